#!/usr/bin/env python
"""Reads the per-warp phase cycle counters a -DRLG_PHASE_TIMING build of the library dumps at rlg_engine_destroy
(RLG_PHASE_DUMP=path) and prints, per phase, the mean over warps (per role) and the block-critical path
(mean over blocks of the max over the block's warps).  Usage: tools/phase_prof.py gpurun_out/phase_prof.bin"""
import sys

import numpy as np

NAMES = ["load", "s0", "p1", "p2(+p3 self)", "p3", "p4+gym", "reset+store", "barrier wait", "total", "-"]


def main(path):
    raw = np.fromfile(path, dtype=np.uint32)
    blocks, warps, slots, roles = (int(x) for x in raw[:4])
    n1 = blocks * warps * slots
    d = raw[4:4 + n1].reshape(blocks, warps, slots).astype(np.float64)
    sub = raw[4 + n1:].reshape(blocks, warps, 32).astype(np.float64) if raw.size > 4 + n1 else None
    groups = warps // roles
    role = np.arange(warps) % roles
    tot = d[:, :, 8].mean()
    print(f"{blocks} blocks x {warps} warps ({groups} groups x {roles} roles); cycles are sums over all launches of the run")
    print(f"{'phase':<16}{'all warps':>10}{'ball':>10}{'cars':>10}{'max/block':>11}{'max/group':>11}   (percent of the mean warp total)")
    for i in range(8):
        allw = d[:, :, i].mean()
        ball = d[:, role == 0, i].mean()
        cars = d[:, role > 0, i].mean()
        mb = d[:, :, i].max(axis=1).mean()
        mg = d[:, :, i].reshape(blocks, groups, roles).max(axis=2).mean()
        print(f"{NAMES[i]:<16}{100 * allw / tot:>9.1f}%{100 * ball / tot:>9.1f}%{100 * cars / tot:>9.1f}%{100 * mb / tot:>10.1f}%{100 * mg / tot:>10.1f}%")
    if sub is not None:
        SUB = ["car p1: clamp/respawn/inertia + mesh candidates", "car p1: vehicle_first (4 wheel rays, friction impulses)",
               "car p1: wheels/air/jump/flip/auto-roll", "car p1: vehicle_second + boost", "car p1: car-ball", "car p1: hitbox-mesh",
               "car p1: hitbox-plane", "ball p1: pads pre-tick", "ball p1: sphere-mesh", "ball p1: sphere-plane",
               "ball p2: damping, car-car pairs", "ball p2: gather + solve + write back", "ball p2: finish", "car p2: own island solve",
               "car p3: integrate, post-tick, pad overlap",
               "car cands pass: scan + queue", "car cands pass: pair evaluation (rays x4, pre-filter)", "car cands pass: fold",
               "car cands pass: tail / serial lanes", "car hitbox pass: scan + queue", "car hitbox pass: pair evaluation (early-out, GJK/SAT)",
               "car hitbox pass: manifold fold", "car hitbox pass: tail", "* island_one: row setup (both roles)",
               "* island_one: split-impulse iterations", "* island_one: velocity iterations", "* island_one: finish"]
        print("sub-phases (mean over the warps of the role, percent of the mean warp total):")
        for i, nm in enumerate(SUB):
            sel = (role > 0) if nm.startswith("car") else ((role == 0) if nm.startswith("ball") else (role >= 0))
            print(f"  {nm:<58}{100 * sub[:, sel, i].mean() / tot:>7.2f}%")
    bt = d[:, :, 8].max(axis=1)
    print(f"block totals: mean {bt.mean():.0f} max {bt.max():.0f} min {bt.min():.0f} cycles; slowest/mean = {bt.max() / bt.mean():.3f}")


if __name__ == "__main__":
    main(sys.argv[1])
