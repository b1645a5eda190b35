"""CPU suite: the hitbox narrowphase restatements (rl_gjk.h pair detector incl. the penetration-depth search of rl_epa.h, rl_boxbox.h)
against the reference's OWN detectors, called directly on random poses: btGjkPairDetector + btGjkEpaPenetrationDepthSolver for
box-vs-triangle / box-vs-sphere (B/BulletCollision/NarrowPhaseCollision/btGjkPairDetector.cpp:690-1000, btGjkEpa2.cpp) and
btBoxBoxDetector (B/BulletCollision/CollisionDispatch/btBoxBoxDetector.cpp).  Needs oracle/_ref (this container); the device
headers are compiled for the host by tests/hostsim (a test tool)."""
import ctypes as C

import numpy as np
import pytest

from hostsim import hostsim
from oracle import refsim

pytestmark = pytest.mark.skipif(not refsim.available(), reason="oracle/_ref not built")

HALF = np.array([118.01 / 2, 84.2 / 2, 36.16 / 2], dtype=np.float32) / 50.0  # Octane hitbox half extents, Bullet units


def _rot(rng):
    q = rng.standard_normal(4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]], dtype=np.float32)


def _p(a):
    return np.ascontiguousarray(a, dtype=np.float32).ctypes.data_as(C.c_void_p)


def _box_t(rng, center):
    return np.concatenate([np.asarray(center, np.float32), _rot(rng).reshape(-1)]).astype(np.float32)


def test_box_triangle_pair_detector_incl_epa():
    """Random hitbox poses around a large triangle, from separated (inside the breaking distance) to centre-deep overlap: same
    hit / miss decision, normal within 1e-4, contact point within 1e-3 BT (0.05 uu), depth within 1e-5 BT (measured: 6e-6, 1.2e-4, 8e-7).  More than a
    third of the hits take the EPA path (btGjkPairDetector::m_lastUsedMethod == 3)."""
    rng = np.random.default_rng(5)
    R, H = refsim.lib(), hostsim.lib()
    tri = np.array([[-6, -5, 0], [7, -4, 0.3], [0.5, 8, -0.2]], dtype=np.float32)
    n_hit = n_epa = 0
    worst = dict(normal=0.0, point=0.0, depth=0.0)
    for i in range(3000):
        c = np.array([rng.uniform(-3, 3), rng.uniform(-3, 3), rng.uniform(-1.6, 1.6)], dtype=np.float32)
        bt = _box_t(rng, c)
        ro, ho = np.zeros(7, np.float32), np.zeros(7, np.float32)
        method = C.c_int(0)
        nr = R.ref_probe_box_triangle(_p(HALF), _p(bt), _p(tri), C.c_float(0.05), _p(ro), C.byref(method))
        nh = H.hs_probe_box_triangle(_p(HALF), _p(bt), _p(tri), C.c_float(0.05), _p(ho))
        nr, nh = int(nr and ro[6] <= 0.05), int(nh and ho[6] <= 0.05)  # what btManifoldResult::addContactPoint keeps
        assert nr == nh, (i, nr, nh, method.value, ro, ho)
        if nr:
            n_hit += 1
            n_epa += method.value == 3
            worst["normal"] = max(worst["normal"], float(np.abs(ro[:3] - ho[:3]).max()))
            worst["point"] = max(worst["point"], float(np.abs(ro[3:6] - ho[3:6]).max()))
            worst["depth"] = max(worst["depth"], float(abs(ro[6] - ho[6])))
    print("hits", n_hit, "epa", n_epa, worst)
    assert n_hit > 1000 and n_epa > n_hit // 3
    assert worst["normal"] < 1e-4 and worst["point"] < 1e-3 and worst["depth"] < 1e-5, worst


def test_box_sphere_pair_detector():
    """Ball against the hitbox incl. the degenerate-small distance and centre-inside-the-core cases (penetration solver)."""
    rng = np.random.default_rng(6)
    R, H = refsim.lib(), hostsim.lib()
    radius = 91.25 / 50.0
    worst = {False: dict(normal=0.0, point=0.0, depth=0.0), True: dict(normal=0.0, point=0.0, depth=0.0)}  # keyed by "took the EPA path"
    n_hit = n_deep = 0
    for i in range(3000):
        deep = i % 3 == 0
        d = rng.standard_normal(3)
        d /= np.linalg.norm(d)
        dist = rng.uniform(0.0, 0.4) if deep else rng.uniform(0.3, 3.4)
        bt = _box_t(rng, rng.uniform(-2, 2, size=3))
        center = (bt[:3] + d * dist).astype(np.float32)
        ro, ho = np.zeros(7, np.float32), np.zeros(7, np.float32)
        method = C.c_int(0)
        nr = R.ref_probe_box_sphere(_p(HALF), _p(bt), _p(center), C.c_float(radius), C.c_float(0.04), _p(ro), C.byref(method))
        nh = H.hs_probe_box_sphere(_p(HALF), _p(bt), _p(center), C.c_float(radius), C.c_float(0.04), _p(ho))
        # the detector's own gate is (distance - margins)^2 < (margins + breaking)^2; what counts is what btManifoldResult::addContactPoint
        # keeps: depth <= breaking
        nr, nh = int(nr and ro[6] <= 0.04), int(nh and ho[6] <= 0.04)
        assert nr == nh, (i, nr, nh, method.value)
        if nr:
            n_hit += 1
            n_deep += method.value == 3
            w = worst[method.value == 3]
            w["normal"] = max(w["normal"], float(np.abs(ro[:3] - ho[:3]).max()))
            w["point"] = max(w["point"], float(np.abs(ro[3:6] - ho[3:6]).max()))
            w["depth"] = max(w["depth"], float(abs(ro[6] - ho[6])))
    print("hits", n_hit, "epa", n_deep, worst)
    assert n_hit > 1000 and n_deep > 20
    # GJK path (every physically reachable ball contact): closed form vs the iterative detector
    assert worst[False]["normal"] < 1e-4 and worst[False]["point"] < 5e-4 and worst[False]["depth"] < 1e-5, worst
    # ball centre inside the hitbox core (a > 91 uu deep overlap: unreachable in play): EPA on two rounded shapes, 1e-4 accuracy exits
    assert worst[True]["normal"] < 2e-3 and worst[True]["point"] < 1e-2 and worst[True]["depth"] < 4e-4, worst


def test_box_box_detector():
    """btBoxBoxDetector: same number of points, same normal, same points / depths in the same order."""
    rng = np.random.default_rng(7)
    R, H = refsim.lib(), hostsim.lib()
    n_hit = 0
    for i in range(3000):
        ta = _box_t(rng, rng.uniform(-1, 1, size=3))
        d = rng.standard_normal(3)
        d /= np.linalg.norm(d)
        tb = _box_t(rng, ta[:3] + d * rng.uniform(0.3, 2.6))
        ro, ho = np.zeros((8, 7), np.float32), np.zeros((8, 7), np.float32)
        nr = R.ref_probe_box_box(_p(HALF), _p(ta), _p(HALF), _p(tb), _p(ro), 8)
        nh = H.hs_probe_box_box(_p(HALF), _p(ta), _p(HALF), _p(tb), _p(ho), 8)
        assert nr == nh, (i, nr, nh)
        if nr:
            n_hit += 1
            assert np.abs(ro[:nr] - ho[:nr]).max() < 2e-4, (i, ro[:nr], ho[:nr])
    print("hits", n_hit)
    assert n_hit > 800
