"""Placeholder soccar collision-mesh set (v1).

The real ``collision_meshes/soccar/*.cmf`` files are dumped from the game and are not
redistributable (reference README.md:41), so BOTH sides of every parity test and of the
benchmark load the byte-identical set generated here (BASELINE.json north_star: "both
sides use the same generated placeholder soccar mesh set").

File format (reference RocketSim/src/CollisionMeshFile/CollisionMeshFile.cpp:11-35):
``int32 numTris, int32 numVerts, numTris x int32[3], numVerts x float32[3]`` little
endian, vertices in Bullet units (uu / 50).

What the set contains — everything soccar does NOT get from its four built-in planes
(floor z=0, ceiling z=2048, side walls x=+-4096; reference Arena.cpp:1060-1101):
  0. orange end: back wall y=+5120 with goal mouth, goal box to y=+6000
  1. blue end:   same, mirrored
  2. the four 45-degree corner walls
  3. 45-degree floor ramps along side walls and back walls, ceiling ramps along side walls
All quads are split into cells no larger than ~768 uu so the BVH is not trivial.
Triangles are wound so that the geometric normal points into the playing field.
"""
from __future__ import annotations

import struct
from typing import List

import numpy as np

UU_TO_BT = 1.0 / 50.0

EXTENT_X = 4096.0
EXTENT_Y = 5120.0
HEIGHT = 2048.0
GOAL_HALF_W = 892.755
GOAL_H = 642.775
GOAL_BACK_Y = 6000.0
CORNER = 1152.0
RAMP = 256.0
MAX_CELL = 768.0


class _MeshBuilder:
    def __init__(self):
        self.verts: List[tuple] = []
        self.tris: List[tuple] = []

    def quad(self, p00, p10, p11, p01, inward):
        """Bilinear quad p00-p10-p11-p01 subdivided into <=MAX_CELL cells; winding chosen so
        the normal has a positive dot with `inward`."""
        p00, p10, p11, p01 = (np.asarray(p, dtype=np.float64) for p in (p00, p10, p11, p01))
        nu = max(1, int(np.ceil(np.linalg.norm(p10 - p00) / MAX_CELL)))
        nv = max(1, int(np.ceil(np.linalg.norm(p01 - p00) / MAX_CELL)))
        n = np.cross(p10 - p00, p01 - p00)
        flip = float(np.dot(n, np.asarray(inward, dtype=np.float64))) < 0
        base = len(self.verts)
        for j in range(nv + 1):
            for i in range(nu + 1):
                u, v = i / nu, j / nv
                p = (1 - u) * (1 - v) * p00 + u * (1 - v) * p10 + u * v * p11 + (1 - u) * v * p01
                self.verts.append(tuple(float(x) for x in p))
        for j in range(nv):
            for i in range(nu):
                a = base + j * (nu + 1) + i
                b = a + 1
                c = a + (nu + 1) + 1
                d = a + (nu + 1)
                t1, t2 = (a, b, c), (a, c, d)
                if flip:
                    t1, t2 = (a, c, b), (a, d, c)
                self.tris.append(t1)
                self.tris.append(t2)

    def to_cmf(self) -> bytes:
        out = [struct.pack("<ii", len(self.tris), len(self.verts))]
        for t in self.tris:
            out.append(struct.pack("<iii", *t))
        for v in self.verts:
            out.append(struct.pack("<fff", *(np.float32(c * UU_TO_BT) for c in v)))
        return b"".join(out)


def _end_wall(sign: float) -> _MeshBuilder:
    """Back wall at y = sign*5120 with goal mouth + goal box."""
    m = _MeshBuilder()
    y = sign * EXTENT_Y
    yb = sign * GOAL_BACK_Y
    inward = (0, -sign, 0)
    xl = EXTENT_X - CORNER  # back wall spans |x| <= 2944; beyond that the corner walls take over
    # left and right of the goal mouth
    m.quad((-xl, y, 0), (-GOAL_HALF_W, y, 0), (-GOAL_HALF_W, y, HEIGHT), (-xl, y, HEIGHT), inward)
    m.quad((GOAL_HALF_W, y, 0), (xl, y, 0), (xl, y, HEIGHT), (GOAL_HALF_W, y, HEIGHT), inward)
    # above the goal mouth
    m.quad((-GOAL_HALF_W, y, GOAL_H), (GOAL_HALF_W, y, GOAL_H), (GOAL_HALF_W, y, HEIGHT), (-GOAL_HALF_W, y, HEIGHT), inward)
    # goal box: two side walls, back net, roof
    m.quad((-GOAL_HALF_W, y, 0), (-GOAL_HALF_W, yb, 0), (-GOAL_HALF_W, yb, GOAL_H), (-GOAL_HALF_W, y, GOAL_H), (1, 0, 0))
    m.quad((GOAL_HALF_W, y, 0), (GOAL_HALF_W, yb, 0), (GOAL_HALF_W, yb, GOAL_H), (GOAL_HALF_W, y, GOAL_H), (-1, 0, 0))
    m.quad((-GOAL_HALF_W, yb, 0), (GOAL_HALF_W, yb, 0), (GOAL_HALF_W, yb, GOAL_H), (-GOAL_HALF_W, yb, GOAL_H), inward)
    m.quad((-GOAL_HALF_W, y, GOAL_H), (GOAL_HALF_W, y, GOAL_H), (GOAL_HALF_W, yb, GOAL_H), (-GOAL_HALF_W, yb, GOAL_H), (0, 0, -1))
    return m


def _corners() -> _MeshBuilder:
    m = _MeshBuilder()
    for sx in (-1.0, 1.0):
        for sy in (-1.0, 1.0):
            a = (sx * (EXTENT_X - CORNER), sy * EXTENT_Y)
            b = (sx * EXTENT_X, sy * (EXTENT_Y - CORNER))
            m.quad((a[0], a[1], 0), (b[0], b[1], 0), (b[0], b[1], HEIGHT), (a[0], a[1], HEIGHT), (-sx, -sy, 0))
    return m


def _ramps() -> _MeshBuilder:
    m = _MeshBuilder()
    yl = EXTENT_Y - CORNER
    xl = EXTENT_X - CORNER
    for sx in (-1.0, 1.0):
        x0, x1 = sx * (EXTENT_X - RAMP), sx * EXTENT_X
        # floor ramp along the side wall
        m.quad((x0, -yl, 0), (x0, yl, 0), (x1, yl, RAMP), (x1, -yl, RAMP), (-sx, 0, 1))
        # ceiling ramp along the side wall
        m.quad((x0, -yl, HEIGHT), (x0, yl, HEIGHT), (x1, yl, HEIGHT - RAMP), (x1, -yl, HEIGHT - RAMP), (-sx, 0, -1))
    for sy in (-1.0, 1.0):
        y0, y1 = sy * (EXTENT_Y - RAMP), sy * EXTENT_Y
        # floor ramps along the back wall, either side of the goal mouth
        m.quad((-xl, y0, 0), (-GOAL_HALF_W, y0, 0), (-GOAL_HALF_W, y1, RAMP), (-xl, y1, RAMP), (0, -sy, 1))
        m.quad((GOAL_HALF_W, y0, 0), (xl, y0, 0), (xl, y1, RAMP), (GOAL_HALF_W, y1, RAMP), (0, -sy, 1))
    return m


def generate_placeholder_soccar() -> List[bytes]:
    """Returns the placeholder set as a list of .cmf blobs, in load order."""
    return [_end_wall(+1.0).to_cmf(), _end_wall(-1.0).to_cmf(), _corners().to_cmf(), _ramps().to_cmf()]


def parse_cmf(blob: bytes):
    """-> (tris int32 [T,3], verts float32 [V,3] in Bullet units)"""
    nt, nv = struct.unpack_from("<ii", blob, 0)
    tris = np.frombuffer(blob, dtype="<i4", count=nt * 3, offset=8).reshape(nt, 3).copy()
    verts = np.frombuffer(blob, dtype="<f4", count=nv * 3, offset=8 + nt * 12).reshape(nv, 3).copy()
    return tris, verts


def read_cmf_folder(folder: str, game_mode: str = "soccar") -> List[bytes]:
    """The .cmf files RocketSim::Init(folder) loads for a game mode (R/RocketSim.cpp:70-212: <folder>/<mode>/*.cmf), as blobs in
    directory order, e.g. the real arena meshes dumped with the reference's asset tool."""
    import os

    d = os.path.join(folder, game_mode)
    if not os.path.isdir(d):
        raise RuntimeError(f"RocketSim::Init: collision mesh folder {d} does not exist")
    names = sorted(n for n in os.listdir(d) if n.lower().endswith(".cmf"))
    if not names:
        raise RuntimeError(f"RocketSim::Init: no .cmf files in {d}")
    blobs = []
    for n in names:
        with open(os.path.join(d, n), "rb") as f:
            blobs.append(f.read())
    return blobs


def write_placeholder_set(folder: str) -> List[str]:
    """Writes the set as <folder>/soccar/placeholder_<i>.cmf (what RocketSim::Init reads)."""
    import os

    d = os.path.join(folder, "soccar")
    os.makedirs(d, exist_ok=True)
    paths = []
    for i, blob in enumerate(generate_placeholder_soccar()):
        p = os.path.join(d, f"placeholder_{i}.cmf")
        with open(p, "wb") as f:
            f.write(blob)
        paths.append(p)
    return paths


if __name__ == "__main__":
    import argparse

    ap = argparse.ArgumentParser(description="write the placeholder soccar set as <out>/soccar/*.cmf (what RocketSim::Init reads)")
    ap.add_argument("--out", default="collision_meshes")
    for path in write_placeholder_set(ap.parse_args().out):
        print(path)
