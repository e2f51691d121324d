"""Generate the marching-cubes triangulation table used by K10 (TSDF -> triangle mesh).

No marching-cubes table exists in the reference (it has no mesh path) or in any package
of this image, so the table is *derived*, not transcribed:

* corner / edge numbering and the per-edge owner shifts follow Open3D's
  `VoxelBlockGrid.extract_triangle_mesh` convention (corner i of the cube based at voxel
  (x,y,z): 0:(0,0,0) 1:(1,0,0) 2:(1,1,0) 3:(0,1,0) 4:(0,0,1) 5:(1,0,1) 6:(1,1,1) 7:(0,1,1);
  edge e joins EDGE_CORNERS[e]; the vertex on edge e is owned by voxel base+shift, axis a);
* case index bit i is set iff tsdf(corner i) < 0;
* on every cube face the crossed edges are joined by one segment per maximal run of negative
  corners (so on an ambiguous face the two negative corners are cut off separately — a rule that
  depends on the four corner signs of the face only, hence identical for the two cubes that share
  the face ⇒ the mesh is watertight across cubes);
* the directed segments chain into closed loops, each loop is fan-triangulated from its
  smallest edge id; winding is chosen so triangle normals point to the positive (free-space)
  side, the direction of the TSDF gradient.

Vertex SET of the mesh is therefore exactly Open3D's (one vertex per sign-changing edge of a
valid cube); triangle connectivity inside ambiguous cubes may differ from Open3D's transcribed
Lorensen/Bourke table (parity unpinned).

Writes the same numbers to  oracle/mc_tables.h  (C, for the oracle) and
textureless_3d_reconstruction_b200/csrc/mc_tables.cuh  (CUDA __constant__).
Run:  python oracle/gen_mc_tables.py
"""
from __future__ import annotations

from pathlib import Path

import numpy as np

CORNERS = np.array([(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)])
EDGE_CORNERS = [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7)]
# faces as corner cycles, counter-clockwise seen from outside the cube
FACES = [(0, 3, 2, 1), (4, 5, 6, 7), (0, 1, 5, 4), (3, 7, 6, 2), (0, 4, 7, 3), (1, 2, 6, 5)]


def edge_shift(e):
    a, b = EDGE_CORNERS[e]
    pa, pb = CORNERS[a], CORNERS[b]
    lo = np.minimum(pa, pb)
    axis = int(np.nonzero(pa != pb)[0][0])
    return (int(lo[0]), int(lo[1]), int(lo[2]), axis)


def edge_id(a, b):
    for e, (p, q) in enumerate(EDGE_CORNERS):
        if (p, q) == (a, b) or (p, q) == (b, a):
            return e
    raise KeyError((a, b))


def check_faces():
    for f in FACES:
        p = CORNERS[list(f)].astype(float)
        n = np.cross(p[1] - p[0], p[2] - p[1])
        out = p.mean(0) - 0.5
        assert np.dot(n, out) > 0, f


def loops_of_case(case):
    neg = [(case >> i) & 1 for i in range(8)]
    nxt = {}
    for f in FACES:
        # maximal runs of negative corners on the face cycle
        for k in range(4):
            a, b = f[k], f[(k + 1) % 4]
            if not neg[a] and neg[b]:  # enter a negative run at edge (a,b)
                j = (k + 1) % 4
                while neg[f[(j + 1) % 4]]:
                    j = (j + 1) % 4
                e_in = edge_id(a, b)
                e_out = edge_id(f[j], f[(j + 1) % 4])
                assert e_in not in nxt
                nxt[e_in] = e_out
    loops = []
    seen = set()
    for e0 in sorted(nxt):
        if e0 in seen:
            continue
        loop = [e0]
        seen.add(e0)
        e = nxt[e0]
        while e != e0:
            loop.append(e)
            seen.add(e)
            e = nxt[e]
        loops.append(loop)
    crossed = {e for e, (a, b) in enumerate(EDGE_CORNERS) if neg[a] != neg[b]}
    assert seen == crossed, (case, seen, crossed)
    return loops


def midpoint(e):
    a, b = EDGE_CORNERS[e]
    return 0.5 * (CORNERS[a] + CORNERS[b])


def build():
    check_faces()
    tris = []
    for case in range(256):
        t = []
        for loop in loops_of_case(case):
            for k in range(1, len(loop) - 1):
                t.append((loop[0], loop[k], loop[k + 1]))
        tris.append(t)
    # orientation: one global flip decision, taken on the single-corner case and verified on all
    # cases whose negative corners form a set with a well-defined direction
    def score(case, t):
        neg = np.array([(case >> i) & 1 for i in range(8)], bool)
        d = CORNERS[~neg].mean(0) - CORNERS[neg].mean(0)
        s = 0.0
        for (a, b, c) in t:
            n = np.cross(midpoint(b) - midpoint(a), midpoint(c) - midpoint(a))
            s += float(np.dot(n, d))
        return s
    flip = score(1, tris[1]) < 0
    if flip:
        tris = [[(a, c, b) for (a, b, c) in t] for t in tris]
    for case in range(1, 255):
        assert score(case, tris[case]) >= -1e-12, case
    # complement symmetry of the edge set, manifoldness inside a cube: every directed edge once
    for case in range(256):
        used = set()
        for (a, b, c) in tris[case]:
            for d in ((a, b), (b, c), (c, a)):
                assert d not in used
                used.add(d)
    return tris


def emit(tris):
    maxt = max(len(t) for t in tris)
    ntri = [len(t) for t in tris]
    flat = []
    for t in tris:
        row = [e for tri in t for e in tri]
        row += [-1] * (maxt * 3 - len(row))
        flat.append(row)
    shifts = [edge_shift(e) for e in range(12)]
    head = ("// GENERATED by oracle/gen_mc_tables.py — do not edit.  Marching-cubes tables for K10\n"
            "// (corner/edge numbering and owner shifts as Open3D's VoxelBlockGrid mesh extraction;\n"
            "// triangulation derived by face-loop tracing, see the generator's docstring).\n")
    def body(q):
        s = f"#define MC_MAX_TRI {maxt}\n"
        s += f"{q} unsigned char MC_NUM_TRI[256] = {{\n"
        for r in range(0, 256, 32):
            s += "  " + ", ".join(str(v) for v in ntri[r:r + 32]) + ",\n"
        s += "};\n"
        s += f"{q} signed char MC_TRI[256][{maxt * 3}] = {{\n"
        for row in flat:
            s += "  {" + ", ".join(f"{v}" for v in row) + "},\n"
        s += "};\n"
        s += f"{q} signed char MC_EDGE_SHIFT[12][4] = {{\n"
        for sh in shifts:
            s += "  {" + ", ".join(str(v) for v in sh) + "},\n"
        s += "};\n"
        s += f"{q} signed char MC_CORNER[8][3] = {{\n"
        for c in CORNERS:
            s += "  {" + ", ".join(str(int(v)) for v in c) + "},\n"
        s += "};\n"
        return s
    root = Path(__file__).resolve().parent.parent
    (root / "oracle" / "mc_tables.h").write_text(head + "#pragma once\n" + body("static const"))
    (root / "textureless_3d_reconstruction_b200" / "csrc" / "mc_tables.cuh").write_text(
        head + "#pragma once\n" + body("__constant__"))
    return maxt, sum(ntri)


if __name__ == "__main__":
    t = build()
    maxt, total = emit(t)
    print(f"max triangles per case {maxt}, total {total}")
