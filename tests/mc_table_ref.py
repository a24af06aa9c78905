"""First-principles generator of the geometry of the marching-cubes case table (N4): what tests/test_meshing.py checks the classic
table the library ships (csrc/se_mc_table.cuh) against -- same directed polygon boundaries in all 256 cases.

Convention (reference numbering, se_core/include/se/algorithms/meshing.hpp:58-104): corner c sits at CORNER[c], edge e joins
EDGE[e] = (source, dest).  On every cube face, walked counter-clockwise as seen from outside the cube, each maximal run of
inside corners is cut off by one segment directed from the edge where the walk enters the run to the edge where it leaves it.
Segments chain into closed polygons; a polygon starts at its lowest edge index, polygons are ordered by that index, and a
polygon (e0 .. ek-1) is the fan (e0, ei, ei+1).  Rows are -1 terminated, 16 wide (at most 5 triangles)."""
import numpy as np

CORNER = [(0, 0, 0), (1, 0, 0), (1, 0, 1), (0, 0, 1), (0, 1, 0), (1, 1, 0), (1, 1, 1), (0, 1, 1)]
EDGE = [(0, 1), (1, 2), (2, 3), (0, 3), (4, 5), (5, 6), (6, 7), (4, 7), (0, 4), (1, 5), (2, 6), (3, 7)]
_CORNER_AT = {p: i for i, p in enumerate(CORNER)}
_EDGE_OF = {}
for _i, (_a, _b) in enumerate(EDGE):
    _EDGE_OF[(_a, _b)] = _i
    _EDGE_OF[(_b, _a)] = _i


def faces():
    """six corner cycles, counter-clockwise seen from outside (right-hand rule about the outward normal)"""
    out = []
    for axis in range(3):
        u, v = (axis + 1) % 3, (axis + 2) % 3
        for side in (0, 1):
            cyc = [(0, 0), (1, 0), (1, 1), (0, 1)] if side else [(0, 0), (0, 1), (1, 1), (1, 0)]
            f = []
            for cu, cv in cyc:
                p = [0, 0, 0]
                p[axis], p[u], p[v] = side, cu, cv
                f.append(_CORNER_AT[tuple(p)])
            out.append(f)
    return out


FACES = faces()


def polygons(index):
    inside = [(index >> c) & 1 for c in range(8)]
    nxt = [-1] * 12
    for cyc in FACES:
        for i in range(4):
            a, b = cyc[i], cyc[(i + 1) % 4]
            if inside[a] or not inside[b]:
                continue
            j = i + 1
            while inside[cyc[(j + 1) % 4]]:
                j += 1
            nxt[_EDGE_OF[(a, b)]] = _EDGE_OF[(cyc[j % 4], cyc[(j + 1) % 4])]
    used = [False] * 12
    loops = []
    for e in range(12):
        if nxt[e] < 0 or used[e]:
            continue
        loop, q = [], e
        while not used[q]:
            used[q] = True
            loop.append(q)
            q = nxt[q]
        loops.append(loop)
    return loops


def triangles(index):
    out = []
    for loop in polygons(index):
        for i in range(1, len(loop) - 1):
            out.append((loop[0], loop[i], loop[i + 1]))
    return out


def table():
    t = np.full((256, 16), -1, np.int8)
    for index in range(256):
        flat = [e for tri in triangles(index) for e in tri]
        assert len(flat) <= 15
        t[index, :len(flat)] = flat
    return t


def classic_table():
    """the table the library ships, read from its data file (no library needed): 256 x 16 int8, -1 terminated rows"""
    import os
    import re
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "supereight_b200", "csrc", "se_mc_table.cuh")
    cases = re.findall(r'"([0-9a-b]*)"', "".join(l for l in open(path) if not l.lstrip().startswith("//")))
    assert len(cases) == 256
    t = np.full((256, 16), -1, np.int8)
    for i, c in enumerate(cases):
        t[i, :len(c)] = [int(ch, 16) for ch in c]
    return t


def boundary(tris):
    """directed boundary edges of a set of triangles (edge-index triples): interior diagonals cancel against their reverse"""
    edges = set()
    for a, b, c in tris:
        for e in ((a, b), (b, c), (c, a)):
            if (e[1], e[0]) in edges:
                edges.remove((e[1], e[0]))
            else:
                edges.add(e)
    return edges
