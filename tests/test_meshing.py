"""N4: marching cubes (se_core/include/se/algorithms/meshing.hpp:158-208, DenseSLAMSystem::dump_mesh).

CPU part: the classic case table the library ships (csrc/se_mc_table.cuh) has, case by case, the directed polygon boundaries a
first-principles generator derives (tests/mc_table_ref.py) and the properties a marching-cubes table must have -- and, where the
reference tree is present, is the reference's own list; the oracle's restatement on closed forms.
GPU part: the CUDA mesh == the oracle's mesh, bit for bit, triangle by triangle."""
import itertools

import numpy as np
import pytest

import mc_table_ref as ref
from oracle_lib import OFUSION, SDF, Oracle
from supereight_b200 import synth

K = (481.2, 480.0, 320.0, 240.0)


def rows(table):
    return [[tuple(int(v) for v in table[i, j:j + 3]) for j in range(0, 15, 3) if table[i, j] != -1] for i in range(256)]


# ---- the table ------------------------------------------------------------------------------------
def test_library_table_is_the_classic_table_and_has_the_generated_geometry():
    import supereight_b200
    lib = supereight_b200.mc_table()
    assert lib.shape == (256, 16) and lib.dtype == np.int8
    assert np.array_equal(lib, ref.classic_table())                  # the C++ decoder of the data file == this one
    gen = rows(ref.table())
    for i, tris in enumerate(rows(lib)):
        assert len(tris) == len(gen[i]), i                           # same number of triangles ...
        assert ref.boundary(tris) == ref.boundary(gen[i]), i         # ... filling the same directed polygons
    assert sum(len(t) for t in rows(lib)) == 820


def test_library_table_is_the_reference_list():
    """edge_tables.h (`triTable`), where the reference tree exists (the development container)"""
    import os
    import sys
    if not os.path.exists("/root/reference/se_core/include/se/algorithms/edge_tables.h"):
        pytest.skip("no reference tree on this machine")
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
    import make_mc_table
    want = make_mc_table.read_reference_rows("/root/reference")
    got = [[int(v) for v in row if v != -1] for row in ref.classic_table()]
    assert got == want


def test_table_rows_are_well_formed():
    t = ref.classic_table()
    for i in range(256):
        row = t[i]
        n = int(np.argmax(row == -1)) if (row == -1).any() else 16
        assert n % 3 == 0 and n <= 15 and (row[n:] == -1).all() and (row[:n] >= 0).all() and (row[:n] < 12).all()
    assert rows(t)[0] == [] and rows(t)[255] == []
    assert rows(t)[1] == [(0, 8, 3)]                      # corner 0 inside: the triangle around it, facing away from it


def test_table_uses_exactly_the_edges_that_change_sign():
    for i, tris in enumerate(rows(ref.classic_table())):
        cut = {e for e, (a, b) in enumerate(ref.EDGE) if ((i >> a) & 1) != ((i >> b) & 1)}
        assert {e for tri in tris for e in tri} == cut, i


def directed_boundary(tris):
    d = set()
    for t in tris:
        for k in range(3):
            a, b = t[k], t[(k + 1) % 3]
            if (b, a) in d:
                d.remove((b, a))
            else:
                assert (a, b) not in d
                d.add((a, b))
    return d


def test_table_polygons_stay_on_cube_faces_and_neighbours_agree():
    """every boundary segment of a cell's surface lies on one cube face, and the cell across that face produces the same
    segment reversed: the mesh has no cracks and a consistent orientation"""
    all_rows = rows(ref.classic_table())
    edge_faces = []
    for a, b in ref.EDGE:
        pa, pb = ref.CORNER[a], ref.CORNER[b]
        edge_faces.append({(ax, pa[ax]) for ax in range(3) if pa[ax] == pb[ax]})
    seg = []
    for i in range(256):
        per_face = {}
        for a, b in directed_boundary(all_rows[i]):
            common = edge_faces[a] & edge_faces[b]
            assert len(common) == 1, (i, a, b)
            per_face.setdefault(next(iter(common)), set()).add((a, b))
        seg.append(per_face)

    def shifted(edge, axis):                      # the same geometric edge seen from the neighbour across the +axis face
        a, b = ref.EDGE[edge]
        pa, pb = list(ref.CORNER[a]), list(ref.CORNER[b])
        pa[axis] = 0; pb[axis] = 0
        return ref._EDGE_OF[(ref._CORNER_AT[tuple(pa)], ref._CORNER_AT[tuple(pb)])]

    for axis in range(3):
        hi = [c for c in range(8) if ref.CORNER[c][axis] == 1]
        lo = [ref._CORNER_AT[tuple(0 if ax == axis else v for ax, v in enumerate(ref.CORNER[c]))] for c in hi]
        for i, j in itertools.product(range(256), range(256)):
            if any(((i >> h) & 1) != ((j >> l) & 1) for h, l in zip(hi, lo)):
                continue                          # j is not a possible +axis neighbour of i
            mine = {(shifted(a, axis), shifted(b, axis)) for a, b in seg[i].get((axis, 1), set())}
            theirs = {(b, a) for a, b in seg[j].get((axis, 0), set())}
            assert mine == theirs, (axis, i, j)


def test_table_orientation_single_corner_cases():
    mid = [np.mean([ref.CORNER[a], ref.CORNER[b]], axis=0) for a, b in ref.EDGE]
    all_rows = rows(ref.classic_table())
    for c in range(8):
        for index, sign in ((1 << c, 1.0), (255 ^ (1 << c), -1.0)):
            (t,) = all_rows[index]
            n = np.cross(mid[t[1]] - mid[t[0]], mid[t[2]] - mid[t[1]])
            away = np.mean([mid[e] for e in t], axis=0) - np.array(ref.CORNER[c], float)
            assert sign * float(n @ away) > 0      # normals point from inside (x < 0) to outside


# ---- the oracle's marching_cube on closed forms ------------------------------------------------------
def plane_map(field, frames=3, size=256, dim=4.8, W=160, H=120, mu=0.1):
    k = tuple(v * W / 640.0 for v in K)
    o = Oracle(field, size, dim, W, H)
    for f in range(frames):
        d, pose = synth.planar_sweep(0, dim, W, H, k, dropout=0.0)
        o.preprocess(d)
        o.integrate(pose, k, mu, f)
    return o, pose, k


def test_oracle_mesh_of_a_wall_lies_on_the_wall_and_faces_the_camera():
    dim, size = 4.8, 256
    o, pose, k = plane_map(SDF)
    tri = o.marching_cube(ref.classic_table())
    assert len(tri) > 1000
    zw = 0.75 * dim
    assert np.abs(tri[..., 2] - zw).max() < 1.0 * dim / size          # the zero crossing is the wall (1 mm depth quantisation)
    n = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 1])
    assert (n[:, 2] < 0).mean() > 0.999                                 # free space (x > 0) is on the camera side, -z
    assert (tri > 0).all() and (tri <= dim).all()                       # checkVertex


def test_oracle_mesh_is_watertight_inside_the_observed_region():
    """inner edges of the wall's mesh are shared by exactly two triangles, in opposite directions"""
    o, pose, k = plane_map(SDF)
    tri = o.marching_cube(ref.classic_table())
    keys = np.round(tri.astype(np.float64) * 1e6).astype(np.int64)
    edges = {}
    for t in keys:
        for a in range(3):
            e = (tuple(t[a]), tuple(t[(a + 1) % 3]))
            edges[e] = edges.get(e, 0) + 1
    assert max(edges.values()) == 1                                     # no directed edge twice
    paired = sum(1 for (a, b) in edges if (b, a) in edges)
    assert paired / len(edges) > 0.97                                   # the rest is the rim of the observed patch


def test_oracle_empty_map_has_no_mesh():
    assert len(Oracle(SDF, 64, 1.0, 8, 8).marching_cube(ref.classic_table())) == 0


def test_oracle_mesh_hand_built_cell():
    """one cell with corner 0 inside: the triangle of case 1 at the linear zero crossings (meshing.hpp:46-55)"""
    o = Oracle(SDF, 64, 6.4, 8, 8)                                       # voxel = 0.1 m
    o.allocate([o.hash(8, 8, 8)])
    for x, y, z in itertools.product(range(8, 11), repeat=3):
        o.set_voxel(x, y, z, 0.5, 1.0)
    o.set_voxel(9, 9, 9, -0.25, 1.0)
    tri = o.marching_cube(ref.classic_table())
    # the inside voxel is corner 0 of cell (9,9,9) and some other corner of its 7 lower neighbours: 8 triangles, an octahedron
    assert tri.shape == (8, 3, 3)
    cell = [t for t in tri if (t >= np.float32(0.9) - 1e-6).all()]
    assert len(cell) == 1
    third = np.float32(0.9) + (np.float32(0.25) * np.float32(0.1)) / np.float32(0.75)
    expect = np.array([[third, 0.9, 0.9], [0.9, third, 0.9], [0.9, 0.9, third]], np.float32)   # edges 0, 8, 3
    assert np.allclose(cell[0], expect, rtol=0, atol=2e-7)


# ---- GPU parity ----------------------------------------------------------------------------------------
def _gpu_map(field, size, dim, W, H):
    from supereight_b200 import Map
    return Map(field, size, dim, W, H)


@pytest.mark.gpu
def test_gpu_mesh_equals_oracle_sdf_sequence():
    size, dim, W, H, mu = 256, 4.8, 160, 120, 0.1
    k = tuple(v * W / 640.0 for v in K)
    g, o = _gpu_map(SDF, size, dim, W, H), Oracle(SDF, size, dim, W, H)
    for f in range(6):
        d, pose = synth.box_room(f * 7, dim, W, H, k, noise_mm=2.0, dropout=0.01)
        g.preprocess(d); o.preprocess(d)
        g.integrate(pose, k, mu, f); o.integrate(pose, k, mu, f)
    want = o.marching_cube(ref.classic_table())
    got = g.mesh()
    assert len(want) > 5000
    assert got.shape == want.shape
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))       # same triangles, same order, same bits
    assert np.array_equal(g.mesh().view(np.uint32), got.view(np.uint32))   # idempotent (buffers are reused)


@pytest.mark.gpu
@pytest.mark.parametrize("field", [SDF, OFUSION])
def test_gpu_mesh_of_an_uploaded_map_equals_oracle(field):
    """the map built by the oracle, moved to the device through the map-file records (N3), meshes identically: covers
    OFusion, whose integration is only parity-within-tolerance, and blocks on the volume border"""
    size, dim, W, H, mu = 128, 2.4, 80, 60, 0.1
    k = tuple(v * W / 640.0 for v in K)
    o = Oracle(field, size, dim, W, H)
    for f in range(4):
        d, pose = synth.box_room(f * 11, dim, W, H, k, noise_mm=1.0, dropout=0.0)     # walls at 0.1 / 0.9 dim
        o.preprocess(d)
        o.integrate(pose, k, mu, f)
    keys, coords, active, data = o.blocks_sorted()
    codes, side, mask, values = o.nodes_sorted()
    g = _gpu_map(field, size, dim, W, H)
    g.upload_nodes(codes, values)
    g.upload_blocks(keys, data)
    want = o.marching_cube(ref.classic_table())
    got = g.mesh()
    assert len(want) > 100
    assert got.shape == want.shape
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.gpu
def test_gpu_mesh_empty_map_and_border_clamp():
    g = _gpu_map(SDF, 64, 6.4, 8, 8)
    assert g.mesh().shape == (0, 3, 3)
    # a surface through the last voxel layer: cells starting at size-1 do not exist (meshing.hpp:178-180), vertices on the
    # lower faces (coordinate 0) are dropped by checkVertex (:151-153)
    o = Oracle(SDF, 64, 6.4, 8, 8)
    keys = [o.hash(56, 56, 56), o.hash(0, 0, 0)]
    o.allocate(keys); g.allocate(keys)
    rng = np.random.default_rng(5)
    xyz = np.array([(x, y, z) for b in (0, 56) for x in range(b, b + 8) for y in range(b, b + 8) for z in range(b, b + 8)], np.int32)
    vox = np.zeros(len(xyz), o.vdtype)
    vox["x"] = rng.uniform(-1, 1, len(xyz)).astype(np.float32)
    vox["y"] = (rng.random(len(xyz)) > 0.05).astype(np.float32)
    for p, v in zip(xyz, vox):
        o.set_voxel(int(p[0]), int(p[1]), int(p[2]), float(v["x"]), float(v["y"]))
    g.set_voxels(xyz, vox)
    want = o.marching_cube(ref.classic_table())
    got = g.mesh()
    assert len(want) > 200
    assert got.shape == want.shape and np.array_equal(got.view(np.uint32), want.view(np.uint32))
