"""The reference's own se_core known-answer tests, re-expressed against the CPU oracle.
Each test names the GTest it restates (paths relative to /root/reference/se_core/test).  These pin the
integer/structural layer of the oracle; nothing here needs a GPU."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib
from oracle_lib import OFUSION, SDF, Oracle

lib = oracle_lib.load()


def enc(x, y, z, level, max_depth):
    return lib.seo_key_encode(x, y, z, level, max_depth)


def decode(code):
    out = (C.c_int * 3)()
    lib.seo_morton_decode(code & ~0x1FF, out)
    return tuple(out)


# ---- utils/morton_unittest.cpp:37-66 --------------------------------------------------------
def test_morton_random_ints():
    rng = np.random.default_rng(7)
    for x, y, z in rng.integers(0, 4097, size=(1000, 3)):
        out = (C.c_int * 3)()
        lib.seo_morton_decode(lib.seo_morton_encode(int(x), int(y), int(z)), out)
        assert tuple(out) == (x, y, z)


def test_morton_exhaustive_slab():
    lib.seo_morton_roundtrip_mismatches.restype = C.c_longlong
    assert lib.seo_morton_roundtrip_mismatches(0, 4096, 2048, 2050, 2048, 4096) == 0


def test_morton_21_bit_corners():
    for v in (0, 1, (1 << 21) - 1):
        out = (C.c_int * 3)()
        lib.seo_morton_decode(lib.seo_morton_encode(v, 0, v), out)
        assert tuple(out) == (v, 0, v)


# ---- octree/octree_unittest.cpp:36-217 ------------------------------------------------------
def test_octant_face_neighbours():
    octant, max_depth, leaves = (112, 80, 160), 8, 5
    code = enc(*octant, leaves, max_depth)
    faces = [(-1, 0, 0), (1, 0, 0), (0, -1, 0), (0, 1, 0), (0, 0, -1), (0, 0, 1)]
    for i, f in enumerate(faces):
        out = (C.c_int * 3)()
        lib.seo_key_face_neighbour(code, i, leaves, max_depth, out)
        assert tuple(out) == tuple(o + 8 * d for o, d in zip(octant, f))


def test_octant_descendant():
    code = enc(110, 80, 159, 5, 8)
    assert lib.seo_key_descendant(code, enc(96, 64, 128, 3, 8), 8) == 1
    assert lib.seo_key_descendant(code, enc(128, 64, 64, 3, 8), 8) == 0


def test_octant_parent_chain():
    code = enc(112, 80, 160, 5, 8)
    p = lib.seo_key_parent(code, 8)
    assert (code & ~0x1FF) == (p & ~0x1FF) and (p & 0x1FF) == 4
    p = lib.seo_key_parent(p, 8)
    assert (p & 0x1FF) == 3 and p == enc(96, 64, 160, 3, 8)
    p = lib.seo_key_parent(p, 8)
    assert (p & 0x1FF) == 2 and p == enc(64, 64, 128, 2, 8)


def test_far_corner():
    cases = {(16, 16, 16): (16, 16, 16), (24, 16, 16): (32, 16, 16), (16, 24, 16): (16, 32, 16),
             (24, 24, 16): (32, 32, 16), (16, 16, 24): (16, 16, 32), (24, 16, 24): (32, 16, 32), (24, 24, 24): (32, 32, 32)}
    for cell, expect in cases.items():
        out = (C.c_int * 3)()
        lib.seo_key_far_corner(enc(*cell, 2, 5), 2, 5, out)
        assert tuple(out) == expect


def test_inner_octant_exterior_neighbours():
    cell = enc(16, 16, 16, 2, 5)
    N = (C.c_uint64 * 7)()
    lib.seo_key_exterior_neighbours(N, cell, 2, 5)
    gt = [enc(15, 16, 16, 2, 5), enc(16, 15, 16, 2, 5), enc(15, 15, 16, 2, 5), enc(16, 16, 15, 2, 5),
          enc(15, 16, 15, 2, 5), enc(16, 15, 15, 2, 5), enc(15, 15, 15, 2, 5)]
    p = lib.seo_key_parent(cell, 5)
    for i in range(7):
        assert N[i] == gt[i]
        assert lib.seo_key_parent(N[i], 5) != p


def test_edge_octant_exterior_neighbours():
    N = (C.c_uint64 * 7)()
    lib.seo_key_exterior_neighbours(N, enc(0, 16, 16, 2, 5), 2, 5)
    for i in range(7):
        assert all(0 <= c <= 31 for c in decode(N[i]))


def test_octant_siblings():
    cell = enc(16, 16, 16, 2, 5)
    s = (C.c_uint64 * 8)()
    lib.seo_key_siblings(s, cell, 5)
    assert s[lib.seo_key_child_id(cell, 2, 5)] == cell
    for i in range(8):
        assert lib.seo_key_parent(s[i], 5) == lib.seo_key_parent(cell, 5)


# ---- allocation/alloc_unittest.cpp:41-123 ---------------------------------------------------
def test_alloc_empty_single_voxel():
    o = Oracle(OFUSION, 256, 5.0, 8, 8)          # OFusion: empty() == initValue() == 0, like the test's float field
    assert o.get(25, 65, 127)[0] == 0.0


def test_alloc_set_single_voxel():
    o = Oracle(SDF, 256, 5.0, 8, 8)
    vox = (25, 65, 127)
    o.allocate([o.hash(*vox)])
    assert o.fetch(*vox)
    o.set_voxel(*vox, 2.0)
    assert o.get(*vox)[0] == 2.0
    assert o.get_fine(*vox)[0] == 2.0


def test_alloc_fetch_octant():
    o = Oracle(SDF, 256, 5.0, 8, 8)
    vox = (25, 65, 127)
    o.allocate([o.hash(*vox)])
    assert o.fetch_octant(*vox, 3)


def test_morton_prefix_mask():
    max_bits, block_side = 21, 8
    size = 1 << max_bits
    rng = np.random.default_rng(3)
    coords = rng.integers(0, size, size=(10, 3))
    leaf_level = max_bits - 3
    shift = max_bits - max_bits
    edge = size // 2
    for level in range(0, leaf_level + 1):
        mask = lib.seo_level_mask(level + shift)
        for x, y, z in coords:
            m = decode(lib.seo_morton_encode(int(x), int(y), int(z)) & mask)
            assert all(c % edge == 0 for c in m)
        edge //= 2


def test_level_mask_table():
    # octree_defines.h:58-80, first/last entries
    assert lib.seo_level_mask(0) == 0x7000000000000000
    assert lib.seo_level_mask(1) == 0x7e00000000000000
    assert lib.seo_level_mask(17) == 0x7ffffffffffffe00
    assert lib.seo_level_mask(20) == 0x7fffffffffffffff


# ---- multiscale/multiscale_unittest.cpp:58-185 ----------------------------------------------
BLOCKS10 = [(56, 12, 254), (87, 32, 423), (128, 128, 128), (136, 128, 128), (128, 136, 128), (136, 136, 128),
            (128, 128, 136), (136, 128, 136), (128, 136, 136), (136, 136, 136)]


def test_multiscale_init():
    o = Oracle(SDF, 512, 5.0, 8, 8)
    assert o.get(137, 138, 130)[0] == 1.0         # initValue


def test_multiscale_plain_alloc():
    o = Oracle(SDF, 512, 5.0, 8, 8)
    o.allocate([o.hash(56, 12, 254), o.hash(87, 32, 423)])
    o.set_voxel(56, 12, 254, 3.0)
    assert o.get(56, 12, 254)[0] == 3.0
    assert o.get(106, 12, 254)[0] == 1.0


def test_multiscale_scaled_alloc():
    o = Oracle(SDF, 512, 5.0, 8, 8)
    o.allocate([o.hash(200, 12, 25, 5), o.hash(87, 32, 423, 5)])
    assert o.fetch_octant(87, 32, 420, 5)
    assert o.set_node_value(87, 32, 420, 5, 0, 10.0)
    assert o.get(87, 32, 420)[0] == 10.0          # coarse get returns the node's value_ where the tree stops


def test_multiscale_iterator_sides():
    o = Oracle(SDF, 512, 5.0, 8, 8)
    o.allocate([o.hash(56, 12, 254)])
    codes, side, mask, values = o.nodes_sorted()
    assert sorted(side.tolist(), reverse=True) == [512, 256, 128, 64, 32, 16]


def test_multiscale_children_mask():
    o = Oracle(SDF, 512, 5.0, 8, 8)
    o.allocate([o.hash(*b, 5) for b in BLOCKS10])
    codes, side, mask, values = o.nodes_sorted()
    # every node's mask bit i <=> a child exists: recount children from the codes
    present = set(int(c) for c in codes)
    for c, s, m in zip(codes, side, mask):
        level = int(c) & 0x1FF
        x, y, z = decode(int(c) & ~0x1FF)
        for i in range(8):
            half = int(s) // 2
            child = lib.seo_key_encode(x + (i & 1) * half, y + ((i >> 1) & 1) * half, z + ((i >> 2) & 1) * half, level + 1, 9)
            has = child in present
            if level + 1 <= 5 and has:
                assert m & (1 << i)


def test_multiscale_octant_alloc():
    o = Oracle(SDF, 512, 5.0, 8, 8)
    keys = [o.hash(*b) for b in BLOCKS10]
    keys[2] = keys[2] | 3
    keys[9] = keys[2] | 5
    o.allocate(keys)
    assert o.fetch_octant(*BLOCKS10[4], 3)
    assert not o.fetch_octant(*BLOCKS10[9], 6)


# ---- algorithms/unique_unittest.cpp:90-118 --------------------------------------------------
def _unique_keys():
    blocks = [(56, 12, 12), (56, 12, 15), (128, 128, 128), (128, 128, 125), (128, 128, 127), (128, 136, 129),
              (128, 136, 127), (136, 128, 136), (128, 240, 136), (128, 241, 136)]
    return np.array([enc(x, y, z, 7, 10) for x, y, z in blocks], np.uint64)


def _multiscale_keys():
    root_side = 2 ** (10 - 4)
    ks = [enc(64, 0, 64, 4, 10), enc(64 + root_side // 2, 0, 64, 5, 10), enc(64 + root_side // 4, 0, 64, 5, 10), enc(128, 24, 80, 5, 10)]
    return np.sort(np.array(ks, np.uint64))


def test_unique_filter_duplicates():
    k = _unique_keys()
    last = lib.seo_keys_unique(k.ctypes.data_as(oracle_lib.u64p), len(k))
    assert all(k[i] != k[i - 1] for i in range(1, last))


def test_filter_ancestors_is_3():
    k = _multiscale_keys()
    last = lib.seo_keys_filter_ancestors(k.ctypes.data_as(oracle_lib.u64p), len(k), 10)
    assert last == 3
    assert all(k[i] != k[i - 1] for i in range(1, last))


def test_unique_multiscale_is_3():
    k = _multiscale_keys()
    last = lib.seo_keys_unique_multiscale(k.ctypes.data_as(oracle_lib.u64p), len(k), 4)
    assert last == 3
    assert all(k[i] != k[i - 1] for i in range(1, last))


# ---- octree/ray_iterator_unittest.cpp:46-87 -------------------------------------------------
def test_ray_iterator_fetch_along_ray():
    o = Oracle(SDF, 512, 5.0, 8, 8)
    p = np.array([1.5, 1.5, 1.5], np.float32)
    d = np.array([0.5, 0.5, 0.5], np.float32)
    d = d / np.sqrt(np.float32((d * d).sum()))
    voxelsize = np.float32(5.0) / np.float32(512)
    stepsize = np.float32(2) * (voxelsize * np.float32(8))
    keys, t = [], np.float32(0.6)
    for _ in range(4):
        vox = ((p + t * d) / voxelsize).astype(np.int32)
        keys.append(o.hash(int(vox[0]), int(vox[1]), int(vox[2])))
        t = t + stepsize
    o.allocate(keys)
    got, tinfo = o.ray_blocks(p, d, 0.4, 4.0)
    assert [int(g) for g in got] == keys


# ---- interp/gather_unittest.cpp:63-187 ------------------------------------------------------
@pytest.mark.parametrize("base,cross", [((136, 128, 136), 0), ((132, 128, 135), 1), ((132, 135, 132), 2), ((132, 135, 135), 3),
                                        ((135, 132, 132), 4), ((135, 132, 135), 5), ((135, 135, 132), 6), ((135, 135, 135), 7)])
def test_gather_all_cross_cases(base, cross):
    o = Oracle(SDF, 512, 5.0, 8, 8)
    o.allocate([o.hash(*b) for b in BLOCKS10])
    assert ((base[0] % 8 == 7) << 2 | (base[1] % 8 == 7) << 1 | (base[2] % 8 == 7)) == cross
    assert np.all(o.gather(*base) == 1.0)         # initValue


def test_gather_reads_the_right_voxels():
    """Stronger than the reference's constant-field test: every voxel gets a distinct value."""
    o = Oracle(SDF, 512, 5.0, 8, 8)
    o.allocate([o.hash(*b) for b in BLOCKS10])
    f = lambda x, y, z: float(x + 1000 * y + 1000000 * (z - 120))
    for x in range(128, 144):
        for y in range(128, 144):
            for z in range(128, 144):
                o.set_voxel(x, y, z, f(x, y, z))
    for base in [(136, 128, 136), (132, 128, 135), (132, 135, 132), (132, 135, 135), (135, 132, 132), (135, 132, 135), (135, 135, 132), (135, 135, 135)]:
        g = o.gather(*base)
        for i in range(8):
            assert g[i] == np.float32(f(base[0] + (i & 1), base[1] + ((i >> 1) & 1), base[2] + ((i >> 2) & 1)))


# ---- multi-level allocation quirk the GPU path reproduces ------------------------------------
def test_allocate_first_key_gets_leaf_chain():
    """Octree::allocate keeps keys[0] in every per-level pass (unique.hpp:63-79 starts at i = 1), so the
    smallest key, when it is not at the leaves level, grows a chain of first children down to a block."""
    o = Oracle(OFUSION, 512, 5.0, 8, 8)
    o.allocate([o.hash(64, 64, 64, 4), o.hash(320, 64, 64, 5)])
    assert o.fetch(64, 64, 64)                     # block at the low corner of the level-4 octant
    assert not o.fetch(320, 64, 64)                # the other (non-first) key stays a childless node
    assert o.fetch_octant(320, 64, 64, 5)
    assert o.block_count() == 1


def test_allocate_zero_keys_is_noop():
    o = Oracle(SDF, 512, 5.0, 8, 8)
    o.allocate([])
    assert o.block_count() == 0 and o.node_count() == 1
