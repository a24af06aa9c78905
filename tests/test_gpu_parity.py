"""GPU parity tests proper: the CUDA path, called through the C ABI (include/se_b200.h), against the CPU
oracle on the same seeded inputs.

Bars (north_star): block allocation set bit-exact; TSDF / vertex / normal within 1e-4 relative.  Because the
library follows the oracle's arithmetic contract (no FMA, fixed summation order), the SDF path is in fact
required to be BIT-EXACT here; OFusion goes through log2 (libm vs device, <= 1 ulp apart) and is held to the
1e-4 tolerance, written out below."""
import numpy as np
import pytest

import oracle_lib
from oracle_lib import OFUSION, SDF, Oracle
from parity_utils import compare_blocks, compare_images, compare_nodes

pytestmark = pytest.mark.gpu

REL_TOL = 1e-4          # north_star: "TSDF values and raycast vertex/normal maps within 1e-4 relative"
K640 = (481.2, 480.0, 320.0, 240.0)


def scaled_k(W):
    return tuple(v * W / 640.0 for v in K640)


def make_pair(field, size, dim, W, H, **kw):
    from supereight_b200 import Map
    return Map(field, size, dim, W, H, **kw), Oracle(field, size, dim, W, H)


def run_sequence(g, o, gen, dim, W, H, k, mu, frames, **genkw):
    pose = None
    for f in frames:
        d, pose = gen(f, dim, W, H, k, **genkw)
        o.preprocess(d); o.integrate(pose, k, mu, f)
        g.preprocess(d); g.integrate(pose, k, mu, f)
    return pose


def assert_sdf_bit_exact(g, o, pose, k, mu):
    cb = compare_blocks(g, o)
    assert cb["keys_equal"], cb
    assert cb["coords_equal"] and cb["active_mismatch"] == 0, cb
    assert cb["x_bit_mismatch"] == 0 and cb["y_mismatch"] == 0, cb
    assert cb["x_max_rel"] <= REL_TOL
    cn = compare_nodes(g, o)
    assert cn["codes_equal"] and cn["side_equal"] and cn["mask_equal"], cn
    assert cn["x_bit_mismatch"] == 0 and cn["y_mismatch"] == 0, cn
    o.raycast(pose, k, mu); g.raycast(pose, k, mu)
    gv, gn = g.vertex_normal()
    ci = compare_images(gv, gn, o.vertex(), o.normal())
    assert ci["hit_mask_mismatch"] == 0, ci
    assert ci["vertex_bit_mismatch"] == 0 and ci["normal_bit_mismatch"] == 0, ci
    assert ci["hits_gpu"] > 0
    return cb, ci


# ---- SDF, the metric's configuration: 640x480 into 512^3 ------------------------------------
def test_sdf_512_full_frame_sequence_bit_exact():
    from supereight_b200 import synth
    dim, mu, W, H = 4.8, 0.1, 640, 480
    g, o = make_pair(SDF, 512, dim, W, H)
    pose = run_sequence(g, o, synth.planar_sweep, dim, W, H, K640, mu, range(4), noise_mm=2.0, dropout=0.01)
    cb, ci = assert_sdf_bit_exact(g, o, pose, K640, mu)
    assert cb["n_gpu"] > 5000 and ci["hits_gpu"] > 250000
    # renderVolume: reuse path and the re-raycast path from another view, renderDepth: byte-exact
    assert np.array_equal(g.render_volume(pose, K640, mu, 0.75 * mu, False), o.render_volume(pose, K640, mu, 0.75 * mu, False))
    view = synth.planar_sweep_pose(40, dim)
    assert np.array_equal(g.render_volume(view, K640, mu, 0.75 * mu, True), o.render_volume(view, K640, mu, 0.75 * mu, True))
    assert np.array_equal(g.render_depth(), o.render_depth())


def test_sdf_8mm_voxels_room_bit_exact():
    """north_star's '512^3 8 mm' reading: dim 4.096 m; box room scene turns the camera (yaw) between frames."""
    from supereight_b200 import synth
    dim, mu, W, H = 4.096, 0.1, 320, 240
    k = scaled_k(W)
    g, o = make_pair(SDF, 512, dim, W, H)
    pose = run_sequence(g, o, synth.box_room, dim, W, H, k, mu, range(0, 30, 6), n_frames=60, dropout=0.02)
    assert_sdf_bit_exact(g, o, pose, k, mu)


def test_sdf_negative_fy_camera():
    """ICL-NUIM intrinsics have fy < 0 (README.md:80): image flipped vertically."""
    from supereight_b200 import synth
    dim, mu, W, H = 4.8, 0.1, 160, 120
    k = (120.3, -120.0, 80.0, 60.0)
    g, o = make_pair(SDF, 256, dim, W, H)
    pose = run_sequence(g, o, synth.planar_sweep, dim, W, H, k, mu, range(3), dropout=0.0)
    assert_sdf_bit_exact(g, o, pose, k, mu)


def test_sdf_ratio2_preprocess_and_empty_frames():
    """compute-size-ratio 2 (320x240 from 640x480 input), an all-zero depth frame (nothing allocated, nothing
    fused), then a normal frame."""
    from supereight_b200 import synth
    dim, mu, W, H = 4.8, 0.1, 320, 240
    k = scaled_k(W)
    g, o = make_pair(SDF, 256, dim, W, H)
    zero = np.zeros((480, 640), np.uint16)
    pose = synth.planar_sweep_pose(0, dim)
    o.preprocess(zero); o.integrate(pose, k, mu, 0)
    g.preprocess(zero); g.integrate(pose, k, mu, 0)
    assert g.block_count() == 0 == o.block_count() and g.node_count() == 1
    o.raycast(pose, k, mu); g.raycast(pose, k, mu)
    gv, gn = g.vertex_normal()
    assert np.all(gn[..., 0] == -2) and np.all(gv == 0)
    d, pose = synth.planar_sweep(1, dim, 640, 480, K640)
    o.preprocess(d); o.integrate(pose, k, mu, 1)
    g.preprocess(d); g.integrate(pose, k, mu, 1)
    assert np.array_equal(g.render_depth(), o.render_depth())
    assert_sdf_bit_exact(g, o, pose, k, mu)


def test_sdf_camera_outside_and_partially_out_of_volume():
    """Rays that leave the volume (samples outside [0,size) are skipped, alloc_impl.hpp:92-94) and a camera
    placed outside the cube."""
    from supereight_b200 import synth
    dim, mu, W, H = 2.0, 0.1, 160, 120
    k = scaled_k(W)
    g, o = make_pair(SDF, 128, dim, W, H)
    for f, (tx, tz) in enumerate([(1.0, -0.6), (1.9, 0.2), (0.05, 0.3)]):
        pose = synth.yaw_pose(tx, 1.0, tz, 0.3 * f)
        d, _ = synth.planar_sweep(0, 3.2, W, H, k, dropout=0.05)       # wall at z = 2.4 m: beyond the 2 m cube for some rays
        o.preprocess(d); o.integrate(pose, k, mu, f)
        g.preprocess(d); g.integrate(pose, k, mu, f)
    assert_sdf_bit_exact(g, o, pose, k, mu)


def test_sdf_weight_saturation_property():
    """maxweight = 100 (DenseSLAMSystem.cpp:235): after 105 static frames every updated voxel has weight 100."""
    from supereight_b200 import Map, synth
    dim, mu, W, H = 4.8, 0.1, 80, 60
    k = scaled_k(W)
    g = Map(SDF, 128, dim, W, H)
    d, pose = synth.planar_sweep(0, dim, W, H, k, dropout=0.0)
    g.preprocess(d)
    for f in range(105):
        g.integrate(pose, k, mu, f)
    keys, coords, active, data = g.blocks_sorted()
    w = np.unique(data["y"])
    assert set(w.tolist()) <= {0.0, 100.0} and 100.0 in w
    assert np.all(np.abs(data["x"]) <= 1.0)


# ---- size-independent properties at BASELINE's largest configuration -------------------------
def test_sdf_2048_deep_tree_properties_and_sampled_parity():
    """configs[3]: SDF 2048^3 @ 2 mm.  The oracle needs ~30 M ray samples per frame here, so parity is checked
    on one 320x240 frame, and the structural invariants of the map on the device side."""
    from supereight_b200 import synth
    dim, mu, W, H = 4.096, 0.1, 320, 240
    k = scaled_k(W)
    g, o = make_pair(SDF, 2048, dim, W, H, max_blocks=400000)
    pose = run_sequence(g, o, synth.box_room, dim, W, H, k, mu, [0], n_frames=300, dropout=0.01)
    cb = compare_blocks(g, o)
    assert cb["keys_equal"] and cb["x_bit_mismatch"] == 0 and cb["y_mismatch"] == 0, cb
    keys, coords, active, _ = g.blocks_sorted(with_data=False)
    assert len(np.unique(keys)) == len(keys) > 20000                       # no block allocated twice
    assert np.all((keys & 0x1FF) == 8)                                     # leaves level = 11 - 3
    assert np.all(coords % 8 == 0) and coords.min() >= 0 and coords.max() < 2048
    lib = oracle_lib.load()
    sample = np.random.default_rng(0).choice(len(keys), 500, replace=False)
    for i in sample:                                                       # key <-> coordinates are consistent
        assert lib.seo_key_encode(int(coords[i, 0]), int(coords[i, 1]), int(coords[i, 2]), 8, 11) == int(keys[i])
    codes, side, mask, _ = g.nodes_sorted()
    assert len(np.unique(codes)) == len(codes)
    present = set(int(c) for c in codes) | set(int(c) for c in keys)
    for c, s, m in list(zip(codes, side, mask))[:2000]:                    # children_mask <=> child exists; side halves per level
        level = int(c) & 0x1FF
        assert int(s) == 2048 >> level
        out = (__import__("ctypes").c_int * 3)()
        lib.seo_morton_decode(int(c) & ~0x1FF, out)
        for i in range(8):
            half = int(s) // 2
            child = lib.seo_key_encode(out[0] + (i & 1) * half, out[1] + ((i >> 1) & 1) * half, out[2] + ((i >> 2) & 1) * half, level + 1, 11)
            assert bool(m & (1 << i)) == (child in present)


# ---- OFusion ---------------------------------------------------------------------------------
# What the OFusion path is held to.  north_star's bar is 1e-4 relative; MEASURED on the B200 (scripts/ofusion_deviation.py, round 2:
# four scenarios, 160x120 .. 640x480 into 256^3 .. 1024^3) the only differences are occupancies 1 ulp apart (<= 4.8e-7 relative,
# where glibc's log2f is not the correctly rounded value the device computes -- 0 to 0.13 % of the voxels), no hit flips, no
# differing vertex bit and at most one normal component 1 ulp off.  The assertions sit a small factor above that.
OFU_X_RTOL = 2e-6          # occupancy (log-odds), relative: 4 ulp
OFU_FLIPS = 2e-5           # fraction of pixels whose hit / miss may differ: a handful per 640x480 image (observed: none)
OFU_VERTEX_ATOL = 2e-6     # metres (observed: bit-identical)
OFU_NORMAL_ATOL = 1e-6     # (observed: <= 1.5e-8)


def assert_ofusion_parity(g, o, pose, k, mu):
    gk, gc, ga, gd = g.blocks_sorted()
    ok, oc, oa, od = o.blocks_sorted()
    assert np.array_equal(gk, ok), (len(gk), len(ok), len(np.setdiff1d(gk, ok)), len(np.setdiff1d(ok, gk)))   # allocation set bit-exact
    assert np.array_equal(gc, oc) and np.array_equal(ga, oa)
    assert np.array_equal(gd["y"], od["y"])                                 # timestamps exact
    np.testing.assert_allclose(gd["x"], od["x"], rtol=OFU_X_RTOL, atol=1e-7)
    gcodes, gs, gm, gvv = g.nodes_sorted()
    ocodes, os_, om, ovv = o.nodes_sorted()
    assert np.array_equal(gcodes, ocodes) and np.array_equal(gs, os_) and np.array_equal(gm, om)
    np.testing.assert_allclose(gvv["x"], ovv["x"], rtol=OFU_X_RTOL, atol=1e-7)
    assert np.array_equal(gvv["y"], ovv["y"])
    o.raycast(pose, k, mu); g.raycast(pose, k, mu)
    gv, gn = g.vertex_normal()
    ov, on = o.vertex(), o.normal()
    ghit, ohit = gn[..., 0] != -2, on[..., 0] != -2
    assert ghit.sum() > 0.5 * ghit.size
    assert np.count_nonzero(ghit != ohit) <= OFU_FLIPS * ghit.size
    both = ghit & ohit
    np.testing.assert_allclose(gv[both], ov[both], rtol=0, atol=OFU_VERTEX_ATOL)
    np.testing.assert_allclose(gn[both], on[both], rtol=0, atol=OFU_NORMAL_ATOL)
    bit_equal = np.count_nonzero(gd["x"].view(np.uint32) != od["x"].view(np.uint32))
    return bit_equal / gd["x"].size


def test_ofusion_1024_room_sequence():
    """configs[2]: OFusion, 1024^3 @ 4.8 m, mu 0.008 (Makefile:38-39), multi-level octant requests."""
    from supereight_b200 import synth
    dim, mu, W, H = 4.8, 0.008, 320, 240
    k = scaled_k(W)
    g, o = make_pair(OFUSION, 1024, dim, W, H)
    pose = run_sequence(g, o, synth.box_room, dim, W, H, k, mu, range(0, 20, 4), n_frames=300, dropout=0.01)
    frac = assert_ofusion_parity(g, o, pose, k, mu)
    assert frac < 2e-3          # bit-identical except where glibc's log2f is not the correctly rounded value (1 ulp; observed: none here, <= 0.13 % elsewhere)
    assert np.array_equal(g.render_volume(pose, k, mu, 0.75 * mu, False)[..., 3], o.render_volume(pose, k, mu, 0.75 * mu, False)[..., 3])


def test_ofusion_plane_512_full_frame():
    from supereight_b200 import synth
    dim, mu, W, H = 4.8, 0.008, 640, 480
    g, o = make_pair(OFUSION, 512, dim, W, H)
    pose = run_sequence(g, o, synth.planar_sweep, dim, W, H, K640, mu, range(3), noise_mm=2.0, dropout=0.01)
    assert_ofusion_parity(g, o, pose, K640, mu)


# ---- explicit key lists, point queries, ray walks: the se_core KATs through the C ABI ---------
BLOCKS10 = [(56, 12, 254), (87, 32, 423), (128, 128, 128), (136, 128, 128), (128, 136, 128), (136, 136, 128),
            (128, 128, 136), (136, 128, 136), (128, 136, 136), (136, 136, 136)]


@pytest.mark.parametrize("field", [SDF, OFUSION])
def test_allocate_keys_matches_oracle(field):
    g, o = make_pair(field, 512, 5.0, 8, 8)
    keys = [o.hash(*b) for b in BLOCKS10]
    keys[2] = keys[2] | 3                 # multiscale_unittest.cpp:129-147 (OctantAlloc): mixed levels
    keys[9] = keys[2] | 5
    keys += keys[:3]                      # duplicates
    o.allocate(keys); g.allocate(keys)
    assert np.array_equal(g.blocks_sorted(False)[0], o.blocks_sorted(False)[0])
    gc, gs, gm, _ = g.nodes_sorted(); oc, os_, om, _ = o.nodes_sorted()
    assert np.array_equal(gc, oc) and np.array_equal(gs, os_) and np.array_equal(gm, om)
    # the keys[0] rule: smallest surviving key below the leaves level grows a first-child chain
    g2, o2 = make_pair(field, 512, 5.0, 8, 8)
    ks = [o2.hash(64, 64, 64, 4), o2.hash(320, 64, 64, 5)]
    o2.allocate(ks); g2.allocate(ks)
    assert np.array_equal(g2.blocks_sorted(False)[0], o2.blocks_sorted(False)[0]) and g2.block_count() == 1
    assert np.array_equal(g2.nodes_sorted()[0], o2.nodes_sorted()[0])
    g2.allocate([])                       # no-op
    assert g2.block_count() == 1


def test_point_queries_match_oracle_all_gather_cases():
    g, o = make_pair(SDF, 512, 5.0, 8, 8)
    keys = [o.hash(*b) for b in BLOCKS10]
    o.allocate(keys); g.allocate(keys)
    rng = np.random.default_rng(5)
    xyz = np.array([(x, y, z) for x in range(126, 146) for y in range(127, 145) for z in range(127, 145)], np.int32)
    vals = np.zeros(len(xyz), g.vdtype)
    vals["x"] = rng.uniform(-1, 1, len(xyz)).astype(np.float32)
    vals["y"] = rng.integers(0, 5, len(xyz)).astype(np.float32)
    g.set_voxels(xyz, vals)
    for p, v in zip(xyz, vals):
        if o.fetch(int(p[0]), int(p[1]), int(p[2])):
            o.set_voxel(int(p[0]), int(p[1]), int(p[2]), float(v["x"]), float(v["y"]))
    got = g.query_voxels(xyz)
    want = np.array([o.get_fine(int(p[0]), int(p[1]), int(p[2])) for p in xyz])
    assert np.array_equal(got["x"], want[:, 0].astype(np.float32)) and np.array_equal(got["y"], want[:, 1].astype(np.float32))
    # interpolation / gradient at positions covering all 8 block-crossing cases, the volume border and unallocated space
    pos = np.concatenate([rng.uniform(126, 145, (3000, 3)), rng.uniform(-1, 3, (50, 3)), rng.uniform(509, 513, (50, 3)),
                          np.array([[135.25, 135.5, 135.75], [135.0, 128.0, 131.0], [127.9, 135.1, 135.9]])]).astype(np.float32)
    gi = g.query_interp(pos)
    oi = np.array([o.interp(float(p[0]), float(p[1]), float(p[2])) for p in pos], np.float32)
    assert np.array_equal(gi.view(np.uint32), oi.view(np.uint32))
    gg = g.query_grad(pos)
    og = np.array([o.grad(float(p[0]), float(p[1]), float(p[2])) for p in pos], np.float32)
    assert np.array_equal(gg.view(np.uint32), og.view(np.uint32))


def test_ray_walk_first_block_matches_oracle():
    """ray_iterator_unittest.cpp:46-87 plus random rays: first block, tmin, tmax, tcmin."""
    g, o = make_pair(SDF, 512, 5.0, 8, 8)
    p = np.array([1.5, 1.5, 1.5], np.float32)
    d = np.array([0.5, 0.5, 0.5], np.float32); d = d / np.sqrt(np.float32((d * d).sum()))
    vs = np.float32(5.0) / np.float32(512)
    keys, t = [], np.float32(0.6)
    for _ in range(4):
        vox = ((p + t * d) / vs).astype(np.int32)
        keys.append(o.hash(int(vox[0]), int(vox[1]), int(vox[2]))); t = t + np.float32(2) * (vs * np.float32(8))
    keys += [o.hash(*b) for b in BLOCKS10]
    o.allocate(keys); g.allocate(keys)
    rng = np.random.default_rng(11)
    rays = [np.concatenate([p, d])]
    for _ in range(400):
        org = rng.uniform(-1.0, 6.0, 3)
        tgt = np.array(BLOCKS10[rng.integers(len(BLOCKS10))], np.float64) * float(vs) + rng.uniform(0, 0.08, 3)
        dr = tgt - org; dr /= np.linalg.norm(dr)
        if rng.random() < 0.1:
            dr[rng.integers(3)] = 0.0     # exercises the epsilon clamp of ray_iterator.hpp:66-71
            dr /= max(np.linalg.norm(dr), 1e-9)
        rays.append(np.concatenate([org, dr]))
    rays = np.array(rays, np.float32)
    gk, gt = g.query_rays(rays, 0.4, 4.0)
    for i, r in enumerate(rays):
        blocks, tinfo = o.ray_blocks(r[:3], r[3:], 0.4, 4.0)
        want = int(blocks[0]) if len(blocks) else 0xFFFFFFFFFFFFFFFF
        assert int(gk[i]) == want, i
        assert np.array_equal(gt[i].view(np.uint32), tinfo.view(np.uint32)), (i, gt[i], tinfo)
    assert int(gk[0]) == keys[0]


# ---- API behaviour -----------------------------------------------------------------------------
def test_error_paths_and_render_track():
    from supereight_b200 import Map, SeB200Error
    g = Map(SDF, 256, 4.8, 160, 120)
    with pytest.raises(SeB200Error, match="Invalid ratio"):
        g.preprocess(np.zeros((100, 160), np.uint16))          # preprocessing.cpp:165-176
    with pytest.raises(SeB200Error, match="Invalid ratio"):
        g.preprocess(np.zeros((240, 480), np.uint16))
    with pytest.raises(SeB200Error):
        Map(SDF, 300, 4.8, 160, 120)                            # not a power of two
    res = np.zeros((120, 160, 8), np.int32)
    codes = [1, -1, -2, -3, -4, -5, 0, 7]
    for i, c in enumerate(codes):
        res[:, i * 20:(i + 1) * 20, 0] = c
    got = g.render_track(res, stride_ints=8)
    want = np.empty((120, 160, 4), np.uint8)
    oracle_lib.load().seo_render_track(want.ctypes.data, res.ctypes.data, 8, 160, 120)
    assert np.array_equal(got, want)
    # pool exhaustion is reported, not silent
    from supereight_b200 import synth
    small = Map(SDF, 256, 4.8, 160, 120, max_blocks=16)
    d, pose = synth.planar_sweep(0, 4.8, 160, 120, scaled_k(160))
    small.preprocess(d); small.integrate(pose, scaled_k(160), 0.1, 0)
    with pytest.raises(SeB200Error, match="pool exhausted"):
        small.block_count()


def test_pool_exhaustion_surfaces_from_the_per_frame_calls():
    """A caller that only drives the stages (the DenseSLAMSystem shim never asks for counters) still learns that a pool ran
    out: the frame's integrate kernel hands the error bits to the host, and a later se_b200_integrate / se_b200_raycast returns
    SE_B200_ERR_POOL -- once; the condition is cleared, and the map keeps working with what fitted."""
    from supereight_b200 import Map, SeB200Error, synth
    k = scaled_k(160)
    small = Map(SDF, 256, 4.8, 160, 120, max_blocks=16)
    raised = 0
    for f in range(6):
        d, pose = synth.planar_sweep(f, 4.8, 160, 120, k)
        small.preprocess(d)
        for call in (lambda: small.integrate(pose, k, 0.1, f), lambda: small.raycast(pose, k, 0.1)):
            try:
                call()
            except SeB200Error as e:
                assert "pool exhausted" in str(e)
                raised += 1
        small.sync()                      # (the per-frame calls do not wait for the device; give the flag time to arrive)
    assert raised >= 1
    # the pool stays full, so every frame raises the condition anew; a map with room never reports anything
    assert small.block_count_nothrow() == 16
    big = Map(SDF, 256, 4.8, 160, 120)
    for f in range(3):
        d, pose = synth.planar_sweep(f, 4.8, 160, 120, k)
        big.preprocess(d); big.integrate(pose, k, 0.1, f); big.raycast(pose, k, 0.1); big.sync()
    assert big.block_count() > 16


@pytest.mark.parametrize("field,mu", [(SDF, 0.1), (OFUSION, 0.008)])
def test_map_export_import_round_trip(field, mu):
    """Octree::save -> Octree::load (octree.hpp:897-950) through the ABI: a map rebuilt from its exported records
    is the same map (same keys, same payloads, same raycast)."""
    from supereight_b200 import Map, synth
    dim, W, H = 4.8, 160, 120
    k = scaled_k(W)
    a = Map(field, 256, dim, W, H)
    for f in range(3):
        d, pose = synth.planar_sweep(f, dim, W, H, k)
        a.preprocess(d); a.integrate(pose, k, mu, f)
    keys, coords, active, data = a.blocks_sorted()
    codes, side, mask, values = a.nodes_sorted()
    b = Map(field, 256, dim, W, H)
    b.upload_nodes(codes, values)
    b.upload_blocks(keys, data)
    keys2, coords2, _, data2 = b.blocks_sorted()
    codes2, side2, mask2, values2 = b.nodes_sorted()
    assert np.array_equal(keys, keys2) and np.array_equal(coords, coords2) and data.tobytes() == data2.tobytes()
    assert np.array_equal(codes, codes2) and np.array_equal(side, side2) and np.array_equal(mask, mask2)
    assert values.tobytes() == values2.tobytes()
    a.raycast(pose, k, mu); b.raycast(pose, k, mu)
    va, na = a.vertex_normal(); vb, nb = b.vertex_normal()
    assert va.tobytes() == vb.tobytes() and na.tobytes() == nb.tobytes()


def test_sdf_ieee_division_fallback_paths(monkeypatch):
    """The integrate kernel has two instantiations: check-free FMA sequences when every parameter
    is 0 or within [2^-20, 2^20], plain IEEE operators otherwise.  Both must be bit-exact: (a) forced by the
    environment, (b) selected automatically by a pose with a 1e-9 rotation entry."""
    from supereight_b200 import synth
    dim, mu, W, H = 4.8, 0.1, 160, 120
    k = scaled_k(W)
    monkeypatch.setenv("SE_B200_IEEE_DIV", "1")
    g, o = make_pair(SDF, 256, dim, W, H)
    pose = run_sequence(g, o, synth.planar_sweep, dim, W, H, k, mu, range(3), noise_mm=2.0)
    assert_sdf_bit_exact(g, o, pose, k, mu)
    monkeypatch.delenv("SE_B200_IEEE_DIV")
    g, o = make_pair(SDF, 256, dim, W, H)
    for f in range(3):
        d, pose = synth.planar_sweep(f, dim, W, H, k)
        pose = pose.copy(); pose[0, 1] = 1e-9; pose[1, 0] = -1e-9      # outside the fast kernel's precondition
        o.preprocess(d); o.integrate(pose, k, mu, f)
        g.preprocess(d); g.integrate(pose, k, mu, f)
    assert_sdf_bit_exact(g, o, pose, k, mu)


def test_ragged_image_sizes_and_tiny_volume():
    """Image sizes that are not multiples of the 8x4 pixel tile, and the smallest supported volume."""
    from supereight_b200 import synth
    dim, mu, W, H = 1.0, 0.05, 150, 77
    k = (110.0, 108.0, 75.0, 38.0)
    g, o = make_pair(SDF, 16, dim, W, H)
    for f in range(3):
        pose = synth.yaw_pose(0.5, 0.5, -0.3, 0.05 * f)
        d, _ = synth.planar_sweep(0, 1.0, W, H, k, dropout=0.1)      # wall at z = 0.75
        d = (d.astype(np.float32) * 1.3).astype(np.uint16)           # push it to ~1 m in front of the camera
        o.preprocess(d); o.integrate(pose, k, mu, f)
        g.preprocess(d); g.integrate(pose, k, mu, f)
    assert g.block_count() == o.block_count() and g.block_count() <= 8
    assert_sdf_bit_exact(g, o, pose, k, mu)
    g2, o2 = make_pair(OFUSION, 64, 2.0, 33, 17)
    kk = (30.0, 30.0, 16.0, 8.0)
    for f in range(2):
        d, pose = synth.box_room(f, 2.0, 33, 17, kk, n_frames=20)
        o2.preprocess(d); o2.integrate(pose, kk, 0.02, f)
        g2.preprocess(d); g2.integrate(pose, kk, 0.02, f)
    assert np.array_equal(g2.blocks_sorted(False)[0], o2.blocks_sorted(False)[0])
    assert np.array_equal(g2.nodes_sorted()[0], o2.nodes_sorted()[0])


def test_host_calls_with_pinned_buffers_match_pageable_ones():
    """se_b200_render_volume_host writes a page-locked destination in place (no staging copy) and
    se_b200_preprocess_depth_host copies asynchronously from a page-locked source: same bytes as with pageable buffers"""
    import ctypes as C

    import torch
    from supereight_b200 import synth
    dim, mu, W, H = 4.8, 0.1, 160, 120
    k = scaled_k(W)
    g, o = make_pair(SDF, 256, dim, W, H)
    pinned_depth = torch.empty((H, W), dtype=torch.int16).pin_memory()
    pinned_out = torch.zeros((H, W, 4), dtype=torch.uint8).pin_memory()
    kk = np.ascontiguousarray(k, np.float32)
    for f in range(4):
        d, pose = synth.planar_sweep(f, dim, W, H, k, noise_mm=2.0, dropout=0.01)
        pinned_depth.numpy()[...] = d.view(np.int16)
        assert g.lib.se_b200_preprocess_depth_host(g.h, C.c_void_p(pinned_depth.data_ptr()), W, H) == 0
        o.preprocess(d)
        g.integrate(pose, k, mu, f); o.integrate(pose, k, mu, f)
    g.raycast(pose, k, mu); o.raycast(pose, k, mu)
    p = np.ascontiguousarray(pose, np.float32)
    for reraycast in (0, 1):
        pinned_out.zero_()
        rc = g.lib.se_b200_render_volume_host(g.h, C.c_void_p(pinned_out.data_ptr()), C.c_void_p(p.ctypes.data), C.c_void_p(kk.ctypes.data),
                                              C.c_float(mu), C.c_float(0.75 * mu), reraycast)
        assert rc == 0
        want = o.render_volume(pose, k, mu, 0.75 * mu, bool(reraycast))
        assert np.array_equal(pinned_out.numpy(), want)
        assert np.array_equal(g.render_volume(pose, k, mu, 0.75 * mu, bool(reraycast)), want)       # pageable destination


def test_overlapped_host_io_gives_the_same_frames():
    """se_b200_preprocess_depth_host_async / se_b200_render_volume_host_async (copy streams, double-buffered staging): issued
    back to back without synchronising, the map, the last two images and the vertex / normal maps equal the oracle's"""
    import ctypes as C

    import torch
    from supereight_b200 import synth
    dim, mu, W, H, frames = 4.8, 0.1, 160, 120, 7
    k = scaled_k(W)
    kk = np.ascontiguousarray(k, np.float32)
    g, o = make_pair(SDF, 256, dim, W, H)
    depth = torch.empty((frames, H, W), dtype=torch.int16).pin_memory()
    outs = [torch.zeros((H, W, 4), dtype=torch.uint8).pin_memory() for _ in range(2)]
    poses, want = [], []
    for f in range(frames):
        d, pose = synth.box_room(f * 5, dim, W, H, k, noise_mm=2.0, dropout=0.01)
        depth[f].numpy()[...] = d.view(np.int16)
        poses.append(np.ascontiguousarray(pose, np.float32))
        o.preprocess(d); o.integrate(pose, k, mu, f); o.raycast(pose, k, mu)
        want.append(o.render_volume(pose, k, mu, 0.75 * mu, False))
    for f in range(frames):                                     # no synchronisation inside the loop
        p = C.c_void_p(poses[f].ctypes.data)
        assert g.lib.se_b200_preprocess_depth_host_async(g.h, C.c_void_p(depth[f].data_ptr()), W, H) == 0
        assert g.lib.se_b200_integrate(g.h, p, C.c_void_p(kk.ctypes.data), C.c_float(mu), f) == 0
        assert g.lib.se_b200_raycast(g.h, p, C.c_void_p(kk.ctypes.data), C.c_float(mu)) == 0
        assert g.lib.se_b200_render_volume_host_async(g.h, C.c_void_p(outs[f & 1].data_ptr()), p, C.c_void_p(kk.ctypes.data),
                                                      C.c_float(mu), C.c_float(0.75 * mu), 0) == 0
    g.sync()
    assert np.array_equal(outs[(frames - 1) & 1].numpy(), want[-1])
    assert np.array_equal(outs[(frames - 2) & 1].numpy(), want[-2])
    res = compare_blocks(g, o)
    assert res["keys_equal"] and res["x_bit_mismatch"] == 0 and res["y_mismatch"] == 0, res
    gv, gn = g.vertex_normal()
    assert np.array_equal(gv.view(np.uint32), o.vertex().view(np.uint32)) and np.array_equal(gn.view(np.uint32), o.normal().view(np.uint32))
