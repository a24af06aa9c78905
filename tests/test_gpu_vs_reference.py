"""The CUDA path DIRECTLY beside the reference's own code at the benchmark's sizes (GPU tier).

The other parity tests compare the device with the oracle, and tests/test_reference_build.py (CPU tier, small sizes) the
oracle with `oracle/_ref` -- the reference's sources compiled where they lie.  These tests close the chain in one hop at the
sizes BASELINE.json quotes: 640x480 frames into the 512^3 TSDF map (configs[1]) and the OFusion map, through the C ABI on one
side and through the reference's DenseSLAMSystem::{preprocessing, integration, raycasting} + renderVolumeKernel on the other
(single-threaded, in a fresh process: tests/_ref_frames_worker.py), plus one full-size frame of configs[2] and configs[3]
against the oracle.  Bars as everywhere: SDF bit-exact; OFusion occupancies within 1e-4 relative (log2f: libm vs device),
timestamps and the allocation set exact, everything else to the measured-level tolerances of tests/test_gpu_parity.py (OFU_*)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle_lib
from oracle_lib import OFUSION, SDF
from parity_utils import compare_blocks, compare_images
from test_gpu_parity import K640, OFU_FLIPS, OFU_NORMAL_ATOL, OFU_VERTEX_ATOL, OFU_X_RTOL, assert_ofusion_parity, make_pair, run_sequence

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def reference_run(tmp_path, field, size, dim, W, H, mu, scene, frames, n_frames=300):
    if not oracle_lib.have_reference_build():
        pytest.skip("oracle/_ref (the reference build) is absent")
    out = str(tmp_path / "reference.npz")
    r = subprocess.run([sys.executable, os.path.join(HERE, "_ref_frames_worker.py"), field, str(size), str(dim), str(W), str(H), str(mu), scene,
                        ",".join(str(v) for v in frames), out, str(n_frames)], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    return np.load(out)


def gpu_run(field, size, dim, W, H, mu, scene, frames, n_frames=300, **kw):
    from supereight_b200 import Map, synth
    g = Map(field, size, dim, W, H, **kw)
    first, step, count = frames
    pose = None
    for i in range(count):
        f = first + i * step
        if scene == "plane":
            d, pose = synth.planar_sweep(f, dim, W, H, K640, noise_mm=2.0, dropout=0.01)
        else:
            d, pose = synth.box_room(f, dim, W, H, K640, n_frames=n_frames, noise_mm=2.0, dropout=0.01)
        g.preprocess(d); g.integrate(pose, K640, mu, f)
    g.raycast(pose, K640, mu)
    return g, pose


def test_sdf_512_headline_frames_equal_the_reference_build(tmp_path):
    """configs[1]: four 640x480 frames of the planar sweep into 512^3 @ 4.8 m -- every bit of the map, the vertex / normal maps
    and the rendered image equal what the reference's own code produces"""
    dim, mu, W, H, frames = 4.8, 0.1, 640, 480, (0, 1, 4)
    ref = reference_run(tmp_path, "sdf", 512, dim, W, H, mu, "plane", frames)
    g, pose = gpu_run(SDF, 512, dim, W, H, mu, "plane", frames)
    keys, coords, active, data = g.blocks_sorted()
    assert len(keys) > 3000 and np.array_equal(keys, ref["keys"]) and np.array_equal(coords, ref["coords"]) and np.array_equal(active, ref["active"])
    assert data.tobytes() == ref["data"].tobytes()
    codes, side, mask, values = g.nodes_sorted()
    assert np.array_equal(codes, ref["codes"]) and np.array_equal(side, ref["side"]) and values.tobytes() == ref["values"].tobytes()
    gv, gn = g.vertex_normal()
    ci = compare_images(gv, gn, ref["vertex"], ref["normal"])
    assert ci["hits_gpu"] > 0.9 * W * H and ci["hit_mask_mismatch"] == 0 and ci["vertex_bit_mismatch"] == 0 and ci["normal_bit_mismatch"] == 0, ci
    assert np.array_equal(g.render_volume(pose, K640, mu, 0.75 * mu, False), ref["image"])
    assert np.array_equal(g.render_depth(), ref["depth_image"])


def test_ofusion_512_frames_against_the_reference_build(tmp_path):
    """the OFusion library of the reference at 640x480 into 512^3: allocation set (multi-level requests), timestamps exact;
    occupancies within 1e-4 relative; vertex map within 1e-4 where both hit"""
    dim, mu, W, H, frames = 4.8, 0.03, 640, 480, (0, 5, 3)
    ref = reference_run(tmp_path, "ofusion", 512, dim, W, H, mu, "room", frames)
    g, pose = gpu_run(OFUSION, 512, dim, W, H, mu, "room", frames)
    keys, coords, active, data = g.blocks_sorted()
    assert len(keys) > 3000 and np.array_equal(keys, ref["keys"]) and np.array_equal(coords, ref["coords"]) and np.array_equal(active, ref["active"])
    assert np.array_equal(data["y"], ref["data"]["y"])
    np.testing.assert_allclose(data["x"], ref["data"]["x"], rtol=OFU_X_RTOL, atol=1e-7)
    codes, side, mask, values = g.nodes_sorted()
    assert np.array_equal(codes, ref["codes"]) and np.array_equal(side, ref["side"]) and np.array_equal(values["y"], ref["values"]["y"])
    np.testing.assert_allclose(values["x"], ref["values"]["x"], rtol=OFU_X_RTOL, atol=1e-7)
    gv, gn = g.vertex_normal()
    ghit, rhit = gn[..., 0] != -2, ref["normal"][..., 0] != -2
    assert ghit.sum() > 0.5 * ghit.size and np.count_nonzero(ghit != rhit) <= OFU_FLIPS * ghit.size
    both = ghit & rhit
    np.testing.assert_allclose(gv[both], ref["vertex"][both], rtol=0, atol=OFU_VERTEX_ATOL)
    np.testing.assert_allclose(gn[both], ref["normal"][both], rtol=0, atol=OFU_NORMAL_ATOL)


def test_ofusion_1024_one_full_size_frame_against_the_oracle():
    """configs[2] at its full image size: one 640x480 frame into OFusion 1024^3, mu 0.008"""
    from supereight_b200 import synth
    dim, mu, W, H = 4.8, 0.008, 640, 480
    g, o = make_pair(OFUSION, 1024, dim, W, H)
    pose = run_sequence(g, o, synth.box_room, dim, W, H, K640, mu, [3], n_frames=300, noise_mm=2.0, dropout=0.01)
    assert_ofusion_parity(g, o, pose, K640, mu)


def test_sdf_2048_one_full_size_frame_against_the_oracle():
    """configs[3] at its full image size: one 640x480 frame into SDF 2048^3 @ 2 mm (100 samples per ray, ~10^5 new blocks)"""
    from supereight_b200 import synth
    dim, mu, W, H = 4.096, 0.1, 640, 480
    g, o = make_pair(SDF, 2048, dim, W, H, max_blocks=600000)
    pose = run_sequence(g, o, synth.box_room, dim, W, H, K640, mu, [0], n_frames=300, noise_mm=2.0, dropout=0.01)
    cb = compare_blocks(g, o)
    assert cb["n_gpu"] > 80000 and cb["keys_equal"] and cb["coords_equal"] and cb["active_mismatch"] == 0, cb
    assert cb["x_bit_mismatch"] == 0 and cb["y_mismatch"] == 0, cb
    o.raycast(pose, K640, mu); g.raycast(pose, K640, mu)
    gv, gn = g.vertex_normal()
    ci = compare_images(gv, gn, o.vertex(), o.normal())
    assert ci["hits_gpu"] > 0.5 * W * H and ci["hit_mask_mismatch"] == 0 and ci["vertex_bit_mismatch"] == 0 and ci["normal_bit_mismatch"] == 0, ci
