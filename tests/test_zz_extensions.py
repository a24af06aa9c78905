"""GPU tests of what this round added last and could not yet run on the device -- the render-target extension, the OFusion
integrate's default instantiation against its plain-operator twin, and a batch of randomised parity scenarios -- kept in a
file that pytest runs after the parity files (it orders files by name), so that `-x` never lets them hide an established test."""
import numpy as np
import pytest

from test_gpu_parity import OFUSION, REL_TOL, SDF, assert_ofusion_parity, make_pair, run_sequence, scaled_k

pytestmark = pytest.mark.gpu


def _pinned_image(H, W):
    """(array view, address, keep-alive) of a page-locked H x W x 4 byte buffer; on the fiber executor (no CUDA driver) a plain
    numpy buffer, which tests/test_simt_emu.py makes look page-locked to the library (SIMT_HOST_IS_PINNED=1)"""
    import os
    if "simt" in os.environ.get("SE_B200_LIB", ""):
        a = np.zeros((H, W, 4), np.uint8)
        return a, a.ctypes.data, a
    import torch
    t = torch.zeros((H, W, 4), dtype=torch.uint8).pin_memory()
    return t.numpy(), t.data_ptr(), t


@pytest.mark.parametrize("field,mu", [(SDF, 0.1), (OFUSION, 0.008)])
def test_render_target_image_equals_render_volume(field, mu):
    """se_b200_set_render_target: the raycast shades into the caller's page-locked buffer and renderVolume's reuse path then
    launches nothing -- same image, same vertex / normal maps, one launch fewer per frame; a different view pose, a re-raycast
    or another destination still render as usual, and clearing the target restores the plain path."""
    from supereight_b200 import Map, SeB200Error, synth
    dim, W, H = 4.8, 160, 120
    k = scaled_k(W)
    a, b = Map(field, 256, dim, W, H), Map(field, 256, dim, W, H)
    img, ptr, keep = _pinned_image(H, W)
    b.set_render_target(ptr)
    gen = synth.planar_sweep if field == SDF else synth.box_room
    kw = dict(noise_mm=2.0) if field == SDF else dict(n_frames=300)
    for f in range(4):
        d, pose = gen(f, dim, W, H, k, dropout=0.01, **kw)
        for m_ in (a, b):
            m_.preprocess(d); m_.integrate(pose, k, mu, f)
        la, lb = a.launch_count(), b.launch_count()
        a.raycast(pose, k, mu); b.raycast(pose, k, mu)
        want = a.render_volume(pose, k, mu, 0.75 * mu, False)
        b.render_volume_host_ptr(ptr, pose, k, mu, 0.75 * mu, False)
        assert (a.launch_count() - la, b.launch_count() - lb) == (2, 1)
        assert np.array_equal(img, want)
        va, na = a.vertex_normal(); vb, nb = b.vertex_normal()
        assert va.tobytes() == vb.tobytes() and na.tobytes() == nb.tobytes()
    assert want[..., 0].max() > 0 and (na[..., 0] != -2.0).sum() > 0.2 * W * H          # a real image: rays hit
    # another destination, another view, a re-raycast: rendered the usual way
    assert np.array_equal(b.render_volume(pose, k, mu, 0.75 * mu, False), want)
    moved = pose.copy(); moved[0, 3] += 0.05
    lb = b.launch_count()
    b.render_volume_host_ptr(ptr, moved, k, mu, 0.75 * mu, False)
    assert b.launch_count() - lb == 1 and np.array_equal(img, a.render_volume(moved, k, mu, 0.75 * mu, False))
    b.render_volume_host_ptr(ptr, moved, k, mu, 0.75 * mu, True)
    assert np.array_equal(img, a.render_volume(moved, k, mu, 0.75 * mu, True))
    # target off: the plain kernels again
    b.set_render_target(None)
    lb = b.launch_count()
    b.raycast(pose, k, mu); b.render_volume_host_ptr(ptr, pose, k, mu, 0.75 * mu, False)
    assert b.launch_count() - lb == 2 and np.array_equal(img, want)
    if "simt" not in __import__("os").environ.get("SE_B200_LIB", ""):
        with pytest.raises(SeB200Error, match="page-locked"):
            b.set_render_target(np.zeros((H, W, 4), np.uint8).ctypes.data)       # pageable host memory is refused


def test_ofusion_plain_operator_instantiation(monkeypatch):
    """OFusion integrates with the check-free sequences and the tabulated log-odds increment by default; the
    instantiation with the plain IEEE operators and the per-voxel log2 (SE_B200_OFUSION_FAST=0, or SE_B200_IEEE_DIV)
    must match the oracle just the same and leave the same map.  (Byte equality of the two instantiations is asserted on
    the fiber executor, tests/test_simt_emu.py; here, with the device's own MUFU approximations underneath, the bar is
    the one the 1024^3 test applies against the oracle: timestamps equal, < 1e-3 of the values off by an ulp.)"""
    from supereight_b200 import synth
    dim, mu, W, H = 4.8, 0.008, 160, 120
    k = scaled_k(W)
    maps = []
    for env in (None, "SE_B200_OFUSION_FAST", "SE_B200_IEEE_DIV"):
        if env:
            monkeypatch.setenv(env, "0" if env.endswith("FAST") else "1")
        g, o = make_pair(OFUSION, 256, dim, W, H)
        pose = run_sequence(g, o, synth.box_room, dim, W, H, k, mu, range(0, 12, 4), n_frames=300, dropout=0.01)
        assert assert_ofusion_parity(g, o, pose, k, mu) < 1e-3
        maps.append(g.blocks_sorted())
        if env:
            monkeypatch.delenv(env)
    k0, c0, a0, d0 = maps[0]
    for k1, c1, a1, d1 in maps[1:]:
        assert np.array_equal(k0, k1) and np.array_equal(c0, c1) and np.array_equal(a0, a1)
        assert np.array_equal(d0["y"], d1["y"])
        assert np.count_nonzero(d0["x"].view(np.uint32) != d1["x"].view(np.uint32)) < 1e-3 * d0["x"].size
        np.testing.assert_allclose(d0["x"], d1["x"], rtol=REL_TOL, atol=1e-5)


def test_randomised_parity_scenarios_on_the_device():
    """scripts/fuzz_parity.py through whatever library the tests run on: 60 random scenarios (volumes 16^3..256^3, cameras inside /
    outside / on the faces of the volume, axis-aligned views, negative fy, noisy / saturated / 1 mm depth, 1-4 frames), each compared
    with the oracle -- block set, voxel and node values, vertex / normal maps, both renderings, interp / grad queries, the mesh."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "scripts", "fuzz_parity.py"), "60", "9000"], cwd=root,
                       capture_output=True, text=True, timeout=1800)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-1000:]
    assert " 0 with differences" in r.stdout, r.stdout[-1000:]


@pytest.mark.parametrize("size", [64, 256])
def test_randomised_scenarios_product_against_the_reference_build(size):
    """scripts/fuzz_parity.py --ref-device: the product DIRECTLY against the reference's own code (oracle/_ref) over the
    reference's defined domain -- cameras and surfaces inside the volume -- one volume size per process."""
    import os
    import subprocess
    import sys
    import oracle_lib
    if not oracle_lib.have_reference_build():
        pytest.skip("oracle/_ref (the reference build) is absent")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "scripts", "fuzz_parity.py"), "--ref-device", str(size), "40", "7000"], cwd=root,
                       capture_output=True, text=True, timeout=1800)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-1000:]
    assert " 0 with differences" in r.stdout, r.stdout[-1000:]


@pytest.mark.parametrize("field,mu", [(SDF, 0.1), (OFUSION, 0.008)])
def test_launch_schedule_changes_no_result(field, mu, monkeypatch):
    """LaunchSchedule (se_kernels.cuh): the raycast -- and the OFusion allocation pass -- start their expensive tile groups first,
    from the costs earlier launches recorded.  Only the order of execution may change: over a sequence long enough for the
    order to be re-partitioned several times, every frame's vertex / normal maps, image and map equal those of a map created
    with SE_B200_LAUNCH_ORDER_OFF=1 (image order, nothing recorded)."""
    from supereight_b200 import Map, synth
    dim, W, H = 4.8, 160, 120
    k = scaled_k(W)
    monkeypatch.setenv("SE_B200_LAUNCH_ORDER_OFF", "1")
    a = Map(field, 256, dim, W, H)
    monkeypatch.delenv("SE_B200_LAUNCH_ORDER_OFF")
    b = Map(field, 256, dim, W, H)
    gen = synth.box_room
    for f in range(6):
        d, pose = gen(f * 7, dim, W, H, k, dropout=0.01, n_frames=300)
        for m_ in (a, b):
            m_.preprocess(d); m_.integrate(pose, k, mu, f); m_.raycast(pose, k, mu)
        va, na = a.vertex_normal(); vb, nb = b.vertex_normal()
        assert va.tobytes() == vb.tobytes() and na.tobytes() == nb.tobytes(), f
        assert np.array_equal(a.render_volume(pose, k, mu, 0.75 * mu, False), b.render_volume(pose, k, mu, 0.75 * mu, False))
    ka, _, _, xa = a.blocks_sorted(); kb, _, _, xb = b.blocks_sorted()
    assert np.array_equal(ka, kb)
    if field == SDF:
        assert xa.tobytes() == xb.tobytes()
    assert (nb[..., 0] != -2.0).sum() > 0.2 * W * H


def test_sdf_long_active_list_is_handed_out_dynamically():
    """ActiveList::Cursor (se_kernels.cuh): a list longer than four entries per resident warp (18 944 on a B200) is handed out by ticket, with the
    warps of an exhausted class stealing from the others.  A 1024^3 room at 320x240 gives the integrate kernel tens of thousands of
    blocks -- several rounds on any GPU, and on the fiber executor -- and every voxel must still equal the oracle's."""
    from supereight_b200 import synth
    dim, mu, W, H = 4.096, 0.1, 320, 240
    k = scaled_k(W)
    g, o = make_pair(SDF, 1024, dim, W, H)
    pose = None
    for f in range(2):
        d, pose = synth.box_room(f * 5, dim, W, H, k, dropout=0.01, n_frames=300)
        o.preprocess(d); o.integrate(pose, k, mu, f)
        g.preprocess(d); g.integrate(pose, k, mu, f)
    assert g.counters()["active"] > 20000          # (23 872 / 27 069 after the two frames)
    from test_gpu_parity import assert_sdf_bit_exact
    assert_sdf_bit_exact(g, o, pose, k, mu)

