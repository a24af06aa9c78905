"""The oracle against THE REFERENCE'S OWN CODE.

oracle/_ref/libse_ref_<field>.so is /root/reference/se_denseslam/src/DenseSLAMSystem.cpp (with everything it includes:
octree, allocation, projective functor, ray iterator, interpolation, rendering, tracking, meshing), compiled where it lies,
unmodified, against the stand-in Eigen / Sophus headers in oracle/ref_standin (the image has neither library), behind the
same seo_* entry points as the oracle (oracle/ref_capi.cpp).  It is built only where /root/reference exists (the
development container); the built files travel with the repo snapshot.  Every array the pipeline produces must be
bit-identical between the two; tests/golden/seq_*.npz are written from this build (tests/golden/make_golden.py).

What this pins and what it does not: the oracle's restatement of the reference's ALGORITHM (control flow, indexing,
formulas, evaluation order as written in the source) is pinned for every stage of the path.  The linear-algebra layer
underneath is the stand-in, which evaluates sums left to right in plain fp32 and inverts rigid / camera matrices in closed
form; real Eigen (SIMD kernels, -march=native) and real Sophus (unit-quaternion rotation) differ from that by a few ulp."""
import json
import os
import subprocess
import sys

import pytest

import oracle_lib

HERE = os.path.dirname(os.path.abspath(__file__))
needs_ref = pytest.mark.skipif(not oracle_lib.have_reference_build(), reason="oracle/_ref not built (needs /root/reference)")

EXACT = ["block_keys", "block_coords", "block_active", "block_x", "block_y", "node_codes", "node_side", "node_mask", "node_x", "node_y",
         "depth", "vertex", "normal", "render_reuse", "render_view", "render_depth", "get", "interp", "grad", "ray_blocks",
         "track_result", "track_error", "track_J", "reduction", "pose", "tracked", "mesh_vertices_differ"] + \
        [f"pyramid{lvl}_{nm}" for lvl in range(3) for nm in ("depth", "vertex", "normal")]


def compare(field, size, dim, W, H, frames):
    r = subprocess.run([sys.executable, os.path.join(HERE, "_ref_compare_worker.py"), field, str(size), str(dim), str(W), str(H), str(frames)],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


@needs_ref
@pytest.mark.parametrize("field,size,dim,W,H,frames", [("sdf", 256, 4.8, 160, 120, 5), ("ofusion", 256, 4.8, 160, 120, 5), ("sdf", 128, 2.4, 80, 60, 4)])
def test_oracle_is_bit_identical_to_the_reference_build(field, size, dim, W, H, frames):
    res = compare(field, size, dim, W, H, frames)
    assert res["n_blocks"] > 100 and res["hits"] > 1000
    wrong = {k: res[k] for k in EXACT if res[k] != 0}
    assert not wrong, wrong
    # N4: the reference meshes with its own edge_tables.h, the oracle with the table the library ships (the same classic list):
    # the same triangles, every vertex bit for bit (the reference's ORDER depends on its OpenMP schedule: compared as sets)
    assert res["mesh_triangles"][0] == res["mesh_triangles"][1] > 1000
    assert res["mesh_identical_triangles_frac"] == 1.0 and res["mesh_triangles_only_in_oracle"] == 0


@needs_ref
@pytest.mark.parametrize("size", [64, 128])
def test_oracle_equals_the_reference_build_on_random_scenarios(size):
    """scripts/fuzz_parity.py --ref: random cameras, depth images and 1-4 frame sequences (both fields) inside the domain where the
    reference's code is defined -- the oracle against the reference's own sources, every array bit for bit.  A fixed batch here;
    some 600 scenarios over five volume sizes were run when it was written: no difference."""
    r = subprocess.run([sys.executable, os.path.join(os.path.dirname(HERE), "scripts", "fuzz_parity.py"), "--ref", str(size), "12", "300"],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-1000:]
    assert " 0 with differences" in r.stdout, r.stdout[-1000:]


@needs_ref
def test_reference_build_exports_what_the_oracle_binding_uses():
    for kind in ("ref_sdf", "ref_ofusion"):
        lib = oracle_lib.load(kind)
        for name in ("seo_create", "seo_preprocess", "seo_integrate", "seo_raycast", "seo_render_volume", "seo_get_blocks_sorted", "seo_tracking", "seo_marching_cube"):
            assert hasattr(lib, name), (kind, name)
        assert oracle_lib.Oracle(1 if kind == "ref_sdf" else 0, 64, 1.0, 8, 8, kind=kind).h.value is None      # the field type is a compile-time choice


REF_TESTS = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "tests")
REF_TEST_NAMES = ["axisaligned", "image", "aabb_collision", "octree_collision", "octree", "ray_iterator", "alloc", "gather", "interpolation", "io",
                  "unique", "math", "morton", "multiscale"]


@pytest.mark.skipif(not os.path.isdir(REF_TESTS), reason="oracle/_ref/tests not built (needs /root/reference)")
@pytest.mark.parametrize("name", REF_TEST_NAMES)
def test_the_references_own_unit_tests_pass_on_the_standin_build(name, tmp_path):
    """se_core/test/**/<name>_unittest.cpp, compiled unmodified against oracle/ref_standin (Eigen + gtest stand-ins): the
    reference's own known-answer tests hold on the build the oracle is compared with (14 files, 56 tests)"""
    exe = os.path.join(REF_TESTS, name + "_unittest")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300, cwd=str(tmp_path))     # io_unittest writes files
    assert r.returncode == 0, r.stdout[-3000:]
    lines = r.stdout.strip().splitlines()
    assert any(ln.startswith("[  PASSED  ]") for ln in lines) and not any(ln.startswith("[  FAILED  ]") for ln in lines)


def test_gtest_standin_reports_failures(tmp_path):
    """the stand-in gtest.h must fail when an assertion fails (fatal ones return, non-fatal ones continue)"""
    src = tmp_path / "t.cpp"
    src.write_text('#include "gtest/gtest.h"\n'
                   'TEST(S, ok) { ASSERT_EQ(2, 1 + 1); }\n'
                   'TEST(S, fatal) { ASSERT_EQ(3, 1 + 1) << "msg"; std::printf("UNREACHED\\n"); }\n'
                   'TEST(S, nonfatal) { EXPECT_TRUE(false); std::printf("CONTINUED\\n"); }\n')
    standin = os.path.join(os.path.dirname(HERE), "oracle", "ref_standin")
    exe = str(tmp_path / "t")
    subprocess.run(["/usr/bin/g++", "-std=c++14", "-I" + standin, str(src), os.path.join(standin, "gtest", "gtest_main.cpp"), "-o", exe], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 1
    assert "[  PASSED  ] 1 tests." in r.stdout and "[  FAILED  ] 2 tests." in r.stdout
    assert "UNREACHED" not in r.stdout and "CONTINUED" in r.stdout and "msg" in r.stdout



def test_half_sample_pyramid_with_odd_level_widths(monkeypatch):
    """halfSampleRobustImageKernel indexes its input with in.width() (preprocessing.cpp:209,217): at 100 x 76 the pyramid is
    100 / 50 / 25 / 12 wide and level 3 reads a parent whose row stride (25) is not twice its own width (the reference
    accepts it: 25 / 12 == 2 in integer arithmetic).  Oracle and reference build, four levels, bit for bit."""
    import numpy as np
    from oracle_lib import SDF, Oracle
    from supereight_b200 import synth
    monkeypatch.setenv("SEO_REF_PYRAMID_LEVELS", "4")
    W, H, dim = 100, 76, 4.8
    k = tuple(v * W / 640.0 for v in synth.DEFAULT_K)
    o, r = Oracle(SDF, 128, dim, W, H), Oracle(SDF, 128, dim, W, H, kind="ref_sdf")
    d, pose = synth.corner_view(2, dim, W, H, k, noise_mm=2.0, dropout=0.01)
    for p in (o, r):
        assert p.preprocess(d) == 0
        p.filter_depth(True, 4)
        p.track(pose, pose, k, 1e-5, [0, 0, 0, 0])           # builds the pyramid and the per-level vertex / normal maps, no ICP iteration
    for lvl in range(4):
        od, ov, on = o.pyramid(lvl); rd, rv, rn = r.pyramid(lvl)
        assert od.shape == (H >> lvl, W >> lvl)
        assert od.tobytes() == rd.tobytes() and ov.tobytes() == rv.tobytes() and on.tobytes() == rn.tobytes(), lvl
    assert (o.pyramid(3)[0] > 0).mean() > 0.5
