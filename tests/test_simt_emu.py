"""The kernels' logic, run where there is no GPU.

tests/simt_emu compiles the library's own sources (supereight_b200/csrc, unmodified up to launch syntax) with g++ against a
stand-in CUDA runtime whose kernel launch runs every CUDA thread as a fiber: warp collectives, __syncthreads, the lock-free
tree insert and the shared-memory pipeline's hand-shakes execute as written.  The GPU parity tests are then run against
that build (SE_B200_LIB), so the CPU tier checks the product kernels -- not only the oracle -- bit for bit.

This is a checker, like the oracle: the package never loads it, it is never timed, and it does not make the `-m gpu` tier
redundant (the real MUFU approximations, memory ordering between concurrent threads, TMA and PDL exist only on the device).
"""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "simt_emu"))


@pytest.fixture(scope="session")
def emu_lib(oracle_built):
    import build as simt_build          # tests/simt_emu/build.py
    return simt_build.build()


def run_pytest_on_emu(emu_lib, nodeid, extra_env=None, timeout=900):
    env = dict(os.environ, SE_B200_LIB=emu_lib, **(extra_env or {}))
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", "-p", "no:cacheprovider", nodeid],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout and "failed" not in r.stdout, r.stdout[-2000:]


# the GPU parity tests that finish in seconds on the fiber executor (the whole file passes on it; the larger cases --
# OFusion 1024^3, SDF 2048^3, the ICP sequence -- take minutes and are left to the device)
FAST_GPU_TESTS = [
    "tests/test_gpu_parity.py::test_sdf_512_full_frame_sequence_bit_exact",          # the metric's configuration, 640x480 -> 512^3
    "tests/test_gpu_parity.py::test_sdf_negative_fy_camera",
    "tests/test_gpu_parity.py::test_sdf_ratio2_preprocess_and_empty_frames",
    "tests/test_gpu_parity.py::test_sdf_camera_outside_and_partially_out_of_volume",
    "tests/test_gpu_parity.py::test_sdf_weight_saturation_property",
    "tests/test_gpu_parity.py::test_sdf_ieee_division_fallback_paths",
    "tests/test_zz_extensions.py::test_ofusion_plain_operator_instantiation",
    "tests/test_gpu_parity.py::test_ragged_image_sizes_and_tiny_volume",
    "tests/test_gpu_parity.py::test_error_paths_and_render_track",
    "tests/test_gpu_parity.py::test_map_export_import_round_trip",
    "tests/test_gpu_parity.py::test_point_queries_match_oracle_all_gather_cases",
    "tests/test_gpu_parity.py::test_ray_walk_first_block_matches_oracle",
    "tests/test_meshing.py::test_gpu_mesh_of_an_uploaded_map_equals_oracle",
    "tests/test_meshing.py::test_gpu_mesh_empty_map_and_border_clamp",
    "tests/test_zz_extensions.py::test_launch_schedule_changes_no_result",             # expensive tile groups first: same results
    "tests/test_zz_extensions.py::test_sdf_long_active_list_is_handed_out_dynamically",  # tickets + stealing in the integrate kernel
]


@pytest.mark.parametrize("nodeid", FAST_GPU_TESTS)
def test_gpu_parity_test_passes_on_the_fiber_executor(emu_lib, nodeid):
    run_pytest_on_emu(emu_lib, nodeid)


def test_render_target_extension(emu_lib):
    """se_b200_set_render_target (raycast + shading fused, image written to the caller's page-locked buffer): the executor's
    cudaPointerGetAttributes reports every pointer as page-locked and mapped when SIMT_HOST_IS_PINNED=1"""
    run_pytest_on_emu(emu_lib, "tests/test_zz_extensions.py::test_render_target_image_equals_render_volume", {"SIMT_HOST_IS_PINNED": "1"})


def test_randomised_parity_scenarios(emu_lib):
    """scripts/fuzz_parity.py: random volumes, cameras (inside / outside / on the faces of the volume, axis-aligned, negative
    fy), depth images and short sequences through the product kernels and the oracle -- a fixed batch of seeds here, any
    number by hand or on the device (some 750 more were run when it was written: no difference)."""
    env = dict(os.environ, SE_B200_LIB=emu_lib)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "fuzz_parity.py"), "16", "1000"], cwd=ROOT, env=env,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-1000:]
    assert " 0 with differences" in r.stdout, r.stdout[-1000:]


def test_tree_descent_without_the_directories(emu_lib):
    """SE_B200_DISABLE_DIRECTORY=1: every fetch is the root-to-leaf descent, the allocation pass de-duplicates in the warp"""
    run_pytest_on_emu(emu_lib, "tests/test_gpu_parity.py::test_sdf_ratio2_preprocess_and_empty_frames", {"SE_B200_DISABLE_DIRECTORY": "1"})
    run_pytest_on_emu(emu_lib, "tests/test_gpu_parity.py::test_map_export_import_round_trip", {"SE_B200_DISABLE_DIRECTORY": "1"})


def run_worker(emu_lib, field, out, extra_env):
    env = dict(os.environ, SE_B200_LIB=emu_lib, **extra_env)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "_simt_worker.py"), field, out], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return np.load(out)


@pytest.mark.parametrize("field", ["sdf", "ofusion"])
def test_check_free_arithmetic_gives_the_ieee_bits(emu_lib, tmp_path, field):
    """The integrate kernels' check-free division / square-root sequences (and, for OFusion, the tabulated log-odds
    increment) against the instantiation with the plain IEEE operators: every array identical, bit for bit."""
    fast = run_worker(emu_lib, field, str(tmp_path / "fast.npz"), {})
    ieee = run_worker(emu_lib, field, str(tmp_path / "ieee.npz"), {"SE_B200_IEEE_DIV": "1"})
    assert len(fast["keys"]) > 500 and (fast["normal"][..., 0] != -2.0).sum() > 5000
    for name in fast.files:
        a, b = fast[name], ieee[name]
        assert a.shape == b.shape and a.tobytes() == b.tobytes(), name
