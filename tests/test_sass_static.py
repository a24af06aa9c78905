"""Static checks of the built library's device code (no GPU needed: cuobjdump reads the embedded sm_100a cubin).

They pin what the ncu captures under profiles/ show, so that a change which silently loses it fails in the CPU tier:
the register budgets the occupancy targets rest on, the TMA / mbarrier pipeline of the integrate kernel, programmatic
dependent launch in every per-frame kernel, packed fp32, the warp-level de-duplication and lock-free insert of the
allocation pass -- and that there is no tensor-core instruction anywhere (nothing on this path is a contraction)."""
import collections
import re
import shutil
import subprocess

import pytest

from supereight_b200 import capi

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")


def demangle(names):
    if shutil.which("c++filt") is None:
        return list(names)
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True, check=True).stdout
    return out.splitlines()


@pytest.fixture(scope="module")
def resources():
    """kernel (demangled, without the parameter list) -> dict(REG=, STACK=, SHARED=, ...)"""
    text = subprocess.run(["cuobjdump", "-res-usage", capi.lib_path()], capture_output=True, text=True, check=True).stdout
    lines = text.splitlines()
    names, usage = [], []
    for i, ln in enumerate(lines):
        m = re.match(r"\s*Function (\S+):", ln)
        if m and i + 1 < len(lines):
            names.append(m.group(1))
            usage.append({k: int(v) for k, v in re.findall(r"(\w+)(?:\[0\])?:(\d+)", lines[i + 1])})
    res = {}
    for n, u in zip(demangle(names), usage):
        res[re.sub(r"^void ", "", n).split("(")[0].replace("se_b200::", "")] = u
    return res


@pytest.fixture(scope="module")
def sass():
    """kernel (mangled) -> opcode counter"""
    text = subprocess.run(["cuobjdump", "-sass", capi.lib_path()], capture_output=True, text=True, check=True).stdout
    per, cur = {}, None
    for ln in text.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = per.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m and cur is not None:
            cur[m.group(1)] += 1
    assert per, "no SASS in the library: was it built for sm_100a?"
    return per


def kernels(sass, substr):
    hit = {k: v for k, v in sass.items() if substr in k}
    assert hit, f"no kernel named *{substr}*"
    return hit


def has(counter, prefix):
    return sum(v for k, v in counter.items() if k.startswith(prefix))


def test_built_for_sm_100a():
    text = subprocess.run(["cuobjdump", "-lelf", capi.lib_path()], capture_output=True, text=True, check=True).stdout
    assert "sm_100a" in text, text


def test_register_budgets(resources):
    # raycast: 8 CTAs x 128 threads per SM need <= 64 registers; a few spilled words are accepted, a frame of kilobytes is not
    # (template arguments: field, DENSE walk, COUNT, SHADE)
    ray = [n for n in resources if n.startswith("k_raycast<")]
    assert len(ray) >= 8, ray
    for name in ray:      # (no frame: the out-of-line slow paths take the map by value, not the kernel parameter by reference)
        assert resources[name]["REG"] <= 64 and resources[name]["STACK"] <= 64, (name, resources[name])
    # integrate (SDF): 4 CTAs x 256 threads per SM -> <= 64 registers, no stack; its 32 KiB of half-block stage buffers are dynamic smem
    for name in ("k_integrate_sdf<true>", "k_integrate_sdf<false>"):
        assert resources[name]["REG"] <= 64 and resources[name]["STACK"] == 0, (name, resources[name])
    # integrate (OFusion): 4 CTAs x 256 threads -> <= 64
    # (the plain-operator fallback may spill a word or two: it carries the in-kernel list builder's frustum test as well)
    for name, stack in (("k_integrate_ofusion<true>", 0), ("k_integrate_ofusion<false>", 16)):
        assert resources[name]["REG"] <= 64 and resources[name]["STACK"] <= stack, (name, resources[name])
    # allocation: 5 CTAs x 256 threads -> <= 51, the per-thread block lists live in (static) shared memory
    assert resources["k_alloc_sdf<SdfVoxel>"]["REG"] <= 51 and resources["k_alloc_sdf<SdfVoxel>"]["STACK"] == 0
    assert resources["k_alloc_ofusion<OfuVoxel>"]["REG"] <= 64 and resources["k_alloc_ofusion<OfuVoxel>"]["STACK"] == 0


def test_integrate_pipeline_is_tma_plus_mbarrier_and_packed_fp32(sass):
    for name, ops in kernels(sass, "k_integrate_sdfILb1").items():
        assert has(ops, "UBLKCP") >= 2, (name, "cp.async.bulk (TMA bulk copy) missing")
        assert has(ops, "SYNCS") >= 3, (name, "mbarrier operations missing")
        assert has(ops, "FFMA2") > 0 and has(ops, "FMUL2") > 0 and has(ops, "FADD2") > 0, (name, "packed fp32 missing")
        assert has(ops, "STG.E.128") > 0 and has(ops, "LDS.128") > 0, (name, "vectorised payload access missing")
    # the plain-operator instantiation keeps the pipeline, not the packed arithmetic
    for name, ops in kernels(sass, "k_integrate_sdfILb0").items():
        assert has(ops, "UBLKCP") >= 2 and has(ops, "FFMA2") == 0, name


def test_every_per_frame_kernel_uses_programmatic_dependent_launch(sass):
    for k in ("k_mm2meters", "k_alloc_sdf", "k_alloc_ofusion", "k_alloc_first_key_chain", "k_integrate_sdf",
              "k_integrate_ofusion", "k_raycast", "k_render_shade", "k_render_volume"):
        for name, ops in kernels(sass, k).items():
            assert has(ops, "ACQBULK") >= 1, (name, "griddepcontrol.wait missing")
            assert has(ops, "PREEXIT") >= 1, (name, "griddepcontrol.launch_dependents missing")


def test_allocation_pass_primitives(sass):
    for k in ("k_alloc_sdf", "k_alloc_ofusion"):
        for name, ops in kernels(sass, k).items():
            assert has(ops, "MATCH.ANY") >= 1, (name, "__match_any_sync de-duplication missing")
            assert has(ops, "VOTE") >= 1, name
    lib = collections.Counter()
    for ops in sass.values():
        lib.update(ops)
    assert any(k.startswith("ATOMG") and ".CAS" in k for k in lib), "atomicCAS on the child slots missing"


def test_no_tensor_core_instructions(sass):
    for name, ops in sass.items():
        for op in ops:
            assert not op.startswith(("HMMA", "IMMA", "DMMA", "QMMA", "OMMA", "UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCMMA", "HGMMA")), (name, op)
