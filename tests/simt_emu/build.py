"""Build the CPU test tier's copy of the library: the product sources compiled with g++ against the stand-in CUDA runtime.

TEST INFRASTRUCTURE ONLY (see include/cuda_runtime.h).  What is compiled is supereight_b200/csrc/* as it stands, with
exactly these textual rewrites, all of launch / declaration SYNTAX that g++ cannot parse:

  kernel<<<grid, block, smem, stream>>>(args)      ->  simt::launch(kernel, grid, block, smem, stream, args)
  extern __shared__ __align__(N) T name[];         ->  T* name = (T*)simt::dynamic_smem();
  __noinline__                                     ->  __attribute__((noinline))
  se_ptx.cuh (inline PTX)                          ->  tests/simt_emu/se_ptx_emu.cuh (the same wrappers in C++)

Output: tests/simt_emu/_build/libse_b200_simt.so, loaded by the tests through SE_B200_LIB (never by the package itself).
Compiled with -ffp-contract=off like every host file of the project: the arithmetic contract (no FMA contraction) holds.
"""
from __future__ import annotations

import os
import re
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "supereight_b200", "csrc")
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(BUILD, "libse_b200_simt.so")

_LAUNCH = re.compile(r"(\bk_\w+(?:<[^<>;]*>)?)\s*<<<(.*?)>>>\s*\(")
_EXTERN_SMEM = re.compile(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?([\w ]+?)\s+(\w+)\[\];")


def rewrite(text: str) -> str:
    text = _LAUNCH.sub(lambda m: f"simt::launch({m.group(1)}, {m.group(2)}, ", text)
    text = text.replace("__noinline__", "__attribute__((noinline))")      # (libstdc++ spells attributes __noinline__ too)
    text = _EXTERN_SMEM.sub(lambda m: f"{m.group(1)}* {m.group(2)} = ({m.group(1)}*)simt::dynamic_smem();", text)
    return text


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh")))


def up_to_date() -> bool:
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in sources()] + [os.path.join(ROOT, "include", "se_b200.h"), __file__,
                                                         os.path.join(HERE, "se_ptx_emu.cuh")]
    for d, _, files in os.walk(os.path.join(HERE, "include")):
        deps += [os.path.join(d, f) for f in files]
    return all(os.path.getmtime(d) <= t for d in deps)


def build(force: bool = False, defines=(), suffix: str = "") -> str:
    """defines / suffix: a build of a compile-time variant (e.g. ("-DSE_INT_STAGE_SLICES=4",), "_half") next to the default one"""
    if suffix:
        return _build(LIB.replace(".so", suffix + ".so"), tuple(defines))
    if not force and up_to_date():
        return LIB
    return _build(LIB, tuple(defines))


def _build(lib: str, defines) -> str:
    src = os.path.join(BUILD, "supereight_b200", "csrc")      # same depth as the original: "../../include/se_b200.h" resolves
    os.makedirs(src, exist_ok=True)
    os.makedirs(os.path.join(BUILD, "include"), exist_ok=True)
    shutil.copy(os.path.join(ROOT, "include", "se_b200.h"), os.path.join(BUILD, "include", "se_b200.h"))
    for f in sources():
        if f == "se_ptx.cuh":
            continue
        with open(os.path.join(CSRC, f)) as fh:
            text = rewrite(fh.read())
        if "<<<" in text or "asm" in re.sub(r"//.*", "", text):
            raise RuntimeError(f"{f}: a launch or an asm statement the rewrite rules do not cover")
        with open(os.path.join(src, f), "w") as fh:
            fh.write(text)
    shutil.copy(os.path.join(HERE, "se_ptx_emu.cuh"), os.path.join(src, "se_ptx.cuh"))
    cmd = ["g++", "-std=c++17", "-O2", "-g1", "-fPIC", "-shared", "-ffp-contract=off", "-mfma", "-fno-strict-aliasing",
           "-Wno-attributes", "-Wno-unused-value", *defines, "-I", os.path.join(HERE, "include"), "-x", "c++", os.path.join(src, "se_b200.cu"),
           "-o", lib]
    subprocess.run(cmd, check=True)
    return lib


if __name__ == "__main__":
    print(build(force=True))
