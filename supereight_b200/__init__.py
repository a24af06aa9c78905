"""supereight_b200 -- the supereight per-frame dense-SLAM hot path, B200-native.

The product is the CUDA library `libse_b200.so` (C ABI in include/se_b200.h, kernels in
supereight_b200/csrc/) and the C++ `DenseSLAMSystem` shim in supereight_b200/host/.  This Python
package is only the ctypes binding that the tests and bench.py use to call the C ABI, plus the
synthetic depth-stream generator.  There is no CPU fallback: loading fails loudly when the
library is missing, and every call fails when there is no CUDA device.
"""
from .capi import Map, SE_B200_SDF, SE_B200_OFUSION, lib_path, load_library, mc_table, SeB200Error  # noqa: F401
