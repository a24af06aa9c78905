#!/bin/bash
# Builds the opt-in compile-time experiments next to the default library (build container, no GPU needed):
#   ab_libs/stage4.so   -DSE_INT_STAGE_SLICES=4     integrate: half-block TMA stages, 4 CTAs / SM   (csrc/se_integrate_staged.cuh)
#   ab_libs/nbhd.so     -DSE_GRAD_NBHD              raycast: gradient's 2x2x2 directory cells from one base index (se_map.cuh)
#   ab_libs/uni.so      -DSE_RAY_UNIFORMS           raycast: ray-independent set-up quantities computed on the host (RayWalk::init_pre)
#   ab_libs/ray.so      the two raycast ones together
#   ab_libs/all.so      all three
#   ab_libs/t32.so      -DSE_RAY_TILE_32X1          ray kernels: 32x1 pixel tiles per warp (128 B store segments; for the render-target A/B)
# ab_libs/*.so are git-ignored and travel to the GPU box with the snapshot; scripts/ab_variants.sh A/Bs them there.
set -e
cd "$(dirname "$0")/.."
mkdir -p ab_libs
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC,-ffp-contract=off -shared"
nvcc $FLAGS -DSE_INT_STAGE_SLICES=4 -o ab_libs/stage4.so supereight_b200/csrc/se_b200.cu &
nvcc $FLAGS -DSE_GRAD_NBHD -o ab_libs/nbhd.so supereight_b200/csrc/se_b200.cu &
nvcc $FLAGS -DSE_RAY_UNIFORMS -o ab_libs/uni.so supereight_b200/csrc/se_b200.cu &
nvcc $FLAGS -DSE_GRAD_NBHD -DSE_RAY_UNIFORMS -o ab_libs/ray.so supereight_b200/csrc/se_b200.cu &
nvcc $FLAGS -DSE_INT_STAGE_SLICES=4 -DSE_GRAD_NBHD -DSE_RAY_UNIFORMS -o ab_libs/all.so supereight_b200/csrc/se_b200.cu &
nvcc $FLAGS -DSE_RAY_TILE_32X1 -o ab_libs/t32.so supereight_b200/csrc/se_b200.cu &
wait
ls -la ab_libs
