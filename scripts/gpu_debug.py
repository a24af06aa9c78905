"""Run a short synthetic sequence through the CUDA path and the oracle, print parity + timings.
Usage: python scripts/gpu_debug.py [sdf|ofusion] [size] [frames] [scene]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import oracle_lib
from parity_utils import compare_blocks, compare_images, compare_nodes
from supereight_b200 import Map, synth

field = {"sdf": 0, "ofusion": 1}[sys.argv[1] if len(sys.argv) > 1 else "sdf"]
size = int(sys.argv[2]) if len(sys.argv) > 2 else 512
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 5
scene = sys.argv[4] if len(sys.argv) > 4 else "plane"
dim = 4.8
mu = 0.1 if field == 0 else 0.008
W, H = 640, 480
k = synth.DEFAULT_K
gen = synth.planar_sweep if scene == "plane" else synth.box_room

g = Map(field, size, dim, W, H)
o = oracle_lib.Oracle(field, size, dim, W, H)
for f in range(frames):
    d, pose = gen(f, dim)
    t0 = time.time(); o.preprocess(d); o.integrate(pose, k, mu, f); t1 = time.time()
    g.preprocess(d); g.integrate(pose, k, mu, f); g.sync()
    print(f"frame {f}: oracle integrate {1e3*(t1-t0):.1f} ms | gpu alloc {g.elapsed_ms('alloc'):.3f} fuse {g.elapsed_ms('fuse'):.3f} ms", g.counters())
    print("  blocks", compare_blocks(g, o))
    print("  nodes ", compare_nodes(g, o))
    gc_, gs_, gm_, gv_ = g.nodes_sorted(); oc_, os_, om_, ov_ = o.nodes_sorted()
    if len(gc_) == len(oc_):
        bad = np.argwhere(gv_["x"].view(np.uint32) != ov_["x"].view(np.uint32))
        for n_, s_ in bad[:6]:
            print(f"    node code {int(gc_[n_]):#x} side {int(gs_[n_])} slot {s_}: gpu {float(gv_['x'][n_, s_])!r} oracle {float(ov_['x'][n_, s_])!r} y {float(gv_['y'][n_, s_])!r}")
    t0 = time.time(); o.raycast(pose, k, mu); t1 = time.time()
    g.raycast(pose, k, mu); gv, gn = g.vertex_normal()
    print(f"  raycast oracle {1e3*(t1-t0):.1f} ms gpu {g.elapsed_ms('raycast'):.3f} ms", compare_images(gv, gn, o.vertex(), o.normal()))
    gi = g.render_volume(pose, k, mu, 0.75 * mu, False)
    oi = o.render_volume(pose, k, mu, 0.75 * mu, False)
    print("  render reuse mismatch", int(np.count_nonzero(gi != oi)), "render ms", g.elapsed_ms('render'))
    if f == frames - 1:
        view = gen(f + 7, dim)[1]
        gi = g.render_volume(view, k, mu, 0.75 * mu, True)
        oi = o.render_volume(view, k, mu, 0.75 * mu, True)
        print("  render re-raycast mismatch", int(np.count_nonzero(gi != oi)), "render ms", g.elapsed_ms('render'))
        print("  render depth mismatch", int(np.count_nonzero(g.render_depth() != o.render_depth())))
print("launches", g.launch_count())
