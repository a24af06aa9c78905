"""Runs the REFERENCE build (oracle/_ref: the reference's own sources, single-threaded) over frames of a synthetic stream in a
fresh process and saves what it produced (a fresh process per volume size: the reference keeps `static const float epsilon`
of the first map it raycasts, ray_iterator.hpp:63).
Usage: python _ref_frames_worker.py sdf|ofusion size dim W H mu plane|room first,step,count out.npz [n_frames]"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import numpy as np

from oracle_lib import OFUSION, SDF, Oracle
from supereight_b200 import synth

name, size, dim, W, H, mu, scene = sys.argv[1], int(sys.argv[2]), float(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), float(sys.argv[6]), sys.argv[7]
first, step, count = (int(v) for v in sys.argv[8].split(","))
out = sys.argv[9]
n_frames = int(sys.argv[10]) if len(sys.argv) > 10 else 300
field = SDF if name == "sdf" else OFUSION
k = tuple(v * W / 640.0 for v in synth.DEFAULT_K)
r = Oracle(field, size, dim, W, H, kind="ref_" + name)
r.lib.seo_set_omp_threads(1)           # allocate_level's children_mask_ update is racy otherwise (octree.hpp:843-849)
pose = None
for i in range(count):
    f = first + i * step
    if scene == "plane":
        d, pose = synth.planar_sweep(f, dim, W, H, k, noise_mm=2.0, dropout=0.01)
    else:
        d, pose = synth.box_room(f, dim, W, H, k, n_frames=n_frames, noise_mm=2.0, dropout=0.01)
    assert r.preprocess(d) == 0
    r.integrate(pose, k, mu, f)
r.raycast(pose, k, mu)
keys, coords, active, data = r.blocks_sorted()
codes, side, mask, values = r.nodes_sorted()
np.savez(out, keys=keys, coords=coords, active=active, data=data, codes=codes, side=side, mask=mask, values=values,
         vertex=r.vertex(), normal=r.normal(), image=r.render_volume(pose, k, mu, 0.75 * mu, False), depth_image=r.render_depth(), pose=pose)
