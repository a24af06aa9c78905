"""What the OFusion path actually differs by from the oracle on the parity scenarios (GPU): hit-mask flips, vertex / normal
deviations, occupancy bits -- the numbers the tolerances in tests/test_gpu_parity.py are set from.
Usage: python scripts/ofusion_deviation.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

from oracle_lib import OFUSION, Oracle
from supereight_b200 import Map, synth

K640 = (481.2, 480.0, 320.0, 240.0)
for name, size, dim, mu, W, H, gen, frames, kw in (
        ("room 1024 320x240", 1024, 4.8, 0.008, 320, 240, synth.box_room, range(0, 20, 4), dict(n_frames=300, dropout=0.01)),
        ("room 1024 640x480", 1024, 4.8, 0.008, 640, 480, synth.box_room, [3], dict(n_frames=300, noise_mm=2.0, dropout=0.01)),
        ("plane 512 640x480", 512, 4.8, 0.03, 640, 480, synth.planar_sweep, range(3), dict(noise_mm=2.0, dropout=0.01)),
        ("room 256 160x120", 256, 4.8, 0.03, 160, 120, synth.box_room, range(0, 30, 5), dict(n_frames=60, noise_mm=2.0, dropout=0.01))):
    k = tuple(v * W / 640.0 for v in K640)
    g, o = Map(OFUSION, size, dim, W, H), Oracle(OFUSION, size, dim, W, H)
    for f in frames:
        d, pose = gen(f, dim, W, H, k, **kw)
        o.preprocess(d); o.integrate(pose, k, mu, f)
        g.preprocess(d); g.integrate(pose, k, mu, f)
    o.raycast(pose, k, mu); g.raycast(pose, k, mu)
    gk, _, _, gd = g.blocks_sorted(); ok, _, _, od = o.blocks_sorted()
    gv, gn = g.vertex_normal(); ov, on = o.vertex(), o.normal()
    ghit, ohit = gn[..., 0] != -2, on[..., 0] != -2
    both = ghit & ohit
    xb = np.count_nonzero(gd["x"].view(np.uint32) != od["x"].view(np.uint32))
    print(f"{name}: keys equal {np.array_equal(gk, ok)}  blocks {len(gk)}  occupancy words differing {xb} of {gd['x'].size}  max rel {np.abs(gd['x'] - od['x']).max():.3g}"
          f"  hits {int(ghit.sum())}  flips {int((ghit != ohit).sum())}  vertex bits differing {int((gv.view(np.uint32) != ov.view(np.uint32)).sum())}"
          f"  max |dv| {np.abs(gv[both] - ov[both]).max():.3g}  normal bits differing {int((gn.view(np.uint32) != on.view(np.uint32)).sum())}  max |dn| {np.abs(gn[both] - on[both]).max():.3g}")
