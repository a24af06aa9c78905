"""Worker of tests/test_simt_emu.py: runs a short sequence through whatever library SE_B200_LIB names (the test sets it to
the fiber-executor build of the product sources, tests/simt_emu) and dumps everything comparable to an .npz.  A separate
process per run, because the library reads its SE_B200_* switches once."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(field_name, out_path):
    from supereight_b200 import Map, SE_B200_OFUSION, SE_B200_SDF, synth
    assert "simt_emu" in os.environ.get("SE_B200_LIB", ""), "this worker is for the emulated build only"
    field = SE_B200_SDF if field_name == "sdf" else SE_B200_OFUSION
    mu = 0.1 if field_name == "sdf" else 0.008
    W, H, size, dim = 160, 120, 256, 4.8
    k = (120.3, 120.0, 80.0, 60.0)
    g = Map(field, size, dim, W, H, device=0)
    pose = None
    for f in (0, 3, 6, 9):
        d, pose = synth.box_room(f, dim, W, H, k, n_frames=60, noise_mm=2.0, dropout=0.01, seed=7)
        g.preprocess(d); g.integrate(pose, k, mu, f)
    g.raycast(pose, k, mu)
    keys, coords, active, data = g.blocks_sorted(True)
    ncodes, nside, nmask, nval = g.nodes_sorted()
    v, n = g.vertex_normal()
    img = g.render_volume(pose, k, mu, 0.75 * mu, False)
    np.savez(out_path, keys=keys, coords=coords, active=active, x=data["x"], y=data["y"], ncodes=ncodes, nmask=nmask,
             nx=nval["x"], ny=nval["y"], vertex=v, normal=n, img=img)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
