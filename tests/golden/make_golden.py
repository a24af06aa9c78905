"""Generates the golden fixtures in this directory.  Run from the repo root in the BUILD container:

    python tests/golden/make_golden.py

* bspline_lut.sha256  : checksum of the float32 image of the reference's own B-spline table
  (/root/reference/se_denseslam/src/bfusion/bspline_lookup.cc:37) -- needs /root/reference.
* seq_*.npz           : inputs (uint16 depth frames, poses, intrinsics) and outputs (sorted block keys,
  voxel payloads, node codes/values, vertex/normal maps, rendered images) of short sequences run through
  THE REFERENCE ITSELF: oracle/_ref/libse_ref_<field>.so is the reference's own DenseSLAMSystem.cpp (and every
  header it includes) compiled where it lies under /root/reference against the stand-in Eigen / Sophus headers of
  oracle/ref_standin (the image has neither library; see oracle/Makefile and ref_standin/Eigen/Dense), run
  single-threaded (its allocate_level races on children_mask_ otherwise).  The script then checks that the oracle
  (oracle/, parity build) reproduces every array bit for bit before writing the file, so the fixtures pin the
  oracle to the reference's code, and the GPU tests get committed vectors that need neither to be rebuilt.
  Without /root/reference (no oracle/_ref) the script refuses to regenerate them.
"""
import hashlib
import subprocess
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_lib  # noqa: E402
from supereight_b200 import synth  # noqa: E402

SEQS = {
    # name: field, size, dim, W, H, mu, scene, frames, noise
    "seq_sdf_plane": (oracle_lib.SDF, 128, 2.4, 80, 60, 0.05, "plane", 4, 2.0),
    "seq_sdf_room": (oracle_lib.SDF, 128, 2.4, 80, 60, 0.05, "room", 4, 0.0),
    "seq_ofusion_room": (oracle_lib.OFUSION, 128, 2.4, 80, 60, 0.008, "room", 4, 0.0),
}


def lut_checksum():
    ref = "/root/reference/se_denseslam/src/bfusion/bspline_lookup.cc"
    if not os.path.exists(ref):
        print("reference tree not present: keeping the committed bspline_lut.sha256")
        return
    src = open(ref).read()
    body = src[src.index("{", src.index("bspline_lookup[")):]
    toks = re.findall(r"[-+]?(?:\d+\.\d*|\.\d+|\d+)(?:[eE][-+]?\d+)?", body)
    table = np.array([float(v) for v in toks[:1000]], dtype=np.float32)
    with open(os.path.join(HERE, "bspline_lut.sha256"), "w") as f:
        f.write(hashlib.sha256(table.tobytes()).hexdigest() + "  float32[1000] of se_denseslam/src/bfusion/bspline_lookup.cc:37 (see make_golden.py)\n")


def run_pipeline(o, name, depth, poses, k):
    field, size, dim, W, H, mu, scene, frames, noise = SEQS[name]
    for f in range(frames):
        o.preprocess(depth[f])
        o.integrate(poses[f], k, mu, f)
    o.raycast(poses[-1], k, mu)
    keys, coords, active, data = o.blocks_sorted()
    codes, side, mask, values = o.nodes_sorted()
    return dict(
        block_keys=keys, block_coords=coords, block_active=active, block_x=data["x"], block_y=data["y"],
        node_codes=codes, node_side=side, node_mask=mask, node_x=values["x"], node_y=values["y"],
        vertex=o.vertex(), normal=o.normal(),
        render_reuse=o.render_volume(poses[-1], k, mu, 0.75 * mu, False),
        render_view=o.render_volume(poses[0], k, mu, 0.75 * mu, True),
        render_depth=o.render_depth(),
    )


def run_sequence(name):
    field, size, dim, W, H, mu, scene, frames, noise = SEQS[name]
    k = np.array([v * W / 640.0 for v in synth.DEFAULT_K], np.float32)
    gen = synth.planar_sweep if scene == "plane" else synth.box_room
    depth = np.empty((frames, H, W), np.uint16)
    poses = np.empty((frames, 4, 4), np.float32)
    for f in range(frames):
        if scene == "plane":
            depth[f], poses[f] = gen(f * 5, dim, W, H, tuple(k), noise_mm=noise, dropout=0.02)
        else:
            depth[f], poses[f] = gen(f * 3, dim, W, H, tuple(k), n_frames=60, noise_mm=noise, dropout=0.02)
    ref = oracle_lib.Oracle(field, size, dim, W, H, kind="ref_sdf" if field == oracle_lib.SDF else "ref_ofusion")
    ref.lib.seo_set_omp_threads(1)
    out = run_pipeline(ref, name, depth, poses, k)
    mine = run_pipeline(oracle_lib.Oracle(field, size, dim, W, H), name, depth, poses, k)
    for key, want in out.items():
        got = mine[key]
        same = got.shape == want.shape and (np.array_equal(got.view(np.uint32), want.view(np.uint32)) if want.dtype == np.float32 else np.array_equal(got, want))
        assert same, f"{name}: the oracle differs from the reference build in {key}"
    out.update(field=field, size=size, dim=np.float32(dim), W=W, H=H, mu=np.float32(mu), k=k, depth=depth, poses=poses)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "(reference build == oracle) blocks", len(out["block_keys"]), "nodes", len(out["node_codes"]),
          "hits", int((out["normal"][..., 0] != -2).sum()), "bytes", os.path.getsize(os.path.join(HERE, name + ".npz")))


if __name__ == "__main__":
    subprocess.run(["make", "-s", "-C", oracle_lib.ORACLE_DIR], check=True)      # the oracle and, where /root/reference exists, oracle/_ref
    lut_checksum()
    if not oracle_lib.have_reference_build():
        sys.exit("oracle/_ref is missing (needs /root/reference): the committed seq_*.npz are kept")
    for n in SEQS:
        run_sequence(n)
