"""Runs the reference build (oracle/_ref) and the oracle side by side in a fresh process and prints mismatch counts as JSON.
A fresh process per volume size because the reference keeps `static const float epsilon` of the FIRST map it sees
(ray_iterator.hpp:63).  Usage: python _ref_compare_worker.py sdf|ofusion size dim W H frames"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import numpy as np

import mc_table_ref
from oracle_lib import OFUSION, SDF, Oracle
from supereight_b200 import synth

name, size, dim, W, H, frames = sys.argv[1], int(sys.argv[2]), float(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
field = SDF if name == "sdf" else OFUSION
mu = 0.1 if field == SDF else 0.008
k = tuple(v * W / 640.0 for v in synth.DEFAULT_K)
o, r = Oracle(field, size, dim, W, H), Oracle(field, size, dim, W, H, kind="ref_" + name)
r.lib.seo_set_omp_threads(1)           # allocate_level's children_mask_ update and the ICP reduction are racy / unordered otherwise
res = {}


def bits(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    if a.shape != b.shape:
        return -1
    if a.dtype == np.float32:
        return int(np.count_nonzero(a.view(np.uint32) != b.view(np.uint32)))
    return int(np.count_nonzero(a != b))


def compare_maps(tag):
    ok, oc, oa, od = o.blocks_sorted(); rk, rc, ra, rd = r.blocks_sorted()
    res[tag + "block_keys"] = bits(ok, rk)
    if res[tag + "block_keys"] == 0:
        res[tag + "block_coords"] = bits(oc, rc); res[tag + "block_active"] = bits(oa, ra)
        res[tag + "block_x"] = bits(od["x"], rd["x"]); res[tag + "block_y"] = bits(od["y"], rd["y"])
    on, rn = o.nodes_sorted(), r.nodes_sorted()
    res[tag + "node_codes"] = bits(on[0], rn[0])
    if res[tag + "node_codes"] == 0:
        res[tag + "node_side"] = bits(on[1], rn[1]); res[tag + "node_mask"] = bits(on[2], rn[2])
        res[tag + "node_x"] = bits(on[3]["x"], rn[3]["x"]); res[tag + "node_y"] = bits(on[3]["y"], rn[3]["y"])
    res[tag + "n_blocks"] = len(ok)


poses = []
for f in range(frames):
    d, pose = synth.box_room(f * 7, dim, W, H, k, noise_mm=2.0, dropout=0.01)
    poses.append(pose)
    assert o.preprocess(d) == 0 and r.preprocess(d) == 0
    o.integrate(pose, k, mu, f); r.integrate(pose, k, mu, f)
    o.raycast(pose, k, mu); r.raycast(pose, k, mu)
compare_maps("")
res["depth"] = bits(o.depth(), r.depth())
res["vertex"] = bits(o.vertex(), r.vertex()); res["normal"] = bits(o.normal(), r.normal())
res["hits"] = int((o.normal()[..., 0] != -2).sum())
res["render_reuse"] = bits(o.render_volume(poses[-1], k, mu, 0.75 * mu, False), r.render_volume(poses[-1], k, mu, 0.75 * mu, False))
res["render_view"] = bits(o.render_volume(poses[0], k, mu, 0.75 * mu, True), r.render_volume(poses[0], k, mu, 0.75 * mu, True))
res["render_depth"] = bits(o.render_depth(), r.render_depth())

# se_core level: get / interp / grad at random positions, the blocks a ray visits
rng = np.random.default_rng(3)
keys, coords, _, _ = o.blocks_sorted()
pts = (coords[rng.integers(0, len(coords), 300)] + rng.uniform(-1, 9, (300, 3))).astype(np.float32)
bad_get = bad_interp = bad_grad = 0
for p in pts:
    ip = [int(v) for v in np.clip(p, 0, size - 1)]
    bad_get += o.get_fine(*ip) != r.get_fine(*ip)
    bad_get += o.get(*ip) != r.get(*ip)
    if (p >= 1).all() and (p < size - 2).all():
        bad_interp += np.float32(o.interp(*p)).view(np.uint32) != np.float32(r.interp(*p)).view(np.uint32)
        bad_grad += bits(o.grad(*p), r.grad(*p))
res["get"], res["interp"], res["grad"] = int(bad_get), int(bad_interp), int(bad_grad)
bad_ray = 0
for _ in range(40):
    origin = poses[-1][:3, 3] + rng.uniform(-0.05, 0.05, 3).astype(np.float32)
    direction = rng.normal(size=3).astype(np.float32); direction /= np.linalg.norm(direction)
    ko, to = o.ray_blocks(origin, direction, 0.4, 4.0)
    kr, tr = r.ray_blocks(origin, direction, 0.4, 4.0)
    bad_ray += (len(ko) != len(kr)) or not np.array_equal(ko, kr) or bits(to, tr) != 0
res["ray_blocks"] = int(bad_ray)

# N1: bilateral filter, pyramid, one ICP run from a perturbed pose
d, pose = synth.box_room((frames - 1) * 7 + 2, dim, W, H, k, noise_mm=2.0, dropout=0.01)
o.preprocess(d); r.preprocess(d)
o.filter_depth(True, 3); r.filter_depth(True, 3)
start = poses[-1].copy(); start[0, 3] += 0.004
po, oko = o.track(start, poses[-1], k, 1e-5, [10, 5, 4])
pr, okr = r.track(start, poses[-1], k, 1e-5, [10, 5, 4])
for lvl in range(3):
    for nm, a, b in zip(("depth", "vertex", "normal"), o.pyramid(lvl), r.pyramid(lvl)):
        res[f"pyramid{lvl}_{nm}"] = bits(a, b)
to_, ro = o.tracking_data(); tr_, rr = r.tracking_data()
res["track_result"] = bits(to_["result"], tr_["result"]); res["track_error"] = bits(to_["error"], tr_["error"]); res["track_J"] = bits(to_["J"], tr_["J"])
res["reduction"] = bits(ro, rr); res["pose"] = bits(po, pr); res["tracked"] = int(oko != okr)

# N4: the reference's own edge table vs the classic list the library ships: the same triangles
mo = o.marching_cube(mc_table_ref.classic_table()); mr = r.marching_cube(mc_table_ref.classic_table())
res["mesh_triangles"] = [int(len(mo)), int(len(mr))]
vo = np.unique(mo.reshape(-1, 3).view(np.uint32), axis=0); vr = np.unique(mr.reshape(-1, 3).view(np.uint32), axis=0)
res["mesh_vertices_differ"] = int(vo.shape != vr.shape or not np.array_equal(vo, vr))
so = {tuple(np.roll(t, -min(range(3), key=lambda i: tuple(t[i])), axis=0).ravel()) for t in mo.view(np.uint32).astype(np.int64)}
sr = {tuple(np.roll(t, -min(range(3), key=lambda i: tuple(t[i])), axis=0).ravel()) for t in mr.view(np.uint32).astype(np.int64)}
res["mesh_identical_triangles_frac"] = len(so & sr) / max(1, len(sr))
res["mesh_triangles_only_in_oracle"] = len(so - sr)
print(json.dumps(res))
