"""Comparison helpers shared by the GPU parity tests and scripts/gpu_debug.py."""
from __future__ import annotations

import numpy as np


def rel_err(a, b, floor=1e-6):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), floor)


def compare_blocks(gpu_map, oracle, with_data=True):
    """Returns a dict of mismatch statistics between the device map and the oracle map."""
    gk, gc, ga, gd = gpu_map.blocks_sorted(with_data)
    ok, oc, oa, od = oracle.blocks_sorted(with_data)
    out = dict(n_gpu=len(gk), n_oracle=len(ok))
    out["keys_equal"] = bool(len(gk) == len(ok) and np.array_equal(gk, ok))
    if not out["keys_equal"]:
        out["only_gpu"] = int(len(np.setdiff1d(gk, ok)))
        out["only_oracle"] = int(len(np.setdiff1d(ok, gk)))
        return out
    out["coords_equal"] = bool(np.array_equal(gc, oc))
    out["active_mismatch"] = int(np.count_nonzero(ga != oa))
    if with_data:
        gx, ox = gd["x"], od["x"]
        gy, oy = gd["y"].astype(np.float64), od["y"].astype(np.float64)
        out["x_bit_mismatch"] = int(np.count_nonzero(gx.view(np.uint32) != ox.view(np.uint32)))
        out["y_mismatch"] = int(np.count_nonzero(gy != oy))
        out["x_max_rel"] = float(rel_err(gx, ox).max()) if gx.size else 0.0
        out["x_max_abs"] = float(np.abs(gx.astype(np.float64) - ox).max()) if gx.size else 0.0
    return out


def compare_nodes(gpu_map, oracle):
    gc, gs, gm, gv = gpu_map.nodes_sorted()
    oc, os_, om, ov = oracle.nodes_sorted()
    out = dict(n_gpu=len(gc), n_oracle=len(oc))
    out["codes_equal"] = bool(len(gc) == len(oc) and np.array_equal(gc, oc))
    if not out["codes_equal"]:
        return out
    out["side_equal"] = bool(np.array_equal(gs, os_))
    out["mask_equal"] = bool(np.array_equal(gm, om))
    out["x_bit_mismatch"] = int(np.count_nonzero(gv["x"].view(np.uint32) != ov["x"].view(np.uint32)))
    out["y_mismatch"] = int(np.count_nonzero(gv["y"].astype(np.float64) != ov["y"].astype(np.float64)))
    out["x_max_rel"] = float(rel_err(gv["x"], ov["x"]).max())
    return out


def compare_images(gv, gn, ov, on):
    """vertex/normal maps: hit-mask equality, bit mismatches and max relative error."""
    out = {}
    ghit = gn[..., 0] != -2.0
    ohit = on[..., 0] != -2.0
    out["hits_gpu"] = int(ghit.sum())
    out["hits_oracle"] = int(ohit.sum())
    out["hit_mask_mismatch"] = int(np.count_nonzero(ghit != ohit))
    both = ghit & ohit
    out["vertex_bit_mismatch"] = int(np.count_nonzero(gv.view(np.uint32) != ov.view(np.uint32)))
    out["normal_bit_mismatch"] = int(np.count_nonzero(gn.view(np.uint32) != on.view(np.uint32)))
    if both.any():
        out["vertex_max_rel"] = float(rel_err(gv[both], ov[both], floor=1e-3).max())
        out["normal_max_abs"] = float(np.abs(gn[both].astype(np.float64) - on[both]).max())
    return out
