"""ctypes binding of the CPU oracle (oracle/_build/liboracle_*.so).  Test infrastructure only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

SDF, OFUSION = 0, 1
SDF_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4")])
OFUSION_DTYPE = np.dtype([("x", "<f4"), ("_pad", "<f4"), ("y", "<f8")])   # {float x; double y;} -> 16 B
FIELD_DTYPE = {SDF: SDF_DTYPE, OFUSION: OFUSION_DTYPE}

_libs = {}

u64p = C.POINTER(C.c_uint64)
f32p = C.POINTER(C.c_float)
i32p = C.POINTER(C.c_int)
u8p = C.POINTER(C.c_uint8)


def build():
    """the oracle libraries only; oracle/_ref (minutes of compilation) is built by __graft_entry__.build() / `make -C oracle`"""
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "oracle"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def lib_file(kind: str) -> str:
    """"parity" / "fast": the oracle (oracle/_build); "ref_sdf", "ref_ofusion" (+ "_fast"): the reference's own sources built
    against the stand-in Eigen / Sophus headers (oracle/_ref, see oracle/Makefile) behind the same seo_* entry points."""
    if kind.startswith("ref_"):
        return os.path.join(ORACLE_DIR, "_ref", f"libse_{kind}.so")
    return os.path.join(ORACLE_DIR, "_build", f"liboracle_{kind}.so")


def have_reference_build() -> bool:
    return all(os.path.exists(lib_file(k)) for k in ("ref_sdf", "ref_ofusion"))


def _set(lib, name, attr, value):
    if hasattr(lib, name):              # the reference build exports the pipeline-level subset only
        setattr(getattr(lib, name), attr, value)


def load(kind: str = "parity"):
    if kind in _libs:
        return _libs[kind]
    path = lib_file(kind)
    if not os.path.exists(path):
        build()
    lib = C.CDLL(path)
    _set(lib, "seo_morton_encode", "restype", C.c_uint64)
    _set(lib, "seo_morton_encode", "argtypes", [C.c_int] * 3)
    _set(lib, "seo_morton_decode", "argtypes", [C.c_uint64, i32p])
    _set(lib, "seo_level_mask", "restype", C.c_uint64)
    _set(lib, "seo_key_encode", "restype", C.c_uint64)
    _set(lib, "seo_key_encode", "argtypes", [C.c_int] * 5)
    _set(lib, "seo_key_descendant", "argtypes", [C.c_uint64, C.c_uint64, C.c_int])
    _set(lib, "seo_key_parent", "restype", C.c_uint64)
    _set(lib, "seo_key_parent", "argtypes", [C.c_uint64, C.c_int])
    _set(lib, "seo_key_child_id", "argtypes", [C.c_uint64, C.c_int, C.c_int])
    _set(lib, "seo_key_far_corner", "argtypes", [C.c_uint64, C.c_int, C.c_int, i32p])
    _set(lib, "seo_key_face_neighbour", "argtypes", [C.c_uint64, C.c_uint, C.c_uint, C.c_uint, i32p])
    _set(lib, "seo_key_exterior_neighbours", "argtypes", [u64p, C.c_uint64, C.c_int, C.c_int])
    _set(lib, "seo_key_siblings", "argtypes", [u64p, C.c_uint64, C.c_int])
    _set(lib, "seo_keys_unique", "argtypes", [u64p, C.c_int])
    _set(lib, "seo_keys_filter_ancestors", "argtypes", [u64p, C.c_int, C.c_int])
    _set(lib, "seo_keys_unique_multiscale", "argtypes", [u64p, C.c_int, C.c_uint])
    _set(lib, "seo_bspline_lut", "restype", C.c_float)
    _set(lib, "seo_create", "restype", C.c_void_p)
    _set(lib, "seo_create", "argtypes", [C.c_int, C.c_int, C.c_float, C.c_int, C.c_int])
    _set(lib, "seo_destroy", "argtypes", [C.c_void_p])
    _set(lib, "seo_preprocess", "argtypes", [C.c_void_p, C.c_void_p, C.c_int, C.c_int])
    _set(lib, "seo_set_depth", "argtypes", [C.c_void_p, C.c_void_p])
    _set(lib, "seo_get_depth", "argtypes", [C.c_void_p, C.c_void_p])
    _set(lib, "seo_integrate", "restype", C.c_uint)
    _set(lib, "seo_integrate", "argtypes", [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_uint])
    _set(lib, "seo_raycast", "argtypes", [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float])
    _set(lib, "seo_render_volume", "argtypes", [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_int])
    _set(lib, "seo_render_depth", "argtypes", [C.c_void_p, C.c_void_p])
    _set(lib, "seo_render_track", "argtypes", [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int])
    _set(lib, "seo_get_vertex", "argtypes", [C.c_void_p, C.c_void_p])
    _set(lib, "seo_get_normal", "argtypes", [C.c_void_p, C.c_void_p])
    _set(lib, "seo_set_vertex_normal", "argtypes", [C.c_void_p, C.c_void_p, C.c_void_p])
    _set(lib, "seo_block_count", "argtypes", [C.c_void_p])
    _set(lib, "seo_node_count", "argtypes", [C.c_void_p])
    _set(lib, "seo_get_blocks_sorted", "argtypes", [C.c_void_p] * 5)
    _set(lib, "seo_get_nodes_sorted", "argtypes", [C.c_void_p] * 5)
    _set(lib, "seo_allocate", "argtypes", [C.c_void_p, C.c_void_p, C.c_int])
    _set(lib, "seo_fetch", "argtypes", [C.c_void_p, C.c_int, C.c_int, C.c_int])
    _set(lib, "seo_fetch_octant", "argtypes", [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int])
    _set(lib, "seo_fetch_octant_code", "restype", C.c_uint64)
    _set(lib, "seo_fetch_octant_code", "argtypes", [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int])
    _set(lib, "seo_get_fine", "argtypes", [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)])
    _set(lib, "seo_get_coarse", "argtypes", [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)])
    _set(lib, "seo_set_voxel", "argtypes", [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double])
    _set(lib, "seo_set_node_value", "argtypes", [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double])
    _set(lib, "seo_interp", "restype", C.c_float)
    _set(lib, "seo_interp", "argtypes", [C.c_void_p, C.c_float, C.c_float, C.c_float])
    _set(lib, "seo_grad", "argtypes", [C.c_void_p, C.c_float, C.c_float, C.c_float, f32p])
    _set(lib, "seo_gather", "argtypes", [C.c_void_p, C.c_int, C.c_int, C.c_int, f32p])
    _set(lib, "seo_ray_blocks", "argtypes", [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_int, C.c_void_p])
    _set(lib, "seo_set_counting", "argtypes", [C.c_void_p, C.c_int])
    _set(lib, "seo_filter_depth", "argtypes", [C.c_void_p, C.c_int, C.c_int])
    _set(lib, "seo_tracking", "restype", C.c_int)
    _set(lib, "seo_tracking", "argtypes", [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_int])
    _set(lib, "seo_get_pyramid", "argtypes", [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p])
    _set(lib, "seo_get_tracking", "argtypes", [C.c_void_p, C.c_void_p, C.c_void_p])
    _set(lib, "seo_marching_cube", "restype", C.c_longlong)
    _set(lib, "seo_marching_cube", "argtypes", [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong])
    _set(lib, "seo_se3_exp", "argtypes", [C.c_void_p, C.c_void_p])
    _set(lib, "seo_solve6", "argtypes", [C.c_void_p, C.c_void_p])
    _set(lib, "seo_reset_counters", "argtypes", [C.c_void_p])
    _set(lib, "seo_get_counters", "argtypes", [C.c_void_p, C.c_void_p])
    _set(lib, "seo_set_omp_threads", "argtypes", [C.c_int])
    _set(lib, "seo_save_map", "argtypes", [C.c_void_p, C.c_char_p])
    _set(lib, "seo_load_map", "restype", C.c_void_p)
    _set(lib, "seo_load_map", "argtypes", [C.c_char_p, C.c_float])
    _set(lib, "seo_map_size", "argtypes", [C.c_void_p])
    _set(lib, "seo_map_dim", "restype", C.c_float)
    _set(lib, "seo_map_dim", "argtypes", [C.c_void_p])
    _libs[kind] = lib
    return lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


COUNTER_NAMES = ("n_get", "n_interp", "n_grad", "n_active", "n_nodes", "n_new_blocks", "n_new_nodes",
                 "n_unique_keys", "n_keys_raw")


class Oracle:
    """One oracle pipeline (map + images), mirroring the stages of DenseSLAMSystem."""

    def __init__(self, field: int, size: int, dim: float, W: int, H: int, kind: str = "parity"):
        self.lib = load(kind)
        self.field, self.size, self.dim, self.W, self.H = field, size, float(dim), W, H
        self.h = C.c_void_p(self.lib.seo_create(field, size, dim, W, H))
        self.vdtype = FIELD_DTYPE[field]

    def close(self):
        if self.h:
            self.lib.seo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- N3 (reference build only): Octree::save, and Octree::load into a stand-alone octree -----------------------
    def save_map(self, path: str):
        self.lib.seo_save_map(self.h, path.encode())

    @classmethod
    def load_map(cls, field: int, path: str, kind: str, dim_fix: float = 0.0):
        """An Oracle whose handle wraps an octree read by the reference's own Octree::load (block / node / point queries
        only).  dim_fix > 0 repairs dim_, which the reference reads as an int (octree.hpp:921-923)."""
        o = cls.__new__(cls)
        o.lib = load(kind)
        o.field, o.vdtype = field, FIELD_DTYPE[field]
        o.h = C.c_void_p(o.lib.seo_load_map(path.encode(), dim_fix))
        o.size, o.dim, o.W, o.H = o.lib.seo_map_size(o.h), float(o.lib.seo_map_dim(o.h)), 0, 0
        return o

    def preprocess(self, depth_mm: np.ndarray):
        d = np.ascontiguousarray(depth_mm, dtype=np.uint16)
        return self.lib.seo_preprocess(self.h, _ptr(d), d.shape[1], d.shape[0])

    def set_depth(self, depth_m):
        d = np.ascontiguousarray(depth_m, dtype=np.float32)
        assert d.size == self.W * self.H
        self.lib.seo_set_depth(self.h, _ptr(d))

    def depth(self):
        d = np.empty((self.H, self.W), np.float32)
        self.lib.seo_get_depth(self.h, _ptr(d))
        return d

    def integrate(self, pose, k, mu, frame):
        p = np.ascontiguousarray(pose, np.float32)
        kk = np.ascontiguousarray(k, np.float32)
        return self.lib.seo_integrate(self.h, _ptr(p), _ptr(kk), mu, frame)

    def raycast(self, pose, k, mu):
        p = np.ascontiguousarray(pose, np.float32)
        kk = np.ascontiguousarray(k, np.float32)
        self.lib.seo_raycast(self.h, _ptr(p), _ptr(kk), mu)

    def vertex(self):
        v = np.empty((self.H, self.W, 3), np.float32)
        self.lib.seo_get_vertex(self.h, _ptr(v))
        return v

    def normal(self):
        v = np.empty((self.H, self.W, 3), np.float32)
        self.lib.seo_get_normal(self.h, _ptr(v))
        return v

    def render_volume(self, viewpose, k, mu, largestep, render: bool):
        out = np.empty((self.H, self.W, 4), np.uint8)
        p = np.ascontiguousarray(viewpose, np.float32)
        kk = np.ascontiguousarray(k, np.float32)
        self.lib.seo_render_volume(self.h, _ptr(out), _ptr(p), _ptr(kk), mu, largestep, int(render))
        return out

    def render_depth(self):
        out = np.empty((self.H, self.W, 4), np.uint8)
        self.lib.seo_render_depth(self.h, _ptr(out))
        return out

    def block_count(self):
        return self.lib.seo_block_count(self.h)

    def node_count(self):
        return self.lib.seo_node_count(self.h)

    def blocks_sorted(self, with_data=True):
        n = self.block_count()
        keys = np.empty(n, np.uint64)
        coords = np.empty((n, 3), np.int32)
        active = np.empty(n, np.uint8)
        data = np.empty((n, 512), self.vdtype) if with_data else None
        self.lib.seo_get_blocks_sorted(self.h, _ptr(keys), _ptr(coords), _ptr(active), _ptr(data) if with_data else None)
        return keys, coords, active, data

    def nodes_sorted(self):
        n = self.node_count()
        codes = np.empty(n, np.uint64)
        side = np.empty(n, np.uint32)
        mask = np.empty(n, np.uint8)
        values = np.empty((n, 8), self.vdtype)
        self.lib.seo_get_nodes_sorted(self.h, _ptr(codes), _ptr(side), _ptr(mask), _ptr(values))
        return codes, side, mask, values

    def allocate(self, keys):
        k = np.ascontiguousarray(keys, np.uint64)
        return self.lib.seo_allocate(self.h, _ptr(k), len(k))

    def hash(self, x, y, z, level=None):
        max_level = int(np.log2(self.size))
        if level is None:
            level = max_level - 3
        return self.lib.seo_key_encode(x, y, z, level, max_level)

    def fetch(self, x, y, z):
        return bool(self.lib.seo_fetch(self.h, x, y, z))

    def fetch_octant(self, x, y, z, depth):
        return bool(self.lib.seo_fetch_octant(self.h, x, y, z, depth))

    def fetch_octant_code(self, x, y, z, depth):
        return self.lib.seo_fetch_octant_code(self.h, x, y, z, depth)

    def get_fine(self, x, y, z):
        out = (C.c_double * 2)()
        self.lib.seo_get_fine(self.h, x, y, z, out)
        return out[0], out[1]

    def get(self, x, y, z):
        out = (C.c_double * 2)()
        self.lib.seo_get_coarse(self.h, x, y, z, out)
        return out[0], out[1]

    def set_voxel(self, x, y, z, vx, vy=0.0):
        self.lib.seo_set_voxel(self.h, x, y, z, vx, vy)

    def set_node_value(self, x, y, z, depth, slot, vx, vy=0.0):
        return self.lib.seo_set_node_value(self.h, x, y, z, depth, slot, vx, vy)

    def interp(self, x, y, z):
        return self.lib.seo_interp(self.h, x, y, z)

    def grad(self, x, y, z):
        out = (C.c_float * 3)()
        self.lib.seo_grad(self.h, x, y, z, out)
        return np.array(out[:], np.float32)

    def gather(self, x, y, z):
        out = (C.c_float * 8)()
        self.lib.seo_gather(self.h, x, y, z, out)
        return np.array(out[:], np.float32)

    def ray_blocks(self, origin, direction, near, far, max_out=4096):
        o = np.ascontiguousarray(origin, np.float32)
        d = np.ascontiguousarray(direction, np.float32)
        out = np.empty(max_out, np.uint64)
        tinfo = np.zeros(3, np.float32)
        n = self.lib.seo_ray_blocks(self.h, _ptr(o), _ptr(d), near, far, _ptr(out), max_out, _ptr(tinfo))
        return out[:min(n, max_out)].copy(), tinfo

    # ---- N1 ------------------------------------------------------------------------------
    def filter_depth(self, filter: bool, levels: int = 3):
        self.lib.seo_filter_depth(self.h, int(filter), levels)

    def track(self, pose, raycast_pose, k, icp_threshold, iterations):
        p = np.ascontiguousarray(pose, np.float32).reshape(4, 4).copy()
        rp = np.ascontiguousarray(raycast_pose, np.float32)
        kk = np.ascontiguousarray(k, np.float32)
        it = np.ascontiguousarray(iterations, np.int32)
        ok = self.lib.seo_tracking(self.h, _ptr(p), _ptr(rp), _ptr(kk), icp_threshold, _ptr(it), len(it))
        return p, bool(ok)

    def pyramid(self, level):
        w, h = self.W >> level, self.H >> level
        d = np.empty((h, w), np.float32); v = np.empty((h, w, 3), np.float32); n = np.empty((h, w, 3), np.float32)
        self.lib.seo_get_pyramid(self.h, level, _ptr(d), _ptr(v), _ptr(n))
        return d, v, n

    def marching_cube(self, table):
        """N4: triangles (n, 3, 3) float32, blocks in key order; `table` = (256, 16) int8 case table"""
        tab = np.ascontiguousarray(table, np.int8)
        n = self.lib.seo_marching_cube(self.h, _ptr(tab), None, 0)
        out = np.empty((n, 3, 3), np.float32)
        if n:
            self.lib.seo_marching_cube(self.h, _ptr(tab), _ptr(out), n)
        return out

    def tracking_data(self):
        td = np.empty((self.H, self.W), np.dtype([("result", "<i4"), ("error", "<f4"), ("J", "<f4", (6,))]))
        red = np.zeros(32, np.float32)
        self.lib.seo_get_tracking(self.h, _ptr(td), _ptr(red))
        return td, red

    def set_counting(self, on=True):
        self.lib.seo_set_counting(self.h, int(on))

    def reset_counters(self):
        self.lib.seo_reset_counters(self.h)

    def counters(self):
        out = np.zeros(9, np.uint64)
        self.lib.seo_get_counters(self.h, _ptr(out))
        return dict(zip(COUNTER_NAMES, (int(v) for v in out)))
