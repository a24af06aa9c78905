"""ctypes binding of include/se_b200.h (libse_b200.so).  Mirrors the C ABI one to one."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

SE_B200_SDF, SE_B200_OFUSION = 0, 1
SDF_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4")])
OFUSION_DTYPE = np.dtype([("x", "<f4"), ("_pad", "<f4"), ("y", "<f8")])
FIELD_DTYPE = {SE_B200_SDF: SDF_DTYPE, SE_B200_OFUSION: OFUSION_DTYPE}
STAGES = ("preprocess", "alloc", "fuse", "raycast", "render")

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None

# every symbol include/se_b200.h declares (tests check that the library exports them all)
EXPORTS = (
    "se_b200_last_error", "se_b200_device_count", "se_b200_bspline_lut", "se_b200_create", "se_b200_destroy", "se_b200_set_stream",
    "se_b200_sync", "se_b200_preprocess_depth_host", "se_b200_preprocess_depth_device", "se_b200_set_depth_m_host",
    "se_b200_integrate", "se_b200_raycast", "se_b200_raycast_count_samples", "se_b200_download_vertex_normal", "se_b200_upload_vertex_normal",
    "se_b200_render_volume_host", "se_b200_render_volume_device", "se_b200_render_depth_host",
    "se_b200_render_track_host", "se_b200_filter_depth", "se_b200_track", "se_b200_download_pyramid",
    "se_b200_download_tracking", "se_b200_block_count", "se_b200_node_count", "se_b200_download_blocks_sorted",
    "se_b200_download_nodes_sorted", "se_b200_upload_blocks", "se_b200_upload_nodes", "se_b200_allocate_keys", "se_b200_query_voxels", "se_b200_query_interp",
    "se_b200_query_grad", "se_b200_set_voxels", "se_b200_query_rays", "se_b200_elapsed_ms", "se_b200_set_stage_timing", "se_b200_counters",
    "se_b200_launch_count", "se_b200_device_image", "se_b200_extract_mesh", "se_b200_download_mesh", "se_b200_mc_table",
    "se_b200_preprocess_depth_host_async", "se_b200_render_volume_host_async", "se_b200_set_render_target",
    "se_b200_register_host_buffer", "se_b200_unregister_host_buffer",
)


class SeB200Error(RuntimeError):
    pass


def lib_path() -> str:
    # SE_B200_LIB: load another build of the same library (A/B measurements of kernel variants)
    return os.environ.get("SE_B200_LIB") or os.path.join(_HERE, "libse_b200.so")


def load_library():
    """Load libse_b200.so (built in-tree by __graft_entry__.build()).  No fallback of any kind."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise SeB200Error(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)")
    lib = C.CDLL(path)
    vp, i32, f32, u32, i64 = C.c_void_p, C.c_int, C.c_float, C.c_uint, C.c_int64
    lib.se_b200_last_error.restype = C.c_char_p
    lib.se_b200_create.argtypes = [C.POINTER(vp), i32, i32, f32, i32, i32, i64, i64, i32]
    lib.se_b200_destroy.argtypes = [vp]
    lib.se_b200_set_stream.argtypes = [vp, vp]
    lib.se_b200_sync.argtypes = [vp]
    lib.se_b200_preprocess_depth_host.argtypes = [vp, vp, i32, i32]
    lib.se_b200_preprocess_depth_host_async.argtypes = [vp, vp, i32, i32]
    lib.se_b200_preprocess_depth_device.argtypes = [vp, vp, i32, i32]
    lib.se_b200_set_depth_m_host.argtypes = [vp, vp]
    lib.se_b200_integrate.argtypes = [vp, vp, vp, f32, u32]
    lib.se_b200_raycast.argtypes = [vp, vp, vp, f32]
    lib.se_b200_raycast_count_samples.argtypes = [vp, vp, vp, f32, vp]
    lib.se_b200_download_vertex_normal.argtypes = [vp, vp, vp]
    lib.se_b200_upload_vertex_normal.argtypes = [vp, vp, vp]
    lib.se_b200_render_volume_host.argtypes = [vp, vp, vp, vp, f32, f32, i32]
    lib.se_b200_render_volume_host_async.argtypes = [vp, vp, vp, vp, f32, f32, i32]
    lib.se_b200_render_volume_device.argtypes = [vp, vp, vp, vp, f32, f32, i32]
    lib.se_b200_render_depth_host.argtypes = [vp, vp]
    lib.se_b200_set_render_target.argtypes = [vp, vp]
    if hasattr(lib, "se_b200_register_host_buffer"):       # (absent from older builds loaded through SE_B200_LIB for A/B runs)
        lib.se_b200_register_host_buffer.argtypes = [vp, C.c_size_t]
        lib.se_b200_unregister_host_buffer.argtypes = [vp]
    lib.se_b200_render_track_host.argtypes = [vp, vp, vp, i32]
    lib.se_b200_filter_depth.argtypes = [vp, i32, i32]
    lib.se_b200_track.argtypes = [vp, vp, vp, vp, f32, vp, i32, C.POINTER(i32)]
    lib.se_b200_download_pyramid.argtypes = [vp, i32, vp, vp, vp]
    lib.se_b200_download_tracking.argtypes = [vp, vp, vp]
    lib.se_b200_extract_mesh.argtypes = [vp, C.POINTER(C.c_int64)]
    lib.se_b200_download_mesh.argtypes = [vp, vp, C.c_int64]
    lib.se_b200_mc_table.argtypes = [vp]
    lib.se_b200_mc_table.restype = None
    lib.se_b200_block_count.argtypes = [vp, C.POINTER(i32)]
    lib.se_b200_node_count.argtypes = [vp, C.POINTER(i32)]
    lib.se_b200_download_blocks_sorted.argtypes = [vp, vp, vp, vp, vp]
    lib.se_b200_download_nodes_sorted.argtypes = [vp, vp, vp, vp, vp]
    lib.se_b200_allocate_keys.argtypes = [vp, vp, i32]
    lib.se_b200_upload_blocks.argtypes = [vp, vp, vp, i32]
    lib.se_b200_upload_nodes.argtypes = [vp, vp, vp, i32]
    lib.se_b200_query_voxels.argtypes = [vp, vp, i32, vp]
    lib.se_b200_query_interp.argtypes = [vp, vp, i32, vp]
    lib.se_b200_query_grad.argtypes = [vp, vp, i32, vp]
    lib.se_b200_set_voxels.argtypes = [vp, vp, vp, i32]
    lib.se_b200_query_rays.argtypes = [vp, vp, i32, f32, f32, vp, vp]
    lib.se_b200_elapsed_ms.argtypes = [vp, i32, C.POINTER(f32)]
    lib.se_b200_set_stage_timing.argtypes = [vp, i32]
    lib.se_b200_counters.argtypes = [vp, vp]
    lib.se_b200_launch_count.argtypes = [vp, C.POINTER(i64)]
    lib.se_b200_device_image.argtypes = [vp, i32, C.POINTER(vp)]
    _lib = lib
    return lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a, n=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if n is not None:
        assert a.size == n, (a.size, n)
    return a


class Map:
    """One device-resident map + its images: the state behind a DenseSLAMSystem."""

    def __init__(self, field: int, size: int, dim: float, W: int, H: int, max_blocks: int = 0, max_nodes: int = 0,
                 device: int = 0):
        self.lib = load_library()
        self.field, self.size, self.dim, self.W, self.H, self.device = field, size, float(dim), W, H, device
        self.vdtype = FIELD_DTYPE[field]
        h = C.c_void_p()
        self._check(self.lib.se_b200_create(C.byref(h), field, size, dim, W, H, max_blocks, max_nodes, device))
        self.h = h

    def _check(self, rc):
        if rc != 0:
            raise SeB200Error(f"se_b200 error {rc}: {self.lib.se_b200_last_error().decode()}")

    def close(self):
        if getattr(self, "h", None):
            self.lib.se_b200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- stages ---------------------------------------------------------------------------
    def set_stream(self, cuda_stream_ptr):
        self._check(self.lib.se_b200_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    def sync(self):
        self._check(self.lib.se_b200_sync(self.h))

    def preprocess(self, depth_mm: np.ndarray):
        d = np.ascontiguousarray(depth_mm, dtype=np.uint16)
        self._check(self.lib.se_b200_preprocess_depth_host(self.h, _ptr(d), d.shape[1], d.shape[0]))
        self._keep = d   # the copy may be asynchronous

    def preprocess_host_ptr(self, ptr: int, inW: int, inH: int):
        self._check(self.lib.se_b200_preprocess_depth_host(self.h, C.c_void_p(ptr), inW, inH))

    def preprocess_device_ptr(self, ptr: int, inW: int, inH: int):
        self._check(self.lib.se_b200_preprocess_depth_device(self.h, C.c_void_p(ptr), inW, inH))

    def set_depth(self, depth_m):
        d = _f32(depth_m, self.W * self.H)
        self._check(self.lib.se_b200_set_depth_m_host(self.h, _ptr(d)))

    def integrate(self, pose, k, mu, frame):
        p, kk = _f32(pose, 16), _f32(k, 4)
        self._check(self.lib.se_b200_integrate(self.h, _ptr(p), _ptr(kk), mu, frame))

    def raycast(self, pose, k, mu):
        p, kk = _f32(pose, 16), _f32(k, 4)
        self._check(self.lib.se_b200_raycast(self.h, _ptr(p), _ptr(kk), mu))

    def raycast_count_samples(self, pose, k, mu):
        p, kk = _f32(pose, 16), _f32(k, 4)
        out = np.zeros(4, np.uint64)
        self._check(self.lib.se_b200_raycast_count_samples(self.h, _ptr(p), _ptr(kk), mu, _ptr(out)))
        return dict(n_get=int(out[0]), n_interp=int(out[1]), n_grad=int(out[2]), n_walk=int(out[3]))

    def vertex_normal(self):
        v = np.empty((self.H, self.W, 3), np.float32)
        n = np.empty((self.H, self.W, 3), np.float32)
        self._check(self.lib.se_b200_download_vertex_normal(self.h, _ptr(v), _ptr(n)))
        return v, n

    def upload_vertex_normal(self, v, n):
        v, n = _f32(v, self.W * self.H * 3), _f32(n, self.W * self.H * 3)
        self._check(self.lib.se_b200_upload_vertex_normal(self.h, _ptr(v), _ptr(n)))

    def render_volume(self, view_pose, k, mu, largestep, reraycast: bool, out=None):
        if out is None:
            out = np.empty((self.H, self.W, 4), np.uint8)
        p, kk = _f32(view_pose, 16), _f32(k, 4)
        self._check(self.lib.se_b200_render_volume_host(self.h, _ptr(out), _ptr(p), _ptr(kk), mu, largestep, int(reraycast)))
        return out

    def render_volume_host_ptr(self, out_ptr: int, view_pose, k, mu, largestep, reraycast: bool):
        p, kk = _f32(view_pose, 16), _f32(k, 4)
        self._check(self.lib.se_b200_render_volume_host(self.h, C.c_void_p(out_ptr), _ptr(p), _ptr(kk), mu, largestep, int(reraycast)))

    def render_volume_device_ptr(self, out_ptr: int, view_pose, k, mu, largestep, reraycast: bool):
        p, kk = _f32(view_pose, 16), _f32(k, 4)
        self._check(self.lib.se_b200_render_volume_device(self.h, C.c_void_p(out_ptr), _ptr(p), _ptr(kk), mu, largestep, int(reraycast)))

    def set_render_target(self, out_ptr):
        """se_b200_set_render_target: `out_ptr` = address of W*H*4 bytes of device or pinned host memory, or None / 0 to turn it off"""
        self._check(self.lib.se_b200_set_render_target(self.h, C.c_void_p(out_ptr or None)))

    def render_depth(self):
        out = np.empty((self.H, self.W, 4), np.uint8)
        self._check(self.lib.se_b200_render_depth_host(self.h, _ptr(out)))
        return out

    def render_track(self, result, stride_ints=1):
        r = np.ascontiguousarray(result, dtype=np.int32)
        out = np.empty((self.H, self.W, 4), np.uint8)
        self._check(self.lib.se_b200_render_track_host(self.h, _ptr(out), _ptr(r), stride_ints))
        return out

    # ---- N1: tracking front-end ---------------------------------------------------------
    def filter_depth(self, filter: bool, levels: int = 3):
        self._check(self.lib.se_b200_filter_depth(self.h, int(filter), levels))

    def track(self, pose, raycast_pose, k, icp_threshold, iterations):
        """Returns (new_pose [4,4] float32, tracked bool)."""
        p = _f32(pose, 16).reshape(4, 4).copy()
        rp, kk = _f32(raycast_pose, 16), _f32(k, 4)
        it = np.ascontiguousarray(iterations, dtype=np.int32)
        ok = C.c_int()
        self._check(self.lib.se_b200_track(self.h, _ptr(p), _ptr(rp), _ptr(kk), icp_threshold, _ptr(it), len(it), C.byref(ok)))
        return p, bool(ok.value)

    def pyramid(self, level):
        w, h = self.W >> level, self.H >> level
        d = np.empty((h, w), np.float32); v = np.empty((h, w, 3), np.float32); n = np.empty((h, w, 3), np.float32)
        self._check(self.lib.se_b200_download_pyramid(self.h, level, _ptr(d), _ptr(v), _ptr(n)))
        return d, v, n

    def tracking_data(self):
        td = np.empty((self.H, self.W), np.dtype([("result", "<i4"), ("error", "<f4"), ("J", "<f4", (6,))]))
        red = np.zeros(32, np.float32)
        self._check(self.lib.se_b200_download_tracking(self.h, _ptr(td), _ptr(red)))
        return td, red

    def render_track_last(self):
        out = np.empty((self.H, self.W, 4), np.uint8)
        self._check(self.lib.se_b200_render_track_host(self.h, _ptr(out), None, 0))
        return out

    # ---- N4: meshing ----------------------------------------------------------------------
    def mesh(self):
        """DenseSLAMSystem::dump_mesh's triangle list: (n, 3, 3) float32, vertexes[0..2] of every triangle, metres"""
        n = C.c_int64()
        self._check(self.lib.se_b200_extract_mesh(self.h, C.byref(n)))
        out = np.empty((n.value, 3, 3), np.float32)
        self._check(self.lib.se_b200_download_mesh(self.h, _ptr(out) if n.value else None, n.value))
        return out

    # ---- inspection -----------------------------------------------------------------------
    def block_count(self):
        n = C.c_int()
        self._check(self.lib.se_b200_block_count(self.h, C.byref(n)))
        return n.value

    def block_count_nothrow(self):
        """the count even when the call reports SE_B200_ERR_POOL (the value is filled in before the status is returned)"""
        n = C.c_int()
        self.lib.se_b200_block_count(self.h, C.byref(n))
        return n.value

    def node_count(self):
        n = C.c_int()
        self._check(self.lib.se_b200_node_count(self.h, C.byref(n)))
        return n.value

    def blocks_sorted(self, with_data=True):
        n = self.block_count()
        keys = np.empty(n, np.uint64)
        coords = np.empty((n, 3), np.int32)
        active = np.empty(n, np.uint8)
        data = np.empty((n, 512), self.vdtype) if with_data else None
        self._check(self.lib.se_b200_download_blocks_sorted(self.h, _ptr(keys), _ptr(coords), _ptr(active), _ptr(data)))
        return keys, coords, active, data

    def nodes_sorted(self):
        n = self.node_count()
        codes = np.empty(n, np.uint64)
        side = np.empty(n, np.uint32)
        mask = np.empty(n, np.uint8)
        values = np.empty((n, 8), self.vdtype)
        self._check(self.lib.se_b200_download_nodes_sorted(self.h, _ptr(codes), _ptr(side), _ptr(mask), _ptr(values)))
        return codes, side, mask, values

    def upload_blocks(self, keys, voxels):
        k = np.ascontiguousarray(keys, dtype=np.uint64)
        v = np.ascontiguousarray(voxels, dtype=self.vdtype).reshape(len(k), 512)
        self._check(self.lib.se_b200_upload_blocks(self.h, _ptr(k), _ptr(v), len(k)))

    def upload_nodes(self, codes, values):
        k = np.ascontiguousarray(codes, dtype=np.uint64)
        v = np.ascontiguousarray(values, dtype=self.vdtype).reshape(len(k), 8)
        self._check(self.lib.se_b200_upload_nodes(self.h, _ptr(k), _ptr(v), len(k)))

    def allocate(self, keys):
        k = np.ascontiguousarray(keys, dtype=np.uint64)
        self._check(self.lib.se_b200_allocate_keys(self.h, _ptr(k), len(k)))

    def query_voxels(self, xyz):
        p = np.ascontiguousarray(xyz, dtype=np.int32).reshape(-1, 3)
        out = np.empty(len(p), self.vdtype)
        self._check(self.lib.se_b200_query_voxels(self.h, _ptr(p), len(p), _ptr(out)))
        return out

    def query_interp(self, pos):
        p = np.ascontiguousarray(pos, dtype=np.float32).reshape(-1, 3)
        out = np.empty(len(p), np.float32)
        self._check(self.lib.se_b200_query_interp(self.h, _ptr(p), len(p), _ptr(out)))
        return out

    def query_grad(self, pos):
        p = np.ascontiguousarray(pos, dtype=np.float32).reshape(-1, 3)
        out = np.empty((len(p), 3), np.float32)
        self._check(self.lib.se_b200_query_grad(self.h, _ptr(p), len(p), _ptr(out)))
        return out

    def set_voxels(self, xyz, values):
        p = np.ascontiguousarray(xyz, dtype=np.int32).reshape(-1, 3)
        v = np.ascontiguousarray(values, dtype=self.vdtype)
        assert len(v) == len(p)
        self._check(self.lib.se_b200_set_voxels(self.h, _ptr(p), _ptr(v), len(p)))

    def query_rays(self, origin_dir, near, far):
        p = np.ascontiguousarray(origin_dir, dtype=np.float32).reshape(-1, 6)
        keys = np.empty(len(p), np.uint64)
        tinfo = np.empty((len(p), 3), np.float32)
        self._check(self.lib.se_b200_query_rays(self.h, _ptr(p), len(p), near, far, _ptr(keys), _ptr(tinfo)))
        return keys, tinfo

    # ---- measurement ----------------------------------------------------------------------
    def elapsed_ms(self, stage):
        ms = C.c_float()
        idx = STAGES.index(stage) if isinstance(stage, str) else int(stage)
        self._check(self.lib.se_b200_elapsed_ms(self.h, idx, C.byref(ms)))
        return ms.value

    def set_stage_timing(self, enable: bool):
        self._check(self.lib.se_b200_set_stage_timing(self.h, int(enable)))

    def counters(self):
        out = np.zeros(8, np.int32)
        self._check(self.lib.se_b200_counters(self.h, _ptr(out)))
        return dict(nodes=int(out[0]), blocks=int(out[1]), active=int(out[2]), error=int(out[3]),
                    blocks_before=int(out[4]), nodes_before=int(out[5]), requests=int(out[6]))

    def launch_count(self):
        n = C.c_int64()
        self._check(self.lib.se_b200_launch_count(self.h, C.byref(n)))
        return n.value

    def device_image(self, which: int) -> int:
        p = C.c_void_p()
        self._check(self.lib.se_b200_device_image(self.h, which, C.byref(p)))
        return p.value


def mc_table():
    """the marching-cubes case table the meshing kernel uses: (256, 16) int8, -1 terminated rows (host only, no GPU)"""
    t = np.empty((256, 16), np.int8)
    load_library().se_b200_mc_table(_ptr(t))
    return t
