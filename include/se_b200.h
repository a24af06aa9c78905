/* se_b200.h -- C ABI of the B200-native supereight hot path.
 *
 * This is the drop-in boundary: plain C, opaque handle, plain pointers
 * and sizes, row-major 4x4 float poses, no Eigen/torch types.  The C++ class
 * `DenseSLAMSystem` in supereight_b200/host/ (same public methods as the reference's
 * se_denseslam/include/se/DenseSLAMSystem.h:58-411) is a thin shim over these calls; a
 * maintainer of the reference binds the same calls from DenseSLAMSystem.cpp.
 *
 * Every entry point names the reference code it replaces (paths relative to the reference
 * tree).  All functions return 0 on success and a negative code on failure;
 * se_b200_last_error() gives the message for the calling thread.  There is no CPU fallback:
 * without a CUDA device every call fails with SE_B200_ERR_CUDA.
 *
 * Threading: like the reference (one caller thread drives the stages in order), a map must
 * not be used from two threads at once.  All work of a map is enqueued on one CUDA stream
 * (se_b200_set_stream); calls that return data to host memory synchronise that stream, the
 * others return as soon as the work is enqueued.
 */
#ifndef SE_B200_H
#define SE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct se_b200_map se_b200_map;

/* SE_FIELD_TYPE of the reference (DenseSLAMSystem.h:54, volume_traits.hpp:41-72) */
enum { SE_B200_SDF = 0, SE_B200_OFUSION = 1 };

enum {
  SE_B200_OK = 0,
  SE_B200_ERR_ARG = -1,       /* bad argument (the reference prints "Invalid ratio." and exit(1)s, preprocessing.cpp:165-176) */
  SE_B200_ERR_CUDA = -2,      /* CUDA runtime error, or no device */
  SE_B200_ERR_POOL = -3       /* node/block pool exhausted (reference: allocation list silently truncated, kfusion/alloc_impl.hpp:103-106) */
};
/* SE_B200_ERR_POOL and the per-frame path.  A pool runs out on the device, in the middle of se_b200_integrate's allocation
 * pass, and the per-frame calls never wait for the device: the octants that did not fit are dropped for that frame (they are
 * requested again by the next one), the frame's integrate kernel hands the error bits to the host through a word of
 * mapped page-locked memory, and the NEXT se_b200_integrate / se_b200_raycast (normally one frame later; at the latest
 * after a synchronising call) returns SE_B200_ERR_POOL once and clears the condition.  se_b200_block_count,
 * se_b200_node_count, se_b200_counters and the map import calls synchronise and report it at once.  Re-create the map with
 * larger max_blocks / max_nodes (se_b200_download_* + se_b200_upload_* carry the content over). */

/* voxel payload layouts, identical to the reference structs (volume_traits.hpp:41-44, 62-65) */
typedef struct { float x; float y; } se_b200_sdf_voxel;                 /* tsdf in [-1,1], weight */
typedef struct { float x; float pad_; double y; } se_b200_ofusion_voxel; /* log-odds occupancy, timestamp [s] */

const char* se_b200_last_error(void);
int se_b200_device_count(void);
/* the 1000-entry B-spline table the OFusion update samples (se_denseslam/src/bfusion/bspline_lookup.cc:36-37),
 * as uploaded to constant memory; host-side, needs no device */
int se_b200_bspline_lut(float out[1000]);

/* ---- lifetime --------------------------------------------------------------------------
 * Replaces se::Octree<FieldType>::init (se_core/include/se/octree.hpp:411-421) and the image
 * members of DenseSLAMSystem (DenseSLAMSystem.cpp:65-126).  The octree lives in device memory
 * as flat index-addressed pools of `max_nodes` nodes and `max_blocks` 8x8x8 VoxelBlocks
 * (0 = default: min((size/8)^3, 2^20) blocks, a quarter as many nodes).  `size` must be a power of two >= 16, `W`x`H` is
 * the computation size. */
int se_b200_create(se_b200_map** out, int field_type, int size, float dim, int W, int H,
                   int64_t max_blocks, int64_t max_nodes, int device);
int se_b200_destroy(se_b200_map* map);
/* Use the caller's CUDA stream (a cudaStream_t passed as void*; NULL = the map's own stream). */
int se_b200_set_stream(se_b200_map* map, void* cuda_stream);
int se_b200_sync(se_b200_map* map);

/* ---- a1: mm2metersKernel (se_denseslam/src/preprocessing.cpp:161-188) -------------------
 * depth_mm is inW x inH uint16 millimetres; the result is the W x H float depth (metres) that
 * integration and renderDepth read.  inW/W == inH/H must be a whole ratio.
 * _host: pointer to host memory (copied to a staging buffer in HBM inside the call, asynchronously if it is page-locked);
 * _device: pointer to device memory already holding the frame.
 * The conversion itself is deferred: the allocation kernel of the next se_b200_integrate reads every pixel exactly once
 * anyway and converts it there (any other consumer of the float image -- se_b200_filter_depth, se_b200_render_depth_host,
 * se_b200_device_image(0) -- triggers a conversion kernel first); results are those of the eager kernel, bit for bit.
 * Buffer lifetime.  _host with PAGE-LOCKED memory: the copy is asynchronous -- do not modify depth_mm until a
 * synchronising call (se_b200_sync, any call that returns data to host memory) has returned; pageable memory is staged
 * before the call returns.  _device: depth_mm_dev is READ LATER by that next consumer -- keep it valid
 * and unchanged until the work of the se_b200_integrate (or other consumer) that follows has been enqueued AND anything
 * you then do to the buffer is ordered after the map's stream; and it must HOLD the frame when the call is made (written and
 * device-visible -- e.g. produced by work the caller has synchronised, or that precedes the previous frame's calls in the
 * map's stream): for SDF maps the allocation kernel that reads it runs on a second, internal stream, ordered behind the
 * previous frame's integrate kernel but not behind its raycast (frames overlap; results are those of the serial order).
 * se_b200_register_host_buffer page-locks (cudaHostRegister) a caller-owned buffer -- the reference application malloc()s its
 * depth and RGBA buffers (se_apps/src/benchmark.cpp:90-97), which makes every transfer a staged, synchronous copy; registering
 * them once makes the _host calls above and below asynchronous / zero-copy.  Unregister before freeing the memory. */
int se_b200_register_host_buffer(void* ptr, size_t bytes);
int se_b200_unregister_host_buffer(void* ptr);
int se_b200_preprocess_depth_host(se_b200_map* map, const uint16_t* depth_mm, int inW, int inH);
int se_b200_preprocess_depth_device(se_b200_map* map, const uint16_t* depth_mm_dev, int inW, int inH);
/* test/IO helper: set the float depth image directly (W*H floats, host memory) */
int se_b200_set_depth_m_host(se_b200_map* map, const float* depth_m);

/* ---- a3..a12: the body of DenseSLAMSystem::integration (DenseSLAMSystem.cpp:211-253) ----
 * buildAllocationList / buildOctantList (kfusion|bfusion/alloc_impl.hpp) + Octree::allocate
 * (octree.hpp:792-856) + se::functor::projective_map (functors/projective_functor.hpp:139-156)
 * with sdf_update / bfusion_update.  pose = camera-to-world, k = (fx, fy, cx, cy), `frame`
 * only feeds the OFusion timestamp frame/30.  The frame gate (frame % rate) stays in the caller. */
int se_b200_integrate(se_b200_map* map, const float pose[16], const float k[4], float mu, unsigned frame);

/* ---- a13..a17: raycastKernel (se_denseslam/src/rendering.cpp:50-90) ---------------------
 * view = pose * K^-1, near/far planes 0.4/4.0 m (constant_parameters.h:27,32), step = one
 * voxel, largestep = one block (DenseSLAMSystem.cpp:197-200).  Vertex and normal maps stay on
 * the device (W*H*3 floats each, row-major, as se::Image<Eigen::Vector3f>). */
int se_b200_raycast(se_b200_map* map, const float pose[16], const float k[4], float mu);
/* same raycast, additionally counting the field samples it takes: samples[0] = VolumeTemplate::get,
 * [1] = interp (8 voxels each), [2] = grad (32 distinct voxels each) -- the inputs of the
 * algorithmic-bytes figure of SURVEY.md 8(d) -- and [3] = octree walk steps.  Measurement only. */
int se_b200_raycast_count_samples(se_b200_map* map, const float pose[16], const float k[4], float mu, uint64_t samples[4]);
int se_b200_download_vertex_normal(se_b200_map* map, float* vertex, float* normal);   /* either may be NULL */
int se_b200_upload_vertex_normal(se_b200_map* map, const float* vertex, const float* normal);

/* ---- a18: renderVolumeKernel / renderDepthKernel / renderTrackKernel --------------------
 * (rendering.cpp:214-283, 111-152, 154-212).  out is W*H*4 bytes RGBA, caller-owned, as in
 * the reference (benchmark.cpp:90-97).  reraycast != 0 is the reference's `render` flag (view
 * pose differs from the raycast pose: cast again from the volume entry with far plane 8 m);
 * 0 shades the stored vertex/normal maps.  light = translation of view_pose.
 * track_result: W*H ints `stride_ints` apart (TrackData::result, commons.h:249-253 -> stride 8);
 * NULL = the TrackData the last se_b200_track left on the device. */
int se_b200_render_volume_host(se_b200_map* map, uint8_t* out, const float view_pose[16], const float k[4],
                               float mu, float largestep, int reraycast);
int se_b200_render_volume_device(se_b200_map* map, uint8_t* out_dev, const float view_pose[16], const float k[4],
                                 float mu, float largestep, int reraycast);
int se_b200_render_depth_host(se_b200_map* map, uint8_t* out);

/* Extension (no counterpart in the reference, whose renderVolume always runs after raycasting has returned): while a render
 * target is set, se_b200_raycast also shades every pixel exactly as renderVolume's reuse path would (rendering.cpp:259-279,
 * light at the raycast pose) and stores the RGBA image to `out` as each ray finishes.  `out` is W*H*4 bytes of device memory
 * or of PAGE-LOCKED host memory (cudaHostAlloc / cudaHostRegister: its mapped alias is written in place, so the image crosses
 * PCIe while the remaining rays are still being cast); NULL turns the mode off.  se_b200_render_volume_host / _device called
 * with that same pointer, reraycast == 0 and a view pose whose translation is the raycast pose's then launch nothing:
 * _host only synchronises, _device returns (the image is there in stream order).  Any other call renders as usual.
 * The caller must keep `out` valid, and leave its contents alone between the raycast and the render call, until the target is
 * changed or the map destroyed. */
int se_b200_set_render_target(se_b200_map* map, uint8_t* out);

/* ---- overlapped host I/O (no counterpart in the reference, whose stages are synchronous) ----------------
 * The same two stages as se_b200_preprocess_depth_host / se_b200_render_volume_host, but the copies run on the map's own
 * upload / download streams through double-buffered staging in HBM, ordered against the kernel stream by events:
 * frame N+1's depth upload and frame N's image download overlap the kernels of the frames in between.  Both calls return
 * without waiting.  depth_mm must stay valid, and out must not be read, until se_b200_sync (which also drains these
 * streams) -- or, for out, until two further se_b200_render_volume_host_async calls have been issued and a later
 * synchronising call returned.  Use page-locked host memory (pageable memory makes the copies synchronous). */
int se_b200_preprocess_depth_host_async(se_b200_map* map, const uint16_t* depth_mm, int inW, int inH);
int se_b200_render_volume_host_async(se_b200_map* map, uint8_t* out, const float view_pose[16], const float k[4],
                                     float mu, float largestep, int reraycast);
int se_b200_render_track_host(se_b200_map* map, uint8_t* out, const int* track_result, int stride_ints);

/* ---- N1: the tracking front-end either side of the hot path ----------------
 * se_b200_filter_depth = second half of DenseSLAMSystem::preprocessing (DenseSLAMSystem.cpp:132-139):
 *   bilateralFilterKernel (preprocessing.cpp:42-87) when filter != 0, else a copy, into level 0 of the
 *   depth pyramid; call it after se_b200_preprocess_depth_*.  `levels` = pyramid depth (config.pyramid.size()).
 * se_b200_track = the body of DenseSLAMSystem::tracking (DenseSLAMSystem.cpp:149-188): half-sample pyramid,
 *   depth2vertex / vertex2normal per level (preprocessing.cpp:89-226), then per level `iterations[level]` rounds
 *   of trackKernel + reduceKernel (tracking.cpp:66-300) and the 6x6 solve + SE3 exponential of updatePoseKernel
 *   (tracking.cpp:302-318), all on the GPU with the pose kept in device memory (no host round trip per iteration;
 *   iterations after a level's convergence return at once), against the vertex / normal maps of the last
 *   se_b200_raycast taken from `raycast_pose`; finally checkPoseKernel (:320-336) on the host from the 32 sums.  pose_io is updated in place (restored when the check
 *   fails), *tracked = the check's verdict.  se_b200_render_track_host(map, out, NULL, 0) renders its result. */
int se_b200_filter_depth(se_b200_map* map, int filter, int levels);
int se_b200_track(se_b200_map* map, float pose_io[16], const float raycast_pose[16], const float k[4], float icp_threshold,
                  const int* iterations, int levels, int* tracked);
/* inspection: one pyramid level (depth (W>>l)*(H>>l), vertex / normal 3 floats per pixel) and the TrackData image
 * (W*H records of 8 x 4 bytes: int result, float error, float J[6]) + the 32 reduced sums of the last iteration */
int se_b200_download_pyramid(se_b200_map* map, int level, float* depth, float* vertex, float* normal);
int se_b200_download_tracking(se_b200_map* map, void* track_data, float reduction[32]);

/* ---- N4: DenseSLAMSystem::dump_mesh (DenseSLAMSystem.cpp:302-322) ---------------------
 * se_b200_extract_mesh = se::algorithms::marching_cube (meshing.hpp:158-208) with dump_mesh's functors (inside = x < 0,
 *   select = x) over every allocated VoxelBlock, into a library-owned device buffer; *n_triangles = how many it kept.
 *   Triangles come out in the order of a serial run over the block list sorted by key (the reference's order depends on
 *   its OpenMP schedule, meshing.hpp:173-203), cells x-outer / z-inner, table order inside a cell.
 * se_b200_download_mesh copies them out: 9 floats per triangle = vertexes[0..2] (x, y, z) of the reference's Triangle
 *   (commons.h:166-168), metres in the volume frame.
 * se_b200_mc_table (host only, no GPU needed): the 256 x 16 case table the kernel uses, -1 terminated rows in the
 *   reference's edge numbering (meshing.hpp:58-104): the classic marching-cubes tabulation, the list the reference's
 *   edge_tables.h transcribes too, so the triangles are the reference's (see csrc/se_meshing.cuh). */
int se_b200_extract_mesh(se_b200_map* map, int64_t* n_triangles);
int se_b200_download_mesh(se_b200_map* map, float* triangles, int64_t capacity_triangles);
void se_b200_mc_table(int8_t table[4096]);

/* ---- inspection: what getMap() / Octree::save expose in the reference --------------------
 * (DenseSLAMSystem.h:295-297, octree.hpp:897-915).  Pool order is arbitrary in the reference
 * (OpenMP scheduling) and here (atomics); "sorted" = ascending key, the comparable form.
 * Any output pointer may be NULL.  voxels: n*512 payload structs; values: n*8. */
int se_b200_block_count(se_b200_map* map, int* out);
int se_b200_node_count(se_b200_map* map, int* out);
int se_b200_download_blocks_sorted(se_b200_map* map, uint64_t* keys, int32_t* coords_xyz, uint8_t* active, void* voxels);
int se_b200_download_nodes_sorted(se_b200_map* map, uint64_t* codes, uint32_t* side, uint8_t* children_mask, void* values);
/* Octree::load (octree.hpp:917-950): re-create nodes / blocks from saved records -- insert(coords, level)
 * followed by a copy of value_[8] / the 512-voxel payload.  keys/codes carry the level in their low 9 bits as
 * written by Octree::save; `voxels` is n*512 payload structs, `values` n*8.  Either call may be used alone. */
int se_b200_upload_blocks(se_b200_map* map, const uint64_t* keys, const void* voxels, int n);
int se_b200_upload_nodes(se_b200_map* map, const uint64_t* codes, const void* values, int n);
/* Octree::allocate for an explicit key list (octree.hpp:792-817), incl. multi-level keys */
int se_b200_allocate_keys(se_b200_map* map, const uint64_t* keys, int n);
/* point queries, n points each: get_fine (octree.hpp:356-377) at integer voxels, interp
 * (:541-563) and grad (:652-737) at float voxel positions, Octree::set (:310-329) */
int se_b200_query_voxels(se_b200_map* map, const int32_t* xyz, int n, void* voxels_out);
int se_b200_query_interp(se_b200_map* map, const float* pos_xyz, int n, float* out);
int se_b200_query_grad(se_b200_map* map, const float* pos_xyz, int n, float* out_xyz);
int se_b200_set_voxels(se_b200_map* map, const int32_t* xyz, const void* voxels, int n);
/* se::ray_iterator (ray_iterator.hpp:49-289): first block on each ray (key, or ~0) and
 * (tmin, tmax, tcmin) in metres.  origin_dir: n * (ox,oy,oz,dx,dy,dz). */
int se_b200_query_rays(se_b200_map* map, const float* origin_dir, int n, float near_plane, float far_plane,
                       uint64_t* first_block_key, float* tmin_tmax_tcmin);

/* ---- measurement ------------------------------------------------------------------------
 * Device time of the most recent run of a stage (CUDA events on the map's stream); stands in
 * for the TICK/TOCK samples of se_shared/timings.h:7-15. */
enum { SE_B200_STAGE_PREPROCESS = 0, SE_B200_STAGE_ALLOC = 1, SE_B200_STAGE_FUSE = 2, SE_B200_STAGE_RAYCAST = 3,
       SE_B200_STAGE_RENDER = 4, SE_B200_NUM_STAGES = 5 };
int se_b200_elapsed_ms(se_b200_map* map, int stage, float* ms);
/* per-stage CUDA event pairs for se_b200_elapsed_ms.  OFF by default: an event record between two kernel launches breaks the
 * programmatic-dependent-launch edge between them (every per-frame kernel overlaps its prologue with its predecessor's tail),
 * so timing costs ~10 % of the frame; enable != 0 turns the events on (benchmarks, DenseSLAMSystem::stageMilliseconds). */
int se_b200_set_stage_timing(se_b200_map* map, int enable);
/* counters of the last integrate: [0] nodes, [1] blocks, [2] active blocks, [3] error bits,
 * [4] blocks before the frame, [5] nodes before the frame, [6] octant requests (OFusion) */
int se_b200_counters(se_b200_map* map, int32_t out[8]);
/* number of kernels this library has launched on behalf of `map` so far */
int se_b200_launch_count(se_b200_map* map, int64_t* out);
/* device pointers of the map's images, for callers that keep data resident:
 * 0 float depth [W*H], 1 vertex [W*H*3], 2 normal [W*H*3] */
int se_b200_device_image(se_b200_map* map, int which, void** ptr);

#ifdef __cplusplus
}
#endif
#endif /* SE_B200_H */
