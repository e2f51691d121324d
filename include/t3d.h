/*
 * t3d.h — C ABI of libt3d.so, the B200 (sm_100a) dense geometric core.
 *
 * This is the drop-in boundary for the hot path behind the reference's
 * depth_to_reconstruction.py (d2r), depth_enhanced_reconstruction.py (der) and
 * depth_processor.py (dp).  The reference is pure Python and has no FFI of its
 * own; every entry point below cites the reference function it
 * replaces.  INTEGRATION.md shows the ctypes stub a reference maintainer adds.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no torch/C++ types.
 *   - All data pointers are DEVICE pointers unless the name ends in `_h`
 *     (host).  The caller owns every buffer; the library never frees caller
 *     memory and only allocates its own ctx / volume storage.
 *   - Every call takes a `t3d_stream` (a cudaStream_t passed as void*) and is
 *     asynchronous w.r.t. the host unless documented otherwise.
 *   - Return value: 0 = OK, negative = error (T3D_E_*).  t3d_last_error()
 *     returns a thread-local message.  No exceptions cross the boundary.
 *   - One t3d_ctx per GPU; a process may hold several (one per device).  Every
 *     entry point runs on ITS ctx's device whatever the caller's current device
 *     is, and restores the caller's current device before it returns.  A ctx is
 *     not thread-safe; distinct ctxs are.
 *   - ONE STREAM AT A TIME PER CTX: a ctx owns scratch that its asynchronous
 *     launches share (scan / ticket state of K1, the cached projection tables,
 *     hash tables and partials of K2/K3/K8).  Calls on the same ctx must be
 *     issued on one stream, or be ordered across streams by the caller (event
 *     or synchronise) — the library adds no cross-stream ordering of its own.
 *     A t3d_tsdf volume is the same: its sequence calls use one internal side
 *     stream that is ordered against the caller's stream on entry and exit.
 *   - There is NO CPU fallback: without a CUDA device every compute call
 *     returns T3D_E_CUDA.
 */
#ifndef T3D_H_
#define T3D_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define T3D_VERSION 100

#define T3D_OK 0
#define T3D_E_INVALID (-1)  /* bad argument */
#define T3D_E_CUDA (-2)     /* CUDA runtime error (message has the detail) */
#define T3D_E_CAPACITY (-3) /* hash table / block pool / output buffer too small */
#define T3D_E_IO (-4)       /* file I/O */
#define T3D_E_NUMERIC (-5)  /* voxel index overflow etc. */

typedef struct t3d_ctx t3d_ctx;
typedef struct t3d_tsdf t3d_tsdf;
typedef void* t3d_stream; /* cudaStream_t */

const char* t3d_last_error(void);
int t3d_version(void);

/* Create / destroy the per-GPU context (scratch buffers, scan state, cached
 * projection-factor tables = the reference's `_projection_cache`, d2r:285-295). */
t3d_ctx* t3d_create(int device);
void t3d_destroy(t3d_ctx* ctx);
/* Number of library kernels launched through this ctx (and its volumes) since
 * creation — bench.py's `gpu_launches` evidence. */
int64_t t3d_launch_count(const t3d_ctx* ctx);

/* ------------------------------------------------------------------------- */
/* K1 — back-projection.                                                      */
/* Replaces DenseReconstructor.depth_to_pointcloud (d2r:328-384),             */
/* DensePointCloudGenerator.depth_to_pointcloud (der:554-613) and             */
/* PointCloudGenerator.generate (dp:371-422).                                 */
/* ------------------------------------------------------------------------- */
typedef struct t3d_backproject_params {
  int32_t H, W;          /* full-resolution frame; depth is H*W C-order        */
  int32_t subsample;     /* [::s, ::s] slice (d2r:349-353), >= 1               */
  int32_t depth_is_f64;  /* 0: depth is float32, 1: float64 (der passes        */
                         /*    depths[i]*scale = f64, der:1135)                */
  int32_t scale_is_f64;  /* NumPy promotion of `depth*scale` (d2r:356): 0 =    */
                         /*    python-float scale (stays f32, thresholds cast  */
                         /*    to f32), 1 = np.float64 scale (everything f64)  */
  int32_t has_pose;      /* 0: pose=None (points stay in the camera frame)     */
  int32_t rgb_out_f32;   /* 0: uint8 RGB (d2r/der); 1: float32 RGB in [0,1]    */
                         /*    = u8.astype(f32)/255 (dp:417)                   */
  int32_t has_color;     /* 0: bgr==NULL, no colours written (dp rgb=None)     */
  double fx, fy, cx, cy; /* pinhole intrinsics (d2r:48-51)                     */
  double scale;          /* depth multiplier (d2r:356); 1.0 for der/dp         */
  double min_depth, max_depth; /* strict inequalities (d2r:359-361)           */
  double R[9];           /* world->camera rotation, row-major (d2r:371-376)    */
  double t[3];           /* world->camera translation                          */
} t3d_backproject_params;

/* depth: H*W (f32|f64), bgr: H*W*3 u8 (BGR, OpenCV order), conf_mask:
 * nullable H*W u8 (extension: pixel kept only if non-zero; the reference has no
 * confidence map).  out_xyz: capacity*3 f32; out_rgb: capacity*3
 * (u8|f32).  capacity must be >= ceil(H/s)*ceil(W/s) or the call fails.
 * out_n: device int64, number of valid points.  Output order = row-major order
 * of the valid sampled pixels (NumPy boolean-mask order, d2r:364-366). */
int t3d_backproject(t3d_ctx* ctx, const void* depth, const uint8_t* bgr,
                    const uint8_t* conf_mask, const t3d_backproject_params* p,
                    float* out_xyz, void* out_rgb, int64_t capacity,
                    int64_t* out_n, t3d_stream stream);

/* Batched K1: n_frames frames of identical geometry (H, W, subsample, intrinsics,
 * scale, depth range, dtype flags taken from `p`; p->R / p->t are ignored, each
 * frame carries its own pose) back-projected by ONE launch per 32 frames into ONE
 * concatenated, frame-ordered output — the reference's per-frame loop followed by
 * np.vstack (d2r:566-581,653-658 + d2r:401-402; der:1130-1145,1229-1237).
 * frames_h: host array.  out_offsets: device int64[n_frames + 1]; frame i owns
 * rows [out_offsets[i], out_offsets[i+1]).  capacity >= n_frames * ceil(H/s)*ceil(W/s).
 * Asynchronous. */
typedef struct t3d_backproject_frame {
  const void* depth;        /* H*W f32|f64 (device)                               */
  const uint8_t* bgr;       /* H*W*3 u8 BGR (device; NULL iff !p->has_color)      */
  const uint8_t* conf_mask; /* nullable H*W u8                                    */
  double R[9];              /* world->camera rotation, row-major (if p->has_pose) */
  double t[3];
} t3d_backproject_frame;

int t3d_backproject_batch(t3d_ctx* ctx, const t3d_backproject_frame* frames_h,
                          int n_frames, const t3d_backproject_params* p,
                          float* out_xyz, void* out_rgb, int64_t capacity,
                          int64_t* out_offsets, t3d_stream stream);

/* ------------------------------------------------------------------------- */
/* K2 — voxel-grid downsample.  Replaces Open3D                               */
/* PointCloud.voxel_down_sample as called by merge_pointclouds (d2r:405-410,  */
/* der:635-640).  Semantics: SURVEY §8c R2.                                   */
/* ------------------------------------------------------------------------- */
/* xyz: n*3 (f32|f64), rgb: n*3 u8 (nullable).  out_xyz: cap*3 f64 voxel
 * means; out_rgb: cap*3 u8 = trunc(mean(c/255)*255) (d2r:417-418); out_rgb_sum
 * (nullable) cap*3 u32 integer colour sums; out_count (nullable) cap u32;
 * out_vox_idx (nullable): cap*3 int32 voxel indices floor((p-minb)/v).
 * out_m: device int64.  min_bound_h (nullable, host): if given it is used as
 * the grid origin `minb` instead of min(p) - v/2 (multi-GPU: the global min).
 * sorted!=0: output in ascending (ix,iy,iz) order (deterministic); otherwise
 * hash-slot order.  Synchronous w.r.t. the host (needs the global bounds). */
int t3d_voxel_downsample(t3d_ctx* ctx, const void* xyz, int xyz_is_f64,
                         const uint8_t* rgb, int64_t n, double voxel,
                         const double* min_bound_h, int sorted,
                         double* out_xyz, uint8_t* out_rgb,
                         uint32_t* out_rgb_sum, uint32_t* out_count,
                         int32_t* out_vox_idx, int64_t capacity,
                         int64_t* out_m, double* out_min_bound_h,
                         t3d_stream stream);

/* Sharded K2 (owner = hash(voxel index) mod G, all_to_all of per-voxel partial sums,
 * owner merges).  Both calls take the GLOBAL grid: min_bound_h = min over all ranks - voxel/2,
 * max_bound_h = max over all ranks (exact all_reduce of t3d_bounds), so that every rank computes
 * identical voxel indices.
 * t3d_voxel_partials: this rank's points -> one 56-byte record per locally occupied voxel
 *   {int32 ix, iy, iz; uint32 count; double sum_x, sum_y, sum_z; uint32 sum_r, sum_g, sum_b, pad},
 *   grouped by owner (`world` owners); out_counts: device int32[world] records per owner;
 *   out_m: device int64 total.  capacity in records.
 * t3d_voxel_merge_partials: records received from every rank -> the owner's voxels, outputs as
 *   t3d_voxel_downsample.  f64 sums of f32 coordinates inside a voxel are exact, so the result is
 *   bit-identical to a single-GPU run over the union of the points. */
int t3d_voxel_partials(t3d_ctx* ctx, const void* xyz, int xyz_is_f64, const uint8_t* rgb,
                       int64_t n, double voxel, const double* min_bound_h,
                       const double* max_bound_h, int world, void* out_records,
                       int64_t capacity, int32_t* out_counts, int64_t* out_m,
                       t3d_stream stream);
int t3d_voxel_merge_partials(t3d_ctx* ctx, const void* records, int64_t n_records, int has_rgb,
                             double voxel, const double* min_bound_h,
                             const double* max_bound_h, int sorted, double* out_xyz,
                             uint8_t* out_rgb, uint32_t* out_rgb_sum, uint32_t* out_count,
                             int32_t* out_vox_idx, int64_t capacity, int64_t* out_m,
                             t3d_stream stream);

/* Min / max bound of a cloud (host result, synchronous). */
int t3d_bounds(t3d_ctx* ctx, const void* xyz, int xyz_is_f64, int64_t n,
               double* min_h, double* max_h, t3d_stream stream);

/* ------------------------------------------------------------------------- */
/* K3 — statistical outlier removal.  Replaces Open3D                         */
/* remove_statistical_outlier(nb_neighbors, std_ratio) (d2r:413-415), R3.     */
/* ------------------------------------------------------------------------- */
/* xyz: n*3 f64.  out_mean_dist (nullable): n f64 mean distance to the nb
 * nearest neighbours (self included).  keep_mask: n u8.  out_kept: device
 * int64.  stats_h (nullable, host, 3 doubles): mu, sigma, threshold.
 * Synchronous. */
int t3d_statistical_outlier(t3d_ctx* ctx, const double* xyz, int64_t n, int nb,
                            double std_ratio, double* out_mean_dist,
                            uint8_t* keep_mask, int64_t* out_kept,
                            double* stats_h, t3d_stream stream);

/* Sharded K3 (replicate the — small, downsampled — cloud, shard the queries).
 * t3d_sor_mean_distances_part: mean distance to the nb nearest neighbours for the points at
 *   grid-sorted positions [part*n/parts, (part+1)*n/parts) only; other entries of out_mean_dist
 *   (n doubles, indexed like xyz) stay untouched — initialise to -inf and all_reduce(MAX) the
 *   ranks' vectors.  The grid order is a stable sort of the full cloud, identical on every rank.
 * t3d_sor_from_mean_distances: mu, sigma, threshold and keep mask from a complete vector
 *   (stats_h: host double[3] = mu, sigma, threshold; nullable).  Both synchronous. */
int t3d_sor_mean_distances_part(t3d_ctx* ctx, const double* xyz, int64_t n, int nb, int part,
                                int parts, double* out_mean_dist, t3d_stream stream);
int t3d_sor_from_mean_distances(t3d_ctx* ctx, const double* mean_dist, int64_t n,
                                double std_ratio, uint8_t* keep_mask, int64_t* out_kept,
                                double* stats_h, t3d_stream stream);

/* Ordered compaction of rows by a byte mask (keeps input order, R3). */
int t3d_compact_rows(t3d_ctx* ctx, const void* rows, int64_t n,
                     int32_t row_bytes, const uint8_t* keep_mask,
                     void* out_rows, int64_t* out_n, t3d_stream stream);

/* ------------------------------------------------------------------------- */
/* K4/K5/K6 — TSDF voxel-block grid (north_star; Open3D                       */
/* t.geometry.VoxelBlockGrid semantics).  No reference code */
/* exists for this part: parity is against oracle/ only ("parity unpinned").  */
/* ------------------------------------------------------------------------- */
typedef struct t3d_tsdf_params {
  float voxel_size;      /* metres (config 2: 0.01)                            */
  float sdf_trunc;       /* metres (config 2: 0.04)                            */
  int32_t block_res;     /* must be 8                                          */
  int32_t pixel_round;   /* 0: ui=floor(u+0.5); 1: ui=(int)u       */
  int64_t block_capacity;/* max voxel blocks resident (10 KiB each)            */
  int64_t hash_capacity; /* 0 = auto (next pow2 >= 2*block_capacity)           */
} t3d_tsdf_params;

typedef struct t3d_frame_view {
  const void* depth;     /* H*W, f32 metres-scale or u16 raw                   */
  const uint8_t* bgr;    /* H*W*3 u8 BGR (nullable: no colour integration)     */
  float K[4];            /* fx, fy, cx, cy                                     */
  float T_cw[12];        /* world->camera extrinsic, row-major 3x4 [R|t]       */
  const uint8_t* conf_mask; /* H*W u8, nullable (north_star "confidence masking"; the reference has no
                          * confidence map): a pixel whose byte is 0 carries no measurement —
                          * K4 casts no ray through it and K5 updates no voxel from it, exactly as if its
                          * depth were 0.  Same convention as t3d_backproject's conf_mask.              */
} t3d_frame_view;

int t3d_tsdf_create(t3d_ctx* ctx, const t3d_tsdf_params* p, t3d_tsdf** out);
void t3d_tsdf_destroy(t3d_tsdf* v);
/* Forget all blocks (O(hash) memset; block storage is lazily re-initialised). */
int t3d_tsdf_reset(t3d_tsdf* v, t3d_stream stream);

/* Fuse `n_frames` (1..64) frames in one pass: K4 touch/allocate over all of
 * them, then K5 integrates every touched block, each voxel applying its
 * frames in index order (identical arithmetic to n_frames sequential calls).
 * frames_h: host array.  depth_is_u16: depth = raw/depth_scale.  Asynchronous. */
int t3d_tsdf_integrate(t3d_tsdf* v, const t3d_frame_view* frames_h,
                       int n_frames, int H, int W, int depth_is_u16,
                       float depth_scale, float depth_max, t3d_stream stream);

/* Fuse a whole sequence with known poses (config 2): batches of `batch`
 * (1..64) frames; K4 of batch b+1 runs on an internal stream underneath K5 of
 * batch b.  Bit-identical to calling t3d_tsdf_integrate batch by batch.
 * Work is ordered after prior work on `stream`, and `stream` ends ordered
 * after all of it.  Asynchronous. */
int t3d_tsdf_integrate_sequence(t3d_tsdf* v, const t3d_frame_view* frames_h,
                                int n_frames, int batch, int H, int W,
                                int depth_is_u16, float depth_scale,
                                float depth_max, t3d_stream stream);

/* The same pipeline with hooks, for block routing that overlaps fusion.  The caller orders the
 * frames so that every frame whose blocks must travel is in batches 0..hook_batch, and every frame that meets
 * incoming blocks is either in those batches too (then the merge needs no ordering against later batches) or in
 * the last batch (then wait_before_last orders it):
 *   nblocks_after_touch0 (device int32, nullable): the block count after K4 of batch `hook_batch` — every block
 *     batches 0..hook_batch touch exists and no allocation is in flight at that point;
 *   after_batch0(user, phase, event_a, event_b): called ON THE HOST, inside this call:
 *     phase b >= hook_batch, right after K5 of batch b was enqueued: event_a (a cudaEvent_t) completes with K4 of
 *       that batch (for b == hook_batch: from then on nblocks_after_touch0 and the keys of those blocks are final),
 *       event_b with its K5.  The callee enqueues its own work on another stream (t3d_stream_wait_event);
 *     phase -1, only when wait_before_last is given, right after K4 of the LAST batch was enqueued: event_a
 *       completes with it (event_b is NULL).  The callee enqueues what must run between K4 and K5 of the last batch
 *       and records wait_before_last before returning;
 *   wait_before_last (cudaEvent_t, nullable): K5 of the last batch waits for it (without a hook: K4 too).
 * The hook runs on the thread that feeds the pipeline: while it blocks (an event or stream synchronise) no later
 * batch is enqueued.  K5's persistent CTAs own every SM, so work the hook enqueues behind K4 of batch b only runs
 * when K5 of batch b retires — a hook that waits for such work in phase b stalls the GPU for a whole batch; wait
 * one phase later, when the next batch is already queued (CopyEngineBlockRouter.fuse_overlapped does).
 * Needs at least 2 batches when any hook is given. */
typedef void (*t3d_sequence_hook)(void* user, int phase, void* event_a, void* event_b);
int t3d_tsdf_integrate_sequence_hooked(t3d_tsdf* v, const t3d_frame_view* frames_h,
                                       int n_frames, int batch, int H, int W,
                                       int depth_is_u16, float depth_scale, float depth_max,
                                       int32_t* nblocks_after_touch0,
                                       t3d_sequence_hook after_batch0, void* user,
                                       void* wait_before_last, int hook_batch, t3d_stream stream);
/* Timing-less CUDA events for the hooks above (handles are cudaEvent_t). */
int t3d_event_create(void** out_event);
int t3d_event_destroy(void* event);
int t3d_event_record(void* event, t3d_stream stream);
int t3d_stream_wait_event(t3d_stream stream, void* event);

/* K4 alone: unique block keys touched by one frame (R4).  out_keys: cap*3
 * int32; out_n device int64.  Does not modify the volume. */
int t3d_tsdf_touch(t3d_tsdf* v, const t3d_frame_view* frame_h, int H, int W,
                   int depth_is_u16, float depth_scale, float depth_max,
                   int32_t* out_keys, int64_t capacity, int64_t* out_n,
                   t3d_stream stream);

/* Synchronous getters. */
int64_t t3d_tsdf_num_blocks(t3d_tsdf* v, t3d_stream stream);
/* counters_h (5 x int64): [0] voxel updates applied since create/reset
 * (sum of V_upd), [1] block-frame pairs integrated (sum of B), [2] frames
 * integrated, [3] voxels changed per block visit (a voxel updated by several
 * frames of one batch counts once), [4] block visits (one per block per batch). */
int t3d_tsdf_counters(t3d_tsdf* v, int64_t* counters_h, t3d_stream stream);

/* Per-kernel timing for bench.py's roofline: when enabled every
 * t3d_tsdf_integrate call is bracketed by CUDA events on its stream.
 * get_profile synchronises the stream and returns accumulated
 * {K4 touch ms, K5 integrate ms, integrate calls} since set_profiling. */
int t3d_tsdf_set_profiling(t3d_tsdf* v, int enable);
int t3d_tsdf_get_profile(t3d_tsdf* v, double* out3_h, t3d_stream stream);

/* Dump blocks for parity / routing.  keys: B*3 int32; tsdf, weight: B*512 f32
 * (x fastest, then y, z); rgb: B*512*3 f32 (0..255).  Any output may be NULL. */
int t3d_tsdf_export_blocks(t3d_tsdf* v, int32_t* keys, float* tsdf,
                           float* weight, float* rgb, int64_t capacity,
                           int64_t* out_b, t3d_stream stream);
/* Same, restricted to blocks whose key[axis] lies in [lo, hi) — the routing
 * record of the multi-GPU path (ownership = z-slab of the block key, SURVEY
 * §8e).  With all outputs NULL only the count is returned.  Synchronous. */
int t3d_tsdf_export_blocks_range(t3d_tsdf* v, int axis, int32_t lo, int32_t hi,
                                 int32_t* keys, float* tsdf, float* weight,
                                 float* rgb, int64_t capacity, int64_t* out_b,
                                 t3d_stream stream);
/* Complement of the above: blocks whose key[axis] is NOT in [lo, hi) — everything a
 * rank does not own, in one call.  If capacity is too small the call fails with
 * T3D_E_CAPACITY but *out_b still holds the required count. */
int t3d_tsdf_export_blocks_outside(t3d_tsdf* v, int axis, int32_t lo, int32_t hi,
                                   int32_t* keys, float* tsdf, float* weight,
                                   float* rgb, int64_t capacity, int64_t* out_b,
                                   t3d_stream stream);
/* Merge partial blocks into the volume (multi-GPU owner reduce):
 * w' = w_a + w_b, tsdf' = (w_a*tsdf_a + w_b*tsdf_b)/w', same for rgb; a voxel with
 * w_a == 0 (or w_b == 0) takes the other side's values unchanged, so restoring a dump
 * into an empty volume is bit-exact. */
int t3d_tsdf_merge_blocks(t3d_tsdf* v, const int32_t* keys, const float* tsdf,
                          const float* weight, const float* rgb, int64_t b,
                          t3d_stream stream);

/* Multi-GPU block routing in wire format (owner(key) = clamp(floor(key[axis] /
 * slab_blocks), 0, world-1)).  A record is 2564 f32 words = 10 256 B:
 *   [kx ky kz owner (int32 bits) | tsdf x512 | weight x512 | rgb x1536 voxel-major].
 * 1. t3d_tsdf_route_counts: counts[d] = number of blocks owned by rank d != self (device).
 * 2. t3d_tsdf_route_export: writes every non-owned block as a record at row
 *    dst_base[owner] + k (dst_base = exclusive scan of counts), i.e. grouped by destination —
 *    directly usable as the send buffer of a variable-size all-to-all.
 * 3. t3d_tsdf_merge_records: merges received records (keys unique within one call) with
 *    the t3d_tsdf_merge_blocks rule.
 * All three are asynchronous (the block count is read on the device). */
int t3d_tsdf_route_counts(t3d_tsdf* v, int axis, int32_t slab_blocks, int world,
                          int self_rank, int32_t* counts, t3d_stream stream);
int t3d_tsdf_route_export(t3d_tsdf* v, int axis, int32_t slab_blocks, int world,
                          int self_rank, const int32_t* dst_base, int32_t* dst_fill,
                          float* records, t3d_stream stream);
int t3d_tsdf_merge_records(t3d_tsdf* v, const float* records, int64_t b,
                           t3d_stream stream);
/* The same two with the block count capped by *n_blocks_dev (device int32, nullable: no cap) — routing
 * that overlaps fusion only looks at the blocks that existed after K4 of batch 0. */
int t3d_tsdf_route_counts_upto(t3d_tsdf* v, int axis, int32_t slab_blocks, int world,
                               int self_rank, const int32_t* n_blocks_dev, int32_t* counts,
                               t3d_stream stream);
int t3d_tsdf_route_export_upto(t3d_tsdf* v, int axis, int32_t slab_blocks, int world,
                               int self_rank, const int32_t* n_blocks_dev,
                               const int32_t* dst_base, int32_t* dst_fill, float* records,
                               t3d_stream stream);
/* Copy-engine routing (the default on one NVLink box): records are packed into a LOCAL send buffer
 * (t3d_tsdf_route_export_upto), each destination's group is pushed into that owner's receive buffer with
 * one t3d_memcpy_async (peer copy: NVLink through the copy engine, no SM involved) plus a 4-byte copy of
 * the count, and after a barrier the owner merges every source's region with ONE call:
 *   recv_base: [int32 count_from[world] | pad to header_bytes][region 0]...[region world-1],
 *   region s = region_bytes bytes = up to region_records records from rank s (own rank: skipped).
 * Blocks arriving from several sources are merged under a per-block lock. */
int t3d_tsdf_merge_records_multi(t3d_tsdf* v, const void* recv_base, int64_t header_bytes,
                                 int64_t region_bytes, int world, int self_rank,
                                 int64_t region_records, t3d_stream stream);

/* Fused export + transfer over peer memory (NVLink/NVSwitch): every non-owned block is stored
 * straight into its owner's receive region — peer_regions_h[d] is, in rank d's memory (opened
 * with t3d_ipc_open), the region reserved for records coming from THIS rank; peer_counts_h[d]
 * the int32 header slot there for this rank's record count, written by the last CTA.
 * local_fill: device int32[world + 2] scratch ([world] = CTA ticket, [world+1] = records that
 * did not fit).  n_blocks_dev (device int32, nullable): only blocks [0, *n_blocks_dev) are
 * examined (see t3d_tsdf_integrate_sequence_hooked).  The owners then call
 * t3d_tsdf_merge_records_dev on their own memory after a barrier.  Asynchronous; no staging copy,
 * no collective on the data path. */
int t3d_tsdf_route_export_p2p(t3d_tsdf* v, int axis, int32_t slab_blocks, int world,
                              int self_rank, void* const* peer_regions_h,
                              int32_t* const* peer_counts_h, int64_t region_records,
                              int32_t* local_fill, const int32_t* n_blocks_dev,
                              t3d_stream stream);
/* t3d_tsdf_merge_records with the record count read from device memory (<= max_b). */
int t3d_tsdf_merge_records_dev(t3d_tsdf* v, const float* records, const int32_t* count_dev,
                               int64_t max_b, t3d_stream stream);

/* Peer-visible device buffers (CUDA IPC).  alloc: cudaMalloc + zero + 64-byte handle to
 * hand to the other ranks; open/close: map/unmap a peer's buffer; free: release one's own. */
int t3d_ipc_alloc(t3d_ctx* ctx, size_t bytes, void** dev_ptr, uint8_t* handle_out64);
int t3d_ipc_open(t3d_ctx* ctx, const uint8_t* handle64, void** dev_ptr);
int t3d_ipc_close(t3d_ctx* ctx, void* dev_ptr);
int t3d_ipc_free(t3d_ctx* ctx, void* dev_ptr);
int t3d_memset_async(void* dev_ptr, int value, size_t bytes, t3d_stream stream);
/* cudaMemcpyAsync(cudaMemcpyDefault) between device buffers, peers included. */
int t3d_memcpy_async(void* dst, const void* src, size_t bytes, t3d_stream stream);

/* K6: surface points (R6).  xyz/nrm: cap*3 f32; rgb: cap*3 u8 (nullable
 * nrm/rgb).  out_n device int64.  Order: deterministic only as a set. */
int t3d_tsdf_extract_points(t3d_tsdf* v, float weight_threshold, float* xyz,
                            float* nrm, uint8_t* rgb, int64_t capacity,
                            int64_t* out_n, t3d_stream stream);

/* K6 restricted to the blocks a camera can see (frame-to-model tracking, cfg 3:
 * the ICP target is the model surface inside the predicted view).  A block is
 * kept iff its bounding sphere (centre (key+0.5)*8*voxel, radius sqrt(3)/2*8*voxel)
 * reaches into depth (0, depth_max) and into the image rectangle of view_h
 * (K, T_cw; depth/bgr pointers ignored).  Asynchronous — block and point counts stay on
 * the device — unless out_blocks_h (nullable, host: number of blocks selected) is given,
 * which costs one stream synchronisation.  If *out_n > capacity the output was truncated. */
int t3d_tsdf_extract_points_view(t3d_tsdf* v, const t3d_frame_view* view_h, int H,
                                 int W, float depth_max, float weight_threshold,
                                 float* xyz, float* nrm, uint8_t* rgb,
                                 int64_t capacity, int64_t* out_n,
                                 int64_t* out_blocks_h, t3d_stream stream);

/* K10 — triangle-mesh extraction (north_star "surface/point extraction to .ply", SURVEY 8f
 * rank 3; no reference code — semantics of Open3D VoxelBlockGrid.extract_triangle_mesh, R6m in
 * the oracle).  A cube (voxel + its 7 upper neighbours) is valid iff all 8 voxels have
 * W >= weight_threshold; case bit i = tsdf(corner i) < 0; every sign-changing edge of a valid
 * cube carries one vertex with R6's position / normal / colour; triangles from the table in
 * csrc/mc_tables.cuh, wound so that their normals follow the TSDF gradient.
 * xyz, nrm: vertex_capacity*3 f32; rgb: vertex_capacity*3 u8 (nrm, rgb nullable);
 * tri: triangle_capacity*3 int32 vertex indices.  out_counts (device int64[2]) = {vertices,
 * triangles}; a count above its capacity means that output was truncated.  Pass both capacities
 * as 0 to count only.  Vertex / triangle order is unspecified.  Synchronous (block count). */
int t3d_tsdf_extract_mesh(t3d_tsdf* v, float weight_threshold, float* xyz, float* nrm,
                          uint8_t* rgb, int64_t vertex_capacity, int32_t* tri,
                          int64_t triangle_capacity, int64_t* out_counts,
                          t3d_stream stream);

/* K6 over the blocks with key[axis] in [lo, hi) only (each owner extracts its own slab;
 * neighbour tests and gradients read every block present, i.e. the halo fetched from the
 * neighbouring ranks).  Asynchronous apart from the block count. */
int t3d_tsdf_extract_points_range(t3d_tsdf* v, int axis, int32_t lo, int32_t hi,
                                  float weight_threshold, float* xyz, float* nrm,
                                  uint8_t* rgb, int64_t capacity, int64_t* out_n,
                                  t3d_stream stream);

/* ------------------------------------------------------------------------- */
/* K7 — normal estimation (north_star; Open3D estimate_normals KNN, R7).      */
/* ------------------------------------------------------------------------- */
/* xyz: n*3 f32; nrm: n*3 f32.  orient_to_h (nullable host double[3]): flip
 * each normal towards that camera location.  Synchronous (grid build). */
int t3d_estimate_normals(t3d_ctx* ctx, const float* xyz, int64_t n, int knn,
                         const double* orient_to_h, float* nrm,
                         t3d_stream stream);

/* ------------------------------------------------------------------------- */
/* K8 — point-to-plane ICP (north_star; Open3D registration_icp with          */
/* TransformationEstimationPointToPlane, R8).                                 */
/* ------------------------------------------------------------------------- */
typedef struct t3d_icp_result {
  double T[16];       /* source->target transform, row-major 4x4               */
  double fitness;     /* |C| / |S|                                             */
  double inlier_rmse; /* sqrt(sum d^2 / |C|)                                   */
  int32_t iterations; /* iterations executed                                   */
  int32_t converged;
  int64_t correspondences;
} t3d_icp_result;

/* src: n_src*3 f32; tgt, tgt_nrm: n_tgt*3 f32.  T0_h: host row-major 4x4.
 * The whole registration runs on the device (correspondences, f64 normal equations,
 * 6x6 solve with the R8 determinant guard, convergence test); the host reads one result
 * record.  Synchronous. */
int t3d_icp_point_to_plane(t3d_ctx* ctx, const float* src, int64_t n_src,
                           const float* tgt, const float* tgt_nrm,
                           int64_t n_tgt, double max_corr_dist,
                           const double* T0_h, int max_iter,
                           double rel_fitness, double rel_rmse,
                           t3d_icp_result* result_h, t3d_stream stream);

/* The correspondence search of the device-resident registration on its own (the kernel R8's iterations spend
 * their time in): for every source point the ORIGINAL index of the nearest target point within max_corr_dist of
 * T0 * p (-1: none; ties -> lowest index) and the squared distance (f64).  With T1 != NULL a second search at T1,
 * seeded with the first one's answers exactly as a registration's later iterations are, replaces the result.
 * Replaces: open3d KDTreeFlann::SearchHybrid(knn = 1) inside registration_icp. */
int t3d_icp_correspondences(t3d_ctx* ctx, const float* src, int64_t n_src, const float* tgt, const float* tgt_nrm,
                            int64_t n_tgt, double max_corr_dist, const double* T0_h, const double* T1_h,
                            int32_t* out_idx, double* out_d2, t3d_stream stream);

/* Same registration with the cloud sizes in DEVICE memory (int64): src/tgt buffers hold up
 * to src_capacity / tgt_capacity points, of which *n_src_dev / *n_tgt_dev are valid.  Made for
 * frame-to-model tracking, where both clouds were just produced on the device (K1, K6): their
 * sizes never visit the host.  If either cloud has fewer than min_points points no registration
 * is attempted: T = T0 and *skipped_h = 1. */
int t3d_icp_point_to_plane_dev(t3d_ctx* ctx, const float* src, int64_t src_capacity,
                               const int64_t* n_src_dev, const float* tgt,
                               const float* tgt_nrm, int64_t tgt_capacity,
                               const int64_t* n_tgt_dev, int min_points,
                               double max_corr_dist, const double* T0_h, int max_iter,
                               double rel_fitness, double rel_rmse,
                               t3d_icp_result* result_h, int* skipped_h, t3d_stream stream);

/* One ICP linearisation: correspondences + 6x6 normal equations only.
 * out27_h: 21 upper-triangular JtJ + 6 Jtr; out_stats_h: sum r^2 (squared
 * point distance), count.  Used by the multi-GPU path (all-reduce of 29
 * doubles between linearise and solve). */
int t3d_icp_linearize(t3d_ctx* ctx, const float* src, int64_t n_src,
                      const float* tgt, const float* tgt_nrm, int64_t n_tgt,
                      double max_corr_dist, const double* T_h,
                      double* out27_h, double* out_stats_h,
                      t3d_stream stream);

/* Nearest neighbour (k=1) within radius: idx n_q int32 (-1 = none), d2 n_q f32. */
int t3d_nearest_neighbor(t3d_ctx* ctx, const float* query, int64_t n_q,
                         const float* ref, int64_t n_ref, double radius,
                         int32_t* out_idx, float* out_d2, t3d_stream stream);

/* ------------------------------------------------------------------------- */
/* K9 — PLY writers (host).  Replace save_reconstruction (d2r:673-703),       */
/* _save_pointcloud (der:1283-1311), save_ply (dp:424-440).                   */
/* ------------------------------------------------------------------------- */
#define T3D_PLY_O3D_BINARY 0 /* Open3D write_point_cloud layout (R9)          */
#define T3D_PLY_REF_ASCII 1  /* in-repo fallback layout (d2r:690-701)         */
/* xyz_h: n*3 (f32|f64), rgb_h: n*3 u8 (nullable), nrm_h: n*3 f32|f64 same
 * dtype as xyz (nullable, binary layout only). */
int t3d_write_ply_h(const char* path, const void* xyz_h, int xyz_is_f64,
                    const uint8_t* rgb_h, const void* nrm_h, int64_t n,
                    int layout);

/* Triangle mesh as Open3D's write_triangle_mesh writes it: binary little-endian, vertex
 * `double x y z [nx ny nz]` + `uchar red green blue`, then `element face` with
 * `property list uchar uint vertex_indices`.  xyz_h, nrm_h: nv*3 f32; tri_h: nt*3 int32. */
int t3d_write_ply_mesh_h(const char* path, const float* xyz_h, const float* nrm_h,
                         const uint8_t* rgb_h, int64_t nv, const int32_t* tri_h,
                         int64_t nt);

/* ------------------------------------------------------------------------- */
/* Formats either side of the path.                              */
/* ------------------------------------------------------------------------- */
/* 16-bit millimetre depth -> float32: out = (float)raw / divisor — the reader's
 * `raw.astype(np.float32) / 1000.0` (d2r:85-90) on the GPU (a depth PNG then crosses
 * PCIe at 2 B/pixel). */
int t3d_depth_u16_to_f32(t3d_ctx* ctx, const uint16_t* raw, int64_t n, float divisor,
                         float* out, t3d_stream stream);
/* float32 depth -> 16-bit: out = (depth * factor).astype(np.uint16), the writer's
 * `(depth * 1000).astype(np.uint16)` (dp:919-921) including its wrap-around for
 * out-of-range / non-finite values (x86-64 NumPy conversion). */
int t3d_depth_f32_to_u16(t3d_ctx* ctx, const float* depth, int64_t n, float factor,
                         uint16_t* out, t3d_stream stream);
/* cv2.resize(depth, (dst_w, dst_h), interpolation=cv2.INTER_LINEAR) for float32
 * single-channel images (d2r:465-467).  Pinned against OpenCV 4.13 outputs
 * (tests/golden/formats.npz); tolerance 1e-6 relative (SIMD contraction differs). */
int t3d_resize_bilinear_f32(t3d_ctx* ctx, const float* src, int src_h, int src_w,
                            float* dst, int dst_h, int dst_w, t3d_stream stream);
/* Depth-scale estimation (d2r:297-326 with gate=1; der:652-697 with gate=0 and
 * min_input_points=5): median over sparse points of Z_i / depth[int(v_i), int(u_i)].
 * depth: device H*W f32.  pts3d_h (n*3) / pts2d_h (n*2): host f64.  Returns 1.0 with
 * fewer than 3 accepted samples.  Synchronous. */
int t3d_estimate_scale(t3d_ctx* ctx, const float* depth, int H, int W,
                       const double* pts3d_h, const double* pts2d_h, int64_t n,
                       int gate, int min_input_points, double* out_scale_h,
                       int64_t* out_samples_h, t3d_stream stream);
/* PointCloud2 records (dp:744-758): out[i] = {x, y, z, rgb} with rgb = the bytes
 * (b, g, r, 0) reinterpreted as float32; colours are f32 in [0,1] (r,g,b =
 * (c*255).astype(uint8), colors_are_f32 != 0) or u8 RGB.  out_records: n*4 f32. */
int t3d_pack_pointcloud2(t3d_ctx* ctx, const float* xyz, const void* colors,
                         int colors_are_f32, int64_t n, float* out_records,
                         t3d_stream stream);

/* ------------------------------------------------------------------------- */
/* Synthetic scenes: device-side generators used by bench.py and */
/* the full-size property tests; tests compare them with the NumPy generator. */
/* ------------------------------------------------------------------------- */
/* scene 0 = tunnel T1 (cfg 2/3/5), scene 1 = relief plane S1 (cfg 1).
 * Writes depth (H*W f32, metres) and bgr (H*W*3 u8) for frame `frame_index`
 * and the world->camera pose [R|t] into T_cw_h (host, 12 doubles row-major 3x4,
 * nullable).  depth==NULL: pose only. */
int t3d_synth_frame(t3d_ctx* ctx, int scene, int frame_index, int H, int W,
                    double fx, double fy, double cx, double cy, uint64_t seed,
                    float noise_sigma, float* depth, uint8_t* bgr,
                    double* T_cw_h, t3d_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* T3D_H_ */
