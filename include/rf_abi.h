/*
 * rf_abi.h — C-ABI of librf_b200.so, the B200-native (sm_100a) mapping hot path of RemixFusion.
 *
 * The reference has no FFI for this path: it JIT-compiles CUDA-C strings with PyCUDA and launches them
 * on raw device pointers, and it reaches tiny-cuda-nn through its torch binding.  Every entry point
 * below names the reference launch / module call it replaces (file:line relative to the reference
 * tree).  Conventions:
 *
 *   - extern "C", plain pointers and sizes, no torch types;
 *   - pointers documented "device" are caller-owned device pointers (a torch tensor's data_ptr());
 *     pointers documented "host" are small parameter blocks read synchronously during the call;
 *   - nothing is allocated, nothing is retained after the call returns;
 *   - the last argument is the cudaStream_t (as void*) the work is enqueued on; calls are asynchronous;
 *   - return value: 0 = ok, <0 = bad argument (RF_E_*), >0 = cudaError_t of the failed launch;
 *     rf_last_error() returns a thread-local human-readable string for the last non-zero return.
 */
#ifndef RF_ABI_H
#define RF_ABI_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RF_ABI_VERSION 1

#define RF_E_NULL      (-1)  /* required pointer is NULL                       */
#define RF_E_RANGE     (-2)  /* size / slab / flag out of range                */
#define RF_E_ALIGN     (-3)  /* pointer not aligned as the layout requires     */
#define RF_E_UNSUPPORTED (-4)/* shape outside what the kernels are built for   */

int         rf_version(void);
const char* rf_last_error(void);

/* ------------------------------------------------------------------------------------------------
 * Stage 1a — local moving volume (tracker).  Replaces the PyCUDA launch of `integrate`
 * model/Volume.py:729-756 (kernel model/Volume.py:196-336).
 *
 * Layout: three fp32 arrays of dx*dy*dz elements, index = z + y*dz + x*dy*dz (z fastest).
 * packed_bgr is the folded colour image floor(B*65536 + G*256 + R) of model/Volume.py:728.
 * origin is truncated toward zero to an integer exactly as the kernel does (model/Volume.py:230-232).
 * x0,x1: x-slab [x0,x1) owned by the caller (multi-GPU sharding; 0,dx for the whole volume).  The
 * array pointers always address the FULL volume's element 0 unless slab_local != 0, in which case
 * they address the first element of slab x0 (a rank that only allocates its own slab).
 * ---------------------------------------------------------------------------------------------- */
int rf_tsdf_integrate_local(float* tsdf, float* weight, float* color,       /* device, in/out */
                            int dx, int dy, int dz,
                            const float origin[3],                           /* host */
                            float voxel_size,
                            const float K[9],                                /* host, row-major 3x3 */
                            const float c2w[16],                             /* host, row-major 4x4 */
                            const float* depth,                              /* device [H*W] metres */
                            const float* packed_bgr,                         /* device [H*W] */
                            int H, int W,
                            float trunc_margin, float obs_weight,
                            int weight_clamp, int reintegrate,
                            const float old_bnd[6],                          /* host; may be NULL if !reintegrate */
                            int x0, int x1, int slab_local,
                            const float* rcp_lambda,                         /* device [H*W] from rf_tsdf_pixel_lambda, or NULL */
                            const float* depth_max,                          /* device scalar from rf_tsdf_depth_max, or NULL */
                            void* stream);

/* N2 — re-centring of the moving volume when the camera leaves it.  Replaces `copy_volume` +
 * `swap_rot_trans` (model/Volume.py:585-610, :128-194; host :796-858, :883-908): every voxel of the NEW volume
 * (dims dx,dy,dz at `origin`) takes the value of the nearest OLD voxel at the same world position, or (1,0,0) outside
 * the old volume.  The caller keeps two sets of arrays and swaps them (no backup copy: 24 instead of 48 B/voxel);
 * new and old arrays must not alias.  `origin` / `old_origin` are the float bounds the reference passes (not truncated). */
int rf_tsdf_recenter(float* tsdf, float* weight, float* color,                       /* device, written: new volume */
                     const float* old_tsdf, const float* old_weight, const float* old_color,
                     int dx, int dy, int dz, const float origin[3],
                     int odx, int ody, int odz, const float old_origin[3],
                     float voxel_size, void* stream);

/* N2, second half — the volume-reading kernels of the random-optimisation tracker.
 *
 * rf_track_vertex_normal replaces `compute_vertex` + `compute_normal` (model/ROtracker.py:273-344, :346-400; host
 * init_depth_vertex :436-456, init_normal :458-470): depth_vertex [H*W][4] = back-projected vertex at depth + a random
 * offset along z, and the TSDF value that offset implies; normal [H*W][3] from central differences (border pixels are not
 * written: zero-fill the map once, as the reference's allocation does).  The reference seeds curand with subsequence =
 * row index, i.e. one random offset per image ROW; `row_sample` (device, H floats) receives those offsets.  `seed` is the
 * reference's `seed_num`.  All pointers are device pointers; K is a host array. */
int rf_track_vertex_normal(const float* depth, int H, int W, const float K[9], float cut_dist, float trunc, int seed,
                           float sample_range, float* row_sample, float* depth_vertex, float* normal, void* stream);

/* rf_track_fitness replaces `compute_tsdf_value` (model/ROtracker.py:144-271; host evaluate_tsdf :536-604): for each of
 * the n pose candidates (candidates [n][6] = translation + quaternion vector part, in units of search_size), transform
 * every `level`-th valid vertex (offset level_index) by candidate o current pose (R, T), and accumulate
 * |tsdf_vol[nearest voxel] - expected tsdf| into search_value[n] and the hit count into search_count[n] (both float,
 * overwritten).  Deterministic (fixed summation order), unlike the reference's system-scope atomics.  vol_origin is
 * truncated to int as the reference kernel does (:163-165).  scratch: rf_track_fitness_scratch_floats() floats, 8-byte
 * aligned.  vol_dim, vol_origin, K, R, T, search_size are host arrays; the rest device pointers. */
int rf_track_fitness(const float* tsdf_vol, const int vol_dim[3], const float vol_origin[3], float voxel_size,
                     const float* depth_vertex, const float* normal, int H, int W, const float K[9],
                     const float R[9], const float T[3], const float* candidates, int n_candidates,
                     const float search_size[6], int level, int level_index,
                     float* search_value, float* search_count, float* scratch, void* stream);
int64_t rf_track_fitness_scratch_floats(int n_candidates, int H, int W, int level);

/* rf_track_cal_transform replaces the Python loop of `cal_transform` (model/ROtracker.py:606-714) on the arrays the call
 * above left on the device: fit_j = search_value[j] / (search_count[j] + 1e-6) (evaluate_tsdf :604); among candidates
 * 1..n-1 with fit_j < fit_0, the first `count_search` in index order are averaged with weights fit_0 - fit_j.
 * out9 (device, 9 floats): success flag (0 / 1), min_tsdf, mean_transform = (tx, ty, tz, qw, qx, qy, qz).
 * No qualifying candidate: (0, fit_0, zeros).  search_size is a host array. */
int rf_track_cal_transform(const float* search_value, const float* search_count, const float* candidates, int n_candidates,
                           const float search_size[6], int count_search, float* out9, void* stream);

/* rf_track_random_optimization replaces the 20-iteration search loop of `RO_tracker.random_optimization`
 * (model/ROtracker.py:716-836) including `get_PST` (:467-493), `evaluate_tsdf` (:536-604), `cal_transform` (:606-714) and
 * `update_PST` (:495-531): every iteration is four launches (fitness, fold, cal_transform, policy) on `stream` with the search
 * state kept on the device, so nothing is read back between iterations (the reference reads 2 x n floats and uploads the search
 * size every iteration).  `pst`: all candidate tables [.., 6] (device); for count_particle k = 0..19 the table starts at float
 * offset pst_offset[k] and holds pst_n[k] candidates (the reference's tiff_index / PST_size), evaluated at pyramid level
 * depth_level[k].  `state` (device, 64 floats, 16-byte aligned): in  [0,9) current_global_R, [9,12) current_global_T,
 * [12,18) search_size, [18,24) previous_search_size, words 39.. = {count_particle 0, level_index 5, 0, 0, 0, 0, n, level,
 * cand_off, chunks of k = 0} as rf_track_state_init fills them; out  the optimised R / T, the final search sizes, word 44 =
 * success of iteration 0 (previous_frame_success), word 49 = bit mask of the successful iterations.
 * search_value / search_count: max(pst_n) floats each; scratch: rf_track_random_optimization_scratch_floats() floats. */
int rf_track_random_optimization(const float* tsdf_vol, const int vol_dim[3], const float vol_origin[3], float voxel_size,
                                 const float* depth_vertex, const float* normal, int H, int W, const float K[9],
                                 const float* pst, const int pst_offset[20], const int pst_n[20], const int depth_level[20],
                                 int iters, int count_search, float scaling_coefficient, int fix_level_index, int iterative_scale,
                                 float beta, float* state, float* search_value, float* search_count, float* scratch, void* stream);
int64_t rf_track_random_optimization_scratch_floats(const int pst_n[20], const int depth_level[20], int H, int W);
/* Host helper: fills the 64-word state image (host memory) for the first iteration. */
int rf_track_state_init(float* state_host, const float R[9], const float T[3], const float search_size[6],
                        const float previous_search_size[6], const int pst_offset[20], const int pst_n[20], const int depth_level[20],
                        int H, int W);

/* The per-pixel factor 1/sqrt(vx^2 + vy^2 + 1), vx = (px - cx)/fx, vy = (py - cy)/fy, of the projective SDF
 * (model/Volume.py:280-283, mp_slam/mapper.py:108-111).  It depends on the intrinsics only, so a caller computes it
 * once per camera and passes it to every integrate; the kernels then load it next to the depth instead of spending two
 * IEEE divides, a square root and a reciprocal per projected voxel.  Same operations in the same order: bit-identical. */
int rf_tsdf_pixel_lambda(const float K[9], int H, int W, float* rcp_lambda /*device [H*W]*/, void* stream);

/* model/Volume.py:723-728 — fold an RGB float image (values 0..255) into packed BGR, on device. */
int rf_pack_bgr(const float* rgb_hw3 /*device [H*W*3]*/, float* packed /*device [H*W]*/, int n_pixels, void* stream);

/* model/Volume.py:561-583 (`clean_tsdf`): tsdf=1, weight=0, colour=0 over n voxels. */
int rf_tsdf_clear_local(float* tsdf, float* weight, float* color, int64_t n_voxels, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Stage 1b — global coarse volume "GBV" (mapper).  Replaces the PyCUDA launch of `map_integrate`
 * mp_slam/mapper.py:843-872 (kernel mp_slam/mapper.py:37-158).
 *
 * Layout: trgb = tiny-cuda-nn Dense-grid params, [R^3][4] fp32 AoS (tsdf,r,g,b), voxel index
 * v = x + y*R + z*R*R (x fastest); wgt = [R^3] fp32.  box = {x0,x1,y0,y1,z0,z1} (mapping.bound).
 * rgb_hw3 is the interleaved float RGB image in [0,1].  z0,z1: z-slab [z0,z1) owned by the caller;
 * slab_local as above.  c2w may live on the device (the reference passes a device tensor,
 * mp_slam/mapper.py:849): set c2w_on_device != 0.
 * ---------------------------------------------------------------------------------------------- */
int rf_tsdf_integrate_global(float* trgb, float* wgt,                        /* device, in/out */
                             int R,
                             const float box[6],                             /* host */
                             const float K[9],                               /* host */
                             const float* c2w, int c2w_on_device,
                             const float* depth,                             /* device [H*W] */
                             const float* rgb_hw3,                           /* device [H*W*3] */
                             int H, int W,
                             float trunc_margin, float obs_weight,
                             int z0, int z1, int slab_local,
                             const float* rcp_lambda,                        /* device [H*W] or NULL (see above) */
                             const float* depth_max,                         /* device scalar or NULL (rf_tsdf_depth_max) */
                             void* stream);

/* Largest depth of the frame -> depth_max (device scalar).  Passing it to the integrate calls adds a far plane to the row
 * clip: no voxel farther than (depth_max + trunc)(1 + 0.5/fx + 0.5/fy) from the camera plane can pass the reference's
 * `sdf >= -trunc` test, so the sweep stops behind the farthest surface.  Results are identical with and without it. */
int rf_tsdf_depth_max(const float* depth, int64_t n, float* depth_max, void* stream);

/* mp_slam/mapper.py:161-183 + :267-282 (`clean_tsdf` / init_mapvolume): trgb[v] = (1,0,0,0). */
int rf_tsdf_clear_global(float* trgb, int64_t n_voxels, void* stream);

/* Count voxels a frame would touch (the metric's numerator): same predicate as the
 * integrate kernels, no writes.  counts (device, uint64[2]) += {n_touched, n_band}. */
int rf_tsdf_count_local(int dx, int dy, int dz, const float origin[3], float voxel_size,
                        const float K[9], const float c2w[16], const float* depth, int H, int W,
                        float trunc_margin, int reintegrate, const float old_bnd[6],
                        int x0, int x1, unsigned long long* counts, void* stream);
int rf_tsdf_count_global(int R, const float box[6], const float K[9], const float* c2w, int c2w_on_device,
                         const float* depth, int H, int W, float trunc_margin,
                         const float* trgb, const float* wgt, float obs_weight,
                         int z0, int z1, int slab_local, unsigned long long* counts, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Stage 2 — encoders (what `tcnn.Encoding.forward/backward` did; model/encodings.py:33-51,65-76,
 * model/scene_rep.py:60-93).  tiny-cuda-nn semantics: SURVEY.md Appendix B.
 * ---------------------------------------------------------------------------------------------- */
#define RF_MAX_LEVELS 16

typedef struct rf_grid_desc {
    int32_t  n_levels;                    /* <= RF_MAX_LEVELS */
    int32_t  n_features;                  /* F: 1, 2 or 4 */
    int32_t  is_hash;                     /* 1: HashGrid, 0: Dense */
    int32_t  _pad;
    float    scale[RF_MAX_LEVELS];        /* exp2f(l*log2f(s))*base - 1 */
    uint32_t resolution[RF_MAX_LEVELS];   /* ceilf(scale)+1 */
    uint32_t size[RF_MAX_LEVELS];         /* entries in the level (after hash cap / x8 round-up) */
    uint32_t offset[RF_MAX_LEVELS + 1];   /* running sum of size[] (entries, not floats) */
} rf_grid_desc;

/* Fill a descriptor exactly as tiny-cuda-nn's GridEncoding constructor does (Appendix B1). */
int rf_grid_desc_init(rf_grid_desc* d, int n_levels, int n_features, int is_hash,
                      int log2_hashmap_size, int base_resolution, double per_level_scale);

/* x: device [n,3] fp32 in normalised coords; params: device fp32 [F*offset[n_levels]];
 * out: device [n, n_levels*F] fp32 row-major. */
int rf_grid_encode_forward(const rf_grid_desc* d, const float* params, const float* x, int64_t n,
                           float* out, void* stream);
/* grad_params (device, same shape as params) += scatter of dout; dx (device [n,3]) written when non-NULL. */
int rf_grid_encode_backward(const rf_grid_desc* d, const float* params, const float* x, int64_t n,
                            const float* dout, float* grad_params, float* dx, void* stream);
/* OneBlob, n_bins per coordinate, out [n, 3*n_bins]. */
int rf_oneblob_forward(const float* x, int64_t n, int n_bins, float* out, void* stream);
int rf_oneblob_backward(const float* x, int64_t n, int n_bins, const float* dout, float* dx, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Stage 2 — fused mixed-representation ray query + render (+ losses) and its backward.
 * Replaces JointEncoding.render_rays/run_network/query_color_sdf/raw2outputs/sdf2weights
 * (model/scene_rep.py:107-127,156-179,314-349,370-456), ColorSDFNet.forward (model/decoder.py:132-146)
 * and, in training mode, the losses of JointEncoding.mapping (model/scene_rep.py:493-517,
 * model/utils.py:170-256).
 * ---------------------------------------------------------------------------------------------- */
typedef struct rf_ray_cfg {
    /* sampling (model/scene_rep.py:417-441) */
    float   range_d;        /* training.range_d */
    int32_t n_range_d;      /* training.n_range_d */
    int32_t n_samples_d;    /* training.n_samples_d */
    float   near_, far_;    /* cam.near, cam.far */
    int32_t perturb;        /* training.perturb > 0 */
    /* query (model/scene_rep.py:329-345) */
    float   c_trunc;        /* training.c_trunc */
    float   trunc;          /* training.trunc */
    int32_t clamp_mode;     /* 0: mapping (clamp +-1); 1: BA (clamp +-clamp_thr, decoder fed +-1) */
    float   clamp_thr;      /* mapping.clamp */
    /* render + losses */
    float   sc_factor;      /* data.sc_factor */
    float   depth_trunc;    /* cam.depth_trunc */
    float   rgb_missing;    /* training.rgb_missing */
    /* decoder shape (model/decoder.py) */
    int32_t hidden;         /* decoder.hidden_dim == hidden_dim_color (32 or 64) */
    int32_t n_bins;         /* pos.n_bins (16) */
    int32_t geo_feat;       /* decoder.geo_feat_dim (15) */
    int32_t mlp_precision;  /* 0: fp32 SIMT decoder (the accuracy anchor); 1: tcgen05 tensor-core decoder (bf16x3
                             * products, fp32 accumulation; ~2^-16 relative per product) fed by feature planes */
    int32_t _pad;
    int64_t n_rays_total;   /* rays in the whole batch when it is sharded over GPUs (loss means); 0 = n_rays */
    double  bbox[6];        /* float64 bounding box {x0,x1,y0,y1,z0,z1} (model/scene_rep.py:388) */
} rf_ray_cfg;

typedef struct rf_ray_params {
    const float* hash_params;  /* device; HashGrid table */
    const float* gbv_params;   /* device; Dense F=4 grid (the TSDF volume) */
    const float* w_sdf0;       /* device [hidden, in_sdf]   nn.Linear weight, row-major [out,in] */
    const float* w_sdf1;       /* device [1+geo, hidden] */
    const float* w_col0;       /* device [hidden, in_col] */
    const float* w_col1;       /* device [3, hidden] */
} rf_ray_params;

/* Depth sampling along rays (model/scene_rep.py:417-441).  target_d [N]; u [N,S] jitter in [0,1) (required when
 * cfg->perturb, the reference draws it with torch.rand on the CPU generator, :441); z_tables (device,
 * [2*n_range_d + n_samples_d]): linspace(-range_d, range_d, n_range_d), linspace(near, far, n_range_d),
 * linspace(near, far, n_samples_d) exactly as torch.linspace produces them on the host; z_vals [N,S] out. */
int rf_ray_sample_z(const rf_ray_cfg* cfg, const float* target_d, const float* u, const float* z_tables,
                    int64_t n_rays, float* z_vals, void* stream);

/* SDF-to-weight compositing alone (JointEncoding.raw2outputs / sdf2weights, model/scene_rep.py:156-179, :107-127):
 * raw [N,S,4], z_vals [N,S] -> rgb_map [N,3], depth_map [N].  Forward only. */
int rf_ray_composite(const rf_ray_cfg* cfg, const float* raw, const float* z_vals, int64_t n_rays, float* rgb_map,
                     float* depth_map, void* stream);

/* The four mapping losses from the (all-reduced) partial sums of rf_ray_query_forward (model/scene_rep.py:501-517,
 * model/utils.py:190-196, :242-245): losses (device float[4]) = rgb, depth, sdf, fs.  n_rays_total: rays in the whole
 * (multi-GPU) batch; n_samples: samples per ray. */
int rf_ray_loss_finalize(const double* loss_partials, int64_t n_rays_total, int n_samples, float* losses, void* stream);

/* Forward.  rays_o, rays_d [N,3]; z_vals [N,S] (from rf_ray_sample_z);
 * outputs: raw [N,S,4], rgb_map [N,3], depth_map [N];
 * loss_partials (device double[8], may be NULL; caller zeroes it): accumulates the sums the four losses of
 * JointEncoding.mapping are built from:
 *   [0] sum (rgb*w - tgt*w)^2   [1] sum_valid (depth - d)^2   [2] n_valid
 *   [3] sum (s*front - front)^2 [4] sum ((z+s*tr)*m - d*m)^2  [5] n_front (pre-mask) [6] n_sdf (pre-mask)
 * target_d [N] and target_rgb [N,3] are required when loss_partials != NULL. */
int rf_ray_query_forward(const rf_ray_cfg* cfg, const rf_grid_desc* hash, const rf_grid_desc* gbv,
                         const rf_ray_params* p,
                         const float* rays_o, const float* rays_d, const float* target_d,
                         const float* target_rgb, const float* z_vals, int64_t n_rays,
                         float* raw, float* rgb_map, float* depth_map,
                         double* loss_partials, float* workspace, void* stream);

/* Floats of `workspace` the forward needs (and the backward reads back): 0 for mlp_precision 0;
 * (2*n_levels + 4 + 3 + 1) * N * S for mlp_precision 1 (hash feature planes, GBV features, normalised positions,
 * depth along the ray). */
int64_t rf_ray_workspace_floats(const rf_ray_cfg* cfg, const rf_grid_desc* hash, int64_t n_rays);
/* Floats of `scratch` the backward needs.  mlp_precision 0: 4*N*S (7*N*S with ray gradients).  mlp_precision 1:
 * (4 + 2*n_levels)*N*S (d_raw + feature-gradient planes) + the replicated gradient tables of the small levels, and with
 * ray gradients 7*N*S more (GBV-texel gradient and the OneBlob part of the position gradient). */
int64_t rf_ray_scratch_floats(const rf_ray_cfg* cfg, const rf_grid_desc* hash, int64_t n_rays, int ray_grads);

typedef struct rf_ray_grads {
    float* g_hash;     /* device, += ; same shape as hash_params (may be NULL) */
    float* g_w_sdf0;   /* device, += (each may be NULL) */
    float* g_w_sdf1;
    float* g_w_col0;
    float* g_w_col1;
    float* g_rays_o;   /* device [N,3], written; NULL = mapping mode (no ray gradients) */
    float* g_rays_d;   /* device [N,3], written; NULL likewise */
} rf_ray_grads;

/* Backward of the forward above (recomputes activations; mlp_precision 1 re-reads the forward's workspace).  Upstream gradients (each may be NULL = zero):
 *   d_rgb_map [N,3], d_depth_map [N], d_raw [N,S,4], and loss_grads (device float[4]: d/d rgb_loss, depth_loss,
 *   sdf_loss, fs_loss) together with the forward's loss_partials (all-reduced over ranks when sharded).
 * workspace: the forward's (NULL for mlp_precision 0); scratch: device, 16-byte aligned, rf_ray_scratch_floats(). */
int rf_ray_query_backward(const rf_ray_cfg* cfg, const rf_grid_desc* hash, const rf_grid_desc* gbv,
                          const rf_ray_params* p,
                          const float* rays_o, const float* rays_d, const float* target_d,
                          const float* target_rgb, int64_t n_rays,
                          const float* z_vals, const float* raw, const float* rgb_map, const float* depth_map,
                          const float* d_rgb_map, const float* d_depth_map, const float* d_raw,
                          const float* loss_grads, const double* loss_partials,
                          const rf_ray_grads* g, const float* workspace, float* scratch, void* stream);

/* Point query (model/scene_rep.py:212-310 — query_sdf_res / query_color_residual / run_network(flat)):
 * x [n,3] normalised coords -> raw [n,4] (rgb, sdf).  variant: 0 = query_color_sdf (cfg->clamp_mode),
 * 1 = query_sdf_res (always +-1 clamp; rgb lanes hold geo garbage-free zeros), 2 = query_color_residual
 * (decoder fed the unscaled GBV tsdf). */
int rf_point_query_forward(const rf_ray_cfg* cfg, const rf_grid_desc* hash, const rf_grid_desc* gbv,
                           const rf_ray_params* p, const float* x, int64_t n, int variant,
                           float* raw, float* workspace, void* stream);
/* Floats of `workspace` for the call above: 0 for mlp_precision 0, (2*n_levels + 4 + 3) * n for mlp_precision 1
 * (the same feature planes as the ray query, one sample per "ray"). */
int64_t rf_point_workspace_floats(const rf_ray_cfg* cfg, const rf_grid_desc* hash, int64_t n);

/* Micro-benchmarks for the gather-bound roofline denominators: random 8-byte loads and
 * random fp32 red.add over a table of table_bytes.  Synchronous: runs `iters` launches after one warm-up and
 * returns the mean ms per launch in *ms (host).  The number of operations actually issued per launch is n_ops
 * rounded up to a multiple of (8 * SMs * 256 * 8) for gathers, (8 * SMs * 256) for atomics. */
int rf_microbench_gather(void* table, int64_t table_bytes, int64_t n_ops, int iters, float* ms, void* stream);
int rf_microbench_atomic(void* table, int64_t table_bytes, int64_t n_ops, int iters, float* ms, void* stream);

/* ------------------------------------------------------------------------------------------------
 * N1 — the optimiser step after each backward: torch.optim.Adam as mp_slam/slam.py:271-286 sets it up
 * (betas 0.9/0.99; decoder weight_decay 1e-6; table eps 1e-15), dense torch semantics, one pass; zero_grad != 0 also
 * clears the gradient (mp_slam/mapper.py:423).  All arrays: device fp32 [n], 16-byte aligned.  step = 1 for the first call;
 * step_dev (device float, may be NULL) overrides it with a step count read by the kernel, so that the call can be
 * captured in a CUDA graph and replayed (torch's `capturable` Adam keeps the same device-side float step).
 * ---------------------------------------------------------------------------------------------- */
int rf_adam_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                 double lr, double beta1, double beta2, double eps, double weight_decay,
                 int64_t step, const float* step_dev, int zero_grad, void* stream);

/* Optional per-kernel timing for bench.py's live roofline numbers.  rf_profile_enable(1) makes the launchers
 * bracket each hot kernel with CUDA events on the launching stream (the last launch of each slot is kept);
 * rf_profile_read(ms) synchronises on those events and writes RF_PROF_SLOTS floats (ms, -1 = not launched). */
#define RF_PROF_SLOTS 64
enum { RF_PROF_TSDF_LOCAL = 0, RF_PROF_TSDF_GLOBAL = 1, RF_PROF_RAY_Z = 2, RF_PROF_RAY_POS = 3, RF_PROF_ENCODE = 4,
       RF_PROF_MLP_FWD = 5, RF_PROF_COMPOSITE_FWD = 6, RF_PROF_COMPOSITE_BWD = 7, RF_PROF_MLP_BWD = 8,
       RF_PROF_SCATTER = 9, RF_PROF_SAMPLE_FWD = 10, RF_PROF_SAMPLE_BWD = 11, RF_PROF_RAY_GRAD = 12, RF_PROF_TSDF_RECENTER = 13, RF_PROF_TRACK_FITNESS = 14,
       RF_PROF_SCATTER_LEVEL0 = 16 /* +level, only with RF_DEBUG_PER_LEVEL=1 */, RF_PROF_ENCODE_LEVEL0 = 40 /* +level */ };
int rf_profile_enable(int on);
int rf_profile_read(float* ms);

/* Hardware self-test of the hand-written tcgen05 building blocks (csrc/umma.cuh): bf16x3 GEMM on one CTA.
 * mode 0: D[128,N] = A[128,K] * B[N,K]^T; mode 1: D[f,j] = sum_m A[m,f] * B[m,j] (A [128,K], B [128,N]).
 * Synchronous.  Returns 0, or 1001 if the MMA never completed. */
int rf_umma_selftest(const float* A, const float* B, float* D, int K, int N, int mode, void* stream);

/* ---- N4 (second half): marching cubes on the device (csrc/marching_cubes.cu) ---------------------------------------------
 * Dual-grid marching cubes of thirdparty/NumpyMarchingCubes (`mcubes.marching_cubes(volume, isovalue, truncation)`,
 * utils.py:169).  volume [X][Y][Z] fp32 (z fastest, the layout query_lattice produces); voxels with |d| >= truncation (or
 * NaN / -inf) are invalid and no surface is built next to them.
 *   rf_mc_count : corner_ws (rf_mc_corner_floats floats) <- dual-grid corner values; counts [X*Y*Z] <- triangles per cell.
 *   rf_mc_emit  : offsets [X*Y*Z] = exclusive prefix sum of counts (int64); triangles [T][3][3] positions in voxel units in
 *                 the reference's cell order; keys [T][3] int64 = identity of each vertex (dual edge or snapped corner) for welding. */
int64_t rf_mc_corner_floats(int X, int Y, int Z);
int rf_mc_count(const float* volume, int X, int Y, int Z, float isovalue, float truncation, float* corner_ws, int* counts, void* stream);
int rf_mc_emit(const float* corner_ws, int X, int Y, int Z, float isovalue, const long long* offsets, float* triangles,
               long long* keys, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RF_ABI_H */
