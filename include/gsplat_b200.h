/*
 * gsplat_b200.h -- C ABI of libgsplat_b200.so: the B200 (sm_100a) implementation of the
 * GaussianRenderer.render hot path of Loveof1ife7/mini-3d-gaussian-splatting.
 *
 * The reference has no FFI (it is pure Python/torch); each entry point below replaces one
 * reference *method* and is what a binding for that method would call.  Citations are
 * file:line into the reference tree.
 *
 *   gs_project_fwd / gs_project_bwd   GaussianRenderer._project_gaussians_3d_to_2d  src/core/renderer.py:117-200
 *                                     + GaussianModel.compute_3d_covariance         src/core/gaussian_model.py:200-207
 *                                     + MathUtils.build_rotation_matrix             src/utils/math_utils.py:9-26
 *                                     + activations get_opacity / sigmoid(features) src/core/gaussian_model.py:120-122, renderer.py:88-94
 *                                     + GaussianRenderer._frustum_culling           src/core/renderer.py:201-220
 *   gs_bin_prepare / gs_bin_sort      GaussianRenderer._sort_gaussians_by_depth     src/core/renderer.py:222-239
 *                                     + the tile-list building loop                 src/core/renderer.py:263-298
 *   gs_raster_fwd / gs_raster_bwd     the pixel loop + epilogue of _tile_rasterization  src/core/renderer.py:300-367
 *                                     (backward = what torch autograd derives from it)
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - the library never allocates or frees device memory: outputs and scratch are caller-owned
 *     (query gs_bin_workspace_bytes, then run);
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); nothing synchronises;
 *   - return value: 0 = ok, negative = GsStatus; text via gs_last_error_string() (thread-local);
 *   - fp32 everywhere; row-major contiguous arrays unless a stride is given;
 *   - re-entrant: no global mutable state.
 */
#ifndef GSPLAT_B200_H_
#define GSPLAT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GS_ABI_VERSION 18

typedef enum GsStatus {
    GS_OK = 0,
    GS_ERR_INVALID_ARGUMENT = -1,
    GS_ERR_CUDA = -2,
    GS_ERR_WORKSPACE_TOO_SMALL = -3,
    GS_ERR_UNSUPPORTED = -4
} GsStatus;

/* Camera block, HOST memory, 20 floats:
 *   [0..8]  Rv  row-major 3x3  = world_view_transform()[:3,:3]      renderer.py:150-151
 *   [9..11] Tv                 = world_view_transform()[:3,3]       renderer.py:152
 *   [12] fx = 0.5*W/tan(FoVx/2)  [13] fy = 0.5*H/tan(FoVy/2)  [14] cx = W/2  [15] cy = H/2   renderer.py:142-147
 *   [16..18] camera centre in world space, -Rv^T Tv (only read when sh_degree > 0)   [19] unused
 * (computed in double on the host and rounded once to fp32, as the reference does). */
#define GS_CAMERA_FLOATS 20

/* Per-splat record consumed by the raster kernels: 12 floats = 3 x float4, 48-byte stride.
 *   {mx, my, c*Q00, c*(Q01+Q10)} {c*Q11, opacity, depth, r} {g, b, regular, log2(opacity)},  c = -0.5*log2(e):
 * the splat weight exp(-0.5*s) of renderer.py:333-334 is then one exp2 of the quadratic form, and for
 * `regular` records opacity * weight is one exp2 of (quadratic form + log2 opacity).
 * `regular` (1.0 / 0.0) marks splats with opacity in [0,1] and a positive-definite, well-conditioned
 * conic, for which the two clamps of renderer.py:335,339 are identities: batches of such entries take
 * a shorter instruction sequence in gs_raster_fwd / gs_raster_bwd (same results; a record with
 * regular = 0 is always handled with the literal clamp sequence). */
#define GS_SPLAT_REC_FLOATS 12

int gs_abi_version(void);
const char* gs_last_error_string(void);
/* Compute capability the kernels were built for (100 for sm_100a). */
int gs_built_for_sm(void);
/* Cumulative number of hand-written kernels this process has launched through the entry points
 * below (CUB's internal sort/scan kernels are not counted).  Statistics only. */
int64_t gs_kernel_launch_count(void);

/* ---------------------------------------------------------------------------------------
 * Stage P+M+C: projection, fused with activations, 3-D covariance and frustum culling.
 *
 * Two input modes, mirroring the two ways the reference renderer is fed:
 *   parameter mode  (scaling_log != NULL && rotation != NULL): raw GaussianModel parameters;
 *        Sigma = R(normalize(rotation)) diag(exp(scaling_log)^2) R^T is formed in registers.
 *   covariance mode (cov3d != NULL): a ready [n,3,3] covariance, as the duck-typed
 *        `gaussians.get_covariance` of the reference's own tests (tests/test_renderer.py:48-53).
 * opacity: [n]; a logit when opacity_is_logit != 0 (sigmoid is fused), else already activated.
 * feat0:   pointer to features[0,0,0]; row i, channel c is feat0[i*feat_stride + c]; the colour
 *          is sigmoid(features[:,0,:]) (renderer.py:88-92).
 * sh_rest / sh_rest_stride / sh_degree: optional view-dependent colour (an extension: the reference
 *          is DC-only).  sh_degree 0 (default) ignores sh_rest.  For degree d in 1..3 the colour is
 *          sigmoid(features[:,0,:] + sum_{k=1}^{(d+1)^2-1} Y_k(dir) features[:,k,:]) with the real SH
 *          basis and dir = normalize(xyz - camera centre); row k >= 1, channel c of splat i is
 *          sh_rest[i*sh_rest_stride + (k-1)*3 + c].  Identical to degree 0 when those rows are zero.
 *
 * Outputs (all length n, invisible splats included -- renderer.py:106-114):
 *   means2d [n,2]  depths [n]  conics [n,2,2]  radii [n] (float)  colors [n,3]  opacities [n]
 *   vis [n] (0/1 bytes = visibility_filter)
 *   tiles_touched [n] int32: tiles of the splat's integer AABB (0 if culled or empty)   renderer.py:278-293
 *   tile_rect [n,4] uint16: tx0, ty0, tx1, ty1 inclusive (valid when tiles_touched > 0)
 *   depth_keys [n] uint32: fp32 bits of depth for splats with tiles_touched > 0; 0xFFFFFFFE for visible
 *                          splats whose AABB is empty; 0xFFFFFFFF for culled splats
 *   splat_rec [n,12]: packed copy for the raster kernels (GS_SPLAT_REC_FLOATS)
 * ------------------------------------------------------------------------------------- */
int gs_project_fwd(int64_t n,
                   const float* xyz,
                   const float* scaling_log, const float* rotation,
                   const float* cov3d,
                   const float* opacity, int32_t opacity_is_logit,
                   const float* feat0, int64_t feat_stride,
                   const float* sh_rest, int64_t sh_rest_stride, int32_t sh_degree,
                   const float* camera_host,
                   int32_t img_w, int32_t img_h, int32_t tile_size,
                   float radius_min, float radius_max,
                   float* means2d, float* depths, float* conics, float* radii,
                   float* colors, float* opacities, uint8_t* vis,
                   int32_t* tiles_touched, uint16_t* tile_rect, uint32_t* depth_keys,
                   float* splat_rec,
                   void* stream);

/* Backward of gs_project_fwd.  Upstream gradients: g_means2d [n,2], g_conics [n,2,2],
 * g_depths [n], g_colors [n,3], g_opacities [n].  Results are WRITTEN (not accumulated):
 *   g_xyz [n,3]; parameter mode: g_scaling_log [n,3], g_rotation [n,4]; covariance mode: g_cov3d [n,3,3];
 *   g_opacity [n] (w.r.t. the logit when opacity_is_logit, else w.r.t. the activated value);
 *   g_feat0: row i channel c at g_feat0[i*g_feat_stride + c]; with sh_degree 0 the other feature
 *            rows are the caller's to zero (the reference yields zeros there -- SURVEY 3.2);
 *   g_sh_rest (sh_degree > 0): all 15 higher-order rows are written (rows beyond the degree as 0).
 * accumulate != 0: every result above is ADDED to what its buffer holds instead of written -- the
 *   multi-view step accumulates the views' parameter gradients straight into one flat buffer
 *   (replaces the `.grad +=` that autograd would otherwise run per tensor).
 * stat_* (all NULL, or all non-NULL, each [n]): densification statistics fused into this pass; for
 *   every splat with stat_vis[i] != 0:  stat_grad_norm[i] += |g_means2d[i]|, stat_count[i] += 1,
 *   stat_max_radii[i] = max(stat_max_radii[i], stat_radii[i]) when accumulate != 0; with accumulate == 0 the
 *   three statistics are WRITTEN for every splat (zeros where not visible), like the gradients -- so the
 *   first view of a step initialises the whole buffer and nothing has to be zeroed.  These are the buffers the reference
 *   allocates as xyz_gradient_accum / denom / max_radii2D (gaussian_model.py:29-31). */
int gs_project_bwd(int64_t n,
                   const float* xyz,
                   const float* scaling_log, const float* rotation,
                   const float* cov3d,
                   const float* opacity, int32_t opacity_is_logit,
                   const float* feat0, int64_t feat_stride,
                   const float* sh_rest, int64_t sh_rest_stride, int32_t sh_degree,
                   const float* camera_host,
                   const float* g_means2d, const float* g_conics, const float* g_depths,
                   const float* g_colors, const float* g_opacities,
                   float* g_xyz, float* g_scaling_log, float* g_rotation, float* g_cov3d,
                   float* g_opacity, float* g_feat0, int64_t g_feat_stride,
                   float* g_sh_rest, int64_t g_sh_rest_stride,
                   int32_t accumulate,
                   const float* stat_radii, const uint8_t* stat_vis,
                   float* stat_grad_norm, float* stat_count, float* stat_max_radii,
                   void* stream);

/* ---------------------------------------------------------------------------------------
 * Stage S+B: global depth order and per-tile lists.
 *
 * The reference sorts visible splats by depth once and appends each, in that order, to the
 * list of every tile its AABB touches.  That is the sequence obtained by a stable sort of
 * (tile_id << 32 | depth_bits) keys emitted in ascending splat index; it is produced here as
 *   gs_bin_prepare: stable radix sort of depth_keys (ties -> ascending index), gather of
 *                   tiles_touched in that order, exclusive prefix sum;
 *                   counters[0] = splats with >= 1 tile, counters[1] = D (tile pairs),
 *                   counters[2] = splats that passed culling (3 x int64)
 *   -- caller reads the counters back (the one host sync of the frame) --
 *   gs_bin_sort:    groups the pairs by tile, keeping depth order.  algo GS_BIN_COUNTING (default via
 *                   GS_BIN_AUTO): hand-written stable chunked counting sort (per-chunk shared-memory
 *                   tile counters walked in depth order, column scan, parallel scatter);
 *                   GS_BIN_RADIX: duplication + library (CUB) radix sort on the tile id + range
 *                   extraction -- kept for tile grids too large for the counters and as a cross-check;
 *                   GS_BIN_BLOCKED: two-level counting sort -- ranks grouped by 8x8-tile block first (one
 *                   radix pass on the block id), then each piece of a block's list is sorted by tile in
 *                   shared memory and written out in coalesced runs.  Requires every tile rectangle to span
 *                   at most 8 tiles per side (radius_max <= 50 px at 16-px tiles); the caller must choose
 *                   another algorithm otherwise.
 * counters_dev (optional, counting sort only): the device address of gs_bin_prepare's counters.  When given,
 *          `num_sorted` and `d` are CAPACITIES (pass n and the size of entry_ids): every kernel reads the
 *          actual sizes on the device, and if D does not fit the capacity nothing is written and
 *          tile_ranges stays zero -- the caller compares the counters with its capacity once they have
 *          arrived on the host and, on overflow, repeats the call with exact sizes.  This lets the whole
 *          forward be enqueued without waiting for the read-back (no bubble on the GPU).
 * Outputs: entry_ids [D] int32 splat ids grouped by tile, front to back;
 *          tile_ranges [num_tiles,2] int32 = [begin,end) into entry_ids;
 *          entry_keys [D] uint64 (optional, may be NULL) = tile_id<<32 | depth_bits of each entry,
 *          for parity checks against the reference order;
 *          tile_order [num_tiles] int32 (optional, flat counting sort only; ignored by the other algorithms): the tiles
 *          sorted by decreasing list length (buckets of 8 entries) -- what gs_tile_order(tile_ranges) would give, produced
 *          by the same launch that scans the tiles;
 *          flag_count (optional, one int32): set to zero -- the counter of tiles whose stored prefix is too short, which the
 *          compositing pass enqueued behind this call raises (truncated lists, see gs_bin_complete).
 * ------------------------------------------------------------------------------------- */
#define GS_BIN_AUTO 0
#define GS_BIN_COUNTING 1
#define GS_BIN_RADIX 2
#define GS_BIN_BLOCKED 3

int64_t gs_bin_workspace_bytes(int64_t n, int64_t d_capacity, int32_t num_tiles);

int gs_bin_prepare(int64_t n,
                   const uint32_t* depth_keys, const int32_t* tiles_touched,
                   void* workspace, int64_t workspace_bytes,
                   int32_t* sorted_ids, int64_t* offsets, int64_t* counters,
                   void* stream);

int gs_bin_sort(int64_t n, int64_t num_sorted, int64_t d,
                const int32_t* sorted_ids, const int64_t* offsets,
                const uint16_t* tile_rect, const uint32_t* depth_keys,
                int32_t tiles_x, int32_t num_tiles, int32_t algo,
                void* workspace, int64_t workspace_bytes,
                int32_t* entry_ids, int32_t* tile_ranges, uint64_t* entry_keys,
                const int64_t* counters_dev, int32_t list_cap,
                int32_t* tile_order, int32_t* flag_count, void* stream);

/* Truncated tile lists (flat counting sort only).  With list_cap > 0 gs_bin_sort stores only the first list_cap
 * entries of every tile's list (tile_ranges still describe the complete lists); gs_raster_fwd, given the same
 * list_cap, composites from that prefix and flags a tile that reaches the end of its stored prefix with pixels
 * still alive (tile_flags[t] = 1, *flag_count += 1).  gs_bin_complete then stores the remaining entries of the
 * flagged tiles -- it needs gs_bin_sort's workspace untouched -- and gs_raster_fwd with rerun != 0 composites the
 * flagged tiles again from their complete lists.  Both completion calls are meant to be enqueued unconditionally:
 * every CTA returns at once when *flag_count == 0.  The result is identical to list_cap = 0; what is saved are
 * the scattered stores of entries no pixel consumes (on config[1]: > 70 % of the 26 M). */
int gs_bin_complete(int64_t n, int64_t num_sorted, int64_t d,
                    const int32_t* sorted_ids, const int64_t* offsets, const uint16_t* tile_rect,
                    int32_t tiles_x, int32_t num_tiles,
                    const void* workspace, int64_t workspace_bytes,
                    int32_t list_cap, const uint8_t* tile_flags, const int32_t* flag_count,
                    int32_t* entry_ids, const int64_t* counters_dev,
                    void* stream);

/* ---------------------------------------------------------------------------------------
 * Stage R: 16x16-tile front-to-back compositing.
 *
 * bg: 3 floats (device).  `any_visible_host` (= counters[2] > 0) says whether any splat passed
 * culling: the reference returns the background ONCE, unclamped, when nothing is visible
 * (renderer.py:74-83) and adds it TWICE otherwise (renderer.py:273 + :359); if counters_dev is not NULL
 * the kernel reads that fact from counters_dev[2] itself and any_visible_host is ignored.
 * tile_order (optional): a permutation of the tile indices giving the order in which CTAs take tiles
 * (gs_tile_order on an estimate of the tiles' work -- e.g. the previous frame's tile_consumed -- puts the
 * heavy tiles first so that light ones fill the last wave); results do not depend on it.
 * list_cap / tile_flags / flag_count / rerun: truncated lists, see gs_bin_complete (list_cap <= 0: complete lists,
 * the three may be NULL / 0).
 * Outputs: image [3,H,W], alpha [1,H,W], depth [1,H,W];
 *   saved for backward: pix_state [H*W,4] = {C_r, C_g, C_b, Dsum} before the epilogue;
 *   tile_consumed [num_tiles] int32 = list entries the tile loaded before all its pixels saturated
 *   (granularity: 8 entries); n_consumed [H*W] int32 (optional, may be NULL) = entries each
 *   pixel walked up to and including the one that terminated it -- a debug/parity output that
 *   selects a kernel variant with one extra predicated move per evaluation.
 * ------------------------------------------------------------------------------------- */
int gs_raster_fwd(int32_t img_w, int32_t img_h, int32_t tile_size,
                  const int32_t* entry_ids, const int32_t* tile_ranges,
                  const float* splat_rec, const float* bg, int32_t any_visible_host,
                  const int64_t* counters_dev, const int32_t* tile_order,
                  int32_t list_cap, uint8_t* tile_flags, int32_t* flag_count, int32_t rerun,
                  float* image, float* alpha, float* depth,
                  float* pix_state, int32_t* n_consumed, int32_t* tile_consumed,
                  void* stream);

/* tile_order = a permutation of [0, num_tiles) sorted by decreasing work estimate (bucketed by 8 entries):
 * tile_consumed when given (what a forward measured), else the tiles' list lengths from tile_ranges. */
int gs_tile_order(int32_t num_tiles, const int32_t* tile_consumed, const int32_t* tile_ranges, int32_t* tile_order,
                  void* stream);

/* Backward of gs_raster_fwd.  g_image [3,H,W], g_alpha [1,H,W], g_depth [1,H,W] upstream; g_alpha and g_depth may be
 * NULL when no gradient flows into that output (the reference's train step differentiates the image only).
 * Gradients are ACCUMULATED (atomic adds) into caller-zeroed g_means2d [n,2], g_conics [n,2,2],
 * g_depths [n], g_colors [n,3], g_opacities [n].
 * tile_order_scratch (optional, [num_tiles] int32): when given, the tiles are first bucketed by their exact
 * work (tile_consumed, heaviest first) and taken in that order -- the grid is only a few waves of one-warp
 * CTAs and the tiles' work varies widely, so the natural order leaves full-size tiles in the last wave.
 * tile_order_ready != 0: tile_order_scratch already holds that order (gs_tile_order(tile_consumed), e.g. computed on
 * another stream while the loss was evaluated) and is only read. */
int gs_raster_bwd(int32_t img_w, int32_t img_h, int32_t tile_size,
                  const int32_t* entry_ids, const int32_t* tile_ranges,
                  const float* splat_rec, const float* bg,
                  const float* alpha, const float* pix_state,
                  const int32_t* tile_consumed, int32_t* tile_order_scratch, int32_t tile_order_ready,
                  const float* g_image, const float* g_alpha, const float* g_depth,
                  float* g_means2d, float* g_conics, float* g_depths,
                  float* g_colors, float* g_opacities,
                  void* stream);

/* ---------------------------------------------------------------------------------------
 * Multi-GPU exchange of the view-sharded step (no reference counterpart: the reference is single
 * device).  Every rank holds the same flat fp32 buffer in peer-mapped (symmetric) memory;
 * peer_ptrs_host[r] is rank r's buffer as seen from THIS process (HOST array of `world` device
 * addresses).  One kernel: this rank reduces its 1/world slice of two regions across all peers --
 * elementwise SUM over [sum_offset, sum_offset+sum_count) and MAX over [max_offset, max_offset+max_count)
 * (floats, multiples of 4) -- and writes the result into every peer's buffer.  The caller must make sure
 * (device-side barrier before) that all peers have finished producing their buffers and (barrier after)
 * that all ranks' results have landed before anyone reads.  Results are bit-identical on all ranks.
 * multicast_ptr: 0, or the NVSwitch multicast address of the same buffers; then the reduction runs inside
 * the switch (multimem.ld_reduce / multimem.st) and the MAX region must hold non-negative floats.
 * flags: GS_PEER_TMA moves the data with bulk asynchronous copies (cp.async.bulk global<->shared, mbarrier completion,
 * a 3-stage ring of 64 KB per CTA) instead of per-thread 16-byte loads/stores; same owner, same summation order, same bits.
 * ------------------------------------------------------------------------------------- */
#define GS_PEER_TMA 1
int gs_peer_allreduce(const uint64_t* peer_ptrs_host, uint64_t multicast_ptr, int32_t world, int32_t rank,
                      int64_t sum_offset, int64_t sum_count, int64_t max_offset, int64_t max_count,
                      int32_t flags, void* stream);

/* ---------------------------------------------------------------------------------------
 * Density control on the device: GaussianModel.density_and_split / density_and_clone / prune_points
 * (src/core/gaussian_model.py:130-197) as driven by DensityController.densify_and_prune
 * (src/core/optimizer.py:43-71), as one plan pass + one apply pass.
 *   clone: |grad[i]| > grad_threshold and mean(exp(scaling_log[i])) < small_sigma -> keep + jittered copy
 *          (xyz + noise[i] * 0.5*mean sigma);
 *   split: |grad[i]| > grad_threshold and mean sigma > large_sigma -> replaced by two children at
 *          xyz -+ R[:,0]*0.5*mean sigma, sigma*0.75, opacity logit clamped to [-6,6];
 *   prune: rows with sigmoid(opacity) <= min_opacity are dropped.
 * gs_densify_plan classifies and scans; counts (device, 5 x int64) = {surviving originals, clone copies,
 * split parents, total rows, clone candidates}.  The caller reads the counts, allocates the six output arrays with
 * `total` rows and calls gs_densify_apply with the same workspace.  Output order: surviving originals, clone
 * copies, "minus" children, "plus" children -- each in index order (what the sequential formulation yields).
 * noise [clone candidates, 3]: row j is the jitter of the j-th splat (in index order) that met the clone
 * criterion -- whether or not its copy survives the opacity test -- i.e. the randn(k, 3) block that
 * gaussian_model.py:171 draws, so the same generator yields the same clones as the sequential formulation.
 * src_row (optional, [total] int32): for every output row the index of the original it continues (surviving originals),
 * or -1 for rows created by this round (clone copies, split children) -- what an optimiser needs to carry its
 * per-row state across the round.
 * ------------------------------------------------------------------------------------- */
int64_t gs_densify_workspace_bytes(int64_t n);

int gs_densify_plan(int64_t n, const float* scaling_log, const float* opacity, const float* grad,
                    float grad_threshold, float small_sigma, float large_sigma, float min_opacity,
                    void* workspace, int64_t workspace_bytes, int64_t* counts, void* stream);

int gs_densify_apply(int64_t n, const void* workspace, int64_t kept, int64_t cloned, int64_t split,
                     const float* xyz, const float* features_dc, const float* features_rest,
                     const float* scaling_log, const float* rotation, const float* opacity, const float* noise,
                     float* o_xyz, float* o_features_dc, float* o_features_rest, float* o_scaling_log,
                     float* o_rotation, float* o_opacity, int32_t* src_row, void* stream);

/* ---------------------------------------------------------------------------------------
 * Loss heads fused into one pass over the rendered planes (the caller of the render path: the reference's
 * train step computes l1_loss(image, target) = |a - b|.mean() with torch ops, src/utils/loss.py, optimizer.py:137-139).
 *
 * Both entry points reduce deterministically (fixed slices, fixed order) into `out` (device, one float) and need a
 * workspace of gs_loss_workspace_bytes() that the caller zero-fills ONCE (the kernels leave it zeroed again).
 *
 * gs_weighted_sum: out = sum_k coeff[k] * dot(x[k], w[k]) over num_terms <= 4 pairs of device arrays of n[k] floats
 *   (x, w, n, coeff are HOST arrays of num_terms entries).  Its gradient w.r.t. x[k] is coeff[k] * w[k]: no kernel.
 * gs_l1_loss: out = mean|x - target| over n floats; when `grad` is given it receives
 *   d(grad_scale * out)/dx = grad_scale * sign(x - target) / n in the same pass (sign(0) = 0, as torch).
 * ------------------------------------------------------------------------------------- */
int64_t gs_loss_workspace_bytes(void);

int gs_weighted_sum(int32_t num_terms, const float* const* x, const float* const* w, const int64_t* n,
                    const float* coeff, float* out, void* workspace, int64_t workspace_bytes, void* stream);

int gs_l1_loss(const float* x, const float* target, int64_t n, float grad_scale, float* grad, float* out,
               void* workspace, int64_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GSPLAT_B200_H_ */
