// Stage R: 16x16-tile front-to-back alpha compositing, forward and backward.
//
// Reference semantics: the pixel loop and epilogue of GaussianRenderer._tile_rasterization,
// src/core/renderer.py:300-367.
//
// Both kernels are bound by instruction issue, not by HBM (ncu: DRAM < 7 % of peak; a tile saturates after a few
// hundred of its thousands of list entries, so ~0.3 GB moves while ~6e8 pixel x splat evaluations execute).  The
// roof is a count of issue slots -- a packed FP32 instruction costs two, everything else one -- measured in
// profiles/r2_issue_model.md (tools/ubench_issue.cu): 122.5 slots per list entry forward, 307 backward, the kernels
// at 0.88 / 0.84 of it.  The design therefore minimises issue slots per evaluation:
//   * one warp per tile, each lane owns a 1x8 pixel strip (16 rows x 2 strips) kept as 4 packed fp32 pairs
//     (FFMA2 / FMUL2 / FADD2): the per-entry work (shared-memory broadcast of the 48-byte record, the dy
//     terms, loop control) is amortised over 8 evaluations, 4 independent dependency chains hide FP32/MUFU
//     latency, and no block-wide barrier exists -- only __syncwarp;
//   * branch-free evaluation: the splat weight is one MUFU.EX2 on a quadratic form whose coefficients were
//     pre-multiplied by -0.5*log2(e) when project_fwd packed the record; per batch of 32 entries a fast
//     path drops both clamps of the reference (identities for records flagged `regular`) and folds the
//     two skip rules into the weight itself (see eval_pair);
//   * backward: each lane pre-sums its 8 pixels, then 10 plain per-lane sums are reduced across the warp
//     through a padded shared-memory transpose (10 STS + 3 LDS.128 per lane instead of 50 shuffles), the
//     entry's conic / opacity factors are applied to the reduced sums by the lane that owns each output, and
//     the eleven gradient values leave as ONE reduction instruction (lanes 0..10 -> 11 addresses);
//   * tiles are taken longest-first (tile_order_kernel): the grids are ~3 waves of one-warp CTAs and the
//     work per tile varies by 40 %, so raster order would leave full-size tiles in the last wave.
//
// Accuracy of the weight: ex2.approx on the pre-scaled form differs from the reference's
// exp(-0.5*s) by <= ~2e-7 relative -- the same size as the error of an "accurate" expf, whose own
// last step is the same MUFU.EX2.  The accumulated opacity A, which decides termination, keeps the
// reference's recurrence A += (1-A)*a with separately rounded operations; forward and backward
// share one inline evaluation function, so they walk bit-identical contributor sets.
#include "common.cuh"

namespace gs {

#ifndef GS_FWD_MINB
#define GS_FWD_MINB 18                 // resident one-warp CTAs per SM asked of ptxas for the forward kernel: 96 registers instead of 111, no spills (325 -> 321 us)
#endif
#ifndef GS_BWD_MINB
#define GS_BWD_MINB 16
#endif
#ifndef GS_RASTER_WARPS
#define GS_RASTER_WARPS 1              // tiles (= warps) per CTA; consecutive slots of the launch order, i.e. tiles of similar work
#endif
#ifndef GS_BWD_RAWSUMS
#define GS_BWD_RAWSUMS 1               // backward: reduce the raw per-lane sums across the warp, apply the per-entry coefficients after
                                       // (0: every lane forms the five conic / mean terms itself before the reduction; 838 vs 817 us)
#endif
#ifndef GS_BWD_GROUP
#define GS_BWD_GROUP 2                 // list entries whose arithmetic runs between two warp barriers (backward)
#endif

constexpr int kWarpsPerCta = GS_RASTER_WARPS;
constexpr int kPx = 8;                 // pixels per lane (1x8 strip) = 4 packed pairs
constexpr int kPairs = kPx / 2;
constexpr int kBatch = 32;             // list entries staged per round (one per lane)
constexpr float kTermA = 0.995f;       // renderer.py:352
constexpr float kMinW = 1e-5f;         // renderer.py:336
constexpr float kLn2 = 0.69314718056f;
constexpr float kNegHalfLog2e = -0.72134752044448170f;
// log2 of the fp32 value of kMinW: on the fast path the opacity is folded into the exponent (a = 2^(s' + log2 opacity)),
// so the reference's `w < 1e-5` rule is tested on the exponent: s' + log2 opacity >= log2(1e-5) + log2 opacity.
constexpr float kLog2MinW = -16.609640510882354f;
// Opacities at or below this are skipped as a whole (warp-uniform).  The reference skips a <= 0
// (renderer.py:340); for 0 < opacity <= 1e-30 it would add < 1e-30 to every accumulator.
constexpr float kTinyOpacity = 1e-30f;

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Blackwell packed fp32: one instruction, two IEEE-rounded fp32 results (FFMA2 / FADD2 / FMUL2).  Each lane keeps its
// 8 pixels as 4 (even, odd) pairs.  Measured (profiles/r2_issue_model.md): a packed instruction holds its scheduler for
// two cycles, so packing buys no FP32 throughput; what it buys is half the instruction count around the arithmetic
// (operand moves, predicates stay per pair) and the registers of one pair per two pixels -- scalar pairs measured
// slower (profiles/r2_experiments.md: 367 -> 389 us forward with two of the four pairs scalar).
__device__ __forceinline__ float2 bc2(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
// The same three operations with a compile-time choice between the packed instruction and two scalar ones (identical
// IEEE results).  A packed op holds the FP32 pipe for two cycles, so the cycle after it must be filled by a non-FP32
// instruction; with as many packed ops as other instructions the schedule would have to alternate perfectly.  Running
// the last GS_*_SCALAR_PAIRS of a lane's four pixel pairs through scalar ops trades issue slots for slack.
template <bool kPk> __device__ __forceinline__ float2 fma2s(float2 a, float2 b, float2 c) {
    if (kPk) return __ffma2_rn(a, b, c);
    return make_float2(__fmaf_rn(a.x, b.x, c.x), __fmaf_rn(a.y, b.y, c.y));
}
template <bool kPk> __device__ __forceinline__ float2 add2s(float2 a, float2 b) {
    if (kPk) return __fadd2_rn(a, b);
    return make_float2(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y));
}
template <bool kPk> __device__ __forceinline__ float2 mul2s(float2 a, float2 b) {
    if (kPk) return __fmul2_rn(a, b);
    return make_float2(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y));
}
#ifndef GS_FWD_SCALAR_PAIRS
#define GS_FWD_SCALAR_PAIRS 0
#endif
#ifndef GS_BWD_SCALAR_PAIRS
#define GS_BWD_SCALAR_PAIRS 0
#endif

// Per-entry values shared by a lane's 8 pixels, pre-broadcast into pairs.
struct EntryRow {
    float2 neg_mx, q00, qsdy, q11dy2;      // fast path: q11dy2 also carries log2(opacity)
    float op;                              // general path: opacity; fast path: log2(kMinW) + log2(opacity)
};

// renderer.py:330-346 for one pixel pair.  `A >= kTermA` encodes "this pixel has terminated (or
// lies outside the image)", so no separate flag is carried.  Forward and backward both call this, so
// they agree bit for bit.
//
// Two variants, chosen per batch of 32 list entries (warp-uniform):
//  * kFast -- every entry of the batch carries the "regular" flag project_fwd stored in its record:
//    opacity in [0,1] and a positive-definite, well-conditioned conic.  Then exp(-s/2) <= 1 (up to one
//    rounding of s) and opacity*w <= 1, so both clamps of the reference are identities and are dropped,
//    and the two skip rules (pixel terminated, w < 1e-5) are folded into the weight itself:
//    wm = live ? e : 0 makes a, the contribution and every gradient term of that pixel an exact zero
//    without further selects.
//  * general -- the literal clamp / skip sequence, for anything else (opacities outside [0,1] through the
//    covariance-mode inputs, degenerate conics).
struct PairEval {
    float2 dx, e, w, u, a, T, contrib;
    bool act0, act1;
};

// wm = (A < kTermA && e >= kMinW) ? e : 0  -- two chained compares and one select
__device__ __forceinline__ float live_weight(float A, float e) {
    float wm;
    asm("{\n\t.reg .pred p;\n\tsetp.lt.f32 p, %1, %2;\n\tsetp.ge.and.f32 p, %3, %4, p;\n\tselp.f32 %0, %3, 0f00000000, p;\n\t}"
        : "=f"(wm) : "f"(A), "f"(kTermA), "f"(e), "f"(kMinW));
    return wm;
}
// fast path: am = (A < kTermA && sp >= thr) ? a : 0 with a = 2^sp = opacity * w and thr = log2(kMinW) + log2(opacity);
// the compares do not wait for the MUFU result
__device__ __forceinline__ float live_alpha(float A, float sp, float thr, float a) {
    float am;
    asm("{\n\t.reg .pred p;\n\tsetp.lt.f32 p, %1, %2;\n\tsetp.ge.and.f32 p, %3, %4, p;\n\tselp.f32 %0, %5, 0f00000000, p;\n\t}"
        : "=f"(am) : "f"(A), "f"(kTermA), "f"(sp), "f"(thr), "f"(a));
    return am;
}
__device__ __forceinline__ float live_select(float A, float w, float c) {
    float r;
    asm("{\n\t.reg .pred p;\n\tsetp.lt.f32 p, %1, %2;\n\tsetp.ge.and.f32 p, %3, %4, p;\n\tselp.f32 %0, %5, 0f00000000, p;\n\t}"
        : "=f"(r) : "f"(A), "f"(kTermA), "f"(w), "f"(kMinW), "f"(c));
    return r;
}

template <bool kFast, bool kPk = true>
__device__ __forceinline__ void eval_pair(float2 fpx, const EntryRow& r, float2 A, PairEval& ev) {
    ev.dx = add2s<kPk>(fpx, r.neg_mx);
    const float2 t = fma2s<kPk>(r.q00, ev.dx, r.qsdy);
    const float2 sp = fma2s<kPk>(ev.dx, t, r.q11dy2);          // -0.5*log2(e) * s
    ev.e = make_float2(ex2_approx(sp.x), ex2_approx(sp.y));
    ev.T = fma2s<kPk>(A, bc2(-1.0f), bc2(1.0f));                              // 1 - A, one rounding
    if (kFast) {
        // the exponent already holds log2(opacity): e = opacity * w, one multiply less per pixel.  `w`, `u` and `a` all
        // name that product here; the backward divides the opacity out once per entry.
        ev.a = make_float2(live_alpha(A.x, sp.x, r.op, ev.e.x), live_alpha(A.y, sp.y, r.op, ev.e.y));
        ev.w = ev.a;
        ev.u = ev.a;
        ev.contrib = mul2s<kPk>(ev.T, ev.a);
        ev.act0 = ev.a.x > 0.f;
        ev.act1 = ev.a.y > 0.f;
    } else {
        ev.w = make_float2(fminf(ev.e.x, 1.0f), fminf(ev.e.y, 1.0f));      // clamp(exp(.), 0, 1); exp >= 0
        ev.u = mul2s<kPk>(bc2(r.op), ev.w);
        ev.a = make_float2(__saturatef(ev.u.x), __saturatef(ev.u.y));      // clamp(opacity * w, 0, 1)
        const float2 c = mul2s<kPk>(ev.T, ev.a);
        // a > 0 and contrib > 0 follow from opacity > kTinyOpacity, w >= 1e-5 and 1 - A >= 0.005;
        // entries with opacity <= kTinyOpacity (or negative) were staged with opacity 0 and contribute
        // exact zeros (a = 0 => contrib = 0; the backward skips such an entry as a whole)
        ev.act0 = (A.x < kTermA) && (ev.w.x >= kMinW);
        ev.act1 = (A.y < kTermA) && (ev.w.y >= kMinW);
        ev.contrib = make_float2(live_select(A.x, ev.w.x, c.x), live_select(A.y, ev.w.y, c.y));
    }
}

// `lop` = log2(opacity) from the record (-inf for an entry staged with opacity 0); only the fast path uses it.
template <bool kFast>
__device__ __forceinline__ void load_entry_row(const float4& r0, const float4& r1, float lop, float fpy, EntryRow& row, float& dy) {
    dy = fpy - r0.y;
    row.neg_mx = bc2(-r0.x);
    row.q00 = bc2(r0.z);
    row.qsdy = bc2(r0.w * dy);
    if (kFast) {
        row.q11dy2 = bc2(r1.x * dy * dy + lop);
        row.op = kLog2MinW + lop;
    } else {
        row.q11dy2 = bc2(r1.x * dy * dy);
        row.op = r1.y;
    }
}

// Stages one list entry's record into shared memory; returns whether it is "regular" (fast path).
__device__ __forceinline__ bool stage_entry(const float4* __restrict__ rec, int id, float4* dst) {
    float4 q1 = __ldg(&rec[(int64_t)id * 3 + 1]);
    const float4 q2 = __ldg(&rec[(int64_t)id * 3 + 2]);
    const bool tiny = !(q1.y > kTinyOpacity);
    const bool regular = q2.z != 0.f;
    q1.y = tiny ? 0.f : q1.y;                        // opacity 0 => a = 0 => the entry adds exact zeros
    dst[0] = __ldg(&rec[(int64_t)id * 3 + 0]);
    dst[1] = q1;
    // shared-memory copy: {g, b, 1/opacity, log2 opacity} -- the flag has been read, its slot carries the reciprocal
    // the backward needs once per entry (dL/d opacity = sum(dL/da * a) / opacity)
    dst[2] = make_float4(q2.x, q2.y, tiny ? 0.f : __frcp_rn(q1.y), tiny ? __int_as_float(0xff800000) : q2.w);
    return regular;
}

// ---- asynchronous staging (GS_PREFETCH=1, off by default): the records of batch b+1 travel global -> shared
// with cp.async while batch b is being composited; the entry ids are read two batches ahead.  Measured: no gain
// (fwd 372.6 vs 371.9 us) or a loss (bwd 1053 vs 1014 us, the two extra live registers spill under the 128 cap) --
// the other resident warps already cover the staging latency (long_scoreboard is 6-8 % of stall samples).
#ifndef GS_PREFETCH
#define GS_PREFETCH 0
#endif
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void stage_entry_async(const float4* __restrict__ rec, int id, float4* dst) {
    const float4* src = rec + (int64_t)id * 3;
    cp_async16(dst + 0, src + 0);
    cp_async16(dst + 1, src + 1);
    cp_async16(dst + 2, src + 2);
}
// after the copy has landed: the lane that staged an entry applies the tiny-opacity rule and reads its flag
__device__ __forceinline__ bool fixup_entry(float4* dst) {
    const float op = dst[1].y;
    const bool tiny = !(op > kTinyOpacity);
    const bool regular = dst[2].z != 0.f;
    if (tiny) { dst[1].y = 0.f; dst[2].w = __int_as_float(0xff800000); }
    dst[2].z = tiny ? 0.f : __frcp_rn(op);
    return regular;
}

// One batch of the forward walk (cnt staged entries); returns how many of them were composited.  Every
// kExitStride entries the warp checks whether any pixel is still alive, so a tile stops within 8 entries of
// its last contributor instead of at the end of the 32-entry batch (5 % of the walk on config[1]); the
// backward inherits the shorter walk through tile_consumed.
constexpr int kExitStride = 8;
template <bool kFast, bool kTrack>
__device__ __forceinline__ int fwd_batch(const float4* srec, int cnt, int first_index, float fpy, const float2 (&fpx)[kPairs],
                                         float2 (&A)[kPairs], float2 (&Cr)[kPairs], float2 (&Cg)[kPairs],
                                         float2 (&Cb)[kPairs], float2 (&Ds)[kPairs], int (&ncons)[kPx]) {
    int done = 0;
    for (int j0 = 0; j0 < cnt; j0 += kExitStride) {
        if (j0) {
            bool alive = false;
#pragma unroll
            for (int p = 0; p < kPairs; ++p) alive |= (A[p].x < kTermA) | (A[p].y < kTermA);
            if (!__any_sync(0xffffffffu, alive)) break;
        }
        const int jn = min(j0 + kExitStride, cnt);
#pragma unroll 2
        for (int j = j0; j < jn; ++j) {
            const float4 r0 = srec[j * 3 + 0];          // mx, my, q00', qs'
            const float4 r1 = srec[j * 3 + 1];          // q11', opacity (0 if tiny), z, r
            const float4 r2 = srec[j * 3 + 2];          // g, b, 1/opacity, log2 opacity
            EntryRow row;
            float dy;
            load_entry_row<kFast>(r0, r1, r2.w, fpy, row, dy);
            const float2 cr = bc2(r1.w), cg = bc2(r2.x), cb = bc2(r2.y), z = bc2(r1.z);
#pragma unroll
            for (int p = 0; p < kPairs; ++p) {
                PairEval ev;
                if (p < kPairs - GS_FWD_SCALAR_PAIRS) {
                    eval_pair<kFast, true>(fpx[p], row, A[p], ev);
                    Cr[p] = fma2(ev.contrib, cr, Cr[p]);
                    Cg[p] = fma2(ev.contrib, cg, Cg[p]);
                    Cb[p] = fma2(ev.contrib, cb, Cb[p]);
                    Ds[p] = fma2(ev.contrib, z, Ds[p]);
                    A[p] = add2(A[p], ev.contrib);
                } else {
                    eval_pair<kFast, false>(fpx[p], row, A[p], ev);
                    Cr[p] = fma2s<false>(ev.contrib, cr, Cr[p]);
                    Cg[p] = fma2s<false>(ev.contrib, cg, Cg[p]);
                    Cb[p] = fma2s<false>(ev.contrib, cb, Cb[p]);
                    Ds[p] = fma2s<false>(ev.contrib, z, Ds[p]);
                    A[p] = add2s<false>(A[p], ev.contrib);
                }
                if (kTrack) {                           // renderer.py:352: the entry that terminates the pixel
                    if (ev.act0 && A[p].x >= kTermA) ncons[2 * p] = first_index + j + 1;
                    if (ev.act1 && A[p].y >= kTermA) ncons[2 * p + 1] = first_index + j + 1;
                }
            }
        }
        done = jn;
    }
    return done;
}

// Tile geometry for any tile size (the reference takes any, renderer.py:24).  A warp always works on a 16x16 pixel block
// (32 lanes x 1x8 strips).  A tile of T x T pixels is covered by sub x sub such blocks, sub = ceil(T / 16), each walked
// by its own warp over the tile's list; pixels of a block beyond the tile's (or the image's) edge are masked for good
// (the "never alive" sentinel).  T = 16: one block per tile and nothing to mask inside the image; T < 16: one block per
// tile with the lanes outside the T x T corner idle; T > 16: the blocks of a tile stop independently.
// A launch SLOT is one (tile, block) pair: slot = tile * sub^2 + block; tile_order / tile_consumed are indexed by slot.
struct TileGeom {
    int tile_size, tiles_x, tiles_y, sub;
};
__host__ __device__ inline TileGeom tile_geom(int img_w, int img_h, int tile_size) {
    TileGeom g;
    g.tile_size = tile_size;
    g.tiles_x = (img_w + tile_size - 1) / tile_size;
    g.tiles_y = (img_h + tile_size - 1) / tile_size;
    g.sub = (tile_size + kTile - 1) / kTile;
    return g;
}
struct WarpBlock {
    int tile, px0, py, x_lim, y_lim;
};
__device__ __forceinline__ WarpBlock locate_block(const TileGeom& g, int slot_id, int lane, int img_w, int img_h) {
    WarpBlock b;
    const int per_tile = g.sub * g.sub;
    b.tile = slot_id / per_tile;
    const int blk = slot_id - b.tile * per_tile;
    const int by = blk / g.sub, bx = blk - by * g.sub;
    const int ty = b.tile / g.tiles_x, tx = b.tile - ty * g.tiles_x;
    b.x_lim = min(img_w, (tx + 1) * g.tile_size);
    b.y_lim = min(img_h, (ty + 1) * g.tile_size);
    b.py = ty * g.tile_size + by * kTile + (lane >> 1);
    b.px0 = tx * g.tile_size + bx * kTile + (lane & 1) * kPx;
    return b;
}

template <bool kTrack>
__global__ void __launch_bounds__(32 * kWarpsPerCta, (GS_FWD_MINB + kWarpsPerCta - 1) / kWarpsPerCta)
raster_fwd_kernel(int img_w, int img_h, TileGeom geom, const int32_t* __restrict__ entry_ids,
                  const int2* __restrict__ tile_ranges, const float4* __restrict__ rec,
                  const float* __restrict__ bg_ptr, int any_visible_host, const int64_t* __restrict__ counters_dev,
                  const int32_t* __restrict__ tile_order, int list_cap, uint8_t* __restrict__ tile_flags,
                  int32_t* __restrict__ flag_count, int rerun, float* __restrict__ image, float* __restrict__ alpha,
                  float* __restrict__ depth, float4* __restrict__ pix_state, int32_t* __restrict__ n_consumed,
                  int32_t* __restrict__ tile_consumed) {
    __shared__ float4 srec_all[kWarpsPerCta][GS_PREFETCH ? 2 : 1][kBatch * 3];
    float4 (*srec)[kBatch * 3] = srec_all[threadIdx.x >> 5];

    // one warp per tile; warps never synchronise with each other, so a warp without a tile simply leaves
    const int slot = (int)blockIdx.x * kWarpsPerCta + (int)(threadIdx.x >> 5);
    grid_dependency_wait();                                        // launched early (PDL): the binning's output is final from here
    if (slot >= geom.tiles_x * geom.tiles_y * geom.sub * geom.sub) return;
    const int slot_id = tile_order ? tile_order[slot] : slot;      // any permutation of the slots
    const int lane = threadIdx.x & 31;
    const WarpBlock wb = locate_block(geom, slot_id, lane, img_w, img_h);
    const int tile = wb.tile;
    // truncated lists (tile sizes <= 16 only: slot == tile): the re-run pass touches only the tiles the first pass flagged
    if (rerun && (*flag_count == 0 || tile_flags[tile] == 0)) return;
    const int py = wb.py, px0 = wb.px0;
    const float bg0 = bg_ptr[0], bg1 = bg_ptr[1], bg2 = bg_ptr[2];
    const float fpy = (float)py;
    const bool any_visible = counters_dev ? (counters_dev[2] > 0) : (any_visible_host != 0);

    float2 fpx[kPairs], A[kPairs], Cr[kPairs], Cg[kPairs], Cb[kPairs], Ds[kPairs];
    int ncons[kPx];
#pragma unroll
    for (int p = 0; p < kPairs; ++p) {
        const bool in0 = (px0 + 2 * p < wb.x_lim) && (py < wb.y_lim);
        const bool in1 = (px0 + 2 * p + 1 < wb.x_lim) && (py < wb.y_lim);
        fpx[p] = make_float2((float)(px0 + 2 * p), (float)(px0 + 2 * p + 1));
        A[p] = make_float2(in0 ? 0.f : 2.0f, in1 ? 0.f : 2.0f);          // 2.0: never alive
        Ds[p] = bc2(0.f);
        Cr[p] = bc2(bg0); Cg[p] = bc2(bg1); Cb[p] = bc2(bg2);            // out_rgb starts at bg (renderer.py:273)
        ncons[2 * p] = ncons[2 * p + 1] = -1;
    }

    int2 range = tile_ranges[tile];
    const int full_end = range.y;
    if (list_cap > 0 && !rerun) range.y = min(range.y, range.x + list_cap);      // the stored prefix of this tile's list
    int walked = 0;
#if GS_PREFETCH
    int buf = 0;
    int id_nxt = (range.x + lane < range.y) ? entry_ids[range.x + lane] : -1;                 // ids of batch 0
    if (id_nxt >= 0) stage_entry_async(rec, id_nxt, &srec[0][lane * 3]);
    cp_async_commit();
    id_nxt = (range.x + kBatch + lane < range.y) ? entry_ids[range.x + kBatch + lane] : -1;   // ids of batch 1
    for (int base = range.x; base < range.y; base += kBatch, buf ^= 1) {
        bool alive = false;
#pragma unroll
        for (int p = 0; p < kPairs; ++p) alive |= (A[p].x < kTermA) | (A[p].y < kTermA);
        if (!__any_sync(0xffffffffu, alive)) break;
        const int cnt = min(kBatch, range.y - base);
        cp_async_wait_all();
        __syncwarp();                                   // this batch has landed; the previous one is fully read
        if (id_nxt >= 0) stage_entry_async(rec, id_nxt, &srec[buf ^ 1][lane * 3]);
        cp_async_commit();
        id_nxt = (base + 2 * kBatch + lane < range.y) ? entry_ids[base + 2 * kBatch + lane] : -1;
        bool regular = true;
        if (lane < cnt) regular = fixup_entry(&srec[buf][lane * 3]);
        const bool all_regular = __all_sync(0xffffffffu, regular);
        __syncwarp();
        walked = base - range.x + (all_regular ? fwd_batch<true, kTrack>(srec[buf], cnt, base - range.x, fpy, fpx, A, Cr, Cg, Cb, Ds, ncons)
                                               : fwd_batch<false, kTrack>(srec[buf], cnt, base - range.x, fpy, fpx, A, Cr, Cg, Cb, Ds, ncons));
    }
    cp_async_wait_all();                                // a prefetch may still be in flight after an early exit
#else
    for (int base = range.x; base < range.y; base += kBatch) {
        bool alive = false;
#pragma unroll
        for (int p = 0; p < kPairs; ++p) alive |= (A[p].x < kTermA) | (A[p].y < kTermA);
        if (!__any_sync(0xffffffffu, alive)) break;
        const int cnt = min(kBatch, range.y - base);
        __syncwarp();                                   // previous batch fully read
        bool regular = true;
        if (lane < cnt) regular = stage_entry(rec, entry_ids[base + lane], &srec[0][lane * 3]);
        const bool all_regular = __all_sync(0xffffffffu, regular);
        __syncwarp();
        walked = base - range.x + (all_regular ? fwd_batch<true, kTrack>(srec[0], cnt, base - range.x, fpy, fpx, A, Cr, Cg, Cb, Ds, ncons)
                                               : fwd_batch<false, kTrack>(srec[0], cnt, base - range.x, fpy, fpx, A, Cr, Cg, Cb, Ds, ncons));
    }
#endif

    if (list_cap > 0 && !rerun) {
        // every tile writes its flag (the array needs no zero-fill): did the walk stop at the end of a stored prefix with
        // pixels still alive?
        bool alive = false;
        if (range.y < full_end) {
#pragma unroll
            for (int p = 0; p < kPairs; ++p) alive |= (A[p].x < kTermA) | (A[p].y < kTermA);
        }
        const bool more = __any_sync(0xffffffffu, alive);
        if (lane == 0) {
            tile_flags[tile] = more ? 1 : 0;
            if (more) atomicAdd(flag_count, 1);
        }
    }

    // epilogue: renderer.py:359-367 (or :74-83 when nothing passed culling)
    const int64_t plane = (int64_t)img_w * img_h;
#pragma unroll
    for (int k = 0; k < kPx; ++k) {
        const int px = px0 + k;
        const float Ak = (k & 1) ? A[k >> 1].y : A[k >> 1].x;
        if (Ak < 1.5f) {                            // pixels outside the tile / image kept the 2.0 sentinel
            const int64_t p = (int64_t)py * img_w + px;
            const float Crk = (k & 1) ? Cr[k >> 1].y : Cr[k >> 1].x;
            const float Cgk = (k & 1) ? Cg[k >> 1].y : Cg[k >> 1].x;
            const float Cbk = (k & 1) ? Cb[k >> 1].y : Cb[k >> 1].x;
            const float Dsk = (k & 1) ? Ds[k >> 1].y : Ds[k >> 1].x;
            float o_r, o_g, o_b, o_a, o_d;
            if (any_visible) {
                const float om = sub_rn(1.f, Ak);
                o_r = __saturatef(add_rn(Crk, mul_rn(om, bg0)));
                o_g = __saturatef(add_rn(Cgk, mul_rn(om, bg1)));
                o_b = __saturatef(add_rn(Cbk, mul_rn(om, bg2)));
                o_a = __saturatef(Ak);
                o_d = div_rn(Dsk, add_rn(Ak, 1e-6f));
            } else {
                o_r = bg0; o_g = bg1; o_b = bg2; o_a = 0.f; o_d = 0.f;
            }
            image[p] = o_r;
            image[plane + p] = o_g;
            image[2 * plane + p] = o_b;
            alpha[p] = o_a;
            depth[p] = o_d;
            pix_state[p] = make_float4(Crk, Cgk, Cbk, Dsk);
            if (kTrack) n_consumed[p] = ncons[k] >= 0 ? ncons[k] : walked;
        }
    }
    // entries this tile loaded (batch granular): bounds the backward walk; the E of the byte formulas
    if (lane == 0) tile_consumed[slot_id] = walked;
}

// Backward: re-walks the list front to back with the forward's recurrence.  With
//   v_k = gC.colour_k + gDs.z_k + gA   and   Total = sum_k T_k a_k v_k = gC.(C - bg) + gDs.Dsum + gA.A
// (all known from the forward's saved state), the gradient w.r.t. a_k is
//   T_k v_k - (Total - prefix_k) / (1 - a_k),
// so no reverse traversal and no division-recovered transmittance is needed.  Every non-final
// contributor has a_k < 0.995 (else the pixel would have terminated there); for the terminating
// contributor the suffix is exactly zero.
// With s' = c*s (c = -0.5*log2 e) and w = 2^s':  h = dL/ds' = ln2 * w * dL/dw.  Per lane only
//   Sh = sum h,  Sx = sum h*dx,  Sxx = sum h*dx^2   (dy is the same for a lane's 8 pixels)
// are accumulated; the five conic / mean sums follow once per entry:
//   gQ00 = c*Sxx, gQ01 = gQ10 = c*dy*Sx, gQ11 = c*dy^2*Sh,
//   g_mx = -(2 q00' Sx + qs' dy Sh),  g_my = -(2 q11' dy Sh + qs' Sx).
constexpr int kBwdGroup = GS_BWD_GROUP;
static_assert(32 % kBwdGroup == 0, "kBatch must be a multiple of the group");
constexpr int kRedVals = 10;     // mx my | q00 q01 q11 | opacity | z | r g b
constexpr int kRedStride = 36;   // floats per value row: 32 lanes + pad, keeps LDS.128 aligned

struct BwdOut {
    float* base;              // lane v < 10: where reduced value v goes, base + id * stride
    int stride;
    float* red;               // [kRedVals][kRedStride] transpose buffer
    const float4* red_src;    // this lane's slice of it
    const int* sid;           // staged splat ids
#if GS_BWD_RAWSUMS
    // raw-sum reduction: the ten values crossing the warp are the lane's plain sums
    //   0 Sx  1 dy*Sh  2 Sxx  3 dy*Sx  4 dy^2*Sh  5 S_opacity  6 z  7 r  8 g  9 b
    // and the per-entry factors (conic, opacity, ln2, c) are applied once per entry AFTER the reduction, by the lane
    // that owns the output: out = coef[ja] * T[ia] + coef[jb] * T[ib], T fetched from lanes ia / ib by shuffle and the
    // entry's coefficient row {-2k q00', -k qs', -2k q11', c k, opacity factor, 1, 0, 0} (k = ln2, or ln2 * opacity on
    // the general path) written to shared memory by the lane that staged the entry.
    const float* coef;        // [kBatch][kCoefRow]
    int ia, ib, ja, jb;
#endif
};
constexpr int kCoefRow = 8;

// Second half of the per-entry reduction: lanes 0..29 each add a third of one value's 32 partials
// (12 + 12 + 8, read as float4), lanes v < 10 collect the three thirds and send ONE vector atomic.
// Lanes 30/31 and the third float4 of lanes >= 20 read in-bounds scratch that is never used.
__device__ __forceinline__ void reduce_finish(const BwdOut& out, int buf, int id, int lane, int entry) {
    const float4* src = out.red_src + buf * (kRedVals * kRedStride / 4);
    const float4 q0 = src[0], q1 = src[1], q2 = src[2];
    const float2 a01 = add2(make_float2(q0.x, q0.y), make_float2(q0.z, q0.w));
    const float2 a23 = add2(make_float2(q1.x, q1.y), make_float2(q1.z, q1.w));
    float2 a45 = add2(make_float2(q2.x, q2.y), make_float2(q2.z, q2.w));
    a45.x = (lane < 20) ? a45.x : 0.f;                              // thirds 0 and 1 hold 12 partials
    a45.y = (lane < 20) ? a45.y : 0.f;
    const float2 acc = add2(add2(a01, a23), a45);
    const float s = acc.x + acc.y;
    const float s2 = __shfl_down_sync(0xffffffffu, s, 10);
    const float s3 = __shfl_down_sync(0xffffffffu, s, 20);
    // lanes 0..9 hold the ten totals; lane 10 takes a copy of Q01's for Q10 (both enter s symmetrically), so that ONE
    // reduction instruction serves all eleven addresses.  32-bit element offset: id * stride < 2^31 for every array the
    // ABI accepts.
    float total = s + s2 + s3;
#if GS_BWD_RAWSUMS
    (void)entry;
    const float ta = __shfl_sync(0xffffffffu, total, out.ia), tb = __shfl_sync(0xffffffffu, total, out.ib);
    const float* row = out.coef + entry * kCoefRow;
    total = fmaf(row[out.jb], tb, row[out.ja] * ta);
#else
    (void)entry;
    const float q01 = __shfl_sync(0xffffffffu, total, 3);
    total = (lane == kRedVals) ? q01 : total;
#endif
    float* dst = out.base + (unsigned)id * (unsigned)out.stride;
    if (lane <= kRedVals) atomicAdd(dst, total);
}

#if GS_BWD_RAWSUMS
// Coefficient rows of one staged batch (see BwdOut); every lane writes the row of the entry it staged.
__device__ __forceinline__ void stage_coef(bool all_regular, int lane, int cnt_pad, const float4* srec, float* coef) {
    if (lane < cnt_pad) {
        const float4 q0 = srec[lane * 3 + 0], q1 = srec[lane * 3 + 1];
        const float inv_op = srec[lane * 3 + 2].z;
        const float op = q1.y;
        const float k = all_regular ? kLn2 : kLn2 * op;
        float4* row = reinterpret_cast<float4*>(coef + lane * kCoefRow);
        row[0] = make_float4(-2.f * k * q0.z, -k * q0.w, -2.f * k * q1.x, kNegHalfLog2e * k);
        row[1] = make_float4(all_regular ? inv_op : (op > 0.f ? 1.f : 0.f), 1.f, 0.f, 0.f);
    }
}
#endif

// Per-lane partial sums of one list entry over the lane's pixels.
struct BwdAcc {
    float2 s_h, s_x, s_xx, s_op, s_z, s_cr, s_cg, s_cb;
};

// Backward arithmetic of one list entry for ONE pixel pair (kPk: packed or scalar FP32 instructions, same results).
// kFirst: the lane's first pair of this entry initialises the partial sums instead of adding to zeros.
// kDepth: a depth gradient exists (otherwise gDs is identically zero: its term and the per-splat depth sum are dropped --
// the reference's own train step differentiates the image only, optimizer.py:137-139).
template <bool kFast, bool kPk, bool kFirst, bool kDepth>
__device__ __forceinline__ void bwd_pair(float2 fpx, const EntryRow& row, float2& A, float2& R, float2 gCr, float2 gCg, float2 gCb,
                                         float2 gDs, float2 gA, float2 cr, float2 cg, float2 cb, float2 z, BwdAcc& acc) {
    PairEval ev;
    eval_pair<kFast, kPk>(fpx, row, A, ev);
    const float2 v = fma2s<kPk>(gCr, cr, fma2s<kPk>(gCg, cg, fma2s<kPk>(gCb, cb, kDepth ? fma2s<kPk>(gDs, z, gA) : gA)));
    R = fma2s<kPk>(ev.contrib, v, R);
    A = add2s<kPk>(A, ev.contrib);
    // suffix / (1 - a); the terminating contributor has an empty suffix.  Otherwise
    // a < 0.995, so 1 - a >= 0.005 and the approximate reciprocal is safe (a pixel that this
    // entry terminates may see 1 - a ~ 0: its inf/NaN is discarded by the select).
    const float2 oma = fma2s<kPk>(ev.a, bc2(-1.0f), bc2(1.0f));
    float2 nsuf = mul2s<kPk>(R, make_float2(rcp_approx(oma.x), rcp_approx(oma.y)));
    nsuf.x = (A.x >= kTermA) ? 0.f : nsuf.x;
    nsuf.y = (A.y >= kTermA) ? 0.f : nsuf.y;
    float2 g_a = fma2s<kPk>(ev.T, v, nsuf);
    if (kFast) {
        // ev.a = opacity * w is already zero where the reference skips the splat, and g_a only ever appears multiplied
        // by it, so no further masking (both clamps are identities on this path).  With s'' = s' + log2(opacity) and
        // a = 2^s'':  dL/ds' = ln2 * (a * g_a), and dL/d opacity = sum(w * g_a) = sum(a * g_a) / opacity: the constant
        // factors are applied once per entry.
        const float2 ga = mul2s<kPk>(g_a, ev.a);
        const float2 gadx = mul2s<kPk>(ga, ev.dx);
        if (kFirst) {
            acc.s_op = ga;
            acc.s_x = gadx;
            acc.s_xx = mul2s<kPk>(gadx, ev.dx);
        } else {
            acc.s_op = add2s<kPk>(acc.s_op, ga);
            acc.s_x = add2s<kPk>(acc.s_x, gadx);
            acc.s_xx = fma2s<kPk>(gadx, ev.dx, acc.s_xx);
        }
    } else {
        // a = clamp(op*w, 0, 1), w = clamp(exp(-s/2), 0, 1): closed-interval pass-through
        g_a.x = (ev.act0 && ev.u.x <= 1.f) ? g_a.x : 0.f;
        g_a.y = (ev.act1 && ev.u.y <= 1.f) ? g_a.y : 0.f;
        float2 h = mul2s<kPk>(ev.w, g_a);
        acc.s_op = kFirst ? h : add2s<kPk>(acc.s_op, h);
        h.x = (ev.e.x <= 1.f) ? h.x : 0.f;
        h.y = (ev.e.y <= 1.f) ? h.y : 0.f;
        const float2 hdx = mul2s<kPk>(h, ev.dx);
        if (kFirst) {
            acc.s_h = h;
            acc.s_x = hdx;
            acc.s_xx = mul2s<kPk>(hdx, ev.dx);
        } else {
            acc.s_h = add2s<kPk>(acc.s_h, h);
            acc.s_x = add2s<kPk>(acc.s_x, hdx);
            acc.s_xx = fma2s<kPk>(hdx, ev.dx, acc.s_xx);
        }
    }
    if (kFirst) {
        acc.s_cr = mul2s<kPk>(ev.contrib, gCr);
        acc.s_cg = mul2s<kPk>(ev.contrib, gCg);
        acc.s_cb = mul2s<kPk>(ev.contrib, gCb);
        if (kDepth) acc.s_z = mul2s<kPk>(ev.contrib, gDs);
    } else {
        acc.s_cr = fma2s<kPk>(ev.contrib, gCr, acc.s_cr);
        acc.s_cg = fma2s<kPk>(ev.contrib, gCg, acc.s_cg);
        acc.s_cb = fma2s<kPk>(ev.contrib, gCb, acc.s_cb);
        if (kDepth) acc.s_z = fma2s<kPk>(ev.contrib, gDs, acc.s_z);
    }
}

// Arithmetic of ONE list entry for this lane's 8 pixels; leaves the 10 per-lane partial sums in
// rb[v * kRedStride + lane] (first half of the transpose reduction).  No barrier inside.
template <bool kFast, bool kDepth>
__device__ __forceinline__ void bwd_entry(const float4* srec_j, int lane, float fpy, const float2 (&fpx)[kPairs],
                                          float2 (&A)[kPairs], float2 (&R)[kPairs], const float2 (&gCr)[kPairs],
                                          const float2 (&gCg)[kPairs], const float2 (&gCb)[kPairs],
                                          const float2 (&gDs)[kPairs], const float2 (&gA)[kPairs], float* rb) {
    const float4 r0 = srec_j[0];
    const float4 r1 = srec_j[1];
    const float4 r2 = srec_j[2];                        // g, b, 1/opacity, log2 opacity
    const float op = r1.y;
    EntryRow row;
    float dy;
    load_entry_row<kFast>(r0, r1, r2.w, fpy, row, dy);
    const float2 cr = bc2(r1.w), cg = bc2(r2.x), cb = bc2(r2.y), z = bc2(r1.z);
    BwdAcc acc;
    acc.s_h = bc2(0.f);
    if (!kDepth) acc.s_z = bc2(0.f);
    bwd_pair<kFast, (kPairs > GS_BWD_SCALAR_PAIRS), true, kDepth>(fpx[0], row, A[0], R[0], gCr[0], gCg[0], gCb[0], gDs[0], gA[0], cr, cg, cb, z, acc);
#pragma unroll
    for (int p = 1; p < kPairs; ++p) {
        if (p < kPairs - GS_BWD_SCALAR_PAIRS)
            bwd_pair<kFast, true, false, kDepth>(fpx[p], row, A[p], R[p], gCr[p], gCg[p], gCb[p], gDs[p], gA[p], cr, cg, cb, z, acc);
        else
            bwd_pair<kFast, false, false, kDepth>(fpx[p], row, A[p], R[p], gCr[p], gCg[p], gCb[p], gDs[p], gA[p], cr, cg, cb, z, acc);
    }
    const float2 s_h = acc.s_h, s_x = acc.s_x, s_xx = acc.s_xx, s_op = acc.s_op, s_z = acc.s_z;
    const float2 s_cr = acc.s_cr, s_cg = acc.s_cg, s_cb = acc.s_cb;
#if GS_BWD_RAWSUMS
    // the lane's plain sums; conic / opacity / ln2 factors follow after the warp reduction (reduce_finish).  An entry staged
    // with opacity 0 has exact-zero sums on the fast path and a zero opacity coefficient on the general one.
    (void)op; (void)r0; (void)r1;
    const float Sop_all = s_op.x + s_op.y;
    const float Sh = kFast ? Sop_all : (s_h.x + s_h.y), Sx = s_x.x + s_x.y;
    const float dySh = dy * Sh;
    rb[0 * kRedStride + lane] = Sx;
    rb[1 * kRedStride + lane] = dySh;
    rb[2 * kRedStride + lane] = s_xx.x + s_xx.y;
    rb[3 * kRedStride + lane] = dy * Sx;
    rb[4 * kRedStride + lane] = dy * dySh;
    rb[5 * kRedStride + lane] = Sop_all;
    rb[6 * kRedStride + lane] = s_z.x + s_z.y;
    rb[7 * kRedStride + lane] = s_cr.x + s_cr.y;
    rb[8 * kRedStride + lane] = s_cg.x + s_cg.y;
    rb[9 * kRedStride + lane] = s_cb.x + s_cb.y;
}
#else
    // an entry staged with opacity 0 (<= kTinyOpacity, or negative) is one the reference skips (a <= 0):
    // every sum below is then an exact zero except the opacity one, which is forced to zero
    // (fast path: the pixel sums hold a * dL/da with a = opacity * w, so the opacity is divided out here and dL/ds' needs ln2 only)
    const float Sop_all = s_op.x + s_op.y;
    const float Sop = (op > 0.f) ? (kFast ? Sop_all * srec_j[2].z : Sop_all) : 0.f;     // 1/opacity: re-read, not kept live
    const float hs = kFast ? kLn2 : kLn2 * op;          // dL/ds' = ln2 * op * w * dL/da
    const float Sh = hs * (kFast ? Sop_all : (s_h.x + s_h.y)), Sx = hs * (s_x.x + s_x.y), Sxx = hs * (s_xx.x + s_xx.y);
    const float dySh = dy * Sh;
    rb[0 * kRedStride + lane] = -fmaf(2.f * r0.z, Sx, r0.w * dySh);             // g_mx
    rb[1 * kRedStride + lane] = -fmaf(2.f * r1.x, dySh, r0.w * Sx);             // g_my
    rb[2 * kRedStride + lane] = kNegHalfLog2e * Sxx;                            // g_Q00
    rb[3 * kRedStride + lane] = kNegHalfLog2e * (dy * Sx);                      // g_Q01 = g_Q10
    rb[4 * kRedStride + lane] = kNegHalfLog2e * (dy * dySh);                    // g_Q11
    rb[5 * kRedStride + lane] = Sop;
    rb[6 * kRedStride + lane] = s_z.x + s_z.y;
    rb[7 * kRedStride + lane] = s_cr.x + s_cr.y;
    rb[8 * kRedStride + lane] = s_cg.x + s_cg.y;
    rb[9 * kRedStride + lane] = s_cb.x + s_cb.y;
}
#endif

// One batch of the backward walk.  `cnt` is a multiple of kBwdGroup (the stager pads with null entries).
// The entries are taken kBwdGroup at a time: their arithmetic runs back to back with no barrier in
// between -- so the scheduler overlaps one entry's shared-memory and MUFU latencies with the next
// entry's independent work (measured: the barrier after every entry, not the atomics, was what bound
// this kernel; profiles/r1_v5_raster.md) -- then one __syncwarp, then the second halves of their reductions.
template <bool kFast, bool kDepth = true>
__device__ __forceinline__ void bwd_batch(const float4* srec, int cnt, int lane, float fpy, const float2 (&fpx)[kPairs],
                                          float2 (&A)[kPairs], float2 (&R)[kPairs], const float2 (&gCr)[kPairs],
                                          const float2 (&gCg)[kPairs], const float2 (&gCb)[kPairs],
                                          const float2 (&gDs)[kPairs], const float2 (&gA)[kPairs], const BwdOut& out) {
    for (int j = 0; j < cnt; j += kBwdGroup) {
#pragma unroll
        for (int g = 0; g < kBwdGroup; ++g)
            bwd_entry<kFast, kDepth>(srec + (j + g) * 3, lane, fpy, fpx, A, R, gCr, gCg, gCb, gDs, gA,
                             out.red + g * (kRedVals * kRedStride));
        __syncwarp();
#pragma unroll
        for (int g = 0; g < kBwdGroup; ++g) reduce_finish(out, g, out.sid[j + g], lane, j + g);
        __syncwarp();
    }
}

template <bool kDepth>
__global__ void __launch_bounds__(32 * kWarpsPerCta, (GS_BWD_MINB + kWarpsPerCta - 1) / kWarpsPerCta)
raster_bwd_kernel(int img_w, int img_h, TileGeom geom, const int32_t* __restrict__ entry_ids,
                  const int2* __restrict__ tile_ranges, const float4* __restrict__ rec,
                  const float* __restrict__ bg_ptr, const float* __restrict__ alpha,
                  const float4* __restrict__ pix_state, const int32_t* __restrict__ tile_consumed,
                  const int32_t* __restrict__ tile_order, const float* __restrict__ g_image, const float* __restrict__ g_alpha,
                  const float* __restrict__ g_depth,
                  float* __restrict__ g_means2d, float* __restrict__ g_conics, float* __restrict__ g_depths,
                  float* __restrict__ g_colors, float* __restrict__ g_opac) {
    __shared__ float4 srec_all[kWarpsPerCta][GS_PREFETCH ? 2 : 1][kBatch * 3];
    __shared__ int sid_all[kWarpsPerCta][GS_PREFETCH ? 2 : 1][kBatch];
    __shared__ __align__(16) float red_all[kWarpsPerCta][kBwdGroup * kRedVals * kRedStride + 16];
    float4 (*srec)[kBatch * 3] = srec_all[threadIdx.x >> 5];
    int (*sid)[kBatch] = sid_all[threadIdx.x >> 5];
    float* red = red_all[threadIdx.x >> 5];
#if GS_BWD_RAWSUMS
    __shared__ __align__(16) float coef_all[kWarpsPerCta][kBatch * kCoefRow];
    float* coef = coef_all[threadIdx.x >> 5];
#endif

    const int slot = (int)blockIdx.x * kWarpsPerCta + (int)(threadIdx.x >> 5);
    grid_dependency_wait();                                        // launched early (PDL)
    if (slot >= geom.tiles_x * geom.tiles_y * geom.sub * geom.sub) return;
    const int slot_id = tile_order ? tile_order[slot] : slot;      // any permutation of the slots
    const int lane = threadIdx.x & 31;
    const WarpBlock wb = locate_block(geom, slot_id, lane, img_w, img_h);
    const int tile = wb.tile;
    const int py = wb.py, px0 = wb.px0;
    const float bg0 = bg_ptr[0], bg1 = bg_ptr[1], bg2 = bg_ptr[2];
    const float fpy = (float)py;
    const int64_t plane = (int64_t)img_w * img_h;

    // where lane v < 10 sends reduced value v:  target = out_base + id * out_stride
    float* out_base;
    int out_stride;
    switch (lane) {
        case 0: out_base = g_means2d; out_stride = 2; break;
        case 1: out_base = g_means2d + 1; out_stride = 2; break;
        case 2: out_base = g_conics; out_stride = 4; break;
        case 3: out_base = g_conics + 1; out_stride = 4; break;   // Q01 (and Q10 below)
        case 4: out_base = g_conics + 3; out_stride = 4; break;
        case 5: out_base = g_opac; out_stride = 1; break;
        case 6: out_base = g_depths; out_stride = 1; break;
        case 7: out_base = g_colors; out_stride = 3; break;
        case 8: out_base = g_colors + 1; out_stride = 3; break;
        case 9: out_base = g_colors + 2; out_stride = 3; break;
        default: out_base = g_conics + 2; out_stride = 4; break;  // lane 10: Q10 (a copy of Q01's sum); lanes > 10 never store
    }
    // lanes 0..29: value red_v, third red_g of the 32 partials (12 + 12 + 8)
    const int red_v = lane % kRedVals, red_g = lane / kRedVals;
    const float4* red_src = reinterpret_cast<const float4*>(&red[red_v * kRedStride + red_g * 12]);

    // R = -(Total - prefix): minus what the contributors still to come will add (the suffix sum)
    float2 fpx[kPairs], A[kPairs], R[kPairs], gCr[kPairs], gCg[kPairs], gCb[kPairs], gDs[kPairs], gA[kPairs];
#pragma unroll
    for (int p = 0; p < kPairs; ++p) {
        float Ai[2], Ti[2], gr[2], gg[2], gb[2], gd_[2], ga_[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int px = px0 + 2 * p + h;
            const bool inside = px < wb.x_lim && py < wb.y_lim;
            Ai[h] = inside ? 0.f : 2.0f;
            Ti[h] = gr[h] = gg[h] = gb[h] = gd_[h] = ga_[h] = 0.f;
            if (inside) {
                const int64_t q = (int64_t)py * img_w + px;
                const float Af = alpha[q];                   // A is always inside [0,1], so alpha == A
                const float4 st = pix_state[q];
                const float om = sub_rn(1.f, Af);
                const float pre_r = add_rn(st.x, mul_rn(om, bg0));
                const float pre_g = add_rn(st.y, mul_rn(om, bg1));
                const float pre_b = add_rn(st.z, mul_rn(om, bg2));
                // torch.clamp passes the gradient on the closed interval
                gr[h] = (pre_r >= 0.f && pre_r <= 1.f) ? g_image[q] : 0.f;
                gg[h] = (pre_g >= 0.f && pre_g <= 1.f) ? g_image[plane + q] : 0.f;
                gb[h] = (pre_b >= 0.f && pre_b <= 1.f) ? g_image[2 * plane + q] : 0.f;
                const float gd = kDepth ? g_depth[q] : 0.f;
                const float den = add_rn(Af, 1e-6f);
                gd_[h] = gd / den;
                ga_[h] = ((kDepth || g_alpha) ? g_alpha[q] : 0.f) - (gr[h] * bg0 + gg[h] * bg1 + gb[h] * bg2) - gd * st.w / (den * den);
                Ti[h] = gr[h] * (st.x - bg0) + gg[h] * (st.y - bg1) + gb[h] * (st.z - bg2) + gd_[h] * st.w + ga_[h] * Af;
            }
        }
        fpx[p] = make_float2((float)(px0 + 2 * p), (float)(px0 + 2 * p + 1));
        A[p] = make_float2(Ai[0], Ai[1]);
        R[p] = make_float2(-Ti[0], -Ti[1]);
        gCr[p] = make_float2(gr[0], gr[1]); gCg[p] = make_float2(gg[0], gg[1]); gCb[p] = make_float2(gb[0], gb[1]);
        gDs[p] = make_float2(gd_[0], gd_[1]); gA[p] = make_float2(ga_[0], ga_[1]);
    }

    BwdOut out;
    out.base = out_base;
    out.stride = out_stride;
    out.red = red;
    out.red_src = red_src;
    out.sid = sid[0];
#if GS_BWD_RAWSUMS
    out.coef = coef;
    // which reduced sums (ia, ib) and which entries of the coefficient row (ja, jb) make this lane's output; lanes > 10
    // compute a finite value nobody stores (jb = 6 is the row's constant 0)
    out.ia = lane <= 9 ? lane : 3;                    // lane 10: Q10 = Q01's sum (value 3)
    out.ib = lane == 0 ? 1 : (lane == 1 ? 0 : out.ia);
    out.ja = lane == 0 ? 0 : lane == 1 ? 2 : (lane <= 4 || lane == 10) ? 3 : lane == 5 ? 4 : 5;
    out.jb = lane <= 1 ? 1 : 6;
#endif

    const int2 range = tile_ranges[tile];
    const int end = range.x + tile_consumed[slot_id];
#if GS_PREFETCH
    int buf = 0;
    int id_cur = -1;
    int id_nxt = (range.x + lane < end) ? entry_ids[range.x + lane] : -1;                     // ids of batch 0
    if (id_nxt >= 0) stage_entry_async(rec, id_nxt, &srec[0][lane * 3]);
    cp_async_commit();
    id_cur = id_nxt;
    id_nxt = (range.x + kBatch + lane < end) ? entry_ids[range.x + kBatch + lane] : -1;       // ids of batch 1
    for (int base = range.x; base < end; base += kBatch, buf ^= 1) {
        bool alive = false;
#pragma unroll
        for (int p = 0; p < kPairs; ++p) alive |= (A[p].x < kTermA) | (A[p].y < kTermA);
        if (!__any_sync(0xffffffffu, alive)) break;
        const int cnt = min(kBatch, end - base);
        const int cnt_pad = (cnt + kBwdGroup - 1) / kBwdGroup * kBwdGroup;
        cp_async_wait_all();
        __syncwarp();                                   // this batch has landed; the previous one is fully read
        if (id_nxt >= 0) stage_entry_async(rec, id_nxt, &srec[buf ^ 1][lane * 3]);
        cp_async_commit();
        const int id_nn = (base + 2 * kBatch + lane < end) ? entry_ids[base + 2 * kBatch + lane] : -1;
        float4* mine = &srec[buf][lane * 3];
        const int first_id = __shfl_sync(0xffffffffu, id_cur, 0);      // by all lanes, outside the divergent part
        bool regular = true;
        if (lane < cnt) {
            sid[buf][lane] = id_cur;
            regular = fixup_entry(mine);
        } else if (lane < cnt_pad) {
            // null entry: far outside every tile (weight exp2(-1e12) = 0) and opacity 0, so all ten sums are
            // exact zeros; they are added to the batch's first splat
            sid[buf][lane] = first_id;
            mine[0] = make_float4(1.0e6f, 0.f, -1.f, 0.f);
            mine[1] = make_float4(0.f, 0.f, 0.f, 0.f);
            mine[2] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const bool all_regular = __all_sync(0xffffffffu, regular);
        __syncwarp();
#if GS_BWD_RAWSUMS
        stage_coef(all_regular, lane, cnt_pad, srec[buf], coef);
        __syncwarp();
#endif
        out.sid = sid[buf];
        if (all_regular) bwd_batch<true, kDepth>(srec[buf], cnt_pad, lane, fpy, fpx, A, R, gCr, gCg, gCb, gDs, gA, out);
        else bwd_batch<false, kDepth>(srec[buf], cnt_pad, lane, fpy, fpx, A, R, gCr, gCg, gCb, gDs, gA, out);
        id_cur = id_nxt;
        id_nxt = id_nn;
    }
    cp_async_wait_all();
#else
    for (int base = range.x; base < end; base += kBatch) {
        bool alive = false;
#pragma unroll
        for (int p = 0; p < kPairs; ++p) alive |= (A[p].x < kTermA) | (A[p].y < kTermA);
        if (!__any_sync(0xffffffffu, alive)) break;
        const int cnt = min(kBatch, end - base);
        __syncwarp();
        const int cnt_pad = (cnt + kBwdGroup - 1) / kBwdGroup * kBwdGroup;
        bool regular = true;
        if (lane < cnt) {
            const int id = entry_ids[base + lane];
            sid[0][lane] = id;
            regular = stage_entry(rec, id, &srec[0][lane * 3]);
        } else if (lane < cnt_pad) {
            sid[0][lane] = entry_ids[base];
            srec[0][lane * 3 + 0] = make_float4(1.0e6f, 0.f, -1.f, 0.f);
            srec[0][lane * 3 + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
            srec[0][lane * 3 + 2] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const bool all_regular = __all_sync(0xffffffffu, regular);
        __syncwarp();
#if GS_BWD_RAWSUMS
        stage_coef(all_regular, lane, cnt_pad, srec[0], coef);
        __syncwarp();
#endif
        out.sid = sid[0];
        if (all_regular) bwd_batch<true, kDepth>(srec[0], cnt_pad, lane, fpy, fpx, A, R, gCr, gCg, gCb, gDs, gA, out);
        else bwd_batch<false, kDepth>(srec[0], cnt_pad, lane, fpy, fpx, A, R, gCr, gCg, gCb, gDs, gA, out);
    }
#endif
}

// Longest-first launch order for the backward pass.  The work of a tile is known exactly (tile_consumed,
// written by the forward) and varies a lot (config[1]: 0..492 entries, mean 312, std 128), while the grid is
// only ~3.4 waves of one-warp CTAs: in raster order the last wave still holds full-size tiles.  One small
// block buckets the tiles by consumed/8 (counting sort, heaviest bucket first); CTAs are dispatched in
// blockIdx order, so the light tiles fill the tail.
constexpr int kOrderBuckets = 256;
__global__ void __launch_bounds__(1024)
tile_order_kernel(int num_tiles, const int32_t* __restrict__ tile_consumed, const int2* __restrict__ tile_ranges,
                  int32_t* __restrict__ order) {
    __shared__ int s_cnt[kOrderBuckets];
    __shared__ int s_off[kOrderBuckets];
    const int tid = threadIdx.x;
    if (tid < kOrderBuckets) s_cnt[tid] = 0;
    __syncthreads();
    // work estimate of a tile: what it consumed (exact, backward), or -- before the forward has run -- the length of
    // its list: a short list IS the work, a long one saturates somewhere, so long lists only need to come first
    auto bucket = [&](int t) {
        int work;
        if (tile_consumed) work = tile_consumed[t];
        else { const int2 r = tile_ranges[t]; work = r.y - r.x; }
        return kOrderBuckets - 1 - min(kOrderBuckets - 1, work >> 3);
    };
    for (int t = tid; t < num_tiles; t += blockDim.x) atomicAdd(&s_cnt[bucket(t)], 1);
    __syncthreads();
    if (tid < 32) {                                   // exclusive scan of 256 counts by one warp (8 per lane)
        int local[kOrderBuckets / 32], sum = 0;
#pragma unroll
        for (int q = 0; q < kOrderBuckets / 32; ++q) { local[q] = sum; sum += s_cnt[tid * (kOrderBuckets / 32) + q]; }
        int inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int x = __shfl_up_sync(0xffffffffu, inc, o);
            if (tid >= o) inc += x;
        }
#pragma unroll
        for (int q = 0; q < kOrderBuckets / 32; ++q) s_off[tid * (kOrderBuckets / 32) + q] = inc - sum + local[q];
    }
    __syncthreads();
    for (int t = tid; t < num_tiles; t += blockDim.x) order[atomicAdd(&s_off[bucket(t)], 1)] = t;
}

}  // namespace gs

using namespace gs;

static int check_raster_args(int32_t img_w, int32_t img_h, int32_t tile_size, const char* fn) {
    if (img_w <= 0 || img_h <= 0) {
        set_error("%s: image size must be positive", fn);
        return GS_ERR_INVALID_ARGUMENT;
    }
    if (tile_size < 1 || tile_size > 4096) {
        set_error("%s: tile_size %d out of range", fn, tile_size);
        return GS_ERR_INVALID_ARGUMENT;
    }
    return GS_OK;
}

extern "C" int gs_raster_fwd(int32_t img_w, int32_t img_h, int32_t tile_size, const int32_t* entry_ids,
                             const int32_t* tile_ranges, const float* splat_rec, const float* bg,
                             int32_t any_visible_host, const int64_t* counters_dev, const int32_t* tile_order,
                             int32_t list_cap, uint8_t* tile_flags, int32_t* flag_count, int32_t rerun, float* image,
                             float* alpha, float* depth, float* pix_state, int32_t* n_consumed, int32_t* tile_consumed,
                             void* stream) {
    const int rc = check_raster_args(img_w, img_h, tile_size, "gs_raster_fwd");
    if (rc != GS_OK) return rc;
    GS_REQUIRE(tile_ranges && bg && image && alpha && depth && pix_state && tile_consumed, "NULL array argument");
    GS_REQUIRE((list_cap <= 0 && !rerun) || (tile_flags && flag_count), "truncated lists need tile_flags and flag_count");
    if (list_cap <= 0) list_cap = 0;
    GS_REQUIRE((list_cap == 0 && !rerun) || tile_size <= kTile, "truncated lists need tile_size <= 16 (one warp per tile)");
    DeviceGuard guard(image);
    const TileGeom geom = tile_geom(img_w, img_h, tile_size);
    const int64_t slots64 = (int64_t)geom.tiles_x * geom.tiles_y * geom.sub * geom.sub;
    GS_REQUIRE(slots64 < (1ll << 31), "too many tiles");
    const int slots = (int)slots64;
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 grid((unsigned)((slots + kWarpsPerCta - 1) / kWarpsPerCta)), block(32 * kWarpsPerCta);
    if (n_consumed) {
        GS_CUDA_TRY(launch_pdl(raster_fwd_kernel<true>, grid, block, 0, st,
                               img_w, img_h, geom, entry_ids, (const int2*)tile_ranges, (const float4*)splat_rec, bg, any_visible_host,
                               counters_dev, tile_order, list_cap, tile_flags, flag_count, rerun, image, alpha, depth,
                               (float4*)pix_state, n_consumed, tile_consumed));
    } else {
        GS_CUDA_TRY(launch_pdl(raster_fwd_kernel<false>, grid, block, 0, st,
                               img_w, img_h, geom, entry_ids, (const int2*)tile_ranges, (const float4*)splat_rec, bg, any_visible_host,
                               counters_dev, tile_order, list_cap, tile_flags, flag_count, rerun, image, alpha, depth,
                               (float4*)pix_state, (int32_t*)nullptr, tile_consumed));
    }
    count_launches(1);
    return GS_OK;
}

extern "C" int gs_raster_bwd(int32_t img_w, int32_t img_h, int32_t tile_size, const int32_t* entry_ids,
                             const int32_t* tile_ranges, const float* splat_rec, const float* bg, const float* alpha,
                             const float* pix_state, const int32_t* tile_consumed, int32_t* tile_order_scratch,
                             int32_t tile_order_ready,
                             const float* g_image, const float* g_alpha, const float* g_depth, float* g_means2d,
                             float* g_conics, float* g_depths, float* g_colors, float* g_opacities, void* stream) {
    const int rc = check_raster_args(img_w, img_h, tile_size, "gs_raster_bwd");
    if (rc != GS_OK) return rc;
    GS_REQUIRE(tile_ranges && bg && alpha && pix_state && tile_consumed && g_image && g_means2d &&
                   g_conics && g_depths && g_colors && g_opacities, "NULL array argument");
    DeviceGuard guard(alpha);
    const TileGeom geom = tile_geom(img_w, img_h, tile_size);
    const int64_t slots64 = (int64_t)geom.tiles_x * geom.tiles_y * geom.sub * geom.sub;
    GS_REQUIRE(slots64 < (1ll << 31), "too many tiles");
    const int slots = (int)slots64;
    if (tile_order_scratch && !tile_order_ready) {
        tile_order_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(slots, tile_consumed, nullptr, tile_order_scratch);
        GS_CUDA_TRY(cudaGetLastError());
        count_launches(1);
    }
    // g_alpha / g_depth may be NULL (no gradient flows into that output); without a depth gradient a leaner instantiation runs
    GS_REQUIRE(g_depth == nullptr || g_alpha != nullptr, "g_depth without g_alpha: pass zeros for g_alpha");
    const dim3 grid((unsigned)((slots + kWarpsPerCta - 1) / kWarpsPerCta)), block(32 * kWarpsPerCta);
    if (g_depth)
        GS_CUDA_TRY(launch_pdl(raster_bwd_kernel<true>, grid, block, 0, (cudaStream_t)stream,
                               img_w, img_h, geom, entry_ids, (const int2*)tile_ranges, (const float4*)splat_rec, bg, alpha,
                               (const float4*)pix_state, tile_consumed, (const int32_t*)tile_order_scratch, g_image, g_alpha, g_depth,
                               g_means2d, g_conics, g_depths, g_colors, g_opacities));
    else
        GS_CUDA_TRY(launch_pdl(raster_bwd_kernel<false>, grid, block, 0, (cudaStream_t)stream,
                               img_w, img_h, geom, entry_ids, (const int2*)tile_ranges, (const float4*)splat_rec, bg, alpha,
                               (const float4*)pix_state, tile_consumed, (const int32_t*)tile_order_scratch, g_image, g_alpha,
                               (const float*)nullptr, g_means2d, g_conics, g_depths, g_colors, g_opacities));
    count_launches(1);
    return GS_OK;
}

extern "C" int gs_tile_order(int32_t num_tiles, const int32_t* tile_consumed, const int32_t* tile_ranges, int32_t* tile_order,
                             void* stream) {
    GS_REQUIRE(num_tiles > 0 && (tile_consumed || tile_ranges) && tile_order, "bad arguments");
    DeviceGuard guard(tile_order);
    tile_order_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(num_tiles, tile_consumed, (const int2*)tile_ranges, tile_order);
    GS_CUDA_TRY(cudaGetLastError());
    count_launches(1);
    return GS_OK;
}
