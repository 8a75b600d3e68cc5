// Stage R: 16x16-tile front-to-back alpha compositing, forward and backward.
//
// Reference semantics: the pixel loop and epilogue of GaussianRenderer._tile_rasterization,
// src/core/renderer.py:300-367.
//
// Both kernels are bound by instruction issue, not by HBM (ncu, profiles/r1_v1_*: DRAM 0.6 %,
// issue slots 81-90 % busy): a tile saturates after a few hundred of its thousands of list
// entries, so ~0.3 GB moves while ~6e8 pixel x splat evaluations execute.  The design therefore
// minimises instructions per evaluation:
//   * one warp per tile, each lane owns a 1x8 pixel strip (16 rows x 2 strips): the per-entry
//     work (shared-memory broadcast of the 44-byte record, the dy terms, loop control) is
//     amortised over 8 evaluations, 8 independent dependency chains hide FP32/MUFU latency, and no
//     block-wide barrier exists -- only __syncwarp;
//   * branch-free evaluation: the splat weight is one MUFU.EX2 on a quadratic form whose
//     coefficients were pre-multiplied by -0.5*log2(e) when project_fwd packed the record, and the
//     skip / termination rules of the reference are predicates on the accumulation, not branches;
//   * backward: each lane pre-sums its 8 pixels, then the 10 per-splat sums are reduced across the
//     warp through a padded shared-memory transpose (10 STS + 11 LDS per lane instead of 50
//     shuffles) and leave as ONE vector atomic instruction (lanes 0..9 -> 10 addresses).
//
// Accuracy of the weight: ex2.approx on the pre-scaled form differs from the reference's
// exp(-0.5*s) by <= ~2e-7 relative -- the same size as the error of an "accurate" expf, whose own
// last step is the same MUFU.EX2.  The accumulated opacity A, which decides termination, keeps the
// reference's recurrence A += (1-A)*a with separately rounded operations; forward and backward
// share one inline evaluation function, so they walk bit-identical contributor sets.
#include "common.cuh"

namespace gs {

constexpr int kPx = 8;                 // pixels per lane (1x8 strip) = 4 packed pairs
constexpr int kPairs = kPx / 2;
constexpr int kBatch = 32;             // list entries staged per round (one per lane)
constexpr float kTermA = 0.995f;       // renderer.py:352
constexpr float kMinW = 1e-5f;         // renderer.py:336
constexpr float kLn2 = 0.69314718056f;
constexpr float kNegHalfLog2e = -0.72134752044448170f;
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

// Blackwell packed fp32: one issue slot, two IEEE-rounded fp32 results (FFMA2 / FADD2 / FMUL2).
// Each lane keeps its 8 pixels as 4 (even, odd) pairs so the per-pixel arithmetic issues at half
// the instruction count -- the kernels are issue-bound, so this is the sm_100-specific lever.
__device__ __forceinline__ float2 bc2(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }

// Per-entry values shared by a lane's 8 pixels, pre-broadcast into pairs.
struct EntryRow {
    float2 neg_mx, q00, qsdy, q11dy2;
    float op;
};

// renderer.py:330-346 for one pixel pair.  `A >= kTermA` encodes "this pixel has terminated (or
// lies outside the image)", so no separate flag is carried.  Outputs: dx, e = exp(-s/2) (unclamped),
// w = clamp(e), u = opacity*w (unclamped), contrib = (1-A)*clamp(u) ZEROED where the reference skips
// the splat, and the two predicates.  Forward and backward both call this, so they agree bit for bit.
struct PairEval {
    float2 dx, e, w, u, a, T, contrib;
    bool act0, act1;
};

__device__ __forceinline__ void eval_pair(float2 fpx, const EntryRow& r, float2 A, PairEval& ev) {
    ev.dx = add2(fpx, r.neg_mx);
    const float2 t = fma2(r.q00, ev.dx, r.qsdy);
    const float2 sp = fma2(ev.dx, t, r.q11dy2);          // -0.5*log2(e) * s
    ev.e = make_float2(ex2_approx(sp.x), ex2_approx(sp.y));
    ev.w = make_float2(fminf(ev.e.x, 1.0f), fminf(ev.e.y, 1.0f));      // clamp(exp(.), 0, 1); exp >= 0
    ev.u = mul2(bc2(r.op), ev.w);
    ev.a = make_float2(__saturatef(ev.u.x), __saturatef(ev.u.y));      // clamp(opacity * w, 0, 1)
    ev.T = fma2(A, bc2(-1.0f), bc2(1.0f));                              // 1 - A, one rounding
    const float2 c = mul2(ev.T, ev.a);
    // a > 0 and contrib > 0 follow from opacity > kTinyOpacity, w >= 1e-5 and 1 - A >= 0.005;
    // entries with opacity <= kTinyOpacity were staged with opacity 0 and contribute exact zeros
    // (a = 0 => contrib = 0; the backward skips such an entry as a whole because its sum of
    // contributions is zero)
    ev.act0 = (A.x < kTermA) && (ev.w.x >= kMinW);
    ev.act1 = (A.y < kTermA) && (ev.w.y >= kMinW);
    ev.contrib = make_float2(ev.act0 ? c.x : 0.f, ev.act1 ? c.y : 0.f);
}

__device__ __forceinline__ void load_entry_row(const float4& r0, const float4& r1, float fpy, EntryRow& row, float& dy) {
    dy = fpy - r0.y;
    row.neg_mx = bc2(-r0.x);
    row.q00 = bc2(r0.z);
    row.qsdy = bc2(r0.w * dy);
    row.q11dy2 = bc2(r1.x * dy * dy);
    row.op = r1.y;
}

template <bool kTrack>
__global__ void __launch_bounds__(32)
raster_fwd_kernel(int img_w, int img_h, int tiles_x, const int32_t* __restrict__ entry_ids,
                  const int2* __restrict__ tile_ranges, const float4* __restrict__ rec,
                  const float* __restrict__ bg_ptr, int any_visible,
                  float* __restrict__ image, float* __restrict__ alpha, float* __restrict__ depth,
                  float4* __restrict__ pix_state, int32_t* __restrict__ n_consumed,
                  int32_t* __restrict__ tile_consumed) {
    __shared__ float4 srec[kBatch * 3];

    const int tile = blockIdx.x;
    const int lane = threadIdx.x;
    const int tx = tile % tiles_x, ty = tile / tiles_x;
    const int py = ty * kTile + (lane >> 1);
    const int px0 = tx * kTile + (lane & 1) * kPx;
    const float bg0 = bg_ptr[0], bg1 = bg_ptr[1], bg2 = bg_ptr[2];
    const float fpy = (float)py;

    float2 fpx[kPairs], A[kPairs], Cr[kPairs], Cg[kPairs], Cb[kPairs], Ds[kPairs];
    int ncons[kPx];
#pragma unroll
    for (int p = 0; p < kPairs; ++p) {
        const bool in0 = (px0 + 2 * p < img_w) && (py < img_h);
        const bool in1 = (px0 + 2 * p + 1 < img_w) && (py < img_h);
        fpx[p] = make_float2((float)(px0 + 2 * p), (float)(px0 + 2 * p + 1));
        A[p] = make_float2(in0 ? 0.f : 2.0f, in1 ? 0.f : 2.0f);          // 2.0: never alive
        Ds[p] = bc2(0.f);
        Cr[p] = bc2(bg0); Cg[p] = bc2(bg1); Cb[p] = bc2(bg2);            // out_rgb starts at bg (renderer.py:273)
        ncons[2 * p] = ncons[2 * p + 1] = -1;
    }

    const int2 range = tile_ranges[tile];
    int walked = 0;
    for (int base = range.x; base < range.y; base += kBatch) {
        bool alive = false;
#pragma unroll
        for (int p = 0; p < kPairs; ++p) alive |= (A[p].x < kTermA) | (A[p].y < kTermA);
        if (!__any_sync(0xffffffffu, alive)) break;
        const int cnt = min(kBatch, range.y - base);
        __syncwarp();                                   // previous batch fully read
        if (lane < cnt) {
            const int id = entry_ids[base + lane];
            float4 q1 = __ldg(&rec[(int64_t)id * 3 + 1]);
            q1.y = (q1.y > kTinyOpacity) ? q1.y : 0.f;      // opacity 0 => a = 0 => the entry adds exact zeros
            srec[lane * 3 + 0] = __ldg(&rec[(int64_t)id * 3 + 0]);
            srec[lane * 3 + 1] = q1;
            srec[lane * 3 + 2] = __ldg(&rec[(int64_t)id * 3 + 2]);
        }
        __syncwarp();
        walked = base - range.x + cnt;
#pragma unroll 2
        for (int j = 0; j < cnt; ++j) {
            const float4 r0 = srec[j * 3 + 0];          // mx, my, q00', qs'
            const float4 r1 = srec[j * 3 + 1];          // q11', opacity (0 if tiny), z, r
            const float2 r2 = *reinterpret_cast<const float2*>(&srec[j * 3 + 2]);   // g, b
            EntryRow row;
            float dy;
            load_entry_row(r0, r1, fpy, row, dy);
            const float2 cr = bc2(r1.w), cg = bc2(r2.x), cb = bc2(r2.y), z = bc2(r1.z);
#pragma unroll
            for (int p = 0; p < kPairs; ++p) {
                PairEval ev;
                eval_pair(fpx[p], row, A[p], ev);
                Cr[p] = fma2(ev.contrib, cr, Cr[p]);
                Cg[p] = fma2(ev.contrib, cg, Cg[p]);
                Cb[p] = fma2(ev.contrib, cb, Cb[p]);
                Ds[p] = fma2(ev.contrib, z, Ds[p]);
                A[p] = add2(A[p], ev.contrib);
                if (kTrack) {                           // renderer.py:352: the entry that terminates the pixel
                    if (ev.act0 && A[p].x >= kTermA) ncons[2 * p] = base - range.x + j + 1;
                    if (ev.act1 && A[p].y >= kTermA) ncons[2 * p + 1] = base - range.x + j + 1;
                }
            }
        }
    }

    // epilogue: renderer.py:359-367 (or :74-83 when nothing passed culling)
    const int64_t plane = (int64_t)img_w * img_h;
#pragma unroll
    for (int k = 0; k < kPx; ++k) {
        const int px = px0 + k;
        if (px < img_w && py < img_h) {
            const int64_t p = (int64_t)py * img_w + px;
            const float Ak = (k & 1) ? A[k >> 1].y : A[k >> 1].x;
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
    if (lane == 0) tile_consumed[tile] = walked;
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
constexpr int kRedVals = 10;     // mx my | q00 q01 q11 | opacity | z | r g b
constexpr int kRedStride = 36;   // floats per value row: 32 lanes + pad, keeps LDS.128 aligned

__global__ void __launch_bounds__(32, 16)
raster_bwd_kernel(int img_w, int img_h, int tiles_x, const int32_t* __restrict__ entry_ids,
                  const int2* __restrict__ tile_ranges, const float4* __restrict__ rec,
                  const float* __restrict__ bg_ptr, const float* __restrict__ alpha,
                  const float4* __restrict__ pix_state, const int32_t* __restrict__ tile_consumed,
                  const float* __restrict__ g_image, const float* __restrict__ g_alpha,
                  const float* __restrict__ g_depth,
                  float* __restrict__ g_means2d, float* __restrict__ g_conics, float* __restrict__ g_depths,
                  float* __restrict__ g_colors, float* __restrict__ g_opac) {
    __shared__ float4 srec[kBatch * 3];
    __shared__ int sid[kBatch];
    __shared__ __align__(16) float red[kRedVals * kRedStride];

    const int tile = blockIdx.x;
    const int lane = threadIdx.x;
    const int tx = tile % tiles_x, ty = tile / tiles_x;
    const int py = ty * kTile + (lane >> 1);
    const int px0 = tx * kTile + (lane & 1) * kPx;
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
        default: out_base = g_colors + 2; out_stride = 3; break;
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
            const bool inside = px < img_w && py < img_h;
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
                const float gd = g_depth[q];
                const float den = add_rn(Af, 1e-6f);
                gd_[h] = gd / den;
                ga_[h] = g_alpha[q] - (gr[h] * bg0 + gg[h] * bg1 + gb[h] * bg2) - gd * st.w / (den * den);
                Ti[h] = gr[h] * (st.x - bg0) + gg[h] * (st.y - bg1) + gb[h] * (st.z - bg2) + gd_[h] * st.w + ga_[h] * Af;
            }
        }
        fpx[p] = make_float2((float)(px0 + 2 * p), (float)(px0 + 2 * p + 1));
        A[p] = make_float2(Ai[0], Ai[1]);
        R[p] = make_float2(-Ti[0], -Ti[1]);
        gCr[p] = make_float2(gr[0], gr[1]); gCg[p] = make_float2(gg[0], gg[1]); gCb[p] = make_float2(gb[0], gb[1]);
        gDs[p] = make_float2(gd_[0], gd_[1]); gA[p] = make_float2(ga_[0], ga_[1]);
    }

    const int2 range = tile_ranges[tile];
    const int end = range.x + tile_consumed[tile];
    for (int base = range.x; base < end; base += kBatch) {
        bool alive = false;
#pragma unroll
        for (int p = 0; p < kPairs; ++p) alive |= (A[p].x < kTermA) | (A[p].y < kTermA);
        if (!__any_sync(0xffffffffu, alive)) break;
        const int cnt = min(kBatch, end - base);
        __syncwarp();
        if (lane < cnt) {
            const int id = entry_ids[base + lane];
            sid[lane] = id;
            float4 q1 = __ldg(&rec[(int64_t)id * 3 + 1]);
            q1.y = (q1.y > kTinyOpacity) ? q1.y : 0.f;
            srec[lane * 3 + 0] = __ldg(&rec[(int64_t)id * 3 + 0]);
            srec[lane * 3 + 1] = q1;
            srec[lane * 3 + 2] = __ldg(&rec[(int64_t)id * 3 + 2]);
        }
        __syncwarp();
        for (int j = 0; j < cnt; ++j) {
            const float4 r0 = srec[j * 3 + 0];
            const float4 r1 = srec[j * 3 + 1];
            const float2 r2 = *reinterpret_cast<const float2*>(&srec[j * 3 + 2]);
            const float op = r1.y;
            EntryRow row;
            float dy;
            load_entry_row(r0, r1, fpy, row, dy);
            const float2 cr = bc2(r1.w), cg = bc2(r2.x), cb = bc2(r2.y), z = bc2(r1.z);
            const float2 hscale = bc2(kLn2 * op);
            float2 s_h = bc2(0.f), s_x = bc2(0.f), s_xx = bc2(0.f), s_op = bc2(0.f), s_z = bc2(0.f);
            float2 s_cr = bc2(0.f), s_cg = bc2(0.f), s_cb = bc2(0.f);
#pragma unroll
            for (int p = 0; p < kPairs; ++p) {
                PairEval ev;
                eval_pair(fpx[p], row, A[p], ev);
                const float2 v = fma2(gCr[p], cr, fma2(gCg[p], cg, fma2(gCb[p], cb, fma2(gDs[p], z, gA[p]))));
                R[p] = fma2(ev.contrib, v, R[p]);
                A[p] = add2(A[p], ev.contrib);
                // suffix / (1 - a); the terminating contributor has an empty suffix.  Otherwise
                // a < 0.995, so 1 - a >= 0.005 and the approximate reciprocal is safe.
                const float2 oma = fma2(ev.a, bc2(-1.0f), bc2(1.0f));
                float2 nsuf = mul2(R[p], make_float2(rcp_approx(oma.x), rcp_approx(oma.y)));
                nsuf.x = (A[p].x >= kTermA) ? 0.f : nsuf.x;
                nsuf.y = (A[p].y >= kTermA) ? 0.f : nsuf.y;
                float2 g_a = fma2(ev.T, v, nsuf);
                // a = clamp(op*w, 0, 1), w = clamp(exp(-s/2), 0, 1): closed-interval pass-through
                g_a.x = (ev.act0 && ev.u.x <= 1.f) ? g_a.x : 0.f;
                g_a.y = (ev.act1 && ev.u.y <= 1.f) ? g_a.y : 0.f;
                s_op = fma2(g_a, ev.w, s_op);
                float2 h = mul2(mul2(ev.w, g_a), hscale);                      // dL/ds'
                h.x = (ev.e.x <= 1.f) ? h.x : 0.f;
                h.y = (ev.e.y <= 1.f) ? h.y : 0.f;
                const float2 hdx = mul2(h, ev.dx);
                s_h = add2(s_h, h);
                s_x = add2(s_x, hdx);
                s_xx = fma2(hdx, ev.dx, s_xx);
                s_cr = fma2(ev.contrib, gCr[p], s_cr);
                s_cg = fma2(ev.contrib, gCg[p], s_cg);
                s_cb = fma2(ev.contrib, gCb[p], s_cb);
                s_z = fma2(ev.contrib, gDs[p], s_z);
            }
            if (!(op > 0.f)) continue;      // staged as 0 when <= kTinyOpacity: the reference skips a <= 0 (warp-uniform)
            const float Sh = s_h.x + s_h.y, Sx = s_x.x + s_x.y, Sxx = s_xx.x + s_xx.y;
            const float dySh = dy * Sh;
            // transpose-reduce the 10 sums over the warp: red[v][lane]
            red[0 * kRedStride + lane] = -fmaf(2.f * r0.z, Sx, r0.w * dySh);             // g_mx
            red[1 * kRedStride + lane] = -fmaf(2.f * r1.x, dySh, r0.w * Sx);             // g_my
            red[2 * kRedStride + lane] = kNegHalfLog2e * Sxx;                            // g_Q00
            red[3 * kRedStride + lane] = kNegHalfLog2e * (dy * Sx);                      // g_Q01 = g_Q10
            red[4 * kRedStride + lane] = kNegHalfLog2e * (dy * dySh);                    // g_Q11
            red[5 * kRedStride + lane] = s_op.x + s_op.y;
            red[6 * kRedStride + lane] = s_z.x + s_z.y;
            red[7 * kRedStride + lane] = s_cr.x + s_cr.y;
            red[8 * kRedStride + lane] = s_cg.x + s_cg.y;
            red[9 * kRedStride + lane] = s_cb.x + s_cb.y;
            __syncwarp();
            float s = 0.f;
            if (lane < 30) {
                const float4 q0 = red_src[0], q1 = red_src[1];
                s = (q0.x + q0.y) + (q0.z + q0.w) + (q1.x + q1.y) + (q1.z + q1.w);
                if (red_g < 2) {
                    const float4 q2 = red_src[2];
                    s += (q2.x + q2.y) + (q2.z + q2.w);
                }
            }
            __syncwarp();
            // lane v < 10 collects the three thirds of value v (lanes v, v+10, v+20)
            const float s2 = __shfl_down_sync(0xffffffffu, s, 10);
            const float s3 = __shfl_down_sync(0xffffffffu, s, 20);
            if (lane < kRedVals) {
                const float total = s + s2 + s3;
                float* dst = out_base + (int64_t)sid[j] * out_stride;
                atomicAdd(dst, total);
                if (lane == 3) atomicAdd(dst + 1, total);      // Q01 and Q10 enter s symmetrically
            }
        }
    }
}

}  // namespace gs

using namespace gs;

static int check_raster_args(int32_t img_w, int32_t img_h, int32_t tile_size, const char* fn) {
    if (img_w <= 0 || img_h <= 0) {
        set_error("%s: image size must be positive", fn);
        return GS_ERR_INVALID_ARGUMENT;
    }
    if (tile_size != kTile) {
        set_error("%s: tile_size %d unsupported (kernels are built for %d)", fn, tile_size, kTile);
        return GS_ERR_UNSUPPORTED;
    }
    return GS_OK;
}

extern "C" int gs_raster_fwd(int32_t img_w, int32_t img_h, int32_t tile_size, const int32_t* entry_ids,
                             const int32_t* tile_ranges, const float* splat_rec, const float* bg,
                             int32_t any_visible_host, float* image, float* alpha, float* depth, float* pix_state,
                             int32_t* n_consumed, int32_t* tile_consumed, void* stream) {
    const int rc = check_raster_args(img_w, img_h, tile_size, "gs_raster_fwd");
    if (rc != GS_OK) return rc;
    GS_REQUIRE(tile_ranges && bg && image && alpha && depth && pix_state && tile_consumed, "NULL array argument");
    DeviceGuard guard(image);
    const int tiles_x = (img_w + kTile - 1) / kTile, tiles_y = (img_h + kTile - 1) / kTile;
    cudaStream_t st = (cudaStream_t)stream;
    if (n_consumed) {
        raster_fwd_kernel<true><<<tiles_x * tiles_y, 32, 0, st>>>(
            img_w, img_h, tiles_x, entry_ids, (const int2*)tile_ranges, (const float4*)splat_rec, bg, any_visible_host,
            image, alpha, depth, (float4*)pix_state, n_consumed, tile_consumed);
    } else {
        raster_fwd_kernel<false><<<tiles_x * tiles_y, 32, 0, st>>>(
            img_w, img_h, tiles_x, entry_ids, (const int2*)tile_ranges, (const float4*)splat_rec, bg, any_visible_host,
            image, alpha, depth, (float4*)pix_state, nullptr, tile_consumed);
    }
    GS_CUDA_TRY(cudaGetLastError());
    count_launches(1);
    return GS_OK;
}

extern "C" int gs_raster_bwd(int32_t img_w, int32_t img_h, int32_t tile_size, const int32_t* entry_ids,
                             const int32_t* tile_ranges, const float* splat_rec, const float* bg, const float* alpha,
                             const float* pix_state, const int32_t* tile_consumed, const float* g_image,
                             const float* g_alpha, const float* g_depth, float* g_means2d, float* g_conics,
                             float* g_depths, float* g_colors, float* g_opacities, void* stream) {
    const int rc = check_raster_args(img_w, img_h, tile_size, "gs_raster_bwd");
    if (rc != GS_OK) return rc;
    GS_REQUIRE(tile_ranges && bg && alpha && pix_state && tile_consumed && g_image && g_alpha && g_depth && g_means2d &&
                   g_conics && g_depths && g_colors && g_opacities, "NULL array argument");
    DeviceGuard guard(alpha);
    const int tiles_x = (img_w + kTile - 1) / kTile, tiles_y = (img_h + kTile - 1) / kTile;
    raster_bwd_kernel<<<tiles_x * tiles_y, 32, 0, (cudaStream_t)stream>>>(
        img_w, img_h, tiles_x, entry_ids, (const int2*)tile_ranges, (const float4*)splat_rec, bg, alpha,
        (const float4*)pix_state, tile_consumed, g_image, g_alpha, g_depth, g_means2d, g_conics, g_depths, g_colors,
        g_opacities);
    GS_CUDA_TRY(cudaGetLastError());
    count_launches(1);
    return GS_OK;
}
