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

constexpr int kPx = 8;                 // pixels per lane (1x8 strip)
constexpr int kBatch = 32;             // list entries staged per round (one per lane)
constexpr float kTermA = 0.995f;       // renderer.py:352
constexpr float kMinW = 1e-5f;         // renderer.py:336
constexpr float kLn2 = 0.69314718056f;
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

// Per-entry values shared by a lane's 8 pixels.
struct EntryRow {
    float mx, q00, qsdy, q11dy2, op;
};

// renderer.py:330-346 for one pixel.  `A >= kTermA` encodes "this pixel has terminated (or lies
// outside the image)", so no separate flag is carried.  Returns the predicate under which the
// reference accumulates; w, a, contrib are always computed (branch-free).
__device__ __forceinline__ bool eval_pixel(float px, const EntryRow& r, float A, float& dx, float& e, float& w,
                                           float& a, float& contrib) {
    dx = px - r.mx;
    const float t = fmaf(r.q00, dx, r.qsdy);
    const float sp = fmaf(dx, t, r.q11dy2);            // -0.5*log2(e) * s
    e = ex2_approx(sp);
    w = fminf(e, 1.0f);                                // clamp(exp(.), 0, 1); exp >= 0
    a = __saturatef(mul_rn(r.op, w));                  // clamp(opacity * w, 0, 1)
    contrib = mul_rn(sub_rn(1.0f, A), a);
    // a > 0 and contrib > 0 follow from opacity > kTinyOpacity, w >= 1e-5 and 1 - A >= 0.005
    return (A < kTermA) && (w >= kMinW);
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
    const float fpy = (float)py, fpx0 = (float)px0;

    float A[kPx], Cr[kPx], Cg[kPx], Cb[kPx], Ds[kPx];
    int ncons[kPx];
#pragma unroll
    for (int k = 0; k < kPx; ++k) {
        const bool inside = (px0 + k < img_w) && (py < img_h);
        A[k] = inside ? 0.f : 2.0f;                     // 2.0: never alive
        Ds[k] = 0.f;
        Cr[k] = bg0; Cg[k] = bg1; Cb[k] = bg2;          // out_rgb starts at bg (renderer.py:273)
        ncons[k] = -1;
    }

    const int2 range = tile_ranges[tile];
    int walked = 0;
    for (int base = range.x; base < range.y; base += kBatch) {
        bool alive = false;
#pragma unroll
        for (int k = 0; k < kPx; ++k) alive |= (A[k] < kTermA);
        if (!__any_sync(0xffffffffu, alive)) break;
        const int cnt = min(kBatch, range.y - base);
        __syncwarp();                                   // previous batch fully read
        if (lane < cnt) {
            const int id = entry_ids[base + lane];
            srec[lane * 3 + 0] = __ldg(&rec[(int64_t)id * 3 + 0]);
            srec[lane * 3 + 1] = __ldg(&rec[(int64_t)id * 3 + 1]);
            srec[lane * 3 + 2] = __ldg(&rec[(int64_t)id * 3 + 2]);
        }
        __syncwarp();
        walked = base - range.x + cnt;
        for (int j = 0; j < cnt; ++j) {
            const float4 r0 = srec[j * 3 + 0];          // mx, my, q00', qs'
            const float4 r1 = srec[j * 3 + 1];          // q11', opacity, z, r
            const float2 r2 = *reinterpret_cast<const float2*>(&srec[j * 3 + 2]);   // g, b
            if (!(r1.y > kTinyOpacity)) continue;       // warp-uniform
            const float dy = fpy - r0.y;
            EntryRow row;
            row.mx = r0.x; row.q00 = r0.z; row.qsdy = r0.w * dy; row.q11dy2 = r1.x * dy * dy; row.op = r1.y;
#pragma unroll
            for (int k = 0; k < kPx; ++k) {
                float dx, e, w, a, contrib;
                if (eval_pixel(fpx0 + (float)k, row, A[k], dx, e, w, a, contrib)) {
                    Cr[k] = fmaf(contrib, r1.w, Cr[k]);
                    Cg[k] = fmaf(contrib, r2.x, Cg[k]);
                    Cb[k] = fmaf(contrib, r2.y, Cb[k]);
                    Ds[k] = fmaf(contrib, r1.z, Ds[k]);
                    A[k] = add_rn(A[k], contrib);
                    if (kTrack && A[k] >= kTermA) ncons[k] = base - range.x + j + 1;   // renderer.py:352
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
            float o_r, o_g, o_b, o_a, o_d;
            if (any_visible) {
                const float om = sub_rn(1.f, A[k]);
                o_r = __saturatef(add_rn(Cr[k], mul_rn(om, bg0)));
                o_g = __saturatef(add_rn(Cg[k], mul_rn(om, bg1)));
                o_b = __saturatef(add_rn(Cb[k], mul_rn(om, bg2)));
                o_a = __saturatef(A[k]);
                o_d = div_rn(Ds[k], add_rn(A[k], 1e-6f));
            } else {
                o_r = bg0; o_g = bg1; o_b = bg2; o_a = 0.f; o_d = 0.f;
            }
            image[p] = o_r;
            image[plane + p] = o_g;
            image[2 * plane + p] = o_b;
            alpha[p] = o_a;
            depth[p] = o_d;
            pix_state[p] = make_float4(Cr[k], Cg[k], Cb[k], Ds[k]);
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
// With s' = c*s (c = -0.5*log2 e) and w = 2^s':  dL/ds' = ln2 * w * dL/dw; the conic sums are
// multiplied by c once per entry after the warp reduction.
constexpr int kRedVals = 10;     // mx my | q00 q01 q11 | opacity | z | r g b

__global__ void __launch_bounds__(32)
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
    __shared__ float red[kRedVals][33];

    const int tile = blockIdx.x;
    const int lane = threadIdx.x;
    const int tx = tile % tiles_x, ty = tile / tiles_x;
    const int py = ty * kTile + (lane >> 1);
    const int px0 = tx * kTile + (lane & 1) * kPx;
    const float bg0 = bg_ptr[0], bg1 = bg_ptr[1], bg2 = bg_ptr[2];
    const float fpy = (float)py, fpx0 = (float)px0;
    const int64_t plane = (int64_t)img_w * img_h;

    // where lane v < 10 sends reduced value v:  target = out_base + id * out_stride
    float* out_base;
    int out_stride;
    float out_scale = 1.0f;
    const float kC = -0.72134752044448170f;            // -0.5 * log2(e)
    switch (lane) {
        case 0: out_base = g_means2d; out_stride = 2; break;
        case 1: out_base = g_means2d + 1; out_stride = 2; break;
        case 2: out_base = g_conics; out_stride = 4; out_scale = kC; break;
        case 3: out_base = g_conics + 1; out_stride = 4; out_scale = kC; break;   // Q01 (and Q10 below)
        case 4: out_base = g_conics + 3; out_stride = 4; out_scale = kC; break;
        case 5: out_base = g_opac; out_stride = 1; break;
        case 6: out_base = g_depths; out_stride = 1; break;
        case 7: out_base = g_colors; out_stride = 3; break;
        case 8: out_base = g_colors + 1; out_stride = 3; break;
        default: out_base = g_colors + 2; out_stride = 3; break;
    }
    const int red_v = lane % kRedVals, red_g = lane / kRedVals;      // lanes 0..29: value, third

    float A[kPx], P[kPx], Total[kPx], gCr[kPx], gCg[kPx], gCb[kPx], gDs[kPx], gA[kPx];
#pragma unroll
    for (int k = 0; k < kPx; ++k) {
        const int px = px0 + k;
        const bool inside = px < img_w && py < img_h;
        A[k] = inside ? 0.f : 2.0f;
        P[k] = 0.f;
        Total[k] = gCr[k] = gCg[k] = gCb[k] = gDs[k] = gA[k] = 0.f;
        if (inside) {
            const int64_t p = (int64_t)py * img_w + px;
            const float Af = alpha[p];                   // A is always inside [0,1], so alpha == A
            const float4 st = pix_state[p];
            const float om = sub_rn(1.f, Af);
            const float pre_r = add_rn(st.x, mul_rn(om, bg0));
            const float pre_g = add_rn(st.y, mul_rn(om, bg1));
            const float pre_b = add_rn(st.z, mul_rn(om, bg2));
            // torch.clamp passes the gradient on the closed interval
            gCr[k] = (pre_r >= 0.f && pre_r <= 1.f) ? g_image[p] : 0.f;
            gCg[k] = (pre_g >= 0.f && pre_g <= 1.f) ? g_image[plane + p] : 0.f;
            gCb[k] = (pre_b >= 0.f && pre_b <= 1.f) ? g_image[2 * plane + p] : 0.f;
            const float gd = g_depth[p];
            const float den = add_rn(Af, 1e-6f);
            gDs[k] = gd / den;
            gA[k] = g_alpha[p] - (gCr[k] * bg0 + gCg[k] * bg1 + gCb[k] * bg2) - gd * st.w / (den * den);
            Total[k] = gCr[k] * (st.x - bg0) + gCg[k] * (st.y - bg1) + gCb[k] * (st.z - bg2) + gDs[k] * st.w +
                       gA[k] * Af;
        }
    }

    const int2 range = tile_ranges[tile];
    const int end = range.x + tile_consumed[tile];
    for (int base = range.x; base < end; base += kBatch) {
        bool alive = false;
#pragma unroll
        for (int k = 0; k < kPx; ++k) alive |= (A[k] < kTermA);
        if (!__any_sync(0xffffffffu, alive)) break;
        const int cnt = min(kBatch, end - base);
        __syncwarp();
        if (lane < cnt) {
            const int id = entry_ids[base + lane];
            sid[lane] = id;
            srec[lane * 3 + 0] = __ldg(&rec[(int64_t)id * 3 + 0]);
            srec[lane * 3 + 1] = __ldg(&rec[(int64_t)id * 3 + 1]);
            srec[lane * 3 + 2] = __ldg(&rec[(int64_t)id * 3 + 2]);
        }
        __syncwarp();
        for (int j = 0; j < cnt; ++j) {
            const float4 r0 = srec[j * 3 + 0];
            const float4 r1 = srec[j * 3 + 1];
            const float2 r2 = *reinterpret_cast<const float2*>(&srec[j * 3 + 2]);
            const float op = r1.y, z = r1.z;
            if (!(op > kTinyOpacity)) continue;         // warp-uniform
            const float dy = fpy - r0.y;
            EntryRow row;
            row.mx = r0.x; row.q00 = r0.z; row.qsdy = r0.w * dy; row.q11dy2 = r1.x * dy * dy; row.op = op;
            const float two_q00 = 2.f * r0.z, two_q11dy = 2.f * r1.x * dy, dy2 = dy * dy;
            float a_mx = 0.f, a_my = 0.f, a_q00 = 0.f, a_q01 = 0.f, a_q11 = 0.f, a_op = 0.f, a_z = 0.f;
            float a_cr = 0.f, a_cg = 0.f, a_cb = 0.f;
            bool any = false;
#pragma unroll
            for (int k = 0; k < kPx; ++k) {
                float dx, e, w, a, contrib;
                const float T = sub_rn(1.f, A[k]);
                if (eval_pixel(fpx0 + (float)k, row, A[k], dx, e, w, a, contrib)) {
                    any = true;
                    const float v = fmaf(gCr[k], r1.w, fmaf(gCg[k], r2.x, fmaf(gCb[k], r2.y, fmaf(gDs[k], z, gA[k]))));
                    P[k] = fmaf(contrib, v, P[k]);
                    A[k] = add_rn(A[k], contrib);
                    // terminating contributor: empty suffix.  Otherwise a < 0.995, so 1 - a >= 0.005.
                    const float suffix = (A[k] >= kTermA) ? 0.f : (Total[k] - P[k]) * rcp_approx(1.f - a);
                    const float g_a = fmaf(T, v, -suffix);
                    a_cr = fmaf(contrib, gCr[k], a_cr);
                    a_cg = fmaf(contrib, gCg[k], a_cg);
                    a_cb = fmaf(contrib, gCb[k], a_cb);
                    a_z = fmaf(contrib, gDs[k], a_z);
                    // a = clamp(op*w, 0, 1), w = clamp(exp(-s/2), 0, 1): closed-interval pass-through
                    const bool pass_a = mul_rn(op, w) <= 1.f;
                    const float g_aw = pass_a ? g_a : 0.f;
                    a_op = fmaf(g_aw, w, a_op);
                    const float h = (e <= 1.f) ? (kLn2 * w) * (g_aw * op) : 0.f;     // dL/ds'
                    a_q00 = fmaf(dx * dx, h, a_q00);
                    a_q01 = fmaf(dx * dy, h, a_q01);
                    a_q11 = fmaf(dy2, h, a_q11);
                    a_mx = fmaf(-h, fmaf(two_q00, dx, r0.w * dy), a_mx);
                    a_my = fmaf(-h, fmaf(r0.w, dx, two_q11dy), a_my);
                }
            }
            if (!__any_sync(0xffffffffu, any)) continue;
            // transpose-reduce the 10 sums over the warp: red[v][lane], row stride 33 (conflict-free)
            red[0][lane] = a_mx;  red[1][lane] = a_my;
            red[2][lane] = a_q00; red[3][lane] = a_q01; red[4][lane] = a_q11;
            red[5][lane] = a_op;  red[6][lane] = a_z;
            red[7][lane] = a_cr;  red[8][lane] = a_cg;  red[9][lane] = a_cb;
            __syncwarp();
            float s = 0.f;
            if (lane < 30) {
                const float* rowp = &red[red_v][red_g * 11];
#pragma unroll
                for (int t = 0; t < 10; ++t) s += rowp[t];
                if (red_g < 2) s += rowp[10];            // thirds cover 11 + 11 + 10 lanes
            }
            __syncwarp();
            // lane v < 10 collects the three thirds of value v (lanes v, v+10, v+20)
            const float s2 = __shfl_down_sync(0xffffffffu, s, 10);
            const float s3 = __shfl_down_sync(0xffffffffu, s, 20);
            if (lane < kRedVals) {
                const float total = (s + s2 + s3) * out_scale;
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
