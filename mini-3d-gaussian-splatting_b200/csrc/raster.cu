// Stage R: 16x16-tile front-to-back alpha compositing, forward and backward.
//
// Reference semantics: the pixel loop and epilogue of GaussianRenderer._tile_rasterization,
// src/core/renderer.py:300-367.
//
// Work decomposition (both directions): one CTA per tile, 64 threads, each thread owns a 1x4
// strip of pixels (same row, 4 consecutive columns).  Four pixels per thread amortise the
// shared-memory broadcast of a splat record over four evaluations, give the FP32 pipe four
// independent dependency chains, and -- in the backward pass -- let each thread pre-sum its own
// four pixels before the per-splat warp reduction, so the reduction + atomics cost is paid once
// per 128 pixel evaluations instead of once per 32.
//
// Splat records (48 B, written by project_fwd) are gathered by list entry into shared memory in
// batches; every thread then walks the batch.  The kernels are bounded by FP32/MUFU issue, not
// by HBM: a tile saturates after a few hundred of its thousands of list entries and
// the CTA leaves as soon as all of its pixels are done.
//
// The recurrence that decides skips and termination (s, w, a, contrib, A) is evaluated with the
// reference's exact fp32 operation order (no FMA contraction) and by the same inline function in
// forward and backward, so both walk bit-identical contributor sets.
#include "common.cuh"

namespace gs {

constexpr int kPx = 4;                 // pixels per thread
constexpr int kRasterThreads = 64;     // 16 rows x 4 strips
constexpr int kBatch = 64;             // list entries staged per round

struct Eval {
    float dx, dy, e, w, a, contrib;
};

// renderer.py:330-346.  Returns false where the reference `continue`s.
__device__ __forceinline__ bool eval_splat(float px, float py, const float4& r0, float q11, float op, float A, Eval& ev) {
    ev.dx = sub_rn(px, r0.x);
    ev.dy = sub_rn(py, r0.y);
    const float t1 = mul_rn(mul_rn(ev.dx, ev.dx), r0.z);
    const float t2 = mul_rn(mul_rn(r0.w, ev.dx), ev.dy);
    const float t3 = mul_rn(mul_rn(ev.dy, ev.dy), q11);
    const float s = add_rn(add_rn(t1, t2), t3);
    ev.e = expf(mul_rn(-0.5f, s));
    ev.w = fminf(fmaxf(ev.e, 0.f), 1.f);
    if (ev.w < 1e-5f) return false;
    ev.a = fminf(fmaxf(mul_rn(op, ev.w), 0.f), 1.f);
    if (ev.a <= 0.f) return false;
    ev.contrib = mul_rn(sub_rn(1.f, A), ev.a);
    if (ev.contrib <= 0.f) return false;
    return true;
}

__global__ void __launch_bounds__(kRasterThreads)
raster_fwd_kernel(int img_w, int img_h, int tiles_x, const int32_t* __restrict__ entry_ids,
                  const int2* __restrict__ tile_ranges, const float4* __restrict__ rec,
                  const float* __restrict__ bg_ptr, int any_visible,
                  float* __restrict__ image, float* __restrict__ alpha, float* __restrict__ depth,
                  float4* __restrict__ pix_state, int32_t* __restrict__ n_consumed,
                  int32_t* __restrict__ tile_consumed) {
    __shared__ float4 srec[kBatch * 3];
    __shared__ int s_max[kRasterThreads / 32];

    const int tile = blockIdx.x;
    const int tid = threadIdx.x;
    const int tx = tile % tiles_x, ty = tile / tiles_x;
    const int py = ty * kTile + (tid >> 2);
    const int px0 = tx * kTile + (tid & 3) * kPx;
    const float bg[3] = {bg_ptr[0], bg_ptr[1], bg_ptr[2]};
    const float fpy = (float)py;

    bool done[kPx];
    float A[kPx], Cr[kPx], Cg[kPx], Cb[kPx], Ds[kPx];
    int ncons[kPx];
#pragma unroll
    for (int k = 0; k < kPx; ++k) {
        done[k] = !(px0 + k < img_w && py < img_h);
        A[k] = 0.f; Ds[k] = 0.f;
        Cr[k] = bg[0]; Cg[k] = bg[1]; Cb[k] = bg[2];     // out_rgb starts at bg (renderer.py:273)
        ncons[k] = 0;
    }

    const int2 range = tile_ranges[tile];
    for (int base = range.x; base < range.y; base += kBatch) {
        const bool mine_done = done[0] && done[1] && done[2] && done[3];
        if (__syncthreads_and(mine_done)) break;          // also fences reuse of srec
        const int cnt = min(kBatch, range.y - base);
        if (tid < cnt) {
            const int id = entry_ids[base + tid];
            srec[tid * 3 + 0] = __ldg(&rec[(int64_t)id * 3 + 0]);
            srec[tid * 3 + 1] = __ldg(&rec[(int64_t)id * 3 + 1]);
            srec[tid * 3 + 2] = __ldg(&rec[(int64_t)id * 3 + 2]);
        }
        __syncthreads();
        for (int j = 0; j < cnt; ++j) {
            const float4 r0 = srec[j * 3 + 0];
            const float4 r1 = srec[j * 3 + 1];
            const float4 r2 = srec[j * 3 + 2];
            const int pos = base - range.x + j + 1;
#pragma unroll
            for (int k = 0; k < kPx; ++k) {
                if (!done[k]) {
                    ncons[k] = pos;
                    Eval ev;
                    if (eval_splat((float)(px0 + k), fpy, r0, r1.x, r1.y, A[k], ev)) {
                        Cr[k] = fmaf(ev.contrib, r1.w, Cr[k]);
                        Cg[k] = fmaf(ev.contrib, r2.x, Cg[k]);
                        Cb[k] = fmaf(ev.contrib, r2.y, Cb[k]);
                        Ds[k] = fmaf(ev.contrib, r1.z, Ds[k]);
                        A[k] = add_rn(A[k], ev.contrib);
                        if (A[k] >= 0.995f) done[k] = true;          // renderer.py:352
                    }
                }
            }
        }
    }

    // epilogue: renderer.py:359-367 (or :74-83 when nothing passed culling)
    int my_max = 0;
#pragma unroll
    for (int k = 0; k < kPx; ++k) {
        const int px = px0 + k;
        if (px < img_w && py < img_h) {
            const int64_t p = (int64_t)py * img_w + px;
            const int64_t plane = (int64_t)img_w * img_h;
            float o_r, o_g, o_b, o_a, o_d;
            if (any_visible) {
                const float om = sub_rn(1.f, A[k]);
                o_r = fminf(fmaxf(add_rn(Cr[k], mul_rn(om, bg[0])), 0.f), 1.f);
                o_g = fminf(fmaxf(add_rn(Cg[k], mul_rn(om, bg[1])), 0.f), 1.f);
                o_b = fminf(fmaxf(add_rn(Cb[k], mul_rn(om, bg[2])), 0.f), 1.f);
                o_a = fminf(fmaxf(A[k], 0.f), 1.f);
                o_d = div_rn(Ds[k], add_rn(A[k], 1e-6f));
            } else {
                o_r = bg[0]; o_g = bg[1]; o_b = bg[2]; o_a = 0.f; o_d = 0.f;
            }
            image[p] = o_r;
            image[plane + p] = o_g;
            image[2 * plane + p] = o_b;
            alpha[p] = o_a;
            depth[p] = o_d;
            pix_state[p] = make_float4(Cr[k], Cg[k], Cb[k], Ds[k]);
            if (n_consumed) n_consumed[p] = ncons[k];
            my_max = max(my_max, ncons[k]);
        }
    }
    // tile_consumed = max over the tile's pixels (bounds the backward walk; also the E statistic)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) my_max = max(my_max, __shfl_xor_sync(0xffffffffu, my_max, o));
    if ((tid & 31) == 0) s_max[tid >> 5] = my_max;
    __syncthreads();
    if (tid == 0) tile_consumed[tile] = max(s_max[0], s_max[1]);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Backward: re-walks the list front to back with the forward's recurrence.  With
//   v_k = gC.colour_k + gDs.z_k + gA   and   Total = sum_k T_k a_k v_k = gC.(C - bg) + gDs.Dsum + gA.A
// (all known from the forward's saved state), the gradient w.r.t. a_k is
//   T_k v_k - (Total - prefix_k) / (1 - a_k),
// so no reverse traversal and no division-recovered transmittance is needed.  Every non-final
// contributor has a_k < 0.995 (else the pixel would have terminated there); for the terminating
// contributor the suffix is exactly zero.
__global__ void __launch_bounds__(kRasterThreads)
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

    const int tile = blockIdx.x;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int tx = tile % tiles_x, ty = tile / tiles_x;
    const int py = ty * kTile + (tid >> 2);
    const int px0 = tx * kTile + (tid & 3) * kPx;
    const float bg[3] = {bg_ptr[0], bg_ptr[1], bg_ptr[2]};
    const float fpy = (float)py;
    const int64_t plane = (int64_t)img_w * img_h;

    bool done[kPx];
    float A[kPx], P[kPx], Total[kPx], gCr[kPx], gCg[kPx], gCb[kPx], gDs[kPx], gA[kPx];
#pragma unroll
    for (int k = 0; k < kPx; ++k) {
        const int px = px0 + k;
        const bool inside = px < img_w && py < img_h;
        done[k] = !inside;
        A[k] = 0.f; P[k] = 0.f;
        Total[k] = gCr[k] = gCg[k] = gCb[k] = gDs[k] = gA[k] = 0.f;
        if (inside) {
            const int64_t p = (int64_t)py * img_w + px;
            const float Af = alpha[p];                   // A is always inside [0,1], so alpha == A
            const float4 st = pix_state[p];
            const float om = sub_rn(1.f, Af);
            const float pre_r = add_rn(st.x, mul_rn(om, bg[0]));
            const float pre_g = add_rn(st.y, mul_rn(om, bg[1]));
            const float pre_b = add_rn(st.z, mul_rn(om, bg[2]));
            // torch.clamp passes the gradient on the closed interval
            gCr[k] = (pre_r >= 0.f && pre_r <= 1.f) ? g_image[p] : 0.f;
            gCg[k] = (pre_g >= 0.f && pre_g <= 1.f) ? g_image[plane + p] : 0.f;
            gCb[k] = (pre_b >= 0.f && pre_b <= 1.f) ? g_image[2 * plane + p] : 0.f;
            const float gd = g_depth[p];
            const float den = add_rn(Af, 1e-6f);
            gDs[k] = gd / den;
            gA[k] = g_alpha[p] - (gCr[k] * bg[0] + gCg[k] * bg[1] + gCb[k] * bg[2]) - gd * st.w / (den * den);
            Total[k] = gCr[k] * (st.x - bg[0]) + gCg[k] * (st.y - bg[1]) + gCb[k] * (st.z - bg[2]) + gDs[k] * st.w +
                       gA[k] * Af;
        }
    }

    const int2 range = tile_ranges[tile];
    const int end = range.x + tile_consumed[tile];
    for (int base = range.x; base < end; base += kBatch) {
        const bool mine_done = done[0] && done[1] && done[2] && done[3];
        if (__syncthreads_and(mine_done)) break;
        const int cnt = min(kBatch, end - base);
        if (tid < cnt) {
            const int id = entry_ids[base + tid];
            sid[tid] = id;
            srec[tid * 3 + 0] = __ldg(&rec[(int64_t)id * 3 + 0]);
            srec[tid * 3 + 1] = __ldg(&rec[(int64_t)id * 3 + 1]);
            srec[tid * 3 + 2] = __ldg(&rec[(int64_t)id * 3 + 2]);
        }
        __syncthreads();
        for (int j = 0; j < cnt; ++j) {
            const float4 r0 = srec[j * 3 + 0];
            const float4 r1 = srec[j * 3 + 1];
            const float4 r2 = srec[j * 3 + 2];
            const float q00 = r0.z, qs = r0.w, q11 = r1.x, op = r1.y, z = r1.z;
            float a_mx = 0.f, a_my = 0.f, a_q00 = 0.f, a_q01 = 0.f, a_q11 = 0.f, a_op = 0.f, a_z = 0.f;
            float a_cr = 0.f, a_cg = 0.f, a_cb = 0.f;
            bool any = false;
#pragma unroll
            for (int k = 0; k < kPx; ++k) {
                if (!done[k]) {
                    Eval ev;
                    if (eval_splat((float)(px0 + k), fpy, r0, q11, op, A[k], ev)) {
                        any = true;
                        const float T = sub_rn(1.f, A[k]);
                        const float v = gCr[k] * r1.w + gCg[k] * r2.x + gCb[k] * r2.y + gDs[k] * z + gA[k];
                        P[k] = fmaf(ev.contrib, v, P[k]);
                        A[k] = add_rn(A[k], ev.contrib);
                        float g_a = T * v;
                        if (A[k] >= 0.995f) {
                            done[k] = true;                     // terminating contributor: empty suffix
                        } else {
                            g_a -= (Total[k] - P[k]) / (1.f - ev.a);
                        }
                        a_cr = fmaf(ev.contrib, gCr[k], a_cr);
                        a_cg = fmaf(ev.contrib, gCg[k], a_cg);
                        a_cb = fmaf(ev.contrib, gCb[k], a_cb);
                        a_z = fmaf(ev.contrib, gDs[k], a_z);
                        // a = clamp(op*w, 0, 1), w = clamp(exp(-s/2), 0, 1): closed-interval pass-through
                        const float u = mul_rn(op, ev.w);
                        if (u >= 0.f && u <= 1.f) {
                            a_op = fmaf(g_a, ev.w, a_op);
                            if (ev.e <= 1.f) {
                                const float g_s = -0.5f * ev.w * (g_a * op);
                                a_q00 = fmaf(ev.dx * ev.dx, g_s, a_q00);
                                a_q01 = fmaf(ev.dx * ev.dy, g_s, a_q01);
                                a_q11 = fmaf(ev.dy * ev.dy, g_s, a_q11);
                                a_mx -= g_s * (2.f * ev.dx * q00 + qs * ev.dy);
                                a_my -= g_s * (2.f * ev.dy * q11 + qs * ev.dx);
                            }
                        }
                    }
                }
            }
            if (__ballot_sync(0xffffffffu, any) == 0u) continue;
            a_mx = warp_sum(a_mx); a_my = warp_sum(a_my);
            a_q00 = warp_sum(a_q00); a_q01 = warp_sum(a_q01); a_q11 = warp_sum(a_q11);
            a_op = warp_sum(a_op); a_z = warp_sum(a_z);
            a_cr = warp_sum(a_cr); a_cg = warp_sum(a_cg); a_cb = warp_sum(a_cb);
            if (lane == 0) {
                const int64_t id = sid[j];
                atomicAdd(&g_means2d[id * 2 + 0], a_mx);
                atomicAdd(&g_means2d[id * 2 + 1], a_my);
                atomicAdd(&g_conics[id * 4 + 0], a_q00);
                atomicAdd(&g_conics[id * 4 + 1], a_q01);     // Q01 and Q10 enter s symmetrically
                atomicAdd(&g_conics[id * 4 + 2], a_q01);
                atomicAdd(&g_conics[id * 4 + 3], a_q11);
                atomicAdd(&g_depths[id], a_z);
                atomicAdd(&g_colors[id * 3 + 0], a_cr);
                atomicAdd(&g_colors[id * 3 + 1], a_cg);
                atomicAdd(&g_colors[id * 3 + 2], a_cb);
                atomicAdd(&g_opac[id], a_op);
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
    raster_fwd_kernel<<<tiles_x * tiles_y, kRasterThreads, 0, (cudaStream_t)stream>>>(
        img_w, img_h, tiles_x, entry_ids, (const int2*)tile_ranges, (const float4*)splat_rec, bg, any_visible_host,
        image, alpha, depth, (float4*)pix_state, n_consumed, tile_consumed);
    GS_CUDA_TRY(cudaGetLastError());
    count_launches(1);
    return GS_OK;
}

extern "C" int gs_raster_bwd(int32_t img_w, int32_t img_h, int32_t tile_size, const int32_t* entry_ids,
                             const int32_t* tile_ranges, const float* splat_rec, const float* bg, const float* alpha,
                             const float* pix_state, const int32_t* n_consumed, const int32_t* tile_consumed,
                             const float* g_image, const float* g_alpha, const float* g_depth, float* g_means2d,
                             float* g_conics, float* g_depths, float* g_colors, float* g_opacities, void* stream) {
    (void)n_consumed;
    const int rc = check_raster_args(img_w, img_h, tile_size, "gs_raster_bwd");
    if (rc != GS_OK) return rc;
    GS_REQUIRE(tile_ranges && bg && alpha && pix_state && tile_consumed && g_image && g_alpha && g_depth && g_means2d &&
                   g_conics && g_depths && g_colors && g_opacities, "NULL array argument");
    DeviceGuard guard(alpha);
    const int tiles_x = (img_w + kTile - 1) / kTile, tiles_y = (img_h + kTile - 1) / kTile;
    raster_bwd_kernel<<<tiles_x * tiles_y, kRasterThreads, 0, (cudaStream_t)stream>>>(
        img_w, img_h, tiles_x, entry_ids, (const int2*)tile_ranges, (const float4*)splat_rec, bg, alpha,
        (const float4*)pix_state, tile_consumed, g_image, g_alpha, g_depth, g_means2d, g_conics, g_depths, g_colors,
        g_opacities);
    GS_CUDA_TRY(cudaGetLastError());
    count_launches(1);
    return GS_OK;
}
