// Loss heads fused into one pass over the rendered planes (the train step).
//
// The reference's step computes L1(image, target) with torch ops (src/core/optimizer.py:137-139, src/utils/loss.py
// l1_loss = |a - b|.mean()) and lets autograd produce sign(a - b) / n: five elementwise / reduction kernels and
// three passes over the image each way.  Here one launch reads the rendered plane and its target once, writes the
// gradient plane and reduces the loss; the same skeleton serves the weighted-sum loss of the parity tests and the
// benchmark (sum(w_img*image) + sum(w_a*alpha) + 0.1*sum(w_d*depth)), whose gradient is the weights.
//
// Deterministic: every block reduces a fixed slice in a fixed order into one double, the last block to finish (ticket
// counter in the workspace) adds the partials in block order.  HBM-bound: 8 B (L1 without gradient), 12 B (L1 with
// gradient) or 8 B (weighted sum) per element.
#include "common.cuh"

namespace gs {

constexpr int kLossThreads = 256;
constexpr int kLossMaxBlocks = 1184;             // 148 SMs x 8 resident CTAs
constexpr int kLossMaxTerms = 4;

struct LossWorkspace {                           // caller-provided, zero-initialised once
    double partial[kLossMaxBlocks];
    unsigned int ticket;
};

struct WeightedTerms {
    const float* x[kLossMaxTerms];
    const float* w[kLossMaxTerms];
    long long n[kLossMaxTerms];
    float coeff[kLossMaxTerms];
    int count;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block sum (fixed order) -> partial[block]; the last block adds the partials in index order and resets the ticket
__device__ __forceinline__ void finish_loss(float thread_sum, float scale, LossWorkspace* ws, float* out) {
    __shared__ float s_warp[kLossThreads / 32];
    __shared__ bool s_last;
    const float wsum = warp_sum(thread_sum);
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = wsum;
    __syncthreads();
    if (threadIdx.x == 0) {
        double b = 0.0;
#pragma unroll
        for (int i = 0; i < kLossThreads / 32; ++i) b += (double)s_warp[i];
        ws->partial[blockIdx.x] = b;
        __threadfence();
        s_last = atomicAdd(&ws->ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (s_last && threadIdx.x < 32) {
        __threadfence();
        double t = 0.0;
        for (int i = threadIdx.x; i < (int)gridDim.x; i += 32) t += ((volatile double*)ws->partial)[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (threadIdx.x == 0) {
            *out = (float)(t * (double)scale);
            ws->ticket = 0;
        }
    }
}

__global__ void __launch_bounds__(kLossThreads)
weighted_sum_kernel(WeightedTerms terms, LossWorkspace* ws, float* out) {
    grid_dependency_wait();                          // launched early (PDL) behind the compositing pass
    float acc = 0.f;
    for (int k = 0; k < terms.count; ++k) {
        const float* __restrict__ x = terms.x[k];
        const float* __restrict__ w = terms.w[k];
        const long long n4 = terms.n[k] >> 2;
        float a = 0.f;
        const bool vec = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w)) & 15) == 0;
        long long done = 0;
        if (vec) {
            // two independent 16-byte loads of each array in flight per thread and iteration
            const long long stride = (long long)gridDim.x * kLossThreads;
            float a2 = 0.f;
            long long i = (long long)blockIdx.x * kLossThreads + threadIdx.x;
            for (; i + stride < n4; i += 2 * stride) {
                const float4 x0 = __ldg(reinterpret_cast<const float4*>(x) + i), x1 = __ldg(reinterpret_cast<const float4*>(x) + i + stride);
                const float4 w0 = __ldg(reinterpret_cast<const float4*>(w) + i), w1 = __ldg(reinterpret_cast<const float4*>(w) + i + stride);
                a = fmaf(x0.x, w0.x, a); a = fmaf(x0.y, w0.y, a); a = fmaf(x0.z, w0.z, a); a = fmaf(x0.w, w0.w, a);
                a2 = fmaf(x1.x, w1.x, a2); a2 = fmaf(x1.y, w1.y, a2); a2 = fmaf(x1.z, w1.z, a2); a2 = fmaf(x1.w, w1.w, a2);
            }
            if (i < n4) {
                const float4 xv = __ldg(reinterpret_cast<const float4*>(x) + i);
                const float4 wv = __ldg(reinterpret_cast<const float4*>(w) + i);
                a = fmaf(xv.x, wv.x, a); a = fmaf(xv.y, wv.y, a); a = fmaf(xv.z, wv.z, a); a = fmaf(xv.w, wv.w, a);
            }
            a += a2;
            done = n4 << 2;
        }
        for (long long i = done + (long long)blockIdx.x * kLossThreads + threadIdx.x; i < terms.n[k]; i += (long long)gridDim.x * kLossThreads)
            a = fmaf(__ldg(x + i), __ldg(w + i), a);
        acc = fmaf(terms.coeff[k], a, acc);
    }
    finish_loss(acc, 1.0f, ws, out);
}

// loss = mean |x - t|;  grad (optional) = sign(x - t) * grad_scale / n   (torch: sign(0) = 0)
__global__ void __launch_bounds__(kLossThreads)
l1_loss_kernel(const float* __restrict__ x, const float* __restrict__ t, long long n, float grad_scale,
               float* __restrict__ grad, LossWorkspace* ws, float* out) {
    grid_dependency_wait();
    float acc = 0.f;
    const float g = grad_scale / (float)n;
    const bool vec = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(t) | reinterpret_cast<uintptr_t>(grad)) & 15) == 0;
    const long long n4 = vec ? (n >> 2) : 0;
    auto sgn = [g](float d) { return d > 0.f ? g : (d < 0.f ? -g : 0.f); };
    for (long long i = (long long)blockIdx.x * kLossThreads + threadIdx.x; i < n4; i += (long long)gridDim.x * kLossThreads) {
        const float4 xv = __ldg(reinterpret_cast<const float4*>(x) + i);
        const float4 tv = __ldg(reinterpret_cast<const float4*>(t) + i);
        const float4 d = make_float4(xv.x - tv.x, xv.y - tv.y, xv.z - tv.z, xv.w - tv.w);
        acc += fabsf(d.x) + fabsf(d.y) + fabsf(d.z) + fabsf(d.w);
        if (grad) reinterpret_cast<float4*>(grad)[i] = make_float4(sgn(d.x), sgn(d.y), sgn(d.z), sgn(d.w));
    }
    for (long long i = (n4 << 2) + (long long)blockIdx.x * kLossThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kLossThreads) {
        const float d = __ldg(x + i) - __ldg(t + i);
        acc += fabsf(d);
        if (grad) grad[i] = sgn(d);
    }
    finish_loss(acc, 1.0f / (float)n, ws, out);
}

static int loss_grid(long long n) {
    const long long want = (n / 4 + kLossThreads - 1) / kLossThreads;
    return (int)(want < 1 ? 1 : (want > kLossMaxBlocks ? kLossMaxBlocks : want));
}

}  // namespace gs

using namespace gs;

extern "C" int64_t gs_loss_workspace_bytes(void) { return (int64_t)sizeof(LossWorkspace); }

extern "C" int gs_weighted_sum(int32_t num_terms, const float* const* x, const float* const* w, const int64_t* n,
                               const float* coeff, float* out, void* workspace, int64_t workspace_bytes, void* stream) {
    GS_REQUIRE(num_terms >= 1 && num_terms <= kLossMaxTerms, "1..4 terms");
    GS_REQUIRE(x && w && n && coeff && out && workspace, "NULL argument");
    if (workspace_bytes < (int64_t)sizeof(LossWorkspace)) {
        set_error("gs_weighted_sum: workspace of %lld bytes, need %lld", (long long)workspace_bytes, (long long)sizeof(LossWorkspace));
        return GS_ERR_WORKSPACE_TOO_SMALL;
    }
    WeightedTerms terms;
    terms.count = num_terms;
    long long total = 0;
    for (int k = 0; k < kLossMaxTerms; ++k) {
        const bool on = k < num_terms;
        GS_REQUIRE(!on || (x[k] && w[k] && n[k] >= 0), "NULL term");
        terms.x[k] = on ? x[k] : nullptr;
        terms.w[k] = on ? w[k] : nullptr;
        terms.n[k] = on ? n[k] : 0;
        terms.coeff[k] = on ? coeff[k] : 0.f;
        total += terms.n[k];
    }
    DeviceGuard guard(out);
    GS_CUDA_TRY(launch_pdl(weighted_sum_kernel, dim3(loss_grid(total)), dim3(kLossThreads), 0, (cudaStream_t)stream,
                           terms, (LossWorkspace*)workspace, out));
    count_launches(1);
    return GS_OK;
}

extern "C" int gs_l1_loss(const float* x, const float* target, int64_t n, float grad_scale, float* grad, float* out,
                          void* workspace, int64_t workspace_bytes, void* stream) {
    GS_REQUIRE(x && target && out && workspace && n > 0, "bad arguments");
    if (workspace_bytes < (int64_t)sizeof(LossWorkspace)) {
        set_error("gs_l1_loss: workspace of %lld bytes, need %lld", (long long)workspace_bytes, (long long)sizeof(LossWorkspace));
        return GS_ERR_WORKSPACE_TOO_SMALL;
    }
    DeviceGuard guard(out);
    GS_CUDA_TRY(launch_pdl(l1_loss_kernel, dim3(loss_grid(n)), dim3(kLossThreads), 0, (cudaStream_t)stream,
                           x, target, (long long)n, grad_scale, grad, (LossWorkspace*)workspace, out));
    count_launches(1);
    return GS_OK;
}
