// Ceiling of the compositing kernels' inner loops (profiles/r2_issue_model.md): the product's own fwd_batch / bwd_batch
// (this file #includes csrc/raster.cu) run over a fixed shared-memory batch of 32 regular entries, again and again, with W
// one-warp CTAs per scheduler -- no staging, no tile ranges, no tail, no load imbalance.  What remains is the instruction
// stream itself, so cycles per list entry per scheduler here is what the real kernels could reach at best.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -o build/ubench_raster \
//        tools/ubench_raster.cu mini-3d-gaussian-splatting_b200/csrc/abi.cu && build/ubench_raster
#include "../mini-3d-gaussian-splatting_b200/csrc/raster.cu"

#include <cstdlib>
#include <vector>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

namespace gs {

__device__ __forceinline__ void fill_batch(float4* srec, int* sid, int lane, int cta, int n_ids) {
    // 32 regular entries around a 16x16 tile at the origin: sigma ~ 12 px, tiny opacity (no pixel ever saturates)
    const float mx = 8.f + 20.f * __sinf(lane * 1.7f), my = 8.f + 20.f * __cosf(lane * 2.3f);
    const float q = kNegHalfLog2e / (144.f + lane);
    srec[lane * 3 + 0] = make_float4(mx, my, q, 0.1f * q);
    srec[lane * 3 + 1] = make_float4(1.1f * q, 1e-4f, 3.f + lane, 0.5f);
    srec[lane * 3 + 2] = make_float4(0.25f, 0.75f, 1e4f, -13.287712f);      // g, b, 1/opacity, log2 opacity
    if (sid) sid[lane] = (cta * 977 + lane * 131) % n_ids;
}

__device__ __forceinline__ unsigned where_am_i() {
    unsigned smid, warpid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(warpid));
    return smid * 256 + warpid;
}

__global__ void __launch_bounds__(32, GS_FWD_MINB) fwd_loop_kernel(int rounds, long long* span, float* sink, unsigned* where) {
    __shared__ float4 srec[kBatch * 3];
    const int lane = threadIdx.x;
    fill_batch(srec, nullptr, lane, blockIdx.x, 1);
    __syncwarp();
    const float fpy = (float)(lane >> 1);
    float2 fpx[kPairs], A[kPairs], Cr[kPairs], Cg[kPairs], Cb[kPairs], Ds[kPairs];
    int ncons[kPx];
#pragma unroll
    for (int p = 0; p < kPairs; ++p) {
        fpx[p] = make_float2((float)((lane & 1) * kPx + 2 * p), (float)((lane & 1) * kPx + 2 * p + 1));
        A[p] = Cr[p] = Cg[p] = Cb[p] = Ds[p] = bc2(0.f);
        ncons[2 * p] = ncons[2 * p + 1] = -1;
    }
    const long long t0 = clock64();
    int walked = 0;
    for (int r = 0; r < rounds; ++r) walked += fwd_batch<true, false>(srec, kBatch, 0, fpy, fpx, A, Cr, Cg, Cb, Ds, ncons);
    const long long t1 = clock64();
    float s = (float)walked;
#pragma unroll
    for (int p = 0; p < kPairs; ++p) s += A[p].x + A[p].y + Cr[p].x + Cr[p].y + Cg[p].x + Cg[p].y + Cb[p].x + Cb[p].y + Ds[p].x + Ds[p].y;
    if (s == 12345.678f) sink[0] = s;
    if (lane == 0) { span[blockIdx.x] = t1 - t0; where[blockIdx.x] = where_am_i(); }
}

__global__ void __launch_bounds__(32, GS_BWD_MINB) bwd_loop_kernel(int rounds, int n_ids, float* grads, long long* span, float* sink, unsigned* where) {
    __shared__ float4 srec[kBatch * 3];
    __shared__ int sid[kBatch];
    __shared__ __align__(16) float red[kBwdGroup * kRedVals * kRedStride + 16];
    const int lane = threadIdx.x;
    fill_batch(srec, sid, lane, blockIdx.x, n_ids);
    __syncwarp();
    const float fpy = (float)(lane >> 1);
    float2 fpx[kPairs], A[kPairs], R[kPairs], gCr[kPairs], gCg[kPairs], gCb[kPairs], gDs[kPairs], gA[kPairs];
#pragma unroll
    for (int p = 0; p < kPairs; ++p) {
        fpx[p] = make_float2((float)((lane & 1) * kPx + 2 * p), (float)((lane & 1) * kPx + 2 * p + 1));
        A[p] = bc2(0.f);
        R[p] = make_float2(-0.3f - lane, -0.2f - p);
        gCr[p] = make_float2(0.1f + lane * 0.01f, 0.2f - p * 0.01f); gCg[p] = make_float2(-0.2f + lane * 0.02f, 0.1f * p);
        gCb[p] = make_float2(0.3f - lane * 0.01f, 0.3f + p); gDs[p] = make_float2(0.01f * lane, 0.02f * p); gA[p] = make_float2(0.05f * lane, 0.04f + p);
    }
    float* g_means2d = grads;
    float* g_conics = grads + 2 * n_ids;
    float* g_opac = grads + 6 * n_ids;
    float* g_depths = grads + 7 * n_ids;
    float* g_colors = grads + 8 * n_ids;
    BwdOut out;
    switch (lane) {
        case 0: out.base = g_means2d; out.stride = 2; break;
        case 1: out.base = g_means2d + 1; out.stride = 2; break;
        case 2: out.base = g_conics; out.stride = 4; break;
        case 3: out.base = g_conics + 1; out.stride = 4; break;
        case 4: out.base = g_conics + 3; out.stride = 4; break;
        case 5: out.base = g_opac; out.stride = 1; break;
        case 6: out.base = g_depths; out.stride = 1; break;
        case 7: out.base = g_colors; out.stride = 3; break;
        case 8: out.base = g_colors + 1; out.stride = 3; break;
        default: out.base = g_colors + 2; out.stride = 3; break;
    }
    const int red_v = lane % kRedVals, red_g = lane / kRedVals;
    out.red = red;
    out.red_src = reinterpret_cast<const float4*>(&red[red_v * kRedStride + red_g * 12]);
    out.sid = sid;
#if GS_BWD_RAWSUMS
    __shared__ __align__(16) float coef[kBatch * kCoefRow];
    stage_coef(true, lane, kBatch, srec, coef);
    __syncwarp();
    out.coef = coef;
    out.ia = lane <= 9 ? lane : 3;
    out.ib = lane == 0 ? 1 : (lane == 1 ? 0 : out.ia);
    out.ja = lane == 0 ? 0 : lane == 1 ? 2 : (lane <= 4 || lane == 10) ? 3 : lane == 5 ? 4 : 5;
    out.jb = lane <= 1 ? 1 : 6;
#endif
    const long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) bwd_batch<true>(srec, kBatch, lane, fpy, fpx, A, R, gCr, gCg, gCb, gDs, gA, out);
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int p = 0; p < kPairs; ++p) s += A[p].x + A[p].y + R[p].x + R[p].y;
    if (s == 12345.678f) sink[0] = s;
    if (lane == 0) { span[blockIdx.x] = t1 - t0; where[blockIdx.x] = where_am_i(); }
}

}  // namespace gs

int main() {
    int sms = 0;
    CHECK(cudaSetDevice(0));
    CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int n_ids = 1 << 20, rounds = 400;
    long long* d_span;
    float *d_sink, *d_grads;
    CHECK(cudaMalloc(&d_span, sms * 32 * sizeof(long long)));
    CHECK(cudaMalloc(&d_sink, 64));
    unsigned* d_where;
    CHECK(cudaMalloc(&d_where, sms * 32 * sizeof(unsigned)));
    std::vector<unsigned> hw(sms * 32);
    CHECK(cudaMalloc(&d_grads, (size_t)11 * n_ids * sizeof(float)));
    CHECK(cudaMemset(d_grads, 0, (size_t)11 * n_ids * sizeof(float)));
    std::vector<long long> h(sms * 32);
    printf("{\"sms\": %d, \"entries_per_warp\": %d, \"unit\": \"cycles per list entry per scheduler = slowest CTA span / (entries per warp x warps per scheduler)\", \"rows\": [\n", sms, rounds * 32);
    const int ws[] = {1, 2, 3, 4, 5, 6, 8};
    for (int which = 0; which < 2; ++which) {
        for (int wi = 0; wi < 7; ++wi) {
            const int w = ws[wi], ctas = sms * 4 * w;
            int resident = 0;
            if (which == 0) CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, gs::fwd_loop_kernel, 32, 0));
            else CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, gs::bwd_loop_kernel, 32, 0));
            if (4 * w > resident) continue;            // would not be co-resident: the span would count queueing
            for (int rep = 0; rep < 2; ++rep) {
                if (which == 0) gs::fwd_loop_kernel<<<ctas, 32>>>(rounds, d_span, d_sink, d_where);
                else gs::bwd_loop_kernel<<<ctas, 32>>>(rounds, n_ids, d_grads, d_span, d_sink, d_where);
                CHECK(cudaDeviceSynchronize());
            }
            CHECK(cudaMemcpy(h.data(), d_span, ctas * sizeof(long long), cudaMemcpyDeviceToHost));
            long long mx = 0;
            double mean = 0;
            for (int i = 0; i < ctas; ++i) { mx = h[i] > mx ? h[i] : mx; mean += (double)h[i] / ctas; }
            CHECK(cudaMemcpy(hw.data(), d_where, ctas * sizeof(unsigned), cudaMemcpyDeviceToHost));
            // how the one-warp CTAs were spread over the four schedulers of each SM (scheduler = hardware warp slot % 4), and
            // the mean span of the warps by the number of warps that shared their scheduler
            {
                std::vector<int> load(sms * 4, 0);
                for (int i = 0; i < ctas; ++i) load[(hw[i] >> 8) * 4 + (hw[i] & 3)]++;
                int hist[12] = {0};
                for (int i = 0; i < sms * 4; ++i) hist[load[i] < 11 ? load[i] : 11]++;
                double by_load[12] = {0}; int cnt_load[12] = {0};
                for (int i = 0; i < ctas; ++i) { const int l = load[(hw[i] >> 8) * 4 + (hw[i] & 3)]; by_load[l < 11 ? l : 11] += (double)h[i]; cnt_load[l < 11 ? l : 11]++; }
                printf("  {\"schedulers_by_resident_warps\": [");
                for (int l = 0; l < 12; ++l) printf("%d%s", hist[l], l < 11 ? ", " : "");
                printf("], \"cycles_per_entry_of_a_warp_by_scheduler_load\": [");
                for (int l = 0; l < 12; ++l) printf("%.1f%s", cnt_load[l] ? by_load[l] / cnt_load[l] / (rounds * 32.0) : 0.0, l < 11 ? ", " : "");
                printf("]},\n");
            }
            printf("  {\"kernel\": \"%s\", \"warps_per_scheduler\": %d, \"resident_ctas_per_sm_max\": %d, \"cycles_per_entry_slowest\": %.1f, \"cycles_per_entry_mean\": %.1f},\n",
                   which == 0 ? "fwd_batch<fast>" : "bwd_batch<fast>", w, resident, (double)mx / (rounds * 32.0 * w), mean / (rounds * 32.0 * w));
        }
    }
    printf("  {}\n]}\n");
    return 0;
}
