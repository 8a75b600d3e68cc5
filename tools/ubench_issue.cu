// Micro-benchmark behind profiles/r2_issue_model.md: what does a packed FP32 instruction (FFMA2 / FMUL2 / FADD2)
// cost on sm_100 -- one issue slot and two FMA-pipe cycles (so that ALU / MUFU / LDS instructions of the same or other
// warps can fill the second cycle), or two issue slots?  The answer fixes the roofline of the compositing kernels.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench_issue tools/ubench_issue.cu && build/ubench_issue
//
// Every kernel runs `iters` rounds of a fixed instruction mix on 8 independent register chains per thread, one CTA per
// SM, W warps per scheduler (CTA = 128*W threads); the slowest CTA's clock64() span / iters is printed as cycles per
// round per scheduler, next to the number of instructions of each kind one warp issues per round.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

enum Mix { kFfma2 = 0, kFfma, kFfma2Alu, kFfmaAlu, kFfma2Mufu, kFfma2Setp, kAlu, kFfma2x2Alu, kFfma2Lds, kMufu, kFfma16, kNumMix };
static const char* kNames[kNumMix] = {
    "8 FFMA2", "8 FFMA", "8 FFMA2 + 8 LOP3", "8 FFMA + 8 LOP3", "8 FFMA2 + 2 MUFU.EX2", "8 FFMA2 + 8 FSETP/FSEL pairs",
    "8 LOP3", "8 FFMA2 + 16 LOP3", "8 FFMA2 + 4 LDS", "2 MUFU.EX2", "16 FFMA"};

template <int kMix>
__global__ void __launch_bounds__(1024, 1) mix_kernel(int iters, float seed, long long* span, float* sink) {
    __shared__ float sh[1024];
    sh[threadIdx.x] = seed;
    __syncthreads();
    float2 acc[8];
    unsigned ia[16];
    float m0 = seed, m1 = seed * 0.5f;
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = make_float2(seed + k, seed - k);
#pragma unroll
    for (int k = 0; k < 16; ++k) ia[k] = threadIdx.x * 7 + k;
    const float2 a = make_float2(1.0001f, 0.9999f), b = make_float2(seed, -seed);
    const long long t0 = clock64();
#pragma unroll 4
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (kMix == kFfma2 || kMix == kFfma2Alu || kMix == kFfma2Mufu || kMix == kFfma2Setp || kMix == kFfma2x2Alu || kMix == kFfma2Lds)
                acc[k] = __ffma2_rn(acc[k], a, b);
            if (kMix == kFfma || kMix == kFfmaAlu) acc[k].x = __fmaf_rn(acc[k].x, a.x, b.x);
            if (kMix == kFfma16) { acc[k].x = __fmaf_rn(acc[k].x, a.x, b.x); acc[k].y = __fmaf_rn(acc[k].y, a.y, b.y); }
            if (kMix == kFfma2Alu || kMix == kFfmaAlu || kMix == kAlu || kMix == kFfma2x2Alu)
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(ia[k]) : "r"(ia[(k + 1) & 7]), "r"(it));
            if (kMix == kFfma2x2Alu)
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(ia[8 + k]) : "r"(ia[8 + ((k + 1) & 7)]), "r"(it));
            if (kMix == kFfma2Setp) {
                // the compositing kernels' live-weight idiom: two chained compares and one select per value
                float r;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.lt.f32 p, %1, %2;\n\tselp.f32 %0, %1, 0f00000000, p;\n\t}"
                             : "=f"(r) : "f"(acc[k].y), "f"(seed * 1e30f));
                acc[k].y = r;
            }
            if ((kMix == kFfma2Mufu || kMix == kMufu) && (k == 0 || k == 4)) {
                if (k) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(m1));
                else asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(m0));
            }
            if (kMix == kFfma2Lds && (k & 1) == 0) {
                float v;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"((unsigned)__cvta_generic_to_shared(&sh[(threadIdx.x + k) & 1023])));
                m0 += v * 0.f;
            }
        }
    }
    const long long t1 = clock64();
    float s = m0 + m1;
    unsigned x = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += acc[k].x + acc[k].y;
#pragma unroll
    for (int k = 0; k < 16; ++k) x ^= ia[k];
    if (s == 12345.678f && x == 77) sink[0] = s;      // never true: keeps the chains alive
    if (threadIdx.x == 0) span[blockIdx.x] = t1 - t0;
}

template <int kMix>
static double run(int warps_per_sched, int iters, long long* d_span, float* d_sink, int sms) {
    mix_kernel<kMix><<<sms, 128 * warps_per_sched>>>(iters, 1.0f, d_span, d_sink);
    mix_kernel<kMix><<<sms, 128 * warps_per_sched>>>(iters, 1.0f, d_span, d_sink);
    CHECK(cudaDeviceSynchronize());
    static long long h[1024];
    CHECK(cudaMemcpy(h, d_span, sms * sizeof(long long), cudaMemcpyDeviceToHost));
    long long mx = 0;
    for (int i = 0; i < sms; ++i) mx = h[i] > mx ? h[i] : mx;
    return (double)mx / iters;
}

int main() {
    int dev = 0, sms = 0;
    CHECK(cudaSetDevice(dev));
    CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    long long* d_span;
    float* d_sink;
    CHECK(cudaMalloc(&d_span, 1024 * sizeof(long long)));
    CHECK(cudaMalloc(&d_sink, 64));
    const int iters = 20000;
    printf("{\"sms\": %d, \"iters\": %d, \"unit\": \"cycles per round per scheduler (all warps of the scheduler together)\", \"rows\": [\n", sms, iters);
    for (int mix = 0; mix < kNumMix; ++mix) {
        printf("  {\"mix\": \"%s\"", kNames[mix]);
        for (int w = 1; w <= 8; w *= 2) {
            double c = 0;
            switch (mix) {
                case kFfma2: c = run<kFfma2>(w, iters, d_span, d_sink, sms); break;
                case kFfma: c = run<kFfma>(w, iters, d_span, d_sink, sms); break;
                case kFfma2Alu: c = run<kFfma2Alu>(w, iters, d_span, d_sink, sms); break;
                case kFfmaAlu: c = run<kFfmaAlu>(w, iters, d_span, d_sink, sms); break;
                case kFfma2Mufu: c = run<kFfma2Mufu>(w, iters, d_span, d_sink, sms); break;
                case kFfma2Setp: c = run<kFfma2Setp>(w, iters, d_span, d_sink, sms); break;
                case kAlu: c = run<kAlu>(w, iters, d_span, d_sink, sms); break;
                case kFfma2x2Alu: c = run<kFfma2x2Alu>(w, iters, d_span, d_sink, sms); break;
                case kFfma2Lds: c = run<kFfma2Lds>(w, iters, d_span, d_sink, sms); break;
                case kMufu: c = run<kMufu>(w, iters, d_span, d_sink, sms); break;
                case kFfma16: c = run<kFfma16>(w, iters, d_span, d_sink, sms); break;
            }
            // cycles per round for ONE warp's round, with w warps sharing the scheduler: span / iters is the time in
            // which each of the w warps finished one round, so the scheduler spent span / iters cycles on w rounds
            printf(", \"w%d\": %.2f", w, c / w);
        }
        printf("}%s\n", mix + 1 < kNumMix ? "," : "");
    }
    printf("]}\n");
    return 0;
}
