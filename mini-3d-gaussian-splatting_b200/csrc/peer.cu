// Gradient / statistics exchange of the view-sharded step over NVLink peer memory.
//
// Every rank holds the same flat buffer [ gradients + statistics | max radii ] in symmetric memory
// (one allocation per rank, all mapped into every rank's address space).  One kernel per rank does the
// whole exchange: rank r owns the r-th slice of both regions, reads that slice from EVERY peer
// (16-byte loads over NVLink, all peers' loads in flight together), reduces it -- SUM for the first
// region, MAX for the second -- in a fixed peer order, and stores the result back into every peer's
// buffer.  Reads are the link's inbound direction and writes its outbound one, so both are busy at once
// and the reduce-scatter and the all-gather of a two-phase all-reduce collapse into one pass with no
// intermediate buffer.  A slice is touched by its owner only, so the kernel needs no synchronisation of
// its own; the caller brackets it with two device-side barriers (all buffers complete / all results landed).
// Every rank ends up with bit-identical values (one owner per element, fixed summation order).
#include "common.cuh"

#include <stdlib.h>

namespace gs {

constexpr int kMaxPeers = 16;
struct PeerPtrs {
    float* p[kMaxPeers];
};

#ifndef GS_PEER_LD
#define GS_PEER_LD 0
#endif
__device__ __forceinline__ float4 ld_peer(const float4* a) {
    float4 v;
#if GS_PEER_LD == 0
    // .cv: never served from a stale line -- the data was written by another GPU
    asm volatile("ld.global.cv.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(a));
#elif GS_PEER_LD == 1
    asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(a));
#else
    asm volatile("ld.global.relaxed.sys.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(a));
#endif
    return v;
}


template <int kWorld, bool kMax>
__device__ __forceinline__ void reduce_slice(const PeerPtrs& peers, int world, int rank, int64_t off4, int64_t n4) {
    constexpr int kW = kWorld > 0 ? kWorld : kMaxPeers;
    // 8 independent 16-byte loads per thread whatever the world size: ~10 MB in flight per GPU
    constexpr int kPeerUnroll = (kWorld > 0 && kWorld < 8) ? 8 / kWorld : 1;
    const int w = kWorld > 0 ? kWorld : world;
    const int64_t per = (n4 + w - 1) / w;
    const int64_t begin = (int64_t)rank * per;
    const int64_t end = min(begin + per, n4);
    const int64_t threads = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t i0 = begin + tid; i0 < end; i0 += threads * kPeerUnroll) {
        float4 v[kPeerUnroll][kW];
#pragma unroll
        for (int u = 0; u < kPeerUnroll; ++u) {
            const int64_t i = i0 + u * threads;
#pragma unroll
            for (int p = 0; p < kW; ++p)
                if (p < w && i < end) v[u][p] = ld_peer(reinterpret_cast<const float4*>(peers.p[p]) + off4 + i);
        }
#pragma unroll
        for (int u = 0; u < kPeerUnroll; ++u) {
            const int64_t i = i0 + u * threads;
            if (i >= end) break;
            float4 acc = v[u][0];
#pragma unroll
            for (int p = 1; p < kW; ++p) {
                if (p < w) {
                    if (kMax) {
                        acc.x = fmaxf(acc.x, v[u][p].x); acc.y = fmaxf(acc.y, v[u][p].y);
                        acc.z = fmaxf(acc.z, v[u][p].z); acc.w = fmaxf(acc.w, v[u][p].w);
                    } else {
                        acc.x += v[u][p].x; acc.y += v[u][p].y; acc.z += v[u][p].z; acc.w += v[u][p].w;
                    }
                }
            }
#pragma unroll
            for (int p = 0; p < kW; ++p)
                if (p < w) reinterpret_cast<float4*>(peers.p[p])[off4 + i] = acc;
        }
    }
}

template <int kWorld>
__global__ void __launch_bounds__(512)
peer_allreduce_kernel(PeerPtrs peers, int world, int rank, int64_t sum_off4, int64_t sum_n4, int64_t max_off4, int64_t max_n4) {
    reduce_slice<kWorld, false>(peers, world, rank, sum_off4, sum_n4);
    reduce_slice<kWorld, true>(peers, world, rank, max_off4, max_n4);
}

// NVSwitch multicast (NVLS) variant: the switch does the arithmetic.  `mc` is the multicast address of the
// same symmetric buffer: a multimem.ld_reduce returns the element reduced over every rank's copy (one
// 16-byte response per 16 bytes of slice instead of one per peer), and a multimem.st writes it into every
// rank's copy (one request instead of one per peer): link traffic per GPU drops from 2*(G-1)/G of the buffer
// to 2/G of it.  The MAX region holds non-negative floats, whose order is that of their bit patterns, so it is
// reduced as .max.u32 (multimem has no f32 max).
__global__ void __launch_bounds__(512)
peer_allreduce_multimem_kernel(float* mc, int world, int rank, int64_t sum_off4, int64_t sum_n4, int64_t max_off4, int64_t max_n4) {
    const int64_t threads = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    {
        const int64_t per = (sum_n4 + world - 1) / world;
        const int64_t begin = (int64_t)rank * per, end = min(begin + per, sum_n4);
        constexpr int kU = 4;                              // independent reductions in flight per thread
        for (int64_t i0 = begin + tid; i0 < end; i0 += threads * kU) {
            float4 v[kU];
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const int64_t i = i0 + u * threads;
                if (i < end)
                    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                                 : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w)
                                 : "l"(reinterpret_cast<float4*>(mc) + sum_off4 + i) : "memory");
            }
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const int64_t i = i0 + u * threads;
                if (i < end)
                    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
                                 :: "l"(reinterpret_cast<float4*>(mc) + sum_off4 + i), "f"(v[u].x), "f"(v[u].y), "f"(v[u].z), "f"(v[u].w)
                                 : "memory");
            }
        }
    }
    {
        const int64_t n = max_n4 * 4, off = max_off4 * 4;
        const int64_t per = (n + world - 1) / world;
        const int64_t begin = (int64_t)rank * per, end = min(begin + per, n);
        for (int64_t i = begin + tid; i < end; i += threads) {
            unsigned* a = reinterpret_cast<unsigned*>(mc) + off + i;
            unsigned v;
            asm volatile("multimem.ld_reduce.relaxed.sys.global.max.u32 %0, [%1];" : "=r"(v) : "l"(a) : "memory");
            asm volatile("multimem.st.relaxed.sys.global.u32 [%0], %1;" :: "l"(a), "r"(v) : "memory");
        }
    }
}

}  // namespace gs

using namespace gs;

extern "C" int gs_peer_allreduce(const uint64_t* peer_ptrs_host, uint64_t multicast_ptr, int32_t world, int32_t rank,
                                 int64_t sum_offset, int64_t sum_count, int64_t max_offset, int64_t max_count,
                                 void* stream) {
    GS_REQUIRE(peer_ptrs_host != nullptr, "peer_ptrs_host is NULL");
    GS_REQUIRE(world >= 1 && world <= kMaxPeers, "world must be 1..16");
    GS_REQUIRE(rank >= 0 && rank < world, "rank out of range");
    GS_REQUIRE(sum_offset >= 0 && sum_count >= 0 && max_offset >= 0 && max_count >= 0, "negative extent");
    GS_REQUIRE((sum_offset | sum_count | max_offset | max_count) % 4 == 0, "offsets and counts must be multiples of 4 floats");
    PeerPtrs peers;
    for (int p = 0; p < kMaxPeers; ++p) peers.p[p] = p < world ? reinterpret_cast<float*>(peer_ptrs_host[p]) : nullptr;
    for (int p = 0; p < world; ++p) {
        GS_REQUIRE(peers.p[p] != nullptr, "NULL peer pointer");
        GS_REQUIRE(reinterpret_cast<uintptr_t>(peers.p[p]) % 16 == 0, "peer buffers must be 16-byte aligned");
    }
    if (world == 1 || (sum_count == 0 && max_count == 0)) return GS_OK;
    DeviceGuard guard(peers.p[rank]);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int mult = 2;
    if (const char* e = getenv("GS_PEER_GRID_MULT")) mult = atoi(e) > 0 ? atoi(e) : 2;      // tuning knob (tools/peer_pieces.py)
    const dim3 grid((unsigned)(sms * mult)), block(512);
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t so = sum_offset / 4, sn = sum_count / 4, mo = max_offset / 4, mn = max_count / 4;
    if (multicast_ptr != 0) {
        GS_REQUIRE(multicast_ptr % 16 == 0, "multicast address must be 16-byte aligned");
        peer_allreduce_multimem_kernel<<<grid, block, 0, st>>>(reinterpret_cast<float*>(multicast_ptr), world, rank, so, sn, mo, mn);
        GS_CUDA_TRY(cudaGetLastError());
        count_launches(1);
        return GS_OK;
    }
    switch (world) {
        case 2: peer_allreduce_kernel<2><<<grid, block, 0, st>>>(peers, world, rank, so, sn, mo, mn); break;
        case 4: peer_allreduce_kernel<4><<<grid, block, 0, st>>>(peers, world, rank, so, sn, mo, mn); break;
        case 8: peer_allreduce_kernel<8><<<grid, block, 0, st>>>(peers, world, rank, so, sn, mo, mn); break;
        default: peer_allreduce_kernel<0><<<grid, block, 0, st>>>(peers, world, rank, so, sn, mo, mn); break;
    }
    GS_CUDA_TRY(cudaGetLastError());
    count_launches(1);
    return GS_OK;
}
