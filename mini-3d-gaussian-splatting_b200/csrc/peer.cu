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
template <int kU>
__global__ void __launch_bounds__(512)
peer_allreduce_multimem_kernel(float* mc, int world, int rank, int64_t sum_off4, int64_t sum_n4, int64_t max_off4, int64_t max_n4) {
    const int64_t threads = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    {
        const int64_t per = (sum_n4 + world - 1) / world;
        const int64_t begin = (int64_t)rank * per, end = min(begin + per, sum_n4);
        // kU independent in-switch reductions in flight per thread (a reduction's latency is the slowest of `world` reads)
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


// ------------------------------------------------------------------------------------------------
// TMA variant (GS_PEER_TMA): the same exchange with the data moved by bulk asynchronous copies instead of per-thread
// 16-byte loads/stores.  An SM can keep only a few dozen peer requests of <= 128 B in flight, which is what pins the
// load/store kernel to ~1/3 of the NVLink rate whatever its occupancy (profiles/r1_v5_multigpu.md); one
// cp.async.bulk moves 8 KB per request, needs no registers for the data in flight and completes on an mbarrier.
// Per CTA a ring of kTmaStages stages; a stage holds the same chunk of this rank's slice from every peer:
//   thread 0:    expect_tx + one bulk load per peer  ->  stage, signalled on the stage's mbarrier
//   all threads: wait, reduce the `world` copies in a fixed order into copy 0 (SUM or MAX), proxy fence, barrier
//   thread 0:    one bulk store of copy 0 per peer (commit group), wait until the stores have READ the stage, refill it
// One owner per element and a fixed summation order, as in the load/store kernel: results are bit-identical on all
// ranks and bit-identical to that kernel.
// ------------------------------------------------------------------------------------------------
constexpr int kTmaThreads = 256;
constexpr int kTmaStages = 3;
constexpr int kTmaStageBytes = 64 * 1024;        // world x chunk bytes

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tGS_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra GS_DONE_%=;\n\tbra GS_WAIT_%=;\n\tGS_DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int kWorld>
__global__ void __launch_bounds__(kTmaThreads, 1)
peer_allreduce_tma_kernel(PeerPtrs peers, int world_rt, int rank, int64_t sum_off4, int64_t sum_n4, int64_t max_off4, int64_t max_n4) {
    extern __shared__ __align__(128) unsigned char tma_smem[];
    const int world = kWorld > 0 ? kWorld : world_rt;
    const int chunk4 = kTmaStageBytes / 16 / world;                    // float4 per peer per stage (multiple of 8 -> 128-byte pieces)
    uint64_t* bars = reinterpret_cast<uint64_t*>(tma_smem);              // kTmaStages mbarriers
    float4* stages = reinterpret_cast<float4*>(tma_smem + 128);
    const int tid = threadIdx.x;

    // this rank's slices of the two regions, cut into chunks
    auto slice = [&](int64_t off4, int64_t n4, int64_t& b, int64_t& e) {
        const int64_t per = (n4 + world - 1) / world;
        b = min((int64_t)rank * per, n4);
        e = min(b + per, n4);
        b += off4; e += off4;
    };
    int64_t sb, se, mb, me;
    slice(sum_off4, sum_n4, sb, se);
    slice(max_off4, max_n4, mb, me);
    const int64_t nc_sum = (se - sb + chunk4 - 1) / chunk4, nc_max = (me - mb + chunk4 - 1) / chunk4;
    const int64_t total = nc_sum + nc_max;
    const int64_t mine = total > (int64_t)blockIdx.x ? (total - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    auto chunk_of = [&](int64_t j, int64_t& start, int& len, bool& is_max) {
        const int64_t t = (int64_t)blockIdx.x + j * gridDim.x;
        is_max = t >= nc_sum;
        const int64_t b = is_max ? mb : sb, e = is_max ? me : se;
        start = b + (is_max ? t - nc_sum : t) * chunk4;
        len = (int)min((int64_t)chunk4, e - start);
    };
    auto issue = [&](int64_t j) {                                       // thread 0 only
        int64_t start; int len; bool is_max;
        chunk_of(j, start, len, is_max);
        const int s = (int)(j % kTmaStages);
        mbar_expect_tx(&bars[s], (uint32_t)(world * len * 16));
        for (int p = 0; p < world; ++p)
            bulk_g2s(stages + ((int64_t)s * world + p) * chunk4, reinterpret_cast<const float4*>(peers.p[p]) + start, (uint32_t)(len * 16), &bars[s]);
    };

    if (tid == 0) {
        for (int s = 0; s < kTmaStages; ++s) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0)
        for (int64_t j = 0; j < mine && j < kTmaStages; ++j) issue(j);

    for (int64_t j = 0; j < mine; ++j) {
        int64_t start; int len; bool is_max;
        chunk_of(j, start, len, is_max);
        const int s = (int)(j % kTmaStages);
        mbar_wait(&bars[s], (uint32_t)((j / kTmaStages) & 1));
        float4* st = stages + (int64_t)s * world * chunk4;
        for (int i = tid; i < len; i += kTmaThreads) {
            float4 acc = st[i];
#pragma unroll
            for (int p = 1; p < (kWorld > 0 ? kWorld : kMaxPeers); ++p) {
                if (p < world) {
                    const float4 v = st[(int64_t)p * chunk4 + i];
                    if (is_max) {
                        acc.x = fmaxf(acc.x, v.x); acc.y = fmaxf(acc.y, v.y); acc.z = fmaxf(acc.z, v.z); acc.w = fmaxf(acc.w, v.w);
                    } else {
                        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
                    }
                }
            }
            st[i] = acc;
        }
        fence_async_smem();                           // generic-proxy writes -> visible to the bulk stores below
        __syncthreads();
        if (tid == 0) {
            for (int p = 0; p < world; ++p)
                bulk_s2g(reinterpret_cast<float4*>(peers.p[p]) + start, st, (uint32_t)(len * 16));
            bulk_commit();
            if (j + kTmaStages < mine) {
                bulk_wait_read_all();                 // the stores have read the stage: it may be overwritten
                issue(j + kTmaStages);
            }
        }
    }
    if (tid == 0) {
        bulk_wait_all();                              // every store of this CTA is complete before the kernel (and the barrier after it) ends
        __threadfence_system();
    }
}

}  // namespace gs

using namespace gs;

extern "C" int gs_peer_allreduce(const uint64_t* peer_ptrs_host, uint64_t multicast_ptr, int32_t world, int32_t rank,
                                 int64_t sum_offset, int64_t sum_count, int64_t max_offset, int64_t max_count,
                                 int32_t flags, void* stream) {
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
        // 16 in-switch reductions in flight per thread, 4 CTAs per SM: 0.148 ms at 8 GPUs against 0.162 ms with 4 / 2
        // (profiles/r2_multigpu.md); both are tuning knobs of tools/peer_pieces.py
        int unroll = 16;
        if (const char* e = getenv("GS_PEER_MC_UNROLL")) unroll = atoi(e);
        float* mcp = reinterpret_cast<float*>(multicast_ptr);
        const dim3 grid((unsigned)(sms * (getenv("GS_PEER_GRID_MULT") ? mult : 4)));
        if (unroll >= 16) peer_allreduce_multimem_kernel<16><<<grid, block, 0, st>>>(mcp, world, rank, so, sn, mo, mn);
        else if (unroll >= 8) peer_allreduce_multimem_kernel<8><<<grid, block, 0, st>>>(mcp, world, rank, so, sn, mo, mn);
        else peer_allreduce_multimem_kernel<4><<<grid, block, 0, st>>>(mcp, world, rank, so, sn, mo, mn);
        GS_CUDA_TRY(cudaGetLastError());
        count_launches(1);
        return GS_OK;
    }
    if (flags & GS_PEER_TMA) {
        const size_t smem = 128 + (size_t)kTmaStages * kTmaStageBytes;
        const dim3 tgrid((unsigned)sms);
#define GS_TMA_LAUNCH(W)                                                                                                   \
    do {                                                                                                                   \
        GS_CUDA_TRY(cudaFuncSetAttribute(peer_allreduce_tma_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        peer_allreduce_tma_kernel<W><<<tgrid, kTmaThreads, smem, st>>>(peers, world, rank, so, sn, mo, mn);                \
    } while (0)
        switch (world) {
            case 2: GS_TMA_LAUNCH(2); break;
            case 4: GS_TMA_LAUNCH(4); break;
            case 8: GS_TMA_LAUNCH(8); break;
            default: GS_TMA_LAUNCH(0); break;
        }
#undef GS_TMA_LAUNCH
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
