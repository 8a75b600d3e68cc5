// Stage S: the global depth order (src/core/renderer.py:222-239), as ONE persistent cooperative kernel.
//
// gs_bin_prepare must deliver: sorted_ids = splat indices in ascending depth, ties in ascending index (what a stable
// argsort gives), splats without tiles behind them (0xFFFFFFFE: visible with an empty AABB, then 0xFFFFFFFF: culled);
// offsets = exclusive prefix sum of tiles_touched in that order; counters = {splats with tiles, tile pairs D, visible}.
//
// A library radix sort does that in ~18 launches (iota, histogram, four 8-bit passes, scan, ...) of 130-190 us at
// N = 1 M, where the data (4 MB of keys) would stream in 2 us: the stage is bound by launch boundaries and by sorting
// bits that never differ.  Here:
//   * one launch: 1 CTA per SM, all co-resident (cooperative launch), phases separated by grid-wide barriers;
//   * ids are implicit in the first pass (no iota), the key range [min, max] is found first and only the bits in which
//     (key - min) can differ are sorted: depths of one scene span a few binades, so 24 bits = 3 passes instead of 4;
//     the two "no tiles" codes are remapped to max-min+1 / max-min+2 and simply sort to the end;
//   * per pass every CTA owns a contiguous chunk of the current sequence: it histograms its chunk (256 digit bins) into
//     a [CTA][digit] table, and after the barrier derives where each of its digit runs starts from the table (all CTAs
//     before it, all smaller digits) -- no decoupled look-back, no atomics on global memory, fully deterministic;
//     the items are ranked inside a tile with warp match_any (stable), reordered through shared memory and stored
//     in runs;
//   * the LAST pass carries a second, weighted table (sum of tiles_touched per CTA and digit) through the same scan, so
//     the exclusive prefix sum of tiles_touched in sorted order -- `offsets` -- falls out of the scatter itself: no
//     gather+scan pass, and the three counters come from the per-CTA partials.
// LSD with stable passes => stable overall; the result is bit-identical to the library path it replaces
// (tests/test_gpu_parity.py::test_sort_keys_bit_exact_vs_oracle and the whole-frame tests compare it with the oracle's sort).
#include "common.cuh"

#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace gs {

constexpr int kSortThreads = 512;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortItems = 14;                          // items per thread per tile
constexpr int kSortTile = kSortThreads * kSortItems;    // 7 168 items: one tile per CTA at N = 1 M on 148 SMs
constexpr int kSortWarpSpan = kSortItems * 32;
constexpr int kRadix = 256;
constexpr int kMaxSortGrid = 1024;                      // workspace is laid out for at most this many CTAs

struct DepthSortArgs {
    int64_t n;
    int64_t chunk;                   // items per CTA
    const uint32_t* keys_in;
    const int32_t* tiles_touched;
    uint32_t* kbuf[2];
    int32_t* ibuf[2];
    uint32_t* wbuf[2];               // tiles_touched carried along with the items (a random gather per item would cost more)
    uint32_t* table;                 // [grid][256] digit counts of the current pass
    unsigned long long* wtable;      // [grid][256] tiles_touched sums, last pass only
    uint32_t* partial;               // [grid][4]: min valid key, max valid key, #0xFFFFFFFE, #0xFFFFFFFF
    int32_t* sorted_ids;
    int64_t* offsets;
    int64_t* counters;
};

struct SortSmem {
    uint32_t key_in[kSortTile];              // the tile as it lies in the source sequence (bulk-copied)
    int32_t id_in[kSortTile];
    uint32_t wgt_in[kSortTile];              // tiles_touched of the items (last pass)
    uint16_t rank[kSortTile];                // rank of an item among the items of its warp with the same digit
    uint32_t key_out[kSortTile];             // the tile sorted by the current digit
    int32_t id_out[kSortTile];
    uint32_t wscan[kSortTile + 4];           // last pass: tiles_touched in sorted order, then its exclusive prefix sum
    uint32_t wcnt[kSortWarps][kRadix];       // per-warp digit counts, then per-warp exclusive bases
    uint32_t hist[kRadix];
    uint32_t whist[kRadix];
    uint32_t tcount[kRadix];                 // digit counts of the current tile
    uint32_t texcl[kRadix + 1];              // tile-local start of each digit run
    uint32_t run[kRadix];                    // global position where this CTA's next item of digit d goes
    unsigned long long wrun[kRadix];         // ... and the prefix sum of tiles_touched before it
    uint32_t scan_tmp[kSortThreads / 32 + 1];
    unsigned long long scan_tmp64[kRadix / 32 + 1];
    uint32_t red[4][kSortWarps];
    unsigned long long mbar;                 // completion barrier of the bulk copies
};

__device__ __forceinline__ uint32_t remap_key(uint32_t k, uint32_t kmin, uint32_t span) {
    // valid keys -> [0, span]; 0xFFFFFFFE -> span + 1; 0xFFFFFFFF -> span + 2 (order preserved)
    return k >= 0xFFFFFFFEu ? span + 1u + (k & 1u) : k - kmin;
}

// Lanes of the warp whose 8-bit digit equals this lane's (invalid lanes match nobody): eight ballots, data-independent cost
// (__match_any_sync takes one micro-coded round per DISTINCT value in the warp, ~30 with random digits).
__device__ __forceinline__ unsigned match_digit(uint32_t d, bool valid) {
    unsigned peers = __ballot_sync(0xffffffffu, valid);
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        const bool bit = (d >> b) & 1u;
        const unsigned vote = __ballot_sync(0xffffffffu, bit);
        peers &= bit ? vote : ~vote;
    }
    return valid ? peers : 0u;
}

// exclusive scan of one value per thread over the first 256 threads (8 warps); returns the exclusive prefix, total in *total
__device__ __forceinline__ uint32_t scan256(uint32_t v, uint32_t* tmp, int tid, uint32_t* total) {
    const int lane = tid & 31, wid = tid >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t x = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += x;
    }
    if (tid < kRadix && lane == 31) tmp[wid] = inc;
    __syncthreads();
    uint32_t base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < kRadix / 32; ++w) {
        const uint32_t t = tmp[w];
        if (w < wid) base += t;
        tot += t;
    }
    __syncthreads();
    if (total) *total = tot;
    return base + inc - v;
}
__device__ __forceinline__ unsigned long long scan256_64(unsigned long long v, unsigned long long* tmp, int tid, unsigned long long* total) {
    const int lane = tid & 31, wid = tid >> 5;
    unsigned long long inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long x = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += x;
    }
    if (tid < kRadix && lane == 31) tmp[wid] = inc;
    __syncthreads();
    unsigned long long base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < kRadix / 32; ++w) {
        const unsigned long long t = tmp[w];
        if (w < wid) base += t;
        tot += t;
    }
    __syncthreads();
    if (total) *total = tot;
    return base + inc - v;
}

// ---- bulk asynchronous copies (TMA, non-tensor form): one request moves a whole tile, completion on an mbarrier ----
__device__ __forceinline__ uint32_t ds_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ds_mbar_init(unsigned long long* bar) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(ds_smem_u32(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void ds_mbar_expect(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ds_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ds_mbar_wait(unsigned long long* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tDS_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DS_DONE_%=;\n\tbra DS_WAIT_%=;\n\tDS_DONE_%=:\n\t}" ::"r"(ds_smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void ds_bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(ds_smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(ds_smem_u32(bar)) : "memory");
}

#ifndef GS_SORT_TIMING
#define GS_SORT_TIMING 0
#endif
#if GS_SORT_TIMING
__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define GS_STAMP(k) do { if (blockIdx.x == 0 && threadIdx.x == 0) reinterpret_cast<unsigned long long*>(a.partial + kMaxSortGrid * 4)[k] = gtimer(); } while (0)
#else
#define GS_STAMP(k) do { } while (0)
#endif

__global__ void __launch_bounds__(kSortThreads, 1)
depth_sort_kernel(DepthSortArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SortSmem& s = *reinterpret_cast<SortSmem*>(smem_raw);
    cg::grid_group grid = cg::this_grid();
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int cta = blockIdx.x, G = gridDim.x;
    const int64_t begin = min((int64_t)cta * a.chunk, a.n), end = min(begin + a.chunk, a.n);
    uint32_t tma_parity = 0u;
    if (tid == 0) ds_mbar_init(&s.mbar);

    GS_STAMP(0);
    // ---- phase 0: key range and the two "no tiles" counts of this chunk -------------------------------------
    {
        uint32_t kmin = 0xFFFFFFFFu, kmax = 0u, n_fe = 0u, n_ff = 0u;
#pragma unroll 4
        for (int64_t i = begin + tid; i < end; i += kSortThreads) {
            const uint32_t k = a.keys_in[i];
            if (k < 0xFFFFFFFEu) { kmin = min(kmin, k); kmax = max(kmax, k); }
            n_fe += (k == 0xFFFFFFFEu);
            n_ff += (k == 0xFFFFFFFFu);
        }
        kmin = __reduce_min_sync(0xffffffffu, kmin);
        kmax = __reduce_max_sync(0xffffffffu, kmax);
        n_fe = __reduce_add_sync(0xffffffffu, n_fe);
        n_ff = __reduce_add_sync(0xffffffffu, n_ff);
        if (lane == 0) { s.red[0][wid] = kmin; s.red[1][wid] = kmax; s.red[2][wid] = n_fe; s.red[3][wid] = n_ff; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < kSortWarps; ++w) {
                kmin = min(kmin, s.red[0][w]); kmax = max(kmax, s.red[1][w]); n_fe += s.red[2][w]; n_ff += s.red[3][w];
            }
            a.partial[cta * 4 + 0] = kmin; a.partial[cta * 4 + 1] = kmax; a.partial[cta * 4 + 2] = n_fe; a.partial[cta * 4 + 3] = n_ff;
        }
    }
    GS_STAMP(1);
    grid.sync();
    GS_STAMP(2);
    uint32_t kmin = 0xFFFFFFFFu, kmax = 0u;
    unsigned long long n_fe = 0, n_ff = 0;
    for (int c = lane; c < G; c += 32) {                 // every warp reduces the G partials for itself
        kmin = min(kmin, a.partial[c * 4 + 0]);
        kmax = max(kmax, a.partial[c * 4 + 1]);
        n_fe += a.partial[c * 4 + 2];
        n_ff += a.partial[c * 4 + 3];
    }
    kmin = __reduce_min_sync(0xffffffffu, kmin);
    kmax = __reduce_max_sync(0xffffffffu, kmax);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        n_fe += __shfl_xor_sync(0xffffffffu, n_fe, o);
        n_ff += __shfl_xor_sync(0xffffffffu, n_ff, o);
    }
    if (kmin > kmax) kmin = kmax = 0u;                   // no splat has tiles
    const uint32_t span = kmax - kmin;                   // < 2^31: valid keys are bit patterns of positive floats
    const int bits = 32 - __clz(span + 2u);
    const int passes = (bits + 7) / 8;
    // a chunk of at most one tile is staged and ranked ONCE per pass: before the barrier for the histogram, and the same
    // ranks serve the scatter after it (N <= 148 x 7 168 on a B200); longer chunks are staged and ranked twice
    const bool single_tile = a.chunk <= kSortTile;

    for (int pass = 0; pass < passes; ++pass) {
        const bool first = pass == 0, last = pass == passes - 1;
        const int shift = pass * 8;
        const uint32_t* ksrc = first ? a.keys_in : a.kbuf[(pass - 1) & 1];
        const int32_t* isrc = first ? nullptr : a.ibuf[(pass - 1) & 1];
        const uint32_t* wsrc = first ? reinterpret_cast<const uint32_t*>(a.tiles_touched) : a.wbuf[(pass - 1) & 1];
        uint32_t* kdst = a.kbuf[pass & 1];
        int32_t* idst = a.ibuf[pass & 1];
        uint32_t* wdst = a.wbuf[pass & 1];

        // Stages tile [tbeg, tbeg + tile_n) in key_in / id_in (one bulk copy each; the <= 3 items beyond the last 16-byte
        // multiple by plain loads), gathers the weights (last pass), and ranks every item inside its warp by digit:
        // afterwards wcnt[w][d] = items of digit d in warp w's span, rank[loc] = the item's rank among them.
        auto stage_and_rank = [&](int64_t tbeg, int tile_n) {
            __syncthreads();                                 // every reader of the previous tile is done
            const int bulk_n = tile_n & ~3;
            if (tid == 0) {
                asm volatile("fence.proxy.async;" ::: "memory");      // sources were written through the generic proxy (previous pass)
                ds_mbar_expect(&s.mbar, (uint32_t)(bulk_n * 4 * (first ? 2 : 3)));
                if (bulk_n) {
                    ds_bulk_load(s.key_in, ksrc + tbeg, (uint32_t)(bulk_n * 4), &s.mbar);
                    ds_bulk_load(s.wgt_in, wsrc + tbeg, (uint32_t)(bulk_n * 4), &s.mbar);
                    if (!first) ds_bulk_load(s.id_in, isrc + tbeg, (uint32_t)(bulk_n * 4), &s.mbar);
                }
            }
            if (tid < 3 && bulk_n + tid < tile_n) {          // ragged tail of the very last tile
                s.key_in[bulk_n + tid] = ksrc[tbeg + bulk_n + tid];
                s.wgt_in[bulk_n + tid] = wsrc[tbeg + bulk_n + tid];
                if (!first) s.id_in[bulk_n + tid] = isrc[tbeg + bulk_n + tid];
            }
#pragma unroll
            for (int q = 0; q < kRadix / 32; ++q) s.wcnt[wid][q * 32 + lane] = 0u;
            ds_mbar_wait(&s.mbar, tma_parity);
            tma_parity ^= 1u;
            __syncthreads();                                 // tail items and counters visible
#pragma unroll 2
            for (int k = 0; k < kSortItems; ++k) {
                const int loc = wid * kSortWarpSpan + k * 32 + lane;
                const bool valid = loc < tile_n;
                uint32_t key = valid ? s.key_in[loc] : 0u;
                if (first && valid) { key = remap_key(key, kmin, span); s.key_in[loc] = key; }
                const uint32_t d = (key >> shift) & 255u;
                const unsigned peers = match_digit(d, valid);
                const uint32_t prev = valid ? s.wcnt[wid][d] : 0u;
                __syncwarp();
                if (valid && lane == (__ffs(peers) - 1)) s.wcnt[wid][d] = prev + (uint32_t)__popc(peers);
                __syncwarp();
                if (valid) s.rank[loc] = (uint16_t)(prev + (uint32_t)__popc(peers & ((1u << lane) - 1u)));
            }
            __syncthreads();
        };

        // Sorts the staged tile by the current digit inside shared memory (stable): per-warp counts -> exclusive bases,
        // tile-local start of every digit run, items moved to key_out / id_out; in the last pass also the exclusive prefix sum
        // of tiles_touched over the sorted tile (wscan), from which the tile's per-digit weight sums follow by differences.
        auto local_sort = [&](int64_t tbeg, int tile_n) {
            // per-warp counts -> per-warp exclusive bases, tile counts, tile-local start of every digit run
            uint32_t cnt = 0u;
            if (tid < kRadix) {
#pragma unroll
                for (int w = 0; w < kSortWarps; ++w) {
                    const uint32_t v = s.wcnt[w][tid];
                    s.wcnt[w][tid] = cnt;
                    cnt += v;
                }
                s.tcount[tid] = cnt;
            }
            const uint32_t ex = scan256(tid < kRadix ? cnt : 0u, s.scan_tmp, tid, nullptr);
            if (tid < kRadix) s.texcl[tid] = ex;
            if (tid == 0) s.texcl[kRadix] = (uint32_t)tile_n;
            __syncthreads();
#pragma unroll 2
            for (int k = 0; k < kSortItems; ++k) {
                const int loc = wid * kSortWarpSpan + k * 32 + lane;
                if (loc < tile_n) {
                    const uint32_t key = s.key_in[loc];
                    const uint32_t d = (key >> shift) & 255u;
                    const uint32_t dst = s.texcl[d] + s.wcnt[wid][d] + (uint32_t)s.rank[loc];
                    s.key_out[dst] = key;
                    s.id_out[dst] = first ? (int)(tbeg + loc) : s.id_in[loc];
                    s.wscan[dst] = s.wgt_in[loc];             // weights travel with the items; the last pass scans them
                }
            }
            __syncthreads();
            if (last) {
                // exclusive prefix sum of tiles_touched over the tile in its sorted order
                uint32_t w[kSortItems], sum = 0u;
#pragma unroll
                for (int k = 0; k < kSortItems; ++k) {
                    const int loc = tid * kSortItems + k;
                    w[k] = loc < tile_n ? s.wscan[loc] : 0u;
                    sum += w[k];
                }
                uint32_t inc = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t x = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += x;
                }
                if (lane == 31) s.scan_tmp[wid] = inc;
                __syncthreads();
                uint32_t base = 0u;
#pragma unroll
                for (int q = 0; q < kSortWarps; ++q) base += (q < wid) ? s.scan_tmp[q] : 0u;
                uint32_t run = base + inc - sum;
#pragma unroll
                for (int k = 0; k < kSortItems; ++k) {
                    const int loc = tid * kSortItems + k;
                    if (loc <= tile_n) s.wscan[loc] = run;          // wscan[tile_n] = the tile's total
                    run += w[k];
                }
                if (tid == kSortThreads - 1) s.wscan[kSortTile] = run;   // full tile: no thread's range reaches loc == tile_n
                __syncthreads();
            }
        };

        GS_STAMP(3 + pass * 5);
        // ---- phase A: digit histogram (and tiles_touched sums) of this CTA's chunk -> table row ----------------
        if (tid < kRadix) { s.hist[tid] = 0u; s.whist[tid] = 0u; }
        for (int64_t tbeg = begin; tbeg < end; tbeg += kSortTile) {
            const int tile_n = (int)min((int64_t)kSortTile, end - tbeg);
            stage_and_rank(tbeg, tile_n);
            if (last || single_tile) {
                // the tile is sorted here already: in the last pass its per-digit tiles_touched sums are differences of the
                // prefix sum over the sorted tile (no atomics); a single-tile chunk keeps the sorted tile for the scatter
                local_sort(tbeg, tile_n);
                if (tid < kRadix) {
                    s.hist[tid] += s.tcount[tid];
                    if (last) s.whist[tid] += s.wscan[s.texcl[tid] + s.tcount[tid]] - s.wscan[s.texcl[tid]];
                }
            } else if (tid < kRadix) {
                uint32_t c = 0u;
#pragma unroll
                for (int w = 0; w < kSortWarps; ++w) c += s.wcnt[w][tid];
                s.hist[tid] += c;
            }
        }
        __syncthreads();
        if (tid < kRadix) {
            a.table[cta * kRadix + tid] = s.hist[tid];
            if (last) a.wtable[cta * kRadix + tid] = (unsigned long long)s.whist[tid];
        }
        GS_STAMP(4 + pass * 5);
        grid.sync();
        GS_STAMP(5 + pass * 5);

        // ---- scan: where does this CTA's run of digit d start?  (all smaller digits) + (digit d in earlier CTAs) ----
        {
            // counts: thread t takes digits 4q..4q+3 (q = t & 63, one 16-byte load per table row) of the rows c = g, g+8, ...
            // (g = t >> 6), eight independent loads in flight; the eight row groups meet in shared memory
            uint32_t before = 0u, total = 0u;
            unsigned long long wbefore = 0ull, wtotal = 0ull;
            {
                const int q = tid & 63, g = tid >> 6;
                uint4 bsum = make_uint4(0u, 0u, 0u, 0u), tsum = make_uint4(0u, 0u, 0u, 0u);
                const uint4* tab = reinterpret_cast<const uint4*>(a.table) + q;
#pragma unroll 8
                for (int c = g; c < G; c += kSortThreads / 64) {
                    const uint4 v = tab[c * (kRadix / 4)];
                    tsum.x += v.x; tsum.y += v.y; tsum.z += v.z; tsum.w += v.w;
                    if (c < cta) { bsum.x += v.x; bsum.y += v.y; bsum.z += v.z; bsum.w += v.w; }
                }
                uint4* redc = reinterpret_cast<uint4*>(s.key_in);                // [8 groups][64 quads][before, total] (key_in / id_in are free here)
                redc[(g * 64 + q) * 2 + 0] = bsum;
                redc[(g * 64 + q) * 2 + 1] = tsum;
                unsigned long long* redw = reinterpret_cast<unsigned long long*>(s.id_in);    // [4 groups][128 pairs][before x2, total x2]
                if (last) {
                    const int q2 = tid & 127, g2 = tid >> 7;
                    unsigned long long b0 = 0ull, b1 = 0ull, t0 = 0ull, t1 = 0ull;
                    const ulonglong2* wtab = reinterpret_cast<const ulonglong2*>(a.wtable) + q2;
#pragma unroll 8
                    for (int c = g2; c < G; c += kSortThreads / 128) {
                        const ulonglong2 v = wtab[c * (kRadix / 2)];
                        t0 += v.x; t1 += v.y;
                        if (c < cta) { b0 += v.x; b1 += v.y; }
                    }
                    unsigned long long* dstw = redw + (g2 * 128 + q2) * 4;
                    dstw[0] = b0; dstw[1] = b1; dstw[2] = t0; dstw[3] = t1;
                }
                __syncthreads();
                if (tid < kRadix) {
                    const uint32_t* rc = reinterpret_cast<const uint32_t*>(s.key_in);
#pragma unroll
                    for (int gg = 0; gg < kSortThreads / 64; ++gg) {
                        before += rc[((gg * 64 + (tid >> 2)) * 2 + 0) * 4 + (tid & 3)];
                        total += rc[((gg * 64 + (tid >> 2)) * 2 + 1) * 4 + (tid & 3)];
                    }
                    if (last) {
#pragma unroll
                        for (int gg = 0; gg < kSortThreads / 128; ++gg) {
                            wbefore += redw[(gg * 128 + (tid >> 1)) * 4 + (tid & 1)];
                            wtotal += redw[(gg * 128 + (tid >> 1)) * 4 + 2 + (tid & 1)];
                        }
                    }
                }
                __syncthreads();
            }
            const uint32_t dbase = scan256(tid < kRadix ? total : 0u, s.scan_tmp, tid, nullptr);
            if (tid < kRadix) s.run[tid] = dbase + before;
            if (last) {
                unsigned long long d_total = 0ull;
                const unsigned long long wbase = scan256_64(tid < kRadix ? wtotal : 0ull, s.scan_tmp64, tid, &d_total);
                if (tid < kRadix) s.wrun[tid] = wbase + wbefore;
                if (cta == 0 && tid == 0) {
                    a.counters[0] = (int64_t)((unsigned long long)a.n - n_fe - n_ff);
                    a.counters[1] = (int64_t)d_total;
                    a.counters[2] = (int64_t)((unsigned long long)a.n - n_ff);
                }
            }
            __syncthreads();
        }

        GS_STAMP(6 + pass * 5);
        // ---- phase B: stable scatter of the chunk, tile by tile ---------------------------------------------
        for (int64_t tbeg = begin; tbeg < end; tbeg += kSortTile) {
            const int tile_n = (int)min((int64_t)kSortTile, end - tbeg);
            if (!single_tile) {
                stage_and_rank(tbeg, tile_n);
                local_sort(tbeg, tile_n);
            }
            GS_STAMP(34 + pass * 4);
#pragma unroll 2
            for (int k = 0; k < kSortItems; ++k) {
                const int loc = k * kSortThreads + tid;
                if (loc < tile_n) {
                    const uint32_t kk = s.key_out[loc];
                    const uint32_t d = (kk >> shift) & 255u;
                    const uint32_t first_of_run = s.texcl[d];
                    const int64_t pos = (int64_t)s.run[d] + (loc - (int)first_of_run);
                    if (last) {
                        a.sorted_ids[pos] = s.id_out[loc];
                        a.offsets[pos] = (int64_t)(s.wrun[d] + (unsigned long long)(s.wscan[loc] - s.wscan[first_of_run]));
                    } else {
                        kdst[pos] = kk;
                        idst[pos] = s.id_out[loc];
                        wdst[pos] = s.wscan[loc];
                    }
                }
            }
            __syncthreads();
            if (tid < kRadix) {
                const uint32_t c = s.tcount[tid];
                if (last) s.wrun[tid] += (unsigned long long)(s.wscan[s.texcl[tid] + c] - s.wscan[s.texcl[tid]]);
                s.run[tid] += c;
            }
        }
        GS_STAMP(7 + pass * 5);
        if (!last) grid.sync();
    }
    GS_STAMP(3 + 4 * 5);
}

int64_t depth_sort_workspace_bytes(int64_t n) {
    auto up = [](int64_t x) { return (x + 255) / 256 * 256; };
    return 6 * up(n * 4) + up((int64_t)kMaxSortGrid * kRadix * 4) + up((int64_t)kMaxSortGrid * kRadix * 8) + up((int64_t)kMaxSortGrid * 16) + 1024;
}

// Enqueues the sort; returns a GsStatus.
int depth_sort_launch(int64_t n, const uint32_t* depth_keys, const int32_t* tiles_touched, void* workspace, int64_t workspace_bytes,
                      int32_t* sorted_ids, int64_t* offsets, int64_t* counters, cudaStream_t st) {
    auto up = [](int64_t x) { return (x + 255) / 256 * 256; };
    if (workspace_bytes < depth_sort_workspace_bytes(n)) {
        set_error("gs_bin_prepare: workspace %lld B < required %lld B", (long long)workspace_bytes, (long long)depth_sort_workspace_bytes(n));
        return GS_ERR_WORKSPACE_TOO_SMALL;
    }
    if ((reinterpret_cast<uintptr_t>(depth_keys) | reinterpret_cast<uintptr_t>(tiles_touched) | reinterpret_cast<uintptr_t>(workspace)) % 16 != 0) {
        set_error("gs_bin_prepare: depth_keys, tiles_touched and workspace must be 16-byte aligned");
        return GS_ERR_INVALID_ARGUMENT;
    }
    static thread_local int cached_dev = -1, cached_grid = 0;
    int dev = 0;
    GS_CUDA_TRY(cudaGetDevice(&dev));
    const size_t smem = sizeof(SortSmem);
    if (dev != cached_dev) {
        int sms = 0, per_sm = 0, coop = 0;
        GS_CUDA_TRY(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
        if (!coop) {
            set_error("gs_bin_prepare: device %d does not support cooperative launches", dev);
            return GS_ERR_UNSUPPORTED;
        }
        GS_CUDA_TRY(cudaFuncSetAttribute(depth_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        GS_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        GS_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, depth_sort_kernel, kSortThreads, smem));
        if (per_sm < 1) {
            set_error("gs_bin_prepare: depth_sort_kernel does not fit an SM (%zu B of shared memory)", smem);
            return GS_ERR_UNSUPPORTED;
        }
        cached_grid = sms < kMaxSortGrid ? sms : kMaxSortGrid;        // one CTA per SM: all co-resident
        cached_dev = dev;
    }
    char* ws = (char*)workspace;
    DepthSortArgs a;
    a.n = n;
    // a CTA's chunk is at least one full tile, so small inputs occupy few CTAs (the others only join the barriers)
    int64_t chunk = (n + cached_grid - 1) / cached_grid;
    if (chunk < kSortTile / 4) chunk = kSortTile / 4;
    chunk = (chunk + 3) / 4 * 4;                  // tiles start on 16-byte boundaries (bulk copies)
    a.chunk = chunk;
    a.keys_in = depth_keys;
    a.tiles_touched = tiles_touched;
    int64_t o = 0;
    a.kbuf[0] = (uint32_t*)(ws + o); o += up(n * 4);
    a.kbuf[1] = (uint32_t*)(ws + o); o += up(n * 4);
    a.ibuf[0] = (int32_t*)(ws + o); o += up(n * 4);
    a.ibuf[1] = (int32_t*)(ws + o); o += up(n * 4);
    a.wbuf[0] = (uint32_t*)(ws + o); o += up(n * 4);
    a.wbuf[1] = (uint32_t*)(ws + o); o += up(n * 4);
    a.table = (uint32_t*)(ws + o); o += up((int64_t)kMaxSortGrid * kRadix * 4);
    a.wtable = (unsigned long long*)(ws + o); o += up((int64_t)kMaxSortGrid * kRadix * 8);
    a.partial = (uint32_t*)(ws + o);
    a.sorted_ids = sorted_ids;
    a.offsets = offsets;
    a.counters = counters;
    void* args[] = {&a};
    GS_CUDA_TRY(cudaLaunchCooperativeKernel((void*)depth_sort_kernel, dim3(cached_grid), dim3(kSortThreads), args, smem, st));
    count_launches(1);
    return GS_OK;
}

}  // namespace gs
