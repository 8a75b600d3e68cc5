// Stage S+B: global depth order, per-tile duplication, stable tile sort, tile ranges.
//
// Reference semantics: src/core/renderer.py:222-239 (one global depth sort of the visible
// splats) and :263-298 (each splat, in that order, appended to every tile its AABB touches).
//
// The reference order equals a stable sort of (tile_id<<32 | depth_bits) keys.  Sorting 64-bit
// keys for every tile pair would move ~24 B x 6 passes per pair; instead the two key halves are
// sorted where they are cheap: depth on the N splats (32-bit keys, N items), then the tile id
// alone on the D pairs (ceil(log2 tiles) bits -> 2 radix passes of 4+4 B).  Same sequence, about
// a fifth of the HBM traffic.  All kernels here are HBM-bound streams.
#include "common.cuh"

#include <cub/cub.cuh>

namespace gs {

static inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

struct GatherCount {
    const int32_t* tiles_touched;
    const int32_t* sorted_ids;
    __host__ __device__ __forceinline__ int64_t operator()(int64_t j) const {
        return (int64_t)tiles_touched[sorted_ids[j]];
    }
};

__global__ void iota_kernel(int64_t n, int32_t* ids) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) ids[i] = (int32_t)i;
}

__device__ __forceinline__ int64_t lower_bound_u32(const uint32_t* a, int64_t n, uint32_t v) {
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (a[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// counters[0] = splats with >=1 tile, counters[1] = total tile pairs, counters[2] = visible splats
__global__ void counters_kernel(int64_t n, const uint32_t* sorted_keys, const int32_t* sorted_ids,
                                const int32_t* tiles_touched, const int64_t* offsets, int64_t* counters) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const int64_t v_tiles = lower_bound_u32(sorted_keys, n, 0xFFFFFFFEu);
        const int64_t v_vis = lower_bound_u32(sorted_keys, n, 0xFFFFFFFFu);
        counters[0] = v_tiles;
        counters[1] = n > 0 ? offsets[n - 1] + (int64_t)tiles_touched[sorted_ids[n - 1]] : 0;
        counters[2] = v_vis;
    }
}

// One warp per 32 consecutive depth ranks; for each rank the warp writes that splat's tile ids
// side by side (coalesced), in row-major tile order.
__global__ void __launch_bounds__(256)
duplicate_kernel(int64_t num_sorted, const int32_t* __restrict__ sorted_ids, const int64_t* __restrict__ offsets,
                 const ushort4* __restrict__ tile_rect, int tiles_x,
                 uint32_t* __restrict__ tile_keys, int32_t* __restrict__ vals) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t j0 = warp * 32;
    if (j0 >= num_sorted) return;
    const int64_t j = j0 + lane;
    int id = 0;
    long long off = 0;
    int tx0 = 0, ty0 = 0, w = 0, cnt = 0;
    if (j < num_sorted) {
        id = sorted_ids[j];
        off = offsets[j];
        const ushort4 r = tile_rect[id];
        tx0 = r.x; ty0 = r.y;
        w = (int)r.z - (int)r.x + 1;
        cnt = w * ((int)r.w - (int)r.y + 1);
    }
    const int limit = (int)min((int64_t)32, num_sorted - j0);
    for (int l = 0; l < limit; ++l) {
        const int s_id = __shfl_sync(0xffffffffu, id, l);
        const long long s_off = __shfl_sync(0xffffffffu, off, l);
        const int s_tx0 = __shfl_sync(0xffffffffu, tx0, l);
        const int s_ty0 = __shfl_sync(0xffffffffu, ty0, l);
        const int s_w = __shfl_sync(0xffffffffu, w, l);
        const int s_cnt = __shfl_sync(0xffffffffu, cnt, l);
        for (int k = lane; k < s_cnt; k += 32) {
            const int ty = s_ty0 + k / s_w;
            const int tx = s_tx0 + k % s_w;
            tile_keys[s_off + k] = (uint32_t)(ty * tiles_x + tx);
            vals[s_off + k] = s_id;
        }
    }
}

__global__ void __launch_bounds__(256)
ranges_kernel(int64_t d, const uint32_t* __restrict__ sorted_tile_keys, int32_t* __restrict__ ranges) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d) return;
    const uint32_t t = sorted_tile_keys[i];
    if (i == 0) {
        ranges[2 * t] = 0;
    } else {
        const uint32_t prev = sorted_tile_keys[i - 1];
        if (prev != t) {
            ranges[2 * prev + 1] = (int32_t)i;
            ranges[2 * t] = (int32_t)i;
        }
    }
    if (i == d - 1) ranges[2 * t + 1] = (int32_t)d;
}

__global__ void __launch_bounds__(256)
entry_keys_kernel(int64_t d, const uint32_t* __restrict__ sorted_tile_keys, const int32_t* __restrict__ entry_ids,
                  const uint32_t* __restrict__ depth_keys, uint64_t* __restrict__ entry_keys) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d) return;
    entry_keys[i] = ((uint64_t)sorted_tile_keys[i] << 32) | (uint64_t)depth_keys[entry_ids[i]];
}

// ------------------------------------------------------------------------------------------------
// Stable counting sort on the tile id (default path).
//
// The pairs are produced in depth order, so grouping them by tile while KEEPING that order is all
// the second sort has to do.  Depth ranks are cut into chunks of kChunk consecutive ranks; one warp
// walks its chunk in rank order with a shared-memory counter per tile, which gives every pair its
// position among the chunk's pairs of the same tile (`local`) and, at the end, the chunk's row of the
// [chunks x tiles] count table.  A column scan over chunks plus a scan over tiles turns the table
// into base offsets, and a fully parallel pass scatters:  pos = tile_start[t] + base[chunk][t] + local.
// Order within a tile is (chunk, local) = depth rank: stable by construction, no atomics, no
// comparison of keys.  Traffic ~ 6 B/pair + the table, against ~36 B/pair for two radix passes.
// ------------------------------------------------------------------------------------------------
constexpr int kChunk = 512;            // depth ranks per chunk (one warp); counts fit uint16

// A splat's tile rectangle packed for warp broadcast: origin tile index, width, tile count, and the
// reciprocal used to split k into (row, col) without an integer division in the inner loop.
struct RectDesc {
    int origin;        // ty0 * tiles_x + tx0
    int w_cnt;         // width | count << 12   (count <= 4095 on this path, else `big`)
    unsigned inv;      // 65536 / w + 1
};
constexpr int kMaxFastCount = 4095;

__device__ __forceinline__ RectDesc make_desc(const ushort4 r, int tiles_x, int& cnt_out) {
    RectDesc d;
    const int w = (int)r.z - (int)r.x + 1;
    const int cnt = w * ((int)r.w - (int)r.y + 1);
    d.origin = (int)r.y * tiles_x + (int)r.x;
    d.w_cnt = w | (min(cnt, kMaxFastCount) << 12);
    d.inv = 65536u / (unsigned)w + 1u;
    cnt_out = cnt;
    return d;
}

// tile index of the k-th tile (row-major inside the rectangle); exact for k <= 4095:
// (k*inv)>>16 over-estimates k/w by < 1/16, so it is at most one too large.
__device__ __forceinline__ int tile_of(int k, int origin, int w, unsigned inv, int tiles_x) {
    int row = (int)(((unsigned)k * inv) >> 16);
    row -= (row * w > k);
    return origin + row * tiles_x + (k - row * w);
}

__global__ void __launch_bounds__(32)
chunk_count_kernel(int64_t num_sorted, const int32_t* __restrict__ sorted_ids, const int64_t* __restrict__ offsets,
                   const ushort4* __restrict__ tile_rect, int tiles_x, int num_tiles,
                   uint16_t* __restrict__ local_pos, uint16_t* __restrict__ counts /* [chunks][tiles] */) {
    extern __shared__ uint16_t s_cnt[];
    const int lane = threadIdx.x;
    const int64_t chunk = blockIdx.x;
    for (int t = lane; t < num_tiles; t += 32) s_cnt[t] = 0;
    __syncwarp();
    const int64_t j_begin = chunk * kChunk;
    const int64_t j_end = min(j_begin + (int64_t)kChunk, num_sorted);
    const int64_t off_base = offsets[j_begin];
    uint16_t* lp = local_pos + off_base;
    for (int64_t j0 = j_begin; j0 < j_end; j0 += 32) {
        const int64_t j = j0 + lane;
        int off = 0, cnt = 0;
        RectDesc d = {0, 1, 65537u};
        if (j < j_end) {
            off = (int)(offsets[j] - off_base);
            d = make_desc(tile_rect[sorted_ids[j]], tiles_x, cnt);
        }
        const int limit = (int)min((int64_t)32, j_end - j0);
#pragma unroll 4
        for (int l = 0; l < limit; ++l) {
            const int s_off = __shfl_sync(0xffffffffu, off, l);
            const int s_origin = __shfl_sync(0xffffffffu, d.origin, l);
            const int s_wc = __shfl_sync(0xffffffffu, d.w_cnt, l);
            const unsigned s_inv = __shfl_sync(0xffffffffu, d.inv, l);
            const int s_w = s_wc & 0xfff, s_cnt_l = s_wc >> 12;
            if (s_cnt_l < kMaxFastCount) {
                for (int k = lane; k < s_cnt_l; k += 32) {          // a splat's tiles are distinct: no conflicts
                    const int t = tile_of(k, s_origin, s_w, s_inv, tiles_x);
                    const uint16_t c = s_cnt[t];
                    s_cnt[t] = (uint16_t)(c + 1);
                    lp[s_off + k] = c;
                }
            } else {                                                 // huge rectangle: exact division
                const int full = __shfl_sync(0xffffffffu, cnt, l);
                for (int k = lane; k < full; k += 32) {
                    const int row = k / s_w;
                    const int t = s_origin + row * tiles_x + (k - row * s_w);
                    const uint16_t c = s_cnt[t];
                    s_cnt[t] = (uint16_t)(c + 1);
                    lp[s_off + k] = c;
                }
            }
            __syncwarp();                                            // next splat (next depth rank) sees these counts
        }
    }
    uint16_t* row = counts + chunk * (int64_t)num_tiles;
    for (int t = lane; t < num_tiles; t += 32) row[t] = s_cnt[t];
}

// Exclusive scan over chunks for every tile.  Block = 32 tiles x 32 contiguous chunk segments.
__global__ void __launch_bounds__(1024)
column_scan_kernel(int num_chunks, int num_tiles, const uint16_t* __restrict__ counts, uint32_t* __restrict__ base,
                   uint32_t* __restrict__ tile_total) {
    __shared__ uint32_t s_sum[32][33];
    const int tl = threadIdx.x, seg = threadIdx.y;
    const int t = blockIdx.x * 32 + tl;
    const int per = (num_chunks + 31) / 32;
    const int c0 = seg * per, c1 = min(c0 + per, num_chunks);
    uint32_t sum = 0;
    if (t < num_tiles)
        for (int c = c0; c < c1; ++c) sum += counts[(int64_t)c * num_tiles + t];
    s_sum[seg][tl] = sum;
    __syncthreads();
    uint32_t run = 0;
    for (int s2 = 0; s2 < seg; ++s2) run += s_sum[s2][tl];
    if (t < num_tiles) {
        if (seg == 31) tile_total[t] = run + sum;
        for (int c = c0; c < c1; ++c) {
            const int64_t at = (int64_t)c * num_tiles + t;
            base[at] = run;
            run += counts[at];
        }
    }
}

// Exclusive scan over tiles -> tile_ranges [begin,end) and tile_start.  One block of 1024 threads:
// per-thread serial sums, warp-shuffle scan, one cross-warp scan.
__global__ void __launch_bounds__(1024)
tile_scan_kernel(int num_tiles, const uint32_t* __restrict__ tile_total, uint32_t* __restrict__ tile_start,
                 int32_t* __restrict__ ranges) {
    __shared__ uint32_t s_warp[32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int per = (num_tiles + 1023) / 1024;
    const int t0 = tid * per, t1 = min(t0 + per, num_tiles);
    uint32_t sum = 0;
    for (int t = t0; t < t1; ++t) sum += tile_total[t];
    uint32_t inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
    }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint32_t w = s_warp[lane], winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += v;
        }
        s_warp[lane] = winc - w;            // exclusive prefix of the warp totals
    }
    __syncthreads();
    uint32_t run = s_warp[wid] + inc - sum;
    for (int t = t0; t < t1; ++t) {
        const uint32_t c = tile_total[t];
        tile_start[t] = run;
        ranges[2 * t] = c ? (int32_t)run : 0;      // empty tiles report (0,0), as the radix path does
        ranges[2 * t + 1] = c ? (int32_t)(run + c) : 0;
        run += c;
    }
}

// One thread per depth rank: its tiles' loads are independent of each other, so they pipeline.
__global__ void __launch_bounds__(256)
scatter_kernel(int64_t num_sorted, const int32_t* __restrict__ sorted_ids, const int64_t* __restrict__ offsets,
               const ushort4* __restrict__ tile_rect, int tiles_x, int num_tiles,
               const uint16_t* __restrict__ local_pos, const uint32_t* __restrict__ base,
               const uint32_t* __restrict__ tile_start, const uint32_t* __restrict__ depth_keys,
               int32_t* __restrict__ entry_ids, uint64_t* __restrict__ entry_keys) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= num_sorted) return;
    const int id = sorted_ids[j];
    const int64_t off = offsets[j];
    const ushort4 r = tile_rect[id];
    const uint32_t* base_row = base + (j / kChunk) * (int64_t)num_tiles;
    const uint16_t* lp = local_pos + off;
    const uint64_t dkey = entry_keys ? (uint64_t)depth_keys[id] : 0ull;
    int k = 0;
    for (int ty = r.y; ty <= (int)r.w; ++ty) {
        const int trow = ty * tiles_x;
#pragma unroll 4
        for (int tx = r.x; tx <= (int)r.z; ++tx, ++k) {
            const int t = trow + tx;
            const uint32_t pos = tile_start[t] + base_row[t] + (uint32_t)lp[k];
            entry_ids[pos] = id;
            if (entry_keys) entry_keys[pos] = ((uint64_t)(uint32_t)t << 32) | dkey;
        }
    }
}

struct CountLayout {
    int64_t counts, base, local_pos, tile_total, tile_start, total;
    int num_chunks;
};
static CountLayout count_layout(int64_t num_sorted_cap, int64_t d, int32_t num_tiles) {
    CountLayout L;
    L.num_chunks = (int)((num_sorted_cap + kChunk - 1) / kChunk);
    if (L.num_chunks < 1) L.num_chunks = 1;
    int64_t o = 0;
    L.counts = o;     o += align_up((int64_t)L.num_chunks * num_tiles * 2, 256);
    L.base = o;       o += align_up((int64_t)L.num_chunks * num_tiles * 4, 256);
    L.local_pos = o;  o += align_up(d * 2, 256);
    L.tile_total = o; o += align_up((int64_t)num_tiles * 4, 256);
    L.tile_start = o; o += align_up((int64_t)num_tiles * 4, 256);
    L.total = o;
    return L;
}
constexpr int kMaxCountingTiles = 100000;      // 2 B of shared memory per tile for the chunk counters

static int tile_bits(int32_t num_tiles) {
    int bits = 1;
    while ((1ll << bits) < (long long)num_tiles) ++bits;
    return bits;
}

struct PrepareLayout {
    int64_t keys_sorted, ids_iota, cub_temp, cub_bytes, total;
};
static PrepareLayout prepare_layout(int64_t n) {
    PrepareLayout L;
    size_t sort_bytes = 0, scan_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                    (const int32_t*)nullptr, (int32_t*)nullptr, n, 0, 32);
    GatherCount op{nullptr, nullptr};
    cub::TransformInputIterator<int64_t, GatherCount, cub::CountingInputIterator<int64_t>> it(
        cub::CountingInputIterator<int64_t>(0), op);
    cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, it, (int64_t*)nullptr, n);
    int64_t o = 0;
    L.keys_sorted = o; o += align_up(n * 4, 256);
    L.ids_iota = o;    o += align_up(n * 4, 256);
    L.cub_temp = o;
    L.cub_bytes = align_up((int64_t)(sort_bytes > scan_bytes ? sort_bytes : scan_bytes), 256);
    o += L.cub_bytes;
    L.total = o;
    return L;
}

struct SortLayout {
    int64_t keys_in, keys_out, vals_in, cub_temp, cub_bytes, total;
};
static SortLayout sort_layout(int64_t d, int32_t num_tiles) {
    SortLayout L;
    size_t sort_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                    (const int32_t*)nullptr, (int32_t*)nullptr, d, 0, tile_bits(num_tiles));
    int64_t o = 0;
    L.keys_in = o;  o += align_up(d * 4, 256);
    L.keys_out = o; o += align_up(d * 4, 256);
    L.vals_in = o;  o += align_up(d * 4, 256);
    L.cub_temp = o;
    L.cub_bytes = align_up((int64_t)sort_bytes, 256);
    o += L.cub_bytes;
    L.total = o;
    return L;
}

}  // namespace gs

using namespace gs;

extern "C" int64_t gs_bin_workspace_bytes(int64_t n, int64_t d_capacity, int32_t num_tiles) {
    if (n < 0 || d_capacity < 0 || num_tiles <= 0) return GS_ERR_INVALID_ARGUMENT;
    const int64_t a = prepare_layout(n > 0 ? n : 1).total;
    const int64_t b = sort_layout(d_capacity > 0 ? d_capacity : 1, num_tiles).total;
    const int64_t c = num_tiles <= kMaxCountingTiles ? count_layout(n > 0 ? n : 1, d_capacity > 0 ? d_capacity : 1, num_tiles).total : 0;
    int64_t m = a > b ? a : b;
    if (c > m) m = c;
    return m + 256;
}

extern "C" int gs_bin_prepare(int64_t n, const uint32_t* depth_keys, const int32_t* tiles_touched, void* workspace,
                              int64_t workspace_bytes, int32_t* sorted_ids, int64_t* offsets, int64_t* counters,
                              void* stream) {
    GS_REQUIRE(n >= 0, "n < 0");
    GS_REQUIRE(counters != nullptr, "counters is NULL");
    cudaStream_t st = (cudaStream_t)stream;
    DeviceGuard guard(counters);
    if (n == 0) {
        GS_CUDA_TRY(cudaMemsetAsync(counters, 0, 3 * sizeof(int64_t), st));
        return GS_OK;
    }
    GS_REQUIRE(n < (1ll << 31), "n must fit int32 ids");
    GS_REQUIRE(depth_keys && tiles_touched && workspace && sorted_ids && offsets, "NULL array argument");
    const PrepareLayout L = prepare_layout(n);
    if (workspace_bytes < L.total) {
        set_error("gs_bin_prepare: workspace %lld B < required %lld B", (long long)workspace_bytes, (long long)L.total);
        return GS_ERR_WORKSPACE_TOO_SMALL;
    }
    char* ws = (char*)workspace;
    uint32_t* keys_sorted = (uint32_t*)(ws + L.keys_sorted);
    int32_t* ids_iota = (int32_t*)(ws + L.ids_iota);
    void* cub_temp = ws + L.cub_temp;
    size_t cub_bytes = (size_t)L.cub_bytes;
    const int threads = 256;
    iota_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, st>>>(n, ids_iota);
    GS_CUDA_TRY(cudaGetLastError());
    // LSD radix sort is stable: equal depths keep ascending splat index.
    GS_CUDA_TRY(cub::DeviceRadixSort::SortPairs(cub_temp, cub_bytes, depth_keys, keys_sorted, (const int32_t*)ids_iota,
                                                sorted_ids, n, 0, 32, st));
    GatherCount op{tiles_touched, sorted_ids};
    cub::TransformInputIterator<int64_t, GatherCount, cub::CountingInputIterator<int64_t>> it(
        cub::CountingInputIterator<int64_t>(0), op);
    cub_bytes = (size_t)L.cub_bytes;
    GS_CUDA_TRY(cub::DeviceScan::ExclusiveSum(cub_temp, cub_bytes, it, offsets, n, st));
    counters_kernel<<<1, 32, 0, st>>>(n, keys_sorted, sorted_ids, tiles_touched, offsets, counters);
    GS_CUDA_TRY(cudaGetLastError());
    count_launches(2);   // iota + counters
    return GS_OK;
}

extern "C" int gs_bin_sort(int64_t n, int64_t num_sorted, int64_t d, const int32_t* sorted_ids, const int64_t* offsets,
                           const uint16_t* tile_rect, const uint32_t* depth_keys, int32_t tiles_x, int32_t num_tiles,
                           int32_t algo, void* workspace, int64_t workspace_bytes, int32_t* entry_ids,
                           int32_t* tile_ranges, uint64_t* entry_keys, void* stream) {
    GS_REQUIRE(n >= 0 && num_sorted >= 0 && num_sorted <= n && d >= 0, "bad sizes");
    GS_REQUIRE(num_tiles > 0 && tiles_x > 0, "bad tile grid");
    GS_REQUIRE(tile_ranges != nullptr, "tile_ranges is NULL");
    GS_REQUIRE(d < (1ll << 31), "tile pairs must fit int32 positions");
    cudaStream_t st = (cudaStream_t)stream;
    DeviceGuard guard(tile_ranges);
    GS_CUDA_TRY(cudaMemsetAsync(tile_ranges, 0, (size_t)num_tiles * 2 * sizeof(int32_t), st));
    if (d == 0 || num_sorted == 0) return GS_OK;
    GS_REQUIRE(sorted_ids && offsets && tile_rect && workspace && entry_ids, "NULL array argument");
    GS_REQUIRE(entry_keys == nullptr || depth_keys != nullptr, "entry_keys needs depth_keys");
    GS_REQUIRE(algo >= 0 && algo <= 2, "algo must be 0 (auto), 1 (counting) or 2 (radix)");
    const bool counting = algo == GS_BIN_COUNTING || (algo == GS_BIN_AUTO && num_tiles <= kMaxCountingTiles);
    if (counting) {
        if (num_tiles > kMaxCountingTiles) {
            set_error("gs_bin_sort: the counting sort supports at most %d tiles (got %d); use algo 0 or 2", kMaxCountingTiles, num_tiles);
            return GS_ERR_UNSUPPORTED;
        }
        const CountLayout C = count_layout(num_sorted, d, num_tiles);
        if (workspace_bytes < C.total) {
            set_error("gs_bin_sort: workspace %lld B < required %lld B", (long long)workspace_bytes, (long long)C.total);
            return GS_ERR_WORKSPACE_TOO_SMALL;
        }
        char* wsc = (char*)workspace;
        uint16_t* counts = (uint16_t*)(wsc + C.counts);
        uint32_t* base = (uint32_t*)(wsc + C.base);
        uint16_t* local_pos = (uint16_t*)(wsc + C.local_pos);
        uint32_t* tile_total = (uint32_t*)(wsc + C.tile_total);
        uint32_t* tile_start = (uint32_t*)(wsc + C.tile_start);
        const size_t smem = (size_t)num_tiles * sizeof(uint16_t);
        if (smem > 48 * 1024) GS_CUDA_TRY(cudaFuncSetAttribute(chunk_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        chunk_count_kernel<<<C.num_chunks, 32, smem, st>>>(num_sorted, sorted_ids, offsets, (const ushort4*)tile_rect, tiles_x,
                                                          num_tiles, local_pos, counts);
        GS_CUDA_TRY(cudaGetLastError());
        column_scan_kernel<<<(num_tiles + 31) / 32, dim3(32, 32), 0, st>>>(C.num_chunks, num_tiles, counts, base, tile_total);
        GS_CUDA_TRY(cudaGetLastError());
        tile_scan_kernel<<<1, 1024, 0, st>>>(num_tiles, tile_total, tile_start, tile_ranges);
        GS_CUDA_TRY(cudaGetLastError());
        scatter_kernel<<<(unsigned)((num_sorted + 255) / 256), 256, 0, st>>>(
            num_sorted, sorted_ids, offsets, (const ushort4*)tile_rect, tiles_x, num_tiles, local_pos, base, tile_start,
            depth_keys, entry_ids, entry_keys);
        GS_CUDA_TRY(cudaGetLastError());
        count_launches(4);
        return GS_OK;
    }
    const SortLayout L = sort_layout(d, num_tiles);
    if (workspace_bytes < L.total) {
        set_error("gs_bin_sort: workspace %lld B < required %lld B", (long long)workspace_bytes, (long long)L.total);
        return GS_ERR_WORKSPACE_TOO_SMALL;
    }
    char* ws = (char*)workspace;
    uint32_t* keys_in = (uint32_t*)(ws + L.keys_in);
    uint32_t* keys_out = (uint32_t*)(ws + L.keys_out);
    int32_t* vals_in = (int32_t*)(ws + L.vals_in);
    void* cub_temp = ws + L.cub_temp;
    size_t cub_bytes = (size_t)L.cub_bytes;
    const int threads = 256;
    {
        const int64_t warps = (num_sorted + 31) / 32;
        const int64_t blocks = (warps * 32 + threads - 1) / threads;
        duplicate_kernel<<<(unsigned)blocks, threads, 0, st>>>(num_sorted, sorted_ids, offsets, (const ushort4*)tile_rect,
                                                               tiles_x, keys_in, vals_in);
        GS_CUDA_TRY(cudaGetLastError());
    }
    GS_CUDA_TRY(cub::DeviceRadixSort::SortPairs(cub_temp, cub_bytes, (const uint32_t*)keys_in, keys_out,
                                                (const int32_t*)vals_in, entry_ids, d, 0, tile_bits(num_tiles), st));
    const unsigned blocks_d = (unsigned)((d + threads - 1) / threads);
    ranges_kernel<<<blocks_d, threads, 0, st>>>(d, keys_out, tile_ranges);
    GS_CUDA_TRY(cudaGetLastError());
    count_launches(2);   // duplicate + ranges
    if (entry_keys) {
        entry_keys_kernel<<<blocks_d, threads, 0, st>>>(d, keys_out, entry_ids, depth_keys, entry_keys);
        GS_CUDA_TRY(cudaGetLastError());
        count_launches(1);
    }
    return GS_OK;
}
