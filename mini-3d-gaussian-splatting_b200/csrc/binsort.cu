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

// Stage S lives in depthsort.cu (one cooperative kernel: key range, 8-bit LSD passes over the differing bits only,
// prefix sum of tiles_touched and the counters fused into the last pass).
int64_t depth_sort_workspace_bytes(int64_t n);
int depth_sort_launch(int64_t n, const uint32_t* depth_keys, const int32_t* tiles_touched, void* workspace, int64_t workspace_bytes,
                      int32_t* sorted_ids, int64_t* offsets, int64_t* counters, cudaStream_t st);

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
// The pairs exist only implicitly: depth rank j touches the tiles of tile_rect[sorted_ids[j]], and
//   pos(j, t) = tile_start[t] + #{ j' < j : rect(j') contains t }.
// Depth ranks are cut into chunks of kChunk consecutive ranks, chunks into super-chunks of kSuper.
//   1. chunk_walk_kernel    one warp per chunk walks its ranks IN ORDER with a uint8 running counter per
//                           tile in shared memory: local[pair] = pairs of the same tile earlier in the
//                           chunk (uint8, written coalesced), counts[chunk][t] (uint8 row) at the end
//   2. column_prefix_kernel base16[chunk][t] = pairs of tile t in earlier chunks of the same super-chunk
//                           (uint16), super_tot[super][t]
//   3. tile_tables_kernel   super_tab -> exclusive prefix over supers, the tiles' totals, the chunk from which each
//                           8x8-tile block is closed (truncated lists), then -- last CTA to finish -- tile_start /
//                           tile_ranges and the forward's tile order.  (tile_scan_kernel alone serves the blocked algo.)
//   4. scatter_kernel       fully parallel, in rank order (so the 4-byte stores of neighbouring list
//                           entries reach L2 close together and merge before they are written back):
//                           entry_ids[tile_start + super_tab + base16 + local] = id
// Order within a tile is (super, chunk, local) = depth rank: stable by construction, no comparison of
// keys, no atomics.  A fused walk+scatter (stores issued from the sequential walk) was measured and
// rejected: every chunk then writes its share of a list sector at an unrelated time, L2 evicts the
// partial sectors, and DRAM traffic grows from ~0.3 GB to ~0.95 GB (profiles/r1_v5_binning.md).
// ------------------------------------------------------------------------------------------------
constexpr int kChunk = 255;            // depth ranks per chunk: running counts fit uint8
constexpr int kSuper = 256;            // chunks per super-chunk: 65 280 ranks, prefix fits uint16
constexpr int kRowAlign = 128;         // table rows padded to 128 tiles (uchar4 / ushort4 / uint4 accesses)

// Sizes of the frame.  Exact mode: the host read the counters back and passes them by value (dev = NULL).
// Optimistic mode: the host passes CAPACITIES and the counters' device address; every kernel reads the
// actual sizes itself and does nothing at all if the pairs would not fit (the caller then repeats the
// stage in exact mode), so the launch does not wait for the read-back.
struct BinSizes {
    int64_t num_sorted;             // exact mode: value; optimistic: capacity (n)
    int64_t d_capacity;
    const int64_t* dev;             // counters {num_sorted, D, visible} or NULL
};
// Truncated tile lists (gs_bin_sort list_cap / gs_bin_complete).  On config[1] a tile consumes at most ~500 of
// the ~3 200 entries of its list before all its pixels saturate, and the scattered stores are what the
// binning costs: the first pass therefore stores only the first `limit` entries of every tile, the compositing
// kernel flags the (rare) tile that reaches the end of its stored prefix with pixels still alive, and a
// completion pass -- enqueued unconditionally, every CTA of which returns at once when nothing is flagged --
// stores the rest for the flagged tiles before those tiles are composited again.
struct ListCap {
    uint32_t limit;                 // first pass: store entries with in-tile index < limit
    const uint8_t* tile_flags;      // completion pass: tiles to complete
    const int32_t* flag_count;      // completion pass: number of flagged tiles (NULL in the first pass)
    const int32_t* close_chunk;     // first pass (optional): per 8x8-tile block, the first chunk from which every tile of
    int blocks_x;                   //   the block already holds `limit` entries -- later ranks skip the block entirely
};
constexpr int kCloseBlk = 8;
__device__ __forceinline__ bool resolve_sizes(const BinSizes& z, int64_t& num_sorted) {
    num_sorted = z.num_sorted;
    if (z.dev != nullptr) {
        num_sorted = z.dev[0];
        if (z.dev[1] > z.d_capacity) return false;
    }
    return true;
}
static_assert(kChunk <= 255 && (kSuper - 1) * kChunk <= 65535, "counter widths");

// A splat's tile rectangle packed for warp broadcast: origin tile index, width, tile count, and the
// reciprocal used to split k into (row, col) without an integer division in the inner loop.
struct RectDesc {
    int origin;        // ty0 * tiles_x + tx0
    int w_cnt;         // width | count << 12   (count capped at kMaxFastCount; larger rectangles take the slow loop)
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

// 1. ordered walk.  The warp takes its ranks 32 at a time (one rectangle descriptor per lane, ids and
// rectangles software-pipelined one and two blocks ahead) and passes them IN RANK ORDER through the
// shared-memory counters; lane k serves tiles k, k+32, ... of the current rectangle.  A splat's tiles
// are distinct, so a splat never conflicts with itself; __syncwarp orders consecutive splats.
__global__ void __launch_bounds__(32)
chunk_walk_kernel(BinSizes sizes, const int32_t* __restrict__ sorted_ids, const int64_t* __restrict__ offsets,
                  const ushort4* __restrict__ tile_rect, int tiles_x, int row_tiles,
                  uint8_t* __restrict__ local_pos, uint8_t* __restrict__ counts /* [chunks][row_tiles] */,
                  uint32_t* __restrict__ zero_words, int zero_count) {
    extern __shared__ uint32_t s_words[];
    uint8_t* s_cnt = reinterpret_cast<uint8_t*>(s_words);
    const int lane = threadIdx.x;
    const int64_t chunk = blockIdx.x;
    const int words = row_tiles >> 2;
    // the tile-table launch's control block (ticket + closing chunks, < 1 KB) is cleared here, two launches ahead of its
    // use, instead of by a memset between the kernels
    if (chunk == 0)
        for (int w = lane; w < zero_count; w += 32) zero_words[w] = 0u;
    int64_t num_sorted;
    if (!resolve_sizes(sizes, num_sorted)) return;
    const int64_t j_begin = chunk * kChunk;
    if (j_begin >= num_sorted) return;                 // optimistic mode launches chunks for all n splats
    for (int w = lane; w < words; w += 32) s_words[w] = 0u;
    __syncwarp();
    const int64_t j_end = min(j_begin + (int64_t)kChunk, num_sorted);
    const ushort4 kNoRect = make_ushort4(1, 1, 0, 0);                  // width 0 -> count 0
    int id_cur = (j_begin + lane < j_end) ? sorted_ids[j_begin + lane] : -1;
    int id_nxt = (j_begin + 32 + lane < j_end) ? sorted_ids[j_begin + 32 + lane] : -1;
    ushort4 r_cur = id_cur >= 0 ? tile_rect[id_cur] : kNoRect;
    const int64_t off_base = offsets[j_begin];
    int off_cur = (j_begin + lane < j_end) ? (int)(offsets[j_begin + lane] - off_base) : 0;
    uint8_t* lp = local_pos + off_base;
    for (int64_t j0 = j_begin; j0 < j_end; j0 += 32) {
        const ushort4 r_nxt = id_nxt >= 0 ? tile_rect[id_nxt] : kNoRect;
        const int id_nn = (j0 + 64 + lane < j_end) ? sorted_ids[j0 + 64 + lane] : -1;
        const int off_nxt = (j0 + 32 + lane < j_end) ? (int)(offsets[j0 + 32 + lane] - off_base) : 0;
        int cnt = 0;
        RectDesc d = {0, 1, 65537u};
        if (id_cur >= 0) d = make_desc(r_cur, tiles_x, cnt);
        const int limit = (int)min((int64_t)32, j_end - j0);
#pragma unroll 4
        for (int l = 0; l < limit; ++l) {
            const int s_off = __shfl_sync(0xffffffffu, off_cur, l);
            const int s_origin = __shfl_sync(0xffffffffu, d.origin, l);
            const int s_wc = __shfl_sync(0xffffffffu, d.w_cnt, l);
            const unsigned s_inv = __shfl_sync(0xffffffffu, d.inv, l);
            const int s_w = s_wc & 0xfff, s_n = s_wc >> 12;
            if (s_n < kMaxFastCount) {
                for (int k = lane; k < s_n; k += 32) {
                    const int t = tile_of(k, s_origin, s_w, s_inv, tiles_x);
                    const uint8_t c = s_cnt[t];
                    s_cnt[t] = (uint8_t)(c + 1);
                    lp[s_off + k] = c;
                }
            } else {                                                 // huge rectangle: exact division
                const int full = __shfl_sync(0xffffffffu, cnt, l);
                for (int k = lane; k < full; k += 32) {
                    const int row = k / s_w;
                    const int t = s_origin + row * tiles_x + (k - row * s_w);
                    const uint8_t c = s_cnt[t];
                    s_cnt[t] = (uint8_t)(c + 1);
                    lp[s_off + k] = c;
                }
            }
            __syncwarp();                                            // the next rank sees these counts
        }
        id_cur = id_nxt; r_cur = r_nxt; id_nxt = id_nn; off_cur = off_nxt;
    }
    uint32_t* row = reinterpret_cast<uint32_t*>(counts + chunk * row_tiles);
    for (int w = lane; w < words; w += 32) row[w] = s_words[w];
}

// 2. exclusive prefix over the chunks of one super-chunk, for 128 tiles per block.
// Block = 32 lanes (4 tiles each) x 32 segments of kSuper/32 chunks.
constexpr int kSegChunks = kSuper / 32;
__global__ void __launch_bounds__(1024)
column_prefix_kernel(BinSizes sizes, int row_tiles, const uint8_t* __restrict__ counts, uint16_t* __restrict__ base16,
                     uint32_t* __restrict__ super_tot /* [supers][row_tiles] */) {
    __shared__ uint4 s_seg[32][33];
    const int lane = threadIdx.x, seg = threadIdx.y;
    grid_dependency_wait();
    int64_t num_sorted;
    if (!resolve_sizes(sizes, num_sorted)) return;
    const int num_chunks = (int)((num_sorted + kChunk - 1) / kChunk);
    if ((int)blockIdx.y * kSuper >= num_chunks) return;
    const int t4 = (blockIdx.x * 32 + lane) * 4;
    const int sup = blockIdx.y;
    const int c0 = sup * kSuper + seg * kSegChunks;
    uchar4 v[kSegChunks];
    uint4 sum = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int q = 0; q < kSegChunks; ++q) {
        v[q] = make_uchar4(0, 0, 0, 0);
        if (c0 + q < num_chunks) v[q] = *reinterpret_cast<const uchar4*>(counts + (int64_t)(c0 + q) * row_tiles + t4);
        sum.x += v[q].x; sum.y += v[q].y; sum.z += v[q].z; sum.w += v[q].w;
    }
    s_seg[seg][lane] = sum;
    __syncthreads();
    uint4 run = make_uint4(0, 0, 0, 0);
    for (int s2 = 0; s2 < seg; ++s2) {
        const uint4 o = s_seg[s2][lane];
        run.x += o.x; run.y += o.y; run.z += o.z; run.w += o.w;
    }
#pragma unroll
    for (int q = 0; q < kSegChunks; ++q) {
        if (c0 + q < num_chunks)
            *reinterpret_cast<ushort4*>(base16 + (int64_t)(c0 + q) * row_tiles + t4) =
                make_ushort4((unsigned short)run.x, (unsigned short)run.y, (unsigned short)run.z, (unsigned short)run.w);
        run.x += v[q].x; run.y += v[q].y; run.z += v[q].z; run.w += v[q].w;
    }
    if (seg == 31) *reinterpret_cast<uint4*>(super_tot + (int64_t)sup * row_tiles + t4) = run;
}

// exclusive scan over tiles -> tile_start and tile_ranges [begin,end).  One block of 1024 threads (blocked algo; the
// flat counting sort does the same scan inside tile_tables_kernel).
__global__ void __launch_bounds__(1024)
tile_scan_kernel(BinSizes sizes, int num_tiles, const uint32_t* __restrict__ tile_total, uint32_t* __restrict__ tile_start,
                 int32_t* __restrict__ ranges) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    int64_t num_sorted_unused;
    if (!resolve_sizes(sizes, num_sorted_unused)) return;
    if (tid == 0) s_carry = 0u;
    __syncthreads();
    for (int base = 0; base < num_tiles; base += 1024) {
        const int t = base + tid;
        const uint32_t v = t < num_tiles ? tile_total[t] : 0u;
        uint32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t x = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += x;
        }
        if (lane == 31) s_warp[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            const uint32_t w = s_warp[lane];
            uint32_t winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t x = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= o) winc += x;
            }
            s_warp[lane] = winc - w;            // exclusive prefix of the warp totals
        }
        __syncthreads();
        const uint32_t begin = s_carry + s_warp[wid] + inc - v;
        if (t < num_tiles) {
            tile_start[t] = begin;
            ranges[2 * t] = v ? (int32_t)begin : 0;       // empty tiles report (0,0), as the radix path does
            ranges[2 * t + 1] = v ? (int32_t)(begin + v) : 0;
        }
        __syncthreads();
        if (tid == 1023) s_carry = begin + v;
        __syncthreads();
    }
}

__global__ void identity_order_kernel(int num_tiles, int32_t* __restrict__ tile_order) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < num_tiles) tile_order[t] = t;
}

// 3.  Super-chunk prefix + closing chunks + tile scan + the forward's longest-first tile order in ONE launch.
// These were four latency-bound launches of 6-10 us each (a fifth of gs_bin_sort).  Phase 1 runs on every CTA, one thread
// per tile: exclusive prefix over the super-chunks (in place), the tile's total, and -- truncated lists -- the chunk from
// which the tile's stored prefix is full: before(c, t) = super_tab + base16 is non-decreasing in the chunk index c, so a
// binary search per tile finds the first chunk with before >= limit (num_chunks if the list is shorter than the limit); the
// maximum over the tiles of an 8x8 block is the chunk from which the whole block is closed.  Depth ranks are sorted, so on
// config[1] ~60 % of the ranks meet closed blocks only and the scatter drops them after reading their rectangle.  The CTA that finishes last (ticket counter) then does the two
// single-CTA steps on the totals all CTAs have published: the exclusive scan over tiles (tile_start / tile_ranges) and the
// bucket sort of the tiles by list length (tile_order, heaviest first).
#ifndef GS_TABLES_THREADS
#define GS_TABLES_THREADS 256
#endif
constexpr int kTablesThreads = GS_TABLES_THREADS;
constexpr int kOrderBuckets2 = 256;
__global__ void __launch_bounds__(kTablesThreads)
tile_tables_kernel(BinSizes sizes, int num_tiles, int tiles_x, int row_tiles, const uint16_t* __restrict__ base16,
                   uint32_t* __restrict__ super_tab, uint32_t* __restrict__ tile_total, uint32_t* __restrict__ tile_start,
                   int32_t* __restrict__ ranges, uint32_t limit, int blocks_x, int32_t* __restrict__ close_chunk,
                   int32_t* __restrict__ tile_order, unsigned int* __restrict__ done_counter, int32_t* __restrict__ flag_count) {
    // truncated lists: this frame's count of tiles that need more than their stored prefix starts at zero (the
    // compositing pass that raises it is enqueued behind this launch)
    grid_dependency_wait();
    if (flag_count != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *flag_count = 0;
    __shared__ int s_cnt[kOrderBuckets2];
    __shared__ int s_off[kOrderBuckets2];
    __shared__ bool s_last;
    const int tid = threadIdx.x;
    const int t = blockIdx.x * blockDim.x + tid;
    int64_t num_sorted;
    if (!resolve_sizes(sizes, num_sorted)) {
        // capacity overflow: no list is written (tile_ranges stay zero) and the caller repeats the stage with exact sizes, but
        // the compositing launch already enqueued behind this one still reads the order: give it a valid one
        if (t < num_tiles) {
            reinterpret_cast<int2*>(ranges)[t] = make_int2(0, 0);
            if (tile_order != nullptr) tile_order[t] = t;
        }
        return;
    }
    const int num_chunks = (int)((num_sorted + kChunk - 1) / kChunk);
    const int num_supers = (num_chunks + kSuper - 1) / kSuper;
    if (t < row_tiles) {
        uint32_t run = 0u;
        // the super-chunk in which this tile's list reaches `limit` entries (truncated lists), noted while the prefixes are
        // still in registers: the search for the closing chunk below then stays inside one super-chunk
        int cross = -1;
        uint32_t cross_base = 0u;
        int sidx = 0;
        for (; sidx + 8 <= num_supers; sidx += 8) {
            uint32_t v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = super_tab[(int64_t)(sidx + q) * row_tiles + t];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                super_tab[(int64_t)(sidx + q) * row_tiles + t] = run;
                if (cross < 0 && run + v[q] >= limit) { cross = sidx + q; cross_base = run; }
                run += v[q];
            }
        }
        for (; sidx < num_supers; ++sidx) {
            const uint32_t v = super_tab[(int64_t)sidx * row_tiles + t];
            super_tab[(int64_t)sidx * row_tiles + t] = run;
            if (cross < 0 && run + v >= limit) { cross = sidx; cross_base = run; }
            run += v;
        }
        tile_total[t] = run;
        if (close_chunk != nullptr && t < num_tiles) {
            // first c in [0, num_chunks] with before(c, t) = super prefix + in-super prefix >= limit; num_chunks: never closed
            // (the whole list fits the stored prefix, or it fills up inside the last chunk).  before() is non-decreasing in c,
            // below `limit` at the first chunk of super-chunk `cross` and at or above it at the first chunk of the next one:
            // 8 dependent 2-byte loads instead of 12 pairs of loads over the whole column.
            int lo = num_chunks, hi = num_chunks;
            if (cross >= 0) {
                lo = cross * kSuper;
                hi = min(lo + kSuper, num_chunks);
            }
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                const uint32_t before = cross_base + (uint32_t)base16[(int64_t)mid * row_tiles + t];
                if (before >= limit) hi = mid; else lo = mid + 1;
            }
            const int ty = t / tiles_x, tx = t - ty * tiles_x;
            atomicMax(&close_chunk[(ty / kCloseBlk) * blocks_x + tx / kCloseBlk], lo);
        }
    }
    // ---- the last CTA to get here owns the two single-CTA steps ----------------------------------------------
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(done_counter, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // exclusive scan over tiles, 256 x 8 tiles per round: every thread loads its 8 consecutive totals up front (independent
    // loads), scans them locally, the 256 thread sums are scanned by warp shuffles, a carry links the rounds
    constexpr int kPer = 8;
    __shared__ uint32_t s_warp[kTablesThreads / 32];
    __shared__ uint32_t s_carry;
    const int lane = tid & 31, wid = tid >> 5;
    if (tid == 0) s_carry = 0u;
    if (tile_order != nullptr && tid < kOrderBuckets2) s_cnt[tid] = 0;
    __syncthreads();
    auto bucket = [&](uint32_t len) { return kOrderBuckets2 - 1 - min(kOrderBuckets2 - 1, (int)(len >> 3)); };
    for (int base = 0; base < num_tiles; base += kTablesThreads * kPer) {
        const int u0 = base + tid * kPer;
        uint32_t v[kPer], sum = 0u;
#pragma unroll
        for (int q = 0; q < kPer; ++q) {
            v[q] = (u0 + q < num_tiles) ? __ldcg(&tile_total[u0 + q]) : 0u;
            sum += v[q];
        }
        uint32_t inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t x = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += x;
        }
        if (lane == 31) s_warp[wid] = inc;
        __syncthreads();
        uint32_t wbase = 0u, round_total = 0u;
#pragma unroll
        for (int w = 0; w < kTablesThreads / 32; ++w) {
            const uint32_t x = s_warp[w];
            wbase += (w < wid) ? x : 0u;
            round_total += x;
        }
        uint32_t begin = s_carry + wbase + inc - sum;
#pragma unroll
        for (int q = 0; q < kPer; ++q) {
            const int u = u0 + q;
            if (u < num_tiles) {
                tile_start[u] = begin;
                reinterpret_cast<int2*>(ranges)[u] = v[q] ? make_int2((int)begin, (int)(begin + v[q])) : make_int2(0, 0);   // empty: (0,0)
                if (tile_order != nullptr) atomicAdd(&s_cnt[bucket(v[q])], 1);
            }
            begin += v[q];
        }
        __syncthreads();
        if (tid == 0) s_carry += round_total;
        __syncthreads();
    }
    if (tile_order == nullptr) return;
    if (tid < 32) {                                           // exclusive scan of the 256 bucket counts by one warp, 8 per lane
        int local[kOrderBuckets2 / 32], acc = 0;
#pragma unroll
        for (int q = 0; q < kOrderBuckets2 / 32; ++q) { local[q] = acc; acc += s_cnt[tid * (kOrderBuckets2 / 32) + q]; }
        int inc = acc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int x = __shfl_up_sync(0xffffffffu, inc, o);
            if (tid >= o) inc += x;
        }
#pragma unroll
        for (int q = 0; q < kOrderBuckets2 / 32; ++q) s_off[tid * (kOrderBuckets2 / 32) + q] = inc - acc + local[q];
    }
    __syncthreads();
    for (int base = 0; base < num_tiles; base += kTablesThreads * kPer) {
        const int u0 = base + tid * kPer;
        uint32_t v[kPer];
#pragma unroll
        for (int q = 0; q < kPer; ++q) v[q] = (u0 + q < num_tiles) ? __ldcg(&tile_total[u0 + q]) : 0u;
#pragma unroll
        for (int q = 0; q < kPer; ++q)
            if (u0 + q < num_tiles) tile_order[atomicAdd(&s_off[bucket(v[q])], 1)] = u0 + q;
    }
}

// 4. parallel scatter in rank order.  One warp per 32 consecutive ranks; for each rank the lanes serve
// its tiles side by side, so the table reads of a rectangle row and the local_pos read are coalesced.
// Measured floor: the 26 M four-byte stores land in 26 M different 32-byte sectors and the kernel runs
// at ~110 G store sectors/s whatever the load side does (batching the loads 4 ranks deep, 2x the
// instructions in flight, changed nothing -- profiles/r1_v5_binning.md).
__global__ void __launch_bounds__(256)
scatter_kernel(BinSizes sizes, const int32_t* __restrict__ sorted_ids, const int64_t* __restrict__ offsets,
               const ushort4* __restrict__ tile_rect, int tiles_x, int row_tiles,
               const uint8_t* __restrict__ local_pos, const uint16_t* __restrict__ base16,
               const uint32_t* __restrict__ super_base, const uint32_t* __restrict__ tile_start,
               const uint32_t* __restrict__ depth_keys, int32_t* __restrict__ entry_ids,
               uint64_t* __restrict__ entry_keys, ListCap cap) {
    const int lane = threadIdx.x & 31;
    const int64_t j0 = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32;
    grid_dependency_wait();
    int64_t num_sorted;
    if (!resolve_sizes(sizes, num_sorted)) return;
    if (j0 >= num_sorted) return;
    if (cap.flag_count != nullptr && *cap.flag_count == 0) return;        // completion pass with nothing to complete
    const int64_t j = j0 + lane;
    int id = 0, cnt = 0;
    long long off = 0;
    RectDesc d = {0, 1, 65537u};
    uint32_t dk = 0u;
    if (j < num_sorted) {
        id = sorted_ids[j];
        off = offsets[j];
        const ushort4 r = tile_rect[id];
        d = make_desc(r, tiles_x, cnt);
        if (entry_keys) dk = depth_keys[id];
        if (cap.close_chunk != nullptr && cap.flag_count == nullptr) {
            // every block this rectangle meets is already closed at this rank's chunk: nothing of it will be stored
            const int c = (int)(j / kChunk);
            bool open = false;
            for (int by = r.y / kCloseBlk; by <= (int)r.w / kCloseBlk; ++by)
                for (int bx = r.x / kCloseBlk; bx <= (int)r.z / kCloseBlk; ++bx) open |= c < cap.close_chunk[by * cap.blocks_x + bx];
            if (!open) cnt = 0;
        }
    }
    // only the ranks that still have something to store are visited
    unsigned todo = __ballot_sync(0xffffffffu, cnt > 0);
    while (todo) {
        const int l = __ffs(todo) - 1;
        todo &= todo - 1;
        const int s_id = __shfl_sync(0xffffffffu, id, l);
        const long long s_off = __shfl_sync(0xffffffffu, off, l);
        const int s_origin = __shfl_sync(0xffffffffu, d.origin, l);
        const int s_wc = __shfl_sync(0xffffffffu, d.w_cnt, l);
        const unsigned s_inv = __shfl_sync(0xffffffffu, d.inv, l);
        const int full = __shfl_sync(0xffffffffu, cnt, l);
        const uint64_t s_dk = (uint64_t)__shfl_sync(0xffffffffu, dk, l);
        const int s_w = s_wc & 0xfff;
        const int64_t chunk = (j0 + l) / kChunk;
        const uint16_t* b16 = base16 + chunk * row_tiles;
        const uint32_t* sb = super_base + (chunk / kSuper) * row_tiles;
        const uint8_t* lp = local_pos + s_off;
        const bool fast = full <= kMaxFastCount;
        for (int k = lane; k < full; k += 32) {
            int t;
            if (fast) {
                t = tile_of(k, s_origin, s_w, s_inv, tiles_x);
            } else {
                const int row = k / s_w;
                t = s_origin + row * tiles_x + (k - row * s_w);
            }
            // truncated lists: the first pass stores a tile's first `limit` entries only; the completion pass
            // stores the rest, for the tiles the compositing kernel flagged
            const uint32_t before = sb[t] + (uint32_t)b16[t];                         // pairs of tile t in earlier chunks
            if (cap.flag_count == nullptr ? before >= cap.limit : cap.tile_flags[t] == 0) continue;
            const uint32_t in_tile = before + (uint32_t)lp[k];                        // index inside tile t's list
            const bool store = cap.flag_count == nullptr ? in_tile < cap.limit : in_tile >= cap.limit;
            if (!store) continue;
            const uint32_t pos = tile_start[t] + in_tile;
            entry_ids[pos] = s_id;
            if (entry_keys) entry_keys[pos] = ((uint64_t)(uint32_t)t << 32) | s_dk;
        }
    }
}

struct CountLayout {
    int64_t counts, base16, super_tab, tile_total, tile_start, local_pos, close_chunk, total;
    int num_chunks, num_supers, row_tiles;
};
static CountLayout count_layout(int64_t num_sorted_cap, int64_t d, int32_t num_tiles) {
    CountLayout L;
    L.num_chunks = (int)((num_sorted_cap + kChunk - 1) / kChunk);
    if (L.num_chunks < 1) L.num_chunks = 1;
    L.num_supers = (L.num_chunks + kSuper - 1) / kSuper;
    L.row_tiles = (int)align_up(num_tiles, kRowAlign);
    int64_t o = 0;
    L.counts = o;     o += align_up((int64_t)L.num_chunks * L.row_tiles, 256);
    L.base16 = o;     o += align_up((int64_t)L.num_chunks * L.row_tiles * 2, 256);
    L.super_tab = o;  o += align_up((int64_t)L.num_supers * L.row_tiles * 4, 256);
    L.tile_total = o; o += align_up((int64_t)L.row_tiles * 4, 256);
    L.tile_start = o; o += align_up((int64_t)L.row_tiles * 4, 256);
    L.local_pos = o;  o += align_up(d, 256);
    L.close_chunk = o; o += align_up(((int64_t)num_tiles / kCloseBlk + 2) * 4, 256) + 256;   // ticket counter (first 256 B) + >= blocks of any grid shape
    L.total = o;
    return L;
}
constexpr int kMaxCountingTiles = 200000;      // 1 B of shared memory per tile for the running counters

static int tile_bits(int32_t num_tiles) {
    int bits = 1;
    while ((1ll << bits) < (long long)num_tiles) ++bits;
    return bits;
}

// ------------------------------------------------------------------------------------------------
// Blocked (two-level) stable counting sort: GS_BIN_BLOCKED.
//
// The flat counting sort above ends in 26 M four-byte stores that each land in a different 32-byte sector
// (~9.3 ps per pair, the floor of that design: profiles/r1_v5_binning.md).  Here the pairs are first grouped
// coarsely, so that the final stores of one CTA are runs of neighbouring list entries:
//   1. coarse_emit      every rank emits (block id, rank) for the <= 4 blocks of 8x8 tiles its rectangle meets
//   2. CUB radix sort   ONE pass on the block id (<= 8 bits for up to 255 blocks), stable => rank order kept
//   3. ranges / piece_map   per-block ranges of the coarse list, cut into pieces of kPiece ranks
//   4. fine_count       per piece: pairs per local tile (shared-memory atomics)         pcount[piece][64]
//   5. block_scan       per block: exclusive prefix over its pieces, tile totals        pbase[piece][64]
//   6. tile_scan        tile_start / tile_ranges
//   7. fine_write       per piece: cover bitsets per local tile (atomicOr), rank-ordered position of every pair
//                       by popcount, pairs grouped by tile in shared memory, then copied out in runs:
//                       entry_ids[tile_start + pbase + i] -- coalesced.
// Requires every rectangle to span at most 8 tiles per side (true for radius_max <= 50 px at 16-px tiles: 101 px);
// the caller (renderer.py) selects the flat sort otherwise.
// ------------------------------------------------------------------------------------------------
constexpr int kBlk = 8;
constexpr int kBlkTiles = kBlk * kBlk;
constexpr int kPiece = 128;               // ranks per fine CTA (4 warps)
constexpr int kPieceWords = kPiece / 32;

struct BlockGrid {
    int tiles_x, tiles_y, blocks_x, num_blocks;
};

__global__ void __launch_bounds__(256)
coarse_emit_kernel(BinSizes sizes, int64_t cap, const int32_t* __restrict__ sorted_ids, const ushort4* __restrict__ tile_rect,
                   BlockGrid g, uint32_t* __restrict__ keys, int32_t* __restrict__ vals) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= cap) return;
    int64_t num_sorted;
    const bool ok = resolve_sizes(sizes, num_sorted);
    uint32_t k[4] = {(uint32_t)g.num_blocks, (uint32_t)g.num_blocks, (uint32_t)g.num_blocks, (uint32_t)g.num_blocks};
    if (ok && j < num_sorted) {
        const ushort4 r = tile_rect[sorted_ids[j]];
        const int bx0 = r.x / kBlk, bx1 = min((int)r.z / kBlk, bx0 + 1);
        const int by0 = r.y / kBlk, by1 = min((int)r.w / kBlk, by0 + 1);
        int q = 0;
        for (int by = by0; by <= by1; ++by)
            for (int bx = bx0; bx <= bx1; ++bx) k[q++] = (uint32_t)(by * g.blocks_x + bx);
    }
    reinterpret_cast<uint4*>(keys)[j] = make_uint4(k[0], k[1], k[2], k[3]);
    reinterpret_cast<int4*>(vals)[j] = make_int4((int)j, (int)j, (int)j, (int)j);
}

// piece_begin[b] = number of pieces of the blocks before b; piece_begin[num_blocks] = all pieces.  One CTA.
__global__ void __launch_bounds__(1024)
piece_map_kernel(BinSizes sizes, int num_blocks, const int32_t* __restrict__ block_ranges, int32_t* __restrict__ piece_begin) {
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    int64_t unused;
    const bool ok = resolve_sizes(sizes, unused);
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < num_blocks; base += 1024) {
        const int b = base + tid;
        int v = 0;
        if (ok && b < num_blocks) v = (block_ranges[2 * b + 1] - block_ranges[2 * b] + kPiece - 1) / kPiece;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int x = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += x;
        }
        if (lane == 31) s_warp[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            const int w = s_warp[lane];
            int winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int x = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= o) winc += x;
            }
            s_warp[lane] = winc - w;
        }
        __syncthreads();
        const int begin = s_carry + s_warp[wid] + inc - v;
        if (b < num_blocks) piece_begin[b] = begin;
        __syncthreads();
        if (tid == 1023) s_carry = begin + v;
        __syncthreads();
    }
    if (tid == 0) piece_begin[num_blocks] = s_carry;
}

// Which piece is this CTA, which block does it belong to, which coarse entries does it cover.
struct PieceInfo {
    int block, bx, by, start, len;
};
__device__ __forceinline__ bool locate_piece(int p, const BlockGrid& g, const int32_t* __restrict__ piece_begin,
                                             const int32_t* __restrict__ block_ranges, PieceInfo& info) {
    if (p >= piece_begin[g.num_blocks]) return false;
    int lo = 0, hi = g.num_blocks;                        // last b with piece_begin[b] <= p
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (piece_begin[mid] <= p) lo = mid; else hi = mid;
    }
    info.block = lo;
    info.by = lo / g.blocks_x;
    info.bx = lo - info.by * g.blocks_x;
    const int kth = p - piece_begin[lo];
    info.start = block_ranges[2 * lo] + kth * kPiece;
    info.len = min(kPiece, block_ranges[2 * lo + 1] - info.start);
    return true;
}

// the part of a rank's rectangle inside block (bx, by), in local tile coordinates (empty: lx0 > lx1)
struct LocalRect {
    int lx0, lx1, ly0, ly1, id;
};
__device__ __forceinline__ LocalRect local_rect(const PieceInfo& info, int i, const int32_t* __restrict__ coarse_ranks,
                                                const int32_t* __restrict__ sorted_ids, const ushort4* __restrict__ tile_rect) {
    LocalRect L = {1, 0, 1, 0, 0};
    if (i < info.len) {
        L.id = sorted_ids[coarse_ranks[info.start + i]];
        const ushort4 r = tile_rect[L.id];
        L.lx0 = max((int)r.x - info.bx * kBlk, 0);
        L.lx1 = min((int)r.z - info.bx * kBlk, kBlk - 1);
        L.ly0 = max((int)r.y - info.by * kBlk, 0);
        L.ly1 = min((int)r.w - info.by * kBlk, kBlk - 1);
    }
    return L;
}

__global__ void __launch_bounds__(kPiece)
fine_count_kernel(BinSizes sizes, BlockGrid g, const int32_t* __restrict__ piece_begin, const int32_t* __restrict__ block_ranges,
                  const int32_t* __restrict__ coarse_ranks, const int32_t* __restrict__ sorted_ids,
                  const ushort4* __restrict__ tile_rect, uint16_t* __restrict__ pcount /* [pieces][64] */) {
    __shared__ int s_cnt[kBlkTiles];
    int64_t unused;
    if (!resolve_sizes(sizes, unused)) return;
    PieceInfo info;
    if (!locate_piece(blockIdx.x, g, piece_begin, block_ranges, info)) return;
    if (threadIdx.x < kBlkTiles) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const LocalRect L = local_rect(info, threadIdx.x, coarse_ranks, sorted_ids, tile_rect);
    for (int ly = L.ly0; ly <= L.ly1; ++ly)
        for (int lx = L.lx0; lx <= L.lx1; ++lx) atomicAdd(&s_cnt[ly * kBlk + lx], 1);
    __syncthreads();
    if (threadIdx.x < kBlkTiles) pcount[(int64_t)blockIdx.x * kBlkTiles + threadIdx.x] = (uint16_t)s_cnt[threadIdx.x];
}

// per block: exclusive prefix over its pieces for each of its 64 tiles; the tiles' totals
__global__ void __launch_bounds__(kBlkTiles)
block_scan_kernel(BinSizes sizes, BlockGrid g, const int32_t* __restrict__ piece_begin, const uint16_t* __restrict__ pcount,
                  uint32_t* __restrict__ pbase, uint32_t* __restrict__ tile_total) {
    int64_t unused;
    const bool ok = resolve_sizes(sizes, unused);
    const int b = blockIdx.x, lt = threadIdx.x;
    const int by = b / g.blocks_x, bx = b - by * g.blocks_x;
    const int tx = bx * kBlk + (lt & (kBlk - 1)), ty = by * kBlk + (lt / kBlk);
    uint32_t run = 0;
    if (ok) {
        const int p0 = piece_begin[b], p1 = piece_begin[b + 1];
        int p = p0;
        for (; p + 4 <= p1; p += 4) {
            uint32_t c[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) c[q] = pcount[(int64_t)(p + q) * kBlkTiles + lt];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                pbase[(int64_t)(p + q) * kBlkTiles + lt] = run;
                run += c[q];
            }
        }
        for (; p < p1; ++p) {
            const uint32_t c = pcount[(int64_t)p * kBlkTiles + lt];
            pbase[(int64_t)p * kBlkTiles + lt] = run;
            run += c;
        }
    }
    if (tx < g.tiles_x && ty < g.tiles_y) tile_total[ty * g.tiles_x + tx] = run;
}

constexpr int kPieceMaxPairs = kPiece * kBlkTiles;      // 8192: every rank covers the whole block
__global__ void __launch_bounds__(kPiece)
fine_write_kernel(BinSizes sizes, BlockGrid g, const int32_t* __restrict__ piece_begin, const int32_t* __restrict__ block_ranges,
                  const int32_t* __restrict__ coarse_ranks, const int32_t* __restrict__ sorted_ids,
                  const ushort4* __restrict__ tile_rect, const uint32_t* __restrict__ pbase,
                  const uint32_t* __restrict__ tile_start, const uint32_t* __restrict__ depth_keys,
                  int32_t* __restrict__ entry_ids, uint64_t* __restrict__ entry_keys) {
    __shared__ uint32_t s_cover[kBlkTiles][kPieceWords];      // which ranks of the piece cover local tile lt
    __shared__ uint16_t s_wprefix[kBlkTiles][kPieceWords];    // covering ranks in the earlier warps
    __shared__ uint32_t s_offset[kBlkTiles + 1];              // where tile lt's pairs start in s_out
    __shared__ uint32_t s_gbase[kBlkTiles];                   // where they start in entry_ids
    __shared__ int32_t s_out[kPieceMaxPairs];
    __shared__ uint8_t s_tile[kPieceMaxPairs];
    int64_t unused;
    if (!resolve_sizes(sizes, unused)) return;
    PieceInfo info;
    if (!locate_piece(blockIdx.x, g, piece_begin, block_ranges, info)) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int w = tid; w < kBlkTiles * kPieceWords; w += kPiece) (&s_cover[0][0])[w] = 0u;
    __syncthreads();
    const LocalRect L = local_rect(info, tid, coarse_ranks, sorted_ids, tile_rect);
    for (int ly = L.ly0; ly <= L.ly1; ++ly)
        for (int lx = L.lx0; lx <= L.lx1; ++lx) atomicOr(&s_cover[ly * kBlk + lx][warp], 1u << lane);
    __syncthreads();
    if (tid < kBlkTiles) {
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < kPieceWords; ++w) {
            s_wprefix[tid][w] = (uint16_t)run;
            run += __popc(s_cover[tid][w]);
        }
        // exclusive scan of the 64 totals over two warps
        uint32_t inc = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t x = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += x;
        }
        s_offset[tid + 1] = inc;                                   // inclusive within the warp; fixed up below
        const int tx = info.bx * kBlk + (tid & (kBlk - 1)), ty = info.by * kBlk + (tid / kBlk);
        s_gbase[tid] = (tx < g.tiles_x && ty < g.tiles_y)
                           ? tile_start[ty * g.tiles_x + tx] + pbase[(int64_t)blockIdx.x * kBlkTiles + tid] : 0u;
    }
    __syncthreads();
    if (tid >= 32 && tid < kBlkTiles) s_offset[tid + 1] += s_offset[32];   // second warp: add the first warp's total
    if (tid == 0) s_offset[0] = 0u;
    __syncthreads();
    for (int ly = L.ly0; ly <= L.ly1; ++ly)
        for (int lx = L.lx0; lx <= L.lx1; ++lx) {
            const int lt = ly * kBlk + lx;
            const uint32_t pos = s_offset[lt] + s_wprefix[lt][warp] + __popc(s_cover[lt][warp] & ((1u << lane) - 1u));
            s_out[pos] = L.id;
            s_tile[pos] = (uint8_t)lt;
        }
    __syncthreads();
    const uint32_t total = s_offset[kBlkTiles];
    for (uint32_t i = tid; i < total; i += kPiece) {
        const int lt = s_tile[i];
        const uint32_t pos = s_gbase[lt] + (i - s_offset[lt]);
        const int id = s_out[i];
        entry_ids[pos] = id;
        if (entry_keys) {
            const int tx = info.bx * kBlk + (lt & (kBlk - 1)), ty = info.by * kBlk + (lt / kBlk);
            entry_keys[pos] = ((uint64_t)(uint32_t)(ty * g.tiles_x + tx) << 32) | (uint64_t)depth_keys[id];
        }
    }
}

struct BlockedLayout {
    int64_t keys_in, keys_out, vals_in, vals_out, cub_temp, cub_bytes, block_ranges, piece_begin, pcount, pbase, tile_total,
        tile_start, total;
    int64_t items, max_pieces;
    int bits;
};
static BlockedLayout blocked_layout(int64_t cap, int32_t num_tiles, int num_blocks_bound) {
    BlockedLayout L;
    L.items = 4 * (cap > 0 ? cap : 1);
    L.bits = 1;
    while ((1ll << L.bits) < (long long)num_blocks_bound + 1) ++L.bits;
    L.max_pieces = (L.items + kPiece - 1) / kPiece + num_blocks_bound + 1;
    size_t sort_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                    (const int32_t*)nullptr, (int32_t*)nullptr, L.items, 0, L.bits);
    int64_t o = 0;
    L.keys_in = o;      o += align_up(L.items * 4, 256);
    L.keys_out = o;     o += align_up(L.items * 4, 256);
    L.vals_in = o;      o += align_up(L.items * 4, 256);
    L.vals_out = o;     o += align_up(L.items * 4, 256);
    L.cub_temp = o;     L.cub_bytes = align_up((int64_t)sort_bytes, 256); o += L.cub_bytes;
    L.block_ranges = o; o += align_up((int64_t)(num_blocks_bound + 1) * 8, 256);
    L.piece_begin = o;  o += align_up((int64_t)(num_blocks_bound + 2) * 4, 256);
    L.pcount = o;       o += align_up(L.max_pieces * kBlkTiles * 2, 256);
    L.pbase = o;        o += align_up(L.max_pieces * kBlkTiles * 4, 256);
    L.tile_total = o;   o += align_up((int64_t)num_tiles * 4, 256);
    L.tile_start = o;   o += align_up((int64_t)num_tiles * 4, 256);
    L.total = o;
    return L;
}
// any tile grid of num_tiles tiles has at most this many 8x8 blocks (a 1 x num_tiles strip)
static int blocks_bound(int32_t num_tiles) { return (num_tiles + kBlk - 1) / kBlk + 1; }

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
    const int64_t a = depth_sort_workspace_bytes(n > 0 ? n : 1);
    const int64_t b = sort_layout(d_capacity > 0 ? d_capacity : 1, num_tiles).total;
    const int64_t c = num_tiles <= kMaxCountingTiles ? count_layout(n > 0 ? n : 1, d_capacity > 0 ? d_capacity : 1, num_tiles).total : 0;
    const int64_t e = blocked_layout(n > 0 ? n : 1, num_tiles, blocks_bound(num_tiles)).total;
    int64_t m = a > b ? a : b;
    if (c > m) m = c;
    if (e > m) m = e;
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
    return depth_sort_launch(n, depth_keys, tiles_touched, workspace, workspace_bytes, sorted_ids, offsets, counters, st);
}

extern "C" int gs_bin_sort(int64_t n, int64_t num_sorted, int64_t d, const int32_t* sorted_ids, const int64_t* offsets,
                           const uint16_t* tile_rect, const uint32_t* depth_keys, int32_t tiles_x, int32_t num_tiles,
                           int32_t algo, void* workspace, int64_t workspace_bytes, int32_t* entry_ids,
                           int32_t* tile_ranges, uint64_t* entry_keys, const int64_t* counters_dev, int32_t list_cap,
                           int32_t* tile_order, int32_t* flag_count, void* stream) {
    GS_REQUIRE(n >= 0 && num_sorted >= 0 && num_sorted <= n && d >= 0, "bad sizes");
    GS_REQUIRE(num_tiles > 0 && tiles_x > 0, "bad tile grid");
    GS_REQUIRE(tile_ranges != nullptr, "tile_ranges is NULL");
    GS_REQUIRE(d < (1ll << 31), "tile pairs must fit int32 positions");
    cudaStream_t st = (cudaStream_t)stream;
    DeviceGuard guard(tile_ranges);
    const bool nothing = (counters_dev == nullptr && (d == 0 || num_sorted == 0)) ||
                         (counters_dev != nullptr && (d == 0 || n == 0));      // no pairs / no capacity: nothing can be written
    const bool counting_path = algo == GS_BIN_COUNTING || (algo == GS_BIN_AUTO && num_tiles <= kMaxCountingTiles);
    // the flat counting sort writes every tile's range itself (zeros on a capacity overflow); every other way out starts from zeros
    if (nothing || !counting_path) GS_CUDA_TRY(cudaMemsetAsync(tile_ranges, 0, (size_t)num_tiles * 2 * sizeof(int32_t), st));
    // flag_count (truncated lists): zeroed by the flat counting sort's tile-table launch; every other way out zeroes it here
    if (flag_count != nullptr && (nothing || !counting_path)) GS_CUDA_TRY(cudaMemsetAsync(flag_count, 0, sizeof(int32_t), st));
    if (nothing) {
        if (tile_order != nullptr) {                    // all lists are empty: any permutation is the longest-first order
            identity_order_kernel<<<(num_tiles + 255) / 256, 256, 0, st>>>(num_tiles, tile_order);
            GS_CUDA_TRY(cudaGetLastError());
            count_launches(1);
        }
        return GS_OK;
    }
    GS_REQUIRE(sorted_ids && offsets && tile_rect && workspace && entry_ids, "NULL array argument");
    GS_REQUIRE(entry_keys == nullptr || depth_keys != nullptr, "entry_keys needs depth_keys");
    GS_REQUIRE(algo >= 0 && algo <= 3, "algo must be 0 (auto), 1 (counting), 2 (radix) or 3 (blocked)");
    GS_REQUIRE(list_cap <= 0 || algo == GS_BIN_COUNTING || (algo == GS_BIN_AUTO && num_tiles <= kMaxCountingTiles),
               "truncated lists (list_cap) need the flat counting sort");
    const bool counting = algo == GS_BIN_COUNTING || (algo == GS_BIN_AUTO && num_tiles <= kMaxCountingTiles);
    GS_REQUIRE(counters_dev == nullptr || counting || algo == GS_BIN_BLOCKED, "device-side sizes (counters_dev) need a counting sort");
    if (algo == GS_BIN_BLOCKED) {
        GS_REQUIRE(num_tiles % tiles_x == 0, "num_tiles must be tiles_x * tiles_y");
        BlockGrid g;
        g.tiles_x = tiles_x;
        g.tiles_y = num_tiles / tiles_x;
        g.blocks_x = (g.tiles_x + kBlk - 1) / kBlk;
        g.num_blocks = g.blocks_x * ((g.tiles_y + kBlk - 1) / kBlk);
        const BlockedLayout L = blocked_layout(num_sorted, num_tiles, blocks_bound(num_tiles));
        if (workspace_bytes < L.total) {
            set_error("gs_bin_sort: workspace %lld B < required %lld B", (long long)workspace_bytes, (long long)L.total);
            return GS_ERR_WORKSPACE_TOO_SMALL;
        }
        char* w = (char*)workspace;
        uint32_t* keys_in = (uint32_t*)(w + L.keys_in);
        uint32_t* keys_out = (uint32_t*)(w + L.keys_out);
        int32_t* vals_in = (int32_t*)(w + L.vals_in);
        int32_t* vals_out = (int32_t*)(w + L.vals_out);
        int32_t* block_ranges = (int32_t*)(w + L.block_ranges);
        int32_t* piece_begin = (int32_t*)(w + L.piece_begin);
        uint16_t* pcount = (uint16_t*)(w + L.pcount);
        uint32_t* pbase = (uint32_t*)(w + L.pbase);
        uint32_t* tile_total = (uint32_t*)(w + L.tile_total);
        uint32_t* tile_start = (uint32_t*)(w + L.tile_start);
        const BinSizes sizes = {num_sorted, d, counters_dev};
        const int64_t cap = num_sorted;
        int bits = 1;
        while ((1 << bits) < g.num_blocks + 1) ++bits;
        coarse_emit_kernel<<<(unsigned)((cap + 255) / 256), 256, 0, st>>>(sizes, cap, sorted_ids, (const ushort4*)tile_rect, g,
                                                                          keys_in, vals_in);
        GS_CUDA_TRY(cudaGetLastError());
        size_t cub_bytes = (size_t)L.cub_bytes;
        GS_CUDA_TRY(cub::DeviceRadixSort::SortPairs(w + L.cub_temp, cub_bytes, (const uint32_t*)keys_in, keys_out,
                                                    (const int32_t*)vals_in, vals_out, 4 * cap, 0, bits, st));
        GS_CUDA_TRY(cudaMemsetAsync(block_ranges, 0, (size_t)(g.num_blocks + 1) * 8, st));
        ranges_kernel<<<(unsigned)((4 * cap + 255) / 256), 256, 0, st>>>(4 * cap, keys_out, block_ranges);
        GS_CUDA_TRY(cudaGetLastError());
        piece_map_kernel<<<1, 1024, 0, st>>>(sizes, g.num_blocks, block_ranges, piece_begin);
        GS_CUDA_TRY(cudaGetLastError());
        const unsigned max_pieces = (unsigned)((4 * cap + kPiece - 1) / kPiece + g.num_blocks);
        fine_count_kernel<<<max_pieces, kPiece, 0, st>>>(sizes, g, piece_begin, block_ranges, vals_out, sorted_ids,
                                                         (const ushort4*)tile_rect, pcount);
        GS_CUDA_TRY(cudaGetLastError());
        block_scan_kernel<<<g.num_blocks, kBlkTiles, 0, st>>>(sizes, g, piece_begin, pcount, pbase, tile_total);
        GS_CUDA_TRY(cudaGetLastError());
        tile_scan_kernel<<<1, 1024, 0, st>>>(sizes, num_tiles, tile_total, tile_start, tile_ranges);
        GS_CUDA_TRY(cudaGetLastError());
        fine_write_kernel<<<max_pieces, kPiece, 0, st>>>(sizes, g, piece_begin, block_ranges, vals_out, sorted_ids,
                                                         (const ushort4*)tile_rect, pbase, tile_start, depth_keys, entry_ids,
                                                         entry_keys);
        GS_CUDA_TRY(cudaGetLastError());
        count_launches(7);
        return GS_OK;
    }
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
        uint8_t* counts = (uint8_t*)(wsc + C.counts);
        uint16_t* base16 = (uint16_t*)(wsc + C.base16);
        uint32_t* super_tab = (uint32_t*)(wsc + C.super_tab);
        uint32_t* tile_total = (uint32_t*)(wsc + C.tile_total);
        uint32_t* tile_start = (uint32_t*)(wsc + C.tile_start);
        uint8_t* local_pos = (uint8_t*)(wsc + C.local_pos);
        const size_t smem_walk = (size_t)C.row_tiles;
        const BinSizes sizes = {num_sorted, d, counters_dev};
        {
            // function attributes once per device and size class, not per frame
            static int walk_smem_set[64] = {0};
            int dev_i = 0;
            cudaGetDevice(&dev_i);
            const int need = (int)smem_walk;
            if (dev_i >= 0 && dev_i < 64 && walk_smem_set[dev_i] < need + 1) {
                if (smem_walk > 48 * 1024)
                    GS_CUDA_TRY(cudaFuncSetAttribute(chunk_walk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, need));
                // one warp + row_tiles bytes per CTA: residency is bounded by shared memory, so ask for the largest carve-out
                GS_CUDA_TRY(cudaFuncSetAttribute(chunk_walk_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
                walk_smem_set[dev_i] = need + 1;
            }
        }
        ListCap cap = {list_cap > 0 ? (uint32_t)list_cap : 0xFFFFFFFFu, nullptr, nullptr, nullptr, 0};
        // control block of the tile-table launch: ticket (256 B) + one closing chunk per 8x8-tile block
        unsigned int* ticket = (unsigned int*)(wsc + C.close_chunk);
        int32_t* close_chunk = (int32_t*)(wsc + C.close_chunk + 256);
        int blocks_x = 0;
        size_t zero_bytes = 256;
        const bool closing = list_cap > 0 && num_tiles % tiles_x == 0;
        if (closing) {
            const int tiles_y = num_tiles / tiles_x;
            blocks_x = (tiles_x + kCloseBlk - 1) / kCloseBlk;
            zero_bytes += (size_t)blocks_x * ((tiles_y + kCloseBlk - 1) / kCloseBlk) * sizeof(int32_t);
        }
        chunk_walk_kernel<<<C.num_chunks, 32, smem_walk, st>>>(sizes, sorted_ids, offsets, (const ushort4*)tile_rect, tiles_x,
                                                              C.row_tiles, local_pos, counts, ticket, (int)(zero_bytes / 4));
        GS_CUDA_TRY(cudaGetLastError());
        // the three launches below follow their predecessor directly (no memset in between): dispatched early (PDL)
        GS_CUDA_TRY(launch_pdl(column_prefix_kernel, dim3(C.row_tiles / kRowAlign, C.num_supers), dim3(32, 32), 0, st,
                               sizes, C.row_tiles, (const uint8_t*)counts, base16, super_tab));
        // super-chunk prefix, closing chunks (truncated lists), tile scan and the forward's tile order: one launch
        GS_CUDA_TRY(launch_pdl(tile_tables_kernel, dim3((C.row_tiles + kTablesThreads - 1) / kTablesThreads), dim3(kTablesThreads), 0, st,
                               sizes, num_tiles, tiles_x, C.row_tiles, (const uint16_t*)base16, super_tab, tile_total, tile_start,
                               tile_ranges, cap.limit, blocks_x, closing ? close_chunk : (int32_t*)nullptr, tile_order, ticket,
                               flag_count));
        if (closing) {
            cap.close_chunk = close_chunk;
            cap.blocks_x = blocks_x;
        }
        const int64_t warps = (num_sorted + 31) / 32;
        GS_CUDA_TRY(launch_pdl(scatter_kernel, dim3((unsigned)((warps * 32 + 255) / 256)), dim3(256), 0, st,
                               sizes, sorted_ids, offsets, (const ushort4*)tile_rect, tiles_x, C.row_tiles, (const uint8_t*)local_pos,
                               (const uint16_t*)base16, (const uint32_t*)super_tab, (const uint32_t*)tile_start, depth_keys, entry_ids,
                               entry_keys, cap));
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

extern "C" int gs_bin_complete(int64_t n, int64_t num_sorted, int64_t d, const int32_t* sorted_ids, const int64_t* offsets,
                               const uint16_t* tile_rect, int32_t tiles_x, int32_t num_tiles, const void* workspace,
                               int64_t workspace_bytes, int32_t list_cap, const uint8_t* tile_flags,
                               const int32_t* flag_count, int32_t* entry_ids, const int64_t* counters_dev, void* stream) {
    GS_REQUIRE(n >= 0 && num_sorted >= 0 && num_sorted <= n && d >= 0, "bad sizes");
    GS_REQUIRE(num_tiles > 0 && tiles_x > 0 && num_tiles <= kMaxCountingTiles, "bad tile grid");
    GS_REQUIRE(list_cap > 0 && tile_flags && flag_count, "gs_bin_complete needs the list_cap of the first pass, tile_flags and flag_count");
    if (d == 0 || num_sorted == 0) return GS_OK;
    GS_REQUIRE(sorted_ids && offsets && tile_rect && workspace && entry_ids, "NULL array argument");
    DeviceGuard guard(entry_ids);
    const CountLayout C = count_layout(num_sorted, d, num_tiles);
    if (workspace_bytes < C.total) {
        set_error("gs_bin_complete: workspace %lld B < required %lld B", (long long)workspace_bytes, (long long)C.total);
        return GS_ERR_WORKSPACE_TOO_SMALL;
    }
    const char* wsc = (const char*)workspace;
    const BinSizes sizes = {num_sorted, d, counters_dev};
    const int64_t warps = (num_sorted + 31) / 32;
    GS_CUDA_TRY(launch_pdl(scatter_kernel, dim3((unsigned)((warps * 32 + 255) / 256)), dim3(256), 0, (cudaStream_t)stream,
                           sizes, sorted_ids, offsets, (const ushort4*)tile_rect, tiles_x, C.row_tiles,
                           (const uint8_t*)(wsc + C.local_pos), (const uint16_t*)(wsc + C.base16), (const uint32_t*)(wsc + C.super_tab),
                           (const uint32_t*)(wsc + C.tile_start), (const uint32_t*)nullptr, entry_ids, (uint64_t*)nullptr,
                           ListCap{(uint32_t)list_cap, tile_flags, flag_count, nullptr, 0}));
    count_launches(1);
    return GS_OK;
}
