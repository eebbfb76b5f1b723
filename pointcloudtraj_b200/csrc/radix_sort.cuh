// radix_sort.cuh -- hand-written stable LSD radix sort of (key, uint32 value) pairs, 8 bits per pass.
//
// Replaces the N sequential kd_insert3 calls of the reference (Utils/kdtree/src/kdtree.c:244-251):
// ordering the cloud along a space-filling curve IS the index construction.  The same sort orders query batches.
//
// Default path ("onesweep": Adinets & Merrill 2022), ONE kernel per 8-bit pass plus one histogram kernel up front:
//   os_histogram : the digit histograms of ALL passes from one read of the keys
//   os_pass      : a CTA takes the next tile (ticket from an atomic counter, so every predecessor tile is already running),
//                  ranks its elements stably (match_any ballots per warp round, warp-level running counters, scan across
//                  warps), publishes the tile's per-digit counts and obtains its per-digit offset by DECOUPLED LOOK-BACK
//                  over the predecessors' published counts / inclusive prefixes (thread d serves digit d), stages the tile
//                  in shared memory in output order and writes it run by run (coalesced)
// Per pass the keys are read once and written once -- the three-kernel path below reads them twice and runs a scan kernel
// in between.  Status words are 2 flag bits + 30 count bits, so sorts of 2^30 or more elements take the three-kernel path:
//   rs_histogram : per-tile digit histograms (warp-private shared-memory counters)
//   rs_scan_rows : exclusive scan of each digit's row of tile counts (one CTA per digit)
//   rs_scatter   : the same stable in-tile ranking and staged, coalesced scatter
// Stability: a warp owns a contiguous slice of the tile and walks it in rounds of 32 consecutive
// elements, so (warp, round, lane) order is the input order.
#pragma once
#include "common.cuh"

#define RS_THREADS 256
#define RS_WARPS (RS_THREADS / 32)
#define RS_RADIX 256

template <typename KeyT>
__device__ __forceinline__ uint32_t rs_digit(KeyT k, int shift)
{
    return (uint32_t)(k >> shift) & (RS_RADIX - 1);
}

// tile_hist layout: [digit][tile] so that a digit's row is contiguous for the scan
template <typename KeyT, int ITEMS>
__global__ void __launch_bounds__(RS_THREADS)
rs_histogram(const KeyT *__restrict__ keys, int64_t n, const unsigned long long *__restrict__ n_dev, int shift,
             uint32_t *__restrict__ tile_hist, int num_tiles)
{
    if (n_dev && (int64_t)*n_dev < n) n = (int64_t)*n_dev;     // element count known only on the device (<= host bound)
    __shared__ uint32_t hist[RS_WARPS][RS_RADIX];
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < RS_WARPS * RS_RADIX; i += RS_THREADS) (&hist[0][0])[i] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * (RS_THREADS * ITEMS);
#pragma unroll
    for (int r = 0; r < ITEMS; r++) {
        int64_t i = base + (int64_t)r * RS_THREADS + tid;
        if (i < n) atomicAdd(&hist[warp][rs_digit(keys[i], shift)], 1u);
    }
    __syncthreads();
    uint32_t s = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; w++) s += hist[w][tid];
    tile_hist[(int64_t)tid * num_tiles + blockIdx.x] = s;
}

// One CTA per digit: in-place exclusive scan of tile_hist[digit][0..num_tiles), total -> digit_total[digit]
__global__ void __launch_bounds__(RS_THREADS)
rs_scan_rows(uint32_t *__restrict__ tile_hist, int num_tiles, uint32_t *__restrict__ digit_total)
{
    __shared__ uint32_t warp_sum[RS_WARPS];
    __shared__ uint32_t carry_s;
    uint32_t *row = tile_hist + (int64_t)blockIdx.x * num_tiles;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < num_tiles; base += RS_THREADS) {
        int i = base + tid;
        uint32_t v = (i < num_tiles) ? row[i] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(PC_FULL_MASK, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_sum[warp] = incl;
        __syncthreads();
        uint32_t woff = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) if (w < warp) woff += warp_sum[w];
        uint32_t carry = carry_s;
        if (i < num_tiles) row[i] = carry + woff + incl - v;
        __syncthreads();
        if (tid == RS_THREADS - 1) carry_s = carry + woff + incl;
        __syncthreads();
    }
    if (tid == 0) digit_total[blockIdx.x] = carry_s;
}

template <typename KeyT, int ITEMS>
__global__ void __launch_bounds__(RS_THREADS)
rs_scatter(const KeyT *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
           KeyT *__restrict__ keys_out, uint32_t *__restrict__ vals_out, int64_t n, const unsigned long long *__restrict__ n_dev,
           int shift, const uint32_t *__restrict__ tile_prefix, int num_tiles, const uint32_t *__restrict__ digit_total)
{
    if (n_dev && (int64_t)*n_dev < n) n = (int64_t)*n_dev;
    if ((int64_t)blockIdx.x * (RS_THREADS * ITEMS) >= n) return;   // tile past the device-side count
    // counter[w][d]: first the running count of digit d seen by warp w, then its output base
    __shared__ uint32_t counter[RS_WARPS][RS_RADIX + 1];
    __shared__ uint32_t digit_base[RS_RADIX], local_base[RS_RADIX], out_base[RS_RADIX];
    __shared__ uint32_t warp_sum[RS_WARPS];
    __shared__ KeyT s_key[RS_THREADS * ITEMS];
    __shared__ uint32_t s_val[RS_THREADS * ITEMS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < RS_WARPS * (RS_RADIX + 1); i += RS_THREADS) (&counter[0][0])[i] = 0;

    // exclusive scan of the 256 digit totals (every CTA repeats it: 256 values, cheaper than a launch)
    {
        uint32_t v = digit_total[tid], incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(PC_FULL_MASK, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_sum[warp] = incl;
        __syncthreads();
        uint32_t woff = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) if (w < warp) woff += warp_sum[w];
        digit_base[tid] = woff + incl - v;
    }
    __syncthreads();

    const int64_t warp_base = (int64_t)blockIdx.x * (RS_THREADS * ITEMS) + (int64_t)warp * (32 * ITEMS);
    KeyT key[ITEMS];
    uint32_t val[ITEMS];
    uint32_t rank[ITEMS];   // rank of the element among equal digits inside this warp's slice
    uint32_t dig[ITEMS];
    const uint32_t lt = pc_lanemask_lt();
#pragma unroll
    for (int r = 0; r < ITEMS; r++) {
        int64_t i = warp_base + r * 32 + lane;
        bool ok = i < n;
        key[r] = ok ? keys_in[i] : (KeyT)0;
        val[r] = ok ? vals_in[i] : 0u;
        dig[r] = ok ? rs_digit(key[r], shift) : (uint32_t)RS_RADIX;   // bin 256 collects the out-of-range lanes
    }
#pragma unroll
    for (int r = 0; r < ITEMS; r++) {
        uint32_t peers = __match_any_sync(PC_FULL_MASK, dig[r]);
        int leader = __ffs(peers) - 1;
        uint32_t before = 0;
        if (lane == leader) {
            before = counter[warp][dig[r]];
            counter[warp][dig[r]] = before + __popc(peers);
        }
        before = __shfl_sync(PC_FULL_MASK, before, leader);
        rank[r] = before + __popc(peers & lt);
        __syncwarp();
    }
    __syncthreads();
    // thread d: the per-warp counts of digit d become offsets inside the digit's run of this tile; the tile's digit counts
    // are scanned into the run's position in the tile (local_base) and paired with its position in the output (out_base)
    {
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) {
            uint32_t c = counter[w][tid];
            counter[w][tid] = run;
            run += c;
        }
        uint32_t incl = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(PC_FULL_MASK, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_sum[warp] = incl;      // (the digit_total scan above is done with warp_sum: synced since)
        __syncthreads();
        uint32_t woff = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) if (w < warp) woff += warp_sum[w];
        local_base[tid] = woff + incl - run;
        out_base[tid] = digit_base[tid] + tile_prefix[(int64_t)tid * num_tiles + blockIdx.x];
    }
    __syncthreads();
    // stage the tile in shared memory in output order (digit runs, stable inside a run) ...
#pragma unroll
    for (int r = 0; r < ITEMS; r++) {
        if (dig[r] < RS_RADIX) {
            const uint32_t lpos = local_base[dig[r]] + counter[warp][dig[r]] + rank[r];
            s_key[lpos] = key[r];
            s_val[lpos] = val[r];
        }
    }
    __syncthreads();
    // ... and write it out run by run: consecutive threads write consecutive addresses inside a run (a run of 16 pairs is two
    // full sectors per array instead of 32 single-element sector writes spread over ITEMS store instructions)
    const int64_t tile_base = (int64_t)blockIdx.x * (RS_THREADS * ITEMS);
    const int tile_n = (int)(n - tile_base < (int64_t)(RS_THREADS * ITEMS) ? n - tile_base : (int64_t)(RS_THREADS * ITEMS));
    for (int j = tid; j < tile_n; j += RS_THREADS) {
        const KeyT k = s_key[j];
        const uint32_t d = rs_digit(k, shift);
        const uint32_t pos = out_base[d] + ((uint32_t)j - local_base[d]);
        keys_out[pos] = k;
        vals_out[pos] = s_val[j];
    }
}

// Host-side driver.  Sorts n pairs on `stream`; the result ends in (keys_a, vals_a) when the number of
// passes is even, otherwise in (keys_b, vals_b): the return value says which (0 = a, 1 = b).
// tile_hist must hold RS_RADIX * num_tiles(n) uint32, digit_total RS_RADIX uint32.
template <int ITEMS>
static inline int rs_num_tiles(int64_t n) { return (int)((n + (int64_t)RS_THREADS * ITEMS - 1) / ((int64_t)RS_THREADS * ITEMS)); }

template <typename KeyT, int ITEMS>
static int rs_sort_pairs(KeyT *keys_a, uint32_t *vals_a, KeyT *keys_b, uint32_t *vals_b, int64_t n,
                         int begin_bit, int end_bit, uint32_t *tile_hist, uint32_t *digit_total,
                         cudaStream_t stream, int64_t *launches, const unsigned long long *n_dev = nullptr)
{
    if (n <= 0) return 0;
    const int tiles = rs_num_tiles<ITEMS>(n);
    int which = 0;
    for (int shift = begin_bit; shift < end_bit; shift += 8) {
        KeyT *kin = which ? keys_b : keys_a, *kout = which ? keys_a : keys_b;
        uint32_t *vin = which ? vals_b : vals_a, *vout = which ? vals_a : vals_b;
        rs_histogram<KeyT, ITEMS><<<tiles, RS_THREADS, 0, stream>>>(kin, n, n_dev, shift, tile_hist, tiles);
        rs_scan_rows<<<RS_RADIX, RS_THREADS, 0, stream>>>(tile_hist, tiles, digit_total);
        rs_scatter<KeyT, ITEMS><<<tiles, RS_THREADS, 0, stream>>>(kin, vin, kout, vout, n, n_dev, shift, tile_hist, tiles, digit_total);
        if (launches) *launches += 3;
        which ^= 1;
    }
    return which;
}

// ---- onesweep -----------------------------------------------------------------------------------------------------------------
#define OS_MAX_PASSES 8
#ifndef OS_MIN_CTAS
#define OS_MIN_CTAS 4                    // resident CTAs per SM the pass kernel is compiled for (<= 64 registers per thread)
#endif
#define OS_FLAG_LOCAL 0x40000000u        // the word holds the tile's own count of this digit
#define OS_FLAG_INCL  0x80000000u        // the word holds the inclusive prefix over tiles 0 .. this one
#define OS_VALUE_MASK 0x3fffffffu
#define OS_MAX_N ((int64_t)1 << 30)

// scratch layout (uint32 words): [0, 16) tile tickets, one per pass; [16, 16 + 256 P) digit histograms of the P passes;
// then P status arrays of tiles x 256 words.  One memset clears all of it before os_histogram.
static inline int64_t os_scratch_words(int64_t tiles, int passes) { return 16 + (int64_t)RS_RADIX * passes + (int64_t)RS_RADIX * tiles * passes; }

static inline uint32_t *os_ghist(uint32_t *scratch) { return scratch + 16; }
static inline void os_clear(uint32_t *scratch, int64_t n, int items, int passes, cudaStream_t stream)
{
    const int64_t tiles = (n + (int64_t)RS_THREADS * items - 1) / ((int64_t)RS_THREADS * items);
    cudaMemsetAsync(scratch, 0, (size_t)os_scratch_words(tiles, passes) * sizeof(uint32_t), stream);
}

template <typename KeyT>
__global__ void __launch_bounds__(RS_THREADS)
os_histogram(const KeyT *__restrict__ keys, int64_t n, const unsigned long long *__restrict__ n_dev, int begin_bit, int passes,
             uint32_t *__restrict__ ghist)
{
    if (n_dev && (int64_t)*n_dev < n) n = (int64_t)*n_dev;
    __shared__ uint32_t hist[OS_MAX_PASSES][RS_RADIX];
    const int tid = threadIdx.x;
    for (int i = tid; i < passes * RS_RADIX; i += RS_THREADS) (&hist[0][0])[i] = 0;
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * RS_THREADS + tid; i < n; i += (int64_t)gridDim.x * RS_THREADS) {
        const KeyT k = keys[i];
        for (int p = 0; p < passes; p++) atomicAdd(&hist[p][rs_digit(k, begin_bit + 8 * p)], 1u);
    }
    __syncthreads();
    for (int i = tid; i < passes * RS_RADIX; i += RS_THREADS) {
        const uint32_t c = (&hist[0][0])[i];
        if (c) atomicAdd(&ghist[i], c);
    }
}

template <typename KeyT, int ITEMS>
__global__ void __launch_bounds__(RS_THREADS, OS_MIN_CTAS)
os_pass(const KeyT *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
        KeyT *__restrict__ keys_out, uint32_t *__restrict__ vals_out, int64_t n, const unsigned long long *__restrict__ n_dev,
        int shift, const uint32_t *__restrict__ ghist, volatile uint32_t *status, uint32_t *ticket)
{
    if (n_dev && (int64_t)*n_dev < n) n = (int64_t)*n_dev;
    __shared__ uint32_t counter[RS_WARPS][RS_RADIX + 1];
    __shared__ uint32_t digit_base[RS_RADIX], local_base[RS_RADIX], out_base[RS_RADIX];
    __shared__ uint32_t warp_sum[RS_WARPS];
    __shared__ uint32_t s_tile;
    __shared__ KeyT s_key[RS_THREADS * ITEMS];
    __shared__ uint32_t s_val[RS_THREADS * ITEMS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);          // tiles are taken in order: all predecessors of a tile are running
    for (int i = tid; i < RS_WARPS * (RS_RADIX + 1); i += RS_THREADS) (&counter[0][0])[i] = 0;
    __syncthreads();
    const int64_t tile = s_tile;
    const int64_t tile_base = tile * (RS_THREADS * ITEMS);
    if (tile_base >= n) return;                             // tile past the device-side count

    // exclusive scan of the 256 digit totals of this pass
    {
        uint32_t v = ghist[tid], incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(PC_FULL_MASK, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_sum[warp] = incl;
        __syncthreads();
        uint32_t woff = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) if (w < warp) woff += warp_sum[w];
        digit_base[tid] = woff + incl - v;
    }

    const int64_t warp_base = tile_base + (int64_t)warp * (32 * ITEMS);
    KeyT key[ITEMS];
    uint32_t val[ITEMS];
    uint32_t rank[ITEMS];   // rank of the element among equal digits inside this warp's slice
    uint32_t dig[ITEMS];
    const uint32_t lt = pc_lanemask_lt();
#pragma unroll
    for (int r = 0; r < ITEMS; r++) {
        int64_t i = warp_base + r * 32 + lane;
        bool ok = i < n;
        key[r] = ok ? keys_in[i] : (KeyT)0;
        val[r] = ok ? vals_in[i] : 0u;
        dig[r] = ok ? rs_digit(key[r], shift) : (uint32_t)RS_RADIX;   // bin 256 collects the out-of-range lanes
    }
#pragma unroll
    for (int r = 0; r < ITEMS; r++) {
        uint32_t peers = __match_any_sync(PC_FULL_MASK, dig[r]);
        int leader = __ffs(peers) - 1;
        uint32_t before = 0;
        if (lane == leader) {
            before = counter[warp][dig[r]];
            counter[warp][dig[r]] = before + __popc(peers);
        }
        before = __shfl_sync(PC_FULL_MASK, before, leader);
        rank[r] = before + __popc(peers & lt);
        __syncwarp();
    }
    __syncthreads();
    // thread d: the per-warp counts of digit d become offsets inside the digit's run of this tile; the tile's count is
    // published, and the tile's offset inside the digit's output range comes from the predecessors (decoupled look-back)
    {
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) {
            uint32_t c = counter[w][tid];
            counter[w][tid] = run;
            run += c;
        }
        volatile uint32_t *mine = status + tile * RS_RADIX + tid;
        *mine = run | (tile == 0 ? OS_FLAG_INCL : OS_FLAG_LOCAL);
        uint32_t incl = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(PC_FULL_MASK, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_sum[warp] = incl;      // (the digit_total scan above is done with warp_sum: synced since)
        __syncthreads();
        uint32_t woff = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) if (w < warp) woff += warp_sum[w];
        local_base[tid] = woff + incl - run;
        // decoupled look-back: thread d walks the predecessors' words of digit d backwards, adding counts until it meets an
        // inclusive prefix.  (A variant that read 8 predecessors per round trip -- one per warp, parked in shared memory -- was
        // slower, 0.234 vs 0.206 ms for three passes over 5.6 M pairs: profiles/r2_variants_ab.txt.)
        uint32_t excl = 0;
        if (tile > 0) {
            int64_t look = tile - 1;
            for (;;) {
                const uint32_t v = status[look * RS_RADIX + tid];
                if ((v & (OS_FLAG_LOCAL | OS_FLAG_INCL)) == 0) continue;      // predecessor is running but has not published yet
                excl += v & OS_VALUE_MASK;
                if (v & OS_FLAG_INCL) break;
                look--;                                                     // tile 0 always publishes an inclusive prefix
            }
            *mine = (excl + run) | OS_FLAG_INCL;
        }
        out_base[tid] = digit_base[tid] + excl;
    }
    __syncthreads();
    // stage the tile in shared memory in output order (digit runs, stable inside a run) ...
#pragma unroll
    for (int r = 0; r < ITEMS; r++) {
        if (dig[r] < RS_RADIX) {
            const uint32_t lpos = local_base[dig[r]] + counter[warp][dig[r]] + rank[r];
            s_key[lpos] = key[r];
            s_val[lpos] = val[r];
        }
    }
    __syncthreads();
    // ... and write it out run by run: consecutive threads write consecutive addresses inside a run
    const int tile_n = (int)(n - tile_base < (int64_t)(RS_THREADS * ITEMS) ? n - tile_base : (int64_t)(RS_THREADS * ITEMS));
    for (int j = tid; j < tile_n; j += RS_THREADS) {
        const KeyT k = s_key[j];
        const uint32_t d = rs_digit(k, shift);
        const uint32_t pos = out_base[d] + ((uint32_t)j - local_base[d]);
        keys_out[pos] = k;
        vals_out[pos] = s_val[j];
    }
}

// Host-side driver of the onesweep path; same contract as rs_sort_pairs.  `scratch` holds os_scratch_words(tiles, passes)
// uint32 words.  hist_done: the caller's key-producing kernel already accumulated the digit histograms into scratch + 16
// (and cleared the scratch before it).
template <typename KeyT, int ITEMS>
static int os_sort_pairs(KeyT *keys_a, uint32_t *vals_a, KeyT *keys_b, uint32_t *vals_b, int64_t n,
                         int begin_bit, int end_bit, uint32_t *scratch, int sm_count,
                         cudaStream_t stream, int64_t *launches, const unsigned long long *n_dev = nullptr, bool hist_done = false)
{
    if (n <= 0) return 0;
    const int passes = (end_bit - begin_bit + 7) / 8;
    const int tiles = rs_num_tiles<ITEMS>(n);
    uint32_t *ticket = scratch, *ghist = os_ghist(scratch), *status = scratch + 16 + RS_RADIX * passes;
    if (!hist_done) {
        os_clear(scratch, n, ITEMS, passes, stream);
        const int64_t want = (n + RS_THREADS * 8 - 1) / (RS_THREADS * 8);
        const int grid = (int)(want < (int64_t)sm_count * 8 ? want : (int64_t)sm_count * 8);
        os_histogram<KeyT><<<grid, RS_THREADS, 0, stream>>>(keys_a, n, n_dev, begin_bit, passes, ghist);
        if (launches) *launches += 1;
    }
    int which = 0;
    for (int p = 0; p < passes; p++) {
        KeyT *kin = which ? keys_b : keys_a, *kout = which ? keys_a : keys_b;
        uint32_t *vin = which ? vals_b : vals_a, *vout = which ? vals_a : vals_b;
        os_pass<KeyT, ITEMS><<<tiles, RS_THREADS, 0, stream>>>(kin, vin, kout, vout, n, n_dev, begin_bit + 8 * p, ghist + RS_RADIX * p,
                                                                status + (int64_t)RS_RADIX * tiles * p, ticket + p);
        if (launches) *launches += 1;
        which ^= 1;
    }
    return which;
}
