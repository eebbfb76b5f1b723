// sampler_kernels.cuh -- the RRT* sample stream generated on the device (no per-sample host traffic).
//
// The reference draws its samples from std::default_random_engine eng(0) (corridor_finder.cpp:12) = minstd_rand0,
//     x' = 16807 x mod (2^31 - 1),
// through std::uniform_real_distribution<double> (genSample, corridor_finder.cpp:333-358).  libstdc++ builds every uniform
// double from TWO draws (generate_canonical<double, 53>: bits/random.tcc), and a sample takes one uniform (the goal bias:
// the sample is the goal) or four (bias + x, y, z).  The stream is therefore a chain: where sample j starts depends on how
// many goal-biased samples came before.  The chain is resolved in parallel:
//
//   position u   = the u-th PAIR of draws; the engine state in front of it is seed * 16807^(2u) (skip-ahead, O(log u))
//   len(u)       = 1 if the uniform made of pair u is <= goal_ratio, else 4   ("a sample that starts at u ends at u + len(u)")
//   jump map     for a chunk of positions: entry offset e in {0,1,2,3} -> (samples started in the chunk, offset into the next
//                chunk); maps of neighbouring chunks compose associatively, so they are scanned like a prefix sum
//
//   pc_sample_mask_kernel   len(u) for every position (one bit each) + the jump map of every tile of 8192 positions
//   pc_sample_tile_kernel   scan of the tile maps: where the first sample of each tile starts and its ordinal
//   pc_sample_emit_kernel   every thread walks its 32 positions from its entry point and writes the samples it starts
//
// Every floating-point step is the IEEE double operation libstdc++ performs (no contraction), so sample j is bit-identical
// with the j-th genSample() of the reference.  The informed-ellipsoid branch (cbrt / acos / sin / cos, :361-378) is not
// offered here: libm's results are not reproducible bit for bit on the device -- the caller falls back to host samples once
// a path is found (pc_rrt.hpp).
#pragma once
#include "common.cuh"

#define PC_LCG_A 16807u
#define PC_LCG_M 2147483647u
#define PC_SMP_CHUNK 32                                    // positions per thread (one mask word)
#define PC_SMP_THREADS 256
#define PC_SMP_TILE (PC_SMP_CHUNK * PC_SMP_THREADS)        // positions per CTA
#define PC_SMP_SCAN_THREADS 1024

struct pc_sampler_dev {
    uint32_t state;                 // engine state in front of the first draw
    double goal_ratio, inlier_sum;  // bias <= goal_ratio: goal; bias <= goal_ratio + inlier_ratio: local box; else global box
    double end_pt[3];
    double lo[3], span[3];          // rand_x / rand_y / rand_z: a, b - a
    double in_lo[3], in_span[3];    // rand_x_in / rand_y_in / rand_z_in
};

__host__ __device__ __forceinline__ uint32_t pc_lcg_mulmod(uint32_t a, uint32_t b)
{
    const uint64_t p = (uint64_t)a * b;                        // < 2^62
    uint64_t x = (p & PC_LCG_M) + (p >> 31);                   // Mersenne modulus: 2^31 = 1 (mod M)
    x = (x & PC_LCG_M) + (x >> 31);
    return (uint32_t)(x >= PC_LCG_M ? x - PC_LCG_M : x);
}

// engine state after `draws` draws
__host__ __device__ inline uint32_t pc_lcg_skip(uint32_t state, uint64_t draws)
{
    uint32_t f = 1u, b = PC_LCG_A;
    for (; draws; draws >>= 1) {
        if (draws & 1) f = pc_lcg_mulmod(f, b);
        b = pc_lcg_mulmod(b, b);
    }
    return pc_lcg_mulmod(state, f);
}

// std::generate_canonical<double, 53>(minstd_rand0) from its two draws (bits/random.tcc:3349-3381): range r = 2147483646,
// sum = (x1 - 1) + (x2 - 1) * r, ret = sum / (double)(r * r), clamped below 1
__device__ __forceinline__ double pc_canonical(uint32_t x1, uint32_t x2)
{
    double s = (double)(x1 - 1u);
    s = __dadd_rn(s, __dmul_rn((double)(x2 - 1u), 2147483646.0));
    const double r = __ddiv_rn(s, 0x1.fffffffp+61);           // (double)(2147483646.0L * 2147483646.0L) = 4611686009837453312
    return r >= 1.0 ? 0x1.fffffffffffffp-1 : r;
}

// jump map of a run of positions: entry offset e -> samples started (cnt[e]) and the offset into the following run (2 bits each)
struct pc_jump { uint32_t cnt[4]; uint32_t exit; };
#define PC_JUMP_IDENTITY_EXIT 0xe4u                               // e -> e

__device__ __forceinline__ pc_jump pc_jump_identity()
{
    pc_jump j; j.cnt[0] = j.cnt[1] = j.cnt[2] = j.cnt[3] = 0u; j.exit = PC_JUMP_IDENTITY_EXIT; return j;
}
// first a, then b
__device__ __forceinline__ pc_jump pc_jump_compose(const pc_jump &a, const pc_jump &b)
{
    pc_jump o; o.exit = 0u;
#pragma unroll
    for (int e = 0; e < 4; e++) {
        const uint32_t x = (a.exit >> (2 * e)) & 3u;
        o.cnt[e] = a.cnt[e] + (x == 0 ? b.cnt[0] : x == 1 ? b.cnt[1] : x == 2 ? b.cnt[2] : b.cnt[3]);
        o.exit |= ((b.exit >> (2 * x)) & 3u) << (2 * e);
    }
    return o;
}
// one mask word: bit i set = a sample starting at position i is the goal (length 1)
__device__ __forceinline__ pc_jump pc_jump_of_mask(uint32_t mask)
{
    pc_jump o; o.exit = 0u;
#pragma unroll
    for (int e = 0; e < 4; e++) {
        int pos = e; uint32_t c = 0;
        while (pos < PC_SMP_CHUNK) { c++; pos += ((mask >> pos) & 1u) ? 1 : 4; }
        o.cnt[e] = c; o.exit |= (uint32_t)(pos - PC_SMP_CHUNK) << (2 * e);
    }
    return o;
}
__device__ __forceinline__ pc_jump pc_jump_shfl_up(const pc_jump &v, int d)
{
    pc_jump o;
#pragma unroll
    for (int e = 0; e < 4; e++) o.cnt[e] = __shfl_up_sync(PC_FULL_MASK, v.cnt[e], d);
    o.exit = __shfl_up_sync(PC_FULL_MASK, v.exit, d);
    return o;
}

// inclusive scan of the per-thread maps of a CTA (blockDim.x = 32 * warps <= 1024); returns the thread's inclusive map, *excl =
// the map of everything in front of the thread
__device__ __forceinline__ pc_jump pc_jump_block_scan(pc_jump v, pc_jump *excl, pc_jump *s_warp /* [32] */)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const pc_jump p = pc_jump_shfl_up(v, d);
        if (lane >= d) v = pc_jump_compose(p, v);
    }
    if (lane == 31) s_warp[w] = v;
    __syncthreads();
    if (w == 0) {
        pc_jump t = lane < nw ? s_warp[lane] : pc_jump_identity();
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const pc_jump p = pc_jump_shfl_up(t, d);
            if (lane >= d) t = pc_jump_compose(p, t);
        }
        if (lane < nw) s_warp[lane] = t;
    }
    __syncthreads();
    pc_jump prev = pc_jump_shfl_up(v, 1);
    if (lane == 0) prev = pc_jump_identity();
    pc_jump before = w > 0 ? pc_jump_compose(s_warp[w - 1], prev) : prev;
    *excl = before;
    return w > 0 ? pc_jump_compose(s_warp[w - 1], v) : v;
}

// len(u) for n_pos positions, 32 per thread; tile_map[blockIdx.x] = the jump map of the CTA's 8192 positions
__global__ void __launch_bounds__(PC_SMP_THREADS)
pc_sample_mask_kernel(uint32_t seed_state, double goal_ratio, uint32_t *__restrict__ masks, pc_jump *__restrict__ tile_map)
{
    __shared__ pc_jump s_warp[32];
    const uint64_t chunk = (uint64_t)blockIdx.x * PC_SMP_THREADS + threadIdx.x;
    uint32_t st = pc_lcg_skip(seed_state, 2ull * PC_SMP_CHUNK * chunk);
    uint32_t mask = 0u;
#pragma unroll 4
    for (int i = 0; i < PC_SMP_CHUNK; i++) {
        const uint32_t x1 = pc_lcg_mulmod(st, PC_LCG_A), x2 = pc_lcg_mulmod(x1, PC_LCG_A);
        st = x2;
        if (pc_canonical(x1, x2) <= goal_ratio) mask |= 1u << i;
    }
    masks[chunk] = mask;
    pc_jump excl;
    const pc_jump incl = pc_jump_block_scan(pc_jump_of_mask(mask), &excl, s_warp);
    if (threadIdx.x == blockDim.x - 1) tile_map[blockIdx.x] = incl;
}

// one CTA: tile_entry[t] = (offset of the first sample start inside tile t, ordinal of that sample), the stream starts with a
// sample at position 0
__global__ void __launch_bounds__(PC_SMP_SCAN_THREADS)
pc_sample_tile_kernel(const pc_jump *__restrict__ tile_map, int64_t n_tiles, uint2 *__restrict__ tile_entry)
{
    __shared__ pc_jump s_warp[32];
    const int64_t per = (n_tiles + blockDim.x - 1) / blockDim.x;
    const int64_t t0 = per * threadIdx.x, t1 = t0 + per < n_tiles ? t0 + per : n_tiles;
    pc_jump mine = pc_jump_identity();
    for (int64_t t = t0; t < t1; t++) mine = pc_jump_compose(mine, tile_map[t]);
    pc_jump excl;
    pc_jump_block_scan(mine, &excl, s_warp);
    uint32_t e = excl.exit & 3u, base = excl.cnt[0];
    for (int64_t t = t0; t < t1; t++) {
        tile_entry[t] = make_uint2(e, base);
        const pc_jump m = tile_map[t];
        base += e == 0 ? m.cnt[0] : e == 1 ? m.cnt[1] : e == 2 ? m.cnt[2] : m.cnt[3];
        e = (m.exit >> (2 * e)) & 3u;
    }
}

// the samples: thread = 32 positions; sample j < k goes to out_xyz[3j..] (double, the planner's Vector3d) and, cast to
// float32 like findNearstVertex does (corridor_finder.cpp:430), to out_q[j] = (x, y, z, 0).  The thread that meets the start of
// sample k stores the engine state in front of it: the state the host engine continues from.
// FUSED: a planner-sized batch (4k + 1 <= 32768 positions) is ONE CTA of 1024 threads that computes its masks itself -- one launch
// instead of three.
template <bool FUSED>
__global__ void __launch_bounds__(FUSED ? PC_SMP_SCAN_THREADS : PC_SMP_THREADS)
pc_sample_emit_kernel(pc_sampler_dev S, const uint32_t *__restrict__ masks, const uint2 *__restrict__ tile_entry, uint64_t k,
                      double *__restrict__ out_xyz, float4 *__restrict__ out_q, uint32_t *__restrict__ out_state)
{
    __shared__ pc_jump s_warp[32];
    const uint64_t chunk = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t mask = 0u;
    if (FUSED) {
        uint32_t st0 = pc_lcg_skip(S.state, 2ull * PC_SMP_CHUNK * chunk);
#pragma unroll 4
        for (int i = 0; i < PC_SMP_CHUNK; i++) {
            const uint32_t x1 = pc_lcg_mulmod(st0, PC_LCG_A), x2 = pc_lcg_mulmod(x1, PC_LCG_A);
            st0 = x2;
            if (pc_canonical(x1, x2) <= S.goal_ratio) mask |= 1u << i;
        }
    } else {
        mask = masks[chunk];
    }
    pc_jump excl;
    pc_jump_block_scan(pc_jump_of_mask(mask), &excl, s_warp);
    const uint2 te = FUSED ? make_uint2(0u, 0u) : tile_entry[blockIdx.x];
    const uint32_t e0 = te.x;
    int pos = (int)((excl.exit >> (2 * e0)) & 3u);
    uint64_t j = (uint64_t)te.y + (e0 == 0 ? excl.cnt[0] : e0 == 1 ? excl.cnt[1] : e0 == 2 ? excl.cnt[2] : excl.cnt[3]);
    if (j > k || pos >= PC_SMP_CHUNK) return;
    uint32_t st = pc_lcg_skip(S.state, 2ull * (PC_SMP_CHUNK * chunk + (uint64_t)pos));
    while (pos < PC_SMP_CHUNK) {
        if (j == k) { *out_state = st; return; }
        uint32_t x1 = pc_lcg_mulmod(st, PC_LCG_A), x2 = pc_lcg_mulmod(x1, PC_LCG_A);
        st = x2;
        const double bias = pc_canonical(x1, x2);
        double p[3];
        if (bias <= S.goal_ratio) {
            p[0] = S.end_pt[0]; p[1] = S.end_pt[1]; p[2] = S.end_pt[2];
            pos += 1;
        } else {
            const bool in = bias <= S.inlier_sum;                 // (bias > goal_ratio holds here)
#pragma unroll
            for (int a = 0; a < 3; a++) {
                x1 = pc_lcg_mulmod(st, PC_LCG_A); x2 = pc_lcg_mulmod(x1, PC_LCG_A);
                st = x2;
                // uniform_real_distribution: canonical * (b - a) + a
                p[a] = __dadd_rn(__dmul_rn(pc_canonical(x1, x2), in ? S.in_span[a] : S.span[a]), in ? S.in_lo[a] : S.lo[a]);
            }
            pos += 4;
        }
        if (out_xyz) { out_xyz[3 * j] = p[0]; out_xyz[3 * j + 1] = p[1]; out_xyz[3 * j + 2] = p[2]; }
        if (out_q) out_q[j] = make_float4((float)p[0], (float)p[1], (float)p[2], 0.f);
        j++;
    }
}

// ---- the speculative expansion batch (pc_expand_batch) -------------------------------------------------------------------
// node centres (double) -> float32 positions of the node tree (kd_insertf takes floats: corridor_finder.cpp:734)
__global__ void pc_node_pos_kernel(const double *__restrict__ coord, int64_t n, float4 *__restrict__ pos)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) pos[i] = make_float4((float)coord[3 * i], (float)coord[3 * i + 1], (float)coord[3 * i + 2], 0.f);
}

// findNearstVertex for a planner-sized batch against a planner-sized node set, without an index: one WARP per sample, the
// lanes stride over the nodes, exact fp64 distances in the reference's operation order (kd_nearestf on float positions:
// kdtree.c:379-382), ties to the lowest index -- the very definition the tree kernels implement, so the answers are identical.
// k x n pair evaluations at the fp64 rate: 4096 samples x 4096 nodes take a few microseconds, where building the node index
// (9 launches) and walking it took ~90 us of a 260 us call.
#define PC_BRUTE_MAX_PAIRS ((int64_t)1 << 26)
__global__ void __launch_bounds__(256)
pc_nearest_brute_kernel(const float4 *__restrict__ pos, int64_t n, const float4 *__restrict__ q, int64_t k, int32_t *__restrict__ out_idx)
{
    const int lane = threadIdx.x & 31;
    const int64_t j = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (j >= k) return;
    const float4 qq = q[j];
    const double qx = (double)qq.x, qy = (double)qq.y, qz = (double)qq.z;
    double best = INFINITY;
    int32_t bi = -1;
    for (int64_t i = lane; i < n; i += 32) {
        const float4 p = __ldg(pos + i);
        const double e = pc_exact_d2(p.x, p.y, p.z, qx, qy, qz);
        if (e < best) { best = e; bi = (int32_t)i; }          // a lane's indices ascend: the lowest index of equal distances stays
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(PC_FULL_MASK, best, o);
        const int32_t oi = __shfl_xor_sync(PC_FULL_MASK, bi, o);
        if (ob < best || (ob == best && (uint32_t)oi < (uint32_t)bi)) { best = ob; bi = oi; }
    }
    if (lane == 0) out_idx[j] = bi;
}

// genNewNode's steering (corridor_finder.cpp:385-402): the sample is pulled onto the surface of the nearest node's sphere.
// In place: xyz[j] sample -> centre, q[j] -> float32 centre for the cloud query.  ok[j] = the nearest vertex exists and is valid
// (:726).
__global__ void pc_steer_kernel(double *__restrict__ xyz, float4 *__restrict__ q, const int32_t *__restrict__ nearest, int64_t k,
                                const double *__restrict__ node_coord, const float *__restrict__ node_radius,
                                const uint8_t *__restrict__ node_valid, uint8_t *__restrict__ ok)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= k) return;
    const int32_t nn = nearest[j];
    if (nn < 0 || !node_valid[nn]) { ok[j] = 0; return; }
    const double s[3] = { xyz[3 * j], xyz[3 * j + 1], xyz[3 * j + 2] };
    const double c[3] = { node_coord[3 * (int64_t)nn], node_coord[3 * (int64_t)nn + 1], node_coord[3 * (int64_t)nn + 2] };
    // getDis: sqrt(pow(dx, 2) + pow(dy, 2) + pow(dz, 2)), left to right
    const double dx = __dsub_rn(c[0], s[0]), dy = __dsub_rn(c[1], s[1]), dz = __dsub_rn(c[2], s[2]);
    const double dis = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
    const double rad = (double)node_radius[nn];
    double o[3] = { s[0], s[1], s[2] };
    if (dis > rad) {
        const double steer = __ddiv_rn(rad, dis);
#pragma unroll
        for (int a = 0; a < 3; a++) o[a] = __dadd_rn(c[a], __dmul_rn(__dsub_rn(s[a], c[a]), steer));
        xyz[3 * j] = o[0]; xyz[3 * j + 1] = o[1]; xyz[3 * j + 2] = o[2];
    }
    q[j] = make_float4((float)o[0], (float)o[1], (float)o[2], 0.f);
    ok[j] = 1;
}

// candidates the expansion loop would not drop at once (corridor_finder.cpp:730: below the floor or radius < safety_margin),
// kept in sample order: count per tile, scan, write
#define PC_CAND_THREADS 256
#define PC_CAND_ITEMS 8
#define PC_CAND_TILE (PC_CAND_THREADS * PC_CAND_ITEMS)

__device__ __forceinline__ bool pc_cand_keep(int64_t j, const double *xyz, const float *radius, const uint8_t *ok, double z_l, double safety_margin)
{
    return ok[j] && !(xyz[3 * j + 2] < z_l || (double)radius[j] < safety_margin);
}

__global__ void __launch_bounds__(PC_CAND_THREADS)
pc_cand_count_kernel(const double *__restrict__ xyz, const float *__restrict__ radius, const uint8_t *__restrict__ ok, int64_t k,
                     double z_l, double safety_margin, uint32_t *__restrict__ tile_count)
{
    __shared__ uint32_t s_sum;
    if (threadIdx.x == 0) s_sum = 0u;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * PC_CAND_TILE;
    uint32_t c = 0;
    for (int i = 0; i < PC_CAND_ITEMS; i++) {
        const int64_t j = base + (int64_t)i * PC_CAND_THREADS + threadIdx.x;
        if (j < k && pc_cand_keep(j, xyz, radius, ok, z_l, safety_margin)) c++;
    }
    c = __reduce_add_sync(PC_FULL_MASK, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s_sum, c);
    __syncthreads();
    if (threadIdx.x == 0) tile_count[blockIdx.x] = s_sum;
}

// one CTA: exclusive scan of the tile counts in place, on top of *base (the candidates of the chunks in front of this one);
// *total = *base + the candidates of this chunk
__global__ void __launch_bounds__(1024)
pc_cand_scan_kernel(uint32_t *__restrict__ tile_count, int64_t n_tiles, const unsigned long long *__restrict__ base, unsigned long long *__restrict__ total)
{
    __shared__ unsigned long long s_warp[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t per = (n_tiles + blockDim.x - 1) / blockDim.x;
    const int64_t t0 = per * threadIdx.x, t1 = t0 + per < n_tiles ? t0 + per : n_tiles;
    unsigned long long mine = 0;
    for (int64_t t = t0; t < t1; t++) mine += tile_count[t];
    unsigned long long v = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const unsigned long long p = __shfl_up_sync(PC_FULL_MASK, v, d); if (lane >= d) v += p; }
    if (lane == 31) s_warp[w] = v;
    __syncthreads();
    if (w == 0) {
        unsigned long long t = s_warp[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const unsigned long long p = __shfl_up_sync(PC_FULL_MASK, t, d); if (lane >= d) t += p; }
        s_warp[lane] = t;
    }
    __syncthreads();
    unsigned long long run = *base + v - mine + (w > 0 ? s_warp[w - 1] : 0ull);
    for (int64_t t = t0; t < t1; t++) { const uint32_t c = tile_count[t]; tile_count[t] = (uint32_t)run; run += c; }
    if (threadIdx.x == blockDim.x - 1) *total = run;
}

struct pc_candidate_dev { double center[3]; float radius; int32_t nearest; };      // = pc_candidate (pc_index.h)

// planner-sized batches: count, scan and write in ONE CTA of 1024 threads (one launch instead of three)
#define PC_CAND_SMALL_MAX 32768
__global__ void __launch_bounds__(1024)
pc_cand_small_kernel(const double *__restrict__ xyz, const float *__restrict__ radius, const uint8_t *__restrict__ ok,
                     const int32_t *__restrict__ nearest, int64_t k, double z_l, double safety_margin,
                     const unsigned long long *__restrict__ base, unsigned long long *__restrict__ total,
                     pc_candidate_dev *__restrict__ out, uint64_t cap)
{
    __shared__ uint32_t s_warp[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned long long run = *base;
    for (int64_t j0 = 0; j0 < k; j0 += 1024) {
        const int64_t j = j0 + threadIdx.x;
        const bool keep = j < k && pc_cand_keep(j, xyz, radius, ok, z_l, safety_margin);
        const uint32_t bal = __ballot_sync(PC_FULL_MASK, keep);
        if (lane == 0) s_warp[w] = __popc(bal);
        __syncthreads();
        uint32_t before = 0, all = 0;
#pragma unroll
        for (int x = 0; x < 32; x++) { const uint32_t c = s_warp[x]; if (x < w) before += c; all += c; }
        if (keep) {
            const uint64_t dst = run + before + __popc(bal & ((1u << lane) - 1u));
            if (dst < cap) {
                pc_candidate_dev c;
                c.center[0] = xyz[3 * j]; c.center[1] = xyz[3 * j + 1]; c.center[2] = xyz[3 * j + 2];
                c.radius = radius[j]; c.nearest = nearest[j];
                out[dst] = c;
            }
        }
        run += all;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = run;
}

__global__ void __launch_bounds__(PC_CAND_THREADS)
pc_cand_write_kernel(const double *__restrict__ xyz, const float *__restrict__ radius, const uint8_t *__restrict__ ok,
                     const int32_t *__restrict__ nearest, int64_t k, double z_l, double safety_margin,
                     const uint32_t *__restrict__ tile_offset, pc_candidate_dev *__restrict__ out, uint64_t cap)
{
    __shared__ uint32_t s_warp[PC_CAND_THREADS / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t base = (int64_t)blockIdx.x * PC_CAND_TILE;
    uint32_t run = tile_offset[blockIdx.x];
    // sample order: item i of all threads comes before item i + 1
    for (int i = 0; i < PC_CAND_ITEMS; i++) {
        const int64_t j = base + (int64_t)i * PC_CAND_THREADS + threadIdx.x;
        const bool keep = j < k && pc_cand_keep(j, xyz, radius, ok, z_l, safety_margin);
        const uint32_t bal = __ballot_sync(PC_FULL_MASK, keep);
        if (lane == 0) s_warp[w] = __popc(bal);
        __syncthreads();
        uint32_t before = 0, all = 0;
#pragma unroll
        for (int x = 0; x < PC_CAND_THREADS / 32; x++) { const uint32_t c = s_warp[x]; if (x < w) before += c; all += c; }
        if (keep) {
            const uint64_t dst = (uint64_t)run + before + __popc(bal & ((1u << lane) - 1u));
            if (dst < cap) {
                pc_candidate_dev c;
                c.center[0] = xyz[3 * j]; c.center[1] = xyz[3 * j + 1]; c.center[2] = xyz[3 * j + 2];
                c.radius = radius[j]; c.nearest = nearest[j];
                out[dst] = c;
            }
        }
        run += all;
        __syncthreads();
    }
}
