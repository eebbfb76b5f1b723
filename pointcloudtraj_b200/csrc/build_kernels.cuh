// build_kernels.cuh -- index construction kernels (bbox, curve keys, leaf gather, bounding-box tree).
//
// Index layout in HBM: ONE float4 array, P = 2^k >= max(2, ceil(n / 2)) leaf slots, heap numbering (root = node 1,
// children of i are 2i and 2i+1, leaf j is node P + j):
//   boxes  tree[0 .. 4P)  : node i's box is tree[2i] = min (xyz), tree[2i+1] = max (xyz); the boxes of the two children
//                           of i are therefore the aligned 64-byte record tree[4i .. 4i+3]
//   points tree[4P .. 6P) : leaf j holds points[2j], points[2j+1] in curve order, (x, y, z, original index as int
//                           bits); when n is odd the last slot repeats the last real point
// A traversal step loads the record of the node it visits: a box pair (64 B) or a leaf's two points (32 B).
// Unused box slots hold the empty box (min = +inf, max = -inf).
// replaces struct kdtree / struct kdnode / struct kdhyperrect (Utils/kdtree/src/kdtree.c:56-80) and
// hyperrect_extend (kdtree.c:729-741).
#pragma once
#include "common.cuh"

#define PC_BUILD_THREADS 256

// ---- bounding box of the cloud: bbox[0..2] = ordered(min), bbox[3..5] = ordered(max) ----------------
__global__ void __launch_bounds__(PC_BUILD_THREADS)
pc_bbox_kernel(const float *__restrict__ xyz, int64_t n, int stride, uint32_t *__restrict__ bbox)
{
    float lo[3] = { INFINITY, INFINITY, INFINITY }, hi[3] = { -INFINITY, -INFINITY, -INFINITY };
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float *p = xyz + i * stride;
#pragma unroll
        for (int a = 0; a < 3; a++) {
            float v = p[a];
            lo[a] = fminf(lo[a], v);   // fminf/fmaxf drop NaNs
            hi[a] = fmaxf(hi[a], v);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; a++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(PC_FULL_MASK, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(PC_FULL_MASK, hi[a], o));
        }
    }
    __shared__ float s_lo[PC_BUILD_THREADS / 32][3], s_hi[PC_BUILD_THREADS / 32][3];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
#pragma unroll
        for (int a = 0; a < 3; a++) { s_lo[warp][a] = lo[a]; s_hi[warp][a] = hi[a]; }
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        int a = threadIdx.x;
        float l = s_lo[0][a], h = s_hi[0][a];
        for (int w = 1; w < PC_BUILD_THREADS / 32; w++) { l = fminf(l, s_lo[w][a]); h = fmaxf(h, s_hi[w][a]); }
        atomicMin(&bbox[a], pc_float_to_ordered(l));
        atomicMax(&bbox[3 + a], pc_float_to_ordered(h));
    }
}

// decode the bbox into the quantisation frame (device-side, no host round trip)
__device__ __forceinline__ pc_frame pc_make_frame(const uint32_t *bbox, int bits)
{
    pc_frame f;
    float ext = 0.0f;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        float lo = pc_ordered_to_float(bbox[a]), hi = pc_ordered_to_float(bbox[3 + a]);
        f.lo[a] = lo;
        ext = fmaxf(ext, hi - lo);
    }
    f.max_cell = (1u << bits) - 1u;
    // cells of edge ext / 2^bits; a degenerate cloud (ext == 0) maps everything to cell 0
    f.inv_cell = (ext > 0.0f && ext < INFINITY) ? ((float)(1u << bits) * (1.0f - 1e-6f)) / ext : 0.0f;
    return f;
}

// Order of the cloud along a space-filling curve.  The tree above it is implicit -- aligned groups of 2^k consecutive
// leaves -- so its boxes are only as tight as consecutive stretches of the curve are compact.  A Hilbert stretch always is
// (consecutive cells are adjacent); a Morton stretch that straddles an octant boundary is not.  Measured on the bench
// workload: radius search 1.84 -> 1.29 ms, unbounded nearest 7.10 -> 4.15 ms per 10 M queries, identical results
// (profiles/r1_sweep9_cloud_hilbert_order.txt).
#ifndef PC_POINT_CURVE
#define PC_POINT_CURVE 1     // 0 = Morton, 1 = Hilbert
#endif
template <typename KeyT>
__global__ void __launch_bounds__(PC_BUILD_THREADS)
pc_keygen_kernel(const float *__restrict__ xyz, int64_t n, int stride, const uint32_t *__restrict__ bbox, int bits,
                 KeyT *__restrict__ keys, uint32_t *__restrict__ vals)
{
    const pc_frame f = pc_make_frame(bbox, bits);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float *p = xyz + i * stride;
#if PC_POINT_CURVE == 1
        if (sizeof(KeyT) == 4) keys[i] = (KeyT)pc_hilbert30(p[0], p[1], p[2], f);
        else keys[i] = (KeyT)pc_hilbert63(p[0], p[1], p[2], f, bits);
#else
        if (sizeof(KeyT) == 4) keys[i] = (KeyT)pc_morton30(p[0], p[1], p[2], f);
        else keys[i] = (KeyT)pc_morton63(p[0], p[1], p[2], f);
#endif
        vals[i] = (uint32_t)i;
    }
}

// ---- leaves: gather the cloud into curve order (one thread per point slot) and box every PC_LEAF slots ----
__global__ void __launch_bounds__(PC_BUILD_THREADS)
pc_leaf_kernel(const float *__restrict__ xyz, int stride, const uint32_t *__restrict__ order, int64_t n,
               int64_t n_leaves, int64_t P, float4 *__restrict__ points, float4 *__restrict__ nodes)
{
    const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // grid covers 8 * n_leaves (rounded up to 8 lanes)
    const bool in_range = slot < n_leaves * PC_LEAF;
    int64_t s = slot < n ? slot : n - 1;                                    // tail slots repeat the last real point
    float x = 0.f, y = 0.f, z = 0.f; uint32_t src = 0;
    if (in_range) {
        src = order[s];
        const float *p = xyz + (int64_t)src * stride;
        x = p[0]; y = p[1]; z = p[2];
        points[slot] = make_float4(x, y, z, __uint_as_float(src));
    }
    // NaN coordinates must not poison the box: fminf/fmaxf ignore them
    float lx = in_range ? x : INFINITY, ly = in_range ? y : INFINITY, lz = in_range ? z : INFINITY;
    float hx = in_range ? x : -INFINITY, hy = in_range ? y : -INFINITY, hz = in_range ? z : -INFINITY;
#pragma unroll
    for (int o = 1; o < PC_LEAF; o <<= 1) {
        lx = fminf(lx, __shfl_xor_sync(PC_FULL_MASK, lx, o)); hx = fmaxf(hx, __shfl_xor_sync(PC_FULL_MASK, hx, o));
        ly = fminf(ly, __shfl_xor_sync(PC_FULL_MASK, ly, o)); hy = fmaxf(hy, __shfl_xor_sync(PC_FULL_MASK, hy, o));
        lz = fminf(lz, __shfl_xor_sync(PC_FULL_MASK, lz, o)); hz = fmaxf(hz, __shfl_xor_sync(PC_FULL_MASK, hz, o));
    }
    if ((threadIdx.x & (PC_LEAF - 1)) == 0) {
        int64_t leaf = slot / PC_LEAF;
        // the slots after the last leaf, up to the next multiple of 4, are written as empty boxes: they are the siblings /
        // cousins of the last leaf that the binary (2 boxes per visit) and the 4-ary (4 boxes per visit) walks read
        if (leaf < ((n_leaves + 4) & ~(int64_t)3) && leaf < P) {
            nodes[2 * (P + leaf)] = make_float4(lx, ly, lz, 0.f);
            nodes[2 * (P + leaf) + 1] = make_float4(hx, hy, hz, 0.f);
        }
    }
}

// ---- upper levels: each CTA folds 2*PC_UP_THREADS nodes of level `lvl0` into up to PC_UP_LEVELS levels above ----
// Level l has base id P >> l and cnt_l = ceil(cnt_{l-1} / 2) real nodes (cnt_0 = n_leaves); the nodes from cnt_l up to the
// next multiple of 4 (the possible sibling and cousins of its last real node) are written as empty boxes.
#define PC_UP_THREADS 128
#define PC_UP_LEVELS 8   // log2(2 * PC_UP_THREADS)

__global__ void __launch_bounds__(PC_UP_THREADS)
pc_upper_kernel(float4 *__restrict__ nodes, int64_t P, int lvl0, int64_t cnt0, int n_levels)
{
    __shared__ float4 s_lo[PC_UP_THREADS], s_hi[PC_UP_THREADS];
    const int t = threadIdx.x;
    int64_t child_first = (int64_t)blockIdx.x * (2 * PC_UP_THREADS);   // index inside level lvl0
    int64_t cnt_child = cnt0;
    float4 lo, hi;
    for (int s = 1; s <= n_levels; s++) {
        const int lvl = lvl0 + s;
        const int64_t base = P >> lvl;                 // id of the first node of this level
        const int64_t cnt = (cnt_child + 1) >> 1;      // real nodes on this level
        const int width = (2 * PC_UP_THREADS) >> s;    // nodes of this level owned by this CTA
        const int64_t k = (child_first >> s) + t;      // node index inside the level
        if (t < width) {
            float4 alo, ahi, blo, bhi;
            if (s == 1) {
                const int64_t cbase = P >> lvl0;
                const int64_t c = 2 * k;
                const float4 e_lo = make_float4(INFINITY, INFINITY, INFINITY, 0.f), e_hi = make_float4(-INFINITY, -INFINITY, -INFINITY, 0.f);
                if (c < cnt_child) { alo = nodes[2 * (cbase + c)]; ahi = nodes[2 * (cbase + c) + 1]; } else { alo = e_lo; ahi = e_hi; }
                if (c + 1 < cnt_child) { blo = nodes[2 * (cbase + c + 1)]; bhi = nodes[2 * (cbase + c + 1) + 1]; } else { blo = e_lo; bhi = e_hi; }
            } else {
                alo = s_lo[2 * t]; ahi = s_hi[2 * t]; blo = s_lo[2 * t + 1]; bhi = s_hi[2 * t + 1];
            }
            lo = make_float4(fminf(alo.x, blo.x), fminf(alo.y, blo.y), fminf(alo.z, blo.z), 0.f);
            hi = make_float4(fmaxf(ahi.x, bhi.x), fmaxf(ahi.y, bhi.y), fmaxf(ahi.z, bhi.z), 0.f);
        }
        __syncthreads();   // everyone has read the previous level from shared memory
        if (t < width) {
            s_lo[t] = lo; s_hi[t] = hi;
            if (k < ((cnt + 4) & ~(int64_t)3) && k < base) {   // pads up to the next multiple of 4; base == number of slots on this level
                nodes[2 * (base + k)] = lo;
                nodes[2 * (base + k) + 1] = hi;
            }
        }
        __syncthreads();
        cnt_child = cnt;
    }
}
