// lbvh.cuh -- groundwork for the prefix-split tree planned in DESIGN.md section 8 (not yet used by libpcindex.so).
//
// Node ranges and splits of a binary radix tree over SORTED curve keys, one independent computation per inner node
// (T. Karras, "Maximizing Parallelism in the Construction of BVHs, Octrees, and k-d Trees", HPG 2012, section 3): every
// node is split where the highest differing bit of its first and last key flips, so node boundaries coincide with the
// boundaries of the curve's cells.  Equal keys are told apart by their position (the key of element i is the pair
// (key[i], i)), so all elements are distinct and the tree is well defined for clouds with coincident points.
//
// The functions compile for the host as well (tests/c/lbvh_check.cpp checks them against a recursive top-down
// construction on the CPU); the CUDA build kernel of the next round calls pc_lbvh_node once per thread.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PC_HD __host__ __device__ __forceinline__
#else
#define PC_HD static inline
#endif

PC_HD int pc_lbvh_clz64(uint64_t v)
{
#if defined(__CUDA_ARCH__)
    return __clzll((long long)v);
#else
    return v ? __builtin_clzll(v) : 64;
#endif
}

// length of the common prefix of elements i and j in the 128-bit string (key, position); -1 when j is out of range
template <typename KeyT>
PC_HD int pc_lbvh_delta(const KeyT *keys, int64_t n, int64_t i, int64_t j)
{
    if (j < 0 || j >= n) return -1;
    const uint64_t a = (uint64_t)keys[i], b = (uint64_t)keys[j];
    if (a != b) return pc_lbvh_clz64(a ^ b);
    return 64 + pc_lbvh_clz64((uint64_t)i ^ (uint64_t)j);
}

// Inner node i (0 <= i <= n - 2) of the radix tree over n >= 2 sorted elements: its element range [first, last] and its
// split: the left child covers [first, split], the right child [split + 1, last].  A child that covers one element is
// that element (a leaf); otherwise it is the inner node with index `split` (left) or `split + 1` (right).
template <typename KeyT>
PC_HD void pc_lbvh_node(const KeyT *keys, int64_t n, int64_t i, int64_t *first, int64_t *last, int64_t *split)
{
    const int d = pc_lbvh_delta(keys, n, i, i + 1) - pc_lbvh_delta(keys, n, i, i - 1) >= 0 ? 1 : -1;
    const int dmin = pc_lbvh_delta(keys, n, i, i - d);
    int64_t lmax = 2;
    while (pc_lbvh_delta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
    int64_t l = 0;
    for (int64_t t = lmax / 2; t >= 1; t /= 2)
        if (pc_lbvh_delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int64_t j = i + l * d;
    const int dnode = pc_lbvh_delta(keys, n, i, j);
    int64_t s = 0, t = l;
    do {
        t = (t + 1) >> 1;
        if (pc_lbvh_delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    const int64_t g = i + s * d + (d < 0 ? -1 : 0);
    *first = i < j ? i : j;
    *last = i < j ? j : i;
    *split = g;
}
