// lbvh.cuh -- the prefix-split tree of the index: per-node range / split function and record layout (host + device).
//
// Node ranges and splits of a binary radix tree over SORTED curve keys, one independent computation per inner node
// (T. Karras, "Maximizing Parallelism in the Construction of BVHs, Octrees, and k-d Trees", HPG 2012, section 3): every
// node is split where the highest differing bit of its first and last key flips, so node boundaries coincide with the
// boundaries of the curve's cells.  Equal keys are told apart by their position (the key of element i is the pair
// (key[i], i)), so all elements are distinct and the tree is well defined for clouds with coincident points.
//
// The functions compile for the host as well (tests/c/lbvh_check.cpp checks them against a recursive top-down
// construction on the CPU); pc_tree_nodes_kernel (build_kernels.cuh) calls pc_lbvh_node once per thread.
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

// ---- record layout and single-query walk (host + device), see DESIGN.md section 3 ---------------------------------------
// One 64-byte record per inner node i, four float4: [min0 | max0 | min1 | max1] = the boxes of its two children; the
// otherwise unused .w words hold the children: min.w = child reference, max.w = number of points when the child is a leaf.
// A child whose range holds <= PC_LBVH_LEAF points is a LEAF: the reference is PC_REF_LEAF | (index of its first point in
// curve order) and the walk scans PC_LBVH_LEAF consecutive points from there -- points past the leaf's own count are real
// points of the cloud too, so scanning them cannot break exactness (the point array is padded with copies of the last
// point).  Inner nodes whose own range is that small are never referenced; their records stay unused.
#ifndef PC_LBVH_LEAF
#define PC_LBVH_LEAF 4
#endif
#define PC_REF_LEAF 0x80000000u

struct pc_f4 { float x, y, z, w; };          // same layout as CUDA's float4

PC_HD uint32_t pc_f2u(float f) { union { float f; uint32_t u; } v; v.f = f; return v.u; }
PC_HD float pc_u2f(uint32_t u) { union { float f; uint32_t u; } v; v.u = u; return v.f; }

// children of inner node i from its (first, last, split) and the PC_LBVH_LEAF rule
PC_HD void pc_lbvh_children(int64_t first, int64_t last, int64_t split, uint32_t *ref0, uint32_t *count0, uint32_t *ref1, uint32_t *count1)
{
    const int64_t n0 = split - first + 1, n1 = last - split;
    *ref0 = n0 <= PC_LBVH_LEAF ? (PC_REF_LEAF | (uint32_t)first) : (uint32_t)split;
    *ref1 = n1 <= PC_LBVH_LEAF ? (PC_REF_LEAF | (uint32_t)(split + 1)) : (uint32_t)(split + 1);
    *count0 = (uint32_t)n0; *count1 = (uint32_t)n1;
}

PC_HD float pc_lbvh_box_d2(const pc_f4 lo, const pc_f4 hi, float qx, float qy, float qz)
{
    float dx = lo.x - qx, ex = qx - hi.x; if (ex > dx) dx = ex; if (!(dx > 0.f)) dx = 0.f;
    float dy = lo.y - qy, ey = qy - hi.y; if (ey > dy) dy = ey; if (!(dy > 0.f)) dy = 0.f;
    float dz = lo.z - qz, ez = qz - hi.z; if (ez > dz) dz = ez; if (!(dz > 0.f)) dz = 0.f;
    return dx * dx + dy * dy + dz * dz;          // any rounding here is covered by the 2^-20 slack of the filter threshold
}

// Exact nearest point of one query: fp32 filter against thr, fp64 decision in the reference's operation order, ties -> lowest
// original index (the rule of DESIGN.md section 2).  `root` is a child reference (the whole tree may be a single leaf).
// On entry *thr is the initial bound on d2 (FLT_MAX or the radius bound), *best_d2 = +inf, *best_idx = -1.
template <typename RoundUp>
PC_HD void pc_lbvh_nearest(const pc_f4 *records, const pc_f4 *points, uint32_t root, float qx, float qy, float qz,
                           double *best_d2, int32_t *best_idx, float *thr, RoundUp thr_from, int64_t *visits = nullptr)
{
    uint32_t stack_ref[64];
    float stack_d[64];
    int sp = 0;
    uint32_t ref = root;
    for (;;) {
        if (ref & PC_REF_LEAF) {
            const pc_f4 *p = points + (ref & 0x7fffffffu);
            for (int i = 0; i < PC_LBVH_LEAF; i++) {
                const float dx = p[i].x - qx, dy = p[i].y - qy, dz = p[i].z - qz;
                const float d = dx * dx + dy * dy + dz * dz;
                if (d <= *thr) {
                    const double ex = (double)p[i].x - (double)qx, ey = (double)p[i].y - (double)qy, ez = (double)p[i].z - (double)qz;
                    double e = ex * ex; e = e + ey * ey; e = e + ez * ez;
                    const int32_t id = (int32_t)pc_f2u(p[i].w);
                    if (e < *best_d2 || (e == *best_d2 && (uint32_t)id < (uint32_t)*best_idx)) {
                        *best_d2 = e; *best_idx = id;
                        const float t = thr_from(e);
                        if (t < *thr) *thr = t;
                    }
                }
            }
            ref = 0xffffffffu;                                  // nothing to descend into
        } else {
            if (visits) (*visits)++;
            const pc_f4 *r = records + 4 * (int64_t)ref;
            const float d0 = pc_lbvh_box_d2(r[0], r[1], qx, qy, qz), d1 = pc_lbvh_box_d2(r[2], r[3], qx, qy, qz);
            const uint32_t r0 = pc_f2u(r[0].w), r1 = pc_f2u(r[2].w);
            const bool first0 = d0 <= d1;
            const uint32_t rn = first0 ? r0 : r1, rf = first0 ? r1 : r0;
            const float dn = first0 ? d0 : d1, df = first0 ? d1 : d0;
            if (df <= *thr) { stack_ref[sp] = rf; stack_d[sp] = df; sp++; }
            ref = dn <= *thr ? rn : 0xffffffffu;
        }
        if (ref != 0xffffffffu) continue;
        bool found = false;
        while (sp > 0) { sp--; if (stack_d[sp] <= *thr) { ref = stack_ref[sp]; found = true; break; } }
        if (!found) break;
    }
}
