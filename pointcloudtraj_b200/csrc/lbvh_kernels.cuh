// lbvh_kernels.cuh -- EXPERIMENTAL (off unless PC_LBVH=1 is set when the index is created): the prefix-split tree of DESIGN.md
// section 8 as a SECOND tree over the same curve-ordered points, used by the ordered-batch packet kernels only.  The layout,
// the per-node range / split function, the bottom-up box fit and the packet walk with mixed leaf / inner children were
// validated on the CPU first (lbvh.cuh, tests/test_lbvh_cpu.py); this file is their CUDA form.
#pragma once
#include "lbvh.cuh"
#include "query_kernels.cuh"

#if PC_LBVH_LEAF != PC_LEAF
#error "the prefix-split tree shares the point array of the implicit tree: PC_LBVH_LEAF must equal PC_LEAF"
#endif

#define PC_LBVH_UNUSED 0xffffffffu

// K1: one thread per inner node of the radix tree: range, split, child references (the .w words of its record), parent links.
// Thread 0 also pads the point array: a leaf scan reads PC_LEAF consecutive points from any start.
template <typename KeyT>
__global__ void __launch_bounds__(256)
pc_lbvh_nodes_kernel(const KeyT *__restrict__ keys, int64_t n, int64_t n_leaves, float4 *__restrict__ rec, int32_t *__restrict__ parent,
                     float4 *__restrict__ points)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        parent[0] = -1;
        const float4 last = points[n - 1];
        for (int j = 0; j < PC_LEAF; j++) points[n_leaves * PC_LEAF + j] = last;
    }
    if (i >= n - 1) return;
    int64_t f, l, s;
    pc_lbvh_node(keys, n, i, &f, &l, &s);
    float *w = reinterpret_cast<float *>(rec + 4 * i);
    if (l - f + 1 <= PC_LBVH_LEAF) { w[3] = __uint_as_float(PC_LBVH_UNUSED); return; }   // collapsed into a leaf of its parent
    uint32_t r0, c0, r1, c1;
    pc_lbvh_children(f, l, s, &r0, &c0, &r1, &c1);
    w[3] = __uint_as_float(r0); w[7] = __uint_as_float(c0); w[11] = __uint_as_float(r1); w[15] = __uint_as_float(c1);
    if (!(r0 & PC_REF_LEAF)) parent[r0] = (int32_t)i;
    if (!(r1 & PC_REF_LEAF)) parent[r1] = (int32_t)i;
}

__device__ __forceinline__ void pc_lbvh_leaf_box(const float4 *__restrict__ points, uint32_t first, uint32_t count, float *lo, float *hi)
{
    float lx = INFINITY, ly = INFINITY, lz = INFINITY, hx = -INFINITY, hy = -INFINITY, hz = -INFINITY;
    for (uint32_t k = 0; k < count; k++) {
        const float4 p = points[first + k];
        lx = fminf(lx, p.x); ly = fminf(ly, p.y); lz = fminf(lz, p.z);      // fminf / fmaxf ignore NaN coordinates
        hx = fmaxf(hx, p.x); hy = fmaxf(hy, p.y); hz = fmaxf(hz, p.z);
    }
    lo[0] = lx; lo[1] = ly; lo[2] = lz; hi[0] = hx; hi[1] = hy; hi[2] = hz;
}

// K2: bottom-up box fit.  One thread per used inner node boxes its LEAF children; whichever thread brings a node's arrival
// counter to 2 merges the node's two child boxes into the slot the node has in its parent's record, and carries on upwards.
__global__ void __launch_bounds__(256)
pc_lbvh_fit_kernel(float4 *rec, const float4 *__restrict__ points, const int32_t *__restrict__ parent, int *ready, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    volatile float *w = reinterpret_cast<volatile float *>(rec + 4 * i);
    const uint32_t r0 = __float_as_uint(w[3]);
    if (r0 == PC_LBVH_UNUSED) return;
    const uint32_t c0 = __float_as_uint(w[7]), r1 = __float_as_uint(w[11]), c1 = __float_as_uint(w[15]);
    int add = 0;
    float lo[3], hi[3];
    if (r0 & PC_REF_LEAF) {
        pc_lbvh_leaf_box(points, r0 & 0x7fffffffu, c0, lo, hi);
        w[0] = lo[0]; w[1] = lo[1]; w[2] = lo[2]; w[4] = hi[0]; w[5] = hi[1]; w[6] = hi[2];
        add++;
    }
    if (r1 & PC_REF_LEAF) {
        pc_lbvh_leaf_box(points, r1 & 0x7fffffffu, c1, lo, hi);
        w[8] = lo[0]; w[9] = lo[1]; w[10] = lo[2]; w[12] = hi[0]; w[13] = hi[1]; w[14] = hi[2];
        add++;
    }
    int64_t node = i;
    while (add > 0) {
        __threadfence();                                       // my box writes before my arrival
        const int old = atomicAdd(&ready[node], add);
        if (old + add < 2) break;                              // the other child's thread completes this node
        __threadfence();                                       // the other child's box writes after its arrival
        const int32_t par = parent[node];
        if (par < 0) break;                                    // the root is complete
        volatile float *c = reinterpret_cast<volatile float *>(rec + 4 * node);
        const float mlx = fminf(c[0], c[8]), mly = fminf(c[1], c[9]), mlz = fminf(c[2], c[10]);
        const float mhx = fmaxf(c[4], c[12]), mhy = fmaxf(c[5], c[13]), mhz = fmaxf(c[6], c[14]);
        volatile float *p = reinterpret_cast<volatile float *>(rec + 4 * (int64_t)par);
        const int slot = __float_as_uint(p[3]) == (uint32_t)node ? 0 : 8;     // left child = first reference of the parent
        p[slot + 0] = mlx; p[slot + 1] = mly; p[slot + 2] = mlz;
        p[slot + 4] = mhx; p[slot + 5] = mhy; p[slot + 6] = mhz;
        node = par;
        add = 1;
    }
}

// leaf scan from an arbitrary (unaligned) start: 128-bit loads
__device__ __forceinline__ void pc_scan_leaf2_u(const float4 *__restrict__ pts, const float qa[3], const float qb[3], pc_best &ba, pc_best &bb)
{
    float4 p[PC_LEAF];
    float da[PC_LEAF], db[PC_LEAF];
#pragma unroll
    for (int i = 0; i < PC_LEAF; i++) p[i] = __ldg(pts + i);
    float mina = FLT_MAX, minb = FLT_MAX;
#pragma unroll
    for (int i = 0; i < PC_LEAF; i++) {
        float dx = p[i].x - qa[0], dy = p[i].y - qa[1], dz = p[i].z - qa[2];
        da[i] = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        mina = fminf(mina, da[i]);
        dx = p[i].x - qb[0]; dy = p[i].y - qb[1]; dz = p[i].z - qb[2];
        db[i] = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        minb = fminf(minb, db[i]);
    }
    if (mina <= ba.thr) {
#pragma unroll
        for (int i = 0; i < PC_LEAF; i++) pc_consider(p[i], da[i], qa[0], qa[1], qa[2], ba);
    }
    if (minb <= bb.thr) {
#pragma unroll
        for (int i = 0; i < PC_LEAF; i++) pc_consider(p[i], db[i], qb[0], qb[1], qb[2], bb);
    }
}

// The 64-query packet walk of pc_packet2_traverse over the prefix-split tree: children come from the record, and the two
// children of a node may be a leaf and an inner node.  The warp stack holds 96 entries (three per lane): the tree is as deep
// as the keys are long plus the position bits that separate coincident points.
__device__ __forceinline__ void pc_packet2_traverse_lbvh(const pc_tree &T, const float qa[3], const float qb[3], pc_best &ba, pc_best &bb, int lane)
{
    if (__ballot_sync(PC_FULL_MASK, ba.thr >= 0.f || bb.thr >= 0.f) == 0) return;
    uint32_t e0 = 0, e1 = 0, e2 = 0;
    int sp = 0;
    uint32_t ref = T.lbvh_root;
    if (ref & PC_REF_LEAF) { pc_scan_leaf2_u(T.points + (ref & 0x7fffffffu), qa, qb, ba, bb); return; }
    for (;;) {
        const float4 *pair = T.lbvh + 4ull * ref;
        float4 lo0, hi0, lo1, hi1;
        pc_load_box(pair, lo0, hi0);
        pc_load_box(pair + 2, lo1, hi1);
        const float a0 = pc_box_d2(lo0, hi0, qa[0], qa[1], qa[2]), a1 = pc_box_d2(lo1, hi1, qa[0], qa[1], qa[2]);
        const float b0 = pc_box_d2(lo0, hi0, qb[0], qb[1], qb[2]), b1 = pc_box_d2(lo1, hi1, qb[0], qb[1], qb[2]);
        const bool wa0 = a0 <= ba.thr, wa1 = a1 <= ba.thr, wb0 = b0 <= bb.thr, wb1 = b1 <= bb.thr;
        const uint32_t w0 = __ballot_sync(PC_FULL_MASK, wa0 || wb0);
        const uint32_t w1 = __ballot_sync(PC_FULL_MASK, wa1 || wb1);
        uint32_t next = 0xffffffffu;
        if (w0 | w1) {
            const bool both = w0 != 0 && w1 != 0;
            bool first0 = w1 == 0;
            if (both) {
                const uint32_t ia = __ballot_sync(PC_FULL_MASK, wa0 || wa1);
                const uint32_t pa = __ballot_sync(PC_FULL_MASK, a0 <= a1) & ia;
                first0 = 2 * __popc(pa) >= __popc(ia);
            }
            const uint32_t r0 = __float_as_uint(lo0.w), r1 = __float_as_uint(lo1.w);
            const uint32_t rn = first0 ? r0 : r1, rf = first0 ? r1 : r0;
            if (rn & PC_REF_LEAF) pc_scan_leaf2_u(T.points + (rn & 0x7fffffffu), qa, qb, ba, bb);
            else next = rn;
            if (both) {
                if (rf & PC_REF_LEAF) {
                    if (__ballot_sync(PC_FULL_MASK, (first0 ? a1 : a0) <= ba.thr || (first0 ? b1 : b0) <= bb.thr))
                        pc_scan_leaf2_u(T.points + (rf & 0x7fffffffu), qa, qb, ba, bb);
                } else if (next != 0xffffffffu) {
                    if (lane == (sp & 31)) { if (sp < 32) e0 = rf; else if (sp < 64) e1 = rf; else e2 = rf; }
                    sp++;
                } else next = rf;
            }
        }
        if (next != 0xffffffffu) { ref = next; continue; }
        if (sp == 0) break;
        sp--;
        ref = __shfl_sync(PC_FULL_MASK, sp < 32 ? e0 : (sp < 64 ? e1 : e2), sp & 31);
    }
}

template <int KIND>
__global__ void __launch_bounds__(PC_QUERY_THREADS)
pc_query_packet2_lbvh_kernel(pc_tree T, pc_radius_dev R, const float *__restrict__ q, int64_t m, int qstride,
                             const uint32_t *__restrict__ perm, const unsigned long long *__restrict__ m_eff,
                             int32_t *__restrict__ out_idx, float *__restrict__ out_f)
{
    const int lane = threadIdx.x & 31;
    const long long m_search = m_eff ? (long long)*m_eff : (long long)m;
    const long long warp_id = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long base = warp_id * 64;
    if (base >= m_search) return;
    float qv[2][3] = { { 0.f, 0.f, 0.f }, { 0.f, 0.f, 0.f } };
    uint32_t k[2] = { 0u, 0u };
    bool valid[2];
    pc_best b[2];
#pragma unroll
    for (int j = 0; j < 2; j++) {
        const long long t = base + 32 * j + lane;
        valid[j] = t < m_search;
        b[j].d2 = INFINITY; b[j].idx = -1; b[j].thr = -1.0f;
        if (valid[j]) {
            k[j] = perm ? perm[t] : (uint32_t)t;
            const float *qq = q + (size_t)k[j] * qstride;
            qv[j][0] = qq[0]; qv[j][1] = qq[1]; qv[j][2] = qq[2];
            bool search = T.n_points > 0;
            if (KIND == PC_KIND_RADIUS && search && !m_eff && pc_radius_early_out((double)qv[j][0], (double)qv[j][1], (double)qv[j][2], R)) search = false;
            if (search) b[j].thr = (KIND == PC_KIND_RADIUS) ? R.bound_thr : FLT_MAX;
            else { pc_write_trivial<KIND>(R, k[j], out_idx, out_f); valid[j] = false; }
        }
    }
    pc_packet2_traverse_lbvh(T, qv[0], qv[1], b[0], b[1], lane);
#pragma unroll
    for (int j = 0; j < 2; j++)
        if (valid[j]) pc_write_result<KIND>(R, b[j], k[j], out_idx, out_f);
}
