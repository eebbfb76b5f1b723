// query_kernels.cuh -- exact nearest / radius / range queries: one thread per query, near-first
// descent of the bounding-box tree with a short per-thread stack.
//
// Replaces kd_nearest_i / kd_nearest3 (Utils/kdtree/src/kdtree.c:345-457,493-500), find_nearest /
// kd_nearest_range3 (kdtree.c:262-293,595-602) and the epilogue of safeRegionRrtStar::radiusSearch
// (Planner/src/corridor_finder.cpp:113-133).
//
// Exactness: boxes and points are filtered in fp32 against `thr`, an upper bound of the current best fp64
// distance inflated by 2^-20 (fp32 evaluation error of d2 is < 2^-22 relative, so no candidate whose fp64
// distance is <= best can be filtered out); survivors are re-evaluated in fp64 with un-fused
// __dsub_rn/__dmul_rn/__dadd_rn in the reference's operation order and ranked by (d2, original index).
#pragma once
#include "common.cuh"
#include "build_kernels.cuh"

#define PC_QUERY_THREADS 128
#define PC_STACK 32
#define PC_THR_SLACK 1.00000095367431640625f   // 1 + 2^-20

struct pc_tree {
    const float4 *__restrict__ nodes;
    const float4 *__restrict__ points;
    int64_t n_points;
    uint32_t P;            // leaf base (power of two >= 2)
};

__device__ __forceinline__ float pc_box_d2(const float4 lo, const float4 hi, float qx, float qy, float qz)
{
    float dx = fmaxf(fmaxf(lo.x - qx, qx - hi.x), 0.0f);
    float dy = fmaxf(fmaxf(lo.y - qy, qy - hi.y), 0.0f);
    float dz = fmaxf(fmaxf(lo.z - qz, qz - hi.z), 0.0f);
    return fmaf(dz, dz, fmaf(dy, dy, dx * dx));
}

// the reference's fp64 expression: s = 0; s += (px-qx)^2; s += (py-qy)^2; s += (pz-qz)^2  (kdtree.c:379-382)
__device__ __forceinline__ double pc_exact_d2(float px, float py, float pz, double qx, double qy, double qz)
{
    double dx = __dsub_rn((double)px, qx), dy = __dsub_rn((double)py, qy), dz = __dsub_rn((double)pz, qz);
    double s = __dmul_rn(dx, dx);
    s = __dadd_rn(s, __dmul_rn(dy, dy));
    s = __dadd_rn(s, __dmul_rn(dz, dz));
    return s;
}

__device__ __forceinline__ float pc_thr_from(double best)
{
    return __fmul_ru(__double2float_ru(best), PC_THR_SLACK);
}

struct pc_best {
    double d2;     // +inf until a point is accepted
    int32_t idx;   // -1 until a point is accepted
    float thr;     // fp32 filter threshold (inclusive)
};

__device__ __forceinline__ void pc_scan_leaf(const float4 *__restrict__ pts, float qx, float qy, float qz,
                                             double qxd, double qyd, double qzd, pc_best &b)
{
    float4 p[PC_LEAF];
#pragma unroll
    for (int i = 0; i < PC_LEAF; i++) p[i] = __ldg(pts + i);
#pragma unroll
    for (int i = 0; i < PC_LEAF; i++) {
        float dx = p[i].x - qx, dy = p[i].y - qy, dz = p[i].z - qz;
        float d = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        if (d <= b.thr) {
            double e = pc_exact_d2(p[i].x, p[i].y, p[i].z, qxd, qyd, qzd);
            int32_t id = __float_as_int(p[i].w);
            if (e < b.d2 || (e == b.d2 && (uint32_t)id < (uint32_t)b.idx)) {
                b.d2 = e; b.idx = id; b.thr = fminf(b.thr, pc_thr_from(e));
            }
        }
    }
}

// Core traversal.  On entry b holds the initial bound (d2 = +inf, idx = -1, thr = bound).
__device__ __forceinline__ void pc_nearest_traverse(const pc_tree &T, float qx, float qy, float qz, pc_best &b)
{
    const double qxd = (double)qx, qyd = (double)qy, qzd = (double)qz;
    uint32_t stack_node[PC_STACK];
    float stack_d[PC_STACK];
    int sp = 0;
    uint32_t node = 1;
    for (;;) {
        // children of `node` are the aligned pair (2 node, 2 node + 1) = nodes[4 node .. 4 node + 3]
        const float4 *pair = T.nodes + 4ull * node;
        const float4 lo0 = __ldg(pair), hi0 = __ldg(pair + 1), lo1 = __ldg(pair + 2), hi1 = __ldg(pair + 3);
        const float d0 = pc_box_d2(lo0, hi0, qx, qy, qz);
        const float d1 = pc_box_d2(lo1, hi1, qx, qy, qz);
        const uint32_t c0 = 2u * node;
        const bool first0 = d0 <= d1;
        const uint32_t cn = first0 ? c0 : c0 + 1, cf = first0 ? c0 + 1 : c0;
        const float dn = fminf(d0, d1), df = fmaxf(d0, d1);
        bool descended = false;
        if (c0 >= T.P) {
            // children are leaves
            if (dn <= b.thr) pc_scan_leaf(T.points + (size_t)(cn - T.P) * PC_LEAF, qx, qy, qz, qxd, qyd, qzd, b);
            if (df <= b.thr) pc_scan_leaf(T.points + (size_t)(cf - T.P) * PC_LEAF, qx, qy, qz, qxd, qyd, qzd, b);
        } else {
            if (df <= b.thr) { stack_node[sp] = cf; stack_d[sp] = df; sp++; }
            if (dn <= b.thr) { node = cn; descended = true; }
        }
        if (descended) continue;
        // pop until a still-promising node is found
        bool found = false;
        while (sp > 0) {
            sp--;
            if (stack_d[sp] <= b.thr) { node = stack_node[sp]; found = true; break; }
        }
        if (!found) break;
    }
}

// ---- kernels ----------------------------------------------------------------------------------------
// perm (nullable): process query perm[t] in slot t (Morton-ordered batches); results go to the original slot.
__global__ void __launch_bounds__(PC_QUERY_THREADS)
pc_nearest_kernel(pc_tree T, const float *__restrict__ q, int64_t m, int qstride, const uint32_t *__restrict__ perm,
                  int32_t *__restrict__ out_idx, float *__restrict__ out_d2)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m) return;
    int64_t k = perm ? (int64_t)perm[t] : t;
    const float *qq = q + k * qstride;
    float qx = qq[0], qy = qq[1], qz = qq[2];
    pc_best b; b.d2 = INFINITY; b.idx = -1; b.thr = FLT_MAX;
    if (T.n_points > 0) pc_nearest_traverse(T, qx, qy, qz, b);
    if (out_idx) out_idx[k] = b.idx;
    if (out_d2) out_d2[k] = (float)b.d2;
}

struct pc_radius_dev {
    double search_margin, max_radius, sample_range;
    double sx, sy, sz;
    float bound_thr;    // fp32 threshold on d2 for the bounded search (FLT_MAX when unbounded)
    int bounded;        // PC_RADIUS_BOUNDED: out_idx = -1 wherever the radius clamps to max_radius
};

// radiusSearch epilogue on a finished search (corridor_finder.cpp:131-132); nothing found inside the bound => clamp
__device__ __forceinline__ double pc_radius_epilogue(const pc_best &b, const pc_radius_dev &R)
{
    if (b.idx < 0) return R.max_radius;
    double radius = __dsub_rn(__dsqrt_rn(b.d2), R.search_margin);
    return radius < R.max_radius ? radius : R.max_radius;
}

// radiusSearch early-out (corridor_finder.cpp:115-116): |p - start| > sample_range + max_radius
__device__ __forceinline__ bool pc_radius_early_out(double px, double py, double pz, const pc_radius_dev &R)
{
    if (R.sample_range < 0.0) return false;
    double dx = __dsub_rn(px, R.sx), dy = __dsub_rn(py, R.sy), dz = __dsub_rn(pz, R.sz);
    double s = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    return __dsqrt_rn(s) > __dadd_rn(R.sample_range, R.max_radius);
}

__global__ void __launch_bounds__(PC_QUERY_THREADS)
pc_radius_kernel(pc_tree T, pc_radius_dev R, const float *__restrict__ q, int64_t m, int qstride,
                 const uint32_t *__restrict__ perm, float *__restrict__ out_radius, int32_t *__restrict__ out_idx)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m) return;
    int64_t k = perm ? (int64_t)perm[t] : t;
    const float *qq = q + k * qstride;
    float qx = qq[0], qy = qq[1], qz = qq[2];
    double radius;
    int32_t idx = -1;
    if (T.n_points == 0 || pc_radius_early_out((double)qx, (double)qy, (double)qz, R)) {
        radius = __dsub_rn(R.max_radius, R.search_margin);
    } else {
        pc_best b; b.d2 = INFINITY; b.idx = -1; b.thr = R.bound_thr;
        pc_nearest_traverse(T, qx, qy, qz, b);
        radius = pc_radius_epilogue(b, R);
        idx = (R.bounded && !(radius < R.max_radius)) ? -1 : b.idx;
    }
    if (out_radius) out_radius[k] = (float)radius;
    if (out_idx) out_idx[k] = idx;
}

// Morton keys of a query batch in the index's frame, keeping only the top bits (coarse cells are enough to make
// the lanes of a warp walk the same part of the tree)
__global__ void __launch_bounds__(256)
pc_query_key_kernel(const float *__restrict__ q, int64_t m, int qstride, const uint32_t *__restrict__ bbox, int drop_bits,
                    uint32_t *__restrict__ keys, uint32_t *__restrict__ vals)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const pc_frame f = pc_make_frame(bbox, 10);
    const float *p = q + i * qstride;
    keys[i] = pc_morton30(p[0], p[1], p[2], f) >> drop_bits;
    vals[i] = (uint32_t)i;
}
