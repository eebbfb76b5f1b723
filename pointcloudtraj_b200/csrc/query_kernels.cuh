// query_kernels.cuh -- exact nearest / radius queries: near-first descent of the bounding-box tree with a short
// per-thread stack, one query per lane.
//
// Replaces kd_nearest_i / kd_nearest3 (Utils/kdtree/src/kdtree.c:345-457,493-500) and the epilogue of
// safeRegionRrtStar::radiusSearch (Planner/src/corridor_finder.cpp:113-133).
//
// Exactness: boxes and points are filtered in fp32 against `thr`, an upper bound of the current best fp64
// distance inflated by 2^-20 (fp32 evaluation error of d2 is < 2^-22 relative, so no candidate whose fp64
// distance is <= best can be filtered out); survivors are re-evaluated in fp64 with un-fused
// __dsub_rn/__dmul_rn/__dadd_rn in the reference's operation order and ranked by (d2, original index).
#pragma once
#include "common.cuh"
#include "build_kernels.cuh"

#define PC_QUERY_THREADS 128
#ifndef PC_QUERY_CURVE
#define PC_QUERY_CURVE 1     // batch ordering curve: 0 = Morton, 1 = Hilbert
#endif
#ifndef PC_PREFETCH_PUSH
#define PC_PREFETCH_PUSH 0
#endif
#ifndef PC_PACKET_ORDER
#define PC_PACKET_ORDER 0
#endif
#define PC_STACK 32
#define PC_THR_SLACK 1.00000095367431640625f   // 1 + 2^-20

struct pc_tree {
    const float4 *__restrict__ nodes;    // boxes: node i -> nodes[2i] (min), nodes[2i+1] (max)
    const float4 *__restrict__ points;   // leaf j -> points[PC_LEAF * j ..]
    int64_t n_points;
    uint32_t P;            // leaf base (power of two >= 2)
    // experimental second tree over the same points (lbvh_kernels.cuh; null unless the index was created with PC_LBVH=1)
    const float4 *__restrict__ lbvh;
    uint32_t lbvh_root;
};

// one box (min, max: 32 bytes, 32-byte aligned) with ONE 256-bit read-only load (sm_100: LDG.E.256)
__device__ __forceinline__ void pc_load_box(const float4 *__restrict__ box, float4 &lo, float4 &hi)
{
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(lo.x), "=f"(lo.y), "=f"(lo.z), "=f"(lo.w), "=f"(hi.x), "=f"(hi.y), "=f"(hi.z), "=f"(hi.w)
        : "l"(box));
}

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

__device__ __forceinline__ void pc_consider(const float4 p, float d, float qx, float qy, float qz, pc_best &b)
{
    if (d <= b.thr) {
        const double e = pc_exact_d2(p.x, p.y, p.z, (double)qx, (double)qy, (double)qz);
        const int32_t id = __float_as_int(p.w);
        if (e < b.d2 || (e == b.d2 && (uint32_t)id < (uint32_t)b.idx)) {
            b.d2 = e; b.idx = id; b.thr = fminf(b.thr, pc_thr_from(e));
        }
    }
}

__device__ __forceinline__ void pc_scan_leaf(const float4 *__restrict__ pts, float qx, float qy, float qz, pc_best &b)
{
    float4 p[PC_LEAF];
    float d[PC_LEAF];
#pragma unroll
    for (int i = 0; i < PC_LEAF; i += 2) pc_load_box(pts + i, p[i], p[i + 1]);     // two points per 256-bit load
    float dmin = FLT_MAX;
#pragma unroll
    for (int i = 0; i < PC_LEAF; i++) {
        const float dx = p[i].x - qx, dy = p[i].y - qy, dz = p[i].z - qz;
        d[i] = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        dmin = fminf(dmin, d[i]);
    }
    if (dmin <= b.thr) {          // one branch for the whole leaf: after the first leaves almost never taken
#pragma unroll
        for (int i = 0; i < PC_LEAF; i++) pc_consider(p[i], d[i], qx, qy, qz, b);
    }
}

// Core traversal of one query by one thread.  On entry b holds the initial bound (d2 = +inf, idx = -1, thr = bound).
__device__ __forceinline__ void pc_nearest_traverse(const pc_tree &T, float qx, float qy, float qz, pc_best &b)
{
    uint32_t stack_node[PC_STACK];
    float stack_d[PC_STACK];
    int sp = 0;
    uint32_t node = 1;
    for (;;) {
        // children of `node` are the aligned pair (2 node, 2 node + 1) = nodes[4 node .. 4 node + 3]
        const float4 *pair = T.nodes + 4ull * node;
        float4 lo0, hi0, lo1, hi1;
        pc_load_box(pair, lo0, hi0);
        pc_load_box(pair + 2, lo1, hi1);
        const float d0 = pc_box_d2(lo0, hi0, qx, qy, qz);
        const float d1 = pc_box_d2(lo1, hi1, qx, qy, qz);
        const uint32_t c0 = 2u * node;
        const bool first0 = d0 <= d1;
        const uint32_t cn = first0 ? c0 : c0 + 1, cf = first0 ? c0 + 1 : c0;
        const float dn = fminf(d0, d1), df = fmaxf(d0, d1);
        bool descended = false;
        if (c0 >= T.P) {
            // children are leaves
            if (dn <= b.thr) pc_scan_leaf(T.points + (size_t)(cn - T.P) * PC_LEAF, qx, qy, qz, b);
            if (df <= b.thr) pc_scan_leaf(T.points + (size_t)(cf - T.P) * PC_LEAF, qx, qy, qz, b);
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

// ---- radiusSearch pieces ------------------------------------------------------------------------------
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

#define PC_KIND_NEAREST 0
#define PC_KIND_RADIUS 1

// result of a query that needs no search: empty cloud (kd_nearest3 -> NULL; corridor_finder.cpp:118-120) or outside the
// sensing range (corridor_finder.cpp:115-116)
template <int KIND>
__device__ __forceinline__ void pc_write_trivial(const pc_radius_dev &R, uint32_t k, int32_t *out_idx, float *out_f)
{
    if (out_idx) out_idx[k] = -1;
    if (out_f) out_f[k] = (KIND == PC_KIND_RADIUS) ? (float)__dsub_rn(R.max_radius, R.search_margin) : INFINITY;
}

template <int KIND>
__device__ __forceinline__ void pc_write_result(const pc_radius_dev &R, const pc_best &b, uint32_t k, int32_t *out_idx, float *out_f)
{
    if (KIND == PC_KIND_RADIUS) {
        const double radius = pc_radius_epilogue(b, R);
        if (out_f) out_f[k] = (float)radius;
        if (out_idx) out_idx[k] = (R.bounded && !(radius < R.max_radius)) ? -1 : b.idx;
    } else {
        if (out_idx) out_idx[k] = b.idx;
        if (out_f) out_f[k] = (float)b.d2;
    }
}

// ---- variant 1: one thread per query, exit when done ----------------------------------------------------
// perm (nullable): process query perm[t] in slot t (Morton-ordered batches); results go to the original slot.
// m_eff (nullable): device-side number of leading entries of perm that need a search (the rest were answered by the
// ordering pass).
template <int KIND>
__global__ void __launch_bounds__(PC_QUERY_THREADS)
pc_query_simple_kernel(pc_tree T, pc_radius_dev R, const float *__restrict__ q, int64_t m, int qstride,
                       const uint32_t *__restrict__ perm, const unsigned long long *__restrict__ m_eff,
                       int32_t *__restrict__ out_idx, float *__restrict__ out_f)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m || (m_eff && t >= (int64_t)*m_eff)) return;
    const uint32_t k = perm ? perm[t] : (uint32_t)t;
    const float *qq = q + (size_t)k * qstride;
    const float qx = qq[0], qy = qq[1], qz = qq[2];
    bool search = T.n_points > 0;
    if (KIND == PC_KIND_RADIUS && search && !m_eff && pc_radius_early_out((double)qx, (double)qy, (double)qz, R)) search = false;
    if (!search) { pc_write_trivial<KIND>(R, k, out_idx, out_f); return; }
    pc_best b; b.d2 = INFINITY; b.idx = -1; b.thr = (KIND == PC_KIND_RADIUS) ? R.bound_thr : FLT_MAX;
    pc_nearest_traverse(T, qx, qy, qz, b);
    pc_write_result<KIND>(R, b, k, out_idx, out_f);
}

// ---- variant 2: persistent warps, lanes refilled as their query finishes -----------------------------------
// A query walks a data-dependent number of tree nodes, so "one thread per query, exit when done" leaves most lanes
// of a warp idle while the longest query of the 32 finishes (measured: 5.7 of 32 lanes active, profiles/r1_full_v1*).
// Here every warp owns a running window of the (Morton-ordered) batch, taken PC_Q_CHUNK queries at a time from a
// global counter and staged in shared memory; finished lanes park their result and, once `min_idle` lanes are
// parked, write the results and take the next queries of the window.  Every loop iteration performs ONE traversal
// step per lane: load the record of the node to visit (a pair of child boxes, or the points of a leaf) and process it.
#define PC_Q_CHUNK 128

template <int KIND>
__global__ void __launch_bounds__(PC_QUERY_THREADS, 8)
pc_query_persist_kernel(pc_tree T, pc_radius_dev R, const float *__restrict__ q, int64_t m, int qstride,
                        const uint32_t *__restrict__ perm, const unsigned long long *__restrict__ m_eff,
                        int32_t *__restrict__ out_idx, float *__restrict__ out_f,
                        unsigned long long *__restrict__ counter, int min_idle)
{
    __shared__ float4 s_q[PC_QUERY_THREADS / 32][PC_Q_CHUNK];   // staged window: x, y, z, caller slot (int bits)
    const uint32_t lt = pc_lanemask_lt();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long m_search = m_eff ? (long long)*m_eff : (long long)m;
    uint32_t stack_node[PC_STACK];
    float stack_d[PC_STACK];
    int sp = 0;
    // warp-uniform window [cur, end) of the batch; s_q[warp][j] holds query win0 + j
    long long cur = 0, end = 0, win0 = 0;
    bool exhausted = false;
    // per-lane query state
    bool has_q = false, parked = false;   // parked: finished, result not yet written
    uint32_t node = 0;          // node to visit next; 0 = take the next one from the stack
    uint32_t k = 0;             // slot of the query in the caller's arrays
    float qx = 0.f, qy = 0.f, qz = 0.f;
    pc_best b; b.d2 = INFINITY; b.idx = -1; b.thr = FLT_MAX;

    for (;;) {
        uint32_t idle = __ballot_sync(PC_FULL_MASK, !has_q);
        if (__popc(idle) >= min_idle) {
            // parked lanes write their results (the epilogue is deferred to here so that it runs for many lanes at once)
            if (parked) { pc_write_result<KIND>(R, b, k, out_idx, out_f); parked = false; }
            // hand queries to the idle lanes; queries that need no search are answered on the spot, so keep going
            // until the idle lanes hold real work or the batch is used up
            for (;;) {
                if (cur == end && !exhausted) {
                    unsigned long long base = 0;
                    if (lane == 0) base = atomicAdd(counter, (unsigned long long)PC_Q_CHUNK);
                    base = __shfl_sync(PC_FULL_MASK, base, 0);
                    if (base >= (unsigned long long)m_search) {
                        exhausted = true;
                    } else {
                        win0 = cur = (long long)base;
                        end = cur + PC_Q_CHUNK < m_search ? cur + PC_Q_CHUNK : m_search;
                        __syncwarp();
                        for (int j = lane; j < (int)(end - cur); j += 32) {
                            const long long t = cur + j;
                            const uint32_t kk = perm ? perm[t] : (uint32_t)t;
                            const float *qq = q + (size_t)kk * qstride;
                            s_q[warp][j] = make_float4(qq[0], qq[1], qq[2], __uint_as_float(kk));
                        }
                        __syncwarp();
                    }
                }
                const int avail = (int)(end - cur);
                const int want = __popc(idle);
                const int rank = __popc(idle & lt);
                if (!has_q && rank < avail) {
                    const float4 v = s_q[warp][(int)(cur - win0) + rank];
                    qx = v.x; qy = v.y; qz = v.z; k = __float_as_uint(v.w);
                    bool search = T.n_points > 0;
                    if (KIND == PC_KIND_RADIUS && search && !m_eff && pc_radius_early_out((double)qx, (double)qy, (double)qz, R)) search = false;
                    if (search) {
                        has_q = true; node = 1; sp = 0;
                        b.d2 = INFINITY; b.idx = -1; b.thr = (KIND == PC_KIND_RADIUS) ? R.bound_thr : FLT_MAX;
                    } else {
                        pc_write_trivial<KIND>(R, k, out_idx, out_f);
                    }
                }
                cur += want < avail ? want : avail;
                idle = __ballot_sync(PC_FULL_MASK, !has_q);
                if (exhausted || __popc(idle) < min_idle) break;
            }
            if (exhausted && idle == PC_FULL_MASK) break;
        }
        // ---- one traversal step --------------------------------------------------------------------------
        if (has_q) {
            if (node >= T.P) {
                pc_scan_leaf(T.points + (size_t)(node - T.P) * PC_LEAF, qx, qy, qz, b);
                node = 0;
            } else {
                const float4 *pair = T.nodes + 4ull * node;
                float4 lo0, hi0, lo1, hi1;
                pc_load_box(pair, lo0, hi0);
                pc_load_box(pair + 2, lo1, hi1);
                const float d0 = pc_box_d2(lo0, hi0, qx, qy, qz);
                const float d1 = pc_box_d2(lo1, hi1, qx, qy, qz);
                const uint32_t c0 = 2u * node;
                const bool first0 = d0 <= d1;
                const uint32_t cn = first0 ? c0 : c0 + 1, cf = first0 ? c0 + 1 : c0;
                const float dn = fminf(d0, d1), df = fmaxf(d0, d1);
                if (df <= b.thr) { stack_node[sp] = cf; stack_d[sp] = df; sp++; }
                node = dn <= b.thr ? cn : 0;
            }
            if (node == 0) {
                while (sp > 0 && stack_d[sp - 1] > b.thr) sp--;          // drop entries the shrinking bound has pruned
                if (sp == 0) { has_q = false; parked = true; }
                else node = stack_node[--sp];
            }
        }
    }
}

// ---- variant 3: warp packets ---------------------------------------------------------------------------------
// After the ordering pass the 32 queries of a warp lie in one small Morton cell, so their searches visit almost the
// same nodes.  The warp therefore walks the tree ONCE for all 32 queries: one shared stack (kept in registers, entry i
// in lane i, read back with a shuffle), every node record loaded once at a warp-uniform address, each lane testing its
// own query against it; a subtree is entered when ANY lane still needs it (ballot), the nearer child is chosen by
// majority vote.  Control flow is warp-uniform, so all 32 lanes are active at every step and there is no per-lane
// stack in local memory; the price is that a lane also visits nodes only its neighbours needed.
// Tried and dropped (profiles/r1_sweep3*, r1_sweep4*): rejecting stale stack entries with a per-entry minimum
// distance and __reduce_min/max_sync (3 % slower: entries are rarely stale), and a "wide" variant that tests the 32
// descendants five levels down against the packet's bounding box, one per lane (40 % slower: one lane whose
// search radius stays at the bound keeps the whole packet's bound large, so far too many leaves survive the
// conservative test and need a per-query re-test).
// The packet walk itself: must be called by all 32 lanes of a warp, converged; lanes without a query pass
// b.thr < 0 and take part in the votes only.
__device__ __forceinline__ void pc_packet_traverse(const pc_tree &T, float qx, float qy, float qz, pc_best &b, int lane)
{
    if (__ballot_sync(PC_FULL_MASK, b.thr >= 0.f) == 0) return;
    uint32_t my_entry = 0;      // warp stack: entry i lives in lane i
    int sp = 0;
    uint32_t node = 1;
    for (;;) {
        const float4 *pair = T.nodes + 4ull * node;
        float4 lo0, hi0, lo1, hi1;
        pc_load_box(pair, lo0, hi0);
        pc_load_box(pair + 2, lo1, hi1);
        const float d0 = pc_box_d2(lo0, hi0, qx, qy, qz);
        const float d1 = pc_box_d2(lo1, hi1, qx, qy, qz);
        const uint32_t w0 = __ballot_sync(PC_FULL_MASK, d0 <= b.thr);
        const uint32_t w1 = __ballot_sync(PC_FULL_MASK, d1 <= b.thr);
        const uint32_t c0 = 2u * node;
        bool pop = true;
        if (w0 | w1) {
#if PC_PACKET_ORDER == 0
            // the child most interested lanes are nearer to goes first (no vote needed when only one child is wanted)
            bool first0 = w1 == 0;
            if (w0 != 0 && w1 != 0) {
                const uint32_t pref0 = __ballot_sync(PC_FULL_MASK, d0 <= d1) & (w0 | w1);
                first0 = 2 * __popc(pref0) >= __popc(w0 | w1);
            }
#else
            // the child more lanes still need goes first: cheaper, but measured 14 % slower on radius batches and 11x slower on
            // unbounded nearest batches (no near-first order while every lane still wants both children); kept for the record
            const bool first0 = __popc(w0) >= __popc(w1);
#endif
            const uint32_t cn = c0 + (first0 ? 0u : 1u), cf = cn ^ 1u;
            const bool both = w0 != 0 && w1 != 0;
            if (c0 >= T.P) {
                pc_scan_leaf(T.points + (size_t)(cn - T.P) * PC_LEAF, qx, qy, qz, b);
                if (both && __ballot_sync(PC_FULL_MASK, (first0 ? d1 : d0) <= b.thr))
                    pc_scan_leaf(T.points + (size_t)(cf - T.P) * PC_LEAF, qx, qy, qz, b);
            } else {
                if (both) {
                    if (lane == sp) my_entry = cf;
                    sp++;
#if PC_PREFETCH_PUSH
                    // the far child will be visited later: start pulling its record towards L2 now (matters when the
                    // index is larger than L2 and a visit would otherwise wait for HBM)
                    if (lane == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(T.nodes + 4ull * cf));
#endif
                }
                node = cn;
                pop = false;
            }
        }
        if (pop) {
            if (sp == 0) break;
            sp--;
            node = __shfl_sync(PC_FULL_MASK, my_entry, sp);
        }
    }
}

template <int KIND>
__global__ void __launch_bounds__(PC_QUERY_THREADS)
pc_query_packet_kernel(pc_tree T, pc_radius_dev R, const float *__restrict__ q, int64_t m, int qstride,
                       const uint32_t *__restrict__ perm, const unsigned long long *__restrict__ m_eff,
                       int32_t *__restrict__ out_idx, float *__restrict__ out_f)
{
    const int lane = threadIdx.x & 31;
    const long long m_search = m_eff ? (long long)*m_eff : (long long)m;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t - lane >= m_search) return;                       // whole warp past the end
    bool valid = t < m_search;
    uint32_t k = 0;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    pc_best b; b.d2 = INFINITY; b.idx = -1; b.thr = -1.0f;  // thr < 0: this lane needs nothing
    if (valid) {
        k = perm ? perm[t] : (uint32_t)t;
        const float *qq = q + (size_t)k * qstride;
        qx = qq[0]; qy = qq[1]; qz = qq[2];
        bool search = T.n_points > 0;
        if (KIND == PC_KIND_RADIUS && search && !m_eff && pc_radius_early_out((double)qx, (double)qy, (double)qz, R)) search = false;
        if (search) b.thr = (KIND == PC_KIND_RADIUS) ? R.bound_thr : FLT_MAX;
        else { pc_write_trivial<KIND>(R, k, out_idx, out_f); valid = false; }
    }
    pc_packet_traverse(T, qx, qy, qz, b, lane);
    if (valid) pc_write_result<KIND>(R, b, k, out_idx, out_f);
}

// ---- variant 4: packets of 64 queries, two per lane ---------------------------------------------------------------
// The packet walk is bound by instruction issue and by L1 register fill (every lane receives the full 64-byte record
// of every visited node).  Giving each lane TWO queries -- slots t and t + 32 of the warp's 64 consecutive ordered
// queries -- halves the record bytes and the control instructions per query-visit; the price is a slightly larger packet.
// (Rejecting stale stack entries with a per-entry minimum distance and __reduce_min/max_sync was measured again on this
// kernel: 12 % of the visits are stale, yet the extra warp reductions cost more than the visits saved -- 2.19 vs 2.03 ms.
// So was a 4-ary walk of the same tree -- the grandchildren 4i..4i+3 of node i are one aligned 128-byte line, so a visit can
// test four boxes and skip a level, children ordered by a packed-vote warp reduction: identical results, 2.06 vs 2.04 ms on
// radius batches and 8.46 vs 7.84 ms on unbounded nearest batches (profiles/r1_sweep6*): the box tests it wastes on
// grandchildren whose parent would have been pruned cost what the saved votes and stack traffic gain.)
__device__ __forceinline__ void pc_scan_leaf2(const float4 *__restrict__ pts, const float qa[3], const float qb[3], pc_best &ba, pc_best &bb)
{
    float4 p[PC_LEAF];
    float da[PC_LEAF], db[PC_LEAF];
#pragma unroll
    for (int i = 0; i < PC_LEAF; i += 2) pc_load_box(pts + i, p[i], p[i + 1]);     // two points per 256-bit load
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

#ifdef PC_STATS
__device__ unsigned long long pc_stats_hist[65];
#endif
__device__ __forceinline__ void pc_packet2_traverse(const pc_tree &T, const float qa[3], const float qb[3], pc_best &ba, pc_best &bb, int lane)
{
    if (__ballot_sync(PC_FULL_MASK, ba.thr >= 0.f || bb.thr >= 0.f) == 0) return;
    uint32_t my_entry = 0;
    int sp = 0;
    uint32_t node = 1;
    for (;;) {
        const float4 *pair = T.nodes + 4ull * node;
        float4 lo0, hi0, lo1, hi1;
        pc_load_box(pair, lo0, hi0);
        pc_load_box(pair + 2, lo1, hi1);
        const float a0 = pc_box_d2(lo0, hi0, qa[0], qa[1], qa[2]), a1 = pc_box_d2(lo1, hi1, qa[0], qa[1], qa[2]);
        const float b0 = pc_box_d2(lo0, hi0, qb[0], qb[1], qb[2]), b1 = pc_box_d2(lo1, hi1, qb[0], qb[1], qb[2]);
        const bool wa0 = a0 <= ba.thr, wa1 = a1 <= ba.thr, wb0 = b0 <= bb.thr, wb1 = b1 <= bb.thr;
        const uint32_t w0 = __ballot_sync(PC_FULL_MASK, wa0 || wb0);
        const uint32_t w1 = __ballot_sync(PC_FULL_MASK, wa1 || wb1);
        const uint32_t c0 = 2u * node;
#ifdef PC_STATS
        {   // histogram of how many of the 64 queries wanted this node's children (diagnostic build only)
            const int na = __popc(__ballot_sync(PC_FULL_MASK, wa0 || wa1)) + __popc(__ballot_sync(PC_FULL_MASK, wb0 || wb1));
            if (lane == 0) atomicAdd(&pc_stats_hist[na], 1ull);
        }
#endif
        bool pop = true;
        if (w0 | w1) {
            const bool both = w0 != 0 && w1 != 0;
            bool first0 = w1 == 0;
            if (both) {
                // majority vote of the lanes' FIRST queries that are interested (not needed when only one child is wanted).
                // Measured (profiles/r1_sweep8_vote_variants.txt): letting both queries of a lane vote costs two more ballots
                // and orders no better (1.955 vs 1.841 ms); a packed __reduce_add_sync vote is no faster than ballots (1.954);
                // letting the first interested lane decide alone is cheaper still but orders worse (1.949).
                const uint32_t ia = __ballot_sync(PC_FULL_MASK, wa0 || wa1);
                const uint32_t pa = __ballot_sync(PC_FULL_MASK, a0 <= a1) & ia;
                first0 = 2 * __popc(pa) >= __popc(ia);
            }
            const uint32_t cn = c0 + (first0 ? 0u : 1u), cf = cn ^ 1u;
            if (c0 >= T.P) {
                pc_scan_leaf2(T.points + (size_t)(cn - T.P) * PC_LEAF, qa, qb, ba, bb);
                if (both && __ballot_sync(PC_FULL_MASK, (first0 ? a1 : a0) <= ba.thr || (first0 ? b1 : b0) <= bb.thr))
                    pc_scan_leaf2(T.points + (size_t)(cf - T.P) * PC_LEAF, qa, qb, ba, bb);
            } else {
                if (both) { if (lane == sp) my_entry = cf; sp++; }
                node = cn;
                pop = false;
            }
        }
        if (pop) {
            if (sp == 0) break;
            sp--;
            node = __shfl_sync(PC_FULL_MASK, my_entry, sp);
        }
    }
}

template <int KIND>
__global__ void __launch_bounds__(PC_QUERY_THREADS)
pc_query_packet2_kernel(pc_tree T, pc_radius_dev R, const float *__restrict__ q, int64_t m, int qstride,
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
    pc_packet2_traverse(T, qv[0], qv[1], b[0], b[1], lane);
#pragma unroll
    for (int j = 0; j < 2; j++)
        if (valid[j]) pc_write_result<KIND>(R, b[j], k[j], out_idx, out_f);
}

// ---- variant 6: a GROUP of lanes per query (small batches) ------------------------------------------------------------
// The planner's own loop asks for ONE radius per call (corridor_finder.cpp:404), and then a search is a chain of dependent
// loads: ~60-100 node visits of 0.5 us each with one thread per query.  Here G lanes (a whole warp for the smallest batches,
// 16 or 8 lanes when there are more queries than the GPU holds warps) work on the same query: the open nodes sit on a LIFO
// frontier in shared memory, every step the group takes the (up to) G most recently pushed -- deepest, nearest -- nodes,
// one per lane, tests their two child boxes (or scans their four points), shares the tightened bound with one warp
// reduction and pushes the surviving children, the nearer ones on top.  The number of dependent steps drops from the
// number of visits to roughly the depth of the tree.  The walk starts from the (real) nodes of level 5 / 4 / 3.
// Exactness as everywhere: fp32 filter against the shared bound, fp64 re-evaluation, and the lanes' private bests are
// merged at the end by (d2, index).  Groups of one warp run different numbers of steps; every warp-level primitive is
// called with the group's own lane mask.
#define PC_COOP_WARPS 4
#define PC_COOP_CAP(G) (32 * (G))          // frontier entries per group: 8 KB of shared memory per warp for every G

template <int KIND, int G>
__global__ void __launch_bounds__(32 * PC_COOP_WARPS)
pc_query_coop_kernel(pc_tree T, pc_radius_dev R, const float *__restrict__ q, int64_t m, int qstride,
                     int32_t *__restrict__ out_idx, float *__restrict__ out_f)
{
    constexpr int GROUPS = 32 / G, CAP = PC_COOP_CAP(G), SEED_LEVEL = G == 32 ? 5 : (G == 16 ? 4 : 3);
    __shared__ uint2 s_front[PC_COOP_WARPS * GROUPS][CAP];      // (node, float bits of its box distance)
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int gl = lane & (G - 1), grp = lane / G;              // lane inside the group, group inside the warp
    const uint32_t gmask = G == 32 ? PC_FULL_MASK : (((1u << G) - 1u) << (grp * G));
    const int64_t k = ((int64_t)blockIdx.x * PC_COOP_WARPS + w) * GROUPS + grp;
    if (k >= m) return;
    const float *qq = q + (size_t)k * qstride;
    const float qx = qq[0], qy = qq[1], qz = qq[2];
    bool search = T.n_points > 0;
    if (KIND == PC_KIND_RADIUS && search && pc_radius_early_out((double)qx, (double)qy, (double)qz, R)) search = false;
    if (!search) { if (gl == 0) pc_write_trivial<KIND>(R, (uint32_t)k, out_idx, out_f); return; }
    pc_best b; b.d2 = INFINITY; b.idx = -1; b.thr = (KIND == PC_KIND_RADIUS) ? R.bound_thr : FLT_MAX;
    uint2 *F = s_front[w * GROUPS + grp];
    int size = 1;
    if (T.P >= 2u * G) {
        // the seed level holds G nodes (ids G .. 2G-1), each covering P / G leaves: only the real ones (the pads behind them
        // are not all written)
        const int64_t per = (int64_t)(T.P >> SEED_LEVEL), n_leaves = (T.n_points + PC_LEAF - 1) / PC_LEAF;
        size = (int)((n_leaves + per - 1) / per);
        if (gl < size) F[gl] = make_uint2((uint32_t)G + (uint32_t)gl, 0u);
    } else if (gl == 0) F[0] = make_uint2(1u, 0u);
    __syncwarp(gmask);
    const uint32_t lt = (1u << gl) - 1u;
    while (size > 0) {
        // take the top of the frontier, one node per lane (grows it by at most G entries); within G + 32 entries of the
        // capacity fall back to one node per step -- a plain DFS, which adds at most one entry per tree level (< 32)
        const int take = (size + G + 32 <= CAP) ? min(size, G) : 1;
        bool active = gl < take;
        uint2 e = make_uint2(0u, 0u);
        if (active) e = F[size - 1 - gl];
        size -= take;
        __syncwarp(gmask);
        active = active && __uint_as_float(e.y) <= b.thr;
        float dn = INFINITY, df = INFINITY;
        uint32_t cn = 0, cf = 0;
        if (active) {
            if (e.x >= T.P) {
                pc_scan_leaf(T.points + (size_t)(e.x - T.P) * PC_LEAF, qx, qy, qz, b);
            } else {
                const float4 *pair = T.nodes + 4ull * e.x;
                float4 lo0, hi0, lo1, hi1;
                pc_load_box(pair, lo0, hi0);
                pc_load_box(pair + 2, lo1, hi1);
                const float d0 = pc_box_d2(lo0, hi0, qx, qy, qz), d1 = pc_box_d2(lo1, hi1, qx, qy, qz);
                const bool first0 = d0 <= d1;
                cn = 2u * e.x + (first0 ? 0u : 1u); cf = cn ^ 1u;
                dn = fminf(d0, d1); df = fmaxf(d0, d1);
            }
        }
        // one bound for the whole group (thr >= 0, so the float order is the order of its bits)
        b.thr = __uint_as_float(__reduce_min_sync(gmask, __float_as_uint(b.thr)));
        const bool wn = dn <= b.thr, wf = df <= b.thr;
        const uint32_t mf = (__ballot_sync(gmask, wf) & gmask) >> (grp * G), mn = (__ballot_sync(gmask, wn) & gmask) >> (grp * G);
        const int nf = __popc(mf), nn = __popc(mn);
        if (wf) F[size + __popc(mf & lt)] = make_uint2(cf, __float_as_uint(df));
        if (wn) F[size + nf + (nn - 1 - __popc(mn & lt))] = make_uint2(cn, __float_as_uint(dn));   // lane 0's child ends on top
        size += nf + nn;
        __syncwarp(gmask);
    }
    // merge the lanes' private bests: smallest (d2, index); d2 >= 0, so its bit pattern orders like the value
    long long key = __double_as_longlong(b.d2);
    uint32_t id = (uint32_t)b.idx;
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
        const long long k2 = __shfl_xor_sync(gmask, key, o);
        const uint32_t i2 = __shfl_xor_sync(gmask, id, o);
        if (k2 < key || (k2 == key && i2 < id)) { key = k2; id = i2; }
    }
    if (gl == 0) {
        b.d2 = __longlong_as_double(key); b.idx = (int32_t)id;
        pc_write_result<KIND>(R, b, (uint32_t)k, out_idx, out_f);
    }
}

// ---- ordering pass of a batch --------------------------------------------------------------------------
// Morton key of every query in the index's frame (top `30 - drop_bits` bits), so that the lanes of a warp walk the
// same part of the tree.  For radius batches the sensing-range early-out (corridor_finder.cpp:115-116) is evaluated
// here, once, coalesced: such queries get their result now; the others are compacted into (key, slot) pairs and
// n_search counts them -- only those are sorted and searched.
template <int KIND>
__global__ void __launch_bounds__(256)
pc_query_key_kernel(const float *__restrict__ q, int64_t m, int qstride, const uint32_t *__restrict__ bbox, int drop_bits,
                    pc_radius_dev R, int32_t *__restrict__ out_idx, float *__restrict__ out_f,
                    uint32_t *__restrict__ keys, uint32_t *__restrict__ vals, unsigned long long *__restrict__ n_search,
                    int shard_rank, int shard_n, int shard_shift)
{
    __shared__ uint32_t s_warp[8];
    __shared__ unsigned long long s_base;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    bool search = false;
    uint32_t key = 0;
    if (i < m) {
        const pc_frame f = pc_make_frame(bbox, 10);
        const float *p = q + i * qstride;
        const float x = p[0], y = p[1], z = p[2];
        search = true;
        if (KIND == PC_KIND_RADIUS && pc_radius_early_out((double)x, (double)y, (double)z, R)) {
            search = false;
            pc_write_trivial<KIND>(R, (uint32_t)i, out_idx, out_f);
        }
#if PC_QUERY_CURVE == 1
        key = pc_hilbert30(x, y, z, f);
#else
        key = pc_morton30(x, y, z, f);
#endif
        // pc_batch_shard: the cells of the curve at a batch-dependent level are dealt to the ranks by a hash of the cell
        // index.  A rank's share is then as dense in space as the whole batch (dense packets) and spread over the whole map
        // (equal cost per rank).  Measured on C5 with 8 GPUs: array slices 37 ms; contiguous stretches of the curve 41 ms
        // and round-robin 32^3 cells 43 ms (both unbalanced: 18..42 ms per rank -- on a flat map the low bits of the cell
        // index encode the z layer); hashed cells: see profiles/r1_c5_strong_scaling.jsonl.
        if (shard_n > 1 && (int)((((key >> shard_shift) * 2654435761u) >> 15) % (uint32_t)shard_n) != shard_rank) search = false;
        key >>= drop_bits;
    }
    // compact the queries that still need a search: only those are sorted and searched.  One atomic per CTA; the slot a
    // query lands in depends on CTA scheduling, which changes the composition of packets but never a result.
    const uint32_t mask = __ballot_sync(PC_FULL_MASK, search);
    if (lane == 0) s_warp[warp] = __popc(mask);
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t tot = 0;
        for (int w = 0; w < 8; w++) { const uint32_t c = s_warp[w]; s_warp[w] = tot; tot += c; }
        s_base = tot ? atomicAdd(n_search, (unsigned long long)tot) : 0ull;
    }
    __syncthreads();
    if (search) {
        const unsigned long long pos = s_base + s_warp[warp] + __popc(mask & pc_lanemask_lt());
        keys[pos] = key;
        vals[pos] = (uint32_t)i;
    }
}
