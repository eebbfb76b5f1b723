// clearance_kernels.cuh -- trajectory clearance: Bezier sampling fused with the nearest-obstacle query and a
// per-trajectory reduction.  Flat path (schedule / eval / finish kernels, bottom of this file) or one warp per trajectory.
//
// Replaces checkSafeTrajectory (Planner/src/sim_planning_demo.cpp:729-781), getPosFromBezier (:715-727) and
// safeRegionRrtStar::checkTrajPtCol (Planner/src/corridor_finder.cpp:412-416).
//
// The reference's time walk accumulates `t += 0.02` / `t_accu += 0.02` in doubles, so sample times are the
// result of REPEATED addition (not k * dt).  To reproduce them bit for bit, lane 0 of the trajectory's warp runs the walk
// sequentially and publishes a schedule of (segment, t) pairs in shared memory, PC_CLR_CHUNK samples at a time;
// the warp then evaluates the samples of the chunk 32 at a time.
// Powers u^j: the reference calls libm's pow (sim_planning_demo.cpp:724), whose result is the correctly rounded power in
// all but ~2^-15 of the cases (glibc's pow carries ~68 bits internally).  Here u^j is built by repeated multiplication in
// double-double arithmetic (error-free products with fma, relative error < 2^-100 after 12 steps) and rounded to double
// ONCE, i.e. it IS the correctly rounded power; the position is then cast to float32 exactly where radiusSearch does
// (corridor_finder.cpp:122-125), which absorbs a last-bit difference of a power in all but ~2^-29 of those rare cases.
#pragma once
#include "query_kernels.cuh"

#define PC_CLR_THREADS 128
#define PC_CLR_CHUNK 128
#define PC_MAX_ORDER 12

__constant__ double pc_binom[PC_MAX_ORDER + 1][PC_MAX_ORDER + 1];

struct pc_traj_dev { int32_t first_seg, num_seg; double t_now; };

__device__ __forceinline__ void pc_power_table(double u, int n, double *p)
{
    p[0] = 1.0;
    if (n >= 1) p[1] = u;
    double hi = u, lo = 0.0;                          // u^(j-1) = hi + lo, |lo| <= ulp(hi) / 2
    for (int j = 2; j <= n; j++) {
        const double ph = __dmul_rn(hi, u);
        const double pe = __fma_rn(hi, u, -ph);       // hi * u = ph + pe exactly
        const double pl = __fma_rn(lo, u, pe);
        hi = __dadd_rn(ph, pl);                       // fast two-sum: |ph| >= |pl|
        lo = __dadd_rn(__dsub_rn(ph, hi), pl);
        p[j] = hi;                                    // = round(hi + lo): hi + lo is normalised
    }
}

// p = T * sum_j C(n,j) c_j u^j (1-u)^(n-j) per axis, term evaluated left to right, accumulated from 0
__device__ __forceinline__ void pc_bezier_pos(const double *__restrict__ c, int n, double u, double T, double out[3])
{
    double pu[PC_MAX_ORDER + 1], pv[PC_MAX_ORDER + 1];
    pc_power_table(u, n, pu);
    pc_power_table(__dsub_rn(1.0, u), n, pv);
    const int nc = n + 1;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        double acc = 0.0;
        for (int j = 0; j <= n; j++) {
            double term = __dmul_rn(__dmul_rn(__dmul_rn(pc_binom[n][j], c[a * nc + j]), pu[j]), pv[n - j]);
            acc = __dadd_rn(acc, term);
        }
        out[a] = __dmul_rn(acc, T);
    }
}

// One WARP per trajectory (four per CTA).  Lane 0 replays the reference's time walk for the next PC_CLR_CHUNK samples into
// the warp's slice of shared memory, then the warp evaluates them 32 at a time: Bezier position in fp64, float32 cast, and the
// packet walk for the 32 consecutive samples (a few centimetres apart -- ideal packets).  No CTA-wide barrier: while one warp's
// lane 0 walks (a chain of dependent fp64 additions), the other warps of the SM evaluate.  (Round 1 ran one CTA per trajectory
// with thread 0 walking while 127 threads waited at a barrier: issue slots 37 % busy, profiles/r2_full_pc_clearance_kernel.txt.)
__global__ void __launch_bounds__(PC_CLR_THREADS)
pc_clearance_kernel(pc_tree T, pc_radius_dev R, const pc_traj_dev *__restrict__ traj, int64_t n_traj,
                    const int32_t *__restrict__ seg_order, const double *__restrict__ seg_T,
                    const int64_t *__restrict__ seg_coef_off, const double *__restrict__ coef,
                    double dt, double horizon,
                    int32_t *__restrict__ out_first_hit, float *__restrict__ out_min_radius, int32_t *__restrict__ out_n_samples)
{
    __shared__ double s_t[PC_CLR_THREADS / 32][PC_CLR_CHUNK];
    __shared__ int32_t s_seg[PC_CLR_THREADS / 32][PC_CLR_CHUNK];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t tr = (int64_t)blockIdx.x * (PC_CLR_THREADS / 32) + w;
    if (tr >= n_traj) return;
    const pc_traj_dev tj = traj[tr];
    const int32_t seg0 = tj.first_seg, nseg = tj.num_seg;

    // walk state (lane 0 only), sim_planning_demo.cpp:735-749
    int i = 0;
    double tt = 0.0, t_accu = 0.0;
    if (lane == 0) {
        double t_s = tj.t_now > 0.0 ? tj.t_now : 0.0;
        for (i = 0; i < nseg; ++i) {
            if (t_s > seg_T[seg0 + i] && i + 1 < nseg) t_s = __dsub_rn(t_s, seg_T[seg0 + i]);
            else break;
        }
        tt = t_s;
    }
    double my_min = INFINITY;
    int my_first = 0x7fffffff;
    int64_t chunk_base = 0;
    for (;;) {
        int cnt = 0, done = 0;
        if (lane == 0) {
            while (i < nseg && cnt < PC_CLR_CHUNK) {
                const double Ti = seg_T[seg0 + i];
                if (!(tt < Ti)) { i++; tt = 0.0; continue; }
                t_accu = __dadd_rn(t_accu, dt);
                if (t_accu > horizon) { i++; tt = 0.0; continue; }
                s_t[w][cnt] = tt; s_seg[w][cnt] = seg0 + i; cnt++;
                tt = __dadd_rn(tt, dt);
            }
            done = (i >= nseg);
        }
        __syncwarp();
        cnt = __shfl_sync(PC_FULL_MASK, cnt, 0);
        done = __shfl_sync(PC_FULL_MASK, done, 0);
        for (int base = 0; base < cnt; base += 32) {
            const int k = base + lane;
            const bool have = k < cnt;
            double radius = INFINITY;
            pc_best b[1];
            b[0].d2 = INFINITY; b[0].idx = -1; b[0].thr = -1.0f;
            float qv[1][3] = { { 0.f, 0.f, 0.f } };
            bool search = false;
            if (have) {
                const int32_t sg = s_seg[w][k];
                const double Ti = seg_T[sg];
                double pos[3];
                pc_bezier_pos(coef + seg_coef_off[sg], seg_order[sg], __ddiv_rn(s_t[w][k], Ti), Ti, pos);
                if (T.n_points == 0 || pc_radius_early_out(pos[0], pos[1], pos[2], R)) {
                    radius = __dsub_rn(R.max_radius, R.search_margin);
                } else {
                    qv[0][0] = (float)pos[0]; qv[0][1] = (float)pos[1]; qv[0][2] = (float)pos[2];
                    search = qv[0][0] == qv[0][0] && qv[0][1] == qv[0][1] && qv[0][2] == qv[0][2];
                    if (search) b[0].thr = R.bound_thr;
                }
            }
            pc_packet_traverse<1>(T, qv, b, lane);
            if (have) {
                if (search || !(radius < INFINITY)) radius = pc_radius_epilogue(b[0], R);
                my_min = fmin(my_min, radius);
                if (radius < 0.0) my_first = min(my_first, (int)(chunk_base + k));
            }
        }
        chunk_base += cnt;
        __syncwarp();                       // the chunk's schedule is consumed before lane 0 overwrites it
        if (done) break;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        my_min = fmin(my_min, __shfl_xor_sync(PC_FULL_MASK, my_min, o));
        my_first = min(my_first, __shfl_xor_sync(PC_FULL_MASK, my_first, o));
    }
    if (lane == 0) {
        if (out_min_radius) out_min_radius[tr] = (float)my_min;
        if (out_first_hit) out_first_hit[tr] = my_first == 0x7fffffff ? -1 : my_first;
        if (out_n_samples) out_n_samples[tr] = (int32_t)chunk_base;
    }
}

// ---- the flat path: every sample of every trajectory is its own unit of work ------------------------------------------------
// One warp per trajectory serialises a trajectory's samples: ONE trajectory of 600 samples (what the planner checks per call,
// sim_planning_demo.cpp:729-781) took 19 dependent packet walks, 0.32 ms.  Three kernels instead:
//   pc_clearance_schedule_kernel   one THREAD per trajectory replays the time walk (the only sequential part: repeated fp64
//                                  additions) and writes the (t, segment) schedule, `stride` slots per trajectory
//   pc_clearance_eval_kernel       one WARP per 32 consecutive samples of a trajectory, any trajectory: Bezier position,
//                                  float32 cast, packet walk, per-packet minimum and first colliding sample
//   pc_clearance_finish_kernel     one thread per trajectory reduces its packets
// Same arithmetic per sample as pc_clearance_kernel, so the results are identical (tests run both).
__global__ void pc_clearance_schedule_kernel(const pc_traj_dev *__restrict__ traj, int64_t n_traj, const double *__restrict__ seg_T,
                                             double dt, double horizon, int64_t stride,
                                             double *__restrict__ sched_t, int32_t *__restrict__ sched_seg, int32_t *__restrict__ n_samples)
{
    const int64_t tr = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tr >= n_traj) return;
    const pc_traj_dev tj = traj[tr];
    const int32_t seg0 = tj.first_seg, nseg = tj.num_seg;
    // sim_planning_demo.cpp:735-749
    int i = 0;
    double t_s = tj.t_now > 0.0 ? tj.t_now : 0.0;
    for (i = 0; i < nseg; ++i) {
        if (t_s > seg_T[seg0 + i] && i + 1 < nseg) t_s = __dsub_rn(t_s, seg_T[seg0 + i]);
        else break;
    }
    double tt = t_s, t_accu = 0.0;
    int64_t cnt = 0;
    double *st = sched_t + tr * stride;
    int32_t *ss = sched_seg + tr * stride;
    while (i < nseg && cnt < stride) {
        const double Ti = seg_T[seg0 + i];
        if (!(tt < Ti)) { i++; tt = 0.0; continue; }
        t_accu = __dadd_rn(t_accu, dt);
        if (t_accu > horizon) { i++; tt = 0.0; continue; }
        st[cnt] = tt; ss[cnt] = seg0 + i; cnt++;
        tt = __dadd_rn(tt, dt);
    }
    n_samples[tr] = (int32_t)cnt;
}

__global__ void __launch_bounds__(PC_CLR_THREADS)
pc_clearance_eval_kernel(pc_tree T, pc_radius_dev R, int64_t n_traj, int64_t stride, int64_t packets_per_traj,
                         const int32_t *__restrict__ seg_order, const double *__restrict__ seg_T,
                         const int64_t *__restrict__ seg_coef_off, const double *__restrict__ coef,
                         const double *__restrict__ sched_t, const int32_t *__restrict__ sched_seg, const int32_t *__restrict__ n_samples,
                         double *__restrict__ packet_min, int32_t *__restrict__ packet_first)
{
    const int lane = threadIdx.x & 31;
    const int64_t pk = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (pk >= n_traj * packets_per_traj) return;
    const int64_t tr = pk / packets_per_traj, base = (pk - tr * packets_per_traj) * 32;
    const int64_t cnt = n_samples[tr];
    if (base >= cnt) return;                               // (the finish kernel reads only the packets that exist)
    const int64_t k = base + lane;
    const bool have = k < cnt;
    double radius = INFINITY;
    pc_best b[1];
    b[0].d2 = INFINITY; b[0].idx = -1; b[0].thr = -1.0f;
    float qv[1][3] = { { 0.f, 0.f, 0.f } };
    bool search = false;
    if (have) {
        const int32_t sg = sched_seg[tr * stride + k];
        const double Ti = seg_T[sg];
        double pos[3];
        pc_bezier_pos(coef + seg_coef_off[sg], seg_order[sg], __ddiv_rn(sched_t[tr * stride + k], Ti), Ti, pos);
        if (T.n_points == 0 || pc_radius_early_out(pos[0], pos[1], pos[2], R)) {
            radius = __dsub_rn(R.max_radius, R.search_margin);
        } else {
            qv[0][0] = (float)pos[0]; qv[0][1] = (float)pos[1]; qv[0][2] = (float)pos[2];
            search = qv[0][0] == qv[0][0] && qv[0][1] == qv[0][1] && qv[0][2] == qv[0][2];
            if (search) b[0].thr = R.bound_thr;
        }
    }
    pc_packet_traverse<1>(T, qv, b, lane);
    double my_min = INFINITY;
    int my_first = 0x7fffffff;
    if (have) {
        if (search || !(radius < INFINITY)) radius = pc_radius_epilogue(b[0], R);
        my_min = radius;
        if (radius < 0.0) my_first = (int)k;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        my_min = fmin(my_min, __shfl_xor_sync(PC_FULL_MASK, my_min, o));
        my_first = min(my_first, __shfl_xor_sync(PC_FULL_MASK, my_first, o));
    }
    if (lane == 0) { packet_min[pk] = my_min; packet_first[pk] = my_first; }
}

__global__ void pc_clearance_finish_kernel(int64_t n_traj, int64_t packets_per_traj, const int32_t *__restrict__ n_samples,
                                           const double *__restrict__ packet_min, const int32_t *__restrict__ packet_first,
                                           int32_t *__restrict__ out_first_hit, float *__restrict__ out_min_radius)
{
    const int64_t tr = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tr >= n_traj) return;
    const int64_t np = ((int64_t)n_samples[tr] + 31) / 32;
    double mn = INFINITY;
    int first = 0x7fffffff;
    for (int64_t p = 0; p < np; p++) {
        mn = fmin(mn, packet_min[tr * packets_per_traj + p]);
        first = min(first, packet_first[tr * packets_per_traj + p]);
    }
    if (out_min_radius) out_min_radius[tr] = (float)mn;
    if (out_first_hit) out_first_hit[tr] = first == 0x7fffffff ? -1 : first;
}
