/*
 * planner_oracle.c -- TEST INFRASTRUCTURE ONLY (CPU parity oracle, never shipped, never timed as product).
 *
 * CPU restatement of the planner-side semantics that sit on top of the nearest-obstacle query.
 * The planner itself cannot be compiled here (needs roscpp, PCL 1.10, Eigen, libmosek64), so
 * these functions restate it with <math.h> only:
 *
 *   po_radius_search         follows Planner/src/corridor_finder.cpp:113-133  (safeRegionRrtStar::radiusSearch)
 *                            with getDis at corridor_finder.cpp:109-111
 *   po_check_traj_pt_col     follows Planner/src/corridor_finder.cpp:412-416  (checkTrajPtCol)
 *   po_binomial              follows Planner/src/bezier_base.cpp:35-48,256-266 (int factorial quotient)
 *   po_bezier_pos            follows Planner/src/sim_planning_demo.cpp:715-727 (getPosFromBezier)
 *   po_check_safe_trajectory follows Planner/src/sim_planning_demo.cpp:729-781 (checkSafeTrajectory)
 *   po_gen_samples           follows Planner/src/corridor_finder.cpp:333-358  (genSample while inform_status is false) on
 *                            std::default_random_engine (minstd_rand0, eng(0) at :12) and libstdc++'s
 *                            uniform_real_distribution<double> (generate_canonical, bits/random.tcc: two draws per double)
 *   po_steer                 follows Planner/src/corridor_finder.cpp:385-402  (the centre genNewNode queries)
 *
 * The cloud query inside radiusSearch is PCL/FLANN float32 in the reference (un-vendored third
 * party, PCL 1.10 EXACT, Planner/CMakeLists.txt:18; no reference test pins its outputs).  Per
 * BASELINE.json's north_star the parity target is Utils/kdtree in double precision instead, so the
 * nearest query here is kdo_nearest3 (kd_oracle.c) on the float32-cast query point, and
 * radius = sqrt(d2_fp64) - search_margin.  "Parity unpinned" applies to the PCL float path only.
 */
#include <stdint.h>
#include <stddef.h>
#include <math.h>

typedef struct kdo_tree kdo_tree;
int kdo_nearest3(const kdo_tree *t, double x, double y, double z, int64_t *idx, double *d2, double *pos3);
int64_t kdo_size(const kdo_tree *t);

/* same field order as pc_radius_params in include/pc_index.h */
typedef struct {
    double search_margin;
    double max_radius;
    double sample_range; /* < 0 disables the out-of-sensing-range early-out */
    double start[3];
} po_radius_params;

static inline double sq(double v) { return v * v; }

/* radiusSearch: early-outs first, then float32 cast of the search point, 1-NN, epilogue. */
double po_radius_search(const kdo_tree *t, const po_radius_params *P, const double p[3], int64_t *nn_idx)
{
    if (nn_idx) *nn_idx = -1;
    if (P->sample_range >= 0.0) {
        double dis = sqrt(sq(p[0] - P->start[0]) + sq(p[1] - P->start[1]) + sq(p[2] - P->start[2]));
        if (dis > P->sample_range + P->max_radius) return P->max_radius - P->search_margin;
    }
    if (kdo_size(t) == 0) return P->max_radius - P->search_margin;
    float fx = (float)p[0], fy = (float)p[1], fz = (float)p[2];
    int64_t idx; double d2;
    kdo_nearest3(t, (double)fx, (double)fy, (double)fz, &idx, &d2, NULL);
    if (nn_idx) *nn_idx = idx;
    double radius = sqrt(d2) - P->search_margin;
    return radius < P->max_radius ? radius : P->max_radius; /* std::min(radius, max_radius) */
}

int po_check_traj_pt_col(const kdo_tree *t, const po_radius_params *P, const double p[3])
{
    return po_radius_search(t, P, p, NULL) < 0.0 ? 1 : 0;
}

int po_radius_batch(const kdo_tree *t, const po_radius_params *P, const float *q, int64_t m, int64_t stride,
                    double *out_radius, int64_t *out_idx)
{
    for (int64_t k = 0; k < m; k++) {
        const float *f = q + k * stride;
        double p[3] = { (double)f[0], (double)f[1], (double)f[2] };
        int64_t idx;
        out_radius[k] = po_radius_search(t, P, p, &idx);
        if (out_idx) out_idx[k] = idx;
    }
    return 0;
}

static int po_factorial(int n)
{
    int f = 1;
    for (int i = n; i > 0; i--) f *= i;
    return f;
}

double po_binomial(int n, int k)
{
    return (double)(po_factorial(n) / (po_factorial(k) * po_factorial(n - k)));
}

/* coef_row: [x_0..x_n, y_0..y_n, z_0..z_n]; out = sum_j C(n,j) c_j u^j (1-u)^(n-j) per axis,
 * accumulated from 0 in j order, each term evaluated left to right with libm pow. */
void po_bezier_pos(const double *coef_row, int order, double u, double out[3])
{
    int nctrl = order + 1;
    for (int a = 0; a < 3; a++) {
        double acc = 0.0;
        for (int j = 0; j < nctrl; j++)
            acc += po_binomial(order, j) * coef_row[a * nctrl + j] * pow(u, (double)j) * pow(1 - u, (double)(order - j));
        out[a] = acc;
    }
}

/*
 * checkSafeTrajectory, restated as a full walk: every sample up to the stop horizon is
 * evaluated (the reference returns at the first colliding sample; its return value is
 * "first_hit >= 0").  Per sample k the scaled position (float32, as radiusSearch casts it) and
 * the radiusSearch value can be recorded.
 *
 * t_now        : max(0, odom stamp - trajectory start), the reference's t_s before the segment walk
 * coef, coef_ld: row-major segment matrix, row i = [x|y|z] blocks of order[i]+1 coefficients
 * returns the ordinal of the first colliding sample or -1; *n_samples = samples within the horizon.
 */
int64_t po_check_safe_trajectory(const kdo_tree *t, const po_radius_params *P,
                                 int32_t n_seg, const int32_t *order, const double *T,
                                 const double *coef, int64_t coef_ld,
                                 double t_now, double stop_time, double dt,
                                 int64_t cap, float *out_pts /*cap*3*/, double *out_radius /*cap*/,
                                 int64_t *n_samples, double *min_radius)
{
    double t_s = t_now > 0.0 ? t_now : 0.0;
    int idx;
    for (idx = 0; idx < n_seg; ++idx) {
        if (t_s > T[idx] && idx + 1 < n_seg) t_s -= T[idx];
        else break;
    }
    int64_t k = 0, first_hit = -1;
    double rmin = INFINITY;
    double t_accu = 0.0;
    for (int i = idx; i < n_seg; i++) {
        double t_ss = (i == idx) ? t_s : 0.0;
        for (double tt = t_ss; tt < T[i]; tt += dt) {
            t_accu += dt;
            if (t_accu > stop_time) break;
            double pos[3];
            po_bezier_pos(coef + (int64_t)i * coef_ld, order[i], tt / T[i], pos);
            pos[0] *= T[i]; pos[1] *= T[i]; pos[2] *= T[i];
            double r = po_radius_search(t, P, pos, NULL);
            if (k < cap) {
                if (out_pts) { out_pts[3 * k] = (float)pos[0]; out_pts[3 * k + 1] = (float)pos[1]; out_pts[3 * k + 2] = (float)pos[2]; }
                if (out_radius) out_radius[k] = r;
            }
            if (r < rmin) rmin = r;
            if (r < 0.0 && first_hit < 0) first_hit = k;
            k++;
        }
    }
    if (n_samples) *n_samples = k;
    if (min_radius) *min_radius = rmin;
    return first_hit;
}

/* ---- the sample stream ------------------------------------------------------------------------------------------------- */
/* same field order as pc_sampler in include/pc_index.h */
typedef struct {
    uint32_t engine_state, reserved;
    double goal_ratio, inlier_ratio;
    double end_pt[3];
    double lo[3], hi[3];
    double in_lo[3], in_hi[3];
} po_sampler;

/* minstd_rand0: x' = 16807 x mod 2147483647 */
static inline uint32_t po_lcg(uint32_t *st) { *st = (uint32_t)(((uint64_t)*st * 16807u) % 2147483647u); return *st; }

/* std::generate_canonical<double, 53>(minstd_rand0): range r = max - min + 1 = 2147483646, m = 2 draws,
 * sum = (x1 - 1) * 1 + (x2 - 1) * r, ret = sum / (double)(r * r) (the long double product rounded to double) */
static double po_canonical(uint32_t *st)
{
    const long double r = 2147483646.0L;
    double sum = 0.0, tmp = 1.0;
    for (int k = 2; k != 0; --k) {
        sum += (double)(po_lcg(st) - 1u) * tmp;
        tmp = (double)(tmp * r);
    }
    double ret = sum / tmp;
    if (ret >= 1.0) ret = nextafter(1.0, 0.0);
    return ret;
}

static inline double po_uniform(uint32_t *st, double a, double b) { return po_canonical(st) * (b - a) + a; }

/* k samples; S->engine_state is advanced */
void po_gen_samples(po_sampler *S, int64_t k, double *out3)
{
    uint32_t st = S->engine_state;
    for (int64_t i = 0; i < k; i++) {
        double *pt = out3 + 3 * i;
        const double bias = po_uniform(&st, 0.0, 1.0);
        if (bias <= S->goal_ratio) { pt[0] = S->end_pt[0]; pt[1] = S->end_pt[1]; pt[2] = S->end_pt[2]; continue; }
        if (bias > S->goal_ratio && bias <= (S->goal_ratio + S->inlier_ratio)) {
            for (int a = 0; a < 3; a++) pt[a] = po_uniform(&st, S->in_lo[a], S->in_hi[a]);
        } else {
            for (int a = 0; a < 3; a++) pt[a] = po_uniform(&st, S->lo[a], S->hi[a]);
        }
    }
    S->engine_state = st;
}

/* genNewNode's centre: the sample pulled onto the surface of the nearest node's sphere */
void po_steer(const double sample[3], const double node[3], float node_radius, double center[3])
{
    const double dis = sqrt(sq(node[0] - sample[0]) + sq(node[1] - sample[1]) + sq(node[2] - sample[2]));
    if (dis > node_radius) {
        const double steer_dis = node_radius / dis;
        for (int a = 0; a < 3; a++) center[a] = node[a] + (sample[a] - node[a]) * steer_dis;
    } else {
        for (int a = 0; a < 3; a++) center[a] = sample[a];
    }
}

void po_steer_batch(const double *samples, int64_t k, const double *node_coord, const float *node_radius, const int32_t *nearest, double *centers)
{
    for (int64_t j = 0; j < k; j++) po_steer(samples + 3 * j, node_coord + 3 * (int64_t)nearest[j], node_radius[nearest[j]], centers + 3 * j);
}
