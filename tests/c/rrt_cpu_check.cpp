// rrt_cpu_check.cpp -- CPU-only run of include/pc_rrt.hpp with the oracle as radius provider (TEST-ONLY): exercises the
// restated expansion / re-validation / refinement logic (sequential and speculative-batch drivers) without a GPU.
// File formats: rrt_io.hpp.  Output = the records of the sequential driver, then those of the batched driver, then the batched
// driver with a brute-force snapshot nearest-vertex provider, then the batched driver with a CPU stand-in for the
// device-generated batch provider (setDeviceBatch: the stream is re-drawn from the ENGINE STATE by oracle/planner_oracle.c's
// po_gen_samples) -- the last two must agree bit for bit: that is the host-side contract of pc_expand_batch.
#include "rrt_io.hpp"

extern "C" {
struct kdo_tree;
kdo_tree *kdo_create(void);
void kdo_free(kdo_tree *);
int kdo_build(kdo_tree *, const float *xyz, int64_t n, int64_t stride, const int64_t *order);
struct po_radius_params { double search_margin, max_radius, sample_range, start[3]; };
double po_radius_search(const kdo_tree *, const po_radius_params *, const double p[3], int64_t *nn_idx);
struct po_sampler { uint32_t engine_state, reserved; double goal_ratio, inlier_ratio, end_pt[3], lo[3], hi[3], in_lo[3], in_hi[3]; };
void po_gen_samples(po_sampler *S, int64_t k, double *out3);
void po_steer(const double sample[3], const double node[3], float node_radius, double center[3]);
}

// nearest of n float32 positions by fp64 distance, ties to the smallest index (what pc_nearest_batch returns)
static int brute_nearest(const float *pos, int n, const float q[3])
{
    int best = -1; double bd = 0.0;
    for (int i = 0; i < n; i++) {
        double s = 0.0;
        for (int a = 0; a < 3; a++) { const double d = (double)pos[3 * i + a] - (double)q[a]; s += d * d; }
        if (best < 0 || s < bd) { best = i; bd = s; }
    }
    return best;
}

int main(int argc, char **argv)
{
    if (argc != 3) return 2;
    RrtInput in;
    if (int rc = rrt_read_input(argv[1], in)) return rc;
    kdo_tree *kt[2] = { kdo_create(), kdo_create() };
    if (kdo_build(kt[0], in.pts.data(), in.n, 3, nullptr)) return 6;
    if (in.second && kdo_build(kt[1], in.pts2.data(), in.n2, 3, nullptr)) return 6;
    int cur = 0;
    po_radius_params P{ in.prm[1], in.prm[2], in.prm[3], { in.start[0], in.start[1], in.start[2] } };
    auto provider = [&](const double *c, int m, double *out) {
        for (int i = 0; i < m; i++) out[i] = (double)(float)po_radius_search(kt[cur], &P, c + 3 * i, nullptr);
    };
    FILE *o = fopen(argv[2], "wb");
    if (!o) return 7;
    for (int mode = 0; mode < 2; mode++) {
        pc::SafeRegionRrtStarDriver d(provider);
        rrt_run(o, d, in, mode == 1, [&](int which) { cur = which; });
    }
    auto snap_nearest = [&](const float *node_pos, int n_nodes, const float *samples, int k, int32_t *out_nearest) {
        for (int j = 0; j < k; j++) out_nearest[j] = brute_nearest(node_pos, n_nodes, samples + 3 * j);
    };
    {
        pc::SafeRegionRrtStarDriver d(provider);
        d.setSnapshotNearest(snap_nearest);
        rrt_run(o, d, in, true, [&](int which) { cur = which; });
    }
    {
        pc::SafeRegionRrtStarDriver d(provider);
        d.setSnapshotNearest(snap_nearest);
        d.setDeviceBatch([&](const pc::SafeRegionRrtStarDriver::DeviceBatchRequest &rq, std::vector<double> &centers, std::vector<double> &radii) {
            po_sampler S;
            S.engine_state = rq.engine_state; S.reserved = 0; S.goal_ratio = rq.goal_ratio; S.inlier_ratio = rq.inlier_ratio;
            for (int a = 0; a < 3; a++) { S.end_pt[a] = rq.end_pt[a]; S.lo[a] = rq.lo[a]; S.hi[a] = rq.hi[a]; S.in_lo[a] = rq.in_lo[a]; S.in_hi[a] = rq.in_hi[a]; }
            std::vector<double> smp((size_t)rq.k * 3);
            po_gen_samples(&S, rq.k, smp.data());
            std::vector<float> pos((size_t)rq.n_nodes * 3);
            for (size_t i = 0; i < pos.size(); i++) pos[i] = (float)rq.node_coord[i];
            centers.clear(); radii.clear();
            for (int j = 0; j < rq.k; j++) {
                const float q[3] = { (float)smp[3 * j], (float)smp[3 * j + 1], (float)smp[3 * j + 2] };
                const int nn = brute_nearest(pos.data(), rq.n_nodes, q);
                if (nn < 0 || !rq.node_valid[nn]) continue;
                double c[3], r;
                po_steer(&smp[(size_t)3 * j], rq.node_coord + 3 * nn, rq.node_radius[nn], c);
                provider(c, 1, &r);
                if (c[2] < rq.z_l || (float)r < rq.safety_margin) continue;
                centers.insert(centers.end(), c, c + 3); radii.push_back(r);
            }
            return S.engine_state;
        });
        rrt_run(o, d, in, true, [&](int which) { cur = which; });
        if (d.device_batches == 0) return 14;
    }
    fclose(o);
    kdo_free(kt[0]); kdo_free(kt[1]);
    return 0;
}
