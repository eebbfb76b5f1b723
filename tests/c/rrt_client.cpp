// rrt_client.cpp -- drives include/pc_rrt.hpp (restated safe-region RRT* expansion, re-validation and refinement) with
// three radius providers:
//   A  GPU, one query per iteration        (pc::SafeRegionCloud::radiusSearch(double[3]))
//   B  CPU oracle, one query per iteration (po_radius_search of oracle/planner_oracle.c)   -- TEST-ONLY checker
//   C  GPU, speculative batches of K       (pc_radius_batch on the float32-cast centres)
//   D  as C + the batch's nearest-vertex queries on the GPU (pc::NodeSnapshotIndex), checked against the CPU node tree
//   E  as C + the 2 x radius neighbourhoods of treeRewire (kd_nearest_rangef, corridor_finder.cpp:462-464) answered for the
//      whole batch by ONE pc_range_batch on the snapshot index: must reproduce C bit for bit
//   F  as D, but while no path is known the whole snapshot phase of a batch (samples, nearest vertex, steering, radius, the
//      loop's early rejections) is ONE pc_expand_batch call that generates the sample stream on the device: must reproduce D
//   G  F + E: device-generated batches and the rewire neighbourhoods from one pc_range_batch per batch: must reproduce F
// A and B must produce bit-identical corridors (replay mode); C is validated by the pytest against the oracle.
// usage: rrt_client <in.bin> <out.bin>      file formats: rrt_io.hpp; output = the records of A, then B, then C, then D, then E, then F, then G
#include <cstdlib>
#include "pc_corridor.hpp"
#include "rrt_io.hpp"

extern "C" {
struct kdo_tree;
kdo_tree *kdo_create(void);
void kdo_free(kdo_tree *);
int kdo_build(kdo_tree *, const float *xyz, int64_t n, int64_t stride, const int64_t *order);
struct po_radius_params { double search_margin, max_radius, sample_range, start[3]; };
double po_radius_search(const kdo_tree *, const po_radius_params *, const double p[3], int64_t *nn_idx);
}

int main(int argc, char **argv)
{
    if (argc != 3) return 2;
    RrtInput in;
    if (int rc = rrt_read_input(argv[1], in)) return rc;

    pc::SafeRegionCloud cloud(0, in.n > in.n2 ? in.n : in.n2);
    cloud.setParam(in.prm[0], in.prm[1], in.prm[2], in.prm[3]);
    cloud.setPt(in.start, in.prm[3]);
    // a new cloud message = a full index rebuild (setInput, corridor_finder.cpp:93-99)
    auto gpu_cloud = [&](int which) {
        if (cloud.setInput(which ? in.pts2.data() : in.pts.data(), which ? in.n2 : in.n, 3) != PC_OK) exit(5);
    };

    kdo_tree *kt[2] = { kdo_create(), kdo_create() };
    if (kdo_build(kt[0], in.pts.data(), in.n, 3, nullptr)) return 6;
    if (in.second && kdo_build(kt[1], in.pts2.data(), in.n2, 3, nullptr)) return 6;
    int cur = 0;
    po_radius_params P{ in.prm[1], in.prm[2], in.prm[3], { in.start[0], in.start[1], in.start[2] } };

    FILE *o = fopen(argv[2], "wb");
    if (!o) return 7;

    pc::SafeRegionRrtStarDriver A([&](const double *c, int m, double *out) { for (int i = 0; i < m; i++) out[i] = cloud.radiusSearch(c + 3 * i); });
    rrt_run(o, A, in, false, gpu_cloud);

    pc::SafeRegionRrtStarDriver B([&](const double *c, int m, double *out) {
        for (int i = 0; i < m; i++) out[i] = (double)(float)po_radius_search(kt[cur], &P, c + 3 * i, nullptr);   // the GPU returns float32
    });
    rrt_run(o, B, in, false, [&](int which) { cur = which; });

    std::vector<float> qf, rf;
    pc::SafeRegionRrtStarDriver C([&](const double *c, int m, double *out) {
        qf.resize((size_t)m * 3); rf.resize((size_t)m);
        for (size_t i = 0; i < qf.size(); i++) qf[i] = (float)c[i];
        if (cloud.radiusSearch(qf.data(), m, 3, rf.data()) != PC_OK) exit(8);
        for (int i = 0; i < m; i++) out[i] = rf[(size_t)i];
    });
    rrt_run(o, C, in, true, gpu_cloud);

    // D  as C, plus the nearest-vertex queries of every batch answered on the GPU against the frozen node set;
    //    every answer is checked here against the driver's own CPU node tree (same fp64 distance; ties may pick another node)
    pc::NodeSnapshotIndex nodes(0, 1 << 16);
    pc::SafeRegionRrtStarDriver D([&](const double *c, int m, double *out) {
        qf.resize((size_t)m * 3); rf.resize((size_t)m);
        for (size_t i = 0; i < qf.size(); i++) qf[i] = (float)c[i];
        if (cloud.radiusSearch(qf.data(), m, 3, rf.data()) != PC_OK) exit(8);
        for (int i = 0; i < m; i++) out[i] = rf[(size_t)i];
    });
    std::vector<float> nd2;
    int batch_no = 0;
    D.setSnapshotNearest([&](const float *node_pos, int n_nodes, const float *samples, int k, int32_t *out_nearest) {
        nd2.resize((size_t)k);
        if (nodes.nearest(node_pos, n_nodes, samples, k, out_nearest, nd2.data()) != PC_OK) exit(9);
        if ((batch_no++ & 3) != 0) return;                      // check every fourth batch (keeps the timing meaningful)
        for (int j = 0; j < k; j++) {
            double want = 0.0;
            D.nodeTree().nearest(samples + 3 * j, &want);
            const float *p = node_pos + 3 * (size_t)out_nearest[j];
            double got = 0.0;
            for (int a = 0; a < 3; a++) { const double d = (double)p[a] - (double)samples[3 * j + a]; got += d * d; }
            if (got != want || nd2[(size_t)j] != (float)want) { fprintf(stderr, "node-tree nearest mismatch: sample %d got %.17g want %.17g\n", j, got, want); exit(10); }
        }
    });
    rrt_run(o, D, in, true, gpu_cloud);

    pc::SafeRegionRrtStarDriver E([&](const double *c, int m, double *out) {
        qf.resize((size_t)m * 3); rf.resize((size_t)m);
        for (size_t i = 0; i < qf.size(); i++) qf[i] = (float)c[i];
        if (cloud.radiusSearch(qf.data(), m, 3, rf.data()) != PC_OK) exit(8);
        for (int i = 0; i < m; i++) out[i] = rf[(size_t)i];
    });
    pc::NodeSnapshotIndex nodes_e(0, 1 << 16);
    double e_build_ms = 0.0, e_range_ms = 0.0, e_max_ms = 0.0;
    int e_calls = 0;
    E.setSnapshotRange([&](const float *node_pos, int n_nodes, const float *centers, const float *ranges, int k,
                           std::vector<int64_t> &offsets, std::vector<int32_t> &idx) {
        RrtTimer t;
        if (nodes_e.build(node_pos, n_nodes) != PC_OK) { fprintf(stderr, "snapshot build: %s\n", nodes_e.lastError()); exit(11); }
        const double b = t.ms();
        if (nodes_e.range(centers, ranges, k, offsets, idx) != PC_OK) { fprintf(stderr, "snapshot range: %s\n", nodes_e.lastError()); exit(11); }
        const double r = t.ms();
        e_build_ms += b; e_range_ms += r; e_calls++;
        if (b + r > e_max_ms) e_max_ms = b + r;
    });
    rrt_run(o, E, in, true, gpu_cloud);
    fprintf(stderr, "E: %d range batches, node index builds %.1f ms, range calls %.1f ms, slowest batch %.1f ms\n", e_calls, e_build_ms, e_range_ms, e_max_ms);
    // the corridor as the planner publishes it (PolynomialTrajectoryExtra.path_* / radii): first sphere repeated
    {
        const pc::CorridorExport ex = pc::exportCorridor(E.path.data(), E.radius.data(), (int64_t)E.radius.size());
        const size_t k = E.radius.size();
        if (ex.size() != (k ? k + 1 : 0)) return 12;
        for (size_t i = 0; i < k; i++)
            if (ex.path_x[i + 1] != E.path[3 * i] || ex.path_y[i + 1] != E.path[3 * i + 1] || ex.path_z[i + 1] != E.path[3 * i + 2] || ex.radii[i + 1] != E.radius[i]) return 12;
        if (k && (ex.path_x[0] != E.path[0] || ex.path_y[0] != E.path[1] || ex.path_z[0] != E.path[2] || ex.radii[0] != E.radius[0])) return 12;
    }
    // F: the device-generated batch path
    {
        pc::SafeRegionRrtStarDriver F([&](const double *c, int m, double *out) {
            qf.resize((size_t)m * 3); rf.resize((size_t)m);
            for (size_t i = 0; i < qf.size(); i++) qf[i] = (float)c[i];
            if (cloud.radiusSearch(qf.data(), m, 3, rf.data()) != PC_OK) exit(8);
            for (int i = 0; i < m; i++) out[i] = rf[(size_t)i];
        });
        pc::NodeSnapshotIndex nodes_f(0, 1 << 16), nodes_dev(0, 1 << 16);
        F.setSnapshotNearest([&](const float *node_pos, int n_nodes, const float *samples, int k, int32_t *out_nearest) {
            if (nodes_f.nearest(node_pos, n_nodes, samples, k, out_nearest, nullptr) != PC_OK) exit(9);
        });
        pc::DeviceExpansion dev(cloud, nodes_dev);
        F.setDeviceBatch([&](const pc::SafeRegionRrtStarDriver::DeviceBatchRequest &rq, std::vector<double> &centers, std::vector<double> &radii) {
            pc_sampler sm;
            sm.engine_state = rq.engine_state; sm.reserved = 0; sm.goal_ratio = rq.goal_ratio; sm.inlier_ratio = rq.inlier_ratio;
            for (int a = 0; a < 3; a++) { sm.end_pt[a] = rq.end_pt[a]; sm.lo[a] = rq.lo[a]; sm.hi[a] = rq.hi[a]; sm.in_lo[a] = rq.in_lo[a]; sm.in_hi[a] = rq.in_hi[a]; }
            pc_node_set set{ rq.n_nodes, rq.node_coord, rq.node_radius, rq.node_valid };
            uint32_t after = 0;
            if (dev.run(sm, set, rq.z_l, rq.k, centers, radii, &after) != PC_OK) { fprintf(stderr, "device batch: %s\n", cloud.lastError()); exit(13); }
            return after;
        });
        rrt_run(o, F, in, true, gpu_cloud);
        if (F.device_batches == 0) return 14;
        fprintf(stderr, "F: %lld device batches\n", (long long)F.device_batches);
        // G: F + E -- device-generated batches AND the 2 x radius neighbourhoods of treeRewire from one pc_range_batch per batch:
        //    no per-sample work left on the host but the sequential insertion itself.  Must reproduce F.
        pc::SafeRegionRrtStarDriver G([&](const double *c, int m, double *out) {
            qf.resize((size_t)m * 3); rf.resize((size_t)m);
            for (size_t i = 0; i < qf.size(); i++) qf[i] = (float)c[i];
            if (cloud.radiusSearch(qf.data(), m, 3, rf.data()) != PC_OK) exit(8);
            for (int i = 0; i < m; i++) out[i] = rf[(size_t)i];
        });
        pc::NodeSnapshotIndex nodes_g(0, 1 << 16);
        G.setSnapshotNearest([&](const float *node_pos, int n_nodes, const float *samples, int k, int32_t *out_nearest) {
            if (nodes_f.nearest(node_pos, n_nodes, samples, k, out_nearest, nullptr) != PC_OK) exit(9);
        });
        G.setSnapshotRange([&](const float *node_pos, int n_nodes, const float *centers, const float *ranges, int k,
                               std::vector<int64_t> &offsets, std::vector<int32_t> &idx) {
            if (nodes_g.build(node_pos, n_nodes) != PC_OK || nodes_g.range(centers, ranges, k, offsets, idx) != PC_OK) exit(11);
        });
        G.setDeviceBatch([&](const pc::SafeRegionRrtStarDriver::DeviceBatchRequest &rq, std::vector<double> &centers, std::vector<double> &radii) {
            pc_sampler sm;
            sm.engine_state = rq.engine_state; sm.reserved = 0; sm.goal_ratio = rq.goal_ratio; sm.inlier_ratio = rq.inlier_ratio;
            for (int a = 0; a < 3; a++) { sm.end_pt[a] = rq.end_pt[a]; sm.lo[a] = rq.lo[a]; sm.hi[a] = rq.hi[a]; sm.in_lo[a] = rq.in_lo[a]; sm.in_hi[a] = rq.in_hi[a]; }
            pc_node_set set{ rq.n_nodes, rq.node_coord, rq.node_radius, rq.node_valid };
            uint32_t after = 0;
            if (dev.run(sm, set, rq.z_l, rq.k, centers, radii, &after) != PC_OK) { fprintf(stderr, "device batch: %s\n", cloud.lastError()); exit(13); }
            return after;
        });
        rrt_run(o, G, in, true, gpu_cloud);
    }
    fclose(o);
    kdo_free(kt[0]); kdo_free(kt[1]);
    return 0;
}
