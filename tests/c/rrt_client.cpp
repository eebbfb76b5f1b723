// rrt_client.cpp -- drives include/pc_rrt.hpp (restated safe-region RRT* expansion) with three radius providers:
//   A  GPU, one query per iteration        (pc::SafeRegionCloud::radiusSearch(double[3]))
//   B  CPU oracle, one query per iteration (po_radius_search of oracle/planner_oracle.c)   -- TEST-ONLY checker
//   C  GPU, speculative batches of K       (pc_radius_batch on the float32-cast centres)
// A and B must produce bit-identical corridors (replay mode); C is validated by the pytest against the oracle.
// usage: rrt_client <in.bin> <out.bin>
//   in : int64 n, int64 max_iter, int64 K, double prm[4] (safety, search, max_radius, range), double start[3], goal[3],
//        double box[6] (xl xh yl yh zl zh), double portions[2] (sample, goal), float pts[n*3]
//   out: for each of A, B, C: int64 k, int64 nodes, int64 cloud_queries, double ms, double path[k*3], double radius[k]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "pc_corridor.hpp"
#include "pc_rrt.hpp"

extern "C" {
struct kdo_tree;
kdo_tree *kdo_create(void);
void kdo_free(kdo_tree *);
int kdo_build(kdo_tree *, const float *xyz, int64_t n, int64_t stride, const int64_t *order);
struct po_radius_params { double search_margin, max_radius, sample_range, start[3]; };
double po_radius_search(const kdo_tree *, const po_radius_params *, const double p[3], int64_t *nn_idx);
}

static void dump(FILE *o, pc::SafeRegionRrtStarDriver &d, double ms)
{
    int64_t k = (int64_t)d.radius.size(), nodes = (int64_t)d.nodeCount(), cq = d.cloud_queries;
    fwrite(&k, 8, 1, o); fwrite(&nodes, 8, 1, o); fwrite(&cq, 8, 1, o); fwrite(&ms, 8, 1, o);
    fwrite(d.path.data(), 8, d.path.size(), o);
    fwrite(d.radius.data(), 8, d.radius.size(), o);
}

int main(int argc, char **argv)
{
    if (argc != 3) return 2;
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 3;
    int64_t n, max_iter, K;
    double prm[4], start[3], goal[3], box[6], portions[2];
    if (fread(&n, 8, 1, f) != 1 || fread(&max_iter, 8, 1, f) != 1 || fread(&K, 8, 1, f) != 1 || fread(prm, 8, 4, f) != 4 ||
        fread(start, 8, 3, f) != 3 || fread(goal, 8, 3, f) != 3 || fread(box, 8, 6, f) != 6 || fread(portions, 8, 2, f) != 2) return 4;
    std::vector<float> pts((size_t)n * 3);
    if (fread(pts.data(), 12, (size_t)n, f) != (size_t)n) return 4;
    fclose(f);

    pc::SafeRegionCloud cloud(0, n);
    cloud.setParam(prm[0], prm[1], prm[2], prm[3]);
    cloud.setPt(start, prm[3]);
    if (cloud.setInput(pts.data(), n, 3) != PC_OK) return 5;

    kdo_tree *kt = kdo_create();
    if (kdo_build(kt, pts.data(), n, 3, nullptr)) return 6;
    po_radius_params P{ prm[1], prm[2], prm[3], { start[0], start[1], start[2] } };

    auto setup = [&](pc::SafeRegionRrtStarDriver &d) {
        d.setParam(prm[0], prm[1], prm[2], prm[3]);
        d.reset();
        d.setPt(start, goal, box[0], box[1], box[2], box[3], box[4], box[5], prm[3], (int)max_iter, portions[0], portions[1]);
    };
    FILE *o = fopen(argv[2], "wb");
    if (!o) return 7;
    using clk = std::chrono::steady_clock;

    pc::SafeRegionRrtStarDriver A([&](const double *c, int m, double *out) { for (int i = 0; i < m; i++) out[i] = cloud.radiusSearch(c + 3 * i); });
    setup(A);
    auto t0 = clk::now();
    A.expand((int)max_iter);
    dump(o, A, std::chrono::duration<double, std::milli>(clk::now() - t0).count());

    pc::SafeRegionRrtStarDriver B([&](const double *c, int m, double *out) {
        for (int i = 0; i < m; i++) out[i] = (double)(float)po_radius_search(kt, &P, c + 3 * i, nullptr);   // the GPU returns float32
    });
    setup(B);
    t0 = clk::now();
    B.expand((int)max_iter);
    dump(o, B, std::chrono::duration<double, std::milli>(clk::now() - t0).count());

    std::vector<float> qf, rf;
    pc::SafeRegionRrtStarDriver C([&](const double *c, int m, double *out) {
        qf.resize((size_t)m * 3); rf.resize((size_t)m);
        for (size_t i = 0; i < qf.size(); i++) qf[i] = (float)c[i];
        if (cloud.radiusSearch(qf.data(), m, 3, rf.data()) != PC_OK) exit(8);
        for (int i = 0; i < m; i++) out[i] = rf[(size_t)i];
    });
    setup(C);
    t0 = clk::now();
    C.expandBatched((int)max_iter, (int)K);
    dump(o, C, std::chrono::duration<double, std::milli>(clk::now() - t0).count());
    fclose(o);
    kdo_free(kt);
    return 0;
}
