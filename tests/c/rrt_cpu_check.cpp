// rrt_cpu_check.cpp -- CPU-only run of include/pc_rrt.hpp with the oracle as radius provider (TEST-ONLY): exercises the
// restated expansion logic (sequential and speculative-batch drivers) without a GPU.  Same file formats as rrt_client.cpp,
// output = two records (sequential, batched).
#include <chrono>
#include <cstdio>
#include <vector>
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
    kdo_tree *kt = kdo_create();
    if (kdo_build(kt, pts.data(), n, 3, nullptr)) return 6;
    po_radius_params P{ prm[1], prm[2], prm[3], { start[0], start[1], start[2] } };
    auto provider = [&](const double *c, int m, double *out) {
        for (int i = 0; i < m; i++) out[i] = (double)(float)po_radius_search(kt, &P, c + 3 * i, nullptr);
    };
    FILE *o = fopen(argv[2], "wb");
    if (!o) return 7;
    using clk = std::chrono::steady_clock;
    for (int mode = 0; mode < 2; mode++) {
        pc::SafeRegionRrtStarDriver d(provider);
        d.setParam(prm[0], prm[1], prm[2], prm[3]);
        d.reset();
        d.setPt(start, goal, box[0], box[1], box[2], box[3], box[4], box[5], prm[3], (int)max_iter, portions[0], portions[1]);
        auto t0 = clk::now();
        if (mode == 0) d.expand((int)max_iter); else d.expandBatched((int)max_iter, (int)K);
        dump(o, d, std::chrono::duration<double, std::milli>(clk::now() - t0).count());
    }
    fclose(o);
    kdo_free(kt);
    return 0;
}
