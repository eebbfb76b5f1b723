// rrt_cpu_check.cpp -- CPU-only run of include/pc_rrt.hpp with the oracle as radius provider (TEST-ONLY): exercises the
// restated expansion / re-validation / refinement logic (sequential and speculative-batch drivers) without a GPU.
// File formats: rrt_io.hpp.  Output = the records of the sequential driver, then those of the batched driver.
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
    fclose(o);
    kdo_free(kt[0]); kdo_free(kt[1]);
    return 0;
}
