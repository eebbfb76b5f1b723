// rrt_ref_check.cpp -- TEST-ONLY: include/pc_rrt.hpp in reference-quirks mode against the UNMODIFIED reference planner
// (Planner/src/corridor_finder.cpp compiled into oracle/_ref/libplanner_ref.so).  Both run the same scenario -- growth on a
// first cloud, then refine + evaluate cycles on a second one -- and after every phase the corridor (path, radii) and the
// whole node list (centre, radius, cost, parent, validity) must be bit-identical.
//   mode "ref": the driver's radius provider is the compiled reference's own radiusSearch -> pins the restated TREE LOGIC
//   mode "gpu": the provider is pc::SafeRegionCloud over libpcindex.so (PC_ARITH_PCL_FLOAT)  -> pins the GPU path through it
// Input file: rrt_io.hpp.  Prints one line per phase: "<phase> nodes=<n> path=<k> queries_ref=<q> queries_drv=<q> identical=<0|1>".
#include <cstring>
#include "rrt_io.hpp"
#include "pc_corridor.hpp"

extern "C" {
int rp_create(double, double, double, double);
void rp_set_input(const float *, long long, long long);
void rp_reset(void);
void rp_set_pt(const double *, const double *, double, double, double, double, double, double, double, int, double, double);
void rp_expand(double);
void rp_refine(double);
void rp_evaluate(void);
int rp_get_path(double *, double *, int);
int rp_get_tree(double *, int);
void rp_stats(long long *);
void rp_radius_batch(const double *, long long, double *);
}

static bool same_bits(const double *a, const double *b, size_t n) { return n == 0 || memcmp(a, b, n * sizeof(double)) == 0; }

static int compare(const char *phase, pc::SafeRegionRrtStarDriver &d)
{
    std::vector<double> rpath(3 * 4096), rrad(4096);
    long long st[6];
    rp_stats(st);
    int rk = rp_get_path(rpath.data(), rrad.data(), 4096);
    if (!st[1]) rk = 0;                                   // no path: the reference reports a 3 x 3 identity placeholder
    std::vector<double> rtree((size_t)st[0] * 7 + 7);
    const int rn = rp_get_tree(rtree.data(), (int)st[0] + 1);
    bool ok = rk == (int)d.radius.size() && (bool)st[1] == d.path_exist_status && rn == (int)d.nodeCount();
    if (ok) ok = same_bits(rpath.data(), d.path.data(), (size_t)rk * 3) && same_bits(rrad.data(), d.radius.data(), (size_t)rk);
    if (ok) {
        const std::vector<pc::RrtNode *> &nl = d.nodeList();
        for (int i = 0; i < rn && ok; i++) {
            const pc::RrtNode *p = nl[(size_t)i];
            int parent = -1;
            if (p->pre) for (int j = 0; j < rn; j++) if (nl[(size_t)j] == p->pre) { parent = j; break; }
            const double mine[7] = { p->coord[0], p->coord[1], p->coord[2], (double)p->radius, (double)p->g, (double)parent, p->valid ? 1.0 : 0.0 };
            ok = same_bits(mine, &rtree[(size_t)i * 7], 7);
            if (!ok) fprintf(stderr, "%s: node %d differs: mine (%.17g %.17g %.17g r=%.9g g=%.9g par=%g) ref (%.17g %.17g %.17g r=%.9g g=%.9g par=%g)\n", phase, i,
                             mine[0], mine[1], mine[2], mine[3], mine[4], mine[5], rtree[i * 7], rtree[i * 7 + 1], rtree[i * 7 + 2], rtree[i * 7 + 3], rtree[i * 7 + 4], rtree[i * 7 + 5]);
        }
    }
    printf("%s nodes=%d/%d path=%d/%d queries_ref=%lld queries_drv=%lld identical=%d\n", phase, rn, (int)d.nodeCount(), rk, (int)d.radius.size(),
           st[2], (long long)d.cloud_queries, ok ? 1 : 0);
    return ok ? 0 : 1;
}

int main(int argc, char **argv)
{
    if (argc != 3) return 2;
    RrtInput in;
    if (int rc = rrt_read_input(argv[1], in)) return rc;
    const bool gpu = strcmp(argv[2], "gpu") == 0;
    rp_create(in.prm[0], in.prm[1], in.prm[2], in.prm[3]);

    pc::SafeRegionCloud *cloud = nullptr;
    if (gpu) {
        cloud = new pc::SafeRegionCloud(0, in.n + (in.second ? in.n2 : 0));
        cloud->setParam(in.prm[0], in.prm[1], in.prm[2], in.prm[3]);
        cloud->setPt(in.start, in.prm[3]);
        if (cloud->setRadiusArith(PC_ARITH_PCL_FLOAT) != PC_OK) return 5;
    }
    auto provider = [&](const double *c, int m, double *out) {
        if (gpu) for (int i = 0; i < m; i++) out[i] = cloud->radiusSearch(c + 3 * i);
        else rp_radius_batch(c, m, out);
    };
    pc::SafeRegionRrtStarDriver d(provider);
    d.setReferenceQuirks(true);

    int bad = 0;
    // phase 1: planInitialTraj (sim_planning_demo.cpp:346-354): reset, setPt, SafeRegionExpansion
    rp_set_input(in.pts.data(), in.n, 3);
    rp_reset();
    rp_set_pt(in.start, in.goal, in.box[0], in.box[1], in.box[2], in.box[3], in.box[4], in.box[5], in.prm[3], (int)in.max_iter, in.portions[0], in.portions[1]);
    rp_expand((double)in.max_iter);
    if (gpu && cloud->setInput(in.pts.data(), in.n, 3) != PC_OK) return 6;
    rrt_setup(d, in);
    d.expand((int)in.max_iter);
    bad += compare("expand", d);
    if (in.second) {
        // a new cloud message (rcvPointCloudCallBack, :159-167), then planning ticks: SafeRegionRefine + SafeRegionEvaluate (:412-413)
        for (int cycle = 0; cycle < 3; cycle++) {
            if (cycle == 0) {
                rp_set_input(in.pts2.data(), in.n2, 3);
                if (gpu && cloud->setInput(in.pts2.data(), in.n2, 3) != PC_OK) return 6;
            }
            rp_refine((double)in.refine_iter);
            d.refine((int)in.refine_iter);
            bad += compare(cycle == 0 ? "refine1" : (cycle == 1 ? "refine2" : "refine3"), d);
            rp_evaluate();
            d.evaluate();
            bad += compare(cycle == 0 ? "evaluate1" : (cycle == 1 ? "evaluate2" : "evaluate3"), d);
        }
    }
    delete cloud;
    return bad ? 1 : 0;
}
