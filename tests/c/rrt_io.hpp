// rrt_io.hpp -- input / output files shared by rrt_client.cpp and rrt_cpu_check.cpp (test code).
//   in : int64 n, int64 max_iter, int64 K, double prm[4] (safety, search, max_radius, range), double start[3], goal[3],
//        double box[6] (xl xh yl yh zl zh), double portions[2] (sample, goal), float pts[n*3]
//        optionally: int64 n2, int64 refine_iter, float pts2[n2*3]   -- a SECOND cloud message (re-validation phase)
//   out: records of int64 k, int64 nodes, int64 cloud_queries, double ms, double path[k*3], double radius[k];
//        per driver: after growth; with a second cloud also: after evaluate() on it, and after cycles of refine(refine_iter) + evaluate() until a path survives
#pragma once
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <vector>
#include "pc_rrt.hpp"

struct RrtInput {
    int64_t n = 0, max_iter = 0, K = 0, n2 = 0, refine_iter = 0;
    double prm[4], start[3], goal[3], box[6], portions[2];
    std::vector<float> pts, pts2;
    bool second = false;
};

static inline int rrt_read_input(const char *path, RrtInput &in)
{
    FILE *f = fopen(path, "rb");
    if (!f) return 3;
    if (fread(&in.n, 8, 1, f) != 1 || fread(&in.max_iter, 8, 1, f) != 1 || fread(&in.K, 8, 1, f) != 1 || fread(in.prm, 8, 4, f) != 4 ||
        fread(in.start, 8, 3, f) != 3 || fread(in.goal, 8, 3, f) != 3 || fread(in.box, 8, 6, f) != 6 || fread(in.portions, 8, 2, f) != 2) return 4;
    in.pts.resize((size_t)in.n * 3);
    if (fread(in.pts.data(), 12, (size_t)in.n, f) != (size_t)in.n) return 4;
    if (fread(&in.n2, 8, 1, f) == 1) {
        if (fread(&in.refine_iter, 8, 1, f) != 1) return 4;
        in.pts2.resize((size_t)in.n2 * 3);
        if (fread(in.pts2.data(), 12, (size_t)in.n2, f) != (size_t)in.n2) return 4;
        in.second = true;
    }
    fclose(f);
    return 0;
}

static inline void rrt_setup(pc::SafeRegionRrtStarDriver &d, const RrtInput &in)
{
    d.setParam(in.prm[0], in.prm[1], in.prm[2], in.prm[3]);
    d.reset();
    d.setPt(in.start, in.goal, in.box[0], in.box[1], in.box[2], in.box[3], in.box[4], in.box[5], in.prm[3], (int)in.max_iter,
            in.portions[0], in.portions[1]);
}

static inline void rrt_dump(FILE *o, pc::SafeRegionRrtStarDriver &d, double ms)
{
    int64_t k = (int64_t)d.radius.size(), nodes = (int64_t)d.nodeCount(), cq = d.cloud_queries;
    fwrite(&k, 8, 1, o); fwrite(&nodes, 8, 1, o); fwrite(&cq, 8, 1, o); fwrite(&ms, 8, 1, o);
    fwrite(d.path.data(), 8, d.path.size(), o);
    fwrite(d.radius.data(), 8, d.radius.size(), o);
}

struct RrtTimer {
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    double ms() { auto t1 = std::chrono::steady_clock::now(); double v = std::chrono::duration<double, std::milli>(t1 - t0).count(); t0 = t1; return v; }
};

// growth on the first cloud, then (if the input holds a second cloud message) re-validation and refinement on it.
// set_cloud(which) switches the radius provider's cloud (0 = first, 1 = second).
template <class SetCloud>
static inline void rrt_run(FILE *o, pc::SafeRegionRrtStarDriver &d, const RrtInput &in, bool batched, SetCloud set_cloud)
{
    set_cloud(0);
    rrt_setup(d, in);
    RrtTimer t;
    if (batched) d.expandBatched((int)in.max_iter, (int)in.K); else d.expand((int)in.max_iter);
    rrt_dump(o, d, t.ms());
    if (!in.second) return;
    set_cloud(1);
    t.ms();
    d.evaluate();
    rrt_dump(o, d, t.ms());
    // the planner's cycle: refine, then lazily re-validate whatever became the best path.  Nodes off the old path still carry
    // radii of the older cloud, so rewiring can pull a fresh path through a stale sphere and the next evaluate() rejects it
    // again (one stale sphere per cycle); run until a path survives, at most 8 cycles
    for (int cycle = 0; cycle < 8; cycle++) {
        if (batched) d.refineBatched((int)in.refine_iter, (int)in.K); else d.refine((int)in.refine_iter);
        d.evaluate();
        if (d.path_exist_status) break;
    }
    rrt_dump(o, d, t.ms());
}
