// lbvh_frontier_check.cpp -- CPU emulation of the warp-per-query frontier walk (pc_query_coop_kernel of query_kernels.cuh) over
// the prefix-split index planned for the next round, and of the warp-per-query RANGE walk (pc_range_coop_kernel) with the
// per-leaf point counts of the records.  Checks exactness against brute force; reports dependent steps per query.
// usage: lbvh_frontier_check <points.bin> <queries.bin> <bound> <range>     (bound 0 = unbounded nearest; range = radius of the range queries)
#include <set>
#include <string>
#include "lbvh_host_build.hpp"

struct Entry { uint32_t ref; float d; uint32_t count; };

// one query, G lanes: every step the G most recently pushed entries are taken, one per lane
static int nearest_frontier(const LhIndex &ix, const float q[3], float thr0, double *best_out, int32_t *idx_out, int G)
{
    std::vector<Entry> F;
    F.push_back(Entry{ ix.root, 0.f, 0 });
    double best = INFINITY; int32_t idx = -1; float thr = thr0;
    int steps = 0;
    while (!F.empty()) {
        steps++;
        const int take = (int)std::min<size_t>(F.size(), (size_t)G);
        std::vector<Entry> lane(F.end() - take, F.end());              // lane 0 = top of the stack = lane.back()
        F.resize(F.size() - take);
        std::vector<Entry> far, near;                                  // pushed after the step: far children first, near on top
        float thr_step = thr;                                          // lanes test against the bound of the step's start ...
        for (int l = take - 1; l >= 0; l--) {
            const Entry e = lane[(size_t)l];
            if (!(e.d <= thr_step)) continue;
            if (e.ref & PC_REF_LEAF) {
                const pc_f4 *pt = ix.pts.data() + (e.ref & 0x7fffffffu);
                for (int i = 0; i < PC_LBVH_LEAF; i++) {
                    const float dx = pt[i].x - q[0], dy = pt[i].y - q[1], dz = pt[i].z - q[2];
                    const float d = dx * dx + dy * dy + dz * dz;
                    if (d <= thr_step) {                               // (a lane only knows its own improvements within a step)
                        const double ex = (double)pt[i].x - (double)q[0], ey = (double)pt[i].y - (double)q[1], ez = (double)pt[i].z - (double)q[2];
                        double v = ex * ex; v = v + ey * ey; v = v + ez * ez;
                        const int32_t id = (int32_t)pc_f2u(pt[i].w);
                        if (v < best || (v == best && (uint32_t)id < (uint32_t)idx)) { best = v; idx = id; thr = std::min(thr, lh_thr_from(v)); }
                    }
                }
            } else {
                const pc_f4 *r = ix.rec.data() + 4 * (int64_t)e.ref;
                const float d0 = pc_lbvh_box_d2(r[0], r[1], q[0], q[1], q[2]), d1 = pc_lbvh_box_d2(r[2], r[3], q[0], q[1], q[2]);
                const bool first0 = d0 <= d1;
                near.push_back(Entry{ pc_f2u(first0 ? r[0].w : r[2].w), first0 ? d0 : d1, 0 });
                far.push_back(Entry{ pc_f2u(first0 ? r[2].w : r[0].w), first0 ? d1 : d0, 0 });
            }
        }
        // ... and the children are filtered against the bound shared at the end of the step
        for (const Entry &e : far) if (e.d <= thr) F.push_back(e);
        for (auto it = near.rbegin(); it != near.rend(); ++it) if (it->d <= thr) F.push_back(*it);     // lane 0's near child ends on top
    }
    *best_out = best; *idx_out = idx;
    return steps;
}

// all points within r (inclusive), as original indices; leaves contribute only their own `count` points
static int range_frontier(const LhIndex &ix, const float q[3], double r, std::vector<int32_t> &hits, int G)
{
    hits.clear();
    const double r2 = r * r;
    float thr = (float)r2; if ((double)thr < r2) thr = nextafterf(thr, INFINITY); thr = nextafterf(thr * 1.00000095367431640625f, INFINITY);
    std::vector<Entry> F;
    F.push_back(Entry{ ix.root, 0.f, (uint32_t)std::min<int64_t>(ix.n, PC_LBVH_LEAF) });
    int steps = 0;
    while (!F.empty()) {
        steps++;
        const int take = (int)std::min<size_t>(F.size(), (size_t)G);
        std::vector<Entry> lane(F.end() - take, F.end());
        F.resize(F.size() - take);
        for (const Entry &e : lane) {
            if (e.ref & PC_REF_LEAF) {
                const pc_f4 *pt = ix.pts.data() + (e.ref & 0x7fffffffu);
                for (uint32_t i = 0; i < e.count; i++) {
                    const float dx = pt[i].x - q[0], dy = pt[i].y - q[1], dz = pt[i].z - q[2];
                    if (dx * dx + dy * dy + dz * dz <= thr) {
                        const double ex = (double)pt[i].x - (double)q[0], ey = (double)pt[i].y - (double)q[1], ez = (double)pt[i].z - (double)q[2];
                        double v = ex * ex; v = v + ey * ey; v = v + ez * ez;
                        if (v <= r2) hits.push_back((int32_t)pc_f2u(pt[i].w));
                    }
                }
            } else {
                const pc_f4 *rc = ix.rec.data() + 4 * (int64_t)e.ref;
                if (pc_lbvh_box_d2(rc[0], rc[1], q[0], q[1], q[2]) <= thr) F.push_back(Entry{ pc_f2u(rc[0].w), 0.f, pc_f2u(rc[1].w) });
                if (pc_lbvh_box_d2(rc[2], rc[3], q[0], q[1], q[2]) <= thr) F.push_back(Entry{ pc_f2u(rc[2].w), 0.f, pc_f2u(rc[3].w) });
            }
        }
    }
    std::sort(hits.begin(), hits.end());
    return steps;
}

int main(int argc, char **argv)
{
    if (argc < 5) return 2;
    const std::vector<float> P = lh_read_f32(argv[1]), Q = lh_read_f32(argv[2]);
    const double bound = atof(argv[3]), range = atof(argv[4]);
    const int64_t m = (int64_t)Q.size() / 3;
    LhIndex ix; ix.build(P);
    float thr0 = FLT_MAX;
    if (bound > 0) { const double b2 = bound * bound * (1.0 + 1e-6); thr0 = nextafterf((float)b2, INFINITY) * 1.00000095367431640625f; }
    int64_t bad = 0, steps32 = 0, steps8 = 0, rsteps = 0, rbad = 0, rhits = 0;
    std::vector<int32_t> hits;
    for (int64_t k = 0; k < m; k++) {
        const float *q = &Q[3 * k];
        double bb = INFINITY; int32_t bi = -1;
        std::vector<int32_t> want;
        for (int64_t i = 0; i < ix.n; i++) {
            const double ex = (double)P[3 * i] - (double)q[0], ey = (double)P[3 * i + 1] - (double)q[1], ez = (double)P[3 * i + 2] - (double)q[2];
            double e = ex * ex; e = e + ey * ey; e = e + ez * ez;
            if (e < bb) { bb = e; bi = (int32_t)i; }
            if (e <= range * range) want.push_back((int32_t)i);
        }
        for (int G : { 32, 8 }) {
            double best = INFINITY; int32_t idx = -1;
            const int s = ix.n > 0 ? nearest_frontier(ix, q, thr0, &best, &idx, G) : 0;
            (G == 32 ? steps32 : steps8) += s;
            const bool inside = bound <= 0 || bb <= (double)thr0;
            if (ix.n > 0 && (inside ? (idx != bi || best != bb) : (idx != -1 && best != bb))) { if (bad < 5) fprintf(stderr, "nearest %lld (G=%d): got (%d, %.17g) want (%d, %.17g)\n", (long long)k, G, idx, best, bi, bb); bad++; }
        }
        if (ix.n > 0) rsteps += range_frontier(ix, q, range, hits, 32); else hits.clear();
        rhits += (int64_t)hits.size();
        if (hits != want) { if (rbad < 5) fprintf(stderr, "range %lld: %zu hits, want %zu\n", (long long)k, hits.size(), want.size()); rbad++; }
    }
    printf("n=%lld m=%lld nearest: steps/query G=32 %.1f, G=8 %.1f, mismatches=%lld; range r=%.2f: %.1f hits/query, %.1f steps/query, mismatches=%lld\n",
           (long long)ix.n, (long long)m, (double)steps32 / m, (double)steps8 / m, (long long)bad, range, (double)rhits / m, (double)rsteps / m, (long long)rbad);
    return bad || rbad ? 1 : 0;
}
