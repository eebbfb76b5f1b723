// lbvh_packet_check.cpp -- CPU emulation of the 64-query packet walk (pc_packet2_traverse of query_kernels.cuh) over the
// prefix-split index planned for the next round: the control flow a warp would execute (shared stack, "any query wants it"
// tests, vote of the first 32 queries on the near child, leaf and inner children mixed under one node), run scalar.
// Checks exactness against an fp64 brute force and reports visits per packet, to compare with the implicit tree at HEAD
// (123 visits per packet on the bench workload, profiles/r1_hot_block_breakdown.txt).
// usage: lbvh_packet_check <points.bin> <queries.bin> <bound> [implicit] [nocheck]
//        bound = search radius bound, 0 = unbounded; "implicit" walks the implicit tree of the index at HEAD instead (same
//        record format); "nocheck" skips the O(n m) brute force (for the large, dense batches used to predict visit counts)
#include <string>
#include "lbvh_host_build.hpp"

struct Packet {
    float q[64][3]; double best[64]; int32_t idx[64]; float thr[64]; int n;
};

static void scan_leaf(const LhIndex &ix, uint32_t ref, Packet &p)
{
    const pc_f4 *pt = ix.pts.data() + (ref & 0x7fffffffu);
    for (int k = 0; k < p.n; k++) {
        for (int i = 0; i < PC_LBVH_LEAF; i++) {
            const float dx = pt[i].x - p.q[k][0], dy = pt[i].y - p.q[k][1], dz = pt[i].z - p.q[k][2];
            const float d = dx * dx + dy * dy + dz * dz;
            if (d <= p.thr[k]) {
                const double ex = (double)pt[i].x - (double)p.q[k][0], ey = (double)pt[i].y - (double)p.q[k][1], ez = (double)pt[i].z - (double)p.q[k][2];
                double e = ex * ex; e = e + ey * ey; e = e + ez * ez;
                const int32_t id = (int32_t)pc_f2u(pt[i].w);
                if (e < p.best[k] || (e == p.best[k] && (uint32_t)id < (uint32_t)p.idx[k])) { p.best[k] = e; p.idx[k] = id; p.thr[k] = std::min(p.thr[k], lh_thr_from(e)); }
            }
        }
    }
}

static void walk(const LhIndex &ix, Packet &p, int64_t &visits, int64_t &leaf_scans, int64_t &stale)
{
    uint32_t stack[96]; int sp = 0;
    uint32_t ref = ix.root;
    if (ref & PC_REF_LEAF) { scan_leaf(ix, ref, p); leaf_scans++; return; }
    for (;;) {
        visits++;
        const pc_f4 *r = ix.rec.data() + 4 * (int64_t)ref;
        float d0[64], d1[64];
        bool w0 = false, w1 = false;
        for (int k = 0; k < p.n; k++) {
            d0[k] = pc_lbvh_box_d2(r[0], r[1], p.q[k][0], p.q[k][1], p.q[k][2]);
            d1[k] = pc_lbvh_box_d2(r[2], r[3], p.q[k][0], p.q[k][1], p.q[k][2]);
            w0 = w0 || d0[k] <= p.thr[k]; w1 = w1 || d1[k] <= p.thr[k];
        }
        uint32_t next = 0xffffffffu;
        if (w0 || w1) {
            const bool both = w0 && w1;
            bool first0 = !w1;
            if (both) {                                   // vote of the lanes' first queries (slots 0..31) that are interested
                int interested = 0, pref0 = 0;
                for (int k = 0; k < std::min(p.n, 32); k++)
                    if (d0[k] <= p.thr[k] || d1[k] <= p.thr[k]) { interested++; if (d0[k] <= d1[k]) pref0++; }
                first0 = 2 * pref0 >= interested;
            }
            const uint32_t r0 = pc_f2u(r[0].w), r1 = pc_f2u(r[2].w);
            const uint32_t rn = first0 ? r0 : r1, rf = first0 ? r1 : r0;
            const float *dfar = first0 ? d1 : d0;
            if (rn & PC_REF_LEAF) { scan_leaf(ix, rn, p); leaf_scans++; } else next = rn;
            if (both) {
                if (rf & PC_REF_LEAF) {
                    bool still = false;                   // re-test against the bounds the near child may have tightened
                    for (int k = 0; k < p.n; k++) still = still || dfar[k] <= p.thr[k];
                    if (still) { scan_leaf(ix, rf, p); leaf_scans++; }
                } else if (next != 0xffffffffu) stack[sp++] = rf;
                else next = rf;
            }
        } else stale++;
        if (next != 0xffffffffu) { ref = next; continue; }
        if (sp == 0) break;
        ref = stack[--sp];
    }
}

int main(int argc, char **argv)
{
    if (argc < 4) return 2;
    bool implicit = false, check = true;
    for (int a = 4; a < argc; a++) { if (std::string(argv[a]) == "implicit") implicit = true; if (std::string(argv[a]) == "nocheck") check = false; }
    const std::vector<float> P = lh_read_f32(argv[1]);
    std::vector<float> Q = lh_read_f32(argv[2]);
    const double bound = atof(argv[3]);
    const int64_t m = (int64_t)Q.size() / 3;
    LhIndex ix;
    if (implicit) ix.build_implicit(P); else ix.build(P);
    // order the batch like the ordering pass: top 24 bits of the query's Hilbert key in the index's frame
    std::vector<int64_t> ord((size_t)m); std::vector<uint32_t> key((size_t)m);
    for (int64_t k = 0; k < m; k++) { ord[(size_t)k] = k; key[(size_t)k] = ix.key_of(&Q[3 * k]) >> 6; }
    std::stable_sort(ord.begin(), ord.end(), [&](int64_t a, int64_t b) { return key[(size_t)a] < key[(size_t)b]; });
    float thr0 = FLT_MAX;
    if (bound > 0) { const double b2 = bound * bound * (1.0 + 1e-6); thr0 = nextafterf((float)b2, INFINITY) * 1.00000095367431640625f; }
    int64_t visits = 0, leaf_scans = 0, stale = 0, packets = 0, bad = 0;
    for (int64_t base = 0; base < m; base += 64, packets++) {
        Packet p; p.n = (int)std::min<int64_t>(64, m - base);
        // lane l holds the queries of slots l and l + 32: the vote uses slots 0..31
        for (int k = 0; k < p.n; k++) { const int64_t s = ord[(size_t)(base + k)]; for (int a = 0; a < 3; a++) p.q[k][a] = Q[3 * s + a]; p.best[k] = INFINITY; p.idx[k] = -1; p.thr[k] = thr0; }
        if (ix.n > 0) walk(ix, p, visits, leaf_scans, stale);
        for (int k = 0; check && k < p.n; k++) {
            const int64_t s = ord[(size_t)(base + k)];
            double bb = INFINITY; int32_t bi = -1;
            for (int64_t i = 0; i < ix.n; i++) {
                const double ex = (double)P[3 * i] - (double)Q[3 * s], ey = (double)P[3 * i + 1] - (double)Q[3 * s + 1], ez = (double)P[3 * i + 2] - (double)Q[3 * s + 2];
                double e = ex * ex; e = e + ey * ey; e = e + ez * ez;
                if (e < bb) { bb = e; bi = (int32_t)i; }
            }
            const bool inside = bound <= 0 || bb <= (double)thr0;      // beyond the bound nothing has to be found
            if (inside ? (bi != p.idx[k] || bb != p.best[k]) : (p.idx[k] != -1 && p.best[k] != bb)) { if (bad < 5) fprintf(stderr, "query %lld: got (%d, %.17g) want (%d, %.17g)\n", (long long)s, p.idx[k], p.best[k], bi, bb); bad++; }
        }
    }
    printf("%s tree: n=%lld m=%lld packets=%lld visits/packet=%.1f leaf_scans/packet=%.1f stale=%.1f%% mismatches=%lld\n", implicit ? "implicit" : "prefix-split", (long long)ix.n, (long long)m, (long long)packets,
           (double)visits / packets, (double)leaf_scans / packets, 100.0 * stale / std::max<int64_t>(visits, 1), (long long)bad);
    return bad ? 1 : 0;
}
