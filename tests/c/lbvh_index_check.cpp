// lbvh_index_check.cpp -- CPU check of the prefix-split index planned for the next round: records built
// from pc_lbvh_node / pc_lbvh_children exactly as the future CUDA build will lay them out, walked by pc_lbvh_nearest, and
// compared with an fp64 brute force (lowest index among exact ties).  Also reports the mean number of node visits.
// usage: lbvh_index_check <points.bin> <queries.bin>      (float32 xyz triples)
#include "lbvh_host_build.hpp"

int main(int argc, char **argv)
{
    if (argc != 3) return 2;
    const std::vector<float> P = lh_read_f32(argv[1]), Q = lh_read_f32(argv[2]);
    const int64_t m = (int64_t)Q.size() / 3;
    LhIndex ix;
    ix.build(P);
    const int64_t n = ix.n, used = ix.used;
    const std::vector<uint32_t> &skey = ix.skey;
    const std::vector<pc_f4> &pts = ix.pts, &rec = ix.rec;
    const uint32_t root = ix.root;
    auto box = [&](int64_t a, int64_t b, pc_f4 &mn, pc_f4 &mx) { ix.box(a, b, mn, mx); };
    // The CUDA build will fit the boxes bottom-up: one thread per used inner node boxes its LEAF children, and whichever
    // thread completes a node (second arrival at its counter) merges the node's two child boxes into the slot the node has in
    // its parent's record, and so on upwards.  Emulate that here with the nodes taken in a shuffled order and require the
    // same records as the direct computation above.
    if (n > PC_LBVH_LEAF) {
        std::vector<int64_t> F((size_t)n - 1), L((size_t)n - 1), S((size_t)n - 1), parent((size_t)n - 1, -1);
        std::vector<char> is_used((size_t)n - 1, 0);
        std::vector<int64_t> order;
        for (int64_t i = 0; i + 1 < n; i++) {
            pc_lbvh_node(skey.data(), n, i, &F[(size_t)i], &L[(size_t)i], &S[(size_t)i]);
            if (L[(size_t)i] - F[(size_t)i] + 1 > PC_LBVH_LEAF) { is_used[(size_t)i] = 1; order.push_back(i); }
        }
        for (int64_t i : order) {
            uint32_t r0, c0, r1, c1; pc_lbvh_children(F[(size_t)i], L[(size_t)i], S[(size_t)i], &r0, &c0, &r1, &c1);
            if (!(r0 & PC_REF_LEAF)) parent[(size_t)r0] = i;
            if (!(r1 & PC_REF_LEAF)) parent[(size_t)r1] = i;
        }
        unsigned long long lcg = 12345;
        for (size_t k = order.size(); k > 1; k--) { lcg = lcg * 6364136223846793005ull + 1442695040888963407ull; std::swap(order[k - 1], order[(size_t)((lcg >> 33) % k)]); }
        std::vector<pc_f4> rec2(rec.size(), pc_f4{ INFINITY, INFINITY, INFINITY, 0.f });
        std::vector<int> ready((size_t)n - 1, 0);
        for (int64_t i : order) {
            uint32_t r0, c0, r1, c1; pc_lbvh_children(F[(size_t)i], L[(size_t)i], S[(size_t)i], &r0, &c0, &r1, &c1);
            pc_f4 *r = &rec2[(size_t)i * 4];
            int arrived = 0;
            if (r0 & PC_REF_LEAF) { box(F[(size_t)i], S[(size_t)i], r[0], r[1]); arrived++; }
            if (r1 & PC_REF_LEAF) { box(S[(size_t)i] + 1, L[(size_t)i], r[2], r[3]); arrived++; }
            r[0].w = pc_u2f(r0); r[1].w = pc_u2f(c0); r[2].w = pc_u2f(r1); r[3].w = pc_u2f(c1);      // (.w words never take part in the merges below)
            int64_t node = i;
            int add = arrived;
            while (add > 0) {
                ready[(size_t)node] += add;
                if (ready[(size_t)node] < 2) break;                 // the other child's thread will complete this node
                const int64_t par = parent[(size_t)node];
                if (par < 0) break;                                 // the root is complete
                const pc_f4 *c = &rec2[(size_t)node * 4];
                pc_f4 *pr = &rec2[(size_t)par * 4];
                const int slot = node == S[(size_t)par] ? 0 : 2;    // left child = node `split`, right child = `split + 1`
                pr[slot].x = std::min(c[0].x, c[2].x); pr[slot].y = std::min(c[0].y, c[2].y); pr[slot].z = std::min(c[0].z, c[2].z);
                pr[slot + 1].x = std::max(c[1].x, c[3].x); pr[slot + 1].y = std::max(c[1].y, c[3].y); pr[slot + 1].z = std::max(c[1].z, c[3].z);
                node = par; add = 1;
            }
        }
        for (int64_t i : order) for (int j = 0; j < 4; j++) {
            const pc_f4 a = rec[(size_t)i * 4 + j], b = rec2[(size_t)i * 4 + j];
            if (a.x != b.x || a.y != b.y || a.z != b.z || pc_f2u(a.w) != pc_f2u(b.w)) { fprintf(stderr, "bottom-up fit differs at node %lld slot %d\n", (long long)i, j); return 1; }
        }
    }
    int64_t visits = 0, bad = 0;
    for (int64_t k = 0; k < m; k++) {
        const float qx = Q[3 * k], qy = Q[3 * k + 1], qz = Q[3 * k + 2];
        double best = INFINITY; int32_t idx = -1; float thr = FLT_MAX;
        if (n > 0) pc_lbvh_nearest(rec.data(), pts.data(), root, qx, qy, qz, &best, &idx, &thr, lh_thr_from, &visits);
        double bb = INFINITY; int32_t bi = -1;                // brute force, reference operation order, lowest index
        for (int64_t i = 0; i < n; i++) {
            const double ex = (double)P[3 * i] - (double)qx, ey = (double)P[3 * i + 1] - (double)qy, ez = (double)P[3 * i + 2] - (double)qz;
            double e = ex * ex; e = e + ey * ey; e = e + ez * ez;
            if (e < bb) { bb = e; bi = (int32_t)i; }
        }
        if (bi != idx || bb != best) { if (bad < 5) fprintf(stderr, "query %lld: got (%d, %.17g) want (%d, %.17g)\n", (long long)k, idx, best, bi, bb); bad++; }
    }
    printf("n=%lld m=%lld used_records=%lld visits/query=%.1f mismatches=%lld\n", (long long)n, (long long)m, (long long)used, m ? (double)visits / m : 0.0, (long long)bad);
    return bad ? 1 : 0;
}
