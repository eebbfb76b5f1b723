// lbvh_index_check.cpp -- CPU check of the prefix-split index planned for the next round: records built
// from pc_lbvh_node / pc_lbvh_children exactly as the future CUDA build will lay them out, walked by pc_lbvh_nearest, and
// compared with an fp64 brute force (lowest index among exact ties).  Also reports the mean number of node visits.
// usage: lbvh_index_check <points.bin> <queries.bin>      (float32 xyz triples)
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../pointcloudtraj_b200/csrc/lbvh.cuh"

static uint32_t spread10(uint32_t v) { v &= 0x3ff; v = (v | (v << 16)) & 0x030000ff; v = (v | (v << 8)) & 0x0300f00f; v = (v | (v << 4)) & 0x030c30c3; v = (v | (v << 2)) & 0x09249249; return v; }
static uint32_t hilbert30(uint32_t x, uint32_t y, uint32_t z)
{
    uint32_t X[3] = { x, y, z }, M = 1u << 9, t;
    for (uint32_t Q = M; Q > 1; Q >>= 1) {
        const uint32_t P = Q - 1;
        for (int i = 0; i < 3; i++) { if (X[i] & Q) X[0] ^= P; else { t = (X[0] ^ X[i]) & P; X[0] ^= t; X[i] ^= t; } }
    }
    X[1] ^= X[0]; X[2] ^= X[1]; t = 0;
    for (uint32_t Q = M; Q > 1; Q >>= 1) if (X[2] & Q) t ^= Q - 1;
    X[0] ^= t; X[1] ^= t; X[2] ^= t;
    return (spread10(X[0]) << 2) | (spread10(X[1]) << 1) | spread10(X[2]);
}
static std::vector<float> read_f32(const char *path)
{
    FILE *f = fopen(path, "rb"); if (!f) { perror(path); exit(2); }
    fseek(f, 0, SEEK_END); long b = ftell(f); fseek(f, 0, SEEK_SET);
    std::vector<float> v((size_t)b / 4); if (fread(v.data(), 4, v.size(), f) != v.size()) exit(2); fclose(f); return v;
}
static float thr_from(double e) { float f = (float)e; if ((double)f < e) f = nextafterf(f, INFINITY); return nextafterf(f * 1.00000095367431640625f, INFINITY); }

int main(int argc, char **argv)
{
    if (argc != 3) return 2;
    const std::vector<float> P = read_f32(argv[1]), Q = read_f32(argv[2]);
    const int64_t n = (int64_t)P.size() / 3, m = (int64_t)Q.size() / 3;
    // curve order (ties by original index), like pc_keygen_kernel + the stable radix sort
    float lo[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, hi[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX }, ext = 0;
    for (int64_t i = 0; i < n; i++) for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], P[3 * i + a]); hi[a] = std::max(hi[a], P[3 * i + a]); }
    for (int a = 0; a < 3; a++) ext = std::max(ext, hi[a] - lo[a]);
    const float inv = ext > 0 ? 1024.0f * (1.0f - 1e-6f) / ext : 0.0f;
    std::vector<uint32_t> key((size_t)n); std::vector<int32_t> ord((size_t)n);
    for (int64_t i = 0; i < n; i++) {
        uint32_t c[3];
        for (int a = 0; a < 3; a++) { float v = (P[3 * i + a] - lo[a]) * inv; v = std::min(std::max(v, 0.0f), 1023.0f); c[a] = (uint32_t)v; }
        key[(size_t)i] = hilbert30(c[0], c[1], c[2]); ord[(size_t)i] = (int32_t)i;
    }
    std::stable_sort(ord.begin(), ord.end(), [&](int32_t a, int32_t b) { return key[(size_t)a] < key[(size_t)b]; });
    std::vector<uint32_t> skey((size_t)n);
    std::vector<pc_f4> pts((size_t)n + PC_LBVH_LEAF);
    for (int64_t i = 0; i < n; i++) { const int32_t s = ord[(size_t)i]; skey[(size_t)i] = key[(size_t)s]; pts[(size_t)i] = pc_f4{ P[3 * s], P[3 * s + 1], P[3 * s + 2], pc_u2f((uint32_t)s) }; }
    for (int i = 0; i < PC_LBVH_LEAF; i++) pts[(size_t)n + i] = pts[(size_t)n - 1];          // padding: copies of the last point

    // records: one per inner node of the radix tree (those of collapsed nodes stay unused)
    std::vector<pc_f4> rec((size_t)std::max<int64_t>(n - 1, 1) * 4, pc_f4{ INFINITY, INFINITY, INFINITY, 0.f });
    auto box = [&](int64_t a, int64_t b, pc_f4 &mn, pc_f4 &mx) {
        mn = pc_f4{ INFINITY, INFINITY, INFINITY, 0.f }; mx = pc_f4{ -INFINITY, -INFINITY, -INFINITY, 0.f };
        for (int64_t i = a; i <= b; i++) { mn.x = std::min(mn.x, pts[(size_t)i].x); mn.y = std::min(mn.y, pts[(size_t)i].y); mn.z = std::min(mn.z, pts[(size_t)i].z);
                                           mx.x = std::max(mx.x, pts[(size_t)i].x); mx.y = std::max(mx.y, pts[(size_t)i].y); mx.z = std::max(mx.z, pts[(size_t)i].z); }
    };
    uint32_t root = PC_REF_LEAF | 0u;                       // n <= PC_LBVH_LEAF: the whole cloud is one leaf
    int64_t used = 0;
    if (n > PC_LBVH_LEAF) {
        root = 0;
        // (a real build fits the boxes bottom-up; here every used record scans its two ranges -- fine for a CPU check)
        std::vector<int64_t> todo{ 0 };
        while (!todo.empty()) {
            const int64_t i = todo.back(); todo.pop_back(); used++;
            int64_t f, l, s; pc_lbvh_node(skey.data(), n, i, &f, &l, &s);
            uint32_t r0, c0, r1, c1; pc_lbvh_children(f, l, s, &r0, &c0, &r1, &c1);
            pc_f4 *r = &rec[(size_t)i * 4];
            box(f, s, r[0], r[1]); box(s + 1, l, r[2], r[3]);
            r[0].w = pc_u2f(r0); r[1].w = pc_u2f(c0); r[2].w = pc_u2f(r1); r[3].w = pc_u2f(c1);
            if (!(r0 & PC_REF_LEAF)) todo.push_back((int64_t)r0);
            if (!(r1 & PC_REF_LEAF)) todo.push_back((int64_t)r1);
        }
    }
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
        if (n > 0) pc_lbvh_nearest(rec.data(), pts.data(), root, qx, qy, qz, &best, &idx, &thr, thr_from, &visits);
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
