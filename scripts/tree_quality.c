/* tree_quality.c -- CPU experiment (no GPU): how many node visits does the bounded nearest-obstacle search need on
 * different bounding-box trees over the same cloud?  Used to plan the next index layout.
 *
 *   python scripts/tree_quality.py            # writes the cloud / queries, compiles and runs this file
 *
 * Trees (all with leaves of <= LEAF points, tight AABBs, near-first traversal, search bounded at radius R):
 *   morton-implicit   cloud sorted by 30-bit Morton key, complete binary tree over aligned groups of consecutive leaves
 *   hilbert-implicit  the same over a 30-bit Hilbert order (what libpcindex builds)
 *   hilbert-prefix    Hilbert order, but every node is split where the highest differing key bit flips (LBVH / Karras)
 *   median-split      top-down: split the node's points at the median of the longest axis of their box (object median)
 * Output: mean inner visits (one visit = test of the two child boxes of a node) and leaf scans per query.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifdef LEAF_OVERRIDE
#define LEAF LEAF_OVERRIDE
#else
#define LEAF 4
#endif
typedef struct { float lo[3], hi[3]; } box_t;
typedef struct { box_t b; int left, right; int first, count; } node_t;   /* leaf: left < 0, points [first, first + count) */

static float *P;      /* n x 3 */
static int64_t N;
static uint32_t *KEY; /* per point */
static int *ORD;      /* permutation */

static uint32_t spread10(uint32_t v) { v &= 0x3ff; v = (v | (v << 16)) & 0x030000ff; v = (v | (v << 8)) & 0x0300f00f; v = (v | (v << 4)) & 0x030c30c3; v = (v | (v << 2)) & 0x09249249; return v; }
static uint32_t morton30(uint32_t x, uint32_t y, uint32_t z) { return spread10(x) | (spread10(y) << 1) | (spread10(z) << 2); }
static uint32_t hilbert30(uint32_t x, uint32_t y, uint32_t z)
{
    uint32_t X[3] = { x, y, z }, M = 1u << 9, t;
    for (uint32_t Q = M; Q > 1; Q >>= 1) {
        uint32_t Pm = Q - 1;
        for (int i = 0; i < 3; i++) {
            if (X[i] & Q) X[0] ^= Pm;
            else { t = (X[0] ^ X[i]) & Pm; X[0] ^= t; X[i] ^= t; }
        }
    }
    X[1] ^= X[0]; X[2] ^= X[1];
    t = 0;
    for (uint32_t Q = M; Q > 1; Q >>= 1) if (X[2] & Q) t ^= Q - 1;
    X[0] ^= t; X[1] ^= t; X[2] ^= t;
    return (spread10(X[0]) << 2) | (spread10(X[1]) << 1) | spread10(X[2]);
}
static int cmp_key(const void *a, const void *b) { uint32_t ka = KEY[*(const int *)a], kb = KEY[*(const int *)b]; return ka < kb ? -1 : (ka > kb ? 1 : (*(const int *)a - *(const int *)b)); }

static void order_by_curve(int hilbert)
{
    float lo[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, hi[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX }, ext = 0;
    for (int64_t i = 0; i < N; i++) for (int a = 0; a < 3; a++) { if (P[3 * i + a] < lo[a]) lo[a] = P[3 * i + a]; if (P[3 * i + a] > hi[a]) hi[a] = P[3 * i + a]; }
    for (int a = 0; a < 3; a++) if (hi[a] - lo[a] > ext) ext = hi[a] - lo[a];
    float inv = 1024.0f * (1.0f - 1e-6f) / ext;
    for (int64_t i = 0; i < N; i++) {
        uint32_t c[3];
        for (int a = 0; a < 3; a++) { float v = (P[3 * i + a] - lo[a]) * inv; if (v < 0) v = 0; if (v > 1023) v = 1023; c[a] = (uint32_t)v; }
        KEY[i] = hilbert ? hilbert30(c[0], c[1], c[2]) : morton30(c[0], c[1], c[2]);
        ORD[i] = (int)i;
    }
    qsort(ORD, (size_t)N, sizeof(int), cmp_key);
}

static node_t *NODES; static int64_t NN, NCAP;
static int new_node(void) { if (NN == NCAP) { NCAP = NCAP ? NCAP * 2 : 1 << 20; NODES = realloc(NODES, (size_t)NCAP * sizeof(node_t)); } return (int)NN++; }
static box_t box_of(int first, int count)
{
    box_t b; for (int a = 0; a < 3; a++) { b.lo[a] = FLT_MAX; b.hi[a] = -FLT_MAX; }
    for (int i = first; i < first + count; i++) for (int a = 0; a < 3; a++) { float v = P[3 * (int64_t)ORD[i] + a]; if (v < b.lo[a]) b.lo[a] = v; if (v > b.hi[a]) b.hi[a] = v; }
    return b;
}
static box_t box_merge(box_t x, box_t y) { for (int a = 0; a < 3; a++) { if (y.lo[a] < x.lo[a]) x.lo[a] = y.lo[a]; if (y.hi[a] > x.hi[a]) x.hi[a] = y.hi[a]; } return x; }

/* implicit: aligned groups of consecutive leaves, like the GPU index (range of LEAVES [l0, l0 + nl), nl a power of two) */
static int build_implicit(int64_t l0, int64_t nl, int64_t n_leaves)
{
    if (l0 >= n_leaves) return -1;
    int id = new_node();
    if (nl == 1) {
        int first = (int)(l0 * LEAF), count = (int)((N - first) < LEAF ? (N - first) : LEAF);
        NODES[id].left = NODES[id].right = -1; NODES[id].first = first; NODES[id].count = count; NODES[id].b = box_of(first, count);
        return id;
    }
    int l = build_implicit(l0, nl / 2, n_leaves), r = build_implicit(l0 + nl / 2, nl / 2, n_leaves);
    NODES[id].left = l; NODES[id].right = r; NODES[id].first = 0; NODES[id].count = 0;
    NODES[id].b = r >= 0 ? box_merge(NODES[l].b, NODES[r].b) : NODES[l].b;
    return id;
}
/* split where the highest differing key bit flips (falls back to the middle when all keys are equal) */
static int build_prefix(int first, int count)
{
    int id = new_node();
    if (count <= LEAF) { NODES[id].left = NODES[id].right = -1; NODES[id].first = first; NODES[id].count = count; NODES[id].b = box_of(first, count); return id; }
    uint32_t ka = KEY[ORD[first]], kb = KEY[ORD[first + count - 1]];
    int split = first + count / 2;
    if (ka != kb) {
        int bit = 31 - __builtin_clz(ka ^ kb);
        int lo = first, hi = first + count - 1;          /* first index whose key has `bit` set (keys sorted) */
        while (lo < hi) { int mid = (lo + hi) / 2; if ((KEY[ORD[mid]] >> bit) & 1u) hi = mid; else lo = mid + 1; }
        split = lo;
    }
    int l = build_prefix(first, split - first), r = build_prefix(split, first + count - split);
    NODES[id].left = l; NODES[id].right = r; NODES[id].first = 0; NODES[id].count = 0; NODES[id].b = box_merge(NODES[l].b, NODES[r].b);
    return id;
}
static int g_axis;
static int cmp_axis(const void *a, const void *b) { float x = P[3 * (int64_t)*(const int *)a + g_axis], y = P[3 * (int64_t)*(const int *)b + g_axis]; return x < y ? -1 : (x > y ? 1 : 0); }
static int build_median(int first, int count)
{
    int id = new_node();
    box_t b = box_of(first, count);
    if (count <= LEAF) { NODES[id].left = NODES[id].right = -1; NODES[id].first = first; NODES[id].count = count; NODES[id].b = b; return id; }
    int ax = 0; for (int a = 1; a < 3; a++) if (b.hi[a] - b.lo[a] > b.hi[ax] - b.lo[ax]) ax = a;
    g_axis = ax; qsort(ORD + first, (size_t)count, sizeof(int), cmp_axis);
    /* left part: a multiple of LEAF points so that leaves stay full */
    int half = ((count / 2 + LEAF - 1) / LEAF) * LEAF; if (half >= count) half = count / 2;
    int l = build_median(first, half), r = build_median(first + half, count - half);
    NODES[id].left = l; NODES[id].right = r; NODES[id].first = 0; NODES[id].count = 0; NODES[id].b = b;
    return id;
}

static double box_d2(const box_t *b, const float q[3])
{
    double s = 0; for (int a = 0; a < 3; a++) { double d = b->lo[a] - q[a]; double e = q[a] - b->hi[a]; if (e > d) d = e; if (d < 0) d = 0; s += d * d; } return s;
}
static void search(int root, const float q[3], double bound2, int64_t *visits, int64_t *leaves, double *best_out)
{
    int stack[128]; double sd[128]; int sp = 0; double best = bound2; int node = root;
    for (;;) {
        const node_t *n = &NODES[node];
        int descended = 0;
        if (n->left < 0) {
            (*leaves)++;
            for (int i = n->first; i < n->first + n->count; i++) { double s = 0; for (int a = 0; a < 3; a++) { double d = P[3 * (int64_t)ORD[i] + a] - q[a]; s += d * d; } if (s < best) best = s; }
        } else {
            (*visits)++;
            double d0 = box_d2(&NODES[n->left].b, q), d1 = n->right >= 0 ? box_d2(&NODES[n->right].b, q) : INFINITY;
            int cn = d0 <= d1 ? n->left : n->right, cf = d0 <= d1 ? n->right : n->left; double dn = d0 <= d1 ? d0 : d1, df = d0 <= d1 ? d1 : d0;
            if (df <= best) { stack[sp] = cf; sd[sp] = df; sp++; }
            if (dn <= best) { node = cn; descended = 1; }
        }
        if (descended) continue;
        int found = 0;
        while (sp > 0) { sp--; if (sd[sp] <= best) { node = stack[sp]; found = 1; break; } }
        if (!found) break;
    }
    *best_out = best;
}

int main(int argc, char **argv)
{
    if (argc < 4) { fprintf(stderr, "usage: tree_quality points.bin queries.bin radius\n"); return 2; }
    FILE *f = fopen(argv[1], "rb"); fseek(f, 0, SEEK_END); N = ftell(f) / 12; fseek(f, 0, SEEK_SET);
    P = malloc((size_t)N * 12); if (fread(P, 12, (size_t)N, f) != (size_t)N) return 3; fclose(f);
    f = fopen(argv[2], "rb"); fseek(f, 0, SEEK_END); int64_t M = ftell(f) / 12; fseek(f, 0, SEEK_SET);
    float *Q = malloc((size_t)M * 12); if (fread(Q, 12, (size_t)M, f) != (size_t)M) return 3; fclose(f);
    const double R = atof(argv[3]), R2 = R > 0 ? R * R : INFINITY;
    KEY = malloc((size_t)N * 4); ORD = malloc((size_t)N * 4);
    printf("# %lld points, %lld queries, search bounded at %.3f (0 = unbounded); leaves of %d points\n", (long long)N, (long long)M, R, LEAF);
    printf("%-18s %12s %12s %10s\n", "tree", "visits/query", "leaves/query", "nodes");
    double ref_sum = -1;
    for (int t = 0; t < 4; t++) {
        const char *name = t == 0 ? "morton-implicit" : t == 1 ? "hilbert-implicit" : t == 2 ? "hilbert-prefix" : "median-split";
        order_by_curve(t != 0);
        NN = 0;
        int root;
        int64_t n_leaves = (N + LEAF - 1) / LEAF, Pw = 1; while (Pw < n_leaves) Pw <<= 1;
        if (t <= 1) root = build_implicit(0, Pw, n_leaves);
        else if (t == 2) root = build_prefix(0, (int)N);
        else root = build_median(0, (int)N);
        int64_t visits = 0, leaves = 0; double sum = 0;
        for (int64_t k = 0; k < M; k++) { double b; search(root, Q + 3 * k, R2, &visits, &leaves, &b); sum += isfinite(b) ? b : 0; }
        if (ref_sum < 0) ref_sum = sum;
        printf("%-18s %12.1f %12.1f %10lld %s\n", name, (double)visits / M, (double)leaves / M, (long long)NN, fabs(sum - ref_sum) <= 1e-9 * fabs(ref_sum) ? "" : "RESULT MISMATCH");
        fflush(stdout);
    }
    return 0;
}
