/*
 * kd_client.c -- ONE client program, two backends.  It is written against the reference's kd-tree API
 * (Utils/kdtree/include/kdtree/kdtree.h:39-122) the way the planner uses it (kd_create / kd_insert3f with the
 * index in the data pointer / kd_nearest3f + kd_res_item / kd_nearest_range3f + iteration / kd_res_free / kd_free):
 *
 *   gcc kd_client.c -L oracle/_ref -lkdtree_ref                 -> runs on the UNMODIFIED reference library
 *   gcc -DUSE_PCINDEX kd_client.c -I include -lpcindex          -> same source on the GPU index via the
 *                                                                   kd_* -> pckd_* renames of pc_kdtree_compat.h
 *
 * tests/test_kd_compat_gpu.py builds both and compares their output files.
 * usage: kd_client <in.bin> <out.bin>     in : int64 n, int64 m, int64 n_range, double range, float pts[n*3], float q[m*3]
 *                                        out: int64 nn_index[m], double nn_pos[m*3], int64 range_count[n_range],
 *                                             int64 range_items[sum counts] (each list sorted ascending)
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#ifdef USE_PCINDEX
#define PC_KDTREE_COMPAT_RENAME
#include "pc_kdtree_compat.h"
#else
/* prototypes of the reference's public API (the header itself only exists in the reference checkout) */
struct kdtree;
struct kdres;
struct kdtree *kd_create(int k);
void kd_free(struct kdtree *tree);
int kd_insert3f(struct kdtree *tree, float x, float y, float z, void *data);
struct kdres *kd_nearest3f(struct kdtree *tree, float x, float y, float z);
struct kdres *kd_nearest_range3f(struct kdtree *tree, float x, float y, float z, float range);
void kd_res_free(struct kdres *set);
int kd_res_size(struct kdres *set);
int kd_res_end(struct kdres *set);
int kd_res_next(struct kdres *set);
void *kd_res_item(struct kdres *set, double *pos);
void *kd_res_item_data(struct kdres *set);
#endif

static int cmp_i64(const void *a, const void *b)
{
    int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
    return x < y ? -1 : (x > y ? 1 : 0);
}

int main(int argc, char **argv)
{
    if (argc != 3) return 2;
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 3;
    int64_t n, m, nr;
    double range;
    if (fread(&n, 8, 1, f) != 1 || fread(&m, 8, 1, f) != 1 || fread(&nr, 8, 1, f) != 1 || fread(&range, 8, 1, f) != 1) return 4;
    float *pts = malloc((size_t)n * 12 + 4), *q = malloc((size_t)m * 12 + 4);
    if (fread(pts, 12, (size_t)n, f) != (size_t)n || fread(q, 12, (size_t)m, f) != (size_t)m) return 4;
    fclose(f);

    struct kdtree *tree = kd_create(3);
    if (!tree) return 5;
    for (int64_t i = 0; i < n; i++)
        if (kd_insert3f(tree, pts[3 * i], pts[3 * i + 1], pts[3 * i + 2], (void *)(intptr_t)(i + 1))) return 6;

    FILE *o = fopen(argv[2], "wb");
    if (!o) return 7;
    int64_t *nn = malloc((size_t)m * 8 + 8);
    double *pos = malloc((size_t)m * 24 + 8);
    for (int64_t k = 0; k < m; k++) {
        struct kdres *r = kd_nearest3f(tree, q[3 * k], q[3 * k + 1], q[3 * k + 2]);
        if (!r) { nn[k] = -1; pos[3 * k] = pos[3 * k + 1] = pos[3 * k + 2] = 0; continue; }
        if (kd_res_size(r) != 1) return 8;
        nn[k] = (int64_t)(intptr_t)kd_res_item(r, pos + 3 * k) - 1;
        kd_res_free(r);
    }
    fwrite(nn, 8, (size_t)m, o);
    fwrite(pos, 24, (size_t)m, o);
    int64_t *cnt = malloc((size_t)nr * 8 + 8);
    int64_t cap = 1 << 16, used = 0, *items = malloc((size_t)cap * 8);
    for (int64_t k = 0; k < nr; k++) {
        struct kdres *r = kd_nearest_range3f(tree, q[3 * k], q[3 * k + 1], q[3 * k + 2], (float)range);
        if (!r) return 9;
        cnt[k] = kd_res_size(r);
        int64_t first = used;
        while (!kd_res_end(r)) {
            if (used == cap) { cap *= 2; items = realloc(items, (size_t)cap * 8); }
            items[used++] = (int64_t)(intptr_t)kd_res_item_data(r) - 1;
            kd_res_next(r);
        }
        if (used - first != cnt[k]) return 10;
        qsort(items + first, (size_t)(used - first), 8, cmp_i64);
        kd_res_free(r);
    }
    fwrite(cnt, 8, (size_t)nr, o);
    fwrite(items, 8, (size_t)used, o);
    fclose(o);
    kd_free(tree);
    return 0;
}
