/*
 * ref_harness.c -- TEST INFRASTRUCTURE ONLY.  Batch driver around the UNMODIFIED reference
 * kd-tree.  oracle/Makefile compiles this file together with the reference's own
 * Utils/kdtree/src/kdtree.c (read in place from /root/reference; no reference source is copied
 * into this repo) into oracle/_ref/libkdtree_ref.so.  Only the reference's public C API
 * (Utils/kdtree/include/kdtree/kdtree.h:39-122) is used.
 *
 * The point index travels in the node's `data` pointer; only the double `...3` entry points
 * are called from threads (the `...f` ones use static buffers, kdtree.c:213,461,563).
 * kd_res_item3 is avoided (it tests *x instead of x and returns 0, kdtree.c:666-684).
 */
#include <stdint.h>
#include <stdlib.h>
#include <math.h>
#include "kdtree/kdtree.h"
#ifdef _OPENMP
#include <omp.h>
#endif

void *refh_create(void) { return kd_create(3); }
void refh_free(void *tree) { kd_free((struct kdtree *)tree); }
void refh_clear(void *tree) { kd_clear((struct kdtree *)tree); }

/* kd_clear + n x kd_insert3(x, y, z, (void*)i), in `order` (NULL: 0..n-1) */
int refh_build(void *tree, const float *xyz, int64_t n, int64_t stride, const int64_t *order)
{
    struct kdtree *t = (struct kdtree *)tree;
    kd_clear(t);
    for (int64_t k = 0; k < n; k++) {
        int64_t i = order ? order[k] : k;
        const float *p = xyz + i * stride;
        if (kd_insert3(t, (double)p[0], (double)p[1], (double)p[2], (void *)(intptr_t)i)) return -1;
    }
    return 0;
}

static int threads_or_all(int nthreads)
{
#ifdef _OPENMP
    return nthreads > 0 ? nthreads : omp_get_max_threads();
#else
    (void)nthreads;
    return 1;
#endif
}

int refh_max_threads(void) { return threads_or_all(0); }

int refh_nearest_batch(void *tree, const float *q, int64_t m, int64_t stride,
                       int64_t *idx, double *d2, int nthreads)
{
    struct kdtree *t = (struct kdtree *)tree;
    nthreads = threads_or_all(nthreads);
#pragma omp parallel for schedule(dynamic, 256) num_threads(nthreads)
    for (int64_t k = 0; k < m; k++) {
        const float *f = q + k * stride;
        double x = f[0], y = f[1], z = f[2];
        struct kdres *r = kd_nearest3(t, x, y, z);
        if (!r) {
            if (idx) idx[k] = -1;
            if (d2) d2[k] = INFINITY;
            continue;
        }
        double pos[3];
        void *data = kd_res_item(r, pos);
        if (idx) idx[k] = (int64_t)(intptr_t)data;
        if (d2) {
            double s = 0.0;
            s += (pos[0] - x) * (pos[0] - x);
            s += (pos[1] - y) * (pos[1] - y);
            s += (pos[2] - z) * (pos[2] - z);
            d2[k] = s;
        }
        kd_res_free(r);
    }
    return 0;
}

/* one range query; writes min(count, cap) indices in the result set's iteration order */
int64_t refh_range3(void *tree, double x, double y, double z, double range, int64_t *out, int64_t cap)
{
    struct kdres *r = kd_nearest_range3((struct kdtree *)tree, x, y, z, range);
    if (!r) return -1;
    int64_t n = kd_res_size(r), w = 0;
    if (out) {
        for (kd_res_rewind(r); !kd_res_end(r) && w < cap; kd_res_next(r))
            out[w++] = (int64_t)(intptr_t)kd_res_item_data(r);
    }
    kd_res_free(r);
    return n;
}

int refh_range_count_batch(void *tree, const float *q, int64_t m, int64_t stride,
                           const double *range, int range_is_scalar, int64_t *out_count, int nthreads)
{
    nthreads = threads_or_all(nthreads);
#pragma omp parallel for schedule(dynamic, 64) num_threads(nthreads)
    for (int64_t k = 0; k < m; k++) {
        const float *f = q + k * stride;
        out_count[k] = refh_range3(tree, f[0], f[1], f[2], range_is_scalar ? range[0] : range[k], NULL, 0);
    }
    return 0;
}

int refh_range_fill_batch(void *tree, const float *q, int64_t m, int64_t stride,
                          const double *range, int range_is_scalar, const int64_t *offsets, int64_t *out_idx,
                          int nthreads)
{
    nthreads = threads_or_all(nthreads);
#pragma omp parallel for schedule(dynamic, 64) num_threads(nthreads)
    for (int64_t k = 0; k < m; k++) {
        const float *f = q + k * stride;
        refh_range3(tree, f[0], f[1], f[2], range_is_scalar ? range[0] : range[k],
                    out_idx + offsets[k], offsets[k + 1] - offsets[k]);
    }
    return 0;
}

/* radiusSearch epilogue on top of the real kd_nearest3 (Planner/src/corridor_finder.cpp:113-133);
 * used as the CPU baseline for the headline metric. params = {search_margin, max_radius, sample_range, sx, sy, sz} */
int refh_radius_batch(void *tree, const float *q, int64_t m, int64_t stride, const double *params,
                      double *out_radius, int nthreads)
{
    struct kdtree *t = (struct kdtree *)tree;
    const double margin = params[0], rmax = params[1], srange = params[2];
    const double sx = params[3], sy = params[4], sz = params[5];
    nthreads = threads_or_all(nthreads);
#pragma omp parallel for schedule(dynamic, 256) num_threads(nthreads)
    for (int64_t k = 0; k < m; k++) {
        const float *f = q + k * stride;
        double x = f[0], y = f[1], z = f[2];
        if (srange >= 0.0) {
            double dis = sqrt((x - sx) * (x - sx) + (y - sy) * (y - sy) + (z - sz) * (z - sz));
            if (dis > srange + rmax) { out_radius[k] = rmax - margin; continue; }
        }
        struct kdres *r = kd_nearest3(t, x, y, z);
        if (!r) { out_radius[k] = rmax - margin; continue; }
        double pos[3];
        kd_res_item(r, pos);
        double s = 0.0;
        s += (pos[0] - x) * (pos[0] - x);
        s += (pos[1] - y) * (pos[1] - y);
        s += (pos[2] - z) * (pos[2] - z);
        kd_res_free(r);
        double radius = sqrt(s) - margin;
        out_radius[k] = radius < rmax ? radius : rmax;
    }
    return 0;
}
