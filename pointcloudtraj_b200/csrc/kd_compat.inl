// kd_compat.inl -- the kd_* call surface of Utils/kdtree over the GPU index (see include/pc_kdtree_compat.h).
// Included by pc_index.cu.  Host-side buffering only; every query goes through the pc_* entry points above.
#include "../../include/pc_kdtree_compat.h"
#include <vector>

struct pckd_tree {
    pc_index *ix = nullptr;
    std::vector<float> xyz;        // 3 floats per inserted point (insertion order = point index)
    std::vector<void *> data;
    void (*destr)(void *) = nullptr;
    bool dirty = false;            // inserts pending: rebuild before the next query
    char err[256] = "";
};

struct pckd_res {
    pckd_tree *tree = nullptr;
    std::vector<int32_t> items;    // point indices
    size_t iter = 0;
};

static bool pckd_is_float(double v) { return (double)(float)v == v || v != v; }

extern "C" struct pckd_tree *pckd_create(int k)
{
    if (k != 3) return nullptr;
    pckd_tree *t = new (std::nothrow) pckd_tree();
    if (!t) return nullptr;
    if (pc_index_create(&t->ix, 0, 0, nullptr) != PC_OK) { delete t; return nullptr; }
    return t;
}

extern "C" void pckd_clear(struct pckd_tree *t)
{
    if (!t) return;
    if (t->destr) for (void *d : t->data) t->destr(d);
    t->xyz.clear(); t->data.clear(); t->dirty = true;
}

extern "C" void pckd_free(struct pckd_tree *t)
{
    if (!t) return;
    pckd_clear(t);
    pc_index_destroy(t->ix);
    delete t;
}

extern "C" void pckd_data_destructor(struct pckd_tree *t, void (*destr)(void *)) { if (t) t->destr = destr; }

extern "C" int pckd_insert3f(struct pckd_tree *t, float x, float y, float z, void *data)
{
    if (!t) return -1;
    try {
        t->xyz.push_back(x); t->xyz.push_back(y); t->xyz.push_back(z);
        t->data.push_back(data);
    } catch (...) { return -1; }
    t->dirty = true;
    return 0;
}

extern "C" int pckd_insert3(struct pckd_tree *t, double x, double y, double z, void *data)
{
    if (!t) return -1;
    if (!pckd_is_float(x) || !pckd_is_float(y) || !pckd_is_float(z)) {
        snprintf(t->err, sizeof t->err, "pckd_insert: coordinate is not representable in float32");
        return -1;
    }
    return pckd_insert3f(t, (float)x, (float)y, (float)z, data);
}

extern "C" int pckd_insert(struct pckd_tree *t, const double *pos, void *data) { return pos ? pckd_insert3(t, pos[0], pos[1], pos[2], data) : -1; }
extern "C" int pckd_insertf(struct pckd_tree *t, const float *pos, void *data) { return pos ? pckd_insert3f(t, pos[0], pos[1], pos[2], data) : -1; }

static int pckd_sync(pckd_tree *t)
{
    if (!t->dirty) return PC_OK;
    int rc = pc_index_build(t->ix, t->xyz.data(), (int64_t)t->data.size(), 3, PC_HOST);
    if (rc != PC_OK) { snprintf(t->err, sizeof t->err, "%s", pc_last_error(t->ix)); return rc; }
    t->dirty = false;
    return PC_OK;
}

extern "C" struct pc_index *pckd_index(struct pckd_tree *t) { return (t && pckd_sync(t) == PC_OK) ? t->ix : nullptr; }
extern "C" int64_t pckd_size(struct pckd_tree *t) { return t ? (int64_t)t->data.size() : 0; }
extern "C" const char *pckd_last_error(struct pckd_tree *t) { return t ? t->err : ""; }

extern "C" struct pckd_res *pckd_nearest3f(struct pckd_tree *t, float x, float y, float z)
{
    if (!t || t->data.empty()) return nullptr;                 // kd_nearest: NULL on an empty tree (kdtree.c:412-413)
    if (pckd_sync(t) != PC_OK) return nullptr;
    const float q[3] = { x, y, z };
    int32_t idx = -1;
    if (pc_nearest_batch(t->ix, q, 1, 3, PC_HOST, PC_QUERY_UNSORTED, &idx, nullptr) != PC_OK || idx < 0) return nullptr;
    pckd_res *r = new (std::nothrow) pckd_res();
    if (!r) return nullptr;
    r->tree = t;
    try { r->items.push_back(idx); } catch (...) { delete r; return nullptr; }
    return r;
}

extern "C" struct pckd_res *pckd_nearest3(struct pckd_tree *t, double x, double y, double z)
{
    if (!t) return nullptr;
    if (!pckd_is_float(x) || !pckd_is_float(y) || !pckd_is_float(z)) {
        snprintf(t->err, sizeof t->err, "pckd_nearest: query coordinate is not representable in float32");
        return nullptr;
    }
    return pckd_nearest3f(t, (float)x, (float)y, (float)z);
}

extern "C" struct pckd_res *pckd_nearest(struct pckd_tree *t, const double *pos) { return pos ? pckd_nearest3(t, pos[0], pos[1], pos[2]) : nullptr; }
extern "C" struct pckd_res *pckd_nearestf(struct pckd_tree *t, const float *pos) { return pos ? pckd_nearest3f(t, pos[0], pos[1], pos[2]) : nullptr; }

static struct pckd_res *pckd_range_impl(struct pckd_tree *t, float x, float y, float z, double range)
{
    if (!t) return nullptr;
    pckd_res *r = new (std::nothrow) pckd_res();
    if (!r) return nullptr;
    r->tree = t;
    if (t->data.empty()) return r;                             // kd_nearest_range: valid empty set (kdtree.c:537-559)
    if (pckd_sync(t) != PC_OK) { delete r; return nullptr; }
    const float q[3] = { x, y, z };
    int64_t off[2] = { 0, 0 };
    if (pc_range_batch(t->ix, q, 1, 3, PC_HOST, &range, 1, off, nullptr, 0) != PC_OK) { delete r; return nullptr; }
    if (off[1] > 0) {
        try { r->items.resize((size_t)off[1]); } catch (...) { delete r; return nullptr; }
        if (pc_range_batch(t->ix, q, 1, 3, PC_HOST, &range, 1, off, r->items.data(), off[1]) != PC_OK) { delete r; return nullptr; }
    }
    return r;
}

extern "C" struct pckd_res *pckd_nearest_range3f(struct pckd_tree *t, float x, float y, float z, float range)
{
    return pckd_range_impl(t, x, y, z, (double)range);
}

extern "C" struct pckd_res *pckd_nearest_range3(struct pckd_tree *t, double x, double y, double z, double range)
{
    if (!t) return nullptr;
    if (!pckd_is_float(x) || !pckd_is_float(y) || !pckd_is_float(z)) {
        snprintf(t->err, sizeof t->err, "pckd_nearest_range: query coordinate is not representable in float32");
        return nullptr;
    }
    return pckd_range_impl(t, (float)x, (float)y, (float)z, range);
}

extern "C" struct pckd_res *pckd_nearest_range(struct pckd_tree *t, const double *pos, double range)
{
    return pos ? pckd_nearest_range3(t, pos[0], pos[1], pos[2], range) : nullptr;
}

extern "C" struct pckd_res *pckd_nearest_rangef(struct pckd_tree *t, const float *pos, float range)
{
    return pos ? pckd_nearest_range3f(t, pos[0], pos[1], pos[2], range) : nullptr;
}

extern "C" void pckd_res_free(struct pckd_res *r) { delete r; }
extern "C" int pckd_res_size(struct pckd_res *r) { return r ? (int)r->items.size() : 0; }
extern "C" void pckd_res_rewind(struct pckd_res *r) { if (r) r->iter = 0; }
extern "C" int pckd_res_end(struct pckd_res *r) { return !r || r->iter >= r->items.size(); }
extern "C" int pckd_res_next(struct pckd_res *r)
{
    if (!r || r->iter >= r->items.size()) return 0;
    r->iter++;
    return r->iter < r->items.size();
}

extern "C" void *pckd_res_item3(struct pckd_res *r, double *x, double *y, double *z)
{
    if (!r || r->iter >= r->items.size()) return nullptr;
    const int32_t i = r->items[r->iter];
    const float *p = &r->tree->xyz[3 * (size_t)i];
    if (x) *x = p[0];
    if (y) *y = p[1];
    if (z) *z = p[2];
    return r->tree->data[(size_t)i];
}

extern "C" void *pckd_res_item(struct pckd_res *r, double *pos)
{
    return pos ? pckd_res_item3(r, pos, pos + 1, pos + 2) : pckd_res_item3(r, nullptr, nullptr, nullptr);
}

extern "C" void *pckd_res_item3f(struct pckd_res *r, float *x, float *y, float *z)
{
    double a, b, c;
    void *d = pckd_res_item3(r, &a, &b, &c);
    if (r && r->iter < r->items.size()) {
        if (x) *x = (float)a;
        if (y) *y = (float)b;
        if (z) *z = (float)c;
    }
    return d;
}

extern "C" void *pckd_res_itemf(struct pckd_res *r, float *pos)
{
    return pos ? pckd_res_item3f(r, pos, pos + 1, pos + 2) : pckd_res_item3f(r, nullptr, nullptr, nullptr);
}

extern "C" void *pckd_res_item_data(struct pckd_res *r) { return pckd_res_item(r, nullptr); }

extern "C" int pckd_nearest_batchf(struct pckd_tree *t, const float *q_xyz, int64_t m, int64_t q_stride,
                                   int32_t *out_index, void **out_data, float *out_d2)
{
    if (!t || m < 0 || (m > 0 && (!q_xyz || !out_index))) return -1;
    if (pckd_sync(t) != PC_OK) return -1;
    if (pc_nearest_batch(t->ix, q_xyz, m, q_stride, PC_HOST, PC_QUERY_AUTO, out_index, out_d2) != PC_OK) {
        snprintf(t->err, sizeof t->err, "%s", pc_last_error(t->ix));
        return -1;
    }
    if (out_data)
        for (int64_t k = 0; k < m; k++) out_data[k] = out_index[k] >= 0 ? t->data[(size_t)out_index[k]] : nullptr;
    return 0;
}
