// grid_kernels.cuh -- EXPERIMENT (off unless PC_GRID=1 is set when the index is created): the index design the north star
// names -- a uniform voxel grid with cell offsets over the cloud sorted by cell, and a ring search -- built next to the tree so
// that the two can be timed on the same batches (scripts/grid_ab.py, profiles/r2_grid_vs_tree.txt).  Used for BOUNDED radius
// batches only: without a bound the ring search has no stopping radius in free space.
//
//   pc_grid_key_kernel    linear cell id of every point (cells of edge h in the cloud's bounding box, x fastest)
//   (radix sort of (cell, point) pairs: radix_sort.cuh)
//   pc_grid_cells_kernel  cell_start[c] = first sorted position whose cell id is >= c (binary search per cell) -- the offset table
//   pc_grid_gather_kernel points in cell order, float4 (x, y, z, original index)
//   pc_radius_grid_kernel one thread per query of the curve-ordered batch (a warp's queries share their cells, so the offset
//                         and point loads of a warp coalesce); rings of cells of growing Chebyshev distance around the
//                         query's cell; a cell is opened when its box is within the current bound; the search stops after
//                         ring r once (r h)^2 exceeds the bound -- every unvisited cell is at least r h away.
// Same exactness rule as the tree kernels (fp32 filter against the inflated bound, fp64 re-rank by (d2, index)), so the
// results are bit-identical with theirs.
#pragma once
#include "query_kernels.cuh"

struct pc_grid {
    const uint32_t *__restrict__ cell_start;   // cells + 1 entries
    const float4 *__restrict__ points;         // cell order
    float lo[3];
    float h, inv_h, eps;                       // eps: bound on the rounding of a point's position relative to its cell's box
    int n[3];                                  // cells per axis
};

__device__ __forceinline__ int pc_grid_coord(float v, float lo, float inv_h, int n)
{
    float c = (v - lo) * inv_h;
    c = fminf(fmaxf(c, 0.0f), (float)(n - 1));      // NaN -> 0
    return (int)c;
}

__global__ void __launch_bounds__(256)
pc_grid_key_kernel(const float *__restrict__ xyz, int64_t n, int stride, pc_grid G, uint32_t *__restrict__ keys, uint32_t *__restrict__ vals)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *p = xyz + i * stride;
    const int cx = pc_grid_coord(p[0], G.lo[0], G.inv_h, G.n[0]), cy = pc_grid_coord(p[1], G.lo[1], G.inv_h, G.n[1]),
              cz = pc_grid_coord(p[2], G.lo[2], G.inv_h, G.n[2]);
    keys[i] = (uint32_t)((cz * G.n[1] + cy) * G.n[0] + cx);
    vals[i] = (uint32_t)i;
}

__global__ void __launch_bounds__(256)
pc_grid_cells_kernel(const uint32_t *__restrict__ sorted_keys, int64_t n, int64_t cells, uint32_t *__restrict__ cell_start)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c > cells) return;
    int64_t lo = 0, hi = n;                          // first position with key >= c
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if ((int64_t)sorted_keys[mid] < c) lo = mid + 1; else hi = mid;
    }
    cell_start[c] = (uint32_t)lo;
}

__global__ void __launch_bounds__(256)
pc_grid_gather_kernel(const float *__restrict__ xyz, int stride, const uint32_t *__restrict__ order, int64_t n, float4 *__restrict__ points)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t src = order[i];
    const float *p = xyz + (int64_t)src * stride;
    points[i] = make_float4(p[0], p[1], p[2], __uint_as_float(src));
}

__global__ void __launch_bounds__(PC_QUERY_THREADS)
pc_radius_grid_kernel(pc_grid G, pc_radius_dev R, const float *__restrict__ q, int64_t m, int qstride,
                      const uint32_t *__restrict__ perm, const float4 *__restrict__ ordered, const unsigned long long *__restrict__ m_eff,
                      int32_t *__restrict__ out_idx, float *__restrict__ out_f)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m || (m_eff && t >= (int64_t)*m_eff)) return;
    uint32_t k;
    float qx, qy, qz;
    if (ordered) { const float4 v = __ldg(ordered + t); qx = v.x; qy = v.y; qz = v.z; k = __float_as_uint(v.w); }
    else { k = perm ? perm[t] : (uint32_t)t; const float *qq = q + (size_t)k * qstride; qx = qq[0]; qy = qq[1]; qz = qq[2]; }
    if (!m_eff && pc_radius_early_out((double)qx, (double)qy, (double)qz, R)) { pc_write_trivial<PC_KIND_RADIUS>(R, k, out_idx, out_f); return; }
    pc_best b; b.d2 = INFINITY; b.idx = -1; b.thr = R.bound_thr;
    const int cx = pc_grid_coord(qx, G.lo[0], G.inv_h, G.n[0]), cy = pc_grid_coord(qy, G.lo[1], G.inv_h, G.n[1]),
              cz = pc_grid_coord(qz, G.lo[2], G.inv_h, G.n[2]);
    const int r_max = max(G.n[0], max(G.n[1], G.n[2]));
    for (int r = 0; r <= r_max; r++) {
        // every cell outside rings 0 .. r-1 is at least (r - 1) h away from any position inside the query's cell -- and from a
        // query outside the grid, whose cell is the nearest border cell -- so once that exceeds the bound nothing is left
        if (r >= 2) { const float g = (float)(r - 1) * G.h - G.eps; if (g * g > b.thr) break; }
        for (int dz = -r; dz <= r; dz++) {
            const int z = cz + dz;
            if (z < 0 || z >= G.n[2]) continue;
            const float bz = fmaxf(fmaxf(G.lo[2] + (float)z * G.h - qz, qz - (G.lo[2] + (float)(z + 1) * G.h)) - G.eps, 0.0f);
            for (int dy = -r; dy <= r; dy++) {
                const int y = cy + dy;
                if (y < 0 || y >= G.n[1]) continue;
                const float by = fmaxf(fmaxf(G.lo[1] + (float)y * G.h - qy, qy - (G.lo[1] + (float)(y + 1) * G.h)) - G.eps, 0.0f);
                const float byz = fmaf(by, by, bz * bz);                     // lower bound of the row's distance
                if (byz > b.thr) continue;
                const bool shell_row = dz == -r || dz == r || dy == -r || dy == r;      // whole row belongs to ring r
                const int step = shell_row ? 1 : 2 * r;                                  // otherwise only its two end cells
                const uint32_t *row = G.cell_start + ((int64_t)z * G.n[1] + y) * G.n[0];
                for (int dx = -r; dx <= r; dx += (step > 0 ? step : 1)) {
                    const int x = cx + dx;
                    if (x < 0 || x >= G.n[0]) continue;
                    const uint32_t s = __ldg(row + x), e = __ldg(row + x + 1);
                    if (s == e) continue;
                    const float bx = fmaxf(fmaxf(G.lo[0] + (float)x * G.h - qx, qx - (G.lo[0] + (float)(x + 1) * G.h)) - G.eps, 0.0f);
                    if (fmaf(bx, bx, byz) > b.thr) continue;
                    for (uint32_t i = s; i < e; i++) {
                        const float4 p = __ldg(G.points + i);
                        const float ddx = p.x - qx, ddy = p.y - qy, ddz = p.z - qz;
                        pc_consider(p, fmaf(ddz, ddz, fmaf(ddy, ddy, ddx * ddx)), qx, qy, qz, b);
                    }
                }
            }
        }
    }
    pc_write_result<PC_KIND_RADIUS>(R, b, k, out_idx, out_f);
}
