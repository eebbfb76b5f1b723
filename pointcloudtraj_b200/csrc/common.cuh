// common.cuh -- shared device helpers for libpcindex (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <float.h>

#ifndef PC_LEAF
#define PC_LEAF 4            // points per leaf (power of two); build with -DPC_LEAF=n to change (scripts/sweep.py)
#endif
#define PC_FULL_MASK 0xffffffffu

// ---- monotone float <-> uint mapping (for atomicMin/atomicMax on floats) -------------------------
__host__ __device__ __forceinline__ uint32_t pc_float_to_ordered(float f)
{
#ifdef __CUDA_ARCH__
    uint32_t b = __float_as_uint(f);
#else
    union { float f; uint32_t u; } c; c.f = f; uint32_t b = c.u;
#endif
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__host__ __device__ __forceinline__ float pc_ordered_to_float(uint32_t k)
{
    uint32_t b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
#ifdef __CUDA_ARCH__
    return __uint_as_float(b);
#else
    union { float f; uint32_t u; } c; c.u = b; return c.f;
#endif
}

// ---- Morton codes --------------------------------------------------------------------------------
// spread the low 10 bits of v so that there are two zero bits between consecutive bits
__device__ __forceinline__ uint32_t pc_spread10(uint32_t v)
{
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

// spread the low 21 bits of v (64-bit Morton)
__device__ __forceinline__ uint64_t pc_spread21(uint64_t v)
{
    v &= 0x1fffffull;
    v = (v | (v << 32)) & 0x001f00000000ffffull;
    v = (v | (v << 16)) & 0x001f0000ff0000ffull;
    v = (v | (v << 8)) & 0x100f00f00f00f00full;
    v = (v | (v << 4)) & 0x10c30c30c30c30c3ull;
    v = (v | (v << 2)) & 0x1249249249249249ull;
    return v;
}

// Quantisation frame of an index: cubic cells of edge 1/inv_cell anchored at the cloud's bbox min.
struct pc_frame {
    float lo[3];
    float inv_cell;   // cells per unit length
    uint32_t max_cell; // 2^bits - 1
};

__device__ __forceinline__ uint32_t pc_cell_coord(float v, float lo, float inv_cell, uint32_t max_cell)
{
    float c = (v - lo) * inv_cell;
    c = fminf(fmaxf(c, 0.0f), (float)max_cell);   // NaN -> 0 via fmaxf
    return (uint32_t)c;
}

__device__ __forceinline__ uint32_t pc_morton30(float x, float y, float z, const pc_frame &f)
{
    uint32_t cx = pc_cell_coord(x, f.lo[0], f.inv_cell, f.max_cell);
    uint32_t cy = pc_cell_coord(y, f.lo[1], f.inv_cell, f.max_cell);
    uint32_t cz = pc_cell_coord(z, f.lo[2], f.inv_cell, f.max_cell);
    return pc_spread10(cx) | (pc_spread10(cy) << 1) | (pc_spread10(cz) << 2);
}

// 30-bit Hilbert index of a 10-bit cell (Skilling's transpose algorithm): consecutive indices are always adjacent
// cells, unlike the Morton curve whose octant jumps put far-apart cells next to each other
__device__ __forceinline__ uint32_t pc_hilbert30_cells(uint32_t x, uint32_t y, uint32_t z)
{
    uint32_t X[3] = { x, y, z };
    const uint32_t M = 1u << 9;
#pragma unroll
    for (uint32_t Q = M; Q > 1; Q >>= 1) {
        const uint32_t P = Q - 1;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            if (X[i] & Q) X[0] ^= P;
            else { const uint32_t t = (X[0] ^ X[i]) & P; X[0] ^= t; X[i] ^= t; }
        }
    }
    X[1] ^= X[0];
    X[2] ^= X[1];
    uint32_t t = 0;
#pragma unroll
    for (uint32_t Q = M; Q > 1; Q >>= 1) if (X[2] & Q) t ^= Q - 1;
    X[0] ^= t; X[1] ^= t; X[2] ^= t;
    return (pc_spread10(X[0]) << 2) | (pc_spread10(X[1]) << 1) | pc_spread10(X[2]);
}

__device__ __forceinline__ uint32_t pc_hilbert30(float x, float y, float z, const pc_frame &f)
{
    return pc_hilbert30_cells(pc_cell_coord(x, f.lo[0], f.inv_cell, f.max_cell), pc_cell_coord(y, f.lo[1], f.inv_cell, f.max_cell),
                              pc_cell_coord(z, f.lo[2], f.inv_cell, f.max_cell));
}

// the same transform for `bits` (11..21) bits per axis, interleaved into a 3 * bits wide key (clouds beyond 4 Mi points)
__device__ __forceinline__ uint64_t pc_hilbert63(float x, float y, float z, const pc_frame &f, int bits)
{
    uint32_t X[3] = { pc_cell_coord(x, f.lo[0], f.inv_cell, f.max_cell), pc_cell_coord(y, f.lo[1], f.inv_cell, f.max_cell),
                      pc_cell_coord(z, f.lo[2], f.inv_cell, f.max_cell) };
    const uint32_t M = 1u << (bits - 1);
    for (uint32_t Q = M; Q > 1; Q >>= 1) {
        const uint32_t P = Q - 1;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            if (X[i] & Q) X[0] ^= P;
            else { const uint32_t t = (X[0] ^ X[i]) & P; X[0] ^= t; X[i] ^= t; }
        }
    }
    X[1] ^= X[0];
    X[2] ^= X[1];
    uint32_t t = 0;
    for (uint32_t Q = M; Q > 1; Q >>= 1) if (X[2] & Q) t ^= Q - 1;
    X[0] ^= t; X[1] ^= t; X[2] ^= t;
    return (pc_spread21(X[0]) << 2) | (pc_spread21(X[1]) << 1) | pc_spread21(X[2]);
}

__device__ __forceinline__ uint64_t pc_morton63(float x, float y, float z, const pc_frame &f)
{
    uint64_t cx = pc_cell_coord(x, f.lo[0], f.inv_cell, f.max_cell);
    uint64_t cy = pc_cell_coord(y, f.lo[1], f.inv_cell, f.max_cell);
    uint64_t cz = pc_cell_coord(z, f.lo[2], f.inv_cell, f.max_cell);
    return pc_spread21(cx) | (pc_spread21(cy) << 1) | (pc_spread21(cz) << 2);
}

__device__ __forceinline__ uint32_t pc_lanemask_lt()
{
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
