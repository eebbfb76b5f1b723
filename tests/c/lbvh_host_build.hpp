// lbvh_host_build.hpp -- host-side construction of the planned prefix-split index (test code shared by lbvh_index_check.cpp and
// lbvh_packet_check.cpp): curve order like pc_keygen_kernel + the stable radix sort, records laid out as the CUDA build of the
// next round will write them (see pointcloudtraj_b200/csrc/lbvh.cuh).
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../pointcloudtraj_b200/csrc/lbvh.cuh"

static inline uint32_t lh_spread10(uint32_t v) { v &= 0x3ff; v = (v | (v << 16)) & 0x030000ff; v = (v | (v << 8)) & 0x0300f00f; v = (v | (v << 4)) & 0x030c30c3; v = (v | (v << 2)) & 0x09249249; return v; }
static inline uint32_t lh_hilbert30(uint32_t x, uint32_t y, uint32_t z)
{
    uint32_t X[3] = { x, y, z }, M = 1u << 9, t;
    for (uint32_t Q = M; Q > 1; Q >>= 1) {
        const uint32_t P = Q - 1;
        for (int i = 0; i < 3; i++) { if (X[i] & Q) X[0] ^= P; else { t = (X[0] ^ X[i]) & P; X[0] ^= t; X[i] ^= t; } }
    }
    X[1] ^= X[0]; X[2] ^= X[1]; t = 0;
    for (uint32_t Q = M; Q > 1; Q >>= 1) if (X[2] & Q) t ^= Q - 1;
    X[0] ^= t; X[1] ^= t; X[2] ^= t;
    return (lh_spread10(X[0]) << 2) | (lh_spread10(X[1]) << 1) | lh_spread10(X[2]);
}
static inline std::vector<float> lh_read_f32(const char *path)
{
    FILE *f = fopen(path, "rb"); if (!f) { perror(path); exit(2); }
    fseek(f, 0, SEEK_END); long b = ftell(f); fseek(f, 0, SEEK_SET);
    std::vector<float> v((size_t)b / 4); if (fread(v.data(), 4, v.size(), f) != v.size()) exit(2); fclose(f); return v;
}
static inline float lh_thr_from(double e) { float f = (float)e; if ((double)f < e) f = nextafterf(f, INFINITY); return nextafterf(f * 1.00000095367431640625f, INFINITY); }

struct LhIndex {
    int64_t n = 0, used = 0;
    float lo[3] = { 0, 0, 0 }, inv = 0;                 // quantisation frame of the curve keys
    std::vector<uint32_t> skey;                         // sorted keys
    std::vector<pc_f4> pts, rec;                        // points in curve order (+ padding), records
    uint32_t root = PC_REF_LEAF;

    uint32_t key_of(const float *p) const
    {
        uint32_t c[3];
        for (int a = 0; a < 3; a++) { float v = (p[a] - lo[a]) * inv; v = std::min(std::max(v, 0.0f), 1023.0f); c[a] = (uint32_t)v; }
        return lh_hilbert30(c[0], c[1], c[2]);
    }
    void box(int64_t a, int64_t b, pc_f4 &mn, pc_f4 &mx) const
    {
        mn = pc_f4{ INFINITY, INFINITY, INFINITY, 0.f }; mx = pc_f4{ -INFINITY, -INFINITY, -INFINITY, 0.f };
        for (int64_t i = a; i <= b; i++) { mn.x = std::min(mn.x, pts[(size_t)i].x); mn.y = std::min(mn.y, pts[(size_t)i].y); mn.z = std::min(mn.z, pts[(size_t)i].z);
                                           mx.x = std::max(mx.x, pts[(size_t)i].x); mx.y = std::max(mx.y, pts[(size_t)i].y); mx.z = std::max(mx.z, pts[(size_t)i].z); }
    }
    // the implicit tree of the index at HEAD (aligned groups of consecutive leaves of PC_LBVH_LEAF points, heap numbering,
    // root = node 1) expressed in the same record format, so that the same walks can count its visits
    void build_implicit(const std::vector<float> &P)
    {
        build(P);
        const int64_t n_leaves = (n + PC_LBVH_LEAF - 1) / PC_LBVH_LEAF;
        int64_t Pw = 2; while (Pw < n_leaves) Pw <<= 1;
        rec.assign((size_t)Pw * 4, pc_f4{ INFINITY, INFINITY, INFINITY, 0.f });
        used = 0;
        root = n > 0 ? 1u : PC_REF_LEAF;
        for (int64_t i = 1; i < Pw; i++) {
            // node i at depth d covers leaves [ (i - 2^d) * span, ... + span ), span = Pw >> d
            int d = 0; while (((int64_t)1 << (d + 1)) <= i) d++;
            const int64_t span = Pw >> d, first_leaf = (i - ((int64_t)1 << d)) * span;
            if (first_leaf >= n_leaves) continue;
            used++;
            pc_f4 *r = &rec[(size_t)i * 4];
            for (int c = 0; c < 2; c++) {
                const int64_t cl0 = first_leaf + c * (span / 2), cl1 = std::min(cl0 + span / 2, n_leaves);     // leaves [cl0, cl1)
                pc_f4 &mn = r[2 * c], &mx = r[2 * c + 1];
                mn = pc_f4{ INFINITY, INFINITY, INFINITY, 0.f }; mx = pc_f4{ -INFINITY, -INFINITY, -INFINITY, 0.f };
                uint32_t ref = PC_REF_LEAF;
                if (cl0 < n_leaves) {
                    box(cl0 * PC_LBVH_LEAF, std::min(cl1 * PC_LBVH_LEAF, n) - 1, mn, mx);
                    ref = span / 2 == 1 ? (PC_REF_LEAF | (uint32_t)(cl0 * PC_LBVH_LEAF)) : (uint32_t)(2 * i + c);
                }
                mn.w = pc_u2f(ref); mx.w = pc_u2f(0u);
            }
        }
    }

    void build(const std::vector<float> &P)
    {
        n = (int64_t)P.size() / 3;
        float hi[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX }, ext = 0;
        for (int a = 0; a < 3; a++) lo[a] = FLT_MAX;
        for (int64_t i = 0; i < n; i++) for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], P[3 * i + a]); hi[a] = std::max(hi[a], P[3 * i + a]); }
        for (int a = 0; a < 3; a++) ext = std::max(ext, hi[a] - lo[a]);
        inv = ext > 0 ? 1024.0f * (1.0f - 1e-6f) / ext : 0.0f;
        std::vector<uint32_t> key((size_t)n); std::vector<int32_t> ord((size_t)n);
        for (int64_t i = 0; i < n; i++) { key[(size_t)i] = key_of(&P[3 * i]); ord[(size_t)i] = (int32_t)i; }
        std::stable_sort(ord.begin(), ord.end(), [&](int32_t a, int32_t b) { return key[(size_t)a] < key[(size_t)b]; });
        skey.resize((size_t)n); pts.resize((size_t)n + PC_LBVH_LEAF);
        for (int64_t i = 0; i < n; i++) { const int32_t s = ord[(size_t)i]; skey[(size_t)i] = key[(size_t)s]; pts[(size_t)i] = pc_f4{ P[3 * s], P[3 * s + 1], P[3 * s + 2], pc_u2f((uint32_t)s) }; }
        for (int i = 0; i < PC_LBVH_LEAF && n > 0; i++) pts[(size_t)n + i] = pts[(size_t)n - 1];      // padding: copies of the last point
        rec.assign((size_t)std::max<int64_t>(n - 1, 1) * 4, pc_f4{ INFINITY, INFINITY, INFINITY, 0.f });
        root = PC_REF_LEAF | 0u;                            // n <= PC_LBVH_LEAF: the whole cloud is one leaf
        used = 0;
        if (n <= PC_LBVH_LEAF) return;
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
};
