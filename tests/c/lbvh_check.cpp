// lbvh_check.cpp -- CPU check of pointcloudtraj_b200/csrc/lbvh.cuh (groundwork for the prefix-split tree):
// the per-node range / split computation must describe exactly the tree a recursive top-down construction builds
// ("split where the highest differing bit of the (key, position) string flips"), for 32- and 64-bit keys, with duplicates.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
#include "../../pointcloudtraj_b200/csrc/lbvh.cuh"

template <typename KeyT>
static long check_range(const std::vector<KeyT> &k, int64_t first, int64_t last, const std::vector<int64_t> &F, const std::vector<int64_t> &L,
                        const std::vector<int64_t> &S, int64_t node, std::vector<char> &seen)
{
    // reference split of [first, last]: the first position whose (key, position) string has the highest differing bit set
    const int64_t n = (int64_t)k.size();
    const int common = pc_lbvh_delta(k.data(), n, first, last);
    int64_t lo = first, hi = last;                         // largest g in [first, last) with delta(first, g) > common
    while (lo < hi) { const int64_t mid = (lo + hi + 1) / 2; if (pc_lbvh_delta(k.data(), n, first, mid) > common) lo = mid; else hi = mid - 1; }
    const int64_t g = lo;
    if (F[node] != first || L[node] != last || S[node] != g) {
        fprintf(stderr, "node %lld: got [%lld,%lld] split %lld, want [%lld,%lld] split %lld\n", (long long)node, (long long)F[node], (long long)L[node],
                (long long)S[node], (long long)first, (long long)last, (long long)g);
        exit(1);
    }
    if (seen[node]) { fprintf(stderr, "node %lld reached twice\n", (long long)node); exit(1); }
    seen[node] = 1;
    long count = 1;
    if (g > first) count += check_range(k, first, g, F, L, S, g, seen);
    if (g + 1 < last) count += check_range(k, g + 1, last, F, L, S, g + 1, seen);
    return count;
}

template <typename KeyT>
static void run(int64_t n, int key_bits, unsigned seed)
{
    std::mt19937_64 rng(seed);
    std::vector<KeyT> k((size_t)n);
    for (auto &v : k) v = (KeyT)(rng() & ((key_bits >= 64 ? ~0ull : ((1ull << key_bits) - 1))));
    std::sort(k.begin(), k.end());
    std::vector<int64_t> F((size_t)n - 1), L((size_t)n - 1), S((size_t)n - 1);
    for (int64_t i = 0; i + 1 < n; i++) pc_lbvh_node(k.data(), n, i, &F[(size_t)i], &L[(size_t)i], &S[(size_t)i]);
    std::vector<char> seen((size_t)n - 1, 0);
    const long visited = check_range(k, 0, n - 1, F, L, S, 0, seen);
    if (visited != n - 1) { fprintf(stderr, "n=%lld bits=%d: %ld of %lld inner nodes reachable from the root\n", (long long)n, key_bits, visited, (long long)n - 1); exit(1); }
}

#include <algorithm>
int main()
{
    for (unsigned seed = 1; seed <= 5; seed++) {
        for (int64_t n : { 2, 3, 4, 5, 17, 256, 1000, 20000 }) {
            run<uint32_t>(n, 30, seed);        // 30-bit curve keys
            run<uint32_t>(n, 6, seed);         // heavy duplicates: position bits decide
            run<uint32_t>(n, 0, seed);         // all keys equal
            run<uint64_t>(n, 63, seed);        // 63-bit curve keys
        }
    }
    printf("lbvh ok\n");
    return 0;
}
