// corridor_client.cpp -- exercises include/pc_corridor.hpp the way sim_planning_demo / corridor_finder use the cloud:
// setParam, setPt, setInput (rebuild per frame), radiusSearch (single + batch), checkTrajPtCol.
// usage: corridor_client <in.bin> <out.bin>   in: int64 n, int64 m, double params[4] (safety, search, max_radius, range),
//                                              double start[3], float pts[n*4] (PointXYZ layout), double q[m*3]
//                                             out: double radius_single[m], float radius_batch[m], uint8 col[m],
//                                                  int64 first_collision (col_rad 0.3), float nearest_dist[m]
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "pc_corridor.hpp"

int main(int argc, char **argv)
{
    if (argc != 3) return 2;
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 3;
    int64_t n, m;
    double prm[4], start[3];
    if (fread(&n, 8, 1, f) != 1 || fread(&m, 8, 1, f) != 1 || fread(prm, 8, 4, f) != 4 || fread(start, 8, 3, f) != 3) return 4;
    std::vector<float> pts((size_t)n * 4);
    std::vector<double> q((size_t)m * 3);
    if (fread(pts.data(), 16, (size_t)n, f) != (size_t)n || fread(q.data(), 24, (size_t)m, f) != (size_t)m) return 4;
    fclose(f);

    pc::SafeRegionCloud planner;
    planner.setParam(prm[0], prm[1], prm[2], prm[3]);
    planner.setPt(start, prm[3]);
    std::vector<double> r1((size_t)m);
    // before any cloud arrives: cloud_empty -> max_radius - search_margin
    double p0[3] = { start[0], start[1], start[2] };
    if (planner.radiusSearch(p0) != prm[2] - prm[1]) return 5;
    if (planner.setInput(pts.data(), n, 4) != PC_OK) return 6;
    for (int64_t k = 0; k < m; k++) r1[(size_t)k] = planner.radiusSearch(&q[3 * (size_t)k]);
    std::vector<float> qf((size_t)m * 3), r2((size_t)m);
    for (size_t i = 0; i < qf.size(); i++) qf[i] = (float)q[i];
    if (planner.radiusSearch(qf.data(), m, 3, r2.data()) != PC_OK) return 7;
    std::vector<uint8_t> col;
    if (planner.checkTrajPtCol(qf.data(), m, 3, col) != PC_OK) return 8;
    for (int64_t k = 0; k < m; k += 97)
        if (planner.checkTrajPtCol(&q[3 * (size_t)k]) != (r1[(size_t)k] < 0.0)) return 9;
    // the neighbouring rows: ground-truth arbiter and LiDAR observation
    std::vector<float> nd;
    const int64_t first_col = planner.firstCollision(qf.data(), m, 3, 0.3, &nd);
    int64_t expect_first = -1;
    for (int64_t k = 0; k < m && expect_first < 0; k++) if (nd[(size_t)k] < 0.3f && (double)nd[(size_t)k] < 0.3) expect_first = k;
    if (first_col != expect_first) return 11;
    std::vector<int32_t> seen;
    if (planner.observe(start, 6.0, seen) != PC_OK) return 12;
    for (size_t i = 1; i < seen.size(); i++) if (seen[i] <= seen[i - 1]) return 13;
    size_t brute = 0;
    for (int64_t i = 0; i < n; i++) {
        const double dx = (double)pts[4 * (size_t)i] - (double)(float)start[0], dy = (double)pts[4 * (size_t)i + 1] - (double)(float)start[1], dz = (double)pts[4 * (size_t)i + 2] - (double)(float)start[2];
        if ((dx * dx + dy * dy) + dz * dz <= 36.0) brute++;
    }
    if (brute != seen.size()) return 14;
    FILE *o = fopen(argv[2], "wb");
    if (!o) return 10;
    fwrite(r1.data(), 8, (size_t)m, o);
    fwrite(r2.data(), 4, (size_t)m, o);
    fwrite(col.data(), 1, (size_t)m, o);
    fwrite(&first_col, 8, 1, o);
    fwrite(nd.data(), 4, (size_t)m, o);
    fclose(o);
    return 0;
}
