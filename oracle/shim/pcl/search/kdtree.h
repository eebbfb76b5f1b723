#pragma once
#include <vector>
#include <kdtree/kdtree.h>
#include "../point_cloud.h"
namespace pcl { namespace search {
// counters the harness reads (number of cloud queries issued by the planner code)
struct ShimStats { long long queries = 0, builds = 0; bool log_on = false; std::vector<float> log_d2; };
inline ShimStats &shim_stats() { static ShimStats s; return s; }

template <typename PointT>
class KdTree {
public:
    KdTree() : tree_(nullptr) {}
    ~KdTree() { if (tree_) kd_free(tree_); }
    KdTree(const KdTree &) = delete;
    KdTree &operator=(const KdTree &) = delete;
    // full rebuild (pcl::search::KdTree::setInputCloud): kd_create + n x kd_insert3, point i carries (void*)(i + 1).
    // Insertion order: a fixed pseudo-random permutation (insertion in the cloud's own order degenerates the unbalanced
    // kd-tree on sorted input); it only decides which of several EXACTLY equidistant points is returned.
    void setInputCloud(const typename PointCloud<PointT>::ConstPtr &cloud)
    {
        if (tree_) kd_free(tree_);
        tree_ = kd_create(3);
        const size_t n = cloud->points.size();
        std::vector<size_t> order(n);
        for (size_t i = 0; i < n; i++) order[i] = i;
        unsigned long long s = 0x9E3779B97F4A7C15ull;
        for (size_t i = n; i > 1; i--) {
            s = s * 6364136223846793005ull + 1442695040888963407ull;
            const size_t j = (size_t)((s >> 33) % i);
            const size_t t = order[i - 1]; order[i - 1] = order[j]; order[j] = t;
        }
        for (size_t k = 0; k < n; k++) {
            const PointT &p = cloud->points[order[k]];
            kd_insert3(tree_, (double)p.x, (double)p.y, (double)p.z, (void *)(order[k] + 1));
        }
        shim_stats().builds++;
    }
    int nearestKSearch(const PointT &q, int k, std::vector<int> &k_indices, std::vector<float> &k_sqr_distances) const
    {
        k_indices.clear(); k_sqr_distances.clear();
        if (!tree_ || k < 1) return 0;
        shim_stats().queries++;
        struct kdres *r = kd_nearest3(tree_, (double)q.x, (double)q.y, (double)q.z);
        if (!r) return 0;
        double pos[3];
        void *data = kd_res_item(r, pos);
        kd_res_free(r);
        const double dx = pos[0] - (double)q.x, dy = pos[1] - (double)q.y, dz = pos[2] - (double)q.z;
        double d2 = 0.0;                               /* the reference's accumulation order, kdtree.c:379-382 */
        d2 += dx * dx; d2 += dy * dy; d2 += dz * dz;
        k_indices.push_back((int)((size_t)data - 1));
        k_sqr_distances.push_back((float)d2);
        if (shim_stats().log_on) shim_stats().log_d2.push_back((float)d2);
        return 1;
    }
private:
    struct kdtree *tree_;
};
} }
