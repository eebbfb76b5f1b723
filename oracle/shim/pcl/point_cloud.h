#pragma once
#include <memory>
#include <vector>
#include <cstdint>
#include "point_types.h"
namespace pcl {
template <typename PointT>
struct PointCloud {
    typedef std::shared_ptr<PointCloud<PointT> > Ptr;
    typedef std::shared_ptr<const PointCloud<PointT> > ConstPtr;
    std::vector<PointT> points;
    uint32_t width = 0, height = 0;
    bool is_dense = true;
    bool empty() const { return points.empty(); }
    size_t size() const { return points.size(); }
    void push_back(const PointT &p) { points.push_back(p); }
    Ptr makeShared() const { return Ptr(new PointCloud<PointT>(*this)); }
};
}
