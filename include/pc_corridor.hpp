// pc_corridor.hpp -- header-only C++14 mirror of the CLOUD-FACING members of safeRegionRrtStar
// (Planner/include/pointcloudTraj/corridor_finder.h:17-150, Planner/src/corridor_finder.cpp) over the C ABI of
// pc_index.h.  Same member names, argument meaning and return values; the per-sample members get batched
// overloads, which is what the planner's loops call after the switch.
//
//   setParam        corridor_finder.cpp:17-23      safety_margin, search_margin, max_radius, sample_range
//   setStartPt      corridor_finder.cpp:43-50      refreshes start_pt (the centre of the sensing-range early-out)
//   setPt           corridor_finder.cpp:52-91      only its effect on radiusSearch: start_pt and sample_range = local_range (:87)
//   setInput        corridor_finder.cpp:93-99      full index rebuild per cloud message; empty cloud -> cloud_empty
//   radiusSearch    corridor_finder.cpp:113-133    early-outs, float32 cast of the point, 1-NN, min(sqrt(d2) - search_margin, max_radius)
//   checkTrajPtCol  corridor_finder.cpp:412-416    radiusSearch(pt) < 0
//   checkSafeTrajectory  Planner/src/sim_planning_demo.cpp:729-781 (a free function there; it only needs the cloud)
//   firstCollision  Planner/src/status_inspector.cpp:33-46   ground-truth collision check of executed positions
//   observe         Planner/src/camera_sensor.cpp:133-145    LiDAR-mode observation (all points within max_dist)
//   NodeSnapshotIndex   corridor_finder.cpp:428-437, 462-464 batched findNearstVertex / treeRewire neighbourhoods against a frozen node set
//   DeviceExpansion     corridor_finder.cpp:720-731, 333-358  one speculative batch of the expansion loop generated and answered on the device
//   exportCorridor      sim_planning_demo.cpp:571-592        (Path, Radius) -> PolynomialTrajectoryExtra.path_* / radii
//
// No Eigen/PCL/ROS dependency: points are plain double[3] / float arrays (pcl::PointXYZ is x,y,z,pad float32 =
// stride 4; Eigen::Vector3d::data() is double[3]).
#ifndef PC_CORRIDOR_HPP_
#define PC_CORRIDOR_HPP_

#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "pc_index.h"

namespace pc {

class SafeRegionCloud {
public:
    explicit SafeRegionCloud(int device = 0, int64_t max_points = 0, void *cuda_stream = nullptr)
    {
        if (pc_index_create(&ix_, device, max_points, cuda_stream) != PC_OK)
            throw std::runtime_error(std::string("pc_index_create: ") + pc_last_error(nullptr));
        params_.search_margin = 0.0; params_.max_radius = 0.0; params_.sample_range = 0.0;
        params_.start[0] = params_.start[1] = params_.start[2] = 0.0;
    }
    ~SafeRegionCloud() { pc_index_destroy(ix_); }
    SafeRegionCloud(const SafeRegionCloud &) = delete;
    SafeRegionCloud &operator=(const SafeRegionCloud &) = delete;

    void setParam(double safety_margin_, double search_margin_, double max_radius_, double sample_range_)
    {
        safety_margin = safety_margin_;
        params_.search_margin = search_margin_;
        params_.max_radius = max_radius_;
        params_.sample_range = sample_range_;
    }
    // PC_ARITH_FP64 (default) or PC_ARITH_PCL_FLOAT: the arithmetic of the radius epilogue (pc_index.h)
    int setRadiusArith(int mode) { return pc_index_set_radius_arith(ix_, mode); }
    void setStartPt(const double startPt[3]) { for (int a = 0; a < 3; a++) params_.start[a] = startPt[a]; }
    void setPt(const double startPt[3], double local_range) { setStartPt(startPt); params_.sample_range = local_range; }

    // cloud: n points, stride_floats = 3 (packed) or 4 (pcl::PointXYZ); host memory.  Returns PC_OK or an error code.
    int setInput(const float *xyz, int64_t n, int64_t stride_floats = 4)
    {
        cloud_empty = (n == 0);
        return pc_index_build(ix_, xyz, n, stride_floats, PC_HOST);
    }

    // one point, as in the reference (double in, double out).  Latency-bound: prefer the batch overload.
    double radiusSearch(const double search_Pt[3])
    {
        // the reference tests the sensing range on the double point BEFORE the float32 cast (corridor_finder.cpp:115-125)
        const double dx = search_Pt[0] - params_.start[0], dy = search_Pt[1] - params_.start[1], dz = search_Pt[2] - params_.start[2];
        if (std::sqrt(dx * dx + dy * dy + dz * dz) > params_.sample_range + params_.max_radius) return params_.max_radius - params_.search_margin;
        if (cloud_empty) return params_.max_radius - params_.search_margin;
        const float q[3] = { (float)search_Pt[0], (float)search_Pt[1], (float)search_Pt[2] };
        pc_radius_params p = params_;
        p.sample_range = -1.0;                      // already decided above
        float r = 0.f;
        check(pc_radius_batch(ix_, q, 1, 3, PC_HOST, PC_QUERY_UNSORTED, &p, &r, nullptr));
        return (double)r;
    }

    // m points (float32, stride 3 or 4, host memory) -> out_radius[m] (and the nearest point's index, nullable)
    int radiusSearch(const float *pts, int64_t m, int64_t stride_floats, float *out_radius, int32_t *out_idx = nullptr)
    {
        return pc_radius_batch(ix_, pts, m, stride_floats, PC_HOST, PC_QUERY_AUTO, &params_, out_radius, out_idx);
    }

    bool checkTrajPtCol(const double pt[3]) { return radiusSearch(pt) < 0.0; }

    int checkTrajPtCol(const float *pts, int64_t m, int64_t stride_floats, std::vector<uint8_t> &collides)
    {
        std::vector<float> r((size_t)m);
        int rc = radiusSearch(pts, m, stride_floats, r.data());
        if (rc != PC_OK) return rc;
        collides.resize((size_t)m);
        for (int64_t k = 0; k < m; k++) collides[(size_t)k] = r[(size_t)k] < 0.f;
        return PC_OK;
    }

    // checkSafeTrajectory for a batch of piecewise Bezier trajectories (layout: pc_clearance_batch in pc_index.h).
    // first_hit[t] >= 0  <=>  the reference's checkSafeTrajectory(stop_time) returns true for trajectory t.
    int checkSafeTrajectory(const std::vector<pc_traj> &traj, const std::vector<int32_t> &seg_order, const std::vector<double> &seg_T,
                            const std::vector<int64_t> &seg_coef_off, const std::vector<double> &coef, double stop_time,
                            std::vector<int32_t> &first_hit, std::vector<float> *min_radius = nullptr, double dt = 0.02)
    {
        first_hit.resize(traj.size());
        if (min_radius) min_radius->resize(traj.size());
        return pc_clearance_batch(ix_, traj.data(), (int64_t)traj.size(), seg_order.data(), seg_T.data(), seg_coef_off.data(),
                                  (int64_t)seg_order.size(), coef.data(), (int64_t)coef.size(), PC_HOST, dt, stop_time, &params_,
                                  first_hit.data(), min_radius ? min_radius->data() : nullptr, nullptr);
    }

    // Ground-truth collision arbiter of the experiment harness (Planner/src/status_inspector.cpp:33-46): for every executed
    // position, is the nearest map point closer than col_rad?  Returns the ordinal of the first colliding position, -1 if none.
    int64_t firstCollision(const float *positions, int64_t m, int64_t stride_floats, double col_rad, std::vector<float> *nearest_dist = nullptr)
    {
        std::vector<float> d2((size_t)m);
        check(pc_nearest_batch(ix_, positions, m, stride_floats, PC_HOST, PC_QUERY_AUTO, nullptr, d2.data()));
        int64_t first = -1;
        if (nearest_dist) nearest_dist->resize((size_t)m);
        for (int64_t k = 0; k < m; k++) {
            const float d = std::sqrt(d2[(size_t)k]);      // sqrt(points_distances[0]) < col_rad: a std::vector<float>, float sqrt
            if (nearest_dist) (*nearest_dist)[(size_t)k] = d;
            if (first < 0 && (double)d < col_rad) first = k;
        }
        return first;
    }

    // LiDAR-mode observation of the sensor node (Planner/src/camera_sensor.cpp:133-145): indices of all map points within
    // max_dist of the sensor position, ascending.
    int observe(const double sensor_pos[3], double max_dist, std::vector<int32_t> &indices)
    {
        int64_t n = 0;
        int rc = pc_sphere_gather(ix_, sensor_pos, max_dist, PC_HOST, nullptr, 0, &n);
        if (rc != PC_OK) return rc;
        indices.resize((size_t)n);
        if (n == 0) return PC_OK;
        return pc_sphere_gather(ix_, sensor_pos, max_dist, PC_HOST, indices.data(), n, &n);
    }

    pc_index *handle() { return ix_; }
    const pc_radius_params &params() const { return params_; }
    const char *lastError() const { return pc_last_error(ix_); }

    double safety_margin = 0.0;     // kept for the planner's own node rejection (corridor_finder.cpp:732), unused here
    bool cloud_empty = true;

private:
    void check(int rc) const
    {
        if (rc != PC_OK) throw std::runtime_error(std::string("pcindex: ") + pc_last_error(ix_));
    }
    pc_index *ix_ = nullptr;
    pc_radius_params params_;
};

// Corridor export: the (Path, Radius) pair of safeRegionRrtStar::getPath as the planner packs it into the path_x / path_y /
// path_z / radii arrays of quadrotor_msgs/PolynomialTrajectoryExtra (Planner/src/sim_planning_demo.cpp:571-592,
// Utils/quadrotor_msgs/msg/PolynomialTrajectoryExtra.msg): k + 1 entries, entry 0 REPEATS the first sphere (:582-585), entries
// 1 .. k are the spheres from the root to the goal.  Plain arrays of doubles (float64[] in the message).
struct CorridorExport {
    std::vector<double> path_x, path_y, path_z, radii;
    size_t size() const { return radii.size(); }
};

// path: k x 3 sphere centres (row-major), radius: k radii -- the outputs of tracePath (pc::SafeRegionRrtStarDriver::path /
// ::radius, or the reference's getPath()); k == 0 gives an empty export
inline CorridorExport exportCorridor(const double *path, const double *radius, int64_t k)
{
    CorridorExport e;
    if (k <= 0) return e;
    e.path_x.resize((size_t)k + 1); e.path_y.resize((size_t)k + 1); e.path_z.resize((size_t)k + 1); e.radii.resize((size_t)k + 1);
    e.path_x[0] = path[0]; e.path_y[0] = path[1]; e.path_z[0] = path[2]; e.radii[0] = radius[0];
    for (int64_t i = 0; i < k; i++) {
        e.path_x[(size_t)i + 1] = path[3 * i]; e.path_y[(size_t)i + 1] = path[3 * i + 1]; e.path_z[(size_t)i + 1] = path[3 * i + 2];
        e.radii[(size_t)i + 1] = radius[i];
    }
    return e;
}

// The RRT* NODE tree's nearest-vertex queries for a whole batch of samples: the reference asks kd_nearestf on
// the node kd-tree once per sample (corridor_finder.cpp:428-437); in the speculative-batch driver the K samples of a batch
// see the same frozen node set, so the set is indexed once (a few thousand centres: ~0.1 ms) and the K queries are one
// exact batched nearest call -- same metric as kd_nearestf (fp64 distance to the float32 centres), ties -> lowest node index.
class NodeSnapshotIndex {
public:
    explicit NodeSnapshotIndex(int device = 0, int64_t max_nodes = 1 << 16)
    {
        if (pc_index_create(&ix_, device, max_nodes, nullptr) != PC_OK)
            throw std::runtime_error(std::string("pc_index_create: ") + pc_last_error(nullptr));
    }
    ~NodeSnapshotIndex() { pc_index_destroy(ix_); }
    NodeSnapshotIndex(const NodeSnapshotIndex &) = delete;
    NodeSnapshotIndex &operator=(const NodeSnapshotIndex &) = delete;

    // index the frozen node set once per batch ...
    int build(const float *node_pos, int64_t n_nodes) { return pc_index_build(ix_, node_pos, n_nodes, 3, PC_HOST); }
    // ... then findNearstVertex for k samples (kd_nearestf, corridor_finder.cpp:428-437) ...
    int nearest(const float *samples, int64_t k, int32_t *out_nearest, float *out_d2 = nullptr)
    {
        return pc_nearest_batch(ix_, samples, k, 3, PC_HOST, PC_QUERY_AUTO, out_nearest, out_d2);
    }
    // ... and the neighbourhoods treeRewire asks for (kd_nearest_rangef(pos, 2 * radius), corridor_finder.cpp:462-464): for
    // centre j all snapshot nodes within ranges[j], as CSR lists of node indices in ascending (= insertion) order
    int range(const float *centers, const float *ranges, int64_t k, std::vector<int64_t> &offsets, std::vector<int32_t> &idx)
    {
        offsets.assign((size_t)k + 1, 0);
        std::vector<double> r((size_t)k);
        for (int64_t j = 0; j < k; j++) r[(size_t)j] = (double)ranges[j];
        // one call when the lists fit the capacity kept from earlier batches, a second one with the exact size otherwise
        if ((int64_t)idx.size() < 32 * k) idx.resize((size_t)(32 * k) + 1);
        int rc = pc_range_batch(ix_, centers, k, 3, PC_HOST, r.data(), 0, offsets.data(), idx.data(), (int64_t)idx.size());
        if (rc != PC_ECAP) return rc;
        idx.resize((size_t)offsets[(size_t)k] + (size_t)offsets[(size_t)k] / 2 + 1);
        return pc_range_batch(ix_, centers, k, 3, PC_HOST, r.data(), 0, offsets.data(), idx.data(), (int64_t)idx.size());
    }
    int nearest(const float *node_pos, int64_t n_nodes, const float *samples, int64_t k, int32_t *out_nearest, float *out_d2 = nullptr)
    {
        int rc = build(node_pos, n_nodes);
        if (rc != PC_OK) return rc;
        return nearest(samples, k, out_nearest, out_d2);
    }
    const char *lastError() const { return pc_last_error(ix_); }
    pc_index *handle() { return ix_; }

private:
    pc_index *ix_ = nullptr;
};

// One speculative batch of the expansion loop without per-sample host traffic (pc_expand_batch): the samples are generated on
// the device from the planner's engine state, steered against the frozen node set, answered against the cloud, and only the
// candidates the loop keeps come back.  `nodes` holds the index of the node set (rebuilt per batch).
class DeviceExpansion {
public:
    DeviceExpansion(SafeRegionCloud &cloud, NodeSnapshotIndex &nodes) : cloud_(cloud), nodes_(nodes) {}

    // returns PC_OK or an error code; centers (count x 3), radii (count), nearest (count, nullable) in sample order
    int run(const pc_sampler &sampler, const pc_node_set &set, double z_l, int64_t k, std::vector<double> &centers, std::vector<double> &radii,
            uint32_t *engine_state_after, std::vector<int32_t> *nearest = nullptr)
    {
        if ((int64_t)buf_.size() < k) buf_.resize((size_t)k);
        int64_t count = 0;
        const int rc = pc_expand_batch(cloud_.handle(), nodes_.handle(), &set, &sampler, &cloud_.params(), z_l, cloud_.safety_margin, k,
                                       buf_.data(), (int64_t)buf_.size(), &count, engine_state_after);
        if (rc != PC_OK) return rc;
        centers.resize((size_t)count * 3); radii.resize((size_t)count);
        if (nearest) nearest->resize((size_t)count);
        for (int64_t i = 0; i < count; i++) {
            for (int a = 0; a < 3; a++) centers[(size_t)i * 3 + a] = buf_[(size_t)i].center[a];
            radii[(size_t)i] = (double)buf_[(size_t)i].radius;
            if (nearest) (*nearest)[(size_t)i] = buf_[(size_t)i].nearest;
        }
        return PC_OK;
    }

private:
    SafeRegionCloud &cloud_;
    NodeSnapshotIndex &nodes_;
    std::vector<pc_candidate> buf_;
};

}  // namespace pc
#endif  // PC_CORRIDOR_HPP_
