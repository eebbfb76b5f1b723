// oracle/shim/pcl -- TEST INFRASTRUCTURE.  Stand-in for the PCL types named by the reference's corridor_finder.{h,cpp}: a
// point, a cloud, and pcl::search::KdTree whose nearestKSearch forwards to the reference's own Utils/kdtree (kd_nearest3,
// fp64) -- the parity target the north star names -- and returns the squared distance rounded to float32, the type PCL's
// interface imposes (std::vector<float> &k_sqr_distances).
#pragma once
namespace pcl { struct PointXYZ { float x, y, z, pad_; PointXYZ() : x(0), y(0), z(0), pad_(1.0f) {} PointXYZ(float x_, float y_, float z_) : x(x_), y(y_), z(z_), pad_(1.0f) {} }; }
