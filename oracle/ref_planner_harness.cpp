// ref_planner_harness.cpp -- TEST INFRASTRUCTURE ONLY (never linked into libpcindex.so).
//
// C entry points around the UNMODIFIED reference planner sources, compiled where they lie under $(REF) by oracle/Makefile
// into oracle/_ref/libplanner_ref.so:
//   * Planner/src/corridor_finder.cpp is #included as a whole (so that its `inline` members -- radiusSearch, genSample,
//     genNewNode ... -- are callable from here), against the stand-in headers of oracle/shim/ (Eigen, pcl, ros: absent from
//     this image) and the reference's own Utils/kdtree (kdtree.c, compiled beside it);
//   * Planner/src/sim_planning_demo.cpp:715-781 (getPosFromBezier, checkSafeTrajectory) is cut out of the reference file by
//     the Makefile at build time (sed, into the git-ignored _ref/ directory) and #included below, after the globals it names.
// The cloud query behind safeRegionRrtStar::radiusSearch is oracle/shim/pcl/search/kdtree.h: kd_nearest3 of the reference
// kd-tree, squared distance rounded to float32 (PCL's interface type).  ros::Time is a step counter (shim/ros/ros.h), so the
// reference's wall-clock budgets are exact iteration budgets.  The planner object lives in zero-filled storage: the members
// the reference never initialises (elli_l, elli_s, which SafeRegionExpansion samples from once an end node exists,
// corridor_finder.cpp:742-743 with :362-383) read as 0.0 instead of stack garbage.
#include <cstring>
#include <new>
#include <string>

#include "src/corridor_finder.cpp"          // -I $(REF)/Planner

using namespace pcl;

// ---- the globals sim_planning_demo.cpp:715-781 names (sim_planning_demo.cpp:56-86) -------------------------------------
struct ShimHeader { ros::Time stamp; std::string frame_id; };
struct ShimOdometry { ShimHeader header; };
struct ShimPointCloud2 { ShimHeader header; };
struct ShimPublisher { template <class T> void publish(const T &) {} };
namespace pcl { template <class CloudT, class MsgT> void toROSMsg(const CloudT &, MsgT &) {} }

static bool _is_traj_exist = false;
static ShimOdometry _odom;
static ros::Time _start_time;
static int _segment_num = 0;
static vector<int> _poly_orderList;
static MatrixXd _PolyCoeff;
static VectorXd _Time;
static vector<VectorXd> _CList;
static double _stop_time = 0.0;
static ShimPointCloud2 traj_stop_pts;
static PointCloud<PointXYZ> traj_stop_pts_pcd;
static ShimPublisher _vis_stop_traj_points;

alignas(64) static unsigned char g_storage[sizeof(safeRegionRrtStar)];
static safeRegionRrtStar *g_planner = nullptr;
#define _rrtPathPlaner (*g_planner)

#include "traj_check_extract.inc"           // -I _ref : sim_planning_demo.cpp:715-781, cut by the Makefile

// Bernstein::setParam's binomial table (Planner/src/bezier_base.cpp:35-48, 256-266): C(order)(k) = order! / (k! (order-k)!)
// in int arithmetic.  bezier_base.cpp itself needs Eigen's LDLT and is not compiled.
static void fill_clist(int order_max)
{
    _CList.clear();
    for (int order = 0; order <= order_max; order++) {
        VectorXd C_(order + 1);
        for (int k = 0; k <= order; k++) {
            int fn = 1, fk = 1, fnk = 1;
            for (int i = order; i > 0; i--) fn *= i;
            for (int i = k; i > 0; i--) fk *= i;
            for (int i = order - k; i > 0; i--) fnk *= i;
            C_(k) = fn / (fk * fnk);
        }
        _CList.push_back(C_);
    }
}

extern "C" {

// (re)create the planner in zero-filled storage and apply setParam (sim_planning_demo.cpp:487)
int rp_create(double safety_margin, double search_margin, double max_radius, double sample_range)
{
    if (g_planner) g_planner->~safeRegionRrtStar();
    memset(g_storage, 0, sizeof g_storage);
    g_planner = new (g_storage) safeRegionRrtStar();
    g_planner->setParam(safety_margin, search_margin, max_radius, sample_range);
    ros::shim_clock() = 0.0; ros::shim_tick() = 0.0; ros::shim_warnings() = 0;
    pcl::search::shim_stats() = pcl::search::ShimStats();
    fill_clist(12);
    return 0;
}

// rcvPointCloudCallBack -> setInput (sim_planning_demo.cpp:159-167): full index rebuild
void rp_set_input(const float *xyz, long long n, long long stride)
{
    pcl::PointCloud<pcl::PointXYZ> cloud;
    cloud.points.reserve((size_t)n);
    for (long long i = 0; i < n; i++) cloud.points.push_back(pcl::PointXYZ(xyz[i * stride], xyz[i * stride + 1], xyz[i * stride + 2]));
    cloud.width = (uint32_t)n; cloud.height = 1;
    g_planner->setInput(cloud);
}

void rp_reset(void) { g_planner->reset(); }

void rp_set_pt(const double *s, const double *e, double xl, double xh, double yl, double yh, double zl, double zh,
               double local_range, int max_iter, double sample_portion, double goal_portion)
{
    g_planner->setPt(Vector3d(s[0], s[1], s[2]), Vector3d(e[0], e[1], e[2]), xl, xh, yl, yh, zl, zh, local_range, max_iter,
                     sample_portion, goal_portion);
}

void rp_set_start_pt(const double *s, const double *e) { g_planner->setStartPt(Vector3d(s[0], s[1], s[2]), Vector3d(e[0], e[1], e[2])); }

// time budgets as iteration budgets: tick 1 s per ros::Time::now(), limit n_iter s  =>  exactly n_iter loop iterations
void rp_expand(double n_iter) { ros::shim_tick() = 1.0; g_planner->SafeRegionExpansion(n_iter); ros::shim_tick() = 0.0; }
void rp_refine(double n_iter) { ros::shim_tick() = 1.0; g_planner->SafeRegionRefine(n_iter); ros::shim_tick() = 0.0; }
void rp_evaluate(void) { ros::shim_tick() = 0.0; g_planner->SafeRegionEvaluate(1e9); }

int rp_reset_root(const double *t) { Vector3d v(t[0], t[1], t[2]); g_planner->resetRoot(v); return g_planner->getGlobalNaviStatus() ? 1 : 0; }

int rp_get_path(double *path, double *radius, int cap)
{
    pair<MatrixXd, VectorXd> pr = g_planner->getPath();
    const int n = pr.first.rows();
    for (int i = 0; i < n && i < cap; i++) {
        for (int c = 0; c < 3; c++) path[3 * i + c] = pr.first(i, c);
        radius[i] = pr.second(i);
    }
    return n;
}

// every node of the tree in NodeList order: x, y, z, radius, g, index of its parent in the same list (-1: none), valid
int rp_get_tree(double *out7, int cap)
{
    vector<NodePtr> nodes = g_planner->getTree();
    const int n = (int)nodes.size();
    for (int i = 0; i < n && i < cap; i++) {
        NodePtr p = nodes[i];
        int parent = -1;
        if (p->preNode_ptr) for (int j = 0; j < n; j++) if (nodes[j] == p->preNode_ptr) { parent = j; break; }
        double *o = out7 + 7 * i;
        o[0] = p->coord(0); o[1] = p->coord(1); o[2] = p->coord(2); o[3] = p->radius; o[4] = p->g; o[5] = parent; o[6] = p->valid ? 1.0 : 0.0;
    }
    return n;
}

// [0] nodes, [1] path exists, [2] cloud queries issued, [3] index rebuilds, [4] ROS_WARN/ROS_ERROR count, [5] global navi status
void rp_stats(long long out[6])
{
    out[0] = (long long)g_planner->getTree().size();
    out[1] = g_planner->getPathExistStatus() ? 1 : 0;
    out[2] = pcl::search::shim_stats().queries;
    out[3] = pcl::search::shim_stats().builds;
    out[4] = ros::shim_warnings();
    out[5] = g_planner->getGlobalNaviStatus() ? 1 : 0;
}

// k calls of the UNMODIFIED genSample (corridor_finder.cpp:333-383) on the planner's own engine (eng(0), :12)
void rp_gen_samples(long long k, double *out3)
{
    for (long long i = 0; i < k; i++) { const Vector3d v = g_planner->genSample(); out3[3 * i] = v(0); out3[3 * i + 1] = v(1); out3[3 * i + 2] = v(2); }
}
double rp_radius_search(const double *p) { Vector3d v(p[0], p[1], p[2]); return g_planner->radiusSearch(v); }
void rp_radius_batch(const double *p, long long n, double *out) { for (long long i = 0; i < n; i++) out[i] = rp_radius_search(p + 3 * i); }
int rp_check_traj_pt_col(const double *p) { Vector3d v(p[0], p[1], p[2]); return g_planner->checkTrajPtCol(v) ? 1 : 0; }
void rp_bezier_pos(int order, const double *coef_row, double u, double *out3)
{
    _poly_orderList.assign(1, order);
    MatrixXd m(1, 3 * (order + 1));
    for (int j = 0; j < 3 * (order + 1); j++) m(0, j) = coef_row[j];
    Vector3d r;
    getPosFromBezier(m, u, 0, r);
    out3[0] = r(0); out3[1] = r(1); out3[2] = r(2);
}

// checkSafeTrajectory (sim_planning_demo.cpp:729-781) on one piecewise trajectory.  coef: row-major, row i = segment i's
// [x|y|z] blocks, ld doubles per row.  t_now = (odom stamp - trajectory start).  Returns the reference's return value
// (1: a sample collides); the float32 sample points it visited (up to and including the colliding one) go to out_pts.
// *n_searched = how many of those samples reached the cloud query (the others took radiusSearch's early-outs);
// *min_d2 = the smallest squared distance (float32, as PCL's interface returns it) any of them saw, +inf if none.
int rp_check_safe_trajectory(int n_seg, const int *order, const double *T, const double *coef, long long ld,
                             double t_now, double stop_time, float *out_pts, long long cap, long long *n_pts,
                             long long *n_searched, float *min_d2)
{
    pcl::search::shim_stats().log_d2.clear();
    pcl::search::shim_stats().log_on = true;
    _is_traj_exist = true;
    _segment_num = n_seg;
    _poly_orderList.assign(order, order + n_seg);
    _Time.resize(n_seg);
    _PolyCoeff.resize(n_seg, (int)ld);
    for (int i = 0; i < n_seg; i++) {
        _Time(i) = T[i];
        for (long long j = 0; j < ld; j++) _PolyCoeff(i, (int)j) = coef[(long long)i * ld + j];
    }
    _start_time = ros::Time(0.0);
    _odom.header.stamp = ros::Time(t_now);
    _stop_time = stop_time;
    const bool hit = checkSafeTrajectory(stop_time);
    pcl::search::shim_stats().log_on = false;
    if (n_searched) *n_searched = (long long)pcl::search::shim_stats().log_d2.size();
    if (min_d2) {
        float m = INFINITY;
        for (float v : pcl::search::shim_stats().log_d2) m = v < m ? v : m;
        *min_d2 = m;
    }
    const long long n = (long long)traj_stop_pts_pcd.points.size();
    for (long long k = 0; k < n && k < cap; k++) {
        out_pts[3 * k] = traj_stop_pts_pcd.points[(size_t)k].x;
        out_pts[3 * k + 1] = traj_stop_pts_pcd.points[(size_t)k].y;
        out_pts[3 * k + 2] = traj_stop_pts_pcd.points[(size_t)k].z;
    }
    if (n_pts) *n_pts = n;
    return hit ? 1 : 0;
}

}  // extern "C"
