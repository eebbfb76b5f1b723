// pc_rrt.hpp -- header-only C++14 restatement of the CALLER of the hot path: the safe-region RRT* expansion of
// safeRegionRrtStar (Planner/src/corridor_finder.cpp), with the per-sample cloud query behind a pluggable, BATCHED
// radius provider.  This is the first "next" row of SURVEY.md section 8(f): it turns queries/s into corridor-growth
// milliseconds and shows how the sequential planner loop consumes batched radius calls.
//
//   genSample            corridor_finder.cpp:333-383      goal bias / local box / global box / informed ellipsoid
//   findNearstVertex     corridor_finder.cpp:428-437      kd_nearestf on the node tree
//   genNewNode           corridor_finder.cpp:385-410      steer to the nearest sphere's surface, radiusSearch(center)
//   checkEnd             corridor_finder.cpp:418-426
//   checkNodeRelation    corridor_finder.cpp:439-455
//   treeRewire           corridor_finder.cpp:458-571      kd_nearest_rangef(2 * radius), choose parent, rewire neighbours
//   treePrune / clearBranchS / removeInvalid   corridor_finder.cpp:150-232
//   updateHeuristicRegion corridor_finder.cpp:298-330
//   tracePath / checkValidEnd / isSuccessor     corridor_finder.cpp:578-641, 669-702
//   SafeRegionExpansion  corridor_finder.cpp:704-763      (the wall-clock budget becomes an iteration budget)
//   SafeRegionRefine     corridor_finder.cpp:765-815      the same loop on the existing tree
//   SafeRegionEvaluate   corridor_finder.cpp:817-936      lazy re-validation of the best path against a new cloud (batched)
//   treeRepair           corridor_finder.cpp:938-1021     re-validation of the failed nodes' neighbourhoods (one batched call)
//
// Two drivers over the same restated logic:
//   expand(max_iter)              one radius query per iteration, exactly the reference's loop order ("replay" mode: with
//                                 two radius providers that return the same values the corridors are bit-identical)
//   expandBatched(max_iter, K)    speculative: K samples are steered against a frozen snapshot of the node tree, ONE
//                                 radius call answers the K centres, then the K candidates are inserted / rewired in order
//                                 (a candidate's centre was chosen against the snapshot; its radius is exact for that
//                                 centre, so every sphere of the resulting corridor is still obstacle-free)
//
//   + setDeviceBatch(fn)          the snapshot phase of every batch (K x genSample, nearest vertex, steering, radius, the loop's
//                                 early rejections) in ONE provider call that is handed the engine state instead of samples
//                                 (pc_expand_batch generates the stream on the device); used while no path is known -- the
//                                 informed-ellipsoid samples (libm calls) stay on the host.  Same corridor, bit for bit.
//
// Deviations from the reference, on purpose: (1) iteration budget instead of ros::Time; (2) the informed-sampling
// ellipsoid is updated when an end node is found, as SafeRegionRefine does (:796) -- in SafeRegionExpansion that call is
// commented out (:742) while inform_status is still set, which samples from uninitialised elli_l / elli_s; (3) the node
// tree is a small dynamic kd-tree written here (float positions like kd_insertf / kd_nearestf / kd_nearest_rangef),
// range results are consumed in ascending insertion order; (4) treePrune compares with a 1e-4 slack (see there); (5)
// SafeRegionRefine is also bounded by max_samples.
// setReferenceQuirks(true) switches (2)-(5) back to the reference's exact behaviour (elli_l = elli_s = 0 until Refine finds
// a better end, range results in kd_nearest_rangef's own order, exact float prune test, unbounded Refine): in that mode
// expand() / refine() / evaluate() reproduce the UNMODIFIED corridor_finder.cpp bit for bit -- node list, parents, costs,
// radii and corridor -- which tests/test_planner_ref.py checks against the compiled reference (oracle/_ref/libplanner_ref.so).
// No Eigen / PCL / ROS.
#ifndef PC_RRT_HPP_
#define PC_RRT_HPP_

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <functional>
#include <limits>
#include <memory>
#include <random>
#include <sstream>
#include <vector>

namespace pc {

struct RrtNode {                 // Planner/include/pointcloudTraj/data_type.h:12-51
    double coord[3];
    float radius;
    bool valid = true, best = false, change = false;
    int rel_id = -2;
    float rel_dis = -1.0f;
    RrtNode *pre = nullptr;
    std::vector<RrtNode *> nxt;
    float g, f;
    int64_t serial = 0;          // insertion order into the node tree (deterministic iteration of range results)
    RrtNode(const double c[3], float radius_, float g_, float f_) : radius(radius_), g(g_), f(f_) { coord[0] = c[0]; coord[1] = c[1]; coord[2] = c[2]; }
};

// dynamic 3-D kd-tree over node centres cast to float (kd_insertf / kd_nearestf / kd_nearest_rangef semantics: split axis
// cycles with depth, strictly-less goes left; nearest = smallest fp64 distance to the float positions)
class NodeKdTree {
public:
    void clear() { items_.clear(); }
    void insert(const float pos[3], RrtNode *n)
    {
        Item it; it.p[0] = pos[0]; it.p[1] = pos[1]; it.p[2] = pos[2]; it.node = n; it.lo = it.hi = -1; it.axis = 0;
        const int me = (int)items_.size();
        if (me > 0) {
            int cur = 0;
            for (;;) {
                const int a = items_[cur].axis;
                int &slot = (it.p[a] < items_[cur].p[a]) ? items_[cur].lo : items_[cur].hi;
                if (slot < 0) { slot = me; it.axis = (a + 1) % 3; break; }
                cur = slot;
            }
        }
        items_.push_back(it);
    }
    RrtNode *nearest(const float q[3], double *out_d2 = nullptr) const
    {
        if (items_.empty()) return nullptr;
        int best = 0;
        double best_d2 = d2(0, q);
        search(0, q, best, best_d2);
        if (out_d2) *out_d2 = best_d2;
        return items_[best].node;
    }
    // all nodes with d2 <= range^2, in the order kd_res_next hands them out after kd_nearest_rangef (kdtree.c:262-293: a node
    // is tested before its subtrees, the side of the query first, the far side only if |dx| < range -- strict --, and
    // rlist_insert puts every hit at the HEAD of the list, :810-828 with dist_sq = -1, so the list is the visit order reversed)
    void rangeReferenceOrder(const float q[3], float r, std::vector<RrtNode *> &out) const
    {
        out.clear();
        if (items_.empty()) return;
        const double qd[3] = { (double)q[0], (double)q[1], (double)q[2] }, rd = (double)r;
        rangeRec(0, qd, rd, out);
        std::reverse(out.begin(), out.end());
    }
    // all nodes with d2 <= range^2, ascending insertion order
    void range(const float q[3], float r, std::vector<RrtNode *> &out) const
    {
        out.clear();
        if (items_.empty()) return;
        std::vector<int> stack{0};
        const double r2 = (double)r * (double)r;
        std::vector<int> hits;
        while (!stack.empty()) {
            const int i = stack.back(); stack.pop_back();
            if (d2(i, q) <= r2) hits.push_back(i);
            const int a = items_[i].axis;
            const double dx = (double)q[a] - (double)items_[i].p[a];
            const int near_c = dx <= 0.0 ? items_[i].lo : items_[i].hi, far_c = dx <= 0.0 ? items_[i].hi : items_[i].lo;
            if (near_c >= 0) stack.push_back(near_c);
            if (far_c >= 0 && std::fabs(dx) <= (double)r) stack.push_back(far_c);
        }
        std::sort(hits.begin(), hits.end());
        for (int i : hits) out.push_back(items_[i].node);
    }
    size_t size() const { return items_.size(); }

private:
    struct Item { float p[3]; RrtNode *node; int lo, hi, axis; };
    void rangeRec(int i, const double q[3], double r, std::vector<RrtNode *> &out) const
    {
        if (i < 0) return;
        double s = 0.0;
        for (int a = 0; a < 3; a++) { const double d = (double)items_[i].p[a] - q[a]; s += d * d; }
        if (s <= r * r) out.push_back(items_[i].node);
        const int a = items_[i].axis;
        const double dx = q[a] - (double)items_[i].p[a];
        rangeRec(dx <= 0.0 ? items_[i].lo : items_[i].hi, q, r, out);
        if (std::fabs(dx) < r) rangeRec(dx <= 0.0 ? items_[i].hi : items_[i].lo, q, r, out);
    }
    double d2(int i, const float q[3]) const
    {
        double s = 0.0;
        for (int a = 0; a < 3; a++) { const double d = (double)items_[i].p[a] - (double)q[a]; s += d * d; }
        return s;
    }
    void search(int i, const float q[3], int &best, double &best_d2) const
    {
        const int a = items_[i].axis;
        const double dx = (double)q[a] - (double)items_[i].p[a];
        const int near_c = dx <= 0.0 ? items_[i].lo : items_[i].hi, far_c = dx <= 0.0 ? items_[i].hi : items_[i].lo;
        if (near_c >= 0) search(near_c, q, best, best_d2);
        const double d = d2(i, q);
        if (d < best_d2) { best_d2 = d; best = i; }
        if (far_c >= 0 && dx * dx < best_d2) search(far_c, q, best, best_d2);
    }
    std::vector<Item> items_;
};

// radius provider: answer n centres (n x 3 doubles) with safeRegionRrtStar::radiusSearch values
using RadiusBatchFn = std::function<void(const double *centers, int n, double *out_radius)>;

class SafeRegionRrtStarDriver {
public:
    explicit SafeRegionRrtStarDriver(RadiusBatchFn radius) : radius_(std::move(radius)), eng_(0) {}   // eng(0): corridor_finder.cpp:12
    ~SafeRegionRrtStarDriver() { destroyTree(); }

    void setParam(double safety_margin_, double search_margin_, double max_radius_, double sample_range_)
    {
        safety_margin = safety_margin_; search_margin = search_margin_; max_radius = max_radius_; sample_range = sample_range_;
    }

    // corridor_finder.cpp:52-91
    void setPt(const double startPt[3], const double endPt[3], double xl, double xh, double yl, double yh, double zl, double zh,
               double local_range, int max_iter, double sample_portion, double goal_portion)
    {
        for (int a = 0; a < 3; a++) { start_pt[a] = startPt[a]; end_pt[a] = endPt[a]; }
        x_l = xl; x_h = xh; y_l = yl; y_h = yh; z_l = zl; z_h = zh;
        rand_x = U(x_l, x_h); rand_y = U(y_l, y_h);
        rand_z = rand_z_in = U(z_l + safety_margin, z_h);
        rand_bias = U(0.0, 1.0);
        rand_x_in = U(start_pt[0] - sample_range, start_pt[0] + sample_range);
        rand_y_in = U(start_pt[1] - sample_range, start_pt[1] + sample_range);
        min_distance = dist(start_pt, end_pt);
        updateEllipsoidFrame(start_pt);
        sample_range = local_range; max_samples = max_iter; inlier_ratio = sample_portion; goal_ratio = goal_portion;
    }

    // corridor_finder.cpp:25-41
    void reset()
    {
        destroyTree();
        end_list_.clear(); invalid_set_.clear(); path_list_.clear();
        best_end_ptr = nullptr; root_node = nullptr;
        path_exist_status = true; inform_status = false; best_distance = inf();
        path.clear(); radius.clear();
    }

    // reproduce the reference bit for bit (see the header comment); off by default
    void setReferenceQuirks(bool on) { quirks_ = on; }

    // SafeRegionExpansion, one cloud query per iteration (corridor_finder.cpp:704-763)
    int expand(int max_iterations) { initRoot(); return growLoop(max_iterations, true); }

    // SafeRegionRefine (corridor_finder.cpp:765-815): the same loop on the existing tree (more samples, more rewiring)
    int refine(int max_iterations) { return growLoop(max_iterations, false); }

    int growLoop(int max_iterations, bool expansion)
    {
        int it = 0;
        in_expansion_ = expansion;
        for (; it < max_iterations && (it < max_samples || (quirks_ && !expansion)); it++) {
            double s[3];
            genSample(s);
            RrtNode *nearest = findNearestVertex(s);
            if (!nearest || !nearest->valid) continue;
            double c[3];
            steer(s, nearest, c);
            double r;
            radius_(c, 1, &r);
            cloud_queries++;
            tryInsert(c, r, nearest);
        }
        removeInvalid();
        tracePath();
        return it;
    }

    // speculative batches of K samples against a frozen snapshot of the node tree
    int expandBatched(int max_iterations, int K) { initRoot(); return refineBatched(max_iterations, K); }

    int refineBatched(int max_iterations, int K)
    {
        in_expansion_ = false;
        std::vector<double> centers((size_t)K * 3), radii((size_t)K), samples((size_t)K * 3);
        std::vector<float> node_pos, sample_pos((size_t)K * 3), center_pos((size_t)K * 3), ranges((size_t)K);
        std::vector<int32_t> nearest_idx((size_t)K);
        std::vector<int64_t> range_off;
        std::vector<int32_t> range_idx;
        std::vector<RrtNode *> snapshot, batch_nodes, found;
        std::vector<double> node_coord;
        std::vector<float> node_radius;
        std::vector<uint8_t> node_valid;
        int it = 0;
        while (it < max_iterations && it < max_samples) {
            const int k_now = std::min(K, std::min(max_iterations, max_samples) - it);
            int n = 0;
            const bool on_device = device_batch_ && !inform_status;      // the informed-ellipsoid samples are drawn here
            if (centers.size() < (size_t)K * 3) centers.resize((size_t)K * 3);
            if (radii.size() < (size_t)K) radii.resize((size_t)K);
            if (!on_device) for (int j = 0; j < k_now; j++) genSample(&samples[(size_t)j * 3]);
            if (snapshot_nearest_ || snapshot_range_ || on_device) {
                // the K nearest-vertex queries of a batch go against the SAME frozen node set, so they are one
                // batched exact-NN call on the node centres (float positions, like kd_nearestf)
                snapshot = node_list_;
                node_pos.resize(snapshot.size() * 3);
                for (size_t i = 0; i < snapshot.size(); i++) for (int a = 0; a < 3; a++) node_pos[3 * i + a] = (float)snapshot[i]->coord[a];
            }
            if (on_device) {
                node_coord.resize(snapshot.size() * 3); node_radius.resize(snapshot.size()); node_valid.resize(snapshot.size());
                for (size_t i = 0; i < snapshot.size(); i++) {
                    for (int a = 0; a < 3; a++) node_coord[3 * i + a] = snapshot[i]->coord[a];
                    node_radius[i] = snapshot[i]->radius; node_valid[i] = snapshot[i]->valid ? 1 : 0;
                }
                DeviceBatchRequest req;
                req.engine_state = engineState(); req.goal_ratio = goal_ratio; req.inlier_ratio = inlier_ratio;
                const U *g[3] = { &rand_x, &rand_y, &rand_z }, *l[3] = { &rand_x_in, &rand_y_in, &rand_z_in };
                for (int a = 0; a < 3; a++) {
                    req.end_pt[a] = end_pt[a];
                    req.lo[a] = g[a]->a(); req.hi[a] = g[a]->b(); req.in_lo[a] = l[a]->a(); req.in_hi[a] = l[a]->b();
                }
                req.z_l = z_l; req.safety_margin = safety_margin; req.k = k_now; req.n_nodes = (int)snapshot.size();
                req.node_coord = node_coord.data(); req.node_radius = node_radius.data(); req.node_valid = node_valid.data();
                eng_.seed(device_batch_(req, centers, radii));
                n = (int)radii.size();
                cloud_queries += k_now;
                radius_calls++; device_batches++;
            } else if (snapshot_nearest_) {
                for (size_t i = 0; i < (size_t)k_now * 3; i++) sample_pos[i] = (float)samples[i];
                snapshot_nearest_(node_pos.data(), (int)snapshot.size(), sample_pos.data(), k_now, nearest_idx.data());
                node_tree_calls++;
            }
            for (int j = 0; j < k_now && !on_device; j++) {
                const double *s = &samples[(size_t)j * 3];
                RrtNode *nearest = snapshot_nearest_ ? (nearest_idx[(size_t)j] >= 0 ? node_list_[(size_t)nearest_idx[(size_t)j]] : nullptr) : findNearestVertex(s);
                if (!nearest || !nearest->valid) continue;
                steer(s, nearest, &centers[(size_t)n * 3]);
                n++;
            }
            if (n > 0 && !on_device) {
                radius_(centers.data(), n, radii.data());
                cloud_queries += n;
                radius_calls++;
            }
            const bool gpu_range = snapshot_range_ && n > 0;
            if (gpu_range) {
                // SURVEY 8f-2, second half: the 2 x radius neighbourhoods treeRewire asks the node tree for
                // (kd_nearest_rangef, corridor_finder.cpp:462-464) -- one batched range call against the same snapshot for all
                // candidates; the nodes this batch itself inserts are added from a short list below.  Invalidated nodes stay
                // in place until the batch is done (their snapshot entries must stay alive), as they do in the reference
                // until invalidSet reaches cach_size.
                for (int j = 0; j < n; j++) {
                    for (int a = 0; a < 3; a++) center_pos[(size_t)j * 3 + a] = (float)centers[(size_t)j * 3 + a];
                    ranges[(size_t)j] = (float)radii[(size_t)j] * 2.0f;
                }
                snapshot_range_(node_pos.data(), (int)snapshot.size(), center_pos.data(), ranges.data(), n, range_off, range_idx);
                node_tree_calls++;
                batch_nodes.clear();
            }
            defer_remove_ = true;          // with either provider, so that both see the same node tree
            for (int j = 0; j < n; j++) {
                // the tree may have grown since the snapshot: connect to the vertex that is nearest NOW
                RrtNode *nearest = findNearestVertex(&centers[(size_t)j * 3]);
                if (!nearest || !nearest->valid) continue;
                if (gpu_range) {
                    found.clear();
                    for (int64_t t = range_off[(size_t)j]; t < range_off[(size_t)j + 1]; t++) found.push_back(snapshot[(size_t)range_idx[(size_t)t]]);
                    const float *c = &center_pos[(size_t)j * 3];
                    const double r2 = (double)ranges[(size_t)j] * (double)ranges[(size_t)j];
                    for (RrtNode *b : batch_nodes) {              // inserted by this batch, in insertion order
                        double s2 = 0.0;
                        for (int a = 0; a < 3; a++) { const double d = (double)(float)b->coord[a] - (double)c[a]; s2 += d * d; }
                        if (s2 <= r2) found.push_back(b);
                    }
                    found_override_ = &found;
                }
                RrtNode *added = tryInsert(&centers[(size_t)j * 3], radii[(size_t)j], nearest);
                found_override_ = nullptr;
                if (gpu_range && added) batch_nodes.push_back(added);
            }
            defer_remove_ = false;
            if ((int)invalid_set_.size() >= cach_size) removeInvalid();
            it += k_now;
        }
        removeInvalid();
        tracePath();
        return it;
    }

    // SafeRegionEvaluate (corridor_finder.cpp:817-936): after a NEW cloud has arrived, lazily re-validate the nodes of the
    // current best path -- shrink radii, drop nodes that became too small or lost the connection to their parent /
    // children, fall back to the next feasible end -- until the best path is valid again or none is left.  A node's new
    // radius depends only on its centre, so each pass of the reference's loop is ONE batched radius call here (the
    // reference issues one cloud query per path node, :835).  The failed nodes are handed to treeRepair (:933).  Returns the
    // number of passes.
    int evaluate()
    {
        if (!path_exist_status) return 0;
        int passes = 0;
        std::vector<double> centers, radii;
        std::vector<FailedNode> fail_list;
        for (;;) {
            passes++;
            centers.clear();
            for (RrtNode *p : path_list_) if (p->pre) { centers.push_back(p->coord[0]); centers.push_back(p->coord[1]); centers.push_back(p->coord[2]); }
            radii.resize(centers.size() / 3);
            if (!radii.empty()) {
                radius_(centers.data(), (int)radii.size(), radii.data());
                cloud_queries += (int64_t)radii.size();
                radius_calls++;
            }
            size_t k = 0;
            for (RrtNode *ptr : path_list_) {
                RrtNode *pre = ptr->pre;
                if (!pre) continue;
                const double update_radius = radii[k++];
                const int ret = checkNodeUpdate(update_radius, ptr->radius);
                const float old_radius = ptr->radius;
                ptr->radius = (float)update_radius;
                if (ret == -1) {
                    ptr->valid = false; invalid_set_.push_back(ptr); clearBranchS(ptr);
                    fail_list.push_back(FailedNode(ptr->coord, old_radius));
                } else if (checkNodeRelation(dist(ptr->coord, pre->coord), ptr, pre) != -1) {
                    if (ptr->valid) {
                        ptr->valid = false; invalid_set_.push_back(ptr); clearBranchS(ptr);
                        fail_list.push_back(FailedNode(ptr->coord, old_radius));
                    }
                } else {
                    const std::vector<RrtNode *> children = ptr->nxt;
                    for (RrtNode *c : children) {
                        if (checkNodeRelation(dist(ptr->coord, c->coord), ptr, c) != -1 && c->valid) {
                            c->valid = false; invalid_set_.push_back(c); clearBranchS(c);
                            fail_list.push_back(FailedNode(c->coord, c->radius));
                        }
                    }
                }
            }
            bool all_valid = true;
            for (RrtNode *p : path_list_) all_valid = all_valid && p->valid;
            if (all_valid) break;
            std::vector<RrtNode *> feasible;
            for (RrtNode *e : end_list_) if (e->valid && checkEnd(e)) feasible.push_back(e);
            end_list_ = feasible;
            if (feasible.empty()) { path_exist_status = false; inform_status = false; best_distance = inf(); break; }
            best_end_ptr = feasible[0];
            double best_cost = inf();
            for (RrtNode *n : feasible) {
                const double cost = n->g + dist(n->coord, end_pt) + dist(root_node->coord, commit_root);
                if (cost < best_cost) { best_end_ptr = n; best_cost = cost; best_distance = best_cost; }
            }
            path_list_.clear();
            for (RrtNode *p = best_end_ptr; p; p = p->pre) path_list_.push_back(p);
        }
        removeInvalid();
        if (repair_after_evaluate) treeRepair(fail_list);
        tracePath();
        return passes;
    }

    // treeRepair (corridor_finder.cpp:938-1021): where nodes failed, their neighbours most likely fail too -- re-query every
    // valid node within 2 x radius of a failed node and invalidate what is too small or disconnected now.  A node's radius
    // depends only on its centre, so the neighbours of ALL failed nodes are gathered first and answered by ONE batched
    // radius call (the reference: one cloud query per neighbour, :973); the invalidation logic then runs in the
    // reference's order on the cached radii.
    struct FailedNode {
        double coord[3]; float radius;
        FailedNode(const double c[3], float r) : radius(r) { coord[0] = c[0]; coord[1] = c[1]; coord[2] = c[2]; }
    };
    void treeRepair(const std::vector<FailedNode> &fail_list)
    {
        std::vector<std::vector<RrtNode *>> near(fail_list.size());
        std::vector<RrtNode *> todo;
        for (size_t i = 0; i < fail_list.size(); i++) {
            const float pos[3] = { (float)fail_list[i].coord[0], (float)fail_list[i].coord[1], (float)fail_list[i].coord[2] };
            if (quirks_) node_tree_.rangeReferenceOrder(pos, fail_list[i].radius * 2.0f, near[i]);
            else node_tree_.range(pos, fail_list[i].radius * 2.0f, near[i]);
            for (RrtNode *p : near[i]) if (p != root_node && p->pre != root_node) todo.push_back(p);
        }
        std::sort(todo.begin(), todo.end(), [](const RrtNode *x, const RrtNode *y) { return x->serial < y->serial; });
        todo.erase(std::unique(todo.begin(), todo.end()), todo.end());
        std::vector<double> centers(todo.size() * 3), radii(todo.size());
        for (size_t i = 0; i < todo.size(); i++) for (int a = 0; a < 3; a++) centers[3 * i + a] = todo[i]->coord[a];
        if (!todo.empty()) {
            radius_(centers.data(), (int)todo.size(), radii.data());
            cloud_queries += (int64_t)todo.size();
            radius_calls++;
        }
        auto cached = [&](const RrtNode *p) {
            const size_t i = std::lower_bound(todo.begin(), todo.end(), p, [](const RrtNode *x, const RrtNode *y) { return x->serial < y->serial; }) - todo.begin();
            return radii[i];
        };
        for (size_t i = 0; i < fail_list.size(); i++) {
            for (RrtNode *ptr : near[i]) {
                if (!ptr->valid) continue;
                RrtNode *pre = ptr->pre;
                if (pre == root_node || ptr == root_node) continue;
                const double update_radius = cached(ptr);
                const int ret = checkNodeUpdate(update_radius, ptr->radius);
                ptr->radius = (float)update_radius;
                if (ret == -1) {
                    ptr->valid = false; invalid_set_.push_back(ptr); clearBranchS(ptr);
                    continue;
                }
                if (pre && checkNodeRelation(dist(pre->coord, ptr->coord), pre, ptr) != -1 && pre->valid) {
                    pre->valid = false; invalid_set_.push_back(pre); clearBranchS(pre);
                    continue;
                }
                const std::vector<RrtNode *> children = ptr->nxt;
                for (RrtNode *c : children) {
                    if (checkNodeRelation(dist(ptr->coord, c->coord), ptr, c) != -1 && c->valid) {
                        c->valid = false; invalid_set_.push_back(c); clearBranchS(c);
                    }
                }
            }
        }
        removeInvalid();
    }
    bool repair_after_evaluate = true;      // the reference always calls treeRepair from SafeRegionEvaluate (:933)

    // results (getPath, corridor_finder.h:131-134): centres (k x 3) and radii of the corridor spheres, root first
    std::vector<double> path;
    std::vector<double> radius;
    bool path_exist_status = true;
    size_t nodeCount() const { return node_list_.size(); }
    const std::vector<RrtNode *> &nodeList() const { return node_list_; }      // getTree (corridor_finder.h:136-139)
    int64_t cloud_queries = 0, radius_calls = 0, node_tree_calls = 0;

    // optional: batched nearest-vertex provider for the snapshot phase of expandBatched / refineBatched.
    // node_pos: n_nodes x 3 float centres in node-list order; out_nearest[j] = index of the node nearest to sample j.
    using SnapshotNearestFn = std::function<void(const float *node_pos, int n_nodes, const float *samples, int k, int32_t *out_nearest)>;
    void setSnapshotNearest(SnapshotNearestFn fn) { snapshot_nearest_ = std::move(fn); }
    // optional: batched range provider for the insertion phase of expandBatched / refineBatched: for candidate j all snapshot
    // nodes with d2 <= ranges[j]^2 (fp64 on the float32 positions, inclusive: kd_nearest_rangef), CSR lists of node indices in
    // ascending order
    using SnapshotRangeFn = std::function<void(const float *node_pos, int n_nodes, const float *centers, const float *ranges, int k,
                                               std::vector<int64_t> &offsets, std::vector<int32_t> &idx)>;
    void setSnapshotRange(SnapshotRangeFn fn) { snapshot_range_ = std::move(fn); }
    const NodeKdTree &nodeTree() const { return node_tree_; }
    // optional: the whole snapshot phase of a batch behind one call.  The provider gets genSample's state (engine state and
    // distribution bounds: corridor_finder.cpp:333-358 with setPt :52-91) and the frozen node set, draws the next k samples of
    // that stream itself, and returns the centres / radii of the candidates the loop would keep (nearest vertex valid, centre
    // z >= z_l, radius >= safety_margin) in sample order, plus the engine state behind the k-th sample.
    struct DeviceBatchRequest {
        uint32_t engine_state;
        double goal_ratio, inlier_ratio, end_pt[3], lo[3], hi[3], in_lo[3], in_hi[3], z_l, safety_margin;
        int k, n_nodes;
        const double *node_coord;      // n_nodes x 3
        const float *node_radius;
        const uint8_t *node_valid;
    };
    using DeviceBatchFn = std::function<uint32_t(const DeviceBatchRequest &req, std::vector<double> &centers, std::vector<double> &radii)>;
    void setDeviceBatch(DeviceBatchFn fn) { device_batch_ = std::move(fn); }
    int64_t device_batches = 0;
    // state of the engine: the next draw is 16807 * state mod 2^31-1 (operator<< of linear_congruential_engine writes it)
    uint32_t engineState() const { std::ostringstream os; os << eng_; return (uint32_t)std::stoul(os.str()); }

    double safety_margin = 0, search_margin = 0, max_radius = 0, sample_range = 0;

private:
    using U = std::uniform_real_distribution<double>;
    static double inf() { return 9999999.0; }             // data_type.h:6
    static double dist(const double a[3], const double b[3])
    {
        return std::sqrt((a[0] - b[0]) * (a[0] - b[0]) + (a[1] - b[1]) * (a[1] - b[1]) + (a[2] - b[2]) * (a[2] - b[2]));
    }

    void initRoot()
    {
        node_tree_.clear();
        for (int a = 0; a < 3; a++) commit_root[a] = start_pt[a];
        double r;
        radius_(start_pt, 1, &r);
        cloud_queries++;
        root_node = new RrtNode(start_pt, (float)r, 0.0f, (float)min_distance);
        recordNode(root_node);
        insertIntoTree(root_node);
    }

    void insertIntoTree(RrtNode *n)
    {
        const float pos[3] = { (float)n->coord[0], (float)n->coord[1], (float)n->coord[2] };
        n->serial = serial_++;
        node_tree_.insert(pos, n);
    }
    void recordNode(RrtNode *n) { node_list_.push_back(n); }

    // corridor_finder.cpp:333-383
    void genSample(double pt[3])
    {
        const double bias = rand_bias(eng_);
        if (bias <= goal_ratio) { pt[0] = end_pt[0]; pt[1] = end_pt[1]; pt[2] = end_pt[2]; return; }
        if (!inform_status) {
            if (bias > goal_ratio && bias <= (goal_ratio + inlier_ratio)) { pt[0] = rand_x_in(eng_); pt[1] = rand_y_in(eng_); pt[2] = rand_z_in(eng_); }
            else { pt[0] = rand_x(eng_); pt[1] = rand_y(eng_); pt[2] = rand_z(eng_); }
        } else {
            const double us = rand_u(eng_), vs = rand_v(eng_), phis = rand_phi(eng_);
            const double as = elli_l / 2.0 * std::cbrt(us), bs = elli_s / 2.0 * std::cbrt(us);
            const double thetas = std::acos(1 - 2 * vs);
            const double l[3] = { as * std::sin(thetas) * std::cos(phis), bs * std::sin(thetas) * std::sin(phis), bs * std::cos(thetas) };
            for (int a = 0; a < 3; a++) pt[a] = rot[a][0] * l[0] + rot[a][1] * l[1] + rot[a][2] * l[2] + translation[a];
            pt[0] = std::min(std::max(pt[0], x_l), x_h);
            pt[1] = std::min(std::max(pt[1], y_l), y_h);
            pt[2] = std::min(std::max(pt[2], z_l), z_h);
        }
    }

    RrtNode *findNearestVertex(const double pt[3])
    {
        const float pos[3] = { (float)pt[0], (float)pt[1], (float)pt[2] };
        return node_tree_.nearest(pos);
    }

    // the steering half of genNewNode (corridor_finder.cpp:387-402)
    static void steer(const double s[3], const RrtNode *nearest, double center[3])
    {
        const double dis = dist(nearest->coord, s);
        if (dis > nearest->radius) {
            const double steer_dis = nearest->radius / dis;
            for (int a = 0; a < 3; a++) center[a] = nearest->coord[a] + (s[a] - nearest->coord[a]) * steer_dis;
        } else {
            for (int a = 0; a < 3; a++) center[a] = s[a];
        }
    }

    bool checkEnd(const RrtNode *n) const { return dist(n->coord, end_pt) + 0.1 < n->radius; }

    // the body of the expansion loop after the cloud query (corridor_finder.cpp:730-755)
    RrtNode *tryInsert(const double c[3], double r, RrtNode *nearest)
    {
        if (c[2] < z_l || (float)r < safety_margin) return nullptr;
        RrtNode *n = new RrtNode(c, (float)r, (float)inf(), (float)dist(c, end_pt));
        treeRewire(n, nearest);
        if (!n->valid) { delete n; return nullptr; }  // (the reference leaks these nodes)
        if (checkEnd(n)) {
            if (!inform_status) best_end_ptr = n;
            end_list_.push_back(n);
            if (!(quirks_ && in_expansion_)) updateHeuristicRegion(n);      // commented out in SafeRegionExpansion (:742)
            inform_status = true;
        }
        insertIntoTree(n);
        recordNode(n);
        treePrune(n);
        if ((int)invalid_set_.size() >= cach_size && !defer_remove_) removeInvalid();
        return n;
    }

    // corridor_finder.cpp:661-667
    int checkNodeUpdate(double new_radius, double old_radius) const
    {
        if (new_radius < safety_margin) return -1;
        return new_radius < old_radius ? 0 : 1;
    }

    static int checkNodeRelation(double dis, const RrtNode *n1, const RrtNode *n2)
    {
        if ((dis + n2->radius) == n1->radius) return 1;
        if ((dis + 0.1) < 0.95 * (n1->radius + n2->radius)) return -1;
        return 0;
    }

    static bool isSuccessor(const RrtNode *cur, const RrtNode *near_)
    {
        for (const RrtNode *p = near_ ? near_->pre : nullptr; p; p = p->pre) if (p == cur) return true;
        return false;
    }

    // corridor_finder.cpp:458-571
    void treeRewire(RrtNode *newPtr, RrtNode *nearestPtr)
    {
        const float range = newPtr->radius * 2.0f;
        const float pos[3] = { (float)newPtr->coord[0], (float)newPtr->coord[1], (float)newPtr->coord[2] };
        std::vector<RrtNode *> found;
        if (found_override_) found = *found_override_;          // answered by the batched snapshot range provider
        else if (quirks_) node_tree_.rangeReferenceOrder(pos, range, found);
        else node_tree_.range(pos, range, found);
        std::vector<RrtNode *> nearPtrList;
        bool isInvalid = false;
        for (RrtNode *nearPtr : found) {
            const double dis = dist(nearPtr->coord, newPtr->coord);
            const int res = checkNodeRelation(dis, nearPtr, newPtr);
            nearPtr->rel_id = res; nearPtr->rel_dis = (float)dis;
            nearPtrList.push_back(nearPtr);
            if (res == 1) { newPtr->valid = false; isInvalid = true; break; }
        }
        if (isInvalid) {
            for (RrtNode *p : nearPtrList) { p->rel_id = -2; p->rel_dis = -1.0f; }
            return;
        }
        double min_cost = nearestPtr->g + dist(nearestPtr->coord, newPtr->coord);
        newPtr->pre = nearestPtr;
        newPtr->g = (float)min_cost;
        nearestPtr->nxt.push_back(newPtr);
        RrtNode *lstParentPtr = nearestPtr;
        std::vector<RrtNode *> nearVertex;
        for (RrtNode *nearPtr : nearPtrList) {
            const int res = nearPtr->rel_id;
            const double dis = nearPtr->rel_dis;
            const double cost = nearPtr->g + dis;
            if (res == -1) {
                if (cost < min_cost) {
                    min_cost = cost;
                    newPtr->pre = nearPtr;
                    newPtr->g = (float)min_cost;
                    lstParentPtr->nxt.pop_back();
                    lstParentPtr = nearPtr;
                    lstParentPtr->nxt.push_back(newPtr);
                }
                nearVertex.push_back(nearPtr);
            }
            nearPtr->rel_id = -2; nearPtr->rel_dis = -1.0f;
        }
        for (RrtNode *nearPtr : nearVertex) {
            if (!nearPtr->valid) continue;
            const double dis = dist(nearPtr->coord, newPtr->coord);
            const double cost = dis + newPtr->g;
            if (cost < nearPtr->g) {
                if (isSuccessor(nearPtr, newPtr->pre)) continue;
                if (nearPtr->pre == nullptr) {
                    nearPtr->pre = newPtr; nearPtr->g = (float)cost;
                } else {
                    RrtNode *lstNearParent = nearPtr->pre;
                    nearPtr->pre = newPtr; nearPtr->g = (float)cost;
                    auto &ch = lstNearParent->nxt;
                    ch.erase(std::remove(ch.begin(), ch.end(), nearPtr), ch.end());
                }
                newPtr->nxt.push_back(nearPtr);
            }
        }
    }

    void clearBranchS(RrtNode *n)
    {
        for (RrtNode *c : n->nxt) {
            if (c->valid) invalid_set_.push_back(c);
            c->valid = false;
            clearBranchS(c);
        }
    }

    void treePrune(RrtNode *n)
    {
        // the reference compares the float sum g + f with the double best_distance (:165); for the end node that has just set
        // best_distance the two differ only by float rounding, which prunes it half of the time -- hence the small slack
        const bool prune = quirks_ ? (double)(n->g + n->f) > best_distance : (double)n->g + (double)n->f > best_distance + 1e-4;
        if (prune) { n->valid = false; invalid_set_.push_back(n); clearBranchS(n); }
    }

    // corridor_finder.cpp:170-232
    void removeInvalid()
    {
        std::vector<RrtNode *> keepNodes, keepEnds;
        node_tree_.clear();
        for (RrtNode *n : node_list_) {
            if (n->valid) {
                const float pos[3] = { (float)n->coord[0], (float)n->coord[1], (float)n->coord[2] };
                node_tree_.insert(pos, n);
                keepNodes.push_back(n);
                if (checkEnd(n)) keepEnds.push_back(n);
            }
        }
        node_list_ = keepNodes;
        end_list_ = keepEnds;
        for (RrtNode *n : invalid_set_) {
            if (n->pre != nullptr) {
                auto &ch = n->pre->nxt;
                ch.erase(std::remove(ch.begin(), ch.end(), n), ch.end());
            }
        }
        for (RrtNode *n : invalid_set_)
            for (RrtNode *c : n->nxt) if (c->valid) c->pre = nullptr;
        std::sort(invalid_set_.begin(), invalid_set_.end());
        invalid_set_.erase(std::unique(invalid_set_.begin(), invalid_set_.end()), invalid_set_.end());
        for (RrtNode *n : invalid_set_) { if (n == best_end_ptr) best_end_ptr = nullptr; delete n; }
        invalid_set_.clear();
    }

    void updateEllipsoidFrame(const double from[3])
    {
        double xtf[3], ytf[3], ztf[3];
        const double down[3] = { 0, 0, -1 };
        for (int a = 0; a < 3; a++) translation[a] = (from[a] + end_pt[a]) / 2.0;
        double nrm = 0;
        for (int a = 0; a < 3; a++) { xtf[a] = end_pt[a] - translation[a]; nrm += xtf[a] * xtf[a]; }
        nrm = std::sqrt(nrm);
        for (int a = 0; a < 3; a++) xtf[a] = nrm > 0 ? xtf[a] / nrm : 0.0;
        ytf[0] = xtf[1] * down[2] - xtf[2] * down[1]; ytf[1] = xtf[2] * down[0] - xtf[0] * down[2]; ytf[2] = xtf[0] * down[1] - xtf[1] * down[0];
        nrm = std::sqrt(ytf[0] * ytf[0] + ytf[1] * ytf[1] + ytf[2] * ytf[2]);
        for (int a = 0; a < 3; a++) ytf[a] = nrm > 0 ? ytf[a] / nrm : 0.0;
        ztf[0] = xtf[1] * ytf[2] - xtf[2] * ytf[1]; ztf[1] = xtf[2] * ytf[0] - xtf[0] * ytf[2]; ztf[2] = xtf[0] * ytf[1] - xtf[1] * ytf[0];
        for (int a = 0; a < 3; a++) { rot[a][0] = xtf[a]; rot[a][1] = ytf[a]; rot[a][2] = ztf[a]; }
    }

    // corridor_finder.cpp:298-330
    void updateHeuristicRegion(RrtNode *end_node)
    {
        const double update_cost = end_node->g + dist(end_node->coord, end_pt) + dist(root_node->coord, commit_root);
        if (update_cost < best_distance) {
            best_distance = update_cost;
            elli_l = best_distance;
            elli_s = std::sqrt(std::max(0.0, best_distance * best_distance - min_distance * min_distance));
            if (inform_status) for (RrtNode *p : node_list_) p->best = false;
            for (RrtNode *p = end_node; p; p = p->pre) p->best = true;
            best_end_ptr = end_node;
        }
    }

    bool checkValidEnd(const RrtNode *end) const
    {
        for (const RrtNode *p = end; p; p = p->pre) {
            if (!p->valid) return false;
            if (dist(p->coord, root_node->coord) < p->radius) return true;
        }
        return false;
    }

    // corridor_finder.cpp:578-641
    void tracePath()
    {
        std::vector<RrtNode *> feasible;
        for (RrtNode *e : end_list_)
            if (checkValidEnd(e) && checkEnd(e) && e->valid) feasible.push_back(e);
        path.clear(); radius.clear(); path_list_.clear();
        if (feasible.empty()) {
            path_exist_status = false; best_distance = inf(); inform_status = false; end_list_.clear();
            return;
        }
        end_list_ = feasible;
        best_end_ptr = feasible[0];
        double best_cost = inf();
        for (RrtNode *n : feasible) {
            const double cost = n->g + dist(n->coord, end_pt) + dist(root_node->coord, commit_root);
            if (cost < best_cost) { best_end_ptr = n; best_cost = cost; best_distance = best_cost; }
        }
        for (RrtNode *p = best_end_ptr; p; p = p->pre) path_list_.push_back(p);
        for (auto it = path_list_.rbegin(); it != path_list_.rend(); ++it) {
            path.push_back((*it)->coord[0]); path.push_back((*it)->coord[1]); path.push_back((*it)->coord[2]);
            radius.push_back((*it)->radius);
        }
        path_exist_status = true;
    }

    void destroyTree()
    {
        node_tree_.clear();
        for (RrtNode *n : node_list_) delete n;
        node_list_.clear();
    }

    RadiusBatchFn radius_;
    SnapshotNearestFn snapshot_nearest_;
    SnapshotRangeFn snapshot_range_;
    DeviceBatchFn device_batch_;
    const std::vector<RrtNode *> *found_override_ = nullptr;
    bool defer_remove_ = false;
    std::default_random_engine eng_;
    U rand_x, rand_y, rand_z, rand_bias, rand_x_in, rand_y_in, rand_z_in;
    U rand_u = U(0.0, 1.0), rand_v = U(0.0, 1.0), rand_phi = U(0.0, 2 * M_PI);
    double start_pt[3] = { 0, 0, 0 }, end_pt[3] = { 0, 0, 0 }, commit_root[3] = { 0, 0, 0 };
    double x_l = 0, x_h = 0, y_l = 0, y_h = 0, z_l = 0, z_h = 0;
    double min_distance = 0, best_distance = inf(), elli_l = 0, elli_s = 0;
    double translation[3] = { 0, 0, 0 }, rot[3][3] = { { 1, 0, 0 }, { 0, 1, 0 }, { 0, 0, 1 } };
    double inlier_ratio = 0, goal_ratio = 0;
    int max_samples = 0, cach_size = 10;
    bool inform_status = false, quirks_ = false, in_expansion_ = false;
    RrtNode *root_node = nullptr, *best_end_ptr = nullptr;
    std::vector<RrtNode *> node_list_, end_list_, invalid_set_, path_list_;
    NodeKdTree node_tree_;
    int64_t serial_ = 0;
};

}  // namespace pc
#endif  // PC_RRT_HPP_
