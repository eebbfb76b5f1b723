/*
 * pc_kdtree_compat.h -- the call surface of the reference's Utils/kdtree (3-D) on top of the GPU index.
 *
 * Every function mirrors the reference function of the same name without the `pc` prefix
 * (Utils/kdtree/include/kdtree/kdtree.h:39-122; implementation Utils/kdtree/src/kdtree.c): same arguments, same
 * return conventions (0 / -1, NULL on allocation failure or on an empty tree for the nearest query, result sets
 * owned by the caller and released with pckd_res_free).  A caller switches by including this header and either
 * renaming its calls or defining PC_KDTREE_COMPAT_RENAME before the include, which maps the kd_* names onto these.
 *
 * What differs, by design:
 *   - k must be 3 (pckd_create returns NULL otherwise); positions are float32 on the GPU, which is the planner's
 *     cloud type (pcl::PointXYZ).  A double coordinate that is not exactly representable in float32 is rejected
 *     (insert returns -1, queries return NULL) instead of being silently rounded.
 *   - inserts are buffered on the host; the GPU index is rebuilt on the first query after a modification
 *     (kd_insert is O(depth) per call in the reference; here a rebuild is one pc_index_build, ~0.1 ms per 300k points).
 *   - ties: among points at exactly the same fp64 distance the LOWEST insertion index is returned (the reference
 *     returns the one its DFS meets first, kdtree.c:383,433); range result sets iterate in insertion order (the
 *     reference iterates in reverse visit order, kdtree.c:810-828) and contain every point with d2 <= range^2
 *     (the reference can miss a point at exactly `range`, kdtree.c:283).
 *   - one query per call is latency-bound (a kernel launch per call); the batch entry points at the end of this
 *     header are what the planner's loops should use.
 *   - pckd_res_item3 / pckd_res_item3f implement the documented behaviour (they fill x, y, z and return the data
 *     pointer); the reference's versions test *x instead of x and return 0 (kdtree.c:666-684).
 */
#ifndef PC_KDTREE_COMPAT_H_
#define PC_KDTREE_COMPAT_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

struct pckd_tree;
struct pckd_res;

struct pckd_tree *pckd_create(int k);                                   /* kd_create            kdtree.c:112 */
void pckd_free(struct pckd_tree *tree);                                 /* kd_free              kdtree.c:128 */
void pckd_clear(struct pckd_tree *tree);                                /* kd_clear             kdtree.c:150 */
void pckd_data_destructor(struct pckd_tree *tree, void (*destr)(void *)); /* kd_data_destructor kdtree.c:161 */

int pckd_insert(struct pckd_tree *tree, const double *pos, void *data);  /* kd_insert           kdtree.c:196 */
int pckd_insertf(struct pckd_tree *tree, const float *pos, void *data);  /* kd_insertf          kdtree.c:211 */
int pckd_insert3(struct pckd_tree *tree, double x, double y, double z, void *data);  /* kd_insert3  kdtree.c:244 */
int pckd_insert3f(struct pckd_tree *tree, float x, float y, float z, void *data);    /* kd_insert3f kdtree.c:253 */

struct pckd_res *pckd_nearest(struct pckd_tree *tree, const double *pos);             /* kd_nearest   kdtree.c:404 */
struct pckd_res *pckd_nearestf(struct pckd_tree *tree, const float *pos);             /* kd_nearestf  kdtree.c:459 */
struct pckd_res *pckd_nearest3(struct pckd_tree *tree, double x, double y, double z); /* kd_nearest3  kdtree.c:493 */
struct pckd_res *pckd_nearest3f(struct pckd_tree *tree, float x, float y, float z);   /* kd_nearest3f kdtree.c:502 */

struct pckd_res *pckd_nearest_range(struct pckd_tree *tree, const double *pos, double range);             /* kdtree.c:537 */
struct pckd_res *pckd_nearest_rangef(struct pckd_tree *tree, const float *pos, float range);              /* kdtree.c:561 */
struct pckd_res *pckd_nearest_range3(struct pckd_tree *tree, double x, double y, double z, double range); /* kdtree.c:595 */
struct pckd_res *pckd_nearest_range3f(struct pckd_tree *tree, float x, float y, float z, float range);    /* kdtree.c:604 */

void pckd_res_free(struct pckd_res *set);      /* kd_res_free   kdtree.c:613 */
int pckd_res_size(struct pckd_res *set);       /* kd_res_size   kdtree.c:620 */
void pckd_res_rewind(struct pckd_res *set);    /* kd_res_rewind kdtree.c:625 */
int pckd_res_end(struct pckd_res *set);        /* kd_res_end    kdtree.c:630 */
int pckd_res_next(struct pckd_res *set);       /* kd_res_next   kdtree.c:635 */
void *pckd_res_item(struct pckd_res *set, double *pos);                         /* kd_res_item   kdtree.c:641 */
void *pckd_res_itemf(struct pckd_res *set, float *pos);                         /* kd_res_itemf  kdtree.c:652 */
void *pckd_res_item3(struct pckd_res *set, double *x, double *y, double *z);    /* kd_res_item3  kdtree.c:666 */
void *pckd_res_item3f(struct pckd_res *set, float *x, float *y, float *z);      /* kd_res_item3f kdtree.c:676 */
void *pckd_res_item_data(struct pckd_res *set);                                 /* kd_res_item_data kdtree.c:686 */

/* ---- batch extensions (what the planner's loops should call) ------------------------------------------- */
/* m float32 queries (stride 3 or 4 floats, host memory): out_index[k] = insertion index of the nearest point
 * (-1 on an empty tree), out_data[k] = its data pointer (nullable), out_d2[k] = squared distance (nullable). */
int pckd_nearest_batchf(struct pckd_tree *tree, const float *q_xyz, int64_t m, int64_t q_stride,
                        int32_t *out_index, void **out_data, float *out_d2);
/* the pc_index handle behind the tree (rebuilds it first if inserts are pending); NULL on failure.  Gives access to
 * pc_radius_batch / pc_range_batch / pc_clearance_batch of include/pc_index.h on the same cloud. */
struct pc_index *pckd_index(struct pckd_tree *tree);
int64_t pckd_size(struct pckd_tree *tree);
const char *pckd_last_error(struct pckd_tree *tree);

#ifdef PC_KDTREE_COMPAT_RENAME
#define kdtree pckd_tree
#define kdres pckd_res
#define kd_create pckd_create
#define kd_free pckd_free
#define kd_clear pckd_clear
#define kd_data_destructor pckd_data_destructor
#define kd_insert pckd_insert
#define kd_insertf pckd_insertf
#define kd_insert3 pckd_insert3
#define kd_insert3f pckd_insert3f
#define kd_nearest pckd_nearest
#define kd_nearestf pckd_nearestf
#define kd_nearest3 pckd_nearest3
#define kd_nearest3f pckd_nearest3f
#define kd_nearest_range pckd_nearest_range
#define kd_nearest_rangef pckd_nearest_rangef
#define kd_nearest_range3 pckd_nearest_range3
#define kd_nearest_range3f pckd_nearest_range3f
#define kd_res_free pckd_res_free
#define kd_res_size pckd_res_size
#define kd_res_rewind pckd_res_rewind
#define kd_res_end pckd_res_end
#define kd_res_next pckd_res_next
#define kd_res_item pckd_res_item
#define kd_res_itemf pckd_res_itemf
#define kd_res_item3 pckd_res_item3
#define kd_res_item3f pckd_res_item3f
#define kd_res_item_data pckd_res_item_data
#endif

#ifdef __cplusplus
}
#endif
#endif /* PC_KDTREE_COMPAT_H_ */
