/* clpt_host.h -- host-side scene preparation that feeds the render boundary.
 *
 * Plain C11.  Everything here runs on the CPU and produces exactly the host
 * data the reference hands to CLSetMeshes / CLSetCameraMatrix, in the
 * reference's wire layout (clpt_types.h).  Each entry point names the
 * reference interface it stands in for (paths relative to the reference tree).
 *
 * All symbols are exported from libclpt.so.
 */
#ifndef CLPT_HOST_H
#define CLPT_HOST_H

#include "clpt_types.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ lists
 * Fat-pointer byte vectors: a {capacity, length} header of two size_t sits
 * immediately before the data pointer.  Sizes are BYTES.  This is how sizes
 * cross the CLState boundary (CLSetMeshes reads list_size() of each member).
 * Replaces include/list.h:10-35, src/list.c:27-111.
 */
extern size_t LIST_INDEX;
void *new_list(size_t capacity_bytes);            /* list.h:21 */
void *init_list(size_t count, size_t elem_size);  /* list.h:23: length = capacity = count*elem_size, data uninitialised */
void *copy_list(const void *list);                /* list.h:25 */
void delete_list(void *list);                     /* list.h:27: NULL is a no-op */
size_t list_grow(void **list_ptr, size_t bytes);  /* list.h:32: returns old_length/bytes */
size_t list_size(const void *list);               /* list.h:34: length in bytes */
void list_concat(void **list1_ptr, const void *list2); /* list.h:19 */

#define vector_append(vec, item)                                         \
    (LIST_INDEX = list_grow((void **)&(vec), sizeof((item))),            \
     (vec)[LIST_INDEX] = (item))
#define vector_length(vec) (list_size(vec) / sizeof(*(vec)))

/* ------------------------------------------------------------------ vec3
 * Replaces include/vector.h:26-54, src/vector.c.  fp32, evaluated left to
 * right without FMA contraction; vec_length takes sqrt in double like the
 * reference (src/vector.c:16-18).
 */
#define Vector3(x, y, z) ((Vector3){ { (x), (y), (z) } })
#define Vector4(x, y, z, w) ((Vector4){ { (x), (y), (z), (w) } })
#define vec_x(v) ((v).s[0])
#define vec_y(v) ((v).s[1])
#define vec_z(v) ((v).s[2])
vec_t vec_dot(Vector3, Vector3);
vec_t vec_length_squared(Vector3);
vec_t vec_length(Vector3);
Vector3 vec_normalized(Vector3);
Vector3 *vec_normalize(Vector3 *);
Vector3 vec_add(Vector3, Vector3);
Vector3 vec_subtract(Vector3, Vector3);
Vector3 vec_cross(Vector3, Vector3);
Vector3 vec_negated(Vector3);
Vector3 *vec_negate(Vector3 *);
Vector3 vec_scaled(Vector3, vec_t);
Vector3 *vec_scale(Vector3 *, vec_t);
Vector3 vec_divide(Vector3, Vector3);
Vector3 vec_min(Vector3, Vector3);
Vector3 vec_max(Vector3, Vector3);

/* ------------------------------------------------------------------ 4x4
 * Replaces include/matrix.h:27-39, src/matrix.c.  Row-major rows[r].s[c];
 * mat_set/mat_get take (column n, row m) like the reference (matrix.c:5-13).
 */
void mat_set(Matrix *, unsigned int n, unsigned int m, vec_t value);
vec_t mat_get(Matrix, unsigned int n, unsigned int m);
Matrix mat_add(Matrix, Matrix);
Matrix mat_multiply(Matrix, Matrix);
Matrix mat_scaled(Matrix, vec_t);
Matrix *mat_scale(Matrix *, vec_t);
Matrix mat_inverse(Matrix, int *err); /* cofactor inverse; det==0 -> zero matrix, *err=1 */

/* ------------------------------------------------------------------ camera
 * cam_matrix = inverse(device(height/2) * projection(near,far,fov) * view).
 * Replaces include/camera.h:14-15, src/camera.c:62-70.
 * cam_matrix_ptr is the same call with pointer arguments for FFI callers. */
Matrix cam_matrix(Camera, int height);
void cam_matrix_ptr(const Camera *cam, int height, Matrix *out);

/* ------------------------------------------------------------------ kd-tree
 * build_kd: binned surface-area build + ropes, preorder 68-byte nodes.
 * Replaces include/kd_tree.h:52-59, src/kd_tree.c:202-320.  `tris` holds one
 * cl_int3 {v, vn, vt} per triangle CORNER (3 per triangle).  Takes ownership
 * of the three input lists (they become members of the returned kd).  When
 * path != NULL also writes "<path>.kd".
 *
 * build_kd(...) == build_kd_ex(..., KD_REF_DEPTH, KD_REF_NBINS) and is
 * byte-identical to the reference builder's node array / tri_indices.
 * build_kd_ex generalises depth and bin count at run time (the reference
 * hard-codes them, src/kd_tree.c:8-9); it is multi-threaded (OpenMP).
 */
#define KD_REF_DEPTH 15
#define KD_REF_NBINS 25
kd build_kd(cl_int3 *tris, Vector3 *verts, Vector3 *norms, const char *path);
kd build_kd_ex(cl_int3 *tris, Vector3 *verts, Vector3 *norms, const char *path,
               int depth, int nbins);
/* Extension (no reference counterpart): standard surface-area heuristic with a
 * leaf-cost termination; candidate planes are all triangle bounds (nbins <= 0),
 * or `nbins` uniform planes per axis in cells of more than 48 triangles
 * (nbins > 0).  Same wire format; ropes pushed down further than the reference
 * does.  Suggested: max_depth 8 + 1.3*log2(triangles), nbins 0, traversal_cost 1,
 * intersect_cost 1, empty_bonus 0.9. */
kd build_kd_sah(cl_int3 *tris, Vector3 *verts, Vector3 *norms, const char *path,
                int max_depth, int nbins, float traversal_cost, float intersect_cost,
                float empty_bonus);
int parse_kd(const char *filename, kd *tree);  /* kd_tree.h:55; returns 0 on success, 1 on I/O error */
int write_kd(const char *filename, const kd *tree); /* the writer half of src/kd_tree.c:239-274 */
void delete_kd(kd tree);                       /* kd_tree.h:58 */

/* Statistics of a built tree (the reference prints the first three at
 * src/kd_tree.c:232-235). */
typedef struct kd_stats {
    long long leaf_tri_refs;
    long long leaf_count;
    long long empty_leaves;
    long long node_count;
    int max_leaf_tris;
    int max_depth;
} kd_stats;
void kd_get_stats(const kd *tree, kd_stats *out);

/* ------------------------------------------------------------------ models
 * LoadModel: ".obj" -> parse + build_kd (also writes "<stem>.kd");
 * ".kd" -> parse_kd.  Returns 0 on success, 1 on failure.
 * Replaces include/model.h:6-7, src/model.c:147-176.  The OBJ reader handles
 * v / vn / vt / f (fan triangulation, negative = relative indices, a missing
 * vn/vt index becomes a negative number like tinyobj's invalid marker). */
int LoadModel(const char *filename, kd *model);
int load_obj_lists(const char *filename, Vector3 **verts, Vector3 **norms, cl_int3 **tris);
int write_obj(const char *filename, const Vector3 *verts, const Vector3 *norms, const cl_int3 *tris);
void kd_set_build_params(int depth, int nbins); /* depth/nbins used by LoadModel's build */
/* build_kd_sah: re-derive a straddling triangle's bounds from the triangle clipped to
 * each child cell ("perfect splits").  Default on: 15-20% fewer triangle tests on
 * scenes with large or thin triangles, ~10% more build time.  Off for per-frame rebuilds. */
void kd_set_sah_clip(int enable);

/* ------------------------------------------------------------------ physics
 * Explicit Euler pos += vel*dt over registered pointer pairs.
 * Replaces include/physics.h:6-13, src/physics.c (drives the animated config). */
void AddPhysObject(Vector3 *position, Vector3 *velocity);
void PhysStep(double stepSize);
void PhysTerminate(void);

#ifdef __cplusplus
}
#endif

#endif /* CLPT_HOST_H */
