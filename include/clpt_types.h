/* clpt_types.h -- wire types of the render boundary.
 *
 * These are the host-side layouts that cross the CLState.h boundary of
 * taylor-santos/CLPathTracer.  The reference gets them from <CL/cl_gl.h>
 * (cl_float3 == cl_float4 == 16 B, cl_int3 == 16 B); this tree has no OpenCL
 * headers, so layout-compatible types are declared here.  If a real <CL/cl.h>
 * was included first, its types are used instead, so a reference host file can
 * include this header unchanged.
 *
 *   type      size  reference
 *   Vector3   16    include/vector.h:12   (cl_float3, .s[4], 16-aligned)
 *   Vector4   16    include/vector.h:13
 *   cl_int3   16    include/kd_tree.h:15  (one per triangle CORNER: {v, vn, vt, pad})
 *   Matrix    64    include/matrix.h:10-12 (row-major rows[r].s[c])
 *   kdnode    68    include/kd_tree.h:31-50 (packed, stride 68)
 *   Object    24    include/object.h:9-21  (packed)
 *   Camera    48    include/camera.h:6-12
 *   kd        40    include/kd_tree.h:10-16 (five list.c fat pointers)
 */
#ifndef CLPT_TYPES_H
#define CLPT_TYPES_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef vec_t
#define vec_t float
#endif

#if !defined(CL_VERSION_1_0) && !defined(__OPENCL_CL_H)
typedef int32_t cl_int;
typedef uint32_t cl_uint;
typedef float cl_float;
typedef union clpt_float4 {
    cl_float s[4];
} __attribute__((aligned(16))) cl_float4;
typedef cl_float4 cl_float3;
typedef union clpt_int4 {
    cl_int s[4];
} __attribute__((aligned(16))) cl_int4;
typedef cl_int4 cl_int3;
#endif

typedef cl_float3 Vector3;
typedef cl_float4 Vector4;
typedef cl_int kd_index;
typedef unsigned int GLuint_t; /* GLuint of CLCreateImage(GLuint), include/CLState.h:26 */

typedef struct Matrix {
    Vector4 rows[4];
} Matrix;

typedef struct Camera {
    vec_t Near;
    vec_t Far;
    vec_t FOV;
    Vector3 Position;
    Vector3 Forward;
} Camera;

typedef enum KD_AXIS { KD_X = 0, KD_Y = 1, KD_Z = 2 } KD_AXIS;

/* Face numbering of a cell; face/2 is the axis, face&1 is the max side. */
typedef enum KD_SIDE {
    KD_LEFT = 0,
    KD_RIGHT = 1,
    KD_DOWN = 2,
    KD_UP = 3,
    KD_BACK = 4,
    KD_FRONT = 5
} KD_SIDE;

enum { KD_SPLIT = 0, KD_LEAF = 1 };

#pragma GCC diagnostic push
#pragma GCC diagnostic ignored "-Wpragmas"
#pragma GCC diagnostic ignored "-Wpacked-not-aligned"
#pragma pack(push, 1)
typedef struct kdnode {
    Vector4 min, max;       /*  0, 16 */
    cl_int type;            /* 32: KD_SPLIT / KD_LEAF */
    union {
        struct {
            vec_t value;    /* 36 */
            cl_int axis;    /* 40 */
            kd_index children[2]; /* 44, 48 */
        } split;
        struct {
            kd_index tris;      /* 36: offset into tri_indices */
            kd_index tri_count; /* 40 */
            kd_index ropes[6];  /* 44..67: neighbour node per KD_SIDE, -1 = outside */
        } leaf;
    };
} kdnode;

typedef struct Object {
    Vector3 position;
    cl_int type; /* OBJ_SPHERE = 0 */
    union {
        struct {
            vec_t radius;
        } sphere;
    };
} Object;
#pragma pack(pop)
#pragma GCC diagnostic pop

enum { OBJ_SPHERE = 0 };

/* A loaded model: five list.c-style fat-pointer vectors (lengths travel in the
 * hidden header, see list_size()). */
typedef struct kd {
    kdnode *node_vec;
    int *tri_indices;
    Vector4 *vert_vec;
    Vector4 *norm_vec;
    cl_int3 *tri_vec;
} kd;

#ifdef __cplusplus
}
#define CLPT_STATIC_ASSERT(c, m) static_assert(c, m)
#else
#define CLPT_STATIC_ASSERT(c, m) _Static_assert(c, m)
#endif

CLPT_STATIC_ASSERT(sizeof(Vector3) == 16, "Vector3 is a 16-byte cl_float3");
CLPT_STATIC_ASSERT(sizeof(cl_int3) == 16, "cl_int3 is 16 bytes");
CLPT_STATIC_ASSERT(sizeof(Matrix) == 64, "Matrix is 4 rows of float4");
CLPT_STATIC_ASSERT(sizeof(kdnode) == 68, "kdnode is 68 bytes packed");
CLPT_STATIC_ASSERT(sizeof(Object) == 24, "Object is 24 bytes packed");
CLPT_STATIC_ASSERT(sizeof(Camera) == 48, "Camera is 48 bytes");
CLPT_STATIC_ASSERT(offsetof(kdnode, type) == 32, "kdnode.type at 32");
CLPT_STATIC_ASSERT(offsetof(kdnode, split.children) == 44, "children at 44");
CLPT_STATIC_ASSERT(offsetof(kdnode, leaf.ropes) == 44, "ropes at 44");

#endif /* CLPT_TYPES_H */
