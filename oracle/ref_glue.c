/* ref_glue.c -- TEST INFRASTRUCTURE.  Pointer-argument entry points into the
 * reference's own host code, linked into oracle/_ref/libref_host.so next to
 * the reference's unmodified sources.  Nothing here computes anything: each
 * function forwards to a reference function so that ctypes callers do not
 * have to pass 48/64-byte structs by value.
 */
#include <stddef.h>
#include "camera.h"   /* reference header, via -I$(REF)/include */
#include "kd_tree.h"
#include "list.h"

void
ref_cam_matrix_ptr(const Camera *cam, int height, Matrix *out) {
    *out = cam_matrix(*cam, height);
}

size_t
ref_sizeof_kdnode(void) {
    return sizeof(kdnode);
}

size_t
ref_sizeof_camera(void) {
    return sizeof(Camera);
}
