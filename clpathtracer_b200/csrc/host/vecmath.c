/* vecmath.c -- vec3, 4x4 matrix and the camera matrix (clpt_host.h).
 *
 * Numerics contract: fp32, one rounding per source operation, evaluated in the
 * reference's order so that results are bit-identical to the reference host
 * code built without FMA contraction (src/vector.c, src/matrix.c,
 * src/camera.c).  Build this file with -ffp-contract=off.
 */
#include <math.h>

#include "clpt_host.h"

/* ---- vec3 (src/vector.c:5-113) ---- */

vec_t
vec_dot(Vector3 a, Vector3 b) {
    return a.s[0] * b.s[0] + a.s[1] * b.s[1] + a.s[2] * b.s[2];
}

vec_t
vec_length_squared(Vector3 v) {
    return vec_dot(v, v);
}

vec_t
vec_length(Vector3 v) {
    /* reference takes the root in double, then narrows (vector.c:16-18) */
    return (vec_t)sqrt((double)vec_length_squared(v));
}

Vector3 *
vec_normalize(Vector3 *v) {
    vec_t len = vec_length(*v);
    for (int k = 0; k < 3; k++) {
        v->s[k] /= len;
    }
    return v;
}

Vector3
vec_normalized(Vector3 v) {
    vec_normalize(&v);
    return v;
}

Vector3
vec_add(Vector3 a, Vector3 b) {
    return Vector3(a.s[0] + b.s[0], a.s[1] + b.s[1], a.s[2] + b.s[2]);
}

Vector3
vec_subtract(Vector3 a, Vector3 b) {
    return Vector3(a.s[0] - b.s[0], a.s[1] - b.s[1], a.s[2] - b.s[2]);
}

Vector3
vec_cross(Vector3 a, Vector3 b) {
    return Vector3(a.s[1] * b.s[2] - a.s[2] * b.s[1],
                   a.s[2] * b.s[0] - a.s[0] * b.s[2],
                   a.s[0] * b.s[1] - a.s[1] * b.s[0]);
}

Vector3 *
vec_negate(Vector3 *v) {
    for (int k = 0; k < 3; k++) {
        v->s[k] = -v->s[k];
    }
    return v;
}

Vector3
vec_negated(Vector3 v) {
    vec_negate(&v);
    return v;
}

Vector3 *
vec_scale(Vector3 *v, vec_t f) {
    for (int k = 0; k < 3; k++) {
        v->s[k] *= f;
    }
    return v;
}

Vector3
vec_scaled(Vector3 v, vec_t f) {
    vec_scale(&v, f);
    return v;
}

Vector3
vec_divide(Vector3 a, Vector3 b) {
    return Vector3(a.s[0] / b.s[0], a.s[1] / b.s[1], a.s[2] / b.s[2]);
}

Vector3
vec_min(Vector3 a, Vector3 b) {
    Vector3 r = a;
    for (int k = 0; k < 3; k++) {
        r.s[k] = a.s[k] < b.s[k] ? a.s[k] : b.s[k];
    }
    r.s[3] = 0;
    return r;
}

Vector3
vec_max(Vector3 a, Vector3 b) {
    Vector3 r = a;
    for (int k = 0; k < 3; k++) {
        r.s[k] = a.s[k] > b.s[k] ? a.s[k] : b.s[k];
    }
    r.s[3] = 0;
    return r;
}

/* ---- 4x4 (src/matrix.c) ---- */

void
mat_set(Matrix *mat, unsigned int n, unsigned int m, vec_t value) {
    mat->rows[m].s[n] = value; /* (column, row) like matrix.c:5-8 */
}

vec_t
mat_get(Matrix mat, unsigned int n, unsigned int m) {
    return mat.rows[m].s[n];
}

Matrix
mat_add(Matrix a, Matrix b) {
    Matrix r;
    for (int i = 0; i < 16; i++) {
        r.rows[i >> 2].s[i & 3] = a.rows[i >> 2].s[i & 3] + b.rows[i >> 2].s[i & 3];
    }
    return r;
}

Matrix
mat_multiply(Matrix a, Matrix b) {
    Matrix r;
    for (int i = 0; i < 4; i++) {
        for (int j = 0; j < 4; j++) {
            /* accumulate from 0 over k ascending (matrix.c:29-35) */
            vec_t acc = 0;
            for (int k = 0; k < 4; k++) {
                acc += a.rows[i].s[k] * b.rows[k].s[j];
            }
            r.rows[i].s[j] = acc;
        }
    }
    return r;
}

Matrix *
mat_scale(Matrix *mat, vec_t f) {
    for (int i = 0; i < 16; i++) {
        mat->rows[i >> 2].s[i & 3] *= f;
    }
    return mat;
}

Matrix
mat_scaled(Matrix mat, vec_t f) {
    mat_scale(&mat, f);
    return mat;
}

/* Cofactor inverse, table-driven.  Every cofactor is the same six-term fp32
 * expression, term for term and in the same association order, as the
 * reference's spelled-out version (src/matrix.c:54-153): a left-to-right sum
 * of triple products with the sign pattern + - - + + - (or its negation, where
 * the leading minus binds to the first factor).  E(i) is flat element i. */
#define E(i) (m.rows[(i) >> 2].s[(i) & 3])
Matrix
mat_inverse(Matrix m, int *err) {
    /* Table of the 16 cofactors.  Row r of the table lists, for inverse
     * element r, the six index triples and whether the expression starts
     * with a negated factor (the reference alternates the sign pattern
     * + - - + + -  and  - + + - - +). */
    static const unsigned char T[16][18] = {
        { 5, 10, 15, 5, 11, 14, 9, 6, 15, 9, 7, 14, 13, 6, 11, 13, 7, 10 },
        { 1, 10, 15, 1, 11, 14, 9, 2, 15, 9, 3, 14, 13, 2, 11, 13, 3, 10 },
        { 1, 6, 15, 1, 7, 14, 5, 2, 15, 5, 3, 14, 13, 2, 7, 13, 3, 6 },
        { 1, 6, 11, 1, 7, 10, 5, 2, 11, 5, 3, 10, 9, 2, 7, 9, 3, 6 },
        { 4, 10, 15, 4, 11, 14, 8, 6, 15, 8, 7, 14, 12, 6, 11, 12, 7, 10 },
        { 0, 10, 15, 0, 11, 14, 8, 2, 15, 8, 3, 14, 12, 2, 11, 12, 3, 10 },
        { 0, 6, 15, 0, 7, 14, 4, 2, 15, 4, 3, 14, 12, 2, 7, 12, 3, 6 },
        { 0, 6, 11, 0, 7, 10, 4, 2, 11, 4, 3, 10, 8, 2, 7, 8, 3, 6 },
        { 4, 9, 15, 4, 11, 13, 8, 5, 15, 8, 7, 13, 12, 5, 11, 12, 7, 9 },
        { 0, 9, 15, 0, 11, 13, 8, 1, 15, 8, 3, 13, 12, 1, 11, 12, 3, 9 },
        { 0, 5, 15, 0, 7, 13, 4, 1, 15, 4, 3, 13, 12, 1, 7, 12, 3, 5 },
        { 0, 5, 11, 0, 7, 9, 4, 1, 11, 4, 3, 9, 8, 1, 7, 8, 3, 5 },
        { 4, 9, 14, 4, 10, 13, 8, 5, 14, 8, 6, 13, 12, 5, 10, 12, 6, 9 },
        { 0, 9, 14, 0, 10, 13, 8, 1, 14, 8, 2, 13, 12, 1, 10, 12, 2, 9 },
        { 0, 5, 14, 0, 6, 13, 4, 1, 14, 4, 2, 13, 12, 1, 6, 12, 2, 5 },
        { 0, 5, 10, 0, 6, 9, 4, 1, 10, 4, 2, 9, 8, 1, 6, 8, 2, 5 },
    };
    /* inverse element e (row-major) starts with '+' iff (row+col) is even */
    Matrix inv;
    for (int e = 0; e < 16; e++) {
        const unsigned char *t = T[e];
        int plus = (((e >> 2) + (e & 3)) & 1) == 0;
        vec_t p0 = E(t[0]), p1 = E(t[3]), p2 = E(t[6]);
        vec_t p3 = E(t[9]), p4 = E(t[12]), p5 = E(t[15]);
        vec_t acc;
        if (plus) {
            /*  a*b*c - a*d*e - f*g*c + f*h*e + i*g*d - i*h*b */
            acc = p0 * E(t[1]) * E(t[2]);
            acc = acc - p1 * E(t[4]) * E(t[5]);
            acc = acc - p2 * E(t[7]) * E(t[8]);
            acc = acc + p3 * E(t[10]) * E(t[11]);
            acc = acc + p4 * E(t[13]) * E(t[14]);
            acc = acc - p5 * E(t[16]) * E(t[17]);
        } else {
            /* -a*b*c + a*d*e + f*g*c - f*h*e - i*g*d + i*h*b ; the leading
             * minus binds to the first factor (exact, sign flip only) */
            acc = -p0 * E(t[1]) * E(t[2]);
            acc = acc + p1 * E(t[4]) * E(t[5]);
            acc = acc + p2 * E(t[7]) * E(t[8]);
            acc = acc - p3 * E(t[10]) * E(t[11]);
            acc = acc - p4 * E(t[13]) * E(t[14]);
            acc = acc + p5 * E(t[16]) * E(t[17]);
        }
        inv.rows[e >> 2].s[e & 3] = acc;
    }
    vec_t det = E(0) * inv.rows[0].s[0] + E(1) * inv.rows[1].s[0] +
                E(2) * inv.rows[2].s[0] + E(3) * inv.rows[3].s[0];
    if (det == 0) {
        if (err) {
            *err = 1;
        }
        Matrix zero = { 0 };
        return zero;
    }
    mat_scale(&inv, 1 / det);
    return inv;
}
#undef E

/* ---- camera (src/camera.c:5-70) ---- */

static Matrix
view_of(const Camera *cam) {
    /* left-handed: x right, y up, z forward.  left = (fz, 0, -fx)/|.|,
     * up = forward x left, translation = basis . (-position)  (camera.c:6-33) */
    Vector3 fwd = cam->Forward;
    Vector3 left = Vector3(fwd.s[2], 0, -fwd.s[0]);
    vec_normalize(&left);
    Vector3 up = vec_cross(fwd, left);
    Vector3 npos = vec_negated(cam->Position);
    Matrix v = { 0 };
    const Vector3 *basis[3] = { &left, &up, &fwd };
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) {
            v.rows[r].s[c] = basis[r]->s[c];
        }
        v.rows[r].s[3] = vec_dot(*basis[r], npos);
    }
    v.rows[3].s[3] = 1;
    return v;
}

static Matrix
projection_of(const Camera *cam) {
    /* camera.c:35-50; cot(fov/2) evaluated in double then narrowed */
    Matrix p = { 0 };
    vec_t n = cam->Near, f = cam->Far;
    vec_t c = 1 / (vec_t)tan((double)cam->FOV / 2);
    p.rows[0].s[0] = c;
    p.rows[1].s[1] = c;
    p.rows[2].s[2] = -(f + n) / (n - f);
    p.rows[2].s[3] = (2 * f * n) / (n - f);
    p.rows[3].s[2] = 1;
    return p;
}

static Matrix
device_of(int height) {
    /* camera.c:52-60: pixel units are half-heights on both axes */
    Matrix d = { 0 };
    d.rows[0].s[0] = (vec_t)height / 2;
    d.rows[1].s[1] = (vec_t)height / 2;
    d.rows[2].s[2] = 1;
    d.rows[3].s[3] = 1;
    return d;
}

void
cam_matrix_ptr(const Camera *cam, int height, Matrix *out) {
    Matrix m = mat_multiply(mat_multiply(device_of(height), projection_of(cam)),
                            view_of(cam));
    *out = mat_inverse(m, NULL);
}

Matrix
cam_matrix(Camera cam, int height) {
    Matrix out;
    cam_matrix_ptr(&cam, height, &out);
    return out;
}
