/* oracle_kernel.c -- CPU restatement of the reference render kernel.
 *
 * TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this.  It is the checker for
 * the CUDA path, never a fallback for it.
 *
 * PARITY STATUS.  The reference ships no tests, golden images or fixtures, and
 * its device code (src/kernel.cl) cannot be compiled in the build container (no
 * OpenCL headers, ICD or clang; gcc/g++ reject its OpenCL-C vector syntax).
 *   mode A (as shipped)  PINNED by a run of the reference's own kernel.cl on the
 *        B200 through the driver's OpenCL ICD (oracle/cl_harness.c; frames
 *        committed as tests/golden/ref_kernel_golden.npz and re-run live by
 *        tests/test_reference_kernel.py): same triangle at every pixel except exact
 *        edge-graze ties, same-triangle colours within 1 ulp (the vendor compiler
 *        contracts FMAs, this file does not).
 *   mode B (the code after the early return), jitter, multi-spp, mode C:
 *        "parity unpinned" -- the reference never executes them; this file is the
 *        definition.  Mode B's primary hits are mode A's.
 * What follows is a line-by-line restatement in C of
 *     src/kernel.cl:79-87    new_Ray
 *     src/kernel.cl:89-94    mul (matrix * point with perspective divide)
 *     src/kernel.cl:101-144  hit_AABB
 *     src/kernel.cl:146-174  traverse_AABB
 *     src/kernel.cl:227-255  hit_triangle (Moller-Trumbore, back-face culled)
 *     src/kernel.cl:296-422  trace_ray (rope traversal; tie rule :344, early
 *                            out :381, p1 update :385, mirror bounce :399-417)
 *     src/kernel.cl:424-473  render (ray generation :443-456)
 * walking the reference's own 68-byte node array (the host half -- tree,
 * tri_indices, camera matrix -- IS pinned byte-for-byte by the reference's
 * unmodified host code, see oracle/Makefile target `ref`).
 *
 * Arithmetic: fp32, one rounding per operation, in the source order of the
 * reference expressions; build with -ffp-contract=off.  Conventions for the
 * OpenCL built-ins, which the standard leaves to the implementation:
 *     dot(a,b)     = a.x*b.x + a.y*b.y + a.z*b.z      (left to right)
 *     cross(a,b)   = (a.y*b.z - a.z*b.y, a.z*b.x - a.x*b.z, a.x*b.y - a.y*b.x)
 *     normalize(v) = v / sqrt(dot(v,v))               (IEEE sqrt, 3 divides)
 * The literal 0.001 at kernel.cl:381 has no suffix; on an fp64-capable device
 * OpenCL C makes it a double, so that one comparison is evaluated in double.
 *
 * Modes
 *   0 "A" as shipped: colour of the first hit's shading normal, white on miss
 *         (the `return` at kernel.cl:396).
 *   1 "B" the code after that return enabled: deterministic mirror bounces,
 *         depth = bounces + 1 (the reference passes depth 2, kernel.cl:468).
 *   2 "C" EXTENSION with no reference behaviour: diffuse/mirror materials,
 *         cosine-weighted sampling; see oracle_path_* below.
 * spp > 1 or flags&ORACLE_JITTER adds sub-pixel jitter from Philox4x32-10;
 * with spp == 1 and no jitter, ray generation is exactly the reference's.
 */
#define _GNU_SOURCE
#include <sched.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { float x, y, z; } f3;
typedef struct { float s[4]; } f4;
typedef struct { int32_t s[4]; } i4;

#pragma pack(push, 1)
typedef struct onode { /* include/kd_tree.h:31-50 == src/kernel.cl:48-65 */
    f4 min, max;
    int32_t type; /* 0 split, 1 leaf */
    union {
        struct { float value; int32_t axis; int32_t children[2]; } split;
        struct { int32_t tris; int32_t tri_count; int32_t ropes[6]; } leaf;
    };
} onode;
#pragma pack(pop)

enum { ORACLE_JITTER = 1, ORACLE_ACCUMULATE = 2 };

typedef struct oracle_material { /* extension; 32 bytes */
    float albedo[3];
    int32_t kind; /* 0 diffuse, 1 mirror */
    float emission[3];
    float pad;
} oracle_material;

typedef struct oracle_params {
    /* scene, in the wire layout CLSetMeshes receives (src/CLState.c:124-202) */
    const onode *nodes;
    const int32_t *tri_indices;
    const i4 *tris;  /* 3 per triangle: {v, vn, vt, pad} */
    const f4 *verts;
    const f4 *norms;
    const float *cam; /* 16 floats, row-major inverse camera matrix */
    int32_t width, height;
    int32_t y0, y1;   /* rows [y0, y1) are rendered */
    int32_t mode, depth, spp, flags;
    uint32_t seed, sample_base;
    int32_t max_leaf_visits; /* cap on rope hops per ray (the reference has none) */
    int32_t threads;
    const oracle_material *materials; /* mode C */
    const int32_t *tri_material;      /* mode C: per-triangle material id or NULL */
    int32_t n_materials;
    int32_t reserved;
    /* outputs (any may be NULL); indexed y*width+x over the FULL image */
    float *rgba;      /* 4 floats per pixel */
    int32_t *prim_id; /* first hit of the primary ray of sample 0, -1 = miss */
    float *t_hit;
    float *uv;        /* 2 floats per pixel */
    float *normal;    /* 3 floats per pixel */
    float *path_edge; /* min over the hits of sample 0's path of min(u, v, |1-u-v|): how close the
                       * path came to a triangle edge (classifies tie-break mismatches); 2 = no hit */
    uint64_t *accum;  /* ORACLE_ACCUMULATE: 4 words per pixel, the running sums of r, g, b in 2^-32
                       * fixed point and the sample count; `rgba` then receives the mean so far */
    /* work counters, summed over all rays: rays, splits, leaves, tri tests,
     * vn-shaded hits, rays stopped by the cap */
    uint64_t counters[6];
} oracle_params;

/* ---------------------------------------------------------------- helpers */
static inline f3 v3(float x, float y, float z) { f3 r = { x, y, z }; return r; }
static inline f3 add3(f3 a, f3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline f3 sub3(f3 a, f3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline f3 scale3(f3 a, float k) { return v3(a.x * k, a.y * k, a.z * k); }
static inline float dot3(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline f3 cross3(f3 a, f3 b) {
    return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static inline f3 normalize3(f3 a) {
    float len = sqrtf(dot3(a, a));
    return v3(a.x / len, a.y / len, a.z / len);
}
static inline f3 xyz(f4 a) { return v3(a.s[0], a.s[1], a.s[2]); }
static inline float comp(f3 a, int k) { return k == 0 ? a.x : (k == 1 ? a.y : a.z); }

typedef struct { f3 orig, dir, invdir; int sign[3]; } ray_t;

static ray_t make_ray(f3 orig, f3 dir) { /* kernel.cl:79-87 */
    ray_t r;
    r.orig = orig;
    r.dir = dir;
    r.invdir = v3(1 / dir.x, 1 / dir.y, 1 / dir.z);
    r.sign[0] = r.invdir.x < 0;
    r.sign[1] = r.invdir.y < 0;
    r.sign[2] = r.invdir.z < 0;
    return r;
}

static f3 unproject(const float *M, f3 X) { /* kernel.cl:89-94 */
    float w = (M[12] * X.x + M[13] * X.y + M[14] * X.z) + M[15];
    float a = (M[0] * X.x + M[1] * X.y + M[2] * X.z) + M[3];
    float b = (M[4] * X.x + M[5] * X.y + M[6] * X.z) + M[7];
    float c = (M[8] * X.x + M[9] * X.y + M[10] * X.z) + M[11];
    return v3(a / w, b / w, c / w);
}

/* kernel.cl:101-144.  bounds[0]=min, bounds[1]=max. */
static int clip_root(const f3 bounds[2], const ray_t *r, float *tmin, float *tmax) {
    float tymin, tymax, tzmin, tzmax;
    *tmin = (bounds[r->sign[0]].x - r->orig.x) * r->invdir.x;
    *tmax = (bounds[1 - r->sign[0]].x - r->orig.x) * r->invdir.x;
    tymin = (bounds[r->sign[1]].y - r->orig.y) * r->invdir.y;
    tymax = (bounds[1 - r->sign[1]].y - r->orig.y) * r->invdir.y;
    if ((*tmin > tymax) || (tymin > *tmax)) return 0;
    if (tymin > *tmin) *tmin = tymin;
    if (tymax < *tmax) *tmax = tymax;
    tzmin = (bounds[r->sign[2]].z - r->orig.z) * r->invdir.z;
    tzmax = (bounds[1 - r->sign[2]].z - r->orig.z) * r->invdir.z;
    if ((*tmin > tzmax) || (tzmin > *tmax)) return 0;
    if (tzmin > *tmin) *tmin = tzmin;
    if (tzmax < *tmax) *tmax = tzmax;
    return *tmax > 0;
}

/* kernel.cl:146-174: slab interval of a leaf and the face the ray leaves by */
static void leaf_exit(const f3 bounds[2], const ray_t *r, float *tmin, float *tmax, int *far) {
    float tymin, tymax, tzmin, tzmax;
    *far = 1 - r->sign[0];
    *tmin = (bounds[r->sign[0]].x - r->orig.x) * r->invdir.x;
    *tmax = (bounds[1 - r->sign[0]].x - r->orig.x) * r->invdir.x;
    tymin = (bounds[r->sign[1]].y - r->orig.y) * r->invdir.y;
    tymax = (bounds[1 - r->sign[1]].y - r->orig.y) * r->invdir.y;
    if (tymin > *tmin) *tmin = tymin;
    if (tymax < *tmax) { *tmax = tymax; *far = 3 - r->sign[1]; }
    tzmin = (bounds[r->sign[2]].z - r->orig.z) * r->invdir.z;
    tzmax = (bounds[1 - r->sign[2]].z - r->orig.z) * r->invdir.z;
    if (tzmin > *tmin) *tmin = tzmin;
    if (tzmax < *tmax) { *tmax = tzmax; *far = 5 - r->sign[2]; }
}

/* kernel.cl:227-255, EPS = 0 */
static int hit_tri(f3 v0, f3 v1, f3 v2, f3 start, f3 dir, float *t, float *u, float *v) {
    f3 e1 = sub3(v1, v0), e2 = sub3(v2, v0);
    f3 pvec = cross3(dir, e2);
    float det = dot3(e1, pvec);
    if (det < 0.0) return 0;
    float inv = 1 / det;
    f3 tvec = sub3(start, v0);
    *u = dot3(tvec, pvec) * inv;
    if (*u < 0 || *u > 1) return 0;
    f3 qvec = cross3(tvec, e1);
    *v = dot3(dir, qvec) * inv;
    if (*v < 0 || *u + *v > 1) return 0;
    *t = dot3(e2, qvec) * inv;
    return *t > 0;
}

typedef struct hit_t {
    int did_hit, prim, capped;
    float t, u, v;
    f3 normal;
} hit_t;

/* The traversal loop of trace_ray, kernel.cl:311-389, for one ray. */
static hit_t closest_hit(const oracle_params *P, const ray_t *r, uint64_t *cnt) {
    hit_t h;
    memset(&h, 0, sizeof(h));
    h.prim = -1;
    const onode *kd = P->nodes;
    f3 rb[2] = { xyz(kd[0].min), xyz(kd[0].max) };
    float tmin, tmax;
    cnt[0]++;
    if (!clip_root(rb, r, &tmin, &tmax)) return h;
    f3 p1 = r->orig;
    if (tmin > 0) p1 = add3(p1, scale3(r->dir, tmin)); /* p1 += tmin * dir */
    int index = 0, visits = 0;
    float minHit = 0;
    while (index != -1) {
        while (kd[index].type == 0) {
            int axis = kd[index].split.axis;
            int cond = comp(p1, axis) > kd[index].split.value;
            index = kd[index].split.children[cond];
            cnt[1]++;
        }
        cnt[2]++;
        const onode *leaf = &kd[index];
        if (leaf->leaf.tris != -1) {
            for (int i = 0; i < leaf->leaf.tri_count; i++) {
                int b = P->tri_indices[leaf->leaf.tris + i];
                i4 c1 = P->tris[3 * b + 0], c2 = P->tris[3 * b + 1], c3 = P->tris[3 * b + 2];
                f3 v1 = xyz(P->verts[c1.s[0]]), v2 = xyz(P->verts[c2.s[0]]),
                   v3_ = xyz(P->verts[c3.s[0]]);
                float t = 0, u = 0, v = 0;
                cnt[3]++;
                if (hit_tri(v1, v2, v3_, r->orig, r->dir, &t, &u, &v)) {
                    if (!h.did_hit || t <= minHit) { /* later triangle wins ties */
                        h.did_hit = 1;
                        minHit = t;
                        h.prim = b;
                        h.u = u;
                        h.v = v;
                        if (c1.s[1] >= 0) {
                            f3 n1 = xyz(P->norms[c1.s[1]]), n2 = xyz(P->norms[c2.s[1]]),
                               n3 = xyz(P->norms[c3.s[1]]);
                            float w = 1.0f - u - v;
                            h.normal = normalize3(add3(add3(scale3(n1, w), scale3(n2, u)),
                                                       scale3(n3, v)));
                        } else {
                            h.normal = normalize3(cross3(sub3(v2, v1), sub3(v3_, v1)));
                        }
                    }
                }
            }
        }
        f3 lb[2] = { xyz(leaf->min), xyz(leaf->max) };
        int far;
        leaf_exit(lb, r, &tmin, &tmax, &far);
        if (h.did_hit && (double)tmin + 0.001 > (double)minHit) break;
        index = leaf->leaf.ropes[far];
        p1 = add3(r->orig, scale3(r->dir, tmax)); /* orig + tmax * dir */
        if (index == -1) break;
        if (++visits >= P->max_leaf_visits) { h.capped = 1; cnt[5]++; break; }
    }
    h.t = minHit;
    if (h.did_hit && P->norms != NULL && P->tris[3 * h.prim].s[1] >= 0) cnt[4]++;
    return h;
}

/* ------------------------------------------------------------ Philox4x32-10
 * Counter-based RNG (Salmon et al., SC'11).  Integer-only, so CPU and GPU
 * streams are identical.  counter = (pixel, sample, dimension block, 0),
 * key = (seed, 'clpt'). */
static void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int round = 0; round < 10; round++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

void oracle_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    memcpy(out, ctr, 16);
    philox4x32_10(out, key[0], key[1]);
}

/* EXTENSION (progressive accumulation has no reference behaviour; this is the definition).
 * A sample's colour enters the running sum as 2^-32 fixed point, rounded to nearest even;
 * negative and NaN count as 0, values saturate at 2^20.  Integer sums do not depend on the order
 * of the additions, so one device adding samples 0,1,2,... and N devices adding every N-th
 * sample each produce the same bits. */
static inline uint64_t fix32(float x) {
    if (!(x > 0.0f)) return 0;
    if (x >= 1048576.0f) return (uint64_t)1 << 52;
    return (uint64_t)llrintf(x * 4294967296.0f);
}

static inline float u01(uint32_t x) { return (float)(x >> 8) * 0x1p-24f; } /* [0,1), exact */

#define ORACLE_KEY1 0x636c7074u

/* ------------------------------------------------------------------ mode C
 * EXTENSION (no reference behaviour; this restatement IS the definition).
 * Path with `depth` segments.  At a hit: add throughput*emission; mirror
 * materials reflect like kernel.cl:400-401; diffuse materials multiply the
 * throughput by the albedo and continue in a cosine-weighted direction about
 * the shading normal.  A miss adds throughput * white (the reference's miss
 * colour, kernel.cl:421).  Sampling uses only + - * / sqrt so that it is
 * bit-reproducible: a point in the unit disk by rejection from Philox pairs
 * (4 tries, then the centre), lifted to the hemisphere (Malley), in the
 * branch-free orthonormal basis of Duff et al. 2017. */
static f3 cosine_dir(f3 n, uint32_t pixel, uint32_t sample, uint32_t bounce, uint32_t seed) {
    float a = 0, b = 0;
    uint32_t c[4] = { pixel, sample, 1u + bounce, 0 };
    philox4x32_10(c, seed, ORACLE_KEY1);
    for (int k = 0; k < 2; k++) {
        float x = 2 * u01(c[2 * k]) - 1, y = 2 * u01(c[2 * k + 1]) - 1;
        if (x * x + y * y <= 1) { a = x; b = y; goto found; }
    }
    c[0] = pixel; c[1] = sample; c[2] = 1u + bounce; c[3] = 1;
    philox4x32_10(c, seed, ORACLE_KEY1);
    for (int k = 0; k < 2; k++) {
        float x = 2 * u01(c[2 * k]) - 1, y = 2 * u01(c[2 * k + 1]) - 1;
        if (x * x + y * y <= 1) { a = x; b = y; goto found; }
    }
found:;
    float zz = 1 - a * a - b * b;
    float z = sqrtf(zz > 0 ? zz : 0);
    float sg = n.z >= 0 ? 1.0f : -1.0f;
    float p = -1 / (sg + n.z);
    float q = n.x * n.y * p;
    f3 t1 = v3(1 + sg * n.x * n.x * p, sg * q, -sg * n.x);
    f3 t2 = v3(q, sg + n.y * n.y * p, -n.y);
    f3 d = add3(add3(scale3(t1, a), scale3(t2, b)), scale3(n, z));
    return normalize3(d);
}

static inline void note_edge(float *edge, const hit_t *h) {
    if (!edge || !h->did_hit) return;
    float w = fabsf(1.0f - h->u - h->v), m = h->u < h->v ? h->u : h->v;
    if (w < m) m = w;
    if (m < *edge) *edge = m;
}

static f3 shade_sample(const oracle_params *P, ray_t r, uint32_t pixel, uint32_t sample,
                       hit_t *first, float *edge, uint64_t *cnt) {
    int depth = P->depth;
    if (P->mode == 0) depth = depth > 0 ? 1 : 0;
    if (P->mode == 2) {
        f3 L = v3(0, 0, 0), T = v3(1, 1, 1);
        for (int seg = 0; seg < depth; seg++) {
            hit_t h = closest_hit(P, &r, cnt);
            if (seg == 0 && first) *first = h;
            note_edge(edge, &h);
            if (!h.did_hit) { L = add3(L, T); return L; }
            int m = P->tri_material ? P->tri_material[h.prim] : 0;
            if (m < 0 || m >= P->n_materials) m = 0;
            oracle_material mat;
            if (P->n_materials > 0) mat = P->materials[m];
            else { mat.albedo[0] = mat.albedo[1] = mat.albedo[2] = 0.5f; mat.kind = 0;
                   mat.emission[0] = mat.emission[1] = mat.emission[2] = 0; }
            L = add3(L, v3(T.x * mat.emission[0], T.y * mat.emission[1], T.z * mat.emission[2]));
            T = v3(T.x * mat.albedo[0], T.y * mat.albedo[1], T.z * mat.albedo[2]);
            f3 hitp = add3(r.orig, scale3(r.dir, h.t));
            f3 nd;
            if (mat.kind == 1) {
                nd = normalize3(sub3(r.dir, scale3(h.normal, 2 * dot3(r.dir, h.normal))));
            } else {
                nd = cosine_dir(h.normal, pixel, sample, (uint32_t)seg, P->seed);
            }
            hitp = add3(hitp, scale3(nd, 0.0001f));
            r = make_ray(hitp, nd);
        }
        return L;
    }
    /* modes A and B: kernel.cl:296-422 with the tail recursion unrolled */
    f3 col = v3(0, 0, 0);
    float str = 1.0f;
    for (; depth > 0; depth--) {
        hit_t h = closest_hit(P, &r, cnt);
        if (first) { *first = h; first = NULL; }
        note_edge(edge, &h);
        if (!h.did_hit) break;
        f3 nc = v3((h.normal.x + 1) / 2, (h.normal.y + 1) / 2, (h.normal.z + 1) / 2);
        if (P->mode == 0) return nc; /* kernel.cl:396 */
        f3 no = add3(r.orig, scale3(r.dir, h.t));
        f3 nd = normalize3(sub3(r.dir, scale3(h.normal, 2 * dot3(r.dir, h.normal))));
        no = add3(no, scale3(nd, 0.0001f));
        col = add3(scale3(col, 1 - str), scale3(nc, str));
        str *= 0.2f;
        r = make_ray(no, nd);
    }
    /* kernel.cl:421 */
    return v3((1 - str) * col.x + str, (1 - str) * col.y + str, (1 - str) * col.z + str);
}

int oracle_render(oracle_params *P) {
    const int W = P->width, H = P->height;
    if ((P->flags & ORACLE_ACCUMULATE) && (!P->accum || !P->rgba)) return 1;
    const float *M = P->cam;
    int spp = P->spp < 1 ? 1 : P->spp;
    int jitter = (P->flags & ORACLE_JITTER) != 0;
    uint64_t total[6] = { 0, 0, 0, 0, 0, 0 };
#ifdef _OPENMP
    if (P->threads > 0) omp_set_num_threads(P->threads);
#endif
#pragma omp parallel
    {
        uint64_t cnt[6] = { 0, 0, 0, 0, 0, 0 };
#pragma omp for schedule(dynamic, 1) nowait
        for (int y = P->y0; y < P->y1; y++) {
            for (int x = 0; x < W; x++) {
                uint32_t pixel = (uint32_t)(y * W + x);
                /* kernel.cl:443-445: eye = column 2 of the inverse / w */
                f3 origin = v3(M[2] / M[14], M[6] / M[14], M[10] / M[14]);
                f3 acc = v3(0, 0, 0);
                uint64_t fsum[3] = { 0, 0, 0 };
                hit_t first;
                float edge = 2.0f;
                memset(&first, 0, sizeof(first));
                first.prim = -1;
                for (int s = 0; s < spp; s++) {
                    uint32_t sample = P->sample_base + (uint32_t)s;
                    float fx = (float)(uint32_t)x - (float)(uint32_t)W / 2;
                    float fy = (float)(uint32_t)y - (float)(uint32_t)H / 2;
                    if (jitter) {
                        uint32_t c[4] = { pixel, sample, 0, 0 };
                        philox4x32_10(c, P->seed, ORACLE_KEY1);
                        fx = fx + (u01(c[0]) - 0.5f);
                        fy = fy + (u01(c[1]) - 0.5f);
                    }
                    f3 ncp = unproject(M, v3(fx, fy, -1));
                    f3 fcp = unproject(M, v3(fx, fy, 1));
                    f3 dir = normalize3(sub3(fcp, ncp));
                    ray_t r = make_ray(origin, dir);
                    f3 c = shade_sample(P, r, pixel, sample, s == 0 ? &first : NULL, s == 0 ? &edge : NULL, cnt);
                    acc = add3(acc, c);
                    fsum[0] += fix32(c.x);
                    fsum[1] += fix32(c.y);
                    fsum[2] += fix32(c.z);
                }
                size_t o = (size_t)y * (size_t)W + (size_t)x;
                if (P->rgba) {
                    float *px = P->rgba + 4 * o;
                    if (P->flags & ORACLE_ACCUMULATE) {
                        uint64_t *a = P->accum + 4 * o;
                        a[0] += fsum[0]; a[1] += fsum[1]; a[2] += fsum[2]; a[3] += (uint64_t)spp;
                        const double k = (double)a[3] * 4294967296.0;
                        px[0] = (float)((double)a[0] / k); px[1] = (float)((double)a[1] / k);
                        px[2] = (float)((double)a[2] / k); px[3] = 1.0f;
                    } else if (spp == 1) {
                        px[0] = acc.x; px[1] = acc.y; px[2] = acc.z; px[3] = 1.0f;
                    } else {
                        float inv = 1.0f / (float)spp;
                        px[0] = acc.x * inv; px[1] = acc.y * inv; px[2] = acc.z * inv; px[3] = 1.0f;
                    }
                }
                if (P->prim_id) P->prim_id[o] = first.did_hit ? first.prim : -1;
                if (P->t_hit) P->t_hit[o] = first.did_hit ? first.t : 0.0f;
                if (P->uv) { P->uv[2 * o] = first.u; P->uv[2 * o + 1] = first.v; }
                if (P->path_edge) P->path_edge[o] = edge;
                if (P->normal) {
                    P->normal[3 * o] = first.normal.x;
                    P->normal[3 * o + 1] = first.normal.y;
                    P->normal[3 * o + 2] = first.normal.z;
                }
            }
        }
#pragma omp critical
        for (int k = 0; k < 6; k++) total[k] += cnt[k];
    }
    for (int k = 0; k < 6; k++) P->counters[k] = total[k];
    return 0;
}

size_t oracle_sizeof_params(void) { return sizeof(oracle_params); }
int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
/* Cores this process may run on (the affinity mask), whatever OMP_NUM_THREADS says:
 * torchrun exports OMP_NUM_THREADS=1, which must not shrink the CPU baseline. */
int oracle_host_cores(void) {
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof(set), &set) == 0 && CPU_COUNT(&set) > 0) return CPU_COUNT(&set);
    return oracle_num_threads();
}
