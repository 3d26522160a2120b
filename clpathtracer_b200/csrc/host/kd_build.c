/* kd_build.c -- kd-tree builder with ropes (clpt_host.h "kd-tree").
 *
 * Produces the reference's wire format: a PREORDER array of 68-byte nodes
 * (left child is always parent+1), a tri_indices array filled leaf by leaf in
 * preorder, and six neighbour links ("ropes") per leaf.  At depth 15 / 25 bins
 * the output is byte-identical to the reference builder
 * (src/kd_tree.c:95-276); tests/test_host_parity.py pins that against the
 * reference's own unmodified code.
 *
 * The split rule is the reference's, restated (src/kd_tree.c:113-157):
 *   for axis in x,y,z (skipped when the extent < 1e-9, compared in double)
 *     for i in 0..nbins-1:  d = (i+1)/(nbins+1),  v = min[axis] + d*extent
 *        SL = half-box area on the low side  + sum of areas of triangles with
 *             a vertex <= v;  SR likewise on the high side (vertex >= v)
 *        cost = NL*SL + NR*SR          (fp32, triangles summed in list order)
 *   keep the FIRST candidate that attains the minimum cost.
 *   Stop (leaf) at <= 1 triangle, depth 0, no candidate, or a plane on the
 *   cell boundary.  Triangles with a vertex within 1e-9 (double) of the plane
 *   go to both children.
 *
 * What is different from the reference is how that rule is evaluated: here
 * triangles live in per-node structure-of-arrays (per-axis min/max + area,
 * 32 B per reference instead of an 80 B struct copy), the candidate planes of
 * large nodes are evaluated concurrently (each candidate still sums its
 * triangles in list order, so every fp32 cost is bit-identical to a serial
 * evaluation), subtrees are built as OpenMP tasks into a linked tree whose
 * nodes know their subtree sizes, so the preorder arrays are then written -- and
 * the ropes linked -- by parallel tasks at known offsets.  Depth and bin count
 * are run-time parameters.
 */
#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "clpt_host.h"

#define KD_EPS 0.000000001 /* double, like src/kd_tree.c:10 */

/* Triangle references of one cell, structure-of-arrays. */
typedef struct tri_set {
    int n;
    int *id;      /* original triangle index */
    float *lo[3]; /* per-axis min over the three vertices */
    float *hi[3]; /* per-axis max over the three vertices */
    float *area;  /* |cross|/2 */
    void *block;
} tri_set;

typedef struct bnode {
    float bmin[3], bmax[3];
    int leaf;
    float value;
    int axis;
    struct bnode *kid[2];
    int *ids; /* leaf: triangle ids in list order */
    int nids;
    int sub_nodes, sub_refs; /* size of this subtree: where its preorder block goes */
} bnode;

static void *
xmalloc(size_t n) {
    void *p = malloc(n ? n : 1);
    if (p == NULL) {
        perror("malloc");
        exit(EXIT_FAILURE);
    }
    return p;
}

/* Build-time nodes and leaf id lists come from per-thread bump chunks that are
 * released together when the build ends: one malloc per 256 KiB instead of two
 * or three per node. */
#define ARENA_CHUNK (256u << 10)
#define ARENA_THREADS 256
typedef struct arena_chunk {
    struct arena_chunk *next;
} arena_chunk;
static arena_chunk *g_arena_chunks = NULL;
static struct {
    char *cur, *end;
    char pad[48]; /* one cache line per thread */
} g_arena[ARENA_THREADS];

static int
arena_thread(void) {
#ifdef _OPENMP
    extern int omp_get_thread_num(void);
    return omp_get_thread_num() % ARENA_THREADS;
#else
    return 0;
#endif
}

/* Team size of the build's parallel regions: every thread owns one arena slot, so a
 * host with more hardware threads than slots must not start more (threads t and
 * t + ARENA_THREADS would otherwise share one unsynchronised bump pointer). */
static int
arena_team(void) {
#ifdef _OPENMP
    extern int omp_get_max_threads(void);
    int n = omp_get_max_threads();
    return n < 1 ? 1 : (n > ARENA_THREADS ? ARENA_THREADS : n);
#else
    return 1;
#endif
}

static void *
arena_alloc(size_t bytes) {
    bytes = (bytes + 15u) & ~(size_t)15u;
    int t = arena_thread();
    if ((size_t)(g_arena[t].end - g_arena[t].cur) < bytes) {
        size_t payload = bytes > ARENA_CHUNK ? bytes : ARENA_CHUNK;
        arena_chunk *c = xmalloc(sizeof(arena_chunk) + 16 + payload);
#pragma omp critical(kd_arena)
        {
            c->next = g_arena_chunks;
            g_arena_chunks = c;
        }
        g_arena[t].cur = (char *)c + 16 + sizeof(arena_chunk) - (sizeof(arena_chunk) % 16);
        g_arena[t].end = (char *)c + sizeof(arena_chunk) + 16 + payload;
    }
    void *p = g_arena[t].cur;
    g_arena[t].cur += bytes;
    return p;
}

static void
arena_release(void) {
    while (g_arena_chunks) {
        arena_chunk *c = g_arena_chunks;
        g_arena_chunks = c->next;
        free(c);
    }
    memset(g_arena, 0, sizeof(g_arena));
}

static tri_set
tri_set_alloc(int n) {
    tri_set s;
    size_t cap = (size_t)(n > 0 ? n : 1);
    s.block = xmalloc(cap * 8 * sizeof(float));
    float *f = s.block;
    s.n = 0;
    s.id = (int *)f;
    for (int a = 0; a < 3; a++) {
        s.lo[a] = f + cap * (1 + a);
        s.hi[a] = f + cap * (4 + a);
    }
    s.area = f + cap * 7;
    return s;
}

static void
tri_set_free(tri_set *s) {
    free(s->block);
    s->block = NULL;
}

static inline void
tri_set_push(tri_set *dst, const tri_set *src, int t) {
    int k = dst->n++;
    dst->id[k] = src->id[t];
    for (int a = 0; a < 3; a++) {
        dst->lo[a][k] = src->lo[a][t];
        dst->hi[a][k] = src->hi[a][t];
    }
    dst->area[k] = src->area[t];
}

static bnode *
make_leaf(const float *bmin, const float *bmax, const tri_set *s) {
    bnode *b = arena_alloc(sizeof(*b));
    memcpy(b->bmin, bmin, sizeof(b->bmin));
    memcpy(b->bmax, bmax, sizeof(b->bmax));
    b->leaf = 1;
    b->kid[0] = b->kid[1] = NULL;
    b->nids = s->n;
    b->ids = arena_alloc(sizeof(int) * (size_t)s->n);
    memcpy(b->ids, s->id, sizeof(int) * (size_t)s->n);
    b->sub_nodes = 1;
    b->sub_refs = s->n;
    return b;
}

/* Cost of one candidate plane.  Sums run over the triangles in list order in
 * fp32 (src/kd_tree.c:126-145).  `bound` is a cost already achieved by another
 * candidate: once the partial cost exceeds it the candidate cannot win (all
 * terms are non-negative, so the cost only grows) and +inf is returned. */
static float
plane_cost(const tri_set *s, int axis, float v, float SL, float SR,
           const volatile float *bound) {
    const float *lo = s->lo[axis], *hi = s->hi[axis], *area = s->area;
    int NL = 0, NR = 0;
    int n = s->n;
    for (int base = 0; base < n; base += 512) {
        int end = base + 512 < n ? base + 512 : n;
        for (int t = base; t < end; t++) {
            int l = lo[t] <= v, r = hi[t] >= v;
            NL += l;
            NR += r;
            SL += l ? area[t] : 0.0f; /* x + 0 == x: same bits as a skipped add */
            SR += r ? area[t] : 0.0f;
        }
        if (bound && NL * SL + NR * SR > *bound) {
            return INFINITY;
        }
    }
    return NL * SL + NR * SR;
}

typedef struct candidate {
    int axis;
    float v, SL0, SR0, cost;
} candidate;

static bnode *
build_cell(tri_set s, const float *bmin, const float *bmax, int depth,
           int nbins) {
    bnode *out;
    if (s.n <= 1 || depth == 0) {
        out = make_leaf(bmin, bmax, &s);
        tri_set_free(&s);
        return out;
    }
    float ext[3];
    for (int a = 0; a < 3; a++) {
        ext[a] = bmax[a] - bmin[a];
    }
    candidate *cand = xmalloc(sizeof(candidate) * 3 * (size_t)nbins);
    int ncand = 0;
    for (int axis = 0; axis < 3; axis++) {
        float e = ext[axis];
        if (e < KD_EPS) {
            continue;
        }
        float p = ext[(axis + 1) % 3], q = ext[(axis + 2) % 3];
        for (int i = 0; i < nbins; i++) {
            float d = (float)(i + 1) / (float)(nbins + 1);
            candidate c;
            c.axis = axis;
            c.v = bmin[axis] + d * e;
            c.SL0 = 2 * (p * q + e * d * (p + q));
            c.SR0 = 2 * (p * q + e * (1 - d) * (p + q));
            c.cost = INFINITY;
            cand[ncand++] = c;
        }
    }
    volatile float bound = INFINITY;
    if (s.n >= 16384 && ncand > 1) {
        /* big cell: candidates in parallel, shared pruning bound */
#pragma omp taskloop grainsize(1) shared(cand, bound, s) default(none) firstprivate(ncand)
        for (int k = 0; k < ncand; k++) {
            float c = plane_cost(&s, cand[k].axis, cand[k].v, cand[k].SL0,
                                 cand[k].SR0, &bound);
            cand[k].cost = c;
            if (c < bound) {
#pragma omp critical(kd_bound)
                if (c < bound) {
                    bound = c;
                }
            }
        }
    } else {
        for (int k = 0; k < ncand; k++) {
            float c = plane_cost(&s, cand[k].axis, cand[k].v, cand[k].SL0,
                                 cand[k].SR0, &bound);
            cand[k].cost = c;
            if (c < bound) {
                bound = c;
            }
        }
    }
    /* first candidate attaining the minimum (strict '<' in the reference) */
    int best = -1;
    for (int k = 0; k < ncand; k++) {
        if (best < 0 || cand[k].cost < cand[best].cost) {
            /* a pruned candidate (inf) can only be taken as the very first */
            best = k;
        }
    }
    int ok = best >= 0;
    float best_v = 0;
    int best_axis = 0;
    if (ok) {
        best_v = cand[best].v;
        best_axis = cand[best].axis;
        if (best_v <= bmin[best_axis] || bmax[best_axis] <= best_v) {
            ok = 0;
        }
    }
    free(cand);
    if (!ok) {
        out = make_leaf(bmin, bmax, &s);
        tri_set_free(&s);
        return out;
    }
    /* partition: a vertex within EPS of the plane puts the triangle on both
     * sides; comparisons in double (src/kd_tree.c:166-183) */
    tri_set L = tri_set_alloc(s.n), R = tri_set_alloc(s.n);
    double vL = (double)best_v + KD_EPS, vR = (double)best_v - KD_EPS;
    const float *lo = s.lo[best_axis], *hi = s.hi[best_axis];
    for (int t = 0; t < s.n; t++) {
        if ((double)lo[t] <= vL) {
            tri_set_push(&L, &s, t);
        }
        if ((double)hi[t] >= vR) {
            tri_set_push(&R, &s, t);
        }
    }
    int big = s.n >= 1024;
    tri_set_free(&s);
    float lmax[3], rmin[3];
    memcpy(lmax, bmax, sizeof(lmax));
    memcpy(rmin, bmin, sizeof(rmin));
    lmax[best_axis] = rmin[best_axis] = best_v;

    out = arena_alloc(sizeof(*out));
    memcpy(out->bmin, bmin, sizeof(out->bmin));
    memcpy(out->bmax, bmax, sizeof(out->bmax));
    out->leaf = 0;
    out->value = best_v;
    out->axis = best_axis;
    out->ids = NULL;
    out->nids = 0;
    bnode *kl = NULL, *kr = NULL;
#pragma omp task shared(kl) firstprivate(L, depth, nbins) if (big)
    kl = build_cell(L, bmin, lmax, depth - 1, nbins);
    kr = build_cell(R, rmin, bmax, depth - 1, nbins);
#pragma omp taskwait
    out->kid[0] = kl;
    out->kid[1] = kr;
    out->sub_nodes = 1 + kl->sub_nodes + kr->sub_nodes;
    out->sub_refs = kl->sub_refs + kr->sub_refs;
    return out;
}


/* ------------------------------------------------------------------ SAH build
 * An ADDITION with no counterpart in the reference:
 * the reference heuristic has no leaf-cost term and a hard depth cap, which at
 * 1M triangles leaves either ~56 triangles per leaf (depth 15) or millions of
 * thin and empty cells (deep).  build_kd_sah uses the standard surface-area
 * heuristic
 *     cost(plane) = Ct + Ci * (SA(L) * NL + SA(R) * NR) / SA(cell)
 * (x empty_bonus when one side is empty), makes a leaf when no plane beats
 * Ci * N, and evaluates as candidate planes every triangle bound inside the cell
 * (sorted sweep, O(n log n) per cell and axis).  With nbins > 0, cells with more
 * than SAH_EXACT_BELOW triangles use a uniform grid of `nbins` planes per axis
 * instead (histograms, O(n + bins)); measured on the 1M-triangle heightfield the
 * exact sweep builds as fast and gives a tree with 12% fewer leaf visits and 44%
 * fewer triangle references (planes on triangle bounds do not cut triangles).
 * Triangle bounds are clipped to the cell as they are handed down.  The
 * preorder wire format and the ropes are the same as above, so the output is
 * consumed by the same traversal.
 */
#ifndef SAH_EXACT_BELOW
#define SAH_EXACT_BELOW 48
#endif

typedef struct sah_params {
    int max_depth, nbins;
    float ct, ci, empty_bonus;
    int exact_below; /* cells with at most this many triangles evaluate every triangle bound */
    int clip;        /* re-derive a straddler's bounds from the triangle clipped to each child */
    const cl_int3 *corners;
    const Vector3 *verts;
} sah_params;

/* Bounds of triangle `id` clipped to the box [cmin, cmax] ("perfect splits", Wald &
 * Havran 2006): Sutherland-Hodgman in double against the six faces, then the polygon's
 * extent, widened outward by one float ulp and intersected with the bounds handed in.
 * A triangle's box overlaps far more cells than the triangle does when it is long and
 * thin; tight bounds put the next candidate planes where the surface really is.
 * Returns 0 (bounds untouched) if nothing is left of the polygon, which can only be a
 * rounding artefact because every reference in a cell came from a clipped parent. */
static int
clip_bounds(const sah_params *P, int id, const float *cmin, const float *cmax, float lo[3], float hi[3]) {
    double a[10][3], b[10][3];
    double (*src)[3] = a, (*dst)[3] = b;
    int np = 3;
    for (int k = 0; k < 3; k++) {
        const Vector3 v = P->verts[P->corners[3 * id + k].s[0]];
        for (int c = 0; c < 3; c++) {
            src[k][c] = v.s[c];
        }
    }
    double tmin[3], tmax[3]; /* the whole triangle's extent: faces it does not reach clip nothing */
    for (int c = 0; c < 3; c++) {
        tmin[c] = fmin(fmin(src[0][c], src[1][c]), src[2][c]);
        tmax[c] = fmax(fmax(src[0][c], src[1][c]), src[2][c]);
    }
    for (int axis = 0; axis < 3 && np > 0; axis++) {
        for (int side = 0; side < 2 && np > 0; side++) {
            const double plane = side ? cmax[axis] : cmin[axis];
            if (side ? tmax[axis] <= plane : tmin[axis] >= plane) {
                continue;
            }
            int nq = 0;
            for (int k = 0; k < np; k++) {
                const double *p = src[k], *q = src[(k + 1) % np];
                const int pin = side ? p[axis] <= plane : p[axis] >= plane;
                const int qin = side ? q[axis] <= plane : q[axis] >= plane;
                if (pin) {
                    memcpy(dst[nq++], p, sizeof(double) * 3);
                }
                if (pin != qin) {
                    const double t = (plane - p[axis]) / (q[axis] - p[axis]);
                    for (int c = 0; c < 3; c++) {
                        dst[nq][c] = p[c] + t * (q[c] - p[c]);
                    }
                    dst[nq++][axis] = plane;
                }
            }
            double (*tmp)[3] = src;
            src = dst;
            dst = tmp;
            np = nq;
        }
    }
    if (np == 0) {
        return 0;
    }
    for (int c = 0; c < 3; c++) {
        double mn = src[0][c], mx = src[0][c];
        for (int k = 1; k < np; k++) {
            mn = src[k][c] < mn ? src[k][c] : mn;
            mx = src[k][c] > mx ? src[k][c] : mx;
        }
        float fl = nextafterf((float)mn, -INFINITY), fh = nextafterf((float)mx, INFINITY);
        if (fl > lo[c]) lo[c] = fl;
        if (fh < hi[c]) hi[c] = fh;
        if (lo[c] > hi[c]) { /* cannot happen for a real overlap; stay safe */
            lo[c] = hi[c] = (float)(0.5 * (mn + mx));
        }
    }
    return 1;
}

static int
cmp_float(const void *a, const void *b) {
    float x = *(const float *)a, y = *(const float *)b;
    return (x > y) - (x < y);
}

static inline float
box_area(const float *e) {
    return 2.0f * (e[0] * e[1] + e[1] * e[2] + e[2] * e[0]);
}

static inline float
sah_cost(const sah_params *P, const float *ext, int axis, float bmin_a, float v, int NL, int NR, float inv_area) {
    float el[3] = { ext[0], ext[1], ext[2] }, er[3] = { ext[0], ext[1], ext[2] };
    el[axis] = v - bmin_a;
    er[axis] = ext[axis] - el[axis];
    float c = P->ct + P->ci * (box_area(el) * (float)NL + box_area(er) * (float)NR) * inv_area;
    if (NL == 0 || NR == 0) {
        c *= P->empty_bonus;
    }
    return c;
}

/* Best candidate plane of one axis: the first plane, in candidate order, with the
 * lowest cost below *best_cost.  Returns 1 and updates *best_cost / *best_v if one exists. */
static int
sah_axis_best(const tri_set *s, const float *bmin, const float *bmax, const float *ext, float inv_area,
              int axis, const sah_params *P, float *best_cost_io, float *best_v_io) {
    const int n = s->n;
    float best_cost = *best_cost_io, best_v = *best_v_io;
    const float cost_in = best_cost;
    const float *lo = s->lo[axis], *hi = s->hi[axis];
    if (n > P->exact_below) {
        /* uniform planes v_i = min + (i+1)/(K+1) * extent; histogram of the
         * first plane each triangle is "left of" and the last it is "right of" */
        const int K = P->nbins;
        float *v = xmalloc(sizeof(float) * (size_t)K);
        int *hl = calloc((size_t)K + 1, sizeof(int)), *hr = calloc((size_t)K + 1, sizeof(int));
        if (!hl || !hr) {
            perror("calloc");
            exit(EXIT_FAILURE);
        }
        for (int i = 0; i < K; i++) {
            v[i] = bmin[axis] + ((float)(i + 1) / (float)(K + 1)) * ext[axis];
        }
        const float scale = (float)(K + 1) / ext[axis];
        for (int t = 0; t < n; t++) {
            /* il = smallest i with GOES_LEFT(t, v_i) (K if none); a triangle flat
             * on this axis (lo == hi) sits left of a plane through it */
            const int flat = lo[t] == hi[t];
            int il = (int)((lo[t] - bmin[axis]) * scale) - 1;
            il = il < 0 ? 0 : (il > K ? K : il);
            while (il > 0 && (lo[t] < v[il - 1] || (flat && lo[t] == v[il - 1]))) il--;
            while (il < K && !(lo[t] < v[il] || (flat && lo[t] == v[il]))) il++;
            hl[il]++;
            /* ir = largest i with hi > v_i (-1 if none), stored shifted by one */
            int ir = (int)((hi[t] - bmin[axis]) * scale) - 1;
            ir = ir < -1 ? -1 : (ir > K - 1 ? K - 1 : ir);
            while (ir < K - 1 && hi[t] > v[ir + 1]) ir++;
            while (ir >= 0 && !(hi[t] > v[ir])) ir--;
            hr[ir + 1]++;
        }
        int NL = 0, NR = n - hr[0];
        for (int i = 0; i < K; i++) {
            NL += hl[i];
            /* NR(i) = triangles whose last right-plane is >= i */
            if (v[i] > bmin[axis] && v[i] < bmax[axis]) {
                float c = sah_cost(P, ext, axis, bmin[axis], v[i], NL, NR, inv_area);
                if (c < best_cost) {
                    best_cost = c;
                                        best_v = v[i];
                }
            }
            NR -= hr[i + 1];
        }
        free(v);
        free(hl);
        free(hr);
    } else {
        /* Every triangle bound strictly inside the cell is a candidate.  The
         * bounds are sorted once and the counts follow by a sweep:
         *   NL(v) = #{lo < v} + #{flat triangles lying in the plane}
         *   NR(v) = #{hi > v} */
        float Lbuf[64], Hbuf[64], Fbuf[64];
        float *L = Lbuf, *H = Hbuf, *F = Fbuf;
        if (n > 64) {
            L = xmalloc(sizeof(float) * 3 * (size_t)n);
            H = L + n;
            F = H + n;
        }
        int nf = 0;
        if (n > 64) {
            memcpy(L, lo, sizeof(float) * (size_t)n);
            memcpy(H, hi, sizeof(float) * (size_t)n);
            for (int t = 0; t < n; t++) {
                if (lo[t] == hi[t]) F[nf++] = lo[t];
            }
            qsort(L, (size_t)n, sizeof(float), cmp_float);
            qsort(H, (size_t)n, sizeof(float), cmp_float);
            qsort(F, (size_t)nf, sizeof(float), cmp_float);
        } else
        for (int t = 0; t < n; t++) { /* insertion sorts for small cells */
            float x = lo[t];
            int k = t;
            while (k > 0 && L[k - 1] > x) { L[k] = L[k - 1]; k--; }
            L[k] = x;
            x = hi[t];
            k = t;
            while (k > 0 && H[k - 1] > x) { H[k] = H[k - 1]; k--; }
            H[k] = x;
            if (lo[t] == hi[t]) {
                x = lo[t];
                k = nf++;
                while (k > 0 && F[k - 1] > x) { F[k] = F[k - 1]; k--; }
                F[k] = x;
            }
        }
        int il = 0, ih = 0;      /* read positions of the merged candidate stream */
        int lo_less = 0, hi_le = 0, f0 = 0;
        while (il < n || ih < n) {
            float v;
            if (ih >= n || (il < n && L[il] <= H[ih])) v = L[il++]; else v = H[ih++];
            while (il < n && L[il] == v) il++; /* skip duplicates of this value */
            while (ih < n && H[ih] == v) ih++;
            if (!(v > bmin[axis] && v < bmax[axis])) {
                continue;
            }
            while (lo_less < n && L[lo_less] < v) lo_less++;
            while (hi_le < n && H[hi_le] <= v) hi_le++;
            while (f0 < nf && F[f0] < v) f0++;
            int flat_here = 0;
            while (f0 + flat_here < nf && F[f0 + flat_here] == v) flat_here++;
            const int NL = lo_less + flat_here, NR = n - hi_le;
            float c = sah_cost(P, ext, axis, bmin[axis], v, NL, NR, inv_area);
            if (c < best_cost) {
                best_cost = c;
                                best_v = v;
            }
        }
        if (L != Lbuf) free(L);
    }
    if (best_cost < cost_in) {
        *best_cost_io = best_cost;
        *best_v_io = best_v;
        return 1;
    }
    return 0;
}

static bnode *
build_cell_sah(tri_set s, const float *bmin, const float *bmax, int depth, const sah_params *P) {
    bnode *out;
    if (s.n == 0 || depth == 0) {
        out = make_leaf(bmin, bmax, &s);
        tri_set_free(&s);
        return out;
    }
    float ext[3];
    for (int a = 0; a < 3; a++) {
        ext[a] = bmax[a] - bmin[a];
    }
    const float area = box_area(ext);
    const float inv_area = area > 0 ? 1.0f / area : 0.0f;
    const float leaf_cost = P->ci * (float)s.n;
    float best_cost = leaf_cost;
    int best_axis = -1;
    float best_v = 0;
    const int n = s.n;

    /* The three axes are independent searches, combined in axis order with a strict
     * comparison: the plane a single sweep over (axis, candidate) would keep.  (Running
     * them as OpenMP tasks in the big cells at the top of the tree was measured: slower,
     * 29 against 20 ms for 50k triangles on 8 cores.) */
    float axis_cost[3] = { leaf_cost, leaf_cost, leaf_cost }, axis_v[3] = { 0, 0, 0 };
    int axis_found[3] = { 0, 0, 0 };
    for (int axis = 0; axis < 3 && inv_area > 0; axis++) {
        if (ext[axis] >= KD_EPS) {
            axis_found[axis] = sah_axis_best(&s, bmin, bmax, ext, inv_area, axis, P, &axis_cost[axis], &axis_v[axis]);
        }
    }
    for (int axis = 0; axis < 3; axis++) {
        if (axis_found[axis] && axis_cost[axis] < best_cost) {
            best_cost = axis_cost[axis];
            best_axis = axis;
            best_v = axis_v[axis];
        }
    }
    if (best_axis < 0) {
        out = make_leaf(bmin, bmax, &s);
        tri_set_free(&s);
        return out;
    }
    /* Partition.  Unlike the reference rule (both sides within 1e-9 of the plane),
     * a triangle that only TOUCHES the plane stays on its own side, and one lying
     * in the plane goes left: on meshes whose vertices sit on the candidate planes
     * the reference rule duplicates every triangle adjacent to a plane.
     * Consequence: a ray lying EXACTLY in a split plane (the traversal sends
     * p == plane to the low child) meets only the low side's copy of an edge that
     * lies in that plane; whether that copy or the high side's wins is an exact
     * edge-graze tie either way.  With candidate planes on triangle bounds
     * (nbins <= 0) and a mesh-aligned, centred camera this shows up as ties along
     * one pixel column; listing such triangles on both sides was tried and costs
     * 2x the triangle tests (the copies are carried down every level). */
    tri_set L = tri_set_alloc(n), R = tri_set_alloc(n);
    const float *lo = s.lo[best_axis], *hi = s.hi[best_axis];
    float lmax[3], rmin[3];
    memcpy(lmax, bmax, sizeof(lmax));
    memcpy(rmin, bmin, sizeof(rmin));
    lmax[best_axis] = rmin[best_axis] = best_v;
    for (int t = 0; t < n; t++) {
        const int straddles = lo[t] < best_v && hi[t] > best_v;
        if (lo[t] < best_v || (lo[t] == hi[t] && lo[t] == best_v)) {
            tri_set_push(&L, &s, t);
            /* clip the handed-down bound to the child */
            if (L.hi[best_axis][L.n - 1] > best_v) L.hi[best_axis][L.n - 1] = best_v;
            if (straddles && P->clip) {
                float b0[3], b1[3];
                for (int a = 0; a < 3; a++) { b0[a] = L.lo[a][L.n - 1]; b1[a] = L.hi[a][L.n - 1]; }
                if (clip_bounds(P, s.id[t], bmin, lmax, b0, b1)) {
                    for (int a = 0; a < 3; a++) { L.lo[a][L.n - 1] = b0[a]; L.hi[a][L.n - 1] = b1[a]; }
                }
            }
        }
        if (hi[t] > best_v) {
            tri_set_push(&R, &s, t);
            if (R.lo[best_axis][R.n - 1] < best_v) R.lo[best_axis][R.n - 1] = best_v;
            if (straddles && P->clip) {
                float b0[3], b1[3];
                for (int a = 0; a < 3; a++) { b0[a] = R.lo[a][R.n - 1]; b1[a] = R.hi[a][R.n - 1]; }
                if (clip_bounds(P, s.id[t], rmin, bmax, b0, b1)) {
                    for (int a = 0; a < 3; a++) { R.lo[a][R.n - 1] = b0[a]; R.hi[a][R.n - 1] = b1[a]; }
                }
            }
        }
    }
    /* a split that separates nothing and removes no volume would recurse forever */
    if (L.n == n && R.n == n) {
        tri_set_free(&L);
        tri_set_free(&R);
        out = make_leaf(bmin, bmax, &s);
        tri_set_free(&s);
        return out;
    }
    int big = n >= 1024;
    tri_set_free(&s);
    out = arena_alloc(sizeof(*out));
    memcpy(out->bmin, bmin, sizeof(out->bmin));
    memcpy(out->bmax, bmax, sizeof(out->bmax));
    out->leaf = 0;
    out->value = best_v;
    out->axis = best_axis;
    out->ids = NULL;
    out->nids = 0;
    bnode *kl = NULL, *kr = NULL;
#pragma omp task shared(kl) firstprivate(L, depth, P) if (big)
    kl = build_cell_sah(L, bmin, lmax, depth - 1, P);
    kr = build_cell_sah(R, rmin, bmax, depth - 1, P);
#pragma omp taskwait
    out->kid[0] = kl;
    out->kid[1] = kr;
    out->sub_nodes = 1 + kl->sub_nodes + kr->sub_nodes;
    out->sub_refs = kl->sub_refs + kr->sub_refs;
    return out;
}

typedef struct emitter {
    kdnode *nodes;
    int *refs;
} emitter;

/* Preorder emission (the left child is always index+1, src/kd_tree.c).  Every subtree
 * knows its size, so its block of nodes and of triangle references starts at a known
 * offset and large subtrees are written as parallel tasks; the arrays are the ones a
 * sequential walk produces. */
static void
emit_preorder(const emitter *em, const bnode *b, int me, int ref_at) {
    kdnode *n = &em->nodes[me];
    memset(n, 0, sizeof(*n));
    for (int a = 0; a < 3; a++) {
        n->min.s[a] = b->bmin[a];
        n->max.s[a] = b->bmax[a];
    }
    if (b->leaf) {
        n->type = KD_LEAF;
        n->leaf.tris = ref_at;
        n->leaf.tri_count = b->nids;
        for (int f = 0; f < 6; f++) {
            n->leaf.ropes[f] = -1;
        }
        memcpy(em->refs + ref_at, b->ids, sizeof(int) * (size_t)b->nids);
        return;
    }
    const int l = me + 1, r = me + 1 + b->kid[0]->sub_nodes;
    n->type = KD_SPLIT;
    n->split.value = b->value;
    n->split.axis = b->axis;
    n->split.children[0] = l;
    n->split.children[1] = r;
    const int big = b->sub_nodes >= 4096;
#pragma omp task firstprivate(em, b, l, ref_at) if (big)
    emit_preorder(em, b->kid[0], l, ref_at);
    emit_preorder(em, b->kid[1], r, ref_at + b->kid[0]->sub_refs);
#pragma omp taskwait
}

/* Ropes (src/kd_tree.c:43-83).  Walking down from the root, each cell carries
 * the six neighbour links of its box.  Before they are handed to the children
 * every link is pushed down the neighbour's subtree for as long as the
 * neighbour is split on an axis other than the face's own and the split plane
 * does not cut this cell's extent on that axis; then the two children link to
 * each other across the split plane. */
static void
push_down_link(const kdnode *nodes, const kdnode *cell, int face, int *link, int full) {
    while (*link != -1 && nodes[*link].type != KD_LEAF) {
        const kdnode *nb = &nodes[*link];
        int ax = nb->split.axis;
        if (face / 2 == ax) {
            if (!full) {
                return; /* the reference stops here (src/kd_tree.c:49-51) */
            }
            /* A split parallel to the face: only the child touching the face can
             * be entered through it -- the low child when the neighbour lies above
             * this cell (face on our max side), the high child otherwise. */
            *link = nb->split.children[(face & 1) ? 0 : 1];
            continue;
        }
        float plane = nb->split.value;
        if (plane >= cell->max.s[ax]) {
            *link = nb->split.children[0];
        } else if (plane <= cell->min.s[ax]) {
            *link = nb->split.children[1];
        } else {
            return;
        }
    }
}

/* full = 0: the reference's rope construction, byte for byte.  full = 1 (SAH
 * builder only): links are also pushed through splits parallel to their face and
 * once more at the leaves, so a ray that leaves a leaf lands deeper in the
 * neighbour's subtree and re-descends less. */
static void
link_cells(kdnode *nodes, int index, int links[6], int full) {
    kdnode *cell = &nodes[index];
    if (cell->type == KD_LEAF) {
        if (full) {
            for (int f = 0; f < 6; f++) {
                push_down_link(nodes, cell, f, &links[f], 1);
            }
        }
        memcpy(cell->leaf.ropes, links, sizeof(int) * 6);
        return;
    }
    for (int f = 0; f < 6; f++) {
        push_down_link(nodes, cell, f, &links[f], full);
    }
    int ax = cell->split.axis;
    int lo_child = cell->split.children[0], hi_child = cell->split.children[1];
    int lo_links[6], hi_links[6];
    memcpy(lo_links, links, sizeof(lo_links));
    memcpy(hi_links, links, sizeof(hi_links));
    lo_links[2 * ax + 1] = hi_child; /* max face of the low child */
    hi_links[2 * ax] = lo_child;     /* min face of the high child */
    /* preorder: the low child's subtree is the index range [lo_child, hi_child).  Subtrees
     * only write their own leaves' ropes and read split planes, so they can run as tasks. */
    const int big = hi_child - lo_child >= 2048;
#pragma omp task firstprivate(nodes, lo_child, lo_links, full) if (big)
    link_cells(nodes, lo_child, lo_links, full);
    link_cells(nodes, hi_child, hi_links, full);
#pragma omp taskwait
}

int
write_kd(const char *filename, const kd *tree) {
    /* [size_t n][n records] for nodes(68) verts(16) norms(16) tri_indices(4)
     * tris(16), host endian, no magic (src/kd_tree.c:250-271) */
    FILE *f = fopen(filename, "wb");
    if (f == NULL) {
        perror(filename);
        return 1;
    }
    const void *blk[5] = { tree->node_vec, tree->vert_vec, tree->norm_vec,
                           tree->tri_indices, tree->tri_vec };
    const size_t esz[5] = { sizeof(kdnode), sizeof(Vector4), sizeof(Vector4),
                            sizeof(int), sizeof(cl_int3) };
    int bad = 0;
    for (int k = 0; k < 5 && !bad; k++) {
        size_t n = blk[k] ? list_size(blk[k]) / esz[k] : 0;
        bad |= fwrite(&n, sizeof(n), 1, f) != 1;
        bad |= n && fwrite(blk[k], esz[k], n, f) != n;
    }
    bad |= fclose(f) != 0;
    return bad;
}

int
parse_kd(const char *filename, kd *tree) {
    /* reader for the layout above (src/kd_tree.c:278-311); unlike the
     * reference, I/O failures are reported instead of ignored */
    FILE *f = fopen(filename, "rb");
    if (f == NULL) {
        perror(filename);
        return 1;
    }
    void *blk[5] = { NULL, NULL, NULL, NULL, NULL };
    const size_t esz[5] = { sizeof(kdnode), sizeof(Vector4), sizeof(Vector4),
                            sizeof(int), sizeof(cl_int3) };
    int bad = 0;
    for (int k = 0; k < 5 && !bad; k++) {
        size_t n = 0;
        if (fread(&n, sizeof(n), 1, f) != 1) {
            bad = 1;
            break;
        }
        blk[k] = init_list(n, esz[k]);
        if (n && fread(blk[k], esz[k], n, f) != n) {
            bad = 1;
        }
    }
    fclose(f);
    if (bad) {
        fprintf(stderr, "%s: truncated .kd file\n", filename);
        for (int k = 0; k < 5; k++) {
            delete_list(blk[k]);
        }
        return 1;
    }
    tree->node_vec = blk[0];
    tree->vert_vec = blk[1];
    tree->norm_vec = blk[2];
    tree->tri_indices = blk[3];
    tree->tri_vec = blk[4];
    return 0;
}

void
delete_kd(kd tree) {
    delete_list(tree.node_vec);
    delete_list(tree.tri_vec);
    delete_list(tree.norm_vec);
    delete_list(tree.vert_vec);
    delete_list(tree.tri_indices);
}

static void
depth_walk(const kdnode *nodes, int index, int depth, int *max_depth) {
    if (depth > *max_depth) {
        *max_depth = depth;
    }
    if (nodes[index].type == KD_SPLIT) {
        depth_walk(nodes, nodes[index].split.children[0], depth + 1, max_depth);
        depth_walk(nodes, nodes[index].split.children[1], depth + 1, max_depth);
    }
}

void
kd_get_stats(const kd *tree, kd_stats *out) {
    memset(out, 0, sizeof(*out));
    size_t n = vector_length(tree->node_vec);
    out->node_count = (long long)n;
    for (size_t i = 0; i < n; i++) {
        const kdnode *k = &tree->node_vec[i];
        if (k->type == KD_LEAF) {
            out->leaf_count++;
            out->leaf_tri_refs += k->leaf.tri_count;
            out->empty_leaves += k->leaf.tri_count == 0;
            if (k->leaf.tri_count > out->max_leaf_tris) {
                out->max_leaf_tris = k->leaf.tri_count;
            }
        }
    }
    if (n) {
        depth_walk(tree->node_vec, 0, 0, &out->max_depth);
    }
}

static double
kd_now_ms(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

static kd
build_tree(cl_int3 *tris, Vector3 *verts, Vector3 *norms, const char *path,
           int depth, int nbins, const sah_params *sah) {
    size_t ncorners = vector_length(tris);
    size_t ntris = ncorners / 3;
    kd tree = { NULL, NULL, verts, norms, tris };
    if (ntris == 0) {
        /* the reference would read verts[tris[0]] out of bounds here; emit a
         * single empty leaf instead */
        tree.node_vec = init_list(1, sizeof(kdnode));
        memset(tree.node_vec, 0, sizeof(kdnode));
        tree.node_vec[0].type = KD_LEAF;
        for (int f = 0; f < 6; f++) {
            tree.node_vec[0].leaf.ropes[f] = -1;
        }
        tree.tri_indices = new_list(0);
        return tree;
    }
    const int timing = getenv("CLPT_BUILD_TIMING") != NULL;
    const double t0 = kd_now_ms();
    tri_set root = tri_set_alloc((int)ntris);
    Vector3 bmin, bmax;
    bmin = bmax = verts[tris[0].s[0]];
#pragma omp parallel for schedule(static) if (ntris >= 4096)
    for (long long i = 0; i < (long long)ntris; i++) {
        Vector3 A = verts[tris[3 * i + 0].s[0]], B = verts[tris[3 * i + 1].s[0]],
                C = verts[tris[3 * i + 2].s[0]];
        Vector3 tmin = vec_min(vec_min(A, B), C), tmax = vec_max(vec_max(A, B), C);
        Vector3 N = vec_cross(vec_subtract(B, A), vec_subtract(C, A));
        root.id[i] = (int)i;
        for (int a = 0; a < 3; a++) {
            root.lo[a][i] = tmin.s[a];
            root.hi[a][i] = tmax.s[a];
        }
        root.area[i] = vec_length(N) / 2;
    }
    /* the scene box in list order, as the reference accumulates it (min/max of +0 and -0
     * depend on the order, and the box is part of the node bytes) */
    for (int a = 0; a < 3; a++) { /* vec_min / vec_max, component by component */
        float mn = bmin.s[a], mx = bmax.s[a];
        const float *lo = root.lo[a], *hi = root.hi[a];
        for (size_t i = 0; i < ntris; i++) {
            mn = mn < lo[i] ? mn : lo[i];
            mx = mx > hi[i] ? mx : hi[i];
        }
        bmin.s[a] = mn;
        bmax.s[a] = mx;
    }
    bmin.s[3] = bmax.s[3] = 0;
    root.n = (int)ntris;

    const double t1 = kd_now_ms();
    bnode *top = NULL;
    const int team = arena_team();
#pragma omp parallel num_threads(team)
#pragma omp single
    top = sah ? build_cell_sah(root, bmin.s, bmax.s, sah->max_depth, sah)
              : build_cell(root, bmin.s, bmax.s, depth, nbins);

    const double t2 = kd_now_ms();
    const size_t nnodes = (size_t)top->sub_nodes, nrefs = (size_t)top->sub_refs;
    emitter em;
    em.nodes = init_list(nnodes, sizeof(kdnode));
    em.refs = init_list(nrefs, sizeof(int));
#pragma omp parallel num_threads(team)
#pragma omp single
    emit_preorder(&em, top, 0, 0);
    const double t3 = kd_now_ms();
    int links[6] = { -1, -1, -1, -1, -1, -1 };
#pragma omp parallel num_threads(team)
#pragma omp single
    link_cells(em.nodes, 0, links, sah != NULL);
    if (timing) {
        fprintf(stderr, "build_kd: bounds %.2f ms, cells %.2f ms, emit %.2f ms, ropes %.2f ms (%zu nodes)\n", t1 - t0,
                t2 - t1, t3 - t2, kd_now_ms() - t3, nnodes);
    }
    tree.node_vec = em.nodes;
    tree.tri_indices = em.refs;
    arena_release();

    if (path != NULL) {
        size_t len = strlen(path) + 4;
        char *kdpath = xmalloc(len);
        snprintf(kdpath, len, "%s.kd", path);
        write_kd(kdpath, &tree);
        free(kdpath);
    }
    return tree;
}

kd
build_kd_ex(cl_int3 *tris, Vector3 *verts, Vector3 *norms, const char *path,
            int depth, int nbins) {
    return build_tree(tris, verts, norms, path, depth, nbins, NULL);
}

static int g_sah_clip = 1;

void
kd_set_sah_clip(int enable) {
    g_sah_clip = enable != 0;
}

kd
build_kd_sah(cl_int3 *tris, Vector3 *verts, Vector3 *norms, const char *path,
             int max_depth, int nbins, float traversal_cost, float intersect_cost,
             float empty_bonus) {
    /* nbins <= 0: every triangle bound is a candidate at every cell size (sorted sweep) */
    sah_params P = { max_depth, nbins, traversal_cost, intersect_cost, empty_bonus,
                     nbins > 0 ? SAH_EXACT_BELOW : 0x7fffffff, g_sah_clip, tris, verts };
    const char *e = getenv("CLPT_SAH_CLIP"); /* overrides kd_set_sah_clip, for experiments */
    if (e) P.clip = atoi(e);
    return build_tree(tris, verts, norms, path, 0, 0, &P);
}

kd
build_kd(cl_int3 *tris, Vector3 *verts, Vector3 *norms, const char *path) {
    return build_kd_ex(tris, verts, norms, path, KD_REF_DEPTH, KD_REF_NBINS);
}
