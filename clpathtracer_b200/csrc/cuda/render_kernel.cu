// render_kernel.cu -- the render hot path for sm_100a.
//
// One launch = one frame (or one rank's row tiles of it): per pixel sample,
// camera ray generation -> stackless kd-tree traversal with ropes ->
// Moller-Trumbore over the leaf's contiguous triangle run -> shading ->
// in-register accumulation over the samples -> one float4 store.
//
// What it computes is what the reference kernel computes (src/kernel.cl:296-473,
// see oracle/oracle_kernel.c for the restatement it is checked against); how
// it computes it is not: the camera matrix comes from the constant bank
// instead of global memory, split nodes are 8 bytes and siblings adjacent,
// leaves are one aligned 64-byte record, triangles are pre-gathered per leaf
// with the edge vectors precomputed, the hit normal is evaluated once per ray
// instead of once per accepted candidate, the tail recursion is a loop, and the
// lanes of a warp trace samples of the SAME pixel (see "Lane mapping" below).
//
// Numerics: every fp32 operation that decides a hit is written with the
// round-to-nearest intrinsics (__fmul_rn, __fadd_rn, ...), which the compiler
// never contracts into FMAs, and in the reference's operand order, so hit
// ids, t, u, v and colours are bit-identical to the restatement compiled with
// -ffp-contract=off.  Division and square root are the IEEE-rounded ones.
#include "clpt_device.cuh"

#ifndef CLPT_MIN_BLOCKS
#define CLPT_MIN_BLOCKS 8 // resident 256-thread blocks per SM the register allocation aims for (measured: 3 -> 2029, 4 -> 2488, 6 -> 2921, 8 -> 3054 Mrays/s)
#endif
#ifndef CLPT_PREFETCH_ROPE
#define CLPT_PREFETCH_ROPE 1
#endif
#ifndef CLPT_BRANCHLESS_TRI
#define CLPT_BRANCHLESS_TRI 0
#endif

namespace {

struct V3 {
    float x, y, z;
};

__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
// 1/x: the correctly rounded reciprocal is the same number as the IEEE quotient 1.0f/x
__device__ __forceinline__ float frcp(float a) { return __frcp_rn(a); }
__device__ __forceinline__ V3 mk(float x, float y, float z) { V3 r = { x, y, z }; return r; }
__device__ __forceinline__ V3 vadd(V3 a, V3 b) { return mk(fadd(a.x, b.x), fadd(a.y, b.y), fadd(a.z, b.z)); }
__device__ __forceinline__ V3 vsub(V3 a, V3 b) { return mk(fsub(a.x, b.x), fsub(a.y, b.y), fsub(a.z, b.z)); }
__device__ __forceinline__ V3 vscale(V3 a, float k) { return mk(fmul(a.x, k), fmul(a.y, k), fmul(a.z, k)); }
__device__ __forceinline__ float vdot(V3 a, V3 b) {
    return fadd(fadd(fmul(a.x, b.x), fmul(a.y, b.y)), fmul(a.z, b.z));
}
__device__ __forceinline__ V3 vcross(V3 a, V3 b) {
    return mk(fsub(fmul(a.y, b.z), fmul(a.z, b.y)), fsub(fmul(a.z, b.x), fmul(a.x, b.z)),
              fsub(fmul(a.x, b.y), fmul(a.y, b.x)));
}
__device__ __forceinline__ V3 vnormalize(V3 a) {
    float len = __fsqrt_rn(vdot(a, a));
    return mk(fdiv(a.x, len), fdiv(a.y, len), fdiv(a.z, len));
}
__device__ __forceinline__ V3 xyz(float4 a) { return mk(a.x, a.y, a.z); }

// What survives the traversal loop is only the winning triangle SLOT and its t:
// primitive id, u and v are re-derived afterwards (hit_details) by the same
// arithmetic on the same operands, which keeps the loop's register footprint low.
struct Hit {
    int ref; // triangle slot of the accepted hit, -1 = none
    float t;
};

struct HitDetails {
    int prim;
    float u, v;
};

struct Counters {
    unsigned int rays, splits, leaves, tris, shade_vn, capped;
};

// Traversal of one ray: src/kernel.cl:311-389.
template <bool COUNT>
__device__ __forceinline__ Hit closest_hit(const ClptScene &S, V3 o, V3 d, int max_visits,
                                           Counters &cn) {
    Hit h;
    h.ref = -1;
    h.t = 0.0f;
    if (COUNT) cn.rays++;

    const V3 inv = mk(frcp(d.x), frcp(d.y), frcp(d.z));
    const bool sx = inv.x < 0.0f, sy = inv.y < 0.0f, sz = inv.z < 0.0f;

    float tmin, tmax;
    { // root clip, kernel.cl:101-144
        const float nx = sx ? S.root_max[0] : S.root_min[0], fx = sx ? S.root_min[0] : S.root_max[0];
        const float ny = sy ? S.root_max[1] : S.root_min[1], fy = sy ? S.root_min[1] : S.root_max[1];
        const float nz = sz ? S.root_max[2] : S.root_min[2], fz = sz ? S.root_min[2] : S.root_max[2];
        tmin = fmul(fsub(nx, o.x), inv.x);
        tmax = fmul(fsub(fx, o.x), inv.x);
        const float tymin = fmul(fsub(ny, o.y), inv.y), tymax = fmul(fsub(fy, o.y), inv.y);
        if ((tmin > tymax) || (tymin > tmax)) return h;
        if (tymin > tmin) tmin = tymin;
        if (tymax < tmax) tmax = tymax;
        const float tzmin = fmul(fsub(nz, o.z), inv.z), tzmax = fmul(fsub(fz, o.z), inv.z);
        if ((tmin > tzmax) || (tzmin > tmax)) return h;
        if (tzmin > tmin) tmin = tzmin;
        if (tzmax < tmax) tmax = tzmax;
        if (!(tmax > 0.0f)) return h;
    }
    V3 p1 = o;
    if (tmin > 0.0f) p1 = vadd(p1, vscale(d, tmin));

    int index = 0;
    int visits = 0;
    float min_hit = 0.0f;
    const uint2 *__restrict__ nodes = S.nodes;
    const float4 *__restrict__ leaves = S.leaves;
    const float4 *__restrict__ tri = S.tri;

    uint2 n = __ldg(nodes);
    for (;;) {
        // descend to the leaf containing p1, kernel.cl:325-330 (selects, no branches)
        while ((n.y & 3u) != 3u) {
            const unsigned axis = n.y & 3u;
            float p = axis == 1u ? p1.y : p1.x;
            p = axis == 2u ? p1.z : p;
            index = (int)((n.y >> 2) + (p > __uint_as_float(n.x) ? 1u : 0u));
            n = __ldg(nodes + index);
            if (COUNT) cn.splits++;
        }
        if (COUNT) cn.leaves++;
        const float4 *L = leaves + 4 * (size_t)n.x;
        const float4 lmin = __ldg(L), lmax = __ldg(L + 1);
        const int first = __float_as_int(lmin.w), count = __float_as_int(lmax.w);

        // leaf slab interval and exit face, kernel.cl:146-174.  It depends only on the
        // leaf box and the ray, so it is evaluated BEFORE the triangle run: three
        // values stay live across the run instead of the eight of the box.
        int far = sx ? 0 : 1;
        {
            const float nx = sx ? lmax.x : lmin.x, fx = sx ? lmin.x : lmax.x;
            const float ny = sy ? lmax.y : lmin.y, fy = sy ? lmin.y : lmax.y;
            const float nz = sz ? lmax.z : lmin.z, fz = sz ? lmin.z : lmax.z;
            tmin = fmul(fsub(nx, o.x), inv.x);
            tmax = fmul(fsub(fx, o.x), inv.x);
            const float tymin = fmul(fsub(ny, o.y), inv.y), tymax = fmul(fsub(fy, o.y), inv.y);
            if (tymin > tmin) tmin = tymin;
            if (tymax < tmax) {
                tmax = tymax;
                far = sy ? 2 : 3;
            }
            const float tzmin = fmul(fsub(nz, o.z), inv.z), tzmax = fmul(fsub(fz, o.z), inv.z);
            if (tzmin > tmin) tmin = tzmin;
            if (tzmax < tmax) {
                tmax = tzmax;
                far = sz ? 4 : 5;
            }
        }
#if CLPT_PREFETCH_ROPE
        // The neighbour across the exit face and its node record are requested now,
        // so the two dependent loads overlap the triangle run instead of following it.
        index = __ldg(reinterpret_cast<const int *>(L + 2) + far);
        uint2 n_next = make_uint2(0u, 3u);
        if (index >= 0) n_next = __ldg(nodes + index);
#endif

        // triangle run of the leaf, kernel.cl:333-368 / 227-255
        for (int i = first; i < first + count; i++) {
            const float4 b = __ldg(tri + 3 * (size_t)i + 1);
            const float4 c = __ldg(tri + 3 * (size_t)i + 2);
            if (COUNT) cn.tris++;
            const V3 e1 = xyz(b), e2 = xyz(c);
            const V3 pvec = vcross(d, e2);
            const float det = vdot(e1, pvec);
#if CLPT_BRANCHLESS_TRI
            // Straight-line variant: the lanes of a warp rarely agree on which test
            // rejects, so the warp pays for the whole test anyway; predicating it
            // removes the divergent branches.  Same comparisons, same order.
            const float4 a = __ldg(tri + 3 * (size_t)i);
            const float idet = frcp(det);
            const V3 tvec = vsub(o, xyz(a));
            const float u = fmul(vdot(tvec, pvec), idet);
            const V3 qvec = vcross(tvec, e1);
            const float v = fmul(vdot(d, qvec), idet);
            const float t = fmul(vdot(e2, qvec), idet);
            const bool ok = !(det < 0.0f) && !(u < 0.0f || u > 1.0f) && !(v < 0.0f || fadd(u, v) > 1.0f) &&
                            (t > 0.0f);
            if (ok && (h.ref < 0 || t <= min_hit)) {
                min_hit = t;
                h.ref = i;
            }
#else
            if (det < 0.0f) continue;
            const float4 a = __ldg(tri + 3 * (size_t)i);
            const float idet = frcp(det);
            const V3 tvec = vsub(o, xyz(a));
            const float u = fmul(vdot(tvec, pvec), idet);
            if (u < 0.0f || u > 1.0f) continue;
            const V3 qvec = vcross(tvec, e1);
            const float v = fmul(vdot(d, qvec), idet);
            if (v < 0.0f || fadd(u, v) > 1.0f) continue;
            const float t = fmul(vdot(e2, qvec), idet);
            if (!(t > 0.0f)) continue;
            if (h.ref < 0 || t <= min_hit) { // the later triangle wins ties (:344)
                min_hit = t;
                h.ref = i;
            }
#endif
        }

        // 0.001 is a double literal in the reference (:381)
        if (h.ref >= 0 && (double)tmin + 0.001 > (double)min_hit) break;
#if !CLPT_PREFETCH_ROPE
        index = __ldg(reinterpret_cast<const int *>(L + 2) + far);
#endif
        p1 = vadd(o, vscale(d, tmax));
        if (index == -1) break;
        if (++visits >= max_visits) {
            if (COUNT) cn.capped++;
            break;
        }
#if CLPT_PREFETCH_ROPE
        n = n_next;
#else
        n = __ldg(nodes + index);
#endif
    }
    h.t = min_hit;
    return h;
}

// Primitive id and barycentrics of the accepted hit: hit_triangle's u and v
// (kernel.cl:243-249) evaluated again for the winning slot.
__device__ __forceinline__ HitDetails hit_details(const ClptScene &S, const Hit &h, V3 o, V3 d) {
    HitDetails r;
    const float4 a = __ldg(S.tri + 3 * (size_t)h.ref);
    const V3 e1 = xyz(__ldg(S.tri + 3 * (size_t)h.ref + 1));
    const V3 e2 = xyz(__ldg(S.tri + 3 * (size_t)h.ref + 2));
    const V3 pvec = vcross(d, e2);
    const float idet = frcp(vdot(e1, pvec));
    const V3 tvec = vsub(o, xyz(a));
    r.prim = __float_as_int(a.w);
    r.u = fmul(vdot(tvec, pvec), idet);
    r.v = fmul(vdot(d, vcross(tvec, e1)), idet);
    return r;
}

// Shading normal of an accepted hit, kernel.cl:349-365.
template <bool COUNT>
__device__ __forceinline__ V3 hit_normal(const ClptScene &S, const Hit &h, V3 o, V3 d, Counters &cn) {
    const int prim = __float_as_int(__ldg(&S.tri[3 * (size_t)h.ref].w));
    const int4 c1 = __ldg(S.corners + 3 * (size_t)prim);
    if (c1.y >= 0) {
        const HitDetails hd = hit_details(S, h, o, d);
        const int4 c2 = __ldg(S.corners + 3 * (size_t)prim + 1);
        const int4 c3 = __ldg(S.corners + 3 * (size_t)prim + 2);
        const V3 n1 = xyz(__ldg(S.norms + c1.y)), n2 = xyz(__ldg(S.norms + c2.y)),
                 n3 = xyz(__ldg(S.norms + c3.y));
        const float w = fsub(fsub(1.0f, hd.u), hd.v);
        if (COUNT) cn.shade_vn++;
        return vnormalize(vadd(vadd(vscale(n1, w), vscale(n2, hd.u)), vscale(n3, hd.v)));
    }
    const V3 e1 = xyz(__ldg(S.tri + 3 * (size_t)h.ref + 1));
    const V3 e2 = xyz(__ldg(S.tri + 3 * (size_t)h.ref + 2));
    return vnormalize(vcross(e1, e2));
}

template <bool COUNT>
__device__ __forceinline__ void write_aov(const ClptScene &S, const ClptFrame &F, const Hit &h, V3 o, V3 d,
                                          int x, int y) {
    const size_t px = (size_t)y * F.width + x;
    if (h.ref >= 0) {
        const HitDetails hd = hit_details(S, h, o, d);
        F.aov_prim[px] = hd.prim;
        F.aov_t[px] = h.t;
        F.aov_uv[px] = make_float2(hd.u, hd.v);
    } else {
        F.aov_prim[px] = -1;
        F.aov_t[px] = 0.0f;
        F.aov_uv[px] = make_float2(0.0f, 0.0f);
    }
}

// Philox4x32-10, counter (pixel, sample, dimension block, lane), key (seed, 'clpt').
__device__ __forceinline__ void philox(unsigned c[4], unsigned k0, unsigned k1) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const unsigned hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const unsigned hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const unsigned n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
        c[0] = n0;
        c[1] = lo1;
        c[2] = n2;
        c[3] = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}
__device__ __forceinline__ float u01(unsigned x) { return fmul((float)(x >> 8), 0x1p-24f); }
#define CLPT_KEY1 0x636c7074u

// Cosine-weighted direction about n (extension; oracle_kernel.c cosine_dir).
__device__ __forceinline__ V3 cosine_dir(V3 n, unsigned pixel, unsigned sample, unsigned bounce,
                                         unsigned seed) {
    float a = 0.0f, b = 0.0f;
    bool found = false;
#pragma unroll 1
    for (unsigned blk = 0; blk < 2 && !found; blk++) {
        unsigned c[4] = { pixel, sample, 1u + bounce, blk };
        philox(c, seed, CLPT_KEY1);
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const float x = fsub(fmul(2.0f, u01(c[2 * k])), 1.0f);
            const float y = fsub(fmul(2.0f, u01(c[2 * k + 1])), 1.0f);
            if (!found && fadd(fmul(x, x), fmul(y, y)) <= 1.0f) {
                a = x;
                b = y;
                found = true;
            }
        }
    }
    const float zz = fsub(fsub(1.0f, fmul(a, a)), fmul(b, b));
    const float z = __fsqrt_rn(zz > 0.0f ? zz : 0.0f);
    const float sg = n.z >= 0.0f ? 1.0f : -1.0f;
    const float p = fdiv(-1.0f, fadd(sg, n.z));
    const float q = fmul(fmul(n.x, n.y), p);
    const V3 t1 = mk(fadd(1.0f, fmul(fmul(fmul(sg, n.x), n.x), p)), fmul(sg, q), fmul(-sg, n.x));
    const V3 t2 = mk(q, fadd(sg, fmul(fmul(n.y, n.y), p)), -n.y);
    return vnormalize(vadd(vadd(vscale(t1, a), vscale(t2, b)), vscale(n, z)));
}

__device__ __forceinline__ V3 unproject(const float *M, V3 X) { // kernel.cl:89-94
    const float w = fadd(fadd(fadd(fmul(M[12], X.x), fmul(M[13], X.y)), fmul(M[14], X.z)), M[15]);
    const float a = fadd(fadd(fadd(fmul(M[0], X.x), fmul(M[1], X.y)), fmul(M[2], X.z)), M[3]);
    const float b = fadd(fadd(fadd(fmul(M[4], X.x), fmul(M[5], X.y)), fmul(M[6], X.z)), M[7]);
    const float c = fadd(fadd(fadd(fmul(M[8], X.x), fmul(M[9], X.y)), fmul(M[10], X.z)), M[11]);
    return mk(fdiv(a, w), fdiv(b, w), fdiv(c, w));
}

// Lane mapping.  A warp works on 32 / S pixels at a time, S lanes per pixel, where
// S = sample_lanes = the largest power of two <= min(spp, 32): the S lanes of a
// pixel trace S jittered samples of that pixel side by side, so the rays of a
// warp share origin, leaves and triangles (warp-coherent by construction), node
// and triangle loads collapse into broadcasts, and lanes finish together.  At
// 1 spp this degenerates to one lane per pixel over an 8x4 pixel tile.
//   pixels per warp   1     2     4     8     16    32
//   warp tile (w x h) 1x1   2x1   2x2   4x2   4x4   8x4
// A block is 8 warps laid out 4 x 2, i.e. (4*tw) x (2*th) pixels.
// The samples of a pixel are summed in ascending sample order (shuffles in a
// fixed order), which keeps the result independent of the mapping.
__host__ __device__ inline void warp_tile_dims(int log2_ppw, int &tw, int &th) {
    tw = 1 << ((log2_ppw + 1) >> 1);
    th = 1 << (log2_ppw >> 1);
}

template <int MODE, bool COUNT>
__device__ __forceinline__ V3 trace_sample(const ClptScene &S, const ClptFrame &F, int x, int y, unsigned pixel,
                                           unsigned sample, bool aov, Counters &cn) {
    const float *M = F.cam;
    const V3 origin = mk(fdiv(M[2], M[14]), fdiv(M[6], M[14]), fdiv(M[10], M[14])); // :443-445
    float fx = fsub((float)(unsigned)x, fdiv((float)(unsigned)F.width, 2.0f));
    float fy = fsub((float)(unsigned)y, fdiv((float)(unsigned)F.height, 2.0f));
    if (F.flags & CLPT_F_JITTER) {
        unsigned c[4] = { pixel, sample, 0u, 0u };
        philox(c, F.seed, CLPT_KEY1);
        fx = fadd(fx, fsub(u01(c[0]), 0.5f));
        fy = fadd(fy, fsub(u01(c[1]), 0.5f));
    }
    const V3 ncp = unproject(M, mk(fx, fy, -1.0f));
    const V3 fcp = unproject(M, mk(fx, fy, 1.0f));
    V3 o = origin;
    V3 d = vnormalize(vsub(fcp, ncp));
    if (MODE == 2) {
        V3 Lsum = mk(0.0f, 0.0f, 0.0f), T = mk(1.0f, 1.0f, 1.0f);
        for (int seg = 0; seg < F.depth; seg++) {
            const Hit h = closest_hit<COUNT>(S, o, d, F.max_leaf_visits, cn);
            if (seg == 0 && aov) write_aov<COUNT>(S, F, h, o, d, x, y);
            if (h.ref < 0) {
                Lsum = vadd(Lsum, T);
                break;
            }
            const V3 nrm = hit_normal<COUNT>(S, h, o, d, cn);
            float al[3] = { 0.5f, 0.5f, 0.5f }, em[3] = { 0.0f, 0.0f, 0.0f };
            int kind = 0;
            if (S.n_materials > 0) {
                int m = S.tri_material ? __ldg(S.tri_material + __float_as_int(__ldg(&S.tri[3 * (size_t)h.ref].w))) : 0;
                if (m < 0 || m >= S.n_materials) m = 0;
                const ClptMaterial *mp = S.materials + m;
                al[0] = mp->albedo[0]; al[1] = mp->albedo[1]; al[2] = mp->albedo[2];
                em[0] = mp->emission[0]; em[1] = mp->emission[1]; em[2] = mp->emission[2];
                kind = mp->kind;
            }
            Lsum = vadd(Lsum, mk(fmul(T.x, em[0]), fmul(T.y, em[1]), fmul(T.z, em[2])));
            T = mk(fmul(T.x, al[0]), fmul(T.y, al[1]), fmul(T.z, al[2]));
            const V3 hp = vadd(o, vscale(d, h.t));
            V3 nd;
            if (kind == 1) {
                nd = vnormalize(vsub(d, vscale(nrm, fmul(2.0f, vdot(d, nrm)))));
            } else {
                nd = cosine_dir(nrm, pixel, sample, (unsigned)seg, F.seed);
            }
            o = vadd(hp, vscale(nd, 0.0001f));
            d = nd;
        }
        return Lsum;
    }
    // modes A/B: kernel.cl:296-422 with the tail recursion as a loop
    V3 col = mk(0.0f, 0.0f, 0.0f);
    float str = 1.0f;
    int depth = F.depth;
    if (MODE == 0) depth = depth > 0 ? 1 : 0;
    const int first_depth = depth;
    for (; depth > 0; depth--) {
        const Hit h = closest_hit<COUNT>(S, o, d, F.max_leaf_visits, cn);
        if (depth == first_depth && aov) write_aov<COUNT>(S, F, h, o, d, x, y);
        if (h.ref < 0) break;
        const V3 nrm = hit_normal<COUNT>(S, h, o, d, cn);
        const V3 nc = mk(fdiv(fadd(nrm.x, 1.0f), 2.0f), fdiv(fadd(nrm.y, 1.0f), 2.0f),
                         fdiv(fadd(nrm.z, 1.0f), 2.0f));
        if (MODE == 0) return nc; // the `return` at :396
        V3 no = vadd(o, vscale(d, h.t));
        const V3 nd = vnormalize(vsub(d, vscale(nrm, fmul(2.0f, vdot(d, nrm)))));
        no = vadd(no, vscale(nd, 0.0001f));
        col = vadd(vscale(col, fsub(1.0f, str)), vscale(nc, str));
        str = fmul(str, 0.2f);
        o = no;
        d = nd;
    }
    const float k = fsub(1.0f, str); // :421
    return mk(fadd(fmul(k, col.x), str), fadd(fmul(k, col.y), str), fadd(fmul(k, col.z), str));
}

template <int MODE, bool COUNT>
__global__ void __launch_bounds__(256, CLPT_MIN_BLOCKS)
render_kernel(const __grid_constant__ ClptScene S, const __grid_constant__ ClptFrame F) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int log2_s = F.log2_sample_lanes, s_lanes = 1 << log2_s;
    int tw, th;
    warp_tile_dims(5 - log2_s, tw, th);
    const int pslot = lane >> log2_s, sslot = lane & (s_lanes - 1);
    const int x = (blockIdx.x * 4 + (warp & 3)) * tw + (pslot & (tw - 1));
    // row within this rank's slab
    const int ly = (blockIdx.y * 2 + (warp >> 2)) * th + pslot / tw;
    // slab row -> image row: tiles of tile_rows rows dealt round-robin to ranks
    const int lt = ly / F.tile_rows;
    const int y = (lt * F.nranks + F.rank) * F.tile_rows + (ly - lt * F.tile_rows);
    const bool valid = x < F.width && y < F.height && ly < F.local_rows;
    const unsigned pixel = (unsigned)(y * F.width + x);
    const int spp = F.spp < 1 ? 1 : F.spp;
    const int group_base = lane & ~(s_lanes - 1);
    Counters cn = { 0, 0, 0, 0, 0, 0 };
    V3 acc = mk(0.0f, 0.0f, 0.0f);
    for (int base = 0; base < spp; base += s_lanes) {
        const int s = base + sslot;
        V3 colour = mk(0.0f, 0.0f, 0.0f);
        if (valid && s < spp) {
            colour = trace_sample<MODE, COUNT>(S, F, x, y, pixel, F.sample_base + (unsigned)s,
                                               s == 0 && F.aov_prim != nullptr, cn);
        }
        // ordered sum over the samples of this round (every lane of the group
        // computes the same running sum)
        for (int j = 0; j < s_lanes; j++) {
            const float cx = __shfl_sync(0xffffffffu, colour.x, group_base + j);
            const float cy = __shfl_sync(0xffffffffu, colour.y, group_base + j);
            const float cz = __shfl_sync(0xffffffffu, colour.z, group_base + j);
            if (base + j < spp) acc = vadd(acc, mk(cx, cy, cz));
        }
    }
    if (valid && sslot == 0) {
        float4 *dst = F.target + (size_t)ly * F.width + x;
        if (F.flags & CLPT_F_ACCUMULATE) {
            const float4 prev = *dst;
            *dst = make_float4(fadd(prev.x, acc.x), fadd(prev.y, acc.y), fadd(prev.z, acc.z),
                               fadd(prev.w, (float)spp));
        } else if (spp == 1) {
            *dst = make_float4(acc.x, acc.y, acc.z, 1.0f);
        } else {
            const float k = fdiv(1.0f, (float)spp);
            *dst = make_float4(fmul(acc.x, k), fmul(acc.y, k), fmul(acc.z, k), 1.0f);
        }
    }
    if (COUNT) {
        unsigned v[6] = { cn.rays, cn.splits, cn.leaves, cn.tris, cn.shade_vn, cn.capped };
#pragma unroll
        for (int k = 0; k < 6; k++) {
            unsigned sum = v[k];
            for (int off = 16; off > 0; off >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, off);
            if (lane == 0 && sum) atomicAdd(F.counters + k, (unsigned long long)sum);
        }
    }
}

// Gathered slabs [rank][slab_rows][width] -> image rows.
__global__ void deinterleave_kernel(const float4 *__restrict__ gathered, float4 *__restrict__ image,
                                    int width, int height, int nranks, int tile_rows,
                                    int slab_rows) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)width * height) return;
    const int y = (int)(i / width), x = (int)(i - (size_t)y * width);
    const int t = y / tile_rows, r = t % nranks, lt = t / nranks;
    const int ly = lt * tile_rows + (y - t * tile_rows);
    image[i] = gathered[((size_t)r * slab_rows + ly) * width + x];
}

__global__ void fill_kernel(float4 *dst, size_t n, float value) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = make_float4(value, value, value, value);
}

// Progressive target (sum, count in .w) -> displayable average.
__global__ void normalise_kernel(const float4 *__restrict__ src, float4 *__restrict__ dst, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 s = src[i];
    if (s.w > 0.0f) {
        const float k = __fdiv_rn(1.0f, s.w);
        dst[i] = make_float4(__fmul_rn(s.x, k), __fmul_rn(s.y, k), __fmul_rn(s.z, k), 1.0f);
    } else {
        dst[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
}

template <int MODE>
void launch_mode(const ClptScene &scene, const ClptFrame &frame, dim3 grid, cudaStream_t stream) {
    if (frame.flags & CLPT_F_COUNTERS) {
        render_kernel<MODE, true><<<grid, 256, 0, stream>>>(scene, frame);
    } else {
        render_kernel<MODE, false><<<grid, 256, 0, stream>>>(scene, frame);
    }
}

} // namespace

void clpt_launch_render(const ClptScene &scene, const ClptFrame &frame, cudaStream_t stream) {
    int tw, th;
    warp_tile_dims(5 - frame.log2_sample_lanes, tw, th);
    const int block_w = 4 * tw, block_h = 2 * th;
    dim3 grid((frame.width + block_w - 1) / block_w, (frame.local_rows + block_h - 1) / block_h);
    if (grid.x == 0 || grid.y == 0) return;
    switch (frame.mode) {
    case 0: launch_mode<0>(scene, frame, grid, stream); break;
    case 1: launch_mode<1>(scene, frame, grid, stream); break;
    default: launch_mode<2>(scene, frame, grid, stream); break;
    }
}

void clpt_launch_deinterleave(const float4 *gathered, float4 *image, int width, int height,
                              int nranks, int tile_rows, int slab_rows, cudaStream_t stream) {
    const size_t n = (size_t)width * height;
    if (n == 0) return;
    deinterleave_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(gathered, image, width, height,
                                                                         nranks, tile_rows, slab_rows);
}

void clpt_launch_fill(float4 *dst, size_t n, float value, cudaStream_t stream) {
    if (n == 0) return;
    fill_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(dst, n, value);
}

void clpt_launch_normalise(const float4 *src, float4 *dst, size_t n, cudaStream_t stream) {
    if (n == 0) return;
    normalise_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(src, dst, n);
}

const void *clpt_render_kernel_symbol(void) {
    return reinterpret_cast<const void *>(&render_kernel<0, false>);
}
