// wavefront.cu -- the wavefront engine of the render path (modes A and B).
//
// The megakernel (render_kernel.cu) keeps a whole path -- every bounce of a
// sample -- in one thread.  On incoherent bounce rays its lanes sit idle half the
// time: ncu shows ~15 of 32 threads active, because lanes whose ray has ended
// wait for the longest traversal of the warp, five times per path.  Here the
// same arithmetic is split along the bounce axis, with the path state parked in
// HBM between steps (a few GB of queues are cheap on a 180 GB part):
//
//   generate : one thread per (pixel, sample): camera ray -> ray queue
//   per segment (depth times):
//     trace  : PERSISTENT kernel, one ray per lane; a lane whose ray has ended
//              writes its hit and takes the next ray of its warp's window of the
//              queue, so lanes stay busy through the whole traversal
//     shade  : one thread per ray: hit normal, blend, mirror ray -> next queue
//              (compacted), or final colour -> per-sample slot
//   resolve  : one thread per pixel: samples summed in ascending order -> target
//
// Every per-ray and per-pixel expression is the one in clpt_trace.cuh, and the
// per-pixel sum runs in sample order, so the frame is bit-identical to the
// megakernel's and to the oracle's.
//
// STATUS: an experiment that lost.  On the bench workload the best setting (no
// refill for camera rays, refill at 16 idle lanes for bounce rays) takes 91.3 ms a
// frame against the megakernel's 57.7 ms: refilled lanes restart at the root and
// de-phase the warp (12.6 threads active per instruction against 15.3), the
// persistent loop adds per-iteration bookkeeping, and generate/shade/resolve cost
// ~20 ms of extra passes.  CLSetEngine(2) still selects it; "automatic" does not.
#include "clpt_trace.cuh"

#include <cstdlib>

#ifndef CLPT_WF_MIN_BLOCKS
#define CLPT_WF_MIN_BLOCKS 8
#endif
#ifndef CLPT_WF_WINDOW
#define CLPT_WF_WINDOW 256 // rays a warp claims from the queue per global atomic
#endif

namespace {

struct WfChunk {
    int row0, rows;       // slab rows [row0, row0 + rows) of this rank
    unsigned n_paths;     // rows * width * spp
};

// Queue layout (structure of arrays, 16-byte records so a warp's accesses coalesce):
//   qa[slot] = origin.xyz, str         qb[slot] = dir.xyz, path id bits
//   qc[slot] = col.xyz, -
// path id = (pixel index inside the chunk) * spp + sample.

__device__ __forceinline__ void path_to_pixel(const ClptFrame &F, const WfChunk &C, unsigned pid, int spp, int &x,
                                              int &ly, int &s) {
    const unsigned pix = pid / (unsigned)spp;
    s = (int)(pid - pix * (unsigned)spp);
    ly = C.row0 + (int)(pix / (unsigned)F.width);
    x = (int)(pix % (unsigned)F.width);
}

// Appends one record per participating lane, in lane order, with one atomic per warp.
__device__ __forceinline__ unsigned warp_append(unsigned *counter, bool take) {
    const unsigned mask = __ballot_sync(0xffffffffu, take);
    if (mask == 0) return 0;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(mask) - 1;
    unsigned base = 0;
    if (lane == leader) base = atomicAdd(counter, (unsigned)__popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    return base + (unsigned)__popc(mask & ((1u << lane) - 1u));
}

__global__ void __launch_bounds__(256)
wf_generate(const __grid_constant__ ClptFrame F, const WfChunk C, float4 *__restrict__ qa, float4 *__restrict__ qb,
            float4 *__restrict__ qc, unsigned *__restrict__ q_count) {
    const unsigned pid = blockIdx.x * 256u + threadIdx.x;
    const int spp = F.spp < 1 ? 1 : F.spp;
    bool valid = pid < C.n_paths;
    int x = 0, ly = 0, s = 0, y = 0;
    V3 o = mk(0, 0, 0), d = mk(0, 0, 0);
    if (valid) {
        path_to_pixel(F, C, pid, spp, x, ly, s);
        y = slab_row_to_image_row(F, ly);
        valid = y < F.height && ly < F.local_rows;
    }
    if (valid) primary_ray(F, x, y, (unsigned)(y * F.width + x), F.sample_base + (unsigned)s, o, d);
    const unsigned slot = warp_append(q_count, valid);
    if (valid) {
        qa[slot] = make_float4(o.x, o.y, o.z, 1.0f); // str = 1 (kernel.cl:470)
        qb[slot] = make_float4(d.x, d.y, d.z, __uint_as_float(pid));
        qc[slot] = make_float4(0.0f, 0.0f, 0.0f, 0.0f); // col = 0 (kernel.cl:469)
    }
}

// Persistent traversal.  Per-lane state machine over the steps of clpt_trace.cuh;
// `slot < 0` means the lane is idle.  A warp claims CLPT_WF_WINDOW consecutive
// rays at a time, so its lanes always hold neighbours in the queue (the samples
// of one or two adjacent pixels): coherent where the rays are, busy where not.
template <bool COUNT>
__global__ void __launch_bounds__(256, CLPT_WF_MIN_BLOCKS)
wf_trace(const __grid_constant__ ClptScene S, const float4 *__restrict__ qa, const float4 *__restrict__ qb,
         int2 *__restrict__ hits, const unsigned *__restrict__ q_count, unsigned *__restrict__ head, int max_visits,
         int refill_min, unsigned long long *__restrict__ counters) {
    const int lane = threadIdx.x & 31;
    const unsigned n_rays = *q_count;
    const uint2 *__restrict__ nodes = S.nodes;
    Counters cn = { 0, 0, 0, 0, 0, 0 };

    int slot = -1;
    V3 o = mk(0, 0, 0), d = mk(0, 0, 0), inv = mk(0, 0, 0), p1 = mk(0, 0, 0);
    uint2 n = make_uint2(0u, CLPT_LEAF_WORD);
    int ref = -1, visits = 0;
    float min_hit = 0.0f;
    unsigned w_next = 0, w_end = 0; // this warp's window of the queue (warp-uniform)
    bool exhausted = false;

    for (;;) {
        const unsigned idle = __ballot_sync(0xffffffffu, slot < 0);
        if (idle != 0 && !exhausted && (__popc(idle) >= refill_min || idle == 0xffffffffu)) {
            if (w_next == w_end) {
                unsigned base = 0;
                if (lane == 0) base = atomicAdd(head, (unsigned)CLPT_WF_WINDOW);
                base = __shfl_sync(0xffffffffu, base, 0);
                if (base >= n_rays) {
                    exhausted = true;
                } else {
                    w_next = base;
                    w_end = min(base + (unsigned)CLPT_WF_WINDOW, n_rays);
                }
            }
            if (!exhausted) {
                const unsigned take = min((unsigned)__popc(idle), w_end - w_next);
                const unsigned rank = (unsigned)__popc(idle & ((1u << lane) - 1u));
                if (slot < 0 && rank < take) {
                    slot = (int)(w_next + rank);
                    const float4 a = __ldg(qa + slot), b = __ldg(qb + slot);
                    o = xyz(a);
                    d = xyz(b);
                    inv = mk(frcp(d.x), frcp(d.y), frcp(d.z));
                    if (COUNT) cn.rays++;
                    float tmin, tmax;
                    ref = -1;
                    min_hit = 0.0f;
                    visits = 0;
                    if (root_clip(S, o, inv, tmin, tmax)) {
                        p1 = o;
                        if (tmin > 0.0f) p1 = vadd(p1, vscale(d, tmin));
                        n = __ldg(nodes + (COUNT ? 0 : start_node(S, p1)));
                    } else { // missed the scene box: done before it started
                        hits[slot] = make_int2(-1, 0);
                        slot = -1;
                    }
                }
                w_next += take;
            }
        }
        if (__ballot_sync(0xffffffffu, slot >= 0) == 0) {
            if (exhausted) break;
            continue;
        }
        if (slot >= 0) {
            n = descend<COUNT>(nodes, n, p1, cn);
            if (COUNT) cn.leaves++;
            const float4 *L = S.leaves + 4 * (size_t)n.x;
            const float4 lmin = __ldg(L), lmax = __ldg(L + 1);
            float tmax;
            int far;
            leaf_exit(lmin, lmax, o, inv, tmax, far);
            triangle_run<COUNT>(S.tri, __float_as_int(lmin.w), __float_as_int(lmax.w), o, d, ref, min_hit, cn);
            bool done = ref >= 0 && hit_is_final(ref, leaf_entry(__ldg(L), __ldg(L + 1), o, inv), min_hit);
            if (!done) done = leave_leaf<COUNT>(nodes, L, far, o, d, tmax, max_visits, p1, n, visits, cn);
            if (done) {
                hits[slot] = make_int2(ref, __float_as_int(min_hit));
                slot = -1;
            }
        }
    }
    if (COUNT) {
        unsigned v[6] = { cn.rays, cn.splits, cn.leaves, cn.tris, 0u, cn.capped };
#pragma unroll
        for (int k = 0; k < 6; k++) {
            unsigned sum = v[k];
            for (int off = 16; off > 0; off >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, off);
            if (lane == 0 && sum) atomicAdd(counters + k, (unsigned long long)sum);
        }
    }
}

// One thread per traced ray: kernel.cl:395-421 for that segment.
template <int MODE, bool COUNT>
__global__ void __launch_bounds__(256)
wf_shade(const __grid_constant__ ClptScene S, const __grid_constant__ ClptFrame F, const WfChunk C, int segment,
         int segments, const float4 *__restrict__ qa, const float4 *__restrict__ qb, const float4 *__restrict__ qc,
         const int2 *__restrict__ hits, const unsigned *__restrict__ q_count, float4 *__restrict__ na,
         float4 *__restrict__ nb, float4 *__restrict__ nc_, unsigned *__restrict__ n_count,
         float4 *__restrict__ sample_rgb) {
    const unsigned slot = blockIdx.x * 256u + threadIdx.x;
    const bool live = slot < *q_count;
    bool bounce = false;
    V3 no = mk(0, 0, 0), nd = mk(0, 0, 0), col = mk(0, 0, 0);
    float str = 0.0f;
    unsigned pid = 0;
    Counters cn = { 0, 0, 0, 0, 0, 0 };
    if (live) {
        const float4 a = qa[slot], b = qb[slot], c = qc[slot];
        const int2 hr = hits[slot];
        const V3 o = xyz(a), d = xyz(b);
        col = xyz(c);
        str = a.w;
        pid = __float_as_uint(b.w);
        Hit h;
        h.ref = hr.x;
        h.t = __int_as_float(hr.y);
        const int spp = F.spp < 1 ? 1 : F.spp;
        if (segment == 0 && F.aov_prim != nullptr) {
            int x, ly, s;
            path_to_pixel(F, C, pid, spp, x, ly, s);
            if (s == 0) write_aov<COUNT>(S, F, h, o, d, x, slab_row_to_image_row(F, ly));
        }
        V3 out;
        bool final_colour = true;
        if (h.ref < 0) {
            const float k = fsub(1.0f, str); // :421
            out = mk(fadd(fmul(k, col.x), str), fadd(fmul(k, col.y), str), fadd(fmul(k, col.z), str));
        } else {
            const V3 nrm = hit_normal<COUNT>(S, h, o, d, cn);
            const V3 ncol = mk(fdiv(fadd(nrm.x, 1.0f), 2.0f), fdiv(fadd(nrm.y, 1.0f), 2.0f),
                               fdiv(fadd(nrm.z, 1.0f), 2.0f));
            if (MODE == 0) {
                out = ncol; // the `return` at :396
            } else {
                col = vadd(vscale(col, fsub(1.0f, str)), vscale(ncol, str));
                str = fmul(str, 0.2f);
                if (segment + 1 >= segments) { // depth exhausted: :421 with the blended colour
                    const float k = fsub(1.0f, str);
                    out = mk(fadd(fmul(k, col.x), str), fadd(fmul(k, col.y), str), fadd(fmul(k, col.z), str));
                } else {
                    no = vadd(o, vscale(d, h.t));
                    nd = vnormalize(vsub(d, vscale(nrm, fmul(2.0f, vdot(d, nrm)))));
                    no = vadd(no, vscale(nd, 0.0001f));
                    final_colour = false;
                    bounce = true;
                }
            }
        }
        if (final_colour) sample_rgb[pid] = make_float4(out.x, out.y, out.z, 1.0f);
    }
    const unsigned dst = warp_append(n_count, bounce);
    if (bounce) {
        na[dst] = make_float4(no.x, no.y, no.z, str);
        nb[dst] = make_float4(nd.x, nd.y, nd.z, __uint_as_float(pid));
        nc_[dst] = make_float4(col.x, col.y, col.z, 0.0f);
    }
    if (COUNT) {
        unsigned sum = cn.shade_vn;
        for (int off = 16; off > 0; off >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, off);
        if ((threadIdx.x & 31) == 0 && sum) atomicAdd(F.counters + 4, (unsigned long long)sum);
    }
}

// One thread per pixel of the chunk: ordered sum of its samples.
__global__ void __launch_bounds__(256)
wf_resolve(const __grid_constant__ ClptFrame F, const WfChunk C, const float4 *__restrict__ sample_rgb) {
    const unsigned pix = blockIdx.x * 256u + threadIdx.x;
    const int spp = F.spp < 1 ? 1 : F.spp;
    if (pix >= (unsigned)C.rows * (unsigned)F.width) return;
    const int ly = C.row0 + (int)(pix / (unsigned)F.width), x = (int)(pix % (unsigned)F.width);
    if (ly >= F.local_rows || slab_row_to_image_row(F, ly) >= F.height) return;
    const float4 *src = sample_rgb + (size_t)pix * spp;
    V3 acc = mk(0.0f, 0.0f, 0.0f);
    for (int s = 0; s < spp; s++) {
        const float4 c = src[s];
        acc = vadd(acc, mk(c.x, c.y, c.z));
    }
    store_pixel(F, x, ly, acc, spp);
}

} // namespace

size_t clpt_wavefront_workspace_bytes(size_t max_paths) {
    // 2 queues x 3 float4 arrays, hits, per-sample colours, 8 counters
    return max_paths * (2 * 3 * sizeof(float4) + sizeof(int2) + sizeof(float4)) + 64;
}

int clpt_launch_wavefront(const ClptScene &scene, const ClptFrame &frame, void *workspace, size_t max_paths,
                          int sm_count, cudaStream_t stream) {
    const int spp = frame.spp < 1 ? 1 : frame.spp;
    const int segments = frame.mode == 0 ? (frame.depth > 0 ? 1 : 0) : frame.depth;
    const bool count = (frame.flags & CLPT_F_COUNTERS) != 0;
    char *base = static_cast<char *>(workspace);
    float4 *q[2][3];
    for (int k = 0; k < 2; k++)
        for (int j = 0; j < 3; j++) {
            q[k][j] = reinterpret_cast<float4 *>(base);
            base += max_paths * sizeof(float4);
        }
    int2 *hits = reinterpret_cast<int2 *>(base);
    base += max_paths * sizeof(int2);
    float4 *sample_rgb = reinterpret_cast<float4 *>(base);
    base += max_paths * sizeof(float4);
    unsigned *ctr = reinterpret_cast<unsigned *>(base); // [0],[1] queue counts, [2] trace head

    const size_t per_row = (size_t)frame.width * spp;
    int rows_per_chunk = (int)(max_paths / per_row);
    if (rows_per_chunk < 1) return -1; // workspace too small for one row
    // whole row tiles per chunk keep the slab-row arithmetic simple (not required for correctness)
    int launches = 0;
    int refill_primary = 32, refill_bounce = 8;
    if (const char *e = getenv("CLPT_WF_REFILL0")) refill_primary = atoi(e);
    if (const char *e = getenv("CLPT_WF_REFILL")) refill_bounce = atoi(e);
    for (int row0 = 0; row0 < frame.local_rows; row0 += rows_per_chunk) {
        WfChunk C;
        C.row0 = row0;
        C.rows = min(rows_per_chunk, frame.local_rows - row0);
        C.n_paths = (unsigned)((size_t)C.rows * per_row);
        cudaMemsetAsync(ctr, 0, 16, stream);
        const unsigned blocks = (C.n_paths + 255u) / 256u;
        if (segments == 0) {
            // depth 0 traces nothing: every sample is the miss colour with col = 0, str = 1 -> white
            // (handled by one shade pass over rays marked as misses)
        }
        wf_generate<<<blocks, 256, 0, stream>>>(frame, C, q[0][0], q[0][1], q[0][2], ctr + 0);
        launches++;
        int cur = 0;
        for (int seg = 0; seg < (segments > 0 ? segments : 1); seg++) {
            const int nxt = cur ^ 1;
            cudaMemsetAsync(ctr + nxt, 0, 4, stream);
            cudaMemsetAsync(ctr + 2, 0, 4, stream);
            if (segments > 0) {
                // idle lanes a warp tolerates before it refills: camera rays are coherent and
                // finish together (refill only when the whole warp is done); bounce rays are not
                const int refill = seg == 0 ? refill_primary : refill_bounce;
                if (count)
                    wf_trace<true><<<sm_count * CLPT_WF_MIN_BLOCKS, 256, 0, stream>>>(
                        scene, q[cur][0], q[cur][1], hits, ctr + cur, ctr + 2, frame.max_leaf_visits, refill, frame.counters);
                else
                    wf_trace<false><<<sm_count * CLPT_WF_MIN_BLOCKS, 256, 0, stream>>>(
                        scene, q[cur][0], q[cur][1], hits, ctr + cur, ctr + 2, frame.max_leaf_visits, refill, frame.counters);
                launches++;
            } else {
                cudaMemsetAsync(hits, 0xff, (size_t)C.n_paths * sizeof(int2), stream); // all misses
            }
#define WF_SHADE(MODE, COUNT)                                                                                  \
    wf_shade<MODE, COUNT><<<blocks, 256, 0, stream>>>(scene, frame, C, seg, segments > 0 ? segments : 1, q[cur][0], \
                                                      q[cur][1], q[cur][2], hits, ctr + cur, q[nxt][0], q[nxt][1],  \
                                                      q[nxt][2], ctr + nxt, sample_rgb)
            if (frame.mode == 0) {
                if (count) WF_SHADE(0, true); else WF_SHADE(0, false);
            } else {
                if (count) WF_SHADE(1, true); else WF_SHADE(1, false);
            }
#undef WF_SHADE
            launches++;
            cur = nxt;
        }
        const unsigned pix_blocks = (unsigned)(((size_t)C.rows * frame.width + 255) / 256);
        wf_resolve<<<pix_blocks, 256, 0, stream>>>(frame, C, sample_rgb);
        launches++;
    }
    return launches;
}
