// Kernels of the B200 foam tracer: scene re-layout, forward, backward, benchmark.
// See foam_device.cuh for the spec references.
#pragma once

#include "foam_device.cuh"

namespace rfb {

// ------------------------------------------------------------------ re-layout
// cells[i] = (point, density); sh_rows[i] = SH coefficients padded to 16 bytes.
// Replaces per-step gathers of 3 + 49 scalars by aligned 128-bit loads.
template <typename AttrT>
__global__ void build_cells_kernel(const float *__restrict__ points,
                                   const AttrT *__restrict__ attrs, uint32_t num_points,
                                   int attr_dim_, int sh_row_, float4 *__restrict__ cells,
                                   float *__restrict__ sh_rows) {
    // one warp per point row (grid-stride): lanes stream the row's SH coefficients, lane 0
    // also writes the cell record -- no per-element 64-bit divisions, row-contiguous traffic
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < num_points; i += warps) {
        const AttrT *row = attrs + (uint64_t)i * attr_dim_;
        float *dst = sh_rows + (uint64_t)i * sh_row_;
        for (int s = lane; s < sh_row_; s += 32)
            dst[s] = (s < attr_dim_ - 1) ? (float)row[s] : 0.0f;
        if (lane == 0)
            cells[i] = make_float4(points[3 * (uint64_t)i], points[3 * (uint64_t)i + 1],
                                   points[3 * (uint64_t)i + 2], (float)row[attr_dim_ - 1]);
    }
}

// faces[padded_begin(i) + f] = half4(RN(points[adj[e]] - points[i]), 0),
// nbr[...] = adj[e]  (the reference's prefetch_adjacent_diff_kernel,
// pipeline.cu:546-568, writes the same values in CSR order).  16 lanes per row.
__global__ void build_faces_kernel(const float *__restrict__ points, uint32_t num_points,
                                   const uint32_t *__restrict__ adj,
                                   const uint32_t *__restrict__ off, uint2 *__restrict__ faces,
                                   uint32_t *__restrict__ nbr) {
    uint32_t group = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    uint32_t lane = threadIdx.x & 15;
    uint32_t stride = (gridDim.x * blockDim.x) >> 4;
    for (uint32_t i = group; i < num_points; i += stride) {
        uint32_t a = __ldg(off + i), b = __ldg(off + i + 1);
        uint32_t dst = padded_begin(a, i);
        float px = __ldg(points + 3 * (uint64_t)i), py = __ldg(points + 3 * (uint64_t)i + 1),
              pz = __ldg(points + 3 * (uint64_t)i + 2);
        uint32_t nf = b - a, nf4 = (nf + 3u) & ~3u;
        for (uint32_t f = lane; f < nf4; f += 16) {
            uint2 rec = make_uint2(0u, 0u); // pad: zero face, dp == 0 never wins
            uint32_t j = 0;
            if (f < nf) {
                j = __ldg(adj + a + f);
                float qx = __ldg(points + 3 * (uint64_t)j), qy = __ldg(points + 3 * (uint64_t)j + 1),
                      qz = __ldg(points + 3 * (uint64_t)j + 2);
                __half2 hxy = __floats2half2_rn(__fsub_rn(qx, px), __fsub_rn(qy, py));
                __half2 hzw = __floats2half2_rn(__fsub_rn(qz, pz), 0.0f);
                rec.x = *reinterpret_cast<uint32_t *>(&hxy);
                rec.y = *reinterpret_cast<uint32_t *>(&hzw);
            }
            faces[dst + f] = rec;
            nbr[dst + f] = j;
        }
    }
}

// Same padded layout, but the offsets are COPIED from a caller-built half4[E] array (what
// trace_benchmark is given, pipeline.h:117-126) instead of derived from the points.
__global__ void build_faces_from_diff_kernel(const uint2 *__restrict__ diff, uint32_t num_points,
                                             const uint32_t *__restrict__ adj,
                                             const uint32_t *__restrict__ off, uint2 *__restrict__ faces,
                                             uint32_t *__restrict__ nbr) {
    uint32_t group = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    uint32_t lane = threadIdx.x & 15;
    uint32_t stride = (gridDim.x * blockDim.x) >> 4;
    for (uint32_t i = group; i < num_points; i += stride) {
        uint32_t a = __ldg(off + i), b = __ldg(off + i + 1);
        uint32_t dst = padded_begin(a, i);
        uint32_t nf = b - a, nf4 = (nf + 3u) & ~3u;
        for (uint32_t f = lane; f < nf4; f += 16) {
            uint2 rec = make_uint2(0u, 0u);
            uint32_t j = 0;
            if (f < nf) {
                rec = ldg2(diff + a + f);
                rec.y &= 0x0000FFFFu; // the 4th half is never read by the reference; keep it zero
                j = __ldg(adj + a + f);
            }
            faces[dst + f] = rec;
            nbr[dst + f] = j;
        }
    }
}

// reference-layout adjacent_diff (CSR order), for rfb_prefetch_adjacent_diff
__global__ void adjacent_diff_kernel(const float *__restrict__ points, uint32_t num_points,
                                     const uint32_t *__restrict__ adj,
                                     const uint32_t *__restrict__ off, uint2 *__restrict__ out) {
    uint32_t group = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    uint32_t lane = threadIdx.x & 15;
    uint32_t stride = (gridDim.x * blockDim.x) >> 4;
    for (uint32_t i = group; i < num_points; i += stride) {
        uint32_t a = __ldg(off + i), b = __ldg(off + i + 1);
        float px = __ldg(points + 3 * (uint64_t)i), py = __ldg(points + 3 * (uint64_t)i + 1),
              pz = __ldg(points + 3 * (uint64_t)i + 2);
        for (uint32_t e = a + lane; e < b; e += 16) {
            uint32_t j = __ldg(adj + e);
            float qx = __ldg(points + 3 * (uint64_t)j), qy = __ldg(points + 3 * (uint64_t)j + 1),
                  qz = __ldg(points + 3 * (uint64_t)j + 2);
            __half2 hxy = __floats2half2_rn(__fsub_rn(qx, px), __fsub_rn(qy, py));
            __half2 hzw = __floats2half2_rn(__fsub_rn(qz, pz), 0.0f);
            uint2 rec;
            rec.x = *reinterpret_cast<uint32_t *>(&hxy);
            rec.y = *reinterpret_cast<uint32_t *>(&hzw);
            out[e] = rec;
        }
    }
}

// ------------------------------------------------------------------ forward
struct ForwardParams {
    const float4 *cells;
    const float *sh_rows;
    const float *rays;
    const uint32_t *start;
    const float *quantiles; // [R][Q] or null
    void *rgba;             // [R][4] f32 or f16
    float *qdepth;          // [R][Q]
    uint32_t *qidx;         // [R][Q]
    uint32_t *nint;         // [R] or null
    void *contrib;          // [N] f32 or f16, or null
    uint32_t num_rays;
    uint32_t num_q;
    uint32_t image_width;
    uint32_t blocks_x;
    float weight_threshold;
    uint32_t max_steps;
    int out_half;
};

template <int DEG, typename Faces>
__global__ void __launch_bounds__(kBlock) forward_kernel(const ForwardParams p, const Faces fa) {
    uint32_t r;
    if (!thread_ray(p.num_rays, p.image_width, p.blocks_x, r))
        return;

    RayGeom ray;
    {
        const float *rp = p.rays + 6 * (uint64_t)r;
        ray.ox = __ldg(rp + 0);
        ray.oy = __ldg(rp + 1);
        ray.oz = __ldg(rp + 2);
        ray.dx = __ldg(rp + 3);
        ray.dy = __ldg(rp + 4);
        ray.dz = __ldg(rp + 5);
        normalize_dir(ray.dx, ray.dy, ray.dz);
    }
    float sh[sh_dim(DEG)];
    sh_basis<DEG>(ray.dx, ray.dy, ray.dz, sh);

    const uint32_t Q = p.quantiles ? p.num_q : 0u;
    const float *qv = p.quantiles + (uint64_t)r * p.num_q;
    uint32_t qi = 0;
    float cq = Q ? __ldg(qv) : 0.0f;

    float T = 1.0f;
    float cr = 0.0f, cg = 0.0f, cb = 0.0f;

    auto cell_fn = [&](uint32_t cell, const float4 &pc, float t0, float t1, const float4 &) -> bool {
        float s = pc.w;
        float r_ = 0.0f, g_ = 0.0f, b_ = 0.0f;
        if (s > 1e-6f)
            sh_to_rgb<DEG>(p.sh_rows + (uint64_t)cell * sh_row(DEG), sh, r_, g_, b_);
        float delta = fmaxf(__fsub_rn(t1, t0), 0.0f);
        float alpha = 1.0f - expf(-s * delta);
        float w = __fmul_rn(T, alpha);
        if (p.contrib) {
            if (p.out_half)
                atomicAdd(reinterpret_cast<__half *>(p.contrib) + cell, __float2half_rn(w));
            else
                atomicAdd(reinterpret_cast<float *>(p.contrib) + cell, w);
        }
        cr = __fmaf_rn(w, r_, cr);
        cg = __fmaf_rn(w, g_, cg);
        cb = __fmaf_rn(w, b_, cb);
        float Tn = __fmul_rn(T, __fsub_rn(1.0f, alpha));
        while (qi < Q && Tn < cq) {
            p.qdepth[(uint64_t)r * Q + qi] = __fadd_rn(t0, __fdiv_rn(logf(__fdiv_rn(T, cq)), s));
            p.qidx[(uint64_t)r * Q + qi] = cell;
            qi++;
            if (qi < Q)
                cq = __ldg(qv + qi);
        }
        T = Tn;
        return T > p.weight_threshold;
    };

    uint32_t n = walk(fa, p.cells, ray, __ldg(p.start + r), p.max_steps, cell_fn);

    while (qi < Q) {
        p.qdepth[(uint64_t)r * Q + qi] = -1.0f;
        p.qidx[(uint64_t)r * Q + qi] = kNone;
        qi++;
    }
    float a = __fsub_rn(1.0f, T);
    if (p.out_half) {
        __half2 lo = __floats2half2_rn(cr, cg), hi = __floats2half2_rn(cb, a);
        uint2 v;
        v.x = *reinterpret_cast<uint32_t *>(&lo);
        v.y = *reinterpret_cast<uint32_t *>(&hi);
        reinterpret_cast<uint2 *>(p.rgba)[r] = v;
    } else {
        reinterpret_cast<float4 *>(p.rgba)[r] = make_float4(cr, cg, cb, a);
    }
    if (p.nint)
        p.nint[r] = n;
}

// ------------------------------------------------------------------ the walk tape
// A training step traces every ray twice: forward, then the backward re-walk
// (pipeline.cu:132-343 repeats trace<>()).  The face scan is ~2/3 of the walk's instructions,
// so when a backward pass is expected the forward records, per ray and visited cell,
// (cell, t1) -- 8 bytes -- and the backward replays that instead of re-scanning faces.  The
// record stream is identical to what the re-walk would compute (same kernel arithmetic), so
// results do not change.  Layout: all lanes of a warp are at the same step index k in the same
// loop iteration, so warp w writes rows of 32 records; rows are grouped in chunks of kTapeChunk
// steps (8 KB) handed out by a bump allocator, chunk ids in table[w][k / kTapeChunk].  If the
// pool runs out the overflow flag is raised and the backward re-walks (still correct); the host
// grows the pool for the next step.
constexpr int kTapeChunk = 32;     // steps per chunk
constexpr uint32_t kTapeNoChunk = 0xFFFFFFFFu;

struct Tape {
    uint2 *pool;        // [capacity][kTapeChunk][32] records (cell, float_as_uint(t1))
    uint32_t *table;    // [num_warps][table_stride] chunk ids
    uint2 *per_ray;     // [R] (number of records, cell entered after the last record)
    uint32_t *ctrl;     // [0] bump cursor, [1] overflow flag
    uint32_t capacity;  // chunks in the pool
    uint32_t table_stride;
};

// Forward with an explicit warp-synchronous loop (all lanes stay in the loop until the warp is
// done, so lane 0 can allocate tape chunks for the warp) that records the tape.
template <int DEG, typename Faces>
__global__ void __launch_bounds__(kBlock) forward_record_kernel(const ForwardParams p, const Faces fa,
                                                                const Tape tape) {
    constexpr unsigned FULL = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t gwarp = blockIdx.x * (kBlock / 32) + (threadIdx.x >> 5);
    uint32_t r;
    bool done = !thread_ray(p.num_rays, p.image_width, p.blocks_x, r);
    const bool has_ray = !done;

    RayGeom ray = {0.f, 0.f, 0.f, 0.f, 0.f, 1.f};
    float sh[sh_dim(DEG)];
    uint32_t Q = 0, qi = 0;
    const float *qv = nullptr;
    float cq = 0.0f;
    uint32_t cur = 0;
    float4 pc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (has_ray) {
        const float *rp = p.rays + 6 * (uint64_t)r;
        ray.ox = __ldg(rp + 0);
        ray.oy = __ldg(rp + 1);
        ray.oz = __ldg(rp + 2);
        ray.dx = __ldg(rp + 3);
        ray.dy = __ldg(rp + 4);
        ray.dz = __ldg(rp + 5);
        normalize_dir(ray.dx, ray.dy, ray.dz);
        Q = p.quantiles ? p.num_q : 0u;
        qv = p.quantiles + (uint64_t)r * p.num_q;
        cq = Q ? __ldg(qv) : 0.0f;
        cur = __ldg(p.start + r);
        pc = ldg4(p.cells + cur);
    }
    sh_basis<DEG>(ray.dx, ray.dy, ray.dz, sh);

    float T = 1.0f, cr = 0.0f, cg = 0.0f, cb = 0.0f, t0 = 0.0f;
    uint32_t n = 0, nrec = 0;
    uint32_t chunk = kTapeNoChunk;
    for (uint32_t k = 0;; ++k) {
        if ((k % kTapeChunk) == 0) { // the warp enters a new chunk of steps
            uint32_t c = kTapeNoChunk;
            if (lane == 0) {
                c = atomicAdd(tape.ctrl, 1u);
                if (c >= tape.capacity || k / kTapeChunk >= tape.table_stride) {
                    atomicExch(tape.ctrl + 1, 1u);
                    c = kTapeNoChunk;
                } else {
                    tape.table[(uint64_t)gwarp * tape.table_stride + k / kTapeChunk] = c;
                }
            }
            chunk = __shfl_sync(FULL, c, 0);
        }
        if (!done) {
            n++;
            if (n > p.max_steps) {
                done = true;
            } else {
                uint32_t begin, nf;
                fa.row(cur, begin, nf);
                float t1 = __int_as_float(0x7f800000);
                uint32_t face = kNone;
                fa.scan(begin, nf, pc.x, pc.y, pc.z, ray, t1, face);
                if (face == kNone) {
                    done = true;
                } else {
                    if (chunk != kTapeNoChunk) // streaming store: the tape is written once, read once
                        __stcs(tape.pool + ((uint64_t)chunk * kTapeChunk + (k % kTapeChunk)) * 32 + lane,
                               make_uint2(cur, __float_as_uint(t1)));
                    nrec++;
                    uint32_t nxt = fa.neighbour(begin, face);
                    float4 pn = ldg4(p.cells + nxt);
                    if (t1 > t0) {
                        float s = pc.w;
                        float r_ = 0.0f, g_ = 0.0f, b_ = 0.0f;
                        if (s > 1e-6f)
                            sh_to_rgb<DEG>(p.sh_rows + (uint64_t)cur * sh_row(DEG), sh, r_, g_, b_);
                        float delta = fmaxf(__fsub_rn(t1, t0), 0.0f);
                        float alpha = 1.0f - expf(-s * delta);
                        float w = __fmul_rn(T, alpha);
                        if (p.contrib) {
                            if (p.out_half)
                                atomicAdd(reinterpret_cast<__half *>(p.contrib) + cur, __float2half_rn(w));
                            else
                                atomicAdd(reinterpret_cast<float *>(p.contrib) + cur, w);
                        }
                        cr = __fmaf_rn(w, r_, cr);
                        cg = __fmaf_rn(w, g_, cg);
                        cb = __fmaf_rn(w, b_, cb);
                        float Tn = __fmul_rn(T, __fsub_rn(1.0f, alpha));
                        while (qi < Q && Tn < cq) {
                            p.qdepth[(uint64_t)r * Q + qi] = __fadd_rn(t0, __fdiv_rn(logf(__fdiv_rn(T, cq)), s));
                            p.qidx[(uint64_t)r * Q + qi] = cur;
                            qi++;
                            if (qi < Q)
                                cq = __ldg(qv + qi);
                        }
                        T = Tn;
                        done = !(T > p.weight_threshold);
                    }
                    t0 = fmaxf(t0, t1);
                    cur = nxt;
                    pc = pn;
                }
            }
        }
        if (!__any_sync(FULL, !done))
            break;
    }
    if (!has_ray)
        return;
    tape.per_ray[r] = make_uint2(nrec, cur);
    while (qi < Q) {
        p.qdepth[(uint64_t)r * Q + qi] = -1.0f;
        p.qidx[(uint64_t)r * Q + qi] = kNone;
        qi++;
    }
    float a = __fsub_rn(1.0f, T);
    if (p.out_half) {
        __half2 lo = __floats2half2_rn(cr, cg), hi = __floats2half2_rn(cb, a);
        uint2 v;
        v.x = *reinterpret_cast<uint32_t *>(&lo);
        v.y = *reinterpret_cast<uint32_t *>(&hi);
        reinterpret_cast<uint2 *>(p.rgba)[r] = v;
    } else {
        reinterpret_cast<float4 *>(p.rgba)[r] = make_float4(cr, cg, cb, a);
    }
    if (p.nint)
        p.nint[r] = n;
}

// EXPERIMENTS (RFB_FWD_VARIANT=1 / 2; not the default, not yet measured on a B200): forward_record_kernel with
//   SCAN == 1  the face scan run warp-synchronously so that 4-face chunks without a front face for any lane are skipped
//              by a vote (PaddedFaces::scan_voted);
//   SCAN == 2  the two-pass scan (PaddedFaces::scan_two_pass): dp of all faces first, ranking of the front faces only.
// Everything else is a copy of the kernel above; results are identical (tests/test_emu_kernels.py).
template <int DEG, typename Faces, int SCAN>
__global__ void __launch_bounds__(kBlock) forward_record_voted_kernel(const ForwardParams p, const Faces fa,
                                                                const Tape tape) {
    constexpr unsigned FULL = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t gwarp = blockIdx.x * (kBlock / 32) + (threadIdx.x >> 5);
    uint32_t r;
    bool done = !thread_ray(p.num_rays, p.image_width, p.blocks_x, r);
    const bool has_ray = !done;

    RayGeom ray = {0.f, 0.f, 0.f, 0.f, 0.f, 1.f};
    float sh[sh_dim(DEG)];
    uint32_t Q = 0, qi = 0;
    const float *qv = nullptr;
    float cq = 0.0f;
    uint32_t cur = 0;
    float4 pc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (has_ray) {
        const float *rp = p.rays + 6 * (uint64_t)r;
        ray.ox = __ldg(rp + 0);
        ray.oy = __ldg(rp + 1);
        ray.oz = __ldg(rp + 2);
        ray.dx = __ldg(rp + 3);
        ray.dy = __ldg(rp + 4);
        ray.dz = __ldg(rp + 5);
        normalize_dir(ray.dx, ray.dy, ray.dz);
        Q = p.quantiles ? p.num_q : 0u;
        qv = p.quantiles + (uint64_t)r * p.num_q;
        cq = Q ? __ldg(qv) : 0.0f;
        cur = __ldg(p.start + r);
        pc = ldg4(p.cells + cur);
    }
    sh_basis<DEG>(ray.dx, ray.dy, ray.dz, sh);

    float T = 1.0f, cr = 0.0f, cg = 0.0f, cb = 0.0f, t0 = 0.0f;
    uint32_t n = 0, nrec = 0;
    uint32_t chunk = kTapeNoChunk;
    for (uint32_t k = 0;; ++k) {
        if ((k % kTapeChunk) == 0) { // the warp enters a new chunk of steps
            uint32_t c = kTapeNoChunk;
            if (lane == 0) {
                c = atomicAdd(tape.ctrl, 1u);
                if (c >= tape.capacity || k / kTapeChunk >= tape.table_stride) {
                    atomicExch(tape.ctrl + 1, 1u);
                    c = kTapeNoChunk;
                } else {
                    tape.table[(uint64_t)gwarp * tape.table_stride + k / kTapeChunk] = c;
                }
            }
            chunk = __shfl_sync(FULL, c, 0);
        }
        // the face scan, hoisted out of the divergent region and run by the whole warp
        bool stepping = !done;
        uint32_t v_begin = 0, v_nf = 0, v_face = kNone;
        float v_t1 = __int_as_float(0x7f800000);
        if (stepping && n + 1 > p.max_steps)
            stepping = false; // the budget check below ends the ray
        if (stepping)
            fa.row(cur, v_begin, v_nf);
        if (SCAN == 1)
            fa.scan_voted(stepping, v_begin, v_nf, pc.x, pc.y, pc.z, ray, v_t1, v_face);
        else if (stepping)
            fa.scan_two_pass(v_begin, v_nf, pc.x, pc.y, pc.z, ray, v_t1, v_face);
        if (!done) {
            n++;
            if (n > p.max_steps) {
                done = true;
            } else {
                const uint32_t begin = v_begin, face = v_face;
                const float t1 = v_t1;
                if (face == kNone) {
                    done = true;
                } else {
                    if (chunk != kTapeNoChunk) // streaming store: the tape is written once, read once
                        __stcs(tape.pool + ((uint64_t)chunk * kTapeChunk + (k % kTapeChunk)) * 32 + lane,
                               make_uint2(cur, __float_as_uint(t1)));
                    nrec++;
                    uint32_t nxt = fa.neighbour(begin, face);
                    float4 pn = ldg4(p.cells + nxt);
                    if (t1 > t0) {
                        float s = pc.w;
                        float r_ = 0.0f, g_ = 0.0f, b_ = 0.0f;
                        if (s > 1e-6f)
                            sh_to_rgb<DEG>(p.sh_rows + (uint64_t)cur * sh_row(DEG), sh, r_, g_, b_);
                        float delta = fmaxf(__fsub_rn(t1, t0), 0.0f);
                        float alpha = 1.0f - expf(-s * delta);
                        float w = __fmul_rn(T, alpha);
                        if (p.contrib) {
                            if (p.out_half)
                                atomicAdd(reinterpret_cast<__half *>(p.contrib) + cur, __float2half_rn(w));
                            else
                                atomicAdd(reinterpret_cast<float *>(p.contrib) + cur, w);
                        }
                        cr = __fmaf_rn(w, r_, cr);
                        cg = __fmaf_rn(w, g_, cg);
                        cb = __fmaf_rn(w, b_, cb);
                        float Tn = __fmul_rn(T, __fsub_rn(1.0f, alpha));
                        while (qi < Q && Tn < cq) {
                            p.qdepth[(uint64_t)r * Q + qi] = __fadd_rn(t0, __fdiv_rn(logf(__fdiv_rn(T, cq)), s));
                            p.qidx[(uint64_t)r * Q + qi] = cur;
                            qi++;
                            if (qi < Q)
                                cq = __ldg(qv + qi);
                        }
                        T = Tn;
                        done = !(T > p.weight_threshold);
                    }
                    t0 = fmaxf(t0, t1);
                    cur = nxt;
                    pc = pn;
                }
            }
        }
        if (!__any_sync(FULL, !done))
            break;
    }
    if (!has_ray)
        return;
    tape.per_ray[r] = make_uint2(nrec, cur);
    while (qi < Q) {
        p.qdepth[(uint64_t)r * Q + qi] = -1.0f;
        p.qidx[(uint64_t)r * Q + qi] = kNone;
        qi++;
    }
    float a = __fsub_rn(1.0f, T);
    if (p.out_half) {
        __half2 lo = __floats2half2_rn(cr, cg), hi = __floats2half2_rn(cb, a);
        uint2 v;
        v.x = *reinterpret_cast<uint32_t *>(&lo);
        v.y = *reinterpret_cast<uint32_t *>(&hi);
        reinterpret_cast<uint2 *>(p.rgba)[r] = v;
    } else {
        reinterpret_cast<float4 *>(p.rgba)[r] = make_float4(cr, cg, cb, a);
    }
    if (p.nint)
        p.nint[r] = n;
}

// ------------------------------------------------------------------ backward
struct BackwardParams {
    const float4 *cells;
    const float *sh_rows;
    const float *rays;
    const uint32_t *start;
    const float *quantiles;
    const uint32_t *qidx;
    const void *rgba;      // saved forward output [R][4]
    const void *rgba_grad; // [R][4]
    const float *depth_grad;
    const void *ray_error; // [R] or null
    void *point_error;     // [N] or null
    float *acc;            // [N][grad_row] fp32 accumulator
    uint32_t num_rays;
    uint32_t num_q;
    uint32_t image_width;
    uint32_t blocks_x;
    float weight_threshold;
    uint32_t max_steps;
    int io_half;
};

// Loads this thread's ray and upstream gradients (pipeline.cu:156-207).
template <int DEG>
__device__ __forceinline__ void backward_ray_setup(const BackwardParams &p, uint32_t r, RayGeom &ray,
                                                   float *sh, BackwardRay &st) {
    const float *rp = p.rays + 6 * (uint64_t)r;
    ray.ox = __ldg(rp + 0);
    ray.oy = __ldg(rp + 1);
    ray.oz = __ldg(rp + 2);
    ray.dx = __ldg(rp + 3);
    ray.dy = __ldg(rp + 4);
    ray.dz = __ldg(rp + 5);
    normalize_dir(ray.dx, ray.dy, ray.dz);
    sh_basis<DEG>(ray.dx, ray.dy, ray.dz, sh);
    st.err = 0.0f;
    if (p.io_half) {
        const __half *o = reinterpret_cast<const __half *>(p.rgba) + 4 * (uint64_t)r;
        const __half *gg = reinterpret_cast<const __half *>(p.rgba_grad) + 4 * (uint64_t)r;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            st.out[c] = __half2float(o[c]);
            st.g[c] = __half2float(gg[c]);
        }
        if (p.ray_error)
            st.err = __half2float(reinterpret_cast<const __half *>(p.ray_error)[r]);
    } else {
        float4 o = __ldg(reinterpret_cast<const float4 *>(p.rgba) + r);
        float4 gg = __ldg(reinterpret_cast<const float4 *>(p.rgba_grad) + r);
        st.out[0] = o.x; st.out[1] = o.y; st.out[2] = o.z; st.out[3] = o.w;
        st.g[0] = gg.x; st.g[1] = gg.y; st.g[2] = gg.z; st.g[3] = gg.w;
        if (p.ray_error)
            st.err = __ldg(reinterpret_cast<const float *>(p.ray_error) + r);
    }
    st.init();
    st.Q = p.quantiles ? p.num_q : 0u;
    st.qv = p.quantiles + (uint64_t)r * p.num_q;
    st.dg = p.depth_grad + (uint64_t)r * p.num_q;
    st.cq = st.Q ? __ldg(st.qv) : 0.0f;
    for (uint32_t i = 0; i < st.Q; ++i) { // pipeline.cu:196-207
        uint32_t pi = __ldg(p.qidx + (uint64_t)r * st.Q + i);
        if (pi != kNone)
            st.cdg = __fadd_rn(st.cdg, __fdiv_rn(__ldg(st.dg + i), ldg4(p.cells + pi).w));
    }
}

__device__ __forceinline__ void add_point_error(const BackwardParams &p, uint32_t cell, float v) {
    if (p.io_half)
        atomicAdd(reinterpret_cast<__half *>(p.point_error) + cell, __float2half_rn(v));
    else
        atomicAdd(reinterpret_cast<float *>(p.point_error) + cell, v);
}

// one gradient row as 128-bit reductions (write_rgb_grad_to_sh, sh_utils.cuh:85-92, plus the
// density slot); a row whose three channel gradients are all zero adds nothing and is skipped
template <int DEG>
__device__ __forceinline__ void reduce_row_direct(float *row, const float *sh, const float dL_drgb[3],
                                                  float dL_ds) {
    constexpr int SR = sh_row(DEG);
    if (dL_drgb[0] != 0.0f || dL_drgb[1] != 0.0f || dL_drgb[2] != 0.0f) {
#pragma unroll
        for (int i = 0; i < SR; i += 4) {
            float v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
                v[k] = (i + k < 3 * sh_dim(DEG)) ? __fmul_rn(sh[(i + k) / 3], dL_drgb[(i + k) % 3]) : 0.0f;
            red_add_v4(row + i, v[0], v[1], v[2], v[3]);
        }
    }
    red_add_v4(row + SR, dL_ds, 0.0f, 0.0f, 0.0f);
}

// ---- backward, direct: every lane reduces its own rows straight to HBM
template <int DEG, typename Faces>
__global__ void __launch_bounds__(kBlock) backward_kernel(const BackwardParams p, const Faces fa) {
    uint32_t r;
    if (!thread_ray(p.num_rays, p.image_width, p.blocks_x, r))
        return;
    constexpr int GR = grad_row(DEG);
    constexpr int SR = sh_row(DEG);
    RayGeom ray;
    float sh[sh_dim(DEG)];
    BackwardRay st;
    backward_ray_setup<DEG>(p, r, ray, sh, st);

    auto cell_fn = [&](uint32_t cell, const float4 &pc, float t0, float t1, const float4 &pn) -> bool {
        float rgb[3] = {0.0f, 0.0f, 0.0f};
        if (pc.w > 1e-6f)
            sh_to_rgb<DEG>(p.sh_rows + (uint64_t)cell * SR, sh, rgb[0], rgb[1], rgb[2]);
        float dL_drgb[3], dL_ds, w, fx, fy, fz;
        bool flush;
        uint32_t flush_idx;
        bool go = st.cell(cell, pc, pn, t0, t1, rgb, ray, p.weight_threshold, dL_drgb, dL_ds, w, flush,
                          flush_idx, fx, fy, fz);
        if (p.point_error)
            add_point_error(p, cell, __fmul_rn(w, st.err));
        if (flush)
            red_add_v4(p.acc + (uint64_t)flush_idx * GR + SR, 0.0f, fx, fy, fz);
        reduce_row_direct<DEG>(p.acc + (uint64_t)cell * GR, sh, dL_drgb, dL_ds);
        return go;
    };
    walk(fa, p.cells, ray, __ldg(p.start + r), p.max_steps, cell_fn);
}

// ---- backward, warp-aggregated
// Same mathematics; what changes is how the per-cell gradient rows reach HBM.  Neighbouring
// rays of an 8x4 tile cross mostly the same cells, a few iterations apart (one ray clips a
// sliver cell the other misses and they drift out of lock-step), so instead of 13 x 128-bit
// reductions per lane per step the warp keeps a small direct-mapped cache of gradient rows in
// shared memory:
//   * lanes are grouped by cell (MATCH.ANY); a lane that is alone in its cell issues its
//     reductions directly, in parallel with the other singleton lanes;
//   * every lane of a multi-lane group writes its row (SH products + density grad) to a
//     staging row in shared memory; the warp sums the staged rows of a group "transposed" --
//     lane j owns elements 2j, 2j+1 of the row -- so the shared-memory update needs no atomics
//     and has no bank conflicts, and adds the sum to the cell's cache row;
//   * a cache row is written to HBM (13 lanes x one RED.128) only when its slot is claimed by
//     another cell, or when the warp's rays are finished.
// Position gradients (3 floats to the previous composited cell) stay direct reductions.
// MIN_GROUP: smallest same-cell lane group routed through the cache (smaller groups reduce
// directly: all such lanes issue their 13 reductions simultaneously, which costs fewer issue
// slots than one serial cache round per group, at the price of more L2 atomic traffic).
// REPLAY = true: the steps come from the forward's tape of (cell, t1) instead of re-scanning
// faces.  When a tape is offered the host launches both instantiations back to back; each looks
// at the tape's overflow flag first and exactly one of them does the work (the other returns at
// once), so a pool overflow needs no host round trip.  (Prefetching the next cell's SH row into
// L1 or L2 as soon as the cell is known was tried and is no faster: 11.8 vs 11.65 ms.)
template <int DEG, typename Faces, int SLOTS, int MIN_GROUP, int MIN_BLOCKS, bool REPLAY>
__global__ void __launch_bounds__(kBlock, MIN_BLOCKS)
    backward_cached_kernel(const BackwardParams p, const Faces fa, const Tape tape) {
    if (tape.pool != nullptr) {
        const bool overflowed = tape.ctrl[1] != 0u;
        if (REPLAY == overflowed) // replay kernel on an overflowed tape / re-walk kernel on a good one
            return;
    }
    constexpr int GR = grad_row(DEG);
    constexpr int SR = sh_row(DEG);
    constexpr int HALF_ROW = GR / 2; // lanes that own two row elements each
    constexpr unsigned FULL = 0xffffffffu;
    static_assert(HALF_ROW <= 32, "row too wide for one warp pass");
    static_assert((SLOTS & (SLOTS - 1)) == 0, "SLOTS must be a power of two");

#ifdef RFB_EMU
    float *smem = rfb_emu_dynamic_smem();
#else
    extern __shared__ __align__(16) float smem[];
#endif
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int WARP_FLOATS = (32 * GR + SLOTS * GR + SLOTS + 3) & ~3; // keeps every warp's rows 16-byte aligned
    float *stage = smem + warp * WARP_FLOATS;                          // [32][GR]
    float *cache = stage + 32 * GR;                                    // [SLOTS][GR]
    uint32_t *tags = reinterpret_cast<uint32_t *>(cache + SLOTS * GR); // [SLOTS]
    for (int i = lane; i < SLOTS; i += 32)
        tags[i] = kNone;
    __syncwarp();

    uint32_t r;
    bool done = !thread_ray(p.num_rays, p.image_width, p.blocks_x, r);
    RayGeom ray = {0.f, 0.f, 0.f, 0.f, 0.f, 1.f};
    float sh[sh_dim(DEG)];
    BackwardRay st;
    uint32_t cur = 0;
    float4 pc = make_float4(0.f, 0.f, 0.f, 0.f);
    constexpr bool replay = REPLAY;
    const uint32_t gwarp = blockIdx.x * (kBlock / 32) + warp;
    uint32_t nrec = 0, last_cell = 0;
    uint2 rec = make_uint2(0u, 0u); // record of the step about to be processed (replay)
    if (!done) {
        backward_ray_setup<DEG>(p, r, ray, sh, st);
        cur = __ldg(p.start + r);
        pc = ldg4(p.cells + cur);
        if (replay) {
            uint2 pr = tape.per_ray[r];
            nrec = pr.x;
            last_cell = pr.y;
            done = nrec == 0;
        }
    } else {
#pragma unroll
        for (int i = 0; i < sh_dim(DEG); ++i)
            sh[i] = 0.0f;
    }
    float t0 = 0.0f;
    uint32_t n = 0;
    // replay pipeline: records k and k+1 are kept in registers and record k+2 is requested
    // while step k is processed, so the streamed tape read is never waited for.  (Also
    // prefetching the cell gather one step ahead costs 8 more registers, spills, and is slower:
    // 12.5 vs 11.6 ms.)
    uint32_t chunk_ahead = 0; // chunk id (warp-uniform) of tape row k + 2
    uint2 recB = make_uint2(0u, 0u);
    auto tape_row = [&](uint32_t chunk_id, uint32_t j) -> uint2 {
        // record j of this lane; past the end: the cell the forward stopped in
        return j < nrec ? __ldcs(tape.pool + ((uint64_t)chunk_id * kTapeChunk + (j % kTapeChunk)) * 32 + lane)
                        : make_uint2(last_cell, 0u);
    };
    if (replay) {
        chunk_ahead = tape.table[(uint64_t)gwarp * tape.table_stride]; // rows 0..31
        if (!done) {
            rec = tape_row(chunk_ahead, 0);
            recB = tape_row(chunk_ahead, 1);
        }
    }

    for (uint32_t k = 0;; ++k) {
        bool c_valid = false; // this lane has a (SH, density) row for cell c_cell this iteration
        uint32_t c_cell = kNone;
        float dL_ds = 0.0f;
        float dL_drgb[3] = {0.0f, 0.0f, 0.0f};

        // one step of the ray: (cur, t1, nxt) from the tape, or from scanning the cell's faces
        bool step = false;
        float t1 = __int_as_float(0x7f800000);
        uint32_t nxt = 0;
        if (replay) {
            if (((k + 2) % kTapeChunk) == 0)
                chunk_ahead = tape.table[(uint64_t)gwarp * tape.table_stride + (k + 2) / kTapeChunk];
            if (!done) {
                uint2 recC = tape_row(chunk_ahead, k + 2);
                t1 = __uint_as_float(rec.y);
                nxt = recB.x;
                rec = recB;
                recB = recC;
                step = true;
            }
        } else if (!done) {
            n++;
            if (n > p.max_steps) {
                done = true;
            } else {
                uint32_t begin, nf;
                fa.row(cur, begin, nf);
                uint32_t face = kNone;
                fa.scan(begin, nf, pc.x, pc.y, pc.z, ray, t1, face);
                if (face == kNone) {
                    done = true;
                } else {
                    nxt = fa.neighbour(begin, face);
                    step = true;
                }
            }
        }
        if (step) {
            float4 pn = ldg4(p.cells + nxt);
            if (t1 > t0) {
                float rgb[3] = {0.0f, 0.0f, 0.0f};
                if (pc.w > 1e-6f)
                    sh_to_rgb<DEG>(p.sh_rows + (uint64_t)cur * SR, sh, rgb[0], rgb[1], rgb[2]);
                float w, fx, fy, fz;
                bool flush;
                uint32_t flush_idx;
                bool go = st.cell(cur, pc, pn, t0, t1, rgb, ray, p.weight_threshold, dL_drgb, dL_ds, w,
                                  flush, flush_idx, fx, fy, fz);
                if (p.point_error)
                    add_point_error(p, cur, __fmul_rn(w, st.err));
                if (flush)
                    red_add_v4(p.acc + (uint64_t)flush_idx * GR + SR, 0.0f, fx, fy, fz);
                c_valid = true;
                c_cell = cur;
                done = !go;
            }
            t0 = fmaxf(t0, t1);
            cur = nxt;
            pc = pn;
            if (replay && k + 1 >= nrec)
                done = true;
        }

        // ---- warp-collective phase: route this iteration's rows
        unsigned grp = __match_any_sync(FULL, c_valid ? c_cell : (0x80000000u | lane));
        bool single = c_valid && __popc(grp) < MIN_GROUP;
        bool staged = c_valid && !single;
        if (single)
            reduce_row_direct<DEG>(p.acc + (uint64_t)c_cell * GR, sh, dL_drgb, dL_ds);
        unsigned todo = __ballot_sync(FULL, staged);
        if (todo) {
            if (staged) {
                float4 *srow = reinterpret_cast<float4 *>(stage + lane * GR);
#pragma unroll
                for (int i = 0; i < SR; i += 4) {
                    float v[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        v[k] = (i + k < 3 * sh_dim(DEG)) ? __fmul_rn(sh[(i + k) / 3], dL_drgb[(i + k) % 3]) : 0.0f;
                    srow[i / 4] = make_float4(v[0], v[1], v[2], v[3]);
                }
                srow[SR / 4] = make_float4(dL_ds, 0.0f, 0.0f, 0.0f);
            }
            __syncwarp();
            while (todo) {
                int leader = __ffs(todo) - 1;
                unsigned gmask = __shfl_sync(FULL, grp, leader);
                uint32_t lc = __shfl_sync(FULL, c_cell, leader);
                todo &= ~gmask;
                float a0 = 0.0f, a1 = 0.0f;
                if (lane < HALF_ROW) {
                    // four staged rows per round: the loads are issued together so their
                    // shared-memory latency overlaps (groups have >= MIN_GROUP members)
                    const float *col = stage + 2 * lane;
                    unsigned m = gmask;
                    while (m) {
                        int s0 = __ffs(m) - 1;
                        m &= m - 1;
                        int s1 = m ? __ffs(m) - 1 : -1;
                        m &= m - 1;
                        int s2 = m ? __ffs(m) - 1 : -1;
                        m &= m - 1;
                        int s3 = m ? __ffs(m) - 1 : -1;
                        m &= m - 1;
                        float2 v0 = *reinterpret_cast<const float2 *>(col + s0 * GR);
                        float2 v1 = s1 >= 0 ? *reinterpret_cast<const float2 *>(col + s1 * GR) : make_float2(0.f, 0.f);
                        float2 v2 = s2 >= 0 ? *reinterpret_cast<const float2 *>(col + s2 * GR) : make_float2(0.f, 0.f);
                        float2 v3 = s3 >= 0 ? *reinterpret_cast<const float2 *>(col + s3 * GR) : make_float2(0.f, 0.f);
                        a0 += (v0.x + v1.x) + (v2.x + v3.x);
                        a1 += (v0.y + v1.y) + (v2.y + v3.y);
                    }
                }
                uint32_t slot = (lc * 2654435761u) >> (32 - __builtin_ctz(SLOTS));
                uint32_t tag = tags[slot];
                __syncwarp();
                float2 *crow = reinterpret_cast<float2 *>(cache + slot * GR + 2 * lane);
                if (tag == lc) {
                    if (lane < HALF_ROW) {
                        float2 c = *crow;
                        c.x += a0;
                        c.y += a1;
                        *crow = c;
                    }
                } else {
                    if (tag != kNone && lane < GR / 4) {
                        float4 v = *reinterpret_cast<const float4 *>(cache + slot * GR + 4 * lane);
                        red_add_v4(p.acc + (uint64_t)tag * GR + 4 * lane, v.x, v.y, v.z, v.w);
                    }
                    __syncwarp();
                    if (lane < HALF_ROW)
                        *crow = make_float2(a0, a1);
                    if (lane == 0)
                        tags[slot] = lc;
                }
                __syncwarp();
            }
        }
        if (!__any_sync(FULL, !done))
            break;
    }

    // drain the cache
    __syncwarp();
    for (int slot = 0; slot < SLOTS; ++slot) {
        uint32_t tag = tags[slot];
        if (tag != kNone && lane < GR / 4) {
            float4 v = *reinterpret_cast<const float4 *>(cache + slot * GR + 4 * lane);
            red_add_v4(p.acc + (uint64_t)tag * GR + 4 * lane, v.x, v.y, v.z, v.w);
        }
    }
}

// ---- backward, pooled rows (EXPERIMENT, not the default: RFB_BWD_VARIANT=4..6; functionally checked on the CPU
// emulator, tests/test_emu_kernels.py, not yet measured on a B200)
// tests/tools/tape_stats.py on a scene with the bench's rays/points ratio: 40 % of the composited lane-steps sit in
// same-cell groups of fewer than 6 lanes, which the kernel above reduces directly and which produce 78 % of its
// 16-byte reductions; another 15 % are the per-step position-gradient reductions.  A 16-row cache fed by EVERY group
// would issue 0.29x the reductions (32 rows: 0.22x) -- but only if a group costs far less than one serial warp round.
// This kernel changes three things:
//  * a lane's record is compact: (dL/drgb, dL/dsigma | position gradient) = 32 bytes in shared memory, plus the
//    ray's SH basis (64 bytes, written once).  The row is rank one in (basis x dL/drgb), so it is expanded while it is
//    summed instead of being staged as 52 floats;
//  * the record of cell i is completed by the position gradient that the walk hands over one composited cell later
//    (quirk A.5.1) and only then routed, so position gradients ride in the row (floats 49..51) instead of costing a
//    reduction of their own; the last cell's record leaves with a zero position gradient, as upstream never flushes it;
//  * groups are summed by QUARTER warps, four groups per round: lane `sub` of a quarter owns SH coefficients
//    k = 2 sub, 2 sub + 1 (6 floats), lane 0 also the 4 trailing floats.  A group of >= 16 lanes takes a round of its
//    own: it is split over the four quarters by lane range and the partial rows are combined by shuffles.  Every
//    group goes through the warp's row cache (SLOTS rows, direct-mapped); two groups of one round that hash to the
//    same row are not scheduled together.  The round scheduler reads the groups from a list the leaders publish in
//    shared memory (no shuffles).  With MIN_GROUP = 2 lone lanes bypass all that and reduce their row directly.
template <int DEG, typename Faces, int SLOTS, int MIN_GROUP, int MIN_BLOCKS, bool REPLAY>
__global__ void __launch_bounds__(kBlock, MIN_BLOCKS)
    backward_pooled_kernel(const BackwardParams p, const Faces fa, const Tape tape) {
    if (tape.pool != nullptr) {
        const bool overflowed = tape.ctrl[1] != 0u;
        if (REPLAY == overflowed)
            return;
    }
    constexpr int GR = grad_row(DEG);
    constexpr int SR = sh_row(DEG);
    constexpr int NK = sh_dim(DEG);
    constexpr unsigned FULL = 0xffffffffu;
    static_assert(DEG == 3, "lane ownership (6 floats per lane, 8 lanes) is laid out for 16 SH coefficients");
    static_assert((SLOTS & (SLOTS - 1)) == 0 && SLOTS <= 32, "SLOTS: power of two, at most 32");

#ifdef RFB_EMU
    float *smem = rfb_emu_dynamic_smem();
#else
    extern __shared__ __align__(16) float smem[];
#endif
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31, sub = lane & 7, quarter = lane >> 3;
    constexpr int WARP_FLOATS = 32 * 8 + 32 * 16 + SLOTS * GR + 128 + SLOTS;
    float *rec = smem + warp * WARP_FLOATS;                              // [32][8]  (g0 g1 g2 ds | gx gy gz 0)
    float *bas = rec + 32 * 8;                                           // [32][16] SH basis of the lane's ray
    float *cache = bas + 32 * 16;                                        // [SLOTS][GR]
    uint4 *glist = reinterpret_cast<uint4 *>(cache + SLOTS * GR);        // [32] this iteration's groups
    uint32_t *tags = reinterpret_cast<uint32_t *>(glist + 32);           // [SLOTS]
    for (int i = lane; i < SLOTS; i += 32)
        tags[i] = kNone;

    uint32_t r;
    bool done = !thread_ray(p.num_rays, p.image_width, p.blocks_x, r);
    RayGeom ray = {0.f, 0.f, 0.f, 0.f, 0.f, 1.f};
    float sh[NK];
    BackwardRay st;
    uint32_t cur = 0;
    float4 pc = make_float4(0.f, 0.f, 0.f, 0.f);
    const uint32_t gwarp = blockIdx.x * (kBlock / 32) + warp;
    uint32_t nrec = 0, last_cell = 0;
    uint2 rec_a = make_uint2(0u, 0u);
    if (!done) {
        backward_ray_setup<DEG>(p, r, ray, sh, st);
        cur = __ldg(p.start + r);
        pc = ldg4(p.cells + cur);
        if (REPLAY) {
            uint2 pr = tape.per_ray[r];
            nrec = pr.x;
            last_cell = pr.y;
            done = nrec == 0;
        }
    } else {
#pragma unroll
        for (int i = 0; i < NK; ++i)
            sh[i] = 0.0f;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i)
        bas[lane * 16 + i] = i < NK ? sh[i] : 0.0f;
    __syncwarp();

    float t0 = 0.0f;
    uint32_t n = 0;
    uint32_t chunk_ahead = 0;
    uint2 rec_b = make_uint2(0u, 0u);
    auto tape_row = [&](uint32_t chunk_id, uint32_t j) -> uint2 {
        return j < nrec ? __ldcs(tape.pool + ((uint64_t)chunk_id * kTapeChunk + (j % kTapeChunk)) * 32 + lane)
                        : make_uint2(last_cell, 0u);
    };
    if (REPLAY) {
        chunk_ahead = tape.table[(uint64_t)gwarp * tape.table_stride];
        if (!done) {
            rec_a = tape_row(chunk_ahead, 0);
            rec_b = tape_row(chunk_ahead, 1);
        }
    }
    uint32_t pend_cell = kNone; // cell of the record waiting in rec[lane] for its position gradient
    float4 *my_rec = reinterpret_cast<float4 *>(rec + lane * 8);
    const float4 *rec4 = reinterpret_cast<const float4 *>(rec);
    const float2 *bas2 = reinterpret_cast<const float2 *>(bas);

    for (uint32_t k = 0;; ++k) {
        bool c_valid = false;
        float dL_ds = 0.0f;
        float dL_drgb[3] = {0.0f, 0.0f, 0.0f};
        uint32_t new_cell = kNone;
        bool emit = false;
        uint32_t emit_cell = kNone;

        bool step = false;
        float t1 = __int_as_float(0x7f800000);
        uint32_t nxt = 0;
        if (REPLAY) {
            if (((k + 2) % kTapeChunk) == 0)
                chunk_ahead = tape.table[(uint64_t)gwarp * tape.table_stride + (k + 2) / kTapeChunk];
            if (!done) {
                uint2 rec_c = tape_row(chunk_ahead, k + 2);
                t1 = __uint_as_float(rec_a.y);
                nxt = rec_b.x;
                rec_a = rec_b;
                rec_b = rec_c;
                step = true;
            }
        } else if (!done) {
            n++;
            if (n > p.max_steps) {
                done = true;
            } else {
                uint32_t begin, nf;
                fa.row(cur, begin, nf);
                uint32_t face = kNone;
                fa.scan(begin, nf, pc.x, pc.y, pc.z, ray, t1, face);
                if (face == kNone) {
                    done = true;
                } else {
                    nxt = fa.neighbour(begin, face);
                    step = true;
                }
            }
        }
        if (step) {
            float4 pn = ldg4(p.cells + nxt);
            if (t1 > t0) {
                float rgb[3] = {0.0f, 0.0f, 0.0f};
                if (pc.w > 1e-6f)
                    sh_to_rgb<DEG>(p.sh_rows + (uint64_t)cur * SR, sh, rgb[0], rgb[1], rgb[2]);
                float w, fx, fy, fz;
                bool flush;
                uint32_t flush_idx;
                bool go = st.cell(cur, pc, pn, t0, t1, rgb, ray, p.weight_threshold, dL_drgb, dL_ds, w, flush,
                                  flush_idx, fx, fy, fz);
                if (p.point_error)
                    add_point_error(p, cur, __fmul_rn(w, st.err));
                if (flush) { // flush_idx == pend_cell: the waiting record is complete now
#ifdef RFB_EMU
                    if (flush_idx != pend_cell)
                        __builtin_trap(); // invariant of the record pipeline, checked on the CPU emulator only
#endif
                    my_rec[1] = make_float4(fx, fy, fz, 0.0f);
                    emit = true;
                    emit_cell = pend_cell;
                }
                c_valid = true;
                new_cell = cur;
                done = !go;
            }
            t0 = fmaxf(t0, t1);
            cur = nxt;
            pc = pn;
            if (REPLAY && k + 1 >= nrec)
                done = true;
        } else if (pend_cell != kNone) {
            // the ray has ended: its last record leaves without a position gradient (never flushed upstream)
            my_rec[1] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            emit = true;
            emit_cell = pend_cell;
            pend_cell = kNone;
        }
        __syncwarp();

        // ---- warp-collective phase: sum the complete records by cell, four groups per round
        const unsigned grp = __match_any_sync(FULL, emit ? emit_cell : (0x80000000u | lane));
        // MIN_GROUP > 1: lanes in groups smaller than that reduce their own complete row directly, all at once (a
        // lone lane is 9 % of the lane-steps but 39 % of the groups, i.e. of the serial rounds)
        const bool direct = MIN_GROUP > 1 && emit && __popc(grp) < MIN_GROUP;
        if (direct) {
            const float4 lo = my_rec[0], hi = my_rec[1];
            float *grow = p.acc + (uint64_t)emit_cell * GR;
            if (lo.x != 0.0f || lo.y != 0.0f || lo.z != 0.0f) {
#pragma unroll
                for (int i = 0; i < SR; i += 4) {
                    float v[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int e = i + j, c = e % 3;
                        v[j] = e < 3 * NK ? __fmul_rn(sh[e / 3], c == 0 ? lo.x : (c == 1 ? lo.y : lo.z)) : 0.0f;
                    }
                    red_add_v4(grow + i, v[0], v[1], v[2], v[3]);
                }
            }
            red_add_v4(grow + SR, lo.w, hi.x, hi.y, hi.z);
        }
        const bool is_leader = emit && !direct && (uint32_t)(__ffs(grp) - 1) == lane;
        const unsigned leaders = __ballot_sync(FULL, is_leader);
        if (leaders == 0u) {
            // nothing to route in this iteration
        } else {
            // the leaders publish (lane mask, cell) in lane order so that the round scheduler reads them from shared
            // memory instead of shuffling them
            if (is_leader) // (lane mask, cell, cache row, lanes): everything the scheduler needs, computed once
                glist[__popc(leaders & ((1u << lane) - 1u))] =
                    make_uint4(grp, emit_cell, (emit_cell * 2654435761u) >> (32 - __builtin_ctz(SLOTS)),
                               (uint32_t)__popc(grp));
            __syncwarp();
        }
        const int num_groups = __popc(leaders);
        int next_group = 0;
        while (next_group < num_groups) {
            unsigned my_members = 0, used = 0;
            uint32_t my_cell = kNone, my_slot = 0;
            bool my_owner = false, split = false;
            int q_next = 0;
            while (next_group < num_groups && q_next < 4) {
                const uint4 g = glist[next_group];
                const unsigned gmask = g.x;
                const uint32_t cell = g.y, slot = g.z;
                const bool big = g.w >= 16u; // split over the four quarters by lane range
                if ((big && q_next != 0) || ((used >> slot) & 1u))
                    break; // next round (the first candidate of a round always fits)
                used |= 1u << slot;
                if (big) {
                    my_members = gmask & (0xFFu << (8 * quarter));
                    my_cell = cell;
                    my_slot = slot;
                    my_owner = quarter == 0;
                    split = true;
                    q_next = 4;
                } else {
                    if ((int)quarter == q_next) {
                        my_members = gmask;
                        my_cell = cell;
                        my_slot = slot;
                        my_owner = true;
                    }
                    q_next++;
                }
                next_group++;
            }
            float a[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, x[4] = {0.f, 0.f, 0.f, 0.f};
            for (unsigned mm = my_members; mm; mm &= mm - 1) {
                // (two members per trip was tried: +12 live registers -> 48-72 bytes of spill at 5 CTAs/SM)
                const int m = __ffs(mm) - 1;
                const float4 lo = rec4[2 * m], hi = rec4[2 * m + 1];
                const float2 b = bas2[m * 8 + sub];
                a[0] = __fmaf_rn(b.x, lo.x, a[0]);
                a[1] = __fmaf_rn(b.x, lo.y, a[1]);
                a[2] = __fmaf_rn(b.x, lo.z, a[2]);
                a[3] = __fmaf_rn(b.y, lo.x, a[3]);
                a[4] = __fmaf_rn(b.y, lo.y, a[4]);
                a[5] = __fmaf_rn(b.y, lo.z, a[5]);
                x[0] += lo.w;
                x[1] += hi.x;
                x[2] += hi.y;
                x[3] += hi.z;
            }
            if (split) { // partial rows of the four quarters
#pragma unroll
                for (int i = 0; i < 6; ++i) {
                    a[i] += __shfl_xor_sync(FULL, a[i], 8);
                    a[i] += __shfl_xor_sync(FULL, a[i], 16);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    x[i] += __shfl_xor_sync(FULL, x[i], 8);
                    x[i] += __shfl_xor_sync(FULL, x[i], 16);
                }
            }
            // cache update by the owning quarter; lane `sub` holds floats [6 sub, 6 sub + 6), lane 0 also [SR, SR + 4).
            // A row that has to make room leaves as 13 x 16-byte reductions (the quarter's lanes take float4 i and
            // i + 8): same sectors per row as the kernel above, not 3 x 8 bytes per lane.
            const bool update = my_owner && my_cell != kNone;
            uint32_t tag = kNone;
            float *crow = cache + my_slot * GR;
            if (update)
                tag = tags[my_slot];
            const bool hit = tag == my_cell;
            if (update && !hit && tag != kNone) {
                float *grow = p.acc + (uint64_t)tag * GR;
                for (int i = (int)sub; i < GR / 4; i += 8) {
                    const float4 v = *reinterpret_cast<const float4 *>(crow + 4 * i);
                    red_add_v4(grow + 4 * i, v.x, v.y, v.z, v.w);
                }
            }
            __syncwarp(); // the old row and its tag have been read
            if (update) {
                float2 *mine = reinterpret_cast<float2 *>(crow + 6 * sub); // 8-byte aligned: rows are 208 bytes
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    float2 v = hit ? mine[i] : make_float2(0.0f, 0.0f);
                    v.x += a[2 * i];
                    v.y += a[2 * i + 1];
                    mine[i] = v;
                }
                if (sub == 0) {
                    float4 *tail = reinterpret_cast<float4 *>(crow + SR);
                    float4 v = hit ? *tail : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                    v.x += x[0];
                    v.y += x[1];
                    v.z += x[2];
                    v.w += x[3];
                    *tail = v;
                    tags[my_slot] = my_cell;
                }
            }
            __syncwarp();
        }
        // the step's own record starts waiting (after the phase above has read the slot)
        if (c_valid) {
            my_rec[0] = make_float4(dL_drgb[0], dL_drgb[1], dL_drgb[2], dL_ds);
            pend_cell = new_cell;
        }
        if (!__any_sync(FULL, !done || pend_cell != kNone))
            break;
    }

    // drain the cache: quarter q writes rows q, q + 4, ...
    __syncwarp();
    for (int slot = (int)quarter; slot < SLOTS; slot += 4) {
        const uint32_t tag = tags[slot];
        if (tag == kNone)
            continue;
        const float *crow = cache + slot * GR;
        float *grow = p.acc + (uint64_t)tag * GR;
        for (int i = (int)sub; i < GR / 4; i += 8) {
            const float4 v = *reinterpret_cast<const float4 *>(crow + 4 * i);
            red_add_v4(grow + 4 * i, v.x, v.y, v.z, v.w);
        }
    }
}

// accumulator -> reference-layout gradient outputs (+ optional finite scrub)
template <typename AttrT>
__global__ void finalize_grads_kernel(const float *__restrict__ acc, uint32_t num_points,
                                      int attr_dim_, int sh_row_, float *__restrict__ points_grad,
                                      AttrT *__restrict__ attr_grad, int scrub) {
    // one warp per accumulator row (grid-stride), row-contiguous reads and writes
    const int gr = sh_row_ + 4;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < num_points; i += warps) {
        const float *row = acc + (uint64_t)i * gr;
        AttrT *out = attr_grad + (uint64_t)i * attr_dim_;
        for (int s = lane; s < attr_dim_; s += 32) {
            float v = (s < attr_dim_ - 1) ? row[s] : row[sh_row_];
            AttrT o = (AttrT)v;
            if (scrub && !isfinite((float)o))
                o = (AttrT)0.0f;
            out[s] = o;
        }
        if (lane < 3) {
            float gq = row[sh_row_ + 1 + lane];
            if (scrub && !isfinite(gq))
                gq = 0.0f;
            points_grad[3 * (uint64_t)i + lane] = gq;
        }
    }
}

// ------------------------------------------------------------------ entry cell
// Nearest point of each query = the Voronoi cell that contains it = the cell a ray from that
// origin starts in (what scene.py:224-234 gets from radfoam.nn over the AABB tree,
// aabb_tree.cu:343-415).  A camera origin is shared by every ray of a frame (and a training
// batch has a few hundred distinct origins), so an exact brute-force scan -- one CTA per query,
// 12 N bytes streamed, mostly from L2 -- is microseconds per query and needs no tree.
// Ties (measure zero) resolve to the lowest index.  Distances in the x0 + (x1 + x2) order.
__global__ void __launch_bounds__(256) nearest_point_kernel(const float *__restrict__ points,
                                                            uint32_t num_points,
                                                            const float *__restrict__ queries,
                                                            uint32_t *__restrict__ out) {
    const float qx = queries[3 * (uint64_t)blockIdx.x], qy = queries[3 * (uint64_t)blockIdx.x + 1],
                qz = queries[3 * (uint64_t)blockIdx.x + 2];
    float best = __int_as_float(0x7f800000);
    uint32_t best_i = kNone;
    for (uint32_t i = threadIdx.x; i < num_points; i += blockDim.x) {
        float dx = points[3 * (uint64_t)i] - qx, dy = points[3 * (uint64_t)i + 1] - qy,
              dz = points[3 * (uint64_t)i + 2] - qz;
        float d2 = __fmaf_rn(dx, dx, __fmaf_rn(dy, dy, __fmul_rn(dz, dz)));
        if (d2 < best) { // strict: the lowest index among this thread's equals stays
            best = d2;
            best_i = i;
        }
    }
    // (distance, index) lexicographic min over the CTA
    __shared__ float s_d[8];
    __shared__ uint32_t s_i[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float od = __shfl_down_sync(0xffffffffu, best, o);
        uint32_t oi = __shfl_down_sync(0xffffffffu, best_i, o);
        if (od < best || (od == best && oi < best_i)) {
            best = od;
            best_i = oi;
        }
    }
    if ((threadIdx.x & 31) == 0) {
        s_d[threadIdx.x >> 5] = best;
        s_i[threadIdx.x >> 5] = best_i;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w)
            if (s_d[w] < best || (s_d[w] == best && s_i[w] < best_i)) {
                best = s_d[w];
                best_i = s_i[w];
            }
        out[blockIdx.x] = best_i;
    }
}

// ------------------------------------------------------------------ farthest neighbour
// Per cell: the adjacent point farthest from it and half the mean neighbour distance ("cell radius"),
// what the densification pass reads (scene.py:434-461; reference kernel triangulation_ops.cu:9-44, one
// thread per point walking its row with dependent gathers of three scalar loads each).
// Arithmetic as the reference's SASS has it: d = q - p, |d|^2 = fma(dx,dx, fma(dy,dy, dz*dz)), IEEE sqrt and
// divide, strict '>' first-max from 0, and `sum += 0.5 * dist` evaluated in fp64 and rounded back to fp32 every
// iteration.  That last step needs no fp64: 0.5*dist is exact and rounding an exact sum to 53 then to 24 bits
// equals rounding it to 24 bits directly whenever 53 >= 2*24 + 2 (double rounding is innocuous for +), so
// fmaf(dist, 0.5f, sum) is bit-identical; the FP64CHAIN variants keep the literal form.
constexpr uint32_t kRowLanes = 8;

struct PackedPoints { // the caller's [N][3] array: three scalar loads per point
    const float *p;
    __device__ __forceinline__ float3 operator[](uint64_t i) const {
        return make_float3(p[3 * i], p[3 * i + 1], p[3 * i + 2]);
    }
};
struct PaddedPoints { // a float4 mirror: one 16-byte load per point
    const float4 *p;
    __device__ __forceinline__ float3 operator[](uint64_t i) const {
        const float4 v = __ldg(p + i);
        return make_float3(v.x, v.y, v.z);
    }
};

__global__ void __launch_bounds__(256) pad_points_kernel(const float *__restrict__ points, uint32_t num_points,
                                                         float4 *__restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < num_points)
        out[i] = make_float4(points[3 * i], points[3 * i + 1], points[3 * i + 2], 0.0f);
}

__device__ __forceinline__ float neighbour_distance(float3 q, float3 p) {
    const float dx = __fsub_rn(q.x, p.x), dy = __fsub_rn(q.y, p.y), dz = __fsub_rn(q.z, p.z);
    return __fsqrt_rn(__fmaf_rn(dx, dx, __fmaf_rn(dy, dy, __fmul_rn(dz, dz))));
}

template <bool FP64CHAIN>
__device__ __forceinline__ float half_distance_sum(float sum, float dist) {
    if (FP64CHAIN)
        return __double2float_rn(__fma_rn((double)dist, 0.5, (double)sum));
    return __fmaf_rn(dist, 0.5f, sum);
}

// 8 lanes share a row: the adjacency read is one 32-byte sector and the 8 neighbour gathers are in flight
// together; the reference's in-order accumulation is then replayed over the 8 distances by shuffle.
template <typename Points, bool FP64CHAIN>
__global__ void __launch_bounds__(256) farthest_neighbor_kernel(Points points,
                                                                const uint32_t *__restrict__ adjacency,
                                                                const uint32_t *__restrict__ offsets,
                                                                uint32_t num_points,
                                                                uint32_t *__restrict__ indices,
                                                                float *__restrict__ cell_radius) {
    const uint32_t lane = threadIdx.x & 31u, sub = lane & (kRowLanes - 1u);
    const uint32_t group_mask = 0xffu << (lane & ~(kRowLanes - 1u));
    const uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) / kRowLanes;
    if (i >= num_points) // the 8 lanes of a row leave together
        return;
    const float3 p = points[i];
    const uint32_t begin = offsets[i], num_faces = offsets[i + 1] - begin;
    float sum = 0.0f, farthest = 0.0f;
    uint32_t farthest_face = kNone;
    for (uint32_t f0 = 0; f0 < num_faces; f0 += kRowLanes) {
        const uint32_t f = f0 + sub;
        float dist = 0.0f;
        if (f < num_faces)
            dist = neighbour_distance(points[adjacency[begin + f]], p);
        const uint32_t count = min(kRowLanes, num_faces - f0);
        for (uint32_t k = 0; k < count; ++k) { // every lane of the row replays the same sequence
            const float dk = __shfl_sync(group_mask, dist, k, kRowLanes);
            sum = half_distance_sum<FP64CHAIN>(sum, dk);
            if (dk > farthest) {
                farthest = dk;
                farthest_face = f0 + k;
            }
        }
    }
    if (sub == 0) {
        indices[i] = farthest_face == kNone ? kNone : adjacency[begin + farthest_face];
        cell_radius[i] = __fdiv_rn(sum, __uint2float_rn(num_faces)); // 0/0 = NaN for an empty row, as upstream
    }
}

// One thread per row (the reference's shape), for comparison: fewer instructions per face, scattered reads.
template <typename Points, bool FP64CHAIN>
__global__ void __launch_bounds__(256) farthest_neighbor_rows_kernel(Points points,
                                                                     const uint32_t *__restrict__ adjacency,
                                                                     const uint32_t *__restrict__ offsets,
                                                                     uint32_t num_points,
                                                                     uint32_t *__restrict__ indices,
                                                                     float *__restrict__ cell_radius) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= num_points)
        return;
    const float3 p = points[i];
    const uint32_t begin = offsets[i], num_faces = offsets[i + 1] - begin;
    float sum = 0.0f, farthest = 0.0f;
    uint32_t farthest_idx = kNone;
#pragma unroll 4
    for (uint32_t f = 0; f < num_faces; ++f) {
        const uint32_t j = adjacency[begin + f];
        const float dist = neighbour_distance(points[j], p);
        sum = half_distance_sum<FP64CHAIN>(sum, dist);
        if (dist > farthest) {
            farthest = dist;
            farthest_idx = j;
        }
    }
    indices[i] = farthest_idx;
    cell_radius[i] = __fdiv_rn(sum, __uint2float_rn(num_faces));
}

// ------------------------------------------------------------------ benchmark
struct CameraParams {
    float position[3], forward[3], right[3], up[3];
    float fov;
    uint32_t width, height;
    int model;
};

// cast_ray (camera.h:56-85)
__device__ __forceinline__ void cast_ray(const CameraParams &c, int i, int j, RayGeom &ray) {
    float aspect = (float)c.width / (float)c.height;
    float x = (float)i / (float)c.width;
    float y = (float)j / (float)c.height;
    float u = (2.0f * x - 1.0f) * aspect;
    float v = (1.0f - 2.0f * y);
    float mask = 1.0f;
    float d[3];
    if (c.model == 0) {
        float w = 1.0f / tanf(c.fov * 0.5f);
#pragma unroll
        for (int k = 0; k < 3; ++k)
            d[k] = w * c.forward[k] + u * c.right[k] + v * c.up[k];
    } else {
        float theta = atan2f(v, u);
        float phi = c.fov * sqrtf(u * u + v * v);
        if (phi >= 3.14159265358979323846f) {
            phi = 3.14159265358979323846f - 1e-6f;
            mask = 0.0f;
        }
#pragma unroll
        for (int k = 0; k < 3; ++k)
            d[k] = sinf(phi) * cosf(theta) * c.right[k] + sinf(phi) * sinf(theta) * c.up[k] +
                   cosf(phi) * c.forward[k];
    }
    float n2 = __fmaf_rn(d[0], d[0], __fmaf_rn(d[1], d[1], __fmul_rn(d[2], d[2])));
    if (n2 > 0.0f) {
        float n = __fsqrt_rn(n2);
        d[0] = __fdiv_rn(d[0], n);
        d[1] = __fdiv_rn(d[1], n);
        d[2] = __fdiv_rn(d[2], n);
    }
    ray.ox = c.position[0];
    ray.oy = c.position[1];
    ray.oz = c.position[2];
    ray.dx = d[0] * mask;
    ray.dy = d[1] * mask;
    ray.dz = d[2] * mask;
}

// make_rgba8 (tracing_utils.cuh:105-115): clamp, truncate
__device__ __forceinline__ uint32_t pack_rgba8(float r, float g, float b, float a) {
    r = fmaxf(0.0f, fminf(1.0f, r));
    g = fmaxf(0.0f, fminf(1.0f, g));
    b = fmaxf(0.0f, fminf(1.0f, b));
    a = fmaxf(0.0f, fminf(1.0f, a));
    int ri = (int)(r * 255.0f), gi = (int)(g * 255.0f), bi = (int)(b * 255.0f), ai = (int)(a * 255.0f);
    return ((uint32_t)ai << 24) | ((uint32_t)bi << 16) | ((uint32_t)gi << 8) | (uint32_t)ri;
}

struct BenchmarkParams {
    const float4 *cells;
    const float *sh_rows;
    const uint32_t *start; // one element
    uint32_t *out;
    CameraParams cam;
    uint32_t blocks_x;
    float weight_threshold;
    uint32_t max_steps;
};

template <int DEG, typename Faces>
__global__ void __launch_bounds__(kBlock) benchmark_kernel(const BenchmarkParams p, const Faces fa) {
    uint32_t r;
    if (!thread_ray(p.cam.width * p.cam.height, p.cam.width, p.blocks_x, r))
        return;
    uint32_t pi = r % p.cam.width, pj = r / p.cam.width;
    RayGeom ray;
    cast_ray(p.cam, (int)pi, (int)pj, ray);
    float nrm = __fsqrt_rn(__fmaf_rn(ray.dx, ray.dx, __fmaf_rn(ray.dy, ray.dy, __fmul_rn(ray.dz, ray.dz))));
    if (nrm < 0.1f) {
        p.out[r] = 0;
        return;
    }
    float sh[sh_dim(DEG)];
    sh_basis<DEG>(ray.dx, ray.dy, ray.dz, sh);
    float T = 1.0f, cr = 0.0f, cg = 0.0f, cb = 0.0f;
    auto cell_fn = [&](uint32_t cell, const float4 &pc, float t0, float t1, const float4 &) -> bool {
        float s = pc.w;
        float r_ = 0.0f, g_ = 0.0f, b_ = 0.0f;
        if (s > 1e-6f)
            sh_to_rgb<DEG>(p.sh_rows + (uint64_t)cell * sh_row(DEG), sh, r_, g_, b_);
        float delta = fmaxf(__fsub_rn(t1, t0), 0.0f);
        float alpha = 1.0f - expf(-s * delta);
        float w = __fmul_rn(T, alpha);
        cr = __fmaf_rn(w, r_, cr);
        cg = __fmaf_rn(w, g_, cg);
        cb = __fmaf_rn(w, b_, cb);
        T = __fmul_rn(T, __fsub_rn(1.0f, alpha));
        return T > p.weight_threshold;
    };
    walk(fa, p.cells, ray, __ldg(p.start), p.max_steps, cell_fn);
    p.out[r] = pack_rgba8(cr, cg, cb, 1.0f);
}

} // namespace rfb
