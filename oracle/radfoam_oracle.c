/*
 * radfoam_oracle.c -- CPU restatement of radfoam's src/tracing hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (radfoam_b200/) may
 * import, link or call this file; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg use it, and only as the checker / the reported
 * CPU baseline.
 *
 * Parity pin: the reference ships no tests, golden vectors or fixtures for this
 * path.  This restatement is pinned against outputs of the
 * reference's OWN kernels (oracle/_ref: /root/reference/src/tracing/pipeline.cu
 * compiled unmodified against oracle/eigen_shim) run on a B200; the vectors are
 * committed under tests/golden/ with the script that made them.
 *
 * Each function cites the reference file:line it restates (paths relative to
 * /root/reference).  Arithmetic that decides the integer traversal is written
 * with explicit fmaf() in the association the reference's sm_100 SASS uses
 * (nvcc -fmad contraction on top of Eigen 3.4's x0 + (x1 + x2) reduction
 * order; read off `cuobjdump -sass oracle/_ref/libradfoam_ref.so`):
 *     dp  = fma(ox, dx, fma(oy, dy, oz * dz))
 *     num = fma(ox, fx - rx, fma(oy, fy - ry, oz * (fz - rz))),  f = P + o/2
 *     t   = num / dp                      (IEEE RN division)
 * The file is compiled with -ffp-contract=off so the compiler adds no fusion
 * of its own.  expf/logf are libm's (the GPU's are libdevice's MUFU-based
 * versions), so composited floats agree to ~1 ulp-level, not bit-for-bit;
 * the cell-to-cell walk itself is bit-reproducible.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef _Float16 half_t;

#define RF_NONE 0xFFFFFFFFu

typedef struct {
    float origin[3];
    float direction[3];
} rfo_ray; /* src/tracing/camera.h:7-10 */

/* ---- SH basis constants, src/tracing/sh_utils.cuh:8-30 ---- */
static const float C0 = 0.28209479177387814f;
static const float C1 = 0.4886025119029199f;
static const float C2[5] = {1.0925484305920792f, -1.0925484305920792f,
                            0.31539156525252005f, -1.0925484305920792f,
                            0.5462742152960396f};
static const float C3[7] = {-0.5900435899266435f, 2.890611442640554f,
                            -0.4570457994644658f, 0.3731763325901154f,
                            -0.4570457994644658f, 1.445305721320277f,
                            -0.5900435899266435f};

/* sh_coefficients<degree>, src/tracing/sh_utils.cuh:34-70.  Products/sums follow
 * the source's left-to-right order with the fusions nvcc applies (a*b - c ->
 * fma(a, b, -c)). */
static void sh_coefficients(int degree, const float d[3], float sh[16]) {
    float x = d[0], y = d[1], z = d[2];
    for (int i = 0; i < 16; ++i)
        sh[i] = 0.0f;
    sh[0] = C0;
    if (degree > 0) {
        sh[1] = -C1 * y;
        sh[2] = C1 * z;
        sh[3] = -C1 * x;
    }
    float xx = x * x, yy = y * y, zz = z * z;
    float xy = x * y, yz = y * z, xz = x * z;
    if (degree > 1) {
        sh[4] = C2[0] * xy;
        sh[5] = C2[1] * yz;
        sh[6] = C2[2] * (((zz + zz) - xx) - yy);
        sh[7] = C2[3] * xz;
        sh[8] = C2[4] * (xx - yy);
    }
    if (degree > 2) {
        sh[9] = (C3[0] * y) * fmaf(3.0f, xx, -yy);
        sh[10] = (C3[1] * xy) * z;
        sh[11] = (C3[2] * y) * (fmaf(4.0f, zz, -xx) - yy);
        sh[12] = (C3[3] * z) * fmaf(-3.0f, yy, fmaf(-3.0f, xx, zz + zz));
        sh[13] = (C3[4] * x) * (fmaf(4.0f, zz, -xx) - yy);
        sh[14] = (C3[5] * z) * (xx - yy);
        sh[15] = (C3[6] * x) * fmaf(-3.0f, yy, xx);
    }
}

static inline float attr_at(const void *attrs, int is_half, size_t i) {
    return is_half ? (float)((const half_t *)attrs)[i] : ((const float *)attrs)[i];
}

/* load_sh_as_rgb, src/tracing/sh_utils.cuh:72-83: rgb[c] = max(0, 0.5 + sum_k
 * coeff[k] * attr[3k + c]) accumulated k = 0.. in one FFMA chain per channel. */
static void load_sh_as_rgb(int sh_dim, const float *coeffs, const void *attrs,
                           int is_half, size_t base, float rgb[3]) {
    rgb[0] = rgb[1] = rgb[2] = 0.5f;
    for (int i = 0; i < 3 * sh_dim; ++i)
        rgb[i % 3] = fmaf(coeffs[i / 3], attr_at(attrs, is_half, base + i), rgb[i % 3]);
    for (int c = 0; c < 3; ++c)
        rgb[c] = (rgb[c] < 0.0f) ? 0.0f : rgb[c]; /* cwiseMax(0) as built: NaN stays NaN */
}

/* load_attributes lambda, src/tracing/pipeline.cu:47-55 */
static void load_attributes(int sh_dim, const float *coeffs, const void *attrs,
                            int is_half, uint32_t idx, float rgb[3], float *s) {
    size_t A = (size_t)(1 + 3 * sh_dim);
    *s = attr_at(attrs, is_half, idx * A + (A - 1));
    if (*s > 1e-6f)
        load_sh_as_rgb(sh_dim, coeffs, attrs, is_half, idx * A, rgb);
    else
        rgb[0] = rgb[1] = rgb[2] = 0.0f;
}

/* ray.direction /= ray.direction.norm(), src/tracing/pipeline.cu:39-40 */
static void normalize_direction(float d[3]) {
    float n2 = fmaf(d[0], d[0], fmaf(d[1], d[1], d[2] * d[2]));
    float n = sqrtf(n2);
    d[0] = d[0] / n;
    d[1] = d[1] / n;
    d[2] = d[2] / n;
}

/* prefetch_adjacent_diff_kernel, src/tracing/pipeline.cu:546-568.
 * adjacent_diff: E x 4 halfs = (RN_half(q - p).xyz, 0). */
void rfo_prefetch_adjacent_diff(const float *points, uint32_t num_points,
                                const uint32_t *adj, const uint32_t *off,
                                uint16_t *adjacent_diff) {
    half_t *out = (half_t *)adjacent_diff;
    for (uint32_t i = 0; i < num_points; ++i) {
        const float *p = points + 3 * (size_t)i;
        for (uint32_t e = off[i]; e < off[i + 1]; ++e) {
            const float *q = points + 3 * (size_t)adj[e];
            out[4 * (size_t)e + 0] = (half_t)(q[0] - p[0]);
            out[4 * (size_t)e + 1] = (half_t)(q[1] - p[1]);
            out[4 * (size_t)e + 2] = (half_t)(q[2] - p[2]);
            out[4 * (size_t)e + 3] = (half_t)0.0f;
        }
    }
}

/* One step of trace(): scan the faces of `cur` (src/tracing/tracing_utils.cuh:43-67).
 * Returns the winning face (RF_NONE if none) and t_1. */
static inline uint32_t scan_faces(const half_t *diff, uint32_t begin, uint32_t nf,
                                  const float P[3], const float o[3],
                                  const float d[3], float *t1_out) {
    float t1 = INFINITY;
    uint32_t next_face = RF_NONE;
    for (uint32_t f = 0; f < nf; ++f) {
        const half_t *h = diff + 4 * (size_t)(begin + f);
        float ox = (float)h[0], oy = (float)h[1], oz = (float)h[2];
        float fx = fmaf(ox, 0.5f, P[0]);
        float fy = fmaf(oy, 0.5f, P[1]);
        float fz = fmaf(oz, 0.5f, P[2]);
        float dp = fmaf(ox, d[0], fmaf(oy, d[1], oz * d[2]));
        float num = fmaf(ox, fx - o[0], fmaf(oy, fy - o[1], oz * (fz - o[2])));
        float t = num / dp;
        if (dp > 0.0f && t < t1) {
            t1 = t;
            next_face = f;
        }
    }
    *t1_out = t1;
    return next_face;
}

/* cell_intersection_grad, src/tracing/tracing_utils.cuh:91-103, in the reference's SASS
 * association: n = q - p; a = fma(p + q, 0.5, -o); dp = fma(nx,dx,fma(ny,dy,nz*dz));
 * num = fma(nx,ax,fma(ny,ay,nz*az)); g_i = fma(num, d_i, dp*(o_i - p_i)) / (dp*dp). */
static void cell_intersection_grad(const float p[3], const float q[3],
                                   const float o[3], const float d[3], float g[3]) {
    float n[3], a[3];
    for (int i = 0; i < 3; ++i) {
        n[i] = q[i] - p[i];
        a[i] = fmaf(p[i] + q[i], 0.5f, -o[i]);
    }
    float dp = fmaf(n[0], d[0], fmaf(n[1], d[1], n[2] * d[2]));
    float num = fmaf(n[0], a[0], fmaf(n[1], a[1], n[2] * a[2]));
    float dp2 = dp * dp;
    for (int i = 0; i < 3; ++i)
        g[i] = fmaf(num, d[i], dp * (o[i] - p[i])) / dp2;
}

static inline void atomic_addf(float *dst, float v, int parallel) {
    if (parallel) {
#pragma omp atomic
        *dst += v;
    } else {
        *dst += v;
    }
}

static inline void store_attr(void *dst, int is_half, size_t i, float v) {
    if (is_half)
        ((half_t *)dst)[i] = (half_t)v;
    else
        ((float *)dst)[i] = v;
}

/* atomicAdd(attr_scalar*, (attr_scalar)v): half mode rounds the addend and the sum to half */
static inline void add_attr(void *dst, int is_half, size_t i, float v, int parallel) {
    if (is_half) {
        half_t *p = (half_t *)dst + i;
        half_t hv = (half_t)v;
#pragma omp critical(rfo_half_add)
        *p = (half_t)(*p + hv);
    } else {
        atomic_addf((float *)dst + i, v, parallel);
    }
}

/* forward kernel + trace(), src/tracing/pipeline.cu:14-130 and
 * src/tracing/tracing_utils.cuh:8-89, one ray. */
static void forward_ray(int sh_degree, int is_half, float weight_threshold,
                        uint32_t max_steps, const float *points, const void *attrs,
                        const uint32_t *adj, const uint32_t *off, const half_t *diff,
                        const float *ray, uint32_t start, uint32_t Q, const float *dq,
                        float rgba[4], float *qdepth, uint32_t *qidx, uint32_t *nint,
                        void *contrib, int parallel) {
    int sh_dim = (sh_degree + 1) * (sh_degree + 1);
    float o[3] = {ray[0], ray[1], ray[2]};
    float d[3] = {ray[3], ray[4], ray[5]};
    normalize_direction(d);
    float sh[16];
    sh_coefficients(sh_degree, d, sh);

    float T = 1.0f;
    float acc[3] = {0.0f, 0.0f, 0.0f};
    uint32_t qi = 0;
    float cq = (dq && Q > 0) ? dq[0] : 0.0f;

    float t0 = 0.0f;
    uint32_t n = 0;
    uint32_t cur = start;
    float P[3] = {points[3 * (size_t)cur], points[3 * (size_t)cur + 1], points[3 * (size_t)cur + 2]};

    for (;;) {
        n++;
        if (n > max_steps)
            break;
        uint32_t b = off[cur], nf = off[cur + 1] - b;
        float t1;
        uint32_t face = scan_faces(diff, b, nf, P, o, d, &t1);
        if (face == RF_NONE)
            break;
        uint32_t nxt = adj[b + face];
        float Pn[3] = {points[3 * (size_t)nxt], points[3 * (size_t)nxt + 1], points[3 * (size_t)nxt + 2]};
        if (t1 > t0) {
            /* cell functor, src/tracing/pipeline.cu:66-102 */
            float rgb[3], s;
            load_attributes(sh_dim, sh, attrs, is_half, cur, rgb, &s);
            float delta = fmaxf(t1 - t0, 0.0f);
            float alpha = 1.0f - expf(-s * delta);
            float w = T * alpha;
            if (contrib)
                add_attr(contrib, is_half, cur, w, parallel);
            for (int c = 0; c < 3; ++c)
                acc[c] = fmaf(w, rgb[c], acc[c]);
            float Tn = T * (1.0f - alpha);
            while (qi < Q && Tn < cq) {
                qdepth[qi] = t0 + logf(T / cq) / s;
                qidx[qi] = cur;
                qi++;
                if (qi < Q)
                    cq = dq[qi];
            }
            T = Tn;
            if (!(T > weight_threshold))
                break;
        }
        t0 = fmaxf(t0, t1);
        cur = nxt;
        P[0] = Pn[0];
        P[1] = Pn[1];
        P[2] = Pn[2];
    }
    while (qi < Q) {
        qdepth[qi] = -1.0f;
        qidx[qi] = RF_NONE;
        qi++;
    }
    rgba[0] = acc[0];
    rgba[1] = acc[1];
    rgba[2] = acc[2];
    rgba[3] = 1.0f - T;
    *nint = n;
}

/* Pipeline::trace_forward, src/tracing/pipeline.cu:595-643 (prefetch + forward). */
int rfo_trace_forward(int sh_degree, int attr_is_half, float weight_threshold,
                      uint32_t max_intersections, uint32_t num_points,
                      const float *points, const void *attributes,
                      uint32_t point_adjacency_size, const uint32_t *adj,
                      const uint32_t *off, uint32_t num_rays, const float *rays,
                      const uint32_t *start_point_index, uint32_t Q,
                      const float *depth_quantiles, void *ray_rgba,
                      float *quantile_depths, uint32_t *quantile_point_indices,
                      uint32_t *num_intersections, void *point_contribution,
                      int num_threads) {
    if (sh_degree < 0 || sh_degree > 3)
        return 1;
    half_t *diff = (half_t *)malloc(sizeof(half_t) * 4 * ((size_t)point_adjacency_size + 32));
    if (!diff)
        return 2;
    rfo_prefetch_adjacent_diff(points, num_points, adj, off, (uint16_t *)diff);
    int parallel = num_threads != 1;
#ifdef _OPENMP
    if (num_threads > 0)
        omp_set_num_threads(num_threads);
#endif
#pragma omp parallel for schedule(dynamic, 256) if (parallel)
    for (int64_t r = 0; r < (int64_t)num_rays; ++r) {
        float rgba[4];
        uint32_t n;
        float qd_tmp[64];
        uint32_t qi_tmp[64];
        float *qd = quantile_depths ? quantile_depths + (size_t)r * Q : qd_tmp;
        uint32_t *qi = quantile_point_indices ? quantile_point_indices + (size_t)r * Q : qi_tmp;
        forward_ray(sh_degree, attr_is_half, weight_threshold, max_intersections, points,
                    attributes, adj, off, diff, rays + 6 * (size_t)r, start_point_index[r],
                    depth_quantiles ? Q : 0, depth_quantiles ? depth_quantiles + (size_t)r * Q : NULL,
                    rgba, qd, qi, &n, point_contribution, parallel);
        for (int c = 0; c < 4; ++c)
            store_attr(ray_rgba, attr_is_half, 4 * (size_t)r + c, rgba[c]);
        if (num_intersections)
            num_intersections[r] = n;
    }
    free(diff);
    return 0;
}

/* backward kernel, src/tracing/pipeline.cu:132-343, one ray. */
static void backward_ray(int sh_degree, int is_half, float weight_threshold,
                         uint32_t max_steps, const float *points, const void *attrs,
                         const uint32_t *adj, const uint32_t *off, const half_t *diff,
                         const float *ray, uint32_t start, uint32_t Q, const float *dq,
                         const uint32_t *qidx, const float rgba[4], const float rgba_grad[4],
                         const float *depth_grad, const float *ray_error, float *points_grad,
                         void *attr_grad, void *point_error, int parallel) {
    int sh_dim = (sh_degree + 1) * (sh_degree + 1);
    size_t A = (size_t)(1 + 3 * sh_dim);
    float o[3] = {ray[0], ray[1], ray[2]};
    float d[3] = {ray[3], ray[4], ray[5]};
    normalize_direction(d);
    float sh[16];
    sh_coefficients(sh_degree, d, sh);

    float error = ray_error ? *ray_error : 0.0f;

    uint32_t qi = 0;
    float cq = (dq && Q > 0) ? dq[0] : 0.0f;
    float cdg = 0.0f; /* current_depth_grad, pipeline.cu:196-207 */
    for (uint32_t i = 0; i < Q; ++i) {
        if (qidx[i] != RF_NONE) {
            float s = attr_at(attrs, is_half, qidx[i] * A + (A - 1));
            cdg += depth_grad[i] / s;
        }
    }

    float T = 1.0f;
    float acc[3] = {0, 0, 0};
    uint32_t prev_idx = RF_NONE;
    float prev_point[3] = {0, 0, 0};
    float prev_grad[3] = {0, 0, 0};
    float cur_grad[3] = {0, 0, 0};
    float next_grad[3] = {0, 0, 0};

    float t0 = 0.0f;
    uint32_t n = 0;
    uint32_t cur = start;
    float P[3] = {points[3 * (size_t)cur], points[3 * (size_t)cur + 1], points[3 * (size_t)cur + 2]};

    for (;;) {
        n++;
        if (n > max_steps)
            break;
        uint32_t b = off[cur], nf = off[cur + 1] - b;
        float t1;
        uint32_t face = scan_faces(diff, b, nf, P, o, d, &t1);
        if (face == RF_NONE)
            break;
        uint32_t nxt = adj[b + face];
        float Pn[3] = {points[3 * (size_t)nxt], points[3 * (size_t)nxt + 1], points[3 * (size_t)nxt + 2]};
        if (t1 > t0) {
            /* cell functor, src/tracing/pipeline.cu:219-331; multiply-adds associated as in
             * the reference's sm_100 SASS (gradient sums cancel heavily, so a differently
             * fused multiply-add is visible at 1e-4 of the result) */
            float rgb[3], s;
            load_attributes(sh_dim, sh, attrs, is_half, cur, rgb, &s);
            float delta = fmaxf(t1 - t0, 0.0f);
            float alpha = 1.0f - expf(s * -delta);
            float oma = 1.0f - alpha;
            float omae = oma + 1e-6f;
            float denom = omae * T;
            float w = alpha * T;
            for (int c = 0; c < 3; ++c)
                acc[c] = fmaf(w, rgb[c], acc[c]);
            if (point_error)
                add_attr(point_error, is_half, cur, w * error, parallel);

            float rest[3];
            for (int c = 0; c < 3; ++c)
                rest[c] = (rgba[c] - acc[c]) / denom;
            float dot = fmaf(rgba_grad[0], rgb[0] - rest[0],
                             fmaf(rgba_grad[1], rgb[1] - rest[1], rgba_grad[2] * (rgb[2] - rest[2])));
            float k_alpha = (1.0f - rgba[3]) * rgba_grad[3];
            float dL_dalpha = fmaf(dot, T, k_alpha / omae);
            float Tn = oma * T;
            float dL_ds = dL_dalpha * (delta * oma);
            float dL_dt0 = 0.0f;
            while (qi < Q && Tn < cq) {
                float gq = depth_grad[qi] / s;
                dL_dt0 = gq + dL_dt0;
                float m = logf(T / cq) * gq;
                dL_ds = dL_ds - m / s;
                cdg = cdg - gq;
                qi++;
                if (qi < Q)
                    cq = dq[qi];
            }
            float dL_ddelta = dL_dalpha * ((delta > 0.0f) ? s * oma : 0.0f);
            if (qi < Q) {
                dL_ds = fmaf(cdg, -delta, dL_ds);
                dL_ddelta = fmaf(cdg, -s, dL_ddelta);
            }
            dL_dt0 = dL_dt0 - dL_ddelta;
            float dL_dt1 = dL_ddelta;

            float g_t0_prev[3] = {0, 0, 0}, g_t1_cur[3], g_t0_cur[3], g_t1_next[3];
            if (prev_idx != RF_NONE)
                cell_intersection_grad(prev_point, P, o, d, g_t0_prev);
            cell_intersection_grad(P, Pn, o, d, g_t1_cur);
            cell_intersection_grad(P, prev_point, o, d, g_t0_cur);
            cell_intersection_grad(Pn, P, o, d, g_t1_next);
            for (int c = 0; c < 3; ++c) {
                prev_grad[c] = fmaf(dL_dt0, g_t0_prev[c], prev_grad[c]);
                cur_grad[c] = cur_grad[c] + fmaf(dL_dt0, g_t0_cur[c], dL_dt1 * g_t1_cur[c]);
                next_grad[c] = dL_dt1 * g_t1_next[c];
            }
            if (prev_idx != RF_NONE)
                for (int c = 0; c < 3; ++c)
                    atomic_addf(points_grad + 3 * (size_t)prev_idx + c, prev_grad[c], parallel);
            for (int c = 0; c < 3; ++c) {
                prev_point[c] = P[c];
                prev_grad[c] = cur_grad[c];
                cur_grad[c] = next_grad[c];
                next_grad[c] = 0.0f;
            }
            prev_idx = cur;
            T = Tn;

            float dL_drgb[3];
            for (int c = 0; c < 3; ++c)
                dL_drgb[c] = (rgb[c] == 0.0f) ? 0.0f : rgba_grad[c] * w;
            /* write_rgb_grad_to_sh, src/tracing/sh_utils.cuh:85-92 */
            for (int i = 0; i < 3 * sh_dim; ++i)
                add_attr(attr_grad, is_half, cur * A + i, sh[i / 3] * dL_drgb[i % 3], parallel);
            add_attr(attr_grad, is_half, cur * A + (A - 1), dL_ds, parallel);

            if (!(T > weight_threshold))
                break;
        }
        t0 = fmaxf(t0, t1);
        cur = nxt;
        P[0] = Pn[0];
        P[1] = Pn[1];
        P[2] = Pn[2];
    }
}

/* Pipeline::trace_backward, src/tracing/pipeline.cu:645-701.  points_grad,
 * attribute_grad and point_error must be zero-initialised by the caller
 * (torch_bindings/pipeline_bindings.cpp:441-455 does it upstream).  ray_grad is
 * never written. */
int rfo_trace_backward(int sh_degree, int attr_is_half, float weight_threshold,
                       uint32_t max_intersections, uint32_t num_points,
                       const float *points, const void *attributes,
                       uint32_t point_adjacency_size, const uint32_t *adj,
                       const uint32_t *off, uint32_t num_rays, const float *rays,
                       const uint32_t *start_point_index, uint32_t Q,
                       const float *depth_quantiles,
                       const uint32_t *quantile_point_indices, const void *ray_rgba,
                       const void *ray_rgba_grad, const float *depth_grad,
                       const void *ray_error, float *points_grad,
                       void *attribute_grad, void *point_error, int num_threads) {
    if (sh_degree < 0 || sh_degree > 3)
        return 1;
    half_t *diff = (half_t *)malloc(sizeof(half_t) * 4 * ((size_t)point_adjacency_size + 32));
    if (!diff)
        return 2;
    rfo_prefetch_adjacent_diff(points, num_points, adj, off, (uint16_t *)diff);
    int parallel = num_threads != 1;
#ifdef _OPENMP
    if (num_threads > 0)
        omp_set_num_threads(num_threads);
#endif
#pragma omp parallel for schedule(dynamic, 256) if (parallel)
    for (int64_t r = 0; r < (int64_t)num_rays; ++r) {
        float rgba[4], g[4], err = 0.0f;
        for (int c = 0; c < 4; ++c) {
            rgba[c] = attr_at(ray_rgba, attr_is_half, 4 * (size_t)r + c);
            g[c] = attr_at(ray_rgba_grad, attr_is_half, 4 * (size_t)r + c);
        }
        if (ray_error)
            err = attr_at(ray_error, attr_is_half, (size_t)r);
        backward_ray(sh_degree, attr_is_half, weight_threshold, max_intersections, points,
                     attributes, adj, off, diff, rays + 6 * (size_t)r, start_point_index[r],
                     depth_quantiles ? Q : 0, depth_quantiles ? depth_quantiles + (size_t)r * Q : NULL,
                     depth_quantiles ? quantile_point_indices + (size_t)r * Q : NULL, rgba, g,
                     depth_quantiles ? depth_grad + (size_t)r * Q : NULL, ray_error ? &err : NULL,
                     points_grad, attribute_grad, point_error, parallel);
    }
    free(diff);
    return 0;
}

/* cast_ray, src/tracing/camera.h:56-85 (model 0 = pinhole, 1 = fisheye). */
static void cast_ray(const float pos[3], const float fwd[3], const float right[3],
                     const float up[3], float fov, uint32_t width, uint32_t height,
                     int model, int i, int j, float o[3], float d[3]) {
    float aspect = (float)width / (float)height;
    float x = (float)i / (float)width;
    float y = (float)j / (float)height;
    float u = fmaf(2.0f, x, -1.0f) * aspect;
    float v = fmaf(-2.0f, y, 1.0f);
    float mask = 1.0f;
    for (int c = 0; c < 3; ++c)
        o[c] = pos[c];
    if (model == 0) {
        float w = 1.0f / tanf(fov * 0.5f);
        for (int c = 0; c < 3; ++c)
            d[c] = fmaf(v, up[c], fmaf(u, right[c], w * fwd[c]));
    } else {
        float theta = atan2f(v, u);
        float phi = fov * sqrtf(fmaf(u, u, v * v));
        if (phi >= 3.14159265358979323846f) {
            phi = 3.14159265358979323846f - 1e-6f;
            mask = 0.0f;
        }
        float sp = sinf(phi), cp = cosf(phi), st = sinf(theta), ct = cosf(theta);
        for (int c = 0; c < 3; ++c)
            d[c] = fmaf(cp, fwd[c], fmaf(sp * st, up[c], (sp * ct) * right[c]));
    }
    float n2 = fmaf(d[0], d[0], fmaf(d[1], d[1], d[2] * d[2]));
    if (n2 > 0.0f) {
        float n = sqrtf(n2);
        for (int c = 0; c < 3; ++c)
            d[c] = d[c] / n;
    }
    for (int c = 0; c < 3; ++c)
        d[c] = d[c] * mask;
}

/* make_rgba8, src/tracing/tracing_utils.cuh:105-115 */
static uint32_t make_rgba8(float r, float g, float b, float a) {
    r = fmaxf(0.0f, fminf(1.0f, r));
    g = fmaxf(0.0f, fminf(1.0f, g));
    b = fmaxf(0.0f, fminf(1.0f, b));
    a = fmaxf(0.0f, fminf(1.0f, a));
    int ri = (int)(r * 255.0f), gi = (int)(g * 255.0f), bi = (int)(b * 255.0f), ai = (int)(a * 255.0f);
    return ((uint32_t)ai << 24) | ((uint32_t)bi << 16) | ((uint32_t)gi << 8) | (uint32_t)ri;
}

/* benchmark kernel + Pipeline::trace_benchmark, src/tracing/pipeline.cu:472-544, 738-765.
 * adjacent_diff is supplied by the caller (benchmark.py:41-54), E x 4 halfs.
 * Note the kernel does NOT renormalise the direction again (cast_ray did). */
int rfo_trace_benchmark(int sh_degree, int attr_is_half, float weight_threshold,
                        uint32_t max_intersections, const float *points,
                        const void *attributes, const uint32_t *adj,
                        const uint32_t *off, const uint16_t *adjacent_diff,
                        const float *cam_position, const float *cam_forward,
                        const float *cam_right, const float *cam_up, float fov,
                        uint32_t width, uint32_t height, int model,
                        uint32_t start_point, uint32_t *output_rgba, int num_threads) {
    if (sh_degree < 0 || sh_degree > 3)
        return 1;
    int sh_dim = (sh_degree + 1) * (sh_degree + 1);
    const half_t *diff = (const half_t *)adjacent_diff;
    int parallel = num_threads != 1;
#ifdef _OPENMP
    if (num_threads > 0)
        omp_set_num_threads(num_threads);
#endif
#pragma omp parallel for schedule(dynamic, 256) if (parallel)
    for (int64_t idx = 0; idx < (int64_t)width * height; ++idx) {
        int pi = (int)(idx % width), pj = (int)(idx / width);
        float o[3], d[3];
        cast_ray(cam_position, cam_forward, cam_right, cam_up, fov, width, height, model, pi, pj, o, d);
        float nrm = sqrtf(fmaf(d[0], d[0], fmaf(d[1], d[1], d[2] * d[2])));
        if (nrm < 0.1f) {
            output_rgba[idx] = 0;
            continue;
        }
        float sh[16];
        sh_coefficients(sh_degree, d, sh);
        float T = 1.0f, acc[3] = {0, 0, 0};
        float t0 = 0.0f;
        uint32_t n = 0, cur = start_point;
        float P[3] = {points[3 * (size_t)cur], points[3 * (size_t)cur + 1], points[3 * (size_t)cur + 2]};
        for (;;) {
            n++;
            if (n > max_intersections)
                break;
            uint32_t b = off[cur], nf = off[cur + 1] - b;
            float t1;
            uint32_t face = scan_faces(diff, b, nf, P, o, d, &t1);
            if (face == RF_NONE)
                break;
            uint32_t nxt = adj[b + face];
            if (t1 > t0) {
                float rgb[3], s;
                load_attributes(sh_dim, sh, attributes, attr_is_half, cur, rgb, &s);
                float delta = fmaxf(t1 - t0, 0.0f);
                float alpha = 1.0f - expf(-s * delta);
                float w = T * alpha;
                for (int c = 0; c < 3; ++c)
                    acc[c] = fmaf(w, rgb[c], acc[c]);
                T = T * (1.0f - alpha);
                if (!(T > weight_threshold))
                    break;
            }
            t0 = fmaxf(t0, t1);
            cur = nxt;
            P[0] = points[3 * (size_t)nxt];
            P[1] = points[3 * (size_t)nxt + 1];
            P[2] = points[3 * (size_t)nxt + 2];
        }
        output_rgba[idx] = make_rgba8(acc[0], acc[1], acc[2], 1.0f);
    }
    return 0;
}

/* farthest_neighbor_kernel, src/delaunay/triangulation_ops.cu:9-44.
 * Row order, strict '>' against a running maximum that starts at 0 (so a neighbour at distance 0 or NaN never
 * wins and an empty row leaves UINT32_MAX); `sum_distance += 0.5 * distance` is a DOUBLE add rounded back to
 * float every iteration; radius = sum / (float)num_faces (0/0 = NaN for an empty row).  |q - p| as the SASS
 * has it: sqrt(fma(dx,dx, fma(dy,dy, dz*dz))), IEEE sqrt. */
void rfo_farthest_neighbor(const float *points, uint32_t num_points, const uint32_t *adj,
                           const uint32_t *off, uint32_t *indices, float *cell_radius) {
    for (uint32_t i = 0; i < num_points; ++i) {
        const float *p = points + 3 * (size_t)i;
        const uint32_t begin = off[i], num_faces = off[i + 1] - begin;
        uint32_t farthest_idx = RF_NONE;
        float sum_distance = 0.0f, max_distance = 0.0f;
        for (uint32_t f = 0; f < num_faces; ++f) {
            const uint32_t j = adj[begin + f];
            const float *q = points + 3 * (size_t)j;
            const float dx = q[0] - p[0], dy = q[1] - p[1], dz = q[2] - p[2];
            const float distance = sqrtf(fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
            sum_distance = (float)((double)sum_distance + 0.5 * (double)distance);
            if (distance > max_distance) {
                max_distance = distance;
                farthest_idx = j;
            }
        }
        indices[i] = farthest_idx;
        cell_radius[i] = sum_distance / (float)num_faces;
    }
}

int rfo_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
