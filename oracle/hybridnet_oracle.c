/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
 *
 * Plain-C CPU restatement of the reference's 3D hot path stages that are index / gather / reduction
 * work (the V2V convolutions are restated with torch-CPU functional ops in hybridnet_oracle.py):
 *
 *   jho_reproject_indices   jarvis/hybridnet/repro_layer.py:26-36 (grid), :40-85 (reprojectPoints)
 *   jho_gather_mean         jarvis/hybridnet/repro_layer.py:88-107 (_get_heatmap_value)
 *   jho_centroid            jarvis/hybridnet/model.py:72-87 (softplus / centroid / confidence)
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library.  Parity is PINNED: tests/test_oracle_golden.py checks every function against
 * fixtures produced by importing the reference itself (tests/golden/make_golden.py).
 *
 * Arithmetic contract: every reference line is its own torch op, i.e. a separately
 * rounded fp32 operation, so this file is compiled with -ffp-contract=off and writes each op on its
 * own statement.  Two sub-steps are library-defined in the reference and were matched empirically
 * against torch 2.11 CPU (tests/golden/make_golden.py records the check):
 *   - the K=4 dot product of torch.matmul (repro_layer.py:50): an FMA chain in k order starting
 *     from the rounded first product  (r = x0*p0; r = fma(x1,p1,r); ...);
 *   - ATen upsample_trilinear3d (repro_layer.py:70-80): nested lerps, innermost dimension first,
 *     each lerp evaluated as fma(w0, a, w1*b).
 * `lerp_mode` selects the lerp contraction (0: fma(w0,a,w1*b) = ATen CPU; 1: fma(w1,b,w0*a);
 * 2: no fma) so the GPU tests can pin whichever one ATen's CUDA kernel uses on the B200.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline float lerp_f(float w0, float a, float w1, float b, int mode)
{
    if (mode == 0) { float t = w1 * b; return fmaf(w0, a, t); }
    if (mode == 1) { float t = w0 * a; return fmaf(w1, b, t); }
    { float t = w0 * a; float u = w1 * b; return t + u; }
}

/* source index / weights of ATen's area_pixel_compute_source_index at scale 0.5, align_corners=False
 * (torch/include/ATen/native/cuda/UpSample.cuh:115-130; CPU twin in UpSampleKernel.cpp). */
static inline void src_index(int dst, int h, int *i0, int *i1, float *l0, float *l1)
{
    float s = 0.5f * ((float)dst + 0.5f) - 0.5f;
    if (s < 0.f) s = 0.f;
    int a = (int)s;
    *i0 = a;
    *i1 = a + ((a < h - 1) ? 1 : 0);
    *l1 = s - (float)a;
    *l0 = 1.f - *l1;
}

/* Coarse projection: a,b of SURVEY.md §9.1 for every coarse grid point and camera.
 * coarse_a / coarse_b : [ncam][h][h][h] fp32 (x fastest = k index, as the reference's view). */
static void project_coarse(const float *cam, const float *intr, const float *dist,
                           const float *center3D, const int32_t *centerHM,
                           int c_begin, int c_end, int h, float spacing, int hs,
                           float *coarse_a, float *coarse_b)
{
    const int half = h / 2;                               /* int(grid_size/2/2), repro_layer.py:28 */
    /* `self.grid + center[0]` (repro_layer.py:113): an int centre (jarvis3D.py:183) is promoted to fp32 (exact), a
     * float centre (validation path, hybridnet.py:284-304) is added as it is */
    const float c3x = center3D[0], c3y = center3D[1], c3z = center3D[2];
    for (int c = c_begin; c < c_end; ++c) {
        const float *P = cam + 12 * c;                    /* [4][3] row-major                      */
        const float fx = intr[9 * c + 0], fy = intr[9 * c + 4];
        const float cx = intr[9 * c + 6], cy = intr[9 * c + 7];
        const float k1 = dist[5 * c + 0], k2 = dist[5 * c + 1];
        const float chx = (float)centerHM[2 * c + 0], chy = (float)centerHM[2 * c + 1];
        /* clamp bounds are integer tensors promoted to fp32 (repro_layer.py:65-68) */
        const float lox = (float)(centerHM[2 * c + 0] - (hs - 1)), hix = (float)(centerHM[2 * c + 0] + hs - 2);
        const float loy = (float)(centerHM[2 * c + 1] - (hs - 1)), hiy = (float)(centerHM[2 * c + 1] + hs - 2);
        const float fhs = (float)hs;
        for (int i = 0; i < h; ++i)
        for (int j = 0; j < h; ++j)
        for (int k = 0; k < h; ++k) {
            /* grid = (idx - half) * spacing * 2 + center   (repro_layer.py:32-36,113) */
            float X = (float)(i - half) * spacing; X = X * 2.f; X = X + c3x;
            float Y = (float)(j - half) * spacing; Y = Y * 2.f; Y = Y + c3y;
            float Z = (float)(k - half) * spacing; Z = Z * 2.f; Z = Z + c3z;
            /* [X,Y,Z,1] @ P : SGEMM K=4 as an FMA chain (repro_layer.py:46-52) */
            float uvw[3];
            for (int q = 0; q < 3; ++q) {
                float r = X * P[q];
                r = fmaf(Y, P[3 + q], r);
                r = fmaf(Z, P[6 + q], r);
                r = fmaf(1.f, P[9 + q], r);
                uvw[q] = r;
            }
            float a = uvw[0] / uvw[2]; a = a - cx;                        /* :54-55 */
            float b = uvw[1] / uvw[2]; b = b - cy;                        /* :56-57 */
            float ax = a / fx; ax = ax * ax;                              /* :58    */
            float by = b / fy; by = by * by;                              /* :59    */
            float r2 = ax + by;
            float d = k2 * r2; d = k1 + d; d = d * r2; d = 1.f + d;       /* :60-61 */
            a = a * d; a = a + cx;                                        /* :62    */
            b = b * d; b = b + cy;                                        /* :63    */
            a = fminf(fmaxf(a, lox), hix); a = a - chx; a = a + fhs; a = a - 1.f;   /* :65-66 */
            b = fminf(fmaxf(b, loy), hiy); b = b - chy; b = b + fhs; b = b - 1.f;   /* :67-68 */
            size_t o = (((size_t)c * h + i) * h + j) * h + k;
            coarse_a[o] = a;
            coarse_b[o] = b;
        }
    }
}

/* res of repro_layer.py:82-83 as int32 flat padded-pixel index y*hs+x, layout [ncam][G][G][G].
 * Optional outputs: coarse_a/coarse_b [ncam][h^3] (pre-interpolation coordinates).
 * Only cameras [c_begin,c_end) are written, so the Python wrapper can split cameras over threads. */
int jho_reproject_indices(const float *cam, const float *intr, const float *dist,
                          const float *center3D, const int32_t *centerHM,
                          int ncam, int G, float spacing, int hs, int lerp_mode,
                          int32_t *idx_out, float *coarse_a_out, float *coarse_b_out,
                          int c_begin, int c_end)
{
    const int h = G / 2;
    if (G <= 0 || (G & 1) || ncam <= 0 || hs < 3) return -1;
    size_t nc = (size_t)ncam * h * h * h;
    float *ca = coarse_a_out ? coarse_a_out : (float *)malloc(nc * sizeof(float));
    float *cb = coarse_b_out ? coarse_b_out : (float *)malloc(nc * sizeof(float));
    if (!ca || !cb) return -2;
    project_coarse(cam, intr, dist, center3D, centerHM, c_begin, c_end, h, spacing, hs, ca, cb);

    for (int c = c_begin; c < c_end; ++c)
    for (int I = 0; I < G; ++I) {
        const float *A = ca + (size_t)c * h * h * h;
        const float *B = cb + (size_t)c * h * h * h;
        int i0, i1; float li0, li1;
        src_index(I, h, &i0, &i1, &li0, &li1);
        for (int J = 0; J < G; ++J) {
            int j0, j1; float lj0, lj1;
            src_index(J, h, &j0, &j1, &lj0, &lj1);
            for (int K = 0; K < G; ++K) {
                int k0, k1; float lk0, lk1;
                src_index(K, h, &k0, &k1, &lk0, &lk1);
                float v[2];
                for (int w = 0; w < 2; ++w) {
                    const float *S = w ? B : A;
#define AT(a, b, c_) S[((size_t)(a) * h + (b)) * h + (c_)]
                    float x00 = lerp_f(lk0, AT(i0, j0, k0), lk1, AT(i0, j0, k1), lerp_mode);
                    float x01 = lerp_f(lk0, AT(i0, j1, k0), lk1, AT(i0, j1, k1), lerp_mode);
                    float x10 = lerp_f(lk0, AT(i1, j0, k0), lk1, AT(i1, j0, k1), lerp_mode);
                    float x11 = lerp_f(lk0, AT(i1, j1, k0), lk1, AT(i1, j1, k1), lerp_mode);
#undef AT
                    float y0 = lerp_f(lj0, x00, lj1, x01, lerp_mode);
                    float y1 = lerp_f(lj0, x10, lj1, x11, lerp_mode);
                    v[w] = lerp_f(li0, y0, li1, y1, lerp_mode);
                }
                float xa = v[0] / 2.f, yb = v[1] / 2.f;                  /* :82 */
                int32_t ix = (int32_t)xa, iy = (int32_t)yb;              /* .int() truncates */
                idx_out[(((size_t)c * G + I) * G + J) * G + K] = iy * hs + ix;
            }
        }
    }
    if (!coarse_a_out) free(ca);
    if (!coarse_b_out) free(cb);
    return 0;
}

/* index_select + mean over cameras (repro_layer.py:97-105).
 * hm: padded maps [ncam][K][hs][hs]; idx [ncam][G^3]; out [K][G^3] = (sum_c hm[c][k][idx[c][v]]) / ncam,
 * cameras accumulated in order.  Only key points [k_begin,k_end) are written (thread split). */
int jho_gather_mean(const float *hm, const int32_t *idx, int ncam, int K, int hs, int G, float *out,
                    int k_begin, int k_end)
{
    const size_t nv = (size_t)G * G * G, plane = (size_t)hs * hs;
    for (int k = k_begin; k < k_end; ++k)
    for (size_t v0 = 0; v0 < nv; v0 += 4096) {
        size_t v1 = v0 + 4096 < nv ? v0 + 4096 : nv;
        for (size_t v = v0; v < v1; ++v) {
            float s = 0.f;
            for (int c = 0; c < ncam; ++c)
                s = s + hm[((size_t)c * K + k) * plane + (size_t)idx[(size_t)c * nv + v]];
            out[(size_t)k * nv + v] = s / (float)ncam;
        }
    }
    return 0;
}

/* softplus (beta 1, threshold 20) -> sum-normalised centroid, confidence, voxel->mm
 * (model.py:73-87); argmax = first maximum of the raw volume ("argmax voxel").
 * v: [K][h][h][h] fp32 with axes (x=i, y=j, z=k) like the ij meshgrid of model.py:44-48.
 * Sums are accumulated in double so the oracle is the low-noise side of the 0.05 mm comparison. */
int jho_centroid(const float *v, int K, int h, float spacing, float roi, const float *center3D,
                 float *points, float *conf, int32_t *argmax)
{
    const size_t nv = (size_t)h * h * h;
    for (int k = 0; k < K; ++k) {
        const float *p = v + (size_t)k * nv;
        double n = 0, sx = 0, sy = 0, sz = 0;
        float best_raw = -INFINITY, best_hf = -INFINITY; int32_t best = 0;
        for (int i = 0; i < h; ++i)
        for (int j = 0; j < h; ++j)
        for (int q = 0; q < h; ++q) {
            size_t o = ((size_t)i * h + j) * h + q;
            float x = p[o];
            float hf = x > 20.f ? x : log1pf(expf(x));
            n += hf; sx += (double)hf * i; sy += (double)hf * j; sz += (double)hf * q;
            if (x > best_raw) { best_raw = x; best = (int32_t)o; }
            if (hf > best_hf) best_hf = hf;
        }
        float c = best_hf > 255.f ? 255.f : best_hf;                      /* model.py:84-85 */
        conf[k] = c / 255.f;
        argmax[k] = best;
        float cx = (float)(sx / n), cy = (float)(sy / n), cz = (float)(sz / n);
        /* points*spacing*2 - roi/2 + center3D  (model.py:86-87) */
        points[3 * k + 0] = cx * spacing * 2.f - roi / 2.f + center3D[0];
        points[3 * k + 1] = cy * spacing * 2.f - roi / 2.f + center3D[1];
        points[3 * k + 2] = cz * spacing * 2.f - roi / 2.f + center3D[2];
    }
    return 0;
}
