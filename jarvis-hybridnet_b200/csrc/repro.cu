// Stage 1: reprojection gather.  Replaces jarvis/hybridnet/repro_layer.py:40-119 (+ F.pad and /255 of
// jarvis/hybridnet/model.py:65-66,72).
//
// Kernels (all HBM/L2-bound integer + gather work, no tensor cores):
//   coarse_project_kernel  half-resolution grid -> per-camera distorted, clamped pixel coordinates
//                          (repro_layer.py:46-68).  Separately rounded fp32 ops in the reference's order;
//                          the K=4 dot product is cuBLAS/MKL's FMA chain.
//   fine_index_kernel      ATen upsample_trilinear3d (align_corners=False, scale 1/2) of both coordinate
//                          volumes, /2, truncate, y*hs+x (repro_layer.py:70-83).  One thread owns the
//                          <=2x2x2 fine voxels that share the same 8 coarse corners, so the corners are
//                          loaded once and the separable lerps are shared (bit-identical to ATen's nested
//                          expression because every intermediate is an fp32 value in both).
//   relayout_kernel        [ncam][K][S][S] planar fp32 -> channels-last [ncam][hs][hs][KP] (fp32 or bf16)
//                          with the 1-px zero border of F.pad materialised, so that one voxel x camera
//                          gather is a single contiguous KP-vector instead of K strided scalars.
//   gather_mean_kernel     index_select + mean over cameras (+ /255): one thread per fine voxel,
//                          vector loads of the KP-vector per camera, cameras accumulated in order.
#include "common.cuh"

namespace jhn {

__device__ __forceinline__ float lerp_rn(float w0, float a, float w1, float b, int mode)
{
    if (mode == JHN_LERP_FMA_FIRST) return __fmaf_rn(w0, a, __fmul_rn(w1, b));
    if (mode == JHN_LERP_FMA_SECOND) return __fmaf_rn(w1, b, __fmul_rn(w0, a));
    return __fadd_rn(__fmul_rn(w0, a), __fmul_rn(w1, b));
}

__global__ void __launch_bounds__(256)
coarse_project_kernel(const float *__restrict__ cam, const float *__restrict__ intr,
                      const float *__restrict__ dist, const int32_t *__restrict__ center3D,
                      const int32_t *__restrict__ centerHM, int B, int ncam, int h, float spacing, int hs,
                      float *__restrict__ ca, float *__restrict__ cb)
{
    const long long total = (long long)B * ncam * h * h * h;
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int k = (int)(t % h); long long r = t / h;
    const int j = (int)(r % h); r /= h;
    const int i = (int)(r % h); r /= h;
    const int c = (int)(r % ncam);
    const int b = (int)(r / ncam);
    const int bc = b * ncam + c;
    const float *P = cam + 12 * bc;
    const float fx = intr[9 * bc + 0], fy = intr[9 * bc + 4];
    const float cx = intr[9 * bc + 6], cy = intr[9 * bc + 7];
    const float k1 = dist[5 * bc + 0], k2 = dist[5 * bc + 1];
    const int chxi = centerHM[2 * bc + 0], chyi = centerHM[2 * bc + 1];
    const float chx = (float)chxi, chy = (float)chyi, fhs = (float)hs;
    const float lox = (float)(chxi - (hs - 1)), hix = (float)(chxi + hs - 2);
    const float loy = (float)(chyi - (hs - 1)), hiy = (float)(chyi + hs - 2);
    const int half = h / 2;
    // grid = (idx - half) * spacing * 2 + center        repro_layer.py:32-36,113
    const float X = __fadd_rn(__fmul_rn(__fmul_rn((float)(i - half), spacing), 2.f), (float)center3D[3 * b + 0]);
    const float Y = __fadd_rn(__fmul_rn(__fmul_rn((float)(j - half), spacing), 2.f), (float)center3D[3 * b + 1]);
    const float Z = __fadd_rn(__fmul_rn(__fmul_rn((float)(k - half), spacing), 2.f), (float)center3D[3 * b + 2]);
    float uvw[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) {                      // [X,Y,Z,1] @ P, FMA chain     :46-52
        float s = __fmul_rn(X, P[q]);
        s = __fmaf_rn(Y, P[3 + q], s);
        s = __fmaf_rn(Z, P[6 + q], s);
        s = __fmaf_rn(1.f, P[9 + q], s);
        uvw[q] = s;
    }
    float a = __fsub_rn(__fdiv_rn(uvw[0], uvw[2]), cx);                                   // :54-55
    float bb = __fsub_rn(__fdiv_rn(uvw[1], uvw[2]), cy);                                  // :56-57
    float ax = __fdiv_rn(a, fx); ax = __fmul_rn(ax, ax);                                  // :58
    float by = __fdiv_rn(bb, fy); by = __fmul_rn(by, by);                                 // :59
    const float r2 = __fadd_rn(ax, by);
    const float d = __fadd_rn(1.f, __fmul_rn(__fadd_rn(k1, __fmul_rn(k2, r2)), r2));      // :60-61
    a = __fadd_rn(__fmul_rn(a, d), cx);                                                   // :62
    bb = __fadd_rn(__fmul_rn(bb, d), cy);                                                 // :63
    a = fminf(fmaxf(a, lox), hix);                                                        // :65-66
    a = __fsub_rn(__fadd_rn(__fsub_rn(a, chx), fhs), 1.f);
    bb = fminf(fmaxf(bb, loy), hiy);                                                      // :67-68
    bb = __fsub_rn(__fadd_rn(__fsub_rn(bb, chy), fhs), 1.f);
    ca[t] = a;
    cb[t] = bb;
}

// Per-dimension description of the fine voxels owned by corner block m in [0, h]:
//   m == 0 : I = {0}         i0 = 0,   lambda1 = 0
//   m == h : I = {G-1}       i0 = h-1, lambda1 = .25
//   else   : I = {2m-1, 2m}  i0 = m-1, lambda1 = {.25, .75}
// (ATen area_pixel_compute_source_index with scale .5: src = .5*(I+.5)-.5 clamped at 0.)
struct DimBlock { int i0, i1, first, count; float l1[2]; };
__device__ __forceinline__ DimBlock dim_block(int m, int h)
{
    DimBlock d;
    if (m == 0) { d.i0 = 0; d.first = 0; d.count = 1; d.l1[0] = 0.f; d.l1[1] = 0.f; }
    else if (m == h) { d.i0 = h - 1; d.first = 2 * h - 1; d.count = 1; d.l1[0] = 0.25f; d.l1[1] = 0.25f; }
    else { d.i0 = m - 1; d.first = 2 * m - 1; d.count = 2; d.l1[0] = 0.25f; d.l1[1] = 0.75f; }
    d.i1 = d.i0 + (d.i0 < h - 1 ? 1 : 0);
    return d;
}

__global__ void __launch_bounds__(128)
fine_index_kernel(const float *__restrict__ ca, const float *__restrict__ cb, int BC, int h, int hs,
                  int lerp_mode, int32_t *__restrict__ idx)
{
    const int hb = h + 1, G = 2 * h;
    const long long total = (long long)BC * hb * hb * hb;
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int mk = (int)(t % hb); long long r = t / hb;
    const int mj = (int)(r % hb); r /= hb;
    const int mi = (int)(r % hb);
    const int bc = (int)(r / hb);
    const DimBlock di = dim_block(mi, h), dj = dim_block(mj, h), dk = dim_block(mk, h);
    const size_t cbase = (size_t)bc * h * h * h;
    float va[2][2][2], vb[2][2][2];
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const size_t o = cbase + ((size_t)(p ? di.i1 : di.i0) * h + (q ? dj.i1 : dj.i0)) * h + (s ? dk.i1 : dk.i0);
                va[p][q][s] = __ldg(ca + o);
                vb[p][q][s] = __ldg(cb + o);
            }
    int32_t *out = idx + (size_t)bc * G * G * G;
    for (int kv = 0; kv < dk.count; ++kv) {
        const float lk1 = dk.l1[kv], lk0 = __fsub_rn(1.f, lk1);
        float xa[2][2], xb[2][2];
#pragma unroll
        for (int p = 0; p < 2; ++p)
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                xa[p][q] = lerp_rn(lk0, va[p][q][0], lk1, va[p][q][1], lerp_mode);
                xb[p][q] = lerp_rn(lk0, vb[p][q][0], lk1, vb[p][q][1], lerp_mode);
            }
        for (int jv = 0; jv < dj.count; ++jv) {
            const float lj1 = dj.l1[jv], lj0 = __fsub_rn(1.f, lj1);
            float ya[2], yb[2];
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                ya[p] = lerp_rn(lj0, xa[p][0], lj1, xa[p][1], lerp_mode);
                yb[p] = lerp_rn(lj0, xb[p][0], lj1, xb[p][1], lerp_mode);
            }
            for (int iv = 0; iv < di.count; ++iv) {
                const float li1 = di.l1[iv], li0 = __fsub_rn(1.f, li1);
                const float fa = lerp_rn(li0, ya[0], li1, ya[1], lerp_mode);
                const float fb = lerp_rn(li0, yb[0], li1, yb[1], lerp_mode);
                const int ix = __float2int_rz(__fmul_rn(fa, 0.5f));      // (val1/2).int()   :82-83
                const int iy = __float2int_rz(__fmul_rn(fb, 0.5f));
                out[((size_t)(di.first + iv) * G + (dj.first + jv)) * G + (dk.first + kv)] = iy * hs + ix;
            }
        }
    }
}

template <typename T> __device__ __forceinline__ T to_store(float v);
template <> __device__ __forceinline__ float to_store<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 to_store<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// one block per padded image row (b, cam, y): planar -> channels-last through shared memory
template <typename T>
__global__ void __launch_bounds__(128)
relayout_kernel(const float *__restrict__ in, int K, int hs, int padded, T *__restrict__ out)
{
    extern __shared__ float tile[];                     // [hs][KP+1]
    const int y = blockIdx.x % hs;
    const int bc = blockIdx.x / hs;
    const int S = padded ? hs : hs - 2;
    const int off = padded ? 0 : 1;
    const int ys = y - off;
    const bool row_ok = ys >= 0 && ys < S;
    for (int k = 0; k < K; ++k) {
        const float *src = in + (((size_t)bc * K + k) * S + (row_ok ? ys : 0)) * S;
        for (int x = threadIdx.x; x < hs; x += blockDim.x) {
            const int xs = x - off;
            float v = 0.f;
            if (row_ok && xs >= 0 && xs < S) v = __ldg(src + xs);
            tile[x * (KP + 1) + k] = v;
        }
    }
    __syncthreads();
    T *dst = out + ((size_t)bc * hs + y) * hs * KP;
    for (int e = threadIdx.x; e < hs * KP; e += blockDim.x) {
        const int x = e / KP, k = e - x * KP;
        dst[e] = to_store<T>(k < K ? tile[x * (KP + 1) + k] : 0.f);
    }
}

template <typename T> struct Vec;
template <> struct Vec<float> {
    static constexpr int N = KP / 4;                    // 6 x float4
    static __device__ __forceinline__ void add(const float *p, float *acc)
    {
        const float4 *q = reinterpret_cast<const float4 *>(p);
        float4 v[N];
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = __ldg(q + i);
#pragma unroll
        for (int i = 0; i < N; ++i) {
            acc[4 * i + 0] = __fadd_rn(acc[4 * i + 0], v[i].x);
            acc[4 * i + 1] = __fadd_rn(acc[4 * i + 1], v[i].y);
            acc[4 * i + 2] = __fadd_rn(acc[4 * i + 2], v[i].z);
            acc[4 * i + 3] = __fadd_rn(acc[4 * i + 3], v[i].w);
        }
    }
};
template <> struct Vec<__nv_bfloat16> {
    static constexpr int N = KP / 8;                    // 3 x 16 B
    static __device__ __forceinline__ void add(const __nv_bfloat16 *p, float *acc)
    {
        const uint4 *q = reinterpret_cast<const uint4 *>(p);
        uint4 v[N];
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = __ldg(q + i);
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const uint32_t w[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {               // bf16 -> fp32 is a 16-bit shift
                acc[8 * i + 2 * j + 0] += __uint_as_float(w[j] << 16);
                acc[8 * i + 2 * j + 1] += __uint_as_float(w[j] & 0xffff0000u);
            }
        }
    }
};

template <typename T, int LAYOUT>
__global__ void __launch_bounds__(256)
gather_mean_kernel(const T *__restrict__ hm, const int32_t *__restrict__ idx, int ncam, int K, int hs, int G,
                   float post_divide, void *__restrict__ out_)
{
    const size_t nv = (size_t)G * G * G;
    const size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (v >= nv) return;
    float acc[KP];
#pragma unroll
    for (int i = 0; i < KP; ++i) acc[i] = 0.f;
    const int32_t *ip = idx + (size_t)b * ncam * nv + v;
    const T *base = hm + (size_t)b * ncam * hs * hs * KP;
    for (int c = 0; c < ncam; ++c) {
        const int32_t flat = __ldg(ip + (size_t)c * nv);
        Vec<T>::add(base + ((size_t)c * hs * hs + flat) * KP, acc);
    }
    const float fn = (float)ncam;
#pragma unroll
    for (int k = 0; k < KP; ++k) {
        float m = __fdiv_rn(acc[k], fn);                            // torch.mean        :103-105
        if (post_divide != 1.f) m = __fdiv_rn(m, post_divide);      // heatmaps3D/255.   model.py:72
        acc[k] = m;
    }
    if (LAYOUT == JHN_VOL_NCDHW_F32) {
        float *out = (float *)out_ + (size_t)b * K * nv + v;
#pragma unroll
        for (int k = 0; k < KP; ++k)
            if (k < K) out[(size_t)k * nv] = acc[k];
    } else {
        // parity-split, channel-blocked bf16 volume read by the tensor-core front convolution (conv_tc.cu):
        // [b][s = (I&1,J&1,Kz&1)][j][zp][pp][8] on the G/2 grid; channel chunks beyond K are written as zeros
        const int I = (int)(v / ((size_t)G * G)), r = (int)(v - (size_t)I * G * G), J = r / G, Kz = r - J * G;
        const int CJ = (K + 15) / 16 * 2, Wh = G / 2 + 2;
        const int s = ((I & 1) * 2 + (J & 1)) * 2 + (Kz & 1);
        uint4 *out = (uint4 *)out_;
        const size_t pos = ((size_t)(I >> 1) + 1) * Wh * Wh + (size_t)((J >> 1) + 1) * Wh + (Kz >> 1) + 1;
        const size_t chunk_stride = (size_t)Wh * Wh * Wh;
        const size_t base = (((size_t)b * 8 + s) * CJ) * chunk_stride + pos;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (j < CJ) {
                uint32_t pk[4] = {0, 0, 0, 0};
                if (j < KP / 8) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        __nv_bfloat162 h2 = __floats2bfloat162_rn(acc[8 * j + 2 * i], acc[8 * j + 2 * i + 1]);
                        pk[i] = *reinterpret_cast<uint32_t *>(&h2);
                    }
                }
                out[base + (size_t)j * chunk_stride] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
        }
    }
}

size_t reproject_workspace(int B, int ncam, int K, int hs, int G, int precision)
{
    const int h = G / 2;
    Arena a(nullptr, 0);
    a.take<float>((size_t)B * ncam * h * h * h);                    // coarse a
    a.take<float>((size_t)B * ncam * h * h * h);                    // coarse b
    a.take<int32_t>((size_t)B * ncam * G * G * G);                  // fine indices
    const size_t px = (size_t)B * ncam * hs * hs * KP;
    if (precision == JHN_FP32) a.take<float>(px); else a.take<__nv_bfloat16>(px);
    return a.off;
}

template <typename T>
static int run_gather(const ReprojectArgs &a, const T *hm_cl, const int32_t *idx, cudaStream_t st)
{
    const size_t nv = (size_t)a.G * a.G * a.G;
    dim3 grid(cdiv(nv, 256), a.B);
    if (a.layout == JHN_VOL_NCDHW_F32) {
        JHN_LAUNCH("gather_mean_kernel", st,
                   (gather_mean_kernel<T, JHN_VOL_NCDHW_F32><<<grid, 256, 0, st>>>(hm_cl, idx, a.ncam, a.K, a.hs, a.G,
                                                                                  a.post_divide, a.volume_out)));
        return JHN_OK;
    }
    const int CJ = (a.K + 15) / 16 * 2;
    JHN_TRY(tc_zero_border_launch(a.volume_out, a.B * 8 * CJ, a.G / 2, st));
    JHN_LAUNCH("gather_mean_kernel", st,
               (gather_mean_kernel<T, JHN_VOL_V2V_BF16><<<grid, 256, 0, st>>>(hm_cl, idx, a.ncam, a.K, a.hs, a.G,
                                                                             a.post_divide, a.volume_out)));
    return JHN_OK;
}

int reproject_launch(const ReprojectArgs &a, void *ws, size_t ws_bytes, cudaStream_t st)
{
    const int h = a.G / 2;
    Arena ar(ws, ws_bytes);
    float *ca = ar.take<float>((size_t)a.B * a.ncam * h * h * h);
    float *cb = ar.take<float>((size_t)a.B * a.ncam * h * h * h);
    int32_t *idx_ws = ar.take<int32_t>((size_t)a.B * a.ncam * a.G * a.G * a.G);
    const size_t px = (size_t)a.B * a.ncam * a.hs * a.hs * KP;
    void *hm_cl = (a.precision == JHN_FP32) ? (void *)ar.take<float>(px) : (void *)ar.take<__nv_bfloat16>(px);
    if (!ar.ok()) return fail(JHN_ERR_WORKSPACE, "reproject workspace: need %zu bytes, got %zu", ar.off, ws_bytes);
    int32_t *idx = a.index_out ? a.index_out : idx_ws;

    const long long nc = (long long)a.B * a.ncam * h * h * h;
    JHN_LAUNCH("coarse_project_kernel", st,
               coarse_project_kernel<<<cdiv(nc, 256), 256, 0, st>>>(a.cam, a.intr, a.dist, a.center3D, a.centerHM, a.B,
                                                                   a.ncam, h, a.spacing, a.hs, ca, cb));
    const long long nb = (long long)a.B * a.ncam * (h + 1) * (h + 1) * (h + 1);
    JHN_LAUNCH("fine_index_kernel", st,
               fine_index_kernel<<<cdiv(nb, 128), 128, 0, st>>>(ca, cb, a.B * a.ncam, h, a.hs, a.lerp_mode, idx));
    const size_t smem = (size_t)a.hs * (KP + 1) * sizeof(float);
    const int rows = a.B * a.ncam * a.hs;
    if (a.precision == JHN_FP32) {
        JHN_LAUNCH("relayout_kernel", st,
                   relayout_kernel<float><<<rows, 128, smem, st>>>(a.heatmaps, a.K, a.hs, a.padded, (float *)hm_cl));
        return run_gather<float>(a, (const float *)hm_cl, idx, st);
    }
    JHN_LAUNCH("relayout_kernel", st,
               relayout_kernel<__nv_bfloat16><<<rows, 128, smem, st>>>(a.heatmaps, a.K, a.hs, a.padded, (__nv_bfloat16 *)hm_cl));
    return run_gather<__nv_bfloat16>(a, (const __nv_bfloat16 *)hm_cl, idx, st);
}

}  // namespace jhn
