// Stage 3: softplus -> sum-normalised centroid -> confidence -> mm.  Replaces the inline tail of
// HybridNetBackbone.forward (jarvis/hybridnet/model.py:73-87).  One block per (frame set, key point):
// a single pass over the h^3 fp32 volume, fused softplus + four sums + max/argmax, block reduction.
// HBM-bound: K*h^3*4 bytes read once, 5 floats written.
#include "common.cuh"

namespace jhn {

struct TailAcc { float n, sx, sy, sz, hf_max, raw_max; int arg; };

__device__ __forceinline__ void tail_merge(TailAcc &a, const TailAcc &b)
{
    a.n += b.n; a.sx += b.sx; a.sy += b.sy; a.sz += b.sz;
    a.hf_max = fmaxf(a.hf_max, b.hf_max);
    if (b.raw_max > a.raw_max || (b.raw_max == a.raw_max && b.arg < a.arg)) { a.raw_max = b.raw_max; a.arg = b.arg; }
}

__global__ void __launch_bounds__(256)
centroid_kernel(const float *__restrict__ v, int K, int h, float spacing, float roi,
                const float *__restrict__ center3D, float *__restrict__ points, float *__restrict__ conf,
                int32_t *__restrict__ argmax)
{
    const int k = blockIdx.x, b = blockIdx.y;
    const int nv = h * h * h, hh = h * h;
    const float *p = v + ((size_t)b * K + k) * nv;
    TailAcc a = {0.f, 0.f, 0.f, 0.f, -INFINITY, -INFINITY, 0x7fffffff};
    for (int o = threadIdx.x; o < nv; o += blockDim.x) {
        const float x = __ldg(p + o);
        const float hf = x > 20.f ? x : log1pf(expf(x));           // nn.Softplus(beta=1, threshold=20)  :73
        const int i = o / hh, r = o - i * hh, j = r / h, q = r - j * h;
        a.n += hf;                                                   // :76
        a.sx = fmaf(hf, (float)i, a.sx);                             // :77-82
        a.sy = fmaf(hf, (float)j, a.sy);
        a.sz = fmaf(hf, (float)q, a.sz);
        a.hf_max = fmaxf(a.hf_max, hf);                              // :84
        if (x > a.raw_max) { a.raw_max = x; a.arg = o; }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        TailAcc o;
        o.n = __shfl_xor_sync(0xffffffffu, a.n, s); o.sx = __shfl_xor_sync(0xffffffffu, a.sx, s);
        o.sy = __shfl_xor_sync(0xffffffffu, a.sy, s); o.sz = __shfl_xor_sync(0xffffffffu, a.sz, s);
        o.hf_max = __shfl_xor_sync(0xffffffffu, a.hf_max, s); o.raw_max = __shfl_xor_sync(0xffffffffu, a.raw_max, s);
        o.arg = __shfl_xor_sync(0xffffffffu, a.arg, s);
        tail_merge(a, o);
    }
    __shared__ TailAcc part[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) part[warp] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        TailAcc t = part[0];
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) tail_merge(t, part[w]);
        const int o = b * K + k;
        conf[o] = fminf(t.hf_max, 255.f) / 255.f;                    // :84-85
        if (argmax) argmax[o] = t.arg;
        const float sc = spacing * 2.f, half = roi / 2.f;            // :86-87
        points[3 * o + 0] = (t.sx / t.n) * sc - half + center3D[3 * b + 0];
        points[3 * o + 1] = (t.sy / t.n) * sc - half + center3D[3 * b + 1];
        points[3 * o + 2] = (t.sz / t.n) * sc - half + center3D[3 * b + 2];
    }
}

int centroid_launch(const float *v, int B, int K, int h, float spacing, float roi, const float *center3D,
                    float *points, float *conf, int32_t *argmax, cudaStream_t st)
{
    JHN_LAUNCH("centroid_kernel", st,
               centroid_kernel<<<dim3(K, B), 256, 0, st>>>(v, K, h, spacing, roi, center3D, points, conf, argmax));
    return JHN_OK;
}

}  // namespace jhn
