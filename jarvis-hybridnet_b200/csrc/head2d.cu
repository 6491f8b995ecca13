// the last layer of the 2D key-point network written in the reprojection gather's layout.
//
// EfficientTrackBackbone ends in  res2 = deconv1(res1)  (jarvis/efficienttrack/model.py:89-95,127: ConvTranspose2d,
// C -> K channels, kernel 4, stride 2, padding 1, no bias; C = 64 / 88 / 160), and HybridNetBackbone then reshapes and
// F.pad-s that tensor (jarvis/hybridnet/model.py:57-66) before the ReprojectionLayer reads single pixels of it.  The
// gather wants one (camera, pixel) as ONE contiguous 24-channel 16-bit vector with the zero border in place
// (JHN_HM_F16_CL), so this kernel computes the transposed convolution and stores its result directly in that form:
// the planar fp32 tensor (18 MB per frame set), F.pad's copy and the staging pass that re-reads both never exist.
//
// Arithmetic: fp32 FFMA, fp32 accumulation (the reference's fp32 layer; cuDNN may use TF32 unless allow_tf32 is off).
// out[n][k][oy][ox] = sum_c sum_{ky,kx} in[n][c][iy][ix] * w[c][k][ky][kx],  oy = 2 iy - 1 + ky,  ox = 2 ix - 1 + kx:
// every output pixel receives 2 x 2 taps, and the 2 x 2 outputs {2iy, 2iy+1} x {2ix, 2ix+1} read the 3 x 3 input
// neighbourhood of (iy, ix).  CTA = 16 x 16 input positions of one image (32 x 32 outputs), 256 threads; thread =
// two x-adjacent positions x 12 of the 24 output channels (96 accumulators), the channel half is warp-uniform so the
// weights are shared-memory broadcasts (48 LDS.128 for 384 FFMA per input channel).  Input channels stream through
// shared memory in chunks of 8.
#include <cuda_fp16.h>

#include "common.cuh"

namespace jhn {

constexpr int H2_T = 16, H2_CH = 8, H2_ROW = H2_T + 2;

// tap (ky) -> output parity py and neighbourhood row: oy = 2 iy' - 1 + ky
__host__ __device__ constexpr int tap_par(int k) { return (k == 0 || k == 2) ? 1 : 0; }
__host__ __device__ constexpr int tap_off(int k) { return k == 0 ? 2 : (k == 3 ? 0 : 1); }   // row / col index in the 3-wide window

template <int FORMAT>
__global__ void __launch_bounds__(256, 1)
efftrack_head_kernel(const float *__restrict__ in, const float *__restrict__ w, int C, int K, int Hq, int Wq, void *__restrict__ out_)
{
    __shared__ __align__(16) float in_s[H2_CH][H2_ROW][H2_ROW];
    __shared__ __align__(16) float w_s[H2_CH][16][KP];
    const int n = blockIdx.y;
    const int tiles_x = (Wq + H2_T - 1) / H2_T;
    const int y0 = (blockIdx.x / tiles_x) * H2_T, x0 = (blockIdx.x % tiles_x) * H2_T;
    const int half = threadIdx.x >> 7, t = threadIdx.x & 127;
    const int ty = t >> 3, tx2 = (t & 7) * 2;
    float acc[2][2][2][12];                                           // [position][py][px][channel]
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b)
#pragma unroll
                for (int k = 0; k < 12; ++k) acc[p][a][b][k] = 0.f;

    for (int c0 = 0; c0 < C; c0 += H2_CH) {
        __syncthreads();
        for (int e = threadIdx.x; e < H2_CH * H2_ROW * H2_ROW; e += 256) {
            const int cc = e / (H2_ROW * H2_ROW), r = e - cc * H2_ROW * H2_ROW, yy = r / H2_ROW, xx = r - yy * H2_ROW;
            const int iy = y0 + yy - 1, ix = x0 + xx - 1, c = c0 + cc;
            const bool ok = c < C && iy >= 0 && iy < Hq && ix >= 0 && ix < Wq;
            in_s[cc][yy][xx] = ok ? __ldg(in + (((size_t)n * C + c) * Hq + iy) * Wq + ix) : 0.f;
        }
        for (int e = threadIdx.x; e < H2_CH * 16 * KP; e += 256) {
            const int cc = e / (16 * KP), r = e - cc * 16 * KP, tap = r / KP, k = r - tap * KP, c = c0 + cc;
            w_s[cc][tap][k] = (c < C && k < K) ? __ldg(w + ((size_t)c * K + k) * 16 + tap) : 0.f;
        }
        __syncthreads();
#pragma unroll 2
        for (int cc = 0; cc < H2_CH; ++cc) {
            float a[3][4];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const float2 lo = *reinterpret_cast<const float2 *>(&in_s[cc][ty + r][tx2]);
                const float2 hi = *reinterpret_cast<const float2 *>(&in_s[cc][ty + r][tx2 + 2]);
                a[r][0] = lo.x; a[r][1] = lo.y; a[r][2] = hi.x; a[r][3] = hi.y;
            }
#pragma unroll
            for (int ky = 0; ky < 4; ++ky)
#pragma unroll
                for (int kx = 0; kx < 4; ++kx) {
                    const float4 *wp = reinterpret_cast<const float4 *>(&w_s[cc][ky * 4 + kx][12 * half]);
                    const float4 w0 = wp[0], w1 = wp[1], w2 = wp[2];
                    const float wk[12] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x, w2.y, w2.z, w2.w};
                    const int py = tap_par(ky), px = tap_par(kx), ro = tap_off(ky), co = tap_off(kx);
#pragma unroll
                    for (int p = 0; p < 2; ++p) {
                        const float v = a[ro][co + p];
#pragma unroll
                        for (int k = 0; k < 12; ++k) acc[p][py][px][k] = fmaf(v, wk[k], acc[p][py][px][k]);
                    }
                }
        }
    }

    const int iy = y0 + ty, ix0 = x0 + tx2;
    if (iy >= Hq) return;
    const int S_y = 2 * Hq, S_x = 2 * Wq;
    if (FORMAT == JHN_HM_F32_PLANAR) {
        float *out = (float *)out_;
#pragma unroll
        for (int k = 0; k < 12; ++k) {
            const int kk = 12 * half + k;
            if (kk >= K) continue;
#pragma unroll
            for (int py = 0; py < 2; ++py) {
                float *row = out + (((size_t)n * K + kk) * S_y + 2 * iy + py) * S_x + 2 * ix0;
                if (ix0 + 1 < Wq && (S_x & 3) == 0)
                    *reinterpret_cast<float4 *>(row) = make_float4(acc[0][py][0][k], acc[0][py][1][k], acc[1][py][0][k], acc[1][py][1][k]);
                else
                    for (int p = 0; p < 2; ++p)
                        if (ix0 + p < Wq) { row[2 * p] = acc[p][py][0][k]; row[2 * p + 1] = acc[p][py][1][k]; }
            }
        }
    } else {
        // channels-last 16-bit, padded: pixel (oy, ox) lives at [oy + 1][ox + 1]; this thread owns bytes 24*half .. +24 of
        // each of its pixel vectors, and the border vectors next to the image edge (zeros: F.pad, model.py:65-66)
        const int hsy = S_y + 2, hsx = S_x + 2;
        uint8_t *img = (uint8_t *)out_ + (size_t)n * hsy * hsx * KP * 2;
        auto store = [&](int yy, int xx, const float *v) {           // v == nullptr: zeros
            uint32_t pk[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                const float f0 = v ? v[2 * i] : 0.f, f1 = v ? v[2 * i + 1] : 0.f;
                if (FORMAT == JHN_HM_F16_CL) {
                    const __half2 h2 = __floats2half2_rn(f0 * JHN_HM_F16_SCALE, f1 * JHN_HM_F16_SCALE);
                    pk[i] = *reinterpret_cast<const uint32_t *>(&h2);
                } else {
                    const __nv_bfloat162 h2 = __floats2bfloat162_rn(f0, f1);
                    pk[i] = *reinterpret_cast<const uint32_t *>(&h2);
                }
            }
            uint2 *dst = reinterpret_cast<uint2 *>(img + ((size_t)yy * hsx + xx) * KP * 2 + 24 * half);
            dst[0] = make_uint2(pk[0], pk[1]); dst[1] = make_uint2(pk[2], pk[3]); dst[2] = make_uint2(pk[4], pk[5]);
        };
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            if (ix0 + p >= Wq) continue;
#pragma unroll
            for (int py = 0; py < 2; ++py)
#pragma unroll
                for (int px = 0; px < 2; ++px) {
                    const int oy = 2 * iy + py, ox = 2 * (ix0 + p) + px;
                    store(oy + 1, ox + 1, acc[p][py][px]);
                    const bool top = oy == 0, bot = oy == S_y - 1, lef = ox == 0, rig = ox == S_x - 1;
                    if (top) store(0, ox + 1, nullptr);
                    if (bot) store(hsy - 1, ox + 1, nullptr);
                    if (lef) store(oy + 1, 0, nullptr);
                    if (rig) store(oy + 1, hsx - 1, nullptr);
                    if (top && lef) store(0, 0, nullptr);
                    if (top && rig) store(0, hsx - 1, nullptr);
                    if (bot && lef) store(hsy - 1, 0, nullptr);
                    if (bot && rig) store(hsy - 1, hsx - 1, nullptr);
                }
        }
    }
}

int efftrack_head_launch(const float *features, const float *weight, int N, int C, int K, int Hq, int Wq, int out_format,
                         void *heatmaps, cudaStream_t st)
{
    const int tiles = cdiv(Hq, H2_T) * cdiv(Wq, H2_T);
    const dim3 grid(tiles, N);
    switch (out_format) {
    case JHN_HM_F32_PLANAR:
        JHN_LAUNCH("efftrack_head_kernel", st, efftrack_head_kernel<JHN_HM_F32_PLANAR><<<grid, 256, 0, st>>>(features, weight, C, K, Hq, Wq, heatmaps));
        return JHN_OK;
    case JHN_HM_F16_CL:
        JHN_LAUNCH("efftrack_head_kernel", st, efftrack_head_kernel<JHN_HM_F16_CL><<<grid, 256, 0, st>>>(features, weight, C, K, Hq, Wq, heatmaps));
        return JHN_OK;
    case JHN_HM_BF16_CL:
        JHN_LAUNCH("efftrack_head_kernel", st, efftrack_head_kernel<JHN_HM_BF16_CL><<<grid, 256, 0, st>>>(features, weight, C, K, Hq, Wq, heatmaps));
        return JHN_OK;
    }
    return fail(JHN_ERR_ARG, "jhn_efftrack_head: unknown heat-map format %d", out_format);
}

}  // namespace jhn
