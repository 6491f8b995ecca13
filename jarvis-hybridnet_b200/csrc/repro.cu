// Stage 1: reprojection gather.  Replaces jarvis/hybridnet/repro_layer.py:40-119 (+ F.pad and /255 of
// jarvis/hybridnet/model.py:65-66,72).
//
// Kernels (all HBM/L2-bound integer + gather work, no tensor cores):
//   coarse_project_kernel  half-resolution grid -> per-camera distorted, clamped pixel coordinates
//                          (repro_layer.py:46-68).  Separately rounded fp32 ops in the reference's order;
//                          the K=4 dot product is cuBLAS/MKL's FMA chain.
//   fine index (phase A of gather_fused_kernel)
//                          ATen upsample_trilinear3d (align_corners=False, scale 1/2) of both coordinate
//                          volumes, /2, truncate, y*hs+x (repro_layer.py:70-83).  One work item owns the
//                          <=2x2x2 fine voxels that share the same 8 coarse corners, so the corners are
//                          read once and the separable lerps are shared (bit-identical to ATen's nested
//                          expression because every intermediate is an fp32 value in both).  The indices
//                          live in shared memory only (the reference materialises them as a 36 MB int64
//                          tensor); they are written to HBM only when the caller asks for the parity dump.
//   relayout_kernel        [ncam][K][S][S] planar fp32 -> channels-last [ncam][hs][hs][KP] (fp32, bf16 or fp16;
//                          the fp16 copy of the streaming gather only inside each camera's pixel box of the grid)
//                          with the 1-px zero border of F.pad materialised, so that one voxel x camera
//                          gather is a single contiguous KP-vector instead of K strided scalars.
//   gather_fused_kernel    (fp32 parity path, and whenever the int32 index dump is requested) per 8x8x8 voxel
//                          tile: coarse corner coordinates of all cameras -> smem, phase A above, then phase B =
//                          index_select + mean over cameras (+ /255), cameras accumulated in order.  Writes NCDHW
//                          fp32 or the bf16 parity-split layout of the first tensor-core convolution.
//   gather_stream_kernel   (bf16 throughput path) the same arithmetic as a persistent, warp-specialised kernel:
//                          pixel boxes of a channels-last fp16 staging copy are TMA-staged in shared memory and
//                          gathered with LDS.128; see the block comment above the kernel.
#include <atomic>

#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace jhn {

__device__ __forceinline__ float lerp_rn(float w0, float a, float w1, float b, int mode)
{
    if (mode == JHN_LERP_FMA_FIRST) return __fmaf_rn(w0, a, __fmul_rn(w1, b));
    if (mode == JHN_LERP_FMA_SECOND) return __fmaf_rn(w1, b, __fmul_rn(w0, a));
    return __fadd_rn(__fmul_rn(w0, a), __fmul_rn(w1, b));
}

// One thread per coarse grid point of one frame set, looping over the cameras: the point's world coordinates are
// formed once and the per-camera chains (two IEEE divisions each) are independent, which gives the scheduler the
// instruction-level parallelism a one-projection-per-thread layout lacks.  Camera constants sit in shared memory.
constexpr int CP_PARAMS = 20;                           // P[12], fx, fy, cx, cy, k1, k2, chx, chy
__global__ void __launch_bounds__(256)
coarse_project_kernel(const float *__restrict__ cam, const float *__restrict__ intr,
                      const float *__restrict__ dist, const float *__restrict__ center3D,
                      const int32_t *__restrict__ centerHM, int B, int ncam, int h, float spacing, int hs,
                      float2 *__restrict__ cab, int *__restrict__ roi)
{
    extern __shared__ float cp[];                       // [ncam][CP_PARAMS], then (roi != null) int [ncam][4] block-level box
    int *sbox = reinterpret_cast<int *>(cp + ncam * CP_PARAMS);
    if (roi)
        for (int e = threadIdx.x; e < ncam * 4; e += blockDim.x) sbox[e] = 0x7f7f7f7f;
    const int b = blockIdx.y;
    for (int e = threadIdx.x; e < ncam * CP_PARAMS; e += blockDim.x) {
        const int c = e / CP_PARAMS, q = e - c * CP_PARAMS, bc = b * ncam + c;
        float v;
        if (q < 12) v = cam[12 * bc + q];
        else if (q == 12) v = intr[9 * bc + 0];
        else if (q == 13) v = intr[9 * bc + 4];
        else if (q == 14) v = intr[9 * bc + 6];
        else if (q == 15) v = intr[9 * bc + 7];
        else if (q == 16) v = dist[5 * bc + 0];
        else if (q == 17) v = dist[5 * bc + 1];
        else v = (float)centerHM[2 * bc + (q - 18)];
        cp[e] = v;
    }
    __syncthreads();
    const int nc = h * h * h;
    const int tt = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = tt < nc;
    const int t = live ? tt : nc - 1;                   // lanes past the end repeat the last point (they only feed the box)
    const int k = t % h, j = (t / h) % h, i = t / (h * h);
    const int half = h / 2;
    const float fhs = (float)hs, fhs1 = (float)(hs - 1), fhs2 = (float)(hs - 2);
    // grid = (idx - half) * spacing * 2 + center        repro_layer.py:32-36,113
    const float X = __fadd_rn(__fmul_rn(__fmul_rn((float)(i - half), spacing), 2.f), center3D[3 * b + 0]);
    const float Y = __fadd_rn(__fmul_rn(__fmul_rn((float)(j - half), spacing), 2.f), center3D[3 * b + 1]);
    const float Z = __fadd_rn(__fmul_rn(__fmul_rn((float)(k - half), spacing), 2.f), center3D[3 * b + 2]);
#pragma unroll 4
    for (int c = 0; c < ncam; ++c) {
        const float *P = cp + c * CP_PARAMS;
        const float fx = P[12], fy = P[13], cx = P[14], cy = P[15], k1 = P[16], k2 = P[17], chx = P[18], chy = P[19];
        // integer-valued floats: chx - (hs-1) and chx + hs-2 are exact, same values as the int arithmetic of :65-68
        const float lox = chx - fhs1, hix = chx + fhs2, loy = chy - fhs1, hiy = chy + fhs2;
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
        const size_t o = ((size_t)b * ncam + c) * nc + t;
        if (live && cab) cab[o] = make_float2(a, bb);        // (x, y) interleaved: one 8-byte access per corner downstream
        if (roi) {
            // pixel box of this camera over the whole voxel grid = min / max over all coarse corners (the fine coordinates
            // are rounded convex combinations of them): warp REDUX -> shared atomics -> one global atomic per block.
            // Stored as {x0, y0, -x1, -y1} so that all four are minima of a buffer memset to 0x7f7f7f7f.
            const int px = __float2int_rz(__fmul_rn(a, 0.5f)), py = __float2int_rz(__fmul_rn(bb, 0.5f));
            const int x0 = __reduce_min_sync(0xffffffffu, px), y0 = __reduce_min_sync(0xffffffffu, py);
            const int x1 = __reduce_max_sync(0xffffffffu, px), y1 = __reduce_max_sync(0xffffffffu, py);
            if ((threadIdx.x & 31) == 0) {
                atomicMin(sbox + 4 * c + 0, x0); atomicMin(sbox + 4 * c + 1, y0);
                atomicMin(sbox + 4 * c + 2, -x1); atomicMin(sbox + 4 * c + 3, -y1);
            }
        }
    }
    if (roi) {
        __syncthreads();
        for (int e = threadIdx.x; e < ncam * 4; e += blockDim.x) atomicMin(roi + ((size_t)b * ncam) * 4 + e, sbox[e]);
    }
}

// fp16 staging of the streaming gather: values are scaled by 2^-4 (exact) so that the sum over up to 64 cameras
// of heat-map values as large as 16 000 stays inside the fp16 range; the gather multiplies the sum back.
constexpr float HALF_STAGE_SCALE = 0.0625f, HALF_STAGE_UNSCALE = 16.f;

template <typename T> __device__ __forceinline__ T to_store(float v);
template <> __device__ __forceinline__ float to_store<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 to_store<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half to_store<__half>(float v) { return __float2half_rn(v * HALF_STAGE_SCALE); }

// 8 consecutive channels of one (camera, pixel) KP-vector, added into fp32 accumulators
__device__ __forceinline__ void add8(const float *p, float *acc)
{
    const float4 a = __ldg(reinterpret_cast<const float4 *>(p)), b = __ldg(reinterpret_cast<const float4 *>(p) + 1);
    acc[0] = __fadd_rn(acc[0], a.x); acc[1] = __fadd_rn(acc[1], a.y); acc[2] = __fadd_rn(acc[2], a.z); acc[3] = __fadd_rn(acc[3], a.w);
    acc[4] = __fadd_rn(acc[4], b.x); acc[5] = __fadd_rn(acc[5], b.y); acc[6] = __fadd_rn(acc[6], b.z); acc[7] = __fadd_rn(acc[7], b.w);
}
__device__ __forceinline__ void add8(const __nv_bfloat16 *p, float *acc)
{
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p));
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {                       // bf16 -> fp32 is a 16-bit shift
        acc[2 * i + 0] += __uint_as_float(w[i] << 16);
        acc[2 * i + 1] += __uint_as_float(w[i] & 0xffff0000u);
    }
}

__device__ __forceinline__ void add8(const __half *p, float *acc)
{
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p));
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {                       // staged values carry the exact factor 2^-4
        const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&w[i]));
        acc[2 * i + 0] += f.x * HALF_STAGE_UNSCALE;
        acc[2 * i + 1] += f.y * HALF_STAGE_UNSCALE;
    }
}

// bf16 channels-last -> the streaming gather's fp16 staging format (x 2^-4), 16 bytes per thread
__global__ void __launch_bounds__(256)
bf16cl_to_f16cl_kernel(const uint4 *__restrict__ in, uint4 *__restrict__ out, size_t n16)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n16) return;
    const uint4 v = __ldg(in + i);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const __half2 h = __floats2half2_rn(__uint_as_float(w[k] << 16) * HALF_STAGE_SCALE, __uint_as_float(w[k] & 0xffff0000u) * HALF_STAGE_SCALE);
        o[k] = *reinterpret_cast<const uint32_t *>(&h);
    }
    out[i] = make_uint4(o[0], o[1], o[2], o[3]);
}

// planar -> channels-last: one thread per padded pixel reads its K channel values (for a fixed channel the lanes
// of a warp read consecutive floats: coalesced, K independent loads in flight per thread); the 16-bit staging
// copy goes through a per-warp shared-memory tile so that each store instruction writes 512 contiguous bytes.
template <typename T>
__global__ void __launch_bounds__(256)
relayout_pixel_kernel(const float *__restrict__ in, int K, int hs, int padded, long long npix, const int4 *__restrict__ roi,
                      T *__restrict__ out)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = idx < npix;                                    // lanes past the end stay for the warp-wide staging below
    const int hh = hs * hs;
    const long long bc = (live ? idx : 0) / hh;
    const int r = (int)((live ? idx : 0) - bc * hh), y = r / hs, x = r - y * hs;
    if (roi) {
        // only the pixels the gather can touch: the camera's box over the whole voxel grid (coarse_project_kernel); whole warps
        // outside it leave without a load or a store (their part of the staging copy is never read)
        const int4 bx = __ldg(roi + bc);
        const bool hit = live && x >= bx.x && x <= -bx.z && y >= bx.y && y <= -bx.w;        // {x0, y0, -x1, -y1}
        if (!__any_sync(0xffffffffu, hit)) return;
    }
    const int S = padded ? hs : hs - 2, off = padded ? 0 : 1;
    const int ys = y - off, xs = x - off;
    const bool inside = live && ys >= 0 && ys < S && xs >= 0 && xs < S;
    float v[KP];
    const float *src = in + ((size_t)bc * K * S + (inside ? ys : 0)) * S + (inside ? xs : 0);
#pragma unroll
    for (int k = 0; k < KP; ++k) v[k] = (inside && k < K) ? __ldg(src + (size_t)k * S * S) : 0.f;
    if (sizeof(T) == 4) {
        if (!live) return;
        T *dst = out + (size_t)idx * KP;
#pragma unroll
        for (int i = 0; i < KP / 4; ++i)
            reinterpret_cast<float4 *>(dst)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    } else {
        // 16-bit staging: a warp's 32 pixel vectors are 1536 contiguous bytes; pass them through shared memory so that
        // every store instruction writes 512 contiguous bytes instead of 32 x 16 B at a 48-byte stride
        __shared__ __align__(16) uint4 stage[256 / 32][32 * (KP / 8)];
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
        for (int i = 0; i < KP / 8; ++i) {
            T t[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) t[j] = to_store<T>(v[8 * i + j]);
            stage[warp][lane * (KP / 8) + i] = *reinterpret_cast<const uint4 *>(t);
        }
        __syncwarp();
        const long long first = idx - lane;                          // the warp's first pixel
        if (first >= npix) return;
        uint4 *dst = reinterpret_cast<uint4 *>(out + (size_t)first * KP);
        const long long n16 = (npix - first) * (KP / 8);             // 16-byte units left in the tensor from `first`
#pragma unroll
        for (int i = 0; i < KP / 8; ++i)
            if (i * 32 + lane < n16) dst[i * 32 + lane] = stage[warp][i * 32 + lane];
    }
}

template <typename T>
static int launch_relayout(const ReprojectArgs &a, T *hm_cl, const int4 *roi, cudaStream_t st)
{
    const long long npix = (long long)a.B * a.ncam * a.hs * a.hs;
    JHN_LAUNCH("relayout_kernel", st,
               relayout_pixel_kernel<T><<<cdiv(npix, 256), 256, 0, st>>>((const float *)a.heatmaps, a.K, a.hs, a.padded, npix, roi, hm_cl));
    return JHN_OK;
}

// jhn_heatmap_convert: whole maps, no pixel-box restriction (the caller may gather any grid from the result)
int heatmap_convert_launch(const float *hm, int padded, int B, int ncam, int K, int hs, int dst_format, void *dst, cudaStream_t st)
{
    ReprojectArgs a{};
    a.heatmaps = hm; a.hm_format = JHN_HM_F32_PLANAR; a.padded = padded; a.B = B; a.ncam = ncam; a.K = K; a.hs = hs;
    if (dst_format == JHN_HM_F16_CL) return launch_relayout<__half>(a, (__half *)dst, nullptr, st);
    if (dst_format == JHN_HM_BF16_CL) return launch_relayout<__nv_bfloat16>(a, (__nv_bfloat16 *)dst, nullptr, st);
    return fail(JHN_ERR_ARG, "jhn_heatmap_convert: dst_format %d is not a channels-last format", dst_format);
}

// jhn_heatmap_boxes: only the per-camera pixel boxes of the voxel grid (no coordinate dump): what a host needs to know to
// upload just the pixels the gather can touch.  boxes: int4 {x0, y0, -x1, -y1} per (frame set, camera).
int heatmap_boxes_launch(const float *cam, const float *intr, const float *dist, const float *center3D, const int32_t *centerHM,
                         int B, int ncam, int hs, int G, float spacing, int32_t *boxes, cudaStream_t st)
{
    const int h = G / 2;
    JHN_CUDA(cudaMemsetAsync(boxes, 0x7f, (size_t)B * ncam * sizeof(int4), st));
    JHN_LAUNCH("coarse_project_kernel", st,
               coarse_project_kernel<<<dim3(cdiv((long long)h * h * h, 256), B), 256,
                                       ncam * (CP_PARAMS * sizeof(float) + 4 * sizeof(int)), st>>>(
                   cam, intr, dist, center3D, centerHM, B, ncam, h, spacing, hs, nullptr, (int *)boxes));
    return JHN_OK;
}

// jhn_heatmap_spans: per (frame set, camera, pixel row) the column range [lo, hi] the gather can touch — tighter than the box
// (the voxel cube projects to a hexagon: the row spans hold 78 % of the box's pixels at the Example shape, and every index
// the reference computes: tests).  A fine voxel's coordinates are rounded convex combinations of the 8 coarse corners of its
// cell and (v / 2).int() is monotone, so its pixel lies inside the integer box of those corners: every cell adds its box's
// columns to the rows its box covers.  Shared-memory atomics per (camera, slab of cells), one global atomic per row and block.
// spans: int2 {lo, -hi} per row, minima of a buffer memset to 0x7f7f7f7f (rows no voxel maps to keep lo > hi).
constexpr int SPAN_MAX_HS = 2048;
__global__ void __launch_bounds__(256)
row_spans_kernel(const float2 *__restrict__ cab, int ncam, int h, int hs, int i_per_block, int2 *__restrict__ spans)
{
    __shared__ int s_lo[SPAN_MAX_HS], s_nhi[SPAN_MAX_HS];
    const int c = blockIdx.y, b = blockIdx.z;
    for (int y = threadIdx.x; y < hs; y += blockDim.x) { s_lo[y] = 0x7f7f7f7f; s_nhi[y] = 0x7f7f7f7f; }
    __syncthreads();
    const float2 *src = cab + ((size_t)b * ncam + c) * h * h * h;
    const int hc = h > 1 ? h - 1 : 1;                                   // cells per dimension (h == 1: one degenerate cell)
    const int i1 = min((int)(blockIdx.x + 1) * i_per_block, hc);
    for (int i = blockIdx.x * i_per_block; i < i1; ++i) {
        const int ip = min(i + 1, h - 1);
        for (int q = threadIdx.x; q < hc * hc; q += blockDim.x) {
            const int j = q / hc, k = q - j * hc, jp = min(j + 1, h - 1), kp = min(k + 1, h - 1);
            int x0 = 0x7fffffff, x1 = -0x7fffffff, y0 = 0x7fffffff, y1 = -0x7fffffff;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float2 v = __ldg(src + ((size_t)((e & 4) ? ip : i) * h + ((e & 2) ? jp : j)) * h + ((e & 1) ? kp : k));
                const int px = __float2int_rz(__fmul_rn(v.x, 0.5f)), py = __float2int_rz(__fmul_rn(v.y, 0.5f));
                x0 = min(x0, px); x1 = max(x1, px); y0 = min(y0, py); y1 = max(y1, py);
            }
            y0 = max(y0, 0); y1 = min(y1, hs - 1);
            for (int y = y0; y <= y1; ++y) { atomicMin(s_lo + y, x0); atomicMin(s_nhi + y, -x1); }
        }
    }
    __syncthreads();
    int *g = reinterpret_cast<int *>(spans + ((size_t)b * ncam + c) * hs);
    for (int y = threadIdx.x; y < hs; y += blockDim.x)
        if (s_lo[y] != 0x7f7f7f7f) { atomicMin(g + 2 * y, s_lo[y]); atomicMin(g + 2 * y + 1, s_nhi[y]); }
}

int heatmap_spans_launch(const float *cam, const float *intr, const float *dist, const float *center3D, const int32_t *centerHM,
                         int B, int ncam, int hs, int G, float spacing, void *scratch, int32_t *boxes, int32_t *spans, cudaStream_t st)
{
    const int h = G / 2;
    if (hs > SPAN_MAX_HS) return fail(JHN_ERR_SHAPE, "jhn_heatmap_spans: maps of up to %d rows (got %d)", SPAN_MAX_HS, hs);
    JHN_CUDA(cudaMemsetAsync(boxes, 0x7f, (size_t)B * ncam * sizeof(int4), st));
    JHN_CUDA(cudaMemsetAsync(spans, 0x7f, (size_t)B * ncam * hs * sizeof(int2), st));
    JHN_LAUNCH("coarse_project_kernel", st,
               coarse_project_kernel<<<dim3(cdiv((long long)h * h * h, 256), B), 256,
                                       ncam * (CP_PARAMS * sizeof(float) + 4 * sizeof(int)), st>>>(
                   cam, intr, dist, center3D, centerHM, B, ncam, h, spacing, hs, (float2 *)scratch, (int *)boxes));
    const int hc = h > 1 ? h - 1 : 1, ipb = 4;
    JHN_LAUNCH("row_spans_kernel", st,
               row_spans_kernel<<<dim3(cdiv(hc, ipb), ncam, B), 256, 0, st>>>((const float2 *)scratch, ncam, h, hs, ipb, (int2 *)spans));
    return JHN_OK;
}

constexpr int TS = 8;                                   // voxel tile side
constexpr int CS = TS / 2 + 2;                          // coarse corners per dimension needed by a tile

// One CTA = one TS^3 tile of fine voxels of frame set blockIdx.y.
template <typename T, int LAYOUT>
__global__ void __launch_bounds__(256)
gather_fused_kernel(const T *__restrict__ hm, const float2 *__restrict__ cab, int ncam, int K,
                    int hs, int G, int lerp_mode, float post_divide, void *__restrict__ out_, int32_t *__restrict__ idx_out)
{
    extern __shared__ __align__(16) uint8_t gsm[];
    float *co = reinterpret_cast<float *>(gsm);                         // [ncam][2][CS^3] coarse a / b
    int32_t *ix = reinterpret_cast<int32_t *>(co + (size_t)ncam * 2 * CS * CS * CS);   // [ncam][TS^3]
    const int h = G / 2, nt = (G + TS - 1) / TS;
    const int b = blockIdx.y;
    const int tk = blockIdx.x % nt, tj = (blockIdx.x / nt) % nt, ti = blockIdx.x / (nt * nt);
    const int I0 = ti * TS, J0 = tj * TS, K0 = tk * TS;
    const size_t nc = (size_t)h * h * h, nv = (size_t)G * G * G;

    // ---- coarse corners -> smem (indices clamped to the grid; clamped duplicates are what ATen reads too)
    for (int e = threadIdx.x; e < ncam * CS * CS * CS; e += blockDim.x) {
        const int lk = e % CS, lj = (e / CS) % CS, li = (e / (CS * CS)) % CS, c = e / (CS * CS * CS);
        const int gi = min(max(I0 / 2 - 1 + li, 0), h - 1), gj = min(max(J0 / 2 - 1 + lj, 0), h - 1),
                  gk = min(max(K0 / 2 - 1 + lk, 0), h - 1);
        const size_t o = ((size_t)b * ncam + c) * nc + ((size_t)gi * h + gj) * h + gk;
        const int l = (li * CS + lj) * CS + lk;
        const float2 ab = __ldg(cab + o);
        co[(c * 2 + 0) * CS * CS * CS + l] = ab.x;
        co[(c * 2 + 1) * CS * CS * CS + l] = ab.y;
    }
    __syncthreads();

    // ---- phase A: indices of the tile.  Work item = (camera, corner block); block m of a dimension owns the
    // fine voxels {I0+2m-1 (lambda1 .25), I0+2m (lambda1 .75)} inside the tile, both between corners m and m+1
    // (ATen area_pixel_compute_source_index, scale .5: src = .5*(I+.5)-.5 clamped at 0, so I=0 has lambda1 0).
    constexpr int NB = TS / 2 + 1;
    for (int w = threadIdx.x; w < ncam * NB * NB * NB; w += blockDim.x) {
        const int mk = w % NB, mj = (w / NB) % NB, mi = (w / (NB * NB)) % NB, c = w / (NB * NB * NB);
        const float *A = co + (c * 2 + 0) * CS * CS * CS, *Bc = co + (c * 2 + 1) * CS * CS * CS;
        float va[2][2][2], vb[2][2][2];
#pragma unroll
        for (int p = 0; p < 2; ++p)
#pragma unroll
            for (int q = 0; q < 2; ++q)
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const int l = ((mi + p) * CS + (mj + q)) * CS + (mk + r);
                    va[p][q][r] = A[l];
                    vb[p][q][r] = Bc[l];
                }
#pragma unroll
        for (int kv = 0; kv < 2; ++kv) {
            const int lk = 2 * mk - 1 + kv, Kz = K0 + lk;
            if (lk < 0 || lk >= TS || Kz >= G) continue;
            const float lk1 = Kz == 0 ? 0.f : (kv ? 0.75f : 0.25f), lk0 = __fsub_rn(1.f, lk1);
            float xa[2][2], xb[2][2];
#pragma unroll
            for (int p = 0; p < 2; ++p)
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    xa[p][q] = lerp_rn(lk0, va[p][q][0], lk1, va[p][q][1], lerp_mode);
                    xb[p][q] = lerp_rn(lk0, vb[p][q][0], lk1, vb[p][q][1], lerp_mode);
                }
#pragma unroll
            for (int jv = 0; jv < 2; ++jv) {
                const int lj = 2 * mj - 1 + jv, J = J0 + lj;
                if (lj < 0 || lj >= TS || J >= G) continue;
                const float lj1 = J == 0 ? 0.f : (jv ? 0.75f : 0.25f), lj0 = __fsub_rn(1.f, lj1);
                float ya[2], yb[2];
#pragma unroll
                for (int p = 0; p < 2; ++p) {
                    ya[p] = lerp_rn(lj0, xa[p][0], lj1, xa[p][1], lerp_mode);
                    yb[p] = lerp_rn(lj0, xb[p][0], lj1, xb[p][1], lerp_mode);
                }
#pragma unroll
                for (int iv = 0; iv < 2; ++iv) {
                    const int li = 2 * mi - 1 + iv, I = I0 + li;
                    if (li < 0 || li >= TS || I >= G) continue;
                    const float li1 = I == 0 ? 0.f : (iv ? 0.75f : 0.25f), li0 = __fsub_rn(1.f, li1);
                    const float fa = lerp_rn(li0, ya[0], li1, ya[1], lerp_mode);
                    const float fb = lerp_rn(li0, yb[0], li1, yb[1], lerp_mode);
                    const int px = __float2int_rz(__fmul_rn(fa, 0.5f));      // (val1/2).int()   :82-83
                    const int py = __float2int_rz(__fmul_rn(fb, 0.5f));
                    const int flat = py * hs + px;
                    ix[c * TS * TS * TS + (li * TS + lj) * TS + lk] = flat;
                    if (idx_out) idx_out[((size_t)b * ncam + c) * nv + ((size_t)I * G + J) * G + Kz] = flat;
                }
            }
        }
    }
    __syncthreads();

    // ---- phase B: gather + camera mean.  Lane = (voxel slot, 8-channel chunk): one warp request covers
    // 10 voxels x 3 chunks, and the 3 chunk lanes of a voxel read one contiguous 48-byte pixel vector, so a
    // request touches ~one cache line per distinct pixel instead of one per lane.
    constexpr int NCH = KP / 8, VPW = 32 / NCH;
    const T *base = hm + (size_t)b * ncam * hs * hs * KP;
    const float fn = (float)ncam;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int vslot = lane / NCH, j = lane - vslot * NCH;
#pragma unroll 1
    for (int g = warp; g * VPW < TS * TS * TS; g += 8) {
        const int v = g * VPW + vslot;
        const int lk = v % TS, lj = (v / TS) % TS, li = v / (TS * TS);
        const int I = I0 + li, J = J0 + lj, Kz = K0 + lk;
        if (vslot >= VPW || v >= TS * TS * TS || I >= G || J >= G || Kz >= G) continue;
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
        for (int c = 0; c < ncam; ++c) {
            const int32_t flat = ix[c * TS * TS * TS + v];
            add8(base + ((size_t)c * hs * hs + flat) * KP + 8 * j, acc);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float m = __fdiv_rn(acc[k], fn);                            // torch.mean        :103-105
            if (post_divide != 1.f) m = __fdiv_rn(m, post_divide);      // heatmaps3D/255.   model.py:72
            acc[k] = m;
        }
        if (LAYOUT == JHN_VOL_NCDHW_F32) {
            float *out = (float *)out_ + (size_t)b * K * nv + ((size_t)I * G + J) * G + Kz;
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (8 * j + k < K) out[(size_t)(8 * j + k) * nv] = acc[k];
        } else {
            // parity-split, channel-blocked bf16 volume read by the tensor-core front convolution (conv_tc.cu):
            // [b][s = (I&1,J&1,Kz&1)][j][zp][pp][8] on the G/2 grid; channel chunks beyond K are written as zeros
            const int CJ = (K + 15) / 16 * 2, Wh = G / 2 + 2;
            const int sv = ((I & 1) * 2 + (J & 1)) * 2 + (Kz & 1);
            uint4 *out = (uint4 *)out_;
            const size_t pos = ((size_t)(I >> 1) + 1) * Wh * Wh + (size_t)((J >> 1) + 1) * Wh + (Kz >> 1) + 1;
            const size_t chunk_stride = (size_t)Wh * Wh * Wh;
            const size_t ob = (((size_t)b * 8 + sv) * CJ) * chunk_stride + pos;
            uint32_t pk[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                __nv_bfloat162 h2 = __floats2bfloat162_rn(acc[2 * i], acc[2 * i + 1]);
                pk[i] = *reinterpret_cast<uint32_t *>(&h2);
            }
            if (j < CJ) out[ob + (size_t)j * chunk_stride] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            if (j == 0)
                for (int jz = NCH; jz < CJ; ++jz) out[ob + (size_t)jz * chunk_stride] = make_uint4(0, 0, 0, 0);
        }
    }
}


// ------------------------------------------------------------------------------------------------
// Shared pieces of the bf16 throughput gather (gather_stream_kernel below): tile geometry, fp16 staging, packed fp32x2
// index arithmetic.
// ------------------------------------------------------------------------------------------------
constexpr int GT = 8, GCELL = GT / 2, GC = GCELL + 1;                          // voxel tile side, coarse cells, corners per x / y
constexpr int G_PIX_BYTES = KP * 2;                                           // 48 B per staged pixel

// Packed fp32x2 arithmetic (sm_100 FMUL2 / FFMA2 / FADD2): the x and the y pixel coordinate of a corner travel
// as one 64-bit register pair through the three nested lerps, each half an IEEE round-to-nearest fp32
// operation exactly like the scalar instruction — half the issue slots of the index chain.
__device__ __forceinline__ uint64_t f2_pack(float x, float y)
{
    uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y)); return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t v, float &x, float &y)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v));
}
__device__ __forceinline__ uint64_t f2_mul(uint64_t a, uint64_t b)
{
    uint64_t r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;
}
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b)
{
    uint64_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;
}
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c)
{
    uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r;
}
template <int MODE> __device__ __forceinline__ uint64_t lerp2(uint64_t w0, uint64_t a, uint64_t w1, uint64_t b)
{
    if (MODE == JHN_LERP_FMA_FIRST) return f2_fma(w0, a, f2_mul(w1, b));
    if (MODE == JHN_LERP_FMA_SECOND) return f2_fma(w1, b, f2_mul(w0, a));
    return f2_add(f2_mul(w0, a), f2_mul(w1, b));
}

// ------------------------------------------------------------------------------------------------
// Streaming gather.  One work item = (8x8x8 voxel tile, camera).  The tile's fine coordinates are lerps of 5x5x6
// coarse corners, so every index of the item falls inside the pixel box spanned by the corners: the box (~140
// pixels x 48 B) is staged in shared memory by TMA and each (voxel, camera) gather is three LDS.128.
//
// An earlier one-tile-per-CTA version of this kernel (profiles/r01_run25) paid the chain  corner LDG -> box -> TMA ->
// first LDS  once per CTA (~13 us per tile against ~1 us of issue work) and parked 27 % of its lanes on voxels
// outside the grid.  Here CTAs are persistent (3 per SM, tiles blockIdx.x, +gridDim.x, ...) and warp-specialised:
//   4 producer warps (setmaxnreg 40), producer p owning box slot p and items p, p + 4, ... of the CTA's sequence:
//       corners of the item: 150 x cp.async (LDGSTS, 8 B) into the item's header, issued one item ahead
//       cp.async.wait_group -> integer min / max of the corners' pixels (REDUX) = the box -> smem row pitch chosen
//       for few bank conflicts -> wait for the slot -> mbarrier.arrive.expect_tx -> one cp.async.bulk per box row
//     (one producer warp cannot keep up: its dependent chain is ~2 us per item; four run their chains in parallel)
//   4 gather warps (setmaxnreg 120), thread = 2x2 voxels (x, y) at one z sharing the separable lerps:
//       wait full -> 8 corner pairs -> nested fp32x2 lerps -> F2I -> 12 x LDS.128 -> 48 x HADD2 -> arrive empty;
//       after the last camera: one fp32 scale, bf16 pack, 16-byte stores into the first convolution's input layout.
// z tiles are unshifted (6 corners along z instead of 5) so that no lane is ever outside the grid along z;
// along x / y whole warps / quarter-warps outside the grid skip the step (no LSU wavefronts).
// ------------------------------------------------------------------------------------------------
constexpr int GCK = GC + 1, GCN = GC * GC * GCK;                               // 5 x 5 x 6 corners per (tile, camera)
#ifndef GS_SLOTS_V
#define GS_SLOTS_V 4
#endif
#ifndef GS_LANE_ARRIVE
#define GS_LANE_ARRIVE 1
#endif
constexpr int GS_PROD = 4, GS_SLOTS = GS_SLOTS_V, GS_THREADS = 32 * (4 + GS_PROD);  // 4 producer warps, 4 gather warps, GS_SLOTS box slots
constexpr int GS_META = GCN * 8, GS_HDR = GS_META + 16;                        // header: corners (1200 B), then the int4 box
constexpr int GS_CORNERS_PER_LANE = (GCN + 31) / 32;

__device__ __forceinline__ void cp_async8(uint32_t dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16cg(uint32_t dst, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// this lane's corners of tile (ti, tj, tk): offsets in the coarse grid, indices clamped like ATen's reads
__device__ __forceinline__ void tile_corner_offsets(int lane, int ti, int tj, int tk, int h, int (&go)[GS_CORNERS_PER_LANE])
{
#pragma unroll
    for (int r = 0; r < GS_CORNERS_PER_LANE; ++r) {
        const int l = min(lane + 32 * r, GCN - 1);
        const int lk = l % GCK, lj = (l / GCK) % GC, li = l / (GCK * GC);
        const int gi = min(max(GCELL * ti - 1 + li, 0), h - 1), gj = min(max(GCELL * tj - 1 + lj, 0), h - 1),
                  gk = min(max(GCELL * tk - 1 + lk, 0), h - 1);
        go[r] = (gi * h + gj) * h + gk;
    }
}

template <int LAYOUT, int MODE>
__global__ void __launch_bounds__(GS_THREADS, 3)
gather_stream_kernel(const __half *__restrict__ hm, const float2 *__restrict__ cab,
                     int ncam, int K, int hs, int G, float post_scale, int cap_bytes, int box_limit, void *__restrict__ out_, int total_tiles)
{
    extern __shared__ __align__(128) uint8_t gsm[];                            // [GS_SLOTS][cap_bytes] pixel boxes
    uint8_t *hdrs = gsm + (size_t)GS_SLOTS * cap_bytes;                        // [2 * GS_SLOTS][GS_HDR]
    uint64_t *bars = reinterpret_cast<uint64_t *>(hdrs + 2 * GS_SLOTS * GS_HDR);
    uint64_t *full = bars, *empty = bars + GS_SLOTS;
    const int h = G / 2, nt = G / GT + 1, ntk = G / GT, tiles_fs = nt * nt * ntk;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t nc = (size_t)h * h * h;
    const int my_tiles = (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (threadIdx.x == 0) {
        for (int i = 0; i < GS_SLOTS; ++i) { mbar_init(smem_u32(full + i), GS_LANE_ARRIVE ? 32 : 1); mbar_init(smem_u32(empty + i), GS_LANE_ARRIVE ? 128 : 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp >= 4) {
        // =================================== producers ========================================
        // Producer p owns box slot p and the items n = p, p + GS_SLOTS, ...  (item n = tile n / ncam, camera n % ncam of
        // this CTA's tile sequence): its chain  corners -> box -> rows  is ~2 us of dependent latency per item, which
        // one warp cannot sustain per item but GS_SLOTS warps, each with GS_SLOTS gather steps of time per item, can.
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        const int p = warp - 4;
        const int total = my_tiles * ncam;
        int go[GS_CORNERS_PER_LANE];
        int go_tl = -1;
        const float2 *csrc = cab;
        auto prefetch = [&](int n, int hidx) {                                 // corners of item n -> header hidx (LDGSTS)
            const int tl = n / ncam, c = n - tl * ncam;
            if (tl != go_tl) {
                go_tl = tl;
                const int t = (int)blockIdx.x + tl * (int)gridDim.x;
                const int b = t / tiles_fs, r0 = t - b * tiles_fs;
                const int tk = r0 % ntk, tj = (r0 / ntk) % nt, ti = r0 / (ntk * nt);
                tile_corner_offsets(lane, ti, tj, tk, h, go);
                csrc = cab + (size_t)b * ncam * nc;
            }
            const float2 *src = csrc + (size_t)c * nc;
            const uint32_t dst = smem_u32(hdrs + hidx * GS_HDR);
#pragma unroll
            for (int r = 0; r < GS_CORNERS_PER_LANE; ++r)
                if (lane + 32 * r < GCN) cp_async8(dst + (uint32_t)(lane + 32 * r) * 8u, src + go[r]);
            cp_async_commit();
        };
        // item n lives in box slot n % GS_SLOTS and header n % (2 GS_SLOTS); producer p takes the items p, p + GS_PROD, ...
        // (slot, wrap parity, header) advance by GS_PROD per item without a division
        if (p < total) prefetch(p, p);
        int slot = p, hidx = p; uint32_t ph = 0;                               // GS_PROD <= GS_SLOTS: the first items start in slot p, use 0
        bool reuse = false;                                                    // the slot has had an earlier item
        for (int n = p; n < total; n += GS_PROD) {
            cp_async_wait<0>();
            __syncwarp();                                                      // all lanes' copies visible to all lanes
            // pixel box bounding every index of the tile: each ATen lerp is a rounded convex combination, so the fine
            // coordinates stay inside the corners' range; (v / 2).int() is monotone, so the box is the integer
            // min / max of the corners' own pixels (REDUX instead of a float shuffle tree)
            uint8_t *hd = hdrs + hidx * GS_HDR;
            int x0, x1, y0, y1;
            {
                const float2 *cn = reinterpret_cast<const float2 *>(hd);
                const float2 v0 = cn[lane];                                    // this lane's own copies (lane < GCN)
                x0 = x1 = __float2int_rz(__fmul_rn(v0.x, 0.5f)); y0 = y1 = __float2int_rz(__fmul_rn(v0.y, 0.5f));
#pragma unroll
                for (int r = 1; r < GS_CORNERS_PER_LANE; ++r) {
                    const int l = lane + 32 * r;
                    const float2 v = cn[l < GCN ? l : lane];
                    const int px = __float2int_rz(__fmul_rn(v.x, 0.5f)), py = __float2int_rz(__fmul_rn(v.y, 0.5f));
                    x0 = min(x0, px); x1 = max(x1, px); y0 = min(y0, py); y1 = max(y1, py);
                }
                x0 = __reduce_min_sync(0xffffffffu, x0); x1 = __reduce_max_sync(0xffffffffu, x1);
                y0 = __reduce_min_sync(0xffffffffu, y0); y1 = __reduce_max_sync(0xffffffffu, y1);
            }
            const int bw = x1 - x0 + 1, bh = y1 - y0 + 1;
            // smem row pitch in pixels.  A pixel vector is three 16-byte bank groups, so two pixels of one LDS.128 phase
            // (8 lanes = 8 voxels along z) collide iff their linear offsets x + pitch * y agree mod 8.  Which pitch
            // residue keeps a z run of pixels apart depends on the camera's view of the z axis, so the warp tries the
            // four odd residues on the tile's central z line (lane = residue x voxel) and takes the one with the fewest
            // colliding lanes; then the smallest pitch >= bw of that residue that fits the slot, else odd, else the
            // bare width; 0 = the box does not fit and the step gathers from global memory.
            int pitch;
            {
                const float2 *cn = reinterpret_cast<const float2 *>(hd) + ((GC / 2) * GC + GC / 2) * GCK;
                const int zz = lane & 7, k0 = (zz + 1) >> 1, res = 2 * (lane >> 3) + 1;
                const float w1 = (zz & 1) ? 0.25f : 0.75f;
                const float2 ca = cn[k0], cb = cn[k0 + 1];
                const int X = __float2int_rz(0.5f * (ca.x + w1 * (cb.x - ca.x))), Y = __float2int_rz(0.5f * (ca.y + w1 * (cb.y - ca.y)));
                const unsigned grp = 0xffu << (lane & 24);
                const unsigned same_bank = __match_any_sync(0xffffffffu, ((X + res * Y) & 7) | (res << 3));
                const unsigned same_pix = __match_any_sync(0xffffffffu, (X & 0xfff) | ((Y & 0xfff) << 12) | (res << 24));
                const unsigned clash = __ballot_sync(0xffffffffu, (same_bank & ~same_pix & grp) != 0u);
                const int c1 = __popc(clash & 0xffu), c3 = __popc(clash & 0xff00u), c5 = __popc(clash & 0xff0000u), c7 = __popc(clash >> 24);
                int best = 3, cost = c3;
                if (c5 < cost) { best = 5; cost = c5; }
                if (c1 < cost) { best = 1; cost = c1; }
                if (c7 < cost) { best = 7; cost = c7; }
                pitch = bw + ((best - bw) & 7);
            }
            if (pitch * bh * G_PIX_BYTES > box_limit) pitch = bw | 1;
            if (pitch * bh * G_PIX_BYTES > box_limit) pitch = bw;
            if (pitch * bh * G_PIX_BYTES > box_limit || x0 < 0 || y0 < 0 || x1 >= hs || y1 >= hs) pitch = 0;
            if (lane == 0) *reinterpret_cast<int4 *>(hd + GS_META) = make_int4(x0, y0, slot * cap_bytes, pitch);
            const uint32_t fb = smem_u32(full + slot);
            if (reuse) mbar_wait_parked(smem_u32(empty + slot), ph ^ 1u);      // item n - GS_SLOTS is done: slot + older header free
            __syncwarp();                                                      // every lane's corner copies + the box precede the arrive
            // GS_LANE_ARRIVE: every producer lane arrives on the full barrier itself (count 32), i.e. each lane RELEASES its own
            // cp.async corner copies instead of handing them to lane 0 through __syncwarp — the same ordering, but in the form
            // compute-sanitizer's racecheck models (it reported the one-arrival form as a write / read hazard on the header).
            if (pitch == 0) {
                if (GS_LANE_ARRIVE || lane == 0) mbar_arrive(fb);              // box too large: gathered from global memory
            } else {
                const uint32_t row_bytes = (uint32_t)(bw * G_PIX_BYTES);
                if (lane == 0) mbar_expect_tx(fb, row_bytes * (uint32_t)bh);
                else if (GS_LANE_ARRIVE) mbar_arrive(fb);
                __syncwarp();
                const int tl = n / ncam, c = n - tl * ncam;
                const int b = ((int)blockIdx.x + tl * (int)gridDim.x) / tiles_fs;
                uint8_t *box = gsm + (size_t)slot * cap_bytes;
                const uint8_t *src = reinterpret_cast<const uint8_t *>(hm) + ((((size_t)b * ncam + c) * hs + y0) * hs + x0) * G_PIX_BYTES;
                const int rowB = pitch * G_PIX_BYTES;
                for (int r = lane; r < bh; r += 32)
                    bulk_load(smem_u32(box + (size_t)r * rowB), src + (size_t)r * hs * G_PIX_BYTES, row_bytes, fb);
            }
            // next item of this producer: slot / header advance by GS_PROD; its corners go into header (n + GS_PROD) % (2 GS_SLOTS),
            // last used by item n + GS_PROD - 2 GS_SLOTS, which was released before item n - GS_SLOTS (waited for above)
            slot += GS_PROD; hidx += GS_PROD;
            if (slot >= GS_SLOTS) { slot -= GS_SLOTS; ph ^= 1u; reuse = true; }
            if (hidx >= 2 * GS_SLOTS) hidx -= 2 * GS_SLOTS;
            if (n + GS_PROD < total) prefetch(n + GS_PROD, hidx);
        }
        return;
    }
    asm volatile("setmaxnreg.inc.sync.aligned.u32 120;");

    // =================================== gather warps ========================================
    const int ci = warp, cj = lane >> 3, z = lane & 7, lk0 = (z + 1) >> 1;
    const int l0 = (ci * GC + cj) * GCK + lk0;
    const uint64_t whalf = f2_pack(0.5f, 0.5f);
    const size_t nv = (size_t)G * G * G;
    int s = 0; uint32_t ph = 0, hsel = 0;                                      // box slot, its phase, header half (0 / GS_SLOTS)
    for (int tl = 0; tl < my_tiles; ++tl) {
        const int t = (int)blockIdx.x + tl * (int)gridDim.x;
        const int b = t / tiles_fs, r0 = t - b * tiles_fs;
        const int tk = r0 % ntk, tj = (r0 / ntk) % nt, ti = r0 / (ntk * nt);
        const int I0 = GT * ti - 1 + 2 * ci, J0 = GT * tj - 1 + 2 * cj, Kz = GT * tk + z;
        // ATen area_pixel_compute_source_index, scale .5: even fine index -> lambda1 .75, odd -> .25, index 0 -> 0
        const float lk1 = Kz == 0 ? 0.f : ((z & 1) ? 0.25f : 0.75f), lk0w = __fsub_rn(1.f, lk1);
        const uint64_t wk0 = f2_pack(lk0w, lk0w), wk1 = f2_pack(lk1, lk1);
        uint64_t wj1[2], wj0[2], wi1[2], wi0[2];
        bool vi[2], vj[2];
#pragma unroll
        for (int v = 0; v < 2; ++v) {
            const float j1 = (J0 + v) == 0 ? 0.f : (v ? 0.75f : 0.25f), j0 = __fsub_rn(1.f, j1);
            const float i1 = (I0 + v) == 0 ? 0.f : (v ? 0.75f : 0.25f), i0 = __fsub_rn(1.f, i1);
            wj1[v] = f2_pack(j1, j1); wj0[v] = f2_pack(j0, j0); wi1[v] = f2_pack(i1, i1); wi0[v] = f2_pack(i0, i0);
            vi[v] = (I0 + v) >= 0 && (I0 + v) < G; vj[v] = (J0 + v) >= 0 && (J0 + v) < G;
        }
        const bool any = (vi[0] || vi[1]) && (vj[0] || vj[1]);
        __half2 acc[4][KP / 2];
#pragma unroll
        for (int v = 0; v < 4; ++v)
#pragma unroll
            for (int i = 0; i < KP / 2; ++i) acc[v][i] = __float2half2_rn(0.f);

        for (int c = 0; c < ncam; ++c) {
            mbar_wait(smem_u32(full + s), ph);
            if (any) {
                const uint8_t *hd = hdrs + (s + (int)hsel) * GS_HDR;
                const uint64_t *Cn = reinterpret_cast<const uint64_t *>(hd) + l0;                   // (x, y) corner pairs
                const int4 bx = *reinterpret_cast<const int4 *>(hd + GS_META);
                const int fits = bx.w, rowB = (fits ? bx.w : hs) * G_PIX_BYTES;
                const int baseB = fits ? -(bx.y * bx.w + bx.x) * G_PIX_BYTES : 0;
                uint64_t xk[2][2];
#pragma unroll
                for (int p = 0; p < 2; ++p)
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const int l = (p * GC + q) * GCK;
                        xk[p][q] = lerp2<MODE>(wk0, Cn[l], wk1, Cn[l + 1]);
                    }
                int off[4];
#pragma unroll
                for (int jv = 0; jv < 2; ++jv) {
                    uint64_t yj[2];
#pragma unroll
                    for (int p = 0; p < 2; ++p) yj[p] = lerp2<MODE>(wj0[jv], xk[p][0], wj1[jv], xk[p][1]);
#pragma unroll
                    for (int iv = 0; iv < 2; ++iv) {
                        float fa, fb2;
                        f2_unpack(f2_mul(lerp2<MODE>(wi0[iv], yj[0], wi1[iv], yj[1]), whalf), fa, fb2);
                        const int px = __float2int_rz(fa), py = __float2int_rz(fb2);                // (val/2).int()   repro_layer.py:82-83
                        off[iv * 2 + jv] = py * rowB + (px * G_PIX_BYTES + baseB);
                    }
                }
                uint4 w[4][3];
                if (fits) {
                    // no per-voxel predicate here: a voxel of this lane that lies outside the grid (x / y edge tiles) still has
                    // clamped corners, so its offset is inside the box; what it accumulates is never stored.  (The predicated
                    // form cost 48 register-zeroing instructions per camera step.)
                    const uint8_t *box = gsm + bx.z;
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        const uint4 *pp = reinterpret_cast<const uint4 *>(box + off[v]);
                        w[v][0] = pp[0]; w[v][1] = pp[1]; w[v][2] = pp[2];
                    }
                } else {
                    const uint8_t *gbase = reinterpret_cast<const uint8_t *>(hm) + ((size_t)b * ncam + c) * hs * hs * G_PIX_BYTES;
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        if (vi[v >> 1] && vj[v & 1]) {
                            const uint4 *pp = reinterpret_cast<const uint4 *>(gbase + off[v]);
                            w[v][0] = __ldg(pp); w[v][1] = __ldg(pp + 1); w[v][2] = __ldg(pp + 2);
                        } else {
                            w[v][0] = w[v][1] = w[v][2] = make_uint4(0, 0, 0, 0);
                        }
                    }
                }
#pragma unroll
                for (int v = 0; v < 4; ++v)
#pragma unroll
                    for (int g = 0; g < 3; ++g) {
                        const uint32_t ww[4] = {w[v][g].x, w[v][g].y, w[v][g].z, w[v][g].w};
#pragma unroll
                        for (int i = 0; i < 4; ++i) acc[v][4 * g + i] = __hadd2(acc[v][4 * g + i], *reinterpret_cast<const __half2 *>(&ww[i]));
                    }
            }
            // release of the slot and of the older header: with GS_LANE_ARRIVE every lane of the four gather warps arrives itself
            // (count 128), i.e. each lane's own reads precede its own arrive — the form racecheck can follow
            if (GS_LANE_ARRIVE) mbar_arrive(smem_u32(empty + s));
            else { __syncwarp(); if (lane == 0) mbar_arrive(smem_u32(empty + s)); }
            if (++s == GS_SLOTS) { s = 0; ph ^= 1u; hsel ^= GS_SLOTS; }
        }

        // ---- mean over cameras (+ /255) as one fp32 scale, store ---------------------------------------
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const int I = I0 + (v >> 1), J = J0 + (v & 1);
            if (!(vi[v >> 1] && vj[v & 1])) continue;
            float m[KP];
#pragma unroll
            for (int i = 0; i < KP / 2; ++i) {
                const float2 f = __half22float2(acc[v][i]);
                m[2 * i] = f.x * post_scale; m[2 * i + 1] = f.y * post_scale;
            }
            if (LAYOUT == JHN_VOL_NCDHW_F32) {
                float *out = (float *)out_ + (size_t)b * K * nv + ((size_t)I * G + J) * G + Kz;
#pragma unroll
                for (int k = 0; k < KP; ++k)
                    if (k < K) out[(size_t)k * nv] = m[k];
            } else {
                const int CJ = (K + 15) / 16 * 2, Wh = G / 2 + 2;
                const int sv = ((I & 1) * 2 + (J & 1)) * 2 + (Kz & 1);
                uint4 *out = (uint4 *)out_;
                const size_t pos = ((size_t)(I >> 1) + 1) * Wh * Wh + (size_t)((J >> 1) + 1) * Wh + (Kz >> 1) + 1;
                const size_t chunk_stride = (size_t)Wh * Wh * Wh;
                const size_t ob = (((size_t)b * 8 + sv) * CJ) * chunk_stride + pos;
#pragma unroll
                for (int j = 0; j < KP / 8; ++j) {
                    if (j < CJ) {
                        uint32_t pk[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            __nv_bfloat162 h2 = __floats2bfloat162_rn(m[8 * j + 2 * i], m[8 * j + 2 * i + 1]);
                            pk[i] = *reinterpret_cast<uint32_t *>(&h2);
                        }
                        out[ob + (size_t)j * chunk_stride] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    }
                }
                // chunks KP / 8 .. CJ - 1 (channel padding of the first convolution) stay zero: launch_stream clears them
            }
        }
    }
}


#ifndef GS_CAP_V
#define GS_CAP_V 15360
#endif
constexpr int GS_CAP = GS_CAP_V;                                               // bytes per pixel-box slot.  The boxes of an 8^3 tile average 96 pixels
                                                                               // (4.6 KB, max 224 at the Example shape); six slots of 10 KB instead of four of
                                                                               // 15 KB (GS_SLOTS_V=6, GS_CAP_V=10240) were measured SLOWER: 0.726 vs 0.690 ms (run 31)
static std::atomic<int> g_box_limit{GS_CAP};                                   // boxes above this gather from global memory (test hook)
int gather_set_box_bytes(int bytes)
{
    const int v = (bytes <= 0 || bytes > GS_CAP) ? GS_CAP : bytes;
    g_box_limit.store(v, std::memory_order_relaxed);
    return v;
}

template <int LAYOUT>
static int launch_stream(const ReprojectArgs &a, const __half *hm_cl, const float2 *cab, int cap, cudaStream_t st)
{
    const size_t gsmem = (size_t)GS_SLOTS * cap + 2 * GS_SLOTS * GS_HDR + 2 * GS_SLOTS * 8;
    const int nt = a.G / GT + 1, ntk = a.G / GT;
    const long long total = (long long)a.B * nt * nt * ntk;
    if (total * a.ncam > 0x7fffffffLL) return fail(JHN_ERR_SHAPE, "too many gather tiles (%lld)", total);
    int dev = 0, sms = 0;                                                       // 3 resident CTAs per SM
    JHN_CUDA(cudaGetDevice(&dev));
    JHN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int ctas = 3 * sms;
    const int grid = (int)(total < ctas ? total : ctas);
    const float post_scale = HALF_STAGE_UNSCALE / ((float)a.ncam * a.post_divide);
    int box_limit = g_box_limit.load(std::memory_order_relaxed);
    if (box_limit > cap) box_limit = cap;
#define JHN_STREAM(MODE)                                                                                         \
    {                                                                                                            \
        auto kern = gather_stream_kernel<LAYOUT, MODE>;                                                          \
        JHN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsmem));           \
        JHN_LAUNCH("gather_stream_kernel", st,                                                                   \
                   kern<<<grid, GS_THREADS, gsmem, st>>>(hm_cl, cab, a.ncam, a.K, a.hs, a.G, post_scale, cap, \
                                                         box_limit, a.volume_out, (int)total));                              \
        return JHN_OK;                                                                                           \
    }
    if (a.lerp_mode == JHN_LERP_FMA_FIRST) JHN_STREAM(JHN_LERP_FMA_FIRST)
    if (a.lerp_mode == JHN_LERP_FMA_SECOND) JHN_STREAM(JHN_LERP_FMA_SECOND)
    JHN_STREAM(JHN_LERP_NO_FMA)
#undef JHN_STREAM
}


size_t reproject_workspace(int B, int ncam, int K, int hs, int G, int precision)
{
    const int h = G / 2;
    Arena a(nullptr, 0);
    a.take<float2>((size_t)B * ncam * h * h * h);                   // coarse (x, y) pixel coordinates
    const size_t px = (size_t)B * ncam * hs * hs * KP;              // staging copy (not touched when the caller hands JHN_HM_F16_CL)
    if (precision == JHN_FP32) a.take<float>(px); else a.take<__nv_bfloat16>(px);
    a.take<int4>((size_t)B * ncam);                                 // per-camera pixel box of the voxel grid
    return a.off;
}

static size_t gather_smem(int ncam) { return (size_t)ncam * (2 * CS * CS * CS * sizeof(float) + TS * TS * TS * sizeof(int32_t)); }

template <typename T>
static int run_gather(const ReprojectArgs &a, const T *hm_cl, const float2 *cab, cudaStream_t st)
{
    const int nt = cdiv(a.G, TS);
    dim3 grid(nt * nt * nt, a.B);
    const size_t smem = gather_smem(a.ncam);
    if (smem > 200 * 1024) return fail(JHN_ERR_SHAPE, "too many cameras (%d) for the gather tile", a.ncam);
    if (a.layout == JHN_VOL_NCDHW_F32) {
        auto kern = gather_fused_kernel<T, JHN_VOL_NCDHW_F32>;
        JHN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        JHN_LAUNCH("gather_fused_kernel", st,
                   kern<<<grid, 256, smem, st>>>(hm_cl, cab, a.ncam, a.K, a.hs, a.G, a.lerp_mode, a.post_divide,
                                                 a.volume_out, a.index_out));
        return JHN_OK;
    }
    const int CJ = (a.K + 15) / 16 * 2;
    if (!a.borders_valid) JHN_TRY(tc_zero_border_launch(a.volume_out, a.B * 8 * CJ, a.G / 2, st));
    auto kern = gather_fused_kernel<T, JHN_VOL_V2V_BF16>;
    JHN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    JHN_LAUNCH("gather_fused_kernel", st,
               kern<<<grid, 256, smem, st>>>(hm_cl, cab, a.ncam, a.K, a.hs, a.G, a.lerp_mode, a.post_divide,
                                             a.volume_out, a.index_out));
    return JHN_OK;
}

int reproject_launch(const ReprojectArgs &a, void *ws, size_t ws_bytes, cudaStream_t st)
{
    const int h = a.G / 2;
    Arena ar(ws, ws_bytes);
    float2 *cab = ar.take<float2>((size_t)a.B * a.ncam * h * h * h);
    const size_t px = (size_t)a.B * a.ncam * a.hs * a.hs * KP;
    void *hm_cl = (a.precision == JHN_FP32) ? (void *)ar.take<float>(px) : (void *)ar.take<__nv_bfloat16>(px);
    int4 *roi = ar.take<int4>((size_t)a.B * a.ncam);
    if (!ar.ok()) return fail(JHN_ERR_WORKSPACE, "reproject workspace: need %zu bytes, got %zu", ar.off, ws_bytes);
    const bool planar = a.hm_format == JHN_HM_F32_PLANAR;
    if (!planar && (a.precision == JHN_FP32 || !a.padded))
        return fail(JHN_ERR_ARG, "channels-last heat maps are a 16-bit, padded format: they need precision JHN_BF16 and heatmaps_padded = 1");

    // the streaming gather's staging copy covers only each camera's pixel box of the voxel grid (collected by the
    // projection kernel); the other paths relayout whole maps
    const bool stream = a.precision != JHN_FP32 && !a.index_out && a.G % GT == 0;
    const bool want_roi = stream && planar;
    if (want_roi) JHN_CUDA(cudaMemsetAsync(roi, 0x7f, (size_t)a.B * a.ncam * sizeof(int4), st));
    JHN_LAUNCH("coarse_project_kernel", st,
               coarse_project_kernel<<<dim3(cdiv((long long)h * h * h, 256), a.B), 256,
                                       a.ncam * (CP_PARAMS * sizeof(float) + 4 * sizeof(int)), st>>>(
                   a.cam, a.intr, a.dist, a.center3D, a.centerHM, a.B, a.ncam, h, a.spacing, a.hs, cab, want_roi ? (int *)roi : nullptr));
    if (a.precision == JHN_FP32) {
        JHN_TRY(launch_relayout<float>(a, (float *)hm_cl, nullptr, st));
        return run_gather<float>(a, (const float *)hm_cl, cab, st);
    }
    if (stream) {
        // throughput path: fp16 staging copy + streaming gather (the index dump needs the in-order kernel below).
        // JHN_HM_F16_CL is that staging format: gathered in place, no copy at all.
        const __half *src16 = (const __half *)hm_cl;
        if (planar) JHN_TRY(launch_relayout<__half>(a, (__half *)hm_cl, roi, st));
        else if (a.hm_format == JHN_HM_F16_CL) src16 = (const __half *)a.heatmaps;
        else {
            const size_t n16 = px / 8;
            JHN_LAUNCH("relayout_kernel", st,
                       bf16cl_to_f16cl_kernel<<<cdiv((long long)n16, 256), 256, 0, st>>>((const uint4 *)a.heatmaps, (uint4 *)hm_cl, n16));
        }
        if (a.layout == JHN_VOL_NCDHW_F32) return launch_stream<JHN_VOL_NCDHW_F32>(a, src16, cab, GS_CAP, st);
        const int CJ = (a.K + 15) / 16 * 2;
        if (!a.borders_valid) {
            JHN_TRY(tc_zero_border_launch(a.volume_out, a.B * 8 * CJ, a.G / 2, st));
            if (CJ > KP / 8) {                                                 // all-padding channel chunks: zeroed here, never written by the kernel
                const size_t Wh = a.G / 2 + 2, chunk_bytes = Wh * Wh * Wh * 16;
                JHN_CUDA(cudaMemset2DAsync((char *)a.volume_out + (KP / 8) * chunk_bytes, CJ * chunk_bytes, 0,
                                           (CJ - KP / 8) * chunk_bytes, (size_t)a.B * 8, st));
            }
        }
        return launch_stream<JHN_VOL_V2V_BF16>(a, src16, cab, GS_CAP, st);
    }
    if (a.hm_format == JHN_HM_F16_CL) return run_gather<__half>(a, (const __half *)a.heatmaps, cab, st);
    if (a.hm_format == JHN_HM_BF16_CL) return run_gather<__nv_bfloat16>(a, (const __nv_bfloat16 *)a.heatmaps, cab, st);
    JHN_TRY(launch_relayout<__nv_bfloat16>(a, (__nv_bfloat16 *)hm_cl, nullptr, st));
    return run_gather<__nv_bfloat16>(a, (const __nv_bfloat16 *)hm_cl, cab, st);
}

}  // namespace jhn
