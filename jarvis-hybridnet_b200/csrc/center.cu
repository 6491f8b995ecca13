// Predictor glue between the centre-detect CNN and the 3D network.  Replaces, for B frame sets
// and without a single host synchronisation,
//   jarvis/prediction/jarvis3D.py:147-166   per-camera argmax + threshold count, scaling to full-resolution pixels,
//                                           ReprojectionTool.reconstructPoint, reprojectPoint, .int(), clamp
//   jarvis/utils/reprojection.py:45-90      (the two ReprojectionTool methods: weighted DLT + SVD, projection)
//   jarvis/prediction/jarvis3D.py:168-177   the 12-iteration Python crop loop (one device sync per slice) + normalise
// The reference branches on the host (`if num_cams_detect >= 2`, a device sync) and syncs again for every crop
// slice; here the decision is a device flag (`valid`) that the crop kernel and the caller read later, so the whole
// predictor step can be enqueued (or graph-captured) at once.
//
//   center_locate_kernel   grid (ncam, B): block = one camera's heat map -> first index of the maximum; the last
//                          block of a frame set to finish (atomic ticket) triangulates: fp32 DLT rows exactly as
//                          the reference forms them, the 4x4 normal matrix in fp64, cyclic Jacobi for its smallest
//                          eigenvector (= last right singular vector of the reference's torch.linalg.svd), then
//                          the fp32 projection chain, truncation and clamp per camera.
//   crop_normalize_kernel  HBM-bound copy of the ncam bounding boxes, (x - mean) / std as two rounded fp32 ops.
#include "common.cuh"

namespace jhn {

constexpr int CL_THREADS = 256, CL_MAX_CAMS = 64;

__device__ __forceinline__ void jacobi4(double (&A)[4][4], double (&V)[4][4])
{
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) V[i][j] = i == j ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 16; ++sweep) {
        double off = 0.0, diag = 0.0;
        for (int p = 0; p < 4; ++p) {
            diag += A[p][p] * A[p][p];
            for (int q = p + 1; q < 4; ++q) off += A[p][q] * A[p][q];
        }
        if (off <= 1e-30 * diag) break;                               // off-diagonal mass below fp64 rounding of the diagonal
        for (int p = 0; p < 4; ++p)
            for (int q = p + 1; q < 4; ++q) {
                if (A[p][q] == 0.0) continue;
                const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 4; ++k) {                      // A <- A J
                    const double akp = A[k][p], akq = A[k][q];
                    A[k][p] = c * akp - s * akq; A[k][q] = s * akp + c * akq;
                }
                for (int k = 0; k < 4; ++k) {                      // A <- J^T A
                    const double apk = A[p][k], aqk = A[q][k];
                    A[p][k] = c * apk - s * aqk; A[q][k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 4; ++k) {
                    const double vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = c * vkp - s * vkq; V[k][q] = s * vkp + c * vkq;
                }
            }
    }
}

__global__ void __launch_bounds__(CL_THREADS)
center_locate_kernel(const float *__restrict__ hm, int ncam, int Hc, int Wc, float sx2, float sy2, int img_w, int img_h,
                     int bbox_hw, float threshold, const float *__restrict__ cam, const float *__restrict__ intr,
                     const float *__restrict__ dist, int32_t *__restrict__ preds, float *__restrict__ maxvals,
                     float *__restrict__ center3D, int32_t *__restrict__ center3D_int, int32_t *__restrict__ centerHM,
                     int32_t *__restrict__ valid, unsigned int *__restrict__ scratch)
{
    __shared__ float s_val[CL_THREADS / 32];
    __shared__ int s_idx[CL_THREADS / 32];
    __shared__ unsigned int s_ticket;
    __shared__ double s_M[10];
    __shared__ float s_X[3];
    const int c = blockIdx.x, b = blockIdx.y, bc = b * ncam + c;
    const int n = Hc * Wc;
    // ---- first index of the maximum (heatmaps_gpu.argmax(2), jarvis3D.py:149-150) ---------------------------
    const float *src = hm + (size_t)bc * n;
    float best = -INFINITY; int bi = 0x7fffffff;
    for (int i = threadIdx.x; i < n; i += CL_THREADS) {
        const float v = __ldg(src + i);
        if (v > best) { best = v; bi = i; }                          // ascending i per thread: first occurrence kept
    }
#pragma unroll
    for (int sh = 16; sh > 0; sh >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, sh);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, sh);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { s_val[threadIdx.x >> 5] = best; s_idx[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < CL_THREADS / 32; ++w)
            if (s_val[w] > best || (s_val[w] == best && s_idx[w] < bi)) { best = s_val[w]; bi = s_idx[w]; }
        preds[2 * bc + 0] = bi % Hc;                                 // `m % shape[2]`, `m // shape[3]`   :151-152
        preds[2 * bc + 1] = bi / Wc;
        maxvals[bc] = __fdiv_rn(best, 255.f);                        // :155
        if (best > threshold) atomicAdd(scratch + 2 * b + 1, 1u);    // num_cams_detect            :154
        __threadfence();
        s_ticket = atomicAdd(scratch + 2 * b, 1u);
    }
    __syncthreads();
    if (s_ticket != (unsigned)(ncam - 1)) return;

    // ---- last block of this frame set: triangulate + reproject ------------------------------------------------
    __threadfence();
    if (threadIdx.x < 10) s_M[threadIdx.x] = 0.0;
    __syncthreads();
    const int t = threadIdx.x;
    if (t < ncam) {
        const int tc = b * ncam + t;
        const float *P = cam + 12 * tc;                              // [4][3]
        const float fx = intr[9 * tc + 0], fy = intr[9 * tc + 4], cx = intr[9 * tc + 6], cy = intr[9 * tc + 7];
        const float k1 = dist[5 * tc + 0], k2 = dist[5 * tc + 1];
        const float w = __ldcg(maxvals + tc);
        // preds * (downsampling_scale * 2), then ReprojectionTool.reconstructPoint            reprojection.py:69-84
        float x = __fsub_rn(__fmul_rn((float)__ldcg(preds + 2 * tc + 0), sx2), cx);
        float y = __fsub_rn(__fmul_rn((float)__ldcg(preds + 2 * tc + 1), sy2), cy);
        const float ax = __fdiv_rn(x, fx), ay = __fdiv_rn(y, fy);
        const float r2 = __fadd_rn(__fmul_rn(ax, ax), __fmul_rn(ay, ay));
        const float d = __fadd_rn(1.f, __fmul_rn(__fadd_rn(k1, __fmul_rn(k2, r2)), r2));
        x = __fadd_rn(__fdiv_rn(x, d), cx);
        y = __fadd_rn(__fdiv_rn(y, d), cy);
        float a0[4], a1[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {                                // row r: p_r * P[:,2] - P[:,r], times maxvals
            a0[k] = __fmul_rn(__fsub_rn(__fmul_rn(x, P[3 * k + 2]), P[3 * k + 0]), w);
            a1[k] = __fmul_rn(__fsub_rn(__fmul_rn(y, P[3 * k + 2]), P[3 * k + 1]), w);
        }
        int e = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = i; j < 4; ++j)
                atomicAdd(&s_M[e++], (double)a0[i] * (double)a0[j] + (double)a1[i] * (double)a1[j]);
    }
    __syncthreads();
    if (t == 0) {
        double A[4][4], V[4][4];
        int e = 0;
        for (int i = 0; i < 4; ++i)
            for (int j = i; j < 4; ++j) { A[i][j] = A[j][i] = s_M[e++]; }
        jacobi4(A, V);
        int m = 0;
        for (int i = 1; i < 4; ++i)
            if (A[i][i] < A[m][m]) m = i;
        const unsigned int ndet = __ldcg(scratch + 2 * b + 1);
        scratch[2 * b] = 0u; scratch[2 * b + 1] = 0u;                // ready for the next launch
        const bool ok = ndet >= 2u;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const float X = ok ? (float)(V[i][m] / V[3][m]) : 0.f;   // X / X[-1]                 reprojection.py:88-90
            s_X[i] = X;
            center3D[3 * b + i] = X;
            center3D_int[3 * b + i] = ok ? __float2int_rz(X) : 0;     // center3D.int()            jarvis3D.py:183
        }
        valid[b] = ok ? 1 : 0;
    }
    __syncthreads();
    if (t < ncam) {
        const int tc = b * ncam + t;
        const float *P = cam + 12 * tc;
        const float fx = intr[9 * tc + 0], fy = intr[9 * tc + 4], cx = intr[9 * tc + 6], cy = intr[9 * tc + 7];
        const float k1 = dist[5 * tc + 0], k2 = dist[5 * tc + 1];
        float uvw[3];
#pragma unroll
        for (int q = 0; q < 3; ++q) {                                // [X,1] @ P as the GEMM's FMA chain   reprojection.py:45-49
            float s = __fmul_rn(s_X[0], P[q]);
            s = __fmaf_rn(s_X[1], P[3 + q], s);
            s = __fmaf_rn(s_X[2], P[6 + q], s);
            s = __fmaf_rn(1.f, P[9 + q], s);
            uvw[q] = s;
        }
        float a = __fsub_rn(__fdiv_rn(uvw[0], uvw[2]), cx), bb = __fsub_rn(__fdiv_rn(uvw[1], uvw[2]), cy);
        const float ax = __fdiv_rn(a, fx), ay = __fdiv_rn(bb, fy);
        const float r2 = __fadd_rn(__fmul_rn(ax, ax), __fmul_rn(ay, ay));
        const float d = __fadd_rn(1.f, __fmul_rn(__fadd_rn(k1, __fmul_rn(k2, r2)), r2));
        a = __fadd_rn(__fmul_rn(a, d), cx);
        bb = __fadd_rn(__fmul_rn(bb, d), cy);
        const bool ok = valid[b] != 0;                                // written by thread 0 of this block before the barrier
        int px = ok ? __float2int_rz(a) : bbox_hw, py = ok ? __float2int_rz(bb) : bbox_hw;   // .int()   jarvis3D.py:161-162
        px = min(max(px, bbox_hw), img_w - bbox_hw);                  // :163-166
        py = min(max(py, bbox_hw), img_h - bbox_hw);
        centerHM[2 * tc + 0] = px; centerHM[2 * tc + 1] = py;
    }
}

// out[b][c][ch][y][x] = (imgs[b][c][ch][cy - hw + y][cx - hw + x] - mean[ch]) / std[ch]; zeros when !valid[b]
__global__ void __launch_bounds__(256)
crop_normalize_kernel(const float *__restrict__ imgs, int H, int W, int bbox, const int32_t *__restrict__ centerHM,
                      const int32_t *__restrict__ valid, int ncam, float m0, float m1, float m2, float s0, float s1,
                      float s2, float *__restrict__ out)
{
    const int bc = blockIdx.z, ch = blockIdx.y, hw = bbox / 2;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;             // one thread = 4 consecutive x of one row
    const int per_row = bbox / 4;
    if (q >= bbox * per_row) return;
    const int y = q / per_row, x4 = (q - y * per_row) * 4;
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid[bc / ncam]) {
        const int cx = centerHM[2 * bc + 0], cy = centerHM[2 * bc + 1];
        const float *src = imgs + (((size_t)bc * 3 + ch) * H + (cy - hw + y)) * W + (cx - hw + x4);
        const float m = ch == 0 ? m0 : ch == 1 ? m1 : m2, s = ch == 0 ? s0 : ch == 1 ? s1 : s2;
        r.x = __fdiv_rn(__fsub_rn(__ldg(src + 0), m), s); r.y = __fdiv_rn(__fsub_rn(__ldg(src + 1), m), s);
        r.z = __fdiv_rn(__fsub_rn(__ldg(src + 2), m), s); r.w = __fdiv_rn(__fsub_rn(__ldg(src + 3), m), s);
    }
    reinterpret_cast<float4 *>(out + (((size_t)bc * 3 + ch) * bbox + y) * bbox)[x4 / 4] = r;
}

int center_locate_launch(const float *hm, int B, int ncam, int Hc, int Wc, int img_w, int img_h, int cdis, int bbox_hw,
                         float threshold, const float *cam, const float *intr, const float *dist, int32_t *preds,
                         float *maxvals, float *center3D, int32_t *center3D_int, int32_t *centerHM, int32_t *valid,
                         void *scratch, cudaStream_t st)
{
    // downsampling_scale = tensor([W / float(cdis), H / float(cdis)]).float(); preds * (downsampling_scale * 2)   jarvis3D.py:135-138,158
    const float sx2 = (float)((double)img_w / (double)cdis) * 2.f, sy2 = (float)((double)img_h / (double)cdis) * 2.f;
    JHN_LAUNCH("center_locate_kernel", st,
               center_locate_kernel<<<dim3(ncam, B), CL_THREADS, 0, st>>>(hm, ncam, Hc, Wc, sx2, sy2, img_w, img_h, bbox_hw, threshold,
                                                                           cam, intr, dist, preds, maxvals, center3D, center3D_int,
                                                                           centerHM, valid, (unsigned int *)scratch));
    return JHN_OK;
}

int crop_normalize_launch(const float *imgs, int B, int ncam, int H, int W, int bbox, const int32_t *centerHM,
                          const int32_t *valid, const float *mean, const float *std, float *out, cudaStream_t st)
{
    const int threads_needed = bbox * (bbox / 4);
    JHN_LAUNCH("crop_normalize_kernel", st,
               crop_normalize_kernel<<<dim3(cdiv(threads_needed, 256), 3, B * ncam), 256, 0, st>>>(
                   imgs, H, W, bbox, centerHM, valid, ncam, mean[0], mean[1], mean[2], std[0], std[1], std[2], out));
    return JHN_OK;
}

}  // namespace jhn
