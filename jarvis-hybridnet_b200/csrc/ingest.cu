// Frame ingest and the returned volumes of HybridNetBackbone.forward (row a11).
//
//   ingest_frames_kernel      jarvis/prediction/predict3D.py:79  `from_numpy(imgs_orig).cuda().float().permute(0,3,1,2)[:, [2,1,0]] / 255.`
//                             decoded frames cross PCIe as the decoder wrote them (uint8, H x W x BGR: 3 B per pixel instead of
//                             the 12 B of the fp32 tensor the reference uploads) and become the reference's fp32 CHW RGB tensor
//                             on the device.  ATen's CUDA division by a Python scalar multiplies by the fp32 reciprocal
//                             (BinaryDivTrueKernel.cu: `a * (1 / b)`), so the reference's tensor is float(u8) * (1.f / 255.f);
//                             the kernel produces those bits (they differ from the IEEE quotient in the last place for
//                             some byte values).
//   crop_normalize_u8_kernel  jarvis/prediction/jarvis3D.py:168-177 straight from the uint8 frames: the key-point detector's
//                             crops ((u8 * (1/255) - mean) / std, three separately rounded fp32 ops as the reference's CUDA path) without
//                             ever materialising the 12 x 3 x 1024 x 1280 fp32 image in HBM.
//   softplus2_kernel          jarvis/hybridnet/model.py:73,88   heatmap_final = softplus(softplus(v2v))  (the returned volume)
//   pad_border_kernel         jarvis/hybridnet/model.py:65-66   heatmaps_padded = F.pad(heatmaps, [1,1,1,1])
// All four are HBM-bound streaming kernels: 16-byte accesses, one pass.
#include <algorithm>
#include <atomic>

#include "common.cuh"

namespace jhn {

// one thread = 4 consecutive pixels of one row: 12 input bytes (three aligned 32-bit loads), one float4 per colour plane
__global__ void __launch_bounds__(256)
ingest_frames_kernel(const uint8_t *__restrict__ frames, int H, int W, long long quads, float *__restrict__ out)
{
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= quads) return;
    const int per_row = W / 4;
    const long long row = q / per_row;                                // image * H + y
    const int x4 = (int)(q - row * per_row) * 4;
    const long long img = row / H;
    const int y = (int)(row - img * H);
    const uint32_t *src = reinterpret_cast<const uint32_t *>(frames + (size_t)row * W * 3 + (size_t)x4 * 3);
    const uint32_t w0 = __ldg(src), w1 = __ldg(src + 1), w2 = __ldg(src + 2);   // B0 G0 R0 B1 | G1 R1 B2 G2 | R2 B3 G3 R3
    const float b0 = (float)(w0 & 0xffu), g0 = (float)((w0 >> 8) & 0xffu), r0 = (float)((w0 >> 16) & 0xffu), b1 = (float)(w0 >> 24);
    const float g1 = (float)(w1 & 0xffu), r1 = (float)((w1 >> 8) & 0xffu), b2 = (float)((w1 >> 16) & 0xffu), g2 = (float)(w1 >> 24);
    const float r2 = (float)(w2 & 0xffu), b3 = (float)((w2 >> 8) & 0xffu), g3 = (float)((w2 >> 16) & 0xffu), r3 = (float)(w2 >> 24);
    const size_t plane = (size_t)H * W;
    float *dst = out + (size_t)img * 3 * plane + (size_t)y * W + x4;
    const float d = __fdiv_rn(1.f, 255.f);                           // ATen CUDA `x / 255.` == x * (1.f / 255.f)
    *reinterpret_cast<float4 *>(dst) = make_float4(__fmul_rn(r0, d), __fmul_rn(r1, d), __fmul_rn(r2, d), __fmul_rn(r3, d));
    *reinterpret_cast<float4 *>(dst + plane) = make_float4(__fmul_rn(g0, d), __fmul_rn(g1, d), __fmul_rn(g2, d), __fmul_rn(g3, d));
    *reinterpret_cast<float4 *>(dst + 2 * plane) = make_float4(__fmul_rn(b0, d), __fmul_rn(b1, d), __fmul_rn(b2, d), __fmul_rn(b3, d));
}

// out[b][c][ch][y][x] = ((frames[b][c][cy - hw + y][cx - hw + x][2 - ch] / 255) - mean[ch]) / std[ch]; zeros when !valid[b].
// One thread = 4 consecutive x of one crop row, all three colour planes (the window start is not 4-byte aligned: byte loads
// through the read-only path, 12 per thread, each cache line used by the neighbouring lanes).
__global__ void __launch_bounds__(256)
crop_normalize_u8_kernel(const uint8_t *__restrict__ frames, int H, int W, int bbox, const int32_t *__restrict__ centerHM,
                         const int32_t *__restrict__ valid, int ncam, float m0, float m1, float m2, float s0, float s1,
                         float s2, float *__restrict__ out)
{
    const int bc = blockIdx.y, hw = bbox / 2;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    const int per_row = bbox / 4;
    if (q >= bbox * per_row) return;
    const int y = q / per_row, x4 = (q - y * per_row) * 4;
    float4 r[3];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) r[ch] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid[bc / ncam]) {
        const int cx = centerHM[2 * bc + 0], cy = centerHM[2 * bc + 1];
        const uint8_t *src = frames + (((size_t)bc * H + (cy - hw + y)) * W + (cx - hw + x4)) * 3;
        float v[12];
#pragma unroll
        const float d = __fdiv_rn(1.f, 255.f);
        for (int i = 0; i < 12; ++i) v[i] = __fmul_rn((float)__ldg(src + i), d);          // pixel i / 3, BGR channel i % 3
        const float m[3] = {m0, m1, m2}, s[3] = {s0, s1, s2};
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {                              // output plane ch = RGB, input byte 2 - ch
            r[ch].x = __fdiv_rn(__fsub_rn(v[0 + 2 - ch], m[ch]), s[ch]); r[ch].y = __fdiv_rn(__fsub_rn(v[3 + 2 - ch], m[ch]), s[ch]);
            r[ch].z = __fdiv_rn(__fsub_rn(v[6 + 2 - ch], m[ch]), s[ch]); r[ch].w = __fdiv_rn(__fsub_rn(v[9 + 2 - ch], m[ch]), s[ch]);
        }
    }
#pragma unroll
    for (int ch = 0; ch < 3; ++ch)
        reinterpret_cast<float4 *>(out + (((size_t)bc * 3 + ch) * bbox + y) * bbox)[x4 / 4] = r[ch];
}

// torch.nn.Softplus (beta 1, threshold 20), applied twice: x > 20 ? x : log1p(exp(x)) with CUDA libm, as ATen's kernel does
__device__ __forceinline__ float softplus1(float x) { return x > 20.f ? x : log1pf(expf(x)); }

__global__ void __launch_bounds__(256)
softplus2_kernel(const float *__restrict__ v, long long n, float *__restrict__ out)
{
    const long long i4 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i4 + 3 < n) {
        const float4 a = __ldg(reinterpret_cast<const float4 *>(v + i4));
        *reinterpret_cast<float4 *>(out + i4) = make_float4(softplus1(softplus1(a.x)), softplus1(softplus1(a.y)),
                                                            softplus1(softplus1(a.z)), softplus1(softplus1(a.w)));
    } else {
        for (long long i = i4; i < n; ++i) out[i] = softplus1(softplus1(v[i]));
    }
}

// [N][S][S] -> [N][S+2][S+2] with a zero border; one thread per output element (rows of S+2 floats are not 16-byte aligned)
__global__ void __launch_bounds__(256)
pad_border_kernel(const float *__restrict__ in, int S, long long n_out, float *__restrict__ out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_out) return;
    const int hs = S + 2;
    const long long img = i / (hs * hs);
    const int r = (int)(i - img * hs * hs), y = r / hs, x = r - y * hs;
    const bool inside = y >= 1 && y <= S && x >= 1 && x <= S;
    out[i] = inside ? __ldg(in + (size_t)img * S * S + (size_t)(y - 1) * S + (x - 1)) : 0.f;
}

// Host -> device transfer of only the pixels the gather reads, executed by the SMs: the channels-last heat maps stay in
// pinned (device-mapped) host memory and this kernel reads each image's pixel box over PCIe and writes it to the same place
// of the device tensor.  A copy engine walks a strided (2-D) copy at ~0.16 us per row — 36 GB/s on these 5 KB rows, less
// than copying the whole tensor contiguously — whereas coalesced 16-byte loads from mapped memory run at link speed for
// any row length, and the boxes never have to visit the host (jhn_heatmap_boxes leaves them on the device).
// A SMALL persistent grid (PULL_CTAS CTAs of 128 threads x 32 registers, no shared memory) walks the work items
// (image, part of its box), 2 loads in flight per thread: ~200 KB in flight cover the link's bandwidth-delay product, and
// one such CTA fits next to the persistent convolution / gather CTAs (they leave 4096 registers and, with
// JHN_SMEM_RESERVE, 3 KB of shared memory per SM free), so the transfer of chunk i+1 overlaps the kernels of chunk i
// without ever keeping a compute CTA off an SM — a grid with one CTA per work item did exactly that.
// (A row-granular variant — a warp per box row, no division per unit, ~5 instead of ~40 instructions per 16 bytes — slowed the
// co-resident convolution CTAs just as much and made the hybrid upload slower, 4.7 vs 4.2 ms per step: run 48.)
constexpr int PULL_MAX_THREADS = 128, PULL_UNROLL = 2;
// Launch shape, settable at run time (jhn_debug_set_pull_config): threads per CTA (32..128), CTAs, parts per image.
static std::atomic<int> g_pull_threads{128}, g_pull_ctas{48}, g_pull_split{8};
void pull_set_config(int threads, int ctas, int split)
{
    if (threads >= 32 && threads <= PULL_MAX_THREADS && threads % 32 == 0) g_pull_threads.store(threads, std::memory_order_relaxed);
    if (ctas >= 1) g_pull_ctas.store(ctas, std::memory_order_relaxed);
    if (split >= 1 && split <= 64) g_pull_split.store(split, std::memory_order_relaxed);
}

__global__ void __launch_bounds__(PULL_MAX_THREADS, 16)
pull_boxes_kernel(const uint4 *__restrict__ host, uint4 *__restrict__ dev, const int4 *__restrict__ boxes, int n_images, int hs,
                  int units_per_pixel, int split, unsigned long long *__restrict__ bytes_out)
{
    const int nthr = blockDim.x;
    for (int w = blockIdx.x; w < n_images * split; w += gridDim.x) {
        const int img = w / split, part = w - img * split;
        const int4 bx = __ldg(boxes + img);                           // {x0, y0, -x1, -y1}
        const int x0 = bx.x, y0 = bx.y, bw = -bx.z - bx.x + 1, bh = -bx.w - bx.y + 1;
        if (x0 < 0 || y0 < 0 || bw < 1 || bh < 1 || x0 + bw > hs || y0 + bh > hs) continue;  // no box: nothing the gather could read
        const int row_units = bw * units_per_pixel, pitch_units = hs * units_per_pixel;
        const int total = row_units * bh;                             // < 2^31: one image
        const size_t base = (size_t)img * hs * pitch_units + (size_t)y0 * pitch_units + (size_t)x0 * units_per_pixel;
        const uint4 *src = host + base;
        uint4 *dst = dev + base;
        const int stride = nthr * split;
        for (int u0 = part * nthr + threadIdx.x; u0 < total; u0 += stride * PULL_UNROLL) {
            uint4 v[PULL_UNROLL];
            int off[PULL_UNROLL];
#pragma unroll
            for (int k = 0; k < PULL_UNROLL; ++k) {
                const int u = u0 + k * stride;
                const int r = u / row_units;
                off[k] = u + r * (pitch_units - row_units);           // r * pitch + (u - r * row_units)
                if (u < total) v[k] = __ldcs(src + off[k]);           // streaming: each byte crosses the link once
            }
#pragma unroll
            for (int k = 0; k < PULL_UNROLL; ++k)
                if (u0 + k * stride < total) dst[off[k]] = v[k];
        }
        if (bytes_out && part == 0 && threadIdx.x == 0) atomicAdd(bytes_out, (unsigned long long)total * 16ull);
    }
}

// An SM's shared-memory / L1 split is reconfigured only while the SM is idle.  A kernel without shared memory prefers the
// smallest carve-out, so a pull CTA sitting alone on an SM pins that configuration — and the next persistent convolution CTA,
// which needs nearly all of the SM's shared memory, must wait until the pull CTA has left: every persistent kernel of the
// forward then runs in two waves (tools/coreside_probe.py with a sleeping dummy kernel: forward 2.98 -> 4.8 ms next to 48
// idle CTAs, 3.06 ms once the dummy prefers the largest carve-out; profiles/r02_e2e_hybrid_upload.txt, runs 49 - 51).
// The transfer kernels therefore ask for the LARGEST carve-out although they use no shared memory.
template <typename F>
static int prefer_max_carveout(F kern)
{
    static std::atomic<unsigned long long> configured{0ull};          // one bit per device
    int dev = 0;
    JHN_CUDA(cudaGetDevice(&dev));
    const unsigned long long bit = 1ull << (dev & 63);
    if (!(configured.load(std::memory_order_acquire) & bit)) {
        JHN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
        configured.fetch_or(bit, std::memory_order_release);
    }
    return JHN_OK;
}

int pull_boxes_launch(const void *host_mapped, void *dev, const int32_t *boxes, int n_images, int hs, int pixel_bytes,
                      unsigned long long *bytes_out, cudaStream_t st)
{
    JHN_TRY(prefer_max_carveout(pull_boxes_kernel));
    const int threads = g_pull_threads.load(std::memory_order_relaxed), split = g_pull_split.load(std::memory_order_relaxed);
    const int ctas = std::min(g_pull_ctas.load(std::memory_order_relaxed), n_images * split);
    JHN_LAUNCH("pull_boxes_kernel", st,
               pull_boxes_kernel<<<ctas, threads, 0, st>>>((const uint4 *)host_mapped, (uint4 *)dev, (const int4 *)boxes, n_images, hs,
                                                           pixel_bytes / 16, split, bytes_out));
    return JHN_OK;
}

// The same transfer restricted to each pixel row's column span (jhn_heatmap_spans): a warp per row, work item = (image, part).
__global__ void __launch_bounds__(PULL_MAX_THREADS, 16)
pull_spans_kernel(const uint4 *__restrict__ host, uint4 *__restrict__ dev, const int2 *__restrict__ spans, int n_images, int hs,
                  int units_per_pixel, int split, unsigned long long *__restrict__ bytes_out)
{
    const int nw = blockDim.x >> 5, wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const ptrdiff_t to_dev = reinterpret_cast<const char *>(dev) - reinterpret_cast<const char *>(host);
    const int pitch_units = hs * units_per_pixel;
    unsigned long long moved = 0;
    for (int w = blockIdx.x; w < n_images * split; w += gridDim.x) {
        const int img = w / split, part = w - img * split;
        for (int y = part * nw + wid; y < hs; y += split * nw) {
            const int2 sp = __ldg(spans + (size_t)img * hs + y);      // {lo, -hi}
            const int lo = sp.x, hi = -sp.y;
            if (lo < 0 || hi >= hs || hi < lo) continue;              // no voxel maps to this row
            const int units = (hi - lo + 1) * units_per_pixel;
            const uint4 *src = host + ((size_t)img * hs + y) * pitch_units + (size_t)lo * units_per_pixel + lane;
            for (int left = units - lane; left > 0; left -= 32 * PULL_UNROLL, src += 32 * PULL_UNROLL) {
                uint4 v[PULL_UNROLL];
#pragma unroll
                for (int k = 0; k < PULL_UNROLL; ++k)
                    if (32 * k < left) v[k] = __ldcs(src + 32 * k);
                uint4 *dst = reinterpret_cast<uint4 *>(reinterpret_cast<char *>(const_cast<uint4 *>(src)) + to_dev);
#pragma unroll
                for (int k = 0; k < PULL_UNROLL; ++k)
                    if (32 * k < left) dst[32 * k] = v[k];
            }
            moved += (unsigned long long)units * 16ull;
        }
    }
    if (bytes_out && lane == 0 && moved) atomicAdd(bytes_out, moved);
}

// Maps of up to PULL_FLAT_HS rows: the spans of an image are flattened into one run of 16-byte units (prefix sums of the row
// widths in 1.5 KB of shared memory — what a convolution CTA leaves free on its SM) and the threads stride through that run
// exactly as pull_boxes_kernel strides through a box: two independent loads in flight per thread, no partly filled warps at
// row ends.  (The warp-per-row kernel above loses ~25 % of the link to those: 4.45 vs 4.11 ms per step, run 53.)
constexpr int PULL_FLAT_HS = 256;
__global__ void __launch_bounds__(PULL_MAX_THREADS, 16)
pull_spans_flat_kernel(const uint4 *__restrict__ host, uint4 *__restrict__ dev, const int2 *__restrict__ spans, int n_images, int hs,
                       int units_per_pixel, int split, unsigned long long *__restrict__ bytes_out)
{
    __shared__ int pre[PULL_FLAT_HS + 1];                             // pixels in the rows before row y
    __shared__ short lo_s[PULL_FLAT_HS];
    const int nthr = blockDim.x, lane = threadIdx.x & 31;
    const int per = (hs + 31) / 32;                                   // rows per lane of the scanning warp
    for (int w = blockIdx.x; w < n_images * split; w += gridDim.x) {
        const int img = w / split, part = w - img * split;
        __syncthreads();                                              // the previous item's tables are no longer read
        for (int y = threadIdx.x; y < hs; y += nthr) {
            const int2 sp = __ldg(spans + (size_t)img * hs + y);      // {lo, -hi}
            const int lo = sp.x, hi = -sp.y;
            const bool ok = lo >= 0 && hi < hs && hi >= lo;
            pre[y + 1] = ok ? hi - lo + 1 : 0;
            lo_s[y] = (short)(ok ? lo : 0);
        }
        __syncthreads();
        if (threadIdx.x < 32) {                                       // exclusive scan: a chunk of rows per lane, then the lanes
            int sum = 0;
            for (int i = 0; i < per; ++i) { const int y = lane * per + i; if (y < hs) sum += pre[y + 1]; }
            int incl = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
            int run = incl - sum;
            for (int i = 0; i < per; ++i) { const int y = lane * per + i; if (y < hs) { const int wd = pre[y + 1]; pre[y + 1] = run + wd; run += wd; } }
            if (lane == 0) pre[0] = 0;
        }
        __syncthreads();
        const int total = pre[hs] * units_per_pixel;
        const size_t base = (size_t)img * hs * hs * units_per_pixel;
        const uint4 *src = host + base;
        uint4 *dst = dev + base;
        const int stride = nthr * split;
        for (int u0 = part * nthr + threadIdx.x; u0 < total; u0 += stride * PULL_UNROLL) {
            uint4 v[PULL_UNROLL];
            int off[PULL_UNROLL];
#pragma unroll
            for (int k = 0; k < PULL_UNROLL; ++k) {
                const int u = u0 + k * stride;
                const int px = u / units_per_pixel;
                int a = 0, b = hs;                                    // last row with pre[row] <= px
                while (b - a > 1) { const int m = (a + b) >> 1; if (pre[m] <= px) a = m; else b = m; }
                off[k] = (a * hs + lo_s[a] - pre[a]) * units_per_pixel + u;
                if (u < total) v[k] = __ldcs(src + off[k]);
            }
#pragma unroll
            for (int k = 0; k < PULL_UNROLL; ++k)
                if (u0 + k * stride < total) dst[off[k]] = v[k];
        }
        if (bytes_out && part == 0 && threadIdx.x == 0) atomicAdd(bytes_out, (unsigned long long)total * 16ull);
    }
}

int pull_spans_launch(const void *host_mapped, void *dev, const int32_t *spans, int n_images, int hs, int pixel_bytes,
                      unsigned long long *bytes_out, cudaStream_t st)
{
    JHN_TRY(prefer_max_carveout(pull_spans_kernel));
    const int threads = g_pull_threads.load(std::memory_order_relaxed), split = g_pull_split.load(std::memory_order_relaxed);
    const int ctas = std::min(g_pull_ctas.load(std::memory_order_relaxed), n_images * split);
    if (hs <= PULL_FLAT_HS) {
        JHN_TRY(prefer_max_carveout(pull_spans_flat_kernel));
        JHN_LAUNCH("pull_spans_flat_kernel", st,
                   pull_spans_flat_kernel<<<ctas, threads, 0, st>>>((const uint4 *)host_mapped, (uint4 *)dev, (const int2 *)spans, n_images, hs,
                                                                    pixel_bytes / 16, split, bytes_out));
        return JHN_OK;
    }
    JHN_LAUNCH("pull_spans_kernel", st,
               pull_spans_kernel<<<ctas, threads, 0, st>>>((const uint4 *)host_mapped, (uint4 *)dev, (const int2 *)spans, n_images, hs,
                                                           pixel_bytes / 16, split, bytes_out));
    return JHN_OK;
}

// A handful of small host tensors (calibration, centres: ~1.4 KB per frame set) read out of mapped host memory by ONE small
// kernel.  Through cudaMemcpyAsync they would queue on the host->device copy engine BEHIND the previous step's heat-map
// transfer (a copy engine serves its streams in issue order), and the step's own transfer — which needs the pixel boxes computed
// from them — could not start until that transfer had drained: ~0.45 ms of idle link per step (profiles/r02_e2e_timeline.txt).
struct PullSegs { int n; const uint32_t *src[8]; uint32_t *dst[8]; unsigned words[8]; };
__global__ void __launch_bounds__(128)
pull_segments_kernel(const __grid_constant__ PullSegs S)
{
    for (int k = 0; k < S.n; ++k)
        for (unsigned i = blockIdx.x * 128u + threadIdx.x; i < S.words[k]; i += gridDim.x * 128u) S.dst[k][i] = __ldcs(S.src[k] + i);
}

int pull_segments_launch(int n, const void *const *src_mapped, void *const *dst, const size_t *bytes, cudaStream_t st)
{
    PullSegs S{};
    S.n = n;
    size_t most = 0;
    for (int k = 0; k < n; ++k) {
        S.src[k] = (const uint32_t *)src_mapped[k]; S.dst[k] = (uint32_t *)dst[k]; S.words[k] = (unsigned)(bytes[k] / 4);
        if (bytes[k] > most) most = bytes[k];
    }
    const int ctas = (int)std::min<size_t>(16, std::max<size_t>(1, most / 4 / (128 * 4)));
    JHN_TRY(prefer_max_carveout(pull_segments_kernel));
    JHN_LAUNCH("pull_segments_kernel", st, pull_segments_kernel<<<ctas, 128, 0, st>>>(S));
    return JHN_OK;
}

int ingest_frames_launch(const uint8_t *frames, int N, int H, int W, float *out, cudaStream_t st)
{
    const long long quads = (long long)N * H * (W / 4);
    JHN_LAUNCH("ingest_frames_kernel", st, ingest_frames_kernel<<<(unsigned)cdiv(quads, 256), 256, 0, st>>>(frames, H, W, quads, out));
    return JHN_OK;
}

int crop_normalize_u8_launch(const uint8_t *frames, int B, int ncam, int H, int W, int bbox, const int32_t *centerHM,
                             const int32_t *valid, const float *mean, const float *std, float *out, cudaStream_t st)
{
    const int threads_needed = bbox * (bbox / 4);
    JHN_LAUNCH("crop_normalize_u8_kernel", st,
               crop_normalize_u8_kernel<<<dim3(cdiv(threads_needed, 256), B * ncam), 256, 0, st>>>(
                   frames, H, W, bbox, centerHM, valid, ncam, mean[0], mean[1], mean[2], std[0], std[1], std[2], out));
    return JHN_OK;
}

int softplus2_launch(const float *v, long long n, float *out, cudaStream_t st)
{
    JHN_LAUNCH("softplus2_kernel", st, softplus2_kernel<<<(unsigned)cdiv(cdiv(n, 4), 256), 256, 0, st>>>(v, n, out));
    return JHN_OK;
}

int pad_border_launch(const float *in, long long N, int S, float *out, cudaStream_t st)
{
    const long long n_out = N * (S + 2) * (S + 2);
    JHN_LAUNCH("pad_border_kernel", st, pad_border_kernel<<<(unsigned)cdiv(n_out, 256), 256, 0, st>>>(in, S, n_out, out));
    return JHN_OK;
}

}  // namespace jhn
