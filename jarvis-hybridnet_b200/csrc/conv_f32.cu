// Stage 2, fp32 parity path: V2VNet (jarvis/hybridnet/v2vnet.py:12-102) with plain FFMA kernels on
// NCDHW fp32 tensors.  This path exists for the fp32 parity bar (key points within 0.05 mm of the
// reference's fp32 run); the throughput path is the bf16 tcgen05 implementation in conv_tc.cu.
//
//   conv3d_f32_kernel<KS>   direct convolution, 128 output voxels x 8 output channels per block,
//                           weights of 8 input channels staged in shared memory per step
//   convT_k2s2_f32_kernel   ConvTranspose3d(k=2,s=2): every output voxel has exactly one tap
//   instnorm_stats_kernel   two-pass mean / biased variance per (sample, channel)  (InstanceNorm3d)
//   norm_act_kernel         (x-mean)*rstd [+ residual] [ReLU] [+ skip]
#include "v2v.cuh"

namespace jhn {

void layer_table(int K, LayerDesc *d)
{
    const int C1 = K, C2 = 2 * K, C4 = 4 * K;
    d[L_FRONT0]  = {C1, C2, 3, 2, 1, 0};     // v2vnet.py:90  Basic3DBlock(C, 2C, 3, 2)
    d[L_FRONT1A] = {C2, C2, 3, 1, 1, 0};     // :91 Res3DBlock
    d[L_FRONT1B] = {C2, C2, 3, 1, 1, 0};
    d[L_POOL]    = {C2, C4, 2, 2, 0, 0};     // :67-68 encoder_pool1 Basic3DBlock(2C, 4C, 2, 2)
    d[L_MIDA]    = {C4, C4, 3, 1, 1, 0};     // :69 mid_res
    d[L_MIDB]    = {C4, C4, 3, 1, 1, 0};
    d[L_UP]      = {C4, C2, 2, 2, 0, 1};     // :70-71 decoder_upsample1 (ConvTranspose3d)
    d[L_DECA]    = {C2, C2, 3, 1, 1, 0};     // :72 decoder_res1
    d[L_DECB]    = {C2, C2, 3, 1, 1, 0};
    d[L_SKIPA]   = {C2, C2, 3, 1, 1, 0};     // :73 skip_res1
    d[L_SKIPB]   = {C2, C2, 3, 1, 1, 0};
    d[L_HEAD]    = {C2, C1, 1, 1, 0, 0};     // :94-95 output_layer
}

__global__ void pack_f32_kernel(const float *__restrict__ src, float *__restrict__ dst, int cout, int cin, int taps,
                                int transposed)
{
    const int n = ((cout + 7) / 8) * cin * taps * 8;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const int j = e & 7; int r = e >> 3;
    const int t = r % taps; r /= taps;
    const int ci = r % cin;
    const int co = (r / cin) * 8 + j;
    float v = 0.f;
    if (co < cout) v = transposed ? src[((size_t)ci * cout + co) * taps + t] : src[((size_t)co * cin + ci) * taps + t];
    dst[e] = v;
}

template <int KS>
__global__ void __launch_bounds__(128)
conv3d_f32_kernel(const float *__restrict__ in, const float *__restrict__ wp, const float *__restrict__ bias,
                  float *__restrict__ out, int Cin, int Cout, int Di, int Do, int stride, int pad)
{
    constexpr int T = KS * KS * KS, CI = 8;
    __shared__ __align__(16) float ws[CI * T * 8];
    const int nvo = Do * Do * Do;
    const long long nvi = (long long)Di * Di * Di;
    const int o = blockIdx.x * 128 + threadIdx.x;
    const int cb = blockIdx.y, b = blockIdx.z;
    const bool active = o < nvo;
    const int zo = o / (Do * Do), r = o - zo * Do * Do, yo = r / Do, xo = r - yo * Do;
    const int z0 = zo * stride - pad, y0 = yo * stride - pad, x0 = xo * stride - pad;
    uint32_t mask = 0;
#pragma unroll
    for (int t = 0; t < T; ++t) {
        const int dz = t / (KS * KS), dy = (t / KS) % KS, dx = t % KS;
        const bool ok = active && (unsigned)(z0 + dz) < (unsigned)Di && (unsigned)(y0 + dy) < (unsigned)Di &&
                        (unsigned)(x0 + dx) < (unsigned)Di;
        mask |= (ok ? 1u : 0u) << t;
    }
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = (cb * 8 + j < Cout) ? __ldg(bias + cb * 8 + j) : 0.f;
    const long long base = (long long)b * Cin * nvi + ((long long)z0 * Di + y0) * Di + x0;
    for (int ci0 = 0; ci0 < Cin; ci0 += CI) {
        const int nci = min(CI, Cin - ci0);
        const float4 *src = reinterpret_cast<const float4 *>(wp + ((size_t)cb * Cin + ci0) * T * 8);
        for (int e = threadIdx.x; e < nci * T * 2; e += 128) reinterpret_cast<float4 *>(ws)[e] = __ldg(src + e);
        __syncthreads();
        for (int c = 0; c < nci; ++c) {
            const long long cbase = base + (long long)(ci0 + c) * nvi;
#pragma unroll
            for (int t = 0; t < T; ++t) {
                if ((mask >> t) & 1u) {
                    const int dz = t / (KS * KS), dy = (t / KS) % KS, dx = t % KS;
                    const float x = __ldg(in + cbase + ((long long)dz * Di + dy) * Di + dx);
                    const float4 w0 = *reinterpret_cast<const float4 *>(&ws[(c * T + t) * 8]);
                    const float4 w1 = *reinterpret_cast<const float4 *>(&ws[(c * T + t) * 8 + 4]);
                    acc[0] = fmaf(x, w0.x, acc[0]); acc[1] = fmaf(x, w0.y, acc[1]);
                    acc[2] = fmaf(x, w0.z, acc[2]); acc[3] = fmaf(x, w0.w, acc[3]);
                    acc[4] = fmaf(x, w1.x, acc[4]); acc[5] = fmaf(x, w1.y, acc[5]);
                    acc[6] = fmaf(x, w1.z, acc[6]); acc[7] = fmaf(x, w1.w, acc[7]);
                }
            }
        }
        __syncthreads();
    }
    if (active) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int co = cb * 8 + j;
            if (co < Cout) out[((size_t)b * Cout + co) * nvo + o] = acc[j];
        }
    }
}

// ConvTranspose3d(k=2, s=2, p=0): out[2z+a,2y+b,2x+c] = bias + sum_ci x[ci,z,y,x] * W[ci,co,a,b,c]   (v2vnet.py:52)
__global__ void __launch_bounds__(128)
convT_k2s2_f32_kernel(const float *__restrict__ in, const float *__restrict__ wp, const float *__restrict__ bias,
                      float *__restrict__ out, int Cin, int Cout, int Di)
{
    constexpr int T = 8, CI = 8;
    __shared__ __align__(16) float ws[CI * T * 8];
    const int Do = 2 * Di, nvo = Do * Do * Do, nvi = Di * Di * Di;
    const int o = blockIdx.x * 128 + threadIdx.x;
    const int cb = blockIdx.y, b = blockIdx.z;
    const bool active = o < nvo;
    const int zo = o / (Do * Do), r = o - zo * Do * Do, yo = r / Do, xo = r - yo * Do;
    const int tap = ((zo & 1) * 2 + (yo & 1)) * 2 + (xo & 1);
    const int vi = ((zo >> 1) * Di + (yo >> 1)) * Di + (xo >> 1);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = (cb * 8 + j < Cout) ? __ldg(bias + cb * 8 + j) : 0.f;
    for (int ci0 = 0; ci0 < Cin; ci0 += CI) {
        const int nci = min(CI, Cin - ci0);
        const float4 *src = reinterpret_cast<const float4 *>(wp + ((size_t)cb * Cin + ci0) * T * 8);
        for (int e = threadIdx.x; e < nci * T * 2; e += 128) reinterpret_cast<float4 *>(ws)[e] = __ldg(src + e);
        __syncthreads();
        if (active) {
            for (int c = 0; c < nci; ++c) {
                const float x = __ldg(in + ((size_t)b * Cin + ci0 + c) * nvi + vi);
                const float4 w0 = *reinterpret_cast<const float4 *>(&ws[(c * T + tap) * 8]);
                const float4 w1 = *reinterpret_cast<const float4 *>(&ws[(c * T + tap) * 8 + 4]);
                acc[0] = fmaf(x, w0.x, acc[0]); acc[1] = fmaf(x, w0.y, acc[1]);
                acc[2] = fmaf(x, w0.z, acc[2]); acc[3] = fmaf(x, w0.w, acc[3]);
                acc[4] = fmaf(x, w1.x, acc[4]); acc[5] = fmaf(x, w1.y, acc[5]);
                acc[6] = fmaf(x, w1.z, acc[6]); acc[7] = fmaf(x, w1.w, acc[7]);
            }
        }
        __syncthreads();
    }
    if (active) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int co = cb * 8 + j;
            if (co < Cout) out[((size_t)b * Cout + co) * nvo + o] = acc[j];
        }
    }
}

__device__ __forceinline__ float block_sum(float v, float *sh)
{
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
    return t;
}

// InstanceNorm3d statistics: eps 1e-5, biased variance, no affine (v2vnet.py:18,33,37,54)
__global__ void __launch_bounds__(256)
instnorm_stats_kernel(const float *__restrict__ x, int nv, float eps, float2 *__restrict__ stats)
{
    __shared__ float sh[8];
    const float *p = x + (size_t)blockIdx.x * nv;
    float s = 0.f;
    for (int i = threadIdx.x; i < nv; i += blockDim.x) s += __ldg(p + i);
    const float mean = block_sum(s, sh) / (float)nv;
    float q = 0.f;
    for (int i = threadIdx.x; i < nv; i += blockDim.x) { const float d = __ldg(p + i) - mean; q = fmaf(d, d, q); }
    const float var = block_sum(q, sh) / (float)nv;
    if (threadIdx.x == 0) stats[blockIdx.x] = make_float2(mean, 1.f / sqrtf(var + eps));
}

__global__ void __launch_bounds__(256)
norm_act_kernel(float *__restrict__ x, const float2 *__restrict__ stats, const float *__restrict__ residual,
                const float *__restrict__ post_add, int relu, int nv)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nv) return;
    const size_t o = (size_t)blockIdx.y * nv + i;
    const float2 st = stats[blockIdx.y];
    float y = (x[o] - st.x) * st.y;
    if (residual) y += residual[o];
    if (relu) y = fmaxf(y, 0.f);
    if (post_add) y += post_add[o];
    x[o] = y;
}

int v2v_f32_pack(jhn_v2v *net, const float *const *tensors, cudaStream_t st)
{
    size_t total = 0;
    for (int l = 0; l < NUM_LAYERS; ++l) {
        const LayerDesc &d = net->desc[l];
        const int taps = d.ks * d.ks * d.ks;
        total += align_up((size_t)((d.cout + 7) / 8) * d.cin * taps * 8, 64) + align_up(d.cout, 64);
    }
    JHN_CUDA(cudaMalloc(&net->blob, total * sizeof(float)));
    float *p = net->blob;
    for (int l = 0; l < NUM_LAYERS; ++l) {
        const LayerDesc &d = net->desc[l];
        const int taps = d.ks * d.ks * d.ks;
        const int n = ((d.cout + 7) / 8) * d.cin * taps * 8;
        net->f32[l].w = p; p += align_up(n, 64);
        net->f32[l].bias = p; p += align_up(d.cout, 64);
        JHN_LAUNCH("pack_f32_kernel", st,
                   pack_f32_kernel<<<cdiv(n, 256), 256, 0, st>>>(tensors[2 * l], net->f32[l].w, d.cout, d.cin, taps, d.transposed));
        JHN_CUDA(cudaMemcpyAsync(net->f32[l].bias, tensors[2 * l + 1], d.cout * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    return JHN_OK;
}

size_t v2v_f32_workspace(const jhn_v2v *net, int B, int G)
{
    const int h = G / 2, q = G / 4, K = net->K;
    Arena a(nullptr, 0);
    for (int i = 0; i < 5; ++i) a.take<float>((size_t)B * 2 * K * h * h * h);
    for (int i = 0; i < 3; ++i) a.take<float>((size_t)B * 4 * K * q * q * q);
    a.take<float2>((size_t)B * 4 * K);
    return a.off;
}

namespace {
struct F32Ctx {
    const jhn_v2v *net; int B; cudaStream_t st; float2 *stats;

    int conv(int l, const float *in, float *out, int Di, int Do) const
    {
        const LayerDesc &d = net->desc[l];
        const LayerF32 &w = net->f32[l];
        const int nvo = Do * Do * Do;
        dim3 grid(cdiv(nvo, 128), (d.cout + 7) / 8, B);
        if (d.transposed)
            JHN_LAUNCH("convT_k2s2_f32_kernel", st,
                       convT_k2s2_f32_kernel<<<grid, 128, 0, st>>>(in, w.w, w.bias, out, d.cin, d.cout, Di));
        else if (d.ks == 3)
            JHN_LAUNCH("conv3d_f32_kernel<3>", st,
                       conv3d_f32_kernel<3><<<grid, 128, 0, st>>>(in, w.w, w.bias, out, d.cin, d.cout, Di, Do, d.stride, d.pad));
        else if (d.ks == 2)
            JHN_LAUNCH("conv3d_f32_kernel<2>", st,
                       conv3d_f32_kernel<2><<<grid, 128, 0, st>>>(in, w.w, w.bias, out, d.cin, d.cout, Di, Do, d.stride, d.pad));
        else
            JHN_LAUNCH("conv3d_f32_kernel<1>", st,
                       conv3d_f32_kernel<1><<<grid, 128, 0, st>>>(in, w.w, w.bias, out, d.cin, d.cout, Di, Do, d.stride, d.pad));
        return JHN_OK;
    }
    // InstanceNorm [+residual] [ReLU] [+post_add], in place on x [B][C][D^3]
    int norm(float *x, int C, int D, const float *residual, bool relu, const float *post_add) const
    {
        const int nv = D * D * D;
        JHN_LAUNCH("instnorm_stats_kernel", st, instnorm_stats_kernel<<<B * C, 256, 0, st>>>(x, nv, 1e-5f, stats));
        JHN_LAUNCH("norm_act_kernel", st,
                   norm_act_kernel<<<dim3(cdiv(nv, 256), B * C), 256, 0, st>>>(x, stats, residual, post_add, relu ? 1 : 0, nv));
        return JHN_OK;
    }
    // Res3DBlock (v2vnet.py:27-43): x -> relu(IN(conv(relu(IN(conv(x))))) + x) [+ post_add]
    int res(int la, int lb, const float *x, float *tmp, float *out, int C, int D, const float *post_add) const
    {
        JHN_TRY(conv(la, x, tmp, D, D));
        JHN_TRY(norm(tmp, C, D, nullptr, true, nullptr));
        JHN_TRY(conv(lb, tmp, out, D, D));
        return norm(out, C, D, x, true, post_add);
    }
};
}  // namespace

int v2v_f32_forward(const jhn_v2v *net, const float *x, int B, int G, float *out, void *ws, size_t ws_bytes,
                    cudaStream_t st)
{
    const int h = G / 2, q = G / 4, K = net->K, C2 = 2 * K, C4 = 4 * K;
    Arena a(ws, ws_bytes);
    float *bufA = a.take<float>((size_t)B * C2 * h * h * h), *bufB = a.take<float>((size_t)B * C2 * h * h * h);
    float *bufC = a.take<float>((size_t)B * C2 * h * h * h), *bufD = a.take<float>((size_t)B * C2 * h * h * h);
    float *bufE = a.take<float>((size_t)B * C2 * h * h * h);
    float *bufP = a.take<float>((size_t)B * C4 * q * q * q), *bufQ = a.take<float>((size_t)B * C4 * q * q * q);
    float *bufR = a.take<float>((size_t)B * C4 * q * q * q);
    float2 *stats = a.take<float2>((size_t)B * C4);
    if (!a.ok()) return fail(JHN_ERR_WORKSPACE, "v2v fp32 workspace: need %zu bytes, got %zu", a.off, ws_bytes);
    F32Ctx c{net, B, st, stats};

    JHN_TRY(c.conv(L_FRONT0, x, bufA, G, h));                                    // front_layers.0  (v2vnet.py:90)
    JHN_TRY(c.norm(bufA, C2, h, nullptr, true, nullptr));
    JHN_TRY(c.res(L_FRONT1A, L_FRONT1B, bufA, bufB, bufC, C2, h, nullptr));      // front_layers.1  x = C
    JHN_TRY(c.res(L_SKIPA, L_SKIPB, bufC, bufA, bufB, C2, h, nullptr));          // skip_res1       s = B   (:76)
    JHN_TRY(c.conv(L_POOL, bufC, bufP, h, q));                                   // encoder_pool1           (:77)
    JHN_TRY(c.norm(bufP, C4, q, nullptr, true, nullptr));
    JHN_TRY(c.res(L_MIDA, L_MIDB, bufP, bufQ, bufR, C4, q, nullptr));            // mid_res         R       (:78)
    JHN_TRY(c.conv(L_UP, bufR, bufA, q, h));                                     // decoder_upsample1 A     (:79)
    JHN_TRY(c.norm(bufA, C2, h, nullptr, true, nullptr));
    JHN_TRY(c.res(L_DECA, L_DECB, bufA, bufD, bufE, C2, h, bufB));               // decoder_res1 (+ s)      (:80-81)
    return c.conv(L_HEAD, bufE, out, h, h);                                      // output_layer            (v2vnet.py:101)
}

}  // namespace jhn
