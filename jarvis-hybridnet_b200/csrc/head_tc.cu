// Stages 2+3 fused tail of the bf16 path: V2VNet's 1x1x1 output layer (jarvis/hybridnet/v2vnet.py:94-101) as a
// tcgen05 GEMM whose epilogue is the centroid tail of HybridNetBackbone.forward (jarvis/hybridnet/model.py:73-87):
// softplus, the four sums of the sum-normalised centroid, the confidence maximum and the argmax voxel.  The raw
// [B,K,h^3] fp32 volume (137 MB per 32 frame sets) is never written: the kernel reads the bf16 activations once
// (HBM-bound) and leaves 32 bytes per (frame set, key point) in global accumulators, which a one-block kernel
// turns into points3D / confidences.
//
//     warp 0      TMA producer: one box of 128 consecutive padded positions x KC chunks per tile (cp.async.bulk)
//     warp 1      MMA issuer: KC/2 tcgen05.mma (M=128, N=NOUT, K=16) per tile, accumulators double-buffered in TMEM
//     warps 2-9   epilogue, two per TMEM lane quadrant, each owning half of the key points: tcgen05.ld -> +bias ->
//                 softplus -> per-thread running sums (n, sum hf*i, sum hf*j, sum hf*k), max hf, (max raw, argmax);
//                 warp-reduced when the CTA's tile range leaves a frame set: the four sums go to the warp's own slot
//                 (plain stores; the finalize kernel adds a frame set's slots in CTA order in fp64, so the key points
//                 are bit-identical from run to run), the maxima are order-independent atomicMax.
// The argmax is an atomicMax on a 64-bit key (order-preserving bits of the raw value, ~flat index), i.e. the first
// flat index of the maximum, independent of scheduling.
#include <atomic>

#include "tc_ptx.cuh"
#include "v2v.cuh"

namespace jhn {

constexpr int HD_THREADS = 320, HD_MAX_SLOTS = 8, HD_CH = 12;                 // HD_CH: key points per epilogue warp (K <= 24)

struct HeadLaunch {
    const uint4 *in;                                   // BP bf16 input [B][KC][D+2][(D+2)^2]
    const __nv_bfloat16 *w;                            // [KC][NOUT][8] bf16
    const float *bias;                                 // [NOUT]
    float *sums;                                       // [4 * (grid + B)][K][4]  per-CTA, per-quadrant partial n, sx, sy, sz (tc_ptx.cuh)
    unsigned int *hfmax;                               // [B][K]     bits of max softplus (>= 0, so ordered as uint)
    unsigned long long *key;                           // [B][K]     (ordered raw bits << 32) | ~flat index
    int B, D, K, KC, NOUT, NT, total_tiles, NS;
};

__device__ __forceinline__ void hd_ld16(uint32_t taddr, uint32_t *v)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                 "tcgen05.wait::ld.sync.aligned;\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr)
                 : "memory");
}

__device__ __forceinline__ unsigned int ordered_bits(float v)
{
    const unsigned int u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void __launch_bounds__(HD_THREADS, 1)
tc_head_centroid_kernel(const HeadLaunch L)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int D = L.D, Wp = D + 2, PP = Wp * Wp, KC = L.KC, NOUT = L.NOUT;
    const int w_bytes = KC * NOUT * 16, slot_bytes = KC * TILE_M * 16;

    uint8_t *w_smem = smem;
    uint8_t *ring = smem + (w_bytes + 127) / 128 * 128;
    float *bias_s = reinterpret_cast<float *>(ring + (size_t)L.NS * slot_bytes);
    uint64_t *bars = reinterpret_cast<uint64_t *>(bias_s + 32);
    uint64_t *full = bars, *empty = bars + HD_MAX_SLOTS, *tfull = bars + 2 * HD_MAX_SLOTS, *tempty = tfull + 2, *wbar = tempty + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(wbar + 1);

    const int t_begin = (int)((long long)L.total_tiles * blockIdx.x / gridDim.x);
    const int t_end = (int)((long long)L.total_tiles * (blockIdx.x + 1) / gridDim.x);

    if (threadIdx.x == 0) {
        for (int i = 0; i < L.NS; ++i) { mbar_init(smem_u32(full + i), 1); mbar_init(smem_u32(empty + i), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(tfull + i), 1); mbar_init(smem_u32(tempty + i), 8); }
        mbar_init(smem_u32(wbar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) bias_s[threadIdx.x] = threadIdx.x < NOUT ? L.bias[threadIdx.x] : 0.f;
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // tile t -> (b, z, pt): pt fastest, 128 consecutive padded positions of plane z starting at the first interior one
    if (warp == 0) {
        if (elect_one()) {
            mbar_expect_tx(smem_u32(wbar), (uint32_t)w_bytes);
            bulk_load(smem_u32(w_smem), L.w, (uint32_t)w_bytes, smem_u32(wbar));
            int slot = 0; uint32_t phase = 0;
            for (int t = t_begin; t < t_end; ++t) {
                const int pt = t % L.NT, bz = t / L.NT, z = bz % D, b = bz / D;
                const int start = Wp + 1 + pt * TILE_M;
                const int npos = min(TILE_M, PP - start);
                const uint32_t run = (uint32_t)npos * 16;
                mbar_wait(smem_u32(empty + slot), phase ^ 1);
                const uint32_t fb = smem_u32(full + slot);
                mbar_expect_tx(fb, run * (uint32_t)KC);
                const uint4 *src = L.in + ((size_t)b * KC * Wp + (z + 1)) * PP + start;
                uint8_t *dst = ring + (size_t)slot * slot_bytes;
                for (int j = 0; j < KC; ++j)
                    bulk_load(smem_u32(dst + (size_t)j * TILE_M * 16), src + (size_t)j * Wp * PP, run, fb);
                if (++slot == L.NS) { slot = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            const uint32_t idesc = umma_idesc_bf16(NOUT);
            const uint64_t hi_c = (uint64_t)(8u | (1u << 14)) << 32;
            const uint32_t lbo_a = (uint32_t)TILE_M << 16, lbo_b = (uint32_t)NOUT << 16;
            const uint32_t w_units = smem_u32(w_smem) >> 4, ring_units = smem_u32(ring) >> 4, slot_units = (uint32_t)slot_bytes >> 4;
            mbar_wait(smem_u32(wbar), 0);
            int slot = 0; uint32_t phase = 0;
            int ab = 0; uint32_t aphase = 0;
            for (int t = t_begin; t < t_end; ++t) {
                mbar_wait(smem_u32(tempty + ab), aphase ^ 1);
                mbar_wait(smem_u32(full + slot), phase);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(ab * 32);
                const uint32_t a0 = ring_units + (uint32_t)slot * slot_units;
                uint32_t acc = 0;
                for (int kc = 0; kc < KC; kc += 2) {
                    const uint32_t a_lo = lbo_a | (a0 + (uint32_t)(kc * TILE_M));
                    const uint32_t b_lo = lbo_b | (w_units + (uint32_t)(kc * NOUT));
                    tc_mma_bf16(d_tmem, hi_c | (uint64_t)a_lo, hi_c | (uint64_t)b_lo, idesc, acc);
                    acc = 1;
                }
                tc_commit(smem_u32(empty + slot));
                tc_commit(smem_u32(tfull + ab));
                if (++slot == L.NS) { slot = 0; phase ^= 1; }
                if (++ab == 2) { ab = 0; aphase ^= 1; }
            }
        }
    } else {
        const int q = warp & 3, half = (warp - 2) >> 2;
        const int row = q * 32 + lane;
        const int CH = (L.K + 1) / 2;                                         // key points of this warp: [k0, k0 + nk)
        const int k0 = half * CH, nk = min(CH, L.K - k0);
        float n[HD_CH], sx[HD_CH], sy[HD_CH], sz[HD_CH], hmax[HD_CH], rmax[HD_CH], bias_r[HD_CH];
        int arg[HD_CH];
#pragma unroll
        for (int i = 0; i < HD_CH; ++i) {
            n[i] = sx[i] = sy[i] = sz[i] = 0.f; hmax[i] = 0.f; rmax[i] = -INFINITY; arg[i] = 0x7fffffff;
            bias_r[i] = bias_s[min(k0 + i, 31)];
        }
        auto flush = [&](int b) {
#pragma unroll
            for (int i = 0; i < HD_CH; ++i) {
                if (i < nk) {                                                 // warp-uniform
                    float a = n[i], bx = sx[i], by = sy[i], bz = sz[i], hm = hmax[i];
                    unsigned long long ky = ((unsigned long long)ordered_bits(rmax[i]) << 32) | (unsigned int)(0xffffffffu - (unsigned int)arg[i]);
#pragma unroll
                    for (int s = 16; s > 0; s >>= 1) {
                        a += __shfl_xor_sync(0xffffffffu, a, s); bx += __shfl_xor_sync(0xffffffffu, bx, s);
                        by += __shfl_xor_sync(0xffffffffu, by, s); bz += __shfl_xor_sync(0xffffffffu, bz, s);
                        hm = fmaxf(hm, __shfl_xor_sync(0xffffffffu, hm, s));
                        const unsigned long long o = __shfl_xor_sync(0xffffffffu, ky, s);
                        ky = o > ky ? o : ky;
                    }
                    if (lane == 0) {
                        const size_t e = (size_t)b * L.K + k0 + i;
                        const size_t slot = (size_t)(((int)blockIdx.x + b) * 4 + q);
                        reinterpret_cast<float4 *>(L.sums)[slot * L.K + k0 + i] = make_float4(a, bx, by, bz);
                        atomicMax(L.hfmax + e, __float_as_uint(hm));
                        atomicMax(L.key + e, ky);
                    }
                }
                n[i] = sx[i] = sy[i] = sz[i] = 0.f; hmax[i] = 0.f; rmax[i] = -INFINITY; arg[i] = 0x7fffffff;
            }
        };
        int stat_b = -1;
        int ab = 0; uint32_t aphase = 0;
        for (int t = t_begin; t < t_end; ++t) {
            const int pt = t % L.NT, bz = t / L.NT, z = bz % D, b = bz / D;
            if (b != stat_b) {
                if (stat_b >= 0) flush(stat_b);
                stat_b = b;
            }
            const int p = Wp + 1 + pt * TILE_M + row;
            const int yp = p / Wp, xp = p - yp * Wp;
            const bool valid = p < PP && xp >= 1 && xp <= D && yp >= 1 && yp <= D;
            const float fi = (float)z, fj = (float)(yp - 1), fk = (float)(xp - 1);
            const int flat = (z * D + (yp - 1)) * D + (xp - 1);
            mbar_wait(smem_u32(tfull + ab), aphase);
            tc_fence_after();
            uint32_t r[16];
            hd_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * 32 + k0), r);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(tempty + ab));
            if (valid) {
#pragma unroll
                for (int i = 0; i < HD_CH; ++i) {
                    const float v = __uint_as_float(r[i]) + bias_r[i];
                    // nn.Softplus(beta=1, threshold=20) (model.py:73) as max(v,0) + log(1 + exp(-|v|)): two MUFU ops, absolute
                    // error < 1e-7 (the threshold branch is implied: for v > 20 the log term is below fp32 resolution of v)
                    const float hf = fmaxf(v, 0.f) + __logf(1.f + __expf(-fabsf(v)));
                    n[i] += hf;                                               // :76
                    sx[i] = fmaf(hf, fi, sx[i]);                              // :77-82
                    sy[i] = fmaf(hf, fj, sy[i]);
                    sz[i] = fmaf(hf, fk, sz[i]);
                    hmax[i] = fmaxf(hmax[i], hf);                             // :84
                    if (v > rmax[i]) { rmax[i] = v; arg[i] = flat; }
                }
            }
            if (++ab == 2) { ab = 0; aphase ^= 1; }
        }
        if (stat_b >= 0) flush(stat_b);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64u) : "memory");
    }
}

// accumulators -> points3D (mm), confidences, argmax   (model.py:84-87).  One warp per (frame set, key point): the
// lanes stride over the frame set's slots in CTA order, then a fixed shuffle tree: deterministic, fp64.
__global__ void centroid_finalize_kernel(const float *__restrict__ sums, const unsigned int *__restrict__ hfmax,
                                         const unsigned long long *__restrict__ key, int BK, int K, float spacing, float roi,
                                         const float *__restrict__ center3D, float *__restrict__ points, float *__restrict__ conf,
                                         int32_t *__restrict__ argmax, int grid, int Tb, long long T)
{
    const int o = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (o >= BK) return;
    const int b = o / K, k = o - b * K;
    const int c0 = stat_owner((long long)b * Tb, grid, T), c1 = stat_owner((long long)(b + 1) * Tb - 1, grid, T);
    const float4 *src = reinterpret_cast<const float4 *>(sums) + (size_t)(c0 + b) * 4 * K + k;
    double dn = 0.0, dx = 0.0, dy = 0.0, dz = 0.0;
    for (int s = lane; s < (c1 - c0 + 1) * 4; s += 32) {
        // a warp (quadrant q, key-point half) only flushes the key points it owns, and only rows it saw: every slot of
        // every CTA in [c0, c1] is written for every key point, because all 8 epilogue warps flush at a frame-set change
        const float4 v = src[(size_t)s * K];
        dn += (double)v.x; dx += (double)v.y; dy += (double)v.z; dz += (double)v.w;
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        dn += __shfl_xor_sync(0xffffffffu, dn, s); dx += __shfl_xor_sync(0xffffffffu, dx, s);
        dy += __shfl_xor_sync(0xffffffffu, dy, s); dz += __shfl_xor_sync(0xffffffffu, dz, s);
    }
    if (lane != 0) return;
    const float n = (float)dn, sx = (float)dx, sy = (float)dy, sz = (float)dz;
    conf[o] = fminf(__uint_as_float(hfmax[o]), 255.f) / 255.f;
    if (argmax) argmax[o] = (int32_t)(0xffffffffu - (unsigned int)(key[o] & 0xffffffffull));
    const float sc = spacing * 2.f, half = roi / 2.f;
    points[3 * o + 0] = (sx / n) * sc - half + center3D[3 * b + 0];
    points[3 * o + 1] = (sy / n) * sc - half + center3D[3 * b + 1];
    points[3 * o + 2] = (sz / n) * sc - half + center3D[3 * b + 2];
}

static size_t head_sum_bytes(int B, int K) { return align_up((size_t)stat_slots(256, B) * K * 16, 256); }
size_t head_acc_bytes(int B, int K) { return head_sum_bytes(B, K) + align_up((size_t)B * K * 4, 256) + align_up((size_t)B * K * 8, 256); }

static size_t head_smem(int KC, int NOUT, int NS)
{
    return ((size_t)KC * NOUT * 16 + 127) / 128 * 128 + (size_t)NS * KC * TILE_M * 16 + 32 * 4 + (2 * HD_MAX_SLOTS + 5) * 8 + 16;
}

bool head_supported(int K, int cin_pad, int cout_pad, int D) { return K <= 2 * HD_CH && cout_pad <= 32 && cin_pad % 16 == 0 && D >= 2; }

// in: BP bf16 activations of decoder output; acc: head_acc_bytes() of scratch.  Launches memset + GEMM/centroid + finalize.
int head_centroid_launch(const void *in, const __nv_bfloat16 *w, const float *bias, int B, int D, int K, int cin_pad, int cout_pad,
                         float spacing, float roi, const float *center3D, float *points, float *conf, int32_t *argmax, void *acc,
                         int sms, int max_smem, cudaStream_t st)
{
    HeadLaunch L;
    L.in = (const uint4 *)in; L.w = w; L.bias = bias; L.B = B; L.D = D; L.K = K; L.KC = cin_pad / 8; L.NOUT = cout_pad;
    char *p = (char *)acc;
    L.sums = (float *)p; p += head_sum_bytes(B, K);
    L.hfmax = (unsigned int *)p; p += align_up((size_t)B * K * 4, 256);
    L.key = (unsigned long long *)p;
    if (sms > 256) return fail(JHN_ERR_ARCH, "device has %d SMs; the partial-sum slots are sized for <= 256", sms);
    const int Wp = D + 2;
    L.NT = cdiv((long long)(D - 1) * Wp + D, TILE_M);
    L.total_tiles = B * D * L.NT;
    int ns = HD_MAX_SLOTS;
    while (ns > 2 && head_smem(L.KC, L.NOUT, ns) > (size_t)max_smem) --ns;
    if (head_smem(L.KC, L.NOUT, ns) > (size_t)max_smem) return fail(JHN_ERR_SHAPE, "fused head: tile does not fit shared memory");
    L.NS = ns;
    const size_t smem = head_smem(L.KC, L.NOUT, ns);
    JHN_CUDA(cudaMemsetAsync((char *)acc + head_sum_bytes(B, K), 0, head_acc_bytes(B, K) - head_sum_bytes(B, K), st));   // the two maxima
    static std::atomic<unsigned long long> configured{0ull};          // one bit per device
    int dev = 0;
    JHN_CUDA(cudaGetDevice(&dev));
    const unsigned long long bit = 1ull << (dev & 63);
    if (!(configured.load(std::memory_order_acquire) & bit)) {
        JHN_CUDA(cudaFuncSetAttribute(tc_head_centroid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        configured.fetch_or(bit, std::memory_order_release);
    }
    const int grid = L.total_tiles < sms ? L.total_tiles : sms;
    JHN_LAUNCH("tc_head_centroid_kernel", st, tc_head_centroid_kernel<<<grid, HD_THREADS, smem, st>>>(L));
    const int BK = B * K;
    JHN_LAUNCH("centroid_finalize_kernel", st,
               centroid_finalize_kernel<<<cdiv(BK, 4), 128, 0, st>>>(L.sums, L.hfmax, L.key, BK, K, spacing, roi, center3D, points,
                                                                     conf, argmax, grid, D * L.NT, (long long)L.total_tiles));
    return JHN_OK;
}

}  // namespace jhn
