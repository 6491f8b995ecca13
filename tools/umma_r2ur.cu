// Microbenchmark: cycles per tcgen05.mma (kind::f16, bf16 -> fp32, M=128, K=16, cta_group::1) issued back to
// back by one thread from shared-memory operands in the no-swizzle K-major layout conv_tc.cu uses.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_bench tools/umma_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout)
{
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}

// mode 0: A rows shift per MMA (like conv taps), same accumulator     mode 1: two accumulators alternate
// mode 2: M=64                                                         mode 3: 128B-swizzle descriptors (data is garbage, timing only)
__global__ void __launch_bounds__(128, 1) bench(int N, int iters, int mode, int lbo_a, long long *out)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ volatile uint32_t zero_s[4];
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 224 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0;
    if (threadIdx.x == 0) { zero_s[0] = 0; zero_s[1] = 0; mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        const int M = mode == 2 ? 64 : 128;
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + (mode >= 4 ? 100 * 1024 : 128 * 1024);
        const uint32_t layout = mode == 3 ? 2u : 0u;
        const uint32_t lboA = mode == 3 ? 16 : (uint32_t)lbo_a, sboA = mode == 3 ? 1024 : 128;
        const uint32_t lboB = mode == 3 ? 16 : (uint32_t)N * 16, sboB = mode == 3 ? 1024 : 128;
        // descriptors precomputed: 4 different A start rows (like taps), issue loop is 4 bare MMAs
        uint64_t ad[4], bd = desc(b0, lboB, sboB, layout);
        for (int i = 0; i < 4; ++i) ad[i] = desc(a0 + (mode == 3 ? 0u : (uint32_t)(i * 48)), lboA, sboA, layout);
        if (mode >= 4) {
            // conv3_tc.cu's exact operand stream: per tile 27 MMAs = 3 planes x 3 dy x 3 K-steps, A = plane box (PB = 204
            // positions, KC = 6 chunks) shifted by dy * 38 rows, B = a different 144-column weight slab per MMA (124 KB
            // resident).  mode 4: B as in the kernel; mode 5: the same B slab for every MMA; mode 6: A unshifted (dy * 40).
            const int PB = 204, KC = 6, N3 = 144, Wp = 38;
            const uint32_t w_units = b0 >> 4, ring_units = a0 >> 4, slot_units = KC * PB;
            const uint64_t hi_c = (uint64_t)(8u | (1u << 14)) << 32;
            const uint32_t lbo_a2 = (uint32_t)PB << 16, lbo_b2 = (uint32_t)N3 << 16;
            const uint32_t id = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N3 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            long long t0 = clock64();
            for (int it = 0; it < iters; it += 27) {
                const uint32_t d = tmem + (((it / 27) & 1) ? 256u : 0u);
                uint32_t acc = 0;
                uint32_t vt = 0;
                if (mode == 9) vt = zero_s[0];
#pragma unroll 1
                for (int dz = 0; dz < 3; ++dz) {
                    uint32_t vg = vt;
                    if (mode == 8) vg = zero_s[0];
                    const uint32_t a00 = ring_units + (uint32_t)(((it / 27) + dz) % 5) * slot_units + vg;
                    const uint32_t b00 = w_units + (false ? 0u : (uint32_t)(dz * 3 * KC * N3));
#pragma unroll
                    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                        for (int kc = 0; kc < KC; kc += 2) {
                            uint32_t vm = 0;
                            if (mode == 7) vm = zero_s[0];
                            const uint32_t a_lo = lbo_a2 | (a00 + vm + (uint32_t)(dy * Wp + kc * PB));
                            const uint32_t b_lo = lbo_b2 | (b00 + (false ? 0u : (uint32_t)((dy * KC + kc) * N3)));
                            mma(d, hi_c | (uint64_t)a_lo, hi_c | (uint64_t)b_lo, id, acc);
                            acc = 1;
                        }
                }
            }
            tc_commit(smem_u32(&bar));
            mbar_wait(smem_u32(&bar), 0);
            long long t1 = clock64();
            if (blockIdx.x == 0) *out = t1 - t0;
        } else {
        const uint32_t d0 = tmem, d1 = tmem + (mode == 1 ? 256u : 0u);
        mma(d0, ad[0], bd, idesc, 0);
        mma(d1, ad[1], bd, idesc, 0);
        long long t0 = clock64();
        for (int it = 0; it < iters; it += 4) {
            mma(d0, ad[0], bd, idesc, 1);
            mma(d1, ad[1], bd, idesc, 1);
            mma(d0, ad[2], bd, idesc, 1);
            mma(d1, ad[3], bd, idesc, 1);
        }
        tc_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0);
        long long t1 = clock64();
        if (blockIdx.x == 0) *out = t1 - t0;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main()
{
    long long *d_out, h;
    cudaMalloc(&d_out, 8);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    const int iters = 4096;
    const int Ns[] = {144};
    printf("grid mode lboA N cycles_per_mma ideal(=N/2)\n");
    for (int grid : {148})
        for (int mode : {0})
            for (int lbo : {3296, 2048})
                for (int N : Ns) {
                    if (mode != 0 && lbo != 3296) continue;
                    bench<<<grid, 128, 224 * 1024>>>(N, 64, mode, lbo, d_out);
                    bench<<<grid, 128, 224 * 1024>>>(N, iters, mode, lbo, d_out);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                    cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
                    printf("%d %d %d %d %.1f %.1f\n", grid, mode, lbo, N, (double)h / iters, N / 2.0);
                }
    for (int grid : {1, 148})
        for (int mode : {4, 7, 8, 9}) {
            const int it27 = 27 * 152;
            bench<<<grid, 128, 224 * 1024>>>(144, 54, mode, 0, d_out);
            bench<<<grid, 128, 224 * 1024>>>(144, it27, mode, 0, d_out);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
            printf("conv3-like grid %d mode %d: %.1f cycles per MMA (27 per tile)\n", grid, mode, (double)h / it27);
        }
    return 0;
}
