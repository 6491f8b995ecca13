// Shared-memory wavefront model probe: how many data-pipe cycles does one LDS.128 take for a given lane -> 16-byte-unit
// pattern?  One warp per SM loops over dependent-free LDS.128s of a fixed pattern and reports cycles per instruction.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/lds_bench tools/lds_bench.cu && tools/lds_bench
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__global__ void probe(const int *units, int npat, long long *out, int iters, int warps)
{
    extern __shared__ uint4 sm[];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = make_uint4(i, i, i, i);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    for (int pat = 0; pat < npat; ++pat) {
        const int u = units[pat * 32 + lane];
        uint4 acc = make_uint4(0, 0, 0, 0);
        __syncthreads();
        const long long t0 = clock64();
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                uint4 v;                                         // +64 units = +1024 B keeps every bank group
                const uint32_t addr = (uint32_t)__cvta_generic_to_shared(sm + ((u + 64 * k) & 4095));
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
                acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
            }
        }
        const long long t1 = clock64();
        __syncthreads();
        if (acc.x == 0x12345678u) out[63] = acc.y + acc.z + acc.w;
        if (threadIdx.x == 0 && blockIdx.x == 0) out[pat] = (t1 - t0);
    }
}

int main()
{
    const int NP = 10;
    int h[NP * 32];
    const char *names[NP] = {
        "T0 lane l -> unit l (contiguous 512 B)",
        "T1 quarter-distinct groups, rows differ per quarter: u = (l%8) + 8*(l/8)*3",
        "T2 8-way inside each quarter: u = (l%8)*8 + (l/8)",
        "T3 same group for l and l+8 (different quarters, same half): u = (l%8) + 8*(l/8)",
        "T4 2-way inside a quarter (lanes 0,1 same group, different rows)",
        "T5 all lanes one address (broadcast)",
        "T6 2 distinct addresses per quarter in distinct groups",
        "T7 lanes l and l+16 same group different row (different halves)",
        "T8 48-byte vectors: lane l -> unit 3*l (stride 48 B)",
        "T9 48-byte vectors, pairs share a pixel: unit 3*(l/2)",
    };
    for (int l = 0; l < 32; ++l) {
        h[0 * 32 + l] = l;
        h[1 * 32 + l] = (l % 8) + 8 * (l / 8) * 3;
        h[2 * 32 + l] = (l % 8) * 8 + (l / 8);
        h[3 * 32 + l] = (l % 8) + 8 * (l / 8);
        h[4 * 32 + l] = (l % 8 == 1) ? (l / 8) * 64 + 8 : (l % 8) + (l / 8) * 64;
        h[5 * 32 + l] = 5;
        h[6 * 32 + l] = (l % 2) * 3 + (l / 8) * 8;
        h[7 * 32 + l] = (l % 16) + (l / 16) * 16 * 5;
        h[8 * 32 + l] = 3 * l;
        h[9 * 32 + l] = 3 * (l / 2);
    }
    int *d; long long *o;
    cudaMalloc(&d, sizeof(h)); cudaMalloc(&o, 64 * 8);
    cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    for (int warps = 1; warps <= 8; warps *= 2) {
        const int iters = 2000;
        probe<<<1, 32 * warps, 65536>>>(d, NP, o, iters, warps);
        cudaDeviceSynchronize();
        long long r[64];
        cudaMemcpy(r, o, sizeof(r), cudaMemcpyDeviceToHost);
        printf("---- %d warp(s) per SM: cycles per LDS.128 warp instruction (x warps) ----\n", warps);
        for (int p = 0; p < NP; ++p)
            printf("  %-80s %6.2f\n", names[p], (double)r[p] / (iters * 16.0) / 1.0);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
