// Probe kernels for tools/coreside_probe.py: what does a small CTA that merely SITS on an SM cost the persistent kernels?
//   mode 0: sleep (nanosleep loop) for `us` microseconds        — residency only
//   mode 1: spin on clock64 (ALU / issue slots)                  — issue pressure
//   mode 2: stream through a device buffer with 16-byte loads/stores (LSU / L1 / L2 traffic)
#include <cuda_runtime.h>
#include <stdint.h>
__global__ void __launch_bounds__(128, 16) dummy_kernel(int mode, long long cycles, uint4 *buf, size_t n)
{
    const long long t0 = clock64();
    if (mode == 0) { while (clock64() - t0 < cycles) __nanosleep(1000); }
    else if (mode == 1) { unsigned x = threadIdx.x; while (clock64() - t0 < cycles) x = x * 1664525u + 1013904223u; if (x == 0xdeadbeefu) buf[0].x = x; }
    else {
        size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
        while (clock64() - t0 < cycles) { uint4 v = __ldcs(buf + (i % n)); buf[(i + n / 2) % n] = v; i += (size_t)gridDim.x * blockDim.x; }
    }
}
// carveout: -1 = driver default, 0..100 = cudaFuncAttributePreferredSharedMemoryCarveout (100 = all shared memory)
extern "C" int dummy_set_carveout(int carveout)
{
    return (int)cudaFuncSetAttribute(dummy_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carveout);
}
extern "C" int dummy_launch(int ctas, int threads, int mode, double us, void *buf, size_t n_units, void *stream)
{
    dummy_kernel<<<ctas, threads, 0, (cudaStream_t)stream>>>(mode, (long long)(us * 1965.0), (uint4 *)buf, n_units);
    return (int)cudaGetLastError();
}
