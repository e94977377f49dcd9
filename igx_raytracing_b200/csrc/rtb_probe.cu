// rtb_probe.cu — read-only streaming microbenchmark for the L2 roofline denominator ("L2 = to be measured
// by the builder with a read-only streaming microbenchmark over a <= 64 MB buffer").  Not on the hot path.
//
// A persistent grid reads a buffer that fits the 126 MB L2 over and over with 16-byte ld.global.cg loads (cached in L2
// only, so L1 hits cannot inflate the figure); the first pass warms L2 and is not timed.
#include <cuda_runtime.h>
#include <stdint.h>

namespace rtb {

__global__ void __launch_bounds__(256) k_stream_read(const uint4* __restrict__ buf, size_t n16, uint32_t passes, uint32_t* __restrict__ sink) {
    uint4 acc = make_uint4(0u, 0u, 0u, 0u);
    const size_t stride = (size_t)gridDim.x * blockDim.x, first = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (uint32_t p = 0; p < passes; ++p) {
        #pragma unroll 8
        for (size_t i = first; i < n16; i += stride) {
            uint4 v;
            asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(buf + i));
            acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
        }
    }
    if ((acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x9E3779B9u) *sink = acc.x;   // never true for the zero-filled buffer; keeps the loads alive
}

// returns GB/s (1e9 bytes per second) or a negative cudaError_t
double measure_l2_read_gbs(size_t bytes, uint32_t passes, cudaStream_t st) {
    const size_t n16 = bytes / 16;
    uint4* buf = nullptr; uint32_t* sink = nullptr;
    cudaError_t e = cudaMalloc(&buf, n16 * 16);
    if (e != cudaSuccess) return -(double)e;
    e = cudaMalloc(&sink, 4);
    if (e != cudaSuccess) { cudaFree(buf); return -(double)e; }
    cudaMemsetAsync(buf, 0, n16 * 16, st);
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int blocks = sms * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_stream_read<<<blocks, 256, 0, st>>>(buf, n16, 2, sink);   // warm L2
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0, st);
        k_stream_read<<<blocks, 256, 0, st>>>(buf, n16, passes, sink);
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    e = cudaGetLastError();
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(buf); cudaFree(sink);
    if (e != cudaSuccess) return -(double)e;
    return (double)(n16 * 16) * passes / (best * 1e-3) / 1e9;
}

}  // namespace rtb
