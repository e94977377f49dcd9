// rtb_sort.cu — ray queues in light-space order (no reference counterpart: the reference traces one occlusion ray per
// thread in pixel order, SH/nv_all.shadow.comp:84-126).
//
// The occlusion rays of a frame go to one light (lightId = 0) and are nearly parallel (a sun) or converge on a point.  Rays
// that travel along the same line need the same nodes, whatever the depth they start from — but in pixel order a warp's rays
// start on triangles at unrelated depths and share nothing.  k_shadowgen therefore appends the LIVE rays to a queue
// (warp-aggregated append: no dead slots travel) and bins each by a 2D light-space coordinate (RayBin): a counting sort over
// Morton-ordered cells puts rays of neighbouring lines next to each other.
//
//   k_shadowgen            ray record + (cell, rank) with rank = atomicAdd(hist[cell], 1)      (rtb_kernels.cu)
//   k_scan_cells / _tops   exclusive scan of the cell histogram (1024-cell blocks, then the block totals)
//   k_scatter_rays         ray r -> position offset(cell) + rank
//
// The order inside a cell depends on the atomics' arrival order; results cannot (every ray sets its own pixel's bit).
#include <cuda_runtime.h>
#include "rtb_kernels.cuh"

namespace rtb {

__global__ void __launch_bounds__(1024) k_scan_cells(uint32_t* __restrict__ hist, uint32_t cells, uint32_t* __restrict__ blockSums) {
    __shared__ uint32_t sWarp[32];
    const uint32_t i = blockIdx.x * 1024u + threadIdx.x, lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t v = i < cells ? hist[i] : 0u;
    uint32_t inc = v;
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, o); if ((int)lane >= o) inc += t; }
    if (lane == 31u) sWarp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = sWarp[lane], winc = w;
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, winc, o); if ((int)lane >= o) winc += t; }
        sWarp[lane] = winc - w;
        if (lane == 31u) blockSums[blockIdx.x] = winc;
    }
    __syncthreads();
    if (i < cells) hist[i] = sWarp[warp] + inc - v;   // exclusive, block-local
}

__global__ void __launch_bounds__(1024) k_scan_tops(uint32_t* __restrict__ blockSums, uint32_t blocks) {
    __shared__ uint32_t sWarp[32];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t v = threadIdx.x < blocks ? blockSums[threadIdx.x] : 0u;
    uint32_t inc = v;
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, o); if ((int)lane >= o) inc += t; }
    if (lane == 31u) sWarp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = sWarp[lane], winc = w;
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, winc, o); if ((int)lane >= o) winc += t; }
        sWarp[lane] = winc - w;
    }
    __syncthreads();
    if (threadIdx.x < blocks) blockSums[threadIdx.x] = sWarp[warp] + inc - v;
}

__global__ void __launch_bounds__(256) k_scatter_rays(const RayQueue q, RayRec* __restrict__ outRays, uint32_t* __restrict__ outSlots) {
    const uint32_t n = *q.count;
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        const uint32_t cell = q.cell[r];
        const uint32_t pos = q.hist[cell] + q.blockSums[cell >> 10] + q.rank[r];
        const float4* src = reinterpret_cast<const float4*>(q.rays + r);
        float4* dst = reinterpret_cast<float4*>(outRays + pos);
        dst[0] = src[0]; dst[1] = src[1];
        outSlots[pos] = q.slotIds[r];
    }
}

void launch_sort_rays(const RayQueue& q, uint32_t cells, uint32_t maxRays, RayRec* outRays, uint32_t* outSlots, cudaStream_t st) {
    const uint32_t blocks = (cells + 1023u) / 1024u;   // <= 1024 (cells <= 2^20)
    k_scan_cells<<<blocks, 1024, 0, st>>>(q.hist, cells, q.blockSums);
    k_scan_tops<<<1, 1024, 0, st>>>(q.blockSums, blocks);
    const uint32_t grid = (maxRays + 255u) / 256u;
    k_scatter_rays<<<grid < 148u * 8u ? (grid ? grid : 1u) : 148u * 8u, 256, 0, st>>>(q, outRays, outSlots);
}

}  // namespace rtb
