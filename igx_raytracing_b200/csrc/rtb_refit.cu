// rtb_refit.cu — device-side refit of the 8-wide compressed BVH after triangles moved.
//
// The reference mutates geometry every frame (test/scene/niels_scene.cpp:61-70) and uploads only the dirty ranges
// (igx/src/helpers/scene_graph.cpp:267-323); it has no acceleration structure to keep up to date.  Here the topology
// of the tree built by rtb_build_accel is kept and every box is recomputed from the triangle buffer as it now is:
//
//   k_refit_maxabs   largest |coordinate| -> the padding of the triangle boxes (DESIGN.md "conservative boxes")
//   k_refit_tris     traversal triangles (p0, e1, e2) rewritten from the uploaded records, same float subtractions
//   k_refit_level    one launch per tree level, deepest first (the tree is stored breadth-first, so a level is an index
//                    range): a node's child boxes are the boxes of its triangles or the already refitted boxes of its
//                    inner children; the node is re-encoded with the builder's own rounding (rtb_node8_encode.h)
//
// A refit of unmoved triangles reproduces the built tree bit for bit (tests/test_gpu_parity.py::test_refit_*).
#include <cuda_runtime.h>
#include "rtb_kernels.cuh"
#include "rtb_node8_encode.h"

namespace rtb {

__global__ void k_refit_maxabs(const TriangleRec* __restrict__ tris, uint32_t n, uint32_t* __restrict__ maxBits) {
    float m = 0.0f;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const TriangleRec t = tris[i];
        for (int a = 0; a < 3; ++a) m = fmaxf(m, fmaxf(fabsf(t.p0[a]), fmaxf(fabsf(t.p1[a]), fabsf(t.p2[a]))));
    }
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(maxBits, __float_as_uint(m));   // non-negative floats order like their bit patterns
}

__global__ void k_refit_tris(const TriangleRec* __restrict__ tris, TravTri* __restrict__ tt, uint32_t n) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t id = tt[k].id;
    const TriangleRec t = tris[id];
    TravTri o;
    for (int a = 0; a < 3; ++a) { o.p0[a] = t.p0[a]; o.e1[a] = t.p1[a] - t.p0[a]; o.e2[a] = t.p2[a] - t.p0[a]; }
    o.id = id; o.pad1 = 0; o.pad2 = 0;
    tt[k] = o;
}

RTB_ENC_HD float boxArea(const Box6& b) {
    const float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
    if (!(dx >= 0.0f && dy >= 0.0f && dz >= 0.0f)) return 0.0f;
    return 2.0f * (dx * dy + dy * dz + dz * dx);
}

// areaSums[0] += area of every node box, areaSums[1] += area x triangle count of every leaf slot: the two sums of the
// builder's SAH cost (rtb_bvh.cpp, buildCwbvh), so that rtb_accel_info.sah_cost follows the deformation
__global__ void k_refit_level(Node8* __restrict__ nodes, uint32_t first, uint32_t count, const TriangleRec* __restrict__ tris,
                              const TravTri* __restrict__ tt, Box6* __restrict__ nodeBox, const uint32_t* __restrict__ maxBits,
                              double* __restrict__ areaSums) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    double leafArea = 0.0, nodeArea = 0.0;
    if (i < count) {
    float maxAbs = __uint_as_float(*maxBits);
    if (!(maxAbs < INFINITY)) maxAbs = 1.0f;
    const float pad = node8Pad(maxAbs);
    const Node8 old = nodes[first + i];
    const uint32_t P = old.valid & 0x00FFFFFFu;
    Box6 cb[8];
    bool used[8];
    Box6 nb;
    for (int a = 0; a < 3; ++a) { nb.lo[a] = INFINITY; nb.hi[a] = -INFINITY; }
    for (int s = 0; s < 8; ++s) {
        used[s] = false;
        if ((old.imask >> s) & 1u) {
            cb[s] = nodeBox[old.childBase + __popc(old.imask & ((1u << s) - 1u))];
            used[s] = true;
        } else {
            const uint32_t cnt = __popc((P >> (3 * s)) & 7u);
            if (!cnt) continue;   // (an empty slot)
            const uint32_t t0 = old.triBase + __popc(P & ((1u << (3 * s)) - 1u));
            for (int a = 0; a < 3; ++a) { cb[s].lo[a] = INFINITY; cb[s].hi[a] = -INFINITY; }
            for (uint32_t k = 0; k < cnt; ++k) {
                const Box6 tb = node8TriangleBox(tris[tt[t0 + k].id], maxAbs, pad);
                for (int a = 0; a < 3; ++a) { cb[s].lo[a] = fminf(cb[s].lo[a], tb.lo[a]); cb[s].hi[a] = fmaxf(cb[s].hi[a], tb.hi[a]); }
            }
            leafArea += (double)boxArea(cb[s]) * (double)cnt;
            used[s] = true;
        }
        for (int a = 0; a < 3; ++a) { nb.lo[a] = fminf(nb.lo[a], cb[s].lo[a]); nb.hi[a] = fmaxf(nb.hi[a], cb[s].hi[a]); }
    }
    nodeBox[first + i] = nb;
    nodeArea = (double)boxArea(nb);
    Node8 out;
    memset(&out, 0, sizeof out);
    double step[3];
    node8Grid(nb, out, step);
    out.imask = old.imask; out.childBase = old.childBase; out.triBase = old.triBase; out.valid = old.valid; out.reserved = old.reserved;
    for (int s = 0; s < 8; ++s) node8Child(out, s, used[s] ? &cb[s] : nullptr, step);
    nodes[first + i] = out;
    }
    for (int o = 16; o > 0; o >>= 1) { nodeArea += __shfl_xor_sync(0xFFFFFFFFu, nodeArea, o); leafArea += __shfl_xor_sync(0xFFFFFFFFu, leafArea, o); }
    if ((threadIdx.x & 31) == 0) { atomicAdd(areaSums, nodeArea); atomicAdd(areaSums + 1, leafArea); }   // one pair of atomics per warp
}

void launch_refit(const TriangleRec* tris, uint32_t triCount, TravTri* tt, uint32_t ttCount, Node8* nodes, const uint32_t* levelFirst,
                  uint32_t levels, float* nodeBox, uint32_t* maxBits, double* areaSums, cudaStream_t st) {
    if (!ttCount || !levels) return;
    cudaMemsetAsync(maxBits, 0, sizeof(uint32_t), st);
    cudaMemsetAsync(areaSums, 0, 2 * sizeof(double), st);
    k_refit_maxabs<<<min((triCount + 255u) / 256u, 1184u), 256, 0, st>>>(tris, triCount, maxBits);
    k_refit_tris<<<(ttCount + 255u) / 256u, 256, 0, st>>>(tris, tt, ttCount);
    for (uint32_t l = levels; l-- > 0;) {
        const uint32_t first = levelFirst[l], count = levelFirst[l + 1] - first;
        if (count) k_refit_level<<<(count + 127u) / 128u, 128, 0, st>>>(nodes, first, count, tris, tt, reinterpret_cast<Box6*>(nodeBox), maxBits, areaSums);
    }
}

}  // namespace rtb
