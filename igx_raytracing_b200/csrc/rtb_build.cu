// rtb_build.cu — the 8-wide compressed BVH built ON THE DEVICE (no reference counterpart: the reference has no acceleration
// structure; this replaces the host build of rtb_bvh.cpp where add / del / compaction of triangles — ref:
// igx/src/helpers/scene_graph.cpp:343-376,378-522 — would otherwise stall the frame for 0.7 s per million triangles).
//
//   k_build_bounds      scene bounds of the triangle centroids (block reduction + ordered-integer atomics)
//   k_build_morton      63-bit keys: 30-bit Morton code of the centroid << 32 | triangle index (unique, so the hierarchy needs no
//                       tie rule); sorted with cub::DeviceRadixSort (library call, like a cuBLAS GEMM would be)
//   k_lbvh_hierarchy    Karras 2012: every internal node of the binary radix tree in parallel (its key range, split, children)
//   k_lbvh_fit          boxes bottom-up: the second thread to arrive at a node merges its children
//   k_collapse_level    one launch per level of the 8-wide tree: a node starts from a binary node's two children and keeps opening
//                       the child with the largest surface area until it has eight (or nothing is left to open); children whose
//                       subtree holds at most leafMax (1 .. 3) triangles become leaf slots, the others nodes of the next level (allocated
//                       consecutively, so the tree is breadth-first and a level is an index range); child slots are assigned by
//                       octant like the host builder's
//   launch_refit        (rtb_refit.cu) fills every box, grid and plane with the builder's own encoder and sums the SAH cost —
//                       the topology is all this file has to produce
//
// The result obeys the same contract as the host builder's (every stored box contains its triangles' padded boxes), so the hits
// are the brute-force loop's (tests/test_gpu_parity.py::test_device_builder_*).
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>
#include <vector>
#include "rtb_kernels.cuh"
#include "rtb_node8_encode.h"

namespace rtb {

namespace {

__device__ __forceinline__ uint32_t orderedU(float f) { const uint32_t b = __float_as_uint(f); return b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u); }
__device__ __forceinline__ float fromOrderedU(uint32_t u) { return __uint_as_float(u ^ ((u >> 31) ? 0x80000000u : 0xFFFFFFFFu)); }

struct TriBox { float lo[3], hi[3]; };
__device__ __forceinline__ TriBox triBox(const TriangleRec& t) {
    TriBox b;
    for (int a = 0; a < 3; ++a) {
        b.lo[a] = fminf(fminf(t.p0[a], t.p1[a]), t.p2[a]);
        b.hi[a] = fmaxf(fmaxf(t.p0[a], t.p1[a]), t.p2[a]);
        if (!(b.lo[a] <= b.hi[a])) { b.lo[a] = 0.0f; b.hi[a] = 0.0f; }   // a NaN vertex: sorted somewhere, kept reachable by the refit's padded box
    }
    return b;
}

// bounds[0..2] = min, bounds[3..5] = max of the centroids, as ordered integers
__global__ void __launch_bounds__(256) k_build_bounds(const TriangleRec* __restrict__ tris, uint32_t n, uint32_t* __restrict__ bounds) {
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const TriBox b = triBox(tris[i]);
        for (int a = 0; a < 3; ++a) { const float c = 0.5f * (b.lo[a] + b.hi[a]); if (isfinite(c)) { lo[a] = fminf(lo[a], c); hi[a] = fmaxf(hi[a], c); } }
    }
    for (int a = 0; a < 3; ++a) {
        for (int o = 16; o > 0; o >>= 1) { lo[a] = fminf(lo[a], __shfl_xor_sync(0xFFFFFFFFu, lo[a], o)); hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xFFFFFFFFu, hi[a], o)); }
        if ((threadIdx.x & 31) == 0) { atomicMin(bounds + a, orderedU(lo[a])); atomicMax(bounds + 3 + a, orderedU(hi[a])); }
    }
}

__device__ __forceinline__ uint32_t expand10(uint32_t v) {   // 10 bits -> every third bit
    v &= 0x3FFu;
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

__global__ void __launch_bounds__(256) k_build_morton(const TriangleRec* __restrict__ tris, uint32_t n, const uint32_t* __restrict__ bounds, unsigned long long* __restrict__ keys) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const TriBox b = triBox(tris[i]);
    uint32_t q[3];
    for (int a = 0; a < 3; ++a) {
        const float lo = fromOrderedU(bounds[a]), hi = fromOrderedU(bounds[3 + a]);
        const float ext = hi - lo;
        float c = ext > 0.0f ? (0.5f * (b.lo[a] + b.hi[a]) - lo) / ext : 0.0f;
        c = fminf(fmaxf(c, 0.0f), 1.0f);
        q[a] = min((uint32_t)(c * 1024.0f), 1023u);
    }
    const uint32_t m = (expand10(q[0]) << 2) | (expand10(q[1]) << 1) | expand10(q[2]);
    keys[i] = ((unsigned long long)m << 32) | i;
}

// binary radix tree over n sorted, unique keys: internal nodes 0 .. n-2.  Child link: >= 0 internal node, < 0 leaf ~index.
struct LbvhNode { int left, right, parent; uint32_t first, last; };

__device__ __forceinline__ int delta(const unsigned long long* keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    return __clzll(keys[i] ^ keys[j]);
}

__global__ void __launch_bounds__(256) k_lbvh_hierarchy(const unsigned long long* __restrict__ keys, int n, LbvhNode* __restrict__ nodes, int* __restrict__ leafParent) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1) >= 0 ? 1 : -1;
    const int dMin = delta(keys, n, i, i - d);
    int lMax = 2;
    while (delta(keys, n, i, i + lMax * d) > dMin) lMax <<= 1;
    int l = 0;
    for (int t = lMax >> 1; t >= 1; t >>= 1) if (delta(keys, n, i, i + (l + t) * d) > dMin) l += t;
    const int j = i + l * d;
    const int dNode = delta(keys, n, i, j);
    int s = 0;
    for (int t = (l + 1) >> 1;; t = (t + 1) >> 1) {
        if (delta(keys, n, i, i + (s + t) * d) > dNode) s += t;
        if (t == 1) break;
    }
    const int gamma = i + s * d + min(d, 0);
    const int lo = min(i, j), hi = max(i, j);
    const int left = lo == gamma ? ~gamma : gamma, right = hi == gamma + 1 ? ~(gamma + 1) : gamma + 1;
    nodes[i].left = left; nodes[i].right = right; nodes[i].first = (uint32_t)lo; nodes[i].last = (uint32_t)hi;
    // a node's parent field is written by its parent's thread (a different field from the ones above: no conflict)
    if (left >= 0) nodes[left].parent = i; else leafParent[~left] = i;
    if (right >= 0) nodes[right].parent = i; else leafParent[~right] = i;
    if (i == 0) nodes[0].parent = -1;
}

__global__ void __launch_bounds__(256) k_lbvh_fit(const unsigned long long* __restrict__ keys, int n, const TriangleRec* __restrict__ tris, const LbvhNode* __restrict__ nodes,
                                                  const int* __restrict__ leafParent, TriBox* boxes, uint32_t* __restrict__ arrived) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int node = leafParent[i];
    while (node >= 0) {
        if (atomicAdd(arrived + node, 1u) == 0u) return;   // the first to arrive leaves; the second finds both children done
        __threadfence();
        const LbvhNode nd = nodes[node];
        // boxes written by other SMs: read them from L2 (a stale L1 line may hold a neighbour of an older read)
        auto boxL2 = [&](int r) { TriBox x; const float* q = reinterpret_cast<const float*>(boxes + r); for (int k = 0; k < 3; ++k) { x.lo[k] = __ldcg(q + k); x.hi[k] = __ldcg(q + 3 + k); } return x; };
        const TriBox a = nd.left >= 0 ? boxL2(nd.left) : triBox(tris[(uint32_t)keys[~nd.left]]);
        const TriBox b = nd.right >= 0 ? boxL2(nd.right) : triBox(tris[(uint32_t)keys[~nd.right]]);
        TriBox m;
        for (int k = 0; k < 3; ++k) { m.lo[k] = fminf(a.lo[k], b.lo[k]); m.hi[k] = fmaxf(a.hi[k], b.hi[k]); }
        boxes[node] = m;
        __threadfence();
        node = nd.parent;
    }
}

__device__ __forceinline__ float areaOf(const TriBox& b) {
    const float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
    return dx * dy + dy * dz + dz * dx;
}

// counters: [0] nodes of the next level, [1] triangles placed so far, [2] leaf slots so far
__global__ void __launch_bounds__(128) k_collapse_level(const unsigned long long* __restrict__ keys, const TriangleRec* __restrict__ tris, const LbvhNode* __restrict__ bn,
                                                        const TriBox* __restrict__ boxes, const int* __restrict__ rootOf, uint32_t first, uint32_t count,
                                                        uint32_t nextFirst, uint32_t capacity, Node8* __restrict__ nodes8, int* __restrict__ rootOfNext, TravTri* __restrict__ tt,
                                                        uint32_t* __restrict__ counters, const uint32_t leafMax) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= count) return;
    const int root = rootOf[w];
    int ref[8];
    int n = 2;
    ref[0] = bn[root].left; ref[1] = bn[root].right;
    auto boxOf = [&](int r) { return r >= 0 ? boxes[r] : triBox(tris[(uint32_t)keys[~r]]); };
    auto trisOf = [&](int r) { return r >= 0 ? bn[r].last - bn[r].first + 1u : 1u; };
    while (n < 8) {   // open the internal child with the largest surface area
        int best = -1; float bestArea = -1.0f;
        for (int k = 0; k < n; ++k) if (ref[k] >= 0) { const float ar = areaOf(boxes[ref[k]]); if (ar > bestArea) { bestArea = ar; best = k; } }
        if (best < 0) break;
        const int r = ref[best];
        ref[best] = bn[r].left; ref[n++] = bn[r].right;
    }
    // slot assignment: child i -> slot s maximising dot(centre_i - centre_node, d_s), d_s = (+-1, +-1, +-1) by the bits of s
    const TriBox nb = boxes[root];
    float cx[8][3];
    for (int k = 0; k < n; ++k) { const TriBox b = boxOf(ref[k]); for (int a = 0; a < 3; ++a) cx[k][a] = 0.5f * (b.lo[a] + b.hi[a]) - 0.5f * (nb.lo[a] + nb.hi[a]); }
    int childAt[8] = {-1, -1, -1, -1, -1, -1, -1, -1};
    uint32_t doneMask = 0, usedMask = 0;
    for (int it = 0; it < n; ++it) {
        int bi = -1, bs = -1; float bc = -INFINITY;
        for (int k = 0; k < n; ++k) if (!((doneMask >> k) & 1u))
            for (int s = 0; s < 8; ++s) if (!((usedMask >> s) & 1u)) {
                const float d = ((s & 1) ? cx[k][0] : -cx[k][0]) + ((s & 2) ? cx[k][1] : -cx[k][1]) + ((s & 4) ? cx[k][2] : -cx[k][2]);
                if (d > bc) { bc = d; bi = k; bs = s; }
            }
        if (bi < 0) { for (int k = 0; k < n && bi < 0; ++k) if (!((doneMask >> k) & 1u)) bi = k; for (int s = 0; s < 8 && bs < 0; ++s) if (!((usedMask >> s) & 1u)) bs = s; }
        doneMask |= 1u << bi; usedMask |= 1u << bs; childAt[bs] = bi;
    }
    uint32_t inner = 0, triTotal = 0, leafSlots = 0, imask = 0, presence = 0;
    for (int s = 0; s < 8; ++s) {
        const int k = childAt[s];
        if (k < 0) continue;
        const uint32_t cnt = trisOf(ref[k]);
        if (cnt > leafMax) { imask |= 1u << s; ++inner; }
        else { presence |= ((1u << cnt) - 1u) << (3 * s); triTotal += cnt; ++leafSlots; }
    }
    const uint32_t childBase = inner ? nextFirst + atomicAdd(counters, inner) : 0u;
    const uint32_t triBase = triTotal ? atomicAdd(counters + 1, triTotal) : 0u;
    if (leafSlots) atomicAdd(counters + 2, leafSlots);
    Node8 out;
    memset(&out, 0, sizeof out);
    out.imask = (uint8_t)imask; out.childBase = childBase; out.triBase = triBase; out.valid = (imask << 24) | presence;
    uint32_t ci = 0, ti = 0;
    for (int s = 0; s < 8; ++s) {
        const int k = childAt[s];
        if (k < 0) continue;
        const int r = ref[k];
        if ((imask >> s) & 1u) {
            if (childBase + ci < capacity) rootOfNext[childBase + ci - nextFirst] = r;
            ++ci;
        } else {
            const uint32_t f = r >= 0 ? bn[r].first : (uint32_t)~r, cnt = trisOf(r);
            for (uint32_t q = 0; q < cnt; ++q) tt[triBase + ti + q].id = (uint32_t)keys[f + q];
            ti += cnt;
        }
    }
    nodes8[first + w] = out;
}

// mean edge length of the nodes that hold only triangles (the packet rule of rtb_api.cu wants it): sums[0] += extent, sums[1] += 1
__global__ void __launch_bounds__(256) k_leaf_extent(const Node8* __restrict__ nodes8, const float* __restrict__ nodeBox, uint32_t count, double* __restrict__ sums) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    double e = 0.0, c = 0.0;
    if (i < count && nodes8[i].imask == 0) {
        const float* b = nodeBox + 6 * (size_t)i;
        e = ((double)(b[3] - b[0]) + (double)(b[4] - b[1]) + (double)(b[5] - b[2])) / 3.0; c = 1.0;
    }
    for (int o = 16; o > 0; o >>= 1) { e += __shfl_xor_sync(0xFFFFFFFFu, e, o); c += __shfl_xor_sync(0xFFFFFFFFu, c, o); }
    if ((threadIdx.x & 31) == 0 && c > 0.0) { atomicAdd(sums, e); atomicAdd(sums + 1, c); }
}

template <class T> cudaError_t devAlloc(T*& p, size_t n) { return cudaMalloc(reinterpret_cast<void**>(&p), (n ? n : 1) * sizeof(T)); }

}  // namespace

// Builds the 8-wide tree over tris[0, n) into nodes8 (capacity nodeCapacity) and tt (capacity n).  Returns cudaSuccess and fills
// the out parameters, or an error; *tooDeep is set when the tree has more levels than maxLevels (the caller falls back).
cudaError_t device_build_cwbvh(const TriangleRec* tris, uint32_t n, Node8* nodes8, uint32_t nodeCapacity, TravTri* tt, float* nodeBox, uint32_t* maxBits,
                               double* areaSums, uint32_t maxLevels, uint32_t leafMax, std::vector<uint32_t>& levelFirst, uint32_t& nodeCount, uint32_t& leafSlots,
                               float& leafNodeExtent, bool* tooDeep, cudaStream_t st) {
    *tooDeep = false;
    levelFirst.clear();
    unsigned long long *keys = nullptr, *keysSorted = nullptr;
    LbvhNode* bn = nullptr; TriBox* boxes = nullptr; int *leafParent = nullptr, *rootA = nullptr, *rootB = nullptr;
    uint32_t *arrived = nullptr, *bounds = nullptr, *counters = nullptr; void* tmp = nullptr; double* extSums = nullptr;
    cudaError_t e = cudaSuccess;
    auto cleanup = [&]() {
        cudaFree(keys); cudaFree(keysSorted); cudaFree(bn); cudaFree(boxes); cudaFree(leafParent); cudaFree(rootA); cudaFree(rootB);
        cudaFree(arrived); cudaFree(bounds); cudaFree(counters); cudaFree(tmp); cudaFree(extSums);
    };
#define RTB_B(call) do { e = (call); if (e != cudaSuccess) { cleanup(); return e; } } while (0)
    RTB_B(devAlloc(keys, n)); RTB_B(devAlloc(keysSorted, n)); RTB_B(devAlloc(bn, n)); RTB_B(devAlloc(boxes, n)); RTB_B(devAlloc(leafParent, n));
    RTB_B(devAlloc(rootA, nodeCapacity)); RTB_B(devAlloc(rootB, nodeCapacity)); RTB_B(devAlloc(arrived, n)); RTB_B(devAlloc(bounds, 6)); RTB_B(devAlloc(counters, 4));
    RTB_B(devAlloc(extSums, 2));
    const uint32_t initBounds[6] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u, 0u};
    RTB_B(cudaMemcpyAsync(bounds, initBounds, sizeof initBounds, cudaMemcpyHostToDevice, st));
    k_build_bounds<<<min((n + 255u) / 256u, 1184u), 256, 0, st>>>(tris, n, bounds);
    k_build_morton<<<(n + 255u) / 256u, 256, 0, st>>>(tris, n, bounds, keys);
    size_t tmpBytes = 0;
    RTB_B(cub::DeviceRadixSort::SortKeys(nullptr, tmpBytes, keys, keysSorted, (int)n, 0, 62, st));
    RTB_B(cudaMalloc(&tmp, tmpBytes ? tmpBytes : 1));
    RTB_B(cub::DeviceRadixSort::SortKeys(tmp, tmpBytes, keys, keysSorted, (int)n, 0, 62, st));
    RTB_B(cudaMemsetAsync(arrived, 0, (size_t)n * 4, st));
    k_lbvh_hierarchy<<<(n + 255u) / 256u, 256, 0, st>>>(keysSorted, (int)n, bn, leafParent);
    k_lbvh_fit<<<(n + 255u) / 256u, 256, 0, st>>>(keysSorted, (int)n, tris, bn, leafParent, boxes, arrived);
    // ---- collapse, level by level ---------------------------------------------------------------------------------------------
    const int zero = 0;
    RTB_B(cudaMemcpyAsync(rootA, &zero, 4, cudaMemcpyHostToDevice, st));   // level 0: the binary root
    RTB_B(cudaMemsetAsync(counters, 0, 16, st));
    uint32_t first = 0, count = 1;
    int* cur = rootA; int* nxt = rootB;
    levelFirst.push_back(0);
    while (count) {
        if (levelFirst.size() > maxLevels) { *tooDeep = true; cleanup(); return cudaSuccess; }
        const uint32_t nextFirst = first + count;
        if (nextFirst > nodeCapacity) { cleanup(); return cudaErrorMemoryAllocation; }
        RTB_B(cudaMemsetAsync(counters, 0, 4, st));
        k_collapse_level<<<(count + 127u) / 128u, 128, 0, st>>>(keysSorted, tris, bn, boxes, cur, first, count, nextFirst, nodeCapacity, nodes8, nxt, tt, counters, leafMax);
        uint32_t next = 0;
        RTB_B(cudaMemcpyAsync(&next, counters, 4, cudaMemcpyDeviceToHost, st));
        RTB_B(cudaStreamSynchronize(st));
        if (nextFirst + next > nodeCapacity) { cleanup(); return cudaErrorMemoryAllocation; }
        levelFirst.push_back(nextFirst);
        first = nextFirst; count = next;
        std::swap(cur, nxt);
    }
    nodeCount = first;
    uint32_t hc[4];
    RTB_B(cudaMemcpyAsync(hc, counters, 16, cudaMemcpyDeviceToHost, st));
    RTB_B(cudaStreamSynchronize(st));
    leafSlots = hc[2];
    // ---- boxes, grids, planes, SAH sums: the refit, with the builder's own encoder ----------------------------------------------------
    launch_refit(tris, n, tt, n, nodes8, levelFirst.data(), (uint32_t)levelFirst.size() - 1, nodeBox, maxBits, areaSums, st);
    RTB_B(cudaMemsetAsync(extSums, 0, 16, st));
    k_leaf_extent<<<(nodeCount + 255u) / 256u, 256, 0, st>>>(nodes8, nodeBox, nodeCount, extSums);
    double hs[2];
    RTB_B(cudaMemcpyAsync(hs, extSums, 16, cudaMemcpyDeviceToHost, st));
    RTB_B(cudaStreamSynchronize(st));
    leafNodeExtent = hs[1] > 0.0 ? (float)(hs[0] / hs[1]) : 0.0f;
    RTB_B(cudaGetLastError());
#undef RTB_B
    cleanup();
    return cudaSuccess;
}

}  // namespace rtb
