// rtb_api.cu — the C ABI of include/rtb200.h: context, device buffers, uploads, pass dispatch, read-back.
// This layer is what replaces ignis (GPUBuffer / Texture / Descriptors / Pipeline / CommandList) under the
// reference's render tasks.  No CPU fallback exists: without a usable CUDA device every call fails.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <initializer_list>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/rtb200.h"
#include "rtb_bvh.h"
#include "rtb_kernels.cuh"

using namespace rtb;

namespace {
std::string g_createError;

template <class T>
struct DevBuf {
    T* p = nullptr; size_t count = 0;
    cudaError_t alloc(size_t n) {
        if (n <= count && p) return cudaSuccess;
        release();
        const cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&p), (n ? n : 1) * sizeof(T));
        if (e == cudaSuccess) count = n ? n : 1; else p = nullptr;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; count = 0; }
    size_t bytes() const { return count * sizeof(T); }
};
}  // namespace

struct rtb_ctx {
    int device = 0;
    rtb_limits limits{};
    cudaStream_t ownStream = nullptr, stream = nullptr;
    std::string error;

    // host mirrors of the small uniform buffers
    CameraRec camera{}; bool cameraSet = false;
    SceneInfoRec info{};
    uint32_t shadowSamplesProp = 1;

    // scene buffers (device)
    DevBuf<TriangleRec> triangles; DevBuf<float4> spheres; DevBuf<float> cubes; DevBuf<float4> planes;
    DevBuf<LightRec> lights; DevBuf<MaterialRec> materials; DevBuf<uint32_t> materialIndices;
    DevBuf<uint2> skybox; uint32_t skyW = 0, skyH = 0;
    DevBuf<SeedRec> seed;
    std::vector<uint8_t> triangleMirror;   // host copy of the triangle buffer: the BVH is built on the host

    // acceleration structure
    DevBuf<BvhNode> nodes; DevBuf<Node8> nodes8; DevBuf<TravTri> travTris;
    uint32_t nodeCount = 0; rtb_accel_mode accelMode = RTB_ACCEL_BRUTE; bool accelValid = false;
    BvhStats stats;
    DevBuf<float> nodeBox; DevBuf<uint32_t> maxBits; DevBuf<double> areaSums;   // refit scratch: 6 floats per node, largest |coordinate|, SAH sums
    uint32_t builtTriangles = 0, refits = 0;

    // frame resources
    uint32_t width = 0, height = 0, samples = 0;
    FrameMap fm{};
    uint32_t tileRank = 0, tileCount = 1;
    DevBuf<float4> dirT, uvN, accum; DevBuf<uint2> lighting; DevBuf<uint32_t> bits, rgba8, rgba8Tiled;
    // per-lane wavefront buffers.  Lane 0 serves the whole frame (per-pass dispatch, rays-in, path frames) or, when a frame runs as
    // two half-frame lanes on two streams (RTB_OPT_FRAME_LANES), the even local blocks; lane 1 the odd ones.
    struct LaneBufs {
        DevBuf<RayRec> rays; DevBuf<TriHit> hits; DevBuf<uint32_t> workCounter;
        DevBuf<RayRec> queueRays; DevBuf<uint32_t> queueSlots, sortedSlots, queueCell, queueRank, queueHist, queueSums, queueCount, queueFallbackCount;
        void release() { rays.release(); hits.release(); workCounter.release(); queueRays.release(); queueSlots.release(); sortedSlots.release(); queueCell.release();
                         queueRank.release(); queueHist.release(); queueSums.release(); queueCount.release(); queueFallbackCount.release(); }
    } lane[2];
    FrameMap laneFm[2]{};             // the two half-frame maps: lane h of rank r is virtual rank h * n + r of 2 n
    cudaStream_t laneStream = nullptr; cudaEvent_t evFork = nullptr, evJoin = nullptr;
    uint32_t lanesOpt = 1;            // RTB_OPT_FRAME_LANES
    // spheres / cubes through their own trees (rtb_trace8s.cuh), built on the device when a type has at least primTreeMin primitives
    struct PrimTreeBufs {
        DevBuf<TriangleRec> proxies; DevBuf<Node8> nodes; DevBuf<TravTri> tt; DevBuf<float> nodeBox; DevBuf<uint32_t> maxBits; DevBuf<double> areaSums;
        std::vector<uint32_t> levelFirst; uint32_t nodeCount = 0, count = 0; bool valid = false, dirty = true;
        void release() { proxies.release(); nodes.release(); tt.release(); nodeBox.release(); maxBits.release(); areaSums.release(); }
    } primTree[2];                    // 0 spheres, 1 cubes
    uint32_t primTreeMin = 64;        // RTB_OPT_PRIMITIVE_TREES: 0 = never
    DevBuf<PrimHit> primHitS[2], primHitC[2], rinPrimS, rinPrimC;   // winners per wavefront slot (per lane) / per rays-in ray
    // beyond the reference: every light evaluated, tile light lists, History blend (LightsView in rtb_kernels.cuh)
    uint32_t lightsOpt = 0;           // RTB_OPT_LIGHTS: 0 reference (light 0 x lightCount), 1 all lights, 2 all lights through tile lists
    float historyAlpha = 0.0f;        // RTB_OPT_HISTORY_ALPHA (float bits): 0 = off
    bool historyValid = false;        // false: the next lighting pass starts the History texture (alpha 1)
    DevBuf<uint32_t> lightTileCount, lightTileList; DevBuf<uint2> history;
    // what lighting.comp derives per (pixel, sample) from the frame size alone (LightsView.lightCache); refilled when the key changes
    DevBuf<float4> lightCache;
    struct LightCacheKey { uint32_t w = 0, h = 0, samples = 0, rank = 0, count = 0, kind = 0; uint64_t lightEpoch = 0;
                           bool operator==(const LightCacheKey& o) const { return w == o.w && h == o.h && samples == o.samples && rank == o.rank && count == o.count && kind == o.kind && lightEpoch == o.lightEpoch; } } lightCacheKey;
    uint64_t light0Epoch = 1;         // bumped when lights[0] is rewritten
    DevBuf<SunFrame> sunFrame; uint64_t sunEpoch = 0; bool sunValid = false;   // SceneView.sun0
    uint32_t lightCacheOpt = 1;       // RTB_OPT_LIGHT_CACHE
    uint32_t bitsLayers = 0;          // layers the shadow-word buffer currently holds
    // RTB_PASS_FRAME as two CUDA graphs (everything before the shade launch / the shade launch), replayed while nothing the
    // recorded launches hold by value has changed: `stamp` counts those changes
    uint32_t graphOpt = 1;            // RTB_OPT_FRAME_GRAPH
    uint64_t stamp = 1, lastFrameStamp = 0, graphStamp = 0;
    cudaGraphExec_t graphA = nullptr, graphB = nullptr;
    bool capturing = false;
    // Consecutive frames overlap (RTB_OPT_FRAME_OVERLAP; recorded frames only): a frame is FRONT = K0 + camera rays + nearest hit +
    // G-buffer on `frontStream` and BACK = shadow rays + occlusion + lighting + composite on `stream`.  What FRONT hands to BACK —
    // G-buffer, wavefront buffers, the Seed as K0 left it — exists twice (`alt` is the set not in use; the sets are swapped at every
    // overlapped frame, so c->dirT etc. always are the latest frame's), so FRONT of frame k+1 runs while BACK of frame k drains:
    // the tails of the persistent launches are filled by the other stream's work.
    uint32_t overlapOpt = 1;
    cudaStream_t frontStream = nullptr, midStream = nullptr;   // FRONT; shadow rays + occlusion when the frame runs in three stages
    cudaEvent_t evMidDone[2] = {nullptr, nullptr};
    cudaEvent_t evFrontDone[2] = {nullptr, nullptr}, evBackDone[2] = {nullptr, nullptr}, evJoinF = nullptr, evJoinB = nullptr;
    bool backDoneSet[2] = {false, false};
    bool frontDirty = false;          // the front stream holds work the back stream has not been ordered after
    bool frontNeedsBack = true;       // the back stream holds work (uploads, direct passes) the front stream must be ordered after
    int setIndex = 0;                 // which physical set c->dirT etc. hold
    bool altAllocated = false;
    struct FrameSet { DevBuf<float4> dirT, uvN; LaneBufs lane; DevBuf<PrimHit> primHitS, primHitC; DevBuf<uint32_t> bits; } alt;
    DevBuf<SeedRec> seedSnap;         // [set]
    SeedRec* seedUse = nullptr;       // non-null while an overlapped frame is being recorded: the set's snapshot
    cudaGraphExec_t ovFront[2] = {nullptr, nullptr}, ovShadow[2] = {nullptr, nullptr}, ovShade[2] = {nullptr, nullptr};
    uint64_t ovStamp[2] = {0, 0};
    DevBuf<TraceCounters> counters;   // [0] primary, [1] shadow
    bool countersOn = false;
    uint32_t countersMode = 0;        // 1: per-ray algorithmic counts (per-lane kernel); 2: what the kernels in use fetch
    uint32_t packetsOpt = 2;          // RTB_OPT_PRIMARY_PACKETS: 0 off, 1 union packets, 2 auto, 3 frustum packets
    int lastPrimaryPackets = 0;       // PACKETS_* of the last camera-ray launch
    uint32_t fuseOpt = 1;             // RTB_OPT_FUSE_PRIMARY
    uint32_t builderOpt = 0;          // RTB_OPT_ACCEL_BUILDER: 0 host (binned SAH + optimal collapse), 1 device (LBVH + greedy collapse)
    uint32_t builtBy = 0;             // builder of the tree in use
    uint32_t shadowOrder = 1;         // RTB_OPT_SHADOW_ORDER: 0 slot order, 1 queue of live rays (default), 2 queue sorted in light space
    LightRec light0{};                // host mirror of lights[0] (the one light the shadow pass samples): picks the sort key
    // wavefront path tracing (rtb_path_frame)
    DevBuf<float4> pathT, pathL, pathDirect; DevBuf<RayRec> pathRays[2], pathShadowRays; DevBuf<uint32_t> pathSlots[2], pathShadowSlots, pathCounts;
    DevBuf<uint8_t> pathOccA, pathOccB;
    std::vector<cudaEvent_t> pathEv;   // 4 per depth: nearest-hit launch begin / end, occlusion launch begin / end; then frame begin / end
    uint32_t pathDepths = 0, pathBounces = 0, pathKernelLaunches = 0; bool pathTimed = false;
    uint32_t releaseBuild = 0;        // RTB_OPT_SHADER_BUILD: 0 = DEBUG build of the reference shaders (what ships), 1 = RELEASE

    // rays-in scratch
    DevBuf<RayRec> rinRays; DevBuf<TriHit> rinHits; DevBuf<uint32_t> rinObj; DevBuf<float> rinT; DevBuf<float2> rinUv; DevBuf<uint8_t> rinOcc, rinOcc2;

    cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // frame-phase boundaries
    // asynchronous read-back (rtb_readback_async): a copy stream ordered after the frame by evReady; the next pass that
    // overwrites the target waits for evCopied, everything before it overlaps the copy
    cudaStream_t copyStream = nullptr; cudaEvent_t evReady = nullptr, evCopied = nullptr;
    int copyTarget = -1;
    bool frameTimed = false;
};

namespace {

int fail(rtb_ctx* c, rtb_status st, const std::string& msg) { if (c) c->error = msg; else g_createError = msg; return st; }
int cudaFail(rtb_ctx* c, cudaError_t e, const char* what) { return fail(c, RTB_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e)); }

#define RTB_CUDA(c, call) do { const cudaError_t e_ = (call); if (e_ != cudaSuccess) return cudaFail((c), e_, #call); } while (0)
#define RTB_BIND(c) do { const cudaError_t e_ = cudaSetDevice((c)->device); if (e_ != cudaSuccess) return cudaFail((c), e_, "cudaSetDevice"); } while (0)

SceneView sceneView(const rtb_ctx* c) {
    SceneView v{};
    v.triangles = c->triangles.p; v.spheres = c->spheres.p; v.cubes = c->cubes.p; v.planes = c->planes.p;
    v.lights = c->lights.p; v.materials = c->materials.p; v.materialIndices = c->materialIndices.p;
    v.skybox = c->skyW ? c->skybox.p : nullptr; v.skyW = c->skyW; v.skyH = c->skyH;
    v.info = c->info;
    v.nodes = c->nodes.p; v.nodes8 = c->nodes8.p; v.travTris = c->travTris.p; v.nodeCount = c->nodeCount;
    v.releaseBuild = c->releaseBuild;
    v.sun0 = c->sunValid ? c->sunFrame.p : nullptr;
    v.sphereTree = c->primTree[0].valid ? 1u : 0u; v.cubeTree = c->primTree[1].valid ? 1u : 0u;
    v.useBvh = !c->accelValid ? ACCEL_KIND_BRUTE : (c->accelMode == RTB_ACCEL_BVH ? ACCEL_KIND_CWBVH : (c->accelMode == RTB_ACCEL_BVH2 ? ACCEL_KIND_BVH2 : ACCEL_KIND_BRUTE));
    return v;
}

void fillFrameMap(FrameMap& fm, uint32_t w, uint32_t h, uint32_t rank, uint32_t nranks, uint32_t tiledMul, uint32_t tiledAdd) {
    fm.w = w; fm.h = h;
    fm.blocksX = (w + 31) / 32; fm.blocksY = (h + 31) / 32;
    fm.rank = rank; fm.nranks = nranks;
    const uint32_t total = fm.blocksX * fm.blocksY;
    fm.localBlocks = total > fm.rank ? (total - fm.rank + fm.nranks - 1) / fm.nranks : 0;
    fm.localSlots = fm.localBlocks * 1024u;
    fm.tiledMul = tiledMul; fm.tiledAdd = tiledAdd;
}
void makeFrameMap(rtb_ctx* c) {
    fillFrameMap(c->fm, c->width, c->height, c->tileRank, c->tileCount, 1, 0);
    // lane h takes the rank's local blocks 2 k + h: global blocks g = (2 k + h) n + r = k (2 n) + (h n + r)
    for (uint32_t h = 0; h < 2; ++h) fillFrameMap(c->laneFm[h], c->width, c->height, h * c->tileCount + c->tileRank, 2 * c->tileCount, 2, h);
}

uint32_t shadowWords(uint32_t w, uint32_t h, uint32_t samples) { return ((w + 15) / 16) * ((h + 1) / 2) * samples; }

void releaseAlt(rtb_ctx* c);
int allocFrame(rtb_ctx* c) {
    makeFrameMap(c);
    if (c->frontStream) { RTB_CUDA(c, cudaStreamSynchronize(c->frontStream)); RTB_CUDA(c, cudaStreamSynchronize(c->midStream)); c->frontDirty = false; }
    releaseAlt(c);   // the second set of FRONT -> BACK buffers (frame overlap) follows the frame size: reallocated when next needed
    const size_t px = (size_t)c->width * c->height;
    RTB_CUDA(c, c->dirT.alloc(px)); RTB_CUDA(c, c->uvN.alloc(px)); RTB_CUDA(c, c->accum.alloc(px));
    RTB_CUDA(c, c->lighting.alloc(px)); RTB_CUDA(c, c->rgba8.alloc(px));
    RTB_CUDA(c, c->bits.alloc(shadowWords(c->width, c->height, c->samples)));
    RTB_CUDA(c, c->rgba8Tiled.alloc((size_t)((c->fm.blocksX * c->fm.blocksY + c->fm.nranks - 1) / c->fm.nranks) * 1024u));
    RTB_CUDA(c, c->lane[0].rays.alloc((size_t)c->fm.localSlots * (c->samples ? c->samples : 1)));
    RTB_CUDA(c, c->lane[0].hits.alloc(c->fm.localSlots));
    RTB_CUDA(c, c->lane[1].rays.alloc((size_t)c->laneFm[1].localSlots * (c->samples ? c->samples : 1)));
    RTB_CUDA(c, c->lane[1].hits.alloc(c->laneFm[1].localSlots));
    RTB_CUDA(c, c->lane[1].workCounter.alloc(1));
    // pixels owned by other ranks are never written: keep them defined
    RTB_CUDA(c, cudaMemsetAsync(c->dirT.p, 0, c->dirT.bytes(), c->stream));
    RTB_CUDA(c, cudaMemsetAsync(c->uvN.p, 0, c->uvN.bytes(), c->stream));
    RTB_CUDA(c, cudaMemsetAsync(c->accum.p, 0, c->accum.bytes(), c->stream));
    RTB_CUDA(c, cudaMemsetAsync(c->lighting.p, 0, c->lighting.bytes(), c->stream));
    RTB_CUDA(c, cudaMemsetAsync(c->rgba8.p, 0, c->rgba8.bytes(), c->stream));
    RTB_CUDA(c, cudaMemsetAsync(c->bits.p, 0, c->bits.bytes(), c->stream));
    return RTB_OK;
}

// before a pass overwrites `target`: let a read-back of it that is still in flight finish first (device-side wait)
int waitCopy(rtb_ctx* c, std::initializer_list<int> targets) {
    if (c->copyTarget < 0 || c->capturing) return RTB_OK;   // (while a frame is being recorded the dispatcher does the waits, outside the graph)
    for (int t : targets)
        if (t == c->copyTarget) { RTB_CUDA(c, cudaStreamWaitEvent(c->stream, c->evCopied, 0)); c->copyTarget = -1; break; }
    return RTB_OK;
}
int drainCopy(rtb_ctx* c) {
    if (c->copyTarget >= 0) { RTB_CUDA(c, cudaEventSynchronize(c->evCopied)); c->copyTarget = -1; }
    return RTB_OK;
}

// the Seed the launches of a frame read: the buffer K0 updates in place, or — overlapped frames — the set's snapshot of it
const SeedRec* seedFor(const rtb_ctx* c) { return c->seedUse ? c->seedUse : c->seed.p; }

// frame overlap: order the context's stream after everything the front stream holds (before anything that reads or writes what
// FRONT writes: G-buffer, Seed, wavefront buffers); joinFront alone keeps the next FRONT free to run ahead
int joinFront(rtb_ctx* c) {
    if (!c->frontStream || !c->frontDirty) return RTB_OK;
    RTB_CUDA(c, cudaEventRecord(c->evJoinF, c->frontStream));
    RTB_CUDA(c, cudaStreamWaitEvent(c->stream, c->evJoinF, 0));
    c->frontDirty = false;
    return RTB_OK;
}
// ... and the context's stream is about to get work the next FRONT depends on (scene uploads, builds, direct passes, allocations)
int quiesce(rtb_ctx* c) { c->frontNeedsBack = true; return joinFront(c); }

// SceneView.sun0: the light-only part of a directional lights[0], re-evaluated on the context's stream after lights[0] was rewritten
// (the launches of a frame in flight precede it on that stream; the front stream's launches do not read lights)
int ensureSunFrame(rtb_ctx* c) {
    const bool want = c->lightCacheOpt && (c->light0.colorBType >> 16) != LIGHT_POINT;
    if (!want) { if (c->sunValid) { c->sunValid = false; ++c->stamp; } return RTB_OK; }
    if (c->sunValid && c->sunEpoch == c->light0Epoch) return RTB_OK;
    RTB_CUDA(c, c->sunFrame.alloc(1));
    launch_sun_frame(c->lights.p, c->sunFrame.p, c->stream);
    if (!c->sunValid) { c->sunValid = true; ++c->stamp; }   // recorded launches hold the pointer (or its absence) by value
    c->sunEpoch = c->light0Epoch;
    return RTB_OK;
}

// scene counts against the capacities of rtb_create, and the acceleration structure against the triangle buffer:
// shared by the dispatch path and the rays-in entry points (both index the scene buffers by these counts)
int checkScene(rtb_ctx* c) {
    if (c->accelMode != RTB_ACCEL_BRUTE && !c->accelValid && c->info.triangleCount)
        return fail(c, RTB_ERR_STATE, "triangles changed since the last rtb_build_accel");
    if (c->info.triangleCount > c->limits.max_triangles || c->info.sphereCount > c->limits.max_spheres || c->info.cubeCount > c->limits.max_cubes ||
        c->info.planeCount > c->limits.max_planes || c->info.lightCount > c->limits.max_lights || c->info.materialCount > c->limits.max_materials)
        return fail(c, RTB_ERR_CAPACITY, "scene info counts exceed the capacities given to rtb_create");
    return RTB_OK;
}

int checkReady(rtb_ctx* c) {
    if (!c->width || !c->height) return fail(c, RTB_ERR_STATE, "rtb_dispatch before rtb_resize");
    if (!c->cameraSet) return fail(c, RTB_ERR_STATE, "rtb_dispatch before the camera was uploaded");
    if (c->camera.width != c->width || c->camera.height != c->height)
        return fail(c, RTB_ERR_STATE, "camera.width/height differ from the size given to rtb_resize");
    return checkScene(c);
}

// Camera rays of an 8x4-pixel patch are walked through the 8-wide tree as one packet (rtb_trace8f.cuh) when the patch is
// small against the tree's leaf nodes: its width at the distance of the scene centre, 8 pixels wide, must stay below
// PACKET_RATIO mean leaf-node edges.  Beyond that the triangles of the union, which every lane tests, grow with the square
// of the ratio.  Measured on B200 with the four-node walk (nearest-hit phase incl. ray generation and G-buffer finish):
// ratio 0.76 (1M-triangle soup, 4K) 1.38 ms against 5.25 per ray; ratio 1.64 (same soup from z = 30) 0.58 against 0.87;
// ratio 16.8 (10M-triangle height field, 1080p) 14.5 against 0.51.
constexpr float PACKET_RATIO = 2.0f;
int primaryPackets(const rtb_ctx* c) {
    if (c->accelMode != RTB_ACCEL_BVH || !c->accelValid || !c->info.triangleCount) return PACKETS_OFF;
    if (c->countersOn && c->countersMode == 1) return PACKETS_OFF;
    if (c->packetsOpt != 2) return (int)c->packetsOpt;
    const CameraRec& cam = c->camera;
    if (cam.projectionType != 0 || !cam.width) return PACKETS_OFF;   // the frustum kernel wants one origin per packet
    float ctr[3], e2c = 0.0f, rad = 0.0f, pw = 0.0f, pd = 0.0f;
    for (int a = 0; a < 3; ++a) {
        ctr[a] = 0.5f * (c->stats.lo[a] + c->stats.hi[a]);
        e2c += (ctr[a] - cam.eye[a]) * (ctr[a] - cam.eye[a]);
        rad += 0.25f * (c->stats.hi[a] - c->stats.lo[a]) * (c->stats.hi[a] - c->stats.lo[a]);
        pw += (cam.p1[a] - cam.p0[a]) * (cam.p1[a] - cam.p0[a]);
        const float mid = 0.5f * (cam.p1[a] + cam.p2[a]) - cam.eye[a];   // centre of the screen plane
        pd += mid * mid;
    }
    const float dist = std::max(std::sqrt(e2c), 0.5f * std::sqrt(rad));
    const float patch = 8.0f * std::sqrt(pw) / (float)cam.width / std::max(std::sqrt(pd), 1e-20f) * dist;
    return patch < PACKET_RATIO * c->stats.leafNodeExtent ? PACKETS_FRUSTUM : PACKETS_OFF;
}

// Spheres / cubes get a tree of their own when there are many (RTB_OPT_PRIMITIVE_TREES) and the triangles are not searched by the
// reference's loop (RTB_ACCEL_BRUTE keeps every loop verbatim).  Built on the device over proxy triangles (rtb_trace8s.cuh) by the
// builder and refit kernels the triangle tree uses; rebuilt when the type's buffer or count changed.  Synchronises when it builds.
int ensurePrimTrees(rtb_ctx* c) {
    for (int kind = 0; kind < 2; ++kind) {
        rtb_ctx::PrimTreeBufs& T = c->primTree[kind];
        const uint32_t n = kind == 0 ? c->info.sphereCount : c->info.cubeCount;
        const bool want = c->primTreeMin && n >= std::max(c->primTreeMin, 2u) && c->accelMode != RTB_ACCEL_BRUTE && (c->accelValid || !c->info.triangleCount);
        if (!want) { if (T.valid) { T.valid = false; ++c->stamp; } continue; }
        if (T.valid && !T.dirty && T.count == n) continue;
        const uint32_t cap = n + 64;
        { const int rc = quiesce(c); if (rc) return rc; }   // (frames in flight on the front stream read the tree)
        RTB_CUDA(c, T.proxies.alloc(n)); RTB_CUDA(c, T.nodes.alloc(cap)); RTB_CUDA(c, T.tt.alloc(n)); RTB_CUDA(c, T.nodeBox.alloc((size_t)cap * 6));
        RTB_CUDA(c, T.maxBits.alloc(1)); RTB_CUDA(c, T.areaSums.alloc(2));
        launch_proxy_triangles(kind, c->spheres.p, c->cubes.p, n, T.proxies.p, c->stream);
        bool tooDeep = false;
        uint32_t nodeCount = 0, leafSlots = 0; float leafExtent = 0.0f;
        const cudaError_t e = device_build_cwbvh(T.proxies.p, n, T.nodes.p, cap, T.tt.p, T.nodeBox.p, T.maxBits.p, T.areaSums.p, 20, 1, T.levelFirst, nodeCount,
                                                 leafSlots, leafExtent, &tooDeep, c->stream);
        if (e != cudaSuccess) return cudaFail(c, e, "device_build_cwbvh (spheres / cubes)");
        T.valid = !tooDeep; T.dirty = false; T.count = n; T.nodeCount = nodeCount;   // too deep (thousands of coincident primitives): the loops stay
        ++c->stamp;
    }
    return RTB_OK;
}
PrimTree primTreeOf(const rtb_ctx* c, int kind) { return PrimTree{c->primTree[kind].nodes.p, c->primTree[kind].tt.p, c->primTree[kind].nodeCount}; }

// `mark` (single-lane frame dispatch only) records an event after each phase so the traversal launches can be timed alone.
// lane < 0: the whole frame on the context's stream; lane 0 / 1: that half-frame lane (lane 1 on its own stream).
struct LaneRef { const FrameMap& fm; rtb_ctx::LaneBufs& b; cudaStream_t st; };
LaneRef laneOf(rtb_ctx* c, int lane) {
    if (lane < 0) return LaneRef{c->fm, c->lane[0], c->stream};
    return LaneRef{c->laneFm[lane], c->lane[lane], lane == 1 ? c->laneStream : c->stream};
}

int passRaygen(rtb_ctx* c, bool mark, int lane = -1) {
    if (lane < 0) { const int rc = waitCopy(c, {RTB_TGT_DIR_T, RTB_TGT_UV_NORMAL}); if (rc) return rc; }
    const LaneRef L = laneOf(c, lane);
    const SceneView sv = sceneView(c);
    if (lane <= 0) c->lastPrimaryPackets = primaryPackets(c);
    const bool trees = sv.sphereTree || sv.cubeTree;
    if (c->lastPrimaryPackets == PACKETS_FRUSTUM && !c->countersOn && c->fuseOpt && !trees) {
        // one launch: rays generated in registers, traced, G-buffer written (the phase events collapse onto the trace phase)
        if (mark) RTB_CUDA(c, cudaEventRecord(c->ev[2], L.st));
        launch_primary_fused(L.fm, sv, &c->camera, seedFor(c), c->dirT.p, c->uvN.p, L.b.workCounter.p, L.st);
        if (mark) RTB_CUDA(c, cudaEventRecord(c->ev[3], L.st));
        return RTB_OK;
    }
    launch_raygen(L.fm, &c->camera, seedFor(c), L.b.rays.p, L.st);
    if (mark) RTB_CUDA(c, cudaEventRecord(c->ev[2], L.st));
    launch_trace_closest(sv, L.b.rays.p, L.fm.localSlots, L.b.hits.p, L.b.workCounter.p, c->countersOn ? c->counters.p : nullptr, c->lastPrimaryPackets, L.st);
    if (mark) RTB_CUDA(c, cudaEventRecord(c->ev[3], L.st));
    const int li = lane == 1 ? 1 : 0;
    const PrimHit *ws = nullptr, *wc = nullptr;
    if (sv.sphereTree) {
        RTB_CUDA(c, c->primHitS[li].alloc(L.fm.localSlots));
        launch_prims_closest(0, sv, primTreeOf(c, 0), L.b.rays.p, L.fm.localSlots, nullptr, L.b.hits.p, nullptr, c->primHitS[li].p, L.st);
        ws = c->primHitS[li].p;
    }
    if (sv.cubeTree) {
        RTB_CUDA(c, c->primHitC[li].alloc(L.fm.localSlots));
        launch_prims_closest(1, sv, primTreeOf(c, 1), L.b.rays.p, L.fm.localSlots, nullptr, L.b.hits.p, ws, c->primHitC[li].p, L.st);
        wc = c->primHitC[li].p;
    }
    launch_finish_primary(L.fm, sv, L.b.rays.p, L.b.hits.p, c->dirT.p, c->uvN.p, L.st, ws, wc);
    if (c->countersOn) launch_count_hits(L.fm, c->dirT.p, c->counters.p, L.st);
    return RTB_OK;
}

// Light-space binning of the occlusion rays (rtb_sort.cu): a sun gets two axes perpendicular to its direction, a point light
// the octahedral map of the direction from the light; the cell grid covers the scene bounds seen from the light.
RayBin shadowBin(const rtb_ctx* c, uint32_t maxRays) {
    RayBin b{};
    uint32_t bits = 4;
    while (bits < 10 && (1ull << (2 * bits)) * 4ull < maxRays) ++bits;   // about four rays per cell
    b.bits = bits;
    const float cells = (float)(1u << bits);
    const LightRec& l = c->light0;
    if ((l.colorBType >> 16) == LIGHT_POINT) {
        b.kind = 2; b.lpos[0] = l.pos[0]; b.lpos[1] = l.pos[1]; b.lpos[2] = l.pos[2];
        b.u0 = -1.0f; b.v0 = -1.0f; b.su = b.sv = cells * 0.5f;
        return b;
    }
    // decodeNormal (SH/primitive.glsl:90-93); any vector near the light direction will do for a sort key
    float d[3] = {(float)(l.dir[0] >> 16) / 65535.0f * 2.0f - 1.0f, (float)(l.dir[0] & 65535u) / 65535.0f * 2.0f - 1.0f, (float)l.dir[1] / 65535.0f * 2.0f - 1.0f};
    const float len = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    if (!(len > 1e-6f)) return b;   // kind 0: no binning
    for (float& v : d) v /= len;
    const int k = std::fabs(d[0]) < std::fabs(d[1]) ? (std::fabs(d[0]) < std::fabs(d[2]) ? 0 : 2) : (std::fabs(d[1]) < std::fabs(d[2]) ? 1 : 2);
    float e[3] = {0, 0, 0}; e[k] = 1.0f;
    float b1[3] = {d[1] * e[2] - d[2] * e[1], d[2] * e[0] - d[0] * e[2], d[0] * e[1] - d[1] * e[0]};
    const float l1 = std::sqrt(b1[0] * b1[0] + b1[1] * b1[1] + b1[2] * b1[2]);
    for (float& v : b1) v /= l1;
    const float b2[3] = {d[1] * b1[2] - d[2] * b1[1], d[2] * b1[0] - d[0] * b1[2], d[0] * b1[1] - d[1] * b1[0]};
    float lo[2] = {3e38f, 3e38f}, hi[2] = {-3e38f, -3e38f};
    for (int corner = 0; corner < 8; ++corner) {
        const float p[3] = {corner & 1 ? c->stats.hi[0] : c->stats.lo[0], corner & 2 ? c->stats.hi[1] : c->stats.lo[1], corner & 4 ? c->stats.hi[2] : c->stats.lo[2]};
        const float u = p[0] * b1[0] + p[1] * b1[1] + p[2] * b1[2], v = p[0] * b2[0] + p[1] * b2[1] + p[2] * b2[2];
        lo[0] = std::min(lo[0], u); hi[0] = std::max(hi[0], u); lo[1] = std::min(lo[1], v); hi[1] = std::max(hi[1], v);
    }
    if (!(hi[0] > lo[0]) || !(hi[1] > lo[1])) return b;
    b.kind = 1;
    for (int a = 0; a < 3; ++a) { b.b1[a] = b1[a]; b.b2[a] = b2[a]; }
    b.u0 = lo[0]; b.v0 = lo[1]; b.su = cells / (hi[0] - lo[0]); b.sv = cells / (hi[1] - lo[1]);
    return b;
}

uint32_t shadowLayers(const rtb_ctx* c) { return c->samples * (c->lightsOpt ? std::max(c->info.lightCount, 1u) : 1u); }

// the shadow-word buffer holds `layers` layers (samples, or lights x samples with RTB_OPT_LIGHTS); grown when a scene gains lights
int ensureShadowWords(rtb_ctx* c) {
    const uint32_t layers = shadowLayers(c);
    if (layers == c->bitsLayers && c->bits.p) return RTB_OK;
    const size_t words = shadowWords(c->width, c->height, layers);
    if (words > c->bits.count) {
        if (c->frontStream) { RTB_CUDA(c, cudaStreamSynchronize(c->frontStream)); RTB_CUDA(c, cudaStreamSynchronize(c->midStream)); c->frontDirty = false; }
        RTB_CUDA(c, cudaStreamSynchronize(c->stream));
        RTB_CUDA(c, c->bits.alloc(words));
        releaseAlt(c);   // the other set's shadow words are too small as well: reallocated when next needed
    }
    RTB_CUDA(c, cudaMemsetAsync(c->bits.p, 0, c->bits.bytes(), c->stream));
    c->bitsLayers = layers;
    return RTB_OK;
}

int lightsView(rtb_ctx* c, LightsView& lv, bool forLighting) {
    lv = LightsView{};
    lv.mode = c->lightsOpt && c->info.lightCount ? c->lightsOpt : 0u;
    lv.tilesX = (c->width + 15u) / 16u;
    lv.lightBegin = 0; lv.lightEnd = c->info.lightCount;
    if (lv.mode == 2u) { lv.tileCount = c->lightTileCount.p; lv.tileList = c->lightTileList.p; }
    if (forLighting && c->lightCacheOpt && c->width && c->info.lightCount) {
        // kind 2 (the direction itself) for the reference's one directional light, kind 1 (the random pair) otherwise; none beyond 1.5 GB
        rtb_ctx::LightCacheKey key;
        key.w = c->width; key.h = c->height; key.samples = c->samples; key.rank = c->tileRank; key.count = c->tileCount;
        key.kind = (!lv.mode && (c->light0.colorBType >> 16) != LIGHT_POINT) ? 2u : 1u;
        key.lightEpoch = key.kind == 2u ? c->light0Epoch : 0u;
        const size_t entries = (size_t)c->width * c->height * c->samples;
        if (entries * sizeof(float4) <= (3ull << 29)) {
            if (!(key == c->lightCacheKey) || !c->lightCache.p) {
                if (!c->capturing) {   // (a recorded frame follows a direct one with the same state: the cache is there by then)
                    if (c->lightCache.count < entries) { RTB_CUDA(c, cudaStreamSynchronize(c->stream)); RTB_CUDA(c, c->lightCache.alloc(entries)); }
                    launch_light_cache(c->fm, sceneView(c), c->samples, key.kind, c->lightCache.p, c->stream);
                    c->lightCacheKey = key;
                    ++c->stamp;   // recorded launches hold the cache kind by value
                }
            }
            if (key == c->lightCacheKey && c->lightCache.p) { lv.cacheKind = key.kind; lv.lightCache = c->lightCache.p; }
        }
    }
    if (forLighting && c->historyAlpha > 0.0f) {
        const size_t px = (size_t)c->width * c->height;
        if (c->history.count < px) { RTB_CUDA(c, cudaStreamSynchronize(c->stream)); RTB_CUDA(c, c->history.alloc(px)); c->historyValid = false; }
        lv.history = c->history.p;
        lv.historyAlpha = c->historyValid ? c->historyAlpha : 1.0f;   // the first frame of a sequence starts the history
    }
    return RTB_OK;
}

// the shadow words start from zero (DEBUG build) or keep what they held where no subgroup has a hit (RELEASE): once per frame,
// before any lane's rays are traced
int clearShadowBits(rtb_ctx* c) {
    { const int rc = waitCopy(c, {RTB_TGT_SHADOW_BITS}); if (rc) return rc; }
    { const int rc = ensureShadowWords(c); if (rc) return rc; }
    if (c->releaseBuild) launch_clear_hit_strips(c->fm, c->dirT.p, shadowLayers(c), c->bits.p, c->stream);
    else RTB_CUDA(c, cudaMemsetAsync(c->bits.p, 0, (size_t)shadowWords(c->width, c->height, shadowLayers(c)) * 4, c->stream));
    return RTB_OK;
}

// occlusion by spheres / cubes that have a tree of their own: the same rays, after the triangles' any-hit launch
void shadowPrimPasses(rtb_ctx* c, const SceneView& sv, const FrameMap& fm, const RayRec* rays, uint32_t n, const uint32_t* countPtr, const uint32_t* slotIds, cudaStream_t st) {
    if (sv.sphereTree) launch_prims_any(0, sv, primTreeOf(c, 0), rays, n, countPtr, nullptr, c->bits.p, slotIds, fm, st);
    if (sv.cubeTree) launch_prims_any(1, sv, primTreeOf(c, 1), rays, n, countPtr, nullptr, c->bits.p, slotIds, fm, st);
}

int passShadow(rtb_ctx* c, bool mark, int lane = -1) {
    // (the RELEASE build's clear reads the G-buffer, so with lanes it runs after both lanes' nearest-hit launches: see the frame dispatch)
    if (lane < 0) { const int rc = clearShadowBits(c); if (rc) return rc; }
    const LaneRef L = laneOf(c, lane);
    rtb_ctx::LaneBufs& B = L.b;
    const SceneView sv = sceneView(c);
    TraceCounters* counters = c->countersOn ? c->counters.p + 1 : nullptr;
    if (c->lightsOpt && sv.info.lightCount) {
        // ---- every light (beyond the reference): the queue is filled and traced in chunks of lights so that it stays bounded ----
        LightsView lv;
        { const int rc = lightsView(c, lv, false); if (rc) return rc; }
        if (lv.mode == 2u) {
            const uint32_t tiles = lv.tilesX * ((c->height + 15u) / 16u);
            RTB_CUDA(c, c->lightTileCount.alloc(tiles)); RTB_CUDA(c, c->lightTileList.alloc((size_t)tiles * LIGHTS_PER_TILE));
            lv.tileCount = c->lightTileCount.p; lv.tileList = c->lightTileList.p;
            launch_light_tiles(L.fm, sv, &c->camera, c->dirT.p, c->lightTileCount.p, c->lightTileList.p, L.st);
        }
        const uint32_t perLight = L.fm.localSlots * c->samples;
        if (!perLight) return RTB_OK;
        const uint32_t lightsPerChunk = std::max(1u, std::min(sv.info.lightCount, (64u << 20) / perLight));
        const uint32_t capacity = lightsPerChunk * perLight;
        RTB_CUDA(c, B.queueCount.alloc(1)); RTB_CUDA(c, B.queueSlots.alloc(capacity)); RTB_CUDA(c, B.queueRays.alloc(capacity));
        RayQueue q{};
        q.rays = B.queueRays.p; q.slotIds = B.queueSlots.p; q.count = B.queueCount.p;
        if (mark) RTB_CUDA(c, cudaEventRecord(c->ev[5], L.st));
        for (uint32_t first = 0; first < sv.info.lightCount; first += lightsPerChunk) {
            lv.lightBegin = first; lv.lightEnd = std::min(sv.info.lightCount, first + lightsPerChunk);
            RTB_CUDA(c, cudaMemsetAsync(B.queueCount.p, 0, 4, L.st));
            launch_shadowgen_lights(L.fm, sv, &c->camera, seedFor(c), c->samples, c->dirT.p, c->bits.p, q, lv, L.st);
            launch_trace_any_bits(L.fm, sv, q.rays, capacity, c->bits.p, B.workCounter.p, counters, q.slotIds, q.count, L.st);
            shadowPrimPasses(c, sv, L.fm, q.rays, capacity, q.count, q.slotIds, L.st);
        }
        return RTB_OK;
    }
    const uint32_t maxRays = L.fm.localSlots * c->samples;
    if (!c->shadowOrder || sv.useBvh != ACCEL_KIND_CWBVH || !sv.info.triangleCount || !maxRays) {
        // slot order: one record per (sample, slot), what the first-generation kernels and the reference loop consume
        launch_shadowgen(L.fm, sv, &c->camera, seedFor(c), c->samples, c->dirT.p, B.rays.p, c->bits.p, nullptr, nullptr, L.st);
        if (mark) RTB_CUDA(c, cudaEventRecord(c->ev[5], L.st));
        launch_trace_any_bits(L.fm, sv, B.rays.p, maxRays, c->bits.p, B.workCounter.p, counters, nullptr, nullptr, L.st);
        shadowPrimPasses(c, sv, L.fm, B.rays.p, maxRays, nullptr, nullptr, L.st);
        return RTB_OK;
    }
    // queue of live rays, optionally sorted in light space
    RayBin bin = c->shadowOrder >= 2 && !(c->countersOn && c->countersMode == 1) ? shadowBin(c, maxRays) : RayBin{};
    const uint32_t cells = bin.kind ? 1u << (2 * bin.bits) : 0u;
    RTB_CUDA(c, B.queueCount.alloc(1)); RTB_CUDA(c, B.queueSlots.alloc(maxRays));
    RTB_CUDA(c, cudaMemsetAsync(B.queueCount.p, 0, 4, L.st));
    RayQueue q{};
    q.count = B.queueCount.p; q.slotIds = B.queueSlots.p;
    if (bin.kind) {
        RTB_CUDA(c, B.queueRays.alloc(maxRays)); RTB_CUDA(c, B.sortedSlots.alloc(maxRays)); RTB_CUDA(c, B.queueCell.alloc(maxRays)); RTB_CUDA(c, B.queueRank.alloc(maxRays));
        RTB_CUDA(c, B.queueHist.alloc(cells)); RTB_CUDA(c, B.queueSums.alloc(1024));
        RTB_CUDA(c, cudaMemsetAsync(B.queueHist.p, 0, (size_t)cells * 4, L.st));
        q.rays = B.queueRays.p; q.cell = B.queueCell.p; q.rank = B.queueRank.p; q.hist = B.queueHist.p; q.blockSums = B.queueSums.p;
    } else
        q.rays = B.rays.p;
    launch_shadowgen(L.fm, sv, &c->camera, seedFor(c), c->samples, c->dirT.p, B.rays.p, c->bits.p, &q, &bin, L.st);
    const uint32_t* slots = B.queueSlots.p;
    if (bin.kind) { launch_sort_rays(q, cells, maxRays, B.rays.p, B.sortedSlots.p, L.st); slots = B.sortedSlots.p; }
    if (mark) RTB_CUDA(c, cudaEventRecord(c->ev[5], L.st));
    if (bin.kind && c->shadowOrder == 3) {   // beam packets over the sorted queue; the pre-sort buffers are free again and take the fall-back rays
        RTB_CUDA(c, B.queueFallbackCount.alloc(1));
        RayQueue fb{};
        fb.rays = B.queueRays.p; fb.slotIds = B.queueSlots.p; fb.count = B.queueFallbackCount.p;
        launch_trace_beam_bits(L.fm, sv, B.rays.p, maxRays, c->bits.p, B.workCounter.p, counters, slots, B.queueCount.p, fb, L.st);
    } else
        launch_trace_any_bits(L.fm, sv, B.rays.p, maxRays, c->bits.p, B.workCounter.p, counters, slots, B.queueCount.p, L.st);
    shadowPrimPasses(c, sv, L.fm, B.rays.p, maxRays, B.queueCount.p, slots, L.st);
    return RTB_OK;
}
int passShade(rtb_ctx* c, int what, int lane = -1) {
    if (lane < 0) { const int rc = waitCopy(c, {RTB_TGT_LIGHTING, RTB_TGT_ACCUM, RTB_TGT_RGBA8, RTB_TGT_RGBA8_TILED}); if (rc) return rc; }
    const LaneRef L = laneOf(c, lane);
    const SceneView sv = sceneView(c);
    LightsView lv;
    { const int rc = lightsView(c, lv, what != SHADE_COMPOSITE); if (rc) return rc; }
    launch_shade(what, L.fm, sv, &c->camera, seedFor(c), c->samples, c->dirT.p, c->uvN.p, c->bits.p, c->lighting.p, c->accum.p,
                 c->rgba8.p, c->tileCount > 1 ? c->rgba8Tiled.p : nullptr, L.st, &lv);
    if (lv.history && !c->historyValid) { c->historyValid = true; ++c->stamp; }   // the recorded launches hold alpha by value
    return RTB_OK;
}

// a frame as two half-frame lanes on two streams: while a persistent launch of one lane drains (its last rays are the longest),
// the other lane's launch takes the freed SM slots.  Same kernels, same pixels; only the launches' slot ranges differ.
bool useLanes(const rtb_ctx* c) {
    return c->lanesOpt == 2 && !c->countersOn && !c->lightsOpt && !(c->historyAlpha > 0.0f) && c->laneFm[1].localSlots >= 64u * 1024u;   // >= 64 blocks per lane: smaller frames are launch-bound
}
int prepareLanes(rtb_ctx* c) {
    if (c->laneStream) return RTB_OK;
    RTB_CUDA(c, cudaStreamCreateWithFlags(&c->laneStream, cudaStreamNonBlocking));
    RTB_CUDA(c, cudaEventCreateWithFlags(&c->evFork, cudaEventDisableTiming));
    RTB_CUDA(c, cudaEventCreateWithFlags(&c->evJoin, cudaEventDisableTiming));
    return RTB_OK;
}
int forkLanes(rtb_ctx* c) {
    RTB_CUDA(c, cudaEventRecord(c->evFork, c->stream));
    RTB_CUDA(c, cudaStreamWaitEvent(c->laneStream, c->evFork, 0));
    return RTB_OK;
}
int joinLanes(rtb_ctx* c) {
    RTB_CUDA(c, cudaEventRecord(c->evJoin, c->laneStream));
    RTB_CUDA(c, cudaStreamWaitEvent(c->stream, c->evJoin, 0));
    return RTB_OK;
}

// RTB_PASS_FRAME, part A: init .. occlusion.  Everything here is stream-ordered work without host synchronisation or waits on
// outside events (the caller has done those), so that it can be captured into a CUDA graph.  `mark`: phase events (one lane only).
int framePartA(rtb_ctx* c, bool mark) {
    int rc;
    if (mark) RTB_CUDA(c, cudaEventRecord(c->ev[0], c->stream));
    launch_init(c->seed.p, c->stream);
    if (mark) RTB_CUDA(c, cudaEventRecord(c->ev[1], c->stream));
    if (!useLanes(c)) {
        if ((rc = passRaygen(c, mark))) return rc;
        if (mark) RTB_CUDA(c, cudaEventRecord(c->ev[4], c->stream));
        if ((rc = passShadow(c, mark))) return rc;
        if (mark) RTB_CUDA(c, cudaEventRecord(c->ev[6], c->stream));
        return RTB_OK;
    }
    // two half-frame lanes on two streams: while a persistent launch of one lane drains, the other lane's launch takes the freed SM slots
    if (!c->releaseBuild) { if ((rc = clearShadowBits(c))) return rc; }
    if ((rc = forkLanes(c))) return rc;
    for (int lane = 0; lane < 2; ++lane) if ((rc = passRaygen(c, false, lane))) return rc;
    if (c->releaseBuild) {   // the RELEASE clear reads the whole G-buffer: join, clear, fork again
        if ((rc = joinLanes(c))) return rc;
        if ((rc = clearShadowBits(c))) return rc;
        if ((rc = forkLanes(c))) return rc;
    }
    for (int lane = 0; lane < 2; ++lane) if ((rc = passShadow(c, false, lane))) return rc;
    return joinLanes(c);
}
// part B: lighting + composite
int framePartB(rtb_ctx* c, bool mark) {
    int rc;
    if (!useLanes(c)) {
        if ((rc = passShade(c, SHADE_BOTH))) return rc;
        if (mark) RTB_CUDA(c, cudaEventRecord(c->ev[7], c->stream));
        return RTB_OK;
    }
    { LightsView lv; if ((rc = lightsView(c, lv, true))) return rc; }   // (fills the light cache, if it must, before the lanes fork)
    if ((rc = forkLanes(c))) return rc;
    for (int lane = 0; lane < 2; ++lane) if ((rc = passShade(c, SHADE_BOTH, lane))) return rc;
    return joinLanes(c);
}

int captureFrame(rtb_ctx* c, int (*part)(rtb_ctx*, bool), cudaGraphExec_t* out) {
    if (*out) { cudaGraphExecDestroy(*out); *out = nullptr; }
    RTB_CUDA(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeRelaxed));
    c->capturing = true;
    const int rc = part(c, false);
    c->capturing = false;
    cudaGraph_t g = nullptr;
    const cudaError_t e = cudaStreamEndCapture(c->stream, &g);
    if (rc) { if (g) cudaGraphDestroy(g); return rc; }
    if (e != cudaSuccess) return cudaFail(c, e, "cudaStreamEndCapture");
    const cudaError_t e2 = cudaGraphInstantiate(out, g, 0);
    cudaGraphDestroy(g);
    if (e2 != cudaSuccess) { *out = nullptr; return cudaFail(c, e2, "cudaGraphInstantiate"); }
    return RTB_OK;
}

// ---- frame overlap -------------------------------------------------------------------------------------------------------------
// (not with the RELEASE shader build: its skipped stores leave what the SAME texture held a frame earlier, so the G-buffer must be one)
bool useOverlap(const rtb_ctx* c) { return c->overlapOpt && c->graphOpt && !c->countersOn && !c->releaseBuild && !useLanes(c) && c->fm.localSlots; }

int overlapFront(rtb_ctx* c, bool) {
    launch_init(c->seed.p, c->stream, c->seedUse);
    return passRaygen(c, false);
}
int overlapShadow(rtb_ctx* c, bool) { return passShadow(c, false); }
int overlapShade(rtb_ctx* c, bool) { return passShade(c, SHADE_BOTH); }

void swapSets(rtb_ctx* c) {
    std::swap(c->dirT, c->alt.dirT); std::swap(c->uvN, c->alt.uvN); std::swap(c->lane[0], c->alt.lane);
    std::swap(c->primHitS[0], c->alt.primHitS); std::swap(c->primHitC[0], c->alt.primHitC);
    std::swap(c->bits, c->alt.bits);
    c->setIndex ^= 1;
}
void releaseAlt(rtb_ctx* c) {
    c->alt.dirT.release(); c->alt.uvN.release(); c->alt.lane.release(); c->alt.primHitS.release(); c->alt.primHitC.release(); c->alt.bits.release();
    c->altAllocated = false;
    for (int k = 0; k < 2; ++k) c->ovStamp[k] = 0;
}
int prepareOverlap(rtb_ctx* c) {
    if (!c->frontStream) {
        RTB_CUDA(c, cudaStreamCreateWithFlags(&c->frontStream, cudaStreamNonBlocking));
        RTB_CUDA(c, cudaStreamCreateWithFlags(&c->midStream, cudaStreamNonBlocking));
        for (int k = 0; k < 2; ++k) {
            RTB_CUDA(c, cudaEventCreateWithFlags(&c->evMidDone[k], cudaEventDisableTiming));
            RTB_CUDA(c, cudaEventCreateWithFlags(&c->evFrontDone[k], cudaEventDisableTiming));
            RTB_CUDA(c, cudaEventCreateWithFlags(&c->evBackDone[k], cudaEventDisableTiming));
        }
        RTB_CUDA(c, cudaEventCreateWithFlags(&c->evJoinF, cudaEventDisableTiming));
        RTB_CUDA(c, cudaEventCreateWithFlags(&c->evJoinB, cudaEventDisableTiming));
        RTB_CUDA(c, c->seedSnap.alloc(2));
    }
    if (!c->altAllocated) {
        const size_t px = (size_t)c->width * c->height;
        RTB_CUDA(c, c->alt.dirT.alloc(px)); RTB_CUDA(c, c->alt.uvN.alloc(px));
        RTB_CUDA(c, c->alt.lane.rays.alloc((size_t)c->fm.localSlots * (c->samples ? c->samples : 1)));
        RTB_CUDA(c, c->alt.lane.hits.alloc(c->fm.localSlots));
        RTB_CUDA(c, c->alt.lane.workCounter.alloc(1));
        RTB_CUDA(c, c->alt.bits.alloc(c->bits.count));   // (the shadow words of a set: written by its shadow pass, read by its shade pass)
        RTB_CUDA(c, cudaMemsetAsync(c->alt.bits.p, 0, c->alt.bits.bytes(), c->stream));
        RTB_CUDA(c, cudaMemsetAsync(c->alt.dirT.p, 0, c->alt.dirT.bytes(), c->stream));   // pixels of other ranks stay defined
        RTB_CUDA(c, cudaMemsetAsync(c->alt.uvN.p, 0, c->alt.uvN.bytes(), c->stream));
        RTB_CUDA(c, cudaMemsetAsync(c->alt.lane.workCounter.p, 0, 4, c->stream));
        c->altAllocated = true;
        c->frontNeedsBack = true;
    }
    return RTB_OK;
}
int captureOn(rtb_ctx* c, cudaStream_t st, int (*part)(rtb_ctx*, bool), cudaGraphExec_t* out) {
    cudaStream_t keep = c->stream;
    c->stream = st;
    const int rc = captureFrame(c, part, out);
    c->stream = keep;
    return rc;
}
// one overlapped frame: swap the sets, FRONT on the front stream, BACK on the context's stream
int overlappedFrame(rtb_ctx* c) {
    int rc;
    if ((rc = prepareOverlap(c))) return rc;
    swapSets(c);
    const int s = c->setIndex;
    if (c->ovStamp[s] != c->stamp || !c->ovFront[s]) {
        c->seedUse = c->seedSnap.p + s;
        rc = captureOn(c, c->frontStream, overlapFront, &c->ovFront[s]);
        if (!rc) rc = captureFrame(c, overlapShadow, &c->ovShadow[s]);
        if (!rc) rc = captureFrame(c, overlapShade, &c->ovShade[s]);
        c->seedUse = nullptr;
        if (rc) { c->ovStamp[s] = 0; return rc; }
        c->ovStamp[s] = c->stamp;
    }
    // FRONT: after what the context's stream holds for it, after the BACK that last read this set, after a read-back of its targets
    if (c->frontNeedsBack) {
        RTB_CUDA(c, cudaEventRecord(c->evJoinB, c->stream));
        RTB_CUDA(c, cudaStreamWaitEvent(c->frontStream, c->evJoinB, 0));
        c->frontNeedsBack = false;
    } else if (c->backDoneSet[s])
        RTB_CUDA(c, cudaStreamWaitEvent(c->frontStream, c->evBackDone[s], 0));
    if (c->copyTarget == RTB_TGT_SEED || c->copyTarget == RTB_TGT_DIR_T || c->copyTarget == RTB_TGT_UV_NORMAL) {
        RTB_CUDA(c, cudaStreamWaitEvent(c->frontStream, c->evCopied, 0));
        RTB_CUDA(c, cudaStreamWaitEvent(c->stream, c->evCopied, 0));
        c->copyTarget = -1;
    }
    RTB_CUDA(c, cudaGraphLaunch(c->ovFront[s], c->frontStream));
    RTB_CUDA(c, cudaEventRecord(c->evFrontDone[s], c->frontStream));
    // MID: shadow rays + occlusion on a stream of their own (so that they run beside the shade launch of the frame before) unless the
    // shadow and the shade pass share more than the set's shadow words (the tile light lists of RTB_OPT_LIGHTS); BACK: shade
    const bool three = c->lightsOpt == 0u;
    cudaStream_t ms = three ? c->midStream : c->stream;
    RTB_CUDA(c, cudaStreamWaitEvent(ms, c->evFrontDone[s], 0));
    c->frontDirty = false;
    if (c->copyTarget == RTB_TGT_SHADOW_BITS) { RTB_CUDA(c, cudaStreamWaitEvent(ms, c->evCopied, 0)); c->copyTarget = -1; }
    RTB_CUDA(c, cudaGraphLaunch(c->ovShadow[s], ms));
    if (three) {
        RTB_CUDA(c, cudaEventRecord(c->evMidDone[s], ms));
        RTB_CUDA(c, cudaStreamWaitEvent(c->stream, c->evMidDone[s], 0));   // (and with it everything the front stream held)
    }
    if ((rc = waitCopy(c, {RTB_TGT_LIGHTING, RTB_TGT_ACCUM, RTB_TGT_RGBA8, RTB_TGT_RGBA8_TILED}))) return rc;
    RTB_CUDA(c, cudaGraphLaunch(c->ovShade[s], c->stream));
    RTB_CUDA(c, cudaEventRecord(c->evBackDone[s], c->stream));
    c->backDoneSet[s] = true;
    c->frameTimed = false;
    return RTB_OK;
}

}  // namespace

extern "C" {

int rtb_create(rtb_ctx** out, int cudaDevice, const rtb_limits* limits) {
    if (!out || !limits) return fail(nullptr, RTB_ERR_ARG, "rtb_create: null argument");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) return fail(nullptr, RTB_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e));
    if (cudaDevice < 0 || cudaDevice >= n) return fail(nullptr, RTB_ERR_ARG, "rtb_create: device index out of range");
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, cudaDevice)) != cudaSuccess) return cudaFail(nullptr, e, "cudaGetDeviceProperties");
    if (prop.major != 10) return fail(nullptr, RTB_ERR_CUDA, "librtb200 is built for sm_100a only; device is sm_" + std::to_string(prop.major * 10 + prop.minor));
    if ((e = cudaSetDevice(cudaDevice)) != cudaSuccess) return cudaFail(nullptr, e, "cudaSetDevice");
    rtb_ctx* c = new rtb_ctx();
    c->device = cudaDevice; c->limits = *limits;
    auto bail = [&](cudaError_t err, const char* what) { const int rc = cudaFail(nullptr, err, what); rtb_destroy(c); return rc; };
    if ((e = cudaStreamCreateWithFlags(&c->ownStream, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
    c->stream = c->ownStream;
    for (auto& ev : c->ev) if ((e = cudaEventCreate(&ev)) != cudaSuccess) return bail(e, "cudaEventCreate");
    if ((e = c->triangles.alloc(limits->max_triangles)) != cudaSuccess) return bail(e, "alloc triangles");
    if ((e = c->spheres.alloc(limits->max_spheres)) != cudaSuccess) return bail(e, "alloc spheres");
    if ((e = c->cubes.alloc((size_t)limits->max_cubes * 6)) != cudaSuccess) return bail(e, "alloc cubes");
    if ((e = c->planes.alloc(limits->max_planes)) != cudaSuccess) return bail(e, "alloc planes");
    if ((e = c->lights.alloc(limits->max_lights)) != cudaSuccess) return bail(e, "alloc lights");
    if ((e = c->materials.alloc(limits->max_materials)) != cudaSuccess) return bail(e, "alloc materials");
    const size_t objects = (size_t)limits->max_triangles + limits->max_spheres + limits->max_cubes + limits->max_planes;
    if ((e = c->materialIndices.alloc(objects)) != cudaSuccess) return bail(e, "alloc materialIndices");
    if ((e = c->seed.alloc(1)) != cudaSuccess) return bail(e, "alloc seed");
    if ((e = c->lane[0].workCounter.alloc(1)) != cudaSuccess) return bail(e, "alloc workCounter");
    if ((e = c->counters.alloc(2)) != cudaSuccess) return bail(e, "alloc counters");
    if ((e = cudaMemsetAsync(c->lights.p, 0, c->lights.bytes(), c->stream)) != cudaSuccess) return bail(e, "cudaMemsetAsync");
    if ((e = cudaMemsetAsync(c->materials.p, 0, c->materials.bytes(), c->stream)) != cudaSuccess) return bail(e, "cudaMemsetAsync");
    if ((e = cudaMemsetAsync(c->materialIndices.p, 0, c->materialIndices.bytes(), c->stream)) != cudaSuccess) return bail(e, "cudaMemsetAsync");
    if ((e = cudaMemsetAsync(c->seed.p, 0, sizeof(SeedRec), c->stream)) != cudaSuccess) return bail(e, "cudaMemsetAsync");
    if ((e = cudaMemsetAsync(c->counters.p, 0, 2 * sizeof(TraceCounters), c->stream)) != cudaSuccess) return bail(e, "cudaMemsetAsync");
    c->triangleMirror.resize((size_t)limits->max_triangles * sizeof(TriangleRec));
    *out = c;
    return RTB_OK;
}

void rtb_destroy(rtb_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->frontStream) cudaStreamSynchronize(c->frontStream);
    if (c->midStream) cudaStreamSynchronize(c->midStream);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->copyStream) { cudaStreamSynchronize(c->copyStream); cudaStreamDestroy(c->copyStream); }
    for (int k = 0; k < 2; ++k) {
        if (c->ovFront[k]) cudaGraphExecDestroy(c->ovFront[k]);
        if (c->ovShadow[k]) cudaGraphExecDestroy(c->ovShadow[k]);
        if (c->ovShade[k]) cudaGraphExecDestroy(c->ovShade[k]);
        if (c->evFrontDone[k]) cudaEventDestroy(c->evFrontDone[k]);
        if (c->evMidDone[k]) cudaEventDestroy(c->evMidDone[k]);
        if (c->evBackDone[k]) cudaEventDestroy(c->evBackDone[k]);
    }
    if (c->evJoinF) cudaEventDestroy(c->evJoinF);
    if (c->evJoinB) cudaEventDestroy(c->evJoinB);
    if (c->frontStream) cudaStreamDestroy(c->frontStream);
    if (c->midStream) cudaStreamDestroy(c->midStream);
    c->alt.dirT.release(); c->alt.uvN.release(); c->alt.lane.release(); c->alt.primHitS.release(); c->alt.primHitC.release(); c->alt.bits.release(); c->seedSnap.release();
    if (c->evReady) cudaEventDestroy(c->evReady);
    if (c->evCopied) cudaEventDestroy(c->evCopied);
    c->triangles.release(); c->spheres.release(); c->cubes.release(); c->planes.release(); c->lights.release(); c->materials.release();
    c->materialIndices.release(); c->skybox.release(); c->seed.release(); c->nodes.release(); c->nodes8.release(); c->travTris.release(); c->nodeBox.release(); c->maxBits.release(); c->areaSums.release();
    c->dirT.release(); c->uvN.release(); c->accum.release(); c->lighting.release(); c->bits.release(); c->rgba8.release(); c->rgba8Tiled.release();
    c->lane[0].release(); c->lane[1].release(); c->counters.release();
    c->lightTileCount.release(); c->lightTileList.release(); c->history.release(); c->lightCache.release(); c->sunFrame.release();
    for (int k = 0; k < 2; ++k) { c->primTree[k].release(); c->primHitS[k].release(); c->primHitC[k].release(); }
    c->rinPrimS.release(); c->rinPrimC.release();
    if (c->graphA) cudaGraphExecDestroy(c->graphA);
    if (c->graphB) cudaGraphExecDestroy(c->graphB);
    if (c->laneStream) cudaStreamDestroy(c->laneStream);
    if (c->evFork) cudaEventDestroy(c->evFork);
    if (c->evJoin) cudaEventDestroy(c->evJoin);
    c->pathT.release(); c->pathL.release(); c->pathDirect.release(); c->pathRays[0].release(); c->pathRays[1].release(); c->pathShadowRays.release();
    c->pathSlots[0].release(); c->pathSlots[1].release(); c->pathShadowSlots.release(); c->pathCounts.release(); c->pathOccA.release(); c->pathOccB.release();
    for (auto& ev : c->pathEv) if (ev) cudaEventDestroy(ev);
    c->rinRays.release(); c->rinHits.release(); c->rinObj.release(); c->rinT.release(); c->rinUv.release(); c->rinOcc.release(); c->rinOcc2.release();
    for (auto& ev : c->ev) if (ev) cudaEventDestroy(ev);
    if (c->ownStream) cudaStreamDestroy(c->ownStream);
    delete c;
}

const char* rtb_last_error(const rtb_ctx* c) { return c ? c->error.c_str() : g_createError.c_str(); }

int rtb_set_option(rtb_ctx* c, rtb_option opt, uint32_t value) {
    if (!c) return RTB_ERR_ARG;
    ++c->stamp;   // recorded frames hold the options' effects
    switch (opt) {
        case RTB_OPT_COUNTERS: c->countersOn = value != 0; c->countersMode = value; return RTB_OK;
        case RTB_OPT_FUSE_PRIMARY: c->fuseOpt = value ? 1u : 0u; return RTB_OK;
        case RTB_OPT_PRIMARY_PACKETS:
            if (value > 3) return fail(c, RTB_ERR_ARG, "RTB_OPT_PRIMARY_PACKETS: 0 off, 1 union packets, 2 auto, 3 frustum packets");
            c->packetsOpt = value; return RTB_OK;
        case RTB_OPT_TILE_RANK:
        case RTB_OPT_TILE_COUNT: {
            // count is set before rank when both change (rtb.py, bench.py, the facade): a rank is valid against the count in force
            const uint32_t rank = opt == RTB_OPT_TILE_RANK ? value : (value > c->tileRank ? c->tileRank : 0u);
            const uint32_t count = opt == RTB_OPT_TILE_COUNT ? value : c->tileCount;
            if (!count) return fail(c, RTB_ERR_ARG, "tile count must be >= 1");
            if (rank >= count) return fail(c, RTB_ERR_ARG, "RTB_OPT_TILE_RANK must be below RTB_OPT_TILE_COUNT (set the count first)");
            const uint32_t oldRank = c->tileRank, oldCount = c->tileCount;
            c->tileRank = rank; c->tileCount = count;
            if (c->width) {
                RTB_BIND(c);
                { const int rc = quiesce(c); if (rc) return rc; }
                RTB_CUDA(c, cudaStreamSynchronize(c->stream));
                const int rc = allocFrame(c);
                if (rc) { c->tileRank = oldRank; c->tileCount = oldCount; c->width = c->height = 0; makeFrameMap(c); return rc; }   // frame resources are gone: resize again
            }
            return RTB_OK;
        }
        case RTB_OPT_PRIMITIVE_TREES: c->primTreeMin = value; return RTB_OK;
        case RTB_OPT_LIGHTS:
            if (value > 2) return fail(c, RTB_ERR_ARG, "RTB_OPT_LIGHTS: 0 reference (light 0 x lightCount), 1 all lights, 2 all lights through tile lists");
            c->lightsOpt = value; c->historyValid = false; return RTB_OK;
        case RTB_OPT_HISTORY_ALPHA: {
            float a; std::memcpy(&a, &value, 4);
            if (!(a >= 0.0f && a <= 1.0f)) return fail(c, RTB_ERR_ARG, "RTB_OPT_HISTORY_ALPHA: the bits of a float in [0, 1]; 0 switches the History blend off");
            c->historyAlpha = a; c->historyValid = false; return RTB_OK;
        }
        case RTB_OPT_FRAME_GRAPH: c->graphOpt = value ? 1u : 0u; return RTB_OK;
        case RTB_OPT_FRAME_OVERLAP: c->overlapOpt = value ? 1u : 0u; return RTB_OK;
        case RTB_OPT_LIGHT_CACHE: c->lightCacheOpt = value ? 1u : 0u; return RTB_OK;
        case RTB_OPT_FRAME_LANES:
            if (value < 1 || value > 2) return fail(c, RTB_ERR_ARG, "RTB_OPT_FRAME_LANES: 1 or 2");
            c->lanesOpt = value; return RTB_OK;
        case RTB_OPT_ACCEL_BUILDER:
            if (value > 2) return fail(c, RTB_ERR_ARG, "RTB_OPT_ACCEL_BUILDER: 0 host builder, 1 device builder, 2 device builder with up to 3 triangles per leaf slot");
            c->builderOpt = value; return RTB_OK;
        case RTB_OPT_SHADOW_ORDER:
            if (value > 3) return fail(c, RTB_ERR_ARG, "RTB_OPT_SHADOW_ORDER: 0 slot order, 1 queue of live rays, 2 queue sorted in light space, 3 sorted queue walked as beam packets");
            c->shadowOrder = value; return RTB_OK;
        case RTB_OPT_SHADER_BUILD:
            if (value > 1) return fail(c, RTB_ERR_ARG, "RTB_OPT_SHADER_BUILD: 0 = DEBUG build (default), 1 = RELEASE build");
            c->releaseBuild = value; return RTB_OK;
    }
    return fail(c, RTB_ERR_ARG, "unknown option");
}

int rtb_set_stream(rtb_ctx* c, void* s) {
    if (!c) return RTB_ERR_ARG;
    RTB_BIND(c);
    { const int rc = quiesce(c); if (rc) return rc; }
    RTB_CUDA(c, cudaStreamSynchronize(c->stream));
    c->stream = s ? static_cast<cudaStream_t>(s) : c->ownStream;
    ++c->stamp;
    return RTB_OK;
}

int rtb_resize(rtb_ctx* c, uint32_t w, uint32_t h, uint32_t shadowSamples) {
    if (!c) return RTB_ERR_ARG;
    if (!w || !h || !shadowSamples || w > 32768 || h > 32768 || shadowSamples > 512) return fail(c, RTB_ERR_ARG, "rtb_resize: size or sample count out of range");
    RTB_BIND(c);
    if (w == c->width && h == c->height && shadowSamples == c->samples) return RTB_OK;
    if ((uint64_t)((w + 31) / 32) * ((h + 31) / 32) * 1024ull * shadowSamples > 0xFFFFFFFFull) return fail(c, RTB_ERR_ARG, "rtb_resize: too many shadow rays per frame");
    { const int rc = quiesce(c); if (rc) return rc; }
    RTB_CUDA(c, cudaStreamSynchronize(c->stream));
    { const int rc = drainCopy(c); if (rc) return rc; }
    c->width = w; c->height = h; c->samples = shadowSamples; c->shadowSamplesProp = shadowSamples;
    ++c->stamp;
    c->historyValid = false; c->bitsLayers = 0;
    const int rc = allocFrame(c);
    if (rc) { c->width = c->height = 0; c->samples = 0; makeFrameMap(c); }   // a failed allocation leaves NO frame: dispatch reports "before rtb_resize"
    return rc;
}

int rtb_upload(rtb_ctx* c, rtb_buffer id, size_t off, size_t bytes, const void* src) {
    if (!c || (!src && bytes)) return c ? fail(c, RTB_ERR_ARG, "rtb_upload: null source") : RTB_ERR_ARG;
    RTB_BIND(c);
    void* dst = nullptr; size_t cap = 0;
    switch (id) {
        case RTB_BUF_CAMERA:
            if (off > sizeof(CameraRec) || bytes > sizeof(CameraRec) - off) return fail(c, RTB_ERR_CAPACITY, "camera upload past 144 bytes");
            if (!c->cameraSet || std::memcmp(reinterpret_cast<uint8_t*>(&c->camera) + off, src, bytes)) ++c->stamp;   // kernels take the camera by value
            std::memcpy(reinterpret_cast<uint8_t*>(&c->camera) + off, src, bytes); c->cameraSet = true;
            return RTB_OK;
        case RTB_BUF_SCENE_INFO: {
            if (off > sizeof(SceneInfoRec) || bytes > sizeof(SceneInfoRec) - off) return fail(c, RTB_ERR_CAPACITY, "scene info upload past 36 bytes");
            const uint32_t before = c->info.triangleCount;
            if (std::memcmp(reinterpret_cast<uint8_t*>(&c->info) + off, src, bytes)) ++c->stamp;
            std::memcpy(reinterpret_cast<uint8_t*>(&c->info) + off, src, bytes);
            if (c->info.triangleCount != before) c->accelValid = false;
            return RTB_OK;
        }
        case RTB_BUF_SHADOW_PROPS:
            if (off > 4 || bytes > 4 - off) return fail(c, RTB_ERR_CAPACITY, "shadow properties upload past 4 bytes");
            if (std::memcmp(reinterpret_cast<uint8_t*>(&c->shadowSamplesProp) + off, src, bytes)) ++c->stamp;   // (the facade flushes it every frame)
            std::memcpy(reinterpret_cast<uint8_t*>(&c->shadowSamplesProp) + off, src, bytes);
            if (c->width && c->shadowSamplesProp != c->samples) return rtb_resize(c, c->width, c->height, c->shadowSamplesProp);
            return RTB_OK;
        case RTB_BUF_SEED: dst = c->seed.p; cap = sizeof(SeedRec); break;
        case RTB_BUF_TRIANGLES: dst = c->triangles.p; cap = (size_t)c->limits.max_triangles * sizeof(TriangleRec); break;
        case RTB_BUF_SPHERES: dst = c->spheres.p; cap = (size_t)c->limits.max_spheres * 16; break;
        case RTB_BUF_CUBES: dst = c->cubes.p; cap = (size_t)c->limits.max_cubes * 24; break;
        case RTB_BUF_PLANES: dst = c->planes.p; cap = (size_t)c->limits.max_planes * 16; break;
        case RTB_BUF_LIGHTS: dst = c->lights.p; cap = (size_t)c->limits.max_lights * 32; break;
        case RTB_BUF_MATERIALS: dst = c->materials.p; cap = (size_t)c->limits.max_materials * 32; break;
        case RTB_BUF_MATERIAL_INDICES:
            dst = c->materialIndices.p;
            cap = ((size_t)c->limits.max_triangles + c->limits.max_spheres + c->limits.max_cubes + c->limits.max_planes) * 4;
            break;
        default: return fail(c, RTB_ERR_ARG, "rtb_upload: unknown buffer id");
    }
    if (off > cap || bytes > cap - off) return fail(c, RTB_ERR_CAPACITY, "rtb_upload: range exceeds the capacity given to rtb_create");
    if (!bytes) return RTB_OK;
    if (id == RTB_BUF_SEED && c->frontStream && !c->frontNeedsBack) {
        // overlapped frames: K0 of the next frame runs on the front stream, ahead of what the context's stream still holds of the
        // last one; the Seed buffer belongs to that stream (the running frame reads its own snapshot)
        RTB_CUDA(c, cudaMemcpyAsync(static_cast<uint8_t*>(dst) + off, src, bytes, cudaMemcpyHostToDevice, c->frontStream));
        c->frontDirty = true;
        return RTB_OK;
    }
    { const int rc = quiesce(c); if (rc) return rc; }
    if (id == RTB_BUF_TRIANGLES) { std::memcpy(c->triangleMirror.data() + off, src, bytes); c->accelValid = false; ++c->stamp; }
    if (id == RTB_BUF_SPHERES) c->primTree[0].dirty = true;
    if (id == RTB_BUF_CUBES) c->primTree[1].dirty = true;
    if (id == RTB_BUF_LIGHTS && off < sizeof(LightRec)) { ++c->stamp; ++c->light0Epoch; }   // lights[0] picks the occlusion rays' sort key and fills the light cache
    if (id == RTB_BUF_LIGHTS && off < sizeof(LightRec)) std::memcpy(reinterpret_cast<uint8_t*>(&c->light0) + off, src, std::min(bytes, sizeof(LightRec) - off));
    // pageable source: cudaMemcpyAsync stages it before returning, so the caller may reuse src at once (like GPUBuffer::flush)
    RTB_CUDA(c, cudaMemcpyAsync(static_cast<uint8_t*>(dst) + off, src, bytes, cudaMemcpyHostToDevice, c->stream));
    return RTB_OK;
}

int rtb_upload_skybox(rtb_ctx* c, uint32_t w, uint32_t h, const uint16_t* px) {
    if (!c) return RTB_ERR_ARG;
    RTB_BIND(c);
    ++c->stamp;
    { const int rc = quiesce(c); if (rc) return rc; }
    if (!w || !h || !px) { c->skyW = c->skyH = 0; return RTB_OK; }
    RTB_CUDA(c, cudaStreamSynchronize(c->stream));
    RTB_CUDA(c, c->skybox.alloc((size_t)w * h));
    RTB_CUDA(c, cudaMemcpyAsync(c->skybox.p, px, (size_t)w * h * 8, cudaMemcpyHostToDevice, c->stream));
    c->skyW = w; c->skyH = h;
    return RTB_OK;
}

int rtb_build_accel(rtb_ctx* c, rtb_accel_mode mode) {
    if (!c) return RTB_ERR_ARG;
    RTB_BIND(c);
    if (mode != RTB_ACCEL_BRUTE && mode != RTB_ACCEL_BVH && mode != RTB_ACCEL_BVH2) return fail(c, RTB_ERR_ARG, "rtb_build_accel: unknown mode");
    { const int rc = quiesce(c); if (rc) return rc; }
    c->accelMode = mode;
    c->accelValid = false;   // until this build has succeeded
    ++c->stamp;
    c->stats = BvhStats();
    c->nodeCount = 0; c->builtTriangles = 0; c->refits = 0;
    if (mode == RTB_ACCEL_BRUTE) { c->accelValid = true; return RTB_OK; }
    if (c->info.triangleCount > c->limits.max_triangles) return fail(c, RTB_ERR_CAPACITY, "triangleCount exceeds max_triangles");
    std::vector<TravTri> tt;
    const TriangleRec* tris = reinterpret_cast<const TriangleRec*>(c->triangleMirror.data());
    c->builtBy = 0;
    if (mode == RTB_ACCEL_BVH && c->builderOpt >= 1 && c->info.triangleCount >= 2) {
        // ---- on the device: no host copy of the triangles is touched; the uploads are already ordered on the stream -------------
        const auto t0 = std::chrono::steady_clock::now();
        const uint32_t n = c->info.triangleCount, cap = n + 64;   // every node but the root holds at least two children
        RTB_CUDA(c, c->nodes8.alloc(cap)); RTB_CUDA(c, c->travTris.alloc(n));
        RTB_CUDA(c, c->nodeBox.alloc((size_t)cap * 6)); RTB_CUDA(c, c->maxBits.alloc(1)); RTB_CUDA(c, c->areaSums.alloc(2));
        bool tooDeep = false;
        uint32_t nodeCount = 0, leafSlots = 0; float leafExtent = 0.0f;
        const uint32_t maxLevels = (64 - 2) / 3;   // the per-ray traversal stack: 3 entries per level + 2
        const cudaError_t e = device_build_cwbvh(c->triangles.p, n, c->nodes8.p, cap, c->travTris.p, c->nodeBox.p, c->maxBits.p, c->areaSums.p, maxLevels, c->builderOpt == 1 ? 1u : 3u,
                                                 c->stats.levelFirst, nodeCount, leafSlots, leafExtent, &tooDeep, c->stream);
        if (e != cudaSuccess) return cudaFail(c, e, "device_build_cwbvh");
        if (!tooDeep) {
            float root[6]; double sums[2];
            RTB_CUDA(c, cudaMemcpyAsync(root, c->nodeBox.p, sizeof root, cudaMemcpyDeviceToHost, c->stream));
            RTB_CUDA(c, cudaMemcpyAsync(sums, c->areaSums.p, sizeof sums, cudaMemcpyDeviceToHost, c->stream));
            RTB_CUDA(c, cudaStreamSynchronize(c->stream));
            for (int a = 0; a < 3; ++a) { c->stats.lo[a] = root[a]; c->stats.hi[a] = root[3 + a]; }
            const double dx = (double)root[3] - root[0], dy = (double)root[4] - root[1], dz = (double)root[5] - root[2];
            const double rootArea = 2.0 * (dx * dy + dy * dz + dz * dx);
            c->stats.sahCost = rootArea > 0.0 ? (float)((sums[0] + sums[1]) / rootArea) : 0.0f;
            c->stats.nodeCount = nodeCount; c->stats.leafCount = leafSlots; c->stats.maxDepth = (uint32_t)c->stats.levelFirst.size() - 1;
            c->stats.leafNodeExtent = leafExtent;
            c->stats.buildMs = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
            c->nodeCount = nodeCount; c->builtTriangles = n; c->builtBy = 1; c->accelValid = true;
            return RTB_OK;
        }
        c->stats = BvhStats();   // deeper than the traversal stack allows (heavily clustered input): the host builder bounds its depth
    }
    if (mode == RTB_ACCEL_BVH) {
        std::vector<Node8> nodes;
        buildCwbvh(tris, c->info.triangleCount, 0, nodes, tt, c->stats);
        // every level can leave three entries on the per-ray traversal stack (64 entries): the remaining sibling group and,
        // when a triangle group is postponed, that group plus the re-pushed node group (rtb_trace8.cuh, steps A and B)
        if (c->stats.maxDepth * 3 + 2 > 64) return fail(c, RTB_ERR_CAPACITY, "rtb_build_accel: 8-wide tree deeper than the traversal stack allows; use RTB_ACCEL_BVH2");
        RTB_CUDA(c, cudaStreamSynchronize(c->stream));
        RTB_CUDA(c, c->nodes8.alloc(nodes.size()));
        if (!nodes.empty()) RTB_CUDA(c, cudaMemcpy(c->nodes8.p, nodes.data(), nodes.size() * sizeof(Node8), cudaMemcpyHostToDevice));
        c->nodeCount = (uint32_t)nodes.size();
    } else {
        std::vector<BvhNode> nodes;
        buildBvh(tris, c->info.triangleCount, 256, 0, nodes, tt, c->stats);
        RTB_CUDA(c, cudaStreamSynchronize(c->stream));
        RTB_CUDA(c, c->nodes.alloc(nodes.size()));
        if (!nodes.empty()) RTB_CUDA(c, cudaMemcpy(c->nodes.p, nodes.data(), nodes.size() * sizeof(BvhNode), cudaMemcpyHostToDevice));
        c->nodeCount = (uint32_t)nodes.size();
    }
    RTB_CUDA(c, c->travTris.alloc(tt.size()));
    if (!tt.empty()) RTB_CUDA(c, cudaMemcpy(c->travTris.p, tt.data(), tt.size() * sizeof(TravTri), cudaMemcpyHostToDevice));
    c->builtTriangles = c->info.triangleCount;
    c->accelValid = true;
    return RTB_OK;
}

int rtb_refit_accel(rtb_ctx* c) {
    if (!c) return RTB_ERR_ARG;
    RTB_BIND(c);
    if (c->accelMode != RTB_ACCEL_BVH || !c->nodeCount || c->builtTriangles != c->info.triangleCount || c->stats.levelFirst.size() < 2)
        return rtb_build_accel(c, c->accelMode);
    { const int rc = quiesce(c); if (rc) return rc; }
    RTB_CUDA(c, c->nodeBox.alloc((size_t)c->nodeCount * 6)); RTB_CUDA(c, c->maxBits.alloc(1)); RTB_CUDA(c, c->areaSums.alloc(2));
    launch_refit(c->triangles.p, c->info.triangleCount, c->travTris.p, c->info.triangleCount, c->nodes8.p, c->stats.levelFirst.data(),
                 (uint32_t)c->stats.levelFirst.size() - 1, c->nodeBox.p, c->maxBits.p, c->areaSums.p, c->stream);
    RTB_CUDA(c, cudaGetLastError());
    float root[6];   // scene bounds for the packet rule (primaryPackets)
    RTB_CUDA(c, cudaMemcpyAsync(root, c->nodeBox.p, sizeof root, cudaMemcpyDeviceToHost, c->stream));
    double sums[2];
    RTB_CUDA(c, cudaMemcpyAsync(sums, c->areaSums.p, sizeof sums, cudaMemcpyDeviceToHost, c->stream));
    RTB_CUDA(c, cudaStreamSynchronize(c->stream));
    for (int a = 0; a < 3; ++a) { c->stats.lo[a] = root[a]; c->stats.hi[a] = root[3 + a]; }
    {   // the builder's cost: (sum of node-box areas + sum of leaf-slot areas x triangle counts) / area of the scene bounds
        const double dx = (double)root[3] - root[0], dy = (double)root[4] - root[1], dz = (double)root[5] - root[2];
        const double rootArea = 2.0 * (dx * dy + dy * dz + dz * dx);
        if (rootArea > 0.0) c->stats.sahCost = (float)((sums[0] + sums[1]) / rootArea);
    }
    ++c->refits;
    c->accelValid = true;
    ++c->stamp;   // the scene bounds feed the packet rule
    return RTB_OK;
}

int rtb_accel_info_get(const rtb_ctx* c, rtb_accel_info* out) {
    if (!c || !out) return RTB_ERR_ARG;
    out->mode = c->accelMode; out->node_count = c->nodeCount; out->node_bytes = c->accelMode == RTB_ACCEL_BVH ? sizeof(Node8) : sizeof(BvhNode); out->leaf_count = c->stats.leafCount;
    out->max_depth = c->stats.maxDepth; out->tri_record_bytes = sizeof(TravTri); out->sah_cost = c->stats.sahCost; out->build_ms = c->stats.buildMs;
    out->leaf_node_extent = c->stats.leafNodeExtent; out->refits = c->refits; out->primary_packets = (uint32_t)c->lastPrimaryPackets; out->builder = c->builtBy;
    out->sphere_tree_nodes = c->primTree[0].valid ? c->primTree[0].nodeCount : 0u; out->cube_tree_nodes = c->primTree[1].valid ? c->primTree[1].nodeCount : 0u;
    return RTB_OK;
}

int rtb_dispatch(rtb_ctx* c, rtb_pass pass) {
    if (!c) return RTB_ERR_ARG;
    RTB_BIND(c);
    if (pass != RTB_PASS_INIT) { const int rc = checkReady(c); if (rc) return rc; }
    if (pass != RTB_PASS_INIT) { const int rc = ensurePrimTrees(c); if (rc) return rc; }
    if (pass != RTB_PASS_INIT) { const int rc = ensureSunFrame(c); if (rc) return rc; }
    int rc = RTB_OK;
    if (pass != RTB_PASS_FRAME && (rc = quiesce(c))) return rc;
    switch (pass) {
        case RTB_PASS_INIT: if ((rc = waitCopy(c, {RTB_TGT_SEED}))) return rc; launch_init(c->seed.p, c->stream); break;
        case RTB_PASS_RAYGEN: rc = passRaygen(c, false); break;
        case RTB_PASS_SHADOW: rc = passShadow(c, false); break;
        case RTB_PASS_LIGHTING: rc = passShade(c, SHADE_LIGHTING); break;
        case RTB_PASS_COMPOSITE: rc = passShade(c, SHADE_COMPOSITE); break;
        case RTB_PASS_FRAME: {
            if (c->countersOn) RTB_CUDA(c, cudaMemsetAsync(c->counters.p, 0, 2 * sizeof(TraceCounters), c->stream));
            if (useLanes(c)) { if ((rc = prepareLanes(c))) return rc; }
            if (useOverlap(c) && c->lastFrameStamp == c->stamp) { rc = overlappedFrame(c); break; }   // recorded frames, FRONT of this one under BACK of the last
            if ((rc = quiesce(c))) return rc;
            // copies still in flight (rtb_readback_async) are waited for outside the recorded work: before part A those of its
            // targets, before part B those of the shade targets — so a frame's read-back overlaps the next frame's traversal
            if ((rc = waitCopy(c, {RTB_TGT_SEED, RTB_TGT_DIR_T, RTB_TGT_UV_NORMAL, RTB_TGT_SHADOW_BITS}))) return rc;
            const bool canGraph = c->graphOpt && !c->countersOn;
            const bool replay = canGraph && c->graphA && c->graphB && c->graphStamp == c->stamp;
            const bool capture = canGraph && !replay && c->lastFrameStamp == c->stamp;   // the second frame with nothing changed
            c->lastFrameStamp = c->stamp;
            if (capture) {
                if ((rc = captureFrame(c, framePartA, &c->graphA))) return rc;
                if ((rc = captureFrame(c, framePartB, &c->graphB))) return rc;
                c->graphStamp = c->stamp;
            }
            const bool graph = replay || capture;
            if (graph) RTB_CUDA(c, cudaGraphLaunch(c->graphA, c->stream));
            else if ((rc = framePartA(c, !useLanes(c)))) return rc;
            if ((rc = waitCopy(c, {RTB_TGT_LIGHTING, RTB_TGT_ACCUM, RTB_TGT_RGBA8, RTB_TGT_RGBA8_TILED}))) return rc;
            if (graph) RTB_CUDA(c, cudaGraphLaunch(c->graphB, c->stream));
            else if ((rc = framePartB(c, !useLanes(c)))) return rc;
            c->frameTimed = !graph && !useLanes(c);
            break;
        }
        default: return fail(c, RTB_ERR_ARG, "rtb_dispatch: unknown pass");
    }
    if (rc) return rc;
    RTB_CUDA(c, cudaGetLastError());
    return RTB_OK;
}

int rtb_device_ptr(rtb_ctx* c, rtb_target t, void** out, size_t* bytes) {
    if (!c || !out) return RTB_ERR_ARG;
    if (c->frontDirty) { RTB_BIND(c); const int rc = joinFront(c); if (rc) return rc; }   // the caller works on the context's stream
    const size_t px = (size_t)c->width * c->height;
    void* p = nullptr; size_t n = 0;
    switch (t) {
        case RTB_TGT_DIR_T: p = c->dirT.p; n = px * 16; break;
        case RTB_TGT_UV_NORMAL: p = c->uvN.p; n = px * 16; break;
        case RTB_TGT_SHADOW_BITS: p = c->bits.p; n = (size_t)shadowWords(c->width, c->height, shadowLayers(c)) * 4; break;
        case RTB_TGT_LIGHTING: p = c->lighting.p; n = px * 8; break;
        case RTB_TGT_ACCUM: p = c->accum.p; n = px * 16; break;
        case RTB_TGT_RGBA8: p = c->rgba8.p; n = px * 4; break;
        case RTB_TGT_SEED: p = c->seed.p; n = sizeof(SeedRec); break;
        case RTB_TGT_RGBA8_TILED: p = c->rgba8Tiled.p; n = (size_t)((c->fm.blocksX * c->fm.blocksY + c->fm.nranks - 1) / c->fm.nranks) * 4096u; break;
        case RTB_TGT_ACCEL_NODES:
            if (c->accelMode == RTB_ACCEL_BVH) { p = c->nodes8.p; n = (size_t)c->nodeCount * sizeof(Node8); } else { p = c->nodes.p; n = (size_t)c->nodeCount * sizeof(BvhNode); }
            break;
        case RTB_TGT_ACCEL_TRIANGLES: p = c->travTris.p; n = c->nodeCount ? (size_t)c->builtTriangles * sizeof(TravTri) : 0; break;
        default: return fail(c, RTB_ERR_ARG, "unknown target");
    }
    if (t != RTB_TGT_SEED && t != RTB_TGT_ACCEL_NODES && t != RTB_TGT_ACCEL_TRIANGLES && !px) return fail(c, RTB_ERR_STATE, "no frame resources before rtb_resize");
    *out = p; if (bytes) *bytes = n;
    return RTB_OK;
}

int rtb_readback(rtb_ctx* c, rtb_target t, void* dst, size_t bytes) {
    if (!c || !dst) return RTB_ERR_ARG;
    RTB_BIND(c);
    void* p; size_t n;
    const int rc = rtb_device_ptr(c, t, &p, &n);
    if (rc) return rc;
    if (bytes > n) return fail(c, RTB_ERR_ARG, "rtb_readback: more bytes requested than the target holds");
    { const int rc2 = joinFront(c); if (rc2) return rc2; }
    RTB_CUDA(c, cudaMemcpyAsync(dst, p, bytes, cudaMemcpyDeviceToHost, c->stream));
    RTB_CUDA(c, cudaStreamSynchronize(c->stream));
    return RTB_OK;
}

int rtb_readback_async(rtb_ctx* c, rtb_target t, void* dst, size_t bytes) {
    if (!c || !dst) return RTB_ERR_ARG;
    RTB_BIND(c);
    void* p; size_t n;
    int rc = rtb_device_ptr(c, t, &p, &n);
    if (rc) return rc;
    if (bytes > n) return fail(c, RTB_ERR_ARG, "rtb_readback_async: more bytes requested than the target holds");
    if ((rc = drainCopy(c))) return rc;   // one read-back in flight at a time
    if ((rc = joinFront(c))) return rc;
    if (!c->copyStream) {
        RTB_CUDA(c, cudaStreamCreateWithFlags(&c->copyStream, cudaStreamNonBlocking));
        RTB_CUDA(c, cudaEventCreateWithFlags(&c->evReady, cudaEventDisableTiming));
        RTB_CUDA(c, cudaEventCreateWithFlags(&c->evCopied, cudaEventDisableTiming));
    }
    RTB_CUDA(c, cudaEventRecord(c->evReady, c->stream));
    RTB_CUDA(c, cudaStreamWaitEvent(c->copyStream, c->evReady, 0));
    RTB_CUDA(c, cudaMemcpyAsync(dst, p, bytes, cudaMemcpyDeviceToHost, c->copyStream));
    RTB_CUDA(c, cudaEventRecord(c->evCopied, c->copyStream));
    c->copyTarget = (int)t;
    return RTB_OK;
}

int rtb_readback_wait(rtb_ctx* c) {
    if (!c) return RTB_ERR_ARG;
    RTB_BIND(c);
    if (c->copyStream) RTB_CUDA(c, cudaEventSynchronize(c->evCopied));   // the host may read dst; the device-side guard (copyTarget) stays armed until the next writer has waited
    return RTB_OK;
}

int rtb_untile(rtb_ctx* c, const void* tiledAll, uint32_t nranks, uint32_t slotsPerRank, void* out) {
    return rtb_untile_on(c, tiledAll, nranks, slotsPerRank, out, nullptr);
}

int rtb_untile_on(rtb_ctx* c, const void* tiledAll, uint32_t nranks, uint32_t slotsPerRank, void* out, void* cudaStream) {
    if (!c || !tiledAll || !nranks) return c ? fail(c, RTB_ERR_ARG, "rtb_untile: bad argument") : RTB_ERR_ARG;
    RTB_BIND(c);
    if (cudaStream && !out) return fail(c, RTB_ERR_ARG, "rtb_untile_on: a caller-owned stream needs a caller-owned output frame");
    if (!out) { const int rc = waitCopy(c, {RTB_TGT_RGBA8}); if (rc) return rc; }
    if (!c->width) return fail(c, RTB_ERR_STATE, "rtb_untile before rtb_resize");
    FrameMap fm = c->fm;
    fm.nranks = nranks; fm.rank = 0;
    const uint32_t total = fm.blocksX * fm.blocksY;
    if ((uint64_t)((total + nranks - 1) / nranks) * 1024ull > slotsPerRank) return fail(c, RTB_ERR_ARG, "rtb_untile: slots_per_rank too small for this frame");
    launch_untile(fm, static_cast<const uint32_t*>(tiledAll), slotsPerRank, out ? static_cast<uint32_t*>(out) : c->rgba8.p,
                  cudaStream ? static_cast<cudaStream_t>(cudaStream) : c->stream);
    RTB_CUDA(c, cudaGetLastError());
    return RTB_OK;
}

int rtb_present_host(rtb_ctx* c, void* hostFrame, const void* tiledSrc, void* cudaStream) {
    if (!c || !hostFrame) return c ? fail(c, RTB_ERR_ARG, "rtb_present_host: null frame") : RTB_ERR_ARG;
    RTB_BIND(c);
    if (!c->width) return fail(c, RTB_ERR_STATE, "rtb_present_host before rtb_resize");
    if (cudaStream && !tiledSrc) return fail(c, RTB_ERR_ARG, "rtb_present_host: a caller-owned stream needs a caller-owned copy of the tiled pixels");
    void* dev = nullptr;
    RTB_CUDA(c, cudaHostGetDevicePointer(&dev, hostFrame, 0));   // fails unless the frame is page-locked and mapped (cudaHostAlloc / cudaHostRegister)
    const uint32_t* tiled = tiledSrc ? static_cast<const uint32_t*>(tiledSrc) : (c->tileCount > 1 ? c->rgba8Tiled.p : nullptr);
    launch_present_host(c->fm, tiled, c->rgba8.p, static_cast<uint32_t*>(dev), cudaStream ? static_cast<cudaStream_t>(cudaStream) : c->stream);
    RTB_CUDA(c, cudaGetLastError());
    return RTB_OK;
}

int rtb_probe_l2_read_gbs(rtb_ctx* c, size_t bytes, double* outGbs) {
    if (!c || !outGbs || bytes < (1u << 20)) return c ? fail(c, RTB_ERR_ARG, "rtb_probe_l2_read_gbs: bad argument") : RTB_ERR_ARG;
    RTB_BIND(c);
    RTB_CUDA(c, cudaStreamSynchronize(c->stream));
    const double v = measure_l2_read_gbs(bytes, 64, c->stream);
    if (v < 0.0) return cudaFail(c, (cudaError_t)(int)(-v), "measure_l2_read_gbs");
    *outGbs = v;
    return RTB_OK;
}

int rtb_sync(rtb_ctx* c) {
    if (!c) return RTB_ERR_ARG;
    RTB_BIND(c);
    if (c->frontStream) { RTB_CUDA(c, cudaStreamSynchronize(c->frontStream)); RTB_CUDA(c, cudaStreamSynchronize(c->midStream)); }
    RTB_CUDA(c, cudaStreamSynchronize(c->stream));   // (a frame in lanes has joined the lane stream back into this one)
    if (c->copyStream) RTB_CUDA(c, cudaStreamSynchronize(c->copyStream));
    return RTB_OK;
}

int rtb_counters_get(rtb_ctx* c, rtb_counters* out) {
    if (!c || !out) return RTB_ERR_ARG;
    RTB_BIND(c);
    TraceCounters h[2];
    RTB_CUDA(c, cudaMemcpyAsync(h, c->counters.p, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    RTB_CUDA(c, cudaStreamSynchronize(c->stream));
    out->primary_rays = h[0].rays; out->primary_nodes = h[0].nodes; out->primary_tris = h[0].tris; out->primary_hits = h[0].hits;
    out->shadow_rays = h[1].rays; out->shadow_nodes = h[1].nodes; out->shadow_tris = h[1].tris; out->shadow_occluded = h[1].hits;
    return RTB_OK;
}

int rtb_last_frame_ms(rtb_ctx* c, float ms[8]) {
    if (!c || !ms) return RTB_ERR_ARG;
    RTB_BIND(c);
    if (!c->frameTimed) return fail(c, RTB_ERR_STATE, "the last RTB_PASS_FRAME recorded no phase events: none dispatched yet, or it ran as a CUDA graph / in two lanes (set RTB_OPT_FRAME_GRAPH = 0 and RTB_OPT_FRAME_LANES = 1 to time the phases)");
    RTB_CUDA(c, cudaEventSynchronize(c->ev[7]));
    for (int i = 0; i < 7; ++i) RTB_CUDA(c, cudaEventElapsedTime(&ms[i], c->ev[i], c->ev[i + 1]));
    RTB_CUDA(c, cudaEventElapsedTime(&ms[7], c->ev[0], c->ev[7]));
    return RTB_OK;
}

// ---- wavefront path tracing: BASELINE.json configs[3] (definition: csrc/rtb_path.cuh) ------------------------------------------
constexpr uint32_t PATH_MAX_BOUNCES = 15;

int rtb_path_frame(rtb_ctx* c, uint32_t bounces) {
    if (!c) return RTB_ERR_ARG;
    RTB_BIND(c);
    if (bounces > PATH_MAX_BOUNCES) return fail(c, RTB_ERR_ARG, "rtb_path_frame: at most 15 bounces");
    { const int rc = checkReady(c); if (rc) return rc; }
    // consecutive path frames overlap like recorded RTB_PASS_FRAMEs (overlappedFrame): init + depth 0 of frame k+1 on the front stream,
    // on the other set of G-buffer / wavefront / Seed-snapshot buffers, under the deep, latency-bound launches of frame k
    const bool ov = c->overlapOpt && !c->countersOn && !c->releaseBuild && c->fm.localSlots;
    if (!ov) { const int rc = quiesce(c); if (rc) return rc; }
    { const int rc = waitCopy(c, {RTB_TGT_ACCUM, RTB_TGT_RGBA8, RTB_TGT_RGBA8_TILED}); if (rc) return rc; }
    const uint32_t slots = c->fm.localSlots, depths = bounces + 1;
    RTB_CUDA(c, c->pathT.alloc(slots)); RTB_CUDA(c, c->pathL.alloc(slots)); RTB_CUDA(c, c->pathDirect.alloc(slots));
    for (int k = 0; k < 2; ++k) { RTB_CUDA(c, c->pathRays[k].alloc(slots)); RTB_CUDA(c, c->pathSlots[k].alloc(slots)); }
    RTB_CUDA(c, c->pathShadowRays.alloc(slots)); RTB_CUDA(c, c->pathShadowSlots.alloc(slots)); RTB_CUDA(c, c->pathOccA.alloc(slots)); RTB_CUDA(c, c->pathOccB.alloc(slots));
    RTB_CUDA(c, c->pathCounts.alloc(2 * (PATH_MAX_BOUNCES + 2)));
    if (c->pathEv.empty()) {
        c->pathEv.resize(4 * (PATH_MAX_BOUNCES + 1) + 2, nullptr);
        for (auto& ev : c->pathEv) RTB_CUDA(c, cudaEventCreate(&ev));
    }
    cudaEvent_t* ev = c->pathEv.data();
    cudaEvent_t evBegin = ev[4 * (PATH_MAX_BOUNCES + 1)], evEnd = ev[4 * (PATH_MAX_BOUNCES + 1) + 1];
    { const int rc = ensurePrimTrees(c); if (rc) return rc; }   // depth 0 (passRaygen) uses them
    { const int rc = ensureSunFrame(c); if (rc) return rc; }
    SceneView sv = sceneView(c);
    sv.sphereTree = sv.cubeTree = 0u;   // the bounce vertices and their shadow rays keep the linear loops
    TraceCounters* cc = c->countersOn ? c->counters.p : nullptr;
    const PathBuffers pb{c->pathT.p, c->pathL.p, c->pathDirect.p};
    auto closestQ = [&](uint32_t d) { RayQueue q{}; q.rays = c->pathRays[d & 1].p; q.slotIds = c->pathSlots[d & 1].p; q.count = c->pathCounts.p + 2 * d; return q; };       // rays traced at depth d (d >= 1)
    auto shadowQ = [&](uint32_t d) { RayQueue q{}; q.rays = c->pathShadowRays.p; q.slotIds = c->pathShadowSlots.p; q.count = c->pathCounts.p + 2 * d + 1; return q; };     // shadow rays of the vertices at depth d
    uint32_t launches = 0;
    RTB_CUDA(c, cudaEventRecord(evBegin, c->stream));
    RTB_CUDA(c, cudaMemsetAsync(c->pathCounts.p, 0, c->pathCounts.bytes(), c->stream));
    if (c->countersOn) RTB_CUDA(c, cudaMemsetAsync(c->counters.p, 0, 2 * sizeof(TraceCounters), c->stream));
    struct SeedUse { rtb_ctx* c; ~SeedUse() { c->seedUse = nullptr; } } seedGuard{c};   // every return below drops the snapshot again
    cudaStream_t back = c->stream;
    int set = 0;
    if (ov) {
        { const int rc = prepareOverlap(c); if (rc) return rc; }
        swapSets(c);
        set = c->setIndex;
        c->seedUse = c->seedSnap.p + set;
        if (c->frontNeedsBack) {
            RTB_CUDA(c, cudaEventRecord(c->evJoinB, back));
            RTB_CUDA(c, cudaStreamWaitEvent(c->frontStream, c->evJoinB, 0));
            c->frontNeedsBack = false;
        } else if (c->backDoneSet[set])
            RTB_CUDA(c, cudaStreamWaitEvent(c->frontStream, c->evBackDone[set], 0));
        c->stream = c->frontStream;   // init + depth 0 are launched by the pass functions on the context's current stream
    }
    int rcFront = waitCopy(c, {RTB_TGT_SEED});
    if (!rcFront) {
        launch_init(c->seed.p, c->stream, c->seedUse); ++launches;
        // depth 0: the reference's G-buffer, by the launches RTB_PASS_FRAME uses
        if (cudaEventRecord(ev[0], c->stream) != cudaSuccess) rcFront = RTB_ERR_CUDA;
        if (!rcFront) rcFront = passRaygen(c, false);
        if (!rcFront && cudaEventRecord(ev[1], c->stream) != cudaSuccess) rcFront = RTB_ERR_CUDA;
    }
    c->stream = back;
    if (rcFront) return rcFront == RTB_ERR_CUDA ? cudaFail(c, cudaGetLastError(), "rtb_path_frame: depth 0") : rcFront;
    if (ov) {
        RTB_CUDA(c, cudaEventRecord(c->evFrontDone[set], c->frontStream));
        RTB_CUDA(c, cudaStreamWaitEvent(c->stream, c->evFrontDone[set], 0));
        c->frontDirty = false;
    }
    launches += (c->lastPrimaryPackets == PACKETS_FRUSTUM && !c->countersOn && c->fuseOpt) ? 1u : 3u;
    // The occlusion launches of depth d and the nearest-hit launch of depth d + 1 both consume what vertex d produced and touch no
    // common buffer: with RTB_OPT_FRAME_OVERLAP they run side by side on two streams — the deeper launches hold a few thousand
    // rays and are bound by the longest ray's chain of dependent fetches, not by throughput.  The path state is still updated in
    // the order of the definition: shadow terms of depth d, then vertex d + 1.
    const bool sideBySide = c->overlapOpt && !c->countersOn && depths > 1;
    if (sideBySide) { const int rc = prepareLanes(c); if (rc) return rc; }
    launch_path_start(c->fm, sv, &c->camera, seedFor(c), bounces, c->dirT.p, c->uvN.p, pb, shadowQ(0), closestQ(1), c->stream); ++launches;
    for (uint32_t d = 0; d < depths; ++d) {
        const bool more = d + 1 < depths;
        const bool fork = sideBySide && more;
        cudaStream_t so = fork ? c->laneStream : c->stream;
        if (fork) { const int rc = forkLanes(c); if (rc) return rc; }
        const RayQueue sq = shadowQ(d);
        RTB_CUDA(c, cudaEventRecord(ev[4 * d + 2], so));
        launch_occlusion_others(sv, sq.rays, slots, c->pathOccA.p, so, sq.count);
        launch_trace_any_bytes(sv, sq.rays, slots, c->pathOccB.p, fork ? c->lane[1].workCounter.p : c->lane[0].workCounter.p, so, cc ? cc + 1 : nullptr, sq.count);
        RTB_CUDA(c, cudaEventRecord(ev[4 * d + 3], so));
        launches += 2;
        const RayQueue in = more ? closestQ(d + 1) : RayQueue{};
        if (more) {
            RTB_CUDA(c, cudaEventRecord(ev[4 * (d + 1)], c->stream));
            launch_trace_closest(sv, in.rays, slots, c->lane[0].hits.p, c->lane[0].workCounter.p, cc, PACKETS_OFF, c->stream, in.count);
            RTB_CUDA(c, cudaEventRecord(ev[4 * (d + 1) + 1], c->stream));
            ++launches;
        }
        if (fork) { const int rc = joinLanes(c); if (rc) return rc; }
        launch_path_shadow_resolve(sq, slots, c->pathOccA.p, c->pathOccB.p, pb, c->stream); ++launches;
        if (more) { launch_path_vertex(c->fm, sv, &c->camera, seedFor(c), d + 1, bounces, in, c->lane[0].hits.p, pb, shadowQ(d + 1), closestQ(d + 1 + 1), c->stream); ++launches; }
    }
    launch_path_resolve(c->fm, sv, &c->camera, seedFor(c), pb, c->accum.p, c->rgba8.p, c->tileCount > 1 ? c->rgba8Tiled.p : nullptr, c->stream); ++launches;
    RTB_CUDA(c, cudaEventRecord(evEnd, c->stream));
    if (ov) { RTB_CUDA(c, cudaEventRecord(c->evBackDone[set], c->stream)); c->backDoneSet[set] = true; }
    RTB_CUDA(c, cudaGetLastError());
    c->pathDepths = depths; c->pathBounces = bounces; c->pathKernelLaunches = launches; c->pathTimed = true;
    return RTB_OK;
}

int rtb_path_stats_get(rtb_ctx* c, rtb_path_stats* out) {
    if (!c || !out) return RTB_ERR_ARG;
    RTB_BIND(c);
    if (!c->pathTimed) return fail(c, RTB_ERR_STATE, "no rtb_path_frame has been dispatched yet");
    std::memset(out, 0, sizeof *out);
    uint32_t counts[2 * (PATH_MAX_BOUNCES + 2)];
    RTB_CUDA(c, cudaMemcpyAsync(counts, c->pathCounts.p, sizeof counts, cudaMemcpyDeviceToHost, c->stream));
    RTB_CUDA(c, cudaStreamSynchronize(c->stream));
    cudaEvent_t* ev = c->pathEv.data();
    out->depths = c->pathDepths; out->kernel_launches = c->pathKernelLaunches;
    out->closest_launches = c->pathDepths; out->shadow_launches = c->pathDepths;
    for (uint32_t d = 0; d < c->pathDepths; ++d) {
        float a = 0.0f, b = 0.0f;
        RTB_CUDA(c, cudaEventElapsedTime(&a, ev[4 * d], ev[4 * d + 1]));
        RTB_CUDA(c, cudaEventElapsedTime(&b, ev[4 * d + 2], ev[4 * d + 3]));
        out->closest_ms_at_depth[d] = a; out->shadow_ms_at_depth[d] = b;
        out->closest_ms += a; out->shadow_ms += b;
        out->closest_rays_at_depth[d] = d == 0 ? (uint64_t)c->width * c->height : counts[2 * d];
        out->shadow_rays_at_depth[d] = counts[2 * d + 1];
    }
    if (c->tileCount > 1) {   // this rank's camera rays: its pixels
        uint64_t px = 0;
        for (uint32_t k = 0; k < c->fm.localBlocks; ++k) {
            const uint32_t g = k * c->fm.nranks + c->fm.rank, bx = g % c->fm.blocksX, by = g / c->fm.blocksX;
            px += (uint64_t)std::min(32u, c->width - bx * 32u) * std::min(32u, c->height - by * 32u);
        }
        out->closest_rays_at_depth[0] = px;
    }
    for (uint32_t d = 0; d < c->pathDepths; ++d) { out->closest_rays += out->closest_rays_at_depth[d]; out->shadow_rays += out->shadow_rays_at_depth[d]; }
    RTB_CUDA(c, cudaEventElapsedTime(&out->total_ms, ev[4 * (PATH_MAX_BOUNCES + 1)], ev[4 * (PATH_MAX_BOUNCES + 1) + 1]));
    return RTB_OK;
}

// ---- rays-in mode ---------------------------------------------------------------------------------------------
static int stageRays(rtb_ctx* c, const float* rays, uint64_t n, const uint32_t* prev, const float* maxDist) {
    std::vector<RayRec> h((size_t)n);
    for (size_t i = 0; i < (size_t)n; ++i) {
        RayRec& r = h[i];
        r.ox = rays[6 * i]; r.oy = rays[6 * i + 1]; r.oz = rays[6 * i + 2]; r.prev = prev ? prev[i] : NO_RAY_HIT;
        r.dx = rays[6 * i + 3]; r.dy = rays[6 * i + 4]; r.dz = rays[6 * i + 5]; r.tmax = maxDist ? maxDist[i] : NO_HIT;
    }
    RTB_CUDA(c, c->rinRays.alloc((size_t)n));
    RTB_CUDA(c, cudaMemcpyAsync(c->rinRays.p, h.data(), (size_t)n * sizeof(RayRec), cudaMemcpyHostToDevice, c->stream));
    RTB_CUDA(c, cudaStreamSynchronize(c->stream));
    return RTB_OK;
}

int rtb_trace_rays(rtb_ctx* c, const float* rays, uint64_t n, const uint32_t* prev, uint32_t* object, float* t, float* uv) {
    if (!c || (!rays && n)) return RTB_ERR_ARG;
    RTB_BIND(c);
    if (n > 0x7FFFFFFFull) return fail(c, RTB_ERR_ARG, "rtb_trace_rays: too many rays");
    { const int rc = checkScene(c); if (rc) return rc; }
    if (!n) return RTB_OK;
    { const int rc = quiesce(c); if (rc) return rc; }
    int rc = stageRays(c, rays, n, prev, nullptr);
    if (rc) return rc;
    RTB_CUDA(c, c->rinHits.alloc((size_t)n)); RTB_CUDA(c, c->rinObj.alloc((size_t)n)); RTB_CUDA(c, c->rinT.alloc((size_t)n)); RTB_CUDA(c, c->rinUv.alloc((size_t)n));
    { const int rc2 = ensurePrimTrees(c); if (rc2) return rc2; }
    const SceneView sv = sceneView(c);
    launch_trace_closest(sv, c->rinRays.p, (uint32_t)n, c->rinHits.p, c->lane[0].workCounter.p, nullptr, c->packetsOpt == 2 ? PACKETS_OFF : (int)c->packetsOpt, c->stream);
    const PrimHit *ws = nullptr, *wc = nullptr;
    if (sv.sphereTree) {
        RTB_CUDA(c, c->rinPrimS.alloc((size_t)n));
        launch_prims_closest(0, sv, primTreeOf(c, 0), c->rinRays.p, (uint32_t)n, nullptr, c->rinHits.p, nullptr, c->rinPrimS.p, c->stream);
        ws = c->rinPrimS.p;
    }
    if (sv.cubeTree) {
        RTB_CUDA(c, c->rinPrimC.alloc((size_t)n));
        launch_prims_closest(1, sv, primTreeOf(c, 1), c->rinRays.p, (uint32_t)n, nullptr, c->rinHits.p, ws, c->rinPrimC.p, c->stream);
        wc = c->rinPrimC.p;
    }
    launch_finish_rays(sv, c->rinRays.p, c->rinHits.p, (uint32_t)n, c->rinObj.p, c->rinT.p, c->rinUv.p, c->stream, ws, wc);
    RTB_CUDA(c, cudaGetLastError());
    if (object) RTB_CUDA(c, cudaMemcpyAsync(object, c->rinObj.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
    if (t) RTB_CUDA(c, cudaMemcpyAsync(t, c->rinT.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
    if (uv) RTB_CUDA(c, cudaMemcpyAsync(uv, c->rinUv.p, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
    RTB_CUDA(c, cudaStreamSynchronize(c->stream));
    return RTB_OK;
}

int rtb_occlusion_rays(rtb_ctx* c, const float* rays, uint64_t n, const float* maxDist, const uint32_t* prev, uint8_t* occluded) {
    if (!c || (!rays && n) || (!occluded && n)) return RTB_ERR_ARG;
    RTB_BIND(c);
    if (n > 0x7FFFFFFFull) return fail(c, RTB_ERR_ARG, "rtb_occlusion_rays: too many rays");
    { const int rc = checkScene(c); if (rc) return rc; }
    if (!n) return RTB_OK;
    { const int rc = quiesce(c); if (rc) return rc; }
    int rc = stageRays(c, rays, n, prev, maxDist);
    if (rc) return rc;
    RTB_CUDA(c, c->rinOcc.alloc((size_t)n)); RTB_CUDA(c, c->rinOcc2.alloc((size_t)n));
    { const int rc2 = ensurePrimTrees(c); if (rc2) return rc2; }
    const SceneView sv = sceneView(c);
    launch_occlusion_others(sv, c->rinRays.p, (uint32_t)n, c->rinOcc.p, c->stream);
    launch_trace_any_bytes(sv, c->rinRays.p, (uint32_t)n, c->rinOcc2.p, c->lane[0].workCounter.p, c->stream);
    if (sv.sphereTree) launch_prims_any(0, sv, primTreeOf(c, 0), c->rinRays.p, (uint32_t)n, nullptr, c->rinOcc.p, nullptr, nullptr, FrameMap{}, c->stream);
    if (sv.cubeTree) launch_prims_any(1, sv, primTreeOf(c, 1), c->rinRays.p, (uint32_t)n, nullptr, c->rinOcc.p, nullptr, nullptr, FrameMap{}, c->stream);
    RTB_CUDA(c, cudaGetLastError());
    std::vector<uint8_t> a((size_t)n), b((size_t)n);
    RTB_CUDA(c, cudaMemcpyAsync(a.data(), c->rinOcc.p, (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    RTB_CUDA(c, cudaMemcpyAsync(b.data(), c->rinOcc2.p, (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    RTB_CUDA(c, cudaStreamSynchronize(c->stream));
    for (size_t i = 0; i < (size_t)n; ++i) occluded[i] = (a[i] | b[i]) ? 1 : 0;
    return RTB_OK;
}

}  // extern "C"
