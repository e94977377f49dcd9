// rtb_kernels.cuh — launch interface between the C-ABI layer (rtb_api.cu) and the kernels (rtb_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <vector>
#include "rtb_types.h"

namespace rtb {

// Which pixels this context renders and where each wavefront slot lives.
// The screen is cut into 32x32-pixel blocks; block g belongs to rank g % nranks.  Local slot i:
//   k = i >> 10 (local block, global block g = k * nranks + rank), s = (i >> 5) & 31 (8x4 sub-tile),
//   lane = i & 31 -> pixel (bx*32 + (s&3)*8 + (lane&7), by*32 + (s>>2)*4 + (lane>>3)).
// One warp of consecutive slots is therefore an 8x4 pixel patch: coherent primary rays and full 32-byte
// sectors on every scan-line-order store.
struct FrameMap {
    uint32_t w, h, blocksX, blocksY, rank, nranks, localBlocks, localSlots;
    // where local slot i lives in the rank's RTB_TGT_RGBA8_TILED buffer: block (i >> 10) * tiledMul + tiledAdd.  (1, 0) for a rank's
    // own map; (2, h) for half-frame lane h, which is virtual rank h * n + r of 2 n and owns the rank's local blocks 2 k + h.
    uint32_t tiledMul, tiledAdd;
};

// what getSunDirection (SH/rand_util.glsl:66-85) and getDirToLight (SH/light.glsl:127-130) derive from a directional light alone:
// computed once per upload of lights[0] by k_sun_frame with the shaders' own expressions, read by every pixel afterwards
struct SunFrame { float h; float dir[3], bitangent[3], tangent[3]; };

struct SceneView {
    const SunFrame* sun0;         // of lights[0] when it is directional and the frame is current; null: evaluated per pixel
    const TriangleRec* triangles; const float4* spheres; const float* cubes; const float4* planes;
    const LightRec* lights; const MaterialRec* materials; const uint32_t* materialIndices;
    const uint2* skybox; uint32_t skyW, skyH;
    SceneInfoRec info;            // host mirror of the reference's SceneData UBO, passed by value
    const BvhNode* nodes; const Node8* nodes8; const TravTri* travTris; uint32_t nodeCount;   // of the structure in use
    uint32_t useBvh;              // ACCEL_KIND_*
    uint32_t releaseBuild;        // RTB_OPT_SHADER_BUILD: 0 = the reference's DEBUG shader build (what ships), 1 = RELEASE
    // spheres / cubes through their own trees (rtb_trace8s.cuh): when set, the linear loops of that type are skipped by
    // finishGeometry (given the traversal's winner) and by occludedByOthers (the traversal answers)
    uint32_t sphereTree, cubeTree;
};
enum : uint32_t { ACCEL_KIND_BRUTE = 0, ACCEL_KIND_CWBVH = 1, ACCEL_KIND_BVH2 = 2 };

struct TraceCounters { unsigned long long rays, nodes, tris, hits; };
// spheres / cubes through their own trees (rtb_trace8s.cuh): the winner of a type's traversal (id within the type, NO_RAY_HIT = none)
struct PrimHit { float t; uint32_t id; };
struct PrimTree { const Node8* nodes; const TravTri* tt; uint32_t nodeCount; };

// Light-space 2D coordinate of an occlusion ray (rtb_sort.cu): rays with equal coordinates travel along the same line.
struct RayBin {
    uint32_t kind;         // 0: no binning, 1: directional light (u, v = origin . b1, origin . b2), 2: point light (octahedral map of origin - lpos)
    uint32_t bits;         // cells per axis = 1 << bits (<= 10); cell index = Morton code of the two cell coordinates
    float b1[3], b2[3], lpos[3];
    float u0, v0, su, sv;  // cell coordinate = clamp((u - u0) * su, 0, cells - 1)
};
// Queue of live rays appended by k_shadowgen (and the path kernels): count[0] rays, each with the wavefront slot it belongs to.
struct RayQueue {
    RayRec* rays; uint32_t* slotIds; uint32_t* count;
    uint32_t* cell; uint32_t* rank; uint32_t* hist; uint32_t* blockSums;   // binning (RayBin.kind != 0)
};

void launch_init(SeedRec* seed, cudaStream_t s, SeedRec* snap = nullptr);
void launch_raygen(const FrameMap& fm, const CameraRec* cam, const SeedRec* seed, RayRec* rays, cudaStream_t s);
// nearest triangle for every ray slot (BVH or brute force according to sv.useBvh).  packets != PACKETS_OFF: 32 consecutive
// slots are a coherent patch and the 8-wide tree is walked once per warp — every lane testing every child box
// (PACKETS_UNION, rtb_trace8p.cuh) or one lane testing one child box against one quadrant's interval ray
// (PACKETS_FRUSTUM, rtb_trace8f.cuh); otherwise one traversal per lane.
enum { PACKETS_OFF = 0, PACKETS_UNION = 1, PACKETS_FRUSTUM = 3 };
void launch_trace_closest(const SceneView& sv, const RayRec* rays, uint32_t n, TriHit* hits, uint32_t* workCounter,
                          TraceCounters* counters, int packets, cudaStream_t s, const uint32_t* countPtr = nullptr);   // countPtr: the rays are a queue of *countPtr (device) entries, n its capacity
// raygen + nearest hit (frustum packets) + G-buffer finish as ONE launch: no ray / hit records in between
void launch_primary_fused(const FrameMap& fm, const SceneView& sv, const CameraRec* cam, const SeedRec* seed, float4* dirT, float4* uvN,
                          uint32_t* workCounter, cudaStream_t s);
// spheres, cubes, planes after the triangles, normal interpolation, G-buffer stores (raygen.comp:39-51)
void launch_finish_primary(const FrameMap& fm, const SceneView& sv, const RayRec* rays, const TriHit* hits,
                           float4* dirT, float4* uvN, cudaStream_t s, const PrimHit* sph = nullptr, const PrimHit* cub = nullptr);
// instrumented frames: counters->hits = pixels of this rank whose nearest hit is any primitive
void launch_count_hits(const FrameMap& fm, const float4* dirT, TraceCounters* counters, cudaStream_t s);
// shadow.comp ray set-up + occlusion by the non-triangle primitives; leaves triangle work in `rays`
// queue == nullptr: one record per (sample, slot), dead slots marked tmax < 0 (slot order).  Otherwise the live rays are
// appended to the queue and, with bin.kind != 0, binned for launch_sort_rays.
void launch_shadowgen(const FrameMap& fm, const SceneView& sv, const CameraRec* cam, const SeedRec* seed, uint32_t samples,
                      const float4* dirT, RayRec* rays, uint32_t* bits, const RayQueue* queue, const RayBin* bin, cudaStream_t s);
// counting sort of a binned queue into (outRays, outSlots); cells = 4^bin.bits, maxRays bounds the queue length
void launch_sort_rays(const RayQueue& q, uint32_t cells, uint32_t maxRays, RayRec* outRays, uint32_t* outSlots, cudaStream_t s);
// any-hit over the triangles; sets the (pixel, sample) bit of every occluded slot.  slotIds / countPtr (8-wide tree only):
// the rays are a queue — ray r belongs to wavefront slot slotIds[r], and *countPtr (device) rays are valid (n = upper bound)
void launch_trace_any_bits(const FrameMap& fm, const SceneView& sv, const RayRec* rays, uint32_t n, uint32_t* bits,
                           uint32_t* workCounter, TraceCounters* counters, const uint32_t* slotIds, const uint32_t* countPtr, cudaStream_t s);
// the same by beam packets over a queue sorted in light space; `fallback` (rays, slotIds, count; capacity n) receives what the beam
// walk does not take and is answered by the per-ray kernel in the same call
void launch_trace_beam_bits(const FrameMap& fm, const SceneView& sv, const RayRec* rays, uint32_t n, uint32_t* bits, uint32_t* workCounter,
                            TraceCounters* counters, const uint32_t* slotIds, const uint32_t* countPtr, const RayQueue& fallback, cudaStream_t s);
// RELEASE shader build only: zero the shadow words of the 16x2 strips that hold at least one hit pixel (the reference's
// subgroups without hits leave early and store nothing: nv_all.shadow.comp:69-82); the DEBUG build zeroes every word
void launch_clear_hit_strips(const FrameMap& fm, const float4* dirT, uint32_t samples, uint32_t* bits, cudaStream_t s);
// any-hit, one byte per ray (rays-in mode)
void launch_trace_any_bytes(const SceneView& sv, const RayRec* rays, uint32_t n, uint8_t* occluded, uint32_t* workCounter,
                            cudaStream_t s, TraceCounters* counters = nullptr, const uint32_t* countPtr = nullptr);
// ---- beyond the reference: every light evaluated, screen-tile light lists, temporal History blend ----
// The reference samples lights[0] only and scales by lightCount (nv_all.shadow.comp:97, nv_all.lighting.comp:88,98).  With
// LightsView.mode != 0 every light gets its own shadow ray per sample and its own Cook-Torrance term; the shadow mask then has
// lightCount * samples layers, layer = light * samples + sample.  mode 2 walks, per 16x16-pixel screen tile, a list of the lights
// that can reach a hit point of the tile (at most LIGHTS_PER_TILE = 32 entries, ascending; a tile with more keeps the full loop):
// a light left out contributes exactly zero to every pixel of the tile, so the result is mode 1's bit for bit.
constexpr uint32_t LIGHTS_PER_TILE = 32;          // ref: res/shaders/defines.glsl:6
constexpr uint32_t LIGHT_TILE_ALL = 0xFFFFFFFFu;  // tile count value: walk every light
struct LightsView {
    uint32_t mode;                 // 0 reference (light 0 x lightCount), 1 all lights, 2 all lights through the tile lists
    uint32_t tilesX;               // 16x16-pixel tiles per row
    const uint32_t* tileCount;     // per tile: entries in its list, or LIGHT_TILE_ALL
    const uint32_t* tileList;      // per tile: LIGHTS_PER_TILE light indices, ascending
    uint32_t lightBegin, lightEnd; // shadow-ray generation only: the lights of this launch (the queue is filled in chunks of lights)
    float historyAlpha;            // > 0: lighting = history * (1 - a) + lighting * a through `history` (rgba16f), stored in both
    uint2* history;
    // lighting.comp derives its random pair from the pixel, the sample index and the sample count alone (its Seed block is unbound
    // and reads as zero: decree D8), so the pair — and, for a directional light 0, the light direction it yields — is the same in
    // every frame of a given size: kept per (sample, pixel) instead of four exact-reduction binary64 sines per pixel and frame.
    //   cacheKind 0: none, computed in the kernel; 1: xy = random; 2: xyz = normalised direction to (directional) light 0
    uint32_t cacheKind;
    const float4* lightCache;      // [sample * w * h + y * w + x]
};
void launch_sun_frame(const LightRec* lights, SunFrame* out, cudaStream_t s);
void launch_light_cache(const FrameMap& fm, const SceneView& sv, uint32_t samples, uint32_t kind, float4* cache, cudaStream_t s);
void launch_light_tiles(const FrameMap& fm, const SceneView& sv, const CameraRec* cam, const float4* dirT, uint32_t* tileCount, uint32_t* tileList, cudaStream_t s);
// shadow rays of every light in [lv.lightBegin, lv.lightEnd) for every hit pixel and sample, appended to `queue`
// (slot id = layer * localSlots + slot); occluders other than triangles are answered here
void launch_shadowgen_lights(const FrameMap& fm, const SceneView& sv, const CameraRec* cam, const SeedRec* seed, uint32_t samples, const float4* dirT,
                             uint32_t* bits, const RayQueue& queue, const LightsView& lv, cudaStream_t s);

// lighting.comp + composite.comp
enum { SHADE_LIGHTING = 1, SHADE_COMPOSITE = 2, SHADE_BOTH = 3 };
void launch_shade(int what, const FrameMap& fm, const SceneView& sv, const CameraRec* cam, const SeedRec* seed, uint32_t samples,
                  const float4* dirT, const float4* uvN, const uint32_t* bits, uint2* lighting, float4* accum,
                  uint32_t* rgba8, uint32_t* rgba8Tiled, cudaStream_t s, const LightsView* lights = nullptr);
// spheres / cubes through their own trees (rtb_trace8s.cuh)
// nearest sphere / cube for every ray given what the earlier stages left (triHits, `before`); kind 0 spheres, 1 cubes
void launch_prims_closest(int kind, const SceneView& sv, const PrimTree& tree, const RayRec* rays, uint32_t n, const uint32_t* countPtr,
                          const TriHit* triHits, const PrimHit* before, PrimHit* out, cudaStream_t s);
// any hit: ORs into `bytes` (rays-in) or into the shadow words (slotIds as launch_trace_any_bits)
void launch_prims_any(int kind, const SceneView& sv, const PrimTree& tree, const RayRec* rays, uint32_t n, const uint32_t* countPtr, uint8_t* bytes,
                      uint32_t* bits, const uint32_t* slotIds, const FrameMap& fm, cudaStream_t s);
void launch_proxy_triangles(int kind, const float4* spheres, const float* cubes, uint32_t n, TriangleRec* out, cudaStream_t s);
// rays-in helpers (sph / cub: the winners of the type's tree traversal, or nullptr for the linear loops)
void launch_finish_rays(const SceneView& sv, const RayRec* rays, const TriHit* hits, uint32_t n, uint32_t* object, float* t,
                        float2* uv, cudaStream_t s, const PrimHit* sph = nullptr, const PrimHit* cub = nullptr);
void launch_occlusion_others(const SceneView& sv, RayRec* rays, uint32_t n, uint8_t* occluded, cudaStream_t s, const uint32_t* countPtr = nullptr);
// rank 0: gathered [nranks][slotsPerRank] tiled pixels -> scan-line rgba8
void launch_untile(const FrameMap& fm, const uint32_t* tiledAll, uint32_t slotsPerRank, uint32_t* rgba8, cudaStream_t s);

// ---- wavefront path tracing (rtb_path.cuh): the diffuse-bounce workload of BASELINE.json configs[3] ----
struct PathBuffers { float4* throughput; float4* radiance; float4* direct; };
void launch_path_start(const FrameMap& fm, const SceneView& sv, const CameraRec* cam, const SeedRec* seed, uint32_t bounces, const float4* dirT, const float4* uvN,
                       const PathBuffers& pb, const RayQueue& shadowQ, const RayQueue& nextQ, cudaStream_t s);
void launch_path_vertex(const FrameMap& fm, const SceneView& sv, const CameraRec* cam, const SeedRec* seed, uint32_t depth, uint32_t bounces, const RayQueue& inQ,
                        const TriHit* hits, const PathBuffers& pb, const RayQueue& shadowQ, const RayQueue& nextQ, cudaStream_t s);
void launch_path_shadow_resolve(const RayQueue& shadowQ, uint32_t maxRays, const uint8_t* occOthers, const uint8_t* occTris, const PathBuffers& pb, cudaStream_t s);
void launch_path_resolve(const FrameMap& fm, const SceneView& sv, const CameraRec* cam, const SeedRec* seed, const PathBuffers& pb, float4* accum, uint32_t* rgba8,
                         uint32_t* rgba8Tiled, cudaStream_t s);

// this rank's pixels (from the tiled target when tiled != nullptr, else from the scan-line rgba8 target) into a mapped host frame
void launch_present_host(const FrameMap& fm, const uint32_t* tiled, const uint32_t* rgba8, uint32_t* hostFrame, cudaStream_t s);

// refit of the 8-wide tree from the current triangle buffer (rtb_refit.cu); levelFirst is a HOST array of levels + 1 entries;
// areaSums (2 doubles, device) receives the two sums of the SAH cost: node-box areas, leaf-slot areas x triangle counts
void launch_refit(const TriangleRec* tris, uint32_t triCount, TravTri* tt, uint32_t ttCount, Node8* nodes, const uint32_t* levelFirst,
                  uint32_t levels, float* nodeBox, uint32_t* maxBits, double* areaSums, cudaStream_t s);

// rtb_build.cu: the 8-wide tree built on the device (Morton sort, radix tree, collapse, then launch_refit for every box).
// levelFirst comes back as the host array launch_refit wants; *tooDeep: more than maxLevels levels, nothing usable was built.
cudaError_t device_build_cwbvh(const TriangleRec* tris, uint32_t n, Node8* nodes8, uint32_t nodeCapacity, TravTri* tt, float* nodeBox, uint32_t* maxBits,
                               double* areaSums, uint32_t maxLevels, uint32_t leafMax, std::vector<uint32_t>& levelFirst, uint32_t& nodeCount, uint32_t& leafSlots,
                               float& leafNodeExtent, bool* tooDeep, cudaStream_t s);

// rtb_probe.cu: L2 read bandwidth (GB/s) over a buffer of `bytes` read `passes` times; negative = -cudaError_t
double measure_l2_read_gbs(size_t bytes, uint32_t passes, cudaStream_t s);

int trace_grid_blocks();   // persistent grid size used by the traversal kernels (for reporting)

}  // namespace rtb
