// rtb_types.h — byte layouts shared by host and device code.
//
// The scene-side records keep the reference's GPU layouts byte for byte:
//   Camera 144 B   ref: igx/include/types/scene_object_types.hpp:32-66, res/shaders/camera.glsl:14-42
//   Seed 24 B      ref: include/rt/structs.hpp:8-14, res/shaders/rand_util.glsl:6-10
//   SceneInfo 36 B ref: igx/include/helpers/scene_graph.hpp:61-74, res/shaders/scene.glsl:6-20
//   Triangle 48 B  ref: scene_object_types.hpp:74-112, res/shaders/primitive.glsl:28-38
//   Light 32 B     ref: scene_object_types.hpp:158-265, primitive.glsl:46-54
//   Material 32 B  ref: scene_object_types.hpp:267-290, primitive.glsl:72-81
// The BVH node and traversal-triangle records are new (the reference has no acceleration structure).
#pragma once
#include <stdint.h>

namespace rtb {

struct CameraRec {
    float eye[3]; uint32_t width;
    float p0[3]; uint32_t height;
    float p1[3]; float ipd;
    float p2[3]; uint32_t projectionType;
    float skyboxColor[3]; float exposure;
    float p3[3]; float focalDistance;
    float p4[3]; float aperature;
    float p5[3]; uint32_t flags;
    float invRes[2]; uint32_t tiles[2];
};
static_assert(sizeof(CameraRec) == 144, "Camera is 144 bytes");

struct SeedRec { float randomX, randomY, cpuOffsetX, cpuOffsetY; uint32_t sampleCount, sampleOffset; };
static_assert(sizeof(SeedRec) == 24, "Seed is 24 bytes");

struct SceneInfoRec {
    uint32_t lightCount, materialCount, triangleCount, sphereCount, cubeCount, planeCount;
    uint32_t directionalLightCount, spotLightCount, pointLightCount;
};
static_assert(sizeof(SceneInfoRec) == 36, "SceneInfo is 36 bytes");

struct TriangleRec { float p0[3]; uint32_t n0; float p1[3]; uint32_t n1; float p2[3]; uint32_t n2; };
static_assert(sizeof(TriangleRec) == 48, "Triangle is 48 bytes");

struct LightRec { float pos[3]; uint32_t radOrigin; uint32_t dir[2]; uint32_t colorRG; uint32_t colorBType; };
static_assert(sizeof(LightRec) == 32, "Light is 32 bytes");

struct MaterialRec { uint32_t albedoMetallic[2]; uint32_t ambientRoughness[2]; uint32_t emissive[2]; float transparency; uint32_t materialInfo; };
static_assert(sizeof(MaterialRec) == 32, "Material is 32 bytes");

enum : uint32_t { CAMERA_USE_UI = 1u, CAMERA_USE_SUPERSAMPLING = 2u };
enum : uint32_t { LIGHT_DIRECTIONAL = 0u, LIGHT_SPOT = 1u, LIGHT_POINT = 2u };

// ---- acceleration structure (new) ------------------------------------------------------------------
//
// Binary BVH, one 64-byte record per inner node holding BOTH children's boxes, so one node fetch
// (4 x 16-byte vector loads, 64-byte aligned = two 32-byte sectors of one 128-byte line... ) decides
// both children.  Child links: >= 0 inner-node index; < 0 leaf: ~link = (firstTri << 3) | (count - 1).
struct alignas(64) BvhNode {
    float c0lox, c0hix, c0loy, c0hiy;   // child 0: x and y slabs
    float c1lox, c1hix, c1loy, c1hiy;   // child 1: x and y slabs
    float c0loz, c0hiz, c1loz, c1hiz;   // both children: z slabs
    int32_t child0, child1; uint32_t pad0, pad1;
};
static_assert(sizeof(BvhNode) == 64, "BVH node is 64 bytes");

// 8-wide compressed node, 128 bytes = one cache line = eight 16-byte vector loads.  The topology fields follow
// Ylitie et al. 2017; the box encoding is ours:
//   p, e        origin and per-axis exponent (biased by 127) of the node's grid: world = p + g * 2^e
//   imask       bit s set: child slot s is an inner node
//   childBase   index of the first inner child (inner children are consecutive, in slot order)
//   triBase     index of the first triangle of the leaf children (consecutive, in slot order, at most 24)
//   valid       imask << 24 | triangle presence: leaf slot s holding c <= 3 triangles sets bits 3s .. 3s+c-1; the
//               triangle at bit b is triBase + popcount(presence bits below b)
//   planes      child box planes as bf16 grid coordinates g in [0, 256), two children per word: slot 2k in the
//               upper half, slot 2k+1 in the lower half.  The traversal reads the upper value by taking the whole
//               word as a float (no decode), which can only enlarge it by less than one bf16 step, so the builder
//               rounds upper-half lo planes down by one extra step; every stored box contains the true box.
//               planes[axis][0] = lo, planes[axis][1] = hi: one 32-byte granule per axis, fetched with ONE 256-bit
//               load whose two destination halves are swapped by the ray's direction sign (near / far).
// The record is four 32-byte granules = four LDG.256 per visit (sm_100 has 256-bit global loads).
struct alignas(128) Node8 {
    float p[3]; uint8_t e[3]; uint8_t imask;
    uint32_t childBase, triBase, valid, reserved;
    uint32_t planes[3][2][4];
    uint32_t* lo(int a) { return planes[a][0]; }
    uint32_t* hi(int a) { return planes[a][1]; }
    const uint32_t* lo(int a) const { return planes[a][0]; }
    const uint32_t* hi(int a) const { return planes[a][1]; }
};
static_assert(sizeof(Node8) == 128, "CWBVH node is 128 bytes");

// Traversal triangle, in leaf order: p0, e1 = p1 - p0, e2 = p2 - p0 (the same float subtractions the
// reference shader performs per test, done once), and the triangle's index in the uploaded array.
struct alignas(16) TravTri {
    float p0[3]; uint32_t id;
    float e1[3]; uint32_t pad1;
    float e2[3]; uint32_t pad2;
};
static_assert(sizeof(TravTri) == 48, "traversal triangle is 48 bytes");

constexpr float NO_HIT = 3.4028235e38f;          // ref: res/shaders/primitive.glsl:6
constexpr uint32_t NO_RAY_HIT = 0xFFFFFFFFu;     // ref: res/shaders/primitive.glsl:7

// One ray of the wavefront: 32 bytes, two 16-byte vector accesses.
//   o.xyz origin, o.w = bits(prev object id to exclude); d.xyz direction, d.w = tmax (< 0: inactive slot)
struct alignas(16) RayRec { float ox, oy, oz; uint32_t prev; float dx, dy, dz, tmax; };
static_assert(sizeof(RayRec) == 32, "ray is 32 bytes");

// Nearest-triangle result of the BVH search: 16 bytes.
struct alignas(16) TriHit { float t; uint32_t id; float u, v; };

}  // namespace rtb
