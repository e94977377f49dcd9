// ref_tu.cpp — one translation unit per reference shader: the adapted shader text is included INSIDE a namespace (so the five
// main()s and their globals stay apart), followed by the harness that binds the host arrays and walks the dispatch grid.
// Compiled five times by the Makefile with -DREF_KIND=<n> -DREF_FILE="gen/<shader>" and -DDEBUG or -DRELEASE, -DVENDOR_NV
// (what res/shaders/compile.sh passes).  TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).  Ours; contains no reference code.
#include "glsl_shim.h"
#include "ref_api.h"
#include "ref_runtime.h"

#define REF_INIT 0
#define REF_RAYGEN 1
#define REF_SHADOW 2
#define REF_LIGHTING 3
#define REF_COMPOSITE 4

namespace glsl {
namespace REF_NS {

#define main shader_main
#include REF_FILE
#undef main

// ---- harness -------------------------------------------------------------------------------------------------------------------
#if REF_KIND != REF_INIT
static_assert(sizeof(Triangle) == 48 && sizeof(Cube) == 24 && sizeof(Light) == 32 && sizeof(Material) == 32, "std430 record sizes");
static_assert(sizeof(Camera) == 144 && sizeof(SceneInfo) == 36 && sizeof(Seed) == 24, "UBO sizes");

static void bindScene(const ref_bind* b) {
    memcpy(&camera, b->camera144, 144);
    memcpy(&sceneInfo, b->scene_info9, 36);
    triangles = (Triangle*)b->triangles; spheres = (vec4*)b->spheres; cubes = (Cube*)b->cubes; planes = (vec4*)b->planes;
    lights = (Light*)b->lights; materials = (Material*)b->materials; materialIndices = (uint*)b->material_indices;
    skybox = sampler2D{b->skybox, b->skybox ? (int)b->sky_w : 0, b->skybox ? (int)b->sky_h : 0, FMT_RGBA16F, 1};
}

// Dispatch(w, h) with local size 16x16x1: every invocation of every group, the out-of-frame ones included
static void dispatch2D(uint w, uint h) {
    const uint gx = (w + 15) / 16, gy = (h + 15) / 16;
    parallel_for((int64_t)gy * 16, [&](int64_t y) {
        for (uint x = 0; x < gx * 16; ++x) { gl_GlobalInvocationID = uvec3(x, (uint)y, 0u); shader_main(); }
    });
}
#endif

}  // namespace REF_NS
}  // namespace glsl

using namespace glsl;
using namespace glsl::REF_NS;

extern "C" {

#if REF_KIND == REF_INIT
void ref_init(const ref_bind* b) {
    memcpy(&seed, b->seed24, 24);
    gl_GlobalInvocationID = uvec3(0u, 0u, 0u);
    shader_main();
    memcpy(b->seed24, &seed, 24);
}
#ifdef DEBUG
int ref_is_debug(void) { return 1; }
#else
int ref_is_debug(void) { return 0; }
#endif
#endif

#if REF_KIND == REF_RAYGEN
void ref_raygen(const ref_bind* b) {
    bindScene(b);
    memcpy(&seed, b->seed24, 24);
    dirObject = image2D{b->dirT, (int)b->width, (int)b->height, FMT_RGBA32F};
    uvNormal = image2D{b->uvN, (int)b->width, (int)b->height, FMT_RGBA32F};
    dispatch2D(b->width, b->height);
}

void ref_trace_rays(const ref_bind* b, const float* rays, uint64_t n, const uint32_t* prev, uint32_t* object, float* t,
                    float* uv, uint32_t* enc_normal2) {
    bindScene(b);
    parallel_for((int64_t)((n + 63) / 64), [&](int64_t c) {
        for (uint64_t i = (uint64_t)c * 64; i < n && i < (uint64_t)(c + 1) * 64; ++i) {
            const Ray ray = Ray(vec3(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]), vec3(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]));
            const Hit hit = traceGeometry(ray, prev ? prev[i] : noRayHit);
            object[i] = hit.hitT == noHit ? noRayHit : hit.object;
            t[i] = hit.hitT;
            if (uv) { uv[2 * i] = hit.uv.x; uv[2 * i + 1] = hit.uv.y; }
            if (enc_normal2) { uvec2 e = encodeNormal(hit.objectNormal); enc_normal2[2 * i] = e.x; enc_normal2[2 * i + 1] = e.y; }
        }
    });
}

void ref_occlusion_rays(const ref_bind* b, const float* rays, uint64_t n, const float* max_dist, const uint32_t* prev,
                        uint8_t* occluded) {
    bindScene(b);
    parallel_for((int64_t)((n + 63) / 64), [&](int64_t c) {
        for (uint64_t i = (uint64_t)c * 64; i < n && i < (uint64_t)(c + 1) * 64; ++i) {
            const Ray ray = Ray(vec3(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]), vec3(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]));
            occluded[i] = traceOcclusion(ray, max_dist ? max_dist[i] : noHit, prev ? prev[i] : noRayHit) ? 1 : 0;
        }
    });
}

void ref_primary_rays(const ref_bind* b, float* rays_out) {
    bindScene(b);
    memcpy(&seed, b->seed24, 24);
    for (uint y = 0; y < b->height; ++y)
        for (uint x = 0; x < b->width; ++x) {
            const Ray ray = calculatePrimary(uvec2(x, y), vec2(seed.randomX, seed.randomY));
            float* o = rays_out + 6 * ((size_t)y * b->width + x);
            o[0] = ray.pos.x; o[1] = ray.pos.y; o[2] = ray.pos.z; o[3] = ray.dir.x; o[4] = ray.dir.y; o[5] = ray.dir.z;
        }
}
#endif

#if REF_KIND == REF_SHADOW
void ref_shadow(const ref_bind* b) {
    bindScene(b);
    memcpy(&seed, b->seed24, 24);
    totalSamples = b->samples;
    shadowOutput32 = b->shadow_bits;
    dirObject = sampler2D{b->dirT, (int)b->width, (int)b->height, FMT_RGBA32F, 0};
    // Dispatch(w, h, samples), local size 16x16x2; a subgroup = 32 consecutive local invocation indices (x + 16*y + 256*z)
    const uint gx = (b->width + 15) / 16, gy = (b->height + 15) / 16, gz = (b->samples + 1) / 2;
    parallel_for((int64_t)gz * gy, [&](int64_t job) {
        const uint gzi = (uint)(job / gy), gyi = (uint)(job % gy);
        for (uint gxi = 0; gxi < gx; ++gxi)
            for (uint sub = 0; sub < 16; ++sub) {   // 512 invocations per group = 16 subgroups
                uvec3 ids[32];
                for (uint l = 0; l < 32; ++l) {
                    const uint li = sub * 32 + l;
                    ids[l] = uvec3(gxi * 16 + (li & 15u), gyi * 16 + ((li >> 4) & 15u), gzi * 2 + (li >> 8));
                }
                run_subgroup(shader_main, ids);
            }
    });
}
#endif

#if REF_KIND == REF_LIGHTING
void ref_lighting(const ref_bind* b) {
    bindScene(b);
    memset(&seed, 0, sizeof seed);   // D8: the host never binds SSBO 8 (ref: src/rt/task/shadow_task.cpp:59-61,128-130); reads return zero
    totalSamples = b->samples;
    shadowOutput32 = b->shadow_bits;
    dirObject = sampler2D{b->dirT, (int)b->width, (int)b->height, FMT_RGBA32F, 0};
    uvNormal = sampler2D{b->uvN, (int)b->width, (int)b->height, FMT_RGBA32F, 0};
    lighting = image2D{b->lighting, (int)b->width, (int)b->height, FMT_RGBA16F};
    dispatch2D(b->width, b->height);
}
#endif

#if REF_KIND == REF_COMPOSITE
void ref_composite(const ref_bind* b) {
    bindScene(b);
    memcpy(&seed, b->seed24, 24);
    rayOutput = image2D{b->rgba8, (int)b->width, (int)b->height, FMT_RGBA8};
    accumulation = image2D{b->accum, (int)b->width, (int)b->height, FMT_RGBA32F};
    ui = sampler2DMS{nullptr, (int)b->width, (int)b->height, 1};
    dirObject = sampler2D{b->dirT, (int)b->width, (int)b->height, FMT_RGBA32F, 0};
    uvNormal = sampler2D{b->uvN, (int)b->width, (int)b->height, FMT_RGBA32F, 0};
    cloutput = sampler2D{nullptr, 0, 0, FMT_RGBA32F, 0};   // never sampled: the cloud term is the constant vec4(0) (ref: composite.comp:93-97)
    lighting = sampler2D{b->lighting, (int)b->width, (int)b->height, FMT_RGBA16F, 0};
#ifdef DEBUG
    debugType = b->debug_type;
    nanOnly = b->nan_only;
#endif
    dispatch2D(b->width, b->height);
}
#endif

}  // extern "C"
