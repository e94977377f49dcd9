/*
 * ref_api.h — C interface of oracle/_ref/libigxref_{debug,release}.so: the reference's own shader sources
 * (res/shaders/*.glsl, *.comp @ 24f24ea2), adapted for syntax by glsl_front.py and compiled for the host through glsl_shim.h.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/oracle.h): loaded by tests/ and scripts/make_golden.py to pin the hand-written oracle.
 * Each ref_<pass> call runs that shader's main() once per invocation of the grid the reference dispatches (ref:
 * src/rt/task/raygen_task.cpp:88-94, shadow_task.cpp:194-215, composite_task.cpp:253-276); the shadow pass runs each 32-wide
 * subgroup in lock-step (fibres) so that ballotARB means what it means on the GPU.
 */
#ifndef REF_API_H
#define REF_API_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct ref_bind {
    const void* camera144;            /* CameraData UBO */
    const uint32_t* scene_info9;      /* SceneData UBO */
    void* seed24;                     /* SeedBuffer (init.comp writes it) */
    const void* triangles; const void* spheres; const void* cubes; const void* planes;
    const void* lights; const void* materials; const uint32_t* material_indices;
    const uint16_t* skybox;           /* rgba16f or NULL (textureSize == 0 -> camera.skyboxColor) */
    uint32_t sky_w, sky_h;
    uint32_t width, height, samples;  /* dispatch size; ShadowProperties.totalSamples */
    float* dirT;                      /* rgba32f */
    float* uvN;                       /* rgba32f */
    uint32_t* shadow_bits;            /* ShadowOutput32 */
    uint16_t* lighting;               /* rgba16f */
    float* accum;                     /* rgba32f */
    uint32_t* rgba8;                  /* rgba8 */
    uint32_t debug_type, nan_only;    /* DebugData UBO of the DEBUG composite (0, 0 = the default view) */
} ref_bind;

int  ref_is_debug(void);              /* 1: compiled with -DDEBUG (what the shipped .spv are), 0: -DRELEASE */
void ref_set_threads(int n);
void ref_init(const ref_bind* b);
void ref_raygen(const ref_bind* b);
void ref_shadow(const ref_bind* b);
void ref_lighting(const ref_bind* b);
void ref_composite(const ref_bind* b);
/* function-level entries: the reference's traceGeometry / traceOcclusion / encodeNormal on explicit rays (6 floats each) */
void ref_trace_rays(const ref_bind* b, const float* rays, uint64_t n, const uint32_t* prev, uint32_t* object, float* t,
                    float* uv, uint32_t* enc_normal2);
void ref_occlusion_rays(const ref_bind* b, const float* rays, uint64_t n, const float* max_dist, const uint32_t* prev,
                        uint8_t* occluded);
/* the reference's calculatePrimary for every pixel: 6 floats (origin, dir) per pixel */
void ref_primary_rays(const ref_bind* b, float* rays_out);

#ifdef __cplusplus
}
#endif
#endif
