/*
 * oracle.h — C interface of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing under oracle/ is part of the product. Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * The oracle is a from-scratch CPU restatement of the igx_raytracing hot path (reference
 * commit 24f24ea2): host-side packing (igx/include/types/scene_object_types.hpp, core2 flp/vec/mat),
 * camera set-up (src/rt/structs.cpp, src/rt/raytracing_interface.cpp:279-328), the default scene
 * (test/scene/niels_scene.cpp) and the five compute passes res/shaders/{init,raygen,nv_all.shadow,
 * nv_all.lighting,composite}.comp with their includes. Every function in oracle.cpp cites the
 * reference file:line it follows.
 *
 * PARITY PINNING: the reference's host (Windows/WGL OpenGL) cannot run here, but its SHADER SOURCES can: oracle/ref_shim/
 * compiles res/shaders/{init,raygen,nv_all.shadow,nv_all.lighting,composite}.comp and their includes, read where they lie
 * under /root/reference and adapted for syntax only, into oracle/_ref/libigxref_{debug,release}.so.  This oracle is held
 * bit-equal to that library (tests/test_oracle_vs_ref.py: whole frames of every fixture, explicit rays through
 * traceGeometry / traceOcclusion, both shader builds), and the committed fixtures under tests/golden/ are that library's
 * output (scripts/make_golden.py).  Host packing is pinned by the 22 f32->f16 known answers of core2/test/test.cpp:8-29.
 * What stays decreed rather than pinned is what GLSL leaves to the driver (D1-D9 in oracle.cpp: operation rounding,
 * transcendental precision, sampler arithmetic); the shim and the oracle implement the same decrees.
 */
#ifndef ORACLE_H
#define ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Views over the raw GPU-layout buffers the reference uploads. */
typedef struct orc_scene {
    const void* triangles;          /* 48 B each  */
    const void* spheres;            /* 16 B each  */
    const void* cubes;              /* 24 B each  */
    const void* planes;             /* 16 B each  */
    const void* lights;             /* 32 B each  */
    const void* materials;          /* 32 B each  */
    const uint32_t* material_indices;
    const uint16_t* skybox;         /* rgba16f, row 0 first; NULL => camera.skyboxColor */
    uint32_t info[9];               /* lightCount, materialCount, triangleCount, sphereCount, cubeCount,
                                       planeCount, directionalLightCount, spotLightCount, pointLightCount */
    uint32_t sky_w, sky_h;
    uint32_t _pad;
} orc_scene;

enum {
    ORC_FLAG_EDGE = 1,      /* a candidate hit lies within eps_bary of a primitive edge / silhouette */
    ORC_FLAG_TIE = 2,       /* two candidates within eps_t relative distance of the nearest */
    ORC_FLAG_PARALLEL = 4,  /* near-parallel triangle / plane, |a| tiny */
    ORC_FLAG_NAN = 8        /* NaN or inf met while evaluating a candidate */
};

enum { ORC_MODE_DEBUG = 0, ORC_MODE_RELEASE = 1 };   /* which build of the shaders is modelled (oracle.cpp D10) */
void orc_set_mode(int mode);
/* Beyond the reference (definitions in oracle.cpp next to g_all_lights): every light evaluated instead of
   light 0 x lightCount (shadow layers = light * samples + sample), and the temporal History blend of the lighting texture */
void orc_set_all_lights(int on);
void orc_set_history(float alpha);
void orc_history_blend(uint16_t* history_f16, uint16_t* lighting_f16, uint64_t texels, int first);
int  orc_get_mode(void);
void orc_set_threads(int n);
int  orc_get_threads(void);

/* ---- host packing (H1-H3, H6) ---- */
uint16_t orc_f16_trunc(float v);                       /* core2 flp.hpp:124-162 */
float    orc_f16_to_f32(uint16_t h);                   /* IEEE half -> float (GLSL unpackHalf2x16) */
uint16_t orc_f32_to_f16_rtne(float v);                 /* image store to rgba16f */
void orc_spheremap(const float n[3], uint16_t out[2]); /* scene_object_types.hpp:68-72 */
void orc_encode_normal_cpu(const float n[3], uint32_t out[2]); /* scene_object_types.hpp:131-139 */
void orc_triangle_flat(const float p[9], void* out48); /* scene_object_types.hpp:98-106 */
void orc_triangle_normals(const float p[9], const float n[9], void* out48); /* :87-96 */
void orc_light_directional(const float dir[3], const float color[3], float angular_extent, void* out32);
void orc_light_point(const float pos[3], const float color[3], float rad, float origin, float specularity, void* out32);
void orc_material(const float albedo[3], const float ambient[3], const float emission[3],
                  float metallic, float roughness, float transparency, void* out32);
/* CPUCamera + RaytracingInterface::resize/update camera part. Angles in radians, fov in degrees. */
void orc_camera(const float eye[3], float pitch, float yaw, float roll, float left_fov, float right_fov,
                float ipd, uint32_t projection, uint32_t width, uint32_t height, uint32_t flags,
                float exposure, const float skybox_color[3], void* out144);
/* NielsScene at a given time; buffers sized for >= 3 tris, 7 spheres, 2 cubes, 1 plane, 3 lights, 8 mats, 13 idx. */
void orc_niels_scene(double time, void* tris, void* spheres, void* cubes, void* planes, void* lights,
                     void* materials, uint32_t* material_indices, uint32_t info[9]);
/* Radiance .hdr -> rgba16f as igxi convert.cpp does. out==NULL: only query w,h. Returns 0 on success. */
int orc_load_hdr(const char* path, uint16_t* out, uint32_t* w, uint32_t* h);

/* the synthetic scenes of BASELINE.json configs[2] / configs[3] (byte-equal to rtb_gen_soup / rtb_gen_heightfield) */
void orc_gen_soup(uint64_t n, uint64_t seed, void* out_triangles48);
void orc_gen_heightfield(uint32_t grid, uint64_t seed, void* out_triangles48);

/* ---- device passes (K0-K4) ---- */
void orc_init_pass(void* seed24);
/* rays_out: 6 floats/pixel (origin, dir) or NULL; flags_out: 1 byte/pixel or NULL. */
void orc_raygen(const orc_scene* s, const void* cam144, const void* seed24,
                float* dirT, float* uvN, float* rays_out, uint8_t* flags_out);
/* rays-in mode: nearest hit for explicit rays. prev may be NULL (= noRayHit for all). */
void orc_trace_rays(const orc_scene* s, const float* rays, uint64_t n, const uint32_t* prev,
                    uint32_t* object, float* t, float* uv, uint32_t* enc_normal2, uint8_t* flags_out);
/* occlusion for explicit rays: hit = traceOcclusion(ray, max_dist, prev) */
void orc_occlusion_rays(const orc_scene* s, const float* rays, uint64_t n, const float* max_dist,
                        const uint32_t* prev, uint8_t* occluded);
void orc_shadow(const orc_scene* s, const void* cam144, const void* seed24, uint32_t samples,
                const float* dirT, uint32_t* bits, float* shadow_rays_out);
void orc_lighting(const orc_scene* s, const void* cam144, uint32_t samples, const float* dirT,
                  const float* uvN, const uint32_t* bits, uint16_t* lighting_f16, float* lighting_f32);
void orc_composite(const orc_scene* s, const void* cam144, const void* seed24, const float* dirT,
                   const float* uvN, const uint16_t* lighting_f16, float* accum, uint32_t* rgba8);
/* whole frame K0..K4 (seed updated in place). Any output may be NULL except rgba8. */
void orc_frame(const orc_scene* s, const void* cam144, void* seed24, uint32_t samples,
               float* dirT, float* uvN, uint32_t* bits, uint16_t* lighting_f16, float* accum, uint32_t* rgba8);
/* Diffuse bounces (BASELINE.json configs[3]).  NOT reference behaviour — the reference traces no secondary rays; this restates the
   definition in igx_raytracing_b200/csrc/rtb_path.cuh from the reference functions so that the CUDA path has a CPU statement of
   the same semantics to be compared with.  K0 + one path-traced frame; depth 0 is the reference's G-buffer.  dirT / uvN / accum /
   radiance3 / rays_out may be NULL (accum must be given under USE_SUPERSAMPLING). */
void orc_path_frame(const orc_scene* s, const void* cam144, void* seed24, uint32_t bounces, float* dirT, float* uvN, float* accum,
                    uint32_t* rgba8, float* radiance3, uint64_t* rays_out);
/* K1..K4 for a list of pixels only (bounded CPU-baseline sample); seed must already be initialised
   by orc_init_pass. Returns the number of rays traced (primary + shadow). */
uint64_t orc_frame_pixels(const orc_scene* s, const void* cam144, const void* seed24, uint32_t samples,
                          const uint32_t* xy, uint64_t n, uint32_t* rgba8, uint32_t* object, float* t);

/* the same with more outputs (any may be NULL): the G-buffer texel dirT (4 floats per pixel), sample 0's shadow bit, and the
   edge / tie / degeneracy flags of the primary ray (ORC_FLAG_*) */
uint64_t orc_frame_pixels_ex(const orc_scene* s, const void* cam144, const void* seed24, uint32_t samples, const uint32_t* xy,
                             uint64_t n, uint32_t* rgba8, uint32_t* object, float* t, float* dirT4, uint8_t* shadowed, uint8_t* flags);

#ifdef __cplusplus
}
#endif
#endif
