/*
 * oracle.cpp — CPU oracle for the igx_raytracing hot path. TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * A from-scratch restatement (new code, same maths) of the reference @ 24f24ea2. "SH/" below means
 * /root/reference/res/shaders/, "CORE2/" means igx/igxi-tool/igxi/ignis/core2/.
 *
 * Build: strict IEEE f32, no contraction (see Makefile: -O2 -ffp-contract=off -fno-fast-math).
 *
 * Decrees for what GLSL leaves implementation-defined (stated once here, mirrored in DESIGN.md):
 *   D1  +,-,*,/ and sqrt are IEEE-754 binary32 round-to-nearest-even, never fused.
 *   D2  dot(a,b) = ((a.x*b.x + a.y*b.y) + a.z*b.z); cross is the textbook formula.
 *   D3  normalize(v) = v * (1 / sqrt(dot(v,v))); length(v) = sqrt(dot(v,v)).
 *   D4  sin cos asin acos atan atan2 exp pow are the correctly rounded binary32 results, obtained by
 *       evaluating in binary64 and rounding once.
 *   D5  float -> uint conversion truncates; NaN and negatives give 0; >= 2^32 saturates.
 *   D6  imageStore to rgba16f rounds to nearest even; to rgba8 computes floor(clamp(c,0,1)*255 + 0.5), NaN -> 0.
 *   D7  texture() with the nearest sampler at a texel centre is a texel fetch; the skybox bilinear
 *       filter uses full binary32 weights and a (0,0,0,0) border.
 *   D8  lighting.comp reads an unbound Seed SSBO: seed.random = (0,0) there.
 *   D9  smoothstep(e0,e1,x) with e0 > e1 evaluates the Hermite form on clamp((x-e0)/(e1-e0),0,1).
 *   D10 two shader builds exist (SH/compile.sh: -DDEBUG or -DRELEASE) and both are modelled (orc_set_mode):
 *         ORC_MODE_DEBUG (default; the shipped .spv are DEBUG builds): every pass stores for every pixel
 *           (uvObjectNormal and a zero lighting texel on misses, zero shadow words for subgroups without hits), triangles with
 *           p1 == p0 are rejected (SH/primitive.glsl:248-253), composite maps a NaN colour to (0,0,10000) (SH/composite.comp:236-239);
 *         ORC_MODE_RELEASE: those stores are skipped (the targets keep what they held), no zero-edge reject, no NaN mapping.
 *
 * PINNING: oracle/_ref is the reference's own shader source compiled for the host (oracle/ref_shim/); tests/test_oracle_vs_ref.py
 * holds this restatement bit-equal to it in both modes, on whole frames and on explicit rays.
 */
#include "oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>
#include <atomic>
#include <thread>

namespace {

// ------------------------------------------------------------------------------------------------
// minimal GLSL vocabulary
// ------------------------------------------------------------------------------------------------
struct vec2 { float x, y; };
struct vec3 { float x, y, z; };

inline vec2 operator+(vec2 a, vec2 b) { return {a.x + b.x, a.y + b.y}; }
inline vec2 operator-(vec2 a, vec2 b) { return {a.x - b.x, a.y - b.y}; }
inline vec2 operator*(vec2 a, vec2 b) { return {a.x * b.x, a.y * b.y}; }
inline vec2 operator*(vec2 a, float s) { return {a.x * s, a.y * s}; }
inline vec2 operator+(vec2 a, float s) { return {a.x + s, a.y + s}; }
inline vec2 operator/(vec2 a, float s) { return {a.x / s, a.y / s}; }

inline vec3 operator+(vec3 a, vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline vec3 operator-(vec3 a, vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline vec3 operator*(vec3 a, vec3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline vec3 operator/(vec3 a, vec3 b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }
inline vec3 operator*(vec3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline vec3 operator*(float s, vec3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline vec3 operator/(vec3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline vec3 operator+(vec3 a, float s) { return {a.x + s, a.y + s, a.z + s}; }
inline vec3 operator-(vec3 a, float s) { return {a.x - s, a.y - s, a.z - s}; }
inline vec3 operator-(float s, vec3 a) { return {s - a.x, s - a.y, s - a.z}; }
inline vec3 operator-(vec3 a) { return {-a.x, -a.y, -a.z}; }

inline float dot(vec2 a, vec2 b) { return a.x * b.x + a.y * b.y; }                       // D2
inline float dot(vec3 a, vec3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }         // D2
inline vec3 cross(vec3 a, vec3 b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline float length(vec3 v) { return std::sqrt(dot(v, v)); }                             // D3
inline vec3 normalize(vec3 v) { float inv = 1.0f / std::sqrt(dot(v, v)); return v * inv; } // D3
inline vec3 reflect(vec3 i, vec3 n) { return i - (2.0f * dot(n, i)) * n; }
inline float fractf(float x) { return x - std::floor(x); }
inline float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }
inline vec3 mix(vec3 a, vec3 b, float t) { return {mixf(a.x, b.x, t), mixf(a.y, b.y, t), mixf(a.z, b.z, t)}; }
// GLSL leaves min/max with a NaN operand undefined; NVIDIA hardware returns the non-NaN operand (IEEE minNum/maxNum)
inline float glsl_max(float a, float b) { return std::fmax(a, b); }
inline float glsl_min(float a, float b) { return std::fmin(a, b); }
inline vec3 vmax(vec3 a, vec3 b) { return {glsl_max(a.x, b.x), glsl_max(a.y, b.y), glsl_max(a.z, b.z)}; }
inline vec3 vmin(vec3 a, vec3 b) { return {glsl_min(a.x, b.x), glsl_min(a.y, b.y), glsl_min(a.z, b.z)}; }
inline float signf(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }

// D4: correctly rounded binary32 elementary functions
inline float cr_sin(float x) { return (float)std::sin((double)x); }
inline float cr_cos(float x) { return (float)std::cos((double)x); }
inline float cr_asin(float x) { return (float)std::asin((double)x); }
inline float cr_acos(float x) { return (float)std::acos((double)x); }
inline float cr_atan(float x) { return (float)std::atan((double)x); }
inline float cr_atan2(float y, float x) { return (float)std::atan2((double)y, (double)x); }
inline float cr_exp(float x) { return (float)std::exp((double)x); }
inline float cr_pow(float x, float y) { return (float)std::pow((double)x, (double)y); }

inline uint32_t f2u(float f) {   // D5
    if (!(f > 0.0f)) return 0u;
    if (f >= 4294967296.0f) return 0xFFFFFFFFu;
    return (uint32_t)f;
}
inline uint32_t fbits(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline float ubits(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }

int g_mode = 0;   // ORC_MODE_DEBUG / ORC_MODE_RELEASE (D10)
// Beyond the reference (the reference hard-codes lightId = 0 and multiplies by lightCount):
//   g_all_lights   every light gets its own shadow ray and its own Cook-Torrance term; layer = light * samples + sample
//   g_history      temporal blend of the lighting texture through the History texture the reference allocates but never
//                  uses (src/rt/task/shadow_task.cpp:20-22,212 "Do denoising"): history = history * (1 - a) + lighting * a
int g_all_lights = 0;
float g_history_alpha = 0.0f;   // 0 = off

const float noHit = 3.4028235e38f;          // SH/primitive.glsl:6
const uint32_t noRayHit = 0xFFFFFFFFu;      // SH/primitive.glsl:7
const float pi = 3.1415927410125732421875f; // SH/rand_util.glsl:13

// ------------------------------------------------------------------------------------------------
// half floats
// ------------------------------------------------------------------------------------------------

// CORE2/include/types/flp.hpp:124-162 (flp::_init, f32 -> f16): copy sign; zero stays zero; rebias the
// exponent; negative rebias collapses to signed zero; NaN -> all ones; overflow -> inf; else truncate.
uint16_t f16_trunc(float v) {
    uint32_t b = fbits(v);
    uint16_t out = (uint16_t)((b >> 31) << 15);
    uint32_t mant = b & 0x7FFFFFu, rawExp = (b >> 23) & 0xFFu;
    if (mant == 0 && rawExp == 0) return out;
    int32_t ne = (int32_t)rawExp - 127 + 15;
    if (ne < 0) return out;
    if (rawExp == 0xFF && mant) return (uint16_t)(out | (0x1Fu << 10) | 0x3FFu);
    if (ne >= 30 && (ne >= 31 || mant > (0x3FFu << 13))) return (uint16_t)(out | (0x1Fu << 10));
    return (uint16_t)(out | ((uint32_t)ne << 10) | (mant >> 13));
}

// IEEE binary16 -> binary32, exact (GLSL unpackHalf2x16 / rgba16f texel fetch)
float f16_to_f32(uint16_t h) {
    uint32_t sign = (uint32_t)(h >> 15) << 31, e = (h >> 10) & 0x1Fu, m = h & 0x3FFu;
    if (e == 0) {
        if (m == 0) return ubits(sign);
        float f = (float)m * 5.9604644775390625e-8f;   // m * 2^-24, exact
        return sign ? -f : f;
    }
    if (e == 31) return ubits(sign | 0x7F800000u | (m << 13));
    return ubits(sign | ((e + 112u) << 23) | (m << 13));
}

// binary32 -> binary16 round-to-nearest-even (D6)
uint16_t f32_to_f16_rtne(float v) {
    uint32_t b = fbits(v);
    uint32_t sign = (b >> 16) & 0x8000u;
    uint32_t a = b & 0x7FFFFFFFu;
    if (a > 0x7F800000u) return (uint16_t)0x7FFFu;            // NaN -> canonical NaN (what cvt.rn.f16.f32 yields)
    if (a >= 0x477FF000u) return (uint16_t)(sign | 0x7C00u);  // >= 65520 rounds to inf (also inf)
    if (a < 0x33000001u) return (uint16_t)sign;               // <= 2^-25 rounds to zero
    int32_t e = (int32_t)(a >> 23) - 127;
    uint32_t m = (a & 0x7FFFFFu) | 0x800000u;
    if (e < -14) {                                            // subnormal half
        uint32_t shift = (uint32_t)(-14 - e) + 13u;           // 14..24
        uint32_t q = m >> shift, rem = m & ((1u << shift) - 1u), half = 1u << (shift - 1u);
        if (rem > half || (rem == half && (q & 1u))) q++;
        return (uint16_t)(sign | q);
    }
    uint32_t q = ((uint32_t)(e + 15) << 10) | ((m >> 13) & 0x3FFu);
    uint32_t rem = m & 0x1FFFu;
    if (rem > 0x1000u || (rem == 0x1000u && (q & 1u))) q++;
    return (uint16_t)(sign | q);
}

inline vec2 unpackHalf2x16(uint32_t v) { return {f16_to_f32((uint16_t)(v & 0xFFFFu)), f16_to_f32((uint16_t)(v >> 16))}; }

// ------------------------------------------------------------------------------------------------
// host-side vector maths (CORE2/include/types/vec.hpp)
// ------------------------------------------------------------------------------------------------

// vec.hpp:153-165: magnitude = f32(sqrt(f64(sum of squares in f32))); normalize = component / magnitude
inline float host_magnitude3(vec3 v) { float s = 0.0f; s += v.x * v.x; s += v.y * v.y; s += v.z * v.z; return (float)std::sqrt((double)s); }
inline vec3 host_normalize3(vec3 v) { float m = host_magnitude3(v); return {v.x / m, v.y / m, v.z / m}; }
inline vec2 host_normalize2(vec2 v) { float s = 0.0f; s += v.x * v.x; s += v.y * v.y; float m = (float)std::sqrt((double)s); return {v.x / m, v.y / m}; }

// scene_object_types.hpp:68-72
void spheremapTransform(uint16_t& nx, uint16_t& ny, vec3 n) {
    vec2 xy = host_normalize2({n.x, n.y});
    float s = std::sqrt(-n.z * 0.5f + 0.5f);
    nx = f16_trunc(xy.x * s);
    ny = f16_trunc(xy.y * s);
}

// scene_object_types.hpp:131-139
void encodeNormalCpu(vec3 n, uint32_t out[2]) {
    vec3 nn = host_normalize3(n);
    nn = {(nn.x * 0.5f + 0.5f) * 65535.0f, (nn.y * 0.5f + 0.5f) * 65535.0f, (nn.z * 0.5f + 0.5f) * 65535.0f};
    out[0] = ((uint32_t)nn.x << 16) | (uint32_t)nn.y;
    out[1] = (uint32_t)nn.z;
}

struct TriangleRec { float p0[3]; uint16_t n0[2]; float p1[3]; uint16_t n1[2]; float p2[3]; uint16_t n2[2]; };
static_assert(sizeof(TriangleRec) == 48, "Triangle is 48 bytes");
struct LightRec { float pos[3]; uint16_t rad, origin; uint32_t dir[2]; uint16_t r, g, b, type; };
static_assert(sizeof(LightRec) == 32, "Light is 32 bytes");
struct MaterialRec {
    uint16_t albedo[3], metallic; uint16_t ambient[3], roughness; uint16_t emission[3], pad2; float transparency; uint32_t materialInfo;
};
static_assert(sizeof(MaterialRec) == 32, "Material is 32 bytes");
struct CameraRec {   // scene_object_types.hpp:32-66, SH/camera.glsl:14-42
    float eye[3]; uint32_t width; float p0[3]; uint32_t height; float p1[3]; float ipd; float p2[3]; uint32_t projectionType;
    float skyboxColor[3]; float exposure; float p3[3]; float focalDistance; float p4[3]; float aperature; float p5[3]; uint32_t flags;
    float invRes[2]; uint32_t tiles[2];
};
static_assert(sizeof(CameraRec) == 144, "Camera is 144 bytes");
struct SeedRec { float randomX, randomY, cpuOffsetX, cpuOffsetY; uint32_t sampleCount, sampleOffset; };   // include/rt/structs.hpp:8-14
static_assert(sizeof(SeedRec) == 24, "Seed is 24 bytes");

inline vec3 v3(const float* p) { return {p[0], p[1], p[2]}; }

// ------------------------------------------------------------------------------------------------
// device-side: random numbers (SH/rand_util.glsl)
// ------------------------------------------------------------------------------------------------

// rand_util.glsl:115-117
inline float rand1(vec2 co) { return fractf(cr_sin(dot(co, vec2{12.9898f, 78.233f})) * 43758.5453f); }
// rand_util.glsl:127-129; 1103515245 converts to the float 1103515264
inline vec2 rand2(vec2 p) { return {rand1(p), rand1(p * 1103515245.0f + 12345.0f)}; }

// rand_util.glsl:96-103
inline uint32_t uVdC(uint32_t s) {
    s = (s << 16) | (s >> 16);
    s = ((s & 0x55555555u) << 1) | ((s & 0xAAAAAAAAu) >> 1);
    s = ((s & 0x33333333u) << 2) | ((s & 0xCCCCCCCCu) >> 2);
    s = ((s & 0x0F0F0F0Fu) << 4) | ((s & 0xF0F0F0F0u) >> 4);
    s = ((s & 0x00FF00FFu) << 8) | ((s & 0xFF00FF00u) >> 8);
    return s;
}
// rand_util.glsl:105-111
inline vec2 hammersley(uint32_t i, uint32_t N) { return {(float)i / (float)N, (float)uVdC(i) * 2.3283064365386963e-10f}; }

// rand_util.glsl:33-39
vec3 randomPointOnUnitSphere(vec2 r) {
    vec2 polar = {(2.0f * pi) * r.x, cr_acos(1.0f - 2.0f * r.y)};
    vec2 s = {cr_sin(polar.x), cr_sin(polar.y)}, c = {cr_cos(polar.x), cr_cos(polar.y)};
    return {s.x * c.y, s.x * s.y, c.x};
}
// rand_util.glsl:43-51 (the flip multiplies by +1: a no-op, kept)
vec3 randomPointOnHemisphere(vec2 r, vec3 origin, float size, vec3 /*l*/) {
    vec3 n = randomPointOnUnitSphere(r);
    return origin + n * size;
}
// rand_util.glsl:55-64
vec3 getPerpendicularVector(vec3 u) {
    vec3 a = {std::fabs(u.x), std::fabs(u.y), std::fabs(u.z)};
    uint32_t xm = (a.x - a.y < 0.0f && a.x - a.z < 0.0f) ? 1u : 0u;
    uint32_t ym = (a.y - a.z < 0.0f) ? (1u ^ xm) : 0u;
    uint32_t zm = 1u ^ (xm | ym);
    return cross(u, vec3{(float)xm, (float)ym, (float)zm});
}
// rand_util.glsl:66-85
vec3 getSunDirection(vec2 random, vec3 direction, float angularExtent) {
    float s = random.x, r = random.y;
    float h = cr_cos(angularExtent);
    float phi = (2.0f * pi) * s;
    float z = h + (1.0f - h) * r;
    float sinT = std::sqrt(1.0f - z * z);
    float x = cr_cos(phi) * sinT;
    float y = cr_sin(phi) * sinT;
    vec3 bitangent = getPerpendicularVector(direction);
    vec3 tangent = cross(bitangent, direction);
    return bitangent * x + tangent * y + direction * z;
}

// ------------------------------------------------------------------------------------------------
// device-side: primitives (SH/primitive.glsl)
// ------------------------------------------------------------------------------------------------
struct Ray { vec3 pos, dir; };
struct Hit { vec3 rayDir; float hitT; vec2 uv; uint32_t object; vec3 geometryNormal; vec3 objectNormal; };

// primitive.glsl:85-88 (GPU encodeNormal)
inline void encodeNormalGpu(vec3 n, uint32_t out[2]) {
    vec3 v = (normalize(n) * 0.5f + 0.5f) * 65535.0f;
    uint32_t x = f2u(v.x), y = f2u(v.y), z = f2u(v.z);
    out[0] = (x << 16) | y;
    out[1] = z;
}
// primitive.glsl:90-93
inline vec3 decodeNormal(uint32_t sx, uint32_t sy) {
    vec3 nh = {(float)(sx >> 16), (float)(sx & 65535u), (float)sy};
    return nh / 65535.0f * 2.0f - 1.0f;
}
// primitive.glsl:95-104
inline vec3 decodeSpheremap(uint32_t n) {
    vec2 nn = unpackHalf2x16(n);
    float l = dot(vec3{nn.x, nn.y, 1.0f}, -vec3{nn.x, nn.y, -1.0f});
    float sq = std::sqrt(l);
    nn = nn * sq;
    return vec3{nn.x, nn.y, l} * 2.0f + vec3{0.0f, 0.0f, -1.0f};
}
// primitive.glsl:128-142
inline vec3 unpackColor3(uint32_t cx, uint32_t cy) { vec2 rg = unpackHalf2x16(cx); return {rg.x, rg.y, unpackHalf2x16(cy).x}; }
inline uint32_t unpackColorA(uint32_t cy) { return cy >> 16; }
inline float unpackColorAUnorm(uint32_t cy) { return (float)(cy >> 16) / 65535.0f; }
// primitive.glsl:146-151
inline vec3 interpolate(vec3 a, vec3 b, vec3 c, vec2 uv) {
    vec3 bary = {uv.x, uv.y, 1.0f - uv.x - uv.y};
    return bary.x * b + bary.y * c + bary.z * a;
}

// primitive.glsl:173-210
bool rayIntersectSphere(const Ray& r, const float* sph, Hit& hit, uint32_t obj, uint32_t prevObj) {
    vec3 c = v3(sph);
    vec3 dif = c - r.pos;
    float t = dot(dif, r.dir);
    vec3 Q = dif - t * r.dir;
    float Q2 = dot(Q, Q);
    float R2 = sph[3] * sph[3];
    bool outOfSphere = Q2 > R2;
    float hitT = t - std::sqrt(R2 - Q2);
    if (!outOfSphere && obj != prevObj && hitT >= 0.0f && hitT < hit.hitT) {
        hit.hitT = hitT;
        vec3 o = hitT * r.dir + r.pos;
        vec3 normal = normalize(c - o);
        hit.geometryNormal = normal;
        float latitude = cr_asin(normal.z);
        float longitude = cr_atan(normal.y / normal.x);
        if (std::isnan(longitude)) longitude = 0.0f;
        hit.uv = vec2{latitude, longitude} * (0.636619746685f * 0.5f) + 0.5f;
        return true;
    }
    return false;
}

// primitive.glsl:212-237
bool rayIntersectPlane(const Ray& r, const float* pl, Hit& hit, uint32_t obj, uint32_t prevObj) {
    vec3 pxyz = v3(pl);
    vec3 dir = normalize(pxyz);
    float dif = dot(r.dir, -dir);
    float hitT = -(dot(r.pos, -dir) + pl[3]) / dif;
    if (hitT >= 0.0f && obj != prevObj && hitT < hit.hitT) {
        hit.hitT = hitT;
        hit.geometryNormal = dif > 0.0f ? -dir : dir;
        vec3 o = hitT * r.dir + r.pos;
        vec3 planeX = cross(pxyz, vec3{0, 0, 1});
        vec3 planeZ = cross(pxyz, vec3{1, 0, 0});
        hit.uv = {dot(o, planeX), dot(o, planeZ)};
        return true;
    }
    return false;
}

// primitive.glsl:239-284 (Möller–Trumbore; the zero-edge reject at :248-253 exists in DEBUG builds only, see D10)
bool rayIntersectTri(const Ray& r, const TriangleRec& tri, Hit& hit, uint32_t obj, uint32_t prevObj) {
    vec3 p0 = v3(tri.p0), p1 = v3(tri.p1), p2 = v3(tri.p2);
    vec3 p1_p0 = p1 - p0, p2_p0 = p2 - p0;
    if (g_mode == ORC_MODE_DEBUG && p1_p0.x == 0.0f && p1_p0.y == 0.0f && p1_p0.z == 0.0f) return false;
    vec3 h = cross(r.dir, p2_p0);
    float a = dot(p1_p0, h);
    if (std::fabs(a) < 0.0f) return false;   // never true; kept as in the reference
    float f = 1.0f / a;
    vec3 s = r.pos - p0;
    float u = f * dot(s, h);
    if (u < 0.0f || u > 1.0f) return false;
    vec3 q = cross(s, p1_p0);
    float v = f * dot(r.dir, q);
    if (v < 0.0f || u + v > 1.0f) return false;
    float t = f * dot(p2_p0, q);
    if (t <= 0.0f || obj == prevObj || t >= hit.hitT) return false;
    hit.uv = {u, v};
    hit.hitT = t;
    hit.geometryNormal = cross(normalize(p1 - p0), normalize(p2 - p0)) * -signf(a);
    return true;
}

// primitive.glsl:286-333
bool rayIntersectCube(const Ray& r, const float* cube, Hit& hit, uint32_t obj, uint32_t prevObj) {
    vec3 revDir = {1.0f / r.dir.x, 1.0f / r.dir.y, 1.0f / r.dir.z};
    vec3 start = {cube[0], cube[1], cube[2]};
    vec3 end = {cube[3], cube[4], cube[5]};
    vec3 startDir = (start - r.pos) * revDir;
    vec3 endDir = (end - r.pos) * revDir;
    vec3 mi = vmin(startDir, endDir);
    vec3 ma = vmax(startDir, endDir);
    float tmin = glsl_max(glsl_max(mi.x, mi.y), mi.z);
    float tmax = glsl_min(glsl_min(ma.x, ma.y), ma.z);
    if (tmax < 0.0f || tmin > tmax || tmin > hit.hitT || obj == prevObj) return false;
    vec3 pos = (r.dir * tmin + r.pos) - start;
    pos = pos / end;
    if (tmin == mi.x) {
        int isLeft = (mi.x == startDir.x) ? 1 : 0;
        hit.geometryNormal = {(float)(isLeft * 2 - 1), 0, 0};
        hit.uv = {pos.y, pos.z};
    } else if (tmin == mi.y) {
        int isDown = (mi.y == startDir.y) ? 1 : 0;
        hit.geometryNormal = {0, (float)(isDown * 2 - 1), 0};
        hit.uv = {pos.x, pos.z};
    } else {
        int isBack = (mi.z == startDir.z) ? 1 : 0;
        hit.geometryNormal = {0, 0, (float)(isBack * 2 - 1)};
        hit.uv = {pos.x, pos.y};
    }
    hit.hitT = tmin;
    return true;
}

// ------------------------------------------------------------------------------------------------
// device-side: scene views + traces (SH/scene.glsl, SH/trace.glsl)
// ------------------------------------------------------------------------------------------------
struct Scene {
    const TriangleRec* tri; const float* sph; const float* cube; const float* plane;
    const LightRec* light; const MaterialRec* mat; const uint32_t* matIdx; const uint16_t* sky;
    uint32_t lightCount, materialCount, triangleCount, sphereCount, cubeCount, planeCount;
    uint32_t skyW, skyH;
};
Scene view(const orc_scene* s) {
    Scene o;
    o.tri = (const TriangleRec*)s->triangles; o.sph = (const float*)s->spheres; o.cube = (const float*)s->cubes;
    o.plane = (const float*)s->planes; o.light = (const LightRec*)s->lights; o.mat = (const MaterialRec*)s->materials;
    o.matIdx = s->material_indices; o.sky = s->skybox;
    o.lightCount = s->info[0]; o.materialCount = s->info[1]; o.triangleCount = s->info[2];
    o.sphereCount = s->info[3]; o.cubeCount = s->info[4]; o.planeCount = s->info[5];
    o.skyW = s->skybox ? s->sky_w : 0; o.skyH = s->skybox ? s->sky_h : 0;
    return o;
}

// trace.glsl:9-66
Hit traceGeometry(const Scene& sc, const Ray& ray, uint32_t prevHit) {
    Hit hit;
    hit.rayDir = ray.dir;
    hit.hitT = noHit;
    hit.uv = {0, 0};
    hit.object = 0;
    hit.geometryNormal = {0, 0, 0};
    uint32_t j = 0;
    for (uint32_t i = 0; i < sc.triangleCount; ++i, ++j)
        if (rayIntersectTri(ray, sc.tri[i], hit, j, prevHit)) hit.object = j;
    for (uint32_t i = 0; i < sc.sphereCount; ++i, ++j)
        if (rayIntersectSphere(ray, sc.sph + 4 * i, hit, j, prevHit)) hit.object = j;
    for (uint32_t i = 0; i < sc.cubeCount; ++i, ++j)
        if (rayIntersectCube(ray, sc.cube + 6 * i, hit, j, prevHit)) hit.object = j;
    for (uint32_t i = 0; i < sc.planeCount; ++i, ++j)
        if (rayIntersectPlane(ray, sc.plane + 4 * i, hit, j, prevHit)) hit.object = j;
    if (hit.object < sc.triangleCount) {
        const TriangleRec& t = sc.tri[hit.object];
        uint32_t e0, e1, e2;
        std::memcpy(&e0, t.n0, 4); std::memcpy(&e1, t.n1, 4); std::memcpy(&e2, t.n2, 4);
        hit.objectNormal = interpolate(decodeSpheremap(e0), decodeSpheremap(e1), decodeSpheremap(e2), hit.uv);
    } else
        hit.objectNormal = hit.geometryNormal;
    return hit;
}

// trace.glsl:70-98 (no early out, as in the reference)
bool traceOcclusion(const Scene& sc, const Ray& ray, float maxDist, uint32_t prevHit) {
    Hit hit;
    hit.hitT = noHit;
    hit.rayDir = ray.dir; hit.uv = {0, 0}; hit.object = 0; hit.geometryNormal = {0, 0, 0}; hit.objectNormal = {0, 0, 0};
    uint32_t j = 0;
    for (uint32_t i = 0; i < sc.triangleCount; ++i, ++j) rayIntersectTri(ray, sc.tri[i], hit, j, prevHit);
    for (uint32_t i = 0; i < sc.sphereCount; ++i, ++j) rayIntersectSphere(ray, sc.sph + 4 * i, hit, j, prevHit);
    for (uint32_t i = 0; i < sc.cubeCount; ++i, ++j) rayIntersectCube(ray, sc.cube + 6 * i, hit, j, prevHit);
    for (uint32_t i = 0; i < sc.planeCount; ++i, ++j) rayIntersectPlane(ray, sc.plane + 4 * i, hit, j, prevHit);
    return hit.hitT < maxDist;
}

// Edge / tie / degeneracy flags for one ray whose nearest accepted distance is bestT.
// Not part of the reference: it marks rays on which a numerically different but equally valid
// evaluation order (a BVH, fused multiply-adds) may legitimately return another primitive.
const float EPS_BARY = 1e-5f, EPS_T = 1e-5f;
uint8_t rayFlags(const Scene& sc, const Ray& r, uint32_t prevHit, float bestT) {
    uint8_t fl = 0;
    const bool miss = !(bestT < noHit);
    if (std::isnan(bestT)) fl |= ORC_FLAG_NAN;
    const float lim = miss ? noHit : (bestT + EPS_T * std::fabs(bestT));
    int nearCount = 0;
    uint32_t j = 0;
    for (uint32_t i = 0; i < sc.triangleCount; ++i, ++j) {
        if (j == prevHit) continue;
        const TriangleRec& tri = sc.tri[i];
        vec3 p0 = v3(tri.p0), e1 = v3(tri.p1) - p0, e2 = v3(tri.p2) - p0;
        vec3 h = cross(r.dir, e2);
        float a = dot(e1, h);
        float f = 1.0f / a;
        vec3 s = r.pos - p0;
        float u = f * dot(s, h);
        vec3 q = cross(s, e1);
        float v = f * dot(r.dir, q);
        float t = f * dot(e2, q);
        if (std::isnan(u) || std::isnan(v) || std::isnan(t)) { fl |= ORC_FLAG_NAN; continue; }
        if (!(t > 0.0f) || t > lim) continue;
        float w = 1.0f - u - v;
        bool nearInside = u >= -EPS_BARY && v >= -EPS_BARY && w >= -EPS_BARY;
        if (!nearInside) continue;
        float scale = std::sqrt(dot(e1, e1)) * std::sqrt(dot(e2, e2));
        if (std::fabs(a) < 1e-7f * scale) fl |= ORC_FLAG_PARALLEL;
        if (std::fabs(u) < EPS_BARY || std::fabs(v) < EPS_BARY || std::fabs(w) < EPS_BARY) fl |= ORC_FLAG_EDGE;
        if (u >= 0.0f && v >= 0.0f && u + v <= 1.0f) nearCount++;
    }
    for (uint32_t i = 0; i < sc.sphereCount; ++i, ++j) {
        if (j == prevHit) continue;
        const float* sph = sc.sph + 4 * i;
        vec3 dif = v3(sph) - r.pos;
        float t = dot(dif, r.dir);
        vec3 Q = dif - t * r.dir;
        float Q2 = dot(Q, Q), R2 = sph[3] * sph[3];
        if (std::fabs(Q2 - R2) <= 1e-4f * R2 && t >= 0.0f && t - std::sqrt(glsl_max(R2 - Q2, 0.0f)) <= lim) fl |= ORC_FLAG_EDGE;  // grazing
        if (Q2 > R2) continue;
        float hitT = t - std::sqrt(R2 - Q2);
        if (std::fabs(hitT) <= EPS_T) fl |= ORC_FLAG_EDGE;      // origin on the surface
        if (hitT >= 0.0f && hitT <= lim) nearCount++;
    }
    for (uint32_t i = 0; i < sc.cubeCount; ++i, ++j) {
        if (j == prevHit) continue;
        const float* c = sc.cube + 6 * i;
        vec3 rd = {1.0f / r.dir.x, 1.0f / r.dir.y, 1.0f / r.dir.z};
        vec3 a = (vec3{c[0], c[1], c[2]} - r.pos) * rd, b = (vec3{c[3], c[4], c[5]} - r.pos) * rd;
        if (std::isnan(a.x) || std::isnan(a.y) || std::isnan(a.z) || std::isnan(b.x) || std::isnan(b.y) || std::isnan(b.z)) { fl |= ORC_FLAG_NAN; continue; }
        vec3 mi = vmin(a, b), ma = vmax(a, b);
        float tmin = glsl_max(glsl_max(mi.x, mi.y), mi.z), tmax = glsl_min(glsl_min(ma.x, ma.y), ma.z);
        float slack = EPS_T * (std::fabs(tmin) + std::fabs(tmax) + 1.0f);
        if (tmax < -slack || tmin > tmax + slack || tmin > lim) continue;
        if (std::fabs(tmax) <= slack || std::fabs(tmin - tmax) <= slack) fl |= ORC_FLAG_EDGE;   // touching / corner graze
        // two entry slabs nearly equal: the face (and so uv / normal) is ambiguous
        float m2 = glsl_min(glsl_min(glsl_max(mi.x, mi.y), glsl_max(mi.y, mi.z)), glsl_max(mi.x, mi.z));
        if (std::fabs(tmin - m2) <= slack) fl |= ORC_FLAG_EDGE;
        if (!(tmax < 0.0f) && !(tmin > tmax)) nearCount++;
    }
    for (uint32_t i = 0; i < sc.planeCount; ++i, ++j) {
        if (j == prevHit) continue;
        const float* pl = sc.plane + 4 * i;
        vec3 dir = normalize(v3(pl));
        float dif = dot(r.dir, -dir);
        if (std::fabs(dif) < 1e-6f) { fl |= ORC_FLAG_PARALLEL; continue; }
        float hitT = -(dot(r.pos, -dir) + pl[3]) / dif;
        if (std::isnan(hitT)) { fl |= ORC_FLAG_NAN; continue; }
        if (hitT >= 0.0f && hitT <= lim) nearCount++;
    }
    if (!miss && nearCount > 1) fl |= ORC_FLAG_TIE;
    return fl;
}

// ------------------------------------------------------------------------------------------------
// device-side: camera (SH/camera.glsl)
// ------------------------------------------------------------------------------------------------

// camera.glsl:53-70
Ray calculateOmni(const CameraRec& cam, vec2 c, bool isLeft) {
    vec2 spherical = vec2{c.x - 0.5f, 0.5f - c.y} * vec2{2.0f * pi, pi};
    vec2 sins = {cr_sin(spherical.x), cr_sin(spherical.y)}, coss = {cr_cos(spherical.x), cr_cos(spherical.y)};
    vec3 pos = v3(cam.eye) + vec3{coss.x, 0.0f, sins.x} * (cam.ipd * 5e-4f) * (isLeft ? -1.0f : 1.0f);
    vec3 dir = {sins.x * coss.y, sins.y, -coss.x * coss.y};
    return {pos, dir};
}
// camera.glsl:72-87
Ray calculateScreen(const CameraRec& cam, vec2 c, bool isRight) {
    vec3 p0 = isRight ? v3(cam.p3) : v3(cam.p0);
    vec3 p1 = isRight ? v3(cam.p4) : v3(cam.p1);
    vec3 p2 = isRight ? v3(cam.p5) : v3(cam.p2);
    vec3 right = p1 - p0, up = p2 - p0;
    vec3 pos = p0 + c.x * right + c.y * up;
    vec3 dir = normalize(pos - v3(cam.eye));
    return {v3(cam.eye), dir};
}
// camera.glsl:113-138 (+ :89-111)
Ray calculatePrimary(const CameraRec& cam, uint32_t lx, uint32_t ly, vec2 randLoc) {
    vec2 loc = {(float)lx, (float)ly};
    vec2 c = (loc + rand2(loc + randLoc)) * vec2{cam.invRes[0], cam.invRes[1]};
    c.y = 1.0f - c.y;
    switch (cam.projectionType) {
        case 1: return calculateOmni(cam, c, false);
        case 2: return calculateOmni(cam, {c.x, fractf(c.y * 2.0f)}, c.y < 0.5f);
        case 4: return calculateOmni(cam, {fractf(c.x * 2.0f), c.y}, c.x < 0.5f);
        case 3:
            if (c.y < 0.5f) return calculateScreen(cam, {c.x, c.y * 2.0f}, false);
            return calculateScreen(cam, {c.x, c.y * 2.0f - 1.0f}, true);
        case 5:
            if (c.x < 0.5f) return calculateScreen(cam, {c.x * 2.0f, c.y}, false);
            return calculateScreen(cam, {c.x * 2.0f - 1.0f, c.y}, true);
        default: return calculateScreen(cam, c, false);
    }
}

// ------------------------------------------------------------------------------------------------
// device-side: skybox (SH/scene.glsl:56-69) — D7
// ------------------------------------------------------------------------------------------------
inline vec3 skyTexel(const Scene& sc, int64_t x, int64_t y) {
    if (x < 0 || y < 0 || x >= (int64_t)sc.skyW || y >= (int64_t)sc.skyH) return {0, 0, 0};
    const uint16_t* p = sc.sky + 4 * ((size_t)y * sc.skyW + (size_t)x);
    return {f16_to_f32(p[0]), f16_to_f32(p[1]), f16_to_f32(p[2])};
}
vec3 sampleSkybox(const Scene& sc, const CameraRec& cam, vec3 dir) {
    if (sc.skyW == 0 || sc.skyH == 0) return v3(cam.skyboxColor);
    vec2 uv = vec2{cr_atan2(dir.x, dir.z), cr_asin(dir.y * -1.0f)} * vec2{0.1591f, 0.3183f} + 0.5f;
    float fx = uv.x * (float)sc.skyW - 0.5f, fy = uv.y * (float)sc.skyH - 0.5f;
    float x0f = std::floor(fx), y0f = std::floor(fy);
    float ax = fx - x0f, ay = fy - y0f;
    if (std::isnan(fx) || std::isnan(fy)) return {ubits(0x7FC00000u), ubits(0x7FC00000u), ubits(0x7FC00000u)};
    int64_t x0 = (int64_t)x0f, y0 = (int64_t)y0f;
    vec3 t00 = skyTexel(sc, x0, y0), t10 = skyTexel(sc, x0 + 1, y0), t01 = skyTexel(sc, x0, y0 + 1), t11 = skyTexel(sc, x0 + 1, y0 + 1);
    vec3 top = t00 * (1.0f - ax) + t10 * ax;
    vec3 bot = t01 * (1.0f - ax) + t11 * ax;
    return top * (1.0f - ay) + bot * ay;
}

// ------------------------------------------------------------------------------------------------
// device-side: lights and shading (SH/light.glsl)
// ------------------------------------------------------------------------------------------------
const float minRoughness = 0.01f, specularEpsilon = 0.001f;   // light.glsl:6-7

inline float smoothstepf(float e0, float e1, float x) {   // D9
    float t = (x - e0) / (e1 - e0);
    t = glsl_min(glsl_max(t, 0.0f), 1.0f);
    return t * t * (3.0f - 2.0f * t);
}

// light.glsl:98-133
vec3 getDirToLight(const LightRec& light, vec3 pos, float& brightness, float& dist, vec2 random) {
    vec3 l;
    brightness = 1.0f;
    dist = -1.0f;
    uint32_t ro; std::memcpy(&ro, &light.rad, 4);
    vec2 radOrigin = unpackHalf2x16(ro);
    radOrigin = {glsl_max(radOrigin.x, 0.0f), glsl_max(radOrigin.y, 0.0f)};
    radOrigin.y = glsl_min(radOrigin.y, radOrigin.x);
    if (light.type == 2) {   // LightType_Point
        l = pos - v3(light.pos);
        dist = length(l);
        vec3 p = randomPointOnHemisphere(random, v3(light.pos), radOrigin.y, normalize(l));
        l = pos - p;
        float r = radOrigin.x - radOrigin.y;
        float d = glsl_max(dist - radOrigin.y, 0.0f);
        brightness = cr_pow(smoothstepf(r, 0.0f, d), ubits(light.dir[0]));
    } else
        l = getSunDirection(random, normalize(decodeNormal(light.dir[0], light.dir[1])), radOrigin.x);
    return normalize(l);
}

// light.glsl:22-55
float ndfGGX(vec3 n, vec3 h, float roughness) {
    float alpha = roughness * roughness;
    float a2 = alpha * alpha;
    float NdotH = glsl_max(dot(n, h), 0.0f);
    float denom = NdotH * NdotH * (a2 - 1.0f) + 1.0f;
    denom *= denom * pi;
    if (denom == 0.0f) return 0.0f;
    return a2 / denom;
}
inline float geomSchlickGGX(float NdotV, float k) { return NdotV / (NdotV * (1.0f - k) + k); }
inline float geomSmith(float NdotV, float NdotL, float k) { return geomSchlickGGX(NdotV, k) * geomSchlickGGX(NdotL, k); }
inline float pow5(float f) { float f2 = f * f; return f2 * f2 * f; }
inline vec3 fresnelSchlick(vec3 F0, vec3 h, vec3 v) { return F0 + (1.0f - F0) * pow5(1.0f - glsl_max(dot(h, v), 0.0f)); }

// light.glsl:64-94
vec3 cookTorrance(vec3 F0, vec3 albedo, const LightRec& light, vec3 n, vec3 l, vec3 v, float NdotV,
                  float invSquareDist, float roughness, float metallic, float k, float NdotL) {
    vec3 h = normalize(l + v);
    float D = ndfGGX(n, h, glsl_max(roughness, minRoughness));
    float G = geomSmith(NdotV, NdotL, k);
    vec3 F = fresnelSchlick(F0, h, v);
    float denom = 4.0f * NdotL * NdotV + specularEpsilon;
    vec3 num = D * G * F;
    vec3 kS = num / denom;
    vec3 kD = (1.0f - F) * (1.0f - metallic);
    vec3 color = kD * albedo + kS;
    uint32_t cx, cy; std::memcpy(&cx, &light.r, 4); std::memcpy(&cy, &light.b, 4);
    return color * unpackColor3(cx, cy) * invSquareDist * NdotL;
}

// light.glsl:135-159
vec3 shadeLight(vec3 F0, vec3 albedo, float roughness, float metallic, const LightRec& light, vec3 pos,
                vec3 n, vec3 v, float NdotV, vec2 random) {
    float brightness, dst;
    vec3 l = getDirToLight(light, pos, brightness, dst, random);
    float k = roughness + 1.0f;
    k *= k / 8.0f;
    float NdotL = glsl_max(dot(n, l), 0.0f);
    return cookTorrance(F0, albedo, light, n, l, v, NdotV, brightness, roughness, metallic, k, NdotL);
}

struct MatU { vec3 albedo, ambient, emissive; float metallic, roughness; };
MatU unpackMaterial(const MaterialRec& m) {
    uint32_t w[6]; std::memcpy(w, &m, 24);
    MatU o;
    o.albedo = unpackColor3(w[0], w[1]); o.metallic = unpackColorAUnorm(w[1]);
    o.ambient = unpackColor3(w[2], w[3]); o.roughness = unpackColorAUnorm(w[3]);
    o.emissive = unpackColor3(w[4], w[5]);
    return o;
}

// light.glsl:58-60
inline vec3 fresnelSchlickRoughness(vec3 F0, float NdotV, float roughness) {
    vec3 r = {1.0f - roughness, 1.0f - roughness, 1.0f - roughness};
    return F0 + (vmax(F0, r) - F0) * cr_pow(1.0f - NdotV, 5.0f);
}
// light.glsl:163-187
vec3 shade(const MaterialRec& mr, vec3 /*pos*/, vec3 /*n*/, vec3 /*v*/, float NdotV, vec3 light, vec3 reflected) {
    MatU m = unpackMaterial(mr);
    vec3 F0 = mix(vec3{0.04f, 0.04f, 0.04f}, m.albedo, m.metallic);
    vec3 kS = fresnelSchlickRoughness(F0, NdotV, m.roughness);
    vec3 kD = (1.0f - kS) * (1.0f - m.metallic);
    return (m.ambient + kD / pi) * m.albedo + kS * reflected + light + m.emissive;
}

// light.glsl:221-229 with the NV constants of SH/light_rt.glsl:21-24 / SH/nv_all.shadow.comp:20-21
inline uint32_t indexToLightNV(uint32_t lx, uint32_t ly, uint32_t w, uint32_t h, uint32_t sample) {
    uint32_t tileX = lx >> 4, tileY = ly >> 1;
    uint32_t tilesX = (w >> 4) + ((w & 15u) != 0u), tilesY = (h >> 1) + ((h & 1u) != 0u);
    return tileX + tileY * tilesX + sample * tilesY * tilesX;
}
inline uint32_t shadowWords(uint32_t w, uint32_t h, uint32_t samples) {
    return (((w >> 4) + ((w & 15u) != 0u)) * ((h >> 1) + ((h & 1u) != 0u))) * samples;
}

// ------------------------------------------------------------------------------------------------
// per-pixel bodies of K1..K4
// ------------------------------------------------------------------------------------------------

// SH/raygen.comp:16-53
void raygenPixel(const Scene& sc, const CameraRec& cam, const SeedRec& seed, uint32_t x, uint32_t y,
                 float dirT[4], float uvN[4], float* rayOut, uint8_t* flagOut) {
    Ray ray = calculatePrimary(cam, x, y, {seed.randomX, seed.randomY});
    Hit hit = traceGeometry(sc, ray, noRayHit);
    vec3 d = hit.rayDir * hit.hitT;
    dirT[0] = d.x; dirT[1] = d.y; dirT[2] = d.z; dirT[3] = ubits(hit.object);
    if (hit.hitT == noHit) { dirT[0] = hit.rayDir.x; dirT[1] = hit.rayDir.y; dirT[2] = hit.rayDir.z; dirT[3] = ubits(noRayHit); }
    uint32_t en[2];
    encodeNormalGpu(hit.objectNormal, en);
    if (g_mode == ORC_MODE_DEBUG || hit.hitT < noHit) {   // raygen.comp:46-51 (D10)
        uvN[0] = hit.uv.x; uvN[1] = hit.uv.y; uvN[2] = ubits(en[0]); uvN[3] = ubits(en[1]);
    }
    if (rayOut) { rayOut[0] = ray.pos.x; rayOut[1] = ray.pos.y; rayOut[2] = ray.pos.z; rayOut[3] = ray.dir.x; rayOut[4] = ray.dir.y; rayOut[5] = ray.dir.z; }
    if (flagOut) *flagOut = rayFlags(sc, ray, noRayHit, hit.hitT);
}

// SH/nv_all.shadow.comp:50-143, one (pixel, sample); returns the lane's `hit`
bool shadowPixel(const Scene& sc, const CameraRec& cam, const SeedRec& seed, uint32_t samples, uint32_t x, uint32_t y,
                 uint32_t i, const float dirT[4], float* rayOut, uint32_t lightId = 0) {
    uint32_t object = fbits(dirT[3]);
    vec3 hitPos = v3(cam.eye) + vec3{dirT[0], dirT[1], dirT[2]};
    vec2 loc = {(float)x, (float)y};
    vec2 uv = (loc + rand2(loc + vec2{seed.randomX, seed.randomY})) / 128.0f;
    uv = uv + hammersley(i, samples);
    vec2 random = rand2(uv);
    const LightRec& light = sc.light[lightId];   // lightId = 0 in the reference, nv_all.shadow.comp:97
    float brightness, dist;
    vec3 l = getDirToLight(light, hitPos, brightness, dist, random);
    Ray ray = {hitPos, -l};
    if (rayOut) { rayOut[0] = ray.pos.x; rayOut[1] = ray.pos.y; rayOut[2] = ray.pos.z; rayOut[3] = ray.dir.x; rayOut[4] = ray.dir.y; rayOut[5] = ray.dir.z; }
    bool hit = false;
    if (object != noRayHit) {
        if (dist >= 0.0f) {
            uint32_t ro; std::memcpy(&ro, &light.rad, 4);
            vec2 radOrigin = unpackHalf2x16(ro);
            if (dist >= radOrigin.y && dist < radOrigin.x) hit = traceOcclusion(sc, ray, dist - radOrigin.y, object);
        } else
            hit = traceOcclusion(sc, ray, noHit, object);
    }
    return hit;
}

// SH/nv_all.lighting.comp:33-103; returns false for a miss (DEBUG stores zero, D10)
bool lightingPixel(const Scene& sc, const CameraRec& cam, uint32_t samples, uint32_t x, uint32_t y, const float dirT[4],
                   const float uvN[4], const uint32_t* bits, uint32_t w, uint32_t h, vec3& out) {
    uint32_t object = fbits(dirT[3]);
    out = {0, 0, 0};
    if (object == noRayHit) return false;
    vec3 dxyz = {dirT[0], dirT[1], dirT[2]};
    vec3 hitPos = v3(cam.eye) + dxyz;
    MatU m = unpackMaterial(sc.mat[sc.matIdx[object]]);
    vec3 F0 = mix(vec3{0.04f, 0.04f, 0.04f}, m.albedo, m.metallic);
    vec3 n = decodeNormal(fbits(uvN[2]), fbits(uvN[3]));
    vec3 v = normalize(dxyz);
    float NdotV = glsl_max(dot(v, -n), 0.0f);
    vec3 light = {0, 0, 0};
    vec2 loc = {(float)x, (float)y};
    vec2 uv = (loc + rand2(loc + vec2{0.0f, 0.0f})) / 128.0f;   // D8: unbound seed reads as zero
    const uint32_t nl = g_all_lights ? sc.lightCount : 1u;
    for (uint32_t L = 0; L < nl; ++L)
        for (uint32_t i = 0; i < samples; ++i) {
            vec2 uvi = uv + hammersley(i, samples);
            vec2 random = rand2(uvi);
            uint32_t word = bits[indexToLightNV(x, y, w, h, L * samples + i)];
            uint32_t bit = (x & 15u) | ((y & 1u) << 4);
            if (!(word & (1u << bit)))
                light = light + shadeLight(F0, m.albedo, m.roughness, m.metallic, sc.light[L], hitPos, n, v, NdotV, random);
        }
    // the reference samples light 0 only and scales by the light count (nv_all.lighting.comp:98); with every light evaluated the sum stands
    out = g_all_lights ? light / (float)samples : light / (float)samples * (float)sc.lightCount;
    return true;
}

inline uint32_t unorm8(float c) {   // D6
    if (std::isnan(c)) return 0;
    c = c < 0.0f ? 0.0f : (c > 1.0f ? 1.0f : c);
    return (uint32_t)std::floor(c * 255.0f + 0.5f);
}

// SH/composite.comp:52-285 (RELEASE colour path; UI blend is out of scope: USE_UI must be clear)
uint32_t compositePixel(const Scene& sc, const CameraRec& cam, const SeedRec& seed, const float dirT[4], const float uvN[4],
                        const uint16_t lighting[4], float* accum) {
    uint32_t object = fbits(dirT[3]);
    vec3 dxyz = {dirT[0], dirT[1], dirT[2]};
    vec3 n = decodeNormal(fbits(uvN[2]), fbits(uvN[3]));
    vec3 rayDir = normalize(dxyz);
    float hitT = length(dxyz);
    if (object == noRayHit) hitT = noHit;
    vec3 eye = v3(cam.eye);
    vec3 hitPos = eye + dxyz;
    vec3 light = {f16_to_f32(lighting[0]), f16_to_f32(lighting[1]), f16_to_f32(lighting[2])};
    vec3 color;
    // light.glsl:205-219 shadeHitFinalRecursion
    if (hitT == noHit)
        color = sampleSkybox(sc, cam, rayDir);
    else {
        vec3 v = rayDir;
        float NdotV = glsl_max(dot(v, -n), 0.0f);
        vec3 reflected = sampleSkybox(sc, cam, reflect(v, n));
        color = shade(sc.mat[sc.matIdx[object]], hitPos, n, v, NdotV, light, reflected);
    }
    color = mix(color, vec3{0, 0, 0}, 0.0f);   // composite.comp:93-97: cloud = vec4(0)
    if (g_mode == ORC_MODE_DEBUG && (std::isnan(color.x) || std::isnan(color.y) || std::isnan(color.z)))
        color = {0.0f, 0.0f, 10000.0f};        // composite.comp:236-239 (DEBUG builds; nanOnly = 0)
    if (cam.flags & 2u) {                      // composite.comp:249-257
        if (seed.sampleCount > 1) color = color + vec3{accum[0], accum[1], accum[2]};
        accum[0] = color.x; accum[1] = color.y; accum[2] = color.z; accum[3] = 0.0f;
        color = color / (float)seed.sampleCount;
    }
    vec3 e = -color * cam.exposure;
    color = vmax(vec3{1.0f, 1.0f, 1.0f} - vec3{cr_exp(e.x), cr_exp(e.y), cr_exp(e.z)}, vec3{0, 0, 0});
    return unorm8(color.x) | (unorm8(color.y) << 8) | (unorm8(color.z) << 16) | (255u << 24);
}

int g_threads = 0;

int threadCount() {
    if (g_threads > 0) return g_threads;
    unsigned h = std::thread::hardware_concurrency();
    return h ? (int)h : 1;
}

// dynamic-chunk parallel loop on std::thread (no OpenMP dependency)
template <class F>
void parallelFor(int64_t n, int64_t chunk, F&& body) {
    const int nt = threadCount();
    if (nt <= 1 || n <= chunk) { for (int64_t i = 0; i < n; ++i) body(i); return; }
    std::atomic<int64_t> next{0};
    auto worker = [&]() {
        for (;;) {
            int64_t b = next.fetch_add(chunk);
            if (b >= n) break;
            int64_t e = std::min(n, b + chunk);
            for (int64_t i = b; i < e; ++i) body(i);
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(worker);
    worker();
    for (auto& t : pool) t.join();
}

}  // namespace

// =================================================================================================
// C interface
// =================================================================================================
extern "C" {

void orc_set_threads(int n) { g_threads = n; }
void orc_set_all_lights(int on) { g_all_lights = on ? 1 : 0; }
void orc_set_history(float alpha) { g_history_alpha = alpha; }
// the History blend (ours): history and lighting are rgba16f; the blend is evaluated in binary32 and stored round-to-nearest-even;
// alpha of the first frame of a sequence is 1 (history = lighting), which the caller expresses by passing first != 0
void orc_history_blend(uint16_t* history_f16, uint16_t* lighting_f16, uint64_t texels, int first) {
    const float a = first ? 1.0f : g_history_alpha;
    for (uint64_t i = 0; i < texels * 4; ++i) {
        const float hv = f16_to_f32(history_f16[i]), lv = f16_to_f32(lighting_f16[i]);
        const uint16_t o = f32_to_f16_rtne(hv * (1.0f - a) + lv * a);
        history_f16[i] = o; lighting_f16[i] = o;   // composite reads the blended texture
    }
}
void orc_set_mode(int mode) { g_mode = mode == ORC_MODE_RELEASE ? ORC_MODE_RELEASE : ORC_MODE_DEBUG; }
int orc_get_mode(void) { return g_mode; }
int orc_get_threads(void) { return threadCount(); }

uint16_t orc_f16_trunc(float v) { return f16_trunc(v); }
float orc_f16_to_f32(uint16_t h) { return f16_to_f32(h); }
uint16_t orc_f32_to_f16_rtne(float v) { return f32_to_f16_rtne(v); }
void orc_spheremap(const float n[3], uint16_t out[2]) { spheremapTransform(out[0], out[1], v3(n)); }
void orc_encode_normal_cpu(const float n[3], uint32_t out[2]) { encodeNormalCpu(v3(n), out); }

// scene_object_types.hpp:98-106
void orc_triangle_flat(const float p[9], void* out48) {
    TriangleRec t;
    std::memset(&t, 0, sizeof t);
    std::memcpy(t.p0, p, 12); std::memcpy(t.p1, p + 3, 12); std::memcpy(t.p2, p + 6, 12);
    vec3 n = cross(host_normalize3(v3(p + 3) - v3(p)), host_normalize3(v3(p + 6) - v3(p)));
    spheremapTransform(t.n0[0], t.n0[1], n);
    spheremapTransform(t.n1[0], t.n1[1], n);
    spheremapTransform(t.n2[0], t.n2[1], n);
    std::memcpy(out48, &t, 48);
}
// scene_object_types.hpp:87-96
void orc_triangle_normals(const float p[9], const float n[9], void* out48) {
    TriangleRec t;
    std::memset(&t, 0, sizeof t);
    std::memcpy(t.p0, p, 12); std::memcpy(t.p1, p + 3, 12); std::memcpy(t.p2, p + 6, 12);
    spheremapTransform(t.n0[0], t.n0[1], v3(n));
    spheremapTransform(t.n1[0], t.n1[1], v3(n + 3));
    spheremapTransform(t.n2[0], t.n2[1], v3(n + 6));
    std::memcpy(out48, &t, 48);
}
// scene_object_types.hpp:166-171
void orc_light_directional(const float dir[3], const float color[3], float angular_extent, void* out32) {
    LightRec l;
    std::memset(&l, 0, sizeof l);
    l.rad = f16_trunc(angular_extent);
    encodeNormalCpu(v3(dir), l.dir);
    l.r = f16_trunc(color[0]); l.g = f16_trunc(color[1]); l.b = f16_trunc(color[2]);
    l.type = 0;
    std::memcpy(out32, &l, 32);
}
// scene_object_types.hpp:173-179
void orc_light_point(const float pos[3], const float color[3], float rad, float origin, float specularity, void* out32) {
    LightRec l;
    std::memset(&l, 0, sizeof l);
    std::memcpy(l.pos, pos, 12);
    l.rad = f16_trunc(rad); l.origin = f16_trunc(origin);
    l.dir[0] = fbits(specularity); l.dir[1] = 0;
    l.r = f16_trunc(color[0]); l.g = f16_trunc(color[1]); l.b = f16_trunc(color[2]);
    l.type = 2;
    std::memcpy(out32, &l, 32);
}
// scene_object_types.hpp:267-288
void orc_material(const float albedo[3], const float ambient[3], const float emission[3], float metallic,
                  float roughness, float transparency, void* out32) {
    MaterialRec m;
    std::memset(&m, 0, sizeof m);
    for (int i = 0; i < 3; ++i) { m.albedo[i] = f16_trunc(albedo[i]); m.ambient[i] = f16_trunc(ambient[i]); m.emission[i] = f16_trunc(emission[i]); }
    m.metallic = (uint16_t)(metallic * 65535.0f);
    m.roughness = (uint16_t)(roughness * 65535.0f);
    m.transparency = transparency;
    std::memcpy(out32, &m, 32);
}

// src/rt/structs.cpp:5-42 (getRot/getView), src/rt/raytracing_interface.cpp:96-107 (resize) and :279-328 (update)
void orc_camera(const float eye[3], float pitch, float yaw, float roll, float left_fov, float right_fov, float ipd,
                uint32_t projection, uint32_t width, uint32_t height, uint32_t flags, float exposure,
                const float skybox_color[3], void* out144) {
    CameraRec c;
    std::memset(&c, 0, sizeof c);
    std::memcpy(c.eye, eye, 12);
    c.width = width; c.height = height;
    c.ipd = ipd; c.projectionType = projection;
    std::memcpy(c.skyboxColor, skybox_color, 12);
    c.exposure = exposure; c.focalDistance = 10.0f; c.aperature = 0.1f; c.flags = flags;
    c.invRes[0] = 1.0f / (float)width; c.invRes[1] = 1.0f / (float)height;
    c.tiles[0] = width / 16; c.tiles[1] = height / 16;

    const float a = roll, b = yaw, g = pitch;
    const float ca = std::cos(a), cb = std::cos(b), cg = std::cos(g), sa = std::sin(a), sb = std::sin(b), sg = std::sin(g);
    const float sbsg = sb * sg, sbcg = sb * cg;
    const vec3 ax = {ca * cb, ca * sbsg - sa * cg, ca * sbcg + sa * sg};
    const vec3 ay = {sa * cb, sa * sbsg + ca * cg, sa * sbcg - ca * sg};
    const vec3 az = {-sb, cb * sg, cb * cg};
    auto viewPos = [&](float eyeOffset) { return v3(eye) + ax * (ipd * 5e-4f * eyeOffset); };
    // CORE2/include/types/mat.hpp:231-240: res[j] += m[i][j] * v[i], i outer, starting from zero
    auto xform = [&](vec3 pos, float x, float y, float z) {
        vec3 r = {0, 0, 0};
        r = r + ax * x; r = r + ay * y; r = r + az * z; r = r + pos * 1.0f;
        return r;
    };
    const bool isStereo = projection == 3 || projection == 5;
    if (projection != 1 && projection != 2 && projection != 4) {
        float rx = (float)width, ry = (float)height;
        if (isStereo) { if (projection == 5) rx /= 2; else ry /= 2; }
        const float aspect = rx / ry;
        const double halfDeg = 0.5 * (3.141592653589793 / 180);   // 0.5_deg, CORE2 units.hpp:53 / data_types.hpp:94-96
        const float nearL = (float)std::tan(left_fov * halfDeg);
        vec3 posL = viewPos(isStereo ? -1.0f : 0.0f);
        vec3 p0 = xform(posL, -aspect, 1, -nearL), p1 = xform(posL, aspect, 1, -nearL), p2 = xform(posL, -aspect, -1, -nearL);
        std::memcpy(c.p0, &p0, 12); std::memcpy(c.p1, &p1, 12); std::memcpy(c.p2, &p2, 12);
        if (isStereo) {
            const float nearR = (float)std::tan(right_fov * halfDeg);
            vec3 posR = viewPos(1.0f);
            vec3 p3 = xform(posR, -aspect, 1, -nearR), p4 = xform(posR, aspect, 1, -nearR), p5 = xform(posR, -aspect, -1, -nearR);
            std::memcpy(c.p3, &p3, 12); std::memcpy(c.p4, &p4, 12); std::memcpy(c.p5, &p5, 12);
        }
    }
    std::memcpy(out144, &c, 144);
}

// test/scene/niels_scene.cpp:5-69 after SceneGraph::update (igx/src/helpers/scene_graph.cpp:267-323, 378-522):
// per-type arrays in insertion order, lights sorted directional < spot < point, material indices by global object id.
void orc_niels_scene(double time, void* tris, void* spheres, void* cubes, void* planes, void* lights, void* materials,
                     uint32_t* material_indices, uint32_t info[9]) {
    const float mats[8][11] = {
        {1, 0.5f, 1, 0.05f, 0.01f, 0.05f, 0, 0, 0, 0, 1},   {0, 1, 0, 0, 0.05f, 0, 0, 0, 0, 0, 1},
        {0, 0, 1, 0, 0, 0.05f, 0, 0, 0, 0, 1},               {1, 0, 1, 0.05f, 0, 0.05f, 0, 0, 0, 0, 1},
        {1, 1, 0, 0.05f, 0.05f, 0, 0, 0, 0, 0, 1},           {0, 1, 1, 0, 0.05f, 0.05f, 0, 0, 0, 0, 1},
        {0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0},                   {0, 0, 0, 0, 0, 0, 0, 0, 0, .25f, .5f}};
    for (int i = 0; i < 8; ++i) orc_material(mats[i], mats[i] + 3, mats[i] + 6, mats[i][9], mats[i][10], 1.0f, (char*)materials + 32 * i);
    const float plane[4] = {0, 1, 0, 0};
    std::memcpy(planes, plane, 16);
    const float cubesv[2][6] = {{0, 0, 0, 1, 1, 1}, {-2, 0, -2, -1, 1, -1}};
    std::memcpy(cubes, cubesv, 48);
    const float tp[3][9] = {{1, 1, 0, -1, 1, 0, 1, 0, 1}, {-1, 4, 0, 1, 4, 0, 1, 3, 1}, {-1, 7, 0, 1, 7, 0, 1, 5, 1}};
    for (int i = 0; i < 3; ++i) orc_triangle_flat(tp[i], (char*)tris + 48 * i);
    const float t = (float)std::sin(time), c = (float)std::cos(time);
    const float sph[7][4] = {{0, 1, 5, 1},  {0, 1, -5, 1}, {3, 1, 0, 1}, {0, 6, 0, 1},
                             {7, 2 + (float)std::sin(0.0), 0, 1}, {-5 + t, 2 + c, 0, 1}, {t, 3 + c, 0, 1}};
    std::memcpy(spheres, sph, sizeof sph);
    const vec3 sd = host_normalize3({-0.5f, -2, -1});
    const float sunDir[3] = {sd.x, sd.y, sd.z}, sunCol[3] = {0.9f, 0.9f, 0.9f};
    orc_light_directional(sunDir, sunCol, (float)(0.533L * (3.141592653589793 / 180)), lights);
    const float p1[3] = {0, 0.1f, 0}, c1[3] = {1, 0, 0}, p2[3] = {2, 2, 2}, c2[3] = {0, 1, 1};
    orc_light_point(p1, c1, 5, 0.3f, 1, (char*)lights + 32);
    orc_light_point(p2, c2, 7, 0.6f, 1, (char*)lights + 64);
    const uint32_t mi[13] = {3, 4, 5, 0, 1, 2, 3, 4, 0, 7, 1, 2, 0};
    std::memcpy(material_indices, mi, sizeof mi);
    const uint32_t inf[9] = {3, 8, 3, 7, 2, 1, 1, 0, 2};
    std::memcpy(info, inf, sizeof inf);
}

// Radiance RGBE reader (what stb_image's stbi__hdr_load does for req_comp = 0: RGB float, value =
// mantissa * 2^(e - 136), e == 0 -> 0) followed by igxi convert.cpp:59-78,209-231,299-323: rgb copied into a
// zero-initialised rgba16f buffer with the truncating f16 conversion, inf/NaN replaced by 65504 (0x7bff).
int orc_load_hdr(const char* path, uint16_t* out, uint32_t* w, uint32_t* h) {
    FILE* f = std::fopen(path, "rb");
    if (!f) return 1;
    std::fseek(f, 0, SEEK_END);
    long size = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    std::vector<uint8_t> buf((size_t)size);
    if (std::fread(buf.data(), 1, buf.size(), f) != buf.size()) { std::fclose(f); return 2; }
    std::fclose(f);
    size_t pos = 0;
    auto readLine = [&]() { std::string s; while (pos < buf.size() && buf[pos] != '\n') s.push_back((char)buf[pos++]); ++pos; return s; };
    std::string first = readLine();
    if (first != "#?RADIANCE" && first != "#?RGBE") return 3;
    bool okFormat = false;
    for (;;) { std::string l = readLine(); if (l.empty()) break; if (l == "FORMAT=32-bit_rle_rgbe") okFormat = true; if (pos >= buf.size()) return 4; }
    if (!okFormat) return 5;
    std::string res = readLine();
    unsigned H = 0, W = 0;
    if (std::sscanf(res.c_str(), "-Y %u +X %u", &H, &W) != 2) return 6;
    *w = W; *h = H;
    if (!out) return 0;
    std::vector<uint8_t> scan((size_t)W * 4);
    auto convert = [&](const uint8_t* rgbe, uint16_t* dst) {
        float v[3] = {0, 0, 0};
        if (rgbe[3] != 0) { float f1 = std::ldexp(1.0f, (int)rgbe[3] - 136); for (int c = 0; c < 3; ++c) v[c] = (float)rgbe[c] * f1; }
        for (int c = 0; c < 3; ++c) {
            uint16_t hv = f16_trunc(v[c]);
            if (((hv >> 10) & 0x1F) == 0x1F) hv = 0x7BFF | 0x03FF;   // flp::max(): ((mask<<mant)-1)|mantMask = 0x7bff
            dst[c] = hv;
        }
        dst[3] = 0;
    };
    for (unsigned y = 0; y < H; ++y) {
        uint16_t* row = out + (size_t)y * W * 4;
        bool rle = false;
        if (W >= 8 && W < 32768 && pos + 4 <= buf.size() && buf[pos] == 2 && buf[pos + 1] == 2 && !(buf[pos + 2] & 0x80)) {
            unsigned len = ((unsigned)buf[pos + 2] << 8) | buf[pos + 3];
            if (len == W) rle = true;
        }
        if (!rle) {   // flat scanline
            if (pos + (size_t)W * 4 > buf.size()) return 7;
            for (unsigned x = 0; x < W; ++x) convert(&buf[pos + 4 * (size_t)x], row + 4 * (size_t)x);
            pos += (size_t)W * 4;
            continue;
        }
        pos += 4;
        for (int k = 0; k < 4; ++k) {
            unsigned i = 0;
            while (i < W) {
                if (pos >= buf.size()) return 8;
                unsigned count = buf[pos++];
                if (count > 128) {
                    count -= 128;
                    if (pos >= buf.size() || i + count > W) return 9;
                    uint8_t value = buf[pos++];
                    for (unsigned z = 0; z < count; ++z) scan[(size_t)(i++) * 4 + k] = value;
                } else {
                    if (count == 0 || pos + count > buf.size() || i + count > W) return 10;
                    for (unsigned z = 0; z < count; ++z) scan[(size_t)(i++) * 4 + k] = buf[pos++];
                }
            }
        }
        for (unsigned x = 0; x < W; ++x) convert(&scan[4 * (size_t)x], row + 4 * (size_t)x);
    }
    return 0;
}

// ---- the synthetic scenes of BASELINE.json configs[2] / configs[3] (no reference counterpart).  The oracle carries its own
// statement of the generators so that the CPU legs of bench.py need nothing from the product library; tests/test_cpu_host.py
// holds them byte-equal to rtb_gen_soup / rtb_gen_heightfield. ----
static inline uint64_t gen_mix64(uint64_t z) {   // splitmix64 finaliser
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline float gen_u01(uint64_t seed, uint64_t k) { return (float)(gen_mix64(seed ^ gen_mix64(k)) >> 40) * (1.0f / 16777216.0f); }

// n flat triangles: centre uniform in [-10,10]^3, three vertex offsets uniform in [-0.05,0.05]^3
void orc_gen_soup(uint64_t n, uint64_t seed, void* out48) {
    parallelFor((int64_t)n, 4096, [&](int64_t i) {
        const uint64_t k = (uint64_t)i * 12;
        const float c[3] = {gen_u01(seed, k) * 20.0f - 10.0f, gen_u01(seed, k + 1) * 20.0f - 10.0f, gen_u01(seed, k + 2) * 20.0f - 10.0f};
        float p[9];
        for (int v = 0; v < 3; ++v)
            for (int a = 0; a < 3; ++a) p[3 * v + a] = c[a] + (gen_u01(seed, k + 3 + 3 * v + a) * 0.1f - 0.05f);
        orc_triangle_flat(p, (char*)out48 + 48 * (size_t)i);
    });
}

// (grid x grid) quads over [-10,10]^2, height = four octaves of lattice value noise, smooth vertex normals (config 4)
void orc_gen_heightfield(uint32_t grid, uint64_t seed, void* out48) {
    const uint32_t V = grid + 1;
    const float step = 20.0f / (float)grid;
    std::vector<float> h((size_t)V * V);
    auto heightAt = [&](float x, float z) {
        float hh = 0.0f, amp = 1.0f, freq = 0.15f;
        for (int o = 0; o < 4; ++o) {
            const float fx = x * freq, fz = z * freq;
            const float x0 = std::floor(fx), z0 = std::floor(fz);
            const float tx = fx - x0, tz = fz - z0;
            auto lat = [&](float ix, float iz) {
                const uint64_t k = ((uint64_t)(int64_t)ix * 0x1F1F1F1Full) ^ ((uint64_t)(int64_t)iz * 0x3D4D51CBull) ^ ((uint64_t)o << 56);
                return gen_u01(seed, k) * 2.0f - 1.0f;
            };
            auto fade = [](float t) { return t * t * t * (t * (t * 6.0f - 15.0f) + 10.0f); };
            const float sx = fade(tx), sz = fade(tz);
            const float a = lat(x0, z0), b = lat(x0 + 1, z0), c = lat(x0, z0 + 1), d = lat(x0 + 1, z0 + 1);
            hh += amp * ((a * (1 - sx) + b * sx) * (1 - sz) + (c * (1 - sx) + d * sx) * sz);
            amp *= 0.5f; freq *= 2.0f;
        }
        return hh * 1.5f;
    };
    parallelFor((int64_t)V * V, 4096, [&](int64_t i) { const uint32_t ix = (uint32_t)(i % V), iz = (uint32_t)(i / V); h[(size_t)i] = heightAt(-10.0f + ix * step, -10.0f + iz * step); });
    auto P = [&](uint32_t ix, uint32_t iz, float* o) { o[0] = -10.0f + ix * step; o[1] = h[(size_t)iz * V + ix]; o[2] = -10.0f + iz * step; };
    auto N = [&](uint32_t ix, uint32_t iz, float* o) {   // central differences, pointing up
        const uint32_t xm = ix ? ix - 1 : ix, xp = ix + 1 < V ? ix + 1 : ix, zm = iz ? iz - 1 : iz, zp = iz + 1 < V ? iz + 1 : iz;
        const float dx = (h[(size_t)iz * V + xp] - h[(size_t)iz * V + xm]) / ((float)(xp - xm) * step);
        const float dz = (h[(size_t)zp * V + ix] - h[(size_t)zm * V + ix]) / ((float)(zp - zm) * step);
        const vec3 n = host_normalize3({-dx, 1.0f, -dz});
        o[0] = n.x; o[1] = n.y; o[2] = n.z;
    };
    parallelFor((int64_t)grid * grid, 2048, [&](int64_t q) {
        const uint32_t ix = (uint32_t)(q % grid), iz = (uint32_t)(q / grid);
        float p[9], n[9];
        P(ix, iz, p); P(ix, iz + 1, p + 3); P(ix + 1, iz, p + 6); N(ix, iz, n); N(ix, iz + 1, n + 3); N(ix + 1, iz, n + 6);
        orc_triangle_normals(p, n, (char*)out48 + 48 * (size_t)(2 * q));
        P(ix + 1, iz, p); P(ix, iz + 1, p + 3); P(ix + 1, iz + 1, p + 6); N(ix + 1, iz, n); N(ix, iz + 1, n + 3); N(ix + 1, iz + 1, n + 6);
        orc_triangle_normals(p, n, (char*)out48 + 48 * (size_t)(2 * q + 1));
    });
}

// SH/init.comp:12-18
void orc_init_pass(void* seed24) {
    SeedRec s;
    std::memcpy(&s, seed24, 24);
    vec2 off = rand2(vec2{s.cpuOffsetX, s.cpuOffsetY} + (float)s.sampleCount);
    s.randomX = off.x; s.randomY = off.y;
    ++s.sampleCount; ++s.sampleOffset;
    std::memcpy(seed24, &s, 24);
}

void orc_raygen(const orc_scene* s, const void* cam144, const void* seed24, float* dirT, float* uvN, float* rays_out,
                uint8_t* flags_out) {
    Scene sc = view(s);
    CameraRec cam; std::memcpy(&cam, cam144, 144);
    SeedRec seed; std::memcpy(&seed, seed24, 24);
    const int64_t W = cam.width, H = cam.height;
    parallelFor(H, 2, [&](int64_t y) {
        for (int64_t x = 0; x < W; ++x) {
            size_t i = (size_t)(y * W + x);
            raygenPixel(sc, cam, seed, (uint32_t)x, (uint32_t)y, dirT + 4 * i, uvN + 4 * i, rays_out ? rays_out + 6 * i : nullptr,
                        flags_out ? flags_out + i : nullptr);
        }
    });
}

void orc_trace_rays(const orc_scene* s, const float* rays, uint64_t n, const uint32_t* prev, uint32_t* object, float* t,
                    float* uv, uint32_t* enc_normal2, uint8_t* flags_out) {
    Scene sc = view(s);
    parallelFor((int64_t)n, 64, [&](int64_t i) {
        Ray r = {v3(rays + 6 * i), v3(rays + 6 * i + 3)};
        uint32_t p = prev ? prev[i] : noRayHit;
        Hit hit = traceGeometry(sc, r, p);
        object[i] = hit.hitT == noHit ? noRayHit : hit.object;
        t[i] = hit.hitT;
        if (uv) { uv[2 * i] = hit.uv.x; uv[2 * i + 1] = hit.uv.y; }
        if (enc_normal2) encodeNormalGpu(hit.objectNormal, enc_normal2 + 2 * i);
        if (flags_out) flags_out[i] = rayFlags(sc, r, p, hit.hitT);
    });
}

void orc_occlusion_rays(const orc_scene* s, const float* rays, uint64_t n, const float* max_dist, const uint32_t* prev,
                        uint8_t* occluded) {
    Scene sc = view(s);
    parallelFor((int64_t)n, 64, [&](int64_t i) {
        Ray r = {v3(rays + 6 * i), v3(rays + 6 * i + 3)};
        occluded[i] = traceOcclusion(sc, r, max_dist ? max_dist[i] : noHit, prev ? prev[i] : noRayHit) ? 1 : 0;
    });
}

void orc_shadow(const orc_scene* s, const void* cam144, const void* seed24, uint32_t samples, const float* dirT,
                uint32_t* bits, float* shadow_rays_out) {
    Scene sc = view(s);
    CameraRec cam; std::memcpy(&cam, cam144, 144);
    SeedRec seed; std::memcpy(&seed, seed24, 24);
    const uint32_t W = cam.width, H = cam.height;
    const int64_t stripsY = (H + 1) / 2, stripsX = (W + 15) / 16;
    // one "warp" = one 16x2 strip of one sample (nv_all.shadow.comp:40-48); the ballot early-out at :69-82
    // leaves the word zero, and lanes whose own pixel missed contribute hit = false.
    const uint32_t layers = g_all_lights ? samples * sc.lightCount : samples;
    parallelFor((int64_t)layers * stripsY, 2, [&](int64_t job) {
        const int64_t layer = job / stripsY, sy = job % stripsY;
        const int64_t i = layer % samples;
        const uint32_t lightId = (uint32_t)(layer / samples);
        {
            for (int64_t sx = 0; sx < stripsX; ++sx) {
                uint32_t word = 0;
                bool any = false;
                for (uint32_t l = 0; l < 32; ++l) {
                    uint32_t x = (uint32_t)sx * 16 + (l & 15), y = (uint32_t)sy * 2 + (l >> 4);
                    if (x < W && y < H && fbits(dirT[4 * ((size_t)y * W + x) + 3]) != noRayHit) any = true;
                }
                if (any)
                    for (uint32_t l = 0; l < 32; ++l) {
                        uint32_t x = (uint32_t)sx * 16 + (l & 15), y = (uint32_t)sy * 2 + (l >> 4);
                        if (x >= W || y >= H) continue;
                        size_t px = (size_t)y * W + x;
                        float* ro = shadow_rays_out ? shadow_rays_out + 6 * ((size_t)layer * W * H + px) : nullptr;
                        if (shadowPixel(sc, cam, seed, samples, x, y, (uint32_t)i, dirT + 4 * px, ro, lightId)) word |= 1u << l;
                    }
                // nv_all.shadow.comp:69-82: a subgroup without hits leaves early; only DEBUG builds store the zero word (D10)
                if (any || g_mode == ORC_MODE_DEBUG) bits[indexToLightNV((uint32_t)sx * 16, (uint32_t)sy * 2, W, H, (uint32_t)layer)] = word;
            }
        }
    });
}

void orc_lighting(const orc_scene* s, const void* cam144, uint32_t samples, const float* dirT, const float* uvN,
                  const uint32_t* bits, uint16_t* lighting_f16, float* lighting_f32) {
    Scene sc = view(s);
    CameraRec cam; std::memcpy(&cam, cam144, 144);
    const int64_t W = cam.width, H = cam.height;
    parallelFor(H, 2, [&](int64_t y) {
        for (int64_t x = 0; x < W; ++x) {
            size_t i = (size_t)(y * W + x);
            vec3 l;
            bool hit = lightingPixel(sc, cam, samples, (uint32_t)x, (uint32_t)y, dirT + 4 * i, uvN + 4 * i, bits, (uint32_t)W, (uint32_t)H, l);
            float a = hit ? 1.0f : 0.0f;   // imageStore(lighting, vec4(light, 1)) / DEBUG vec4(0) on a miss
            if (!hit && g_mode != ORC_MODE_DEBUG) continue;   // nv_all.lighting.comp:55-62: RELEASE builds store nothing on a miss (D10)
            if (lighting_f32) { lighting_f32[4 * i] = l.x; lighting_f32[4 * i + 1] = l.y; lighting_f32[4 * i + 2] = l.z; lighting_f32[4 * i + 3] = a; }
            lighting_f16[4 * i] = f32_to_f16_rtne(l.x); lighting_f16[4 * i + 1] = f32_to_f16_rtne(l.y);
            lighting_f16[4 * i + 2] = f32_to_f16_rtne(l.z); lighting_f16[4 * i + 3] = f32_to_f16_rtne(a);
        }
    });
}

void orc_composite(const orc_scene* s, const void* cam144, const void* seed24, const float* dirT, const float* uvN,
                   const uint16_t* lighting_f16, float* accum, uint32_t* rgba8) {
    Scene sc = view(s);
    CameraRec cam; std::memcpy(&cam, cam144, 144);
    SeedRec seed; std::memcpy(&seed, seed24, 24);
    const int64_t W = cam.width, H = cam.height;
    parallelFor(H, 2, [&](int64_t y) {
        float dummy[4];
        for (int64_t x = 0; x < W; ++x) {
            size_t i = (size_t)(y * W + x);
            rgba8[i] = compositePixel(sc, cam, seed, dirT + 4 * i, uvN + 4 * i, lighting_f16 + 4 * i, accum ? accum + 4 * i : dummy);
        }
    });
}

void orc_frame(const orc_scene* s, const void* cam144, void* seed24, uint32_t samples, float* dirT, float* uvN,
               uint32_t* bits, uint16_t* lighting_f16, float* accum, uint32_t* rgba8) {
    CameraRec cam; std::memcpy(&cam, cam144, 144);
    const size_t px = (size_t)cam.width * cam.height;
    std::vector<float> dT, uN; std::vector<uint32_t> b; std::vector<uint16_t> lf;
    if (!dirT) { dT.resize(4 * px); dirT = dT.data(); }
    if (!uvN) { uN.resize(4 * px); uvN = uN.data(); }
    if (!bits) { b.resize(shadowWords(cam.width, cam.height, samples * (g_all_lights ? s->info[0] : 1u))); bits = b.data(); }
    if (!lighting_f16) { lf.resize(4 * px); lighting_f16 = lf.data(); }
    orc_init_pass(seed24);
    orc_raygen(s, cam144, seed24, dirT, uvN, nullptr, nullptr);
    orc_shadow(s, cam144, seed24, samples, dirT, bits, nullptr);
    orc_lighting(s, cam144, samples, dirT, uvN, bits, lighting_f16, nullptr);
    orc_composite(s, cam144, seed24, dirT, uvN, lighting_f16, accum, rgba8);
}

// ---- diffuse bounces: BASELINE.json configs[3].  NOT reference behaviour (the reference traces no secondary rays):
// a restatement of the definition in igx_raytracing_b200/csrc/rtb_path.cuh, built from the reference functions above, so that the
// CUDA path can be held bit-equal to a CPU statement of the same semantics.  Depth 0 is the reference's own G-buffer. ----
namespace {
struct PathOut { bool shadow, bounce; Ray sray; float maxDist; Ray bray; vec3 direct; };
PathOut pathVertex(const Scene& sc, vec2 uvBase, uint32_t depth, uint32_t bounces, vec3 pos, vec3 v, uint32_t object, vec3 n, vec3& T, vec3& L) {
    PathOut o;
    o.shadow = false; o.bounce = false; o.direct = {0, 0, 0}; o.maxDist = -1.0f;
    const MatU m = unpackMaterial(sc.mat[sc.matIdx[object]]);
    L = L + T * m.emissive;
    const uint32_t N = 2u * (bounces + 1u);
    if (sc.lightCount) {
        const vec2 random = rand2(uvBase + hammersley(2u * depth, N));
        const LightRec& light = sc.light[0];
        const vec3 F0 = mix(vec3{0.04f, 0.04f, 0.04f}, m.albedo, m.metallic);
        const float NdotV = glsl_max(dot(v, -n), 0.0f);
        const vec3 c = shadeLight(F0, m.albedo, m.roughness, m.metallic, light, pos, n, v, NdotV, random) * (float)sc.lightCount;
        const vec3 contribution = T * c;
        float brightness, dist;
        const vec3 l = getDirToLight(light, pos, brightness, dist, random);
        float maxDist = -1.0f;
        if (dist >= 0.0f) {
            uint32_t ro; std::memcpy(&ro, &light.rad, 4);
            const vec2 radOrigin = unpackHalf2x16(ro);
            if (dist >= radOrigin.y && dist < radOrigin.x) maxDist = dist - radOrigin.y;
        } else
            maxDist = noHit;
        if (maxDist != -1.0f) { o.shadow = true; o.direct = contribution; o.sray = {pos, -l}; o.maxDist = maxDist; }
        else L = L + contribution;
    }
    if (depth < bounces) {
        const vec2 r = rand2(uvBase + hammersley(2u * depth + 1u, N));
        const vec3 nn = normalize(n);
        const vec3 nf = dot(v, nn) > 0.0f ? -nn : nn;
        const float phi = (2.0f * pi) * r.x;
        const float cosT = std::sqrt(1.0f - r.y), sinT = std::sqrt(r.y);
        const float x = cr_cos(phi) * sinT, y = cr_sin(phi) * sinT;
        const vec3 bitangent = normalize(getPerpendicularVector(nf));
        const vec3 tangent = cross(bitangent, nf);
        const vec3 d = normalize(bitangent * x + tangent * y + nf * cosT);
        T = T * m.albedo;
        o.bounce = true; o.bray = {pos, d};
    }
    return o;
}
}  // namespace

void orc_path_frame(const orc_scene* s, const void* cam144, void* seed24, uint32_t bounces, float* dirT, float* uvN, float* accum,
                    uint32_t* rgba8, float* radiance3, uint64_t* rays_out) {
    Scene sc = view(s);
    CameraRec cam; std::memcpy(&cam, cam144, 144);
    orc_init_pass(seed24);
    SeedRec seed; std::memcpy(&seed, seed24, 24);
    const int64_t W = cam.width, H = cam.height;
    std::vector<float> dT, uN;
    if (!dirT) { dT.resize((size_t)W * H * 4); dirT = dT.data(); }
    if (!uvN) { uN.resize((size_t)W * H * 4); uvN = uN.data(); }
    std::atomic<uint64_t> rays{0};
    parallelFor(H, 1, [&](int64_t y) {
        uint64_t nrays = 0;
        for (int64_t x = 0; x < W; ++x) {
            const size_t i = (size_t)(y * W + x);
            float* dt = dirT + 4 * i; float* un = uvN + 4 * i;
            raygenPixel(sc, cam, seed, (uint32_t)x, (uint32_t)y, dt, un, nullptr, nullptr);
            ++nrays;
            const vec3 dxyz = {dt[0], dt[1], dt[2]};
            uint32_t object = fbits(dt[3]);
            vec3 T = {1, 1, 1}, L = {0, 0, 0};
            const vec2 loc = {(float)x, (float)y};
            const vec2 uvBase = (loc + rand2(loc + vec2{seed.randomX, seed.randomY})) / 128.0f;
            if (object == noRayHit)
                L = sampleSkybox(sc, cam, normalize(dxyz));
            else {
                vec3 pos = v3(cam.eye) + dxyz, v = normalize(dxyz), n = decodeNormal(fbits(un[2]), fbits(un[3]));
                for (uint32_t depth = 0;; ++depth) {
                    PathOut o = pathVertex(sc, uvBase, depth, bounces, pos, v, object, n, T, L);
                    if (o.shadow) { ++nrays; if (!traceOcclusion(sc, o.sray, o.maxDist, object)) L = L + o.direct; }
                    if (!o.bounce) break;
                    ++nrays;
                    const Hit hit = traceGeometry(sc, o.bray, object);
                    if (hit.hitT == noHit) { L = L + T * sampleSkybox(sc, cam, o.bray.dir); break; }
                    uint32_t en[2];
                    encodeNormalGpu(hit.objectNormal, en);
                    n = decodeNormal(en[0], en[1]);
                    pos = o.bray.pos + o.bray.dir * hit.hitT;
                    v = o.bray.dir;
                    object = hit.object;
                }
            }
            if (radiance3) { radiance3[3 * i] = L.x; radiance3[3 * i + 1] = L.y; radiance3[3 * i + 2] = L.z; }
            vec3 color = L;
            if (g_mode == ORC_MODE_DEBUG && (std::isnan(color.x) || std::isnan(color.y) || std::isnan(color.z))) color = {0.0f, 0.0f, 10000.0f};
            if (cam.flags & 2u) {
                float* a = accum + 4 * i;
                if (seed.sampleCount > 1) color = color + vec3{a[0], a[1], a[2]};
                a[0] = color.x; a[1] = color.y; a[2] = color.z; a[3] = 0.0f;
                color = color / (float)seed.sampleCount;
            }
            const vec3 e = -color * cam.exposure;
            color = vmax(vec3{1.0f, 1.0f, 1.0f} - vec3{cr_exp(e.x), cr_exp(e.y), cr_exp(e.z)}, vec3{0, 0, 0});
            rgba8[i] = unorm8(color.x) | (unorm8(color.y) << 8) | (unorm8(color.z) << 16) | (255u << 24);
        }
        rays.fetch_add(nrays, std::memory_order_relaxed);
    });
    if (rays_out) *rays_out = rays.load();
}

uint64_t orc_frame_pixels(const orc_scene* s, const void* cam144, const void* seed24, uint32_t samples, const uint32_t* xy,
                          uint64_t n, uint32_t* rgba8, uint32_t* object, float* t) {
    return orc_frame_pixels_ex(s, cam144, seed24, samples, xy, n, rgba8, object, t, nullptr, nullptr, nullptr);
}

uint64_t orc_frame_pixels_ex(const orc_scene* s, const void* cam144, const void* seed24, uint32_t samples, const uint32_t* xy,
                             uint64_t n, uint32_t* rgba8, uint32_t* object, float* t, float* dirT4, uint8_t* shadowed, uint8_t* flags) {
    Scene sc = view(s);
    CameraRec cam; std::memcpy(&cam, cam144, 144);
    SeedRec seed; std::memcpy(&seed, seed24, 24);
    std::atomic<uint64_t> raysTotal{0};
    parallelFor((int64_t)n, 16, [&](int64_t k) {
        uint64_t rays = 0;
        const uint32_t x = xy[2 * k], y = xy[2 * k + 1];
        float dirT[4], uvN[4], accum[4] = {0, 0, 0, 0};
        raygenPixel(sc, cam, seed, x, y, dirT, uvN, nullptr, flags ? flags + k : nullptr);
        rays += 1;
        if (dirT4) std::memcpy(dirT4 + 4 * k, dirT, 16);
        const bool isHit = fbits(dirT[3]) != noRayHit;
        // per-pixel stand-in for the shadow-mask buffer: bit (x&15 | (y&1)<<4) of one word per sample
        std::vector<uint32_t> words(samples, 0u);
        if (isHit)
            for (uint32_t i = 0; i < samples; ++i) {
                if (shadowPixel(sc, cam, seed, samples, x, y, i, dirT, nullptr)) words[i] = 0xFFFFFFFFu;
                rays += 1;
            }
        if (shadowed) shadowed[k] = (isHit && samples && words[0]) ? 1 : 0;   // sample 0's bit
        vec3 l = {0, 0, 0};
        {
            // lightingPixel indexes the real mask; emulate with a 1-word-per-sample view
            uint32_t object_ = fbits(dirT[3]);
            if (object_ != noRayHit) {
                vec3 dxyz = {dirT[0], dirT[1], dirT[2]};
                vec3 hitPos = v3(cam.eye) + dxyz;
                MatU m = unpackMaterial(sc.mat[sc.matIdx[object_]]);
                vec3 F0 = mix(vec3{0.04f, 0.04f, 0.04f}, m.albedo, m.metallic);
                vec3 nn = decodeNormal(fbits(uvN[2]), fbits(uvN[3]));
                vec3 v = normalize(dxyz);
                float NdotV = glsl_max(dot(v, -nn), 0.0f);
                vec2 loc = {(float)x, (float)y};
                vec2 uv = (loc + rand2(loc + vec2{0.0f, 0.0f})) / 128.0f;
                vec3 light = {0, 0, 0};
                for (uint32_t i = 0; i < samples; ++i) {
                    vec2 random = rand2(uv + hammersley(i, samples));
                    if (!words[i]) light = light + shadeLight(F0, m.albedo, m.roughness, m.metallic, sc.light[0], hitPos, nn, v, NdotV, random);
                }
                l = light / (float)samples * (float)sc.lightCount;
            }
        }
        uint16_t lf[4] = {f32_to_f16_rtne(l.x), f32_to_f16_rtne(l.y), f32_to_f16_rtne(l.z), f32_to_f16_rtne(isHit ? 1.0f : 0.0f)};
        CameraRec camNoAccum = cam;
        camNoAccum.flags &= ~2u;
        rgba8[k] = compositePixel(sc, camNoAccum, seed, dirT, uvN, lf, accum);
        if (object) object[k] = fbits(dirT[3]);
        if (t) t[k] = isHit ? length(vec3{dirT[0], dirT[1], dirT[2]}) : noHit;
        raysTotal.fetch_add(rays, std::memory_order_relaxed);
    });
    return raysTotal.load();
}

}  // extern "C"
