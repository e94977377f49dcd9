// rtb_math.cuh — device-side maths of the hot path: the arithmetic of the reference shaders
// (res/shaders/*.glsl, "SH/" below) written for CUDA.
//
// Numerical contract (DESIGN.md "numerics"): this translation unit is compiled with -fmad=false, IEEE
// division and square root, no flush-to-zero, so +,-,*,/ and sqrt are single IEEE binary32 operations in
// source order.  Where a fused multiply-add is wanted (BVH slab tests, which have no reference
// counterpart) it is written explicitly as fmaf().  Transcendentals (sin cos asin acos atan atan2 exp
// pow) are evaluated in binary64 and rounded once, which is what GLSL's "implementation-defined"
// precision is pinned to for this project.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>
#include "rtb_types.h"
#include "rtb_crmath.h"

namespace rtb {

#define RTB_DI __device__ __forceinline__

struct vec2 { float x, y; };
struct vec3 { float x, y, z; };

RTB_DI vec2 mk2(float x, float y) { vec2 r; r.x = x; r.y = y; return r; }
RTB_DI vec3 mk3(float x, float y, float z) { vec3 r; r.x = x; r.y = y; r.z = z; return r; }
RTB_DI vec3 mk3(const float* p) { return mk3(p[0], p[1], p[2]); }

RTB_DI vec2 operator+(vec2 a, vec2 b) { return mk2(a.x + b.x, a.y + b.y); }
RTB_DI vec2 operator*(vec2 a, vec2 b) { return mk2(a.x * b.x, a.y * b.y); }
RTB_DI vec2 operator*(vec2 a, float s) { return mk2(a.x * s, a.y * s); }
RTB_DI vec2 operator+(vec2 a, float s) { return mk2(a.x + s, a.y + s); }
RTB_DI vec2 operator/(vec2 a, float s) { return mk2(a.x / s, a.y / s); }

RTB_DI vec3 operator+(vec3 a, vec3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
RTB_DI vec3 operator-(vec3 a, vec3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
RTB_DI vec3 operator*(vec3 a, vec3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
RTB_DI vec3 operator/(vec3 a, vec3 b) { return mk3(a.x / b.x, a.y / b.y, a.z / b.z); }
RTB_DI vec3 operator*(vec3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
RTB_DI vec3 operator*(float s, vec3 a) { return mk3(s * a.x, s * a.y, s * a.z); }
RTB_DI vec3 operator/(vec3 a, float s) { return mk3(a.x / s, a.y / s, a.z / s); }
RTB_DI vec3 operator+(vec3 a, float s) { return mk3(a.x + s, a.y + s, a.z + s); }
RTB_DI vec3 operator-(vec3 a, float s) { return mk3(a.x - s, a.y - s, a.z - s); }
RTB_DI vec3 operator-(float s, vec3 a) { return mk3(s - a.x, s - a.y, s - a.z); }
RTB_DI vec3 operator-(vec3 a) { return mk3(-a.x, -a.y, -a.z); }

RTB_DI float dot(vec2 a, vec2 b) { return a.x * b.x + a.y * b.y; }
RTB_DI float dot(vec3 a, vec3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
RTB_DI vec3 cross(vec3 a, vec3 b) { return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
RTB_DI float length(vec3 v) { return sqrtf(dot(v, v)); }
RTB_DI vec3 normalize(vec3 v) { float inv = 1.0f / sqrtf(dot(v, v)); return v * inv; }
RTB_DI vec3 reflect(vec3 i, vec3 n) { return i - (2.0f * dot(n, i)) * n; }
RTB_DI float fractf(float x) { return x - floorf(x); }
RTB_DI float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }
RTB_DI vec3 mix(vec3 a, vec3 b, float t) { return mk3(mixf(a.x, b.x, t), mixf(a.y, b.y, t), mixf(a.z, b.z, t)); }
RTB_DI vec3 vmax(vec3 a, vec3 b) { return mk3(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); }
RTB_DI vec3 vmin(vec3 a, vec3 b) { return mk3(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)); }
RTB_DI float signf(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }

// binary64 evaluation, one rounding
RTB_DI float cr_sin(float x) { return cr_sin_f(x); }   // rtb_crmath.h: binary64 accuracy without the Payne-Hanek path
RTB_DI float cr_cos(float x) { return cr_cos_f(x); }
RTB_DI float cr_asin(float x) { return (float)asin((double)x); }
RTB_DI float cr_acos(float x) { return (float)acos((double)x); }
RTB_DI float cr_atan(float x) { return (float)atan((double)x); }
RTB_DI float cr_atan2(float y, float x) { return (float)atan2((double)y, (double)x); }
RTB_DI float cr_exp(float x) { return (float)exp((double)x); }
RTB_DI float cr_pow(float x, float y) { return (float)pow((double)x, (double)y); }

RTB_DI uint32_t f2u(float f) { return __float2uint_rz(f); }   // NaN, negatives -> 0; saturates
RTB_DI uint32_t fbits(float f) { return __float_as_uint(f); }
RTB_DI float ubits(uint32_t u) { return __uint_as_float(u); }

constexpr float PI_F = 3.1415927410125732421875f;   // SH/rand_util.glsl:13

RTB_DI float h2f(uint32_t h16) { return __half2float(__ushort_as_half((unsigned short)h16)); }
RTB_DI vec2 unpackHalf2x16(uint32_t v) { return mk2(h2f(v & 0xFFFFu), h2f(v >> 16)); }
RTB_DI uint32_t f2h_rn(float f) { return (uint32_t)__half_as_ushort(__float2half_rn(f)); }

// ---- random numbers: SH/rand_util.glsl ---------------------------------------------------------------
RTB_DI float rand1(vec2 co) { return fractf(cr_sin(dot(co, mk2(12.9898f, 78.233f))) * 43758.5453f); }   // :115-117
RTB_DI vec2 rand2(vec2 p) { return mk2(rand1(p), rand1(p * 1103515245.0f + 12345.0f)); }                // :127-129
RTB_DI vec2 hammersley(uint32_t i, uint32_t N) {                                                         // :96-111
    return mk2((float)i / (float)N, (float)__brev(i) * 2.3283064365386963e-10f);
}
RTB_DI vec3 randomPointOnUnitSphere(vec2 r) {                                                            // :33-39
    float px = (2.0f * PI_F) * r.x, py = cr_acos(1.0f - 2.0f * r.y);
    float sx = cr_sin(px), sy = cr_sin(py), cx = cr_cos(px), cy = cr_cos(py);
    return mk3(sx * cy, sx * sy, cx);
}
RTB_DI vec3 getPerpendicularVector(vec3 u) {                                                             // :55-64
    float ax = fabsf(u.x), ay = fabsf(u.y), az = fabsf(u.z);
    uint32_t xm = (ax - ay < 0.0f && ax - az < 0.0f) ? 1u : 0u;
    uint32_t ym = (ay - az < 0.0f) ? (1u ^ xm) : 0u;
    uint32_t zm = 1u ^ (xm | ym);
    return cross(u, mk3((float)xm, (float)ym, (float)zm));
}
RTB_DI vec3 getSunDirection(vec2 random, vec3 direction, float angularExtent) {                          // :66-85
    float h = cr_cos(angularExtent);
    float phi = (2.0f * PI_F) * random.x;
    float z = h + (1.0f - h) * random.y;
    float sinT = sqrtf(1.0f - z * z);
    float x = cr_cos(phi) * sinT;
    float y = cr_sin(phi) * sinT;
    vec3 bitangent = getPerpendicularVector(direction);
    vec3 tangent = cross(bitangent, direction);
    return bitangent * x + tangent * y + direction * z;
}

// ---- primitives: SH/primitive.glsl ------------------------------------------------------------------
struct Ray { vec3 pos, dir; };
struct Hit { float hitT; vec2 uv; uint32_t object; vec3 geometryNormal; };

RTB_DI void encodeNormalGpu(vec3 n, uint32_t& ex, uint32_t& ey) {                                        // :85-88
    vec3 v = (normalize(n) * 0.5f + 0.5f) * 65535.0f;
    ex = (f2u(v.x) << 16) | f2u(v.y);
    ey = f2u(v.z);
}
RTB_DI vec3 decodeNormal(uint32_t sx, uint32_t sy) {                                                     // :90-93
    vec3 nh = mk3((float)(sx >> 16), (float)(sx & 65535u), (float)sy);
    return nh / 65535.0f * 2.0f - 1.0f;
}
RTB_DI vec3 decodeSpheremap(uint32_t n) {                                                                // :95-104
    vec2 nn = unpackHalf2x16(n);
    float l = dot(mk3(nn.x, nn.y, 1.0f), -mk3(nn.x, nn.y, -1.0f));
    float sq = sqrtf(l);
    nn = nn * sq;
    return mk3(nn.x, nn.y, l) * 2.0f + mk3(0.0f, 0.0f, -1.0f);
}
RTB_DI vec3 unpackColor3(uint32_t cx, uint32_t cy) { vec2 rg = unpackHalf2x16(cx); return mk3(rg.x, rg.y, unpackHalf2x16(cy).x); }  // :128-133
RTB_DI float unpackColorAUnorm(uint32_t cy) { return (float)(cy >> 16) / 65535.0f; }                     // :139-142
RTB_DI vec3 interpolate(vec3 a, vec3 b, vec3 c, vec2 uv) {                                               // :146-151
    float bz = 1.0f - uv.x - uv.y;
    return uv.x * b + uv.y * c + bz * a;
}

// Möller–Trumbore on (p0, e1, e2): SH/primitive.glsl:239-284.  Returns the candidate (u, v, t) when the
// barycentric tests pass; the caller applies the distance rule (t > 0, t < best, ties by index).
RTB_DI bool triCandidate(vec3 ro, vec3 rd, vec3 p0, vec3 e1, vec3 e2, float& u, float& v, float& t, float& a) {
    vec3 h = cross(rd, e2);
    a = dot(e1, h);
    float f = 1.0f / a;
    vec3 s = ro - p0;
    u = f * dot(s, h);
    if (u < 0.0f || u > 1.0f) return false;
    vec3 q = cross(s, e1);
    v = f * dot(rd, q);
    if (v < 0.0f || u + v > 1.0f) return false;
    t = f * dot(e2, q);
    return true;
}

RTB_DI bool rayIntersectSphere(const Ray& r, float4 sph, Hit& hit, uint32_t obj, uint32_t prevObj) {     // :173-210
    vec3 c = mk3(sph.x, sph.y, sph.z);
    vec3 dif = c - r.pos;
    float t = dot(dif, r.dir);
    vec3 Q = dif - t * r.dir;
    float Q2 = dot(Q, Q);
    float R2 = sph.w * sph.w;
    bool outOfSphere = Q2 > R2;
    float hitT = t - sqrtf(R2 - Q2);
    if (!outOfSphere && obj != prevObj && hitT >= 0.0f && hitT < hit.hitT) {
        hit.hitT = hitT;
        vec3 o = hitT * r.dir + r.pos;
        vec3 normal = normalize(c - o);
        hit.geometryNormal = normal;
        float latitude = cr_asin(normal.z);
        float longitude = cr_atan(normal.y / normal.x);
        if (isnan(longitude)) longitude = 0.0f;
        hit.uv = mk2(latitude, longitude) * (0.636619746685f * 0.5f) + 0.5f;
        return true;
    }
    return false;
}
// occlusion only needs the distance
RTB_DI float sphereCandidateT(const Ray& r, float4 sph) {
    vec3 c = mk3(sph.x, sph.y, sph.z);
    vec3 dif = c - r.pos;
    float t = dot(dif, r.dir);
    vec3 Q = dif - t * r.dir;
    float Q2 = dot(Q, Q);
    float R2 = sph.w * sph.w;
    float hitT = t - sqrtf(R2 - Q2);
    return (!(Q2 > R2) && hitT >= 0.0f) ? hitT : NO_HIT;
}

RTB_DI bool rayIntersectPlane(const Ray& r, float4 pl, Hit& hit, uint32_t obj, uint32_t prevObj) {       // :212-237
    vec3 pxyz = mk3(pl.x, pl.y, pl.z);
    vec3 dir = normalize(pxyz);
    float dif = dot(r.dir, -dir);
    float hitT = -(dot(r.pos, -dir) + pl.w) / dif;
    if (hitT >= 0.0f && obj != prevObj && hitT < hit.hitT) {
        hit.hitT = hitT;
        hit.geometryNormal = dif > 0.0f ? -dir : dir;
        vec3 o = hitT * r.dir + r.pos;
        vec3 planeX = cross(pxyz, mk3(0.0f, 0.0f, 1.0f));
        vec3 planeZ = cross(pxyz, mk3(1.0f, 0.0f, 0.0f));
        hit.uv = mk2(dot(o, planeX), dot(o, planeZ));
        return true;
    }
    return false;
}

RTB_DI bool rayIntersectCube(const Ray& r, const float* cube, Hit& hit, uint32_t obj, uint32_t prevObj) { // :286-333
    vec3 revDir = mk3(1.0f / r.dir.x, 1.0f / r.dir.y, 1.0f / r.dir.z);
    vec3 start = mk3(cube[0], cube[1], cube[2]);
    vec3 end = mk3(cube[3], cube[4], cube[5]);
    vec3 startDir = (start - r.pos) * revDir;
    vec3 endDir = (end - r.pos) * revDir;
    vec3 mi = vmin(startDir, endDir);
    vec3 ma = vmax(startDir, endDir);
    float tmin = fmaxf(fmaxf(mi.x, mi.y), mi.z);
    float tmax = fminf(fminf(ma.x, ma.y), ma.z);
    if (tmax < 0.0f || tmin > tmax || tmin > hit.hitT || obj == prevObj) return false;
    vec3 pos = (r.dir * tmin + r.pos) - start;
    pos = pos / end;
    if (tmin == mi.x) {
        int isLeft = (mi.x == startDir.x) ? 1 : 0;
        hit.geometryNormal = mk3((float)(isLeft * 2 - 1), 0.0f, 0.0f);
        hit.uv = mk2(pos.y, pos.z);
    } else if (tmin == mi.y) {
        int isDown = (mi.y == startDir.y) ? 1 : 0;
        hit.geometryNormal = mk3(0.0f, (float)(isDown * 2 - 1), 0.0f);
        hit.uv = mk2(pos.x, pos.z);
    } else {
        int isBack = (mi.z == startDir.z) ? 1 : 0;
        hit.geometryNormal = mk3(0.0f, 0.0f, (float)(isBack * 2 - 1));
        hit.uv = mk2(pos.x, pos.y);
    }
    hit.hitT = tmin;
    return true;
}

// ---- camera: SH/camera.glsl ---------------------------------------------------------------------------
RTB_DI Ray calculateOmni(const CameraRec& cam, vec2 c, bool isLeft) {                                     // :53-70
    vec2 spherical = mk2(c.x - 0.5f, 0.5f - c.y) * mk2(2.0f * PI_F, PI_F);
    float sx = cr_sin(spherical.x), sy = cr_sin(spherical.y), cx = cr_cos(spherical.x), cy = cr_cos(spherical.y);
    Ray r;
    r.pos = mk3(cam.eye) + mk3(cx, 0.0f, sx) * (cam.ipd * 5e-4f) * (isLeft ? -1.0f : 1.0f);
    r.dir = mk3(sx * cy, sy, -cx * cy);
    return r;
}
RTB_DI Ray calculateScreen(const CameraRec& cam, vec2 c, bool isRight) {                                  // :72-87
    vec3 p0 = isRight ? mk3(cam.p3) : mk3(cam.p0);
    vec3 p1 = isRight ? mk3(cam.p4) : mk3(cam.p1);
    vec3 p2 = isRight ? mk3(cam.p5) : mk3(cam.p2);
    vec3 right = p1 - p0, up = p2 - p0;
    vec3 pos = p0 + c.x * right + c.y * up;
    Ray r;
    r.pos = mk3(cam.eye);
    r.dir = normalize(pos - mk3(cam.eye));
    return r;
}
RTB_DI Ray calculatePrimary(const CameraRec& cam, uint32_t lx, uint32_t ly, vec2 randLoc) {               // :113-138
    vec2 loc = mk2((float)lx, (float)ly);
    vec2 c = (loc + rand2(loc + randLoc)) * mk2(cam.invRes[0], cam.invRes[1]);
    c.y = 1.0f - c.y;
    switch (cam.projectionType) {
        case 1: return calculateOmni(cam, c, false);
        case 2: return calculateOmni(cam, mk2(c.x, fractf(c.y * 2.0f)), c.y < 0.5f);
        case 4: return calculateOmni(cam, mk2(fractf(c.x * 2.0f), c.y), c.x < 0.5f);
        case 3:
            if (c.y < 0.5f) return calculateScreen(cam, mk2(c.x, c.y * 2.0f), false);
            return calculateScreen(cam, mk2(c.x, c.y * 2.0f - 1.0f), true);
        case 5:
            if (c.x < 0.5f) return calculateScreen(cam, mk2(c.x * 2.0f, c.y), false);
            return calculateScreen(cam, mk2(c.x * 2.0f - 1.0f, c.y), true);
        default: return calculateScreen(cam, c, false);
    }
}

// ---- skybox: SH/scene.glsl:56-69 ----------------------------------------------------------------------
// rgba16f equirect, LOD 0, bilinear with binary32 weights, clamp-to-border with the GL default border (0,0,0,0).
struct SkyView { const uint2* texels; uint32_t w, h; };   // one uint2 = 4 halfs
RTB_DI vec3 skyTexel(const SkyView& s, long long x, long long y) {
    if (x < 0 || y < 0 || x >= (long long)s.w || y >= (long long)s.h) return mk3(0.0f, 0.0f, 0.0f);
    uint2 p = __ldg(s.texels + (size_t)y * s.w + (size_t)x);
    return mk3(h2f(p.x & 0xFFFFu), h2f(p.x >> 16), h2f(p.y & 0xFFFFu));
}
RTB_DI vec3 sampleSkybox(const SkyView& s, const CameraRec& cam, vec3 dir) {
    if (s.w == 0 || s.h == 0) return mk3(cam.skyboxColor);
    vec2 uv = mk2(cr_atan2(dir.x, dir.z), cr_asin(dir.y * -1.0f)) * mk2(0.1591f, 0.3183f) + 0.5f;
    float fx = uv.x * (float)s.w - 0.5f, fy = uv.y * (float)s.h - 0.5f;
    float x0f = floorf(fx), y0f = floorf(fy);
    float ax = fx - x0f, ay = fy - y0f;
    if (isnan(fx) || isnan(fy)) { float q = ubits(0x7FC00000u); return mk3(q, q, q); }
    long long x0 = (long long)x0f, y0 = (long long)y0f;
    vec3 t00 = skyTexel(s, x0, y0), t10 = skyTexel(s, x0 + 1, y0), t01 = skyTexel(s, x0, y0 + 1), t11 = skyTexel(s, x0 + 1, y0 + 1);
    vec3 top = t00 * (1.0f - ax) + t10 * ax;
    vec3 bot = t01 * (1.0f - ax) + t11 * ax;
    return top * (1.0f - ay) + bot * ay;
}

// ---- lights and shading: SH/light.glsl ----------------------------------------------------------------
constexpr float MIN_ROUGHNESS = 0.01f, SPECULAR_EPSILON = 0.001f;   // :6-7

RTB_DI float smoothstepf(float e0, float e1, float x) {   // reversed edges evaluate the Hermite form (DESIGN.md decree D9)
    float t = (x - e0) / (e1 - e0);
    t = fminf(fmaxf(t, 0.0f), 1.0f);
    return t * t * (3.0f - 2.0f * t);
}

// getSunDirection with everything that depends on the light alone taken from `f` (k_sun_frame evaluates those expressions once)
RTB_DI vec3 getSunDirectionPre(vec2 random, const SunFrame& f) {
    float h = f.h;
    float phi = (2.0f * PI_F) * random.x;
    float z = h + (1.0f - h) * random.y;
    float sinT = sqrtf(1.0f - z * z);
    float x = cr_cos(phi) * sinT;
    float y = cr_sin(phi) * sinT;
    return mk3(f.bitangent[0], f.bitangent[1], f.bitangent[2]) * x + mk3(f.tangent[0], f.tangent[1], f.tangent[2]) * y + mk3(f.dir[0], f.dir[1], f.dir[2]) * z;
}

RTB_DI vec3 getDirToLight(const LightRec& light, vec3 pos, float& brightness, float& dist, vec2 random, const SunFrame* sun = nullptr) {   // :98-133
    vec3 l;
    brightness = 1.0f;
    dist = -1.0f;
    vec2 radOrigin = unpackHalf2x16(light.radOrigin);
    radOrigin = mk2(fmaxf(radOrigin.x, 0.0f), fmaxf(radOrigin.y, 0.0f));
    radOrigin.y = fminf(radOrigin.y, radOrigin.x);
    if ((light.colorBType >> 16) == LIGHT_POINT) {
        l = pos - mk3(light.pos);
        dist = length(l);
        vec3 p = mk3(light.pos) + randomPointOnUnitSphere(random) * radOrigin.y;   // SH/rand_util.glsl:43-51, flip is a no-op
        l = pos - p;
        float r = radOrigin.x - radOrigin.y;
        float d = fmaxf(dist - radOrigin.y, 0.0f);
        brightness = cr_pow(smoothstepf(r, 0.0f, d), ubits(light.dir[0]));
    } else if (sun)
        l = getSunDirectionPre(random, *sun);
    else
        l = getSunDirection(random, normalize(decodeNormal(light.dir[0], light.dir[1])), radOrigin.x);
    return normalize(l);
}

RTB_DI float ndfGGX(vec3 n, vec3 h, float roughness) {                                                     // :22-35
    float alpha = roughness * roughness;
    float a2 = alpha * alpha;
    float NdotH = fmaxf(dot(n, h), 0.0f);
    float denom = NdotH * NdotH * (a2 - 1.0f) + 1.0f;
    denom *= denom * PI_F;
    if (denom == 0.0f) return 0.0f;
    return a2 / denom;
}
RTB_DI float geomSchlickGGX(float NdotV, float k) { return NdotV / (NdotV * (1.0f - k) + k); }
RTB_DI float pow5(float f) { float f2 = f * f; return f2 * f2 * f; }

// the Cook-Torrance term of one light once its direction and brightness are known (the body of shadeLight after getDirToLight)
RTB_DI vec3 shadeLightDir(vec3 F0, vec3 albedo, float roughness, float metallic, const LightRec& light, vec3 l, float brightness,
                          vec3 n, vec3 v, float NdotV) {                                                    // :139-159, :64-94
    float k = roughness + 1.0f;
    k *= k / 8.0f;
    float NdotL = fmaxf(dot(n, l), 0.0f);
    vec3 h = normalize(l + v);
    float D = ndfGGX(n, h, fmaxf(roughness, MIN_ROUGHNESS));
    float G = geomSchlickGGX(NdotV, k) * geomSchlickGGX(NdotL, k);
    vec3 F = F0 + (1.0f - F0) * pow5(1.0f - fmaxf(dot(h, v), 0.0f));
    float denom = 4.0f * NdotL * NdotV + SPECULAR_EPSILON;
    vec3 num = D * G * F;
    vec3 kS = num / denom;
    vec3 kD = (1.0f - F) * (1.0f - metallic);
    vec3 color = kD * albedo + kS;
    return color * unpackColor3(light.colorRG, light.colorBType) * brightness * NdotL;
}
RTB_DI vec3 shadeLight(vec3 F0, vec3 albedo, float roughness, float metallic, const LightRec& light, vec3 pos,
                       vec3 n, vec3 v, float NdotV, vec2 random, const SunFrame* sun = nullptr) {           // :135-159
    float brightness, dst;
    vec3 l = getDirToLight(light, pos, brightness, dst, random, sun);
    return shadeLightDir(F0, albedo, roughness, metallic, light, l, brightness, n, v, NdotV);
}

struct MatU { vec3 albedo, ambient, emissive; float metallic, roughness; };
RTB_DI MatU unpackMaterial(const MaterialRec* m) {
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(m));
    const uint2 b = __ldg(reinterpret_cast<const uint2*>(m) + 2);
    MatU o;
    o.albedo = unpackColor3(a.x, a.y); o.metallic = unpackColorAUnorm(a.y);
    o.ambient = unpackColor3(a.z, a.w); o.roughness = unpackColorAUnorm(a.w);
    o.emissive = unpackColor3(b.x, b.y);
    return o;
}

// SH/light.glsl:58-60 and :163-187
RTB_DI vec3 shade(const MatU& m, float NdotV, vec3 light, vec3 reflected) {
    vec3 F0 = mix(mk3(0.04f, 0.04f, 0.04f), m.albedo, m.metallic);
    float r1 = 1.0f - m.roughness;
    vec3 kS = F0 + (vmax(F0, mk3(r1, r1, r1)) - F0) * cr_pow5_f(1.0f - NdotV);
    vec3 kD = (1.0f - kS) * (1.0f - m.metallic);
    return (m.ambient + kD / PI_F) * m.albedo + kS * reflected + light + m.emissive;
}

// shadow-mask addressing: SH/light.glsl:221-229 with the 32-wide constants of SH/light_rt.glsl:21-24
RTB_DI uint32_t shadowTilesX(uint32_t w) { return (w >> 4) + ((w & 15u) != 0u); }
RTB_DI uint32_t shadowTilesY(uint32_t h) { return (h >> 1) + ((h & 1u) != 0u); }
RTB_DI uint32_t indexToLight(uint32_t lx, uint32_t ly, uint32_t w, uint32_t h, uint32_t sample) {
    uint32_t tx = shadowTilesX(w), ty = shadowTilesY(h);
    return (lx >> 4) + (ly >> 1) * tx + sample * ty * tx;
}

RTB_DI uint32_t unorm8(float c) {   // imageStore to rgba8
    if (isnan(c)) return 0u;
    c = c < 0.0f ? 0.0f : (c > 1.0f ? 1.0f : c);
    return (uint32_t)floorf(c * 255.0f + 0.5f);
}

}  // namespace rtb
