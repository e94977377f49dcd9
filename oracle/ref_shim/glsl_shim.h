/*
 * glsl_shim.h — the GLSL 4.50 vocabulary the reference's hot-path shaders use, as C++20, so that the reference's OWN shader
 * sources (adapted for syntax only by glsl_front.py) compile for the host and can be executed.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).  This file is ours; it contains no reference code.  Its job is to give
 * every GLSL built-in the meaning the numerics contract decrees (header of oracle/oracle.cpp, D1-D9), so that a bitwise
 * difference between oracle/ and oracle/_ref can only come from the oracle's restatement of the shader code:
 *
 *   D1  + - * / sqrt are single IEEE binary32 operations (the Makefile compiles with -ffp-contract=off, no fast-math, and
 *       -fsingle-precision-constant so that GLSL's unsuffixed literals are 32-bit as in GLSL);
 *   D2  dot = ((x*x' + y*y') + z*z') [+ w*w'], cross = textbook;   D3  normalize(v) = v * (1/sqrt(dot(v,v))), length = sqrt(dot);
 *   D4  sin cos asin acos atan exp pow: evaluated in binary64, rounded once;
 *   D5  float -> uint truncates, NaN / negative -> 0, saturating;  D6  image stores: rgba16f round-to-nearest-even, rgba8
 *       floor(clamp(c,0,1)*255+0.5) with NaN -> 0;  D7  nearest sampling = texel fetch, linear = binary32 weights, border 0;
 *   D9  smoothstep evaluates the Hermite form on clamp((x-e0)/(e1-e0),0,1) whatever the edge order;
 *   min/max return the non-NaN operand (IEEE minNum/maxNum, what NVIDIA hardware does); mix(a,b,t) = a*(1-t) + b*t;
 *   reflect(i,n) = i - (2*dot(n,i))*n; fract(x) = x - floor(x); mod(x,y) = x - y*floor(x/y).
 */
#ifndef GLSL_SHIM_H
#define GLSL_SHIM_H
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <type_traits>

namespace glsl {

typedef uint32_t uint;
struct vec2; struct vec3; struct vec4; struct uvec2; struct uvec3; struct uvec4; struct ivec2; struct bvec2; struct bvec3; struct bvec4;

template <class S> using if_arith = std::enable_if_t<std::is_arithmetic<S>::value, int>;

// D5
inline uint f2u(float f) { if (!(f > 0.0f)) return 0u; if (f >= 4294967296.0f) return 0xFFFFFFFFu; return (uint)f; }
inline uint cvt_u(float f) { return f2u(f); }
inline uint cvt_u(uint u) { return u; }
inline uint cvt_u(int i) { return (uint)i; }
inline uint cvt_u(bool b) { return b ? 1u : 0u; }

// ---- swizzles: empty proxy objects living in a union with the components -----------------------------------------------
template <class V, class T, int A, int B> struct swz2 {
    operator V() const { const T* d = reinterpret_cast<const T*>(this); return V(d[A], d[B]); }
    swz2& operator=(const V& v) { T* d = reinterpret_cast<T*>(this); T a = v.x, b = v.y; d[A] = a; d[B] = b; return *this; }
    swz2& operator=(const swz2& o) { return *this = V(o); }
    template <class R> swz2& operator+=(const R& r) { return *this = V(*this) + r; }
    template <class R> swz2& operator-=(const R& r) { return *this = V(*this) - r; }
    template <class R> swz2& operator*=(const R& r) { return *this = V(*this) * r; }
    template <class R> swz2& operator/=(const R& r) { return *this = V(*this) / r; }
};
template <class V, class T, int A, int B, int C> struct swz3 {
    operator V() const { const T* d = reinterpret_cast<const T*>(this); return V(d[A], d[B], d[C]); }
    swz3& operator=(const V& v) { T* d = reinterpret_cast<T*>(this); T a = v.x, b = v.y, c = v.z; d[A] = a; d[B] = b; d[C] = c; return *this; }
    swz3& operator=(const swz3& o) { return *this = V(o); }
    template <class R> swz3& operator+=(const R& r) { return *this = V(*this) + r; }
    template <class R> swz3& operator-=(const R& r) { return *this = V(*this) - r; }
    template <class R> swz3& operator*=(const R& r) { return *this = V(*this) * r; }
    template <class R> swz3& operator/=(const R& r) { return *this = V(*this) / r; }
};
template <class V, class T, int A, int B, int C, int D> struct swz4 {
    operator V() const { const T* d = reinterpret_cast<const T*>(this); return V(d[A], d[B], d[C], d[D]); }
};

// ---- vector types ---------------------------------------------------------------------------------------------------------
struct bvec2 { bool x, y; };
struct bvec3 { bool x, y, z; };
struct bvec4 { bool x, y, z, w; };

struct ivec2 {
    int x, y;
    ivec2() = default;
    template <class S, if_arith<S> = 0> explicit ivec2(S s) : x((int)s), y((int)s) {}
    template <class A, class B, if_arith<A> = 0, if_arith<B> = 0> ivec2(A a, B b) : x((int)a), y((int)b) {}
    explicit ivec2(const uvec2& u);
};

struct uvec2 {
    union {
        struct { uint x, y; };
        struct { uint r, g; };
        swz2<uvec2, uint, 0, 1> xy, rg;
    };
    static const int N = 2;
    uvec2() = default;
    uvec2(const uvec2& o) : x(o.x), y(o.y) {}
    uvec2& operator=(const uvec2& o) { x = o.x; y = o.y; return *this; }
    template <class S, if_arith<S> = 0> explicit uvec2(S s) : x(cvt_u(s)), y(cvt_u(s)) {}
    template <class A, class B, if_arith<A> = 0, if_arith<B> = 0> uvec2(A a, B b) : x(cvt_u(a)), y(cvt_u(b)) {}
    explicit uvec2(const bvec2& b) : x(b.x), y(b.y) {}
    explicit uvec2(const vec2& v);
    uint& operator[](int i) { return (&x)[i]; }
    const uint& operator[](int i) const { return (&x)[i]; }
};
struct uvec3 {
    union {
        struct { uint x, y, z; };
        struct { uint r, g, b; };
        swz2<uvec2, uint, 0, 1> xy, rg;
        swz3<uvec3, uint, 0, 1, 2> xyz, rgb;
    };
    static const int N = 3;
    uvec3() = default;
    uvec3(const uvec3& o) : x(o.x), y(o.y), z(o.z) {}
    uvec3& operator=(const uvec3& o) { x = o.x; y = o.y; z = o.z; return *this; }
    template <class S, if_arith<S> = 0> explicit uvec3(S s) : x(cvt_u(s)), y(cvt_u(s)), z(cvt_u(s)) {}
    template <class A, class B, class C, if_arith<A> = 0, if_arith<B> = 0, if_arith<C> = 0> uvec3(A a, B b, C c) : x(cvt_u(a)), y(cvt_u(b)), z(cvt_u(c)) {}
    explicit uvec3(const vec3& v);
    uint& operator[](int i) { return (&x)[i]; }
    const uint& operator[](int i) const { return (&x)[i]; }
};
struct uvec4 {
    union { struct { uint x, y, z, w; }; struct { uint r, g, b, a; }; };
    static const int N = 4;
    uvec4() = default;
    uvec4(const vec2& a, const vec2& b);
    template <class A, class B, class C, class D, if_arith<A> = 0, if_arith<B> = 0, if_arith<C> = 0, if_arith<D> = 0>
    uvec4(A a_, B b_, C c_, D d_) : x(cvt_u(a_)), y(cvt_u(b_)), z(cvt_u(c_)), w(cvt_u(d_)) {}
    uint& operator[](int i) { return (&x)[i]; }
    const uint& operator[](int i) const { return (&x)[i]; }
};

struct vec2 {
    union {
        struct { float x, y; };
        struct { float r, g; };
        swz2<vec2, float, 0, 1> xy, rg;
        swz2<vec2, float, 1, 0> yx;
        swz3<vec3, float, 0, 0, 0> xxx, rrr;
        swz4<vec4, float, 0, 0, 0, 0> xxxx;
        swz4<vec4, float, 1, 1, 1, 1> yyyy;
    };
    static const int N = 2;
    vec2() = default;
    vec2(const vec2& o) : x(o.x), y(o.y) {}
    vec2& operator=(const vec2& o) { x = o.x; y = o.y; return *this; }
    template <class S, if_arith<S> = 0> explicit vec2(S s) : x((float)s), y((float)s) {}
    template <class A, class B, if_arith<A> = 0, if_arith<B> = 0> vec2(A a, B b) : x((float)a), y((float)b) {}
    vec2(const uvec2& u) : x((float)u.x), y((float)u.y) {}       // GLSL implicit uint -> float
    vec2(const ivec2& i) : x((float)i.x), y((float)i.y) {}       // GLSL implicit int -> float
    explicit vec2(const bvec2& b) : x(b.x ? 1.0f : 0.0f), y(b.y ? 1.0f : 0.0f) {}
    float& operator[](int i) { return (&x)[i]; }
    const float& operator[](int i) const { return (&x)[i]; }
};
struct vec3 {
    union {
        struct { float x, y, z; };
        struct { float r, g, b; };
        swz2<vec2, float, 0, 1> xy, rg;
        swz2<vec2, float, 1, 2> yz;
        swz2<vec2, float, 0, 2> xz;
        swz3<vec3, float, 0, 1, 2> xyz, rgb;
        swz3<vec3, float, 0, 0, 0> xxx, rrr;
    };
    static const int N = 3;
    vec3() = default;
    vec3(const vec3& o) : x(o.x), y(o.y), z(o.z) {}
    vec3& operator=(const vec3& o) { x = o.x; y = o.y; z = o.z; return *this; }
    template <class S, if_arith<S> = 0> explicit vec3(S s) : x((float)s), y((float)s), z((float)s) {}
    template <class A, class B, class C, if_arith<A> = 0, if_arith<B> = 0, if_arith<C> = 0> vec3(A a, B b, C c) : x((float)a), y((float)b), z((float)c) {}
    vec3(const vec2& v, float z_) : x(v.x), y(v.y), z(z_) {}
    vec3(float x_, const vec2& v) : x(x_), y(v.x), z(v.y) {}
    vec3(const uvec3& u) : x((float)u.x), y((float)u.y), z((float)u.z) {}   // GLSL implicit uint -> float
    explicit vec3(const bvec3& b_) : x(b_.x ? 1.0f : 0.0f), y(b_.y ? 1.0f : 0.0f), z(b_.z ? 1.0f : 0.0f) {}
    float& operator[](int i) { return (&x)[i]; }
    const float& operator[](int i) const { return (&x)[i]; }
};
struct vec4 {
    union {
        struct { float x, y, z, w; };
        struct { float r, g, b, a; };
        swz2<vec2, float, 0, 1> xy, rg;
        swz2<vec2, float, 2, 3> zw, ba;
        swz3<vec3, float, 0, 1, 2> xyz, rgb;
        swz3<vec3, float, 1, 2, 3> yzw;
        swz3<vec3, float, 2, 3, 3> zww;
        swz3<vec3, float, 1, 1, 2> yyz;
        swz3<vec3, float, 0, 0, 0> xxx, rrr;
        swz3<vec3, float, 3, 3, 3> www, aaa;
        swz4<vec4, float, 0, 0, 0, 0> xxxx;
        swz4<vec4, float, 1, 1, 1, 1> yyyy;
    };
    static const int N = 4;
    vec4() = default;
    vec4(const vec4& o) : x(o.x), y(o.y), z(o.z), w(o.w) {}
    vec4& operator=(const vec4& o) { x = o.x; y = o.y; z = o.z; w = o.w; return *this; }
    template <class S, if_arith<S> = 0> explicit vec4(S s) : x((float)s), y((float)s), z((float)s), w((float)s) {}
    template <class A, class B, class C, class D, if_arith<A> = 0, if_arith<B> = 0, if_arith<C> = 0, if_arith<D> = 0>
    vec4(A a_, B b_, C c_, D d_) : x((float)a_), y((float)b_), z((float)c_), w((float)d_) {}
    vec4(const vec3& v, float w_) : x(v.x), y(v.y), z(v.z), w(w_) {}
    vec4(const vec2& p, const vec2& q) : x(p.x), y(p.y), z(q.x), w(q.y) {}
    vec4(const uvec4& u) : x((float)u.x), y((float)u.y), z((float)u.z), w((float)u.w) {}   // GLSL implicit uint -> float
    explicit vec4(const bvec4& b_) : x(b_.x ? 1.0f : 0.0f), y(b_.y ? 1.0f : 0.0f), z(b_.z ? 1.0f : 0.0f), w(b_.w ? 1.0f : 0.0f) {}
    float& operator[](int i) { return (&x)[i]; }
    const float& operator[](int i) const { return (&x)[i]; }
};

inline ivec2::ivec2(const uvec2& u) : x((int)u.x), y((int)u.y) {}
inline uvec2::uvec2(const vec2& v) : x(f2u(v.x)), y(f2u(v.y)) {}
inline uvec3::uvec3(const vec3& v) : x(f2u(v.x)), y(f2u(v.y)), z(f2u(v.z)) {}
inline uvec4::uvec4(const vec2& p, const vec2& q) : x(f2u(p.x)), y(f2u(p.y)), z(f2u(q.x)), w(f2u(q.y)) {}

static_assert(sizeof(vec2) == 8 && sizeof(vec3) == 12 && sizeof(vec4) == 16 && alignof(vec4) == 4, "std430-compatible vectors");
static_assert(sizeof(uvec2) == 8 && sizeof(uvec3) == 12, "std430-compatible vectors");

// ---- component-wise operators (each one a single IEEE operation per component: D1) ---------------------------------------
#define GLSL_ARITH(V, S, OP)                                                                                                  \
    inline V operator OP(const V& a, const V& b) { V r; for (int i = 0; i < V::N; ++i) r[i] = a[i] OP b[i]; return r; }         \
    inline V operator OP(const V& a, S b) { V r; for (int i = 0; i < V::N; ++i) r[i] = a[i] OP b; return r; }                   \
    inline V operator OP(S a, const V& b) { V r; for (int i = 0; i < V::N; ++i) r[i] = a OP b[i]; return r; }                   \
    inline V& operator OP##=(V& a, const V& b) { for (int i = 0; i < V::N; ++i) a[i] = a[i] OP b[i]; return a; }                \
    inline V& operator OP##=(V& a, S b) { for (int i = 0; i < V::N; ++i) a[i] = a[i] OP b; return a; }
#define GLSL_FLOAT_VEC(V) GLSL_ARITH(V, float, +) GLSL_ARITH(V, float, -) GLSL_ARITH(V, float, *) GLSL_ARITH(V, float, /)       \
    inline V operator-(const V& a) { V r; for (int i = 0; i < V::N; ++i) r[i] = -a[i]; return r; }
#define GLSL_UINT_VEC(V) GLSL_ARITH(V, uint, +) GLSL_ARITH(V, uint, -) GLSL_ARITH(V, uint, *) GLSL_ARITH(V, uint, /)            \
    GLSL_ARITH(V, uint, &) GLSL_ARITH(V, uint, |) GLSL_ARITH(V, uint, >>) GLSL_ARITH(V, uint, <<)
GLSL_FLOAT_VEC(vec2) GLSL_FLOAT_VEC(vec3) GLSL_FLOAT_VEC(vec4)
GLSL_UINT_VEC(uvec2) GLSL_UINT_VEC(uvec3) GLSL_UINT_VEC(uvec4)

// ---- scalar built-ins -----------------------------------------------------------------------------------------------------------
inline float sin(float x) { return (float)::sin((double)x); }      // D4
inline float cos(float x) { return (float)::cos((double)x); }
inline float asin(float x) { return (float)::asin((double)x); }
inline float acos(float x) { return (float)::acos((double)x); }
inline float atan(float x) { return (float)::atan((double)x); }
inline float atan(float y, float x) { return (float)::atan2((double)y, (double)x); }
inline float exp(float x) { return (float)::exp((double)x); }
inline float pow(float x, float y) { return (float)::pow((double)x, (double)y); }
inline float sqrt(float x) { return ::sqrtf(x); }                   // D1: correctly rounded binary32
inline float abs(float x) { return ::fabsf(x); }
inline float floor(float x) { return ::floorf(x); }
inline float fract(float x) { return x - ::floorf(x); }
inline float mod(float x, float y) { return x - y * ::floorf(x / y); }
inline float sign(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }
inline float max(float a, float b) { return ::fmaxf(a, b); }
inline float min(float a, float b) { return ::fminf(a, b); }
inline uint max(uint a, uint b) { return a > b ? a : b; }
inline uint min(uint a, uint b) { return a < b ? a : b; }
inline float clamp(float x, float lo, float hi) { return min(max(x, lo), hi); }
inline float mix(float a, float b, float t) { return a * (1.0f - t) + b * t; }
inline float step(float edge, float x) { return x < edge ? 0.0f : 1.0f; }
inline float smoothstep(float e0, float e1, float x) { float t = clamp((x - e0) / (e1 - e0), 0.0f, 1.0f); return t * t * (3.0f - 2.0f * t); }   // D9
inline bool isnan(float x) { return __builtin_isnan(x); }
inline float uintBitsToFloat(uint u) { float f; memcpy(&f, &u, 4); return f; }
inline uint floatBitsToUint(float f) { uint u; memcpy(&u, &f, 4); return u; }
inline vec2 uintBitsToFloat(const uvec2& u) { return vec2(uintBitsToFloat(u.x), uintBitsToFloat(u.y)); }
inline uvec2 floatBitsToUint(const vec2& f) { return uvec2(floatBitsToUint(f.x), floatBitsToUint(f.y)); }

// IEEE binary16 <-> binary32 (unpack exact; pack round-to-nearest-even)
float half_to_float(uint16_t h);
uint16_t float_to_half_rtne(float v);
inline vec2 unpackHalf2x16(uint v) { return vec2(half_to_float((uint16_t)(v & 0xFFFFu)), half_to_float((uint16_t)(v >> 16))); }
inline uint packHalf2x16(const vec2& v) { return (uint)float_to_half_rtne(v.x) | ((uint)float_to_half_rtne(v.y) << 16); }

// ---- vector built-ins -----------------------------------------------------------------------------------------------------------
#define GLSL_MAP1(F, V) inline V F(const V& a) { V r; for (int i = 0; i < V::N; ++i) r[i] = F(a[i]); return r; }
#define GLSL_MAP1_ALL(F) GLSL_MAP1(F, vec2) GLSL_MAP1(F, vec3) GLSL_MAP1(F, vec4)
GLSL_MAP1_ALL(sin) GLSL_MAP1_ALL(cos) GLSL_MAP1_ALL(exp) GLSL_MAP1_ALL(sqrt) GLSL_MAP1_ALL(abs) GLSL_MAP1_ALL(floor) GLSL_MAP1_ALL(fract)
#define GLSL_MAP2(F, V)                                                                                                       \
    inline V F(const V& a, const V& b) { V r; for (int i = 0; i < V::N; ++i) r[i] = F(a[i], b[i]); return r; }                  \
    inline V F(const V& a, float b) { V r; for (int i = 0; i < V::N; ++i) r[i] = F(a[i], b); return r; }
#define GLSL_MAP2_ALL(F) GLSL_MAP2(F, vec2) GLSL_MAP2(F, vec3) GLSL_MAP2(F, vec4)
GLSL_MAP2_ALL(max) GLSL_MAP2_ALL(min) GLSL_MAP2_ALL(mod) GLSL_MAP2_ALL(pow)
#define GLSL_STEP(V) inline V step(const V& e, const V& x) { V r; for (int i = 0; i < V::N; ++i) r[i] = step(e[i], x[i]); return r; }
GLSL_STEP(vec2) GLSL_STEP(vec3) GLSL_STEP(vec4)
#define GLSL_CLAMP_MIX(V)                                                                                                     \
    inline V clamp(const V& x, float lo, float hi) { V r; for (int i = 0; i < V::N; ++i) r[i] = clamp(x[i], lo, hi); return r; } \
    inline V mix(const V& a, const V& b, float t) { V r; for (int i = 0; i < V::N; ++i) r[i] = mix(a[i], b[i], t); return r; }
GLSL_CLAMP_MIX(vec2) GLSL_CLAMP_MIX(vec3) GLSL_CLAMP_MIX(vec4)

inline float dot(const vec2& a, const vec2& b) { return a.x * b.x + a.y * b.y; }                                   // D2
inline float dot(const vec3& a, const vec3& b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline float dot(const vec4& a, const vec4& b) { return ((a.x * b.x + a.y * b.y) + a.z * b.z) + a.w * b.w; }
inline vec3 cross(const vec3& a, const vec3& b) { return vec3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline float length(const vec2& v) { return sqrt(dot(v, v)); }                                                     // D3
inline float length(const vec3& v) { return sqrt(dot(v, v)); }
inline vec2 normalize(const vec2& v) { float inv = 1.0f / sqrt(dot(v, v)); return v * inv; }
inline vec3 normalize(const vec3& v) { float inv = 1.0f / sqrt(dot(v, v)); return v * inv; }
inline vec3 reflect(const vec3& i, const vec3& n) { return i - (2.0f * dot(n, i)) * n; }

#define GLSL_COMPARE(NAME, OP)                                                                                                \
    inline bvec2 NAME(const vec2& a, const vec2& b) { return bvec2{a.x OP b.x, a.y OP b.y}; }                                   \
    inline bvec3 NAME(const vec3& a, const vec3& b) { return bvec3{a.x OP b.x, a.y OP b.y, a.z OP b.z}; }                       \
    inline bvec4 NAME(const vec4& a, const vec4& b) { return bvec4{a.x OP b.x, a.y OP b.y, a.z OP b.z, a.w OP b.w}; }           \
    inline bvec2 NAME(const uvec2& a, const uvec2& b) { return bvec2{a.x OP b.x, a.y OP b.y}; }                                 \
    inline bvec2 NAME(const ivec2& a, const ivec2& b) { return bvec2{a.x OP b.x, a.y OP b.y}; }
GLSL_COMPARE(lessThan, <) GLSL_COMPARE(lessThanEqual, <=) GLSL_COMPARE(greaterThan, >) GLSL_COMPARE(greaterThanEqual, >=)
GLSL_COMPARE(equal, ==) GLSL_COMPARE(notEqual, !=)
inline bool any(const bvec2& b) { return b.x || b.y; }
inline bool any(const bvec3& b) { return b.x || b.y || b.z; }
inline bool all(const bvec2& b) { return b.x && b.y; }
inline bool all(const bvec3& b) { return b.x && b.y && b.z; }
inline bvec3 isnan(const vec3& v) { return bvec3{isnan(v.x), isnan(v.y), isnan(v.z)}; }

// ---- opaque types (bound by the harness) ----------------------------------------------------------------------------------------
enum { FMT_RGBA32F = 0, FMT_RGBA16F = 1, FMT_RGBA8 = 2 };
struct sampler2D { const void* data; int w, h, fmt, linear; };   // linear == 0: nearest, clamp to edge; 1: bilinear, border (0,0,0,0)
struct sampler2DMS { const void* data; int w, h, samples; };
struct image2D { void* data; int w, h, fmt; };

vec4 texture(const sampler2D& s, const vec2& uv);
inline ivec2 textureSize(const sampler2D& s, int) { return ivec2(s.w, s.h); }
inline int textureSamples(const sampler2DMS& s) { return s.samples; }
vec4 texelFetch(const sampler2DMS& s, const ivec2& p, int sample);
void imageStore(const image2D& img, const ivec2& p, const vec4& v);
vec4 imageLoad(const image2D& img, const ivec2& p);

// ---- invocation state (set by the harness before each invocation / fibre switch) ---------------------------------------------
extern thread_local uvec3 gl_GlobalInvocationID;
extern thread_local uint gl_SubGroupInvocationARB;
uint64_t ballotARB(bool v);   // ARB_shader_ballot over the 32 invocations the harness runs in lock-step

}  // namespace glsl
#endif
