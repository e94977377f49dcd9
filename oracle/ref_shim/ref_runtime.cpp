// ref_runtime.cpp — what the GLSL shim needs at run time: half conversions, texture / image access with the decreed
// semantics (D6, D7), the lock-step subgroup scheduler behind ballotARB, and a small parallel loop.
// TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).  Ours; contains no reference code.
#include "glsl_shim.h"
#include "ref_runtime.h"

#include <ucontext.h>

#include <atomic>
#include <cstdlib>
#include <memory>
#include <thread>
#include <vector>

namespace glsl {

thread_local uvec3 gl_GlobalInvocationID;
thread_local uint gl_SubGroupInvocationARB;

float half_to_float(uint16_t h) {
    uint32_t sign = (uint32_t)(h >> 15) << 31, e = (h >> 10) & 0x1Fu, m = h & 0x3FFu;
    if (e == 0) {
        if (m == 0) return uintBitsToFloat(sign);
        float f = (float)m * 5.9604644775390625e-8f;   // m * 2^-24, exact
        return sign ? -f : f;
    }
    if (e == 31) return uintBitsToFloat(sign | 0x7F800000u | (m << 13));
    return uintBitsToFloat(sign | ((e + 112u) << 23) | (m << 13));
}

uint16_t float_to_half_rtne(float v) {   // D6
    uint32_t b = floatBitsToUint(v);
    uint32_t sign = (b >> 16) & 0x8000u, a = b & 0x7FFFFFFFu;
    if (a > 0x7F800000u) return (uint16_t)0x7FFFu;
    if (a >= 0x477FF000u) return (uint16_t)(sign | 0x7C00u);
    if (a < 0x33000001u) return (uint16_t)sign;
    int32_t e = (int32_t)(a >> 23) - 127;
    uint32_t m = (a & 0x7FFFFFu) | 0x800000u;
    if (e < -14) {
        uint32_t shift = (uint32_t)(-14 - e) + 13u;
        uint32_t q = m >> shift, rem = m & ((1u << shift) - 1u), half = 1u << (shift - 1u);
        if (rem > half || (rem == half && (q & 1u))) q++;
        return (uint16_t)(sign | q);
    }
    uint32_t q = ((uint32_t)(e + 15) << 10) | ((m >> 13) & 0x3FFu);
    uint32_t rem = m & 0x1FFFu;
    if (rem > 0x1000u || (rem == 0x1000u && (q & 1u))) q++;
    return (uint16_t)(sign | q);
}

static vec4 fetch(const sampler2D& s, int64_t x, int64_t y) {
    if (s.fmt == FMT_RGBA16F) {
        const uint16_t* p = (const uint16_t*)s.data + 4 * ((size_t)y * s.w + (size_t)x);
        return vec4(half_to_float(p[0]), half_to_float(p[1]), half_to_float(p[2]), half_to_float(p[3]));
    }
    vec4 r;
    memcpy(&r, (const float*)s.data + 4 * ((size_t)y * s.w + (size_t)x), 16);
    return r;
}

vec4 texture(const sampler2D& s, const vec2& uv) {
    if (!s.linear) {   // D7: nearest, clamp to edge
        int64_t x = (int64_t)::floorf(uv.x * (float)s.w), y = (int64_t)::floorf(uv.y * (float)s.h);
        x = x < 0 ? 0 : (x >= s.w ? s.w - 1 : x);
        y = y < 0 ? 0 : (y >= s.h ? s.h - 1 : y);
        return fetch(s, x, y);
    }
    // D7: bilinear, binary32 weights, clamp to a (0,0,0,0) border
    float fx = uv.x * (float)s.w - 0.5f, fy = uv.y * (float)s.h - 0.5f;
    float x0f = ::floorf(fx), y0f = ::floorf(fy);
    float ax = fx - x0f, ay = fy - y0f;
    if (isnan(fx) || isnan(fy)) { float q = uintBitsToFloat(0x7FC00000u); return vec4(q, q, q, q); }
    int64_t x0 = (int64_t)x0f, y0 = (int64_t)y0f;
    auto tap = [&](int64_t x, int64_t y) { return (x < 0 || y < 0 || x >= s.w || y >= s.h) ? vec4(0.0f, 0.0f, 0.0f, 0.0f) : fetch(s, x, y); };
    vec4 t00 = tap(x0, y0), t10 = tap(x0 + 1, y0), t01 = tap(x0, y0 + 1), t11 = tap(x0 + 1, y0 + 1);
    vec4 top = t00 * (1.0f - ax) + t10 * ax;
    vec4 bot = t01 * (1.0f - ax) + t11 * ax;
    return top * (1.0f - ay) + bot * ay;
}

vec4 texelFetch(const sampler2DMS& s, const ivec2& p, int sample) {
    if (!s.data || p.x < 0 || p.y < 0 || p.x >= s.w || p.y >= s.h) return vec4(0.0f, 0.0f, 0.0f, 0.0f);
    vec4 r;
    memcpy(&r, (const float*)s.data + 4 * (((size_t)p.y * s.w + (size_t)p.x) * s.samples + sample), 16);
    return r;
}

static uint32_t unorm8(float c) {   // D6
    if (isnan(c)) return 0;
    c = c < 0.0f ? 0.0f : (c > 1.0f ? 1.0f : c);
    return (uint32_t)::floorf(c * 255.0f + 0.5f);
}

void imageStore(const image2D& img, const ivec2& p, const vec4& v) {
    if (p.x < 0 || p.y < 0 || p.x >= img.w || p.y >= img.h) return;   // out-of-bounds image stores are discarded
    size_t i = (size_t)p.y * img.w + (size_t)p.x;
    if (img.fmt == FMT_RGBA32F) memcpy((float*)img.data + 4 * i, &v, 16);
    else if (img.fmt == FMT_RGBA16F) {
        uint16_t* d = (uint16_t*)img.data + 4 * i;
        d[0] = float_to_half_rtne(v.x); d[1] = float_to_half_rtne(v.y); d[2] = float_to_half_rtne(v.z); d[3] = float_to_half_rtne(v.w);
    } else
        ((uint32_t*)img.data)[i] = unorm8(v.x) | (unorm8(v.y) << 8) | (unorm8(v.z) << 16) | (unorm8(v.w) << 24);
}

vec4 imageLoad(const image2D& img, const ivec2& p) {
    vec4 r(0.0f, 0.0f, 0.0f, 0.0f);
    if (p.x < 0 || p.y < 0 || p.x >= img.w || p.y >= img.h) return r;
    memcpy(&r, (const float*)img.data + 4 * ((size_t)p.y * img.w + (size_t)p.x), 16);
    return r;
}

// ---- lock-step subgroups ----------------------------------------------------------------------------------------------------
// One fibre per invocation of a 32-wide subgroup.  A fibre runs until it returns or reaches ballotARB; when every live
// fibre waits at the ballot the scheduler forms the mask from the waiting (= active) invocations and resumes them.
namespace {
enum { READY = 0, WAITING = 1, DONE = 2 };
const size_t STACK_BYTES = 256 * 1024;
struct Warp {
    ucontext_t sched, ctx[32];
    char* stacks = nullptr;
    int state[32];
    bool vote[32];
    uvec3 gid[32];
    uint64_t result = 0;
    int cur = 0;
    void (*entry)() = nullptr;
    ~Warp() { free(stacks); }
};
thread_local Warp* t_warp = nullptr;    // the running subgroup (null outside run_subgroup)
thread_local std::unique_ptr<Warp> t_pool;

void fibre_main() {
    Warp* w = t_warp;
    w->entry();
    w->state[w->cur] = DONE;            // uc_link returns to the scheduler
}
}  // namespace

uint64_t ballotARB(bool v) {
    Warp* w = t_warp;
    if (!w) return v ? 1u : 0u;
    int l = w->cur;
    w->vote[l] = v;
    w->state[l] = WAITING;
    swapcontext(&w->ctx[l], &w->sched);
    return t_warp->result;
}

void run_subgroup(void (*entry)(), const uvec3 ids[32]) {
    if (!t_pool) { t_pool.reset(new Warp); t_pool->stacks = (char*)malloc(32 * STACK_BYTES); }
    Warp* w = t_pool.get();
    w->entry = entry;
    for (int l = 0; l < 32; ++l) {
        w->gid[l] = ids[l];
        w->state[l] = READY;
        getcontext(&w->ctx[l]);
        w->ctx[l].uc_stack.ss_sp = w->stacks + (size_t)l * STACK_BYTES;
        w->ctx[l].uc_stack.ss_size = STACK_BYTES;
        w->ctx[l].uc_link = &w->sched;
        makecontext(&w->ctx[l], fibre_main, 0);
    }
    t_warp = w;
    for (;;) {
        for (int l = 0; l < 32; ++l)
            if (w->state[l] == READY) {
                w->cur = l;
                gl_GlobalInvocationID = w->gid[l];
                gl_SubGroupInvocationARB = (uint)l;
                swapcontext(&w->sched, &w->ctx[l]);
            }
        uint64_t mask = 0;
        int waiting = 0;
        for (int l = 0; l < 32; ++l)
            if (w->state[l] == WAITING) { ++waiting; if (w->vote[l]) mask |= 1ull << l; }
        if (!waiting) break;
        w->result = mask;
        for (int l = 0; l < 32; ++l)
            if (w->state[l] == WAITING) w->state[l] = READY;
    }
    t_warp = nullptr;
}

static int g_threads = 0;
void set_threads(int n) { g_threads = n; }

void parallel_for(int64_t n, const std::function<void(int64_t)>& body) {
    int nt = g_threads > 0 ? g_threads : (int)std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    if (nt == 1 || n <= 1) { for (int64_t i = 0; i < n; ++i) body(i); return; }
    std::atomic<int64_t> next{0};
    auto worker = [&]() { for (;;) { int64_t i = next.fetch_add(1); if (i >= n) break; body(i); } };
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(worker);
    worker();
    for (auto& t : pool) t.join();
}

}  // namespace glsl

extern "C" void ref_set_threads(int n) { glsl::set_threads(n); }
