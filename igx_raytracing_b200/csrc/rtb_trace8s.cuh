// rtb_trace8s.cuh — spheres and cubes through their own 8-wide trees (included by rtb_kernels.cu after rtb_trace8.cuh).
//
// The reference tests every sphere and every cube against every ray (SH/trace.glsl:31-40, :83-90) and allows 32768 spheres and
// 16384 cubes (igx/include/helpers/scene_graph.hpp:137-142); in round 1 they stayed in those loops.  With RTB_OPT_PRIMITIVE_TREES
// a type with many primitives gets a tree of its own (built on the device over the primitives' boxes: rtb_build.cu through proxy
// triangles) and the loops of that type are replaced by a traversal that returns what the loop returns:
//
//   spheres, nearest   the sphere with the smallest candidate distance BELOW the distance the triangles left (strict <, as
//                      rayIntersectSphere accepts), lowest index on equal distances (the loop's first-wins);
//   cubes, nearest     among the cubes the slab test accepts (tmax >= 0, tmin <= tmax) with tmin <= the distance left by triangles
//                      and spheres, the smallest tmin, HIGHEST index on equal tmin — rayIntersectCube accepts tmin == hitT, so in
//                      the loop a later cube replaces an earlier one at the same distance — negative tmin (origin inside) included;
//   any hit            some primitive other than the ray's own object with a candidate distance < maxDist: the loops only ever
//                      lower hitT, so "hitT < maxDist at the end" decomposes per type (rtb_kernels.cu, occludedByOthers).
//
// The winner's normal and uv are then computed by the reference's own function on that one primitive (finishGeometry).  One ray
// per thread, a per-thread stack of node groups in local memory: these launches are not the hot path of any BASELINE workload,
// they remove an O(primitives) loop per ray.
#pragma once

namespace rtb {

// PrimHit (rtb_kernels.cuh): t, id = index within the type, NO_RAY_HIT = none

// slab distances of rayIntersectCube (SH/primitive.glsl:286-300), the same operations in the same order
RTB_DI bool cubeCandidate(const Ray& r, const float* cube, float& tmin) {
    const vec3 revDir = mk3(1.0f / r.dir.x, 1.0f / r.dir.y, 1.0f / r.dir.z);
    const vec3 start = mk3(cube[0], cube[1], cube[2]), end = mk3(cube[3], cube[4], cube[5]);
    const vec3 startDir = (start - r.pos) * revDir, endDir = (end - r.pos) * revDir;
    const vec3 mi = vmin(startDir, endDir), ma = vmax(startDir, endDir);
    tmin = fmaxf(fmaxf(mi.x, mi.y), mi.z);
    const float tmax = fminf(fminf(ma.x, ma.y), ma.z);
    return !(tmax < 0.0f || tmin > tmax);
}

struct PrimTraceArgs {
    const RayRec* rays; uint32_t n; const uint32_t* countPtr;
    const uint4* nodes8; const uint4* nodes8Alias; const float4* tt;   // the type's tree; tt: 48-byte records, .w of the first float4 = primitive index
    const float4* spheres; const float* cubes; uint32_t firstObject;    // global object id of the type's primitive 0
    const TriHit* triHits; const PrimHit* before;   // nearest: the distance left by the triangles and by the type searched before (may be null)
    PrimHit* out;                                    // nearest
    uint8_t* bytes; uint32_t* bits; const uint32_t* slotIds; FrameMap fm;   // any hit: OR into bytes (rays-in) or into the shadow words
};

template <int KIND, bool ANY>   // KIND 0 spheres, 1 cubes
__global__ void __launch_bounds__(128) k_trace_prims(const PrimTraceArgs a) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n = a.countPtr ? __ldg(a.countPtr) : a.n;
    if (r >= n) return;
    const float4 o = __ldg(reinterpret_cast<const float4*>(a.rays + r)), d = __ldg(reinterpret_cast<const float4*>(a.rays + r) + 1);
    if (!ANY) a.out[r] = PrimHit{NO_HIT, NO_RAY_HIT};
    if (!(d.w >= 0.0f)) return;   // a dead slot
    Ray ray; ray.pos = mk3(o.x, o.y, o.z); ray.dir = mk3(d.x, d.y, d.z);
    const uint32_t prev = fbits(o.w);
    // the distance the earlier stages left: candidates must undercut it (spheres: strictly; cubes: or equal)
    float best = d.w;
    if (!ANY) {
        if (a.triHits) { const float t = __ldg(&a.triHits[r].t); if (__ldg(&a.triHits[r].id) != NO_RAY_HIT) best = fminf(best, t); }
        if (a.before) { const PrimHit b = a.before[r]; if (b.id != NO_RAY_HIT) best = fminf(best, b.t); }
    }
    uint32_t bestId = NO_RAY_HIT;
    bool occluded = false;
    const float tiny = 8.271806e-25f;
    const float idx = 1.0f / (fabsf(ray.dir.x) > tiny ? ray.dir.x : copysignf(tiny, ray.dir.x));
    const float idy = 1.0f / (fabsf(ray.dir.y) > tiny ? ray.dir.y : copysignf(tiny, ray.dir.y));
    const float idz = 1.0f / (fabsf(ray.dir.z) > tiny ? ray.dir.z : copysignf(tiny, ray.dir.z));
    const uint32_t octinv = (idx < 0.0f ? 0u : 1u) | (idy < 0.0f ? 0u : 2u) | (idz < 0.0f ? 0u : 4u);
    const uint32_t negX = ~octinv & 1u, negY = ~octinv & 2u, negZ = ~octinv & 4u;
    uint2 stack[64];
    int sp = 0;
    uint2 G = make_uint2(0u, 0x80000000u);   // the root as a one-node group
    for (;;) {
        if (!(G.y & 0xFF000000u)) {
            if (sp == 0) break;
            G = stack[--sp];
            continue;
        }
        const uint32_t hits = G.y;
        const uint32_t bit = 31u - (uint32_t)__clz(hits);
        const uint32_t childSlot = (bit - 24u) ^ octinv;
        const uint32_t nodeIdx = G.x + (uint32_t)__popc(hits & 0xFFu & ~(0xFFFFFFFFu << childSlot));
        G.y &= ~(1u << bit);
        if ((G.y & 0xFF000000u) && sp < 64) stack[sp++] = G;
        uint4 n0, n1, wnx, wny, wnz, wfx, wfy, wfz;
        const char* p = reinterpret_cast<const char*>(a.nodes8) + (size_t)nodeIdx * 128u;
        const char* p2 = reinterpret_cast<const char*>(a.nodes8Alias) + (size_t)nodeIdx * 128u;
        ldg256(p, n0, n1);
        ldg256swap(p + 32, p2 + 32, negX, wnx, wfx);
        ldg256swap(p + 64, p2 + 64, negY, wny, wfy);
        ldg256swap(p + 96, p2 + 96, negZ, wnz, wfz);
        const float kx = __uint_as_float((n0.w & 0xFFu) << 23) * idx, ky = __uint_as_float((n0.w << 15) & 0x7F800000u) * idy, kz = __uint_as_float((n0.w << 7) & 0x7F800000u) * idz;
        const float cx = (__uint_as_float(n0.x) - ray.pos.x) * idx, cy = (__uint_as_float(n0.y) - ray.pos.y) * idy, cz = (__uint_as_float(n0.z) - ray.pos.z) * idz;
        // the limit of the box test: a cube may be accepted AT the current distance, and its slab arithmetic (a multiplication by
        // 1 / d) differs from the box test's by rounding: a little slack on the limit keeps every acceptable primitive reachable
        // (a negative best — a cube entered from inside — must not prune the nodes that contain the origin: a cube with a still
        // smaller tmin would be accepted by the loop)
        const float lim0 = fmaxf(best, 0.0f);
        const float limit = lim0 < NO_HIT ? lim0 + lim0 * 1e-5f + 1e-30f : lim0;
        uint32_t hitmask = 0;
        testPair<0>(wnx.x, wny.x, wnz.x, wfx.x, wfy.x, wfz.x, kx, ky, kz, cx, cy, cz, limit, hitmask);
        testPair<2>(wnx.y, wny.y, wnz.y, wfx.y, wfy.y, wfz.y, kx, ky, kz, cx, cy, cz, limit, hitmask);
        testPair<4>(wnx.z, wny.z, wnz.z, wfx.z, wfy.z, wfz.z, kx, ky, kz, cx, cy, cz, limit, hitmask);
        testPair<6>(wnx.w, wny.w, wnz.w, wfx.w, wfy.w, wfz.w, kx, ky, kz, cx, cy, cz, limit, hitmask);
        hitmask &= n1.z;
        uint32_t top = hitmask >> 24;
        if (octinv & 1u) top = ((top & 0x55u) << 1) | ((top >> 1) & 0x55u);
        if (octinv & 2u) top = ((top & 0x33u) << 2) | ((top >> 2) & 0x33u);
        if (octinv & 4u) top = ((top & 0x0Fu) << 4) | (top >> 4);
        const uint32_t P = n1.z & 0x00FFFFFFu;
        uint32_t T = hitmask & 0x00FFFFFFu;
        G = make_uint2(n1.x, (top << 24) | (n0.w >> 24));
        while (T) {
            const uint32_t tb = 31u - (uint32_t)__clz(T);
            T &= ~(1u << tb);
            const uint32_t id = fbits(__ldg(a.tt + (size_t)(n1.y + (uint32_t)__popc(P & ~(0xFFFFFFFFu << tb))) * 3).w);
            if (a.firstObject + id == prev) continue;   // the ray's own object (obj == prevObj in the reference's tests)
            if (KIND == 0) {
                const float t = sphereCandidateT(ray, __ldg(a.spheres + id));   // NO_HIT when the sphere is not hit in front
                if (ANY) { if (t < best) occluded = true; }
                else if (t < best || (t == best && bestId != NO_RAY_HIT && id < bestId)) { best = t; bestId = id; }
            } else {
                float c[6];
                const float2* cp = reinterpret_cast<const float2*>(a.cubes + 6 * (size_t)id);
                const float2 c0 = __ldg(cp), c1 = __ldg(cp + 1), c2 = __ldg(cp + 2);
                c[0] = c0.x; c[1] = c0.y; c[2] = c1.x; c[3] = c1.y; c[4] = c2.x; c[5] = c2.y;
                float tmin;
                if (!cubeCandidate(ray, c, tmin)) continue;
                if (ANY) { if (tmin < best) occluded = true; }
                else if (tmin < best || (tmin == best && (bestId == NO_RAY_HIT || id > bestId))) { best = tmin; bestId = id; }
            }
        }
        if (ANY && occluded) break;
    }
    if (!ANY) { a.out[r] = PrimHit{best, bestId}; return; }
    if (!occluded) return;
    if (a.bytes) a.bytes[r] = 1;
    if (a.bits) {
        const uint32_t j = a.slotIds ? __ldg(a.slotIds + r) : r;
        const uint32_t layer = j / a.fm.localSlots, i = j - layer * a.fm.localSlots;
        uint32_t x, y;
        slotToPixel(a.fm, i, x, y);
        atomicOr(a.bits + indexToLight(x, y, a.fm.w, a.fm.h, layer), 1u << ((x & 15u) | ((y & 1u) << 4)));
    }
}

// boxes of spheres / cubes as degenerate "triangles" (p0 = lower corner, p1 = upper corner, p2 = lower corner), so that the device
// builder and the refit kernels — which box triangles — build and maintain these trees unchanged
__global__ void __launch_bounds__(256) k_proxy_triangles(int kind, const float4* __restrict__ spheres, const float* __restrict__ cubes, uint32_t n, TriangleRec* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float lo[3], hi[3];
    if (kind == 0) {
        const float4 s = spheres[i];
        const float r = fabsf(s.w);
        lo[0] = s.x - r; lo[1] = s.y - r; lo[2] = s.z - r; hi[0] = s.x + r; hi[1] = s.y + r; hi[2] = s.z + r;
    } else {
        for (int a = 0; a < 3; ++a) { const float u = cubes[6 * (size_t)i + a], v = cubes[6 * (size_t)i + 3 + a]; lo[a] = fminf(u, v); hi[a] = fmaxf(u, v); }
    }
    TriangleRec t;
    for (int a = 0; a < 3; ++a) { t.p0[a] = lo[a]; t.p1[a] = hi[a]; t.p2[a] = lo[a]; }
    t.n0 = t.n1 = t.n2 = 0u;
    out[i] = t;
}
void launch_proxy_triangles(int kind, const float4* spheres, const float* cubes, uint32_t n, TriangleRec* out, cudaStream_t st) {
    if (n) k_proxy_triangles<<<(n + 255u) / 256u, 256, 0, st>>>(kind, spheres, cubes, n, out);
}

void launch_trace_prims(int kind, bool any, const PrimTraceArgs& a, cudaStream_t st) {
    if (!a.n) return;
    const uint32_t g = (a.n + 127u) / 128u;
    if (kind == 0) { if (any) k_trace_prims<0, true><<<g, 128, 0, st>>>(a); else k_trace_prims<0, false><<<g, 128, 0, st>>>(a); }
    else { if (any) k_trace_prims<1, true><<<g, 128, 0, st>>>(a); else k_trace_prims<1, false><<<g, 128, 0, st>>>(a); }
}

}  // namespace rtb
