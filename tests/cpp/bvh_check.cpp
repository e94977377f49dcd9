// bvh_check.cpp — host-only check of the BVH builder (no GPU): builds the tree over a random soup and walks it on the
// CPU with the same slab test, link encoding and tie rule the traversal kernel uses, against a linear loop.
// Prints "OK <nodes> <leaves> <depth> <sah> <ms>" or a diagnostic and a non-zero exit code.
//   g++ -O2 -std=c++17 -pthread -I igx_raytracing_b200/csrc tests/cpp/bvh_check.cpp igx_raytracing_b200/csrc/rtb_bvh.cpp
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "rtb_bvh.h"

using namespace rtb;

static uint64_t s_state = 0x1234567ull;
static float rnd() { s_state = s_state * 6364136223846793005ull + 1442695040888963407ull; return (float)((s_state >> 40) & 0xFFFFFF) / 16777216.0f; }

struct V { float x, y, z; };
static V sub(V a, V b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static V cross(V a, V b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
static float dot(V a, V b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }

static bool tri(V ro, V rd, V p0, V e1, V e2, float& t) {
    V h = cross(rd, e2);
    float a = dot(e1, h), f = 1.0f / a;
    V s = sub(ro, p0);
    float u = f * dot(s, h);
    if (u < 0.0f || u > 1.0f) return false;
    V q = cross(s, e1);
    float v = f * dot(rd, q);
    if (v < 0.0f || u + v > 1.0f) return false;
    t = f * dot(e2, q);
    return t > 0.0f;
}


// ---- CPU mirror of k_trace_cwbvh (rtb_trace8.cuh): same octant ordering, bf16 plane decode (upper half undecoded), fma planes ----
static uint32_t fbitsU(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }

static bool walk8(const std::vector<Node8>& nodes, const std::vector<TravTri>& tt, V ro, V rd, float& best, uint32_t& bestId, int& maxSp) {
    const float tiny = 8.271806e-25f;
    const float idx = 1.0f / (std::fabs(rd.x) > tiny ? rd.x : std::copysign(tiny, rd.x));
    const float idy = 1.0f / (std::fabs(rd.y) > tiny ? rd.y : std::copysign(tiny, rd.y));
    const float idz = 1.0f / (std::fabs(rd.z) > tiny ? rd.z : std::copysign(tiny, rd.z));
    const uint32_t octinv = (idx < 0 ? 0u : 1u) | (idy < 0 ? 0u : 2u) | (idz < 0 ? 0u : 4u), octinv4 = octinv * 0x01010101u;
    best = 3.4028235e38f; bestId = 0xFFFFFFFFu;
    struct G2 { uint32_t x, y; };
    G2 stack[64]; int sp = 0; G2 G = {0u, 0x80000000u};
    for (;;) {
        G2 T = {0u, 0u}; uint32_t P = 0;
        if (G.y & 0xFF000000u) {
            const uint32_t hits = G.y;
            const uint32_t bit = 31u - (uint32_t)__builtin_clz(hits);
            const uint32_t childSlot = (bit - 24u) ^ octinv;
            const uint32_t rel = (uint32_t)__builtin_popcount(hits & 0xFFu & ~(0xFFFFFFFFu << childSlot));
            const uint32_t ni = G.x + rel;
            G.y &= ~(1u << bit);
            if (G.y & 0xFF000000u) { if (sp >= 64) return false; stack[sp++] = G; }
            if (ni >= nodes.size()) return false;
            const Node8& n = nodes[ni];
            auto asf = [](uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; };
            const float kx = asf((uint32_t)n.e[0] << 23) * idx, ky = asf((uint32_t)n.e[1] << 23) * idy, kz = asf((uint32_t)n.e[2] << 23) * idz;
            const float cx = (n.p[0] - ro.x) * idx, cy = (n.p[1] - ro.y) * idy, cz = (n.p[2] - ro.z) * idz;
            uint32_t hitmask = 0;
            for (int word = 0; word < 4; ++word) {
                const uint32_t nx = idx < 0 ? n.hi(0)[word] : n.lo(0)[word], fx = idx < 0 ? n.lo(0)[word] : n.hi(0)[word];
                const uint32_t ny = idy < 0 ? n.hi(1)[word] : n.lo(1)[word], fy = idy < 0 ? n.lo(1)[word] : n.hi(1)[word];
                const uint32_t nz = idz < 0 ? n.hi(2)[word] : n.lo(2)[word], fz = idz < 0 ? n.lo(2)[word] : n.hi(2)[word];
                for (int half = 0; half < 2; ++half) {   // half 0: upper 16 bits read as the whole word; half 1: lower 16 bits shifted up
                    const int sl = 2 * word + half, sh = half ? 16 : 0;
                    const float tnx = std::fma(asf(nx << sh), kx, cx), tny = std::fma(asf(ny << sh), ky, cy), tnz = std::fma(asf(nz << sh), kz, cz);
                    const float tfx = std::fma(asf(fx << sh), kx, cx), tfy = std::fma(asf(fy << sh), ky, cy), tfz = std::fma(asf(fz << sh), kz, cz);
                    const float cmin = std::fmax(std::fmax(tnx, tny), std::fmax(tnz, 0.0f)), cmax = std::fmin(std::fmin(tfx, tfy), std::fmin(tfz, best));
                    if (cmin <= cmax) hitmask |= (1u << (24 + sl)) | (7u << (3 * sl));
                }
            }
            hitmask &= n.valid;
            uint32_t top = hitmask >> 24;
            if (octinv & 1u) top = ((top & 0x55u) << 1) | ((top >> 1) & 0x55u);
            if (octinv & 2u) top = ((top & 0x33u) << 2) | ((top >> 2) & 0x33u);
            if (octinv & 4u) top = ((top & 0x0Fu) << 4) | (top >> 4);
            hitmask = (hitmask & 0x00FFFFFFu) | (top << 24);
            P = n.valid & 0x00FFFFFFu;
            G = {n.childBase, (hitmask & 0xFF000000u) | n.imask};
            T = {n.triBase, hitmask & 0x00FFFFFFu};
        } else { return false; /* the CPU mirror never postpones */ }
        while (T.y) {
            const uint32_t bit = 31u - (uint32_t)__builtin_clz(T.y);
            T.y &= ~(1u << bit);
            const uint32_t ti = T.x + (uint32_t)__builtin_popcount(P & ~(0xFFFFFFFFu << bit));
            if (ti >= tt.size()) return false;
            const TravTri& t = tt[ti];
            float tv;
            if (tri(ro, rd, V{t.p0[0], t.p0[1], t.p0[2]}, V{t.e1[0], t.e1[1], t.e1[2]}, V{t.e2[0], t.e2[1], t.e2[2]}, tv))
                if (tv < best || (tv == best && t.id < bestId)) { best = tv; bestId = t.id; }
        }
        maxSp = std::max(maxSp, sp);
        if (!(G.y & 0xFF000000u)) { if (sp == 0) break; G = stack[--sp]; }
    }
    (void)fbitsU;
    return true;
}

int main(int argc, char** argv) {
    const uint32_t n = argc > 1 ? (uint32_t)std::atoi(argv[1]) : 20000;
    const int threads = argc > 2 ? std::atoi(argv[2]) : 0;
    const uint32_t nRays = argc > 3 ? (uint32_t)std::atoi(argv[3]) : 2000;
    std::vector<TriangleRec> tris(n);
    for (auto& t : tris) {
        std::memset(&t, 0, sizeof t);
        const float c[3] = {rnd() * 20 - 10, rnd() * 20 - 10, rnd() * 20 - 10};
        for (int a = 0; a < 3; ++a) { t.p0[a] = c[a] + rnd() * 0.4f - 0.2f; t.p1[a] = c[a] + rnd() * 0.4f - 0.2f; t.p2[a] = c[a] + rnd() * 0.4f - 0.2f; }
    }
    // a few exact duplicates: equal t must resolve to the lower index
    for (uint32_t i = 0; i + 1 < n && i < 64; i += 2) tris[n - 1 - i / 2] = tris[i];
    std::vector<BvhNode> nodes; std::vector<TravTri> tt; BvhStats st;
    buildBvh(tris.data(), n, 256, threads, nodes, tt, st);
    if (tt.size() != n) { std::printf("FAIL travTris %zu\n", tt.size()); return 1; }
    std::vector<uint8_t> seen(n, 0);
    for (auto& t : tt) { if (t.id >= n || seen[t.id]) { std::printf("FAIL permutation\n"); return 1; } seen[t.id] = 1; }
    if (st.maxDepth > BVH_MAX_DEPTH) { std::printf("FAIL depth %u\n", st.maxDepth); return 1; }

    std::vector<Node8> nodes8; std::vector<TravTri> tt8; BvhStats st8;
    buildCwbvh(tris.data(), n, threads, nodes8, tt8, st8);
    if (tt8.size() != n) { std::printf("FAIL cwbvh travTris %zu\n", tt8.size()); return 1; }
    { std::vector<uint8_t> seen8(n, 0); for (auto& t : tt8) { if (t.id >= n || seen8[t.id]) { std::printf("FAIL cwbvh permutation\n"); return 1; } seen8[t.id] = 1; } }
    int maxSp = 0;
    uint32_t hits = 0;
    for (uint32_t r = 0; r < nRays; ++r) {
        V ro = {rnd() * 24 - 12, rnd() * 24 - 12, 30.0f};
        V rd = {rnd() * 0.6f - 0.3f, rnd() * 0.6f - 0.3f, -1.0f};
        if (r % 7 == 0) { ro = {tris[r % n].p0[0], tris[r % n].p0[1], 30.0f}; rd = {0.0f, 0.0f, -1.0f}; }   // axis-parallel
        const float il = 1.0f / std::sqrt(dot(rd, rd)); rd = {rd.x * il, rd.y * il, rd.z * il};
        // linear reference
        float best = 3.4028235e38f; uint32_t bestId = 0xFFFFFFFFu;
        for (uint32_t i = 0; i < n; ++i) {
            V p0 = {tris[i].p0[0], tris[i].p0[1], tris[i].p0[2]};
            V e1 = sub(V{tris[i].p1[0], tris[i].p1[1], tris[i].p1[2]}, p0), e2 = sub(V{tris[i].p2[0], tris[i].p2[1], tris[i].p2[2]}, p0);
            float t;
            if (tri(ro, rd, p0, e1, e2, t) && t < best) { best = t; bestId = i; }
        }
        // BVH walk
        const float tiny = 8.271806e-25f;
        const float idx = 1.0f / (std::fabs(rd.x) > tiny ? rd.x : std::copysign(tiny, rd.x));
        const float idy = 1.0f / (std::fabs(rd.y) > tiny ? rd.y : std::copysign(tiny, rd.y));
        const float idz = 1.0f / (std::fabs(rd.z) > tiny ? rd.z : std::copysign(tiny, rd.z));
        const float ox = ro.x * idx, oy = ro.y * idy, oz = ro.z * idz;
        float b2 = 3.4028235e38f; uint32_t id2 = 0xFFFFFFFFu;
        int stack[64]; int sp = 0; stack[0] = 0x7FFFFFFF; int node = 0;
        while (node != 0x7FFFFFFF) {
            if (node >= 0) {
                const BvhNode& nd = nodes[(size_t)node];
                auto slab = [&](float lox, float hix, float loy, float hiy, float loz, float hiz, float& tmin) {
                    const float ax = std::fma(lox, idx, -ox), bx = std::fma(hix, idx, -ox), ay = std::fma(loy, idy, -oy), by = std::fma(hiy, idy, -oy);
                    const float az = std::fma(loz, idz, -oz), bz = std::fma(hiz, idz, -oz);
                    tmin = std::fmax(std::fmax(std::fmin(ax, bx), std::fmin(ay, by)), std::fmax(std::fmin(az, bz), 0.0f));
                    const float tmax = std::fmin(std::fmin(std::fmax(ax, bx), std::fmax(ay, by)), std::fmin(std::fmax(az, bz), b2));
                    return tmax >= tmin;
                };
                float m0, m1;
                const bool h0 = slab(nd.c0lox, nd.c0hix, nd.c0loy, nd.c0hiy, nd.c0loz, nd.c0hiz, m0);
                const bool h1 = slab(nd.c1lox, nd.c1hix, nd.c1loy, nd.c1hiy, nd.c1loz, nd.c1hiz, m1);
                if (!h0 && !h1) { node = stack[sp--]; continue; }
                node = h0 ? nd.child0 : nd.child1;
                if (h0 && h1) { int far = nd.child1; if (m1 < m0) { far = nd.child0; node = nd.child1; } if (sp >= 62) { std::printf("FAIL stack\n"); return 1; } stack[++sp] = far; }
            } else {
                const uint32_t link = ~(uint32_t)node, first = link >> 3, cnt = (link & 7u) + 1u;
                for (uint32_t k = 0; k < cnt; ++k) {
                    const TravTri& t = tt[first + k];
                    float tv;
                    if (tri(ro, rd, V{t.p0[0], t.p0[1], t.p0[2]}, V{t.e1[0], t.e1[1], t.e1[2]}, V{t.e2[0], t.e2[1], t.e2[2]}, tv))
                        if (tv < b2 || (tv == b2 && t.id < id2)) { b2 = tv; id2 = t.id; }
                }
                node = stack[sp--];
            }
        }
        if (id2 != bestId || (bestId != 0xFFFFFFFFu && b2 != best)) { std::printf("FAIL ray %u: linear (%u, %g) bvh (%u, %g)\n", r, bestId, best, id2, b2); return 1; }
        float b8; uint32_t id8;
        if (!walk8(nodes8, tt8, ro, rd, b8, id8, maxSp)) { std::printf("FAIL ray %u: cwbvh walk out of range\n", r); return 1; }
        if (id8 != bestId || (bestId != 0xFFFFFFFFu && b8 != best)) { std::printf("FAIL ray %u: linear (%u, %g) cwbvh (%u, %g)\n", r, bestId, best, id8, b8); return 1; }
        hits += bestId != 0xFFFFFFFFu;
    }
    std::printf("OK %u %u %u %.3f %.1f hits=%u | cwbvh nodes %u leaves %u depth %u sah %.3f %.1f ms maxsp %d\n", st.nodeCount, st.leafCount, st.maxDepth, st.sahCost,
                st.buildMs, hits, st8.nodeCount, st8.leafCount, st8.maxDepth, st8.sahCost, st8.buildMs, maxSp);
    return 0;
}
