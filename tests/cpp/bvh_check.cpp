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
        hits += bestId != 0xFFFFFFFFu;
    }
    std::printf("OK %u %u %u %.3f %.1f hits=%u\n", st.nodeCount, st.leafCount, st.maxDepth, st.sahCost, st.buildMs, hits);
    return 0;
}
