// rtb_trace8f.cuh — frustum packets: warp-cooperative nearest-hit traversal of the 8-wide compressed BVH in which the
// BOX tests are done once per packet instead of once per ray (included by rtb_kernels.cu after rtb_trace8p.cuh).
//
// rtb_trace8p.cuh walks the union of the nodes the 32 rays of an 8x4-pixel patch need, but every lane still tests all
// eight child boxes of every node against its own ray: ~130 of the ~190 instructions per node, on a kernel bound by
// instruction issue.  Camera rays of a patch share their origin and differ by a few pixel angles, so the patch can be
// bounded by four thin frusta (its 4x2-pixel quadrants), each an interval ray: origin o, direction intervals
// [dmin, dmax] per axis.  One lane tests ONE child box against ONE quadrant (8 children x 4 quadrants = 32 lanes):
//
//     entry >= min over the interval of (near plane - o) / d,   exit <= max over the interval of (far plane - o) / d
//
// with interval end points widened by 2^-20 and the final comparison given 1e-5 of slack.  The test is conservative —
// a child is visited whenever any ray of the quadrant could enter its (already padded, outward-rounded) box before the
// packet's largest nearest-hit distance — so the set of triangles tested is a superset of what each ray needs, and the
// triangles themselves are tested by every lane with the reference's exact Möller–Trumbore arithmetic and tie rule.
// The hits are therefore bit-identical to the per-ray kernel's (tests/test_gpu_parity.py: packets 0 / 1 / 3 compared
// with array equality, and against the brute-force loop).
//
// An axis on which a quadrant's directions straddle zero (or are tiny), and every axis when the rays of a packet do not
// share one origin, is left unconstrained: correct, only slower.  The host picks this kernel for the Default projection
// (one eye) when the patch is small against the leaf nodes (rtb_api.cu, primaryPackets).
#pragma once

namespace rtb {

// Measured on B200 (1M-triangle soup, 3840x2160): 4 blocks/SM (64 registers) 2.93 ms; 3 blocks (72 registers) 3.10 ms;
// 5 blocks (48 registers, spills) 3.26 ms.  ncu (profiles/r1o_frustum_trace_full.md): issue slots 85 % busy, 143 instructions
// per node visit against ~190 in rtb_trace8p.cuh.
#ifndef RTB_FR_MINBLOCKS
#define RTB_FR_MINBLOCKS 4
#endif

template <bool COUNT>
__global__ void __launch_bounds__(TRACE_THREADS, RTB_FR_MINBLOCKS) k_trace_cwbvh_frustum(const TraceArgs a) {
    __shared__ uint2 sStack[TRACE_THREADS / 32][PACKET_STACK];
    __shared__ uint8_t sPerm[8][256];     // hit-mask bits from slot order to traversal order: bit s -> bit s ^ octant
    __shared__ uint32_t sSpread[256];     // child bit c -> the three triangle bits 3c .. 3c+2
    for (uint32_t i = threadIdx.x; i < 2048u; i += TRACE_THREADS) {
        const uint32_t o = i >> 8, m = i & 255u;
        uint32_t r = 0;
        for (uint32_t b = 0; b < 8u; ++b) r |= ((m >> b) & 1u) << (b ^ o);
        sPerm[o][m] = (uint8_t)r;
    }
    for (uint32_t m = threadIdx.x; m < 256u; m += TRACE_THREADS) {
        uint32_t r = 0;
        for (uint32_t b = 0; b < 8u; ++b) if ((m >> b) & 1u) r |= 7u << (3u * b);
        sSpread[m] = r;
    }
    __syncthreads();

    const unsigned lane = threadIdx.x & 31u;
    uint2* stack = sStack[threadIdx.x >> 5];
    // test role of this lane: child slot c of its own quadrant (the 8 lanes of a quadrant differ in lane bits 0, 1, 3)
    const uint32_t c = (lane & 3u) | ((lane >> 1) & 4u);
    const uint32_t planeOff = 32u + (c >> 1) * 4u;     // byte offset of the word holding slot c's lo plane on x
    const bool lowerHalf = (c & 1u) != 0u;             // odd slots live in the lower 16 bits of their word
    unsigned long long cRays = 0, cNodes = 0, cTris = 0, cHits = 0;

    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(a.workCounter, 32u);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (base >= a.n) break;
        const uint32_t slot = base + lane;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f), d = make_float4(0.f, 0.f, 1.f, -1.0f);
        if (slot < a.n) {
            o = __ldg(reinterpret_cast<const float4*>(a.rays + slot));
            d = __ldg(reinterpret_cast<const float4*>(a.rays + slot) + 1);
        }
        const bool live = d.w >= 0.0f;
        const unsigned liveMask = __ballot_sync(0xFFFFFFFFu, live);
        const float ox = o.x, oy = o.y, oz = o.z, dx = d.x, dy = d.y, dz = d.z;
        const uint32_t prev = fbits(o.w);
        float best = live ? d.w : -1.0f;
        float bu = 0.0f, bv = 0.0f;
        uint32_t bestId = NO_RAY_HIT;
        if (liveMask) {
            if (COUNT && live) cRays++;
            // ---- the quadrant's interval ray ------------------------------------------------------------------------
            const float INF = __int_as_float(0x7F800000);
            float lox = live ? dx : INF, hix = live ? dx : -INF, loy = live ? dy : INF, hiy = live ? dy : -INF, loz = live ? dz : INF, hiz = live ? dz : -INF;
            #pragma unroll
            for (int s = 0; s < 3; ++s) {
                const int m = s == 0 ? 1 : (s == 1 ? 2 : 8);
                lox = fminf(lox, __shfl_xor_sync(0xFFFFFFFFu, lox, m)); hix = fmaxf(hix, __shfl_xor_sync(0xFFFFFFFFu, hix, m));
                loy = fminf(loy, __shfl_xor_sync(0xFFFFFFFFu, loy, m)); hiy = fmaxf(hiy, __shfl_xor_sync(0xFFFFFFFFu, hiy, m));
                loz = fminf(loz, __shfl_xor_sync(0xFFFFFFFFu, loz, m)); hiz = fmaxf(hiz, __shfl_xor_sync(0xFFFFFFFFu, hiz, m));
            }
            const int first = __ffs(liveMask) - 1;
            const float fox = __shfl_sync(0xFFFFFFFFu, ox, first), foy = __shfl_sync(0xFFFFFFFFu, oy, first), foz = __shfl_sync(0xFFFFFFFFu, oz, first);
            const bool oneOrigin = __all_sync(0xFFFFFFFFu, !live || (ox == fox && oy == foy && oz == foz));
            const bool quadLive = lox <= hix;                      // some live ray in this lane's quadrant
            const float tiny = 8.271806e-25f;                      // 2^-80
            const float widen = 9.5367431640625e-7f;               // 2^-20
            // reciprocal interval per axis; an axis is "free" (unconstrained) when the interval touches zero
            const bool zeroX = !(lox > tiny || hix < -tiny), zeroY = !(loy > tiny || hiy < -tiny), zeroZ = !(loz > tiny || hiz < -tiny);
            const bool freeX = !oneOrigin || zeroX, freeY = !oneOrigin || zeroY, freeZ = !oneOrigin || zeroZ;
            // an axis whose direction interval contains zero has no reciprocal interval; it is tested as a wedge instead:
            // up to the exit distance t1 given by the other axes, the quadrant's rays stay within [t1 * min(lo, 0), t1 * max(hi, 0)]
            const bool wedge = oneOrigin && (zeroX || zeroY || zeroZ);
            const float wlo_x = fminf(lox, 0.0f), whi_x = fmaxf(hix, 0.0f), wlo_y = fminf(loy, 0.0f), whi_y = fmaxf(hiy, 0.0f), wlo_z = fminf(loz, 0.0f), whi_z = fmaxf(hiz, 0.0f);
            float ilx = freeX ? 0.0f : 1.0f / hix, ihx = freeX ? 0.0f : 1.0f / lox;
            float ily = freeY ? 0.0f : 1.0f / hiy, ihy = freeY ? 0.0f : 1.0f / loy;
            float ilz = freeZ ? 0.0f : 1.0f / hiz, ihz = freeZ ? 0.0f : 1.0f / loz;
            ilx -= fabsf(ilx) * widen; ihx += fabsf(ihx) * widen;
            ily -= fabsf(ily) * widen; ihy += fabsf(ihy) * widen;
            ilz -= fabsf(ilz) * widen; ihz += fabsf(ihz) * widen;
            const bool negX = hix < 0.0f, negY = hiy < 0.0f, negZ = hiz < 0.0f;   // the quadrant travels towards - on that axis
            // child order for the warp: octant of the first live ray
            const uint32_t octLane = (dx < 0.0f ? 0u : 1u) | (dy < 0.0f ? 0u : 2u) | (dz < 0.0f ? 0u : 4u);
            const uint32_t woct = __shfl_sync(0xFFFFFFFFu, octLane, first);
            const uint8_t* permRow = sPerm[woct];
            uint32_t limitBits = __reduce_max_sync(0xFFFFFFFFu, live ? fbits(best) : 0u);   // largest nearest-hit distance in the packet

            int sp = 0;
            uint2 G = make_uint2(0u, 0x80000000u);   // the root as a one-node group
            for (;;) {
                // ---- one node for the warp ------------------------------------------------------------------------
                const uint32_t hits = G.y;
                const uint32_t bit = 31u - (uint32_t)__clz(hits);
                const uint32_t childSlot = (bit - 24u) ^ woct;
                const uint32_t nodeIdx = G.x + (uint32_t)__popc(hits & 0xFFu & ~(0xFFFFFFFFu << childSlot));
                G.y &= ~(1u << bit);
                if (G.y & 0xFF000000u) { stack[sp] = G; ++sp; }
                const char* p = reinterpret_cast<const char*>(a.nodes8) + (size_t)nodeIdx * 128u;
                uint4 n0, n1;
                ldg256(p, n0, n1);
                const char* q = p + planeOff;
                const uint32_t wlx = __ldg(reinterpret_cast<const uint32_t*>(q)), whx = __ldg(reinterpret_cast<const uint32_t*>(q + 16));
                const uint32_t wly = __ldg(reinterpret_cast<const uint32_t*>(q + 32)), why = __ldg(reinterpret_cast<const uint32_t*>(q + 48));
                const uint32_t wlz = __ldg(reinterpret_cast<const uint32_t*>(q + 64)), whz = __ldg(reinterpret_cast<const uint32_t*>(q + 80));
                if (COUNT && lane == 0) cNodes++;
                // this lane's child box in world space, relative to the origin: (p + g * 2^e) - o
                const float sx = __uint_as_float((n0.w & 0xFFu) << 23), sy = __uint_as_float((n0.w << 15) & 0x7F800000u), sz = __uint_as_float((n0.w << 7) & 0x7F800000u);
                const float px = __uint_as_float(n0.x) - fox, py = __uint_as_float(n0.y) - foy, pz = __uint_as_float(n0.z) - foz;
                const float glx = __uint_as_float(lowerHalf ? wlx << 16 : wlx & 0xFFFF0000u), ghx = __uint_as_float(lowerHalf ? whx << 16 : whx | 0xFFFFu);
                const float gly = __uint_as_float(lowerHalf ? wly << 16 : wly & 0xFFFF0000u), ghy = __uint_as_float(lowerHalf ? why << 16 : why | 0xFFFFu);
                const float glz = __uint_as_float(lowerHalf ? wlz << 16 : wlz & 0xFFFF0000u), ghz = __uint_as_float(lowerHalf ? whz << 16 : whz | 0xFFFFu);
                const float Wlx = fmaf(glx, sx, px), Whx = fmaf(ghx, sx, px), Wly = fmaf(gly, sy, py), Why = fmaf(ghy, sy, py), Wlz = fmaf(glz, sz, pz), Whz = fmaf(ghz, sz, pz);
                const float nx = negX ? Whx : Wlx, fx = negX ? Wlx : Whx, ny = negY ? Why : Wly, fy = negY ? Wly : Why, nz = negZ ? Whz : Wlz, fz = negZ ? Wlz : Whz;
                const float ex = freeX ? -INF : fminf(nx * ilx, nx * ihx), xx = freeX ? INF : fmaxf(fx * ilx, fx * ihx);
                const float ey = freeY ? -INF : fminf(ny * ily, ny * ihy), xy = freeY ? INF : fmaxf(fy * ily, fy * ihy);
                const float ez = freeZ ? -INF : fminf(nz * ilz, nz * ihz), xz = freeZ ? INF : fmaxf(fz * ilz, fz * ihz);
                const float entry = fmaxf(fmaxf(ex, ey), fmaxf(ez, 0.0f));
                const float exit = fminf(fminf(xx, xy), fminf(xz, __uint_as_float(limitBits)));
                bool hit = quadLive && entry <= exit * 1.00001f + 1e-30f;
                if (wedge) {
                    const float t1 = fminf(exit, 1e30f) * 1.00001f;
                    if (zeroX) hit = hit && Wlx <= t1 * whi_x + 1e-30f && Whx >= t1 * wlo_x - 1e-30f;
                    if (zeroY) hit = hit && Wly <= t1 * whi_y + 1e-30f && Why >= t1 * wlo_y - 1e-30f;
                    if (zeroZ) hit = hit && Wlz <= t1 * whi_z + 1e-30f && Whz >= t1 * wlo_z - 1e-30f;
                }
                const uint32_t any8 = __reduce_or_sync(0xFFFFFFFFu, hit ? (1u << c) : 0u);
                const uint32_t imask = n0.w >> 24;
                const uint32_t top = permRow[any8 & imask];
                const uint32_t P = n1.z & 0x00FFFFFFu;
                uint32_t T = sSpread[any8 & ~imask & 0xFFu] & P;

                // ---- the triangles of the hit leaf slots, every lane against its own ray ------------------------------
                if (T) {
                    do {
                        const uint32_t tb = 31u - (uint32_t)__clz(T);
                        T &= ~(1u << tb);
                        const float4* tp = a.tris + (size_t)(n1.y + (uint32_t)__popc(P & ~(0xFFFFFFFFu << tb))) * 3;
                        const float4 t0 = __ldg(tp), t1 = __ldg(tp + 1), t2 = __ldg(tp + 2);
                        if (COUNT && lane == 0) cTris++;
                        float u, v, t, aa;
                        if (triCandidate(mk3(ox, oy, oz), mk3(dx, dy, dz), mk3(t0.x, t0.y, t0.z), mk3(t1.x, t1.y, t1.z), mk3(t2.x, t2.y, t2.z), u, v, t, aa)) {
                            const uint32_t id = fbits(t0.w);
                            // reference: strict t < hitT in index order => on equal t the lower index wins
                            if (t > 0.0f && id != prev && (t < best || (t == best && id < bestId))) { best = t; bestId = id; bu = u; bv = v; }
                        }
                    } while (T);
                    limitBits = __reduce_max_sync(0xFFFFFFFFu, live ? fbits(best) : 0u);
                }

                // ---- descend, or pop ----------------------------------------------------------------------------------
                if (top) G = make_uint2(n1.x, (top << 24) | imask);
                else if (sp > 0) { --sp; G = stack[sp]; }
                else break;
            }
        }
        if (slot < a.n) {
            TriHit h; h.t = bestId == NO_RAY_HIT ? NO_HIT : best; h.id = bestId; h.u = bu; h.v = bv;
            *reinterpret_cast<float4*>(a.hits + slot) = *reinterpret_cast<float4*>(&h);
            if (COUNT && bestId != NO_RAY_HIT) cHits++;
        }
    }

    if (COUNT) {
        for (int o = 16; o > 0; o >>= 1) {
            cRays += __shfl_xor_sync(0xFFFFFFFFu, cRays, o); cNodes += __shfl_xor_sync(0xFFFFFFFFu, cNodes, o);
            cTris += __shfl_xor_sync(0xFFFFFFFFu, cTris, o); cHits += __shfl_xor_sync(0xFFFFFFFFu, cHits, o);
        }
        if (lane == 0) {
            atomicAdd(&a.counters->rays, cRays); atomicAdd(&a.counters->nodes, cNodes);
            atomicAdd(&a.counters->tris, cTris); atomicAdd(&a.counters->hits, cHits);
        }
    }
}

}  // namespace rtb
