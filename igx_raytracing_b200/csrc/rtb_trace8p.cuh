// rtb_trace8p.cuh — warp-cooperative ("packet") nearest-hit traversal of the 8-wide compressed BVH for COHERENT ray
// patches (included by rtb_kernels.cu after rtb_trace8.cuh).
//
// The wavefront is laid out so that 32 consecutive slots are an 8x4 pixel patch (FrameMap): the primary rays of a
// warp share their origin and differ by a few pixel angles.  When that patch is small against the leaf nodes of the
// tree (the host decides: rtb_api.cu, primaryPackets), the 32 rays visit nearly the same nodes — but in the per-lane
// kernel they drift out of step after the first differing child test, and from then on every lane pays its own L1
// wavefronts for the same cache lines (profiles/r1d_*: 4 wavefronts per lane per node, data pipe 88 % busy).
//
// Here the WARP owns the traversal state: one stack, one current node group, one triangle group.  A node record is
// fetched once per warp (every lane reads the same address: one wavefront per 256-bit load), each lane tests the eight
// child boxes against its OWN ray and its OWN nearest distance, and the warp descends into every child that any lane
// hits (one REDUX.OR over the hit masks).  Triangles of the union are tested by every lane with the reference's
// Möller–Trumbore arithmetic; testing a triangle whose leaf box a lane missed cannot change that lane's result, since
// the exact test is what decides (the reference tests every triangle against every ray, SH/trace.glsl:25-29).
// Child order is the octant order of the first live lane; it only affects how early `best` shrinks.
#pragma once

namespace rtb {

constexpr int PACKET_STACK = 64;   // node groups only (no postponed triangle groups): at most one per tree level
#ifndef RTB_PK_MINBLOCKS
#define RTB_PK_MINBLOCKS 3
#endif
#ifndef RTB_PK_PERM_LUT
#define RTB_PK_PERM_LUT 1
#endif

// Measured on B200 (1M-triangle soup, 3840x2160, profiles/r1g_*): 3 blocks/SM at 77 registers 3.88 ms; forcing 4 blocks
// (64 registers) 4.16 ms; requesting the hit children and triangles early (LDGSTS prefetch into L1) 4.11 ms.  The kernel
// is bound by instruction issue (77 % of peak, ~200 instructions per node of which ~95 on the half-rate ALU pipe), not by
// memory latency, so neither occupancy nor prefetching helps.
template <bool COUNT>
__global__ void __launch_bounds__(TRACE_THREADS, RTB_PK_MINBLOCKS) k_trace_cwbvh_packet(const TraceArgs a) {
    __shared__ uint2 sStack[TRACE_THREADS / 32][PACKET_STACK];
#if RTB_PK_PERM_LUT
    // hit-mask bits from slot order to traversal order (bit s -> bit s ^ octant) as one shared-memory byte load instead of
    // three conditional swap stages: 8 octants x 256 masks
    __shared__ uint8_t sPerm[8][256];
    for (uint32_t i = threadIdx.x; i < 2048u; i += TRACE_THREADS) {
        const uint32_t o = i >> 8, m = i & 255u;
        uint32_t r = 0;
        for (uint32_t b = 0; b < 8u; ++b) r |= ((m >> b) & 1u) << (b ^ o);
        sPerm[o][m] = (uint8_t)r;
    }
    __syncthreads();
#endif
    const unsigned lane = threadIdx.x & 31u;
    uint2* stack = sStack[threadIdx.x >> 5];
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
        float best = live ? d.w : -1.0f;   // a negative limit fails every box and every triangle test
        float bu = 0.0f, bv = 0.0f;
        uint32_t bestId = NO_RAY_HIT;
        if (liveMask) {
            const float tiny = 8.271806e-25f;   // 2^-80: keeps 1/d finite for axis-parallel rays
            const float idx = 1.0f / (fabsf(dx) > tiny ? dx : copysignf(tiny, dx));
            const float idy = 1.0f / (fabsf(dy) > tiny ? dy : copysignf(tiny, dy));
            const float idz = 1.0f / (fabsf(dz) > tiny ? dz : copysignf(tiny, dz));
            const uint32_t octinv = (idx < 0.0f ? 0u : 1u) | (idy < 0.0f ? 0u : 2u) | (idz < 0.0f ? 0u : 4u);
            const uint32_t negX = ~octinv & 1u, negY = ~octinv & 2u, negZ = ~octinv & 4u;   // per lane: near / far of its own ray
            const uint32_t woct = __shfl_sync(0xFFFFFFFFu, octinv, __ffs(liveMask) - 1);    // per warp: child order
            if (COUNT && live) cRays++;
#if RTB_PK_PERM_LUT
            const uint8_t* permRow = sPerm[woct];
#endif

            int sp = 0;
            uint2 G = make_uint2(0u, 0x80000000u);   // the root as a one-node group
            for (;;) {
                // ---- one node for the warp ------------------------------------------------------------------------
                const uint32_t hits = G.y;
                const uint32_t bit = 31u - (uint32_t)__clz(hits);
                const uint32_t childSlot = (bit - 24u) ^ woct;
                const uint32_t nodeIdx = G.x + (uint32_t)__popc(hits & 0xFFu & ~(0xFFFFFFFFu << childSlot));
                G.y &= ~(1u << bit);
                if (G.y & 0xFF000000u) { if (lane == 0) stack[sp] = G; ++sp; }   // one writer; __syncwarp below orders it before any pop
                uint4 n0, n1, wnx, wny, wnz, wfx, wfy, wfz;
                const char* p = reinterpret_cast<const char*>(a.nodes8) + (size_t)nodeIdx * 128u;
                const char* p2 = reinterpret_cast<const char*>(a.nodes8Alias) + (size_t)nodeIdx * 128u;
                ldg256(p, n0, n1);
                ldg256swap(p + 32, p2 + 32, negX, wnx, wfx);
                ldg256swap(p + 64, p2 + 64, negY, wny, wfy);
                ldg256swap(p + 96, p2 + 96, negZ, wnz, wfz);
                if (COUNT && lane == 0) cNodes++;
                const float kx = __uint_as_float((n0.w & 0xFFu) << 23) * idx;
                const float ky = __uint_as_float((n0.w << 15) & 0x7F800000u) * idy;
                const float kz = __uint_as_float((n0.w << 7) & 0x7F800000u) * idz;
                const float cx = (__uint_as_float(n0.x) - ox) * idx;
                const float cy = (__uint_as_float(n0.y) - oy) * idy;
                const float cz = (__uint_as_float(n0.z) - oz) * idz;
                uint32_t hitmask = 0;
                testPair<0>(wnx.x, wny.x, wnz.x, wfx.x, wfy.x, wfz.x, kx, ky, kz, cx, cy, cz, best, hitmask);
                testPair<2>(wnx.y, wny.y, wnz.y, wfx.y, wfy.y, wfz.y, kx, ky, kz, cx, cy, cz, best, hitmask);
                testPair<4>(wnx.z, wny.z, wnz.z, wfx.z, wfy.z, wfz.z, kx, ky, kz, cx, cy, cz, best, hitmask);
                testPair<6>(wnx.w, wny.w, wnz.w, wfx.w, wfy.w, wfz.w, kx, ky, kz, cx, cy, cz, best, hitmask);
                const uint32_t any = __reduce_or_sync(0xFFFFFFFFu, hitmask) & n1.z;   // valid: imask << 24 | triangle presence
#if RTB_PK_PERM_LUT
                const uint32_t top = permRow[any >> 24];
#else
                uint32_t top = any >> 24;
                if (woct & 1u) top = ((top & 0x55u) << 1) | ((top >> 1) & 0x55u);
                if (woct & 2u) top = ((top & 0x33u) << 2) | ((top >> 2) & 0x33u);
                if (woct & 4u) top = ((top & 0x0Fu) << 4) | (top >> 4);
#endif
                const uint32_t P = n1.z & 0x00FFFFFFu;
                uint32_t T = any & 0x00FFFFFFu;

                // ---- the union's triangles, every lane against its own ray ------------------------------------------
                while (T) {
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
                }

                // ---- descend, or pop ----------------------------------------------------------------------------------
                if (top) G = make_uint2(n1.x, (top << 24) | (n0.w >> 24));
                else if (sp > 0) { __syncwarp(); --sp; G = stack[sp]; __syncwarp(); }   // (the second barrier: the slot is not rewritten before every lane has read it)
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
