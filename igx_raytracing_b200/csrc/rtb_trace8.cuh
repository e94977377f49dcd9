// rtb_trace8.cuh — persistent traversal of the 8-wide compressed BVH (included by rtb_kernels.cu).
//
// One warp = 32 rays in flight.  Per iteration, in lock-step:
//   A  every lane whose current node group still has hit inner children takes the nearest one (highest bit of the
//      octant-permuted hit mask), fetches its 128-byte record (4 x 256-bit vector loads from one cache line, or
//      shared memory for the top of the tree) and tests the 8 child boxes; the result is a new node group and a
//      triangle group;
//   B  triangle groups are drained with Möller–Trumbore (the reference's arithmetic, SH/primitive.glsl:239-284);
//      when only a few lanes hold triangles and they still have node work, they postpone the group onto their stack;
//   C  lanes with nothing left in their group pop their stack, or retire the ray;
//   D  when fewer than REFILL8 lanes are still busy the idle lanes take new rays from the wavefront with one
//      warp-aggregated atomic (ballot + shfl).
//
// Box planes are bf16 grid coordinates packed two per word (Node8 in rtb_types.h).  The plane in the upper half is
// used as it stands (the word read as a float), the one in the lower half costs one shift; which of the lo / hi
// vectors is "near" is decided once per ray and applied as a load offset, so decoding takes no ALU-pipe work and the
// slab arithmetic is 48 FFMA per node.  The kernel is bound by L1 data-pipe wavefronts (profiles/r1c_*: 96 % of
// peak with 8 x LDG.128 per node), so a node is fetched with four LDG.256: header, and one granule per axis holding
// (lo, hi); the near / far assignment is a pair of complementary predicated loads with swapped destinations.
//
// Child order of the occlusion rays: nearest first like the nearest-hit search; farthest first visits 27.3 instead of
// 28.0 nodes per ray on the soup frame and runs 2 % slower (2.17 vs 2.14 ms) — the order hardly matters there.
// Also measured and rejected for the node fetch (occlusion launch, 2.14 ms with the four LDG.256): eight 16-byte texture
// fetches of the same record through a linear uint4 texture object, near / far by texel index: 2.42 ms.
//
// The per-lane stack holds 8-byte (base, mask) groups: the first SM_STACK entries in shared memory ([entry][thread],
// conflict-free for any mix of depths), the rest in local memory.
#pragma once

namespace rtb {

#ifndef RTB_CW_TOP_NODES
#define RTB_CW_TOP_NODES 0     // measured with 256-bit node loads: staging the top of the tree is 1-4 % slower than leaving it to L1
#endif
constexpr int TOP8_NODES = RTB_CW_TOP_NODES;   // 16 KB of shared memory: the breadth-first top of the tree (0 = not staged)
constexpr int SM_STACK = 8;          // 16 KB of shared memory per 256-thread block
constexpr int LOCAL_STACK = 56;      // a level can leave 3 entries (sibling group; postponed triangles + the re-pushed node group):
                                     // rtb_build_accel refuses trees with 3 * depth + 2 > SM_STACK + LOCAL_STACK
// (refill / postpone thresholds swept again on the queue of live rays, round 2, occlusion launch of the 4K soup frame:
//  22/8 1.994 ms, 22/4 1.992, 24/8 2.019, 26/12 2.063, 28/8 2.126)
#ifndef RTB_CW_REFILL
#define RTB_CW_REFILL 22
#endif
#ifndef RTB_CW_POSTPONE
#define RTB_CW_POSTPONE 8
#endif
constexpr int REFILL8 = RTB_CW_REFILL;
constexpr int POSTPONE8 = RTB_CW_POSTPONE;

// 32 bytes from global memory through the read-only path in one instruction (LDG.E.256, sm_100)
RTB_DI void ldg256(const char* p, uint4& a, uint4& b) {
    asm("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p));
}
// the same, first half -> a and second half -> b, or the other way round when swap != 0 (per lane): two complementary
// predicated loads, so the near / far choice costs no ALU work and no extra L1 wavefronts.  p2 == p, computed from a
// second kernel parameter holding the same base: given one address ptxas merges the pair into a single load followed
// by 16 predicated moves.
RTB_DI void ldg256swap(const char* p, const char* p2, uint32_t swap, uint4& a, uint4& b) {
    asm("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %10, 0;\n\t"
        "@!q ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n\t"
        "@q ld.global.nc.v8.u32 {%4,%5,%6,%7,%0,%1,%2,%3}, [%9];\n\t}"
        : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p), "l"(p2), "r"(swap));
}

template <int S>
RTB_DI void testPair(uint32_t nx, uint32_t ny, uint32_t nz, uint32_t fx, uint32_t fy, uint32_t fz, float kx, float ky, float kz,
                     float cx, float cy, float cz, float best, uint32_t& hitmask) {
    {
        const float tnx = fmaf(__uint_as_float(nx), kx, cx), tny = fmaf(__uint_as_float(ny), ky, cy), tnz = fmaf(__uint_as_float(nz), kz, cz);
        const float tfx = fmaf(__uint_as_float(fx), kx, cx), tfy = fmaf(__uint_as_float(fy), ky, cy), tfz = fmaf(__uint_as_float(fz), kz, cz);
        const float cmin = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, 0.0f));
        const float cmax = fminf(fminf(tfx, tfy), fminf(tfz, best));
        if (cmin <= cmax) hitmask |= (1u << (24 + S)) | (7u << (3 * S));
    }
    {
        const float tnx = fmaf(__uint_as_float(nx << 16), kx, cx), tny = fmaf(__uint_as_float(ny << 16), ky, cy), tnz = fmaf(__uint_as_float(nz << 16), kz, cz);
        const float tfx = fmaf(__uint_as_float(fx << 16), kx, cx), tfy = fmaf(__uint_as_float(fy << 16), ky, cy), tfz = fmaf(__uint_as_float(fz << 16), kz, cz);
        const float cmin = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, 0.0f));
        const float cmax = fminf(fminf(tfx, tfy), fminf(tfz, best));
        if (cmin <= cmax) hitmask |= (1u << (25 + S)) | (7u << (3 * S + 3));
    }
}

template <int MODE, bool COUNT>
#ifndef RTB_CW_MINBLOCKS
#define RTB_CW_MINBLOCKS 3   // 80 registers: the 256-bit loads need aligned register octets; 4 blocks/SM (64 regs) spills in the loop (5.95 vs 5.03 ms)
#endif
__global__ void __launch_bounds__(TRACE_THREADS, RTB_CW_MINBLOCKS) k_trace_cwbvh(const TraceArgs a) {
    __shared__ uint4 sTop[TOP8_NODES > 0 ? TOP8_NODES * 8 : 1];
    __shared__ uint2 sStack[SM_STACK][TRACE_THREADS];
    const int topN = min((int)a.nodeCount, TOP8_NODES);
    if (TOP8_NODES > 0) {
        for (int i = threadIdx.x; i < topN * 8; i += TRACE_THREADS) sTop[i] = __ldg(a.nodes8 + i);
        __syncthreads();
    }

    const unsigned lane = threadIdx.x & 31u;
    const unsigned lanesBelow = (1u << lane) - 1u;
    uint2 lstack[LOCAL_STACK];
    const uint32_t nRays = a.countPtr ? __ldg(a.countPtr) : a.n;   // a queue's length lives on the device

    bool active = false, exhausted = false;
    uint32_t slot = 0, prev = 0, bestId = NO_RAY_HIT, octinv = 0;
    float ox = 0, oy = 0, oz = 0, dx = 0, dy = 0, dz = 0, idx = 0, idy = 0, idz = 0;
    float best = 0, bu = 0, bv = 0;
    uint2 G = make_uint2(0u, 0u);
    int sp = 0;
    unsigned long long cRays = 0, cNodes = 0, cTris = 0, cHits = 0;

    auto push = [&](uint2 e) { if (sp < SM_STACK) sStack[sp][threadIdx.x] = e; else lstack[sp - SM_STACK] = e; ++sp; };
    auto pop = [&]() { --sp; return sp < SM_STACK ? sStack[sp][threadIdx.x] : lstack[sp - SM_STACK]; };

    for (;;) {
        // ---- D: refill ------------------------------------------------------------------------------------
        const unsigned busy = __ballot_sync(0xFFFFFFFFu, active);
        if (!exhausted && __popc(busy) < REFILL8) {
            const unsigned need = ~busy;
            uint32_t base = 0;
            const int leader = __ffs(need) - 1;
            if ((int)lane == leader) base = atomicAdd(a.workCounter, (uint32_t)__popc(need));
            base = __shfl_sync(0xFFFFFFFFu, base, leader);
            exhausted = base + (uint32_t)__popc(need) >= nRays;
            if (!active) {
                slot = base + (uint32_t)__popc(need & lanesBelow);
                if (slot < nRays) {
                    const float4 o = __ldg(reinterpret_cast<const float4*>(a.rays + slot));
                    const float4 d = __ldg(reinterpret_cast<const float4*>(a.rays + slot) + 1);
                    if (d.w >= 0.0f) {
                        ox = o.x; oy = o.y; oz = o.z; prev = fbits(o.w);
                        dx = d.x; dy = d.y; dz = d.z; best = d.w;
                        const float tiny = 8.271806e-25f;   // 2^-80: keeps 1/d finite for axis-parallel rays
                        idx = 1.0f / (fabsf(dx) > tiny ? dx : copysignf(tiny, dx));
                        idy = 1.0f / (fabsf(dy) > tiny ? dy : copysignf(tiny, dy));
                        idz = 1.0f / (fabsf(dz) > tiny ? dz : copysignf(tiny, dz));
                        // bit set = the ray travels towards + on that axis; slot s ^ octinv orders children far -> near
                        octinv = (idx < 0.0f ? 0u : 1u) | (idy < 0.0f ? 0u : 2u) | (idz < 0.0f ? 0u : 4u);
                        bestId = NO_RAY_HIT; bu = 0.0f; bv = 0.0f;
                        sp = 0; G = make_uint2(0u, 0x80000000u);   // the root as a one-node group
                        active = true;
                        if (COUNT) cRays++;
                    } else if (MODE == MODE_CLOSEST) {
                        TriHit h; h.t = NO_HIT; h.id = NO_RAY_HIT; h.u = 0.0f; h.v = 0.0f;
                        *reinterpret_cast<float4*>(a.hits + slot) = *reinterpret_cast<float4*>(&h);
                    } else if (MODE == MODE_ANY_BYTES) {
                        a.bytes[slot] = 0;
                    }
                }
            }
        }
        if (!__any_sync(0xFFFFFFFFu, active)) {
            if (exhausted) break;
            continue;
        }

        // ---- A: one node per lane -----------------------------------------------------------------------------
        uint2 T = make_uint2(0u, 0u);   // triangle group: (first triangle of the node, hit bits at 3 * slot + k)
        uint32_t P = 0u;                // triangles present in the node, same bit positions: index = base + popc(P below the bit)
        if (active) {
            if (G.y & 0xFF000000u) {
                const uint32_t hits = G.y;
                const uint32_t bit = 31u - (uint32_t)__clz(hits);
                const uint32_t childSlot = (bit - 24u) ^ octinv;
                const uint32_t rel = (uint32_t)__popc(hits & 0xFFu & ~(0xFFFFFFFFu << childSlot));
                const uint32_t nodeIdx = G.x + rel;
                G.y &= ~(1u << bit);
                if (G.y & 0xFF000000u) push(G);
                uint4 n0, n1, wnx, wny, wnz, wfx, wfy, wfz;
                // a ray travelling towards - on an axis meets the hi plane first: swap the two halves of that granule
                const uint32_t negX = ~octinv & 1u, negY = ~octinv & 2u, negZ = ~octinv & 4u;
                if (TOP8_NODES > 0 && (int)nodeIdx < topN) {
                    const char* p = reinterpret_cast<const char*>(sTop) + nodeIdx * 128u;
                    const uint32_t nearX = 32u + (negX << 4), nearY = 64u + (negY << 3), nearZ = 96u + (negZ << 2);
                    n0 = *reinterpret_cast<const uint4*>(p); n1 = *reinterpret_cast<const uint4*>(p + 16);
                    wnx = *reinterpret_cast<const uint4*>(p + nearX); wny = *reinterpret_cast<const uint4*>(p + nearY); wnz = *reinterpret_cast<const uint4*>(p + nearZ);
                    wfx = *reinterpret_cast<const uint4*>(p + (nearX ^ 16u)); wfy = *reinterpret_cast<const uint4*>(p + (nearY ^ 16u)); wfz = *reinterpret_cast<const uint4*>(p + (nearZ ^ 16u));
                } else {
                    const char* p = reinterpret_cast<const char*>(a.nodes8) + (size_t)nodeIdx * 128u;
                    ldg256(p, n0, n1);
                    const char* p2 = reinterpret_cast<const char*>(a.nodes8Alias) + (size_t)nodeIdx * 128u;
                    ldg256swap(p + 32, p2 + 32, negX, wnx, wfx);
                    ldg256swap(p + 64, p2 + 64, negY, wny, wfy);
                    ldg256swap(p + 96, p2 + 96, negZ, wnz, wfz);
                }
                if (COUNT) cNodes++;
                // plane at grid coordinate g along x: t = (p.x + g * 2^e - o.x) / d.x = g * kx + cx
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
                hitmask &= n1.z;   // valid: imask << 24 | the triangle bits the leaf slots own
                // inner bits from slot order to traversal order: bit 24 + s -> 24 + (s ^ octinv)
                uint32_t top = hitmask >> 24;
                if (octinv & 1u) top = ((top & 0x55u) << 1) | ((top >> 1) & 0x55u);
                if (octinv & 2u) top = ((top & 0x33u) << 2) | ((top >> 2) & 0x33u);
                if (octinv & 4u) top = ((top & 0x0Fu) << 4) | (top >> 4);
                P = n1.z & 0x00FFFFFFu;
                G = make_uint2(n1.x, (top << 24) | (n0.w >> 24));
                T = make_uint2(n1.y, hitmask & 0x00FFFFFFu);
            } else {
                T = G;                     // a postponed triangle group: its presence mask was pushed beneath it
                G = make_uint2(0u, 0u);
                P = pop().x;
            }
        }

        // ---- B: triangles -------------------------------------------------------------------------------------
        for (;;) {
            const unsigned m = __ballot_sync(0xFFFFFFFFu, T.y != 0u);
            if (!m) break;
            if (T.y != 0u && __popc(m) < POSTPONE8 && (G.y & 0xFF000000u)) { push(make_uint2(P, 0u)); push(T); T.y = 0u; }   // postpone: this lane has node work
            if (T.y != 0u) {
                const uint32_t bit = 31u - (uint32_t)__clz(T.y);
                T.y &= ~(1u << bit);
                const float4* tp = a.tris + (size_t)(T.x + (uint32_t)__popc(P & ~(0xFFFFFFFFu << bit))) * 3;
                const float4 t0 = __ldg(tp), t1 = __ldg(tp + 1), t2 = __ldg(tp + 2);
                if (COUNT) cTris++;
                float u, v, t, aa;
                if (triCandidate(mk3(ox, oy, oz), mk3(dx, dy, dz), mk3(t0.x, t0.y, t0.z), mk3(t1.x, t1.y, t1.z), mk3(t2.x, t2.y, t2.z), u, v, t, aa)) {
                    const uint32_t id = fbits(t0.w);
                    if (t > 0.0f && id != prev) {
                        if (MODE == MODE_CLOSEST) {
                            // reference: strict t < hitT in index order => on equal t the lower index wins
                            if (t < best || (t == best && id < bestId)) { best = t; bestId = id; bu = u; bv = v; }
                        } else if (t < best) {
                            bestId = id; T.y = 0u; G = make_uint2(0u, 0u); sp = 0;   // any hit ends the ray
                        }
                    }
                }
            }
        }

        // ---- C: pop or retire -----------------------------------------------------------------------------------
        if (active && !(G.y & 0xFF000000u)) {
            if (sp > 0) G = pop();
            else {
                active = false;
                if (MODE == MODE_CLOSEST) {
                    TriHit h; h.t = bestId == NO_RAY_HIT ? NO_HIT : best; h.id = bestId; h.u = bu; h.v = bv;
                    *reinterpret_cast<float4*>(a.hits + slot) = *reinterpret_cast<float4*>(&h);
                    if (COUNT && bestId != NO_RAY_HIT) cHits++;
                } else if (MODE == MODE_ANY_BITS) {
                    if (bestId != NO_RAY_HIT) {
                        const uint32_t j = a.slotIds ? __ldg(a.slotIds + slot) : slot;
                        const uint32_t sample = j / a.fm.localSlots, i = j - sample * a.fm.localSlots;
                        uint32_t x, y;
                        slotToPixel(a.fm, i, x, y);
                        atomicOr(a.bits + indexToLight(x, y, a.fm.w, a.fm.h, sample), 1u << ((x & 15u) | ((y & 1u) << 4)));
                        if (COUNT) cHits++;
                    }
                } else {
                    a.bytes[slot] = bestId != NO_RAY_HIT ? 1 : 0;
                }
            }
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
