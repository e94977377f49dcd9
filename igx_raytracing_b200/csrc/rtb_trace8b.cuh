// rtb_trace8b.cuh — beam packets: warp-cooperative OCCLUSION traversal of the 8-wide compressed BVH for rays that travel along
// neighbouring lines (included by rtb_kernels.cu after rtb_trace8f.cuh; replaces the triangle loop of SH/trace.glsl:78-81 for
// the shadow rays when RTB_OPT_SHADOW_ORDER = 3).
//
// The occlusion rays of a frame go to one light: nearly parallel (a sun) or converging on a point.  In pixel order a warp's 32
// rays start on unrelated triangles and share nothing; sorted by their 2D light-space coordinate (rtb_sort.cu) 32 consecutive
// rays form a thin beam — whatever depth each starts from — and the frustum-packet idea of rtb_trace8f.cuh applies with the roles
// of origin and direction swapped: the directions are (nearly) one, the origins spread.
//
//   rebasing   every ray is re-parametrised from the plane perpendicular to the beam's dominant axis through the rearmost origin:
//              o' = o + s d with s <= 0, so the beam's origins o' lie in a small rectangle of that plane (the thickness of the
//              beam), not in the fat axis-aligned box an oblique beam's true origins span; ray i is the part t' >= t0_i = -s of
//              its line;
//   box test   one lane tests ONE child box against the interval ray (origin box [Omin, Omax], reciprocal direction intervals
//              [al, ah], everything mirrored so that the beam travels towards +):
//                  entry >= min over the intervals of (near - o') / d = min((near - Omax) al, (near - Omax) ah)
//                  exit  <= max over the intervals of (far  - o') / d = max((far  - Omin) al, (far  - Omin) ah)
//              intersected like a slab test with [0, largest t0_i + tmax_i of the rays still undecided].  Conservative by
//              construction (intervals widened by 2^-20, origin box by the rebasing's rounding, 1e-5 of slack on the compare,
//              boxes already padded and rounded outwards): a child is visited whenever ANY ray of the beam could enter it;
//   triangles  the triangles of the hit leaf slots are tested by every undecided lane against its ORIGINAL ray (o, d, tmax, prev)
//              with the reference's Möller–Trumbore arithmetic and accept rule (t > 0, t < maxDist, not the ray's own object), so
//              the shadow bits are the per-ray kernel's bits (tests/test_gpu_parity.py::test_shadow_order_modes_identical,
//              tests/test_gpu_headline.py);
//   walk       four nodes per step, a per-warp (node, entry) stack, as walkFrustumPacket4; a lane that has found an occluder
//              stops testing, the packet ends when every lane has or the stack is empty.
//
// Packets the walk does not take — direction signs differ between the rays, a direction component is (almost) zero, the stack
// would overflow — are appended to a fall-back queue that the per-ray kernel answers afterwards.
#pragma once

namespace rtb {

struct BeamPlain {
    uint32_t smx, smy, smz;          // sign of the mirroring as a bit mask for the grid step 2^e
    float sgx, sgy, sgz;             // +-1
    float moNx, moNy, moNz;          // -(largest mirrored origin coordinate): offset of the NEAR planes
    float moFx, moFy, moFz;          // -(smallest mirrored origin coordinate): offset of the FAR planes
    float alx, ahx, aly, ahy, alz, ahz;   // reciprocal interval of |d|, widened: 0 < al <= ah
};

// float <-> unsigned with the same order (for redux.min / redux.max over signed floats)
RTB_DI uint32_t orderedBits(float f) { const uint32_t b = fbits(f); return b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u); }
RTB_DI float fromOrderedBits(uint32_t u) { return ubits(u ^ ((u >> 31) ? 0x80000000u : 0xFFFFFFFFu)); }
RTB_DI float warpMinF(float v, bool on) { return fromOrderedBits(__reduce_min_sync(0xFFFFFFFFu, on ? orderedBits(v) : 0xFFFFFFFFu)); }
RTB_DI float warpMaxF(float v, bool on) { return fromOrderedBits(__reduce_max_sync(0xFFFFFFFFu, on ? orderedBits(v) : 0u)); }

// returns true when the stack would overflow (the caller forwards the undecided rays to the fall-back queue)
template <bool COUNT>
RTB_DI bool walkBeamPacket4(const TraceArgs& a, uint2* stack, uint32_t woct, bool live, float ox, float oy, float oz, float dx, float dy, float dz,
                            uint32_t prev, float tmax, float tEnd, const BeamPlain& qp, bool& occluded,
                            unsigned long long& cNodes, unsigned long long& cTris, unsigned lane) {
    const uint32_t lanesBelow = (1u << lane) - 1u;
    const uint32_t g = 3u - (lane >> 3);                 // which popped entry this lane works on: lanes 24..31 take the top of the stack
    const uint32_t cs = (lane & 7u) ^ woct;              // its child slot: within a group, a higher lane is a nearer child
    const uint32_t planeOff = 32u + (cs >> 1) * 4u;
    const uint32_t halfSel = (cs & 1u) ? 0x1044u : 0x3244u;
    const uint32_t csBelow = (1u << cs) - 1u, triShift = 3u * cs;
    uint32_t limitBits = __reduce_max_sync(0xFFFFFFFFu, (live && !occluded) ? fbits(tEnd) : 0u);   // non-negative floats order like their bits
    if (lane == 0) stack[0] = make_uint2(0u, 0u);        // the root, entry distance 0
    int sp = 1;
    __syncwarp();
    while (sp > 0) {
        const int nPop = sp > PACKET4_STACK - 40 ? 1 : min(sp, 4);
        uint2 e = make_uint2(0u, 0xFFFFFFFFu);
        if ((int)g < nPop) e = stack[sp - 1 - (int)g];
        sp -= nPop;
        const bool activeNode = e.y <= limitBits;
        __syncwarp();                                    // every pop has been read before this step's pushes land
        bool inner = false, leaf = false;
        uint32_t childIdx = 0, triFirst = 0, triCnt = 0, entryBits = 0;
        if (activeNode) {
            const char* p = reinterpret_cast<const char*>(a.nodes8) + (size_t)e.x * 128u;
            uint4 n0, n1;
            ldg256(p, n0, n1);
            const char* q = p + planeOff;
            const uint32_t wlx = __ldg(reinterpret_cast<const uint32_t*>(q)), whx = __ldg(reinterpret_cast<const uint32_t*>(q + 16));
            const uint32_t wly = __ldg(reinterpret_cast<const uint32_t*>(q + 32)), why = __ldg(reinterpret_cast<const uint32_t*>(q + 48));
            const uint32_t wlz = __ldg(reinterpret_cast<const uint32_t*>(q + 64)), whz = __ldg(reinterpret_cast<const uint32_t*>(q + 80));
            if (COUNT && (lane & 7u) == 0u) cNodes++;
            const float glx = __uint_as_float(__byte_perm(wlx, 0u, halfSel)), ghx = __uint_as_float(__byte_perm(whx, 0u, halfSel));
            const float gly = __uint_as_float(__byte_perm(wly, 0u, halfSel)), ghy = __uint_as_float(__byte_perm(why, 0u, halfSel));
            const float glz = __uint_as_float(__byte_perm(wlz, 0u, halfSel)), ghz = __uint_as_float(__byte_perm(whz, 0u, halfSel));
            // mirrored plane coordinates sgn * (p + g * 2^e); the near plane is the smaller of the two, whatever the sign was
            const float sx = __uint_as_float(((n0.w & 0xFFu) << 23) ^ qp.smx), sy = __uint_as_float(((n0.w << 15) & 0x7F800000u) ^ qp.smy), sz = __uint_as_float(((n0.w << 7) & 0x7F800000u) ^ qp.smz);
            const float px = __uint_as_float(n0.x) * qp.sgx, py = __uint_as_float(n0.y) * qp.sgy, pz = __uint_as_float(n0.z) * qp.sgz;
            const float ax = fmaf(glx, sx, px), bx = fmaf(ghx, sx, px), ay = fmaf(gly, sy, py), by = fmaf(ghy, sy, py), az = fmaf(glz, sz, pz), bz = fmaf(ghz, sz, pz);
            // relative to the origin box: the near plane against the largest origin, the far plane against the smallest
            const float nx = fminf(ax, bx) + qp.moNx, fx = fmaxf(ax, bx) + qp.moFx;
            const float ny = fminf(ay, by) + qp.moNy, fy = fmaxf(ay, by) + qp.moFy;
            const float nz = fminf(az, bz) + qp.moNz, fz = fmaxf(az, bz) + qp.moFz;
            const float ex = fminf(nx * qp.alx, nx * qp.ahx), xx = fmaxf(fx * qp.alx, fx * qp.ahx);
            const float ey = fminf(ny * qp.aly, ny * qp.ahy), xy = fmaxf(fy * qp.aly, fy * qp.ahy);
            const float ez = fminf(nz * qp.alz, nz * qp.ahz), xz = fmaxf(fz * qp.alz, fz * qp.ahz);
            const float entry = fmaxf(fmaxf(ex, ey), fmaxf(ez, 0.0f));
            const float exit = fminf(fminf(xx, xy), fminf(xz, __uint_as_float(limitBits)));
            const bool hit = entry <= fmaf(exit, 1.00001f, 1e-30f);
            const uint32_t imask = n0.w >> 24, P = n1.z & 0x00FFFFFFu;
            triCnt = (uint32_t)__popc((P >> triShift) & 7u);
            inner = hit && ((imask >> cs) & 1u);
            leaf = hit && triCnt != 0u;                  // a slot is an inner node or holds triangles, never both
            childIdx = n1.x + (uint32_t)__popc(imask & csBelow);
            triFirst = n1.y + (uint32_t)__popc(P & ((1u << triShift) - 1u));
            entryBits = fbits(entry);
        }
        const uint32_t mInner = __ballot_sync(0xFFFFFFFFu, inner);
        uint32_t mLeaf = __ballot_sync(0xFFFFFFFFu, leaf);
        if (sp + __popc(mInner) > PACKET4_STACK) return true;
        if (inner) stack[sp + __popc(mInner & lanesBelow)] = make_uint2(childIdx, entryBits);
        sp += __popc(mInner);

        // ---- the triangles of the hit leaf slots, every undecided lane against its own (original) ray -------------------------
        if (mLeaf) {
            do {
                const int L = 31 - __clz(mLeaf);
                mLeaf &= ~(1u << L);
                const uint32_t tf = __shfl_sync(0xFFFFFFFFu, triFirst, L), tc = __shfl_sync(0xFFFFFFFFu, triCnt, L);
                for (uint32_t k = 0; k < tc; ++k) {
                    const float4* tp = a.tris + (size_t)(tf + k) * 3;
                    const float4 t0 = __ldg(tp), t1 = __ldg(tp + 1), t2 = __ldg(tp + 2);
                    if (COUNT && lane == 0) cTris++;
                    if (live && !occluded) {
                        float u, v, t, aa;
                        if (triCandidate(mk3(ox, oy, oz), mk3(dx, dy, dz), mk3(t0.x, t0.y, t0.z), mk3(t1.x, t1.y, t1.z), mk3(t2.x, t2.y, t2.z), u, v, t, aa)) {
                            if (t > 0.0f && fbits(t0.w) != prev && t < tmax) occluded = true;   // SH/primitive.glsl:268, SH/trace.glsl:97
                        }
                    }
                }
            } while (mLeaf);
            limitBits = __reduce_max_sync(0xFFFFFFFFu, (live && !occluded) ? fbits(tEnd) : 0u);
            if (!__any_sync(0xFFFFFFFFu, live && !occluded)) return false;   // every ray of the beam has its occluder
        }
        __syncwarp();                                    // pushes are visible to the next step's pops
    }
    return false;
}

template <bool COUNT>
__global__ void __launch_bounds__(TRACE_THREADS, RTB_FR_MINBLOCKS) k_trace_cwbvh_beam(const TraceArgs a, const RayQueue fallback) {
    __shared__ uint2 sStack[TRACE_THREADS / 32][PACKET4_STACK];
    const unsigned lane = threadIdx.x & 31u;
    uint2* stack = sStack[threadIdx.x >> 5];
    const uint32_t nRays = a.countPtr ? __ldg(a.countPtr) : a.n;
    unsigned long long cRays = 0, cNodes = 0, cTris = 0, cHits = 0;

    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(a.workCounter, 32u);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (base >= nRays) break;
        const uint32_t slot = base + lane;
        float4 o = make_float4(0.f, 0.f, 0.f, ubits(NO_RAY_HIT)), d = make_float4(0.f, 0.f, 1.f, -1.0f);
        if (slot < nRays) {
            o = __ldg(reinterpret_cast<const float4*>(a.rays + slot));
            d = __ldg(reinterpret_cast<const float4*>(a.rays + slot) + 1);
        }
        const bool live = d.w >= 0.0f;
        const unsigned liveMask = __ballot_sync(0xFFFFFFFFu, live);
        if (!liveMask) continue;
        const uint32_t rayId = slot < nRays ? (a.slotIds ? __ldg(a.slotIds + slot) : slot) : 0u;
        const float ox = o.x, oy = o.y, oz = o.z, dx = d.x, dy = d.y, dz = d.z, tmax = d.w;
        const uint32_t prev = fbits(o.w);
        if (COUNT && live) cRays++;
        bool occluded = false;

        // ---- can the packet be walked as one beam? ------------------------------------------------------------------------------
        const float small = 1e-4f;   // a component this close to zero makes the reciprocal interval useless (and its sign fragile)
        const unsigned ngx = __ballot_sync(0xFFFFFFFFu, live && dx < 0.0f), ngy = __ballot_sync(0xFFFFFFFFu, live && dy < 0.0f), ngz = __ballot_sync(0xFFFFFFFFu, live && dz < 0.0f);
        bool beam = (ngx == 0u || ngx == liveMask) && (ngy == 0u || ngy == liveMask) && (ngz == 0u || ngz == liveMask);
        beam = beam && __all_sync(0xFFFFFFFFu, !live || (fabsf(dx) > small && fabsf(dy) > small && fabsf(dz) > small && isfinite(ox + oy + oz)));
        bool forward = !beam;
        if (beam) {
            const int first = __ffs(liveMask) - 1;
            const bool negX = ngx != 0u, negY = ngy != 0u, negZ = ngz != 0u;
            // dominant axis of the first live ray; the plane through the rearmost origin along it
            const float fdx = fabsf(__shfl_sync(0xFFFFFFFFu, dx, first)), fdy = fabsf(__shfl_sync(0xFFFFFFFFu, dy, first)), fdz = fabsf(__shfl_sync(0xFFFFFFFFu, dz, first));
            const int k = fdx >= fdy ? (fdx >= fdz ? 0 : 2) : (fdy >= fdz ? 1 : 2);
            const float ok = k == 0 ? ox : (k == 1 ? oy : oz), dk = k == 0 ? dx : (k == 1 ? dy : dz);
            const bool negK = k == 0 ? negX : (k == 1 ? negY : negZ);
            const float K = negK ? warpMaxF(ok, live) : warpMinF(ok, live);
            const float s = (K - ok) / dk;                      // <= 0: back along the ray to the plane
            const float t0 = fmaxf(-s, 0.0f);
            const float rx = fmaf(s, dx, ox), ry = fmaf(s, dy, oy), rz = fmaf(s, dz, oz);   // the rebased origin o'
            BeamPlain qp;
            qp.sgx = negX ? -1.0f : 1.0f; qp.sgy = negY ? -1.0f : 1.0f; qp.sgz = negZ ? -1.0f : 1.0f;
            qp.smx = negX ? 0x80000000u : 0u; qp.smy = negY ? 0x80000000u : 0u; qp.smz = negZ ? 0x80000000u : 0u;
            // mirrored origin box, widened by the rounding of the rebasing (a few ulp of the coordinates involved)
            const float mag = warpMaxF(fmaxf(fmaxf(fabsf(rx), fabsf(ry)), fmaxf(fabsf(rz), fmaxf(fabsf(ox), fmaxf(fabsf(oy), fabsf(oz))))), live);
            const float slack = mag * 4.76837158203125e-7f + 1e-30f;   // 2^-21 * magnitude: 4 ulp
            const float mx = rx * qp.sgx, my = ry * qp.sgy, mz = rz * qp.sgz;
            qp.moNx = -(warpMaxF(mx, live) + slack); qp.moFx = -(warpMinF(mx, live) - slack);
            qp.moNy = -(warpMaxF(my, live) + slack); qp.moFy = -(warpMinF(my, live) - slack);
            qp.moNz = -(warpMaxF(mz, live) + slack); qp.moFz = -(warpMinF(mz, live) - slack);
            const float widen = 9.5367431640625e-7f;   // 2^-20
            const float mnx = __uint_as_float(__reduce_min_sync(0xFFFFFFFFu, live ? fbits(fabsf(dx)) : 0x7F800000u)), mxx = __uint_as_float(__reduce_max_sync(0xFFFFFFFFu, live ? fbits(fabsf(dx)) : 0u));
            const float mny = __uint_as_float(__reduce_min_sync(0xFFFFFFFFu, live ? fbits(fabsf(dy)) : 0x7F800000u)), mxy = __uint_as_float(__reduce_max_sync(0xFFFFFFFFu, live ? fbits(fabsf(dy)) : 0u));
            const float mnz = __uint_as_float(__reduce_min_sync(0xFFFFFFFFu, live ? fbits(fabsf(dz)) : 0x7F800000u)), mxz = __uint_as_float(__reduce_max_sync(0xFFFFFFFFu, live ? fbits(fabsf(dz)) : 0u));
            qp.alx = (1.0f / mxx) * (1.0f - widen); qp.ahx = (1.0f / mnx) * (1.0f + widen);
            qp.aly = (1.0f / mxy) * (1.0f - widen); qp.ahy = (1.0f / mny) * (1.0f + widen);
            qp.alz = (1.0f / mxz) * (1.0f - widen); qp.ahz = (1.0f / mnz) * (1.0f + widen);
            // the ray ends at t' = t0 + tmax (infinite for a sun), a little later for the rounding of t0
            const float tEnd = (t0 + tmax) * 1.00001f;
            const uint32_t octLane = (dx < 0.0f ? 0u : 1u) | (dy < 0.0f ? 0u : 2u) | (dz < 0.0f ? 0u : 4u);
            const uint32_t woct = __shfl_sync(0xFFFFFFFFu, octLane, first);
            forward = walkBeamPacket4<COUNT>(a, stack, woct, live, ox, oy, oz, dx, dy, dz, prev, tmax, tEnd, qp, occluded, cNodes, cTris, lane);
            __syncwarp();
        }
        if (forward) queueAppend(fallback, live && !occluded, o, d, rayId);   // uniform branch: the per-ray kernel answers these
        if (live && occluded) {
            const uint32_t sample = rayId / a.fm.localSlots, i = rayId - sample * a.fm.localSlots;
            uint32_t x, y;
            slotToPixel(a.fm, i, x, y);
            atomicOr(a.bits + indexToLight(x, y, a.fm.w, a.fm.h, sample), 1u << ((x & 15u) | ((y & 1u) << 4)));
            if (COUNT) cHits++;
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
