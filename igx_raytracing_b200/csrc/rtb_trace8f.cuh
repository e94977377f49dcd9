// rtb_trace8f.cuh — frustum packets: warp-cooperative nearest-hit traversal of the 8-wide compressed BVH in which the
// BOX tests are done once per packet instead of once per ray (included by rtb_kernels.cu after rtb_trace8p.cuh).
//
// rtb_trace8p.cuh walks the union of the nodes the 32 rays of an 8x4-pixel patch need, but every lane still tests all
// eight child boxes of every node against its own ray: ~130 of the ~190 instructions per node, on a kernel bound by
// instruction issue — although the warp only needs to know whether ANY ray enters a child.  Camera rays of a patch share
// their origin and differ by a few pixel angles, so they are bounded by an interval ray: origin o, direction intervals
// [dmin, dmax] per axis.  One lane tests ONE child box against the interval ray:
//
//     entry >= min over the interval of (near plane - o) / d,   exit <= max over the interval of (far plane - o) / d
//
// with interval end points widened by 2^-20 and the final comparison given 1e-5 of slack.  The test is conservative —
// a child is visited whenever any ray of the interval could enter its (already padded, outward-rounded) box before the
// packet's largest nearest-hit distance — so the set of triangles tested is a superset of what each ray needs, and the
// triangles themselves are tested by every lane with the reference's exact Möller–Trumbore arithmetic and tie rule.
// The hits are therefore bit-identical to the per-ray kernel's (tests/test_gpu_parity.py: packets 0 / 1 / 3 compared
// with array equality, and against the brute-force loop).
//
// Three walks, chosen per packet:
//   walkFrustumPacket4          all 32 rays share the sign of every direction component (99.6 % of a frame): one interval
//                               ray for the patch, 8 lanes per node, FOUR nodes per step
//   walkFrustumPacket<PLAIN>    the patch crosses an axis plane of direction space but each 4x2-pixel quadrant has its
//                               signs: one interval ray per quadrant, 8 children x 4 quadrants = 32 lanes, one node per step
//   walkFrustumPacket<generic>  a quadrant's directions straddle zero on an axis (tested as a wedge there), or the rays of
//                               the packet do not share one origin (every axis unconstrained: correct, only slow)
// The host picks this kernel for the Default projection (one eye) when the patch is small against the leaf nodes
// (rtb_api.cu, primaryPackets).
#pragma once

namespace rtb {

// Measured on B200 (1M-triangle soup, 3840x2160), nearest-hit launch: union packets 3.81 ms -> one node per step, per-quadrant
// interval rays 2.93 ms (143 instructions per node; 3 blocks/SM 3.10 ms, 5 blocks 3.26 ms) -> mirrored fast path 2.43 ms
// -> four nodes per step 1.26 ms (4 blocks/SM; 3 blocks 1.38 ms) -> fused with ray generation and G-buffer finish 1.38 ms
// for what took 0.09 + 1.26 + 0.13.  ncu (profiles/r1z_trace_full.md): issue slots 79 % busy, L1 wavefronts 63 %.
#ifndef RTB_FR_MINBLOCKS
#define RTB_FR_MINBLOCKS 4
#endif

// FUSED: the kernel generates the camera rays itself (raygen.comp:16-37) and finishes the G-buffer in its epilogue
// (trace.glsl:31-63, raygen.comp:39-51) instead of reading 32-byte ray records and writing 16-byte hit records for two
// more kernels to pick up: the same device functions k_raygen and k_finish_primary call, so the same bits.
struct FusedArgs {
    FrameMap fm; CameraRec cam; const SeedRec* seed; SceneView sv; float4* dirT; float4* uvN;
};
RTB_DI void finishGeometry(const SceneView& sv, const Ray& ray, uint32_t prev, const TriHit& th, Hit& hit, vec3& objectNormal,
                           const PrimHit* sph, const PrimHit* cub);   // rtb_kernels.cu

// Per-lane description of the quadrant's interval ray, in the form each of the two walks wants it.
struct QuadPlain {      // every axis has a sign: mirrored so that the quadrant travels towards +
    float sgx, sgy, sgz;          // +-1
    uint32_t smx, smy, smz;       // the same sign as a bit mask for the grid step 2^e
    float mox, moy, moz;          // -origin * sgn
    float alx, ahx, aly, ahy, alz, ahz;   // reciprocal interval of |d|, widened: 0 < al <= ah
};
struct QuadGeneric {    // some axis straddles zero, or the packet has several origins
    float lox, hix, loy, hiy, loz, hiz;   // direction interval of the quadrant
    float fox, foy, foz;
    bool oneOrigin;
};

// One packet through the tree.  PLAIN selects the box test; everything else is shared.
template <bool COUNT, bool PLAIN>
RTB_DI void walkFrustumPacket(const TraceArgs& a, uint2* stack, const uint8_t* permRow, const uint32_t* sSpread, uint32_t woct, uint32_t bitC,
                              uint32_t planeOff, uint32_t halfSel, bool live, float ox, float oy, float oz, float dx, float dy, float dz, uint32_t prev,
                              const QuadPlain& qp, const QuadGeneric& qg, float& best, uint32_t& bestId, float& bu, float& bv,
                              unsigned long long& cNodes, unsigned long long& cTris, unsigned lane) {
    const float INF = __int_as_float(0x7F800000);
    uint32_t limitBits = __reduce_max_sync(0xFFFFFFFFu, live ? fbits(best) : 0u);   // largest nearest-hit distance in the packet
    int sp = 0;
    uint2 G = make_uint2(0u, 0x80000000u);   // the root as a one-node group
    for (;;) {
        // ---- one node for the warp --------------------------------------------------------------------------------
        const uint32_t hits = G.y;
        const uint32_t bit = 31u - (uint32_t)__clz(hits);
        const uint32_t childSlot = (bit - 24u) ^ woct;
        const uint32_t nodeIdx = G.x + (uint32_t)__popc(hits & 0xFFu & ~(0xFFFFFFFFu << childSlot));
        G.y &= ~(1u << bit);
        if (G.y & 0xFF000000u) { if (lane == 0) stack[sp] = G; ++sp; }   // one writer; __syncwarp below orders it before any pop
        const char* p = reinterpret_cast<const char*>(a.nodes8) + (size_t)nodeIdx * 128u;
        uint4 n0, n1;
        ldg256(p, n0, n1);
        const char* q = p + planeOff;
        const uint32_t wlx = __ldg(reinterpret_cast<const uint32_t*>(q)), whx = __ldg(reinterpret_cast<const uint32_t*>(q + 16));
        const uint32_t wly = __ldg(reinterpret_cast<const uint32_t*>(q + 32)), why = __ldg(reinterpret_cast<const uint32_t*>(q + 48));
        const uint32_t wlz = __ldg(reinterpret_cast<const uint32_t*>(q + 64)), whz = __ldg(reinterpret_cast<const uint32_t*>(q + 80));
        if (COUNT && lane == 0) cNodes++;
        // this lane's child box: grid coordinates g (its slot's bf16, exact) -> world, relative to the origin
        const float glx = __uint_as_float(__byte_perm(wlx, 0u, halfSel)), ghx = __uint_as_float(__byte_perm(whx, 0u, halfSel));
        const float gly = __uint_as_float(__byte_perm(wly, 0u, halfSel)), ghy = __uint_as_float(__byte_perm(why, 0u, halfSel));
        const float glz = __uint_as_float(__byte_perm(wlz, 0u, halfSel)), ghz = __uint_as_float(__byte_perm(whz, 0u, halfSel));
        bool hit;
        if (PLAIN) {
            // mirrored: sgn * ((p + g * 2^e) - o); the near plane is the smaller of the two, whatever the sign was
            const float sx = __uint_as_float(((n0.w & 0xFFu) << 23) ^ qp.smx), sy = __uint_as_float(((n0.w << 15) & 0x7F800000u) ^ qp.smy), sz = __uint_as_float(((n0.w << 7) & 0x7F800000u) ^ qp.smz);
            const float px = fmaf(__uint_as_float(n0.x), qp.sgx, qp.mox), py = fmaf(__uint_as_float(n0.y), qp.sgy, qp.moy), pz = fmaf(__uint_as_float(n0.z), qp.sgz, qp.moz);
            const float ax = fmaf(glx, sx, px), bx = fmaf(ghx, sx, px), ay = fmaf(gly, sy, py), by = fmaf(ghy, sy, py), az = fmaf(glz, sz, pz), bz = fmaf(ghz, sz, pz);
            const float nx = fminf(ax, bx), fx = fmaxf(ax, bx), ny = fminf(ay, by), fy = fmaxf(ay, by), nz = fminf(az, bz), fz = fmaxf(az, bz);
            const float ex = fminf(nx * qp.alx, nx * qp.ahx), xx = fmaxf(fx * qp.alx, fx * qp.ahx);
            const float ey = fminf(ny * qp.aly, ny * qp.ahy), xy = fmaxf(fy * qp.aly, fy * qp.ahy);
            const float ez = fminf(nz * qp.alz, nz * qp.ahz), xz = fmaxf(fz * qp.alz, fz * qp.ahz);
            const float entry = fmaxf(fmaxf(ex, ey), fmaxf(ez, 0.0f));
            const float exit = fminf(fminf(xx, xy), fminf(xz, __uint_as_float(limitBits)));
            hit = entry <= fmaf(exit, 1.00001f, 1e-30f);
        } else {
            // rare packets (a quadrant crosses an axis plane of direction space; rays-in packets with several origins):
            // nothing is kept in registers across nodes, the interval set-up is redone here
            const float tiny = 8.271806e-25f, widen = 9.5367431640625e-7f;   // 2^-80, 2^-20
            const bool zeroX = !(qg.lox > tiny || qg.hix < -tiny), zeroY = !(qg.loy > tiny || qg.hiy < -tiny), zeroZ = !(qg.loz > tiny || qg.hiz < -tiny);
            const bool freeX = !qg.oneOrigin || zeroX, freeY = !qg.oneOrigin || zeroY, freeZ = !qg.oneOrigin || zeroZ;
            float ilx = freeX ? 0.0f : 1.0f / qg.hix, ihx = freeX ? 0.0f : 1.0f / qg.lox;
            float ily = freeY ? 0.0f : 1.0f / qg.hiy, ihy = freeY ? 0.0f : 1.0f / qg.loy;
            float ilz = freeZ ? 0.0f : 1.0f / qg.hiz, ihz = freeZ ? 0.0f : 1.0f / qg.loz;
            ilx -= fabsf(ilx) * widen; ihx += fabsf(ihx) * widen;
            ily -= fabsf(ily) * widen; ihy += fabsf(ihy) * widen;
            ilz -= fabsf(ilz) * widen; ihz += fabsf(ihz) * widen;
            const bool negX = qg.hix < 0.0f, negY = qg.hiy < 0.0f, negZ = qg.hiz < 0.0f;
            const float sx = __uint_as_float((n0.w & 0xFFu) << 23), sy = __uint_as_float((n0.w << 15) & 0x7F800000u), sz = __uint_as_float((n0.w << 7) & 0x7F800000u);
            const float px = __uint_as_float(n0.x) - qg.fox, py = __uint_as_float(n0.y) - qg.foy, pz = __uint_as_float(n0.z) - qg.foz;
            const float Wlx = fmaf(glx, sx, px), Whx = fmaf(ghx, sx, px), Wly = fmaf(gly, sy, py), Why = fmaf(ghy, sy, py), Wlz = fmaf(glz, sz, pz), Whz = fmaf(ghz, sz, pz);
            const float nx = negX ? Whx : Wlx, fx = negX ? Wlx : Whx, ny = negY ? Why : Wly, fy = negY ? Wly : Why, nz = negZ ? Whz : Wlz, fz = negZ ? Wlz : Whz;
            const float ex = freeX ? -INF : fminf(nx * ilx, nx * ihx), xx = freeX ? INF : fmaxf(fx * ilx, fx * ihx);
            const float ey = freeY ? -INF : fminf(ny * ily, ny * ihy), xy = freeY ? INF : fmaxf(fy * ily, fy * ihy);
            const float ez = freeZ ? -INF : fminf(nz * ilz, nz * ihz), xz = freeZ ? INF : fmaxf(fz * ilz, fz * ihz);
            const float entry = fmaxf(fmaxf(ex, ey), fmaxf(ez, 0.0f));
            const float exit = fminf(fminf(xx, xy), fminf(xz, __uint_as_float(limitBits)));
            hit = entry <= exit * 1.00001f + 1e-30f;
            if (qg.oneOrigin) {
                // an axis whose direction interval contains zero has no reciprocal interval; it is tested as a wedge: up to the
                // exit distance t1 the other axes give, the quadrant's rays stay within [t1 * min(lo, 0), t1 * max(hi, 0)]
                const float t1 = fminf(exit, 1e30f) * 1.00001f;
                if (zeroX) hit = hit && Wlx <= t1 * fmaxf(qg.hix, 0.0f) + 1e-30f && Whx >= t1 * fminf(qg.lox, 0.0f) - 1e-30f;
                if (zeroY) hit = hit && Wly <= t1 * fmaxf(qg.hiy, 0.0f) + 1e-30f && Why >= t1 * fminf(qg.loy, 0.0f) - 1e-30f;
                if (zeroZ) hit = hit && Wlz <= t1 * fmaxf(qg.hiz, 0.0f) + 1e-30f && Whz >= t1 * fminf(qg.loz, 0.0f) - 1e-30f;
            }
        }
        const uint32_t any8 = __reduce_or_sync(0xFFFFFFFFu, hit ? bitC : 0u);
        const uint32_t imask = n0.w >> 24;
        const uint32_t top = permRow[any8 & imask];
        const uint32_t P = n1.z & 0x00FFFFFFu;
        uint32_t T = sSpread[any8 & ~imask & 0xFFu] & P;

        // ---- the triangles of the hit leaf slots, every lane against its own ray --------------------------------------
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

        // ---- descend, or pop ------------------------------------------------------------------------------------------
        if (top) G = make_uint2(n1.x, (top << 24) | imask);
        else if (sp > 0) { __syncwarp(); --sp; G = stack[sp]; __syncwarp(); }   // (the second barrier: the slot is not rewritten before every lane has read it)
        else break;
    }
}


// ---- four nodes per step ---------------------------------------------------------------------------------------------
// When all 32 rays of the packet share the sign of every direction component (all but the patches that cross an axis
// plane of direction space), ONE interval ray bounds the whole patch, and since the warp descends into a child as soon as
// any quadrant hits it, nothing is lost by testing the patch instead of its quadrants.  That frees three quarters of the
// lanes: 8 lanes test the 8 children of one node, and the warp works on FOUR nodes per step.  The traversal state is a
// per-warp stack of (node, entry distance) in shared memory; a step pops up to four entries (nearest on top), tests
// 4 x 8 child boxes, pushes the hit inner children with one parallel store (positions from a ballot) and tests the
// triangles of the hit leaf slots on every lane.  Entries whose entry distance has fallen behind the packet's largest
// nearest-hit distance are dropped when popped.
#ifndef RTB_PK4_STACK
#define RTB_PK4_STACK 256
#endif
constexpr int PACKET4_STACK = RTB_PK4_STACK;     // entries per warp; near the top a step pops one node only, and a packet that would
                                       // still overflow (no real tree does) is handed to the one-node walk, whose stack is
                                       // bounded by the tree depth

template <bool COUNT>
RTB_DI bool walkFrustumPacket4(const TraceArgs& a, uint2* stack, uint32_t woct, bool live, float ox, float oy, float oz, float dx, float dy, float dz,
                               uint32_t prev, const QuadPlain& qp, float& best, uint32_t& bestId, float& bu, float& bv,
                               unsigned long long& cNodes, unsigned long long& cTris, unsigned lane) {
    const uint32_t lanesBelow = (1u << lane) - 1u;
    const uint32_t g = 3u - (lane >> 3);                 // which popped entry this lane works on: lanes 24..31 take the top of the stack
    const uint32_t cs = (lane & 7u) ^ woct;              // its child slot: within a group, a higher lane is a nearer child
    const uint32_t planeOff = 32u + (cs >> 1) * 4u;
    const uint32_t halfSel = (cs & 1u) ? 0x1044u : 0x3244u;
    const uint32_t csBelow = (1u << cs) - 1u, triShift = 3u * cs;
    uint32_t limitBits = __reduce_max_sync(0xFFFFFFFFu, live ? fbits(best) : 0u);
    if (lane == 0) stack[0] = make_uint2(0u, 0u);        // the root, entry distance 0
    int sp = 1;
    __syncwarp();
    while (sp > 0) {
        const int nPop = sp > PACKET4_STACK - 40 ? 1 : min(sp, 4);
        uint2 e = make_uint2(0u, 0xFFFFFFFFu);
        if ((int)g < nPop) e = stack[sp - 1 - (int)g];
        sp -= nPop;
        const bool activeNode = e.y <= limitBits;        // non-negative floats order like their bit patterns
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
            const float sx = __uint_as_float(((n0.w & 0xFFu) << 23) ^ qp.smx), sy = __uint_as_float(((n0.w << 15) & 0x7F800000u) ^ qp.smy), sz = __uint_as_float(((n0.w << 7) & 0x7F800000u) ^ qp.smz);
            const float px = fmaf(__uint_as_float(n0.x), qp.sgx, qp.mox), py = fmaf(__uint_as_float(n0.y), qp.sgy, qp.moy), pz = fmaf(__uint_as_float(n0.z), qp.sgz, qp.moz);
            const float ax = fmaf(glx, sx, px), bx = fmaf(ghx, sx, px), ay = fmaf(gly, sy, py), by = fmaf(ghy, sy, py), az = fmaf(glz, sz, pz), bz = fmaf(ghz, sz, pz);
            const float nx = fminf(ax, bx), fx = fmaxf(ax, bx), ny = fminf(ay, by), fy = fmaxf(ay, by), nz = fminf(az, bz), fz = fmaxf(az, bz);
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
        if (sp + __popc(mInner) > PACKET4_STACK) return true;   // uniform: hits found so far stay valid, the caller restarts from the root
        if (inner) stack[sp + __popc(mInner & lanesBelow)] = make_uint2(childIdx, entryBits);
        sp += __popc(mInner);

        // ---- the triangles of the hit leaf slots, every lane against its own ray --------------------------------------
        if (mLeaf) {
            do {
                const int L = 31 - __clz(mLeaf);
                mLeaf &= ~(1u << L);
                const uint32_t tf = __shfl_sync(0xFFFFFFFFu, triFirst, L), tc = __shfl_sync(0xFFFFFFFFu, triCnt, L);
                for (uint32_t k = 0; k < tc; ++k) {
                    const float4* tp = a.tris + (size_t)(tf + k) * 3;
                    const float4 t0 = __ldg(tp), t1 = __ldg(tp + 1), t2 = __ldg(tp + 2);
                    if (COUNT && lane == 0) cTris++;
                    float u, v, t, aa;
                    if (triCandidate(mk3(ox, oy, oz), mk3(dx, dy, dz), mk3(t0.x, t0.y, t0.z), mk3(t1.x, t1.y, t1.z), mk3(t2.x, t2.y, t2.z), u, v, t, aa)) {
                        const uint32_t id = fbits(t0.w);
                        // reference: strict t < hitT in index order => on equal t the lower index wins
                        if (t > 0.0f && id != prev && (t < best || (t == best && id < bestId))) { best = t; bestId = id; bu = u; bv = v; }
                    }
                }
            } while (mLeaf);
            limitBits = __reduce_max_sync(0xFFFFFFFFu, live ? fbits(best) : 0u);
        }
        __syncwarp();                                    // pushes are visible to the next step's pops
    }
    return false;
}

template <bool COUNT, bool FUSED>
__global__ void __launch_bounds__(TRACE_THREADS, RTB_FR_MINBLOCKS) k_trace_cwbvh_frustum(const TraceArgs a, const FusedArgs f) {
    __shared__ uint2 sStack[TRACE_THREADS / 32][PACKET4_STACK];
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
    const uint32_t planeOff = 32u + (c >> 1) * 4u;            // byte offset of the word holding slot c's lo plane on x
    const uint32_t halfSel = (c & 1u) ? 0x1044u : 0x3244u;    // PRMT selector: this slot's bf16 (odd slots: lower half of the word) into the upper half, zeros below
    unsigned long long cRays = 0, cNodes = 0, cTris = 0, cHits = 0;

    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(a.workCounter, 32u);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (base >= a.n) break;
        const uint32_t slot = base + lane;
        float4 o = make_float4(0.f, 0.f, 0.f, ubits(NO_RAY_HIT)), d = make_float4(0.f, 0.f, 1.f, -1.0f);
        uint32_t pxX = 0, pxY = 0;
        if (FUSED) {
            if (slot < a.n && slotToPixel(f.fm, slot, pxX, pxY)) {
                const Ray r = calculatePrimary(f.cam, pxX, pxY, mk2(__ldg(&f.seed->randomX), __ldg(&f.seed->randomY)));
                o = make_float4(r.pos.x, r.pos.y, r.pos.z, ubits(NO_RAY_HIT));
                d = make_float4(r.dir.x, r.dir.y, r.dir.z, NO_HIT);
            }
        } else if (slot < a.n) {
            o = __ldg(reinterpret_cast<const float4*>(a.rays + slot));
            d = __ldg(reinterpret_cast<const float4*>(a.rays + slot) + 1);
        }
        const bool live = d.w >= 0.0f;
        const unsigned liveMask = __ballot_sync(0xFFFFFFFFu, live);
        const float ox = o.x, oy = o.y, oz = o.z, dx = d.x, dy = d.y, dz = d.z;
        const uint32_t prev = fbits(o.w);
        float best = live ? d.w : -1.0f;   // a negative limit fails every triangle test
        float bu = 0.0f, bv = 0.0f;
        uint32_t bestId = NO_RAY_HIT;
        if (liveMask) {
            if (COUNT && live) cRays++;
            // ---- the quadrant's interval ray ------------------------------------------------------------------------
            const float INF = __int_as_float(0x7F800000);
            QuadGeneric qg;
            qg.lox = live ? dx : INF; qg.hix = live ? dx : -INF; qg.loy = live ? dy : INF; qg.hiy = live ? dy : -INF; qg.loz = live ? dz : INF; qg.hiz = live ? dz : -INF;
            #pragma unroll
            for (int s = 0; s < 3; ++s) {
                const int m = s == 0 ? 1 : (s == 1 ? 2 : 8);
                qg.lox = fminf(qg.lox, __shfl_xor_sync(0xFFFFFFFFu, qg.lox, m)); qg.hix = fmaxf(qg.hix, __shfl_xor_sync(0xFFFFFFFFu, qg.hix, m));
                qg.loy = fminf(qg.loy, __shfl_xor_sync(0xFFFFFFFFu, qg.loy, m)); qg.hiy = fmaxf(qg.hiy, __shfl_xor_sync(0xFFFFFFFFu, qg.hiy, m));
                qg.loz = fminf(qg.loz, __shfl_xor_sync(0xFFFFFFFFu, qg.loz, m)); qg.hiz = fmaxf(qg.hiz, __shfl_xor_sync(0xFFFFFFFFu, qg.hiz, m));
            }
            const int first = __ffs(liveMask) - 1;
            qg.fox = __shfl_sync(0xFFFFFFFFu, ox, first); qg.foy = __shfl_sync(0xFFFFFFFFu, oy, first); qg.foz = __shfl_sync(0xFFFFFFFFu, oz, first);
            qg.oneOrigin = __all_sync(0xFFFFFFFFu, !live || (ox == qg.fox && oy == qg.foy && oz == qg.foz));
            const bool quadLive = qg.lox <= qg.hix;                // some live ray in this lane's quadrant
            const float tiny = 8.271806e-25f, widen = 9.5367431640625e-7f;   // 2^-80, 2^-20
            const bool zeroX = !(qg.lox > tiny || qg.hix < -tiny), zeroY = !(qg.loy > tiny || qg.hiy < -tiny), zeroZ = !(qg.loz > tiny || qg.hiz < -tiny);
            const bool plain = __all_sync(0xFFFFFFFFu, qg.oneOrigin && !(quadLive && (zeroX || zeroY || zeroZ)));
            const uint32_t bitC = quadLive ? (1u << c) : 0u;
            // child order for the warp: octant of the first live ray
            const uint32_t octLane = (dx < 0.0f ? 0u : 1u) | (dy < 0.0f ? 0u : 2u) | (dz < 0.0f ? 0u : 4u);
            const uint32_t woct = __shfl_sync(0xFFFFFFFFu, octLane, first);
            const uint8_t* permRow = sPerm[woct];
            QuadPlain qp;
            // the whole patch on one side of every axis plane of direction space: one interval ray for all 32 rays
            const unsigned ngx = __ballot_sync(0xFFFFFFFFu, live && dx < 0.0f), ngy = __ballot_sync(0xFFFFFFFFu, live && dy < 0.0f), ngz = __ballot_sync(0xFFFFFFFFu, live && dz < 0.0f);
            const bool plain4 = plain && (ngx == 0u || ngx == liveMask) && (ngy == 0u || ngy == liveMask) && (ngz == 0u || ngz == liveMask);
            if (plain4) {
                const bool negX = ngx != 0u, negY = ngy != 0u, negZ = ngz != 0u;
                qp.sgx = negX ? -1.0f : 1.0f; qp.sgy = negY ? -1.0f : 1.0f; qp.sgz = negZ ? -1.0f : 1.0f;
                qp.smx = negX ? 0x80000000u : 0u; qp.smy = negY ? 0x80000000u : 0u; qp.smz = negZ ? 0x80000000u : 0u;
                qp.mox = -qg.fox * qp.sgx; qp.moy = -qg.foy * qp.sgy; qp.moz = -qg.foz * qp.sgz;
                // |d| over the live rays of the patch (positive floats order like their bit patterns)
                const float mnx = __uint_as_float(__reduce_min_sync(0xFFFFFFFFu, live ? fbits(fabsf(dx)) : 0x7F800000u)), mxx = __uint_as_float(__reduce_max_sync(0xFFFFFFFFu, live ? fbits(fabsf(dx)) : 0u));
                const float mny = __uint_as_float(__reduce_min_sync(0xFFFFFFFFu, live ? fbits(fabsf(dy)) : 0x7F800000u)), mxy = __uint_as_float(__reduce_max_sync(0xFFFFFFFFu, live ? fbits(fabsf(dy)) : 0u));
                const float mnz = __uint_as_float(__reduce_min_sync(0xFFFFFFFFu, live ? fbits(fabsf(dz)) : 0x7F800000u)), mxz = __uint_as_float(__reduce_max_sync(0xFFFFFFFFu, live ? fbits(fabsf(dz)) : 0u));
                qp.alx = (1.0f / mxx) * (1.0f - widen); qp.ahx = (1.0f / mnx) * (1.0f + widen);
                qp.aly = (1.0f / mxy) * (1.0f - widen); qp.ahy = (1.0f / mny) * (1.0f + widen);
                qp.alz = (1.0f / mxz) * (1.0f - widen); qp.ahz = (1.0f / mnz) * (1.0f + widen);
                if (walkFrustumPacket4<COUNT>(a, stack, woct, live, ox, oy, oz, dx, dy, dz, prev, qp, best, bestId, bu, bv, cNodes, cTris, lane)) {
                    __syncwarp();
                    walkFrustumPacket<COUNT, true>(a, stack, permRow, sSpread, woct, bitC, planeOff, halfSel, live, ox, oy, oz, dx, dy, dz, prev, qp, qg, best, bestId, bu, bv, cNodes, cTris, lane);
                }
            } else if (plain) {
                const bool negX = qg.hix < 0.0f, negY = qg.hiy < 0.0f, negZ = qg.hiz < 0.0f;   // the quadrant travels towards - on that axis
                qp.sgx = negX ? -1.0f : 1.0f; qp.sgy = negY ? -1.0f : 1.0f; qp.sgz = negZ ? -1.0f : 1.0f;
                qp.smx = negX ? 0x80000000u : 0u; qp.smy = negY ? 0x80000000u : 0u; qp.smz = negZ ? 0x80000000u : 0u;
                qp.mox = -qg.fox * qp.sgx; qp.moy = -qg.foy * qp.sgy; qp.moz = -qg.foz * qp.sgz;
                // |d| in [min, max] -> 1 / |d| in [1 / max, 1 / min]; a dead quadrant (lo = +inf, hi = -inf) gets a harmless interval
                const float mnx = quadLive ? fminf(fabsf(qg.lox), fabsf(qg.hix)) : 1.0f, mxx = quadLive ? fmaxf(fabsf(qg.lox), fabsf(qg.hix)) : 1.0f;
                const float mny = quadLive ? fminf(fabsf(qg.loy), fabsf(qg.hiy)) : 1.0f, mxy = quadLive ? fmaxf(fabsf(qg.loy), fabsf(qg.hiy)) : 1.0f;
                const float mnz = quadLive ? fminf(fabsf(qg.loz), fabsf(qg.hiz)) : 1.0f, mxz = quadLive ? fmaxf(fabsf(qg.loz), fabsf(qg.hiz)) : 1.0f;
                qp.alx = (1.0f / mxx) * (1.0f - widen); qp.ahx = (1.0f / mnx) * (1.0f + widen);
                qp.aly = (1.0f / mxy) * (1.0f - widen); qp.ahy = (1.0f / mny) * (1.0f + widen);
                qp.alz = (1.0f / mxz) * (1.0f - widen); qp.ahz = (1.0f / mnz) * (1.0f + widen);
                walkFrustumPacket<COUNT, true>(a, stack, permRow, sSpread, woct, bitC, planeOff, halfSel, live, ox, oy, oz, dx, dy, dz, prev, qp, qg, best, bestId, bu, bv, cNodes, cTris, lane);
            } else {
                walkFrustumPacket<COUNT, false>(a, stack, permRow, sSpread, woct, bitC, planeOff, halfSel, live, ox, oy, oz, dx, dy, dz, prev, qp, qg, best, bestId, bu, bv, cNodes, cTris, lane);
            }
        }
        TriHit h; h.t = bestId == NO_RAY_HIT ? NO_HIT : best; h.id = bestId; h.u = bu; h.v = bv;
        if (FUSED) {
            if (live) {   // k_finish_primary, on registers
                Ray ray; ray.pos = mk3(ox, oy, oz); ray.dir = mk3(dx, dy, dz);
                Hit hit; vec3 objectNormal;
                finishGeometry(f.sv, ray, NO_RAY_HIT, h, hit, objectNormal, nullptr, nullptr);
                float4 out0;
                if (hit.hitT == NO_HIT) out0 = make_float4(ray.dir.x, ray.dir.y, ray.dir.z, ubits(NO_RAY_HIT));
                else { const vec3 v = ray.dir * hit.hitT; out0 = make_float4(v.x, v.y, v.z, ubits(hit.object)); }
                uint32_t ex, ey;
                encodeNormalGpu(objectNormal, ex, ey);
                const size_t px = (size_t)pxY * f.fm.w + pxX;
                f.dirT[px] = out0;
                if (!f.sv.releaseBuild || hit.hitT < NO_HIT) f.uvN[px] = make_float4(hit.uv.x, hit.uv.y, ubits(ex), ubits(ey));
            }
        } else if (slot < a.n) {
            *reinterpret_cast<float4*>(a.hits + slot) = *reinterpret_cast<float4*>(&h);
        }
        if (COUNT && slot < a.n && bestId != NO_RAY_HIT) cHits++;
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
