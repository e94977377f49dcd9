// rtb_kernels.cu — the sm_100a kernels of the hot path.
//
//   k_init            init.comp                     (SH/init.comp:12-18)
//   k_raygen          primary-ray generation        (SH/raygen.comp:16-37, SH/camera.glsl)
//   k_trace_bvh       persistent, warp-refilled, stack-based BVH traversal with Möller–Trumbore fused
//                     into the leaf loop (replaces the triangle loop of SH/trace.glsl:25-29 / :78-81)
//   k_trace_brute     the reference's linear triangle loop, shared-memory tiled (GPU-side cross-check)
//   k_finish_primary  sphere/cube/plane loops, normal interpolation, G-buffer stores (SH/trace.glsl:31-63,
//                     SH/raygen.comp:39-51)
//   k_shadowgen       shadow-ray set-up + non-triangle occluders (SH/nv_all.shadow.comp:50-143)
//   k_shade           lighting.comp + composite.comp fused, coalesced rgba8 store
//
// "SH/" = res/shaders/ of the reference.  Compiled with -fmad=false (see rtb_math.cuh).
#include "rtb_kernels.cuh"
#include "rtb_math.cuh"

namespace rtb {

// --------------------------------------------------------------------------------------------------------
// slot <-> pixel
// --------------------------------------------------------------------------------------------------------
RTB_DI bool slotToPixel(const FrameMap& fm, uint32_t i, uint32_t& x, uint32_t& y) {
    const uint32_t k = i >> 10, s = (i >> 5) & 31u, lane = i & 31u;
    const uint32_t g = k * fm.nranks + fm.rank;
    if (g >= fm.blocksX * fm.blocksY) return false;
    const uint32_t bx = g % fm.blocksX, by = g / fm.blocksX;
    x = bx * 32u + (s & 3u) * 8u + (lane & 7u);
    y = by * 32u + (s >> 2) * 4u + (lane >> 3);
    return x < fm.w && y < fm.h;
}
RTB_DI uint32_t tiledSlot(const FrameMap& fm, uint32_t i) { return (((i >> 10) * fm.tiledMul + fm.tiledAdd) << 10) | (i & 1023u); }

// append one ray to a queue with one atomic per warp; all 32 lanes of the warp must call
RTB_DI void queueAppend(const RayQueue& q, bool live, float4 ro, float4 rd, uint32_t slot) {
    const unsigned m = __ballot_sync(0xFFFFFFFFu, live);
    if (!m) return;
    const unsigned lane = threadIdx.x & 31u;
    const int leader = __ffs(m) - 1;
    uint32_t base = 0;
    if ((int)lane == leader) base = atomicAdd(q.count, (uint32_t)__popc(m));
    base = __shfl_sync(0xFFFFFFFFu, base, leader);
    if (!live) return;
    const uint32_t r = base + (uint32_t)__popc(m & ((1u << lane) - 1u));
    float4* out = reinterpret_cast<float4*>(q.rays + r);
    out[0] = ro; out[1] = rd;
    q.slotIds[r] = slot;
}

// --------------------------------------------------------------------------------------------------------
// K0
// --------------------------------------------------------------------------------------------------------
// `snap` (may be null): a second copy of the result — the copy the frame's other launches read when consecutive frames overlap
// (rtb_api.cu, frame overlap): the next frame's K0 then rewrites `seed` while this frame's shadow and shade launches still run
__global__ void k_init(SeedRec* seed, SeedRec* snap) {
    SeedRec s = *seed;
    vec2 off = rand2(mk2(s.cpuOffsetX, s.cpuOffsetY) + (float)s.sampleCount);
    s.randomX = off.x; s.randomY = off.y;
    ++s.sampleCount; ++s.sampleOffset;
    *seed = s;
    if (snap) *snap = s;
}
void launch_init(SeedRec* seed, cudaStream_t st, SeedRec* snap) { k_init<<<1, 1, 0, st>>>(seed, snap); }

// --------------------------------------------------------------------------------------------------------
// K1a: primary rays
// --------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_raygen(const FrameMap fm, const CameraRec cam, const SeedRec* __restrict__ seed,
                                                RayRec* __restrict__ rays) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= fm.localSlots) return;
    uint32_t x, y;
    float4 o, d;
    if (slotToPixel(fm, i, x, y)) {
        const vec2 rnd = mk2(__ldg(&seed->randomX), __ldg(&seed->randomY));
        const Ray r = calculatePrimary(cam, x, y, rnd);
        o = make_float4(r.pos.x, r.pos.y, r.pos.z, ubits(NO_RAY_HIT));
        d = make_float4(r.dir.x, r.dir.y, r.dir.z, NO_HIT);
    } else {
        o = make_float4(0.f, 0.f, 0.f, ubits(NO_RAY_HIT));
        d = make_float4(0.f, 0.f, 1.f, -1.0f);
    }
    float4* out = reinterpret_cast<float4*>(rays + i);
    out[0] = o; out[1] = d;
}
void launch_raygen(const FrameMap& fm, const CameraRec* cam, const SeedRec* seed, RayRec* rays, cudaStream_t st) {
    if (!fm.localSlots) return;
    k_raygen<<<(fm.localSlots + 255) / 256, 256, 0, st>>>(fm, *cam, seed, rays);
}

// --------------------------------------------------------------------------------------------------------
// BVH traversal
// --------------------------------------------------------------------------------------------------------
constexpr int TRACE_THREADS = 256;
constexpr int TOP_NODES = 256;              // top of the tree (breadth-first prefix) staged in shared memory: 16 KB
constexpr int STACK_SIZE = 64;              // the builder bounds the depth to STACK_SIZE - 2
constexpr int REFILL_THRESHOLD = 20;        // refill when fewer lanes than this still traverse
constexpr int SENTINEL = 0x7FFFFFFF;

enum { MODE_CLOSEST = 0, MODE_ANY_BITS = 1, MODE_ANY_BYTES = 2 };

struct TraceArgs {
    const RayRec* rays; uint32_t n;
    const float4* nodes; const uint4* nodes8; const float4* tris; uint32_t nodeCount;
    const uint4* nodes8Alias;   // == nodes8, a second name the compiler cannot prove equal (rtb_trace8.cuh, ldg256swap)
    TriHit* hits;               // MODE_CLOSEST
    uint32_t* bits; FrameMap fm;// MODE_ANY_BITS
    const uint32_t* slotIds;    // MODE_ANY_BITS on a queue: the wavefront slot of ray r (nullptr: r itself)
    const uint32_t* countPtr;   // rays in the queue (device; nullptr: n)
    uint8_t* bytes;             // MODE_ANY_BYTES
    uint32_t* workCounter;
    TraceCounters* counters;
};

template <int MODE, bool COUNT>
__global__ void __launch_bounds__(TRACE_THREADS) k_trace_bvh(const TraceArgs a) {
    __shared__ float4 sTop[TOP_NODES * 4];
    const int topN = min((int)a.nodeCount, TOP_NODES);
    for (int i = threadIdx.x; i < topN * 4; i += TRACE_THREADS) sTop[i] = __ldg(a.nodes + i);
    __syncthreads();

    const unsigned lane = threadIdx.x & 31u;
    const unsigned lanesBelow = (1u << lane) - 1u;
    int stack[STACK_SIZE];
    const uint32_t nRays = a.countPtr ? __ldg(a.countPtr) : a.n;   // a queue's length lives on the device

    bool active = false, exhausted = false;
    uint32_t slot = 0, prev = 0, bestId = NO_RAY_HIT;
    float ox = 0, oy = 0, oz = 0, dx = 0, dy = 0, dz = 0, idx = 0, idy = 0, idz = 0, oodx = 0, oody = 0, oodz = 0;
    float best = 0, bu = 0, bv = 0;
    int node = SENTINEL, sp = 0;
    unsigned long long cRays = 0, cNodes = 0, cTris = 0, cHits = 0;

    for (;;) {
        // ---- refill: idle lanes take the next rays of the wavefront (one atomic per warp) -------------
        const unsigned need = __ballot_sync(0xFFFFFFFFu, !active);
        if (need && !exhausted) {
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
                        oodx = ox * idx; oody = oy * idy; oodz = oz * idz;
                        bestId = NO_RAY_HIT; bu = 0.0f; bv = 0.0f;
                        stack[0] = SENTINEL; sp = 0; node = 0;
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

        // ---- traverse in lock-step until too few lanes are left -----------------------------------------
        for (;;) {
            if (active) {
                // inner nodes: both children's slabs come with the node
                while ((unsigned)node < (unsigned)SENTINEL) {
                    float4 n0, n1, n2, n3;
                    if (node < topN) {
                        const float4* p = sTop + node * 4;
                        n0 = p[0]; n1 = p[1]; n2 = p[2]; n3 = p[3];
                    } else {
                        const float4* p = a.nodes + (size_t)node * 4;
                        n0 = __ldg(p); n1 = __ldg(p + 1); n2 = __ldg(p + 2); n3 = __ldg(p + 3);
                    }
                    if (COUNT) cNodes++;
                    const float c0lox = fmaf(n0.x, idx, -oodx), c0hix = fmaf(n0.y, idx, -oodx);
                    const float c0loy = fmaf(n0.z, idy, -oody), c0hiy = fmaf(n0.w, idy, -oody);
                    const float c1lox = fmaf(n1.x, idx, -oodx), c1hix = fmaf(n1.y, idx, -oodx);
                    const float c1loy = fmaf(n1.z, idy, -oody), c1hiy = fmaf(n1.w, idy, -oody);
                    const float c0loz = fmaf(n2.x, idz, -oodz), c0hiz = fmaf(n2.y, idz, -oodz);
                    const float c1loz = fmaf(n2.z, idz, -oodz), c1hiz = fmaf(n2.w, idz, -oodz);
                    const float c0min = fmaxf(fmaxf(fminf(c0lox, c0hix), fminf(c0loy, c0hiy)), fmaxf(fminf(c0loz, c0hiz), 0.0f));
                    const float c0max = fminf(fminf(fmaxf(c0lox, c0hix), fmaxf(c0loy, c0hiy)), fminf(fmaxf(c0loz, c0hiz), best));
                    const float c1min = fmaxf(fmaxf(fminf(c1lox, c1hix), fminf(c1loy, c1hiy)), fmaxf(fminf(c1loz, c1hiz), 0.0f));
                    const float c1max = fminf(fminf(fmaxf(c1lox, c1hix), fmaxf(c1loy, c1hiy)), fminf(fmaxf(c1loz, c1hiz), best));
                    const bool h0 = c0max >= c0min, h1 = c1max >= c1min;
                    const int c0 = __float_as_int(n3.x), c1 = __float_as_int(n3.y);
                    if (!h0 && !h1) {
                        node = stack[sp]; sp--;
                    } else {
                        node = h0 ? c0 : c1;
                        if (h0 && h1) {
                            int far = c1;
                            if (c1min < c0min) { far = c0; node = c1; }
                            stack[++sp] = far;
                        }
                    }
                }
                // one leaf
                if (node < 0) {
                    const uint32_t link = ~(uint32_t)node;
                    const uint32_t first = link >> 3, count = (link & 7u) + 1u;
                    const vec3 ro = mk3(ox, oy, oz), rd = mk3(dx, dy, dz);
                    for (uint32_t k = 0; k < count; ++k) {
                        const float4* tp = a.tris + (size_t)(first + k) * 3;
                        const float4 t0 = __ldg(tp), t1 = __ldg(tp + 1), t2 = __ldg(tp + 2);
                        if (COUNT) cTris++;
                        float u, v, t, aa;
                        if (!triCandidate(ro, rd, mk3(t0.x, t0.y, t0.z), mk3(t1.x, t1.y, t1.z), mk3(t2.x, t2.y, t2.z), u, v, t, aa)) continue;
                        const uint32_t id = fbits(t0.w);
                        if (!(t > 0.0f) || id == prev) continue;
                        if (MODE == MODE_CLOSEST) {
                            // reference: strict t < hitT in index order => on equal t the lower index wins
                            if (t < best || (t == best && id < bestId)) { best = t; bestId = id; bu = u; bv = v; }
                        } else {
                            if (t < best) { bestId = id; break; }
                        }
                    }
                    if (MODE != MODE_CLOSEST && bestId != NO_RAY_HIT) node = SENTINEL;
                    else { node = stack[sp]; sp--; }
                }
                if (node == SENTINEL) {
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
            if (__popc(__ballot_sync(0xFFFFFFFFu, active)) < REFILL_THRESHOLD) break;
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
#include "rtb_trace8.cuh"
#include "rtb_trace8s.cuh"
#include "rtb_trace8p.cuh"
#include "rtb_trace8f.cuh"
#include "rtb_trace8b.cuh"
namespace rtb {

// persistent grid: one resident wave (SM count x blocks that fit per SM), fewer when the wavefront is small.  The size is cached
// per kernel AND per device (a process may drive contexts on several devices)
struct OccCache { int v[32] = {}; int& get() { int d = 0; cudaGetDevice(&d); return v[d & 31]; } };
static int g_traceBlocks = 0;
template <int MODE, bool COUNT, bool WIDE>
static void launchTraceBvh(const TraceArgs& a, cudaStream_t st) {
    static OccCache cache;
    int& blocks = cache.get();
    auto kernel = WIDE ? k_trace_cwbvh<MODE, COUNT> : k_trace_bvh<MODE, COUNT>;
    if (!blocks) {
        int dev = 0, sms = 0, perSm = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kernel, TRACE_THREADS, 0);
        blocks = sms * (perSm > 0 ? perSm : 1);
        if (!COUNT && MODE == MODE_CLOSEST) g_traceBlocks = blocks;
    }
    cudaMemsetAsync(a.workCounter, 0, sizeof(uint32_t), st);
    const uint32_t warpsNeeded = (a.n + 31u) / 32u, blocksNeeded = (warpsNeeded + TRACE_THREADS / 32 - 1) / (TRACE_THREADS / 32);
    kernel<<<min((uint32_t)blocks, blocksNeeded ? blocksNeeded : 1u), TRACE_THREADS, 0, st>>>(a);
}
template <int MODE>
static void launchTrace(const SceneView& sv, const TraceArgs& a, bool count, cudaStream_t st) {
    if (sv.useBvh == ACCEL_KIND_CWBVH) { if (count) launchTraceBvh<MODE, true, true>(a, st); else launchTraceBvh<MODE, false, true>(a, st); }
    else { if (count) launchTraceBvh<MODE, true, false>(a, st); else launchTraceBvh<MODE, false, false>(a, st); }
}
int trace_grid_blocks() { return g_traceBlocks; }

template <class K>
static uint32_t persistentBlocks(K kernel, int& cache, uint32_t n) {
    if (!cache) {
        int dev = 0, sms = 0, perSm = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kernel, TRACE_THREADS, 0);
        cache = sms * (perSm > 0 ? perSm : 1);
    }
    const uint32_t warpsNeeded = (n + 31u) / 32u, blocksNeeded = (warpsNeeded + TRACE_THREADS / 32 - 1) / (TRACE_THREADS / 32);
    return min((uint32_t)cache, blocksNeeded ? blocksNeeded : 1u);
}
template <bool COUNT, bool FRUSTUM>
static void launchPacket(const TraceArgs& a, cudaStream_t st) {
    static OccCache cache;
    int& blocks = cache.get();
    cudaMemsetAsync(a.workCounter, 0, sizeof(uint32_t), st);
    if (FRUSTUM) {
        auto kernel = k_trace_cwbvh_frustum<COUNT, false>;
        kernel<<<persistentBlocks(kernel, blocks, a.n), TRACE_THREADS, 0, st>>>(a, FusedArgs{});
    } else {
        auto kernel = k_trace_cwbvh_packet<COUNT>;
        kernel<<<persistentBlocks(kernel, blocks, a.n), TRACE_THREADS, 0, st>>>(a);
    }
}

// --------------------------------------------------------------------------------------------------------
// brute force: the reference's triangle loop (SH/trace.glsl:25-29, :78-81), triangles staged through shared memory
// --------------------------------------------------------------------------------------------------------
constexpr int BRUTE_THREADS = 256, BRUTE_CHUNK = 256;

template <int MODE>
__global__ void __launch_bounds__(BRUTE_THREADS) k_trace_brute(const RayRec* __restrict__ rays, uint32_t n,
                                                               const TriangleRec* __restrict__ tris, uint32_t triCount,
                                                               TriHit* __restrict__ hits, uint32_t* __restrict__ bits, const FrameMap fm,
                                                               uint8_t* __restrict__ bytes, const bool rejectZeroEdge,
                                                               const uint32_t* __restrict__ slotIds, const uint32_t* __restrict__ countPtr) {
    __shared__ float4 sTri[BRUTE_CHUNK * 3];
    const uint32_t slot = blockIdx.x * BRUTE_THREADS + threadIdx.x;
    const bool inRange = slot < (countPtr ? __ldg(countPtr) : n);
    float4 o = make_float4(0, 0, 0, 0), d = make_float4(0, 0, 1, -1.0f);
    if (inRange) { o = __ldg(reinterpret_cast<const float4*>(rays + slot)); d = __ldg(reinterpret_cast<const float4*>(rays + slot) + 1); }
    const bool live = inRange && d.w >= 0.0f;
    const vec3 ro = mk3(o.x, o.y, o.z), rd = mk3(d.x, d.y, d.z);
    const uint32_t prev = fbits(o.w);
    float hitT = live ? d.w : 0.0f, bu = 0.0f, bv = 0.0f;
    uint32_t object = NO_RAY_HIT;
    for (uint32_t base = 0; base < triCount; base += BRUTE_CHUNK) {
        const uint32_t cnt = min((uint32_t)BRUTE_CHUNK, triCount - base);
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < cnt * 3; i += BRUTE_THREADS) sTri[i] = __ldg(reinterpret_cast<const float4*>(tris + base) + i);
        __syncthreads();
        if (!live) continue;
        for (uint32_t k = 0; k < cnt; ++k) {
            const float4 a0 = sTri[k * 3], a1 = sTri[k * 3 + 1], a2 = sTri[k * 3 + 2];
            const vec3 p0 = mk3(a0.x, a0.y, a0.z), e1 = mk3(a1.x, a1.y, a1.z) - p0;
            if (rejectZeroEdge && e1.x == 0.0f && e1.y == 0.0f && e1.z == 0.0f) continue;   // DEBUG shader build: SH/primitive.glsl:248-253
            float u, v, t, aa;
            if (!triCandidate(ro, rd, p0, e1, mk3(a2.x, a2.y, a2.z) - p0, u, v, t, aa)) continue;
            const uint32_t id = base + k;
            if (t <= 0.0f || id == prev || t >= hitT) continue;   // SH/primitive.glsl:268 (a NaN t passes, as in the reference)
            hitT = t; object = id; bu = u; bv = v;
        }
    }
    if (!inRange) return;
    if (MODE == MODE_CLOSEST) {
        TriHit h; h.t = object == NO_RAY_HIT ? NO_HIT : hitT; h.id = object; h.u = bu; h.v = bv;
        *reinterpret_cast<float4*>(hits + slot) = *reinterpret_cast<float4*>(&h);
    } else if (MODE == MODE_ANY_BITS) {
        if (live && object != NO_RAY_HIT) {
            const uint32_t j = slotIds ? __ldg(slotIds + slot) : slot;
            const uint32_t sample = j / fm.localSlots, i = j - sample * fm.localSlots;
            uint32_t x, y;
            slotToPixel(fm, i, x, y);
            atomicOr(bits + indexToLight(x, y, fm.w, fm.h, sample), 1u << ((x & 15u) | ((y & 1u) << 4)));
        }
    } else {
        bytes[slot] = (live && object != NO_RAY_HIT) ? 1 : 0;
    }
}

__global__ void k_fill_miss(TriHit* hits, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    TriHit h; h.t = NO_HIT; h.id = NO_RAY_HIT; h.u = 0.0f; h.v = 0.0f;
    *reinterpret_cast<float4*>(hits + i) = *reinterpret_cast<float4*>(&h);
}

static TraceArgs makeArgs(const SceneView& sv, const RayRec* rays, uint32_t n, uint32_t* workCounter, TraceCounters* counters) {
    TraceArgs a{};
    a.rays = rays; a.n = n;
    a.nodes = reinterpret_cast<const float4*>(sv.nodes); a.nodes8 = reinterpret_cast<const uint4*>(sv.nodes8); a.nodes8Alias = a.nodes8;
    a.tris = reinterpret_cast<const float4*>(sv.travTris); a.nodeCount = sv.nodeCount;
    a.workCounter = workCounter; a.counters = counters;
    return a;
}

// camera rays generated, traced and finished into the G-buffer by one launch (frustum packets; see rtb_trace8f.cuh, FusedArgs)
void launch_primary_fused(const FrameMap& fm, const SceneView& sv, const CameraRec* cam, const SeedRec* seed, float4* dirT, float4* uvN,
                          uint32_t* workCounter, cudaStream_t st) {
    if (!fm.localSlots) return;
    static OccCache cache;
    int& blocks = cache.get();
    TraceArgs a = makeArgs(sv, nullptr, fm.localSlots, workCounter, nullptr);
    FusedArgs f{fm, *cam, seed, sv, dirT, uvN};
    cudaMemsetAsync(workCounter, 0, sizeof(uint32_t), st);
    auto kernel = k_trace_cwbvh_frustum<false, true>;
    kernel<<<persistentBlocks(kernel, blocks, a.n), TRACE_THREADS, 0, st>>>(a, f);
}

void launch_trace_closest(const SceneView& sv, const RayRec* rays, uint32_t n, TriHit* hits, uint32_t* workCounter,
                          TraceCounters* counters, int packets, cudaStream_t st, const uint32_t* countPtr) {
    if (!n) return;
    if (sv.info.triangleCount == 0) { k_fill_miss<<<(n + 255) / 256, 256, 0, st>>>(hits, n); return; }
    if (sv.useBvh == ACCEL_KIND_CWBVH && packets && !countPtr) {
        TraceArgs a = makeArgs(sv, rays, n, workCounter, counters);
        a.hits = hits;
        if (packets == PACKETS_FRUSTUM) { if (counters) launchPacket<true, true>(a, st); else launchPacket<false, true>(a, st); }
        else { if (counters) launchPacket<true, false>(a, st); else launchPacket<false, false>(a, st); }
    } else if (sv.useBvh) {
        TraceArgs a = makeArgs(sv, rays, n, workCounter, counters);
        a.hits = hits; a.countPtr = countPtr;
        launchTrace<MODE_CLOSEST>(sv, a, counters != nullptr, st);
    } else {
        FrameMap fm{};
        k_trace_brute<MODE_CLOSEST><<<(n + BRUTE_THREADS - 1) / BRUTE_THREADS, BRUTE_THREADS, 0, st>>>(rays, n, sv.triangles, sv.info.triangleCount, hits, nullptr, fm, nullptr, !sv.releaseBuild, nullptr, countPtr);
    }
}

void launch_trace_any_bits(const FrameMap& fm, const SceneView& sv, const RayRec* rays, uint32_t n, uint32_t* bits,
                           uint32_t* workCounter, TraceCounters* counters, const uint32_t* slotIds, const uint32_t* countPtr, cudaStream_t st) {
    if (!n || sv.info.triangleCount == 0) return;
    if (sv.useBvh) {
        TraceArgs a = makeArgs(sv, rays, n, workCounter, counters);
        a.bits = bits; a.fm = fm; a.slotIds = slotIds; a.countPtr = countPtr;
        launchTrace<MODE_ANY_BITS>(sv, a, counters != nullptr, st);
    } else {
        k_trace_brute<MODE_ANY_BITS><<<(n + BRUTE_THREADS - 1) / BRUTE_THREADS, BRUTE_THREADS, 0, st>>>(rays, n, sv.triangles, sv.info.triangleCount, nullptr, bits, fm, nullptr, !sv.releaseBuild, slotIds, countPtr);
    }
}

// occlusion by beam packets over a light-space sorted queue (rtb_trace8b.cuh); what the beam walk does not take goes through
// `fallback` to the per-ray kernel
void launch_trace_beam_bits(const FrameMap& fm, const SceneView& sv, const RayRec* rays, uint32_t n, uint32_t* bits, uint32_t* workCounter,
                            TraceCounters* counters, const uint32_t* slotIds, const uint32_t* countPtr, const RayQueue& fallback, cudaStream_t st) {
    if (!n || sv.info.triangleCount == 0) return;
    static OccCache cache, cacheCount;
    int &blocks = cache.get(), &blocksCount = cacheCount.get();
    TraceArgs a = makeArgs(sv, rays, n, workCounter, counters);
    a.bits = bits; a.fm = fm; a.slotIds = slotIds; a.countPtr = countPtr;
    cudaMemsetAsync(workCounter, 0, sizeof(uint32_t), st);
    cudaMemsetAsync(fallback.count, 0, sizeof(uint32_t), st);
    if (counters) { auto kernel = k_trace_cwbvh_beam<true>; kernel<<<persistentBlocks(kernel, blocksCount, n), TRACE_THREADS, 0, st>>>(a, fallback); }
    else { auto kernel = k_trace_cwbvh_beam<false>; kernel<<<persistentBlocks(kernel, blocks, n), TRACE_THREADS, 0, st>>>(a, fallback); }
    launch_trace_any_bits(fm, sv, fallback.rays, n, bits, workCounter, counters, fallback.slotIds, fallback.count, st);
}

void launch_trace_any_bytes(const SceneView& sv, const RayRec* rays, uint32_t n, uint8_t* occluded, uint32_t* workCounter, cudaStream_t st,
                            TraceCounters* counters, const uint32_t* countPtr) {
    if (!n) return;
    if (sv.info.triangleCount == 0) { cudaMemsetAsync(occluded, 0, n, st); return; }
    if (sv.useBvh) {
        TraceArgs a = makeArgs(sv, rays, n, workCounter, counters);
        a.bytes = occluded; a.countPtr = countPtr;
        launchTrace<MODE_ANY_BYTES>(sv, a, counters != nullptr, st);
    } else {
        FrameMap fm{};
        k_trace_brute<MODE_ANY_BYTES><<<(n + BRUTE_THREADS - 1) / BRUTE_THREADS, BRUTE_THREADS, 0, st>>>(rays, n, sv.triangles, sv.info.triangleCount, nullptr, nullptr, fm, occluded, !sv.releaseBuild, nullptr, countPtr);
    }
}

// --------------------------------------------------------------------------------------------------------
// K1b: the rest of traceGeometry after the triangle loop, then the G-buffer stores
// --------------------------------------------------------------------------------------------------------
// SH/trace.glsl:31-63: spheres, cubes, planes in order with running ids; then the normal selection.
// sph / cub: the winner of the type's tree traversal (rtb_trace8s.cuh) when that type has a tree — the reference's function then
// runs on that one primitive (it is accepted: its distance undercuts what the earlier stages left) — or nullptr: the linear loop.
RTB_DI void finishGeometry(const SceneView& sv, const Ray& ray, uint32_t prev, const TriHit& th, Hit& hit, vec3& objectNormal,
                           const PrimHit* sph = nullptr, const PrimHit* cub = nullptr) {
    hit.hitT = th.id == NO_RAY_HIT ? NO_HIT : th.t;
    hit.uv = mk2(th.u, th.v);
    hit.object = th.id == NO_RAY_HIT ? 0u : th.id;
    hit.geometryNormal = mk3(0.0f, 0.0f, 0.0f);
    uint32_t j = sv.info.triangleCount;
    if (sph) {
        if (sph->id != NO_RAY_HIT && rayIntersectSphere(ray, __ldg(sv.spheres + sph->id), hit, j + sph->id, prev)) hit.object = j + sph->id;
        j += sv.info.sphereCount;
    } else
    for (uint32_t i = 0; i < sv.info.sphereCount; ++i, ++j)
        if (rayIntersectSphere(ray, __ldg(sv.spheres + i), hit, j, prev)) hit.object = j;
    if (cub) {
        if (cub->id != NO_RAY_HIT) {
            float c[6];
            const float2* cp = reinterpret_cast<const float2*>(sv.cubes + 6 * (size_t)cub->id);
            const float2 c0 = __ldg(cp), c1 = __ldg(cp + 1), c2 = __ldg(cp + 2);
            c[0] = c0.x; c[1] = c0.y; c[2] = c1.x; c[3] = c1.y; c[4] = c2.x; c[5] = c2.y;
            if (rayIntersectCube(ray, c, hit, j + cub->id, prev)) hit.object = j + cub->id;
        }
        j += sv.info.cubeCount;
    } else
    for (uint32_t i = 0; i < sv.info.cubeCount; ++i, ++j) {
        float c[6];
        const float2* cp = reinterpret_cast<const float2*>(sv.cubes + 6 * (size_t)i);
        const float2 c0 = __ldg(cp), c1 = __ldg(cp + 1), c2 = __ldg(cp + 2);
        c[0] = c0.x; c[1] = c0.y; c[2] = c1.x; c[3] = c1.y; c[4] = c2.x; c[5] = c2.y;
        if (rayIntersectCube(ray, c, hit, j, prev)) hit.object = j;
    }
    for (uint32_t i = 0; i < sv.info.planeCount; ++i, ++j)
        if (rayIntersectPlane(ray, __ldg(sv.planes + i), hit, j, prev)) hit.object = j;
    if (hit.object < sv.info.triangleCount) {   // also taken on a miss when the scene has triangles (object == 0)
        const TriangleRec* t = sv.triangles + hit.object;
        const uint32_t e0 = __ldg(&t->n0), e1 = __ldg(&t->n1), e2 = __ldg(&t->n2);
        objectNormal = interpolate(decodeSpheremap(e0), decodeSpheremap(e1), decodeSpheremap(e2), hit.uv);
    } else
        objectNormal = hit.geometryNormal;
}

__global__ void __launch_bounds__(256) k_finish_primary(const FrameMap fm, const SceneView sv, const RayRec* __restrict__ rays,
                                                        const TriHit* __restrict__ hits, float4* __restrict__ dirT, float4* __restrict__ uvN,
                                                        const PrimHit* __restrict__ sph, const PrimHit* __restrict__ cub) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= fm.localSlots) return;
    uint32_t x, y;
    if (!slotToPixel(fm, i, x, y)) return;
    const float4 o = __ldg(reinterpret_cast<const float4*>(rays + i)), d = __ldg(reinterpret_cast<const float4*>(rays + i) + 1);
    const float4 hv = __ldg(reinterpret_cast<const float4*>(hits + i));
    TriHit th; th.t = hv.x; th.id = fbits(hv.y); th.u = hv.z; th.v = hv.w;
    Ray ray; ray.pos = mk3(o.x, o.y, o.z); ray.dir = mk3(d.x, d.y, d.z);
    Hit hit; vec3 objectNormal;
    PrimHit ws, wc;
    if (sph) ws = sph[i];
    if (cub) wc = cub[i];
    finishGeometry(sv, ray, NO_RAY_HIT, th, hit, objectNormal, sph ? &ws : nullptr, cub ? &wc : nullptr);
    // SH/raygen.comp:39-51
    float4 out0;
    if (hit.hitT == NO_HIT) out0 = make_float4(ray.dir.x, ray.dir.y, ray.dir.z, ubits(NO_RAY_HIT));
    else { const vec3 v = ray.dir * hit.hitT; out0 = make_float4(v.x, v.y, v.z, ubits(hit.object)); }
    uint32_t ex, ey;
    encodeNormalGpu(objectNormal, ex, ey);
    const size_t px = (size_t)y * fm.w + x;
    dirT[px] = out0;
    if (!sv.releaseBuild || hit.hitT < NO_HIT)   // the RELEASE build stores uvObjectNormal for hits only (SH/raygen.comp:46-51)
        uvN[px] = make_float4(hit.uv.x, hit.uv.y, ubits(ex), ubits(ey));
}
void launch_finish_primary(const FrameMap& fm, const SceneView& sv, const RayRec* rays, const TriHit* hits, float4* dirT, float4* uvN, cudaStream_t st,
                           const PrimHit* sph, const PrimHit* cub) {
    if (!fm.localSlots) return;
    k_finish_primary<<<(fm.localSlots + 255) / 256, 256, 0, st>>>(fm, sv, rays, hits, dirT, uvN, sph, cub);
}

// instrumented frames only: pixels whose nearest hit is ANY primitive (= the shadow rays the reference traces per sample)
__global__ void __launch_bounds__(256) k_count_hits(const FrameMap fm, const float4* __restrict__ dirT, TraceCounters* __restrict__ counters) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t x, y;
    const bool hit = i < fm.localSlots && slotToPixel(fm, i, x, y) && fbits(dirT[(size_t)y * fm.w + x].w) != NO_RAY_HIT;
    const unsigned m = __ballot_sync(0xFFFFFFFFu, hit);
    if ((threadIdx.x & 31u) == 0 && m) atomicAdd(&counters->hits, (unsigned long long)__popc(m));
}
void launch_count_hits(const FrameMap& fm, const float4* dirT, TraceCounters* counters, cudaStream_t st) {
    if (!fm.localSlots) return;
    cudaMemsetAsync(&counters->hits, 0, sizeof(unsigned long long), st);
    k_count_hits<<<(fm.localSlots + 255) / 256, 256, 0, st>>>(fm, dirT, counters);
}

__global__ void __launch_bounds__(256) k_finish_rays(const SceneView sv, const RayRec* __restrict__ rays, const TriHit* __restrict__ hits,
                                                     uint32_t n, uint32_t* __restrict__ object, float* __restrict__ t, float2* __restrict__ uv,
                                                     const PrimHit* __restrict__ sph, const PrimHit* __restrict__ cub) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 o = __ldg(reinterpret_cast<const float4*>(rays + i)), d = __ldg(reinterpret_cast<const float4*>(rays + i) + 1);
    const float4 hv = __ldg(reinterpret_cast<const float4*>(hits + i));
    TriHit th; th.t = hv.x; th.id = fbits(hv.y); th.u = hv.z; th.v = hv.w;
    Ray ray; ray.pos = mk3(o.x, o.y, o.z); ray.dir = mk3(d.x, d.y, d.z);
    Hit hit; vec3 objectNormal;
    PrimHit ws, wc;
    if (sph) ws = sph[i];
    if (cub) wc = cub[i];
    finishGeometry(sv, ray, fbits(o.w), th, hit, objectNormal, sph ? &ws : nullptr, cub ? &wc : nullptr);
    if (object) object[i] = hit.hitT == NO_HIT ? NO_RAY_HIT : hit.object;
    if (t) t[i] = hit.hitT;
    if (uv) uv[i] = make_float2(hit.uv.x, hit.uv.y);
}
void launch_finish_rays(const SceneView& sv, const RayRec* rays, const TriHit* hits, uint32_t n, uint32_t* object, float* t, float2* uv, cudaStream_t st,
                        const PrimHit* sph, const PrimHit* cub) {
    if (!n) return;
    k_finish_rays<<<(n + 255) / 256, 256, 0, st>>>(sv, rays, hits, n, object, t, uv, sph, cub);
}

static PrimTraceArgs primArgs(int kind, const SceneView& sv, const PrimTree& tree, const RayRec* rays, uint32_t n, const uint32_t* countPtr) {
    PrimTraceArgs a{};
    a.rays = rays; a.n = n; a.countPtr = countPtr;
    a.nodes8 = reinterpret_cast<const uint4*>(tree.nodes); a.nodes8Alias = a.nodes8; a.tt = reinterpret_cast<const float4*>(tree.tt);
    a.spheres = sv.spheres; a.cubes = sv.cubes;
    a.firstObject = sv.info.triangleCount + (kind == 1 ? sv.info.sphereCount : 0u);
    return a;
}
void launch_prims_closest(int kind, const SceneView& sv, const PrimTree& tree, const RayRec* rays, uint32_t n, const uint32_t* countPtr,
                          const TriHit* triHits, const PrimHit* before, PrimHit* out, cudaStream_t st) {
    PrimTraceArgs a = primArgs(kind, sv, tree, rays, n, countPtr);
    a.triHits = triHits; a.before = before; a.out = out;
    launch_trace_prims(kind, false, a, st);
}
void launch_prims_any(int kind, const SceneView& sv, const PrimTree& tree, const RayRec* rays, uint32_t n, const uint32_t* countPtr, uint8_t* bytes,
                      uint32_t* bits, const uint32_t* slotIds, const FrameMap& fm, cudaStream_t st) {
    PrimTraceArgs a = primArgs(kind, sv, tree, rays, n, countPtr);
    a.bytes = bytes; a.bits = bits; a.slotIds = slotIds; a.fm = fm;
    launch_trace_prims(kind, true, a, st);
}

// --------------------------------------------------------------------------------------------------------
// K2: shadow rays
// --------------------------------------------------------------------------------------------------------
// Occlusion by spheres, cubes and planes: the tail of traceOcclusion (SH/trace.glsl:83-96).  The loops only ever lower
// hitT, so "hitT < maxDist after all loops" is the same as "some primitive's candidate distance is < maxDist", which
// lets the triangles be searched separately (any-hit) without changing the result.
RTB_DI bool occludedByOthers(const SceneView& sv, const Ray& ray, float maxDist, uint32_t prev) {
    Hit hit;
    hit.hitT = NO_HIT; hit.uv = mk2(0.0f, 0.0f); hit.object = 0; hit.geometryNormal = mk3(0.0f, 0.0f, 0.0f);
    uint32_t j = sv.info.triangleCount;
    if (sv.sphereTree) j += sv.info.sphereCount;   // (answered by the spheres' tree traversal)
    else
    for (uint32_t i = 0; i < sv.info.sphereCount; ++i, ++j) {
        if (j == prev) continue;
        const float t = sphereCandidateT(ray, __ldg(sv.spheres + i));
        if (t < hit.hitT) hit.hitT = t;
    }
    if (sv.cubeTree) j += sv.info.cubeCount;
    else
    for (uint32_t i = 0; i < sv.info.cubeCount; ++i, ++j) {
        float c[6];
        const float2* cp = reinterpret_cast<const float2*>(sv.cubes + 6 * (size_t)i);
        const float2 c0 = __ldg(cp), c1 = __ldg(cp + 1), c2 = __ldg(cp + 2);
        c[0] = c0.x; c[1] = c0.y; c[2] = c1.x; c[3] = c1.y; c[4] = c2.x; c[5] = c2.y;
        rayIntersectCube(ray, c, hit, j, prev);
    }
    for (uint32_t i = 0; i < sv.info.planeCount; ++i, ++j) rayIntersectPlane(ray, __ldg(sv.planes + i), hit, j, prev);
    return hit.hitT < maxDist;
}

// Hit and miss pixels do very different amounts of work (a shadow ray and Cook-Torrance against one sky lookup), and in
// a scene like the triangle soup they alternate inside every 8x4 patch: half the lanes of every warp idle through the
// other half's branch (ncu r1r: 16.8 of 32 lanes active in k_shade and k_shadowgen).  The 256 slots of a block are
// therefore dealt to its threads hits first: all-hit warps, one mixed warp, all-miss warps.  Every slot is still
// processed exactly once by the same code, so nothing changes but which thread does it.  All 256 threads must call.
RTB_DI uint32_t hitsFirst(bool isHit) {
    __shared__ uint16_t sOrder[256];
    __shared__ uint32_t sWarpHits[8];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, isHit);
    if (lane == 0) sWarpHits[warp] = (uint32_t)__popc(m);
    __syncthreads();
    uint32_t before = 0, total = 0;
    #pragma unroll
    for (uint32_t w = 0; w < 8u; ++w) { const uint32_t c = sWarpHits[w]; total += c; if (w < warp) before += c; }
    const uint32_t hitsBeforeMe = before + (uint32_t)__popc(m & ((1u << lane) - 1u));
    sOrder[isHit ? hitsBeforeMe : total + (tid - hitsBeforeMe)] = (uint16_t)tid;
    __syncthreads();
    return sOrder[tid];
}

// Both shading kernels are chains of dependent binary64 operations: latency-bound, so occupancy pays even at the price
// of spills (B200, 4K soup frame: k_shadowgen 0.231 ms at 4 blocks/SM -> 0.202 at 8; k_shade 0.535 ms at 3 -> 0.411 at 6).
#ifndef RTB_SHADOWGEN_MINBLOCKS
#define RTB_SHADOWGEN_MINBLOCKS 8
#endif
__global__ void __launch_bounds__(256, RTB_SHADOWGEN_MINBLOCKS) k_shadowgen(const FrameMap fm, const SceneView sv, const CameraRec cam, const SeedRec* __restrict__ seed,
                                                   uint32_t samples, const float4* __restrict__ dirT, RayRec* __restrict__ rays,
                                                   uint32_t* __restrict__ bits, const RayQueue q, const RayBin bin) {
    // localSlots is a multiple of 1024: a block of 256 slots lies in one sample and is never ragged
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    {
        const uint32_t sample0 = j / fm.localSlots, i0 = j - sample0 * fm.localSlots;
        uint32_t x0, y0;
        const bool hit0 = j < fm.localSlots * samples && slotToPixel(fm, i0, x0, y0) && fbits(__ldg(dirT + (size_t)y0 * fm.w + x0).w) != NO_RAY_HIT;
        j = blockIdx.x * blockDim.x + hitsFirst(hit0);
    }
    const bool inRange = j < fm.localSlots * samples;   // always true: the grid covers localSlots * samples exactly
    const uint32_t sample = j / fm.localSlots, i = j - sample * fm.localSlots;
    float4 ro = make_float4(0.f, 0.f, 0.f, ubits(NO_RAY_HIT)), rd = make_float4(0.f, 0.f, 1.f, -1.0f);
    uint32_t x, y;
    if (inRange && slotToPixel(fm, i, x, y)) {
        const float4 dt = __ldg(dirT + (size_t)y * fm.w + x);
        const uint32_t object = fbits(dt.w);
        if (object != NO_RAY_HIT) {
            // SH/nv_all.shadow.comp:84-126
            const vec3 hitPos = mk3(cam.eye) + mk3(dt.x, dt.y, dt.z);
            const vec2 loc = mk2((float)x, (float)y);
            vec2 uv = (loc + rand2(loc + mk2(__ldg(&seed->randomX), __ldg(&seed->randomY)))) / 128.0f;
            uv = uv + hammersley(sample, samples);
            const vec2 random = rand2(uv);
            const LightRec light = sv.lights[0];   // lightId = 0 (SH/nv_all.shadow.comp:97)
            float brightness, dist;
            const vec3 l = getDirToLight(light, hitPos, brightness, dist, random, sv.sun0);
            Ray ray; ray.pos = hitPos; ray.dir = -l;
            float maxDist = -1.0f;
            if (dist >= 0.0f) {
                const vec2 radOrigin = unpackHalf2x16(light.radOrigin);
                if (dist >= radOrigin.y && dist < radOrigin.x) maxDist = dist - radOrigin.y;
            } else
                maxDist = NO_HIT;
            // maxDist <= 0 can never be undercut by a triangle (t > 0) but a cube entered from inside reports t < 0
            if (maxDist != -1.0f) {
                if (occludedByOthers(sv, ray, maxDist, object))
                    atomicOr(bits + indexToLight(x, y, fm.w, fm.h, sample), 1u << ((x & 15u) | ((y & 1u) << 4)));
                else if (maxDist > 0.0f || sv.sphereTree || sv.cubeTree) {   // (a cube entered from inside occludes at a distance below zero)
                    ro = make_float4(ray.pos.x, ray.pos.y, ray.pos.z, ubits(object));
                    rd = make_float4(ray.dir.x, ray.dir.y, ray.dir.z, maxDist);
                }
            }
        }
    }
    if (!q.count) {   // slot order: one record per slot, dead ones marked
        if (inRange) { float4* out = reinterpret_cast<float4*>(rays + j); out[0] = ro; out[1] = rd; }
        return;
    }
    // queue: live rays only, one atomic per warp
    const bool live = rd.w >= 0.0f;
    const unsigned m = __ballot_sync(0xFFFFFFFFu, live);
    if (!m) return;
    const unsigned lane = threadIdx.x & 31u;
    uint32_t base = 0;
    if (lane == (unsigned)(__ffs(m) - 1)) base = atomicAdd(q.count, (uint32_t)__popc(m));
    base = __shfl_sync(0xFFFFFFFFu, base, __ffs(m) - 1);
    if (!live) return;
    const uint32_t r = base + (uint32_t)__popc(m & ((1u << lane) - 1u));
    float4* out = reinterpret_cast<float4*>(q.rays + r);
    out[0] = ro; out[1] = rd;
    q.slotIds[r] = j;
    if (bin.kind) {
        float u, v;
        if (bin.kind == 1u) {
            u = ro.x * bin.b1[0] + ro.y * bin.b1[1] + ro.z * bin.b1[2];
            v = ro.x * bin.b2[0] + ro.y * bin.b2[1] + ro.z * bin.b2[2];
        } else {   // octahedral map of the direction from the light to the ray origin
            const float dx = ro.x - bin.lpos[0], dy = ro.y - bin.lpos[1], dz = ro.z - bin.lpos[2];
            const float inv = 1.0f / fmaxf(fabsf(dx) + fabsf(dy) + fabsf(dz), 1e-30f);
            u = dx * inv; v = dy * inv;
            if (dz < 0.0f) { const float tu = (1.0f - fabsf(v)) * (u >= 0.0f ? 1.0f : -1.0f); v = (1.0f - fabsf(u)) * (v >= 0.0f ? 1.0f : -1.0f); u = tu; }
        }
        const float cmax = (float)((1u << bin.bits) - 1u);
        const uint32_t cu = (uint32_t)fminf(fmaxf((u - bin.u0) * bin.su, 0.0f), cmax), cv = (uint32_t)fminf(fmaxf((v - bin.v0) * bin.sv, 0.0f), cmax);
        auto spread = [](uint32_t a) { a = (a | (a << 8)) & 0x00FF00FFu; a = (a | (a << 4)) & 0x0F0F0F0Fu; a = (a | (a << 2)) & 0x33333333u; a = (a | (a << 1)) & 0x55555555u; return a; };
        const uint32_t cell = spread(cu) | (spread(cv) << 1);
        q.cell[r] = cell;
        q.rank[r] = atomicAdd(q.hist + cell, 1u);
    }
}
void launch_shadowgen(const FrameMap& fm, const SceneView& sv, const CameraRec* cam, const SeedRec* seed, uint32_t samples,
                      const float4* dirT, RayRec* rays, uint32_t* bits, const RayQueue* queue, const RayBin* bin, cudaStream_t st) {
    const uint32_t n = fm.localSlots * samples;
    if (!n) return;
    k_shadowgen<<<(n + 255) / 256, 256, 0, st>>>(fm, sv, *cam, seed, samples, dirT, rays, bits, queue ? *queue : RayQueue{}, bin ? *bin : RayBin{});
}

// RELEASE shader build: one thread per shadow word (a 16x2 pixel strip of one sample)
__global__ void __launch_bounds__(256) k_clear_hit_strips(const FrameMap fm, const float4* __restrict__ dirT, uint32_t samples, uint32_t* __restrict__ bits) {
    const uint32_t tilesX = (fm.w + 15u) >> 4, tilesY = (fm.h + 1u) >> 1, perSample = tilesX * tilesY;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= perSample) return;
    const uint32_t tx = i % tilesX, ty = i / tilesX;
    bool any = false;
    for (uint32_t l = 0; l < 32u && !any; ++l) {
        const uint32_t x = tx * 16u + (l & 15u), y = ty * 2u + (l >> 4);
        if (x < fm.w && y < fm.h && fbits(__ldg(dirT + (size_t)y * fm.w + x).w) != NO_RAY_HIT) any = true;
    }
    if (any) for (uint32_t s = 0; s < samples; ++s) bits[i + s * perSample] = 0u;
}
void launch_clear_hit_strips(const FrameMap& fm, const float4* dirT, uint32_t samples, uint32_t* bits, cudaStream_t st) {
    const uint32_t n = ((fm.w + 15u) >> 4) * ((fm.h + 1u) >> 1);
    if (!n) return;
    k_clear_hit_strips<<<(n + 255) / 256, 256, 0, st>>>(fm, dirT, samples, bits);
}

__global__ void __launch_bounds__(256) k_occlusion_others(const SceneView sv, RayRec* __restrict__ rays, uint32_t n, uint8_t* __restrict__ occluded,
                                                          const uint32_t* __restrict__ countPtr) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (countPtr ? __ldg(countPtr) : n)) return;
    const float4 o = *reinterpret_cast<const float4*>(rays + i), d = *(reinterpret_cast<const float4*>(rays + i) + 1);
    Ray ray; ray.pos = mk3(o.x, o.y, o.z); ray.dir = mk3(d.x, d.y, d.z);
    const bool occ = occludedByOthers(sv, ray, d.w, fbits(o.w));
    occluded[i] = occ ? 1 : 0;
    if (occ || (!(d.w > 0.0f) && !sv.sphereTree && !sv.cubeTree)) rays[i].tmax = -1.0f;   // nothing left for the searches that follow
}
void launch_occlusion_others(const SceneView& sv, RayRec* rays, uint32_t n, uint8_t* occluded, cudaStream_t st, const uint32_t* countPtr) {
    if (!n) return;
    k_occlusion_others<<<(n + 255) / 256, 256, 0, st>>>(sv, rays, n, occluded, countPtr);
}

// --------------------------------------------------------------------------------------------------------
// K3 + K4: lighting and composite, one thread per pixel
// --------------------------------------------------------------------------------------------------------
// DO_LIGHT: evaluate lighting.comp; DO_COMP: evaluate composite.comp.  Both: the fused frame path (the rgba16f
// rounding of the lighting texture is reproduced in registers); one only: the reference's separate dispatches.
#ifndef RTB_SHADE_MINBLOCKS
#define RTB_SHADE_MINBLOCKS 6
#endif
template <bool DO_LIGHT, bool DO_COMP, bool EXT>
__global__ void __launch_bounds__(256, RTB_SHADE_MINBLOCKS) k_shade(const FrameMap fm, const SceneView sv, const CameraRec cam, const SeedRec* __restrict__ seed,
                                               uint32_t samples, const float4* __restrict__ dirT, const float4* __restrict__ uvN,
                                               const uint32_t* __restrict__ bits, uint2* __restrict__ lighting, float4* __restrict__ accum,
                                               uint32_t* __restrict__ rgba8, uint32_t* __restrict__ rgba8Tiled, const LightsView lv) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    {
        uint32_t x0, y0;
        const bool hit0 = i < fm.localSlots && slotToPixel(fm, i, x0, y0) && fbits(__ldg(dirT + (size_t)y0 * fm.w + x0).w) != NO_RAY_HIT;
        i = blockIdx.x * blockDim.x + hitsFirst(hit0);
    }
    if (i >= fm.localSlots) return;
    uint32_t x, y;
    if (!slotToPixel(fm, i, x, y)) { if (DO_COMP && rgba8Tiled) rgba8Tiled[tiledSlot(fm, i)] = 0u; return; }
    const size_t px = (size_t)y * fm.w + x;
    const float4 dt = __ldg(dirT + px), un = __ldg(uvN + px);
    const uint32_t object = fbits(dt.w);
    const vec3 dxyz = mk3(dt.x, dt.y, dt.z);
    const vec3 eye = mk3(cam.eye);
    const vec3 n = decodeNormal(fbits(un.z), fbits(un.w));
    const bool isHit = object != NO_RAY_HIT;

    MatU m;
    if (isHit) m = unpackMaterial(sv.materials + __ldg(sv.materialIndices + object));

    uint32_t lx, ly, lz;
    if (DO_LIGHT) {
        // ---- lighting.comp (SH/nv_all.lighting.comp:33-103) --------------------------------------------
        vec3 light = mk3(0.0f, 0.0f, 0.0f);
        if (isHit) {
            const vec3 hitPos = eye + dxyz;
            const vec3 F0 = mix(mk3(0.04f, 0.04f, 0.04f), m.albedo, m.metallic);
            const vec3 v = normalize(dxyz);
            const float NdotV = fmaxf(dot(v, -n), 0.0f);
            const vec2 loc = mk2((float)x, (float)y);
            vec2 uv = mk2(0.0f, 0.0f);
            if (!lv.cacheKind) uv = (loc + rand2(loc + mk2(0.0f, 0.0f))) / 128.0f;   // Seed is unbound in lighting.comp: reads as zero (decree D8)
            const size_t plane = (size_t)fm.w * fm.h;
            const uint32_t bit = 1u << ((x & 15u) | ((y & 1u) << 4));
            if (!EXT || lv.mode == 0u) {
                const LightRec l0 = sv.lights[0];
                for (uint32_t s = 0; s < samples; ++s) {
                    const uint32_t word = __ldg(bits + indexToLight(x, y, fm.w, fm.h, s));
                    if (word & bit) continue;
                    if (lv.cacheKind == 2u) {   // the direction to the directional light 0 as getDirToLight returns it, from the cache
                        const float4 cv = __ldg(lv.lightCache + s * plane + px);
                        light = light + shadeLightDir(F0, m.albedo, m.roughness, m.metallic, l0, mk3(cv.x, cv.y, cv.z), 1.0f, n, v, NdotV);
                    } else {
                        vec2 random;
                        if (lv.cacheKind == 1u) { const float4 cv = __ldg(lv.lightCache + s * plane + px); random = mk2(cv.x, cv.y); }
                        else random = rand2(uv + hammersley(s, samples));
                        light = light + shadeLight(F0, m.albedo, m.roughness, m.metallic, l0, hitPos, n, v, NdotV, random, sv.sun0);
                    }
                }
                light = light / (float)samples * (float)sv.info.lightCount;
            } else {
                // every light, in ascending order (the order of the sum is part of the result); the tile list leaves out only lights
                // whose term is exactly zero for every pixel of the tile
                uint32_t cnt = sv.info.lightCount;
                const uint32_t* list = nullptr;
                if (lv.mode == 2u) {
                    const uint32_t tile = (y >> 4) * lv.tilesX + (x >> 4), c = __ldg(lv.tileCount + tile);
                    if (c != LIGHT_TILE_ALL) { cnt = c; list = lv.tileList + (size_t)tile * LIGHTS_PER_TILE; }
                }
                for (uint32_t e = 0; e < cnt; ++e) {
                    const uint32_t L = list ? __ldg(list + e) : e;
                    const LightRec ll = sv.lights[L];
                    for (uint32_t s = 0; s < samples; ++s) {
                        vec2 random;
                        if (lv.cacheKind == 1u) { const float4 cv = __ldg(lv.lightCache + s * plane + px); random = mk2(cv.x, cv.y); }
                        else random = rand2(uv + hammersley(s, samples));
                        const uint32_t word = __ldg(bits + indexToLight(x, y, fm.w, fm.h, L * samples + s));
                        if (!(word & bit)) light = light + shadeLight(F0, m.albedo, m.roughness, m.metallic, ll, hitPos, n, v, NdotV, random, L == 0u ? sv.sun0 : nullptr);
                    }
                }
                light = light / (float)samples;
            }
        }
        // imageStore to rgba16f (alpha 1 on a hit; the shipped DEBUG shader stores vec4(0) on a miss)
        lx = f2h_rn(light.x); ly = f2h_rn(light.y); lz = f2h_rn(light.z);
        uint32_t la = isHit ? 0x3C00u : 0u;
        if (EXT && lv.historyAlpha > 0.0f) {   // temporal blend through the History texture (rgba16f), evaluated in binary32, stored in both
            const uint2 hv = lv.history[px];
            const float a = lv.historyAlpha, b = 1.0f - a;
            lx = f2h_rn(h2f(hv.x & 0xFFFFu) * b + h2f(lx) * a); ly = f2h_rn(h2f(hv.x >> 16) * b + h2f(ly) * a);
            lz = f2h_rn(h2f(hv.y & 0xFFFFu) * b + h2f(lz) * a); la = f2h_rn(h2f(hv.y >> 16) * b + h2f(la) * a);
            lv.history[px] = make_uint2(lx | (ly << 16), lz | (la << 16));
            if (lighting) lighting[px] = make_uint2(lx | (ly << 16), lz | (la << 16));
        } else
        if (lighting && (isHit || !sv.releaseBuild)) lighting[px] = make_uint2(lx | (ly << 16), lz | (la << 16));
        if (!DO_COMP) return;
    } else {
        const uint2 lv = __ldg(lighting + px);
        lx = lv.x & 0xFFFFu; ly = lv.x >> 16; lz = lv.y & 0xFFFFu;
    }
    const vec3 light = mk3(h2f(lx), h2f(ly), h2f(lz));

    // ---- composite.comp (SH/composite.comp:52-285, SH/light.glsl:205-219) ---------------------------------
    const SkyView sky = {sv.skybox, sv.skyW, sv.skyH};
    const vec3 rayDir = normalize(dxyz);
    vec3 color;
    if (!isHit)
        color = sampleSkybox(sky, cam, rayDir);
    else {
        const float NdotV = fmaxf(dot(rayDir, -n), 0.0f);
        const vec3 reflected = sampleSkybox(sky, cam, reflect(rayDir, n));
        color = shade(m, NdotV, light, reflected);
    }
    color = mix(color, mk3(0.0f, 0.0f, 0.0f), 0.0f);   // cloud term is vec4(0) (SH/composite.comp:93-97)
    if (!sv.releaseBuild && (isnan(color.x) || isnan(color.y) || isnan(color.z)))
        color = mk3(0.0f, 0.0f, 10000.0f);             // DEBUG build: NaN shown as bright blue (SH/composite.comp:236-239)
    if (cam.flags & CAMERA_USE_SUPERSAMPLING) {          // SH/composite.comp:249-257
        const uint32_t sampleCount = __ldg(&seed->sampleCount);
        if (sampleCount > 1) { const float4 p = accum[px]; color = color + mk3(p.x, p.y, p.z); }
        accum[px] = make_float4(color.x, color.y, color.z, 0.0f);
        color = color / (float)sampleCount;
    }
    const vec3 e = -color * cam.exposure;
    color = vmax(mk3(1.0f, 1.0f, 1.0f) - mk3(cr_exp(e.x), cr_exp(e.y), cr_exp(e.z)), mk3(0.0f, 0.0f, 0.0f));
    const uint32_t out = unorm8(color.x) | (unorm8(color.y) << 8) | (unorm8(color.z) << 16) | (255u << 24);
    if (rgba8) rgba8[px] = out;
    if (rgba8Tiled) rgba8Tiled[tiledSlot(fm, i)] = out;
}
// SunFrame of lights[0]: the light-only subexpressions of getDirToLight / getSunDirection, evaluated by the same functions
__global__ void k_sun_frame(const LightRec* __restrict__ lights, SunFrame* __restrict__ out) {
    const LightRec light = lights[0];
    vec2 radOrigin = unpackHalf2x16(light.radOrigin);
    radOrigin = mk2(fmaxf(radOrigin.x, 0.0f), fmaxf(radOrigin.y, 0.0f));
    const vec3 direction = normalize(decodeNormal(light.dir[0], light.dir[1]));
    const vec3 bitangent = getPerpendicularVector(direction);
    const vec3 tangent = cross(bitangent, direction);
    SunFrame f;
    f.h = cr_cos(radOrigin.x);
    f.dir[0] = direction.x; f.dir[1] = direction.y; f.dir[2] = direction.z;
    f.bitangent[0] = bitangent.x; f.bitangent[1] = bitangent.y; f.bitangent[2] = bitangent.z;
    f.tangent[0] = tangent.x; f.tangent[1] = tangent.y; f.tangent[2] = tangent.z;
    *out = f;
}
void launch_sun_frame(const LightRec* lights, SunFrame* out, cudaStream_t st) { k_sun_frame<<<1, 1, 0, st>>>(lights, out); }

// LightsView.lightCache: what lighting.comp computes per (pixel, sample) before it looks at the hit — the same expressions as k_shade
__global__ void __launch_bounds__(256) k_light_cache(const FrameMap fm, const SceneView sv, uint32_t samples, uint32_t kind, float4* __restrict__ cache) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= fm.localSlots * samples) return;
    const uint32_t s = j / fm.localSlots, i = j - s * fm.localSlots;
    uint32_t x, y;
    if (!slotToPixel(fm, i, x, y)) return;
    const vec2 loc = mk2((float)x, (float)y);
    const vec2 uv = (loc + rand2(loc + mk2(0.0f, 0.0f))) / 128.0f;
    const vec2 random = rand2(uv + hammersley(s, samples));
    float4 out = make_float4(random.x, random.y, 0.0f, 0.0f);
    if (kind == 2u) {
        float brightness, dist;
        const vec3 l = getDirToLight(sv.lights[0], mk3(0.0f, 0.0f, 0.0f), brightness, dist, random, sv.sun0);   // directional: the position is not used
        out = make_float4(l.x, l.y, l.z, 0.0f);
    }
    cache[(size_t)s * fm.w * fm.h + (size_t)y * fm.w + x] = out;
}
void launch_light_cache(const FrameMap& fm, const SceneView& sv, uint32_t samples, uint32_t kind, float4* cache, cudaStream_t st) {
    const uint32_t n = fm.localSlots * samples;
    if (n) k_light_cache<<<(n + 255) / 256, 256, 0, st>>>(fm, sv, samples, kind, cache);
}

void launch_shade(int what, const FrameMap& fm, const SceneView& sv, const CameraRec* cam, const SeedRec* seed, uint32_t samples,
                  const float4* dirT, const float4* uvN, const uint32_t* bits, uint2* lighting, float4* accum,
                  uint32_t* rgba8, uint32_t* rgba8Tiled, cudaStream_t st, const LightsView* lights) {
    if (!fm.localSlots) return;
    const uint32_t g = (fm.localSlots + 255) / 256;
    const bool ext = lights && (lights->mode != 0u || lights->historyAlpha > 0.0f);
    const LightsView lv = lights ? *lights : LightsView{};
    if (ext) {   // separate instantiations: the reference path keeps its registers
        if (what == SHADE_LIGHTING) k_shade<true, false, true><<<g, 256, 0, st>>>(fm, sv, *cam, seed, samples, dirT, uvN, bits, lighting, accum, rgba8, rgba8Tiled, lv);
        else if (what == SHADE_COMPOSITE) k_shade<false, true, false><<<g, 256, 0, st>>>(fm, sv, *cam, seed, samples, dirT, uvN, bits, lighting, accum, rgba8, rgba8Tiled, lv);
        else k_shade<true, true, true><<<g, 256, 0, st>>>(fm, sv, *cam, seed, samples, dirT, uvN, bits, lighting, accum, rgba8, rgba8Tiled, lv);
        return;
    }
    if (what == SHADE_LIGHTING) k_shade<true, false, false><<<g, 256, 0, st>>>(fm, sv, *cam, seed, samples, dirT, uvN, bits, lighting, accum, rgba8, rgba8Tiled, lv);
    else if (what == SHADE_COMPOSITE) k_shade<false, true, false><<<g, 256, 0, st>>>(fm, sv, *cam, seed, samples, dirT, uvN, bits, lighting, accum, rgba8, rgba8Tiled, lv);
    else k_shade<true, true, false><<<g, 256, 0, st>>>(fm, sv, *cam, seed, samples, dirT, uvN, bits, lighting, accum, rgba8, rgba8Tiled, lv);
}

// --------------------------------------------------------------------------------------------------------
// every light (LightsView.mode != 0): tile light lists and shadow-ray generation.  No reference counterpart beyond the sketch in
// "Raytracing optimization.md":1-14 and LIGHTS_PER_TILE (res/shaders/defines.glsl:6).
// --------------------------------------------------------------------------------------------------------
// Can light `l` add anything to a hit point at distance >= its radius?  getDirToLight (SH/light.glsl:107-126): brightness =
// pow(smoothstep(r, 0, d), specularity) with r = rad - origin, d = max(dist - origin, 0); for dist >= rad the Hermite value is 0 and
// 0^specularity = 0 when specularity > 0 — so a point light with rad > origin >= 0 and specularity > 0 is bounded by its radius.
// Everything else (directional lights, specularity <= 0, degenerate radii) reaches every pixel.
RTB_DI bool lightIsBounded(const LightRec& l, float& rad) {
    if ((l.colorBType >> 16) != LIGHT_POINT) return false;
    vec2 ro = unpackHalf2x16(l.radOrigin);
    ro = mk2(fmaxf(ro.x, 0.0f), fmaxf(ro.y, 0.0f));
    ro.y = fminf(ro.y, ro.x);
    const float spec = ubits(l.dir[0]);
    rad = ro.x;
    return ro.x > ro.y && spec > 0.0f && isfinite(ro.x);
}

// one block per 16x16-pixel tile: the box of its hit points, then every light against it, in index order
__global__ void __launch_bounds__(256) k_light_tiles(const FrameMap fm, const SceneView sv, const CameraRec cam, const float4* __restrict__ dirT,
                                                     uint32_t* __restrict__ tileCount, uint32_t* __restrict__ tileList) {
    __shared__ float sLo[3][8], sHi[3][8];
    __shared__ uint32_t sWarpCount[8], sBase;
    const uint32_t tilesX = (fm.w + 15u) >> 4;
    const uint32_t tile = blockIdx.x, tx = tile % tilesX, ty = tile / tilesX;
    const uint32_t x = tx * 16u + (threadIdx.x & 15u), y = ty * 16u + (threadIdx.x >> 4);
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    if (x < fm.w && y < fm.h) {
        const float4 dt = __ldg(dirT + (size_t)y * fm.w + x);
        if (fbits(dt.w) != NO_RAY_HIT) {
            const vec3 p = mk3(cam.eye) + mk3(dt.x, dt.y, dt.z);
            lo[0] = hi[0] = p.x; lo[1] = hi[1] = p.y; lo[2] = hi[2] = p.z;
        }
    }
    for (int a = 0; a < 3; ++a) {
        for (int o = 16; o > 0; o >>= 1) { lo[a] = fminf(lo[a], __shfl_xor_sync(0xFFFFFFFFu, lo[a], o)); hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xFFFFFFFFu, hi[a], o)); }
        if (lane == 0) { sLo[a][warp] = lo[a]; sHi[a][warp] = hi[a]; }
    }
    if (threadIdx.x == 0) sBase = 0u;
    __syncthreads();
    for (int a = 0; a < 3; ++a) for (int w = 0; w < 8; ++w) { lo[a] = fminf(lo[a], sLo[a][w]); hi[a] = fmaxf(hi[a], sHi[a][w]); }
    const bool anyHit = lo[0] <= hi[0];
    if (!anyHit) { if (threadIdx.x == 0) tileCount[tile] = 0u; return; }   // (uniform: every thread holds the same box)
    bool overflow = false;
    for (uint32_t first = 0; first < sv.info.lightCount && !overflow; first += 256u) {
        const uint32_t L = first + threadIdx.x;
        bool reaches = false;
        if (L < sv.info.lightCount) {
            const LightRec l = sv.lights[L];
            float rad;
            if (!lightIsBounded(l, rad)) reaches = true;
            else {   // squared distance from the light to the box against the radius, with slack for the rounding of both
                float d2 = 0.0f;
                for (int a = 0; a < 3; ++a) { const float c = l.pos[a], d = fmaxf(fmaxf(lo[a] - c, c - hi[a]), 0.0f); d2 += d * d; }
                const float r = rad * 1.0001f + 1e-6f;
                reaches = !(d2 > r * r);   // a NaN position stays in the list
            }
        }
        const unsigned m = __ballot_sync(0xFFFFFFFFu, reaches);
        if (lane == 0) sWarpCount[warp] = (uint32_t)__popc(m);
        __syncthreads();
        uint32_t before = sBase, total = sBase;
        for (uint32_t w = 0; w < 8u; ++w) { if (w < warp) before += sWarpCount[w]; total += sWarpCount[w]; }
        if (total > LIGHTS_PER_TILE) overflow = true;   // uniform
        else if (reaches) tileList[(size_t)tile * LIGHTS_PER_TILE + before + (uint32_t)__popc(m & ((1u << lane) - 1u))] = L;
        __syncthreads();
        if (threadIdx.x == 0) sBase = total;
        __syncthreads();
    }
    if (threadIdx.x == 0) tileCount[tile] = overflow ? LIGHT_TILE_ALL : sBase;
}
void launch_light_tiles(const FrameMap& fm, const SceneView& sv, const CameraRec* cam, const float4* dirT, uint32_t* tileCount, uint32_t* tileList, cudaStream_t st) {
    const uint32_t tiles = ((fm.w + 15u) >> 4) * ((fm.h + 15u) >> 4);
    if (!tiles) return;
    k_light_tiles<<<tiles, 256, 0, st>>>(fm, sv, *cam, dirT, tileCount, tileList);
}

// one thread per (sample, slot): the sample's random pair once (shadow.comp:84-95 — it does not depend on the light), then one
// shadow ray per light of the launch's range, built as shadow.comp builds the one for light 0
__global__ void __launch_bounds__(256) k_shadowgen_lights(const FrameMap fm, const SceneView sv, const CameraRec cam, const SeedRec* __restrict__ seed,
                                                          uint32_t samples, const float4* __restrict__ dirT, uint32_t* __restrict__ bits,
                                                          const RayQueue q, const LightsView lv) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t sample = j / fm.localSlots, i = j - sample * fm.localSlots;
    uint32_t x = 0, y = 0, object = NO_RAY_HIT;
    vec3 hitPos = mk3(0.0f, 0.0f, 0.0f);
    vec2 random = mk2(0.0f, 0.0f);
    if (sample < samples && slotToPixel(fm, i, x, y)) {
        const float4 dt = __ldg(dirT + (size_t)y * fm.w + x);
        object = fbits(dt.w);
        if (object != NO_RAY_HIT) {
            hitPos = mk3(cam.eye) + mk3(dt.x, dt.y, dt.z);
            const vec2 loc = mk2((float)x, (float)y);
            vec2 uv = (loc + rand2(loc + mk2(__ldg(&seed->randomX), __ldg(&seed->randomY)))) / 128.0f;
            uv = uv + hammersley(sample, samples);
            random = rand2(uv);
        }
    }
    const bool isHit = object != NO_RAY_HIT;
    // the lights this lane walks: the tile's list or every light, restricted to the launch's range
    uint32_t cnt = isHit ? sv.info.lightCount : 0u;
    const uint32_t* list = nullptr;
    if (isHit && lv.mode == 2u) {
        const uint32_t tile = (y >> 4) * lv.tilesX + (x >> 4), c = __ldg(lv.tileCount + tile);
        if (c != LIGHT_TILE_ALL) { cnt = c; list = lv.tileList + (size_t)tile * LIGHTS_PER_TILE; }
    }
    for (uint32_t e = 0; __any_sync(0xFFFFFFFFu, e < cnt); ++e) {
        float4 ro = make_float4(0.f, 0.f, 0.f, ubits(NO_RAY_HIT)), rd = make_float4(0.f, 0.f, 1.f, -1.0f);
        uint32_t L = 0;
        bool live = false;
        if (e < cnt) {
            L = list ? __ldg(list + e) : e;
            if (L >= lv.lightBegin && L < lv.lightEnd) {
                const LightRec light = sv.lights[L];
                float brightness, dist;
                const vec3 l = getDirToLight(light, hitPos, brightness, dist, random, L == 0u ? sv.sun0 : nullptr);
                Ray ray; ray.pos = hitPos; ray.dir = -l;
                float maxDist = -1.0f;
                if (dist >= 0.0f) {
                    const vec2 radOrigin = unpackHalf2x16(light.radOrigin);
                    if (dist >= radOrigin.y && dist < radOrigin.x) maxDist = dist - radOrigin.y;
                } else
                    maxDist = NO_HIT;
                if (maxDist != -1.0f) {
                    if (occludedByOthers(sv, ray, maxDist, object))
                        atomicOr(bits + indexToLight(x, y, fm.w, fm.h, L * samples + sample), 1u << ((x & 15u) | ((y & 1u) << 4)));
                    else if (maxDist > 0.0f || sv.sphereTree || sv.cubeTree) {
                        ro = make_float4(ray.pos.x, ray.pos.y, ray.pos.z, ubits(object));
                        rd = make_float4(ray.dir.x, ray.dir.y, ray.dir.z, maxDist);
                        live = true;
                    }
                }
            }
        }
        queueAppend(q, live, ro, rd, (L * samples + sample) * fm.localSlots + i);
    }
}
void launch_shadowgen_lights(const FrameMap& fm, const SceneView& sv, const CameraRec* cam, const SeedRec* seed, uint32_t samples, const float4* dirT,
                             uint32_t* bits, const RayQueue& queue, const LightsView& lv, cudaStream_t st) {
    const uint32_t n = fm.localSlots * samples;
    if (!n) return;
    k_shadowgen_lights<<<(n + 255) / 256, 256, 0, st>>>(fm, sv, *cam, seed, samples, dirT, bits, queue, lv);
}

// --------------------------------------------------------------------------------------------------------
// rank 0 after the gather: [nranks][slotsPerRank] tiled pixels -> scan-line order
// --------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_untile(const FrameMap fm, const uint32_t* __restrict__ tiledAll, uint32_t slotsPerRank,
                                                uint32_t* __restrict__ rgba8) {
    const uint32_t gslot = blockIdx.x * blockDim.x + threadIdx.x;   // over all global blocks * 1024
    const uint32_t g = gslot >> 10;
    if (g >= fm.blocksX * fm.blocksY) return;
    const uint32_t s = (gslot >> 5) & 31u, lane = gslot & 31u;
    const uint32_t bx = g % fm.blocksX, by = g / fm.blocksX;
    const uint32_t x = bx * 32u + (s & 3u) * 8u + (lane & 7u), y = by * 32u + (s >> 2) * 4u + (lane >> 3);
    if (x >= fm.w || y >= fm.h) return;
    const uint32_t rank = g % fm.nranks, k = g / fm.nranks;
    rgba8[(size_t)y * fm.w + x] = __ldg(tiledAll + (size_t)rank * slotsPerRank + (size_t)k * 1024u + (gslot & 1023u));
}
void launch_untile(const FrameMap& fm, const uint32_t* tiledAll, uint32_t slotsPerRank, uint32_t* rgba8, cudaStream_t st) {
    const uint32_t n = fm.blocksX * fm.blocksY * 1024u;
    if (!n) return;
    k_untile<<<(n + 255) / 256, 256, 0, st>>>(fm, tiledAll, slotsPerRank, rgba8);
}


// --------------------------------------------------------------------------------------------------------
// presentation straight into host memory: this rank's pixels to their scan-line positions of a mapped, page-locked frame.
// One warp writes one 32-pixel row of a block = 128 contiguous bytes (one PCIe write of a useful size).
// --------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_present_host(const FrameMap fm, const uint32_t* __restrict__ tiled, const uint32_t* __restrict__ rgba8,
                                                      uint32_t* __restrict__ hostFrame) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;      // local block k, then 5 bits of row, 5 bits of column
    const uint32_t k = t >> 10, r = (t >> 5) & 31u, cx = t & 31u;
    if (k >= fm.localBlocks) return;
    const uint32_t g = k * fm.nranks + fm.rank;
    const uint32_t bx = g % fm.blocksX, by = g / fm.blocksX;
    const uint32_t x = bx * 32u + cx, y = by * 32u + r;
    if (x >= fm.w || y >= fm.h) return;
    const uint32_t slot = k * 1024u + ((r >> 2) * 4u + (cx >> 3)) * 32u + (r & 3u) * 8u + (cx & 7u);
    hostFrame[(size_t)y * fm.w + x] = tiled ? __ldg(tiled + slot) : __ldg(rgba8 + (size_t)y * fm.w + x);
}
void launch_present_host(const FrameMap& fm, const uint32_t* tiled, const uint32_t* rgba8, uint32_t* hostFrame, cudaStream_t st) {
    const uint32_t n = fm.localBlocks * 1024u;
    if (!n) return;
    k_present_host<<<(n + 255) / 256, 256, 0, st>>>(fm, tiled, rgba8, hostFrame);
}

}  // namespace rtb
#include "rtb_path.cuh"
