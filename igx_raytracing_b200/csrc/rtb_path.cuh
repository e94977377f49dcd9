// rtb_path.cuh — wavefront path tracing: the diffuse-bounce workload of BASELINE.json configs[3] (included by rtb_kernels.cu).
//
// The reference has NO bounce rays (its "reflection" is one skybox tap, SH/light.glsl:205-219), so what a bounce
// is is defined here, after SURVEY.md 8(d) config 4, from the reference's own building blocks:
//
//   vertex 0     the reference's G-buffer: primary ray, nearest hit, shading normal through the 3 x 16-bit encoding
//                (raygen.comp) — produced by the same launches RTB_PASS_FRAME uses;
//   per vertex   radiance += throughput * emissive;
//                one shadow ray to lights[0] exactly as shadow.comp builds it (getDirToLight, self-exclusion by object id, the
//                point-light range rule), random = rand(uvBase + hammersley(2 * depth, 2 * (bounces + 1))) with
//                uvBase = (pixel + rand(pixel + seed.random)) / 128 — at depth 0 that is shadow.comp's sample 0 of 1;
//                unoccluded: radiance += throughput * shadeLight(...) * lightCount (lighting.comp's Cook-Torrance term);
//   bounce       direction cosine-distributed about the shading normal turned against the incoming ray
//                (phi = 2 pi u1, cos theta = sqrt(1 - u2), frame from getPerpendicularVector), (u1, u2) =
//                rand(uvBase + hammersley(2 * depth + 1, 2 * (bounces + 1))); throughput *= albedo (the cosine and 1/pi cancel
//                against the density); origin = hit point, previous object excluded by id like the reference's shadow rays;
//                no Russian roulette;
//   miss         radiance += throughput * sampleSkybox(direction); the path ends;
//   pixel        composite.comp's tail: optional progressive accumulation, 1 - exp(-c * exposure), rgba8.
//
// Parity: depth 0 is the reference path (checked bit for bit like RTB_PASS_RAYGEN); deeper vertices are held against a CPU
// statement of THIS definition kept with the tests, on small scenes with every primitive type, and against RTB_ACCEL_BRUTE on
// a 1/64 tile subsample of the 10M-triangle frame (tests/test_gpu_path.py).  All arithmetic follows the numerics contract of rtb_math.cuh.
//
// Wavefront: rays live in compact queues (warp-aggregated append, as k_shadowgen's), one nearest-hit launch and one occlusion
// launch per depth over the queue's device-side count; a path's state is its wavefront slot's throughput and radiance.
#pragma once

namespace rtb {

// PathBuffers (rtb_kernels.cuh), per wavefront slot: throughput.xyz, radiance.xyz, and `direct` = what the vertex's shadow ray
// adds when it is not occluded

struct VertexOut { bool shadow, bounce; float4 so, sd, bo, bd; vec3 direct; };

// One path vertex: pos, unit incoming direction v, object id, shading normal n as the G-buffer decodes it.
RTB_DI VertexOut pathVertex(const SceneView& sv, vec2 uvBase, uint32_t depth, uint32_t bounces, vec3 pos, vec3 v, uint32_t object, vec3 n,
                            vec3& T, vec3& L) {
    VertexOut o;
    o.shadow = false; o.bounce = false; o.direct = mk3(0.0f, 0.0f, 0.0f);
    const MatU m = unpackMaterial(sv.materials + __ldg(sv.materialIndices + object));
    L = L + T * m.emissive;
    const uint32_t N = 2u * (bounces + 1u);
    // ---- direct light: shadow.comp's ray, lighting.comp's term --------------------------------------------------------
    if (sv.info.lightCount) {
        const vec2 random = rand2(uvBase + hammersley(2u * depth, N));
        const LightRec light = sv.lights[0];
        const vec3 F0 = mix(mk3(0.04f, 0.04f, 0.04f), m.albedo, m.metallic);
        const float NdotV = fmaxf(dot(v, -n), 0.0f);
        const vec3 c = shadeLight(F0, m.albedo, m.roughness, m.metallic, light, pos, n, v, NdotV, random, sv.sun0) * (float)sv.info.lightCount;
        const vec3 contribution = T * c;
        float brightness, dist;
        const vec3 l = getDirToLight(light, pos, brightness, dist, random, sv.sun0);
        float maxDist = -1.0f;
        if (dist >= 0.0f) {
            const vec2 radOrigin = unpackHalf2x16(light.radOrigin);
            if (dist >= radOrigin.y && dist < radOrigin.x) maxDist = dist - radOrigin.y;
        } else
            maxDist = NO_HIT;
        if (maxDist != -1.0f) {   // a shadow ray decides
            o.shadow = true; o.direct = contribution;
            const vec3 d = -l;
            o.so = make_float4(pos.x, pos.y, pos.z, ubits(object));
            o.sd = make_float4(d.x, d.y, d.z, maxDist);
        } else
            L = L + contribution;   // out of the light's traced range: shadow.comp reports "not occluded"
    }
    // ---- bounce ----------------------------------------------------------------------------------------------------------
    if (depth < bounces) {
        const vec2 r = rand2(uvBase + hammersley(2u * depth + 1u, N));
        const vec3 nn = normalize(n);
        const vec3 nf = dot(v, nn) > 0.0f ? -nn : nn;            // against the incoming ray
        const float phi = (2.0f * PI_F) * r.x;
        const float cosT = sqrtf(1.0f - r.y), sinT = sqrtf(r.y);
        const float x = cr_cos(phi) * sinT, y = cr_sin(phi) * sinT;
        const vec3 bitangent = normalize(getPerpendicularVector(nf));
        const vec3 tangent = cross(bitangent, nf);
        const vec3 d = normalize(bitangent * x + tangent * y + nf * cosT);
        T = T * m.albedo;
        o.bounce = true;
        o.bo = make_float4(pos.x, pos.y, pos.z, ubits(object));
        o.bd = make_float4(d.x, d.y, d.z, NO_HIT);
    }
    return o;
}

RTB_DI vec2 pathUvBase(uint32_t x, uint32_t y, const SeedRec* seed) {
    const vec2 loc = mk2((float)x, (float)y);
    return (loc + rand2(loc + mk2(__ldg(&seed->randomX), __ldg(&seed->randomY)))) / 128.0f;
}

// depth 0: from the G-buffer
__global__ void __launch_bounds__(256) k_path_start(const FrameMap fm, const SceneView sv, const CameraRec cam, const SeedRec* __restrict__ seed,
                                                    uint32_t bounces, const float4* __restrict__ dirT, const float4* __restrict__ uvN,
                                                    const PathBuffers pb, const RayQueue shadowQ, const RayQueue nextQ) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t x = 0, y = 0;
    const bool valid = i < fm.localSlots && slotToPixel(fm, i, x, y);
    VertexOut o;
    o.shadow = false; o.bounce = false;
    if (valid) {
        const size_t px = (size_t)y * fm.w + x;
        const float4 dt = __ldg(dirT + px), un = __ldg(uvN + px);
        const uint32_t object = fbits(dt.w);
        const vec3 dxyz = mk3(dt.x, dt.y, dt.z);
        vec3 T = mk3(1.0f, 1.0f, 1.0f), L = mk3(0.0f, 0.0f, 0.0f);
        if (object == NO_RAY_HIT) {
            const SkyView sky = {sv.skybox, sv.skyW, sv.skyH};
            L = sampleSkybox(sky, cam, normalize(dxyz));
        } else {
            const vec3 pos = mk3(cam.eye) + dxyz;
            o = pathVertex(sv, pathUvBase(x, y, seed), 0u, bounces, pos, normalize(dxyz), object, decodeNormal(fbits(un.z), fbits(un.w)), T, L);
        }
        pb.throughput[i] = make_float4(T.x, T.y, T.z, 0.0f);
        pb.radiance[i] = make_float4(L.x, L.y, L.z, 0.0f);
        if (o.shadow) pb.direct[i] = make_float4(o.direct.x, o.direct.y, o.direct.z, 0.0f);
    }
    queueAppend(shadowQ, o.shadow, o.so, o.sd, i);
    queueAppend(nextQ, o.bounce, o.bo, o.bd, i);
}

// depth >= 1: one thread per ray of the queue the nearest-hit launch has just answered
__global__ void __launch_bounds__(256) k_path_vertex(const FrameMap fm, const SceneView sv, const CameraRec cam, const SeedRec* __restrict__ seed,
                                                     uint32_t depth, uint32_t bounces, const RayQueue inQ, const TriHit* __restrict__ hits,
                                                     const PathBuffers pb, const RayQueue shadowQ, const RayQueue nextQ) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = r < __ldg(inQ.count);
    VertexOut o;
    o.shadow = false; o.bounce = false;
    uint32_t slot = 0;
    if (valid) {
        slot = inQ.slotIds[r];
        const float4 ro = __ldg(reinterpret_cast<const float4*>(inQ.rays + r)), rd = __ldg(reinterpret_cast<const float4*>(inQ.rays + r) + 1);
        const float4 hv = __ldg(reinterpret_cast<const float4*>(hits + r));
        TriHit th; th.t = hv.x; th.id = fbits(hv.y); th.u = hv.z; th.v = hv.w;
        Ray ray; ray.pos = mk3(ro.x, ro.y, ro.z); ray.dir = mk3(rd.x, rd.y, rd.z);
        Hit hit; vec3 objectNormal;
        finishGeometry(sv, ray, fbits(ro.w), th, hit, objectNormal);
        const float4 t4 = pb.throughput[slot], l4 = pb.radiance[slot];
        vec3 T = mk3(t4.x, t4.y, t4.z), L = mk3(l4.x, l4.y, l4.z);
        if (hit.hitT == NO_HIT) {
            const SkyView sky = {sv.skybox, sv.skyW, sv.skyH};
            L = L + T * sampleSkybox(sky, cam, ray.dir);
        } else {
            uint32_t ex, ey, x, y;
            encodeNormalGpu(objectNormal, ex, ey);            // the shading normal as the reference's G-buffer would hold it
            slotToPixel(fm, slot, x, y);
            const vec3 pos = ray.pos + ray.dir * hit.hitT;
            o = pathVertex(sv, pathUvBase(x, y, seed), depth, bounces, pos, ray.dir, hit.object, decodeNormal(ex, ey), T, L);
        }
        pb.throughput[slot] = make_float4(T.x, T.y, T.z, 0.0f);
        pb.radiance[slot] = make_float4(L.x, L.y, L.z, 0.0f);
        if (o.shadow) pb.direct[slot] = make_float4(o.direct.x, o.direct.y, o.direct.z, 0.0f);
    }
    queueAppend(shadowQ, o.shadow, o.so, o.sd, slot);
    queueAppend(nextQ, o.bounce, o.bo, o.bd, slot);
}

// the vertex's shadow ray has been answered: add the direct term of the unoccluded ones
__global__ void __launch_bounds__(256) k_path_shadow_resolve(const RayQueue shadowQ, const uint8_t* __restrict__ occOthers, const uint8_t* __restrict__ occTris,
                                                             const PathBuffers pb) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= __ldg(shadowQ.count)) return;
    if (occOthers[r] | occTris[r]) return;
    const uint32_t slot = shadowQ.slotIds[r];
    const float4 l4 = pb.radiance[slot], d4 = pb.direct[slot];
    pb.radiance[slot] = make_float4(l4.x + d4.x, l4.y + d4.y, l4.z + d4.z, 0.0f);
}

// composite.comp's tail on the path's radiance (SH/composite.comp:249-285)
__global__ void __launch_bounds__(256) k_path_resolve(const FrameMap fm, const SceneView sv, const CameraRec cam, const SeedRec* __restrict__ seed, const PathBuffers pb,
                                                      float4* __restrict__ accum, uint32_t* __restrict__ rgba8, uint32_t* __restrict__ rgba8Tiled) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= fm.localSlots) return;
    uint32_t x, y;
    if (!slotToPixel(fm, i, x, y)) { if (rgba8Tiled) rgba8Tiled[tiledSlot(fm, i)] = 0u; return; }
    const size_t px = (size_t)y * fm.w + x;
    const float4 l4 = pb.radiance[i];
    vec3 color = mk3(l4.x, l4.y, l4.z);
    if (!sv.releaseBuild && (isnan(color.x) || isnan(color.y) || isnan(color.z))) color = mk3(0.0f, 0.0f, 10000.0f);
    if (cam.flags & CAMERA_USE_SUPERSAMPLING) {
        const uint32_t sampleCount = __ldg(&seed->sampleCount);
        if (sampleCount > 1) { const float4 p = accum[px]; color = color + mk3(p.x, p.y, p.z); }
        accum[px] = make_float4(color.x, color.y, color.z, 0.0f);
        color = color / (float)sampleCount;
    }
    const vec3 e = -color * cam.exposure;
    color = vmax(mk3(1.0f, 1.0f, 1.0f) - mk3(cr_exp(e.x), cr_exp(e.y), cr_exp(e.z)), mk3(0.0f, 0.0f, 0.0f));
    const uint32_t out = unorm8(color.x) | (unorm8(color.y) << 8) | (unorm8(color.z) << 16) | (255u << 24);
    rgba8[px] = out;
    if (rgba8Tiled) rgba8Tiled[tiledSlot(fm, i)] = out;
}

void launch_path_start(const FrameMap& fm, const SceneView& sv, const CameraRec* cam, const SeedRec* seed, uint32_t bounces, const float4* dirT, const float4* uvN,
                       const PathBuffers& pb, const RayQueue& shadowQ, const RayQueue& nextQ, cudaStream_t st) {
    if (!fm.localSlots) return;
    k_path_start<<<(fm.localSlots + 255) / 256, 256, 0, st>>>(fm, sv, *cam, seed, bounces, dirT, uvN, pb, shadowQ, nextQ);
}
void launch_path_vertex(const FrameMap& fm, const SceneView& sv, const CameraRec* cam, const SeedRec* seed, uint32_t depth, uint32_t bounces, const RayQueue& inQ,
                        const TriHit* hits, const PathBuffers& pb, const RayQueue& shadowQ, const RayQueue& nextQ, cudaStream_t st) {
    if (!fm.localSlots) return;
    k_path_vertex<<<(fm.localSlots + 255) / 256, 256, 0, st>>>(fm, sv, *cam, seed, depth, bounces, inQ, hits, pb, shadowQ, nextQ);
}
void launch_path_shadow_resolve(const RayQueue& shadowQ, uint32_t maxRays, const uint8_t* occOthers, const uint8_t* occTris, const PathBuffers& pb, cudaStream_t st) {
    if (!maxRays) return;
    k_path_shadow_resolve<<<(maxRays + 255) / 256, 256, 0, st>>>(shadowQ, occOthers, occTris, pb);
}
void launch_path_resolve(const FrameMap& fm, const SceneView& sv, const CameraRec* cam, const SeedRec* seed, const PathBuffers& pb, float4* accum, uint32_t* rgba8,
                         uint32_t* rgba8Tiled, cudaStream_t st) {
    if (!fm.localSlots) return;
    k_path_resolve<<<(fm.localSlots + 255) / 256, 256, 0, st>>>(fm, sv, *cam, seed, pb, accum, rgba8, rgba8Tiled);
}

}  // namespace rtb
