/*
 * rtb200.h — C ABI of librtb200, the B200 (sm_100a) replacement for the compute-dispatch and
 * buffer-upload layer under igx_raytracing's hot path.
 *
 * The reference drives its five compute shaders (init -> raygen -> shadow -> lighting -> composite)
 * by recording ignis commands: FlushBuffer / FlushImage / BindDescriptors / BindPipeline / Dispatch,
 * then Graphics::present / presentToCpu.  Every entry point below names the reference call site it
 * replaces ("ref:" paths are relative to the reference root; "IGNIS/" = igx/igxi-tool/igxi/ignis/).
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success and a non-zero
 * rtb_status on failure (the reference aborts through log()->fatal; we return a code and keep a
 * message, see rtb_last_error).  A context is bound to one CUDA device and, like ignis::Graphics
 * (ref: IGNIS/api/opengl/src/graphics/gl_graphics.cpp:88), must be driven by one thread at a time.
 * All work is stream-ordered on the context's stream; rtb_readback / rtb_sync synchronise.
 *
 * There is no CPU fallback: every call fails with RTB_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef RTB200_H
#define RTB200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct rtb_ctx rtb_ctx;

typedef enum rtb_status {
    RTB_OK = 0,
    RTB_ERR_ARG = 1,      /* bad argument / out of range (reference: oicAssert / silently skipped command) */
    RTB_ERR_CUDA = 2,     /* CUDA runtime error, message holds cudaGetErrorString */
    RTB_ERR_STATE = 3,    /* call order violated (e.g. dispatch before resize) */
    RTB_ERR_CAPACITY = 4  /* upload past the capacity given at rtb_create (reference: add() returns 0) */
} rtb_status;

/* Object capacities.  The reference hard-codes 65536/32768/16384/256 in the SceneGraph constructor
 * (ref: igx/include/helpers/scene_graph.hpp:132-146, initialiser igx/src/helpers/scene_graph.cpp:86-95);
 * here they are explicit so 1M / 10M-triangle scenes fit. */
typedef struct rtb_limits {
    uint32_t max_triangles, max_spheres, max_cubes, max_planes, max_lights, max_materials;
} rtb_limits;

/* Buffers a host may upload: one per reference GPUBuffer on the path. */
typedef enum rtb_buffer {
    RTB_BUF_CAMERA = 0,        /* 144 B  ref: src/rt/raytracing_interface.cpp:21-26,327-328 */
    RTB_BUF_SEED = 1,          /* 24 B   ref: src/rt/task/composite_task.cpp:31-36,239-243 */
    RTB_BUF_SCENE_INFO = 2,    /* 36 B   ref: igx/src/helpers/scene_graph.cpp:144-154 */
    RTB_BUF_SHADOW_PROPS = 3,  /* 4 B    ref: src/rt/task/shadow_task.cpp:32-35,189-192 */
    RTB_BUF_TRIANGLES = 4,     /* 48 B each   ref: igx/src/helpers/scene_graph.cpp:116-128,297-303 */
    RTB_BUF_SPHERES = 5,       /* 16 B each */
    RTB_BUF_CUBES = 6,         /* 24 B each */
    RTB_BUF_PLANES = 7,        /* 16 B each */
    RTB_BUF_LIGHTS = 8,        /* 32 B each */
    RTB_BUF_MATERIALS = 9,     /* 32 B each */
    RTB_BUF_MATERIAL_INDICES = 10, /* u32 per global object id  ref: igx/src/helpers/scene_graph.cpp:130-138,478-489 */
    RTB_BUF_COUNT = 11
} rtb_buffer;

/* Passes: one per reference Dispatch, plus the whole frame. */
typedef enum rtb_pass {
    RTB_PASS_INIT = 0,       /* init.comp        ref: src/rt/task/composite_task.cpp:260-264 */
    RTB_PASS_RAYGEN = 1,     /* raygen.comp      ref: src/rt/task/raygen_task.cpp:88-94 */
    RTB_PASS_SHADOW = 2,     /* shadow.comp      ref: src/rt/task/shadow_task.cpp:194-204 */
    RTB_PASS_LIGHTING = 3,   /* lighting.comp    ref: src/rt/task/shadow_task.cpp:206-210 */
    RTB_PASS_COMPOSITE = 4,  /* composite.comp   ref: src/rt/task/composite_task.cpp:266-276 */
    RTB_PASS_FRAME = 5       /* the recorded command list replayed once: INIT..COMPOSITE
                                ref: src/rt/raytracing_interface.cpp:144-179 */
} rtb_pass;

/* Render targets that can be read back: one per reference texture / buffer written by the path. */
typedef enum rtb_target {
    RTB_TGT_DIR_T = 0,        /* rgba32f  w*h*16 B   ref: src/rt/task/raygen_task.cpp:18-20 */
    RTB_TGT_UV_NORMAL = 1,    /* rgba32f  w*h*16 B */
    RTB_TGT_SHADOW_BITS = 2,  /* u32[ceil(w/16)*ceil(h/2)*samples]  ref: src/rt/task/shadow_task.cpp:146-165 (x lightCount with RTB_OPT_LIGHTS) */
    RTB_TGT_LIGHTING = 3,     /* rgba16f  w*h*8 B    ref: src/rt/task/shadow_task.cpp:20-22 */
    RTB_TGT_ACCUM = 4,        /* rgba32f  w*h*16 B   ref: src/rt/task/composite_task.cpp:20-22 */
    RTB_TGT_RGBA8 = 5,        /* rgba8    w*h*4 B    (what presentToCpu copies out) */
    RTB_TGT_SEED = 6,         /* 24 B: the Seed buffer after K0 */
    RTB_TGT_RGBA8_TILED = 7,  /* multi-GPU only: this rank's pixels in wavefront-slot order; ceil(blocks/n)*4096 B on every rank (see rtb_untile) */
    RTB_TGT_ACCEL_NODES = 8,  /* the acceleration structure's node records (rtb_accel_info.node_count * node_bytes): inspection / tests */
    RTB_TGT_ACCEL_TRIANGLES = 9, /* its traversal triangles (48 B each, leaf order) */
    RTB_TGT_COUNT = 10
} rtb_target;

/* How the nearest-hit search runs.  The reference has exactly one way: a linear loop over every
 * primitive (ref: res/shaders/trace.glsl:25-45,78-94).  BVH is new and returns the same hits. */
typedef enum rtb_accel_mode {
    RTB_ACCEL_BRUTE = 0,   /* the reference algorithm, verbatim, on the GPU */
    RTB_ACCEL_BVH = 1,     /* triangles through the 8-wide compressed BVH (128-byte nodes); spheres and cubes through trees of their own from
                              RTB_OPT_PRIMITIVE_TREES primitives of a type on, in the reference's loops below that; planes stay linear */
    RTB_ACCEL_BVH2 = 2     /* triangles through the binary BVH (64-byte two-box nodes): the first-generation kernel, kept for comparison */
} rtb_accel_mode;

typedef struct rtb_accel_info {
    uint32_t mode;            /* rtb_accel_mode in effect */
    uint32_t node_count;      /* BVH nodes */
    uint32_t node_bytes;      /* bytes per node record */
    uint32_t leaf_count;
    uint32_t max_depth;
    uint32_t tri_record_bytes;/* bytes per traversal triangle record */
    float    sah_cost;
    float    build_ms;        /* host wall clock of the last build */
    float    leaf_node_extent;/* 8-wide tree: mean edge length of the nodes that hold only triangles (world units) */
    uint32_t refits;          /* rtb_refit_accel calls served by a device refit since the last host build */
    uint32_t primary_packets; /* how the last RAYGEN / FRAME dispatch walked the camera rays: 0 per ray, 1 union packets, 3 frustum packets */
    uint32_t builder;         /* who built the tree in use: 0 host (binned SAH + optimal collapse), 1 device (RTB_OPT_ACCEL_BUILDER) */
    uint32_t sphere_tree_nodes, cube_tree_nodes; /* nodes of the sphere / cube trees the last dispatch or rays-in call searched
                                                    (RTB_OPT_PRIMITIVE_TREES); 0 = that type is in the reference's linear loop */
} rtb_accel_info;

/* Counters of the last instrumented dispatch (rtb_set_option(RTB_OPT_COUNTERS, 1)); the timed build
 * never touches them.  Used for the roofline's algorithmic bytes per ray. */
typedef struct rtb_counters {
    uint64_t primary_rays, shadow_rays;          /* rays handed to the nearest-hit / occlusion search */
    uint64_t primary_nodes, primary_tris;        /* node records fetched, triangle records tested */
    uint64_t shadow_nodes, shadow_tris;
    uint64_t primary_hits, shadow_occluded;      /* pixels whose nearest hit is any primitive; shadow rays found occluded by a triangle */
} rtb_counters;

typedef enum rtb_option {
    RTB_OPT_COUNTERS = 0,     /* 0: off.  1: instrumented per-ray traversal (the ALGORITHMIC node / triangle records each ray
                                 needs).  2: instrumented build of the kernels in use (records actually fetched: a packet
                                 fetch serves 32 rays and counts once) */
    RTB_OPT_TILE_RANK = 1,    /* multi-GPU screen partition: this context renders tiles t with t % count == rank */
    RTB_OPT_TILE_COUNT = 2,
    RTB_OPT_SHADER_BUILD = 3,      /* which build of the reference shaders the passes behave like (ref: res/shaders/compile.sh:3-9).
                                      0 (default) = DEBUG, what the shipped .spv binaries are: every pass stores for every pixel
                                      (uvObjectNormal and a zero lighting texel on misses, zero shadow words for subgroups
                                      without hits: raygen.comp:46-51, nv_all.shadow.comp:69-82, nv_all.lighting.comp:55-62),
                                      triangles with p1 == p0 are rejected (primitive.glsl:248-253; only the reference loop
                                      RTB_ACCEL_BRUTE can accept one at all), a NaN colour is shown as (0,0,10000)
                                      (composite.comp:236-239).  1 = RELEASE: those stores are skipped (the targets keep what
                                      they held; zero after rtb_resize), no reject, no NaN mapping. */
    RTB_OPT_PRIMARY_PACKETS = 4,   /* nearest-hit search of the camera rays with RTB_ACCEL_BVH.  0 = one traversal per ray.
                                      1 = one warp-cooperative traversal per 8x4-pixel patch, every ray testing every
                                      child box of the union ("union packets").  3 = the same walk with the box tests
                                      done once per packet against the interval rays of the patch's four quadrants
                                      ("frustum packets"; meant for rays sharing one origin — other packets are still
                                      answered exactly, only slowly).  2 = default: frustum packets when the projection
                                      is Default and the patch is small against the tree's leaf nodes, else per ray.
                                      The hits are identical in every mode.  With 1 or 3 the rays-in call rtb_trace_rays
                                      also walks its rays in packets of 32 consecutive rays. */
    RTB_OPT_FUSE_PRIMARY = 5,      /* 0/1 (default 1): with frustum packets, generate the camera rays inside the traversal launch and
                                      write the G-buffer from its epilogue (one launch instead of three; same bits).  0 keeps
                                      the three launches, e.g. to time them apart. */
    RTB_OPT_ACCEL_BUILDER = 7,     /* who builds the 8-wide tree in rtb_build_accel(RTB_ACCEL_BVH).  0 (default) = the host: multi-threaded
                                      binned SAH, SAH-optimal collapse; 0.7 s per million triangles.  1 = the device: Morton sort, radix tree
                                      (Karras 2012), greedy collapse by surface area, boxes by the refit kernels — tens of
                                      milliseconds, so that add / del / compaction of triangles (ref: igx/src/helpers/
                                      scene_graph.cpp:343-376,378-522) need not stall the frame; same hits, somewhat higher SAH
                                      cost.  Falls back to the host builder for fewer than 2 triangles or a tree too deep for
                                      the traversal stack. */
    RTB_OPT_FRAME_OVERLAP = 13,    /* default 1: recorded frames (RTB_OPT_FRAME_GRAPH) of consecutive RTB_PASS_FRAME dispatches overlap — init +
                                      camera rays + nearest hit + G-buffer of frame k+1 run on a second stream while shadow rays, lighting
                                      and composite of frame k drain (shadow rays + occlusion on a third stream beside the shade launch before
                                      them), on two sets of the buffers one stage hands to the next (shadow words and the
                                      Seed as init.comp left it included).  Same launches, same pixels; the tails of the persistent
                                      launches are filled (one rank of eight: 0.65 -> 0.57 ms per 4K soup frame).  Everything else the
                                      API offers is ordered after all of them; pointers from rtb_device_ptr(DIR_T / UV_NORMAL / SHADOW_BITS)
                                      name the latest frame's set and are valid until the next RTB_PASS_FRAME (or rtb_path_frame).  0: one frame after the other. */
    RTB_OPT_LIGHT_CACHE = 14,      /* default 1: the random pair lighting.comp derives per pixel and shadow sample depends on the pixel, the
                                      sample index and the sample count only (its Seed block is unbound: it reads zeros), so it — and for a
                                      directional light 0 the light direction it yields — is computed once per frame size (and light 0) and
                                      read back by the lighting pass: 16 bytes per pixel and sample instead of four exact-reduction binary64
                                      sines (and the sun-disc sample) per pixel and frame.  Same values, same pixels.  Not kept beyond
                                      1.5 GB.  0: evaluated in the kernel every frame. */
    RTB_OPT_PRIMITIVE_TREES = 12,  /* spheres and cubes stay in the reference's linear loops (ref: res/shaders/trace.glsl:31-40,83-90) while a
                                      type has fewer than this many primitives (default 64; 0 = always); from there on the type gets
                                      an 8-wide tree of its own, built on the device over the primitives' boxes and rebuilt when its
                                      buffer or count changes, and a traversal that returns what the loop returns — lowest index on
                                      equal sphere distances, highest on equal cube distances, cubes entered from inside included
                                      (the reference allows 32768 spheres and 16384 cubes: ref: igx/include/helpers/
                                      scene_graph.hpp:137-142).  RTB_ACCEL_BRUTE keeps every loop.  Planes are infinite: always looped. */
    RTB_OPT_LIGHTS = 10,           /* BEYOND THE REFERENCE, which samples lights[0] only and scales by lightCount (ref: res/shaders/
                                      nv_all.shadow.comp:97, nv_all.lighting.comp:88,98).  0 (default) = that.
                                      1 = every light gets its own shadow ray per sample and its own Cook-Torrance term; the shadow
                                      mask then has lightCount * samples layers (layer = light * samples + sample) and the sum is
                                      not scaled.  2 = the same result through per-tile light lists: for every 16x16-pixel screen
                                      tile the lights that can reach one of its hit points (at most LIGHTS_PER_TILE = 32 — ref:
                                      res/shaders/defines.glsl:6, "Raytracing optimization.md":1-14; a tile with more keeps the
                                      full loop), so that shadow-ray set-up and lighting walk 32 lights instead of thousands. */
    RTB_OPT_HISTORY_ALPHA = 11,    /* BEYOND THE REFERENCE: the bits of a float a in [0,1].  a > 0: the lighting texture is blended
                                      over frames through the History texture the reference allocates but never uses (ref: src/rt/
                                      task/shadow_task.cpp:20-22,212 "Do denoising"): history = history * (1 - a) + lighting * a,
                                      stored in both; composite reads the blend.  The first frame after rtb_resize or an option
                                      change starts the history (a = 1).  0 (default) = off. */
    RTB_OPT_FRAME_GRAPH = 9,       /* 1 (default): RTB_PASS_FRAME is recorded into two CUDA graphs (everything before the shade launch /
                                      the shade launch) the second time it is dispatched with nothing changed, and replayed from
                                      then on — what the reference does with its CommandList (recorded once, replayed per frame
                                      and N times for an export: ref: src/rt/raytracing_interface.cpp:144-179,222-226).  Anything
                                      the launches hold by value (camera, scene info, options, size, acceleration structure)
                                      re-records; uploads into device buffers (seed, geometry, materials) do not.  0: every frame
                                      is launched directly (rtb_last_frame_ms needs this). */
    RTB_OPT_FRAME_LANES = 8,       /* 2: RTB_PASS_FRAME runs as two half-frame lanes (the rank's even / odd 32x32 blocks) on two
                                      streams, so that while one lane's persistent traversal launch drains its last, longest rays
                                      the other lane's launch takes the freed SM slots; frames of fewer than 128 blocks and
                                      instrumented frames run as one lane.  1 (default): one lane.  Same pixels either way; measured
                                      on B200 the halved launches cost more than the filled tails give back. */
    RTB_OPT_SHADOW_ORDER = 6       /* how the occlusion rays reach the traversal kernel (RTB_ACCEL_BVH; same shadow bits in every
                                      mode).  0 = one record per pixel and sample in wavefront-slot order, misses included.
                                      1 (default) = the live rays appended to a queue (no dead records travel: occlusion launch
                                      2.14 -> 1.99 ms on the 4K soup frame).  2 = that queue sorted by a 2D light-space
                                      coordinate (counting sort over Morton-ordered cells), so that rays travelling along
                                      neighbouring lines — whatever depth they start from — sit in one warp: 1.92 ms, but the
                                      sort costs 0.13 ms, because a per-ray traversal pays its L1 wavefronts per lane whether
                                      or not its neighbours want the same node.  3 = the sorted queue walked
                                      as beam packets (csrc/rtb_trace8b.cuh: one interval-ray box test per packet): meant for
                                      lights without an angular extent — a cone-sampled sun spreads 32 neighbouring rays over
                                      several leaf nodes and the walk is slower than the per-ray kernel (62 ms on the soup). */
} rtb_option;

/* ---- lifetime ------------------------------------------------------------------------------------ */
/* replaces ignis::Graphics + FactoryContainer creation (ref: test/main.cpp:8-12) */
int  rtb_create(rtb_ctx** out, int cuda_device, const rtb_limits* limits);
void rtb_destroy(rtb_ctx* ctx);
const char* rtb_last_error(const rtb_ctx* ctx);   /* ctx may be NULL: error of the last failed rtb_create */
int  rtb_set_option(rtb_ctx* ctx, rtb_option opt, uint32_t value);
/* run on a caller-owned cudaStream_t (passed as void*); NULL = the context's own stream */
int  rtb_set_stream(rtb_ctx* ctx, void* cuda_stream);

/* ---- resources ----------------------------------------------------------------------------------- */
/* replaces TextureRenderTask::resize / ShadowTask::resize / CompositeTask::resize
 * (ref: igx/include/helpers/render_task.hpp:67-97, src/rt/task/shadow_task.cpp:139-187,
 *  src/rt/task/composite_task.cpp:205-233) */
int  rtb_resize(rtb_ctx* ctx, uint32_t width, uint32_t height, uint32_t shadow_samples);
/* replaces GPUBuffer::getBuffer()+flush() and cmd::FlushBuffer
 * (ref: IGNIS/include/graphics/command/commands.hpp FlushBuffer; call sites listed per rtb_buffer) */
int  rtb_upload(rtb_ctx* ctx, rtb_buffer id, size_t byte_offset, size_t bytes, const void* src);
/* replaces cmd::FlushImage(skybox) (ref: igx/src/helpers/scene_graph.cpp:97-101,253-257).
 * rgba16f, row 0 first, w*h*8 bytes; w == 0 or pixels == NULL removes the skybox (camera.skyboxColor is used). */
int  rtb_upload_skybox(rtb_ctx* ctx, uint32_t width, uint32_t height, const uint16_t* rgba16f);
/* NEW (no reference counterpart): (re)build the acceleration structure over the uploaded triangles.
 * Must be called after triangle uploads and before a dispatch when mode == RTB_ACCEL_BVH. */
int  rtb_build_accel(rtb_ctx* ctx, rtb_accel_mode mode);
/* NEW: after triangles MOVED (same count; uploads into RTB_BUF_TRIANGLES), recompute every box of the existing tree on
 * the device instead of rebuilding it on the host — what a per-frame SceneGraph::update of dirty triangle ranges needs
 * (ref: igx/src/helpers/scene_graph.cpp:267-323, test/scene/niels_scene.cpp:61-70).  Stream-ordered, no host work
 * beyond the launches.  Topology is kept, so hits stay exact but traversal cost grows with the deformation; rebuild
 * when it matters: rtb_accel_info.sah_cost is recomputed by every refit.  Falls back to rtb_build_accel when there is no 8-wide tree of the current
 * triangle count (first call, other mode, count changed). */
int  rtb_refit_accel(rtb_ctx* ctx);
int  rtb_accel_info_get(const rtb_ctx* ctx, rtb_accel_info* out);

/* ---- execution ----------------------------------------------------------------------------------- */
/* replaces cmd::BindPipeline + cmd::BindDescriptors + cmd::Dispatch (ref: IGNIS/api/opengl/src/graphics/
 * command/gl_command_list.cpp:311-335); asynchronous, stream-ordered */
int  rtb_dispatch(rtb_ctx* ctx, rtb_pass pass);
/* replaces Graphics::presentToCpu + wait (ref: src/rt/raytracing_interface.cpp:222-226,
 * IGNIS/api/opengl/src/graphics/gl_graphics.cpp:221-242); blocks until the copy has landed */
int  rtb_readback(rtb_ctx* ctx, rtb_target target, void* dst, size_t bytes);
/* The same copy without blocking the host, the way the reference's presentToCpu works (a PBO copy plus a fence, picked up later:
 * ref: IGNIS/api/opengl/src/graphics/gl_graphics.cpp:221-256,546-595).  The copy is ordered after everything dispatched so
 * far and runs on its own stream; following dispatches overlap it, and only the pass that overwrites `target` waits for it on
 * the device.  dst should be page-locked memory; it may be read after rtb_readback_wait (or rtb_sync).  One read-back is in
 * flight at a time: a second call first waits for the previous one. */
int  rtb_readback_async(rtb_ctx* ctx, rtb_target target, void* dst, size_t bytes);
int  rtb_readback_wait(rtb_ctx* ctx);
/* device address of a target (for zero-copy consumers such as an NCCL gather); valid until the next rtb_resize */
int  rtb_device_ptr(rtb_ctx* ctx, rtb_target target, void** out_ptr, size_t* out_bytes);
/* Multi-GPU presentation (no reference counterpart; the reference is single-GPU).  With RTB_OPT_TILE_COUNT = n > 1 each
 * context renders the 32x32-pixel screen blocks b with b % n == rank and leaves them in RTB_TGT_RGBA8_TILED.  After the
 * caller has gathered the n tiled buffers (NCCL) into one DEVICE array of n * slots_per_rank words on this context's
 * GPU, rtb_untile writes the scan-line rgba8 frame to rgba8_out_device (NULL: into RTB_TGT_RGBA8).  Stream-ordered. */
int  rtb_untile(rtb_ctx* ctx, const void* tiled_all_device, uint32_t nranks, uint32_t slots_per_rank, void* rgba8_out_device);
/* the same on a caller-owned cudaStream_t (NULL = the context's stream), so that the gather and the lay-out of frame k can run
 * under the rendering of frame k + 1; rgba8_out_device must then be the caller's own frame */
int  rtb_untile_on(rtb_ctx* ctx, const void* tiled_all_device, uint32_t nranks, uint32_t slots_per_rank, void* rgba8_out_device, void* cuda_stream);
/* Presentation to the host without a gather: this context's pixels (all of them, or its tiles with RTB_OPT_TILE_COUNT > 1) are
 * written by a kernel to their scan-line positions of a page-locked, MAPPED host frame of width*height*4 bytes (cudaHostAlloc /
 * cudaHostRegister with the mapped flag; with several ranks: one frame in shared memory registered by every rank, so that every
 * GPU uses its own PCIe link instead of rank 0's).  Stream-ordered; the frame may be read after the stream has been synchronised
 * on every rank.  tiled_src_device: NULL = this context's own pixels; otherwise a copy of its RTB_TGT_RGBA8_TILED buffer (so that
 * the next frame may overwrite the target while this one travels), which is required with a caller-owned cuda_stream (NULL = the
 * context's stream).  Replaces Graphics::presentToCpu's PBO copy (ref: IGNIS/api/opengl/src/graphics/gl_graphics.cpp:221-242). */
int  rtb_present_host(rtb_ctx* ctx, void* host_frame_rgba8, const void* tiled_src_device, void* cuda_stream);
/* replaces Graphics::wait (ref: IGNIS/api/opengl/src/graphics/gl_graphics.cpp:546-595) */
int  rtb_sync(rtb_ctx* ctx);
int  rtb_counters_get(rtb_ctx* ctx, rtb_counters* out);
/* milliseconds spent in each phase of the last RTB_PASS_FRAME (CUDA events on the context's stream): init, ray generation,
 * nearest-hit traversal, G-buffer finish, shadow-ray set-up, occlusion traversal, lighting+composite, total.  Synchronises. */
int  rtb_last_frame_ms(rtb_ctx* ctx, float out_ms[8]);

/* ---- diffuse bounces (BASELINE.json configs[3]; NO reference counterpart: the reference traces no secondary rays, its
 * "reflection" is one skybox tap, ref: res/shaders/light.glsl:205-219) ------------------------------------------------------------
 * One frame of wavefront path tracing, 1 sample per pixel: init, the reference's G-buffer (primary ray + nearest hit, as
 * RTB_PASS_RAYGEN), then per path vertex one shadow ray to lights[0] built as shadow.comp builds it, the Cook-Torrance term of
 * lighting.comp when it is not occluded, and — up to `bounces` times — a cosine-distributed bounce about the shading normal with
 * throughput *= albedo; the sky on a miss; composite.comp's tail (progressive accumulation under USE_SUPERSAMPLING, exposure,
 * rgba8).  The full definition is in csrc/rtb_path.cuh (tests hold it against a CPU statement of the same).  bounces = 0 gives the
 * direct-lighting image.  Targets written: DIR_T, UV_NORMAL, ACCUM (if supersampling), RGBA8, RGBA8_TILED, SEED. */
int  rtb_path_frame(rtb_ctx* ctx, uint32_t bounces);
typedef struct rtb_path_stats {
    uint32_t depths;                      /* bounces + 1 */
    uint32_t closest_launches, shadow_launches, kernel_launches;   /* per frame */
    uint64_t closest_rays, shadow_rays;   /* rays of the last frame: camera + bounce rays, shadow rays (this context's tiles) */
    uint64_t closest_rays_at_depth[16], shadow_rays_at_depth[16];
    float    closest_ms, shadow_ms, total_ms;   /* CUDA events on the context's stream around the traversal launches / the frame */
    float    closest_ms_at_depth[16], shadow_ms_at_depth[16];
} rtb_path_stats;
int  rtb_path_stats_get(rtb_ctx* ctx, rtb_path_stats* out);   /* synchronises */

/* Measurement aid, not on the hot path: read-only streaming bandwidth (1e9 B/s) over a buffer of `bytes` (choose <= 64 MiB so
 * that it stays L2-resident) with L2-only cached loads — the denominator of the L2 roofline. */
int  rtb_probe_l2_read_gbs(rtb_ctx* ctx, size_t bytes, double* out_gbs);

/* ---- rays-in mode (parity harness: same traversal, explicit rays) --------------------------------- */
/* rays: n * 6 floats (origin, dir) in HOST memory; prev: n object ids to exclude or NULL (= none).
 * Outputs (host): object id (0xFFFFFFFF on a miss), t (3.4028235e38 on a miss), uv (2 floats) — any may be NULL. */
int  rtb_trace_rays(rtb_ctx* ctx, const float* rays, uint64_t n, const uint32_t* prev,
                    uint32_t* object, float* t, float* uv);
/* occluded[i] = 1 when any primitive other than prev[i] is hit at a distance < max_dist[i] (NULL: 3.4028235e38) */
int  rtb_occlusion_rays(rtb_ctx* ctx, const float* rays, uint64_t n, const float* max_dist,
                        const uint32_t* prev, uint8_t* occluded);

/* ---- host-side packing (the igx:: POD constructors; see include/igx_rt.hpp for the C++ faces) ------ */
/* ref: igx/include/types/scene_object_types.hpp:98-106 (3-point ctor) and :87-96 (with normals) */
void rtb_pack_triangle(const float p[9], const float* normals9_or_null, void* out48);
/* ref: scene_object_types.hpp:166-171 / :173-179 */
void rtb_pack_light_directional(const float dir[3], const float color[3], float angular_extent, void* out32);
void rtb_pack_light_point(const float pos[3], const float color[3], float rad, float origin, float specularity, void* out32);
/* ref: scene_object_types.hpp:267-288 */
void rtb_pack_material(const float albedo[3], const float ambient[3], const float emission[3],
                       float metallic, float roughness, float transparency, void* out32);
/* CPUCamera::getView + RaytracingInterface::resize/update (ref: src/rt/structs.cpp:5-42,
 * src/rt/raytracing_interface.cpp:96-107,279-328).  Angles in radians, fov in degrees. */
void rtb_pack_camera(const float eye[3], float pitch, float yaw, float roll, float left_fov, float right_fov,
                     float ipd, uint32_t projection, uint32_t width, uint32_t height, uint32_t flags,
                     float exposure, const float skybox_color[3], void* out144);
/* Radiance .hdr -> rgba16f the way igxi-convert does (ref: igx/igxi-tool/src/igxi/convert.cpp:59-78,148-153,209-231).
 * out == NULL only queries the size. */
int  rtb_load_hdr(const char* path, uint16_t* out, uint32_t* width, uint32_t* height);

/* Frame export: an rgba8 frame as read back from RTB_TGT_RGBA8 -> 8-bit RGBA PNG at `path`.  flip_vertically != 0 writes the
 * last row first, which is what the reference does (row 0 of the frame is the bottom of the view)
 * (ref: src/rt/raytracing_interface.cpp:124-142 onRenderFinish -> igxi::Helper::toDiskExternal,
 *  igx/igxi-tool/src/igxi/convert.cpp:747-781,904-928).  Host-only; returns 0 on success. */
int  rtb_write_png(const char* path, uint32_t width, uint32_t height, const void* rgba8, int flip_vertically);

/* ---- synthetic scenes of BASELINE.json (no reference counterpart; deterministic) ------------------- */
/* n flat-shaded triangles: centre uniform in [-10,10]^3, three offsets uniform in [-0.05,0.05]^3 */
void rtb_gen_soup(uint64_t n, uint64_t seed, void* out_triangles48);
/* (grid x grid) quads of a displaced height field over [-10,10]^2, smooth normals: 2*grid*grid triangles */
void rtb_gen_heightfield(uint32_t grid, uint64_t seed, void* out_triangles48);

#ifdef __cplusplus
}
#endif
#endif
