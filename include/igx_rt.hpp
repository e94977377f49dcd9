// igx_rt.hpp — C++ host side of the drop-in: the igx:: scene types, igx::SceneGraph and the
// igx::rt:: render tasks of the reference, re-implemented above the rtb200 C ABI (include/rtb200.h).
//
// Same names, argument meaning and byte layouts as the reference so that a user of
//   igx::SceneGraph::{add, addGeometry, addNonGeometry, update, del, update(dt)}        (ref: igx/include/helpers/scene_graph.hpp:132-200)
//   igx::rt::{RaygenTask, ShadowTask, CompositeTask}::{resize, update, switchToScene, prepareCommandList}
//                                                                                       (ref: include/rt/task/*.hpp)
//   igx::rt::RaytracingInterface::{resize, update, render}                               (ref: include/rt/raytracing_interface.hpp:95-155)
// can switch over.  What changes is what prepareCommandList records: rtb_dispatch calls instead of ignis
// BindPipeline/BindDescriptors/Dispatch commands, and uploads go through rtb_upload instead of
// GPUBuffer::flush + cmd::FlushBuffer.  GUI, window, swapchain and the cloud task are out of scope
// (SURVEY.md §2 rows 4, 6, 9).
//
// Header-only; link with librtb200.so.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <string>
#include <unordered_map>
#include <vector>

#include "rtb200.h"

namespace igx {

using u8 = uint8_t; using u16 = uint16_t; using u32 = uint32_t; using u64 = uint64_t; using usz = size_t;
using f32 = float; using f64 = double;
template <class T> using List = std::vector<T>;
using String = std::string;

static constexpr f64 PI_CONST = 3.141592653589793;
static constexpr f64 TO_RAD = PI_CONST / 180;
constexpr f64 operator""_deg(long double v) { return f64(v * TO_RAD); }
constexpr f64 operator""_deg(unsigned long long v) { return f64(int64_t(v)) * TO_RAD; }

// 16-bit float with core2's conversion: truncation toward zero, values below the smallest normal collapse
// to signed zero, overflow to infinity (ref: core2/include/types/flp.hpp:124-162; known answers core2/test/test.cpp:8-29).
struct f16 {
    u16 value = 0;
    f16() = default;
    f16(f32 v) {
        u32 b; std::memcpy(&b, &v, 4);
        value = u16((b >> 31) << 15);
        const u32 mantissa = b & 0x7FFFFFu, rawExponent = (b >> 23) & 0xFFu;
        if (!mantissa && !rawExponent) return;
        const int32_t e = int32_t(rawExponent) - 127 + 15;
        if (e < 0) return;
        if (rawExponent == 0xFFu && mantissa) { value |= u16((0x1Fu << 10) | 0x3FFu); return; }
        if (e >= 30 && (e >= 31 || mantissa > (0x3FFu << 13))) { value |= u16(0x1Fu << 10); return; }
        value |= u16((u32(e) << 10) | (mantissa >> 13));
    }
    operator f32() const {
        const u32 sign = u32(value >> 15) << 31, e = (value >> 10) & 0x1Fu, m = value & 0x3FFu;
        u32 b;
        if (e == 0) { f32 f = f32(m) * 5.9604644775390625e-8f; return sign ? -f : f; }
        if (e == 31) b = sign | 0x7F800000u | (m << 13);
        else b = sign | ((e + 112u) << 23) | (m << 13);
        f32 f; std::memcpy(&f, &b, 4); return f;
    }
};

// ---- minimal vectors with core2's semantics (ref: core2/include/types/vec.hpp:153-165) ------------------
struct Vec2f32 {
    f32 x = 0, y = 0;
    Vec2f32() = default; Vec2f32(f32 x, f32 y) : x(x), y(y) {}
    f32 magnitude() const { f32 s = 0; s += x * x; s += y * y; return f32(std::sqrt(f64(s))); }
    Vec2f32 normalize() const { const f32 m = magnitude(); return {x / m, y / m}; }
    Vec2f32 operator*(f32 s) const { return {x * s, y * s}; }
    f32 aspect() const { return x / y; }
};
struct Vec2u32 { u32 x = 0, y = 0; Vec2u32() = default; Vec2u32(u32 x, u32 y) : x(x), y(y) {} bool operator==(const Vec2u32& o) const { return x == o.x && y == o.y; } bool operator!=(const Vec2u32& o) const { return !(*this == o); } };
struct Vec3f32 {
    f32 x = 0, y = 0, z = 0;
    Vec3f32() = default; Vec3f32(f32 x, f32 y, f32 z) : x(x), y(y), z(z) {}
    Vec2f32 xy() const { return {x, y}; }
    Vec3f32 operator+(const Vec3f32& o) const { return {x + o.x, y + o.y, z + o.z}; }
    Vec3f32 operator-(const Vec3f32& o) const { return {x - o.x, y - o.y, z - o.z}; }
    Vec3f32 operator*(f32 s) const { return {x * s, y * s, z * s}; }
    Vec3f32 operator/(f32 s) const { return {x / s, y / s, z / s}; }
    Vec3f32& operator+=(const Vec3f32& o) { x += o.x; y += o.y; z += o.z; return *this; }
    f32 magnitude() const { f32 s = 0; s += x * x; s += y * y; s += z * z; return f32(std::sqrt(f64(s))); }
    Vec3f32 normalize() const { return *this / magnitude(); }
    Vec3f32 cross(const Vec3f32& o) const { return {y * o.z - z * o.y, z * o.x - x * o.z, x * o.y - y * o.x}; }
};
struct Vec4f32 { f32 x = 0, y = 0, z = 0, w = 0; Vec4f32() = default; Vec4f32(f32 x, f32 y, f32 z, f32 w) : x(x), y(y), z(z), w(w) {} };
struct Mat3x3f32 { Vec3f32 xAxis, yAxis, zAxis; };
// column-major; M * v accumulates column i times v[i] starting from zero (ref: core2/include/types/mat.hpp:231-240)
struct Mat4x4f32 {
    Vec3f32 x, y, z, pos;
    Vec3f32 transform(const Vec4f32& v) const {
        Vec3f32 r;
        r += x * v.x; r += y * v.y; r += z * v.z; r += pos * v.w;
        return r;
    }
};

// ---- scene PODs (ref: igx/include/types/scene_object_types.hpp) ---------------------------------------------
enum class LightType : u16 { Directional, Spot, Point, count };
enum class ProjectionType : u32 { Default, Omnidirectional, Stereoscopic_omnidirectional_TB, Stereoscopic_TB, Stereoscopic_omnidirectional_LR, Stereoscopic_LR };
enum class CameraFlags : u32 { NONE = 0, USE_UI = 1 << 0, USE_SUPERSAMPLING = 1 << 1 };
inline CameraFlags operator|(CameraFlags a, CameraFlags b) { return CameraFlags(u32(a) | u32(b)); }
inline CameraFlags operator&(CameraFlags a, CameraFlags b) { return CameraFlags(u32(a) & u32(b)); }
inline CameraFlags operator~(CameraFlags a) { return CameraFlags(~u32(a)); }

struct Camera {   // :32-66
    Vec3f32 eye{4, 2, -2}; u32 width = 0;
    Vec3f32 p0; u32 height = 0;
    Vec3f32 p1; f32 ipd = 62;
    Vec3f32 p2; ProjectionType projectionType = ProjectionType::Default;
    Vec3f32 skyboxColor{0.25f, 0.5f, 1.f}; f32 exposure = 1.f;
    Vec3f32 p3; f32 focalDistance = 10.f;
    Vec3f32 p4; f32 aperature = 0.1f;
    Vec3f32 p5; CameraFlags flags = CameraFlags::USE_UI;
    Vec2f32 invRes; Vec2u32 tiles;
};
static_assert(sizeof(Camera) == 144, "Camera must stay 144 bytes");

inline void spheremapTransform(f16& nx, f16& ny, const Vec3f32& n) {   // :68-72
    const Vec2f32 enc = n.xy().normalize() * std::sqrt(-n.z * 0.5f + 0.5f);
    nx = enc.x; ny = enc.y;
}

struct Triangle {   // :74-112
    Vec3f32 p0; f16 n0x, n0y;
    Vec3f32 p1; f16 n1x, n1y;
    Vec3f32 p2; f16 n2x, n2y;
    Triangle() = default;
    Triangle(const Vec3f32& p0, const Vec3f32& p1, const Vec3f32& p2, const Vec3f32& n0, const Vec3f32& n1, const Vec3f32& n2) : p0(p0), p1(p1), p2(p2) {
        spheremapTransform(n0x, n0y, n0); spheremapTransform(n1x, n1y, n1); spheremapTransform(n2x, n2y, n2);
    }
    Triangle(const Vec3f32& p0, const Vec3f32& p1, const Vec3f32& p2) : p0(p0), p1(p1), p2(p2) {
        const Vec3f32 n = (p1 - p0).normalize().cross((p2 - p0).normalize());   // not re-normalised, as in the reference
        spheremapTransform(n0x, n0y, n); spheremapTransform(n1x, n1y, n); spheremapTransform(n2x, n2y, n);
    }
    Vec3f32 edge0() const { return p1 - p0; }
    Vec3f32 edge1() const { return p2 - p0; }
    Vec3f32 edge2() const { return p2 - p1; }
};
static_assert(sizeof(Triangle) == 48, "Triangle must stay 48 bytes");

struct Cube { Vec3f32 min, max; };
struct Sphere { Vec3f32 Position; f32 Radius = 1; Sphere() = default; Sphere(const Vec3f32& p, f32 r) : Position(p), Radius(r) {} };
struct Plane { Vec3f32 dir; f32 dist = 0; Plane() = default; Plane(const Vec3f32& d, f32 dist) : dir(d), dist(dist) {} };
static_assert(sizeof(Cube) == 24 && sizeof(Sphere) == 16 && sizeof(Plane) == 16, "primitive layouts");

inline Vec2u32 encodeNormal(const Vec3f32& n) {   // :131-139
    const Vec3f32 u = n.normalize();
    const Vec3f32 nn{(u.x * 0.5f + 0.5f) * 65535.0f, (u.y * 0.5f + 0.5f) * 65535.0f, (u.z * 0.5f + 0.5f) * 65535.0f};
    return Vec2u32(u32(nn.x) << 16 | u32(nn.y), u32(nn.z));
}
inline Vec3f32 decodeNormal(const Vec2u32& e) {   // :141-144
    const Vec3f32 nn = Vec3f32(f32(e.x >> 16), f32(u16(e.x)), f32(e.y)) / 65535.0f;
    return nn * 2 - Vec3f32(1, 1, 1);
}

struct Light {   // :158-265
    Vec3f32 pos; f16 rad, origin;
    Vec2u32 dir; f16 r, g, b; LightType type;
    Light() : type(LightType::Directional) {}
    Light(Vec3f32 dir, Vec3f32 color, f32 angularExtent = f32(0.533_deg)) : rad(angularExtent), dir(encodeNormal(dir)), r(color.x), g(color.y), b(color.z), type(LightType::Directional) {}
    Light(Vec3f32 pos, Vec3f32 color, f32 rad, f32 origin, f32 specularity = 1) : pos(pos), rad(rad), origin(origin), r(color.x), g(color.y), b(color.z), type(LightType::Point) {
        u32 s; std::memcpy(&s, &specularity, 4); dir = Vec2u32(s, 0);
    }
};
static_assert(sizeof(Light) == 32, "Light must stay 32 bytes");

struct Material {   // :267-290
    f16 albedoR, albedoG, albedoB; u16 metallic;
    f16 ambientR, ambientG, ambientB; u16 roughness;
    f16 emissionR, emissionG, emissionB, pad2;
    f32 transparency; u32 materialInfo{};
    Material() : metallic(0), roughness(0), transparency(0) {}
    Material(Vec3f32 albedo, Vec3f32 ambient, Vec3f32 emission, f32 metallic, f32 roughness, f32 transparency)
        : albedoR(albedo.x), albedoG(albedo.y), albedoB(albedo.z), metallic(u16(metallic * 65535)),
          ambientR(ambient.x), ambientG(ambient.y), ambientB(ambient.z), roughness(u16(roughness * 65535)),
          emissionR(emission.x), emissionG(emission.y), emissionB(emission.z), transparency(transparency) {}
};
static_assert(sizeof(Material) == 32, "Material must stay 32 bytes");

// ---- SceneGraph (ref: igx/include/helpers/scene_graph.hpp, igx/src/helpers/scene_graph.cpp) -------------------
enum class SceneObjectType : u8 { LIGHT, MATERIAL, TRIANGLE, SPHERE, CUBE, PLANE, COUNT, FIRST = LIGHT };
template <class T> struct TSceneObjectType { static constexpr SceneObjectType type = SceneObjectType::COUNT; static constexpr bool isGeometry = false; };
template <> struct TSceneObjectType<Triangle> { static constexpr SceneObjectType type = SceneObjectType::TRIANGLE; static constexpr bool isGeometry = true; };
template <> struct TSceneObjectType<Light> { static constexpr SceneObjectType type = SceneObjectType::LIGHT; static constexpr bool isGeometry = false; };
template <> struct TSceneObjectType<Material> { static constexpr SceneObjectType type = SceneObjectType::MATERIAL; static constexpr bool isGeometry = false; };
template <> struct TSceneObjectType<Cube> { static constexpr SceneObjectType type = SceneObjectType::CUBE; static constexpr bool isGeometry = true; };
template <> struct TSceneObjectType<Sphere> { static constexpr SceneObjectType type = SceneObjectType::SPHERE; static constexpr bool isGeometry = true; };
template <> struct TSceneObjectType<Plane> { static constexpr SceneObjectType type = SceneObjectType::PLANE; static constexpr bool isGeometry = true; };

union SceneGraphInfo {   // scene_graph.hpp:61-74
    u32 fields[9]{};
    struct { u32 objectCount[6]; u32 lightsCount[3]; };
    struct { u32 lightCount, materialCount, triangleCount, sphereCount, cubeCount, planeCount, directionalLightCount, spotLightCount, pointLightCount; };
};
static_assert(sizeof(SceneGraphInfo) == 36, "SceneGraphInfo must stay 36 bytes");

// The "Graphics + FactoryContainer" of the rebuild: owns the rtb200 context of one GPU.
class Device {
    rtb_ctx* ctx = nullptr;
public:
    // capacities are explicit (the reference hard-codes 65536 / 32768 / 16384 / 256)
    explicit Device(int cudaDevice = 0, u32 maxTriangles = 65536, u32 maxLights = 65536, u32 maxMaterials = 65536,
                    u32 maxCubes = 32768, u32 maxSpheres = 16384, u32 maxPlanes = 256) {
        rtb_limits l{maxTriangles, maxSpheres, maxCubes, maxPlanes, maxLights, maxMaterials};
        limits_ = l;
        if (rtb_create(&ctx, cudaDevice, &l) != RTB_OK) { error_ = rtb_last_error(nullptr); ctx = nullptr; }
    }
    ~Device() { if (ctx) rtb_destroy(ctx); }
    Device(const Device&) = delete; Device& operator=(const Device&) = delete;
    bool valid() const { return ctx != nullptr; }
    rtb_ctx* get() const { return ctx; }
    const rtb_limits& limits() const { return limits_; }
    String error() const { return ctx ? String(rtb_last_error(ctx)) : error_; }
private:
    rtb_limits limits_{}; String error_;
};

// What the reference calls a CommandList: the ordered list of GPU actions one frame replays
// (ref: IGNIS/include/graphics/command/command_list.hpp).  Here a command is a call into the C ABI.
class CommandList {
public:
    // what a recorded command is, as far as replay needs to know: the five dispatches of the reference's frame and the shadow
    // properties flush are recognised, everything else (uploads, builds, a user's pass) is opaque
    enum class Kind : u8 { OTHER, INIT, RAYGEN, SHADOW_PROPS, SHADOW, LIGHTING, COMPOSITE };
private:
    struct Cmd { std::function<int(rtb_ctx*)> run; Kind kind; };
    List<Cmd> cmds;
public:
    // The reference's frame is INIT, RAYGEN, FlushBuffer(shadow properties), SHADOW, LIGHTING, COMPOSITE, recorded by three tasks
    // (composite_task.cpp:249-277, raygen_task.cpp:88-94, shadow_task.cpp:189-210).  Replayed as ONE RTB_PASS_FRAME when the six stand
    // together — same kernels' arithmetic, same targets (tests/test_gpu_parity.py: per-pass dispatches == RTB_PASS_FRAME), but the
    // library may then fuse lighting + composite, record the frame as CUDA graphs and overlap consecutive frames.  A pass injected
    // between them (addPrepass / addPostpass record before / after the run, never inside) keeps the per-pass dispatches.
    bool fuseFrame = true;
    void add(std::function<int(rtb_ctx*)> c, Kind k = Kind::OTHER) { cmds.push_back(Cmd{std::move(c), k}); }
    void addDispatch(rtb_pass pass) {
        static constexpr Kind kinds[5] = {Kind::INIT, Kind::RAYGEN, Kind::SHADOW, Kind::LIGHTING, Kind::COMPOSITE};
        add([pass](rtb_ctx* c) { return rtb_dispatch(c, pass); }, pass <= RTB_PASS_COMPOSITE ? kinds[pass] : Kind::OTHER);
    }
    void clear() { cmds.clear(); }
    bool empty() const { return cmds.empty(); }
    usz size() const { return cmds.size(); }
    // Graphics::execute: replay in order; the first failing command stops the list and is returned
    int execute(rtb_ctx* ctx) const {
        static constexpr Kind frame[6] = {Kind::INIT, Kind::RAYGEN, Kind::SHADOW_PROPS, Kind::SHADOW, Kind::LIGHTING, Kind::COMPOSITE};
        for (usz i = 0; i < cmds.size();) {
            bool whole = fuseFrame && i + 6 <= cmds.size();
            for (usz k = 0; whole && k < 6; ++k) whole = cmds[i + k].kind == frame[k];
            if (whole) {
                int rc = cmds[i + 2].run(ctx);                         // the properties flush (a no-op unless the sample count changed) ...
                if (!rc) rc = rtb_dispatch(ctx, RTB_PASS_FRAME);       // ... then the five dispatches as one
                if (rc) return rc;
                i += 6;
                continue;
            }
            const int rc = cmds[i].run(ctx);
            if (rc) return rc;
            ++i;
        }
        return 0;
    }
};

class SceneGraph {
public:
    struct Entry { u32 index, material; SceneObjectType type; };
    enum class Flags : u32 { NONE = 0 };

private:
    // One pool per object type.  Slot i holds the record at bytes[i * stride]; owner[i] is its handle (0 = free slot).
    // `device` mirrors what the device buffer holds, so that changed records travel as contiguous runs.
    struct Object {
        List<u8> cpuData, gpuData;      // working copy / what the device holds
        List<bool> markedForUpdate;     // slot changed since the last upload
        List<u64> toIndex;              // owner handle per slot
        u32 holes = 0;                  // free slots below the pool's count (add() only searches when there are any)
    };
    static constexpr usz strides[6] = {sizeof(Light), sizeof(Material), sizeof(Triangle), sizeof(Sphere), sizeof(Cube), sizeof(Plane)};
    static constexpr rtb_buffer bufferIds[6] = {RTB_BUF_LIGHTS, RTB_BUF_MATERIALS, RTB_BUF_TRIANGLES, RTB_BUF_SPHERES, RTB_BUF_CUBES, RTB_BUF_PLANES};

    Device& device;
    Object objects[6];
    std::unordered_map<u64, Entry> entries;
    u64 counter = 0;
    SceneGraphInfo info{}, limits{};
    List<u32> materialByObject;
    bool geometryDirty = true;      // triangles changed since the last acceleration-structure build or refit
    bool topologyDirty = true;      // ... and their number or order changed too: a refit is not enough
    bool builtOnce = false;
    String sceneName;
    List<u16> skyboxPixels; u32 skyW = 0, skyH = 0; bool skyboxDirty = false;
    int lastError = 0;

public:
    SceneGraph(const SceneGraph&) = delete; SceneGraph& operator=(const SceneGraph&) = delete;

    // skyboxName: path of a Radiance .hdr (the reference loads "./res/textures/qwantani_4k.hdr" from its virtual file
    // system, test/scene/niels_scene.cpp:6); empty = no skybox, camera.skyboxColor is used.
    SceneGraph(Device& device, const String& sceneName, const String& skyboxName, Flags = Flags::NONE) : device(device), sceneName(sceneName) {
        const rtb_limits& l = device.limits();
        limits.lightCount = l.max_lights; limits.materialCount = l.max_materials; limits.triangleCount = l.max_triangles;
        limits.sphereCount = l.max_spheres; limits.cubeCount = l.max_cubes; limits.planeCount = l.max_planes;
        for (int t = 0; t < 6; ++t) {
            objects[t].markedForUpdate.assign(limits.objectCount[t], false);
            objects[t].toIndex.assign(limits.objectCount[t], 0);
        }
        materialByObject.assign(usz(limits.triangleCount) + limits.sphereCount + limits.cubeCount + limits.planeCount, 0);
        if (!skyboxName.empty()) {
            if (rtb_load_hdr(skyboxName.c_str(), nullptr, &skyW, &skyH) == 0) {
                skyboxPixels.resize(usz(skyW) * skyH * 4);
                if (rtb_load_hdr(skyboxName.c_str(), skyboxPixels.data(), &skyW, &skyH) == 0) skyboxDirty = true;
                else { skyboxPixels.clear(); skyW = skyH = 0; }
            }
        }
    }
    virtual ~SceneGraph() = default;

    void del(const List<u64>& ids) {
        for (u64 i : ids) {
            auto it = entries.find(i);
            if (it == entries.end()) continue;
            const u8 t = u8(it->second.type);
            objects[t].toIndex[it->second.index] = 0;
            objects[t].markedForUpdate[it->second.index] = false;
            ++objects[t].holes;
            if (it->second.type == SceneObjectType::TRIANGLE) geometryDirty = topologyDirty = true;
            entries.erase(it);
        }
    }

    template <class T> u64 addNonGeometry(const T& object) {
        static_assert(TSceneObjectType<T>::type != SceneObjectType::COUNT && !TSceneObjectType<T>::isGeometry, "expects Light or Material");
        return addInternal(TSceneObjectType<T>::type, &object, sizeof(T), 0);
    }
    template <class T> u64 addGeometry(const T& object, const u32 material) {
        static_assert(TSceneObjectType<T>::isGeometry, "expects Triangle, Sphere, Cube or Plane");
        return addInternal(TSceneObjectType<T>::type, &object, sizeof(T), material);
    }
    // add(light, material, triangle, 3u, sphere, 0u, ...): geometry is followed by its material index
    template <class T, class T2, class... Args> void add(const T& a, const T2& b, const Args&... rest) {
        if constexpr (TSceneObjectType<T>::isGeometry) {
            static_assert(std::is_same_v<T2, u32>, "geometry requires a u32 material right after it");
            addGeometry(a, b);
            if constexpr (sizeof...(Args) > 0) add(rest...);
        } else { addNonGeometry(a); add(b, rest...); }
    }
    template <class T> void add(const T& object) { addNonGeometry(object); }

    auto find(u64 index) const { return entries.find(index); }
    bool exists(u64 index) const { return find(index) != entries.end(); }

    // returns false for an unknown handle or a handle of another type
    template <class T> bool update(u64 index, const T& object) {
        constexpr SceneObjectType type = TSceneObjectType<T>::type;
        static_assert(type != SceneObjectType::COUNT, "expects a scene object");
        auto it = entries.find(index);
        if (it == entries.end() || it->second.type != type) return false;
        Object& obj = objects[u8(type)];
        u8* target = obj.cpuData.data() + usz(it->second.index) * sizeof(T);
        if (std::memcmp(&object, target, sizeof(T)) == 0) return true;
        obj.markedForUpdate[it->second.index] = true;
        std::memcpy(target, &object, sizeof(T));
        if (type == SceneObjectType::TRIANGLE) geometryDirty = true;
        return true;
    }

    // Behaviour of the reference's per-frame update (scene_graph.cpp:267-323, 378-522), restated: every pool is squeezed (holes
    // out, order kept; lights additionally grouped directional < spot < point, order kept inside a group), records that
    // changed or moved are uploaded as contiguous runs, and the material-index table lists the geometry in the order the
    // shaders number it: triangles, spheres, cubes, planes.
    virtual void update(f64) {
        lastError = 0;
        for (u8 t = 0; t < 6; ++t) {
            squeeze(SceneObjectType(t));
            Object& obj = objects[t];
            const usz stride = strides[t];
            const u32 n = info.objectCount[t];
            for (u32 i = 0; i < n;) {   // one upload per dirty run
                if (!obj.markedForUpdate[i]) { ++i; continue; }
                u32 j = i;
                while (j < n && obj.markedForUpdate[j]) obj.markedForUpdate[j++] = false;
                std::memcpy(obj.gpuData.data() + stride * i, obj.cpuData.data() + stride * i, usz(j - i) * stride);
                note(rtb_upload(device.get(), bufferIds[t], stride * i, usz(j - i) * stride, obj.gpuData.data() + stride * i));
                i = j;
            }
        }
        syncMaterialIndices();
        note(rtb_upload(device.get(), RTB_BUF_SCENE_INFO, 0, sizeof(info), &info));
    }

    // the copy commands of the frame (scene_graph.cpp:253-265); the skybox is flushed when it changed, and the
    // acceleration structure — which the reference does not have — follows the triangles: update<Triangle>() (the
    // reference's per-frame mutation path, niels_scene.cpp:61-70) costs a device refit, add / del / compaction a rebuild
    void fillCommandList(CommandList* cl) {
        cl->add([this](rtb_ctx* c) {
            if (skyboxDirty) { skyboxDirty = false; const int rc = rtb_upload_skybox(c, skyW, skyH, skyboxPixels.empty() ? nullptr : skyboxPixels.data()); if (rc) return rc; }
            if (geometryDirty) {
                int rc;
                if (topologyDirty) {
                    // the first tree is built by the host builder (best SAH cost); a scene whose triangles come and go afterwards
                    // (add / del / compaction, scene_graph.cpp:343-376,378-522) rebuilds on the device: milliseconds, not a stall
                    rtb_set_option(c, RTB_OPT_ACCEL_BUILDER, (builtOnce && rebuildOnDevice) ? 1u : 0u);
                    rc = rtb_build_accel(c, accelMode);
                } else
                    rc = rtb_refit_accel(c);
                if (rc) return rc;   // stays dirty: the next frame tries again
                builtOnce = true;
                geometryDirty = topologyDirty = false;
            }
            return 0;
        });
    }

    const SceneGraphInfo& getInfo() const { return info; }
    const SceneGraphInfo& getLimits() const { return limits; }
    // CPU mirrors (what the reference keeps in Object::cpuData / materialByObject)
    const List<u32>& getMaterialIndices() const { return materialByObject; }
    const u8* getCpuData(SceneObjectType t) const { return objects[u8(t)].cpuData.data(); }
    Device& getDevice() const { return device; }
    int error() const { return lastError; }
    rtb_accel_mode accelMode = RTB_ACCEL_BVH;
    bool rebuildOnDevice = true;    // RTB_OPT_ACCEL_BUILDER = 1 for every build after the first

private:
    void note(int rc) { if (rc && !lastError) lastError = rc; }

    // A new object takes the lowest free slot of its pool, or the next one; 0 when the pool is at capacity.
    // (behaviour of scene_graph.cpp:343-376)
    u64 addInternal(SceneObjectType t, const void* v, usz siz, u32 mat) {
        Object& pool = objects[u8(t)];
        u32& used = info.objectCount[u8(t)];
        u32 slot = used;
        if (pool.holes) slot = u32(std::find(pool.toIndex.begin(), pool.toIndex.begin() + used, u64(0)) - pool.toIndex.begin());
        if (slot == used) {
            if (used == limits.objectCount[u8(t)]) return 0;
            ++used;
        } else
            --pool.holes;
        u64 handle = counter;
        do { ++handle; } while (!handle || entries.count(handle));
        counter = handle;
        const usz end = usz(slot + 1) * siz;
        if (pool.cpuData.size() < end) { const usz want = std::max(end, pool.cpuData.size() * 2); pool.cpuData.resize(want); pool.gpuData.resize(want); }
        std::memcpy(pool.cpuData.data() + siz * slot, v, siz);
        pool.toIndex[slot] = handle;
        pool.markedForUpdate[slot] = true;
        entries[handle] = Entry{slot, mat, t};
        if (t == SceneObjectType::TRIANGLE) geometryDirty = topologyDirty = true;
        return handle;
    }

    static u32 lightGroup(const u8* record) { const u16 ty = u16(reinterpret_cast<const Light*>(record)->type); return ty < 3 ? ty : 2u; }

    // Squeeze one pool: the live slots keep their relative order (lights: grouped by type first), move to the front, and every
    // record whose slot changed is marked for upload.  One pass builds the new order, one pass moves.
    void squeeze(SceneObjectType type) {
        Object& pool = objects[u8(type)];
        const usz stride = strides[u8(type)];
        u32& used = info.objectCount[u8(type)];
        List<u32> order;   // order[newSlot] = oldSlot
        order.reserve(used);
        for (u32 i = 0; i < used; ++i) if (pool.toIndex[i]) order.push_back(i);
        if (type == SceneObjectType::LIGHT) {
            const u8* rec = pool.cpuData.data();
            std::stable_sort(order.begin(), order.end(), [&](u32 a, u32 b) { return lightGroup(rec + usz(a) * stride) < lightGroup(rec + usz(b) * stride); });
            u32 groups[3] = {0, 0, 0};
            for (u32 o : order) ++groups[lightGroup(rec + usz(o) * stride)];
            for (int k = 0; k < 3; ++k) info.lightsCount[k] = groups[k];
        }
        bool identity = order.size() == used;
        for (u32 p = 0; identity && p < order.size(); ++p) identity = order[p] == p;
        if (identity) return;
        // move: the device mirror doubles as scratch (it is rewritten run by run from cpuData afterwards)
        List<u64> owners(pool.toIndex.size(), 0);
        List<bool> marks(pool.markedForUpdate.size(), false);
        u8* scratch = pool.gpuData.data();
        for (u32 p = 0; p < order.size(); ++p) {
            const u32 q = order[p];
            std::memcpy(scratch + usz(p) * stride, pool.cpuData.data() + usz(q) * stride, stride);
            owners[p] = pool.toIndex[q];
            marks[p] = pool.markedForUpdate[q] || p != q;
            entries[owners[p]].index = p;
        }
        std::memcpy(pool.cpuData.data(), scratch, order.size() * stride);
        pool.toIndex.swap(owners);
        pool.markedForUpdate.swap(marks);
        pool.holes = 0;
        used = u32(order.size());
        if (type == SceneObjectType::TRIANGLE) geometryDirty = topologyDirty = true;
    }

    // materialIndices[global object id] in the shaders' numbering (triangles, then spheres, cubes, planes); changed entries
    // travel as runs (behaviour of scene_graph.cpp:464-489)
    void syncMaterialIndices() {
        u32 g = 0;
        if (materialSent.size() < materialByObject.size()) materialSent.resize(materialByObject.size(), false);
        auto flushRun = [&](u32 first, u32 last) { if (last > first) note(rtb_upload(device.get(), RTB_BUF_MATERIAL_INDICES, usz(first) * 4, usz(last - first) * 4, materialByObject.data() + first)); };
        u32 runStart = 0;
        bool inRun = false;
        for (SceneObjectType t : {SceneObjectType::TRIANGLE, SceneObjectType::SPHERE, SceneObjectType::CUBE, SceneObjectType::PLANE}) {
            const Object& pool = objects[u8(t)];
            for (u32 i = 0; i < info.objectCount[u8(t)]; ++i, ++g) {
                const u32 want = entries[pool.toIndex[i]].material;
                const bool changed = materialByObject[g] != want || !materialSent[g];
                if (changed) { materialByObject[g] = want; materialSent[g] = true; if (!inRun) { inRun = true; runStart = g; } }
                else if (inRun) { flushRun(runStart, g); inRun = false; }
            }
        }
        if (inRun) flushRun(runStart, g);
    }

    List<bool> materialSent;
};

// ---- RenderTask hierarchy (ref: igx/include/helpers/render_task.hpp:21-65) -------------------------------------
enum class RenderMode : u8 { MQ };
class RenderTask {
protected:
    Device& device;
    Vec2u32 size_;
    bool dirty = true;
public:
    explicit RenderTask(Device& d) : device(d) {}
    virtual ~RenderTask() = default;
    virtual void prepareCommandList(CommandList* cl) = 0;
    virtual void update(f64 dt) = 0;
    virtual void resize(const Vec2u32& size) { size_ = size; dirty = true; }
    virtual void switchToScene(SceneGraph*) { dirty = true; }
    virtual void prepareMode(RenderMode) {}
    virtual bool needsCommandUpdate() const { return dirty; }
    const Vec2u32& size() const { return size_; }
    void markClean() { dirty = false; }
};

}  // namespace igx

namespace igx::rt {

struct Seed { f32 randomX = 0, randomY = 0, cpuOffsetX = 0, cpuOffsetY = 0; u32 sampleCount = 0, sampleOffset = 0; };   // include/rt/structs.hpp:8-14
static_assert(sizeof(Seed) == 24, "Seed must stay 24 bytes");

struct CPUCamera : public Camera {   // include/rt/structs.hpp:16-39, src/rt/structs.cpp:5-42
    f32 pitch = 0, yaw = 0, roll = 0;
    f32 speed = 5, leftFov = 70, rightFov = 70;
    Mat3x3f32 getRot() const {
        const f32 a = roll, b = yaw, g = pitch, ca = std::cos(a), cb = std::cos(b), cg = std::cos(g), sa = std::sin(a), sb = std::sin(b), sg = std::sin(g),
                  sbsg = sb * sg, sbcg = sb * cg;
        return Mat3x3f32{Vec3f32{ca * cb, ca * sbsg - sa * cg, ca * sbcg + sa * sg}, Vec3f32{sa * cb, sa * sbsg + ca * cg, sa * sbcg - ca * sg}, Vec3f32{-sb, cb * sg, cb * cg}};
    }
    // eyeOffset 0: centre, -1: left eye, 1: right eye
    Mat4x4f32 getView(f32 eyeOffset) const {
        const Mat3x3f32 rot = getRot();
        Mat4x4f32 res;
        res.x = rot.xAxis; res.y = rot.yAxis; res.z = rot.zAxis;
        res.pos = eye + rot.xAxis * (ipd * 5e-4f * eyeOffset);
        return res;
    }
    // the camera part of RaytracingInterface::resize (src/rt/raytracing_interface.cpp:96-107)
    void setSize(const Vec2u32& size) {
        width = size.x; height = size.y;
        invRes = Vec2f32(1.f / f32(size.x), 1.f / f32(size.y));
        tiles = Vec2u32(size.x / 16, size.y / 16);
    }
    // the camera part of RaytracingInterface::update (src/rt/raytracing_interface.cpp:286-325): screen-plane corners
    void updatePlanes() {
        const bool isStereo = projectionType == ProjectionType::Stereoscopic_TB || projectionType == ProjectionType::Stereoscopic_LR;
        const Mat4x4f32 vLeft = getView(isStereo ? -1.f : 0.f);
        if (projectionType == ProjectionType::Omnidirectional || projectionType == ProjectionType::Stereoscopic_omnidirectional_LR ||
            projectionType == ProjectionType::Stereoscopic_omnidirectional_TB)
            return;
        Vec2f32 res{f32(width), f32(height)};
        if (isStereo) { if (projectionType == ProjectionType::Stereoscopic_LR) res.x /= 2; else res.y /= 2; }
        const f32 aspect = res.aspect();
        const f32 nearPlaneLeft = f32(std::tan(leftFov * 0.5_deg));
        p0 = vLeft.transform(Vec4f32(-aspect, 1, -nearPlaneLeft, 1));
        p1 = vLeft.transform(Vec4f32(aspect, 1, -nearPlaneLeft, 1));
        p2 = vLeft.transform(Vec4f32(-aspect, -1, -nearPlaneLeft, 1));
        if (isStereo) {
            const Mat4x4f32 vRight = getView(1);
            const f32 nearPlaneRight = f32(std::tan(rightFov * 0.5_deg));
            p3 = vRight.transform(Vec4f32(-aspect, 1, -nearPlaneRight, 1));
            p4 = vRight.transform(Vec4f32(aspect, 1, -nearPlaneRight, 1));
            p5 = vRight.transform(Vec4f32(-aspect, -1, -nearPlaneRight, 1));
        }
    }
};

// RaygenTask: owns dirT + uvObjectNormal, dispatches raygen (ref: src/rt/task/raygen_task.cpp:9-94)
class RaygenTask : public RenderTask {
public:
    explicit RaygenTask(Device& d) : RenderTask(d) {}
    void prepareCommandList(CommandList* cl) override { cl->addDispatch(RTB_PASS_RAYGEN); dirty = false; }
    void update(f64) override {}
};

struct ShadowProperties { u32 Shadow_samples = 2; };   // include/rt/task/shadow_task.hpp:12-18

// ShadowTask: shadow + lighting dispatches, owns the shadow mask and the lighting texture (ref: src/rt/task/shadow_task.cpp:10-215)
class ShadowTask : public RenderTask {
    u32 cachedSamples = 0;
public:
    ShadowProperties properties;
    explicit ShadowTask(Device& d) : RenderTask(d) {}
    bool needsCommandUpdate() const override { return dirty || cachedSamples != properties.Shadow_samples; }
    void prepareCommandList(CommandList* cl) override {
        cachedSamples = properties.Shadow_samples;
        const u32 samples = cachedSamples;
        cl->add([samples](rtb_ctx* c) { return rtb_upload(c, RTB_BUF_SHADOW_PROPS, 0, 4, &samples); }, CommandList::Kind::SHADOW_PROPS);   // FlushBuffer(shadowProperties)
        cl->addDispatch(RTB_PASS_SHADOW);
        cl->addDispatch(RTB_PASS_LIGHTING);
        dirty = false;
    }
    void update(f64) override {}
    u32 samples() const { return properties.Shadow_samples; }
};

// CompositeTask: seed buffer, init dispatch, child tasks, composite dispatch (ref: src/rt/task/composite_task.cpp:14-277).
// The cloud task of the reference records nothing (src/rt/task/cloud/cloud_task.cpp:138-148) and is not reproduced.
class CompositeTask : public RenderTask {
    Seed seed;
    RaygenTask raygen;
    ShadowTask shadow;
    std::function<void(f32&, f32&)> offsetSource;   // the reference draws cpuOffsetX/Y from oic::Random in [-1000, 1000)
    u64 lcg = 0x9E3779B97F4A7C15ull;
public:
    explicit CompositeTask(Device& d) : RenderTask(d), raygen(d), shadow(d) {}
    RaygenTask& getRaygen() { return raygen; }
    ShadowTask& getShadow() { return shadow; }
    const Seed& getSeed() const { return seed; }
    // deterministic runs (tests, goldens) install their own source; default is a private LCG
    void setOffsetSource(std::function<void(f32&, f32&)> f) { offsetSource = std::move(f); }
    bool needsCommandUpdate() const override { return dirty || raygen.needsCommandUpdate() || shadow.needsCommandUpdate(); }
    void resize(const Vec2u32& size) override {
        RenderTask::resize(size); raygen.resize(size); shadow.resize(size);
    }
    void switchToScene(SceneGraph* sg) override { RenderTask::switchToScene(sg); raygen.switchToScene(sg); shadow.switchToScene(sg); }
    // composite_task.cpp:235-247: restart accumulation and re-draw the CPU offsets
    void update(f64 dt) override {
        seed.sampleCount = 0;
        if (offsetSource) offsetSource(seed.cpuOffsetX, seed.cpuOffsetY);
        else {
            auto next = [this]() { lcg = lcg * 6364136223846793005ull + 1442695040888963407ull; return f32(f64(lcg >> 11) / f64(1ull << 53) * 2000.0 - 1000.0); };
            seed.cpuOffsetX = next(); seed.cpuOffsetY = next();
        }
        // composite_task.cpp:243: seedBuffer->flush(0, offsetof(Seed, sampleOffset)) — 20 bytes; sampleOffset is owned by the GPU
        // (init.comp increments it) and keeps counting across updates
        rtb_upload(device.get(), RTB_BUF_SEED, 0, offsetof(Seed, sampleOffset), &seed);
        raygen.update(dt); shadow.update(dt);
    }
    // composite_task.cpp:249-277: FlushBuffer(seed), init, children, composite
    void prepareCommandList(CommandList* cl) override {
        cl->addDispatch(RTB_PASS_INIT);
        raygen.prepareCommandList(cl);
        shadow.prepareCommandList(cl);
        cl->addDispatch(RTB_PASS_COMPOSITE);
        dirty = false;
    }
};

// Headless RaytracingInterface: camera buffer + CompositeTask + pre/post passes, one command list replayed per frame
// (ref: src/rt/raytracing_interface.cpp:19-54,96-120,144-179,194-255,258-339).  Window, swapchain, input and GUI are out of scope.
// Output presets and export settings (include/rt/raytracing_interface.hpp:20-91).  UI-only members (fps readout, the
// Nuklear button) are not modelled; exportToPNG() arms the next render() exactly like the reference's button does.
enum class Resolution : u8 { CUSTOM, SD, HD, FHD, QHD, UHD_4K, UHD_8K };
static constexpr u16 pixelsByResolution[7][2] = {{0, 0}, {720, 480}, {1280, 720}, {1920, 1080}, {2560, 1440}, {3840, 2160}, {7680, 4320}};

struct RaytracingProperties {
    String targetOutput = "./output/0";       // ".png" is appended (igxi::Helper::toDiskExternal, convert.cpp:877-886,925)
    u16 targetSizeX = 7680, targetSizeY = 4320;
    u16 targetSamples = 128;                  // ui::Slider<u16, 1, 4096>
    Resolution res = Resolution::UHD_8K;
    bool shouldOutputNextFrame = false, isPortrait = false;
    void exportToPNG() { shouldOutputNextFrame = true; }
    void setResolution(Resolution r) { res = r; if (r != Resolution::CUSTOM) { targetSizeX = pixelsByResolution[u8(r)][0]; targetSizeY = pixelsByResolution[u8(r)][1]; } }
    Vec2u32 getRes() const { return isPortrait ? Vec2u32(targetSizeY, targetSizeX) : Vec2u32(targetSizeX, targetSizeY); }
};

class RaytracingInterface {
    Device& device;
    SceneGraph* sceneGraph;
    CompositeTask compositeTask;
    List<RenderTask*> prePasses, postPasses;
    CommandList cl;
    Vec2u32 res;
    int lastError = 0;
public:
    CPUCamera camera;
    u32 targetSamples = 1;          // replays of the command list per interactive render() (1 in the reference)
    RaytracingProperties properties;
    String lastExport;              // path written by the last export ("" when it failed)

    RaytracingInterface(Device& d, SceneGraph* sg) : device(d), sceneGraph(sg), compositeTask(d) {
        camera.flags = CameraFlags::NONE;   // headless: no UI blend (the export path clears USE_UI the same way, raytracing_interface.cpp:203-204)
        compositeTask.switchToScene(sg);
    }
    ~RaytracingInterface() { for (auto* p : prePasses) delete p; for (auto* p : postPasses) delete p; }
    void addPrepass(RenderTask* t) { prePasses.push_back(t); }
    void addPostpass(RenderTask* t) { postPasses.push_back(t); }
    CompositeTask& getCompositeTask() { return compositeTask; }
    int error() const { return lastError; }

    void resize(const Vec2u32& size) {
        rtb_sync(device.get());
        cl.clear();
        res = size;
        camera.setSize(size);
        note(rtb_resize(device.get(), size.x, size.y, compositeTask.getShadow().samples()));
        for (auto* p : prePasses) p->resize(size);
        compositeTask.resize(size);
        for (auto* p : postPasses) p->resize(size);
    }

    void update(f64 dt) {
        camera.updatePlanes();
        note(rtb_upload(device.get(), RTB_BUF_CAMERA, 0, sizeof(Camera), static_cast<const Camera*>(&camera)));
        sceneGraph->update(dt);
        note(sceneGraph->error());
        for (auto* p : prePasses) p->update(dt);
        compositeTask.update(dt);
        for (auto* p : postPasses) p->update(dt);
    }

    // re-record when a task is dirty, then replay `targetSamples` times (progressive accumulation when > 1).
    // With properties.shouldOutputNextFrame set, first renders the export frame the way the reference does
    // (raytracing_interface.cpp:196-242): resize to the target size, no UI, USE_SUPERSAMPLING when more than one sample,
    // update(0), the command list replayed targetSamples times (each replay K0 advances the sample counter and the
    // composite pass accumulates), read the rgba8 frame back, write <targetOutput>.png, restore size and flags.
    void render() {
        if (properties.shouldOutputNextFrame) exportFrame();
        renderOnce(targetSamples ? targetSamples : 1);
    }

    bool exportFrame() {
        const Vec2u32 oldRes = res;
        const CameraFlags oldFlags = camera.flags;
        resize(properties.getRes());
        u32 flags = u32(camera.flags) & ~u32(CameraFlags::USE_UI);
        if (properties.targetSamples > 1) flags |= u32(CameraFlags::USE_SUPERSAMPLING);
        camera.flags = CameraFlags(flags);
        update(0);
        renderOnce(properties.targetSamples ? properties.targetSamples : 1);
        List<u32> px;
        bool ok = readPixels(px);
        lastExport = properties.targetOutput + ".png";
        if (ok) ok = rtb_write_png(lastExport.c_str(), res.x, res.y, px.data(), 1) == 0;
        if (!ok) { lastExport.clear(); note(RTB_ERR_ARG); }
        camera.flags = oldFlags;
        properties.shouldOutputNextFrame = false;
        if (oldRes.x && oldRes.y) { resize(oldRes); update(0); }
        return ok;
    }

private:
    void renderOnce(u32 replays) {
        bool record = cl.empty() || compositeTask.needsCommandUpdate();
        for (auto* p : prePasses) record |= p->needsCommandUpdate();
        for (auto* p : postPasses) record |= p->needsCommandUpdate();
        if (record) {
            // Shadow_samples is baked into the shadow-mask size (ShadowTask::resize)
            note(rtb_resize(device.get(), res.x, res.y, compositeTask.getShadow().samples()));
            cl.clear();
            sceneGraph->fillCommandList(&cl);
            for (auto* p : prePasses) p->prepareCommandList(&cl);
            compositeTask.prepareCommandList(&cl);
            for (auto* p : postPasses) p->prepareCommandList(&cl);
        }
        for (u32 i = 0; i < replays; ++i) note(cl.execute(device.get()));
    }
public:

    // presentToCpu: the rgba8 frame, row 0 first (which is the bottom of the view)
    bool readPixels(List<u32>& out) {
        out.resize(usz(res.x) * res.y);
        const int rc = rtb_readback(device.get(), RTB_TGT_RGBA8, out.data(), out.size() * 4);
        note(rc);
        return rc == 0;
    }
    // the same without blocking (presentToCpu + fence): dst must stay valid until waitPixels(); rendering may go on meanwhile
    bool readPixelsAsync(u32* dst) { const int rc = rtb_readback_async(device.get(), RTB_TGT_RGBA8, dst, usz(res.x) * res.y * 4); note(rc); return rc == 0; }
    bool waitPixels() { const int rc = rtb_readback_wait(device.get()); note(rc); return rc == 0; }
private:
    void note(int rc) { if (rc && !lastError) lastError = rc; }
};

}  // namespace igx::rt
