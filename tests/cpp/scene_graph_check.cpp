// scene_graph_check.cpp — host logic of igx::SceneGraph / igx::rt::* in include/igx_rt.hpp, runnable without a GPU
// (uploads through the C ABI then fail with an error code, which the facade records; the bookkeeping still runs).
// With a GPU (argv[1] == "render") it also renders NielsScene through RaytracingInterface and prints a checksum.
#include <cstdio>
#include <cstring>

#include "igx_rt.hpp"

using namespace igx;
using namespace igx::rt;

#define CHECK(c) do { if (!(c)) { std::printf("FAIL line %d: %s\n", __LINE__, #c); return 1; } } while (0)

static void nielsScene(SceneGraph& sg, u64 dyn[3]) {
    // test/scene/niels_scene.cpp:10-56, same order of add() calls
    sg.add(Material({1, 0.5f, 1}, {0.05f, 0.01f, 0.05f}, {0, 0, 0}, 0, 1, 1), Material({0, 1, 0}, {0, 0.05f, 0}, {0, 0, 0}, 0, 1, 1),
           Material({0, 0, 1}, {0, 0, 0.05f}, {0, 0, 0}, 0, 1, 1), Material({1, 0, 1}, {0.05f, 0, 0.05f}, {0, 0, 0}, 0, 1, 1),
           Material({1, 1, 0}, {0.05f, 0.05f, 0}, {0, 0, 0}, 0, 1, 1), Material({0, 1, 1}, {0, 0.05f, 0.05f}, {0, 0, 0}, 0, 1, 1),
           Material({0, 0, 0}, {0, 0, 0}, {0, 0, 0}, 1, 0, 1), Material({0, 0, 0}, {0, 0, 0}, {0, 0, 0}, 0.25f, 0.5f, 1));
    sg.add(Plane(Vec3f32(0, 1, 0), 0), 0u, Cube{Vec3f32(0, 0, 0), Vec3f32(1, 1, 1)}, 1u, Cube{Vec3f32(-2, 0, -2), Vec3f32(-1, 1, -1)}, 2u,
           Triangle(Vec3f32(1, 1, 0), Vec3f32(-1, 1, 0), Vec3f32(1, 0, 1)), 3u, Triangle(Vec3f32(-1, 4, 0), Vec3f32(1, 4, 0), Vec3f32(1, 3, 1)), 4u,
           Triangle(Vec3f32(-1, 7, 0), Vec3f32(1, 7, 0), Vec3f32(1, 5, 1)), 5u,
           Sphere(Vec3f32(0, 1, 5), 1), 0u, Sphere(Vec3f32(0, 1, -5), 1), 1u, Sphere(Vec3f32(3, 1, 0), 1), 2u, Sphere(Vec3f32(0, 6, 0), 1), 3u);
    // point lights first on purpose: update() must still put the directional light at index 0
    sg.add(Light(Vec3f32(0, 0.1f, 0), Vec3f32(1, 0, 0), 5, 0.3f), Light(Vec3f32(2, 2, 2), Vec3f32(0, 1, 1), 7, 0.6f),
           Light(Vec3f32(-0.5f, -2, -1).normalize(), Vec3f32(0.9f, 0.9f, 0.9f)));
    dyn[0] = sg.addGeometry(Sphere(Vec3f32(7, 2, 0), 1), 4);
    dyn[1] = sg.addGeometry(Sphere(Vec3f32(-5, 3, 0), 1), 0);
    dyn[2] = sg.addGeometry(Sphere(Vec3f32(0, 4, 0), 1), 7);
}

int main(int argc, char** argv) {
    const bool render = argc > 1 && !std::strcmp(argv[1], "render");
    Device dev(0, 64, 8, 16, 4, 16, 2);   // triangles, lights, materials, cubes, spheres, planes
    if (render) CHECK(dev.valid());
    SceneGraph sg(dev, "Niels scene", "");
    u64 dyn[3];
    nielsScene(sg, dyn);
    CHECK(dyn[0] && dyn[1] && dyn[2]);
    sg.update(0);
    const SceneGraphInfo& info = sg.getInfo();
    CHECK(info.lightCount == 3 && info.materialCount == 8 && info.triangleCount == 3 && info.sphereCount == 7 && info.cubeCount == 2 && info.planeCount == 1);
    CHECK(info.directionalLightCount == 1 && info.spotLightCount == 0 && info.pointLightCount == 2);
    const u32 want[13] = {3, 4, 5, 0, 1, 2, 3, 4, 0, 7, 1, 2, 0};
    for (int i = 0; i < 13; ++i) CHECK(sg.getMaterialIndices()[i] == want[i]);
    const Light* lights = reinterpret_cast<const Light*>(sg.getCpuData(SceneObjectType::LIGHT));
    CHECK(lights[0].type == LightType::Directional && lights[1].type == LightType::Point && lights[2].type == LightType::Point);
    CHECK(lights[1].pos.y == 0.1f && lights[2].pos.y == 2.0f);   // stable within a type

    // handles: wrong type / unknown handle
    CHECK(!sg.update(dyn[0], Cube{}));
    CHECK(!sg.update(u64(123456), Sphere()));
    CHECK(sg.update(dyn[0], Sphere(Vec3f32(7, 2.5f, 0), 1)));
    // capacity: 2 planes allowed
    CHECK(sg.addGeometry(Plane(Vec3f32(1, 0, 0), 3), 1) != 0);
    CHECK(sg.addGeometry(Plane(Vec3f32(1, 0, 0), 4), 1) == 0);
    // delete the second static sphere: later spheres move down, material table follows
    u64 extra = sg.addGeometry(Sphere(Vec3f32(9, 9, 9), 2), 6);
    CHECK(extra);
    sg.del({dyn[1]});
    sg.update(0);
    CHECK(sg.getInfo().sphereCount == 7 && sg.getInfo().planeCount == 2);
    const Sphere* sph = reinterpret_cast<const Sphere*>(sg.getCpuData(SceneObjectType::SPHERE));
    CHECK(sph[4].Position.y == 2.5f && sph[5].Position.y == 4.0f && sph[6].Radius == 2.0f);
    const u32 want2[15] = {3, 4, 5, 0, 1, 2, 3, 4, 7, 6, 1, 2, 0, 1};
    for (int i = 0; i < 14; ++i) CHECK(sg.getMaterialIndices()[i] == want2[i]);
    CHECK(sg.exists(dyn[0]) && !sg.exists(dyn[1]));
    CHECK(sg.find(dyn[2])->second.index == 5);
    // a freed slot is reused by the next add
    u64 again = sg.addGeometry(Sphere(Vec3f32(1, 1, 1), 3), 2);
    CHECK(again && sg.find(again)->second.index == 7);

    // camera maths: default pose at 640x360
    CPUCamera cam;
    cam.setSize(Vec2u32(640, 360));
    cam.updatePlanes();
    CHECK(std::fabs(cam.p0.x - 2.2222223f) < 1e-6f && cam.p0.y == 3.0f && std::fabs(cam.p0.z + 2.7002075f) < 1e-6f);
    CHECK(std::fabs(cam.p1.x - 5.7777777f) < 1e-6f && cam.p2.y == 1.0f);
    CHECK(cam.tiles.x == 40 && cam.tiles.y == 22);

    // export settings: presets and portrait mode (include/rt/raytracing_interface.hpp:20-91)
    RaytracingProperties props;
    CHECK(props.getRes().x == 7680 && props.getRes().y == 4320 && props.targetSamples == 128 && !props.shouldOutputNextFrame);
    props.setResolution(Resolution::QHD); props.isPortrait = true;
    CHECK(props.getRes().x == 1440 && props.getRes().y == 2560);
    props.setResolution(Resolution::CUSTOM);
    CHECK(props.targetSizeX == 2560);   // CUSTOM keeps the size as it is
    props.exportToPNG();
    CHECK(props.shouldOutputNextFrame);
    // PNG writer, host-only: 3x2 frame, flipped on write like the reference's stb call
    {
        const u32 img[6] = {0xFF0000FFu, 0xFF00FF00u, 0xFFFF0000u, 0x80FFFFFFu, 0x00000000u, 0x7F102030u};
        const char* pngPath = argc > 5 ? argv[5] : "/tmp/rtb_scene_graph_check.png";
        CHECK(rtb_write_png(pngPath, 3, 2, img, 1) == 0);
        CHECK(rtb_write_png("/nonexistent-dir/x.png", 3, 2, img, 1) != 0);
    }

    if (!render) {
        CHECK(!dev.valid());                       // no GPU here: creation fails loudly, nothing falls back
        CHECK(sg.error() != 0);
        std::printf("OK host-only (device error: %s)\n", dev.error().c_str());
        return 0;
    }

    // ---- with a GPU: the reference's frame loop, headless ------------------------------------------------
    Device dev2(0, 64, 8, 16, 4, 16, 2);
    SceneGraph sg2(dev2, "Niels scene", "");
    nielsScene(sg2, dyn);
    RaytracingInterface rti(dev2, &sg2);
    rti.getCompositeTask().setOffsetSource([](f32& x, f32& y) { x = 0; y = 0; });
    rti.getCompositeTask().getShadow().properties.Shadow_samples = 1;
    rti.camera.eye = Vec3f32(6, 5, 12);
    rti.resize(Vec2u32(640, 360));
    rti.update(0);
    rti.render();
    List<u32> px;
    CHECK(rti.readPixels(px));
    CHECK(rti.error() == 0);
    u64 sum = 0;
    for (u32 p : px) sum = sum * 1099511628211ull + p;
    std::printf("OK render checksum %016llx first %08x\n", (unsigned long long)sum, px[0]);
    // second frame after moving a sphere: dirty-range upload + accel rebuild path
    sg2.update(dyn[1], Sphere(Vec3f32(-4, 2.5f, 0), 1));
    rti.update(0.1);
    rti.render();
    CHECK(rti.readPixels(px));
    CHECK(rti.error() == 0);
    std::FILE* f = argc > 2 ? std::fopen(argv[2], "wb") : nullptr;
    if (f) { std::fwrite(px.data(), 4, px.size(), f); std::fclose(f); }
    // third frame after moving a TRIANGLE with update<T>(): dirty-range upload + device refit (no host rebuild)
    u64 tri1 = 0;
    for (u64 hnd = 1; hnd < 64 && !tri1; ++hnd) {
        auto it = sg2.find(hnd);
        if (sg2.exists(hnd) && it->second.type == SceneObjectType::TRIANGLE && it->second.index == 1) tri1 = hnd;
    }
    CHECK(tri1 != 0);
    rtb_accel_info ai{};
    CHECK(rtb_accel_info_get(dev2.get(), &ai) == 0 && ai.refits == 0);
    CHECK(sg2.update(tri1, Triangle(Vec3f32(-3, 6, 1), Vec3f32(2, 5, 0.5f), Vec3f32(1, 2, 3))));
    rti.update(0.2);
    rti.render();
    CHECK(rti.readPixels(px));
    CHECK(rti.error() == 0);
    CHECK(rtb_accel_info_get(dev2.get(), &ai) == 0 && ai.refits == 1);
    f = argc > 3 ? std::fopen(argv[3], "wb") : nullptr;
    if (f) { std::fwrite(px.data(), 4, px.size(), f); std::fclose(f); }
    std::printf("OK refit frame\n");
    // export path: RaytracingProperties::exportToPNG() arms the next render(): SD preset, 4 accumulated samples
    if (argc > 4) {
        rti.properties.targetOutput = argv[4];
        rti.properties.setResolution(Resolution::SD);
        rti.properties.targetSamples = 4;
        rti.properties.exportToPNG();
        rti.render();
        CHECK(rti.error() == 0);
        CHECK(!rti.properties.shouldOutputNextFrame);
        CHECK(rti.lastExport == String(argv[4]) + ".png");
        CHECK(rti.readPixels(px) && px.size() == 640u * 360u);   // the interactive size is back
        std::printf("OK export %s\n", rti.lastExport.c_str());
    }
    return 0;
}
