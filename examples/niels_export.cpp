// niels_export.cpp — the reference's demo (test/main.cpp + test/scene/niels_scene.cpp), headless: build the default scene
// through igx::SceneGraph, drive igx::rt::RaytracingInterface the way the viewport thread does (resize / update / render),
// move the three animated spheres (NielsScene::update), and write a frame through the reference's export path
// (RaytracingProperties::exportToPNG).  Everything below the facade is the C ABI of librtb200 (include/rtb200.h).
//
//   g++ -O2 -std=c++17 -pthread -I include examples/niels_export.cpp -L igx_raytracing_b200 -lrtb200 \
//       -Wl,-rpath,$PWD/igx_raytracing_b200 -o niels_export
//   ./niels_export out/frame [width height samples]        -> out/frame.png
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "igx_rt.hpp"

using namespace igx;
using namespace igx::rt;

int main(int argc, char** argv) {
    const char* out = argc > 1 ? argv[1] : "./niels";
    const u16 w = argc > 2 ? (u16)std::atoi(argv[2]) : 1920, h = argc > 3 ? (u16)std::atoi(argv[3]) : 1080;
    const u16 samples = argc > 4 ? (u16)std::atoi(argv[4]) : 16;

    Device dev(/*cuda device*/ 0, /*triangles*/ 64, /*lights*/ 8, /*materials*/ 16, /*cubes*/ 4, /*spheres*/ 16, /*planes*/ 2);
    if (!dev.valid()) { std::fprintf(stderr, "no usable CUDA device: %s\n", dev.error().c_str()); return 2; }

    // test/scene/niels_scene.cpp:10-56
    SceneGraph scene(dev, "Niels scene", "");
    scene.add(Material({1, 0.5f, 1}, {0.05f, 0.01f, 0.05f}, {0, 0, 0}, 0, 1, 1), Material({0, 1, 0}, {0, 0.05f, 0}, {0, 0, 0}, 0, 1, 1),
              Material({0, 0, 1}, {0, 0, 0.05f}, {0, 0, 0}, 0, 1, 1), Material({1, 0, 1}, {0.05f, 0, 0.05f}, {0, 0, 0}, 0, 1, 1),
              Material({1, 1, 0}, {0.05f, 0.05f, 0}, {0, 0, 0}, 0, 1, 1), Material({0, 1, 1}, {0, 0.05f, 0.05f}, {0, 0, 0}, 0, 1, 1),
              Material({0, 0, 0}, {0, 0, 0}, {0, 0, 0}, 1, 0, 1), Material({0, 0, 0}, {0, 0, 0}, {0, 0, 0}, 0.25f, 0.5f, 1));
    scene.add(Plane(Vec3f32(0, 1, 0), 0), 0u, Cube{Vec3f32(0, 0, 0), Vec3f32(1, 1, 1)}, 1u, Cube{Vec3f32(-2, 0, -2), Vec3f32(-1, 1, -1)}, 2u,
              Triangle(Vec3f32(1, 1, 0), Vec3f32(-1, 1, 0), Vec3f32(1, 0, 1)), 3u, Triangle(Vec3f32(-1, 4, 0), Vec3f32(1, 4, 0), Vec3f32(1, 3, 1)), 4u,
              Triangle(Vec3f32(-1, 7, 0), Vec3f32(1, 7, 0), Vec3f32(1, 5, 1)), 5u,
              Sphere(Vec3f32(0, 1, 5), 1), 0u, Sphere(Vec3f32(0, 1, -5), 1), 1u, Sphere(Vec3f32(3, 1, 0), 1), 2u, Sphere(Vec3f32(0, 6, 0), 1), 3u);
    scene.add(Light(Vec3f32(-0.5f, -2, -1).normalize(), Vec3f32(0.9f, 0.9f, 0.9f)), Light(Vec3f32(0, 0.1f, 0), Vec3f32(1, 0, 0), 5, 0.3f),
              Light(Vec3f32(2, 2, 2), Vec3f32(0, 1, 1), 7, 0.6f));
    const u64 dyn[3] = {scene.addGeometry(Sphere(Vec3f32(7, 2, 0), 1), 4), scene.addGeometry(Sphere(Vec3f32(-5, 3, 0), 1), 0),
                        scene.addGeometry(Sphere(Vec3f32(0, 4, 0), 1), 7)};

    RaytracingInterface rti(dev, &scene);
    rti.camera.eye = Vec3f32(6, 5, 12);
    rti.resize(Vec2u32(640, 360));             // the "window"

    // a few interactive frames with the animated spheres (test/scene/niels_scene.cpp:61-70)
    f64 time = 0;
    for (int frame = 0; frame < 8; ++frame, time += 1.0 / 60) {
        scene.update(dyn[0], Sphere(Vec3f32(7 + (f32)std::sin(time), 2, 0), 1));
        scene.update(dyn[1], Sphere(Vec3f32(-5, 3 + (f32)std::cos(time), 0), 1));
        scene.update(dyn[2], Sphere(Vec3f32(0, 4, (f32)std::sin(time) * 2), 1));
        rti.update(1.0 / 60);
        rti.render();
    }

    // the export button: target size and sample count, then the next render() writes <out>.png
    rti.properties.targetOutput = out;
    rti.properties.setResolution(Resolution::CUSTOM);
    rti.properties.targetSizeX = w; rti.properties.targetSizeY = h;
    rti.properties.targetSamples = samples;
    rti.properties.exportToPNG();
    rti.render();
    if (rti.error() || rti.lastExport.empty()) { std::fprintf(stderr, "export failed (%d): %s\n", rti.error(), dev.error().c_str()); return 1; }
    std::printf("wrote %s (%ux%u, %u samples)\n", rti.lastExport.c_str(), (unsigned)w, (unsigned)h, (unsigned)samples);
    return 0;
}
