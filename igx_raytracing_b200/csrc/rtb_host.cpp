// rtb_host.cpp — host-side entry points of the C ABI that need no GPU: the igx:: POD constructors
// (through include/igx_rt.hpp, so the C++ facade and the C ABI cannot drift apart), the Radiance .hdr
// loader, and the synthetic scene generators of BASELINE.json.
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/igx_rt.hpp"

using namespace igx;

extern "C" {

// ref: igx/include/types/scene_object_types.hpp:87-106
void rtb_pack_triangle(const float p[9], const float* n, void* out48) {
    const Vec3f32 p0(p[0], p[1], p[2]), p1(p[3], p[4], p[5]), p2(p[6], p[7], p[8]);
    Triangle t = n ? Triangle(p0, p1, p2, Vec3f32(n[0], n[1], n[2]), Vec3f32(n[3], n[4], n[5]), Vec3f32(n[6], n[7], n[8])) : Triangle(p0, p1, p2);
    std::memcpy(out48, &t, 48);
}

// ref: scene_object_types.hpp:166-171
void rtb_pack_light_directional(const float dir[3], const float color[3], float angularExtent, void* out32) {
    Light l(Vec3f32(dir[0], dir[1], dir[2]), Vec3f32(color[0], color[1], color[2]), angularExtent);
    std::memcpy(out32, &l, 32);   // pos and origin stay zero for a directional light
}

// ref: scene_object_types.hpp:173-179
void rtb_pack_light_point(const float pos[3], const float color[3], float rad, float origin, float specularity, void* out32) {
    Light l(Vec3f32(pos[0], pos[1], pos[2]), Vec3f32(color[0], color[1], color[2]), rad, origin, specularity);
    std::memcpy(out32, &l, 32);
}

// ref: scene_object_types.hpp:267-288
void rtb_pack_material(const float albedo[3], const float ambient[3], const float emission[3], float metallic, float roughness,
                       float transparency, void* out32) {
    Material m(Vec3f32(albedo[0], albedo[1], albedo[2]), Vec3f32(ambient[0], ambient[1], ambient[2]), Vec3f32(emission[0], emission[1], emission[2]),
               metallic, roughness, transparency);
    std::memcpy(out32, &m, 32);
}

// ref: src/rt/structs.cpp:5-42, src/rt/raytracing_interface.cpp:96-107,279-328
void rtb_pack_camera(const float eye[3], float pitch, float yaw, float roll, float leftFov, float rightFov, float ipd, uint32_t projection,
                     uint32_t width, uint32_t height, uint32_t flags, float exposure, const float skyboxColor[3], void* out144) {
    rt::CPUCamera c;
    c.eye = Vec3f32(eye[0], eye[1], eye[2]);
    c.pitch = pitch; c.yaw = yaw; c.roll = roll; c.leftFov = leftFov; c.rightFov = rightFov; c.ipd = ipd;
    c.projectionType = ProjectionType(projection);
    c.flags = CameraFlags(flags);
    c.exposure = exposure;
    c.skyboxColor = Vec3f32(skyboxColor[0], skyboxColor[1], skyboxColor[2]);
    c.setSize(Vec2u32(width, height));
    c.updatePlanes();
    std::memcpy(out144, static_cast<const Camera*>(&c), 144);
}

// Radiance RGBE (what stb_image's HDR path yields for 3 float channels: mantissa * 2^(e-136), e == 0 -> 0), then igxi-convert's
// packing into rgba16f: truncating f32 -> f16, alpha 0, inf/NaN replaced by the largest finite half
// (ref: igx/igxi-tool/src/igxi/convert.cpp:59-78,148-153,209-231).
int rtb_load_hdr(const char* path, uint16_t* out, uint32_t* width, uint32_t* height) {
    FILE* f = std::fopen(path, "rb");
    if (!f) return 1;
    std::vector<uint8_t> buf;
    {
        std::fseek(f, 0, SEEK_END);
        const long size = std::ftell(f);
        std::fseek(f, 0, SEEK_SET);
        if (size <= 0) { std::fclose(f); return 2; }
        buf.resize((size_t)size);
        const size_t got = std::fread(buf.data(), 1, buf.size(), f);
        std::fclose(f);
        if (got != buf.size()) return 2;
    }
    size_t pos = 0;
    auto line = [&]() { std::string s; while (pos < buf.size() && buf[pos] != '\n') s.push_back((char)buf[pos++]); ++pos; return s; };
    const std::string magic = line();
    if (magic != "#?RADIANCE" && magic != "#?RGBE") return 3;
    bool rgbe = false;
    for (;;) {
        if (pos >= buf.size()) return 4;
        const std::string l = line();
        if (l.empty()) break;
        if (l == "FORMAT=32-bit_rle_rgbe") rgbe = true;
    }
    if (!rgbe) return 5;
    unsigned H = 0, W = 0;
    if (std::sscanf(line().c_str(), "-Y %u +X %u", &H, &W) != 2) return 6;
    *width = W; *height = H;
    if (!out) return 0;

    auto texel = [](const uint8_t* p, uint16_t* dst) {
        float v[3] = {0.f, 0.f, 0.f};
        if (p[3]) { const float scale = std::ldexp(1.0f, (int)p[3] - 136); for (int c = 0; c < 3; ++c) v[c] = (float)p[c] * scale; }
        for (int c = 0; c < 3; ++c) { f16 h(v[c]); if (((h.value >> 10) & 0x1F) == 0x1F) h.value = 0x7BFF; dst[c] = h.value; }
        dst[3] = 0;
    };
    std::vector<uint8_t> scan((size_t)W * 4);
    for (unsigned y = 0; y < H; ++y) {
        uint16_t* row = out + (size_t)y * W * 4;
        const bool rle = W >= 8 && W < 32768 && pos + 4 <= buf.size() && buf[pos] == 2 && buf[pos + 1] == 2 && !(buf[pos + 2] & 0x80) &&
                         ((((unsigned)buf[pos + 2]) << 8) | buf[pos + 3]) == W;
        if (!rle) {
            if (pos + (size_t)W * 4 > buf.size()) return 7;
            for (unsigned x = 0; x < W; ++x) texel(&buf[pos + 4 * (size_t)x], row + 4 * (size_t)x);
            pos += (size_t)W * 4;
            continue;
        }
        pos += 4;
        for (int ch = 0; ch < 4; ++ch)
            for (unsigned x = 0; x < W;) {
                if (pos >= buf.size()) return 8;
                unsigned run = buf[pos++];
                if (run > 128) {
                    run -= 128;
                    if (pos >= buf.size() || x + run > W) return 9;
                    const uint8_t v = buf[pos++];
                    while (run--) scan[(size_t)(x++) * 4 + ch] = v;
                } else {
                    if (!run || pos + run > buf.size() || x + run > W) return 10;
                    while (run--) scan[(size_t)(x++) * 4 + ch] = buf[pos++];
                }
            }
        for (unsigned x = 0; x < W; ++x) texel(&scan[4 * (size_t)x], row + 4 * (size_t)x);
    }
    return 0;
}

}  // extern "C"

// ---- synthetic scenes ---------------------------------------------------------------------------------------
namespace {

// counter-based generator (splitmix64 finaliser): value k of stream `seed` is independent of evaluation order,
// so the scene is the same for any thread count
inline uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
inline float u01(uint64_t seed, uint64_t k) { return (float)(mix64(seed ^ mix64(k)) >> 40) * (1.0f / 16777216.0f); }

template <class F>
void parallelRows(uint64_t n, F&& body) {
    unsigned threads = std::thread::hardware_concurrency();
    if (!threads) threads = 1;
    if (n < 65536) threads = 1;
    std::atomic<uint64_t> next{0};
    const uint64_t chunk = 16384;
    auto worker = [&]() { for (;;) { const uint64_t b = next.fetch_add(chunk); if (b >= n) break; body(b, std::min(n, b + chunk)); } };
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < threads; ++t) pool.emplace_back(worker);
    worker();
    for (auto& t : pool) t.join();
}

inline float heightAt(uint64_t seed, float x, float z) {
    // four octaves of value noise on a hashed lattice, smooth (quintic) interpolation
    float h = 0.0f, amp = 1.0f, freq = 0.15f;
    for (int o = 0; o < 4; ++o) {
        const float fx = x * freq, fz = z * freq;
        const float x0 = std::floor(fx), z0 = std::floor(fz);
        const float tx = fx - x0, tz = fz - z0;
        auto lat = [&](float ix, float iz) {
            const uint64_t k = ((uint64_t)(int64_t)ix * 0x1F1F1F1Full) ^ ((uint64_t)(int64_t)iz * 0x3D4D51CBull) ^ ((uint64_t)o << 56);
            return u01(seed, k) * 2.0f - 1.0f;
        };
        auto fade = [](float t) { return t * t * t * (t * (t * 6.0f - 15.0f) + 10.0f); };
        const float sx = fade(tx), sz = fade(tz);
        const float a = lat(x0, z0), b = lat(x0 + 1, z0), c = lat(x0, z0 + 1), d = lat(x0 + 1, z0 + 1);
        h += amp * ((a * (1 - sx) + b * sx) * (1 - sz) + (c * (1 - sx) + d * sx) * sz);
        amp *= 0.5f; freq *= 2.0f;
    }
    return h * 1.5f;
}

}  // namespace

extern "C" {

void rtb_gen_soup(uint64_t n, uint64_t seed, void* outTriangles) {
    Triangle* out = static_cast<Triangle*>(outTriangles);
    parallelRows(n, [&](uint64_t b, uint64_t e) {
        for (uint64_t i = b; i < e; ++i) {
            const uint64_t k = i * 12;
            const Vec3f32 c(u01(seed, k) * 20.0f - 10.0f, u01(seed, k + 1) * 20.0f - 10.0f, u01(seed, k + 2) * 20.0f - 10.0f);
            Vec3f32 p[3];
            for (int v = 0; v < 3; ++v)
                p[v] = c + Vec3f32(u01(seed, k + 3 + 3 * v) * 0.1f - 0.05f, u01(seed, k + 4 + 3 * v) * 0.1f - 0.05f, u01(seed, k + 5 + 3 * v) * 0.1f - 0.05f);
            out[i] = Triangle(p[0], p[1], p[2]);
        }
    });
}

void rtb_gen_heightfield(uint32_t grid, uint64_t seed, void* outTriangles) {
    Triangle* out = static_cast<Triangle*>(outTriangles);
    const uint32_t V = grid + 1;
    const float step = 20.0f / (float)grid;
    std::vector<float> h((size_t)V * V);
    parallelRows((uint64_t)V * V, [&](uint64_t b, uint64_t e) {
        for (uint64_t i = b; i < e; ++i) { const uint32_t ix = (uint32_t)(i % V), iz = (uint32_t)(i / V); h[i] = heightAt(seed, -10.0f + ix * step, -10.0f + iz * step); }
    });
    auto P = [&](uint32_t ix, uint32_t iz) { return Vec3f32(-10.0f + ix * step, h[(size_t)iz * V + ix], -10.0f + iz * step); };
    auto N = [&](uint32_t ix, uint32_t iz) {   // central differences of the height field, pointing up (+y)
        const uint32_t xm = ix ? ix - 1 : ix, xp = ix + 1 < V ? ix + 1 : ix, zm = iz ? iz - 1 : iz, zp = iz + 1 < V ? iz + 1 : iz;
        const float dx = (h[(size_t)iz * V + xp] - h[(size_t)iz * V + xm]) / ((float)(xp - xm) * step);
        const float dz = (h[(size_t)zp * V + ix] - h[(size_t)zm * V + ix]) / ((float)(zp - zm) * step);
        return Vec3f32(-dx, 1.0f, -dz).normalize();
    };
    parallelRows((uint64_t)grid * grid, [&](uint64_t b, uint64_t e) {
        for (uint64_t q = b; q < e; ++q) {
            const uint32_t ix = (uint32_t)(q % grid), iz = (uint32_t)(q / grid);
            const Vec3f32 a = P(ix, iz), bb = P(ix + 1, iz), c = P(ix, iz + 1), d = P(ix + 1, iz + 1);
            const Vec3f32 na = N(ix, iz), nb = N(ix + 1, iz), nc = N(ix, iz + 1), nd = N(ix + 1, iz + 1);
            out[2 * q] = Triangle(a, c, bb, na, nc, nb);
            out[2 * q + 1] = Triangle(bb, c, d, nb, nc, nd);
        }
    });
}

// ---- frame export: rgba8 -> PNG (ref: igx/igxi-tool/src/igxi/convert.cpp:747-781, stbiWrite: 4 channels, 8 bits, rows
// flipped on write because row 0 of the frame is the bottom of the view) -----------------------------------------------
namespace {
uint32_t crc32Of(const uint8_t* p, size_t n, uint32_t crc) {
    static uint32_t table[256];
    static std::atomic<bool> ready{false};
    if (!ready.load(std::memory_order_acquire)) {
        for (uint32_t i = 0; i < 256; ++i) { uint32_t c = i; for (int k = 0; k < 8; ++k) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1; table[i] = c; }
        ready.store(true, std::memory_order_release);
    }
    crc = ~crc;
    for (size_t i = 0; i < n; ++i) crc = table[(crc ^ p[i]) & 0xFFu] ^ (crc >> 8);
    return ~crc;
}
void be32(std::vector<uint8_t>& v, uint32_t x) { v.push_back(uint8_t(x >> 24)); v.push_back(uint8_t(x >> 16)); v.push_back(uint8_t(x >> 8)); v.push_back(uint8_t(x)); }
bool writeChunk(FILE* f, const char type[4], const uint8_t* data, size_t n) {
    std::vector<uint8_t> head; be32(head, (uint32_t)n); head.insert(head.end(), type, type + 4);
    uint32_t crc = crc32Of(head.data() + 4, 4, 0u);
    if (n) crc = crc32Of(data, n, crc);
    std::vector<uint8_t> tail; be32(tail, crc);
    return std::fwrite(head.data(), 1, 8, f) == 8 && (!n || std::fwrite(data, 1, n, f) == n) && std::fwrite(tail.data(), 1, 4, f) == 4;
}
// zlib stream of `raw`: through libz when the process can load it (looked up at run time: no link dependency), otherwise
// as stored deflate blocks (valid, just not compressed)
std::vector<uint8_t> zlibStream(const std::vector<uint8_t>& raw) {
    using compressBoundFn = unsigned long (*)(unsigned long);
    using compress2Fn = int (*)(uint8_t*, unsigned long*, const uint8_t*, unsigned long, int);
    static void* lib = dlopen("libz.so.1", RTLD_NOW | RTLD_LOCAL);
    if (lib && !std::getenv("RTB_PNG_STORED")) {   // RTB_PNG_STORED: force the fallback (tests)
        auto bound = reinterpret_cast<compressBoundFn>(dlsym(lib, "compressBound"));
        auto comp = reinterpret_cast<compress2Fn>(dlsym(lib, "compress2"));
        if (bound && comp) {
            std::vector<uint8_t> out(bound((unsigned long)raw.size()));
            unsigned long n = (unsigned long)out.size();
            if (comp(out.data(), &n, raw.data(), (unsigned long)raw.size(), 1) == 0) { out.resize(n); return out; }
        }
    }
    std::vector<uint8_t> out;
    out.reserve(raw.size() + raw.size() / 65535 * 5 + 16);
    out.push_back(0x78); out.push_back(0x01);
    uint32_t a = 1, b = 0;
    size_t pos = 0;
    do {
        const size_t n = std::min<size_t>(65535, raw.size() - pos);
        out.push_back(pos + n == raw.size() ? 1 : 0);
        out.push_back(uint8_t(n)); out.push_back(uint8_t(n >> 8)); out.push_back(uint8_t(~n)); out.push_back(uint8_t((~n) >> 8));
        out.insert(out.end(), raw.begin() + pos, raw.begin() + pos + n);
        for (size_t i = 0; i < n; ++i) { a += raw[pos + i]; if (a >= 65521) a -= 65521; b += a; if (b >= 65521) b -= 65521; }
        pos += n;
    } while (pos < raw.size());
    be32(out, (b << 16) | a);
    return out;
}
}  // namespace

int rtb_write_png(const char* path, uint32_t width, uint32_t height, const void* rgba8, int flipVertically) {
    if (!path || !rgba8 || !width || !height) return 1;
    const uint8_t* px = static_cast<const uint8_t*>(rgba8);
    const size_t row = (size_t)width * 4;
    std::vector<uint8_t> raw((row + 1) * height);
    for (uint32_t y = 0; y < height; ++y) {
        const uint32_t src = flipVertically ? height - 1 - y : y;
        raw[(row + 1) * y] = 0;   // filter type None
        std::memcpy(raw.data() + (row + 1) * y + 1, px + row * src, row);
    }
    const std::vector<uint8_t> z = zlibStream(raw);
    FILE* f = std::fopen(path, "wb");
    if (!f) return 2;
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    std::vector<uint8_t> ihdr; be32(ihdr, width); be32(ihdr, height);
    ihdr.push_back(8); ihdr.push_back(6); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);   // 8 bits, RGBA, deflate, adaptive, no interlace
    bool ok = std::fwrite(sig, 1, 8, f) == 8 && writeChunk(f, "IHDR", ihdr.data(), ihdr.size());
    for (size_t pos = 0; ok && pos < z.size(); pos += (size_t)1 << 30) ok = writeChunk(f, "IDAT", z.data() + pos, std::min<size_t>((size_t)1 << 30, z.size() - pos));
    ok = ok && writeChunk(f, "IEND", nullptr, 0);
    ok = (std::fclose(f) == 0) && ok;
    return ok ? 0 : 3;
}

}  // extern "C"
