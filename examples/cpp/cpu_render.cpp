// The reference's examples/cpu_render.rs, on the B200 path through the C++ mirror: same 64/8 scene (cpu_render.rs:13-43),
// same camera recipe (:49-94), one GPU launch instead of the per-pixel CPU loop (:104-136). Writes a PPM and prints the
// hit count and an FNV-1a digest of the three framebuffer planes. Also replays the literal ray of the reference's
// test_edge_case_detailed_brick_z_edge_error (src/raytracing/tests.rs:598-628) through get_by_ray.
//   usage: cpu_render [width height [out.ppm [ox oy oz dx dy dz]]]   (camera origin / unit direction as %.9g floats)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "shocovox_b200.hpp"

using namespace svx;

static uint64_t fnv1a(const void* data, size_t n, uint64_t h = 1469598103934665603ull) {
    const unsigned char* p = static_cast<const unsigned char*>(data);
    for (size_t i = 0; i < n; ++i) h = (h ^ p[i]) * 1099511628211ull;
    return h;
}

int main(int argc, char** argv) {
    const uint32_t W = argc > 2 ? (uint32_t)std::atoi(argv[1]) : 150, H = argc > 2 ? (uint32_t)std::atoi(argv[2]) : 150;
    constexpr uint32_t BRICK_DIMENSION = 8, TREE_SIZE = 64;
    try {
        Octree tree = Octree::create(TREE_SIZE, BRICK_DIMENSION);
        tree.insert({1, 3, 3}, Albedo::from(0x645097FF));
        for (uint32_t x = 0; x < TREE_SIZE; ++x)
            for (uint32_t y = 0; y < TREE_SIZE; ++y)
                for (uint32_t z = 0; z < TREE_SIZE; ++z)
                    if (((x < TREE_SIZE / 4 || y < TREE_SIZE / 4 || z < TREE_SIZE / 4) && x % 2 == 0 && y % 4 == 0 && z % 2 == 0) ||
                        (TREE_SIZE / 2 <= x && TREE_SIZE / 2 <= y && TREE_SIZE / 2 <= z))
                        tree.insert({x, y, z}, Albedo{(uint8_t)(255.0f * (float)x / (float)TREE_SIZE), (uint8_t)(255.0f * (float)y / (float)TREE_SIZE),
                                                      (uint8_t)(255.0f * (float)z / (float)TREE_SIZE), 255});
        // camera of cpu_render.rs:49-94 with angle = 40 (origin bits from the command line would remove the libm dependency)
        const float radius = 2.0f * (float)TREE_SIZE, angle = 40.0f;
        const V3c<float> origin{std::sin(angle) * radius, radius, std::cos(angle) * radius};
        const float len = std::sqrt((origin.x * origin.x) + (origin.y * origin.y) + (origin.z * origin.z));
        Viewport vp;
        vp.origin = origin;
        vp.direction = {(0.0f - origin.x) / len, (0.0f - origin.y) / len, (0.0f - origin.z) / len};
        if (argc > 9) {  // an exact camera from the caller (libm's sinf/cosf may differ by an ulp between platforms)
            vp.origin = {std::strtof(argv[4], nullptr), std::strtof(argv[5], nullptr), std::strtof(argv[6], nullptr)};
            vp.direction = {std::strtof(argv[7], nullptr), std::strtof(argv[8], nullptr), std::strtof(argv[9], nullptr)};
        }
        vp.frustum = {4.0f, 4.0f, 3.0f};
        vp.fov = 3.0f;

        OctreeGPUHost host(tree);
        OctreeGPUView view = host.create_new_view(64, vp, {W, H});
        // define light: V3c::new(0., -1., 1.).normalized() (cpu_render.rs:97)
        const float ll = std::sqrt((0.0f * 0.0f) + (-1.0f * -1.0f) + (1.0f * 1.0f));
        view.set_shading({0.0f / ll, -1.0f / ll, 1.0f / ll});
        const Frame frame = view.render_to_host();
        const std::vector<uint32_t> shaded = view.read_shaded();  // the image the reference example builds
        size_t hits = 0;
        for (uint32_t id : frame.hit_id) hits += id != 0xFFFFFFFFu;
        uint64_t digest = fnv1a(frame.hit_id.data(), frame.hit_id.size() * 4);
        digest = fnv1a(frame.albedo.data(), frame.albedo.size() * 4, digest);
        digest = fnv1a(frame.distance.data(), frame.distance.size() * 4, digest);
        std::printf("frame %ux%u hits %zu digest %016llx\n", W, H, hits, (unsigned long long)digest);
        if (argc > 3) {
            if (FILE* f = std::fopen(argv[3], "wb")) {
                std::fprintf(f, "P6\n%u %u\n255\n", W, H);
                for (size_t i = 0; i < shaded.size(); ++i) {
                    const uint32_t c = shaded[i];
                    const unsigned char rgb[3] = {(unsigned char)(c & 0xFF), (unsigned char)((c >> 8) & 0xFF), (unsigned char)((c >> 16) & 0xFF)};
                    std::fwrite(rgb, 1, 3, f);
                }
                std::fclose(f);
            }
        }
        // src/raytracing/tests.rs:598-628
        Octree t2 = Octree::create(8, 2);
        for (uint32_t x = 1; x < 8; ++x)
            for (uint32_t y = 1; y < 8; ++y)
                for (uint32_t z = 1; z < 8; ++z) t2.insert({x, y, z}, Albedo::from(z));
        OctreeGPUHost h2(t2);
        const auto hit = h2.get_by_ray(Ray{{11.92238f, 16.0f, -10.670372f}, {-0.30062392f, -0.6361918f, 0.7105529f}});
        const bool ok = hit && hit->entry == OctreeEntry::Visual(Albedo::from(1)) && hit->normal.x == 0.0f && hit->normal.y == 0.0f &&
                        hit->normal.z == -1.0f;
        std::printf("detailed_brick_z_edge_error %s\n", ok ? "ok" : "FAILED");
        return ok ? 0 : 1;
    } catch (const OctreeError& e) {
        std::fprintf(stderr, "OctreeError %s\n", e.what());
        return 2;
    }
}
