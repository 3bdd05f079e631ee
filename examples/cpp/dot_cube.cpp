// The reference's examples/dot_cube.rs, on the B200 path through the C++ mirror: its 256/32 tree (dot_cube.rs:54-104), its
// viewport (origin (2S, S/2, -2S) looking at 0, frustum (10, 10, 200), fov 3, :48-52 / :108-116) and the CPU render it
// does when Tab is pressed (:196-259: glass at direction * frustum.z, diffuse shading, grey background) - as one GPU frame
// with the shaded fourth plane. With `--mips` the tree gets MIP maps and the frame is rendered through get_by_ray_at_lod
// at viewing distance frustum.z, which is what the reference's own GPU shader does with this viewport.
//   usage: dot_cube [width height [out.ppm]] [--mips]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "shocovox_b200.hpp"

using namespace svx;

int main(int argc, char** argv) {
    bool mips = false;
    const char* pos[3] = {nullptr, nullptr, nullptr};
    int n_pos = 0;
    for (int i = 1; i < argc; ++i) {
        if (std::strcmp(argv[i], "--mips") == 0)
            mips = true;
        else if (n_pos < 3)
            pos[n_pos++] = argv[i];
    }
    const uint32_t W = n_pos >= 2 ? (uint32_t)std::atoi(pos[0]) : 64, H = n_pos >= 2 ? (uint32_t)std::atoi(pos[1]) : 64;
    constexpr uint32_t BRICK_DIMENSION = 32, TREE_SIZE = 256;
    try {
        Octree tree = Octree::create(TREE_SIZE, BRICK_DIMENSION);
        if (mips) tree.albedo_mip_map_resampling_strategy().switch_albedo_mip_maps(true);  // kept current by every insert
        auto channel = [](uint32_t c) -> uint8_t {
            return 0 == c % (TREE_SIZE / 4) ? (uint8_t)(uint32_t)((float)c / (float)TREE_SIZE * 255.0f) : (uint8_t)128;
        };
        uint64_t inserted = 0;
        for (uint32_t x = 0; x < TREE_SIZE; ++x)
            for (uint32_t y = 0; y < TREE_SIZE; ++y)
                for (uint32_t z = 0; z < TREE_SIZE; ++z)
                    if (((x < TREE_SIZE / 4 || y < TREE_SIZE / 4 || z < TREE_SIZE / 4) && 0 == x % 2 && 0 == y % 4 && 0 == z % 2) ||
                        (TREE_SIZE / 2 <= x && TREE_SIZE / 2 <= y && TREE_SIZE / 2 <= z)) {
                        tree.insert({x, y, z}, Albedo{channel(x), channel(y), channel(z), 255});
                        ++inserted;
                    }
        Viewport vp;
        vp.origin = {(float)TREE_SIZE * 2.0f, (float)TREE_SIZE / 2.0f, (float)TREE_SIZE * -2.0f};
        const V3c<float> d{0.0f - vp.origin.x, 0.0f - vp.origin.y, 0.0f - vp.origin.z};
        const float len = std::sqrt((d.x * d.x) + (d.y * d.y) + (d.z * d.z));
        vp.direction = {d.x / len, d.y / len, d.z / len};
        vp.frustum = {10.0f, 10.0f, 200.0f};
        vp.fov = 3.0f;

        OctreeGPUHost host(tree);
        OctreeGPUView view = host.create_new_view(50, vp, {W, H});
        view.set_glass_mode(SVX_GLASS_AT_FRUSTUM_Z);  // viewport_bottom_left = origin + direction * frustum.z (:208-209)
        if (mips) view.set_viewing_distance(vp.frustum.z);
        const float ll = std::sqrt((0.0f * 0.0f) + (-1.0f * -1.0f) + (1.0f * 1.0f));
        view.set_shading({0.0f / ll, -1.0f / ll, 1.0f / ll});  // V3c::new(0., -1., 1.).normalized() (:214)
        const svx_frame on_device = view.render();
        const Frame frame = view.render_to_host();
        const std::vector<uint32_t> shaded = view.read_shaded();
        size_t hits = 0;
        for (uint32_t id : frame.hit_id) hits += id != 0xFFFFFFFFu;
        std::printf("dot_cube: %llu voxels, %llu nodes on the device, %ux%u frame, %zu hits, kernel %.3f ms%s\n",
                    (unsigned long long)inserted, (unsigned long long)host.stats().nodes, W, H, hits, on_device.kernel_ms,
                    mips ? " (MIP maps on, viewing distance = frustum.z)" : "");
        if (n_pos >= 3) {
            if (FILE* f = std::fopen(pos[2], "wb")) {
                std::fprintf(f, "P6\n%u %u\n255\n", W, H);
                for (uint32_t c : shaded) {
                    const unsigned char rgb[3] = {(unsigned char)(c & 0xFF), (unsigned char)((c >> 8) & 0xFF), (unsigned char)((c >> 16) & 0xFF)};
                    std::fwrite(rgb, 1, 3, f);
                }
                std::fclose(f);
            }
        }
        return hits > 0 ? 0 : 1;
    } catch (const OctreeError& e) {
        std::fprintf(stderr, "OctreeError %s\n", e.what());
        return 2;
    }
}
