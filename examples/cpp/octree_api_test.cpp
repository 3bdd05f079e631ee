// Host-side (no GPU) tests of the C++ mirror, written like the reference's own tests
// (src/octree/update/tests.rs, src/octree/mod.rs). Build: see tests/test_cpp_mirror.py.
#include <cstdio>
#include <cstdlib>

#include "shocovox_b200.hpp"

using namespace svx;

#define EXPECT(cond)                                                     \
    do {                                                                 \
        if (!(cond)) {                                                   \
            std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            std::exit(1);                                                \
        }                                                                \
    } while (0)

template <typename F>
static svx_status error_of(F&& f) {
    try {
        f();
    } catch (const OctreeError& e) {
        return e.code;
    }
    return SVX_OK;
}

// update/tests.rs:20-46
static void test_simple_insert_and_get() {
    const Albedo red = Albedo::from(0xFF0000FF), green = Albedo::from(0x00FF00FF), blue = Albedo::from(0x0000FFFF);
    Octree tree = Octree::create(2, 1);
    tree.set_auto_simplify(false);
    tree.insert({1, 0, 0}, red);
    tree.insert({0, 1, 0}, green);
    tree.insert({0, 0, 1}, blue);
    EXPECT(tree.get({1, 0, 0}) == OctreeEntry::Visual(red));
    EXPECT(tree.get({0, 1, 0}) == OctreeEntry::Visual(green));
    EXPECT(tree.get({0, 0, 1}) == OctreeEntry::Visual(blue));
    EXPECT(tree.get({1, 1, 1}) == OctreeEntry::Empty());
    tree.insert({1, 0, 0}, green);
    EXPECT(tree.get({1, 0, 0}) == OctreeEntry::Visual(green));
}

// update/tests.rs:57-110
static void test_complex_insert_and_get() {
    const Albedo red = Albedo::from(0xFF0000FF), green = Albedo::from(0x00FF00FF);
    Octree tree = Octree::create(2, 1);
    tree.set_auto_simplify(false);
    tree.insert({1, 0, 0}, OctreeEntry::Complex(red, 3));
    tree.insert({0, 1, 0}, OctreeEntry::Complex(green, 1));
    tree.insert({0, 0, 1}, OctreeEntry::Informative(2));
    EXPECT(tree.get({1, 0, 0}) == OctreeEntry::Complex(red, 3));
    EXPECT(tree.get({0, 1, 0}) == OctreeEntry::Complex(green, 1));
    EXPECT(tree.get({0, 0, 1}) == OctreeEntry::Informative(2));
    tree.update({1, 0, 0}, OctreeEntry::Informative(4));  // update/tests.rs:376-388
    EXPECT(tree.get({1, 0, 0}) == OctreeEntry::Complex(red, 4));
}

// update/tests.rs:143-191 and :1281-1315
static void test_insert_and_clear_at_lod() {
    const Albedo c = Albedo::from(0xFFAAEEFF);
    Octree tree = Octree::create(8, 1);
    tree.insert_at_lod({0, 0, 0}, 4, OctreeEntry::Visual(c));
    tree.clear_at_lod({0, 0, 0}, 2);
    int hits = 0;
    for (uint32_t x = 0; x < 4; ++x)
        for (uint32_t y = 0; y < 4; ++y)
            for (uint32_t z = 0; z < 4; ++z) {
                const OctreeEntry e = tree.get({x, y, z});
                if (e.is_some()) {
                    EXPECT(e == OctreeEntry::Visual(c));
                    ++hits;
                }
            }
    EXPECT(hits == 64 - 8);
    tree.clear({3, 3, 3});
    EXPECT(tree.get({3, 3, 3}).is_none());
}

// src/octree/mod.rs:173-187 (validation order), insert.rs:108-114
static void test_errors() {
    EXPECT(error_of([] { Octree::create(0, 8); }) == SVX_E_INVALID_BRICK_DIMENSION);
    EXPECT(error_of([] { Octree::create(64, 3); }) == SVX_E_INVALID_BRICK_DIMENSION);
    EXPECT(error_of([] { Octree::create(4, 8); }) == SVX_E_INVALID_SIZE);
    EXPECT(error_of([] { Octree::create(8, 8); }) == SVX_E_INVALID_STRUCTURE);
    Octree tree = Octree::create(4, 1);
    EXPECT(error_of([&] { tree.insert({4, 0, 0}, Albedo::from(0xFF0000FF)); }) == SVX_E_INVALID_POSITION);
    tree.insert({3, 0, 0}, Albedo{});  // an empty entry is a no-op Ok (insert.rs:117-119)
    EXPECT(tree.get({3, 0, 0}).is_none());
    EXPECT(tree.get_size() == 4);
}

// src/convert/bytecode_tests.rs:182-222 (test_octree_file_io) through the C++ mirror, plus the error paths
static void test_save_load() {
    const Albedo red = Albedo::from(0xFF0000FF);
    Octree tree = Octree::create(8, 1);
    tree.insert_at_lod({0, 0, 0}, 4, OctreeEntry::Visual(red));
    tree.clear_at_lod({0, 0, 0}, 2);
    tree.save("test_junk_octree_cpp");
    Octree copy = Octree::load("test_junk_octree_cpp");
    std::remove("test_junk_octree_cpp");
    int hits = 0;
    for (uint32_t x = 0; x < 4; ++x)
        for (uint32_t y = 0; y < 4; ++y)
            for (uint32_t z = 0; z < 4; ++z) {
                EXPECT(tree.get({x, y, z}) == copy.get({x, y, z}));
                if (copy.get({x, y, z}).is_some()) ++hits;
            }
    EXPECT(hits == 64 - 8);
    EXPECT(copy.structure_hash() == tree.structure_hash());
    const std::vector<uint8_t> bytes = tree.to_bytes();
    EXPECT(Octree::from_bytes(bytes).to_bytes() == bytes);
    EXPECT(error_of([] { Octree::load("no/such/file"); }) == SVX_E_IO);
    EXPECT(error_of([&] { Octree::from_bytes(std::vector<uint8_t>(bytes.begin(), bytes.end() - 3)); }) == SVX_E_DECODE);
}

// src/octree/tests.rs:365-468 (mipmap_tests::test_mixed_mip_lvl2_where_dim_is_4) through the StrategyUpdater mirror
static void test_mip_maps() {
    const Albedo red = Albedo::from(0xFF0000FF), green = Albedo::from(0x00FF00FF), blue = Albedo::from(0x0000FFFF);
    Octree tree = Octree::create(16, 4);
    tree.set_auto_simplify(false);
    tree.albedo_mip_map_resampling_strategy()
        .switch_albedo_mip_maps(true)
        .set_method_at(1, MIPResamplingMethods::BoxFilter())
        .set_method_at(2, MIPResamplingMethods::BoxFilter());
    const std::pair<V3c<uint32_t>, Albedo> voxels[] = {
        {{0, 0, 0}, red}, {{0, 0, 1}, green}, {{0, 1, 0}, red}, {{0, 1, 1}, green}, {{1, 0, 0}, red}, {{1, 0, 1}, green},
        {{8, 0, 0}, red}, {{8, 0, 1}, green}, {{8, 1, 0}, blue}, {{8, 1, 1}, green}, {{9, 1, 0}, red}, {{9, 0, 1}, blue}};
    for (const auto& v : voxels) tree.insert(v.first, v.second);
    const uint8_t m2 = 180, m3 = 147;  // sqrt(255^2 / 2), sqrt(255^2 / 3), truncated
    StrategyUpdater mips = tree.albedo_mip_map_resampling_strategy();
    EXPECT(mips.is_enabled());
    EXPECT(mips.sample_root_mip(0, {0, 0, 0}) == OctreeEntry::Visual(Albedo{m2, m2, 0, 255}));
    EXPECT(mips.sample_root_mip(1, {0, 0, 0}) == OctreeEntry::Visual(Albedo{m3, m3, m3, 255}));
    EXPECT(mips.sample_root_mip(8, {0, 0, 0}) == OctreeEntry::Visual(Albedo{m2, m2, 0, 255}));
    EXPECT(mips.sample_root_mip(8, {2, 0, 0}) == OctreeEntry::Visual(Albedo{m3, m3, m3, 255}));
    EXPECT(mips.sample_root_mip(8, {1, 1, 1}).is_none());
    EXPECT(mips.get_method_at(1) == MIPResamplingMethods::BoxFilter());
    EXPECT(mips.set_method_at(3, MIPResamplingMethods::Posterize(2.0f)).get_method_at(3) == MIPResamplingMethods::Posterize(1.0f));
    EXPECT(mips.get_new_color_similarity_at(3) == 0.05f);
    // the MIPs and their strategy travel with the tree
    Octree copy = Octree::from_bytes(tree.to_bytes());
    EXPECT(copy.albedo_mip_map_resampling_strategy().is_enabled());
    EXPECT(copy.albedo_mip_map_resampling_strategy().sample_root_mip(8, {2, 0, 0}) == OctreeEntry::Visual(Albedo{m3, m3, m3, 255}));
    EXPECT(!mips.reset().is_enabled());
}

// without a CUDA device the ray path must throw, never fall back to the CPU
static void test_no_cpu_fallback() {
    if (svx_cuda_device_count() > 0) return;
    Octree tree = Octree::create(4, 1);
    tree.insert({1, 1, 1}, Albedo::from(0xFF0000FF));
    EXPECT(error_of([&] { OctreeGPUHost host(tree); }) == SVX_E_CUDA);
}

// Octree::load_vox_file through the mirror (src/convert/magicavoxel.rs:266-289), plain and with a MIP strategy installed first
static void test_load_vox_file(const char* path) {
    Octree plain = Octree::load_vox_file(path, 8);
    std::FILE* f = std::fopen(path, "rb");
    EXPECT(f != nullptr);
    std::vector<uint8_t> bytes;
    for (int c; (c = std::fgetc(f)) != EOF;) bytes.push_back((uint8_t)c);
    std::fclose(f);
    Octree with_mips = Octree::load_vox_file(bytes, 8, [](StrategyUpdater s) { s.switch_albedo_mip_maps(true); });
    EXPECT(plain.get_size() == with_mips.get_size());
    // (the structure hash covers the palette, which the MIPs extend with averaged colours: compare the voxels themselves)
    size_t filled = 0;
    for (uint32_t x = 0; x < plain.get_size(); x += 3)
        for (uint32_t y = 0; y < plain.get_size(); y += 3)
            for (uint32_t z = 0; z < plain.get_size(); z += 3) {
                EXPECT(plain.get({x, y, z}) == with_mips.get({x, y, z}));
                filled += plain.get({x, y, z}).is_some() ? 1 : 0;
            }
    EXPECT(filled > 10);
    EXPECT(with_mips.albedo_mip_map_resampling_strategy().is_enabled());
    std::printf("vox size %u hash %016llx\n", plain.get_size(), (unsigned long long)plain.structure_hash());
    EXPECT(error_of([&] { Octree::load_vox_file("/nonexistent.vox", 8); }) == SVX_E_IO);
}

int main(int argc, char** argv) {
    if (argc > 1) test_load_vox_file(argv[1]);
    test_simple_insert_and_get();
    test_complex_insert_and_get();
    test_insert_and_clear_at_lod();
    test_errors();
    test_save_load();
    test_mip_maps();
    test_no_cpu_fallback();
    std::puts("octree_api_test: ok");
    return 0;
}
