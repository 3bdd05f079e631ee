/* The multi-GPU entry points of the C ABI from plain C (what a Rust `extern "C"` block would bind): one frame of the
 * reference's examples/cpu_render.rs scene, tile-sharded over the listed CUDA devices by svx_multi_*, must equal the frame
 * one GPU renders, byte for byte - gathered over peer stores into devices[0]'s framebuffer (both wire formats), assembled in
 * host memory over every GPU's own PCIe link, and as a pose batch.
 *   usage: multi_gpu [device ...]        default "0 0": two members on one device (each with its own tree replica and stream)
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "shocovox_b200.h"

#define CHECK(call)                                                                                          \
    do {                                                                                                     \
        int32_t s__ = (call);                                                                                \
        if (s__ != SVX_OK) {                                                                                 \
            fprintf(stderr, "%s failed: %d (%s)\n", #call, (int)s__, svx_last_error_message());              \
            return 1;                                                                                        \
        }                                                                                                    \
    } while (0)

enum { W = 640, H = 363, TREE = 64, DIM = 8, POSES = 5 };

static svx_viewport camera(int k) { /* examples/cpu_render.rs:49-94 with angle = 40 + 0.05 k */
    svx_viewport vp;
    const float radius = 2.0f * (float)TREE, angle = 40.0f + 0.05f * (float)k;
    const float o[3] = {sinf(angle) * radius, radius, cosf(angle) * radius};
    const float len = sqrtf((o[0] * o[0]) + (o[1] * o[1]) + (o[2] * o[2]));
    for (int i = 0; i < 3; ++i) {
        vp.origin[i] = o[i];
        vp.direction[i] = (0.0f - o[i]) / len;
    }
    vp.frustum[0] = 4.0f;
    vp.frustum[1] = 4.0f;
    vp.frustum[2] = 3.0f;
    vp.fov = 3.0f;
    return vp;
}

int main(int argc, char** argv) {
    int32_t devices[16] = {0, 0};
    uint32_t n = 2;
    if (argc > 1) {
        n = 0;
        for (int i = 1; i < argc && n < 16; ++i) devices[n++] = (int32_t)atoi(argv[i]);
    }
    svx_octree* tree = NULL;
    CHECK(svx_octree_new(TREE, DIM, &tree));
    for (uint32_t x = 0; x < TREE; ++x)
        for (uint32_t y = 0; y < TREE; ++y)
            for (uint32_t z = 0; z < TREE; ++z)
                if (((x < TREE / 4 || y < TREE / 4 || z < TREE / 4) && x % 2 == 0 && y % 4 == 0 && z % 2 == 0) ||
                    (TREE / 2 <= x && TREE / 2 <= y && TREE / 2 <= z)) {
                    svx_entry e;
                    memset(&e, 0, sizeof(e));
                    e.kind = SVX_ENTRY_VISUAL;
                    e.albedo.r = (uint8_t)(255.0f * (float)x / (float)TREE);
                    e.albedo.g = (uint8_t)(255.0f * (float)y / (float)TREE);
                    e.albedo.b = (uint8_t)(255.0f * (float)z / (float)TREE);
                    e.albedo.a = 255;
                    CHECK(svx_octree_insert(tree, x, y, z, &e));
                }
    const size_t px = (size_t)W * H;
    uint32_t* want = (uint32_t*)malloc(POSES * 3 * px * 4);
    uint32_t* got = (uint32_t*)malloc(POSES * 3 * px * 4);
    if (!want || !got) return 1;

    /* the single-GPU frames */
    svx_gpu_host* host = NULL;
    svx_view* view = NULL;
    svx_viewport vp = camera(0);
    CHECK(svx_gpu_host_create(tree, devices[0], &host));
    CHECK(svx_gpu_host_create_view(host, 64, &vp, W, H, &view));
    for (int k = 0; k < POSES; ++k) {
        vp = camera(k);
        CHECK(svx_view_set_viewport(view, &vp));
        CHECK(svx_view_render_to_host(view, want + (size_t)k * px, want + (POSES + k) * px, (float*)(want + (2 * POSES + k) * px)));
    }
    size_t hits = 0;
    for (size_t i = 0; i < px; ++i) hits += want[i] != 0xFFFFFFFFu;

    for (int wire = SVX_WIRE_THREE_PLANES; wire <= SVX_WIRE_ID_DISTANCE; ++wire) {
        svx_multi* multi = NULL;
        vp = camera(0);
        CHECK(svx_multi_create(tree, devices, n, &vp, W, H, 8, wire, &multi));
        for (int k = 0; k < 3; ++k) {
            vp = camera(k);
            CHECK(svx_multi_set_viewport(multi, &vp));
            /* gathered in devices[0]'s framebuffer by the viewport kernels themselves */
            svx_frame frame;
            CHECK(svx_multi_render(multi, &frame));
            CHECK(svx_view_read_frame(svx_multi_view(multi, 0), got, got + px, (float*)(got + 2 * px)));
            if (memcmp(got, want + (size_t)k * px, px * 4) || memcmp(got + px, want + (POSES + k) * px, px * 4) ||
                memcmp(got + 2 * px, want + (2 * POSES + k) * px, px * 4)) {
                fprintf(stderr, "wire %d pose %d: the gathered frame differs from the single-GPU frame\n", wire, k);
                return 2;
            }
            /* assembled in host memory, every GPU copying its own rows */
            memset(got, 0xAB, 3 * px * 4);
            CHECK(svx_multi_render_to_host(multi, got, got + px, (float*)(got + 2 * px)));
            if (memcmp(got, want + (size_t)k * px, px * 4) || memcmp(got + px, want + (POSES + k) * px, px * 4) ||
                memcmp(got + 2 * px, want + (2 * POSES + k) * px, px * 4)) {
                fprintf(stderr, "wire %d pose %d: the host-assembled frame differs from the single-GPU frame\n", wire, k);
                return 3;
            }
        }
        /* pose batch: pose k on device k % n */
        svx_viewport poses[POSES];
        for (int k = 0; k < POSES; ++k) poses[k] = camera(k);
        float ms = 0.0f;
        CHECK(svx_multi_render_poses(multi, poses, POSES, got, got + POSES * px, (float*)(got + 2 * POSES * px), &ms));
        if (memcmp(got, want, POSES * 3 * px * 4)) {
            fprintf(stderr, "wire %d: the pose batch differs from the single-GPU frames\n", wire);
            return 4;
        }
        svx_multi_free(multi);
    }
    /* a degenerate camera is refused, not rendered (every ray would be NaN) */
    vp = camera(0);
    vp.direction[0] = 0.0f;
    vp.direction[1] = -1.0f;
    vp.direction[2] = 0.0f;
    if (svx_view_set_viewport(view, &vp) != SVX_E_INVALID_ARGUMENT) {
        fprintf(stderr, "a viewport looking along the up vector was accepted\n");
        return 5;
    }
    svx_view_free(view);
    svx_gpu_host_free(host);
    svx_octree_free(tree);
    free(want);
    free(got);
    printf("multi_gpu: ok (%u members, %zu hits per frame, %s)\n", n, hits, svx_version());
    return 0;
}
