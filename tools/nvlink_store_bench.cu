// How fast can the GPUs of one box store framebuffer pixels into GPU 0's memory from inside a kernel, as a function of the
// store pattern? One process, peer access, every GPU != 0 sends its interleaved row bands of a 3840x2160 plane set at once.
//   pattern 32 : a warp covers an 8x4 pixel tile, one 4-byte store per lane  -> four 32-byte segments per plane (the viewport
//                kernel's own stores)
//   pattern 64 : 16x2 tile                                                   -> two 64-byte segments
//   pattern 128: 32x1                                                        -> one 128-byte line
//   pattern 512: a warp writes 4 rows x 128 bytes with 16-byte stores (uint4 per lane; what a shared-memory staged CTA does)
// Prints inbound GB/s at GPU 0 per pattern and plane count (2 planes = 8 B/pixel wire format, 3 = 12 B/pixel).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/nvlink_store_bench tools/nvlink_store_bench.cu
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e = (x);                                                                   \
        if (e != cudaSuccess) {                                                                \
            std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e));                       \
            std::exit(1);                                                                      \
        }                                                                                      \
    } while (0)

constexpr int W = 3840, H = 2160, BAND = 8;

// local row lr of `rank` -> image row (interleaved bands of 8 rows)
__device__ __forceinline__ int image_row(int lr, int rank, int world) { return ((lr / BAND) * world + rank) * BAND + (lr % BAND); }

template <int LW>  // tile width 2^LW: 3 -> 8x4, 4 -> 16x2, 5 -> 32x1
__global__ void store_tiles(uint32_t* base, size_t plane_words, int planes, int rank, int world, int rows_local) {
    const int TW = 1 << LW, TH = 32 >> LW;
    const int tiles_x = W / TW, tiles_y = rows_local / TH;
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < tiles_x * tiles_y; t += warps) {
        const int tx = t % tiles_x, ty = t / tiles_x;
        const int x = tx * TW + (lane & (TW - 1)), lr = ty * TH + (lane >> LW);
        const int row = image_row(lr, rank, world);
        if (row >= H) continue;
        const size_t i = (size_t)row * W + x;
        for (int p = 0; p < planes; ++p) base[p * plane_words + i] = (uint32_t)i + p;
    }
}

// a warp writes a 32x4 pixel block as 4 rows x 128 bytes, 16 bytes per lane
__global__ void store_rows16(uint32_t* base, size_t plane_words, int planes, int rank, int world, int rows_local) {
    const int tiles_x = W / 32, tiles_y = rows_local / 4;
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < tiles_x * tiles_y; t += warps) {
        const int tx = t % tiles_x, ty = t / tiles_x;
        const int x = tx * 32 + (lane & 7) * 4, lr = ty * 4 + (lane >> 3);
        const int row = image_row(lr, rank, world);
        if (row >= H) continue;
        const size_t i = (size_t)row * W + x;
        for (int p = 0; p < planes; ++p)
            *reinterpret_cast<uint4*>(base + p * plane_words + i) = make_uint4((uint32_t)i, (uint32_t)i + 1, (uint32_t)i + 2, (uint32_t)i + 3 + p);
    }
}

int main(int argc, char** argv) {
    int n = 0;
    CK(cudaGetDeviceCount(&n));
    if (argc > 1) n = std::min(n, std::atoi(argv[1]));
    if (n < 2) {
        std::printf("{\"error\": \"needs at least 2 GPUs\"}\n");
        return 0;
    }
    const size_t plane_words = (size_t)W * H;
    uint32_t* target = nullptr;
    CK(cudaSetDevice(0));
    CK(cudaMalloc(&target, 3 * plane_words * 4));
    std::vector<cudaStream_t> streams(n);
    std::vector<cudaEvent_t> e0(n), e1(n);
    for (int d = 1; d < n; ++d) {
        CK(cudaSetDevice(d));
        int can = 0;
        CK(cudaDeviceCanAccessPeer(&can, d, 0));
        if (!can) {
            std::printf("{\"error\": \"no peer access %d -> 0\"}\n", d);
            return 0;
        }
        CK(cudaDeviceEnablePeerAccess(0, 0));
        CK(cudaStreamCreate(&streams[d]));
        CK(cudaEventCreate(&e0[d]));
        CK(cudaEventCreate(&e1[d]));
    }
    const int bands = H / BAND;
    std::printf("{\"gpus\": %d, \"pixels\": %d", n, W * H);
    for (int planes = 2; planes <= 3; ++planes)
        for (int pattern : {32, 64, 128, 512})
            for (int scope = 0; scope < 2; ++scope) {  // 0: every sender at once (inbound limit of GPU 0), 1: GPU 1 alone (one sender's limit)
                const int first = 1, last = scope == 0 ? n - 1 : 1;
                float worst = 0.0f;
                double bytes = 0;
                for (int rep = 0; rep < 4; ++rep) {
                    for (int d = first; d <= last; ++d) {
                        CK(cudaSetDevice(d));
                        int owned = 0;
                        for (int b = d; b < bands; b += n) ++owned;
                        const int rows_local = owned * BAND;
                        CK(cudaEventRecord(e0[d], streams[d]));
                        const int grid = 148 * 8;
                        if (pattern == 32) store_tiles<3><<<grid, 128, 0, streams[d]>>>(target, plane_words, planes, d, n, rows_local);
                        if (pattern == 64) store_tiles<4><<<grid, 128, 0, streams[d]>>>(target, plane_words, planes, d, n, rows_local);
                        if (pattern == 128) store_tiles<5><<<grid, 128, 0, streams[d]>>>(target, plane_words, planes, d, n, rows_local);
                        if (pattern == 512) store_rows16<<<grid, 128, 0, streams[d]>>>(target, plane_words, planes, d, n, rows_local);
                        CK(cudaEventRecord(e1[d], streams[d]));
                    }
                    worst = 0.0f;
                    bytes = 0;
                    for (int d = first; d <= last; ++d) {
                        CK(cudaSetDevice(d));
                        CK(cudaStreamSynchronize(streams[d]));
                        float ms = 0.0f;
                        CK(cudaEventElapsedTime(&ms, e0[d], e1[d]));
                        worst = ms > worst ? ms : worst;
                        int owned = 0;
                        for (int b = d; b < bands; b += n) ++owned;
                        bytes += (double)owned * BAND * W * 4 * planes;
                    }
                }
                std::printf(", \"planes%d_seg%d_%s\": {\"ms\": %.4f, \"inbound_gbs\": %.1f}", planes, pattern, scope == 0 ? "all_senders" : "one_sender", worst,
                            bytes / (worst * 1e-3) / 1e9);
            }
    std::printf("}\n");
    return 0;
}
