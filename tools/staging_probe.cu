// Throughput of StagedUploader (csrc/host_staging.h) alone: 1 GiB of pageable memory to the device, by thread count.
// Build: nvcc -O3 -std=c++17 -o tools/staging_probe.bin tools/staging_probe.cu -lpthread
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include "../acoustic_image_generation_b200/csrc/host_staging.h"

int main() {
    const size_t bytes = size_t(1) << 30;
    char* src = static_cast<char*>(malloc(bytes));
    memset(src, 1, bytes);
    void* dst; cudaMalloc(&dst, bytes);
    cudaStream_t s; cudaStreamCreate(&s);
    auto now = [] { return std::chrono::steady_clock::now(); };
    {
        cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s); cudaStreamSynchronize(s);
        auto t0 = now();
        cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s); cudaStreamSynchronize(s);
        double dt = std::chrono::duration<double>(now() - t0).count();
        printf("driver pageable path: %.1f GB/s\n", bytes / dt / 1e9);
    }
    for (int threads : {1, 2, 3, 4, 6, 8, 12}) {
        aig::StagedUploader up;
        if (!up.start(threads)) { printf("start failed\n"); return 1; }
        up.upload(dst, src, bytes, s); cudaStreamSynchronize(s);
        auto t0 = now();
        for (int rep = 0; rep < 3; ++rep) up.upload(dst, src, bytes, s);
        cudaStreamSynchronize(s);
        double dt = std::chrono::duration<double>(now() - t0).count() / 3;
        // the same in 64 MiB uploads, as stream_host_rows issues them
        auto t1 = now();
        for (size_t off = 0; off < bytes; off += size_t(64) << 20) up.upload(static_cast<char*>(dst) + off, src + off, size_t(64) << 20, s);
        cudaStreamSynchronize(s);
        double dt64 = std::chrono::duration<double>(now() - t1).count();
        printf("staged, %2d threads: %.1f GB/s in one upload, %.1f GB/s as 16 uploads of 64 MiB\n", threads, bytes / dt / 1e9, bytes / dt64 / 1e9);
        // one 16-frame call's worth (56.6 MB), synchronised each time, with a pause in between like separate API calls
        const size_t small = size_t(16) * 3538944;
        double total = 0;
        for (int rep = 0; rep < 10; ++rep) {
            std::this_thread::sleep_for(std::chrono::milliseconds(2));
            auto t2 = now();
            up.upload(dst, src, small, s);
            cudaStreamSynchronize(s);
            total += std::chrono::duration<double>(now() - t2).count();
        }
        printf("            56.6 MB uploads, synchronised: %.3f ms each = %.1f GB/s\n", total / 10 * 1e3, small / (total / 10) / 1e9);
    }
    return 0;
}
