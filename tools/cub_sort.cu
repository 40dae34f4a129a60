// Comparison only (BASELINE.json north_star: "CUB timed only as a comparison"): cub::DeviceRadixSort::SortPairs
// on (uint32 key, uint32 index) pairs with the product's key widths. NOT part of the product; never linked into
// libpbf_b200.so.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/cub_sort tools/cub_sort.cu
// Usage: tools/cub_sort N CELLS [reps]   -> prints one JSON line with the best and median time in microseconds.
#include <cub/cub.cuh>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void fill(uint32_t* k, uint32_t* v, size_t n, uint32_t cells) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t x = (uint32_t)i * 2654435761u + 12345u;   // a hash: uniform keys in [0, cells)
    x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13;
    k[i] = x % cells;
    v[i] = (uint32_t)i;
}

int main(int argc, char** argv) {
    const size_t n = argc > 1 ? strtoull(argv[1], 0, 10) : 1048576;
    const uint32_t cells = argc > 2 ? (uint32_t)strtoul(argv[2], 0, 10) : 552960;
    const int reps = argc > 3 ? atoi(argv[3]) : 20;
    int bits = 0;
    while ((1ull << bits) < cells) bits++;
    uint32_t *k0, *k1, *v0, *v1;
    CK(cudaMalloc(&k0, n * 4)); CK(cudaMalloc(&k1, n * 4)); CK(cudaMalloc(&v0, n * 4)); CK(cudaMalloc(&v1, n * 4));
    void* tmp = nullptr;
    size_t tmp_bytes = 0;
    CK(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k0, k1, v0, v1, n, 0, bits));
    CK(cudaMalloc(&tmp, tmp_bytes));
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    std::vector<float> us;
    for (int r = 0; r < reps + 3; r++) {
        fill<<<(unsigned)((n + 255) / 256), 256>>>(k0, v0, n, cells);
        CK(cudaEventRecord(a));
        CK(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k0, k1, v0, v1, n, 0, bits));
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms;
        CK(cudaEventElapsedTime(&ms, a, b));
        if (r >= 3) us.push_back(ms * 1e3f);
    }
    std::sort(us.begin(), us.end());
    printf("{\"impl\": \"cub::DeviceRadixSort::SortPairs\", \"n\": %zu, \"cells\": %u, \"key_bits\": %d, \"best_us\": %.1f, \"median_us\": %.1f, \"temp_bytes\": %zu}\n",
           n, cells, bits, us.front(), us[us.size() / 2], tmp_bytes);
    return 0;
}
