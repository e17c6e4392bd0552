// cub_scan_bar.cu -- MEASUREMENT ONLY (not product, not linked into libxtb200.so).
// Times CUB's DeviceScan::InclusiveSum (decoupled look-back, register tiles) and a plain
// device-to-device copy on this box, as the bar for xtb_scan's flat cumsum.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/_build/cub_scan_bar tools/cub_scan_bar.cu
#include <cstdio>
#include <cstdlib>
#include <cub/device/device_scan.cuh>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <class T> int run(const char* name, size_t n) {
    T *in, *out;
    CK(cudaMalloc(&in, n * sizeof(T)));
    CK(cudaMalloc(&out, n * sizeof(T)));
    CK(cudaMemset(in, 0, n * sizeof(T)));
    void* tmp = nullptr;
    size_t tmp_bytes = 0;
    cub::DeviceScan::InclusiveSum(tmp, tmp_bytes, in, out, (int64_t) n);
    CK(cudaMalloc(&tmp, tmp_bytes));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) cub::DeviceScan::InclusiveSum(tmp, tmp_bytes, in, out, (int64_t) n);
    CK(cudaDeviceSynchronize());
    const int iters = 20;
    cudaEventRecord(e0);
    for (int i = 0; i < iters; ++i) cub::DeviceScan::InclusiveSum(tmp, tmp_bytes, in, out, (int64_t) n);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= iters;
    printf("cub InclusiveSum %-4s n=%zu  %.4f ms  %.1f GB/s\n", name, n, ms, 2.0 * n * sizeof(T) / ms / 1e6);
    cudaEventRecord(e0);
    for (int i = 0; i < iters; ++i) cudaMemcpyAsync(out, in, n * sizeof(T), cudaMemcpyDeviceToDevice);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= iters;
    printf("cudaMemcpy D2D   %-4s n=%zu  %.4f ms  %.1f GB/s\n", name, n, ms, 2.0 * n * sizeof(T) / ms / 1e6);
    cudaFree(in); cudaFree(out); cudaFree(tmp);
    return 0;
}

int main() {
    if (run<float>("f32", (size_t) 1 << 26)) return 1;
    if (run<int>("i32", (size_t) 1 << 26)) return 1;
    if (run<double>("f64", (size_t) 1 << 25)) return 1;
    if (run<float>("f32", (size_t) 1 << 28)) return 1;
    return 0;
}
