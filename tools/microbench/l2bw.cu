// L2 / HBM bandwidth probe: read, write and copy over buffers of increasing size (development aid).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void rd(const double2* __restrict__ p, size_t n, int reps, double* sink) {
    double acc = 0; size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < reps; ++r)
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += stride) {
            double2 v; asm volatile("ld.global.cg.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p + i));
            acc += v.x + v.y;
        }
    if (acc == 1.2345) *sink = acc;
}
__global__ void wr(double2* p, size_t n, int reps) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < reps; ++r)
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += stride) p[i] = make_double2(r, i);
}
__global__ void cp(const double2* __restrict__ a, double2* b, size_t n, int reps) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < reps; ++r)
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += stride) {
            double2 v; asm volatile("ld.global.cg.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(a + i));
            b[i] = v;
        }
}
int main() {
    double* sink; cudaMalloc(&sink, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (size_t mb : {8, 16, 32, 48, 64, 96, 128, 256, 1024}) {
        size_t n = mb * 1024 * 1024 / 16; double2 *a, *b; cudaMalloc(&a, n * 16); cudaMalloc(&b, n * 16);
        cudaMemset(a, 0, n * 16); cudaMemset(b, 0, n * 16);
        int reps = (int)(8192 / mb) + 1; float ms;
        for (int k = 0; k < 3; ++k) {
            const char* nm[3] = {"read", "write", "copy"};
            for (int w = 0; w < 2; ++w) {
                cudaEventRecord(e0);
                if (k == 0) rd<<<148 * 8, 256>>>(a, n, reps, sink);
                if (k == 1) wr<<<148 * 8, 256>>>(a, n, reps);
                if (k == 2) cp<<<148 * 8, 256>>>(a, b, n, reps);
                cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
            }
            double bytes = (double)n * 16 * reps * (k == 2 ? 2 : 1);
            printf("%5zu MB %-5s %8.1f GB/s\n", mb, nm[k], bytes / ms / 1e6);
        }
        cudaFree(a); cudaFree(b);
    }
    return 0;
}
