// granule.cu -- what HBM rate do the access PATTERNS of the FFT kernels allow, with the arithmetic taken out?
// Every pattern moves 2 GiB in + 2 GiB out in units of 128 KB; a unit is copied by one CTA pass
// (256 threads, 8 x 16-byte loads in flight per thread, several CTAs per SM).  G = contiguous granule in bytes.
//   copy      : contiguous -> contiguous                                     (the denominator)
//   rows(G)   : contiguous 128 KB row -> 128K/G segments of G bytes, stride n_units*G (row kernel's store)
//   colsA(G)  : 128K/G segments of G bytes strided by 128*G -> contiguous     (column level-A load)
//   colsB(G)  : contiguous -> segments of G bytes strided by pitch (~cy*16 B)*128   (column level-B store)
//   fused(G)  : colsA read pattern -> colsB write pattern                       (what the fused column kernel shows HBM)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o granule granule.cu ; run: ./granule
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

enum Pat { CONTIG = 0, ROWS = 1, COLSA = 2, COLSB = 3 };
constexpr unsigned long long UNIT = 128 * 1024;

// byte offset of segment i of unit u
__device__ __forceinline__ unsigned long long seg_addr(int pat, unsigned long long u, unsigned long long i, unsigned long long G,
                                                       unsigned long long n_units)
{
    const unsigned long long spu = UNIT / G; // segments per unit
    switch (pat) {
    case ROWS: return i * (n_units * G) + u * G;                                  // tile i, row u
    case COLSA: {                                                                 // unit = (strip ct, x2); segment i = x1
        const unsigned long long n2 = 128, per_strip = n2;                        // n2 units per strip
        const unsigned long long ct = u / per_strip, x2 = u % per_strip;
        return ct * (spu * n2 * G) + (i * n2 + x2) * G;
    }
    case COLSB: {                                                                 // unit = (strip ct, k1); segment i = k2, row kx = k1 + n1*k2
        const unsigned long long n1 = 128, tiles = (n_units * UNIT) / (spu * n1 * G); // strips
        const unsigned long long ct = u / n1, k1 = u % n1;
        const unsigned long long pitch = tiles * G + 16;                          // odd pitch like cy*16
        return (k1 + n1 * i) * pitch + ct * G;
    }
    default: return u * UNIT + i * G;
    }
}

__global__ void __launch_bounds__(256) copy_pattern(const char *__restrict__ src, char *__restrict__ dst, int rpat, int wpat,
                                                    unsigned long long G, unsigned long long n_units)
{
    const unsigned long long tps = G / 16;        // threads per segment
    for (unsigned long long u = blockIdx.x; u < n_units; u += gridDim.x) {
        // 8192 16-byte elements per unit, 256 threads -> 32 per thread, in 4 batches of 8
        for (int b = 0; b < 4; ++b) {
            double2 v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const unsigned long long idx = (unsigned long long) (b * 8 + e) * 256 + threadIdx.x; // element in unit
                const unsigned long long i = idx / tps, o = (idx % tps) * 16;
                const double2 *p = (const double2 *) (src + seg_addr(rpat, u, i, G, n_units) + o);
                asm volatile("ld.global.cs.v2.f64 {%0, %1}, [%2];" : "=d"(v[e].x), "=d"(v[e].y) : "l"(p));
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const unsigned long long idx = (unsigned long long) (b * 8 + e) * 256 + threadIdx.x;
                const unsigned long long i = idx / tps, o = (idx % tps) * 16;
                double2 *p = (double2 *) (dst + seg_addr(wpat, u, i, G, n_units) + o);
                asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v[e].x), "d"(v[e].y) : "memory");
            }
        }
    }
}

int main(int argc, char **argv)
{
    const unsigned long long n_units = 16384, bytes = n_units * UNIT; // 2 GiB each way
    char *src, *dst;
    const size_t slack = 64ull << 20;
    if (cudaMalloc(&src, bytes + slack) != cudaSuccess || cudaMalloc(&dst, bytes + slack) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMemset(src, 1, bytes + slack);
    cudaMemset(dst, 0, bytes + slack);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    struct Case { const char *name; int r, w; } cases[] = {{"copy", CONTIG, CONTIG}, {"rows", CONTIG, ROWS}, {"colsA", COLSA, CONTIG},
                                                           {"colsB", CONTIG, COLSB}, {"fused", COLSA, COLSB}};
    printf("pattern,granule_B,ctas_per_sm,ms,GBps\n");
    for (auto &c : cases)
        for (unsigned long long G : {64ull, 128ull, 256ull, 512ull, 1024ull, 2048ull})
            for (int cps : {1, 2, 4, 8}) {
                if (c.r == CONTIG && c.w == CONTIG && G != 256) continue;
                float best = 1e30f;
                for (int rep = 0; rep < 4; ++rep) {
                    cudaEventRecord(a);
                    copy_pattern<<<148 * cps, 256>>>(src, dst, c.r, c.w, G, n_units);
                    cudaEventRecord(b);
                    cudaEventSynchronize(b);
                    float ms;
                    cudaEventElapsedTime(&ms, a, b);
                    if (rep && ms < best) best = ms;
                }
                if (cudaGetLastError() != cudaSuccess) { printf("kernel error\n"); return 2; }
                printf("%s,%llu,%d,%.4f,%.1f\n", c.name, G, cps, best, 2.0 * bytes / best / 1e6);
            }
    return 0;
}
