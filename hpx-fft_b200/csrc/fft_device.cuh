// fft_device.cuh -- register-level FP64 building blocks shared by the row (r2c) and column (c2c)
// kernels: complex helpers, constant-twiddle multiplies and fully unrolled radix-2^k DIF butterflies.
//
// These replace what FFTW's codelets do inside fftw_execute_dft / fftw_execute_dft_r2c
// (reference call sites core/src/util/adapter_fftw.cpp:14,34).  Forward sign: exp(-2 pi i jk/n).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace hpxfft_b200 {

typedef double2 cd;

__device__ __forceinline__ cd cadd(cd a, cd b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cd csub(cd a, cd b) { return make_double2(a.x - b.x, a.y - b.y); }
// Diagnostic builds (never shipped, see tools/diag_builds.sh): HPXFFT_B200_DIAG_NOMATH removes the FP64 work so that
// a kernel's memory / shared-memory / barrier skeleton can be timed alone (results are garbage).
__device__ __forceinline__ cd cmul(cd a, cd b)
{
#ifdef HPXFFT_B200_DIAG_NOMATH
    return a;
#else
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
#endif
}
// a * conj(b)
__device__ __forceinline__ cd cmulc(cd a, cd b)
{
#ifdef HPXFFT_B200_DIAG_NOMATH
    return a;
#else
    return make_double2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
#endif
}
__device__ __forceinline__ cd cconj(cd a) { return make_double2(a.x, -a.y); }

__host__ __device__ constexpr int ilog2(int v) { return v <= 1 ? 0 : 1 + ilog2(v >> 1); }
__host__ __device__ constexpr int bitrev(int i, int bits)
{
    int r = 0;
    for (int b = 0; b < bits; ++b) r |= ((i >> b) & 1) << (bits - 1 - b);
    return r;
}

// multiply by exp(-2 pi i idx/32), idx in [0,16).  idx is a compile-time constant after unrolling,
// so the branches fold away and the table entries become immediates.
__device__ __forceinline__ cd mulw32(cd a, int idx)
{
    constexpr double C[16] = {1.0,
                              0.98078528040323044913,
                              0.92387953251128675613,
                              0.83146961230254523708,
                              0.70710678118654752440,
                              0.55557023301960222474,
                              0.38268343236508977173,
                              0.19509032201612826785,
                              0.0,
                              -0.19509032201612826785,
                              -0.38268343236508977173,
                              -0.55557023301960222474,
                              -0.70710678118654752440,
                              -0.83146961230254523708,
                              -0.92387953251128675613,
                              -0.98078528040323044913};
    constexpr double S[16] = {0.0,
                              0.19509032201612826785,
                              0.38268343236508977173,
                              0.55557023301960222474,
                              0.70710678118654752440,
                              0.83146961230254523708,
                              0.92387953251128675613,
                              0.98078528040323044913,
                              1.0,
                              0.98078528040323044913,
                              0.92387953251128675613,
                              0.83146961230254523708,
                              0.70710678118654752440,
                              0.55557023301960222474,
                              0.38268343236508977173,
                              0.19509032201612826785};
    constexpr double R2 = 0.70710678118654752440;
#ifdef HPXFFT_B200_DIAG_NOMATH
    return a;
#endif
    if (idx == 0) return a;
    if (idx == 8) return make_double2(a.y, -a.x);
    if (idx == 4) return make_double2(R2 * (a.x + a.y), R2 * (a.y - a.x));
    if (idx == 12) return make_double2(R2 * (a.y - a.x), -R2 * (a.x + a.y));
    // (a.x + i a.y)(c - i s) = (a.x c + a.y s) + i (a.y c - a.x s)
    return make_double2(a.x * C[idx] + a.y * S[idx], a.y * C[idx] - a.x * S[idx]);
}

// In-place forward DFT of R = 2^k points held in registers, radix-2 decimation in frequency.
// Output is left bit-reversed:  v[i] = X[bitrev(i)].
template <int R>
__device__ __forceinline__ void fft_dif(cd (&v)[R])
{
    static_assert(R >= 1 && R <= 32 && (R & (R - 1)) == 0, "radix must be a power of two <= 32");
    constexpr int L = ilog2(R);
#ifdef HPXFFT_B200_DIAG_NOMATH
    return;
#endif
#pragma unroll
    for (int s = 0; s < L; ++s) {
        const int h = (R / 2) >> s;
#pragma unroll
        for (int b = 0; b < R; b += 2 * h) {
#pragma unroll
            for (int q = 0; q < h; ++q) {
                const cd a = v[b + q], c = v[b + q + h];
                v[b + q] = cadd(a, c);
                v[b + q + h] = mulw32(csub(a, c), q * (16 / h));
            }
        }
    }
}

// 128-bit read-only load of a twiddle (tables are immutable for the life of a plan)
__device__ __forceinline__ cd ldtw(const cd *__restrict__ tab, unsigned idx) { return __ldg(tab + idx); }

// streaming (evict-first) global accesses for the big arrays: every element is touched once per pass
__device__ __forceinline__ cd ld_stream(const cd *p)
{
    cd r;
    asm volatile("ld.global.cs.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream(cd *p, cd v)
{
#ifdef HPXFFT_B200_DIAG_STORE_CG
    asm volatile("st.global.cg.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory"); // diagnostic: normal L2 eviction priority
#else
    asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
#endif
}
// L2-only (bypass L1) accesses for data produced by other CTAs of the same launch
__device__ __forceinline__ cd ld_cg(const cd *p)
{
    cd r;
    asm volatile("ld.global.cg.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_cg(cd *p, cd v)
{
    asm volatile("st.global.cg.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}

// 256-bit streaming store of two adjacent complex values (sm_100: STG.E.EF.ENL2.256); p must be 32-byte aligned
__device__ __forceinline__ void st_stream_pair(cd *p, cd v0, cd v1)
{
    asm volatile("st.global.cs.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(v0.x), "d"(v0.y), "d"(v1.x), "d"(v1.y) : "memory");
}

// L2 eviction-priority hints (createpolicy).  evict_last: data that is re-read soon by the same kernel (per-CTA scratch,
// a row that the next sub-FFT reads again) must survive the flood of streaming traffic that passes through L2 in the
// meantime; evict_first on the last use hands the lines back.
__device__ __forceinline__ unsigned long long l2_policy_evict_last()
{
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ unsigned long long l2_policy_evict_first()
{
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
// L2 prefetch issued by ONE thread through the bulk-copy (TMA) unit: unlike prefetch.global.L2 it does not occupy a slot of
// the load/store queue per 32 bytes.  p and bytes must be multiples of 16.
__device__ __forceinline__ void l2_prefetch_bulk(const void *p, unsigned bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ cd ld_cg_hint(const cd *p, unsigned long long policy)
{
    cd r;
    asm volatile("ld.global.cg.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(r.x), "=d"(r.y) : "l"(p), "l"(policy));
    return r;
}
__device__ __forceinline__ void st_cg_hint(cd *p, cd v, unsigned long long policy)
{
    asm volatile("st.global.cg.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(v.x), "d"(v.y), "l"(policy) : "memory");
}

}  // namespace hpxfft_b200
