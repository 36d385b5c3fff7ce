// kernels_rows_v2.cuh -- r2c FFT of rows with ny = 16384 (m = 8192 complex): the row length of BASELINE config 2.
//
// Same role, same input/output contract and same final pass as rows_r2c_kernel<8192,1> (kernels_rows.cuh); what
// changes is how the first two thirds of the transform move through shared memory.
//
// m = 16 x 512.  The row is staged in NATURAL order (cp.async, overlapped with the previous row's tail), which makes
// the 16 stride-16 sub-sequences y_j1[j2] = z[j1 + 16 j2] visible as they are.  Each of the 8 warps owns two of
// them and transforms them IN PLACE with a decimation-in-frequency 512-point FFT (radix 16, then radix 32: one lane
// per butterfly, 32 points per lane) -- reads and writes of a butterfly hit the same addresses, a sub-sequence never
// leaves its warp, so these two passes need no CTA barrier at all, only a __syncwarp between them.  Warps drift apart
// and the shared-memory phases of one overlap the FP64 phases of another; the Stockham version serialised every
// pass behind two CTA barriers and measured the same time with all its data in L2 as with HBM traffic (diagnostic
// build "wrap", profiles/r2_diag_wrap_16384.json) -- i.e. it was bound by that serialisation, not by memory.
// The final pass (radix 16 across the sub-sequences on paired columns + Hermitian split + transposed store) is the
// old one with new shared-memory addresses.  CTA barriers per row: 3 instead of 6.
//
// Shared-memory address of element p (16-byte units): P(p) = p ^ (((p >> 4) ^ (p >> 9)) & 7) -- an XOR swizzle of the
// position inside the 128-byte bank row.  Conflict-free for all access patterns (8 consecutive lanes differ in bits 4-6 or
// 9-11 of p, or are contiguous) and, unlike padding, it keeps the staged row aligned to the bank rows.
#pragma once
#include "kernels_rows.cuh"

namespace hpxfft_b200 {

namespace rv2 {
constexpr int M = 8192, PP = 512, JW = PP / 2 + 1; // columns of the final pass; table width (mirror column uses the conjugate)
constexpr int LP = M;
// pencil | twA[s*32+u] = w_512^(u s) | tw2[r*JW+j] = w_M^(r j) | tw3[j] = w_n^j
constexpr int TW_ENTRIES = 512 + 16 * JW + JW;
constexpr size_t SMEM = (size_t) (LP + TW_ENTRIES) * sizeof(cd);
__device__ __forceinline__ int pad(int p) { return p ^ (((p >> 4) ^ (p >> 9)) & 7); }
}  // namespace rv2

// PF: one lane per warp asks the bulk-copy unit to pull the NEXT row of this CTA into L2 as soon as the current one has landed, so
// that the cp.async refill two barriers later is served at L2 latency
// ILV: the 32 cp.async of the refill are issued in four groups between the register-only steps of the tail (twiddles, the two
// butterflies) instead of one burst: a burst backs up the load/store queue and stalls all eight warps at the same time
template <bool FASTADDR, bool PF = false, bool ILV = false>
__global__ void __launch_bounds__(ROW_THREADS, 1)
    rows_r2c_v2_kernel(const cd *__restrict__ V, unsigned pitch, unsigned nxl, RowDst dst, const cd *__restrict__ tw)
{
    using namespace rv2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr unsigned MM = (unsigned) M;
    const int lt = threadIdx.x, warp = lt >> 5, lane = lt & 31;
    cd *sm = reinterpret_cast<cd *>(smem_raw);
    cd *twA = sm + LP;
    cd *tw2 = twA + 512;
    cd *tw3 = tw2 + 16 * JW;

    // tables from tw = w_n^i, n = 2M (published by the first barrier of the row loop)
    for (int i = lt; i < 512; i += ROW_THREADS) {
        const int s = i >> 5, u = i & 31;
        twA[i] = ldtw(tw, (unsigned) (u * s) * (unsigned) (2 * M / 512));
    }
    for (int i = lt; i < 16 * JW; i += ROW_THREADS) {
        const int r = i / JW, j = i - r * JW;
        tw2[i] = ldtw(tw, 2u * (unsigned) (r * j));
    }
    for (int i = lt; i < JW; i += ROW_THREADS) tw3[i] = ldtw(tw, (unsigned) i);

    auto row_ptr = [&](unsigned row) -> const cd * {
#ifdef HPXFFT_B200_DIAG_WRAP
        row &= 63u;
#endif
        return V + (unsigned long long) row * pitch;
    };
    // coalesced staging: thread lt copies elements lt + 256 e
    auto stage_part = [&](const cd *zrow, int e0, int e1) {
#pragma unroll
        for (int e = e0; e < e1; ++e) {
            const int p = lt + e * ROW_THREADS;
            cp_async16(sm + pad(p), zrow + p);
        }
    };
    // Eight copies that must not be issued before `after[8..15]` exist.  ptxas schedules the inlined cp.async freely among
    // register-only arithmetic (operand constraints of the asm statement do not survive into PTX), and packs the groups back into
    // one burst; a true data dependence is the only thing it honours, so the source address of copy i carries a term that is zero
    // at run time (pitch < 2^31) but that the compiler cannot fold: (high word of after[8 + i].x) & (pitch >> 31).
    auto stage_after = [&](const cd *zrow, int e0, const cd (&after)[16]) {
        const unsigned never = pitch >> 31;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int p = lt + (e0 + i) * ROW_THREADS;
            cp_async16(sm + pad(p), zrow + p + ((unsigned) __double2hiint(after[8 + i].x) & never));
        }
    };
    auto stage = [&](const cd *zrow) { stage_part(zrow, 0, ROW_PT); };
    if (blockIdx.x < nxl) stage(row_ptr(blockIdx.x));

    for (unsigned row0 = blockIdx.x; row0 < nxl; row0 += gridDim.x) {
#ifdef HPXFFT_B200_DIAG_WRAP
        const unsigned row = row0 & 63u;
#else
        const unsigned row = row0;
#endif
        cp_async_wait_all();
        __syncthreads(); // (1) the whole row has landed and is visible to every warp
        if constexpr (PF) {
            if (lane == 0 && row0 + gridDim.x < nxl)
                l2_prefetch_bulk(row_ptr(row0 + gridDim.x) + warp * (M / 8), (unsigned) (M / 8 * sizeof(cd)));
        }

        // ---- pass A: radix 16 over j2 = u + 32 r for the two sub-sequences of this warp, in place ----
        {
            const int u = lane;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int j1 = 2 * warp + h;
                // element j1 + 16 u + 512 r sits at (that & ~7) | ((j1 ^ u ^ r) & 7)
                cd *base = sm + ((j1 & 8) + 16 * u);
                const int x = (j1 ^ u) & 7;
                cd a[16];
#pragma unroll
                for (int r = 0; r < 16; ++r) a[r] = base[512 * r + (x ^ (r & 7))];
                fft_dif<16>(a);
#pragma unroll
                for (int s = 0; s < 16; ++s) {
                    cd o = a[bitrev(s, 4)];
                    if (s) o = cmul(o, twA[s * 32 + u]);
                    base[512 * s + (x ^ (s & 7))] = o;
                }
            }
        }
        __syncwarp(); // pass B reads what the other lanes of this warp just wrote
        // ---- pass B: radix 32 over u for fixed s: one lane per (sub-sequence, s), in place ----
        {
            const int h = lane >> 4, s = lane & 15, j1 = 2 * warp + h;
            cd *base = sm + ((j1 & 8) + 512 * s); // element j1 + 16 u + 512 s sits at (that & ~7) | ((j1 ^ u ^ s) & 7)
            const int x = (j1 ^ s) & 7;
            cd c[32];
#pragma unroll
            for (int u = 0; u < 32; ++u) c[u] = base[16 * u + (x ^ (u & 7))];
            fft_dif<32>(c);
#pragma unroll
            for (int t = 0; t < 32; ++t) base[16 * t + (x ^ (t & 7))] = c[bitrev(t, 5)]; // F_j1[s + 16 t]
        }
        __syncthreads(); // (2) all 16 sub-spectra are complete

        // ---- final pass: radix 16 over j1 on the paired columns jA = lt and jB = PP - lt (lt = 0: PP/2) ----
        const int jA = lt, jB = lt ? PP - lt : PP / 2;
        cd A[16], B[16];
        {
            // column k2 = s + 16 t, sub-sequence r: element r + 16 t + 512 s sits at (that & ~7) | ((r ^ t ^ s) & 7)
            const cd *pa = sm + (16 * (jA >> 4) + 512 * (jA & 15));
            const cd *pb = sm + (16 * (jB >> 4) + 512 * (jB & 15));
            const int xa = ((jA >> 4) ^ jA) & 7, xb = ((jB >> 4) ^ jB) & 7;
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                A[r] = pa[(r & 8) + (xa ^ (r & 7))];
                B[r] = pb[(r & 8) + (xb ^ (r & 7))];
            }
        }
        __syncthreads(); // (3) the pencil buffer is dead: refill it with the next row while this one finishes in registers
        const bool refill = row0 + gridDim.x < nxl;
        const cd *znext = row_ptr(refill ? row0 + gridDim.x : row0);
        if (refill) {
            if constexpr (ILV)
                stage_part(znext, 0, ROW_PT / 4);
            else
                stage(znext);
        }

        // column jB got the conjugate twiddles (w_M^(r (PP - j)) = w_16^r conj(w_M^(r j))): its natural output s sits at
        // butterfly output (s + 1) & 15
        if (lt != 0) {
#pragma unroll
            for (int r = 1; r < 16; ++r) {
                const cd t = tw2[r * JW + jA];
                A[r] = cmul(A[r], t);
                B[r] = cmulc(B[r], t);
            }
        } else {
#pragma unroll
            for (int r = 1; r < 16; ++r) B[r] = mulw32(B[r], r); // jA = 0, jB = PP/2: w_M^(r PP/2) = w_32^r
        }
        // ILV: twiddles | group 2 (needs the twiddled A[8..15]) | butterfly A | group 3 (needs its outputs) | butterfly B | group 4
        // (needs its outputs) | Hermitian split and stores
        if constexpr (ILV) {
            if (refill) stage_after(znext, ROW_PT / 4, A);
        }
        fft_dif<16>(A);
        if constexpr (ILV) {
            if (refill) stage_after(znext, ROW_PT / 2, A);
        }
        fft_dif<16>(B);
        if constexpr (ILV) {
            if (refill) stage_after(znext, 3 * ROW_PT / 4, B);
        }
        if (lt != 0) {
            const cd wb = tw3[jA]; // w_n^jA
            cd *pk = nullptr, *pm = nullptr;
            long long step = 0;
            if constexpr (FASTADDR) {
                const unsigned k0 = (unsigned) jA, m0 = MM - k0;
                pk = dst.base[0] + (unsigned long long) (k0 >> CW_SHIFT) * dst.tile_stride + (unsigned long long) row * CW + (k0 & (unsigned) (CW - 1));
                pm = dst.base[0] + (unsigned long long) (m0 >> CW_SHIFT) * dst.tile_stride + (unsigned long long) row * CW + (m0 & (unsigned) (CW - 1));
                step = (long long) (PP >> CW_SHIFT) * (long long) dst.tile_stride;
            }
#pragma unroll
            for (int s = 0; s < 16; ++s) {
                cd xk, xmk;
                // pair Z[kA] with Z[M - kA] = natural output 15-s of column jB = butterfly output (16-s)&15
                herm_pair(A[bitrev(s, 4)], B[bitrev((16 - s) & 15, 4)], mulw32(wb, s), xk, xmk);
                if constexpr (FASTADDR) {
                    st_stream(pk + s * step, xk);
                    st_stream(pm - s * step, xmk);
                } else {
                    const unsigned kA = (unsigned) (jA + s * PP);
                    st_stream(rowdst_ptr(dst, row, kA), xk);
                    st_stream(rowdst_ptr(dst, row, MM - kA), xmk);
                }
            }
        } else {
            // columns 0 and PP/2 are their own partners
            const cd z0 = A[0];
            st_stream(rowdst_ptr(dst, row, 0u), make_double2(z0.x + z0.y, 0.0));
            st_stream(rowdst_ptr(dst, row, MM), make_double2(z0.x - z0.y, 0.0));
#pragma unroll
            for (int s = 1; s < 8; ++s) {
                const unsigned k = (unsigned) (s * PP);
                cd xk, xmk;
                herm_pair(A[bitrev(s, 4)], A[bitrev(16 - s, 4)], mulw32(make_double2(1.0, 0.0), s), xk, xmk);
                st_stream(rowdst_ptr(dst, row, k), xk);
                st_stream(rowdst_ptr(dst, row, MM - k), xmk);
            }
            st_stream(rowdst_ptr(dst, row, (unsigned) (8 * PP)), cconj(A[bitrev(8, 4)]));
            const cd wh = tw3[PP / 2]; // w_n^(PP/2)
#pragma unroll
            for (int s = 0; s < 8; ++s) {
                const unsigned k = (unsigned) (PP / 2 + s * PP);
                cd xk, xmk;
                herm_pair(B[bitrev(s, 4)], B[bitrev(15 - s, 4)], mulw32(wh, s), xk, xmk);
                st_stream(rowdst_ptr(dst, row, k), xk);
                st_stream(rowdst_ptr(dst, row, MM - k), xmk);
            }
        }
    }
}

}  // namespace hpxfft_b200
