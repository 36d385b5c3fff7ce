// layout.cuh -- HBM data layouts of the 2-D r2c pipeline (see DESIGN.md "Data layout in HBM").
//
//   V  : the caller's slab, row-major, nxl rows x (ny+2) doubles == nxl x cy complex (cy = ny/2+1)
//        -- byte-identical to hpxfft::util::vector_2d (core/include/hpxfft/util/vector_2d.hpp:198-213).
//   I  : intermediate between the row pass and the column pass, *column-tiled*:
//        I_q[r][ct][j][c]  r = source rank, ct = tile of CW adjacent ky owned by rank q,
//        j = local row of rank r, c = column inside the tile.  Every (r, ct) block is one contiguous
//        run of nxl*CW complex, so (i) the row kernel writes CW*16-byte segments, (ii) the exchange
//        r -> q is ONE contiguous message, and (iii) the column kernel reads aligned CW*16-byte
//        segments with no transpose anywhere.  This replaces split_vec + communicate + transpose_y_to_x
//        of the reference (core/src/distributed/loop.cpp:19-27, 39-84, 87-106) and, for one rank,
//        transpose_shared_y_to_x (core/src/shared/loop.cpp:18-25).
//   S  : scratch between the two levels of a long column FFT, S[ct][k1][x2][c].
#pragma once
#include "fft_device.cuh"

namespace hpxfft_b200 {

#ifndef HPXFFT_B200_CW
#define HPXFFT_B200_CW 16
#endif
constexpr int CW = HPXFFT_B200_CW;   // columns per tile: 16 complex = 256-byte segments
constexpr int CW_SHIFT = CW == 16 ? 4 : (CW == 32 ? 5 : (CW == 64 ? 6 : -1));
static_assert(CW_SHIFT > 0, "CW must be 16, 32 or 64");
constexpr int MAXP = 16; // max ranks

// Destination of the row pass: column k of local row j.
struct RowDst {
    cd *base[MAXP];             // per destination rank q: start of the [ct][j][c] block written by this rank
    unsigned long long tile_stride; // nxl * CW
    unsigned cy;                // ny/2 + 1
    unsigned wq0;               // columns owned by every rank but the last: cy / P
    unsigned P;
};

__device__ __forceinline__ cd *rowdst_ptr(const RowDst &d, unsigned j, unsigned k)
{
    unsigned q = 0, kl = k;
    if (d.P > 1) {
        q = k / d.wq0;
        if (q >= d.P) q = d.P - 1;
        kl = k - q * d.wq0;
    }
    return d.base[q] + (unsigned long long) (kl / CW) * d.tile_stride + (unsigned long long) j * CW + (kl % CW);
}

// The column pass's view of I on the owning rank.
struct InterView {
    const cd *base;
    unsigned long long rank_stride; // ntiles * nxl * CW
    unsigned long long tile_stride; // nxl * CW
    unsigned nxl;
    unsigned shift;                 // log2(nxl) when nxl is a power of two, else 0xffffffff
};

__host__ __device__ inline unsigned pow2_shift(unsigned v)
{
    if (v == 0 || (v & (v - 1))) return 0xffffffffu;
    unsigned s = 0;
    while ((1u << s) < v) ++s;
    return s;
}

__device__ __forceinline__ const cd *inter_ptr(const InterView &v, unsigned x, unsigned ct, unsigned c)
{
    unsigned r, j;
    if (v.shift != 0xffffffffu) {
        r = x >> v.shift;
        j = x & (v.nxl - 1);
    } else {
        r = x / v.nxl;
        j = x - r * v.nxl;
    }
    return v.base + (unsigned long long) r * v.rank_stride + (unsigned long long) ct * v.tile_stride +
           (unsigned long long) j * CW + c;
}

// Destination of the column pass: local column kl (< w) of global row kx.
struct ColDst {
    cd *base[MAXP];         // per destination rank r (owner of row kx)
    unsigned pitch[MAXP];   // complex elements per destination row
    int col0[MAXP];         // destination column of local column 0 (negative inside a chunked send buffer)
    unsigned nxl;           // rows per rank
    unsigned w;             // valid local columns on this rank
    unsigned shift;         // log2(nxl) when nxl is a power of two, else 0xffffffff
    // odd-radix pre-stage (nx = vt * q, kernels_generic.cuh): the power-of-two kernels then run on vt "virtual strips" per
    // real strip, virtual strip v = ct * vt + k1 holding the length-q transform whose output row k2 is global row k1 + vt * k2
    unsigned vt;            // 1 = no pre-stage
};

// real strip and first output row of a (virtual) strip
__device__ __forceinline__ void coldst_strip(const ColDst &d, unsigned strip, unsigned &ct, unsigned &row0)
{
    if (d.vt > 1) {
        ct = strip / d.vt;
        row0 = strip - ct * d.vt;
    } else {
        ct = strip;
        row0 = 0;
    }
}

__device__ __forceinline__ cd *coldst_ptr(const ColDst &d, unsigned kx, unsigned kl)
{
    unsigned r, j;
    if (d.shift != 0xffffffffu) {
        r = kx >> d.shift;
        j = kx & (d.nxl - 1);
    } else {
        r = kx / d.nxl;
        j = kx - r * d.nxl;
    }
    return d.base[r] + ((long long) j * (long long) d.pitch[r] + (long long) d.col0[r] + (long long) kl);
}

// chunked unpack of exchange #2 (pipelined NCCL transport): one dense [nxl][wc_q] block per source rank
struct UnpackChunk {
    unsigned long long src_off[MAXP]; // element offset of the block inside the receive buffer
    unsigned dst_col[MAXP];           // first destination column in V
    unsigned wc[MAXP];                // block width (0 = nothing from this rank)
};

}  // namespace hpxfft_b200
