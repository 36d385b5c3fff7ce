// launch_generic.cu -- lengths n = t * q with an odd factor t (kernels_generic.cuh): plan-side set-up of the power-of-two
// stage that follows the odd-radix column pre-stage, and the launchers.
#include "kernels_generic.cuh"
#include "launch_util.h"

#include <cstdlib>

namespace hpxfft_b200 {

void gen_factor(size_t n, unsigned &t, unsigned &q, unsigned &lg)
{
    q = 1;
    lg = 0;
    while (n % 2 == 0 && n > 0) {
        n /= 2;
        q *= 2;
        ++lg;
    }
    t = (unsigned) n;
}

bool gen_rows_supported(size_t m)
{
    unsigned t, q, lg;
    gen_factor(m, t, q, lg);
    return t > 1 && t < GEN_TMAX_ROWS && m <= 8192;
}

bool gen_cols_supported(size_t nx)
{
    unsigned t, q, lg;
    gen_factor(nx, t, q, lg);
    return t > 1 && t <= GEN_TMAX_COLS && q <= (1u << 18);
}

template <int TT, int NT>
int launch_rows_mixed_t(const hpxfft_b200_plan *p, const RowDst &dst, unsigned nrows, const cd *V, unsigned pitch, const RowsMixedArgs &a)
{
    const size_t m = (size_t) a.t * a.q;
    const size_t smem = (m + a.t + a.q) * sizeof(cd);
    // the attribute is set once per device for the largest row this kernel accepts (the footprint varies with m): m <= 8192, q <= m / 3
    if (int rc = ensure_smem(rows_mixed_kernel<TT, NT>, (size_t) (8192 + GEN_TMAX_ROWS + 8192 / 3 + 1) * sizeof(cd), p->device)) return rc;
    int per_sm = 1;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rows_mixed_kernel<TT, NT>, NT, smem));
    if (per_sm < 1) per_sm = 1;
    const unsigned cap = (unsigned) (p->sm_count * per_sm);
    const unsigned grid = nrows < cap ? nrows : cap;
    rows_mixed_kernel<TT, NT><<<grid, NT, smem, p->stream>>>(V, pitch, nrows, dst, a, p->tw_row);
    CU(cudaGetLastError());
    return 0;
}

int launch_rows_mixed(const hpxfft_b200_plan *p, const RowDst &dst, unsigned nrows, const cd *V, unsigned pitch, size_t m)
{
    RowsMixedArgs a;
    gen_factor(m, a.t, a.q, a.lg);
    if (a.t <= 3) return launch_rows_mixed_t<4, 512>(p, dst, nrows, V, pitch, a);
    if (a.t <= 7) return launch_rows_mixed_t<8, 512>(p, dst, nrows, V, pitch, a);
    if (a.t <= 15) return launch_rows_mixed_t<16, 512>(p, dst, nrows, V, pitch, a);
    return launch_rows_mixed_t<32, 256>(p, dst, nrows, V, pitch, a);
}

// plan-view of a power-of-two column stage: a column FFT of length q (= 2^lg) over `nstrips` strips, with its own twiddle tables,
// scratch ring and counters; shares the parent's device and stream
int make_col_stage(const hpxfft_b200_plan *parent, unsigned q, unsigned nstrips, hpxfft_b200_plan **out)
{
    unsigned lg = 0;
    while ((1u << lg) < q) ++lg;
    hpxfft_b200_plan *s = new hpxfft_b200_plan();
    *out = s;
    s->device = parent->device;
    s->sm_count = parent->sm_count;
    s->stream = parent->stream; // shared, not owned
    s->nx = s->nxl = q;
    s->ntiles = nstrips;
    s->w = nstrips * CW;
    // same decomposition rules as the main plan (plan.cu: choose_col_split)
    if (q <= 256) {
        s->two_level = false;
        s->n1 = q;
        s->n2 = 1;
    } else {
        s->two_level = true;
        s->n1 = 1u << ((lg + 1) / 2);
        s->n2 = 1u << (lg / 2);
    }
    std::vector<double2> tq;
    make_twiddles(tq, q);
    CU(cudaMalloc(&s->tw_col, tq.size() * sizeof(double2)));
    CU(cudaMemcpy(s->tw_col, tq.data(), tq.size() * sizeof(double2), cudaMemcpyHostToDevice));
    if (s->two_level) {
        std::vector<double2> w2((size_t) q);
        for (size_t x2 = 0; x2 < s->n2; ++x2)
            for (size_t k1 = 0; k1 < s->n1; ++k1) w2[x2 * s->n1 + k1] = tq[(k1 * x2) % q];
        CU(cudaMalloc(&s->tw_il, w2.size() * sizeof(double2)));
        CU(cudaMemcpy(s->tw_il, w2.data(), w2.size() * sizeof(double2), cudaMemcpyHostToDevice));
        const char *e = getenv("HPXFFT_B200_FUSED");
        s->fused = !(e && e[0] == '0') && fused_pair_exists(s->n1, s->n2);
        size_t bytesS = (size_t) s->ntiles * q * CW * sizeof(cd);
        if (s->fused) {
            int bps = 1;
            if (int rc = fused_blocks_per_sm(s->n1, s->n2, 1, &bps)) return rc;
            s->fused_grid = (unsigned) (bps * s->sm_count);
            const unsigned per_group = s->n1 + s->n2;
            s->lag = (unsigned) ((3 * (size_t) s->fused_grid / 2 + per_group - 1) / per_group) + 1;
            s->nslot = 2 * s->lag + 1;
            if (s->nslot > s->ntiles) s->nslot = s->ntiles > 0 ? s->ntiles : 1;
            bytesS = (size_t) s->nslot * q * CW * sizeof(cd);
            CU(cudaMalloc(&s->ctl, (1 + 2 * (size_t) s->ntiles) * sizeof(unsigned)));
        }
        CU(cudaMalloc(&s->S, bytesS));
    }
    return 0;
}

void free_col_stage(hpxfft_b200_plan *s)
{
    if (!s) return;
    cudaFree(s->tw_col);
    cudaFree(s->tw_il);
    cudaFree(s->S);
    cudaFree(s->ctl);
    s->stream = nullptr;
    delete s;
}

// runs the stage on the first `nstrips` strips of `iv` (fused persistent kernel when the pair exists)
int run_col_stage(const hpxfft_b200_plan *s, const InterView &iv, const ColDst &out, unsigned nstrips, int *launches)
{
    if (s->fused) {
        if (launches) *launches += 1;
        return launch_cols_fused(s, iv, out, 0u, nstrips);
    }
    return launch_cols(s, iv, out, nstrips, s->S, (unsigned) s->nx, s->n1, s->n2, s->two_level, launches, nullptr);
}

int gen_setup_col_stage(hpxfft_b200_plan *p)
{
    unsigned t, q, lg;
    gen_factor(p->nx, t, q, lg);
    p->gen_ct = t;
    p->gen_cq = q;
    CU(cudaMalloc(&p->S1, (size_t) p->ntiles * p->nx * CW * sizeof(cd)));
    if (q == 1) return 0;
    return make_col_stage(p, q, p->ntiles * t, &p->colsub);
}

void gen_free_col_stage(hpxfft_b200_plan *p)
{
    cudaFree(p->S1);
    p->S1 = nullptr;
    free_col_stage(p->colsub);
    p->colsub = nullptr;
}

int launch_cols_mixed(const hpxfft_b200_plan *p, const InterView &in, const ColDst &out, int *launches)
{
    ColsOddArgs a;
    a.t = p->gen_ct;
    a.q = p->gen_cq;
    a.nx = (unsigned) p->nx;
    a.S1 = p->S1;
    // x2 values per CTA: keep the tile within 48 KB and at least 256 outputs per CTA
    unsigned xb = (48u * 1024u - a.t * (unsigned) sizeof(cd)) / (a.t * CW * (unsigned) sizeof(cd));
    if (xb < 1) xb = 1;
    if (xb > 16) xb = 16;
    if (xb > a.q) xb = a.q;
    a.xb = xb;
    const size_t smem = ((size_t) a.t * xb * CW + a.t) * sizeof(cd); // <= 48 KB by construction of xb: no opt-in needed
    cols_odd_kernel<<<dim3((a.q + xb - 1) / xb, p->ntiles), GEN_THREADS, smem, p->stream>>>(in, out, a, p->tw_col, a.q == 1);
    CU(cudaGetLastError());
    if (launches) *launches += 1;
    if (a.q == 1) return 0;
    const hpxfft_b200_plan *s = p->colsub;
    InterView iv;
    iv.base = p->S1;
    iv.nxl = a.q;
    iv.shift = pow2_shift(iv.nxl);
    iv.tile_stride = (unsigned long long) a.q * CW;
    iv.rank_stride = 0;
    ColDst o2 = out;
    o2.vt = a.t;
    return run_col_stage(s, iv, o2, s->ntiles, launches);
}

}  // namespace hpxfft_b200
