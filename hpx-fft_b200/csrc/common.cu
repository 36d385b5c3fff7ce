// common.cu -- error string, twiddle generation, per-device kernel attribute bookkeeping.
#include "internal.h"

#include <cmath>
#include <cstdarg>
#include <cstdio>

namespace hpxfft_b200 {

namespace {
thread_local char g_err[512] = "";
}

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

const char *last_error_string() { return g_err; }

// exp(-2 pi i k / n) rounded from long double; exact on the axes and diagonals so that small
// integer inputs (the reference's 4x4 known-answer test) transform exactly.
void make_twiddles(std::vector<double2> &t, size_t n)
{
    t.resize(n);
    const long double PI_L = 3.14159265358979323846264338327950288L;
    for (size_t k = 0; k < n; ++k) {
        // reduce to the first octant, evaluate there, map back by symmetry
        const size_t k8 = (8 * k) / n;           // octant 0..7
        const bool on_oct = (8 * k) % n == 0;
        long double c, s; // cos, sin of 2 pi k / n
        if (on_oct) {
            static const long double r2 = 0.70710678118654752440084436210484903928L;
            const long double C[8] = {1, r2, 0, -r2, -1, -r2, 0, r2};
            const long double S[8] = {0, r2, 1, r2, 0, -r2, -1, -r2};
            c = C[k8];
            s = S[k8];
        } else {
            const long double a = 2.0L * PI_L * (long double) k / (long double) n;
            c = cosl(a);
            s = sinl(a);
        }
        t[k] = make_double2((double) c, (double) (-s));
    }
}

}  // namespace hpxfft_b200
