// agas client surfaces (core/include/hpxfft/shared/agas.hpp:13-27, core/include/hpxfft/distributed/agas.hpp:13-29):
// initialize() -> future<void>, fft_2d_r2c() -> future<vector_2d>; the known-answer case of
// test/src/test_shared_agas.cpp / test_distributed_agas.cpp (plan flag "measure"), compared with ==.  The transform future is
// fulfilled by a CUDA stream callback; the test also overlaps two clients.  Needs a GPU.
#include "check.hpp"
#include "hpxfft/distributed/agas.hpp"
#include "hpxfft/shared/agas.hpp"

#include <string>

using hpxfft::shared::vector_2d;

static vector_2d golden_input()
{
    const double row[6] = {1.0, 2.0, 3.0, 4.0, 0.0, 0.0};
    vector_2d v(4, 6, 0.0);
    for (std::size_t i = 0; i < v.n_row(); ++i)
        for (std::size_t j = 0; j < v.n_col(); ++j) v(i, j) = row[j];
    return v;
}

static vector_2d golden_output()
{
    const double row0[6] = {40.0, 0.0, -8.0, 8.0, -8.0, 0.0};
    vector_2d v(4, 6, 0.0);
    for (std::size_t j = 0; j < v.n_col(); ++j) v(0, j) = row0[j];
    return v;
}

int main()
{
    {
        hpxfft::shared::agas fft;
        fft.initialize(golden_input(), "measure").get();
        auto fut = fft.fft_2d_r2c();
        vector_2d out = fut.get();
        REQUIRE(out == golden_output());
        REQUIRE(fft.get_measurement("total") >= 0.0);
        REQUIRE_THROWS_AS(fft.initialize(vector_2d(4, 6, 0.0), "fast"), std::invalid_argument);
    }
    {
        // two clients in flight at once: both futures are pending while the host is free
        hpxfft::shared::agas a, b;
        auto ia = a.initialize(vector_2d(512, 1026, 1.0), "estimate"), ib = b.initialize(golden_input(), "estimate");
        ia.get();
        ib.get();
        auto fa = a.fft_2d_r2c();
        auto fb = b.fft_2d_r2c();
        vector_2d ob = fb.get(), oa = fa.get();
        REQUIRE(ob == golden_output());
        REQUIRE(oa(0, 0) == 512.0 * 1024.0);   // DC bin of an all-ones 512 x 1024 array
        REQUIRE(oa(1, 0) == 0.0 && oa(0, 2) == 0.0);
    }
    {
        hpxfft::distributed::agas fft;   // one locality
        fft.initialize(golden_input(), "scatter", "measure").get();
        vector_2d out = fft.fft_2d_r2c().get();
        REQUIRE(out == golden_output());
    }
    std::puts("test_agas ok");
    return 0;
}
