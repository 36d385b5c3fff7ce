// hpxfft::shared::loop drop-in: the reference's known-answer case (test/src/test_shared_loop.cpp:15-34,53 --
// four rows [1,2,3,4,0,0] transform to row 0 = [40,0,-8,8,-8,0], everything else 0, compared with ==),
// through both entry points, plus the error behaviour of the class.  Needs a GPU.
#include "check.hpp"
#include "hpxfft/shared/loop.hpp"

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

static void known_answer(bool sequential_entry_point)
{
    hpxfft::shared::loop fft;
    fft.initialize(golden_input(), "estimate");
    vector_2d out = sequential_entry_point ? fft.fft_2d_r2c_seq() : fft.fft_2d_r2c_par();
    REQUIRE(out == golden_output());                       // exact, like the reference's REQUIRE(out2 == expected_output)
    REQUIRE(fft.get_measurement("total") >= 0.0);
    REQUIRE(fft.get_measurement("plan_flops") > 0.0);
    REQUIRE(fft.get_measurement("no such key") == 0.0);    // std::map::operator[] behaviour, shared/loop.cpp:192
}

static void error_behaviour()
{
    hpxfft::shared::loop bad_flag;
    REQUIRE_THROWS_AS(bad_flag.initialize(vector_2d(4, 6, 0.0), "fast"), std::invalid_argument);  // adapter_fftw.hpp:40-43
    hpxfft::shared::loop not_initialised;
    REQUIRE_THROWS_AS(not_initialised.fft_2d_r2c_par(), std::runtime_error);
    hpxfft::shared::loop ok;
    ok.initialize(vector_2d(4, 6, 0.0), "measure");
    REQUIRE_THROWS_AS(ok.write_plans_to_file("/nonexistent_dir/plan.txt"), std::runtime_error);    // shared/loop.cpp:198-201
}

int main()
{
    known_answer(false);
    known_answer(true);
    error_behaviour();
    std::puts("test_shared_loop ok");
    return 0;
}
