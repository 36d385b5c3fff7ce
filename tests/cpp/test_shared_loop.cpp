// Port of test/src/test_shared_loop.cpp (HPX-FFT): 4x6 rows [1,2,3,4,0,0] -> row 0 [40,0,-8,8,-8,0], exact ==.
// Needs a GPU.
#include "check.hpp"
#include "hpxfft/shared/loop.hpp"
#include <string>

int main()
{
    const std::size_t n_row = 4, n_col = 6;
    for (int variant = 0; variant < 2; ++variant)
    {
        hpxfft::shared::vector_2d values_vec(n_row, n_col, 0.0);
        for (std::size_t i = 0; i < n_row; ++i)
        {
            values_vec(i, 0) = 1.0;
            values_vec(i, 1) = 2.0;
            values_vec(i, 2) = 3.0;
            values_vec(i, 3) = 4.0;
        }
        hpxfft::shared::vector_2d expected_output(n_row, n_col, 0.0);
        expected_output(0, 0) = 40.0;
        expected_output(0, 2) = -8.0;
        expected_output(0, 3) = 8.0;
        expected_output(0, 4) = -8.0;

        hpxfft::shared::loop fft;
        std::string plan_flag = "estimate";
        fft.initialize(std::move(values_vec), plan_flag);
        hpxfft::shared::vector_2d out = variant == 0 ? fft.fft_2d_r2c_par() : fft.fft_2d_r2c_seq();
        auto total = fft.get_measurement(std::string("total"));
        auto flops = fft.get_measurement(std::string("plan_flops"));
        REQUIRE(total >= 0.0);
        REQUIRE(flops > 0.0);
        REQUIRE(out == expected_output);
        REQUIRE(fft.get_measurement("unknown") == 0.0);
    }
    {   // util/adapter_fftw.hpp:40-43
        hpxfft::shared::loop fft;
        REQUIRE_THROWS_AS(fft.initialize(hpxfft::shared::vector_2d(4, 6, 0.0), "fast"), std::invalid_argument);
        hpxfft::shared::loop fft2;
        fft2.initialize(hpxfft::shared::vector_2d(4, 6, 0.0), "measure");
        REQUIRE_THROWS_AS(fft2.write_plans_to_file("/nonexistent_dir/plan.txt"), std::runtime_error);
    }
    std::puts("test_shared_loop ok");
    return 0;
}
