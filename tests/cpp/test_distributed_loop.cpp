// Port of test/src/test_distributed_loop.cpp and test_distributed_agas.cpp (HPX-FFT).  Run as one
// process (the only configuration the reference's own test pins) or SPMD with RANK / WORLD_SIZE set:
// the expected data is non-zero on locality 0 only (test_distributed_loop.cpp:33-39).  Needs GPUs.
#include "check.hpp"
#include "hpxfft/distributed/agas.hpp"
#include "hpxfft/distributed/loop.hpp"
#include <string>

using hpxfft::distributed::vector_2d;

static vector_2d make_input(std::size_t n_x_local, std::size_t n_col)
{
    vector_2d values_vec(n_x_local, n_col, 0.0);
    for (std::size_t i = 0; i < n_x_local; ++i)
    {
        values_vec(i, 0) = 1.0;
        values_vec(i, 1) = 2.0;
        values_vec(i, 2) = 3.0;
        values_vec(i, 3) = 4.0;
    }
    return values_vec;
}

int main(int argc, char **argv)
{
    const std::string comm = argc > 1 ? argv[1] : "scatter";
    const std::size_t n_row = 4, n_col = 6;
    hpxfft::distributed::loop fft;
    const std::size_t this_locality = fft.this_locality(), num_localities = fft.num_localities();
    const std::size_t n_x_local = n_row / num_localities;
    vector_2d expected_output(n_x_local, n_col, 0.0);
    if (this_locality == 0)
    {
        expected_output(0, 0) = 40.0;
        expected_output(0, 2) = -8.0;
        expected_output(0, 3) = 8.0;
        expected_output(0, 4) = -8.0;
    }
    fft.initialize(make_input(n_x_local, n_col), comm, "estimate");
    vector_2d values_vec = fft.fft_2d_r2c();
    REQUIRE(fft.get_measurement(std::string("total")) >= 0.0);
    REQUIRE(values_vec == expected_output);

    if (num_localities == 1)
    {   // agas client surface, plan flag "measure" (test_distributed_agas.cpp)
        hpxfft::distributed::agas a;
        a.initialize(make_input(n_row, n_col), "all_to_all", "measure").get();
        vector_2d out = a.fft_2d_r2c().get();
        REQUIRE(out == expected_output);
        // unknown COMM flag: message, no exception, data returned untouched (distributed/loop.cpp:342-346)
        hpxfft::distributed::loop bad;
        bad.initialize(make_input(n_row, n_col), "gather", "estimate");
        vector_2d same = bad.fft_2d_r2c();
        REQUIRE(same == make_input(n_row, n_col));
    }
    std::printf("test_distributed_loop ok (locality %zu of %zu, %s)\n", this_locality, num_localities, comm.c_str());
    return 0;
}
