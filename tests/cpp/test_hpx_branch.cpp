// The HPXFFT_B200_WITH_HPX branches of the drop-in headers, compiled against tests/cpp/hpx_stub (the subset of the HPX API they
// use; HPX itself is absent from this image): vector_2d serialisation (core/include/hpxfft/util/vector_2d.hpp:73-85), the HPX
// bootstrap (hpx::get_locality_id / get_num_localities / collectives, core/src/distributed/loop.cpp:281-282,324-327) and -- with
// a GPU (argument "gpu") -- both agas clients returning hpx::future.
#include "check.hpp"
#include "hpxfft/distributed/agas.hpp"
#include "hpxfft/shared/agas.hpp"

#include <cstring>
#include <string>
#include <type_traits>

using hpxfft::shared::vector_2d;

int main(int argc, char **argv)
{
    static_assert(std::is_same_v<decltype(std::declval<hpxfft::shared::agas &>().fft_2d_r2c()), hpx::future<vector_2d>>, "hpx::future surface");
    static_assert(std::is_same_v<decltype(std::declval<hpxfft::distributed::agas &>().initialize(vector_2d(), "", "")), hpx::future<void>>,
                  "hpx::future surface");
    {   // serialisation round trip
        vector_2d v(3, 4, 0.0);
        for (std::size_t i = 0; i < 3; ++i)
            for (std::size_t j = 0; j < 4; ++j) v(i, j) = 10.0 * i + j;
        std::vector<char> bytes;
        hpx::serialization::output_archive oa(bytes);
        oa << v;
        REQUIRE(bytes.size() == 3 * sizeof(std::size_t) + 12 * sizeof(double));   // element-wise, like the reference
        vector_2d w;
        hpx::serialization::input_archive ia(bytes);
        ia >> w;
        REQUIRE(w == v);
    }
    {   // bootstrap through the HPX calls
        hpxfft::distributed::bootstrap boot;
        REQUIRE(boot.this_locality == 0 && boot.num_localities == 1);
        auto all = boot.all_gather("uid", std::string(128, 'x'));
        REQUIRE(all.size() == 1 && all[0] == std::string(128, 'x'));
    }
    if (argc > 1 && !std::strcmp(argv[1], "gpu"))
    {
        const double row[6] = {1.0, 2.0, 3.0, 4.0, 0.0, 0.0};
        vector_2d in(4, 6, 0.0), in2(4, 6, 0.0);
        for (std::size_t i = 0; i < 4; ++i)
            for (std::size_t j = 0; j < 6; ++j) in(i, j) = in2(i, j) = row[j];
        hpxfft::shared::agas a;
        a.initialize(std::move(in), "measure").get();
        hpx::future<vector_2d> f = a.fft_2d_r2c();
        vector_2d out = f.get();
        REQUIRE(out(0, 0) == 40.0 && out(0, 2) == -8.0 && out(0, 3) == 8.0 && out(0, 4) == -8.0 && out(1, 0) == 0.0);
        hpxfft::distributed::agas d;
        d.initialize(std::move(in2), "all_to_all", "estimate").get();
        vector_2d out2 = d.fft_2d_r2c().get();
        REQUIRE(out2 == out);
    }
    std::puts("test_hpx_branch ok");
    return 0;
}
