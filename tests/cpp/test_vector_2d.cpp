// Port of test/src/test_vector_2d.cpp (HPX-FFT) against the drop-in header.  Needs no GPU.
#include "check.hpp"
#include "hpxfft/util/vector_2d.hpp"
#include <utility>

int main()
{
    {   // "Vector 2D constant: Initialization"  (test_vector_2d.cpp:6-16)
        hpxfft::util::vector_2d<double> vec(3, 3, 4.0);
        REQUIRE(vec.n_row() == 3);
        REQUIRE(vec.n_col() == 3);
        REQUIRE(vec(0, 0) == 4.0);
        REQUIRE(vec(2, 2) == 4.0);
        REQUIRE(vec(1, 2) == 4.0);
        REQUIRE(vec.size() == 9);
    }
    {   // "Vector 2D: Access Out of Range"  (test_vector_2d.cpp:18-23)
        hpxfft::util::vector_2d<double> vec(3, 3, 1.0);
        REQUIRE_THROWS_AS(vec.at(3, 3), std::runtime_error);
    }
    {   // "Compare two Vector 2D instances"  (test_vector_2d.cpp:25-33)
        hpxfft::util::vector_2d<double> vec1(2, 2, 5.0), vec2(2, 2, 5.0), vec3(2, 2, 6.0);
        REQUIRE(vec1 == vec2);
        REQUIRE(!(vec1 == vec3));
    }
    {   // layout + move semantics the loop classes rely on
        hpxfft::util::vector_2d<double> v(2, 6);
        REQUIRE(v(1, 5) == 0.0);
        v(1, 2) = 7.0;
        REQUIRE(v.data()[1 * 6 + 2] == 7.0 && v.row(1)[2] == 7.0 && v.values_[8] == 7.0);
        hpxfft::util::vector_2d<double> w(std::move(v));
        REQUIRE(w.n_row_ == 2 && w.n_col_ == 6 && w.size_ == 12 && v.size() == 0 && v.values_ == nullptr);
        hpxfft::util::vector_2d<double> c = w;
        REQUIRE(c == w && c.data() != w.data());
    }
    std::puts("test_vector_2d ok");
    return 0;
}
