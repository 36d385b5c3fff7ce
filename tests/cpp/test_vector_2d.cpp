// vector_2d drop-in header: the three behaviours the reference's unit tests pin
// (test/src/test_vector_2d.cpp:6-33 -- constant fill + sizes, at() range check, exact ==) plus the
// layout / ownership rules the loop classes rely on.  Runs without a GPU (falls back to new[]).
#include "check.hpp"
#include "hpxfft/util/vector_2d.hpp"

#include <utility>

using grid = hpxfft::util::vector_2d<double>;

static void filled_constructor_reports_shape_and_value()
{
    const double fill = 4.0;
    grid g(3, 3, fill);
    REQUIRE(g.n_row() == 3 && g.n_col() == 3 && g.size() == 9);
    for (std::size_t i = 0; i < 3; ++i)
        for (std::size_t j = 0; j < 3; ++j) REQUIRE(g(i, j) == fill);
    grid z(2, 5);  // value-initialised
    for (double x : z) REQUIRE(x == 0.0);
}

static void at_checks_the_flat_index()
{
    grid g(3, 3, 1.0);
    REQUIRE(g.at(2, 2) == 1.0);
    REQUIRE_THROWS_AS(g.at(3, 3), std::runtime_error);
    const grid &cg = g;
    REQUIRE_THROWS_AS(cg.at(9, 0), std::runtime_error);
}

static void equality_is_exact_and_shape_aware()
{
    grid a(2, 2, 5.0), same(2, 2, 5.0), other(2, 2, 6.0), wide(2, 3, 5.0);
    REQUIRE(a == same);
    REQUIRE(!(a == other));
    REQUIRE(!(a == wide));
}

static void storage_is_row_major_and_moves_cheaply()
{
    grid v(2, 6);
    v(1, 2) = 7.0;
    REQUIRE(v.data()[1 * 6 + 2] == 7.0 && v.row(1)[2] == 7.0 && v.values_[8] == 7.0);
    const double *before = v.data();
    grid w(std::move(v));
    REQUIRE(w.data() == before && w.n_row_ == 2 && w.n_col_ == 6 && w.size_ == 12);
    REQUIRE(v.size() == 0 && v.values_ == nullptr);
    grid c = w;  // deep copy
    REQUIRE(c == w && c.data() != w.data());
    grid d;
    d = std::move(c);
    REQUIRE(d == w && c.size() == 0);
}

int main()
{
    filled_constructor_reports_shape_and_value();
    at_checks_the_flat_index();
    equality_is_exact_and_shape_aware();
    storage_is_row_major_and_moves_cheaply();
    std::puts("test_vector_2d ok");
    return 0;
}
