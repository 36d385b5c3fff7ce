"""Host-side mirrors of the reference classes: vector_2d unit tests (ports of
test/src/test_vector_2d.cpp) and error behaviour of the loop classes.  CPU only."""
import numpy as np
import pytest


def test_vector_2d_constant_initialisation(pkg):  # test_vector_2d.cpp:6-16
    vec = pkg.vector_2d(3, 3, 4.0)
    assert vec.n_row() == 3 and vec.n_col() == 3
    assert vec(0, 0) == 4.0 and vec(2, 2) == 4.0 and vec(1, 2) == 4.0
    assert vec.size() == 9


def test_vector_2d_access_out_of_range(pkg):  # test_vector_2d.cpp:18-23
    vec = pkg.vector_2d(3, 3, 1.0)
    with pytest.raises(RuntimeError):
        vec.at(3, 3)
    assert vec.at(2, 2) == 1.0


def test_vector_2d_compare(pkg):  # test_vector_2d.cpp:25-33
    a, b, c = pkg.vector_2d(2, 2, 5.0), pkg.vector_2d(2, 2, 5.0), pkg.vector_2d(2, 2, 6.0)
    assert a == b and not (a == c)
    assert not (a == pkg.vector_2d(2, 3, 5.0))


def test_vector_2d_row_major_layout(pkg):  # vector_2d.hpp:198-213
    v = pkg.vector_2d(2, 6)
    v.set(1, 2, 7.0)
    assert v.data().reshape(-1)[1 * 6 + 2] == 7.0 and v.row(1)[2] == 7.0
    assert v.data().flags["C_CONTIGUOUS"]


def test_invalid_plan_flag_raises(pkg):  # util/adapter_fftw.hpp:40-43 -> std::invalid_argument
    with pytest.raises(ValueError, match="Invalid FFTW plan flag string"):
        pkg.shared.loop().initialize(pkg.vector_2d(4, 6), "fast")
    with pytest.raises(ValueError):
        pkg.distributed.loop(device=-1).initialize(pkg.vector_2d(4, 6), "all_to_all", "fast")
    with pytest.raises(ValueError):
        pkg.distributed.agas(device=-1).initialize(pkg.vector_2d(4, 6), "all_to_all", "fast")


def test_invalid_comm_flag_prints_and_does_not_transform(pkg, capsys):  # distributed/loop.cpp:342-346,175-179
    fft = pkg.distributed.loop(device=-1)
    v = pkg.vector_2d.from_array(np.arange(24.0).reshape(4, 6))
    fft.initialize(v, "gather", "estimate")
    assert "Specify communication scheme: scatter or all_to_all" in capsys.readouterr().out
    out = fft.fft_2d_r2c()
    assert "Communication scheme not specified during initialization" in capsys.readouterr().out
    assert np.array_equal(out.data(), np.arange(24.0).reshape(4, 6))


def test_unknown_measurement_is_zero(pkg):  # shared/loop.cpp:192
    assert pkg.shared.loop().get_measurement("nonsense") == 0.0
    assert pkg.distributed.loop(device=-1).get_measurement("total") == 0.0


def test_fft_before_initialize_fails_loudly(pkg):
    with pytest.raises(RuntimeError):
        pkg.shared.loop().fft_2d_r2c_par()
