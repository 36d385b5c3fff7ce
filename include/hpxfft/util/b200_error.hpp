// Maps C-ABI return codes onto the exception types the reference throws.
#ifndef HPXFFT_B200_ERROR_HPP
#define HPXFFT_B200_ERROR_HPP

#include "../../hpxfft_b200.h"

#include <stdexcept>
#include <string>

namespace hpxfft::util
{
// plan-flag check of core/include/hpxfft/util/adapter_fftw.hpp:22-44 (same message, same exception)
inline void check_plan_flag(const std::string &flag)
{
    if (flag != "estimate" && flag != "measure" && flag != "patient" && flag != "exhaustive")
        throw std::invalid_argument("Invalid FFTW plan flag string");
}

inline void b200_check(int rc)
{
    if (rc == HPXFFT_B200_OK) return;
    const std::string msg = hpxfft_b200_last_error();
    if (rc == HPXFFT_B200_EPLANFLAG) throw std::invalid_argument(msg);
    throw std::runtime_error("hpxfft_b200: " + msg);
}
}  // namespace hpxfft::util
#endif
