// hpx_stub (see README.md): locality queries of a single-locality run.  NOT HPX.
#pragma once
#include "future.hpp"
#include <cstddef>
#include <cstdint>
namespace hpx
{
namespace launch
{
struct sync_policy
{
};
inline constexpr sync_policy sync{};
}  // namespace launch
inline std::uint32_t get_locality_id() { return 0; }
inline std::uint32_t get_num_localities(launch::sync_policy) { return 1; }
}  // namespace hpx
