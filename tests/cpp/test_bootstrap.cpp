// File-rendezvous bootstrap of hpxfft::distributed::loop (include/hpxfft/distributed/bootstrap.hpp):
// N processes with RANK / WORLD_SIZE exchange byte strings in locality order, twice (two generations),
// the way the NCCL id and the IPC handles travel when HPX is not available.  Needs no GPU.
#include "check.hpp"
#include "hpxfft/distributed/bootstrap.hpp"

#include <string>

int main()
{
    hpxfft::distributed::bootstrap boot;
    const std::size_t me = boot.this_locality, n = boot.num_localities;
    REQUIRE(me < n);
    std::string uid(128, static_cast<char>('a' + me));
    auto ids = boot.all_gather("uid", uid);
    REQUIRE(ids.size() == n);
    for (std::size_t r = 0; r < n; ++r) REQUIRE(ids[r] == std::string(128, static_cast<char>('a' + r)));
    std::string handles(64 * 2, static_cast<char>('A' + me));
    auto all = boot.all_gather("ipc", handles);
    for (std::size_t r = 0; r < n; ++r) REQUIRE(all[r] == std::string(128, static_cast<char>('A' + r)));
    REQUIRE(boot.local_device() >= 0);
    std::printf("test_bootstrap ok (locality %zu of %zu)\n", me, n);
    return 0;
}
