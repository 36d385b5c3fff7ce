// hpx_stub (see README.md): communicator + all_gather of a single-locality run.  NOT HPX.
#pragma once
#include "../future.hpp"
#include <cstddef>
#include <string>
#include <vector>
namespace hpx::collectives
{
struct num_sites_arg
{
    std::size_t n;
    explicit num_sites_arg(std::size_t v) : n(v) {}
};
struct this_site_arg
{
    std::size_t i;
    explicit this_site_arg(std::size_t v) : i(v) {}
};
struct communicator
{
    std::string name;
    std::size_t sites, site;
};
inline communicator create_communicator(char const *basename, num_sites_arg n, this_site_arg i) { return {basename, n.n, i.i}; }
template <class T> hpx::future<std::vector<T>> all_gather(communicator const &c, T &&local)
{
    std::vector<T> all(c.sites);
    all[c.site] = std::forward<T>(local);
    hpx::promise<std::vector<T>> p;
    p.set_value(std::move(all));
    return p.get_future();
}
}  // namespace hpx::collectives
