// hpx_stub (see README.md): hpx::future / hpx::promise / hpx::async on top of <future>.  NOT HPX.
#pragma once
#include <future>
#include <utility>
namespace hpx
{
template <class T> using future = std::future<T>;
template <class T> using promise = std::promise<T>;
template <class F, class... A> auto async(F &&f, A &&...a)
{
    return std::async(std::launch::async, std::forward<F>(f), std::forward<A>(a)...);
}
}  // namespace hpx
