// Futures fulfilled by a CUDA stream callback (hpxfft_b200_on_complete -> cudaLaunchHostFunc): no thread blocks on the GPU.
// hpx::future when built with HPX (-DHPXFFT_B200_WITH_HPX), std::future otherwise -- the agas client surfaces of the
// reference return hpx::future (core/include/hpxfft/shared/agas.hpp:19-24, core/include/hpxfft/distributed/agas.hpp:21-26).
#ifndef HPXFFT_B200_STREAM_FUTURE_HPP
#define HPXFFT_B200_STREAM_FUTURE_HPP

#include "b200_error.hpp"

#include <exception>
#include <memory>
#include <type_traits>
#include <utility>

#if defined(HPXFFT_B200_WITH_HPX)
#include <hpx/future.hpp>
namespace hpxfft::util
{
template <class T> using future = hpx::future<T>;
template <class T> using promise = hpx::promise<T>;
}  // namespace hpxfft::util
#else
#include <future>
namespace hpxfft::util
{
template <class T> using future = std::future<T>;
template <class T> using promise = std::promise<T>;
}  // namespace hpxfft::util
#endif

namespace hpxfft::util
{
// Enqueues `finish` behind everything already on the plan's stream; the returned future becomes ready when it has run.
// `finish` runs on a CUDA-internal thread: it may touch host memory but must not call CUDA or this library.
template <class T, class F> future<T> when_stream_reaches(hpxfft_b200_plan *plan, F finish)
{
    struct state
    {
        promise<T> p;
        F f;
    };
    auto *st = new state{promise<T>(), std::move(finish)};
    future<T> fut = st->p.get_future();
    const int rc = hpxfft_b200_on_complete(
        plan,
        [](void *user)
        {
            std::unique_ptr<state> s(static_cast<state *>(user));
            try
            {
                if constexpr (std::is_void_v<T>)
                {
                    s->f();
                    s->p.set_value();
                }
                else
                    s->p.set_value(s->f());
            }
            catch (...)
            {
                s->p.set_exception(std::current_exception());
            }
        },
        st);
    if (rc != HPXFFT_B200_OK)
    {
        delete st;
        b200_check(rc);
    }
    return fut;
}
}  // namespace hpxfft::util
#endif
