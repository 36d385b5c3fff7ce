// hpxfft::distributed::agas -- client surface of core/include/hpxfft/distributed/agas.hpp:13-29:
//   future<void> initialize(vector_2d, COMM_FLAG, PLAN_FLAG);   future<vector_2d> fft_2d_r2c();
// The reference schedules one HPX action per row behind this client (core/src/distributed/agas.cpp:149-300);
// on a GPU the whole transform is a handful of kernel launches, so the client wraps distributed::loop: initialize()
// (plan creation, communicator bootstrap: host work) runs on a worker, fft_2d_r2c() only ENQUEUES the transform and the
// copy back and returns a future that a CUDA stream callback fulfils -- no thread waits on the GPU.
// hpx::future when built with HPX, std::future otherwise.
#ifndef HPXFFT_B200_DISTRIBUTED_AGAS_HPP
#define HPXFFT_B200_DISTRIBUTED_AGAS_HPP

#include "loop.hpp"

#include <memory>

#if defined(HPXFFT_B200_WITH_HPX)
#include <hpx/future.hpp>
#define HPXFFT_B200_FUTURE hpx::future
#define HPXFFT_B200_ASYNC(...) hpx::async(__VA_ARGS__)
#else
#include <future>
#define HPXFFT_B200_FUTURE std::future
#define HPXFFT_B200_ASYNC(...) std::async(std::launch::async, __VA_ARGS__)
#endif

namespace hpxfft::distributed
{
struct agas
{
    explicit agas() : impl_(std::make_shared<loop>()) {}

    HPXFFT_B200_FUTURE<vector_2d> fft_2d_r2c()
    {
        if (!impl_->has_plan())  // bad COMM_FLAG: the reference prints and returns the data untouched (distributed/loop.cpp:175-179)
        {
            auto impl = impl_;
            return HPXFFT_B200_ASYNC([impl]() { return impl->fft_2d_r2c(); });
        }
        return impl_->fft_2d_r2c_async();
    }

    HPXFFT_B200_FUTURE<void> initialize(vector_2d values_vec, const std::string COMM_FLAG, const std::string PLAN_FLAG)
    {
        hpxfft::util::check_plan_flag(PLAN_FLAG);
        auto impl = impl_;
        auto data = std::make_shared<vector_2d>(std::move(values_vec));
        return HPXFFT_B200_ASYNC([impl, data, COMM_FLAG, PLAN_FLAG]() { impl->initialize(std::move(*data), COMM_FLAG, PLAN_FLAG); });
    }

    real get_measurement(std::string name) { return impl_->get_measurement(std::move(name)); }

    ~agas() = default;

  private:
    std::shared_ptr<loop> impl_;
};
}  // namespace hpxfft::distributed
#endif
