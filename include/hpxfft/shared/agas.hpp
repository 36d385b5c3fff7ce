// hpxfft::shared::agas -- client surface of core/include/hpxfft/shared/agas.hpp:13-27:
//   future<void> initialize(vector_2d, PLAN_FLAG);   future<vector_2d> fft_2d_r2c();
// The reference puts the four loop phases behind an HPX component and schedules them as actions
// (core/src/shared/agas.cpp:47-86); here the client wraps shared::loop: initialize() (plan creation: host work) runs on a
// worker, fft_2d_r2c() only ENQUEUES the transform and the copy back and returns a future that a CUDA stream callback
// fulfils.  hpx::future when built with HPX, std::future otherwise.
#ifndef HPXFFT_B200_SHARED_AGAS_HPP
#define HPXFFT_B200_SHARED_AGAS_HPP

#include "loop.hpp"

#include <memory>

#if defined(HPXFFT_B200_WITH_HPX)
#include <hpx/future.hpp>
#define HPXFFT_B200_SHARED_ASYNC(...) hpx::async(__VA_ARGS__)
#else
#include <future>
#define HPXFFT_B200_SHARED_ASYNC(...) std::async(std::launch::async, __VA_ARGS__)
#endif

namespace hpxfft::shared
{
struct agas
{
    explicit agas() : impl_(std::make_shared<loop>()) {}

    hpxfft::util::future<vector_2d> fft_2d_r2c() { return impl_->fft_2d_r2c_async(); }

    hpxfft::util::future<void> initialize(vector_2d values_vec, const std::string PLAN_FLAG)
    {
        hpxfft::util::check_plan_flag(PLAN_FLAG);  // std::invalid_argument at the call site, like the by-value action argument
        auto impl = impl_;
        auto data = std::make_shared<vector_2d>(std::move(values_vec));
        return HPXFFT_B200_SHARED_ASYNC([impl, data, PLAN_FLAG]() { impl->initialize(std::move(*data), PLAN_FLAG); });
    }

    real get_measurement(std::string name) { return impl_->get_measurement(std::move(name)); }
    void set_device(int device) { impl_->set_device(device); }

    ~agas() = default;

  private:
    std::shared_ptr<loop> impl_;
};
}  // namespace hpxfft::shared
#endif
