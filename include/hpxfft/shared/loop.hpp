// hpxfft::shared::loop -- drop-in for core/include/hpxfft/shared/loop.hpp + core/src/shared/loop.cpp.
// Header-only wrapper over libhpxfft_b200.so (include/hpxfft_b200.h); needs neither HPX nor FFTW.
//
//   hpxfft::shared::loop fft;                       // examples/hpxfft/shared_loop_2d.cpp:44-55
//   fft.initialize(std::move(values_vec), "estimate");
//   values_vec = fft.fft_2d_r2c_par();              // or fft_2d_r2c_seq(): same GPU kernels
//   fft.get_measurement("total");                   // seconds, keys of core/src/shared/loop.cpp:106-110,182,188
#ifndef HPXFFT_B200_SHARED_LOOP_HPP
#define HPXFFT_B200_SHARED_LOOP_HPP

#include "../util/b200_error.hpp"
#include "../util/stream_future.hpp"
#include "../util/vector_2d.hpp"

#include <string>
#include <utility>

typedef double real;

namespace hpxfft::shared
{
using vector_2d = hpxfft::util::vector_2d<real>;

struct loop
{
  public:
    loop() = default;
    loop(const loop &) = delete;
    loop &operator=(const loop &) = delete;

    // takes the array by value and keeps it (core/src/shared/loop.cpp:161); plans; uploads
    void initialize(vector_2d values_vec, const std::string PLAN_FLAG)
    {
        hpxfft::util::check_plan_flag(PLAN_FLAG);
        reset();
        values_vec_ = std::move(values_vec);
        hpxfft::util::b200_check(hpxfft_b200_create(
            &plan_, values_vec_.n_row(), values_vec_.n_col(), 0, 1, device_, nullptr, PLAN_FLAG.c_str(), nullptr));
        hpxfft::util::b200_check(hpxfft_b200_upload(plan_, values_vec_.data()));
    }

    // extension: the data already lives on the GPU.  `device_slab` is a device pointer to n_row x n_col doubles in the
    // vector_2d layout; nothing crosses PCIe.  Pair with fft_2d_r2c_device().
    void initialize_device(const real *device_slab, std::size_t n_row, std::size_t n_col, const std::string PLAN_FLAG)
    {
        hpxfft::util::check_plan_flag(PLAN_FLAG);
        reset();
        hpxfft::util::b200_check(hpxfft_b200_create(&plan_, n_row, n_col, 0, 1, device_, nullptr, PLAN_FLAG.c_str(), nullptr));
        hpxfft::util::b200_check(hpxfft_b200_upload(plan_, device_slab));  // device-to-device (unified addressing)
    }
    // transforms and writes the result to `device_out` (may equal the input pointer); reusable: call again after another
    // hpxfft_b200_upload / initialize_device
    void fft_2d_r2c_device(real *device_out)
    {
        if (!plan_) throw std::runtime_error("hpxfft::shared::loop: initialize() has not been called");
        hpxfft::util::b200_check(hpxfft_b200_execute(plan_));
        hpxfft::util::b200_check(hpxfft_b200_download(plan_, device_out));
    }

    vector_2d fft_2d_r2c_par() { return run(); }
    // the reference's _seq differs only in CPU scheduling; on the GPU both run the same kernels
    vector_2d fft_2d_r2c_seq() { return run(); }
    vector_2d fft_2d_r2c() { return run(); }

    // extension (agas client surface): the transform and the copy back are only enqueued; the future is fulfilled by a
    // stream callback when the result has landed in the (page-locked) vector_2d storage
    hpxfft::util::future<vector_2d> fft_2d_r2c_async()
    {
        if (!plan_) throw std::runtime_error("hpxfft::shared::loop: initialize() has not been called");
        hpxfft::util::b200_check(hpxfft_b200_execute_async(plan_));
        hpxfft::util::b200_check(hpxfft_b200_download_async(plan_, values_vec_.data()));
        return hpxfft::util::when_stream_reaches<vector_2d>(plan_, [this]() { return std::move(values_vec_); });
    }

    real get_measurement(std::string name)
    {
        if (!plan_) return 0.0;
        hpxfft_b200_synchronize(plan_);  // refreshes the timers of transforms that were only enqueued
        return hpxfft_b200_measurement(plan_, name.c_str());
    }

    void write_plans_to_file(std::string file_path)
    {
        if (!plan_ || hpxfft_b200_write_plans(plan_, file_path.c_str()) != HPXFFT_B200_OK)
            throw std::runtime_error("Failed to open file: " + file_path);  // core/src/shared/loop.cpp:198-201
    }

    // extension: choose the CUDA device before initialize() (-1 = current device)
    void set_device(int device) { device_ = device; }

    ~loop() { reset(); }

  private:
    vector_2d run()
    {
        if (!plan_) throw std::runtime_error("hpxfft::shared::loop: initialize() has not been called");
        hpxfft::util::b200_check(hpxfft_b200_execute(plan_));
        hpxfft::util::b200_check(hpxfft_b200_download(plan_, values_vec_.data()));
        return std::move(values_vec_);  // single-use like the reference (core/src/shared/loop.cpp:112)
    }
    void reset()
    {
        if (plan_) hpxfft_b200_destroy(plan_);
        plan_ = nullptr;
    }

    hpxfft_b200_plan *plan_ = nullptr;
    int device_ = -1;
    vector_2d values_vec_;
};
}  // namespace hpxfft::shared
#endif
