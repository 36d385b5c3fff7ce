// hpxfft::distributed::loop -- drop-in for core/include/hpxfft/distributed/loop.hpp +
// core/src/distributed/loop.cpp.  Header-only wrapper over libhpxfft_b200.so.
//
//   hpxfft::distributed::loop fft;                         // examples/hpxfft/distributed_loop_2d.cpp:40-45
//   fft.initialize(std::move(slab), "all_to_all", "estimate");   // or "scatter"; "p2p" is an extension
//   slab = fft.fft_2d_r2c();                               // collective across localities
//
// Result layout: natural order (what fftw_mpi_plan_dft_r2c_2d returns,
// examples/fftw/fftw_mpi_omp_2d.cpp:119-120); with one locality identical to shared::loop.
#ifndef HPXFFT_B200_DISTRIBUTED_LOOP_HPP
#define HPXFFT_B200_DISTRIBUTED_LOOP_HPP

#include "../util/b200_error.hpp"
#include "../util/stream_future.hpp"
#include "../util/vector_2d.hpp"
#include "bootstrap.hpp"

#include <iostream>
#include <string>
#include <utility>

typedef double real;

namespace hpxfft::distributed
{
using vector_2d = hpxfft::util::vector_2d<real>;

struct loop
{
  public:
    loop() = default;
    loop(const loop &) = delete;
    loop &operator=(const loop &) = delete;

    void initialize(vector_2d values_vec, const std::string COMM_FLAG, const std::string PLAN_FLAG)
    {
        hpxfft::util::check_plan_flag(PLAN_FLAG);
        reset();
        values_vec_ = std::move(values_vec);
        COMM_FLAG_ = COMM_FLAG;
        if (COMM_FLAG != "scatter" && COMM_FLAG != "all_to_all" && COMM_FLAG != "p2p")
        {
            // core/src/distributed/loop.cpp:342-346: message, no exception (the reference also calls hpx::finalize)
            std::cout << "Specify communication scheme: scatter or all_to_all\n";
            return;
        }
        const int rank = static_cast<int>(boot_.this_locality), nranks = static_cast<int>(boot_.num_localities);
        std::string uid(HPXFFT_B200_UNIQUE_ID_BYTES, '\0');
        if (nranks > 1)
        {
            if (rank == 0) hpxfft::util::b200_check(hpxfft_b200_get_unique_id(&uid[0]));
            uid = boot_.all_gather("uid", uid)[0];
        }
        hpxfft::util::b200_check(hpxfft_b200_create(&plan_, values_vec_.n_row(), values_vec_.n_col(), rank, nranks,
                                                   device_ >= 0 ? device_ : boot_.local_device(), COMM_FLAG.c_str(),
                                                   PLAN_FLAG.c_str(), nranks > 1 ? uid.data() : nullptr));
        if (const int cnt = hpxfft_b200_ipc_count(plan_); cnt > 0)  // peer windows (copy-engine / fused transports)
        {
            std::string mine(static_cast<std::size_t>(cnt) * HPXFFT_B200_IPC_HANDLE_BYTES, '\0');
            hpxfft::util::b200_check(hpxfft_b200_ipc_export(plan_, &mine[0]));
            std::string all;
            for (const auto &s : boot_.all_gather("ipc", mine)) all += s;
            hpxfft::util::b200_check(hpxfft_b200_ipc_import(plan_, all.data()));
        }
        hpxfft::util::b200_check(hpxfft_b200_upload(plan_, values_vec_.data()));
    }

    vector_2d fft_2d_r2c()
    {
        if (!plan_)
        {
            std::cout << "Communication scheme not specified during initialization\n";  // distributed/loop.cpp:175-179
            return std::move(values_vec_);
        }
        hpxfft::util::b200_check(hpxfft_b200_execute(plan_));
        hpxfft::util::b200_check(hpxfft_b200_download(plan_, values_vec_.data()));
        return std::move(values_vec_);
    }

    // extension (agas client surface): enqueue only; the future is fulfilled by a stream callback.  Collective like fft_2d_r2c.
    hpxfft::util::future<vector_2d> fft_2d_r2c_async()
    {
        if (!plan_) throw std::runtime_error("hpxfft::distributed::loop: no plan (initialize() missing or bad COMM_FLAG)");
        hpxfft::util::b200_check(hpxfft_b200_execute_async(plan_));
        hpxfft::util::b200_check(hpxfft_b200_download_async(plan_, values_vec_.data()));
        return hpxfft::util::when_stream_reaches<vector_2d>(plan_, [this]() { return std::move(values_vec_); });
    }

    bool has_plan() const { return plan_ != nullptr; }

    real get_measurement(std::string name)
    {
        if (!plan_) return 0.0;
        hpxfft_b200_synchronize(plan_);  // refreshes the timers of transforms that were only enqueued
        return hpxfft_b200_measurement(plan_, name.c_str());
    }

    void set_device(int device) { device_ = device; }
    std::size_t this_locality() const { return boot_.this_locality; }
    std::size_t num_localities() const { return boot_.num_localities; }

    ~loop() { reset(); }

  private:
    void reset()
    {
        if (plan_) hpxfft_b200_destroy(plan_);
        plan_ = nullptr;
    }

    hpxfft_b200_plan *plan_ = nullptr;
    int device_ = -1;
    bootstrap boot_;
    std::string COMM_FLAG_;
    vector_2d values_vec_;
};
}  // namespace hpxfft::distributed
#endif
