// hpxfft_shared_loop on B200: same options, console report and runtimes CSV as
// examples/hpxfft/shared_loop_2d.cpp of HPX-FFT (--nx --ny --plan --run --header --result; sizes are
// literal, defaults 8 x 14 as in the reference; non-power-of-two lengths take the direct-DFT kernels).
#include <chrono>
#include <fstream>

#include "cli_common.hpp"
#include "hpxfft/shared/loop.hpp"

int main(int argc, char *argv[])
{
    cli::options opt(argc, argv, {{"result", "0"}, {"nx", "8"}, {"ny", "14"}, {"plan", "estimate"}, {"run", "par"}, {"header", "0"}});
    const std::string run_flag = opt.str("run"), plan_flag = opt.str("plan");
    const std::size_t dim_c_x = opt.num("nx"), dim_r_y = opt.num("ny"), dim_c_y = dim_r_y / 2 + 1;

    // ramp input v(i, j) = j  (examples/hpxfft/shared_loop_2d.cpp:33-40)
    hpxfft::shared::vector_2d values_vec(dim_c_x, 2 * dim_c_y);
    for (std::size_t i = 0; i < dim_c_x; ++i)
        for (std::size_t j = 0; j < dim_r_y; ++j) values_vec(i, j) = static_cast<double>(j);

    hpxfft::shared::loop fft_computer;
    const double start_total = cli::now_s();
    fft_computer.initialize(std::move(values_vec), plan_flag);
    const double stop_init = cli::now_s();
    values_vec = run_flag == "seq" ? fft_computer.fft_2d_r2c_seq() : fft_computer.fft_2d_r2c_par();
    const double stop_total = cli::now_s();
    if (opt.flag("result")) cli::print_vector_2d(values_vec);

    const double total = stop_total - start_total, init = stop_init - start_total;
    auto m = [&](const char *k) { return fft_computer.get_measurement(k); };
    std::cout << "\nLocality 0 - shared - " << run_flag << "\nTotal runtime : " << total << "\nInitialization: " << init
              << "\nFFT 2D runtime: " << m("total") << "\nFFTW r2c      : " << m("first_fftw") << "\nFirst trans   : "
              << m("first_trans") << "\nFFTW c2c      : " << m("second_fftw") << "\nSecond trans  : " << m("second_trans")
              << "\nPlan time     : " << m("plan") << "\nPlan flops    : " << m("plan_flops") << "\n";

    const std::string header = "n_threads;n_x;n_y;plan;run_flag;total;initialization;fft_2d_total;first_fftw;first_trans;"
                               "second_fftw;second_trans;plan_time;plan_flops;\n";
    auto line = [&](std::ostream &os) {
        os << 1 << ";" << dim_c_x << ";" << dim_r_y << ";" << plan_flag << ";" << run_flag << ";" << total << ";" << init << ";"
           << m("total") << ";" << m("first_fftw") << ";" << m("first_trans") << ";" << m("second_fftw") << ";"
           << m("second_trans") << ";" << m("plan") << ";" << m("plan_flops") << ";\n";
    };
    const std::string runtime_file_path = "runtimes/runtimes_hpx_shared_loop.txt";
    cli::create_parent_dir(runtime_file_path);
    std::ofstream runtime_file(runtime_file_path, std::ios_base::app);
    if (opt.flag("header")) runtime_file << header;
    line(runtime_file);
    runtime_file.close();

    const std::string plan_file_path = "plans/plan_hpx_shared_loop.txt";
    cli::create_parent_dir(plan_file_path);
    std::ofstream plan_info_file(plan_file_path, std::ios_base::app);
    plan_info_file << header;
    line(plan_info_file);
    plan_info_file.close();
    fft_computer.write_plans_to_file(plan_file_path);
    return 0;
}
