// hpxfft_distributed_loop on B200: same options, console report and runtimes CSV as
// examples/hpxfft/distributed_loop_2d.cpp of HPX-FFT (--run scatter|all_to_all, plus p2p).
// SPMD: start one process per GPU with RANK / WORLD_SIZE / LOCAL_RANK set (torchrun, srun, mpirun).
#include <chrono>
#include <fstream>
#include <thread>

#include "cli_common.hpp"
#include "hpxfft/distributed/loop.hpp"

int main(int argc, char *argv[])
{
    cli::options opt(argc, argv, {{"result", "0"}, {"nx", "8"}, {"ny", "14"}, {"plan", "estimate"}, {"run", "scatter"}, {"header", "0"}});
    const std::string run_flag = opt.str("run"), plan_flag = opt.str("plan");
    hpxfft::distributed::loop fft_computer;
    const std::size_t this_locality = fft_computer.this_locality(), num_localities = fft_computer.num_localities();
    const std::size_t dim_c_x = opt.num("nx"), dim_r_y = opt.num("ny"), dim_c_y = dim_r_y / 2 + 1;
    const std::size_t n_x_local = dim_c_x / num_localities;  // examples/hpxfft/distributed_loop_2d.cpp:25

    hpxfft::distributed::vector_2d values_vec(n_x_local, 2 * dim_c_y);
    for (std::size_t i = 0; i < n_x_local; ++i)
        for (std::size_t j = 0; j < dim_r_y; ++j) values_vec(i, j) = static_cast<double>(j);

    const double start_total = cli::now_s();
    fft_computer.initialize(std::move(values_vec), run_flag, plan_flag);
    const double stop_init = cli::now_s();
    values_vec = fft_computer.fft_2d_r2c();
    const double stop_total = cli::now_s();
    if (opt.flag("result"))
    {
        std::this_thread::sleep_for(std::chrono::seconds(this_locality));
        cli::print_vector_2d(values_vec);
    }
    if (this_locality != 0) return 0;

    const double total = stop_total - start_total, init = stop_init - start_total;
    auto m = [&](const char *k) { return fft_computer.get_measurement(k); };
    const char *keys[] = {"total", "first_fftw", "first_split", "first_comm", "first_trans", "second_fftw", "second_split",
                          "second_comm", "second_trans"};
    std::cout << "\nLocality 0 - " << run_flag << "\nTotal runtime : " << total << "\nInitialization: " << init << "\n";
    for (const char *k : keys) std::cout << k << ": " << m(k) << "\n";

    const std::string runtime_file_path = "runtimes/runtimes_hpx_distributed_loop.txt";
    cli::create_parent_dir(runtime_file_path);
    std::ofstream f(runtime_file_path, std::ios_base::app);
    if (opt.flag("header"))
        f << "n_threads;n_x;n_y;plan;run_flag;total;initialization;fft_2d_total;first_fftw;first_split;first_comm;first_trans;"
             "second_fftw;second_split;second_comm;second_trans;\n";
    f << num_localities << ";" << dim_c_x << ";" << dim_r_y << ";" << plan_flag << ";" << run_flag << ";" << total << ";" << init
      << ";";
    for (const char *k : keys) f << m(k) << ";";
    f << "\n";
    return 0;
}
