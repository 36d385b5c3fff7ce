// Locality discovery and the few hundred bytes of plan bootstrap for hpxfft::distributed::loop.
//
// With HPX (-DHPXFFT_B200_WITH_HPX): hpx::get_locality_id / get_num_localities and HPX collectives,
// exactly the calls the reference makes (core/src/distributed/loop.cpp:281-282, 324-327).
// Without HPX: SPMD launch by any process launcher that exports RANK / WORLD_SIZE (torchrun, srun with
// SLURM_PROCID / SLURM_NTASKS, mpirun with OMPI_COMM_WORLD_*), and a shared-directory rendezvous
// (HPXFFT_B200_RENDEZVOUS, default /tmp) for the NCCL id and the IPC handles.
#ifndef HPXFFT_B200_BOOTSTRAP_HPP
#define HPXFFT_B200_BOOTSTRAP_HPP

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#if defined(HPXFFT_B200_WITH_HPX)
#include <hpx/hpx.hpp>
#include <hpx/modules/collectives.hpp>
#endif

namespace hpxfft::distributed
{
struct bootstrap
{
    std::size_t this_locality = 0, num_localities = 1;
    unsigned generation = 0;

    bootstrap()
    {
#if defined(HPXFFT_B200_WITH_HPX)
        this_locality = hpx::get_locality_id();
        num_localities = hpx::get_num_localities(hpx::launch::sync);
#else
        this_locality = env_size({"RANK", "SLURM_PROCID", "OMPI_COMM_WORLD_RANK", "PMI_RANK"}, 0);
        num_localities = env_size({"WORLD_SIZE", "SLURM_NTASKS", "OMPI_COMM_WORLD_SIZE", "PMI_SIZE"}, 1);
#endif
    }

    int local_device() const
    {
        for (const char *k : {"LOCAL_RANK", "SLURM_LOCALID", "OMPI_COMM_WORLD_LOCAL_RANK"})
            if (const char *v = std::getenv(k)) return std::atoi(v);
        return static_cast<int>(this_locality);
    }

    // every locality contributes `mine`; returns all contributions in locality order
    std::vector<std::string> all_gather(const std::string &tag, const std::string &mine)
    {
        std::vector<std::string> all(num_localities);
        if (num_localities == 1)
        {
            all[0] = mine;
            return all;
        }
#if defined(HPXFFT_B200_WITH_HPX)
        auto comm = hpx::collectives::create_communicator(
            ("hpxfft_b200_" + tag).c_str(), hpx::collectives::num_sites_arg(num_localities),
            hpx::collectives::this_site_arg(this_locality));
        std::vector<char> v(mine.begin(), mine.end());
        auto res = hpx::collectives::all_gather(comm, std::move(v)).get();
        for (std::size_t i = 0; i < num_localities; ++i) all[i].assign(res[i].begin(), res[i].end());
#else
        const std::string dir = std::getenv("HPXFFT_B200_RENDEZVOUS") ? std::getenv("HPXFFT_B200_RENDEZVOUS") : "/tmp";
        const std::string job = std::getenv("MASTER_PORT") ? std::getenv("MASTER_PORT") : "0";
        const std::string base = dir + "/hpxfft_b200_" + job + "_" + tag + "_" + std::to_string(generation) + "_";
        {
            const std::string tmp = base + std::to_string(this_locality) + ".tmp";
            std::ofstream f(tmp, std::ios::binary);
            f.write(mine.data(), static_cast<std::streamsize>(mine.size()));
            f.close();
            std::rename(tmp.c_str(), (base + std::to_string(this_locality)).c_str());
        }
        for (std::size_t i = 0; i < num_localities; ++i)
        {
            const std::string path = base + std::to_string(i);
            for (int tries = 0;; ++tries)
            {
                std::ifstream f(path, std::ios::binary);
                if (f)
                {
                    all[i].assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
                    if (all[i].size() == mine.size()) break;
                }
                if (tries > 60000) throw std::runtime_error("hpxfft_b200 bootstrap: timeout waiting for " + path);
                std::this_thread::sleep_for(std::chrono::milliseconds(1));
            }
        }
#endif
        ++generation;
        return all;
    }

  private:
    static std::size_t env_size(std::initializer_list<const char *> keys, std::size_t dflt)
    {
        for (const char *k : keys)
            if (const char *v = std::getenv(k)) return static_cast<std::size_t>(std::strtoull(v, nullptr, 10));
        return dflt;
    }
};
}  // namespace hpxfft::distributed
#endif
