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
#include <iterator>
#include <random>
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
        // File rendezvous.  Every file name carries a SESSION nonce that rank 0 draws afresh for this run and
        // that every rank has acknowledged (open_session), so files left behind by an earlier run in the same
        // directory, or by an earlier loop object of this process, can never be mistaken for this exchange.
        const session_t &ses = open_session();
        const std::string base = ses.prefix + "_" + tag + "_" + std::to_string(next_generation()) + "_";
        publish(base + std::to_string(this_locality), mine);
        for (std::size_t i = 0; i < num_localities; ++i) all[i] = await(base + std::to_string(i), [](const std::string &) { return true; });
        // closing barrier: once everybody has read everything, every rank removes its own files
        publish(base + "done_" + std::to_string(this_locality), "1");
        for (std::size_t i = 0; i < num_localities; ++i) await(base + "done_" + std::to_string(i), [](const std::string &) { return true; });
        std::remove((base + std::to_string(this_locality)).c_str());
        if (!ses.last_done.empty()) std::remove(ses.last_done.c_str());   // everyone is past the previous barrier by now
        session().last_done = base + "done_" + std::to_string(this_locality);
#endif
        return all;
    }

  private:
#if !defined(HPXFFT_B200_WITH_HPX)
    struct session_t
    {
        bool open = false;
        std::string prefix;     // <dir>/hpxfft_b200_<job>_<nonce of this run>
        std::string last_done;  // this rank's barrier file of the previous exchange
        unsigned generation = 0;
    };
    static session_t &session()
    {
        static session_t s;  // one session per process: every loop object of the run shares it
        return s;
    }
    static unsigned next_generation() { return session().generation++; }

    static std::string random_token()
    {
        std::random_device rd;
        const unsigned long long v = (static_cast<unsigned long long>(rd()) << 32) ^ rd() ^
                                     static_cast<unsigned long long>(std::chrono::steady_clock::now().time_since_epoch().count());
        char buf[32];
        std::snprintf(buf, sizeof(buf), "%016llx", v);
        return buf;
    }
    static void publish(const std::string &path, const std::string &data)
    {
        const std::string tmp = path + ".tmp";
        {
            std::ofstream f(tmp, std::ios::binary | std::ios::trunc);
            f.write(data.data(), static_cast<std::streamsize>(data.size()));
        }
        std::rename(tmp.c_str(), path.c_str());  // atomic: readers see nothing or everything
    }
    static bool slurp(const std::string &path, std::string &out)
    {
        std::ifstream f(path, std::ios::binary);
        if (!f) return false;
        out.assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
        return true;
    }
    template <class Pred> static std::string await(const std::string &path, Pred ok)
    {
        std::string data;
        for (int tries = 0;; ++tries)
        {
            if (slurp(path, data) && ok(data)) return data;
            if (tries > 120000) throw std::runtime_error("hpxfft_b200 bootstrap: timeout waiting for " + path);
            std::this_thread::sleep_for(std::chrono::milliseconds(1));
        }
    }

    // Handshake that defeats stale files: every rank publishes a fresh random token; rank 0 draws a nonce and
    // publishes "nonce token_0 ... token_{n-1}" with the tokens it currently sees, re-publishing (with a new nonce)
    // whenever one changes; a rank accepts the first session line that carries ITS OWN token and acknowledges that
    // nonce; rank 0 is done when every acknowledgement names its current nonce.
    const session_t &open_session()
    {
        session_t &ses = session();
        if (ses.open) return ses;
        const std::string dir = std::getenv("HPXFFT_B200_RENDEZVOUS") ? std::getenv("HPXFFT_B200_RENDEZVOUS") : "/tmp";
        const std::string job = std::getenv("MASTER_PORT") ? std::getenv("MASTER_PORT") : "0";
        const std::string stem = dir + "/hpxfft_b200_" + job;
        const std::string me = std::to_string(this_locality), token = random_token();
        publish(stem + "_hello_" + me, token);
        std::string nonce;
        if (this_locality == 0)
        {
            std::vector<std::string> seen(num_localities);
            for (int tries = 0;; ++tries)
            {
                bool complete = true, changed = false;
                for (std::size_t i = 0; i < num_localities; ++i)
                {
                    std::string t;
                    if (!slurp(stem + "_hello_" + std::to_string(i), t) || t.size() != token.size()) { complete = false; continue; }
                    if (t != seen[i]) { seen[i] = t; changed = true; }
                }
                if (complete && (changed || nonce.empty()))
                {
                    nonce = random_token();
                    std::string line = nonce;
                    for (const auto &t : seen) line += " " + t;
                    publish(stem + "_session", line);
                }
                if (complete && !nonce.empty())
                {
                    bool acked = true;
                    for (std::size_t i = 1; i < num_localities && acked; ++i)
                    {
                        std::string a;
                        acked = slurp(stem + "_ack_" + std::to_string(i), a) && a == nonce;
                    }
                    if (acked) break;
                }
                if (tries > 120000) throw std::runtime_error("hpxfft_b200 bootstrap: timeout opening the rendezvous session");
                std::this_thread::sleep_for(std::chrono::milliseconds(1));
            }
        }
        else
        {
            // keep acknowledging the newest session line that names my token until rank 0 settles (it removes the
            // session file's hello inputs only after every acknowledgement matched, so the loop below terminates)
            for (int tries = 0;; ++tries)
            {
                std::string line;
                if (slurp(stem + "_session", line))
                {
                    std::vector<std::string> f;
                    std::size_t p = 0;
                    while (p <= line.size())
                    {
                        const std::size_t q = line.find(' ', p);
                        f.push_back(line.substr(p, q == std::string::npos ? std::string::npos : q - p));
                        if (q == std::string::npos) break;
                        p = q + 1;
                    }
                    if (f.size() == num_localities + 1 && f[1 + this_locality] == token)
                    {
                        if (f[0] != nonce)
                        {
                            nonce = f[0];
                            publish(stem + "_ack_" + me, nonce);
                        }
                        // settled when rank 0 publishes the go file for this nonce
                        std::string go;
                        if (slurp(stem + "_go", go) && go == nonce) break;
                    }
                }
                if (tries > 120000) throw std::runtime_error("hpxfft_b200 bootstrap: timeout joining the rendezvous session");
                std::this_thread::sleep_for(std::chrono::milliseconds(1));
            }
        }
        if (this_locality == 0) publish(stem + "_go", nonce);
        ses.prefix = stem + "_" + nonce;
        ses.open = true;
        return ses;
    }
#endif

    static std::size_t env_size(std::initializer_list<const char *> keys, std::size_t dflt)
    {
        for (const char *k : keys)
            if (const char *v = std::getenv(k)) return static_cast<std::size_t>(std::strtoull(v, nullptr, 10));
        return dflt;
    }
};
}  // namespace hpxfft::distributed
#endif
