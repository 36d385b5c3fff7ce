// Shared helpers of the two example drivers: --key=value / --key value parsing with the reference's
// option names (examples/hpxfft/shared_loop_2d.cpp:139-146), parent-directory creation
// (core/src/util/create_dir.cpp:3-18) and the "(re im)" result printer (util/print_vector.hpp:10-38).
#pragma once
#include <cstdio>
#include <cstdlib>
#include <filesystem>
#include <iostream>
#include <map>
#include <stdexcept>
#include <string>

#include "hpxfft/util/vector_2d.hpp"

namespace cli
{
struct options
{
    std::map<std::string, std::string> kv;
    options(int argc, char **argv, std::map<std::string, std::string> defaults) : kv(std::move(defaults))
    {
        for (int i = 1; i < argc; ++i)
        {
            std::string a = argv[i];
            if (a.rfind("--", 0) != 0) continue;
            a = a.substr(2);
            auto eq = a.find('=');
            if (eq != std::string::npos)
                kv[a.substr(0, eq)] = a.substr(eq + 1);
            else if (i + 1 < argc && std::string(argv[i + 1]).rfind("--", 0) != 0)
                kv[a] = argv[++i];
            else
                kv[a] = "1";
        }
    }
    std::string str(const std::string &k) const { return kv.at(k); }
    std::size_t num(const std::string &k) const { return std::strtoull(kv.at(k).c_str(), nullptr, 10); }
    bool flag(const std::string &k) const { return kv.at(k) != "0" && kv.at(k) != "false"; }
};

inline void create_parent_dir(const std::string &file_path)
{
    const auto parent = std::filesystem::path(file_path).parent_path();
    if (parent.empty()) return;
    std::error_code ec;
    std::filesystem::create_directories(parent, ec);
    if (ec) throw std::runtime_error("Failed to create directory: " + parent.string());
}

template <typename T>
void print_vector_2d(const hpxfft::util::vector_2d<T> &v)
{
    for (std::size_t i = 0; i < v.n_row(); ++i)
    {
        for (std::size_t j = 0; j + 1 < v.n_col(); j += 2) std::cout << "(" << v(i, j) << " " << v(i, j + 1) << ") ";
        std::cout << "\n";
    }
    std::cout << "\n" << std::flush;
}

inline double now_s()
{
    using clk = std::chrono::steady_clock;
    return std::chrono::duration<double>(clk::now().time_since_epoch()).count();
}
}  // namespace cli
