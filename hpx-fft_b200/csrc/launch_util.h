// launch_util.h -- opt-in to > 48 KB dynamic shared memory once per (kernel, device), thread-safe.
#pragma once
#include "internal.h"

#include <mutex>
#include <set>
#include <utility>

namespace hpxfft_b200 {

template <class K> int ensure_smem(K kernel, size_t bytes, int device)
{
    static std::mutex mu;
    static std::set<std::pair<const void *, int>> done;
    std::lock_guard<std::mutex> lk(mu);
    const std::pair<const void *, int> key((const void *) kernel, device);
    if (done.count(key)) return 0;
    if (int rc = set_smem(kernel, bytes)) return rc;
    done.insert(key);
    return 0;
}

}  // namespace hpxfft_b200
