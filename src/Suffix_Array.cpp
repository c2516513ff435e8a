// Class shell over the C-ABI (include/caps_sa_gpu.h).  Host-side only: argument plumbing,
// pinned result arrays, the reference's stderr lines and dump layout.
#include "Suffix_Array.hpp"
#include "caps_sa_gpu.h"

#include <chrono>
#include <cstdlib>
#include <algorithm>
#include <atomic>
#include <iostream>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <unistd.h>

namespace CaPS_SA
{

namespace
{

[[noreturn]] void die(const char* what)
{
    std::cerr << what << ": " << caps_sa_gpu_last_error() << "\nAborting.\n";
    std::exit(EXIT_FAILURE);
}

// One engine per process (device from CAPS_SA_DEVICE, default 0), created on first use.
caps_sa_gpu_engine* shared_engine()
{
    static caps_sa_gpu_engine* engine = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* env = std::getenv("CAPS_SA_DEVICE");
        engine = caps_sa_gpu_engine_create(env ? std::atoi(env) : 0);
        if(!engine)
            die("Cannot initialise the CUDA engine");
    });
    return engine;
}

template <typename idx_t>
idx_t* pinned_array(std::size_t count)
{
    void* const p = caps_sa_gpu_host_alloc((count ? count : 1) * sizeof(idx_t));
    if(!p)
        die("Cannot allocate pinned host memory for the suffix array");
    return static_cast<idx_t*>(p);
}

// Devices of the sharded construction: CAPS_SA_GPUS = a count ("4" -> devices 0..3) or an explicit
// list ("0,2,3").  Empty / "1" / unset selects the single-device path.
std::vector<int> sharded_devices()
{
    std::vector<int> devices;
    const char* env = std::getenv("CAPS_SA_GPUS");
    if(!env || !*env)
        return devices;
    const std::string spec(env);
    if(spec.find(',') == std::string::npos)
    {
        const int count = std::atoi(spec.c_str());
        for(int d = 0; count > 1 && d < count; ++d)
            devices.push_back(d);
        return devices;
    }
    std::size_t at = 0;
    while(at <= spec.size())
    {
        const std::size_t comma = std::min(spec.find(',', at), spec.size());
        if(comma > at)
            devices.push_back(std::atoi(spec.substr(at, comma - at).c_str()));
        at = comma + 1;
    }
    return devices;
}

inline double seconds_since(const std::chrono::steady_clock::time_point t0)
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// The reference prints one "... Time taken: X seconds." line per stage of its samplesort
// (src/Suffix_Array.cpp:144,157,183,221,248,367,408,446,462); the stages here are different ones,
// reported in the same form from the engine's CUDA-event times.
void report_stages(const caps_sa_gpu_stats& s)
{
    const auto line = [](const char* what, const float ms) { std::cerr << what << " Time taken: " << ms / 1e3 << " seconds.\n"; };
    if(s.ms_h2d > 0)
        line("Staged the text on the device.", s.ms_h2d);
    line("Packed the text into fixed-width codes.", s.ms_pack);
    if(s.ms_partition > 0)
        line("Partitioned the suffixes among the devices.", s.ms_partition);
    line("Sorted the suffixes by their prefix keys.", s.ms_sort);
    line("Resolved the tied suffixes.", s.ms_refine);
    line("Computed the remaining LCPs.", s.ms_deep_lcp);
    if(s.ms_d2h > 0)
        line("Copied the suffix array and the LCP array to the host (overlapped).", s.ms_d2h);
}

}


template <typename T_idx_>
Suffix_Array<T_idx_>::Suffix_Array(const char* const T, const idx_t n, const idx_t subproblem_count, const idx_t max_context):
    T_(T),
    n_(n),
    SA_((shared_engine(), pinned_array<idx_t>(n))),
    LCP_(pinned_array<idx_t>(n)),
    subproblem_hint_(subproblem_count),
    max_context_(max_context),
    constructed_(false)
{
    // The reference derives p = min(subproblem_count ? subproblem_count : 8192, n / 16) (src/Suffix_Array.cpp:24)
    // before it tests p > n (:33-37), so that test never fires for any argument (and n < 16 dies
    // earlier, by a division by zero, :27): every subproblem count is accepted.  Here the count is a
    // hint the construction ignores — the output does not depend on it.
}


template <typename T_idx_>
Suffix_Array<T_idx_>::~Suffix_Array()
{
    caps_sa_gpu_host_free(SA_);
    caps_sa_gpu_host_free(LCP_);
}


template <>
void Suffix_Array<uint32_t>::construct()
{
    const auto t0 = std::chrono::steady_clock::now();
    const std::vector<int> devices = sharded_devices();
    std::vector<caps_sa_gpu_stats> rank_stats(devices.size() > 1 ? devices.size() : 1);
    const int rc = devices.size() > 1
        ? caps_sa_gpu_construct_multi_u32(devices.data(), static_cast<int>(devices.size()), T_, n_, SA_, LCP_, subproblem_hint_, max_context_, rank_stats.data())
        : caps_sa_gpu_construct_u32(shared_engine(), T_, n_, SA_, LCP_, subproblem_hint_, max_context_);
    if(rc != CAPS_SA_GPU_OK)
        die("Suffix array construction failed");
    if(devices.size() <= 1)
        caps_sa_gpu_engine_stats(shared_engine(), rank_stats.data());
    report_stages(rank_stats[0]);
    constructed_ = true;
    std::cerr << "Constructed the suffix array. Time taken: " << seconds_since(t0) << " seconds.\n";
}


template <>
void Suffix_Array<uint64_t>::construct()
{
    const auto t0 = std::chrono::steady_clock::now();
    const std::vector<int> devices = sharded_devices();
    std::vector<caps_sa_gpu_stats> rank_stats(devices.size() > 1 ? devices.size() : 1);
    const int rc = devices.size() > 1
        ? caps_sa_gpu_construct_multi_u64(devices.data(), static_cast<int>(devices.size()), T_, n_, SA_, LCP_, subproblem_hint_, max_context_, rank_stats.data())
        : caps_sa_gpu_construct_u64(shared_engine(), T_, n_, SA_, LCP_, subproblem_hint_, max_context_);
    if(rc != CAPS_SA_GPU_OK)
        die("Suffix array construction failed");
    if(devices.size() <= 1)
        caps_sa_gpu_engine_stats(shared_engine(), rank_stats.data());
    report_stages(rank_stats[0]);
    constructed_ = true;
    std::cerr << "Constructed the suffix array. Time taken: " << seconds_since(t0) << " seconds.\n";
}


template <typename T_idx_>
void Suffix_Array<T_idx_>::dump(std::ofstream& output)
{
    const auto t0 = std::chrono::steady_clock::now();

    const std::size_t len = n_;
    output.write(reinterpret_cast<const char*>(&len), sizeof(len));
    output.write(reinterpret_cast<const char*>(SA_), static_cast<std::streamsize>(len * sizeof(idx_t)));
    output.write(reinterpret_cast<const char*>(LCP_), static_cast<std::streamsize>(len * sizeof(idx_t)));

    std::cerr << "Dumped the suffix array. Time taken: " << seconds_since(t0) << " seconds.\n";
}



template <typename T_idx_>
bool Suffix_Array<T_idx_>::dump(const char* const path)
{
    const auto t0 = std::chrono::steady_clock::now();

    const int fd = ::open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if(fd < 0)
        return false;

    const std::size_t len = n_;
    const std::size_t arr_bytes = len * sizeof(idx_t);
    std::atomic<bool> ok(true);
    // the file is header || SA || LCP; every thread writes a contiguous share of each array
    const auto write_all = [&](const char* src, std::size_t bytes, std::size_t at) {
        while(bytes > 0)
        {
            const std::size_t step = std::min<std::size_t>(bytes, std::size_t(1) << 26);
            const ssize_t done = ::pwrite(fd, src, step, static_cast<off_t>(at));
            if(done <= 0) { ok = false; return; }
            src += done, at += static_cast<std::size_t>(done), bytes -= static_cast<std::size_t>(done);
        }
    };
    write_all(reinterpret_cast<const char*>(&len), sizeof(len), 0);
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const std::size_t workers = std::min<std::size_t>(hw, std::max<std::size_t>(1, arr_bytes >> 24));
    std::vector<std::thread> pool;
    for(std::size_t w = 0; w < workers; ++w)
        pool.emplace_back([&, w] {
            const std::size_t lo = arr_bytes / workers * w, hi = w + 1 == workers ? arr_bytes : arr_bytes / workers * (w + 1);
            write_all(reinterpret_cast<const char*>(SA_) + lo, hi - lo, sizeof(len) + lo);
            write_all(reinterpret_cast<const char*>(LCP_) + lo, hi - lo, sizeof(len) + arr_bytes + lo);
        });
    for(std::thread& t : pool)
        t.join();
    if(::close(fd) != 0)
        ok = false;

    std::cerr << "Dumped the suffix array. Time taken: " << seconds_since(t0) << " seconds.\n";
    return ok;
}

}


template class CaPS_SA::Suffix_Array<uint32_t>;
template class CaPS_SA::Suffix_Array<uint64_t>;
