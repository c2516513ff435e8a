// caps_sa <input_path> <output_path> [subproblem-count] [bounded-context]
//
// Same command line, byte mapping, index-width rule and output file as the reference driver
// (reference src/main.cpp:43-93); the mapping loop (:61-70) runs as a CUDA kernel on the staged text,
// the file is read and the dump written by several threads (pread / pwrite).
#include "Suffix_Array.hpp"
#include "caps_sa_gpu.h"

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <algorithm>
#include <atomic>
#include <limits>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <unistd.h>


namespace
{

struct Pinned_Text
{
    char* data = nullptr;
    std::size_t size = 0;

    ~Pinned_Text() { caps_sa_gpu_host_free(data); }
};


// Whole-file read into pinned memory (so the host->device copy runs at PCIe speed), by several
// threads with pread: one stream through an ifstream is the slowest part of reading a multi-GB file
// that is in the page cache.
void read_input(const std::string& path, Pinned_Text& text)
{
    std::error_code ec;
    const auto file_size = std::filesystem::file_size(path, ec);
    if(ec)
    {
        std::cerr << path << " : " << ec.message() << "\n";
        std::exit(EXIT_FAILURE);
    }

    text.size = file_size;
    text.data = static_cast<char*>(caps_sa_gpu_host_alloc(file_size ? file_size : 1));
    if(!text.data)
    {
        std::cerr << "Cannot allocate " << file_size << " bytes of pinned memory: " << caps_sa_gpu_last_error() << "\n";
        std::exit(EXIT_FAILURE);
    }

    const int fd = ::open(path.c_str(), O_RDONLY);
    if(fd < 0)
    {
        std::cerr << path << " : cannot open\n";
        std::exit(EXIT_FAILURE);
    }
    std::atomic<bool> ok(true);
    const std::size_t workers = std::min<std::size_t>(std::max(1u, std::thread::hardware_concurrency()), std::max<std::size_t>(1, file_size >> 24));
    std::vector<std::thread> pool;
    for(std::size_t w = 0; w < workers; ++w)
        pool.emplace_back([&, w] {
            std::size_t at = file_size / workers * w;
            const std::size_t stop = w + 1 == workers ? file_size : file_size / workers * (w + 1);
            while(at < stop)
            {
                const ssize_t got = ::pread(fd, text.data + at, std::min<std::size_t>(stop - at, std::size_t(1) << 26), static_cast<off_t>(at));
                if(got <= 0) { ok = false; return; }
                at += static_cast<std::size_t>(got);
            }
        });
    for(std::thread& t : pool)
        t.join();
    ::close(fd);
    if(!ok)
    {
        std::cerr << path << " : short read\n";
        std::exit(EXIT_FAILURE);
    }
}


// Text form of the result: the suffix array on one line, the LCP array on the next, entries
// separated by blanks (the reference declares this as `pretty_print`, src/main.cpp:31-40, and
// lists `--pretty-print` in its usage line, :49, but never parses the flag; here it is live).
template <typename idx_t>
void pretty_print(const CaPS_SA::Suffix_Array<idx_t>& suf_arr, std::ofstream& output)
{
    const std::size_t n = suf_arr.n();
    for(std::size_t i = 0; i < n; ++i)
        output << suf_arr.SA()[i] << " \n"[i == n - 1];
    for(std::size_t i = 0; i < n; ++i)
        output << suf_arr.LCP()[i] << " \n"[i == n - 1];
}


template <typename idx_t>
void build_and_dump(const Pinned_Text& text, const std::size_t subproblems, const std::size_t context, const bool pretty, const std::string& op_path)
{
    CaPS_SA::Suffix_Array<idx_t> suf_arr(text.data, static_cast<idx_t>(text.size), static_cast<idx_t>(subproblems), static_cast<idx_t>(context));
    suf_arr.construct();
    if(pretty)
    {
        std::ofstream output(op_path, std::ios::binary);
        pretty_print(suf_arr, output);
        output.close();
    }
    else if(!suf_arr.dump(op_path.c_str()))
    {
        std::cerr << op_path << " : cannot write the suffix array\n";
        std::exit(EXIT_FAILURE);
    }
}

}


int main(int argc, char* argv[])
{
    if(argc < 3)
    {
        std::cerr << "Usage: CaPS_SA <input_path> <output_path> <(optional)-subproblem-count> <(optional)-bounded-context> <(optional)--pretty-print>\n";
        return EXIT_FAILURE;
    }

    // `--pretty-print` may stand anywhere after the two paths; the other arguments keep the
    // reference's positions.
    bool pretty = false;
    std::vector<const char*> args;
    for(int i = 0; i < argc; ++i)
        if(i >= 3 && std::string(argv[i]) == "--pretty-print")
            pretty = true;
        else
            args.push_back(argv[i]);

    const std::string ip_path(args[1]);
    const std::string op_path(args[2]);
    const std::size_t subproblem_count(args.size() >= 4 ? std::atoi(args[3]) : 0);
    const std::size_t max_context(args.size() >= 5 ? std::atoi(args[4]) : 0);

    Pinned_Text text;
    read_input(ip_path, text);

    // Every byte — FASTA headers and newlines included — becomes one of A, C, T, G (reference
    // src/main.cpp:61-70).  The mapping runs on the device copy of the text, as part of its staging:
    // the text crosses PCIe once.
    caps_sa_gpu_set_cli_byte_mapping(1);

    const std::size_t n = text.size;
    std::cerr << "Text length: " << n << ".\n";
    if(n <= std::numeric_limits<uint32_t>::max())
        build_and_dump<uint32_t>(text, subproblem_count, max_context, pretty, op_path);
    else
        build_and_dump<uint64_t>(text, subproblem_count, max_context, pretty, op_path);

    return 0;
}
