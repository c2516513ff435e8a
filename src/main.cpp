// caps_sa <input_path> <output_path> [subproblem-count] [bounded-context]
//
// Same command line, byte mapping, index-width rule and output file as the reference driver
// (reference src/main.cpp:43-93); the mapping loop (:61-70) runs as a CUDA kernel.
#include "Suffix_Array.hpp"
#include "caps_sa_gpu.h"

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <limits>
#include <string>
#include <vector>


namespace
{

struct Pinned_Text
{
    char* data = nullptr;
    std::size_t size = 0;

    ~Pinned_Text() { caps_sa_gpu_host_free(data); }
};


// Whole-file read into pinned memory (so the host->device copy runs at PCIe speed).
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

    std::ifstream input(path, std::ios::binary);
    input.read(text.data, static_cast<std::streamsize>(file_size));
    if(static_cast<std::size_t>(input.gcount()) != file_size)
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
void build_and_dump(const Pinned_Text& text, const std::size_t subproblems, const std::size_t context, const bool pretty, std::ofstream& output)
{
    CaPS_SA::Suffix_Array<idx_t> suf_arr(text.data, static_cast<idx_t>(text.size), static_cast<idx_t>(subproblems), static_cast<idx_t>(context));
    suf_arr.construct();
    if(pretty)
        pretty_print(suf_arr, output);
    else
        suf_arr.dump(output);
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

    caps_sa_gpu_engine* const engine = caps_sa_gpu_engine_create(std::getenv("CAPS_SA_DEVICE") ? std::atoi(std::getenv("CAPS_SA_DEVICE")) : 0);
    if(!engine)
    {
        std::cerr << "Cannot initialise the CUDA engine: " << caps_sa_gpu_last_error() << "\n";
        return EXIT_FAILURE;
    }

    Pinned_Text text;
    read_input(ip_path, text);

    // Every byte — FASTA headers and newlines included — becomes one of A, C, T, G.
    if(caps_sa_gpu_map_acgt(engine, text.data, text.size) != CAPS_SA_GPU_OK)
    {
        std::cerr << "Byte mapping failed: " << caps_sa_gpu_last_error() << "\n";
        return EXIT_FAILURE;
    }
    caps_sa_gpu_engine_destroy(engine);

    std::ofstream output(op_path, std::ios::binary);

    const std::size_t n = text.size;
    std::cerr << "Text length: " << n << ".\n";
    if(n <= std::numeric_limits<uint32_t>::max())
        build_and_dump<uint32_t>(text, subproblem_count, max_context, pretty, output);
    else
        build_and_dump<uint64_t>(text, subproblem_count, max_context, pretty, output);

    output.close();

    return 0;
}
