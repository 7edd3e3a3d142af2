// Reads a bincode-serialised Lapper<I, u32> from the file argv[2] (argv[1]: "u32" or "u64" = I), loads it through the
// C++ facade (index rebuilt on the GPU), and writes `to_bincode()` of the loaded lapper to argv[3] followed by
// count(28974798, 33141355) on stdout.  tests/test_cpp_facade.py compares the bytes with the Python codec's.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iterator>
#include <vector>

#include "lapper_b200.hpp"

template <typename I>
static int run(const char *in, const char *out) {
    std::ifstream f(in, std::ios::binary);
    const std::vector<uint8_t> buf((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    try {
        auto lap = lapper_b200::Lapper<uint32_t, I>::from_bincode(buf);
        const std::vector<uint8_t> back = lap.to_bincode();
        std::ofstream o(out, std::ios::binary);
        o.write(reinterpret_cast<const char *>(back.data()), (std::streamsize)back.size());
        std::printf("%llu\n", (unsigned long long)lap.count(28974798, 33141355));
    } catch (const lapper_b200::Error &e) {
        std::printf("error %d: %s\n", e.code, e.what());
        return 2;
    }
    return 0;
}

int main(int argc, char **argv) {
    if (argc != 4) return 64;
    return std::strcmp(argv[1], "u64") == 0 ? run<uint64_t>(argv[2], argv[3]) : run<uint32_t>(argv[2], argv[3]);
}
