// `aim_genpairs <seed> <first_pair> <n> <length> <error> <read_size> <out-file>` - write synthetic pairs in the
// Datasets file format (">pattern" / "<text" lines) with the generator of include/aim_b200.h (aim_generate_pairs:
// WFA `generate_dataset` semantics, Datasets/README.md:19-25).  A stand-alone host-only program (no CUDA, no
// libaim_b200.so): bench.py's reference arm uses it so that the timed reference process tree never loads product code.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "aim_b200.h"

int main(int argc, char *argv[])
{
    if (argc != 8) {
        fprintf(stderr, "usage: %s <seed> <first_pair> <n> <length> <error> <read_size> <out-file>\n", argv[0]);
        return 2;
    }
    const uint64_t seed = strtoull(argv[1], nullptr, 10), first = strtoull(argv[2], nullptr, 10);
    const uint32_t n = (uint32_t)strtoul(argv[3], nullptr, 10);
    const int32_t length = atoi(argv[4]), rs = atoi(argv[6]);
    const double error = atof(argv[5]);
    std::vector<int32_t> plen(n), tlen(n);
    std::vector<char> pats((size_t)n * rs), txts((size_t)n * rs);
    int rc = aim_generate_pairs(seed, first, n, length, error, rs, plen.data(), tlen.data(), pats.data(), txts.data(), (int32_t)std::max(1u, std::thread::hardware_concurrency()));
    if (rc == AIM_OK) rc = aim_write_pairs(argv[7], n, rs, plen.data(), tlen.data(), pats.data(), txts.data());
    if (rc != AIM_OK) { fprintf(stderr, "aim_genpairs: %s\n", aim_last_error()); return 1; }
    return 0;
}
