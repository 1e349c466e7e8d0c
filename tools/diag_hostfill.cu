// How fast can host threads rebuild the reference's op rows ('M'-filled 2*READ_SIZE rows) in pinned memory, alone and beside a
// saturating H2D copy?  (decides whether aim_align_batch may download run-length rows and expand them on the host)
//   diag_hostfill [pairs=10000000] [row=336]
#include <cuda_runtime.h>
#include <immintrin.h>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
__attribute__((target("avx2"))) static void fill_nt(char *p, size_t n)
{
    const __m256i v = _mm256_set1_epi8('M');
    for (size_t i = 0; i < n; i += 32) _mm256_stream_si256((__m256i *)(p + i), v);
    _mm_sfence();
}
int main(int argc, char **argv)
{
    const size_t pairs = argc > 1 ? (size_t)atoll(argv[1]) : 10000000, row = argc > 2 ? (size_t)atoll(argv[2]) : 336;
    const size_t bytes = pairs * row;
    char *h_ops, *h_in, *d_in;
    CK(cudaHostAlloc(&h_ops, bytes, cudaHostAllocPortable));
    CK(cudaHostAlloc(&h_in, bytes, cudaHostAllocPortable));
    CK(cudaMalloc(&d_in, bytes));
    memset(h_ops, 1, bytes); memset(h_in, 2, bytes);
    cudaStream_t st; CK(cudaStreamCreate(&st));
    const unsigned hw = std::thread::hardware_concurrency();
    printf("{\"host_threads\": %u, \"pairs\": %zu, \"row\": %zu, \"cases\": [\n", hw, pairs, row);
    bool first = true;
    for (int mode = 0; mode < 3; ++mode)          // 0 memset whole row, 1 non-temporal whole row, 2 memset the span only (160 of 336 bytes)
        for (int beside = 0; beside < 2; ++beside)  // a 55 GB/s H2D copy running at the same time
            for (unsigned T : {4u, 8u, 16u, 32u}) {
                if (T > hw) continue;
                double best = 1e30, best_copy = 0;
                for (int rep = 0; rep < 3; ++rep) {
                    std::atomic<size_t> next{0};
                    const size_t blk = 4096;  // pairs per grab
                    const double t0 = now();
                    if (beside) CK(cudaMemcpyAsync(d_in, h_in, bytes, cudaMemcpyHostToDevice, st));
                    std::vector<std::thread> th;
                    for (unsigned t = 0; t < T; ++t)
                        th.emplace_back([&]() {
                            for (;;) {
                                const size_t b = next.fetch_add(blk);
                                if (b >= pairs) break;
                                const size_t e = std::min(pairs, b + blk);
                                if (mode == 0) memset(h_ops + b * row, 'M', (e - b) * row);
                                else if (mode == 1) fill_nt(h_ops + b * row, (e - b) * row);
                                else for (size_t i = b; i < e; ++i) memset(h_ops + i * row + 170, 'M', 166);
                            }
                        });
                    for (auto &x : th) x.join();
                    const double t1 = now();
                    CK(cudaStreamSynchronize(st));
                    const double t2 = now();
                    if (t1 - t0 < best) { best = t1 - t0; best_copy = t2 - t0; }
                }
                printf("%s{\"mode\": \"%s\", \"beside_h2d\": %d, \"threads\": %u, \"fill_ms\": %.1f, \"pairs_per_s\": %.3g, \"fill_gbs\": %.1f, \"h2d_ms\": %.1f}", first ? "" : ",\n",
                       mode == 0 ? "memset_row" : mode == 1 ? "nt_row" : "memset_span", beside, T, best * 1e3, pairs / best,
                       (mode == 2 ? pairs * 166.0 : (double)bytes) / best * 1e-9, beside ? best_copy * 1e3 : 0.0);
                first = false;
                fflush(stdout);
            }
    printf("\n]}\n");
    return 0;
}
