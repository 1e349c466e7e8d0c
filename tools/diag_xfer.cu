// diag_xfer — copy-only ceiling of host<->device transfers on this box, for 1..N GPUs at once.
//
//   build/diag_xfer [--mb 1024] [--reps 3] [--gpus 8]
//
// For g = 1, 2, 4, ... GPUs driven concurrently (one host thread + two streams per GPU, the same structure as
// aim_align_batch's per-GPU pipeline): H2D alone, D2H alone, both directions together; H2D from write-combined pinned memory;
// and both directions while a third host thread per GPU streams through ordinary memory (what a caller's parser/printer does).
// Prints one JSON object: aggregate and per-GPU GB/s per case.  bench.py's e2e at N GPUs is judged against these numbers
// (DESIGN.md section 6): the reference layout moves 696 B per pair, so pairs/s <= bidirectional GB/s / 348 B per direction.
#include <cuda_runtime.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

struct Gpu {
    int dev;
    char *h_in = nullptr, *h_in_wc = nullptr, *h_out = nullptr, *d_in = nullptr, *d_out = nullptr;
    cudaStream_t s_in = nullptr, s_out = nullptr;
};

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char **argv)
{
    size_t mb = 1024;
    int reps = 3, want = 0;
    for (int i = 1; i + 1 < argc; i += 2) {
        if (!strcmp(argv[i], "--mb")) mb = (size_t)atol(argv[i + 1]);
        else if (!strcmp(argv[i], "--reps")) reps = atoi(argv[i + 1]);
        else if (!strcmp(argv[i], "--gpus")) want = atoi(argv[i + 1]);
    }
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (want > 0 && want < ndev) ndev = want;
    const size_t bytes = mb << 20;
    const size_t chunk = (size_t)96 << 20;  // the library's chunk size for short-read WFA
    std::vector<Gpu> G((size_t)ndev);
    for (int d = 0; d < ndev; ++d) {
        Gpu &g = G[(size_t)d];
        g.dev = d;
        CK(cudaSetDevice(d));
        CK(cudaHostAlloc(&g.h_in, bytes, cudaHostAllocPortable));
        CK(cudaHostAlloc(&g.h_in_wc, bytes, cudaHostAllocPortable | cudaHostAllocWriteCombined));
        CK(cudaHostAlloc(&g.h_out, bytes, cudaHostAllocPortable));
        memset(g.h_in, 1, bytes);
        memset(g.h_in_wc, 1, bytes);
        memset(g.h_out, 0, bytes);
        CK(cudaMalloc(&g.d_in, bytes));
        CK(cudaMalloc(&g.d_out, bytes));
        CK(cudaMemset(g.d_out, 2, bytes));
        CK(cudaStreamCreateWithFlags(&g.s_in, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&g.s_out, cudaStreamNonBlocking));
    }
    // case: bit 0 = H2D, bit 1 = D2H, bit 2 = H2D source is write-combined, bit 3 = host threads stream memory beside the copies
    struct Case { const char *name; int mask; };
    const Case cases[] = {{"h2d", 1}, {"d2h", 2}, {"bidir", 3}, {"h2d_wc", 5}, {"bidir_wc", 7}, {"bidir_host_busy", 11}};
    printf("{\"mb_per_gpu_per_direction\": %zu, \"chunk_mb\": %zu, \"visible_gpus\": %d, \"host_threads\": %u, \"cases\": [\n", mb, chunk >> 20, ndev,
           std::thread::hardware_concurrency());
    bool first = true;
    for (int g = 1; g <= ndev; g *= 2) {
        for (const Case &c : cases) {
            double best = 1e30;
            std::vector<double> per_gpu((size_t)g, 0.0);
            for (int r = 0; r < reps; ++r) {
                std::atomic<int> ready{0};
                std::atomic<bool> go{false}, stop{false};
                std::vector<double> t((size_t)g, 0.0);
                std::vector<std::thread> th, busy;
                for (int d = 0; d < g; ++d) {
                    th.emplace_back([&, d]() {
                        Gpu &u = G[(size_t)d];
                        CK(cudaSetDevice(u.dev));
                        ++ready;
                        while (!go.load()) std::this_thread::yield();
                        const double t0 = now();
                        for (size_t off = 0; off < bytes; off += chunk) {
                            const size_t m = std::min(chunk, bytes - off);
                            if (c.mask & 1) CK(cudaMemcpyAsync(u.d_in + off, ((c.mask & 4) ? u.h_in_wc : u.h_in) + off, m, cudaMemcpyHostToDevice, u.s_in));
                            if (c.mask & 2) CK(cudaMemcpyAsync(u.h_out + off, u.d_out + off, m, cudaMemcpyDeviceToHost, u.s_out));
                        }
                        CK(cudaStreamSynchronize(u.s_in));
                        CK(cudaStreamSynchronize(u.s_out));
                        t[(size_t)d] = now() - t0;
                    });
                    if (c.mask & 8)
                        busy.emplace_back([&, d]() {
                            std::vector<char> a((size_t)256 << 20, 1), b((size_t)256 << 20);
                            while (!stop.load()) memcpy(b.data(), a.data(), a.size());
                            (void)d;
                        });
                }
                while (ready.load() < g) std::this_thread::yield();
                const double t0 = now();
                go.store(true);
                for (auto &x : th) x.join();
                const double dt = now() - t0;
                stop.store(true);
                for (auto &x : busy) x.join();
                if (dt < best) {
                    best = dt;
                    for (int d = 0; d < g; ++d) per_gpu[(size_t)d] = (double)bytes / t[(size_t)d] / 1e9;
                }
            }
            const int dirs = ((c.mask & 1) ? 1 : 0) + ((c.mask & 2) ? 1 : 0);
            printf("%s {\"gpus\": %d, \"case\": \"%s\", \"agg_gbs_per_direction\": %.1f, \"agg_gbs_total\": %.1f, \"per_gpu_gbs_per_direction\": [", first ? "" : ",",
                   g, c.name, (double)bytes * g / best / 1e9, (double)bytes * g * dirs / best / 1e9);
            for (int d = 0; d < g; ++d) printf("%s%.1f", d ? ", " : "", per_gpu[(size_t)d]);
            printf("]}\n");
            first = false;
            fflush(stdout);
        }
    }
    printf("]}\n");
    return 0;
}
