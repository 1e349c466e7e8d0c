// aim_align_file(): `host <pairs-file> <out-file> <N>` as ONE streaming pipeline (tools/host.cpp calls it).
//
// Replaces, for the NW / SWG / WFA programs, everything between the argument checks and the final fclose of the reference
// host (WFA/DPU-MRAM/host/host.c:196-353): get_reads per DPU, the four pushes, the launch, the two pulls and the print loop.
// The host threads only move bytes; parsing and printing are GPU kernels (aim_file.cu):
//
//   reader (caller's thread + I/O threads)  pread() the next chunk of the pair file into a pinned buffer, count its newlines
//                                           and cut it after a whole number of line pairs (the next chunk starts there)
//   GPU c % ngpus, three streams            H2D of the file bytes | parse -> rows, alignment kernels, text formatting | D2H
//   writer thread                           in chunk order: learn the chunk's text size, fetch the text, pwrite() it
//
// Chunks are independent (each starts at a pattern line and numbers its pairs from the running total), so with several
// GPUs chunk c simply goes to GPU c % ngpus and the writer restores the order.  Pairs processed = the reference's
// min(pairs in file, NR_DPUS * roundup8(N / NR_DPUS)) (host.c:191,201-209).
#include <cuda_runtime.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "aim_internal.h"

namespace aim {

namespace {

#define FP_CUDA(call)                                                                          \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            set_error(std::string(#call) + ": " + cudaGetErrorString(e__));                    \
            return AIM_ERR_CUDA;                                                               \
        }                                                                                      \
    } while (0)

constexpr int kSlots = 4;

struct Slot {
    int device = 0;
    // host (pinned)
    char *h_in = nullptr;
    char *h_out = nullptr;
    size_t h_out_cap = 0;
    uint32_t *h_counters = nullptr;  // [0] newlines found, [1] too-long flag, [2] text bytes, [3] OR of 1 << status
    // device
    char *d_in = nullptr, *d_pat = nullptr, *d_txt = nullptr, *d_ops = nullptr, *d_out = nullptr;
    size_t d_out_cap = 0;
    int32_t *d_plen = nullptr, *d_tlen = nullptr;
    aim_result *d_res = nullptr;
    uint32_t *d_nl_pos = nullptr, *d_tiles = nullptr, *d_counters = nullptr, *d_lens = nullptr, *d_offs = nullptr;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // h2d start/done, kernels start/done, d2h start/done
    // the chunk in the slot
    uint32_t pairs = 0;
    size_t text_bytes = 0;
    bool busy = false;
};

struct Gpu {
    int device = 0;
    Scratch *scratch = nullptr;
    std::mutex *mu = nullptr;
    cudaStream_t s_h2d = nullptr, s_kernel = nullptr, s_d2h = nullptr, s_text = nullptr;  // s_text: the writer's text fetches
    Slot slot[kSlots];
};

struct Pipe {
    aim_params p{};
    size_t chunk_bytes = 0;
    uint32_t cap_pairs = 0;
    std::vector<Gpu> gpu;
    int fd_out = -1;
    int write_threads = 8;
    int fmt_mode = 0;  // output line format: 0 NW/SWG/WFA, 1 GenASM-DC, 2 GenASM-filter
    bool out_mappable = true, out_seekable = true, map_populate = false;
    // reader -> writer
    std::mutex m;
    std::condition_variable cv;
    std::deque<std::pair<int, int>> order;   // (gpu, slot) in chunk order: reader -> fetch stage
    std::deque<std::pair<int, int>> order2;  // fetch stage -> file stage
    bool reader_done = false, fetch_done = false;
    int rc = AIM_OK;  // first failure (either side)
    std::string err;
    uint32_t status_or = 0;
    double ph[3] = {0, 0, 0};
    uint64_t out_bytes = 0;
    int launches = 0;
    // host-side time per stage (AIM_VERBOSE): reader {slot wait, pread + newline count, cut + enqueue}, writer {chunk wait, text D2H, pwrite}
    double t_reader[3] = {0, 0, 0}, t_writer[3] = {0, 0, 0};
};

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

void free_slot(Slot &s)
{
    cudaSetDevice(s.device);
    cudaFreeHost(s.h_in); cudaFreeHost(s.h_out); cudaFreeHost(s.h_counters);
    cudaFree(s.d_in); cudaFree(s.d_pat); cudaFree(s.d_txt); cudaFree(s.d_ops); cudaFree(s.d_out); cudaFree(s.d_plen); cudaFree(s.d_tlen);
    cudaFree(s.d_res); cudaFree(s.d_nl_pos); cudaFree(s.d_tiles); cudaFree(s.d_counters); cudaFree(s.d_lens); cudaFree(s.d_offs);
    for (auto &e : s.ev) if (e) cudaEventDestroy(e);
    s = Slot();
}

int alloc_slot(Slot &s, int device, const Pipe &P)
{
    s.device = device;
    const size_t rs = (size_t)P.p.read_size, cap = P.cap_pairs;
    const size_t in_bytes = P.chunk_bytes + 64;
    FP_CUDA(cudaHostAlloc(&s.h_in, in_bytes, cudaHostAllocPortable));
    FP_CUDA(cudaHostAlloc(&s.h_counters, 64, cudaHostAllocPortable));
    FP_CUDA(cudaMalloc(&s.d_in, in_bytes));
    FP_CUDA(cudaMalloc(&s.d_plen, cap * 4));
    FP_CUDA(cudaMalloc(&s.d_tlen, cap * 4));
    FP_CUDA(cudaMalloc(&s.d_pat, cap * rs));
    FP_CUDA(cudaMalloc(&s.d_txt, cap * rs));
    FP_CUDA(cudaMalloc(&s.d_res, cap * sizeof(aim_result)));
    if (P.p.backtrace) FP_CUDA(cudaMalloc(&s.d_ops, cap * 2 * rs));
    FP_CUDA(cudaMalloc(&s.d_nl_pos, (2 * cap + 16) * 4));
    FP_CUDA(cudaMalloc(&s.d_tiles, std::max(file_parse_scratch_bytes(in_bytes), file_format_scratch_bytes((uint32_t)cap))));
    FP_CUDA(cudaMalloc(&s.d_counters, 64));
    FP_CUDA(cudaMalloc(&s.d_lens, cap * 4));
    FP_CUDA(cudaMalloc(&s.d_offs, cap * 4));
    // text: "%d, %d, \n" is at most 27 bytes; a CIGAR is usually a few runs.  Grown on demand (fmt_write_kernel's guard).
    s.d_out_cap = s.h_out_cap = cap * (size_t)(P.p.backtrace ? 96 : 32) + 4096;  // (grown on demand: fmt_write_kernel's guard)
    FP_CUDA(cudaMalloc(&s.d_out, s.d_out_cap));
    FP_CUDA(cudaHostAlloc(&s.h_out, s.h_out_cap, cudaHostAllocPortable));
    for (auto &e : s.ev) FP_CUDA(cudaEventCreate(&e));
    return AIM_OK;
}

// I/O threads of one aim_align_file call: a chunk is read in 1 MiB blocks pulled from an atomic counter (no stragglers), every
// block's newlines are counted while it is hot in the reading core's cache.
struct IoPool {
    std::vector<std::thread> th;
    std::mutex m;
    std::condition_variable cv_job, cv_done;
    // the current job
    int fd = -1;
    char *buf = nullptr;
    size_t n = 0;
    uint64_t off = 0;
    std::atomic<size_t> next{0};
    std::atomic<size_t> newlines{0};
    std::atomic<int> bad{0};
    uint64_t job_id = 0;
    int running = 0;
    bool quit = false;
    static constexpr size_t kBlock = (size_t)1 << 20;

    void worker()
    {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(m);
                cv_job.wait(lk, [&] { return quit || job_id != seen; });
                if (quit) return;
                seen = job_id;
            }
            size_t local = 0;
            for (;;) {
                const size_t a = next.fetch_add(kBlock);
                if (a >= n) break;
                const size_t b = std::min(n, a + kBlock);
                size_t done = a;
                while (done < b) {
                    const ssize_t r = pread(fd, buf + done, b - done, (off_t)(off + done));
                    if (r <= 0) { bad.store(1); break; }
                    done += (size_t)r;
                }
                if (done < b) break;
                local += count_newlines(buf + a, b - a);
            }
            newlines.fetch_add(local);
            {
                std::lock_guard<std::mutex> lk(m);
                --running;
            }
            cv_done.notify_all();
        }
    }
    void start(int T)
    {
        for (int t = 0; t < T; ++t) th.emplace_back([this] { worker(); });
    }
    int read(int fd_, char *buf_, size_t n_, uint64_t off_, size_t *nl)
    {
        {
            std::lock_guard<std::mutex> lk(m);
            fd = fd_; buf = buf_; n = n_; off = off_;
            next.store(0); newlines.store(0); bad.store(0);
            running = (int)th.size();
            ++job_id;
        }
        cv_job.notify_all();
        std::unique_lock<std::mutex> lk(m);
        cv_done.wait(lk, [&] { return running == 0; });
        if (bad.load()) { set_error("pair file: read failed"); return AIM_ERR_IO; }
        *nl = newlines.load();
        return AIM_OK;
    }
    ~IoPool()
    {
        {
            std::lock_guard<std::mutex> lk(m);
            quit = true;
        }
        cv_job.notify_all();
        for (auto &t : th) t.join();
    }
};

// Chunk slots are kept from call to call (pinned and device allocations cost hundreds of milliseconds); aim_shutdown frees them.
struct FileCache {
    std::mutex mu;  // one aim_align_file at a time
    std::vector<Gpu> gpu;
    size_t chunk_bytes = 0;
    uint32_t cap_pairs = 0;
    int read_size = 0, backtrace = 0, first_device = -1;
};
FileCache g_cache;

void cache_release(FileCache &C)
{
    for (Gpu &G : C.gpu) {
        cudaSetDevice(G.device);
        cudaDeviceSynchronize();
        for (Slot &s : G.slot) free_slot(s);
        if (G.s_h2d) cudaStreamDestroy(G.s_h2d);
        if (G.s_kernel) cudaStreamDestroy(G.s_kernel);
        if (G.s_d2h) cudaStreamDestroy(G.s_d2h);
        if (G.s_text) cudaStreamDestroy(G.s_text);
    }
    C.gpu.clear();
    C.first_device = -1;
}

// The writer: chunks in order -> output file.
// The writer, in two stages so that fetching chunk c+1's text overlaps copying chunk c's into the file:
//   fetch_main: chunks in order -> wait for the chunk, check its flags, D2H of the text into the slot's pinned buffer
//   file_main:  chunks in order -> the text into the output file, slot released
void fetch_main(Pipe *P)
{
    for (;;) {
        std::pair<int, int> it;
        {
            std::unique_lock<std::mutex> lk(P->m);
            P->cv.wait(lk, [&] { return !P->order.empty() || P->reader_done; });
            if (P->order.empty()) break;
            it = P->order.front();
            P->order.pop_front();
        }
        Gpu &G = P->gpu[(size_t)it.first];
        Slot &S = G.slot[it.second];
        int rc = AIM_OK;
        std::string err;
        auto cu = [&](cudaError_t e, const char *what) { if (e != cudaSuccess && rc == AIM_OK) { rc = AIM_ERR_CUDA; err = std::string(what) + ": " + cudaGetErrorString(e); } };
        const double tw0 = now_s();
        cu(cudaSetDevice(G.device), "cudaSetDevice");
        cu(cudaEventSynchronize(S.ev[5]), "chunk sync");
        const double tw1 = now_s();
        P->t_writer[0] += tw1 - tw0;
        bool failed_before;
        { std::lock_guard<std::mutex> lk(P->m); failed_before = P->rc != AIM_OK; }
        S.text_bytes = 0;
        if (rc == AIM_OK && !failed_before) {
            const uint32_t lines = 2 * S.pairs;
            if (S.h_counters[1]) { rc = AIM_ERR_LENGTH; err = "READ LENGTH less than length of the input reads"; }
            else if (S.h_counters[0] < lines) { rc = AIM_ERR_CUDA; err = "file pipeline: the device found fewer lines than the host"; }
            else {
                const size_t total = S.h_counters[2];
                if (total > S.d_out_cap) {  // grow the text buffers and write the text again
                    cudaFree(S.d_out); S.d_out = nullptr;
                    cudaFreeHost(S.h_out); S.h_out = nullptr;
                    S.d_out_cap = S.h_out_cap = total + total / 4 + 4096;
                    cu(cudaMalloc(&S.d_out, S.d_out_cap), "cudaMalloc(text)");
                    cu(cudaHostAlloc(&S.h_out, S.h_out_cap, cudaHostAllocPortable), "cudaHostAlloc(text)");
                    if (rc == AIM_OK && launch_file_format_write(S.d_res, S.d_ops, P->p.read_size, P->p.backtrace, P->fmt_mode, S.pairs, S.d_offs, S.d_counters, S.d_out,
                                                                 S.d_out_cap, G.s_text, nullptr) != AIM_OK) { rc = AIM_ERR_CUDA; err = aim_last_error(); }
                }
                if (rc == AIM_OK && total) {
                    cu(cudaMemcpyAsync(S.h_out, S.d_out, total, cudaMemcpyDeviceToHost, G.s_text), "text D2H");
                    cu(cudaStreamSynchronize(G.s_text), "text D2H sync");
                }
                if (rc == AIM_OK) S.text_bytes = total;
                float t;
                for (int k = 0; k < 3; ++k) { if (cudaEventElapsedTime(&t, S.ev[2 * k], S.ev[2 * k + 1]) == cudaSuccess) P->ph[k] += t; }
                std::lock_guard<std::mutex> lk(P->m);
                P->status_or |= S.h_counters[3];
            }
        }
        P->t_writer[1] += now_s() - tw1;
        {
            std::lock_guard<std::mutex> lk(P->m);
            if (rc != AIM_OK && P->rc == AIM_OK) { P->rc = rc; P->err = err; }
            P->order2.push_back(it);
        }
        P->cv.notify_all();
    }
    {
        std::lock_guard<std::mutex> lk(P->m);
        P->fetch_done = true;
    }
    P->cv.notify_all();
}

void file_main(Pipe *P)
{
    uint64_t file_off = 0;
    for (;;) {
        std::pair<int, int> it;
        bool failed;
        {
            std::unique_lock<std::mutex> lk(P->m);
            P->cv.wait(lk, [&] { return !P->order2.empty() || P->fetch_done; });
            if (P->order2.empty()) break;
            it = P->order2.front();
            P->order2.pop_front();
            failed = P->rc != AIM_OK;
        }
        Slot &S = P->gpu[(size_t)it.first].slot[it.second];
        int rc = AIM_OK;
        std::string err;
        const size_t total = failed ? 0 : S.text_bytes;
        const double tw2 = now_s();
        if (total) {
            // Buffered write()s to one file are serialised by the kernel (about 3 GB/s into fresh page-cache pages whatever
            // the thread count), so the text is copied through a shared mapping of the file's next `total` bytes by a few
            // threads; files that cannot be mapped (pipes, devices) get plain writes.
            bool mapped = false;
            if (P->out_mappable && ftruncate(P->fd_out, (off_t)(file_off + total)) == 0) {
                const uint64_t page = 4096, map_off = file_off & ~(page - 1);
                const size_t len = (size_t)(file_off + total - map_off);
                void *m = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_SHARED | (P->map_populate ? MAP_POPULATE : 0), P->fd_out, (off_t)map_off);
                if (m != MAP_FAILED) {
                    char *dst = (char *)m + (file_off - map_off);
                    const int W = (int)std::max<size_t>(1, std::min<size_t>((size_t)P->write_threads, total / ((size_t)1 << 20) + 1));
                    auto piece = [&](int t) {
                        const size_t a = total / (size_t)W * (size_t)t, b = t == W - 1 ? total : total / (size_t)W * (size_t)(t + 1);
                        memcpy(dst + a, S.h_out + a, b - a);
                    };
                    std::vector<std::thread> th;
                    for (int t = 1; t < W; ++t) th.emplace_back(piece, t);
                    piece(0);
                    for (auto &x : th) x.join();
                    munmap(m, len);
                    mapped = true;
                } else {
                    P->out_mappable = false;
                }
            }
            size_t done = mapped ? total : 0;
            while (done < total) {
                ssize_t w = P->out_seekable ? pwrite(P->fd_out, S.h_out + done, total - done, (off_t)(file_off + done))
                                            : write(P->fd_out, S.h_out + done, total - done);
                if (w <= 0) { rc = AIM_ERR_IO; err = "output file: write failed"; break; }
                done += (size_t)w;
            }
            file_off += total;
        }
        P->t_writer[2] += now_s() - tw2;
        {
            std::lock_guard<std::mutex> lk(P->m);
            if (rc != AIM_OK && P->rc == AIM_OK) { P->rc = rc; P->err = err; }
            S.busy = false;
            P->out_bytes = file_off;
        }
        P->cv.notify_all();
    }
}

}  // namespace

void file_pipeline_shutdown()
{
    std::lock_guard<std::mutex> lk(g_cache.mu);
    cache_release(g_cache);
}

}  // namespace aim

using namespace aim;

extern "C" int aim_align_file(const aim_params *params, const char *pairs_path, const char *out_path, uint32_t n_arg, uint32_t nr_dpus,
                              uint64_t *pairs_done, uint32_t *status_mask, double phase_ms[3], int32_t *launches)
{
    if (pairs_done) *pairs_done = 0;
    if (status_mask) *status_mask = 0;
    if (phase_ms) phase_ms[0] = phase_ms[1] = phase_ms[2] = 0.0;
    if (!params || !pairs_path || !out_path) { set_error("NULL argument"); return AIM_ERR_ARG; }
    if (!params_valid_for_file(params)) return AIM_ERR_ARG;
    const int ndev = aim_device_count();
    if (ndev == 0) { set_error("no CUDA device (aim_b200 has no CPU fallback)"); return AIM_ERR_NO_DEVICE; }
    const int g = params->ngpus <= 1 ? 1 : params->ngpus;
    if (params->device < 0 || params->device + g > ndev) { set_error("device range exceeds visible GPUs"); return AIM_ERR_ARG; }
    if (nr_dpus == 0) nr_dpus = 1;
    const uint64_t want = (uint64_t)(((uint64_t)(n_arg / nr_dpus) + 7) / 8 * 8) * nr_dpus;  // host.c:191,201-209

    const double t_setup0 = now_s();
    const int fd = open(pairs_path, O_RDONLY);
    if (fd < 0) { set_error(std::string("Input file '") + pairs_path + "' couldn't be opened"); return AIM_ERR_IO; }
    struct stat st;
    if (fstat(fd, &st) != 0) { close(fd); set_error("fstat failed"); return AIM_ERR_IO; }
    const uint64_t file_size = (uint64_t)st.st_size;
    const int fd_out = open(out_path, O_RDWR | O_CREAT | O_TRUNC, 0644);
    if (fd_out < 0) { close(fd); set_error(std::string("Output file '") + out_path + "' couldn't be opened"); return AIM_ERR_IO; }

    Pipe P;
    P.p = params_normalized(params);
    P.p.ngpus = 1;
    P.fmt_mode = P.p.algo == AIM_ALGO_GENASM_DC ? 1 : P.p.algo == AIM_ALGO_GENASM_FILTER ? 2 : 0;
    P.fd_out = fd_out;
    {
        struct stat so;
        P.out_seekable = lseek(fd_out, 0, SEEK_CUR) != (off_t)-1;
        P.out_mappable = P.out_seekable && fstat(fd_out, &so) == 0 && S_ISREG(so.st_mode) && !getenv("AIM_FILE_NO_MMAP");
    }
    const size_t rs = (size_t)params->read_size;
    {   // chunk: tens of thousands of resident pair slots deep for short reads, several pairs per slot for long reads
        size_t mb = rs >= 2048 ? 256 : 96;
        if (const char *e = getenv("AIM_FILE_CHUNK_MB")) { const long v = atol(e); if (v >= 1 && v <= 1024) mb = (size_t)v; }
        P.chunk_bytes = std::min<size_t>(mb << 20, std::max<size_t>((size_t)file_size / (size_t)(2 * g) + 4096, (size_t)1 << 20));
        P.chunk_bytes = (P.chunk_bytes + 4095) & ~(size_t)4095;
        // pairs a chunk may hold: lines of about READ_SIZE bytes are expected; shorter ones make the reader cut earlier
        P.cap_pairs = (uint32_t)std::max<size_t>(1024, P.chunk_bytes / std::max<size_t>(16, rs / 2));
    }
    int io_threads = (int)std::min(32u, std::max(1u, std::thread::hardware_concurrency()));
    if (const char *e = getenv("AIM_IO_THREADS")) { const int v = atoi(e); if (v >= 1) io_threads = std::min(v, 64); }

    int rc = AIM_OK;
    std::lock_guard<std::mutex> cache_lock(g_cache.mu);
    {   // chunk slots: reuse the cached ones when they fit this call
        FileCache &C = g_cache;
        const bool fits = C.first_device == params->device && (int)C.gpu.size() == g && C.read_size == params->read_size &&
                          C.backtrace >= (P.p.backtrace ? 1 : 0) && C.chunk_bytes >= P.chunk_bytes && C.chunk_bytes <= 4 * P.chunk_bytes + ((size_t)8 << 20);
        if (fits) {
            P.chunk_bytes = C.chunk_bytes;
            P.cap_pairs = C.cap_pairs;
        } else {
            cache_release(C);
            C.gpu.resize((size_t)g);
            for (int d = 0; d < g && rc == AIM_OK; ++d) {
                Gpu &G = C.gpu[(size_t)d];
                G.device = params->device + d;
                void *mu = nullptr;
                rc = device_state(G.device, &G.scratch, &mu);
                G.mu = static_cast<std::mutex *>(mu);
                if (rc != AIM_OK) break;
                if (cudaSetDevice(G.device) != cudaSuccess) { set_error("cudaSetDevice failed"); rc = AIM_ERR_CUDA; break; }
                cudaError_t e = cudaStreamCreateWithFlags(&G.s_h2d, cudaStreamNonBlocking);
                if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&G.s_kernel, cudaStreamNonBlocking);
                if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&G.s_d2h, cudaStreamNonBlocking);
                if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&G.s_text, cudaStreamNonBlocking);
                if (e != cudaSuccess) { set_error(std::string("stream: ") + cudaGetErrorString(e)); rc = AIM_ERR_CUDA; break; }
                for (int k = 0; k < kSlots && rc == AIM_OK; ++k) rc = alloc_slot(G.slot[k], G.device, P);
            }
            if (rc == AIM_OK) {
                C.first_device = params->device;
                C.read_size = params->read_size;
                C.backtrace = P.p.backtrace ? 1 : 0;
                C.chunk_bytes = P.chunk_bytes;
                C.cap_pairs = P.cap_pairs;
            } else {
                const std::string keep = aim_last_error();
                cache_release(C);
                set_error(keep);
            }
        }
        P.gpu = std::move(C.gpu);  // handed back below
        C.gpu.clear();
    }
    IoPool pool;
    pool.start(io_threads);
    P.write_threads = std::max(1, std::min(8, io_threads / 2));
    if (const char *e = getenv("AIM_WRITE_THREADS")) { const int v = atoi(e); if (v >= 1) P.write_threads = std::min(v, 32); }
    if (const char *e = getenv("AIM_FILE_POPULATE")) P.map_populate = atoi(e) != 0;

    const double t_setup1 = now_s();
    std::thread writer, filer;
    if (rc == AIM_OK) { writer = std::thread(fetch_main, &P); filer = std::thread(file_main, &P); }
    uint64_t off = 0, pairs_total = 0, chunk_no = 0;
    int nlaunch = 0;
    while (rc == AIM_OK && off < file_size && pairs_total < want) {
        const int gi = (int)(chunk_no % (uint64_t)g), si = (int)((chunk_no / (uint64_t)g) % kSlots);
        Gpu &G = P.gpu[(size_t)gi];
        Slot &S = G.slot[si];
        const double tr0 = now_s();
        {   // the slot's previous chunk must have left through the writer
            std::unique_lock<std::mutex> lk(P.m);
            P.cv.wait(lk, [&] { return !S.busy || P.rc != AIM_OK; });
            if (P.rc != AIM_OK) break;
        }
        const double tr1 = now_s();
        P.t_reader[0] += tr1 - tr0;
        const size_t got = (size_t)std::min<uint64_t>(P.chunk_bytes, file_size - off);
        size_t newlines = 0;
        rc = pool.read(fd, S.h_in, got, off, &newlines);
        if (rc != AIM_OK) break;
        const double tr2 = now_s();
        P.t_reader[1] += tr2 - tr1;
        const bool at_eof = off + got == file_size;
        // lines as getline sees them: every '\n' ends one; at end of file an unterminated rest is a line too
        const bool tail_line = at_eof && got > 0 && S.h_in[got - 1] != '\n';
        size_t lines = newlines + (tail_line ? 1 : 0);
        uint64_t pairs = lines / 2;
        pairs = std::min<uint64_t>(pairs, std::min<uint64_t>(want - pairs_total, P.cap_pairs));
        if (pairs == 0) {
            if (at_eof) break;  // an odd last line (or nothing): dropped, as get_reads' second getline fails (host.c:108-110)
            // not even one pair in a whole chunk: some line is far longer than READ_SIZE
            set_error("READ LENGTH less than length of the input reads");
            rc = AIM_ERR_LENGTH;
            break;
        }
        // cut after line 2 * pairs (the chunk may end inside a line: those bytes belong to the next chunk)
        size_t cut;            // bytes of this chunk handed to the GPU
        int unterminated = 0;  // the last of those lines has no '\n' (end of file)
        auto after_last_nl = [&](size_t n) { const char *q = (const char *)memrchr(S.h_in, '\n', n); return q ? (size_t)(q - S.h_in) + 1 : (size_t)0; };
        if (2 * pairs == lines) {
            if (tail_line) { cut = got; unterminated = 1; }
            else cut = after_last_nl(got);
        } else if (2 * pairs + 1 == lines) {  // one line too many: the unterminated rest, or the last complete line
            cut = after_last_nl(got);
            if (!tail_line) cut = after_last_nl(cut - 1);
        } else {  // fewer pairs than the chunk holds (N reached, or more pairs than the slot's buffers): walk to newline 2 * pairs
            const char *q = S.h_in;
            for (uint64_t k = 0; k < 2 * pairs; ++k) q = (const char *)memchr(q, '\n', (size_t)(S.h_in + got - q)) + 1;
            cut = (size_t)(q - S.h_in);
        }
        size_t nbytes = cut;
        if (unterminated) S.h_in[nbytes++] = '\n';  // the virtual newline extract_rows_kernel discounts
        memset(S.h_in + nbytes, 0, 16);

        S.pairs = (uint32_t)pairs;
        {
            std::lock_guard<std::mutex> lk(*G.mu);  // launch sequences on this device's scratch are serialised
            if (cudaSetDevice(G.device) != cudaSuccess) { set_error("cudaSetDevice failed"); rc = AIM_ERR_CUDA; break; }
            cudaError_t e = cudaEventRecord(S.ev[0], G.s_h2d);
            if (e == cudaSuccess) e = cudaMemcpyAsync(S.d_in, S.h_in, (nbytes + 31) & ~(size_t)15, cudaMemcpyHostToDevice, G.s_h2d);
            if (e == cudaSuccess) e = cudaEventRecord(S.ev[1], G.s_h2d);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(G.s_kernel, S.ev[1], 0);
            if (e == cudaSuccess) e = cudaEventRecord(S.ev[2], G.s_kernel);
            if (e == cudaSuccess) e = cudaMemsetAsync(S.d_counters, 0, 64, G.s_kernel);
            if (e != cudaSuccess) { set_error(std::string("file pipeline enqueue: ") + cudaGetErrorString(e)); rc = AIM_ERR_CUDA; break; }
            rc = launch_file_parse(S.d_in, nbytes, (uint32_t)(2 * pairs), unterminated, P.p.read_size, S.d_tiles, S.d_nl_pos, S.d_counters, S.d_plen,
                                   S.d_tlen, S.d_pat, S.d_txt, G.s_kernel, &nlaunch);
            if (rc != AIM_OK) break;
            KernelArgs a{P.p, (uint32_t)pairs, (uint32_t)pairs_total, S.d_plen, S.d_tlen, S.d_pat, S.d_txt, S.d_res, P.p.backtrace ? S.d_ops : nullptr};
            rc = launch_algo(a, G.scratch, G.s_kernel, &nlaunch);
            if (rc != AIM_OK) break;
            rc = launch_file_format(S.d_res, S.d_ops, P.p.read_size, P.p.backtrace, P.fmt_mode, (uint32_t)pairs, S.d_lens, S.d_offs, S.d_tiles, S.d_counters, S.d_out,
                                    S.d_out_cap, G.s_kernel, &nlaunch);
            if (rc != AIM_OK) break;
            e = cudaEventRecord(S.ev[3], G.s_kernel);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(G.s_d2h, S.ev[3], 0);
            if (e == cudaSuccess) e = cudaEventRecord(S.ev[4], G.s_d2h);
            if (e == cudaSuccess) e = cudaMemcpyAsync(S.h_counters, S.d_counters, 64, cudaMemcpyDeviceToHost, G.s_d2h);
            if (e == cudaSuccess) e = cudaEventRecord(S.ev[5], G.s_d2h);
            if (e != cudaSuccess) { set_error(std::string("file pipeline enqueue: ") + cudaGetErrorString(e)); rc = AIM_ERR_CUDA; break; }
        }
        {
            std::lock_guard<std::mutex> lk(P.m);
            S.busy = true;
            P.order.emplace_back(gi, si);
        }
        P.cv.notify_all();
        pairs_total += pairs;
        off += cut;
        ++chunk_no;
        P.t_reader[2] += now_s() - tr2;
    }
    {
        std::lock_guard<std::mutex> lk(P.m);
        if (rc != AIM_OK && P.rc == AIM_OK) { P.rc = rc; P.err = aim_last_error(); }
        P.reader_done = true;
    }
    P.cv.notify_all();
    if (writer.joinable()) writer.join();
    if (filer.joinable()) filer.join();
    const double t_stream1 = now_s();
    rc = P.rc;
    for (Gpu &G : P.gpu) {  // drain, then hand the slots back to the cache
        cudaSetDevice(G.device);
        cudaDeviceSynchronize();
        for (Slot &s : G.slot) { s.busy = false; s.pairs = 0; }
    }
    g_cache.gpu = std::move(P.gpu);
    if (rc != AIM_OK && rc != AIM_ERR_LENGTH) cache_release(g_cache);  // after a CUDA failure nothing is kept
    close(fd);
    if (getenv("AIM_VERBOSE"))
        fprintf(stderr, "aim_align_file: %llu chunks of %zu MiB, %d I/O threads; setup %.1f ms, stream %.1f ms, teardown %.1f ms; reader: slot wait %.1f, pread+count %.1f, "
                "cut+enqueue %.1f ms; writer: chunk wait %.1f, text D2H %.1f, pwrite %.1f ms\n", (unsigned long long)chunk_no, P.chunk_bytes >> 20, io_threads,
                (t_setup1 - t_setup0) * 1e3, (t_stream1 - t_setup1) * 1e3, (now_s() - t_stream1) * 1e3, P.t_reader[0] * 1e3, P.t_reader[1] * 1e3, P.t_reader[2] * 1e3,
                P.t_writer[0] * 1e3, P.t_writer[1] * 1e3, P.t_writer[2] * 1e3);
    // the reference writes nothing when it exits before its print loop (a too-long read: host.c:119-123; a DPU fault)
    const bool fatal_status = (P.status_or & ((1u << AIM_STATUS_BACKTRACE) | (1u << AIM_STATUS_ARENA))) != 0;
    if (rc != AIM_OK || fatal_status) { if (ftruncate(fd_out, 0) != 0) { /* best effort */ } }
    close(fd_out);
    if (rc != AIM_OK) { set_error(P.err); return rc; }
    if (pairs_done) *pairs_done = pairs_total;
    if (status_mask) *status_mask = P.status_or;
    if (phase_ms) for (int k = 0; k < 3; ++k) phase_ms[k] = P.ph[k];
    if (launches) *launches = nlaunch;
    return AIM_OK;
}
