// The device boundary: C-ABI entry points, per-GPU contexts, pinned double-buffered transfers.
//
// Replaces the dpu_alloc / dpu_prepare_xfer / dpu_push_xfer / dpu_launch sequence of the
// reference host (WFA/DPU-MRAM/host/host.c:186-330): the four input pushes (:246-268), the
// synchronous launch (:289) and the two result pulls (:316-326) become, per GPU, a chunked
// pipeline on two CUDA streams so that chunk c+1's upload, chunk c's alignment and chunk c-1's
// CIGAR download overlap.  Pairs shard contiguously over GPUs exactly as the reference shards
// them over DPUs (host.c:201-209); no collective is involved.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "aim_internal.h"

namespace aim {

namespace {

#define AIM_CUDA(call)                                                                         \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            set_error(std::string(#call) + ": " + cudaGetErrorString(e__));                    \
            return AIM_ERR_CUDA;                                                               \
        }                                                                                      \
    } while (0)

constexpr int kNumBuf = 4;  // chunk buffers in flight: upload / align / download of different chunks overlap

struct ChunkBuf {
    // device
    int32_t *d_plen = nullptr, *d_tlen = nullptr;
    char *d_pat = nullptr, *d_txt = nullptr, *d_ops = nullptr, *d_cig = nullptr;
    size_t cig_cap = 0;  // bytes of d_cig (aim_align_batch_cigars only)
    aim_result *d_res = nullptr;
    // pinned staging (used only for caller buffers that are not pinned)
    int32_t *h_plen = nullptr, *h_tlen = nullptr;
    char *h_pat = nullptr, *h_txt = nullptr, *h_ops = nullptr;
    aim_result *h_res = nullptr;
    unsigned char *h_runs = nullptr;  // run rows of the chunk as downloaded (op rows are rebuilt from them on the host)
    size_t h_runs_cap = 0;
    // stage boundaries: [0] h2d start, [1] h2d done, [2] kernel start, [3] kernel done, [4] d2h start, [5] d2h done
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    uint32_t pairs_cap = 0;
    int32_t read_size = 0;
    bool has_ops = false, has_staging = false;
};

struct DeviceCtx {
    int device = -1;
    bool ready = false;
    Scratch scratch;
    ChunkBuf chunk[kNumBuf];
    cudaStream_t s_h2d = nullptr, s_kernel = nullptr, s_d2h = nullptr;  // one stream per pipeline stage
    cudaStream_t s_fix = nullptr;  // op rows whose runs did not fit their run row, fetched as they are
    std::mutex mu;
};

std::mutex g_mu;
std::vector<DeviceCtx *> g_ctx;

int get_ctx(int device, DeviceCtx **out)
{
    std::lock_guard<std::mutex> lk(g_mu);
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_error(std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "count = 0") +
                  " (aim_b200 has no CPU fallback)");
        return AIM_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= count) { set_error("device ordinal out of range"); return AIM_ERR_ARG; }
    if ((int)g_ctx.size() < count) g_ctx.resize((size_t)count, nullptr);
    if (!g_ctx[(size_t)device]) {
        DeviceCtx *c = new DeviceCtx();
        c->device = device;
        cudaDeviceProp prop;
        AIM_CUDA(cudaGetDeviceProperties(&prop, device));
        if (prop.major < 10) {
            set_error("device is not sm_100 class; this library carries sm_100a code only");
            delete c;
            return AIM_ERR_NO_DEVICE;
        }
        c->scratch.sm_count = prop.multiProcessorCount;
        c->scratch.device = device;
        c->ready = true;
        g_ctx[(size_t)device] = c;
    }
    *out = g_ctx[(size_t)device];
    return AIM_OK;
}

void free_chunk(ChunkBuf &b)
{
    cudaFree(b.d_plen); cudaFree(b.d_tlen); cudaFree(b.d_pat); cudaFree(b.d_txt); cudaFree(b.d_ops); cudaFree(b.d_res); cudaFree(b.d_cig);
    cudaFreeHost(b.h_plen); cudaFreeHost(b.h_tlen); cudaFreeHost(b.h_pat); cudaFreeHost(b.h_txt);
    cudaFreeHost(b.h_ops); cudaFreeHost(b.h_res); cudaFreeHost(b.h_runs);
    for (auto &e : b.ev) if (e) cudaEventDestroy(e);
    b = ChunkBuf();
}

int ensure_chunk(ChunkBuf &b, uint32_t pairs, int32_t read_size, bool ops, bool staging)
{
    if (b.pairs_cap >= pairs && b.read_size == read_size && (b.has_ops || !ops) && (b.has_staging || !staging)) return AIM_OK;
    free_chunk(b);
    const size_t rs = (size_t)read_size;
    // (blocking sync: a host thread that waits for a chunk sleeps instead of spinning - the cores are needed by the threads
    // that rebuild op rows, and by the other ranks of a multi-GPU job)
    for (auto &e : b.ev) AIM_CUDA(cudaEventCreateWithFlags(&e, cudaEventBlockingSync));
    AIM_CUDA(cudaMalloc(&b.d_plen, pairs * sizeof(int32_t)));
    AIM_CUDA(cudaMalloc(&b.d_tlen, pairs * sizeof(int32_t)));
    AIM_CUDA(cudaMalloc(&b.d_pat, pairs * rs));
    AIM_CUDA(cudaMalloc(&b.d_txt, pairs * rs));
    AIM_CUDA(cudaMalloc(&b.d_res, pairs * sizeof(aim_result)));
    if (ops) AIM_CUDA(cudaMalloc(&b.d_ops, pairs * 2 * rs));
    if (staging) {
        AIM_CUDA(cudaHostAlloc(&b.h_plen, pairs * sizeof(int32_t), cudaHostAllocDefault));
        AIM_CUDA(cudaHostAlloc(&b.h_tlen, pairs * sizeof(int32_t), cudaHostAllocDefault));
        AIM_CUDA(cudaHostAlloc(&b.h_pat, pairs * rs, cudaHostAllocDefault));
        AIM_CUDA(cudaHostAlloc(&b.h_txt, pairs * rs, cudaHostAllocDefault));
        AIM_CUDA(cudaHostAlloc(&b.h_res, pairs * sizeof(aim_result), cudaHostAllocDefault));
        if (ops) AIM_CUDA(cudaHostAlloc(&b.h_ops, pairs * 2 * rs, cudaHostAllocDefault));
    }
    b.pairs_cap = pairs;
    b.read_size = read_size;
    b.has_ops = ops;
    b.has_staging = staging;
    return AIM_OK;
}

bool is_pinned(const void *p)
{
    if (!p) return true;
    cudaPointerAttributes at;
    cudaError_t e = cudaPointerGetAttributes(&at, p);
    if (e != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

int validate(const aim_params *p, bool need_ops, const void *ops)
{
    if (!p) { set_error("params is NULL"); return AIM_ERR_ARG; }
    if (p->algo < AIM_ALGO_NW || p->algo > AIM_ALGO_GENASM_FILTER) { set_error("unknown algo"); return AIM_ERR_ARG; }
    if (p->read_size <= 0 || (p->read_size % 8) != 0) { set_error("read_size must be a positive multiple of 8"); return AIM_ERR_ARG; }
    // the run scripts' penalty validation (run-wfa-pim-mram.py:44-46; NW has no gap_ext)
    if (p->match > 0 || p->mismatch <= 0 || p->gap_open <= 0 || (p->algo != AIM_ALGO_NW && p->gap_ext <= 0)) {
        set_error("Wrong affine gap penalties must be  m <= 0 and g, a, x > 0");
        return AIM_ERR_ARG;
    }
    if (p->max_score < 0) { set_error("max_score must be >= 0"); return AIM_ERR_ARG; }
    if (need_ops && p->backtrace && !ops) { set_error("ops buffer required when backtrace is set"); return AIM_ERR_ARG; }
    return AIM_OK;
}

// GenASM-DC always returns its CIGAR string in the ops rows, the filter never has ops
aim_params normalized(const aim_params *p)
{
    aim_params q = *p;
    if (q.algo == AIM_ALGO_GENASM_DC) q.backtrace = 1;
    if (q.algo == AIM_ALGO_GENASM_FILTER) q.backtrace = 0;
    return q;
}

int launch(const KernelArgs &a, Scratch *s, cudaStream_t stream, int *launches)
{
    int rc = scratch_acquire(s, stream);
    if (rc != AIM_OK) return rc;
    if (a.p.algo == AIM_ALGO_WFA) rc = launch_wfa(a, s, stream, launches);
    else if (a.p.algo == AIM_ALGO_GENASM_DC || a.p.algo == AIM_ALGO_GENASM_FILTER) rc = launch_genasm(a, s, stream, launches);
    else rc = launch_dp(a, s, stream, launches);
    const int rc2 = scratch_release(s, stream);
    return rc != AIM_OK ? rc : rc2;
}

// Chunk size (pairs) of the host pipeline: big enough to fill the GPU many times over, small enough that pipeline
// fill/drain is short (the full-table DP kernels hold one pair per thread for milliseconds: give them several times the
// resident thread count per chunk so the last, partially filled pass stays short; long reads: a chunk must hold several
// pairs per resident pair slot, ~6.5 K slots on a B200).  With several GPUs pulling chunks from one queue the batch is cut
// into at least ~4 chunks per GPU so that variable-cost pairs (long reads) balance.
uint32_t pick_chunk_pairs(const aim_params &p, uint32_t n, size_t per_pair, int ngpus)
{
    size_t chunk_bytes = (p.algo != AIM_ALGO_WFA) ? (384u << 20) : (p.read_size >= 2048 ? (768u << 20) : (96u << 20));
    if (const char *cm = getenv("AIM_CHUNK_MB")) { const long v = atol(cm); if (v >= 1 && v <= 4096) chunk_bytes = (size_t)v << 20; }
    uint32_t chunk_pairs = (uint32_t)std::max<size_t>(16384, std::min<size_t>(chunk_bytes / per_pair, 1u << 20));
    // a batch of a few such chunks (config 2: 2 M pairs = 3) spends as long filling and draining the pipeline as in it: at least
    // ~8 chunks per call, none below 256 K pairs (a DP kernel wave is ~150 K pairs)
    if (p.algo != AIM_ALGO_WFA && !getenv("AIM_CHUNK_MB")) chunk_pairs = std::min(chunk_pairs, std::max(1u << 18, n / 8u));
    if (ngpus > 1) {
        const uint32_t floor_pairs = p.read_size >= 2048 ? 4096u : 16384u;
        chunk_pairs = std::max(floor_pairs, std::min(chunk_pairs, n / (4u * (uint32_t)ngpus)));
    }
    return std::max(1u, std::min(chunk_pairs, n));
}

// One GPU's work on a host batch of n pairs: chunks of chunk_pairs pairs are PULLED from `queue` (shared by the GPUs of an
// ngpus > 1 call: SURVEY 8e, variable-cost pairs balance themselves; a private counter for a single GPU), each chunk
// goes through upload / align / download on this GPU's three streams, results land at the chunk's offset of the caller's
// arrays - already in pair order.
// cigars != NULL (aim_align_batch_cigars): the op rows stay on the device and pitch-byte CIGAR text rows come back instead.
int run_shard(const aim_params &p, int device, std::atomic<uint32_t> *queue, uint32_t chunk_pairs, uint32_t n, uint32_t idx_base,
              const int32_t *plen, const int32_t *tlen, const char *patterns, const char *texts,
              aim_result *results, char *ops, double phase_ms[3], std::string *err, char *cigars = nullptr, int32_t pitch = 0)
{
    auto fail = [&](int rc) { if (err) *err = aim_last_error(); return rc; };
    DeviceCtx *ctx = nullptr;
    int rc = get_ctx(device, &ctx);
    if (rc != AIM_OK) return fail(rc);
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (cudaSetDevice(device) != cudaSuccess) { set_error("cudaSetDevice failed"); return fail(AIM_ERR_CUDA); }
    if (n == 0) return AIM_OK;

    const size_t rs = (size_t)p.read_size;
    const bool bt = p.backtrace != 0;
    // Op rows come back as RUN rows and are rebuilt by host threads (aim_file.cu op_runs_kernel, aim_host.cpp expand_op_runs):
    // the device-to-host direction is the scarcer one and an op row is a handful of runs.  AIM_SPARSE_OPS=0: as they are.
    int32_t runs_pitch = 0;
    bool strings = false;  // GenASM-DC: the rows hold CIGAR strings; their heads come back instead of run rows
    if (bt && !cigars) {
        const char *e = getenv("AIM_SPARSE_OPS");
        if (!(e && atoi(e) == 0)) {
            runs_pitch = op_rows_download_bytes(p);
            strings = runs_pitch > 0 && p.algo == AIM_ALGO_GENASM_DC;
        }
    }
    const bool sparse = runs_pitch > 0;
    const bool pin_in = is_pinned(plen) && is_pinned(tlen) && is_pinned(patterns) && is_pinned(texts);
    // (CIGAR rows are copied straight into the caller's buffer, pinned or not: no staging copy for them; rebuilt op rows are
    // written by host threads)
    const bool pin_out = cigars ? true : (is_pinned(results) && (!bt || sparse || is_pinned(ops)));
    const bool staging = !(pin_in && pin_out);

    const uint32_t nchunks = (n + chunk_pairs - 1) / chunk_pairs;
    const int nbuf = (int)std::min<uint32_t>(kNumBuf, nchunks);
    if (!ctx->s_h2d) {
        AIM_CUDA(cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking));
        AIM_CUDA(cudaStreamCreateWithFlags(&ctx->s_kernel, cudaStreamNonBlocking));
        AIM_CUDA(cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking));
    }
    for (int b = 0; b < nbuf; ++b) {
        rc = ensure_chunk(ctx->chunk[b], chunk_pairs, p.read_size, bt, staging);
        if (rc != AIM_OK) return fail(rc);
        ChunkBuf &B = ctx->chunk[b];
        const size_t cig_need = (size_t)chunk_pairs * (size_t)(sparse ? runs_pitch : cigars ? pitch : 0);
        if (B.cig_cap < cig_need) {
            cudaFree(B.d_cig);
            B.d_cig = nullptr;
            B.cig_cap = 0;
            AIM_CUDA(cudaMalloc(&B.d_cig, cig_need));
            B.cig_cap = cig_need;
        }
        if (sparse && B.h_runs_cap < cig_need) {
            cudaFreeHost(B.h_runs);
            B.h_runs = nullptr;
            B.h_runs_cap = 0;
            AIM_CUDA(cudaHostAlloc(&B.h_runs, cig_need, cudaHostAllocDefault));
            B.h_runs_cap = cig_need;
        }
    }
    if (sparse && !ctx->s_fix) AIM_CUDA(cudaStreamCreateWithFlags(&ctx->s_fix, cudaStreamNonBlocking));
    double ph[3] = {0, 0, 0};
    std::vector<uint32_t> mine;  // chunks this GPU pulled, in pull order (the j-th uses buffer j % nbuf)
    mine.reserve(nchunks);       // (the coordinator thread reads entries while later ones are appended)

    auto finish = [&](uint32_t j) -> int {
        ChunkBuf &B = ctx->chunk[j % (uint32_t)nbuf];
        const uint32_t off = mine[j] * chunk_pairs, m = std::min(chunk_pairs, n - off);
        cudaError_t e = cudaEventSynchronize(B.ev[5]);
        if (e != cudaSuccess) { set_error(std::string("chunk sync: ") + cudaGetErrorString(e)); return AIM_ERR_CUDA; }
        if (!pin_out) {
            memcpy(results + off, B.h_res, (size_t)m * sizeof(aim_result));
            if (bt && !cigars && !sparse) memcpy(ops + (size_t)off * 2 * rs, B.h_ops, (size_t)m * 2 * rs);
        }
        if (sparse) {
            std::vector<uint32_t> ov;
            char *o_ops = ops + (size_t)off * 2 * rs;
            if (strings) expand_str_rows(B.h_runs, runs_pitch, m, p.read_size, o_ops, &ov);
            else expand_op_runs(B.h_runs, runs_pitch, m, p.read_size, o_ops, &ov);
            if (!ov.empty()) {  // rows with more runs than a run row holds: fetched as they are (a few: one by one; many: the chunk's rows)
                if (ov.size() <= 64) {
                    for (uint32_t i : ov) {
                        e = cudaMemcpyAsync(o_ops + (size_t)i * 2 * rs, B.d_ops + (size_t)i * 2 * rs, 2 * rs, cudaMemcpyDeviceToHost, ctx->s_fix);
                        if (e != cudaSuccess) break;
                    }
                } else {
                    e = cudaMemcpyAsync(o_ops, B.d_ops, (size_t)m * 2 * rs, cudaMemcpyDeviceToHost, ctx->s_fix);
                }
                if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->s_fix);
                if (e != cudaSuccess) { set_error(std::string("op row fetch: ") + cudaGetErrorString(e)); return AIM_ERR_CUDA; }
            }
        }
        float t;
        for (int k = 0; k < 3; ++k) { cudaEventElapsedTime(&t, B.ev[2 * k], B.ev[2 * k + 1]); ph[k] += t; }
        return AIM_OK;
    };

    // sparse: a coordinator thread completes the chunks in submission order (event wait, row rebuild on the host pool) while
    // this thread keeps submitting; a chunk buffer is reused once its chunk is complete
    struct Coord {
        std::mutex m;
        std::condition_variable cv;
        uint32_t submitted = 0, done = 0;
        bool closing = false;
        int rc = AIM_OK;
        std::string err;
        std::thread th;
        void close()
        {
            if (!th.joinable()) return;
            { std::lock_guard<std::mutex> lk(m); closing = true; }
            cv.notify_all();
            th.join();
        }
        ~Coord() { close(); }
    } co;
    if (sparse) try {
        co.th = std::thread([&]() {
            cudaSetDevice(device);
            for (uint32_t j = 0;; ++j) {
                {
                    std::unique_lock<std::mutex> lk(co.m);
                    co.cv.wait(lk, [&] { return co.submitted > j || co.closing; });
                    if (co.submitted <= j) break;
                }
                const int r = finish(j);
                {
                    std::lock_guard<std::mutex> lk(co.m);
                    if (r != AIM_OK && co.rc == AIM_OK) { co.rc = r; co.err = aim_last_error(); }
                    co.done = j + 1;
                }
                co.cv.notify_all();
            }
        });
    } catch (...) {
        set_error("cannot start the chunk coordinator thread");
        return fail(AIM_ERR_NOMEM);
    }
    auto wait_done = [&](uint32_t j) -> int {  // chunk j complete (sparse)
        std::unique_lock<std::mutex> lk(co.m);
        co.cv.wait(lk, [&] { return co.done > j; });
        if (co.rc != AIM_OK) { set_error(co.err); return co.rc; }
        return AIM_OK;
    };

    // three-stage pipeline on three streams: all uploads in order on s_h2d, all kernels in order on s_kernel
    // (they share the per-device scratch), all downloads in order on s_d2h; events carry the dependencies.
    for (;;) {
        const uint32_t c = queue->fetch_add(1u);
        if (c >= nchunks) break;
        const uint32_t j = (uint32_t)mine.size();
        mine.push_back(c);
        ChunkBuf &B = ctx->chunk[j % (uint32_t)nbuf];
        if (j >= (uint32_t)nbuf) {
            rc = sparse ? wait_done(j - (uint32_t)nbuf) : finish(j - (uint32_t)nbuf);
            if (rc != AIM_OK) { if (sparse) { queue->store(nchunks); cudaDeviceSynchronize(); } return fail(rc); }
        }
        const uint32_t off = c * chunk_pairs, m = std::min(chunk_pairs, n - off);
        const int32_t *s_plen = plen + off, *s_tlen = tlen + off;
        {   // host.c:119-123 rejects reads longer than READ_SIZE (there: message + exit)
            int32_t lo = 0, hi = 0;
            for (uint32_t i = 0; i < m; ++i) {
                lo = std::min(lo, std::min(s_plen[i], s_tlen[i]));
                hi = std::max(hi, std::max(s_plen[i], s_tlen[i]));
            }
            if (lo < 0 || hi > p.read_size) {
                queue->store(nchunks);    // the other GPUs of this call stop pulling
                cudaDeviceSynchronize();  // drain the chunks in flight before the caller reclaims its buffers
                if (lo < 0) { set_error("negative sequence length"); return fail(AIM_ERR_ARG); }
                set_error("READ LENGTH less than length of the input reads");
                return fail(AIM_ERR_LENGTH);
            }
        }
        const char *s_pat = patterns + (size_t)off * rs, *s_txt = texts + (size_t)off * rs;
        if (!pin_in) {
            memcpy(B.h_plen, s_plen, (size_t)m * 4); memcpy(B.h_tlen, s_tlen, (size_t)m * 4);
            memcpy(B.h_pat, s_pat, (size_t)m * rs); memcpy(B.h_txt, s_txt, (size_t)m * rs);
            s_plen = B.h_plen; s_tlen = B.h_tlen; s_pat = B.h_pat; s_txt = B.h_txt;
        }
        AIM_CUDA(cudaEventRecord(B.ev[0], ctx->s_h2d));
        AIM_CUDA(cudaMemcpyAsync(B.d_plen, s_plen, (size_t)m * 4, cudaMemcpyHostToDevice, ctx->s_h2d));
        AIM_CUDA(cudaMemcpyAsync(B.d_tlen, s_tlen, (size_t)m * 4, cudaMemcpyHostToDevice, ctx->s_h2d));
        AIM_CUDA(cudaMemcpyAsync(B.d_pat, s_pat, (size_t)m * rs, cudaMemcpyHostToDevice, ctx->s_h2d));
        AIM_CUDA(cudaMemcpyAsync(B.d_txt, s_txt, (size_t)m * rs, cudaMemcpyHostToDevice, ctx->s_h2d));
        AIM_CUDA(cudaEventRecord(B.ev[1], ctx->s_h2d));

        AIM_CUDA(cudaStreamWaitEvent(ctx->s_kernel, B.ev[1], 0));
        AIM_CUDA(cudaEventRecord(B.ev[2], ctx->s_kernel));
        KernelArgs a{p, m, idx_base + off, B.d_plen, B.d_tlen, B.d_pat, B.d_txt, B.d_res, bt ? B.d_ops : nullptr};
        rc = launch(a, &ctx->scratch, ctx->s_kernel, nullptr);
        if (rc != AIM_OK) { queue->store(nchunks); cudaDeviceSynchronize(); return fail(rc); }
        if (cigars) {
            a.cigars = B.d_cig;
            a.cigar_pitch = pitch;
            rc = launch_cigar_rows(a, ctx->s_kernel, nullptr);
            if (rc != AIM_OK) { queue->store(nchunks); cudaDeviceSynchronize(); return fail(rc); }
        }
        if (sparse) {
            rc = strings ? launch_str_rows(B.d_ops, p.read_size, m, reinterpret_cast<unsigned char *>(B.d_cig), runs_pitch, ctx->s_kernel, nullptr)
                         : launch_op_runs(B.d_ops, p.read_size, m, reinterpret_cast<unsigned char *>(B.d_cig), runs_pitch, ctx->s_kernel, nullptr);
            if (rc != AIM_OK) { queue->store(nchunks); cudaDeviceSynchronize(); return fail(rc); }
        }
        AIM_CUDA(cudaEventRecord(B.ev[3], ctx->s_kernel));

        AIM_CUDA(cudaStreamWaitEvent(ctx->s_d2h, B.ev[3], 0));
        AIM_CUDA(cudaEventRecord(B.ev[4], ctx->s_d2h));
        aim_result *o_res = pin_out ? results + off : B.h_res;
        AIM_CUDA(cudaMemcpyAsync(o_res, B.d_res, (size_t)m * sizeof(aim_result), cudaMemcpyDeviceToHost, ctx->s_d2h));
        if (cigars) {
            AIM_CUDA(cudaMemcpyAsync(cigars + (size_t)off * (size_t)pitch, B.d_cig, (size_t)m * (size_t)pitch, cudaMemcpyDeviceToHost, ctx->s_d2h));
        } else if (sparse) {
            AIM_CUDA(cudaMemcpyAsync(B.h_runs, B.d_cig, (size_t)m * (size_t)runs_pitch, cudaMemcpyDeviceToHost, ctx->s_d2h));
        } else if (bt) {
            char *o_ops = pin_out ? ops + (size_t)off * 2 * rs : B.h_ops;
            AIM_CUDA(cudaMemcpyAsync(o_ops, B.d_ops, (size_t)m * 2 * rs, cudaMemcpyDeviceToHost, ctx->s_d2h));
        }
        AIM_CUDA(cudaEventRecord(B.ev[5], ctx->s_d2h));
        if (sparse) {
            { std::lock_guard<std::mutex> lk(co.m); co.submitted = j + 1; }
            co.cv.notify_all();
        }
    }
    const uint32_t pulled = (uint32_t)mine.size();
    if (sparse) {
        co.close();
        if (co.rc != AIM_OK) { set_error(co.err); return fail(co.rc); }
    } else {
        for (uint32_t j = (pulled >= (uint32_t)nbuf ? pulled - (uint32_t)nbuf : 0); j < pulled; ++j) {
            rc = finish(j);
            if (rc != AIM_OK) return fail(rc);
        }
    }
    if (phase_ms) for (int k = 0; k < 3; ++k) phase_ms[k] = ph[k];
    return AIM_OK;
}

// One GPU's share [first, first + n) of a PACKED host batch (aim_align_packed): per chunk, lengths + packed rows + flag
// words up, kernels, results + CIGAR rows down; same three-stream pipeline as run_shard.  The chunk's device buffers are
// reused: d_pat holds the packed rows, d_cig the flag words followed by the CIGAR rows (sized from cigar_pitch), d_ops the
// op rows the backtrace writes (they never leave the device).
int run_shard_packed(const aim_params &p, int device, uint32_t first, uint32_t n, uint32_t idx_base, const int32_t *plen,
                     const int32_t *tlen, const uint32_t *packed, const uint32_t *flags, aim_result *results, char *cigars,
                     int32_t pitch, double phase_ms[3], std::string *err)
{
    auto fail = [&](int rc) { if (err) *err = aim_last_error(); return rc; };
    DeviceCtx *ctx = nullptr;
    int rc = get_ctx(device, &ctx);
    if (rc != AIM_OK) return fail(rc);
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (cudaSetDevice(device) != cudaSuccess) { set_error("cudaSetDevice failed"); return fail(AIM_ERR_CUDA); }
    if (n == 0) return AIM_OK;
    const size_t rs = (size_t)p.read_size;
    const size_t row2 = 2 * (size_t)aim_packed_row_bytes(p.read_size);  // packed bytes per pair
    if (row2 > rs) { set_error("read_size too small for the packed entry"); return fail(AIM_ERR_ARG); }
    // chunks start on multiples of 32 pairs so that every chunk owns whole flag words
    const uint32_t lead = (32u - (first & 31u)) & 31u;  // pairs of a shard that starts inside a flag word: handled by an unaligned first chunk
    uint32_t chunk_pairs = (uint32_t)std::max<size_t>(16384, std::min<size_t>(((size_t)96 << 20) / (row2 + (size_t)pitch + 32), 1u << 20)) & ~31u;
    if (const char *cm = getenv("AIM_CHUNK_PAIRS")) { const long v = atol(cm); if (v >= 32) chunk_pairs = (uint32_t)v & ~31u; }
    if (!ctx->s_h2d) {
        AIM_CUDA(cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking));
        AIM_CUDA(cudaStreamCreateWithFlags(&ctx->s_kernel, cudaStreamNonBlocking));
        AIM_CUDA(cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking));
    }
    // chunk list: an unaligned head (if any), then aligned chunks
    std::vector<std::pair<uint32_t, uint32_t>> ch;  // (offset inside the shard, pairs)
    {
        uint32_t off = 0;
        if (lead && lead < n) { ch.emplace_back(0u, lead); off = lead; }
        while (off < n) { const uint32_t m = std::min(chunk_pairs, n - off); ch.emplace_back(off, m); off += m; }
    }
    const uint32_t nchunks = (uint32_t)ch.size();
    const int nbuf = (int)std::min<uint32_t>(kNumBuf, nchunks);
    const uint32_t cap = std::min(std::max(chunk_pairs, lead), std::max(n, 32u));
    // flag words (rounded to 256 B) + CIGAR rows live in the chunk's d_cig buffer, sized from the pitch
    const size_t flag_cap = ((size_t)(cap + 31) / 32 * 4 + 255) / 256 * 256;
    for (int b = 0; b < nbuf; ++b) {
        rc = ensure_chunk(ctx->chunk[b], cap, p.read_size, true, false);
        if (rc != AIM_OK) return fail(rc);
        ChunkBuf &B = ctx->chunk[b];
        const size_t need = flag_cap + (size_t)cap * (size_t)pitch;
        if (B.cig_cap < need) {
            cudaFree(B.d_cig);
            B.d_cig = nullptr;
            B.cig_cap = 0;
            AIM_CUDA(cudaMalloc(&B.d_cig, need));
            B.cig_cap = need;
        }
    }
    double ph[3] = {0, 0, 0};
    auto finish = [&](uint32_t c) -> int {
        ChunkBuf &B = ctx->chunk[c % (uint32_t)nbuf];
        cudaError_t e = cudaEventSynchronize(B.ev[5]);
        if (e != cudaSuccess) { set_error(std::string("chunk sync: ") + cudaGetErrorString(e)); return AIM_ERR_CUDA; }
        float t;
        for (int k = 0; k < 3; ++k) { cudaEventElapsedTime(&t, B.ev[2 * k], B.ev[2 * k + 1]); ph[k] += t; }
        return AIM_OK;
    };
    for (uint32_t c = 0; c < nchunks; ++c) {
        ChunkBuf &B = ctx->chunk[c % (uint32_t)nbuf];
        if (c >= (uint32_t)nbuf) { rc = finish(c - (uint32_t)nbuf); if (rc != AIM_OK) return fail(rc); }
        const uint32_t off = ch[c].first, m = ch[c].second;
        const uint32_t g0 = first + off;  // first pair of the chunk in the caller's arrays
        const int32_t *s_plen = plen + g0, *s_tlen = tlen + g0;
        {   // host.c:119-123 rejects reads longer than READ_SIZE
            int32_t lo = 0, hi = 0;
            for (uint32_t i = 0; i < m; ++i) {
                lo = std::min(lo, std::min(s_plen[i], s_tlen[i]));
                hi = std::max(hi, std::max(s_plen[i], s_tlen[i]));
            }
            if (lo < 0 || hi > p.read_size) {
                cudaDeviceSynchronize();
                set_error(lo < 0 ? "negative sequence length" : "READ LENGTH less than length of the input reads");
                return fail(lo < 0 ? AIM_ERR_ARG : AIM_ERR_LENGTH);
            }
        }
        // flag words of the chunk: the caller's words when the chunk starts on a word boundary, else shifted on the host
        const uint32_t fwords = (m + 31) / 32;
        uint32_t *d_flags = reinterpret_cast<uint32_t *>(B.d_cig);
        char *d_cig = B.d_cig + flag_cap;
        std::vector<uint32_t> shifted;
        const uint32_t *s_flags = flags + (g0 >> 5);
        if (g0 & 31u) {
            shifted.assign(fwords, 0u);
            for (uint32_t i = 0; i < m; ++i) {
                const uint32_t g = g0 + i;
                if ((flags[g >> 5] >> (g & 31)) & 1u) shifted[i >> 5] |= 1u << (i & 31);
            }
            s_flags = shifted.data();
        }
        AIM_CUDA(cudaEventRecord(B.ev[0], ctx->s_h2d));
        AIM_CUDA(cudaMemcpyAsync(B.d_plen, s_plen, (size_t)m * 4, cudaMemcpyHostToDevice, ctx->s_h2d));
        AIM_CUDA(cudaMemcpyAsync(B.d_tlen, s_tlen, (size_t)m * 4, cudaMemcpyHostToDevice, ctx->s_h2d));
        AIM_CUDA(cudaMemcpyAsync(B.d_pat, reinterpret_cast<const char *>(packed) + (size_t)g0 * row2, (size_t)m * row2, cudaMemcpyHostToDevice, ctx->s_h2d));
        AIM_CUDA(cudaMemcpyAsync(d_flags, s_flags, (size_t)fwords * 4, cudaMemcpyHostToDevice, ctx->s_h2d));
        if (!shifted.empty()) AIM_CUDA(cudaStreamSynchronize(ctx->s_h2d));  // the temporary must outlive the copy (unaligned head only)
        AIM_CUDA(cudaEventRecord(B.ev[1], ctx->s_h2d));

        AIM_CUDA(cudaStreamWaitEvent(ctx->s_kernel, B.ev[1], 0));
        AIM_CUDA(cudaEventRecord(B.ev[2], ctx->s_kernel));
        KernelArgs a{p, m, idx_base + g0, B.d_plen, B.d_tlen, nullptr, nullptr, B.d_res, B.d_ops};
        a.packed = reinterpret_cast<const uint32_t *>(B.d_pat);
        a.pflags = d_flags;
        a.cigars = d_cig;
        a.cigar_pitch = pitch;
        rc = launch(a, &ctx->scratch, ctx->s_kernel, nullptr);
        if (rc != AIM_OK) { cudaDeviceSynchronize(); return fail(rc); }
        AIM_CUDA(cudaEventRecord(B.ev[3], ctx->s_kernel));

        AIM_CUDA(cudaStreamWaitEvent(ctx->s_d2h, B.ev[3], 0));
        AIM_CUDA(cudaEventRecord(B.ev[4], ctx->s_d2h));
        AIM_CUDA(cudaMemcpyAsync(results + g0, B.d_res, (size_t)m * sizeof(aim_result), cudaMemcpyDeviceToHost, ctx->s_d2h));
        AIM_CUDA(cudaMemcpyAsync(cigars + (size_t)g0 * (size_t)pitch, d_cig, (size_t)m * (size_t)pitch, cudaMemcpyDeviceToHost, ctx->s_d2h));
        AIM_CUDA(cudaEventRecord(B.ev[5], ctx->s_d2h));
    }
    for (uint32_t c = (nchunks >= (uint32_t)nbuf ? nchunks - (uint32_t)nbuf : 0); c < nchunks; ++c) {
        rc = finish(c);
        if (rc != AIM_OK) return fail(rc);
    }
    if (phase_ms) for (int k = 0; k < 3; ++k) phase_ms[k] = ph[k];
    return AIM_OK;
}

}  // namespace

int scratch_reserve(Scratch *s, size_t bytes)
{
    if (s->bytes >= bytes) return AIM_OK;
    // the scratch may still be in use by an earlier launch on another stream
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess && s->buf) e = cudaFree(s->buf);
    s->buf = nullptr;
    s->bytes = 0;
    if (e == cudaSuccess) e = cudaMalloc(&s->buf, bytes);
    if (e != cudaSuccess) {
        set_error(std::string("scratch allocation of ") + std::to_string(bytes >> 20) + " MiB: " + cudaGetErrorString(e));
        cudaGetLastError();
        return e == cudaErrorMemoryAllocation ? AIM_ERR_NOMEM : AIM_ERR_CUDA;
    }
    s->bytes = bytes;
    return AIM_OK;
}

int device_state(int device, Scratch **scratch, void **mu)
{
    DeviceCtx *ctx = nullptr;
    int rc = get_ctx(device, &ctx);
    if (rc != AIM_OK) return rc;
    *scratch = &ctx->scratch;
    *mu = &ctx->mu;
    return AIM_OK;
}

int launch_algo(const KernelArgs &a, Scratch *s, void *stream, int *launches) { return launch(a, s, (cudaStream_t)stream, launches); }

bool params_valid_for_file(const aim_params *p) { return validate(p, false, nullptr) == AIM_OK; }
aim_params params_normalized(const aim_params *p) { return normalized(p); }

namespace {
struct PlanEntry { std::vector<unsigned char> host; void *dev; };
struct PlanCache { std::vector<PlanEntry> e; };
}  // namespace

const void *cached_plan(Scratch *s, const void *host, size_t bytes)
{
    if (!s->plan_cache) s->plan_cache = new PlanCache();
    PlanCache *pc = static_cast<PlanCache *>(s->plan_cache);
    for (const PlanEntry &e : pc->e)
        if (e.host.size() == bytes && memcmp(e.host.data(), host, bytes) == 0) return e.dev;
    if (pc->e.size() >= 32) {  // a long-lived process cycling through many penalty sets: start over (nothing may still read them)
        cudaDeviceSynchronize();
        for (PlanEntry &e : pc->e) cudaFree(e.dev);
        pc->e.clear();
    }
    PlanEntry ne;
    ne.host.assign(static_cast<const unsigned char *>(host), static_cast<const unsigned char *>(host) + bytes);
    ne.dev = nullptr;
    cudaError_t e = cudaMalloc(&ne.dev, std::max<size_t>(bytes, 256));
    if (e == cudaSuccess) e = cudaMemcpy(ne.dev, host, bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        set_error(std::string("plan upload: ") + cudaGetErrorString(e));
        cudaGetLastError();
        if (ne.dev) cudaFree(ne.dev);
        return nullptr;
    }
    pc->e.push_back(std::move(ne));
    return pc->e.back().dev;
}

int scratch_acquire(Scratch *s, void *stream)
{
    if (s->busy_valid && s->busy_stream != stream) {
        cudaError_t e = cudaStreamWaitEvent((cudaStream_t)stream, (cudaEvent_t)s->busy_event, 0);
        if (e != cudaSuccess) { set_error(std::string("scratch hand-over: ") + cudaGetErrorString(e)); return AIM_ERR_CUDA; }
    }
    return AIM_OK;
}

int scratch_release(Scratch *s, void *stream)
{
    cudaError_t e = cudaSuccess;
    if (!s->busy_event) {
        cudaEvent_t ev = nullptr;
        e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
        s->busy_event = ev;
    }
    if (e == cudaSuccess) e = cudaEventRecord((cudaEvent_t)s->busy_event, (cudaStream_t)stream);
    if (e != cudaSuccess) { set_error(std::string("scratch hand-over: ") + cudaGetErrorString(e)); return AIM_ERR_CUDA; }
    s->busy_stream = stream;
    s->busy_valid = true;
    return AIM_OK;
}

void scratch_destroy(Scratch *s)
{
    cudaFree(s->buf);
    s->buf = nullptr;
    s->bytes = 0;
    if (s->plan_cache) {
        PlanCache *pc = static_cast<PlanCache *>(s->plan_cache);
        for (PlanEntry &e : pc->e) cudaFree(e.dev);
        delete pc;
        s->plan_cache = nullptr;
    }
    if (s->busy_event) cudaEventDestroy((cudaEvent_t)s->busy_event);
    s->busy_event = nullptr;
    s->busy_valid = false;
    if (s->side_stream) cudaStreamDestroy((cudaStream_t)s->side_stream);
    s->side_stream = nullptr;
    for (auto &ev : s->side_ev) { if (ev) cudaEventDestroy((cudaEvent_t)ev); ev = nullptr; }
}

}  // namespace aim

using namespace aim;

extern "C" int aim_device_count(void)
{
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) { cudaGetLastError(); return 0; }
    return count;
}

extern "C" void *aim_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) {
        set_error("cudaHostAlloc failed");
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
extern "C" void aim_host_free(void *p) { if (p) cudaFreeHost(p); }

extern "C" void aim_shutdown(void)
{
    file_pipeline_shutdown();
    host_pool_shutdown();
    std::lock_guard<std::mutex> lk(g_mu);
    for (DeviceCtx *c : g_ctx) {
        if (!c) continue;
        cudaSetDevice(c->device);
        cudaDeviceSynchronize();
        for (auto &b : c->chunk) free_chunk(b);
        if (c->s_h2d) { cudaStreamDestroy(c->s_h2d); cudaStreamDestroy(c->s_kernel); cudaStreamDestroy(c->s_d2h); }
        if (c->s_fix) cudaStreamDestroy(c->s_fix);
        scratch_destroy(&c->scratch);
        delete c;
    }
    g_ctx.clear();
}

extern "C" int aim_align_device(const aim_params *params, int device, uint32_t n, uint32_t idx_base,
                                const int32_t *d_plen, const int32_t *d_tlen, const char *d_patterns,
                                const char *d_texts, aim_result *d_results, char *d_ops, void *stream,
                                float *kernel_ms, int32_t *launches)
{
    if (!params) { set_error("params is NULL"); return AIM_ERR_ARG; }
    const aim_params norm = normalized(params);
    params = &norm;
    int rc = validate(params, true, d_ops);
    if (rc != AIM_OK) return rc;
    if (!d_plen || !d_tlen || !d_patterns || !d_texts || !d_results) { set_error("NULL device buffer"); return AIM_ERR_ARG; }
    DeviceCtx *ctx = nullptr;
    rc = get_ctx(device, &ctx);
    if (rc != AIM_OK) return rc;
    std::lock_guard<std::mutex> lk(ctx->mu);
    AIM_CUDA(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (kernel_ms) {
        AIM_CUDA(cudaEventCreate(&e0));
        AIM_CUDA(cudaEventCreate(&e1));
        AIM_CUDA(cudaEventRecord(e0, st));
    }
    KernelArgs a{*params, n, idx_base, d_plen, d_tlen, d_patterns, d_texts, d_results, params->backtrace ? d_ops : nullptr};
    int nl = 0;
    rc = launch(a, &ctx->scratch, st, &nl);
    if (launches) *launches = nl;
    if (kernel_ms) {
        if (rc == AIM_OK) {
            AIM_CUDA(cudaEventRecord(e1, st));
            AIM_CUDA(cudaEventSynchronize(e1));
            AIM_CUDA(cudaEventElapsedTime(kernel_ms, e0, e1));
        }
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
    }
    return rc;
}

static int align_batch_once(const aim_params *params, uint32_t n, uint32_t idx_base, const int32_t *plen,
                            const int32_t *tlen, const char *patterns, const char *texts,
                            aim_result *results, char *ops, double phase_ms[3], char *cigars, int32_t pitch)
{
    if (!params) { set_error("params is NULL"); return AIM_ERR_ARG; }
    const aim_params norm = normalized(params);
    params = &norm;
    int rc = validate(params, cigars == nullptr, ops);
    if (rc != AIM_OK) return rc;
    if (n > 0 && (!plen || !tlen || !patterns || !texts || !results)) { set_error("NULL host buffer"); return AIM_ERR_ARG; }
    // (sequence lengths are validated chunk by chunk inside run_shard, overlapped with the GPU's work)
    if (phase_ms) phase_ms[0] = phase_ms[1] = phase_ms[2] = 0.0;
    int ndev = aim_device_count();
    if (ndev == 0) { set_error("no CUDA device (aim_b200 has no CPU fallback)"); return AIM_ERR_NO_DEVICE; }
    int g = params->ngpus <= 1 ? 1 : params->ngpus;
    if (params->device < 0 || params->device + g > ndev) { set_error("device range exceeds visible GPUs"); return AIM_ERR_ARG; }
    const size_t per_pair = 2 * (size_t)params->read_size + (cigars ? (size_t)pitch : (params->backtrace ? 2 * (size_t)params->read_size : 0)) + sizeof(aim_result) + 8;
    const uint32_t chunk_pairs = pick_chunk_pairs(*params, n, per_pair, g);
    std::atomic<uint32_t> queue{0};
    if (n == 0) return AIM_OK;
    if (g == 1) return run_shard(*params, params->device, &queue, chunk_pairs, n, idx_base, plen, tlen, patterns, texts, results, ops, phase_ms, nullptr, cigars, pitch);

    // one host thread + stream set per GPU, all pulling chunks from one queue (the reference splits contiguously per DPU,
    // host.c:201-209; a queue gives the same output - every chunk lands at its own offset - and balances uneven pairs)
    // the host threads that rebuild the op rows serve g GPUs now: 3/4 of the cores, less the g submitters (one process drives
    // the box's GPUs here, so the cores are not shared with other ranks)
    host_pool_want((int)std::thread::hardware_concurrency() * 3 / 4 - g);
    std::vector<std::thread> th;
    std::vector<int> rcs((size_t)g, AIM_OK);
    std::vector<std::string> errs((size_t)g);
    std::vector<double> ph((size_t)g * 3, 0.0);
    for (int d = 0; d < g; ++d) {
        th.emplace_back([&, d]() {
            rcs[(size_t)d] = run_shard(*params, params->device + d, &queue, chunk_pairs, n, idx_base, plen, tlen, patterns, texts,
                                       results, ops, &ph[(size_t)d * 3], &errs[(size_t)d], cigars, pitch);
            if (rcs[(size_t)d] != AIM_OK && errs[(size_t)d].empty()) errs[(size_t)d] = aim_last_error();
        });
    }
    for (auto &t : th) t.join();
    for (int d = 0; d < g; ++d) {
        if (rcs[(size_t)d] != AIM_OK) { set_error("gpu " + std::to_string(params->device + d) + ": " + errs[(size_t)d]); return rcs[(size_t)d]; }
        if (phase_ms) for (int k = 0; k < 3; ++k) phase_ms[k] = std::max(phase_ms[k], ph[(size_t)d * 3 + k]);
    }
    return AIM_OK;
}

// aim_align_batch + arena-overflow recovery.  Long-read WFA with backtrace keeps a pair's wavefront history in a fixed arena
// (the reference's per-tasklet MRAM segment; there a pair that outgrows it ends the run: "Out of memory MRAM",
// dpu_allocator_mram.c:6-10).  Here such pairs come back flagged AIM_STATUS_ARENA and are aligned again, alone, with an arena
// eight times larger, up to three times; what still does not fit keeps its flag.
static int align_batch_impl(const aim_params *params, uint32_t n, uint32_t idx_base, const int32_t *plen,
                            const int32_t *tlen, const char *patterns, const char *texts,
                            aim_result *results, char *ops, double phase_ms[3], char *cigars, int32_t pitch)
{
    int rc = align_batch_once(params, n, idx_base, plen, tlen, patterns, texts, results, ops, phase_ms, cigars, pitch);
    if (rc != AIM_OK || !params || params->algo != AIM_ALGO_WFA || !params->backtrace || params->read_size < 512 || n == 0) return rc;
    aim_params q = normalized(params);
    for (int round = 0; round < 3; ++round) {
        std::vector<uint32_t> bad;
        for (uint32_t i = 0; i < n; ++i) if (results[i].status == AIM_STATUS_ARENA) bad.push_back(i);
        if (bad.empty()) break;
        q.arena_mb = (q.arena_mb > 0 ? q.arena_mb : 8) * 8;
        if (q.arena_mb > 4096) break;
        const size_t rs = (size_t)q.read_size, m = bad.size();
        std::vector<int32_t> pl(m), tl(m);
        std::vector<char> pa(m * rs), tx(m * rs), op(cigars ? 0 : m * 2 * rs), cg(cigars ? m * (size_t)pitch : 0);
        std::vector<aim_result> rr(m);
        for (size_t k = 0; k < m; ++k) {
            pl[k] = plen[bad[k]]; tl[k] = tlen[bad[k]];
            memcpy(&pa[k * rs], patterns + (size_t)bad[k] * rs, rs);
            memcpy(&tx[k * rs], texts + (size_t)bad[k] * rs, rs);
        }
        double ph[3] = {0, 0, 0};
        rc = align_batch_once(&q, (uint32_t)m, 0, pl.data(), tl.data(), pa.data(), tx.data(), rr.data(), cigars ? nullptr : op.data(), ph,
                              cigars ? cg.data() : nullptr, pitch);
        if (rc != AIM_OK) return rc;
        for (size_t k = 0; k < m; ++k) {
            rr[k].idx = results[bad[k]].idx;
            results[bad[k]] = rr[k];
            if (cigars) memcpy(cigars + (size_t)bad[k] * (size_t)pitch, &cg[k * (size_t)pitch], (size_t)pitch);
            else memcpy(ops + (size_t)bad[k] * 2 * rs, &op[k * 2 * rs], 2 * rs);
        }
        if (phase_ms) for (int k = 0; k < 3; ++k) phase_ms[k] += ph[k];
    }
    return AIM_OK;
}

extern "C" int aim_align_batch(const aim_params *params, uint32_t n, uint32_t idx_base, const int32_t *plen,
                               const int32_t *tlen, const char *patterns, const char *texts,
                               aim_result *results, char *ops, double phase_ms[3])
{
    return align_batch_impl(params, n, idx_base, plen, tlen, patterns, texts, results, ops, phase_ms, nullptr, 0);
}

extern "C" int aim_align_batch_cigars(const aim_params *params, uint32_t n, uint32_t idx_base, const int32_t *plen,
                                      const int32_t *tlen, const char *patterns, const char *texts,
                                      aim_result *results, char *cigars, int32_t cigar_pitch, double phase_ms[3])
{
    if (!params) { set_error("params is NULL"); return AIM_ERR_ARG; }
    aim_params q = *params;
    if (q.algo != AIM_ALGO_GENASM_DC && q.algo != AIM_ALGO_GENASM_FILTER) q.backtrace = 1;  // the CIGAR is the point
    if (q.algo == AIM_ALGO_GENASM_FILTER) { set_error("GenASM-filter has no CIGAR"); return AIM_ERR_ARG; }
    if (!cigars || cigar_pitch < 16 || (cigar_pitch % 16) != 0 || cigar_pitch > 2 * q.read_size) {
        set_error("cigars required; cigar_pitch must be a multiple of 16 in 16..2*read_size");
        return AIM_ERR_ARG;
    }
    return align_batch_impl(&q, n, idx_base, plen, tlen, patterns, texts, results, nullptr, phase_ms, cigars, cigar_pitch);
}

extern "C" int aim_align_packed(const aim_params *params, uint32_t n, uint32_t idx_base, const int32_t *plen, const int32_t *tlen,
                                const uint32_t *packed, const uint32_t *flags, aim_result *results, char *cigars,
                                int32_t cigar_pitch, double phase_ms[3])
{
    int rc = validate(params, false, nullptr);
    if (rc != AIM_OK) return rc;
    if (params->algo != AIM_ALGO_WFA || !params->backtrace) { set_error("aim_align_packed serves WFA with backtrace"); return AIM_ERR_ARG; }
    if (cigar_pitch < 16 || (cigar_pitch % 16) != 0 || cigar_pitch > 2 * params->read_size || params->read_size < 64) {
        set_error("cigar_pitch must be a multiple of 16 in 16..2*read_size (read_size >= 64)");
        return AIM_ERR_ARG;
    }
    if (n > 0 && (!plen || !tlen || !packed || !flags || !results || !cigars)) { set_error("NULL host buffer"); return AIM_ERR_ARG; }
    if (phase_ms) phase_ms[0] = phase_ms[1] = phase_ms[2] = 0.0;
    int ndev = aim_device_count();
    if (ndev == 0) { set_error("no CUDA device (aim_b200 has no CPU fallback)"); return AIM_ERR_NO_DEVICE; }
    const int g = params->ngpus <= 1 ? 1 : params->ngpus;
    if (params->device < 0 || params->device + g > ndev) { set_error("device range exceeds visible GPUs"); return AIM_ERR_ARG; }
    if (g == 1) return run_shard_packed(*params, params->device, 0, n, idx_base, plen, tlen, packed, flags, results, cigars, cigar_pitch, phase_ms, nullptr);
    // contiguous index ranges per GPU on multiples of 32 pairs (whole flag words), one host thread + stream set each
    const uint32_t per = ((n + (uint32_t)g - 1) / (uint32_t)g + 31u) & ~31u;
    std::vector<std::thread> th;
    std::vector<int> rcs((size_t)g, AIM_OK);
    std::vector<std::string> errs((size_t)g);
    std::vector<double> ph((size_t)g * 3, 0.0);
    for (int d = 0; d < g; ++d) {
        const uint32_t first = std::min<uint64_t>(n, (uint64_t)d * per), cnt = std::min(per, n - first);
        th.emplace_back([&, d, first, cnt]() {
            rcs[(size_t)d] = run_shard_packed(*params, params->device + d, first, cnt, idx_base, plen, tlen, packed, flags, results,
                                              cigars, cigar_pitch, &ph[(size_t)d * 3], &errs[(size_t)d]);
            if (rcs[(size_t)d] != AIM_OK && errs[(size_t)d].empty()) errs[(size_t)d] = aim_last_error();
        });
    }
    for (auto &t : th) t.join();
    for (int d = 0; d < g; ++d) {
        if (rcs[(size_t)d] != AIM_OK) { set_error("gpu " + std::to_string(params->device + d) + ": " + errs[(size_t)d]); return rcs[(size_t)d]; }
        if (phase_ms) for (int k = 0; k < 3; ++k) phase_ms[k] = std::max(phase_ms[k], ph[(size_t)d * 3 + k]);
    }
    return AIM_OK;
}
