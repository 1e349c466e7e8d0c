// Host-side logic of the drop-in boundary that needs no GPU: knob derivation, the pairs-file
// reader, the result writer, the pairs-to-process rule and the synthetic-pair generator.
// Mirrors the reference's host program (citations: paths under the reference checkout).
#include "aim_b200.h"
#include "aim_internal.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace aim {
thread_local std::string g_last_error;
void set_error(const std::string &msg) { g_last_error = msg; }

// Number of '\n' bytes in [p, p + n): 32 bytes per step with AVX2 where the CPU has it (the file pipeline counts the
// newlines of every chunk it reads, at memory speed), 8 bytes per step otherwise.
#if defined(__x86_64__)
__attribute__((target("avx2"))) static size_t count_newlines_avx2(const char *p, size_t n)
{
    const __m256i nl = _mm256_set1_epi8('\n');
    size_t c = 0, i = 0;
    for (; i + 32 <= n; i += 32)
        c += (size_t)__builtin_popcount((unsigned)_mm256_movemask_epi8(_mm256_cmpeq_epi8(_mm256_loadu_si256((const __m256i *)(p + i)), nl)));
    for (; i < n; ++i) c += p[i] == '\n';
    return c;
}
#endif
size_t count_newlines(const char *p, size_t n)
{
#if defined(__x86_64__)
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2) return count_newlines_avx2(p, n);
#endif
    size_t c = 0, i = 0;
    for (; i + 8 <= n; i += 8) {
        uint64_t w;
        memcpy(&w, p + i, 8);
        w ^= 0x0a0a0a0a0a0a0a0aull;  // zero bytes where '\n'
        const uint64_t nz = ((w & 0x7f7f7f7f7f7f7f7full) + 0x7f7f7f7f7f7f7f7full) | w;  // high bit of every NON-zero byte
        c += (size_t)__builtin_popcountll(~nz & 0x8080808080808080ull);
    }
    for (; i < n; ++i) c += p[i] == '\n';
    return c;
}

// ---- host pool: a few long-lived threads for the byte work that has to happen on the host beside the DMA ----
namespace {
class HostPool {
  public:
    ~HostPool() { shutdown(); }
    // fn(b) for every b in [0, nblocks), on the pool's threads and the caller's; returns when all are done
    void run(uint32_t nblocks, const std::function<void(uint32_t)> &fn)
    {
        std::lock_guard<std::mutex> call(call_mu_);  // one job at a time (several GPUs' coordinators share the pool)
        if (nblocks <= 1 || threads() == 0) { for (uint32_t b = 0; b < nblocks; ++b) fn(b); return; }
        {
            std::lock_guard<std::mutex> lk(m_);
            start_threads();
            if (th_.empty()) { for (uint32_t b = 0; b < nblocks; ++b) fn(b); return; }
            fn_ = &fn;
            nblocks_ = nblocks;
            next_.store(0);
            active_ = (int)th_.size();
            ++gen_;
        }
        cv_work_.notify_all();
        for (uint32_t b; (b = next_.fetch_add(1)) < nblocks;) fn(b);
        std::unique_lock<std::mutex> lk(m_);
        cv_done_.wait(lk, [&] { return active_ == 0; });
        fn_ = nullptr;
    }
    void shutdown()
    {
        std::lock_guard<std::mutex> call(call_mu_);
        { std::lock_guard<std::mutex> lk(m_); stop_ = true; }
        cv_work_.notify_all();
        for (auto &t : th_) t.join();
        th_.clear();
        stop_ = false;
    }
    // helper threads beside the caller's: AIM_HOST_THREADS - 1, else this process's share of the cores (torchrun:
    // LOCAL_WORLD_SIZE ranks) less one for the submitting thread, at most 8 threads in all: more fight the DMA-feeding and
    // CUDA threads for the cores and add nothing (measured)
    // more helpers for a caller that drives several GPUs from one process (never fewer)
    void want(int helpers)
    {
        std::lock_guard<std::mutex> call(call_mu_);
        std::lock_guard<std::mutex> lk(m_);
        if (getenv("AIM_HOST_THREADS")) return;
        extra_ = std::max(extra_, std::min(helpers, 63) - threads());
        if (!th_.empty()) grow();
    }
    static int threads()
    {
        static const int n = [] {
            if (const char *e = getenv("AIM_HOST_THREADS")) return std::max(0, std::min(atoi(e), 64) - 1);
            int hw = (int)std::thread::hardware_concurrency(), ranks = 1;
            if (const char *e = getenv("LOCAL_WORLD_SIZE")) ranks = std::max(1, atoi(e));
            return std::max(1, std::min(hw / ranks - 1, 8) - 1);
        }();
        return n;
    }

  private:
    void start_threads() { grow(); }
    void grow()  // (m_ held; a thread that starts now has seen no generation yet and joins the next job)
    {
        const int n = threads() + std::max(0, extra_);
        try {
            while ((int)th_.size() < n) { const uint64_t g = gen_; th_.emplace_back([this, g] { worker(g); }); }
        } catch (...) {  // no more threads to be had: the ones there are (or the caller alone) do the work
        }
    }
    void worker(uint64_t seen)
    {
        for (;;) {
            const std::function<void(uint32_t)> *fn;
            uint32_t nb;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_work_.wait(lk, [&] { return stop_ || gen_ != seen; });
                if (stop_) return;
                seen = gen_;
                fn = fn_;
                nb = nblocks_;
            }
            for (uint32_t b; (b = next_.fetch_add(1)) < nb;) (*fn)(b);
            std::lock_guard<std::mutex> lk(m_);
            if (--active_ == 0) cv_done_.notify_all();
        }
    }
    std::mutex call_mu_, m_;
    std::condition_variable cv_work_, cv_done_;
    std::vector<std::thread> th_;
    const std::function<void(uint32_t)> *fn_ = nullptr;
    uint32_t nblocks_ = 0;
    std::atomic<uint32_t> next_{0};
    uint64_t gen_ = 0;
    int active_ = 0, extra_ = 0;
    bool stop_ = false;
};
HostPool g_pool;

// dst <- src with stores that bypass the cache (the rows are written once, at memory speed, and read by nobody on this
// core).  dst and src are 16-byte aligned and congruent modulo 32, n is a multiple of 16.
#if defined(__x86_64__)
__attribute__((target("avx2"))) void stream_block_avx2(char *dst, const char *src, size_t n)
{
    size_t k = 0;
    if (((uintptr_t)dst & 31u) && n) { _mm_stream_si128((__m128i *)dst, _mm_load_si128((const __m128i *)src)); k = 16; }
    for (; k + 32 <= n; k += 32) _mm256_stream_si256((__m256i *)(dst + k), _mm256_load_si256((const __m256i *)(src + k)));
    if (k < n) _mm_stream_si128((__m128i *)(dst + k), _mm_load_si128((const __m128i *)(src + k)));
}
void stream_block_sse2(char *dst, const char *src, size_t n)
{
    for (size_t k = 0; k < n; k += 16) _mm_stream_si128((__m128i *)(dst + k), _mm_load_si128((const __m128i *)(src + k)));
}
#endif
}  // namespace

// Inverse of op_runs_kernel (aim_file.cu): every pair's 2 * read_size op row from its run row (run count, then position |
// length << 16 | op << 24 per run of bytes other than 'M').  Sixteen rows at a time are put together in a cache-resident
// buffer - 'M' fill, then the few runs as byte stores - and streamed out.
void expand_op_runs(const unsigned char *runs, int pitch, uint32_t m, int read_size, char *ops, std::vector<uint32_t> *overflow)
{
    const size_t row = 2 * (size_t)read_size;
    const uint32_t cap = (uint32_t)(pitch >> 2) - 1u;
    constexpr uint32_t BLK = 2048, GRP = 16;
    std::mutex ov_mu;
#if defined(__x86_64__)
    static const bool avx2 = __builtin_cpu_supports("avx2");
    const bool aligned16 = ((uintptr_t)ops & 15u) == 0;  // (row is a multiple of 16)
#endif
    g_pool.run((m + BLK - 1) / BLK, [&](uint32_t b) {
        alignas(64) char store[GRP * 2048 + 64];
        std::vector<uint32_t> ov;
        const uint32_t i1 = std::min(m, (b + 1) * BLK);
        for (uint32_t g0 = b * BLK; g0 < i1; g0 += GRP) {
            const uint32_t g1 = std::min(i1, g0 + GRP);
            char *dst0 = ops + (size_t)g0 * row;
            char *buf = store + ((uintptr_t)dst0 & 31u);  // the group's rows, laid out as at dst0 modulo 32
            memset(buf, 'M', (size_t)(g1 - g0) * row);
            uint32_t bad = 0;  // bit per row of the group that is not rebuilt
            for (uint32_t i = g0; i < g1; ++i) {
                const uint32_t *rp = reinterpret_cast<const uint32_t *>(runs + (size_t)i * (size_t)pitch);
                char *rb = buf + (size_t)(i - g0) * row;
                const uint32_t c = rp[0];
                bool ok = c <= cap;  // 0xffffffff = more runs than the run row holds (op_runs_kernel)
                for (uint32_t j = 0; ok && j < c; ++j) {
                    const uint32_t e = rp[1 + j];
                    const size_t pos = e & 0xffffu, len = (e >> 16) & 0xffu;
                    const char op = (char)(e >> 24);
                    if (pos + len > row) { ok = false; break; }  // (cannot happen with rows op_runs_kernel wrote)
                    rb[pos] = op;
                    for (size_t q = 1; q < len; ++q) rb[pos + q] = op;
                }
                if (!ok) { ov.push_back(i); bad |= 1u << (i - g0); }
            }
            if (bad == 0) {
#if defined(__x86_64__)
                if (aligned16) {
                    if (avx2) stream_block_avx2(dst0, buf, (size_t)(g1 - g0) * row);
                    else stream_block_sse2(dst0, buf, (size_t)(g1 - g0) * row);
                    continue;
                }
#endif
                memcpy(dst0, buf, (size_t)(g1 - g0) * row);
            } else {  // rows that are not rebuilt stay untouched
                for (uint32_t i = g0; i < g1; ++i)
                    if (!((bad >> (i - g0)) & 1u)) memcpy(ops + (size_t)i * row, buf + (size_t)(i - g0) * row, row);
            }
        }
#if defined(__x86_64__)
        _mm_sfence();
#endif
        if (!ov.empty()) { std::lock_guard<std::mutex> lk(ov_mu); overflow->insert(overflow->end(), ov.begin(), ov.end()); }
    });
}

// What crosses PCIe per pair for the op row of this parameter set (0 = the row itself).  WFA: an alignment of score s has at most
// s / min(x, o + e) runs of ops other than 'M' (each costs at least a mismatch or a gap of one), and MAX_SCORE bounds s, so the
// run row needs no more words than that (config 4: 11 words = 44 bytes instead of 64); a row with more runs is fetched as it is anyway.
int32_t op_rows_download_bytes(const aim_params &p)
{
    if (!p.backtrace && p.algo != AIM_ALGO_GENASM_DC) return 0;
    if (p.algo == AIM_ALGO_GENASM_DC) return str_rows_pitch(p.read_size, p.max_score);
    if (p.algo != AIM_ALGO_NW && p.algo != AIM_ALGO_SWG && p.algo != AIM_ALGO_WFA) return 0;
    int32_t pitch = op_runs_pitch(p.read_size);
    if (pitch > 0 && p.algo == AIM_ALGO_WFA && p.max_score >= 0) {
        const int32_t unit = std::max(1, std::min(p.mismatch, p.gap_open + p.gap_ext));
        const int64_t by_score = 4 * (1 + (int64_t)p.max_score / unit);
        if (by_score >= 16 && by_score < pitch) pitch = (int32_t)by_score;
    }
    return pitch;
}

// GenASM-DC strings: k error levels make at most 2k + 1 runs of at most 5 characters ("1000M"); pieces of up to half a row
int32_t str_rows_pitch(int32_t read_size, int32_t max_score)
{
    const int64_t need = (2 * (int64_t)std::max(max_score, 0) + 1) * 5 + 1;
    const int32_t pitch = (int32_t)std::min<int64_t>((need + 15) / 16 * 16, 1 << 20);
    return (read_size >= 32 && pitch <= read_size) ? pitch : 0;
}

void expand_str_rows(const unsigned char *rows, int pitch, uint32_t m, int read_size, char *ops, std::vector<uint32_t> *overflow)
{
    const size_t row = 2 * (size_t)read_size;
    constexpr uint32_t BLK = 4096;
    std::mutex ov_mu;
    g_pool.run((m + BLK - 1) / BLK, [&](uint32_t b) {
        std::vector<uint32_t> ov;
        const uint32_t i1 = std::min(m, (b + 1) * BLK);
        for (uint32_t i = b * BLK; i < i1; ++i) {
            const unsigned char *r = rows + (size_t)i * (size_t)pitch;
            const size_t len = strnlen(reinterpret_cast<const char *>(r), (size_t)pitch);
            if (len == (size_t)pitch) { ov.push_back(i); continue; }
            memcpy(ops + (size_t)i * row, r, len + 1);
        }
        if (!ov.empty()) { std::lock_guard<std::mutex> lk(ov_mu); overflow->insert(overflow->end(), ov.begin(), ov.end()); }
    });
}

void host_pool_shutdown() { g_pool.shutdown(); }
void host_pool_want(int helpers) { g_pool.want(helpers); }

// bytes of a run row for READ_SIZE-wide pairs: room for ~3/32 * READ_SIZE runs of ops other than 'M' (a 4 % error rate makes
// ~0.04 per base); 0 = rows of this size are downloaded as they are
int32_t op_runs_pitch(int32_t read_size)
{
    if (read_size < 32 || read_size > 1024) return 0;
    return std::max(32, (read_size * 3 / 8 + 15) / 16 * 16);
}
}  // namespace aim

extern "C" int32_t aim_op_runs_pitch(int32_t read_size) { return aim::op_runs_pitch(read_size); }
extern "C" int32_t aim_str_rows_pitch(int32_t read_size, int32_t max_score) { return aim::str_rows_pitch(read_size, max_score); }
extern "C" int32_t aim_op_rows_download_bytes(const aim_params *params) { return params ? aim::op_rows_download_bytes(*params) : 0; }

extern "C" int aim_expand_op_runs(const unsigned char *runs, int32_t pitch, uint32_t n, int32_t read_size, char *ops,
                                  uint32_t *overflow, uint32_t overflow_cap, uint32_t *overflow_count)
{
    if (overflow_count) *overflow_count = 0;
    if (n == 0) return AIM_OK;
    if (!runs || !ops || pitch < 8 || (pitch % 4) != 0 || ((uintptr_t)runs & 3u) != 0 || read_size <= 0 || (read_size % 8) != 0 || read_size > 1024) {
        aim::set_error("aim_expand_op_runs: runs (4-byte aligned) and ops required, pitch a multiple of 4, read_size a multiple of 8 in 8..1024");
        return AIM_ERR_ARG;
    }
    std::vector<uint32_t> ov;
    aim::expand_op_runs(runs, pitch, n, read_size, ops, &ov);
    std::sort(ov.begin(), ov.end());
    if (overflow_count) *overflow_count = (uint32_t)ov.size();
    if (overflow) for (size_t k = 0; k < ov.size() && k < overflow_cap; ++k) overflow[k] = ov[k];
    return AIM_OK;
}

extern "C" const char *aim_last_error(void) { return aim::g_last_error.c_str(); }
extern "C" int aim_abi_version(void) { return AIM_B200_ABI_VERSION; }

extern "C" const char *aim_strerror(int code)
{
    switch (code) {
    case AIM_OK: return "ok";
    case AIM_ERR_ARG: return "invalid argument";
    case AIM_ERR_LENGTH: return "READ LENGTH less than length of the input reads";
    case AIM_ERR_CUDA: return "CUDA failure";
    case AIM_ERR_NO_DEVICE: return "no CUDA device (there is no CPU fallback)";
    case AIM_ERR_IO: return "I/O failure";
    case AIM_ERR_NOMEM: return "out of memory";
    default: return "unknown error";
    }
}

// WFA/DPU-MRAM/run-wfa-pim-mram.py:58-67, NW/DPU-MRAM/run-nw-pim-mram.py:51-60.  Python float
// arithmetic is IEEE double in the same operation order; math.ceil == ceil.
extern "C" int aim_derive_knobs(int32_t algo, int32_t read_length, double error, int32_t mismatch,
                                int32_t gap_open, int32_t gap_ext, int32_t *max_score, int32_t *read_size)
{
    if (read_length <= 0 || !max_score || !read_size) return AIM_ERR_ARG;
    double w = (double)read_length * error;
    double a = w * (double)mismatch;
    double b = (algo == AIM_ALGO_NW) ? w * (double)gap_open : w * (double)(gap_open + gap_ext);
    *max_score = (int32_t)std::ceil(a > b ? a : b);  // Python max(a, b): returns b only if b > a
    *read_size = (int32_t)std::ceil((((double)read_length + w) + 7.0) / 8.0) * 8;
    return AIM_OK;
}

// host.c:191 (ROUND_UP_MULTIPLE_8(N / nr_dpus)) and :201-209 (each DPU reads up to that many).
extern "C" uint32_t aim_pairs_to_process(uint32_t pairs_in_file, uint32_t n_arg, uint32_t nr_dpus)
{
    if (nr_dpus == 0) nr_dpus = 1;
    uint64_t per = ((uint64_t)(n_arg / nr_dpus) + 7) / 8 * 8;
    uint64_t cap = per * nr_dpus;
    return (uint32_t)(cap < pairs_in_file ? cap : pairs_in_file);
}

namespace {

// Read-only mapping of a whole file (empty files map to nothing).
struct Mapped {
    const char *p = nullptr;
    size_t n = 0;
    int fd = -1;
    bool ok = false;
    explicit Mapped(const char *path)
    {
        fd = open(path, O_RDONLY);
        if (fd < 0) return;
        struct stat st;
        if (fstat(fd, &st) != 0) return;
        n = (size_t)st.st_size;
        if (n) {
            void *m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
            if (m == MAP_FAILED) return;
            madvise(m, n, MADV_SEQUENTIAL);
            p = (const char *)m;
        }
        ok = true;
    }
    ~Mapped()
    {
        if (p) munmap((void *)p, n);
        if (fd >= 0) close(fd);
    }
};

int io_threads(size_t work_bytes)
{
    int t = (int)std::thread::hardware_concurrency();
    if (const char *e = getenv("AIM_IO_THREADS")) { int v = atoi(e); if (v >= 1) return std::min(v, 64); }  // exact, for tests
    t = std::max(1, std::min(t, 64));
    const size_t by_size = work_bytes / (4u << 20) + 1;  // at least ~4 MB per thread
    return (int)std::min<size_t>((size_t)t, by_size);
}

size_t count_nl(const char *p, size_t n)
{
    size_t c = 0;
    const char *e = p + n;
    while (p < e) {
        const char *q = (const char *)memchr(p, '\n', (size_t)(e - p));
        if (!q) break;
        ++c;
        p = q + 1;
    }
    return c;
}

template <typename F>
void parallel_for(int nthreads, F &&fn)
{
    if (nthreads <= 1) { fn(0); return; }
    std::vector<std::thread> th;
    th.reserve((size_t)nthreads);
    for (int t = 0; t < nthreads; ++t) th.emplace_back([&fn, t]() { fn(t); });
    for (auto &x : th) x.join();
}

// Lines of a mapped file as getline(3) sees them: every '\n' ends a line, a final unterminated run is a line too.
struct LineIndex {
    int nthreads = 1;
    std::vector<size_t> begin;  // byte range of every slice: [begin[t], begin[t+1])
    std::vector<size_t> nl_before;  // newlines in [0, begin[t])
    size_t lines = 0;
};

LineIndex index_lines(const Mapped &m)
{
    LineIndex ix;
    ix.nthreads = io_threads(m.n);
    const int T = ix.nthreads;
    ix.begin.resize((size_t)T + 1);
    for (int t = 0; t <= T; ++t) ix.begin[(size_t)t] = m.n / (size_t)T * (size_t)t;
    ix.begin[(size_t)T] = m.n;
    std::vector<size_t> nl((size_t)T, 0);
    parallel_for(T, [&](int t) { nl[(size_t)t] = count_nl(m.p + ix.begin[(size_t)t], ix.begin[(size_t)t + 1] - ix.begin[(size_t)t]); });
    ix.nl_before.assign((size_t)T + 1, 0);
    for (int t = 0; t < T; ++t) ix.nl_before[(size_t)t + 1] = ix.nl_before[(size_t)t] + nl[(size_t)t];
    ix.lines = ix.nl_before[(size_t)T] + ((m.n && m.p[m.n - 1] != '\n') ? 1 : 0);
    return ix;
}

}  // namespace

extern "C" int64_t aim_count_pairs(const char *path)
{
    Mapped m(path);
    if (!m.ok) { aim::set_error(std::string("Input file '") + path + "' couldn't be opened"); return AIM_ERR_IO; }
    const LineIndex ix = index_lines(m);
    return (int64_t)(ix.lines / 2);
}

// host.c:91-134 get_reads.  pattern = line+1, length = line_length-2 (the first character and the
// last one, assumed '\n', are dropped without being looked at).  The file is mapped and cut into
// one slice per host thread; a newline count per slice gives every slice the number of its first
// line, so all slices are copied into the n x read_size rows concurrently.
extern "C" int64_t aim_read_pairs(const char *path, uint32_t max_pairs, int32_t read_size,
                                  int32_t *plen, int32_t *tlen, char *patterns, char *texts)
{
    Mapped m(path);
    if (!m.ok) { aim::set_error(std::string("Input file '") + path + "' couldn't be opened"); return AIM_ERR_IO; }
    const LineIndex ix = index_lines(m);
    const size_t want = std::min<size_t>(ix.lines / 2, max_pairs);
    std::vector<int64_t> bad((size_t)ix.nthreads, -1);  // first pair with a read longer than read_size, per slice
    parallel_for(ix.nthreads, [&](int t) {
        size_t pos = ix.begin[(size_t)t];
        const size_t end = ix.begin[(size_t)t + 1];
        size_t line = ix.nl_before[(size_t)t];  // number of the line starting at pos, if one starts there
        if (pos > 0 && m.p[pos - 1] != '\n') {  // skip the tail of a line that started in an earlier slice
            const char *q = (const char *)memchr(m.p + pos, '\n', m.n - pos);
            if (!q) return;
            pos = (size_t)(q - m.p) + 1;
            ++line;
        }
        for (; pos < end && line < 2 * want; ++line) {  // lines that START in this slice
            const char *q = (const char *)memchr(m.p + pos, '\n', m.n - pos);
            const size_t len_with_nl = q ? (size_t)(q - (m.p + pos)) + 1 : m.n - pos;  // what getline returns
            long sl = (long)len_with_nl - 2;
            const size_t pair = line >> 1;
            if (sl > read_size) {
                if (bad[(size_t)t] < 0) bad[(size_t)t] = (int64_t)pair;
            } else {
                if (sl < 0) sl = 0;  // a 1-character line: the reference would index pattern[-1]; we clamp
                char *dst = ((line & 1) ? texts : patterns) + pair * (size_t)read_size;
                memcpy(dst, m.p + pos + 1, (size_t)sl);
                if (sl < read_size) memset(dst + sl, 0, (size_t)(read_size - sl));
                ((line & 1) ? tlen : plen)[pair] = (int32_t)sl;
            }
            pos += len_with_nl;
        }
    });
    for (int t = 0; t < ix.nthreads; ++t) {
        if (bad[(size_t)t] >= 0) {
            aim::set_error("READ LENGTH less than length of the input reads");
            return AIM_ERR_LENGTH;
        }
    }
    return (int64_t)want;
}

namespace {
// decimal digits of v at out (no NUL); returns the number written
inline size_t put_uint(char *out, uint32_t v)
{
    char tmp[10];
    size_t k = 0;
    do { tmp[k++] = (char)('0' + v % 10); v /= 10; } while (v);
    for (size_t i = 0; i < k; ++i) out[i] = tmp[k - 1 - i];
    return k;
}
inline size_t put_int(char *out, int32_t v)
{
    if (v < 0) { *out = '-'; return 1 + put_uint(out + 1, (uint32_t)(-(int64_t)v)); }
    return put_uint(out, (uint32_t)v);
}
// length of the run of bytes equal to ops[0], at most n (> 0)
inline size_t run_length(const char *ops, size_t n)
{
    const unsigned char c = (unsigned char)ops[0];
    const uint64_t pat = 0x0101010101010101ull * c;
    size_t i = 1;
    while (i + 8 <= n) {
        uint64_t w;
        memcpy(&w, ops + i, 8);
        const uint64_t d = w ^ pat;
        if (d) return i + (size_t)(__builtin_ctzll(d) >> 3);
        i += 8;
    }
    while (i < n && (unsigned char)ops[i] == c) ++i;
    return i;
}
// RLE of ops[b..e) (b < e) at out; the caller guarantees room for 11 bytes per op
inline size_t put_cigar(char *out, const char *ops, int32_t b, int32_t e)
{
    size_t pos = 0;
    ptrdiff_t i = b;  // (an empty alignment has b = -1: the reference prints the byte before its span, host.c:73)
    const ptrdiff_t end = e;
    while (i < end) {
        const size_t r = run_length(ops + i, (size_t)(end - i));
        pos += put_uint(out + pos, (uint32_t)r);
        out[pos++] = ops[i];
        i += (ptrdiff_t)r;
    }
    return pos;
}
}  // namespace

// host.c:69-89 edit_cigar_print.
extern "C" int aim_cigar_rle(const char *ops, int32_t begin_offset, int32_t end_offset, char *out, size_t cap)
{
    // a span that starts before the row (an empty pair has begin_offset = -1; the reference then prints the byte before its
    // buffer, an 'M' of the DPU's memset at best): "1M", as the device-side printer (cigar_rle_kernel) writes it
    if (begin_offset < 0) {
        if (cap < 2) return -1;
        out[0] = '1'; out[1] = 'M';
        return 2;
    }
    // edit_cigar_print always emits the op at begin_offset, even for an empty span
    const int32_t e = end_offset > begin_offset ? end_offset : begin_offset + 1;
    if ((size_t)(e - begin_offset) * 11 <= cap) return (int)put_cigar(out, ops, begin_offset, e);
    std::vector<char> tmp((size_t)(e - begin_offset) * 11);
    const size_t len = put_cigar(tmp.data(), ops, begin_offset, e);
    if (len > cap) return -1;
    memcpy(out, tmp.data(), len);
    return (int)len;
}

// host.c:332-353: "%d, %d, \n" then (BACKTRACE only) the RLE CIGAR on its own line.  Blocks of pairs
// are formatted by all host threads into private buffers and written in pair order.
extern "C" int aim_write_results(const char *path, uint32_t n, int32_t read_size, int32_t backtrace,
                                 const aim_result *results, const char *ops)
{
    FILE *f = fopen(path, "w");
    if (!f) { aim::set_error(std::string("Output file '") + path + "' couldn't be opened"); return AIM_ERR_IO; }
    const size_t rs2 = (size_t)read_size * 2;
    const uint32_t block = 16384;  // pairs per formatting block
    const int T = io_threads((size_t)n * (backtrace ? rs2 : 64));
    const size_t per_pair_cap = 40 + (backtrace ? rs2 * 11 + 1 : 0);
    std::vector<std::vector<char>> buf((size_t)T);
    std::vector<size_t> len((size_t)T, 0);
    int rc = AIM_OK;
    for (uint64_t base = 0; base < n && rc == AIM_OK; base += (uint64_t)block * (uint64_t)T) {
        parallel_for(T, [&](int t) {
            const uint64_t lo = base + (uint64_t)t * block, hi = std::min<uint64_t>(n, lo + block);
            len[(size_t)t] = 0;
            if (lo >= hi) return;
            std::vector<char> &b = buf[(size_t)t];
            size_t pos = 0;
            for (uint64_t i = lo; i < hi; ++i) {
                if (b.size() < pos + per_pair_cap) b.resize(std::max(b.size() * 2, pos + per_pair_cap));
                char *o = b.data();
                pos += put_int(o + pos, (int32_t)results[i].idx);
                o[pos++] = ','; o[pos++] = ' ';
                pos += put_int(o + pos, results[i].score);
                o[pos++] = ','; o[pos++] = ' '; o[pos++] = '\n';
                if (backtrace) {
                    const int32_t bo = results[i].begin_offset;
                    const int32_t eo = std::min<int64_t>(results[i].end_offset > bo ? results[i].end_offset : (int64_t)bo + 1, (int64_t)rs2);
                    if (bo < 0 || (size_t)bo >= rs2) { o[pos++] = '1'; o[pos++] = 'M'; }  // span outside the row: see aim_cigar_rle
                    else pos += put_cigar(o + pos, ops + i * rs2, bo, eo);
                    o[pos++] = '\n';
                }
            }
            len[(size_t)t] = pos;
        });
        for (int t = 0; t < T; ++t) {
            if (len[(size_t)t] && fwrite(buf[(size_t)t].data(), 1, len[(size_t)t], f) != len[(size_t)t]) { rc = AIM_ERR_IO; break; }
        }
    }
    if (ferror(f)) rc = AIM_ERR_IO;
    fclose(f);
    return rc;
}

// ---- compact transfers (include/aim_b200.h): host-side 2-bit packing and the printer for CIGAR rows ----
namespace {
// c = eight 2-bit codes, one in the low bits of each byte (byte j = base j) -> 16 bits, base 0 in the two most significant bits
uint32_t gather_swar(uint64_t c)
{
    uint64_t g = c | (c >> 6);  // two bases in the low nibble of every even byte (base j low, j + 1 high)
    g = (g & 0x000f000f000f000full) | ((g >> 12) & 0x00f000f000f000f0ull);  // four bases in the low byte of every 16-bit lane
    const uint32_t lo = (uint32_t)g, hi = (uint32_t)(g >> 32);
    const uint32_t b8 = (lo & 0xffu) | ((hi & 0xffu) << 8);  // base j at bits [2j + 1, 2j]
    uint32_t r = ((b8 & 0x3333u) << 2) | ((b8 >> 2) & 0x3333u);  // reverse the order of the eight 2-bit fields
    r = ((r & 0x0f0fu) << 4) | ((r >> 4) & 0x0f0fu);
    r = ((r & 0x00ffu) << 8) | ((r >> 8) & 0x00ffu);
    return r;
}
// one sequence row -> its packed words; returns false if a byte below len is not one of A, C, G, T.  The gather is a template
// argument so that the BMI2 variant inlines pext
template <typename Gather>
inline __attribute__((always_inline)) bool pack_row(const unsigned char *row, int32_t len, uint32_t words, uint32_t *out, uint16_t *h16, Gather gather)
{
    const int32_t groups = (len + 7) / 8, full = len / 8;
    uint64_t badacc = 0;
    auto code = [&](uint64_t x) -> uint64_t {
        const uint64_t c = (x >> 1) & 0x0303030303030303ull;
        const uint64_t is2 = (c >> 1) & ~c & 0x0101010101010101ull;  // code 2 <=> 'T' (0x54 = 0x41 + 2*2 + 15)
        badacc |= (0x4141414141414141ull + 2ull * c + 15ull * is2) ^ x;
        return c;
    };
    for (int32_t g = 0; g < full; ++g) {
        uint64_t x;
        memcpy(&x, row + 8 * g, 8);
        h16[g] = (uint16_t)gather(code(x));
    }
    if (groups > full) {  // last, partial group: pad with 'A' (rows are READ_SIZE, a multiple of 8, bytes: never read past the row)
        uint64_t x;
        memcpy(&x, row + 8 * full, 8);
        const uint64_t keep = (1ull << (8 * (len - 8 * full))) - 1ull;
        x = (x & keep) | (0x4141414141414141ull & ~keep);
        h16[full] = (uint16_t)gather(code(x));
    }
    for (uint32_t w = 0; w < words; ++w) {
        const uint32_t hi = (int32_t)(2 * w) < groups ? h16[2 * w] : 0u, lo = (int32_t)(2 * w + 1) < groups ? h16[2 * w + 1] : 0u;
        out[w] = (hi << 16) | lo;
    }
    return badacc == 0;
}
struct PackJob {
    uint32_t n; int32_t read_size; const int32_t *plen, *tlen; const char *patterns, *texts; uint32_t *packed, *flags; uint32_t words;
};
// flag words [w0, w1): threads own whole flag words (32 pairs), so no two threads touch the same word
template <typename Gather>
inline __attribute__((always_inline)) bool pack_range(const PackJob &J, uint32_t w0, uint32_t w1, Gather gather)
{
    const size_t rs = (size_t)J.read_size;
    std::vector<uint16_t> h16(rs / 8 + 2);
    bool lengths_ok = true;
    for (uint32_t fw = w0; fw < w1; ++fw) {
        uint32_t fbits = 0;
        const uint32_t hi = std::min(J.n, (fw + 1) * 32);
        for (uint32_t i = fw * 32; i < hi; ++i) {
            bool ok = true;
            for (int q = 0; q < 2; ++q) {
                const int32_t len = q ? J.tlen[i] : J.plen[i];
                if (len < 0 || len > J.read_size) { lengths_ok = false; continue; }
                const unsigned char *row = reinterpret_cast<const unsigned char *>((q ? J.texts : J.patterns) + (size_t)i * rs);
                ok &= pack_row(row, len, J.words, J.packed + ((size_t)i * 2 + (size_t)q) * J.words, h16.data(), gather);
            }
            if (!ok) fbits |= 1u << (i & 31);
        }
        J.flags[fw] = fbits;
    }
    return lengths_ok;
}
bool pack_range_swar(const PackJob &J, uint32_t w0, uint32_t w1) { return pack_range(J, w0, w1, [](uint64_t c) { return gather_swar(c); }); }
#if defined(__x86_64__) && defined(__GNUC__)
// 32 bases per step: codes (c >> 1) & 3, the validity test as a table look-up of the letter each code stands for, four codes to
// a byte with two multiply-adds (64 16 4 1, then pairs of 16-bit sums), the four bytes of a word gathered most significant first.
__attribute__((target("avx2"))) bool pack_row_avx2(const unsigned char *row, int32_t len, int32_t row_bytes, uint32_t words, uint32_t *out)
{
    const __m256i three = _mm256_set1_epi8(3), ones16 = _mm256_set1_epi16(1);
    const __m256i lut = _mm256_setr_epi8('A', 'C', 'T', 'G', 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 'A', 'C', 'T', 'G', 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0);
    const __m256i wts = _mm256_setr_epi8(64, 16, 4, 1, 64, 16, 4, 1, 64, 16, 4, 1, 64, 16, 4, 1, 64, 16, 4, 1, 64, 16, 4, 1, 64, 16, 4, 1, 64, 16, 4, 1);
    const __m256i gat = _mm256_setr_epi8(12, 8, 4, 0, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, 12, 8, 4, 0, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1);
    uint32_t w = 0, bad = 0;
    int32_t pos = 0;
    auto chunk = [&](__m256i x, uint32_t counts) __attribute__((target("avx2"))) {
        const __m256i c = _mm256_and_si256(_mm256_srli_epi16(x, 1), three);
        bad |= ~(uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(_mm256_shuffle_epi8(lut, c), x)) & counts;
        const __m256i g = _mm256_shuffle_epi8(_mm256_madd_epi16(_mm256_maddubs_epi16(c, wts), ones16), gat);
        if (w < words) out[w] = (uint32_t)_mm256_extract_epi32(g, 0);
        if (w + 1 < words) out[w + 1] = (uint32_t)_mm256_extract_epi32(g, 4);
        w += 2;
    };
    for (; pos + 32 <= len; pos += 32) chunk(_mm256_loadu_si256(reinterpret_cast<const __m256i *>(row + pos)), 0xffffffffu);
    if (pos < len) {  // last, partial step: 'A' past the end, and never a byte read past the row
        alignas(32) static const unsigned char ramp[64] = {0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff,
                                                           0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff};
        const int32_t rem = len - pos;
        __m256i x;
        if (pos + 32 <= row_bytes) {
            x = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(row + pos));
        } else {
            alignas(32) unsigned char tmp[32] = {};
            memcpy(tmp, row + pos, (size_t)rem);
            x = _mm256_load_si256(reinterpret_cast<const __m256i *>(tmp));
        }
        const __m256i keep = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(ramp + 32 - rem));  // 0xff for the rem bytes that count
        x = _mm256_blendv_epi8(_mm256_set1_epi8('A'), x, keep);
        chunk(x, (1u << rem) - 1u);
    }
    for (; w < words; ++w) out[w] = 0u;
    return bad == 0;
}
__attribute__((target("avx2"))) bool pack_range_avx2(const PackJob &J, uint32_t w0, uint32_t w1)
{
    const size_t rs = (size_t)J.read_size;
    bool lengths_ok = true;
    for (uint32_t fw = w0; fw < w1; ++fw) {
        uint32_t fbits = 0;
        const uint32_t hi = std::min(J.n, (fw + 1) * 32);
        for (uint32_t i = fw * 32; i < hi; ++i) {
            bool ok = true;
            for (int q = 0; q < 2; ++q) {
                const int32_t len = q ? J.tlen[i] : J.plen[i];
                if (len < 0 || len > J.read_size) { lengths_ok = false; continue; }
                const unsigned char *row = reinterpret_cast<const unsigned char *>((q ? J.texts : J.patterns) + (size_t)i * rs);
                ok &= pack_row_avx2(row, len, J.read_size, J.words, J.packed + ((size_t)i * 2 + (size_t)q) * J.words);
            }
            if (!ok) fbits |= 1u << (i & 31);
        }
        J.flags[fw] = fbits;
    }
    return lengths_ok;
}
#endif
#if defined(__x86_64__) && defined(__GNUC__)
__attribute__((target("bmi2"))) bool pack_range_bmi2(const PackJob &J, uint32_t w0, uint32_t w1)
{
    return pack_range(J, w0, w1, [](uint64_t c) __attribute__((target("bmi2"))) { return (uint32_t)__builtin_ia32_pext_di(__builtin_bswap64(c), 0x0303030303030303ull); });
}
#endif
}  // namespace

extern "C" int32_t aim_packed_row_bytes(int32_t read_size)
{
    if (read_size <= 0) return 0;
    const uint32_t words = ((uint32_t)read_size / 16 + 2 + 3) / 4 * 4;  // the short-read kernel's row: one spare word, multiple of 4
    return (int32_t)(words * 4);
}

extern "C" int aim_pack_pairs(uint32_t n, int32_t read_size, const int32_t *plen, const int32_t *tlen, const char *patterns,
                              const char *texts, uint32_t *packed, uint32_t *flags, int32_t nthreads)
{
    if (!plen || !tlen || !patterns || !texts || !packed || !flags || read_size <= 0) { aim::set_error("aim_pack_pairs: NULL argument"); return AIM_ERR_ARG; }
    const size_t rs = (size_t)read_size;
    PackJob J{n, read_size, plen, tlen, patterns, texts, packed, flags, (uint32_t)aim_packed_row_bytes(read_size) / 4};
    const uint32_t fwords = (n + 31) / 32;
    bool (*range)(const PackJob &, uint32_t, uint32_t) = pack_range_swar;
#if defined(__x86_64__) && defined(__GNUC__)
    if (__builtin_cpu_supports("bmi2") && !getenv("AIM_NO_BMI2")) range = pack_range_bmi2;
    if (__builtin_cpu_supports("avx2") && !getenv("AIM_NO_AVX2")) range = pack_range_avx2;
#endif
    int T = nthreads > 0 ? nthreads : io_threads((size_t)n * rs * 2);
    T = std::max(1, std::min<int>(T, (int)std::max<uint32_t>(1, fwords)));
    const uint32_t per = (fwords + (uint32_t)T - 1) / (uint32_t)T;
    std::vector<int> bad((size_t)T, 0);
    parallel_for(T, [&](int t) {
        const uint32_t w0 = std::min(fwords, (uint32_t)t * per), w1 = std::min(fwords, w0 + per);
        if (!range(J, w0, w1)) bad[(size_t)t] = 1;
    });
    for (int t = 0; t < T; ++t) if (bad[(size_t)t]) { aim::set_error("READ LENGTH less than length of the input reads"); return AIM_ERR_LENGTH; }
    return AIM_OK;
}

extern "C" int aim_write_results_packed(const char *path, uint32_t n, const aim_result *results, const char *cigars, int32_t cigar_pitch)
{
    if (!cigars || cigar_pitch <= 0) { aim::set_error("cigars buffer required"); return AIM_ERR_ARG; }
    FILE *f = fopen(path, "w");
    if (!f) { aim::set_error(std::string("Output file '") + path + "' couldn't be opened"); return AIM_ERR_IO; }
    // blocks of pairs formatted by the I/O threads into their own buffers, written in order (as aim_write_results)
    const uint32_t block = 32768;
    const int T = io_threads((size_t)n * ((size_t)cigar_pitch + 24));
    const size_t per_pair_cap = 40 + (size_t)cigar_pitch;
    std::vector<std::vector<char>> buf((size_t)T);
    std::vector<size_t> len((size_t)T, 0);
    int rc = AIM_OK;
    for (uint64_t base = 0; base < n && rc == AIM_OK; base += (uint64_t)block * (uint64_t)T) {
        parallel_for(T, [&](int t) {
            const uint64_t lo = base + (uint64_t)t * block, hi = std::min<uint64_t>(n, lo + block);
            len[(size_t)t] = 0;
            if (lo >= hi) return;
            std::vector<char> &b = buf[(size_t)t];
            if (b.size() < (size_t)(hi - lo) * per_pair_cap) b.resize((size_t)(hi - lo) * per_pair_cap);
            char *o = b.data();
            size_t pos = 0;
            for (uint64_t i = lo; i < hi; ++i) {
                pos += put_int(o + pos, (int32_t)results[i].idx);
                o[pos++] = ','; o[pos++] = ' ';
                pos += put_int(o + pos, results[i].score);
                o[pos++] = ','; o[pos++] = ' '; o[pos++] = '\n';
                const char *c = cigars + (size_t)i * (size_t)cigar_pitch;
                const size_t l = strnlen(c, (size_t)cigar_pitch);
                memcpy(o + pos, c, l);
                pos += l;
                o[pos++] = '\n';
            }
            len[(size_t)t] = pos;
        });
        for (int t = 0; t < T; ++t) {
            if (len[(size_t)t] && fwrite(buf[(size_t)t].data(), 1, len[(size_t)t], f) != len[(size_t)t]) { rc = AIM_ERR_IO; break; }
        }
    }
    if (ferror(f)) rc = AIM_ERR_IO;
    fclose(f);
    return rc;
}

// GenASM printers: DC "%d, %d, %s\n" (idx, score, the DPU's CIGAR string up to its NUL:
// aim-genasm/GenASM/DPU-WRAM-DC/host/host.c:286-296), filter "%d, %d\n" (DPU-WRAM-filter/host/host.c:272).
extern "C" int aim_write_results_genasm(const char *path, uint32_t n, int32_t read_size, int32_t dc,
                                        const aim_result *results, const char *cigars)
{
    if (dc && !cigars) { aim::set_error("cigars buffer required for GenASM-DC"); return AIM_ERR_ARG; }
    FILE *f = fopen(path, "w");
    if (!f) { aim::set_error(std::string("Output file '") + path + "' couldn't be opened"); return AIM_ERR_IO; }
    const size_t rs2 = (size_t)read_size * 2;
    std::vector<char> buf((size_t)1 << 20);
    size_t pos = 0;
    int rc = AIM_OK;
    for (uint32_t i = 0; i < n; ++i) {
        if (pos + rs2 + 64 > buf.size()) {
            if (fwrite(buf.data(), 1, pos, f) != pos) { rc = AIM_ERR_IO; break; }
            pos = 0;
        }
        char *o = buf.data();
        pos += put_int(o + pos, (int32_t)results[i].idx);
        o[pos++] = ','; o[pos++] = ' ';
        pos += put_int(o + pos, results[i].score);
        if (dc) {
            o[pos++] = ','; o[pos++] = ' ';
            const char *c = cigars + (size_t)i * rs2;
            const size_t len = strnlen(c, rs2);
            memcpy(o + pos, c, len);
            pos += len;
        }
        o[pos++] = '\n';
    }
    if (rc == AIM_OK && pos && fwrite(buf.data(), 1, pos, f) != pos) rc = AIM_ERR_IO;
    if (ferror(f)) rc = AIM_ERR_IO;
    fclose(f);
    return rc;
}

extern "C" int aim_write_pairs(const char *path, uint32_t n, int32_t read_size, const int32_t *plen,
                               const int32_t *tlen, const char *patterns, const char *texts)
{
    FILE *f = fopen(path, "w");
    if (!f) { aim::set_error(std::string("Output file '") + path + "' couldn't be opened"); return AIM_ERR_IO; }
    std::vector<char> obuf(1 << 22);
    setvbuf(f, obuf.data(), _IOFBF, obuf.size());
    for (uint32_t i = 0; i < n; ++i) {
        fputc('>', f);
        fwrite(patterns + (size_t)i * read_size, 1, (size_t)plen[i], f);
        fputc('\n', f);
        fputc('<', f);
        fwrite(texts + (size_t)i * read_size, 1, (size_t)tlen[i], f);
        fputc('\n', f);
    }
    int rc = ferror(f) ? AIM_ERR_IO : AIM_OK;
    fclose(f);
    return rc;
}

// ---- synthetic pairs (Datasets/README.md:19-25 names smarco/WFA's generate_dataset; the tool
// itself is not in the reference tree, so its published behaviour is restated here) ----------
namespace {
struct SplitMix64 {
    uint64_t s;
    explicit SplitMix64(uint64_t seed) : s(seed) {}
    uint64_t next()
    {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    uint32_t below(uint32_t n) { return (uint32_t)(((next() >> 32) * (uint64_t)n) >> 32); }
};

void generate_one(uint64_t seed, uint64_t pair, int32_t length, int32_t nerr, int32_t read_size,
                  int32_t *plen, int32_t *tlen, char *pat, char *txt)
{
    static const char kBases[4] = {'A', 'C', 'G', 'T'};
    SplitMix64 rng(seed * 0xD1B54A32D192ED03ull + pair * 0x9E3779B97F4A7C15ull + 0x2545F4914F6CDD1Dull);
    for (int i = 0; i < length; ++i) pat[i] = kBases[rng.below(4)];
    memcpy(txt, pat, (size_t)length);
    int tl = length;
    for (int e = 0; e < nerr; ++e) {
        uint32_t kind = rng.below(3);
        if (kind == 0 && tl > 0) {  // mismatch to a different base
            uint32_t pos = rng.below((uint32_t)tl);
            char c;
            do { c = kBases[rng.below(4)]; } while (c == txt[pos]);
            txt[pos] = c;
        } else if (kind == 1 && tl > 0) {  // delete one base
            uint32_t pos = rng.below((uint32_t)tl);
            memmove(txt + pos, txt + pos + 1, (size_t)(tl - 1 - (int)pos));
            --tl;
        } else if (tl < read_size) {  // insert one uniform base
            uint32_t pos = rng.below((uint32_t)tl + 1);
            memmove(txt + pos + 1, txt + pos, (size_t)(tl - (int)pos));
            txt[pos] = kBases[rng.below(4)];
            ++tl;
        }
    }
    if (length < read_size) memset(pat + length, 0, (size_t)(read_size - length));
    if (tl < read_size) memset(txt + tl, 0, (size_t)(read_size - tl));
    *plen = length;
    *tlen = tl;
}
}  // namespace

extern "C" int aim_generate_pairs(uint64_t seed, uint64_t first_pair, uint32_t n, int32_t length, double error,
                                  int32_t read_size, int32_t *plen, int32_t *tlen, char *patterns, char *texts,
                                  int32_t nthreads)
{
    if (length <= 0 || length > read_size || !plen || !tlen || !patterns || !texts) return AIM_ERR_ARG;
    int32_t nerr = (int32_t)std::ceil((double)length * error);
    if (nthreads < 1) nthreads = 1;
    auto work = [&](uint32_t a, uint32_t b) {
        for (uint32_t i = a; i < b; ++i)
            generate_one(seed, first_pair + i, length, nerr, read_size, &plen[i], &tlen[i],
                         patterns + (size_t)i * read_size, texts + (size_t)i * read_size);
    };
    if (nthreads == 1 || n < 1024) { work(0, n); return AIM_OK; }
    std::vector<std::thread> th;
    uint32_t per = (n + (uint32_t)nthreads - 1) / (uint32_t)nthreads;
    for (int t = 0; t < nthreads; ++t) {
        uint32_t a = (uint32_t)t * per, b = a + per;
        if (a >= n) break;
        if (b > n) b = n;
        th.emplace_back(work, a, b);
    }
    for (auto &t : th) t.join();
    return AIM_OK;
}
