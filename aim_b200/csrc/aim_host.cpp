// Host-side logic of the drop-in boundary that needs no GPU: knob derivation, the pairs-file
// reader, the result writer, the pairs-to-process rule and the synthetic-pair generator.
// Mirrors the reference's host program (citations: paths under the reference checkout).
#include "aim_b200.h"
#include "aim_internal.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace aim {
thread_local std::string g_last_error;
void set_error(const std::string &msg) { g_last_error = msg; }
}  // namespace aim

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
struct LineReader {
    FILE *f;
    char *buf = nullptr;
    size_t cap = 0;
    explicit LineReader(FILE *fp) : f(fp) {}
    ~LineReader() { free(buf); }
    // getline(3) semantics: length including the newline, -1 at EOF.
    long next() { return (long)getline(&buf, &cap, f); }
};
}  // namespace

extern "C" int64_t aim_count_pairs(const char *path)
{
    FILE *f = fopen(path, "r");
    if (!f) { aim::set_error(std::string("Input file '") + path + "' couldn't be opened"); return AIM_ERR_IO; }
    int64_t lines = 0;
    std::vector<char> chunk(1 << 20);
    size_t got;
    char last = '\n';
    while ((got = fread(chunk.data(), 1, chunk.size(), f)) > 0) {
        for (size_t i = 0; i < got; ++i) lines += chunk[i] == '\n';
        last = chunk[got - 1];
    }
    if (last != '\n') ++lines;  // getline also returns a final unterminated line
    fclose(f);
    return lines / 2;
}

// host.c:91-134 get_reads.  pattern = line+1, length = line_length-2 (the first character and the
// last one, assumed '\n', are dropped without being looked at).
extern "C" int64_t aim_read_pairs(const char *path, uint32_t max_pairs, int32_t read_size,
                                  int32_t *plen, int32_t *tlen, char *patterns, char *texts)
{
    FILE *f = fopen(path, "r");
    if (!f) { aim::set_error(std::string("Input file '") + path + "' couldn't be opened"); return AIM_ERR_IO; }
    LineReader l1(f), l2(f);
    int64_t n = 0;
    for (; n < (int64_t)max_pairs; ++n) {
        long len1 = l1.next();
        if (len1 == -1) break;
        long len2 = l2.next();
        if (len2 == -1) break;
        long pl = len1 - 2, tl = len2 - 2;
        if (tl > read_size || pl > read_size) {
            fclose(f);
            aim::set_error("READ LENGTH less than length of the input reads");
            return AIM_ERR_LENGTH;
        }
        if (pl < 0) pl = 0;  // a 1-character line: the reference would index pattern[-1]; we clamp
        if (tl < 0) tl = 0;
        char *pd = patterns + (size_t)n * read_size, *td = texts + (size_t)n * read_size;
        memcpy(pd, l1.buf + 1, (size_t)pl);
        memcpy(td, l2.buf + 1, (size_t)tl);
        if (pl < read_size) memset(pd + pl, 0, (size_t)(read_size - pl));
        if (tl < read_size) memset(td + tl, 0, (size_t)(read_size - tl));
        plen[n] = (int32_t)pl;
        tlen[n] = (int32_t)tl;
    }
    fclose(f);
    return n;
}

// host.c:69-89 edit_cigar_print.
extern "C" int aim_cigar_rle(const char *ops, int32_t begin_offset, int32_t end_offset, char *out, size_t cap)
{
    size_t pos = 0;
    char last_op = ops[begin_offset];
    int last_len = 1;
    for (int i = begin_offset + 1; i < end_offset; ++i) {
        if (ops[i] == last_op) { ++last_len; continue; }
        int w = snprintf(out + pos, cap - pos, "%d%c", last_len, last_op);
        if (w < 0 || (size_t)w >= cap - pos) return -1;
        pos += (size_t)w;
        last_op = ops[i];
        last_len = 1;
    }
    int w = snprintf(out + pos, cap - pos, "%d%c", last_len, last_op);
    if (w < 0 || (size_t)w >= cap - pos) return -1;
    return (int)(pos + (size_t)w);
}

// host.c:332-353: "%d, %d, \n" then (BACKTRACE only) the RLE CIGAR on its own line.
extern "C" int aim_write_results(const char *path, uint32_t n, int32_t read_size, int32_t backtrace,
                                 const aim_result *results, const char *ops)
{
    FILE *f = fopen(path, "w");
    if (!f) { aim::set_error(std::string("Output file '") + path + "' couldn't be opened"); return AIM_ERR_IO; }
    std::vector<char> obuf(1 << 22);
    setvbuf(f, obuf.data(), _IOFBF, obuf.size());
    std::vector<char> line((size_t)read_size * 2 * 12 + 64);
    for (uint32_t i = 0; i < n; ++i) {
        fprintf(f, "%d, %d, \n", (int)results[i].idx, results[i].score);
        if (backtrace) {
            int len = aim_cigar_rle(ops + (size_t)i * 2 * read_size, results[i].begin_offset,
                                    results[i].end_offset, line.data(), line.size() - 1);
            if (len < 0) { fclose(f); return AIM_ERR_IO; }
            line[(size_t)len] = '\n';
            fwrite(line.data(), 1, (size_t)len + 1, f);
        }
    }
    int rc = ferror(f) ? AIM_ERR_IO : AIM_OK;
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
