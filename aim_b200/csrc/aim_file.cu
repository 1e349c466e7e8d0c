// `host <pairs-file> <out-file> <N>` at device rate: the pair-file parser and the result printer of the reference host
// (WFA/DPU-MRAM/host/host.c: get_reads :91-134, edit_cigar_print :69-89, the print loop :332-353) as CUDA kernels, so that
// the host CPU only moves bytes: read() the file into pinned chunks, write() the formatted text.
//
// Per chunk of file bytes (cut by the host after a whole number of line PAIRS, so chunks are independent):
//   nl_count / tile_scan / nl_scatter   newline positions of the chunk            (get_reads' two getline() calls)
//   extract_rows                        line l -> row (l >> 1) of the pattern (even l) or text (odd l) buffer, first and
//                                       last character of the line dropped unchecked (host.c:112-117),
//                                       length = line_length - 2, rows zero-padded to READ_SIZE; a length > READ_SIZE
//                                       raises the flag the reference turns into message + exit(0) (host.c:119-123)
//   <the alignment kernels on those device buffers: aim_dispatch.cu launch()>
//   fmt_len / scan / fmt_write          "%d, %d, \n" (idx, score) and, with BACKTRACE, the run-length CIGAR of
//                                       ops[begin_offset, end_offset) + "\n" (host.c:340-350, :69-89), packed densely in
//                                       pair order: the bytes of the output file
// All of it is HBM-bound byte work (about 1 KB per pair at config 4, a few ms per 10 M pairs).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <string>

#include "aim_internal.h"

namespace aim {

namespace {

constexpr int TILE_THREADS = 256;
constexpr int TILE_BYTES = TILE_THREADS * 16;  // one uint4 per thread

// bit j of the result = byte j of the 16-byte vector is '\n'
__device__ __forceinline__ uint32_t nl_mask16(uint4 v)
{
    const uint32_t nl = 0x0a0a0a0au;
    // __vcmpeq4 gives 0xff per equal byte; keep one bit per byte and gather the four bits of each word
    auto bits = [](uint32_t eq) { const uint32_t b = eq & 0x01010101u; return (b | (b >> 7) | (b >> 14) | (b >> 21)) & 0xfu; };
    return bits(__vcmpeq4(v.x, nl)) | (bits(__vcmpeq4(v.y, nl)) << 4) | (bits(__vcmpeq4(v.z, nl)) << 8) | (bits(__vcmpeq4(v.w, nl)) << 12);
}

__device__ __forceinline__ uint4 load16_guard(const char *buf, size_t pos, size_t nbytes)
{
    // buf is 16-byte aligned and padded to a multiple of 16 by the caller; bytes at or beyond nbytes are ignored
    uint4 v = __ldg(reinterpret_cast<const uint4 *>(buf + pos));
    if (pos + 16 > nbytes) {
        uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const size_t p = pos + 4 * (size_t)k;
            if (p >= nbytes) w[k] = 0;
            else if (p + 4 > nbytes) w[k] &= (1u << (8 * (uint32_t)(nbytes - p))) - 1u;
        }
        v = make_uint4(w[0], w[1], w[2], w[3]);
    }
    return v;
}

__global__ void __launch_bounds__(TILE_THREADS) nl_count_kernel(const char *buf, size_t nbytes, uint32_t *tile_count)
{
    const size_t pos = ((size_t)blockIdx.x * TILE_THREADS + threadIdx.x) * 16;
    uint32_t c = 0;
    if (pos < nbytes) c = (uint32_t)__popc(nl_mask16(load16_guard(buf, pos, nbytes)));
    __shared__ uint32_t ws[TILE_THREADS / 32];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t s = 0;
#pragma unroll
        for (int k = 0; k < TILE_THREADS / 32; ++k) s += ws[k];
        tile_count[blockIdx.x] = s;
    }
}

// exclusive scan of `count[0..n)` in place by ONE block (n is a few thousand); total -> *total
__global__ void __launch_bounds__(1024) tile_scan_kernel(uint32_t *count, uint32_t n, uint32_t *total)
{
    __shared__ uint32_t ws[32];
    __shared__ uint32_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < n ? count[i] : 0u;
        uint32_t x = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, d); if ((threadIdx.x & 31) >= (uint32_t)d) x += y; }
        if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t w = ws[threadIdx.x];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, w, d); if (threadIdx.x >= (uint32_t)d) w += y; }
            ws[threadIdx.x] = w;
        }
        __syncthreads();
        const uint32_t carry = carry_s;
        const uint32_t incl = x + (threadIdx.x >= 32 ? ws[(threadIdx.x >> 5) - 1] : 0u) + carry;
        if (i < n) count[i] = incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry_s;
}

// nl_pos[r] = byte offset of the r-th '\n' of the chunk (r < max_lines)
__global__ void __launch_bounds__(TILE_THREADS) nl_scatter_kernel(const char *buf, size_t nbytes, const uint32_t *tile_base, uint32_t *nl_pos,
                                                                  uint32_t max_lines)
{
    const size_t pos = ((size_t)blockIdx.x * TILE_THREADS + threadIdx.x) * 16;
    uint32_t m = 0;
    if (pos < nbytes) m = nl_mask16(load16_guard(buf, pos, nbytes));
    const uint32_t c = (uint32_t)__popc(m);
    uint32_t x = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, d); if ((threadIdx.x & 31) >= (uint32_t)d) x += y; }
    __shared__ uint32_t ws[TILE_THREADS / 32];
    if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = x;
    __syncthreads();
    uint32_t before = tile_base[blockIdx.x] + x - c;
    for (uint32_t k = 0; k < (threadIdx.x >> 5); ++k) before += ws[k];
    while (m) {
        const int j = __ffs((int)m) - 1;
        m &= m - 1;
        if (before < max_lines) nl_pos[before] = (uint32_t)pos + (uint32_t)j;
        ++before;
    }
}

// One sub-warp of 8 lanes per line: line l spans (nl_pos[l-1], nl_pos[l]] (the newline included, as getline returns it);
// the sequence is line[1 .. len_with_nl - 1), its length len_with_nl - 2 (clamped at 0).  `unterminated` : the chunk's last
// line has no '\n' (end of file): nl_pos[lines-1] is then the offset one past the last byte and the line is one byte shorter.
__global__ void __launch_bounds__(256) extract_rows_kernel(const char *buf, const uint32_t *nl_pos, uint32_t lines, int unterminated, int RS,
                                                           int32_t *plen, int32_t *tlen, char *patterns, char *texts, uint32_t *too_long)
{
    const uint32_t gl = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;  // line handled by this 8-lane group
    const uint32_t sl = threadIdx.x & 7;
    if (gl >= lines) return;
    const uint32_t start = gl ? nl_pos[gl - 1] + 1u : 0u;
    uint32_t len_with_nl = nl_pos[gl] + 1u - start;
    if (unterminated && gl == lines - 1) --len_with_nl;
    int len = (int)len_with_nl - 2;
    const uint32_t pair = gl >> 1;
    if (len > RS) {  // host.c:119-123
        if (sl == 0) atomicOr(too_long, 1u);
        return;
    }
    if (len < 0) len = 0;  // a 1-character line: the reference indexes pattern[-1]; clamped (as aim_read_pairs does)
    if (sl == 0) ((gl & 1u) ? tlen : plen)[pair] = len;
    uint32_t *row = reinterpret_cast<uint32_t *>(((gl & 1u) ? texts : patterns) + (size_t)pair * RS);
    const uint32_t src0 = start + 1u;  // first sequence byte
    const uint32_t *in32 = reinterpret_cast<const uint32_t *>(buf);
    for (int w = (int)sl; w * 4 < RS; w += 8) {
        uint32_t v = 0;
        if (w * 4 < len) {
            const uint32_t s = src0 + 4u * (uint32_t)w;
            const uint32_t a = __ldg(in32 + (s >> 2)), b = __ldg(in32 + (s >> 2) + 1);  // (the buffer is padded by 8 bytes)
            v = __funnelshift_r(a, b, 8u * (s & 3u));
            const int rem = len - w * 4;
            if (rem < 4) v &= (1u << (8 * rem)) - 1u;
        }
        row[w] = v;
    }
}

// ---- output text ----
__device__ __forceinline__ int ndigits(uint32_t v)
{
    return v < 10u ? 1 : v < 100u ? 2 : v < 1000u ? 3 : v < 10000u ? 4 : v < 100000u ? 5 : v < 1000000u ? 6 : v < 10000000u ? 7
           : v < 100000000u ? 8 : v < 1000000000u ? 9 : 10;
}
__device__ __forceinline__ int ndigits_signed(int v) { return v < 0 ? 1 + ndigits((uint32_t)(-(int64_t)v)) : ndigits((uint32_t)v); }
__device__ __forceinline__ char *put_uint(char *o, uint32_t v)
{
    const int n = ndigits(v);
    for (int k = n - 1; k >= 0; --k) { o[k] = (char)('0' + v % 10u); v /= 10u; }
    return o + n;
}
__device__ __forceinline__ char *put_int(char *o, int v)
{
    if (v < 0) { *o++ = '-'; return put_uint(o, (uint32_t)(-(int64_t)v)); }
    return put_uint(o, (uint32_t)v);
}

// Runs of ops[b, e) (b >= 0, e > b), eight ops per step: f(run_length, op) in order.  Rows are 8-byte aligned.
template <typename F>
__device__ __forceinline__ void for_each_run(const char *row, int b, int e, F &&f)
{
    const unsigned long long *row64 = reinterpret_cast<const unsigned long long *>(row);
    const int w0 = b >> 3, w1 = (e - 1) >> 3;
    int run_start = b;
    unsigned long long x = row64[w0];
    for (int w = w0; w <= w1; ++w) {
        const unsigned long long nx = w < w1 ? row64[w + 1] : x;
        // byte k of diff is non-zero where op 8w+k differs from op 8w+k+1
        unsigned long long diff = x ^ ((x >> 8) | (nx << 56));
        const int lo_k = max(b - 8 * w, 0), hi_k = min(e - 1 - 8 * w, 8);  // boundaries between ops j, j+1 with b <= j, j+1 <= e-1
        if (lo_k > 0) diff &= ~0ull << (8 * lo_k);
        if (hi_k < 8) diff &= hi_k > 0 ? ~(~0ull << (8 * hi_k)) : 0ull;
        while (diff) {
            const int k = (__ffsll((long long)diff) - 1) >> 3;
            const int j = 8 * w + k;
            f(j - run_start + 1, (char)((x >> (8 * k)) & 0xffull));
            run_start = j + 1;
            diff &= ~(0xffull << (8 * k));
        }
        x = nx;
    }
    f(e - run_start, row[e - 1]);
}

// bytes of pair i's output lines; also ORs the pair's status into *status_or (bit s set = some pair has status s)
__global__ void __launch_bounds__(128) fmt_len_kernel(const aim_result *results, const char *ops, int RS, int bt, int mode, uint32_t m, uint32_t *lens,
                                                      uint32_t *status_or)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const aim_result r = results[i];
    if (r.status != AIM_STATUS_OK) atomicOr(status_or, 1u << (r.status & 31));
    if (mode) {  // GenASM: "%d, %d, %s\n" (DC: the CIGAR string of the op row, aim-genasm/GenASM/DPU-WRAM-DC/host/host.c:286-296) / "%d, %d\n" (filter)
        uint32_t len = (uint32_t)(ndigits_signed((int)r.idx) + ndigits_signed(r.score) + 3);
        if (mode == 1) {
            const char *row = ops + (size_t)i * 2 * RS;
            int sl = 0;
            while (sl < 2 * RS && row[sl]) ++sl;
            len += 2u + (uint32_t)sl;
        }
        lens[i] = len;
        return;
    }
    uint32_t len = (uint32_t)(ndigits_signed((int)r.idx) + ndigits_signed(r.score) + 5);  // "%d, %d, \n"
    if (bt) {
        // edit_cigar_print always prints the op at begin_offset, even for an empty span; a span that starts before the
        // row (empty pair, begin_offset = -1) prints "1M" (cigar_rle_kernel, aim_cigar_rle)
        const int b = r.begin_offset, e = min(r.end_offset > b ? r.end_offset : b + 1, 2 * RS);
        if (b < 0 || b >= 2 * RS) len += 2;
        else for_each_run(ops + (size_t)i * 2 * RS, b, e, [&](int run, char) { len += (uint32_t)ndigits((uint32_t)run) + 1u; });
        len += 1;  // '\n'
    }
    lens[i] = len;
}

// (nothing is written when the chunk's text does not fit the buffer: counters[2] = total bytes; the host grows the buffer
// and runs this kernel again)
__global__ void __launch_bounds__(128) fmt_write_kernel(const aim_result *results, const char *ops, int RS, int bt, int mode, uint32_t m,
                                                        const uint32_t *offs, const uint32_t *counters, size_t out_cap, char *out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m || (size_t)counters[2] > out_cap) return;
    const aim_result r = results[i];
    char *o = out + offs[i];
    o = put_int(o, (int)r.idx);  // the reference prints the uint32 idx with %d
    *o++ = ','; *o++ = ' ';
    o = put_int(o, r.score);
    if (mode) {
        if (mode == 1) {
            *o++ = ','; *o++ = ' ';
            const char *row = ops + (size_t)i * 2 * RS;
            for (int sl = 0; sl < 2 * RS && row[sl]; ++sl) *o++ = row[sl];
        }
        *o++ = '\n';
        return;
    }
    *o++ = ','; *o++ = ' '; *o++ = '\n';
    if (bt) {
        const int b = r.begin_offset, e = min(r.end_offset > b ? r.end_offset : b + 1, 2 * RS);
        if (b < 0 || b >= 2 * RS) { *o++ = '1'; *o++ = 'M'; }
        else for_each_run(ops + (size_t)i * 2 * RS, b, e, [&](int run, char op) { o = put_uint(o, (uint32_t)run); *o++ = op; });
        *o++ = '\n';
    }
}

// ---- exclusive scan of m uint32 values (m up to a few million): block sums, scan of the sums, local scan ----
constexpr int SCAN_THREADS = 256, SCAN_PER_THREAD = 4, SCAN_TILE = SCAN_THREADS * SCAN_PER_THREAD;

__global__ void __launch_bounds__(SCAN_THREADS) scan_sums_kernel(const uint32_t *v, uint32_t m, uint32_t *sums)
{
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_PER_THREAD;
    uint32_t c = 0;
#pragma unroll
    for (int k = 0; k < SCAN_PER_THREAD; ++k) if (base + k < m) c += v[base + k];
    __shared__ uint32_t ws[SCAN_THREADS / 32];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t s = 0;
#pragma unroll
        for (int k = 0; k < SCAN_THREADS / 32; ++k) s += ws[k];
        sums[blockIdx.x] = s;
    }
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_local_kernel(const uint32_t *v, uint32_t m, const uint32_t *tile_base, uint32_t *out)
{
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_PER_THREAD;
    uint32_t a[SCAN_PER_THREAD], c = 0;
#pragma unroll
    for (int k = 0; k < SCAN_PER_THREAD; ++k) { a[k] = base + k < m ? v[base + k] : 0u; c += a[k]; }
    uint32_t x = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, d); if ((threadIdx.x & 31) >= (uint32_t)d) x += y; }
    __shared__ uint32_t ws[SCAN_THREADS / 32];
    if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = x;
    __syncthreads();
    uint32_t before = tile_base[blockIdx.x] + x - c;
    for (uint32_t k = 0; k < (threadIdx.x >> 5); ++k) before += ws[k];
#pragma unroll
    for (int k = 0; k < SCAN_PER_THREAD; ++k) {
        if (base + k < m) out[base + k] = before;
        before += a[k];
    }
}

}  // namespace

size_t file_parse_scratch_bytes(size_t chunk_bytes) { return ((chunk_bytes + TILE_BYTES - 1) / TILE_BYTES + 64) * sizeof(uint32_t); }
size_t file_format_scratch_bytes(uint32_t max_pairs) { return ((size_t)(max_pairs + SCAN_TILE - 1) / SCAN_TILE + 64) * sizeof(uint32_t); }

// Newline positions of the chunk + rows.  d_buf: the chunk's bytes (16-byte aligned, readable 16 bytes past nbytes);
// lines = number of lines the host counted (2 * pairs); tiles: file_parse_scratch_bytes(); counters[0] receives the
// newline count the GPU found (the host's count must agree), counters[1] the too-long flag.
int launch_file_parse(const char *d_buf, size_t nbytes, uint32_t lines, int unterminated, int read_size, uint32_t *d_tiles, uint32_t *d_nl_pos,
                      uint32_t *d_counters, int32_t *d_plen, int32_t *d_tlen, char *d_pat, char *d_txt, void *stream_v, int *launches)
{
    cudaStream_t st = (cudaStream_t)stream_v;
    if (lines == 0) return AIM_OK;
    const uint32_t ntiles = (uint32_t)((nbytes + TILE_BYTES - 1) / TILE_BYTES);
    nl_count_kernel<<<ntiles, TILE_THREADS, 0, st>>>(d_buf, nbytes, d_tiles);
    tile_scan_kernel<<<1, 1024, 0, st>>>(d_tiles, ntiles, d_counters);
    nl_scatter_kernel<<<ntiles, TILE_THREADS, 0, st>>>(d_buf, nbytes, d_tiles, d_nl_pos, lines);
    const uint64_t threads = (uint64_t)lines * 8;
    extract_rows_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(d_buf, d_nl_pos, lines, unterminated, read_size, d_plen, d_tlen, d_pat,
                                                                          d_txt, d_counters + 1);
    if (cudaGetLastError() != cudaSuccess) { set_error("file parse launch failed"); return AIM_ERR_CUDA; }
    if (launches) *launches += 4;
    return AIM_OK;
}

// Output text of m pairs, densely packed at d_out; d_lens / d_offs: m words each; d_tiles: file_format_scratch_bytes();
// counters[2] receives the total byte count, counters[3] the OR of (1 << status) over the pairs.
int launch_file_format(const aim_result *d_res, const char *d_ops, int read_size, int backtrace, int mode, uint32_t m, uint32_t *d_lens, uint32_t *d_offs,
                       uint32_t *d_tiles, uint32_t *d_counters, char *d_out, size_t out_cap, void *stream_v, int *launches)
{
    cudaStream_t st = (cudaStream_t)stream_v;
    if (m == 0) return AIM_OK;
    const uint32_t ntiles = (m + SCAN_TILE - 1) / SCAN_TILE;
    fmt_len_kernel<<<(m + 127) / 128, 128, 0, st>>>(d_res, d_ops, read_size, backtrace, mode, m, d_lens, d_counters + 3);
    scan_sums_kernel<<<ntiles, SCAN_THREADS, 0, st>>>(d_lens, m, d_tiles);
    tile_scan_kernel<<<1, 1024, 0, st>>>(d_tiles, ntiles, d_counters + 2);
    scan_local_kernel<<<ntiles, SCAN_THREADS, 0, st>>>(d_lens, m, d_tiles, d_offs);
    fmt_write_kernel<<<(m + 127) / 128, 128, 0, st>>>(d_res, d_ops, read_size, backtrace, mode, m, d_offs, d_counters, out_cap, d_out);
    if (cudaGetLastError() != cudaSuccess) { set_error("file format launch failed"); return AIM_ERR_CUDA; }
    if (launches) *launches += 5;
    return AIM_OK;
}

int launch_file_format_write(const aim_result *d_res, const char *d_ops, int read_size, int backtrace, int mode, uint32_t m, const uint32_t *d_offs,
                             const uint32_t *d_counters, char *d_out, size_t out_cap, void *stream_v, int *launches)
{
    if (m == 0) return AIM_OK;
    fmt_write_kernel<<<(m + 127) / 128, 128, 0, (cudaStream_t)stream_v>>>(d_res, d_ops, read_size, backtrace, mode, m, d_offs, d_counters, out_cap, d_out);
    if (cudaGetLastError() != cudaSuccess) { set_error("file format launch failed"); return AIM_ERR_CUDA; }
    if (launches) ++*launches;
    return AIM_OK;
}


// ---- op rows as RUN rows (aim_align_batch's download) ----
// The op row the reference pulls back per pair is 2 * READ_SIZE bytes (host.c:316-326), 'M' but for a handful of runs; the
// device-to-host direction is the scarcer one of the platform, so the row crosses PCIe as its non-'M' runs and
// the host writes the row out again (expand_op_runs, aim_host.cpp).  A run row = `pitch` bytes = 32-bit words: the number of
// runs, then one word per run of bytes other than 'M' over the WHOLE row: position | length (1..255) << 16 | op << 24 (so
// the rebuilt row is the device row byte for byte); 0xffffffff in the first word = more runs than the row holds, the op row
// is then fetched as it is.  One pair per thread, eight ops per step, all-'M' words skipped.
namespace {
__global__ void __launch_bounds__(128) op_runs_kernel(const char *ops, int row_bytes, unsigned char *runs, int pitch, uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long *row = reinterpret_cast<const unsigned long long *>(ops + (size_t)i * row_bytes);
    uint32_t *out = reinterpret_cast<uint32_t *>(runs + (size_t)i * pitch);
    const int nw = row_bytes >> 3;
    const uint32_t cap = (uint32_t)(pitch >> 2) - 1u;
    const unsigned long long M8 = 0x4d4d4d4d4d4d4d4dull;
    uint32_t cnt = 0;
    int start = -1;  // first byte of the open non-'M' run
    uint32_t op = 0;
    auto close = [&](int end) {
        for (int s = start; s < end; s += 255) {
            if (cnt < cap) out[1 + cnt] = (uint32_t)s | ((uint32_t)min(end - s, 255) << 16) | (op << 24);
            ++cnt;
        }
        start = -1;
    };
    for (int w = 0; w < nw; ++w) {
        const unsigned long long x = row[w];
        if (x == M8) {
            if (start >= 0) close(8 * w);
            continue;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t c = (uint32_t)(x >> (8 * k)) & 0xffu;
            if (start >= 0 && c != op) close(8 * w + k);
            if (c != 'M' && start < 0) { start = 8 * w + k; op = c; }
        }
    }
    if (start >= 0) close(row_bytes);
    out[0] = cnt <= cap ? cnt : 0xffffffffu;
}
}  // namespace

// GenASM-DC: the op row holds the DPU's CIGAR string (NUL-terminated, a few dozen bytes of 2 * READ_SIZE): its first `pitch` bytes
// are what crosses PCIe (one 16-byte piece per thread); a string that does not end inside them is fetched with its whole row.
namespace {
__global__ void __launch_bounds__(256) str_rows_kernel(const char *ops, int row_bytes, unsigned char *rows, int pitch, uint32_t n)
{
    const uint32_t per = (uint32_t)pitch >> 4;
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (uint64_t)n * per) return;
    const uint32_t i = (uint32_t)(t / per), c = (uint32_t)(t % per);
    reinterpret_cast<uint4 *>(rows + (size_t)i * pitch)[c] = reinterpret_cast<const uint4 *>(ops + (size_t)i * row_bytes)[c];
}
}  // namespace

int launch_str_rows(const char *d_ops, int read_size, uint32_t m, unsigned char *d_rows, int pitch, void *stream_v, int *launches)
{
    if (m == 0) return AIM_OK;
    const uint64_t total = (uint64_t)m * (uint64_t)(pitch >> 4);
    str_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream_v>>>(d_ops, 2 * read_size, d_rows, pitch, m);
    if (cudaGetLastError() != cudaSuccess) { set_error("string rows launch failed"); return AIM_ERR_CUDA; }
    if (launches) ++*launches;
    return AIM_OK;
}

int launch_op_runs(const char *d_ops, int read_size, uint32_t m, unsigned char *d_runs, int pitch, void *stream_v, int *launches)
{
    if (m == 0) return AIM_OK;
    op_runs_kernel<<<(m + 127) / 128, 128, 0, (cudaStream_t)stream_v>>>(d_ops, 2 * read_size, d_runs, pitch, m);
    if (cudaGetLastError() != cudaSuccess) { set_error("op runs launch failed"); return AIM_ERR_CUDA; }
    if (launches) ++*launches;
    return AIM_OK;
}

}  // namespace aim
