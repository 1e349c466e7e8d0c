// libaim_dpu.so — the UPMEM host-API calls of include/dpu.h on top of the C ABI of libaim_b200.so.
//
// The reference's hosts (WFA/DPU-MRAM/host/host.c:186-372 and its five siblings) talk to their DPUs through
// dpu_alloc / dpu_load / dpu_prepare_xfer / dpu_push_xfer / dpu_launch; with this library under them they compile
// unchanged and the "DPU program" is the B200 aligner.  Every virtual DPU owns an MRAM-heap IMAGE in host memory laid out by
// the host itself (host.c:215-241: DPUParams at heap offset 0, then requests, results, patterns, texts, ops); dpu_launch
// gathers the pairs of all images, aligns them in one aim_align_batch() call and writes result_t records and op rows back
// where the DPU program's main() would have (WFA/DPU-MRAM/dpu/wfa.c:507-533).  Host-only code: no CUDA here.
#include <algorithm>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "aim_b200.h"
#include "dpu.h"

struct aim_dpu_image {
    uint8_t *mem = nullptr;
    size_t size = 0, cap = 0;
    void *prepared = nullptr;  // dpu_prepare_xfer's pointer, consumed by the next dpu_push_xfer
};

struct aim_dpu_system {
    std::vector<aim_dpu_image> dpu;
    bool loaded = false;
    aim_params params{};
    aim_dpu_knobs knobs{};
    double phase_ms[3] = {0, 0, 0};
    // pinned gather buffers, kept from launch to launch
    size_t cap_pairs = 0, cap_rs = 0;
    bool cap_ops = false;
    int32_t *plen = nullptr, *tlen = nullptr;
    char *patterns = nullptr, *texts = nullptr, *ops = nullptr;
    aim_result *results = nullptr;
};

namespace {

thread_local std::string g_msg;

dpu_error_t fail(dpu_error_t e, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_msg = buf;
    return e;
}

// DPUParams (WFA/DPU-MRAM/common/common.h:189-199): 8 x uint32
struct WireParams {
    uint32_t dpuNumReads, dpuRequests_m, dpuResults_m, dpuPatterns_m, dpuTexts_m, dpuOperations_m, mramTotalAllocated, padding;
};

bool image_reserve(aim_dpu_image &im, size_t bytes)
{
    if (bytes <= im.size) return true;
    if (bytes > im.cap) {
        size_t cap = std::max<size_t>(bytes, im.cap + im.cap / 2);
        cap = (cap + 4095) & ~(size_t)4095;
        uint8_t *m = (uint8_t *)realloc(im.mem, cap);
        if (!m) return false;
        im.mem = m;
        im.cap = cap;
    }
    memset(im.mem + im.size, 0, bytes - im.size);
    im.size = bytes;
    return true;
}

void free_gather(aim_dpu_system *s)
{
    aim_host_free(s->plen); aim_host_free(s->tlen); aim_host_free(s->patterns); aim_host_free(s->texts);
    aim_host_free(s->ops); aim_host_free(s->results);
    s->plen = s->tlen = nullptr;
    s->patterns = s->texts = s->ops = nullptr;
    s->results = nullptr;
    s->cap_pairs = 0;
}

// run fn(d) for every DPU of the set, on a few host threads when the copies are large
template <class F>
void for_each_dpu(aim_dpu_system *s, int32_t one, size_t bytes_each, F &&fn)
{
    const size_t nd = s->dpu.size();
    if (one >= 0) { fn((size_t)one); return; }
    const size_t total = bytes_each * nd;
    unsigned nt = total >= ((size_t)32 << 20) ? std::min<unsigned>(16, std::max(1u, std::thread::hardware_concurrency())) : 1;
    nt = (unsigned)std::min<size_t>(nt, nd);
    if (nt <= 1) { for (size_t d = 0; d < nd; ++d) fn(d); return; }
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t)
        th.emplace_back([&, t]() { for (size_t d = t; d < nd; d += nt) fn(d); });
    for (auto &x : th) x.join();
}

}  // namespace

extern "C" const char *dpu_error_to_string(dpu_error_t status)
{
    static thread_local std::string out;
    static const char *names[] = {"success", "internal error", "system error", "driver error", "allocation error", "invalid DPU set",
                                  "invalid symbol access", "unknown symbol", "invalid MRAM access", "transfer already set",
                                  "different DPU programs", "no program loaded", "DPU fault", "no such ELF file"};
    out = (status >= 0 && status <= DPU_ERR_ELF_NO_SUCH_FILE) ? names[status] : "unknown error";
    if (status != DPU_OK && !g_msg.empty()) out += ": " + g_msg;
    return out.c_str();
}

extern "C" dpu_error_t dpu_alloc(uint32_t nr_dpus, const char *profile, struct dpu_set_t *dpu_set)
{
    (void)profile;
    if (!dpu_set) return fail(DPU_ERR_INVALID_DPU_SET, "dpu_alloc: NULL set");
    if (nr_dpus == DPU_ALLOCATE_ALL) {  // "every DPU of the machine": AIM_NR_DPUS or one virtual DPU per 1 Mi pairs is the caller's business
        const char *v = getenv("AIM_NR_DPUS");
        nr_dpus = v && atoi(v) > 0 ? (uint32_t)atoi(v) : 64u;
    }
    if (nr_dpus == 0) return fail(DPU_ERR_ALLOCATION, "dpu_alloc: 0 DPUs requested");
    aim_dpu_system *s = new (std::nothrow) aim_dpu_system();
    if (!s) return fail(DPU_ERR_ALLOCATION, "dpu_alloc: out of memory");
    s->dpu.resize(nr_dpus);
    dpu_set->sys = s;
    dpu_set->dpu = -1;
    return DPU_OK;
}

extern "C" dpu_error_t dpu_free(struct dpu_set_t set)
{
    if (!set.sys) return fail(DPU_ERR_INVALID_DPU_SET, "dpu_free: NULL set");
    for (auto &im : set.sys->dpu) free(im.mem);
    free_gather(set.sys);
    delete set.sys;
    aim_shutdown();
    return DPU_OK;
}

extern "C" dpu_error_t dpu_get_nr_dpus(struct dpu_set_t set, uint32_t *nr_dpus)
{
    if (!set.sys || !nr_dpus) return fail(DPU_ERR_INVALID_DPU_SET, "dpu_get_nr_dpus: NULL argument");
    *nr_dpus = set.dpu >= 0 ? 1u : (uint32_t)set.sys->dpu.size();
    return DPU_OK;
}

extern "C" struct aim_dpu_iterator aim_dpu_iterator_from(struct dpu_set_t *set)
{
    struct aim_dpu_iterator it;
    it.sys = set ? set->sys : nullptr;
    it.count = (set && set->sys && set->dpu >= 0) ? (uint32_t)set->dpu : 0u;
    it.total = !(set && set->sys) ? 0u : set->dpu >= 0 ? (uint32_t)set->dpu + 1u : (uint32_t)set->sys->dpu.size();
    return it;
}

extern "C" struct dpu_set_t aim_dpu_iterator_at(const struct aim_dpu_iterator *it)
{
    struct dpu_set_t one;
    one.sys = it->sys;
    one.dpu = (int32_t)std::min(it->count, it->total ? it->total - 1 : 0u);
    return one;
}

// The program being replaced and its compile-time knobs -> aim_params (include/aim_b200.h).
extern "C" dpu_error_t aim_dpu_load(struct dpu_set_t set, const char *binary_path, struct dpu_program_t **program,
                                    const struct aim_dpu_knobs *k)
{
    if (program) *program = nullptr;
    if (!set.sys) return fail(DPU_ERR_INVALID_DPU_SET, "dpu_load: NULL set");
    if (!binary_path || !k || k->struct_bytes != (int32_t)sizeof(aim_dpu_knobs)) return fail(DPU_ERR_INTERNAL, "dpu_load: knob block mismatch (dpu.h and libaim_dpu.so differ)");
    std::string base = binary_path;
    const size_t slash = base.find_last_of('/');
    if (slash != std::string::npos) base = base.substr(slash + 1);
    aim_params p;
    memset(&p, 0, sizeof p);
    if (base.find("wfa") != std::string::npos) p.algo = AIM_ALGO_WFA;
    else if (base.find("swg") != std::string::npos) p.algo = AIM_ALGO_SWG;
    else if (base.find("nw") != std::string::npos) p.algo = AIM_ALGO_NW;
    else return fail(DPU_ERR_ELF_NO_SUCH_FILE, "dpu_load: '%s' names no program this library replaces (wfa_dpu, nw_dpu, swg_dpu)", binary_path);
    auto val = [](int32_t v, int32_t dflt) { return v == AIM_DPU_UNSET ? dflt : v; };
    p.match = val(k->match, 0);
    p.mismatch = val(k->mismatch, 3);
    if (p.algo == AIM_ALGO_NW) {
        const int32_t gi = val(k->gap_i, 4), gd = val(k->gap_d, gi);
        if (gi != gd) return fail(DPU_ERR_INTERNAL, "dpu_load: GAP_I != GAP_D is not served (the run script sets both from -g)");
        p.gap_open = gi;
        p.gap_ext = 1;
    } else {
        p.gap_open = val(k->gap_o, 4);
        p.gap_ext = val(k->gap_e, 1);
    }
    p.max_score = val(k->max_score, 0);
    p.read_size = val(k->read_size, 0);
    p.backtrace = k->backtrace;
    p.reduce = p.algo == AIM_ALGO_WFA ? k->reduce : 0;
    p.variant = (p.algo == AIM_ALGO_SWG && k->swg_w8) ? 1 : 0;
    if (p.read_size <= 0 || p.read_size % 8) return fail(DPU_ERR_INTERNAL, "dpu_load: READ_SIZE %d must be a positive multiple of 8 (host.c:225-227 asserts it)", p.read_size);
    if ((k->request_bytes != 8 && k->request_bytes != 16) || (k->result_bytes != 24 && k->result_bytes != 32) || k->params_bytes != 32)
        return fail(DPU_ERR_INTERNAL, "dpu_load: unknown wire layout (request_t %d B, result_t %d B, DPUParams %d B)", k->request_bytes, k->result_bytes, k->params_bytes);
    const char *ng = getenv("AIM_NGPUS");
    p.ngpus = !ng ? 1 : (!strcmp(ng, "all") ? std::max(1, aim_device_count()) : std::max(1, atoi(ng)));
    const char *dv = getenv("AIM_DEVICE");
    p.device = dv ? atoi(dv) : 0;
    set.sys->params = p;
    set.sys->knobs = *k;
    set.sys->loaded = true;
    return DPU_OK;
}

extern "C" dpu_error_t dpu_prepare_xfer(struct dpu_set_t set, void *buffer)
{
    if (!set.sys) return fail(DPU_ERR_INVALID_DPU_SET, "dpu_prepare_xfer: NULL set");
    if (set.dpu >= 0) {
        if ((size_t)set.dpu >= set.sys->dpu.size()) return fail(DPU_ERR_INVALID_DPU_SET, "dpu_prepare_xfer: no such DPU");
        set.sys->dpu[(size_t)set.dpu].prepared = buffer;
    } else {
        for (auto &im : set.sys->dpu) im.prepared = buffer;
    }
    return DPU_OK;
}

extern "C" dpu_error_t dpu_push_xfer(struct dpu_set_t set, dpu_xfer_t xfer, const char *symbol_name, uint32_t symbol_offset,
                                     size_t length, dpu_xfer_flags_t flags)
{
    if (!set.sys) return fail(DPU_ERR_INVALID_DPU_SET, "dpu_push_xfer: NULL set");
    if (!symbol_name || strcmp(symbol_name, DPU_MRAM_HEAP_POINTER_NAME) != 0)
        return fail(DPU_ERR_UNKNOWN_SYMBOL, "dpu_push_xfer: only %s is served (symbol '%s')", DPU_MRAM_HEAP_POINTER_NAME, symbol_name ? symbol_name : "(null)");
    if ((symbol_offset % 8) || (length % 8)) return fail(DPU_ERR_INVALID_MRAM_ACCESS, "dpu_push_xfer: MRAM transfers are 8-byte aligned");
    aim_dpu_system *s = set.sys;
    bool oom = false;
    const size_t end = (size_t)symbol_offset + length;
    for_each_dpu(s, set.dpu, length, [&](size_t d) {
        aim_dpu_image &im = s->dpu[d];
        if (!im.prepared) return;
        if (xfer == DPU_XFER_TO_DPU) {
            if (!image_reserve(im, end)) { oom = true; return; }
            memcpy(im.mem + symbol_offset, im.prepared, length);
        } else {
            const size_t have = im.size > symbol_offset ? std::min(length, im.size - symbol_offset) : 0;
            if (have) memcpy(im.prepared, im.mem + symbol_offset, have);
            if (have < length) memset((uint8_t *)im.prepared + have, 0, length - have);
        }
        if (!(flags & DPU_XFER_NO_RESET)) im.prepared = nullptr;
    });
    return oom ? fail(DPU_ERR_ALLOCATION, "dpu_push_xfer: out of host memory for the MRAM images") : DPU_OK;
}

extern "C" dpu_error_t dpu_broadcast_to(struct dpu_set_t set, const char *symbol_name, uint32_t symbol_offset, const void *src,
                                        size_t length, dpu_xfer_flags_t flags)
{
    dpu_error_t e = dpu_prepare_xfer(set, const_cast<void *>(src));
    return e != DPU_OK ? e : dpu_push_xfer(set, DPU_XFER_TO_DPU, symbol_name, symbol_offset, length, flags);
}

extern "C" dpu_error_t dpu_launch(struct dpu_set_t set, dpu_launch_policy_t policy)
{
    (void)policy;  // the alignment runs to completion either way; dpu_sync has nothing left to wait for
    aim_dpu_system *s = set.sys;
    if (!s) return fail(DPU_ERR_INVALID_DPU_SET, "dpu_launch: NULL set");
    if (!s->loaded) return fail(DPU_ERR_NO_PROGRAM_LOADED, "dpu_launch: dpu_load was not called");
    const aim_params &p = s->params;
    const size_t rs = (size_t)p.read_size;
    const size_t req_b = (size_t)s->knobs.request_bytes, res_b = (size_t)s->knobs.result_bytes;
    const bool bt = p.backtrace != 0;
    const size_t nd = s->dpu.size();
    const size_t d0 = set.dpu >= 0 ? (size_t)set.dpu : 0, d1 = set.dpu >= 0 ? (size_t)set.dpu + 1 : nd;

    // ---- every image's table (the DPU's main() reads it from heap offset 0: wfa.c:418-421) ----
    std::vector<WireParams> wp(nd);
    std::vector<size_t> first(nd + 1, 0);
    for (size_t d = d0; d < d1; ++d) {
        const aim_dpu_image &im = s->dpu[d];
        WireParams w{};
        if (im.size >= sizeof w) memcpy(&w, im.mem, sizeof w);
        const size_t n = w.dpuNumReads;
        if (n) {
            if ((size_t)w.dpuRequests_m + n * req_b > im.size || (size_t)w.dpuPatterns_m + n * rs > im.size || (size_t)w.dpuTexts_m + n * rs > im.size)
                return fail(DPU_ERR_INVALID_MRAM_ACCESS, "dpu_launch: DPU %zu's table points outside what was pushed (READ_SIZE %d, %zu reads)", d, p.read_size, n);
        }
        wp[d] = w;
        first[d + 1] = first[d] + n;
    }
    for (size_t d = d1; d < nd; ++d) first[d + 1] = first[d];
    const size_t total = first[nd];
    if (total > 0xffffffffull) return fail(DPU_ERR_INTERNAL, "dpu_launch: more than 2^32 - 1 pairs");
    if (total == 0) return DPU_OK;
    if (aim_device_count() <= 0) return fail(DPU_ERR_DRIVER, "no CUDA device: the B200 library is the DPU program and has no CPU fallback");

    if (total > s->cap_pairs || rs != s->cap_rs || (bt && !s->cap_ops)) {
        free_gather(s);
        s->plen = (int32_t *)aim_host_alloc(total * 4);
        s->tlen = (int32_t *)aim_host_alloc(total * 4);
        s->patterns = (char *)aim_host_alloc(total * rs);
        s->texts = (char *)aim_host_alloc(total * rs);
        s->results = (aim_result *)aim_host_alloc(total * sizeof(aim_result));
        s->ops = bt ? (char *)aim_host_alloc(total * 2 * rs) : nullptr;
        if (!s->plen || !s->tlen || !s->patterns || !s->texts || !s->results || (bt && !s->ops)) {
            free_gather(s);
            return fail(DPU_ERR_ALLOCATION, "dpu_launch: %s", aim_last_error());
        }
        s->cap_pairs = total;
        s->cap_rs = rs;
        s->cap_ops = bt;
    }

    // ---- gather: request_t (WFA: int16 plen, int16 tlen, u32 idx; NW/SWG: int plen, int tlen, int pad, u32 idx) + rows ----
    std::vector<uint32_t> idx(total);
    for_each_dpu(s, set.dpu, (size_t)(total / std::max<size_t>(1, d1 - d0)) * 2 * rs, [&](size_t d) {
        const WireParams &w = wp[d];
        const aim_dpu_image &im = s->dpu[d];
        const size_t n = w.dpuNumReads, o = first[d];
        for (size_t i = 0; i < n; ++i) {
            const uint8_t *r = im.mem + w.dpuRequests_m + i * req_b;
            if (req_b == 8) {
                int16_t a, b;
                memcpy(&a, r, 2); memcpy(&b, r + 2, 2); memcpy(&idx[o + i], r + 4, 4);
                s->plen[o + i] = a; s->tlen[o + i] = b;
            } else {
                memcpy(&s->plen[o + i], r, 4); memcpy(&s->tlen[o + i], r + 4, 4); memcpy(&idx[o + i], r + 12, 4);
            }
        }
        if (n) {
            memcpy(s->patterns + o * rs, im.mem + w.dpuPatterns_m, n * rs);
            memcpy(s->texts + o * rs, im.mem + w.dpuTexts_m, n * rs);
        }
    });

    int rc = aim_align_batch(&p, (uint32_t)total, 0, s->plen, s->tlen, s->patterns, s->texts, s->results, s->ops, s->phase_ms);
    if (rc != AIM_OK) return fail(rc == AIM_ERR_NO_DEVICE ? DPU_ERR_DRIVER : DPU_ERR_DPU_FAULT, "%s: %s", aim_strerror(rc), aim_last_error());
    // what makes a DPU fault in the reference: a backtrace dead end (wfa_backtracing.c:343-344, swg.c:131-133: message + exit) and
    // the history store running out of MRAM (dpu_allocator_mram.c:6-10)
    for (size_t i = 0; i < total; ++i) {
        if (s->results[i].status == AIM_STATUS_BACKTRACE)
            return fail(DPU_ERR_DPU_FAULT, p.algo == AIM_ALGO_SWG ? "SWG backtrace. No backtrace operation found (pair %u)" : "Backtrace error: No link found during backtrace (pair %u)", idx[i]);
        if (s->results[i].status == AIM_STATUS_ARENA) return fail(DPU_ERR_DPU_FAULT, "Out of memory MRAM (pair %u)", idx[i]);
    }

    // ---- scatter: result_t {max_operations, begin_offset, end_offset, score, [u64 cycles | int pad], idx} and the op rows ----
    bool oom = false;
    for_each_dpu(s, set.dpu, (size_t)(total / std::max<size_t>(1, d1 - d0)) * 2 * rs, [&](size_t d) {
        const WireParams &w = wp[d];
        aim_dpu_image &im = s->dpu[d];
        const size_t n = w.dpuNumReads, o = first[d];
        if (!n) return;
        size_t need = (size_t)w.dpuResults_m + n * res_b;
        if (bt) need = std::max(need, (size_t)w.dpuOperations_m + n * 2 * rs);
        if (!image_reserve(im, need)) { oom = true; return; }
        for (size_t i = 0; i < n; ++i) {
            uint8_t *r = im.mem + w.dpuResults_m + i * res_b;
            const aim_result &g = s->results[o + i];
            memset(r, 0, res_b);
            memcpy(r, &g.max_operations, 4); memcpy(r + 4, &g.begin_offset, 4); memcpy(r + 8, &g.end_offset, 4); memcpy(r + 12, &g.score, 4);
            memcpy(r + (res_b == 32 ? 24 : 20), &idx[o + i], 4);
        }
        if (bt) memcpy(im.mem + w.dpuOperations_m, s->ops + o * 2 * rs, n * 2 * rs);
    });
    return oom ? fail(DPU_ERR_ALLOCATION, "dpu_launch: out of host memory for the MRAM images") : DPU_OK;
}

extern "C" dpu_error_t dpu_sync(struct dpu_set_t set) { return set.sys ? DPU_OK : fail(DPU_ERR_INVALID_DPU_SET, "dpu_sync: NULL set"); }

extern "C" dpu_error_t dpu_log_read(struct dpu_set_t set, FILE *stream)
{
    (void)stream;
    return set.sys ? DPU_OK : fail(DPU_ERR_INVALID_DPU_SET, "dpu_log_read: NULL set");
}

extern "C" void aim_dpu_last_phases(struct dpu_set_t set, double phase_ms[3])
{
    for (int k = 0; k < 3; ++k) phase_ms[k] = set.sys ? set.sys->phase_ms[k] : 0.0;
}
