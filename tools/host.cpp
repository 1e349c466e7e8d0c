// `host <pairs-file> <out-file> <N>` — drop-in for the six reference host programs
// ({NW,SWG,WFA}/DPU-{WRAM,MRAM}/host/host.c; citations below are WFA/DPU-MRAM/host/host.c).
// Same argv, same stdout lines, same pair-file parsing, same output bytes.  The knobs the
// reference bakes in with -D at compile time arrive through the environment, under the same
// names (the run-*-pim-*.py wrappers set them):
//   AIM_ALGO=nw|swg|wfa|genasm_dc|genasm_filter  MAX_SCORE READ_SIZE MATCH MISMATCH GAP_O GAP_E (GAP_I/GAP_D for NW)
//   BACKTRACE=0|1  REDUCE=0|1  NR_DPUS (only feeds the pairs-to-process rule, host.c:191)
//   NR_TASKLETS WRAM_SEGMENT (accepted, ignored)  AIM_NGPUS  AIM_DEVICE  AIM_VARIANT=wram|mram
//   AIM_HOST_PATH=stream|batch  stream (default): aim_align_file - the pair file is parsed and the output text formatted
//                               ON THE GPU, host threads only read() and write();  batch: get_reads on the host
//                               (aim_read_pairs) + aim_align_batch[_cigars] + the host printer
#include <sys/time.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "aim_b200.h"

static long env_int(const char *name, long dflt)
{
    const char *v = getenv(name);
    return (v && *v) ? atol(v) : dflt;
}

static double now_s()
{
    struct timeval tv;
    gettimeofday(&tv, nullptr);  // timer.h:34-45 uses gettimeofday as well
    return (double)tv.tv_sec + (double)tv.tv_usec * 1e-6;
}

int main(int argc, char *argv[])
{
    if (argc != 4) {  // host.c:150-154
        printf("wrong number of arguments\n");
        exit(1);
    }
    const char *in = argv[1];
    const char *out = argv[2];
    uint32_t total_nb_reads = (uint32_t)atoi(argv[3]);

    const char *algo_s = getenv("AIM_ALGO");
    std::string algo = algo_s ? algo_s : "wfa";
    const char *variant_s = getenv("AIM_VARIANT");
    std::string variant = variant_s ? variant_s : "mram";
    aim_params p;
    memset(&p, 0, sizeof(p));
    p.algo = algo == "nw" ? AIM_ALGO_NW : algo == "swg" ? AIM_ALGO_SWG : algo == "genasm_dc" ? AIM_ALGO_GENASM_DC
             : algo == "genasm_filter" ? AIM_ALGO_GENASM_FILTER : AIM_ALGO_WFA;
    // the aim-genasm hosts (aim-genasm/GenASM/DPU-*-{DC,filter}/host/host.c) share this skeleton; they differ in the defaults
    // (common.h:50-56: MAX_SCORE 5, READ_SIZE 120), in the output line and in having no BACKTRACE knob
    const bool genasm = p.algo == AIM_ALGO_GENASM_DC || p.algo == AIM_ALGO_GENASM_FILTER;
    // defaults of */common/common.h
    p.match = (int32_t)env_int("MATCH", 0);
    p.mismatch = (int32_t)env_int("MISMATCH", 3);
    p.gap_open = (int32_t)env_int(p.algo == AIM_ALGO_NW ? "GAP_I" : "GAP_O", 4);
    p.gap_ext = (int32_t)env_int("GAP_E", 1);
    p.max_score = (int32_t)env_int("MAX_SCORE", genasm ? 5 : p.algo == AIM_ALGO_WFA ? 250 : p.algo == AIM_ALGO_SWG ? 400 : 40);
    p.read_size = (int32_t)env_int("READ_SIZE", genasm ? 120 : p.algo == AIM_ALGO_WFA ? 110 : p.algo == AIM_ALGO_SWG ? 560 : 56);
    p.read_size = (p.read_size + 7) / 8 * 8;
    p.backtrace = (int32_t)env_int("BACKTRACE", 0);
    if (genasm) p.backtrace = p.algo == AIM_ALGO_GENASM_DC;
    p.variant = (genasm && variant == "mram") ? 1 : 0;
    if (p.algo == AIM_ALGO_SWG && variant == "wram") p.variant = 1;  // SWG/DPU-WRAM: int8 cells when MAX_SCORE < 127
    p.reduce = (int32_t)env_int("REDUCE", 0);
    p.ngpus = (int32_t)env_int("AIM_NGPUS", 1);
    p.device = (int32_t)env_int("AIM_DEVICE", 0);
    p.arena_mb = (int32_t)env_int("AIM_ARENA_MB", 0);
    const uint32_t nr_dpus = (uint32_t)env_int("NR_DPUS", 1);

    // host.c:161-184, in the same order; the side file is `dpu-out` (`dpu_out` for NW/DPU-WRAM)
    FILE *input_file = fopen(in, "r");
    FILE *output_file = fopen(out, "w");
    FILE *dpu_file = fopen((p.algo == AIM_ALGO_NW && variant == "wram") ? "dpu_out" : "dpu-out", "w");
    if (input_file == NULL) {
        fprintf(stderr, "Input file '%s' couldn't be opened\n", in);
        exit(1);
    }
    if (output_file == NULL) {
        fprintf(stderr, "Output file '%s' couldn't be opened\n", out);
        exit(1);
    }
    if (total_nb_reads <= 0) {
        fprintf(stderr, "Invalid nb of reads\n");
        exit(1);
    }
    if (total_nb_reads <= nr_dpus) {
        printf("Allocated DPUs more than needed\n");
        exit(1);
    }
    fclose(input_file);
    fclose(output_file);

    printf("Allocated %d DPU(s)\n", (int)nr_dpus);  // host.c:189
    const uint32_t nb_reads_per_dpu = (uint32_t)(((total_nb_reads / nr_dpus) + 7) / 8 * 8);
    printf("NumReads per dpu = %u\n", nb_reads_per_dpu);  // host.c:192

    const char *path_s = getenv("AIM_HOST_PATH");
    if (!(path_s && std::string(path_s) == "batch")) {
        // ---- streaming path: host.c:196-353 as one call ----
        uint64_t done = 0;
        uint32_t status_mask = 0;
        double phase[3] = {0, 0, 0};
        const double t0 = now_s();
        const int rc = aim_align_file(&p, in, out, total_nb_reads, nr_dpus, &done, &status_mask, phase, nullptr);
        const double wall_ms = (now_s() - t0) * 1e3;
        if (rc == AIM_ERR_LENGTH) {  // host.c:119-123
            printf("READ LENGTH less than length of the input reads");
            exit(0);
        }
        if (rc != AIM_OK) {
            fprintf(stderr, "aim_b200: %s: %s\n", aim_strerror(rc), aim_last_error());
            exit(1);
        }
        const bool nw = p.algo == AIM_ALGO_NW;
        printf("Copying data to DPU\n");
        printf("CPU-DPU: %f ms\n", phase[0]);
        printf("Run program on DPU(s)\n");
        printf((nw && variant == "wram") ? "DPU Kernel Time: %f ms\n" : "DPU Kernel: %f ms\n", phase[1]);
        printf("Retrieve results\n");
        printf(nw ? "DPU-CPU Time: %f ms\n" : "DPU-CPU: %f ms\n", phase[2]);
        if (getenv("AIM_VERBOSE")) printf("B200 pipeline wall: %f ms (%llu pairs, %d GPU(s), parsed and printed on the GPU)\n", wall_ms, (unsigned long long)done, p.ngpus);
        // the reference exits the whole process on a backtrace dead end (swg.c:131-133, wfa_backtracing.c:343-344) or when the
        // history store runs out of MRAM (dpu_allocator_mram.c:6-10); aim_align_file left the output file empty
        if (status_mask & (1u << AIM_STATUS_BACKTRACE)) {
            printf(p.algo == AIM_ALGO_SWG ? "SWG backtrace. No backtrace operation found" : "Backtrace error: No link found during backtrace\n");
            exit(1);
        }
        if (status_mask & (1u << AIM_STATUS_ARENA)) {
            printf("Out of memory MRAM\n");
            exit(-1);
        }
        if (dpu_file) fclose(dpu_file);
        aim_shutdown();
        return 0;
    }

    const uint64_t cap = (uint64_t)nb_reads_per_dpu * nr_dpus;
    int64_t in_file = aim_count_pairs(in);
    if (in_file < 0) {
        fprintf(stderr, "Input file '%s' couldn't be opened\n", in);
        exit(1);
    }
    const uint32_t want = (uint32_t)((uint64_t)in_file < cap ? (uint64_t)in_file : cap);
    const size_t rs = (size_t)p.read_size;
    const size_t alloc_n = want ? want : 1;
    int32_t *plen = (int32_t *)aim_host_alloc(alloc_n * sizeof(int32_t));
    int32_t *tlen = (int32_t *)aim_host_alloc(alloc_n * sizeof(int32_t));
    char *patterns = (char *)aim_host_alloc(alloc_n * rs);
    char *texts = (char *)aim_host_alloc(alloc_n * rs);
    aim_result *results = (aim_result *)aim_host_alloc(alloc_n * sizeof(aim_result));
    // (the op rows are allocated only if they are fetched: GenASM-DC's CIGAR strings, or the fallback below)
    char *ops = (p.backtrace && genasm) ? (char *)aim_host_alloc(alloc_n * 2 * rs) : nullptr;
    if (!plen || !tlen || !patterns || !texts || !results || (p.backtrace && genasm && !ops)) {
        fprintf(stderr, "aim_b200: %s\n", aim_last_error());
        exit(1);
    }
    int64_t n = aim_read_pairs(in, want, p.read_size, plen, tlen, patterns, texts);
    if (n == AIM_ERR_LENGTH) {  // host.c:119-123
        printf("READ LENGTH less than length of the input reads");
        exit(0);
    }
    if (n < 0) {
        fprintf(stderr, "aim_b200: %s\n", aim_last_error());
        exit(1);
    }

    double phase[3] = {0, 0, 0};
    double t0 = now_s();
    int rc;
    // With a CIGAR to print, the text edit_cigar_print (host.c:69-89) would write is built on the GPU and only that crosses
    // PCIe (aim_align_batch_cigars); the op rows are fetched instead when some CIGAR does not fit the row (AIM_CIGAR_ROWS=0 forces that).
    char *cigars = nullptr;
    int32_t cigar_pitch = 0;
    if (p.backtrace && !genasm && env_int("AIM_CIGAR_ROWS", 1) != 0) {
        cigar_pitch = (int32_t)std::min<size_t>(2 * rs, 128) / 16 * 16;
        if (cigar_pitch >= 16) cigars = (char *)aim_host_alloc(alloc_n * (size_t)cigar_pitch);
    }
    if (cigars) {
        rc = aim_align_batch_cigars(&p, (uint32_t)n, 0, plen, tlen, patterns, texts, results, cigars, cigar_pitch, phase);
        bool overflow = false;
        for (int64_t i = 0; rc == AIM_OK && i < n; ++i) overflow |= results[i].status == AIM_STATUS_CIGAR_OVERFLOW;
        if (rc != AIM_OK || overflow) { aim_host_free(cigars); cigars = nullptr; }
    }
    if (!cigars) {
        if (p.backtrace && !ops) ops = (char *)aim_host_alloc(alloc_n * 2 * rs);
        if (p.backtrace && !ops) { fprintf(stderr, "aim_b200: %s\n", aim_last_error()); exit(1); }
        rc = aim_align_batch(&p, (uint32_t)n, 0, plen, tlen, patterns, texts, results, ops, phase);
    }
    double wall_ms = (now_s() - t0) * 1e3;
    if (rc != AIM_OK) {
        fprintf(stderr, "aim_b200: %s: %s\n", aim_strerror(rc), aim_last_error());
        exit(1);
    }
    // host.c:244-330: the three phases, with the label spelling of the program being replaced
    const bool nw = p.algo == AIM_ALGO_NW;
    printf("Copying data to DPU\n");
    printf("CPU-DPU: %f ms\n", phase[0]);
    printf("Run program on DPU(s)\n");
    printf((nw && variant == "wram") ? "DPU Kernel Time: %f ms\n" : "DPU Kernel: %f ms\n", phase[1]);
    printf("Retrieve results\n");
    printf(nw ? "DPU-CPU Time: %f ms\n" : "DPU-CPU: %f ms\n", phase[2]);
    if (getenv("AIM_VERBOSE")) printf("B200 pipeline wall: %f ms (%u pairs, %d GPU(s))\n", wall_ms, (unsigned)n, p.ngpus);

    // the reference exits the whole process on a backtrace dead end (swg.c:131-133, wfa_backtracing.c:343-344)
    for (int64_t i = 0; i < n; ++i) {
        if (results[i].status == AIM_STATUS_BACKTRACE) {
            printf(p.algo == AIM_ALGO_SWG ? "SWG backtrace. No backtrace operation found" : "Backtrace error: No link found during backtrace\n");
            exit(1);
        }
        if (results[i].status == AIM_STATUS_ARENA) {
            printf("Out of memory MRAM\n");  // dpu_allocator_mram.c:6-10
            exit(-1);
        }
    }
    rc = genasm ? aim_write_results_genasm(out, (uint32_t)n, p.read_size, p.algo == AIM_ALGO_GENASM_DC, results, ops)
         : cigars ? aim_write_results_packed(out, (uint32_t)n, results, cigars, cigar_pitch)
                  : aim_write_results(out, (uint32_t)n, p.read_size, p.backtrace, results, ops);
    if (rc != AIM_OK) {
        fprintf(stderr, "Output file '%s' couldn't be opened\n", out);
        exit(1);
    }
    if (dpu_file) fclose(dpu_file);
    aim_host_free(plen); aim_host_free(tlen); aim_host_free(patterns); aim_host_free(texts);
    aim_host_free(results); aim_host_free(ops); aim_host_free(cigars);
    aim_shutdown();
    return 0;
}
