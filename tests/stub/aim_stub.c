/* TEST INFRASTRUCTURE: stands in for libaim_b200.so's GPU entry points so that the HOST-SIDE plumbing of
 * libaim_dpu.so (include/dpu.h: MRAM images, gather, scatter, wire layouts) can be exercised on a machine without a
 * GPU.  Loaded with LD_PRELOAD by tests/test_upmem_adapter.py only; aim_align_batch here is the CPU oracle
 * (oracle/aim_oracle.c), which the product never links.  Not built by the Makefile, not shipped. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "aim_b200.h"

typedef struct { int32_t algo, match, mismatch, gap_open, gap_ext, max_score, read_size, backtrace, reduce, variant; } orc_params;
typedef struct { int32_t max_operations, begin_offset, end_offset, score, status; } orc_result;
int orc_align_batch(const orc_params *p, uint32_t n, const int32_t *plen, const int32_t *tlen, const char *patterns,
                    const char *texts, orc_result *results, char *ops, int nthreads);

int aim_device_count(void) { return 1; }
void *aim_host_alloc(size_t bytes) { return malloc(bytes ? bytes : 1); }
void aim_host_free(void *p) { free(p); }
void aim_shutdown(void) {}
const char *aim_last_error(void) { return "stub"; }
const char *aim_strerror(int code) { (void)code; return "stub error"; }

int aim_align_batch(const aim_params *params, uint32_t n, uint32_t idx_base, const int32_t *plen, const int32_t *tlen,
                    const char *patterns, const char *texts, aim_result *results, char *ops, double phase_ms[3])
{
    orc_params p = {params->algo, params->match, params->mismatch, params->gap_open, params->gap_ext, params->max_score,
                    params->read_size, params->backtrace, params->reduce, params->variant};
    orc_result *r = (orc_result *)calloc(n ? n : 1, sizeof *r);
    int rc = orc_align_batch(&p, n, plen, tlen, patterns, texts, r, ops, 4);
    for (uint32_t i = 0; i < n; ++i) {
        results[i].max_operations = r[i].max_operations; results[i].begin_offset = r[i].begin_offset;
        results[i].end_offset = r[i].end_offset; results[i].score = r[i].score; results[i].status = r[i].status;
        results[i].idx = idx_base + i;
    }
    free(r);
    if (phase_ms) phase_ms[0] = phase_ms[1] = phase_ms[2] = 0.0;
    return rc == 0 ? AIM_OK : AIM_ERR_ARG;
}
