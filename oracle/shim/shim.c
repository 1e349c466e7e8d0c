/* TEST INFRASTRUCTURE (oracle/): native runtime behind the UPMEM stand-in headers.
 * One simulated DPU = one 64 MiB MRAM image; dpu_launch runs dpu_main() once per
 * tasklet id, DPUs spread over AIM_SHIM_THREADS host threads (default 1).
 * AIM_SHIM_NR_DPUS overrides the DPU count returned by dpu_get_nr_dpus so one
 * binary serves any thread count (host.c only uses the compile-time NR_DPUS in
 * its "N <= NR_DPUS" check, WFA/DPU-MRAM/host/host.c:180). */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "dpu.h"

#define SHIM_MRAM_BYTES (64u << 20)
#ifndef SHIM_WRAM_BYTES
#define SHIM_WRAM_BYTES (64u << 20) /* not the 64 KB of real WRAM: the sources' own 62000-B guard still applies */
#endif

uint32_t shim_nr_dpus = 0;
__thread uint32_t shim_tasklet_id = 0;
__thread uint8_t *shim_mram = NULL;
static __thread uint8_t *wram_heap = NULL;
static __thread size_t wram_used = 0;

static uint8_t **mram_images;
static void **xfer_ptr;
extern int dpu_main(void);

void mem_reset(void) { wram_used = 0; }
void *mem_alloc(size_t size)
{
    if (!wram_heap) wram_heap = (uint8_t *)malloc(SHIM_WRAM_BYTES);
    size = (size + 7) & ~(size_t)7;
    if (wram_used + size > SHIM_WRAM_BYTES) { fprintf(stderr, "shim: WRAM heap exhausted\n"); exit(1); }
    void *p = wram_heap + wram_used;
    wram_used += size;
    return p;
}

dpu_error_t dpu_alloc(uint32_t nr, const char *profile, struct dpu_set_t *set)
{
    (void)profile;
    const char *env = getenv("AIM_SHIM_NR_DPUS");
    shim_nr_dpus = env ? (uint32_t)atoi(env) : nr;
    if (shim_nr_dpus == 0) shim_nr_dpus = 1;
    mram_images = (uint8_t **)calloc(shim_nr_dpus, sizeof(*mram_images));
    xfer_ptr = (void **)calloc(shim_nr_dpus, sizeof(*xfer_ptr));
    for (uint32_t i = 0; i < shim_nr_dpus; ++i) {
        mram_images[i] = (uint8_t *)calloc(SHIM_MRAM_BYTES, 1);
        if (!mram_images[i]) { fprintf(stderr, "shim: out of host memory for MRAM images\n"); exit(1); }
    }
    set->idx = -1;
    return DPU_OK;
}
dpu_error_t dpu_load(struct dpu_set_t set, const char *binary, void *unused) { (void)set; (void)binary; (void)unused; return DPU_OK; }
dpu_error_t dpu_get_nr_dpus(struct dpu_set_t set, uint32_t *nr) { (void)set; *nr = shim_nr_dpus; return DPU_OK; }
dpu_error_t dpu_prepare_xfer(struct dpu_set_t dpu, void *buffer) { xfer_ptr[dpu.idx] = buffer; return DPU_OK; }
dpu_error_t dpu_push_xfer(struct dpu_set_t set, dpu_xfer_t dir, const char *symbol, uint32_t offset, size_t length, int flags)
{
    (void)set; (void)symbol; (void)flags;
    for (uint32_t i = 0; i < shim_nr_dpus; ++i) {
        if (dir == DPU_XFER_TO_DPU) memcpy(mram_images[i] + offset, xfer_ptr[i], length);
        else memcpy(xfer_ptr[i], mram_images[i] + offset, length);
    }
    return DPU_OK;
}

typedef struct { uint32_t first, step; } worker_t;
static void *worker(void *arg)
{
    worker_t *w = (worker_t *)arg;
    for (uint32_t d = w->first; d < shim_nr_dpus; d += w->step) {
        shim_mram = mram_images[d];
        for (uint32_t t = 0; t < NR_TASKLETS; ++t) { shim_tasklet_id = t; dpu_main(); }
    }
    free(wram_heap); wram_heap = NULL;
    return NULL;
}
dpu_error_t dpu_launch(struct dpu_set_t set, int policy)
{
    (void)set; (void)policy;
    const char *env = getenv("AIM_SHIM_THREADS");
    uint32_t nt = env ? (uint32_t)atoi(env) : 1;
    if (nt < 1) nt = 1;
    if (nt > shim_nr_dpus) nt = shim_nr_dpus;
    if (nt == 1) { worker_t w = {0, 1}; worker(&w); return DPU_OK; }
    pthread_t *th = (pthread_t *)malloc(nt * sizeof(*th));
    worker_t *ws = (worker_t *)malloc(nt * sizeof(*ws));
    for (uint32_t i = 0; i < nt; ++i) { ws[i].first = i; ws[i].step = nt; pthread_create(&th[i], NULL, worker, &ws[i]); }
    for (uint32_t i = 0; i < nt; ++i) pthread_join(th[i], NULL);
    free(th); free(ws);
    return DPU_OK;
}
dpu_error_t dpu_log_read(struct dpu_set_t dpu, FILE *f) { (void)dpu; (void)f; return DPU_OK; }
dpu_error_t dpu_free(struct dpu_set_t set)
{
    (void)set;
    for (uint32_t i = 0; i < shim_nr_dpus; ++i) free(mram_images[i]);
    free(mram_images); free(xfer_ptr);
    return DPU_OK;
}
