/* TEST INFRASTRUCTURE (oracle/): stand-in for the host side of the UPMEM SDK
 * (v2021.3.0, not vendored by safaad/aim) so that the reference's own host.c and
 * DPU sources compile natively with gcc and act as the parity oracle.
 * Nothing here is arithmetic; it is transfer/launch plumbing only.
 * Interface modelled: dpu_alloc/dpu_load/dpu_get_nr_dpus/DPU_FOREACH/
 * dpu_prepare_xfer/dpu_push_xfer/dpu_launch/dpu_log_read/dpu_free as called from
 * WFA/DPU-MRAM/host/host.c:186-372 (same calls in the other five hosts). */
#ifndef AIM_ORACLE_SHIM_DPU_H
#define AIM_ORACLE_SHIM_DPU_H
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

typedef int dpu_error_t;
#define DPU_OK 0
struct dpu_set_t { int idx; };

#define DPU_ASSERT(stmt) do { if ((stmt) != DPU_OK) { fprintf(stderr, "shim: DPU_ASSERT failed: %s\n", #stmt); exit(1); } } while (0)
#define DPU_MRAM_HEAP_POINTER_NAME "__sys_used_mram_end"
typedef enum { DPU_XFER_TO_DPU, DPU_XFER_FROM_DPU } dpu_xfer_t;
#define DPU_XFER_DEFAULT 0
#define DPU_SYNCHRONOUS 0

extern uint32_t shim_nr_dpus;
#define DPU_FOREACH(set, d, i) for ((i) = 0, (d).idx = 0; (uint32_t)(i) < shim_nr_dpus; ++(i), (d).idx = (int)(i))

dpu_error_t dpu_alloc(uint32_t nr, const char *profile, struct dpu_set_t *set);
dpu_error_t dpu_load(struct dpu_set_t set, const char *binary, void *unused);
dpu_error_t dpu_get_nr_dpus(struct dpu_set_t set, uint32_t *nr);
dpu_error_t dpu_prepare_xfer(struct dpu_set_t dpu, void *buffer);
dpu_error_t dpu_push_xfer(struct dpu_set_t set, dpu_xfer_t dir, const char *symbol, uint32_t offset, size_t length, int flags);
dpu_error_t dpu_launch(struct dpu_set_t set, int policy);
dpu_error_t dpu_log_read(struct dpu_set_t dpu, FILE *f);
dpu_error_t dpu_free(struct dpu_set_t set);
#endif
