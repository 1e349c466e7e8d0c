/* TEST INFRASTRUCTURE (oracle/): stand-in for UPMEM <mram.h>: MRAM is a 64 MiB
 * host allocation per simulated DPU, mram_read/mram_write are memcpy. */
#ifndef AIM_ORACLE_SHIM_MRAM_H
#define AIM_ORACLE_SHIM_MRAM_H
#include <stdint.h>
#include <string.h>
#include "defs.h"
extern __thread uint8_t *shim_mram;
#define DPU_MRAM_HEAP_POINTER ((void *)0)
static inline void mram_read(const void *from, void *to, unsigned int n) { memcpy(to, shim_mram + (uintptr_t)from, n); }
static inline void mram_write(const void *from, void *to, unsigned int n) { memcpy(shim_mram + (uintptr_t)to, from, n); }
#endif
