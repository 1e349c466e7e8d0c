/* TEST INFRASTRUCTURE (oracle/): stand-in for UPMEM <defs.h> (tasklet id). */
#ifndef AIM_ORACLE_SHIM_DEFS_H
#define AIM_ORACLE_SHIM_DEFS_H
#include <stdint.h>
#define __mram_ptr
#define __host
extern __thread uint32_t shim_tasklet_id;
static inline uint32_t me(void) { return shim_tasklet_id; }
#endif
