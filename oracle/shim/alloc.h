/* TEST INFRASTRUCTURE (oracle/): stand-in for UPMEM <alloc.h> (WRAM heap). */
#ifndef AIM_ORACLE_SHIM_ALLOC_H
#define AIM_ORACLE_SHIM_ALLOC_H
#include <stddef.h>
void mem_reset(void);
void *mem_alloc(size_t size);
#endif
