/* TEST INFRASTRUCTURE (oracle/): stand-in for UPMEM <barrier.h> (unused by the sources). */
#ifndef AIM_ORACLE_SHIM_BARRIER_H
#define AIM_ORACLE_SHIM_BARRIER_H
#endif
