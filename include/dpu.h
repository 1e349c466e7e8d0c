/* dpu.h — the slice of the UPMEM host API (SDK v2021.3.0 <dpu.h>, README.md:45) that safaad/aim's six host programs
 * call, served by B200 GPUs through libaim_b200.so.  With this header first on the include path and -laim_dpu on the
 * link line, {NW,SWG,WFA}/DPU-{WRAM,MRAM}/host/host.c compile UNCHANGED and run on a B200 (route C).
 *
 * What each call becomes (call sites: WFA/DPU-MRAM/host/host.c:186-372; the other five hosts are the same):
 *   dpu_alloc(NR_DPUS, NULL, &set)        host.c:186  a set of NR_DPUS virtual DPUs (nothing is allocated on a device yet)
 *   dpu_load(set, DPU_BINARY, NULL)       host.c:187  no binary is loaded: the call records WHICH program is replaced (basename
 *                                                     wfa_dpu / nw_dpu / swg_dpu) and the -D knobs the DPU binary would have been
 *                                                     compiled with - captured by the macro below from the very macros common.h
 *                                                     (<prog>/common/common.h, included before <dpu.h> by every host) defines
 *   dpu_get_nr_dpus(set, &n)              host.c:188
 *   DPU_FOREACH(set, dpu[, i])            host.c:216  iterates the virtual DPUs; i is assigned from the iterator every pass
 *   dpu_prepare_xfer(dpu, ptr)            host.c:249  records the host pointer for that DPU
 *   dpu_push_xfer(set, TO_DPU, DPU_MRAM_HEAP_POINTER_NAME, off, len, ..)  host.c:251-268  copies len bytes of every prepared
 *                                                     buffer into that DPU's MRAM-heap image at off (host memory; the host frees
 *                                                     its buffers right after, host.c:274-279)
 *   dpu_launch(set, DPU_SYNCHRONOUS)      host.c:289  reads every image's DPUParams table (common.h:189-199) and request_t array,
 *                                                     gathers all pairs and aligns them with ONE aim_align_batch() call on the
 *                                                     GPU(s); result_t records and op rows land in the images where the DPU program
 *                                                     would have written them (wfa.c:507-533)
 *   dpu_push_xfer(set, FROM_DPU, ...)     host.c:316-326  copies from the images into the prepared buffers
 *   dpu_log_read(dpu, file)               host.c:361  the DPU programs log nothing on success: writes nothing
 *   dpu_free(set)                         host.c:372
 * Errors: a dpu_error_t != DPU_OK, which DPU_ASSERT turns into message + exit, as the SDK does.  There is no CPU fallback:
 * without a B200, dpu_launch fails with the library's "no CUDA device" error.
 * Environment: AIM_NGPUS (default 1; "all" = every visible GPU), AIM_DEVICE (first ordinal, default 0).
 */
#ifndef AIM_B200_DPU_H
#define AIM_B200_DPU_H

#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum dpu_error_t {
    DPU_OK = 0,
    DPU_ERR_INTERNAL,
    DPU_ERR_SYSTEM,
    DPU_ERR_DRIVER,
    DPU_ERR_ALLOCATION,
    DPU_ERR_INVALID_DPU_SET,
    DPU_ERR_INVALID_SYMBOL_ACCESS,
    DPU_ERR_UNKNOWN_SYMBOL,
    DPU_ERR_INVALID_MRAM_ACCESS,
    DPU_ERR_TRANSFER_ALREADY_SET,
    DPU_ERR_DIFFERENT_DPU_PROGRAMS,
    DPU_ERR_NO_PROGRAM_LOADED,
    DPU_ERR_DPU_FAULT,
    DPU_ERR_ELF_NO_SUCH_FILE
} dpu_error_t;

typedef enum dpu_xfer_t { DPU_XFER_TO_DPU, DPU_XFER_FROM_DPU } dpu_xfer_t;
typedef enum dpu_xfer_flags_t { DPU_XFER_DEFAULT = 0, DPU_XFER_NO_RESET = 1 << 0, DPU_XFER_ASYNC = 1 << 1 } dpu_xfer_flags_t;
typedef enum dpu_launch_policy_t { DPU_ASYNCHRONOUS, DPU_SYNCHRONOUS } dpu_launch_policy_t;

#define DPU_ALLOCATE_ALL ((uint32_t)-1)
#define DPU_MRAM_HEAP_POINTER_NAME "__sys_used_mram_end"

struct aim_dpu_system;   /* opaque: the virtual DPUs of one dpu_alloc */
struct dpu_program_t;    /* opaque, never materialised */
/* Passed by value like the SDK's: `dpu` < 0 = the whole set, otherwise one DPU of it. */
struct dpu_set_t {
    struct aim_dpu_system *sys;
    int32_t dpu;
};

const char *dpu_error_to_string(dpu_error_t status);

#define DPU_ASSERT(statement)                                                                                   \
    do {                                                                                                        \
        dpu_error_t aim_dpu_status__ = (statement);                                                             \
        if (aim_dpu_status__ != DPU_OK) {                                                                       \
            fprintf(stderr, "%s:%d(%s): DPU Error (%s)\n", __FILE__, __LINE__, __func__,                        \
                    dpu_error_to_string(aim_dpu_status__));                                                     \
            exit(EXIT_FAILURE);                                                                                 \
        }                                                                                                       \
    } while (0)

/* DPU_FOREACH(set, dpu) / DPU_FOREACH(set, dpu, i) */
struct aim_dpu_iterator {
    struct aim_dpu_system *sys;
    uint32_t count, total;
};
struct aim_dpu_iterator aim_dpu_iterator_from(struct dpu_set_t *set);
struct dpu_set_t aim_dpu_iterator_at(const struct aim_dpu_iterator *it);
#define AIM_DPU_FOREACH_2(set, one)                                                                             \
    for (struct aim_dpu_iterator aim_dpu_it__ = aim_dpu_iterator_from(&(set));                                  \
         (one) = aim_dpu_iterator_at(&aim_dpu_it__), aim_dpu_it__.count < aim_dpu_it__.total; ++aim_dpu_it__.count)
#define AIM_DPU_FOREACH_3(set, one, i)                                                                          \
    for (struct aim_dpu_iterator aim_dpu_it__ = aim_dpu_iterator_from(&(set));                                  \
         (one) = aim_dpu_iterator_at(&aim_dpu_it__), (i) = aim_dpu_it__.count, aim_dpu_it__.count < aim_dpu_it__.total; \
         ++aim_dpu_it__.count)
#define AIM_DPU_FOREACH_PICK(a, b, c, name, ...) name
#define DPU_FOREACH(...) AIM_DPU_FOREACH_PICK(__VA_ARGS__, AIM_DPU_FOREACH_3, AIM_DPU_FOREACH_2, )(__VA_ARGS__)

/* ---- the -D knobs of the program being replaced, captured where dpu_load() is written ---------------------------
 * The reference compiles host and DPU binary with the same FLAGS (<prog>/Makefile:26-27, run-wfa-pim-mram.py:133-139), so the
 * macros visible in the host translation unit ARE the DPU program's knobs.  AIM_DPU_UNSET marks a macro the program
 * does not have (NW has GAP_I/GAP_D, no GAP_O/GAP_E). */
#define AIM_DPU_UNSET INT32_MIN
struct aim_dpu_knobs {
    int32_t struct_bytes;
    int32_t match, mismatch, gap_o, gap_e, gap_i, gap_d, max_score, read_size;
    int32_t backtrace, reduce;   /* #ifdef BACKTRACE / REDUCE */
    int32_t swg_w8;              /* SWG/DPU-WRAM common.h:71-79 defines SWG_W8 when MAX_SCORE < 127: int8 cells */
    int32_t request_bytes, result_bytes, params_bytes; /* sizeof(request_t), sizeof(result_t), sizeof(struct DPUParams) */
};
#ifdef MATCH
#define AIM_DPU_K_MATCH (MATCH)
#else
#define AIM_DPU_K_MATCH AIM_DPU_UNSET
#endif
#ifdef MISMATCH
#define AIM_DPU_K_MISMATCH (MISMATCH)
#else
#define AIM_DPU_K_MISMATCH AIM_DPU_UNSET
#endif
#ifdef GAP_O
#define AIM_DPU_K_GAP_O (GAP_O)
#else
#define AIM_DPU_K_GAP_O AIM_DPU_UNSET
#endif
#ifdef GAP_E
#define AIM_DPU_K_GAP_E (GAP_E)
#else
#define AIM_DPU_K_GAP_E AIM_DPU_UNSET
#endif
#ifdef GAP_I
#define AIM_DPU_K_GAP_I (GAP_I)
#else
#define AIM_DPU_K_GAP_I AIM_DPU_UNSET
#endif
#ifdef GAP_D
#define AIM_DPU_K_GAP_D (GAP_D)
#else
#define AIM_DPU_K_GAP_D AIM_DPU_UNSET
#endif
#ifdef MAX_SCORE
#define AIM_DPU_K_MAX_SCORE (MAX_SCORE)
#else
#define AIM_DPU_K_MAX_SCORE AIM_DPU_UNSET
#endif
#ifdef READ_SIZE
#define AIM_DPU_K_READ_SIZE (READ_SIZE)
#else
#define AIM_DPU_K_READ_SIZE AIM_DPU_UNSET
#endif
#ifdef BACKTRACE
#define AIM_DPU_K_BACKTRACE 1
#else
#define AIM_DPU_K_BACKTRACE 0
#endif
#ifdef REDUCE
#define AIM_DPU_K_REDUCE 1
#else
#define AIM_DPU_K_REDUCE 0
#endif
#ifdef SWG_W8
#define AIM_DPU_K_SWG_W8 1
#else
#define AIM_DPU_K_SWG_W8 0
#endif
#ifdef AIM_DPU_NO_WIRE_TYPES /* a caller without the reference's common.h: sizes default to WFA/DPU-MRAM's */
#define AIM_DPU_K_SIZES 8, 32, 32
#else
#define AIM_DPU_K_SIZES (int32_t)sizeof(request_t), (int32_t)sizeof(result_t), (int32_t)sizeof(struct DPUParams)
#endif

dpu_error_t aim_dpu_load(struct dpu_set_t set, const char *binary_path, struct dpu_program_t **program,
                         const struct aim_dpu_knobs *knobs);
#ifdef __cplusplus
#define dpu_load(set, binary, program)                                                                          \
    aim_dpu_load((set), (binary), (program),                                                                    \
                 [] { static const aim_dpu_knobs k = {(int32_t)sizeof(aim_dpu_knobs), AIM_DPU_K_MATCH, AIM_DPU_K_MISMATCH,      \
                      AIM_DPU_K_GAP_O, AIM_DPU_K_GAP_E, AIM_DPU_K_GAP_I, AIM_DPU_K_GAP_D, AIM_DPU_K_MAX_SCORE,   \
                      AIM_DPU_K_READ_SIZE, AIM_DPU_K_BACKTRACE, AIM_DPU_K_REDUCE, AIM_DPU_K_SWG_W8, AIM_DPU_K_SIZES}; return &k; }())
#else
#define dpu_load(set, binary, program)                                                                          \
    aim_dpu_load((set), (binary), (program),                                                                    \
                 &(const struct aim_dpu_knobs){(int32_t)sizeof(struct aim_dpu_knobs), AIM_DPU_K_MATCH, AIM_DPU_K_MISMATCH,      \
                     AIM_DPU_K_GAP_O, AIM_DPU_K_GAP_E, AIM_DPU_K_GAP_I, AIM_DPU_K_GAP_D, AIM_DPU_K_MAX_SCORE,    \
                     AIM_DPU_K_READ_SIZE, AIM_DPU_K_BACKTRACE, AIM_DPU_K_REDUCE, AIM_DPU_K_SWG_W8, AIM_DPU_K_SIZES})
#endif

dpu_error_t dpu_alloc(uint32_t nr_dpus, const char *profile, struct dpu_set_t *dpu_set);
dpu_error_t dpu_free(struct dpu_set_t dpu_set);
dpu_error_t dpu_get_nr_dpus(struct dpu_set_t dpu_set, uint32_t *nr_dpus);
dpu_error_t dpu_prepare_xfer(struct dpu_set_t dpu_set, void *buffer);
dpu_error_t dpu_push_xfer(struct dpu_set_t dpu_set, dpu_xfer_t xfer, const char *symbol_name, uint32_t symbol_offset,
                          size_t length, dpu_xfer_flags_t flags);
dpu_error_t dpu_broadcast_to(struct dpu_set_t dpu_set, const char *symbol_name, uint32_t symbol_offset, const void *src,
                             size_t length, dpu_xfer_flags_t flags);
dpu_error_t dpu_launch(struct dpu_set_t dpu_set, dpu_launch_policy_t policy);
dpu_error_t dpu_sync(struct dpu_set_t dpu_set);
dpu_error_t dpu_log_read(struct dpu_set_t set, FILE *stream);

/* The three phases of the last dpu_launch as aim_align_batch measured them with CUDA events (H2D, kernels, D2H; ms):
 * the host's own "DPU Kernel" timer (host.c:282-299) brackets all three. */
void aim_dpu_last_phases(struct dpu_set_t set, double phase_ms[3]);

#ifdef __cplusplus
}
#endif
#endif /* AIM_B200_DPU_H */
