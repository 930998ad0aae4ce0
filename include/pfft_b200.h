/* include/pfft_b200.h -- extensions of pfft_b200 beyond the PFFT API (plain C ABI).
 *
 * Nothing here exists in the reference; these entry points expose what a GPU
 * library needs in addition: the CUDA stream, device-resident buffers, the job
 * bootstrap used by non-MPI hosts (Python / torchrun), and introspection of the
 * schedule for tests and for the roofline accounting of bench.py.
 */
#ifndef PFFT_B200_EXT_H
#define PFFT_B200_EXT_H 1
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* library version string, e.g. "pfft_b200 0.1 (sm_100a)" */
const char *pfftb200_version(void);

/* Last error message of the calling process ("" if none).  Planner functions return
 * NULL like the reference does (kernel/partrafo.c:337-373); this says why. */
const char *pfftb200_last_error(void);

/* Describe the schedule rank `pid` of a mesh would execute, without MPI, CUDA or any
 * allocation.  `kind`: 0 c2c, 1 r2c, 2 c2r, 3 r2r.  iblock/oblock/kinds/skip may be NULL.
 * Writes a NUL-terminated JSON document into buf (truncated to buflen) and returns
 * the length it needs. */
size_t pfftb200_describe_schedule(int kind, int rnk_n, const ptrdiff_t *n, const ptrdiff_t *ni,
                                  const ptrdiff_t *no, ptrdiff_t howmany, const ptrdiff_t *iblock,
                                  const ptrdiff_t *oblock, int rnk_pm, const int *np, int pid, int sign,
                                  const int *kinds, const int *skip_trafos, unsigned pfft_flags,
                                  char *buf, size_t buflen);

/* Integer layer without a communicator (same results as pfft_local_block_many_*). */
void pfftb200_local_block(int kind, int rnk_n, const ptrdiff_t *ni, const ptrdiff_t *no,
                          const ptrdiff_t *iblock, const ptrdiff_t *oblock, int rnk_pm, const int *np, int pid,
                          unsigned pfft_flags, ptrdiff_t *local_ni, ptrdiff_t *local_i_start,
                          ptrdiff_t *local_no, ptrdiff_t *local_o_start);

/* Device selection and stream. By default a rank uses the calling thread's current
 * CUDA device (so torch.cuda.set_device(local_rank) is respected); launched through
 * pfftrun it uses rank % device_count. */
int pfftb200_set_device(int device);
int pfftb200_get_device(void);
void pfftb200_set_stream(void *cuda_stream);   /* stream used by plans created afterwards */
void *pfftb200_get_stream(void);

/* Device-resident buffers for benchmarks / GPU-native callers (plans accept them
 * wherever they accept pfft_alloc_* memory). */
void *pfftb200_malloc_device(size_t bytes);
void pfftb200_free_device(void *p);

/* Transport for the global exchanges: "auto" (default), "nccl", "p2p".  Returns 0 on success. */
int pfftb200_set_transport(const char *name);
const char *pfftb200_get_transport(void);

/* Introspection of a live plan (both precisions share the layout of this call:
 * pass the pfft_plan / pfftf_plan handle as void*). */
size_t pfftb200_plan_describe(const void *plan, char *buf, size_t buflen);
/* Asynchronous execute on the plan's stream (pfft_execute synchronises like the
 * blocking reference call does); in/out may be NULL to use the planned arrays. */
void pfftb200_execute_async(const void *plan, void *in, void *out);
/* kernels launched by this process since load (evidence for "gpu_launches") */
unsigned long long pfftb200_launch_count(void);
/* per-stage device time of the last execute with timing enabled, milliseconds; returns #stages */
void pfftb200_enable_stage_timing(const void *plan, int on);
int pfftb200_stage_times(const void *plan, double *ms, int max_stages);
/* host time spent ordering each exchange of the last blocking execute (barriers / flag waits), milliseconds */
int pfftb200_exchange_times(const void *plan, double *ms, int max_exchanges);


/* TEST SUPPORT (tests/test_kernel_emulation.py): stage `stage` of rank `pid`'s schedule run through the body
 * of the any-length CUDA stage kernel on the CPU -- the same host+device source the GPU executes, threads
 * emulated between barriers -- on host buffers.  prec 0: double, 1: float.  Nothing in the library calls this;
 * it is not a CPU fallback.  Returns 0 on success. */
int pfftb200_emulate_stage(int prec, int kind, int rnk_n, const ptrdiff_t *n, const ptrdiff_t *ni, const ptrdiff_t *no,
                           ptrdiff_t howmany, const ptrdiff_t *iblock, const ptrdiff_t *oblock, int rnk_pm,
                           const int *np, int pid, int sign, const int *kinds, const int *skip_trafos,
                           unsigned pfft_flags, int stage, const void *in, void *const *outs);

/* Introspection without a device: the kernel family ("pow2" | "reg" | "mixed") and tile geometry the planner picks
 * for every stage of rank `pid` (JSON list, same calling convention as pfftb200_describe_schedule). */
size_t pfftb200_describe_kernels(int prec, int kind, int rnk_n, const ptrdiff_t *n, const ptrdiff_t *ni, const ptrdiff_t *no,
                                 ptrdiff_t howmany, const ptrdiff_t *iblock, const ptrdiff_t *oblock, int rnk_pm,
                                 const int *np, int pid, int sign, const int *kinds, const int *skip_trafos,
                                 unsigned pfft_flags, char *buf, size_t buflen);

/* Introspection without a device: the device-side ordering of the p2p transport for rank `pid` (JSON: per boundary
 * remote / receive area / storing ranks, per stage the [rank, previous-execute, stages-completed] triples it waits for). */
size_t pfftb200_describe_exchange_ordering(int kind, int rnk_n, const ptrdiff_t *n, const ptrdiff_t *ni, const ptrdiff_t *no,
                                           ptrdiff_t howmany, const ptrdiff_t *iblock, const ptrdiff_t *oblock, int rnk_pm,
                                           const int *np, int pid, int sign, const int *kinds, const int *skip_trafos,
                                           unsigned pfft_flags, char *buf, size_t buflen);

#ifdef __cplusplus
}
#endif
#endif
