/*
 * pwa_debug.h -- debug-only entry points of libpwa_b200.so.  NOT part of the product ABI (include/pwa.h): they are compiled
 * in only by `make -C <package>/csrc TIMELINE=1` (test infrastructure for tools/timeline*.py and tools/ws_profile.py) and,
 * unlike every product entry point, allocate device memory and synchronise.
 */
#ifndef PWA_DEBUG_H_
#define PWA_DEBUG_H_
#ifdef __cplusplus
extern "C" {
#endif

/* With PWA_TIMELINE=1 in the environment, CTA 0 of the tcgen05 attention kernels records clock64 stamps / per-role cycle
 * counters into a library-owned device buffer; this copies them to the host.  Returns the number of bytes copied, 0 when
 * no timeline exists. */
int pwa_debug_fwd_timeline(void* host_dst, int bytes);

#ifdef __cplusplus
}
#endif
#endif
