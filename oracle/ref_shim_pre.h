/*
 * ref_shim_pre.h -- TEST INFRASTRUCTURE (oracle).  Not part of the product path.
 *
 * Prepended (through a pipe, see oracle/build_ref.sh) to the UNMODIFIED text of a reference
 * simulator read from /root/reference at build time.  It redirects the handful of libc
 * entry points the reference's `main` uses so that the same source can be
 *   (a) driven as a function (`main` -> ref_main, argv built by ref_shim_post.h),
 *   (b) run in DRAW-TAPE mode: every `rand()` call site of the state machine
 *       (RandomAccessWithNOMA.c:393,394,478,502,514,540,670,685,701;
 *        RandomAccessSimulatorBeta.c:232,253,264,290,374,387,403;
 *        RandomAccessSimulator.c:160,170,187,238,250,251) has `user` (the UE) and `time`
 *       (the ms) in scope, so a function-like macro can key the draw by (UE, ms, k),
 *   (c) observed: the UE array is captured when the reference frees it, the scalars it
 *       prints are captured from its own printf calls,
 *   (d) kept off the file system (mkdir / fopen of result files are redirected).
 * No reference text is stored in this repository; the only edits are the sed lines in
 * build_ref.sh (narrowing the nUE sweep to one point, and for B the parameter locals).
 */
#ifndef REF_SHIM_PRE_H
#define REF_SHIM_PRE_H

#include <stdio.h>
#include <stdlib.h>
#include <memory.h>
#include <string.h>
#include <math.h>
#include <time.h>
#include <unistd.h>
#include <dirent.h>
#include <sys/stat.h>
#include <sys/types.h>
#include <complex.h>
#include <setjmp.h>
#include <stdarg.h>

int   ref_tape_rand(int ue, int ms);
int   ref_tape_rand_bs(int sector, int ms);
void  ref_tape_srand(unsigned seed);
void* ref_calloc_hook(size_t n, size_t sz);
void  ref_free_hook(void* p);
int   ref_printf_hook(const char* fmt, ...);
FILE* ref_fopen_hook(const char* name, const char* mode);
void  ref_exit_hook(int code);

/* parameter globals read by the sed-patched sweep / parameter lines */
extern int ref_nue;
extern int ref_p_nPreamble, ref_p_backoff, ref_p_nGrantUL, ref_p_maxRarWindow,
           ref_p_maxMsg2TxCount, ref_p_accessTime, ref_p_distribution;

#define rand()            ref_tape_rand(user->idx, time)
#define srand(s)          ref_tape_srand(s)
#define calloc(n, sz)     ref_calloc_hook((n), (sz))
#define free(p)           ref_free_hook((void*)(p))
#define printf(...)       ref_printf_hook(__VA_ARGS__)
#define fopen(name, mode) ref_fopen_hook((name), (mode))
#define exit(c)           ref_exit_hook(c)
#define mkdir(path, mode) ((void)0)
#define main              ref_main

#endif
