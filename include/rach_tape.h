/*
 * rach_tape.h -- the Philox4x32-10 "draw tape" shared by host, device and oracle.
 *
 * The reference simulators draw from libc rand(), one global stream consumed in
 * UE-index order (RandomAccessWithNOMA.c:219 srand, call sites :393 :394 :478 :502
 * :514 :540 :670 :685 :701; RandomAccessSimulatorBeta.c:232 :253 :264 :290 :374 :387
 * :403; RandomAccessSimulator.c:160 :170 :187 :238 :250 :251; NOMA.c:133 :142 :168
 * :187 :284 :286 :460 :482 :503 :519 :520).  That stream cannot be produced in
 * parallel, so every build of the state machine in this repository (the reference
 * sources compiled in tape mode under oracle/_ref, the C restatement in oracle/, and
 * the CUDA engine) replaces `rand()` by a counter-based draw:
 *
 *     rand31(seed, rep, ue, ms, k) = philox4x32_10(ctr, key)[k & 3] >> 1
 *       key = (seed_lo, seed_hi)
 *       ctr = (ue, ms, rep, (k >> 2) | (tag << 16))
 *
 * where k is the running number of draws that UE has made in that millisecond
 * (0,1,2,...), and `tag` separates streams that are not owned by a UE
 * (RACH_TAPE_TAG_UE = 0 for UE draws, RACH_TAPE_TAG_BS = 1 for the base-station-side
 * draws of NOMA.c:284/286 where `ue` carries sector*64+pair).  The value range is
 * that of glibc/macOS rand(): [0, 2^31-1] == [0, RAND_MAX].
 *
 * Plain C99 / C++ / CUDA: everything is static inline, no dependencies.
 */
#ifndef RACH_TAPE_H
#define RACH_TAPE_H

#include <stdint.h>

#ifdef __CUDACC__
#define RACH_HD __host__ __device__ __forceinline__
#else
#define RACH_HD static inline
#endif

#define RACH_TAPE_TAG_UE 0u
#define RACH_TAPE_TAG_BS 1u

#define RACH_PHILOX_M0 0xD2511F53u
#define RACH_PHILOX_M1 0xCD9E8D57u
#define RACH_PHILOX_W0 0x9E3779B9u
#define RACH_PHILOX_W1 0xBB67AE85u

typedef struct { uint32_t v[4]; } rach_u32x4;

RACH_HD uint32_t rach_mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

/* Philox4x32-10 (Salmon et al., SC'11), bit-compatible with Random123. */
RACH_HD rach_u32x4 rach_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                      uint32_t k0, uint32_t k1) {
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = rach_mulhi32(RACH_PHILOX_M0, c0), lo0 = RACH_PHILOX_M0 * c0;
        uint32_t hi1 = rach_mulhi32(RACH_PHILOX_M1, c2), lo1 = RACH_PHILOX_M1 * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += RACH_PHILOX_W0; k1 += RACH_PHILOX_W1;
    }
    rach_u32x4 out;
    out.v[0] = c0; out.v[1] = c1; out.v[2] = c2; out.v[3] = c3;
    return out;
}

/* One tape block = draws 4*block .. 4*block+3 of (seed, rep, ue, ms, tag). */
RACH_HD rach_u32x4 rach_tape_block(uint64_t seed, uint32_t rep, uint32_t ue, uint32_t ms,
                                   uint32_t block, uint32_t tag) {
    return rach_philox4x32_10(ue, ms, rep, block | (tag << 16),
                              (uint32_t)seed, (uint32_t)(seed >> 32));
}

/* Draw k of (seed, rep, ue, ms): the replacement for one rand() call. */
RACH_HD int rach_tape_rand31(uint64_t seed, uint32_t rep, uint32_t ue, uint32_t ms,
                             uint32_t k, uint32_t tag) {
    rach_u32x4 b = rach_tape_block(seed, rep, ue, ms, k >> 2, tag);
    return (int)(b.v[k & 3u] >> 1);
}

/* (float)r / (float)RAND_MAX > 0.1  (RandomAccessWithNOMA.c:670-671).  (float)RAND_MAX
 * rounds to 2^31, the division by a power of two is exact, so the test is on
 * (float)r alone; written with the reference's own expression so that the compiler's
 * float semantics decide, not an integer threshold derived by hand. */
RACH_HD int rach_msg3_success(int r31) {
    float p = (float)r31 / (float)2147483647;
    return p > 0.1;
}

#endif /* RACH_TAPE_H */
