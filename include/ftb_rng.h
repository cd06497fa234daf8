/* ftb_rng.h — the counter-based random-number CONTRACT of functracer_b200.
 *
 * The reference draws soft-shadow directions and depth-of-field perturbations from an
 * unseeded System.Random() created per call (FuncTracer/Jitter.fs:27), so its output on
 * scenes with `softdirectional` lights or camera `focus` is not reproducible.  This repo
 * replaces that one source of entropy — and nothing else of Jitter.fs — by a stateless hash
 * keyed on (seed, primary-sample index, bounce depth, light index, sample k, rejection
 * attempt, dimension).  It is a specification, not an algorithm of the reference: both the
 * CUDA kernels and the CPU oracle must produce exactly these bits, so the definition lives
 * in this one header (plain C, also valid CUDA device code).
 *
 * Every uniform has 24 bits of resolution so that 2u-1 is exactly representable in both
 * float and double: the FP32 kernels, the FP64 verification kernels and the oracle consume
 * identical sample positions.
 */
#ifndef FTB_RNG_H
#define FTB_RNG_H

#include <stdint.h>

#if defined(__CUDACC__)
#define FTB_HD __host__ __device__ __forceinline__
#else
#define FTB_HD static inline
#endif

#define FTB_RNG_STREAM_CAMERA 0xFFFFu /* `light` key of the depth-of-field draw */

FTB_HD uint64_t ftb_mix64(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

/* 24 random bits for the given key. */
FTB_HD uint32_t ftb_rng_bits24(uint64_t seed, uint64_t sample, uint32_t depth, uint32_t light,
                               uint32_t k, uint32_t attempt, uint32_t dim)
{
    uint64_t h = ftb_mix64(seed ^ (sample * 0x9E3779B97F4A7C15ULL + 0x632BE59BD9B4E019ULL));
    h = ftb_mix64(h ^ ((uint64_t)depth << 48) ^ ((uint64_t)light << 32) ^ (uint64_t)k);
    h = ftb_mix64(h ^ ((uint64_t)attempt << 1) ^ (uint64_t)dim);
    return (uint32_t)(h >> 40);
}

/* Jitter.uniform (Jitter.fs:9-10): 2*NextDouble()-1, here on the 24-bit lattice in [-1, 1). */
#define FTB_RNG_TO_UNIT(bits) (((double)(int32_t)(2 * (int32_t)(bits)-16777216)) * (1.0 / 16777216.0))

#endif
