/* rt_rng.h — the shared counter-based RNG that stands in for libc rand().
 *
 * The reference draws every random number from rand() (Src/Math.h:17-20,
 * Src/MathHelper.cpp:17-20) seeded by srand(time) (Src/RayTracerProgram.cpp:439), and reads its
 * diffuse directions through one racy static cursor (Src/Math.cpp:36-39).  Neither is
 * reproducible across threads, let alone devices.  To compare a GPU render with the reference's
 * CPU code, both sides use this stateless generator instead: the n-th rand() call made while
 * tracing camera ray (pixel, sample) returns rt_rand31(rt_rng_key(seed, pixel, sample), n).
 * The oracle harness interposes rand() with it; the CUDA path calls it directly.
 *
 * RAND_MAX is glibc's 2147483647, so RMath::Random() = (float)r / RAND_MAX = (float)r * 2^-31
 * (the int->float conversion of RAND_MAX rounds to 2^31; Math.h:19).
 */
#ifndef RT_RNG_H
#define RT_RNG_H

#include <stdint.h>

#if defined(__CUDACC__)
#define RT_HD __host__ __device__ __forceinline__
#else
#define RT_HD static inline
#endif

#define RT_RAND_MAX 2147483647

/* stream ids that are not camera rays */
#define RT_RNG_TABLE_PIXEL 0xFFFFFFFFu  /* PseudoRandomUnitVectors initialisation stream */

RT_HD uint32_t rt_mix32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352du;
    x ^= x >> 15; x *= 0x846ca68bu;
    x ^= x >> 16;
    return x;
}

/* sample = pass*4 + sub-sample for antialiased passes, = pass for single-ray passes */
RT_HD uint32_t rt_rng_key_pixel(uint32_t seed, uint32_t pixel)
{
    uint32_t k = rt_mix32(seed ^ 0xA511E9B3u);
    return rt_mix32(k + pixel);
}

RT_HD uint32_t rt_rng_key_sample(uint32_t pixel_key, uint32_t sample)
{
    return rt_mix32(pixel_key ^ (sample * 0x9E3779B1u + 0x7F4A7C15u));
}

RT_HD uint32_t rt_rng_key(uint32_t seed, uint32_t pixel, uint32_t sample)
{
    return rt_rng_key_sample(rt_rng_key_pixel(seed, pixel), sample);
}

/* n-th draw of a stream, in [0, RAND_MAX] like rand() */
RT_HD int32_t rt_rand31(uint32_t key, uint32_t n)
{
    uint32_t r = rt_mix32(key + n * 0x9E3779B9u);
    r = rt_mix32(r ^ key);
    return (int32_t)(r >> 1);
}

/* RMath::Random(), Math.h:17-20 */
RT_HD float rt_random01(uint32_t key, uint32_t n)
{
    return (float)rt_rand31(key, n) / 2147483648.0f;
}

#endif /* RT_RNG_H */
