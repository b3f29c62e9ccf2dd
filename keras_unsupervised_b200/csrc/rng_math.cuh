// Counter-based uniforms and the elementwise math of the Gibbs chain.
//
// The reference draws `K.random_uniform(shape)` (float32 in [0,1)) for every
// sampling node (ku/ebm/rbm.py:46,52,121) and thresholds with a strict `<`
// (K.less).  Here the draw is Philox4x32-10 keyed by the engine seed with the
// counter (column/4, global row, draw id) so that a sample depends only on its
// coordinates - not on tiling, launch shape or the number of GPUs.  The float
// lattice is the 2^-23 grid TensorFlow's RandomUniform uses (23 mantissa bits).
#pragma once
#include <cstdint>

namespace kucd {

struct Philox4 {
  uint32_t x, y, z, w;
};

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                          uint32_t k0, uint32_t k1) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
#ifdef __CUDA_ARCH__
    const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
#else
    const uint64_t p0 = static_cast<uint64_t>(M0) * c0, p1 = static_cast<uint64_t>(M1) * c2;
    const uint32_t hi0 = static_cast<uint32_t>(p0 >> 32), lo0 = static_cast<uint32_t>(p0);
    const uint32_t hi1 = static_cast<uint32_t>(p1 >> 32), lo1 = static_cast<uint32_t>(p1);
#endif
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0;
    c1 = n1;
    c2 = n2;
    c3 = n3;
    k0 += W0;
    k1 += W1;
  }
  return Philox4{c0, c1, c2, c3};
}

// 23 random mantissa bits -> float in [0, 1) on the 2^-23 lattice (exact).
__host__ __device__ __forceinline__ float u01_from_bits(uint32_t bits) {
  return static_cast<float>(bits >> 9) * 1.1920928955078125e-07f;  // 2^-23
}

#ifdef __CUDACC__
// logistic in fp32: exp via ex2.approx (2 ulp), reciprocal via rcp.approx (1 ulp)
__device__ __forceinline__ float sigmoid_f32(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

// log(1 + e^x), overflow-free; equals the reference's naive K.log(1 + K.exp(x))
// (ku/ebm/rbm.py:74) wherever that does not overflow.
__device__ __forceinline__ float softplus_f32(float x) { return fmaxf(x, 0.0f) + log1pf(__expf(-fabsf(x))); }

// Column statistics (db, dc of rbm.py:129-134) are summed with one fp32 atomicAdd per (32- or 64-row group, column), in
// whatever order the warps arrive - and fp32 addition does not commute with rounding, so two runs of the same step used to
// differ in the last bits (1e-8 on a bias), which a later discrete decision (the bf16 rounding of a stored probability, a
// Bernoulli threshold) now and then amplified to lr * 2^-10.  Rounded to the power-of-two grid 2^(ceil(log2 bound) - 24),
// every partial sum is a multiple of a grid on which all sums up to `bound` are exact in fp32: the adds commute and the
// statistic is the same in every run, on every schedule and tiling.  The grid is the fp32 spacing at magnitude `bound`,
// i.e. what the running sum of an atomic accumulation of that size rounds to anyway.  0/1 counts are integers: unchanged.
__device__ __forceinline__ float stat_grid_round(float s, int bound) {
  const int e = 32 - __clz((bound > 2 ? bound : 2) - 1);  // ceil(log2(bound))
  const float g = __int_as_float((127 + e - 24) << 23), inv = __int_as_float((127 - e + 24) << 23);
  return rintf(s * inv) * g;
}
#endif

}  // namespace kucd
