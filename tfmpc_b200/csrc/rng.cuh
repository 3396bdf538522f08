// Counter-based random numbers for the stochastic plant (GymEnv.step, reference tfmpc/envs/gymenv.py:15-25) and for the
// random initial actions of iLQR.start (ilqr.py:59-70): Philox4x32-10 (Salmon et al., SC'11), keyed by the caller's seed,
// counter = (row, component, call offset, stream tag).  No state is kept on the device: the same (seed, offset) reproduces
// the same draws on any grid shape, and two calls differ only through the offset the caller advances.
#pragma once
#include <stdint.h>

#include "common.cuh"

struct Philox {
  uint32_t c[4], k[2];
  __device__ __forceinline__ Philox(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
    c[0] = c0; c[1] = c1; c[2] = c2; c[3] = c3;
    k[0] = (uint32_t)seed; k[1] = (uint32_t)(seed >> 32);
  }
  __device__ __forceinline__ void round() {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    const uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0], hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k[0], n2 = hi0 ^ c[3] ^ k[1];
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
  }
  // 4 x 32 random bits for the current counter; advances the last counter word so that the next call is a fresh block
  __device__ __forceinline__ void next(uint32_t out[4]) {
    const uint32_t s0 = c[0], s1 = c[1], s2 = c[2], s3 = c[3], k0 = k[0], k1 = k[1];
#pragma unroll
    for (int r = 0; r < 10; r++) {
      round();
      k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
    }
#pragma unroll
    for (int i = 0; i < 4; i++) out[i] = c[i];
    c[0] = s0; c[1] = s1; c[2] = s2; c[3] = s3 + 0x10000u; k[0] = k0; k[1] = k1;
  }
};

// uniform in the OPEN interval (0, 1): 24 (float) / 53 (double) random mantissa bits, centred in their cell
__device__ __forceinline__ double u01(uint32_t a, uint32_t b) {
  const uint64_t m = (((uint64_t)a << 32) | b) >> 11;
  return ((double)m + 0.5) * (1.0 / 9007199254740992.0);
}

// tf.random.truncated_normal(mean 0, stddev sigma): values beyond 2 sigma are re-drawn, i.e. the normal law conditioned on
// |z| <= 2 -- sampled exactly by inverting the CDF on [Phi(-2), Phi(2)] (one uniform, no rejection loop)
__device__ __forceinline__ double trunc_normal2(double u) {
  const double lo = 0.022750131948179195;   // Phi(-2)
  return normcdfinv(lo + u * (1.0 - 2.0 * lo));
}

// tf.random.gamma(alpha, beta = 1 / scale): Marsaglia & Tsang (2000) squeeze-free form, with the alpha < 1 boost
// Gamma(alpha) = Gamma(alpha + 1) U^(1 / alpha).  `g` supplies fresh blocks; the loop accepts with probability > 0.95.
__device__ __forceinline__ double gamma_draw(Philox &g, double alpha, double scale) {
  uint32_t r[4];
  double boost = 1.0;
  if (alpha < 1.0) {
    g.next(r);
    boost = pow(u01(r[0], r[1]), 1.0 / alpha);
    alpha += 1.0;
  }
  const double d = alpha - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
  for (int tries = 0; tries < 64; tries++) {
    g.next(r);
    const double x = normcdfinv(u01(r[0], r[1])), u = u01(r[2], r[3]);
    const double t = 1.0 + c * x;
    if (t <= 0.0) continue;
    const double v = t * t * t;
    if (log(u) < 0.5 * x * x + d - d * v + d * log(v)) return d * v * scale * boost;
  }
  return d * scale * boost;   // (probability ~1e-80)
}
