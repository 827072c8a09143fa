// Philox4x32-10 counter-based generator, shared by host and device code.
//
// Throughput mode of the sweep draws its Metropolis uniforms on the device: key = 64-bit seed,
// counter = (proposal_index >> 1, global chain index, global sweep index lo, hi); each block of four
// 32-bit outputs yields two doubles in [0, 1) built from 53 random bits the way NumPy's legacy
// `random_sample` builds them ((a >> 5) * 2^26 + (b >> 6)) / 2^53.  Keying by the *global* chain and
// sweep index makes a chain's stream independent of how chains are sharded over GPUs.
// The reference draws from the global MT19937 stream instead (lqmc.py:317); parity runs feed that
// stream's numbers through the `uniforms` argument of the C ABI.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define LQMC_HD __host__ __device__ __forceinline__
#else
#define LQMC_HD static inline
#endif

LQMC_HD void lqmc_philox_mulhilo(uint32_t a, uint32_t b, uint32_t* hi, uint32_t* lo) {
  const uint64_t p = (uint64_t)a * (uint64_t)b;
  *hi = (uint32_t)(p >> 32);
  *lo = (uint32_t)p;
}

LQMC_HD void lqmc_philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int round = 0; round < 10; ++round) {
    uint32_t hi0, lo0, hi1, lo1;
    lqmc_philox_mulhilo(0xD2511F53u, c[0], &hi0, &lo0);
    lqmc_philox_mulhilo(0xCD9E8D57u, c[2], &hi1, &lo1);
    const uint32_t n0 = hi1 ^ c[1] ^ k0;
    const uint32_t n1 = lo1;
    const uint32_t n2 = hi0 ^ c[3] ^ k1;
    const uint32_t n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

// Uniform in [0,1) for proposal `pidx` of (chain, sweep) under `seed`.
LQMC_HD double lqmc_philox_uniform(uint64_t seed, uint64_t chain, uint64_t sweep, uint32_t pidx) {
  uint32_t c[4] = {pidx >> 1, (uint32_t)chain, (uint32_t)sweep, (uint32_t)(sweep >> 32) ^ (uint32_t)(chain >> 32)};
  lqmc_philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  const uint32_t a = (pidx & 1u) ? c[2] : c[0];
  const uint32_t b = (pidx & 1u) ? c[3] : c[1];
  return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}
