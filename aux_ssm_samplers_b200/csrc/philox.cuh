// Counter-based Philox4x32-10 draws shared by the fill kernels (host_side.cu) and the in-kernel generation of the
// cSMC sweeps (draw pointers NULL).  One definition of the element layout, so an array filled by
// auxssm_philox_{uniform,normal}_* with (seed, call) and the same logical array generated inside a sweep kernel are
// bit-identical; oracle/philox.py restates it on the CPU.
//
// Layout: element e of an array belongs to Philox block e / PER (counter words: block lo, block hi, call lo, call hi;
// key words: seed lo, seed hi), slot e % PER, with PER = 4 (fp32) or 2 (fp64).
//   uniform  fp32: ((word >> 8) + 0.5) / 2^24       fp64: (((w0 << 32 | w1) >> 11) + 0.5) / 2^53, (w2, w3) likewise
//   normal   Box-Muller on consecutive uniform pairs (a, b): sqrt(-2 log a) (cos 2 pi b, sin 2 pi b)
#pragma once
#include <stdint.h>

namespace auxssm {

__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

template <typename T> struct PhiloxPer { static constexpr int N = (sizeof(T) == 4) ? 4 : 2; };

__device__ __forceinline__ float philox_u01(uint32_t a) { return ((float)(a >> 8) + 0.5f) * (1.0f / 16777216.0f); }
__device__ __forceinline__ double philox_u01(uint32_t a, uint32_t b) {
  const uint64_t m = (((uint64_t)a << 32) | b) >> 11;  // 53 bits
  return ((double)m + 0.5) * (1.0 / 9007199254740992.0);
}

enum { PHILOX_UNIFORM = 0, PHILOX_NORMAL = 1 };

// all PER values of one block (not inlined: the sweep kernels call it from register-tight loops)
static __device__ __noinline__ void philox_block(int mode, uint64_t seed, uint64_t call, uint64_t blk, float (&v)[4]) {
  uint32_t c[4] = {(uint32_t)blk, (uint32_t)(blk >> 32), (uint32_t)call, (uint32_t)(call >> 32)};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  const float a0 = philox_u01(c[0]), a1 = philox_u01(c[1]), a2 = philox_u01(c[2]), a3 = philox_u01(c[3]);
  if (mode == PHILOX_UNIFORM) {
    v[0] = a0; v[1] = a1; v[2] = a2; v[3] = a3;
  } else {
    // fp32 Box-Muller on the special-function unit: MUFU.LG2 / RSQ / SIN / COS (absolute errors ~2^-21, i.e. the
    // fp32 rounding level of the result) instead of the ~20-instruction libm paths of logf / sincospif: in the
    // cSMC sweeps with in-kernel draws the generator was about half of the instructions.  The angle is folded into
    // [-pi, pi), where MUFU.SIN / COS are most accurate.  ((word >> 8) + 0.5) rounds to 2^24 in fp32 when the 24 bits
    // are all ones (one uniform in 2^24 is exactly 1.0f): then -2 log a = -0.0, rsqrt(-0.0) = -inf and the radius
    // -0 * -inf = NaN -- a NaN normal every ~1.7e7 pairs (it froze 32 of the 8192 configs[4] chains at their
    // initial state).  The clamp makes that radius ~0, which is what sqrt(-2 log 1) is; all other draws keep their bits.
    const float t0 = fmaxf(-2.0f * __logf(a0), 1e-30f), t1 = fmaxf(-2.0f * __logf(a2), 1e-30f);
    const float r0 = t0 * rsqrtf(t0), r1 = t1 * rsqrtf(t1);
    const float w0 = 6.283185307179586f * (a1 - (a1 >= 0.5f ? 1.0f : 0.0f));
    const float w1 = 6.283185307179586f * (a3 - (a3 >= 0.5f ? 1.0f : 0.0f));
    v[0] = r0 * __cosf(w0); v[1] = r0 * __sinf(w0); v[2] = r1 * __cosf(w1); v[3] = r1 * __sinf(w1);
  }
}
static __device__ __noinline__ void philox_block(int mode, uint64_t seed, uint64_t call, uint64_t blk, double (&v)[2]) {
  uint32_t c[4] = {(uint32_t)blk, (uint32_t)(blk >> 32), (uint32_t)call, (uint32_t)(call >> 32)};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  const double a0 = philox_u01(c[0], c[1]), a1 = philox_u01(c[2], c[3]);
  if (mode == PHILOX_UNIFORM) {
    v[0] = a0; v[1] = a1;
  } else {
    const double r0 = sqrt(-2.0 * log(a0));
    double s0, c0;
    sincospi(2.0 * a1, &s0, &c0);
    v[0] = r0 * c0; v[1] = r0 * s0;
  }
}

__device__ __forceinline__ float philox_pick(const float (&v)[4], int s) {
  return s == 0 ? v[0] : (s == 1 ? v[1] : (s == 2 ? v[2] : v[3]));
}
__device__ __forceinline__ double philox_pick(const double (&v)[2], int s) { return s == 0 ? v[0] : v[1]; }

// one element
template <typename T> __device__ __forceinline__ T philox_elem(int mode, uint64_t seed, uint64_t call, uint64_t e) {
  constexpr int PER = PhiloxPer<T>::N;
  T v[PER];
  philox_block(mode, seed, call, e / PER, v);
  return philox_pick(v, (int)(e % PER));
}

// n <= NMAX consecutive elements starting at e0 (a Philox block is evaluated once per block touched)
template <typename T, int NMAX>
__device__ __forceinline__ void philox_run(int mode, uint64_t seed, uint64_t call, uint64_t e0, int n, T (&out)[NMAX]) {
  constexpr int PER = PhiloxPer<T>::N;
  T v[PER];
  uint64_t cur = ~(uint64_t)0;
#pragma unroll
  for (int k = 0; k < NMAX; ++k) {
    if (k < n) {
      const uint64_t e = e0 + (uint64_t)k, blk = e / PER;
      if (blk != cur) { philox_block(mode, seed, call, blk, v); cur = blk; }
      out[k] = philox_pick(v, (int)(e % PER));
    }
  }
}

// Draw-source of a cSMC sweep: external arrays, or in-kernel generation when `normals == NULL`.
// Calls: normals = call0, u_resample = call0 + 1, u_kill = call0 + 2 (killing only), u_backward = the next one --
// the order in which the host mirror draws the same arrays from a PhiloxDraws key.
struct RngCtl {
  uint64_t seed, call0;
  int on;  // 1: generate in the kernel
};

}  // namespace auxssm
