// In-shared-memory complex FFTs for the cross-correlation kernels (sm_100a).
//
// Replaces extern/RealFFT (FFTReal<float>::do_fft / do_ifft, FFTReal.hpp:145-245): the
// reference runs 12 radix-2 real transforms of N points per strand-pair; here two real
// channels ride one complex N-point transform (A + iC, G + iT), the forward transform is
// an in-place decimation-in-frequency network that leaves the spectrum in mixed-radix
// digit-reversed order, and the inverse is its exact adjoint (decimation-in-time) that
// consumes that order and returns natural order -- so no bit-reversal pass exists at all:
// the spectral product in between is element-wise and order-agnostic.
//
// One pass = every thread pulls R points (stride S) from shared memory into registers,
// does an R-point DFT, applies the inter-stage twiddles and writes back in place.
#pragma once
#include <cuda_runtime.h>

namespace sx {

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// a * conj(b)
__device__ __forceinline__ float2 cmulc(float2 a, float2 b) {
  return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
// multiply by -i (forward) or +i (inverse)
template <bool INV>
__device__ __forceinline__ float2 rot90(float2 a) {
  return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
}

// ---- small DFTs in registers: X[p] = sum_m x[m] e^{-/+ 2 pi i p m / R} -------------------
template <bool INV>
__device__ __forceinline__ void dft2(float2 &a, float2 &b) {
  float2 t = a;
  a = cadd(t, b);
  b = csub(t, b);
}

template <bool INV>
__device__ __forceinline__ void dft4(float2 &x0, float2 &x1, float2 &x2, float2 &x3) {
  float2 t0 = cadd(x0, x2), t1 = csub(x0, x2);
  float2 t2 = cadd(x1, x3), t3 = rot90<INV>(csub(x1, x3));
  x0 = cadd(t0, t2);
  x2 = csub(t0, t2);
  x1 = cadd(t1, t3);
  x3 = csub(t1, t3);
}

// e^{-/+ 2 pi i k / 16}, k = 1..3 (others by symmetry)
#define SX_C1 0.92387953251128674f
#define SX_S1 0.38268343236508977f
#define SX_R2 0.70710678118654752f

template <bool INV>
__device__ __forceinline__ float2 w16(int k) {  // k compile-time after unrolling
  float c, s;
  switch (k) {
    case 0: c = 1.f; s = 0.f; break;
    case 1: c = SX_C1; s = SX_S1; break;
    case 2: c = SX_R2; s = SX_R2; break;
    case 3: c = SX_S1; s = SX_C1; break;
    case 4: c = 0.f; s = 1.f; break;
    case 6: c = -SX_R2; s = SX_R2; break;
    case 9: c = -SX_C1; s = -SX_S1; break;  // 9 = 8 + 1
    default: c = 1.f; s = 0.f; break;
  }
  return make_float2(c, INV ? s : -s);
}

template <int R, bool INV>
struct Dft;

template <bool INV>
struct Dft<4, INV> {
  static __device__ __forceinline__ void run(float2 *x) { dft4<INV>(x[0], x[1], x[2], x[3]); }
};

// R = 8 as 4 x 2: inputs m = j + 2a (j<2, a<4); X[p + 4k] = DFT2_j( W8^{jp} DFT4_a x[j+2a] )
template <bool INV>
struct Dft<8, INV> {
  static __device__ __forceinline__ void run(float2 *x) {
    dft4<INV>(x[0], x[2], x[4], x[6]);
    dft4<INV>(x[1], x[3], x[5], x[7]);
    // twiddles on the odd branch: W8^p, p = 0..3
    x[3] = cmul(x[3], w16<INV>(2));
    x[5] = rot90<INV>(x[5]);
    x[7] = cmul(x[7], w16<INV>(6));
    // z0[p] = x[2p], z1[p] = x[2p+1]  ->  X[p] = z0+z1, X[p+4] = z0-z1
    float2 y[8];
#pragma unroll
    for (int p = 0; p < 4; p++) {
      y[p] = cadd(x[2 * p], x[2 * p + 1]);
      y[p + 4] = csub(x[2 * p], x[2 * p + 1]);
    }
#pragma unroll
    for (int p = 0; p < 8; p++) x[p] = y[p];
  }
};

// R = 16 as 4 x 4: inputs m = j + 4a; X[p + 4k] = DFT4_j( W16^{jp} DFT4_a x[j+4a] )
template <bool INV>
struct Dft<16, INV> {
  static __device__ __forceinline__ void run(float2 *x) {
#pragma unroll
    for (int j = 0; j < 4; j++) dft4<INV>(x[j], x[j + 4], x[j + 8], x[j + 12]);
    // now x[j + 4p] = z_j[p]; twiddle by W16^{j p}
    x[1 + 4] = cmul(x[1 + 4], w16<INV>(1));
    x[1 + 8] = cmul(x[1 + 8], w16<INV>(2));
    x[1 + 12] = cmul(x[1 + 12], w16<INV>(3));
    x[2 + 4] = cmul(x[2 + 4], w16<INV>(2));
    x[2 + 8] = rot90<INV>(x[2 + 8]);
    x[2 + 12] = cmul(x[2 + 12], w16<INV>(6));
    x[3 + 4] = cmul(x[3 + 4], w16<INV>(3));
    x[3 + 8] = cmul(x[3 + 8], w16<INV>(6));
    x[3 + 12] = cmul(x[3 + 12], w16<INV>(9));
    // for each p: X[p + 4k] = DFT4 over j of z_j[p] = x[j + 4p]
    float2 y[16];
#pragma unroll
    for (int p = 0; p < 4; p++) {
      float2 a = x[0 + 4 * p], b = x[1 + 4 * p], c = x[2 + 4 * p], d = x[3 + 4 * p];
      dft4<INV>(a, b, c, d);
      y[p] = a;
      y[p + 4] = b;
      y[p + 8] = c;
      y[p + 12] = d;
    }
#pragma unroll
    for (int p = 0; p < 16; p++) x[p] = y[p];
  }
};

// ---- one in-place pass over a shared-memory buffer of N complex points ---------------------
// L = current sub-transform length, R = radix, S = L/R.  Butterfly b: j = b mod S,
// block = b div S, points at block*L + j + m*S.
// Forward (DIF): DFT_R then multiply output p by W_L^{j p};  inverse (DIT): conj-twiddle then DFT_R^*.
template <int N, int L, int R, bool INV, int NT>
__device__ __forceinline__ void fft_pass(float2 *buf, int tid) {
  constexpr int S = L / R;
#pragma unroll 1
  for (int b = tid; b < N / R; b += NT) {
    const int j = b & (S - 1);
    const int base = (b / S) * L + j;
    float2 x[R];
#pragma unroll
    for (int m = 0; m < R; m++) x[m] = buf[base + m * S];
    float2 w[R];  // W_L^{j p}; built by binary powering from W_L^{j} (few roundings)
    if (S > 1) {
      float sn, cs;
      sincospif(2.0f * (float)j / (float)L, &sn, &cs);
      w[1] = make_float2(cs, -sn);  // forward sign; conjugated below when INV
#pragma unroll
      for (int p = 2; p < R; p++) w[p] = (p & 1) ? cmul(w[p - 1], w[1]) : cmul(w[p / 2], w[p / 2]);
    }
    if (INV && S > 1) {
#pragma unroll
      for (int p = 1; p < R; p++) x[p] = cmulc(x[p], w[p]);
    }
    Dft<R, INV>::run(x);
    if (!INV && S > 1) {
#pragma unroll
      for (int p = 1; p < R; p++) x[p] = cmul(x[p], w[p]);
    }
#pragma unroll
    for (int m = 0; m < R; m++) buf[base + m * S] = x[m];
  }
}

// ---- plans: radices per transform length (product = N) -------------------------------------
template <int LOG2N>
struct Plan;
template <>
struct Plan<11> { static constexpr int R0 = 16, R1 = 16, R2 = 8, R3 = 1; };
template <>
struct Plan<12> { static constexpr int R0 = 16, R1 = 16, R2 = 16, R3 = 1; };
template <>
struct Plan<13> { static constexpr int R0 = 16, R1 = 8, R2 = 8, R3 = 8; };
template <>
struct Plan<14> { static constexpr int R0 = 16, R1 = 16, R2 = 8, R3 = 8; };

// Position of natural-order bin k in the scrambled (DIF output) order.
template <int LOG2N>
__host__ __device__ constexpr int scrambled_pos(int k) {
  int L = 1 << LOG2N, r = 0;
  const int rad[4] = {Plan<LOG2N>::R0, Plan<LOG2N>::R1, Plan<LOG2N>::R2, Plan<LOG2N>::R3};
  for (int i = 0; i < 4; i++) {
    if (rad[i] == 1) break;
    int S = L / rad[i];
    r += (k % rad[i]) * S;
    k /= rad[i];
    L = S;
  }
  return r;
}

// Forward transform, natural order in -> scrambled order out.  Ends with a barrier.
template <int LOG2N, int NT>
__device__ __forceinline__ void fft_forward(float2 *buf, int tid) {
  constexpr int N = 1 << LOG2N;
  using P = Plan<LOG2N>;
  constexpr int L1 = N / P::R0, L2 = L1 / P::R1, L3 = L2 / P::R2;
  fft_pass<N, N, P::R0, false, NT>(buf, tid);
  __syncthreads();
  fft_pass<N, L1, P::R1, false, NT>(buf, tid);
  __syncthreads();
  fft_pass<N, L2, P::R2, false, NT>(buf, tid);
  __syncthreads();
  if constexpr (P::R3 > 1) {
    fft_pass<N, L3, (P::R3 > 1 ? P::R3 : 4), false, NT>(buf, tid);
    __syncthreads();
  }
}

// Inverse (unscaled) transform, scrambled order in -> natural order out.  Ends with a barrier.
template <int LOG2N, int NT>
__device__ __forceinline__ void fft_inverse(float2 *buf, int tid) {
  constexpr int N = 1 << LOG2N;
  using P = Plan<LOG2N>;
  constexpr int L1 = N / P::R0, L2 = L1 / P::R1, L3 = L2 / P::R2;
  if constexpr (P::R3 > 1) {
    fft_pass<N, L3, (P::R3 > 1 ? P::R3 : 4), true, NT>(buf, tid);
    __syncthreads();
  }
  fft_pass<N, L2, P::R2, true, NT>(buf, tid);
  __syncthreads();
  fft_pass<N, L1, P::R1, true, NT>(buf, tid);
  __syncthreads();
  fft_pass<N, N, P::R0, true, NT>(buf, tid);
  __syncthreads();
}

}  // namespace sx
