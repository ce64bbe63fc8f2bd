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
//
// Shared-memory layout: logical element l lives at physical slot swz(l) = l ^ f(l), where f is an
// XOR-linear function of address bits 4..7 that only changes the low 4 bits.  With 8-byte elements
// (16 banks per half-warp) this makes EVERY pass conflict-free, including the last two whose
// natural strides (8 and 1 butterflies apart = 64 and 8 elements) would be 2- and 8-way conflicted.
// Because f is linear and base / m*S never share bits, swz(base + m*S) = (base + m*S) ^ f(base) ^ f(m*S)
// with f(m*S) a compile-time constant: one extra XOR per element.
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

// low-4-bit XOR mask of a logical index (linear in the address bits)
__host__ __device__ constexpr int swz_f(int l) { return ((l >> 4) & 7) ^ ((((l >> 6) ^ (l >> 7)) & 1) << 3); }
__host__ __device__ constexpr int swz(int l) { return l ^ swz_f(l); }

// ---- small DFTs in registers: X[p] = sum_m x[m] e^{-/+ 2 pi i p m / R} -------------------
template <bool INV>
__device__ __forceinline__ void dft4(float2 &x0, float2 &x1, float2 &x2, float2 &x3) {
  float2 t0 = cadd(x0, x2), t1 = csub(x0, x2);
  float2 t2 = cadd(x1, x3), t3 = rot90<INV>(csub(x1, x3));
  x0 = cadd(t0, t2);
  x2 = csub(t0, t2);
  x1 = cadd(t1, t3);
  x3 = csub(t1, t3);
}

#define SX_C1 0.92387953251128674f
#define SX_S1 0.38268343236508977f
#define SX_R2 0.70710678118654752f

// e^{-/+ 2 pi i k / 16} for the k used below
template <bool INV>
__device__ __forceinline__ float2 w16(int k) {  // k is a compile-time constant after inlining
  float c, s;
  switch (k) {
    case 1: c = SX_C1; s = SX_S1; break;
    case 2: c = SX_R2; s = SX_R2; break;
    case 3: c = SX_S1; s = SX_C1; break;
    case 6: c = -SX_R2; s = SX_R2; break;
    case 9: c = -SX_C1; s = -SX_S1; break;
    default: c = 1.f; s = 0.f; break;
  }
  return make_float2(c, INV ? s : -s);
}

template <int R, bool INV>
struct Dft;

// R = 8 as 4 x 2: inputs m = j + 2a (j<2, a<4); X[p + 4k] = DFT2_j( W8^{jp} DFT4_a x[j+2a] )
template <bool INV>
struct Dft<8, INV> {
  static __device__ __forceinline__ void run(float2 *x) {
    dft4<INV>(x[0], x[2], x[4], x[6]);
    dft4<INV>(x[1], x[3], x[5], x[7]);
    x[3] = cmul(x[3], w16<INV>(2));
    x[5] = rot90<INV>(x[5]);
    x[7] = cmul(x[7], w16<INV>(6));
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
    x[1 + 4] = cmul(x[1 + 4], w16<INV>(1));
    x[1 + 8] = cmul(x[1 + 8], w16<INV>(2));
    x[1 + 12] = cmul(x[1 + 12], w16<INV>(3));
    x[2 + 4] = cmul(x[2 + 4], w16<INV>(2));
    x[2 + 8] = rot90<INV>(x[2 + 8]);
    x[2 + 12] = cmul(x[2 + 12], w16<INV>(6));
    x[3 + 4] = cmul(x[3 + 4], w16<INV>(3));
    x[3 + 8] = cmul(x[3 + 8], w16<INV>(6));
    x[3 + 12] = cmul(x[3 + 12], w16<INV>(9));
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

// Inter-stage twiddle table: tw[t] = e^{-2 pi i t / 16384}, t < 2048 (filled by upload_tables()).
// A transform of length N uses entries t * (16384 / N); every pass needs indices < N/8 only.
#define SX_TW_BASE_LOG2 14
#define SX_TW_ENTRIES 2048
static __device__ float2 g_twiddle[SX_TW_ENTRIES];  // this header is compiled into one translation unit only

// ---- one in-place pass over a (swizzled) shared-memory buffer of N complex points ----------
// L = current sub-transform length, R = radix, S = L/R.  Butterfly b: j = b mod S,
// block = b div S, points at block*L + j + m*S.
// Forward (DIF): DFT_R then multiply output p by W_L^{j p};  inverse (DIT): conj-twiddle then DFT_R^*.
template <int LOG2N, int L, int R, bool INV, int NT>
__device__ __forceinline__ void fft_pass(float2 *buf, int tid) {
  constexpr int N = 1 << LOG2N;
  constexpr int S = L / R;
#pragma unroll 1
  for (int b = tid; b < N / R; b += NT) {
    const int j = b & (S - 1);
    const int base = (b / S) * L + j;
    const int fb = swz_f(base);
    float2 x[R];
#pragma unroll
    for (int m = 0; m < R; m++) x[m] = buf[(base + m * S) ^ (fb ^ swz_f(m * S))];
    float2 w[R];  // W_L^{j p}: W_L^{j} from the table, the powers by binary powering (<= 4 roundings)
    if (S > 1) {
      w[1] = __ldg(&g_twiddle[j * ((N / L) << (SX_TW_BASE_LOG2 - LOG2N))]);
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
    for (int m = 0; m < R; m++) buf[(base + m * S) ^ (fb ^ swz_f(m * S))] = x[m];
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

// LOGICAL position of natural-order bin k in the scrambled (DIF output) order; its shared-memory
// slot is swz() of this.
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
  fft_pass<LOG2N, N, P::R0, false, NT>(buf, tid);
  __syncthreads();
  fft_pass<LOG2N, L1, P::R1, false, NT>(buf, tid);
  __syncthreads();
  fft_pass<LOG2N, L2, P::R2, false, NT>(buf, tid);
  __syncthreads();
  if constexpr (P::R3 > 1) {
    fft_pass<LOG2N, L3, P::R3, false, NT>(buf, tid);
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
    fft_pass<LOG2N, L3, P::R3, true, NT>(buf, tid);
    __syncthreads();
  }
  fft_pass<LOG2N, L2, P::R2, true, NT>(buf, tid);
  __syncthreads();
  fft_pass<LOG2N, L1, P::R1, true, NT>(buf, tid);
  __syncthreads();
  fft_pass<LOG2N, N, P::R0, true, NT>(buf, tid);
  __syncthreads();
}

}  // namespace sx
