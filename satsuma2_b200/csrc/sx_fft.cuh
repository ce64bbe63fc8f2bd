// In-shared-memory complex FFTs for the cross-correlation kernels (sm_100a).
//
// Replaces extern/RealFFT (FFTReal<float>::do_fft / do_ifft, FFTReal.hpp:145-245): the
// reference runs 12 radix-2 real transforms of N points per strand-pair.  Here
//   * two real channels ride one complex transform (A + iC, G + iT);
//   * the N-point transform is split once by decimation in frequency into two INDEPENDENT
//     H = N/2 point transforms: even bins = FFT_H(z[n] + z[n+H]), odd bins = FFT_H((z[n] - z[n+H]) w_N^n).
//     A chunk never exceeds H bases in the default configuration, so z[n+H] = 0 and the split costs
//     nothing: the encoder writes z and z*w_N^n side by side.  The spectrum is stored as
//     [even half | odd half]; the inverse runs the two H-point inverses and one radix-2 combine;
//   * each H-point transform is an in-place decimation-in-frequency network that leaves its bins in
//     mixed-radix digit-reversed order, the inverse is its exact adjoint (decimation in time) that
//     consumes that order -- no reordering pass exists: the spectral product in between is
//     element-wise (plus a bin <-> N - bin pairing that stays inside each half, see scrambled_neg).
//
// One pass = every thread pulls R points (stride S) from shared memory into registers,
// does an R-point DFT, applies the inter-stage twiddles and writes back in place.
//
// Shared-memory layout: logical element l lives at physical slot swz(l) = l ^ f(l), where f is an
// XOR-linear function of address bits 4..7 that only changes the low 4 bits.  With 8-byte elements
// (16 banks per half-warp) this makes EVERY pass of every plan conflict-free (tools/fft_bank_check.py).
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

// Inter-stage twiddle table: tw[t] = e^{-2 pi i t / 32768}, t < 4096 (filled by upload_tables()).
// A transform of length L uses entries j * (32768 / L); every pass needs angles below 2 pi / 8 only.
#define SX_TW_BASE_LOG2 15
#define SX_TW_ENTRIES 4096
static __device__ float2 g_twiddle[SX_TW_ENTRIES];  // this header is compiled into one translation unit only

// ---- one in-place pass over a (swizzled) shared-memory buffer of NPTS complex points --------
// L = current sub-transform length, R = radix, S = L/R.  Butterfly b: j = b mod S,
// block = b div S, points at block*L + j + m*S.  NPTS is a multiple of L (N for the two halves in
// one CTA, H when a CTA of a cluster holds one half).
// Forward (DIF): DFT_R then multiply output p by W_L^{j p};  inverse (DIT): conj-twiddle then DFT_R^*.
// tw: table of this pass in shared memory, tw[j] = W_L^j for j < S (see TwTables), or nullptr = read the global
// table.  The inverse applies the twiddles BEFORE the butterfly, so the latency of that load is exposed: its
// kernels stage the few hundred entries they need in shared memory; the forward side hides it behind the math.
template <int NPTS, int L, int R, bool INV, int NT, bool SMEM_TW = false>
__device__ __forceinline__ void fft_pass(float2 *buf, int tid, const float2 *tw = nullptr) {
  constexpr int S = L / R;
  constexpr int LOG2L = __builtin_ctz(L);
#pragma unroll 1
  for (int b = tid; b < NPTS / R; b += NT) {
    const int j = b & (S - 1);
    const int base = (b / S) * L + j;
    const int fb = swz_f(base);
    float2 x[R];
#pragma unroll
    for (int m = 0; m < R; m++) x[m] = buf[(base + m * S) ^ (fb ^ swz_f(m * S))];
    float2 w[R];  // W_L^{j p}: W_L^{j} from the table, the powers by binary powering (<= 4 roundings)
    if (S > 1) {
      w[1] = SMEM_TW ? tw[j] : __ldg(&g_twiddle[j << (SX_TW_BASE_LOG2 - LOG2L)]);
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

// barrier over the NT/2 threads that own one half of a grouped transform (ids 1 and 2; 0 is __syncthreads)
template <int NT>
__device__ __forceinline__ void group_barrier(int tid) {
  if (tid < NT / 2)  // literal barrier ids: a register id would make ptxas reserve all 16 barriers for the CTA
    asm volatile("bar.sync 1, %0;" ::"n"(NT / 2) : "memory");
  else
    asm volatile("bar.sync 2, %0;" ::"n"(NT / 2) : "memory");
}

// ---- plans: radices of the H = N/2 point transforms (product = H) ---------------------------
template <int LOG2N>
struct Plan;
template <>
struct Plan<11> { static constexpr int R0 = 16, R1 = 8, R2 = 8, R3 = 1; };
template <>
struct Plan<12> { static constexpr int R0 = 16, R1 = 16, R2 = 8, R3 = 1; };
template <>
struct Plan<13> { static constexpr int R0 = 16, R1 = 16, R2 = 16, R3 = 1; };
template <>
struct Plan<14> { static constexpr int R0 = 16, R1 = 8, R2 = 8, R3 = 8; };
template <>
struct Plan<15> { static constexpr int R0 = 16, R1 = 16, R2 = 8, R3 = 8; };

// LOGICAL position (inside its half) of bin m of an H-point transform in the scrambled (DIF output)
// order; the shared-memory slot is swz() of it.  Natural bin k of the N-point transform lives in
// half (k & 1) at scrambled_pos(k >> 1).
template <int LOG2N>
__host__ __device__ constexpr int scrambled_pos(int m) {
  int L = 1 << (LOG2N - 1), r = 0;
  const int rad[4] = {Plan<LOG2N>::R0, Plan<LOG2N>::R1, Plan<LOG2N>::R2, Plan<LOG2N>::R3};
  for (int i = 0; i < 4; i++) {
    if (rad[i] == 1) break;
    int S = L / rad[i];
    r += (m % rad[i]) * S;
    m /= rad[i];
    L = S;
  }
  return r;
}
// inverse of scrambled_pos
template <int LOG2N>
__host__ __device__ constexpr int natural_bin(int r) {
  int L = 1 << (LOG2N - 1), m = 0, mul = 1;
  const int rad[4] = {Plan<LOG2N>::R0, Plan<LOG2N>::R1, Plan<LOG2N>::R2, Plan<LOG2N>::R3};
  for (int i = 0; i < 4; i++) {
    if (rad[i] == 1) break;
    int S = L / rad[i];
    m += (r / S) * mul;
    r %= S;
    mul *= rad[i];
    L = S;
  }
  return m;
}
// physical slot (index into a [even half | odd half] spectrum or FFT buffer) of natural bin k, 0 <= k < N
template <int LOG2N>
__host__ __device__ constexpr int bin_slot(int k) {
  return (k & 1) * (1 << (LOG2N - 1)) + swz(scrambled_pos<LOG2N>(k >> 1));
}
// radix of the last pass = number of consecutive scrambled positions that differ only in the top digit of m
template <int LOG2N>
__host__ __device__ constexpr int last_radix() {
  return Plan<LOG2N>::R3 > 1 ? Plan<LOG2N>::R3 : Plan<LOG2N>::R2;
}

// The H-point transforms of a buffer of NPTS points (NPTS = N: both halves; NPTS = H: one half),
// natural order in -> scrambled order out.  Ends with a barrier.
template <int LOG2N, int NPTS, int NT>
__device__ __forceinline__ void fft_forward_halves(float2 *buf, int tid) {
  constexpr int H = 1 << (LOG2N - 1);
  using P = Plan<LOG2N>;
  constexpr int L1 = H / P::R0, L2 = L1 / P::R1, L3 = L2 / P::R2;
  fft_pass<NPTS, H, P::R0, false, NT>(buf, tid);
  __syncthreads();
  fft_pass<NPTS, L1, P::R1, false, NT>(buf, tid);
  __syncthreads();
  fft_pass<NPTS, L2, P::R2, false, NT>(buf, tid);
  __syncthreads();
  if constexpr (P::R3 > 1) {
    fft_pass<NPTS, L3, P::R3, false, NT>(buf, tid);
    __syncthreads();
  }
}

// Shared-memory twiddle tables of the passes of an H-point transform: pass i (sub-transform length L_i, stride
// S_i = L_i / R_i) owns S_i entries W_{L_i}^j at offset OFF_i; passes with S = 1 need none.
template <int LOG2N>
struct TwTables {
  using P = Plan<LOG2N>;
  static constexpr int H = 1 << (LOG2N - 1);
  static constexpr int L0 = H, L1 = L0 / P::R0, L2 = L1 / P::R1, L3 = L2 / P::R2;
  static constexpr int S0 = L0 / P::R0, S1 = L1 / P::R1, S2 = L2 / P::R2, S3 = P::R3 > 1 ? L3 / P::R3 : 1;
  static constexpr int OFF0 = 0, OFF1 = OFF0 + (S0 > 1 ? S0 : 0), OFF2 = OFF1 + (S1 > 1 ? S1 : 0),
                       OFF3 = OFF2 + (S2 > 1 ? S2 : 0), TOTAL = OFF3 + (S3 > 1 ? S3 : 0);
  template <int NT>
  static __device__ __forceinline__ void load(float2 *tw, int tid) {  // the caller synchronises before use
    if (S0 > 1) for (int j = tid; j < S0; j += NT) tw[OFF0 + j] = __ldg(&g_twiddle[j << (SX_TW_BASE_LOG2 - __builtin_ctz(L0))]);
    if (S1 > 1) for (int j = tid; j < S1; j += NT) tw[OFF1 + j] = __ldg(&g_twiddle[j << (SX_TW_BASE_LOG2 - __builtin_ctz(L1))]);
    if (S2 > 1) for (int j = tid; j < S2; j += NT) tw[OFF2 + j] = __ldg(&g_twiddle[j << (SX_TW_BASE_LOG2 - __builtin_ctz(L2))]);
    if (S3 > 1) for (int j = tid; j < S3; j += NT) tw[OFF3 + j] = __ldg(&g_twiddle[j << (SX_TW_BASE_LOG2 - __builtin_ctz(L3))]);
  }
};

// Inverse (unscaled) H-point transforms, scrambled order in -> natural order out.  tw: TwTables<LOG2N> in shared
// memory.  Ends with a barrier.
template <int LOG2N, int NPTS, int NT>
__device__ __forceinline__ void fft_inverse_halves(float2 *buf, int tid, const float2 *tw) {
  constexpr int H = 1 << (LOG2N - 1);
  using P = Plan<LOG2N>;
  using T = TwTables<LOG2N>;
  constexpr int L1 = H / P::R0, L2 = L1 / P::R1, L3 = L2 / P::R2;
  // A buffer with both halves (NPTS = N) is worked on by two thread groups, threads [0, NT/2) on the first half and
  // the rest on the second; the halves never exchange data, so each group synchronises on its own (group_barrier).
  // (Measured: 2 % on the pair kernel; the same split made the forward transforms of the encoder slower.)
  if constexpr (NPTS == 2 * H) {
    constexpr int NG = NT / 2;
    float2 *hb = buf + (tid >= NG ? H : 0);
    const int t = tid >= NG ? tid - NG : tid;
    if constexpr (P::R3 > 1) {
      fft_pass<H, L3, P::R3, true, NG, true>(hb, t, tw + T::OFF3);
      group_barrier<NT>(tid);
    }
    fft_pass<H, L2, P::R2, true, NG, true>(hb, t, tw + T::OFF2);
    group_barrier<NT>(tid);
    fft_pass<H, L1, P::R1, true, NG, true>(hb, t, tw + T::OFF1);
    group_barrier<NT>(tid);
    fft_pass<H, H, P::R0, true, NG, true>(hb, t, tw + T::OFF0);
  } else {
    if constexpr (P::R3 > 1) {
      fft_pass<NPTS, L3, P::R3, true, NT, true>(buf, tid, tw + T::OFF3);
      __syncthreads();
    }
    fft_pass<NPTS, L2, P::R2, true, NT, true>(buf, tid, tw + T::OFF2);
    __syncthreads();
    fft_pass<NPTS, L1, P::R1, true, NT, true>(buf, tid, tw + T::OFF1);
    __syncthreads();
    fft_pass<NPTS, H, P::R0, true, NT, true>(buf, tid, tw + T::OFF0);
  }
  __syncthreads();
}

}  // namespace sx
