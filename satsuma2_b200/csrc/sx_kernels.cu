// Hand-written sm_100a kernels for Satsuma2's chunk-pair cross-correlation path.
//
//   K1 encode_fft_kernel      (a)+(b)  DNA -> 4-channel entropy-weighted signal -> spectra
//        replaces CCSignal::SetSequence / ComputeEntropy / SeqToPCM (analysis/CrossCorr.cc:35-134,
//        179-206), DNAVector::ReverseComplement (analysis/DNAVector.cc:482-521) and the forward
//        FFTReal::do_fft calls of CrossCorrelation::DoOne (CrossCorr.cc:471-474)
//   K2 xcorr_pair_kernel (both strands of a chunk pair from the forward query spectrum) /
//      xcorr_findtop_kernel (one strand-pair)   (c)+(d)  spectral product (+ reference quirk bins),
//        channel sum, one inverse FFT, half rotation, RMS envelope, threshold, ordered compaction
//        replaces CrossCorrelation::DoOne / CrossCorrelate (CrossCorr.cc:386-507) and
//        SeqAnalyzer::FindTop (CrossCorr.cc:878-944)
//   K3 scan_score_kernel[_generic] (e)  diagonal sliding-window scan + match probability
//        replaces SeqAnalyzer::MatchUp / DoOne (CrossCorr.cc:583-605, 667-724),
//        GetMatchProbabilityEx (analysis/AlignProbability.cc:62-127), ProbTable lookup
//        (analysis/ProbTable.cc:58-74, 105-140) and the filter of FilterMatches
//        (analysis/HomologyByXCorrSlave.cc:168-219)
//
//   K0 encode_prep_kernel     one warp per chunk signal ahead of K1: validation, 2-bit planes, entropy weights, means
//   pair_fused_kernel         K1 + K2 in one kernel for chunk pairs whose spectra nobody else needs
//        (sx_config::fuse_pairs; the spectra never reach HBM)
//
//   N = 32768 (more complex points than one CTA's shared memory holds): encode_fft_half_kernel -- two CTAs per
//        transform, one per half of the radix-2 split; xcorr_cluster_kernel -- a cluster of two CTAs per strand-pair,
//        halves combined through distributed shared memory (xcorr_half_kernel + combine_findtop_kernel: the same
//        through an HBM scratch buffer, kept for comparison)
//
// Pure A/C/G/T chunks are transformed in THREE-CHANNEL form (sx_kernels.h): T = -(A + C + G) sample by sample, the G
// channels of two chunks share one complex transform; the correlation kernels recover the per-channel spectra from
// bin pairs (k, N - k).
//
// No tensor cores (nothing here is a dense contraction), no cuFFT, no CPU fallback.
#include "sx_kernels.h"
#include "sx_fft.cuh"

#include <math.h>
#include <string.h>

#include <cooperative_groups.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

namespace sx {

// ------------------------------------------------------------------------------------------------
// Lookup tables.  The IUPAC codec (analysis/DNAVector.cc:13-58): fraction of A,C,G,T per letter,
// complement letter; match scores (int)(100*DNA_EqualAmb + 0.5) (CrossCorr.cc:547-553).
// Bytes >= 128 are undefined behaviour in the reference (negative table index); here they are
// mapped to NUL when a chunk is loaded, so every table has 128 rows.
// ------------------------------------------------------------------------------------------------
__constant__ uint16_t c_fcode[128];  // 4 x 4-bit value codes (A,C,G,T): 0:0 1:1 2:1/2 3:1/3 4:1/4
__constant__ uint8_t c_comp[128];    // complement letter (0 for unknown)
__constant__ uint8_t c_base2[128];   // 0..3 for A,C,G,T; 4 otherwise; |8 for B,V,H,D (thirds)
__device__ uint8_t g_score[128 * 128];

static double h_frac[128 * 4];
static uint8_t h_comp[128];
static uint8_t h_score[128 * 128];
static uint16_t h_fcode[128];
static uint8_t h_base2[128];
static bool h_tables_ready = false;

static void host_put(int ch, int a, int c, int g, int t, int comp) {
  static const double val[5] = {0., 1., 0.5, 1. / 3., 0.25};
  h_fcode[ch] = (uint16_t)(a | (c << 4) | (g << 8) | (t << 12));
  h_frac[ch * 4 + 0] = val[a];
  h_frac[ch * 4 + 1] = val[c];
  h_frac[ch * 4 + 2] = val[g];
  h_frac[ch * 4 + 3] = val[t];
  h_comp[ch] = (uint8_t)comp;
}

static void build_host_tables() {
  if (h_tables_ready) return;
  memset(h_frac, 0, sizeof(h_frac));
  memset(h_comp, 0, sizeof(h_comp));
  memset(h_fcode, 0, sizeof(h_fcode));
  host_put('A', 1, 0, 0, 0, 'T');
  host_put('C', 0, 1, 0, 0, 'G');
  host_put('G', 0, 0, 1, 0, 'C');
  host_put('T', 0, 0, 0, 1, 'A');
  host_put('K', 0, 0, 2, 2, 'M');
  host_put('M', 2, 2, 0, 0, 'K');
  host_put('R', 2, 0, 2, 0, 'Y');
  host_put('Y', 0, 2, 0, 2, 'R');
  host_put('S', 0, 2, 2, 0, 'S');
  host_put('W', 2, 0, 0, 2, 'W');
  host_put('B', 0, 3, 3, 3, 'V');
  host_put('V', 3, 3, 3, 0, 'B');
  host_put('H', 3, 3, 0, 3, 'D');
  host_put('D', 3, 0, 3, 3, 'H');
  host_put('-', 0, 0, 0, 0, '-');
  host_put('N', 4, 4, 4, 4, 'N');
  host_put('X', 4, 4, 4, 4, 'X');
  for (int i = 0; i < 128; i++) {
    h_base2[i] = 4;
    if (i == 'B' || i == 'V' || i == 'H' || i == 'D') h_base2[i] = 4 | 8;
  }
  h_base2['A'] = 0;
  h_base2['C'] = 1;
  h_base2['G'] = 2;
  h_base2['T'] = 3;
  for (int a = 0; a < 128; a++)
    for (int b = 0; b < 128; b++) {
      double pa = h_frac[a * 4 + 0] * h_frac[b * 4 + 0];
      double pc = h_frac[a * 4 + 1] * h_frac[b * 4 + 1];
      double pg = h_frac[a * 4 + 2] * h_frac[b * 4 + 2];
      double pt = h_frac[a * 4 + 3] * h_frac[b * 4 + 3];
      double dot = pa + pc + pg + pt;
      double amb = (a == b && a != 'N') ? 1. : dot;
      h_score[a * 128 + b] = (uint8_t)(int)(amb * 100. + 0.5);
    }
  h_tables_ready = true;
}

const double *host_frac_table() { build_host_tables(); return h_frac; }
const uint8_t *host_comp_table() { build_host_tables(); return h_comp; }
const uint8_t *host_score_table() { build_host_tables(); return h_score; }

cudaError_t upload_tables() {
  build_host_tables();
  cudaError_t e;
  if ((e = cudaMemcpyToSymbol(c_fcode, h_fcode, sizeof(h_fcode))) != cudaSuccess) return e;
  if ((e = cudaMemcpyToSymbol(c_comp, h_comp, sizeof(h_comp))) != cudaSuccess) return e;
  if ((e = cudaMemcpyToSymbol(c_base2, h_base2, sizeof(h_base2))) != cudaSuccess) return e;
  if ((e = cudaMemcpyToSymbol(g_score, h_score, sizeof(h_score))) != cudaSuccess) return e;
  static float2 h_tw[SX_TW_ENTRIES];
  for (int t = 0; t < SX_TW_ENTRIES; t++) {
    const double a = -2.0 * 3.14159265358979323846 * (double)t / (double)(1 << SX_TW_BASE_LOG2);
    h_tw[t] = make_float2((float)cos(a), (float)sin(a));
  }
  if ((e = cudaMemcpyToSymbol(g_twiddle, h_tw, sizeof(h_tw))) != cudaSuccess) return e;
  return cudaSuccess;
}

bool log2n_supported(int log2n) { return log2n >= 11 && log2n <= 15; }
bool log2n_split(int log2n) { return log2n >= 15; }

void fill_wn_table(int log2n, float2 *out) {
  const int N = 1 << log2n;
  for (int n = 0; n < N / 2; n++) {
    const double a = -2.0 * 3.14159265358979323846 * (double)n / (double)N;
    out[n] = make_float2((float)cos(a), (float)sin(a));
  }
}
size_t slot_spec_elems(int log2n) { return (size_t)2 << log2n; }

// ---- the reference's oscillator twiddles (sx_kernels.h) ---------------------------------------------------------
// w~[i] = OscSinCos after i steps of angle pi / M from (1, 0), all in float, i < M / 2
static void osc_twiddles(int M, std::vector<double> &wr, std::vector<double> &wi) {
  const float sc = (float)cos(3.14159265358979323846 / (double)M), ss = (float)sin(3.14159265358979323846 / (double)M);
  volatile float pc = 1.f, ps = 0.f;  // volatile: every product and sum rounded to float on its own (no FMA, no excess precision)
  wr.assign((size_t)M / 2, 1.);
  wi.assign((size_t)M / 2, 0.);
  for (int i = 1; i < M / 2; i++) {
    const float oc = pc, os = ps;
    volatile float a = oc * sc, b = os * ss, c2 = oc * ss, d = os * sc;
    pc = a - b;
    ps = c2 + d;
    wr[(size_t)i] = (double)pc;
    wi[(size_t)i] = (double)ps;
  }
}
// ratio (twiddle the reference effectively applies to bin i < M when it builds length 2M) / (exact twiddle), in the
// reference's convention e^{+j pi i / M}: the oscillator value for i < M/2, exact at i = 0 and i = M/2 (FFTReal.hpp:
// 622-626), -conj(value at M - i) above (those bins are produced as conjugates of the lower ones)
static void drift_ratio(const std::vector<double> &wr, const std::vector<double> &wi, int M, int i, double *re, double *im) {
  if (i == 0 || 2 * i == M) {
    *re = 1.;
    *im = 0.;
    return;
  }
  double er, ei;
  if (2 * i < M) {
    er = wr[(size_t)i];
    ei = wi[(size_t)i];
  } else {
    er = -wr[(size_t)(M - i)];
    ei = wi[(size_t)(M - i)];
  }
  const double ang = 3.14159265358979323846 * (double)i / (double)M, cr = cos(ang), ci = sin(ang);
  *re = er * cr + ei * ci;  // eff * conj(exact)
  *im = ei * cr - er * ci;
}
size_t drift_table_elems(int log2n) { return log2n == 15 ? (size_t)3 << (log2n - 2) : 0; }
void fill_drift_table(int log2n, float2 *out) {
  if (log2n != 15) return;
  const int N = 1 << log2n, H = N / 2, Q = N / 4;
  std::vector<double> w14r, w14i, w13r, w13i;
  osc_twiddles(H, w14r, w14i);      // the pass that builds length N from two of length H
  osc_twiddles(H / 2, w13r, w13i);  // the pass below it
  for (int k = 0; k < Q; k++) {
    double re, im;
    // standard-DFT convention (negative exponent) = conjugates of the reference-convention ratios
    drift_ratio(w14r, w14i, H, k, &re, &im);
    out[k] = make_float2((float)re, (float)-im);              // rho0
    drift_ratio(w14r, w14i, H, k + Q, &re, &im);
    out[Q + k] = make_float2((float)re, (float)-im);          // rho1
    drift_ratio(w13r, w13i, H / 2, k, &re, &im);
    out[2 * Q + k] = make_float2((float)(0.5 * (re - 1.)), (float)(-0.5 * im));  // phi
  }
}

// Entropy terms of CCSignal::ComputeEntropy (analysis/CrossCorr.cc:28-33, 70-88) for integer base counts:
// out[k * (WIN + 1) + c] = p < 0.001 ? 0 : p * log(p) / 0.69314718056 with p = c / k, for window lengths
// k = 1 .. WIN (WIN = N / 512) and counts c = 0 .. k.  Built with the host's libm, like the reference.
size_t ent_table_elems(int log2n) {
  const size_t win = ((size_t)1 << log2n) / 512;
  return (win + 1) * (win + 1);
}
void fill_ent_table(int log2n, double *out) {
  const int win = (1 << log2n) / 512;
  for (int k = 0; k <= win; k++)
    for (int c = 0; c <= win; c++) {
      double e = 0.;
      if (k > 0 && c <= k) {
        const double p = (double)c / (double)k;
        e = (p < 0.001) ? 0.0 : p * log(p) / 0.69314718056;
      }
      out[(size_t)k * (win + 1) + c] = e;
    }
}

// fraction of channel c (0..3) for a table entry; exact doubles {0,1,1/2,1/3,1/4}
__device__ __forceinline__ double frac_of(uint32_t fcode, int c) {
  const uint32_t v = (fcode >> (4 * c)) & 15u;
  // 1/3 must be the same double the reference computes (1./3.)
  return v == 0 ? 0.0 : v == 1 ? 1.0 : v == 2 ? 0.5 : v == 3 ? (1.0 / 3.0) : 0.25;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- bulk asynchronous copy shared -> global (the TMA engine without a tensor map) ------------------------
__device__ __forceinline__ void bulk_store_fence() {  // generic-proxy writes to shared memory -> visible to the copy engine
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void bulk_store(void *gdst, const void *ssrc, int bytes) {  // 16-byte aligned, multiple of 16
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
               "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// =================================================================================================
// K1: encode + forward transform.  One CTA per chunk signal.
// =================================================================================================
// e^{-2 pi i j / N} for small j as a compile-time constant (Taylor series in double: the angle is
// below 2 pi * 64 / 2048, six terms are exact to double rounding)
__host__ __device__ constexpr double cx_cos_small(double a) {
  const double a2 = a * a;
  return 1. - a2 / 2. * (1. - a2 / 12. * (1. - a2 / 30. * (1. - a2 / 56. * (1. - a2 / 90. * (1. - a2 / 132.)))));
}
__host__ __device__ constexpr double cx_sin_small(double a) {
  const double a2 = a * a;
  return a * (1. - a2 / 6. * (1. - a2 / 20. * (1. - a2 / 42. * (1. - a2 / 72. * (1. - a2 / 110. * (1. - a2 / 156.))))));
}

// Steps 1-2 of the encoder, shared by the whole-signal and the half-signal kernels: stage the raw chunk,
// orient (reverse-complement) and sanitise it into sb[], write the 2-bit planes / oriented bytes (and the
// other orientation's, see SigDesc::rc_slot1) when `side` is set, entropy weights per window into went[],
// channel means into s_off[].  Returns the slot flags.  `raw` may alias any buffer that is free until the
// signal is generated.  Ends with a barrier.
template <int LOG2N, int NT>
__device__ __forceinline__ int encode_prepare(const SigDesc &sd, const Slots &ws, uint8_t *raw, uint8_t *sb, float *went,
                                              uint16_t *s_fcode, uint8_t *s_comp, uint8_t *s_base2,
                                              double (*s_red)[NT / 32], double *s_off, int *s_flags_p, bool side) {
  constexpr int N = 1 << LOG2N, NW = N / 32, WIN = N / 512, NWARP = NT / 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int len = sd.len;
  const uint8_t *__restrict__ src = sd.src;
  int &s_flags = *s_flags_p;
  if (tid < 128) {
    s_fcode[tid] = c_fcode[tid];
    s_comp[tid] = c_comp[tid];
    s_base2[tid] = c_base2[tid];
  }
  if (tid == 0) s_flags = 0;
  __syncthreads();
  const int len32 = (len + 31) & ~31, nwords = len32 >> 5;
  // plane words of this signal, [2][NW], kept in shared memory inside the (still unused) FFT buffer behind the
  // staged chunk: the reverse-complement derivation and the window counts below read them
  uint32_t *s_pl = reinterpret_cast<uint32_t *>(raw + N);

  // ---- 1 (fast). forward strand, pure A/C/G/T: 16 bases per thread straight from HBM, no per-base work.
  // Four bases per 32-bit word: code = ((b >> 1) & 3) ^ ((b >> 2) & 1) maps A,C,G,T -> 0,1,2,3; the word is
  // valid iff rebuilding the letters from the codes gives it back; the code bits of four bytes are gathered
  // into a nibble by one multiplication.  Anything else (a letter outside A/C/G/T, a reverse-strand signal)
  // takes the byte-wise path below, which is the general statement of the same thing.
  bool fast_done = false;
  if (sd.strand == 0) {
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 3u);
    const uint32_t *s32 = reinterpret_cast<const uint32_t *>(src - mis);  // chunk stores are padded by 32 bytes (the funnel below reads up to 20 bytes past the last base)
    const bool al16 = (reinterpret_cast<uintptr_t>(src) & 15u) == 0;
    uint32_t bad = 0;
    for (int t = tid; t < (len32 >> 4); t += NT) {
      uint32_t w4[4] = {0u, 0u, 0u, 0u};
      if (t * 16 < len) {
        if (al16) {
          const uint4 v = __ldg(reinterpret_cast<const uint4 *>(src) + t);
          w4[0] = v.x; w4[1] = v.y; w4[2] = v.z; w4[3] = v.w;
        } else {
          uint32_t x[5];
#pragma unroll
          for (int i = 0; i < 5; i++) x[i] = __ldg(s32 + t * 4 + i);
#pragma unroll
          for (int i = 0; i < 4; i++) w4[i] = __funnelshift_r(x[i], x[i + 1], 8 * mis);
        }
      }
      uint32_t lo16 = 0, hi16 = 0;
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int nv = min(max(len - (t * 16 + i * 4), 0), 4);  // bases of this word inside the chunk
        const uint32_t bm = nv == 0 ? 0u : (0xffffffffu >> (8 * (4 - nv)));
        const uint32_t w = w4[i] & bm;
        w4[i] = w;
        const uint32_t c = ((w >> 1) & 0x03030303u) ^ ((w >> 2) & 0x01010101u);
        const uint32_t c0 = c & 0x01010101u & bm, c1 = (c >> 1) & 0x01010101u & bm, c01 = c0 & c1;
        const uint32_t letters = (0x41414141u & bm) + 2u * c0 + 6u * c1 + 11u * c01;  // A, C = A+2, G = A+6, T = A+19
        bad |= letters ^ w;
        lo16 |= ((c0 * 0x01020408u) >> 24) << (4 * i);
        hi16 |= ((c1 * 0x01020408u) >> 24) << (4 * i);
      }
      reinterpret_cast<uint4 *>(sb)[t] = make_uint4(w4[0], w4[1], w4[2], w4[3]);
      reinterpret_cast<uint16_t *>(s_pl)[t] = (uint16_t)lo16;
      reinterpret_cast<uint16_t *>(s_pl + NW)[t] = (uint16_t)hi16;
    }
    fast_done = __syncthreads_or(bad != 0u) == 0;
    if (fast_done && side) {
      uint32_t *planes = ws.planes + (size_t)sd.slot * 2 * NW;
      for (int w = tid; w < NW; w += NT) {
        planes[w] = w < nwords ? s_pl[w] : 0u;
        planes[NW + w] = w < nwords ? s_pl[NW + w] : 0u;
      }
    }
  }
  if (!fast_done) {
    // ---- 1a. stage the raw chunk in shared memory (the FFT buffer is free until step 3) -----------
    if ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
      const uint4 *s16 = reinterpret_cast<const uint4 *>(src);
      for (int i = tid; i < (len + 15) / 16; i += NT) reinterpret_cast<uint4 *>(raw)[i] = __ldg(s16 + i);
    } else {
      for (int i = tid; i < len; i += NT) raw[i] = src[i];
    }
    __syncthreads();
  }

  // ---- 1b. orient (reverse-complement), sanitise, 2-bit planes, byte by byte --------------------------
  // Only the first len32 positions are visited; the rest of the planes is zero-filled with plain stores.
  auto orient_round = [&](int strand, int slot, bool to_hbm) {
    int myflags = 0;
    uint32_t *planes = ws.planes + (size_t)slot * 2 * NW;
    for (int k0 = warp * 32; k0 < len32; k0 += NT) {
      const int k = k0 + lane;
      uint32_t b = 0, code = 4;
      if (k < len) {
        b = strand ? raw[len - 1 - k] : raw[k];
        if (b >= 128u) b = 0;
        if (strand) b = s_comp[b];
        code = s_base2[b];
        if (code & 4u) myflags |= SLOT_NONACGT;
        if (code & 8u) myflags |= 2;
      }
      sb[k] = (uint8_t)b;
      const uint32_t lo = __ballot_sync(0xffffffffu, (code & 5u) == 1u);  // C or T -> bit0
      const uint32_t hi = __ballot_sync(0xffffffffu, (code & 6u) == 2u);  // G or T -> bit1
      if (lane == 0) {
        s_pl[k0 >> 5] = lo;
        s_pl[NW + (k0 >> 5)] = hi;
      }
    }
    if (myflags) atomicOr(&s_flags, myflags);
    __syncthreads();
    if (to_hbm) {
      for (int w = tid; w < NW; w += NT) {
        planes[w] = w < nwords ? s_pl[w] : 0u;
        planes[NW + w] = w < nwords ? s_pl[NW + w] : 0u;
      }
      // The oriented bases go to HBM only when the chunk holds a letter other than A/C/G/T: the byte-wise scan
      // kernel rebuilds pure chunks from their planes (SlotMeta::flags tells which).  16 bytes per thread.
      if (s_flags & SLOT_NONACGT) {
        uint4 *gb = reinterpret_cast<uint4 *>(ws.bytes + (size_t)slot * N);
        for (int i = tid; i < len32 / 16; i += NT) gb[i] = reinterpret_cast<const uint4 *>(sb)[i];
      }
    }
  };
  if (!fast_done) orient_round(sd.strand, sd.slot, side);
  const int flags = s_flags;
  if (sd.rc_slot1) {
    // The OTHER orientation's planes / bytes / meta go to slot rc_slot1 - 1: the host derives that strand's
    // correlation from this signal's spectrum, it gets no transform of its own.
    const int oslot = sd.rc_slot1 - 1;
    if (!(flags & SLOT_NONACGT)) {
      // pure A/C/G/T: reverse complement = complemented codes in reverse order, i.e. bit-reversed inverted
      // plane words: position k of the other strand is position len-1-k of this one.  With len-1 = 32 q + r
      // word j is (R[q-j] >> (31-r)) | (R[q-j-1] << (r+1)), R[x] = brev(~plane[x]); bits past the end fall
      // out at the bottom of R[q] and are cut at the top of word q.  The bit-parallel scan kernel needs no bytes.
      if (side) {
        uint32_t *planes = ws.planes + (size_t)oslot * 2 * NW;
        const int q = (len - 1) >> 5, r = (len - 1) & 31;
        for (int j = tid; j < NW; j += NT) {
          uint32_t lo = 0u, hi = 0u;
          if (j <= q && len > 0) {
            const uint32_t l0 = __brev(~s_pl[q - j]), h0 = __brev(~s_pl[NW + q - j]);
            const uint32_t l1 = j < q ? __brev(~s_pl[q - j - 1]) : 0u, h1 = j < q ? __brev(~s_pl[NW + q - j - 1]) : 0u;
            lo = __funnelshift_r(l0, l1, 31 - r);
            hi = __funnelshift_r(h0, h1, 31 - r);
            if (j == q) {
              const uint32_t vm = r == 31 ? 0xffffffffu : ((2u << r) - 1u);
              lo &= vm;
              hi &= vm;
            }
          }
          planes[j] = lo;
          planes[NW + j] = hi;
        }
      }
    } else {
      // IUPAC / unknown letters: the byte-wise scan kernel needs the other strand's bytes; orient it in full,
      // then restore this signal's own orientation in sb
      __syncthreads();
      orient_round(sd.strand ^ 1, oslot, side);
      __syncthreads();
      orient_round(sd.strand, sd.slot, false);
    }
    if (tid == 0 && side) {
      SlotMeta m;
      m.len = len;
      m.flags = flags & SLOT_NONACGT;  // same letters in both orientations (complement keeps the class)
      m.q_re = m.q_im = m.q_nyq = 0.f;
      m.zmode = ZM_FOUR;
      m.zslot = 0;
      m.pad = 0;
      ws.meta[oslot] = m;
    }
  }
  __syncthreads();

  // ---- 2. window sums, entropy weights (ComputeEntropy) and channel means (SeqToPCM) ------------
  double tot[4] = {0., 0., 0., 0.};
  const int nwin = (len + WIN - 1) / WIN;
  for (int w = tid; w < nwin; w += NT) {
    double s4[4] = {0., 0., 0., 0.};
    const int i0 = w * WIN;
    int k = 0;
    double s = 0.;
    if (!(flags & SLOT_NONACGT)) {
      // pure A/C/G/T: the counts come from the plane bits of the window (integer counts are the exact double
      // sums), the terms p*log(p)/ln2 from a table the host built with its libm: [k][count], k = window length
      k = min(WIN, len - i0);
      uint32_t lo, hi;
      int cntT, cntC, cntG;
      if (WIN <= 32) {
        const uint32_t wm = k >= 32 ? 0xffffffffu : ((1u << k) - 1u);
        lo = (s_pl[i0 >> 5] >> (i0 & 31)) & wm;
        hi = (s_pl[NW + (i0 >> 5)] >> (i0 & 31)) & wm;
        cntT = __popc(lo & hi);
        cntC = __popc(lo & ~hi);
        cntG = __popc(hi & ~lo);
      } else {  // windows of 64 bases = two plane words (bits past the end of the chunk are zero)
        cntT = cntC = cntG = 0;
#pragma unroll
        for (int h = 0; h < WIN / 32; h++) {
          const bool in = (i0 >> 5) + h < nwords;  // plane words are only written up to the chunk's last word
          lo = in ? s_pl[(i0 >> 5) + h] : 0u;
          hi = in ? s_pl[NW + (i0 >> 5) + h] : 0u;
          cntT += __popc(lo & hi);
          cntC += __popc(lo & ~hi);
          cntG += __popc(hi & ~lo);
        }
      }
      const int cnt[4] = {k - cntT - cntC - cntG, cntC, cntG, cntT};
      const double *et = ws.ent_table + (size_t)k * (WIN + 1);
#pragma unroll
      for (int c = 0; c < 4; c++) {
        s4[c] = (double)cnt[c];
        const double e = __ldg(et + cnt[c]);
        s = (c == 0) ? e : __dadd_rn(s, e);
      }
#pragma unroll
      for (int c = 0; c < 4; c++) tot[c] += s4[c];
    } else {
      for (int j = i0; j < i0 + WIN && j < len; j++, k++) {
        const uint32_t fc = s_fcode[sb[j]];
#pragma unroll
        for (int c = 0; c < 4; c++) s4[c] = __dadd_rn(s4[c], frac_of(fc, c));
      }
#pragma unroll
      for (int c = 0; c < 4; c++) tot[c] += s4[c];
#pragma unroll
      for (int c = 0; c < 4; c++) {
        const double p = __ddiv_rn(s4[c], (double)k);
        const double e = (p < 0.001) ? 0.0 : __ddiv_rn(__dmul_rn(p, log(p)), 0.69314718056);
        s = (c == 0) ? e : __dadd_rn(s, e);
      }
    }
    float v = __double2float_rn(-s);
    if (v < 0.f) v = 0.f;
    went[w] = v;
  }
  // Means: all fractions except 1/3 are dyadic, so partial sums are exact in any order; with
  // B/V/H/D present fall back to the reference's sequential summation order.
#pragma unroll
  for (int c = 0; c < 4; c++) {
    const double r = warp_sum(tot[c]);
    if (lane == 0) s_red[c][warp] = r;
  }
  __syncthreads();
  if (tid < 4) {
    double sum = 0.;
    if (flags & 2) {
      for (int i = 0; i < len; i++) sum = __dadd_rn(sum, frac_of(s_fcode[sb[i]], tid));
    } else {
      for (int w = 0; w < NWARP; w++) sum += s_red[tid][w];
    }
    s_off[tid] = __ddiv_rn(sum, (double)len);
  }
  __syncthreads();
  return flags;
}

// K0: preparation of the common case -- forward strand, pure A/C/G/T, at most H bases -- one WARP per signal, no
// block barriers.  These steps are a chain of dependent latencies (bases from HBM, plane words, window counts, table
// look-ups, reductions); inside the transform kernel they held a whole CTA and its 75 KB of shared memory idle for a
// third of its time.  Here thousands of warps are in flight and hide them; the transform kernel only picks up the
// planes, 256 window weights and four means.  Same arithmetic as encode_prepare, statement for statement.
template <int LOG2N>
__global__ void __launch_bounds__(256) encode_prep_kernel(const SigDesc *__restrict__ sigs, int nsig, Slots ws, PrepBuf prep) {
  constexpr int N = 1 << LOG2N, H = N / 2, NW = N / 32, WIN = N / 512, WMAX = H / WIN;
  static_assert(WIN <= 64, "at most two plane words per entropy window");
  __shared__ uint32_t s_plw[8][2 * (NW / 2)];  // plane words of at most H bases, per warp
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sig = blockIdx.x * 8 + warp;
  if (sig >= nsig) return;
  const SigDesc sd = sigs[sig];
  const int len = sd.len;
  if (len > H || len <= 0) {
    if (lane == 0) prep.flag[sig] = 0;
    return;
  }
  // an explicit reverse-strand signal (a query whose reverse strand cannot be derived from its forward spectrum): its
  // planes are the bit-reversed inverted forward planes, written to its slot first and read back for the window counts
  const bool rc_sig = sd.strand != 0;
  constexpr int NWH = NW / 2;
  uint32_t *s_pl = s_plw[warp];
  const uint8_t *__restrict__ src = sd.src;
  const int len32 = (len + 31) & ~31, nwords = len32 >> 5;
  const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 3u);
  const uint32_t *s32 = reinterpret_cast<const uint32_t *>(src - mis);  // chunk stores are padded by 32 bytes
  const bool al16 = (reinterpret_cast<uintptr_t>(src) & 15u) == 0;
  uint32_t bad = 0;
  for (int t = lane; t < (len32 >> 4); t += 32) {
    uint32_t w4[4] = {0u, 0u, 0u, 0u};
    if (t * 16 < len) {
      if (al16) {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(src) + t);
        w4[0] = v.x; w4[1] = v.y; w4[2] = v.z; w4[3] = v.w;
      } else {
        uint32_t x[5];
#pragma unroll
        for (int i = 0; i < 5; i++) x[i] = __ldg(s32 + t * 4 + i);
#pragma unroll
        for (int i = 0; i < 4; i++) w4[i] = __funnelshift_r(x[i], x[i + 1], 8 * mis);
      }
    }
    uint32_t lo16 = 0, hi16 = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int nv = min(max(len - (t * 16 + i * 4), 0), 4);  // bases of this word inside the chunk
      const uint32_t bm = nv == 0 ? 0u : (0xffffffffu >> (8 * (4 - nv)));
      const uint32_t w = w4[i] & bm;
      const uint32_t c = ((w >> 1) & 0x03030303u) ^ ((w >> 2) & 0x01010101u);
      const uint32_t c0 = c & 0x01010101u & bm, c1 = (c >> 1) & 0x01010101u & bm, c01 = c0 & c1;
      const uint32_t letters = (0x41414141u & bm) + 2u * c0 + 6u * c1 + 11u * c01;  // A, C = A+2, G = A+6, T = A+19
      bad |= letters ^ w;
      lo16 |= ((c0 * 0x01020408u) >> 24) << (4 * i);
      hi16 |= ((c1 * 0x01020408u) >> 24) << (4 * i);
    }
    reinterpret_cast<uint16_t *>(s_pl)[t] = (uint16_t)lo16;
    reinterpret_cast<uint16_t *>(s_pl + NWH)[t] = (uint16_t)hi16;
  }
  if (__any_sync(0xffffffffu, bad != 0u)) {  // a letter other than A/C/G/T: the general path of the transform kernel
    if (lane == 0) prep.flag[sig] = 0;
    return;
  }
  __syncwarp();
  uint32_t *own_planes = ws.planes + (size_t)sd.slot * 2 * NW;
  if (!rc_sig) {
    for (int w = lane; w < NW; w += 32) {
      own_planes[w] = w < nwords ? s_pl[w] : 0u;
      own_planes[NW + w] = w < nwords ? s_pl[NWH + w] : 0u;
    }
  }
  if (sd.rc_slot1 || rc_sig) {  // planes (and meta) of the other orientation (see encode_prepare)
    const int oslot = rc_sig ? sd.slot : sd.rc_slot1 - 1;
    uint32_t *planes = ws.planes + (size_t)oslot * 2 * NW;
    const int q = (len - 1) >> 5, r = (len - 1) & 31;
    for (int j = lane; j < NW; j += 32) {
      uint32_t lo = 0u, hi = 0u;
      if (j <= q) {
        const uint32_t l0 = __brev(~s_pl[q - j]), h0 = __brev(~s_pl[NWH + q - j]);
        const uint32_t l1 = j < q ? __brev(~s_pl[q - j - 1]) : 0u, h1 = j < q ? __brev(~s_pl[NWH + q - j - 1]) : 0u;
        lo = __funnelshift_r(l0, l1, 31 - r);
        hi = __funnelshift_r(h0, h1, 31 - r);
        if (j == q) {
          const uint32_t vm = r == 31 ? 0xffffffffu : ((2u << r) - 1u);
          lo &= vm;
          hi &= vm;
        }
      }
      planes[j] = lo;
      planes[NW + j] = hi;
    }
    if (lane == 0 && !rc_sig) {
      SlotMeta m;
      m.len = len;
      m.flags = 0;
      m.q_re = m.q_im = m.q_nyq = 0.f;
      m.zmode = ZM_FOUR;
      m.zslot = 0;
      m.pad = 0;
      ws.meta[oslot] = m;
    }
    __syncwarp();  // rc_sig: the window counts below read these words back
  }
  // plane word w of this signal's own orientation (zero past the chunk's last word)
  auto plane_word = [&](int p, int w) -> uint32_t {
    if (w >= nwords) return 0u;
    return rc_sig ? own_planes[p * NW + w] : s_pl[p * NWH + w];
  };
  // entropy weights per window and base totals (integer counts are the exact double sums of the reference)
  int totC = 0, totG = 0, totT = 0;
  const int nwin = (len + WIN - 1) / WIN;
  float *wout = prep.went + (size_t)sig * WMAX;
  for (int w = lane; w < WMAX; w += 32) {
    float v = 0.f;
    if (w < nwin) {
      const int i0 = w * WIN, k = min(WIN, len - i0);
      int cntT = 0, cntC = 0, cntG = 0;
      if (WIN <= 32) {
        const uint32_t wm = k >= 32 ? 0xffffffffu : ((1u << k) - 1u);
        const uint32_t lo = (plane_word(0, i0 >> 5) >> (i0 & 31)) & wm, hi = (plane_word(1, i0 >> 5) >> (i0 & 31)) & wm;
        cntT = __popc(lo & hi);
        cntC = __popc(lo & ~hi);
        cntG = __popc(hi & ~lo);
      } else {  // windows of 64 bases = two plane words (bits past the end of the chunk are zero)
#pragma unroll
        for (int h = 0; h < WIN / 32; h++) {
          const uint32_t lo = plane_word(0, (i0 >> 5) + h), hi = plane_word(1, (i0 >> 5) + h);
          cntT += __popc(lo & hi);
          cntC += __popc(lo & ~hi);
          cntG += __popc(hi & ~lo);
        }
      }
      const int cnt[4] = {k - cntT - cntC - cntG, cntC, cntG, cntT};
      totC += cntC;
      totG += cntG;
      totT += cntT;
      const double *et = ws.ent_table + (size_t)k * (WIN + 1);
      double s = 0.;
#pragma unroll
      for (int c = 0; c < 4; c++) {
        const double e = __ldg(et + cnt[c]);
        s = (c == 0) ? e : __dadd_rn(s, e);
      }
      v = __double2float_rn(-s);
      if (v < 0.f) v = 0.f;
    }
    wout[w] = v;
  }
  totC = __reduce_add_sync(0xffffffffu, totC);
  totG = __reduce_add_sync(0xffffffffu, totG);
  totT = __reduce_add_sync(0xffffffffu, totT);
  if (lane < 4) {
    const int cnt = lane == 0 ? len - totC - totG - totT : lane == 1 ? totC : lane == 2 ? totG : totT;
    prep.off[(size_t)sig * 4 + lane] = __ddiv_rn((double)cnt, (double)len);
  }
  if (lane == 0) prep.flag[sig] = 1;
}

template <int LOG2N, int NT>
__global__ void __launch_bounds__(NT, (LOG2N <= 13 ? 3 : 1))
    encode_fft_kernel(const SigDesc *__restrict__ sigs, Slots ws, float *__restrict__ tap, PrepBuf prep,
                      const uint32_t *__restrict__ jobs, const unsigned int *__restrict__ njobs) {
  constexpr int N = 1 << LOG2N, H = N / 2, WIN = N / 512, NWARP = NT / 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2 *buf = reinterpret_cast<float2 *>(smem_raw);            // N complex (swizzled slots)
  uint8_t *sb = smem_raw + (size_t)N * sizeof(float2);            // N oriented bases
  float *went = reinterpret_cast<float *>(sb + N);                // 512 window weights
  uint16_t *s_fcode = reinterpret_cast<uint16_t *>(went + 512);   // 128 x u16
  uint8_t *s_comp = reinterpret_cast<uint8_t *>(s_fcode + 128);   // 128
  uint8_t *s_base2 = s_comp + 128;                                // 128
  __shared__ double s_red[4][NWARP];
  __shared__ double s_off[4];
  __shared__ double s_poff;  // three-channel form: the partner's G mean
  __shared__ int s_flags;

  const int tid = threadIdx.x;
  // which signals: jobs == nullptr: signal blockIdx.x; jobs given: signal jobs[blockIdx.x]; jobs and njobs given: a fixed
  // grid strides over the *njobs entries of a list written on the device (the fused kernel's fall-backs, normally none)
  for (int jb = blockIdx.x; njobs != nullptr ? jb < (int)*njobs : jb == (int)blockIdx.x; jb += gridDim.x) {
  const int sig = jobs != nullptr ? (int)jobs[jb] : jb;
  const SigDesc sd = sigs[sig];
  const int len = sd.len;
  const float2 *__restrict__ wn = ws.wn;

  // prepared by encode_prep_kernel (the common case): pick up planes, window weights and means; otherwise do it here
  const bool prepared = prep.flag != nullptr && tap == nullptr && prep.flag[sig] != 0;
  uint32_t *s_pl2 = reinterpret_cast<uint32_t *>(sb);  // prepared: plane words [2][NW] in place of the bases
  int flags = 0;
  // Three-channel form (sx_kernels.h): a prepared signal is pure A/C/G/T, its four channels sum to zero, so the T
  // spectrum is never transformed; the second transform carries G of this signal and G of its partner.
  int zmode = ZM_FOUR, partner = -1, plen = 0;
  if (prepared) {
    if (sd.g_mode == G_OWNER) {
      zmode = ZM_RE;
      if (sd.g_partner >= 0 && prep.flag[sd.g_partner] != 0) partner = sd.g_partner;
    } else if (sd.g_mode == G_MEMBER && sd.g_partner >= 0) {
      zmode = prep.flag[sd.g_partner] != 0 ? ZM_IM : ZM_RE;  // owner not pure: this signal transforms (G + i 0) itself
    }
    constexpr int NWp = N / 32;
    const uint32_t *pl = ws.planes + (size_t)sd.slot * 2 * NWp;
    for (int w = tid; w < 2 * NWp; w += NT) s_pl2[w] = pl[w];
    const float *wsrc = prep.went + (size_t)sig * (H / WIN);
    for (int w = tid; w < H / WIN; w += NT) went[w] = wsrc[w];
    if (tid < 4) s_off[tid] = prep.off[(size_t)sig * 4 + tid];
    if (partner >= 0) {  // the partner's planes, weights and G mean behind this signal's
      plen = sd.g_plen;
      const uint32_t *ppl = ws.planes + (size_t)sd.g_pslot * 2 * NWp;
      for (int w = tid; w < 2 * NWp; w += NT) s_pl2[2 * NWp + w] = ppl[w];
      const float *pw = prep.went + (size_t)partner * (H / WIN);
      for (int w = tid; w < H / WIN; w += NT) went[H / WIN + w] = pw[w];
      if (tid == 4) s_poff = prep.off[(size_t)partner * 4 + 2];
    }
    __syncthreads();
  } else {
    flags = encode_prepare<LOG2N, NT>(sd, ws, smem_raw, sb, went, s_fcode, s_comp, s_base2, s_red, s_off, &s_flags, true);
  }
  const bool flat = len < 1024;  // "Skip entropy": weight 1 everywhere (CrossCorr.cc:39-44)

  if (tap != nullptr) {
    float *te = tap + (size_t)sig * 5 * N;
    for (int k = tid; k < N; k += NT) te[k] = flat ? 1.f : (k < len ? went[k / WIN] : 0.f);
  }

  // ---- 3. two complex transforms: (A + iC) then (G + iT) -----------------------------------------
  // The N-point transform is split once (sx_fft.cuh): half 0 of the buffer gets e[n] = z[n] + z[n+H],
  // half 1 gets o[n] = (z[n] - z[n+H]) w_N^n; a chunk of at most H bases has z[n+H] = 0.
  float acc_re = 0.f, acc_im = 0.f, acc_ny = 0.f;
  constexpr int PH1 = bin_slot<LOG2N>(H - 1), PH = bin_slot<LOG2N>(H), PH2 = bin_slot<LOG2N>(H + 1);
  const bool pure = !(flags & SLOT_NONACGT);
  const bool fastgen = pure && len <= H;
  const int rounds = zmode == ZM_IM ? 1 : 2;  // a member's G rides its owner's second transform
#pragma unroll 1
  for (int pr = 0; pr < rounds; pr++) {
    // the channel in the real part (c0) and in the imaginary part (c1) of this round's packed signal; in the
    // three-channel form the second round is (G of this signal) + i (G of the partner, or nothing)
    const bool g_round = pr == 1 && zmode != ZM_FOUR;
    const uint32_t c0 = 2 * pr, c1 = g_round ? 2u : 2 * pr + 1;
    const double off0 = s_off[c0], off1 = g_round ? (partner >= 0 ? s_poff : 0.0) : s_off[c1];
    const int len1 = g_round ? plen : len;  // plen = 0 without a partner: imaginary part all zero
    // sample = (float)(weight * (fraction - mean)); for A/C/G/T the fraction is 1 or 0
    const double hit0 = __dsub_rn(1.0, off0), miss0 = __dsub_rn(0.0, off0);
    const double hit1 = __dsub_rn(1.0, off1), miss1 = __dsub_rn(0.0, off1);
    if (fastgen) {
      // Fast path: one thread per entropy window (N/512 bases).  Inside a window the sample is one of
      // two floats per channel -- (float)(w*(1-mean)) or (float)(w*(0-mean)) -- rounded exactly as the
      // per-base double product of the reference; the base only selects between them.
      const bool flat1 = g_round ? plen < 1024 : flat;
      const int src1 = (g_round && partner >= 0) ? 1 : 0;  // whose planes / weights feed the imaginary part
      for (int w = tid; w < H / WIN; w += NT) {
        const int k0 = w * WIN;
        const double e0 = flat ? 1.0 : (double)went[w];
        const double e1 = flat1 ? 1.0 : (double)went[src1 * (H / WIN) + w];
        const float h0 = __double2float_rn(__dmul_rn(e0, hit0)), m0 = __double2float_rn(__dmul_rn(e0, miss0));
        const float h1 = __double2float_rn(__dmul_rn(e1, hit1)), m1 = __double2float_rn(__dmul_rn(e1, miss1));
        const float2 wb = __ldg(wn + k0);  // w_N^{k0}; w_N^{k0 + j} = wb * (compile-time) w_N^j
        // bit j of sel0 / sel1: base k0 + j is the channel's letter (codes A,C,G,T -> 0,1,2,3 as two bit planes)
        uint32_t sel0, sel1;
        if (prepared) {
          constexpr int NWp = N / 32;
          const uint32_t lo = s_pl2[k0 >> 5] >> (k0 & 31), hi = s_pl2[NWp + (k0 >> 5)] >> (k0 & 31);
          const uint32_t *p1 = s_pl2 + src1 * 2 * NWp;
          const uint32_t lo1 = p1[k0 >> 5] >> (k0 & 31), hi1 = p1[NWp + (k0 >> 5)] >> (k0 & 31);
          sel0 = ((c0 & 1u) ? lo : ~lo) & ((c0 & 2u) ? hi : ~hi);
          sel1 = ((c1 & 1u) ? lo1 : ~lo1) & ((c1 & 2u) ? hi1 : ~hi1);
        } else {
          sel0 = sel1 = 0u;
#pragma unroll
          for (int j = 0; j < (WIN + 3) / 4; j++) {
            const uint32_t p4 = reinterpret_cast<const uint32_t *>(sb + k0)[j];
            const uint32_t c4 = ((p4 >> 1) & 0x03030303u) ^ ((p4 >> 2) & 0x01010101u);
#pragma unroll
            for (int b = 0; b < 4; b++) {
              const uint32_t code = (c4 >> (8 * b)) & 3u;
              sel0 |= (uint32_t)(code == c0) << (4 * j + b);
              sel1 |= (uint32_t)(code == c1) << (4 * j + b);
            }
          }
        }
        // bases past the end of a chunk give zero samples: one mask per window instead of a test per sample
        const uint32_t in0 = k0 + WIN <= len ? 0xffffffffu : (k0 < len ? (1u << (len - k0)) - 1u : 0u);
        const uint32_t in1 = k0 + WIN <= len1 ? 0xffffffffu : (k0 < len1 ? (1u << (len1 - k0)) - 1u : 0u);
        const bool full = (in0 & in1) == 0xffffffffu;
#pragma unroll
        for (int j = 0; j < WIN; j++) {
          const int k = k0 + j;
          float2 v;
          v.x = ((sel0 >> j) & 1u) ? h0 : m0;
          v.y = ((sel1 >> j) & 1u) ? h1 : m1;
          if (!full) {
            if (!((in0 >> j) & 1u)) v.x = 0.f;
            if (!((in1 >> j) & 1u)) v.y = 0.f;
          }
          constexpr double ang = -2.0 * 3.14159265358979323846 / (double)N;
          const float2 st = make_float2((float)cx_cos_small(ang * j), (float)cx_sin_small(ang * j));
          const float2 wk = j == 0 ? wb : cmul(wb, st);
          buf[swz(k)] = v;
          buf[H + swz(k)] = cmul(v, wk);
        }
      }
      if (tap != nullptr) {  // taps read back exactly what the transform is about to see
        __syncthreads();
        float *ts = tap + (size_t)sig * 5 * N + (size_t)(1 + 2 * pr) * N;
        for (int k = tid; k < N; k += NT) {
          const float2 v = k < H ? buf[swz(k)] : make_float2(0.f, 0.f);
          ts[k] = v.x;
          ts[N + k] = v.y;
        }
      }
    } else {
      // general path (IUPAC letters, or a query chunk longer than H): the whole signal in natural
      // order first, then the radix-2 split in place
      for (int k = tid; k < N; k += NT) {
        float2 v = make_float2(0.f, 0.f);
        if (k < len) {
          const double e = flat ? 1.0 : (double)went[k / WIN];
          if (pure) {
            const uint32_t code = s_base2[sb[k]];
            v.x = __double2float_rn(__dmul_rn(e, code == (uint32_t)(2 * pr) ? hit0 : miss0));
            v.y = __double2float_rn(__dmul_rn(e, code == (uint32_t)(2 * pr + 1) ? hit1 : miss1));
          } else {
            const uint32_t fc = s_fcode[sb[k]];
            v.x = __double2float_rn(__dmul_rn(e, __dsub_rn(frac_of(fc, 2 * pr), off0)));
            v.y = __double2float_rn(__dmul_rn(e, __dsub_rn(frac_of(fc, 2 * pr + 1), off1)));
          }
        }
        buf[swz(k)] = v;
        if (tap != nullptr) {
          float *ts = tap + (size_t)sig * 5 * N + (size_t)(1 + 2 * pr) * N;
          ts[k] = v.x;
          ts[N + k] = v.y;
        }
      }
      __syncthreads();
      for (int n = tid; n < H; n += NT) {
        const float2 a = buf[swz(n)], b = buf[H + swz(n)];
        buf[swz(n)] = cadd(a, b);
        buf[H + swz(n)] = cmul(csub(a, b), __ldg(wn + n));
      }
    }
    __syncthreads();
    fft_forward_halves<LOG2N, N, NT>(buf, tid);
    // spectra out in shared-memory slot order ([even | odd] halves, scrambled bins, swizzled slots): one bulk
    // asynchronous copy (TMA, shared -> global) issued by one thread instead of 16 load/store pairs per thread;
    // the threads go on while the copy engine drains the buffer, and only wait for it to have READ the buffer
    // before the next signal is generated into it
    bulk_store_fence();
    __syncthreads();
    if (tid == 0) {
      float2 *dst = ws.spec + ((size_t)sd.slot * 2 + pr) * N;
      constexpr int PIECE = N < 4096 ? N : 4096;  // float2 per copy (at most 32 KiB)
#pragma unroll
      for (int o = 0; o < N; o += PIECE) bulk_store(dst + o, buf + o, PIECE * (int)sizeof(float2));
      bulk_store_commit();
    }
    if (tid == 0) {
      // per-channel bins from the packed transform: X_re-channel[k] + X_im-channel[k]
      //   = (a+b)/2 - i (a-b)/2  with a = Z[k], b = conj(Z[N-k])
      const float2 a = buf[PH1], z2 = buf[PH2], zn = buf[PH];
      const float2 b = make_float2(z2.x, -z2.y);
      const float2 s = cadd(a, b), d = csub(a, b);
      acc_re += 0.5f * s.x + 0.5f * d.y;
      acc_im += 0.5f * s.y - 0.5f * d.x;
      acc_ny += zn.x + zn.y;
      bulk_store_wait_read();  // the buffer may be overwritten (and, after the last round, the CTA may leave)
    }
    __syncthreads();
  }
  if (tid == 0) {
    SlotMeta m;
    m.len = len;
    m.flags = flags & SLOT_NONACGT;
    // three-channel form: the channel sum of the spectrum is zero by construction (the reference's is float noise)
    m.q_re = zmode == ZM_FOUR ? acc_re : 0.f;
    m.q_im = zmode == ZM_FOUR ? acc_im : 0.f;
    m.q_nyq = zmode == ZM_FOUR ? acc_ny : 0.f;
    m.zmode = zmode;
    m.zslot = zmode == ZM_IM ? sd.g_pslot : sd.slot;
    m.pad = 0;
    ws.meta[sd.slot] = m;
  }
  __syncthreads();  // list mode: the shared-memory state of this signal is dead
  }
}

// ---- the reference's twiddle drift at N = 32768, applied to one half (bins of one parity) of a spectrum held in
// shared memory in scrambled order (sx_kernels.h, fill_drift_table).  Bins k, k + N/4, k + N/2, k + 3N/4 (k < N/4)
// form a quad: the reference's last pass ties k to k + N/2, the pass below it ties k to k + N/4.  In this half's
// H-point transform they are m, m + H/4, m + H/2, m + 3H/4 with m = k >> 1: the top digit of the scrambled order
// moves by 2, so a quad is positions r, r+2, r+4, r+6 of a block of 8.
//   FORWARD: exact spectrum -> what the reference's two oscillator passes produce
//   INVERSE: product -> the spectrum whose EXACT inverse equals the reference's inverse with its oscillator passes
template <int LOG2N, int NT, bool INVERSE>
__device__ __forceinline__ void drift_correct_half(float2 *buf, const float2 *__restrict__ drift, int half, int tid) {
  constexpr int N = 1 << LOG2N, H = N / 2, Q = N / 4;
  static_assert(LOG2N == 15, "the quad layout below is that of Plan<15> (last radix 8)");
  for (int q = tid; q < H / 4; q += NT) {
    const int r = ((q >> 1) << 3) | (q & 1);  // block of 8 positions, quad 0 (even positions) or 1 (odd)
    const int k = 2 * natural_bin<LOG2N>(r) + half;  // < N/4
    float2 rho0 = __ldg(drift + k), rho1 = __ldg(drift + Q + k), phi = __ldg(drift + 2 * Q + k);
    if (INVERSE) {
      rho0.y = -rho0.y;
      rho1.y = -rho1.y;
      phi.y = -phi.y;
    }
    const int pa = swz(r), pb = swz(r + 2), pc = swz(r + 4), pd = swz(r + 6);
    const float2 xa = buf[pa], xb = buf[pb], xc = buf[pc], xd = buf[pd];
    float2 s0 = cadd(xa, xc), d0 = csub(xa, xc), s1 = cadd(xb, xd), d1 = csub(xb, xd);
    if (!INVERSE) {
      s0 = make_float2(0.5f * s0.x, 0.5f * s0.y);
      d0 = make_float2(0.5f * d0.x, 0.5f * d0.y);
      s1 = make_float2(0.5f * s1.x, 0.5f * s1.y);
      d1 = make_float2(0.5f * d1.x, 0.5f * d1.y);
      const float2 w = cmul(phi, make_float2(d0.x + d1.y, d0.y - d1.x));  // phi (D0 - i D1)
      const float2 t0 = cmul(rho0, cadd(d0, w));
      const float2 t1 = cmul(rho1, make_float2(d1.x - w.y, d1.y + w.x));  // rho1 (D1 + i W)
      const float2 de = cmul(phi, csub(s0, s1));
      const float2 e0 = cadd(s0, de), e1 = csub(s1, de);
      buf[pa] = cadd(e0, t0);
      buf[pc] = csub(e0, t0);
      buf[pb] = cadd(e1, t1);
      buf[pd] = csub(e1, t1);
    } else {
      const float2 a0 = cmul(rho0, d0), a1 = cmul(rho1, d1);
      const float2 v = cmul(phi, make_float2(a0.x + a1.y, a0.y - a1.x));  // chi (A0 - i A1)
      const float2 u0 = cadd(a0, v), u1 = make_float2(a1.x - v.y, a1.y + v.x);  // A1 + i V
      const float2 ds = cmul(phi, csub(s0, s1));
      const float2 e0 = cadd(s0, ds), e1 = csub(s1, ds);
      buf[pa] = make_float2(0.5f * (e0.x + u0.x), 0.5f * (e0.y + u0.y));
      buf[pc] = make_float2(0.5f * (e0.x - u0.x), 0.5f * (e0.y - u0.y));
      buf[pb] = make_float2(0.5f * (e1.x + u1.x), 0.5f * (e1.y + u1.y));
      buf[pd] = make_float2(0.5f * (e1.x - u1.x), 0.5f * (e1.y - u1.y));
    }
  }
  __syncthreads();
}

// K1 for transforms that do not fit one CTA's shared memory (N = 32768: 256 KiB of complex points): the
// two H-point transforms of the split (sx_fft.cuh) are independent, so a signal is handled by TWO CTAs,
// blockIdx.y = half: 0 builds e[n] = z[n] + z[n+H] (even bins), 1 builds o[n] = (z[n] - z[n+H]) w_N^n (odd
// bins).  Both repeat the cheap encoding steps; only half 0 writes planes / bytes / taps.  The quirk bins
// H-1 and H+1 are odd (half 1), bin H is even (half 0): the meta fields are written field by field.
template <int LOG2N, int NT>
__global__ void __launch_bounds__(NT, 1)
    encode_fft_half_kernel(const SigDesc *__restrict__ sigs, Slots ws, float *__restrict__ tap, PrepBuf prep) {
  constexpr int N = 1 << LOG2N, H = N / 2, WIN = N / 512, NWARP = NT / 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2 *buf = reinterpret_cast<float2 *>(smem_raw);            // H complex (swizzled slots)
  uint8_t *sb = smem_raw + (size_t)H * sizeof(float2);            // N oriented bases
  float *went = reinterpret_cast<float *>(sb + N);                // 512 window weights
  uint16_t *s_fcode = reinterpret_cast<uint16_t *>(went + 512);   // 128 x u16
  uint8_t *s_comp = reinterpret_cast<uint8_t *>(s_fcode + 128);   // 128
  uint8_t *s_base2 = s_comp + 128;                                // 128
  __shared__ double s_red[4][NWARP];
  __shared__ double s_off[4];
  __shared__ double s_poff;  // three-channel form: the partner's G mean
  __shared__ int s_flags;

  const int tid = threadIdx.x;
  const int half = blockIdx.y;
  const SigDesc sd = sigs[blockIdx.x];
  const int len = sd.len;
  const float2 *__restrict__ wn = ws.wn;

  // prepared by encode_prep_kernel (pure A/C/G/T, at most H bases; either strand): pick up the oriented planes, window
  // weights and means; otherwise both CTAs of the signal prepare it themselves
  const bool prepared = prep.flag != nullptr && tap == nullptr && prep.flag[blockIdx.x] != 0;
  uint32_t *s_pl2 = reinterpret_cast<uint32_t *>(sb);  // prepared: plane words [2][N/32] in place of the bases
  int flags = 0;
  // three-channel form, as in encode_fft_kernel: the second transform carries G of this signal and G of its partner
  int zmode = ZM_FOUR, partner = -1, plen = 0;
  if (prepared) {
    if (sd.g_mode == G_OWNER) {
      zmode = ZM_RE;
      if (sd.g_partner >= 0 && prep.flag[sd.g_partner] != 0) partner = sd.g_partner;
    } else if (sd.g_mode == G_MEMBER && sd.g_partner >= 0) {
      zmode = prep.flag[sd.g_partner] != 0 ? ZM_IM : ZM_RE;
    }
    constexpr int NWp = N / 32;
    const uint32_t *pl = ws.planes + (size_t)sd.slot * 2 * NWp;
    for (int w = tid; w < 2 * NWp; w += NT) s_pl2[w] = pl[w];
    const float *wsrc = prep.went + (size_t)blockIdx.x * (H / WIN);
    for (int w = tid; w < H / WIN; w += NT) went[w] = wsrc[w];
    if (tid < 4) s_off[tid] = prep.off[(size_t)blockIdx.x * 4 + tid];
    if (partner >= 0) {
      plen = sd.g_plen;
      const uint32_t *ppl = ws.planes + (size_t)sd.g_pslot * 2 * NWp;
      for (int w = tid; w < 2 * NWp; w += NT) s_pl2[2 * NWp + w] = ppl[w];
      const float *pw = prep.went + (size_t)partner * (H / WIN);
      for (int w = tid; w < H / WIN; w += NT) went[H / WIN + w] = pw[w];
      if (tid == 4) s_poff = prep.off[(size_t)partner * 4 + 2];
    }
    __syncthreads();
  } else {
    flags = encode_prepare<LOG2N, NT>(sd, ws, smem_raw, sb, went, s_fcode, s_comp, s_base2, s_red, s_off, &s_flags, half == 0);
  }
  const bool flat = len < 1024;  // "Skip entropy": weight 1 everywhere (CrossCorr.cc:39-44)
  const bool pure = !(flags & SLOT_NONACGT);
  if (tap != nullptr && half == 0) {
    float *te = tap + (size_t)blockIdx.x * 5 * N;
    for (int k = tid; k < N; k += NT) te[k] = flat ? 1.f : (k < len ? went[k / WIN] : 0.f);
  }

  float acc_re = 0.f, acc_im = 0.f, acc_ny = 0.f;
  // slots inside this CTA's half
  constexpr int PH1 = bin_slot<LOG2N>(H - 1) - H, PH2 = bin_slot<LOG2N>(H + 1) - H, PH = bin_slot<LOG2N>(H);
  const int rounds = zmode == ZM_IM ? 1 : 2;
#pragma unroll 1
  for (int pr = 0; pr < rounds; pr++) {
    const bool g_round = pr == 1 && zmode != ZM_FOUR;  // (G of this signal) + i (G of the partner, or nothing)
    const double off0 = s_off[2 * pr], off1 = g_round ? (partner >= 0 ? s_poff : 0.0) : s_off[2 * pr + 1];
    // sample k of the packed signal: (float)(weight * (fraction - mean)) per channel (SeqToPCM), 0 past the end
    auto sample = [&](int k) -> float2 {
      float2 v = make_float2(0.f, 0.f);
      if (k < len) {
        const double e = flat ? 1.0 : (double)went[k / WIN];
        double f0, f1;
        if (pure) {
          const uint32_t code = s_base2[sb[k]];
          f0 = code == (uint32_t)(2 * pr) ? 1.0 : 0.0;
          f1 = code == (uint32_t)(2 * pr + 1) ? 1.0 : 0.0;
        } else {
          const uint32_t fc = s_fcode[sb[k]];
          f0 = frac_of(fc, 2 * pr);
          f1 = frac_of(fc, 2 * pr + 1);
        }
        v.x = __double2float_rn(__dmul_rn(e, __dsub_rn(f0, off0)));
        v.y = __double2float_rn(__dmul_rn(e, __dsub_rn(f1, off1)));
      }
      return v;
    };
    if (tap != nullptr && half == 0) {
      float *ts = tap + (size_t)blockIdx.x * 5 * N + (size_t)(1 + 2 * pr) * N;
      for (int k = tid; k < N; k += NT) {
        const float2 v = sample(k);
        ts[k] = v.x;
        ts[N + k] = v.y;
      }
    }
    if (prepared) {
      // the chunk has at most H bases, so z[n + H] = 0: e[n] = z[n], o[n] = z[n] w_N^n.  One thread per 32 samples of
      // a window: inside a window the sample is one of two floats per channel, rounded exactly as the per-base double
      // product of the reference (see encode_fft_kernel); the base only selects between them.
      constexpr int NWp = N / 32;
      const uint32_t c0 = 2 * pr, c1 = g_round ? 2u : 2 * pr + 1;
      const int len1 = g_round ? plen : len;  // plen = 0 without a partner: imaginary part all zero
      const bool flat1 = g_round ? plen < 1024 : flat;
      const int src1 = (g_round && partner >= 0) ? 1 : 0;  // whose planes / weights feed the imaginary part
      const double hit0 = __dsub_rn(1.0, off0), miss0 = __dsub_rn(0.0, off0);
      const double hit1 = __dsub_rn(1.0, off1), miss1 = __dsub_rn(0.0, off1);
      constexpr int SEG = WIN < 32 ? WIN : 32;  // samples per thread step
      for (int sg = tid; sg < H / SEG; sg += NT) {
        const int k0 = sg * SEG, w = k0 / WIN;
        const double e = flat ? 1.0 : (double)went[w];
        const double e1 = flat1 ? 1.0 : (double)went[src1 * (H / WIN) + w];
        const float h0 = __double2float_rn(__dmul_rn(e, hit0)), m0 = __double2float_rn(__dmul_rn(e, miss0));
        const float h1 = __double2float_rn(__dmul_rn(e1, hit1)), m1 = __double2float_rn(__dmul_rn(e1, miss1));
        const uint32_t lo = s_pl2[k0 >> 5] >> (k0 & 31), hi = s_pl2[NWp + (k0 >> 5)] >> (k0 & 31);
        const uint32_t *p1 = s_pl2 + src1 * 2 * NWp;
        const uint32_t lo1 = p1[k0 >> 5] >> (k0 & 31), hi1 = p1[NWp + (k0 >> 5)] >> (k0 & 31);
        const uint32_t sel0 = ((c0 & 1u) ? lo : ~lo) & ((c0 & 2u) ? hi : ~hi);
        const uint32_t sel1 = ((c1 & 1u) ? lo1 : ~lo1) & ((c1 & 2u) ? hi1 : ~hi1);
        const uint32_t in = k0 + SEG <= len ? 0xffffffffu : (k0 < len ? (1u << (len - k0)) - 1u : 0u);
        const uint32_t in1 = k0 + SEG <= len1 ? 0xffffffffu : (k0 < len1 ? (1u << (len1 - k0)) - 1u : 0u);
        const float2 wb = half ? __ldg(wn + k0) : make_float2(1.f, 0.f);
#pragma unroll
        for (int j = 0; j < SEG; j++) {
          float2 v;
          v.x = ((sel0 >> j) & 1u) ? h0 : m0;
          v.y = ((sel1 >> j) & 1u) ? h1 : m1;
          if (!((in >> j) & 1u)) v.x = 0.f;
          if (!((in1 >> j) & 1u)) v.y = 0.f;
          if (half) {
            constexpr double ang = -2.0 * 3.14159265358979323846 / (double)N;
            const float2 st = make_float2((float)cx_cos_small(ang * j), (float)cx_sin_small(ang * j));
            v = cmul(v, j == 0 ? wb : cmul(wb, st));
          }
          buf[swz(k0 + j)] = v;
        }
      }
    } else {
      for (int n = tid; n < H; n += NT) {
        const float2 a = sample(n), b = sample(n + H);
        buf[swz(n)] = half == 0 ? cadd(a, b) : cmul(csub(a, b), __ldg(wn + n));
      }
    }
    __syncthreads();
    fft_forward_halves<LOG2N, H, NT>(buf, tid);
    if constexpr (LOG2N == 15) drift_correct_half<LOG2N, NT, false>(buf, ws.drift, half, tid);
    float4 *dst = reinterpret_cast<float4 *>(ws.spec + ((size_t)sd.slot * 2 + pr) * N + (size_t)half * H);
    const float4 *s4p = reinterpret_cast<const float4 *>(buf);
    for (int k = tid; k < H / 2; k += NT) dst[k] = s4p[k];
    if (tid == 0) {
      if (half == 1) {  // per-channel bins from the packed transform, see encode_fft_kernel
        const float2 a = buf[PH1], z2 = buf[PH2];
        const float2 b = make_float2(z2.x, -z2.y);
        const float2 s = cadd(a, b), d = csub(a, b);
        acc_re += 0.5f * s.x + 0.5f * d.y;
        acc_im += 0.5f * s.y - 0.5f * d.x;
      } else {
        const float2 zn = buf[PH];
        acc_ny += zn.x + zn.y;
      }
    }
    __syncthreads();
  }
  if (tid == 0) {
    SlotMeta *m = ws.meta + sd.slot;
    const bool four = zmode == ZM_FOUR;  // three-channel form: the channel sum of the spectrum is zero by construction
    if (half == 0) {
      m->len = len;
      m->flags = flags & SLOT_NONACGT;
      m->q_nyq = four ? acc_ny : 0.f;
      m->zmode = zmode;
      m->zslot = zmode == ZM_IM ? sd.g_pslot : sd.slot;
      m->pad = 0;
    } else {
      m->q_re = four ? acc_re : 0.f;
      m->q_im = four ? acc_im : 0.f;
    }
  }
}

// =================================================================================================
// K2: spectral product + inverse transform + FindTop.
//   xcorr_pair_kernel   one CTA per CHUNK PAIR: both strands from the forward query spectrum, one
//                       inverse transform (forward strand in the real part, reverse in the imaginary)
//   xcorr_findtop_kernel one CTA per strand-pair (reverse-strand signal with its own spectrum)
// =================================================================================================
// FindTop (SeqAnalyzer::FindTop, CrossCorr.cc:878-944) over xc[i] = COMP(buf[(i + off) mod N]) * scale:
// RMS envelope per 256 lags (float square, double accumulate), threshold env*cutoff + 1, ballot
// mask, block scan, ordered compaction into the candidate pool (one atomic per strand-pair).
template <int LOG2N, int NT, class XcAt>
__device__ __forceinline__ void findtop_impl(XcAt xc_at, double co, uint32_t *mask,
                                             unsigned int *s_wtot, unsigned int *s_base, int spi,
                                             uint16_t *__restrict__ cand_pool, unsigned int pool_cap,
                                             uint2 *__restrict__ cand_ref, BatchCounters *ctr,
                                             float *__restrict__ xc_tap) {
  constexpr int N = 1 << LOG2N, NW = N / 32, NB = N / 256, NWARP = NT / 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (xc_tap != nullptr) {
    float *o = xc_tap + (size_t)spi * N;
    for (int i = tid; i < N; i += NT) o[i] = xc_at(i);
  }
  // one warp per block of 256 lags: its 8 values per lane stay in registers between the envelope and the
  // threshold test (every lane gets the block sum from the butterfly reduction)
  for (int b = warp; b < NB; b += NWARP) {
    float v[8];
    double acc = 0.;
#pragma unroll
    for (int r = 0; r < 8; r++) {
      v[r] = xc_at(b * 256 + r * 32 + lane);
      if (NB > 8) acc += (double)__fmul_rn(v[r], v[r]);
    }
    if (NB > 8) acc = warp_sum(acc);
    const double thr = __dadd_rn(__dmul_rn(__dsqrt_rn(__ddiv_rn(acc, 256.0)), co), 1.0);
    // (double)v > thr for a float v is v > (largest float <= thr): exact, and the compares stay in FP32
    const float thr_f = __double2float_rd(thr);
#pragma unroll
    for (int r = 0; r < 8; r++) {
      const uint32_t m = __ballot_sync(0xffffffffu, v[r] > thr_f);
      if (lane == 0) mask[b * 8 + r] = m;
    }
  }
  __syncthreads();
  // ---- ordered compaction: exclusive scan of per-word popcounts -----------------------------------
  constexpr int IPT = (NW + NT - 1) / NT;  // words per thread (contiguous)
  unsigned int cnt[IPT], mine = 0;
#pragma unroll
  for (int r = 0; r < IPT; r++) {
    const int w = tid * IPT + r;
    cnt[r] = (w < NW) ? __popc(mask[w]) : 0;
    mine += cnt[r];
  }
  unsigned int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) s_wtot[warp] = incl;
  __syncthreads();
  unsigned int wbase = 0, total = 0;
#pragma unroll
  for (int w = 0; w < NWARP; w++) {
    if (w < warp) wbase += s_wtot[w];
    total += s_wtot[w];
  }
  if (tid == 0) {
    unsigned int base = 0xffffffffu;
    if (total > 0) {
      base = atomicAdd(&ctr->cand_used, total);
      if (base + total > pool_cap) {
        atomicOr(&ctr->status, (unsigned int)ST_CAND_OVERFLOW);
        base = 0xffffffffu;
      }
    } else {
      base = 0;
    }
    *s_base = base;
    cand_ref[spi] = make_uint2(base, total);
  }
  __syncthreads();
  const unsigned int base = *s_base;
  if (base != 0xffffffffu && total > 0) {
    unsigned int o = base + wbase + (incl - mine);
#pragma unroll
    for (int r = 0; r < IPT; r++) {
      const int w = tid * IPT + r;
      if (w < NW) {
        uint32_t m = mask[w];
        while (m) {
          const int bit = __ffs(m) - 1;
          m &= m - 1;
          cand_pool[o++] = (uint16_t)(w * 32 + bit);
        }
      }
    }
  }
  __syncthreads();  // mask / s_base are reused by the next strand
}

// FindTop over component COMP of a swizzled complex buffer holding the unscaled inverse transform
template <int LOG2N, int NT, int COMP>
__device__ __forceinline__ void findtop(const float2 *buf, int off, float scale, double co, uint32_t *mask,
                                        unsigned int *s_wtot, unsigned int *s_base, int spi,
                                        uint16_t *__restrict__ cand_pool, unsigned int pool_cap,
                                        uint2 *__restrict__ cand_ref, BatchCounters *ctr, float *__restrict__ xc_tap) {
  constexpr int N = 1 << LOG2N;
  auto xc_at = [&](int i) -> float {
    const float2 v = buf[swz((i + off) & (N - 1))];
    return (COMP ? v.y : v.x) * scale;
  };
  findtop_impl<LOG2N, NT>(xc_at, co, mask, s_wtot, s_base, spi, cand_pool, pool_cap, cand_ref, ctr, xc_tap);
}

// the two H-point inverses, then the radix-2 combine x[n] = e[n] + w_N^{-n} o[n], x[n+H] = e[n] - w_N^{-n} o[n]
// in place (natural order, swizzled slots).  Ends with a barrier.
template <int LOG2N, int NT>
__device__ __forceinline__ void inverse_full(float2 *buf, const float2 *__restrict__ wn, int tid, const float2 *tw) {
  constexpr int N = 1 << LOG2N, H = N / 2;
  fft_inverse_halves<LOG2N, N, NT>(buf, tid, tw);
  for (int n = tid; n < H; n += NT) {
    const float2 e = buf[swz(n)], o = buf[H + swz(n)];
    const float2 t = cmulc(o, __ldg(wn + n));
    buf[swz(n)] = cadd(e, t);
    buf[H + swz(n)] = csub(e, t);
  }
  __syncthreads();
}

// ---- spectral product of a chunk pair from the stored spectra, bin pair (k, N - k) by bin pair -----------------
// Where the spectra of a signal live: Z1 = spectrum of (A + iC); Z3 = spectrum of (G + iT) (four-channel form) or
// of (G_owner + i G_member) shared with another signal (three-channel form, sx_kernels.h).
struct SpecSrc {
  const float2 *z1, *z3;
  int mode;  // ZM_FOUR / ZM_RE / ZM_IM
};
__device__ __forceinline__ SpecSrc spec_src(const Slots &ws, int slot, const SlotMeta &m, int n) {
  SpecSrc r;
  r.z1 = ws.spec + ((size_t)slot * 2) * n;
  r.z3 = ws.spec + ((size_t)(m.zmode == ZM_FOUR ? slot : m.zslot) * 2 + 1) * n;
  r.mode = m.zmode;
  return r;
}
// Twice the per-channel spectra at bin k from the packed values at bin k (a) and at its mirror N - k (b):
// a real sequence has X[N-k] = conj X[k], so for Z = X + iY: 2X = Za + conj(Zb), 2Y = -i (Za - conj(Zb)).
// Three-channel form: T = -(A + C + G).
struct Chan4 { float2 a, c, g, t; };
template <int MODE>
__device__ __forceinline__ Chan4 channels2(float2 z1a, float2 z1b, float2 z3a, float2 z3b) {
  Chan4 r;
  r.a = make_float2(z1a.x + z1b.x, z1a.y - z1b.y);
  r.c = make_float2(z1a.y + z1b.y, z1b.x - z1a.x);
  const float2 re = make_float2(z3a.x + z3b.x, z3a.y - z3b.y), im = make_float2(z3a.y + z3b.y, z3b.x - z3a.x);
  if (MODE == ZM_FOUR) {
    r.g = re;
    r.t = im;
  } else {
    r.g = MODE == ZM_RE ? re : im;
    r.t = make_float2(-(r.a.x + r.c.x + r.g.x), -(r.a.y + r.c.y + r.g.y));
  }
  return r;
}
// One bin pair -> the values the inverse transform gets at bin k (oa) and at N - k (ob), times 4:
//   forward strand  Pf[k]  = conj(At) Aq + conj(Ct) Cq + conj(Gt) Gq + conj(Tt) Tq
//   reverse strand  the reverse-complement signal is the forward one reversed with the channels swapped A<->T,
//                   C<->G (exact when the entropy windows line up, which the host checks); reversed about index 0
//                   its channel spectra are conj(Tq), conj(Gq), conj(Cq), conj(Aq), so
//                   Pr'[k] = conj(At Tq + Ct Gq + Gt Cq + Tt Aq), and its correlation is the true one rotated by
//                   qlen - 1 lags.
// Both are spectra of real sequences (P[N-k] = conj P[k]); BOTH packs them as Pf + i Pr' so that ONE inverse
// transform gives the forward strand in its real part and the reverse strand in its imaginary part.
template <bool BOTH>
__device__ __forceinline__ void pair_product(const Chan4 &t, const Chan4 &q, float2 &oa, float2 &ob) {
  float2 pf;
  pf.x = (t.a.x * q.a.x + t.a.y * q.a.y) + (t.c.x * q.c.x + t.c.y * q.c.y) + (t.g.x * q.g.x + t.g.y * q.g.y) +
         (t.t.x * q.t.x + t.t.y * q.t.y);
  pf.y = (t.a.x * q.a.y - t.a.y * q.a.x) + (t.c.x * q.c.y - t.c.y * q.c.x) + (t.g.x * q.g.y - t.g.y * q.g.x) +
         (t.t.x * q.t.y - t.t.y * q.t.x);
  if (BOTH) {
    float2 rr;
    rr.x = (t.a.x * q.t.x - t.a.y * q.t.y) + (t.c.x * q.g.x - t.c.y * q.g.y) + (t.g.x * q.c.x - t.g.y * q.c.y) +
           (t.t.x * q.a.x - t.t.y * q.a.y);
    rr.y = (t.a.x * q.t.y + t.a.y * q.t.x) + (t.c.x * q.g.y + t.c.y * q.g.x) + (t.g.x * q.c.y + t.g.y * q.c.x) +
           (t.t.x * q.a.y + t.t.y * q.a.x);
    oa = make_float2(pf.x + rr.y, pf.y + rr.x);  // Pf + i conj(Rr)
    ob = make_float2(pf.x - rr.y, rr.x - pf.y);  // conj(Pf) + i Rr
  } else {
    oa = pf;
    ob = make_float2(pf.x, -pf.y);
  }
}
// The whole product into the (swizzled, scrambled-order) transform buffer.  Two neighbouring slots (r, r ^ 1) per
// step with 16-byte loads: their mirror slots are neighbours too (even bins: top digit d <-> R - 1 - d while the
// lower digits are not all zero; odd bins: r <-> H - 1 - r), and the swizzle only XORs the low four bits with a
// per-block constant, so pairs stay pairs.  The caller synchronises.
// half = -1: both halves into a buffer of N slots; 0 / 1: only the even / odd bins, into a buffer of H slots (a CTA that
// holds one H-point transform, N = 32768)
template <int LOG2N, int NT, bool BOTH, int TM, int QM>
__device__ __forceinline__ void spectral_product_m(float2 *buf, const SpecSrc T, const SpecSrc Q, int tid, int half) {
  constexpr int N = 1 << LOG2N, H = N / 2;
  constexpr int LR = last_radix<LOG2N>();
  constexpr int PPB = LR / 4;  // slot pairs per block of LR slots with top digit < LR / 2
  const int it0 = half == 1 ? H / 4 : 0, it1 = half == 0 ? H / 4 : H / 2;
  const int boff = half == 1 ? H : 0;  // slot offset of the buffer's first element
  const bool same3 = T.z3 == Q.z3;  // partners: one shared G transform
  const float4 *T1 = reinterpret_cast<const float4 *>(T.z1), *T3 = reinterpret_cast<const float4 *>(T.z3);
  const float4 *Q1 = reinterpret_cast<const float4 *>(Q.z1), *Q3 = reinterpret_cast<const float4 *>(Q.z3);
  auto lo = [](const float4 &v) { return make_float2(v.x, v.y); };
  auto hi = [](const float4 &v) { return make_float2(v.z, v.w); };
#pragma unroll 2
  for (int it = it0 + tid; it < it1; it += NT) {
    int pa0, pb0;
    if (it < H / 4) {  // even bins: m <-> (H - m) mod H; m < H/2 <=> top digit (last in scrambled order) < LR/2
      const int r = (it / PPB) * LR + 2 * (it % PPB);
      if (r < LR) continue;  // lower digits all zero: mirrors are d <-> R - d, done one by one below
      const int m2 = (H - natural_bin<LOG2N>(r)) & (H - 1);
      pa0 = swz(r);
      pb0 = swz(scrambled_pos<LOG2N>(m2));
    } else {  // odd bins: m <-> H - 1 - m, i.e. scrambled position r <-> H - 1 - r
      const int r = 2 * (it - H / 4);
      pa0 = H + swz(r);
      pb0 = H + swz(H - 1 - r);
    }
    const int qa = pa0 >> 1, qb = pb0 >> 1;  // float4 index
    const float4 t1a = __ldg(T1 + qa), t3a = __ldg(T3 + qa), q1a = __ldg(Q1 + qa);
    float4 t1b = __ldg(T1 + qb), t3b = __ldg(T3 + qb), q1b = __ldg(Q1 + qb);
    float4 q3a = t3a, q3b = t3b;
    if (!same3) {
      q3a = __ldg(Q3 + qa);
      q3b = __ldg(Q3 + qb);
    }
    // slot r sits in half (pa0 & 1) of the a-quad and its mirror in half (pb0 & 1) of the b-quad; slot r ^ 1 and
    // its mirror sit in the other halves: the lower slot of the a-quad mirrors the lower slot of the b-quad when the
    // two parities agree, the upper one otherwise -- one conditional swap of the b-quads instead of per-value selects
    const bool x = (pa0 ^ pb0) & 1;
    auto swp = [&](float4 &v) {
      if (x) v = make_float4(v.z, v.w, v.x, v.y);
    };
    swp(t1b);
    swp(t3b);
    swp(q1b);
    if (!same3) swp(q3b); else q3b = t3b;
    float2 oa0, ob0, oa1, ob1;
    pair_product<BOTH>(channels2<TM>(lo(t1a), lo(t1b), lo(t3a), lo(t3b)), channels2<QM>(lo(q1a), lo(q1b), lo(q3a), lo(q3b)),
                       oa0, ob0);
    pair_product<BOTH>(channels2<TM>(hi(t1a), hi(t1b), hi(t3a), hi(t3b)), channels2<QM>(hi(q1a), hi(q1b), hi(q3a), hi(q3b)),
                       oa1, ob1);
    reinterpret_cast<float4 *>(buf - boff)[qa] = make_float4(oa0.x, oa0.y, oa1.x, oa1.y);
    reinterpret_cast<float4 *>(buf - boff)[qb] = x ? make_float4(ob1.x, ob1.y, ob0.x, ob0.y) : make_float4(ob0.x, ob0.y, ob1.x, ob1.y);
  }
  if (half == 1) return;
  for (int r = tid; r < LR / 2; r += NT) {  // first block of the even bins, slot by slot (bins 0 and H/2 mirror themselves)
    const int m2 = (H - natural_bin<LOG2N>(r)) & (H - 1);
    const int pa = swz(r), pb = swz(scrambled_pos<LOG2N>(m2));
    float2 oa, ob;
    pair_product<BOTH>(channels2<TM>(__ldg(T.z1 + pa), __ldg(T.z1 + pb), __ldg(T.z3 + pa), __ldg(T.z3 + pb)),
                       channels2<QM>(__ldg(Q.z1 + pa), __ldg(Q.z1 + pb), __ldg(Q.z3 + pa), __ldg(Q.z3 + pb)), oa, ob);
    buf[pa] = oa;
    buf[pb] = ob;
  }
}
// the loop specialised for the forms of both chunks (uniform over the CTA)
template <int LOG2N, int NT, bool BOTH>
__device__ __forceinline__ void spectral_product(float2 *buf, const SpecSrc T, const SpecSrc Q, int tid, int half = -1) {
#define SX_PM(TM, QM) spectral_product_m<LOG2N, NT, BOTH, TM, QM>(buf, T, Q, tid, half)
  switch (T.mode * 3 + Q.mode) {
    case ZM_FOUR * 3 + ZM_FOUR: SX_PM(ZM_FOUR, ZM_FOUR); break;
    case ZM_FOUR * 3 + ZM_RE: SX_PM(ZM_FOUR, ZM_RE); break;
    case ZM_FOUR * 3 + ZM_IM: SX_PM(ZM_FOUR, ZM_IM); break;
    case ZM_RE * 3 + ZM_FOUR: SX_PM(ZM_RE, ZM_FOUR); break;
    case ZM_RE * 3 + ZM_RE: SX_PM(ZM_RE, ZM_RE); break;
    case ZM_RE * 3 + ZM_IM: SX_PM(ZM_RE, ZM_IM); break;
    case ZM_IM * 3 + ZM_FOUR: SX_PM(ZM_IM, ZM_FOUR); break;
    case ZM_IM * 3 + ZM_RE: SX_PM(ZM_IM, ZM_RE); break;
    default: SX_PM(ZM_IM, ZM_IM); break;
  }
#undef SX_PM
}

template <int LOG2N, int NT>
__global__ void __launch_bounds__(NT, (LOG2N <= 13 ? 3 : 1))
    xcorr_pair_kernel(const uint32_t *__restrict__ pair_list, const SpDesc *__restrict__ sps, Slots ws, double cutoff,
                      double cutoff_fast, uint16_t *__restrict__ cand_pool, unsigned int pool_cap,
                      uint2 *__restrict__ cand_ref, BatchCounters *ctr, float *__restrict__ xc_tap,
                      const unsigned int *__restrict__ n_listed) {
  constexpr int N = 1 << LOG2N, H = N / 2, NW = N / 32, NWARP = NT / 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2 *buf = reinterpret_cast<float2 *>(smem_raw);
  uint32_t *mask = reinterpret_cast<uint32_t *>(smem_raw + (size_t)N * sizeof(float2));  // NW words
  float2 *s_tw = reinterpret_cast<float2 *>(mask + NW);  // TwTables<LOG2N>::TOTAL inverse-pass twiddles
  __shared__ unsigned int s_wtot[NWARP];
  __shared__ unsigned int s_base;

  const int tid = threadIdx.x;
  TwTables<LOG2N>::template load<NT>(s_tw, tid);  // used after the barriers below
  // n_listed == nullptr: one CTA per entry of pair_list; otherwise a fixed grid strides over the *n_listed entries of a
  // list written on the device (the fused kernel's fall-backs, normally none).  FindTop ends with a barrier.
  for (int jb = blockIdx.x; n_listed != nullptr ? jb < (int)*n_listed : jb == (int)blockIdx.x; jb += gridDim.x) {
  const int spi = (int)pair_list[jb];  // forward strand-pair; the reverse one is spi + 1
  const SpDesc sp = sps[spi];
  const SlotMeta tm = ws.meta[sp.t_slot], qm = ws.meta[sp.q_slot];
  const int qlen = qm.len;

  // ---- products of both strands packed for one inverse transform: buf = 4 (Pf + i Pr') --------------------------
  spectral_product<LOG2N, NT, true>(buf, spec_src(ws, sp.t_slot, tm, N), spec_src(ws, sp.q_slot, qm, N), tid);
  __syncthreads();
  // ---- reference quirk (CrossCorr.cc:480-492): bins H-1 and H keep the TARGET spectrum, on both strands;
  //      the reverse strand's copy carries the rotation phase e^{+2 pi i k (qlen-1) / N}
  if (tid == 0) {
    constexpr int PH1 = bin_slot<LOG2N>(H - 1), PH = bin_slot<LOG2N>(H), PH2 = bin_slot<LOG2N>(H + 1);
    const int rot = qlen - 1;
    auto put = [&](int slot, int k, float2 t) {
      float sn, cs;
      sincospif(2.0f * (float)(((long long)k * rot) & (N - 1)) / (float)N, &sn, &cs);
      const float2 tr = cmul(t, make_float2(cs, sn));
      buf[slot] = make_float2(4.f * (t.x - tr.y), 4.f * (t.y + tr.x));  // 4 (t + i tr)
    };
    put(PH1, H - 1, make_float2(tm.q_re, tm.q_im));
    put(PH2, H + 1, make_float2(tm.q_re, -tm.q_im));
    put(PH, H, make_float2(tm.q_nyq, 0.f));
  }
  __syncthreads();
  inverse_full<LOG2N, NT>(buf, ws.wn, tid, s_tw);

  // xc[i] = x[(i + H) mod N] / N (rescale + half rotation, CrossCorr.cc:493-505); the factor 4 above
  const float scale = 0.25f / (float)N;
  const double co = (sp.flags & SP_FAST) ? cutoff_fast : cutoff;
  findtop<LOG2N, NT, 0>(buf, H, scale, co, mask, s_wtot, &s_base, spi, cand_pool, pool_cap, cand_ref, ctr, xc_tap);
  findtop<LOG2N, NT, 1>(buf, (H - (qlen - 1)) & (N - 1), scale, co, mask, s_wtot, &s_base, spi + 1, cand_pool,
                        pool_cap, cand_ref, ctr, xc_tap);
  }
}

// FindTop alone over a correlation vector the CALLER supplies (SeqAnalyzer::MatchUp takes `xc` as an argument,
// analysis/CrossCorr.h:261): the candidates of strand-pair `spi` come from it instead of from the pair kernel.
template <int LOG2N, int NT>
__global__ void __launch_bounds__(NT)
    findtop_external_kernel(const float *__restrict__ xc, int spi, double co, uint16_t *__restrict__ cand_pool,
                            unsigned int pool_cap, uint2 *__restrict__ cand_ref, BatchCounters *ctr) {
  constexpr int N = 1 << LOG2N, NW = N / 32, NWARP = NT / 32;
  __shared__ uint32_t mask[NW];
  __shared__ unsigned int s_wtot[NWARP];
  __shared__ unsigned int s_base;
  auto xc_at = [&](int i) -> float { return xc[i]; };
  findtop_impl<LOG2N, NT>(xc_at, co, mask, s_wtot, &s_base, spi, cand_pool, pool_cap, cand_ref, ctr, nullptr);
}

template <int LOG2N, int NT>
__global__ void __launch_bounds__(NT)
    xcorr_findtop_kernel(const uint32_t *__restrict__ direct_list, const SpDesc *__restrict__ sps, Slots ws,
                         double cutoff, double cutoff_fast, uint16_t *__restrict__ cand_pool, unsigned int pool_cap,
                         uint2 *__restrict__ cand_ref, BatchCounters *ctr, float *__restrict__ xc_tap) {
  constexpr int N = 1 << LOG2N, H = N / 2, NW = N / 32, NWARP = NT / 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2 *buf = reinterpret_cast<float2 *>(smem_raw);
  uint32_t *mask = reinterpret_cast<uint32_t *>(smem_raw + (size_t)N * sizeof(float2));  // NW words
  float2 *s_tw = reinterpret_cast<float2 *>(mask + NW);  // TwTables<LOG2N>::TOTAL inverse-pass twiddles
  __shared__ unsigned int s_wtot[NWARP];
  __shared__ unsigned int s_base;

  const int tid = threadIdx.x;
  TwTables<LOG2N>::template load<NT>(s_tw, tid);  // used after the barriers below
  const int spi = (int)direct_list[blockIdx.x];
  const SpDesc sp = sps[spi];
  const SlotMeta tm = ws.meta[sp.t_slot];

  // ---- product of one strand: buf = 4 Pf (the spectrum of a real sequence) ------------------------------------
  spectral_product<LOG2N, NT, false>(buf, spec_src(ws, sp.t_slot, tm, N), spec_src(ws, sp.q_slot, ws.meta[sp.q_slot], N), tid);
  __syncthreads();
  // ---- reference quirk (CrossCorr.cc:480-492): bins H-1 and H keep the TARGET spectrum ------------
  if (tid == 0) {
    constexpr int PH1 = bin_slot<LOG2N>(H - 1), PH = bin_slot<LOG2N>(H), PH2 = bin_slot<LOG2N>(H + 1);
    buf[PH1] = make_float2(4.f * tm.q_re, 4.f * tm.q_im);
    buf[PH2] = make_float2(4.f * tm.q_re, -4.f * tm.q_im);
    buf[PH] = make_float2(4.f * tm.q_nyq, 0.f);
  }
  __syncthreads();
  inverse_full<LOG2N, NT>(buf, ws.wn, tid, s_tw);

  // xc[i] = Re x[(i + H) mod N] / N   (rescale + half rotation, CrossCorr.cc:493-505); the factor 4 above
  const float scale = 0.25f / (float)N;
  const double co = (sp.flags & SP_FAST) ? cutoff_fast : cutoff;
  findtop<LOG2N, NT, 0>(buf, H, scale, co, mask, s_wtot, &s_base, spi, cand_pool, pool_cap, cand_ref, ctr, xc_tap);
}

// =================================================================================================
// K1 + K2 fused for chunk pairs whose spectra nobody else needs (pair mode: independent chunk pairs, the guided
// refinement pass): one CTA per chunk pair, the spectra never leave the SM.
//   Three-channel form: the pair needs three complex N-point transforms, (A + iC) of the target, (A + iC) of the
//   query and (G_target + i G_query).  Every N-point transform is two independent H-point transforms (even / odd bins,
//   sx_fft.cuh) and the spectral product pairs bin k with N - k, which stay inside a half.  So the CTA works half by
//   half: three thread groups of 128 transform one signal each into three 32 KiB buffers (own named barriers, they
//   exchange nothing), the whole CTA forms the product in place and runs its H-point inverse.  The inverse of the even
//   half (32 KiB) is parked in an L2-resident scratch line of this CTA while the odd half goes through the same
//   buffers; the radix-2 combine then writes both strands' correlation vectors over the two free buffers and FindTop
//   runs on them.  96 KiB of transform buffers -> two CTAs (24 warps) per SM; per chunk pair 8 KB of bases are read
//   and 64 KB go through L2, against 384 KB through HBM for separate kernels.
//   Chunks the preparation kernel did not accept (a letter other than A/C/G/T) are handed back: the pair is appended to
//   the fall-back lists, which the separate kernels work off right behind this one.
// =================================================================================================
#define SX_FUSED_NT 384
#define SX_FUSED_NG 128
template <int LOG2N>
struct FusedCfg {
  static constexpr int N = 1 << LOG2N, H = N / 2, NW = N / 32;
  static constexpr size_t SMEM = (size_t)3 * H * sizeof(float2) + (size_t)2 * 2 * NW * 4 + (size_t)2 * 256 * 4 +
                                 (size_t)TwTables<LOG2N>::TOTAL * sizeof(float2) + (size_t)NW * 4;
};

__device__ __forceinline__ void fused_group_barrier(int grp) {  // literal ids (a register id reserves all 16 barriers)
  if (grp == 0)
    asm volatile("bar.sync 1, %0;" ::"n"(SX_FUSED_NG) : "memory");
  else if (grp == 1)
    asm volatile("bar.sync 2, %0;" ::"n"(SX_FUSED_NG) : "memory");
  else
    asm volatile("bar.sync 3, %0;" ::"n"(SX_FUSED_NG) : "memory");
}

template <int LOG2N>
__global__ void __launch_bounds__(SX_FUSED_NT, 2)
    pair_fused_kernel(const FusedJob *__restrict__ jobs, int njobs, const SigDesc *__restrict__ sigs,
                      const SpDesc *__restrict__ sps, Slots ws, PrepBuf prep, float2 *__restrict__ scratch, double cutoff,
                      double cutoff_fast, uint16_t *__restrict__ cand_pool, unsigned int pool_cap,
                      uint2 *__restrict__ cand_ref, BatchCounters *ctr, FusedFail fail) {
  constexpr int N = 1 << LOG2N, H = N / 2, NW = N / 32, WIN = N / 512, NT = SX_FUSED_NT, NG = SX_FUSED_NG, NWARP = NT / 32;
  constexpr int LR = last_radix<LOG2N>(), PPB = LR / 4;
  using P = Plan<LOG2N>;
  using TW = TwTables<LOG2N>;
  constexpr int L1 = H / P::R0, L2 = L1 / P::R1, L3 = L2 / P::R2;
  static_assert(WIN <= 32, "one plane word per entropy window");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2 *buf = reinterpret_cast<float2 *>(smem_raw);                          // [3][H] swizzled slots
  uint32_t *s_pl = reinterpret_cast<uint32_t *>(buf + 3 * H);                  // [2 chunks][2 planes][NW]
  float *s_went = reinterpret_cast<float *>(s_pl + 2 * 2 * NW);                // [2][256] window weights
  float2 *s_tw = reinterpret_cast<float2 *>(s_went + 2 * 256);                 // inverse-pass twiddles
  uint32_t *mask = reinterpret_cast<uint32_t *>(s_tw + TW::TOTAL);             // NW words (FindTop)
  float *xf = reinterpret_cast<float *>(buf), *xr = reinterpret_cast<float *>(buf + H);  // N lags each, over buffers 0 / 1
  __shared__ double s_off[2][4];
  __shared__ unsigned int s_wtot[NWARP];
  __shared__ unsigned int s_base;

  const int tid = threadIdx.x, grp = tid / NG, gt = tid - grp * NG;
  const float2 *__restrict__ wn = ws.wn;
  TW::template load<NT>(s_tw, tid);
  float4 *park = reinterpret_cast<float4 *>(scratch + (size_t)blockIdx.x * H);  // this CTA's scratch line (L2-resident)

  for (int job = blockIdx.x; job < njobs; job += gridDim.x) {
    __syncthreads();  // the previous pair's shared-memory contents are dead
    const FusedJob J = jobs[job];
    if (prep.flag[J.t_sig] == 0 || prep.flag[J.q_sig] == 0) {  // not pure A/C/G/T (or otherwise not prepared): hand back
      if (tid == 0) {
        const unsigned int i = atomicAdd(fail.n_pairs, 1u);
        fail.pairs[i] = J.spi;
        const unsigned int k = atomicAdd(fail.n_sigs, 2u);
        fail.sigs[k] = J.t_sig;
        fail.sigs[k + 1] = J.q_sig;
      }
      continue;
    }
    const SigDesc td = sigs[J.t_sig], qd = sigs[J.q_sig];
    const int tlen = td.len, qlen = qd.len;
    {
      const uint32_t *pt = ws.planes + (size_t)td.slot * 2 * NW, *pq = ws.planes + (size_t)qd.slot * 2 * NW;
      for (int w = tid; w < 2 * NW; w += NT) {
        s_pl[w] = pt[w];
        s_pl[2 * NW + w] = pq[w];
      }
      const float *wt = prep.went + (size_t)J.t_sig * 256, *wq = prep.went + (size_t)J.q_sig * 256;
      for (int w = tid; w < 256; w += NT) {
        s_went[w] = wt[w];
        s_went[256 + w] = wq[w];
      }
      if (tid < 8) s_off[tid >> 2][tid & 3] = prep.off[(size_t)(tid < 4 ? J.t_sig : J.q_sig) * 4 + (tid & 3)];
    }
    __syncthreads();

    // what this thread's group transforms: group 0 (A + iC) of the target, 1 (A + iC) of the query, 2 (G_t + i G_q)
    const int sx_ = grp == 1 ? 1 : 0, sy = grp == 0 ? 0 : 1;  // chunk feeding the real / imaginary part
    const uint32_t c0 = grp == 2 ? 2u : 0u, c1 = grp == 2 ? 2u : 1u;
    const int len0 = sx_ ? qlen : tlen, len1 = sy ? qlen : tlen;
    const bool flat0 = len0 < 1024, flat1 = len1 < 1024;  // "Skip entropy": weight 1 everywhere (CrossCorr.cc:39-44)
    const double off0 = s_off[sx_][c0], off1 = s_off[sy][c1];
    const double hit0 = __dsub_rn(1.0, off0), miss0 = __dsub_rn(0.0, off0);
    const double hit1 = __dsub_rn(1.0, off1), miss1 = __dsub_rn(0.0, off1);
    float2 *gb = buf + grp * H;
    float4 *B0 = reinterpret_cast<float4 *>(buf), *B1 = reinterpret_cast<float4 *>(buf + H),
           *B2 = reinterpret_cast<float4 *>(buf + 2 * H);

#pragma unroll 1
    for (int half = 0; half < 2; half++) {
      // ---- signal of this half: e[n] = z[n] (even bins) / o[n] = z[n] w_N^n (odd bins); a chunk has z[n + H] = 0.
      //      Same arithmetic as encode_fft_kernel's fast path, sample for sample.
      for (int w = gt; w < H / WIN; w += NG) {
        const int k0 = w * WIN;
        const double e0 = flat0 ? 1.0 : (double)s_went[sx_ * 256 + w];
        const double e1 = flat1 ? 1.0 : (double)s_went[sy * 256 + w];
        const float h0 = __double2float_rn(__dmul_rn(e0, hit0)), m0 = __double2float_rn(__dmul_rn(e0, miss0));
        const float h1 = __double2float_rn(__dmul_rn(e1, hit1)), m1 = __double2float_rn(__dmul_rn(e1, miss1));
        const uint32_t *p0 = s_pl + sx_ * 2 * NW, *p1 = s_pl + sy * 2 * NW;
        const uint32_t lo0 = p0[k0 >> 5] >> (k0 & 31), hi0 = p0[NW + (k0 >> 5)] >> (k0 & 31);
        const uint32_t lo1 = p1[k0 >> 5] >> (k0 & 31), hi1 = p1[NW + (k0 >> 5)] >> (k0 & 31);
        const uint32_t sel0 = ((c0 & 1u) ? lo0 : ~lo0) & ((c0 & 2u) ? hi0 : ~hi0);
        const uint32_t sel1 = ((c1 & 1u) ? lo1 : ~lo1) & ((c1 & 2u) ? hi1 : ~hi1);
        const float2 wb = half ? __ldg(wn + k0) : make_float2(1.f, 0.f);
        // bases past the end of a chunk give zero samples: cut the selection masks once per window instead of
        // testing every sample (zero = neither "hit" nor "miss": a second mask)
        const uint32_t in0 = k0 + WIN <= len0 ? 0xffffffffu : (k0 < len0 ? (1u << (len0 - k0)) - 1u : 0u);
        const uint32_t in1 = k0 + WIN <= len1 ? 0xffffffffu : (k0 < len1 ? (1u << (len1 - k0)) - 1u : 0u);
        const bool full = (in0 & in1) == 0xffffffffu;
#pragma unroll
        for (int j = 0; j < WIN; j++) {
          const int k = k0 + j;
          float2 v;
          v.x = ((sel0 >> j) & 1u) ? h0 : m0;
          v.y = ((sel1 >> j) & 1u) ? h1 : m1;
          if (!full) {
            if (!((in0 >> j) & 1u)) v.x = 0.f;
            if (!((in1 >> j) & 1u)) v.y = 0.f;
          }
          if (half) {
            constexpr double ang = -2.0 * 3.14159265358979323846 / (double)N;
            const float2 st = make_float2((float)cx_cos_small(ang * j), (float)cx_sin_small(ang * j));
            const float2 wk = j == 0 ? wb : cmul(wb, st);
            v = cmul(v, wk);
          }
          gb[swz(k)] = v;
        }
      }
      fused_group_barrier(grp);
      // ---- forward H-point transform of this group's buffer
      fft_pass<H, H, P::R0, false, NG>(gb, gt);
      fused_group_barrier(grp);
      fft_pass<H, L1, P::R1, false, NG>(gb, gt);
      fused_group_barrier(grp);
      fft_pass<H, L2, P::R2, false, NG>(gb, gt);
      if constexpr (P::R3 > 1) {
        fused_group_barrier(grp);
        fft_pass<H, L3, P::R3, false, NG>(gb, gt);
      }
      __syncthreads();
      // ---- product of this half in place over buffer 2 (see spectral_product): 4 (Pf + i Pr')
      for (int it = tid; it < H / 4; it += NT) {
        int pa0, pb0;
        if (half == 0) {
          const int r = (it / PPB) * LR + 2 * (it % PPB);
          if (r < LR) continue;
          const int m2 = (H - natural_bin<LOG2N>(r)) & (H - 1);
          pa0 = swz(r);
          pb0 = swz(scrambled_pos<LOG2N>(m2));
        } else {
          const int r = 2 * it;
          pa0 = swz(r);
          pb0 = swz(H - 1 - r);
        }
        const int qa = pa0 >> 1, qb = pb0 >> 1;
        const float4 t1a = B0[qa], q1a = B1[qa], z3a = B2[qa];
        float4 t1b = B0[qb], q1b = B1[qb], z3b = B2[qb];
        const bool x = (pa0 ^ pb0) & 1;
        if (x) {
          t1b = make_float4(t1b.z, t1b.w, t1b.x, t1b.y);
          q1b = make_float4(q1b.z, q1b.w, q1b.x, q1b.y);
          z3b = make_float4(z3b.z, z3b.w, z3b.x, z3b.y);
        }
        auto lo = [](const float4 &v) { return make_float2(v.x, v.y); };
        auto hi = [](const float4 &v) { return make_float2(v.z, v.w); };
        float2 oa0, ob0, oa1, ob1;
        pair_product<true>(channels2<ZM_RE>(lo(t1a), lo(t1b), lo(z3a), lo(z3b)), channels2<ZM_IM>(lo(q1a), lo(q1b), lo(z3a), lo(z3b)),
                           oa0, ob0);
        pair_product<true>(channels2<ZM_RE>(hi(t1a), hi(t1b), hi(z3a), hi(z3b)), channels2<ZM_IM>(hi(q1a), hi(q1b), hi(z3a), hi(z3b)),
                           oa1, ob1);
        B2[qa] = make_float4(oa0.x, oa0.y, oa1.x, oa1.y);
        B2[qb] = x ? make_float4(ob1.x, ob1.y, ob0.x, ob0.y) : make_float4(ob0.x, ob0.y, ob1.x, ob1.y);
      }
      if (half == 0) {
        for (int r = tid; r < LR / 2; r += NT) {  // first block of the even bins, slot by slot
          const int m2 = (H - natural_bin<LOG2N>(r)) & (H - 1);
          const int pa = swz(r), pb = swz(scrambled_pos<LOG2N>(m2));
          float2 oa, ob;
          pair_product<true>(channels2<ZM_RE>(buf[pa], buf[pb], buf[2 * H + pa], buf[2 * H + pb]),
                             channels2<ZM_IM>(buf[H + pa], buf[H + pb], buf[2 * H + pa], buf[2 * H + pb]), oa, ob);
          buf[2 * H + pa] = oa;
          buf[2 * H + pb] = ob;
        }
      }
      __syncthreads();
      // reference quirk (CrossCorr.cc:480-492): bins H-1, H (and the mirror H+1) keep the target spectrum summed over
      // the channels -- zero in the three-channel form
      if (tid == 0) {
        if (half == 0) {
          buf[2 * H + bin_slot<LOG2N>(H)] = make_float2(0.f, 0.f);
        } else {
          buf[2 * H + bin_slot<LOG2N>(H - 1) - H] = make_float2(0.f, 0.f);
          buf[2 * H + bin_slot<LOG2N>(H + 1) - H] = make_float2(0.f, 0.f);
        }
      }
      __syncthreads();
      // ---- inverse H-point transform of the product
      if constexpr (P::R3 > 1) {
        fft_pass<H, L3, P::R3, true, NT, true>(buf + 2 * H, tid, s_tw + TW::OFF3);
        __syncthreads();
      }
      fft_pass<H, L2, P::R2, true, NT, true>(buf + 2 * H, tid, s_tw + TW::OFF2);
      __syncthreads();
      fft_pass<H, L1, P::R1, true, NT, true>(buf + 2 * H, tid, s_tw + TW::OFF1);
      __syncthreads();
      fft_pass<H, H, P::R0, true, NT, true>(buf + 2 * H, tid, s_tw + TW::OFF0);
      __syncthreads();
      if (half == 0) {  // park e[n]; every thread reads back exactly what it wrote
        for (int i = tid; i < H / 2; i += NT) park[i] = B2[i];
        __syncthreads();  // buffer 2 is generated into next
      }
    }
    // ---- radix-2 combine x[n] = e[n] + w_N^{-n} o[n], x[n + H] = e[n] - w_N^{-n} o[n]; forward strand = real part,
    //      reverse strand = imaginary part rotated by qlen - 1 lags; xc[i] = x[(i + H) mod N] / N (rescale + half
    //      rotation, CrossCorr.cc:493-505), the factor 4 of the product
    {
      const float scale = 0.25f / (float)N;
      const int offr = (H - (qlen - 1)) & (N - 1);
      for (int i = tid; i < H / 2; i += NT) {
        const float4 e4 = park[i], o4 = B2[i];
#pragma unroll
        for (int c = 0; c < 2; c++) {
          const int n = swz(2 * i + c);  // logical index of this physical slot (the swizzle is its own inverse)
          const float2 e = c ? make_float2(e4.z, e4.w) : make_float2(e4.x, e4.y);
          const float2 o = c ? make_float2(o4.z, o4.w) : make_float2(o4.x, o4.y);
          const float2 t = cmulc(o, __ldg(wn + n));
          const float2 lo = cadd(e, t), hi = csub(e, t);
          xf[n + H] = lo.x * scale;
          xf[n] = hi.x * scale;
          xr[(n - offr) & (N - 1)] = lo.y * scale;
          xr[(n + H - offr) & (N - 1)] = hi.y * scale;
        }
      }
    }
    __syncthreads();
    const SpDesc sp = sps[J.spi];
    const double co = (sp.flags & SP_FAST) ? cutoff_fast : cutoff;
    auto at_f = [&](int i) -> float { return xf[i]; };
    auto at_r = [&](int i) -> float { return xr[i]; };
    findtop_impl<LOG2N, NT>(at_f, co, mask, s_wtot, &s_base, (int)J.spi, cand_pool, pool_cap, cand_ref, ctr, nullptr);
    findtop_impl<LOG2N, NT>(at_r, co, mask, s_wtot, &s_base, (int)J.spi + 1, cand_pool, pool_cap, cand_ref, ctr, nullptr);
    if (tid < 2) {  // what the scan kernel reads of the two slots (the planes are the preparation kernel's)
      SlotMeta m;
      m.len = tid ? qlen : tlen;
      m.flags = 0;
      m.q_re = m.q_im = m.q_nyq = 0.f;
      m.zmode = ZM_FOUR;  // no stored spectrum
      m.zslot = 0;
      m.pad = 0;
      ws.meta[tid ? qd.slot : td.slot] = m;
    }
  }
}

// ---- K2 for transforms that do not fit one CTA (N = 32768) ------------------------------------------
// xcorr_half_kernel: grid (jobs, 2); job < n_pairs is a chunk pair handled like xcorr_pair_kernel, the
// rest are single strand-pairs handled like xcorr_findtop_kernel; blockIdx.y = half (even / odd bins).
// Product, quirk bins and the H-point inverse of one half in shared memory; the natural-order result
// e[n] (half 0) / o[n] (half 1) goes to a scratch buffer in HBM (L2-resident for the next kernel).
// combine_findtop_kernel: one CTA per strand-pair: x[n] = e[n] + w_N^{-n} o[n], x[n+H] = e[n] - w_N^{-n} o[n],
// the component of its strand, rescale + rotation, then FindTop on the real vector in shared memory.
template <int LOG2N, int NT>
__global__ void __launch_bounds__(NT, 1)
    xcorr_half_kernel(const uint32_t *__restrict__ pair_list, int n_pairs, const uint32_t *__restrict__ direct_list,
                      const SpDesc *__restrict__ sps, Slots ws, float2 *__restrict__ scratch) {
  constexpr int N = 1 << LOG2N, H = N / 2;
  constexpr int LR = last_radix<LOG2N>();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2 *buf = reinterpret_cast<float2 *>(smem_raw);  // H complex: this CTA's half
  float2 *s_tw = buf + H;                              // TwTables<LOG2N>::TOTAL inverse-pass twiddles
  const int tid = threadIdx.x, half = blockIdx.y, job = blockIdx.x;
  TwTables<LOG2N>::template load<NT>(s_tw, tid);  // used after the barriers below
  const bool is_pair = job < n_pairs;
  const int spi = (int)(is_pair ? pair_list[job] : direct_list[job - n_pairs]);
  const SpDesc sp = sps[spi];
  const SlotMeta tm = ws.meta[sp.t_slot];
  const int qlen = ws.meta[sp.q_slot].len;
  const float2 *U1 = ws.spec + ((size_t)sp.t_slot * 2) * N + (size_t)half * H, *U2 = U1 + N;
  const float2 *V1 = ws.spec + ((size_t)sp.q_slot * 2) * N + (size_t)half * H, *V2 = V1 + N;
  constexpr int PH1 = bin_slot<LOG2N>(H - 1) - H, PH2 = bin_slot<LOG2N>(H + 1) - H, PH = bin_slot<LOG2N>(H);
  if (is_pair) {
    // both strands, Hermitian-symmetrised and packed (see xcorr_pair_kernel)
#pragma unroll 2
    for (int it = tid; it < H / 2; it += NT) {
      int pa, pb;
      if (half == 0) {  // even bins: m <-> (H - m) mod H
        const int r = (it / (LR / 2)) * LR + (it % (LR / 2));
        const int m2 = (H - natural_bin<LOG2N>(r)) & (H - 1);
        pa = swz(r);
        pb = swz(scrambled_pos<LOG2N>(m2));
      } else {  // odd bins: scrambled position r <-> H - 1 - r
        pa = swz(it);
        pb = swz(H - 1 - it);
      }
      const float2 u1a = __ldg(U1 + pa), u2a = __ldg(U2 + pa), v1a = __ldg(V1 + pa), v2a = __ldg(V2 + pa);
      const float2 u1b = __ldg(U1 + pb), u2b = __ldg(U2 + pb), v1b = __ldg(V1 + pb), v2b = __ldg(V2 + pb);
      const float2 a = cadd(cmulc(v1a, u1a), cmulc(v2a, u2a));
      const float2 b = cadd(cmulc(v1b, u1b), cmulc(v2b, u2b));
      const float2 g = cadd(cmul(u1a, v2a), cmul(u2a, v1a));
      const float2 h = cadd(cmul(u1b, v2b), cmul(u2b, v1b));
      const float sx_ = a.x + b.x, dx = g.x - h.x, sy = g.y + h.y, dy = a.y - b.y;
      buf[pa] = make_float2(sx_ - dx, sy + dy);
      buf[pb] = make_float2(sx_ + dx, sy - dy);
    }
    __syncthreads();
    if (tid == 0) {  // quirk bins with the reverse strand's rotation phase (see xcorr_pair_kernel)
      const int rot = qlen - 1;
      auto put = [&](int slot, int k, float2 t) {
        float sn, cs;
        sincospif(2.0f * (float)(((long long)k * rot) & (N - 1)) / (float)N, &sn, &cs);
        const float2 tr = cmul(t, make_float2(cs, sn));
        buf[slot] = make_float2(2.f * (t.x - tr.y), 2.f * (t.y + tr.x));
      };
      if (half == 1) {
        put(PH1, H - 1, make_float2(tm.q_re, tm.q_im));
        put(PH2, H + 1, make_float2(tm.q_re, -tm.q_im));
      } else {
        put(PH, H, make_float2(tm.q_nyq, 0.f));
      }
    }
  } else {
    const float4 *u1 = reinterpret_cast<const float4 *>(U1), *u2 = reinterpret_cast<const float4 *>(U2);
    const float4 *v1 = reinterpret_cast<const float4 *>(V1), *v2 = reinterpret_cast<const float4 *>(V2);
    float4 *dst = reinterpret_cast<float4 *>(buf);
#pragma unroll 2
    for (int k = tid; k < H / 2; k += NT) {
      const float4 a1 = __ldg(u1 + k), a2 = __ldg(u2 + k), b1 = __ldg(v1 + k), b2 = __ldg(v2 + k);
      float4 p;
      p.x = (b1.x * a1.x + b1.y * a1.y) + (b2.x * a2.x + b2.y * a2.y);
      p.y = (b1.y * a1.x - b1.x * a1.y) + (b2.y * a2.x - b2.x * a2.y);
      p.z = (b1.z * a1.z + b1.w * a1.w) + (b2.z * a2.z + b2.w * a2.w);
      p.w = (b1.w * a1.z - b1.z * a1.w) + (b2.w * a2.z - b2.z * a2.w);
      dst[k] = p;
    }
    __syncthreads();
    if (tid == 0) {
      if (half == 1) {
        buf[PH1] = make_float2(tm.q_re, tm.q_im);
        buf[PH2] = make_float2(tm.q_re, -tm.q_im);
      } else {
        buf[PH] = make_float2(tm.q_nyq, 0.f);
      }
    }
  }
  __syncthreads();
  if constexpr (LOG2N == 15) drift_correct_half<LOG2N, NT, true>(buf, ws.drift, half, tid);
  fft_inverse_halves<LOG2N, H, NT>(buf, tid, s_tw);
  float2 *out = scratch + ((size_t)job * 2 + half) * H;
  for (int n = tid; n < H; n += NT) out[n] = buf[swz(n)];
}

// ---- N = 32768 as ONE kernel: a cluster of two CTAs per strand-pair, one per half of the radix-2 split ----------
// Product, quirk bins, drift pre-correction and the H-point inverse of its half as in xcorr_half_kernel; then the
// radix-2 combine through DISTRIBUTED SHARED MEMORY: CTA 0 holds e[n], CTA 1 holds o[n]; index n is handled by exactly
// one thread of the cluster (CTA 0 the lower half of the indices, CTA 1 the upper), which reads e[n] and o[n] -- one of
// them from the other SM -- and writes x[n] = e[n] + w_N^{-n} o[n] over e[n] and x[n + H] = e[n] - w_N^{-n} o[n] over
// o[n].  No scratch buffer, no second kernel.  After the rescale + half rotation lag i is x[(i + H) mod N], so CTA 1
// now holds lags [0, H) and CTA 0 lags [H, N): each runs FindTop over its own lags (RMS envelope per 256 lags,
// threshold, ballot mask) and the two compact into ONE ascending candidate list: they exchange their counts, the CTA
// with the lower lags reserves the pool entries for both.
template <int LOG2N, int NT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NT, 1)
    xcorr_cluster_kernel(const uint32_t *__restrict__ direct_list, const SpDesc *__restrict__ sps, Slots ws, double cutoff,
                         double cutoff_fast, uint16_t *__restrict__ cand_pool, unsigned int pool_cap,
                         uint2 *__restrict__ cand_ref, BatchCounters *ctr, float *__restrict__ xc_tap) {
  namespace cg = cooperative_groups;
  constexpr int N = 1 << LOG2N, H = N / 2, NWL = H / 32, NBL = H / 256, NWARP = NT / 32;
  static_assert(NWL <= NT, "one mask word per thread");
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2 *buf = reinterpret_cast<float2 *>(smem_raw);         // H complex: this CTA's half
  float2 *s_tw = buf + H;                                      // TwTables<LOG2N>::TOTAL inverse-pass twiddles
  uint32_t *mask = reinterpret_cast<uint32_t *>(s_tw + TwTables<LOG2N>::TOTAL);  // NWL words: this CTA's lags
  __shared__ unsigned int s_wtot[NWARP];
  __shared__ unsigned int s_peer_total, s_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int half = (int)cluster.block_rank(), job = blockIdx.x >> 1;
  TwTables<LOG2N>::template load<NT>(s_tw, tid);  // used after the barriers below
  const int spi = (int)direct_list[job];
  const SpDesc sp = sps[spi];
  const SlotMeta tm = ws.meta[sp.t_slot];
  {
    // product of one strand over this half's bins, bin pair (k, N - k) by bin pair (both in this half), from the
    // three- or four-channel spectra of the two chunks: 4 Pf, the spectrum of a real sequence (see spectral_product)
    spectral_product<LOG2N, NT, false>(buf, spec_src(ws, sp.t_slot, tm, N), spec_src(ws, sp.q_slot, ws.meta[sp.q_slot], N), tid, half);
    __syncthreads();
    if (tid == 0) {  // reference quirk (CrossCorr.cc:480-492): bins H-1 and H keep the TARGET spectrum
      constexpr int PH1 = bin_slot<LOG2N>(H - 1) - H, PH2 = bin_slot<LOG2N>(H + 1) - H, PH = bin_slot<LOG2N>(H);
      if (half == 1) {
        buf[PH1] = make_float2(4.f * tm.q_re, 4.f * tm.q_im);
        buf[PH2] = make_float2(4.f * tm.q_re, -4.f * tm.q_im);
      } else {
        buf[PH] = make_float2(4.f * tm.q_nyq, 0.f);
      }
    }
    __syncthreads();
  }
  if constexpr (LOG2N == 15) drift_correct_half<LOG2N, NT, true>(buf, ws.drift, half, tid);
  fft_inverse_halves<LOG2N, H, NT>(buf, tid, s_tw);  // natural order, swizzled slots; ends with a barrier
  cluster.sync();
  {
    float2 *peer = cluster.map_shared_rank(buf, half ^ 1);
    float2 *eb = half == 0 ? buf : peer, *ob = half == 0 ? peer : buf;  // CTA 0's buffer: e -> x[0, H); CTA 1's: o -> x[H, N)
    const float2 *__restrict__ wn = ws.wn;
    for (int n = half * (H / 2) + tid; n < (half + 1) * (H / 2); n += NT) {
      const int slot = swz(n);
      const float2 e = eb[slot], o = ob[slot];
      const float2 t = cmulc(o, __ldg(wn + n));
      eb[slot] = cadd(e, t);
      ob[slot] = csub(e, t);
    }
  }
  cluster.sync();
  // ---- FindTop over this CTA's lags: lag = lag0 + j with xc = Re buf[j] / N
  const int lag0 = half == 1 ? 0 : H;
  const float scale = 0.25f / (float)N;  // the factor 4 of the product
  const double co = (sp.flags & SP_FAST) ? cutoff_fast : cutoff;
  auto xc_at = [&](int j) -> float { return buf[swz(j)].x * scale; };
  if (xc_tap != nullptr) {
    float *o = xc_tap + (size_t)spi * N + lag0;
    for (int j = tid; j < H; j += NT) o[j] = xc_at(j);
  }
  for (int b = warp; b < NBL; b += NWARP) {  // see findtop_impl
    float v[8];
    double acc = 0.;
#pragma unroll
    for (int r = 0; r < 8; r++) {
      v[r] = xc_at(b * 256 + r * 32 + lane);
      acc += (double)__fmul_rn(v[r], v[r]);
    }
    acc = warp_sum(acc);
    const double thr = __dadd_rn(__dmul_rn(__dsqrt_rn(__ddiv_rn(acc, 256.0)), co), 1.0);
    const float thr_f = __double2float_rd(thr);
#pragma unroll
    for (int r = 0; r < 8; r++) {
      const uint32_t m = __ballot_sync(0xffffffffu, v[r] > thr_f);
      if (lane == 0) mask[b * 8 + r] = m;
    }
  }
  __syncthreads();
  const unsigned int mine = tid < NWL ? __popc(mask[tid]) : 0u;
  unsigned int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) s_wtot[warp] = incl;
  __syncthreads();
  unsigned int wbase = 0, total = 0;
#pragma unroll
  for (int w = 0; w < NWARP; w++) {
    if (w < warp) wbase += s_wtot[w];
    total += s_wtot[w];
  }
  if (tid == 0) *cluster.map_shared_rank(&s_peer_total, half ^ 1) = total;
  cluster.sync();
  if (half == 1 && tid == 0) {  // the CTA with the lower lags reserves for both
    const unsigned int both = total + s_peer_total;
    unsigned int base = 0;
    if (both > 0) {
      base = atomicAdd(&ctr->cand_used, both);
      if (base + both > pool_cap) {
        atomicOr(&ctr->status, (unsigned int)ST_CAND_OVERFLOW);
        base = 0xffffffffu;
      }
    }
    cand_ref[spi] = make_uint2(base, both);
    s_base = base;
    *cluster.map_shared_rank(&s_base, 0) = base == 0xffffffffu ? base : base + total;
  }
  cluster.sync();
  const unsigned int base = s_base;
  if (base != 0xffffffffu && mine > 0) {
    unsigned int o = base + wbase + (incl - mine);
    uint32_t m = mask[tid];
    while (m) {
      const int bit = __ffs(m) - 1;
      m &= m - 1;
      cand_pool[o++] = (uint16_t)(lag0 + tid * 32 + bit);
    }
  }
}

template <int LOG2N, int NT>
__global__ void __launch_bounds__(NT, 1)
    combine_findtop_kernel(const uint32_t *__restrict__ pair_list, int n_pairs, const uint32_t *__restrict__ direct_list,
                           const SpDesc *__restrict__ sps, Slots ws, const float2 *__restrict__ scratch, double cutoff,
                           double cutoff_fast, uint16_t *__restrict__ cand_pool, unsigned int pool_cap,
                           uint2 *__restrict__ cand_ref, BatchCounters *ctr, float *__restrict__ xc_tap) {
  constexpr int N = 1 << LOG2N, H = N / 2, NWARP = NT / 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float *xs = reinterpret_cast<float *>(smem_raw);                                       // N lags
  uint32_t *mask = reinterpret_cast<uint32_t *>(smem_raw + (size_t)N * sizeof(float));  // NW words
  __shared__ unsigned int s_wtot[NWARP];
  __shared__ unsigned int s_base;
  const int tid = threadIdx.x, sj = blockIdx.x;
  int job, comp, spi;
  if (sj < 2 * n_pairs) {
    job = sj >> 1;
    comp = sj & 1;
    spi = (int)pair_list[job] + comp;
  } else {
    job = n_pairs + (sj - 2 * n_pairs);
    comp = 0;
    spi = (int)direct_list[sj - 2 * n_pairs];
  }
  const SpDesc sp = sps[spi];
  const int qlen = ws.meta[sp.q_slot].len;
  // xc[i] = x[(i + off) mod N] * scale: rescale + half rotation (CrossCorr.cc:493-505); the derived reverse
  // strand is additionally rotated by qlen - 1 lags and the pair product carries a factor 2
  const float scale = (sj < 2 * n_pairs ? 0.5f : 1.0f) / (float)N;
  const int off = comp ? ((H - (qlen - 1)) & (N - 1)) : H;
  const float2 *e = scratch + (size_t)job * 2 * H, *o = e + H;
  const float2 *__restrict__ wn = ws.wn;
  for (int n = tid; n < H; n += NT) {
    const float2 ev = __ldg(e + n), ov = __ldg(o + n);
    const float2 t = cmulc(ov, __ldg(wn + n));
    const float2 lo = cadd(ev, t), hi = csub(ev, t);
    xs[(n - off) & (N - 1)] = (comp ? lo.y : lo.x) * scale;
    xs[(n + H - off) & (N - 1)] = (comp ? hi.y : hi.x) * scale;
  }
  __syncthreads();
  const double co = (sp.flags & SP_FAST) ? cutoff_fast : cutoff;
  auto xc_at = [&](int i) -> float { return xs[i]; };
  findtop_impl<LOG2N, NT>(xc_at, co, mask, s_wtot, &s_base, spi, cand_pool, pool_cap, cand_ref, ctr, xc_tap);
}

// =================================================================================================
// Match probability (AlignProbability.cc:11-40, 62-127 / ProbTable.cc:58-74, 105-140), FP64, no
// FMA contraction so that it matches the reference's x86-64 arithmetic operation by operation
// (erf/exp come from the CUDA math library: <= 2 ulp from glibc's).
// Returns true when the segment is kept ("if (prob < m_minProb) continue", Slave.cc:186,194).
// =================================================================================================
__device__ __forceinline__ bool score_counts(double matches, double gcT, double gcQ, int len,
                                             const ScoreParams &prm, double &prob, double &ident) {
  const double dl = (double)len;
  ident = __ddiv_rn(matches, dl);
  const double gc_target = __ddiv_rn(gcT, dl);
  const double at_target = __dsub_rn(1., gc_target);
  double r = __dmul_rn(gcQ, gc_target);
  r = __dadd_rn(r, __dmul_rn(__dsub_rn(dl, gcQ), at_target));
  const double p_match = __ddiv_rn(__ddiv_rn(r, dl), 2.);
  if (prm.use_table) {
    const int index = (int)__dmul_rn(p_match, 511.0);
    if (index < 1 || index > 511) {
      prob = 0.;  // reference reads out of bounds for index 0: defined as "reject"
    } else {
      const int l = len >= 2048 ? 2047 : len;
      prob = (ident >= __ldg(prm.table + (size_t)index * 2048 + l)) ? prm.table_value : 0.;
    }
  } else {
    const double s = __dsqrt_rn(__dmul_rn(__dmul_rn(p_match, __dsub_rn(1., p_match)), dl));
    const double m = __dmul_rn(p_match, dl);
    const double x = __dmul_rn(dl, ident);
    const double z = __ddiv_rn(__ddiv_rn(__dsub_rn(m, x), s), 1.414213562);
    // exact early reject: for z > z_cut the host has verified cdf(z)*T >= 2*(-ln min_prob), i.e.
    // prob <= min_prob^2 < min_prob whatever the last-ulp behaviour of erf/exp (NaN falls through)
    if (z > prm.z_cut) {
      prob = 0.;
      return false;
    }
    const double cdf = __dmul_rn(0.5, __dadd_rn(1., erf(z)));
    const double expect = __dmul_rn(cdf, prm.target_total);
    prob = exp(-expect);
  }
  if (len < prm.min_len) return false;
  return !(prob < prm.min_prob);
}

__device__ __forceinline__ void emit_result(const SpDesc &sp, int start_t, int shift, int len, double prob,
                                            double ident, ResultRec *res_pool, unsigned int res_cap,
                                            BatchCounters *ctr) {
  const unsigned int slot = atomicAdd(&ctr->res_used, 1u);
  if (slot >= res_cap) {
    atomicOr(&ctr->status, (unsigned int)ST_RES_OVERFLOW);
    return;
  }
  ResultRec r;
  r.pair = sp.pair;
  r.strand = sp.flags & SP_REVERSE;
  r.start_t = start_t;
  r.start_q = start_t + shift;
  r.len = len;
  r.shift = shift;
  r.prob = prob;
  r.ident = ident;
  res_pool[slot] = r;
}

__device__ __forceinline__ void tap_segment(int spi, int start_t, int shift, int len, SegRec *seg_tap,
                                            unsigned int seg_tap_cap, BatchCounters *ctr) {
  if (seg_tap == nullptr) return;
  const unsigned int slot = atomicAdd(&ctr->seg_tap_used, 1u);
  if (slot >= seg_tap_cap) {
    atomicOr(&ctr->status, (unsigned int)ST_TAP_OVERFLOW);
    return;
  }
  SegRec s;
  s.sp = spi;
  s.start_t = start_t;
  s.shift = shift;
  s.len = len;
  seg_tap[slot] = s;
}

__device__ __forceinline__ uint32_t range_mask(int lo, int hi) {  // bits [lo, hi) of a 32-bit word
  if (hi <= 0 || lo >= 32 || lo >= hi) return 0u;
  const uint32_t upper = hi >= 32 ? 0xffffffffu : ((1u << hi) - 1u);
  const uint32_t lower = lo <= 0 ? 0u : ((1u << lo) - 1u);
  return upper & ~lower;
}

#define SX_SEGQ_CAP 3072
#include "sx_scan.cuh"

// =================================================================================================
// K3 (generic path, any IUPAC / unknown byte present): the reference's loop verbatim in spirit --
// running integer score sum over a 46-wide window with the 128x128 score table, one diagonal per
// thread, then FP64 scoring with sequential summation in the reference's order.
// =================================================================================================
__device__ __forceinline__ bool score_generic(const uint8_t *tb, const uint8_t *qb, const uint16_t *fcode,
                                              int start_t, int shift, int len, const ScoreParams &prm,
                                              double &prob, double &ident) {
  double matches = 0., gct = 0., gcq = 0.;
  for (int i = 0; i < len; i++) {
    const uint32_t fa = fcode[tb[start_t + i]], fb = fcode[qb[start_t + shift + i]];
    // DNA_Equal (DNAVector.cc:390-403): a + c + g + t, left to right
    double dot = __dmul_rn(frac_of(fa, 0), frac_of(fb, 0));
    dot = __dadd_rn(dot, __dmul_rn(frac_of(fa, 1), frac_of(fb, 1)));
    dot = __dadd_rn(dot, __dmul_rn(frac_of(fa, 2), frac_of(fb, 2)));
    dot = __dadd_rn(dot, __dmul_rn(frac_of(fa, 3), frac_of(fb, 3)));
    matches = __dadd_rn(matches, dot);
    gct = __dadd_rn(gct, __dadd_rn(frac_of(fa, 1), frac_of(fa, 2)));
    gcq = __dadd_rn(gcq, __dadd_rn(frac_of(fb, 1), frac_of(fb, 2)));
  }
  return score_counts(matches, gct, gcq, len, prm, prob, ident);
}

template <int LOG2N, int NT>
__global__ void __launch_bounds__(NT)
    scan_score_generic_kernel(const SpDesc *__restrict__ sps, int nsp, Slots ws, const uint16_t *__restrict__ cand_pool,
                              const uint2 *__restrict__ cand_ref, ScoreParams prm,
                              ResultRec *__restrict__ res_pool, unsigned int res_cap,
                              SegRec *__restrict__ seg_tap, unsigned int seg_tap_cap, BatchCounters *ctr) {
  constexpr int N = 1 << LOG2N, H = N / 2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint8_t *s_t = smem_raw;              // N
  uint8_t *s_q = s_t + N;               // N
  uint8_t *s_score = s_q + N;           // 128*128
  uint16_t *s_fcode = reinterpret_cast<uint16_t *>(s_score + 128 * 128);  // 128
  uint2 *s_segq = reinterpret_cast<uint2 *>(s_fcode + 128);               // SX_SEGQ_CAP
  __shared__ unsigned int s_nseg;

  const int tid = threadIdx.x;
  if (ctr->n_generic == 0) return;  // the bit-parallel kernel saw nothing for this one (the usual case)
  unsigned long long my_segments = 0, my_positions = 0;
  // a fixed grid strides over the strand-pairs
  for (int spi = blockIdx.x; spi < nsp; spi += gridDim.x) {
  __syncthreads();  // the previous strand-pair's shared-memory contents are dead
  const SpDesc sp = sps[spi];
  const uint2 cref = cand_ref[spi];
  const int ncand = (int)cref.y;
  if (ncand == 0 || cref.x == 0xffffffffu) continue;
  const SlotMeta tm = ws.meta[sp.t_slot], qm = ws.meta[sp.q_slot];
  if (!((tm.flags | qm.flags) & SLOT_NONACGT)) continue;  // handled by the bit-parallel kernel
  const int tlen = tm.len, qlen = qm.len;
  {
    // bases of both chunks: from the oriented bytes when the chunk holds IUPAC / unknown letters, rebuilt from
    // the 2-bit planes when it is pure A/C/G/T (the encoder stores bytes only for the former)
    constexpr int NWp = N / 32;
    auto load_bases = [&](uint8_t *dst, int slot, const SlotMeta &m) {
      if (m.flags & SLOT_NONACGT) {
        const uint4 *src16 = reinterpret_cast<const uint4 *>(ws.bytes + (size_t)slot * N);
        for (int i = tid; i < (m.len + 15) / 16; i += NT) reinterpret_cast<uint4 *>(dst)[i] = src16[i];
      } else {
        const uint32_t *pl = ws.planes + (size_t)slot * 2 * NWp;
        for (int i = tid; i < (m.len + 3) / 4; i += NT) {  // 4 bases per thread
          const uint32_t lo = (pl[i >> 3] >> ((i & 7) * 4)) & 15u, hi = (pl[NWp + (i >> 3)] >> ((i & 7) * 4)) & 15u;
          uint32_t w = 0;
#pragma unroll
          for (int b = 0; b < 4; b++) {
            const uint32_t code = ((lo >> b) & 1u) | (((hi >> b) & 1u) << 1);
            w |= ((0x54474341u >> (8 * code)) & 0xffu) << (8 * b);  // "ACGT"[code]
          }
          reinterpret_cast<uint32_t *>(dst)[i] = w;
        }
      }
    };
    load_bases(s_t, sp.t_slot, tm);
    load_bases(s_q, sp.q_slot, qm);
    const uint4 *sc = reinterpret_cast<const uint4 *>(g_score);
    for (int i = tid; i < 128 * 128 / 16; i += NT) reinterpret_cast<uint4 *>(s_score)[i] = sc[i];
    if (tid < 128) s_fcode[tid] = c_fcode[tid];
    if (tid == 0) s_nseg = 0;
  }
  __syncthreads();

  for (int c0 = 0; c0 < ncand; c0 += NT) {
    const int c = c0 + tid;
    if (c < ncand) {
      const int shift = (int)cand_pool[cref.x + c] - H;
      const int i0 = shift < 0 ? -shift : 0;
      int i_end = qlen - shift;
      if (tlen - 1 < i_end) i_end = tlen - 1;
      int sum = 0, open = -1;
      if (i_end > i0) my_positions += (unsigned long long)(i_end - i0);
      auto close_segment = [&](int i) {
        const int seg_len = i - open;
        my_segments++;
        tap_segment(spi, open, shift, seg_len, seg_tap, seg_tap_cap, ctr);
        const unsigned int slot = atomicAdd(&s_nseg, 1u);
        if (slot < SX_SEGQ_CAP) {
          s_segq[slot] = make_uint2((uint32_t)open | ((uint32_t)seg_len << 16), (uint32_t)shift);
        } else {
          double prob, ident;
          if (score_generic(s_t, s_q, s_fcode, open, shift, seg_len, prm, prob, ident))
            emit_result(sp, open, shift, seg_len, prob, ident, res_pool, res_cap, ctr);
        }
        open = -1;
      };
      for (int i = i0; i < i_end; i++) {
        sum += s_score[s_t[i] * 128 + s_q[i + shift]];
        if (i - i0 > 45) {
          sum -= s_score[s_t[i - 46] * 128 + s_q[i - 46 + shift]];
          if (sum > 1889) {  // (int)(45 * 0.42 * 100) evaluates to 1889 in IEEE double (CrossCorr.cc:675)
            if (open < 0) open = i - 45;
          } else if (open >= 0) {
            close_segment(i);
          }
        }
      }
      if (open >= 0 && i_end > i0) close_segment(i_end);
    }
    __syncthreads();
    const int nq = min((int)s_nseg, SX_SEGQ_CAP);
    for (int s = tid; s < nq; s += NT) {
      const uint2 q = s_segq[s];
      const int start_t = (int)(q.x & 0xffffu), seg_len = (int)(q.x >> 16), shift = (int)q.y;
      double prob, ident;
      if (score_generic(s_t, s_q, s_fcode, start_t, shift, seg_len, prm, prob, ident))
        emit_result(sp, start_t, shift, seg_len, prob, ident, res_pool, res_cap, ctr);
    }
    __syncthreads();
    if (tid == 0) s_nseg = 0;
    __syncthreads();
  }
  }  // strand-pairs
  for (int o = 16; o > 0; o >>= 1) my_segments += __shfl_xor_sync(0xffffffffu, my_segments, o);
  if ((tid & 31) == 0 && my_segments) atomicAdd(&ctr->n_segments, my_segments);
  for (int o = 16; o > 0; o >>= 1) my_positions += __shfl_xor_sync(0xffffffffu, my_positions, o);
  if ((tid & 31) == 0 && my_positions) atomicAdd(&ctr->n_positions, my_positions);
}

// =================================================================================================
// Launchers
// =================================================================================================
// Experiment knob (tools/dual_context_probe.py): SX_FFT_SMEM_KB=<n> requests at least n KiB of dynamic shared memory for
// the transform kernels, i.e. caps their CTAs per SM so that CTAs of another context's scan kernel fit beside them.
static size_t fft_smem_floor() {
  static const size_t v = [] {
    const char *e = getenv("SX_FFT_SMEM_KB");
    return e ? (size_t)atoi(e) * 1024 : (size_t)0;
  }();
  return v;
}

template <int LOG2N>
struct Cfg {
  static constexpr int NT = (LOG2N >= 14) ? 512 : 256;
  // N = 32768: the complex buffer (256 KiB) exceeds one CTA's shared memory; every transform is handled
  // by two CTAs, one per half of the radix-2 split, with the combine step going through HBM/L2
  static constexpr bool SPLIT = LOG2N >= 15;
};

template <int LOG2N>
static cudaError_t encode_launch(const SigDesc *sigs, int nsig, Slots ws, float *tap, PrepBuf prep, const uint32_t *enc_list,
                                 int n_enc, cudaStream_t st) {
  constexpr int N = 1 << LOG2N, NT = Cfg<LOG2N>::NT;
  if constexpr (Cfg<LOG2N>::SPLIT) {  // one CTA per half signal
    const size_t smem = (size_t)(N / 2) * 8 + N + 512 * 4 + 128 * 2 + 256;
    auto k = encode_fft_half_kernel<LOG2N, NT>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (tap != nullptr) prep.flag = nullptr;  // the signal tap shows the in-kernel route, stage by stage
    if (prep.flag != nullptr) {
      encode_prep_kernel<LOG2N><<<(nsig + 7) / 8, 256, 0, st>>>(sigs, nsig, ws, prep);
      if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    k<<<dim3(nsig, 2), NT, smem, st>>>(sigs, ws, tap, prep);
    return cudaGetLastError();
  } else {
  const size_t smem = std::max((size_t)N * 8 + N + 512 * 4 + 128 * 2 + 256, fft_smem_floor());
  auto k = encode_fft_kernel<LOG2N, NT>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  if (tap != nullptr) prep.flag = nullptr;  // the signal tap shows the in-kernel route, stage by stage
  if (prep.flag != nullptr) {
    encode_prep_kernel<LOG2N><<<(nsig + 7) / 8, 256, 0, st>>>(sigs, nsig, ws, prep);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
  }
  if (enc_list != nullptr && n_enc == 0) return cudaSuccess;
  k<<<enc_list != nullptr ? n_enc : nsig, NT, smem, st>>>(sigs, ws, tap, prep, enc_list, nullptr);
  return cudaGetLastError();
  }
}

template <int LOG2N>
static cudaError_t xcorr_launch(const SpDesc *sps, const uint32_t *pair_list, int n_pairs, const uint32_t *direct_list,
                                int n_direct, Slots ws, double cutoff, double cutoff_fast, uint16_t *cand_pool,
                                unsigned int pool_cap, uint2 *cand_ref, BatchCounters *ctr, float *xc_tap,
                                float2 *scratch, bool split_three_kernels, cudaStream_t st) {
  constexpr int N = 1 << LOG2N, NT = Cfg<LOG2N>::NT;
  if constexpr (Cfg<LOG2N>::SPLIT) {
    if (n_pairs == 0 && !split_three_kernels) {  // one cluster of two CTAs per strand-pair, combine through DSMEM
      auto kc = xcorr_cluster_kernel<LOG2N, NT>;
      const size_t smemc = (size_t)(N / 2) * 8 + (size_t)TwTables<LOG2N>::TOTAL * 8 + (size_t)(N / 64) * 4;
      cudaError_t e = cudaFuncSetAttribute(kc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemc);
      if (e != cudaSuccess) return e;
      kc<<<2 * n_direct, NT, smemc, st>>>(direct_list, sps, ws, cutoff, cutoff_fast, cand_pool, pool_cap, cand_ref, ctr, xc_tap);
      return cudaGetLastError();
    }
    if (scratch == nullptr) return cudaErrorInvalidValue;
    const int jobs = n_pairs + n_direct;
    auto k1 = xcorr_half_kernel<LOG2N, NT>;
    const size_t smem1 = (size_t)(N / 2) * 8 + (size_t)TwTables<LOG2N>::TOTAL * 8;
    cudaError_t e = cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1);
    if (e != cudaSuccess) return e;
    k1<<<dim3(jobs, 2), NT, smem1, st>>>(pair_list, n_pairs, direct_list, sps, ws, scratch);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    auto k2 = combine_findtop_kernel<LOG2N, NT>;
    const size_t smem2 = (size_t)N * 4 + (N / 32) * 4;
    e = cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
    if (e != cudaSuccess) return e;
    k2<<<2 * n_pairs + n_direct, NT, smem2, st>>>(pair_list, n_pairs, direct_list, sps, ws, scratch, cutoff, cutoff_fast,
                                                 cand_pool, pool_cap, cand_ref, ctr, xc_tap);
    return cudaGetLastError();
  } else {
  const size_t smem = std::max((size_t)N * 8 + (N / 32) * 4 + (size_t)TwTables<LOG2N>::TOTAL * 8, fft_smem_floor());
  if (n_pairs > 0) {
    auto k = xcorr_pair_kernel<LOG2N, NT>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k<<<n_pairs, NT, smem, st>>>(pair_list, sps, ws, cutoff, cutoff_fast, cand_pool, pool_cap, cand_ref, ctr, xc_tap, nullptr);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
  }
  if (n_direct > 0) {
    auto k = xcorr_findtop_kernel<LOG2N, NT>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k<<<n_direct, NT, smem, st>>>(direct_list, sps, ws, cutoff, cutoff_fast, cand_pool, pool_cap, cand_ref, ctr, xc_tap);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
  }
  return cudaSuccess;
  }
}

template <int LOG2N>
static cudaError_t fused_launch(const FusedJob *jobs, int njobs, const SigDesc *sigs, const SpDesc *sps, Slots ws, PrepBuf prep,
                                float2 *scratch, int grid, double cutoff, double cutoff_fast, uint16_t *cand_pool,
                                unsigned int pool_cap, uint2 *cand_ref, BatchCounters *ctr, FusedFail fail, cudaStream_t st) {
  if constexpr (LOG2N > 13) {
    return cudaErrorInvalidValue;  // 3 x H complex points do not fit two CTAs per SM: the separate kernels do these sizes
  } else {
    constexpr int N = 1 << LOG2N, NT = Cfg<LOG2N>::NT;
    auto k = pair_fused_kernel<LOG2N>;
    const size_t smem = FusedCfg<LOG2N>::SMEM;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k<<<grid, SX_FUSED_NT, smem, st>>>(jobs, njobs, sigs, sps, ws, prep, scratch, cutoff, cutoff_fast, cand_pool, pool_cap,
                                       cand_ref, ctr, fail);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    // fall-backs (chunks with a letter other than A/C/G/T), normally none: fixed grids that leave at once
    const int fb_grid = 148;
    auto k1 = encode_fft_kernel<LOG2N, NT>;
    const size_t smem1 = std::max((size_t)N * 8 + N + 512 * 4 + 128 * 2 + 256, fft_smem_floor());
    if ((e = cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1)) != cudaSuccess) return e;
    k1<<<fb_grid, NT, smem1, st>>>(sigs, ws, nullptr, prep, fail.sigs, fail.n_sigs);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    auto k2 = xcorr_pair_kernel<LOG2N, NT>;
    const size_t smem2 = std::max((size_t)N * 8 + (N / 32) * 4 + (size_t)TwTables<LOG2N>::TOTAL * 8, fft_smem_floor());
    if ((e = cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2)) != cudaSuccess) return e;
    k2<<<fb_grid, NT, smem2, st>>>(fail.pairs, sps, ws, cutoff, cutoff_fast, cand_pool, pool_cap, cand_ref, ctr, nullptr,
                                   fail.n_pairs);
    return cudaGetLastError();
  }
}

template <int LOG2N>
static cudaError_t scan_launch(const SpDesc *sps, int nsp, Slots ws, const uint16_t *cand_pool,
                               const uint2 *cand_ref, ScoreParams prm, ResultRec *res_pool, unsigned int res_cap,
                               SegRec *seg_tap, unsigned int seg_tap_cap,
                               BatchCounters *ctr, cudaStream_t st) {
  constexpr int N = 1 << LOG2N, NT = 256;
  constexpr size_t scan_smem = ScanCfg<LOG2N>::SMEM;
  constexpr int SPC = ScanCfg<LOG2N>::SPC;
  cudaError_t e = cudaSuccess;
  if (scan_smem > 48 * 1024) {
    e = cudaFuncSetAttribute(scan_score_kernel<LOG2N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scan_smem);
    if (e != cudaSuccess) return e;
  }
  scan_score_kernel<LOG2N><<<(nsp + SPC - 1) / SPC, SX_SCAN_NT, scan_smem, st>>>(
      sps, nsp, ws, cand_pool, cand_ref, prm, res_pool, res_cap, seg_tap, seg_tap_cap, ctr);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const size_t smem = (size_t)2 * N + 128 * 128 + 128 * 2 + (size_t)SX_SEGQ_CAP * 8;
  auto k = scan_score_generic_kernel<LOG2N, NT>;
  e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k<<<nsp < 592 ? nsp : 592, NT, smem, st>>>(sps, nsp, ws, cand_pool, cand_ref, prm, res_pool, res_cap, seg_tap, seg_tap_cap,
                                           ctr);  // 148 SMs x 4
  return cudaGetLastError();
}

#define SX_DISPATCH(log2n, CALL)       \
  switch (log2n) {                     \
    case 11: return CALL(11);          \
    case 12: return CALL(12);          \
    case 13: return CALL(13);          \
    case 14: return CALL(14);          \
    case 15: return CALL(15);          \
    default: return cudaErrorInvalidValue; \
  }

cudaError_t launch_encode_fft(int log2n, const SigDesc *sigs, int nsig, Slots ws, float *tap5n, PrepBuf prep,
                              const uint32_t *enc_list, int n_enc, cudaStream_t stream) {
  if (nsig <= 0) return cudaSuccess;
#define CALL(L) encode_launch<L>(sigs, nsig, ws, tap5n, prep, enc_list, n_enc, stream)
  SX_DISPATCH(log2n, CALL)
#undef CALL
}

cudaError_t launch_xcorr_findtop(int log2n, const SpDesc *sps, const uint32_t *pair_list, int n_pairs,
                                 const uint32_t *direct_list, int n_direct, Slots ws, double cutoff,
                                 double cutoff_fast, uint16_t *cand_pool, unsigned int pool_cap,
                                 uint2 *cand_ref, BatchCounters *ctr, float *xc_tap, float2 *scratch,
                                 bool split_three_kernels, cudaStream_t stream) {
  if (n_pairs <= 0 && n_direct <= 0) return cudaSuccess;
#define CALL(L) xcorr_launch<L>(sps, pair_list, n_pairs, direct_list, n_direct, ws, cutoff, cutoff_fast, cand_pool, pool_cap, cand_ref, ctr, xc_tap, scratch, split_three_kernels, stream)
  SX_DISPATCH(log2n, CALL)
#undef CALL
}

bool log2n_fusable(int log2n) { return log2n >= 11 && log2n <= 13; }
cudaError_t launch_pair_fused(int log2n, const FusedJob *jobs, int njobs, const SigDesc *sigs, const SpDesc *sps, Slots ws,
                              PrepBuf prep, float2 *scratch, int grid, double cutoff, double cutoff_fast, uint16_t *cand_pool,
                              unsigned int pool_cap, uint2 *cand_ref, BatchCounters *ctr, FusedFail fail,
                              cudaStream_t stream) {
  if (njobs <= 0) return cudaSuccess;
#define CALL(L) fused_launch<L>(jobs, njobs, sigs, sps, ws, prep, scratch, grid, cutoff, cutoff_fast, cand_pool, pool_cap, cand_ref, ctr, fail, stream)
  SX_DISPATCH(log2n, CALL)
#undef CALL
}

template <int LOG2N>
static cudaError_t findtop_external_launch(const float *xc, int spi, double co, uint16_t *cand_pool, unsigned int pool_cap,
                                           uint2 *cand_ref, BatchCounters *ctr, cudaStream_t st) {
  findtop_external_kernel<LOG2N, 256><<<1, 256, 0, st>>>(xc, spi, co, cand_pool, pool_cap, cand_ref, ctr);
  return cudaGetLastError();
}
cudaError_t launch_findtop_external(int log2n, const float *xc, int spi, double cutoff, uint16_t *cand_pool,
                                    unsigned int pool_cap, uint2 *cand_ref, BatchCounters *ctr, cudaStream_t stream) {
#define CALL(L) findtop_external_launch<L>(xc, spi, cutoff, cand_pool, pool_cap, cand_ref, ctr, stream)
  SX_DISPATCH(log2n, CALL)
#undef CALL
}

cudaError_t launch_scan_score(int log2n, const SpDesc *sps, int nsp, Slots ws, const uint16_t *cand_pool,
                              const uint2 *cand_ref, ScoreParams prm, ResultRec *res_pool,
                              unsigned int res_cap, SegRec *seg_tap, unsigned int seg_tap_cap, BatchCounters *ctr,
                              cudaStream_t stream) {
  if (nsp <= 0) return cudaSuccess;
#define CALL(L) scan_launch<L>(sps, nsp, ws, cand_pool, cand_ref, prm, res_pool, res_cap, seg_tap, seg_tap_cap, ctr, stream)
  SX_DISPATCH(log2n, CALL)
#undef CALL
}

}  // namespace sx
