// K3 (fast path, both chunks pure A/C/G/T): bit-parallel diagonal scan + scoring.
// Included by sx_kernels.cu after score_counts / emit_result / tap_segment are defined.
//
// Replaces SeqAnalyzer::MatchUp / DoOne (analysis/CrossCorr.cc:583-605, 667-724).  For A/C/G/T the
// reference's integer scores are 100 (equal) / 0 (different), so "window sum > 1889" (46-wide
// window; (int)(45*0.42*100) is 1889 in IEEE double) is "at least 19 of the last 46 positions match".
//
// One CTA per strand-pair, the 2-bit base planes of both chunks in shared memory; each WARP takes
// groups of 32 candidate lags.  Everything is done in DIAGONAL coordinates k = i - i0, 32 positions
// (one word of match bits) at a time.  Three phases per group:
//
//  A. filter (one diagonal per lane, warp-uniform loop): match word m = XOR of the planes, then an
//     exact NECESSARY condition for any window ending in this word to reach 19: for each quarter of
//     the word, the popcount of the 53 positions that cover all windows ending in that quarter.  On
//     random DNA ~88 % of the words fail it and need nothing more.  Survivors are recorded as bits.
//  B. units: maximal runs of surviving words.  The word before and after a unit has no passing
//     position, so units are independent.  They are spread over the lanes; a lane evaluates a unit
//     with the exact bit-sliced 46-window count (doubling: windows 2,4,8,16,32, then 32+8+4+2; ">= 19"
//     is three logic ops on the six count planes), starting two words early with empty history --
//     every window ending inside the unit lies within those words, so the result is exact.  Run
//     starts/ends inside the unit give the segments, exactly as the reference's sequential loop.
//  C. scoring: segments are spread over the lanes; popcounts over the planes, an FP32 pre-reject, the
//     exact early reject and the reference's FP64 formula (score_counts).
// A full queue spills to a global list scored by score_spill_kernel -- nothing is dropped.
#pragma once

#define SX_SCAN_NT 128
#define SX_UNIT_CAP 512  // units queued per warp and group (more are evaluated in place)
#define SX_SEG_CAP 448   // segments queued per warp and group (more spill to the global list)

// the unit list of a warp is split in three by unit length: 1, 2, >= 3 words
__host__ __device__ constexpr int unit_cap(int bucket) { return bucket == 0 ? 288 : bucket == 1 ? 144 : 80; }
__host__ __device__ constexpr int unit_base(int bucket) { return bucket == 0 ? 0 : bucket == 1 ? 288 : 432; }

struct PlanePtrs {
  const uint32_t *tlo, *thi, *qlo, *qhi;  // each readable up to word NW+1 (zero padded)
};

// 32 positions of both sequences starting at target bit tw*32+tsh / query bit qw*32+qsh
__device__ __forceinline__ void diag_words(const PlanePtrs &P, int tw, int tsh, int qw, int qsh, uint32_t &tl,
                                           uint32_t &th, uint32_t &ql, uint32_t &qh) {
  tl = __funnelshift_r(P.tlo[tw], P.tlo[tw + 1], tsh);
  th = __funnelshift_r(P.thi[tw], P.thi[tw + 1], tsh);
  ql = __funnelshift_r(P.qlo[qw], P.qlo[qw + 1], qsh);
  qh = __funnelshift_r(P.qhi[qw], P.qhi[qw + 1], qsh);
}

// Counts over one segment [start_t, start_t+len) on lag `shift`, then the probability filter.
__device__ __forceinline__ bool score_fast(const PlanePtrs &P, int start_t, int shift, int len,
                                           const ScoreParams &prm, double &prob, double &ident) {
  int matches = 0, gct = 0, gcq = 0;
  const int qoff = start_t + shift;
  for (int k = 0; k < len; k += 32) {
    uint32_t tl, th, ql, qh;
    diag_words(P, (start_t + k) >> 5, (start_t + k) & 31, (qoff + k) >> 5, (qoff + k) & 31, tl, th, ql, qh);
    const int rem = len - k;
    const uint32_t vm = rem >= 32 ? 0xffffffffu : ((1u << rem) - 1u);
    matches += __popc(~((tl ^ ql) | (th ^ qh)) & vm);
    gct += __popc((tl ^ th) & vm);
    gcq += __popc((ql ^ qh) & vm);
  }
  // Cheap exact-by-margin reject in FP32 before the FP64 arithmetic: all inputs are small integers
  // (exact in float), the float estimate of the normalised deviation z is within ~1e-3 of the FP64
  // value, and z_cut already carries a factor-2 safety margin in probability, so "estimate > z_cut
  // + 0.25" can only drop segments whose probability is far below min_prob.  NaN/inf fall through.
  if (!prm.use_table && len >= prm.min_len) {
    const float fl = (float)len;
    const float num = (float)(gcq * gct + (len - gcq) * (len - gct));  // <= 2*len^2 < 2^30, exact enough
    const float p = num / (2.f * fl * fl);
    const float zf = (p * fl - (float)matches) * rsqrtf(p * (1.f - p) * fl) * 0.70710678f;
    if (zf > (float)prm.z_cut + 0.25f) return false;
  }
  return score_counts((double)matches, (double)gct, (double)gcq, len, prm, prob, ident);
}

// geometry of one candidate diagonal in a strand-pair
struct Diag {
  int shift, i0, L, nwords, tw0, tsh, qw0, qsh;
  uint32_t lastmask;
};

__device__ __forceinline__ Diag make_diag(int shift, int tlen, int qlen) {
  Diag d;
  d.shift = shift;
  d.i0 = shift < 0 ? -shift : 0;
  int i_end = qlen - shift;                // first i with j >= qlen
  if (tlen - 1 < i_end) i_end = tlen - 1;  // the last target base is never scored
  d.L = i_end - d.i0;
  if (d.L <= 46) d.L = 0;  // no window is ever evaluated (needs n > 45)
  d.nwords = (d.L + 31) >> 5;
  const int j0 = d.i0 + shift;
  d.tw0 = d.i0 >> 5;
  d.tsh = d.i0 & 31;
  d.qw0 = j0 >> 5;
  d.qsh = j0 & 31;
  d.lastmask = (d.L & 31) ? ((1u << (d.L & 31)) - 1u) : 0xffffffffu;
  return d;
}

// match bits of diagonal word kw (positions 32kw..32kw+31 of the diagonal), zero outside [0, L)
template <int NW>
__device__ __forceinline__ uint32_t match_bits(const PlanePtrs &P, const Diag &d, int kw) {
  uint32_t tl, th, ql, qh;
  diag_words(P, min(d.tw0 + kw, NW), d.tsh, min(d.qw0 + kw, NW), d.qsh, tl, th, ql, qh);
  const uint32_t vm = kw < d.nwords - 1 ? 0xffffffffu : (kw == d.nwords - 1 ? d.lastmask : 0u);
  return ~((tl ^ ql) | (th ^ qh)) & vm;
}

// one finished segment -> warp queue (shared-memory atomic), spill list when full
__device__ __forceinline__ void push_segment(int spi, int start_t, int shift, int seg_len, uint2 *wq, unsigned int *wqn,
                                             SegRec *spill, unsigned int spill_cap, SegRec *seg_tap,
                                             unsigned int seg_tap_cap, BatchCounters *ctr) {
  tap_segment(spi, start_t, shift, seg_len, seg_tap, seg_tap_cap, ctr);
  const unsigned int slot = atomicAdd(wqn, 1u);
  if (slot < SX_SEG_CAP) {
    wq[slot] = make_uint2((uint32_t)start_t | ((uint32_t)seg_len << 16), (uint32_t)shift);
  } else {
    const unsigned int gs = atomicAdd(&ctr->spill_used, 1u);
    if (gs < spill_cap) {
      SegRec r;
      r.sp = spi;
      r.start_t = start_t;
      r.shift = shift;
      r.len = seg_len;
      spill[gs] = r;
    } else {
      atomicOr(&ctr->status, (unsigned int)ST_SPILL_OVERFLOW);
    }
  }
}

// Exact evaluation of one unit = diagonal words [ks, ks+k) of diagonal d: bit-sliced 46-window count,
// run extraction, segments pushed.  Starts two words early with empty history (see file header).
// WARP-COLLECTIVE: all 32 lanes call it together (k = 0 for a lane without a unit); the step loop
// runs for the longest unit of the warp and re-converges every step.
template <int NW>
__device__ __forceinline__ int eval_unit(const PlanePtrs &P, const Diag &d, int ks, int k, int spi, uint2 *wq,
                                         unsigned int *wqn, SegRec *spill, unsigned int spill_cap, SegRec *seg_tap,
                                         unsigned int seg_tap_cap, BatchCounters *ctr) {
  uint32_t m_prev = 0, carry = 0;
  uint32_t s1p0 = 0, s1p1 = 0, s1q0 = 0, s1q1 = 0;
  uint32_t s2p0 = 0, s2p1 = 0, s2p2 = 0, s2q0 = 0, s2q1 = 0, s2q2 = 0;
  uint32_t s3p0 = 0, s3p1 = 0, s3p2 = 0, s3p3 = 0;
  uint32_t s4p0 = 0, s4p1 = 0, s4p2 = 0, s4p3 = 0, s4p4 = 0;
  int open = -1;  // start of the open run (diagonal coordinates), -1 = none
  int nseg = 0;
  const int k_end = ks + k;
  const int kw0 = max(ks - 2, 0);
  const int max_steps = __reduce_max_sync(0xffffffffu, k > 0 ? k_end - kw0 : 0);
#pragma unroll 1
  for (int step = 0; step < max_steps; step++) {  // warp-uniform
   const int kw = kw0 + step;
   if (kw < k_end && k > 0) {
    const uint32_t m = match_bits<NW>(P, d, kw);
    uint32_t cy;
    // window 2
    const uint32_t m1 = __funnelshift_l(m_prev, m, 1);
    const uint32_t s10 = m ^ m1, s11 = m & m1;
    // window 4 = s1 + (s1 delayed by 2)
    const uint32_t a0 = __funnelshift_l(s1p0, s10, 2), a1 = __funnelshift_l(s1p1, s11, 2);
    const uint32_t s20 = s10 ^ a0;
    cy = s10 & a0;
    const uint32_t s21 = s11 ^ a1 ^ cy;
    const uint32_t s22 = (s11 & a1) | (cy & (s11 ^ a1));
    // window 8 = s2 + (s2 delayed by 4)
    const uint32_t b0 = __funnelshift_l(s2p0, s20, 4), b1 = __funnelshift_l(s2p1, s21, 4),
                   b2 = __funnelshift_l(s2p2, s22, 4);
    const uint32_t s30 = s20 ^ b0;
    cy = s20 & b0;
    const uint32_t s31 = s21 ^ b1 ^ cy;
    cy = (s21 & b1) | (cy & (s21 ^ b1));
    const uint32_t s32 = s22 ^ b2 ^ cy;
    const uint32_t s33 = (s22 & b2) | (cy & (s22 ^ b2));
    // window 16 = s3 + (s3 delayed by 8)
    const uint32_t c0 = __funnelshift_l(s3p0, s30, 8), c1 = __funnelshift_l(s3p1, s31, 8),
                   c2 = __funnelshift_l(s3p2, s32, 8), c3 = __funnelshift_l(s3p3, s33, 8);
    const uint32_t s40 = s30 ^ c0;
    cy = s30 & c0;
    const uint32_t s41 = s31 ^ c1 ^ cy;
    cy = (s31 & c1) | (cy & (s31 ^ c1));
    const uint32_t s42 = s32 ^ c2 ^ cy;
    cy = (s32 & c2) | (cy & (s32 ^ c2));
    const uint32_t s43 = s33 ^ c3 ^ cy;
    const uint32_t s44 = (s33 & c3) | (cy & (s33 ^ c3));
    if (kw >= ks) {  // warm-up words only feed the history
      // window 32 = s4 + (s4 delayed by 16)
      const uint32_t d0 = __funnelshift_l(s4p0, s40, 16), d1 = __funnelshift_l(s4p1, s41, 16),
                     d2 = __funnelshift_l(s4p2, s42, 16), d3 = __funnelshift_l(s4p3, s43, 16),
                     d4 = __funnelshift_l(s4p4, s44, 16);
      const uint32_t s50 = s40 ^ d0;
      cy = s40 & d0;
      const uint32_t s51 = s41 ^ d1 ^ cy;
      cy = (s41 & d1) | (cy & (s41 ^ d1));
      const uint32_t s52 = s42 ^ d2 ^ cy;
      cy = (s42 & d2) | (cy & (s42 ^ d2));
      const uint32_t s53 = s43 ^ d3 ^ cy;
      cy = (s43 & d3) | (cy & (s43 ^ d3));
      const uint32_t s54 = s44 ^ d4 ^ cy;
      const uint32_t s55 = (s44 & d4) | (cy & (s44 ^ d4));
      // tail of the 46-window: positions 32..39 = s3 one word back, 40..43 = s2 delayed 40, 44..45 = s1 delayed 44
      const uint32_t e20 = __funnelshift_l(s2q0, s2p0, 8), e21 = __funnelshift_l(s2q1, s2p1, 8),
                     e22 = __funnelshift_l(s2q2, s2p2, 8);
      const uint32_t e10 = __funnelshift_l(s1q0, s1p0, 12), e11 = __funnelshift_l(s1q1, s1p1, 12);
      const uint32_t u0 = e20 ^ e10;  // u = e2 + e1 <= 6
      cy = e20 & e10;
      const uint32_t u1 = e21 ^ e11 ^ cy;
      cy = (e21 & e11) | (cy & (e21 ^ e11));
      const uint32_t u2 = e22 ^ cy;
      const uint32_t v0 = s3p0 ^ u0;  // v = s3(prev word) + u <= 14
      cy = s3p0 & u0;
      const uint32_t v1 = s3p1 ^ u1 ^ cy;
      cy = (s3p1 & u1) | (cy & (s3p1 ^ u1));
      const uint32_t v2 = s3p2 ^ u2 ^ cy;
      cy = (s3p2 & u2) | (cy & (s3p2 ^ u2));
      const uint32_t v3 = s3p3 ^ cy;
      const uint32_t n0 = s50 ^ v0;  // count = s5 + v <= 46
      cy = s50 & v0;
      const uint32_t n1 = s51 ^ v1 ^ cy;
      cy = (s51 & v1) | (cy & (s51 ^ v1));
      const uint32_t n2 = s52 ^ v2 ^ cy;
      cy = (s52 & v2) | (cy & (s52 ^ v2));
      const uint32_t n3 = s53 ^ v3 ^ cy;
      cy = (s53 & v3) | (cy & (s53 ^ v3));
      const uint32_t n4 = s54 ^ cy;
      const uint32_t n5 = s55 ^ (s54 & cy);
      // count >= 19 (0b010011), only where a full window has been seen (k >= 46) and inside the diagonal
      const uint32_t vm = kw < d.nwords - 1 ? 0xffffffffu : (kw == d.nwords - 1 ? d.lastmask : 0u);
      uint32_t pass = (n5 | (n4 & (n3 | n2 | (n1 & n0)))) & vm;
      if (kw < 2) pass &= (kw == 0) ? 0u : 0xffffc000u;
      // run boundaries inside this word: rise = a run starts here, fall = first position after a run
      const uint32_t prevp = (pass << 1) | carry;
      const uint32_t rise = pass & ~prevp;
      uint32_t fall = ~pass & prevp;
      carry = pass >> 31;
      const int kb = kw * 32;
      while (fall) {
        const int f = __ffs(fall) - 1;
        fall &= fall - 1u;
        const uint32_t below = rise & ((1u << f) - 1u);
        // the run started at the closest rise below f, or in an earlier word (carried in `open`)
        const int start_k = below ? (kb + 31 - __clz(below) - 45) : open;  // lastStart = i - m_minLen
        push_segment(spi, d.i0 + start_k, d.shift, kb + f - start_k, wq, wqn, spill, spill_cap, seg_tap, seg_tap_cap,
                     ctr);
        nseg++;
      }
      open = carry ? (rise ? (kb + 31 - __clz(rise) - 45) : open) : -1;
    }
    m_prev = m;
    s1q0 = s1p0; s1q1 = s1p1; s1p0 = s10; s1p1 = s11;
    s2q0 = s2p0; s2q1 = s2p1; s2q2 = s2p2; s2p0 = s20; s2p1 = s21; s2p2 = s22;
    s3p0 = s30; s3p1 = s31; s3p2 = s32; s3p3 = s33;
    s4p0 = s40; s4p1 = s41; s4p2 = s42; s4p3 = s43; s4p4 = s44;
   }
   __syncwarp();
  }
  // The word after a unit has no passing position (or the diagonal ends): an open run closes at the
  // word boundary / at the stop position L.
  if (open >= 0) {
    const int close = min(k_end * 32, d.L);
    push_segment(spi, d.i0 + open, d.shift, close - open, wq, wqn, spill, spill_cap, seg_tap, seg_tap_cap, ctr);
    nseg++;
  }
  return nseg;
}

template <int LOG2N>
constexpr size_t scan_smem_bytes() {
  constexpr size_t N = (size_t)1 << LOG2N, NW = N / 32, PW = NW + 2, NWARP = SX_SCAN_NT / 32, NBW = (NW + 31) / 32;
  return NWARP * SX_SEG_CAP * 8 + 4 * PW * 4 + NWARP * NBW * 32 * 4 + NWARP * SX_UNIT_CAP * 4 + NWARP * 32 * 4;
}

template <int LOG2N>
__global__ void __launch_bounds__(SX_SCAN_NT)
    scan_score_kernel(const SpDesc *__restrict__ sps, Slots ws, const uint16_t *__restrict__ cand_pool,
                      const uint2 *__restrict__ cand_ref, ScoreParams prm, ResultRec *__restrict__ res_pool,
                      unsigned int res_cap, SegRec *__restrict__ seg_tap, unsigned int seg_tap_cap,
                      SegRec *__restrict__ spill, unsigned int spill_cap, BatchCounters *ctr) {
  constexpr int N = 1 << LOG2N, H = N / 2, NW = N / 32, PW = NW + 2, NWARP = SX_SCAN_NT / 32;
  constexpr int NBW = (NW + 31) / 32;  // 32-bit words of the per-diagonal "word survives the filter" bitset
  // dynamic shared memory (scan_smem_bytes<LOG2N>(): above the 48 KiB static limit for N = 32768)
  extern __shared__ __align__(16) unsigned char scan_smem[];
  uint2(*s_wq)[SX_SEG_CAP] = reinterpret_cast<uint2(*)[SX_SEG_CAP]>(scan_smem);
  uint32_t *s_tlo = reinterpret_cast<uint32_t *>(s_wq + NWARP), *s_thi = s_tlo + PW, *s_qlo = s_thi + PW,
           *s_qhi = s_qlo + PW;
  uint32_t(*s_need)[NBW][32] = reinterpret_cast<uint32_t(*)[NBW][32]>(s_qhi + PW);
  uint32_t(*s_unit)[SX_UNIT_CAP] = reinterpret_cast<uint32_t(*)[SX_UNIT_CAP]>(s_need + NWARP);
  int(*s_shift)[32] = reinterpret_cast<int(*)[32]>(s_unit + NWARP);
  __shared__ unsigned int s_nunit[NWARP][4], s_wqn[NWARP];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const SpDesc sp = sps[blockIdx.x];
  const uint2 cref = cand_ref[blockIdx.x];
  const int ncand = (int)cref.y;
  if (ncand == 0 || cref.x == 0xffffffffu) return;
  const SlotMeta tm = ws.meta[sp.t_slot], qm = ws.meta[sp.q_slot];
  if ((tm.flags | qm.flags) & SLOT_NONACGT) return;  // handled by the generic kernel
  const int tlen = tm.len, qlen = qm.len;
  {
    const uint32_t *tp = ws.planes + (size_t)sp.t_slot * 2 * NW;
    const uint32_t *qp = ws.planes + (size_t)sp.q_slot * 2 * NW;
    for (int i = tid; i < PW; i += SX_SCAN_NT) {
      const bool in = i < NW;
      s_tlo[i] = in ? tp[i] : 0u;
      s_thi[i] = in ? tp[NW + i] : 0u;
      s_qlo[i] = in ? qp[i] : 0u;
      s_qhi[i] = in ? qp[NW + i] : 0u;
    }
    if (tid < NWARP) {
      s_nunit[tid][0] = s_nunit[tid][1] = s_nunit[tid][2] = s_nunit[tid][3] = 0;
      s_wqn[tid] = 0;
    }
  }
  __syncthreads();
  PlanePtrs P;
  P.tlo = s_tlo;
  P.thi = s_thi;
  P.qlo = s_qlo;
  P.qhi = s_qhi;
  uint2 *wq = s_wq[warp];
  unsigned int *wqn = &s_wqn[warp];
  uint32_t *units = s_unit[warp];
  unsigned int *nunit = s_nunit[warp];
  const int spi = blockIdx.x;

  unsigned int my_segments = 0;
  unsigned long long my_positions = 0;
  for (int g0 = warp * 32; g0 < ncand; g0 += SX_SCAN_NT) {  // warp-uniform
    // Candidates are sorted by lag, so diagonal lengths rise and fall like a triangle.  Pairing the
    // k-th from the front with the k-th from the back keeps the 32 diagonals of a warp close in length.
    const int p = g0 + lane;
    const int c = (p & 1) ? (ncand - 1 - (p >> 1)) : (p >> 1);
    int shift = 0;
    Diag d = make_diag(0, 0, 0);
    if (p < ncand) {
      shift = (int)cand_pool[cref.x + c] - H;  // pos = idx - N/2 (CrossCorr.cc:600-602)
      d = make_diag(shift, tlen, qlen);
    }
    s_shift[warp][lane] = shift;
    my_positions += (unsigned long long)d.L;
    const int nw_max = __reduce_max_sync(0xffffffffu, d.nwords);

    // ---- A. filter -------------------------------------------------------------------------------
    // Quarter q of word kw (positions 32kw+8q .. +7): every window ending there lies inside the 53
    // positions [32kw+8q-45, 32kw+8q+7].  From the previous words only a few popcounts are needed:
    //   q0: m[kw-2] bits 19..31 + m[kw-1]            + m[kw] bits 0..7
    //   q1: m[kw-2] bits 27..31 + m[kw-1]            + m[kw] bits 0..15
    //   q2:                       m[kw-1] bits 3..31  + m[kw] bits 0..23
    //   q3:                       m[kw-1] bits 11..31 + m[kw]
    int pp_all = 0, pp_ge3 = 0, pp_ge11 = 0;  // of m[kw-1]
    int p2_ge19 = 0, p2_ge27 = 0;             // of m[kw-2]
    int q_ge19 = 0, q_ge27 = 0;               // of m[kw-1], become p2_* next step
    uint32_t acc = 0;
    uint32_t rtl = s_tlo[min(d.tw0, NW)], rth = s_thi[min(d.tw0, NW)], rql = s_qlo[min(d.qw0, NW)],
             rqh = s_qhi[min(d.qw0, NW)];
#pragma unroll 2
    for (int kw = 0; kw < nw_max; kw++) {  // warp-uniform trip count
      const int tn = min(d.tw0 + kw + 1, NW + 1), qn = min(d.qw0 + kw + 1, NW + 1);
      const uint32_t ntl = s_tlo[tn], nth = s_thi[tn], nql = s_qlo[qn], nqh = s_qhi[qn];
      const uint32_t tl = __funnelshift_r(rtl, ntl, d.tsh), th = __funnelshift_r(rth, nth, d.tsh);
      const uint32_t ql = __funnelshift_r(rql, nql, d.qsh), qh = __funnelshift_r(rqh, nqh, d.qsh);
      rtl = ntl; rth = nth; rql = nql; rqh = nqh;
      const uint32_t vm = kw < d.nwords - 1 ? 0xffffffffu : (kw == d.nwords - 1 ? d.lastmask : 0u);
      const uint32_t m = ~((tl ^ ql) | (th ^ qh)) & vm;
      const int c_all = __popc(m), c_lo8 = __popc(m & 0xffu), c_lo16 = __popc(m & 0xffffu),
                c_lo24 = __popc(m & 0xffffffu);
      const int u0 = p2_ge19 + pp_all + c_lo8;
      const int u1 = p2_ge27 + pp_all + c_lo16;
      const int u2 = pp_ge3 + c_lo24;
      const int u3 = pp_ge11 + c_all;
      const bool need = max(max(u0, u1), max(u2, u3)) >= 19;
      acc |= (need ? 1u : 0u) << (kw & 31);
      if ((kw & 31) == 31) {  // warp-uniform
        s_need[warp][kw >> 5][lane] = acc;
        acc = 0;
      }
      p2_ge19 = q_ge19;
      p2_ge27 = q_ge27;
      pp_all = c_all;
      pp_ge3 = __popc(m >> 3);
      pp_ge11 = __popc(m >> 11);
      q_ge19 = __popc(m >> 19);
      q_ge27 = __popc(m >> 27);
    }
    if (nw_max & 31) s_need[warp][nw_max >> 5][lane] = acc;
    __syncwarp();

    // ---- B/C in rounds: enumerate units until the list is full, evaluate them, score the segments --
    const int nbw = (nw_max + 31) >> 5;
    int e_w = 0, e_pos = 0, e_start = -1;  // enumeration state: bitset word, bit position, open run start
    bool e_done = (d.nwords == 0);
    do {
      // B1. units = maximal runs of surviving words, (lane:5 | first word:13 | words:14)
      while (!e_done) {
        // find the next run end from (e_w, e_pos, e_start) without committing the state yet
        int w = e_w, pos = e_pos, start = e_start, ks = -1, kk = 0;
        while (w < nbw) {
          const uint32_t bits = s_need[warp][w][lane];
          if (start < 0) {
            const uint32_t r = pos < 32 ? (bits >> pos) : 0u;
            if (!r) { w++; pos = 0; continue; }
            pos += __ffs(r) - 1;
            start = w * 32 + pos;
          }
          // length of the run of ones at pos (zeros shifted in from the top end it at bit 32)
          const uint32_t z = pos < 32 ? ~(bits >> pos) : 1u;
          const int run = (pos == 0 && bits == 0xffffffffu) ? 32 : __ffs(z) - 1;
          pos += run;
          if (pos < 32 || w == nbw - 1) {  // the run ends here (or the diagonal does)
            ks = start;
            kk = w * 32 + pos - start;
            start = -1;
            break;
          }
          w++;  // the run continues in the next stretch
          pos = 0;
        }
        if (ks < 0) {
          e_done = true;
          break;
        }
        // three lists by unit length (1, 2, >= 3 words) so that the 32 units evaluated together take
        // the same number of steps
        const int bucket = min(kk, 3) - 1;
        const unsigned int slot = atomicAdd(&nunit[bucket], 1u);
        if (slot >= (unsigned int)unit_cap(bucket)) break;  // list full: this unit is found again in the next round
        units[unit_base(bucket) + slot] = (uint32_t)lane | ((uint32_t)ks << 5) | ((uint32_t)kk << 18);
        e_w = w;
        e_pos = pos;
        e_start = start;
      }
      __syncwarp();

      // B2. evaluate the units, one lane each, 32 at a time
      for (int bucket = 0; bucket < 3; bucket++) {
        const int nu = min((int)nunit[bucket], unit_cap(bucket));
        const uint32_t *ul = units + unit_base(bucket);
        for (int u0 = 0; u0 < nu; u0 += 32) {  // warp-uniform
          const int u = u0 + lane;
          const uint32_t rec = u < nu ? ul[u] : 0u;
          const int owner = (int)(rec & 31u), ks = (int)((rec >> 5) & 0x1fffu), kk = (int)(rec >> 18);
          const Diag od = make_diag(s_shift[warp][owner], tlen, qlen);
          my_segments += eval_unit<NW>(P, od, ks, kk, spi, wq, wqn, spill, spill_cap, seg_tap, seg_tap_cap, ctr);
        }
      }
      __syncwarp();

      // C. score the segments found so far, spread over the lanes
      const int nq = min((int)*wqn, SX_SEG_CAP);
      for (int sidx = lane; sidx < nq; sidx += 32) {
        const uint2 q = wq[sidx];
        const int start_t = (int)(q.x & 0xffffu), seg_len = (int)(q.x >> 16), sh = (int)q.y;
        double prob, ident;
        if (score_fast(P, start_t, sh, seg_len, prm, prob, ident))
          emit_result(sp, start_t, sh, seg_len, prob, ident, res_pool, res_cap, ctr);
      }
      __syncwarp();
      if (lane == 0) {
        *wqn = 0;
        nunit[0] = nunit[1] = nunit[2] = 0;
      }
      __syncwarp();
    } while (__any_sync(0xffffffffu, !e_done));
  }
  my_segments = __reduce_add_sync(0xffffffffu, my_segments);
  if (lane == 0 && my_segments) atomicAdd(&ctr->n_segments, (unsigned long long)my_segments);
  for (int o = 16; o > 0; o >>= 1) my_positions += __shfl_xor_sync(0xffffffffu, my_positions, o);
  if (lane == 0 && my_positions) atomicAdd(&ctr->n_positions, my_positions);
}

// Segments that did not fit a warp queue: same scoring, planes read from global memory.
template <int LOG2N>
__global__ void __launch_bounds__(128)
    score_spill_kernel(const SpDesc *__restrict__ sps, Slots ws, const SegRec *__restrict__ spill, ScoreParams prm,
                       ResultRec *__restrict__ res_pool, unsigned int res_cap, unsigned int spill_cap,
                       BatchCounters *ctr) {
  constexpr int N = 1 << LOG2N, NW = N / 32;
  const unsigned int n = min(ctr->spill_used, spill_cap);
  for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const SegRec r = spill[i];
    const SpDesc sp = sps[r.sp];
    const uint32_t *tp = ws.planes + (size_t)sp.t_slot * 2 * NW;
    const uint32_t *qp = ws.planes + (size_t)sp.q_slot * 2 * NW;
    int matches = 0, gct = 0, gcq = 0;
    const int qoff = r.start_t + r.shift;
    for (int k = 0; k < r.len; k += 32) {
      const int tw = (r.start_t + k) >> 5, tsh = (r.start_t + k) & 31, qw = (qoff + k) >> 5, qsh = (qoff + k) & 31;
      auto ld = [&](const uint32_t *p, int w) -> uint32_t { return w < NW ? p[w] : 0u; };
      const uint32_t tl = __funnelshift_r(ld(tp, tw), ld(tp, tw + 1), tsh);
      const uint32_t th = __funnelshift_r(ld(tp + NW, tw), ld(tp + NW, tw + 1), tsh);
      const uint32_t ql = __funnelshift_r(ld(qp, qw), ld(qp, qw + 1), qsh);
      const uint32_t qh = __funnelshift_r(ld(qp + NW, qw), ld(qp + NW, qw + 1), qsh);
      const int rem = r.len - k;
      const uint32_t vm = rem >= 32 ? 0xffffffffu : ((1u << rem) - 1u);
      matches += __popc(~((tl ^ ql) | (th ^ qh)) & vm);
      gct += __popc((tl ^ th) & vm);
      gcq += __popc((ql ^ qh) & vm);
    }
    double prob, ident;
    if (score_counts((double)matches, (double)gct, (double)gcq, r.len, prm, prob, ident))
      emit_result(sp, r.start_t, r.shift, r.len, prob, ident, res_pool, res_cap, ctr);
  }
}
