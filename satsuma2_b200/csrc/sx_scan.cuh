// K3 (fast path, both chunks pure A/C/G/T): bit-parallel diagonal scan + scoring.
// Included by sx_kernels.cu after score_counts / emit_result / tap_segment are defined.
//
// Replaces SeqAnalyzer::MatchUp / DoOne (analysis/CrossCorr.cc:583-605, 667-724).  For A/C/G/T the
// reference's integer scores are 100 (equal) / 0 (different), so "window sum > 1889" (46-wide
// window; (int)(45*0.42*100) is 1889 in IEEE double) is "at least 19 of the last 46 positions match".
//
// Work mapping: one CTA per strand-pair, the base planes of both chunks in shared memory; each WARP
// independently takes groups of 32 consecutive candidate lags, one diagonal per lane.  A lane walks
// its diagonal 32 positions per step in DIAGONAL coordinates k = i - i0 (both planes funnel-shifted
// to bit 0), so every lane starts at word 0 and the loop trip count is warp-uniform (max over lanes):
// the warp stays converged, the only divergent code is the few instructions that record a run
// start/end.  Per step: match bits by XOR of the 2-bit planes; the 46-wide sliding count built
// bit-sliced by doubling (windows 2,4,8,16,32, then 32+8+4+2); ">= 19" is three logic ops on the six
// count planes; run boundaries are the set bits of pass ^ (pass << 1 | carry).
// Finished segments go to a per-warp shared-memory queue and are scored (FP64) by the same warp
// after the group, 32 segments at a time; a full queue spills to a global list scored by
// score_spill_kernel -- nothing is dropped.
#pragma once

#define SX_SCAN_NT 128
#define SX_LQ_CAP 40  // queued segments per lane (diagonal); more spill to a global list

struct PlanePtrs {
  const uint32_t *tlo, *thi, *qlo, *qhi;  // each readable up to word NW+1 (zero padded)
};

// 32 positions of both sequences starting at diagonal position 32*kw: t bit offset toff, q bit offset qoff
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

// Appends one finished segment to the calling lane's private queue (slot-major layout: entry r of
// lane l at [r*32 + l], so simultaneous pushes of different lanes never conflict); a full queue spills
// to the global list.  Only start and length are stored: the lane that scanned the diagonal scores it.
__device__ __forceinline__ void push_segment(int start_t, int shift, int seg_len, int lane, uint32_t *wq, int &cnt,
                                             SegRec *spill, unsigned int spill_cap, SegRec *seg_tap,
                                             unsigned int seg_tap_cap, BatchCounters *ctr) {
  tap_segment(blockIdx.x, start_t, shift, seg_len, seg_tap, seg_tap_cap, ctr);
  if (cnt < SX_LQ_CAP) {
    wq[cnt * 32 + lane] = (uint32_t)start_t | ((uint32_t)seg_len << 16);
  } else {
    const unsigned int gs = atomicAdd(&ctr->spill_used, 1u);
    if (gs < spill_cap) {
      SegRec r;
      r.sp = blockIdx.x;
      r.start_t = start_t;
      r.shift = shift;
      r.len = seg_len;
      spill[gs] = r;
    } else {
      atomicOr(&ctr->status, (unsigned int)ST_SPILL_OVERFLOW);
    }
  }
  cnt++;
}

template <int LOG2N>
__global__ void __launch_bounds__(SX_SCAN_NT)
    scan_score_kernel(const SpDesc *__restrict__ sps, Slots ws, const uint16_t *__restrict__ cand_pool,
                      const uint2 *__restrict__ cand_ref, ScoreParams prm, ResultRec *__restrict__ res_pool,
                      unsigned int res_cap, SegRec *__restrict__ seg_tap, unsigned int seg_tap_cap,
                      SegRec *__restrict__ spill, unsigned int spill_cap, BatchCounters *ctr) {
  constexpr int N = 1 << LOG2N, H = N / 2, NW = N / 32, PW = NW + 2, NWARP = SX_SCAN_NT / 32;
  __shared__ uint32_t s_tlo[PW], s_thi[PW], s_qlo[PW], s_qhi[PW];
  __shared__ uint32_t s_wq[NWARP][SX_LQ_CAP * 32];

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
  }
  __syncthreads();
  PlanePtrs P;
  P.tlo = s_tlo;
  P.thi = s_thi;
  P.qlo = s_qlo;
  P.qhi = s_qhi;
  uint32_t *wq = s_wq[warp];

  unsigned int my_segments = 0;
  unsigned long long my_positions = 0;
  for (int g0 = warp * 32; g0 < ncand; g0 += SX_SCAN_NT) {  // warp-uniform
    // Candidates are sorted by lag, so diagonal lengths rise and fall like a triangle.  Pairing the
    // k-th from the front with the k-th from the back keeps the 32 diagonals of a warp close in
    // length (the loop runs for the longest one).
    const int p = g0 + lane;
    const int c = (p & 1) ? (ncand - 1 - (p >> 1)) : (p >> 1);
    int shift = 0, i0 = 0, L = 0;
    if (p < ncand) {
      shift = (int)cand_pool[cref.x + c] - H;  // pos = idx - N/2 (CrossCorr.cc:600-602)
      i0 = shift < 0 ? -shift : 0;
      int i_end = qlen - shift;               // first i with j >= qlen
      if (tlen - 1 < i_end) i_end = tlen - 1;  // the last target base is never scored
      L = i_end - i0;
      if (L <= 46) L = 0;  // no window is ever evaluated (needs n > 45)
    }
    my_positions += (unsigned long long)L;
    const int nwords = (L + 31) >> 5;
    const int nw_max = __reduce_max_sync(0xffffffffu, nwords);
    const int j0 = i0 + shift;
    const int tw0 = i0 >> 5, tsh = i0 & 31, qw0 = j0 >> 5, qsh = j0 & 31;
    const uint32_t lastmask = (L & 31) ? ((1u << (L & 31)) - 1u) : 0xffffffffu;

    uint32_t m_prev = 0, carry = 0;
    uint32_t s1p0 = 0, s1p1 = 0, s1q0 = 0, s1q1 = 0;
    uint32_t s2p0 = 0, s2p1 = 0, s2p2 = 0, s2q0 = 0, s2q1 = 0, s2q2 = 0;
    uint32_t s3p0 = 0, s3p1 = 0, s3p2 = 0, s3p3 = 0;
    uint32_t s4p0 = 0, s4p1 = 0, s4p2 = 0, s4p3 = 0, s4p4 = 0;
    int open = -1;  // start of the open run in diagonal coordinates, -1 = none
    int cnt = 0;    // segments this lane has found on its diagonal
    // raw plane words are carried from step to step: 4 shared-memory loads per step instead of 8
    uint32_t rtl = s_tlo[min(tw0, NW)], rth = s_thi[min(tw0, NW)], rql = s_qlo[min(qw0, NW)],
             rqh = s_qhi[min(qw0, NW)];

#pragma unroll 2
    for (int kw = 0; kw < nw_max; kw++) {  // warp-uniform trip count
      const int tn = min(tw0 + kw + 1, NW + 1), qn = min(qw0 + kw + 1, NW + 1);
      const uint32_t ntl = s_tlo[tn], nth = s_thi[tn], nql = s_qlo[qn], nqh = s_qhi[qn];
      const uint32_t tl = __funnelshift_r(rtl, ntl, tsh), th = __funnelshift_r(rth, nth, tsh);
      const uint32_t ql = __funnelshift_r(rql, nql, qsh), qh = __funnelshift_r(rqh, nqh, qsh);
      rtl = ntl; rth = nth; rql = nql; rqh = nqh;
      const uint32_t vm = kw < nwords - 1 ? 0xffffffffu : (kw == nwords - 1 ? lastmask : 0u);
      const uint32_t m = ~((tl ^ ql) | (th ^ qh)) & vm;
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
      uint32_t pass = (n5 | (n4 & (n3 | n2 | (n1 & n0)))) & vm;
      if (kw < 2) pass &= (kw == 0) ? 0u : 0xffffc000u;  // warp-uniform branch
      // run boundaries inside this word: rise = a run starts here, fall = first position after a run
      const uint32_t prevp = (pass << 1) | carry;
      const uint32_t rise = pass & ~prevp;
      uint32_t fall = ~pass & prevp;
      carry = pass >> 31;
      const int kb = kw * 32;
      if (__any_sync(0xffffffffu, fall != 0u)) {
        // one iteration per run that ENDS in this word (usually one); lanes work in the same instructions
        do {
          const int f = __ffs(fall) - 1;  // -1: nothing left for this lane
          if (f >= 0) {
            fall &= fall - 1u;
            const uint32_t below = rise & ((1u << f) - 1u);
            // the run started at the closest rise below f, or in an earlier word (carried in `open`)
            const int start_k = below ? (kb + 31 - __clz(below) - 45) : open;  // lastStart = i - m_minLen
            push_segment(i0 + start_k, shift, kb + f - start_k, lane, wq, cnt, spill, spill_cap, seg_tap, seg_tap_cap,
                         ctr);
          }
        } while (__any_sync(0xffffffffu, fall != 0u));
      }
      // state at the end of the word: a run still open started at the last rise of this word, if any
      open = carry ? (rise ? (kb + 31 - __clz(rise) - 45) : open) : -1;
      m_prev = m;
      s1q0 = s1p0; s1q1 = s1p1; s1p0 = s10; s1p1 = s11;
      s2q0 = s2p0; s2q1 = s2p1; s2q2 = s2p2; s2p0 = s20; s2p1 = s21; s2p2 = s22;
      s3p0 = s30; s3p1 = s31; s3p2 = s32; s3p3 = s33;
      s4p0 = s40; s4p1 = s41; s4p2 = s42; s4p3 = s43; s4p4 = s44;
    }
    // a run that reaches the stop position exactly at a word boundary closes there
    if (open >= 0) push_segment(i0 + open, shift, L - open, lane, wq, cnt, spill, spill_cap, seg_tap, seg_tap_cap, ctr);
    my_segments += (unsigned int)cnt;
    __syncwarp();
    // ---- score the segments: every lane walks its own queue ----------------------------------------
    // The queues have different fill levels; spread the segments evenly over the lanes: flat index
    // f -> (owner lane, slot) through an exclusive prefix of the counts.
    const int nq = min(cnt, SX_LQ_CAP);
    int off = nq;  // inclusive scan ...
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, off, o);
      if (lane >= o) off += v;
    }
    const int total = __shfl_sync(0xffffffffu, off, 31);
    off -= nq;  // ... made exclusive
    for (int f0 = 0; f0 < total; f0 += 32) {  // warp-uniform
      const int f = f0 + lane;
      int owner = 0;
#pragma unroll
      for (int step = 16; step >= 1; step >>= 1) {  // largest lane whose offset is <= f
        const int cand = owner + step;
        const int v = __shfl_sync(0xffffffffu, off, cand & 31);
        if (cand < 32 && v <= f) owner = cand;
      }
      const int o_off = __shfl_sync(0xffffffffu, off, owner);
      const int o_shift = __shfl_sync(0xffffffffu, shift, owner);
      if (f < total) {
        const uint32_t q = wq[(f - o_off) * 32 + owner];
        const int start_t = (int)(q & 0xffffu), seg_len = (int)(q >> 16);
        double prob, ident;
        if (score_fast(P, start_t, o_shift, seg_len, prm, prob, ident))
          emit_result(sp, start_t, o_shift, seg_len, prob, ident, res_pool, res_cap, ctr);
      }
    }
    __syncwarp();
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
    // global planes are exactly NW words per plane: stage the (few) words this segment needs with
    // bounds checks into a local window (segment + 1 word on each side)
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
