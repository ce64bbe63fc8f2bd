// K3 (fast path, both chunks pure A/C/G/T): bit-parallel diagonal scan + scoring.
// Included by sx_kernels.cu after score_counts / emit_result / tap_segment are defined.
//
// Replaces SeqAnalyzer::MatchUp / DoOne (analysis/CrossCorr.cc:583-605, 667-724).  For A/C/G/T the
// reference's integer scores are 100 (equal) / 0 (different), so "window sum > 1889" (46-wide
// window; (int)(45*0.42*100) is 1889 in IEEE double) is "at least 19 of the last 46 positions match".
//
// One CTA (4 warps) per chunk pair = two consecutive strand-pairs; the 2-bit base planes of the chunks
// sit in shared memory.  The candidate lags are cut into groups of 32 which the warps take from a
// shared counter, so no warp idles while a sibling still has groups.  A warp works on its group alone,
// one diagonal per lane, in DIAGONAL coordinates k = i - i0, 32 positions (one word of match bits) at a
// time:
//
//  A. filter (warp-uniform loop over the words): match word m = XOR of the planes, then an exact
//     NECESSARY condition for any window ending in this word to reach 19: for each quarter of the word,
//     the popcount of the 53 positions that cover all windows ending in that quarter.  On random DNA
//     ~89 % of the words fail it and need nothing more.  Survivors are recorded as bits.
//  B. every surviving word becomes one work item (lane:5 | word:10); a warp scan gives each lane its
//     slots in the item list, and the list is evaluated 32 items at a time, one per lane: the exact
//     bit-sliced 46-window count (doubling: windows 2,4,8,16,32, then 32+8+4+2; ">= 19" is three logic
//     ops on the six count planes) from the word and its two predecessors -- every window ending in the
//     word lies within those three, so the result is exact -- as straight-line code that computes only
//     the planes that can influence it.  The pass words go to shared memory.
//  C. runs of passing positions are the segments, exactly as the reference's sequential loop emits
//     them.  Every run END (item, bit) is queued as soon as its item is evaluated (slots from a warp
//     scan); the queue is worked off one run per lane: find the start of the run, going back through
//     the previous items of the same diagonal if need be (a word that is not an item has no passing
//     position), then score the segment: popcounts over the planes, an FP32 pre-reject, the exact early
//     reject and the reference's FP64 formula (score_counts).
#pragma once

#define SX_SCAN_NT 128
#define SX_SCAN_WARPS (SX_SCAN_NT / 32)
#define SX_RUN_CAP 256  // run ends (= segments) queued per warp before they are worked off

template <int LOG2N>
struct ScanCfg {
  static constexpr int NW = (1 << LOG2N) / 32;                // words per plane
  static constexpr int PAD = 2;                               // zero words in front of every plane
  static constexpr int PW = NW + PAD + 2;                     // padded plane length (word NW + 1 is still readable)
  // a diagonal runs over target positions, at most N/2 of them: N/64 words
  static constexpr int NBW = (NW / 2 + 31) / 32;              // words of a lane's survivor bitset
  static constexpr int SPC = LOG2N <= 14 ? 2 : 1;             // strand-pairs per CTA
  static constexpr int ITEM_CAP = NW / 2 > 512 ? NW / 2 : 512;  // items listed per round (>= the most one lane can have)
  static constexpr size_t WARP_BYTES = (size_t)SX_RUN_CAP * 4 + (size_t)NBW * 32 * 4 + (size_t)ITEM_CAP * 4 +
                                       (size_t)ITEM_CAP * 2 + 32 * 4;
  static constexpr size_t SMEM = (size_t)SPC * 4 * PW * 4 + SX_SCAN_WARPS * WARP_BYTES;
};

struct PlanePtrs {
  const uint2 *t, *q;  // plane words {lo, hi} of the target / query chunk, readable from word -2 to word NW+1 (zero padded)
};

// 32 positions of both sequences starting at target bit tw*32+tsh / query bit qw*32+qsh
__device__ __forceinline__ void diag_words(const PlanePtrs &P, int tw, int tsh, int qw, int qsh, uint32_t &tl,
                                           uint32_t &th, uint32_t &ql, uint32_t &qh) {
  const uint2 t0 = P.t[tw], t1 = P.t[tw + 1], q0 = P.q[qw], q1 = P.q[qw + 1];
  tl = __funnelshift_r(t0.x, t1.x, tsh);
  th = __funnelshift_r(t0.y, t1.y, tsh);
  ql = __funnelshift_r(q0.x, q1.x, qsh);
  qh = __funnelshift_r(q0.y, q1.y, qsh);
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

// ---- bit-sliced sliding counts -------------------------------------------------------------------
// Plane j of a struct holds bit j of the count of matches in the window ENDING at each of the 32
// positions of a word.  "prev" = the same quantity one word earlier on the diagonal.
struct W2 { uint32_t p0, p1; };              // window 2,  count <= 2
struct W4 { uint32_t p0, p1, p2; };          // window 4,  count <= 4
struct W8 { uint32_t p0, p1, p2, p3; };      // window 8,  count <= 8
struct W16 { uint32_t p0, p1, p2, p3, p4; };  // window 16, count <= 16

__device__ __forceinline__ uint32_t maj3(uint32_t a, uint32_t b, uint32_t c) { return (a & b) | (c & (a ^ b)); }

__device__ __forceinline__ W2 win2(uint32_t m_prev, uint32_t m) {
  const uint32_t m1 = __funnelshift_l(m_prev, m, 1);
  W2 r;
  r.p0 = m ^ m1;
  r.p1 = m & m1;
  return r;
}
__device__ __forceinline__ W4 win4(const W2 &pv, const W2 &c) {  // c + (c delayed by 2)
  const uint32_t a0 = __funnelshift_l(pv.p0, c.p0, 2), a1 = __funnelshift_l(pv.p1, c.p1, 2);
  W4 r;
  r.p0 = c.p0 ^ a0;
  const uint32_t cy = c.p0 & a0;
  r.p1 = c.p1 ^ a1 ^ cy;
  r.p2 = maj3(c.p1, a1, cy);
  return r;
}
__device__ __forceinline__ W8 win8(const W4 &pv, const W4 &c) {  // c + (c delayed by 4)
  const uint32_t b0 = __funnelshift_l(pv.p0, c.p0, 4), b1 = __funnelshift_l(pv.p1, c.p1, 4),
                 b2 = __funnelshift_l(pv.p2, c.p2, 4);
  W8 r;
  r.p0 = c.p0 ^ b0;
  uint32_t cy = c.p0 & b0;
  r.p1 = c.p1 ^ b1 ^ cy;
  cy = maj3(c.p1, b1, cy);
  r.p2 = c.p2 ^ b2 ^ cy;
  r.p3 = maj3(c.p2, b2, cy);
  return r;
}
__device__ __forceinline__ W16 win16(const W8 &pv, const W8 &c) {  // c + (c delayed by 8)
  const uint32_t c0 = __funnelshift_l(pv.p0, c.p0, 8), c1 = __funnelshift_l(pv.p1, c.p1, 8),
                 c2 = __funnelshift_l(pv.p2, c.p2, 8), c3 = __funnelshift_l(pv.p3, c.p3, 8);
  W16 r;
  r.p0 = c.p0 ^ c0;
  uint32_t cy = c.p0 & c0;
  r.p1 = c.p1 ^ c1 ^ cy;
  cy = maj3(c.p1, c1, cy);
  r.p2 = c.p2 ^ c2 ^ cy;
  cy = maj3(c.p2, c2, cy);
  r.p3 = c.p3 ^ c3 ^ cy;
  r.p4 = maj3(c.p3, c3, cy);
  return r;
}
// Positions of the word whose 46-window holds >= 19 matches.  s4 / s4p: window-16 counts of this word and
// the previous one; s3p, s2p, s1p: window-8/4/2 counts of the previous word; s2q, s1q: two words back.
// 46 = 32 (s4 + s4 delayed 16) + 8 (s3 one word back) + 4 (s2 delayed 40) + 2 (s1 delayed 44).
__device__ __forceinline__ uint32_t pass_word(const W16 &s4p, const W16 &s4, const W8 &s3p, const W4 &s2q,
                                              const W4 &s2p, const W2 &s1q, const W2 &s1p) {
  const uint32_t d0 = __funnelshift_l(s4p.p0, s4.p0, 16), d1 = __funnelshift_l(s4p.p1, s4.p1, 16),
                 d2 = __funnelshift_l(s4p.p2, s4.p2, 16), d3 = __funnelshift_l(s4p.p3, s4.p3, 16),
                 d4 = __funnelshift_l(s4p.p4, s4.p4, 16);
  const uint32_t s50 = s4.p0 ^ d0;
  uint32_t cy = s4.p0 & d0;
  const uint32_t s51 = s4.p1 ^ d1 ^ cy;
  cy = maj3(s4.p1, d1, cy);
  const uint32_t s52 = s4.p2 ^ d2 ^ cy;
  cy = maj3(s4.p2, d2, cy);
  const uint32_t s53 = s4.p3 ^ d3 ^ cy;
  cy = maj3(s4.p3, d3, cy);
  const uint32_t s54 = s4.p4 ^ d4 ^ cy;
  const uint32_t s55 = maj3(s4.p4, d4, cy);
  const uint32_t e20 = __funnelshift_l(s2q.p0, s2p.p0, 8), e21 = __funnelshift_l(s2q.p1, s2p.p1, 8),
                 e22 = __funnelshift_l(s2q.p2, s2p.p2, 8);
  const uint32_t e10 = __funnelshift_l(s1q.p0, s1p.p0, 12), e11 = __funnelshift_l(s1q.p1, s1p.p1, 12);
  const uint32_t u0 = e20 ^ e10;  // u = e2 + e1 <= 6
  cy = e20 & e10;
  const uint32_t u1 = e21 ^ e11 ^ cy;
  cy = maj3(e21, e11, cy);
  const uint32_t u2 = e22 ^ cy;
  const uint32_t v0 = s3p.p0 ^ u0;  // v = s3(prev word) + u <= 14
  cy = s3p.p0 & u0;
  const uint32_t v1 = s3p.p1 ^ u1 ^ cy;
  cy = maj3(s3p.p1, u1, cy);
  const uint32_t v2 = s3p.p2 ^ u2 ^ cy;
  cy = maj3(s3p.p2, u2, cy);
  const uint32_t v3 = s3p.p3 ^ cy;
  const uint32_t n0 = s50 ^ v0;  // count = s5 + v <= 46
  cy = s50 & v0;
  const uint32_t n1 = s51 ^ v1 ^ cy;
  cy = maj3(s51, v1, cy);
  const uint32_t n2 = s52 ^ v2 ^ cy;
  cy = maj3(s52, v2, cy);
  const uint32_t n3 = s53 ^ v3 ^ cy;
  cy = maj3(s53, v3, cy);
  const uint32_t n4 = s54 ^ cy;
  const uint32_t n5 = s55 ^ (s54 & cy);
  return n5 | (n4 & (n3 | n2 | (n1 & n0)));  // count >= 19 (0b010011)
}
// a * b + c as one multiply-add on the FMA pipe, also when b is a power of two (the logic pipe is this kernel's limit)
__device__ __forceinline__ uint32_t mad_u32(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}
// exclusive prefix sum over the warp; *total = sum over all lanes
__device__ __forceinline__ unsigned int warp_excl_scan(unsigned int v, unsigned int *total) {
  const int lane = threadIdx.x & 31;
  unsigned int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  *total = __shfl_sync(0xffffffffu, incl, 31);
  return incl - v;
}

// Exact pass word of diagonal word kw (kw < d.nwords): the three match words kw-2 .. kw from four consecutive
// words of every plane, then only the count planes that can reach the result: word kw-2 feeds windows 2 / 4
// (their top bits), word kw-1 windows 2 .. 16.
__device__ __forceinline__ uint32_t eval_word(const PlanePtrs &P, const Diag &d, int kw) {
  const uint2 *tp = P.t + d.tw0 + kw - 2, *qp = P.q + d.qw0 + kw - 2;
  const uint2 t0 = tp[0], t1 = tp[1], t2 = tp[2], t3 = tp[3];
  const uint2 q0 = qp[0], q1 = qp[1], q2 = qp[2], q3 = qp[3];
  const int ts = d.tsh, qs = d.qsh;
  uint32_t mA = ~((__funnelshift_r(t0.x, t1.x, ts) ^ __funnelshift_r(q0.x, q1.x, qs)) |
                  (__funnelshift_r(t0.y, t1.y, ts) ^ __funnelshift_r(q0.y, q1.y, qs)));
  uint32_t mB = ~((__funnelshift_r(t1.x, t2.x, ts) ^ __funnelshift_r(q1.x, q2.x, qs)) |
                  (__funnelshift_r(t1.y, t2.y, ts) ^ __funnelshift_r(q1.y, q2.y, qs)));
  uint32_t mC = ~((__funnelshift_r(t2.x, t3.x, ts) ^ __funnelshift_r(q2.x, q3.x, qs)) |
                  (__funnelshift_r(t2.y, t3.y, ts) ^ __funnelshift_r(q2.y, q3.y, qs)));
  if (kw < 2) mA = 0u;  // before the diagonal starts
  if (kw < 1) mB = 0u;
  if (kw == d.nwords - 1) mC &= d.lastmask;
  const W2 zero2 = {0u, 0u};
  const W8 zero8 = {0u, 0u, 0u, 0u};
  // only the top bits of word A's planes are ever read: empty history below them is exact
  const W2 s1A = win2(0u, mA);
  const W4 s2A = win4(zero2, s1A);
  const W2 s1B = win2(mA, mB);
  const W4 s2B = win4(s1A, s1B);
  const W8 s3B = win8(s2A, s2B);
  const W16 s4B = win16(zero8, s3B);  // only its top 16 bits are read: they depend on s3B alone
  const W2 s1C = win2(mB, mC);
  const W4 s2C = win4(s1B, s1C);
  const W8 s3C = win8(s2B, s2C);
  const W16 s4C = win16(s3B, s3C);
  uint32_t pass = pass_word(s4B, s4C, s3B, s2A, s2B, s1A, s1B);
  // a window is only evaluated once 46 positions have been seen (k >= 46) and inside the diagonal
  if (kw == d.nwords - 1) pass &= d.lastmask;
  if (kw < 2) pass &= (kw == 0) ? 0u : 0xffffc000u;
  return pass;
}

template <int LOG2N>
__global__ void __launch_bounds__(SX_SCAN_NT, 8)
    scan_score_kernel(const SpDesc *__restrict__ sps, int nsp, Slots ws, const uint16_t *__restrict__ cand_pool,
                      const uint2 *__restrict__ cand_ref, ScoreParams prm, ResultRec *__restrict__ res_pool,
                      unsigned int res_cap, SegRec *__restrict__ seg_tap, unsigned int seg_tap_cap, BatchCounters *ctr) {
  using C = ScanCfg<LOG2N>;
  constexpr int N = 1 << LOG2N, H = N / 2, NW = C::NW, PW = C::PW, NBW = C::NBW, SPC = C::SPC, PAD = C::PAD;
  constexpr int ITEM_CAP = C::ITEM_CAP;
  extern __shared__ __align__(16) unsigned char scan_smem[];
  uint2 *s_planes = reinterpret_cast<uint2 *>(scan_smem);  // [SPC][target, query][PW] plane words {lo, hi}
  __shared__ unsigned int s_next;
  __shared__ int s_ncand[SPC];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned char *wbase = scan_smem + (size_t)SPC * 4 * PW * 4 + (size_t)warp * C::WARP_BYTES;
  uint32_t *wq = reinterpret_cast<uint32_t *>(wbase);                         // SX_RUN_CAP run ends: item | position << 16
  uint32_t(*s_need)[32] = reinterpret_cast<uint32_t(*)[32]>(wq + SX_RUN_CAP);  // [NBW][32]
  uint32_t *passw = reinterpret_cast<uint32_t *>(s_need + NBW);               // ITEM_CAP
  uint16_t *items = reinterpret_cast<uint16_t *>(passw + ITEM_CAP);           // ITEM_CAP
  int *s_shift = reinterpret_cast<int *>(items + ITEM_CAP);                   // 32

  const int sp0 = blockIdx.x * SPC;
  if (tid < SPC) {
    int nc = 0;
    const int spi = sp0 + tid;
    if (spi < nsp) {
      const uint2 cref = cand_ref[spi];
      const SpDesc sp = sps[spi];
      const bool generic = ((ws.meta[sp.t_slot].flags | ws.meta[sp.q_slot].flags) & SLOT_NONACGT) != 0;  // other kernel
      if (cref.x != 0xffffffffu && !generic) nc = (int)cref.y;
      if (cref.x != 0xffffffffu && generic && cref.y > 0) atomicAdd(&ctr->n_generic, 1u);
    }
    s_ncand[tid] = nc;
  }
  if (tid == 0) s_next = 0;
  __syncthreads();
  int ncand[SPC], gstart[SPC + 1];
  gstart[0] = 0;
#pragma unroll
  for (int j = 0; j < SPC; j++) {
    ncand[j] = s_ncand[j];
    gstart[j + 1] = gstart[j] + ((ncand[j] + 31) >> 5);
  }
  if (gstart[SPC] == 0) return;
#pragma unroll
  for (int j = 0; j < SPC; j++) {
    if (ncand[j] == 0) continue;
    const SpDesc sp = sps[sp0 + j];
    const uint32_t *tp = ws.planes + (size_t)sp.t_slot * 2 * NW;
    const uint32_t *qp = ws.planes + (size_t)sp.q_slot * 2 * NW;
    uint2 *pl = s_planes + (size_t)j * 2 * PW;
    for (int i = tid; i < PW; i += SX_SCAN_NT) {
      const int w = i - PAD;
      const bool in = w >= 0 && w < NW;
      pl[i] = in ? make_uint2(__ldg(tp + w), __ldg(tp + NW + w)) : make_uint2(0u, 0u);
      pl[PW + i] = in ? make_uint2(__ldg(qp + w), __ldg(qp + NW + w)) : make_uint2(0u, 0u);
    }
  }
  __syncthreads();

  unsigned int my_segments = 0;
  unsigned long long my_positions = 0;
  unsigned int qn = 0;  // run ends in this warp's queue (warp-uniform)
  const unsigned int run_cap = prm.run_cap > 0 ? min((unsigned int)prm.run_cap, (unsigned int)SX_RUN_CAP) : SX_RUN_CAP;
  const int run_min = seg_tap != nullptr ? 0 : prm.run_min;  // the segment tap wants every raw segment
  const bool byte_filter = run_min >= 15;
  for (;;) {  // groups of 32 candidate lags, handed out dynamically
    unsigned int g = 0;
    if (lane == 0) g = atomicAdd(&s_next, 1u);
    g = __shfl_sync(0xffffffffu, g, 0);
    if (g >= (unsigned int)gstart[SPC]) break;
    int j = 0, nc = ncand[0], gbase = 0;
#pragma unroll
    for (int jj = 1; jj < SPC; jj++)
      if ((int)g >= gstart[jj]) {
        j = jj;
        nc = ncand[jj];
        gbase = gstart[jj];
      }
    const int spi = sp0 + j;
    const SpDesc sp = sps[spi];
    const uint2 cref = cand_ref[spi];
    const int tlen = ws.meta[sp.t_slot].len, qlen = ws.meta[sp.q_slot].len;
    PlanePtrs P;
    P.t = s_planes + (size_t)j * 2 * PW + PAD;
    P.q = P.t + PW;

    // One queued run end (item i, position f in 0..32 of its word: the first failing position after the run,
    // 32 = the run reaches the end of the word and the next word holds no passing position) -> its segment ->
    // probability filter.  A run starts at its first passing position minus 45 ("lastStart = i - m_minLen",
    // CrossCorr.cc:700); the start is in this word or, through words that pass everywhere, in an earlier item
    // of the same diagonal.  The run ends at f or where the diagonal stops.
    auto do_end = [&](int i, int f) {
      const uint32_t it = items[i], pass = passw[i];
      const Diag od = make_diag(s_shift[it & 31u], tlen, qlen);
      const int kb = (int)(it >> 5) * 32;
      const int e = min(kb + f, od.L);
      // highest failing position below f, if any, is right before the start
      uint32_t below = ~pass & (f >= 32 ? 0xffffffffu : ((1u << f) - 1u));
      int start_k;
      if (below) {
        start_k = kb + (32 - __clz(below));
      } else {  // the run covers the word from bit 0: it began in an earlier item (or at bit 0 of this one)
        int j = i - 1, back = kb;
        uint32_t jt = it - 32u;
        for (;;) {
          if (j < 0 || items[j] != jt) {  // the previous word is no item: the run starts at this word's first bit
            start_k = back;
            break;
          }
          const uint32_t pj = passw[j];
          if (pj != 0xffffffffu) {
            start_k = (pj >> 31) ? (back - __clz(~pj)) : back;
            break;
          }
          back -= 32;
          j--;
          jt -= 32u;
        }
      }
      // run_min pruning (ScoreParams::run_min): a run of r passing windows that follows a failing window holds
      // at most 18 + r matches in its 45 + r positions; the host has verified that no such segment passes the
      // probability filter for r < run_min.  Runs that start at the first evaluated window (k = 46) are exempt.
      if (e - start_k < run_min && start_k > 46) return;
      start_k -= 45;
      const int start_t = od.i0 + start_k, seg_len = e - start_k;
      tap_segment(spi, start_t, od.shift, seg_len, seg_tap, seg_tap_cap, ctr);
      double prob, ident;
      if (score_fast(P, start_t, od.shift, seg_len, prm, prob, ident))
        emit_result(sp, start_t, od.shift, seg_len, prob, ident, res_pool, res_cap, ctr);
    };
    // work off the queued run ends, one per lane (warp-collective); they only look back, so this may happen
    // at any time
    auto flush_queue = [&]() {
      __syncwarp();
      for (int r = lane; r < (int)qn; r += 32) {
        const uint32_t rec = wq[r];
        do_end((int)(rec & 0xffffu), (int)(rec >> 16));
      }
      __syncwarp();
      qn = 0;
    };

    // Candidates are sorted by lag, so diagonal lengths rise and fall like a triangle.  Pairing the
    // k-th from the front with the k-th from the back keeps the 32 diagonals of a warp close in length.
    const int p = ((int)g - gbase) * 32 + lane;
    const int c = (p & 1) ? (nc - 1 - (p >> 1)) : (p >> 1);
    int shift = 0;
    Diag d = make_diag(0, 0, 0);
    if (p < nc) {
      shift = (int)cand_pool[cref.x + c] - H;  // pos = idx - N/2 (CrossCorr.cc:600-602)
      d = make_diag(shift, tlen, qlen);
    }
    s_shift[lane] = shift;
    my_positions += (unsigned long long)d.L;
    const int nw_max = __reduce_max_sync(0xffffffffu, d.nwords);
    const int nbw = (nw_max + 31) >> 5;

    // ---- A. filter -------------------------------------------------------------------------------
    // Quarter q of word kw (positions 32kw+8q .. +7): every window ending there lies inside the 53
    // positions [32kw+8q-45, 32kw+8q+7].  From the previous words only a few popcounts are needed:
    //   q0: m[kw-2] bits 19..31 + m[kw-1]            + m[kw] bits 0..7
    //   q1: m[kw-2] bits 27..31 + m[kw-1]            + m[kw] bits 0..15
    //   q2:                       m[kw-1] bits 3..31  + m[kw] bits 0..23
    //   q3:                       m[kw-1] bits 11..31 + m[kw]
    unsigned int n_items = 0;
    {
      const uint2 *pt = P.t + d.tw0, *pq = P.q + d.qw0;
      uint2 rt = pt[0], rq = pq[0];
      // match word kw of this lane's diagonal; FULL = every lane's diagonal covers the whole word (no masking,
      // no clamping).  Returns false past the end of the diagonal (m = 0 there).
      auto match_word = [&](int kw, bool full, uint32_t &m) -> bool {
        const int kn = full ? kw + 1 : min(kw + 1, d.nwords);  // a shorter diagonal idles, never reads past its planes
        const uint2 nt = pt[kn], nq = pq[kn];
        const uint32_t tl = __funnelshift_r(rt.x, nt.x, d.tsh), th = __funnelshift_r(rt.y, nt.y, d.tsh);
        const uint32_t ql = __funnelshift_r(rq.x, nq.x, d.qsh), qh = __funnelshift_r(rq.y, nq.y, d.qsh);
        rt = nt;
        rq = nq;
        m = ~((tl ^ ql) | (th ^ qh));
        if (full) return true;
        // the diagonal's last word is partial, words after it are empty
        m &= kw < d.nwords - 1 ? 0xffffffffu : (kw == d.nwords - 1 ? d.lastmask : 0u);
        return kw < d.nwords;
      };
      // words [0, nw_full) are full on every diagonal of the group (empty lanes do not count: their bits are dropped)
      const bool live = d.nwords > 0;
      const int nw_full = min(__reduce_min_sync(0xffffffffu, live ? d.nwords - 1 : 0x7fffffff), nw_max);
      if (!byte_filter) {
        // The four sums ride one register, a byte each (nothing exceeds 53): pre = {c_lo8, c_lo16, c_lo24, c_all} of
        // m[kw] (four popcounts, packed by multiply-adds on the FMA pipe), ge = {popc(m >> 3), >> 11, >> 19, >> 27} =
        // c_all - (pre << 8) - p3 with p3 = the popcounts of the low three bits of every byte, taken in parallel
        // (y - (y >> 1) - (y >> 2) on 3-bit fields).  u = pre + {ge19, ge27 of m[kw-2] ; ge3, ge11 of m[kw-1]} +
        // c_all(m[kw-1]) * {1, 1, 0, 0}.  Four popcounts per word instead of eight: the XU pipe (16 lanes per clock)
        // was this loop's limit.
        uint32_t g1 = 0, g2 = 0, call1 = 0;  // ge of m[kw-1], of m[kw-2]; c_all of m[kw-1]
        uint32_t acc = 0, bit = 1u;
        auto step = [&](int kw, bool full) {
          uint32_t m;
          const bool inside = match_word(kw, full, m);
          const uint32_t call = (uint32_t)__popc(m);
          uint32_t pre = (uint32_t)__popc(mad_u32(m, 0x1000000u, 0u));
          pre = mad_u32((uint32_t)__popc(mad_u32(m, 0x10000u, 0u)), 0x100u, pre);
          pre = mad_u32((uint32_t)__popc(mad_u32(m, 0x100u, 0u)), 0x10000u, pre);
          pre = mad_u32(call, 0x1000000u, pre);
          const uint32_t y = m & 0x07070707u;
          const uint32_t p3 = y - ((y >> 1) & 0x03030303u) - ((y >> 2) & 0x01010101u);
          const uint32_t ge = call * 0x01010101u - (pre << 8) - p3;
          const uint32_t u = pre + __byte_perm(g2, g1, 0x5432) + call1 * 0x00000101u;
          if (inside && ((u + 0x6d6d6d6du) & 0x80808080u)) acc |= bit;  // some byte >= 19
          bit <<= 1;
          g2 = g1;
          g1 = ge;
          call1 = call;
        };
        for (int w = 0; w < nbw; w++) {  // warp-uniform
          acc = 0;
          bit = 1u;
          const int kw_end = min(nw_max, w * 32 + 32), kw_full = max(w * 32, min(nw_full, kw_end));
          int kw = w * 32;
#pragma unroll 4
          for (; kw < kw_full; kw++) step(kw, true);  // warp-uniform trip counts
#pragma unroll 2
          for (; kw < kw_end; kw++) step(kw, false);
          if (!live) acc = 0;
          s_need[w][lane] = acc;
          n_items += __popc(acc);
        }
      } else {
        // Byte filter (run_min >= 15).  Only runs of >= run_min passing windows are of interest; such a run
        // contains a whole aligned byte of positions [8g, 8g+7].  Its last window, ending at 8g+7, lies inside bytes
        // g-5 .. g: S6(g) = sum of their match counts >= 19; its first window, ending at 8g, lies inside bytes
        // g-6 .. g-1 plus one position of byte g: S7(g) - c(g) + 1 >= 19.  F marks the words with such a byte.
        // A run also touches words where it covers no whole byte; every word between those and the byte's word is
        // covered entirely, hence in F.  Where the run crosses from word i into word i+1 the top window of word i
        // passes (its top byte has S6 >= 19: H) and the bottom window of word i+1 passes (S7 of its bottom byte
        // >= 19: L): X(i) = H(i) & L(i+1).  items = F | X(i) & F(i+1) | X(i-1) & F(i-1).
        // Word 1 holds the first evaluated window (k = 46, positions 1 .. 46, inside bytes 0 .. 5): a run that
        // starts there is scored whatever its length, so word 1 is an item when those bytes hold 19 matches.
        // Counts are kept four bytes to a register: c = match counts of the word's bytes, c * 0x01010101 their
        // prefix sums; nothing exceeds 56, so the byte lanes never carry into each other.
        uint32_t cp = 0, cpp = 0;  // byte counts of words kw-1, kw-2
        uint32_t bn = 0;           // for word kw-1: {T, T, T - P0, T - P1}: what its bytes add to S6 of the next word's bytes
        uint32_t accF = 0, accH = 0, accL = 0, bit = 1u;
        auto step = [&](int kw, bool full) {
          uint32_t m;
          const bool inside = match_word(kw, full, m);
          // byte j of pre: matches in bytes 0 .. j of the word (three masks, four popcounts, packed by
          // multiply-adds, which run beside the logic pipe); c = pre - (pre << 8): the bytes' own counts
          // (the low bytes are isolated by multiplications -- left shifts on the FMA pipe -- instead of AND masks)
          uint32_t pre = (uint32_t)__popc(mad_u32(m, 0x1000000u, 0u));
          pre = mad_u32((uint32_t)__popc(mad_u32(m, 0x10000u, 0u)), 0x100u, pre);
          pre = mad_u32((uint32_t)__popc(mad_u32(m, 0x100u, 0u)), 0x10000u, pre);
          pre = mad_u32((uint32_t)__popc(m), 0x1000000u, pre);
          const uint32_t c = pre * 0xffffff01u;
          const uint32_t s6 = pre + bn + (cpp >> 24);               // byte j: counts of bytes g-5 .. g
          const uint32_t s7 = s6 + __byte_perm(cpp, cp, 0x5432);    // byte j: counts of bytes g-6 .. g
          const uint32_t t6 = (s6 + 0x6d6d6d6du) & 0x80808080u;     // bytes with S6 >= 19
          const uint32_t t7 = (s7 - c + 0x6e6e6e6eu) & 0x80808080u;  // bytes with S7 - c >= 18
          if (inside && (t6 & t7)) accF |= bit;
          // H and L are single bits of words that exist anyway: they are shifted in from the top, one funnel shift each
          // (the block's words end up in reversed order; put right after the loop), instead of a test and a predicated OR
          const uint32_t l31 = mad_u32(s7, 0x1000000u, 0x6d000000u);  // bit 31: S7 of the bottom byte >= 19
          accH = __funnelshift_l(inside ? t6 : 0u, accH, 1);
          accL = __funnelshift_l(inside ? l31 : 0u, accL, 1);
          bit <<= 1;
          const uint32_t tot = pre >> 24;
          bn = tot * 0x01010101u - (pre << 16);
          cpp = cp;
          cp = c;
        };
        bool w1 = false;
        if (d.nwords > 1) {
          uint32_t tl, th, ql, qh;
          diag_words(P, d.tw0, d.tsh, d.qw0, d.qsh, tl, th, ql, qh);
          int cnt = __popc(~((tl ^ ql) | (th ^ qh)));
          diag_words(P, d.tw0 + 1, d.tsh, d.qw0 + 1, d.qsh, tl, th, ql, qh);
          cnt += __popc(~((tl ^ ql) | (th ^ qh)) & 0xffffu & (d.nwords == 2 ? d.lastmask : 0xffffffffu));
          w1 = cnt >= 19;
        }
        uint32_t carryF = 0u, carryH = 0u;  // F / H of word 32w - 1
        for (int w = 0; w < nbw; w++) {  // warp-uniform
          accF = accH = accL = 0;
          bit = 1u;
          const int kw_end = min(nw_max, w * 32 + 32), kw_full = max(w * 32, min(nw_full, kw_end));
          int kw = w * 32;
#pragma unroll 4
          for (; kw < kw_full; kw++) step(kw, true);  // warp-uniform trip counts
#pragma unroll 2
          for (; kw < kw_end; kw++) step(kw, false);
          {
            const int n_steps = kw_end - w * 32;  // 1 .. 32: word k of the block sits at bit n_steps - 1 - k
            accH = __brev(accH) >> (32 - n_steps);
            accL = __brev(accL) >> (32 - n_steps);
          }
          const uint32_t x = accH & (accL >> 1);  // crossings between words of this block
          uint32_t acc = accF | (x & (accF >> 1)) | ((x & accF) << 1);
          if (carryH & accL & 1u) {  // a run may cross from the previous block's last word into this block's first
            if (carryF) acc |= 1u;
            if ((accF & 1u) && live && !(s_need[w - 1][lane] >> 31)) {
              s_need[w - 1][lane] |= 0x80000000u;
              n_items++;
            }
          }
          carryF = accF >> 31;
          carryH = accH >> 31;
          if (w == 0 && w1) acc |= 2u;
          if (!live) acc = 0;
          s_need[w][lane] = acc;
          n_items += __popc(acc);
        }
      }
    }
    __syncwarp();

    // ---- B. one item per surviving word; rounds of as many lanes as fit the item list (normally one round)
    unsigned int total_items;
    const unsigned int off_all = warp_excl_scan(n_items, &total_items);
    unsigned int done_upto = 0;  // items of lanes whose offset is below this have been handled
    while (done_upto < total_items) {  // warp-uniform
      // lanes [first, last): the longest run of lanes from the first unhandled one whose items fit the list
      const bool fits = off_all >= done_upto && off_all + n_items <= done_upto + ITEM_CAP;
      const unsigned int fit_mask = __ballot_sync(0xffffffffu, fits);
      const unsigned int started = __ballot_sync(0xffffffffu, off_all >= done_upto);
      const int first = __ffs(started) - 1;
      // first lane at or after `first` that does not fit ends the round
      const unsigned int nofit_after = ~fit_mask & (0xffffffffu << first);
      const int last = nofit_after ? (__ffs(nofit_after) - 1) : 32;
      const bool mine = lane >= first && lane < last;
      const unsigned int round_base = __shfl_sync(0xffffffffu, off_all, first);
      const unsigned int round_end =
          last < 32 ? __shfl_sync(0xffffffffu, off_all, last & 31) : total_items;
      const int n_round = (int)(round_end - round_base);
      if (mine) {
        unsigned int o = off_all - round_base;
        for (int w = 0; w < nbw; w++) {
          uint32_t b = s_need[w][lane];
          while (b) {
            const int pos = __ffs(b) - 1;
            b &= b - 1u;
            items[o++] = (uint16_t)(lane | ((w * 32 + pos) << 5));
          }
        }
      }
      __syncwarp();
      // B2 + C1. exact pass words, one item per lane; every run end found is queued (slots from a warp scan)
      uint32_t last_it = 0x80000000u, last_pass = 0u;  // item / pass word of the previous batch's last lane
      for (int i0 = 0; i0 < n_round; i0 += 32) {  // warp-uniform
        const int i = i0 + lane;
        uint32_t it = 0x80000000u, pass = 0u, next_it = 0x80000000u;
        if (i < n_round) {
          it = items[i];
          if (i + 1 < n_round) next_it = items[i + 1];
          const Diag od = make_diag(s_shift[it & 31u], tlen, qlen);
          pass = eval_word(P, od, (int)(it >> 5));
          passw[i] = pass;
        }
        // the previous item in list order: same diagonal and previous word -> its top bit carries into this word
        uint32_t p_it = __shfl_up_sync(0xffffffffu, it, 1), p_pass = __shfl_up_sync(0xffffffffu, pass, 1);
        if (lane == 0) {
          p_it = last_it;
          p_pass = last_pass;
        }
        last_it = __shfl_sync(0xffffffffu, it, 31);
        last_pass = __shfl_sync(0xffffffffu, pass, 31);
        const uint32_t carry_in = (p_it == it - 32u) ? (p_pass >> 31) : 0u;
        uint32_t fall = ~pass & ((pass << 1) | carry_in);
        // a run that reaches the end of the word ends there unless the next word is an item too
        const bool edge = (pass >> 31) && next_it != it + 32u;
        my_segments += (unsigned int)__popc(fall) + (edge ? 1u : 0u);
        if (run_min > 1) {
          // Runs that lie inside this word and are shorter than run_min never reach the queue (see do_end).  Kept:
          // run ends with >= run_min passing positions right below them (erosion of the pass word), the lowest run
          // end when the run reaches bit 0 (it may have begun in an earlier word; do_end decides) and, in word 1,
          // the run that starts at the first evaluated window (bit 14).
          uint32_t er = pass;
          for (int have = 1; have < run_min && er; ) {  // warp-uniform trip count except for the early exit
            const int sft = min(have, run_min - have);
            er &= er >> sft;
            have += sft;
          }
          uint32_t keep = run_min < 32 ? (er << run_min) : 0u;
          keep |= (~pass) & (0u - (~pass));  // lowest failing position: the run below it (if any) reaches bit 0
          if ((it >> 5) == 1u) {
            const uint32_t z = ~(pass | 0x3fffu);
            keep |= z & (0u - z);
          }
          fall &= keep;
        }
        const unsigned int mine_n = (unsigned int)__popc(fall) + (edge ? 1u : 0u);
        unsigned int tot;
        unsigned int slot = warp_excl_scan(mine_n, &tot);
        if (tot != 0) {
          if (qn + tot > run_cap) flush_queue();
          if (tot > run_cap) {  // more run ends in one batch than the queue holds: done in place
            __syncwarp();          // (passw of this batch is visible)
            while (fall) {
              const int f = __ffs(fall) - 1;
              fall &= fall - 1u;
              do_end(i, f);
            }
            if (edge) do_end(i, 32);
          } else {
            slot += qn;
            qn += tot;
            while (fall) {
              const int f = __ffs(fall) - 1;
              fall &= fall - 1u;
              wq[slot++] = (uint32_t)i | ((uint32_t)f << 16);
            }
            if (edge) wq[slot++] = (uint32_t)i | (32u << 16);
          }
        }
        __syncwarp();
      }
      flush_queue();  // C2. before the item list is reused (the planes and `sp` of this group are still current)
      done_upto = round_end;
    }
  }
  my_segments = __reduce_add_sync(0xffffffffu, my_segments);
  if (lane == 0 && my_segments) atomicAdd(&ctr->n_segments, (unsigned long long)my_segments);
  for (int o = 16; o > 0; o >>= 1) my_positions += __shfl_xor_sync(0xffffffffu, my_positions, o);
  if (lane == 0 && my_positions) atomicAdd(&ctr->n_positions, my_positions);
}
