// libsatsuma_b200: k-mer seeding on the GPU (include/satsuma_kmatch.h) -- SURVEY 8(f) rank 4.
//
// Replaces the reference's KMatch program (kmatch/KMatch.cc): canonical k-mer positions of both genomes
// (kmer_array_from_fasta, :20-147), sorted, filtered by frequency, joined on equal k-mers (merge_positions, :156-189),
// the k-mer matches sorted by query position and chained along their diagonals into blocks
// (dump_matching_blocks, :196-318) that SatsumaSynteny2 loads as seed t_result records.
// The reference does this with std::sort on one thread per genome and a sequential sweep with a list of open matches.
// Here every step is data-parallel: a window kernel, an 8-bit LSD radix sort (histogram / scan / stable scatter),
// flag + scan compactions, a binary-search join, and the sweep restated as "maximal chains of same-diagonal matches
// whose query positions are at most max_jump apart" found by a second stable sort on the diagonal and a max-scan.
// Quirks of the reference that decide what is emitted are kept (see the comments marked Q).  No CPU fallback.
#include "../../include/satsuma_kmatch.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace {

thread_local std::string g_kerr;
int kfail(int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_kerr = buf;
  return code;
}
#define KCU(expr)                                                                                      \
  do {                                                                                                 \
    cudaError_t _e = (expr);                                                                           \
    if (_e != cudaSuccess)                                                                             \
      return kfail(_e == cudaErrorMemoryAllocation ? SX_ERR_NOMEM : SX_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

constexpr long long CHR = 10000000000LL;  // KMATCH_POSITION_CHR_CNST (kmatch/KMatch.h:19)

template <typename T>
struct Dev {
  T *p = nullptr;
  size_t n = 0;
  ~Dev() { if (p) cudaFree(p); }
  cudaError_t alloc(size_t want) {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
    if (want == 0) want = 1;
    cudaError_t e = cudaMalloc((void **)&p, want * sizeof(T));
    if (e == cudaSuccess) n = want;
    return e;
  }
};

// ---- generic scans over 32-bit values (three-phase: per-block scan + block totals, scan of the totals, add back) ----
struct OpAdd { __device__ static unsigned int id() { return 0u; } __device__ static unsigned int f(unsigned int a, unsigned int b) { return a + b; } };
struct OpMax { __device__ static unsigned int id() { return 0u; } __device__ static unsigned int f(unsigned int a, unsigned int b) { return a > b ? a : b; } };

constexpr int SCAN_NT = 256, SCAN_IPT = 4, SCAN_TILE = SCAN_NT * SCAN_IPT;

// inclusive scan of a tile; block total to sums[blockIdx.x]
template <class Op>
__global__ void __launch_bounds__(SCAN_NT) scan_tiles(const unsigned int *__restrict__ in, unsigned int *__restrict__ out,
                                                     unsigned int *__restrict__ sums, size_t n) {
  __shared__ unsigned int s_warp[SCAN_NT / 32];
  const size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_IPT;
  unsigned int v[SCAN_IPT], acc = Op::id();
#pragma unroll
  for (int i = 0; i < SCAN_IPT; i++) {
    v[i] = base + i < n ? in[base + i] : Op::id();
    acc = Op::f(acc, v[i]);
    v[i] = acc;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned int incl = acc;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl = Op::f(t, incl);
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  unsigned int wbase = Op::id();
  for (int w = 0; w < warp; w++) wbase = Op::f(wbase, s_warp[w]);
  const unsigned int excl_thread = Op::f(wbase, __shfl_up_sync(0xffffffffu, incl, 1));
  const unsigned int tbase = lane == 0 ? wbase : excl_thread;
#pragma unroll
  for (int i = 0; i < SCAN_IPT; i++)
    if (base + i < n) out[base + i] = Op::f(tbase, v[i]);
  if (threadIdx.x == SCAN_NT - 1 && sums) sums[blockIdx.x] = Op::f(wbase, incl);
}
template <class Op>
__global__ void __launch_bounds__(SCAN_NT) scan_add_back(unsigned int *__restrict__ out, const unsigned int *__restrict__ sums_incl, size_t n) {
  if (blockIdx.x == 0) return;
  const unsigned int add = sums_incl[blockIdx.x - 1];
  const size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_IPT;
#pragma unroll
  for (int i = 0; i < SCAN_IPT; i++)
    if (base + i < n) out[base + i] = Op::f(add, out[base + i]);
}
// out[i] = op(in[0..i]) (inclusive), in place allowed
template <class Op>
int scan_inclusive(const unsigned int *in, unsigned int *out, size_t n, cudaStream_t st) {
  if (n == 0) return SX_OK;
  const size_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
  Dev<unsigned int> sums;
  KCU(sums.alloc(nb));
  scan_tiles<Op><<<(unsigned int)nb, SCAN_NT, 0, st>>>(in, out, sums.p, n);
  KCU(cudaGetLastError());
  if (nb > 1) {
    int rc = scan_inclusive<Op>(sums.p, sums.p, nb, st);
    if (rc != SX_OK) return rc;
    scan_add_back<Op><<<(unsigned int)nb, SCAN_NT, 0, st>>>(out, sums.p, n);
    KCU(cudaGetLastError());
  }
  KCU(cudaStreamSynchronize(st));  // `sums` is freed on return
  return SX_OK;
}

// ---- stable LSD radix sort of (u64 key, u64 value) pairs, 8 bits per pass ----------------------------------------
constexpr int RS_NT = 256, RS_IPT = 8, RS_TILE = RS_NT * RS_IPT;

__global__ void __launch_bounds__(RS_NT) rs_histogram(const unsigned long long *__restrict__ keys, size_t n, int shift,
                                                      unsigned int *__restrict__ ghist, unsigned int nblocks) {
  __shared__ unsigned int h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const size_t base = (size_t)blockIdx.x * RS_TILE;
#pragma unroll
  for (int r = 0; r < RS_IPT; r++) {
    const size_t i = base + (size_t)r * RS_NT + threadIdx.x;
    if (i < n) atomicAdd(&h[(unsigned int)(keys[i] >> shift) & 255u], 1u);
  }
  __syncthreads();
  ghist[(size_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];  // digit-major: one scan gives every tile its bases
}
// offsets = EXCLUSIVE scan of ghist (passed as inclusive scan `gscan`; exclusive value = inclusive - own count)
__global__ void __launch_bounds__(RS_NT) rs_scatter(const unsigned long long *__restrict__ keys, const unsigned long long *__restrict__ vals,
                                                    unsigned long long *__restrict__ okeys, unsigned long long *__restrict__ ovals,
                                                    size_t n, int shift, const unsigned int *__restrict__ ghist,
                                                    const unsigned int *__restrict__ gscan, unsigned int nblocks) {
  __shared__ unsigned int s_base[256];           // global position of the next element of every digit of this tile
  __shared__ unsigned int s_wcnt[RS_NT / 32][256];  // per warp: elements of every digit in the current round
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  {
    const size_t gi = (size_t)threadIdx.x * nblocks + blockIdx.x;
    s_base[threadIdx.x] = gscan[gi] - ghist[gi];
  }
  const size_t base = (size_t)blockIdx.x * RS_TILE;
  for (int r = 0; r < RS_IPT; r++) {  // rounds of 256 consecutive elements keep the sort stable
#pragma unroll
    for (int w = 0; w < RS_NT / 32; w++) s_wcnt[w][threadIdx.x] = 0;
    __syncthreads();
    const size_t i = base + (size_t)r * RS_NT + threadIdx.x;
    const bool in = i < n;
    unsigned long long k = 0, v = 0;
    unsigned int d = 0, rank = 0;
    if (in) {
      k = keys[i];
      v = vals[i];
      d = (unsigned int)(k >> shift) & 255u;
    }
    const unsigned int act = __ballot_sync(0xffffffffu, in);
    if (in) {
      const unsigned int peers = __match_any_sync(act, d);
      rank = __popc(peers & ((1u << lane) - 1u));
      if (rank == 0) s_wcnt[warp][d] = __popc(peers);
    }
    __syncthreads();
    unsigned int pos = 0;
    if (in) {
      unsigned int before = 0;
      for (int w = 0; w < warp; w++) before += s_wcnt[w][d];
      pos = s_base[d] + before + rank;
    }
    __syncthreads();
    {
      unsigned int tot = 0;
#pragma unroll
      for (int w = 0; w < RS_NT / 32; w++) tot += s_wcnt[w][threadIdx.x];
      s_base[threadIdx.x] += tot;
    }
    if (in) {
      okeys[pos] = k;
      ovals[pos] = v;
    }
    __syncthreads();
  }
}
// sorts by bits [0, nbits) of the key; result ends up in (k0, v0) or (k1, v1): returns which through *in_first
int radix_sort(unsigned long long *k0, unsigned long long *v0, unsigned long long *k1, unsigned long long *v1, size_t n,
               int nbits, bool *in_first, cudaStream_t st) {
  *in_first = true;
  if (n == 0) return SX_OK;
  if (n >= 0xffffffffull) return kfail(SX_ERR_ARG, "kmatch: more than 2^32 elements in one sort");
  const unsigned int nblocks = (unsigned int)((n + RS_TILE - 1) / RS_TILE);
  Dev<unsigned int> ghist, gscan;
  KCU(ghist.alloc((size_t)256 * nblocks));
  KCU(gscan.alloc((size_t)256 * nblocks));
  unsigned long long *ki = k0, *vi = v0, *ko = k1, *vo = v1;
  for (int shift = 0; shift < nbits; shift += 8) {
    rs_histogram<<<nblocks, RS_NT, 0, st>>>(ki, n, shift, ghist.p, nblocks);
    KCU(cudaGetLastError());
    int rc = scan_inclusive<OpAdd>(ghist.p, gscan.p, (size_t)256 * nblocks, st);
    if (rc != SX_OK) return rc;
    rs_scatter<<<nblocks, RS_NT, 0, st>>>(ki, vi, ko, vo, n, shift, ghist.p, gscan.p, nblocks);
    KCU(cudaGetLastError());
    std::swap(ki, ko);
    std::swap(vi, vo);
    *in_first = !*in_first;
  }
  KCU(cudaStreamSynchronize(st));
  return SX_OK;
}

// ---- step A: canonical k-mer of every window (kmer_array_from_fasta, KMatch.cc:44-112) ---------------------------
// window w of sequence s starts at base p (0-based); position = p + 1 + (s + 1) * CHR, negative when the reverse
// complement is the smaller k-mer.  Q: the reference sizes its array for EVERY window but writes only the valid ones
// (no letter outside ACGTacgt among the K bases) from the front; the rest stay value-initialised {kmer 0, position 0}
// and take part in the sort and the frequency filter like real entries.  Invalid windows are emitted as exactly that.
__global__ void __launch_bounds__(256) kmer_windows(const unsigned char *__restrict__ bases, const long long *__restrict__ seq_off,
                                                    const long long *__restrict__ win_base, int n_seq, int K, size_t n_win,
                                                    unsigned long long *__restrict__ keys, unsigned long long *__restrict__ vals) {
  const size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n_win) return;
  int lo = 0, hi = n_seq - 1;  // sequence of this window: last s with win_base[s] <= w
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if ((size_t)win_base[mid] <= w) lo = mid; else hi = mid - 1;
  }
  const long long p = (long long)(w - (size_t)win_base[lo]);
  const unsigned char *b = bases + seq_off[lo] + p;
  unsigned long long f = 0, r = 0;
  bool ok = true;
  for (int i = 0; i < K; i++) {
    unsigned int c;
    switch (b[i]) {
      case 'A': case 'a': c = 0; break;
      case 'C': case 'c': c = 1; break;
      case 'G': case 'g': c = 2; break;
      case 'T': case 't': c = 3; break;
      default: c = 0; ok = false; break;
    }
    f = (f << 2) | c;
    r = (r >> 2) | ((unsigned long long)(3u - c) << (2 * (K - 1)));
  }
  unsigned long long key = 0;
  long long pos = 0;
  if (ok) {
    const long long P = p + 1 + (long long)(lo + 1) * CHR;
    if (f <= r) { key = f; pos = P; } else { key = r; pos = -P; }
  }
  keys[w] = key;
  vals[w] = (unsigned long long)pos;
}

// ---- step C: frequency filter (KMatch.cc:121-139): groups of more than max_freq equal k-mers are dropped ----------
__global__ void __launch_bounds__(256) freq_keep(const unsigned long long *__restrict__ keys, size_t n, int max_freq,
                                                 unsigned int *__restrict__ keep) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // i sits in a group larger than max_freq iff some max_freq + 1 consecutive entries around it are equal
  const unsigned long long k = keys[i];
  bool big = false;
  const size_t a0 = i >= (size_t)max_freq ? i - (size_t)max_freq : 0;
  for (size_t a = a0; a <= i && !big; a++)
    if (a + (size_t)max_freq < n && keys[a] == k && keys[a + (size_t)max_freq] == k) big = true;
  keep[i] = big ? 0u : 1u;
}
__global__ void __launch_bounds__(256) compact_pairs(const unsigned long long *__restrict__ keys, const unsigned long long *__restrict__ vals,
                                                     const unsigned int *__restrict__ keep, const unsigned int *__restrict__ incl, size_t n,
                                                     unsigned long long *__restrict__ okeys, unsigned long long *__restrict__ ovals) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !keep[i]) return;
  okeys[incl[i] - 1] = keys[i];
  ovals[incl[i] - 1] = vals[i];
}

// ---- step D: join on equal k-mers (merge_positions, KMatch.cc:156-189) ---------------------------------------------
__device__ __forceinline__ size_t lower_bound_u64(const unsigned long long *a, size_t n, unsigned long long k) {
  size_t lo = 0, hi = n;
  while (lo < hi) {
    const size_t mid = (lo + hi) >> 1;
    if (a[mid] < k) lo = mid + 1; else hi = mid;
  }
  return lo;
}
__global__ void __launch_bounds__(256) join_count(const unsigned long long *__restrict__ qk, size_t nq, const unsigned long long *__restrict__ tk,
                                                  size_t nt, unsigned int *__restrict__ cnt, unsigned int *__restrict__ first) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  const unsigned long long k = qk[i];
  const size_t lb = lower_bound_u64(tk, nt, k);
  unsigned int c = 0;
  while (lb + c < nt && tk[lb + c] == k) c++;
  cnt[i] = c;
  first[i] = (unsigned int)lb;
}
// one k-mer match: key = query position, value = target position | reverse << 63
__global__ void __launch_bounds__(256) join_write(const unsigned long long *__restrict__ qv, size_t nq, const unsigned long long *__restrict__ tv,
                                                  const unsigned int *__restrict__ cnt, const unsigned int *__restrict__ first,
                                                  const unsigned int *__restrict__ incl, unsigned long long *__restrict__ mk,
                                                  unsigned long long *__restrict__ mv) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq || cnt[i] == 0) return;
  const long long qp = (long long)qv[i];
  size_t o = (size_t)incl[i] - cnt[i];
  for (unsigned int j = 0; j < cnt[i]; j++, o++) {
    const long long tp = (long long)tv[first[i] + j];
    bool rev = false;
    long long q = qp, t = tp;
    if (!(qp > 0)) { rev = true; q = -qp; }    // KMatch.cc:171-176 (a phantom position 0 counts as "reverse")
    if (!(tp > 0)) { rev = !rev; t = -tp; }    // :177-182
    mk[o] = (unsigned long long)q;
    mv[o] = (unsigned long long)t | ((unsigned long long)(rev ? 1 : 0) << 63);
  }
}

// ---- steps F-G: blocks = maximal chains on one diagonal with query gaps <= max_jump (dump_matching_blocks) -------
// diagonal key of a match: orientation in the top bit, then t - q (forward) or t + q (reverse), biased to be positive
__global__ void __launch_bounds__(256) diag_keys(const unsigned long long *__restrict__ mk, const unsigned long long *__restrict__ mv, size_t n,
                                                 unsigned long long *__restrict__ dk, unsigned long long *__restrict__ dv) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long q = (long long)mk[i], t = (long long)(mv[i] & 0x7fffffffffffffffull);
  const unsigned long long rev = mv[i] >> 63;
  const long long d = rev ? t + q : t - q + (1ll << 61);
  dk[i] = (rev << 62) | (unsigned long long)d;  // q, t < 2^60 (CHR * sequences)
  dv[i] = (unsigned long long)q;                 // t follows from the diagonal
}
__global__ void __launch_bounds__(256) chain_heads(const unsigned long long *__restrict__ dk, const unsigned long long *__restrict__ dq, size_t n,
                                                   long long max_jump, unsigned int *__restrict__ head_idx) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const bool head = i == 0 || dk[i] != dk[i - 1] || (long long)(dq[i] - dq[i - 1]) > max_jump;
  head_idx[i] = head ? (unsigned int)i + 1u : 0u;  // 1-based so that a max-scan carries the last head forward
}
__global__ void __launch_bounds__(256) chain_emit_flags(const unsigned long long *__restrict__ dk, const unsigned long long *__restrict__ dq,
                                                        const unsigned int *__restrict__ head_scan, size_t n, long long max_jump, int K,
                                                        int min_length, unsigned int *__restrict__ emit) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const bool tail = i + 1 == n || dk[i + 1] != dk[i] || (long long)(dq[i + 1] - dq[i]) > max_jump;
  unsigned int e = 0;
  if (tail) {
    const size_t h = (size_t)head_scan[i] - 1;
    const long long length = (long long)(dq[i] - dq[h]);
    if (length + K >= min_length) e = 1;  // KMatch.cc:224
  }
  emit[i] = e;
}
__global__ void __launch_bounds__(256) chain_emit(const unsigned long long *__restrict__ dk, const unsigned long long *__restrict__ dq,
                                                  const unsigned int *__restrict__ head_scan, const unsigned int *__restrict__ emit,
                                                  const unsigned int *__restrict__ emit_incl, size_t n, int K,
                                                  const long long *__restrict__ q_len, int n_q, sx_result *__restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !emit[i]) return;
  const size_t h = (size_t)head_scan[i] - 1;
  const unsigned long long key = dk[h];
  const bool rev = (key >> 62) & 1ull;
  const long long d = (long long)(key & ((1ull << 62) - 1ull));
  const long long q_start = (long long)dq[h];
  const long long t_start = rev ? d - q_start : d - (1ll << 61) + q_start;
  const long long length = (long long)(dq[i] - dq[h]);
  sx_result r;
  memset(&r, 0, sizeof(r));
  const long long qid = q_start / CHR - 1;  // KMatch.cc:228-241
  r.query_id = (unsigned long long)qid;
  r.target_id = (unsigned long long)(t_start / CHR - 1);
  r.query_size = qid >= 0 && qid < n_q ? (unsigned long long)q_len[qid] : 0ull;
  r.qstart = (unsigned long long)(q_start % CHR - 1);
  r.tstart = (unsigned long long)((rev ? t_start - length : t_start) % CHR - 1);
  r.len = (unsigned long long)(length + K);
  r.reverse = rev ? 1 : 0;
  r.prob = 1.;
  r.ident = 1.;
  out[emit_incl[i] - 1] = r;
}

struct Genome {  // device copy of one FASTA's sequences
  Dev<unsigned char> bases;
  Dev<long long> seq_off, win_base, lens;
  size_t n_win = 0;
  int n_seq = 0;
};

int upload_genome(Genome &g, const char *bases, const int64_t *offsets, const int64_t *lens, int32_t n, int K, const char *what) {
  if (n < 0 || (n > 0 && (!bases || !offsets || !lens))) return kfail(SX_ERR_ARG, "sx_kmatch: %s: null argument", what);
  std::vector<long long> off((size_t)n), wb((size_t)n + 1, 0), ln((size_t)n);
  int64_t blob = 0;
  for (int i = 0; i < n; i++) {
    if (lens[i] < 0 || offsets[i] < 0) return kfail(SX_ERR_ARG, "sx_kmatch: %s: negative length or offset", what);
    // the reference grows its array by size + 1 - K (unsigned): a sequence shorter than K - 1 wraps around and aborts
    if (lens[i] < (int64_t)K - 1) return kfail(SX_ERR_ARG, "sx_kmatch: %s: sequence %d is shorter than K - 1 (%lld bases)", what, i, (long long)lens[i]);
    off[(size_t)i] = offsets[i];
    ln[(size_t)i] = lens[i];
    wb[(size_t)i + 1] = wb[(size_t)i] + (lens[i] + 1 - K);
    blob = std::max<int64_t>(blob, offsets[i] + lens[i]);
  }
  g.n_seq = n;
  g.n_win = (size_t)wb[(size_t)n];
  KCU(g.bases.alloc((size_t)blob + 32));
  KCU(g.seq_off.alloc((size_t)n));
  KCU(g.win_base.alloc((size_t)n + 1));
  KCU(g.lens.alloc((size_t)n));
  if (blob) KCU(cudaMemcpy(g.bases.p, bases, (size_t)blob, cudaMemcpyHostToDevice));
  if (n) {
    KCU(cudaMemcpy(g.seq_off.p, off.data(), sizeof(long long) * n, cudaMemcpyHostToDevice));
    KCU(cudaMemcpy(g.lens.p, ln.data(), sizeof(long long) * n, cudaMemcpyHostToDevice));
  }
  KCU(cudaMemcpy(g.win_base.p, wb.data(), sizeof(long long) * ((size_t)n + 1), cudaMemcpyHostToDevice));
  return SX_OK;
}

inline unsigned int grid_for(size_t n) { return (unsigned int)((n + 255) / 256); }

// windows -> sorted -> frequency-filtered (k-mer, position) list of one genome; *n_out entries in (keys, vals)
int kmer_list(const Genome &g, int K, int max_freq, Dev<unsigned long long> &keys, Dev<unsigned long long> &vals, size_t *n_out,
              cudaStream_t st) {
  *n_out = 0;
  const size_t n = g.n_win;
  if (n == 0) return SX_OK;
  Dev<unsigned long long> k0, v0, k1, v1;
  KCU(k0.alloc(n)); KCU(v0.alloc(n)); KCU(k1.alloc(n)); KCU(v1.alloc(n));
  kmer_windows<<<grid_for(n), 256, 0, st>>>(g.bases.p, g.seq_off.p, g.win_base.p, g.n_seq, K, n, k0.p, v0.p);
  KCU(cudaGetLastError());
  bool first;
  int rc = radix_sort(k0.p, v0.p, k1.p, v1.p, n, 2 * K, &first, st);
  if (rc != SX_OK) return rc;
  unsigned long long *sk = first ? k0.p : k1.p, *sv = first ? v0.p : v1.p;
  unsigned long long *ok = first ? k1.p : k0.p, *ov = first ? v1.p : v0.p;
  Dev<unsigned int> keep, incl;
  KCU(keep.alloc(n)); KCU(incl.alloc(n));
  freq_keep<<<grid_for(n), 256, 0, st>>>(sk, n, max_freq, keep.p);
  KCU(cudaGetLastError());
  if ((rc = scan_inclusive<OpAdd>(keep.p, incl.p, n, st)) != SX_OK) return rc;
  unsigned int kept = 0;
  KCU(cudaMemcpy(&kept, incl.p + (n - 1), sizeof(kept), cudaMemcpyDeviceToHost));
  compact_pairs<<<grid_for(n), 256, 0, st>>>(sk, sv, keep.p, incl.p, n, ok, ov);
  KCU(cudaGetLastError());
  // Q: "kposv.resize(wi - 1)" (KMatch.cc:141-146): the last kept entry is cut off as well
  const size_t m = kept > 0 ? (size_t)kept - 1 : 0;
  KCU(keys.alloc(m)); KCU(vals.alloc(m));
  if (m) {
    KCU(cudaMemcpyAsync(keys.p, ok, sizeof(unsigned long long) * m, cudaMemcpyDeviceToDevice, st));
    KCU(cudaMemcpyAsync(vals.p, ov, sizeof(unsigned long long) * m, cudaMemcpyDeviceToDevice, st));
  }
  KCU(cudaStreamSynchronize(st));
  *n_out = m;
  return SX_OK;
}

}  // namespace

extern "C" const char *sx_kmatch_last_error(void) { return g_kerr.c_str(); }

extern "C" void sx_kmatch_default_config(sx_kmatch_config *c) {
  memset(c, 0, sizeof(*c));
  c->k = 31;
  c->max_freq = 1;   // SatsumaSynteny2 -max_seed_kmer_freq (analysis/SatsumaSynteny2.cc:249)
  c->min_length = 31;
  c->max_jump = 30;
  c->device = 0;
}

extern "C" int sx_kmatch(const sx_kmatch_config *cfg, const char *q_bases, const int64_t *q_offsets, const int64_t *q_lens,
                         int32_t n_q, const char *t_bases, const int64_t *t_offsets, const int64_t *t_lens, int32_t n_t,
                         sx_result *out, int64_t cap, int64_t *n_out, sx_kmatch_stats *stats) {
  if (!cfg || !n_out) return kfail(SX_ERR_ARG, "sx_kmatch: null argument");
  *n_out = 0;
  if (stats) memset(stats, 0, sizeof(*stats));
  const int K = cfg->k;
  if (K < 3 || K > 31 || (K & 1) == 0) return kfail(SX_ERR_ARG, "sx_kmatch: K = %d (odd values 3 .. 31)", K);  // KMatch.cc:325
  if (cfg->max_freq < 1 || cfg->max_freq > 4096) return kfail(SX_ERR_ARG, "sx_kmatch: max_freq = %d (1 .. 4096)", cfg->max_freq);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return kfail(SX_ERR_CUDA, "sx_kmatch: no CUDA device; this library has no CPU fallback");
  if (cfg->device < 0 || cfg->device >= ndev) return kfail(SX_ERR_ARG, "sx_kmatch: device %d of %d", cfg->device, ndev);
  KCU(cudaSetDevice(cfg->device));
  KCU(cudaFree(0));  // context creation is not part of the timed device work
  cudaStream_t st = nullptr;
  const auto t_begin = std::chrono::steady_clock::now();
  struct Timer {
    sx_kmatch_stats *s;
    std::chrono::steady_clock::time_point t0;
    ~Timer() {
      if (s) s->gpu_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
  } timer{stats, t_begin};
  Genome Q, T;
  int rc = upload_genome(Q, q_bases, q_offsets, q_lens, n_q, K, "query");
  if (rc == SX_OK) rc = upload_genome(T, t_bases, t_offsets, t_lens, n_t, K, "target");
  if (rc != SX_OK) return rc;
  Dev<unsigned long long> qk, qv, tk, tv;
  size_t nq = 0, nt = 0;
  if ((rc = kmer_list(Q, K, cfg->max_freq, qk, qv, &nq, st)) != SX_OK) return rc;
  if ((rc = kmer_list(T, K, cfg->max_freq, tk, tv, &nt, st)) != SX_OK) return rc;
  if (stats) {
    stats->query_windows = (int64_t)Q.n_win;
    stats->target_windows = (int64_t)T.n_win;
    stats->query_kmers = (int64_t)nq;
    stats->target_kmers = (int64_t)nt;
  }
  if (nq == 0 || nt == 0) return SX_OK;
  // join
  Dev<unsigned int> cnt, first, incl;
  KCU(cnt.alloc(nq)); KCU(first.alloc(nq)); KCU(incl.alloc(nq));
  join_count<<<grid_for(nq), 256, 0, st>>>(qk.p, nq, tk.p, nt, cnt.p, first.p);
  KCU(cudaGetLastError());
  if ((rc = scan_inclusive<OpAdd>(cnt.p, incl.p, nq, st)) != SX_OK) return rc;
  unsigned int n_match = 0;
  KCU(cudaMemcpy(&n_match, incl.p + (nq - 1), sizeof(n_match), cudaMemcpyDeviceToHost));
  if (stats) stats->kmer_matches = (int64_t)n_match;
  if (n_match < 2) return SX_OK;  // Q: the sweep starts at the second match (KMatch.cc:206)
  Dev<unsigned long long> mk0, mv0, mk1, mv1;
  KCU(mk0.alloc(n_match)); KCU(mv0.alloc(n_match)); KCU(mk1.alloc(n_match)); KCU(mv1.alloc(n_match));
  join_write<<<grid_for(nq), 256, 0, st>>>(qv.p, nq, tv.p, cnt.p, first.p, incl.p, mk0.p, mv0.p);
  KCU(cudaGetLastError());
  // sort by query position (KMatch.cc:200); positions stay below CHR * (sequences + 1) < 2^60
  bool in_first;
  int qbits = 1;
  while (qbits < 62 && (1ull << qbits) <= (unsigned long long)CHR * (unsigned long long)(n_q + 1)) qbits++;
  if ((rc = radix_sort(mk0.p, mv0.p, mk1.p, mv1.p, n_match, qbits, &in_first, st)) != SX_OK) return rc;
  unsigned long long *sk = in_first ? mk0.p : mk1.p, *sv = in_first ? mv0.p : mv1.p;
  unsigned long long *wk = in_first ? mk1.p : mk0.p, *wv = in_first ? mv1.p : mv0.p;
  // Q: the match with the smallest query position never takes part (the loop runs from i = 1; KMatch.cc:206).
  // Q: at its last step the reference reads one element past the end of the match array (i == kmsize) before it
  //    flushes the open matches; that read is undefined and is taken as "extends nothing" here.
  const size_t m = (size_t)n_match - 1;
  diag_keys<<<grid_for(m), 256, 0, st>>>(sk + 1, sv + 1, m, wk, wv);
  KCU(cudaGetLastError());
  // stable sort on the diagonal keeps the query order inside every diagonal
  if ((rc = radix_sort(wk, wv, sk, sv, m, 63, &in_first, st)) != SX_OK) return rc;
  const unsigned long long *dk = in_first ? wk : sk, *dq = in_first ? wv : sv;
  Dev<unsigned int> head, emit, emit_incl;
  KCU(head.alloc(m)); KCU(emit.alloc(m)); KCU(emit_incl.alloc(m));
  chain_heads<<<grid_for(m), 256, 0, st>>>(dk, dq, m, (long long)cfg->max_jump, head.p);
  KCU(cudaGetLastError());
  if ((rc = scan_inclusive<OpMax>(head.p, head.p, m, st)) != SX_OK) return rc;
  chain_emit_flags<<<grid_for(m), 256, 0, st>>>(dk, dq, head.p, m, (long long)cfg->max_jump, K, cfg->min_length, emit.p);
  KCU(cudaGetLastError());
  if ((rc = scan_inclusive<OpAdd>(emit.p, emit_incl.p, m, st)) != SX_OK) return rc;
  unsigned int n_blocks = 0;
  KCU(cudaMemcpy(&n_blocks, emit_incl.p + (m - 1), sizeof(n_blocks), cudaMemcpyDeviceToHost));
  *n_out = (int64_t)n_blocks;
  if (stats) stats->blocks = (int64_t)n_blocks;
  if ((int64_t)n_blocks > cap || (n_blocks > 0 && !out))
    return kfail(SX_ERR_CAPACITY, "sx_kmatch: %u blocks, buffer holds %lld", n_blocks, (long long)cap);
  if (n_blocks == 0) return SX_OK;
  Dev<sx_result> d_out;
  KCU(d_out.alloc(n_blocks));
  chain_emit<<<grid_for(m), 256, 0, st>>>(dk, dq, head.p, emit.p, emit_incl.p, m, K, Q.lens.p, n_q, d_out.p);
  KCU(cudaGetLastError());
  KCU(cudaMemcpy(out, d_out.p, sizeof(sx_result) * n_blocks, cudaMemcpyDeviceToHost));
  return SX_OK;
}
