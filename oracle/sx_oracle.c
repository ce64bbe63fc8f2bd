/* TEST INFRASTRUCTURE ONLY -- see sx_oracle.h.  CPU restatement (plain C) of Satsuma2's
 * chunk-pair cross-correlation path, written from the behaviour documented in SURVEY.md
 * section 8 and pinned against the compiled reference.  Citations: /root/reference paths.
 *
 * Build: gcc -O2 -std=c11 -ffp-contract=off (no FMA contraction: the reference's x86-64
 * build has none, and the double-precision stages are compared bit-for-bit).
 */
#define _GNU_SOURCE
#include "sx_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * a4. Codec: IUPAC letter -> (A,C,G,T) fractions and complement.
 * Follows DNACodec::DNACodec (analysis/DNAVector.cc:13-58): ACGT=1; K,M,R,Y,S,W=1/2;
 * B,V,H,D=1/3; N,X=1/4; '-'=0; everything else 0 with complement NUL.
 * ------------------------------------------------------------------------------------------ */
static double g_frac[256][4];
static char g_comp[256];
static double g_equal[256][256];
static int g_score[256][256];
static int g_ready = 0;
static pthread_once_t g_once = PTHREAD_ONCE_INIT;

static void put(int ch, double a, double c, double g, double t, int comp) {
  g_frac[ch][0] = a;
  g_frac[ch][1] = c;
  g_frac[ch][2] = g;
  g_frac[ch][3] = t;
  g_comp[ch] = (char)comp;
}

static void build_tables(void) {
  const double third = 1. / 3.;
  memset(g_frac, 0, sizeof(g_frac));
  memset(g_comp, 0, sizeof(g_comp));
  put('A', 1, 0, 0, 0, 'T');
  put('C', 0, 1, 0, 0, 'G');
  put('G', 0, 0, 1, 0, 'C');
  put('T', 0, 0, 0, 1, 'A');
  put('K', 0, 0, .5, .5, 'M');
  put('M', .5, .5, 0, 0, 'K');
  put('R', .5, 0, .5, 0, 'Y');
  put('Y', 0, .5, 0, .5, 'R');
  put('S', 0, .5, .5, 0, 'S');
  put('W', .5, 0, 0, .5, 'W');
  put('B', 0, third, third, third, 'V');
  put('V', third, third, third, 0, 'B');
  put('H', third, third, 0, third, 'D');
  put('D', third, 0, third, third, 'H');
  put('-', 0, 0, 0, 0, '-');
  put('N', .25, .25, .25, .25, 'N');
  put('X', .25, .25, .25, .25, 'X');
  for (int a = 0; a < 256; a++) {
    for (int b = 0; b < 256; b++) {
      /* DNA_Equal (DNAVector.cc:390-403): dot product, summed A,C,G,T in that order */
      double pa = g_frac[a][0] * g_frac[b][0];
      double pc = g_frac[a][1] * g_frac[b][1];
      double pg = g_frac[a][2] * g_frac[b][2];
      double pt = g_frac[a][3] * g_frac[b][3];
      double dot = pa + pc + pg + pt;
      g_equal[a][b] = dot;
      /* DNA_EqualAmb (DNAVector.cc:372-387): identical letters other than 'N' count 1 */
      double amb = (a == b && a != 'N') ? 1. : dot;
      /* LookupMatch ctor (CrossCorr.cc:547-553): integer score, scale 100 */
      g_score[a][b] = (int)(amb * 100. + 0.5);
    }
  }
  g_ready = 1;
}

static inline void ensure_tables(void) {
  if (!g_ready) pthread_once(&g_once, build_tables);
}

void sxo_codec(int byte, double acgt[4]) {
  ensure_tables();
  memcpy(acgt, g_frac[byte & 255], 4 * sizeof(double));
}
char sxo_rc_base(int byte) {
  ensure_tables();
  return g_comp[byte & 255];
}
double sxo_equal(int a, int b) {
  ensure_tables();
  return g_equal[a & 255][b & 255];
}
int sxo_score(int a, int b) {
  ensure_tables();
  return g_score[a & 255][b & 255];
}

/* DNAVector::ReverseComplement (DNAVector.cc:482-521) */
void sxo_revcomp(const char *in, int len, char *out) {
  ensure_tables();
  for (int i = 0; i < len; i++) out[i] = g_comp[(unsigned char)in[len - 1 - i]];
}

/* ------------------------------------------------------------------------------------------
 * a1-a3. Signal encoding.
 * ------------------------------------------------------------------------------------------ */
static double ent_term(double p) { /* Ent, CrossCorr.cc:28-33 */
  if (p < 0.001) return 0;
  return p * log(p) / 0.69314718056;
}

void sxo_encode(const char *bases, int len, int N, float *out5) {
  ensure_tables();
  float *ent = out5;
  memset(out5, 0, sizeof(float) * 5 * (size_t)N);
  /* ComputeEntropy, CrossCorr.cc:35-93 */
  if (len < 1024) {
    for (int i = 0; i < N; i++) ent[i] = 1.f;
  } else {
    int win = N / 512;
    for (int i = 0; i < len; i += win) {
      double s4[4] = {0, 0, 0, 0};
      int k = 0;
      for (int j = i; j < i + win && j < len; j++) {
        const double *f = g_frac[(unsigned char)bases[j]];
        for (int c = 0; c < 4; c++) s4[c] += f[c];
        k++;
      }
      for (int c = 0; c < 4; c++) s4[c] /= (double)k;
      double s = ent_term(s4[0]) + ent_term(s4[1]) + ent_term(s4[2]) + ent_term(s4[3]);
      float v = (float)(-s);
      if (v < 0.) v = 0.f;
      for (int j = i; j < i + k; j++) ent[j] = v;
    }
  }
  /* SeqToPCM, CrossCorr.cc:97-134: mean over the real length, padding stays 0 */
  for (int c = 0; c < 4; c++) {
    float *sig = out5 + (size_t)(c + 1) * N;
    double sum = 0;
    for (int i = 0; i < len; i++) sum += g_frac[(unsigned char)bases[i]][c];
    double off = sum / (double)len;
    for (int i = 0; i < len; i++) sig[i] = (float)((double)ent[i] * (g_frac[(unsigned char)bases[i]][c] - off));
  }
}

/* ------------------------------------------------------------------------------------------
 * c1+c2. Cross-correlation, float64 model of FFTReal-based DoOne/CrossCorrelate.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  double re, im;
} cplx;

/* In-place iterative radix-2 DFT, X[k] = sum x[n] e^{sign*2*pi*i*k*n/N}. */
static void fft_c(cplx *a, int n, int sign) {
  for (int i = 1, j = 0; i < n; i++) {
    int bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) {
      cplx t = a[i];
      a[i] = a[j];
      a[j] = t;
    }
  }
  for (int len = 2; len <= n; len <<= 1) {
    int half = len >> 1;
    for (int k = 0; k < half; k++) {
      double ang = sign * 2.0 * M_PI * (double)k / (double)len;
      double wr = cos(ang), wi = sin(ang);
      for (int s = 0; s < n; s += len) {
        cplx u = a[s + k];
        cplx v = a[s + k + half];
        double tr = v.re * wr - v.im * wi;
        double ti = v.re * wi + v.im * wr;
        a[s + k].re = u.re + tr;
        a[s + k].im = u.im + ti;
        a[s + k + half].re = u.re - tr;
        a[s + k + half].im = u.im - ti;
      }
    }
  }
}

/* ---- the reference's twiddle drift from N = 16384 on ---------------------------------
 * FFTReal builds a real transform of length 2M from two of length M ("pass" p, M = 2^p) with twiddles
 * cos/sin(i*pi/M), i < M/2.  Up to p = 12 they come from an exact table; for p > 12 (FFTReal.h:78) from
 * OscSinCos (OscSinCos.hpp:88-96): a float rotation recurrence restarted at (1, 0) for every group
 * (FFTReal.hpp:605-656 forward, 783-830 inverse), whose rounding errors accumulate to ~1e-4 after some
 * thousand steps.  The model below keeps every other operation in float64 and uses exactly those float
 * twiddles, so it follows the reference (not the ideal transform) at N = 16384 / 32768. */
static void osc_twiddles(int M, double *wr, double *wi) { /* w[i] = c_i + j s_i for i < M/2; i = 0 exact */
  const float sc = (float)cos(M_PI / (double)M), ss = (float)sin(M_PI / (double)M);
  float pc = 1.f, ps = 0.f;
  wr[0] = 1.;
  wi[0] = 0.;
  for (int i = 1; i < M / 2; i++) {
    const float oc = pc, os = ps;
    const float a = oc * sc, b = os * ss, c = oc * ss, d = os * sc; /* separate roundings: no FMA on x86-64 */
    pc = a - b;
    ps = c + d;
    wr[i] = (double)pc;
    wi[i] = (double)ps;
  }
}
/* twiddle the reference effectively applies to bin i < M of the odd half when it builds length 2M:
 * i < M/2: the oscillator value; i = M/2: exact (the "extreme coefficients" are handled separately,
 * FFTReal.hpp:622-626); i > M/2: -conj of the value at M - i (those bins are produced as conjugates of the
 * lower ones, dfi[-i] / dfi[nbr_coef - i]).  `osc` = NULL: the exact twiddle (passes <= 12). */
static void eff_twiddle(const double *owr, const double *owi, int M, int i, double *re, double *im) {
  if (!owr || i == 0 || 2 * i == M) {
    const double ang = M_PI * (double)i / (double)M;
    *re = cos(ang);
    *im = sin(ang);
  } else if (2 * i < M) {
    *re = owr[i];
    *im = owi[i];
  } else {
    *re = -owr[M - i];
    *im = owi[M - i];
  }
}
/* forward transform as FFTReal::do_fft computes it (positive exponent): exact below M = 8192, oscillator
 * twiddles in the passes that build lengths 16384 and 32768 */
static void fft_ref_forward(cplx *a, int n) {
  if (n < 16384) {
    fft_c(a, n, +1);
    return;
  }
  for (int i = 1, j = 0; i < n; i++) { /* bit reversal */
    int bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) {
      cplx t = a[i];
      a[i] = a[j];
      a[j] = t;
    }
  }
  for (int len = 2; len <= n; len <<= 1) {
    const int half = len >> 1; /* = M */
    double *owr = NULL, *owi = NULL;
    if (half >= 8192) { /* pass p = log2(M) > 12 */
      owr = (double *)malloc(sizeof(double) * half);
      owi = (double *)malloc(sizeof(double) * half);
      osc_twiddles(half, owr, owi);
    }
    for (int k = 0; k < half; k++) {
      double wr, wi;
      eff_twiddle(owr, owi, half, k, &wr, &wi);
      for (int s = 0; s < n; s += len) {
        cplx u = a[s + k], v = a[s + k + half];
        const double tr = v.re * wr - v.im * wi, ti = v.re * wi + v.im * wr;
        a[s + k].re = u.re + tr;
        a[s + k].im = u.im + ti;
        a[s + k + half].re = u.re - tr;
        a[s + k + half].im = u.im - ti;
      }
    }
    free(owr);
    free(owi);
  }
}
/* inverse as FFTReal::do_ifft computes it (negative exponent, unscaled): it SPLITS first -- E = G[k] + G[k+M],
 * O = (G[k] - G[k+M]) * conj(twiddle) -- so the oscillator twiddles act on the spectrum, level by level, before
 * the exact shorter transforms; even time samples come from E, odd ones from O. */
static void fft_ref_inverse(cplx *g, int n) {
  if (n < 16384) {
    fft_c(g, n, -1);
    return;
  }
  const int M = n / 2;
  double *owr = (double *)malloc(sizeof(double) * M), *owi = (double *)malloc(sizeof(double) * M);
  osc_twiddles(M, owr, owi);
  cplx *e = (cplx *)malloc(sizeof(cplx) * M), *o = (cplx *)malloc(sizeof(cplx) * M);
  for (int k = 0; k < M; k++) {
    double wr, wi;
    eff_twiddle(owr, owi, M, k, &wr, &wi);
    const cplx s = {g[k].re + g[k + M].re, g[k].im + g[k + M].im}, d = {g[k].re - g[k + M].re, g[k].im - g[k + M].im};
    e[k] = s;
    o[k].re = d.re * wr + d.im * wi; /* d * conj(w) */
    o[k].im = d.im * wr - d.re * wi;
  }
  free(owr);
  free(owi);
  fft_ref_inverse(e, M);
  fft_ref_inverse(o, M);
  for (int m = 0; m < M; m++) {
    g[2 * m] = e[m];
    g[2 * m + 1] = o[m];
  }
  free(e);
  free(o);
}

void sxo_xcorr(const float *tsig4, const float *qsig4, int N, float *out) {
  int H = N / 2;
  cplx *f1 = (cplx *)malloc(sizeof(cplx) * N);
  cplx *f2 = (cplx *)malloc(sizeof(cplx) * N);
  cplx *g = (cplx *)malloc(sizeof(cplx) * N);
  for (int i = 0; i < N; i++) out[i] = 0.f; /* CrossCorrelate: out.resize(N, 0) cc:388-389 */
  for (int c = 0; c < 4; c++) {
    const float *t = tsig4 + (size_t)c * N;
    const float *q = qsig4 + (size_t)c * N;
    for (int i = 0; i < N; i++) {
      f1[i].re = t[i];
      f1[i].im = 0;
      f2[i].re = q[i];
      f2[i].im = 0;
    }
    /* FFTReal::do_fft uses the positive exponent (extern/RealFFT/readme.txt:127) */
    fft_ref_forward(f1, N);
    fft_ref_forward(f2, N);
    /* DoOne, CrossCorr.cc:477-492: DC real*real; bins 1..H-2 get conj(F1)*F2;
     * bins H-1 and H are left as F1 (quirk Q1). */
    g[0].re = f1[0].re * f2[0].re;
    g[0].im = 0;
    for (int k = 1; k <= H - 2; k++) {
      /* conj(F1)*F2 */
      g[k].re = f1[k].re * f2[k].re + f1[k].im * f2[k].im;
      g[k].im = f1[k].re * f2[k].im - f1[k].im * f2[k].re;
    }
    g[H - 1] = f1[H - 1];
    g[H].re = f1[H].re;
    g[H].im = 0;
    for (int k = 1; k < H; k++) { /* Hermitian extension of the packed half spectrum */
      g[N - k].re = g[k].re;
      g[N - k].im = -g[k].im;
    }
    /* do_ifft: negative exponent, unscaled; rescale by 1/N (FFTReal.hpp:206-293) */
    fft_ref_inverse(g, N);
    /* rotation by H (cc:500-505) and float accumulation over channels in order A,C,G,T (cc:395-403) */
    for (int i = 0; i < N; i++) {
      float x = (float)(g[(i + H) % N].re / (double)N);
      out[i] += x;
    }
  }
  free(f1);
  free(f2);
  free(g);
}

/* ------------------------------------------------------------------------------------------
 * d1. FindTop (CrossCorr.cc:878-944): RMS envelope per 256 lags, threshold env*cutoff + 1.
 * ------------------------------------------------------------------------------------------ */
int sxo_findtop(const float *xc, int N, double cutoff, int32_t *idx, int cap, double *env_out) {
  const int envSize = 256;
  int nPoints = N / envSize;
  if (nPoints == 0) nPoints = 1;
  double *env = (double *)calloc((size_t)nPoints, sizeof(double));
  for (int i = 0; i < N; i++) {
    float sq = xc[i] * xc[i]; /* float product, then widened (cc:902) */
    if (nPoints > 8) env[i / envSize] += (double)sq;
  }
  for (int b = 0; b < nPoints; b++) env[b] = sqrt(env[b] / (double)envSize);
  int n = 0;
  for (int i = 0; i < N; i++) {
    if ((double)xc[i] > env[i / envSize] * cutoff + 1.) {
      if (n < cap) idx[n] = i;
      n++;
    }
  }
  if (env_out) memcpy(env_out, env, sizeof(double) * nPoints);
  free(env);
  return n;
}

/* ------------------------------------------------------------------------------------------
 * e2. Diagonal scan (SeqAnalyzer::DoOne, CrossCorr.cc:667-724; constants cc:557-563).
 * ------------------------------------------------------------------------------------------ */
int sxo_diag(const char *q, int qlen, const char *t, int tlen, int shift, sxo_seg *out, int cap) {
  ensure_tables();
  const int minLen = 45;
  const int threshold = (int)((double)minLen * 0.42 * (double)100); /* cc:675 */
  int sum = 0, seen = 0, open = -1, nout = 0;
  for (int i = 0; i < tlen; i++) {
    int j = i + shift;
    if (j < 0) continue;
    if (j >= qlen || i + 1 >= tlen) {
      if (open != -1) {
        if (nout < cap) {
          out[nout].start_target = open;
          out[nout].start_query = open + shift;
          out[nout].len = i - open;
        }
        nout++;
      }
      break;
    }
    sum += g_score[(unsigned char)t[i]][(unsigned char)q[j]];
    if (seen > minLen) {
      sum -= g_score[(unsigned char)t[i - minLen - 1]][(unsigned char)q[j - minLen - 1]];
      if (sum > threshold) {
        if (open == -1) open = i - minLen;
      } else {
        if (open != -1) {
          if (nout < cap) {
            out[nout].start_target = open;
            out[nout].start_query = open + shift;
            out[nout].len = i - open;
          }
          nout++;
        }
        open = -1;
      }
    }
    seen++;
  }
  return nout;
}

int sxo_matchup(const char *q, int qlen, const char *t, int tlen, const float *xc, int N, double cutoff,
                sxo_seg *out, int cap) {
  int32_t *idx = (int32_t *)malloc(sizeof(int32_t) * (size_t)N);
  int nc = sxo_findtop(xc, N, cutoff, idx, N, NULL);
  int n = 0;
  for (int c = 0; c < nc; c++) {
    int shift = idx[c] - N / 2; /* cc:600-602 */
    int room = cap - n > 0 ? cap - n : 0;
    n += sxo_diag(q, qlen, t, tlen, shift, out + (n < cap ? n : cap), room);
  }
  free(idx);
  return n;
}

/* ------------------------------------------------------------------------------------------
 * e4. Match probability (AlignProbability.cc:11-40, 62-127).
 * ------------------------------------------------------------------------------------------ */
static double prob_from_counts(int len, double ident, double p_match, double target_size) {
  double s = sqrt(p_match * (1. - p_match) * (double)len); /* Sigma */
  double m = p_match * (double)len;
  double x = (double)len * ident;
  double cdf = 0.5 * (1. + erf((m - x) / s / 1.414213562)); /* CDF(m, x, s) */
  double expect = cdf * target_size;
  return exp(-expect);
}

static void seg_counts(const char *t, const char *q, int startT, int startQ, int len, double *matches,
                       double *gcT, double *gcQ) {
  double m = 0, gt = 0, gq = 0;
  for (int i = 0; i < len; i++) {
    int a = (unsigned char)t[i + startT], b = (unsigned char)q[i + startQ];
    m += g_equal[a][b];
    gt += g_frac[a][1] + g_frac[a][2];
    gq += g_frac[b][1] + g_frac[b][2];
  }
  *matches = m;
  *gcT = gt;
  *gcQ = gq;
}

static double gc_adjust_expect(double gc, int n, double gc_target) { /* AlignProbability.cc:26-40 */
  double at_target = 1. - gc_target;
  double r = gc * gc_target;
  r += ((double)n - gc) * at_target;
  return r / (double)n / 2.;
}

double sxo_match_prob(const char *t, const char *q, int startT, int startQ, int len, double target_size,
                      double *ident_out) {
  ensure_tables();
  double matches, gcT, gcQ;
  seg_counts(t, q, startT, startQ, len, &matches, &gcT, &gcQ);
  double ident = matches / (double)len;
  double p_match = gc_adjust_expect(gcQ, len, gcT / (double)len);
  if (ident_out) *ident_out = ident;
  return prob_from_counts(len, ident, p_match, target_size);
}

/* ------------------------------------------------------------------------------------------
 * e5. ProbTable (ProbTable.cc:15-56): for p_match = i/511 (i=1..511) and len = 1..2047 the
 * smallest identity (bisection to 1e-8) whose probability is NON-ZERO (quirk Q11).
 * ------------------------------------------------------------------------------------------ */
void sxo_prob_table_build(double target_size, double *table) {
  const int rows = 512, maxLen = 2048;
  for (int j = 0; j < maxLen; j++) table[j] = 0.; /* row 0 is never filled (Q12) */
  for (int i = 1; i < rows; i++) {
    double *row = table + (size_t)i * maxLen;
    double ident_expect = (double)i / ((double)rows - 1);
    row[0] = 2.;
    for (int j = 1; j < maxLen; j++) {
      double lo = 0, hi = 1;
      while (hi - lo > 0.00000001) {
        double mid = (hi + lo) / 2.0;
        if (prob_from_counts(j, mid, ident_expect, target_size) != 0.)
          hi = mid;
        else
          lo = mid;
      }
      row[j] = (hi + lo) / 2.0;
    }
  }
}

double sxo_prob_table_lookup(const double *table, double table_value, const char *t, const char *q,
                             int startT, int startQ, int len, double *ident_out) {
  ensure_tables();
  const int rows = 512, maxLen = 2048;
  double matches, gcT, gcQ;
  seg_counts(t, q, startT, startQ, len, &matches, &gcT, &gcQ);
  double ident = matches / (double)len;
  double p_match = gc_adjust_expect(gcQ, len, gcT / (double)len);
  if (ident_out) *ident_out = ident;
  int index = (int)(p_match * (double)(rows - 1)); /* ExpectToIndex, ProbTable.cc:71-74 */
  if (index < 1 || index >= rows) return 0.;       /* reference reads out of bounds here (Q12): reject */
  int l = len >= maxLen ? maxLen - 1 : len;
  return ident >= table[(size_t)index * maxLen + l] ? table_value : 0.;
}

/* ------------------------------------------------------------------------------------------
 * Full path for one chunk pair, both strands: HomologyByXCorr::Align + FilterMatches
 * (analysis/HomologyByXCorrSlave.cc:168-253) with create_signals (254-268).
 * ------------------------------------------------------------------------------------------ */
static long filter_matches(const sxo_params *p, const sxo_chunk *t, const sxo_chunk *q, const char *qseq,
                           const sxo_seg *segs, int nseg, int reverse, sxo_result *out, long cap, long n) {
  for (int s = 0; s < nseg; s++) {
    int len = segs[s].len;
    if (len < p->min_len) continue;
    int tStart = t->start + segs[s].start_target;
    int qStart;
    if (!reverse)
      qStart = q->start + segs[s].start_query;
    else /* RCQuery (Slave.cc:56-60) with the -q_chunk FLAG value (quirk Q10) */
      qStart = segs[s].start_query + q->seq_size - q->start - p->q_chunk;
    double prob, ident;
    if (p->use_prob_table && p->prob_table) {
      prob = sxo_prob_table_lookup(p->prob_table, p->table_value, t->bases, qseq, segs[s].start_target,
                                   segs[s].start_query, len, &ident);
    } else {
      prob = sxo_match_prob(t->bases, qseq, segs[s].start_target, segs[s].start_query, len,
                            p->target_total, &ident);
    }
    if (prob < p->min_prob) continue;
    if (n < cap) {
      sxo_result *r = &out[n];
      memset(r, 0, sizeof(*r));
      r->query_id = (uint64_t)(int64_t)q->seq_id;
      r->target_id = (uint64_t)(int64_t)t->seq_id;
      r->query_size = (uint64_t)(int64_t)q->seq_size;
      r->qstart = (uint64_t)(int64_t)qStart; /* int -> unsigned long, sign-extended like the reference */
      r->tstart = (uint64_t)(int64_t)tStart;
      r->len = (uint64_t)(int64_t)len;
      r->reverse = (uint8_t)reverse;
      r->prob = prob;
      r->ident = ident;
    }
    n++;
  }
  return n;
}

long sxo_align_pair(const sxo_params *p, const sxo_chunk *t, const sxo_chunk *q, int fast, sxo_result *out,
                    long cap) {
  ensure_tables();
  int N = 2 * p->t_chunk;
  double cutoff = fast ? p->cutoff_fast : p->cutoff;
  float *tsig = (float *)malloc(sizeof(float) * 5 * (size_t)N);
  float *qsig = (float *)malloc(sizeof(float) * 5 * (size_t)N);
  float *xc = (float *)malloc(sizeof(float) * (size_t)N);
  char *rc = (char *)malloc((size_t)q->len + 1);
  int segcap = 1 << 16;
  sxo_seg *segs = (sxo_seg *)malloc(sizeof(sxo_seg) * (size_t)segcap);
  long n = 0;
  sxo_encode(t->bases, t->len, N, tsig);
  for (int strand = 0; strand < 2; strand++) {
    const char *qseq = q->bases;
    if (strand) {
      sxo_revcomp(q->bases, q->len, rc);
      qseq = rc;
    }
    sxo_encode(qseq, q->len, N, qsig);
    sxo_xcorr(tsig + N, qsig + N, N, xc);
    int nseg = sxo_matchup(qseq, q->len, t->bases, t->len, xc, N, cutoff, segs, segcap);
    if (nseg > segcap) { /* grow and redo: never truncate */
      segcap = nseg;
      segs = (sxo_seg *)realloc(segs, sizeof(sxo_seg) * (size_t)segcap);
      nseg = sxo_matchup(qseq, q->len, t->bases, t->len, xc, N, cutoff, segs, segcap);
    }
    n = filter_matches(p, t, q, qseq, segs, nseg, strand, out, cap, n);
  }
  free(tsig);
  free(qsig);
  free(xc);
  free(rc);
  free(segs);
  return n;
}

typedef struct {
  const sxo_params *p;
  const sxo_chunk *targets, *queries;
  const int32_t *pairs;
  long n;
  int fast;
  long next;
  pthread_mutex_t lock;
  sxo_result *out;
  long cap, nout;
} mt_job;

static void *mt_worker(void *arg) {
  mt_job *job = (mt_job *)arg;
  long lcap = 4096;
  sxo_result *local = (sxo_result *)malloc(sizeof(sxo_result) * (size_t)lcap);
  for (;;) {
    pthread_mutex_lock(&job->lock);
    long i = job->next++;
    pthread_mutex_unlock(&job->lock);
    if (i >= job->n) break;
    const sxo_chunk *t = &job->targets[job->pairs[2 * i]];
    const sxo_chunk *q = &job->queries[job->pairs[2 * i + 1]];
    long k = sxo_align_pair(job->p, t, q, job->fast, local, lcap);
    if (k > lcap) {
      lcap = k;
      local = (sxo_result *)realloc(local, sizeof(sxo_result) * (size_t)lcap);
      k = sxo_align_pair(job->p, t, q, job->fast, local, lcap);
    }
    pthread_mutex_lock(&job->lock);
    for (long r = 0; r < k; r++) {
      if (job->nout < job->cap) job->out[job->nout] = local[r];
      job->nout++;
    }
    pthread_mutex_unlock(&job->lock);
  }
  free(local);
  return NULL;
}

long sxo_align_pairs_mt(const sxo_params *p, const sxo_chunk *targets, const sxo_chunk *queries,
                        const int32_t *pairs, long n, int fast, int threads, sxo_result *out, long cap) {
  ensure_tables();
  mt_job job;
  memset(&job, 0, sizeof(job));
  job.p = p;
  job.targets = targets;
  job.queries = queries;
  job.pairs = pairs;
  job.n = n;
  job.fast = fast;
  job.out = out;
  job.cap = cap;
  pthread_mutex_init(&job.lock, NULL);
  if (threads < 1) threads = 1;
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
  for (int i = 0; i < threads; i++) pthread_create(&th[i], NULL, mt_worker, &job);
  for (int i = 0; i < threads; i++) pthread_join(th[i], NULL);
  free(th);
  pthread_mutex_destroy(&job.lock);
  return job.nout;
}
