/* TEST INFRASTRUCTURE ONLY -- CPU restatement of Satsuma2's chunk-pair cross-correlation
 * hot path (the checker for the CUDA path; never linked into or called by the product).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
 * may load this library.
 *
 * Parity status: PINNED.  Every stage below is checked against the unmodified reference
 * compiled from /root/reference (oracle/_ref/libsatsuma_ref.so, see oracle/Makefile and
 * tests/test_oracle_vs_reference.py) and against tests/golden/ fixtures generated from it
 * (tests/golden/make_golden.py).  Integer/byte stages are bit-exact; the correlation
 * stage is a float64 model of the reference's float32 FFT (measured deviation ~2e-7 of
 * max|xc|, tolerance 1e-4).
 *
 * All file:line citations are relative to /root/reference.
 */
#ifndef SX_ORACLE_H_
#define SX_ORACLE_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* analysis/WorkQueue.h:23-33 (t_result), 72 bytes, native endianness and padding */
typedef struct {
  uint64_t query_id, target_id, query_size, qstart, tstart, len;
  uint8_t reverse;
  uint8_t pad[7];
  double prob, ident;
} sxo_result;

/* analysis/CrossCorr.h:165-196 (SeqMatch) without the unused ident field */
typedef struct {
  int32_t start_target, start_query, len;
} sxo_seg;

/* Parameters of one run: the globals of analysis/HomologyByXCorrSlave.cc:28-53 */
typedef struct {
  int32_t t_chunk;       /* -t_chunk; signal/FFT length N = 2*t_chunk (Slave.cc:259,263) */
  int32_t q_chunk;       /* -q_chunk; used only by the RC coordinate formula (Slave.cc:180) */
  double cutoff;         /* -cutoff (1.8) */
  double cutoff_fast;    /* -cutoff_fast (2.9), used when t_pair.fast (Slave.cc:237) */
  int32_t min_len;       /* -l (0) */
  int32_t use_prob_table;/* -prob_table */
  double min_prob;       /* effective filter threshold: 0.99 in the slave (Slave.cc:76) */
  double table_value;    /* value ProbTable returns for "good" (= -min_prob flag, ProbTable.cc:136) */
  double target_total;   /* sum of all target sequence lengths (Slave.cc:405-408) */
  const double *prob_table; /* 512 x 2048 doubles (row 0 unused) or NULL */
} sxo_params;

/* One chunk as the slave holds it: bases (upper-case ASCII) + SeqChunk info */
typedef struct {
  const char *bases;
  int32_t len;
  int32_t start;    /* SeqChunk::GetStart */
  int32_t seq_id;   /* SeqChunk::GetID */
  int32_t seq_size; /* ChunkManager::GetSize(seq_id) */
} sxo_chunk;

/* ---- a4: codec (analysis/DNAVector.cc:13-58, 351-403, 482-521) */
void sxo_codec(int byte, double acgt[4]);
char sxo_rc_base(int byte);
double sxo_equal(int a, int b);       /* DNA_Equal */
int sxo_score(int a, int b);          /* (int)(100*DNA_EqualAmb + 0.5), CrossCorr.cc:547-553 */
void sxo_revcomp(const char *in, int len, char *out);

/* ---- a1-a3: CCSignal::SetSequence on a fresh object (CrossCorr.cc:35-134, 179-206)
 * out5 = entropy[N], A[N], C[N], G[N], T[N] */
void sxo_encode(const char *bases, int len, int N, float *out5);

/* ---- c1+c2: CrossCorrelation::CrossCorrelate(out, target, query) (CrossCorr.cc:386-507)
 * tsig/qsig: 4 channels x N floats (A,C,G,T). float64 model incl. quirks Q1-Q4. */
void sxo_xcorr(const float *tsig4, const float *qsig4, int N, float *out);

/* ---- d1: SeqAnalyzer::FindTop (CrossCorr.cc:878-944); env (N/256 doubles) optional */
int sxo_findtop(const float *xc, int N, double cutoff, int32_t *idx, int cap, double *env);

/* ---- e2: SeqAnalyzer::DoOne (CrossCorr.cc:667-724) */
int sxo_diag(const char *q, int qlen, const char *t, int tlen, int shift, sxo_seg *out, int cap);
/* ---- e1: SeqAnalyzer::MatchUp (CrossCorr.cc:583-605) */
int sxo_matchup(const char *q, int qlen, const char *t, int tlen, const float *xc, int N, double cutoff,
                sxo_seg *out, int cap);

/* ---- e4: GetMatchProbabilityEx (AlignProbability.cc:62-127) */
double sxo_match_prob(const char *t, const char *q, int startT, int startQ, int len, double target_size,
                      double *ident);
/* ---- e5: ProbTable::Setup (ProbTable.cc:15-56) -> 512x2048, and lookup (58-74, 105-140) */
void sxo_prob_table_build(double target_size, double *table);
double sxo_prob_table_lookup(const double *table, double table_value, const char *t, const char *q,
                             int startT, int startQ, int len, double *ident);

/* ---- full path: HomologyByXCorr::Align + FilterMatches for one chunk pair, both strands
 * (Slave.cc:168-253).  Appends to out (cap entries); returns number produced (may exceed cap). */
long sxo_align_pair(const sxo_params *p, const sxo_chunk *t, const sxo_chunk *q, int fast, sxo_result *out,
                    long cap);

/* Many independent (target index, query index) pairs on `threads` pthreads ("port" CPU baseline). */
long sxo_align_pairs_mt(const sxo_params *p, const sxo_chunk *targets, const sxo_chunk *queries,
                        const int32_t *pairs /* n x 2 */, long n, int fast, int threads, sxo_result *out,
                        long cap);

#ifdef __cplusplus
}
#endif
#endif
