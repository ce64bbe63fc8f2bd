// TEST INFRASTRUCTURE ONLY -- never linked into or called from the product path.
//
// Thin extern "C" harness around the UNMODIFIED Satsuma2 reference, compiled from
// the sources where they lie under /root/reference (see oracle/Makefile, target
// `ref`).  Nothing from the reference is copied into this repository: this file
// only #includes the reference translation unit that holds the hot-path driver
// (analysis/HomologyByXCorrSlave.cc, with its main() renamed) so that the class
// HomologyByXCorr (Align / FilterMatches / align_target, Slave.cc:168-300) and
// the globals it works on (Slave.cc:28-53) are reachable, and exposes
//   * per-stage taps  (CCSignal, CrossCorrelate, FindTop, MatchUp, probability)
//   * the block driver (align_target) single- and multi-threaded
// The output (oracle/_ref/libsatsuma_ref.so) is used to (1) pin the C restatement
// in oracle/sx_oracle.c, (2) generate tests/golden/*, (3) time the reference's own
// CPU implementation in `bench.py --impl reference`.

#include <iostream>
#include <map>
#include <fstream>
#include <string.h>
#include <algorithm>
#include <sstream>
#include <vector>
#include <string>
#include <queue>
#include <thread>
#include <mutex>
#include <atomic>
#include <memory>
#include <cstdint>
#include <set>
#include <list>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <chrono>

// Reach private taps (FindTop, m_entropy, ChunkManager::m_lengths) without editing
// the reference.  Must come after the std headers (libstdc++ breaks otherwise).
#define private public
#define protected public
#define main satsuma_slave_main_unused
#include "analysis/HomologyByXCorrSlave.cc"
#undef main
#undef private
#undef protected
#include "analysis/MatchDynProg.h"

namespace {

// cout is extremely chatty in the reference ("select=..", "chunks: ..", "worker created");
// silence it for the duration of a call.
struct CoutSilencer {
  std::streambuf *old;
  std::ostringstream sink;
  CoutSilencer() { old = std::cout.rdbuf(sink.rdbuf()); }
  ~CoutSilencer() { std::cout.rdbuf(old); }
};

void fill_dna(DNAVector &d, const char *bases, int len) {
  d.resize(len);
  for (int i = 0; i < len; i++) d[i] = bases[i];
}

}  // namespace

extern "C" {

struct ref_result {  // mirrors t_result (analysis/WorkQueue.h:23-33), 72 bytes
  unsigned long query_id, target_id, query_size, qstart, tstart, len;
  unsigned char reverse;
  unsigned char pad[7];
  double prob, ident;
};

int ref_sizeof_t_result() { return (int)sizeof(t_result); }
int ref_sizeof_t_pair() { return (int)sizeof(t_pair); }

// ---------------------------------------------------------------- configuration
// Mirrors the flag parsing in Slave main() (Slave.cc:370-384).
void ref_configure(int t_chunk, int q_chunk, double cutoff, double cutoff_fast, int min_len,
                   int use_prob_table, double min_prob_flag) {
  targetChunk = t_chunk;
  queryChunk = q_chunk;
  topCutoff = cutoff;
  topCutoffFast = cutoff_fast;
  minLen = min_len;
  prob_table = use_prob_table != 0;
  minProb = min_prob_flag;
}

// Builds the ProbTable exactly as Slave.cc:413-415 does. Call after chunks are set.
void ref_build_prob_table() {
  CoutSilencer s;
  probt = ProbTable(targetTotal, minProb);
}

double ref_target_total() { return targetTotal; }
void ref_set_target_total(double t) { targetTotal = t; }

// ---------------------------------------------------------------- chunk loading
// (1) the reference's own loader + chunker, exactly as Slave.cc:388-408.
int ref_load_fasta(const char *target_fasta, const char *query_fasta) {
  CoutSilencer s;
  targetRaw.clear(); queryRaw.clear(); targetNames.clear(); queryNames.clear();
  target.clear(); query.clear(); targetInfo.clear(); queryInfo.clear();
  queryRaw.Read(query_fasta, queryNames);
  delete cmQuery;
  cmQuery = new ChunkManager(queryChunk, 0);
  cmQuery->ChunkIt(query, queryInfo, queryRaw, queryNames, 0, 0);
  queryRaw.clear();
  targetRaw.Read(target_fasta, targetNames);
  delete cmTarget;
  cmTarget = new ChunkManager(targetChunk, targetChunk / 4);
  cmTarget->ChunkIt(target, targetInfo, targetRaw, targetNames, 0, 0);
  targetRaw.clear();
  targetTotal = 0;
  for (int i = 0; i < cmTarget->GetCount(); i++) targetTotal += (double)cmTarget->GetSize(i);
  return 0;
}

int ref_num_chunks(int is_target) { return is_target ? (int)target.size() : (int)query.size(); }
int ref_num_seqs(int is_target) { return is_target ? cmTarget->GetCount() : cmQuery->GetCount(); }
int ref_seq_size(int is_target, int id) { return is_target ? cmTarget->GetSize(id) : cmQuery->GetSize(id); }
int ref_chunk_len(int is_target, int i) { return is_target ? target[i].size() : query[i].size(); }
int ref_chunk_start(int is_target, int i) { return is_target ? targetInfo[i].GetStart() : queryInfo[i].GetStart(); }
int ref_chunk_seq(int is_target, int i) { return is_target ? targetInfo[i].GetID() : queryInfo[i].GetID(); }
void ref_chunk_bases(int is_target, int i, char *out) {
  const DNAVector &d = is_target ? target[i] : query[i];
  for (int k = 0; k < d.size(); k++) out[k] = d[k];
}

// (2) direct chunk injection (synthetic workloads; bypasses FASTA + ChunkManager):
// chunk i = bases[offsets[i] .. offsets[i]+lens[i]), belongs to sequence seq_ids[i] at
// coordinate starts[i]; seq_sizes[n_seqs] are the full sequence lengths
// (ChunkManager::m_lengths).  target_total as Slave.cc:405-408 = sum of target seq sizes.
void ref_set_chunks(int is_target, int n, const char *bases, const long *offsets, const int *lens,
                    const int *seq_ids, const int *starts, int n_seqs, const int *seq_sizes) {
  CoutSilencer s;
  vecDNAVector &v = is_target ? target : query;
  std::vector<SeqChunk> &info = is_target ? targetInfo : queryInfo;
  v.clear();
  v.resize(n);
  info.clear();
  info.resize(n);
  for (int i = 0; i < n; i++) {
    fill_dna(v[i], bases + offsets[i], lens[i]);
    info[i].Set("seq", starts[i], seq_ids[i]);
  }
  ChunkManager *&cm = is_target ? cmTarget : cmQuery;
  delete cm;
  cm = new ChunkManager(is_target ? targetChunk : queryChunk, is_target ? targetChunk / 4 : 0);
  cm->m_lengths.assign(seq_sizes, seq_sizes + n_seqs);
  if (is_target) {
    targetTotal = 0;
    for (int i = 0; i < n_seqs; i++) targetTotal += (double)seq_sizes[i];
  }
}

// ---------------------------------------------------------------- block driver
// One t_pair through HomologyByXCorr::align_target (Slave.cc:270-300).  Returns the
// number of t_result records produced (all are copied if cap is large enough).
long ref_align_block(int tFrom, int tTo, int qFrom, int qTo, int fast, ref_result *out, long cap) {
  CoutSilencer s;
  t_pair p;
  memset(&p, 0, sizeof(p));
  p.targetFrom = tFrom; p.targetTo = tTo; p.queryFrom = qFrom; p.queryTo = qTo; p.fast = fast != 0;
  results.clear();
  HomologyByXCorr h;
  h.align_target(p);
  long n = (long)results.size();
  for (long i = 0; i < n && i < cap; i++) memcpy(&out[i], &results[i], sizeof(t_result));
  results.clear();
  return n;
}

// Many t_pairs on `threads` worker threads, each with its own HomologyByXCorr object
// popping from a shared queue -- the structure of launch_worker/work (Slave.cc:302-325)
// without the sleep(1) polling and the TCP loop.  Returns number of records; wall time
// (seconds, align work only) in *seconds.
long ref_align_pairs_mt(const int *tpairs /* n x 5: tFrom,tTo,qFrom,qTo,fast */, long n, int threads,
                        ref_result *out, long cap, double *seconds) {
  CoutSilencer s;
  results.clear();
  std::atomic<long> next(0);
  auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> pool;
  for (int w = 0; w < threads; w++) {
    pool.emplace_back([&]() {
      HomologyByXCorr h;
      for (;;) {
        long i = next.fetch_add(1);
        if (i >= n) break;
        t_pair p;
        memset(&p, 0, sizeof(p));
        p.targetFrom = tpairs[5 * i + 0]; p.targetTo = tpairs[5 * i + 1];
        p.queryFrom = tpairs[5 * i + 2]; p.queryTo = tpairs[5 * i + 3];
        p.fast = tpairs[5 * i + 4] != 0;
        h.align_target(p);
      }
    });
  }
  for (auto &t : pool) t.join();
  auto t1 = std::chrono::steady_clock::now();
  if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
  long nres = (long)results.size();
  for (long i = 0; i < nres && i < cap; i++) memcpy(&out[i], &results[i], sizeof(t_result));
  results.clear();
  return nres;
}

// ---------------------------------------------------------------- stage taps
// a1-a3: CCSignal::SetSequence(seq, size) -> entropy[size], A,C,G,T[size]  (CrossCorr.cc:179-206)
void ref_signal(const char *bases, int len, int size, float *out5 /* 5*size: ent,A,C,G,T */) {
  DNAVector d;
  fill_dna(d, bases, len);
  CCSignal sig;
  sig.SetSequence(d, size);
  memcpy(out5, sig.m_entropy.data(), sizeof(float) * size);
  for (int c = 0; c < 4; c++) memcpy(out5 + (size_t)(c + 1) * size, sig.Get(c).data(), sizeof(float) * size);
}

// DNAVector::ReverseComplement (DNAVector.cc:482-521)
void ref_revcomp(const char *bases, int len, char *out) {
  DNAVector d;
  fill_dna(d, bases, len);
  d.ReverseComplement();
  for (int i = 0; i < len; i++) out[i] = d[i];
}

// c1+c2: CrossCorrelation::CrossCorrelate(out, target, query)  (CrossCorr.cc:386-403)
void ref_xcorr(const char *t, int tlen, const char *q, int qlen, int size, float *out) {
  DNAVector dt, dq;
  fill_dna(dt, t, tlen);
  fill_dna(dq, q, qlen);
  CCSignal st, sq;
  st.SetSequence(dt, size);
  sq.SetSequence(dq, size);
  CrossCorrelation xc;
  std::vector<float> res;
  xc.CrossCorrelate(res, st, sq);
  memcpy(out, res.data(), sizeof(float) * size);
}

// b2: one forward transform through FFTReal<float>::do_fft (packed layout, readme.txt:140-160)
void ref_fft(const float *x, int n, float *f) {
  FFTReal<float> fft(n);
  fft.do_fft(f, x);
}

// d1: SeqAnalyzer::FindTop (CrossCorr.cc:878-944). Returns count, indices ascending.
int ref_findtop(const float *xc, int n, double cutoff, int *out, int cap) {
  std::vector<float> v(xc, xc + n);
  std::vector<int> top;
  SeqAnalyzer sa;
  sa.FindTop(top, v, cutoff);
  for (size_t i = 0; i < top.size() && (int)i < cap; i++) out[i] = top[i];
  return (int)top.size();
}

// e1+e2: SeqAnalyzer::MatchUp(DNAVector...) (CrossCorr.cc:583-605, 667-724).
// out: n x 3 ints (startTarget, startQuery, len). Returns count.
int ref_matchup(const char *q, int qlen, const char *t, int tlen, const float *xc, int n, double cutoff,
                int *out, int cap) {
  DNAVector dt, dq;
  fill_dna(dt, t, tlen);
  fill_dna(dq, q, qlen);
  std::vector<float> v(xc, xc + n);
  vecSeqMatch m;
  SeqAnalyzer sa;
  sa.SetTopCutoff(cutoff);
  sa.MatchUp(m, dq, dt, v);
  for (int i = 0; i < m.size() && i < cap; i++) {
    out[3 * i + 0] = m[i].GetStartTarget();
    out[3 * i + 1] = m[i].GetStartQuery();
    out[3 * i + 2] = m[i].GetLength();
  }
  return m.size();
}

// e2 alone: one diagonal through SeqAnalyzer::DoOne (CrossCorr.cc:667-724).
int ref_diag(const char *q, int qlen, const char *t, int tlen, int shift, int *out, int cap) {
  DNAVector dt, dq;
  fill_dna(dt, t, tlen);
  fill_dna(dq, q, qlen);
  vecSeqMatch m;
  SeqAnalyzer sa;
  sa.DoOne(m, dq, dt, shift);
  for (int i = 0; i < m.size() && i < cap; i++) {
    out[3 * i + 0] = m[i].GetStartTarget();
    out[3 * i + 1] = m[i].GetStartQuery();
    out[3 * i + 2] = m[i].GetLength();
  }
  return m.size();
}

// e4: GetMatchProbabilityEx (AlignProbability.cc:62-127)
double ref_match_prob(const char *t, int tlen, const char *q, int qlen, int startT, int startQ, int len,
                      double targetSize, double *ident) {
  DNAVector dt, dq;
  fill_dna(dt, t, tlen);
  fill_dna(dq, q, qlen);
  double id = 0;
  double p = GetMatchProbabilityEx(id, dt, dq, startT, startQ, len, targetSize);
  if (ident) *ident = id;
  return p;
}

// e5: ProbTable built for (targetSize, cutoff) -> raw table rows 0..511 x 2048 (row 0 unfilled = all 0 here)
void ref_prob_table(double targetSize, double cutoff, double *out /* 512*2048 */) {
  ProbTable pt(targetSize, cutoff);
  for (int i = 0; i < 512; i++)
    for (int j = 0; j < 2048; j++)
      out[(size_t)i * 2048 + j] = (i == 0 || pt.m_table[i].empty()) ? 0.0 : pt.m_table[i][j];
}

// e6: PrintMatch(query, target, m, silent) identity (CrossCorr.cc:967-1037)
double ref_ident(const char *q, int qlen, const char *t, int tlen, int startT, int startQ, int len) {
  DNAVector dt, dq;
  fill_dna(dt, t, tlen);
  fill_dna(dq, q, qlen);
  SeqMatch m(startT, startQ, len, 0.);
  return PrintMatch(dq, dt, m, true);
}

// Match-file reader of the reference (MultiMatches::Read, analysis/SequenceMatch.cc:113-249): used to
// prove that files written by the B200 host tools are consumable by MergeXCorrMatches & co.
// out: n x 10 doubles (tID,qID,qLen,startT,startQ,len,rc,matches,prob,ident). Returns n or -1.
long ref_read_match_file(const char *path, double *out, long cap, int *n_targets, int *n_queries) {
  CoutSilencer s;
  MultiMatches mm;
  mm.Read(path);
  if (n_targets) *n_targets = mm.GetTargetCount();
  if (n_queries) *n_queries = mm.GetQueryCount();
  long n = mm.GetMatchCount();
  for (long i = 0; i < n && i < cap; i++) {
    const SingleMatch &m = mm.GetMatch((int)i);
    double *o = out + 10 * i;
    o[0] = m.GetTargetID(); o[1] = m.GetQueryID(); o[2] = m.m_queryLen; o[3] = m.GetStartTarget();
    o[4] = m.GetStartQuery(); o[5] = m.GetLength(); o[6] = m.IsRC() ? 1 : 0; o[7] = m.GetMatches();
    o[8] = m.GetProbability(); o[9] = m.GetIdentity();
  }
  return n;
}

// MultiMatches::Sort + Collapse (analysis/SequenceMatch.h:211-215, SequenceMatch.cc:418-469) on n records of
// 10 doubles (layout of ref_read_match_file); returns the number of records left, written back to `io`.
long ref_sort_collapse(double *io, long n, int do_collapse) {
  CoutSilencer s;
  MultiMatches mm;
  for (long i = 0; i < n; i++) {
    const double *r = io + 10 * i;
    SingleMatch m;
    m.SetQueryTargetID((int)r[1], (int)r[0], (int)r[2]);
    m.SetPos((int)r[4], (int)r[3], (int)r[5], r[6] != 0.);
    m.AddMatches(r[7]);
    m.SetProbability(r[8]);
    m.SetIdentity(r[9]);
    mm.AddMatch(m);
  }
  mm.Sort();
  if (do_collapse && mm.GetMatchCount() > 0) mm.Collapse();
  const long k = mm.GetMatchCount();
  for (long i = 0; i < k; i++) {
    const SingleMatch &m = mm.GetMatch((int)i);
    double *o = io + 10 * i;
    o[0] = m.GetTargetID(); o[1] = m.GetQueryID(); o[2] = m.m_queryLen; o[3] = m.GetStartTarget();
    o[4] = m.GetStartQuery(); o[5] = m.GetLength(); o[6] = m.IsRC() ? 1 : 0; o[7] = m.GetMatches();
    o[8] = m.GetProbability(); o[9] = m.GetIdentity();
  }
  return k;
}

// Timing of the list operations alone (no marshalling): Sort, Collapse, RunMatchDynProg[Mult] on n records.
// secs[0] = Sort, secs[1] = Collapse, secs[2] = chain.  Returns the chain length.
long ref_sort_collapse_chain_timed(const double *in, long n, int n_targets, int n_queries, const int *tsize,
                                   const int *qsize, int dups, double *secs, long *n_collapsed) {
  CoutSilencer s;
  MultiMatches mm, out;
  mm.SetCounts(n_targets, n_queries);
  for (int i = 0; i < n_targets; i++) mm.SetTargetSize(i, tsize[i]);
  for (int i = 0; i < n_queries; i++) mm.SetQuerySize(i, qsize[i]);
  for (long i = 0; i < n; i++) {
    const double *r = in + 10 * i;
    SingleMatch m;
    m.SetQueryTargetID((int)r[1], (int)r[0], (int)r[2]);
    m.SetPos((int)r[4], (int)r[3], (int)r[5], r[6] != 0.);
    m.AddMatches(r[7]);
    m.SetProbability(r[8]);
    m.SetIdentity(r[9]);
    mm.AddMatch(m);
  }
  auto t0 = std::chrono::steady_clock::now();
  mm.Sort();
  auto t1 = std::chrono::steady_clock::now();
  mm.Collapse();
  auto t2 = std::chrono::steady_clock::now();
  if (n_collapsed) *n_collapsed = mm.GetMatchCount();
  if (dups)
    RunMatchDynProgMult(out, mm);
  else
    RunMatchDynProg(out, mm);
  auto t3 = std::chrono::steady_clock::now();
  secs[0] = std::chrono::duration<double>(t1 - t0).count();
  secs[1] = std::chrono::duration<double>(t2 - t1).count();
  secs[2] = std::chrono::duration<double>(t3 - t2).count();
  return out.GetMatchCount();
}

// RunMatchDynProg (analysis/MatchDynProg.cc:401-561) on n records (layout of ref_read_match_file) that are already
// sorted + collapsed; sequence sizes as the match file carries them.  Returns the chain length, records in `io`.
long ref_chain(double *io, long n, int n_targets, int n_queries, const int *tsize, const int *qsize, int dups) {
  CoutSilencer s;
  MultiMatches in, out;
  in.SetCounts(n_targets, n_queries);
  for (int i = 0; i < n_targets; i++) in.SetTargetSize(i, tsize[i]);
  for (int i = 0; i < n_queries; i++) in.SetQuerySize(i, qsize[i]);
  for (long i = 0; i < n; i++) {
    const double *r = io + 10 * i;
    SingleMatch m;
    m.SetQueryTargetID((int)r[1], (int)r[0], (int)r[2]);
    m.SetPos((int)r[4], (int)r[3], (int)r[5], r[6] != 0.);
    m.AddMatches(r[7]);
    m.SetProbability(r[8]);
    m.SetIdentity(r[9]);
    in.AddMatch(m);
  }
  if (dups)
    RunMatchDynProgMult(out, in);  // analysis/MatchDynProg.cc:245-399
  else
    RunMatchDynProg(out, in);
  const long k = out.GetMatchCount();
  for (long i = 0; i < k && i < n; i++) {
    const SingleMatch &m = out.GetMatch((int)i);
    double *o = io + 10 * i;
    o[0] = m.GetTargetID(); o[1] = m.GetQueryID(); o[2] = m.m_queryLen; o[3] = m.GetStartTarget();
    o[4] = m.GetStartQuery(); o[5] = m.GetLength(); o[6] = m.IsRC() ? 1 : 0; o[7] = m.GetMatches();
    o[8] = m.GetProbability(); o[9] = m.GetIdentity();
  }
  return k;
}

// a4: codec tables (DNAVector.cc:13-58, 351-403) for all 256 byte values (as signed char index, i.e.
// what DNA_A(char) sees for bytes < 128; bytes >= 128 are UB in the reference and not tabulated).
void ref_codec(double *acgt /* 128*4 */, char *rc /* 128 */, double *equal /* 128*128 */,
               double *equal_amb /* 128*128 */) {
  for (int i = 0; i < 128; i++) {
    acgt[4 * i + 0] = DNA_A((char)i);
    acgt[4 * i + 1] = DNA_C((char)i);
    acgt[4 * i + 2] = DNA_G((char)i);
    acgt[4 * i + 3] = DNA_T((char)i);
    rc[i] = GetRC((char)i);
    for (int j = 0; j < 128; j++) {
      equal[i * 128 + j] = DNA_Equal((char)i, (char)j);
      equal_amb[i * 128 + j] = DNA_EqualAmb((char)i, (char)j);
    }
  }
}

}  // extern "C"
