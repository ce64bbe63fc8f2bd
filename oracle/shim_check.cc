// TEST INFRASTRUCTURE: the class-level shims (satsuma2_b200/host/binding/crosscorr_shim.h) next to the reference's own
// classes, compiled against the reference headers where they lie (oracle/make_refslave_b200.py builds this too).
//   shim_check <target fasta> <query fasta>   (first record of each; sequences cut to 4096 bases)
// For the pair and for the reverse-complemented query: signals bit-equal, correlation within 1e-4 of max|xc|,
// MatchUp on the REFERENCE's correlation vector: identical segment lists.  Prints "SHIM OK ..." and exits 0.
#include <cmath>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "analysis/CrossCorr.h"
#include "crosscorr_shim.h"

static std::string first_record(const char *path) {
  std::ifstream in(path);
  std::string line, seq;
  bool started = false;
  while (std::getline(in, line)) {
    if (!line.empty() && line[0] == '>') {
      if (started) break;
      started = true;
    } else {
      seq += line;
    }
  }
  return seq.substr(0, 4096);
}

int main(int argc, char **argv) {
  if (argc != 3) return 2;
  std::streambuf *old = std::cout.rdbuf();
  std::ostringstream sink;
  std::cout.rdbuf(sink.rdbuf());  // the reference is chatty
  DNAVector t, q;
  const std::string ts = first_record(argv[1]), qs = first_record(argv[2]);
  t.SetFromBases(ts);
  q.SetFromBases(qs);
  const int N = 8192;
  int total_segs = 0;
  double worst = 0;
  for (int strand = 0; strand < 2; strand++) {
    if (strand) q.ReverseComplement();
    ::CCSignal rt, rq;
    rt.SetSequence(t, N);
    rq.SetSequence(q, N);
    sx_shim::CCSignal st, sq;
    st.SetSequence(t, N);
    sq.SetSequence(q, N);
    for (int ch = 0; ch < 4; ch++)
      for (int i = 0; i < N; i++)
        if (rt.Get(ch)[i] != st.Get(ch)[i] || rq.Get(ch)[i] != sq.Get(ch)[i]) {
          fprintf(stderr, "signal differs: strand %d channel %d sample %d\n", strand, ch, i);
          return 1;
        }
    std::vector<float> rxc, sxc;
    ::CrossCorrelation rc;
    rc.CrossCorrelate(rxc, rt, rq);
    sx_shim::CrossCorrelation sc;
    sc.CrossCorrelate(sxc, st, sq);
    double mx = 0, err = 0;
    for (int i = 0; i < N; i++) {
      mx = std::max(mx, (double)std::fabs(rxc[i]));
      err = std::max(err, (double)std::fabs(rxc[i] - sxc[i]));
    }
    worst = std::max(worst, err / mx);
    if (err > 1e-4 * mx) {
      fprintf(stderr, "correlation differs: %g of max\n", err / mx);
      return 1;
    }
    vecSeqMatch rm, sm;
    ::SeqAnalyzer ra;
    ra.SetTopCutoff(1.8);
    ra.MatchUp(rm, q, t, rxc);
    sx_shim::SeqAnalyzer sa;
    sa.SetTopCutoff(1.8);
    sa.MatchUp(sm, q, t, rxc);  // the same (reference) correlation vector on both sides
    if (rm.size() != sm.size()) {
      fprintf(stderr, "segment counts differ: %d vs %d\n", rm.size(), sm.size());
      return 1;
    }
    for (int i = 0; i < rm.size(); i++)
      if (rm[i].GetStartTarget() != sm[i].GetStartTarget() || rm[i].GetStartQuery() != sm[i].GetStartQuery() ||
          rm[i].GetLength() != sm[i].GetLength()) {
        fprintf(stderr, "segment %d differs\n", i);
        return 1;
      }
    total_segs += rm.size();
  }
  std::cout.rdbuf(old);
  printf("SHIM OK signals bit-equal, xc rel err %.2e, %d segments identical\n", worst, total_segs);
  return 0;
}
