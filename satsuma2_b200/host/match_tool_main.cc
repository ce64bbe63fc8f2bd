// XCorrMatchTool: the master-side list operations on an xcorr match file (version 3) that become the next
// bottleneck once the slaves run on GPUs: MultiMatches::Sort and MultiMatches::Collapse
// (analysis/SequenceMatch.h:211-215, SequenceMatch.cc:418-469) as SatsumaSynteny2 and ChainMatches apply them
// (analysis/SatsumaSynteny2.cc:468-476, 506-507, 605-606; tools/analysis/ChainMatches.cc:65-66).
// and the synteny chain of ChainMatches (tools/analysis/ChainMatches.cc:60-75 = Sort, Collapse, RunMatchDynProg,
// analysis/MatchDynProg.cc:401-561; `-dups 1` = RunMatchDynProgMult, :245-399).
//   XCorrMatchTool -i <in> -o <out> [-sort 1] [-collapse 1] [-chain 0] [-dups 0]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <string>

#include "sx_host.h"

int main(int argc, char **argv) {
  std::map<std::string, std::string> a;
  for (int i = 1; i + 1 < argc; i += 2) a[argv[i]] = argv[i + 1];
  if (!a.count("-i") || !a.count("-o")) {
    fprintf(stderr, "usage: %s -i <match file> -o <match file> [-sort 1] [-collapse 1] [-chain 0] [-dups 0]\n", argv[0]);
    return 2;
  }
  const bool do_sort = !a.count("-sort") || atoi(a["-sort"].c_str()) != 0;
  const bool do_collapse = !a.count("-collapse") || atoi(a["-collapse"].c_str()) != 0;
  sxh::MatchFile mf;
  std::string err;
  if (!mf.read(a["-i"], &err)) {
    fprintf(stderr, "%s\n", err.c_str());
    return 1;
  }
  printf("Matches read: %zu\n", mf.matches.size());
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto secs = [](std::chrono::steady_clock::time_point a0, std::chrono::steady_clock::time_point a1) {
    return std::chrono::duration<double>(a1 - a0).count();
  };
  auto t0 = now();
  if (do_sort) mf.sort();
  auto t1 = now();
  if (do_collapse) {
    printf("Matches before collapse: %zu\n", mf.matches.size());
    mf.collapse();
    printf("Matches after collapse:  %zu\n", mf.matches.size());
  }
  auto t2 = now();
  if (a.count("-chain") && atoi(a["-chain"].c_str()) != 0) {
    sxh::MatchFile chained;
    if (a.count("-dups") && atoi(a["-dups"].c_str()) != 0)
      mf.chain_dups(chained);
    else
      mf.chain(chained);
    printf("Matches in the chain:    %zu\n", chained.matches.size());
    mf = chained;
  }
  auto t3 = now();
  printf("seconds: sort %.4f collapse %.4f chain %.4f\n", secs(t0, t1), secs(t1, t2), secs(t2, t3));
  if (!mf.write(a["-o"], &err)) {
    fprintf(stderr, "%s\n", err.c_str());
    return 1;
  }
  return 0;
}
