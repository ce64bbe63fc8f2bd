// HomologyByXCorr on B200: drop-in for the reference's standalone cross-correlation tool
// (tools/analysis/HomologyByXCorr.cc:505-860): same flags for the supported modes, same chunking
// (target overlap t_chunk/2, query overlap 0, -nblocks/-block sharding), same filter (-min_prob
// applied, RC coordinate from the real chunk length) and the same version-3 binary match file
// (`-o`), so MergeXCorrMatches / ChainMatches consume the output unchanged.
// Not supported here: -guide, -proteins, -select, -pairs, -line, -chain (not on the DNA hot path).
// Known, documented difference: the reference tool reuses one CCSignal object, so a short chunk
// that follows a full one keeps stale samples; this tool uses fresh signals as the
// slave does.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>

#include "sx_host.h"

using namespace sxh;

static const char *flag(std::map<std::string, std::string> &a, const char *k, const char *def) {
  auto it = a.find(k);
  return it == a.end() ? def : it->second.c_str();
}

int main(int argc, char **argv) {
  std::map<std::string, std::string> a;
  for (int i = 1; i + 1 < argc; i += 2) a[argv[i]] = argv[i + 1];
  for (const char *bad : {"-guide", "-proteins", "-select", "-pairs", "-line", "-chain"})
    if (a.count(bad) && std::string(a[bad]) != "" && std::string(a[bad]) != "0" && std::string(a[bad]) != "false") {
      fprintf(stderr, "HomologyByXCorr(B200): %s is not supported by the GPU path\n", bad);
      return 2;
    }
  const std::string q = flag(a, "-q", ""), t = flag(a, "-t", ""), out = flag(a, "-o", "");
  if (q.empty() || t.empty() || (out.empty() && !a.count("-dump_chunks"))) {
    fprintf(stderr,
            "usage: %s -q <query fasta> -t <target fasta> -o <match file> [-l 0] [-q_chunk 4096] [-t_chunk 4096]\n"
            "          [-min_prob 0.9999] [-cutoff 1.8] [-nblocks 0 -block 0] [-nblocks_query 0 -block_query 0]\n"
            "          [-same_only 0] [-device 0] [-dump_chunks 1]\n",
            argv[0]);
    return 2;
  }
  HomologyByXCorr::Options opt;
  opt.device = atoi(flag(a, "-device", "0"));
  opt.t_chunk = atoi(flag(a, "-t_chunk", "4096"));
  opt.q_chunk = atoi(flag(a, "-q_chunk", "4096"));
  opt.cutoff = atof(flag(a, "-cutoff", "1.8"));
  opt.min_len = atoi(flag(a, "-l", "0"));
  opt.min_prob_flag = atof(flag(a, "-min_prob", "0.9999"));
  opt.standalone_semantics = true;
  opt.sort_results = true;
  const int nblocks = atoi(flag(a, "-nblocks", "0")), block = atoi(flag(a, "-block", "0"));
  const int nblocks_q = atoi(flag(a, "-nblocks_query", "0")), block_q = atoi(flag(a, "-block_query", "0"));
  const bool same_only = atoi(flag(a, "-same_only", "0")) != 0;

  std::vector<Sequence> qs, ts;
  std::string err;
  if (!read_fasta(q, qs, &err) || !read_fasta(t, ts, &err)) {
    fprintf(stderr, "%s\n", err.c_str());
    return 1;
  }
  ChunkList tc, qc;
  chunk_sequences(ts, opt.t_chunk, opt.t_chunk / 2, nblocks, block, tc);  // tools/...:635
  chunk_sequences(qs, opt.q_chunk, 0, nblocks_q, block_q, qc);
  printf("Query sequence:  %s\nTarget sequence: %s\n", q.c_str(), t.c_str());
  printf("chunks: target %d query %d\n", tc.n(), qc.n());
  if (a.count("-dump_chunks")) {
    for (int which = 0; which < 2; which++) {
      const ChunkList &c = which ? qc : tc;
      for (int i = 0; i < c.n(); i++)
        printf("%s %d seq=%d start=%d len=%d\n", which ? "Q" : "T", i, c.seq_ids[i], c.starts[i], c.lens[i]);
    }
    return 0;
  }

  HomologyByXCorr hx;
  if (!hx.init(opt, tc, qc)) {
    fprintf(stderr, "HomologyByXCorr(B200): %s\n", hx.error().c_str());
    return 1;
  }
  printf("Keeping alignments more like real than %g\n", opt.min_prob_flag);

  MatchFile mf;
  mf.target_names = tc.names;
  mf.query_names = qc.names;
  mf.target_sizes = tc.seq_sizes;
  mf.query_sizes = qc.seq_sizes;
  // target-major like the reference's main loop (tools/...:733-831): one block per target chunk,
  // every query chunk against it, both orientations
  std::vector<t_pair> blocks;
  auto flush = [&]() -> bool {
    if (blocks.empty()) return true;
    const bool ok = hx.align_targets(blocks.data(), (int)blocks.size(), mf.matches);
    blocks.clear();
    return ok;
  };
  for (int j = 0; j < tc.n(); j++) {
    if (tc.lens[j] == 0 || qc.n() == 0) continue;
    if (!same_only) {
      t_pair p;
      memset(&p, 0, sizeof(p));
      p.target_from = p.target_to = j;
      p.query_from = 0;
      p.query_to = qc.n() - 1;
      blocks.push_back(p);
    } else {
      for (int i = 0; i < qc.n(); i++) {
        if (tc.names[tc.seq_ids[j]] != qc.names[qc.seq_ids[i]]) continue;
        t_pair p;
        memset(&p, 0, sizeof(p));
        p.target_from = p.target_to = j;
        p.query_from = p.query_to = i;
        blocks.push_back(p);
      }
    }
    if (blocks.size() >= 64 && !flush()) break;
  }
  if (!flush()) {
    fprintf(stderr, "HomologyByXCorr(B200): %s\n", hx.error().c_str());
    return 1;
  }
  printf("MultiMatches dump: %zu matches\n", mf.matches.size());
  if (!mf.write(out, &err)) {
    fprintf(stderr, "%s\n", err.c_str());
    return 1;
  }
  sx_stats st;
  if (hx.stats(&st))
    printf("chunk pairs: %lld  candidates: %lld  segments: %lld  matches: %lld  kernel launches: %lld\n",
           (long long)st.chunk_pairs, (long long)st.candidates, (long long)st.segments, (long long)st.matches,
           (long long)st.kernel_launches);
  return 0;
}
