// HomologyByXCorr on B200: drop-in for the reference's standalone cross-correlation tool
// (tools/analysis/HomologyByXCorr.cc:505-860): same flags for the supported modes, same chunking
// (target overlap t_chunk/2, query overlap 0, -nblocks/-block sharding), same filter (-min_prob
// applied, RC coordinate from the real chunk length) and the same version-3 binary match file
// (`-o`), so MergeXCorrMatches / ChainMatches consume the output unchanged.
// `-guide <chained match file>` is the refinement pass of SatsumaSynteny2 -do_refine (tools/...:206-330, 651-675,
// 714-717, 786-790, 833-837): pieces of the gaps between chained matches, piece i against pieces i-3 .. i+3, one
// forced orientation per piece, targetSize = t_chunk, then the guide's matches merged in and the list sorted.
// Not supported here: -proteins, -select, -pairs, -line, -chain (not on the DNA hot path).
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
  for (const char *bad : {"-proteins", "-select", "-pairs", "-line", "-chain"})
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
  const std::string guide = flag(a, "-guide", "");
  if (!guide.empty() && nblocks != 0) {
    fprintf(stderr, "Guided mode, chunking not available!!\n");  // tools/...:605-608
    return 2;
  }

  std::vector<Sequence> qs, ts;
  std::string err;
  if (!read_fasta(q, qs, &err) || !read_fasta(t, ts, &err)) {
    fprintf(stderr, "%s\n", err.c_str());
    return 1;
  }
  ChunkList tc, qc;
  chunk_sequences(ts, opt.t_chunk, opt.t_chunk / 2, nblocks, block, tc);  // tools/...:635
  chunk_sequences(qs, opt.q_chunk, 0, nblocks_q, block_q, qc);
  std::vector<int> orientation;  // guided mode: per piece +1 forward only, -1 reverse only
  MatchFile chained;
  if (!guide.empty()) {
    printf("Reading %s\n", guide.c_str());
    if (!chained.read(guide, &err)) {
      fprintf(stderr, "%s\n", err.c_str());
      return 1;
    }
    printf("Done loading, recomputing chunks.\n");
    guide_chunks(ts, qs, chained, opt.t_chunk, tc, qc, orientation);
    opt.target_total = (double)opt.t_chunk;
    printf("Using target size (guided) %d\n", opt.t_chunk);
  }
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
  if (!guide.empty()) {
    // piece j of the target against pieces j-3 .. j+3 of the query (bOneOnOne); a run of query pieces with the same
    // forced orientation is one block, and only records of that orientation are kept
    for (int pass = 0; pass < 2 && tc.n() > 0; pass++) {  // pass 0: forward-only pieces, pass 1: reverse-only
      const int want = pass == 0 ? 1 : -1;
      for (int j = 0; j < tc.n(); j++) {
        if (tc.lens[j] == 0) continue;
        const int lo = j - 3 < 0 ? 0 : j - 3, hi = j + 3 >= qc.n() ? qc.n() - 1 : j + 3;
        for (int i = lo; i <= hi;) {
          if (orientation[i] != want || qc.lens[i] == 0 || (same_only && tc.names[tc.seq_ids[j]] != qc.names[qc.seq_ids[i]])) {
            i++;
            continue;
          }
          int e = i;
          while (e + 1 <= hi && orientation[e + 1] == want && qc.lens[e + 1] != 0 &&
                 !(same_only && tc.names[tc.seq_ids[j]] != qc.names[qc.seq_ids[e + 1]]))
            e++;
          t_pair p;
          memset(&p, 0, sizeof(p));
          p.target_from = p.target_to = j;
          p.query_from = i;
          p.query_to = e;
          blocks.push_back(p);
          i = e + 1;
        }
      }
      const size_t before = mf.matches.size();
      if (!flush()) {
        fprintf(stderr, "HomologyByXCorr(B200): %s\n", hx.error().c_str());
        return 1;
      }
      size_t w = before;
      for (size_t r = before; r < mf.matches.size(); r++)
        if ((mf.matches[r].reverse != 0) == (want < 0)) mf.matches[w++] = mf.matches[r];
      mf.matches.resize(w);
    }
    // "Merging w/ guide..." (tools/...:833-837): MultiMatches::MergeRead takes names and sizes from the guide file
    // and appends its matches; then the list is sorted
    printf("Merging w/ guide...\n");
    mf.n_matches.clear();
    for (const t_result &m : mf.matches) mf.n_matches.push_back(m.ident * (double)(int32_t)m.len);
    mf.target_names = chained.target_names;
    mf.query_names = chained.query_names;
    mf.target_sizes = chained.target_sizes;
    mf.query_sizes = chained.query_sizes;
    for (size_t r = 0; r < chained.matches.size(); r++) {
      mf.matches.push_back(chained.matches[r]);
      mf.n_matches.push_back(r < chained.n_matches.size() ? chained.n_matches[r]
                                                          : chained.matches[r].ident * (double)(int32_t)chained.matches[r].len);
    }
    mf.sort();
  }
  for (int j = 0; guide.empty() && j < tc.n(); j++) {
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
    if (blocks.size() >= 64 && !flush()) {  // a GPU failure must not end in a truncated match file and exit code 0
      fprintf(stderr, "HomologyByXCorr(B200): %s\n", hx.error().c_str());
      return 1;
    }
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
