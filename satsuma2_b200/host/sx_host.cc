// Host side above the C ABI -- see sx_host.h.  Behaviour follows the reference files cited there;
// no reference code is reused.
#include "sx_host.h"

#include <cctype>
#include <algorithm>
#include <atomic>
#include <thread>
#include <cmath>
#include <cstdio>
#include <cstring>

namespace sxh {

// ------------------------------------------------------------------------------------------------
// FASTA.  Reference behaviour (analysis/DNAVector.cc:1190-1282, base/FileParser.cc:140-152,
// util/mutil.cc:349-377, 1493-1526): lines end at '\n' only (a '\r' stays in the token); tokens are
// separated by ' ' and '\t'; empty lines are skipped; a line whose first token starts with '>' opens a
// record whose name is all tokens joined by '_' (the '>' is dropped when names are exported);
// otherwise only the FIRST token of the line is sequence; bases seen before the first header are
// kept and end up in front of the first record; everything is upper-cased with toupper.
// ------------------------------------------------------------------------------------------------
static void split_tokens(const std::string &line, std::vector<std::string> &tok) {
  tok.clear();
  std::string cur;
  for (size_t i = 0; i <= line.size(); i++) {
    const bool end = i == line.size();
    if (end || line[i] == ' ' || line[i] == '\t') {
      if (!cur.empty()) tok.push_back(cur);
      cur.clear();
    } else {
      cur.push_back(line[i]);
    }
  }
}

// vecDNAVector::ReadQ (analysis/DNAVector.cc:1150-1173): records of four lines; the name is the WHOLE header
// line (blanks and the leading '@' included), the bases are the whole second line exactly as written, the '+'
// and quality lines are skipped; blank lines between records are skipped.
static bool read_one_fastq(const std::string &file, std::vector<Sequence> &out, std::string *err) {
  FILE *f = fopen(file.c_str(), "rb");
  if (!f) {
    if (err) *err = "cannot open " + file;
    return false;
  }
  std::vector<std::string> lines;
  std::string line;
  char buf[1 << 16];
  while (fgets(buf, sizeof(buf), f)) {
    size_t n = strlen(buf);
    const bool eol = n > 0 && buf[n - 1] == '\n';
    if (eol) buf[n - 1] = 0;
    line += buf;
    if (eol) {
      lines.push_back(line);
      line.clear();
    }
  }
  if (!line.empty()) lines.push_back(line);
  fclose(f);
  std::vector<std::string> tok;
  for (size_t i = 0; i < lines.size();) {
    split_tokens(lines[i], tok);
    if (tok.empty()) {
      i++;
      continue;
    }
    Sequence s;
    s.name = lines[i];
    s.bases = i + 1 < lines.size() ? lines[i + 1] : std::string();
    s.keep_case = true;
    out.push_back(s);
    i += 4;
  }
  return true;
}

static bool read_one_fasta(const std::string &file, std::vector<Sequence> &out, std::string *err) {
  FILE *f = fopen(file.c_str(), "rb");
  if (!f) {
    if (err) *err = "cannot open " + file;
    return false;
  }
  std::string line, pending;  // pending = bases collected for the current record
  std::vector<std::string> tok;
  bool have_record = false;
  char buf[1 << 16];
  bool first_line = true, is_fastq = false;
  auto flush_line = [&](const std::string &ln) -> bool {
    split_tokens(ln, tok);
    if (tok.empty()) return true;
    if (first_line && tok[0][0] == '@') {  // "It's a fastq file!!!" (DNAVector.cc:1223-1228): re-read as FASTQ
      is_fastq = true;
      return false;
    }
    first_line = false;
    if (tok[0][0] == '>') {
      if (have_record) {
        out.back().bases.swap(pending);
        pending.clear();
      }
      std::string name = tok[0].substr(1);
      for (size_t i = 1; i < tok.size(); i++) name += "_" + tok[i];
      { Sequence rec; rec.name = name; out.push_back(rec); }
      have_record = true;
    } else {
      pending += tok[0];
    }
    return true;
  };
  while (fgets(buf, sizeof(buf), f)) {
    size_t n = strlen(buf);
    const bool eol = n > 0 && buf[n - 1] == '\n';
    if (eol) buf[n - 1] = 0;
    line += buf;
    if (eol) {
      if (!flush_line(line)) {
        fclose(f);
        return is_fastq ? read_one_fastq(file, out, err) : false;
      }
      line.clear();
    }
  }
  if (!line.empty() && !flush_line(line)) {
    fclose(f);
    return is_fastq ? read_one_fastq(file, out, err) : false;
  }
  fclose(f);
  if (have_record) out.back().bases.swap(pending);
  return true;
}

bool read_fasta(const std::string &files, std::vector<Sequence> &out, std::string *err) {
  out.clear();
  size_t pos = 0;
  while (pos <= files.size()) {
    size_t c = files.find(',', pos);
    if (c == std::string::npos) c = files.size();
    const std::string one = files.substr(pos, c - pos);
    if (!one.empty() && !read_one_fasta(one, out, err)) return false;
    pos = c + 1;
  }
  for (Sequence &s : out)
    if (!s.keep_case)  // ReadQ does not upper-case
      for (char &ch : s.bases) ch = (char)toupper((unsigned char)ch);
  return true;
}

// ------------------------------------------------------------------------------------------------
void chunk_sequences(const std::vector<Sequence> &seqs, int size, int overlap, int n_blocks, int my_block,
                     ChunkList &out) {
  out = ChunkList();
  const int stride = size - overlap;
  int64_t total = 0;
  int n = 0;
  for (const Sequence &s : seqs) {
    total += (int64_t)s.bases.size();
    if ((int)s.bases.size() >= 6) n += 1 + (int)s.bases.size() / stride;
  }
  out.blob.reserve((size_t)total + 16);
  out.seq_sizes.resize(seqs.size());
  out.names.resize(seqs.size());
  int first = 0, last = n;
  if (n_blocks > 0) {  // SeqChunk.cc:104-116
    const int bsize = (n + n_blocks - 1) / n_blocks;
    first = my_block * bsize;
    last = (my_block + 1) * bsize;
    if (first == last) {
      first = 0;
      last = n;
    }
  }
  int k = 0;
  for (size_t i = 0; i < seqs.size(); i++) {
    const std::string &b = seqs[i].bases;
    const int64_t base_off = (int64_t)out.blob.size();
    out.blob += b;
    out.seq_sizes[i] = (int32_t)b.size();
    out.names[i] = seqs[i].name;
    const int l = (int)b.size();
    if (l < 6) continue;
    const int nchunks = 1 + l / stride;
    for (int j = 0; j < nchunks; j++, k++) {
      const int start = j * stride;
      int end = (j + 1) * stride + overlap;
      if (end >= l) end = l;
      int len = end > start ? end - start : 0;
      if (k < first || k >= last) len = 0;  // not this process's block: chunk stays empty
      if (len > 0) {  // all N/X -> emptied (SeqChunk.cc:150-153)
        int nn = 0;
        for (int x = start; x < start + len; x++) nn += (b[x] == 'N' || b[x] == 'X');
        if (nn >= len) len = 0;
      }
      out.offsets.push_back(base_off + start);
      out.lens.push_back(len);
      out.starts.push_back(start);
      out.seq_ids.push_back((int32_t)i);
    }
  }
}

// ------------------------------------------------------------------------------------------------
void guide_chunks(const std::vector<Sequence> &targets, const std::vector<Sequence> &queries, const MatchFile &chained,
                  int size, ChunkList &t_out, ChunkList &q_out, std::vector<int> &orientation) {
  t_out = ChunkList();
  q_out = ChunkList();
  orientation.clear();
  std::vector<int64_t> t_base(targets.size()), q_base(queries.size());
  for (size_t i = 0; i < targets.size(); i++) {
    t_base[i] = (int64_t)t_out.blob.size();
    t_out.blob += targets[i].bases;
    t_out.seq_sizes.push_back((int32_t)targets[i].bases.size());
    t_out.names.push_back(targets[i].name);
  }
  for (size_t i = 0; i < queries.size(); i++) {
    q_base[i] = (int64_t)q_out.blob.size();
    q_out.blob += queries[i].bases;
    q_out.seq_sizes.push_back((int32_t)queries[i].bases.size());
    q_out.names.push_back(queries[i].name);
  }
  const int max_len = 100000, min_len = 20, lap = 32;
  for (size_t i = 1; i < chained.matches.size(); i++) {
    const t_result &m = chained.matches[i], &n = chained.matches[i - 1];
    const int t_id = (int32_t)m.target_id, q_id = (int32_t)m.query_id;
    if (t_id != (int32_t)n.target_id || q_id != (int32_t)n.query_id || (m.reverse != 0) != (n.reverse != 0)) continue;
    if (t_id < 0 || t_id >= (int)targets.size() || q_id < 0 || q_id >= (int)queries.size()) continue;
    const int force = m.reverse ? -1 : 1;
    const int len = (int32_t)n.len;
    int start_t = (int32_t)n.tstart + len - lap, start_q = (int32_t)(int64_t)n.qstart + len - lap;
    if (start_t < 0) start_t = 0;
    if (start_q < 0) start_q = 0;
    const int end_t = (int32_t)m.tstart + lap;
    int end_q = (int32_t)(int64_t)m.qstart + lap;
    const int target_size = end_t - start_t, query_size = end_q - start_q;
    if (target_size > max_len || query_size > max_len) continue;
    if (target_size < min_len || query_size < min_len) continue;
    const int mx = target_size > query_size ? target_size : query_size;
    const int pieces = 1 + mx / size;
    const int t_chunk = target_size / pieces, q_chunk = query_size / pieces;
    const int t_len = (int)targets[t_id].bases.size(), q_len = (int)queries[q_id].bases.size();
    if (m.reverse) {  // the match's query coordinates are on the reverse strand: take the forward window
      start_q = q_len - end_q;
      end_q = start_q + query_size;
    }
    int t_iter = start_t, q_iter = start_q;
    while (q_iter < end_q && t_iter < end_t) {
      if (q_iter + q_chunk >= q_len) break;
      if (t_iter + t_chunk >= t_len) break;
      if (q_iter < 0) q_iter = 0;  // "Minor error (q)"
      if (t_iter < 0) t_iter = 0;
      t_out.offsets.push_back(t_base[t_id] + t_iter);
      t_out.lens.push_back(t_chunk);
      t_out.starts.push_back(t_iter);
      t_out.seq_ids.push_back(t_id);
      q_out.offsets.push_back(q_base[q_id] + q_iter);
      q_out.lens.push_back(q_chunk);
      q_out.starts.push_back(q_iter);
      q_out.seq_ids.push_back(q_id);
      orientation.push_back(force);
      q_iter += t_chunk / 2;  // sic: each side advances by half of the OTHER side's piece
      t_iter += q_chunk / 2;
    }
  }
}

// ------------------------------------------------------------------------------------------------
HomologyByXCorr::HomologyByXCorr() {}
HomologyByXCorr::~HomologyByXCorr() {
  if (ctx_) sx_destroy(ctx_);
  if (multi_) sx_multi_destroy(multi_);
}

bool HomologyByXCorr::init(const Options &o, const ChunkList &target, const ChunkList &query) {
  sx_config cfg;
  sx_default_config(&cfg);
  cfg.device = o.device;
  cfg.t_chunk = o.t_chunk;
  cfg.q_chunk = o.q_chunk;
  cfg.cutoff = o.cutoff;
  cfg.cutoff_fast = o.cutoff_fast;
  cfg.min_len = o.min_len;
  cfg.use_prob_table = o.prob_table ? 1 : 0;
  cfg.prob_table_value = o.min_prob_flag;
  // the slave never applies -min_prob (it keeps prob >= 0.99, Slave.cc:76); the standalone tool does
  cfg.min_prob = o.standalone_semantics ? o.min_prob_flag : 0.99;
  cfg.rc_coord_mode = o.standalone_semantics ? 1 : 0;
  cfg.max_batch_pairs = o.max_batch_pairs;
  cfg.sort_results = o.sort_results ? 1 : 0;
  target_total_ = 0;  // Slave.cc:405-408: ALL target sequence lengths
  for (int32_t s : target.seq_sizes) target_total_ += (double)s;
  if (o.target_total > 0) target_total_ = o.target_total;  // "Using target size (guided)", tools/...:714-717
  cfg.target_total = target_total_;
  if (o.n_gpus > 0) {  // several GPUs behind one handle: every GPU gets its range of the target list
    std::vector<int32_t> devs;
    for (int d = 0; d < o.n_gpus; d++) devs.push_back(d);
    if (sx_multi_create(&cfg, devs.data(), (int32_t)devs.size(), 0, 1, &multi_) != SX_OK) {
      err_ = sx_multi_last_error();
      return false;
    }
    if (o.prob_table) {
      std::vector<double> tab((size_t)512 * 2048);
      if (sx_build_prob_table(target_total_, tab.data()) != SX_OK || sx_multi_set_prob_table(multi_, tab.data()) != SX_OK) {
        err_ = sx_multi_last_error();
        return false;
      }
    }
    query_blob_ = query.blob;  // sx_multi fetches query ranges on demand: the blob must outlive the calls
    if (sx_multi_set_targets(multi_, target.blob.data(), target.offsets.data(), target.lens.data(), target.starts.data(),
                             target.seq_ids.data(), target.n(), target.seq_sizes.data(),
                             (int32_t)target.seq_sizes.size()) != SX_OK ||
        sx_multi_set_queries(multi_, query_blob_.data(), query.offsets.data(), query.lens.data(), query.starts.data(),
                             query.seq_ids.data(), query.n(), query.seq_sizes.data(),
                             (int32_t)query.seq_sizes.size()) != SX_OK) {
      err_ = sx_multi_last_error();
      return false;
    }
    return true;
  }
  if (sx_create(&cfg, &ctx_) != SX_OK) {
    err_ = sx_last_error();
    return false;
  }
  if (o.prob_table) {
    std::vector<double> tab((size_t)512 * 2048);
    if (sx_build_prob_table(target_total_, tab.data()) != SX_OK || sx_set_prob_table(ctx_, tab.data()) != SX_OK) {
      err_ = sx_last_error();
      return false;
    }
  }
  if (sx_set_targets(ctx_, target.blob.data(), target.offsets.data(), target.lens.data(), target.starts.data(),
                     target.seq_ids.data(), target.n(), target.seq_sizes.data(), (int32_t)target.seq_sizes.size()) !=
          SX_OK ||
      sx_set_queries(ctx_, query.blob.data(), query.offsets.data(), query.lens.data(), query.starts.data(),
                     query.seq_ids.data(), query.n(), query.seq_sizes.data(), (int32_t)query.seq_sizes.size()) != SX_OK) {
    err_ = sx_last_error();
    return false;
  }
  return true;
}

bool HomologyByXCorr::align_targets(const t_pair *p, int n, std::vector<t_result> &results) {
  if (!ctx_ && !multi_) {
    err_ = "not initialised";
    return false;
  }
  auto call = [&](t_result *out, int64_t cap, int64_t *got) {
    return multi_ ? sx_multi_align_blocks(multi_, p, n, out, cap, got) : sx_align_blocks(ctx_, p, n, out, cap, got);
  };
  const size_t base = results.size();
  int64_t cap = 1 << 16, got = 0;
  results.resize(base + (size_t)cap);
  int rc = call(results.data() + base, cap, &got);
  if (rc == SX_ERR_CAPACITY) {  // never truncated: ask again with the size the library reported
    cap = got;
    results.resize(base + (size_t)cap);
    rc = call(results.data() + base, cap, &got);
  }
  if (rc != SX_OK) {
    results.resize(base);
    err_ = multi_ ? sx_multi_last_error() : sx_last_error();
    return false;
  }
  results.resize(base + (size_t)got);
  return true;
}

bool HomologyByXCorr::align_target(const t_pair &p, std::vector<t_result> &results) {
  return align_targets(&p, 1, results);
}

bool HomologyByXCorr::stats(sx_stats *s) const {
  if (multi_) return sx_multi_get_stats(multi_, -1, s) == SX_OK;
  return ctx_ && sx_get_stats(ctx_, s) == SX_OK;
}

// ------------------------------------------------------------------------------------------------
namespace {
struct Writer {
  FILE *f;
  bool ok = true;
  template <typename T>
  void put(const T &v) {
    if (ok && fwrite(&v, sizeof(T), 1, f) != 1) ok = false;
  }
  void str(const std::string &s) {  // CMWriteFileStream::WriteString: long length incl. NUL, then the bytes
    const long len = (long)s.size() + 1;
    put(len);
    if (ok && fwrite(s.c_str(), (size_t)len, 1, f) != 1) ok = false;
  }
};
struct Reader {
  FILE *f;
  bool ok = true;
  template <typename T>
  void get(T &v) {
    if (ok && fread(&v, sizeof(T), 1, f) != 1) ok = false;
  }
  void str(std::string &s) {
    long len = 0;
    get(len);
    if (!ok || len < 1 || len > (1L << 24)) {
      ok = false;
      return;
    }
    std::vector<char> b((size_t)len);
    if (fread(b.data(), (size_t)len, 1, f) != 1) ok = false;
    s.assign(b.data(), strnlen(b.data(), (size_t)len));
  }
};
}  // namespace

bool MatchFile::write(const std::string &path, std::string *err) const {
  FILE *f = fopen(path.c_str(), "wb");
  if (!f) {
    if (err) *err = "cannot create " + path;
    return false;
  }
  Writer w{f};
  const int32_t ver = 3;
  w.put(ver);
  w.put((int32_t)target_names.size());
  for (const std::string &s : target_names) w.str(s);
  w.put((int32_t)query_names.size());
  for (const std::string &s : query_names) w.str(s);
  w.put((int32_t)matches.size());
  for (size_t mi = 0; mi < matches.size(); mi++) {
    const t_result &m = matches[mi];
    // SingleMatch::Write (SequenceMatch.cc:49-64): all coordinates are 32-bit ints in the file
    w.put((int32_t)m.target_id);
    w.put((int32_t)m.query_id);
    w.put((int32_t)m.query_size);
    w.put((int32_t)m.tstart);
    w.put((int32_t)(int64_t)m.qstart);
    w.put((int32_t)m.len);
    w.put((int32_t)(m.reverse ? 1 : 0));
    const double nmatch = mi < n_matches.size() ? n_matches[mi] : m.ident * (double)(int32_t)m.len;  // AddMatches(ident * len)
    w.put(nmatch);
    w.put(m.prob);
    w.put(m.ident);
  }
  for (size_t i = 0; i < target_names.size(); i++) w.put(i < target_sizes.size() ? target_sizes[i] : (int32_t)0);
  for (size_t i = 0; i < query_names.size(); i++) w.put(i < query_sizes.size() ? query_sizes[i] : (int32_t)0);
  const bool ok = w.ok && fclose(f) == 0;
  if (!ok && err) *err = "write error on " + path;
  return ok;
}

namespace {
struct MatchRec {
  t_result r;
  double nmatch;
};
std::vector<MatchRec> zip_matches(const MatchFile &mf) {
  std::vector<MatchRec> v(mf.matches.size());
  for (size_t i = 0; i < v.size(); i++) {
    v[i].r = mf.matches[i];
    v[i].nmatch = i < mf.n_matches.size() ? mf.n_matches[i] : mf.matches[i].ident * (double)(int32_t)mf.matches[i].len;
  }
  return v;
}
void unzip_matches(const std::vector<MatchRec> &v, MatchFile &mf) {
  mf.matches.resize(v.size());
  mf.n_matches.resize(v.size());
  for (size_t i = 0; i < v.size(); i++) {
    mf.matches[i] = v[i].r;
    mf.n_matches[i] = v[i].nmatch;
  }
}
}  // namespace

void MatchFile::sort() {
  std::vector<MatchRec> v = zip_matches(*this);
  std::sort(v.begin(), v.end(), [](const MatchRec &x, const MatchRec &y) {
    const t_result &a = x.r, &b = y.r;
    // the file holds 32-bit ints (SingleMatch); compare what the reference compares
    const int32_t at = (int32_t)a.target_id, bt = (int32_t)b.target_id, aq = (int32_t)a.query_id, bq = (int32_t)b.query_id;
    if (at != bt) return at < bt;
    if (aq != bq) return aq < bq;
    if ((a.reverse != 0) != (b.reverse != 0)) return a.reverse == 0;
    const int32_t as = (int32_t)a.tstart, bs = (int32_t)b.tstart;
    if (as == bs) return (int32_t)a.len < (int32_t)b.len;
    return as < bs;
  });
  unzip_matches(v, *this);
}

void MatchFile::collapse() {
  if (matches.empty()) return;  // the reference reads element 0 of an empty vector here
  auto laps = [](int32_t a, int32_t b) {
    const int32_t c = a < b ? b - a : a - b;
    return c < 4;
  };
  const std::vector<MatchRec> in = zip_matches(*this);
  std::vector<MatchRec> out;
  out.reserve(in.size());
  MatchRec n = in[0];
  for (size_t i = 1; i < in.size(); i++) {
    const t_result &s1 = in[i - 1].r, &s2 = in[i].r;
    if ((int32_t)s1.target_id == (int32_t)s2.target_id && (s1.reverse != 0) == (s2.reverse != 0) &&
        laps((int32_t)s1.tstart, (int32_t)s2.tstart) && laps((int32_t)(int64_t)s1.qstart, (int32_t)(int64_t)s2.qstart)) {
      const int32_t nq = (int32_t)(int64_t)n.r.qstart, s2q = (int32_t)(int64_t)s2.qstart;
      if (s2q < nq) {
        n.r.len = (uint64_t)(int64_t)(nq + (int32_t)n.r.len - s2q);
        n.r.qstart = (uint64_t)(int64_t)s2q;
      } else {
        n.r.len = (uint64_t)(int64_t)(s2q + (int32_t)s2.len - nq);
      }
    } else {
      out.push_back(n);
      n = in[i];
    }
  }
  unzip_matches(out, *this);  // the running match `n` is not stored: MultiMatches::Collapse ends the same way
}

namespace {
// TransPenalty (MatchDynProg.cc:43-92), operation for operation
double trans_penalty(int startT1, int startQ1, bool rc1, int startT2, int startQ2, bool rc2) {
  if (rc1 != rc2) return 20.;  // RC_PENALTY
  double expect = (double)(startT2 - startT1);
  double observe = (double)(startQ2 - startQ1);
  double sigma = 0;
  if (expect > 1000000) sigma = std::sqrt(expect) / 10. - 100;
  if (rc1 == true && rc2 == true) observe = -observe;
  double diff = observe - expect;
  if (diff < 0.) diff = -diff;
  double flat = 0.01;
  if (diff == 1) flat = 0;
  if (expect < 300.) expect = 300;
  double p = diff / expect + sigma + flat;
  if (p > 40) p = 40;  // MAX_PENALTY
  return p;
}
// GetRepeatScore (MatchDynProg.cc:94-112)
double repeat_score(int v) {
  if (v <= 1) return 10.;
  if (v == 2) return 5.;
  if (v == 3) return 2.5;
  if (v == 4) return 1.2;
  if (v >= 25) return 0.000001;
  if (v >= 10) return 0.001;
  return 0.5;
}
struct ChainRec {  // SingleMatchDP (MatchDynProg.cc:115-163)
  MatchRec m;
  double score = 999999999999999.;  // PRETTY_INFINITE
  double pen = -1.;
  int back = -1;
  double rep = 100.;
  void update(double s, int b) {
    if (s < score) {
      score = s;
      back = b;
    }
  }
  double get_score() {
    if (pen < 0.) pen = -200 * std::log(m.r.prob);  // MatchPenalty
    return score + pen;
  }
};
bool match_less(const t_result &a, const t_result &b) {  // SingleMatch::operator< (SequenceMatch.h:93-109)
  const int32_t at = (int32_t)a.target_id, bt = (int32_t)b.target_id, aq = (int32_t)a.query_id, bq = (int32_t)b.query_id;
  if (at != bt) return at < bt;
  if (aq != bq) return aq < bq;
  if ((a.reverse != 0) != (b.reverse != 0)) return a.reverse == 0;
  const int32_t as = (int32_t)a.tstart, bs = (int32_t)b.tstart;
  if (as == bs) return (int32_t)a.len < (int32_t)b.len;
  return as < bs;
}
// MatchDynProg::Chain (MatchDynProg.cc:199-243)
void chain_one_target(std::vector<ChainRec> &v, std::vector<MatchRec> &out) {
  std::sort(v.begin(), v.end(), [](const ChainRec &a, const ChainRec &b) { return (int32_t)a.m.r.tstart < (int32_t)b.m.r.tstart; });
  if (v.empty()) return;
  const int la_limit = 2000, la_dist = 250000;
  v[0].update(0., -1);
  const int n = (int)v.size();
  for (int i = 0; i < n; i++) {
    ChainRec &one = v[i];
    double skip_score = 0.;
    int fed = 0;
    for (int j = i + 1; j < n; j++) {
      ChainRec &two = v[j];
      if ((int32_t)two.m.r.tstart - (int32_t)one.m.r.tstart > la_dist && fed > 10) break;
      if (j - i > la_limit) break;
      fed++;
      double trans = 150;  // GetHighMaxPenalty
      if ((int32_t)one.m.r.query_id == (int32_t)two.m.r.query_id)
        trans = trans_penalty((int32_t)one.m.r.tstart, (int32_t)(int64_t)one.m.r.qstart, one.m.r.reverse != 0,
                              (int32_t)two.m.r.tstart, (int32_t)(int64_t)two.m.r.qstart, two.m.r.reverse != 0);
      trans += skip_score;
      skip_score += two.rep;
      two.update(one.get_score() + trans, i);
    }
  }
  std::vector<MatchRec> chain;
  for (int i = n - 1; i >= 0; i = v[i].back) chain.push_back(v[i].m);
  std::sort(chain.begin(), chain.end(), [](const MatchRec &a, const MatchRec &b) { return match_less(a.r, b.r); });
  out.insert(out.end(), chain.begin(), chain.end());
}
}  // namespace

void MatchFile::chain(MatchFile &out) const {
  out = MatchFile();
  out.target_names = target_names;
  out.query_names = query_names;
  out.target_sizes = target_sizes;
  out.query_sizes = query_sizes;
  const int n_t = (int)target_sizes.size(), n_q = (int)query_sizes.size();
  const std::vector<MatchRec> in = zip_matches(*this);
  const int n_in = (int)in.size();
  // "Filling out repeat lists": how many matches cover every base (saturating at 100), MatchDynProg.cc:411-481
  std::vector<std::vector<char>> mult_t((size_t)n_t), mult_q((size_t)n_q);
  for (int i = 0; i < n_t; i++) mult_t[i].assign((size_t)std::max(target_sizes[i], 0), 0);
  for (int i = 0; i < n_q; i++) mult_q[i].assign((size_t)std::max(query_sizes[i], 0), 0);
  int last_t = -1, last_q = -1, last_st = 0, last_sq = 0, last_len = 0;
  for (const MatchRec &mr : in) {
    const int tid = (int32_t)mr.r.target_id, qid = (int32_t)mr.r.query_id;
    const int st = (int32_t)mr.r.tstart, sq = (int32_t)(int64_t)mr.r.qstart, len = (int32_t)mr.r.len;
    if (qid < 0) continue;
    int start_t = st, start_q = sq;
    if (tid == last_t && qid == last_q && st <= last_st + last_len && sq <= last_sq + last_len && st > last_st &&
        sq > last_sq) {
      start_t = last_st + last_len;
      start_q = last_sq + last_len;
    }
    last_t = tid; last_q = qid; last_st = st; last_sq = sq; last_len = len;
    if (tid < 0 || tid >= n_t || qid >= n_q) continue;  // the reference indexes out of bounds here
    std::vector<char> &t = mult_t[tid], &q = mult_q[qid];
    for (int j = start_t; j < st + len; j++)
      if (j > 0 && j < (int)t.size() && t[j] < 100) t[j]++;
    for (int j = start_q; j < sq + len; j++)
      if (j > 0 && j < (int)q.size() && q[j] < 100) q[j]++;
  }
  // per target: the index range its matches span in the (sorted) list
  std::vector<int> first_i((size_t)n_t, n_in + 1), last_i((size_t)n_t, -1);
  for (int i = 0; i < n_in; i++) {
    const int id = (int32_t)in[i].r.target_id;
    if (id < 0 || id >= n_t) continue;
    if (i < first_i[id]) first_i[id] = i;
    if (i > last_i[id]) last_i[id] = i;
  }
  // One chain per target sequence (MatchDynProg.cc:483-557); the chains do not depend on each other, so they run on
  // as many host threads as there are targets (at most the core count) and are appended in target order -- the
  // output is the sequential one, record for record.  (The reference does them one after the other on one thread.)
  std::vector<std::vector<MatchRec>> per_target((size_t)n_t);
  std::atomic<int> next_target(0);
  auto worker = [&]() {
  for (int j = next_target.fetch_add(1); j < n_t; j = next_target.fetch_add(1)) {
    std::vector<MatchRec> &result = per_target[(size_t)j];
    const int last = last_i[j], first = last == -1 ? 0 : first_i[j];
    const std::vector<char> &t = mult_t[j];
    std::vector<ChainRec> dp;
    for (int i = first; i <= last; i++) {
      const t_result &m = in[i].r;
      const int qid = (int32_t)m.query_id;
      if (qid < 0 || qid >= n_q) continue;  // out of bounds in the reference
      const std::vector<char> &q = mult_q[qid];
      const int mid_t = (int32_t)m.tstart + (int32_t)m.len / 2, mid_q = (int32_t)(int64_t)m.qstart + (int32_t)m.len / 2;
      if (mid_t < 0 || mid_t >= (int)t.size() || mid_q < 0 || mid_q >= (int)q.size()) continue;  // ditto
      double rep = repeat_score(t[mid_t]);
      const double rep2 = repeat_score(q[mid_q]);
      if (rep2 < rep) rep = rep2;
      if (rep > 0.0001) {
        ChainRec c;
        c.m = in[i];
        c.rep = rep;
        dp.push_back(c);
      }
    }
    chain_one_target(dp, result);
  }
  };
  const int n_threads = std::max(1, std::min<int>(n_t, (int)std::thread::hardware_concurrency()));
  std::vector<std::thread> pool;
  for (int w = 1; w < n_threads; w++) pool.emplace_back(worker);
  worker();
  for (std::thread &th : pool) th.join();
  std::vector<MatchRec> result;
  for (const std::vector<MatchRec> &r : per_target) result.insert(result.end(), r.begin(), r.end());
  unzip_matches(result, out);
}

void MatchFile::chain_dups(MatchFile &out) const {
  MatchFile tmp;
  chain(tmp);
  tmp.sort();
  MatchFile rest;
  rest.target_names = target_names;
  rest.query_names = query_names;
  rest.target_sizes = target_sizes;
  rest.query_sizes = query_sizes;
  const int n_t = (int)target_sizes.size(), n_q = (int)query_sizes.size();
  const std::vector<MatchRec> in = zip_matches(*this);
  std::vector<MatchRec> rest_recs;
  for (int j = 0; j < n_t; j++) {
    // the query with the most matches in this target's first chain ("best query"); ties keep the lowest id
    std::vector<int> q_hits((size_t)n_q, 0);
    for (const t_result &m : tmp.matches) {
      if ((int32_t)m.target_id != j) continue;
      const int q = (int32_t)m.query_id;
      if (q >= 0 && q < n_q) q_hits[q]++;
    }
    int best_query = -1, max_hits = 0;
    for (int i = 0; i < n_q; i++)
      if (q_hits[i] > max_hits) {
        max_hits = q_hits[i];
        best_query = i;
      }
    for (const MatchRec &mr : in) {
      if ((int32_t)mr.r.target_id != j) continue;
      if ((int32_t)mr.r.query_id != best_query) rest_recs.push_back(mr);
    }
  }
  unzip_matches(rest_recs, rest);
  rest.chain(out);
  std::vector<MatchRec> all = zip_matches(out), first = zip_matches(tmp);
  all.insert(all.end(), first.begin(), first.end());
  unzip_matches(all, out);
  out.sort();
}

bool MatchFile::read(const std::string &path, std::string *err) {
  FILE *f = fopen(path.c_str(), "rb");
  if (!f) {
    if (err) *err = "cannot open " + path;
    return false;
  }
  Reader r{f};
  int32_t ver = 0, nt = 0, nq = 0, n = 0;
  r.get(ver);
  r.get(nt);
  if (!r.ok || ver != 3 || nt < 0) {
    fclose(f);
    if (err) *err = path + ": not a version-3 xcorr match file";
    return false;
  }
  target_names.resize((size_t)nt);
  for (auto &s : target_names) r.str(s);
  r.get(nq);
  if (!r.ok || nq < 0) nq = 0;
  query_names.resize((size_t)nq);
  for (auto &s : query_names) r.str(s);
  r.get(n);
  if (!r.ok || n < 0) n = 0;
  matches.resize((size_t)n);
  n_matches.assign((size_t)n, 0.);
  for (size_t mi = 0; mi < matches.size(); mi++) {
    t_result &m = matches[mi];
    int32_t tid, qid, qlen, st, sq, len, rc;
    double nmatch = 0.;
    memset(&m, 0, sizeof(m));
    r.get(tid); r.get(qid); r.get(qlen); r.get(st); r.get(sq); r.get(len); r.get(rc);
    r.get(nmatch); r.get(m.prob); r.get(m.ident);
    n_matches[mi] = nmatch;
    m.target_id = (uint64_t)(int64_t)tid;
    m.query_id = (uint64_t)(int64_t)qid;
    m.query_size = (uint64_t)(int64_t)qlen;
    m.tstart = (uint64_t)(int64_t)st;
    m.qstart = (uint64_t)(int64_t)sq;
    m.len = (uint64_t)(int64_t)len;
    m.reverse = (uint8_t)(rc != 0);
  }
  target_sizes.resize((size_t)nt);
  query_sizes.resize((size_t)nq);
  for (auto &v : target_sizes) r.get(v);
  for (auto &v : query_sizes) r.get(v);
  fclose(f);
  if (!r.ok && err) *err = path + ": truncated match file";
  return r.ok;
}

}  // namespace sxh
