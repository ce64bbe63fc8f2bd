// libsatsuma_b200: several GPUs behind one handle (include/satsuma_xcorr.h, sx_multi_*).
//
// The reference spreads the chunk-pair grid over machines by target range: `-nblocks / -block` of the standalone
// tool (analysis/SeqChunk.cc:104-116) and N slave processes fed by the master's WorkQueue
// (analysis/WorkQueue.cc:290-312).  Here the same split happens inside one process: the flat target chunk list is
// cut into `shard_world * n_devices` contiguous ranges, this handle owns `n_devices` of them (one per GPU), every GPU
//   * receives ONLY the bases of its target range and keeps the spectra of those chunks in its HBM,
//   * receives, per call, only the query chunks its share of the block list touches,
//   * runs its share of every t_pair block (blocks that straddle a boundary are split between the neighbours)
//     on its own host thread and stream,
// and the records of all GPUs are gathered into the caller's single buffer.  No GPU talks to another: there is no
// collective on this path.  With shard_world > 1 several processes (torchrun ranks, slaves on other machines) own
// disjoint parts of the same split and the union of their outputs is the unsharded result, pair for pair.
#include "../../include/satsuma_xcorr.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace {

struct Shard {
  sx_ctx *ctx = nullptr;
  int device = 0;
  int32_t t_lo = 0, t_hi = 0;    // target chunks [t_lo, t_hi) of the caller's list
  int32_t q_lo = 0, q_hi = 0;    // query chunks currently resident on this GPU (empty when q_lo == q_hi)
  int64_t q_epoch = -1;          // sx_multi_set_queries call the resident range was cut from
  std::vector<sx_pair> blocks;   // this call's share, indices local to the shard
  std::vector<sx_result> out;
  int64_t n_out = 0;
  int rc = SX_OK;
  std::string err;
  int64_t h2d_bytes = 0;
};

struct ChunkTable {  // the caller's flat chunk list (host copies of the small arrays, the blob stays the caller's)
  const char *bases = nullptr;
  std::vector<int64_t> offsets;
  std::vector<int32_t> lens, starts, seq_ids, seq_sizes;
  int32_t n = 0;
};

thread_local std::string g_merr;

int mfail(int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_merr = buf;
  return code;
}

// chunks [lo, hi) of `tab` as a self-contained list: offsets rebased to the first byte any of them uses
int set_range(sx_ctx *ctx, bool targets, const ChunkTable &tab, int32_t lo, int32_t hi, int64_t *bytes) {
  const int32_t n = hi - lo;
  std::vector<int64_t> off((size_t)std::max(n, 0));
  int64_t b0 = INT64_MAX, b1 = 0;
  for (int32_t i = lo; i < hi; i++) {
    if (tab.lens[i] == 0) continue;
    b0 = std::min(b0, tab.offsets[i]);
    b1 = std::max(b1, tab.offsets[i] + tab.lens[i]);
  }
  if (b0 == INT64_MAX) b0 = b1 = 0;
  for (int32_t i = lo; i < hi; i++) off[i - lo] = tab.lens[i] == 0 ? 0 : tab.offsets[i] - b0;
  if (bytes) *bytes += b1 - b0;
  auto fn = targets ? sx_set_targets : sx_set_queries;
  static const char dummy[1] = {0};
  return fn(ctx, n > 0 ? tab.bases + b0 : dummy, off.data(), tab.lens.data() + lo, tab.starts.data() + lo,
            tab.seq_ids.data() + lo, n, tab.seq_sizes.data(), (int32_t)tab.seq_sizes.size());
}

int fill_table(ChunkTable &tab, const char *bases, const int64_t *offsets, const int32_t *lens, const int32_t *starts,
               const int32_t *seq_ids, int32_t n, const int32_t *seq_sizes, int32_t n_seqs, const char *what) {
  if (n < 0 || (n > 0 && (!bases || !offsets || !lens))) return mfail(SX_ERR_ARG, "%s: null argument", what);
  if (seq_ids && (!seq_sizes || n_seqs <= 0)) return mfail(SX_ERR_ARG, "%s: seq_ids given without seq_sizes", what);
  tab.bases = bases;
  tab.n = n;
  tab.offsets.assign(offsets, offsets + n);
  tab.lens.assign(lens, lens + n);
  if (starts) tab.starts.assign(starts, starts + n); else tab.starts.assign(n, 0);
  if (seq_ids) tab.seq_ids.assign(seq_ids, seq_ids + n); else tab.seq_ids.assign(n, 0);
  if (seq_sizes && n_seqs > 0) tab.seq_sizes.assign(seq_sizes, seq_sizes + n_seqs); else tab.seq_sizes.assign(1, 0);
  return SX_OK;
}

}  // namespace

struct sx_multi {
  sx_config cfg;
  int32_t shard_rank = 0, shard_world = 1;
  std::vector<Shard> shards;
  ChunkTable T, Q;
  int64_t q_epoch = 0;
  bool have_table = false;
};

extern "C" const char *sx_multi_last_error(void) { return g_merr.c_str(); }

extern "C" int sx_multi_create(const sx_config *cfg, const int32_t *devices, int32_t n_devices, int32_t shard_rank,
                               int32_t shard_world, sx_multi **out) {
  if (!cfg || !out) return mfail(SX_ERR_ARG, "sx_multi_create: null argument");
  if (shard_world < 1 || shard_rank < 0 || shard_rank >= shard_world)
    return mfail(SX_ERR_ARG, "sx_multi_create: shard %d of %d", shard_rank, shard_world);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return mfail(SX_ERR_CUDA, "sx_multi_create: no CUDA device; this library has no CPU fallback");
  std::vector<int32_t> devs;
  if (n_devices <= 0 || !devices) {  // every visible GPU
    for (int d = 0; d < ndev; d++) devs.push_back(d);
  } else {
    devs.assign(devices, devices + n_devices);
  }
  for (int32_t d : devs)
    if (d < 0 || d >= ndev) return mfail(SX_ERR_ARG, "sx_multi_create: device %d of %d", d, ndev);
  sx_multi *m = new (std::nothrow) sx_multi();
  if (!m) return mfail(SX_ERR_NOMEM, "sx_multi_create: out of host memory");
  m->cfg = *cfg;
  m->shard_rank = shard_rank;
  m->shard_world = shard_world;
  m->shards.resize(devs.size());
  for (size_t i = 0; i < devs.size(); i++) {
    sx_config c = *cfg;
    c.device = devs[i];
    m->shards[i].device = devs[i];
    const int rc = sx_create(&c, &m->shards[i].ctx);
    if (rc != SX_OK) {
      g_merr = sx_last_error();
      for (Shard &s : m->shards)
        if (s.ctx) sx_destroy(s.ctx);
      delete m;
      return rc;
    }
  }
  *out = m;
  return SX_OK;
}

extern "C" void sx_multi_destroy(sx_multi *m) {
  if (!m) return;
  for (Shard &s : m->shards)
    if (s.ctx) sx_destroy(s.ctx);
  delete m;
}

extern "C" int32_t sx_multi_device_count(const sx_multi *m) { return m ? (int32_t)m->shards.size() : 0; }

extern "C" int sx_multi_target_range(const sx_multi *m, int32_t shard, int32_t *t_lo, int32_t *t_hi) {
  if (!m || shard < 0 || shard >= (int32_t)m->shards.size()) return mfail(SX_ERR_ARG, "sx_multi_target_range: bad shard");
  if (t_lo) *t_lo = m->shards[shard].t_lo;
  if (t_hi) *t_hi = m->shards[shard].t_hi;
  return SX_OK;
}

// one host thread per GPU for the duration of a call
template <class F>
static int for_each_shard(sx_multi *m, F fn) {
  std::vector<std::thread> pool;
  for (size_t i = 1; i < m->shards.size(); i++) pool.emplace_back([&, i]() { fn(m->shards[i]); });
  fn(m->shards[0]);
  for (auto &t : pool) t.join();
  for (Shard &s : m->shards)
    if (s.rc != SX_OK) {
      g_merr = "device " + std::to_string(s.device) + ": " + s.err;
      return s.rc;
    }
  return SX_OK;
}

extern "C" int sx_multi_set_targets(sx_multi *m, const char *bases, const int64_t *offsets, const int32_t *lens,
                                    const int32_t *starts, const int32_t *seq_ids, int32_t n, const int32_t *seq_sizes,
                                    int32_t n_seqs) {
  if (!m) return mfail(SX_ERR_ARG, "sx_multi_set_targets: null handle");
  int rc = fill_table(m->T, bases, offsets, lens, starts, seq_ids, n, seq_sizes, n_seqs, "sx_multi_set_targets");
  if (rc != SX_OK) return rc;
  // targetTotal covers ALL target sequences, whoever holds their chunks (Slave.cc:405-408): every GPU gets the whole
  // table of sequence sizes, from which sx_set_targets derives it (or the caller fixed it in sx_config::target_total)
  const int64_t parts = (int64_t)m->shard_world * (int64_t)m->shards.size();
  const int64_t per = (n + parts - 1) / parts;
  for (size_t i = 0; i < m->shards.size(); i++) {
    const int64_t idx = (int64_t)m->shard_rank * (int64_t)m->shards.size() + (int64_t)i;
    m->shards[i].t_lo = (int32_t)std::min<int64_t>(n, idx * per);
    m->shards[i].t_hi = (int32_t)std::min<int64_t>(n, idx * per + per);
  }
  return for_each_shard(m, [&](Shard &s) {
    s.rc = set_range(s.ctx, true, m->T, s.t_lo, s.t_hi, &s.h2d_bytes);
    if (s.rc != SX_OK) s.err = sx_last_error();
  });
}

extern "C" int sx_multi_set_queries(sx_multi *m, const char *bases, const int64_t *offsets, const int32_t *lens,
                                    const int32_t *starts, const int32_t *seq_ids, int32_t n, const int32_t *seq_sizes,
                                    int32_t n_seqs) {
  if (!m) return mfail(SX_ERR_ARG, "sx_multi_set_queries: null handle");
  // nothing travels yet: every GPU fetches the range its blocks touch when a call needs it.  The caller keeps
  // `bases` alive until the handle is destroyed or the queries are replaced.
  m->q_epoch++;
  return fill_table(m->Q, bases, offsets, lens, starts, seq_ids, n, seq_sizes, n_seqs, "sx_multi_set_queries");
}

extern "C" int sx_multi_set_prob_table(sx_multi *m, const double *table) {
  if (!m || !table) return mfail(SX_ERR_ARG, "sx_multi_set_prob_table: null argument");
  return for_each_shard(m, [&](Shard &s) {
    s.rc = sx_set_prob_table(s.ctx, table);
    if (s.rc != SX_OK) s.err = sx_last_error();
  });
}

extern "C" int sx_multi_invalidate_spectra(sx_multi *m) {
  if (!m) return mfail(SX_ERR_ARG, "null handle");
  for (Shard &s : m->shards) sx_invalidate_spectra(s.ctx);
  return SX_OK;
}

extern "C" int sx_multi_align_blocks(sx_multi *m, const sx_pair *blocks, int32_t n_blocks, sx_result *out, int64_t cap,
                                     int64_t *n_out) {
  if (!m || (n_blocks > 0 && !blocks)) return mfail(SX_ERR_ARG, "sx_multi_align_blocks: null argument");
  for (int32_t b = 0; b < n_blocks; b++) {
    const sx_pair &p = blocks[b];
    if (p.target_from < 0 || p.target_to >= m->T.n || p.query_from < 0 || p.query_to >= m->Q.n)
      return mfail(SX_ERR_ARG, "sx_multi_align_blocks: block %d ranges t[%d,%d] q[%d,%d] out of bounds", b, p.target_from,
                   p.target_to, p.query_from, p.query_to);
  }
  // each GPU's share: the blocks clipped to its target range
  for (Shard &s : m->shards) {
    s.blocks.clear();
    s.rc = SX_OK;
    s.n_out = 0;
    int32_t q_lo = INT32_MAX, q_hi = -1;
    for (int32_t b = 0; b < n_blocks; b++) {
      sx_pair p = blocks[b];
      p.target_from = std::max(p.target_from, s.t_lo);
      p.target_to = std::min(p.target_to, s.t_hi - 1);
      if (p.target_from > p.target_to || p.query_from > p.query_to) continue;
      q_lo = std::min(q_lo, p.query_from);
      q_hi = std::max(q_hi, p.query_to);
      s.blocks.push_back(p);
    }
    if (s.blocks.empty()) continue;
    // queries: keep what is resident when it covers the need, else fetch exactly the range needed
    const bool covered = s.q_epoch == m->q_epoch && s.q_lo <= q_lo && q_hi < s.q_hi;
    if (!covered) {
      s.q_lo = q_lo;
      s.q_hi = q_hi + 1;
      s.q_epoch = -2;  // fetch in the worker
    }
    for (sx_pair &p : s.blocks) {
      p.target_from -= s.t_lo;
      p.target_to -= s.t_lo;
      p.query_from -= s.q_lo;
      p.query_to -= s.q_lo;
    }
  }
  const bool direct = m->shards.size() == 1;  // one GPU: its records go straight into the caller's buffer
  int rc = for_each_shard(m, [&](Shard &s) {
    if (s.blocks.empty()) return;
    if (s.q_epoch == -2) {
      s.rc = set_range(s.ctx, false, m->Q, s.q_lo, s.q_hi, &s.h2d_bytes);
      if (s.rc != SX_OK) {
        s.err = sx_last_error();
        s.q_epoch = -1;
        return;
      }
      s.q_epoch = m->q_epoch;
    }
    if (direct) {
      s.rc = sx_align_blocks(s.ctx, s.blocks.data(), (int32_t)s.blocks.size(), out, cap, &s.n_out);
      if (s.rc != SX_OK) s.err = sx_last_error();
      return;
    }
    if (s.out.size() < 1024) s.out.resize(1024);
    for (int attempt = 0; attempt < 2; attempt++) {
      s.rc = sx_align_blocks(s.ctx, s.blocks.data(), (int32_t)s.blocks.size(), s.out.data(), (int64_t)s.out.size(),
                             &s.n_out);
      if (s.rc != SX_ERR_CAPACITY) break;
      s.out.resize((size_t)s.n_out);  // never truncated: redo with the size the device asked for
    }
    if (s.rc != SX_OK) s.err = sx_last_error();
  });
  if (direct) {
    if (n_out) *n_out = m->shards[0].blocks.empty() ? 0 : m->shards[0].n_out;
    return rc;
  }
  if (rc != SX_OK) return rc;
  // host gather: one list in the caller's buffer, shard after shard
  int64_t total = 0;
  for (Shard &s : m->shards) total += s.blocks.empty() ? 0 : s.n_out;
  if (n_out) *n_out = total;
  if (total > cap || (total > 0 && !out)) return mfail(SX_ERR_CAPACITY, "align: %lld records, buffer holds %lld", (long long)total, (long long)cap);
  int64_t at = 0;
  for (Shard &s : m->shards) {
    if (s.blocks.empty() || s.n_out == 0) continue;
    memcpy(out + at, s.out.data(), sizeof(sx_result) * (size_t)s.n_out);
    at += s.n_out;
  }
  return SX_OK;
}

extern "C" int sx_multi_get_stats(sx_multi *m, int32_t shard, sx_stats *out) {
  if (!m || !out) return mfail(SX_ERR_ARG, "sx_multi_get_stats: null argument");
  if (shard >= (int32_t)m->shards.size()) return mfail(SX_ERR_ARG, "sx_multi_get_stats: shard %d", shard);
  if (shard >= 0) return sx_get_stats(m->shards[shard].ctx, out);
  memset(out, 0, sizeof(*out));  // shard < 0: summed over the GPUs (times: the slowest GPU)
  for (Shard &s : m->shards) {
    sx_stats st;
    const int rc = sx_get_stats(s.ctx, &st);
    if (rc != SX_OK) return rc;
    out->chunk_pairs += st.chunk_pairs; out->strand_pairs += st.strand_pairs; out->signals += st.signals;
    out->candidates += st.candidates; out->segments += st.segments; out->matches += st.matches;
    out->kernel_launches += st.kernel_launches; out->batches += st.batches; out->h2d_bytes += st.h2d_bytes;
    out->d2h_bytes += st.d2h_bytes; out->retries += st.retries; out->positions += st.positions;
    out->ms_encode_fft = std::max(out->ms_encode_fft, st.ms_encode_fft);
    out->ms_xcorr = std::max(out->ms_xcorr, st.ms_xcorr);
    out->ms_scan_score = std::max(out->ms_scan_score, st.ms_scan_score);
    out->ms_total = std::max(out->ms_total, st.ms_total);
  }
  return SX_OK;
}

extern "C" int sx_multi_reset_stats(sx_multi *m) {
  if (!m) return mfail(SX_ERR_ARG, "null handle");
  for (Shard &s : m->shards) sx_reset_stats(s.ctx);
  return SX_OK;
}

extern "C" int sx_multi_stream(sx_multi *m, int32_t shard, void **stream_out) {
  if (!m || shard < 0 || shard >= (int32_t)m->shards.size()) return mfail(SX_ERR_ARG, "sx_multi_stream: bad shard");
  return sx_stream(m->shards[shard].ctx, stream_out);
}
