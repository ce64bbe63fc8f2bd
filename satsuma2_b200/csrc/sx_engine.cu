// libsatsuma_b200: context, HBM layout, batching and the C ABI (include/satsuma_xcorr.h).
//
// Host-side role: what HomologyByXCorr::align_target / Align / FilterMatches do around the
// kernels in the reference (analysis/HomologyByXCorrSlave.cc:168-300): expand block requests
// into chunk pairs, make sure every chunk signal needed has a spectrum resident in HBM
// (target spectra are kept across calls and reused by every query, the reference recomputes
// them per pair), launch the three kernels per batch, map chunk-local matches to sequence
// coordinates and hand t_result-compatible records back.
#include "../../include/satsuma_xcorr.h"
#include "sx_kernels.h"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

using namespace sx;

static thread_local std::string g_err;

static int fail(int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CU(expr)                                                                         \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess)                                                               \
      return fail(_e == cudaErrorMemoryAllocation ? SX_ERR_NOMEM : SX_ERR_CUDA, "%s: %s", #expr, \
                  cudaGetErrorString(_e));                                               \
  } while (0)

namespace {

struct ChunkStore {
  int32_t n = 0;
  uint8_t *d_bases = nullptr;
  size_t blob_bytes = 0, cap_bytes = 0;
  std::vector<int64_t> offsets;
  std::vector<int32_t> lens, starts, seq_ids, seq_sizes;
  // asynchronous upload: the blob travels in pieces on the copy stream, one event per piece; pieces are
  // enqueued when a batch needs them (plus a look-ahead), so target and query pieces interleave in the
  // order the batches consume them
  std::vector<cudaEvent_t> piece_ev;
  size_t piece_bytes = 0, n_pieces = 0, enq_pieces = 0;
  const char *h_src = nullptr;
  bool async_pending = false;
};

struct PairReq {
  int32_t t, q;
  int32_t fast;
};

template <typename T>
struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  int ensure(size_t want) {
    if (want <= n) return SX_OK;
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
    cudaError_t e = cudaMalloc((void **)&p, want * sizeof(T));
    if (e != cudaSuccess) return fail(SX_ERR_NOMEM, "cudaMalloc(%zu bytes): %s", want * sizeof(T), cudaGetErrorString(e));
    n = want;
    return SX_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
};

template <typename T>
struct PinBuf {
  T *p = nullptr;
  size_t n = 0;
  int ensure(size_t want) {
    if (want <= n) return SX_OK;
    if (p) cudaFreeHost(p);
    p = nullptr;
    n = 0;
    cudaError_t e = cudaMallocHost((void **)&p, want * sizeof(T));
    if (e != cudaSuccess) return fail(SX_ERR_NOMEM, "cudaMallocHost(%zu bytes): %s", want * sizeof(T), cudaGetErrorString(e));
    n = want;
    return SX_OK;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    n = 0;
  }
};

}  // namespace

struct sx_ctx {
  sx_config cfg;
  int log2n = 0, N = 0;
  std::mutex mu;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;  // async_upload: host->device copies of the chunk blobs
  // Descriptor uploads and the read-back of counters / records run on their own stream, tied to the compute stream by
  // events: queued in the compute stream they would sit between the kernels of consecutive batches (measured: 0.19 ms
  // of idle device per 5 ms batch)
  cudaStream_t aux_stream = nullptr;   // read-backs (waits for the batch's last kernel)
  cudaStream_t desc_stream = nullptr;  // descriptor uploads (never wait for a kernel)
  cudaEvent_t ev_desc[2] = {nullptr, nullptr};  // descriptors of the batch are on the device
  cudaEvent_t ev_done[2] = {nullptr, nullptr};  // the last kernel of the batch has finished
  cudaEvent_t ev_ctr[2] = {nullptr, nullptr};   // its counter block is on the host
  cudaEvent_t ev_res = nullptr;                 // the records of the last fetched batch have left the result pool
  cudaEvent_t ev[2][7] = {};  // per batch in flight: 0 early start, 1 transforms done, 5 fused kernel done, 6 rest starts, 2 correlation done, 3 scan done, 4 records fetched
  bool profiling = false;
  sx_stats stats;

  ChunkStore T, Q;
  double target_total = 0;
  double z_cut = INFINITY;
  int run_min = 0;                 // compute_run_min() for the parameters below
  double rm_z_cut = NAN;           // ... cached: it only changes with z_cut / the table / min_len
  int rm_use_table = -1, rm_min_len = -1;
  std::vector<double> h_table;     // host copy of the ProbTable (run_min in -prob_table mode)

  // signal slots: [0, n_persist) = cached target spectra (slot == target chunk index),
  // then two per-batch workspaces of n_transient slots each: while the device still works on batch k, batch
  // k+1 is already being encoded into the other one
  size_t n_persist = 0, n_transient = 0;
  DevBuf<float2> spec;
  DevBuf<uint32_t> planes;
  DevBuf<uint8_t> sbytes;
  DevBuf<SlotMeta> meta;
  DevBuf<float2> wn;  // e^{-2 pi i n / N}, n < N/2
  DevBuf<double> ent_table;  // fill_ent_table()
  DevBuf<float2> drift;      // fill_drift_table() (N = 32768)
  std::vector<uint8_t> t_valid;  // persistent target slot holds a spectrum

  // per-batch device buffers
  DevBuf<SigDesc> d_sigs[2];  // one per batch in flight (the next batch's encode is queued behind the current scan)
  DevBuf<int32_t> d_prep_flag[2];  // preparation kernel -> transform / fused kernel (PrepBuf), one set per batch in
  DevBuf<float> d_prep_went[2];    // flight: a batch that is re-run after a pool overflow still finds its own
  DevBuf<double> d_prep_off[2];
  // fused transform + correlation kernel (pair_fused_kernel)
  DevBuf<FusedJob> d_fused[2];      // (the buffers marked [2] exist once per batch in flight: the fused kernel of the next
                                    // batch is queued while the current one still scans)
  DevBuf<uint32_t> d_enc_list[2];   // signals left to the transform kernel when a batch has fused pairs
  DevBuf<unsigned int> d_fail_ctr[2];  // {pairs, signals} handed back by the fused kernel
  DevBuf<uint32_t> d_fail_pairs[2], d_fail_sigs[2];
  DevBuf<float2> d_fused_scratch;   // one parked half-transform per resident CTA
  int fused_grid = 0;               // 2 CTAs per SM
  std::vector<int32_t> sig_of_slot, slot_uses;  // per slot, for the batch being staged (propose_fused)
  DevBuf<SpDesc> d_sps[2];
  DevBuf<uint2> d_cand_ref[2];
  DevBuf<uint32_t> d_lists[2];  // [pair list | direct list] of strand-pair indices (launch_xcorr_findtop)
  DevBuf<uint16_t> d_cand_pool[2];
  DevBuf<ResultRec> d_res;
  DevBuf<SegRec> d_seg_tap;
  DevBuf<BatchCounters> d_ctr[2];
  DevBuf<double> d_table;
  DevBuf<float> d_tap;
  DevBuf<float2> d_scratch;  // split transforms (N = 32768): e[n] / o[n] of every correlation job between the two kernels
  bool have_table = false;

  PinBuf<SigDesc> h_sigs[2];  // descriptor staging, one per batch in flight / being assembled
  PinBuf<SpDesc> h_sps[2];
  PinBuf<uint32_t> h_lists[2];
  PinBuf<FusedJob> h_fused[2];
  PinBuf<uint32_t> h_enc[2];
  PinBuf<ResultRec> h_res;
  PinBuf<BatchCounters> h_ctr;


  Slots slots() const {
    Slots s;
    s.spec = spec.p;
    s.planes = planes.p;
    s.bytes = sbytes.p;
    s.meta = meta.p;
    s.wn = wn.p;
    s.ent_table = ent_table.p;
    s.drift = drift.p;
    return s;
  }
};

// ------------------------------------------------------------------------------------------------
static int pick_log2n(int t_chunk) {
  int n = 2 * t_chunk, l = 0;
  while ((1 << l) < n) l++;
  if ((1 << l) != n) return -1;
  return l;
}

extern "C" void sx_default_config(sx_config *c) {
  memset(c, 0, sizeof(*c));
  c->abi_version = SX_ABI_VERSION;
  c->device = 0;
  c->t_chunk = 4096;
  c->q_chunk = 4096;
  c->cutoff = 1.8;
  c->cutoff_fast = 2.9;
  c->min_len = 0;
  c->use_prob_table = 0;
  c->min_prob = 0.99;
  c->prob_table_value = 0.9999;
  c->target_total = 0;
  c->rc_coord_mode = 0;
  c->max_batch_pairs = 0;
  c->spectra_cache_bytes = 0;
  c->sort_results = 0;
}

extern "C" int sx_abi_version(void) { return SX_ABI_VERSION; }
extern "C" int sx_device_count(void) {
  int n = 0;
  return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0;
}
extern "C" const char *sx_last_error(void) { return g_err.c_str(); }

extern "C" int sx_create(const sx_config *cfg, sx_ctx **out) {
  if (!cfg || !out) return fail(SX_ERR_ARG, "sx_create: null argument");
  if (cfg->abi_version != SX_ABI_VERSION) return fail(SX_ERR_ARG, "sx_create: ABI version %d != %d", cfg->abi_version, SX_ABI_VERSION);
  const int l = pick_log2n(cfg->t_chunk);
  if (l < 0 || !log2n_supported(l))
    return fail(SX_ERR_ARG, "sx_create: t_chunk=%d unsupported (2*t_chunk must be a power of two in [2048,32768])", cfg->t_chunk);
  if (cfg->q_chunk < 1 || cfg->q_chunk > 2 * cfg->t_chunk)
    return fail(SX_ERR_ARG, "sx_create: q_chunk=%d must be in [1, 2*t_chunk]", cfg->q_chunk);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(SX_ERR_CUDA, "sx_create: no CUDA device (%s); this library has no CPU fallback", cudaGetErrorString(e));
  if (cfg->device < 0 || cfg->device >= ndev) return fail(SX_ERR_ARG, "sx_create: device %d of %d", cfg->device, ndev);
  CU(cudaSetDevice(cfg->device));
  sx_ctx *c = new (std::nothrow) sx_ctx();
  if (!c) return fail(SX_ERR_NOMEM, "sx_create: out of host memory");
  c->cfg = *cfg;
  c->log2n = l;
  c->N = 1 << l;
  if (c->cfg.max_batch_pairs <= 0) c->cfg.max_batch_pairs = log2n_split(l) ? 4096 : 16384;
  if (c->cfg.spectra_cache_bytes == 0) c->cfg.spectra_cache_bytes = (int64_t)48 << 30;
  memset(&c->stats, 0, sizeof(c->stats));
  c->target_total = cfg->target_total;
  cudaError_t ce = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
  if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking);
  if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking);
  if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&c->desc_stream, cudaStreamNonBlocking);
  for (int i = 0; i < 2 && ce == cudaSuccess; i++) {
    ce = cudaEventCreateWithFlags(&c->ev_desc[i], cudaEventDisableTiming);
    if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&c->ev_done[i], cudaEventDisableTiming);
    if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&c->ev_ctr[i], cudaEventDisableTiming);
  }
  if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&c->ev_res, cudaEventDisableTiming);
  if (ce == cudaSuccess)
    for (int i = 0; i < 14 && ce == cudaSuccess; i++) ce = cudaEventCreate(&c->ev[i / 7][i % 7]);
  if (ce == cudaSuccess) ce = upload_tables();
  if (ce == cudaSuccess) {
    int sms = 0;
    ce = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg->device);
    c->fused_grid = 2 * sms;  // pair_fused_kernel: two resident CTAs per SM stride over the pairs
  }
  if (ce != cudaSuccess) {
    delete c;
    return fail(SX_ERR_CUDA, "sx_create: %s", cudaGetErrorString(ce));
  }
  int rc = c->d_ctr[0].ensure(1);
  if (rc == SX_OK) rc = c->d_ctr[1].ensure(1);
  if (rc == SX_OK) rc = c->h_ctr.ensure(1);
  if (rc == SX_OK) rc = c->wn.ensure((size_t)c->N / 2);
  if (rc == SX_OK) {
    std::vector<float2> w((size_t)c->N / 2);
    fill_wn_table(c->log2n, w.data());
    if (cudaMemcpy(c->wn.p, w.data(), sizeof(float2) * w.size(), cudaMemcpyHostToDevice) != cudaSuccess)
      rc = fail(SX_ERR_CUDA, "sx_create: twiddle table upload failed");
  }
  if (rc == SX_OK) rc = c->ent_table.ensure(ent_table_elems(c->log2n));
  if (rc == SX_OK) {
    std::vector<double> et(ent_table_elems(c->log2n));
    fill_ent_table(c->log2n, et.data());
    if (cudaMemcpy(c->ent_table.p, et.data(), sizeof(double) * et.size(), cudaMemcpyHostToDevice) != cudaSuccess)
      rc = fail(SX_ERR_CUDA, "sx_create: entropy table upload failed");
  }
  if (rc == SX_OK && drift_table_elems(c->log2n) > 0) {
    std::vector<float2> dt(drift_table_elems(c->log2n));
    fill_drift_table(c->log2n, dt.data());
    rc = c->drift.ensure(dt.size());
    if (rc == SX_OK && cudaMemcpy(c->drift.p, dt.data(), sizeof(float2) * dt.size(), cudaMemcpyHostToDevice) != cudaSuccess)
      rc = fail(SX_ERR_CUDA, "sx_create: drift table upload failed");
  }
  if (rc != SX_OK) {
    delete c;
    return rc;
  }
  *out = c;
  return SX_OK;
}

extern "C" void sx_destroy(sx_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->cfg.device);
  cudaStreamSynchronize(c->stream);
  if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
  if (c->aux_stream) cudaStreamSynchronize(c->aux_stream);
  if (c->desc_stream) cudaStreamSynchronize(c->desc_stream);
  for (ChunkStore *S : {&c->T, &c->Q})
    for (cudaEvent_t e : S->piece_ev) cudaEventDestroy(e);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->aux_stream) cudaStreamDestroy(c->aux_stream);
  if (c->desc_stream) cudaStreamDestroy(c->desc_stream);
  for (int i = 0; i < 2; i++) {
    if (c->ev_desc[i]) cudaEventDestroy(c->ev_desc[i]);
    if (c->ev_done[i]) cudaEventDestroy(c->ev_done[i]);
    if (c->ev_ctr[i]) cudaEventDestroy(c->ev_ctr[i]);
  }
  if (c->ev_res) cudaEventDestroy(c->ev_res);
  if (c->T.d_bases) cudaFree(c->T.d_bases);
  if (c->Q.d_bases) cudaFree(c->Q.d_bases);
  c->spec.release(); c->planes.release(); c->sbytes.release(); c->meta.release(); c->wn.release(); c->ent_table.release(); c->drift.release();
  c->d_lists[0].release(); c->d_lists[1].release(); c->h_lists[0].release(); c->h_lists[1].release();
  c->d_sigs[0].release(); c->d_sigs[1].release(); for (int i = 0; i < 2; i++) { c->d_prep_flag[i].release(); c->d_prep_went[i].release(); c->d_prep_off[i].release(); c->d_enc_list[i].release(); c->h_fused[i].release(); c->h_enc[i].release(); }
  for (int i = 0; i < 2; i++) { c->d_fused[i].release(); c->d_fail_ctr[i].release(); c->d_fail_pairs[i].release(); c->d_fail_sigs[i].release(); c->d_sps[i].release(); c->d_cand_ref[i].release(); c->d_cand_pool[i].release(); c->d_ctr[i].release(); }
  c->d_fused_scratch.release();
  c->d_scratch.release(); c->d_res.release(); c->d_seg_tap.release(); c->d_table.release(); c->d_tap.release();
  c->h_sigs[0].release(); c->h_sigs[1].release(); c->h_sps[0].release(); c->h_sps[1].release(); c->h_res.release(); c->h_ctr.release();
  for (int i = 0; i < 14; i++)
    if (c->ev[i / 7][i % 7]) cudaEventDestroy(c->ev[i / 7][i % 7]);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

// ------------------------------------------------------------------------------------------------
static int upload_pieces(sx_ctx *c, ChunkStore &S, size_t upto);

static int set_store(sx_ctx *c, ChunkStore &S, const char *bases, const int64_t *offsets, const int32_t *lens,
                     const int32_t *starts, const int32_t *seq_ids, int32_t n, const int32_t *seq_sizes,
                     int32_t n_seqs, int max_len, const char *what) {
  if (n < 0 || (n > 0 && (!bases || !offsets || !lens))) return fail(SX_ERR_ARG, "%s: null argument", what);
  size_t blob = 0;
  for (int i = 0; i < n; i++) {
    if (lens[i] < 0 || lens[i] > max_len)
      return fail(SX_ERR_ARG, "%s: chunk %d has length %d (allowed 0..%d)", what, i, lens[i], max_len);
    if (offsets[i] < 0) return fail(SX_ERR_ARG, "%s: negative offset", what);
    blob = std::max(blob, (size_t)offsets[i] + (size_t)lens[i]);
    if (seq_ids && (seq_ids[i] < 0 || seq_ids[i] >= n_seqs))
      return fail(SX_ERR_ARG, "%s: chunk %d refers to sequence %d of %d", what, i, seq_ids[i], n_seqs);
  }
  if (seq_ids && (!seq_sizes || n_seqs <= 0)) return fail(SX_ERR_ARG, "%s: seq_ids given without seq_sizes", what);
  // allocate first, commit the metadata only when the device buffer exists: a failed call leaves an empty store
  // (and, for targets, no cached spectra) behind, never chunk lists that point at a null blob
  S.n = 0;
  S.async_pending = false;
  if (&S == &c->T) std::fill(c->t_valid.begin(), c->t_valid.end(), 0);
  if (S.d_bases && S.cap_bytes < blob + 32) {  // keep the device buffer across calls when it is big enough
    cudaFree(S.d_bases);
    S.d_bases = nullptr;
    S.cap_bytes = 0;
  }
  if (blob > 0 && !S.d_bases) {
    CU(cudaMalloc((void **)&S.d_bases, blob + 32));
    S.cap_bytes = blob + 32;
  }
  S.n = n;
  S.blob_bytes = blob;
  S.offsets.assign(offsets, offsets + n);
  S.lens.assign(lens, lens + n);
  if (starts) S.starts.assign(starts, starts + n); else S.starts.assign(n, 0);
  if (seq_ids) S.seq_ids.assign(seq_ids, seq_ids + n); else S.seq_ids.assign(n, 0);
  if (seq_sizes && n_seqs > 0) S.seq_sizes.assign(seq_sizes, seq_sizes + n_seqs); else S.seq_sizes.assign(1, 0);
  if (blob > 0) {
    S.async_pending = false;
    if (c->cfg.async_upload) {
      // the first pieces (what the first device batches read) start travelling right away, while the caller is still
      // preparing its next call; upload_pieces() enqueues the rest as the batches ask for them
      CU(cudaStreamSynchronize(c->copy_stream));
      S.piece_bytes = (size_t)16 << 20;
      S.n_pieces = (blob + S.piece_bytes - 1) / S.piece_bytes;
      S.enq_pieces = 0;
      S.h_src = bases;
      while (S.piece_ev.size() < S.n_pieces) {
        cudaEvent_t e;
        CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        S.piece_ev.push_back(e);
      }
      S.async_pending = true;
      const size_t first = ((size_t)2 * (size_t)c->cfg.max_batch_pairs * (size_t)c->cfg.t_chunk) / S.piece_bytes + 1;
      int rc = upload_pieces(c, S, first);
      if (rc != SX_OK) return rc;
    } else {
      CU(cudaMemcpyAsync(S.d_bases, bases, blob, cudaMemcpyHostToDevice, c->stream));
      CU(cudaStreamSynchronize(c->stream));
    }
    c->stats.h2d_bytes += (int64_t)blob;
  }
  return SX_OK;
}

// async_upload: enqueue blob pieces [enq_pieces, upto) on the copy stream
static int upload_pieces(sx_ctx *c, ChunkStore &S, size_t upto) {
  upto = std::min(upto, S.n_pieces);
  for (size_t pi = S.enq_pieces; pi < upto; pi++) {
    const size_t off = pi * S.piece_bytes, sz = std::min(S.piece_bytes, S.blob_bytes - off);
    CU(cudaMemcpyAsync(S.d_bases + off, S.h_src + off, sz, cudaMemcpyHostToDevice, c->copy_stream));
    CU(cudaEventRecord(S.piece_ev[pi], c->copy_stream));
  }
  S.enq_pieces = std::max(S.enq_pieces, upto);
  return SX_OK;
}

// async_upload: everything not yet sent goes now and the host waits for it -- after this the caller's
// buffers are no longer referenced (the contract of sx_config::async_upload)
static int finish_uploads(sx_ctx *c) {
  bool any = false;
  for (ChunkStore *S : {&c->T, &c->Q}) {
    if (!S->async_pending) continue;
    int rc = upload_pieces(c, *S, S->n_pieces);
    if (rc != SX_OK) return rc;
    any = true;
  }
  if (any) CU(cudaStreamSynchronize(c->copy_stream));
  c->T.async_pending = c->Q.async_pending = false;
  return SX_OK;
}

static size_t slot_bytes(const sx_ctx *c) {
  const size_t N = (size_t)c->N;
  return 2 * N * sizeof(float2) + 2 * (N / 32) * sizeof(uint32_t) + N + sizeof(SlotMeta);
}

static int alloc_slots(sx_ctx *c) {
  // persistent region for targets if it fits the budget
  const size_t sb = slot_bytes(c);
  size_t persist = 0;
  if (c->cfg.spectra_cache_bytes > 0 && (size_t)c->T.n * sb <= (size_t)c->cfg.spectra_cache_bytes) persist = (size_t)c->T.n;
  const size_t transient = (size_t)3 * (size_t)c->cfg.max_batch_pairs;
  const size_t total = persist + 2 * transient;
  const size_t N = (size_t)c->N;
  int rc;
  if ((rc = c->spec.ensure(total * 2 * N)) != SX_OK) return rc;
  if ((rc = c->planes.ensure(total * 2 * (N / 32))) != SX_OK) return rc;
  if ((rc = c->sbytes.ensure(total * N)) != SX_OK) return rc;
  if ((rc = c->meta.ensure(total)) != SX_OK) return rc;
  c->n_persist = persist;
  c->n_transient = transient;
  c->t_valid.assign(persist, 0);
  return SX_OK;
}

extern "C" int sx_set_targets(sx_ctx *c, const char *bases, const int64_t *offsets, const int32_t *lens,
                              const int32_t *starts, const int32_t *seq_ids, int32_t n, const int32_t *seq_sizes,
                              int32_t n_seqs) {
  if (!c) return fail(SX_ERR_ARG, "sx_set_targets: null context");
  std::lock_guard<std::mutex> lk(c->mu);
  CU(cudaSetDevice(c->cfg.device));
  // a target chunk holds at most t_chunk = N/2 bases (a query chunk up to N): the scan kernel sizes its per-warp lists
  // for diagonals of at most N/2 target positions
  int rc = set_store(c, c->T, bases, offsets, lens, starts, seq_ids, n, seq_sizes, n_seqs, c->N / 2, "sx_set_targets");
  if (rc != SX_OK) return rc;
  if (c->cfg.target_total <= 0) {  // Slave.cc:405-408
    double tot = 0;
    for (int32_t s : c->T.seq_sizes) tot += (double)s;
    c->target_total = tot;
  }
  return alloc_slots(c);
}

extern "C" int sx_set_queries(sx_ctx *c, const char *bases, const int64_t *offsets, const int32_t *lens,
                              const int32_t *starts, const int32_t *seq_ids, int32_t n, const int32_t *seq_sizes,
                              int32_t n_seqs) {
  if (!c) return fail(SX_ERR_ARG, "sx_set_queries: null context");
  std::lock_guard<std::mutex> lk(c->mu);
  CU(cudaSetDevice(c->cfg.device));
  return set_store(c, c->Q, bases, offsets, lens, starts, seq_ids, n, seq_sizes, n_seqs, c->N, "sx_set_queries");
}

extern "C" int sx_invalidate_spectra(sx_ctx *c) {
  if (!c) return fail(SX_ERR_ARG, "null context");
  std::lock_guard<std::mutex> lk(c->mu);
  std::fill(c->t_valid.begin(), c->t_valid.end(), 0);
  return SX_OK;
}

extern "C" int sx_build_prob_table(double target_total, double *table) {
  // ProbTable::Setup (analysis/ProbTable.cc:15-56): rows p_match = i/511, columns len = 1..2047;
  // smallest identity (27-step bisection) whose probability is NON-ZERO.  Built with
  // the host's libm so that it is bit-identical to what the reference builds.
  if (!table) return fail(SX_ERR_ARG, "sx_build_prob_table: null table");
  const int rows = 512, cols = 2048;
  for (int j = 0; j < cols; j++) table[j] = 0.;
  for (int i = 1; i < rows; i++) {
    double *row = table + (size_t)i * cols;
    const double ident_expect = (double)i / ((double)rows - 1);
    row[0] = 2.;
    for (int j = 1; j < cols; j++) {
      double lo = 0, hi = 1;
      while (hi - lo > 0.00000001) {
        const double mid = (hi + lo) / 2.0;
        // GetMatchProbabilityRaw (ProbTable.cc:143-165); built without FMA contraction (-ffp-contract=off)
        const double s = std::sqrt(ident_expect * (1. - ident_expect) * (double)j);
        const double m = ident_expect * (double)j;
        const double x = (double)j * mid;
        const double cdf = 0.5 * (1. + std::erf((m - x) / s / 1.414213562));
        const double expect = cdf * target_total;
        if (std::exp(-expect) != 0.)
          hi = mid;
        else
          lo = mid;
      }
      row[j] = (hi + lo) / 2.0;
    }
  }
  return SX_OK;
}

extern "C" int sx_set_prob_table(sx_ctx *c, const double *table) {
  if (!c || !table) return fail(SX_ERR_ARG, "sx_set_prob_table: null argument");
  std::lock_guard<std::mutex> lk(c->mu);
  CU(cudaSetDevice(c->cfg.device));
  int rc = c->d_table.ensure((size_t)512 * 2048);
  if (rc != SX_OK) return rc;
  CU(cudaMemcpy(c->d_table.p, table, sizeof(double) * 512 * 2048, cudaMemcpyHostToDevice));
  c->have_table = true;
  c->h_table.assign(table, table + (size_t)512 * 2048);
  c->rm_use_table = -1;  // run_min depends on the table
  return SX_OK;
}

extern "C" int sx_set_profiling(sx_ctx *c, int32_t enabled) {
  if (!c) return fail(SX_ERR_ARG, "null context");
  std::lock_guard<std::mutex> lk(c->mu);
  c->profiling = enabled != 0;
  return SX_OK;
}
extern "C" int sx_stream(sx_ctx *c, void **out) {
  if (!c || !out) return fail(SX_ERR_ARG, "null argument");
  *out = (void *)c->stream;
  return SX_OK;
}
extern "C" int sx_get_stats(sx_ctx *c, sx_stats *out) {
  if (!c || !out) return fail(SX_ERR_ARG, "null argument");
  std::lock_guard<std::mutex> lk(c->mu);
  *out = c->stats;
  return SX_OK;
}
extern "C" int sx_reset_stats(sx_ctx *c) {
  if (!c) return fail(SX_ERR_ARG, "null context");
  std::lock_guard<std::mutex> lk(c->mu);
  memset(&c->stats, 0, sizeof(c->stats));
  return SX_OK;
}

// ------------------------------------------------------------------------------------------------
// One device batch.
// ------------------------------------------------------------------------------------------------
namespace {

struct Batch {
  std::vector<SigDesc> sigs;
  std::vector<SpDesc> sps;
  std::vector<PairReq> pairs;  // batch-local pair index -> chunk indices
  std::vector<uint32_t> pair_list, direct_list;  // strand-pair indices for the two correlation kernels
  std::vector<FusedJob> fused;      // chunk pairs handed to the fused kernel (taken out of pair_list)
  std::vector<uint32_t> enc_list;   // signals the transform kernel still does (when fused is not empty)
  size_t slot_base = 0;  // first slot of this batch's workspace
  size_t transient_used = 0;
  size_t t_need = 0, q_need = 0;  // highest byte of the target / query blob this batch reads (+1)
  std::unordered_map<int32_t, int32_t> tslot;  // target chunk -> slot (transient mode)
  std::unordered_map<int32_t, int32_t> qslot;  // query chunk  -> forward slot (rc slot = +1)
  void clear() {
    sigs.clear(); sps.clear(); pairs.clear(); tslot.clear(); qslot.clear(); pair_list.clear(); direct_list.clear();
    fused.clear(); enc_list.clear();
    transient_used = 0;
    t_need = q_need = 0;
  }
};

// where the records of an align call go: straight into the caller's buffer while they fit; beyond its
// capacity they are only counted, so that the call can report the size it needs (nothing is truncated silently)
struct ResultSink {
  sx_result *out = nullptr;
  int64_t cap = 0, n = 0;
};

struct TapRequest {
  float *sig5n = nullptr;   // host, 5N per signal
  float *xc = nullptr;      // host, N per strand-pair
  std::vector<int32_t> *cands = nullptr;
  std::vector<SegRec> *segs = nullptr;
  int sp_select = -1;  // >= 0: xc / cands / segs of this strand-pair only
  const float *ext_xc = nullptr;  // host, N floats: candidates of strand-pair sp_select come from THIS vector
  double ext_cutoff = 0;
};

}  // namespace

// Largest-safe early-reject bound for the normalised deviation z (see score_counts): the smallest z
// with 0.5*(1+erf(z))*T >= 2*(-ln min_prob), found with the host's libm.  For such z the probability
// is <= min_prob^2, far below min_prob compared with any last-ulp difference between erf/exp
// implementations, so rejecting without evaluating erf/exp cannot change the emitted set.
static double compute_z_cut(double min_prob, double target_total) {
  if (!(min_prob > 0.) || !(min_prob < 1.) || !(target_total > 0.)) return INFINITY;
  const double need = 2. * -std::log(min_prob);
  double lo = -40., hi = 10.;
  if (0.5 * (1. + std::erf(hi)) * target_total < need) return INFINITY;
  for (int it = 0; it < 200; it++) {
    const double mid = 0.5 * (lo + hi);
    if (0.5 * (1. + std::erf(mid)) * target_total >= need) hi = mid; else lo = mid;
  }
  return hi + 1e-9;
}

// Shortest run of passing windows whose segment can still pass the probability filter (ScoreParams::run_min).
// A run of r consecutive passing windows that follows a failing window starts with exactly 19 matches in its
// first window (the 46-window count moves by at most one per position), so its segment -- the union of the r
// windows, 45 + r positions -- holds at most 18 + r matches.  r is prunable when, for EVERY base composition
// (gcT, gcQ) that admits that many matches, score_counts rejects a segment of 45 + r positions with
// min(18 + r, feasible) matches: more matches only lower z / raise the identity, so that is the best case.
// The test mirrors score_counts operation for operation; "z <= z_cut" counts as kept (the exact path decides),
// NaN counts as kept.  Returns the first r that is NOT prunable; every shorter run is never scored.
static int compute_run_min(double z_cut, int use_table, const double *table, int min_len) {
  const int R_MAX = 256;
  for (int r = 1; r < R_MAX; r++) {
    const int len = 45 + r;
    if (len < min_len) continue;  // dropped by -l whatever its score (Slave.cc:172)
    const double dl = (double)len;
    bool keepable = false;
    for (int gt = 0; gt <= len && !keepable; gt++) {
      for (int gq = 0; gq <= len; gq++) {
        const int feasible = std::min(gq, gt) + std::min(len - gq, len - gt);  // matches need equal bases
        const int m = std::min(18 + r, feasible);
        if (m < 19) continue;  // the first window alone holds 19 matches
        const double ident = (double)m / dl;
        const double gc_target = (double)gt / dl;
        const double at_target = 1. - gc_target;
        double rr = (double)gq * gc_target;
        rr = rr + (dl - (double)gq) * at_target;
        const double p_match = rr / dl / 2.;
        if (use_table) {
          const int index = (int)(p_match * 511.0);
          if (index < 1 || index > 511) continue;
          const int l = len >= 2048 ? 2047 : len;
          if (!table || ident >= table[(size_t)index * 2048 + l]) { keepable = true; break; }
        } else {
          const double sd = std::sqrt(p_match * (1. - p_match) * dl);
          const double z = (p_match * dl - dl * ident) / sd / 1.414213562;
          if (!(z > z_cut)) { keepable = true; break; }
        }
      }
    }
    if (keepable) return r;
  }
  return R_MAX;
}

static ScoreParams score_params(const sx_ctx *c) {
  ScoreParams p;
  p.target_total = c->target_total;
  p.min_prob = c->cfg.min_prob;
  p.table_value = c->cfg.prob_table_value;
  p.table = c->d_table.p;
  p.min_len = c->cfg.min_len;
  p.use_table = (c->cfg.use_prob_table && c->have_table) ? 1 : 0;
  p.z_cut = c->z_cut;
  p.run_cap = c->cfg.debug_small_pools ? 8 : 0;  // forces the scan kernel's queue-overflow paths
  p.run_min = (c->cfg.debug_flags & 1) ? 0 : c->run_min;
  return p;
}

static void to_result(const sx_ctx *c, const ResultRec &r, const PairReq &pr, sx_result *o) {
  // FilterMatches coordinate mapping (Slave.cc:176-181, 202-212) and RCQuery (Slave.cc:56-60)
  const ChunkStore &T = c->T, &Q = c->Q;
  const int32_t qid = Q.seq_ids[pr.q], tid = T.seq_ids[pr.t];
  const int32_t qsize = Q.seq_sizes[qid];
  const int32_t tstart = T.starts[pr.t] + r.start_t;
  int32_t qstart;
  if (!r.strand) {
    qstart = Q.starts[pr.q] + r.start_q;
  } else {
    const int32_t chunk = c->cfg.rc_coord_mode == 1 ? Q.lens[pr.q] : c->cfg.q_chunk;
    qstart = r.start_q + qsize - Q.starts[pr.q] - chunk;
  }
  memset(o, 0, sizeof(*o));
  o->query_id = (uint64_t)(int64_t)qid;
  o->target_id = (uint64_t)(int64_t)tid;
  o->query_size = (uint64_t)(int64_t)qsize;
  o->qstart = (uint64_t)(int64_t)qstart;  // int -> unsigned long, sign-extended as in the reference
  o->tstart = (uint64_t)(int64_t)tstart;
  o->len = (uint64_t)(int64_t)r.len;
  o->reverse = (uint8_t)(r.strand ? 1 : 0);
  o->prob = r.prob;
  o->ident = r.ident;
}

namespace {
struct Run {  // one batch on the device: launched asynchronously, completed by batch_finish
  Batch *b = nullptr;
  TapRequest *tap = nullptr;
  int nsig = 0, nsp = 0, n_pairlist = 0, n_direct = 0, n_fused = 0, n_enc = 0;
  bool need_encode = false, need_xcorr = true, active = false;
  bool early_done = false;  // descriptors of the signals uploaded and the encode kernel queued
  bool fused_valid = false;  // the fused kernel queued by batch_launch_early has filled this batch's candidate pool;
                             // the counter block was zeroed there
  unsigned long long n_cand_seen = 0;
  float *d_sig_tap = nullptr, *d_xc_tap = nullptr;
  SegRec *d_seg_tap = nullptr;
  unsigned int seg_tap_cap = 0;
  int stage = 0;             // which pinned descriptor staging buffer holds this batch
  bool staged = false;
  unsigned int fetch_n = 0;  // records copied to the host, waiting for conversion
  bool fetching = false;
  BatchCounters done_ctr;    // counters of the completed batch
};
}  // namespace

// device pools of a batch (per batch in flight); debug_small_pools starts with pools that must overflow
static int ensure_pools(sx_ctx *c, Run &r) {
  const int nsp = r.nsp;
  int rc;
  if ((rc = c->d_sps[r.stage].ensure(std::max(nsp, 1))) != SX_OK) return rc;
  if ((rc = c->d_cand_ref[r.stage].ensure(std::max(nsp, 1))) != SX_OK) return rc;
  if (c->cfg.debug_small_pools) {  // test hook: start with pools that must overflow, so the grow-and-retry paths run
    if (c->d_cand_pool[r.stage].n == 0 && (rc = c->d_cand_pool[r.stage].ensure(64)) != SX_OK) return rc;
  } else if (c->d_cand_pool[r.stage].n < std::max<size_t>((size_t)nsp * 640, 1 << 16)) {
    if ((rc = c->d_cand_pool[r.stage].ensure(std::max<size_t>((size_t)nsp * 640, 1 << 16))) != SX_OK) return rc;
  }
  if (r.n_fused) {
    if ((rc = c->d_fused[r.stage].ensure((size_t)r.n_fused)) != SX_OK) return rc;
    if ((rc = c->d_fail_ctr[r.stage].ensure(2)) != SX_OK) return rc;
    if ((rc = c->d_fail_pairs[r.stage].ensure((size_t)r.n_fused)) != SX_OK) return rc;
    if ((rc = c->d_fail_sigs[r.stage].ensure((size_t)2 * r.n_fused)) != SX_OK) return rc;
    if ((rc = c->d_fused_scratch.ensure((size_t)c->fused_grid * ((size_t)c->N / 2))) != SX_OK) return rc;
  }
  return SX_OK;
}

// The fused kernel of a batch (chunk pairs whose spectra nobody else needs: transforms, product, inverse and FindTop
// in one kernel) and, right behind it, the separate kernels over the pairs it hands back (letters other than A/C/G/T).
// Writes this batch's own candidate pool and counter block.
static int launch_fused(sx_ctx *c, Run &r) {
  cudaStream_t st = c->stream;
  CU(cudaMemsetAsync(c->d_fail_ctr[r.stage].p, 0, 2 * sizeof(unsigned int), st));
  PrepBuf prep = {c->d_prep_flag[r.stage].p, c->d_prep_went[r.stage].p, c->d_prep_off[r.stage].p};
  FusedFail ff = {c->d_fail_ctr[r.stage].p, c->d_fail_ctr[r.stage].p + 1, c->d_fail_pairs[r.stage].p, c->d_fail_sigs[r.stage].p};
  CU(launch_pair_fused(c->log2n, c->d_fused[r.stage].p, r.n_fused, c->d_sigs[r.stage].p, c->d_sps[r.stage].p, c->slots(), prep,
                       c->d_fused_scratch.p, std::min(c->fused_grid, r.n_fused), c->cfg.cutoff, c->cfg.cutoff_fast,
                       c->d_cand_pool[r.stage].p, (unsigned int)std::min<size_t>(c->d_cand_pool[r.stage].n, 0xfffffff0u),
                       c->d_cand_ref[r.stage].p, c->d_ctr[r.stage].p, ff, st));
  c->stats.kernel_launches += 3;
  return SX_OK;
}

// (re)launch the correlation and scan kernels of a batch and the asynchronous read-back of its counter block
// (the encode kernel -- and the fused kernel, the first time round -- have been queued by batch_launch_early)
static int batch_kernels(sx_ctx *c, Run &r) {
  cudaStream_t st = c->stream;
  const bool prof = c->profiling;
  const Slots ws = c->slots();
  const ScoreParams prm = score_params(c);
  cudaEvent_t *ev = c->ev[r.stage];
  if (prof) CU(cudaEventRecord(ev[6], st));
  const bool keep_fused = r.fused_valid && r.need_xcorr;  // first attempt: counters and candidates of the fused pairs stand
  if (!keep_fused) CU(cudaMemsetAsync(c->d_ctr[r.stage].p, 0, sizeof(BatchCounters), st));
  if (r.nsp && r.need_xcorr && r.tap && r.tap->ext_xc) {
    // SeqAnalyzer::MatchUp with the caller's correlation vector: no strand-pair has candidates except the selected
    // one, whose candidates are FindTop of that vector
    const size_t N = (size_t)c->N;
    int rc = c->d_tap.ensure(N);
    if (rc != SX_OK) return rc;
    CU(cudaMemcpyAsync(c->d_tap.p, r.tap->ext_xc, sizeof(float) * N, cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(c->d_cand_ref[r.stage].p, 0, sizeof(uint2) * (size_t)r.nsp, st));
    CU(launch_findtop_external(c->log2n, c->d_tap.p, r.tap->sp_select, r.tap->ext_cutoff, c->d_cand_pool[r.stage].p,
                               (unsigned int)std::min<size_t>(c->d_cand_pool[r.stage].n, 0xfffffff0u), c->d_cand_ref[r.stage].p, c->d_ctr[r.stage].p, st));
    c->stats.kernel_launches += 1;
  } else if (r.nsp && r.need_xcorr) {
    if (r.n_fused && !keep_fused) {  // a re-run after a pool overflow
      int rc = launch_fused(c, r);
      if (rc != SX_OK) return rc;
    }
    CU(launch_xcorr_findtop(c->log2n, c->d_sps[r.stage].p, c->d_lists[r.stage].p, r.n_pairlist, c->d_lists[r.stage].p + r.n_pairlist, r.n_direct, ws,
                            c->cfg.cutoff, c->cfg.cutoff_fast, c->d_cand_pool[r.stage].p,
                            (unsigned int)std::min<size_t>(c->d_cand_pool[r.stage].n, 0xfffffff0u), c->d_cand_ref[r.stage].p, c->d_ctr[r.stage].p,
                            r.d_xc_tap, c->d_scratch.p, (c->cfg.debug_flags & 8) != 0, st));
    const bool three = (c->cfg.debug_flags & 8) != 0 || r.n_pairlist > 0;  // split sizes: half kernels + combine kernel
    c->stats.kernel_launches += log2n_split(c->log2n) ? (three ? 2 : 1) : (r.n_pairlist > 0) + (r.n_direct > 0);
  }
  r.fused_valid = false;  // any further attempt starts from zeroed counters
  if (prof) CU(cudaEventRecord(ev[2], st));
  if (r.nsp) {
    CU(cudaStreamWaitEvent(st, c->ev_res, 0));  // the previous batch's records have left the result pool
    CU(launch_scan_score(c->log2n, c->d_sps[r.stage].p, r.nsp, ws, c->d_cand_pool[r.stage].p, c->d_cand_ref[r.stage].p, prm, c->d_res.p,
                         (unsigned int)std::min<size_t>(c->d_res.n, 0xfffffff0u), r.d_seg_tap, r.seg_tap_cap, c->d_ctr[r.stage].p, st));
    c->stats.kernel_launches += 2;
  }
  if (prof) CU(cudaEventRecord(ev[3], st));
  // the counter block comes back on the auxiliary stream: the compute stream goes straight on to the next batch
  CU(cudaEventRecord(c->ev_done[r.stage], st));
  CU(cudaStreamWaitEvent(c->aux_stream, c->ev_done[r.stage], 0));
  CU(cudaMemcpyAsync(c->h_ctr.p, c->d_ctr[r.stage].p, sizeof(BatchCounters), cudaMemcpyDeviceToHost, c->aux_stream));
  CU(cudaEventRecord(c->ev_ctr[r.stage], c->aux_stream));
  return SX_OK;
}

// Fused pairs (sx_kernels.cu, pair_fused_kernel): a chunk pair whose two spectra nobody else in the batch needs -- the
// target is not cached, target and query each take part in this one pair -- is transformed, multiplied and searched
// for peaks by ONE kernel; its spectra never reach HBM.  That is pair mode (independent chunk pairs, the guided
// refinement pass).  The host proposes (it cannot see the letters); pairs with a chunk that is not pure A/C/G/T are
// handed back by the kernel and done by the separate kernels.
static void propose_fused(sx_ctx *c, Batch &b, bool enable) {
  b.fused.clear();
  b.enc_list.clear();
  for (SigDesc &s : b.sigs) {
    s.g_mode = G_NONE;
    s.g_partner = -1;
  }
  const size_t n_slots = c->n_persist + 2 * c->n_transient;
  if (!enable || b.pair_list.empty() || n_slots == 0) return;
  const int32_t H = c->N / 2;
  // per slot: the signal of this batch that fills it, and how many chunk pairs of the batch read it (entries touched
  // here are reset below; the vectors stay all -1 / 0 between calls)
  if (c->sig_of_slot.size() != n_slots) {
    c->sig_of_slot.assign(n_slots, -1);
    c->slot_uses.assign(n_slots, 0);
  }
  for (size_t i = 0; i < b.sigs.size(); i++) c->sig_of_slot[(size_t)b.sigs[i].slot] = (int32_t)i;
  for (const SpDesc &sp : b.sps) {
    if (sp.flags & SP_REVERSE) continue;  // one forward strand-pair per chunk pair
    c->slot_uses[(size_t)sp.t_slot]++;
    c->slot_uses[(size_t)sp.q_slot]++;
  }
  size_t kept = 0;
  for (const uint32_t spi : b.pair_list) {
    const SpDesc &sp = b.sps[spi];
    const int32_t ti = c->sig_of_slot[(size_t)sp.t_slot], qi = c->sig_of_slot[(size_t)sp.q_slot];
    bool ok = ti >= 0 && qi >= 0 && c->slot_uses[(size_t)sp.t_slot] == 1 && c->slot_uses[(size_t)sp.q_slot] == 1;
    if (ok) {
      const SigDesc &t = b.sigs[(size_t)ti], &q = b.sigs[(size_t)qi];
      ok = t.strand == 0 && q.strand == 0 && t.len > 0 && t.len <= H && q.len > 0 && q.len <= H;
    }
    if (!ok) {
      b.pair_list[kept++] = spi;
      continue;
    }
    FusedJob j;
    j.spi = spi;
    j.t_sig = (uint32_t)ti;
    j.q_sig = (uint32_t)qi;
    b.fused.push_back(j);
    b.sigs[(size_t)ti].g_mode = b.sigs[(size_t)qi].g_mode = G_FUSED;
    b.sigs[(size_t)ti].g_partner = qi;
    b.sigs[(size_t)qi].g_partner = ti;
    // a cacheable target that goes through the fused kernel leaves no spectrum behind: a later batch that needs it
    // transforms it then
    if ((size_t)sp.t_slot < c->n_persist) c->t_valid[(size_t)sp.t_slot] = 0;
  }
  b.pair_list.resize(kept);
  for (const SigDesc &sd : b.sigs) c->sig_of_slot[(size_t)sd.slot] = -1;
  for (const SpDesc &sp : b.sps) c->slot_uses[(size_t)sp.t_slot] = c->slot_uses[(size_t)sp.q_slot] = 0;
  if (!b.fused.empty())
    for (size_t i = 0; i < b.sigs.size(); i++)
      if (b.sigs[i].g_mode != G_FUSED) b.enc_list.push_back((uint32_t)i);
}

// Three-channel pairing (sx_kernels.h): the four channel signals of a pure A/C/G/T chunk sum to zero, so only A, C
// and G are transformed and the G channels of two chunks share one complex transform.  The host proposes the pairs --
// consecutive eligible signals of the batch -- and the device settles them (a chunk with any other letter keeps all
// four channels; the transform kernel looks at the preparation kernel's verdict on both partners).
// A cached target's G must live in a cached slot: when one partner is persistent it is the owner.
static void propose_g_partners(const sx_ctx *c, Batch &b, bool enable) {
  const int32_t H = c->N / 2;
  int open = -1;
  for (size_t i = 0; i < b.sigs.size(); i++) {
    SigDesc &s = b.sigs[i];
    if (s.g_mode == G_FUSED) continue;
    if (!enable || s.len <= 0 || s.len > H) continue;  // either strand: the preparation kernel does both
    if (open < 0) {
      open = (int)i;
      continue;
    }
    int o = open, m = (int)i;
    if ((size_t)b.sigs[m].slot < c->n_persist && (size_t)b.sigs[o].slot >= c->n_persist) std::swap(o, m);
    b.sigs[o].g_mode = G_OWNER;
    b.sigs[o].g_partner = m;
    b.sigs[o].g_pslot = b.sigs[m].slot;
    b.sigs[o].g_plen = b.sigs[m].len;
    b.sigs[m].g_mode = G_MEMBER;
    b.sigs[m].g_partner = o;
    b.sigs[m].g_pslot = b.sigs[o].slot;
    b.sigs[m].g_plen = b.sigs[o].len;
    open = -1;
  }
}

// host part of a launch: copy the descriptors of a batch into pinned staging buffer `stage`
static int batch_stage(sx_ctx *c, Run &r, Batch &b, TapRequest *tap, int stage) {
  r = Run();
  r.b = &b;
  r.tap = tap;
  r.stage = stage;
  const int nsig = r.nsig = (int)b.sigs.size(), nsp = r.nsp = (int)b.sps.size();
  // the three-channel form and the fused kernel both build on the preparation kernel (batch_launch_early)
  const bool prep_on = !(c->cfg.debug_flags & 2) && !(tap && tap->sig5n);
  // (the round-1 three-kernel route of N = 32768, debug_flags bit 3, multiplies the packed four-channel spectra)
  const bool three = prep_on && !(c->cfg.debug_flags & 4) && !(log2n_split(c->log2n) && (c->cfg.debug_flags & 8));
  propose_fused(c, b, three && c->cfg.fuse_pairs == 1 && log2n_fusable(c->log2n) && tap == nullptr && c->fused_grid > 0);
  propose_g_partners(c, b, three);
  int rc;
  if ((rc = c->h_sigs[stage].ensure(std::max(nsig, 1))) != SX_OK) return rc;
  if ((rc = c->h_sps[stage].ensure(std::max(nsp, 1))) != SX_OK) return rc;
  if (nsig) memcpy(c->h_sigs[stage].p, b.sigs.data(), sizeof(SigDesc) * nsig);
  if (nsp) memcpy(c->h_sps[stage].p, b.sps.data(), sizeof(SpDesc) * nsp);
  r.n_pairlist = (int)b.pair_list.size();
  r.n_direct = (int)b.direct_list.size();
  if ((rc = c->h_lists[stage].ensure(std::max(r.n_pairlist + r.n_direct, 1))) != SX_OK) return rc;
  if (r.n_pairlist) memcpy(c->h_lists[stage].p, b.pair_list.data(), sizeof(uint32_t) * r.n_pairlist);
  if (r.n_direct) memcpy(c->h_lists[stage].p + r.n_pairlist, b.direct_list.data(), sizeof(uint32_t) * r.n_direct);
  r.n_fused = (int)b.fused.size();
  r.n_enc = (int)b.enc_list.size();
  if ((rc = c->h_fused[stage].ensure(std::max(r.n_fused, 1))) != SX_OK) return rc;
  if ((rc = c->h_enc[stage].ensure(std::max(r.n_enc, 1))) != SX_OK) return rc;
  if (r.n_fused) memcpy(c->h_fused[stage].p, b.fused.data(), sizeof(FusedJob) * r.n_fused);
  if (r.n_enc) memcpy(c->h_enc[stage].p, b.enc_list.data(), sizeof(uint32_t) * r.n_enc);
  r.staged = true;
  return SX_OK;
}

// First half of launching a batch: upload the signal descriptors and queue the encode kernel.  It only
// writes this batch's own workspace slots (and not yet valid cached target slots), so it may be queued
// while the previous batch is still running -- the device then goes from that batch's scan straight into
// this one's encode while the host reads the previous counters.
static int batch_launch_early(sx_ctx *c, Run &r) {
  const int nsig = r.nsig, nsp = r.nsp;
  if (!r.staged || r.early_done || (nsp == 0 && nsig == 0)) return SX_OK;
  TapRequest *tap = r.tap;
  const size_t N = (size_t)c->N;
  int rc;
  if ((rc = c->d_sigs[r.stage].ensure(std::max(nsig, 1))) != SX_OK) return rc;
  if (tap && (tap->sig5n || tap->xc)) {
    if ((rc = c->d_tap.ensure((size_t)std::max(nsig, 1) * 5 * N + (size_t)std::max(nsp, 1) * N)) != SX_OK) return rc;
    if (tap->sig5n) r.d_sig_tap = c->d_tap.p;
    if (tap->xc) r.d_xc_tap = c->d_tap.p + (size_t)std::max(nsig, 1) * 5 * N;
  }
  cudaStream_t st = c->stream;
  // async_upload: send the blob pieces this batch reads (normally already sent as look-ahead) and make
  // the compute stream wait for the last of them
  for (int which = 0; which < 2; which++) {
    ChunkStore &S = which ? c->Q : c->T;
    const size_t need = which ? r.b->q_need : r.b->t_need;
    if (!S.async_pending || S.n_pieces == 0 || need == 0) continue;
    const size_t upto = std::min((need - 1) / S.piece_bytes + 1, S.n_pieces);
    if ((rc = upload_pieces(c, S, upto)) != SX_OK) return rc;
    CU(cudaStreamWaitEvent(st, S.piece_ev[upto - 1], 0));
  }
  // every descriptor of the batch travels on the auxiliary stream, beside whatever the compute stream is still doing
  // for the previous batch; the compute stream picks them up through an event
  if ((rc = ensure_pools(c, r)) != SX_OK) return rc;
  if ((rc = c->d_lists[r.stage].ensure(std::max(r.n_pairlist + r.n_direct, 1))) != SX_OK) return rc;
  if (r.n_fused && (rc = c->d_enc_list[r.stage].ensure((size_t)std::max(r.n_enc, 1))) != SX_OK) return rc;
  {
    cudaStream_t ax = c->desc_stream;
    int64_t bytes = 0;
    auto up = [&](void *dst, const void *src, size_t n) -> cudaError_t {
      if (n == 0) return cudaSuccess;
      bytes += (int64_t)n;
      return cudaMemcpyAsync(dst, src, n, cudaMemcpyHostToDevice, ax);
    };
    CU(up(c->d_sigs[r.stage].p, c->h_sigs[r.stage].p, sizeof(SigDesc) * nsig));
    if (r.n_fused) CU(up(c->d_enc_list[r.stage].p, c->h_enc[r.stage].p, sizeof(uint32_t) * r.n_enc));
    CU(up(c->d_sps[r.stage].p, c->h_sps[r.stage].p, sizeof(SpDesc) * nsp));
    CU(up(c->d_fused[r.stage].p, c->h_fused[r.stage].p, sizeof(FusedJob) * r.n_fused));
    CU(up(c->d_lists[r.stage].p, c->h_lists[r.stage].p, sizeof(uint32_t) * (r.n_pairlist + r.n_direct)));
    c->stats.h2d_bytes += bytes;
    CU(cudaEventRecord(c->ev_desc[r.stage], ax));
    CU(cudaStreamWaitEvent(st, c->ev_desc[r.stage], 0));
  }
  if (c->profiling) CU(cudaEventRecord(c->ev[r.stage][0], st));
  if (nsig) {
    PrepBuf prep = {nullptr, nullptr, nullptr};
    if (r.d_sig_tap == nullptr && !(c->cfg.debug_flags & 2)) {
      if ((rc = c->d_prep_flag[r.stage].ensure((size_t)nsig)) != SX_OK) return rc;
      if ((rc = c->d_prep_went[r.stage].ensure((size_t)nsig * 256)) != SX_OK) return rc;
      if ((rc = c->d_prep_off[r.stage].ensure((size_t)nsig * 4)) != SX_OK) return rc;
      prep.flag = c->d_prep_flag[r.stage].p;
      prep.went = c->d_prep_went[r.stage].p;
      prep.off = c->d_prep_off[r.stage].p;
    }
    // with fused pairs the transform kernel only does the signals the fused kernel does not take
    const uint32_t *enc_list = r.n_fused ? c->d_enc_list[r.stage].p : nullptr;
    CU(launch_encode_fft(c->log2n, c->d_sigs[r.stage].p, nsig, c->slots(), r.d_sig_tap, prep, enc_list, r.n_enc, st));
    c->stats.kernel_launches += (prep.flag ? 1 : 0) + ((enc_list == nullptr || r.n_enc > 0) ? 1 : 0);
  }
  if (c->profiling) CU(cudaEventRecord(c->ev[r.stage][1], st));
  // the fused kernel fills this batch's own candidate pool: the device goes from the previous batch's scan straight
  // into it
  if (r.n_fused) {
    CU(cudaMemsetAsync(c->d_ctr[r.stage].p, 0, sizeof(BatchCounters), st));
    if ((rc = launch_fused(c, r)) != SX_OK) return rc;
    r.fused_valid = true;
    if (c->profiling) CU(cudaEventRecord(c->ev[r.stage][5], st));
  }
  r.need_encode = nsig > 0;  // only tells batch_wait that this batch had an encode kernel to account for
  r.early_done = true;
  return SX_OK;
}

// Second half: upload the strand-pair descriptors, launch the correlation and scan kernels; returns without
// waiting for the device.  The previous batch must have finished with the shared device pools.
static int batch_launch(sx_ctx *c, Run &r) {
  const int nsig = r.nsig, nsp = r.nsp;
  if (!r.staged || (nsp == 0 && nsig == 0)) return SX_OK;
  TapRequest *tap = r.tap;
  const size_t N = (size_t)c->N;
  int rc;
  if ((rc = batch_launch_early(c, r)) != SX_OK) return rc;
  if (log2n_split(c->log2n) && ((c->cfg.debug_flags & 8) != 0 || r.n_pairlist > 0) &&
      (rc = c->d_scratch.ensure((size_t)std::max(r.n_pairlist + r.n_direct, 1) * N)) != SX_OK)
    return rc;
  if (c->cfg.debug_small_pools) {
    if (c->d_res.n == 0 && (rc = c->d_res.ensure(4)) != SX_OK) return rc;
  } else {
    if (c->d_res.n == 0 && (rc = c->d_res.ensure(std::max<size_t>((size_t)nsp * 8, 1 << 16))) != SX_OK) return rc;
  }
  if (tap && tap->segs) {
    if ((rc = c->d_seg_tap.ensure((size_t)1 << 20)) != SX_OK) return rc;
    r.d_seg_tap = c->d_seg_tap.p;
    r.seg_tap_cap = (unsigned int)c->d_seg_tap.n;
  }
  c->z_cut = c->cfg.use_prob_table ? INFINITY : compute_z_cut(c->cfg.min_prob, c->target_total);
  {
    const int use_table = (c->cfg.use_prob_table && c->have_table) ? 1 : 0;
    if (!(c->rm_z_cut == c->z_cut) || c->rm_use_table != use_table || c->rm_min_len != c->cfg.min_len) {
      c->run_min = compute_run_min(c->z_cut, use_table, c->h_table.empty() ? nullptr : c->h_table.data(), c->cfg.min_len);
      c->rm_z_cut = c->z_cut;
      c->rm_use_table = use_table;
      c->rm_min_len = c->cfg.min_len;
    }
  }
  r.need_xcorr = true;
  r.active = true;
  if ((rc = batch_kernels(c, r)) != SX_OK) return rc;
  // async_upload look-ahead: the next batches' bases travel while this one computes.  Queued AFTER this
  // batch's descriptor copies: the host->device copy engine works in order, bulk pieces ahead of the
  // descriptors would delay the kernels by their transfer time.
  for (int which = 0; which < 2; which++) {
    ChunkStore &S = which ? c->Q : c->T;
    const size_t need = which ? r.b->q_need : r.b->t_need;
    if (!S.async_pending || S.n_pieces == 0) continue;
    const size_t upto = need > 0 ? std::min((need - 1) / S.piece_bytes + 1, S.n_pieces) : 0;
    const size_t ahead = (2 * (size_t)c->cfg.max_batch_pairs * (size_t)c->cfg.t_chunk) / S.piece_bytes + 2;
    if ((rc = upload_pieces(c, S, upto + ahead)) != SX_OK) return rc;
  }
  return SX_OK;
}

// wait for a launched batch, grow-and-retry on pool overflow, then START copying its records to the
// host (asynchronously: the next batch can be launched behind that copy)
static int batch_wait(sx_ctx *c, Run &r) {
  if (!r.active) return SX_OK;
  r.active = false;
  const int nsig = r.nsig, nsp = r.nsp;
  const bool prof = c->profiling;
  int rc;
  for (int attempt = 0; attempt < 8; attempt++) {
    if (attempt > 0 && (rc = batch_kernels(c, r)) != SX_OK) return rc;
    CU(cudaEventSynchronize(c->ev_ctr[r.stage]));
    c->stats.d2h_bytes += (int64_t)sizeof(BatchCounters);
    if (prof) {
      float ms = 0;
      cudaEvent_t *ev = c->ev[r.stage];
      // ev[0] .. ev[1]: preparation + transform kernels; ev[1] .. ev[5]: the fused kernel (queued with them); ev[6] ..
      // ev[2]: separate correlation kernels (and the fused one again when a batch is re-run); ev[2] .. ev[3]: scan
      float tot = 0;
      if (r.need_encode) { cudaEventElapsedTime(&ms, ev[0], ev[1]); c->stats.ms_encode_fft += ms; tot += ms; }
      if (r.need_encode && r.n_fused) { cudaEventElapsedTime(&ms, ev[1], ev[5]); c->stats.ms_xcorr += ms; tot += ms; }
      if (nsp && r.need_xcorr) { cudaEventElapsedTime(&ms, ev[6], ev[2]); c->stats.ms_xcorr += ms; tot += ms; }
      if (nsp) { cudaEventElapsedTime(&ms, ev[2], ev[3]); c->stats.ms_scan_score += ms; tot += ms; }
      c->stats.ms_total += tot;
      if (getenv("SX_GAP_DEBUG") && attempt == 0) {  // where the stream idles between the kernels of consecutive batches
        static double g_a = 0, g_b = 0, g_c = 0;
        static int g_n = 0;
        cudaEvent_t *pv = c->ev[r.stage ^ 1];
        if (r.need_encode && cudaEventElapsedTime(&ms, pv[3], ev[0]) == cudaSuccess && ms > 0 && ms < 50) g_a += ms;
        if (cudaEventElapsedTime(&ms, r.n_fused ? ev[5] : ev[1], ev[6]) == cudaSuccess) g_b += ms;
        if (cudaEventElapsedTime(&ms, ev[0], ev[3]) == cudaSuccess) g_c += ms;
        (void)cudaGetLastError();  // an event that was never recorded (first batch) makes cudaEventElapsedTime fail
        if (++g_n % 64 == 0) fprintf(stderr, "[gap] batches %d: scan(b-1)->start(b) %.3f ms, early->rest %.3f ms, start->end %.3f ms (avg per batch)\n", g_n, g_a / g_n, g_b / g_n, g_c / g_n);
      }
    }
    const BatchCounters ctr = *c->h_ctr.p;
    r.need_encode = false;  // spectra of this batch are in place now
    if (r.need_xcorr) r.n_cand_seen = ctr.cand_used;  // every candidate reserves one pool entry, overflowing or not
    if (ctr.status & ST_CAND_OVERFLOW) {
      // the candidate pool was too small: grow to what the kernel asked for and redo K2+K3
      const size_t want = std::max<size_t>((size_t)ctr.cand_used + (ctr.cand_used >> 2), c->d_cand_pool[r.stage].n * 2);
      if ((rc = c->d_cand_pool[r.stage].ensure(want)) != SX_OK) return rc;
      c->stats.retries++;
      r.need_xcorr = true;
      continue;
    }
    if (ctr.status & ST_RES_OVERFLOW) {
      const size_t want = std::max<size_t>((size_t)ctr.res_used + (ctr.res_used >> 2), c->d_res.n * 2);
      if ((rc = c->d_res.ensure(want)) != SX_OK) return rc;
      c->stats.retries++;
      r.need_xcorr = false;  // candidates are valid; only the scan is repeated
      continue;
    }
    if (ctr.status & ST_INTERNAL) return fail(SX_ERR_CUDA, "scan kernel: internal round limit hit");
    if (ctr.status & ST_TAP_OVERFLOW) return fail(SX_ERR_CAPACITY, "segment tap overflow (%u records)", ctr.seg_tap_used);

    // ---- success: account, start fetching the records ------------------------------------------------
    c->stats.batches++;
    c->stats.signals += nsig;
    c->stats.strand_pairs += nsp;
    c->stats.chunk_pairs += (int64_t)r.b->pairs.size();
    c->stats.fused_pairs += (int64_t)r.n_fused;
    c->stats.candidates += (int64_t)r.n_cand_seen;
    c->stats.segments += (int64_t)ctr.n_segments;
    c->stats.positions += (int64_t)ctr.n_positions;
    c->stats.matches += (int64_t)ctr.res_used;
    r.done_ctr = ctr;
    r.fetch_n = ctr.res_used;
    r.fetching = true;
    if (ctr.res_used) {
      if ((rc = c->h_res.ensure(ctr.res_used)) != SX_OK) return rc;
      CU(cudaMemcpyAsync(c->h_res.p, c->d_res.p, sizeof(ResultRec) * ctr.res_used, cudaMemcpyDeviceToHost, c->aux_stream));
      c->stats.d2h_bytes += (int64_t)(sizeof(ResultRec) * ctr.res_used);
    }
    CU(cudaEventRecord(c->ev[r.stage][4], c->aux_stream));
    CU(cudaEventRecord(c->ev_res, c->aux_stream));
    return SX_OK;
  }
  return fail(SX_ERR_CUDA, "device pools kept overflowing after 8 attempts");
}

// second half of completing a batch: wait for the record copy, convert to t_result, serve taps
static int batch_collect(sx_ctx *c, Run &r, ResultSink *results) {
  if (!r.fetching) return SX_OK;
  r.fetching = false;
  Batch &b = *r.b;
  TapRequest *tap = r.tap;
  const int nsig = r.nsig, nsp = r.nsp;
  const size_t N = (size_t)c->N;
  const BatchCounters ctr = r.done_ctr;
  CU(cudaEventSynchronize(c->ev[r.stage][4]));
  if (r.fetch_n && results) {
    ResultRec *rr = c->h_res.p;
    if (c->cfg.sort_results) {
      // reference emission order: pair, forward before reverse, candidate lag ascending, position ascending
      std::sort(rr, rr + r.fetch_n, [](const ResultRec &a, const ResultRec &b2) {
        if (a.pair != b2.pair) return a.pair < b2.pair;
        if (a.strand != b2.strand) return a.strand < b2.strand;
        if (a.shift != b2.shift) return a.shift < b2.shift;
        return a.start_t < b2.start_t;
      });
    }
    if (results->out && results->n + (int64_t)r.fetch_n <= results->cap)
      for (unsigned int i = 0; i < r.fetch_n; i++) to_result(c, rr[i], b.pairs[rr[i].pair], results->out + results->n + i);
    results->n += (int64_t)r.fetch_n;
  }
  if (tap) {
    if (tap->sig5n && nsig) CU(cudaMemcpy(tap->sig5n, r.d_sig_tap, sizeof(float) * nsig * 5 * N, cudaMemcpyDeviceToHost));
    const int sel = tap->sp_select;
    if (tap->xc && nsp) {
      if (sel >= 0)
        CU(cudaMemcpy(tap->xc, r.d_xc_tap + (size_t)sel * N, sizeof(float) * N, cudaMemcpyDeviceToHost));
      else
        CU(cudaMemcpy(tap->xc, r.d_xc_tap, sizeof(float) * nsp * N, cudaMemcpyDeviceToHost));
    }
    if (tap->cands && nsp) {
      std::vector<uint2> refs(nsp);
      CU(cudaMemcpy(refs.data(), c->d_cand_ref[r.stage].p, sizeof(uint2) * nsp, cudaMemcpyDeviceToHost));
      tap->cands->clear();
      for (int s2 = 0; s2 < nsp; s2++) {
        if (sel >= 0 && s2 != sel) continue;
        std::vector<uint16_t> tmp(refs[s2].y);
        if (refs[s2].y) CU(cudaMemcpy(tmp.data(), c->d_cand_pool[r.stage].p + refs[s2].x, sizeof(uint16_t) * refs[s2].y, cudaMemcpyDeviceToHost));
        for (uint16_t v : tmp) tap->cands->push_back((int32_t)v);
      }
    }
    if (tap->segs) {
      std::vector<SegRec> all(ctr.seg_tap_used);
      if (ctr.seg_tap_used) CU(cudaMemcpy(all.data(), r.d_seg_tap, sizeof(SegRec) * ctr.seg_tap_used, cudaMemcpyDeviceToHost));
      tap->segs->clear();
      for (const SegRec &sr : all)
        if (sel < 0 || sr.sp == sel) tap->segs->push_back(sr);
    }
  }
  return SX_OK;
}

static int run_batch(sx_ctx *c, Batch &b, ResultSink *results, TapRequest *tap) {
  Run r;
  int rc = batch_stage(c, r, b, tap, 0);
  if (rc == SX_OK) rc = batch_launch(c, r);
  if (rc == SX_OK) rc = batch_wait(c, r);
  if (rc == SX_OK) rc = batch_collect(c, r, results);
  const int rc2 = finish_uploads(c);
  return rc != SX_OK ? rc : rc2;
}

// ------------------------------------------------------------------------------------------------
// Batching of pair requests.
// ------------------------------------------------------------------------------------------------
static int32_t target_slot(sx_ctx *c, Batch &b, int32_t t) {
  if (c->n_persist) {
    if (!c->t_valid[t]) {
      SigDesc s;
      s.src = c->T.d_bases + c->T.offsets[t];
      s.len = c->T.lens[t];
      b.t_need = std::max(b.t_need, (size_t)c->T.offsets[t] + (size_t)c->T.lens[t]);
      s.strand = 0;
      s.slot = t;
      s.rc_slot1 = 0;
      b.sigs.push_back(s);
      c->t_valid[t] = 1;
    }
    return t;
  }
  auto it = b.tslot.find(t);
  if (it != b.tslot.end()) return it->second;
  const int32_t slot = (int32_t)(b.slot_base + b.transient_used++);
  SigDesc s;
  s.src = c->T.d_bases + c->T.offsets[t];
  s.len = c->T.lens[t];
  b.t_need = std::max(b.t_need, (size_t)c->T.offsets[t] + (size_t)c->T.lens[t]);
  s.strand = 0;
  s.slot = slot;
  s.rc_slot1 = 0;
  b.sigs.push_back(s);
  b.tslot.emplace(t, slot);
  return slot;
}

// The reverse-strand signal of a chunk is the forward signal reversed with the channels swapped
// (A<->T, C<->G) whenever the entropy windows of both orientations cover the same bases: the chunk
// length is a multiple of the window (N/512), or below 1024 where every weight is 1.  Then
// no reverse spectrum is computed at all: xcorr_pair_kernel derives both strands from the forward one.
static bool rc_derivable(const sx_ctx *c, int32_t len) {
  // N = 32768 reproduces the reference's twiddle drift (sx_kernels.h), which is not symmetric under reversal: the
  // reverse-complement signal gets its own transform there
  if (log2n_split(c->log2n)) return false;
  const int win = c->N / 512;
  return len >= 1 && (len % win == 0 || len < 1024);
}

static int32_t query_slot(sx_ctx *c, Batch &b, int32_t q) {
  auto it = b.qslot.find(q);
  if (it != b.qslot.end()) return it->second;
  const int32_t slot = (int32_t)(b.slot_base + b.transient_used);
  b.transient_used += 2;
  const bool derive = rc_derivable(c, c->Q.lens[q]);
  for (int strand = 0; strand < (derive ? 1 : 2); strand++) {
    SigDesc s;
    s.src = c->Q.d_bases + c->Q.offsets[q];
    s.len = c->Q.lens[q];
    b.q_need = std::max(b.q_need, (size_t)c->Q.offsets[q] + (size_t)c->Q.lens[q]);
    s.strand = strand;
    s.slot = slot + strand;
    s.rc_slot1 = derive ? slot + 2 : 0;  // planes / bytes / meta of the other orientation go to slot + 1
    b.sigs.push_back(s);
  }
  b.qslot.emplace(q, slot);
  return slot;
}

// the two strand-pairs of a chunk pair: consecutive entries of sps, forward first
static void push_pair(sx_ctx *c, Batch &b, int32_t ts, int32_t qs, int32_t qlen, int32_t pidx, int fast) {
  const uint32_t first = (uint32_t)b.sps.size();
  for (int strand = 0; strand < 2; strand++) {
    SpDesc sp;
    sp.t_slot = ts;
    sp.q_slot = qs + strand;
    sp.pair = pidx;
    sp.flags = (strand ? SP_REVERSE : 0) | (fast ? SP_FAST : 0);
    b.sps.push_back(sp);
  }
  if (rc_derivable(c, qlen)) {
    b.pair_list.push_back(first);
  } else {
    b.direct_list.push_back(first);
    b.direct_list.push_back(first + 1);
  }
}

static int align_list_inner(sx_ctx *c, const PairReq *reqs, int64_t n, sx_result *out, int64_t cap, int64_t *n_out) {
  if (c->T.n == 0 && n > 0) return fail(SX_ERR_STATE, "align: no targets loaded");
  if (c->Q.n == 0 && n > 0) return fail(SX_ERR_STATE, "align: no queries loaded");
  // every request is checked before anything is queued: a bad one must not leave half a batch behind
  for (int64_t i = 0; i < n; i++)
    if (reqs[i].t < 0 || reqs[i].t >= c->T.n || reqs[i].q < 0 || reqs[i].q >= c->Q.n)
      return fail(SX_ERR_ARG, "align: pair %lld = (target %d, query %d) out of range", (long long)i, reqs[i].t, reqs[i].q);
  CU(cudaSetDevice(c->cfg.device));
  ResultSink sink;
  sink.out = out;
  sink.cap = cap;
  // Two batch descriptors: while the device works on one, the host assembles AND stages the next;
  // the next batch is launched right behind the asynchronous record copy of the previous one, and
  // only then are those records converted on the host.
  Batch bufs[2];
  Run runs[2];
  bufs[0].slot_base = c->n_persist;
  bufs[1].slot_base = c->n_persist + c->n_transient;
  int cur = 0;
  bool have_prev = false;
  auto submit = [&](int idx) -> int {
    int rc2 = batch_stage(c, runs[idx], bufs[idx], nullptr, idx);  // host only, overlaps the device
    if (rc2 != SX_OK) return rc2;
    if ((rc2 = batch_launch_early(c, runs[idx])) != SX_OK) return rc2;  // encode queued behind the previous scan
    if (have_prev && (rc2 = batch_wait(c, runs[idx ^ 1])) != SX_OK) return rc2;
    if ((rc2 = batch_launch(c, runs[idx])) != SX_OK) return rc2;
    if (have_prev && (rc2 = batch_collect(c, runs[idx ^ 1], &sink)) != SX_OK) return rc2;
    have_prev = true;
    return SX_OK;
  };
  auto drain = [&]() -> int {
    if (!have_prev) return SX_OK;
    int rc2 = batch_wait(c, runs[cur]);
    if (rc2 == SX_OK) rc2 = batch_collect(c, runs[cur], &sink);
    have_prev = false;
    return rc2;
  };
  const size_t maxpairs = (size_t)c->cfg.max_batch_pairs;
  for (int64_t i = 0; i < n; i++) {
    const PairReq &r = reqs[i];
    Batch *b = &bufs[cur];
    // worst case this pair needs 3 fresh transient slots
    if (b->pairs.size() >= maxpairs || b->transient_used + 3 > c->n_transient) {
      int rc = submit(cur);
      if (rc != SX_OK) return rc;
      cur ^= 1;
      b = &bufs[cur];
      b->clear();
    }
    const int32_t ts = target_slot(c, *b, r.t);
    const int32_t qs = query_slot(c, *b, r.q);
    const int32_t pidx = (int32_t)b->pairs.size();
    b->pairs.push_back(r);
    push_pair(c, *b, ts, qs, c->Q.lens[r.q], pidx, r.fast);
  }
  int rc = submit(cur);
  if (rc != SX_OK) return rc;
  rc = drain();
  if (rc != SX_OK) return rc;
  const int64_t total = sink.n;
  if (n_out) *n_out = total;
  if (total > cap || (total > 0 && !out))
    return fail(SX_ERR_CAPACITY, "align: %lld records, buffer holds %lld", (long long)total, (long long)cap);
  return SX_OK;
}

static int align_list(sx_ctx *c, const PairReq *reqs, int64_t n, sx_result *out, int64_t cap, int64_t *n_out) {
  const int rc = align_list_inner(c, reqs, n, out, cap, n_out);
  if (rc != SX_OK && rc != SX_ERR_CAPACITY && rc != SX_ERR_ARG) {
    // a call that failed half way may have marked cached target spectra valid whose encode kernel never ran
    const std::string keep = g_err;
    cudaStreamSynchronize(c->stream);
    std::fill(c->t_valid.begin(), c->t_valid.end(), 0);
    g_err = keep;
  }
  const int rc2 = finish_uploads(c);  // bases no batch asked for still leave the caller's buffers now
  return rc != SX_OK ? rc : rc2;
}

extern "C" int sx_align_pairs(sx_ctx *c, const int32_t *pairs, int64_t n, int32_t fast, sx_result *out, int64_t cap,
                              int64_t *n_out) {
  if (!c || (n > 0 && !pairs)) return fail(SX_ERR_ARG, "sx_align_pairs: null argument");
  std::lock_guard<std::mutex> lk(c->mu);
  std::vector<PairReq> reqs((size_t)n);
  for (int64_t i = 0; i < n; i++) {
    reqs[i].t = pairs[2 * i];
    reqs[i].q = pairs[2 * i + 1];
    reqs[i].fast = fast ? 1 : 0;
  }
  return align_list(c, reqs.data(), n, out, cap, n_out);
}

extern "C" int sx_align_blocks(sx_ctx *c, const sx_pair *blocks, int32_t n_blocks, sx_result *out, int64_t cap,
                               int64_t *n_out) {
  if (!c || (n_blocks > 0 && !blocks)) return fail(SX_ERR_ARG, "sx_align_blocks: null argument");
  std::lock_guard<std::mutex> lk(c->mu);
  std::vector<PairReq> reqs;
  for (int32_t bi = 0; bi < n_blocks; bi++) {
    const sx_pair &p = blocks[bi];
    if (p.target_from < 0 || p.target_to >= c->T.n || p.query_from < 0 || p.query_to >= c->Q.n)
      return fail(SX_ERR_ARG, "sx_align_blocks: block %d ranges t[%d,%d] q[%d,%d] out of bounds", bi, p.target_from,
                  p.target_to, p.query_from, p.query_to);
    // loop order of align_target (Slave.cc:290-299): queries outer, targets inner
    for (int32_t q = p.query_from; q <= p.query_to; q++)
      for (int32_t t = p.target_from; t <= p.target_to; t++) {
        PairReq r;
        r.t = t;
        r.q = q;
        r.fast = p.fast ? 1 : 0;
        reqs.push_back(r);
      }
  }
  return align_list(c, reqs.data(), (int64_t)reqs.size(), out, cap, n_out);
}

// ------------------------------------------------------------------------------------------------
// Stage taps (parity tests).  They run the production kernels on one signal / one strand-pair in
// the transient slot region and copy the intermediate out.
// ------------------------------------------------------------------------------------------------
static int tap_prepare(sx_ctx *c, int32_t target, int32_t query, int32_t strand, int32_t fast, Batch &b) {
  if (target < 0 || target >= c->T.n || query < 0 || query >= c->Q.n || (strand != 0 && strand != 1))
    return fail(SX_ERR_ARG, "tap: (target %d, query %d, strand %d) out of range", target, query, strand);
  if (c->n_transient < 3) return fail(SX_ERR_STATE, "tap: no workspace (call sx_set_targets first)");
  CU(cudaSetDevice(c->cfg.device));
  b.slot_base = c->n_persist;
  b.t_need = c->T.blob_bytes;
  b.q_need = c->Q.blob_bytes;
  // the chunk pair exactly as align_list builds it (both strands, production kernels); the tap then
  // picks the strand-pair it was asked for
  const int32_t base = (int32_t)c->n_persist;
  SigDesc s;
  s.src = c->T.d_bases + c->T.offsets[target];
  s.len = c->T.lens[target];
  s.strand = 0;
  s.slot = base;
  s.rc_slot1 = 0;
  b.sigs.push_back(s);
  b.transient_used = 1;
  const int32_t qs = query_slot(c, b, query);
  PairReq pr;
  pr.t = target;
  pr.q = query;
  pr.fast = fast;
  b.pairs.push_back(pr);
  push_pair(c, b, base, qs, c->Q.lens[query], 0, fast);
  return SX_OK;
}

extern "C" int sx_tap_signal(sx_ctx *c, int32_t is_target, int32_t chunk, int32_t strand, float *out5) {
  if (!c || !out5) return fail(SX_ERR_ARG, "sx_tap_signal: null argument");
  std::lock_guard<std::mutex> lk(c->mu);
  const ChunkStore &S = is_target ? c->T : c->Q;
  if (chunk < 0 || chunk >= S.n || (strand != 0 && strand != 1)) return fail(SX_ERR_ARG, "sx_tap_signal: chunk %d out of range", chunk);
  if (c->n_transient < 1) return fail(SX_ERR_STATE, "tap: no workspace (call sx_set_targets first)");
  CU(cudaSetDevice(c->cfg.device));
  Batch b;
  b.t_need = c->T.blob_bytes;
  b.q_need = c->Q.blob_bytes;
  SigDesc s;
  s.src = S.d_bases + S.offsets[chunk];
  s.len = S.lens[chunk];
  s.strand = strand;
  s.slot = (int32_t)c->n_persist;
  s.rc_slot1 = 0;
  b.sigs.push_back(s);
  TapRequest t;
  t.sig5n = out5;
  return run_batch(c, b, nullptr, &t);
}

extern "C" int sx_tap_xcorr(sx_ctx *c, int32_t target, int32_t query, int32_t strand, float *out) {
  if (!c || !out) return fail(SX_ERR_ARG, "sx_tap_xcorr: null argument");
  std::lock_guard<std::mutex> lk(c->mu);
  Batch b;
  int rc = tap_prepare(c, target, query, strand, 0, b);
  if (rc != SX_OK) return rc;
  TapRequest t;
  t.xc = out;
  t.sp_select = strand;
  return run_batch(c, b, nullptr, &t);
}

extern "C" int sx_tap_candidates(sx_ctx *c, int32_t target, int32_t query, int32_t strand, int32_t fast, int32_t *idx,
                                 int32_t cap, int32_t *n_out) {
  if (!c || !n_out) return fail(SX_ERR_ARG, "sx_tap_candidates: null argument");
  std::lock_guard<std::mutex> lk(c->mu);
  Batch b;
  int rc = tap_prepare(c, target, query, strand, fast, b);
  if (rc != SX_OK) return rc;
  std::vector<int32_t> cands;
  TapRequest t;
  t.cands = &cands;
  t.sp_select = strand;
  if ((rc = run_batch(c, b, nullptr, &t)) != SX_OK) return rc;
  *n_out = (int32_t)cands.size();
  if ((int32_t)cands.size() > cap) return fail(SX_ERR_CAPACITY, "sx_tap_candidates: %zu candidates", cands.size());
  if (idx && !cands.empty()) memcpy(idx, cands.data(), sizeof(int32_t) * cands.size());
  return SX_OK;
}

extern "C" int sx_tap_segments(sx_ctx *c, int32_t target, int32_t query, int32_t strand, int32_t fast, sx_segment *out,
                               int32_t cap, int32_t *n_out) {
  if (!c || !n_out) return fail(SX_ERR_ARG, "sx_tap_segments: null argument");
  std::lock_guard<std::mutex> lk(c->mu);
  Batch b;
  int rc = tap_prepare(c, target, query, strand, fast, b);
  if (rc != SX_OK) return rc;
  std::vector<SegRec> segs;
  TapRequest t;
  t.segs = &segs;
  t.sp_select = strand;
  if ((rc = run_batch(c, b, nullptr, &t)) != SX_OK) return rc;
  std::sort(segs.begin(), segs.end(), [](const SegRec &a, const SegRec &b2) {
    if (a.shift != b2.shift) return a.shift < b2.shift;
    return a.start_t < b2.start_t;
  });
  *n_out = (int32_t)segs.size();
  if ((int32_t)segs.size() > cap) return fail(SX_ERR_CAPACITY, "sx_tap_segments: %zu segments", segs.size());
  for (size_t i = 0; i < segs.size() && out; i++) {
    out[i].start_target = segs[i].start_t;
    out[i].start_query = segs[i].start_t + segs[i].shift;
    out[i].len = segs[i].len;
  }
  return SX_OK;
}

extern "C" int sx_tap_matchup(sx_ctx *c, int32_t target, int32_t query, double cutoff, const float *xc, sx_segment *out,
                              int32_t cap, int32_t *n_out) {
  if (!c || !n_out || !xc) return fail(SX_ERR_ARG, "sx_tap_matchup: null argument");
  std::lock_guard<std::mutex> lk(c->mu);
  Batch b;
  int rc = tap_prepare(c, target, query, 0, 0, b);
  if (rc != SX_OK) return rc;
  std::vector<SegRec> segs;
  TapRequest t;
  t.segs = &segs;
  t.sp_select = 0;  // the query as given; the caller passes the reverse complement itself, as the reference does
  t.ext_xc = xc;
  t.ext_cutoff = cutoff;
  if ((rc = run_batch(c, b, nullptr, &t)) != SX_OK) return rc;
  std::sort(segs.begin(), segs.end(), [](const SegRec &a, const SegRec &b2) {
    if (a.shift != b2.shift) return a.shift < b2.shift;
    return a.start_t < b2.start_t;
  });
  *n_out = (int32_t)segs.size();
  if ((int32_t)segs.size() > cap) return fail(SX_ERR_CAPACITY, "sx_tap_matchup: %zu segments", segs.size());
  for (size_t i = 0; i < segs.size() && out; i++) {
    out[i].start_target = segs[i].start_t;
    out[i].start_query = segs[i].start_t + segs[i].shift;
    out[i].len = segs[i].len;
  }
  return SX_OK;
}
