/*
 * satsuma_xcorr.h -- C ABI of libsatsuma_b200.so: Satsuma2's chunk-pair cross-correlation
 * hot path on NVIDIA B200 (sm_100a).
 *
 * The reference (bioinfologics/satsuma2) has no plugin/FFI seam for this path: it is a C++
 * class API inside one executable.  This header is the seam a maintainer binds instead.
 * Each entry point cites the reference interface it replaces (paths relative to the
 * reference tree).  Conventions:
 *   - plain C types, caller-owned buffers with capacity + count; nothing is truncated:
 *     when an output buffer is too small the call returns SX_ERR_CAPACITY and *n_out holds
 *     the required number of records;
 *   - every function returns an int status (0 = ok, < 0 = error); no exceptions, no exit();
 *     sx_last_error() gives a message for the calling thread's last failure;
 *   - a context is bound to one CUDA device; calls on one context are serialised by an
 *     internal mutex, different contexts may be driven from different host threads
 *     (the reference runs `-p` worker threads each with private CrossCorrelation /
 *     SeqAnalyzer objects, analysis/HomologyByXCorrSlave.cc:229-231, 302-325);
 *   - there is no CPU fallback: without a usable CUDA device sx_create fails.
 */
#ifndef SATSUMA_XCORR_H_
#define SATSUMA_XCORR_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SX_ABI_VERSION 1

enum {
  SX_OK = 0,
  SX_ERR_ARG = -1,      /* invalid argument / unsupported configuration */
  SX_ERR_CUDA = -2,     /* CUDA runtime failure (no device, launch error, ...) */
  SX_ERR_NOMEM = -3,    /* host or device allocation failed */
  SX_ERR_CAPACITY = -4, /* caller's output buffer too small; *n_out = records needed */
  SX_ERR_STATE = -5     /* call order violated (e.g. align before set_targets) */
};

typedef struct sx_ctx sx_ctx;

/* Run parameters.  These are the globals HomologyByXCorrSlave parses from its flags
 * (analysis/HomologyByXCorrSlave.cc:28-53, 334-384) plus device-side sizing knobs. */
typedef struct {
  int32_t abi_version;     /* = SX_ABI_VERSION */
  int32_t device;          /* CUDA device ordinal */
  int32_t t_chunk;         /* -t_chunk; FFT length N = 2*t_chunk, N in {2048..32768}; a target chunk holds at most
                              t_chunk bases (ChunkManager never makes longer ones), a query chunk at most 2*t_chunk */
  int32_t q_chunk;         /* -q_chunk; <= 2*t_chunk; also the RC coordinate constant (Slave.cc:180) */
  double cutoff;           /* -cutoff      (1.8) */
  double cutoff_fast;      /* -cutoff_fast (2.9), used when sx_pair.fast != 0 (Slave.cc:237) */
  int32_t min_len;         /* -l (0): drop segments shorter than this (Slave.cc:172) */
  int32_t use_prob_table;  /* -prob_table: ProbTable lookup instead of erf/exp (Slave.cc:184-186) */
  double min_prob;         /* effective keep threshold; the slave hard-wires 0.99 (Slave.cc:76) */
  double prob_table_value; /* what ProbTable returns for a "good" match = the -min_prob flag (ProbTable.cc:136) */
  double target_total;     /* sum of all target sequence lengths (Slave.cc:405-408); <=0: derive in sx_set_targets */
  int32_t rc_coord_mode;   /* 0 = slave formula with q_chunk (Slave.cc:56-60,180); 1 = real chunk length
                              (tools/analysis/HomologyByXCorr.cc:173,799) */
  int32_t max_batch_pairs; /* chunk pairs per device batch (0 = default 16384) */
  int64_t spectra_cache_bytes; /* HBM budget for keeping target spectra resident across calls
                                  (0 = default 48 GiB, < 0 = never keep) */
  int32_t sort_results;    /* != 0: return records ordered as the reference emits them */
  int32_t debug_small_pools; /* test hook: start with tiny device pools so the grow-and-retry paths run */
  int32_t async_upload;    /* != 0: sx_set_targets/queries return before the host->device copy has finished;
                              the caller keeps `bases` alive and unmodified until the next sx_align_* call on
                              this context returns.  Batches start as soon as the bases they need have arrived. */
  int32_t debug_flags;     /* test hooks; bit 0: score every raw segment (no run-length pruning in the scan
                              kernel); bit 1: prepare every signal inside the transform kernel (no preparation kernel);
                              bit 2: transform all four channels of every chunk (no three-channel form);
                              bit 3: N = 32768 through the three-kernel route (two half kernels + combine kernel over an
                              HBM scratch buffer) instead of one kernel with a two-CTA cluster per strand-pair */
  int32_t fuse_pairs;      /* 1: chunk pairs whose spectra nobody else in the batch needs (independent pairs, the guided
                              refinement pass) go through ONE kernel -- transforms, product, inverse, peak scan -- and
                              their spectra never reach HBM.  Identical records.  Default 0: on B200 the separate
                              kernels are 3 % faster (both are bound by instruction issue, not by HBM) */
  int32_t reserved[3];
} sx_config;

/* Fills *cfg with the reference's defaults (slave semantics). */
void sx_default_config(sx_config *cfg);

/* One block request == t_pair (analysis/WorkQueue.h:17-22), 28 bytes, same layout:
 * inclusive chunk-index ranges into the flat target/query chunk lists. */
typedef struct {
  int32_t target_from, target_to, query_from, query_to;
  uint8_t fast;
  uint8_t pad0[3];
  int32_t slave_id;
  uint8_t status;
  uint8_t pad1[3];
} sx_pair;

/* One emitted match == t_result (analysis/WorkQueue.h:23-33), 72 bytes, same layout. */
typedef struct {
  uint64_t query_id, target_id, query_size, qstart, tstart, len;
  uint8_t reverse;
  uint8_t pad[7];
  double prob, ident;
} sx_result;

/* Raw diagonal segment == SeqMatch (analysis/CrossCorr.h:165-196) in chunk-local coordinates. */
typedef struct {
  int32_t start_target, start_query, len;
} sx_segment;

/* Counters and per-kernel device time (ms, CUDA events on the launching stream; only
 * accumulated while profiling is enabled). */
typedef struct {
  int64_t chunk_pairs;    /* chunk pairs processed (both strands each) */
  int64_t strand_pairs;
  int64_t signals;        /* chunk signals encoded + transformed */
  int64_t candidates;     /* lags above the FindTop threshold */
  int64_t segments;       /* raw diagonal segments scored */
  int64_t matches;        /* records kept */
  int64_t kernel_launches;
  int64_t batches;
  int64_t h2d_bytes, d2h_bytes;
  int64_t retries;        /* batches re-run because a device pool overflowed */
  double ms_encode_fft;   /* kernel (a)+(b) */
  double ms_xcorr;        /* kernel (c)+(d) */
  double ms_scan_score;   /* kernel (e) */
  double ms_total;        /* first launch to last kernel end, per batch, summed */
  int64_t positions;      /* diagonal positions (base comparisons) scanned by kernel (e) */
  int64_t fused_pairs;    /* chunk pairs handled by the fused transform + correlation kernel (their device time is
                             in ms_xcorr; ms_encode_fft then only holds the preparation kernel) */
} sx_stats;

/* ------------------------------------------------------------------ lifecycle */
int sx_create(const sx_config *cfg, sx_ctx **out);
void sx_destroy(sx_ctx *ctx);
const char *sx_last_error(void);
int sx_abi_version(void);
int sx_device_count(void); /* usable CUDA devices (0 when there is none) */

/* ------------------------------------------------------------------ chunk loading
 * Replaces the chunk vectors `target` / `query` + `targetInfo` / `queryInfo` that the slave
 * fills with ChunkManager::ChunkIt (Slave.cc:388-402; analysis/SeqChunk.cc:72-165).
 * Chunk i = bases[offsets[i] .. offsets[i]+lens[i]) : upper-case IUPAC ASCII exactly as
 * DNAVector holds it (1 byte/base).  starts[i] = SeqChunk::GetStart, seq_ids[i] =
 * SeqChunk::GetID, seq_sizes[id] = ChunkManager::GetSize(id).  Host pointers; the bases are
 * copied to the device.  Reverse complements, signals and spectra are produced on the GPU. */
int sx_set_targets(sx_ctx *ctx, const char *bases, const int64_t *offsets, const int32_t *lens,
                   const int32_t *starts, const int32_t *seq_ids, int32_t n_chunks,
                   const int32_t *seq_sizes, int32_t n_seqs);
int sx_set_queries(sx_ctx *ctx, const char *bases, const int64_t *offsets, const int32_t *lens,
                   const int32_t *starts, const int32_t *seq_ids, int32_t n_chunks,
                   const int32_t *seq_sizes, int32_t n_seqs);
/* Forget cached target spectra (they are rebuilt on next use). */
int sx_invalidate_spectra(sx_ctx *ctx);

/* ------------------------------------------------------------------ the hot path
 * sx_align_blocks: replaces `void HomologyByXCorr::align_target(t_pair p)`
 * (Slave.cc:270-300) for n block requests: every target chunk x every query chunk of each
 * block, both orientations (Align, Slave.cc:221-253), filtered and mapped to sequence
 * coordinates (FilterMatches, Slave.cc:168-219).  Records are written to out[0..*n_out). */
int sx_align_blocks(sx_ctx *ctx, const sx_pair *blocks, int32_t n_blocks, sx_result *out, int64_t cap,
                    int64_t *n_out);
/* sx_align_pairs: the same for an explicit list of (target chunk, query chunk) index pairs
 * (pairs = n x 2 int32) -- the inner `Align` call for independent pairs. */
int sx_align_pairs(sx_ctx *ctx, const int32_t *pairs, int64_t n_pairs, int32_t fast, sx_result *out,
                   int64_t cap, int64_t *n_out);

/* ------------------------------------------------------------------ several GPUs behind one handle
 * The reference spreads the chunk-pair grid over processes by target range: `-nblocks/-block`
 * (analysis/SeqChunk.cc:104-116) and N slaves fed by the master (analysis/WorkQueue.cc:290-312).  sx_multi does that
 * split in one process: the flat target chunk list is cut into shard_world * n_devices contiguous ranges, this handle
 * owns n_devices of them (those of shard_rank), one per GPU.  Every GPU is sent only the bases of its target range
 * (spectra cached in its HBM) and, per call, only the query chunks its share of the blocks touches; blocks that
 * straddle a range boundary are split between the neighbours; every GPU works on its own host thread; the records of
 * all GPUs are gathered into the caller's one buffer.  No collective: GPUs never exchange data.
 * devices == NULL or n_devices <= 0: every visible GPU.  shard_rank / shard_world = 0 / 1 for a single process; with
 * shard_world > 1 several processes (ranks, slaves) own disjoint parts of the same split and the union of their
 * outputs is the unsharded result.  Chunk indices in sx_pair stay those of the caller's full lists.  The caller keeps
 * the query `bases` alive while the handle may still fetch from them (until destroy / the next sx_multi_set_queries). */
typedef struct sx_multi sx_multi;
int sx_multi_create(const sx_config *cfg, const int32_t *devices, int32_t n_devices, int32_t shard_rank,
                    int32_t shard_world, sx_multi **out);
void sx_multi_destroy(sx_multi *m);
const char *sx_multi_last_error(void);
int32_t sx_multi_device_count(const sx_multi *m);
int sx_multi_target_range(const sx_multi *m, int32_t shard, int32_t *t_lo, int32_t *t_hi); /* chunks [t_lo, t_hi) of GPU `shard` */
int sx_multi_set_targets(sx_multi *m, const char *bases, const int64_t *offsets, const int32_t *lens,
                         const int32_t *starts, const int32_t *seq_ids, int32_t n_chunks, const int32_t *seq_sizes,
                         int32_t n_seqs);
int sx_multi_set_queries(sx_multi *m, const char *bases, const int64_t *offsets, const int32_t *lens,
                         const int32_t *starts, const int32_t *seq_ids, int32_t n_chunks, const int32_t *seq_sizes,
                         int32_t n_seqs);
int sx_multi_set_prob_table(sx_multi *m, const double *table);
int sx_multi_invalidate_spectra(sx_multi *m);
/* align_target (Slave.cc:270-300) for n blocks over all GPUs of the handle; records gathered into out[0..*n_out) */
int sx_multi_align_blocks(sx_multi *m, const sx_pair *blocks, int32_t n_blocks, sx_result *out, int64_t cap,
                          int64_t *n_out);
int sx_multi_get_stats(sx_multi *m, int32_t shard, sx_stats *out); /* shard < 0: counters summed, times of the slowest GPU */
int sx_multi_reset_stats(sx_multi *m);
int sx_multi_stream(sx_multi *m, int32_t shard, void **stream_out);

/* ------------------------------------------------------------------ stage taps (parity tests)
 * strand: 0 = query as given, 1 = reverse-complemented query (DNAVector::ReverseComplement). */
/* CCSignal::SetSequence (analysis/CrossCorr.cc:179-206): out5 = entropy[N],A[N],C[N],G[N],T[N] */
int sx_tap_signal(sx_ctx *ctx, int32_t is_target, int32_t chunk, int32_t strand, float *out5);
/* CrossCorrelation::CrossCorrelate(out, target, query) (CrossCorr.cc:386-403): out[N] */
int sx_tap_xcorr(sx_ctx *ctx, int32_t target, int32_t query, int32_t strand, float *out);
/* SeqAnalyzer::FindTop (CrossCorr.cc:878-944): ascending lag indices */
int sx_tap_candidates(sx_ctx *ctx, int32_t target, int32_t query, int32_t strand, int32_t fast,
                      int32_t *idx, int32_t cap, int32_t *n_out);
/* SeqAnalyzer::MatchUp (CrossCorr.cc:583-605): raw segments before probability filtering,
 * ordered by (candidate lag, position) as the reference emits them */
int sx_tap_segments(sx_ctx *ctx, int32_t target, int32_t query, int32_t strand, int32_t fast,
                    sx_segment *out, int32_t cap, int32_t *n_out);

/* SeqAnalyzer::MatchUp(out, query, target, xc) with the CALLER's correlation vector (CrossCorr.cc:583-605): FindTop
 * of xc[0..N) at `cutoff` (SeqAnalyzer::SetTopCutoff), then the diagonal scan of every candidate; the query chunk is
 * taken as given (the reference's callers pass the reverse complement themselves).  Same order as sx_tap_segments. */
int sx_tap_matchup(sx_ctx *ctx, int32_t target, int32_t query, double cutoff, const float *xc, sx_segment *out,
                   int32_t cap, int32_t *n_out);

/* ------------------------------------------------------------------ measurement */
int sx_set_profiling(sx_ctx *ctx, int32_t enabled); /* per-kernel CUDA-event timing on/off */
int sx_get_stats(sx_ctx *ctx, sx_stats *out);
int sx_reset_stats(sx_ctx *ctx);
/* The CUDA stream (cudaStream_t) every kernel and copy of this context is issued on, so a caller
 * can bracket calls with its own CUDA events. */
int sx_stream(sx_ctx *ctx, void **stream_out);
/* Host-side ProbTable::Setup (analysis/ProbTable.cc:15-56) with libm: fills 512*2048 doubles. */
int sx_build_prob_table(double target_total, double *table);
/* Installs the 512x2048 table used when use_prob_table != 0 (copied to the device). */
int sx_set_prob_table(sx_ctx *ctx, const double *table);

#ifdef __cplusplus
}
#endif
#endif /* SATSUMA_XCORR_H_ */
