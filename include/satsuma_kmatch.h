/*
 * satsuma_kmatch.h -- C ABI of the k-mer seeding step of libsatsuma_b200.so.
 *
 * Replaces the reference's KMatch program (kmatch/KMatch.cc), which SatsumaSynteny2 runs before the cross-correlation
 * search (analysis/SatsumaSynteny2.cc:415-434: `KMatch query.fa target.fa K out K K-1 max_freq` for a series of K) and
 * whose output -- raw t_result records (kmatch/matchresult.h == analysis/WorkQueue.h:23-33), prob = ident = 1 -- it
 * loads as seeds.  Same conventions as satsuma_xcorr.h: plain C types, caller-owned buffers, int status, no CPU
 * fallback.
 */
#ifndef SATSUMA_KMATCH_H_
#define SATSUMA_KMATCH_H_

#include "satsuma_xcorr.h"

#ifdef __cplusplus
extern "C" {
#endif

/* KMatch's command line (kmatch/KMatch.cc:320-343): K (odd, <= 31), min_length, max_jump, max_freq */
typedef struct {
  int32_t k;          /* k-mer size; the reference refuses even values (KMatch.cc:325) */
  int32_t max_freq;   /* k-mers occurring more often than this in a genome are dropped (KMatch.cc:121-139) */
  int32_t min_length; /* blocks shorter than this (bases) are not written (KMatch.cc:224) */
  int32_t max_jump;   /* largest gap between consecutive k-mer matches of a block (KMatch.cc:216, 222) */
  int32_t device;     /* CUDA device ordinal */
  int32_t reserved[3];
} sx_kmatch_config;

typedef struct {
  int64_t query_windows, target_windows; /* k-mer windows per genome (incl. those with a letter outside ACGTacgt) */
  int64_t query_kmers, target_kmers;     /* entries left after the frequency filter */
  int64_t kmer_matches;                  /* joined (query position, target position) pairs */
  int64_t blocks;                        /* records emitted */
  double gpu_ms;                         /* upload + every kernel + download, wall clock around the device work */
} sx_kmatch_stats;

void sx_kmatch_default_config(sx_kmatch_config *cfg);
const char *sx_kmatch_last_error(void);

/* One run of KMatch: kmer_array_from_fasta for both genomes (KMatch.cc:20-147), merge_positions (:156-189),
 * dump_matching_blocks (:196-318).  Sequence i of a genome = bases[offsets[i] .. offsets[i] + lens[i]) exactly as the
 * FASTA lines concatenate (case is kept; anything but ACGTacgt breaks the k-mers that contain it); sequence ids in the
 * records are the 0-based order of the sequences.  Records go to out[0 .. *n_out); SX_ERR_CAPACITY with the count
 * needed in *n_out when cap is too small.  The set of records equals the reference's for max_freq = 1 (its default);
 * for larger values the reference's own output depends on the unspecified order std::sort leaves equal query positions
 * in.  stats may be NULL. */
int sx_kmatch(const sx_kmatch_config *cfg, const char *q_bases, const int64_t *q_offsets, const int64_t *q_lens,
              int32_t n_q, const char *t_bases, const int64_t *t_offsets, const int64_t *t_lens, int32_t n_t,
              sx_result *out, int64_t cap, int64_t *n_out, sx_kmatch_stats *stats);

#ifdef __cplusplus
}
#endif
#endif /* SATSUMA_KMATCH_H_ */
