// C++ host side above the C ABI (include/satsuma_xcorr.h): what stays on the host in Satsuma2's
// slave -- FASTA loading, chunking, block requests, result records, the binary match file -- with
// the reference's behaviour, so that its drivers can switch to the B200 path.
//
//   read_fasta        <- vecDNAVector::Read / ReadOne        (analysis/DNAVector.cc:1175-1294)
//   chunk_sequences   <- ChunkManager::ChunkItSelect         (analysis/SeqChunk.cc:72-165)
//   HomologyByXCorr   <- class HomologyByXCorr (slave)       (analysis/HomologyByXCorrSlave.cc:62-325)
//   MatchFile         <- SingleMatch / MultiMatches::Write   (analysis/SequenceMatch.cc:49-64, 320-360)
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/satsuma_xcorr.h"

namespace sxh {

typedef sx_pair t_pair;      // analysis/WorkQueue.h:17-22
typedef sx_result t_result;  // analysis/WorkQueue.h:23-33

struct Sequence {
  std::string name;   // header tokens joined by '_', leading '>' dropped
  std::string bases;  // upper-cased (FASTA); as written (FASTQ, keep_case)
  bool keep_case = false;
};

// Comma-separated list of FASTA files, concatenated.  Lines are split on blanks/tabs; a header's
// tokens are joined with '_'; of a sequence line only the first token counts; bases are upper-cased.
// A file whose first record line starts with '@' is read as FASTQ (vecDNAVector::ReadQ).
bool read_fasta(const std::string &files, std::vector<Sequence> &out, std::string *err);

struct ChunkList {
  std::string blob;              // all sequences back to back (chunks index into it, overlaps share bytes)
  std::vector<int64_t> offsets;  // per chunk
  std::vector<int32_t> lens, starts, seq_ids;
  std::vector<int32_t> seq_sizes;  // ChunkManager::GetSize(seq)
  std::vector<std::string> names;  // per sequence
  int32_t n() const { return (int32_t)lens.size(); }
};

// chunk j of a sequence of length l >= 6 spans [j*stride, min((j+1)*stride + overlap, l)),
// stride = size - overlap, 1 + l/stride chunks; all-N/X chunks are emptied but keep their index;
// with n_blocks > 0 only chunks of block `my_block` keep their bases (-nblocks/-block).
void chunk_sequences(const std::vector<Sequence> &seqs, int size, int overlap, int n_blocks, int my_block,
                     ChunkList &out);

struct MatchFile;
// Guided refinement (tools/analysis/HomologyByXCorr.cc:206-330, SetGuideChunks): between every two consecutive
// matches of a chained match list that share target, query and orientation, the gap (with 32-base laps into both
// matches) is cut into `pieces = 1 + max(gap_t, gap_q) / size` pieces per side; piece k of the target list is paired
// with piece k of the query list (same index in both lists), and `orientation[k]` (+1 forward only, -1 reverse
// only) is forced.  Every quirk of the reference is kept: gaps above 100 000 or below 20 bases are skipped, the
// iterators advance by HALF of the OTHER side's piece length, reverse matches take their query window from the
// far end of the forward sequence.
void guide_chunks(const std::vector<Sequence> &targets, const std::vector<Sequence> &queries, const MatchFile &chained,
                  int size, ChunkList &t_out, ChunkList &q_out, std::vector<int> &orientation);

// Slave-side driver with the reference's interface: load once, then align_target(t_pair).
class HomologyByXCorr {
 public:
  struct Options {
    int device = 0;
    int n_gpus = 0;  // > 0: drive this many GPUs (devices 0 .. n-1) through sx_multi, target list split by range
    int t_chunk = 4096, q_chunk = 4096;
    double cutoff = 1.8, cutoff_fast = 2.9;
    int min_len = 0;
    bool prob_table = false;
    double min_prob_flag = 0.9999;  // -min_prob: only the value ProbTable returns (slave) / the filter (standalone)
    bool standalone_semantics = false;  // tools/analysis/HomologyByXCorr: filter at -min_prob, RC coordinate by chunk length
    int max_batch_pairs = 0;
    bool sort_results = false;
    double target_total = 0;  // > 0: instead of the sum of the target sequence lengths (guided mode: t_chunk)
  };
  HomologyByXCorr();
  ~HomologyByXCorr();
  // returns false and sets error() on failure
  bool init(const Options &opt, const ChunkList &target, const ChunkList &query);
  // replaces HomologyByXCorr::align_target (Slave.cc:270-300): appends t_result records
  bool align_target(const t_pair &p, std::vector<t_result> &results);
  bool align_targets(const t_pair *p, int n, std::vector<t_result> &results);
  double target_total() const { return target_total_; }
  const std::string &error() const { return err_; }
  bool stats(sx_stats *s) const;

 private:
  sx_ctx *ctx_ = nullptr;
  sx_multi *multi_ = nullptr;
  std::string query_blob_;  // kept for sx_multi, which fetches query ranges when a call needs them
  double target_total_ = 0;
  std::string err_;
};

// xcorr match file, version 3 (MultiMatches::Write):
//   int32 ver=3; int32 nT; nT x {int64 len incl. NUL; bytes}; int32 nQ; nQ x {...}; int32 n;
//   n x {int32 tID,qID,qLen,startT,startQ,len,rc; double matches(=ident*len), prob, ident};
//   int32 tSize[nT]; int32 qSize[nQ]
struct MatchFile {
  std::vector<std::string> target_names, query_names;
  std::vector<int32_t> target_sizes, query_sizes;
  std::vector<t_result> matches;
  // SingleMatch::m_matches per record (= ident * len when the slave record is turned into a match; it is NOT
  // updated when Collapse changes the length).  Empty, or shorter than `matches`: ident * len is written.
  std::vector<double> n_matches;
  bool write(const std::string &path, std::string *err) const;
  bool read(const std::string &path, std::string *err);
  // MultiMatches::Sort (analysis/SequenceMatch.h:211-215): std::sort with SingleMatch::operator< (:93-109):
  // target id, query id, forward before reverse, start in target, length.  Same comparator, same algorithm.
  void sort();
  // MultiMatches::Collapse (analysis/SequenceMatch.cc:418-469) on a sorted list, quirks included: neighbours
  // are fused when they share the TARGET id and orientation and start within 3 bases of each other in both
  // sequences (the query id is not compared); the fused match keeps the first one's scores; and the last
  // group is dropped (the reference never stores its running match after the loop).
  void collapse();
  // RunMatchDynProg (analysis/MatchDynProg.cc:401-561; Chain :199-243; penalties :24-112): the synteny chain
  // through a sorted + collapsed match list, per target sequence: per-base coverage counts give every match a
  // repeat score, matches in heavily covered places are dropped, a forward DP over the matches ordered by their
  // start in the target (look-ahead 2000 matches / 250 kb) minimises transition + skip + match penalties, and the
  // best chain is traced back from the last match.  ChainMatches = sort(); collapse(); chain().
  void chain(MatchFile &out) const;
  // RunMatchDynProgMult (analysis/MatchDynProg.cc:245-399), `-dups 1`: chain once, then for every target drop the
  // matches of the query that dominates its first chain, chain what is left, and return both chains, sorted.
  void chain_dups(MatchFile &out) const;
};

}  // namespace sxh
