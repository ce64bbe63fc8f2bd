// Class-level seams of the reference on top of libsatsuma_b200: drop-in classes with the
// reference's method signatures for callers that use the cross-correlation objects directly instead of
// HomologyByXCorr::align_target --
//   sx_shim::CCSignal::SetSequence(const DNAVector &, int)                      (analysis/CrossCorr.h:16)
//   sx_shim::CrossCorrelation::CrossCorrelate(vector<float> &, one, two)         (analysis/CrossCorr.h:151)
//   sx_shim::SeqAnalyzer::MatchUp(vecSeqMatch &, query, target, vector<float> &) (analysis/CrossCorr.h:261)
// Include AFTER the reference's "analysis/CrossCorr.h" (DNAVector, SeqMatch, vecSeqMatch come from there) and link
// -lsatsuma_b200.  These are convenience seams for debugging and small callers: every call is its own device round
// trip (one chunk, one chunk pair); throughput lives in sx_align_blocks.
#pragma once
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "satsuma_xcorr.h"

namespace sx_shim {

inline void check(int rc) {
  if (rc != SX_OK) throw std::runtime_error(std::string("libsatsuma_b200: ") + sx_last_error());
}
// one context per transform size, created on first use (device: SX_DEVICE or 0)
inline sx_ctx *context(int size) {
  static std::map<int, sx_ctx *> ctxs;
  auto it = ctxs.find(size);
  if (it != ctxs.end()) return it->second;
  sx_config cfg;
  sx_default_config(&cfg);
  const char *dev = getenv("SX_DEVICE");
  cfg.device = dev ? atoi(dev) : 0;
  cfg.t_chunk = size / 2;
  cfg.q_chunk = size;  // both signals are sized `size`; a query may fill all of it
  cfg.target_total = 1.;
  sx_ctx *c = nullptr;
  check(sx_create(&cfg, &c));
  ctxs[size] = c;
  return c;
}
inline void load(sx_ctx *c, bool target, const std::string &bases) {
  const int64_t off = 0;
  const int32_t len = (int32_t)bases.size(), zero = 0, size = len;
  static const char empty[1] = {0};
  check((target ? sx_set_targets : sx_set_queries)(c, len ? bases.data() : empty, &off, &len, &zero, &zero, 1, &size, 1));
}

class CCSignal {
 public:
  CCSignal() : m_size(0) {}
  virtual ~CCSignal() {}
  virtual void SetSequence(const DNAVector &b, int size) {
    m_bases.resize((size_t)b.size());
    for (int i = 0; i < (int)b.size(); i++) m_bases[(size_t)i] = b[i];
    m_size = size;
    sx_ctx *c = context(size);
    load(c, true, m_bases);  // the workspace of a context is sized when targets are set
    std::vector<float> all((size_t)5 * size);
    check(sx_tap_signal(c, 1, 0, 0, all.data()));
    for (int ch = 0; ch < 4; ch++) m_ch[ch].assign(all.begin() + (size_t)(ch + 1) * size, all.begin() + (size_t)(ch + 2) * size);
  }
  virtual int GetFullSize() const { return m_size; }
  virtual int GetCount() const { return 4; }
  virtual const std::vector<float> &Get(int i) const { return m_ch[i >= 0 && i < 4 ? i : 0]; }
  const std::string &Bases() const { return m_bases; }

 private:
  std::string m_bases;
  int m_size;
  std::vector<float> m_ch[4];
};

class CrossCorrelation {
 public:
  // out = sum over the four channels of DoOne(one, two): `one` is the target signal, `two` the query signal
  void CrossCorrelate(std::vector<float> &out, const CCSignal &one, const CCSignal &two) {
    const int size = one.GetFullSize();
    sx_ctx *c = context(size);
    load(c, true, one.Bases());
    load(c, false, two.Bases());
    out.assign((size_t)size, 0.f);
    check(sx_tap_xcorr(c, 0, 0, 0, out.data()));
  }
};

class SeqAnalyzer {
 public:
  SeqAnalyzer() : m_topCutoff(1.8) {}
  void SetTopCutoff(double c) { m_topCutoff = c; }
  void MatchUp(vecSeqMatch &out, const DNAVector &query, const DNAVector &target, std::vector<float> &xc) {
    sx_ctx *c = context((int)xc.size());
    std::string q((size_t)query.size(), 0), t((size_t)target.size(), 0);
    for (int i = 0; i < (int)query.size(); i++) q[(size_t)i] = query[i];
    for (int i = 0; i < (int)target.size(); i++) t[(size_t)i] = target[i];
    load(c, true, t);
    load(c, false, q);
    std::vector<sx_segment> segs((size_t)1 << 16);
    int32_t n = 0;
    int rc = sx_tap_matchup(c, 0, 0, m_topCutoff, xc.data(), segs.data(), (int32_t)segs.size(), &n);
    if (rc == SX_ERR_CAPACITY) {
      segs.resize((size_t)n);
      rc = sx_tap_matchup(c, 0, 0, m_topCutoff, xc.data(), segs.data(), (int32_t)segs.size(), &n);
    }
    check(rc);
    out.clear();
    for (int32_t i = 0; i < n; i++) out.push_back(SeqMatch(segs[(size_t)i].start_target, segs[(size_t)i].start_query, segs[(size_t)i].len, 0.));
  }

 private:
  double m_topCutoff;
};

}  // namespace sx_shim
