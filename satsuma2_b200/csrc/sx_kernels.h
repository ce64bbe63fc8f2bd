// Device-side data layout and kernel launchers of libsatsuma_b200 (internal header).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sx {

// ---- HBM layout -------------------------------------------------------------------------------
// A "signal slot" holds everything later kernels need about one chunk in one orientation:
//   spec   [slot][2][N] float2 : spectra of (A + iC) and (G + iT); each as [even bins | odd bins], every
//                                half in the scrambled (DIF) order of the H-point transform (sx_fft.cuh).
//                                THREE-CHANNEL form (pure A/C/G/T chunks, SlotMeta::zmode != 0): the four channel
//                                signals of such a chunk sum to zero sample by sample (weight * (1 - sum of means)),
//                                so T = -(A + C + G) and only three real transforms are needed: spec[slot][0] =
//                                (A + iC) as before, and the G channels of TWO chunks share one complex transform
//                                (G_owner + i G_member) kept in spec[owner slot][1]; the member's spec[.][1] is unused
//   planes [slot][2][N/32] u32 : 2-bit base codes as two bit-planes (lo, hi); A=0 C=1 G=2 T=3
//   bytes  [slot][N] u8        : the oriented bases (only the first len are meaningful)
//   meta   [slot]              : length, flags, the three "quirk" spectrum values
struct SlotMeta {
  int32_t len;
  int32_t flags;  // bit0: contains a byte other than A,C,G,T
  float q_re, q_im, q_nyq;  // sum over channels of the target spectrum at bins H-1 (complex) and H (real)
  int32_t zmode;  // ZM_FOUR: spec[slot][1] = (G + iT); ZM_RE / ZM_IM: three-channel form, the G spectrum is the
                  // real-sequence / imaginary-sequence part of spec[zslot][1]
  int32_t zslot;
  int32_t pad;
};
enum { SLOT_NONACGT = 1 };
enum { ZM_FOUR = 0, ZM_RE = 1, ZM_IM = 2 };

struct Slots {
  float2 *spec;
  uint32_t *planes;
  uint8_t *bytes;
  SlotMeta *meta;
  const float2 *wn;  // e^{-2 pi i n / N}, n < N/2 (per context: the radix-2 split / combine twiddles)
  const double *ent_table;  // entropy terms for integer base counts, fill_ent_table()
  const float2 *drift;      // N = 32768 only: [3][N/4] = rho0, rho1, phi of fill_drift_table(); nullptr otherwise
};

struct SigDesc {  // one chunk signal to encode + transform
  const uint8_t *src;  // chunk bases, forward orientation, device memory
  int32_t len;
  int32_t strand;  // 1: reverse-complement while loading
  int32_t slot;
  int32_t rc_slot1;  // != 0: also write the reverse-complement planes / bytes / meta (no spectrum) to slot
                     // rc_slot1 - 1; its correlation is derived from the forward spectrum (xcorr_pair_kernel)
  // three-channel pairing proposed by the host (honoured only for signals the preparation kernel accepted, i.e. pure
  // A/C/G/T; decided on the device from PrepBuf::flag of both partners): G_OWNER transforms (G_self + i G_partner)
  // into its own spec[.][1], G_MEMBER skips its second transform when its owner is pure too
  int32_t g_mode = 0;      // G_NONE / G_OWNER / G_MEMBER
  int32_t g_partner = -1;  // index of the partner in the same descriptor array (-1: none)
  int32_t g_pslot = 0;     // the partner's slot and length (copied here by the host: one dependent load less at the
  int32_t g_plen = 0;      // start of the transform kernel's CTA)
};
enum { G_NONE = 0, G_OWNER = 1, G_MEMBER = 2, G_FUSED = 3 };  // G_FUSED: handled by pair_fused_kernel (g_partner = the
                                                               // other chunk of the pair); the transform kernel skips it

struct SpDesc {  // one strand-pair = target slot x query slot
  int32_t t_slot, q_slot;
  int32_t pair;   // chunk-pair index inside the batch
  int32_t flags;  // bit0: reverse strand, bit1: fast cutoff
};
enum { SP_REVERSE = 1, SP_FAST = 2 };

struct ResultRec {  // one kept match, chunk-local coordinates
  int32_t pair, strand;
  int32_t start_t, start_q, len, shift;
  double prob, ident;
};

struct SegRec {  // raw segment (tap only)
  int32_t sp, start_t, shift, len;
};

struct ScoreParams {
  double target_total, min_prob, table_value;
  const double *table;  // 512 x 2048 or nullptr
  int32_t min_len;
  int32_t use_table;
  double z_cut;  // early reject when the normalised deviation z exceeds this (+inf disables)
  int32_t run_cap;  // test hook: run ends queued per warp in the scan kernel (0 = the built-in capacity)
  int32_t run_min;  // runs of fewer passing windows (behind a failing one) cannot pass the filter: never scored
};

struct BatchCounters {  // device-side, zeroed per batch
  unsigned int cand_used;     // candidate pool entries reserved
  unsigned int res_used;      // result records reserved
  unsigned int seg_tap_used;  // tap records reserved
  unsigned int status;        // bit0: candidate pool overflow, bit1: result pool overflow, bit2: tap overflow
  unsigned int n_generic;     // strand-pairs with candidates that the byte-wise scan kernel has to do (set by the
                              // bit-parallel one; 0 lets the byte-wise kernel leave at once)
  unsigned int pad_;
  unsigned long long reserved_, n_segments, n_positions;  // n_positions: diagonal positions scanned
};
enum { ST_CAND_OVERFLOW = 1, ST_RES_OVERFLOW = 2, ST_TAP_OVERFLOW = 4, ST_INTERNAL = 16 };

// ---- launchers (sx_kernels.cu); all asynchronous on `stream`, return cudaGetLastError() --------
cudaError_t upload_tables();  // constant/global lookup tables, once per device

bool log2n_supported(int log2n);
bool log2n_split(int log2n);  // transforms handled by two CTAs (N = 32768): launch_xcorr_findtop needs `scratch`,
                              // N float2 per chunk-pair job / single strand-pair job
size_t slot_spec_elems(int log2n);  // float2 elements per slot

// Per-signal results of the preparation kernel (one warp per signal: load, validate, 2-bit planes, entropy weights,
// channel means) that the transform kernel picks up; flag[i] == 0: signal i is prepared inside the transform kernel
// (letters other than A/C/G/T, explicit reverse-strand signals, chunks longer than half the transform)
struct PrepBuf {
  int32_t *flag;  // [nsig]
  float *went;    // [nsig][256] entropy weight per window
  double *off;    // [nsig][4] channel means
};
// enc_list != nullptr: only the n_enc signals listed there are transformed (the preparation kernel still sees all nsig)
cudaError_t launch_encode_fft(int log2n, const SigDesc *sigs, int nsig, Slots ws, float *tap5n, PrepBuf prep,
                              const uint32_t *enc_list, int n_enc, cudaStream_t stream);

// Fused transform + correlation of chunk pairs whose spectra nobody else needs (sx_kernels.cu, pair_fused_kernel).
struct FusedJob {
  uint32_t spi;           // forward strand-pair of the chunk pair (the reverse one is spi + 1)
  uint32_t t_sig, q_sig;  // indices of the target / query signal in the batch's SigDesc array
};
struct FusedFail {  // pairs the fused kernel hands back (a chunk was not pure A/C/G/T); counters zeroed by the caller
  unsigned int *n_pairs, *n_sigs;
  uint32_t *pairs;  // forward strand-pair indices
  uint32_t *sigs;   // signal indices, two per pair
};
bool log2n_fusable(int log2n);
// grid CTAs (<= 2 per SM) stride over the jobs; scratch: grid x N/2 float2.  Also queues the fall-back kernels.
cudaError_t launch_pair_fused(int log2n, const FusedJob *jobs, int njobs, const SigDesc *sigs, const SpDesc *sps, Slots ws,
                              PrepBuf prep, float2 *scratch, int grid, double cutoff, double cutoff_fast, uint16_t *cand_pool,
                              unsigned int pool_cap, uint2 *cand_ref, BatchCounters *ctr, FusedFail fail,
                              cudaStream_t stream);
// pair_list: index of the forward strand-pair of every chunk pair whose reverse strand is derived from the
// forward query spectrum (one CTA does both strands); direct_list: strand-pairs correlated one by one
cudaError_t launch_xcorr_findtop(int log2n, const SpDesc *sps, const uint32_t *pair_list, int n_pairs,
                                 const uint32_t *direct_list, int n_direct, Slots ws, double cutoff,
                                 double cutoff_fast, uint16_t *cand_pool, unsigned int pool_cap,
                                 uint2 *cand_ref, BatchCounters *ctr, float *xc_tap, float2 *scratch,
                                 bool split_three_kernels, cudaStream_t stream);  // split_three_kernels: N = 32768 through
                                 // two half kernels + a combine kernel and the scratch buffer (the round-1 route, kept for
                                 // comparison) instead of one kernel with a cluster of two CTAs per strand-pair
// candidates of strand-pair `spi` from a correlation vector supplied by the caller (N floats, device memory)
cudaError_t launch_findtop_external(int log2n, const float *xc, int spi, double cutoff, uint16_t *cand_pool,
                                    unsigned int pool_cap, uint2 *cand_ref, BatchCounters *ctr, cudaStream_t stream);
void fill_wn_table(int log2n, float2 *host_out);  // N/2 entries
// The reference's float FFT drifts from the exact transform from N = 16384 on: its last passes take their twiddles
// from a float rotation recurrence (extern/RealFFT/OscSinCos.hpp:88-96, FFTReal.hpp:605-656, 783-830).  At N = 32768
// the deviation (2e-4 of max|xc|) exceeds the 1e-4 parity contract, so the kernels reproduce it: the exact spectrum
// of every signal is corrected bin quad by bin quad to what the reference's two oscillator passes give, and the
// product is pre-corrected for the two oscillator passes of its inverse.  Table: 3 x N/4 complex factors.
size_t drift_table_elems(int log2n);  // 0 unless the transform size uses the correction
void fill_drift_table(int log2n, float2 *host_out);
size_t ent_table_elems(int log2n);
void fill_ent_table(int log2n, double *host_out);  // ent_table_elems() entries
cudaError_t launch_scan_score(int log2n, const SpDesc *sps, int nsp, Slots ws, const uint16_t *cand_pool,
                              const uint2 *cand_ref, ScoreParams prm, ResultRec *res_pool,
                              unsigned int res_cap, SegRec *seg_tap, unsigned int seg_tap_cap, BatchCounters *ctr,
                              cudaStream_t stream);

// host copies of the lookup tables (also used by the host-side ProbTable builder)
const double *host_frac_table();    // [128][4]
const uint8_t *host_comp_table();   // [128]
const uint8_t *host_score_table();  // [128][128]

}  // namespace sx
