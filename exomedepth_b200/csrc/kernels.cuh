// kernels.cuh — host-callable launchers of the sm_100a kernels (internal; the public boundary is
// include/exomedepth_b200.h).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "edb200_math.cuh"
#include "viterbi_step.h"

namespace edb {

// ---- optional per-kernel timing (edb200_profile): one CUDA event on the launching stream before every kernel ----
struct KernelTimer {
    virtual void mark(const char* name, cudaStream_t st) = 0;   // name == nullptr closes the open interval
};
extern KernelTimer* g_timer;                                    // null unless profiling is on (capi.cu)
inline void prof_mark(const char* name, cudaStream_t st)
{
    if (g_timer) g_timer->mark(name, st);
}

// Per-device one-time kernel attributes (cudaFuncSetAttribute is per device, a process may drive several): `seen` is a
// function-local static of the launcher, one slot per device ordinal.
struct PerDevice {
    size_t v[64] = {};
    // true when `want` exceeds what was configured on the current device so far (and records it)
    bool raise(size_t want)
    {
        int d = 0;
        cudaGetDevice(&d);
        d &= 63;
        if (want <= v[d]) return false;
        v[d] = want;
        return true;
    }
};

// ---- emission -----------------------------------------------------------------------------------
struct CountsView {
    const int32_t* observed;      // [n_samples][obs_stride]
    int64_t obs_stride;
    const int32_t* other;         // reference counts (total = observed + other) or totals
    int64_t other_stride;         // 0: one vector shared by all samples
    int other_is_total;           // 1: `other` already holds total = test + reference (reference API)
};

struct LLView {                   // ll[sample*sample_stride + state*state_stride + bin]
    double* ptr;
    int64_t sample_stride;
    int64_t state_stride;
};

// per (sample, state) constants from per-sample scalar phi / expected
void launch_state_setup(int n_samples, int n_states, const double* phi, const double* expected,
                        const double* odds, StateConst* consts, cudaStream_t st);

// Per-cell log of the GSL errors of a `.Call`-shaped emission (src/error.c:35-52 prints every failing call and carries on):
// event = (bin low word, bin high word << 8 | state, sites of the data term's gsl_sf_lnbeta, sites of the normalising term's)
struct GslEventLog {
    unsigned* count;              // events raised (may exceed cap); null: no log
    uint4* events;
    unsigned cap;
};
// reference-API shaped: per-bin phi / expected vectors of ONE sample (src/CNV_estimate.cpp:52-85)
void launch_emission_bins(const double* phi, const double* expected, const int32_t* total,
                          const int32_t* observed, int64_t n_bins, int n_states, const double* odds,
                          LLView out, unsigned* flags, GslEventLog log, cudaStream_t st);

// gsl_sf_lnbeta (src/beta.c:161-164) through the device's faithful GSL restatement; parity tests only
void launch_lnbeta(const double* x, const double* y, int64_t n, double* out, unsigned* flags, cudaStream_t st);

// batched form of the same: per-bin phi / expected of EVERY sample, [n_samples][pb_stride] (reference-API-faithful cohort)
void launch_emission_bins_batch(CountsView c, const double* phi, const double* expected, int64_t pb_stride, int n_samples,
                                int n_states, int64_t n_bins, const double* odds, LLView out, unsigned* flags, cudaStream_t st);

// batched, per-sample scalar phi/expected, lgamma differences evaluated in registers
void launch_emission_direct(CountsView c, const StateConst* consts, int n_samples, int n_states,
                            int64_t n_bins, LLView out, unsigned* flags, cudaStream_t st);

// batched, per (sample,state) lgamma-difference tables in shared memory + gather
struct TableDims { int K, R, N; };    // entries for observed, other (=total-observed), total
constexpr int kMaxBinRanges = 32;
struct BinRanges {                    // bins [b0[q], b1[q]) for q < n: the whole matrix, or one chromosome group
    int n;
    int64_t b0[kMaxBinRanges], b1[kMaxBinRanges];
};
size_t emission_table_smem_bytes(TableDims d);
// `queue`: one int of device scratch per concurrent launch (the kernel's work-item counter; zeroed on `st` here).
// `lattices`: optional HBM copy of the per-item lattices, n_items * (K + R + N) doubles (K + R + N even);
// lattice_mode 0 = build only, 1 = build and save, 2 = reload (the same items' later bin ranges).
void launch_emission_table(CountsView c, const StateConst* consts, int n_samples, int n_states,
                           const BinRanges& ranges, TableDims dims, LLView out, unsigned* flags, int* queue, int n_sms,
                           double* lattices, int lattice_mode, int* spill, int64_t spill_stride, cudaStream_t st);
// `spill`: parking list of out-of-lattice bins beyond the shared-memory list, emission_table_spill_ctas(...) * spill_stride ints,
// spill_stride >= the bins the ranges cover
int emission_table_max_ctas(int n_sms);

// ---- Viterbi ------------------------------------------------------------------------------------
// One chain template per chromosome; shared by every sample of the batch.
struct ChainDesc {
    int64_t lt_row0;   // log-transition row of observation i is lt_row0 + i   (i = 1 .. nobs-1)
    int64_t em_off;    // emission bin of observation i is em_off + i          (i = 1 .. n_em)
    int64_t out_off;   // path_out index of observation i is out_off + i       (i = out_first .. out_last)
    int32_t nobs;      // observations incl. the two CallCNVs dummies when framed
    int32_t n_em;      // observations 1..n_em read the emission matrix; later ones use the tail constants
    int32_t out_first; // first / last observation written to path_out
    int32_t out_last;
    int32_t call_shift;   // added to start.p / end.p (CallCNVs: -1 for the dummy + per-chromosome shift)
    int32_t pad;
};


struct ViterbiArgs {
    const ChainDesc* chains;      // [n_chains]
    int n_chains;
    const int32_t* chain_list;    // the chains this launch covers (device, [n_list]); null = chains 0 .. n_list-1
    int n_list;
    int max_list_tiles;           // most tiles of any chain in the list (tilemap grid)
    int64_t flat_records;         // > 0: the launch covers ALL chains — the tilemap kernel walks the records 0 .. flat_records-1 as they lie
    int n_samples;
    int n_states;
    int groups;                   // ceil(n_samples / (32 / n_states)): warps per chromosome
    int warps_per_cta;            // sweep warps per CTA: 4 or 8 (viterbi_pick_warps)
    int n_slots;                  // sweep warps in the grid (a multiple of warps_per_cta)
    const int32_t* sched_begin;   // [n_slots + 1] first work item of each sweep warp (viterbi_schedule)
    const int32_t* sched_items;   // [2 * n_items] (chain, group) pairs
    const double* ll;             // emission matrix, see LLView strides; sample stride must be n_states * state stride
    int64_t ll_sample_stride;
    int64_t ll_state_stride;
    const void* ll_map;           // host pointer to the CUtensorMap of ll (2-D: rows = sample x state, box 16 bins x 32/S*S rows, 128B swizzle)
    int perm[kMaxStates];         // HMM state j reads emission column perm[j] (CallCNVs: c(2,1,3))
    const double* lt;             // [rows + tile][lt_pitch] log-transition table (host libm), row = S(j) x pitch/S (k, padded)
    uint32_t* bp;                 // scratch records, viterbi_record_bytes() each: [chain tiles][group][tile]
    const int32_t* bp_tile_base;  // [n_chains] prefix sum of the chains' tile counts
    double tail_other;            // emission of the non-normal states at the dummy last observation (-100)
    int8_t* path;                 // [n_samples][path_stride]
    int64_t path_stride;
    int32_t* chain_calls;         // scratch [n_samples][n_chains][chain_call_cap][4]
    int32_t* chain_ncalls;        // scratch [n_samples][n_chains]
    int chain_call_cap;
    int32_t* calls;               // out [n_samples][call_cap][4]  (start.p, end.p, type, nexons)
    int32_t* ncalls;              // out [n_samples]
    int call_cap;
    unsigned* flags;
    // one-thread-per-chain sweep for CallCNVs-structured transition rows (viterbi_tpc.cu); the schedule then deals
    // (chain, group of 32 samples) items and warps_per_cta counts that kernel's sweep warps
    int tpc;                      // 1: use it
    const StructRow* srows;       // [rows + tile] structured rows (host_tables.h: build_struct_rows)
    double c0, c1;                // log t(0 -> 0), log t(0 -> j > 0)
    const void* ll_map_tpc;       // host pointer to the CUtensorMap with box 16 bins x 32*S rows
    // segmented sweep (viterbi_seam.h): every (chain, 32 samples) line of tiles is cut into pieces that are swept
    // concurrently; sched_items then holds piece ids
    int seg;                      // 1: the schedule deals pieces
    int seg_warm;                 // warm-up tiles in front of a piece that does not start its chain
    const int4* seg_desc;         // [n_pieces] (chain, 32-sample group, first recorded tile, end tile)
    const int32_t* seg_first;     // [n_list * n_g32 + 1] first piece of every line (chain_list order); a line's pieces are consecutive
                                  // (for the statistics; the kernels go by seg_desc)
    double* seam_in;              // [n_pieces][S][32] V of a piece after its warm-up
    double* seam_out;             // [n_pieces][S][32] V of a piece after its last observation
    unsigned* seam_mag;           // [n_pieces][kSeamWords][32] magnitudes and error multipliers of a piece (viterbi_seam.h: PieceErr)
    int4* seg_close;              // [seg_close_cap] decisions with a lead below kSegTau: (piece, sample, observation, destination | lead)
    int seg_close_cap;
    // repair flags, zeroed per run (layout: seg_off_* below)
    int32_t* seg_flags;
    int only_bad;                 // 1: the repair pass — only the lines / chains whose flags are raised
    int seg_force_repair;         // test hook: the check kernel sends every chain to the repair pass
};
// views into ViterbiArgs::seg_flags: [0] listed decisions, [1] any chain refused, then one word per chain, per line
// (chain, 32-sample group) and per (chain, sample) ...
__host__ __device__ inline int seg_n_g32(int n_samples) { return (n_samples + 31) / 32; }
__host__ __device__ inline size_t seg_off_chain(int chain) { return 2 + (size_t)chain; }
__host__ __device__ inline size_t seg_off_line(int n_chains, int n_samples, int chain, int g32) { return 2 + (size_t)n_chains + (size_t)chain * seg_n_g32(n_samples) + g32; }
__host__ __device__ inline size_t seg_off_pair(int n_chains, int n_samples, int chain, int smp)
{
    return 2 + (size_t)n_chains * (1 + seg_n_g32(n_samples)) + (size_t)chain * n_samples + smp;
}
// ... and, as doubles, the bound of |C| per line and lane (viterbi_seam.h: piece_cabs_share), 8-byte aligned
__host__ __device__ inline size_t seg_off_cabs(int n_chains, int n_samples) { return (seg_off_pair(n_chains, n_samples, n_chains, 0) + 1) & ~(size_t)1; }
__host__ __device__ inline size_t seg_flag_ints(int n_chains, int n_samples)
{
    return seg_off_cabs(n_chains, n_samples) + 2 * (size_t)n_chains * seg_n_g32(n_samples) * 32;
}

// enqueues sweep, tilemap, trace and expand for the chains of a.chain_list (the schedule must cover exactly those);
// launch_viterbi_compact then concatenates the per-chromosome call tables of ALL chains.  Both return the number of launches.
int launch_viterbi(const ViterbiArgs& a, cudaStream_t st);
int launch_viterbi_compact(const ViterbiArgs& a, cudaStream_t st);
size_t viterbi_smem_bytes(int n_states, int warps_per_cta);
int launch_viterbi_tpc_sweep(const ViterbiArgs& a, cudaStream_t st);   // viterbi_tpc.cu; 3, 5 or 7 states
int launch_viterbi_seg_check(const ViterbiArgs& a, int n_pieces, cudaStream_t st);   // viterbi_tpc.cu: certifies a segmented sweep (after expand)
size_t viterbi_tpc_smem_bytes(int n_states, int warps_per_cta);
int viterbi_tpc_max_warps(int n_states);
int viterbi_seg_warps(int n_states);      // sweep warps per CTA of the segmented sweep (4, 6 or 8)
int viterbi_pick_warps(const int32_t* chain_nobs, int n_chains, int groups, int n_sms);
int viterbi_lt_pitch(int n_states);      // doubles per table row: S destination rows of S doubles padded to an even count
int viterbi_tile();                      // observations per tile; the table carries this many spare rows
size_t viterbi_record_bytes();           // scratch bytes per (warp, tile): packed back-pointers + tile maps
// tiles a chain spans (tiles follow the 128-byte lines of the emission rows)
inline int viterbi_chain_tiles(const ChainDesc& cd)
{
    if (cd.nobs <= 1) return 0;
    return (int)(((cd.em_off + cd.nobs - 1) >> 4) - ((cd.em_off + 1) >> 4) + 1);
}

// ---- CallCNVs post-processing (callcnvs.cu) -----------------------------------------------------------
// Per-call sums of R/class_definition.R:393-400 and the per-sample cor(test, reference) of :338.
struct CallSummaryArgs {
    int n_samples;
    int n_states;
    int64_t n_bins;
    CountsView counts;            // test counts + reference (or total) counts
    const double* expected;       // [n_samples], or [n_samples][expected_stride] per bin
    int64_t expected_stride;      // 0: one value per sample
    const double* ll;             // emission matrix, see LLView strides
    int64_t ll_sample_stride;
    int64_t ll_state_stride;
    int perm[kMaxStates];         // HMM state j is likelihood column perm[j]; perm[0] is the normal column
    const int32_t* calls;         // [n_samples][call_cap][4] as written by the Viterbi (1-based global bin indices)
    const int32_t* ncalls;        // [n_samples]
    int call_cap;
    double* stats;                // out [n_samples][call_cap][3]: sum(ll_type - ll_normal), sum(total*expected), sum(test); or null
    double* cor;                  // out [n_samples] Pearson correlation of test and reference counts; or null
};
int launch_call_summary(const CallSummaryArgs& a, cudaStream_t st);   // returns the number of launches

// 16-bit counts (65535 = see the overflow list) -> int32 rows, for the bins of `rg`; returns the number of launches
int launch_widen_counts(const uint16_t* src, int64_t src_stride, int32_t* dst, int64_t dst_stride, int n_samples, int64_t n_bins,
                        const BinRanges& rg, const int64_t* ovf_index, const int32_t* ovf_value, int64_t n_overflow, cudaStream_t st);

// 12-bit counts (rows of little-endian 12-bit fields, 4095 = see the overflow list) -> int32 rows, whole rows
int launch_unpack12_counts(const uint8_t* src, int64_t src_stride, int32_t* dst, int64_t dst_stride, int n_samples, int64_t n_bins, cudaStream_t st);

int launch_patch_overflow(int32_t* dst, int64_t dst_stride, int64_t n_bins, const BinRanges& rg, const int64_t* ovf_index,
                          const int32_t* ovf_value, int64_t n_overflow, cudaStream_t st, int64_t s_lo = 0, int64_t s_hi = INT64_MAX);

// ---- select.reference.set correlation sweep (refset.cu) -----------------------------------------------
// z: [n_samples][k_pad] standardised rows over the selected bins (k_pad = n_sel rounded up to 16, zero padded)
void launch_refset_standardize(const int32_t* counts, int64_t stride, int n_samples, const double* bin_length,
                               const int32_t* selected, int64_t n_sel, int64_t k_pad, double* z, cudaStream_t st);
int refset_gram_slices(int m, int n, int64_t k_pad, int n_sms);       // K-slices the Gram kernel is split into
constexpr int kMaxPeers = 16;
struct PeerRows {                     // rows of the B operand: row j lives at base[j / rows_per_rank] + (j % rows_per_rank) * k_pad
    const double* base[kMaxPeers];
    int rows_per_rank;
};
// c[m][n] = za . zb^T clamped to [-1, 1]; partial: scratch of n_slices * m * n doubles
void launch_refset_gram(const double* za, int m, const PeerRows& zb, int n, int64_t k_pad, int n_slices, double* partial,
                        double* c, cudaStream_t st);

// ---- beta-binomial maximum-likelihood fit (betabin.cu) ---------------------------------------------------
// One CTA per sample; dims = caps of the exceedance arrays (observed, reference, total counts).
// info[s] >= 0: Newton iterations used; -1 degenerate sample; -2 negative counts or more than ovf_cap bins beyond the
// caps; -3 iteration cap; -4 binomial limit.
size_t betabin_fit_smem_bytes(TableDims d);
// overflow: device scratch of n_samples * ovf_cap int2 for the bins whose counts exceed the caps
void launch_betabin_fit(CountsView c, int n_samples, int64_t n_bins, TableDims dims, int max_iter, void* overflow, int ovf_cap,
                        double* mu, double* phi, double* loglik, int32_t* info, cudaStream_t st);

// expected Bayes factor of get.power.betabinom (R/tools.R:128-166, the deterministic beta-binomial branch), one CTA per problem
void launch_power_betabinom(const int32_t* size, const double* phi, const double* p, const double* alt_p, int n, double* out, cudaStream_t st);

// ---- forward pass / transition-probability grid (extension, forward.cu) -------------------------------
struct ForwardArgs {
    const ChainDesc* chains;      // [n_chains]
    int n_chains;
    int n_samples;
    int n_states;
    int n_grid;
    const double* ll;             // emission matrix, see LLView strides
    int64_t ll_sample_stride;
    int64_t ll_state_stride;
    int perm[kMaxStates];         // HMM state j reads emission column perm[j]
    const double* decay;          // [rows] exp(-dist / L) per observation (host libm, hmm.cpp:62-64)
    const double* T_grid;         // [n_grid][S*S] transition matrices, column-major T[k + S*j] = P(k -> j)
    double tail_other;            // emission of the non-normal states at the dummy last observation
    double* chain_loglik;         // scratch [n_samples][n_chains][n_grid]
    double* loglik;               // out [n_samples][n_grid]
    int32_t* best;                // out [n_samples] first maximiser over the grid, or null
};
int launch_forward(const ForwardArgs& a, cudaStream_t st);   // returns the number of launches

}  // namespace edb
