// host_tables.h — host-side shared HMM metadata (see host_tables.cpp).
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

#ifdef __CUDACC__
#define EDB_HD __host__ __device__
#else
#define EDB_HD
#endif

namespace edb {
// transition row of destination state j: S doubles padded to an even count, so that rows are 16-byte aligned — and for
// S = 7 to 10, not 8: lane (chain, j) reads row j with 128-bit loads, and rows 64 bytes apart put j = 0, 2, 4, 6 on the
// same four banks (16 wavefronts per load instead of 4; the S = 7 sweep was shared-memory bound, 80 wavefronts per
// warp-step, profiles/r1l_sweep_s7_summary.txt); 80 bytes apart the seven rows, and the four chains' exchange slots, fall
// on distinct banks.  Strides of 4 (S = 3, 4) and 6 doubles (S = 5, 6) are conflict-free as they are.
EDB_HD constexpr int lt_jstride(int S) { return S == 7 ? 10 : S + (S & 1); }
EDB_HD constexpr int lt_pitch(int S) { return S * lt_jstride(S); }          // doubles per observation
// CallCNVs transition matrix for `tp` (R/class_definition.R:343-347), column-major T[k + S*j] = P(k -> j)
void callcnvs_transitions(int S, double tp, double* T);
// lt[i*pitch + j*(pitch/S) + k] = log(t_{k->j} at observation i), i = 1..nobs-1 (src/hmm.cpp:62-79); row 0 and padding zeroed
void build_log_transition_rows(int S, const double* T, const int32_t* pos, int32_t nobs, double L, double* lt, int pitch);
// decay[i] = exp(-(pos[i] - pos[i-1]) / L), i = 1..nobs-1 (src/hmm.cpp:62-64, host libm); decay[0] = 0
void build_decay_rows(const int32_t* pos, int32_t nobs, double L, double* decay);
// CallCNVs framing of one chromosome's positions (R/class_definition.R:368); pos has nb+2 entries. 0 = ok
int frame_positions(int64_t nb, const int32_t* start, const int32_t* end, double L, int32_t* pos);

// Encoders of the count-matrix ingestion layouts (include/exomedepth_b200.h: edb200_batch.observed16 / observed12), on the
// host's threads.  bits = 16: out = uint16 rows, out_stride in elements; bits = 12: out = rows of 12-bit fields, out_stride in
// bytes.  Returns the number of overflow entries the matrix has (only the first `cap` are written), -1 for a negative count.
int64_t pack_counts(int bits, const int32_t* counts, int64_t stride, int32_t n_samples, int64_t n_bins, void* out, int64_t out_stride,
                    int64_t* ovf_index, int32_t* ovf_value, int64_t cap);
// in-place NaN -> -Inf for the device copy of the table: in the recurrence a NaN candidate and a -Inf candidate
// behave identically (neither can satisfy the strict '>' of src/hmm.cpp:81)
void nan_to_neg_inf(double* v, size_t n);
// CallCNVs-structured view of a built table (viterbi_step.h): per observation row the three distance-dependent terms
// b0 = log t(k>0 -> 0), sf = log t(j -> j), ot = log t(k -> j) for k outside {0, j}, plus the two row-independent
// terms c0 = log t(0 -> 0), c1 = log t(0 -> j>0).  Returns 1 when EVERY row of `lt` (device copy: NaN already -Inf)
// has exactly that structure, bit for bit — the one-thread-per-chain sweep is then exact by construction — else 0
// (arbitrary matrices keep the general sweep).  All-zero rows (row 0 of every chain, never read) are skipped.
struct StructRow;
int build_struct_rows(int S, const double* lt, int pitch, int64_t n_rows, StructRow* rows, double* c0, double* c1);
// Placement of the Viterbi sweep's work items (chromosome x group of samples) on the sweep warps: items are
// dealt longest first to the least loaded SM sub-partition (warp slots w and w + 4 of a CTA share one), then to
// its less loaded warp.  The sweep is a latency-bound dependent chain, so the longest chromosomes end up alone
// on their sub-partition.  begin has n_slots + 1 entries, items 2 per work item (chain, group).
void viterbi_schedule(const int32_t* chain_nobs, int n_chains, int groups, int n_ctas, int warps_per_cta,
                      std::vector<int32_t>& begin, std::vector<int32_t>& items);
// Pieces of the segmented sweep (viterbi_seam.h): desc = 4 ints per piece (chain, 32-sample group, first tile, end tile),
// a line's pieces consecutive; first[chain * n_g32 + g] = the line's first piece (n_chains * n_g32 + 1 entries);
// begin / items = the pieces of every sweep warp, 2 ints per item (piece, 0).
void viterbi_cut_pieces(const int32_t* chain_tiles, int n_chains, int n_g32, int n_ctas, int warps_per_cta, int warm, int min_piece,
                        std::vector<int32_t>& begin, std::vector<int32_t>& items, std::vector<int32_t>& desc, std::vector<int32_t>& first);
}  // namespace edb
