// capi.cu — the C ABI declared in include/exomedepth_b200.h: context, device workspaces, and the
// host-/device-pointer entry points that stand in for the reference's two .Call routines
// (src/ExomeDepth_init.c:14-24) and for the per-sample loop that R drives around them.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include <cuda.h>

#include "../../include/exomedepth_b200.h"
#include "host_tables.h"
#include "kernels.cuh"
#include "viterbi_seam.h"

namespace {

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

struct Context {
    bool ready = false;
    int device = 0;
    int n_sms = 148;
    int cc_major = 0, cc_minor = 0;
    size_t smem_optin = 0;
    std::string name;
    cudaStream_t stream = nullptr;       // used by the host-pointer entry points
    cudaStream_t stream2 = nullptr;      // drains results to the host while `stream` computes the next chunk
    cudaEvent_t chunk_done[16] = {};
    // chromosome-group pipeline: copy-in, emission and one Viterbi stream per group, forked from / joined to the caller's stream
    static constexpr int kMaxParts = 6;
    cudaStream_t s_copy = nullptr, s_em = nullptr, s_vit[kMaxParts] = {};
    cudaStream_t s_widen = nullptr, s_cor = nullptr;      // sample-chunk pipeline: widening gates a chunk's emission (highest priority), cor(test, reference) is only needed at the end (lowest)
    cudaEvent_t ev_fork = nullptr, ev_copy[kMaxParts] = {}, ev_em[kMaxParts] = {}, ev_vit[kMaxParts] = {}, ev_setup = nullptr;
    int* d_queue = nullptr;              // work-item counter of the emission lattice kernel
    DevBuf cold_spill;                   // lattice kernel: parking list of out-of-lattice bins beyond the shared-memory list (per CTA)
    unsigned* d_flags = nullptr;         // sticky device warning word
    unsigned* d_gsl_count = nullptr;     // per-cell GSL error log of the `.Call`-shaped emission (kernels.cuh: GslEventLog)
    uint4* d_gsl_events = nullptr;
    unsigned sticky = 0;
    std::vector<DevBuf*> bufs;
};

Context g;
std::mutex g_mu;
thread_local std::string t_err;
std::atomic<long long> g_launches{0};

// per-kernel timing: consecutive marks on a stream bracket the kernel launched between them
struct EventTimer : edb::KernelTimer {
    struct Interval { std::string name; cudaEvent_t e0, e1; };
    std::vector<Interval> done;
    std::vector<cudaEvent_t> pool;
    struct Open { cudaStream_t st; std::string name; cudaEvent_t ev; };
    std::vector<Open> open;           // one open interval per stream (the group pipeline launches on several)
    bool print_timeline = false;
    cudaEvent_t get()
    {
        if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
        cudaEvent_t e;
        cudaEventCreate(&e);
        return e;
    }
    void mark(const char* name, cudaStream_t st) override
    {
        cudaEvent_t e = get();
        cudaEventRecord(e, st);
        size_t k = 0;
        while (k < open.size() && open[k].st != st) k++;
        if (k < open.size()) {
            done.push_back({open[k].name, open[k].ev, e});
            if (name) { open[k].name = name; open[k].ev = e; }
            else open.erase(open.begin() + k);
        } else if (name) {
            open.push_back({st, name, e});
        }
        // an event that only closes an interval is owned by that interval; one that opens the next is shared
    }
    std::string read()
    {
        cudaDeviceSynchronize();
        std::vector<std::string> names;
        std::vector<double> total, longest;
        std::vector<int> count;
        const bool timeline = print_timeline;                            // edb200_profile(2): start / duration of every launch
        for (auto& iv : done) {
            float ms = 0;
            cudaEventElapsedTime(&ms, iv.e0, iv.e1);
            if (timeline) {
                float t0 = 0;
                cudaEventElapsedTime(&t0, done.front().e0, iv.e0);
                fprintf(stderr, "[timeline] %-18s start %8.3f ms  dur %7.3f ms\n", iv.name.c_str(), t0, ms);
            }
            size_t k = 0;
            while (k < names.size() && names[k] != iv.name) k++;
            if (k == names.size()) { names.push_back(iv.name); total.push_back(0); longest.push_back(0); count.push_back(0); }
            total[k] += ms;
            if (ms > longest[k]) longest[k] = ms;
            count[k]++;
        }
        done.clear();               // events are left to the driver: a profile run is short
        std::string out;
        char line[160];
        for (size_t k = 0; k < names.size(); k++) {
            snprintf(line, sizeof line, "%s%s:%d:%.6f:%.6f", k ? ";" : "", names[k].c_str(), count[k], total[k], longest[k]);
            out += line;
        }
        return out;
    }
};
EventTimer g_event_timer;

int fail(int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    t_err = buf;
    return code;
}

#define CU(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess)                                                                       \
            return fail(EDB200_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// bumped whenever a workspace moves: captured graphs (edb200_cohort_capture_device) hold workspace addresses
std::atomic<long long> g_realloc_gen{0};

int ensure(DevBuf& b, size_t bytes)
{
    if (bytes <= b.cap) return 0;
    g_realloc_gen++;
    if (b.p) CU(cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    CU(cudaMalloc(&b.p, want));
    b.cap = want;
    return 0;
}

void release(DevBuf& b)
{
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
}

int need_ctx()
{
    if (g.ready) return 0;
    return edb200_init(-1);
}

int check_kernel(const char* what)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(EDB200_ERR_CUDA, "%s launch failed: %s", what, cudaGetErrorString(e));
    return 0;
}

void release_refset_scratch();
void release_fit_scratch();
void release_hmm_cache();

// scratch used by the single-call (reference-shaped) entry points
struct CallScratch {
    DevBuf phi, expected, total, observed, odds, ll, consts, lt, chains, bp, path, ccalls, cncalls, calls, ncalls, sched_begin, sched_items;
} cs;

// CUtensorMap of an emission matrix for the Viterbi sweep's 2-D TMA loads: rows = (sample, state) pairs of `cols`
// doubles (row pitch `pitch` doubles), box = 16 bins x (32/S)*S rows, 128-byte swizzle, zero fill out of bounds.
// cuTensorMapEncodeTiled is fetched through the runtime (no link-time dependency on libcuda).
int make_ll_map(const double* ll, int64_t rows, int64_t cols, int64_t pitch, int box_rows, CUtensorMap* out, int box_cols = 16)
{
    typedef CUresult (*Encode)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static Encode encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn)
            return fail(EDB200_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
        encode = (Encode)fn;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)pitch * 8};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(ll), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, box_cols == 16 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                              // (L2 promotion to 128 / 256 bytes for the 64-byte rows of the half-tile boxes: measured for the latency-bound and for
                              //  the segmented sweep, no change; an L2 prefetch of the box two stages ahead: 0.67 -> 1.18 ms)
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(EDB200_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d): ll %p rows %lld cols %lld pitch %lld", (int)r, (const void*)ll,
                                       (long long)rows, (long long)cols, (long long)pitch);
    return 0;
}


}  // namespace

edb::KernelTimer* edb::g_timer = nullptr;

// =====================================================================================================
struct edb200_graph {
    cudaGraphExec_t exec = nullptr;      // null once the cohort it was captured from is destroyed
    edb200_cohort* cohort = nullptr;
    long long realloc_gen = 0;           // g_realloc_gen at capture
    int kernels = 0;                     // kernel nodes (edb200_launch_count advances by this per replay)
};

struct edb200_cohort {
    int64_t n_bins = 0;
    int32_t n_chains = 0;
    int32_t S = 3;
    int64_t total_rows = 0;
    double L = 50000.0;
    double odds[EDB200_MAX_STATES];
    double T[EDB200_MAX_STATES * EDB200_MAX_STATES];
    int perm[EDB200_MAX_STATES];
    std::vector<edb::ChainDesc> chains_h;
    int64_t total_tiles = 0;
    int lt_pitch = 0;
    DevBuf chains, lt, odds_d, tile_base, decay, srows;
    // CallCNVs-structured view of the table for the one-thread-per-chain sweep (built lazily from the device table, so
    // that ranks that received the table by broadcast get it too): 0 = not examined, 1 = structured, -1 = arbitrary matrix
    int struct_state = 0;
    double c0 = 0, c1 = 0;
    // edb200_cohort_set_option
    int opt_sweep = 0;                   // 0 auto, 1 one lane per (chain, state), 2 one thread per chain
    int opt_parts = 0;                   // chromosome groups of a pipelined batch (0 auto)
    int opt_vsplit = -1;                 // device-resident Viterbi as two concurrent passes (-1 auto)
    int opt_crit_warps = 0, opt_sweep_warps = 0, opt_packplan = 0;
    int opt_segments = -1, opt_seg_warm = 0, opt_seg_min = 0, opt_seg_repair = 0, opt_reserve = 0, opt_chunks = 0;
    // segmented sweep (viterbi_seam.h); seg_ok: the transition terms are small enough for the error bound (ensure_struct)
    int seg_ok = 0;
    int consts_first = -1;               // >= 0: the per-state constants of the whole host batch are built; a chunk's start at this sample
    int seg_slot = 0;                    // which sample chunk of a host call is being processed (its pieces and flags are its own)
    int vit_slots = 1, vit_slot = 0, vit_slot_samples = 0;     // (slots sized for vit_slot_samples, the largest chunk)
    // sample-chunk pipeline: a set of Viterbi scratch (back-pointers, per-chain call tables) per chunk
    int emission_sms = 0;                // SMs the emission launch of the current chunk may take (0 = all)
    bool in_host_call = false;
    // the pieces of one chromosome group for `key` (samples, warps, warm-up, shortest piece), its scratch and flags
    struct SegPlan {
        long long key = -1;
        int pieces = 0, ctas = 0;
        DevBuf desc, first, begin, items, seam_in, seam_out, seam_mag, close, flags;
    };
    // Chromosome groups ("parts"): the chains are split by length so that the emission of the long chromosomes can
    // finish — and their sweeps, the critical path, can start — while the rest is still being computed (or uploaded).
    struct Part {
        std::vector<int32_t> chains;     // ascending chain ids
        edb::BinRanges ranges;           // 16-bin aligned, merged bin ranges covering the chains
        int max_tiles = 0;
        DevBuf chain_list, sched_begin, sched_items;   // device chain list; sweep schedule for `sched_groups` sample groups
        int sched_groups = 0;
        int sched_warps = 4;             // sweep warps per CTA the schedule was built for
        int sched_ctas = 1;              // sweep CTAs that have work
        int sched_avail = 0;             // SMs the schedule was allowed to use
        SegPlan seg;
    };
    std::vector<Part> plans[Context::kMaxParts + 1];   // plans[n]: the split into n parts (built on first use);
                                                       // plans[0]: {the longest chains, the rest} for the device-resident Viterbi
    // segmented sweeps: all chains as one group, per (samples, chunk slot) — pieces, scratch and flags depend on the sample
    // count, and the sample chunks of one host call keep their flags apart (edb200_cohort_segment_stats adds them up)
    std::map<std::pair<int, int>, Part> seg_parts;
    std::vector<std::pair<Part*, int>> seg_used;       // (group, samples) of the segmented passes since the last top-level call
    // per-batch scratch
    DevBuf consts, bp, ccalls, cncalls, fw_grid, fw_chain, fw_out, fw_best, lattices, cold_spill;
    std::vector<edb200_graph*> graphs;   // captured replays of this cohort (invalidated when the cohort goes)
    int last_host_samples = 0;           // samples whose likelihoods the last host-pointer run left in h_ll
    // host-mode staging
    DevBuf h_obs, h_ref, h_phi, h_exp, h_ll, h_path, h_calls, h_ncalls, h_stats, h_cor, h_obs16, h_ovf_i, h_ovf_v;
};

extern "C" {

int edb200_init(int device)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (g.ready && (device < 0 || device == g.device)) return 0;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(EDB200_ERR_CUDA, "no CUDA device available (%s); exomedepth_b200 has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0) {
        const char* lr = getenv("LOCAL_RANK");
        device = lr ? atoi(lr) % count : 0;
    }
    if (device >= count) return fail(EDB200_ERR_ARG, "device %d out of range (%d devices)", device, count);
    CU(cudaSetDevice(device));
    cudaDeviceProp p;
    CU(cudaGetDeviceProperties(&p, device));
    g.device = device;
    g.n_sms = p.multiProcessorCount;
    g.cc_major = p.major;
    g.cc_minor = p.minor;
    g.smem_optin = p.sharedMemPerBlockOptin;
    g.name = p.name;
    if (p.major < 10)
        return fail(EDB200_ERR_CUDA, "device %d (%s, sm_%d%d) is not a Blackwell sm_100 part; this library ships sm_100a code only",
                    device, p.name, p.major, p.minor);
    if (!g.stream) CU(cudaStreamCreateWithFlags(&g.stream, cudaStreamNonBlocking));
    if (!g.stream2) {
        CU(cudaStreamCreateWithFlags(&g.stream2, cudaStreamNonBlocking));
        for (cudaEvent_t& e : g.chunk_done) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    if (!g.d_flags) {
        CU(cudaMalloc(&g.d_flags, sizeof(unsigned)));
        CU(cudaMemset(g.d_flags, 0, sizeof(unsigned)));
    }
    if (!g.s_em) {
        CU(cudaStreamCreateWithFlags(&g.s_copy, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&g.s_em, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&g.ev_fork, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&g.ev_setup, cudaEventDisableTiming));
        for (int p = 0; p < Context::kMaxParts; p++) {
            CU(cudaStreamCreateWithFlags(&g.s_vit[p], cudaStreamNonBlocking));
            CU(cudaEventCreateWithFlags(&g.ev_copy[p], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&g.ev_em[p], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&g.ev_vit[p], cudaEventDisableTiming));
        }
        int prio_lo = 0, prio_hi = 0;                       // (numerically lower = served first)
        CU(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        CU(cudaStreamCreateWithPriority(&g.s_widen, cudaStreamNonBlocking, prio_hi));
        CU(cudaStreamCreateWithPriority(&g.s_cor, cudaStreamNonBlocking, prio_lo));
        CU(cudaMalloc(&g.d_queue, sizeof(int)));
    }
    g.ready = true;
    return 0;
}

void edb200_shutdown(void)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g.ready) return;
    cudaDeviceSynchronize();
    DevBuf* all[] = {&cs.phi, &cs.expected, &cs.total, &cs.observed, &cs.odds, &cs.ll, &cs.consts, &cs.lt,
                     &cs.chains, &cs.bp, &cs.path, &cs.ccalls, &cs.cncalls, &cs.calls, &cs.ncalls, &cs.sched_begin, &cs.sched_items};
    for (DevBuf* b : all) release(*b);
    release_hmm_cache();
    release_refset_scratch();            // also releases the fit scratch
    if (g.d_flags) cudaFree(g.d_flags);
    g.d_flags = nullptr;
    if (g.d_gsl_count) cudaFree(g.d_gsl_count);
    if (g.d_gsl_events) cudaFree(g.d_gsl_events);
    g.d_gsl_count = nullptr;
    g.d_gsl_events = nullptr;
    if (g.stream) cudaStreamDestroy(g.stream);
    g.stream = nullptr;
    if (g.stream2) {
        cudaStreamDestroy(g.stream2);
        for (cudaEvent_t& e : g.chunk_done) cudaEventDestroy(e);
    }
    g.stream2 = nullptr;
    if (g.s_em) {
        cudaStreamDestroy(g.s_copy);
        cudaStreamDestroy(g.s_em);
        cudaStreamDestroy(g.s_widen);
        cudaStreamDestroy(g.s_cor);
        cudaEventDestroy(g.ev_fork);
        cudaEventDestroy(g.ev_setup);
        for (int p = 0; p < Context::kMaxParts; p++) {
            cudaStreamDestroy(g.s_vit[p]);
            cudaEventDestroy(g.ev_copy[p]);
            cudaEventDestroy(g.ev_em[p]);
            cudaEventDestroy(g.ev_vit[p]);
        }
        cudaFree(g.d_queue);
        release(g.cold_spill);
    }
    g.s_em = nullptr;
    g.ready = false;
}

const char* edb200_last_error(void) { return t_err.c_str(); }

int edb200_device_info(char* buf, int buflen, int* n_sms, int* cc_major, int* cc_minor)
{
    if (int rc = need_ctx()) return rc;
    if (buf && buflen > 0) snprintf(buf, buflen, "%s sm_%d%d %d SMs", g.name.c_str(), g.cc_major, g.cc_minor, g.n_sms);
    if (n_sms) *n_sms = g.n_sms;
    if (cc_major) *cc_major = g.cc_major;
    if (cc_minor) *cc_minor = g.cc_minor;
    return 0;
}

int64_t edb200_launch_count(int reset)
{
    long long n = g_launches.load();
    if (reset) g_launches.store(0);
    return n;
}

void* edb200_host_alloc(size_t bytes)
{
    if (need_ctx()) return nullptr;
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
        fail(EDB200_ERR_CUDA, "cudaHostAlloc(%zu) failed", bytes);
        return nullptr;
    }
    return p;
}

void edb200_host_free(void* p)
{
    if (p) cudaFreeHost(p);
}

// host-side encoders of the ingestion layouts (host_tables.cpp:pack_counts); no device involved
static int64_t pack_checked(int bits, const int32_t* counts, int64_t stride, int32_t n_samples, int64_t n_bins, void* out, int64_t out_stride,
                            int64_t min_stride, int64_t* ovf_index, int32_t* ovf_value, int64_t cap)
{
    if (n_samples < 0 || n_bins < 0 || cap < 0 || (cap > 0 && (!ovf_index || !ovf_value))) return fail(EDB200_ERR_ARG, "pack_counts: bad size or overflow list"), -1;
    if (n_samples == 0 || n_bins == 0) return 0;
    if (!counts || !out || stride < n_bins || out_stride < min_stride) return fail(EDB200_ERR_ARG, "pack_counts: null matrix or stride smaller than a row"), -1;
    const int64_t n = edb::pack_counts(bits, counts, stride, n_samples, n_bins, out, out_stride, ovf_index, ovf_value, cap);
    if (n < 0) fail(EDB200_ERR_ARG, "pack_counts: negative read count");
    return n;
}

int64_t edb200_pack_counts16(const int32_t* counts, int64_t stride, int32_t n_samples, int64_t n_bins, uint16_t* out16, int64_t out_stride,
                             int64_t* overflow_index, int32_t* overflow_value, int64_t overflow_cap)
{
    return pack_checked(16, counts, stride, n_samples, n_bins, out16, out_stride, n_bins, overflow_index, overflow_value, overflow_cap);
}

int64_t edb200_pack_counts12(const int32_t* counts, int64_t stride, int32_t n_samples, int64_t n_bins, uint8_t* out12, int64_t out_stride,
                             int64_t* overflow_index, int32_t* overflow_value, int64_t overflow_cap)
{
    return pack_checked(12, counts, stride, n_samples, n_bins, out12, out_stride, (n_bins + 1) / 2 * 3, overflow_index, overflow_value, overflow_cap);
}

int edb200_profile(int enable)
{
    if (int rc = need_ctx()) return rc;
    cudaDeviceSynchronize();
    g_event_timer.done.clear();
    g_event_timer.open.clear();
    g_event_timer.print_timeline = enable == 2;
    edb::g_timer = enable ? &g_event_timer : nullptr;
    return 0;
}

int edb200_profile_read(char* buf, int buflen)
{
    if (int rc = need_ctx()) return rc;
    if (!buf || buflen < 1) return fail(EDB200_ERR_ARG, "null buffer");
    const std::string s = g_event_timer.read();
    snprintf(buf, buflen, "%s", s.c_str());
    return (int)s.size() >= buflen ? EDB200_ERR_ARG : 0;
}

int edb200_status(int reset)
{
    if (int rc = need_ctx()) return rc;
    unsigned f = 0;
    CU(cudaMemcpy(&f, g.d_flags, sizeof f, cudaMemcpyDeviceToHost));
    g.sticky |= f;
    int out = (int)g.sticky;
    if (reset) {
        g.sticky = 0;
        CU(cudaMemset(g.d_flags, 0, sizeof(unsigned)));
    }
    return out;
}

}  // extern "C"

// =====================================================================================================
namespace {

constexpr int kHostChunks = 8;                     // sample chunks of the host-pointer cohort call (<= 16 events)
// lattice caps: (3072 + 2 * 11776) * 8 B = 208 KB of shared memory.  The split follows the counts: the test count is the
// fraction e (0.08 .. 0.3, more inside a duplication) of the total, so K ~ N / 4 leaves the fewest cells outside — on the
// synthetic cohort 0.01 % (0.08 % with 2048 + 2 * 12288), 0.5 % instead of 1.9 % at counts twice as deep
constexpr int kTableK = 3072, kTableRN = 11776;

// ---- per-cell GSL error log (src/error.c:35-52) -------------------------------------------------------------
constexpr unsigned kGslLogCap = 1u << 16;
struct GslSite { const char* file; int line; const char* reason; };
// bit order = edb200_math.cuh kSite* (the order the reference reaches them inside one gsl_sf_lnbeta call)
const GslSite kGslSites[10] = {{"beta.c", 56, "domain error"},      {"beta.c", 59, "domain error"},      {"VP_gamma.c", 1338, "domain error"},
                               {"VP_log.c", 202, "domain error"},   {"VP_gamma.c", 1239, "domain error"}, {"VP_gamma.c", 1253, "domain error"},
                               {"VP_gamma.c", 1261, "error"},       {"VP_gamma.c", 803, "error"},         {"VP_gamma.c", 1283, "error"},
                               {"beta.c", 44, "domain error"}};
std::vector<uint4> g_gsl_log;            // events of the last `.Call`-shaped emission, in the reference's order (bin, state)
long long g_gsl_raised = 0;              // events raised (the log keeps the first kGslLogCap)

edb::GslEventLog gsl_log_begin(cudaStream_t st)
{
    g_gsl_log.clear();
    g_gsl_raised = 0;
    if (!g.d_gsl_count) {
        if (cudaMalloc(&g.d_gsl_count, 16) != cudaSuccess || cudaMalloc(&g.d_gsl_events, (size_t)kGslLogCap * sizeof(uint4)) != cudaSuccess)
            return edb::GslEventLog{nullptr, nullptr, 0};
    }
    cudaMemsetAsync(g.d_gsl_count, 0, 4, st);
    return edb::GslEventLog{g.d_gsl_count, g.d_gsl_events, kGslLogCap};
}
int gsl_log_end(cudaStream_t st)
{
    if (!g.d_gsl_count) return 0;
    unsigned n = 0;
    CU(cudaMemcpyAsync(&n, g.d_gsl_count, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    g_gsl_raised = n;
    g_gsl_log.resize(std::min(n, kGslLogCap));
    if (!g_gsl_log.empty()) {
        CU(cudaMemcpyAsync(g_gsl_log.data(), g.d_gsl_events, g_gsl_log.size() * sizeof(uint4), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        // the reference walks the bins in order and, per bin, the states (src/CNV_estimate.cpp:71-80)
        std::sort(g_gsl_log.begin(), g_gsl_log.end(), [](const uint4& a, const uint4& b) {
            const unsigned long long ka = ((unsigned long long)(a.y >> 8) << 32 | a.x), kb = ((unsigned long long)(b.y >> 8) << 32 | b.x);
            return ka != kb ? ka < kb : (a.y & 0xFFu) < (b.y & 0xFFu);
        });
    }
    return 0;
}

int pull_flags(cudaStream_t st, int* warn)
{
    unsigned f = 0;
    CU(cudaMemcpyAsync(&f, g.d_flags, sizeof f, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (f) {
        CU(cudaMemsetAsync(g.d_flags, 0, sizeof(unsigned), st));
        g.sticky |= f;
    }
    *warn = (f & edb::kFlagNaN) ? EDB200_WARN_NAN : 0;
    return 0;
}

// emission for device-resident inputs; consts scratch must hold n_samples*S StateConst
int run_emission_scalar(edb::CountsView cv, const double* d_phi, const double* d_expected, const double* d_odds,
                        edb::StateConst* d_consts, int n_samples, int S, int64_t n_bins, edb::LLView out,
                        int mode, cudaStream_t st)
{
    edb::prof_mark("state_setup", st);
    edb::launch_state_setup(n_samples, S, d_phi, d_expected, d_odds, d_consts, st);
    edb::prof_mark(mode == EDB200_EMISSION_DIRECT ? "emission_direct" : "emission", st);
    g_launches++;
    bool table = mode == EDB200_EMISSION_TABLE;
    if (mode == EDB200_EMISSION_AUTO) table = n_bins >= 4 * (int64_t)(kTableK + 2 * kTableRN);
    if (table) {
        edb::TableDims d{kTableK, kTableRN, kTableRN};
        if (edb::emission_table_smem_bytes(d) > g.smem_optin)
            return fail(EDB200_ERR_CUDA, "device offers %zu B of shared memory per CTA; the lattice kernel needs %zu",
                        g.smem_optin, edb::emission_table_smem_bytes(d));
        edb::BinRanges all{};
        all.n = 1;
        all.b0[0] = 0;
        all.b1[0] = n_bins;
        if (int rc = ensure(g.cold_spill, (size_t)edb::emission_table_max_ctas(g.n_sms) * n_bins * 4)) return rc;
        edb::launch_emission_table(cv, d_consts, n_samples, S, all, d, out, g.d_flags, g.d_queue, g.n_sms, nullptr, 0, (int*)g.cold_spill.p, n_bins, st);
    } else {
        edb::launch_emission_direct(cv, d_consts, n_samples, S, n_bins, out, g.d_flags, st);
    }
    edb::prof_mark(nullptr, st);
    g_launches++;
    return check_kernel("emission");
}

}  // namespace

extern "C" {

int edb200_emission(const double* phi, const double* expected, const int32_t* total, const int32_t* observed,
                    int64_t n, int32_t n_states, const double* odds, double* ll_out)
{
    if (int rc = need_ctx()) return rc;
    if (n_states < 2 || n_states > EDB200_MAX_STATES) return fail(EDB200_ERR_NSTATES, "n_states=%d not in [2,%d]", n_states, EDB200_MAX_STATES);
    if (n < 0 || (n > 0 && (!phi || !expected || !total || !observed || !ll_out)) || !odds) return fail(EDB200_ERR_ARG, "null argument");
    if (n == 0) return 0;
    std::lock_guard<std::mutex> lk(g_mu);
    cudaStream_t st = g.stream;
    const int S = n_states;
    if (int rc = ensure(cs.total, n * 4)) return rc;
    if (int rc = ensure(cs.observed, n * 4)) return rc;
    if (int rc = ensure(cs.odds, S * 8)) return rc;
    if (int rc = ensure(cs.ll, (size_t)n * S * 8)) return rc;
    CU(cudaMemcpyAsync(cs.total.p, total, n * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(cs.observed.p, observed, n * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(cs.odds.p, odds, S * 8, cudaMemcpyHostToDevice, st));
    edb::LLView out{(double*)cs.ll.p, 0, n};

    // R hands over full-length vectors even when the fit is a single (phi, expected) pair
    // (R/class_definition.R:119, 168); detect that and hoist the per-state constants.
    bool constant = true;
    for (int64_t i = 1; i < n && constant; i++) constant = phi[i] == phi[0] && expected[i] == expected[0];
    if (constant) {
        if (int rc = ensure(cs.phi, 8)) return rc;
        if (int rc = ensure(cs.expected, 8)) return rc;
        if (int rc = ensure(cs.consts, S * sizeof(edb::StateConst))) return rc;
        CU(cudaMemcpyAsync(cs.phi.p, phi, 8, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(cs.expected.p, expected, 8, cudaMemcpyHostToDevice, st));
        edb::CountsView cv{(const int32_t*)cs.observed.p, n, (const int32_t*)cs.total.p, 0, 1};
        if (int rc = run_emission_scalar(cv, (double*)cs.phi.p, (double*)cs.expected.p, (double*)cs.odds.p,
                                         (edb::StateConst*)cs.consts.p, 1, S, n, out, EDB200_EMISSION_AUTO, st))
            return rc;
    } else {
        if (int rc = ensure(cs.phi, n * 8)) return rc;
        if (int rc = ensure(cs.expected, n * 8)) return rc;
        CU(cudaMemcpyAsync(cs.phi.p, phi, n * 8, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(cs.expected.p, expected, n * 8, cudaMemcpyHostToDevice, st));
        edb::launch_emission_bins((double*)cs.phi.p, (double*)cs.expected.p, (int32_t*)cs.total.p, (int32_t*)cs.observed.p,
                                  n, S, (double*)cs.odds.p, out, g.d_flags, gsl_log_begin(st), st);
        g_launches++;
        if (int rc = check_kernel("emission_bins")) return rc;
        if (int rc = gsl_log_end(st)) return rc;
    }
    CU(cudaMemcpyAsync(ll_out, cs.ll.p, (size_t)n * S * 8, cudaMemcpyDeviceToHost, st));
    int warn = 0;
    if (int rc = pull_flags(st, &warn)) return rc;
    if (constant) {
        g_gsl_log.clear();
        g_gsl_raised = 0;
        if (warn & EDB200_WARN_NAN) {
            // the hoisted kernels report only THAT a call failed; the reference names every failing call (src/error.c:45-48):
            // the per-bin kernel — same arithmetic, same bits — runs once more over the broadcast pair to list them
            std::vector<double> ph((size_t)n, phi[0]), ex((size_t)n, expected[0]);
            if (int rc = ensure(cs.phi, n * 8)) return rc;
            if (int rc = ensure(cs.expected, n * 8)) return rc;
            CU(cudaMemcpyAsync(cs.phi.p, ph.data(), n * 8, cudaMemcpyHostToDevice, st));
            CU(cudaMemcpyAsync(cs.expected.p, ex.data(), n * 8, cudaMemcpyHostToDevice, st));
            edb::launch_emission_bins((double*)cs.phi.p, (double*)cs.expected.p, (int32_t*)cs.total.p, (int32_t*)cs.observed.p,
                                      n, S, (double*)cs.odds.p, out, g.d_flags, gsl_log_begin(st), st);
            g_launches++;
            if (int rc = check_kernel("emission_bins")) return rc;
            if (int rc = gsl_log_end(st)) return rc;
            int again = 0;
            if (int rc = pull_flags(st, &again)) return rc;
        }
    }
    return warn;
}

int edb200_lnbeta(const double* x, const double* y, int64_t n, double* out)
{
    if (int rc = need_ctx()) return rc;
    if (n < 0 || (n > 0 && (!x || !y || !out))) return fail(EDB200_ERR_ARG, "null argument");
    if (n == 0) return 0;
    std::lock_guard<std::mutex> lk(g_mu);
    cudaStream_t st = g.stream;
    if (int rc = ensure(cs.phi, n * 8)) return rc;
    if (int rc = ensure(cs.expected, n * 8)) return rc;
    if (int rc = ensure(cs.ll, n * 8)) return rc;
    CU(cudaMemcpyAsync(cs.phi.p, x, n * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(cs.expected.p, y, n * 8, cudaMemcpyHostToDevice, st));
    edb::launch_lnbeta((const double*)cs.phi.p, (const double*)cs.expected.p, n, (double*)cs.ll.p, g.d_flags, st);
    g_launches++;
    if (int rc = check_kernel("lnbeta")) return rc;
    CU(cudaMemcpyAsync(out, cs.ll.p, n * 8, cudaMemcpyDeviceToHost, st));
    int warn = 0;
    if (int rc = pull_flags(st, &warn)) return rc;
    return warn;
}

int64_t edb200_gsl_error_log(char* buf, size_t buflen, int64_t first_event, int64_t* next_event)
{
    std::lock_guard<std::mutex> lk(g_mu);
    size_t used = 0;
    int64_t ev = first_event < 0 ? 0 : first_event;
    if (buf && buflen) buf[0] = 0;
    auto put = [&](const char* file, int line, const char* reason) {
        // src/error.c:45-48
        const int w = snprintf(buf + used, buflen - used, "ERROR %s %i %s\nDefault GSL error handler invoked.\n", file, line, reason);
        if (w < 0 || (size_t)w >= buflen - used) return false;
        used += (size_t)w;
        return true;
    };
    for (; buf && ev < (int64_t)g_gsl_log.size(); ev++) {
        const size_t mark = used;
        bool fits = true;
        for (int call = 0; call < 2 && fits; call++) {
            const unsigned sites = call == 0 ? g_gsl_log[ev].z : g_gsl_log[ev].w;
            if (!sites) continue;
            for (int b = 0; b < 10 && fits; b++)
                if (sites >> b & 1u) fits = put(kGslSites[b].file, kGslSites[b].line, kGslSites[b].reason);
            // the value wrapper reports the failed call once more (src/beta.c:163, eval.h:3-9)
            if (fits) fits = put("beta.c", 163, "gsl_sf_lnbeta_e(x, y, &result)");
        }
        if (!fits) {
            used = mark;
            buf[used] = 0;
            break;
        }
    }
    if (next_event) *next_event = ev;
    return g_gsl_raised;
}

int edb200_get_loglike_matrix(const double* phi, const double* expected, const int32_t* total,
                              const int32_t* observed, double mixture, int64_t n, double* ll_out)
{
    const double odds[3] = {1 - 0.5 * mixture, 1.0, 1 + 0.5 * mixture};   // src/CNV_estimate.cpp:65-66
    return edb200_emission(phi, expected, total, observed, n, 3, odds, ll_out);
}


// ------------------------------------------------------------------------------------------------ cohort
int edb200_cohort_create(const edb200_cohort_spec* sp, edb200_cohort** out)
{
    if (int rc = need_ctx()) return rc;
    if (!sp || !out || sp->n_bins < 1 || sp->n_chains < 1 || !sp->chain_offsets || !sp->start || !sp->end)
        return fail(EDB200_ERR_ARG, "bad cohort spec");
    const int S = sp->n_states;
    if (S != 3 && S != 5 && S != 7 && !(sp->odds && sp->transitions && S >= 2 && S <= EDB200_MAX_STATES))
        return fail(EDB200_ERR_NSTATES, "n_states=%d: 3, 5, 7 have built-in tables; other values in [2,7] need explicit odds and transitions", S);
    if (sp->chain_offsets[0] != 0 || sp->chain_offsets[sp->n_chains] != sp->n_bins)
        return fail(EDB200_ERR_ARG, "chain_offsets must start at 0 and end at n_bins");
    std::lock_guard<std::mutex> lk(g_mu);
    edb200_cohort* c = new edb200_cohort();
    c->n_bins = sp->n_bins;
    c->n_chains = sp->n_chains;
    c->S = S;
    c->L = sp->expected_cnv_length;
    // likelihood-column order is by copy number; the normal column is 1 for S=3 (CN 1,2,3) and 2 otherwise (CN 0..)
    const int normal = S == 3 ? 1 : 2;
    if (sp->odds) memcpy(c->odds, sp->odds, S * 8);
    else {
        for (int s = 0; s < S; s++) c->odds[s] = 1 + ((S == 3 ? s + 1 : s) - 2) / 2.0 * sp->mixture;
        c->odds[normal] = 1.0;
        if (S != 3 && c->odds[0] < 0.05) c->odds[0] = 0.05;
    }
    if (sp->transitions) memcpy(c->T, sp->transitions, S * S * 8);
    else edb::callcnvs_transitions(S, sp->transition_probability, c->T);
    c->perm[0] = normal;                                       // R/class_definition.R:364  c(2, 1, 3)
    for (int s = 0, j = 1; s < S; s++) if (s != normal) c->perm[j++] = s;

    // frame every chromosome and build its log-transition rows
    c->chains_h.resize(c->n_chains);
    int64_t rows = 0;
    for (int ch = 0; ch < c->n_chains; ch++) {
        const int64_t b0 = sp->chain_offsets[ch], b1 = sp->chain_offsets[ch + 1];
        if (b1 <= b0) { delete c; return fail(EDB200_ERR_ARG, "empty chromosome %d", ch); }
        edb::ChainDesc& cd = c->chains_h[ch];
        cd.lt_row0 = rows;
        cd.nobs = (int32_t)(b1 - b0 + 2);
        cd.n_em = (int32_t)(b1 - b0);
        cd.em_off = b0 - 1;
        cd.out_off = b0 - 1;
        cd.out_first = 1;
        cd.out_last = (int32_t)(b1 - b0);
        cd.call_shift = (int32_t)(b0 - 1);
        cd.pad = 0;
        rows += cd.nobs;
    }
    c->total_rows = rows;
    const int pitch = edb::viterbi_lt_pitch(S);
    c->lt_pitch = pitch;
    std::vector<double> lt(((size_t)rows + edb::viterbi_tile()) * pitch, 0.0);
    std::vector<int32_t> tile_base(c->n_chains);
    for (int ch = 0; ch < c->n_chains; ch++) {
        tile_base[ch] = (int32_t)c->total_tiles;
        c->total_tiles += edb::viterbi_chain_tiles(c->chains_h[ch]);
    }
    std::vector<int32_t> pos;
    std::vector<double> decay((size_t)rows + edb::viterbi_tile(), 0.0);
    for (int ch = 0; ch < c->n_chains; ch++) {
        const int64_t b0 = sp->chain_offsets[ch], nb = sp->chain_offsets[ch + 1] - b0;
        pos.resize(nb + 2);
        if (edb::frame_positions(nb, sp->start + b0, sp->end + b0, c->L, pos.data())) {
            delete c;
            return fail(EDB200_ERR_ARG, "framed position of chromosome %d does not fit an R integer", ch);
        }
        edb::build_decay_rows(pos.data(), (int32_t)(nb + 2), c->L, decay.data() + c->chains_h[ch].lt_row0);
        if (!sp->skip_table_build)
            edb::build_log_transition_rows(S, c->T, pos.data(), (int32_t)(nb + 2), c->L,
                                           lt.data() + (size_t)c->chains_h[ch].lt_row0 * pitch, pitch);
    }
    int rc = 0;
    if ((rc = ensure(c->lt, lt.size() * 8)) || (rc = ensure(c->chains, c->n_chains * sizeof(edb::ChainDesc))) ||
        (rc = ensure(c->odds_d, S * 8)) || (rc = ensure(c->tile_base, c->n_chains * 4))) {
        delete c;
        return rc;
    }
    CU(cudaMemcpy(c->tile_base.p, tile_base.data(), c->n_chains * 4, cudaMemcpyHostToDevice));
    if ((rc = ensure(c->decay, decay.size() * 8))) {
        delete c;
        return rc;
    }
    CU(cudaMemcpy(c->decay.p, decay.data(), decay.size() * 8, cudaMemcpyHostToDevice));
    edb::nan_to_neg_inf(lt.data(), lt.size());     // device copy only: a NaN term is "never selected", like -Inf (hmm.cpp:81)
    CU(cudaMemcpy(c->lt.p, lt.data(), lt.size() * 8, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(c->chains.p, c->chains_h.data(), c->n_chains * sizeof(edb::ChainDesc), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(c->odds_d.p, c->odds, S * 8, cudaMemcpyHostToDevice));
    *out = c;
    return 0;
}

static void destroy_cohort_locked(edb200_cohort* c);
void edb200_cohort_destroy(edb200_cohort* c)
{
    if (!c) return;
    std::lock_guard<std::mutex> lk(g_mu);
    destroy_cohort_locked(c);
}
static void destroy_cohort_locked(edb200_cohort* c)
{
    cudaDeviceSynchronize();
    for (edb200_graph* gr : c->graphs) {
        if (gr->exec) cudaGraphExecDestroy(gr->exec);
        gr->exec = nullptr;
        gr->cohort = nullptr;
    }
    for (auto& plan : c->plans)
        for (auto& part : plan) {
            release(part.chain_list);
            release(part.sched_begin);
            release(part.sched_items);
            for (DevBuf* b : {&part.seg.desc, &part.seg.first, &part.seg.begin, &part.seg.items, &part.seg.seam_in, &part.seg.seam_out,
                              &part.seg.seam_mag, &part.seg.close, &part.seg.flags})
                release(*b);
        }
    for (auto& kv : c->seg_parts) {
        edb200_cohort::Part& part = kv.second;
        for (DevBuf* b : {&part.chain_list, &part.sched_begin, &part.sched_items, &part.seg.desc, &part.seg.first, &part.seg.begin, &part.seg.items,
                          &part.seg.seam_in, &part.seg.seam_out, &part.seg.seam_mag, &part.seg.close, &part.seg.flags})
            release(*b);
    }
    DevBuf* all[] = {&c->chains, &c->lt, &c->odds_d, &c->tile_base, &c->decay, &c->srows, &c->fw_grid, &c->fw_chain, &c->fw_out, &c->fw_best, &c->consts, &c->bp, &c->ccalls, &c->cncalls, &c->lattices, &c->cold_spill,
                     &c->h_obs, &c->h_ref, &c->h_phi, &c->h_exp, &c->h_ll, &c->h_path, &c->h_calls, &c->h_ncalls, &c->h_stats, &c->h_cor,
                     &c->h_obs16, &c->h_ovf_i, &c->h_ovf_v};
    for (DevBuf* b : all) release(*b);
    delete c;
}

int edb200_cohort_table(edb200_cohort* c, void** device_ptr, size_t* bytes)
{
    if (!c) return fail(EDB200_ERR_ARG, "null cohort");
    if (device_ptr) *device_ptr = c->lt.p;
    if (bytes) *bytes = ((size_t)c->total_rows + edb::viterbi_tile()) * c->lt_pitch * 8;
    return 0;
}

int edb200_cohort_table_copy(edb200_cohort* c, void* device_buf, int direction, void* cuda_stream)
{
    if (!c || !device_buf) return fail(EDB200_ERR_ARG, "null argument");
    const size_t bytes = ((size_t)c->total_rows + edb::viterbi_tile()) * c->lt_pitch * 8;
    if (direction == 0) CU(cudaMemcpyAsync(device_buf, c->lt.p, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)cuda_stream));
    else {
        CU(cudaMemcpyAsync(c->lt.p, device_buf, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)cuda_stream));
        c->struct_state = 0;                        // the structured view follows the table
    }
    return 0;
}

int edb200_cohort_set_option(edb200_cohort* c, int option, int value)
{
    if (!c) return fail(EDB200_ERR_ARG, "null cohort");
    std::lock_guard<std::mutex> lk(g_mu);
    switch (option) {
        case EDB200_OPT_SWEEP:
            if (value < 0 || value > 2) return fail(EDB200_ERR_ARG, "EDB200_OPT_SWEEP: 0 auto, 1 lane per state, 2 thread per chain");
            c->opt_sweep = value;
            break;
        case EDB200_OPT_PARTS:
            if (value < 0 || value > Context::kMaxParts) return fail(EDB200_ERR_ARG, "EDB200_OPT_PARTS: 0 (auto) .. %d", Context::kMaxParts);
            c->opt_parts = value;
            break;
        case EDB200_OPT_VSPLIT: c->opt_vsplit = value < 0 ? -1 : value != 0; break;
        case EDB200_OPT_CRIT_WARPS: c->opt_crit_warps = value; break;
        case EDB200_OPT_SWEEP_WARPS: c->opt_sweep_warps = value; break;
        case EDB200_OPT_PACKPLAN: c->opt_packplan = value; break;
        case EDB200_OPT_SEGMENTS: c->opt_segments = value < 0 ? -1 : value != 0; break;
        case EDB200_OPT_SEG_WARM:
            if (value < 0 || value > 64) return fail(EDB200_ERR_ARG, "EDB200_OPT_SEG_WARM: 0 (default) .. 64 tiles");
            c->opt_seg_warm = value;
            break;
        case EDB200_OPT_SEG_MIN:
            if (value < 0) return fail(EDB200_ERR_ARG, "EDB200_OPT_SEG_MIN: tiles >= 0");
            c->opt_seg_min = value;
            break;
        case EDB200_OPT_SEG_REPAIR: c->opt_seg_repair = value != 0; break;
        case EDB200_OPT_RESERVE: c->opt_reserve = value < 0 ? 0 : value; break;
        case EDB200_OPT_CHUNKS: c->opt_chunks = value < 0 ? 0 : value; break;
        default: return fail(EDB200_ERR_ARG, "unknown option %d", option);
    }
    for (auto& plan : c->plans)
        for (auto& part : plan) {
            part.sched_groups = 0;                          // schedules are rebuilt under the new options
            part.seg.key = -1;
        }
    for (auto& kv : c->seg_parts) {
        kv.second.sched_groups = 0;
        kv.second.seg.key = -1;
    }
    return 0;
}

int edb200_cohort_segment_stats(edb200_cohort* c, int32_t out[10])
{
    if (!c || !out) return fail(EDB200_ERR_ARG, "null argument");
    std::lock_guard<std::mutex> lk(g_mu);
    for (int i = 0; i < 10; i++) out[i] = 0;
    if (c->seg_used.empty()) return 0;
    CU(cudaDeviceSynchronize());
    for (auto& used : c->seg_used) {                           // one set of flags per segmented pass
        edb200_cohort::Part& part = *used.first;
        const int ns = used.second;
        if (!part.seg.flags.p || part.seg.key < 0) continue;
        std::vector<int32_t> f(edb::seg_flag_ints(c->n_chains, ns));
        CU(cudaMemcpy(f.data(), part.seg.flags.p, f.size() * 4, cudaMemcpyDeviceToHost));
        out[0] += part.seg.pieces;
        out[1] += f[0];
        for (int i = 0; i < c->n_chains; i++) out[2] += f[edb::seg_off_chain(i)] != 0;
        for (size_t i = edb::seg_off_pair(c->n_chains, ns, 0, 0); i < edb::seg_off_pair(c->n_chains, ns, c->n_chains, 0); i++) {
            out[3] += f[i] != 0;
            for (int b = 0; b < 6; b++) out[4 + b] += (f[i] >> b) & 1;
        }
    }
    return 0;
}

// The structured view of the log-transition table (host_tables.h: build_struct_rows), from the DEVICE copy so that a
// rank that received the table by broadcast sees the same rows.  Once per cohort (and per table copy).
static int ensure_struct(edb200_cohort* c)
{
    if (c->struct_state != 0) return 0;
    const int S = c->S;
    c->struct_state = -1;
    if (S != 3 && S != 5 && S != 7) return 0;
    const size_t n_rows = (size_t)c->total_rows + edb::viterbi_tile();
    std::vector<double> lt(n_rows * c->lt_pitch);
    CU(cudaMemcpy(lt.data(), c->lt.p, lt.size() * 8, cudaMemcpyDeviceToHost));
    std::vector<edb::StructRow> rows(n_rows);
    if (!edb::build_struct_rows(S, lt.data(), c->lt_pitch, (int64_t)n_rows, rows.data(), &c->c0, &c->c1)) return 0;
    // the error bound of the segmented sweep (viterbi_seam.h) takes |log t| <= 1024 for every finite term
    c->seg_ok = std::isfinite(c->c0) && std::isfinite(c->c1) && std::fabs(c->c0) <= 1024.0 && std::fabs(c->c1) <= 1024.0;
    for (const edb::StructRow& r : rows)
        for (double v : {r.b0, r.sf, r.ot})
            if (v != v || v == HUGE_VAL || (v > -HUGE_VAL && std::fabs(v) > 1024.0)) c->seg_ok = 0;
    if (int rc = ensure(c->srows, n_rows * sizeof(edb::StructRow))) return rc;
    CU(cudaMemcpy(c->srows.p, rows.data(), n_rows * sizeof(edb::StructRow), cudaMemcpyHostToDevice));
    c->struct_state = 1;
    return 0;
}

// ---- chromosome-group plans --------------------------------------------------------------------------
// Split the chains into n_parts groups by descending length (cumulative bin fractions below) and describe each
// group by its chain list and by the 16-bin aligned ranges of the likelihood matrix its chains read.
static int build_plan(edb200_cohort* c, int n_parts)
{
    static const double kCuts[Context::kMaxParts + 1][Context::kMaxParts] = {
        {}, {1.0}, {0.45, 1.0}, {0.35, 0.7, 1.0}, {0.3, 0.6, 0.85, 1.0}, {0.25, 0.5, 0.75, 0.9, 1.0}, {0.09, 0.28, 0.5, 0.7, 0.88, 1.0}};
    // (6 parts: the longest chromosome on its own — its sweep is the longest dependent chain of the batch and can start as soon
    // as a tenth of the counts has arrived)
    std::vector<edb200_cohort::Part>& plan = c->plans[n_parts];
    if (!plan.empty()) return 0;
    std::vector<int> order(c->n_chains);
    for (int i = 0; i < c->n_chains; i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return c->chains_h[x].nobs > c->chains_h[y].nobs; });
    if (n_parts == 0) {
        // the chains within 25 % of the longest one (their sweep is the critical path) | all others
        plan.resize(2);
        const double lim = 0.75 * c->chains_h[order[0]].nobs;
        for (int oc = 0; oc < c->n_chains; oc++) plan[c->chains_h[order[oc]].nobs > lim ? 0 : 1].chains.push_back(order[oc]);
    } else {
        plan.resize(n_parts);
        double cuts[Context::kMaxParts];
        for (int p = 0; p < n_parts; p++) cuts[p] = kCuts[n_parts][p];
        int64_t cum = 0;
        int part = 0;
        for (int oc = 0; oc < c->n_chains; oc++) {
            const int ch = order[oc];
            while (part + 1 < n_parts && !plan[part].chains.empty() && (double)cum >= cuts[part] * (double)c->n_bins) part++;
            plan[part].chains.push_back(ch);
            cum += c->chains_h[ch].n_em;
        }
    }
    while (!plan.empty() && plan.back().chains.empty()) plan.pop_back();
    for (auto& pt : plan) {
        std::sort(pt.chains.begin(), pt.chains.end());
        // EXACT bin ranges of the part's chains (adjacent chains merged).  Not rounded out to whole 16-bin tiles: the bins of
        // a neighbouring chromosome belong to another part, whose sweep may be running — the lattice kernel stores a
        // clamped-gather value for out-of-lattice cells before its cold pass writes the right one, and a rewrite of
        // somebody else's bins exposes that transient to a concurrent reader (seen as one extra call in one sample every
        // few calls with pinned buffers).  The kernels take unaligned ranges: scalar head and tail, vector body.
        std::vector<std::pair<int64_t, int64_t>> rs;
        for (int ch : pt.chains) {
            const edb::ChainDesc& cd = c->chains_h[ch];
            const int64_t b0 = cd.em_off + 1;
            const int64_t b1 = std::min<int64_t>(c->n_bins, cd.em_off + 1 + cd.n_em);
            if (!rs.empty() && b0 <= rs.back().second) rs.back().second = std::max(rs.back().second, b1);
            else rs.push_back({b0, b1});
            pt.max_tiles = std::max(pt.max_tiles, edb::viterbi_chain_tiles(cd));
        }
        while ((int)rs.size() > edb::kMaxBinRanges) {          // (more than 32 disjoint runs in one part: close the smallest gaps;
                                                               //  cannot happen with up to 32 chromosomes per part)
            size_t best = 1;
            for (size_t i = 2; i < rs.size(); i++)
                if (rs[i].first - rs[i - 1].second < rs[best].first - rs[best - 1].second) best = i;
            rs[best - 1].second = rs[best].second;
            rs.erase(rs.begin() + best);
        }
        pt.ranges.n = (int)rs.size();
        for (size_t i = 0; i < rs.size(); i++) { pt.ranges.b0[i] = rs[i].first; pt.ranges.b1[i] = rs[i].second; }
        if (int rc = ensure(pt.chain_list, pt.chains.size() * 4)) return rc;
        CU(cudaMemcpy(pt.chain_list.p, pt.chains.data(), pt.chains.size() * 4, cudaMemcpyHostToDevice));
    }
    return 0;
}

static bool use_table(const edb200_cohort* c, int mode)
{
    if (mode == EDB200_EMISSION_TABLE) return true;
    return mode == EDB200_EMISSION_AUTO && c->n_bins >= 4 * (int64_t)(kTableK + 2 * kTableRN);
}

// Panels (a few thousand bins, many samples): the full lattice does not amortise its 26,624-entry build over the bins of
// one (sample, state), the in-register kernel pays ~450 FP64 instructions per cell — a lattice sized to the panel, built
// by all threads over one shared index space (emission.cu, kPanel), costs ~65 per entry and a gather per cell.  Worth it
// once there are enough (sample, state) items to fill the SMs; counts beyond it take the in-register path cell by cell.
static bool panel_table(const edb200_cohort* c, int64_t n_items, int mode, edb::TableDims* d)
{
    if (mode == EDB200_EMISSION_DIRECT || mode == EDB200_EMISSION_TABLE) return false;
    if (mode == EDB200_EMISSION_AUTO && (c->n_bins < 4096 || n_items < 2 * (int64_t)g.n_sms)) return false;
    *d = c->n_bins < 16384 ? edb::TableDims{2048, 4096, 4096} : edb::TableDims{2048, 8192, 8192};
    return true;
}

// how many chromosome groups a batch is pipelined over (1 = emission, then Viterbi, on the caller's stream)
static int pick_parts(const edb200_cohort* c, int mode, int wanted)
{
    if (c->opt_parts > 0) wanted = c->opt_parts;
    if (!use_table(c, mode) || c->n_chains < 2 * wanted) return 1;      // small panels: launch-bound, nothing to overlap
    return wanted < 1 ? 1 : wanted > Context::kMaxParts ? Context::kMaxParts : wanted;
}

// lattice_mode: 0 = one launch covers the batch; 1 = first part of a pipelined batch (keeps the lattices in HBM);
// 2 = later part (reloads them)
static int emission_part(edb200_cohort* c, const edb200_batch* b, const edb::BinRanges& rg, bool whole, int mode, int lattice_mode,
                         cudaStream_t st, int n_sms = 0)
{
    if (n_sms <= 0 || n_sms > g.n_sms) n_sms = g.n_sms;
    const int S = c->S, ns = b->n_samples;
    if (b->per_bin_stride) {                                 // per-bin phi / expected: constants per bin, in registers
        if (!whole) return fail(EDB200_ERR_ARG, "internal: the per-bin emission kernel covers whole rows only");
        edb::prof_mark("emission_bins", st);
        edb::launch_emission_bins_batch(edb::CountsView{b->observed, b->obs_stride, b->reference, b->ref_stride, 0}, b->phi, b->expected,
                                        b->per_bin_stride, ns, S, c->n_bins, (const double*)c->odds_d.p,
                                        edb::LLView{b->ll, (int64_t)S * b->ll_stride, b->ll_stride}, g.d_flags, st);
        edb::prof_mark(nullptr, st);
        g_launches++;
        return check_kernel("emission_bins");
    }
    edb::CountsView cv{b->observed, b->obs_stride, b->reference, b->ref_stride, 0};
    edb::LLView out{b->ll, (int64_t)S * b->ll_stride, b->ll_stride};
    edb::StateConst* consts = (edb::StateConst*)c->consts.p + (c->consts_first > 0 ? (size_t)c->consts_first * S : 0);
    edb::TableDims d{kTableK, kTableRN, kTableRN};
    const bool panel = !use_table(c, mode) && whole && lattice_mode == 0 && panel_table(c, (int64_t)ns * S, mode, &d);
    if (panel || use_table(c, mode)) {
        if (edb::emission_table_smem_bytes(d) > g.smem_optin)
            return fail(EDB200_ERR_CUDA, "device offers %zu B of shared memory per CTA; the lattice kernel needs %zu",
                        g.smem_optin, edb::emission_table_smem_bytes(d));
        if (lattice_mode)
            if (int rc = ensure(c->lattices, (size_t)ns * S * (kTableK + 2 * kTableRN) * 8)) return rc;
        edb::prof_mark(panel ? "emission_panel" : "emission", st);
        int64_t span = 0;
        for (int q = 0; q < rg.n; q++) span += rg.b1[q] - rg.b0[q];
        if (int rc = ensure(c->cold_spill, (size_t)edb::emission_table_max_ctas(g.n_sms) * span * 4)) return rc;
        edb::launch_emission_table(cv, consts, ns, S, rg, d, out, g.d_flags, g.d_queue, n_sms, (double*)c->lattices.p, lattice_mode, (int*)c->cold_spill.p, span, st);
    } else {
        if (!whole) return fail(EDB200_ERR_ARG, "internal: the in-register emission kernel covers whole rows only");
        edb::prof_mark("emission_direct", st);
        edb::launch_emission_direct(cv, consts, ns, S, c->n_bins, out, g.d_flags, st);
    }
    edb::prof_mark(nullptr, st);
    g_launches++;
    return check_kernel("emission");
}

static int state_setup(edb200_cohort* c, const edb200_batch* b, cudaStream_t st)
{
    if (b->per_bin_stride) return 0;                        // nothing to hoist: the constants change with the bin
    if (int rc = ensure(c->consts, (size_t)b->n_samples * c->S * sizeof(edb::StateConst))) return rc;
    edb::prof_mark("state_setup", st);
    edb::launch_state_setup(b->n_samples, c->S, b->phi, b->expected, (const double*)c->odds_d.p, (edb::StateConst*)c->consts.p, st);
    edb::prof_mark(nullptr, st);
    g_launches++;
    return check_kernel("state_setup");
}

// scratch + arguments shared by every part of a Viterbi pass over batch b
// Which sweep kernel: both are exact; they differ in what bounds them (DESIGN.md "Viterbi").  One lane per (chain, state)
// steps in ~150 cycles but a warp carries only 32/S chains; one thread per chain steps in ~205 cycles and a warp carries 32.
// The pass ends when the longest chain does, or when the warps have turned over all chain-steps: whichever kernel's larger
// bound is smaller wins (256 samples x 200k bins: lane per state, 1.5 vs 2.0 ms; 2,000 samples: thread per chain, 14.5 vs 2.6 ms).
static bool use_tpc(const edb200_cohort* c, int n_samples)
{
    if (c->struct_state != 1 || c->opt_sweep == 1) return false;
    if (c->opt_sweep == 2) return true;
    const int S = c->S, G = 32 / S;
    int64_t longest = 0, total = 0;
    for (const auto& cd : c->chains_h) {
        longest = std::max<int64_t>(longest, cd.nobs);
        total += cd.nobs;
    }
    const double slots = 4.0 * g.n_sms;                  // one sweep warp per SM sub-partition
    const double lane = std::max(150.0 * longest, 165.0 * total * ((n_samples + G - 1) / G) / slots);
    const double tpc = std::max(205.0 * longest, 205.0 * total * ((n_samples + 31) / 32) / slots);
    return tpc < lane;
}

// ll_map: two tensor maps (lane-per-state box, thread-per-chain box)
static int viterbi_prepare(edb200_cohort* c, const edb200_batch* b, edb::ViterbiArgs& a, CUtensorMap* ll_map)
{
    if (int rc = ensure_struct(c)) return rc;
    if (c->opt_sweep == 2 && c->struct_state != 1)
        return fail(EDB200_ERR_ARG, "EDB200_OPT_SWEEP = 2 needs 3, 5 or 7 states and the CallCNVs transition structure");
    const int S = c->S, ns = b->n_samples;
    if (!b->path || !b->calls || !b->ncalls || b->call_cap < 1 || b->path_stride < c->n_bins)
        return fail(EDB200_ERR_ARG, "Viterbi outputs missing in batch");
    if ((b->ll_stride & 15) || (reinterpret_cast<uintptr_t>(b->ll) & 127))
        return fail(EDB200_ERR_ARG, "Viterbi needs ll_stride to be a multiple of 16 and ll 128-byte aligned (whole 128-byte lines per row tile)");
    const int ccap = b->call_cap;
    const int G = 32 / S;
    const int64_t groups = (ns + G - 1) / G;
    // (slot sizes rounded up to 256 bytes; a later, smaller chunk of the same call fits the slots of the first)
    const int slot_ns = std::max(ns, c->vit_slots > 1 ? c->vit_slot_samples : 0);
    const size_t bp_bytes = ((size_t)((slot_ns + G - 1) / G) * (c->total_tiles + 1) * edb::viterbi_record_bytes() + 255) & ~(size_t)255;
    const size_t cc_bytes = ((size_t)slot_ns * c->n_chains * ccap * 16 + 255) & ~(size_t)255, cn_bytes = ((size_t)slot_ns * c->n_chains * 4 + 255) & ~(size_t)255;
    if (int rc = ensure(c->bp, c->vit_slots * bp_bytes)) return rc;
    if (int rc = ensure(c->ccalls, c->vit_slots * cc_bytes)) return rc;
    if (int rc = ensure(c->cncalls, c->vit_slots * cn_bytes)) return rc;
    a = edb::ViterbiArgs{};
    a.chains = (const edb::ChainDesc*)c->chains.p;
    a.n_chains = c->n_chains;
    a.n_samples = ns;
    a.n_states = S;
    a.ll = b->ll;
    a.ll_sample_stride = (int64_t)S * b->ll_stride;
    a.ll_state_stride = b->ll_stride;
    if (int rc = make_ll_map(b->ll, (int64_t)ns * S, b->ll_stride, b->ll_stride, (32 / S) * S, ll_map)) return rc;
    a.ll_map = ll_map;
    a.tpc = use_tpc(c, ns) ? 1 : 0;
    if (c->struct_state == 1 && c->opt_sweep != 1) {
        if (int rc = make_ll_map(b->ll, (int64_t)ns * S, b->ll_stride, b->ll_stride, 32 * S, ll_map + 1, 8)) return rc;
        a.ll_map_tpc = ll_map + 1;
        a.srows = (const edb::StructRow*)c->srows.p;
        a.c0 = c->c0;
        a.c1 = c->c1;
    }
    for (int j = 0; j < S; j++) a.perm[j] = c->perm[j];
    a.groups = (int)groups;
    a.lt = (const double*)c->lt.p;
    a.bp = (uint32_t*)((char*)c->bp.p + c->vit_slot * bp_bytes);
    a.bp_tile_base = (const int32_t*)c->tile_base.p;
    a.tail_other = -100.0;                                 // R/class_definition.R:364
    a.path = b->path;
    a.path_stride = b->path_stride;
    a.chain_calls = (int32_t*)((char*)c->ccalls.p + c->vit_slot * cc_bytes);
    a.chain_ncalls = (int32_t*)((char*)c->cncalls.p + c->vit_slot * cn_bytes);
    a.chain_call_cap = ccap;
    a.calls = b->calls;
    a.ncalls = b->ncalls;
    a.call_cap = b->call_cap;
    a.flags = g.d_flags;
    return 0;
}

// Lane-per-state sweep: warps per CTA of the pass that carries the longest chains (`packed` = 1 below).  Two warps per
// CTA: four warps keep an SM's shared-memory pipe busy for 118 of a step's 160 cycles; measured 1.617 (4 warps) /
// 1.514 (2) / 1.526 ms (1) for the critical sweep at 256 x 200k x 5 (profiles/r2a_knob_ab.log).
static int crit_warps(const edb200_cohort* c)
{
    const int v = c->opt_crit_warps;
    return v == 1 || v == 2 || v == 4 ? v : 2;
}

// sweep, tilemap, trace, expand of one part on `st`; the schedule is rebuilt when the number of sample groups changes
// `packed` > 0: the sweep shares the GPU with the emission kernel of the next part (a sweep CTA owns its SM's shared
// memory, so does an emission CTA): its CTAs are filled instead of spread over all SMs — 1: one warp per SM
// sub-partition, which keeps the long chains of the first part at full speed; 2: two per sub-partition (shorter chains,
// half the SMs).
static int viterbi_part(edb200_cohort* c, edb200_cohort::Part& pt, edb::ViterbiArgs a, int packed, cudaStream_t st, int avail_sms = 0,
                        int tpc_warps = 0)
{
    if (avail_sms <= 0 || avail_sms > g.n_sms) avail_sms = g.n_sms;
    // tpc_warps > 0: a pipelined part — the one-thread-per-chain sweep (when the table allows it) with that many warps per
    // CTA: it needs 32/S times fewer warps, hence SMs, than the lane-per-state sweep, and the SMs are what the emission of
    // the next chromosome group is waiting for
    if (tpc_warps > 0 && c->struct_state == 1 && c->opt_sweep != 1 && a.srows) a.tpc = 1;
    // thread-per-chain sweep: work items are (chain, 32 samples); lane-per-state sweep: (chain, 32/S samples)
    const int sched_groups = a.tpc ? (a.n_samples + 31) / 32 : a.groups;
    const int key = (sched_groups * 2 + a.tpc) * 8 + tpc_warps;
    if (pt.sched_groups != key || pt.sched_avail != avail_sms) {
        std::vector<int32_t> nobs(pt.chains.size());
        for (size_t i = 0; i < pt.chains.size(); i++) nobs[i] = c->chains_h[pt.chains[i]].nobs;
        const int64_t n_items = (int64_t)nobs.size() * sched_groups;
        int sched_ctas;
        if (a.tpc) {
            // one sweep warp per SM sub-partition; as few warps per CTA as place the items on the available SMs, because
            // the CTA's shared memory is split between its warps' rings (the fewer, the deeper)
            const int max_w = edb::viterbi_tpc_max_warps(a.n_states);
            int w = c->opt_sweep_warps > 0 ? c->opt_sweep_warps : tpc_warps > 0 ? tpc_warps : (int)((n_items + avail_sms - 1) / avail_sms);
            pt.sched_warps = w < 1 ? 1 : w > max_w ? max_w : w;
            sched_ctas = (int)std::min<int64_t>(avail_sms, (n_items + pt.sched_warps - 1) / pt.sched_warps);
        } else {
            const int force = c->opt_sweep_warps;
            pt.sched_warps = force ? (force == 8 ? 8 : 4) : packed == 1 ? crit_warps(c) : packed == 2 ? 8 : edb::viterbi_pick_warps(nobs.data(), (int)nobs.size(), a.groups, avail_sms);
            sched_ctas = packed ? (int)std::min<int64_t>(avail_sms, (n_items + pt.sched_warps - 1) / pt.sched_warps) : avail_sms;
        }
        std::vector<int32_t> begin, items;
        edb::viterbi_schedule(nobs.data(), (int)nobs.size(), sched_groups, sched_ctas, pt.sched_warps, begin, items);
        for (size_t i = 0; i < items.size(); i += 2) items[i] = pt.chains[items[i]];      // index in the part -> chain id
        // trailing sweep CTAs without work are not launched
        int n_ctas = sched_ctas;
        while (n_ctas > 1 && begin[(size_t)(n_ctas - 1) * pt.sched_warps] == begin[(size_t)n_ctas * pt.sched_warps]) n_ctas--;
        pt.sched_ctas = n_ctas;
        if (int rc = ensure(pt.sched_begin, begin.size() * 4)) return rc;
        if (int rc = ensure(pt.sched_items, items.size() * 4 + 8)) return rc;
        CU(cudaMemcpyAsync(pt.sched_begin.p, begin.data(), begin.size() * 4, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(pt.sched_items.p, items.data(), items.size() * 4, cudaMemcpyHostToDevice, st));
        CU(cudaStreamSynchronize(st));          // the vectors go out of scope
        pt.sched_groups = key;
        pt.sched_avail = avail_sms;
    }
    a.chain_list = (const int32_t*)pt.chain_list.p;
    a.n_list = (int)pt.chains.size();
    a.max_list_tiles = pt.max_tiles;
    if (!a.only_bad) a.flat_records = (int)pt.chains.size() == c->n_chains ? (int64_t)c->total_tiles * a.groups : 0;
    a.warps_per_cta = pt.sched_warps;
    a.n_slots = pt.sched_ctas * pt.sched_warps;
    a.sched_begin = (const int32_t*)pt.sched_begin.p;
    a.sched_items = (const int32_t*)pt.sched_items.p;
    g_launches += edb::launch_viterbi(a, st);
    return check_kernel("viterbi");
}

// ---- segmented sweep (viterbi_seam.h) ------------------------------------------------------------------------
// Defaults: a warm-up of 2 tiles (32 observations; 8 already close every seam of the synthetic cohorts, and a seam that does
// not close only costs its chain the repair pass), pieces of at least 12 tiles.
constexpr int kSegWarm = 2, kSegMinPiece = 12;
constexpr int kChunkReserve = 40;        // SMs the emission of a middle chunk leaves to the Viterbi kernels of the chunk before it
// Whether to cut the chains: when an even share of all tiles per sweep warp (plus its warm-up) is well below what bounds
// the plain sweeps — the longest chain, or the whole lines packed onto the warps.
static bool use_segments(const edb200_cohort* c, int n_samples)
{
    if (c->struct_state != 1 || !c->seg_ok || c->opt_sweep == 1 || c->opt_segments == 0) return false;
    if (c->opt_segments == 1) return true;
    const int S = c->S, G = 32 / S;
    // measured: 3 and 5 states (256 x 200k x 5: sweep 1.51 -> 0.60 ms; 2,000 samples: 4.4 -> 4.3).  The 7-state instantiation
    // spills (255 registers) and runs six warps per CTA: 0.53 ms where the plain sweeps take 0.17 on the 512 x 5,000 panel.
    if (S == 7) return false;
    int64_t longest = 0, total = 0;
    for (const auto& cd : c->chains_h) {
        longest = std::max<int64_t>(longest, cd.nobs);
        total += cd.nobs;
    }
    const double slots = 4.0 * g.n_sms;
    const double lane = std::max(150.0 * longest, 165.0 * total * ((n_samples + G - 1) / G) / slots);
    const double tpc = std::max(205.0 * longest, 205.0 * total * ((n_samples + 31) / 32) / slots);
    const double min_piece = 16.0 * (c->opt_seg_min > 0 ? c->opt_seg_min : kSegMinPiece), warm = 16.0 * (c->opt_seg_warm > 0 ? c->opt_seg_warm : kSegWarm);
    // (two sweep warps per SM sub-partition: ~330 cycles per step and warp, twice the warps)
    const double seg_slots = (double)edb::viterbi_seg_warps(S) * g.n_sms;
    const double share = std::max((double)total * ((n_samples + 31) / 32) / seg_slots, std::min((double)longest, min_piece));
    // ... and only where a chain is long enough to bound the batch (panels of a few thousand bins are launch- and
    // tail-bound: the check and repair launches of the segmented pass cost more than the cut saves)
    return longest >= 4096 && 330.0 * (share + warm) < 0.8 * std::min(lane, tpc);
}

// all chains as one group, for `ns` samples and chunk slot `slot`
static int seg_part(edb200_cohort* c, int ns, int slot, edb200_cohort::Part** out)
{
    edb200_cohort::Part& pt = c->seg_parts[{ns, slot}];
    if (pt.chains.empty()) {
        pt.chains.resize(c->n_chains);
        for (int i = 0; i < c->n_chains; i++) {
            pt.chains[i] = i;
            pt.max_tiles = std::max(pt.max_tiles, edb::viterbi_chain_tiles(c->chains_h[i]));
        }
        if (int rc = ensure(pt.chain_list, pt.chains.size() * 4)) return rc;
        CU(cudaMemcpy(pt.chain_list.p, pt.chains.data(), pt.chains.size() * 4, cudaMemcpyHostToDevice));
    }
    *out = &pt;
    return 0;
}

// seg sweep, tilemap, trace, expand, check, and the (normally empty) repair pass over the chains of one chromosome group
// on `st`.  The pieces are cut for every SM: they are independent, so CTAs that find the SMs taken
// (by the emission of the next group) simply run when one frees up.
static int viterbi_segmented(edb200_cohort* c, edb200_cohort::Part& pt, edb::ViterbiArgs a, cudaStream_t st)
{
    const int S = c->S, ns = a.n_samples, n_g32 = edb::seg_n_g32(ns);
    const int warm = c->opt_seg_warm > 0 ? c->opt_seg_warm : kSegWarm;
    const int min_piece = c->opt_seg_min > 0 ? c->opt_seg_min : kSegMinPiece;
    constexpr int kCloseCap = 1 << 16;
    const int sw = c->opt_sweep_warps;
    const int kW = sw == 4 || ((sw == 6 || sw == 8) && sw <= edb::viterbi_seg_warps(S)) ? sw : edb::viterbi_seg_warps(S);
    edb200_cohort::SegPlan& sp = pt.seg;
    const long long key = ((long long)ns << 28) | ((long long)kW << 24) | ((long long)warm << 16) | (long long)std::min(min_piece, 65535);
    const int n_list = (int)pt.chains.size();
    if (sp.key != key) {
        std::vector<int32_t> tiles(n_list), begin, items, desc, first;
        for (int i = 0; i < n_list; i++) tiles[i] = edb::viterbi_chain_tiles(c->chains_h[pt.chains[i]]);
        edb::viterbi_cut_pieces(tiles.data(), n_list, n_g32, g.n_sms, kW, warm, min_piece, begin, items, desc, first);
        for (size_t i = 0; i < desc.size(); i += 4) desc[i] = pt.chains[desc[i]];          // index in the group -> chain id
        int n_ctas = g.n_sms;
        while (n_ctas > 1 && begin[(size_t)(n_ctas - 1) * kW] == begin[(size_t)n_ctas * kW]) n_ctas--;
        sp.ctas = n_ctas;
        sp.pieces = (int)(desc.size() / 4);
        if (int rc = ensure(sp.desc, desc.size() * 4 + 16)) return rc;
        if (int rc = ensure(sp.first, first.size() * 4)) return rc;
        if (int rc = ensure(sp.begin, begin.size() * 4)) return rc;
        if (int rc = ensure(sp.items, items.size() * 4 + 8)) return rc;
        if (int rc = ensure(sp.seam_in, (size_t)sp.pieces * S * 32 * 8 + 8)) return rc;
        if (int rc = ensure(sp.seam_out, (size_t)sp.pieces * S * 32 * 8 + 8)) return rc;
        if (int rc = ensure(sp.seam_mag, (size_t)sp.pieces * edb::kSeamWords * 32 * 4 + 8)) return rc;
        if (int rc = ensure(sp.close, (size_t)kCloseCap * 16)) return rc;
        if (int rc = ensure(sp.flags, edb::seg_flag_ints(c->n_chains, ns) * 4)) return rc;
        CU(cudaMemcpyAsync(sp.desc.p, desc.data(), desc.size() * 4, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(sp.first.p, first.data(), first.size() * 4, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(sp.begin.p, begin.data(), begin.size() * 4, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(sp.items.p, items.data(), items.size() * 4, cudaMemcpyHostToDevice, st));
        CU(cudaStreamSynchronize(st));          // the vectors go out of scope
        sp.key = key;
    }
    c->seg_used.push_back({&pt, ns});
    a.tpc = 1;
    a.seg = 1;
    a.seg_warm = warm;
    a.seg_desc = (const int4*)sp.desc.p;
    a.seg_first = (const int32_t*)sp.first.p;
    a.seam_in = (double*)sp.seam_in.p;
    a.seam_out = (double*)sp.seam_out.p;
    a.seam_mag = (unsigned*)sp.seam_mag.p;
    a.seg_close = (int4*)sp.close.p;
    a.seg_close_cap = kCloseCap;
    a.seg_flags = (int32_t*)sp.flags.p;
    a.seg_force_repair = c->opt_seg_repair;
    a.chain_list = (const int32_t*)pt.chain_list.p;
    a.n_list = n_list;
    a.max_list_tiles = pt.max_tiles;
    a.flat_records = n_list == c->n_chains ? (int64_t)c->total_tiles * a.groups : 0;
    a.warps_per_cta = kW;
    a.n_slots = sp.ctas * kW;
    a.sched_begin = (const int32_t*)sp.begin.p;
    a.sched_items = (const int32_t*)sp.items.p;
    CU(cudaMemsetAsync(sp.flags.p, 0, edb::seg_flag_ints(c->n_chains, ns) * 4, st));
    g_launches += edb::launch_viterbi(a, st);
    edb::prof_mark("viterbi_seg_check", st);
    g_launches += edb::launch_viterbi_seg_check(a, sp.pieces, st);
    edb::prof_mark(nullptr, st);
    if (int rc = check_kernel("viterbi (segmented)")) return rc;
    // repair pass: the plain exact sweep and its post-processing over the refused chains (normally none: four launches
    // that find nothing to do)
    a.seg = 0;
    a.only_bad = 1;
    return viterbi_part(c, pt, a, 0, st, 0, 4);
}

static int call_summary(edb200_cohort* c, const edb200_batch* b, bool stats, bool cor, cudaStream_t st)
{
    stats = stats && b->call_stats;
    cor = cor && b->cor;
    if (!stats && !cor) return 0;
    if (stats && (!b->calls || !b->ncalls || b->call_cap < 1))
        return fail(EDB200_ERR_ARG, "CallCNVs post-processing needs the calls / ncalls of a Viterbi pass");
    const int S = c->S;
    edb::CallSummaryArgs a{};
    a.n_samples = b->n_samples;
    a.n_states = S;
    a.n_bins = c->n_bins;
    a.counts = edb::CountsView{b->observed, b->obs_stride, b->reference, b->ref_stride, 0};
    a.expected = b->expected;
    a.expected_stride = b->per_bin_stride;
    a.ll = b->ll;
    a.ll_sample_stride = (int64_t)S * b->ll_stride;
    a.ll_state_stride = b->ll_stride;
    for (int j = 0; j < S; j++) a.perm[j] = c->perm[j];
    a.calls = b->calls;
    a.ncalls = b->ncalls;
    a.call_cap = b->call_cap;
    a.stats = stats ? b->call_stats : nullptr;
    a.cor = cor ? b->cor : nullptr;
    g_launches += edb::launch_call_summary(a, st);
    return check_kernel("call_summary");
}

int edb200_cohort_run_device(edb200_cohort* c, const edb200_batch* b, int what, int emission_mode, void* cuda_stream)
{
    if (int rc = need_ctx()) return rc;
    if (!c || !b) return fail(EDB200_ERR_ARG, "null argument");
    if (b->n_samples <= 0) return 0;
    if (!b->ll || !b->observed || !b->reference || !b->phi || !b->expected) return fail(EDB200_ERR_ARG, "null device pointer in batch");
    if (b->obs_stride < c->n_bins || b->ll_stride < c->n_bins) return fail(EDB200_ERR_ARG, "stride smaller than n_bins");
    cudaStream_t st = (cudaStream_t)cuda_stream;
    edb::ViterbiArgs va{};
    alignas(64) CUtensorMap ll_map[2];
    if ((what & 2)) if (int rc = viterbi_prepare(c, b, va, ll_map)) return rc;
    if ((what & 2) && !c->in_host_call) c->seg_used.clear();

    // Device-resident batches run emission, then Viterbi (1 part) unless EDB200_PARTS asks otherwise: both kernels own
    // their SM's shared memory, so overlapping them takes SMs away from the sweep's critical chains — measured slower
    // (3.0 -> 3.5 ms at 256 x 200k x 5).  The host-pointer call pipelines over PCIe instead (edb200_cohort_run_host).
    if (b->per_bin_stride && b->per_bin_stride < c->n_bins) return fail(EDB200_ERR_ARG, "per_bin_stride smaller than n_bins");
    const int n_parts = (what & 3) == 3 && !b->per_bin_stride ? pick_parts(c, emission_mode, 1) : 1;
    if (int rc = build_plan(c, n_parts)) return rc;
    std::vector<edb200_cohort::Part>& plan = c->plans[n_parts];
    if (plan.size() <= 1) {
        // ---- one pass: emission, then Viterbi, on the caller's stream
        if (what & 1) {
            if (c->consts_first < 0)
                if (int rc = state_setup(c, b, st)) return rc;
            edb::BinRanges all{};
            all.n = 1;
            all.b0[0] = 0;
            all.b1[0] = c->n_bins;
            if (int rc = emission_part(c, b, all, true, emission_mode, 0, st, c->emission_sms)) return rc;
        }
        if (what & 2) {
            // 0: one pass (tests, experiments).  The thread-per-chain sweep is split only while there are fewer work items than
            // sweep warps on the GPU (the passes then just let the short chromosomes' post-processing start early); beyond
            // that the work is throughput-bound and one balanced launch is better (2,000 samples: 5.6 -> see DESIGN.md)
            if (use_segments(c, b->n_samples)) {
                edb200_cohort::Part* sp = nullptr;
                if (int rc = seg_part(c, b->n_samples, c->seg_slot, &sp)) return rc;
                if (int rc = viterbi_segmented(c, *sp, va, st)) return rc;
            } else {
            const bool split = c->opt_vsplit != 0 && (c->opt_vsplit == 1 || !va.tpc || (int64_t)c->n_chains * ((b->n_samples + 31) / 32) <= 4LL * g.n_sms);
            if (split && c->n_chains >= 4 && va.groups >= 8)
                if (int rc = build_plan(c, 0)) return rc;
            // (chains of near-equal length — small panels — all fall into the first group: nothing to split)
            if (split && c->n_chains >= 4 && va.groups >= 8 && c->plans[0].size() == 2) {
                // The longest chromosomes' sweep is the critical path; everything behind a sweep (tilemap, trace, expand)
                // scales with the chains it covers.  Two concurrent passes — {longest chains} on a few SMs of their own,
                // {all others} on the rest — leave only the longest chains' own post-processing behind the critical sweep.
                std::vector<edb200_cohort::Part>& vp = c->plans[0];
                const int64_t items0 = (int64_t)vp[0].chains.size() * va.groups;
                const int ctas0 = (int)std::min<int64_t>(g.n_sms / 2, (items0 + crit_warps(c) - 1) / crit_warps(c));
                CU(cudaEventRecord(g.ev_fork, st));
                for (int p = 0; p < 2; p++) {
                    if (vp[p].chains.empty()) continue;
                    CU(cudaStreamWaitEvent(g.s_vit[p], g.ev_fork, 0));
                    if (int rc = viterbi_part(c, vp[p], va, p == 0 ? 1 : 0, g.s_vit[p], p == 0 ? ctas0 : g.n_sms - ctas0)) return rc;
                    CU(cudaEventRecord(g.ev_vit[p], g.s_vit[p]));
                    CU(cudaStreamWaitEvent(st, g.ev_vit[p], 0));
                }
            } else if (int rc = viterbi_part(c, plan[0], va, 0, st)) return rc;
            }
        }
    } else {
        // ---- chromosome-group pipeline: the emission of group p+1 runs while group p is being swept.  Forked from
        // and joined back into the caller's stream, so the call keeps its "enqueue on cuda_stream" contract.
        CU(cudaEventRecord(g.ev_fork, st));
        CU(cudaStreamWaitEvent(g.s_em, g.ev_fork, 0));
        if (int rc = state_setup(c, b, g.s_em)) return rc;
        for (size_t p = 0; p < plan.size(); p++) {
            if (int rc = emission_part(c, b, plan[p].ranges, false, emission_mode, p == 0 ? 1 : 2, g.s_em)) return rc;
            CU(cudaEventRecord(g.ev_em[p], g.s_em));
            CU(cudaStreamWaitEvent(g.s_vit[p], g.ev_em[p], 0));
            // first part: one warp per sub-partition on SMs of its own; later parts spread over the others
            const int64_t items0 = (int64_t)plan[0].chains.size() * va.groups;
            const int ctas0 = (int)std::min<int64_t>(g.n_sms / 2, (items0 + crit_warps(c) - 1) / crit_warps(c));
            if (int rc = viterbi_part(c, plan[p], va, p == 0 ? 1 : 0, g.s_vit[p], p == 0 ? ctas0 : g.n_sms - ctas0)) return rc;
            CU(cudaEventRecord(g.ev_vit[p], g.s_vit[p]));
        }
        for (size_t p = 0; p < plan.size(); p++) CU(cudaStreamWaitEvent(st, g.ev_vit[p], 0));
    }
    if (what & 2) {
        g_launches += edb::launch_viterbi_compact(va, st);
        if (int rc = check_kernel("viterbi_compact")) return rc;
    }
    if (what & 4)
        if (int rc = call_summary(c, b, true, true, st)) return rc;
    return 0;
}

// ---- C_hmm-shaped entry point ------------------------------------------------------------------------------------
// R calls C_hmm once per sample and chromosome with the SAME positions, transition matrix and expected length for every
// sample of a run (R/class_definition.R:354-369).  The host-libm log-transition table of a (positions, T, L) triple, its
// structured view and everything sized by the chain live in a one-chain cohort that is kept — most recently used first, 64
// of them — so a repeat costs an emission upload, the kernels and a path download; and the call goes through the cohort's
// own Viterbi pass: for the CallCNVs matrix and a genome-scale chromosome that is the segmented sweep (section 4.2b of
// DESIGN.md), ~0.1 ms instead of 19,803 dependent steps (1.5 ms) for chromosome 1.
}  // extern "C"
namespace {
struct HmmEntry {
    int S = 0, nobs = 0;
    double L = 0;
    std::vector<double> T;
    std::vector<int32_t> pos;
    edb200_cohort* co = nullptr;
};
std::vector<HmmEntry> g_hmm_cache;           // most recently used first
constexpr size_t kHmmCacheCap = 64;

int hmm_entry(int S, int nobs, const double* T, const int32_t* pos, double L, edb200_cohort** out)
{
    for (size_t i = 0; i < g_hmm_cache.size(); i++) {
        HmmEntry& e = g_hmm_cache[i];
        if (e.S == S && e.nobs == nobs && e.L == L && memcmp(e.T.data(), T, (size_t)S * S * 8) == 0 && memcmp(e.pos.data(), pos, (size_t)nobs * 4) == 0) {
            if (i) std::rotate(g_hmm_cache.begin(), g_hmm_cache.begin() + i, g_hmm_cache.begin() + i + 1);
            *out = g_hmm_cache[0].co;
            return 0;
        }
    }
    edb200_cohort* c = new edb200_cohort();
    auto bail = [&](int rc) {
        destroy_cohort_locked(c);
        return rc;
    };
    c->n_bins = nobs;                                        // (the path row: one byte per observation)
    c->n_chains = 1;
    c->S = S;
    c->L = L;
    memcpy(c->T, T, (size_t)S * S * 8);
    for (int j = 0; j < S; j++) c->perm[j] = j;              // C_hmm's columns are the HMM's states as they are
    edb::ChainDesc cd{};
    cd.nobs = nobs;
    cd.n_em = nobs - 1;
    cd.out_first = 0;
    cd.out_last = nobs - 1;
    c->chains_h.assign(1, cd);
    c->total_rows = nobs;
    c->total_tiles = edb::viterbi_chain_tiles(cd);
    const int pitch = edb::viterbi_lt_pitch(S);
    c->lt_pitch = pitch;
    const size_t n_rows = (size_t)nobs + edb::viterbi_tile();
    std::vector<double> lt(n_rows * pitch, 0.0);
    edb::build_log_transition_rows(S, T, pos, nobs, L, lt.data(), pitch);
    edb::nan_to_neg_inf(lt.data(), lt.size());               // device copy only: a NaN term is "never selected", like -Inf (hmm.cpp:81)
    const int32_t tile_base0 = 0;
    if (int rc = ensure(c->lt, lt.size() * 8)) return bail(rc);
    if (int rc = ensure(c->chains, sizeof cd)) return bail(rc);
    if (int rc = ensure(c->tile_base, 4)) return bail(rc);
    if (cudaMemcpy(c->lt.p, lt.data(), lt.size() * 8, cudaMemcpyHostToDevice) != cudaSuccess || cudaMemcpy(c->chains.p, &cd, sizeof cd, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(c->tile_base.p, &tile_base0, 4, cudaMemcpyHostToDevice) != cudaSuccess)
        return bail(fail(EDB200_ERR_CUDA, "edb200_hmm: uploading the log-transition table failed"));
    if (int rc = ensure_struct(c)) return bail(rc);          // structured rows when the matrix has the CallCNVs form (else the general sweep)
    if (g_hmm_cache.size() >= kHmmCacheCap) {
        destroy_cohort_locked(g_hmm_cache.back().co);
        g_hmm_cache.pop_back();
    }
    HmmEntry e;
    e.S = S, e.nobs = nobs, e.L = L, e.co = c;
    e.T.assign(T, T + (size_t)S * S);
    e.pos.assign(pos, pos + nobs);
    g_hmm_cache.insert(g_hmm_cache.begin(), std::move(e));
    *out = c;
    return 0;
}
void release_hmm_cache()
{
    for (HmmEntry& e : g_hmm_cache) destroy_cohort_locked(e.co);
    g_hmm_cache.clear();
}
}  // namespace
extern "C" {

int edb200_hmm(int32_t nstates, int32_t nobs, const double* transitions, const double* probabilities,
               const int32_t* positions, double expected_length, int32_t* path_out, int32_t* calls_out,
               int32_t call_cap, int32_t* ncalls_out)
{
    if (int rc = need_ctx()) return rc;
    if (nstates < 2 || nstates > EDB200_MAX_STATES) return fail(EDB200_ERR_NSTATES, "nstates=%d not in [2,%d]", nstates, EDB200_MAX_STATES);
    if (nobs < 1 || !transitions || !probabilities || !positions || !path_out || !ncalls_out || call_cap < 0 || (call_cap && !calls_out))
        return fail(EDB200_ERR_ARG, "bad argument");
    std::lock_guard<std::mutex> lk(g_mu);
    cudaStream_t st = g.stream;
    const int S = nstates;
    const int cap = call_cap > 0 ? call_cap : 1;
    edb200_cohort* co = nullptr;
    if (int rc = hmm_entry(S, nobs, transitions, positions, expected_length, &co)) return rc;
    const int64_t nobs_p = ((int64_t)nobs + 15) & ~(int64_t)15;     // emission rows padded to whole 128-byte lines
    if (int rc = ensure(cs.ll, (size_t)nobs_p * S * 8)) return rc;
    if (int rc = ensure(cs.path, (size_t)nobs_p)) return rc;
    if (int rc = ensure(cs.calls, (size_t)cap * 16)) return rc;
    if (int rc = ensure(cs.ncalls, 4)) return rc;
    CU(cudaMemcpy2DAsync(cs.ll.p, nobs_p * 8, probabilities, (size_t)nobs * 8, (size_t)nobs * 8, S, cudaMemcpyHostToDevice, st));
    edb200_batch b{};
    b.n_samples = 1;
    // (the Viterbi pass reads the likelihood rows only; the count / fit slots of the batch are not touched)
    b.observed = b.reference = (const int32_t*)cs.ll.p;
    b.phi = b.expected = (const double*)cs.ll.p;
    b.obs_stride = nobs;
    b.ll = (double*)cs.ll.p;
    b.ll_stride = nobs_p;
    b.path = (int8_t*)cs.path.p;
    b.path_stride = nobs_p;
    b.calls = (int32_t*)cs.calls.p;
    b.ncalls = (int32_t*)cs.ncalls.p;
    b.call_cap = cap;
    if (int rc = edb200_cohort_run_device(co, &b, 2, EDB200_EMISSION_AUTO, st)) return rc;

    std::vector<int8_t> p8(nobs);
    int32_t nc = 0;
    CU(cudaMemcpyAsync(p8.data(), cs.path.p, nobs, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(&nc, cs.ncalls.p, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    for (int i = 0; i < nobs; i++) path_out[i] = p8[i];
    *ncalls_out = nc;
    const int have = nc < call_cap ? nc : call_cap;
    if (have > 0) CU(cudaMemcpy(calls_out, cs.calls.p, (size_t)have * 16, cudaMemcpyDeviceToHost));
    return nc > call_cap ? EDB200_WARN_CALLCAP : 0;
}

// ---- CUDA-graph replay of a device-resident batch (small panels: the step is bound by its ~10 launches) ----
int edb200_cohort_capture_device(edb200_cohort* c, const edb200_batch* b, int what, int emission_mode, edb200_graph** out)
{
    if (int rc = need_ctx()) return rc;
    if (!c || !b || !out) return fail(EDB200_ERR_ARG, "null argument");
    *out = nullptr;
    if (b->n_samples <= 0) return fail(EDB200_ERR_ARG, "empty batch");
    if (edb::g_timer) return fail(EDB200_ERR_ARG, "edb200_profile is on: its event marks cannot be part of a captured graph");
    cudaStream_t st = g.stream;
    // 1. one plain run: workspaces are sized, sweep schedules uploaded, argument errors surface un-captured
    int warn = edb200_cohort_run_device(c, b, what, emission_mode, st);
    if (warn & (EDB200_ERR_NSTATES | EDB200_ERR_CUDA | EDB200_ERR_ARG)) return warn;
    CU(cudaStreamSynchronize(st));
    // 2. the same enqueue again, recorded.  Thread-local mode: an allocation or a synchronisation inside the recorded
    // call (there is none after step 1) fails the capture instead of being silently left out of the graph.
    const long long gen = g_realloc_gen.load();
    const long long launches0 = g_launches.load();
    CU(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    const int rc = edb200_cohort_run_device(c, b, what, emission_mode, st);
    cudaGraph_t graph = nullptr;
    const cudaError_t e_end = cudaStreamEndCapture(st, &graph);
    const int kernels = (int)(g_launches.load() - launches0);
    g_launches.store(launches0);                                    // nothing ran
    if (rc & (EDB200_ERR_NSTATES | EDB200_ERR_CUDA | EDB200_ERR_ARG)) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        return rc;
    }
    if (e_end != cudaSuccess || !graph) {
        cudaGetLastError();
        return fail(EDB200_ERR_CUDA, "stream capture failed: %s", cudaGetErrorString(e_end));
    }
    if (gen != g_realloc_gen.load()) {
        cudaGraphDestroy(graph);
        return fail(EDB200_ERR_CUDA, "internal: a workspace moved during capture");
    }
    cudaGraphExec_t exec = nullptr;
    const cudaError_t e_inst = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e_inst != cudaSuccess) return fail(EDB200_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e_inst));
    edb200_graph* gr = new edb200_graph;
    gr->exec = exec;
    gr->cohort = c;
    gr->realloc_gen = gen;
    gr->kernels = kernels;
    c->graphs.push_back(gr);
    *out = gr;
    return warn;
}

int edb200_graph_launch(edb200_graph* gr, void* cuda_stream)
{
    if (int rc = need_ctx()) return rc;
    if (!gr) return fail(EDB200_ERR_ARG, "null graph");
    if (!gr->exec) return fail(EDB200_ERR_ARG, "the cohort this graph was captured from has been destroyed");
    if (gr->realloc_gen != g_realloc_gen.load())
        return fail(EDB200_ERR_ARG, "a library workspace was re-allocated after this graph was captured (a larger batch ran since); capture it again");
    CU(cudaGraphLaunch(gr->exec, (cudaStream_t)cuda_stream));
    g_launches += gr->kernels;
    return 0;
}

void edb200_graph_destroy(edb200_graph* gr)
{
    if (!gr) return;
    std::lock_guard<std::mutex> lk(g_mu);
    if (gr->exec) {
        cudaDeviceSynchronize();
        cudaGraphExecDestroy(gr->exec);
    }
    if (gr->cohort) {
        auto& v = gr->cohort->graphs;
        v.erase(std::remove(v.begin(), v.end(), gr), v.end());
    }
    delete gr;
}

// transition matrices of the grid: the CallCNVs matrix for every tp, or the cohort's own matrix when tp_grid is null
static int forward_common(edb200_cohort* c, const double* d_ll, int64_t ll_stride, int ns, const double* tp_grid, int n_grid,
                          double* d_loglik, int32_t* d_best, cudaStream_t st)
{
    const int S = c->S;
    if (!tp_grid) n_grid = 1;
    if (n_grid < 1 || n_grid > 4096) return fail(EDB200_ERR_ARG, "n_grid=%d not in [1, 4096]", n_grid);
    std::vector<double> grid((size_t)n_grid * S * S);
    for (int gi = 0; gi < n_grid; gi++) {
        if (tp_grid) edb::callcnvs_transitions(S, tp_grid[gi], grid.data() + (size_t)gi * S * S);
        else memcpy(grid.data(), c->T, S * S * 8);
    }
    if (int rc = ensure(c->fw_grid, grid.size() * 8)) return rc;
    if (int rc = ensure(c->fw_chain, (size_t)ns * c->n_chains * n_grid * 8)) return rc;
    CU(cudaMemcpyAsync(c->fw_grid.p, grid.data(), grid.size() * 8, cudaMemcpyHostToDevice, st));
    CU(cudaStreamSynchronize(st));          // `grid` goes out of scope
    edb::ForwardArgs a{};
    a.chains = (const edb::ChainDesc*)c->chains.p;
    a.n_chains = c->n_chains;
    a.n_samples = ns;
    a.n_states = S;
    a.n_grid = n_grid;
    a.ll = d_ll;
    a.ll_sample_stride = (int64_t)S * ll_stride;
    a.ll_state_stride = ll_stride;
    for (int j = 0; j < S; j++) a.perm[j] = c->perm[j];
    a.decay = (const double*)c->decay.p;
    a.T_grid = (const double*)c->fw_grid.p;
    a.tail_other = -100.0;                                     // R/class_definition.R:364
    a.chain_loglik = (double*)c->fw_chain.p;
    a.loglik = d_loglik;
    a.best = d_best;
    g_launches += edb::launch_forward(a, st);
    return check_kernel("forward");
}

int edb200_cohort_forward_device(edb200_cohort* c, const edb200_batch* b, const double* tp_grid, int32_t n_grid,
                                 double* loglik, int32_t* best, void* cuda_stream)
{
    if (int rc = need_ctx()) return rc;
    if (!c || !b || !b->ll || !loglik || b->n_samples < 1) return fail(EDB200_ERR_ARG, "bad argument");
    if (b->ll_stride < c->n_bins) return fail(EDB200_ERR_ARG, "stride smaller than n_bins");
    return forward_common(c, b->ll, b->ll_stride, b->n_samples, tp_grid, n_grid, loglik, best, (cudaStream_t)cuda_stream);
}

int edb200_cohort_forward_last(edb200_cohort* c, const double* tp_grid, int32_t n_grid, double* loglik, int32_t* best)
{
    if (int rc = need_ctx()) return rc;
    if (!c || !loglik) return fail(EDB200_ERR_ARG, "null argument");
    if (c->last_host_samples < 1) return fail(EDB200_ERR_ARG, "no likelihoods resident: call edb200_cohort_run_host first");
    std::lock_guard<std::mutex> lk(g_mu);
    cudaStream_t st = g.stream;
    const int ns = c->last_host_samples, ng = tp_grid ? n_grid : 1;
    const int64_t nbp = (c->n_bins + 15) & ~(int64_t)15;
    if (ng < 1 || ng > 4096) return fail(EDB200_ERR_ARG, "n_grid=%d not in [1, 4096]", ng);
    if (int rc = ensure(c->fw_out, (size_t)ns * ng * 8)) return rc;
    if (int rc = ensure(c->fw_best, (size_t)ns * 4)) return rc;
    if (int rc = forward_common(c, (const double*)c->h_ll.p, nbp, ns, tp_grid, ng, (double*)c->fw_out.p, (int32_t*)c->fw_best.p, st)) return rc;
    CU(cudaMemcpyAsync(loglik, c->fw_out.p, (size_t)ns * ng * 8, cudaMemcpyDeviceToHost, st));
    if (best) CU(cudaMemcpyAsync(best, c->fw_best.p, (size_t)ns * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return 0;
}

int edb200_cohort_run_host(edb200_cohort* c, const edb200_batch* b, int emission_mode)
{
    if (int rc = need_ctx()) return rc;
    if (!c || !b) return fail(EDB200_ERR_ARG, "null argument");
    if (b->n_samples <= 0) return 0;
    std::lock_guard<std::mutex> lk(g_mu);
    struct HostCall {
        edb200_cohort* c;
        explicit HostCall(edb200_cohort* c_) : c(c_) { c->in_host_call = true; c->seg_used.clear(); }
        ~HostCall() { c->in_host_call = false; c->consts_first = -1; c->seg_slot = 0; c->vit_slots = 1; c->vit_slot = 0; c->vit_slot_samples = 0; }
    } host_call(c);
    cudaStream_t st = g.stream;
    const int S = c->S, ns = b->n_samples;
    const int64_t nb = c->n_bins;
    const bool shared_ref = b->ref_stride == 0;
    const int cap = b->call_cap > 0 ? b->call_cap : 1;
    const int64_t nbp = (nb + 15) & ~(int64_t)15;          // device rows padded to whole 128-byte lines
    int rc = 0;
    if ((rc = ensure(c->h_obs, (size_t)ns * nb * 4)) || (rc = ensure(c->h_ref, (size_t)(shared_ref ? 1 : ns) * nb * 4)) ||
        (rc = ensure(c->h_phi, (size_t)ns * (b->per_bin_stride ? nb : 1) * 8)) || (rc = ensure(c->h_exp, (size_t)ns * (b->per_bin_stride ? nb : 1) * 8)) || (rc = ensure(c->h_ll, (size_t)ns * S * nbp * 8)) ||
        (rc = ensure(c->h_path, (size_t)ns * nb)) || (rc = ensure(c->h_calls, (size_t)ns * cap * 16)) ||
        (rc = ensure(c->h_ncalls, ns * 4)) || (rc = ensure(c->h_stats, b->call_stats ? (size_t)ns * cap * 24 : 8)) ||
        (rc = ensure(c->h_cor, ns * 8)))
        return rc;

    edb200_batch d = *b;
    d.observed = (const int32_t*)c->h_obs.p;
    d.obs_stride = nb;
    d.reference = (const int32_t*)c->h_ref.p;
    d.ref_stride = shared_ref ? 0 : nb;
    d.phi = (const double*)c->h_phi.p;
    d.expected = (const double*)c->h_exp.p;
    d.per_bin_stride = b->per_bin_stride ? nb : 0;
    d.ll = (double*)c->h_ll.p;
    d.ll_stride = nbp;
    d.path = (int8_t*)c->h_path.p;
    d.path_stride = nb;
    d.calls = (int32_t*)c->h_calls.p;
    d.ncalls = (int32_t*)c->h_ncalls.p;
    d.call_cap = cap;
    d.call_stats = b->call_stats ? (double*)c->h_stats.p : nullptr;
    d.cor = b->cor ? (double*)c->h_cor.p : nullptr;
    const bool want_vit = b->path || b->calls || b->ncalls || b->call_stats;
    c->last_host_samples = ns;

    // (the 12-bit layout goes up in whole rows: sample chunks, not chromosome groups)
    const int n_parts = want_vit && !b->per_bin_stride && !b->observed12 ? pick_parts(c, emission_mode, Context::kMaxParts) : 1;
    if ((rc = build_plan(c, n_parts))) return rc;
    std::vector<edb200_cohort::Part>& plan = c->plans[n_parts];

    // 16-bit ingestion layout: the overflow list goes up once, every group's columns are widened behind their upload
    const bool p12 = b->observed12 != nullptr;
    const bool u16 = !p12 && b->observed16 != nullptr;
    const int64_t row12 = (nb + 1) / 2 * 3, r12 = (row12 + 3) & ~(int64_t)3;         // bytes of a 12-bit row; its pitch on the device
    if (p12) {
        if (b->obs12_stride < row12 || (b->obs12_stride & 3) || (reinterpret_cast<uintptr_t>(b->observed12) & 3) || b->n_overflow < 0 ||
            (b->n_overflow > 0 && (!b->overflow_index || !b->overflow_value)))
            return fail(EDB200_ERR_ARG, "bad 12-bit count layout (stride: a multiple of 4, >= 3 * ceil(n_bins / 2); rows 4-byte aligned; overflow list)");
        if ((rc = ensure(c->h_obs16, (size_t)ns * r12)) || (rc = ensure(c->h_ovf_i, (size_t)b->n_overflow * 8 + 8)) ||
            (rc = ensure(c->h_ovf_v, (size_t)b->n_overflow * 4 + 8)))
            return rc;
    } else if (u16) {
        if (b->obs16_stride < nb || b->n_overflow < 0 || (b->n_overflow > 0 && (!b->overflow_index || !b->overflow_value)))
            return fail(EDB200_ERR_ARG, "bad 16-bit count layout (stride / overflow list)");
        if ((rc = ensure(c->h_obs16, (size_t)ns * nb * 2)) || (rc = ensure(c->h_ovf_i, (size_t)b->n_overflow * 8 + 8)) ||
            (rc = ensure(c->h_ovf_v, (size_t)b->n_overflow * 4 + 8)))
            return rc;
    } else if (!b->observed)
        return fail(EDB200_ERR_ARG, "no test counts in batch (observed, observed16 or observed12)");

    // ---- sample-chunk pipeline (segmented sweeps).  With the chains cut into pieces no chromosome is a critical path any
    // more, so the batch goes through in chunks of SAMPLES: chunk k+1 uploads (contiguous rows) while chunk k runs emission,
    // then its segmented sweep, on ONE compute stream — both kernels fill every SM; side by side they only trade SMs.
    // T ~ U / k + G + k * o  (U upload, G device time of the batch, o ~ 0.1-0.25 ms of launches, tails and partly filled rounds
    // per chunk): k ~ sqrt(U / o).
    int seg_k = 0, seg_per = 0;
    bool tables_copied = false;                              // call tables / per-call columns already went back chunk by chunk
    if (want_vit && !b->per_bin_stride && use_table(c, emission_mode) && c->opt_parts == 0) {
        if ((rc = ensure_struct(c))) return rc;
        // (the 12-bit layout is counted like the 16-bit one: measured best is 4 chunks for both at 256 x 200k — the call is GPU-bound)
        const double upload_ms = (double)ns * nb * (p12 || u16 ? 2.0 : 4.0) / 53e6, per_chunk_ms = 0.12;
        int k = c->opt_chunks > 0 ? c->opt_chunks : (int)std::lround(std::sqrt(upload_ms / per_chunk_ms));       // (4 at 256 x 200k, 16-bit: measured best with 56 SMs reserved)
        k = std::max(1, std::min({k, Context::kMaxParts, ns / 24}));
        // chunk size: a whole number of rounds of the emission kernel's (sample, state) items over the SMs (64 samples x 5
        // states on 148 SMs are 2.16 rounds and cost 3)
        // balanced chunks; when a chunk is a little more than a whole number of rounds of the emission kernel's (sample, state)
        // items over the SMs (64 samples x 5 states on 148 SMs: 2.16 rounds, which cost 3), one more chunk of whole rounds
        int per = (ns + k - 1) / k;
        const double rounds = (double)per * S / g.n_sms;
        if (rounds > 1.0 && rounds - std::floor(rounds) < 0.25 && k < Context::kMaxParts) {
            const int per2 = (int)std::floor(rounds) * g.n_sms / S;
            if ((ns + per2 - 1) / per2 <= k + 1 && ns - per2 * k >= per2 / 2) per = per2;
        }
        if (use_segments(c, per)) {
            seg_per = per;
            seg_k = (ns + per - 1) / per;
        }
    }
    if (seg_k > 0) {
        cudaStream_t sc = g.s_copy, sx = g.s_em, sw = g.s_widen, sr = g.s_cor;
        if (shared_ref) CU(cudaMemcpyAsync(c->h_ref.p, b->reference, nb * 4, cudaMemcpyHostToDevice, sc));
        CU(cudaMemcpyAsync(c->h_phi.p, b->phi, ns * 8, cudaMemcpyHostToDevice, sc));
        CU(cudaMemcpyAsync(c->h_exp.p, b->expected, ns * 8, cudaMemcpyHostToDevice, sc));
        if ((u16 || p12) && b->n_overflow > 0) {
            CU(cudaMemcpyAsync(c->h_ovf_i.p, b->overflow_index, (size_t)b->n_overflow * 8, cudaMemcpyHostToDevice, sc));
            CU(cudaMemcpyAsync(c->h_ovf_v.p, b->overflow_value, (size_t)b->n_overflow * 4, cudaMemcpyHostToDevice, sc));
        }
        // the per-state constants of every sample need phi and expected only: built once, ahead of the counts (a setup launch
        // per chunk in front of its emission kernel waits for an SM beside the running kernels; worth ~1 % of the call)
        if ((rc = state_setup(c, &d, sc))) return rc;
        edb::BinRanges all{};
        all.n = 1;
        all.b0[0] = 0;
        all.b1[0] = nb;
        for (int k = 0, s0 = 0; s0 < ns; k++, s0 += seg_per) {
            const int cnt = std::min(seg_per, ns - s0);
            c->consts_first = s0;
            edb::prof_mark("h2d_counts", sc);
            if (p12)
                CU(cudaMemcpy2DAsync((uint8_t*)c->h_obs16.p + (size_t)s0 * r12, r12, b->observed12 + (size_t)s0 * b->obs12_stride, b->obs12_stride,
                                     row12, cnt, cudaMemcpyHostToDevice, sc));
            else if (u16)
                CU(cudaMemcpy2DAsync((uint16_t*)c->h_obs16.p + (size_t)s0 * nb, nb * 2, b->observed16 + (size_t)s0 * b->obs16_stride, b->obs16_stride * 2,
                                     nb * 2, cnt, cudaMemcpyHostToDevice, sc));
            else
                CU(cudaMemcpy2DAsync((int32_t*)c->h_obs.p + (size_t)s0 * nb, nb * 4, b->observed + (size_t)s0 * b->obs_stride, b->obs_stride * 4,
                                     nb * 4, cnt, cudaMemcpyHostToDevice, sc));
            if (!shared_ref)
                CU(cudaMemcpy2DAsync((int32_t*)c->h_ref.p + (size_t)s0 * nb, nb * 4, b->reference + (size_t)s0 * b->ref_stride, b->ref_stride * 4,
                                     nb * 4, cnt, cudaMemcpyHostToDevice, sc));
            edb::prof_mark(nullptr, sc);
            CU(cudaEventRecord(g.ev_copy[k], sc));
            if (u16 || p12) {
                // widened on a stream of its own: on the copy stream the kernel — whose blocks wait for an SM while an emission
                // grid is resident — would hold up the next chunk's upload (0.1 ms per chunk)
                CU(cudaStreamWaitEvent(sw, g.ev_copy[k], 0));
                if (p12)
                    g_launches += edb::launch_unpack12_counts((const uint8_t*)c->h_obs16.p + (size_t)s0 * r12, r12, (int32_t*)c->h_obs.p + (size_t)s0 * nb, nb, cnt, nb, sw);
                else
                    g_launches += edb::launch_widen_counts((const uint16_t*)c->h_obs16.p + (size_t)s0 * nb, nb, (int32_t*)c->h_obs.p + (size_t)s0 * nb, nb, cnt, nb,
                                                           all, nullptr, nullptr, 0, sw);
                g_launches += edb::launch_patch_overflow((int32_t*)c->h_obs.p, nb, nb, all, (const int64_t*)c->h_ovf_i.p, (const int32_t*)c->h_ovf_v.p,
                                                         b->n_overflow, sw, s0, s0 + cnt);
                CU(cudaEventRecord(g.ev_copy[k], sw));
            }
            CU(cudaStreamWaitEvent(sx, g.ev_copy[k], 0));
            edb200_batch e = d;
            e.n_samples = cnt;
            e.observed = d.observed + (size_t)s0 * nb;
            if (!shared_ref) e.reference = d.reference + (size_t)s0 * nb;
            e.phi = d.phi + s0;
            e.expected = d.expected + s0;
            e.ll = d.ll + (size_t)s0 * S * nbp;
            e.path = d.path + (size_t)s0 * nb;
            e.calls = d.calls + (size_t)s0 * cap * 4;
            e.ncalls = d.ncalls + s0;
            if (d.call_stats) e.call_stats = d.call_stats + (size_t)s0 * cap * 3;
            if (d.cor) e.cor = d.cor + s0;
            c->seg_slot = k;
            // cor(test, reference) needs the counts only: on a stream of its own behind the chunk's upload (one CTA per sample:
            // 0.18 ms for 64 samples if it ran in line, and on the copy stream it would hold up the next upload)
            if (d.cor) {
                CU(cudaStreamWaitEvent(sr, g.ev_copy[k], 0));
                if ((rc = call_summary(c, &e, false, true, sr))) return rc;
            }
            e.cor = nullptr;
            // two compute streams: the emission of chunk k+1 behind the emission of chunk k, the Viterbi of chunk k behind its
            // emission and behind the Viterbi of chunk k-1 (they share the back-pointer scratch) — the SMs a sweep's last CTAs
            // and the small kernels behind it leave idle go to the next emission
            // every chunk but the last leaves `reserve` SMs to the Viterbi kernels of the chunk before it: an emission CTA owns its
            // SM (208 KB of shared memory, 1,024 threads) for the whole launch, so behind a full-width emission grid the sweep, the
            // tile maps, ... of the previous chunk simply wait
            const bool last = s0 + cnt >= ns;
            c->emission_sms = last || k == 0 ? 0 : g.n_sms - (c->opt_reserve > 0 ? std::min(c->opt_reserve, g.n_sms / 2) : kChunkReserve);
            rc = edb200_cohort_run_device(c, &e, 1, emission_mode, sx);
            c->emission_sms = 0;
            if (rc) return rc;
            CU(cudaEventRecord(g.ev_em[k], sx));
            // a Viterbi stream and a scratch set per chunk: the sweep of chunk k fills the SMs while the small, latency-bound kernels
            // behind the sweeps of the chunks before it (tile maps, trace, expand, check, repair, compaction, call sums) are still
            // running — or still waiting for an SM beside the emission grid of chunk k+1 (with two alternating sets the sweep of the
            // last chunk sat 0.2 ms behind the starved expand kernel of the chunk two before it, profiles/r2j_e2e_timeline.txt)
            cudaStream_t sv = g.s_vit[k];
            c->vit_slots = seg_k;
            c->vit_slot = k;
            c->vit_slot_samples = seg_per;
            CU(cudaStreamWaitEvent(sv, g.ev_em[k], 0));
            if ((rc = edb200_cohort_run_device(c, &e, 2 | (d.call_stats ? 4 : 0), emission_mode, sv))) return rc;
            CU(cudaEventRecord(g.ev_vit[k], sv));
            // results drain per chunk on the copy-back stream, behind the chunk's Viterbi pass and call sums: the call tables and
            // per-call columns too (cap x 40 bytes per sample — 10.5 MB at 256 samples and a capacity of 1,024 calls: 0.19 ms if
            // they all went back behind the last chunk's kernels)
            CU(cudaStreamWaitEvent(g.stream2, g.ev_vit[k], 0));
            if (b->ll)
                CU(cudaMemcpy2DAsync(b->ll + (size_t)s0 * S * b->ll_stride, b->ll_stride * 8, e.ll, nbp * 8, nb * 8, (size_t)cnt * S,
                                     cudaMemcpyDeviceToHost, g.stream2));
            if (b->path)
                CU(cudaMemcpy2DAsync(b->path + (size_t)s0 * b->path_stride, b->path_stride, e.path, nb, nb, cnt, cudaMemcpyDeviceToHost, g.stream2));
            if (b->calls && b->call_cap > 0)
                CU(cudaMemcpyAsync(b->calls + (size_t)s0 * cap * 4, e.calls, (size_t)cnt * cap * 16, cudaMemcpyDeviceToHost, g.stream2));
            if (b->ncalls) CU(cudaMemcpyAsync(b->ncalls + s0, e.ncalls, (size_t)cnt * 4, cudaMemcpyDeviceToHost, g.stream2));
            if (b->call_stats && b->call_cap > 0)
                CU(cudaMemcpyAsync(b->call_stats + (size_t)s0 * cap * 3, e.call_stats, (size_t)cnt * cap * 24, cudaMemcpyDeviceToHost, g.stream2));
            tables_copied = true;
        }
        CU(cudaEventRecord(g.ev_setup, sr));
        CU(cudaStreamWaitEvent(st, g.ev_setup, 0));
        for (int k = 0; k < seg_k; k++) CU(cudaStreamWaitEvent(st, g.ev_vit[k], 0));
    } else if (plan.size() > 1) {
        // ---- chromosome-group pipeline over PCIe: the counts of the long chromosomes go up first; their emission and
        // sweep (the critical path) run while the other groups are still uploading; results drain per group.
        edb::ViterbiArgs va{};
        alignas(64) CUtensorMap ll_map[2];
        if ((rc = viterbi_prepare(c, &d, va, ll_map))) return rc;
        cudaStream_t sc = g.s_copy;
        if (shared_ref) CU(cudaMemcpyAsync(c->h_ref.p, b->reference, nb * 4, cudaMemcpyHostToDevice, sc));
        CU(cudaMemcpyAsync(c->h_phi.p, b->phi, ns * 8, cudaMemcpyHostToDevice, sc));
        CU(cudaMemcpyAsync(c->h_exp.p, b->expected, ns * 8, cudaMemcpyHostToDevice, sc));
        if (u16 && b->n_overflow > 0) {
            CU(cudaMemcpyAsync(c->h_ovf_i.p, b->overflow_index, (size_t)b->n_overflow * 8, cudaMemcpyHostToDevice, sc));
            CU(cudaMemcpyAsync(c->h_ovf_v.p, b->overflow_value, (size_t)b->n_overflow * 4, cudaMemcpyHostToDevice, sc));
        }
        CU(cudaEventRecord(g.ev_setup, sc));
        CU(cudaStreamWaitEvent(g.s_em, g.ev_setup, 0));
        if ((rc = state_setup(c, &d, g.s_em))) return rc;
        // SM budget.  With the one-thread-per-chain sweep the Viterbi work of the whole batch is (samples / 32 / 4 warps per CTA)
        // x bins x ~104 ns of SM time, the uploads take samples x bins x bytes / ~53 GB/s: their ratio — the SMs the sweeps
        // need in order to finish with the uploads — does not depend on the batch: ~22 for 16-bit counts, ~11 for 32-bit; the
        // longest chromosome's lane-per-state sweep (two warps per CTA) takes another 22 at 256 samples.
        // The emission launches of the later groups leave that many SMs alone (an emission CTA and a sweep CTA both own their
        // SM's shared memory; emission CTAs are persistent over the launch, so sweep CTAs launched behind them would wait).
        const bool tpc_parts = c->struct_state == 1 && c->opt_sweep != 1;
        // segmented sweeps (viterbi_seam.h) for every group when cutting pays for the batch as a whole: no chain is a critical
        // path any more, a group's sweep is ~0.67 ms x its share of the bins x 148 / the SMs it finds free
        const bool seg_parts = use_segments(c, ns);
        const int reserve = c->opt_reserve > 0 ? std::min(c->opt_reserve, g.n_sms / 2) : !tpc_parts ? 0 : std::min(g.n_sms / 3, u16 ? 44 : 33);
        for (size_t p = 0; p < plan.size(); p++) {
            const edb::BinRanges& rg = plan[p].ranges;
            edb::prof_mark("h2d_counts", sc);
            for (int q = 0; q < rg.n; q++) {
                const int64_t r0 = rg.b0[q], w = rg.b1[q] - r0;
                if (u16)
                    CU(cudaMemcpy2DAsync((uint16_t*)c->h_obs16.p + r0, nb * 2, b->observed16 + r0, b->obs16_stride * 2, w * 2, ns, cudaMemcpyHostToDevice, sc));
                else
                    CU(cudaMemcpy2DAsync((int32_t*)c->h_obs.p + r0, nb * 4, b->observed + r0, b->obs_stride * 4, w * 4, ns, cudaMemcpyHostToDevice, sc));
                if (!shared_ref)
                    CU(cudaMemcpy2DAsync((int32_t*)c->h_ref.p + r0, nb * 4, b->reference + r0, b->ref_stride * 4, w * 4, ns, cudaMemcpyHostToDevice, sc));
            }
            edb::prof_mark(nullptr, sc);
            if (u16)
                g_launches += edb::launch_widen_counts((const uint16_t*)c->h_obs16.p, nb, (int32_t*)c->h_obs.p, nb, ns, nb, rg,
                                                       (const int64_t*)c->h_ovf_i.p, (const int32_t*)c->h_ovf_v.p, b->n_overflow, sc);
            CU(cudaEventRecord(g.ev_copy[p], sc));
            CU(cudaStreamWaitEvent(g.s_em, g.ev_copy[p], 0));
            if ((rc = emission_part(c, &d, rg, false, emission_mode, p == 0 ? 1 : 2, g.s_em, p == 0 ? g.n_sms : g.n_sms - reserve))) return rc;
            CU(cudaEventRecord(g.ev_em[p], g.s_em));
            if (b->ll) {
                CU(cudaStreamWaitEvent(g.stream2, g.ev_em[p], 0));
                for (int q = 0; q < rg.n; q++) {
                    const int64_t r0 = rg.b0[q], w = rg.b1[q] - r0;
                    CU(cudaMemcpy2DAsync(b->ll + r0, b->ll_stride * 8, d.ll + r0, nbp * 8, w * 8, (size_t)ns * S, cudaMemcpyDeviceToHost, g.stream2));
                }
            }
            CU(cudaStreamWaitEvent(g.s_vit[p], g.ev_em[p], 0));
            // packing per part: "1" = one sweep warp per SM sub-partition (fast), "2" = two (half the SMs)
            // EDB200_OPT_PACKPLAN (experiments): decimal digits, one per part from the left, e.g. 222111
            int pack = p == 0 ? 1 : 2;
            if (c->opt_packplan > 0) {
                char pp[16];
                snprintf(pp, sizeof pp, "%d", c->opt_packplan);
                if (strlen(pp) > p) pack = pp[p] - '0';
            }
            // the first group (the longest chains) keeps the lane-per-state sweep, two warps per CTA: ~150 instead of ~205 cycles
            // per step of the chain everything else waits for; the other groups take the thread-per-chain sweep
            if (seg_parts) {
                if ((rc = viterbi_segmented(c, plan[p], va, g.s_vit[p]))) return rc;
            } else if ((rc = viterbi_part(c, plan[p], va, pack, g.s_vit[p], 0, tpc_parts && p > 0 ? 4 : 0))) return rc;
            if (b->path)
                for (int q = 0; q < rg.n; q++) {
                    const int64_t r0 = rg.b0[q], w = rg.b1[q] - r0;
                    CU(cudaMemcpy2DAsync(b->path + r0, b->path_stride, (int8_t*)c->h_path.p + r0, nb, w, ns, cudaMemcpyDeviceToHost, g.s_vit[p]));
                }
            CU(cudaEventRecord(g.ev_vit[p], g.s_vit[p]));
        }
        // cor(test, reference) needs the counts only: behind the last upload, beside the kernels of the last parts
        if ((rc = call_summary(c, &d, false, true, sc))) return rc;
        CU(cudaEventRecord(g.ev_setup, sc));
        CU(cudaStreamWaitEvent(st, g.ev_setup, 0));
        for (size_t p = 0; p < plan.size(); p++) CU(cudaStreamWaitEvent(st, g.ev_vit[p], 0));
        g_launches += edb::launch_viterbi_compact(va, st);
        if ((rc = check_kernel("viterbi_compact"))) return rc;
        if ((rc = call_summary(c, &d, true, false, st))) return rc;
    } else {
        if (shared_ref) CU(cudaMemcpyAsync(c->h_ref.p, b->reference, nb * 4, cudaMemcpyHostToDevice, st));
        else CU(cudaMemcpy2DAsync(c->h_ref.p, nb * 4, b->reference, b->ref_stride * 4, nb * 4, ns, cudaMemcpyHostToDevice, st));
        if (b->per_bin_stride) {
            CU(cudaMemcpy2DAsync(c->h_phi.p, nb * 8, b->phi, b->per_bin_stride * 8, nb * 8, ns, cudaMemcpyHostToDevice, st));
            CU(cudaMemcpy2DAsync(c->h_exp.p, nb * 8, b->expected, b->per_bin_stride * 8, nb * 8, ns, cudaMemcpyHostToDevice, st));
        } else {
            CU(cudaMemcpyAsync(c->h_phi.p, b->phi, ns * 8, cudaMemcpyHostToDevice, st));
            CU(cudaMemcpyAsync(c->h_exp.p, b->expected, ns * 8, cudaMemcpyHostToDevice, st));
        }
        // The likelihood matrix is 8*S bytes per bin and sample on the way back, against 4 on the way in: the call is
        // bound by the device-to-host copy.  Samples therefore go through in chunks: the counts of chunk i+1 upload and
        // its emission kernel runs on `st` while the likelihoods of chunk i drain on the second stream.
        const int n_chunks = (b->ll && ns >= 2 * kHostChunks) ? kHostChunks : 1;
        const int per = (ns + n_chunks - 1) / n_chunks;
        // the emission kernel is chosen once for the call, from the whole batch: a (smaller) last chunk must not fall
        // back to another kernel — a sample's likelihoods do not depend on its slot in the batch
        int chunk_mode = emission_mode;
        if (emission_mode == EDB200_EMISSION_AUTO && !use_table(c, emission_mode)) {
            edb::TableDims unused{};
            chunk_mode = panel_table(c, (int64_t)ns * S, emission_mode, &unused) ? EDB200_EMISSION_PANEL : EDB200_EMISSION_DIRECT;
        }
        for (int k = 0, s0 = 0; s0 < ns; k++, s0 += per) {
            const int cnt = ns - s0 < per ? ns - s0 : per;
            if (u16 || p12) {
                if (p12)
                    CU(cudaMemcpy2DAsync((uint8_t*)c->h_obs16.p + (size_t)s0 * r12, r12, b->observed12 + (size_t)s0 * b->obs12_stride, b->obs12_stride,
                                         row12, cnt, cudaMemcpyHostToDevice, st));
                else
                    CU(cudaMemcpy2DAsync((uint16_t*)c->h_obs16.p + (size_t)s0 * nb, nb * 2, b->observed16 + (size_t)s0 * b->obs16_stride, b->obs16_stride * 2,
                                         nb * 2, cnt, cudaMemcpyHostToDevice, st));
                if (k == 0 && b->n_overflow > 0) {
                    CU(cudaMemcpyAsync(c->h_ovf_i.p, b->overflow_index, (size_t)b->n_overflow * 8, cudaMemcpyHostToDevice, st));
                    CU(cudaMemcpyAsync(c->h_ovf_v.p, b->overflow_value, (size_t)b->n_overflow * 4, cudaMemcpyHostToDevice, st));
                }
                edb::BinRanges all{};
                all.n = 1;
                all.b0[0] = 0;
                all.b1[0] = nb;
                if (p12)
                    g_launches += edb::launch_unpack12_counts((const uint8_t*)c->h_obs16.p + (size_t)s0 * r12, r12, (int32_t*)c->h_obs.p + (size_t)s0 * nb, nb, cnt, nb, st);
                else
                    g_launches += edb::launch_widen_counts((const uint16_t*)c->h_obs16.p + (size_t)s0 * nb, nb, (int32_t*)c->h_obs.p + (size_t)s0 * nb, nb, cnt, nb, all,
                                                           nullptr, nullptr, 0, st);
                // the overflow entries of every sample, after every chunk: rows not yet uploaded are overwritten by their own widen
                // pass and patched again then (the entries are idempotent)
                g_launches += edb::launch_patch_overflow((int32_t*)c->h_obs.p, nb, nb, all, (const int64_t*)c->h_ovf_i.p, (const int32_t*)c->h_ovf_v.p, b->n_overflow, st);
            } else
                CU(cudaMemcpy2DAsync((int32_t*)c->h_obs.p + (size_t)s0 * nb, nb * 4, b->observed + (size_t)s0 * b->obs_stride, b->obs_stride * 4,
                                     nb * 4, cnt, cudaMemcpyHostToDevice, st));
            edb200_batch e = d;
            e.n_samples = cnt;
            e.observed = d.observed + (size_t)s0 * nb;
            if (!shared_ref) e.reference = d.reference + (size_t)s0 * nb;
            e.phi = d.phi + (size_t)s0 * (b->per_bin_stride ? nb : 1);
            e.expected = d.expected + (size_t)s0 * (b->per_bin_stride ? nb : 1);
            e.ll = d.ll + (size_t)s0 * S * nbp;
            if ((rc = edb200_cohort_run_device(c, &e, 1, chunk_mode, st))) return rc;
            if (b->ll) {
                CU(cudaEventRecord(g.chunk_done[k], st));
                CU(cudaStreamWaitEvent(g.stream2, g.chunk_done[k], 0));
                CU(cudaMemcpy2DAsync(b->ll + (size_t)s0 * S * b->ll_stride, b->ll_stride * 8, e.ll, nbp * 8, nb * 8, (size_t)cnt * S,
                                     cudaMemcpyDeviceToHost, g.stream2));
            }
        }
        if (want_vit && (rc = edb200_cohort_run_device(c, &d, 2, emission_mode, st))) return rc;
        if ((d.call_stats || d.cor) && (rc = edb200_cohort_run_device(c, &d, 4, emission_mode, st))) return rc;
        if (b->path) CU(cudaMemcpy2DAsync(b->path, b->path_stride, c->h_path.p, nb, nb, ns, cudaMemcpyDeviceToHost, st));
    }

    if (!tables_copied) {
        if (b->calls && b->call_cap > 0) CU(cudaMemcpyAsync(b->calls, c->h_calls.p, (size_t)ns * cap * 16, cudaMemcpyDeviceToHost, st));
        if (b->ncalls) CU(cudaMemcpyAsync(b->ncalls, c->h_ncalls.p, ns * 4, cudaMemcpyDeviceToHost, st));
        if (b->call_stats && b->call_cap > 0) CU(cudaMemcpyAsync(b->call_stats, c->h_stats.p, (size_t)ns * cap * 24, cudaMemcpyDeviceToHost, st));
    }
    if (b->cor) CU(cudaMemcpyAsync(b->cor, c->h_cor.p, ns * 8, cudaMemcpyDeviceToHost, st));
    int warn = 0;
    if ((rc = pull_flags(st, &warn))) return rc;
    CU(cudaStreamSynchronize(g.stream2));
    if (b->ncalls && b->call_cap > 0)
        for (int s = 0; s < ns; s++) if (b->ncalls[s] > b->call_cap) warn |= EDB200_WARN_CALLCAP;
    return warn;
}

// ------------------------------------------------------------------------------------------------ reference-set sweep
namespace {
struct RefsetScratch {
    DevBuf counts, bl, sel, z, partial, c, block;      // block: this rank's standardised rows, shared with the peers over IPC
    void* peer[edb::kMaxPeers] = {};
    int n_peers = 0, my_rank = 0;
} rs;
}  // namespace

}  // extern "C"
namespace {
void release_refset_scratch()
{
    DevBuf* all[] = {&rs.counts, &rs.bl, &rs.sel, &rs.z, &rs.partial, &rs.c, &rs.block};
    for (DevBuf* b : all) release(*b);
    release_fit_scratch();
}
}  // namespace
extern "C" {

int64_t edb200_refset_kpad(int64_t n_selected) { return (n_selected + 15) & ~(int64_t)15; }

int edb200_refset_standardize_device(const int32_t* counts, int64_t stride, int32_t n_samples, const double* bin_length,
                                     const int32_t* selected, int64_t n_selected, double* z, void* cuda_stream)
{
    if (int rc = need_ctx()) return rc;
    if (!counts || !selected || !z || n_samples < 0 || n_selected < 2) return fail(EDB200_ERR_ARG, "bad argument (at least 2 selected bins)");
    edb::launch_refset_standardize(counts, stride, n_samples, bin_length, selected, n_selected, edb200_refset_kpad(n_selected), z,
                                   (cudaStream_t)cuda_stream);
    g_launches++;
    return check_kernel("refset_standardize");
}

static int refset_gram_common(const double* za, int32_t m, const edb::PeerRows& zb, int32_t n, int64_t n_selected, double* cor_out,
                              cudaStream_t st)
{
    std::lock_guard<std::mutex> lk(g_mu);           // the K-slice scratch is shared
    const int64_t k_pad = edb200_refset_kpad(n_selected);
    const int slices = edb::refset_gram_slices(m, n, k_pad, g.n_sms);
    if (int rc = ensure(rs.partial, (size_t)slices * m * n * 8)) return rc;
    edb::launch_refset_gram(za, m, zb, n, k_pad, slices, (double*)rs.partial.p, cor_out, st);
    g_launches += 2;
    return check_kernel("refset_gram");
}

int edb200_refset_gram_device(const double* za, int32_t m, const double* zb, int32_t n, int64_t n_selected, double* cor_out,
                              void* cuda_stream)
{
    if (int rc = need_ctx()) return rc;
    if (!za || !zb || !cor_out || m < 0 || n < 0) return fail(EDB200_ERR_ARG, "bad argument");
    if (m == 0 || n == 0) return 0;
    edb::PeerRows rows{};
    rows.base[0] = zb;
    rows.rows_per_rank = n > 0 ? n : 1;
    return refset_gram_common(za, m, rows, n, n_selected, cor_out, (cudaStream_t)cuda_stream);
}

// ---- the sharded sweep without an all-gather: peers' standardised blocks are mapped with CUDA IPC -----------------
int edb200_refset_block_alloc(int32_t rows_per_rank, int64_t n_selected, void** z_dev, void* ipc_handle_out)
{
    if (int rc = need_ctx()) return rc;
    if (rows_per_rank < 1 || n_selected < 2 || !z_dev || !ipc_handle_out) return fail(EDB200_ERR_ARG, "bad argument");
    std::lock_guard<std::mutex> lk(g_mu);
    const size_t bytes = (size_t)rows_per_rank * edb200_refset_kpad(n_selected) * 8;
    // a fresh cudaMalloc allocation: IPC handles cover whole allocations, so the block must not share one with anything
    if (rs.block.p) CU(cudaFree(rs.block.p));
    rs.block.p = nullptr;
    rs.block.cap = 0;
    CU(cudaMalloc(&rs.block.p, bytes));
    rs.block.cap = bytes;
    CU(cudaMemset(rs.block.p, 0, bytes));
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, rs.block.p));
    static_assert(sizeof(h) == EDB200_IPC_HANDLE_BYTES, "CUDA IPC handle size");
    memcpy(ipc_handle_out, &h, sizeof h);
    *z_dev = rs.block.p;
    return 0;
}

int edb200_refset_peers_open(const void* handles, int32_t world, int32_t my_rank)
{
    if (int rc = need_ctx()) return rc;
    if (!handles || world < 1 || world > edb::kMaxPeers || my_rank < 0 || my_rank >= world) return fail(EDB200_ERR_ARG, "bad argument (at most %d ranks)", edb::kMaxPeers);
    if (!rs.block.p) return fail(EDB200_ERR_ARG, "edb200_refset_block_alloc has not been called");
    std::lock_guard<std::mutex> lk(g_mu);
    for (int r = 0; r < world; r++) {
        if (r == my_rank) { rs.peer[r] = rs.block.p; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char*)handles + (size_t)r * sizeof h, sizeof h);
        void* p = nullptr;
        CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        rs.peer[r] = p;
    }
    rs.n_peers = world;
    rs.my_rank = my_rank;
    return 0;
}

int edb200_refset_peers_close(void)
{
    std::lock_guard<std::mutex> lk(g_mu);
    cudaDeviceSynchronize();
    for (int r = 0; r < rs.n_peers; r++)
        if (r != rs.my_rank && rs.peer[r]) cudaIpcCloseMemHandle(rs.peer[r]);
    rs.n_peers = 0;
    return 0;
}

int edb200_refset_gram_peers_device(int32_t m, int32_t rows_per_rank, int32_t n_total, int64_t n_selected, double* cor_out, void* cuda_stream)
{
    if (int rc = need_ctx()) return rc;
    if (rs.n_peers < 1) return fail(EDB200_ERR_ARG, "edb200_refset_peers_open has not been called");
    if (!cor_out || m < 0 || m > rows_per_rank || n_total < 1 || (int64_t)rs.n_peers * rows_per_rank < n_total) return fail(EDB200_ERR_ARG, "bad argument");
    if (m == 0) return 0;
    edb::PeerRows rows{};
    for (int r = 0; r < rs.n_peers; r++) rows.base[r] = (const double*)rs.peer[r];
    rows.rows_per_rank = rows_per_rank;
    return refset_gram_common((const double*)rs.block.p, m, rows, n_total, n_selected, cor_out, (cudaStream_t)cuda_stream);
}

int edb200_refset_correlations(const int32_t* counts, int64_t stride, int32_t n_samples, const double* bin_length,
                               const int32_t* selected, int64_t n_selected, int32_t row0, int32_t n_rows, double* cor_out)
{
    if (int rc = need_ctx()) return rc;
    if (!counts || !selected || !cor_out || n_samples < 1 || n_selected < 2 || row0 < 0 || n_rows < 0 || row0 + n_rows > n_samples)
        return fail(EDB200_ERR_ARG, "bad argument");
    int64_t n_bins = 0;
    for (int64_t i = 0; i < n_selected; i++) {
        if (selected[i] < 0 || selected[i] >= stride) return fail(EDB200_ERR_ARG, "selected[%lld] = %d is outside the rows", (long long)i, selected[i]);
        if (selected[i] + 1 > n_bins) n_bins = selected[i] + 1;
    }
    cudaStream_t st = g.stream;
    const int64_t k_pad = edb200_refset_kpad(n_selected);
    {
        std::lock_guard<std::mutex> lk(g_mu);
        int rc = 0;
        if ((rc = ensure(rs.counts, (size_t)n_samples * n_bins * 4)) || (rc = ensure(rs.sel, n_selected * 4)) ||
            (rc = ensure(rs.bl, n_bins * 8)) || (rc = ensure(rs.z, (size_t)n_samples * k_pad * 8)) || (rc = ensure(rs.c, (size_t)n_rows * n_samples * 8 + 8)))
            return rc;
        CU(cudaMemcpy2DAsync(rs.counts.p, n_bins * 4, counts, stride * 4, n_bins * 4, n_samples, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(rs.sel.p, selected, n_selected * 4, cudaMemcpyHostToDevice, st));
        if (bin_length) CU(cudaMemcpyAsync(rs.bl.p, bin_length, n_bins * 8, cudaMemcpyHostToDevice, st));
    }
    if (int rc = edb200_refset_standardize_device((const int32_t*)rs.counts.p, n_bins, n_samples, bin_length ? (const double*)rs.bl.p : nullptr,
                                                  (const int32_t*)rs.sel.p, n_selected, (double*)rs.z.p, st))
        return rc;
    if (int rc = edb200_refset_gram_device((const double*)rs.z.p + (size_t)row0 * k_pad, n_rows, (const double*)rs.z.p, n_samples, n_selected,
                                           (double*)rs.c.p, st))
        return rc;
    CU(cudaMemcpyAsync(cor_out, rs.c.p, (size_t)n_rows * n_samples * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return 0;
}

// ------------------------------------------------------------------------------------------------ beta-binomial fit
}  // extern "C"
namespace {
struct FitScratch {
    DevBuf obs, ref, mu, phi, ll, info, overflow, power;
} fs;
constexpr int kFitOverflow = 4096;                 // bins per sample that may exceed the caps
std::mutex g_fit_mu;
constexpr int kFitK = 6144, kFitRN = 22528;       // exceedance-array caps: (6144 + 2 * 22528) * 4 B = 200 KB of shared memory
void release_fit_scratch()
{
    DevBuf* all[] = {&fs.obs, &fs.ref, &fs.mu, &fs.phi, &fs.ll, &fs.info, &fs.overflow, &fs.power};
    for (DevBuf* b : all) release(*b);
}
}  // namespace
extern "C" {

int edb200_betabin_fit_device(const int32_t* observed, int64_t obs_stride, const int32_t* reference, int64_t ref_stride,
                              int32_t n_samples, int64_t n_bins, double* mu, double* phi, double* loglik, int32_t* info,
                              void* cuda_stream)
{
    if (int rc = need_ctx()) return rc;
    if (!observed || !reference || !mu || !phi || !loglik || !info || n_samples < 0 || n_bins < 1) return fail(EDB200_ERR_ARG, "bad argument");
    edb::TableDims d{kFitK, kFitRN, kFitRN};
    if (edb::betabin_fit_smem_bytes(d) > g.smem_optin)
        return fail(EDB200_ERR_CUDA, "device offers %zu B of shared memory per CTA; the fit kernel needs %zu", g.smem_optin, edb::betabin_fit_smem_bytes(d));
    edb::CountsView cv{observed, obs_stride, reference, ref_stride, 0};
    {
        std::lock_guard<std::mutex> lk(g_fit_mu);       // grows the overflow scratch (kept until shutdown)
        if (int rc = ensure(fs.overflow, (size_t)n_samples * kFitOverflow * 8)) return rc;
    }
    edb::launch_betabin_fit(cv, n_samples, n_bins, d, 200, fs.overflow.p, kFitOverflow, mu, phi, loglik, info, (cudaStream_t)cuda_stream);
    g_launches++;
    return check_kernel("betabin_fit");
}

int edb200_power_betabinom(const int32_t* size, const double* phi, const double* p, const double* alt_p, int32_t n, double* expected_bf)
{
    if (int rc = need_ctx()) return rc;
    if (n < 0 || (n > 0 && (!size || !phi || !p || !alt_p || !expected_bf))) return fail(EDB200_ERR_ARG, "bad argument");
    if (n == 0) return 0;
    std::lock_guard<std::mutex> lk(g_mu);
    cudaStream_t st = g.stream;
    // one staging block: [size | phi | p | alt_p | out]
    const size_t pad = ((size_t)n * 4 + 7) & ~(size_t)7;
    if (int rc = ensure(fs.power, pad + (size_t)n * 8 * 4)) return rc;
    char* base = (char*)fs.power.p;
    double* d_phi = (double*)(base + pad);
    CU(cudaMemcpyAsync(base, size, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_phi, phi, (size_t)n * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_phi + n, p, (size_t)n * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_phi + 2 * (size_t)n, alt_p, (size_t)n * 8, cudaMemcpyHostToDevice, st));
    edb::launch_power_betabinom((const int32_t*)base, d_phi, d_phi + n, d_phi + 2 * (size_t)n, n, d_phi + 3 * (size_t)n, st);
    g_launches++;
    if (int rc = check_kernel("power_betabinom")) return rc;
    CU(cudaMemcpyAsync(expected_bf, d_phi + 3 * (size_t)n, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return 0;
}

int edb200_betabin_fit(const int32_t* observed, int64_t obs_stride, const int32_t* reference, int64_t ref_stride, int32_t n_samples,
                       int64_t n_bins, double* mu, double* phi, double* loglik, int32_t* info)
{
    if (int rc = need_ctx()) return rc;
    if (!observed || !reference || !mu || !phi || n_samples < 1 || n_bins < 1 || obs_stride < n_bins || (ref_stride && ref_stride < n_bins))
        return fail(EDB200_ERR_ARG, "bad argument");
    std::lock_guard<std::mutex> lk(g_mu);
    cudaStream_t st = g.stream;
    const bool shared_ref = ref_stride == 0;
    int rc = 0;
    if ((rc = ensure(fs.obs, (size_t)n_samples * n_bins * 4)) || (rc = ensure(fs.ref, (size_t)(shared_ref ? 1 : n_samples) * n_bins * 4)) ||
        (rc = ensure(fs.mu, n_samples * 8)) || (rc = ensure(fs.phi, n_samples * 8)) || (rc = ensure(fs.ll, n_samples * 8)) ||
        (rc = ensure(fs.info, n_samples * 4)))
        return rc;
    CU(cudaMemcpy2DAsync(fs.obs.p, n_bins * 4, observed, obs_stride * 4, n_bins * 4, n_samples, cudaMemcpyHostToDevice, st));
    if (shared_ref) CU(cudaMemcpyAsync(fs.ref.p, reference, n_bins * 4, cudaMemcpyHostToDevice, st));
    else CU(cudaMemcpy2DAsync(fs.ref.p, n_bins * 4, reference, ref_stride * 4, n_bins * 4, n_samples, cudaMemcpyHostToDevice, st));
    if ((rc = edb200_betabin_fit_device((const int32_t*)fs.obs.p, n_bins, (const int32_t*)fs.ref.p, shared_ref ? 0 : n_bins, n_samples, n_bins,
                                        (double*)fs.mu.p, (double*)fs.phi.p, (double*)fs.ll.p, (int32_t*)fs.info.p, st)))
        return rc;
    std::vector<int32_t> inf(n_samples);
    CU(cudaMemcpyAsync(mu, fs.mu.p, n_samples * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(phi, fs.phi.p, n_samples * 8, cudaMemcpyDeviceToHost, st));
    if (loglik) CU(cudaMemcpyAsync(loglik, fs.ll.p, n_samples * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(inf.data(), fs.info.p, n_samples * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    int warn = 0;
    for (int s = 0; s < n_samples; s++) {
        if (info) info[s] = inf[s];
        if (inf[s] < 0) warn = EDB200_WARN_NAN;
    }
    return warn;
}

}  // extern "C"
