// common.cuh -- context, error handling and the device scratch pool of libpcdgpu.so.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/pcdgpu.h"
#include "ec.cuh"

#define PCD_CUDA(ctx, call)                                                                      \
  do {                                                                                           \
    cudaError_t e__ = (call);                                                                    \
    if (e__ != cudaSuccess) {                                                                    \
      (ctx)->set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));   \
      return PCDGPU_E_CUDA;                                                                      \
    }                                                                                            \
  } while (0)

#define PCD_TRY(expr)           \
  do {                          \
    int rc__ = (expr);          \
    if (rc__ != 0) return rc__; \
  } while (0)

// Per (field, log_n) NTT tables, built on first use and kept for the life of the context.
struct NttTables {
  void* twiddles = nullptr;   // omega^k, k < n/2
  void* coset_pow = nullptr;  // g^i, i < n
  void* coset_inv = nullptr;  // g^-i / n, i < n
};

// Profiling classes: CUDA-event spans recorded around groups of launches on the context's stream
// (pcdgpu_profile_enable / pcdgpu_profile_read); bench.py derives per-kernel time shares and the
// roofline figures from them.
enum {
  PROF_MSM_SORT = 0,  // digits + histogram, scan, scatter
  PROF_MSM_ACC_G1,    // the accumulate kernel of a large MSM (units: the entries IT walks -- heavy buckets excluded), G1 curves
  PROF_MSM_ACC_G2,    // same, G2 over Fq2 (MNT4)
  PROF_MSM_REDUCE,    // bucket reduction, per-window tree sum
  PROF_MSM_TAIL,      // Horner over windows
  PROF_NTT,           // all passes of one transform
  PROF_SPMV,          // CSR mat-vec + pointwise QAP combine
  PROF_ASSEMBLE,      // proof assembly (scalar multiplications, normalisation)
  PROF_MSM_ACC_G2Q3,  // bucket accumulation over Fq3 (MNT6 G2), kept apart from Fq2 so that the work per entry is exact
  PROF_MSM_ACC_SMALL, // bucket accumulation of MSMs below 2^14 points (default-circuit proofs): latency-bound regime
  PROF_MSM_ACC_TAIL,  // after the accumulate kernel of a large MSM: part fold and the heavy-bucket kernels (latency-bound
                      // trees over the few buckets that hold repeated witness values); units = the entries they add
  PROF_NSLOT
};

struct ProfSpan {
  cudaEvent_t a, b;
  int slot;
  double units;      // algorithmic units processed inside the span (entries, butterflies, rows)
  int units_pinned;  // >= 0: index into ctx->prof_pinned holding the unit count read back from the device
};

static const int PCD_E_RETRY = -100;  // internal: abandon the graph capture and redo the call eagerly (never returned)

struct pcdgpu_ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  int msm_window = 0;
  bool in_proof = false;     // inside a prover entry point: its MSMs run side by side on the lanes
  std::atomic<bool> busy{false};  // a prover call is running (InProofGuard): concurrent use from a second thread is refused
  bool key_upload = false;  // set while a proving key's queries are uploaded (window rule of concurrent MSMs)
  char err[512] = {0};
  // Lanes: the five MSMs of a proof are independent, so each runs on its own stream with its own
  // scratch and their latency-bound phases (bucket reduction, window combination) overlap the
  // other lanes' accumulation.  Lane 0 is the context's (or the caller's) stream.
  static const int NLANE = 7;        // lanes 5, 6: the two extra MSMs of a small Groth16 proof (see pcdgpu_groth16_prove_dev)
  static const int NLANE_PROOF = 5;  // lanes every prover forks
  static const int SLOTS_PER_LANE = 24;
  int lane = 0;
  bool concurrent = true;
  cudaStream_t lane_stream[NLANE] = {nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[NLANE] = {nullptr};
  // Order of the accumulation grids inside a proof (set by the prover before an MSM, consumed by it): the next MSM's
  // accumulate kernel waits for gate_wait[*] and records gate_done right after its launch.  ev_acc: a, b_g1, b_g2.
  cudaEvent_t ev_acc[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t gate_wait[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t gate_done = nullptr;
  // ev_sorted[k]: recorded after the scatter kernel (end of the sorting phase) of the MSM that was given sort_done
  cudaEvent_t ev_sorted[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t sort_done = nullptr;
  cudaStream_t cur() const { return lane == 0 ? stream : lane_stream[lane]; }
  // grow-only scratch slots (slot ids are fixed per use and lane so concurrent phases never alias)
  static const int NSLOT = SLOTS_PER_LANE * NLANE;
  void* slot[NSLOT] = {nullptr};
  size_t slot_bytes[NSLOT] = {0};
  std::map<int, NttTables> ntt_tables;  // key = field * 64 + log_n
  void* pinned = nullptr;               // small pinned staging buffer
  size_t pinned_bytes = 0;
  // profiling / counters
  bool profiling = false;
  std::vector<ProfSpan> spans;
  size_t spans_used = 0;
  unsigned* prof_pinned = nullptr;  // 4096 u32, pinned
  unsigned long long launches = 0;  // kernels launched by this context since the last read
  // CUDA graphs of whole proofs (pcdgpu_set_proof_graphs, off by default -- measured slower, see include/pcdgpu.h): a
  // proof's ~150 - 250 launches on up to seven streams depend only on the key, the constraint system and the addresses
  // involved, not on the assignment, so the second call with the same (key, system, assignment address) is captured
  // and every later one is ONE cudaGraphLaunch.  A graph bakes scratch addresses in: scratch_epoch counts
  // (re)allocations, an entry of another epoch is dropped; while capturing, scratch() refuses to grow (PCD_E_RETRY: the
  // call is redone eagerly).
  struct ProofGraph {
    unsigned long long pk_uid = 0, r1cs_uid = 0, epoch = 0, last_use = 0, launches = 0;
    const void* d_z = nullptr;
    int seen = 0;
    cudaGraphExec_t exec = nullptr;
  };
  static const int NGRAPH = 8;
  ProofGraph graphs[NGRAPH];
  unsigned long long scratch_epoch = 0, graph_clock = 0, graphs_captured = 0, graphs_replayed = 0;
  bool capturing = false, use_graphs = false;  // pcdgpu_set_proof_graphs: measured slower than eager enqueueing
  // multi-GPU (comm.cu): NCCL communicator of the ranks that share one proof / MSM; world 1 = none
  void* nccl_comm = nullptr;
  int comm_rank = 0, comm_world = 1;

  int prof_begin(int slot, double units) {
    if (!profiling) return -1;
    if (spans_used == spans.size()) {
      if (spans.size() >= 4096) return -1;
      ProfSpan sp;
      cudaEventCreate(&sp.a);
      cudaEventCreate(&sp.b);
      spans.push_back(sp);
    }
    ProfSpan& sp = spans[spans_used];
    sp.slot = slot;
    sp.units = units;
    sp.units_pinned = -1;
    cudaEventRecord(sp.a, cur());
    return (int)spans_used++;
  }
  void prof_end(int id) {
    if (id >= 0) cudaEventRecord(spans[id].b, cur());
  }

  // after a failure inside a forked region: let every lane drain before the caller sees the error, so that the
  // next call cannot reuse scratch a lane is still reading
  void drain_lanes() {
    for (int l = 1; l < NLANE; l++)
      if (lane_stream[l]) cudaStreamSynchronize(lane_stream[l]);
    lane = 0;
  }

  void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err, sizeof(err), fmt, ap);
    va_end(ap);
  }
  // scratch: returns a device buffer of at least `bytes` for slot `id`
  int scratch(int id, size_t bytes, void** out) {
    id += lane * SLOTS_PER_LANE;
    if (bytes > slot_bytes[id]) {
      if (capturing) {
        set_error("scratch slot %d would grow during graph capture", id);
        return PCD_E_RETRY;
      }
      scratch_epoch++;
      if (slot[id]) {
        cudaStreamSynchronize(cur());
        cudaFree(slot[id]);
        slot[id] = nullptr;
        slot_bytes[id] = 0;
      }
      size_t want = bytes + bytes / 8 + 256;
      cudaError_t e = cudaMalloc(&slot[id], want);
      if (e != cudaSuccess) {
        set_error("cudaMalloc(%zu) for scratch slot %d: %s", want, id, cudaGetErrorString(e));
        return PCDGPU_E_NOMEM;
      }
      slot_bytes[id] = want;
    }
    *out = slot[id];
    return 0;
  }
};

// Marks the context as "inside a prover call" (MSM launch heuristics) and enforces the header's rule that a context is
// used by one host thread at a time: a second prover call that finds the context busy fails (`ok` false) instead of
// sharing lanes, scratch and the pinned staging buffer with the first.
struct InProofGuard {
  pcdgpu_ctx* c;
  bool ok;
  explicit InProofGuard(pcdgpu_ctx* ctx) : c(ctx), ok(!ctx->busy.exchange(true)) {
    if (ok) c->in_proof = true;
  }
  ~InProofGuard() {
    if (ok) {
      c->in_proof = false;
      c->busy.store(false);
    }
  }
};

enum {
  SLOT_NTT = 0,      // ping-pong buffer of a multi-pass NTT
  SLOT_IO = 1,       // host-pointer entry points: staged input / output
  SLOT_IO2 = 2,
  SLOT_MSM_DIG = 3,  // MSM digits
  SLOT_MSM_ENT = 4,  // MSM sorted entries
  SLOT_MSM_CNT = 5,  // MSM bucket counts / offsets / cursors
  SLOT_MSM_BKT = 6,  // MSM bucket sums
  SLOT_MSM_SEG = 7,  // MSM reduction partials + result
  SLOT_WM_A = 8,     // witness map vectors
  SLOT_WM_B = 9,
  SLOT_WM_C = 10,
  SLOT_Z = 11,       // assignment on device
  SLOT_MISC = 12,
  SLOT_CUB = 13,
  SLOT_MSM_HP = 14,  // heavy-bucket partial sums
  SLOT_CUB2 = 15,
  SLOT_NTT_MIXED = 16,  // ping-pong buffer of a mixed-radix transform
  SLOT_SAP_FULL = 17,   // GM17: the SAP assignment (z and the extra variables); small Groth16 proofs: s z | r z
  SLOT_COMM = 18        // multi-GPU: this rank's partial sums and the gathered ones
};

template <class T>
static inline T* dptr(void* p) { return reinterpret_cast<T*>(p); }

static inline int ilog2_ceil(size_t n) {
  int l = 0;
  while (((size_t)1 << l) < n) l++;
  return l;
}
