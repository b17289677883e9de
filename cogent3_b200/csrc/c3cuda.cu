// libc3cuda: host side of the engine + the C ABI declared in include/c3cuda.h.
//
// One c3_engine holds, resident in HBM for the life of a likelihood function:
//   * the pattern tables of every node (int32 gather indexes, tip indicator rows),
//   * the partial likelihoods of every internal node, FP64, [row][bin][MP],
//   * every P[edge][bin] and the pre-multiplied tip tables,
//   * the eigen-systems / rate matrices ("q slots") the P's are derived from.
// An evaluation uploads ONE small parameter block (dirty expm jobs, per-level dirty
// node lists, pi, bin probs), then launches: batched expm -> one pruning kernel per
// dirty tree level -> the site-weighted reduction, and reads one double back.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/c3cuda.h"
#include "c3_types.cuh"
#include "compress_kernels.cuh"
#include "expm_kernels.cuh"
#include "prune_kernels.cuh"
#include "hmm_kernels.cuh"

using namespace c3;

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

#define C3_CUDA(call)                                                                      \
  do {                                                                                     \
    cudaError_t err__ = (call);                                                            \
    if (err__ != cudaSuccess) {                                                            \
      const int code__ = (err__ == cudaErrorMemoryAllocation) ? C3_ERR_MEMORY : C3_ERR_CUDA; \
      return fail(code__, std::string(#call) + ": " + cudaGetErrorString(err__));          \
    }                                                                                      \
  } while (0)

#define C3_REQUIRE(cond, msg)                              \
  do {                                                     \
    if (!(cond)) return fail(C3_ERR_INVALID, (msg));       \
  } while (0)

inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

struct Slot {
  int mode = -1;  // ExpmMode or -1 (unset)
  std::vector<double> data;
  bool dirty = false;
};

}  // namespace

struct c3_engine {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int n_sm = 148;

  int n_nodes = 0, M = 0, MP = 0, K = 0, root = 0, n_levels = 0;
  bool use_dmma = false;
  bool use_strip = true;  // pipelined 4-state kernel (C3_NO_STRIP=1 selects the first version)
  std::vector<std::vector<int>> children;
  std::vector<int> parent, level, child_begin;
  std::vector<std::vector<int>> level_nodes;  // internal nodes per level

  // host copies of the pattern tables until finalize
  std::vector<int64_t> n_rows;
  std::vector<std::vector<double>> tip_table;
  std::vector<std::vector<int32_t>> idx32;
  std::vector<double> counts;
  bool finalized = false;

  // tables built on the device by c3_compress_alignment (moved into the pool by c3_finalize)
  int64_t n_columns = -1;
  std::vector<int32_t*> d_index;    // per node: pattern number of every column, int32 [L] (root always kept)
  std::vector<int32_t*> d_idx_tmp;  // per internal node: gather table int32 [C][rows]
  int32_t* d_root_index = nullptr;  // == d_index[root]; survives finalize

  // device memory
  char* d_pool = nullptr;  // everything sized at finalize
  int64_t pool_bytes = 0, partial_bytes = 0;
  std::vector<double*> d_rows;      // per node: partials (internal) / tip_inner (tips)
  std::vector<double*> d_tip_table;
  std::vector<int32_t*> d_idx;
  double* d_P = nullptr;
  NodeDesc* d_nodes = nullptr;
  ChildRef* d_childs = nullptr;
  EdgeDesc* d_edges = nullptr;
  double* d_counts = nullptr;
  double* d_lh = nullptr;
  double* d_tile_sums = nullptr;
  unsigned int* d_counter = nullptr;
  double* d_lnl = nullptr;
  std::vector<NodeDesc> h_nodes;
  bool nodes_dirty = false;

  double* d_slots = nullptr;
  int64_t slot_stride = 0;
  int n_slots_alloc = 0;
  std::vector<Slot> slots;

  // per-evaluation parameter block
  char* h_block = nullptr;  // pinned
  char* d_block = nullptr;
  int64_t block_bytes = 0;
  int64_t off_pi = 0, off_bprobs = 0, off_jobs = 0, off_work = 0, off_prefix = 0, off_small = 0;
  bool use_small = true;  // fused small-levels kernel (C3_NO_SMALL=1 disables)
  bool use_small_reduce = true;  // ... which also reduces when the root is in it (C3_NO_SMALL_REDUCE=1 disables)
  // "medium" levels (more units than one cluster takes, up to grid_units_per_thread units per thread of the whole
  // GPU) join the small-levels group, which then runs as ONE cooperative launch over every SM with grid barriers
  // between levels instead of a launch per level (C3_NO_GRID_SMALL=1 disables, C3_GRID_UNITS=n)
  bool use_grid_small = false;  // opt-in (C3_GRID_SMALL=1): measured no faster than the launches it replaces, DESIGN.md lesson 17
  int grid_units_per_thread = 4;
  unsigned int* d_grid_bar = nullptr;
  char* d_hmm_scratch = nullptr;  // c3_hmm_reduce: per-block products + the result
  // recomputing kernel (prune4_rec_kernel): a run of consecutive small / medium levels (at most rec_units units
  // each, at least one of them too big for the small-levels cluster) becomes launches of up to REC_DEPTH + 1 levels
  // without any barrier.  C3_REC=1 enables, C3_REC_UNITS=n.
  // independent root subtrees on their own streams: the sweep is launched per (subtree of a root child, level)
  // instead of per level, every subtree on a stream of its own, so that one subtree's launch ramps up on the SMs
  // another one's launch is draining from (the ~14 us of ramp + drain per streaming launch, lesson 16).
  // Opt-in (C3_STREAMS=1): bit-identical, measured slower (C2 0.394 vs 0.364 ms: 22 smaller launches instead of 10,
  // each with its own ramp and drain; DESIGN.md lesson 16e).
  bool use_streams = false;
  std::vector<cudaStream_t> aux_streams;
  std::vector<cudaEvent_t> aux_done;
  cudaEvent_t aux_fork = nullptr;
  std::vector<int> subtree_of;  // [n_nodes] index of the root child above the node (the root: -1)
  bool use_rec = false;  // opt-in (C3_REC=1): bit-identical, no faster (DESIGN.md lesson 16d)
  int64_t rec_units = 400000;
  int64_t off_rec = 0;
  int path_variant = 1;   // 0: 1024 threads x 3 units, 1: 512 x 6, 2: 256 x 11 (C3_PATH_VARIANT)
  bool use_path = false;  // opt-in (C3_PATH=1): single-branch updates of small problems by ONE CTA that walks the dirty
                          // path with the rows in shared memory -- measured slower than the small-levels launch
                          // (one SM has to issue every instruction of a level), DESIGN.md lesson 18
  // chain kernel (prune4_chain_kernel): single-branch updates of small 4-state problems walked per ROOT row through
  // composed gather tables, no barrier between the levels (C3_CHAIN=0 disables)
  bool use_chain = true;
  bool skip_block_copy = false;  // (inside a batch capture: the batch's copy kernel brings every engine's block)
  // tree-walk kernel (prune4_walk_kernel): every other evaluation of such a problem -- full sweeps included -- the
  // same way, with a per-thread value stack for the dirty children
  // OPT-IN (C3_WALK=1): bit-identical; brca1's full sweep walks in 31-34 us against the small-levels launch's 40, but
  // every root row recomputes every node -- 64 engines at once run at 100k against 153k evaluations/s, a 64-taxon
  // problem at 104 against 74 us (DESIGN lesson 17h, ncu of the first version: profiles/r2_prune4_walk_ncu.txt)
  bool use_walk = false;
  int64_t off_ops = 0;
  // C3_CHAIN_FUSED=1: a single-branch update in the real eigen form as ONE launch -- parameters by value, the
  // branch's P computed inside the chain launch, no parameter-block copy, no expm launch, no graph.  Bit-identical;
  // opt-in because it measured no faster: what is left of an update is the launch / completion round trip between
  // host and GPU (~20 us on the bench box whichever way the work is submitted), not device time (DESIGN lesson 17g)
  bool use_chain_fused = false;
  bool chain_ok = false;              // the composed tables exist (small problem, <= 3 children everywhere)
  int32_t* d_chain_maps = nullptr;    // [n_nodes][root_rows]: root row -> row of the node
  const int32_t** d_node_maps = nullptr;   // [n_nodes]
  const int32_t** d_child_maps = nullptr;  // [child refs]
  unsigned int* d_chain_ctr = nullptr;     // [1 + reduction tiles] completion counters of the chain launch
  bool use_pdl = true;    // programmatic dependent launch between levels (C3_NO_PDL=1 disables)
  // M != 4: every node stores its rows already multiplied by the P of the branch above it, one
  // contraction per NODE row (prune_dmma_edge_kernel); C3_EDGE_MAJOR=0: the parent-major kernels
  bool edge_major = false;
  bool use_dmma_reg = true;  // warp-private register pipeline for M != 4 (C3_DMMA_REG=0: first version)
  bool use_reg2 = true;   // binary nodes: two strips in flight per warp (C3_REG1=1: the one-strip regpipe kernel)
  // C3_REG2_U2=1: binary nodes by 16 warps x 2 rows per lane (strips of 64 units) instead of 8 x 4 (128): the same
  // bytes in flight per SM in steps half as long -- the drain and the last-strip imbalance of a launch scale with
  // the step
  bool reg2_u2 = false;
  bool use_dyn = false;   // C3_DYN=1: strips past a warp's first five are claimed from a per-launch counter
  int64_t off_ctr = 0;    // [DYN_MAX_LAUNCHES] int32 zeros in the parameter block
  int root_u = 3;  // rows per lane of the three-children instance: 3 (8 warps, 255 registers) or 2 (12 warps, C3_ROOT_U=2)
  // lean root (C3_LEAN_ROOT=0 disables): the regpipe2 root launch writes mix[p] (8 B per row) instead of lh[p][K]
  // and the reduction reads mix + counts only; lh is recomputed by the root launch alone when an accessor asks for
  // it (ensure_lh), and stays materialised for the next LH_KEEP evaluations after such a request
  bool lean_root = true;
  bool lh_valid = true;   // d_lh holds the last evaluation's per-pattern likelihoods
  bool last_lean = false; // the evaluation being planned / last issued has a lean root launch
  bool last_chain = false; // ... is one chain launch
  bool last_walk = false;  // ... is one tree-walk launch
  int lh_keep = 0;        // evaluations that still write lh because an accessor read it recently
  double* d_mix = nullptr;
  struct RootLaunch { int work_off = 0, prefix_off = 0, n_work = 0, total_tiles = 0, kind = 0; } root_launch;
  // developer timeline (C3_TIMELINE=1): globaltimer stamps of every pruning / reduction launch of the LAST
  // evaluation, printed by c3_destroy
  bool timeline = false;
  bool hostprof = false;   // C3_HOSTPROF=1: host nanoseconds of the blocking path (plan | issue or replay | wait), printed by c3_destroy
  long long hp_plan_ns = 0, hp_issue_ns = 0, hp_wait_ns = 0, hp_n = 0;
  unsigned long long* d_tl = nullptr;
  unsigned long long* h_tl_init = nullptr;
  int tl_launches = 0;
  std::vector<int> tl_kinds;
  bool use_ring = false;  // C3_RING=1: the cp.async shared-memory ring version of the strip kernel
  // Nothing on the path reads the ROOT's partial rows (only lh = inner(plh_root, pi) goes on, and the
  // root is recomputed whenever anything changed), so by default they are neither allocated nor written
  // -- a quarter of the root launch's traffic at 4 states; c3_get_partials(root) recomputes them.
  // C3_STORE_ROOT=1, or a kernel family without the check, keeps them.
  bool store_root = false;
  // dataflow kernel: consecutive big binary levels in one launch.  Opt-in (C3_FLOW=1): its kernel
  // time is lower than the per-level launches' sum (C2 0.267 vs 0.289 ms, C3 0.435 vs 0.585 ms) but
  // the per-level launches overlap through programmatic dependent launch, so the STEP is not faster
  bool use_flow = false;
  // wavefront kernel (prune4_wave_kernel): every dirty binary node of the sweep in one persistent launch, all
  // levels streaming at once; the static strip schedule of a structure is built on the host the second time the
  // structure comes up and kept on the device.
  bool use_wave = false;  // opt-in (C3_WAVE=1): measured slower than the per-level launches, DESIGN.md lesson 16
  int wave_lag = 4;          // chunks between a strip and the strips it depends on (C3_WAVE_LAG)
  int wave_group = 1;        // chunks per completion signal (C3_WAVE_GROUP)
  int64_t wave_prefix = 0;   // > 0: the wave launch takes only the tree's lower levels, up to the first level with
                             // more strips than this; the big levels keep their own launches (C3_WAVE_PREFIX)
  int wave_prio = 128;       // strips a parent runs behind its children in the schedule's priority order (C3_WAVE_PRIO)
  int64_t wave_min_strips = 4096;  // smaller sweeps keep the per-level / small-levels launches (C3_WAVE_MIN)
  std::vector<char> idx_mono;      // [n_nodes] idx_c[p] <= p for every child and row (first-seen numbering)
  struct WaveSched {
    std::vector<int32_t> work;  // the structure: dirty non-root nodes, ascending
    int2* d_sched = nullptr;
    unsigned* d_done = nullptr;
    int n_chunks = 0, grid = 0;
    int64_t bytes = 0, strips = 0, noops = 0;
    uint64_t last_use = 0;
    double build_ms = 0;
  };
  std::unordered_map<uint64_t, WaveSched> waves;
  std::unordered_map<uint64_t, int> wave_seen;
  int64_t wave_bytes = 0;
  unsigned int* d_done = nullptr;  // per work item completion counters of the dataflow launches
  int64_t off_dep = 0;
  cudaEvent_t block_copied = nullptr;
  bool block_free = false;  // the last evaluation's result has been collected: the staging buffer is free without asking
  // mapped pinned result: the value and, after it, the sequence number of the evaluation that wrote it
  // (the blocking path spins on that word instead of synchronising the stream)
  double* h_lnl = nullptr;
  double* h_lnl_dev = nullptr;
  volatile unsigned long long* h_seq = nullptr;
  unsigned long long* h_seq_dev = nullptr;
  unsigned long long* d_seq = nullptr;    // device-side count of published results
  unsigned long long* d_epoch = nullptr;  // device-side evaluation number of the fused all-reduce
  unsigned long long seq_issued = 0;      // results the issued evaluations will have published
  bool spin_wait = true;                  // C3_SPIN=0: cudaStreamSynchronize
  bool auto_futile = false;  // rescaling was tried for a -inf and did not help (a pattern with likelihood 0)

  // model state (host)
  std::vector<int32_t> qslot;       // [n_nodes*K]
  std::vector<double> dist;         // [n_nodes*K]
  std::vector<char> pdirty;         // [n_nodes*K] P must be recomputed
  std::vector<char> pfixed;         // [n_nodes*K] P supplied by host
  std::vector<char> node_dirty;     // [n_nodes]
  std::vector<double> pi, bprobs;
  std::vector<int32_t> fixed_motif;
  bool reduce_dirty = true;
  bool have_pi = false, have_dist = false;
  double cached_lnl = NAN;
  bool have_cached = false;
  bool deferred_pending = false;  // c3_evaluate_many: launched, result not collected yet
  bool async_pending = false;     // c3_evaluate_device work may still be running on the stream

  // sharded runs: fused all-reduce through peer mailboxes (c3_comm_*)
  CommArgs comm{};            // nranks == 0: not attached
  CommSlot* d_mailbox = nullptr;
  int mailbox_ranks = 0;
  std::vector<void*> ipc_opened;

  // CUDA graphs: the stream operations of an evaluation (parameter-block copy, expm, one launch per
  // dirty level, reduction) depend only on its STRUCTURE -- which nodes are dirty, how many expm jobs --
  // while the numbers travel in the parameter block.  An optimiser revisits the same structures over
  // and over (Powell: every parameter, every iteration; a full sweep for every rate-matrix
  // parameter; bootstraps and batches of small alignments: always the full sweep), so a structure
  // that has come up `graph_after` times is captured once and from then on replayed with ONE
  // cudaGraphLaunch instead of 4-15 launch calls.  C3_GRAPH=0 disables.
  bool use_graphs = true;
  struct GraphEntry { std::vector<int32_t> sig; cudaGraphExec_t exec = nullptr; int64_t n_launch = 0; uint64_t last_use = 0; };
  std::unordered_map<uint64_t, GraphEntry> graphs;
  // c3_evaluate_many batches: 0 = normal; 1 = issue into the caller's capture (no graph of its own);
  // 2 = dry plan: fill the parameter block, report the structure hash, launch nothing, mark nothing
  int batch_mode = 0;
  uint64_t last_sig_hash = 0;
  bool last_skipped = false;  // nothing was dirty: the evaluation returned the cached value
  struct PlanStats { int64_t expm_jobs = 0, nodes = 0, node_rows = 0, alg_bytes = 0, flops = 0, h2d_bytes = 0; };
  PlanStats plan;
  std::unordered_map<uint64_t, int> graph_seen;  // structure hash -> times seen before its capture
  int graph_after = 16;                          // sightings before a structure is captured (C3_GRAPH_AFTER)
  uint64_t graph_clock = 0;

  // per-pattern underflow rescaling (c3_set_rescaling; kernels in prune_kernels.cuh)
  int rescale_mode = C3_RESCALE_AUTO;
  int rescale_threshold = 16;  // a subtree becomes a scaled node once it holds this many unscaled tips
  bool rescale_active = false;
  char* d_scale_pool = nullptr;
  int64_t scale_pool_bytes = 0;
  std::vector<char> is_scaled;  // [n_nodes]
  struct ScaleRec { int32_t* d_exps = nullptr; const ScaleSrc* d_src = nullptr; int n_src = 0; };
  std::vector<ScaleRec> scale_rec;  // [n_nodes]; the root's d_exps are the per-pattern exponents of lh
  int n_scaled = 0;

  c3_stats stats{};

  // per-launch profiling
  bool profiling = false;
  struct LaunchRec { int kind; int64_t bytes, flops; cudaEvent_t a, b; float ms; };
  std::vector<LaunchRec> prof;
  std::vector<cudaEvent_t> event_pool;
  size_t events_used = 0;
};

namespace {

// cudaFuncSetAttribute is per device (context): one engine per GPU in one process (c3_create(device, ...),
// c3_comm_attach(device_ptrs=...)) must opt in on EVERY device it runs on, not only the first
// Engines are single-threaded objects (like cogent3's Calculator), but SEVERAL engines may be driven from
// several host threads at once (cogent3_b200.batch: ctypes releases the GIL during a call): what engines share
// -- the per-device opt-in flags, the batch-graph registry -- is guarded.
std::mutex g_once_mutex;
std::recursive_mutex g_batch_mutex;

struct DeviceOnce {
  uint64_t done[4] = {0, 0, 0, 0};
  bool need(int device) {
    std::lock_guard<std::mutex> lock(g_once_mutex);
    const int d = device & 255;
    if (done[d >> 6] & (1ull << (d & 63))) return false;
    done[d >> 6] |= 1ull << (d & 63);
    return true;
  }
};

template <typename Kern>
int set_max_smem(Kern kern, size_t bytes) {
  if (bytes > 48 * 1024) {
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (err != cudaSuccess) return fail(C3_ERR_CUDA, std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(err));
  }
  return C3_OK;
}

template <int MP>
int launch_dmma(c3_engine* e, const PruneArgs& a, bool is_root, int grid) {
  const size_t smem = DmmaCfg<MP>::SMEM_BYTES;
  static DeviceOnce configured;
  if (configured.need(e->device)) {
    int rc = set_max_smem(prune_dmma_kernel<MP, false>, smem);
    if (rc) return rc;
    rc = set_max_smem(prune_dmma_kernel<MP, true>, smem);
    if (rc) return rc;
  }
  if (is_root)
    prune_dmma_kernel<MP, true><<<grid, DMMA_THREADS, smem, e->stream>>>(a);
  else
    prune_dmma_kernel<MP, false><<<grid, DMMA_THREADS, smem, e->stream>>>(a);
  return C3_OK;
}

template <int MP>
int launch_dmma_reg(c3_engine* e, const PruneArgs& a, bool is_root) {
  const size_t smem = DRegCfg<MP>::SMEM_BYTES;
  static DeviceOnce configured;
  if (configured.need(e->device)) {
    int rc = set_max_smem(prune_dmma_reg_kernel<MP, false>, smem);
    if (rc) return rc;
    rc = set_max_smem(prune_dmma_reg_kernel<MP, true>, smem);
    if (rc) return rc;
  }
  const int grid = std::max(1, std::min(a.total_tiles, e->n_sm));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(DR_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = e->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = e->use_pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t err = is_root ? cudaLaunchKernelEx(&cfg, prune_dmma_reg_kernel<MP, true>, a)
                            : cudaLaunchKernelEx(&cfg, prune_dmma_reg_kernel<MP, false>, a);
  if (err != cudaSuccess) return fail(C3_ERR_CUDA, std::string("dmma-reg launch: ") + cudaGetErrorString(err));
  return C3_OK;
}

constexpr int DE_RA = 1;  // row groups per warp of the contraction instance (2: 8 warps x 16 rows, measured slower, DESIGN lesson 13)
constexpr bool DE_STG = false;  // the next tile's rows staged in shared memory (cp.async), 16 warps instead of 12: measured slower (DESIGN lesson 13)

template <int MP>
int launch_dmma_edge(c3_engine* e, const PruneArgs& a, bool root_only3) {
  using Cfg = DECfg<DE_RA, DE_STG>;
  const size_t smem_p = (size_t)MP * (MP + 8) * sizeof(double);
  const size_t smem = smem_p + (DE_STG ? (size_t)Cfg::WARPS * 2 * 8 * (MP + 8) * sizeof(double) : 0);
  static DeviceOnce configured;
  if (configured.need(e->device)) {
    int rc = set_max_smem(prune_dmma_edge_kernel<MP, 2, true, DE_RA, DE_STG>, smem);
    if (rc) return rc;
    rc = set_max_smem(prune_dmma_edge_kernel<MP, 3, false, 1>, smem_p);
    if (rc) return rc;
  }
  const int grid = std::max(1, std::min(a.total_tiles, e->n_sm));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(root_only3 ? DECfg<1>::THREADS : Cfg::THREADS);
  cfg.dynamicSmemBytes = root_only3 ? smem_p : smem;
  cfg.stream = e->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = e->use_pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // a launch holding only the trifurcating root: three prefetched children, no contraction
  cudaError_t err = root_only3 ? cudaLaunchKernelEx(&cfg, prune_dmma_edge_kernel<MP, 3, false, 1>, a)
                               : cudaLaunchKernelEx(&cfg, prune_dmma_edge_kernel<MP, 2, true, DE_RA, DE_STG>, a);
  if (err != cudaSuccess) return fail(C3_ERR_CUDA, std::string("dmma-edge launch: ") + cudaGetErrorString(err));
  return C3_OK;
}

int dmma_blocks_per_sm(int MP) {
  const size_t smem = (size_t)(DMMA_TILE_ROWS * (MP + 4) + 2 * MP * (MP + 4)) * sizeof(double);
  int per_sm = (int)((227 * 1024) / (smem + 1024));
  return std::max(1, std::min(per_sm, 2048 / DMMA_THREADS));
}

template <int K, int NC>
int launch_strip(c3_engine* e, const PruneArgs& a, bool is_root) {
  using Cfg = P4SCfg<K, NC>;
  static DeviceOnce configured;
  if (configured.need(e->device)) {
    int rc = set_max_smem(prune4_strip_kernel<K, NC, false>, Cfg::SMEM_BYTES);
    if (rc) return rc;
    rc = set_max_smem(prune4_strip_kernel<K, NC, true>, Cfg::SMEM_BYTES);
    if (rc) return rc;
  }
  // one CTA per SM (shared memory bound); never more warps than strips
  const int grid = std::max(1, std::min(e->n_sm, (a.total_tiles + Cfg::WARPS - 1) / Cfg::WARPS));
  // programmatic dependent launch: this kernel's descriptor staging and first index loads
  // overlap the tail of the previous launch (it calls griddepcontrol.wait before gathering)
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(Cfg::WARPS * 32);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = e->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = e->use_pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t err = is_root ? cudaLaunchKernelEx(&cfg, prune4_strip_kernel<K, NC, true>, a)
                            : cudaLaunchKernelEx(&cfg, prune4_strip_kernel<K, NC, false>, a);
  if (err != cudaSuccess) return fail(C3_ERR_CUDA, std::string("strip launch: ") + cudaGetErrorString(err));
  return C3_OK;
}

template <int K, int NC>
int launch_reg(c3_engine* e, const PruneArgs& a, bool is_root) {
  constexpr int WARPS = RegCfg<NC>::WARPS;
  const int grid = std::max(1, std::min(e->n_sm, (a.total_tiles + WARPS - 1) / WARPS));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(WARPS * 32);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = e->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = e->use_pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t err = is_root ? cudaLaunchKernelEx(&cfg, prune4_reg_kernel<K, NC, true>, a)
                            : cudaLaunchKernelEx(&cfg, prune4_reg_kernel<K, NC, false>, a);
  if (err != cudaSuccess) return fail(C3_ERR_CUDA, std::string("regpipe launch: ") + cudaGetErrorString(err));
  return C3_OK;
}

template <int K>
int launch_flow(c3_engine* e, const PruneArgs& a, const int32_t* d_dep, unsigned int* d_done) {
  static DeviceOnce configured;
  if (configured.need(e->device)) {
    int rc = set_max_smem(prune4_flow_kernel<K>, sizeof(FlowSmem));
    if (rc) return rc;
  }
  // every CTA must be resident (the kernel waits on other CTAs' progress): one CTA per SM
  const int grid = std::max(1, std::min(e->n_sm, (a.total_tiles + 7) / 8));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = sizeof(FlowSmem);
  cfg.stream = e->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = e->use_pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t err = cudaLaunchKernelEx(&cfg, prune4_flow_kernel<K>, a, d_dep, d_done);
  if (err != cudaSuccess) return fail(C3_ERR_CUDA, std::string("dataflow launch: ") + cudaGetErrorString(err));
  return C3_OK;
}

constexpr int DYN_MAX_LAUNCHES = 64;
template <int K, int NC, int U, int WARPS>
int launch_reg2_cfg(c3_engine* e, const PruneArgs& a, bool is_root);

template <int K, int NC>
int launch_reg2(c3_engine* e, const PruneArgs& a, bool is_root) {
  if (NC == 2) return e->reg2_u2 ? launch_reg2_cfg<K, 2, 2, 16>(e, a, is_root) : launch_reg2_cfg<K, 2, 4, 8>(e, a, is_root);
  return e->root_u == 3 ? launch_reg2_cfg<K, 3, 3, 8>(e, a, is_root) : launch_reg2_cfg<K, 3, 2, 12>(e, a, is_root);
}

template <int K, int NC, int U, int WARPS>
int launch_reg2_cfg(c3_engine* e, const PruneArgs& a, bool is_root) {
  const int grid0 = std::max(1, std::min(e->n_sm, (a.total_tiles + WARPS - 1) / WARPS));
  if (e->use_dyn && a.strip_ctr != nullptr && a.total_tiles > 8 * grid0 * WARPS) {
    // dynamic strips (C3_DYN=1): the same kernel, strips past a warp's first five claimed from a counter
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid0);
    cfg.blockDim = dim3(WARPS * 32);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = e->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = e->use_pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t err = is_root ? cudaLaunchKernelEx(&cfg, prune4_reg2_kernel<K, NC, U, WARPS, true, true>, a)
                              : cudaLaunchKernelEx(&cfg, prune4_reg2_kernel<K, NC, U, WARPS, false, true>, a);
    if (err != cudaSuccess) return fail(C3_ERR_CUDA, std::string("regpipe2 (dynamic) launch: ") + cudaGetErrorString(err));
    return C3_OK;
  }
  const int grid = std::max(1, std::min(e->n_sm, (a.total_tiles + WARPS - 1) / WARPS));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(WARPS * 32);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = e->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = e->use_pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t err = is_root ? cudaLaunchKernelEx(&cfg, prune4_reg2_kernel<K, NC, U, WARPS, true>, a)
                            : cudaLaunchKernelEx(&cfg, prune4_reg2_kernel<K, NC, U, WARPS, false>, a);
  if (err != cudaSuccess) return fail(C3_ERR_CUDA, std::string("regpipe2 launch: ") + cudaGetErrorString(err));
  return C3_OK;
}

// ---- wavefront kernel: the static strip schedule of one structure (set of dirty non-root nodes) -----------
// Strips are visited in priority order  w = k + prio * height(node)  (a parent `prio` strips behind its
// children) and placed greedily into chunks of GW entries: a strip goes into the first chunk with room that lies
// at least `lag` chunks after the chunks of the child strips it depends on (strip k of a node needs strips <= k of
// its children: idx_c[p] <= p).  Entries of a chunk are ordered by (node, strip), so a warp keeps meeting the same
// node chunk after chunk; positions nobody could fill are no-op entries.
void free_wave(c3_engine* e, c3_engine::WaveSched& ws) {
  if (ws.d_sched) cudaFree(ws.d_sched);
  if (ws.d_done) cudaFree(ws.d_done);
  e->wave_bytes -= ws.bytes;
  e->stats.device_bytes -= ws.bytes;
  ws.d_sched = nullptr;
  ws.d_done = nullptr;
  ws.bytes = 0;
}

int build_wave_schedule(c3_engine* e, const std::vector<int>& nodes, c3_engine::WaveSched& ws) {
  const auto t0 = std::chrono::steady_clock::now();
  const int K = e->K, ROWS = 128 / K, W = (int)nodes.size();
  const int grid = e->n_sm, GW = grid * WAVE_WARPS, lag = e->wave_lag, prio = e->wave_prio, R = e->wave_group;
  // the kernel gathers for chunk j once every GROUP that ends at or before chunk j - lag is complete: a strip
  // whose dependency sits in chunk d may go into chunk j >= ceil((d + 1) / R) R + lag - 1
  auto after = [&](int64_t d) { return (d + 1 + R - 1) / R * R + lag - 1; };
  std::vector<int> wi(e->n_nodes, -1);
  for (int i = 0; i < W; ++i) wi[nodes[i]] = i;
  std::vector<int64_t> strips(W), off(W + 1, 0), start(W);
  std::vector<int> h(W, 0), kid0(W), kid1(W);
  for (int i = 0; i < W; ++i) {  // ascending node ids: children before parents
    const int n = nodes[i];
    strips[i] = (e->n_rows[n] + ROWS - 1) / ROWS;
    off[i + 1] = off[i] + strips[i];
    kid0[i] = wi[e->children[n][0]];
    kid1[i] = wi[e->children[n][1]];
    if (kid0[i] >= 0) h[i] = std::max(h[i], h[kid0[i]] + 1);
    if (kid1[i] >= 0) h[i] = std::max(h[i], h[kid1[i]] + 1);
    start[i] = (int64_t)prio * h[i];
  }
  const int64_t total = off[W];
  if (total >= INT32_MAX) return fail(C3_ERR_INVALID, "too many strips for a wave schedule");
  std::vector<int32_t> chunk_of((size_t)total);
  std::vector<int32_t> fill;
  fill.reserve((size_t)(total / GW + 64));
  int64_t first_open = 0;
  // items by starting wave
  std::vector<int> order(W);
  for (int i = 0; i < W; ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return start[a] < start[b]; });
  std::vector<int> active;
  int next = 0;
  int64_t remaining = total;
  for (int64_t w = 0; remaining > 0; ++w) {
    bool grew = false;
    while (next < W && start[order[next]] <= w) {
      active.push_back(order[next++]);
      grew = true;
    }
    if (active.empty()) {
      w = start[order[next]] - 1;  // (nothing active: jump to the next start)
      continue;
    }
    if (grew) std::sort(active.begin(), active.end());
    size_t keep = 0;
    for (size_t ai = 0; ai < active.size(); ++ai) {
      const int i = active[ai];
      const int64_t k = w - start[i];
      if (k >= strips[i]) continue;  // finished: dropped from the active list
      int64_t earliest = 0;
      if (kid0[i] >= 0) earliest = std::max<int64_t>(earliest, after(chunk_of[off[kid0[i]] + std::min(k, strips[kid0[i]] - 1)]));
      if (kid1[i] >= 0) earliest = std::max<int64_t>(earliest, after(chunk_of[off[kid1[i]] + std::min(k, strips[kid1[i]] - 1)]));
      int64_t j = std::max(earliest, first_open);
      while (j < (int64_t)fill.size() && fill[j] == GW) ++j;
      if (j >= (int64_t)fill.size()) fill.resize((size_t)j + 1, 0);
      fill[j] += 1;
      chunk_of[off[i] + k] = (int32_t)j;
      while (first_open < (int64_t)fill.size() && fill[first_open] == GW) ++first_open;
      --remaining;
      if (k + 1 < strips[i]) active[keep++] = i;
    }
    active.resize(keep);
  }
  const int64_t n_chunks = (int64_t)fill.size();
  if (n_chunks * GW >= ((int64_t)1 << 31)) return fail(C3_ERR_INVALID, "wave schedule too long");
  std::vector<int2> sched((size_t)(n_chunks * GW), make_int2(-1, 0));
  std::vector<int32_t> cursor((size_t)n_chunks, 0);
  for (int i = 0; i < W; ++i)
    for (int64_t k = 0; k < strips[i]; ++k) {
      const int32_t c = chunk_of[off[i] + k];
      sched[(size_t)c * GW + cursor[c]++] = make_int2(i, (int)k);
    }
  if (getenv("C3_WAVE_VERIFY")) {
    // every strip exactly once; every dependency at least `lag` chunks earlier; no strip that depends on
    // anything inside the first `lag` chunks
    std::vector<char> seen((size_t)total, 0);
    for (int64_t c = 0; c < n_chunks; ++c)
      for (int g = 0; g < GW; ++g) {
        const int2 en = sched[(size_t)c * GW + g];
        if (en.x < 0) continue;
        if (en.x >= W || en.y < 0 || en.y >= strips[en.x] || seen[off[en.x] + en.y] || chunk_of[off[en.x] + en.y] != c)
          return fail(C3_ERR_INVALID, "wave schedule: bad entry");
        seen[off[en.x] + en.y] = 1;
        for (int kid : {kid0[en.x], kid1[en.x]}) {
          if (kid < 0) continue;
          const int64_t ck = std::min<int64_t>(en.y, strips[kid] - 1);
          if (after(chunk_of[off[kid] + ck]) > c) return fail(C3_ERR_INVALID, "wave schedule: dependency too close");
        }
      }
    for (int64_t t = 0; t < total; ++t)
      if (!seen[t]) return fail(C3_ERR_INVALID, "wave schedule: strip missing");
  }
  ws.work.assign(nodes.begin(), nodes.end());
  ws.n_chunks = (int)n_chunks;
  ws.grid = grid;
  ws.strips = total;
  ws.noops = n_chunks * GW - total;
  const int64_t bytes = (int64_t)sched.size() * sizeof(int2) + n_chunks * (int64_t)sizeof(unsigned);
  // keep the schedules of a few structures (a full sweep: 8 B per strip); least recently used go first
  while (!e->waves.empty() && e->wave_bytes + bytes > ((int64_t)1 << 30)) {
    auto victim = e->waves.begin();
    for (auto it = e->waves.begin(); it != e->waves.end(); ++it)
      if (it->second.last_use < victim->second.last_use) victim = it;
    cudaStreamSynchronize(e->stream);
    free_wave(e, victim->second);
    e->waves.erase(victim);
  }
  C3_CUDA(cudaMalloc(&ws.d_sched, sched.size() * sizeof(int2)));
  cudaError_t err = cudaMalloc(&ws.d_done, (size_t)std::max<int64_t>(1, n_chunks) * sizeof(unsigned));
  if (err != cudaSuccess) {
    cudaFree(ws.d_sched);
    ws.d_sched = nullptr;
    return fail(C3_ERR_MEMORY, "device memory exhausted (wave schedule)");
  }
  C3_CUDA(cudaMemcpy(ws.d_sched, sched.data(), sched.size() * sizeof(int2), cudaMemcpyHostToDevice));
  ws.bytes = bytes;
  e->wave_bytes += bytes;
  e->stats.device_bytes += bytes;
  ws.build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  if (getenv("C3_WAVE_VERBOSE"))
    fprintf(stderr, "[c3cuda] wave schedule: %d nodes, %lld strips, %lld chunks of %d (%.1f%% no-ops), lag %d, group %d, built in %.1f ms\n",
            W, (long long)total, (long long)n_chunks, GW, 100.0 * ws.noops / std::max<int64_t>(1, n_chunks * GW), lag, R,
            ws.build_ms);
  return C3_OK;
}

template <int K>
int launch_wave(c3_engine* e, const PruneArgs& a, const c3_engine::WaveSched& ws) {
  // every CTA must be resident (warps wait for other warps' progress): a cooperative launch, one CTA per SM
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ws.grid);
  cfg.blockDim = dim3(WAVE_WARPS * 32);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = e->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t err = cudaLaunchKernelEx(&cfg, prune4_wave_kernel<K>, a, (const int2*)ws.d_sched, ws.n_chunks, ws.d_done,
                                       e->wave_lag, e->wave_group);
  if (err != cudaSuccess) return fail(C3_ERR_CUDA, std::string("wave launch: ") + cudaGetErrorString(err));
  return C3_OK;
}

// kind: 0 = generic kernels (prune4_kernel / prune_dmma_kernel); 2, 3 = strip kernel for
// nodes with that many children (4-state, K in {1,2,4}); 6 = register-pipeline DMMA kernel
int launch_prune(c3_engine* e, const PruneArgs& a, bool is_root, int kind) {
  if (a.total_tiles <= 0) return C3_OK;
  if (kind == 8) {
    int rc = C3_OK;
    const bool r3 = is_root && a.n_work == 1 && e->children[e->root].size() == 3;
    switch (e->MP) {
      case 8: rc = launch_dmma_edge<8>(e, a, r3); break;
      case 16: rc = launch_dmma_edge<16>(e, a, r3); break;
      case 24: rc = launch_dmma_edge<24>(e, a, r3); break;
      case 32: rc = launch_dmma_edge<32>(e, a, r3); break;
      case 40: rc = launch_dmma_edge<40>(e, a, r3); break;
      case 48: rc = launch_dmma_edge<48>(e, a, r3); break;
      case 56: rc = launch_dmma_edge<56>(e, a, r3); break;
      case 64: rc = launch_dmma_edge<64>(e, a, r3); break;
      default: return fail(C3_ERR_INVALID, "unsupported padded state count");
    }
    if (rc) return rc;
  } else if (kind == 6) {
    int rc = C3_OK;
    switch (e->MP) {
      case 8: rc = launch_dmma_reg<8>(e, a, is_root); break;
      case 16: rc = launch_dmma_reg<16>(e, a, is_root); break;
      case 24: rc = launch_dmma_reg<24>(e, a, is_root); break;
      case 32: rc = launch_dmma_reg<32>(e, a, is_root); break;
      case 40: rc = launch_dmma_reg<40>(e, a, is_root); break;
      case 48: rc = launch_dmma_reg<48>(e, a, is_root); break;
      case 56: rc = launch_dmma_reg<56>(e, a, is_root); break;
      case 64: rc = launch_dmma_reg<64>(e, a, is_root); break;
      default: return fail(C3_ERR_INVALID, "unsupported padded state count");
    }
    if (rc) return rc;
  } else if (kind >= 2) {
    int rc = C3_ERR_INVALID;
#define C3_STRIP(KK)                                                                              \
  case KK:                                                                                        \
    if (e->use_ring)                                                                              \
      rc = (kind == 2) ? launch_strip<KK, 2>(e, a, is_root) : launch_strip<KK, 3>(e, a, is_root); \
    else if (e->use_reg2)                                                                         \
      rc = (kind == 2) ? launch_reg2<KK, 2>(e, a, is_root) : launch_reg2<KK, 3>(e, a, is_root);   \
    else                                                                                          \
      rc = (kind == 2) ? launch_reg<KK, 2>(e, a, is_root) : launch_reg<KK, 3>(e, a, is_root);     \
    break;
    switch (e->K) {
      C3_STRIP(1)
      C3_STRIP(2)
      C3_STRIP(4)
      default: return fail(C3_ERR_INVALID, "strip kernel: unsupported K");
    }
#undef C3_STRIP
    if (rc) return rc;
  } else if (!e->use_dmma) {
    const int grid = std::min(a.total_tiles, e->n_sm * 32);
#define C3_P4(KT)                                                                    \
  do {                                                                               \
    if (is_root)                                                                     \
      prune4_kernel<KT, true><<<grid, PRUNE4_THREADS, 0, e->stream>>>(a);            \
    else                                                                             \
      prune4_kernel<KT, false><<<grid, PRUNE4_THREADS, 0, e->stream>>>(a);           \
  } while (0)
    switch (e->K) {
      case 1: C3_P4(1); break;
      case 2: C3_P4(2); break;
      case 4: C3_P4(4); break;
      default: C3_P4(0); break;
    }
#undef C3_P4
  } else {
    const int grid = std::min(a.total_tiles, e->n_sm * dmma_blocks_per_sm(e->MP));
    int rc = C3_OK;
    switch (e->MP) {
      case 8: rc = launch_dmma<8>(e, a, is_root, grid); break;
      case 16: rc = launch_dmma<16>(e, a, is_root, grid); break;
      case 24: rc = launch_dmma<24>(e, a, is_root, grid); break;
      case 32: rc = launch_dmma<32>(e, a, is_root, grid); break;
      case 40: rc = launch_dmma<40>(e, a, is_root, grid); break;
      case 48: rc = launch_dmma<48>(e, a, is_root, grid); break;
      case 56: rc = launch_dmma<56>(e, a, is_root, grid); break;
      case 64: rc = launch_dmma<64>(e, a, is_root, grid); break;
      default: return fail(C3_ERR_INVALID, "unsupported padded state count");
    }
    if (rc) return rc;
  }
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return fail(C3_ERR_CUDA, std::string("prune launch: ") + cudaGetErrorString(err));
  return C3_OK;
}

// which kernel family handles node n (see launch_prune)
int kind_of_node(const c3_engine* e, int n) {
  if (e->use_dmma && e->edge_major) return 8;
  if (e->use_dmma) {
    // register-pipeline DMMA kernel; nodes with more children / contraction children than it
    // stages fall back to the first version
    if (e->use_dmma_reg && (int)e->children[n].size() <= DR_MAXC) {
      int n_contract = 0;
      for (int c : e->children[n]) n_contract += !e->children[c].empty();
      if (n_contract <= DR_MAXP) return 6;
    }
    return 0;
  }
  if (!e->use_strip) return 0;
  if (e->K != 1 && e->K != 2 && e->K != 4) return 0;
  const int nc = (int)e->children[n].size();
  return (nc == 2 || nc == 3) ? nc : 0;
}

int64_t tiles_of_node(const c3_engine* e, int n, int kind) {
  if (kind == 8) {
    // the trifurcating root is always alone in its launch and runs the 3-children instance (12 warps x 8 rows)
    const int tile_rows = (n == e->root && e->children[n].size() == 3) ? DECfg<1>::TILE_ROWS : DECfg<DE_RA, DE_STG>::TILE_ROWS;
    return (int64_t)e->K * ((e->n_rows[n] + tile_rows - 1) / tile_rows);
  }
  if (kind == 6) return (int64_t)e->K * ((e->n_rows[n] + DR_TILE_ROWS - 1) / DR_TILE_ROWS);
  if (kind >= 2) {
    // strips of 128 (row, bin) units; regpipe2 walks three-children nodes in strips of 64
    const int units = (!e->use_reg2 || e->use_ring) ? 128 : (kind == 3 ? 32 * e->root_u : (e->reg2_u2 ? 64 : 128));
    const int rows = units / e->K;
    return (e->n_rows[n] + rows - 1) / rows;
  }
  if (!e->use_dmma) {
    const int rpt = prune4_rows_per_tile(e->K);
    return (e->n_rows[n] + rpt - 1) / rpt;
  }
  return (int64_t)e->K * ((e->n_rows[n] + DMMA_TILE_ROWS - 1) / DMMA_TILE_ROWS);
}

// bookkeeping after the stream operations of a planned evaluation were issued (or replayed): everything is clean
void commit_evaluation(c3_engine* e, int64_t n_launch, bool replayed) {
  std::fill(e->pdirty.begin(), e->pdirty.end(), 0);
  std::fill(e->node_dirty.begin(), e->node_dirty.end(), 0);
  e->reduce_dirty = false;
  if (e->lh_keep > 0) --e->lh_keep;
  if (e->last_lean) e->stats.lean_evaluations += 1;
  if (e->last_chain) e->stats.chain_launches += 1;
  if (e->last_walk) e->stats.walk_launches += 1;
  e->stats.evaluations += 1;
  e->stats.kernel_launches += n_launch;
  e->stats.last_launches = n_launch;
  e->stats.last_expm_jobs = e->plan.expm_jobs;
  e->stats.last_nodes = e->plan.nodes;
  e->stats.last_node_rows = e->plan.node_rows;
  e->stats.last_alg_bytes = e->plan.alg_bytes;
  e->stats.last_flops = e->plan.flops;
  e->stats.last_h2d_bytes = e->plan.h2d_bytes;
  if (replayed) e->stats.graph_replays += 1;
}

// c3_evaluate_many: ONE graph over all engines of a batch (each engine's operations on its own
// stream, forked from / joined to the first engine's stream inside the capture)
struct BatchGraph {
  std::vector<c3_engine*> engines;
  std::vector<uint64_t> hashes;  // structure hash of every engine's evaluation
  std::vector<int64_t> n_launch;
  cudaGraphExec_t exec = nullptr;
  int seen = 0;
  void* d_copies = nullptr;  // [n] BatchCopy: the engines' parameter blocks, copied by ONE kernel at the head of the graph
};
std::unordered_map<uint64_t, BatchGraph> g_batches;  // (single-threaded use, like the engines themselves)
cudaEvent_t g_fork = nullptr;
std::vector<cudaEvent_t> g_joins;

void clear_batches() {
  std::lock_guard<std::recursive_mutex> lock(g_batch_mutex);
  for (auto& b : g_batches) {
    if (b.second.exec) cudaGraphExecDestroy(b.second.exec);
    if (b.second.d_copies) cudaFree(b.second.d_copies);
  }
  g_batches.clear();
}

// captured graphs hold device pointers: drop them whenever an allocation they name is replaced
void clear_graphs(c3_engine* e) {
  for (auto& g : e->graphs)
    if (g.second.exec) cudaGraphExecDestroy(g.second.exec);
  e->graphs.clear();
  e->graph_seen.clear();
  clear_batches();
}

int ensure_slots(c3_engine* e, int n) {
  if ((int)e->slots.size() < n) e->slots.resize(n);
  if (n <= e->n_slots_alloc) return C3_OK;
  const int new_n = std::max(n, std::max(4, 2 * e->n_slots_alloc));
  double* fresh = nullptr;
  C3_CUDA(cudaMalloc(&fresh, (size_t)new_n * e->slot_stride * sizeof(double)));
  clear_graphs(e);  // the expm launches name d_slots
  if (e->d_slots) {
    C3_CUDA(cudaStreamSynchronize(e->stream));
    C3_CUDA(cudaMemcpy(fresh, e->d_slots, (size_t)e->n_slots_alloc * e->slot_stride * sizeof(double),
                       cudaMemcpyDeviceToDevice));
    C3_CUDA(cudaFree(e->d_slots));
  }
  e->stats.device_bytes += (int64_t)(new_n - e->n_slots_alloc) * e->slot_stride * sizeof(double);
  e->d_slots = fresh;
  e->n_slots_alloc = new_n;
  return C3_OK;
}

int run_evaluation(c3_engine* e, double* device_out, bool blocking, double* host_out, bool deferred = false);
int finish_deferred(c3_engine* e, double* host_out);
int collect_result(c3_engine* e);
void resync_results(c3_engine* e);
bool wants_auto_rescale(const c3_engine* e);
int after_auto_rescale(c3_engine* e);
int evaluate_many_batched(c3_engine** engines, int n, double* lnL, int* used);
int activate_rescaling(c3_engine* e);
void deactivate_rescaling(c3_engine* e);
int launch_rescale(c3_engine* e, int node);
int ensure_lh(c3_engine* e);

cudaEvent_t next_event(c3_engine* e) {
  if (e->events_used == e->event_pool.size()) {
    cudaEvent_t ev;
    cudaEventCreate(&ev);
    e->event_pool.push_back(ev);
  }
  return e->event_pool[e->events_used++];
}

struct ProfScope {
  c3_engine* e;
  bool on;
  ProfScope(c3_engine* e_, int kind, int64_t bytes, int64_t flops) : e(e_), on(e_->profiling) {
    if (!on) return;
    c3_engine::LaunchRec r{kind, bytes, flops, next_event(e), next_event(e), 0.f};
    cudaEventRecord(r.a, e->stream);
    e->prof.push_back(r);
  }
  ~ProfScope() {
    if (on) cudaEventRecord(e->prof.back().b, e->stream);
  }
};

}  // namespace

// =========================================================================================
// C ABI
// =========================================================================================
extern "C" {

int c3_abi_version(void) { return C3_ABI_VERSION; }
const char* c3_last_error(void) { return g_last_error.c_str(); }

int c3_device_count(int* n) {
  C3_REQUIRE(n != nullptr, "null out pointer");
  cudaError_t err = cudaGetDeviceCount(n);
  if (err != cudaSuccess) {
    *n = 0;
    return fail(C3_ERR_CUDA, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(err));
  }
  return C3_OK;
}

int c3_create(int device, int n_nodes, int n_states, int n_bins, const int32_t* child_ptr,
              const int32_t* child_idx, c3_engine** out) {
  C3_REQUIRE(out != nullptr && child_ptr != nullptr, "null pointer");
  C3_REQUIRE(n_nodes >= 2, "need at least two nodes");
  C3_REQUIRE(n_states >= 2 && n_states <= 64, "n_states must be in [2, 64] (cogent3 asserts < 65)");
  C3_REQUIRE(n_bins >= 1 && n_bins <= 64, "n_bins must be in [1, 64]");
  int n_dev = 0;
  C3_CUDA(cudaGetDeviceCount(&n_dev));
  if (n_dev == 0) return fail(C3_ERR_CUDA, "no CUDA device: libc3cuda has no CPU fallback");
  C3_REQUIRE(device >= 0 && device < n_dev, "bad device ordinal");
  C3_CUDA(cudaSetDevice(device));
  c3_engine* e = new c3_engine();
  e->device = device;
  e->n_nodes = n_nodes;
  e->M = n_states;
  e->K = n_bins;
  e->use_dmma = (n_states != 4);
  if (const char* env = getenv("C3_FORCE_DMMA")) e->use_dmma = e->use_dmma || atoi(env) != 0;
  e->MP = e->use_dmma ? (int)align_up(n_states, 8) : 4;
  if (const char* env = getenv("C3_NO_STRIP")) e->use_strip = atoi(env) == 0;
  if (const char* env = getenv("C3_NO_SMALL")) e->use_small = atoi(env) == 0;
  if (const char* env = getenv("C3_NO_PDL")) e->use_pdl = atoi(env) == 0;
  if (const char* env = getenv("C3_NO_SMALL_REDUCE")) e->use_small_reduce = atoi(env) == 0;
  if (const char* env = getenv("C3_PATH")) e->use_path = atoi(env) != 0;
  if (const char* env = getenv("C3_CHAIN")) e->use_chain = atoi(env) != 0;
  if (const char* env = getenv("C3_WALK")) e->use_walk = atoi(env) != 0;
  if (const char* env = getenv("C3_CHAIN_FUSED")) e->use_chain_fused = atoi(env) != 0;
  if (const char* env = getenv("C3_PATH_VARIANT")) e->path_variant = std::max(0, std::min(2, atoi(env)));
  if (const char* env = getenv("C3_REC")) e->use_rec = atoi(env) != 0;
  if (const char* env = getenv("C3_STREAMS")) e->use_streams = atoi(env) != 0;
  if (const char* env = getenv("C3_REC_UNITS")) e->rec_units = std::max(1, atoi(env));
  if (const char* env = getenv("C3_GRID_SMALL")) e->use_grid_small = atoi(env) != 0;
  if (const char* env = getenv("C3_GRID_UNITS")) e->grid_units_per_thread = std::max(1, std::min(64, atoi(env)));
  if (const char* env = getenv("C3_DMMA_REG")) e->use_dmma_reg = atoi(env) != 0;
  e->edge_major = e->use_dmma;
  if (const char* env = getenv("C3_EDGE_MAJOR")) e->edge_major = e->use_dmma && atoi(env) != 0;
  if (const char* env = getenv("C3_RING")) e->use_ring = atoi(env) != 0;
  if (const char* env = getenv("C3_REG1")) e->use_reg2 = atoi(env) == 0;
  if (const char* env = getenv("C3_REG2_U2")) e->reg2_u2 = atoi(env) != 0;
  if (const char* env = getenv("C3_DYN")) e->use_dyn = atoi(env) != 0;
  if (const char* env = getenv("C3_ROOT_U")) e->root_u = atoi(env) == 2 ? 2 : 3;
  if (const char* env = getenv("C3_LEAN_ROOT")) e->lean_root = atoi(env) != 0;
  if (const char* env = getenv("C3_TIMELINE")) e->timeline = atoi(env) != 0;
  if (const char* env = getenv("C3_HOSTPROF")) e->hostprof = atoi(env) != 0;
  if (const char* env = getenv("C3_FLOW")) e->use_flow = atoi(env) != 0;
  if (const char* env = getenv("C3_WAVE")) e->use_wave = atoi(env) != 0;
  if (const char* env = getenv("C3_WAVE_LAG")) e->wave_lag = std::max(3, std::min(12, atoi(env)));
  if (const char* env = getenv("C3_WAVE_PRIO")) e->wave_prio = std::max(1, atoi(env));
  if (const char* env = getenv("C3_WAVE_MIN")) e->wave_min_strips = std::max(1, atoi(env));
  if (const char* env = getenv("C3_WAVE_GROUP")) e->wave_group = std::max(1, std::min(32, atoi(env)));
  if (const char* env = getenv("C3_WAVE_PREFIX")) e->wave_prefix = std::max(0, atoi(env));
  e->reg2_u2 = e->reg2_u2 && e->use_strip && e->use_reg2 && !e->use_ring && !e->use_flow && !e->use_wave;
  e->store_root = !(e->use_dmma ? e->edge_major
                                : (e->use_strip && e->use_reg2 && !e->use_ring && !e->use_flow));
  if (const char* env = getenv("C3_STORE_ROOT")) e->store_root = e->store_root || atoi(env) != 0;
  if (const char* env = getenv("C3_GRAPH")) e->use_graphs = atoi(env) != 0;
  if (const char* env = getenv("C3_GRAPH_AFTER")) e->graph_after = std::max(1, atoi(env));
  if (const char* env = getenv("C3_RESCALE")) e->rescale_mode = std::max(0, std::min(2, atoi(env)));
  if (const char* env = getenv("C3_RESCALE_TIPS")) e->rescale_threshold = std::max(2, atoi(env));
  e->root = n_nodes - 1;
  e->children.resize(n_nodes);
  e->parent.assign(n_nodes, -1);
  e->level.assign(n_nodes, 0);
  e->child_begin.assign(n_nodes, 0);
  int running = 0;
  for (int n = 0; n < n_nodes; ++n) {
    e->child_begin[n] = running;
    for (int k = child_ptr[n]; k < child_ptr[n + 1]; ++k) {
      const int c = child_idx[k];
      if (c < 0 || c >= n) {
        delete e;
        return fail(C3_ERR_INVALID, "nodes must be in post-order (child id < parent id)");
      }
      if (e->parent[c] != -1) {
        delete e;
        return fail(C3_ERR_INVALID, "node has two parents");
      }
      e->parent[c] = n;
      e->children[n].push_back(c);
      e->level[n] = std::max(e->level[n], e->level[c] + 1);
      ++running;
    }
  }
  for (int n = 0; n < n_nodes - 1; ++n)
    if (e->parent[n] < 0) {
      delete e;
      return fail(C3_ERR_INVALID, "only the last node may be the root");
    }
  if (e->children[e->root].empty()) {
    delete e;
    return fail(C3_ERR_INVALID, "root has no children");
  }
  e->subtree_of.assign(n_nodes, -1);
  {
    int t = 0;
    for (int c : e->children[e->root]) e->subtree_of[c] = t++;
    for (int n = n_nodes - 2; n >= 0; --n)  // parents before children (post-order numbering)
      if (e->subtree_of[n] < 0 && e->parent[n] >= 0) e->subtree_of[n] = e->subtree_of[e->parent[n]];
  }
  e->n_levels = e->level[e->root];
  e->level_nodes.resize(e->n_levels + 1);
  for (int n = 0; n < n_nodes; ++n)
    if (!e->children[n].empty()) e->level_nodes[e->level[n]].push_back(n);
  e->n_rows.assign(n_nodes, 0);
  e->idx_mono.assign(n_nodes, 0);
  e->tip_table.resize(n_nodes);
  e->idx32.resize(n_nodes);
  e->qslot.assign((size_t)n_nodes * n_bins, 0);
  e->dist.assign((size_t)n_nodes * n_bins, NAN);
  e->pdirty.assign((size_t)n_nodes * n_bins, 1);
  e->pfixed.assign((size_t)n_nodes * n_bins, 0);
  e->node_dirty.assign(n_nodes, 1);
  e->pi.assign(e->MP, 0.0);
  e->bprobs.assign(n_bins, 1.0 / n_bins);
  e->fixed_motif.assign(n_nodes, -1);
  e->slot_stride = 2 * (int64_t)n_states + 4 * (int64_t)n_states * n_states;
  cudaDeviceProp prop;
  C3_CUDA(cudaGetDeviceProperties(&prop, device));
  e->n_sm = prop.multiProcessorCount;
  C3_CUDA(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
  e->own_stream = true;
  C3_CUDA(cudaEventCreateWithFlags(&e->block_copied, cudaEventDisableTiming));
  C3_CUDA(cudaHostAlloc(&e->h_lnl, 64, cudaHostAllocMapped));
  std::memset(e->h_lnl, 0, 64);
  C3_CUDA(cudaHostGetDevicePointer(&e->h_lnl_dev, e->h_lnl, 0));
  e->h_seq = reinterpret_cast<volatile unsigned long long*>(e->h_lnl + 1);
  e->h_seq_dev = reinterpret_cast<unsigned long long*>(e->h_lnl_dev + 1);
  if (const char* env = getenv("C3_SPIN")) e->spin_wait = atoi(env) != 0;
  e->stats.n_levels = e->n_levels;
  e->stats.padded_states = e->MP;
  *out = e;
  return C3_OK;
}

// developer timeline of the LAST evaluation (C3_TIMELINE=1): microseconds after the first launch's entry
void dump_timeline(c3_engine* e) {
  if (!e->timeline || !e->d_tl || e->tl_launches <= 0) return;
  if (e->stream) cudaStreamSynchronize(e->stream);
  std::vector<unsigned long long> t((size_t)e->tl_launches * 4);
  if (cudaMemcpy(t.data(), e->d_tl, t.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost) != cudaSuccess) return;
  unsigned long long t0 = ~0ull;
  for (int i = 0; i < e->tl_launches; ++i) t0 = std::min(t0, t[4 * i]);
  fprintf(stderr, "c3 timeline (us) after %lld evaluations: launch kind | entry | start (after the dependency wait) | first warp out | last warp out\n",
          (long long)e->stats.evaluations);
  for (int i = 0; i < e->tl_launches; ++i) {
    auto us = [&](unsigned long long v) { return (v == ~0ull || v == 0ull) ? -1.0 : (double)(v - t0) * 1e-3; };
    fprintf(stderr, "c3 timeline %2d kind %3d | %8.2f | %8.2f | %8.2f | %8.2f\n", i,
            i < (int)e->tl_kinds.size() ? e->tl_kinds[i] : -1, us(t[4 * i]), us(t[4 * i + 1]), us(t[4 * i + 2]),
            us(t[4 * i + 3]));
  }
}

void dump_hostprof(c3_engine* e) {
  if (e->hostprof && e->hp_n > 0)
    fprintf(stderr, "c3 hostprof: %lld blocking evaluations since the last report, us per evaluation: plan %.2f | issue / replay %.2f | wait %.2f\n",
            e->hp_n, e->hp_plan_ns * 1e-3 / e->hp_n, e->hp_issue_ns * 1e-3 / e->hp_n, e->hp_wait_ns * 1e-3 / e->hp_n);
  e->hp_plan_ns = e->hp_issue_ns = e->hp_wait_ns = e->hp_n = 0;
}

int c3_destroy(c3_engine* e) {
  if (!e) return C3_OK;
  cudaSetDevice(e->device);
  if (e->stream) cudaStreamSynchronize(e->stream);
  dump_timeline(e);
  dump_hostprof(e);
  if (e->d_chain_maps) cudaFree(e->d_chain_maps);
  if (e->d_node_maps) cudaFree(e->d_node_maps);
  if (e->d_child_maps) cudaFree(e->d_child_maps);
  if (e->d_chain_ctr) cudaFree(e->d_chain_ctr);
  if (e->d_tl) cudaFree(e->d_tl);
  if (e->h_tl_init) cudaFreeHost(e->h_tl_init);
  for (int32_t* p : e->d_index)
    if (p) cudaFree(p);
  for (int32_t* p : e->d_idx_tmp)
    if (p) cudaFree(p);
  for (void* p : e->ipc_opened) cudaIpcCloseMemHandle(p);
  if (e->d_mailbox) cudaFree(e->d_mailbox);
  if (e->d_pool) cudaFree(e->d_pool);
  if (e->d_scale_pool) cudaFree(e->d_scale_pool);
  if (e->d_hmm_scratch) cudaFree(e->d_hmm_scratch);
  clear_graphs(e);
  for (auto& w : e->waves) free_wave(e, w.second);
  e->waves.clear();
  if (e->d_slots) cudaFree(e->d_slots);
  if (e->d_block) cudaFree(e->d_block);
  if (e->h_block) cudaFreeHost(e->h_block);
  if (e->h_lnl) cudaFreeHost(e->h_lnl);
  if (e->block_copied) cudaEventDestroy(e->block_copied);
  for (cudaStream_t st : e->aux_streams) cudaStreamDestroy(st);
  for (cudaEvent_t ev : e->aux_done) cudaEventDestroy(ev);
  if (e->aux_fork) cudaEventDestroy(e->aux_fork);
  for (cudaEvent_t ev : e->event_pool) cudaEventDestroy(ev);
  if (e->own_stream && e->stream) cudaStreamDestroy(e->stream);
  delete e;
  return C3_OK;
}

int c3_set_stream(c3_engine* e, void* cuda_stream) {
  C3_REQUIRE(e != nullptr, "null engine");
  C3_CUDA(cudaSetDevice(e->device));
  C3_CUDA(cudaStreamSynchronize(e->stream));
  if (e->own_stream) {
    C3_CUDA(cudaStreamDestroy(e->stream));
    e->own_stream = false;
  }
  if (cuda_stream == nullptr) {
    C3_CUDA(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    e->own_stream = true;
  } else {
    e->stream = (cudaStream_t)cuda_stream;
  }
  return C3_OK;
}

int c3_set_tip(c3_engine* e, int node, const double* table, int64_t n_rows) {
  C3_REQUIRE(e && table, "null pointer");
  C3_REQUIRE(!e->finalized, "pattern tables are frozen after c3_finalize");
  C3_REQUIRE(node >= 0 && node < e->n_nodes && e->children[node].empty(), "not a tip node");
  C3_REQUIRE(n_rows >= 1 && n_rows < INT32_MAX, "bad row count");
  e->n_rows[node] = n_rows;
  std::vector<double>& t = e->tip_table[node];
  t.assign((size_t)n_rows * e->MP, 0.0);
  for (int64_t u = 0; u < n_rows; ++u)
    for (int j = 0; j < e->M; ++j) t[u * e->MP + j] = table[u * e->M + j];
  return C3_OK;
}

int c3_set_node_patterns(c3_engine* e, int node, const int64_t* indexes, int64_t n_rows) {
  C3_REQUIRE(e && indexes, "null pointer");
  C3_REQUIRE(!e->finalized, "pattern tables are frozen after c3_finalize");
  C3_REQUIRE(node >= 0 && node < e->n_nodes && !e->children[node].empty(), "not an internal node");
  C3_REQUIRE(n_rows >= 1 && n_rows < INT32_MAX, "bad row count");
  const size_t C = e->children[node].size();
  e->n_rows[node] = n_rows;
  std::vector<int32_t>& ix = e->idx32[node];
  ix.resize(C * (size_t)n_rows);
  bool mono = true;
  for (size_t c = 0; c < C; ++c) {
    const int child = e->children[node][c];
    const int64_t child_rows = e->n_rows[child];
    C3_REQUIRE(child_rows > 0, "set the children's tables before their parent's");
    for (int64_t p = 0; p < n_rows; ++p) {
      const int64_t v = indexes[c * n_rows + p];
      if (v < 0 || v >= child_rows) return fail(C3_ERR_INVALID, "child row index out of range");
      ix[c * n_rows + p] = (int32_t)v;
      mono = mono && v <= p;
    }
  }
  e->idx_mono[node] = mono ? 1 : 0;  // cogent3's first-seen numbering guarantees it (wave kernel precondition)
  return C3_OK;
}

int c3_set_root_counts(c3_engine* e, const double* counts, int64_t n_rows) {
  C3_REQUIRE(e && counts, "null pointer");
  C3_REQUIRE(!e->finalized, "pattern tables are frozen after c3_finalize");
  C3_REQUIRE(n_rows == e->n_rows[e->root], "counts length must equal the root's row count");
  e->counts.assign(counts, counts + n_rows);
  return C3_OK;
}

// ---- device-side pattern compression ------------------------------------------------------
namespace {
struct CmpScratch {
  unsigned long long* keys = nullptr;
  unsigned* first = nullptr;
  unsigned* rank = nullptr;
  int32_t* slot_of_col = nullptr;
  unsigned* pos = nullptr;
  unsigned* tile_sums = nullptr;
  unsigned* total = nullptr;
  int64_t table_cap = 0;
  ~CmpScratch() {
    cudaFree(keys); cudaFree(first); cudaFree(rank); cudaFree(slot_of_col); cudaFree(pos);
    cudaFree(tile_sums); cudaFree(total);
  }
};

// first-seen numbering of key(col) = a[col]*rows_b + b[col] (or codes[col]); writes index[col],
// leaves pos[col] (first-occurrence columns) in the scratch; returns the number of distinct keys
int cmp_first_seen(c3_engine* e, CmpScratch& sc, const int32_t* a, const int32_t* b, const uint8_t* codes,
                   int64_t rows_a, int64_t rows_b, int64_t L, int32_t* index_out, int64_t* n_unique) {
  if (L == 0) {
    *n_unique = 0;
    return C3_OK;
  }
  // table size: a power of two >= 2 * (upper bound on distinct keys)
  const double span = codes ? 256.0 : (double)rows_a * (double)rows_b;
  const int64_t bound = (int64_t)std::min((double)L, span);
  int64_t size = 1024;
  while (size < 2 * bound) size *= 2;
  if (size > sc.table_cap) return fail(C3_ERR_INVALID, "compression hash table too small (internal)");
  CmpTable t{sc.keys, sc.first, sc.rank, (unsigned)(size - 1)};
  C3_CUDA(cudaMemsetAsync(sc.keys, 0xFF, (size_t)size * sizeof(unsigned long long), e->stream));
  C3_CUDA(cudaMemsetAsync(sc.first, 0xFF, (size_t)size * sizeof(unsigned), e->stream));
  const int grid = (int)std::min<int64_t>((L + CMP_THREADS - 1) / CMP_THREADS, (int64_t)e->n_sm * 16);
  const int n_tiles = (int)((L + CMP_SCAN_TILE - 1) / CMP_SCAN_TILE);
  cmp_insert_kernel<<<grid, CMP_THREADS, 0, e->stream>>>(a, b, codes, rows_b, L, t, sc.slot_of_col);
  cmp_count_kernel<<<n_tiles, CMP_THREADS, 0, e->stream>>>(L, t, sc.slot_of_col, sc.tile_sums);
  cmp_scan_tiles_kernel<<<1, CMP_THREADS, 0, e->stream>>>(sc.tile_sums, n_tiles, sc.total);
  cmp_number_kernel<<<n_tiles, CMP_THREADS, 0, e->stream>>>(L, t, sc.slot_of_col, sc.tile_sums, sc.pos);
  cmp_index_kernel<<<grid, CMP_THREADS, 0, e->stream>>>(L, t, sc.slot_of_col, index_out);
  unsigned total = 0;
  C3_CUDA(cudaMemcpyAsync(&total, sc.total, sizeof(unsigned), cudaMemcpyDeviceToHost, e->stream));
  C3_CUDA(cudaStreamSynchronize(e->stream));
  C3_CUDA(cudaGetLastError());
  *n_unique = total;
  return C3_OK;
}
}  // namespace

int c3_compress_alignment(c3_engine* e, const uint8_t* codes, int64_t n_columns, const double* symbol_rows,
                          int n_symbols, int keep_index) {
  C3_REQUIRE(e && symbol_rows && (codes || n_columns == 0), "null pointer");
  C3_REQUIRE(!e->finalized, "pattern tables are frozen after c3_finalize");
  C3_REQUIRE(n_columns >= 0 && n_columns < INT32_MAX - 1, "bad column count");
  C3_REQUIRE(n_symbols >= 1 && n_symbols <= 256, "n_symbols must be in [1, 256]");
  C3_CUDA(cudaSetDevice(e->device));
  const int N = e->n_nodes, M = e->M, MP = e->MP;
  const int64_t L = n_columns;
  int n_tips = 0;
  for (int n = 0; n < N; ++n) n_tips += e->children[n].empty();

  CmpScratch sc;
  {
    int64_t cap = 1024;
    while (cap < 2 * std::max<int64_t>(L, 256)) cap *= 2;
    sc.table_cap = cap;
    const size_t Lp = (size_t)std::max<int64_t>(L, 1);
    const size_t tiles = (Lp + CMP_SCAN_TILE - 1) / CMP_SCAN_TILE;
    C3_CUDA(cudaMalloc(&sc.keys, (size_t)cap * sizeof(unsigned long long)));
    C3_CUDA(cudaMalloc(&sc.first, (size_t)cap * sizeof(unsigned)));
    C3_CUDA(cudaMalloc(&sc.rank, (size_t)cap * sizeof(unsigned)));
    C3_CUDA(cudaMalloc(&sc.slot_of_col, Lp * sizeof(int32_t)));
    C3_CUDA(cudaMalloc(&sc.pos, Lp * sizeof(unsigned)));
    C3_CUDA(cudaMalloc(&sc.tile_sums, tiles * sizeof(unsigned)));
    C3_CUDA(cudaMalloc(&sc.total, sizeof(unsigned)));
  }
  for (int32_t* p : e->d_index)
    if (p) cudaFree(p);
  for (int32_t* p : e->d_idx_tmp)
    if (p) cudaFree(p);
  e->d_index.assign(N, nullptr);
  e->d_idx_tmp.assign(N, nullptr);
  e->n_columns = L;
  const size_t Lb = (size_t)std::max<int64_t>(L, 1);
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((L + CMP_THREADS - 1) / CMP_THREADS, (int64_t)e->n_sm * 16));

  uint8_t* d_codes = nullptr;
  C3_CUDA(cudaMalloc(&d_codes, Lb));
  int32_t* d_small = nullptr;  // first-seen values of the first columns (<= 256 for a tip)
  C3_CUDA(cudaMalloc(&d_small, 257 * sizeof(int32_t)));
  int32_t* d_fold = nullptr;  // intermediate numbering of a polytomy's folded children
  int rc = C3_OK;
  int tip_no = 0;
  for (int n = 0; n < N && rc == C3_OK; ++n) {
    const std::vector<int>& kids = e->children[n];
    cudaError_t err = cudaMalloc(&e->d_index[n], Lb * sizeof(int32_t));
    if (err != cudaSuccess) { rc = fail(C3_ERR_MEMORY, "device memory exhausted during compression"); break; }
    int64_t U = 0;
    if (kids.empty()) {
      // ---- tip (make_likelihood_tree_leaf) ---------------------------------------------------
      if (L > 0) {
        err = cudaMemcpyAsync(d_codes, codes + (int64_t)tip_no * L, (size_t)L, cudaMemcpyHostToDevice, e->stream);
        if (err != cudaSuccess) { rc = fail(C3_ERR_CUDA, cudaGetErrorString(err)); break; }
      }
      rc = cmp_first_seen(e, sc, nullptr, nullptr, d_codes, 1, 1, L, e->d_index[n], &U);
      if (rc) break;
      std::vector<int32_t> uniq((size_t)U);
      if (U > 0) {
        cmp_gather_first_u8_kernel<<<grid, CMP_THREADS, 0, e->stream>>>(L, sc.pos, d_codes, d_small);
        err = cudaMemcpyAsync(uniq.data(), d_small, (size_t)U * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream);
        if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);
        if (err != cudaSuccess) { rc = fail(C3_ERR_CUDA, cudaGetErrorString(err)); break; }
      }
      e->n_rows[n] = U + 1;
      std::vector<double>& tab = e->tip_table[n];
      tab.assign((size_t)(U + 1) * MP, 0.0);
      for (int64_t u = 0; u < U && rc == C3_OK; ++u) {
        if (uniq[u] >= n_symbols) rc = fail(C3_ERR_INVALID, "symbol code out of range");
        else
          for (int j = 0; j < M; ++j) tab[u * MP + j] = symbol_rows[(int64_t)uniq[u] * M + j];
      }
      for (int j = 0; j < M; ++j) tab[U * MP + j] = 1.0;  // the all-gap row
      ++tip_no;
    } else {
      // ---- internal node (LikelihoodTreeEdge.__init__) -------------------------------------
      const int C = (int)kids.size();
      const int32_t* acc = e->d_index[kids[0]];
      int64_t acc_rows = e->n_rows[kids[0]];
      if (C == 1) {
        // a single child: patterns are the child's; number them through the same machinery
        // key = 2 a + a is injective in a; handled like any other key
        rc = cmp_first_seen(e, sc, acc, acc, nullptr, acc_rows, 2, L, e->d_index[n], &U);
      }
      for (int c = 1; c < C && rc == C3_OK; ++c) {
        const bool last = (c == C - 1);
        int32_t* out = e->d_index[n];
        if (!last) {
          if (!d_fold) {
            err = cudaMalloc(&d_fold, 2 * Lb * sizeof(int32_t));
            if (err != cudaSuccess) { rc = fail(C3_ERR_MEMORY, "device memory exhausted during compression"); break; }
          }
          out = d_fold + ((c & 1) ? 0 : Lb);  // ping-pong: acc may point at the other half
        }
        rc = cmp_first_seen(e, sc, acc, e->d_index[kids[c]], nullptr, acc_rows, e->n_rows[kids[c]], L, out, &U);
        acc = out;
        acc_rows = U;  // intermediate numbering has no gap row
      }
      if (rc) break;
      const int64_t rows = U + 1;
      e->n_rows[n] = rows;
      err = cudaMalloc(&e->d_idx_tmp[n], (size_t)C * rows * sizeof(int32_t));
      if (err != cudaSuccess) { rc = fail(C3_ERR_MEMORY, "device memory exhausted during compression"); break; }
      std::vector<int32_t> gap(C);
      for (int c = 0; c < C; ++c) {
        if (L > 0)
          cmp_gather_first_kernel<<<grid, CMP_THREADS, 0, e->stream>>>(L, sc.pos, e->d_index[kids[c]],
                                                                      e->d_idx_tmp[n] + (int64_t)c * rows);
        gap[c] = (int32_t)(e->n_rows[kids[c]] - 1);
        err = cudaMemcpyAsync(e->d_idx_tmp[n] + (int64_t)c * rows + U, &gap[c], sizeof(int32_t),
                              cudaMemcpyHostToDevice, e->stream);
        if (err != cudaSuccess) { rc = fail(C3_ERR_CUDA, cudaGetErrorString(err)); break; }
      }
      // idx_c[p] <= p (first-seen numbering): checked, not assumed (sc.total is free again here)
      cudaMemsetAsync(sc.total, 0, sizeof(unsigned), e->stream);
      cmp_check_monotone_kernel<<<grid, CMP_THREADS, 0, e->stream>>>(e->d_idx_tmp[n], C, rows, sc.total);
      unsigned not_mono = 1;
      err = cudaMemcpyAsync(&not_mono, sc.total, sizeof(unsigned), cudaMemcpyDeviceToHost, e->stream);
      if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);
      if (err != cudaSuccess) { rc = fail(C3_ERR_CUDA, cudaGetErrorString(err)); break; }
      e->idx_mono[n] = not_mono ? 0 : 1;
      if (!keep_index)
        for (int k : kids) {
          cudaFree(e->d_index[k]);
          e->d_index[k] = nullptr;
        }
      if (n == e->root) {
        // counts (float64, like the reference's): histogram of the column numbering + the gap row
        unsigned* d_cnt = nullptr;
        double* d_cntf = nullptr;
        err = cudaMalloc(&d_cnt, (size_t)rows * sizeof(unsigned));
        if (err == cudaSuccess) err = cudaMalloc(&d_cntf, (size_t)rows * sizeof(double));
        if (err != cudaSuccess) { cudaFree(d_cnt); rc = fail(C3_ERR_MEMORY, "device memory exhausted during compression"); break; }
        cudaMemsetAsync(d_cnt, 0, (size_t)rows * sizeof(unsigned), e->stream);
        if (L > 0) cmp_histogram_kernel<<<grid, CMP_THREADS, 0, e->stream>>>(L, e->d_index[n], d_cnt);
        cmp_counts_to_double_kernel<<<(int)std::min<int64_t>((rows + 255) / 256, 4096), CMP_THREADS, 0, e->stream>>>(
            rows, d_cnt, d_cntf);
        e->counts.assign((size_t)rows, 0.0);
        err = cudaMemcpyAsync(e->counts.data(), d_cntf, (size_t)rows * sizeof(double), cudaMemcpyDeviceToHost, e->stream);
        if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);
        cudaFree(d_cnt);
        cudaFree(d_cntf);
        if (err != cudaSuccess) { rc = fail(C3_ERR_CUDA, cudaGetErrorString(err)); break; }
      }
    }
  }
  cudaFree(d_codes);
  cudaFree(d_small);
  cudaFree(d_fold);
  if (rc != C3_OK) return rc;
  e->d_root_index = e->d_index[e->root];
  return C3_OK;
}

int c3_get_node_rows(c3_engine* e, int64_t* rows) {
  C3_REQUIRE(e && rows, "null pointer");
  for (int n = 0; n < e->n_nodes; ++n) rows[n] = e->n_rows[n];
  return C3_OK;
}

int c3_get_node_indexes(c3_engine* e, int node, int64_t* out) {
  C3_REQUIRE(e && out, "null pointer");
  C3_REQUIRE(node >= 0 && node < e->n_nodes && !e->children[node].empty(), "not an internal node");
  C3_CUDA(cudaSetDevice(e->device));
  const size_t n = e->children[node].size() * (size_t)e->n_rows[node];
  std::vector<int32_t> tmp(n);
  const int32_t* src = nullptr;
  if (e->finalized) src = e->d_idx[node];
  else if (!e->d_idx_tmp.empty() && e->d_idx_tmp[node]) src = e->d_idx_tmp[node];
  if (src) {
    C3_CUDA(cudaStreamSynchronize(e->stream));
    C3_CUDA(cudaMemcpy(tmp.data(), src, n * sizeof(int32_t), cudaMemcpyDeviceToHost));
  } else {
    C3_REQUIRE(e->idx32[node].size() == n, "the node has no pattern table yet");
    tmp = e->idx32[node];
  }
  for (size_t i = 0; i < n; ++i) out[i] = tmp[i];
  return C3_OK;
}

int c3_get_tip_table(c3_engine* e, int node, double* out) {
  C3_REQUIRE(e && out, "null pointer");
  C3_REQUIRE(node >= 0 && node < e->n_nodes && e->children[node].empty(), "not a tip node");
  C3_CUDA(cudaSetDevice(e->device));
  const int M = e->M, MP = e->MP;
  const int64_t rows = e->n_rows[node];
  std::vector<double> tmp((size_t)rows * MP);
  if (e->finalized) {
    C3_CUDA(cudaStreamSynchronize(e->stream));
    C3_CUDA(cudaMemcpy(tmp.data(), e->d_tip_table[node], tmp.size() * sizeof(double), cudaMemcpyDeviceToHost));
  } else {
    C3_REQUIRE(e->tip_table[node].size() == tmp.size(), "the tip has no table yet");
    tmp = e->tip_table[node];
  }
  for (int64_t u = 0; u < rows; ++u)
    for (int j = 0; j < M; ++j) out[u * M + j] = tmp[u * MP + j];
  return C3_OK;
}

int c3_get_root_counts(c3_engine* e, double* out) {
  C3_REQUIRE(e && out, "null pointer");
  C3_CUDA(cudaSetDevice(e->device));
  if (e->finalized) {
    C3_CUDA(cudaStreamSynchronize(e->stream));
    C3_CUDA(cudaMemcpy(out, e->d_counts, (size_t)e->n_rows[e->root] * sizeof(double), cudaMemcpyDeviceToHost));
  } else {
    C3_REQUIRE(!e->counts.empty(), "root counts not set");
    std::copy(e->counts.begin(), e->counts.end(), out);
  }
  return C3_OK;
}

int c3_get_column_index(c3_engine* e, int node, int64_t* out) {
  C3_REQUIRE(e && out, "null pointer");
  C3_REQUIRE(node >= 0 && node < e->n_nodes, "bad node");
  C3_REQUIRE(e->n_columns >= 0 && !e->d_index.empty() && e->d_index[node] != nullptr,
             "no column numbering on the device for this node (c3_compress_alignment keeps the root's; "
             "keep_index=1 keeps all)");
  C3_CUDA(cudaSetDevice(e->device));
  std::vector<int32_t> tmp((size_t)e->n_columns);
  C3_CUDA(cudaStreamSynchronize(e->stream));
  if (e->n_columns > 0)
    C3_CUDA(cudaMemcpy(tmp.data(), e->d_index[node], tmp.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
  for (size_t i = 0; i < tmp.size(); ++i) out[i] = tmp[i];
  return C3_OK;
}

int c3_set_column_index(c3_engine* e, const int64_t* index, int64_t n_columns) {
  C3_REQUIRE(e && (index || n_columns == 0), "null pointer");
  C3_REQUIRE(n_columns >= 0 && n_columns < INT32_MAX - 1, "bad column count");
  C3_CUDA(cudaSetDevice(e->device));
  const int64_t rows = e->n_rows[e->root];
  C3_REQUIRE(rows > 0, "set the root's pattern table first");
  std::vector<int32_t> tmp((size_t)n_columns);
  for (int64_t i = 0; i < n_columns; ++i) {
    if (index[i] < 0 || index[i] >= rows) return fail(C3_ERR_INVALID, "column index out of range");
    tmp[i] = (int32_t)index[i];
  }
  if (e->d_index.empty()) e->d_index.assign(e->n_nodes, nullptr);
  if (e->d_index[e->root]) cudaFree(e->d_index[e->root]);
  e->d_index[e->root] = nullptr;
  C3_CUDA(cudaMalloc(&e->d_index[e->root], std::max<size_t>(1, tmp.size()) * sizeof(int32_t)));
  C3_CUDA(cudaMemcpy(e->d_index[e->root], tmp.data(), tmp.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  e->d_root_index = e->d_index[e->root];
  e->n_columns = n_columns;
  return C3_OK;
}

int c3_get_site_lh(c3_engine* e, int bin, double* out, int64_t n_columns) {
  C3_REQUIRE(e && (out || n_columns == 0), "null pointer");
  C3_REQUIRE(e->finalized && e->have_cached, "no evaluation yet");
  C3_REQUIRE(bin >= 0 && bin < e->K, "bad bin");
  C3_REQUIRE(e->d_root_index != nullptr, "the column numbering is not on the device (c3_compress_alignment / c3_set_column_index)");
  C3_REQUIRE(n_columns == e->n_columns, "bad column count");
  if (n_columns == 0) return C3_OK;
  C3_CUDA(cudaSetDevice(e->device));
  { const int rc_lh = ensure_lh(e); if (rc_lh) return rc_lh; }
  double* d_out = nullptr;
  C3_CUDA(cudaMalloc(&d_out, (size_t)n_columns * sizeof(double)));
  const int grid = (int)std::min<int64_t>((n_columns + CMP_THREADS - 1) / CMP_THREADS, (int64_t)e->n_sm * 16);
  site_lh_kernel<<<grid, CMP_THREADS, 0, e->stream>>>(n_columns, e->d_root_index, e->d_lh, e->K, bin, d_out,
                                                      e->rescale_active ? e->scale_rec[e->root].d_exps : nullptr);
  cudaError_t err = cudaMemcpyAsync(out, d_out, (size_t)n_columns * sizeof(double), cudaMemcpyDeviceToHost, e->stream);
  if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);
  cudaFree(d_out);
  if (err != cudaSuccess) return fail(C3_ERR_CUDA, std::string("c3_get_site_lh: ") + cudaGetErrorString(err));
  return C3_OK;
}

int c3_hmm_reduce(c3_engine* e, int n_patches, const int32_t* patch_of_bin, const double* weight_of_bin,
                  const double* switch_matrix, const double* patch_probs, double* lnL) {
  C3_REQUIRE(e && patch_of_bin && weight_of_bin && switch_matrix && patch_probs && lnL, "null pointer");
  C3_REQUIRE(e->finalized && e->have_cached, "no evaluation yet");
  C3_REQUIRE(n_patches == 2, "the device reduction handles two patches (cogent3's PatchSiteDistribution)");
  C3_REQUIRE(e->d_root_index != nullptr && e->n_columns >= 0,
             "the column numbering is not on the device (c3_compress_alignment / c3_set_column_index)");
  C3_REQUIRE(e->K <= 64, "too many bins");
  C3_CUDA(cudaSetDevice(e->device));
  { const int rc_lh = ensure_lh(e); if (rc_lh) return rc_lh; }
  HmmArgs a{};
  a.lh = e->d_lh;
  a.index = e->d_root_index;
  a.root_exp = e->rescale_active ? e->scale_rec[e->root].d_exps : nullptr;
  a.n_sites = e->n_columns;
  a.K = e->K;
  for (int b = 0; b < e->K; ++b) {
    C3_REQUIRE(patch_of_bin[b] == 0 || patch_of_bin[b] == 1, "bad patch number");
    a.patch_of_bin[b] = patch_of_bin[b];
    a.weight_of_bin[b] = weight_of_bin[b];
  }
  for (int i = 0; i < 4; ++i) a.sw[i] = switch_matrix[i];
  a.patch_probs[0] = patch_probs[0];
  a.patch_probs[1] = patch_probs[1];
  const int n_blocks = (int)std::max<int64_t>(1, std::min<int64_t>((e->n_columns + HMM_THREADS * 16 - 1) / (HMM_THREADS * 16),
                                                                 (int64_t)e->n_sm * 8));
  // scratch of the reduction: allocated on first use, kept (an optimiser calls this once per evaluation)
  const size_t scratch_bytes = (size_t)e->n_sm * 8 * sizeof(Hmm2) + sizeof(double);
  if (e->d_hmm_scratch == nullptr) {
    C3_CUDA(cudaMalloc(&e->d_hmm_scratch, scratch_bytes));
    e->stats.device_bytes += (int64_t)scratch_bytes;
  }
  Hmm2* d_blocks = reinterpret_cast<Hmm2*>(e->d_hmm_scratch);
  double* d_out = reinterpret_cast<double*>(d_blocks + (size_t)e->n_sm * 8);
  hmm_chunk_kernel<<<n_blocks, HMM_THREADS, 0, e->stream>>>(a, d_blocks);
  hmm_final_kernel<<<1, 32, 0, e->stream>>>(a, d_blocks, n_blocks, d_out);
  e->stats.kernel_launches += 2;
  cudaError_t err = cudaMemcpyAsync(lnL, d_out, sizeof(double), cudaMemcpyDeviceToHost, e->stream);
  if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);
  if (err != cudaSuccess) return fail(C3_ERR_CUDA, std::string("c3_hmm_reduce: ") + cudaGetErrorString(err));
  return C3_OK;
}

int c3_finalize(c3_engine* e) {
  C3_REQUIRE(e != nullptr, "null engine");
  C3_REQUIRE(!e->finalized, "already finalized");
  C3_CUDA(cudaSetDevice(e->device));
  const int N = e->n_nodes, K = e->K, MP = e->MP;
  for (int n = 0; n < N; ++n) C3_REQUIRE(e->n_rows[n] > 0, "a node has no pattern table");
  C3_REQUIRE(!e->counts.empty(), "root counts not set");

  // ---- carve one pool -----------------------------------------------------
  int64_t off = 0;
  auto carve = [&](int64_t bytes) {
    const int64_t o = off;
    off += align_up(bytes, 256);
    return o;
  };
  std::vector<int64_t> o_rows(N), o_tab(N, -1), o_idx(N, -1), o_state(N, -1);
  int64_t n_child_total = 0, partial_bytes = 0;
  for (int n = 0; n < N; ++n) {
    const int64_t bytes = (n == e->root && !e->store_root) ? 0 : e->n_rows[n] * K * MP * (int64_t)sizeof(double);
    o_rows[n] = carve(bytes);
    if (e->children[n].empty()) {
      o_tab[n] = carve(e->n_rows[n] * MP * (int64_t)sizeof(double));
      o_state[n] = carve(e->n_rows[n] * (int64_t)sizeof(int32_t));
    } else {
      partial_bytes += bytes;
      o_idx[n] = carve((int64_t)e->children[n].size() * e->n_rows[n] * (int64_t)sizeof(int32_t));
      n_child_total += (int64_t)e->children[n].size();
    }
  }
  const int64_t o_P = carve((int64_t)N * K * MP * MP * sizeof(double));
  const int64_t o_nodes = carve((int64_t)N * sizeof(NodeDesc));
  const int64_t o_childs = carve(n_child_total * sizeof(ChildRef));
  const int64_t o_edges = carve((int64_t)N * sizeof(EdgeDesc));
  const int64_t root_rows = e->n_rows[e->root];
  const int64_t o_counts = carve(root_rows * sizeof(double));
  const int64_t o_lh = carve(root_rows * K * sizeof(double));
  const int64_t o_mix = carve(root_rows * sizeof(double));
  const int64_t n_red_tiles = (root_rows + REDUCE_ROWS_PER_TILE - 1) / REDUCE_ROWS_PER_TILE;
  const int64_t o_tiles = carve(n_red_tiles * sizeof(double));
  const int64_t o_counter = carve(256);
  int n_internal_nodes = 0;
  for (int n = 0; n < N; ++n) n_internal_nodes += !e->children[n].empty();
  const int64_t o_done = carve((int64_t)n_internal_nodes * sizeof(unsigned int));
  const int64_t o_lnl = carve(256);

  size_t free_b = 0, total_b = 0;
  C3_CUDA(cudaMemGetInfo(&free_b, &total_b));
  if ((size_t)off > free_b) {
    char msg[256];
    snprintf(msg, sizeof msg,
             "engine needs %.2f GB of device memory (%.2f GB partial likelihoods) but only %.2f GB is free",
             off / 1e9, partial_bytes / 1e9, free_b / 1e9);
    return fail(C3_ERR_MEMORY, msg);
  }
  C3_CUDA(cudaMalloc(&e->d_pool, (size_t)off));
  e->pool_bytes = off;
  e->partial_bytes = partial_bytes;
  C3_CUDA(cudaMemsetAsync(e->d_pool + o_P, 0, (size_t)N * K * MP * MP * sizeof(double), e->stream));
  C3_CUDA(cudaMemsetAsync(e->d_pool + o_counter, 0, 512, e->stream));

  e->d_rows.resize(N);
  e->d_tip_table.assign(N, nullptr);
  e->d_idx.assign(N, nullptr);
  for (int n = 0; n < N; ++n) {
    e->d_rows[n] = (n == e->root && !e->store_root) ? nullptr : reinterpret_cast<double*>(e->d_pool + o_rows[n]);
    if (o_tab[n] >= 0) {
      e->d_tip_table[n] = reinterpret_cast<double*>(e->d_pool + o_tab[n]);
      C3_CUDA(cudaMemcpyAsync(e->d_tip_table[n], e->tip_table[n].data(),
                              e->tip_table[n].size() * sizeof(double), cudaMemcpyHostToDevice, e->stream));
      // which rows are plain indicators (one entry 1.0, the rest 0.0): their inner product with P is a column
      std::vector<int32_t> state((size_t)e->n_rows[n], -1);
      for (int64_t u = 0; u < e->n_rows[n]; ++u) {
        const double* row = e->tip_table[n].data() + u * MP;
        int hot = -1, nonzero = 0;
        for (int j = 0; j < MP; ++j)
          if (row[j] != 0.0) { ++nonzero; hot = j; }
        if (nonzero == 1 && row[hot] == 1.0 && hot < e->M) state[(size_t)u] = hot;
      }
      C3_CUDA(cudaMemcpyAsync(e->d_pool + o_state[n], state.data(), state.size() * sizeof(int32_t),
                              cudaMemcpyHostToDevice, e->stream));
      C3_CUDA(cudaStreamSynchronize(e->stream));  // `state` goes out of scope
      // padded tip_inner columns must be zero
      C3_CUDA(cudaMemsetAsync(e->d_rows[n], 0, (size_t)e->n_rows[n] * K * MP * sizeof(double), e->stream));
    } else {
      e->d_idx[n] = reinterpret_cast<int32_t*>(e->d_pool + o_idx[n]);
      if (!e->d_idx_tmp.empty() && e->d_idx_tmp[n] != nullptr) {
        C3_CUDA(cudaMemcpyAsync(e->d_idx[n], e->d_idx_tmp[n],
                                e->children[n].size() * (size_t)e->n_rows[n] * sizeof(int32_t),
                                cudaMemcpyDeviceToDevice, e->stream));
      } else {
        C3_CUDA(cudaMemcpyAsync(e->d_idx[n], e->idx32[n].data(), e->idx32[n].size() * sizeof(int32_t),
                                cudaMemcpyHostToDevice, e->stream));
      }
    }
  }
  e->d_P = reinterpret_cast<double*>(e->d_pool + o_P);
  e->d_nodes = reinterpret_cast<NodeDesc*>(e->d_pool + o_nodes);
  e->d_childs = reinterpret_cast<ChildRef*>(e->d_pool + o_childs);
  e->d_edges = reinterpret_cast<EdgeDesc*>(e->d_pool + o_edges);
  e->d_counts = reinterpret_cast<double*>(e->d_pool + o_counts);
  e->d_lh = reinterpret_cast<double*>(e->d_pool + o_lh);
  e->d_mix = reinterpret_cast<double*>(e->d_pool + o_mix);
  e->lh_valid = true;
  e->d_tile_sums = reinterpret_cast<double*>(e->d_pool + o_tiles);
  e->d_counter = reinterpret_cast<unsigned int*>(e->d_pool + o_counter);
  e->d_seq = reinterpret_cast<unsigned long long*>(e->d_pool + o_counter + 8);
  e->d_epoch = reinterpret_cast<unsigned long long*>(e->d_pool + o_counter + 16);
  e->d_grid_bar = reinterpret_cast<unsigned int*>(e->d_pool + o_counter + 32);
  e->d_lnl = reinterpret_cast<double*>(e->d_pool + o_lnl);
  e->d_done = reinterpret_cast<unsigned int*>(e->d_pool + o_done);
  C3_CUDA(cudaMemcpyAsync(e->d_counts, e->counts.data(), root_rows * sizeof(double),
                          cudaMemcpyHostToDevice, e->stream));

  // ---- descriptors -----------------------------------------------------------
  e->h_nodes.assign(N, NodeDesc{});
  std::vector<ChildRef> h_childs((size_t)n_child_total);
  std::vector<EdgeDesc> h_edges(N);
  for (int n = 0; n < N; ++n) {
    NodeDesc& nd = e->h_nodes[n];
    nd.out = e->children[n].empty() ? nullptr : e->d_rows[n];
    nd.n_rows = e->n_rows[n];
    nd.child_begin = e->child_begin[n];
    nd.n_children = (int)e->children[n].size();
    nd.fixed_motif = -1;
    nd.is_root = (n == e->root);
    nd.P_own = (e->edge_major && n != e->root && !e->children[n].empty()) ? e->d_P + (int64_t)n * K * MP * MP : nullptr;
    for (size_t c = 0; c < e->children[n].size(); ++c) {
      const int ch = e->children[n][c];
      ChildRef& cr = h_childs[e->child_begin[n] + c];
      cr.src = e->d_rows[ch];
      cr.idx = e->d_idx[n] + c * e->n_rows[n];
      // edge-major: every child's rows are already multiplied by its P (like tips always were)
      cr.P = (e->children[ch].empty() || e->edge_major) ? nullptr : e->d_P + (int64_t)ch * K * MP * MP;
      cr.src_rows = e->n_rows[ch];
    }
    EdgeDesc& ed = h_edges[n];
    ed.P = e->d_P + (int64_t)n * K * MP * MP;
    ed.tip_table = e->d_tip_table[n];
    ed.tip_state = o_state[n] >= 0 ? reinterpret_cast<const int32_t*>(e->d_pool + o_state[n]) : nullptr;
    ed.tip_inner = e->children[n].empty() ? e->d_rows[n] : nullptr;
    ed.tip_rows = e->children[n].empty() ? e->n_rows[n] : 0;
  }
  C3_CUDA(cudaMemcpyAsync(e->d_nodes, e->h_nodes.data(), N * sizeof(NodeDesc), cudaMemcpyHostToDevice, e->stream));
  C3_CUDA(cudaMemcpyAsync(e->d_childs, h_childs.data(), h_childs.size() * sizeof(ChildRef),
                          cudaMemcpyHostToDevice, e->stream));
  C3_CUDA(cudaMemcpyAsync(e->d_edges, h_edges.data(), N * sizeof(EdgeDesc), cudaMemcpyHostToDevice, e->stream));

  // ---- chain kernel: root row -> row of every node (composition of the gather tables from the root down) ----
  e->chain_ok = false;
  if (e->use_chain && !e->use_dmma && (K == 1 || K == 2 || K == 4) &&
      root_rows <= (int64_t)SMALL_REDUCE_MAX_TILES * REDUCE_ROWS_PER_TILE) {
    bool ok = true;
    for (int n = 0; n < N; ++n) ok = ok && e->children[n].size() <= (size_t)CHAIN_MAX_CHILDREN;
    if (ok) {
      C3_CUDA(cudaMalloc(&e->d_chain_maps, (size_t)N * root_rows * sizeof(int32_t)));
      C3_CUDA(cudaMalloc(&e->d_node_maps, (size_t)N * sizeof(int32_t*)));
      C3_CUDA(cudaMalloc(&e->d_child_maps, (size_t)std::max<int64_t>(1, (int64_t)n_child_total) * sizeof(int32_t*)));
      C3_CUDA(cudaMalloc(&e->d_chain_ctr, (1 + SMALL_REDUCE_MAX_TILES) * sizeof(unsigned int)));
      C3_CUDA(cudaMemsetAsync(e->d_chain_ctr, 0, (1 + SMALL_REDUCE_MAX_TILES) * sizeof(unsigned int), e->stream));
      std::vector<int32_t> iota((size_t)root_rows);
      for (int64_t p = 0; p < root_rows; ++p) iota[p] = (int32_t)p;
      auto map_of = [&](int n) { return e->d_chain_maps + (int64_t)n * root_rows; };
      C3_CUDA(cudaMemcpyAsync(map_of(e->root), iota.data(), root_rows * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
      std::vector<const int32_t*> node_maps(N), child_maps((size_t)std::max<int64_t>(1, (int64_t)n_child_total), nullptr);
      const int grid = (int)std::min<int64_t>((root_rows + 255) / 256, 1024);
      for (int n = N - 1; n >= 0; --n) {  // parents before children (post-order numbering)
        node_maps[n] = map_of(n);
        for (size_t c = 0; c < e->children[n].size(); ++c) {
          const int ch = e->children[n][c];
          compose_map_kernel<<<grid, 256, 0, e->stream>>>(map_of(ch), map_of(n), e->d_idx[n] + c * e->n_rows[n], root_rows);
          child_maps[e->child_begin[n] + c] = map_of(ch);
        }
      }
      // one storing root row per node row: the others get the top bit (map_first_kernel / map_mark_kernel); done
      // after every composition has read the clean parent maps (compose masks the bit anyway)
      {
        int64_t max_rows = 1;
        for (int n = 0; n < N; ++n) max_rows = std::max<int64_t>(max_rows, e->n_rows[n]);
        int32_t* d_first = nullptr;
        C3_CUDA(cudaMalloc(&d_first, (size_t)max_rows * sizeof(int32_t)));
        for (int n = 0; n < N; ++n) {
          if (e->children[n].empty() || n == e->root) continue;  // (tips are only gathered; the root map is the identity)
          C3_CUDA(cudaMemsetAsync(d_first, 0x7f, (size_t)e->n_rows[n] * sizeof(int32_t), e->stream));
          map_first_kernel<<<grid, 256, 0, e->stream>>>(map_of(n), d_first, root_rows);
          map_mark_kernel<<<grid, 256, 0, e->stream>>>(map_of(n), d_first, root_rows);
        }
        cudaError_t ferr = cudaStreamSynchronize(e->stream);
        cudaFree(d_first);
        if (ferr != cudaSuccess) return fail(C3_ERR_CUDA, std::string("chain maps: ") + cudaGetErrorString(ferr));
      }
      C3_CUDA(cudaGetLastError());
      C3_CUDA(cudaMemcpyAsync(e->d_node_maps, node_maps.data(), node_maps.size() * sizeof(int32_t*), cudaMemcpyHostToDevice, e->stream));
      C3_CUDA(cudaMemcpyAsync(e->d_child_maps, child_maps.data(), child_maps.size() * sizeof(int32_t*), cudaMemcpyHostToDevice, e->stream));
      C3_CUDA(cudaStreamSynchronize(e->stream));  // (the host vectors go out of scope)
      e->stats.device_bytes += (int64_t)N * root_rows * sizeof(int32_t);
      e->chain_ok = true;
    }
  }

  // ---- per-evaluation parameter block -------------------------------------
  int n_internal = 0;
  for (int n = 0; n < N; ++n) n_internal += !e->children[n].empty();
  int64_t b = 0;
  // the expm job list goes last: one copy of [0, end of used jobs) per evaluation
  e->off_pi = b;      b += align_up(MP * sizeof(double), 16);
  e->off_bprobs = b;  b += align_up(K * sizeof(double), 16);
  e->off_work = b;    b += align_up((int64_t)n_internal * sizeof(int32_t), 16);
  e->off_prefix = b;  b += align_up((int64_t)(2 * n_internal + 4 * e->n_levels + 8) * sizeof(int32_t), 16);
  e->off_small = b;   b += align_up((int64_t)(e->n_levels + 1) * sizeof(SmallLevel), 16);
  e->off_dep = b;     b += align_up((int64_t)2 * n_internal * sizeof(int32_t), 16);
  e->off_rec = b;     b += align_up((int64_t)REC_MAX_CHILDREN * n_internal * sizeof(int32_t), 16);
  e->off_ctr = b;     b += align_up((int64_t)DYN_MAX_LAUNCHES * sizeof(int32_t), 16);
  e->off_ops = b;     b += e->chain_ok ? align_up((int64_t)n_internal * sizeof(WalkOp), 16) : 0;
  e->off_jobs = b;    b += align_up((int64_t)N * K * sizeof(ExpmJob), 16);
  e->block_bytes = b;
  // (mapped: the copy kernel of a batch graph reads the block straight from this buffer)
  C3_CUDA(cudaHostAlloc(&e->h_block, (size_t)b, cudaHostAllocMapped | cudaHostAllocPortable));
  std::memset(e->h_block, 0, (size_t)b);  // (the claim counters of the dynamic-strip launches travel as zeros)
  C3_CUDA(cudaMalloc(&e->d_block, (size_t)b));
  C3_CUDA(cudaStreamSynchronize(e->stream));

  // host / temporary copies no longer needed
  for (auto& p : e->d_idx_tmp)
    if (p) {
      cudaFree(p);
      p = nullptr;
    }
  for (auto& v : e->tip_table) std::vector<double>().swap(v);
  for (auto& v : e->idx32) std::vector<int32_t>().swap(v);
  e->stats.device_bytes += off + b;
  e->stats.partial_bytes = partial_bytes;
  e->finalized = true;
  return C3_OK;
}

int c3_set_eigen(c3_engine* e, int q_slot, const double* roots, const double* evT, const double* evI,
                 int is_complex) {
  C3_REQUIRE(e && roots && evT && evI, "null pointer");
  C3_REQUIRE(q_slot >= 0 && q_slot < (1 << 20), "bad q slot");
  C3_CUDA(cudaSetDevice(e->device));
  int rc = ensure_slots(e, q_slot + 1);
  if (rc) return rc;
  Slot& s = e->slots[q_slot];
  const int M = e->M, MM = M * M;
  s.mode = is_complex ? EXPM_EIGEN_COMPLEX : EXPM_EIGEN_REAL;
  s.data.assign((size_t)e->slot_stride, 0.0);
  const int w = is_complex ? 2 : 1;
  std::memcpy(s.data.data(), roots, sizeof(double) * w * M);
  std::memcpy(s.data.data() + 2 * M, evT, sizeof(double) * w * MM);
  std::memcpy(s.data.data() + 2 * M + 2 * MM, evI, sizeof(double) * w * MM);
  s.dirty = true;
  return C3_OK;
}

int c3_set_Q(c3_engine* e, int q_slot, const double* Q) {
  C3_REQUIRE(e && Q, "null pointer");
  C3_REQUIRE(q_slot >= 0 && q_slot < (1 << 20), "bad q slot");
  C3_CUDA(cudaSetDevice(e->device));
  int rc = ensure_slots(e, q_slot + 1);
  if (rc) return rc;
  Slot& s = e->slots[q_slot];
  s.mode = EXPM_PADE;
  s.data.assign((size_t)e->slot_stride, 0.0);
  std::memcpy(s.data.data(), Q, sizeof(double) * e->M * e->M);
  s.dirty = true;
  return C3_OK;
}

int c3_set_tn93(c3_engine* e, int q_slot, const double* pi, double alpha1, double alpha2) {
  C3_REQUIRE(e && pi, "null pointer");
  C3_REQUIRE(e->M == 4 && !e->use_dmma, "the closed form is for four nucleotide states (T, C, A, G)");
  C3_REQUIRE(q_slot >= 0 && q_slot < (1 << 20), "bad q slot");
  C3_CUDA(cudaSetDevice(e->device));
  int rc = ensure_slots(e, q_slot + 1);
  if (rc) return rc;
  Slot& s = e->slots[q_slot];
  s.mode = EXPM_TN93;
  s.data.assign((size_t)e->slot_stride, 0.0);
  for (int i = 0; i < 4; ++i) s.data[i] = pi[i];
  s.data[4] = alpha1;
  s.data[5] = alpha2;
  s.dirty = true;
  return C3_OK;
}

int c3_set_edge_qslots(c3_engine* e, const int32_t* q) {
  C3_REQUIRE(e && q, "null pointer");
  for (int n = 0; n < e->n_nodes - 1; ++n)
    for (int b = 0; b < e->K; ++b) {
      const size_t i = (size_t)n * e->K + b;
      if (q[i] < 0) continue;  // negative: leave this (edge, bin) as it is (e.g. a supplied P)
      if (q[i] != e->qslot[i] || e->pfixed[i]) {
        e->qslot[i] = q[i];
        e->pfixed[i] = 0;
        e->pdirty[i] = 1;
      }
    }
  return C3_OK;
}

int c3_set_psub(c3_engine* e, int node, int bin, const double* P) {
  C3_REQUIRE(e && P, "null pointer");
  C3_REQUIRE(e->finalized, "call c3_finalize first");
  C3_REQUIRE(node >= 0 && node < e->n_nodes - 1 && bin >= 0 && bin < e->K, "bad edge / bin");
  C3_CUDA(cudaSetDevice(e->device));
  const int M = e->M, MP = e->MP;
  std::vector<double> padded((size_t)MP * MP, 0.0);
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < M; ++j) padded[i * MP + j] = P[i * M + j];
  // (lean root: per-pattern likelihoods of the LAST evaluation are recomputed from the P matrices on the device when
  //  an accessor asks -- materialise them before this call overwrites one of those matrices)
  if (!e->lh_valid && e->have_cached) {
    const int keep = e->lh_keep;
    const int rc = ensure_lh(e);
    e->lh_keep = keep;
    if (rc) return rc;
  }
  C3_CUDA(cudaStreamSynchronize(e->stream));
  C3_CUDA(cudaMemcpy(e->d_P + ((int64_t)node * e->K + bin) * MP * MP, padded.data(),
                     padded.size() * sizeof(double), cudaMemcpyHostToDevice));
  const size_t i = (size_t)node * e->K + bin;
  e->pfixed[i] = 1;
  e->pdirty[i] = 1;
  return C3_OK;
}

int c3_set_distances(c3_engine* e, const double* d) {
  C3_REQUIRE(e && d, "null pointer");
  const size_t n = (size_t)(e->n_nodes - 1) * e->K;
  for (size_t i = 0; i < n; ++i) {
    // bitwise compare: NaN (the initial state) is always "changed"
    if (std::memcmp(&d[i], &e->dist[i], sizeof(double)) != 0) {
      e->dist[i] = d[i];
      if (!e->pfixed[i]) e->pdirty[i] = 1;
    }
  }
  e->have_dist = true;
  return C3_OK;
}

int c3_set_root_mprobs(c3_engine* e, const double* pi) {
  C3_REQUIRE(e && pi, "null pointer");
  bool changed = !e->have_pi;
  for (int i = 0; i < e->M; ++i) changed = changed || (pi[i] != e->pi[i]);
  if (changed) {
    std::fill(e->pi.begin(), e->pi.end(), 0.0);
    std::copy(pi, pi + e->M, e->pi.begin());
    e->node_dirty[e->root] = 1;
    e->have_pi = true;
  }
  return C3_OK;
}

int c3_set_bin_probs(c3_engine* e, const double* bp) {
  C3_REQUIRE(e && bp, "null pointer");
  for (int b = 0; b < e->K; ++b)
    if (bp[b] != e->bprobs[b]) {
      e->bprobs[b] = bp[b];
      e->reduce_dirty = true;
    }
  return C3_OK;
}

int c3_set_fixed_motifs(c3_engine* e, const int32_t* fm) {
  C3_REQUIRE(e && fm, "null pointer");
  C3_REQUIRE(e->finalized, "call c3_finalize first");
  for (int n = 0; n < e->n_nodes; ++n) {
    const int32_t v = (fm[n] < 0) ? -1 : fm[n];
    C3_REQUIRE(v < e->M, "fixed motif out of range");
    if (v != e->fixed_motif[n]) {
      e->fixed_motif[n] = v;
      if (!e->children[n].empty()) {
        e->h_nodes[n].fixed_motif = v;
        e->node_dirty[n] = 1;
        e->nodes_dirty = true;
      }
    }
  }
  return C3_OK;
}

int c3_invalidate(c3_engine* e) {
  C3_REQUIRE(e != nullptr, "null engine");
  for (size_t i = 0; i < e->pdirty.size(); ++i) e->pdirty[i] = 1;
  std::fill(e->node_dirty.begin(), e->node_dirty.end(), 1);
  e->reduce_dirty = true;
  e->have_cached = false;
  return C3_OK;
}

int c3_set_rescaling(c3_engine* e, int mode, int tip_threshold) {
  C3_REQUIRE(e != nullptr, "null engine");
  C3_REQUIRE(mode >= C3_RESCALE_OFF && mode <= C3_RESCALE_AUTO, "bad rescaling mode");
  if (tip_threshold > 0 && std::max(2, tip_threshold) != e->rescale_threshold) {
    e->rescale_threshold = std::max(2, tip_threshold);
    if (e->rescale_active) {
      // a different set of scaled nodes: rebuild the tables at the next evaluation
      C3_CUDA(cudaSetDevice(e->device));
      deactivate_rescaling(e);
      if (mode == C3_RESCALE_AUTO) mode = C3_RESCALE_ON;
    }
  }
  e->rescale_mode = mode;  // ON / OFF take effect at the next evaluation
  if (mode != C3_RESCALE_AUTO && (mode == C3_RESCALE_ON) != e->rescale_active) {
    e->reduce_dirty = true;
    e->have_cached = false;
  }
  return C3_OK;
}

int c3_get_rescaling(c3_engine* e, int* active, int* n_scaled_nodes) {
  C3_REQUIRE(e != nullptr, "null engine");
  if (active) *active = e->rescale_active ? 1 : 0;
  if (n_scaled_nodes) *n_scaled_nodes = e->rescale_active ? e->n_scaled : 0;
  return C3_OK;
}

int c3_get_root_exponents(c3_engine* e, int32_t* out, int64_t n_rows) {
  C3_REQUIRE(e && out, "null pointer");
  C3_REQUIRE(e->finalized && e->have_cached, "no evaluation yet");
  C3_REQUIRE(n_rows == e->n_rows[e->root], "bad row count");
  if (!e->rescale_active) {
    std::fill(out, out + n_rows, 0);
    return C3_OK;
  }
  C3_CUDA(cudaSetDevice(e->device));
  C3_CUDA(cudaStreamSynchronize(e->stream));
  C3_CUDA(cudaMemcpy(out, e->scale_rec[e->root].d_exps, (size_t)n_rows * sizeof(int32_t), cudaMemcpyDeviceToHost));
  return C3_OK;
}

int c3_get_scaled_lh(c3_engine* e, int bin, double* out, int64_t n_rows) {
  C3_REQUIRE(e && out, "null pointer");
  C3_REQUIRE(e->finalized && e->have_cached, "no evaluation yet");
  C3_REQUIRE(bin >= 0 && bin < e->K, "bad bin");
  C3_REQUIRE(n_rows == e->n_rows[e->root], "bad row count");
  C3_CUDA(cudaSetDevice(e->device));
  { const int rc_lh = ensure_lh(e); if (rc_lh) return rc_lh; }
  C3_CUDA(cudaStreamSynchronize(e->stream));
  C3_CUDA(cudaMemcpy2D(out, sizeof(double), e->d_lh + bin, e->K * sizeof(double), sizeof(double),
                       (size_t)n_rows, cudaMemcpyDeviceToHost));
  return C3_OK;
}

int c3_evaluate(c3_engine* e, double* lnL) {
  C3_REQUIRE(e && lnL, "null pointer");
  return run_evaluation(e, nullptr, true, lnL);
}

int c3_evaluate_device(c3_engine* e, double* device_lnL) {
  C3_REQUIRE(e && device_lnL, "null pointer");
  return run_evaluation(e, device_lnL, false, nullptr);
}

int c3_evaluate_many(c3_engine** engines, int n, double* lnL) {
  C3_REQUIRE(engines && lnL && n >= 0, "null pointer");
  for (int i = 0; i < n; ++i) C3_REQUIRE(engines[i] != nullptr, "null engine");
  {
    int used_batch = 0;
    const int brc = evaluate_many_batched(engines, n, lnL, &used_batch);
    if (brc != C3_OK) return brc;
    if (used_batch) return C3_OK;
  }
  // launch every engine's evaluation on its own stream (and device), then collect: small problems
  // overlap, and engines whose reductions exchange shard values (c3_comm_attach) are all in flight
  // before anybody waits
  int rc = C3_OK;
  int launched = 0;
  for (; launched < n; ++launched) {
    rc = run_evaluation(engines[launched], nullptr, false, nullptr, true);
    if (rc != C3_OK) break;
  }
  bool any_auto = false;
  for (int i = 0; i < launched; ++i) {
    c3_engine* e = engines[i];
    if (e->deferred_pending) {
      e->deferred_pending = false;
      const int rc2 = collect_result(e);
      if (rc == C3_OK) rc = rc2;
    }
    lnL[i] = e->cached_lnl;
    any_auto = any_auto || (rc == C3_OK && wants_auto_rescale(e));
  }
  if (rc != C3_OK || launched < n || !any_auto) return rc;
  // a -inf somewhere: the engines concerned switch to rescaled rows TOGETHER (sharded engines all see
  // the same total and must re-evaluate in lock-step) and are evaluated again, launched before waited for
  std::vector<c3_engine*> again;
  for (int i = 0; i < n; ++i)
    if (wants_auto_rescale(engines[i])) again.push_back(engines[i]);
  for (c3_engine* e : again) {
    rc = activate_rescaling(e);
    if (rc != C3_OK) return rc;
  }
  for (c3_engine* e : again) {
    rc = run_evaluation(e, nullptr, false, nullptr, true);
    if (rc != C3_OK) return rc;
  }
  for (c3_engine* e : again) {
    e->deferred_pending = false;
    const int rc2 = collect_result(e);
    if (rc == C3_OK) rc = rc2;
  }
  if (rc == C3_OK)
    for (c3_engine* e : again) after_auto_rescale(e);
  for (int i = 0; i < n; ++i) lnL[i] = engines[i]->cached_lnl;
  return rc;
}

int c3_update_many(c3_engine** engines, int n, const double* distances, int64_t stride, double* lnL) {
  C3_REQUIRE(engines && distances && lnL && n >= 0, "null pointer");
  for (int i = 0; i < n; ++i) {
    C3_REQUIRE(engines[i] != nullptr, "null engine");
    C3_REQUIRE(stride == 0 || stride >= (int64_t)engines[i]->n_nodes * engines[i]->K, "stride shorter than an engine's distance table");
    const int rc = c3_set_distances(engines[i], distances + (int64_t)i * stride);
    if (rc != C3_OK) return rc;
  }
  return c3_evaluate_many(engines, n, lnL);
}

int c3_comm_mailbox(c3_engine* e, int nranks, void** device_ptr, unsigned char* ipc_handle) {
  C3_REQUIRE(e != nullptr, "null engine");
  C3_REQUIRE(nranks >= 1 && nranks <= C3_MAX_RANKS, "nranks must be in [1, 16]");
  C3_CUDA(cudaSetDevice(e->device));
  if (e->d_mailbox == nullptr || e->mailbox_ranks != nranks) {
    C3_REQUIRE(e->comm.nranks == 0, "detach before resizing the mailbox");
    if (e->d_mailbox) C3_CUDA(cudaFree(e->d_mailbox));
    e->d_mailbox = nullptr;
    const size_t bytes = (size_t)2 * nranks * sizeof(CommSlot);
    C3_CUDA(cudaMalloc(&e->d_mailbox, bytes));
    C3_CUDA(cudaMemset(e->d_mailbox, 0, bytes));  // epoch 0 is never used by an evaluation
    e->mailbox_ranks = nranks;
  }
  if (device_ptr) *device_ptr = e->d_mailbox;
  if (ipc_handle) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
    cudaIpcMemHandle_t h;
    C3_CUDA(cudaIpcGetMemHandle(&h, e->d_mailbox));
    std::memcpy(ipc_handle, &h, sizeof h);
  }
  return C3_OK;
}

int c3_comm_attach(c3_engine* e, int nranks, int rank, const unsigned char* ipc_handles, void* const* device_ptrs) {
  C3_REQUIRE(e != nullptr, "null engine");
  C3_REQUIRE(nranks >= 1 && nranks <= C3_MAX_RANKS && rank >= 0 && rank < nranks, "bad rank / nranks");
  C3_REQUIRE(e->d_mailbox != nullptr && e->mailbox_ranks == nranks, "call c3_comm_mailbox(nranks) first");
  C3_REQUIRE(ipc_handles != nullptr || device_ptrs != nullptr || nranks == 1, "no peer mailboxes given");
  C3_REQUIRE(e->comm.nranks == 0, "already attached");
  C3_CUDA(cudaSetDevice(e->device));
  C3_REQUIRE(e->finalized, "call c3_finalize first");
  CommArgs c{};
  c.nranks = nranks;
  c.rank = rank;
  c.epoch_ctr = e->d_epoch;
  C3_CUDA(cudaStreamSynchronize(e->stream));
  // evaluations are numbered from 1 (epoch 0 is what an empty mailbox slot holds).  The mailbox itself
  // is zeroed when it is allocated and again by c3_comm_detach -- never here: a peer that attached
  // earlier may already be writing into it
  C3_CUDA(cudaMemset(e->d_epoch, 0, sizeof(unsigned long long)));
  for (int r = 0; r < nranks; ++r) {
    if (r == rank) {
      c.peer[r] = e->d_mailbox;
    } else if (device_ptrs != nullptr) {
      C3_REQUIRE(device_ptrs[r] != nullptr, "null peer mailbox");
      // an engine on another GPU of this process: its mailbox must be mapped into this device
      cudaPointerAttributes attr{};
      if (cudaPointerGetAttributes(&attr, device_ptrs[r]) == cudaSuccess && attr.type == cudaMemoryTypeDevice &&
          attr.device != e->device) {
        int can = 0;
        C3_CUDA(cudaDeviceCanAccessPeer(&can, e->device, attr.device));
        if (!can) return fail(C3_ERR_CUDA, "no peer access between the devices of the sharded engines");
        const cudaError_t perr = cudaDeviceEnablePeerAccess(attr.device, 0);
        if (perr != cudaSuccess && perr != cudaErrorPeerAccessAlreadyEnabled)
          return fail(C3_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(perr));
        cudaGetLastError();
      } else {
        cudaGetLastError();
      }
      c.peer[r] = static_cast<CommSlot*>(device_ptrs[r]);
    } else {
      cudaIpcMemHandle_t h;
      std::memcpy(&h, ipc_handles + (size_t)r * sizeof h, sizeof h);
      void* p = nullptr;
      cudaError_t err = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
      if (err != cudaSuccess) {
        for (void* q : e->ipc_opened) cudaIpcCloseMemHandle(q);
        e->ipc_opened.clear();
        return fail(C3_ERR_CUDA, std::string("cudaIpcOpenMemHandle (peer access between the GPUs?): ") +
                                     cudaGetErrorString(err));
      }
      e->ipc_opened.push_back(p);
      c.peer[r] = static_cast<CommSlot*>(p);
    }
  }
  e->comm = c;
  clear_graphs(e);  // the reduction's arguments changed
  // a cached value is a shard value: it must not be served as the total
  e->have_cached = false;
  e->reduce_dirty = true;
  return C3_OK;
}

int c3_comm_detach(c3_engine* e) {
  C3_REQUIRE(e != nullptr, "null engine");
  C3_CUDA(cudaSetDevice(e->device));
  C3_CUDA(cudaStreamSynchronize(e->stream));
  for (void* p : e->ipc_opened) cudaIpcCloseMemHandle(p);
  e->ipc_opened.clear();
  // every rank has consumed the values of the last common evaluation: nobody writes here any more
  if (e->d_mailbox) C3_CUDA(cudaMemset(e->d_mailbox, 0, (size_t)2 * e->mailbox_ranks * sizeof(CommSlot)));
  e->comm = CommArgs{};
  clear_graphs(e);
  e->have_cached = false;
  e->reduce_dirty = true;
  return C3_OK;
}

int c3_update(c3_engine* e, const double* d, double* lnL) {
  int rc = c3_set_distances(e, d);
  if (rc) return rc;
  return c3_evaluate(e, lnL);
}

int c3_get_lh(c3_engine* e, int bin, double* out, int64_t n_rows) {
  C3_REQUIRE(e && out, "null pointer");
  C3_REQUIRE(e->finalized && e->have_cached, "no evaluation yet");
  C3_REQUIRE(bin >= 0 && bin < e->K, "bad bin");
  C3_REQUIRE(n_rows == e->n_rows[e->root], "bad row count");
  C3_CUDA(cudaSetDevice(e->device));
  { const int rc_lh = ensure_lh(e); if (rc_lh) return rc_lh; }
  C3_CUDA(cudaStreamSynchronize(e->stream));
  C3_CUDA(cudaMemcpy2D(out, sizeof(double), e->d_lh + bin, e->K * sizeof(double), sizeof(double),
                       (size_t)n_rows, cudaMemcpyDeviceToHost));
  if (e->rescale_active) {
    // fold the per-pattern exponents back: exact unless the true value leaves the FP64 range
    std::vector<int32_t> ex((size_t)n_rows);
    C3_CUDA(cudaMemcpy(ex.data(), e->scale_rec[e->root].d_exps, (size_t)n_rows * sizeof(int32_t),
                       cudaMemcpyDeviceToHost));
    for (int64_t p = 0; p < n_rows; ++p) out[p] = std::ldexp(out[p], ex[p]);
  }
  return C3_OK;
}

int c3_get_psub(c3_engine* e, int node, int bin, double* out) {
  C3_REQUIRE(e && out, "null pointer");
  C3_REQUIRE(e->finalized && e->have_cached, "no evaluation yet");
  C3_REQUIRE(node >= 0 && node < e->n_nodes - 1 && bin >= 0 && bin < e->K, "bad edge / bin");
  C3_CUDA(cudaSetDevice(e->device));
  C3_CUDA(cudaStreamSynchronize(e->stream));
  const int M = e->M, MP = e->MP;
  C3_CUDA(cudaMemcpy2D(out, M * sizeof(double), e->d_P + ((int64_t)node * e->K + bin) * MP * MP,
                       MP * sizeof(double), M * sizeof(double), M, cudaMemcpyDeviceToHost));
  return C3_OK;
}

int c3_get_partials(c3_engine* e, int node, int bin, double* out, int64_t n_rows) {
  C3_REQUIRE(e && out, "null pointer");
  C3_REQUIRE(e->finalized && e->have_cached, "no evaluation yet");
  C3_REQUIRE(node >= 0 && node < e->n_nodes && bin >= 0 && bin < e->K, "bad node / bin");
  C3_REQUIRE(n_rows == e->n_rows[node], "bad row count");
  C3_CUDA(cudaSetDevice(e->device));
  C3_CUDA(cudaStreamSynchronize(e->stream));
  const int M = e->M, MP = e->MP;
  if ((e->edge_major && node != e->root && !e->children[node].empty()) || (node == e->root && !e->store_root)) {
    // the stored rows are P . plh (edge-major) or not stored at all (the root); the accessor returns
    // plh = product of the children's (P-multiplied) rows
    double* tmp = nullptr;
    C3_CUDA(cudaMalloc(&tmp, (size_t)n_rows * MP * sizeof(double)));
    const int grid = (int)std::min<int64_t>((n_rows * MP + 255) / 256, (int64_t)e->n_sm * 8);
    edge_major_plh_kernel<<<grid, 256, 0, e->stream>>>(e->h_nodes[node], e->d_childs, e->K, M, MP, bin, tmp);
    cudaError_t err = cudaStreamSynchronize(e->stream);
    if (err == cudaSuccess)
      err = cudaMemcpy2D(out, M * sizeof(double), tmp, MP * sizeof(double), M * sizeof(double), (size_t)n_rows,
                         cudaMemcpyDeviceToHost);
    cudaFree(tmp);
    if (err != cudaSuccess) return fail(C3_ERR_CUDA, std::string("c3_get_partials: ") + cudaGetErrorString(err));
    return C3_OK;
  }
  // rows are [row][bin][MP]; for a tip this returns the P-multiplied tip table
  C3_CUDA(cudaMemcpy2D(out, M * sizeof(double), e->d_rows[node] + (int64_t)bin * MP,
                       (size_t)e->K * MP * sizeof(double), M * sizeof(double), (size_t)n_rows,
                       cudaMemcpyDeviceToHost));
  return C3_OK;
}

int c3_get_stats(c3_engine* e, c3_stats* out) {
  C3_REQUIRE(e && out, "null pointer");
  *out = e->stats;
  return C3_OK;
}

int c3_set_profiling(c3_engine* e, int on) {
  C3_REQUIRE(e != nullptr, "null engine");
  if (on != 0 && !e->profiling) {
    dump_timeline(e);  // (C3_TIMELINE=1: the last UN-profiled evaluation)
    dump_hostprof(e);
  }
  e->profiling = on != 0;
  return C3_OK;
}

int c3_get_launch_profile(c3_engine* e, int max_n, int32_t* kind, double* ms, int64_t* alg_bytes,
                          int64_t* flops, int* n) {
  C3_REQUIRE(e && kind && ms && alg_bytes && flops && n, "null pointer");
  const int m = std::min<int>(max_n, (int)e->prof.size());
  for (int i = 0; i < m; ++i) {
    kind[i] = e->prof[i].kind;
    ms[i] = e->prof[i].ms;
    alg_bytes[i] = e->prof[i].bytes;
    flops[i] = e->prof[i].flops;
  }
  *n = m;
  return C3_OK;
}

int c3_measure_dmma_peak(int device, double* tflops) {
  C3_REQUIRE(tflops != nullptr, "null pointer");
  C3_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  C3_CUDA(cudaGetDeviceProperties(&prop, device));
  double* d_out = nullptr;
  C3_CUDA(cudaMalloc(&d_out, 64));
  cudaEvent_t a, b;
  C3_CUDA(cudaEventCreate(&a));
  C3_CUDA(cudaEventCreate(&b));
  const int grid = prop.multiProcessorCount * 8, iters = 20000;
  dmma_peak_kernel<<<grid, 256>>>(d_out, 200);
  C3_CUDA(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    C3_CUDA(cudaEventRecord(a));
    dmma_peak_kernel<<<grid, 256>>>(d_out, iters);
    C3_CUDA(cudaEventRecord(b));
    C3_CUDA(cudaEventSynchronize(b));
    float ms = 0;
    C3_CUDA(cudaEventElapsedTime(&ms, a, b));
    best = std::min(best, ms);
  }
  const double flops = (double)grid * 8 /*warps*/ * iters * 8 /*dmma per iter*/ * (2.0 * 8 * 8 * 4);
  *tflops = flops / (best * 1e-3) / 1e12;
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  cudaFree(d_out);
  return C3_OK;
}

}  // extern "C"

// =========================================================================================
// evaluation
// =========================================================================================
namespace {

// deferred: launch everything like a blocking call (the scalar also goes to the mapped host word)
// but do not wait; finish_deferred() collects it.  Lets many engines run concurrently.
// Lean root: the last evaluation's root launch wrote mix[p], not lh[p][K].  An accessor that needs lh runs that
// ONE launch again (same work list and parameter block, still on the device; the children's rows, the P
// matrices and pi are those of the last evaluation) with the lh output -- no reduction, no result, nothing
// committed: input changes still pending stay pending.  The next LH_KEEP evaluations write lh themselves.
constexpr int LH_KEEP = 8;
int ensure_lh(c3_engine* e) {
  e->lh_keep = LH_KEEP;
  if (e->lh_valid) return C3_OK;
  if (e->deferred_pending) {
    const int rc = collect_result(e);
    e->deferred_pending = false;
    if (rc) return rc;
  }
  const c3_engine::RootLaunch& rl = e->root_launch;
  PruneArgs a;
  a.nodes = e->d_nodes;
  a.childs = e->d_childs;
  a.work_node = reinterpret_cast<const int32_t*>(e->d_block + e->off_work) + rl.work_off;
  a.tile_prefix = reinterpret_cast<const int32_t*>(e->d_block + e->off_prefix) + rl.prefix_off;
  a.n_work = rl.n_work;
  a.total_tiles = rl.total_tiles;
  a.K = e->K;
  a.M = e->M;
  a.pi = reinterpret_cast<const double*>(e->d_block + e->off_pi);
  a.lh_out = e->d_lh;
  const int rc = launch_prune(e, a, true, rl.kind);
  if (rc) return rc;
  C3_CUDA(cudaStreamSynchronize(e->stream));
  e->stats.kernel_launches += 1;
  e->lh_valid = true;
  return C3_OK;
}

int run_evaluation(c3_engine* e, double* device_out, bool blocking, double* host_out, bool deferred) {
  C3_REQUIRE(e->finalized, "call c3_finalize first");
  const auto hp_t0 = std::chrono::steady_clock::now();
  e->deferred_pending = false;
  C3_REQUIRE(e->have_pi, "root motif probabilities not set");
  C3_CUDA(cudaSetDevice(e->device));
  const int N = e->n_nodes, K = e->K, M = e->M, MP = e->MP;
  if (e->rescale_mode == C3_RESCALE_ON && !e->rescale_active) {
    int rc = activate_rescaling(e);
    if (rc) return rc;
  } else if (e->rescale_mode == C3_RESCALE_OFF && e->rescale_active) {
    deactivate_rescaling(e);
  }

  // slots whose contents changed dirty every (edge, bin) that uses them
  bool any_slot_dirty = false;
  for (const Slot& s : e->slots) any_slot_dirty = any_slot_dirty || s.dirty;
  if (any_slot_dirty) {
    for (int n = 0; n < N - 1; ++n)
      for (int b = 0; b < K; ++b) {
        const size_t i = (size_t)n * K + b;
        if (!e->pfixed[i] && e->qslot[i] < (int)e->slots.size() && e->slots[e->qslot[i]].dirty)
          e->pdirty[i] = 1;
      }
  }

  // ---- expm jobs ----------------------------------------------------------------
  // staging buffer free again?  (after a blocking evaluation whose result has been collected the stream is idle:
  // nothing to wait for -- one driver call less on the latency path)
  if (!e->block_free) C3_CUDA(cudaEventSynchronize(e->block_copied));
  e->block_free = false;
  ExpmJob* jobs = reinterpret_cast<ExpmJob*>(e->h_block + e->off_jobs);
  int n_eigen = 0, n_pade = 0;
  std::vector<char> edge_touched(N, 0);
  // first pass: eigen-type and fixed jobs, second: Pade (two launches)
  for (int pass = 0; pass < 2; ++pass) {
    for (int n = 0; n < N - 1; ++n)
      for (int b = 0; b < K; ++b) {
        const size_t i = (size_t)n * K + b;
        if (!e->pdirty[i]) continue;
        int mode;
        if (e->pfixed[i]) {
          mode = EXPM_FIXED;
        } else {
          const int s = e->qslot[i];
          if (s >= (int)e->slots.size() || e->slots[s].mode < 0)
            return fail(C3_ERR_INVALID, "an (edge, bin) uses a q slot that was never set");
          // (a NaN distance that WAS set flows through like in the reference: NaN P, NaN lnL, which the
          //  optimiser treats as out of bounds, maths/optimisers.py:116-118)
          if (!e->have_dist) return fail(C3_ERR_INVALID, "distances not set");
          mode = e->slots[s].mode;
        }
        const bool is_pade = (mode == EXPM_PADE);
        if (is_pade != (pass == 1)) continue;
        ExpmJob& j = jobs[n_eigen + n_pade];
        j.node = n;
        j.bin = b;
        j.slot = e->pfixed[i] ? 0 : e->qslot[i];
        j.mode = mode;
        j.t = e->pfixed[i] ? 0.0 : e->dist[i];
        if (is_pade) ++n_pade; else ++n_eigen;
        edge_touched[n] = 1;
      }
  }

  // ---- dirty nodes, level by level --------------------------------------------------
  std::vector<char> recompute(N, 0);
  int n_dirty_nodes = 0;
  for (int n = 0; n < N; ++n) {
    if (e->children[n].empty()) continue;
    bool d = e->node_dirty[n] != 0;
    if (e->edge_major) {
      // a node's rows include the P of the branch above it: that branch dirties the node itself
      for (int c : e->children[n]) d = d || recompute[c] || (edge_touched[c] && e->children[c].empty());
      d = d || (n != e->root && edge_touched[n]);
    } else {
      for (int c : e->children[n]) d = d || edge_touched[c] || recompute[c];
    }
    // (lean root: the mixed values of the last evaluation are stale once the bin probabilities change, and there
    //  is no lh to mix again: the root launch runs again)
    if (n == e->root && e->reduce_dirty && !e->lh_valid) d = true;
    if (d) {
      recompute[n] = 1;
      ++n_dirty_nodes;
    }
  }
  const bool root_dirty = recompute[e->root] != 0;
  e->last_skipped = false;
  if (n_dirty_nodes == 0 && n_eigen + n_pade == 0 && !e->reduce_dirty && e->have_cached) {
    e->last_skipped = true;
    if (e->batch_mode == 2) return C3_OK;
    if (host_out) *host_out = e->cached_lnl;
    if (device_out)
      C3_CUDA(cudaMemcpyAsync(device_out, e->d_lnl, sizeof(double), cudaMemcpyDeviceToDevice, e->stream));
    e->stats.last_launches = 0;
    return C3_OK;
  }

  // ---- wavefront launch?  every dirty non-root node binary, first-seen tables, enough strips, and the
  // structure has come up before (its schedule is built once and kept on the device) --------------------
  c3_engine::WaveSched* wave = nullptr;
  std::vector<int> wave_nodes;
  if (e->use_wave && !e->use_dmma && e->use_strip && e->use_reg2 && !e->use_ring && !e->use_flow &&
      (K == 1 || K == 2 || K == 4) && !e->rescale_active && e->batch_mode == 0 && recompute[e->root]) {
    bool ok = true;
    int64_t strips = 0;
    const int rows_per_strip = 128 / K;
    // prefix mode: only the lower levels, up to (not including) the first level with more strips than
    // wave_prefix; the big levels above keep their own launches
    int top_level = e->n_levels;
    if (e->wave_prefix > 0) {
      top_level = 0;
      int n_levels_in = 0;
      for (int l = 1; l <= e->n_levels; ++l) {
        int64_t s_l = 0;
        for (int n : e->level_nodes[l])
          if (recompute[n] && n != e->root) s_l += (e->n_rows[n] + rows_per_strip - 1) / rows_per_strip;
        if (s_l > e->wave_prefix) break;
        top_level = l;
        n_levels_in += s_l > 0;
      }
      ok = n_levels_in >= 2 && top_level < e->n_levels;
    }
    for (int n = 0; n < N - 1 && ok; ++n) {
      if (!recompute[n] || e->level[n] > top_level) continue;
      ok = e->children[n].size() == 2 && e->idx_mono[n];
      wave_nodes.push_back(n);
      strips += (e->n_rows[n] + rows_per_strip - 1) / rows_per_strip;
    }
    ok = ok && !wave_nodes.empty() && (int)wave_nodes.size() <= WAVE_MAXW &&
         strips >= (e->wave_prefix > 0 ? 64 : e->wave_min_strips);
    if (ok) {
      uint64_t h = 1469598103934665603ull;
      for (int n : wave_nodes) { h ^= (uint32_t)n; h *= 1099511628211ull; }
      auto it = e->waves.find(h);
      if (it != e->waves.end() && it->second.work.size() == wave_nodes.size() &&
          std::equal(wave_nodes.begin(), wave_nodes.end(), it->second.work.begin())) {
        wave = &it->second;
      } else if (it == e->waves.end() && ++e->wave_seen[h] >= 2) {
        c3_engine::WaveSched ws;
        const int rc = build_wave_schedule(e, wave_nodes, ws);
        if (rc == C3_OK) {
          wave = &e->waves.emplace(h, std::move(ws)).first->second;
          e->wave_seen.erase(h);
        } else if (rc != C3_ERR_MEMORY) {
          return rc;
        }
      } else if (e->wave_seen.size() > 4096) {
        e->wave_seen.clear();
      }
    }
    if (wave) wave->last_use = ++e->graph_clock;
  }
  std::vector<char> in_wave(N, 0);
  if (wave)
    for (int n : wave_nodes) in_wave[n] = 1;

  int32_t* work = reinterpret_cast<int32_t*>(e->h_block + e->off_work);
  int32_t* prefix = reinterpret_cast<int32_t*>(e->h_block + e->off_prefix);
  // kind: 0 generic, 1 fused small levels, 2/3 strip kernel (see launch_prune), 7 the wavefront launch
  struct LevelLaunch {
    int work_off, prefix_off, n_work, total_tiles;
    bool is_root;
    int64_t bytes, flops;
    int kind;
    int small_off, n_small;  // kind 1: range of SmallLevel records
    int grid;                // kind 1: > 0 = a cooperative launch of that many CTAs with grid barriers (medium levels)
    int stream;              // 0: the engine's stream; t > 0: the stream of root subtree t
  };
  std::vector<LevelLaunch> launches;
  SmallLevel* small = reinterpret_cast<SmallLevel*>(e->h_block + e->off_small);
  int n_small_total = 0;
  int w_off = 0, p_off = 0;
  int64_t node_rows = 0, alg_bytes = 0, flops = 0;
  auto account = [&](int n) {
    const int64_t rows = e->n_rows[n];
    node_rows += rows;
    if (n != e->root || e->store_root) alg_bytes += rows * K * M * 8;  // write
    for (int c : e->children[n]) {
      alg_bytes += e->n_rows[c] * K * M * 8 + rows * 4;  // child rows once + gather index
      if (!e->edge_major && !e->children[c].empty())
        flops += rows * K * 2 * (int64_t)(e->use_dmma ? MP * MP : M * M);
    }
    if (e->edge_major && n != e->root) flops += rows * K * 2 * (int64_t)MP * MP;  // one contraction per node row
    if (n == e->root) alg_bytes += rows * K * 8;
  };
  int open_small = -1;  // index into launches of the small group being extended
  // dataflow group (kind 5): the dirty binary nodes of consecutive big levels, one launch
  int open_flow = -1;
  int64_t flow_tiles = 0;
  std::vector<int> flow_item(N, -1);  // node -> work item of the open dataflow group
  int32_t* dep = reinterpret_cast<int32_t*>(e->h_block + e->off_dep);
  bool any_flow = false;
  auto close_flow = [&]() {
    if (open_flow < 0) return;
    LevelLaunch& g = launches[open_flow];
    prefix[p_off++] = (int32_t)flow_tiles;
    g.total_tiles = (int)flow_tiles;
    for (int i = 0; i < g.n_work; ++i) flow_item[work[g.work_off + i]] = -1;
    open_flow = -1;
  };
  const bool flow_ok = e->use_flow && !e->use_dmma && e->use_strip && !e->use_ring && (K == 1 || K == 2 || K == 4) &&
                       !e->rescale_active;  // rescaling runs between levels
  if (wave) {
    // one launch for every dirty node below the root; the per-level loop below then only sees the root
    LevelLaunch ll{w_off, p_off, 0, 0, false, 0, 0, 7, 0, 0};
    const int64_t bytes0 = alg_bytes, flops0 = flops;
    for (int n : wave_nodes) {
      work[w_off + ll.n_work++] = n;
      account(n);
      recompute[n] = 0;  // (handled; n_dirty_nodes keeps counting it)
    }
    ll.total_tiles = (int)wave->strips;
    ll.bytes = alg_bytes - bytes0;
    ll.flops = flops - flops0;
    w_off += ll.n_work;
    launches.push_back(ll);
  }
  // ---- path launch?  the dirty set is ONE path to the root (a single branch length moved: the optimiser's
  // inner loop) on a problem small enough for one CTA's shared memory: prune4_path_kernel, which also reduces
  // ---- chain launch?  the dirty set is ONE path to the root on a small problem: prune4_chain_kernel walks it
  // per root row without a barrier, and reduces
  bool chain_launch = false, chain_fused = false;
  int chain_edge_node = -1, chain_edge_slot = 0;
  if (e->chain_ok && e->use_chain && !wave && !e->rescale_active && e->comm.nranks <= 1 && recompute[e->root] &&
      e->use_small_reduce) {
    std::vector<int> chain;
    bool ok = true;
    for (int n = 0; n < N && ok; ++n) {
      if (!recompute[n]) continue;
      if (!chain.empty() && e->parent[chain.back()] != n) ok = false;  // (ascending ids: a path is visited bottom-up)
      chain.push_back(n);
    }
    ok = ok && !chain.empty() && chain.back() == e->root && (int)chain.size() <= CHAIN_MAX_NODES;
    if (ok) {
      LevelLaunch ll{w_off, p_off, 0, 0, true, 0, 0, 11, 0, 0};
      const int64_t bytes0 = alg_bytes, flops0 = flops;
      for (size_t i = 0; i < chain.size(); ++i) {
        const int n = chain[i];
        work[w_off + ll.n_work] = n;
        int slot = -1;
        if (i > 0)
          for (size_t c = 0; c < e->children[n].size(); ++c)
            if (e->children[n][c] == chain[i - 1]) slot = (int)c;
        dep[w_off + ll.n_work] = slot;
        ++ll.n_work;
        account(n);
        recompute[n] = 0;
      }
      ll.total_tiles = (int)(e->n_rows[e->root] * K);
      ll.bytes = alg_bytes - bytes0;
      ll.flops = flops - flops0;
      w_off += ll.n_work;
      launches.push_back(ll);
      chain_launch = true;
      // ONE launch for the whole update?  the jobs are the K rate classes of one branch under the lowest path node,
      // all in the real eigen form; a plain blocking evaluation (batches and profiled runs keep the launch sequence)
      if (n_pade == 0 && n_eigen == K && K <= 4 && M == 4 && e->batch_mode == 0 && !e->profiling && e->use_chain_fused) {
        bool one = true;
        for (int j = 0; j < K; ++j)
          one = one && jobs[j].node == jobs[0].node && jobs[j].bin == j && jobs[j].mode == EXPM_EIGEN_REAL;
        one = one && e->parent[jobs[0].node] == chain[0];
        if (one) {
          chain_fused = true;
          chain_edge_node = jobs[0].node;
          chain_edge_slot = 0;
          for (size_t c = 0; c < e->children[chain[0]].size(); ++c)
            if (e->children[chain[0]][c] == chain_edge_node) chain_edge_slot = (int)c;
        }
      }
    }
  }
  // ---- walk launch?  any other dirty set of such a problem: prune4_walk_kernel evaluates every dirty node per root
  // row, dirty children from a per-thread value stack (slots assigned here), clean ones from memory; it reduces
  bool walk_launch = false;
  int walk_slots = 0, walk_nP = 0, walk_ops = 0;
  if (e->chain_ok && e->use_walk && !chain_launch && !wave && !e->rescale_active && e->comm.nranks <= 1 &&
      recompute[e->root] && e->use_small_reduce && n_dirty_nodes <= WALK_MAX_OPS) {
    WalkOp* ops = reinterpret_cast<WalkOp*>(e->h_block + e->off_ops);
    std::vector<int> slot_of(N, -1);
    std::vector<char> busy(WALK_MAX_SLOTS + 1, 0);
    bool ok = true;
    int n_ops = 0;
    const int64_t root_rows = e->n_rows[e->root];
    for (int n = 0; n < N && ok; ++n) {
      if (!recompute[n]) continue;
      WalkOp& op = ops[n_ops++];
      std::memset(&op, 0, sizeof op);
      op.out = e->h_nodes[n].out;
      op.map_own = e->d_chain_maps + (int64_t)n * root_rows;
      op.n_children = (int8_t)e->children[n].size();
      op.fixed_motif = (int8_t)e->h_nodes[n].fixed_motif;
      op.is_root = n == e->root;
      for (int c = 0; c < 3; ++c) {
        op.slot[c] = -1;
        op.p_idx[c] = -1;
      }
      for (size_t c = 0; c < e->children[n].size(); ++c) {
        const int ch = e->children[n][c];
        op.src[c] = e->d_rows[ch];
        op.P[c] = e->children[ch].empty() ? nullptr : e->d_P + (int64_t)ch * K * 16;
        if (op.P[c] != nullptr) op.p_idx[c] = (int16_t)walk_nP++;
        op.map_child[c] = e->d_chain_maps + (int64_t)ch * root_rows;
        if (recompute[ch]) {
          op.slot[c] = (int8_t)slot_of[ch];
          busy[slot_of[ch]] = 0;  // (read before this node's own value is pushed: the slot may be taken again)
        }
        const int kind = recompute[ch] ? 3 : (op.P[c] != nullptr ? 2 : 1);
        op.shape = (uint8_t)(op.shape | (kind << (2 * c)));
      }
      op.out_slot = -1;
      if (n != e->root) {
        int sl = 0;
        while (sl < WALK_MAX_SLOTS && busy[sl]) ++sl;
        if (sl == WALK_MAX_SLOTS) { ok = false; break; }
        busy[sl] = 1;
        slot_of[n] = sl;
        op.out_slot = (int8_t)sl;
        walk_slots = std::max(walk_slots, sl + 1);
      }
    }
    walk_ops = n_ops;
    auto walk_smem = [&]() {
      return (size_t)walk_ops * sizeof(WalkOp) + (size_t)walk_slots * 2 * WALK_THREADS * sizeof(double2) +
             (size_t)walk_nP * K * 16 * sizeof(double) + (size_t)walk_nP * sizeof(double*);
    };
    ok = ok && walk_smem() <= (size_t)200 * 1024;
    if (ok && n_ops > 0) {
      LevelLaunch ll{w_off, p_off, 0, 0, true, 0, 0, 12, 0, 0};
      const int64_t bytes0 = alg_bytes, flops0 = flops;
      for (int n = 0; n < N; ++n) {
        if (!recompute[n]) continue;
        work[w_off + ll.n_work++] = n;
        account(n);
        recompute[n] = 0;
      }
      ll.total_tiles = (int)(root_rows * K);
      ll.bytes = alg_bytes - bytes0;
      ll.flops = flops - flops0;
      w_off += ll.n_work;
      launches.push_back(ll);
      walk_launch = true;
    }
  }
  bool path_launch = false;
  int path_buf_units = 0;
  if (e->use_path && !chain_launch && !walk_launch && !wave && !e->use_dmma && !e->rescale_active && e->comm.nranks <= 1 && recompute[e->root] &&
      e->n_rows[e->root] <= (int64_t)SMALL_REDUCE_MAX_TILES * REDUCE_ROWS_PER_TILE) {
    std::vector<int> chain;
    bool ok = true;
    int64_t max_units = 0;
    for (int n = 0; n < N && ok; ++n) {
      if (!recompute[n]) continue;
      if (!chain.empty() && e->parent[chain.back()] != n) ok = false;  // (ascending ids: a path is visited bottom-up)
      ok = ok && e->children[n].size() <= (size_t)PATH_MAX_CHILDREN;
      chain.push_back(n);
      if (n != e->root) max_units = std::max<int64_t>(max_units, e->n_rows[n] * K);
    }
    ok = ok && !chain.empty() && chain.back() == e->root && (int)chain.size() <= PATH_MAX_NODES;
    const size_t head = (sizeof(PathSmemHead) + 255) / 256 * 256;
    const size_t smem = head + (size_t)chain.size() * PATH_MAX_CHILDREN * K * 16 * sizeof(double) +
                        2 * (size_t)std::max<int64_t>(max_units, 1) * 4 * sizeof(double);
    ok = ok && smem <= 220 * 1024;
    if (ok) {
      LevelLaunch ll{w_off, p_off, 0, 0, true, 0, 0, 9, 0, 0};
      const int64_t bytes0 = alg_bytes, flops0 = flops;
      for (size_t i = 0; i < chain.size(); ++i) {
        const int n = chain[i];
        work[w_off + ll.n_work] = n;
        int slot = -1;
        if (i > 0)
          for (size_t c = 0; c < e->children[n].size(); ++c)
            if (e->children[n][c] == chain[i - 1]) slot = (int)c;
        dep[w_off + ll.n_work] = slot;
        ++ll.n_work;
        account(n);
        recompute[n] = 0;
      }
      ll.total_tiles = (int)smem;  // (dynamic shared memory of the launch)
      ll.bytes = alg_bytes - bytes0;
      ll.flops = flops - flops0;
      w_off += ll.n_work;
      launches.push_back(ll);
      path_launch = true;
      path_buf_units = (int)std::max<int64_t>(max_units, 1);
    }
  }
  // ---- launch groups: the dirty nodes of one tree level -- or, with several root subtrees streaming, of one
  // (root subtree, level): a subtree only depends on itself, so its launches go to a stream of their own and
  // ramp up while another subtree's launch drains (group 0 unused; the root is the last group, on stream 0)
  std::vector<std::vector<int>> lv_nodes;
  std::vector<int> lv_stream;
  int n_streams_used = 1;
  {
    bool streams = e->use_streams && !e->rescale_active && wave == nullptr && !path_launch && !chain_launch && !walk_launch && !e->use_flow &&
                   !e->profiling && e->batch_mode == 0 && e->children[e->root].size() >= 2;
    if (streams) {
      // worth it only when at least two subtrees have a level that streams (more units than the cluster takes)
      std::vector<char> big(e->children[e->root].size(), 0);
      for (int n = 0; n < N - 1; ++n)
        if (recompute[n] && e->n_rows[n] * K > SMALL_LEVEL_UNITS) big[e->subtree_of[n]] = 1;
      int n_big = 0;
      for (char b : big) n_big += b;
      streams = n_big >= 2;
    }
    lv_nodes.emplace_back();
    lv_stream.push_back(0);
    if (chain_launch || walk_launch || path_launch) {
      // (every dirty node is in that one launch: nothing left to plan level by level)
    } else if (!streams) {
      for (int l = 1; l <= e->n_levels; ++l) {
        lv_nodes.push_back(e->level_nodes[l]);
        lv_stream.push_back(0);
      }
    } else {
      const int T = (int)e->children[e->root].size();
      n_streams_used = T;
      // subtree-major: a subtree's consecutive small levels can still fuse into one small-levels launch; the
      // streams decouple execution order from issue order
      for (int t = 0; t < T; ++t)
        for (int l = 1; l <= e->n_levels; ++l) {
          std::vector<int> g;
          for (int n : e->level_nodes[l])
            if (n != e->root && e->subtree_of[n] == t) g.push_back(n);
          if (g.empty()) continue;
          lv_nodes.push_back(std::move(g));
          lv_stream.push_back(t);  // subtree 0 stays on the engine's stream
        }
      lv_nodes.push_back(std::vector<int>{e->root});
      lv_stream.push_back(0);
    }
  }
  const int n_groups = (int)lv_nodes.size() - 1;
  const bool by_level = n_streams_used == 1;  // group l == tree level l
  // ---- recomputing launches: runs of consecutive small / medium levels, REC_DEPTH + 1 levels per launch -------
  std::vector<int> rec_phase(std::max(e->n_levels, n_groups) + 2, -1);  // level -> phase (launch) number, -1: the usual launches
  std::vector<int> rec_first;                        // first level of every phase
  if (e->use_rec && by_level && !e->use_dmma && e->use_small && !e->rescale_active && !path_launch && !chain_launch && !walk_launch) {
    std::vector<int64_t> level_units(e->n_levels + 2, 0);
    std::vector<char> level_ok(e->n_levels + 2, 0);
    for (int l = 1; l <= e->n_levels; ++l) {
      bool ok = true;
      int n_dirty_l = 0;
      for (int n : e->level_nodes[l])
        if (recompute[n]) {
          level_units[l] += e->n_rows[n] * K;
          ++n_dirty_l;
          ok = ok && e->children[n].size() <= (size_t)REC_MAX_CHILDREN;
        }
      level_ok[l] = ok && n_dirty_l > 0 && level_units[l] <= e->rec_units;
    }
    int l = 1;
    while (l <= e->n_levels) {
      if (!level_ok[l]) { ++l; continue; }
      int l2 = l;
      bool any_medium = false;
      int n_items = 0;
      while (l2 <= e->n_levels && level_ok[l2]) {
        any_medium = any_medium || level_units[l2] > SMALL_LEVEL_UNITS;
        ++l2;
      }
      // (a run of tiny levels only -- a brca1-sized problem -- stays with the cluster kernel, which also reduces)
      if (any_medium && l2 - l >= 2) {
        for (int a0 = l; a0 < l2;) {
          int a1 = a0;
          n_items = 0;
          int64_t units_sum = 0;
          while (a1 < l2 && a1 - a0 <= REC_DEPTH) {
            int items_l = 0;
            for (int n : e->level_nodes[a1]) items_l += recompute[n] ? 1 : 0;
            if (n_items + items_l > REC_MAX_WORK || units_sum + level_units[a1] > INT32_MAX / 2) break;
            n_items += items_l;
            units_sum += level_units[a1];
            ++a1;
          }
          if (a1 == a0) break;  // (a single level with more work items than a launch stages: the usual launches)
          if (a1 - a0 >= 2) {
            for (int x = a0; x < a1; ++x) rec_phase[x] = (int)rec_first.size();
            rec_first.push_back(a0);
          }
          a0 = a1;
        }
      }
      l = l2;
    }
  }
  int32_t* rec_child = reinterpret_cast<int32_t*>(e->h_block + e->off_rec);
  int open_small_stream = 0;
  for (int l = 1; l <= n_groups; ++l) {
    const int g_stream = lv_stream[l];
    if (g_stream != open_small_stream) {
      open_small = -1;  // (a fused small-levels launch never spans two streams)
      open_small_stream = g_stream;
    }
    if (rec_phase[l] >= 0) {
      if (rec_first[rec_phase[l]] != l) continue;  // emitted with the phase's first level
      open_small = -1;
      close_flow();
      const int phase = rec_phase[l];
      LevelLaunch ll{w_off, p_off, 0, 0, false, 0, 0, 10, 0, 0};
      const int64_t bytes0 = alg_bytes, flops0 = flops;
      std::vector<int> item_of(N, -1);
      int64_t u = 0;
      for (int x = l; x <= e->n_levels && rec_phase[x] == phase; ++x)
        for (int n : e->level_nodes[x]) {
          if (!recompute[n]) continue;
          item_of[n] = ll.n_work;
          work[w_off + ll.n_work] = n;
          prefix[p_off + ll.n_work] = (int32_t)u;
          u += e->n_rows[n] * K;
          for (int c = 0; c < REC_MAX_CHILDREN; ++c) {
            const bool has = c < (int)e->children[n].size();
            // children live on lower levels: numbered before their parents, so item_of is final for them
            rec_child[(w_off + ll.n_work) * REC_MAX_CHILDREN + c] = has ? item_of[e->children[n][c]] : -1;
          }
          ++ll.n_work;
          ll.is_root = ll.is_root || (n == e->root);
          account(n);
        }
      prefix[p_off + ll.n_work] = (int32_t)u;
      ll.total_tiles = (int)u;
      ll.bytes = alg_bytes - bytes0;
      ll.flops = flops - flops0;
      w_off += ll.n_work;
      p_off += ll.n_work + 1;
      launches.push_back(ll);
      continue;
    }
    int64_t units = 0;
    int n_dirty = 0;
    bool other_kinds = false;
    for (int n : lv_nodes[l])
      if (recompute[n]) {
        units += e->n_rows[n] * K;
        ++n_dirty;
        other_kinds = other_kinds || kind_of_node(e, n) != 2;
      }
    if (n_dirty == 0) continue;
    const int64_t small_limit =
        (e->use_grid_small && !e->rescale_active)
            ? std::max<int64_t>(SMALL_LEVEL_UNITS, (int64_t)e->grid_units_per_thread * e->n_sm * SMALL_THREADS)
            : SMALL_LEVEL_UNITS;
    if (e->use_small && !e->use_dmma && units <= small_limit) {
      close_flow();
      // extend (or open) a fused launch covering consecutive small levels
      if (open_small < 0) {
        launches.push_back(LevelLaunch{w_off, p_off, 0, 0, false, 0, 0, 1, n_small_total, 0, 0, g_stream});
        open_small = (int)launches.size() - 1;
      }
      LevelLaunch& g = launches[open_small];
      SmallLevel& sl = small[n_small_total++];
      sl.work_off = w_off - g.work_off;
      sl.prefix_off = p_off - g.prefix_off;
      sl.n_work = 0;
      int64_t u = 0;
      const int64_t bytes0 = alg_bytes, flops0 = flops;
      for (int n : lv_nodes[l]) {
        if (!recompute[n]) continue;
        work[w_off + sl.n_work] = n;
        prefix[p_off + sl.n_work] = (int32_t)u;
        u += e->n_rows[n] * K;
        ++sl.n_work;
        g.is_root = g.is_root || (n == e->root);
        account(n);
      }
      prefix[p_off + sl.n_work] = (int32_t)u;
      sl.total_units = (int32_t)u;
      w_off += sl.n_work;
      p_off += sl.n_work + 1;
      g.n_work += sl.n_work;
      g.n_small += 1;
      g.total_tiles += (int)u;
      g.bytes += alg_bytes - bytes0;
      g.flops += flops - flops0;
      if (u > SMALL_LEVEL_UNITS)  // a medium level: the whole group runs grid-wide, one thread per unit where possible
        g.grid = std::max(g.grid, (int)std::min<int64_t>(e->n_sm, (u + SMALL_THREADS - 1) / SMALL_THREADS));
      // a scaled node is rescaled right after its launch: it must be in the group's LAST level
      if (e->rescale_active)
        for (int n : lv_nodes[l])
          if (recompute[n] && e->is_scaled[n]) open_small = -1;
      continue;
    }
    open_small = -1;
    const std::vector<int>& order = lv_nodes[l];
    // nodes the dataflow kernel cannot take are launched on their own, BEFORE the group that
    // this level's binary nodes open (later levels of the group may depend on them)
    if (other_kinds || !flow_ok) close_flow();
    for (int kind : {0, 3, 6, 8, 2}) {
      if (kind == 2 && flow_ok) {
        for (int n : order) {
          if (!recompute[n] || kind_of_node(e, n) != 2) continue;
          const int64_t t = tiles_of_node(e, n, 2);
          if (open_flow >= 0 && (launches[open_flow].n_work == FLOW_MAXW || flow_tiles + t > INT32_MAX))
            close_flow();
          if (open_flow < 0) {
            launches.push_back(LevelLaunch{w_off, p_off, 0, 0, false, 0, 0, 5, 0, 0});
            open_flow = (int)launches.size() - 1;
            flow_tiles = 0;
            any_flow = true;
          }
          LevelLaunch& g = launches[open_flow];
          for (int c = 0; c < 2; ++c) dep[2 * w_off + c] = flow_item[e->children[n][c]];
          flow_item[n] = g.n_work;
          work[w_off++] = n;
          prefix[p_off++] = (int32_t)flow_tiles;
          flow_tiles += t;
          ++g.n_work;
          g.is_root = g.is_root || (n == e->root);
          const int64_t bytes0 = alg_bytes, flops0 = flops;
          account(n);
          g.bytes += alg_bytes - bytes0;
          g.flops += flops - flops0;
        }
        continue;
      }
      LevelLaunch ll{w_off, p_off, 0, 0, false, 0, 0, kind, 0, 0, 0, g_stream};
      int64_t tiles = 0;
      int64_t bytes0 = alg_bytes, flops0 = flops;
      for (int n : lv_nodes[l]) {
        if (!recompute[n] || kind_of_node(e, n) != kind) continue;
        if (kind >= 2 && ll.n_work == (kind == 8 ? DE_MAXW : (kind == 6 ? DR_MAXW : P4S_MAXW))) {
          // the strip kernel stages at most P4S_MAXW descriptors: close this launch, open another
          ll.bytes = alg_bytes - bytes0;
          ll.flops = flops - flops0;
          prefix[p_off + ll.n_work] = (int32_t)tiles;
          ll.total_tiles = (int)tiles;
          w_off += ll.n_work;
          p_off += ll.n_work + 1;
          launches.push_back(ll);
          ll = LevelLaunch{w_off, p_off, 0, 0, false, 0, 0, kind, 0, 0, 0, g_stream};
          tiles = 0;
          bytes0 = alg_bytes;
          flops0 = flops;
        }
        work[w_off + ll.n_work] = n;
        prefix[p_off + ll.n_work] = (int32_t)tiles;
        tiles += tiles_of_node(e, n, kind);
        if (tiles > INT32_MAX) return fail(C3_ERR_INVALID, "too many tiles in one level");
        ++ll.n_work;
        ll.is_root = ll.is_root || (n == e->root);
        account(n);
      }
      if (ll.n_work == 0) continue;
      ll.bytes = alg_bytes - bytes0;
      ll.flops = flops - flops0;
      prefix[p_off + ll.n_work] = (int32_t)tiles;
      ll.total_tiles = (int)tiles;
      w_off += ll.n_work;
      p_off += ll.n_work + 1;
      launches.push_back(ll);
    }
  }
  close_flow();

  // ---- upload: dirty slots, descriptors, the parameter block -------------------------
  int64_t h2d_bytes = 0;
  for (size_t s = 0; s < e->slots.size(); ++s) {
    Slot& sl = e->slots[s];
    if (!sl.dirty) continue;
    h2d_bytes += e->slot_stride * (int64_t)sizeof(double);
    // the vector stays alive and unmodified until the next set_* call; pageable copies
    // are staged by the runtime before the call returns
    C3_CUDA(cudaMemcpyAsync(e->d_slots + (int64_t)s * e->slot_stride, sl.data.data(),
                            (size_t)e->slot_stride * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    sl.dirty = false;
  }
  if (e->nodes_dirty) {
    C3_CUDA(cudaMemcpyAsync(e->d_nodes, e->h_nodes.data(), N * sizeof(NodeDesc), cudaMemcpyHostToDevice,
                            e->stream));
    e->nodes_dirty = false;
  }
  std::memcpy(e->h_block + e->off_pi, e->pi.data(), MP * sizeof(double));
  std::memcpy(e->h_block + e->off_bprobs, e->bprobs.data(), K * sizeof(double));
  // one contiguous copy of the used part of the block
  const int64_t used = e->off_jobs + (int64_t)(n_eigen + n_pade) * sizeof(ExpmJob);
  if (!chain_fused) h2d_bytes += std::min(used, e->block_bytes);
  int64_t n_launch = 0;
  e->prof.clear();
  e->events_used = 0;
  const CommArgs comm = e->comm;  // (its evaluation number lives in device memory: every rank evaluates in lock-step)
  const bool to_host = blocking || deferred;
  const ResultSink sink{e->d_lnl, device_out, to_host ? e->h_lnl_dev : nullptr, e->h_seq_dev, e->d_seq};

  // small problems whose root sits in the (last) small-levels launch: that launch also reduces
  // (prune4_small_kernel's tail: the same tiles and summation trees, one launch fewer)
  e->last_chain = chain_launch;
  e->last_walk = walk_launch;
  const bool fuse_reduce = path_launch || chain_launch || walk_launch ||
                           (e->use_small_reduce && !launches.empty() && launches.back().kind == 1 &&
                            launches.back().grid == 0 && launches.back().is_root && !e->rescale_active && comm.nranks <= 1 &&
                            e->n_rows[e->root] <= (int64_t)SMALL_REDUCE_MAX_TILES * REDUCE_ROWS_PER_TILE);

  // lean root: the launch that holds the root is a regpipe2 launch of its own, several rate classes, no
  // rescaling pass over lh, nobody has read lh lately
  bool lean = false;
  int root_li = -1;
  for (size_t li = 0; li < launches.size(); ++li)
    if (launches[li].is_root) root_li = (int)li;
  if (e->lean_root && root_dirty && root_li >= 0 && !fuse_reduce && (K == 2 || K == 4) && !e->rescale_active &&
      e->lh_keep == 0 && e->use_strip && e->use_reg2 && !e->use_ring && !e->use_dmma &&
      (launches[root_li].kind == 2 || launches[root_li].kind == 3) && launches[root_li].n_work == 1) {
    // (the ALGORITHMIC bytes of the step keep counting lh[p][K] written once and read once: SURVEY §8d's figure
    //  describes the algorithm, not what this launch happens to move -- 8 (K - 1) B per root row less, twice)
    lean = true;
  }
  e->last_lean = lean;
  if (root_dirty && root_li >= 0) {
    const LevelLaunch& rl = launches[root_li];
    e->root_launch.work_off = rl.work_off;
    e->root_launch.prefix_off = rl.prefix_off;
    e->root_launch.n_work = rl.n_work;
    e->root_launch.total_tiles = rl.total_tiles;
    e->root_launch.kind = rl.kind;
    e->lh_valid = !lean;
  }
  const int64_t reduce_row_bytes = (int64_t)(K + 1) * 8;
  if (e->timeline && e->d_tl == nullptr) {
    C3_CUDA(cudaMalloc(&e->d_tl, 64 * 4 * sizeof(unsigned long long)));
    C3_CUDA(cudaHostAlloc(&e->h_tl_init, 64 * 4 * sizeof(unsigned long long), cudaHostAllocDefault));
    for (int i = 0; i < 64; ++i) {
      e->h_tl_init[4 * i + 0] = e->h_tl_init[4 * i + 1] = e->h_tl_init[4 * i + 2] = ~0ull;
      e->h_tl_init[4 * i + 3] = 0ull;
    }
  }

  // every stream operation of the evaluation (issued directly, or into a graph capture)
  bool capturing = false;
  auto issue = [&]() -> int {
  int tl_n = 0;
  if (e->timeline) {
    C3_CUDA(cudaMemcpyAsync(e->d_tl, e->h_tl_init, 64 * 4 * sizeof(unsigned long long), cudaMemcpyHostToDevice, e->stream));
    e->tl_kinds.clear();
  }
  auto tl_slot = [&](int kind) -> unsigned long long* {
    if (!e->timeline || tl_n >= 64) return nullptr;
    e->tl_kinds.push_back(kind);
    e->tl_launches = tl_n + 1;
    return e->d_tl + 4 * tl_n++;
  };
  if (!chain_fused && !e->skip_block_copy) {  // (the fused chain launch carries its parameters by value)
  C3_CUDA(cudaMemcpyAsync(e->d_block, e->h_block, (size_t)std::min(used, e->block_bytes),
                          cudaMemcpyHostToDevice, e->stream));
  if (!capturing) C3_CUDA(cudaEventRecord(e->block_copied, e->stream));  // the staging buffer is free again
  }
  if (any_flow) C3_CUDA(cudaMemsetAsync(e->d_done, 0, (size_t)w_off * sizeof(unsigned int), e->stream));
  const ExpmJob* d_jobs = reinterpret_cast<const ExpmJob*>(e->d_block + e->off_jobs);
  if (n_eigen > 0 && !chain_fused) {
    const int threads = (M <= 4) ? 32 : (M <= 16 ? 64 : (M <= 32 ? 256 : 1024));
    // sP + exp(roots) + (large M) the staged W / evI planes, complex-capable
    const size_t smem = (size_t)(MP * MP + 2 * M + (M > 8 ? 4 * M * (M | 1) : 0)) * sizeof(double);
    static DeviceOnce cfg;
    if (cfg.need(e->device)) { int rc = set_max_smem(expm_eigen_kernel, (size_t)(64 * 64 + 128 + 4 * 64 * 65) * 8); if (rc) return rc; }
    ProfScope ps(e, 0, 0, 0);
    expm_eigen_kernel<<<n_eigen, threads, smem, e->stream>>>(d_jobs, e->d_slots, e->slot_stride, e->d_edges,
                                                            M, MP, K);
    ++n_launch;
  }
  if (n_pade > 0) {
    const int threads = (M <= 4) ? 32 : (M <= 16 ? 128 : (M <= 32 ? 256 : 1024));
    const size_t smem = (size_t)5 * MP * MP * sizeof(double);
    static DeviceOnce cfg;
    if (cfg.need(e->device)) { int rc = set_max_smem(expm_pade_kernel, (size_t)5 * 64 * 64 * 8); if (rc) return rc; }
    ProfScope ps(e, 1, 0, 0);
    expm_pade_kernel<<<n_pade, threads, smem, e->stream>>>(d_jobs + n_eigen, e->d_slots, e->slot_stride,
                                                          e->d_edges, M, MP, K);
    ++n_launch;
  }
  {
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return fail(C3_ERR_CUDA, std::string("expm launch: ") + cudaGetErrorString(err));
  }

  const int32_t* d_work = reinterpret_cast<const int32_t*>(e->d_block + e->off_work);
  const int32_t* d_prefix = reinterpret_cast<const int32_t*>(e->d_block + e->off_prefix);
  // several root subtrees streaming: fork the subtrees' streams off the engine's stream here (the parameter block
  // and the P matrices are in place), join them before the first launch that runs on the engine's stream again
  // after them (the root)
  const cudaStream_t main_stream = e->stream;
  std::vector<char> forked(n_streams_used, 0);
  struct StreamRestore {
    c3_engine* e;
    cudaStream_t s;
    ~StreamRestore() { e->stream = s; }
  } stream_restore{e, main_stream};
  if (n_streams_used > 1) {
    while ((int)e->aux_streams.size() < n_streams_used - 1) {
      cudaStream_t st;
      C3_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
      e->aux_streams.push_back(st);
      cudaEvent_t ev;
      C3_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      e->aux_done.push_back(ev);
    }
    if (!e->aux_fork) C3_CUDA(cudaEventCreateWithFlags(&e->aux_fork, cudaEventDisableTiming));
    C3_CUDA(cudaEventRecord(e->aux_fork, main_stream));
  }
  bool joined = n_streams_used == 1;
  for (size_t li = 0; li < launches.size(); ++li) {
    const LevelLaunch& ll = launches[li];
    if (n_streams_used > 1) {
      if (ll.stream > 0) {
        const cudaStream_t st = e->aux_streams[ll.stream - 1];
        if (!forked[ll.stream]) {
          C3_CUDA(cudaStreamWaitEvent(st, e->aux_fork, 0));
          forked[ll.stream] = 1;
        }
        e->stream = st;
      } else {
        e->stream = main_stream;
        if (!joined && ll.is_root) {
          for (int t = 1; t < n_streams_used; ++t)
            if (forked[t]) {
              C3_CUDA(cudaEventRecord(e->aux_done[t - 1], e->aux_streams[t - 1]));
              C3_CUDA(cudaStreamWaitEvent(main_stream, e->aux_done[t - 1], 0));
            }
          joined = true;
        }
      }
    }
    PruneArgs a;
    a.nodes = e->d_nodes;
    a.childs = e->d_childs;
    a.work_node = d_work + ll.work_off;
    a.tile_prefix = d_prefix + ll.prefix_off;
    a.n_work = ll.n_work;
    a.total_tiles = ll.total_tiles;
    a.K = K;
    a.M = M;
    a.pi = reinterpret_cast<const double*>(e->d_block + e->off_pi);
    a.lh_out = e->d_lh;
    if (lean && (int)li == root_li) {
      a.mix_out = e->d_mix;
      a.bprobs = reinterpret_cast<const double*>(e->d_block + e->off_bprobs);
    }
    a.tl = tl_slot(ll.is_root ? 100 + ll.kind : ll.kind);
    if (e->use_dyn && li < (size_t)DYN_MAX_LAUNCHES)
      a.strip_ctr = reinterpret_cast<int32_t*>(e->d_block + e->off_ctr) + li;
    ProfScope ps(e, (ll.kind == 1 || ll.kind == 9 || ll.kind == 10 || ll.kind == 11 || ll.kind == 12) ? 5 : (ll.is_root ? 3 : 2), ll.bytes, ll.flops);
    if (ll.kind == 1) {
      const SmallLevel* d_small = reinterpret_cast<const SmallLevel*>(e->d_block + e->off_small) + ll.small_off;
      if (ll.grid > 0) {
        // medium levels: every SM, a cooperative launch (the grid barrier needs every CTA resident)
        static DeviceOnce grid_cfg;
        if (grid_cfg.need(e->device)) {
          int rc = set_max_smem(prune4_small_kernel<true>, sizeof(SmallSmem));
          if (rc) return rc;
        }
        C3_CUDA(cudaMemsetAsync(e->d_grid_bar, 0, sizeof(unsigned int), e->stream));
        cudaLaunchConfig_t gcfg = {};
        gcfg.gridDim = dim3(ll.grid);
        gcfg.blockDim = dim3(SMALL_THREADS);
        gcfg.dynamicSmemBytes = sizeof(SmallSmem);
        gcfg.stream = e->stream;
        cudaLaunchAttribute gattr[1];
        gattr[0].id = cudaLaunchAttributeCooperative;
        gattr[0].val.cooperative = 1;
        gcfg.attrs = gattr;
        gcfg.numAttrs = 1;
        cudaError_t gerr = cudaLaunchKernelEx(&gcfg, prune4_small_kernel<true>, a, d_small, ll.n_small, ll.n_work,
                                              ll.n_work + ll.n_small, ReduceTail{}, e->d_grid_bar);
        if (gerr != cudaSuccess) return fail(C3_ERR_CUDA, std::string("medium-level launch: ") + cudaGetErrorString(gerr));
        ++n_launch;
        continue;
      }
      // one cluster; fewer CTAs when there is little work (cluster size must divide the grid)
      int ctas = SMALL_CLUSTER;
      while (ctas > 1 && (int64_t)(ctas / 2) * SMALL_THREADS >= ll.total_tiles / std::max(1, ll.n_small)) ctas /= 2;
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(ctas);
      cfg.blockDim = dim3(SMALL_THREADS);
      cfg.dynamicSmemBytes = sizeof(SmallSmem);
      static DeviceOnce small_cfg;
      if (small_cfg.need(e->device)) {
        int rc = set_max_smem(prune4_small_kernel<false>, sizeof(SmallSmem));
        if (rc) return rc;
      }
      cfg.stream = e->stream;
      cudaLaunchAttribute attr[2];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = ctas;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[1].val.programmaticStreamSerializationAllowed = e->use_pdl ? 1 : 0;
      cfg.attrs = attr;
      cfg.numAttrs = 2;
      ReduceTail tail{};
      if (fuse_reduce && li + 1 == launches.size()) {
        tail.bprobs = reinterpret_cast<const double*>(e->d_block + e->off_bprobs);
        tail.counts = e->d_counts;
        tail.n_rows = e->n_rows[e->root];
        tail.n_tiles = (int)((tail.n_rows + REDUCE_ROWS_PER_TILE - 1) / REDUCE_ROWS_PER_TILE);
        tail.sink = sink;
        tail.tile_sums = e->d_tile_sums;
      }
      cudaError_t err = cudaLaunchKernelEx(&cfg, prune4_small_kernel<false>, a, d_small, ll.n_small, ll.n_work,
                                           ll.n_work + ll.n_small, tail, (unsigned*)nullptr);
      if (err != cudaSuccess) return fail(C3_ERR_CUDA, std::string("small-level launch: ") + cudaGetErrorString(err));
    } else if (ll.kind == 10) {
      static DeviceOnce rec_cfg;
      if (rec_cfg.need(e->device)) {
        int rc = set_max_smem(prune4_rec_kernel, sizeof(RecSmem));
        if (rc) return rc;
      }
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)std::max<int64_t>(1, std::min<int64_t>((ll.total_tiles + REC_THREADS - 1) / REC_THREADS,
                                                                         (int64_t)e->n_sm * 4)));
      cfg.blockDim = dim3(REC_THREADS);
      cfg.dynamicSmemBytes = sizeof(RecSmem);
      cfg.stream = e->stream;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = e->use_pdl ? 1 : 0;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      const int32_t* d_child = reinterpret_cast<const int32_t*>(e->d_block + e->off_rec) + (int64_t)ll.work_off * REC_MAX_CHILDREN;
      cudaError_t err = cudaLaunchKernelEx(&cfg, prune4_rec_kernel, a, d_child);
      if (err != cudaSuccess) return fail(C3_ERR_CUDA, std::string("rec launch: ") + cudaGetErrorString(err));
    } else if (ll.kind == 12) {
      const size_t smem = (size_t)walk_ops * sizeof(WalkOp) + (size_t)walk_slots * 2 * WALK_THREADS * sizeof(double2) +
                          (size_t)walk_nP * K * 16 * sizeof(double) + (size_t)walk_nP * sizeof(double*);
      static DeviceOnce walk_cfg;
      if (walk_cfg.need(e->device)) {
        int rc = set_max_smem(prune4_walk_kernel, (size_t)200 * 1024);
        if (rc) return rc;
      }
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)((ll.total_tiles + WALK_THREADS - 1) / WALK_THREADS));
      cfg.blockDim = dim3(WALK_THREADS);
      cfg.dynamicSmemBytes = smem;
      cfg.stream = e->stream;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = e->use_pdl ? 1 : 0;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      ReduceTail tail{};
      tail.bprobs = reinterpret_cast<const double*>(e->d_block + e->off_bprobs);
      tail.counts = e->d_counts;
      tail.n_rows = e->n_rows[e->root];
      tail.n_tiles = (int)((tail.n_rows + REDUCE_ROWS_PER_TILE - 1) / REDUCE_ROWS_PER_TILE);
      tail.sink = sink;
      const WalkOp* d_ops = reinterpret_cast<const WalkOp*>(e->d_block + e->off_ops);
      cudaError_t err = cudaLaunchKernelEx(&cfg, prune4_walk_kernel, a, d_ops, ll.n_work, walk_slots, walk_nP, tail,
                                           e->d_chain_ctr, e->d_tile_sums);
      if (err != cudaSuccess) return fail(C3_ERR_CUDA, std::string("walk launch: ") + cudaGetErrorString(err));
    } else if (ll.kind == 11) {
      static DeviceOnce chain_cfg;
      if (chain_cfg.need(e->device)) {
        int rc = set_max_smem(prune4_chain_kernel, sizeof(ChainSmem));
        if (rc) return rc;
      }
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)((ll.total_tiles + CHAIN_THREADS - 1) / CHAIN_THREADS));
      cfg.blockDim = dim3(CHAIN_THREADS);
      cfg.dynamicSmemBytes = sizeof(ChainSmem);
      cfg.stream = e->stream;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = e->use_pdl ? 1 : 0;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      ReduceTail tail{};
      tail.bprobs = reinterpret_cast<const double*>(e->d_block + e->off_bprobs);
      tail.counts = e->d_counts;
      tail.n_rows = e->n_rows[e->root];
      tail.n_tiles = (int)((tail.n_rows + REDUCE_ROWS_PER_TILE - 1) / REDUCE_ROWS_PER_TILE);
      tail.sink = sink;
      ChainParams cp{};
      for (int i = 0; i < ll.n_work; ++i) {
        cp.work_node[i] = work[ll.work_off + i];
        cp.dirty_child[i] = dep[ll.work_off + i];
      }
      for (int i = 0; i < 4; ++i) {
        cp.pi[i] = e->pi[i];
        cp.bprobs[i] = i < K ? e->bprobs[i] : 0.0;
      }
      cp.fused = chain_fused ? 1 : 0;
      cp.edge_node = chain_edge_node;
      cp.edge_slot = chain_edge_slot;
      if (chain_fused)
        for (int j = 0; j < K; ++j) {
          cp.q_slot[j] = jobs[j].slot;
          cp.t[j] = jobs[j].t;
        }
      cudaError_t err = cudaLaunchKernelEx(&cfg, prune4_chain_kernel, a, cp, (const int32_t* const*)e->d_node_maps,
                                           (const int32_t* const*)e->d_child_maps, tail, e->d_chain_ctr, e->d_tile_sums,
                                           (const double*)e->d_slots, (int64_t)e->slot_stride, (const EdgeDesc*)e->d_edges);
      if (err != cudaSuccess) return fail(C3_ERR_CUDA, std::string("chain launch: ") + cudaGetErrorString(err));
    } else if (ll.kind == 9) {
      const size_t smem = (size_t)ll.total_tiles;
      static DeviceOnce path_cfg;
      if (path_cfg.need(e->device)) {
        int rc = set_max_smem(prune4_path_kernel<1024, 3>, 220 * 1024);
        if (rc == C3_OK) rc = set_max_smem(prune4_path_kernel<512, 6>, 220 * 1024);
        if (rc == C3_OK) rc = set_max_smem(prune4_path_kernel<256, 11>, 220 * 1024);
        if (rc) return rc;
      }
      const int threads = e->path_variant == 0 ? 1024 : (e->path_variant == 1 ? 512 : 256);
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(1);
      cfg.blockDim = dim3(threads);
      cfg.dynamicSmemBytes = smem;
      cfg.stream = e->stream;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = e->use_pdl ? 1 : 0;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      ReduceTail tail{};
      tail.bprobs = reinterpret_cast<const double*>(e->d_block + e->off_bprobs);
      tail.counts = e->d_counts;
      tail.n_rows = e->n_rows[e->root];
      tail.n_tiles = (int)((tail.n_rows + REDUCE_ROWS_PER_TILE - 1) / REDUCE_ROWS_PER_TILE);
      tail.sink = sink;
      const int32_t* d_dirty = reinterpret_cast<const int32_t*>(e->d_block + e->off_dep) + ll.work_off;
      cudaError_t err;
      if (e->path_variant == 0) err = cudaLaunchKernelEx(&cfg, prune4_path_kernel<1024, 3>, a, d_dirty, path_buf_units, tail);
      else if (e->path_variant == 1) err = cudaLaunchKernelEx(&cfg, prune4_path_kernel<512, 6>, a, d_dirty, path_buf_units, tail);
      else err = cudaLaunchKernelEx(&cfg, prune4_path_kernel<256, 11>, a, d_dirty, path_buf_units, tail);
      if (err != cudaSuccess) return fail(C3_ERR_CUDA, std::string("path launch: ") + cudaGetErrorString(err));
    } else if (ll.kind == 7) {
      C3_CUDA(cudaMemsetAsync(wave->d_done, 0, (size_t)std::max(1, wave->n_chunks) * sizeof(unsigned), e->stream));
      int rc = C3_ERR_INVALID;
      switch (K) {
        case 1: rc = launch_wave<1>(e, a, *wave); break;
        case 2: rc = launch_wave<2>(e, a, *wave); break;
        case 4: rc = launch_wave<4>(e, a, *wave); break;
      }
      if (rc) return rc;
      e->stats.wave_launches += 1;
      e->stats.wave_entries = (int64_t)wave->n_chunks * wave->grid * WAVE_WARPS;
      e->stats.wave_noops = wave->noops;
    } else if (ll.kind == 5) {
      const int32_t* d_dep = reinterpret_cast<const int32_t*>(e->d_block + e->off_dep) + 2 * ll.work_off;
      int rc = C3_ERR_INVALID;
      unsigned int* d_node = e->d_done + ll.work_off;
      switch (K) {
        case 1: rc = launch_flow<1>(e, a, d_dep, d_node); break;
        case 2: rc = launch_flow<2>(e, a, d_dep, d_node); break;
        case 4: rc = launch_flow<4>(e, a, d_dep, d_node); break;
      }
      if (rc) return rc;
    } else {
      int rc = launch_prune(e, a, ll.is_root, ll.kind);
      if (rc) return rc;
    }
    ++n_launch;
    if (e->rescale_active) {
      // scaled nodes of this launch: normalise their rows before the next level reads them
      for (int i = 0; i < ll.n_work; ++i) {
        const int n = work[ll.work_off + i];
        if (!e->is_scaled[n] && n != e->root) continue;
        int rc = launch_rescale(e, n);
        if (rc) return rc;
        ++n_launch;
      }
    }
  }
  e->stream = main_stream;
  if (!joined) {
    for (int t = 1; t < n_streams_used; ++t)
      if (forked[t]) {
        C3_CUDA(cudaEventRecord(e->aux_done[t - 1], e->aux_streams[t - 1]));
        C3_CUDA(cudaStreamWaitEvent(main_stream, e->aux_done[t - 1], 0));
      }
    joined = true;
  }
  if (!fuse_reduce) {
    const int64_t rows = e->n_rows[e->root];
    const int n_tiles = (int)((rows + REDUCE_ROWS_PER_TILE - 1) / REDUCE_ROWS_PER_TILE);
    const int grid = std::min(n_tiles, e->n_sm * 8);
    ProfScope ps(e, 4, rows * reduce_row_bytes, 0);
    ResultSink rsink = sink;
    rsink.tl = tl_slot(4);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(REDUCE_THREADS);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = e->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = e->use_pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const double* a_lh = e->d_lh;
    const double* a_bp = reinterpret_cast<const double*>(e->d_block + e->off_bprobs);
    const double* a_counts = e->d_counts;
    const double* a_mixed = lean ? e->d_mix : nullptr;
    cudaError_t err = cudaLaunchKernelEx(&cfg, lnl_reduce_kernel, a_lh, a_bp, a_counts, (int64_t)rows, K, n_tiles,
                                         e->d_tile_sums, e->d_counter, rsink, comm,
                                         (const int32_t*)(e->rescale_active ? e->scale_rec[e->root].d_exps : nullptr),
                                         a_mixed);
    ++n_launch;
    if (err == cudaSuccess) err = cudaGetLastError();
    if (err != cudaSuccess) return fail(C3_ERR_CUDA, std::string("reduce launch: ") + cudaGetErrorString(err));
  }
  return C3_OK;
  };  // issue

  // ---- direct launches, or capture / replay of this structure's graph ---------------------------
  const auto hp_t1 = std::chrono::steady_clock::now();
  bool graphed = false;
  e->plan = c3_engine::PlanStats{n_eigen + n_pade, n_dirty_nodes, node_rows, alg_bytes, flops, h2d_bytes};
  const bool graph_ok = e->use_graphs && !e->profiling && !any_flow && wave == nullptr &&  // (cooperative launch: not captured)
                        !chain_fused;  // (one launch with by-value parameters: nothing to replay)
  if (e->batch_mode != 0 && !graph_ok) return fail(C3_ERR_INVALID, "engine cannot take part in a batch graph");
  if (graph_ok) {
    std::vector<int32_t> sig;
    sig.reserve(16 + 8 * launches.size() + w_off);
    sig.push_back(n_eigen); sig.push_back(n_pade); sig.push_back((int32_t)std::min(used, e->block_bytes));
    sig.push_back((e->rescale_active ? 1 : 0) | (fuse_reduce ? 2 : 0) | (lean ? 4 : 0) | (walk_slots << 8) | (walk_nP << 16));
    sig.push_back((int32_t)((uintptr_t)device_out & 0xffffffffu)); sig.push_back((int32_t)((uintptr_t)device_out >> 32));
    sig.push_back((blocking || deferred) ? 1 : 0);
    sig.push_back((int32_t)((uintptr_t)e->stream & 0xffffffffu)); sig.push_back((int32_t)((uintptr_t)e->stream >> 32));
    for (const LevelLaunch& ll : launches) {
      sig.push_back(ll.kind); sig.push_back(ll.work_off); sig.push_back(ll.prefix_off); sig.push_back(ll.n_work);
      sig.push_back(ll.total_tiles); sig.push_back(ll.is_root ? 1 : 0); sig.push_back(ll.small_off); sig.push_back(ll.n_small);
      sig.push_back(ll.grid); sig.push_back(ll.stream);
    }
    for (int i = 0; i < w_off; ++i) sig.push_back(work[i]);  // (rescale launches name their nodes)
    uint64_t h = 1469598103934665603ull;
    for (int32_t v : sig) { h ^= (uint32_t)v; h *= 1099511628211ull; }
    e->last_sig_hash = h;
    if (e->batch_mode == 2) return C3_OK;  // dry plan: the parameter block is filled, nothing else happened
    ++e->graph_clock;
    auto it = e->graphs.find(h);
    if (e->batch_mode == 1) {
      // the caller is capturing a batch graph over several engines' streams: issue into it
      capturing = true;
      const int rc = issue();
      capturing = false;
      if (rc) return rc;
      graphed = true;
    } else if (it != e->graphs.end() && it->second.sig == sig) {
      C3_CUDA(cudaGraphLaunch(it->second.exec, e->stream));
      it->second.last_use = e->graph_clock;
      n_launch = it->second.n_launch;
      graphed = true;
    } else if (it == e->graphs.end() && ++e->graph_seen[h] >= e->graph_after) {
      // this structure keeps coming up: capture it (capture + instantiation cost ~100 launches'
      // worth of host time, a replay saves a few microseconds: only frequent structures pay)
      cudaGraph_t graph = nullptr;
      cudaError_t err = cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal);
      if (err == cudaSuccess) {
        capturing = true;
        const int rc = issue();
        capturing = false;
        err = cudaStreamEndCapture(e->stream, &graph);
        if (rc != C3_OK || err != cudaSuccess || graph == nullptr) {
          if (graph) cudaGraphDestroy(graph);
          cudaGetLastError();
          e->use_graphs = false;  // not capturable here (e.g. a stream that is already being captured): plain launches
          graph = nullptr;
          n_launch = 0;
        }
      } else {
        cudaGetLastError();
        e->use_graphs = false;
      }
      if (graph != nullptr) {
        c3_engine::GraphEntry entry;
        err = cudaGraphInstantiate(&entry.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (err == cudaSuccess) {
          if (e->graphs.size() >= 512) {  // evict the least recently used structure
            auto victim = e->graphs.begin();
            for (auto g = e->graphs.begin(); g != e->graphs.end(); ++g)
              if (g->second.last_use < victim->second.last_use) victim = g;
            cudaGraphExecDestroy(victim->second.exec);
            e->graphs.erase(victim);
          }
          entry.sig = sig;
          entry.n_launch = n_launch;
          entry.last_use = e->graph_clock;
          C3_CUDA(cudaGraphLaunch(entry.exec, e->stream));
          e->graphs.emplace(h, std::move(entry));
          graphed = true;
        } else {
          cudaGetLastError();
          e->use_graphs = false;
          n_launch = 0;
        }
      }
    } else if (e->graph_seen.size() > 4096) {
      e->graph_seen.clear();
    }
  }
  if (!graphed) {
    int rc = issue();
    if (rc) return rc;
  } else if (e->batch_mode == 0) {
    C3_CUDA(cudaEventRecord(e->block_copied, e->stream));  // (a replay frees the staging buffer when it is through)
  }
  if (to_host) ++e->seq_issued;  // the launches just issued (or replayed, or captured) publish one more result
  commit_evaluation(e, n_launch, graphed && e->batch_mode == 0);

  if (deferred) {
    e->deferred_pending = true;
    return C3_OK;
  }
  if (blocking) {
    const auto hp_t2 = std::chrono::steady_clock::now();
    int rc = collect_result(e);
    if (rc) return rc;
    if (e->hostprof) {
      const auto hp_t3 = std::chrono::steady_clock::now();
      auto ns = [](auto a, auto b) { return (long long)std::chrono::duration_cast<std::chrono::nanoseconds>(b - a).count(); };
      e->hp_plan_ns += ns(hp_t0, hp_t1);
      e->hp_issue_ns += ns(hp_t1, hp_t2);
      e->hp_wait_ns += ns(hp_t2, hp_t3);
      e->hp_n += 1;
    }
    if (wants_auto_rescale(e)) {
      // some pattern's likelihood left the FP64 range (the reference returns -inf here,
      // evolve/likelihood_tree.py:10): switch to rescaled rows and evaluate again.
      // Sharded runs: every rank sees the same total, so every rank takes this branch.
      rc = activate_rescaling(e);
      if (rc) return rc;
      rc = run_evaluation(e, device_out, blocking, nullptr, deferred);
      if (rc) return rc;
      rc = after_auto_rescale(e);
      if (rc) return rc;
    }
    if (host_out) *host_out = e->cached_lnl;
  } else {
    e->async_pending = true;
    e->have_cached = true;  // value lives in d_lnl; the host copy is refreshed lazily
    e->cached_lnl = NAN;
    e->reduce_dirty = true;  // a later blocking call must re-run the reduction for the host value
  }
  return C3_OK;
}

// after a real stream synchronisation the mapped words hold the last result: take its number as issued
void resync_results(c3_engine* e) {
  cudaStreamSynchronize(e->stream);
  e->seq_issued = *e->h_seq;
}

// Wait for the result of the last evaluation issued towards the host and cache it.  The reduction
// writes the value and then (system fence in between) its sequence number into mapped pinned memory;
// spinning on that word returns within a PCIe write of the kernel's last store, where
// cudaStreamSynchronize adds the driver's completion path (several microseconds -- a tenth of a
// brca1-sized evaluation).  Long evaluations fall back to the blocking synchronisation.
int collect_result(c3_engine* e) {
  cudaError_t err = cudaSuccess;
  bool synced = false;
  if (e->spin_wait && !e->profiling) {
    const auto t0 = std::chrono::steady_clock::now();
    unsigned n = 0;
    // (the device writes value, sequence number and check word without a fence between them: the triple counts
    //  only when it is consistent -- publish_result)
    auto arrived = [&]() {
      if (e->h_seq[0] != e->seq_issued) return false;
      unsigned long long bits;
      const double v = *reinterpret_cast<volatile double*>(e->h_lnl);
      std::memcpy(&bits, &v, sizeof bits);
      return e->h_seq[1] == (bits ^ e->seq_issued) && e->h_seq[0] == e->seq_issued;
    };
    while (!arrived()) {
      if ((++n & 255u) == 0 &&
          std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(400)) {
        err = cudaStreamSynchronize(e->stream);
        synced = true;
        break;
      }
#if defined(__x86_64__) || defined(__i386__)
      __builtin_ia32_pause();
#endif
    }
    std::atomic_thread_fence(std::memory_order_acquire);
  } else {
    err = cudaStreamSynchronize(e->stream);
    synced = true;
  }
  if (err != cudaSuccess) {
    e->have_cached = false;
    c3_invalidate(e);
    cudaGetLastError();
    e->seq_issued = *e->h_seq;
    return fail(C3_ERR_CUDA, std::string("evaluation failed: ") + cudaGetErrorString(err));
  }
  if (synced) e->seq_issued = *e->h_seq;  // stream order: the words hold the last result whatever its number
  for (auto& r : e->prof) cudaEventElapsedTime(&r.ms, r.a, r.b);
  e->async_pending = false;  // stream order: everything issued before the reduction is through as well
  e->block_free = true;      // ... the copy of the parameter block included
  e->cached_lnl = *e->h_lnl;
  e->have_cached = true;
  unsigned long long bits;
  std::memcpy(&bits, &e->cached_lnl, sizeof bits);
  if (bits == C3_TIMEOUT_NAN_BITS) {
    // the fused all-reduce gave up waiting for a peer: the ranks are out of step, no value to return
    e->have_cached = false;
    c3_invalidate(e);
    return fail(C3_ERR_CUDA, "fused all-reduce: a peer rank did not deliver its shard value within 30 s");
  }
  if (std::isfinite(e->cached_lnl)) e->auto_futile = false;
  return C3_OK;
}

// AUTO rescaling switches on for a log-likelihood of -inf only: NaN (bad parameters) and +inf are the
// caller's business (the optimiser treats them as out of bounds), as they are with the reference
bool wants_auto_rescale(const c3_engine* e) {
  return e->rescale_mode == C3_RESCALE_AUTO && !e->rescale_active && !e->auto_futile &&
         std::isinf(e->cached_lnl) && e->cached_lnl < 0;
}

// ... and stays on only if it helped: a pattern whose likelihood is truly 0 (an impossible fixed motif,
// a zero in P) is -inf with and without rescaling, which would otherwise cost 5-38 % for good
int after_auto_rescale(c3_engine* e) {
  if (e->rescale_active && std::isinf(e->cached_lnl) && e->cached_lnl < 0) {
    const double keep = e->cached_lnl;
    deactivate_rescaling(e);  // (everything is recomputed, plainly, by the next evaluation)
    e->auto_futile = true;
    e->cached_lnl = keep;
  }
  return C3_OK;
}

int finish_deferred(c3_engine* e, double* host_out) {
  if (e->deferred_pending) {
    e->deferred_pending = false;
    int rc = collect_result(e);
    if (rc) return rc;
    if (wants_auto_rescale(e)) {
      rc = activate_rescaling(e);
      if (rc) return rc;
      rc = run_evaluation(e, nullptr, true, nullptr);
      if (rc) return rc;
      rc = after_auto_rescale(e);
      if (rc) return rc;
    }
  }
  if (host_out) *host_out = e->cached_lnl;
  return C3_OK;
}

struct BatchCopy {
  const int4* src;  // an engine's pinned parameter staging buffer (device alias)
  int4* dst;        // its device copy
  int64_t n16;      // 16-byte words
};
__global__ void __launch_bounds__(256) batch_copy_kernel(const BatchCopy* __restrict__ copies) {
  const BatchCopy c = copies[blockIdx.x];
  for (int64_t i = threadIdx.x; i < c.n16; i += blockDim.x) c.dst[i] = c.src[i];
}
// C3_BATCH_ONE_COPY=0: one copy node per engine inside a batch graph (the first version)
const bool g_batch_one_copy = [] {
  const char* env = getenv("C3_BATCH_ONE_COPY");
  return env == nullptr || atoi(env) != 0;
}();

// ---- c3_evaluate_many as ONE graph ---------------------------------------------------------------
// Small alignments are bound by the host's launch calls (~14 us per engine even with per-engine
// graphs).  When the same set of engines keeps being evaluated with the same structures (bootstraps,
// batches of genes: always the full sweep), the whole batch is captured once -- every engine's
// operations on its own stream, forked from and joined to the first engine's stream -- and replayed
// with one cudaGraphLaunch; per evaluation the host only plans (fills every engine's parameter block).
// *used = 0: not applicable this time, the caller takes the one-launch-sequence-per-engine path.
int evaluate_many_batched(c3_engine** engines, int n, double* lnL, int* used) {
  *used = 0;
  if (n < 2) return C3_OK;
  std::lock_guard<std::recursive_mutex> lock(g_batch_mutex);
  uint64_t key = 1469598103934665603ull;
  for (int i = 0; i < n; ++i) {
    c3_engine* e = engines[i];
    if (!e->finalized || !e->use_graphs || e->profiling || e->comm.nranks > 1 || e->use_flow || !e->own_stream ||
        e->device != engines[0]->device || e->nodes_dirty || !e->have_pi)
      return C3_OK;
    for (const Slot& s : e->slots)
      if (s.dirty) return C3_OK;  // rate matrices changed: their upload is not part of the graph
    // a pending change of the rescaling mode is applied by the per-engine path (it drops every
    // batch graph: nothing below may hold a reference into g_batches across it)
    if ((e->rescale_mode == C3_RESCALE_ON) != e->rescale_active && e->rescale_mode != C3_RESCALE_AUTO) return C3_OK;
    for (int j = 0; j < i; ++j)
      if (engines[j] == e) return C3_OK;
    key ^= (uint64_t)(uintptr_t)e;
    key *= 1099511628211ull;
  }
  BatchGraph& bg = g_batches[key];
  if (bg.engines.empty()) bg.engines.assign(engines, engines + n);
  if (bg.engines.size() != (size_t)n || !std::equal(engines, engines + n, bg.engines.begin())) return C3_OK;  // (hash collision)
  C3_CUDA(cudaSetDevice(engines[0]->device));
  if (bg.exec == nullptr) {
    if (++bg.seen < engines[0]->graph_after) return C3_OK;
    // capture: every engine issues into the capture on its own stream
    if (g_fork == nullptr) C3_CUDA(cudaEventCreateWithFlags(&g_fork, cudaEventDisableTiming));
    while ((int)g_joins.size() < n) {
      cudaEvent_t ev;
      C3_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      g_joins.push_back(ev);
    }
    cudaStream_t s0 = engines[0]->stream;
    // the parameter blocks of the whole batch by ONE kernel at the head of the graph (it reads every engine's
    // pinned staging buffer directly and writes the device copy): n separate host-to-device copy nodes of a few KB
    // each serialise on the copy engine -- a third of a 64-engine brca1 batch
    bool one_copy = g_batch_one_copy;
    if (one_copy && bg.d_copies == nullptr) {
      std::vector<BatchCopy> copies(n);
      for (int i = 0; i < n; ++i) {
        void* dev_alias = nullptr;
        if (cudaHostGetDevicePointer(&dev_alias, engines[i]->h_block, 0) != cudaSuccess) {
          cudaGetLastError();
          one_copy = false;
          break;
        }
        copies[i] = BatchCopy{reinterpret_cast<const int4*>(dev_alias), reinterpret_cast<int4*>(engines[i]->d_block),
                              engines[i]->block_bytes / 16};
      }
      if (one_copy) {
        C3_CUDA(cudaMalloc(&bg.d_copies, (size_t)n * sizeof(BatchCopy)));
        C3_CUDA(cudaMemcpy(bg.d_copies, copies.data(), (size_t)n * sizeof(BatchCopy), cudaMemcpyHostToDevice));
      }
    }
    if (cudaStreamBeginCapture(s0, cudaStreamCaptureModeRelaxed) != cudaSuccess) {
      cudaGetLastError();
      bg.seen = -(1 << 30);  // never again for this set of engines
      return C3_OK;
    }
    int rc = C3_OK;
    bool skipped = false;
    if (one_copy) batch_copy_kernel<<<n, 256, 0, s0>>>(reinterpret_cast<const BatchCopy*>(bg.d_copies));
    cudaError_t err = cudaEventRecord(g_fork, s0);
    std::vector<uint64_t> hashes(n);
    std::vector<int64_t> n_launch(n);
    for (int i = 0; i < n && rc == C3_OK && err == cudaSuccess; ++i) {
      c3_engine* e = engines[i];
      if (e->stream != s0) err = cudaStreamWaitEvent(e->stream, g_fork, 0);
      if (err != cudaSuccess) break;
      e->batch_mode = 1;
      e->skip_block_copy = one_copy;
      rc = run_evaluation(e, nullptr, false, nullptr, true);
      e->skip_block_copy = false;
      e->batch_mode = 0;
      skipped = skipped || e->last_skipped;
      hashes[i] = e->last_sig_hash;
      n_launch[i] = e->stats.last_launches;
      if (e->stream != s0 && rc == C3_OK) {
        err = cudaEventRecord(g_joins[i], e->stream);
        if (err == cudaSuccess) err = cudaStreamWaitEvent(s0, g_joins[i], 0);
      }
    }
    cudaGraph_t graph = nullptr;
    const cudaError_t end_err = cudaStreamEndCapture(s0, &graph);
    if (rc != C3_OK || err != cudaSuccess || end_err != cudaSuccess || graph == nullptr || skipped) {
      // not capturable (or an engine had nothing to do): everything is evaluated again from scratch, plainly
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      bg.seen = skipped ? 0 : -(1 << 30);
      for (int i = 0; i < n; ++i) {
        engines[i]->deferred_pending = false;
        c3_invalidate(engines[i]);
        resync_results(engines[i]);  // (what was captured never ran)
      }
      return C3_OK;
    }
    err = cudaGraphInstantiate(&bg.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (err != cudaSuccess) {
      cudaGetLastError();
      bg.exec = nullptr;
      bg.seen = -(1 << 30);
      for (int i = 0; i < n; ++i) {
        engines[i]->deferred_pending = false;
        c3_invalidate(engines[i]);
        resync_results(engines[i]);
      }
      return C3_OK;
    }
    bg.hashes = hashes;
    bg.n_launch = n_launch;
  } else {
    // replay: plan every engine (fills its parameter block), check that the structures are the captured ones
    bool same = true;
    for (int i = 0; i < n && same; ++i) {
      c3_engine* e = engines[i];
      e->batch_mode = 2;
      const int rc = run_evaluation(e, nullptr, false, nullptr, true);
      e->batch_mode = 0;
      if (rc != C3_OK) return rc;
      same = !e->last_skipped && e->last_sig_hash == bg.hashes[i];
    }
    if (!same) return C3_OK;  // e.g. a single-branch update this time: the per-engine path plans again
    for (int i = 0; i < n; ++i) {
      commit_evaluation(engines[i], bg.n_launch[i], true);
      ++engines[i]->seq_issued;
      engines[i]->deferred_pending = true;
    }
  }
  cudaStream_t s0 = engines[0]->stream;
  // the graph runs from s0: asynchronous work still pending on another engine's own stream
  // (c3_evaluate_device) must be through before the replay touches that engine's buffers
  for (int i = 0; i < n; ++i)
    if (engines[i]->async_pending) {
      C3_CUDA(cudaStreamSynchronize(engines[i]->stream));
      engines[i]->async_pending = false;
    }
  C3_CUDA(cudaGraphLaunch(bg.exec, s0));
  // (no block_copied records: the call waits for the whole graph below, so every staging buffer is free again
  // by the time anybody plans the next evaluation)
  cudaError_t err = cudaStreamSynchronize(s0);
  if (err != cudaSuccess) {
    for (int i = 0; i < n; ++i) {
      engines[i]->deferred_pending = false;
      engines[i]->have_cached = false;
      c3_invalidate(engines[i]);
      engines[i]->seq_issued = *engines[i]->h_seq;
    }
    return fail(C3_ERR_CUDA, std::string("batched evaluation failed: ") + cudaGetErrorString(err));
  }
  int rc = C3_OK;
  for (int i = 0; i < n; ++i) {
    c3_engine* e = engines[i];
    e->deferred_pending = false;
    e->async_pending = false;
    e->seq_issued = *e->h_seq;  // (the whole graph is through: the mapped words hold this batch's results)
    e->cached_lnl = *e->h_lnl;
    e->have_cached = true;
    if (std::isfinite(e->cached_lnl)) e->auto_futile = false;
    lnL[i] = e->cached_lnl;
  }
  for (int i = 0; i < n; ++i) {
    c3_engine* e = engines[i];
    if (wants_auto_rescale(e)) {
      int rc2 = activate_rescaling(e);  // (drops the batch graphs: the structures change)
      if (rc2 == C3_OK) rc2 = run_evaluation(e, nullptr, true, nullptr);
      if (rc2 == C3_OK) rc2 = after_auto_rescale(e);
      lnL[i] = e->cached_lnl;
      if (rc == C3_OK) rc = rc2;
      if (rc2 != C3_OK) break;
    }
  }
  *used = 1;
  return rc;
}

// ---- per-pattern underflow rescaling: set-up -------------------------------------------------
void deactivate_rescaling(c3_engine* e) {
  if (!e->rescale_active) return;
  cudaStreamSynchronize(e->stream);
  clear_graphs(e);
  if (e->d_scale_pool) cudaFree(e->d_scale_pool);
  e->d_scale_pool = nullptr;
  e->stats.device_bytes -= e->scale_pool_bytes;
  e->scale_pool_bytes = 0;
  e->rescale_active = false;
  e->n_scaled = 0;
  std::fill(e->node_dirty.begin(), e->node_dirty.end(), 1);  // stored rows carry the old scaling
  e->reduce_dirty = true;
  e->have_cached = false;
}

int activate_rescaling(c3_engine* e) {
  if (e->rescale_active) return C3_OK;
  C3_REQUIRE(e->finalized, "call c3_finalize first");
  const int N = e->n_nodes, root = e->root;
  // scaled nodes: post-order, a subtree is cut off once it has gathered `threshold` unscaled tips
  e->is_scaled.assign(N, 0);
  std::vector<int> unscaled(N, 0);
  int n_scaled = 0;
  for (int n = 0; n < N; ++n) {
    if (e->children[n].empty()) { unscaled[n] = 1; continue; }
    int u = 0;
    for (int c : e->children[n]) u += e->is_scaled[c] ? 0 : unscaled[c];
    unscaled[n] = u;
    if (n != root && u >= e->rescale_threshold) { e->is_scaled[n] = 1; ++n_scaled; }
  }
  // nearest scaled descendants of every scaled node and of the root
  std::vector<std::vector<int>> near(N);
  int64_t bytes = 0, tmp_rows = 0;
  auto carve = [&](int64_t b) { const int64_t o = bytes; bytes += align_up(b, 256); return o; };
  std::vector<int64_t> o_exps(N, -1), o_src(N, -1);
  std::vector<std::vector<int64_t>> o_map(N);
  for (int x = 0; x < N; ++x) {
    if (!e->is_scaled[x] && x != root) continue;
    std::vector<int> stack(e->children[x].begin(), e->children[x].end());
    while (!stack.empty()) {
      const int c = stack.back();
      stack.pop_back();
      if (e->is_scaled[c]) near[x].push_back(c);
      else for (int g : e->children[c]) stack.push_back(g);
    }
    std::sort(near[x].begin(), near[x].end());
    o_exps[x] = carve(e->n_rows[x] * (int64_t)sizeof(int32_t));
    o_src[x] = carve((int64_t)std::max<size_t>(1, near[x].size()) * sizeof(ScaleSrc));
    for (size_t i = 0; i < near[x].size(); ++i) o_map[x].push_back(carve(e->n_rows[x] * (int64_t)sizeof(int32_t)));
    if (!near[x].empty()) tmp_rows = std::max(tmp_rows, e->n_rows[x]);
  }
  const int64_t o_tmp = carve(std::max<int64_t>(1, tmp_rows) * sizeof(int32_t));
  C3_CUDA(cudaSetDevice(e->device));
  size_t free_b = 0, total_b = 0;
  C3_CUDA(cudaMemGetInfo(&free_b, &total_b));
  if ((size_t)bytes > free_b) return fail(C3_ERR_MEMORY, "not enough device memory for the rescaling tables");
  C3_CUDA(cudaMalloc(&e->d_scale_pool, (size_t)bytes));
  e->scale_pool_bytes = bytes;
  e->stats.device_bytes += bytes;
  e->scale_rec.assign(N, c3_engine::ScaleRec{});
  int32_t* d_tmp = reinterpret_cast<int32_t*>(e->d_scale_pool + o_tmp);
  for (int x = 0; x < N; ++x) {
    if (o_exps[x] < 0) continue;
    c3_engine::ScaleRec& rec = e->scale_rec[x];
    rec.d_exps = reinterpret_cast<int32_t*>(e->d_scale_pool + o_exps[x]);
    rec.d_src = reinterpret_cast<const ScaleSrc*>(e->d_scale_pool + o_src[x]);
    rec.n_src = (int)near[x].size();
    std::vector<ScaleSrc> h_src(near[x].size());
    const int64_t rows = e->n_rows[x];
    const int grid = (int)std::min<int64_t>((rows + 255) / 256, (int64_t)e->n_sm * 16);
    for (size_t i = 0; i < near[x].size(); ++i) {
      const int d = near[x][i];
      // the path x -> ... -> d: compose the gather tables on the way down
      std::vector<int> path;  // nodes from d up to (excluding) x
      for (int a = d; a != x; a = e->parent[a]) path.push_back(a);
      int32_t* final_map = reinterpret_cast<int32_t*>(e->d_scale_pool + o_map[x][i]);
      const int32_t* cur = nullptr;
      int a = x;
      for (int s = (int)path.size() - 1; s >= 0; --s) {
        const int c = path[s];
        int ci = 0;
        while (e->children[a][ci] != c) ++ci;
        // rows of `a` index rows of `c`; alternate buffers so that the last step lands in final_map
        int32_t* out = (s % 2 == 0) ? final_map : d_tmp;
        compose_map_kernel<<<grid, 256, 0, e->stream>>>(out, cur, e->d_idx[a] + (int64_t)ci * e->n_rows[a], rows);
        cur = out;
        a = c;
      }
      h_src[i].map = final_map;
      h_src[i].exps = reinterpret_cast<const int32_t*>(e->d_scale_pool + o_exps[d]);
    }
    if (!h_src.empty())
      C3_CUDA(cudaMemcpyAsync(e->d_scale_pool + o_src[x], h_src.data(), h_src.size() * sizeof(ScaleSrc),
                              cudaMemcpyHostToDevice, e->stream));
    C3_CUDA(cudaStreamSynchronize(e->stream));  // h_src goes out of scope
  }
  {
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return fail(C3_ERR_CUDA, std::string("rescaling set-up: ") + cudaGetErrorString(err));
  }
  clear_graphs(e);
  e->rescale_active = true;
  e->n_scaled = n_scaled;
  std::fill(e->node_dirty.begin(), e->node_dirty.end(), 1);
  e->reduce_dirty = true;
  e->have_cached = false;
  return C3_OK;
}

// after the launch that computed `node`: scale its rows (a scaled node) / sum the exponents (the root)
int launch_rescale(c3_engine* e, int node) {
  const c3_engine::ScaleRec& rec = e->scale_rec[node];
  const int64_t rows = e->n_rows[node];
  if (node == e->root) {
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((rows + 255) / 256, (int64_t)e->n_sm * 16));
    root_exp_kernel<<<grid, 256, 0, e->stream>>>(rec.d_exps, rows, rec.d_src, rec.n_src);
  } else {
    const int row_len = e->K * e->MP, chunks = row_len / 4;
    int lpr = 1;
    while (lpr < chunks && lpr < 32) lpr *= 2;
    const int64_t warps = (rows + (32 / lpr) - 1) / (32 / lpr);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((warps + 7) / 8, (int64_t)e->n_sm * 8));
    double* r = e->d_rows[node];
    switch (lpr) {
      case 1: rescale_rows_kernel<1><<<grid, 256, 0, e->stream>>>(r, rows, row_len, rec.d_exps, rec.d_src, rec.n_src); break;
      case 2: rescale_rows_kernel<2><<<grid, 256, 0, e->stream>>>(r, rows, row_len, rec.d_exps, rec.d_src, rec.n_src); break;
      case 4: rescale_rows_kernel<4><<<grid, 256, 0, e->stream>>>(r, rows, row_len, rec.d_exps, rec.d_src, rec.n_src); break;
      case 8: rescale_rows_kernel<8><<<grid, 256, 0, e->stream>>>(r, rows, row_len, rec.d_exps, rec.d_src, rec.n_src); break;
      case 16: rescale_rows_kernel<16><<<grid, 256, 0, e->stream>>>(r, rows, row_len, rec.d_exps, rec.d_src, rec.n_src); break;
      default: rescale_rows_kernel<32><<<grid, 256, 0, e->stream>>>(r, rows, row_len, rec.d_exps, rec.d_src, rec.n_src); break;
    }
  }
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return fail(C3_ERR_CUDA, std::string("rescale launch: ") + cudaGetErrorString(err));
  return C3_OK;
}

}  // namespace
