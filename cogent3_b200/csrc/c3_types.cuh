// Shared host/device descriptors of libc3cuda.
#pragma once
#include <cstdint>

namespace c3 {

// One child edge of an internal node, as seen by the pruning kernels.
//   src : the child's rows.  Internal child: its partial likelihoods
//         [rows][K][MP].  Tip child: the pre-multiplied table
//         tip_inner[u][b][:] = P_b . indicator_u  (numpy.inner(tip rows, psub),
//         evolve/likelihood_calculation.py:86 applied to a tip), written by the
//         expm kernels.
//   idx : LikelihoodTreeEdge.indexes[c] narrowed to int32 [parent rows].
//   P   : P[b][i][j] of the edge for an internal child, nullptr when src is
//         already multiplied by P (tips).
struct ChildRef {
  const double* src;
  const int32_t* idx;
  const double* P;
  int64_t src_rows;
};

struct NodeDesc {
  double* out;       // [rows][K][MP]
  int64_t n_rows;    // P+1
  int32_t child_begin;
  int32_t n_children;
  int32_t fixed_motif;  // -1: none (PartialLikelihoodProductDefnFixedMotif)
  int32_t is_root;
  // edge-major engines (M != 4): the node stores its rows already multiplied by the P of the
  // branch above it; P_own = that P [K][MP][MP], nullptr for the root
  const double* P_own;
};

// per edge (node with a branch above it)
struct EdgeDesc {
  double* P;                // [K][MP][MP]
  const double* tip_table;  // [tip_rows][MP] indicator rows, nullptr for internal
  double* tip_inner;        // [tip_rows][K][MP]
  const int32_t* tip_state; // [tip_rows] the state of a one-hot row, -1 for ambiguity codes / the gap row
  int64_t tip_rows;
};

enum ExpmMode : int32_t {
  EXPM_EIGEN_REAL = 0,
  EXPM_EIGEN_COMPLEX = 1,
  EXPM_PADE = 2,
  EXPM_FIXED = 3,  // P supplied by the host; only the tip table is refreshed
  EXPM_TN93 = 4,   // closed form for F81 / HKY85 / TN93 (calc_TN93_P): slot = pi[4], alpha1, alpha2
};

struct ExpmJob {
  int32_t node;
  int32_t bin;
  int32_t slot;
  int32_t mode;
  double t;
};

struct PruneArgs {
  const NodeDesc* nodes;
  const ChildRef* childs;
  const int32_t* work_node;    // dirty nodes of this level
  const int32_t* tile_prefix;  // [n_work+1] first tile of each work item
  int32_t n_work;
  int32_t total_tiles;
  int32_t K;
  int32_t M;
  const double* pi;  // root only: [MP] (zero padded)
  double* lh_out;    // root only: [rows][K]
  // "lean root" (regpipe2 root launch, K > 1): the launch mixes the rate classes itself and writes
  // mix[p] = sum_b bprob_b lh_b[p] (8 B per row) instead of lh[p][K]; the reduction then reads mix and counts only.
  // lh is recomputed when an accessor asks for it.
  double* mix_out = nullptr;
  const double* bprobs = nullptr;
  // developer timeline (C3_TIMELINE=1): [entry_min, start_min (after the dependency wait), end_min, end_max] in
  // globaltimer ns, per launch
  unsigned long long* tl = nullptr;
  // dynamic strips (prune4_reg2_kernel<..., DYN>): this launch's claim counter, zero at launch
  int32_t* strip_ctr = nullptr;
};

// Site-pattern shards on several GPUs of one node: the reduction's last CTA exchanges the shard
// log-likelihoods through peer memory (NVLink / NVSwitch P2P stores into every rank's mailbox).
constexpr int C3_MAX_RANKS = 16;
struct CommSlot {
  double value;
  unsigned long long epoch;  // evaluation number the value belongs to (written after the value)
};
// The evaluation number lives in DEVICE memory (epoch_ctr, bumped by the reduction's last CTA), not in a
// kernel argument: the launch is then identical from one evaluation to the next and can be replayed from
// a CUDA graph.  Every rank evaluates in lock-step, so the counters agree.
struct CommArgs {
  int32_t nranks;  // 0: not sharded
  int32_t rank;
  unsigned long long* epoch_ctr;
  CommSlot* peer[C3_MAX_RANKS];  // mailbox of every rank: [2 parities][nranks] slots (own included)
};
// a peer that never shows up (30 s): the reduction writes this NaN, the host turns it into an error
constexpr unsigned long long C3_TIMEOUT_NAN_BITS = 0x7ff8dead00000001ull;

// where the scalar of an evaluation goes.  h_lnl / h_seq are MAPPED pinned host words: the value, then
// (after a system fence) a sequence number the host spins on -- no stream synchronisation on the
// blocking path.  seq_ctr is the device-side count of published results (graph replays bump it too).
struct ResultSink {
  double* d_lnl;
  double* d_lnl2;  // optional second device destination (c3_evaluate_device)
  volatile double* h_lnl;
  volatile unsigned long long* h_seq;
  unsigned long long* seq_ctr;
  unsigned long long* tl = nullptr;  // developer timeline slot of the launch that publishes (see PruneArgs::tl)
};

// one tree level inside a fused "small levels" launch
struct SmallLevel {
  int32_t work_off;    // into PruneArgs::work_node
  int32_t prefix_off;  // into PruneArgs::tile_prefix (prefix in (row, bin) units)
  int32_t n_work;
  int32_t total_units;
};

constexpr int SMALL_THREADS = 1024;
constexpr int SMALL_CLUSTER = 8;            // CTAs (SMs) cooperating on the small levels
constexpr int SMALL_LEVEL_UNITS = SMALL_THREADS * SMALL_CLUSTER;  // one unit per thread and level
constexpr int SMALL_MAX_WORK = 192;  // work items whose descriptors are staged in shared memory
constexpr int SMALL_MAX_LEVELS = 64;
constexpr int SMALL_STAGED_CHILDREN = 3;

constexpr int FLOW_MAXW = 256;  // work items (dirty nodes) per dataflow launch

constexpr int PRUNE4_THREADS = 256;
#ifndef PRUNE4_UNROLL
#define PRUNE4_UNROLL 1
#endif
#ifndef PRUNE4_MIN_BLOCKS
#define PRUNE4_MIN_BLOCKS 4
#endif
constexpr int DMMA_THREADS = 256;
constexpr int DMMA_TILE_ROWS = 64;
constexpr int REDUCE_THREADS = 256;
constexpr int REDUCE_ROWS_PER_TILE = 1024;

inline int prune4_rows_per_tile(int K) { return (PRUNE4_THREADS / K) * PRUNE4_UNROLL; }

}  // namespace c3
