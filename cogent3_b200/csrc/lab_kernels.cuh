// LAB kernels: superseded kernel families and measured-and-dropped experiments.  NONE of these is on the default
// path; each is selected by an environment switch (DESIGN.md "Switches") and exists so that the parity tests can run
// the same problems through it (bit-identical results) and so that the measurements of DESIGN.md can be reproduced:
//   prune4_strip_kernel   cp.async ring version of the strip kernel          C3_RING=1      lesson 3-5
//   prune4_reg_kernel     one-strip register pipeline                         C3_REG1=1      lesson 5, 9
//   prune4_flow_kernel    dataflow launch over consecutive big levels         C3_FLOW=1      lesson 8
//   prune4_wave_kernel    wavefront launch, all levels at once                C3_WAVE=1      lesson 16a/b
//   prune4_rec_kernel     barrier-free recomputing launch                     C3_REC=1       lesson 16d
//   prune4_path_kernel    single-branch updates by one CTA                    C3_PATH=1      lesson 16c
//   prune_dmma_kernel, prune_dmma_reg_kernel   parent-major DMMA kernels      C3_EDGE_MAJOR=0 lesson 11-12
//   prune4_walk_kernel    any dirty set of a small problem per root row      C3_WALK=1      lesson 17h
// Included at the end of prune_kernels.cuh (it uses that file's helpers and descriptors).
#pragma once

namespace c3 {

// ---------------------------------------------------------------------------------------
// 4-state kernel, pipelined ("strip") version -- the production path for K in {1,2,4} and
// nodes with 2 or 3 children.
//
// The first version above is latency bound (ncu: long-scoreboard stalls, 0.4 eligible
// warps/cycle, L1 data pipe 70% busy of which 2/3 are LDS of the P matrices): a thread
// waits for its gather index, then for the gathered rows, and block-wide barriers per
// tile serialise the warps.  Here every WARP runs its own software pipeline:
//   * a strip = 128 (row, bin) units = 128/K pattern rows; per child that is 4 KB of
//     gathered rows, fetched with cp.async (LDGSTS, 16 B per lane, no registers held)
//     into a warp-private ring of P4S_STAGES stages in shared memory;
//   * the gather indexes of strip s+STAGES are loaded (coalesced, one int per lane)
//     while strip s+STAGES-1.. are in flight, so neither latency is exposed;
//   * compute: each lane owns one bin and 4 rows of the strip (register tiling: each
//     P element read from shared memory is used for 4 rows), FMA in registers, one
//     256-bit coalesced store per row;
//   * no block barrier anywhere: only __syncwarp around the warp's own ring.
// ---------------------------------------------------------------------------------------
#ifndef P4S_WARPS
#define P4S_WARPS 8
#endif
#ifndef P4S_STAGES_NC2
#define P4S_STAGES_NC2 2
#endif
#ifndef P4S_WARPS_NC3
#define P4S_WARPS_NC3 6
#endif
template <int K, int NC>
struct P4SCfg {
  static constexpr int ROWS = 128 / K;  // pattern rows per strip
  // lo halves (first 16 B) of the 128 units, then -- 64 B further, so that the cp.async
  // writes of one quarter-warp fall into distinct banks -- the hi halves
  static constexpr int HI_OFFSET = 2048 + 64;
  static constexpr int CHILD_BYTES = 4096 + 128;
  static constexpr int STAGES = (NC <= 2) ? P4S_STAGES_NC2 : 2;  // ring depth (shared memory bound)
  static constexpr int WARPS = (NC <= 2) ? P4S_WARPS : P4S_WARPS_NC3;
  static constexpr int STAGE_BYTES = NC * CHILD_BYTES;
  static constexpr int P_DOUBLES = NC * K * P4S_PSTRIDE;  // warp-private copy of the node's P's
  static constexpr int WARP_BYTES = STAGES * STAGE_BYTES + ((P_DOUBLES * 8 + 127) / 128) * 128;
  static constexpr size_t SMEM_BYTES = (size_t)WARPS * WARP_BYTES;
};

template <int K, int NC, bool ROOT>
__global__ void __launch_bounds__(P4SCfg<K, NC>::WARPS * 32, 1) prune4_strip_kernel(const PruneArgs a) {
  using Cfg = P4SCfg<K, NC>;
  constexpr int ROWS = Cfg::ROWS;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int NQ = (ROWS + 31) / 32;  // gather-index registers per lane and child
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ StripSmem<NC> sd;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* wbase = smem_raw + (size_t)warp * Cfg::WARP_BYTES;
  double* sP = reinterpret_cast<double*>(wbase + STAGES * Cfg::STAGE_BYTES);
  const unsigned ring = (unsigned)__cvta_generic_to_shared(wbase);

  pdl_trigger();
  // ---- stage the descriptors (static data: may overlap the previous level's tail) ---------
  for (int i = threadIdx.x; i <= a.n_work; i += blockDim.x) sd.prefix[i] = a.tile_prefix[i];
  for (int i = threadIdx.x; i < a.n_work; i += blockDim.x) {
    const NodeDesc nd = a.nodes[a.work_node[i]];
    WorkDesc<NC>& wd = sd.work[i];
    wd.out = nd.out;
    wd.n_rows = nd.n_rows;
    wd.fixed_motif = nd.fixed_motif;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const ChildRef ch = a.childs[nd.child_begin + c];
      wd.src[c] = ch.src;
      wd.idx[c] = ch.idx;
      wd.P[c] = ch.P;
    }
  }
  __syncthreads();

  const int gw = blockIdx.x * Cfg::WARPS + warp;
  const int GW = gridDim.x * Cfg::WARPS;
  const int total = a.total_tiles;  // strips in this launch
  const int b = lane % K;           // this lane's rate class
  const int slot = lane / K;        // its row slot within a pass (32/K slots)

  // fetch cursor (runs STAGES strips ahead of the compute cursor); strips are numbered
  // across the work list and a warp only moves forward, so the lookup is a short scan
  int f_w = 0;
  auto f_seek = [&](int s) {
    while (s >= sd.prefix[f_w + 1]) ++f_w;
  };
  int32_t f_idx[NC][NQ];
  auto load_idx = [&](int s) {
    const WorkDesc<NC>& wd = sd.work[f_w];
    const int64_t row0 = (int64_t)(s - sd.prefix[f_w]) * ROWS;
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        int64_t p = row0 + q * 32 + lane;
        if (p >= wd.n_rows) p = wd.n_rows - 1;
        f_idx[c][q] = __ldg(wd.idx[c] + p);
      }
  };
  // gathers of one strip into ring stage st: row r of the strip has its index in lane
  // r%32, register r/32
  auto issue = [&](int st) {
    const WorkDesc<NC>& wd = sd.work[f_w];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const unsigned dst0 = ring + st * Cfg::STAGE_BYTES + c * Cfg::CHILD_BYTES;
      const double* src = wd.src[c];
      constexpr int CHUNKS_PER_ROW = K * 2;                // 16-byte chunks per pattern row
      constexpr int ROWS_PER_INSTR = 32 / CHUNKS_PER_ROW;  // rows per warp-wide cp.async
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int r = j * ROWS_PER_INSTR + lane / CHUNKS_PER_ROW;  // row within the strip
        const int k = lane % CHUNKS_PER_ROW;                       // chunk within the row
        const int32_t src_row = __shfl_sync(0xffffffffu, f_idx[c][(j * ROWS_PER_INSTR) / 32], r % 32);
        const int unit = r * K + (k >> 1);  // (row, bin) unit index, 0..127
        const unsigned dst = dst0 + ((k & 1) ? (unsigned)Cfg::HI_OFFSET : 0u) + (unsigned)unit * 16u;
        cp_async16_ca(dst, src + ((int64_t)src_row * K) * 4 + k * 2);
      }
    }
  };

  // ---- prologue: fill the ring ----------------------------------------------------------
  int s_fetch = gw;
  bool waited = false;
#pragma unroll 1
  for (int st = 0; st < STAGES; ++st) {
    if (s_fetch < total) {
      f_seek(s_fetch);
      load_idx(s_fetch);  // gather indexes are static too
      if (!waited) {
        pdl_wait();  // the child rows are the previous launches' output
        waited = true;
      }
      issue(st);
    }
    cp_async_commit();
    s_fetch += GW;
  }
  if (!waited) pdl_wait();
  bool have_next = s_fetch < total;
  if (have_next) {  // idx of the next strip to fetch, loaded one iteration early
    f_seek(s_fetch);
    load_idx(s_fetch);
  }

  // ---- main loop ----------------------------------------------------------------------------
  int c_w = 0, p_w = -1;  // compute cursor; node whose P matrices are in sP
  int st = 0;
#pragma unroll 1
  for (int s = gw; s < total; s += GW) {
    while (s >= sd.prefix[c_w + 1]) ++c_w;
    const WorkDesc<NC>& wd = sd.work[c_w];
    if (c_w != p_w) {
      __syncwarp();
      for (int o = lane; o < NC * K * 16; o += 32) {
        const int c = o / (K * 16);
        const int rem = o - c * (K * 16);  // bin * 16 + element
        const double* P = wd.P[c];
        // pre-multiplied children (tips) use the identity: 1*x + 0*... is exact
        sP[(c * K + (rem >> 4)) * P4S_PSTRIDE + (rem & 15)] =
            P ? P[rem] : (((rem & 15) % 5 == 0) ? 1.0 : 0.0);
      }
      p_w = c_w;
      __syncwarp();
    }
    const int64_t node_row0 = (int64_t)(s - sd.prefix[c_w]) * ROWS;
    const int64_t c_rows = wd.n_rows;
    double* c_out = wd.out;
    const int c_fixed = wd.fixed_motif;
    cp_async_wait_group<STAGES - 1>();
    __syncwarp();

    const unsigned char* stage = wbase + st * Cfg::STAGE_BYTES;
    d4 acc[P4S_U];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      d4 x[P4S_U];
#pragma unroll
      for (int u = 0; u < P4S_U; ++u) {
        const int unit = u * 32 + lane;  // == (u*(32/K) + slot) * K + b
        const double2 lo = *reinterpret_cast<const double2*>(stage + c * Cfg::CHILD_BYTES + unit * 16);
        const double2 hi =
            *reinterpret_cast<const double2*>(stage + c * Cfg::CHILD_BYTES + Cfg::HI_OFFSET + unit * 16);
        x[u].x = lo.x; x[u].y = lo.y; x[u].z = hi.x; x[u].w = hi.y;
      }
      const double* Pc = sP + (c * K + b) * P4S_PSTRIDE;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const double2 p01 = *reinterpret_cast<const double2*>(Pc + 4 * i);
        const double2 p23 = *reinterpret_cast<const double2*>(Pc + 4 * i + 2);
#pragma unroll
        for (int u = 0; u < P4S_U; ++u) {
          const double y = fma(x[u].w, p23.y, fma(x[u].z, p23.x, fma(x[u].y, p01.y, x[u].x * p01.x)));
          double& t = (i == 0) ? acc[u].x : (i == 1) ? acc[u].y : (i == 2) ? acc[u].z : acc[u].w;
          t = (c == 0) ? y : t * y;
        }
      }
    }
    __syncwarp();  // every lane is done reading this stage

    // refill the stage with the strip STAGES ahead (its idx was loaded last iteration)
    if (have_next) issue(st);
    cp_async_commit();
    s_fetch += GW;
    have_next = s_fetch < total;
    if (have_next) {
      f_seek(s_fetch);
      load_idx(s_fetch);
    }

#pragma unroll
    for (int u = 0; u < P4S_U; ++u) {
      const int64_t p = node_row0 + u * (32 / K) + slot;
      if (c_fixed >= 0) {
        if (c_fixed != 0) acc[u].x = 0.0;
        if (c_fixed != 1) acc[u].y = 0.0;
        if (c_fixed != 2) acc[u].z = 0.0;
        if (c_fixed != 3) acc[u].w = 0.0;
      }
      if (p < c_rows) {
        st256(c_out + (p * K + b) * 4, acc[u]);
        if (ROOT) {
          const double lh =
              fma(acc[u].w, a.pi[3], fma(acc[u].z, a.pi[2], fma(acc[u].y, a.pi[1], acc[u].x * a.pi[0])));
          a.lh_out[p * K + b] = lh;
        }
      }
    }
    st = (st + 1 == STAGES) ? 0 : st + 1;
  }
  cp_async_wait_group<0>();
}

// ---------------------------------------------------------------------------------------
// 4-state kernel, register-pipelined version ("regpipe").
//
// ncu on the cp.async ring version: DRAM 56 %, L1 data pipe 74 % -- the pipe that moves the
// gathered rows global -> shared (LDGSTS) and shared -> registers (LDS) plus the P matrices
// is the binding unit, and the ring's shared memory leaves little L1 for the small child
// tables.  Here the gathered rows go straight to registers with 256-bit loads and the
// software pipeline lives in registers: while a warp computes strip s from x_cur it already
// has the 8 (4 rows x 2 children) 32-byte loads of strip s+1 in flight into x_next, and the
// gather indexes of strip s+2.  No shared-memory ring => ~1/3 fewer L1 wavefronts per strip
// and ~200 KB of L1 for the child tables.  Same arithmetic (bit-identical results), same
// strip numbering, same staged descriptors and programmatic dependent launch as the ring
// version.
// Binary nodes now run regpipe2 below (two strips in flight); this kernel remains for nodes with
// three children (the trifurcating root: its row buffers alone are 96 registers per strip).
// ---------------------------------------------------------------------------------------
template <int NC>
struct RegCfg {
  static constexpr int WARPS = (NC <= 2) ? 8 : 6;
};

template <int K, int NC, bool ROOT>
__global__ void __launch_bounds__(RegCfg<NC>::WARPS * 32, 1) prune4_reg_kernel(const PruneArgs a) {
  constexpr int ROWS = 128 / K;
  constexpr int WARPS = RegCfg<NC>::WARPS;
  __shared__ StripSmem<NC> sd;
  __shared__ __align__(16) double sPall[WARPS][NC * K * P4S_PSTRIDE];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* sP = sPall[warp];

  pdl_trigger();
  for (int i = threadIdx.x; i <= a.n_work; i += blockDim.x) sd.prefix[i] = a.tile_prefix[i];
  for (int i = threadIdx.x; i < a.n_work; i += blockDim.x) {
    const NodeDesc nd = a.nodes[a.work_node[i]];
    WorkDesc<NC>& wd = sd.work[i];
    wd.out = nd.out;
    wd.n_rows = nd.n_rows;
    wd.fixed_motif = nd.fixed_motif;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const ChildRef ch = a.childs[nd.child_begin + c];
      wd.src[c] = ch.src;
      wd.idx[c] = ch.idx;
      wd.P[c] = ch.P;
    }
  }
  __syncthreads();

  const int gw = blockIdx.x * WARPS + warp;
  const int GW = gridDim.x * WARPS;
  const int total = a.total_tiles;
  const int b = lane % K;
  const int slot = lane / K;

  // idx cursor runs two strips ahead, gather cursor one strip ahead of the compute cursor
  int i_w = 0, g_w = 0, c_w = 0, p_w = -1;
  int32_t idx_n[NC][P4S_U];  // gather rows of the strip after next
  auto load_idx = [&](int s) {
    while (s >= sd.prefix[i_w + 1]) ++i_w;
    const WorkDesc<NC>& wd = sd.work[i_w];
    const int64_t row0 = (int64_t)(s - sd.prefix[i_w]) * ROWS;
#pragma unroll
    for (int u = 0; u < P4S_U; ++u) {
      int64_t p = row0 + u * (32 / K) + slot;
      if (p >= wd.n_rows) p = wd.n_rows - 1;
#pragma unroll
      for (int c = 0; c < NC; ++c) idx_n[c][u] = __ldg(wd.idx[c] + p);
    }
  };
  d4 x_n[NC][P4S_U];  // rows of the next strip, in flight
  auto gather = [&](int s) {
    while (s >= sd.prefix[g_w + 1]) ++g_w;
    const WorkDesc<NC>& wd = sd.work[g_w];
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
      for (int u = 0; u < P4S_U; ++u)
        x_n[c][u] = ld256_nc(wd.src[c] + ((int64_t)idx_n[c][u] * K + b) * 4);
  };

  pdl_wait();
  if (gw < total) {
    load_idx(gw);
    gather(gw);
  }
  if (gw + GW < total) load_idx(gw + GW);

#pragma unroll 1
  for (int s = gw; s < total; s += GW) {
    while (s >= sd.prefix[c_w + 1]) ++c_w;
    const WorkDesc<NC>& wd = sd.work[c_w];
    if (c_w != p_w) {
      __syncwarp();
      for (int o = lane; o < NC * K * 16; o += 32) {
        const int c = o / (K * 16);
        const int rem = o - c * (K * 16);
        const double* P = wd.P[c];
        sP[(c * K + (rem >> 4)) * P4S_PSTRIDE + (rem & 15)] =
            P ? P[rem] : (((rem & 15) % 5 == 0) ? 1.0 : 0.0);
      }
      p_w = c_w;
      __syncwarp();
    }
    // take over the rows that were in flight, start the next strip's loads
    d4 x[NC][P4S_U];
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
      for (int u = 0; u < P4S_U; ++u) x[c][u] = x_n[c][u];
    if (s + GW < total) gather(s + GW);
    if (s + 2 * GW < total) load_idx(s + 2 * GW);

    d4 acc[P4S_U];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const double* Pc = sP + (c * K + b) * P4S_PSTRIDE;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const double2 p01 = *reinterpret_cast<const double2*>(Pc + 4 * i);
        const double2 p23 = *reinterpret_cast<const double2*>(Pc + 4 * i + 2);
#pragma unroll
        for (int u = 0; u < P4S_U; ++u) {
          const double y =
              fma(x[c][u].w, p23.y, fma(x[c][u].z, p23.x, fma(x[c][u].y, p01.y, x[c][u].x * p01.x)));
          double& t = (i == 0) ? acc[u].x : (i == 1) ? acc[u].y : (i == 2) ? acc[u].z : acc[u].w;
          t = (c == 0) ? y : t * y;
        }
      }
    }
    const int64_t node_row0 = (int64_t)(s - sd.prefix[c_w]) * ROWS;
    const int c_fixed = wd.fixed_motif;
#pragma unroll
    for (int u = 0; u < P4S_U; ++u) {
      const int64_t p = node_row0 + u * (32 / K) + slot;
      if (c_fixed >= 0) {
        if (c_fixed != 0) acc[u].x = 0.0;
        if (c_fixed != 1) acc[u].y = 0.0;
        if (c_fixed != 2) acc[u].z = 0.0;
        if (c_fixed != 3) acc[u].w = 0.0;
      }
      if (p < wd.n_rows) {
        st256(wd.out + (p * K + b) * 4, acc[u]);
        if (ROOT) {
          const double lh =
              fma(acc[u].w, a.pi[3], fma(acc[u].z, a.pi[2], fma(acc[u].y, a.pi[1], acc[u].x * a.pi[0])));
          a.lh_out[p * K + b] = lh;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// 4-state WAVEFRONT kernel ("wave"): every dirty binary node of the sweep in ONE persistent launch,
// all tree levels streaming AT THE SAME TIME.
//
// A launch per level costs ~11 us of ramp, drain and imbalance whatever its size (fit over the C2
// levels: time = 11-15 us + bytes / 6.2 TB/s -- a third of a C2 step for 10 launches), and a
// parent's gather finds its children's rows in HBM although they were written microseconds ago.
// cogent3 numbers patterns in first-seen order, so for every node idx_c[p] <= p (the distinct child
// patterns seen in the first columns are no more than the distinct parent tuples; checked per node
// when the tables arrive): strip k of a node (rows [k R, (k+1) R)) only needs strips <= k of its
// children.  So the strips of ALL nodes are laid out by the host in one static schedule, sorted by
// wave index  w = k + LAG * height(node): a parent runs LAG strips behind its children, every level
// is active at once, and what a parent gathers is what its children stored a few microseconds
// earlier -- still in the 126 MB L2.  DRAM then mostly sees the writes.
//
// Mechanics.  The schedule is cut into CHUNKS of GW entries (GW = resident warps); warp g takes entry
// g of every chunk, chunk after chunk, with the register pipeline of prune4_reg2_kernel (two strips
// in flight, gather indexes two steps ahead).  The host places a strip only into a chunk that lies
// at least `lag` chunks after the chunks of the child strips it depends on (greedy list scheduling;
// a position with nothing eligible is a no-op entry).  Completion is tracked per chunk:
//   * after the stores of its entry of chunk i a warp fences and counts itself in a shared-memory slot
//     of its CTA; the CTA's last warp does  red  chunk_done[i] += 1  (148 atomics per chunk, not 1184);
//   * before it GATHERS rows for its entry of chunk j it makes sure that chunk_done[j - lag] == #CTAs,
//     i.e. EVERY warp has stored its entry of chunk j - lag (and so of all earlier chunks: a warp
//     counts its chunks in order).  The counter is read early (relaxed, at L2) and only looked at
//     after the FMA block; in steady state it is already there.
// A warp only ever waits for strictly earlier chunks, so the warp holding the smallest unfinished
// chunk can always proceed: no deadlock as long as every CTA is resident (cooperative launch).
// Rows of a node are read by no SM before that wait has passed, so no L1 can hold a stale line.
// Arithmetic per row is that of every other 4-state kernel (same FMA order): bit-identical results.
// ---------------------------------------------------------------------------------------
constexpr int WAVE_MAXW = 256;   // work items (dirty nodes) per wave launch
constexpr int WAVE_WARPS = 8;
constexpr int WAVE_RING = 16;    // schedule entries staged per warp (two halves of 8)

struct WaveSmem {
  WorkDesc<2> work[WAVE_MAXW];
};

// schedule entry: .x = work item (-1: nothing to do at this position), .y = strip number inside the node
template <int K>
__global__ void __launch_bounds__(WAVE_WARPS * 32, 1)
prune4_wave_kernel(const PruneArgs a, const int2* __restrict__ sched, int n_chunks, unsigned* __restrict__ chunk_done,
                   int lag, int group) {
  constexpr int NC = 2, U = 4;
  constexpr int ROWS = 32 * U / K;  // pattern rows per strip
  __shared__ WaveSmem sd;
  __shared__ __align__(16) double sPall[WAVE_WARPS][NC * K * P4S_PSTRIDE];
  __shared__ int2 sRing[WAVE_WARPS][WAVE_RING];
  __shared__ unsigned sCount[16];  // warps of this CTA through with chunk i (slot i & 15: the warps of the launch are
                                   // never more than `lag` < 16 chunks apart)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* sP = sPall[warp];
  int2* ring = sRing[warp];

  pdl_trigger();
  for (int i = threadIdx.x; i < a.n_work; i += blockDim.x) {
    const NodeDesc nd = a.nodes[a.work_node[i]];
    WorkDesc<NC>& wd = sd.work[i];
    wd.out = nd.out;
    wd.n_rows = nd.n_rows;
    wd.fixed_motif = nd.fixed_motif;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const ChildRef ch = a.childs[nd.child_begin + c];
      wd.src[c] = ch.src;
      wd.idx[c] = ch.idx;
      wd.P[c] = ch.P;
    }
  }
  // entry g of a chunk goes to warp (g / #CTAs) of CTA (g % #CTAs): a chunk that is only partly filled
  // (the ramps of the schedule) still spreads over every SM
  const int gw = warp * gridDim.x + blockIdx.x;
  const int GW = gridDim.x * WAVE_WARPS;
  const int b = lane % K;
  const int slot = lane / K;
  // this warp's entries of chunks [0, 16) into its ring
  if (lane < WAVE_RING) ring[lane] = (lane < n_chunks) ? __ldg(sched + (int64_t)lane * GW + gw) : make_int2(-1, 0);
  if (threadIdx.x < 16) sCount[threadIdx.x] = 0u;
  const unsigned n_cta = gridDim.x;
  __syncthreads();

  auto entry = [&](int i) -> int2 { return (i < n_chunks) ? ring[i & (WAVE_RING - 1)] : make_int2(-1, 0); };
  auto load_idx = [&](int32_t(&ix)[NC][U], const int2 e) {
    if (e.x < 0) return;
    const WorkDesc<NC>& wd = sd.work[e.x];
    const int64_t row0 = (int64_t)e.y * ROWS;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      int64_t p = row0 + u * (32 / K) + slot;
      if (p >= wd.n_rows) p = wd.n_rows - 1;
#pragma unroll
      for (int c = 0; c < NC; ++c) ix[c][u] = __ldg(wd.idx[c] + p);
    }
  };
  // Completion is signalled per GROUP of `group` consecutive chunks (one fence per warp and group: the fence
  // waits for everything the warp has in flight, loads included -- per chunk it cost more than the chunk):
  // chunk_done[g] counts the CTAs through with chunks [g group, (g + 1) group).
  int known_done = -1;  // every GROUP <= known_done is complete
  bool stored = false;  // this warp has stores that no fence has covered yet
  // rows this launch wrote earlier: plain (coherent) loads after the chunk they depend on is complete
  auto gather = [&](d4(&x)[NC][U], const int32_t(&ix)[NC][U], const int2 e) {
    if (e.x < 0) return;
    const WorkDesc<NC>& wd = sd.work[e.x];
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
      for (int u = 0; u < U; ++u) x[c][u] = ld256(wd.src[c] + ((int64_t)ix[c][u] * K + b) * 4);
  };
  auto wait_chunk = [&](int need, unsigned early) {
    if (need <= known_done) return;
    if (early != n_cta) {
      while (ld_acquire_gpu_u32(chunk_done + need) != n_cta) __nanosleep(64);
    }
    known_done = need;
  };

  int p_w = -1;
  // compute the entry of chunk i from x; refill x with the entry of chunk i + 2 (indexes ix, requested two
  // steps ago); then request the indexes of the entry of chunk i + 4 into ix
  auto step = [&](d4(&x)[NC][U], int32_t(&ix)[NC][U], int i) {
    const int2 cur = entry(i);
    const int2 nxt = entry(i + 2);
    // the counter the refill depends on (the last group that lies entirely at or before chunk i + 2 - lag; the
    // host places a strip accordingly), requested now, looked at after the FMAs
    const int last_ok = i + 2 - lag + 1;  // chunks < last_ok must be complete
    const int need = (last_ok >= group) ? last_ok / group - 1 : -1;
    unsigned early = n_cta;
    if (need > known_done) early = ld_relaxed_gpu_u32(chunk_done + need);
    // every 8 chunks: the entries of chunks [i + 8, i + 16) into the ring half that was just left
    // (the half holding chunks [i, i + 8) is still being read)
    int2 fresh = make_int2(-1, 0);
    const bool reload = (i & 7) == 0 && i + 8 < n_chunks;
    if (reload && lane < 8 && i + 8 + lane < n_chunks) fresh = __ldg(sched + (int64_t)(i + 8 + lane) * GW + gw);
    d4 acc[U];
    if (cur.x >= 0) {
      const WorkDesc<NC>& wd = sd.work[cur.x];
      if (cur.x != p_w) {
        __syncwarp();
        for (int o = lane; o < NC * K * 16; o += 32) {
          const int c = o / (K * 16);
          const int rem = o - c * (K * 16);
          const double* P = wd.P[c];
          sP[(c * K + (rem >> 4)) * P4S_PSTRIDE + (rem & 15)] = P ? P[rem] : (((rem & 15) % 5 == 0) ? 1.0 : 0.0);
        }
        p_w = cur.x;
        __syncwarp();
      }
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const double* Pc = sP + (c * K + b) * P4S_PSTRIDE;
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
          const double2 p01 = *reinterpret_cast<const double2*>(Pc + 4 * k4);
          const double2 p23 = *reinterpret_cast<const double2*>(Pc + 4 * k4 + 2);
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const double y =
                fma(x[c][u].w, p23.y, fma(x[c][u].z, p23.x, fma(x[c][u].y, p01.y, x[c][u].x * p01.x)));
            double& t = (k4 == 0) ? acc[u].x : (k4 == 1) ? acc[u].y : (k4 == 2) ? acc[u].z : acc[u].w;
            t = (c == 0) ? y : t * y;
          }
        }
      }
    }
    // every warp waits, also one with nothing to gather: it keeps the warps of the launch within `lag`
    // chunks of each other (the shared-memory slots above and the host's placement rely on it)
    wait_chunk(need, early);
    gather(x, ix, nxt);
    load_idx(ix, entry(i + 4));
    if (cur.x >= 0) {
      const WorkDesc<NC>& wd = sd.work[cur.x];
      const int64_t node_row0 = (int64_t)cur.y * ROWS;
      const int c_fixed = wd.fixed_motif;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t p = node_row0 + u * (32 / K) + slot;
        if (c_fixed >= 0) {
          if (c_fixed != 0) acc[u].x = 0.0;
          if (c_fixed != 1) acc[u].y = 0.0;
          if (c_fixed != 2) acc[u].z = 0.0;
          if (c_fixed != 3) acc[u].w = 0.0;
        }
        if (p < wd.n_rows) st256(wd.out + (p * K + b) * 4, acc[u]);
      }
      stored = true;
    }
    // this warp's entries of a whole group are stored (or were empty): tell everybody
    if ((i + 1) % group == 0 || i + 1 == n_chunks) {
      const int g = i / group;
      __syncwarp();
      if (lane == 0) {
        if (stored) __threadfence();  // (cumulative over the other lanes' stores through __syncwarp)
        if (atomicAdd(&sCount[g & 15], 1u) == WAVE_WARPS - 1) {
          sCount[g & 15] = 0u;  // the slot is next used for group g + 16
          __threadfence();
          atomicAdd(chunk_done + g, 1u);
        }
      }
      stored = false;
    }
    if (reload) {
      // (all lanes are past their reads of the other half: entry(i + 4) above was the last one of this step
      //  and lies in [i, i + 8))
      __syncwarp();
      if (lane < 8) ring[(i + 8 + lane) & (WAVE_RING - 1)] = fresh;
      __syncwarp();
    }
  };

  d4 xa[NC][U], xb[NC][U];
  int32_t ia[NC][U], ib[NC][U];
  // prologue: entries of chunks 0, 1 into the buffers; indexes of chunks 2, 3 into the index sets
  load_idx(ia, entry(0));  // static data: may overlap the previous launch
  load_idx(ib, entry(1));
  pdl_wait();
  // (the host never places a strip that depends on anything into the first `lag` chunks)
  if (n_chunks > 0) gather(xa, ia, entry(0));
  if (n_chunks > 1) gather(xb, ib, entry(1));
  load_idx(ia, entry(2));
  load_idx(ib, entry(3));

#pragma unroll 1
  for (int i = 0; i < n_chunks; i += 2) {
    step(xa, ia, i);
    if (i + 1 < n_chunks) step(xb, ib, i + 1);
  }
}

// ---------------------------------------------------------------------------------------
// 4-state DATAFLOW kernel: several consecutive tree levels in ONE persistent launch.
//
// Per-level launches cost ~15 us each of ramp-up / tail / launch latency (fit over the C2
// levels: time = 15-20 us + bytes / 6 TB/s), a third of a C2 step.  Here the dirty binary
// nodes of a whole run of levels form one work list in level order; strips are numbered
// across it and dealt to the resident warps round-robin, exactly as in the regpipe kernel.
// Instead of a kernel boundary between levels, every node has a completion counter:
//   * when a warp leaves node n (its next strip belongs to another node) it does
//     fence.acq_rel.gpu + red.relaxed  node_done[n] += (strips of n it stored)  (the fence
//     is cumulative over the other lanes' stores through __syncwarp);
//   * before a warp gathers rows for a strip of node n it makes sure, once per node, that
//     node_done[child] == strips(child) for the children computed in this launch
//     (ld.acquire.gpu poll).
// Deadlock freedom: strips are in level order and every warp visits its strips in
// increasing order; a warp only ever blocks when it ENTERS a node, and by then it has
// signalled everything it stored; inside a node (whose children are complete) it never
// blocks.  So the owner of the smallest unfinished strip can always proceed.  The one trap
// -- a warp PREFETCHING strip s+GW must not block on a dependency before it has stored and
// signalled strip s -- is avoided by prefetching early only when the dependency is already
// satisfied, and otherwise gathering after the signal.  All CTAs are resident (grid <= SM
// count, one CTA per SM).
// Measured and dropped: a release per strip (the MEMBAR.GPU is a store round trip per strip:
// 20 % slower), and strip-granular dependencies through "super-round" counters so that level
// l+1 streams in right behind level l (cogent3's first-seen pattern numbering guarantees
// idx[p] <= p): 4-10 % slower than node granularity at 4..32 rounds per signal -- the
// kernel is issue bound and the level transitions were not where the time goes.  Walking
// odd levels downwards so that parents first gather what is still in L2: no change.
// ---------------------------------------------------------------------------------------
struct FlowWork {
  double* out;
  int64_t n_rows;
  const double* src[2];
  const int32_t* idx[2];
  const double* P[2];
  int32_t dep[2];  // work item producing child c in this launch, -1: nothing to wait for
  int32_t fixed_motif;
  int32_t is_root;
  int32_t n_strips;
  int32_t pad;
};
struct FlowSmem {
  FlowWork work[FLOW_MAXW];
  int32_t prefix[FLOW_MAXW + 1];
};

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void red_relaxed_gpu_add(unsigned* p, unsigned v) {
  asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int K>
__global__ void __launch_bounds__(256, 1)
prune4_flow_kernel(const PruneArgs a, const int32_t* __restrict__ dep, unsigned* __restrict__ node_done) {
  constexpr int ROWS = 128 / K;
  constexpr int WARPS = 8;
  constexpr int NC = 2;
  extern __shared__ __align__(16) unsigned char flow_raw[];
  FlowSmem& sd = *reinterpret_cast<FlowSmem*>(flow_raw);
  __shared__ __align__(16) double sPall[WARPS][NC * K * P4S_PSTRIDE];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* sP = sPall[warp];

  pdl_trigger();
  for (int i = threadIdx.x; i <= a.n_work; i += blockDim.x) sd.prefix[i] = a.tile_prefix[i];
  for (int i = threadIdx.x; i < a.n_work; i += blockDim.x) {
    const NodeDesc nd = a.nodes[a.work_node[i]];
    FlowWork& wd = sd.work[i];
    wd.out = nd.out;
    wd.n_rows = nd.n_rows;
    wd.fixed_motif = nd.fixed_motif;
    wd.is_root = nd.is_root;
    wd.n_strips = a.tile_prefix[i + 1] - a.tile_prefix[i];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const ChildRef ch = a.childs[nd.child_begin + c];
      wd.src[c] = ch.src;
      wd.idx[c] = ch.idx;
      wd.P[c] = ch.P;
      wd.dep[c] = dep[2 * i + c];
    }
  }
  __syncthreads();

  const int gw = blockIdx.x * WARPS + warp;
  const int GW = gridDim.x * WARPS;
  const int total = a.total_tiles;
  const int b = lane % K;
  const int slot = lane / K;

  int i_w = 0, g_w = 0, c_w = 0, p_w = -1;
  int ready_w = -1;  // last work item whose children are known to be complete
  // is every child of work item w complete?  Uniform across the warp (same-address polls).
  auto deps_done = [&](int w) -> bool {
    bool ok = true;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int d = sd.work[w].dep[c];
      if (d >= 0 && ld_acquire_gpu(node_done + d) < (unsigned)sd.work[d].n_strips) ok = false;
    }
    __syncwarp();
    return ok;
  };
  auto strip_row0 = [&](int s, int w) -> int64_t { return (int64_t)(s - sd.prefix[w]) * ROWS; };
  int32_t idx_n[NC][P4S_U];
  auto load_idx = [&](int s) {
    while (s >= sd.prefix[i_w + 1]) ++i_w;
    const FlowWork& wd = sd.work[i_w];
    const int64_t row0 = strip_row0(s, i_w);
#pragma unroll
    for (int u = 0; u < P4S_U; ++u) {
      int64_t p = row0 + u * (32 / K) + slot;
      if (p >= wd.n_rows) p = wd.n_rows - 1;
#pragma unroll
      for (int c = 0; c < NC; ++c) idx_n[c][u] = __ldg(wd.idx[c] + p);
    }
  };
  d4 x_n[NC][P4S_U];
  // rows written earlier in this launch: plain (coherent) loads, never the .nc path
  auto gather = [&]() {
    const FlowWork& wd = sd.work[g_w];
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
      for (int u = 0; u < P4S_U; ++u)
        x_n[c][u] = ld256(wd.src[c] + ((int64_t)idx_n[c][u] * K + b) * 4);
  };
  // make strip s the gather target; returns whether its dependencies are satisfied
  auto seek_gather = [&](int s, bool block) -> bool {
    while (s >= sd.prefix[g_w + 1]) ++g_w;
    if (g_w == ready_w) return true;
    bool ok = deps_done(g_w);
    while (!ok && block) {
      __nanosleep(100);
      ok = deps_done(g_w);
    }
    if (ok) ready_w = g_w;
    return ok;
  };

  if (gw < total) load_idx(gw);  // gather indexes are static: may overlap the previous launch
  pdl_wait();
  bool pending = false;   // the next strip's gather still has to be issued (dependency not ready)
  int n_unsignalled = 0;  // strips of the current node stored but not yet added to node_done[]
  if (gw < total) {
    seek_gather(gw, true);
    gather();
  }
  if (gw + GW < total) load_idx(gw + GW);

#pragma unroll 1
  for (int s = gw; s < total; s += GW) {
    while (s >= sd.prefix[c_w + 1]) ++c_w;
    const FlowWork& wd = sd.work[c_w];
    if (c_w != p_w) {
      __syncwarp();
      for (int o = lane; o < NC * K * 16; o += 32) {
        const int c = o / (K * 16);
        const int rem = o - c * (K * 16);
        const double* P = wd.P[c];
        sP[(c * K + (rem >> 4)) * P4S_PSTRIDE + (rem & 15)] =
            P ? P[rem] : (((rem & 15) % 5 == 0) ? 1.0 : 0.0);
      }
      p_w = c_w;
      __syncwarp();
    }
    d4 x[NC][P4S_U];
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
      for (int u = 0; u < P4S_U; ++u) x[c][u] = x_n[c][u];
    // prefetch the next strip now if its node's children are complete, else after the signal
    pending = false;
    const bool more = s + GW < total;
    if (more) {
      if (seek_gather(s + GW, false)) gather();
      else pending = true;
    }
    const bool more_idx = s + 2 * GW < total;
    if (more_idx && !pending) load_idx(s + 2 * GW);

    d4 acc[P4S_U];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const double* Pc = sP + (c * K + b) * P4S_PSTRIDE;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const double2 p01 = *reinterpret_cast<const double2*>(Pc + 4 * i);
        const double2 p23 = *reinterpret_cast<const double2*>(Pc + 4 * i + 2);
#pragma unroll
        for (int u = 0; u < P4S_U; ++u) {
          const double y =
              fma(x[c][u].w, p23.y, fma(x[c][u].z, p23.x, fma(x[c][u].y, p01.y, x[c][u].x * p01.x)));
          double& t = (i == 0) ? acc[u].x : (i == 1) ? acc[u].y : (i == 2) ? acc[u].z : acc[u].w;
          t = (c == 0) ? y : t * y;
        }
      }
    }
    const int64_t node_row0 = strip_row0(s, c_w);
    const int c_fixed = wd.fixed_motif;
#pragma unroll
    for (int u = 0; u < P4S_U; ++u) {
      const int64_t p = node_row0 + u * (32 / K) + slot;
      if (c_fixed >= 0) {
        if (c_fixed != 0) acc[u].x = 0.0;
        if (c_fixed != 1) acc[u].y = 0.0;
        if (c_fixed != 2) acc[u].z = 0.0;
        if (c_fixed != 3) acc[u].w = 0.0;
      }
      if (p < wd.n_rows) {
        st256(wd.out + (p * K + b) * 4, acc[u]);
        if (wd.is_root) {
          const double lh =
              fma(acc[u].w, a.pi[3], fma(acc[u].z, a.pi[2], fma(acc[u].y, a.pi[1], acc[u].x * a.pi[0])));
          a.lh_out[p * K + b] = lh;
        }
      }
    }
    // signal the strips this warp stored for node c_w -- once, when it leaves the node
    ++n_unsignalled;
    if (!more || g_w != c_w) {
      __syncwarp();
      if (lane == 0) {
        fence_acq_rel_gpu();
        red_relaxed_gpu_add(node_done + c_w, (unsigned)n_unsignalled);
      }
      n_unsignalled = 0;
    }
    if (pending) {
      // the next strip's node was not ready before: now it is safe to block
      seek_gather(s + GW, true);
      gather();
      if (more_idx) load_idx(s + 2 * GW);
    }
  }
}

// ---------------------------------------------------------------------------------------
// 4-state RECOMPUTING kernel ("rec"): several consecutive tree levels in one launch WITHOUT any barrier.
//
// The lower levels of a tree hold little data (C2: 0.3, 1.1 and 15 MB for levels 1-2, 3, 4) but each costs a
// launch or a barrier plus two dependent global latencies -- 47 us of a 360 us C2 step, most of a brca1-sized
// one.  What makes a level wait for the one below is only that it LOADS the rows the level below has stored.
// Here a unit (row, bin) of a node does not load the rows of a child that is computed in the same launch: it
// RECOMPUTES them, recursively, down to children whose rows are already there (tips, clean nodes, nodes of an
// earlier launch).  A node REC_DEPTH levels above the bottom of the launch evaluates a subtree of at most
// 2^REC_DEPTH - 1 nodes per unit -- a few dozen FP64 FMAs that nobody waits for -- and in exchange every node of
// all the launch's levels is computed at once: no barrier, one chain of ~REC_DEPTH + 1 dependent loads.  Every
// node still stores its own rows (the resident state later updates start from).  Each row is the same expression
// over the same inputs whoever evaluates it: bit-identical to the level-by-level kernels.
// ---------------------------------------------------------------------------------------
constexpr int REC_THREADS = 256;
constexpr int REC_MAX_WORK = 192;    // work items (dirty nodes) per launch
constexpr int REC_MAX_CHILDREN = 3;
constexpr int REC_DEPTH = 3;         // recursion depth: a launch spans at most REC_DEPTH + 1 levels
constexpr int REC_MAX_P = 192;       // P matrices kept in shared memory

struct RecSmem {
  NodeDesc nodes[REC_MAX_WORK];
  ChildRef childs[REC_MAX_WORK * REC_MAX_CHILDREN];
  int32_t child_work[REC_MAX_WORK * REC_MAX_CHILDREN];  // work item that computes the child in THIS launch, or -1
  int32_t prefix[REC_MAX_WORK + 1];
  double P[REC_MAX_P * 16];
};

template <int D>
__device__ d4 rec_eval(const RecSmem& sm, bool p_staged, int wi, int64_t p, int b, int K) {
  const NodeDesc& nd = sm.nodes[wi];
  d4 acc;
  for (int c = 0; c < nd.n_children; ++c) {
    const ChildRef& ch = sm.childs[wi * REC_MAX_CHILDREN + c];
    const int32_t r = ch.idx[p];
    const int cw = sm.child_work[wi * REC_MAX_CHILDREN + c];
    d4 x;
    if (D > 0 && cw >= 0) {
      x = rec_eval<(D > 0 ? D - 1 : 0)>(sm, p_staged, cw, r, b, K);  // computed in this launch: recompute, do not wait
    } else {
      x = ld256(ch.src + ((int64_t)r * K + b) * 4);  // rows that are already there
    }
    d4 y = x;
    if (ch.P != nullptr) {
      y = p_staged ? matvec4(sm.P + ((wi * REC_MAX_CHILDREN + c) * K + b) * 16, x) : matvec4(ch.P + b * 16, x);
    }
    if (c == 0) acc = y;
    else { acc.x *= y.x; acc.y *= y.y; acc.z *= y.z; acc.w *= y.w; }
  }
  if (nd.fixed_motif >= 0) {
    if (nd.fixed_motif != 0) acc.x = 0.0;
    if (nd.fixed_motif != 1) acc.y = 0.0;
    if (nd.fixed_motif != 2) acc.z = 0.0;
    if (nd.fixed_motif != 3) acc.w = 0.0;
  }
  return acc;
}

__global__ void __launch_bounds__(REC_THREADS)
prune4_rec_kernel(const PruneArgs a, const int32_t* __restrict__ child_work) {
  extern __shared__ __align__(16) unsigned char rec_raw[];
  RecSmem& sm = *reinterpret_cast<RecSmem*>(rec_raw);
  pdl_trigger();
  const int K = a.K;
  const int n_work = a.n_work;
  for (int i = threadIdx.x; i <= n_work; i += REC_THREADS) sm.prefix[i] = a.tile_prefix[i];
  for (int i = threadIdx.x; i < n_work; i += REC_THREADS) {
    const NodeDesc nd = a.nodes[a.work_node[i]];
    sm.nodes[i] = nd;
    for (int c = 0; c < nd.n_children && c < REC_MAX_CHILDREN; ++c) {
      sm.childs[i * REC_MAX_CHILDREN + c] = a.childs[nd.child_begin + c];
      sm.child_work[i * REC_MAX_CHILDREN + c] = child_work[i * REC_MAX_CHILDREN + c];
    }
  }
  __syncthreads();
  pdl_wait();  // the P matrices / tip tables of the expm launch, rows of earlier launches
  const bool p_staged = (int64_t)n_work * REC_MAX_CHILDREN * K <= REC_MAX_P;
  if (p_staged) {
    const int per_item = REC_MAX_CHILDREN * K * 16;
    for (int o = threadIdx.x; o < n_work * per_item; o += REC_THREADS) {
      const int i = o / per_item, rem = o - i * per_item;
      const int c = rem / (K * 16), e16 = rem - c * (K * 16);
      const double* P = (c < sm.nodes[i].n_children) ? sm.childs[i * REC_MAX_CHILDREN + c].P : nullptr;
      if (P != nullptr) sm.P[o] = P[e16];
    }
    __syncthreads();
  }
  const int total = a.total_tiles;  // (row, bin) units of all work items
  for (int u = blockIdx.x * REC_THREADS + threadIdx.x; u < total; u += gridDim.x * REC_THREADS) {
    const int wi = find_work(sm.prefix, n_work, u);
    const int local = u - sm.prefix[wi];
    const int64_t p = local / K;
    const int b = local - (int)p * K;
    const d4 acc = rec_eval<REC_DEPTH>(sm, p_staged, wi, p, b, K);
    const NodeDesc& nd = sm.nodes[wi];
    if (nd.out != nullptr) st256(nd.out + (p * K + b) * 4, acc);
    if (nd.is_root) {
      const double lh = fma(acc.w, a.pi[3], fma(acc.z, a.pi[2], fma(acc.y, a.pi[1], acc.x * a.pi[0])));
      a.lh_out[p * K + b] = lh;
    }
  }
}

// ---------------------------------------------------------------------------------------
// 4-state PATH kernel: the optimiser's inner loop on real-data sizes.
//
// Powell / simulated annealing move ONE branch length at a time (maths/scipy_optimize.py:371-383,
// 488-541); the dirty set of such an update is a single path from the node above the changed branch
// to the root: one node per level, each the parent of the previous one.  Through the small-levels
// launch that is one cluster barrier plus two dependent global latencies per level -- ~2.5 us x 11
// levels on brca1, the bulk of a 65 us update.  Here ONE CTA walks the path and keeps the rows of the
// node it has just computed in SHARED memory: the next node gathers its dirty child from there (no
// global round trip, a block barrier instead of a cluster barrier); its clean children come from
// global memory, where they have been all along.  Every node's rows are still written to global
// memory (they are the resident state later updates start from), the root's lh likewise, and the
// site-weighted reduction follows in the same launch -- the same tiles and summation trees as
// lnl_reduce_kernel, the same per-unit arithmetic as every other 4-state kernel: bit-identical.
// Applies when rows x K of every node on the path fit two shared-memory buffers (brca1: 2764 rows).
// ---------------------------------------------------------------------------------------
constexpr int PATH_THREADS_MAX = 1024;
constexpr int PATH_MAX_NODES = 64;
constexpr int PATH_MAX_CHILDREN = 3;

struct PathSmemHead {
  NodeDesc nodes[PATH_MAX_NODES];
  ChildRef childs[PATH_MAX_NODES * PATH_MAX_CHILDREN];
  int32_t dirty_child[PATH_MAX_NODES];
};

// dirty_child[i]: which child slot of work item i is work item i - 1 (-1 for the first)
// <threads per CTA, units per thread worked as one batch>: brca1 has 2764 units per level
template <int PATH_THREADS, int PATH_UPT>
__global__ void __launch_bounds__(PATH_THREADS, 1)
prune4_path_kernel(const PruneArgs a, const int32_t* __restrict__ dirty_child, int buf_units, const ReduceTail tail) {
  extern __shared__ __align__(32) unsigned char path_raw[];
  PathSmemHead& sh = *reinterpret_cast<PathSmemHead*>(path_raw);
  constexpr size_t HEAD = (sizeof(PathSmemHead) + 255) / 256 * 256;
  double* sP = reinterpret_cast<double*>(path_raw + HEAD);  // [n_work][3][K][16]
  const int K = a.K;
  const int n_work = a.n_work;
  double* buf0 = sP + (size_t)n_work * PATH_MAX_CHILDREN * K * 16;
  double* buf1 = buf0 + (size_t)buf_units * 4;
  const int t = (int)threadIdx.x;

  pdl_trigger();
  for (int i = t; i < n_work; i += PATH_THREADS) {
    const NodeDesc nd = a.nodes[a.work_node[i]];
    sh.nodes[i] = nd;
    sh.dirty_child[i] = dirty_child[i];
    for (int c = 0; c < nd.n_children && c < PATH_MAX_CHILDREN; ++c)
      sh.childs[i * PATH_MAX_CHILDREN + c] = a.childs[nd.child_begin + c];
  }
  __syncthreads();
  pdl_wait();  // the P matrices / tip tables of the expm launch
  {
    const int per_item = PATH_MAX_CHILDREN * K * 16;
    for (int o = t; o < n_work * per_item; o += PATH_THREADS) {
      const int i = o / per_item, rem = o - i * per_item;
      const int c = rem / (K * 16), e16 = rem - c * (K * 16);
      const double* P = (c < sh.nodes[i].n_children) ? sh.childs[i * PATH_MAX_CHILDREN + c].P : nullptr;
      if (P != nullptr) sP[o] = P[e16];
    }
  }
  __syncthreads();

  double* prev = buf0;  // rows of work item i - 1, [rows][K][4]
  double* cur = buf1;
  // Every thread owns up to PATH_UPT units of a level (u = t + j * PATH_THREADS) and works them as a BATCH:
  // the gather indexes of a level are requested while the previous level is finishing, the rows of a clean
  // child are requested for the whole batch before any of them is used -- a level costs about one global
  // latency, not (units per thread) x (two dependent latencies).  Levels with more units fall back to the
  // unit-at-a-time loop.
  constexpr int UPT = PATH_UPT;
  int32_t ix[UPT][PATH_MAX_CHILDREN];
  auto request_idx = [&](int i) {
    if (i >= n_work) return;
    const NodeDesc& nd = sh.nodes[i];
    const int units = (int)nd.n_rows * K;
    if (units > UPT * PATH_THREADS) return;
#pragma unroll
    for (int j = 0; j < UPT; ++j) {
      const int u = t + j * PATH_THREADS;
      if (u < units) {
        const int p = u / K;
#pragma unroll
        for (int c = 0; c < PATH_MAX_CHILDREN; ++c)
          if (c < nd.n_children) ix[j][c] = __ldg(sh.childs[i * PATH_MAX_CHILDREN + c].idx + p);
      }
    }
  };
  request_idx(0);
  for (int i = 0; i < n_work; ++i) {
    const NodeDesc nd = sh.nodes[i];
    const ChildRef* ch = sh.childs + i * PATH_MAX_CHILDREN;
    const double* Pi = sP + (size_t)i * PATH_MAX_CHILDREN * K * 16;
    const int dc = sh.dirty_child[i];
    const int units = (int)nd.n_rows * K;
    auto finish = [&](d4 acc, int u) {
      const int p = u / K, b = u - p * K;
      if (nd.fixed_motif >= 0) {
        if (nd.fixed_motif != 0) acc.x = 0.0;
        if (nd.fixed_motif != 1) acc.y = 0.0;
        if (nd.fixed_motif != 2) acc.z = 0.0;
        if (nd.fixed_motif != 3) acc.w = 0.0;
      }
      if (nd.out != nullptr) st256(nd.out + ((int64_t)p * K + b) * 4, acc);
      if (nd.is_root) {
        const double lh = fma(acc.w, a.pi[3], fma(acc.z, a.pi[2], fma(acc.y, a.pi[1], acc.x * a.pi[0])));
        a.lh_out[(int64_t)p * K + b] = lh;
      } else {
        double* dst = cur + (size_t)u * 4;
        dst[0] = acc.x; dst[1] = acc.y; dst[2] = acc.z; dst[3] = acc.w;
      }
    };
    if (units <= UPT * PATH_THREADS) {
      d4 acc[UPT];
#pragma unroll
      for (int c = 0; c < PATH_MAX_CHILDREN; ++c) {
        if (c >= nd.n_children) break;
        d4 x[UPT];
        if (c == dc) {
#pragma unroll
          for (int j = 0; j < UPT; ++j) {
            const int u = t + j * PATH_THREADS;
            if (u < units) {
              const double* src = prev + ((size_t)ix[j][c] * K + (u % K)) * 4;
              x[j].x = src[0]; x[j].y = src[1]; x[j].z = src[2]; x[j].w = src[3];
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < UPT; ++j) {  // the whole batch in flight together
            const int u = t + j * PATH_THREADS;
            if (u < units) x[j] = ld256(ch[c].src + ((int64_t)ix[j][c] * K + (u % K)) * 4);
          }
        }
#pragma unroll
        for (int j = 0; j < UPT; ++j) {
          const int u = t + j * PATH_THREADS;
          if (u < units) {
            const d4 y = (ch[c].P != nullptr) ? matvec4(Pi + (c * K + (u % K)) * 16, x[j]) : x[j];
            if (c == 0) acc[j] = y;
            else { acc[j].x *= y.x; acc[j].y *= y.y; acc[j].z *= y.z; acc[j].w *= y.w; }
          }
        }
      }
      request_idx(i + 1);  // (static tables: in flight across the stores and the barrier)
#pragma unroll
      for (int j = 0; j < UPT; ++j) {
        const int u = t + j * PATH_THREADS;
        if (u < units) finish(acc[j], u);
      }
    } else {
      for (int u = t; u < units; u += PATH_THREADS) {
        const int p = u / K, b = u - p * K;
        d4 acc;
#pragma unroll
        for (int c = 0; c < PATH_MAX_CHILDREN; ++c) {
          if (c >= nd.n_children) break;
          const int32_t r = ch[c].idx[p];
          d4 x;
          if (c == dc) {
            const double* src = prev + ((size_t)r * K + b) * 4;
            x.x = src[0]; x.y = src[1]; x.z = src[2]; x.w = src[3];
          } else {
            x = ld256(ch[c].src + ((int64_t)r * K + b) * 4);
          }
          const d4 y = (ch[c].P != nullptr) ? matvec4(Pi + (c * K + b) * 16, x) : x;
          if (c == 0) acc = y;
          else { acc.x *= y.x; acc.y *= y.y; acc.z *= y.z; acc.w *= y.w; }
        }
        finish(acc, u);
      }
      request_idx(i + 1);
    }
    __syncthreads();
    double* tmp = prev; prev = cur; cur = tmp;
  }
  if (tail.n_tiles > 0) {
    // (the lh values were written by this CTA: visible to it after the barrier above)
    static_assert(PATH_THREADS >= REDUCE_THREADS, "the reduction tail works with REDUCE_THREADS threads");
    __shared__ double s_warp[PATH_THREADS_MAX / 32];
    __shared__ double s_tiles[SMALL_REDUCE_MAX_TILES];
    for (int tile = 0; tile < tail.n_tiles; ++tile) {
      const double acc = (t < REDUCE_THREADS)
                             ? reduce_tile_partial(a.lh_out, tail.bprobs, tail.counts, tail.n_rows, K, tile, t, nullptr)
                             : 0.0;
      const double total = block_sum(acc, s_warp);  // (the warps past REDUCE_THREADS add +0.0)
      if (t == 0) s_tiles[tile] = total;
    }
    __syncthreads();
    double acc = 0.0;
    if (t < REDUCE_THREADS)
      for (int i = t; i < tail.n_tiles; i += REDUCE_THREADS) acc += s_tiles[i];
    const double total = block_sum(acc, s_warp);
    if (t == 0) publish_result(tail.sink, total);
  }
}

template <int MP>
struct DmmaCfg {
  static constexpr int LDS = MP + 4;  // row stride (doubles): conflict-free fragment loads
  static constexpr int NI = MP / 8;   // n-blocks of 8 columns
  static constexpr int KS = MP / 4;   // k-steps of 4
  static constexpr int A_DOUBLES = DMMA_TILE_ROWS * LDS;
  static constexpr int B_DOUBLES = MP * LDS;
  static constexpr int NB = 2;  // resident P matrices
  static constexpr size_t SMEM_BYTES = (size_t)(A_DOUBLES + NB * B_DOUBLES) * sizeof(double);
};

template <int MP>
__device__ __forceinline__ void load_P_to_smem(double* __restrict__ sB, const double* __restrict__ P) {
  constexpr int LDS = DmmaCfg<MP>::LDS;
  for (int o = threadIdx.x; o < MP * MP / 2; o += DMMA_THREADS) {
    const int i = (2 * o) / MP, j = (2 * o) - i * MP;
    const double2 v = *reinterpret_cast<const double2*>(P + i * MP + j);
    *reinterpret_cast<double2*>(sB + i * LDS + j) = v;
  }
}

template <int MP, bool ROOT>
__global__ void __launch_bounds__(DMMA_THREADS) prune_dmma_kernel(const PruneArgs a) {
  using Cfg = DmmaCfg<MP>;
  constexpr int LDS = Cfg::LDS, NI = Cfg::NI, KS = Cfg::KS;
  extern __shared__ __align__(16) double smem[];
  double* sA = smem;
  double* sB = smem + Cfg::A_DOUBLES;  // [NB][MP*LDS]
  const int K = a.K;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, q = lane & 3;  // fragment row group / quad lane

  int cur_node = -1, cur_bin = -1;
  for (int t = blockIdx.x; t < a.total_tiles; t += gridDim.x) {
    const int w = find_work(a.tile_prefix, a.n_work, t);
    const int node_id = a.work_node[w];
    const NodeDesc nd = a.nodes[node_id];
    const int64_t row_tiles = (nd.n_rows + DMMA_TILE_ROWS - 1) / DMMA_TILE_ROWS;
    const int local = t - a.tile_prefix[w];
    const int b = (int)(local / row_tiles);
    const int64_t row0 = (int64_t)(local - b * row_tiles) * DMMA_TILE_ROWS;
    const int nc = nd.n_children;

    // which children need the contraction; the first NB of them keep P resident
    if (node_id != cur_node || b != cur_bin) {
      __syncthreads();
      int slot = 0;
      for (int c = 0; c < nc && slot < Cfg::NB; ++c) {
        const double* P = a.childs[nd.child_begin + c].P;
        if (P != nullptr) {
          load_P_to_smem<MP>(sB + slot * Cfg::B_DOUBLES, P + (int64_t)b * MP * MP);
          ++slot;
        }
      }
      cur_node = node_id;
      cur_bin = b;
      __syncthreads();
    }

    const int64_t my_row = row0 + warp * 8 + g;  // the C-fragment row of this thread
    const bool row_ok = my_row < nd.n_rows;
    const int64_t my_row_c = row_ok ? my_row : nd.n_rows - 1;

    double prod[NI][2];
    int dmma_seen = 0;
    for (int c = 0; c < nc; ++c) {
      const ChildRef ch = a.childs[nd.child_begin + c];
      double cur[NI][2];
      if (ch.P == nullptr) {
        // pre-multiplied (tip) child: gather the finished row straight into the C layout
        const int32_t r = __ldg(ch.idx + my_row_c);
        const double* src = ch.src + ((int64_t)r * K + b) * MP + 2 * q;
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) {
          const double2 v = __ldg(reinterpret_cast<const double2*>(src + 8 * ni));
          cur[ni][0] = v.x;
          cur[ni][1] = v.y;
        }
      } else {
        const double* sBc;
        if (dmma_seen < Cfg::NB) {
          sBc = sB + dmma_seen * Cfg::B_DOUBLES;
        } else {
          // more contraction children than resident slots (e.g. a trifurcating root of
          // internal nodes): stream this P through slot 0, restore on the next tile
          __syncthreads();
          load_P_to_smem<MP>(sB, ch.P + (int64_t)b * MP * MP);
          cur_node = -1;
          sBc = sB;
        }
        ++dmma_seen;
        // gather 64 child rows into shared memory (16 B per lane, coalesced per row)
        constexpr int CHUNKS = MP / 2;  // 16-byte chunks per row
        for (int o = tid; o < DMMA_TILE_ROWS * CHUNKS; o += DMMA_THREADS) {
          const int rr = o / CHUNKS, ck = o - rr * CHUNKS;
          int64_t p = row0 + rr;
          if (p >= nd.n_rows) p = nd.n_rows - 1;
          const int32_t r = __ldg(ch.idx + p);
          cp_async16(sA + rr * LDS + 2 * ck, ch.src + ((int64_t)r * K + b) * MP + 2 * ck);
        }
        cp_async_commit();
        cp_async_wait_all();
        __syncthreads();
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) cur[ni][0] = cur[ni][1] = 0.0;
        const double* aRow = sA + (warp * 8 + g) * LDS + q;
        const double* bRow = sBc + g * LDS + q;
#pragma unroll 4
        for (int ks = 0; ks < KS; ++ks) {
          const double av = aRow[4 * ks];
#pragma unroll
          for (int ni = 0; ni < NI; ++ni) {
            const double bv = bRow[(8 * ni) * LDS + 4 * ks];
            dmma_m8n8k4(cur[ni][0], cur[ni][1], av, bv);
          }
        }
        __syncthreads();  // sA (and a streamed sB) may be overwritten by the next child
      }
      if (c == 0) {
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) { prod[ni][0] = cur[ni][0]; prod[ni][1] = cur[ni][1]; }
      } else {
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) { prod[ni][0] *= cur[ni][0]; prod[ni][1] *= cur[ni][1]; }
      }
    }
    if (nd.fixed_motif >= 0) {
#pragma unroll
      for (int ni = 0; ni < NI; ++ni) {
        if (8 * ni + 2 * q != nd.fixed_motif) prod[ni][0] = 0.0;
        if (8 * ni + 2 * q + 1 != nd.fixed_motif) prod[ni][1] = 0.0;
      }
    }
    if (row_ok) {
      double* dst = nd.out + ((int64_t)my_row * K + b) * MP + 2 * q;
#pragma unroll
      for (int ni = 0; ni < NI; ++ni)
        *reinterpret_cast<double2*>(dst + 8 * ni) = make_double2(prod[ni][0], prod[ni][1]);
    }
    if (ROOT) {
      double s = 0.0;
#pragma unroll
      for (int ni = 0; ni < NI; ++ni) {
        s = fma(prod[ni][0], a.pi[8 * ni + 2 * q], s);
        s = fma(prod[ni][1], a.pi[8 * ni + 2 * q + 1], s);
      }
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      if (row_ok && q == 0) a.lh_out[my_row * K + b] = s;
    }
  }
}

// ---------------------------------------------------------------------------------------
// DMMA kernel, warp-private register pipeline ("dmma-reg").
//
// ncu on the first version (C4): FP64 tensor pipe 50 % active -- every (tile, child) unit is
// "gather 64 rows into shared memory, wait, barrier, contract, barrier", two CTAs per SM taking
// turns.  Here the gathered rows never touch shared memory: with the k index of the contraction
// PERMUTED (step s, quad lane q  ->  column 8 (s>>1) + 2 q + (s&1)) the m8n8k4 A fragment of a
// lane is a run of adjacent column PAIRS of its row, i.e. MP/8 128-bit loads straight from
// global memory -- the same access pattern as the C fragment that tip children (pre-multiplied
// rows) and the output use.  P is stored in shared memory with the same permutation applied to
// its columns, so the B-fragment loads keep the conflict-free pattern of the first version.
// Each warp owns 8 rows of a 96-row super-tile and runs its own software pipeline: while the
// tensor core works on unit u, the rows of unit u+1 are in flight into a second register
// buffer and the gather index of unit u+2 is being fetched.  (Refilling the consumed buffer in
// place for unit u+2 spills at 168 registers and measured 1.61 vs 1.27 ms on C4: dropped.
// Taking the tip children (tables of <= 62 rows, L1 / shared-memory resident) out of the
// pipeline so that they do not displace the next contraction unit's prefetch: 1.55 ms with the
// tables staged in shared memory, 1.50 ms with the rows loaded at use from L1 -- both dropped.
// Two units in flight WITHOUT spills (8 warps per SM, 236 registers, gather indexes four units
// ahead as in regpipe2): 1.50 ms -- the tensor pipe wants the 12 warps more than the depth.
// Starting the three warps of a scheduler a third of a unit apart: no change.)  No barrier except when the CTA
// moves to another (node, rate class) and reloads P; CTAs take contiguous ranges of super-tiles
// so that happens once or twice per launch.  One CTA of 12 warps per SM (register bound).
// Summation order over k differs from the first version (permuted), so results differ from it
// in the last bits; within the kernel they are deterministic.
// ---------------------------------------------------------------------------------------
constexpr int DR_WARPS = 12;
constexpr int DR_THREADS = DR_WARPS * 32;
constexpr int DR_TILE_ROWS = DR_WARPS * 8;
constexpr int DR_MAXP = 4;   // P matrices resident in shared memory (contraction children per node)
constexpr int DR_MAXC = 4;   // children per node
constexpr int DR_MAXW = 48;  // work items per launch

struct RWork {
  double* out;
  int64_t n_rows;
  int64_t row_tiles;
  const double* src[DR_MAXC];
  const int32_t* idx[DR_MAXC];
  const double* P[DR_MAXC];
  int32_t slot[DR_MAXC];  // shared-memory slot of child c's P (contraction children only)
  int32_t nc;
  int32_t fixed_motif;
};
struct RSmem {
  RWork work[DR_MAXW];
  int32_t prefix[DR_MAXW + 1];
};
struct RCur {
  int st;        // super-tile
  int c;         // child
  int w;         // work item
  int b;         // rate class
  int64_t row0;  // first row of the super-tile
};

template <int MP>
struct DRegCfg {
  static constexpr int LDS = MP + 4;
  static constexpr int NI = MP / 8;
  static constexpr int KS = MP / 4;
  static constexpr int B_DOUBLES = MP * LDS;
  static constexpr size_t SMEM_BYTES = (size_t)DR_MAXP * B_DOUBLES * sizeof(double);
};

template <int MP, bool ROOT>
__global__ void __launch_bounds__(DR_THREADS, 1) prune_dmma_reg_kernel(const PruneArgs a) {
  using Cfg = DRegCfg<MP>;
  constexpr int LDS = Cfg::LDS, NI = Cfg::NI, KS = Cfg::KS;
  extern __shared__ __align__(16) double smem[];  // [DR_MAXP][MP * LDS], columns permuted
  __shared__ RSmem sd;
  const int K = a.K;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, q = lane & 3;

  pdl_trigger();
  for (int i = tid; i <= a.n_work; i += DR_THREADS) sd.prefix[i] = a.tile_prefix[i];
  for (int i = tid; i < a.n_work; i += DR_THREADS) {
    const NodeDesc nd = a.nodes[a.work_node[i]];
    RWork& wd = sd.work[i];
    wd.out = nd.out;
    wd.n_rows = nd.n_rows;
    wd.row_tiles = (nd.n_rows + DR_TILE_ROWS - 1) / DR_TILE_ROWS;
    wd.nc = nd.n_children;
    wd.fixed_motif = nd.fixed_motif;
    int slot = 0;
    for (int c = 0; c < nd.n_children && c < DR_MAXC; ++c) {
      const ChildRef ch = a.childs[nd.child_begin + c];
      wd.src[c] = ch.src;
      wd.idx[c] = ch.idx;
      wd.P[c] = ch.P;
      wd.slot[c] = (ch.P != nullptr) ? slot++ : -1;
    }
  }
  __syncthreads();

  // this CTA's contiguous range of super-tiles
  const int per_cta = (a.total_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
  const int st0 = (int)blockIdx.x * per_cta;
  const int st1 = min(a.total_tiles, st0 + per_cta);
  auto place = [&](RCur& cur) {
    while (cur.st >= sd.prefix[cur.w + 1]) ++cur.w;
    const RWork& wd = sd.work[cur.w];
    const int local = cur.st - sd.prefix[cur.w];
    cur.b = (int)(local / wd.row_tiles);
    cur.row0 = (int64_t)(local - cur.b * wd.row_tiles) * DR_TILE_ROWS;
  };
  auto advance = [&](RCur& cur) {
    if (++cur.c >= sd.work[cur.w].nc) {
      cur.c = 0;
      if (++cur.st < st1) place(cur);
    }
  };
  auto my_row = [&](const RCur& cur) -> int64_t {  // clamped: loads of rows past the end repeat the last
    const int64_t p = cur.row0 + warp * 8 + g;
    const int64_t last = sd.work[cur.w].n_rows - 1;
    return p < last ? p : last;
  };

  RCur cc{st0, 0, 0, 0, 0};
  if (cc.st < st1) place(cc);
  RCur dc = cc, ic = cc;
  int32_t idx_i = 0;
  double2 vn[NI];
  auto fetch = [&](const RCur& cur, int32_t r) {
    const double* src = sd.work[cur.w].src[cur.c] + ((int64_t)r * K + cur.b) * MP + 2 * q;
#pragma unroll
    for (int ni = 0; ni < NI; ++ni) vn[ni] = ld128_nc(src + 8 * ni);
  };
  if (ic.st < st1) idx_i = __ldg(sd.work[ic.w].idx[ic.c] + my_row(ic));
  pdl_wait();
  if (dc.st < st1) {
    const int32_t r = idx_i;
    advance(ic);
    if (ic.st < st1) idx_i = __ldg(sd.work[ic.w].idx[ic.c] + my_row(ic));
    fetch(dc, r);
  }

  int res_w = -1, res_b = -1;
  double prod[NI][2];
#pragma unroll 1
  while (cc.st < st1) {
    const RWork& wd = sd.work[cc.w];
    if (cc.c == 0 && (cc.w != res_w || cc.b != res_b)) {
      // another (node, rate class): every warp is done with the old P matrices
      __syncthreads();
      for (int c = 0; c < wd.nc; ++c) {
        if (wd.P[c] == nullptr) continue;
        double* sB = smem + wd.slot[c] * Cfg::B_DOUBLES;
        const double* P = wd.P[c] + (int64_t)cc.b * MP * MP;
        // coalesced reads of P; column j = 8a + 2q + b lands at position 4 (2a + b) + q
        for (int o = tid; o < MP * MP; o += DR_THREADS) {
          const int i = o / MP, j = o - i * MP;
          const int pos = (j & ~7) + ((j & 1) << 2) + ((j >> 1) & 3);
          sB[i * LDS + pos] = P[o];
        }
      }
      res_w = cc.w;
      res_b = cc.b;
      __syncthreads();
    }
    // take over the rows in flight, request the next unit's
    double2 v[NI];
#pragma unroll
    for (int ni = 0; ni < NI; ++ni) v[ni] = vn[ni];
    advance(dc);
    if (dc.st < st1) {
      const int32_t r = idx_i;
      advance(ic);
      if (ic.st < st1) idx_i = __ldg(sd.work[ic.w].idx[ic.c] + my_row(ic));
      fetch(dc, r);
    }

    double cur[NI][2];
    if (wd.P[cc.c] == nullptr) {
#pragma unroll
      for (int ni = 0; ni < NI; ++ni) { cur[ni][0] = v[ni].x; cur[ni][1] = v[ni].y; }
    } else {
#pragma unroll
      for (int ni = 0; ni < NI; ++ni) cur[ni][0] = cur[ni][1] = 0.0;
      const double* bRow = smem + wd.slot[cc.c] * Cfg::B_DOUBLES + g * LDS + q;
#pragma unroll
      for (int s = 0; s < KS; ++s) {
        const double av = (s & 1) ? v[s >> 1].y : v[s >> 1].x;
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) dmma_m8n8k4(cur[ni][0], cur[ni][1], av, bRow[(8 * ni) * LDS + 4 * s]);
      }
    }
#pragma unroll
    for (int ni = 0; ni < NI; ++ni) {
      if (cc.c == 0) { prod[ni][0] = cur[ni][0]; prod[ni][1] = cur[ni][1]; }
      else { prod[ni][0] *= cur[ni][0]; prod[ni][1] *= cur[ni][1]; }
    }
    if (cc.c == wd.nc - 1) {
      const int64_t p = cc.row0 + warp * 8 + g;
      const bool row_ok = p < wd.n_rows;
      if (wd.fixed_motif >= 0) {
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) {
          if (8 * ni + 2 * q != wd.fixed_motif) prod[ni][0] = 0.0;
          if (8 * ni + 2 * q + 1 != wd.fixed_motif) prod[ni][1] = 0.0;
        }
      }
      if (row_ok) {
        double* dst = wd.out + (p * K + cc.b) * MP + 2 * q;
#pragma unroll
        for (int ni = 0; ni < NI; ++ni)
          *reinterpret_cast<double2*>(dst + 8 * ni) = make_double2(prod[ni][0], prod[ni][1]);
      }
      if (ROOT) {
        double sdot = 0.0;
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) {
          sdot = fma(prod[ni][0], a.pi[8 * ni + 2 * q], sdot);
          sdot = fma(prod[ni][1], a.pi[8 * ni + 2 * q + 1], sdot);
        }
        sdot += __shfl_xor_sync(0xffffffffu, sdot, 1);
        sdot += __shfl_xor_sync(0xffffffffu, sdot, 2);
        if (row_ok && q == 0) a.lh_out[p * K + cc.b] = sdot;
      }
    }
    advance(cc);
  }
}


// ---------------------------------------------------------------------------------------
// 4-state "tree walk" kernel: ANY evaluation of a small problem (full sweeps included) in one launch with no
// barrier between the levels -- the chain kernel's idea for an arbitrary set of dirty nodes.
//
// The thread of (root row p, bin b) evaluates every dirty node, children before parents, for the ONE row of
// each node that root row p looks at (map_n[p]).  A clean child's row (a tip's pre-multiplied row, an internal
// node that did not change) is gathered from memory at map_child[p]; a dirty child's value was computed by this
// very thread a few steps earlier and waits in a slot of a per-thread value stack in shared memory (slots
// assigned by the host: a value's slot is free again once its parent is done).  Every node row is also stored
// (later partial updates read it; root rows that share a lower row store the same bits), lh is written at the
// root, the tiles are reduced as in the chain kernel.  The work is redundant -- a lower node has fewer distinct
// rows than the root has patterns -- and irrelevant at this size: brca1 (35 internal nodes, 2764 x K units) is
// ~100k matvecs more, against 11 levels x (cluster barrier + two dependent global latencies) saved.
// Arithmetic per row as in every 4-state kernel: bit-identical.
// ---------------------------------------------------------------------------------------
constexpr int WALK_MAX_OPS = 224;    // dirty nodes one launch walks (descriptors staged in shared memory)
constexpr int WALK_MAX_SLOTS = 20;   // live values per thread (8 KB of shared memory per slot)
constexpr int WALK_THREADS = REDUCE_THREADS;

struct WalkOp {
  double* out;             // the node's rows (nullptr: the root's are not stored)
  const int32_t* map_own;  // root row -> row of the node
  const double* src[3];    // clean child: its rows
  const double* P[3];      // P of the branch to child c (nullptr: the child's rows are already multiplied, tips)
  const int32_t* map_child[3];
  int8_t n_children;
  int8_t slot[3];          // >= 0: child c is dirty, its value waits in this stack slot; -1: gather from memory
  int8_t out_slot;         // where this node's value goes (-1: nobody above needs it, the root)
  int8_t fixed_motif;
  int8_t is_root;
  uint8_t shape;           // kind of child 0 | kind of child 1 << 2 | kind of child 2 << 4: 0 none, 1 memory without P
                           // (a tip's pre-multiplied row), 2 memory with P (a clean internal node), 3 stack (dirty, P)
  int16_t p_idx[3];        // P[c] staged in shared memory: entry number ([K][16] doubles each), -1: none
  int16_t pad2;
  int64_t pad3;
};
static_assert(sizeof(WalkOp) == 112, "WalkOp layout");

__global__ void __launch_bounds__(WALK_THREADS)
prune4_walk_kernel(const PruneArgs a, const WalkOp* __restrict__ ops, int n_ops, int n_slots, int n_P, const ReduceTail tail,
                   unsigned int* __restrict__ counters, double* __restrict__ tile_sums) {
  extern __shared__ __align__(16) unsigned char walk_raw[];
  // [n_ops] descriptors | [n_slots][2][threads] value stack | [n_P][K][16] P matrices | [n_P] their addresses
  WalkOp* s_ops = reinterpret_cast<WalkOp*>(walk_raw);
  double2* s_stack = reinterpret_cast<double2*>(walk_raw + (size_t)n_ops * sizeof(WalkOp));
  double* s_P = reinterpret_cast<double*>(s_stack + (size_t)n_slots * 2 * WALK_THREADS);
  const double** s_Pptr = reinterpret_cast<const double**>(s_P + (size_t)n_P * a.K * 16);
  __shared__ double s_warp[WALK_THREADS / 32];
  __shared__ bool s_last;
  const int K = a.K, t = (int)threadIdx.x;
  pdl_trigger();
  if (t == 0) tl_first(a.tl, 0);
  pdl_wait();  // the parameter block is this evaluation's; P matrices / tip tables of the expm launch
  if (t == 0) tl_first(a.tl, 1);
  {
    const int4* g = reinterpret_cast<const int4*>(ops);
    int4* d = reinterpret_cast<int4*>(s_ops);
    for (int i = t; i < n_ops * (int)(sizeof(WalkOp) / 16); i += WALK_THREADS) d[i] = g[i];
  }
  __syncthreads();
  const int64_t total = tail.n_rows * K;
  const int64_t u = (int64_t)blockIdx.x * WALK_THREADS + t;
  const bool active = u < total;
  const int64_t p = active ? u / K : 0;
  const int b = active ? (int)(u - p * K) : 0;
  auto push = [&](int slot, const d4& v) {
    s_stack[(slot * 2 + 0) * WALK_THREADS + t] = make_double2(v.x, v.y);
    s_stack[(slot * 2 + 1) * WALK_THREADS + t] = make_double2(v.z, v.w);
  };
  auto pop = [&](int slot) {
    const double2 lo = s_stack[(slot * 2 + 0) * WALK_THREADS + t], hi = s_stack[(slot * 2 + 1) * WALK_THREADS + t];
    return d4{lo.x, lo.y, hi.x, hi.y};
  };
  // software pipeline over the ops (every address is static: it only depends on p)
  auto load_idx = [&](int j, int32_t(&ix)[4]) {
    if (j >= n_ops) return;
    const WalkOp& op = s_ops[j];
    ix[3] = (op.out != nullptr) ? __ldg(op.map_own + p) : -1;  // (negative: another root row stores this row)
    const unsigned shape = op.shape;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const unsigned kind = (shape >> (2 * c)) & 3u;  // 1, 2: gathered from memory
      if (kind == 1u || kind == 2u) ix[c] = __ldg(op.map_child[c] + p);  // (top bit masked at use: no wait here)
    }
  };
  auto load_rows = [&](int j, const int32_t(&ix)[4], d4(&x)[3]) {
    if (j >= n_ops) return;
    const WalkOp& op = s_ops[j];
    const unsigned shape = op.shape;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const unsigned kind = (shape >> (2 * c)) & 3u;
      if (kind == 1u || kind == 2u) x[c] = ld256(op.src[c] + ((int64_t)(ix[c] & 0x7fffffff) * K + b) * 4);
    }
  };
  // the P matrices of the walk into shared memory (read at use from global memory they cost every op an L2 round
  // trip per child -- a launch starts with a cold L1 -- which is what the first version spent its time on)
  for (int i = t; i < n_ops * 3; i += WALK_THREADS) {
    const int j = i / 3, c = i - j * 3;
    if (s_ops[j].p_idx[c] >= 0) s_Pptr[s_ops[j].p_idx[c]] = s_ops[j].P[c];
  }
  __syncthreads();
  for (int i = t; i < n_P * K * 16; i += WALK_THREADS) s_P[i] = s_Pptr[i / (K * 16)][i % (K * 16)];
  __syncthreads();
  // FOUR ops in flight per thread, each with its own row buffer and index set (no register is ever copied from one
  // stage to the next: a move would wait for the load it moves): after op j is computed from buffer q, the rows
  // of op j + 4 are requested into the same buffer with the indexes requested four ops earlier, and the indexes
  // of op j + 8 are requested -- an op costs a quarter of an L2 round trip
  // one child's contribution, specialised on its kind (std::integral_constant tag): no predicate survives
  auto child = [&](auto tag, const WalkOp& op, int c, const d4& xc) -> d4 {
    constexpr int KIND = decltype(tag)::value;
    d4 v = xc;
    if constexpr (KIND == 3) v = pop(op.slot[c]);
    if constexpr (KIND >= 2) return matvec4(s_P + ((int)op.p_idx[c] * K + b) * 16, v);
    return v;
  };
  auto binary = [&](auto t0, auto t1, const WalkOp& op, const d4(&x)[3]) -> d4 {
    const d4 y0 = child(t0, op, 0, x[0]), y1 = child(t1, op, 1, x[1]);
    return d4{y0.x * y1.x, y0.y * y1.y, y0.z * y1.z, y0.w * y1.w};
  };
  using k1 = std::integral_constant<int, 1>;
  using k2 = std::integral_constant<int, 2>;
  using k3 = std::integral_constant<int, 3>;
  auto step = [&](d4(&x)[3], int32_t(&ix)[4], int32_t& own, int j) {
    const WalkOp& op = s_ops[j];
    d4 acc{1.0, 1.0, 1.0, 1.0};
    // binary nodes (every node but a trifurcating root) through a body specialised on the two children's kinds:
    // the generic loop below spends most of its ~220 instructions per node on predicates and descriptor decoding
    switch (op.shape) {
      case 1 | 1 << 2: acc = binary(k1{}, k1{}, op, x); break;
      case 2 | 1 << 2: acc = binary(k2{}, k1{}, op, x); break;
      case 3 | 1 << 2: acc = binary(k3{}, k1{}, op, x); break;
      case 1 | 2 << 2: acc = binary(k1{}, k2{}, op, x); break;
      case 2 | 2 << 2: acc = binary(k2{}, k2{}, op, x); break;
      case 3 | 2 << 2: acc = binary(k3{}, k2{}, op, x); break;
      case 1 | 3 << 2: acc = binary(k1{}, k3{}, op, x); break;
      case 2 | 3 << 2: acc = binary(k2{}, k3{}, op, x); break;
      case 3 | 3 << 2: acc = binary(k3{}, k3{}, op, x); break;
      default:
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          if (c < op.n_children) {
            const d4 v = (op.slot[c] >= 0) ? pop(op.slot[c]) : x[c];
            d4 y = v;
            if (op.p_idx[c] >= 0) y = matvec4(s_P + ((int)op.p_idx[c] * K + b) * 16, v);
            if (c == 0) acc = y;
            else { acc.x *= y.x; acc.y *= y.y; acc.z *= y.z; acc.w *= y.w; }
          }
        }
    }
    if (op.fixed_motif >= 0) {
      if (op.fixed_motif != 0) acc.x = 0.0;
      if (op.fixed_motif != 1) acc.y = 0.0;
      if (op.fixed_motif != 2) acc.z = 0.0;
      if (op.fixed_motif != 3) acc.w = 0.0;
    }
    if (op.out_slot >= 0) push(op.out_slot, acc);
    if (active) {
      if (own >= 0) st256(op.out + ((int64_t)own * K + b) * 4, acc);
      if (op.is_root) {
        const double lh = fma(acc.w, a.pi[3], fma(acc.z, a.pi[2], fma(acc.y, a.pi[1], acc.x * a.pi[0])));
        a.lh_out[p * K + b] = lh;
      }
    }
    if (j + 4 < n_ops) {
      load_rows(j + 4, ix, x);
      own = ix[3];
      load_idx(j + 8, ix);
    }
  };
  int32_t ia[4] = {0, 0, 0, 0}, ib[4] = {0, 0, 0, 0}, ic[4] = {0, 0, 0, 0}, id[4] = {0, 0, 0, 0};
  int32_t oa, ob, oc, od;
  d4 xa[3], xb[3], xc[3], xd[3];
  load_idx(0, ia);
  load_idx(1, ib);
  load_idx(2, ic);
  load_idx(3, id);
  load_rows(0, ia, xa);
  load_rows(1, ib, xb);
  load_rows(2, ic, xc);
  load_rows(3, id, xd);
  oa = ia[3]; ob = ib[3]; oc = ic[3]; od = id[3];
  load_idx(4, ia);
  load_idx(5, ib);
  load_idx(6, ic);
  load_idx(7, id);
#pragma unroll 1
  for (int j = 0; j < n_ops; j += 4) {
    step(xa, ia, oa, j);
    if (j + 1 < n_ops) step(xb, ib, ob, j + 1);
    if (j + 2 < n_ops) step(xc, ic, oc, j + 2);
    if (j + 3 < n_ops) step(xd, id, od, j + 3);
  }
  if (t == 0) tl_end(a.tl);
  // ---- reduction: as in the chain kernel ----
  __threadfence();
  __syncthreads();
  const int ctas_per_tile = REDUCE_ROWS_PER_TILE * K / WALK_THREADS;
  const int tile = (int)blockIdx.x / ctas_per_tile;
  const int ctas_here = min(ctas_per_tile, (int)gridDim.x - tile * ctas_per_tile);
  if (t == 0) s_last = atomicAdd(counters + 1 + tile, 1u) == (unsigned)(ctas_here - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  {
    const double acc = reduce_tile_partial(a.lh_out, tail.bprobs, tail.counts, tail.n_rows, K, tile, t, nullptr);
    const double total_t = block_sum(acc, s_warp);
    if (t == 0) {
      tile_sums[tile] = total_t;
      counters[1 + tile] = 0u;
      __threadfence();
      s_last = atomicAdd(counters, 1u) == (unsigned)(tail.n_tiles - 1);
    }
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  double acc = 0.0;
  for (int i = t; i < tail.n_tiles; i += REDUCE_THREADS) acc += __ldcg(tile_sums + i);
  const double result = block_sum(acc, s_warp);
  if (t == 0) {
    counters[0] = 0u;
    publish_result(tail.sink, result);
    tl_end(a.tl);
  }
}


}  // namespace c3
