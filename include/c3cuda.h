/*
 * libc3cuda -- C ABI of the B200-native likelihood engine for cogent3's
 * substitution-model hot path.
 *
 * This is the drop-in boundary.  cogent3 has no FFI for this
 * path (it is pure Python + numba); every entry point below therefore cites the
 * reference function whose arithmetic it replaces (paths relative to
 * /root/reference/src/cogent3).  The binding a cogent3 maintainer would add is a
 * ctypes stub -- see INTEGRATION.md and cogent3_b200/_lib.py.
 *
 * Conventions
 *   - all functions return C3_OK (0) or a C3_ERR_* code; c3_last_error() gives a
 *     thread-local message.  No exceptions cross the ABI.
 *   - plain pointers and sizes only; host pointers unless the name says "device".
 *   - nodes are numbered in post-order (every child id < its parent id), the root
 *     is node n_nodes-1; "the edge of node n" is the branch above node n.
 *   - a node's pattern table has n_rows = P+1 rows; the last is the all-gap row
 *     with count 0 (evolve/likelihood_tree.py:49-51, 230-232).
 *   - P[i*M + j] = Pr(i -> j) (maths/matrix_exponentiation.py:35-40).
 *   - per-(edge, bin) arrays are laid out [node][bin] (node-major).
 */
#ifndef C3CUDA_H
#define C3CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define C3_ABI_VERSION 1

#if defined(__GNUC__)
#define C3_API __attribute__((visibility("default")))
#else
#define C3_API
#endif

#define C3_OK 0
#define C3_ERR_INVALID 1 /* bad argument / call order                           */
#define C3_ERR_CUDA 2    /* CUDA runtime error                                  */
#define C3_ERR_MEMORY 3  /* device memory exhausted                             */
#define C3_ERR_NUMERIC 4 /* recoverable numeric failure -> ArithmeticError      */

typedef struct c3_engine c3_engine;

typedef struct c3_stats {
  int64_t evaluations;        /* c3_evaluate* calls that launched work          */
  int64_t kernel_launches;    /* kernels launched since creation                */
  int64_t last_launches;      /* kernels launched by the last evaluation        */
  int64_t last_expm_jobs;     /* P matrices recomputed by the last evaluation   */
  int64_t last_nodes;         /* internal nodes recomputed by the last eval     */
  int64_t last_node_rows;     /* sum of pattern rows over those nodes           */
  int64_t last_alg_bytes;     /* algorithmic bytes of the pruning kernels (§8d) */
  int64_t last_flops;         /* executed FP64 flops of the pruning kernels     */
  int64_t last_h2d_bytes;     /* host->device bytes copied by the last evaluation */
  int64_t device_bytes;       /* device memory owned by the engine              */
  int64_t partial_bytes;      /* ... of which resident partial likelihoods      */
  int64_t graph_replays;      /* evaluations issued as ONE cudaGraphLaunch      */
  int32_t n_levels;           /* tree levels (kernel launches of a full sweep)  */
  int32_t padded_states;      /* row stride of a partial row, in doubles        */
  int64_t wave_launches;      /* sweeps issued as ONE wavefront launch (all levels at once) */
  int64_t wave_entries;       /* schedule entries of the last wavefront launch  */
  int64_t wave_noops;         /* ... of which no-op positions                   */
  int64_t lean_evaluations;   /* evaluations whose root launch wrote the class mixture (8 B per row) instead of lh */
  int64_t chain_launches;     /* single-path updates walked per root row by the chain kernel (no level barriers) */
  int64_t walk_launches;      /* other evaluations of small problems walked per root row (tree-walk kernel)     */
} c3_stats;

C3_API int c3_abi_version(void);
C3_API const char* c3_last_error(void);
C3_API int c3_device_count(int* n);

/* ---- construction --------------------------------------------------------
 * replaces the Defn sub-DAG built by make_total_loglikelihood_defn
 * (evolve/likelihood_calculation.py:125-167): one engine per (locus) likelihood
 * function.  child_ptr[n_nodes+1] / child_idx[] is the CSR list of ordered
 * children (PhyloNode.children order, likelihood_calculation.py:83-87).        */
C3_API int c3_create(int device, int n_nodes, int n_states, int n_bins,
              const int32_t* child_ptr, const int32_t* child_idx, c3_engine** out);
C3_API int c3_destroy(c3_engine* e);
/* run on the caller's CUDA stream (e.g. torch's current stream); NULL = a stream of
 * the engine's own.  The legacy default stream is named by its CUDA handle
 * cudaStreamLegacy ((void*)1), not by NULL.                                      */
C3_API int c3_set_stream(c3_engine* e, void* cuda_stream);

/* ---- pattern tables (set once per alignment) -------------------------------
 * c3_set_tip: LikelihoodTreeLeaf.input_likelihoods [n_rows x M]
 *             (evolve/likelihood_tree.py:223-278; LeafPartialLikelihoodDefn
 *             likelihood_calculation.py:32-37)
 * c3_set_node_patterns: LikelihoodTreeEdge.indexes int64 [C x n_rows],
 *             C-contiguous (evolve/likelihood_tree.py:55-58)
 * c3_set_root_counts: root LikelihoodTreeEdge.counts (likelihood_tree.py:62)
 * c3_finalize: allocates the resident partial-likelihood arrays in HBM
 *             (make_partial_likelihoods_array, likelihood_tree.py:133-134).    */
C3_API int c3_set_tip(c3_engine* e, int node, const double* table, int64_t n_rows);
C3_API int c3_set_node_patterns(c3_engine* e, int node, const int64_t* indexes, int64_t n_rows);
C3_API int c3_set_root_counts(c3_engine* e, const double* counts, int64_t n_rows);
C3_API int c3_finalize(c3_engine* e);

/* ---- pattern compression on the device -------------
 * c3_compress_alignment: builds EVERY table above from the raw alignment, on the
 *             GPU, bit-identical to the reference's: per tip the distinct symbol
 *             codes in first-seen column order (make_likelihood_tree_leaf,
 *             evolve/likelihood_tree.py:223-278, `_indexed` :186-203), per internal
 *             node the distinct tuples of child pattern numbers in first-seen
 *             order (LikelihoodTreeEdge.__init__ :13-70), root counts (:62), each
 *             table followed by its all-gap row.  It replaces the c3_set_tip /
 *             c3_set_node_patterns / c3_set_root_counts calls; c3_finalize follows.
 *             codes: uint8 [n_tips][n_columns] (host), tips in node order;
 *             symbol_rows: float64 [n_symbols][M], the 0/1 row of every code
 *             (get_matched_array, likelihood_tree.py:206-220).  keep_index != 0
 *             keeps every node's column numbering on the device (tests); the
 *             root's (LikelihoodTreeEdge.index) is always kept.
 * c3_get_node_rows / _indexes / c3_get_tip_table / c3_get_root_counts /
 * c3_get_column_index: read the tables back (uniq = indexes transposed; `index`).
 * c3_set_column_index: the root's `index` when the tables came from the host
 *             (enables c3_get_site_lh for the drop-in).                           */
C3_API int c3_compress_alignment(c3_engine* e, const uint8_t* codes, int64_t n_columns,
                                 const double* symbol_rows, int n_symbols, int keep_index);
C3_API int c3_get_node_rows(c3_engine* e, int64_t* n_rows_of_node);
C3_API int c3_get_node_indexes(c3_engine* e, int node, int64_t* indexes);
C3_API int c3_get_tip_table(c3_engine* e, int node, double* table);
C3_API int c3_get_root_counts(c3_engine* e, double* counts);
C3_API int c3_get_column_index(c3_engine* e, int node, int64_t* index);
C3_API int c3_set_column_index(c3_engine* e, const int64_t* index, int64_t n_columns);

/* ---- model state -------------------------------------------------------------
 * A "q slot" is one distinct rate matrix (one value of the Q / Qd Defns,
 * evolve/substitution_model.py:636-641).
 * c3_set_eigen: the EigenExponentiator state (maths/matrix_exponentiation.py:23-40)
 *             as produced by Fast/CheckedExponentiator (:132-146): roots[M],
 *             evT[M x M], evI[M x M]; complex128 interleaved (re, im) when
 *             is_complex != 0, float64 otherwise.
 * c3_set_Q:   the rate matrix for the Pade path (PadeExponentiator :92-129).
 * c3_set_edge_qslots: which slot each (edge, bin) exponentiates [n_nodes x K]; a
 *             negative entry leaves that (edge, bin) unchanged (e.g. a supplied P).
 * c3_set_psub: a P matrix supplied directly (discrete-time models,
 *             evolve/discrete_markov.py:15-80); the (edge, bin) stops using expm.
 * c3_set_distances: distance[edge, bin] = length * rate (ProductDefn,
 *             recalculation/definition.py:679-683; psubs = Qd(distance),
 *             substitution_model.py:702-710).  The engine diffs against the
 *             previous values: only changed edges are re-exponentiated and only
 *             their ancestors re-pruned (the role of Calculator.change,
 *             recalculation/calculation.py:402-477).
 * c3_set_root_mprobs / c3_set_bin_probs / c3_set_fixed_motifs: the other inputs
 *             of the lh / tll / plh Defns (likelihood_calculation.py:50-59,
 *             147-148, 174-185).                                               */
C3_API int c3_set_eigen(c3_engine* e, int q_slot, const double* roots, const double* evT,
                 const double* evI, int is_complex);
C3_API int c3_set_Q(c3_engine* e, int q_slot, const double* Q);
/* c3_set_tn93: the closed-form P of the F81 / HKY85 / TN93 family for models built with
 *             rate_matrix_required=False (PredefinedNucleotide.calc_psub_matrix, evolve/solved_models.py:43-49
 *             -> calc_TN93_P, evolve/solved_models_numba.py:9-51): pi[4] in T, C, A, G order, alpha1 = kappa_y,
 *             alpha2 = kappa_r (1, 1 for F81; kappa, kappa for HKY85).  P(t) is then computed on the device
 *             for every (edge, bin) that uses the slot, like the eigen and Pade slots.                       */
C3_API int c3_set_tn93(c3_engine* e, int q_slot, const double* pi, double alpha1, double alpha2);
C3_API int c3_set_edge_qslots(c3_engine* e, const int32_t* q_slot_of_edge_bin);
C3_API int c3_set_psub(c3_engine* e, int node, int bin, const double* P);
C3_API int c3_set_distances(c3_engine* e, const double* distance_of_edge_bin);
C3_API int c3_set_root_mprobs(c3_engine* e, const double* pi);
C3_API int c3_set_bin_probs(c3_engine* e, const double* bprobs);
C3_API int c3_set_fixed_motifs(c3_engine* e, const int32_t* fixed_motif_of_node);
C3_API int c3_invalidate(c3_engine* e); /* force a full recomputation next time        */

/* ---- per-pattern underflow rescaling ----------------------------------------------
 * The reference multiplies plain FP64 partials on this path and has no rescaling
 * (evolve/likelihood_tree.py:10,151-153; its only scaling, BASE = 2**100, is in the
 * phylo-HMM reducer :168-176): past a few hundred divergent taxa a pattern's
 * likelihood underflows to 0 and lnL is -inf.  The engine can carry a per-pattern
 * power-of-two exponent instead: the rows of "scaled" nodes (every subtree that has
 * gathered tip_threshold unscaled tips) are divided by 2^ilogb(largest element)
 * and the integer exponents are summed up the tree and folded back in the
 * reduction (log(mix) + E ln 2) and in c3_get_lh / c3_get_site_lh (ldexp).  Powers
 * of two are exact, so wherever the reference does not underflow the per-pattern
 * values are bit-identical with and without rescaling.
 *   C3_RESCALE_OFF   never (the reference's arithmetic, -inf included)
 *   C3_RESCALE_ON    always
 *   C3_RESCALE_AUTO  (default) plain products until a blocking c3_evaluate / c3_update /
 *                    c3_evaluate_many returns a non-finite lnL; the engine then switches
 *                    rescaling on for good and repeats that evaluation.  The
 *                    asynchronous c3_evaluate_device cannot look at the value: callers
 *                    that use it switch modes themselves.
 * tip_threshold <= 0 keeps the current value (default 16).  While rescaling is active,
 * c3_get_partials returns the stored (scaled) rows.
 * c3_get_rescaling: whether rescaling is active, and the number of scaled nodes.
 * c3_get_root_exponents: E[p] of every root pattern (all zero when inactive).
 * c3_get_scaled_lh: like c3_get_lh but WITHOUT folding the exponent back: the true value is
 *             out[p] * 2^E[p], the same E for every bin of a pattern -- what a caller needs
 *             for ratios (posterior bin probabilities, likelihood_calculation.py:247-257)
 *             and logs of patterns whose likelihood is below the FP64 range.             */
#define C3_RESCALE_OFF 0
#define C3_RESCALE_ON 1
#define C3_RESCALE_AUTO 2
C3_API int c3_set_rescaling(c3_engine* e, int mode, int tip_threshold);
C3_API int c3_get_rescaling(c3_engine* e, int* active, int* n_scaled_nodes);
C3_API int c3_get_root_exponents(c3_engine* e, int32_t* out, int64_t n_rows);
C3_API int c3_get_scaled_lh(c3_engine* e, int bin, double* out, int64_t n_rows);

/* ---- evaluation ----------------------------------------------------------------
 * c3_evaluate: batched expm for the dirty (edge, bin)s, level-by-level pruning of
 *             the dirty nodes, site-weighted reduction; blocks for the scalar
 *             (replaces Calculator.__call__'s cell sweep for the psubs / inner /
 *             plh / lh / tll cells and get_log_sum_across_sites,
 *             evolve/likelihood_tree_numba.py:36-42).
 * c3_evaluate_device: same, asynchronous on the engine's stream; the scalar is
 *             written to device memory (input of the NCCL all-reduce when site
 *             patterns are sharded over GPUs).
 * c3_update:  distances + evaluate in one call (one FFI crossing per optimiser
 *             step).                                                           */
C3_API int c3_evaluate(c3_engine* e, double* lnL);
C3_API int c3_evaluate_device(c3_engine* e, double* device_lnL);
C3_API int c3_update(c3_engine* e, const double* distance_of_edge_bin, double* lnL);

/* ---- site-pattern shards on the GPUs of one node ----------------------
 * Each rank (one process per GPU) holds a column block of the alignment in its own engine
 * and the log-likelihood is the sum over ranks (the multi-locus / multi-cpu sum of
 * SumDefn, recalculation/definition.py:672-676; likelihood_calculation.py:140-145).  The
 * sum is FUSED into the reduction kernel: its last CTA stores the shard value into every
 * rank's mailbox over NVLink peer memory, waits for the other shards' values and adds them
 * in rank order, so c3_evaluate returns the TOTAL on every rank without a collective call.
 * c3_comm_mailbox: allocate this rank's mailbox; returns its device pointer and / or its
 *             64-byte CUDA IPC handle (to be all-gathered by the host, any transport).
 * c3_comm_attach: the peers' mailboxes, as IPC handles [nranks][64] (other processes) or as
 *             device pointers valid in this process (several engines in one process).
 *             Every rank must then evaluate the same number of times.
 * c3_comm_detach: back to shard-local values.                                           */
C3_API int c3_comm_mailbox(c3_engine* e, int nranks, void** device_ptr, unsigned char* ipc_handle);
C3_API int c3_comm_attach(c3_engine* e, int nranks, int rank, const unsigned char* ipc_handles,
                          void* const* device_ptrs);
C3_API int c3_comm_detach(c3_engine* e);
/* c3_evaluate_many: c3_evaluate for n independent engines (each on its own stream): all are
 *             launched before any is waited for, so small alignments -- the app layer's
 *             workload, hundreds of independent likelihood functions (app/evo.py:317-386,
 *             evolve/bootstrap.py:68) -- overlap on the GPU instead of paying the launch /
 *             level latency one after the other.                          */
C3_API int c3_evaluate_many(c3_engine** engines, int n, double* lnL);
/* c3_update_many: c3_set_distances for every engine (engine i reads distances + i * stride;
 *             stride 0: the same table for all), then c3_evaluate_many -- one FFI crossing per
 *             batch.  Engines that keep being evaluated together with the same structure are
 *             captured into ONE CUDA graph (every engine on its own stream inside it).        */
C3_API int c3_update_many(c3_engine** engines, int n, const double* distances, int64_t stride, double* lnL);

/* ---- accessors (what likelihood_function.py reads back) -------------------------
 * c3_get_lh: per-pattern root likelihood of one bin, "lh" Defn
 *             (likelihood_function.py:415-425); n_rows = root rows.  The values are those of
 *             the LAST evaluation.  On large problems the root launch only keeps the class
 *             mixture (nothing on the hot path reads lh[p][K]); the first accessor that asks
 *             (c3_get_lh, c3_get_site_lh, c3_get_scaled_lh, c3_hmm_reduce) re-runs that one
 *             launch with the lh output -- nothing is committed, pending input changes stay
 *             pending -- and the next few evaluations write lh themselves.
 * c3_get_psub: "psubs" value (likelihood_function.py:272-320).
 * c3_get_partials: "plh" of one node and bin [n_rows x M].                     */
C3_API int c3_get_lh(c3_engine* e, int bin, double* out, int64_t n_rows);
C3_API int c3_get_psub(c3_engine* e, int node, int bin, double* out);
C3_API int c3_get_partials(c3_engine* e, int node, int bin, double* out, int64_t n_rows);
C3_API int c3_get_stats(c3_engine* e, c3_stats* out);
/* c3_get_site_lh: per-COLUMN likelihoods of one bin, lh[root.index] gathered on the
 *             device (LikelihoodTreeEdge.get_full_length_likelihoods,
 *             evolve/likelihood_tree.py:93-94; likelihood_function.py:427-431).   */
C3_API int c3_get_site_lh(c3_engine* e, int bin, double* out, int64_t n_columns);

/* ---- phylo-HMM: the reduction over the SITES ------------------------
 * c3_hmm_reduce: with sites_independent=False the per-bin root likelihoods of the last evaluation are
 *             reduced by a hidden Markov model over two "patches" of bins instead of being mixed per
 *             pattern: plhs[site][patch] = sum_{b in patch} weight_b lh_b[index[site]]
 *             (PatchSiteDistribution.get_weighted_sum_lhs, evolve/likelihood_calculation.py:209-216),
 *             state <- (switch . state) * plhs[site] over all sites with BASE = 2**100 rescaling, lnL =
 *             log(sum(state)) + exponent log(BASE) (SiteHmm.__call__ :265-269; log_dot_reduce,
 *             evolve/likelihood_tree.py:168-176 -- a Python loop over the columns).  On the device: a
 *             parallel product of the 2 x 2 matrices diag(plhs[site]) . switch in site order.
 *             patch_of_bin[K] in {0, 1}, weight_of_bin[K], switch_matrix[2 x 2] row-major
 *             (SiteClassTransitionMatrix.Matrix), patch_probs[2] (its StationaryProbs).           */
C3_API int c3_hmm_reduce(c3_engine* e, int n_patches, const int32_t* patch_of_bin, const double* weight_of_bin,
                         const double* switch_matrix, const double* patch_probs, double* lnL);

/* ---- per-launch timing (CUDA events on the engine's stream) ----------------------
 * With profiling on, every kernel launch of a blocking c3_evaluate is bracketed by
 * events.  c3_get_launch_profile returns, for the last evaluation and per launch:
 * kind (0 expm-eigen, 1 expm-pade, 2 prune level, 3 prune root level, 4 reduction, 5 the fused
 * small-levels launch, which may hold the root and the reduction too),
 * duration in ms, algorithmic bytes and executed flops of that launch.            */
C3_API int c3_set_profiling(c3_engine* e, int on);
C3_API int c3_get_launch_profile(c3_engine* e, int max_n, int32_t* kind, double* ms,
                                 int64_t* alg_bytes, int64_t* flops, int* n);

/* ---- calibration helper: sustained FP64 tensor (DMMA) rate of this GPU, used as
 * the roofline denominator of the 61-state kernel.                 */
C3_API int c3_measure_dmma_peak(int device, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* C3CUDA_H */
