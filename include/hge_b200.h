/*
 * hge_b200.h -- C ABI of the B200-native HOBE hot path (libhge_b200.so).
 *
 * The reference (JSybrandt/HypergraphEmbedding) is pure Python and has no FFI of its own
 * (SURVEY.md section 8b): its boundary is a set of Python callables.  This header is the
 * C-level boundary those callables bind to in this build; every entry point names the
 * reference function (file:line, relative to the reference tree) whose arithmetic it
 * replaces.  The Python mirror of the reference interface lives in
 * hypergraphembedding_b200/{algebraic_distance,hg2v_weighting,hg2v_sample}.py and calls
 * these symbols through ctypes (INTEGRATION.md shows the stub).
 *
 * Conventions
 *   - Every function returns int: 0 = HGE_OK, < 0 = error.  hge_last_error() returns a
 *     thread-local message for the last failing call.  Nothing throws across the boundary.
 *   - Pointers are plain C pointers.  "mem" arguments say where they live:
 *     HGE_MEM_HOST (the library stages through its own device buffers, copies included) or
 *     HGE_MEM_DEVICE (borrowed device pointers, e.g. torch tensor storage; the caller keeps
 *     them alive until hge_ctx_sync()).
 *   - Row pointers are int64, column ids int32, values fp32, row-major.
 *   - One context per device; calls on one context are not re-entrant.
 *   - There is no CPU fallback: without a usable CUDA device every compute entry point
 *     fails with HGE_ERR_CUDA.
 */
#ifndef HGE_B200_H_
#define HGE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HGE_OK 0
#define HGE_ERR_INVALID (-1)     /* bad argument (the reference would assert)            */
#define HGE_ERR_CUDA (-2)        /* CUDA runtime / launch failure, or no device          */
#define HGE_ERR_EMPTY_ROW (-3)   /* a node or edge without incidences: the reference     */
                                 /* divides 0/0 there (algebraic_distance.py:49)         */
#define HGE_ERR_NOMEM (-4)
#define HGE_ERR_UNSUPPORTED (-5)

#define HGE_MEM_HOST 0
#define HGE_MEM_DEVICE 1

typedef struct hge_ctx hge_ctx;
typedef struct hge_incidence hge_incidence;
typedef struct hge_algdist hge_algdist;

/* ---- library / context ------------------------------------------------------------ */

int hge_version(void);
const char* hge_last_error(void);

/* stream: the cudaStream_t every launch and copy of this context is queued on (e.g. torch's
 * current stream); NULL is the legacy default stream.  The caller owns the stream. */
int hge_ctx_create(int device, void* stream, hge_ctx** out);
int hge_ctx_destroy(hge_ctx* ctx);
int hge_ctx_set_stream(hge_ctx* ctx, void* stream);
int hge_ctx_sync(hge_ctx* ctx);
/* Tuning knobs of the relaxation schedule (0 keeps the default):
 *   light_max_deg  rows up to this degree are gathered by one sub-warp (<= 255)
 *   chunk          incidences per warp work item for longer rows
 *   blocks_per_sm  persistent grid size = SMs * blocks_per_sm */
int hge_ctx_set_tuning(hge_ctx* ctx, int light_max_deg, int chunk, int blocks_per_sm);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
int64_t hge_ctx_launch_count(const hge_ctx* ctx);

/* ---- incidence storage ------------------------------------------------------------
 * Replaces ToCsrMatrix / ToEdgeCsrMatrix (hypergraph_util.py:96-135): the node->edge
 * incidence as int32 CSR (n2e) and the edge->node incidence as int32 CSR (e2n, i.e. the
 * CSC of the same matrix when the hypergraph is consistent).  Column ids must be sorted
 * and unique inside a row, as scipy's canonical CSR is.  Builds the degree-binned gather
 * schedule and the inverse neighbour-weight sums used by the relaxation.
 * Returns HGE_ERR_EMPTY_ROW if any of the N nodes / E edges has no incidence. */
int hge_incidence_create(hge_ctx* ctx, int32_t num_nodes, int32_t num_edges,
                         const int64_t* n2e_ptr, const int32_t* n2e_idx,
                         const int64_t* e2n_ptr, const int32_t* e2n_idx, int mem,
                         hge_incidence** out);
int hge_incidence_destroy(hge_incidence* inc);
int64_t hge_incidence_nnz(const hge_incidence* inc);

/* ---- algebraic-distance relaxation --------------------------------------------------
 * Replaces the loop of EmbedAlgebraicDistance (algebraic_distance.py:149-164):
 * `iterations` x { node half (_update_alg_dist, :34-51, old edge rows), edge half (new node
 * rows), joint per-column min/max rescale (_helper_scale_embeddings, :97-123) }.
 * xn [N, R] and xe [E, R] hold the initial vectors on entry (algebraic_distance.py:140-141,
 * drawn by the caller) and the rescaled result on return, fp32 row-major, dense.
 * lohi (optional, host, [iterations][2][R]) receives each sweep's per-column (min, max). */
int hge_algdist_run(hge_ctx* ctx, hge_incidence* inc, float* xn, float* xe, int R,
                    int iterations, int mem, float* lohi);

/* Stepwise form of the same loop, used when the node rows are sharded over several GPUs
 * and the host interleaves collectives (DESIGN.md "Multi-GPU").  All pointers are device
 * pointers. */
int hge_algdist_create(hge_ctx* ctx, hge_incidence* inc, int R, int max_iterations,
                       hge_algdist** out);
int hge_algdist_destroy(hge_algdist* st);
int hge_algdist_load(hge_algdist* st, const float* xn, const float* xe, int mem);
int hge_algdist_node_half(hge_algdist* st, int sweep);
int hge_algdist_edge_half(hge_algdist* st, int sweep);
/* Sharded edge half: writes the un-normalised local sums  sum_{n local} w_n * xn'[n]  for
 * every edge into partial [E, ld]; after the caller's all-reduce(sum) the second call
 * blends, rescales and stores the edge rows.  ld = hge_algdist_ld(). */
int hge_algdist_edge_partial(hge_algdist* st, int sweep, float* partial);
int hge_algdist_edge_finalize(hge_algdist* st, int sweep, const float* partial,
                              const float* inv_s_edge_global);
/* Device pointer to this sweep's order-preserving int32 encoded (min[ld], max[ld]) slots,
 * for an all-reduce(MIN) / (MAX) across shards. */
int hge_algdist_minmax_ptr(hge_algdist* st, int sweep, int32_t** out);
int hge_algdist_ld(const hge_algdist* st);
int hge_algdist_store(hge_algdist* st, int sweeps_done, float* xn, float* xe, int mem);

#ifdef __cplusplus
}
#endif
#endif /* HGE_B200_H_ */
