/* gnngls_b200 — C ABI of the B200-native gnngls inference hot path.
 *
 * The reference (proroklab/gnngls) is pure Python and has no FFI of its own; this header is
 * the boundary a maintainer would bind (ctypes stub in INTEGRATION.md) to replace, function by
 * function, the reference code cited beside each entry point (paths relative to the reference
 * repository root).
 *
 * Conventions
 *  - every pointer is a caller-owned DEVICE pointer unless the name ends in _host;
 *  - `stream` is a cudaStream_t passed as void*; all entry points are asynchronous w.r.t. the
 *    host and never allocate device memory (workspaces are caller-provided, sizes are queried);
 *  - return value: 0 on success, negative gnngls_status on failure; a human-readable message
 *    for the calling thread's last failure is available from gnngls_last_error_string();
 *  - tours are int32[n+1] with tour[0] == tour[n] == 0 (the depot), distance matrices are
 *    row-major fp64 [n,n]; batches are dense leading dimensions.
 */
#ifndef GNNGLS_B200_H_
#define GNNGLS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GNNGLS_B200_ABI_VERSION 1

typedef enum gnngls_status {
    GNNGLS_OK = 0,
    GNNGLS_ERR_BAD_ARG = -1,
    GNNGLS_ERR_UNSUPPORTED = -2,
    GNNGLS_ERR_CUDA = -3,
    GNNGLS_ERR_WORKSPACE = -4
} gnngls_status;

typedef enum gnngls_move_op {
    GNNGLS_OP_TWO_OPT = 0,   /* gnngls/operators.py:6-29  */
    GNNGLS_OP_RELOCATE = 1   /* gnngls/operators.py:76-103 */
} gnngls_move_op;

/* How a guide (edge attribute used for the GLS utility / nearest-neighbour init) is stored. */
typedef enum gnngls_guide_kind {
    GNNGLS_GUIDE_MATRIX_F64 = 0,  /* [B, n_guides, n, n] fp64, symmetric                          */
    GNNGLS_GUIDE_EDGEVEC_F32 = 1  /* [B, n_guides, n(n-1)/2] fp32, line-graph node order (i<j);    *
                                   * widened to fp64 on read, as scripts/test.py:81-83 does         */
} gnngls_guide_kind;

/* per-instance status bits written by the search kernels */
#define GNNGLS_INST_EVENTS_TRUNCATED 1   /* more events than max_events; count is still exact */
#define GNNGLS_INST_PENALTY_OVERFLOW 2
#define GNNGLS_INST_STALLED 4            /* perturbation loop hit the safety cap (reference would spin) */

int gnngls_abi_version(void);
const char *gnngls_last_error_string(void);

/* ---------------------------------------------------------------------------------------------
 * Move evaluation — gnngls/operators.py:32-50 (two_opt_a2a), :129-147 (relocate_a2a),
 * :53-73 (two_opt_o2a), :106-126 (relocate_o2a).
 * One CTA per instance -- or, when the batch is far below the SM count (B <= SMs, n >= 48), one
 * thread-block cluster of up to 16 CTAs per instance whose members share the rows of the scan and
 * exchange their winners through distributed shared memory (a2a sweeps, local_search, GLS and the
 * nearest-neighbour constructor; results are bit-identical, the choice is made per launch;
 * GNNGLS_CLUSTER=0 disables it).  fp64 deltas in the reference's association order, no FMA; the
 * winner equals the reference's sequential first-strict-minimum scan.
 *   D            [B,n,n] fp64, or one [n,n] matrix shared by the batch when d_batch_stride == 0
 *   tours        [B,n+1] int32
 *   pos          [B] int32 (o2a only): the fixed index i, 0 < i < n
 *   out_delta    [B] fp64  : best delta, or 0.0 when no improving move exists
 *   out_move     [B,2] int32: (i,j) of the chosen move, (-1,-1) when none
 *   out_tours    [B,n+1] int32 or NULL: tour after applying the move (copy of input when none)
 * ------------------------------------------------------------------------------------------- */
int gnngls_moves_eval_a2a(int op, const double *D, int64_t d_batch_stride, const int32_t *tours,
                          int B, int n, int first_improvement,
                          double *out_delta, int32_t *out_move, int32_t *out_tours, void *stream);

int gnngls_moves_eval_o2a(int op, const double *D, int64_t d_batch_stride, const int32_t *tours,
                          const int32_t *pos, int B, int n, int first_improvement,
                          double *out_delta, int32_t *out_move, int32_t *out_tours, void *stream);

/* ---------------------------------------------------------------------------------------------
 * local_search — gnngls/algorithms.py:111-132.  tours/costs are updated in place.
 *   events   [B,max_events] fp64 or NULL: cost after every accepted move (search_progress['cost'])
 *   n_events [B] int32 or NULL;  status [B] int32 or NULL (GNNGLS_INST_* bits)
 *   counters [B,4] int64 or NULL: {two-opt sweeps, relocate sweeps, o2a scans, accepted moves}
 * ------------------------------------------------------------------------------------------- */
int gnngls_local_search_batch(const double *D, int32_t *tours, double *costs, int B, int n,
                              int first_improvement, double *events, int32_t *n_events,
                              int max_events, int32_t *status, int64_t *counters, void *stream);

/* ---------------------------------------------------------------------------------------------
 * guided_local_search — gnngls/algorithms.py:135-195, persistent on device, one CTA per
 * instance, with the wall-clock test of :146 replaced by an explicit range of outer iterations
 * so that callers can either run a fixed count or poll the clock between chunks.
 * ------------------------------------------------------------------------------------------- */
typedef struct gnngls_gls_args {
    int32_t B, n;
    const double *D;            /* [B,n,n]                                                        */
    int32_t guide_kind;         /* gnngls_guide_kind                                              */
    int32_t n_guides;           /* guide used in outer iteration it is guides[it % n_guides]      */
    const void *guides;
    int32_t *cur_tours;         /* [B,n+1] in: init_tour (or state when resume); out: current     */
    double *cur_costs;          /* [B]     in: init_cost (or state);             out: current     */
    int32_t *best_tours;        /* [B,n+1] out (in/out when resume)                               */
    double *best_costs;         /* [B]     out (in/out when resume)                               */
    double *k;                  /* [B] out when !resume (0.1*init_cost/n, :137), in when resume   */
    int32_t *penalties;         /* [B,n,n] int32 in/out; may be NULL when !resume (not persisted);
                                   the L2-resident and cluster tiers need it                      */
    int32_t resume;             /* 0: zero penalties, run the initial local_search (:138-143)     */
    int32_t iter_begin;         /* index of the first outer iteration of this call                */
    int32_t n_iters;            /* number of outer iterations to run                              */
    int32_t perturbation_moves;
    int32_t first_improvement;
    double *events;             /* [B,max_events] or NULL                                          */
    int32_t *n_events;          /* [B] or NULL: total events of THIS call                         */
    int32_t max_events;
    int32_t *status;            /* [B] or NULL                                                    */
    int64_t *counters;          /* [B,4] or NULL (see gnngls_local_search_batch)                  */
} gnngls_gls_args;

int gnngls_gls_batch(const gnngls_gls_args *args_host, void *stream);
/* sizeof(gnngls_gls_args), for FFI bindings to verify their struct mirror */
size_t gnngls_sizeof_gls_args(void);

/* nearest_neighbor + tour_cost — gnngls/algorithms.py:9-18, gnngls/__init__.py:17-21 as used at
 * scripts/test.py:85-90.  Greedy tour from `depot` on the guide (first minimum in ascending
 * node id wins), then the sequential fp64 cost of that tour under D.
 *   guide: guide_kind layout with n_guides == 1;  D may be NULL (then out_costs is not written) */
int gnngls_nn_init_batch(int guide_kind, const void *guide, const double *D, int B, int n, int depot,
                         int32_t *out_tours, double *out_costs, void *stream);

/* gnngls/__init__.py:17-21 for a batch of tours. */
int gnngls_tour_cost_batch(const double *D, const int32_t *tours, int B, int n, double *out_costs,
                           void *stream);

/* ---------------------------------------------------------------------------------------------
 * Edge-regret model — gnngls/models.py:44-70 (+ dgl.nn.GATConv reached from models.py:23).
 * Activations are fp32 row-major [M,128]; M = total line-graph nodes of the batch.
 * ------------------------------------------------------------------------------------------- */
#define GNNGLS_EMBED_DIM 128
#define GNNGLS_HEADS 8
#define GNNGLS_HEAD_DIM 16
#define GNNGLS_HIDDEN_DIM 512

/* input construction: datasets.py:14-20 + MinMaxScaler.transform (datasets.py:85) for line-graph
 * node v = rank(i,j):  x0 = float(D[b,i,j]);  x1 = float(double(x0)*scale);  x = float(double(x1)+min_)
 * (sklearn applies `X *= scale_; X += min_` in place on the float32 array with fp64 scalars: each
 * step is evaluated in fp64 and rounded to fp32 — verified against sklearn in tests). */
int gnngls_edge_features(const double *D, int B, int n, double scale, double min_, float *x, void *stream);

/* storage format of a tensor-core operand in HBM: the projected features ft[M,128] handed from fc to the
 * aggregates, and the operand copies of the activations (below) */
typedef enum gnngls_ft_dtype {
    GNNGLS_FT_F32 = 0,    /* fp32, unrounded (pure-fp32 debug path)                                   */
    GNNGLS_FT_TF32 = 1,   /* fp32 storage, values rounded to TF32 (10-bit mantissa)                   */
    GNNGLS_FT_F16 = 2     /* IEEE fp16 (same 10-bit mantissa, half the bytes; saturates at +-65504)   */
} gnngls_ft_dtype;

/* embed_layer (models.py:57,66): h[M,128] = x[M,in_dim] * W[128,in_dim]^T + b
 * h_op (nullable): operand copy of h for the first fc, as op_dtype says (GNNGLS_FT_TF32 or GNNGLS_FT_F16). */
int gnngls_embed_forward(const float *x, int64_t M, int in_dim, const float *W, const float *b,
                         float *h, void *h_op, int op_dtype, void *stream);

/* dense implementation selector for the fc / feed-forward contractions */
typedef enum gnngls_dense_impl {
    GNNGLS_DENSE_TCGEN05 = 0,   /* TMA-fed tcgen05.mma kind::tf32, accumulators in TMEM (default) */
    GNNGLS_DENSE_SIMT = 1,      /* plain fp32 CUDA-core kernel: debug cross-check only            */
    GNNGLS_DENSE_TCGEN05_F16 = 2 /* tcgen05.mma kind::f16: weights (and the fc input) passed as fp16  */
} gnngls_dense_impl;

/* GATConv.fc + attention scores (Appendix A of SURVEY.md):
 *   ft[M,128] = h * Wfc[128,128]^T ; el[M,8] = log2(e) * sum_f ft*attn_l ; er[M,8] = log2(e) * sum_f ft*attn_r
 * The scores are stored in the log2 domain because the aggregates evaluate the edge softmax with ex2;
 * leaky_relu is positively homogeneous so softmax(leaky_relu(el+er)) is unchanged.  el/er are always
 * computed from the unrounded fp32 accumulators; `ft` is stored as `ft_dtype` says (its only consumer
 * is the aggregate, whose tensor-core operand has a 10-bit mantissa anyway).
 * h / Wfc: fp32 (GNNGLS_DENSE_SIMT: exact; GNNGLS_DENSE_TCGEN05: the TF32-rounded copies) or fp16
 * (GNNGLS_DENSE_TCGEN05_F16: the fp16 operand copy written by embed / feed-forward, fp16 weights).    */
int gnngls_fc_forward(int impl, const void *h, int64_t M, const void *Wfc, const float *attn_l,
                      const float *attn_r, void *ft, int ft_dtype, float *el, float *er, void *stream);

/* Per-channel affine form of eval-mode BatchNorm1d: y = x*scale + shift
 * (scale = gamma/sqrt(var+eps), shift = beta - mean*scale; models.py:27,35).
 *
 * *_tf32 / *_op outputs (embed / aggregate / feed-forward, all nullable): a second copy of the produced
 * activation for the next tensor-core GEMM: fp32 rounded to TF32 (cvt.rna), or fp16 (same mantissa, half the bytes).  The tcgen05 kind::tf32 MMA ignores the low 13 mantissa
 * bits of its operands (truncation, a biased error ~10x larger than rounding over this 8-layer
 * model); feeding it the pre-rounded copy makes that truncation exact, while skip connections keep
 * reading the unrounded fp32 activation.  Pass NULL on the pure-fp32 debug path. */

/* GAT aggregate over an arbitrary destination-sorted CSR graph + skip + BatchNorm1 (models.py:12-15,27):
 *   h1[v] = BN1(h[v] + sum_u softmax_u(leaky_relu(el[u]+er[v])) ft[u] + gat_bias)               */
int gnngls_gat_aggregate_csr(const int32_t *indptr, const int32_t *indices, int64_t M,
                             const void *ft, int ft_dtype, const float *el, const float *er, const float *h,
                             const float *gat_bias /* [128] or NULL */, const float *bn_scale,
                             const float *bn_shift, float *h1, float *h1_tf32, void *stream);

/* Same op for a batch of line graphs of K_n with the adjacency computed arithmetically
 * (neighbours of (i,j) are (i,k) and (k,j)); M = B*n(n-1)/2.  `workspace` must hold
 * gnngls_gat_kn_workspace_bytes(B,n) bytes. */
size_t gnngls_gat_kn_workspace_bytes(int B, int n);
int gnngls_gat_aggregate_kn(int B, int n, const void *ft, int ft_dtype, const float *el, const float *er,
                            const float *h, const float *gat_bias, const float *bn_scale,
                            const float *bn_shift, float *h1, float *h1_tf32,
                            void *workspace, size_t workspace_bytes, void *stream);

/* feed-forward block (models.py:28-35):
 *   h_out = BN2(h1 + W2 * relu(W1 * h1 + b1) + b2),  W1[512,128], W2[128,512]
 * h1_tf32 (nullable) is a pre-rounded operand copy for the first contraction; when NULL the tensor-core
 * kernel rounds h1 to TF32 itself while staging the tile into tensor memory.  The skip always reads h1.
 * GNNGLS_DENSE_TCGEN05_F16: W1/W2 are IEEE fp16 (same row-major shapes), h1 is packed to fp16 while it is staged
 * into tensor memory (h1_tf32 is ignored) and the hidden activations are kept as fp16 pairs: same 10-bit
 * mantissa as the TF32 variant, half the MMAs and half the streamed weight bytes; values saturate at +-65504.
 * `workspace` must hold gnngls_ff_workspace_bytes(impl, M) bytes (may be 0). */
size_t gnngls_ff_workspace_bytes(int impl, int64_t M);
int gnngls_ff_forward(int impl, const float *h1, const float *h1_tf32, int64_t M, const void *W1,
                      const float *b1, const void *W2, const float *b2, const float *bn_scale,
                      const float *bn_shift, float *h_out, void *h_out_op, int op_dtype, void *workspace,
                      size_t workspace_bytes, void *stream);

/* decision_layer (models.py:63,69): y[M,out_dim] = h * Wd[out_dim,128]^T + bd */
int gnngls_decision_forward(const float *h, int64_t M, int out_dim, const float *Wd, const float *bd,
                            float *y, void *stream);

/* scripts/test.py:79-83: MinMaxScaler.inverse_transform on the float32 array then clamp:
 *   r1 = float(double(y) - min_);  r2 = float(double(r1) / scale);  regret = max(r2, 0)
 * In place allowed. */
int gnngls_regret_postprocess(const float *y, int64_t M, double scale, double min_, float *regret, void *stream);

/* ---- whole model forward for a batch of line graphs of K_n behind ONE call ---------------------------------------
 * EdgePropertyPredictionModel.forward (gnngls/models.py:65-70): embed -> n_layers x [GATConv fc, edge-softmax aggregate +
 * skip + BN1, feed-forward + skip + BN2] -> decision, i.e. the sequence of the entry points above, issued from C on
 * `stream` with no host round trip in between (26 kernels for the shipped 8-layer model).  Weights are caller-owned
 * device arrays in the formats the per-op entry points take: for GNNGLS_DENSE_TCGEN05_F16 Wfc/W1/W2 are fp16, for
 * GNNGLS_DENSE_TCGEN05 fp32 rounded to TF32, for GNNGLS_DENSE_SIMT plain fp32. */
typedef struct gnngls_layer_params {
    const void *Wfc;                    /* [128,128]  GATConv.fc.weight                                   */
    const float *attn_l, *attn_r;       /* [128]      GATConv.attn_l / attn_r, head-major                 */
    const float *gat_bias;              /* [128] or NULL (DGL >= 0.7 checkpoints)                         */
    const float *bn1_scale, *bn1_shift; /* [128]      eval-mode BatchNorm1d after the message passing     */
    const void *W1;                     /* [512,128]                                                      */
    const float *b1;                    /* [512]                                                          */
    const void *W2;                     /* [128,512]                                                      */
    const float *b2;                    /* [128]                                                          */
    const float *bn2_scale, *bn2_shift; /* [128]                                                          */
} gnngls_layer_params;

typedef struct gnngls_model_args {
    int32_t B, n;                       /* batch of line graphs of K_n: M = B*n(n-1)/2 nodes              */
    int32_t in_dim, out_dim, n_layers;
    int32_t dense_impl;                 /* gnngls_dense_impl                                              */
    int32_t ft_dtype;                   /* gnngls_ft_dtype of the projected features                      */
    int32_t reserved;
    const float *x;                     /* [M,in_dim]                                                     */
    const float *We, *be;               /* embed_layer  [128,in_dim], [128]                               */
    const float *Wd, *bd;               /* decision_layer [out_dim,128], [out_dim]                        */
    const gnngls_layer_params *layers;  /* HOST array of n_layers entries (device pointers inside)        */
    float *y;                           /* [M,out_dim]                                                    */
} gnngls_model_args;

size_t gnngls_sizeof_model_args(void);
size_t gnngls_model_forward_workspace_bytes(int B, int n, int dense_impl);
int gnngls_model_forward(const gnngls_model_args *args, void *workspace, size_t workspace_bytes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* GNNGLS_B200_H_ */
