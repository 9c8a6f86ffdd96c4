/*
 * dgcn.h - C ABI of the B200-native message-passing hot path of deep_gcns_torch.
 *
 * The reference (lightaime/deep_gcns_torch) has no FFI / plugin layer: its
 * boundary is the Python nn.Module API.  This header is the
 * boundary a binding for that API talks to.  Every entry point names the
 * reference code it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the name ends in _host;
 *   - tensors are fp32, dense, in the reference's own layouts:
 *       dense path   x (B, C, N, 1)  -> element (b,c,n) at x[b*stride_b + c*stride_c + n]
 *                    edge_index (2, B, N, k) int64, plane 0 = neighbour j, plane 1 = centre i
 *       sparse path  x (N, C) row-major, edge_index (2, E) int64 row 0 = source, row 1 = target
 *   - `stream` is a cudaStream_t; nothing synchronises the host, nothing
 *     allocates: scratch memory is a caller-owned workspace sized by the
 *     matching *_workspace_bytes() query;
 *   - return value: DGCN_OK or a negative dgcn_status; no exceptions cross.
 *   - kernels are compiled for sm_100a only.
 */
#ifndef DGCN_H_
#define DGCN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* dgcn_stream_t; /* cudaStream_t */

enum dgcn_status {
  DGCN_OK = 0,
  DGCN_ERR_BAD_ARG = -1,      /* null pointer / negative or inconsistent size */
  DGCN_ERR_UNSUPPORTED = -2,  /* valid request outside what the kernels cover */
  DGCN_ERR_WORKSPACE = -3,    /* workspace too small                          */
  DGCN_ERR_CUDA = -4          /* launch failed; see dgcn_last_cuda_error()    */
};

/* gcn_lib/dense/torch_nn.py:9-21 (act_layer) */
enum dgcn_act { DGCN_ACT_NONE = 0, DGCN_ACT_RELU = 1, DGCN_ACT_LEAKYRELU = 2, DGCN_ACT_PRELU = 3 };
/* gcn_lib/dense/torch_nn.py:24-33 (norm_layer): BatchNorm2d in eval (running
 * statistics) or train (batch statistics) mode.  InstanceNorm2d cannot be
 * constructed through the reference's BasicConv (torch_nn.py:71 dereferences a
 * None weight), so it is not part of the path. */
enum dgcn_norm { DGCN_NORM_NONE = 0, DGCN_NORM_BATCH_EVAL = 1, DGCN_NORM_BATCH_TRAIN = 2 };
/* gcn_lib/dense/torch_vertex.py:44-49 (GraphConv2d dispatch) */
enum dgcn_conv { DGCN_CONV_EDGE = 0, DGCN_CONV_MR = 1 };
/* gcn_lib/sparse/torch_message.py:44-85 (GenMessagePassing.aggregate) */
enum dgcn_aggr {
  DGCN_AGGR_SOFTMAX = 0,     /* 'softmax' and 'softmax_sg' (identical forward) */
  DGCN_AGGR_SOFTMAX_SUM = 1,
  DGCN_AGGR_POWER = 2,
  DGCN_AGGR_POWER_SUM = 3,
  DGCN_AGGR_ADD = 4,
  DGCN_AGGR_MEAN = 5,
  DGCN_AGGR_MAX = 6
};

int dgcn_version(void);
const char* dgcn_status_string(int status);
/* text of the last CUDA error seen by this thread ("" if none) */
const char* dgcn_last_cuda_error(void);

/* ------------------------------------------------------------------------
 * Dense path
 * --------------------------------------------------------------------- */

/* Parameters of `BasicConv([2*C_in, C_out], act, norm, bias)`
 * (gcn_lib/dense/torch_nn.py:48-72; state_dict keys nn.0.weight/bias,
 * nn.<i>.weight/bias/running_mean/running_var). */
typedef struct dgcn_basic_conv {
  const float* weight;       /* (C_out, 2*C_in) row-major = Conv2d 1x1 weight */
  const float* bias;         /* (C_out) or NULL                               */
  int32_t act;               /* dgcn_act                                      */
  float slope;               /* leakyrelu negative slope (reference: 0.2)     */
  const float* prelu_weight; /* device scalar for DGCN_ACT_PRELU, else NULL   */
  int32_t norm;              /* dgcn_norm                                     */
  const float* bn_weight;    /* gamma (C_out) or NULL (=1)                    */
  const float* bn_bias;      /* beta  (C_out) or NULL (=0)                    */
  const float* bn_mean;      /* running mean (eval)                           */
  const float* bn_var;       /* running var  (eval)                           */
  float bn_eps;              /* 1e-5                                          */
  float* batch_mean_out;     /* train: batch mean (C_out) out, may be NULL    */
  float* batch_var_out;      /* train: biased batch variance (C_out) out      */
} dgcn_basic_conv;

/* Which ranks of the sorted neighbour list survive
 * (gcn_lib/dense/torch_edge.py:19-29, DenseDilated): rank l*dilation for
 * l < k, or - stochastic branch - the k ranks listed in cols_host (a HOST
 * array drawn by the caller from the CPU generator, torch_edge.py:22-24). */
/* flags: DGCN_KNN_EXACT_FP32 ranks with the fp32 FMA kernels only (no tensor-core
 * pre-filter); DGCN_KNN_TC_TILE_PER_CTA keeps the tensor-core pre-filter on the
 * one-tile-per-CTA kernel where the four-tile warp-specialised kernel would be chosen.
 * The result is the same list either way - the switches exist for A/B tests and
 * measurements and are a property of the CALL, not of the process. */
enum dgcn_knn_flags {
  DGCN_KNN_DEFAULT = 0,
  DGCN_KNN_EXACT_FP32 = 1,
  DGCN_KNN_TC_TILE_PER_CTA = 2 /* tensor-core path: always the one-tile-per-CTA kernel, never the four-tile one */
};
typedef struct dgcn_dilation {
  int64_t k;
  int64_t dilation;
  const int32_t* cols_host; /* NULL or k entries, each in [0, k*dilation) */
  int32_t flags;            /* dgcn_knn_flags */
  int32_t reserved;         /* must be 0 */
} dgcn_dilation;

/* Dilated kNN graph of every cloud of a batch.
 * Replaces gcn_lib/dense/torch_edge.py:32-58 (pairwise_distance +
 * dense_knn_matrix) and :61-76 (DenseDilatedKnnGraph.forward); with
 * exclude_self != 0 it replaces :79-101 (DilatedKnnGraph over
 * torch_cluster.knn_graph, loop=False).
 * Ranking: ascending D = (|x_i|^2 + (-2 x_i.x_j)) + |x_j|^2 evaluated in fp32,
 * ties broken towards the smaller j.  No (B,N,N) matrix is materialised for
 * k*dilation <= 64; above that one L2-sized slab of rows lives in `ws`.
 * Outputs (either may be NULL): edge_index (2,B,N,k) int64 exactly as the
 * reference returns it; nbr (B,N,k) int32 = plane 0 only, for the fused
 * consumers below. */
size_t dgcn_knn_graph_workspace_bytes(int64_t B, int64_t C, int64_t N, int64_t K);
int dgcn_knn_graph(const float* x, int64_t B, int64_t C, int64_t N, int64_t stride_b,
                   int64_t stride_c, const dgcn_dilation* dil, int32_t exclude_self,
                   int64_t* edge_index, int32_t* nbr, void* ws, size_t ws_bytes,
                   dgcn_stream_t stream);

/* Static graph convolution on a given graph.
 * Replaces gcn_lib/dense/torch_vertex.py:38-52 (GraphConv2d.forward) =
 * EdgeConv2d.forward :31-35 / MRConv2d.forward :16-20, including
 * batched_index_select (gcn_lib/dense/torch_nn.py:75-96) and BasicConv
 * (torch_nn.py:48-58: conv1x1 -> act -> norm) and the max over neighbours.
 * The graph comes either as the reference's int64 edge_index (2,B,N,k) with
 * arbitrary centres, or as nbr (B,N,k) int32 with centre = own index.
 * out: (B, C_out, N) contiguous. */
size_t dgcn_graph_conv_workspace_bytes(int32_t conv, int64_t B, int64_t C_in, int64_t C_out,
                                       int64_t N, int64_t k);
int dgcn_graph_conv_forward(int32_t conv, const float* x, int64_t B, int64_t C_in, int64_t N,
                            int64_t stride_b, int64_t stride_c, const int64_t* edge_index,
                            const int32_t* nbr, int64_t k, const dgcn_basic_conv* p,
                            int64_t C_out, float* out, void* ws, size_t ws_bytes,
                            dgcn_stream_t stream);

/* Dynamic graph convolution: dilated kNN graph on x, then the convolution, in
 * one call; the neighbour list goes from the selection kernel's shared memory
 * straight into the gather/max and never reaches HBM unless nbr_out != NULL.
 * Replaces gcn_lib/dense/torch_vertex.py:55-72 (DynConv2d.forward with
 * knn='matrix'). */
size_t dgcn_dyn_conv_workspace_bytes(int32_t conv, int64_t B, int64_t C_in, int64_t C_out,
                                     int64_t N, int64_t K);
int dgcn_dyn_conv_forward(int32_t conv, const float* x, int64_t B, int64_t C_in, int64_t N,
                          int64_t stride_b, int64_t stride_c, const dgcn_dilation* dil,
                          const dgcn_basic_conv* p, int64_t C_out, float* out, int32_t* nbr_out,
                          void* ws, size_t ws_bytes, dgcn_stream_t stream);

/* Block epilogue around the dynamic convolution, inference (no train-mode BatchNorm):
 *   residual != NULL: out = conv(x) + residual * res_scale   (ResDynBlock2d.forward, gcn_lib/dense/torch_vertex.py:101;
 *     residual (B, C_out, N) with strides res_stride_b / res_stride_c, unit stride along points)
 *   out_stride_b != 0: batch stride of `out` in floats - `out` is a channel slice of a wider (B, C_total, N) buffer
 *     (DenseDynBlock2d's torch.cat, :116, or a model's fusion buffer, examples/sem_seg_dense/architecture.py:52). */
typedef struct dgcn_block_fusion {
  const float* residual;
  int64_t res_stride_b, res_stride_c;
  float res_scale;
  int64_t out_stride_b;
} dgcn_block_fusion;
int dgcn_dyn_conv_forward_fused(int32_t conv, const float* x, int64_t B, int64_t C_in, int64_t N,
                                int64_t stride_b, int64_t stride_c, const dgcn_dilation* dil,
                                const dgcn_basic_conv* p, int64_t C_out, float* out, int32_t* nbr_out,
                                const dgcn_block_fusion* fus /* may be NULL */, void* ws, size_t ws_bytes,
                                dgcn_stream_t stream);

/* Gradient of dgcn_graph_conv_forward w.r.t. x and the BasicConv parameters
 * (what torch autograd derives for torch_vertex.py:16-35 + torch_nn.py:48-58).
 * The graph is the one used in forward (edge_index (2,B,N,k) int64 or nbr
 * (B,N,k) int32 with centre = own index; the kNN graph itself is
 * non-differentiable, torch_edge.py:53-56).  With DGCN_NORM_BATCH_TRAIN,
 * p->bn_mean / p->bn_var must hold the BATCH statistics the forward returned.
 * grad_x is (B, C_in, N) contiguous.
 * grad_weight (C_out,2*C_in), grad_bias (C_out), grad_bn_weight/bias (C_out),
 * grad_prelu (1) are OVERWRITTEN; any of them may be NULL. */
size_t dgcn_graph_conv_backward_workspace_bytes(int32_t conv, int64_t B, int64_t C_in,
                                                int64_t C_out, int64_t N, int64_t k);
int dgcn_graph_conv_backward(int32_t conv, const float* x, int64_t B, int64_t C_in, int64_t N,
                             int64_t stride_b, int64_t stride_c, const int64_t* edge_index,
                             const int32_t* nbr, int64_t k,
                             const dgcn_basic_conv* p, int64_t C_out, const float* grad_out,
                             float* grad_x, float* grad_weight, float* grad_bias,
                             float* grad_bn_weight, float* grad_bn_bias, float* grad_prelu,
                             void* ws, size_t ws_bytes, dgcn_stream_t stream);

/* ------------------------------------------------------------------------
 * Sparse path
 * --------------------------------------------------------------------- */

/* One-time COO -> CSR-by-destination build of the graph the reference keeps
 * implicit in edge_index (PyG propagate, gcn_lib/sparse/torch_vertex.py:68).
 * Stable: within a destination row, edges keep their edge_index order.
 * rowptr (N+1) int32, src (E) int32 = source node per CSR slot, eid (E) int32 =
 * position of that edge in edge_index (for edge_attr lookup). */
size_t dgcn_csr_build_workspace_bytes(int64_t N, int64_t E);
int dgcn_csr_build(const int64_t* edge_index, int64_t E, int64_t N, int32_t* rowptr,
                   int32_t* src, int32_t* eid, void* ws, size_t ws_bytes, dgcn_stream_t stream);

/* Scalars that the reference keeps either as python floats or as (1,)
 * nn.Parameters (gcn_lib/sparse/torch_message.py:19-40): if the _dev pointer
 * is non-NULL the kernel reads the device scalar, else the host value. */
typedef struct dgcn_genconv_params {
  int32_t aggr;             /* dgcn_aggr */
  float t; const float* t_dev;
  float p; const float* p_dev;
  float y; const float* y_dev;   /* *_sum only: out *= deg^sigmoid(y) */
  float eps;                /* message eps, 1e-7 (torch_vertex.py:26,85) */
  int32_t msg_norm;         /* MsgNorm on/off (torch_message.py:88-99) */
  float msg_scale; const float* msg_scale_dev;
  int32_t add_residual;     /* 1: out = x + m (torch_vertex.py:73); 0: out = m */
  int32_t raw_message;      /* 1: msg_e = x_src[src_e] as is (GenMessagePassing.aggregate on
                               explicit messages, torch_message.py:44); 0: relu(.)+eps */
} dgcn_genconv_params;

/* Fused GENConv message + aggregate + MsgNorm + residual:
 *   msg_e = relu(x[src_e] + edge_attr[eid_e]) + eps     torch_vertex.py:78-85
 *   m_i   = aggregate_{e -> i}(msg_e)                   torch_message.py:44-85
 *   m_i   = scale * |x_i|_2 * m_i / max(|m_i|_2, 1e-12) torch_message.py:95-99
 *   out_i = x_i + m_i                                   torch_vertex.py:73
 * x_src (N_src, C) holds the rows that sources index (on one GPU the same array
 * as x_dst; under node partitioning local rows followed by halo rows);
 * x_dst (N, C) the rows of the destinations this call owns (may be NULL when neither
 * msg_norm nor add_residual is set).
 * edge_attr (E, C) in edge_index order or NULL.  out (N, C). */
/* Long rows ("hubs" of power-law graphs): rows of in-degree >= min_degree are cut into segments of
 * seg_edges edges.  dgcn_csr_hub_rows lists them once per graph, entirely on the device:
 *   items  (2 * max_items int32): (row, segment) pairs, max_items = E / seg_edges + N_hub <= E/seg + E/min_degree + 1
 *   rows   (3 * max_rows  int32): (row, first item, #segments), max_rows <= E / min_degree + 1
 *   counts (2 int32): number of items, number of rows.
 * Given to dgcn_genconv_aggregate (with `partial`, a scratch of items * 3 * C floats) those rows are
 * aggregated by one CTA per segment plus a fixed-order merge instead of one warp per row. */
typedef struct dgcn_csr_hubs {
  const int32_t* items;
  const int32_t* rows;
  const int32_t* counts;
  int32_t min_degree;
  int32_t seg_edges;
  float* partial;
} dgcn_csr_hubs;
int dgcn_csr_hub_rows(const int32_t* rowptr, int64_t N, int64_t E, int32_t min_degree, int32_t seg_edges,
                      int32_t* items, int32_t* rows, int32_t* counts, dgcn_stream_t stream);

int dgcn_genconv_aggregate(const float* x_src, const float* x_dst, int64_t N, int64_t C,
                           const int32_t* rowptr, const int32_t* src, const int32_t* eid,
                           const float* edge_attr, const dgcn_genconv_params* prm,
                           const dgcn_csr_hubs* hubs /* may be NULL */, float* out,
                           dgcn_stream_t stream);

/* Block fusion around the aggregate (DeeperGCN 'res+' block,
 * examples/ogb/ogbn_arxiv/model.py:91-106: h <- GENConv(relu(norm(h))) + h) and the split launch
 * that lets a node-partitioned layer overlap its halo exchange:
 *   pre_scale / pre_shift (C) or both NULL: every row read from x_src and x_dst is taken as
 *     act(pre_scale * x + pre_shift) (eval-mode BatchNorm1d folded to an affine, act = relu when
 *     pre_relu != 0), so the normalised / activated copy of h is never written to HBM;
 *   row_list (n_rows int32) or NULL: destination rows this launch processes (NULL = all N rows):
 *     interior rows (all sources local) first, boundary rows once the halo has arrived;
 *   skip_hubs != 0: rows of degree >= hubs->min_degree are left to a later launch. */
typedef struct dgcn_genconv_fusion {
  const float* pre_scale;
  const float* pre_shift;
  int32_t pre_relu;
  int32_t skip_hubs;
  const int32_t* row_list;
  int64_t n_rows;
} dgcn_genconv_fusion;
int dgcn_genconv_aggregate_fused(const float* x_src, const float* x_dst, int64_t N, int64_t C,
                                 const int32_t* rowptr, const int32_t* src, const int32_t* eid,
                                 const float* edge_attr, const dgcn_genconv_params* prm,
                                 const dgcn_csr_hubs* hubs /* may be NULL */,
                                 const dgcn_genconv_fusion* fus /* may be NULL */, float* out,
                                 dgcn_stream_t stream);

/* Row-wise Linear with fused bias and skip connection on the tcgen05 tensor cores:
 *   out[n][m] = sum_k a[n][k] * weight[m][k] (+ bias[m]) (+ res[n][m])
 * = the Linear that ends GENConv's MLP (gcn_lib/sparse/torch_nn.py:56-68 with mlp_layers = 1; nn.Linear
 * weight layout (M, K)) plus the `+ h` of DeeperGCN's res+ block (examples/ogb/ogbn_arxiv/model.py:104).
 * fp32 in / out; the product runs as a two-plane bf16 split of both operands (4 tensor-core products, fp32
 * accumulation in TMEM, error <= ~2^-16 * sum_k |a||w|).  K in {64, 128, 256}, M a multiple of 32 up to 256,
 * 16-byte aligned rows; anything else returns DGCN_ERR_UNSUPPORTED (callers keep cuBLAS for those).
 * bias, res may be NULL; out may alias res.  ws: dgcn_linear_residual_workspace_bytes(K, M) (0 = unsupported). */
size_t dgcn_linear_residual_workspace_bytes(int64_t K, int64_t M);
int dgcn_linear_residual(const float* a, int64_t N, int64_t K, const float* weight, const float* bias,
                         int64_t M, const float* res, float* out, void* ws, size_t ws_bytes,
                         dgcn_stream_t stream);

/* Gradient of dgcn_genconv_aggregate w.r.t. x (both roles), edge_attr and the
 * scalar parameters.  The softmax weights carry gradient only when
 * softmax_grad != 0 (reference: learn_t, torch_message.py:51-55).
 * grad_x_src (N_src, C) must be zero-initialised by the caller (rows are
 * accumulated with atomics); grad_x_dst (N, C) is overwritten and may alias
 * nothing.  grad_scalars (4) = d/dt, d/dp, d/dy, d/dmsg_scale, accumulated with
 * atomics into a zero-initialised array; may be NULL. */
int dgcn_genconv_aggregate_backward(const float* x_src, const float* x_dst, int64_t N,
                                    int64_t N_src, int64_t C, const int32_t* rowptr,
                                    const int32_t* src, const int32_t* eid,
                                    const float* edge_attr, const dgcn_genconv_params* prm,
                                    int32_t softmax_grad, const float* grad_out,
                                    float* grad_x_src, float* grad_x_dst, float* grad_edge_attr,
                                    float* grad_scalars, dgcn_stream_t stream);

/* Halo packing for node-partitioned graphs (new functionality; the reference
 * has no multi-GPU sparse path): out[r,:] = x[rows[r],:]. */
int dgcn_gather_rows(const float* x, int64_t C, const int32_t* rows, int64_t R, float* out,
                     dgcn_stream_t stream);

/* ------------------------------------------------------------------------
 * Measurement hook (bench.py): when enabled, the dominant kernel of each path
 * ("knn" = the fused selection kernel(s), "aggregate" = the GENConv kernel) is
 * bracketed by CUDA events on the launch stream.  _read() synchronises those
 * events, returns the summed duration and launch count for `tag` and resets it.
 * --------------------------------------------------------------------- */
int dgcn_debug_kernel_timing(int32_t enable);
int dgcn_debug_kernel_timing_read(const char* tag, double* total_ms, int64_t* launches);

#ifdef __cplusplus
}
#endif
#endif /* DGCN_H_ */
