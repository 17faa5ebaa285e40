/*
 * b3d.h — C ABI of libb3d.so: the B200 (sm_100a) kernels behind the Batch3DMOT
 * tracking-graph GNN hot path.
 *
 * The reference (robot-learning-freiburg/Batch3DMOT) has no FFI layer for this
 * path: its boundary is the Python layer API in batch_3dmot/models/pose_gnn.py
 * and clr_att_gnn.py, which reaches device code only through PyTorch /
 * torch_geometric / torch_scatter / torch_cluster calls. Each entry point below
 * names the reference call site(s) it replaces (file:line relative to the
 * reference tree). The host-side binding is batch3dmot_b200/_lib.py (ctypes);
 * INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless
 *     stated otherwise; matrices are row-major fp32 with an explicit leading
 *     dimension (in elements);
 *   - every call enqueues work on `stream` (a cudaStream_t passed as void*),
 *     never allocates, never synchronises; scratch memory is passed in by the
 *     caller and sized by the *_workspace_bytes functions;
 *   - return value 0 = success, otherwise a cudaError_t (or -1 for an argument
 *     error); b3d_last_error() returns a static description;
 *   - results are deterministic: no floating-point atomics anywhere.
 */
#ifndef B3D_H_
#define B3D_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B3D_MAX_SEGS 8

/* Operand modifiers for b3d_seg_t.mask_mode / epilogue activations. */
enum { B3D_MASK_NONE = 0, B3D_MASK_RELU = 1, B3D_MASK_SIGMOID = 2 };
enum { B3D_ACT_NONE = 0, B3D_ACT_RELU = 1, B3D_ACT_SIGMOID = 2,
       B3D_ACT_MASKBITS = 3 };      /* b3d_chain_run only: multiply by a ReLU mask given as sign bits (backward chains) */
enum { B3D_NARROW_INPUT_RELU = 0x100 };   /* flag of b3d_narrow_mlp_*'s final_act, see there */
enum { B3D_FLAG_ACCUMULATE = 1,     /* out += result instead of out = result */
       B3D_FLAG_OUT_BF16 = 2,       /* b3d_segment_sum: `out` is really __nv_bfloat16* (bf16 source, no accumulate) */
       B3D_FLAG_SPLIT = 4 };        /* the tensor-core arithmetic of the 1e-4 parity mode. b3d_linear_tc: "tf32 x3" — every fp32
                                       operand enters the tensor core as hi + lo tf32 words (22 significant bits) and each K
                                       step issues hi*hi + lo*hi + hi*lo with kind::tf32 (~2^-21 relative per product, fp32
                                       accumulation); weights packed with B3D_PACK_SPLIT. b3d_wgrad_tc: the same scheme with
                                       bf16 hi + lo parts (~2^-16; gradients are held to 1e-3). */
enum { B3D_PACK_TRANSPOSE = 1, B3D_PACK_SPLIT = 2 };   /* bits of b3d_tc_pack_weights' `transpose` argument; a split pack
                                                          needs 4 x b3d_tc_packed_bytes */
enum { B3D_F32 = 0, B3D_BF16 = 1 }; /* element type of a segment / output (bf16: tensor-core entry points only) */
/* mask_dtype only: the ReLU mask of a layer output as SIGN BITS, uint32 words [ceil(N/32)][M]
 * (word (c / 32) * M + r holds columns 32*(c/32) .. +31 of row r, bit c % 32 set iff output > 0).
 * Written by the forward tensor-core layers (relu_bits_out), read by the input-gradient layers and
 * by b3d_gather_rows: 1/16 of the bytes of the bf16 activation and one coalesced word per row. */
enum { B3D_BITS = 2 };
/* x_dtype of the narrow MLP chains only: float64 input rows (edge_attr is stored as float64 and cast with
 * `.float()` at pose_gnn.py:67 / clr_att_gnn.py:123; the cast is folded into the row load). */
enum { B3D_F64 = 3 };

/* One column block of a (virtually) concatenated, optionally row-gathered
 * operand: rows r = 0..M-1 read ptr[(idx ? idx[r] : r) * ld + 0..width-1].
 * Replaces torch.cat([...], dim=1) of index_select'ed tensors
 * (pose_gnn.py:210,215,222; clr_att_gnn.py:161-163,314,319,326) and PyG's
 * MessagePassing.__collect__ gathers (pose_gnn.py:180).
 * mask (optional, same row addressing as ptr but leading dimension ldmask)
 * turns the value v into  v * (mask > 0)            for B3D_MASK_RELU,
 *                         v * mask * (1 - mask)     for B3D_MASK_SIGMOID,
 * which is how the backward pass of ReLU / Sigmoid is folded into the GEMMs. */
typedef struct {
  const float* ptr;
  const int32_t* idx;
  const float* mask;
  int32_t width;
  int32_t ld;
  int32_t ldmask;
  int32_t mask_mode;
  int32_t dtype;   /* B3D_F32 (ptr is float*) or B3D_BF16 (ptr is really __nv_bfloat16*, ld in elements) */
} b3d_seg_t;

const char* b3d_last_error(void);
/* Number of kernels launched by this library since the last reset (host counter). */
int64_t b3d_launch_count(void);
void b3d_reset_launch_count(void);

/* ---- graph plumbing ------------------------------------------------------
 * One-time stable LSD radix sort of edge_index into CSC (by target,
 * edge_index[1]) and CSR (by source, edge_index[0]). Replaces the atomics of
 * torch_scatter.scatter(reduce='add') (pose_gnn.py:240, :190-191;
 * clr_att_gnn.py:344, :293-294) by segment tables. Bit-exact spec:
 * perm == torch.argsort(index, stable=True), rowptr == cumsum(bincount).
 * edge_index: int64 [2,E] contiguous. src32/dst32: int32 [E] copies.
 * status: int32[1], set non-zero if an index is outside [0,N). */
size_t b3d_csr_workspace_bytes(int64_t E, int64_t N);
int b3d_csr_build(const int64_t* edge_index, int64_t E, int64_t N,
                  int32_t* src32, int32_t* dst32,
                  int32_t* rowptr_dst, int32_t* perm_dst,
                  int32_t* rowptr_src, int32_t* perm_src,
                  void* workspace, size_t workspace_bytes, int32_t* status, void* stream);

/* src_dtype: B3D_F32 or B3D_BF16 (bf16 rows are accumulated in fp32).
 * out[n, 0:C] (+)= sum_{k in [rowptr[n], rowptr[n+1])} src[(perm ? perm[k] : k), 0:C]
 * summed in ascending k (== edge order within the node, as sequential CPU
 * scatter_add_ does). Warp-per-node segmented reduction.
 * Replaces torch_scatter.scatter(reduce='add') pose_gnn.py:190-191,240. */
int b3d_segment_sum(const void* src, int32_t src_dtype, int32_t ld_src, const int32_t* perm,
                    const int32_t* rowptr, int64_t N, int32_t C, float* out, int32_t ld_out, int32_t flags,
                    void* stream);

/* out[r, 0:C] = src[idx[r], 0:C]  (index_select; backward of segment_sum). out_dtype B3D_F32 or
 * B3D_BF16 (rounded on store). relu_mask (optional, bf16 [M,C], bf16 output only): the summed rows
 * were ReLU outputs, so the gathered gradient is zeroed where relu_mask <= 0. */
int b3d_gather_rows(const float* src, int32_t ld_src, const int32_t* idx, int64_t M, int32_t C,
                    void* out, int32_t out_dtype, int32_t ld_out, const void* relu_mask, int32_t ld_mask,
                    int32_t mask_dtype /* bf16 output: B3D_BF16 or B3D_BITS; fp32 output: B3D_F32 */, int32_t src_dtype /* B3D_F32, or B3D_BF16
                    (src is really __nv_bfloat16*, ld in elements; bf16 output only) */, void* stream);

/* out[M,C] = sum_k ins[k][M,C] (n <= 8 dense fp32/bf16 inputs of equal width C % 8 == 0, each with its
 * own leading dimension; fp32 accumulation in argument order). One pass for the gradient of a tensor
 * with several consumers: the edge features feed edge_update of the next iteration and both message
 * MLPs (pose_gnn.py:210,215,222), att_edge_attr feeds all six iterations (clr_att_gnn.py:186). */
int b3d_add_n(const b3d_seg_t* ins /*host*/, int32_t n, int64_t M, void* out, int32_t out_dtype,
              int32_t ldo, void* stream);

/* ---- dense layers with fused gather / concat / activation ------------------
 * Y[M,Nout] (+)= act( cat_s(A_s)[M,K] * op(W) + bias ) [* (out_mask > 0)] [row_mask]
 *   trans_w == 0: W is [Nout,K] row-major (nn.Linear.weight), Y = A W^T   (forward)
 *   trans_w == 1: W is [K,Nout] row-major,                   Y = A W     (input gradient)
 * Replaces nn.Linear/addmm + torch.cat + index_select + ReLU/Sigmoid chains
 * (pose_gnn.py:29-53,94-120,210-223; clr_att_gnn.py:35-91,196-222,314-327).
 * out_mask (optional [M,Nout], ld ldm): multiply result by (out_mask > 0) (ReLU backward
 * of the producing layer). row_mask (optional uint8[M]): rows with 0 are written as 0
 * (clr_att_gnn.py:132-133,140-141 zero-fill of missing modalities).
 * adds (optional, nadd <= 2, fp32, width == Nout): row-gathered addends summed into the
 * pre-activation, result = act(A W^T + bias + sum_a adds[a][idx_a[r], :]). This is how the
 * node-side blocks of a first-layer weight are applied per NODE instead of per edge
 * (W [x_i|x_j|e] . cat[x_i,x_j,e] = W_xi x[dst] + W_xj x[src] + W_e e). */
int b3d_linear(const b3d_seg_t* segs /*host*/, int32_t nseg, const float* W, int32_t ldw,
               int32_t trans_w, const float* bias, float* Y, int32_t ldy, int64_t M, int32_t Nout,
               int32_t act, int32_t flags, const float* out_mask, int32_t ldm,
               const uint8_t* row_mask, const b3d_seg_t* adds /*host*/, int32_t nadd, void* stream);

/* dW[Nout,K] (+)= dY^T cat_s(A_s),  db[Nout] (+)= colsum(dY)   (weight gradient)
 * dy: a single segment (may carry a ReLU/Sigmoid mask). Deterministic split over rows
 * with a fixed-order second pass; workspace from b3d_wgrad_workspace_bytes. */
size_t b3d_wgrad_workspace_bytes(int64_t M, int32_t Nout, int32_t K);
int b3d_wgrad(const b3d_seg_t* dy /*host*/, const b3d_seg_t* segs /*host*/, int32_t nseg,
              float* dW, int32_t lddw, float* db, int64_t M, int32_t Nout, int32_t flags,
              void* workspace, size_t workspace_bytes, void* stream);

/* ---- bf16 tensor-core variants (tcgen05.mma, fp32 accumulate in TMEM; 2e-2 tolerance mode) ----
 * Same math as b3d_linear / b3d_wgrad with operands rounded to bf16. Segments may be fp32 (converted
 * while staged into shared memory) or bf16 (copied with cp.async); operand masks are not supported
 * (the backward pass pre-masks gradients in the producing kernel's epilogue via out_mask).
 * Requirements: every segment width is a multiple of 8 and rows are 16-byte aligned; otherwise -1
 * is returned and the caller uses the fp32 entry point.
 * Weights are packed once per weight version into bf16 K-major slabs:
 *   transpose == 0: B[n][k] = W[n*ldw + k]  (forward;  n_logical = out_features, k_logical = in_features)
 *   transpose == 1: B[n][k] = W[k*ldw + n]  (input gradient; n_logical = in_features, k_logical = out_features)
 * Y / out_mask element types are given by y_dtype / mask_dtype (B3D_F32 or B3D_BF16). */
size_t b3d_tc_packed_bytes(int32_t n_logical, int32_t k_logical);
int b3d_tc_pack_weights(const float* W, int32_t ldw, int32_t n_logical, int32_t k_logical,
                        int32_t transpose, void* Wp, void* stream);
int b3d_linear_tc(const b3d_seg_t* segs /*host*/, int32_t nseg, const void* Wp, int32_t n_logical,
                  int32_t k_logical, const float* bias, void* Y, int32_t ldy, int32_t y_dtype, int64_t M,
                  int32_t act, int32_t flags, const void* out_mask, int32_t ldm, int32_t mask_dtype,
                  const uint8_t* row_mask, const b3d_seg_t* adds /*host*/, int32_t nadd,
                  void* relu_bits_out /* optional uint32 [ceil(N/32)][M] */, void* stream);
/* TMA-fed persistent variant of b3d_linear_tc for DENSE bf16 operands (1-2 row-major segments, no
 * gather; the leading segment's width a multiple of 64): TMA producer warp, single-thread tcgen05
 * issuer, 4 epilogue warps, weight block resident in shared memory, double-buffered TMEM
 * accumulators. Weights are packed row-major bf16 [Npad][Kpad] by b3d_tma_pack_weights. */
/* b3d_linear_tma takes 1-6 bf16 segments; a segment with `idx` is ROW-GATHERED by the TMA unit itself
 * (cp.async.bulk.tensor tile::gather4: four table rows per request into the swizzled operand tile) — this is how
 * the reference's x[edge_index[1]] / x[edge_index[0]] operands (MessagePassing.__collect__, pose_gnn.py:180;
 * torch.cat at :210,215,222 / clr_att_gnn.py:314,319,326) enter the tensor core without an epilogue gather. At most
 * two distinct index arrays per launch; gathered tables may hold up to 2^28 rows. Every segment occupies whole
 * 64-column chunks of the packed weights: pack with b3d_tma_pack_weights_segs(widths) unless only the last segment
 * is ragged (then b3d_tma_pack_weights gives the same image). */
size_t b3d_tma_packed_bytes_segs(int32_t n_logical, const int32_t* widths /*host*/, int32_t nseg);
int b3d_tma_pack_weights_segs(const float* W, int32_t ldw, int32_t n_logical, const int32_t* widths /*host*/,
                              int32_t nseg, int32_t transpose, void* Wr, void* stream);
size_t b3d_tma_packed_bytes(int32_t n_logical, int32_t k_logical);
int b3d_tma_pack_weights(const float* W, int32_t ldw, int32_t n_logical, int32_t k_logical,
                         int32_t transpose, void* Wr, void* stream);
int b3d_linear_tma(const b3d_seg_t* segs /*host*/, int32_t nseg, const void* Wr, int32_t n_logical,
                   int32_t k_logical, const float* bias, void* Y, int32_t ldy, int32_t y_dtype, int64_t M,
                   int32_t act, int32_t flags, const void* out_mask, int32_t ldm, int32_t mask_dtype,
                   const uint8_t* row_mask, const b3d_seg_t* adds /*host*/, int32_t nadd,
                   void* relu_bits_out /* optional */, void* stream);
size_t b3d_wgrad_tc_workspace_bytes(int64_t M, int32_t Nout, int32_t K);
/* TMA-fed variant of b3d_wgrad_tc for DENSE bf16 dy / segments (see b3d_linear_tma). */
size_t b3d_wgrad_tma_workspace_bytes(int64_t M, int32_t Nout, int32_t K);
int b3d_wgrad_tma(const b3d_seg_t* dy /*host*/, const b3d_seg_t* segs /*host*/, int32_t nseg,
                  float* dW, int32_t lddw, float* db, int64_t M, int32_t Nout, int32_t flags,
                  void* workspace, size_t workspace_bytes, void* stream);
int b3d_wgrad_tc(const b3d_seg_t* dy /*host*/, const b3d_seg_t* segs /*host*/, int32_t nseg,
                 float* dW, int32_t lddw, float* db, int64_t M, int32_t Nout, int32_t flags,
                 void* workspace, size_t workspace_bytes, void* stream);

/* ---- narrow MLP chains: a whole nn.Sequential per kernel ---------------------
 * The edge encoders 4->16->32->64 / 4->8->16->32 (clr_att_gnn.py:35-41, pose_gnn.py:29-35) and the
 * edge classifiers 64->32->16->8->1 [+Sigmoid] / 32->16->8->4->1 (clr_att_gnn.py:49-57,
 * pose_gnn.py:45-53): Linear/ReLU/.../Linear[/Sigmoid] with every width <= 64. One thread per row,
 * parameters in shared memory, hidden activations in registers: HBM sees X and Y only. fp32
 * arithmetic; X / Y / dY / dX may be stored as B3D_F32 or B3D_BF16.
 * dims: host int32[nl+1] = in, hidden..., out (nl = 2, 3 or 4 Linear layers). W / b / dW / db: HOST
 * arrays of nl DEVICE pointers to contiguous nn.Linear parameters ([out,in] row-major; b[l], dW[l],
 * db[l] may be null). final_act: activation after the LAST layer — B3D_ACT_NONE, B3D_ACT_SIGMOID (chains of
 * >= 3 layers) or B3D_ACT_RELU (2-layer chains) — optionally OR-ed with B3D_NARROW_INPUT_RELU (the chain
 * starts with a ReLU on its input rows; dX is masked accordingly). The last two serve the bf16 mode, where the
 * widest layer of a chain runs as a tensor-core layer in front of (classifier: 64->32 | ReLU,32->16->8->1) or
 * behind (edge encoder: 4->16->32,ReLU | 32->64) the narrow rest.
 * Backward recomputes the hidden activations, so only X is needed; dX may be null. Weight gradients
 * are reduced per CTA in registers and summed over CTAs in fixed order (deterministic). */
int b3d_narrow_mlp_supported(int32_t nl, const int32_t* dims /*host*/);
int b3d_narrow_mlp_fwd(const void* X, int32_t x_dtype, int32_t ldx, int64_t M, int32_t nl,
                       const int32_t* dims /*host*/, const float* const* W /*host*/,
                       const float* const* b /*host*/, int32_t final_act, void* Y, int32_t y_dtype,
                       int32_t ldy, void* stream);
size_t b3d_narrow_mlp_bwd_workspace_bytes(int64_t M, int32_t nl, const int32_t* dims /*host*/);
int b3d_narrow_mlp_bwd(const void* X, int32_t x_dtype, int32_t ldx, int64_t M, int32_t nl,
                       const int32_t* dims /*host*/, const float* const* W /*host*/,
                       const float* const* b /*host*/, int32_t final_act, const void* dY,
                       int32_t dy_dtype, int32_t lddy, void* dX, int32_t dx_dtype, int32_t lddx,
                       float* const* dW /*host*/, float* const* db /*host*/, int32_t flags,
                       void* workspace, size_t workspace_bytes, void* stream);

/* ---- frame-wise k-NN + attention-weighted convolution ----------------------
 * Brute-force frame-local k-NN with warp-level top-k selection. Replaces
 * torch_geometric.nn.knn_graph(x_t, k=20, loop=False) per timestamp
 * (pose_gnn.py:76-78, clr_att_gnn.py:180-182). Spec: squared L2
 * accumulated in fp32 in ascending feature order with separate multiply and
 * add, ordered by (distance, neighbour id), self excluded by id,
 * k_eff = min(k, n_frame-1). Nodes must be grouped by frame; frame_ptr int32
 * [F+1]. idx_out: int64 [N,k] global neighbour ids, -1 padded. k <= 128 (upstream knn_graph allows 100).
 * scratch: int32 [F+1]. */
int b3d_knn_frames(const float* x, int32_t ldx, int32_t D, const int32_t* frame_ptr, int32_t F,
                   int64_t N, int32_t k, int64_t* idx_out, int32_t* scratch, void* stream);

/* GATConv(D,D,heads=1,add_self_loops=False) aggregation over a padded
 * neighbour table (pose_gnn.py:55,79; clr_att_gnn.py:93,183):
 * given h = x W^T [N,D] (b3d_linear), a_s = h.att_src, a_d = h.att_dst,
 * z = leaky_relu(a_s[nbr] + a_d[t], slope), alpha = softmax_t(z) with PyG's +1e-16,
 * out[t] = sum alpha h[nbr] + bias. alpha_out (optional [N,k]) saves alpha for backward.
 * scratch: float [2N] (a_s, a_d). */
int b3d_gat_aggregate(const float* h, int32_t ldh, int32_t D, const float* att_src,
                      const float* att_dst, const float* bias, const int64_t* nbr, int32_t k,
                      int64_t N, float slope, float* out, int32_t ldo, float* alpha_out,
                      float* scratch, void* stream);

/* Backward of b3d_gat_aggregate (the intended k-NN update trains through GATConv; the reference's own
 * call discards the result, pose_gnn.py:80). Inputs: dout [N,D]; h, att_src, att_dst, nbr, slope as in the
 * forward; alpha [N,k] and a_s_a_d [2N] (the forward's alpha_out and scratch); rowptr_src [N+2] / perm_src
 * [N*k]: CSR of the flattened neighbour table grouped by SOURCE node over N+1 nodes, padding (-1) mapped
 * to the dummy node N (b3d_csr_build on edge_index = [nbr or N ; position / k]).
 * Outputs: dh [N,D] = d loss / d h including the att_src / att_dst score paths; ds [N,k] scratch;
 * das_dad [2N]: d loss / d a_s, d a_d, from which d att_src = das^T h and d att_dst = dad^T h
 * (b3d_wgrad). d bias = column sum of dout. Deterministic: per-source sums run in ascending position. */
int b3d_gat_bwd(const float* dout, int32_t ldg, const float* h, int32_t ldh, int32_t D,
                const float* att_src, const float* att_dst, const int64_t* nbr, int32_t k, int64_t N,
                float slope, const float* alpha, const float* a_s_a_d, const int32_t* rowptr_src,
                const int32_t* perm_src, float* dh, int32_t lddh, float* ds, float* das_dad, void* stream);

/* ---- multimodal front end ---------------------------------------------------
 * mask[n] = (sum(feats[n, 0:row_len]) != 0): modality-present predicate
 * (clr_att_gnn.py:107-121, 2N host syncs in the reference). */
int b3d_row_nonzero(const float* feats, int64_t row_len, int64_t N, uint8_t* mask, void* stream);

/* ---- loss and optimiser ------------------------------------------------------
 * Weighted BCE (train.py:111,136-141): loss = scale * mean_e w_e * bce(p_e, y_e).
 * from_logits=1 for PoseGNN (returns logits, pose_gnn.py:86), 0 for GNN (Sigmoid,
 * clr_att_gnn.py:57). y int64 [E]; w optional. loss_out: float[1]; grad_out [E] =
 * d loss / d input. partials: float scratch [b3d_bce_partials(E)]. */
int64_t b3d_bce_partials(int64_t E);
int b3d_bce_fwd_bwd(const float* input, const int64_t* y, const float* w, int64_t E, float scale,
                    int32_t from_logits, float* loss_out, float* grad_out, float* partials,
                    void* stream);

/* Focal edge loss named by the multimodal training config ("focal/BCE edge loss"); the reference
 * itself only has BCELoss (train.py:111), so this follows the published definition (Lin et al.):
 * loss = scale * mean_e w_e * a_t (1 - p_t)^gamma * (-log p_t), p_t = p_e or 1 - p_e, a_t = alpha or
 * 1 - alpha. Same buffers and two-stage deterministic reduction as b3d_bce_fwd_bwd. */
int b3d_focal_fwd_bwd(const float* input, const int64_t* y, const float* w, int64_t E, float scale,
                      int32_t from_logits, float alpha, float gamma, float* loss_out, float* grad_out,
                      float* partials, void* stream);

/* torch.optim.Adam step on flat buffers (train.py:106-109,160): L2 weight decay,
 * bias correction by `step` (1-based). grad_scale multiplies g first (1/world for DP). */
int b3d_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                  float beta2, float eps, float weight_decay, int32_t step, float grad_scale,
                  void* stream);
/* The same step with the step number in device memory (*step_counter = number of steps taken so far; read for
 * the bias corrections, then incremented): lets a captured CUDA graph of the whole training step be replayed. */
int b3d_adam_step_dev(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                      float beta2, float eps, float weight_decay, int32_t* step_counter, float grad_scale,
                      void* stream);

/* ---- fused MLP chains (tcgen05, hidden activations stay in shared memory) ------------------------------------
 * One launch runs a whole nn.Sequential(Linear, ReLU, ..., Linear) over dense bf16 edge rows — edge_update
 * (clr_att_gnn.py:196-201 applied at :314-317), the first layers of create_future_msgs / create_past_msgs
 * (:203-213, :319-327; two branches reading the same e'), att_edge_encoder (:82-91, :164) — or the chain of
 * input-gradient GEMMs of their backward pass (act = B3D_ACT_MASKBITS). Layers form a small DAG in program
 * order: layer l reads the chain input (src = -1) or the output of an earlier layer (src < l).
 * Per layer: y = act(x W^T + bias + sum_t add_t[sel_t(row)])  with bf16 row-gathered addends (the node-side
 * first-layer blocks applied per node: add_idx 0 -> idx0[row], 1 -> idx1[row], -1 -> row). `out` (bf16 [M,N],
 * optional) receives the layer output, `bits_out` (optional) its sign bits in the B3D_BITS layout; outputs read
 * by a later layer never leave the SM. Constraints: 1-2 dense bf16 input segments (leading width % 64, total
 * K % 16), every N % 64 == 0 and <= 512, at most B3D_CHAIN_MAX_LAYERS layers, and the activation working set of a
 * 128-row tile must fit beside a >= 2-stage weight ring in 227 KB (b3d_chain_supported tells).
 * Weights are packed once per weight version by b3d_chain_pack_weights (layer by layer, fp32 [N,K] row-major
 * with leading dimension ldw, or its transpose) into a pre-swizzled chunk stream of b3d_chain_packed_bytes. */
#define B3D_CHAIN_MAX_LAYERS 6
typedef struct {
  int32_t K, N;
  int32_t src;                 /* -1: chain input; else index of the producing layer */
  int32_t act;                 /* B3D_ACT_NONE / B3D_ACT_RELU / B3D_ACT_MASKBITS */
  int32_t nadd;
  int32_t add_idx[2];
  int32_t add_ld[2];
  const void* add_ptr[2];      /* bf16 [*, N], rows 32-byte aligned (ld % 16 == 0) */
  const float* bias;           /* [N] or NULL */
  void* out;                   /* bf16 [M, N] (rows 32-byte aligned) or NULL */
  int32_t ldo;
  void* bits_out;              /* uint32 [N/32][M] or NULL */
  const void* bits_in;         /* B3D_ACT_MASKBITS: uint32 [N/32][M] */
} b3d_chain_layer_t;
int b3d_chain_supported(const b3d_chain_layer_t* layers /*host*/, int32_t nl, int32_t k_in);
size_t b3d_chain_packed_bytes(const b3d_chain_layer_t* layers /*host*/, int32_t nl, int32_t k_in);
int b3d_chain_pack_weights(const b3d_chain_layer_t* layers /*host*/, int32_t nl, int32_t k_in, int32_t layer,
                           const float* W, int32_t ldw, int32_t transpose, void* packed, void* stream);
int b3d_chain_run(const b3d_seg_t* in_segs /*host*/, int32_t nseg, const b3d_chain_layer_t* layers /*host*/,
                  int32_t nl, const void* packed, const int32_t* idx0, const int32_t* idx1, int64_t M, void* stream);

/* ---- graph construction: same-category k-NN under the normalised motion metric ---------------------------------
 * get_knn_nodes_in_graph (batch_3dmot/utils/graph_utils.py:33-88) for all current nodes of a window at once
 * (construct_detection_graph_disjoint_parallel_only_poses.py:204-224): row r is current node cur[r]; its
 * candidates are order[first[r] .. first[r] + cnt[r]) (same-category nodes of earlier frames, ascending id).
 * center / velocity: float64 [N,3], yaw: float64 [N]. ex: int64 [R, kmax] receives the min(top_knn, cnt[r])
 * selected node ids in ascending (metric, candidate position) order, -1 padded. flags[r]: bit 0 = the metric holds
 * a NaN (0/0 when a maximum is 0; the caller resolves such rows with the reference's 1-D sequence), bit 1 = an
 * exact tie among the first k+1 values (order = candidate position; unspecified upstream). */
int b3d_window_knn(const double* center, const double* velocity, const double* yaw, const int64_t* order,
                   const int64_t* cur, const int64_t* first, const int64_t* cnt, int64_t R, int32_t top_knn,
                   int32_t kmax, int64_t* ex, int32_t* flags, void* stream);

/* ---- track assembly, host side (the only entry point that takes HOST pointers and runs on the CPU) -------------
 * create_trajectories(mode='hier') + track-id numbering (predict.py:308-373, :437-446) over the surviving edges
 * of one or many scenes: edges (e_out -> e_in, float64 score) in the reference's greedy_edges insertion order,
 * clustered in stable descending-score order (new / prepend / append / join, the join gated by the per-class
 * threshold of the IN node). track_id[n] = position of the node's track among its scene's tracks in insertion
 * order (-1: none), track_pos[n] = position of the node inside its track, tracks_per_scene[s] = track count.
 * Returns -1 if an edge would close a cycle inside one track (the reference corrupts its state there). */
int b3d_hier_tracks_host(const int64_t* e_out, const int64_t* e_in, const double* score, int64_t m,
                         const int64_t* node_class, const int32_t* scene_of_node /*nullable*/, int64_t n,
                         int32_t n_scenes, const double* thresholds, int32_t n_classes,
                         int64_t* track_id, int64_t* track_pos, int64_t* tracks_per_scene);

#ifdef __cplusplus
}
#endif
#endif /* B3D_H_ */
