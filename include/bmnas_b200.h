/*
 * bmnas_b200.h -- C ABI of the B200-native BM-NAS search-step kernels.
 *
 * The reference (Somedaywilldo/BM-NAS) has no FFI layer: its "operator API" is the
 * Python class surface of models/search/darts/.  Each entry point below replaces
 * the chain of eager ATen calls behind one of those classes; the reference
 * file:line each one stands in for is cited on the declaration.  INTEGRATION.md
 * shows the ctypes binding a maintainer adds on the reference side.
 *
 * Rules of the ABI
 *   - plain C: POD parameter blocks, raw DEVICE pointers, explicit sizes; no torch types;
 *   - all tensors are fp32, contiguous, laid out (B, C, L) with L fastest
 *     (exactly the reference's activation layout);
 *   - no allocation and no host synchronisation inside; the caller passes outputs
 *     and workspaces; every call is stream-ordered on `stream` (a cudaStream_t) and
 *     legal under CUDA-graph capture;
 *   - returns 0 on success, a negative BMNAS_E* code otherwise (bmnas_strerror()).
 *   - "counter" workspaces are unsigned ints that must be zero before the first
 *     use; every kernel leaves them zero again.
 *
 * NOTE: bm-nas_b200/bmnas/native.py parses this file to build its ctypes
 * structures.  Keep declarations in the simple `type name;` / `type name[MACRO];`
 * form, one per line.
 */
#ifndef BMNAS_B200_H
#define BMNAS_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define BMNAS_MAX_MIX 16 /* candidate inputs of one edge mix                  */
#define BMNAS_MAX_SRC 4  /* tensors in a virtual channel concat               */
#define BMNAS_MAX_SEG 4  /* stacked output-channel segments of one conv GEMM  */
#define BMNAS_MAX_OPS 8  /* candidate primitives of one NodeMixedOp           */
#define BMNAS_MAX_PREP 16     /* convs per bmnas_wprep launch                  */
#define BMNAS_MAX_PREP_SEG 64 /* BMNAS_MAX_PREP * BMNAS_MAX_SEG                */
#define BMNAS_MAX_PREP_P1 17  /* BMNAS_MAX_PREP + 1                            */

#define BMNAS_OK 0
#define BMNAS_EINVAL -1   /* bad shape / unsupported size / null pointer */
#define BMNAS_ELAUNCH -2  /* CUDA launch failure                         */

/* step-node primitive ids (models/search/darts/node_operations.py:9-14, 66-82) */
#define BMNAS_OP_SUM 0
#define BMNAS_OP_ATTN 1
#define BMNAS_OP_GLU 2
#define BMNAS_OP_FC_RELU 3
#define BMNAS_OP_FC_MISH 4

const char* bmnas_strerror(int code);
int bmnas_abi_version(void);

/* ------------------------------------------------------------------------
 * Edge mix: out = sum_j w_j[skip] * x_j   (PRIMITIVES = [none, skip])
 * replaces FusionMixedOp.forward + the python sum over edges:
 *   models/search/darts/operations.py:95-106 (Zero :14-20, Identity :88-93),
 *   model_search.py:58, node_search.py:54.
 * w is (n,2).  w_is_logits=1: w holds raw alpha/beta rows and the per-edge
 * 2-way softmax (model_search.py:95, node_search.py:102) is taken in-kernel;
 * the backward then returns d/d(logits).
 * Chained mix (optional, out2 / gout2 / w2 / n2): the first inner edge mix of a searchable NodeCell reads the
 * cell-level mix twice (states = [x, y] with x is y, node_search.py:49-54), i.e. it is s2 * out with
 * s2 = sum_{j<n2} w2_j[skip].  The forward writes out2 = s2 * out from the same registers; the backward takes
 * the upstream gradient as gout + s2 * gout2 (either may be NULL).  d/d(w2) is a separate dot-only call.
 * ---------------------------------------------------------------------- */
typedef struct bmnas_mix_params {
    int n;
    int w_is_logits;
    long long numel;
    const float* x[BMNAS_MAX_MIX];
    const float* w;
    float* out;
    const float* gout;
    float* gx[BMNAS_MAX_MIX];
    int gx_accum[BMNAS_MAX_MIX];
    float* gw;
    float* partials;
    unsigned int* counter;
    int n2;
    const float* w2;
    float* out2;
    const float* gout2;
} bmnas_mix_params;
int bmnas_mix_fwd(const bmnas_mix_params* p, void* stream);
int bmnas_mix_bwd(const bmnas_mix_params* p, void* stream);
long long bmnas_mix_partials_size(const bmnas_mix_params* p); /* floats */

/* ------------------------------------------------------------------------
 * 1x1 convolution over a virtual channel concat, with fused bias and
 * train-mode BatchNorm statistics:
 *     Z[b, m, l] = sum_k Weff[m, k] * U[b, k, l] + bias[m]
 * U = channel-concat of src[0..n_src) (never materialised); the output rows are
 * the stacked segments W[0..n_seg) (LinearGLU's 2C rows and ConcatFC's C rows
 * share one GEMM).  Each W[i] is (seg_M[i], w_fold*K) row major and
 * Weff[m,k] = sum_f W[m, f*K + k]: w_fold=2 folds cat([t, t]) (search mode calls
 * every step node with x is y: model_search.py:59, node_search.py:55).
 * replaces torch.cat + nn.Conv1d(k=1) + the statistics half of nn.BatchNorm1d:
 *   node_operations.py:30-34 (LinearGLU), :49-53 (ConcatFC), :75-79 (CatConvMish),
 *   node_search.py:59-62 (out_conv + bn).
 * bn_mode 1: batch mean / rstd over (B, L) -> mean, rstd; running statistics are
 * updated (momentum, unbiased variance) and num_batches_tracked incremented.
 * bn_mode 2: mean/rstd are filled from the running statistics (eval).
 * dgrad / wgrad are the two backward GEMMs.  The upstream gradient operand is
 *     dz[b,m,l] = coef_a[m]*GV[b,m,l] + coef_b[m]*Z[b,m,l] + coef_c[m]
 * (BatchNorm backward folded into the operand load; coef_* NULL => dz = GV).
 * gW / gbias are accumulated atomically (red.add) onto whatever the caller put
 * there: zero for a gradient, the broadcast bias when conv_wgrad is used as the
 * NT GEMM of the classifier forward.
 * ---------------------------------------------------------------------- */
typedef struct bmnas_conv_params {
    int B;
    int L;
    int K;
    int w_fold;
    int n_src;
    int n_seg;
    int M;
    int bn_mode;
    int splits;
    float momentum;
    float eps;
    int src_C[BMNAS_MAX_SRC];
    int seg_M[BMNAS_MAX_SEG];
    const float* src[BMNAS_MAX_SRC];
    const float* W[BMNAS_MAX_SEG];
    const float* bias[BMNAS_MAX_SEG];
    float* Z;
    float* stat_part;
    unsigned int* counter;
    float* mean;
    float* rstd;
    float* running_mean[BMNAS_MAX_SEG];
    float* running_var[BMNAS_MAX_SEG];
    long long* num_batches_tracked[BMNAS_MAX_SEG];
    const float* GV;
    const float* coef_a;
    const float* coef_b;
    const float* coef_c;
    float* gsrc[BMNAS_MAX_SRC];
    int gsrc_accum[BMNAS_MAX_SRC];
    float* gW[BMNAS_MAX_SEG];
    float* gbias[BMNAS_MAX_SEG];
    const float* wimg_fwd;
    const float* wimg_dgrad;
    int wimg_fmt;
    int early_ok;
} bmnas_conv_params;
/* wimg_fwd / wimg_dgrad (optional): images of the stacked, folded weight produced by bmnas_wprep (below), in
 * format wimg_fmt: 0 = tcgen05 tf32 hi/lo slabs (the tensor-core GEMMs fetch them with TMA bulk copies), 2 = bf16
 * slabs (128 rows x 64 k, forward image only; bmnas_mixed_fwd in gemm mode 3), 1 = plain fp32
 * (tile major [row tile of 32][reduction][32], fold applied; the small-N cp.async FFMA GEMMs of gemm_sg.cu).  They must be
 * refreshed (bmnas_wprep) whenever W changes.  bmnas_conv_image_fmt() picks the format for a problem size. */
int bmnas_conv_fwd(const bmnas_conv_params* p, void* stream);
int bmnas_conv_dgrad(const bmnas_conv_params* p, void* stream);
int bmnas_conv_wgrad(const bmnas_conv_params* p, void* stream);
long long bmnas_conv_stat_part_size(const bmnas_conv_params* p); /* floats  */
int bmnas_conv_num_counters(const bmnas_conv_params* p);         /* uints   */

/* ------------------------------------------------------------------------
 * Weight images for the tcgen05 GEMMs: for each of n convs (conv i: stacked segments
 * W[i*BMNAS_MAX_SEG + s] of seg_M[...] rows, w_fold*K columns) write
 *   img_fwd[i]   rows m, reduction k   (bmnas_wimg_floats(M, K, 0) floats)
 *   img_dgrad[i] rows k, reduction m   (bmnas_wimg_floats(M, K, 1) floats)
 * of Weff[m,k] = sum_f W[m, f*K + k], split into tf32 hi / lo parts, zero padded to 128-row x 32-element
 * slabs and stored slab by slab as the K-major core-matrix shared-memory picture the MMA descriptors read.
 * q_start is the exclusive prefix sum of bmnas_wprep_items(M, K, fmt) over the convs (n + 1 entries);
 * fmt[i] selects the image format of conv i (see bmnas_conv_params.wimg_fmt).
 * Either image pointer of a conv may be NULL (that image is then not written): a conv can so be listed twice with two
 * formats, e.g. tcgen05 slabs for the fused forward kernel and the plain fp32 image for the small-N dgrad GEMM.
 * One launch per forward replaces the per-CTA fold / transpose / split of the weights
 * (torch.cat([x, x]) + nn.Conv1d weight use in node_operations.py:30-34, 49-53, node_search.py:59-62).
 * ---------------------------------------------------------------------- */
typedef struct bmnas_wprep_params {
    int n;
    int M[BMNAS_MAX_PREP];
    int K[BMNAS_MAX_PREP];
    int w_fold[BMNAS_MAX_PREP];
    int n_seg[BMNAS_MAX_PREP];
    int seg_M[BMNAS_MAX_PREP_SEG];
    const float* W[BMNAS_MAX_PREP_SEG];
    float* img_fwd[BMNAS_MAX_PREP];
    float* img_dgrad[BMNAS_MAX_PREP];
    long long q_start[BMNAS_MAX_PREP_P1];
    int fmt[BMNAS_MAX_PREP];
    unsigned long long* rng_state; /* optional: also advance the dropout step counter (saves the bmnas_rng_advance launch) */
} bmnas_wprep_params;
int bmnas_wprep(const bmnas_wprep_params* p, void* stream);
long long bmnas_wimg_floats(int M, int K, int which);               /* fmt 0 */
long long bmnas_wimg_floats_fmt(int M, int K, int which, int fmt);  /* floats of one image in format fmt */
long long bmnas_wprep_items(int M, int K, int fmt);                 /* work items of one conv (q_start increments) */
/* image format for a conv over B*L columns: -1 = none (shape not eligible), 0 = tcgen05 slabs, 1 = plain fp32
 * (small column counts: the fp32 cp.async GEMMs beat 3xTF32 UMMA there; see gemm_sg.cu) */
int bmnas_conv_image_fmt(int B, int L, int K, int M);
/* the same choice for the dgrad GEMM of that conv (its weight image may use another format than the forward one:
 * bmnas_wprep writes either image in either format) */
int bmnas_conv_image_fmt_dgrad(int B, int L, int K, int M);

/* ------------------------------------------------------------------------
 * Adaptive max pooling of a raw backbone feature map onto the (C_in, L) grid: the first stage of the reshape
 * layers right upstream of the fusion cells.
 * replaces nn.AdaptiveMaxPool2d((L, 1)) in ReshapeInputLayer.forward (models/auxiliary/aux_models.py:61-69;
 * the F.interpolate(size=L) after it is the identity) and nn.AdaptiveMaxPool2d((sqrt L, sqrt L)) in
 * ReshapeInputLayer_MMIMDB.forward (:102-110).  The Conv1d -> BatchNorm1d -> ReLU -> Dropout block that follows
 * runs on bmnas_conv_* + bmnas_node_* (it is ConcatFC over a single source).
 * x (B, C, H, W) contiguous -> out (B, C, OH*OW); bin (i, j) covers rows [floor(i*H/OH), ceil((i+1)*H/OH)) and
 * columns likewise (ATen rule).  argmax (int32, flat h*W+w per bin) is optional in the forward and required by
 * the backward, which gathers: gx[e] = sum of gout over the bins that elected e (+ gx if gx_accum).
 * ---------------------------------------------------------------------- */
typedef struct bmnas_pool_params {
    int B;
    int C;
    int H;
    int W;
    int OH;
    int OW;
    int gx_accum;
    const float* x;
    float* out;
    int* argmax;
    const float* gout;
    float* gx;
} bmnas_pool_params;
int bmnas_pool_fwd(const bmnas_pool_params* p, void* stream);
int bmnas_pool_bwd(const bmnas_pool_params* p, void* stream);

/* ------------------------------------------------------------------------
 * Step-node mixed op: out = sum_k gamma~_k * op_k(x, y), evaluated per sample
 * from the x / y tiles staged once in shared memory; per-op outputs are never
 * written to HBM.
 * replaces NodeMixedOp.forward and every primitive's forward:
 *   node_operations.py:110-120 (mixed), :16-20 Sum, :84-108 ScaledDotAttn,
 *   :22-39 LinearGLU (BN apply + GLU + dropout), :41-56 ConcatFC, :66-82 CatConvMish.
 * gamma NULL => every listed op has weight 1 (found network / stand-alone op,
 * node.py:45-62).  gamma_is_logits=1 => softmax over n_ops taken in-kernel
 * (node_search.py:103) and g_gamma is d/d(logits).
 * Dropout: mask[k] (uint8 keep mask, (B,C,L)) if given, else Philox keyed by
 * (rng_state[0], rng_state[1], op_uid[k], global element index) when
 * training && p_drop[k] > 0.
 * Backward recomputes the primitives from (x, y, Z), writes GV = dL/d(BN output)
 * for the conv-backed ops, reduces dL/dgamma, BN affine grads and the
 * coefficients for the conv backward operand (see bmnas_conv_params).
 * ---------------------------------------------------------------------- */
typedef struct bmnas_node_params {
    int B;
    int C;
    int L;
    int n_ops;
    int M;
    int training;
    int gamma_is_logits;
    int alias_xy;
    int gx_accum;
    int gy_accum;
    long long sample_offset;
    int op_type[BMNAS_MAX_OPS];
    int z_off[BMNAS_MAX_OPS];
    float p_drop[BMNAS_MAX_OPS];
    unsigned int op_uid[BMNAS_MAX_OPS];
    const float* x;
    const float* y;
    const float* Z;
    const float* mean;
    const float* rstd;
    const float* bn_w[BMNAS_MAX_OPS];
    const float* bn_b[BMNAS_MAX_OPS];
    const float* ln_w[BMNAS_MAX_OPS];
    const float* ln_b[BMNAS_MAX_OPS];
    const unsigned char* mask[BMNAS_MAX_OPS];
    const unsigned long long* rng_state;
    const float* gamma;
    float* out;
    const float* gout;
    float* gx;
    float* gy;
    float* GV;
    float* g_gamma;
    float* g_bn_w[BMNAS_MAX_OPS];
    float* g_bn_b[BMNAS_MAX_OPS];
    float* g_ln_w[BMNAS_MAX_OPS];
    float* g_ln_b[BMNAS_MAX_OPS];
    float* coef_a;
    float* coef_b;
    float* coef_c;
    float* partials;
    unsigned int* counter;
    int early_ok;
    /* Chained edge mix (optional; one-CTA-per-sample kernels only): the NEXT inner edge mix of a searchable NodeCell
     * (node_search.py:52-55) is sum_j w_j * states[j] over the earlier states and this op's own output, so the
     * forward also writes   out2 = sum_{j<n_chain} cw_j * chain_x[j] + cw_{n_chain} * out
     * (cw = skip weights of chain_w rows, 2-way softmax in-kernel when chain_is_logits), and the backward takes
     * gout + cw_{n_chain} * gout2 as its upstream gradient (gout may then be NULL) and adds cw_j * gout2 into
     * chain_gx[j] (overwrite when chain_gx_accum[j] == 0; NULL = not wanted).  d/d(chain_w) is a separate
     * dot-only bmnas_mix_bwd call. */
    int n_chain;
    int chain_is_logits;
    const float* chain_w;
    const float* chain_x[BMNAS_MAX_SRC];
    float* out2;
    const float* gout2;
    float* chain_gx[BMNAS_MAX_SRC];
    int chain_gx_accum[BMNAS_MAX_SRC];
} bmnas_node_params;
int bmnas_node_fwd(const bmnas_node_params* p, void* stream);
int bmnas_node_bwd(const bmnas_node_params* p, void* stream);
long long bmnas_node_partials_size(const bmnas_node_params* p); /* floats */

/* ------------------------------------------------------------------------
 * Fused step-node mixed op (the north-star kernel): bmnas_conv_fwd + bmnas_node_fwd of a searchable NodeMixedOp in
 * ONE persistent cooperative launch on the tcgen05 tensor cores, without the pre-BatchNorm activations Z ever
 * leaving the SM unless the caller asks for them:
 *   out = sum_k softmax(gamma)_k * op_k(t, t),   ops a canonical-order subset of
 *         {Sum, ScaleDotAttn, LinearGLU, ConcatFC | CatConvMish} containing LinearGLU and one FC-type op
 * replaces NodeMixedOp.forward with everything below it: node_operations.py:118-120 (weighted sum), :19-20 (Sum),
 * :92-108 (ScaledDotAttn: QK^T, softmax, PV, dropout, LayerNorm), :30-39 (LinearGLU: cat, Conv1d, BatchNorm1d with
 * batch statistics, GLU, dropout), :49-56 (ConcatFC), :75-82 (CatConvMish), called with x is y (node_search.py:55).
 * Per 64-column tile (64/L samples): producer warps stage the activation tile once (fp32 -> hi/lo tf32 or bf16,
 * K-major SWIZZLE_128B), a TMA warp streams the prepared weight slabs (bmnas_wprep image), one thread issues the
 * UMMAs for the three 128-row output tiles (GLU value | GLU gate | FC) AND the Gram matrix t^T t whose diagonal
 * L x L blocks are the attention scores; accumulators are double buffered in tensor memory; eight epilogue warps
 * reduce the BatchNorm batch statistics per output row (Welford per CTA -> fp64 atomics -> ONE grid barrier),
 * then apply BN / GLU / ReLU / Mish / softmax / PV / dropout / LayerNorm / the gamma-weighted sum straight from
 * tensor memory.  Batches larger than one resident wave (2 tiles per SM) recompute the GEMM in a second pass.
 * cv: the conv block (B, L, K = C = 128, w_fold, W / bias / running statistics, mean / rstd outputs, bn_mode,
 *     wimg_fwd in format 0 (3xTF32) or 2 (bf16); cv->Z NULL = do not write Z (no-grad forward), else Z (B, 3C, L) is
 *     written for the backward kernels).   nd: the node block (ops, gamma, dropout sites, LayerNorm affine, out).
 * workspace: bmnas_mixed_workspace_bytes() bytes, zeroed once by the caller, owned by this (cv, nd) pair.
 * Returns BMNAS_EINVAL for shapes it does not take (bmnas_mixed_supported() == 0): the caller then uses
 * bmnas_conv_fwd + bmnas_node_fwd.
 * ---------------------------------------------------------------------- */
int bmnas_mixed_fwd(const bmnas_conv_params* cv, const bmnas_node_params* nd, void* workspace, void* stream);
int bmnas_mixed_supported(const bmnas_conv_params* cv, const bmnas_node_params* nd);
long long bmnas_mixed_workspace_bytes(void);

/* ------------------------------------------------------------------------
 * The same fused mixed op for batches SMALLER than the machine (the reference batch: NTU 96 x 8 = 768 columns), where the
 * problem is latency, not throughput: fp32 FFMA from a TMA-staged 96-row x 32-column tile per CTA (32 channels: their
 * GLU value / GLU gate / FC rows; 32 columns = 32/L whole samples), grid = (B*L/32) x (C/32) co-resident CTAs, the
 * attention primitive computed from the same activation tile while the BatchNorm partials travel, ONE grid barrier,
 * epilogue from registers.  Also writes the NEXT inner edge mix (nd->out2, n_chain, chain_x: node_search.py:52-55).
 * Needs cv->wimg_fwd in format 1 (tile-major fp32), C % 32 == 0, C <= 256, L in {4, 8, 16} and a grid that fits the
 * machine (bmnas_mixed_small_supported); same parameter blocks and call sites as bmnas_mixed_fwd; the workspace is
 * bmnas_mixed_small_workspace_bytes(cv, nd) bytes, zeroed once.
 * ---------------------------------------------------------------------- */
int bmnas_mixed_small_fwd(const bmnas_conv_params* cv, const bmnas_node_params* nd, void* workspace, void* stream);
int bmnas_mixed_small_supported(const bmnas_conv_params* cv, const bmnas_node_params* nd);
long long bmnas_mixed_small_workspace_bytes(const bmnas_conv_params* cv, const bmnas_node_params* nd);

/* ------------------------------------------------------------------------
 * LayerNorm block over a virtual channel concat.
 * mode 0 (CAT):  v = cat(src[0..n_src)) [+ residual];  out = LN_{[Ctot,L]}(v) [ReLU]
 *   replaces FusionCell tail  model_search.py:63-67  (cat -> LayerNorm -> ReLU -> view)
 *   and NodeCell tail with node_multiplier == 1  node_search.py:59,67-68.
 * mode 1 (TAIL): v = dropout(ReLU(BN(src[0]))) + residual;  out = LN_{[C,L]}(v)
 *   src[0] is the out_conv GEMM output; replaces node_search.py:62-68 / node.py:69-74.
 * Backward writes the source grads (CAT) or GV = dL/d(BN output) (TAIL, gsrc[0]),
 * the residual grad, accumulates LN affine grads atomically and finalises BN
 * affine grads + conv-backward coefficients.
 * ---------------------------------------------------------------------- */
typedef struct bmnas_ln_params {
    int B;
    int L;
    int Ctot;
    int n_src;
    int mode;
    int relu_out;
    int training;
    int gres_accum;
    float p_drop;
    unsigned int op_uid;
    long long sample_offset;
    int src_C[BMNAS_MAX_SRC];
    int gsrc_accum[BMNAS_MAX_SRC];
    const float* src[BMNAS_MAX_SRC];
    const float* residual;
    const float* mean;
    const float* rstd;
    const float* bn_w;
    const float* bn_b;
    const unsigned char* mask;
    const unsigned long long* rng_state;
    const float* ln_w;
    const float* ln_b;
    float* out;
    const float* gout;
    float* gsrc[BMNAS_MAX_SRC];
    float* gresidual;
    float* g_ln_w;
    float* g_ln_b;
    float* g_bn_w;
    float* g_bn_b;
    float* coef_a;
    float* coef_b;
    float* coef_c;
    float* partials;
    unsigned int* counter;
} bmnas_ln_params;
int bmnas_ln_fwd(const bmnas_ln_params* p, void* stream);
int bmnas_ln_bwd(const bmnas_ln_params* p, void* stream);
long long bmnas_ln_partials_size(const bmnas_ln_params* p); /* floats */

/* ------------------------------------------------------------------------
 * Loss head: mean cross-entropy (kind 0, int64 labels) or mean
 * BCE-with-logits (kind 1, float targets), criterion of
 * ntu_darts_searchable.py:25 / mmimdb_darts_searchable.py:22.
 * fwd writes the scalar loss and glogits = d loss / d logits; bwd scales
 * glogits by the upstream scalar *gscale into gout_logits.
 * bias helpers for the classifier (nn.Linear, ntu_darts_searchable.py:100-101):
 * bmnas_bias_rows broadcasts bias into (rows, n); bmnas_colsum reduces (rows,n)->(n).
 * ---------------------------------------------------------------------- */
typedef struct bmnas_loss_params {
    int B;
    int n_classes;
    int kind;
    const float* logits;
    const long long* labels;
    const float* targets;
    float* loss;
    float* glogits;
    const float* gscale;
    float* gout_logits;
    float* partials;
    unsigned int* counter;
} bmnas_loss_params;
int bmnas_loss_fwd(const bmnas_loss_params* p, void* stream);
int bmnas_loss_bwd(const bmnas_loss_params* p, void* stream);
long long bmnas_loss_partials_size(const bmnas_loss_params* p); /* floats */
int bmnas_bias_rows(float* out, const float* bias, int rows, int n, void* stream);
int bmnas_colsum(float* out, const float* in, int rows, int n, void* stream);

/* ------------------------------------------------------------------------
 * Classifier head nn.Linear (central_classifier: ntu_darts_searchable.py:100-101, 176-178;
 * mmimdb_darts_searchable.py / ego_darts_searchable.py likewise): x (B,K) row major, W (N,K), out (B,N).
 *   fwd: out = x W^T + bias                      (bias may be NULL)
 *   bwd: gx = gout W (if gx), gW = gout^T x (if gW), gbias = column sums of gout (if gbias);
 *        every result OVERWRITES its destination.  K % 4 == 0, 16-byte aligned x / W / gx / gW.
 * ---------------------------------------------------------------------- */
typedef struct bmnas_linear_params {
    int B;
    int K;
    int N;
    const float* x;
    const float* W;
    const float* bias;
    float* out;
    const float* gout;
    float* gx;
    float* gW;
    float* gbias;
} bmnas_linear_params;
int bmnas_linear_fwd(const bmnas_linear_params* p, void* stream);
int bmnas_linear_bwd(const bmnas_linear_params* p, void* stream);

/* ------------------------------------------------------------------------
 * Classifier head + criterion + their backward in ONE launch (the tail of a forward and the head of a backward
 * of the search step):
 *   logits = x W^T + bias                       central_classifier (ntu_darts_searchable.py:100-101, 176-178)
 *   loss   = mean CE(logits, labels)   kind 0   nn.CrossEntropyLoss   (ntu_darts_searchable.py:25)
 *          | mean BCE-with-logits      kind 1   nn.BCEWithLogitsLoss  (mmimdb_darts_searchable.py:22)
 *   glogits = d loss / d logits                 what loss.backward() (train_searchable/ntu.py:88, architect.py:27-28)
 *   gx      = glogits W                         hands to the classifier and through it to the fusion network
 * gx NULL: forward + loss only.  glogits NULL: not stored.  gW / gbias are bmnas_linear_bwd's from glogits.
 * Labels outside [0, N) give a NaN loss and NaN gradients for that sample (as bmnas_loss_fwd).
 * partials: bmnas_head_partials_size floats of workspace; counter: one zeroed unsigned int (the kernel re-zeroes it).
 * K % 4 == 0, N <= 128, 16-byte aligned x / W / gx; bmnas_head_supported tells (0: use linear_fwd + loss_fwd + ...).
 * ---------------------------------------------------------------------- */
typedef struct bmnas_head_params {
    int B;
    int K;
    int N;
    int kind;
    const float* x;
    const float* W;
    const float* bias;
    const long long* labels;
    const float* targets;
    float* logits;
    float* loss;
    float* glogits;
    float* gx;
    float* partials;
    unsigned int* counter;
} bmnas_head_params;
int bmnas_head_supported(const bmnas_head_params* p);
long long bmnas_head_partials_size(const bmnas_head_params* p); /* floats */
int bmnas_head_fused(const bmnas_head_params* p, void* stream);

/* ------------------------------------------------------------------------
 * Multi-tensor Adam, identical arithmetic to torch.optim.Adam (coupled L2 weight
 * decay, bias correction, eps outside the sqrt):
 *   weights  ntu_darts_searchable.py:42    arch  :46-47   (architect.py:24 step)
 * tensors is a DEVICE array of n_tensors descriptors.  lr and step live in
 * device memory so a captured CUDA graph sees the per-iteration schedule value
 * (replaces scheduler.update_optimizer's state_dict round trip, scheduler.py:42-46).
 * grad_scale multiplies every gradient first (1/world_size after an all-reduce).
 * lr_ring > 0: lr points at a ring of lr_ring schedule values and the step uses lr[step % lr_ring] -- the host
 * uploads the cosine-restart schedule (scheduler.py:25-40) a few hundred steps ahead instead of writing a scalar before
 * every replay; lr_ring == 0: lr[0].
 * ---------------------------------------------------------------------- */
typedef struct bmnas_adam_tensor {
    float* p;
    const float* g;
    float* m;
    float* v;
    long long n;
    long long block_start;
} bmnas_adam_tensor;
typedef struct bmnas_adam_params {
    int n_tensors;
    int block_elems;
    long long total_blocks;
    const bmnas_adam_tensor* tensors;
    const float* lr;
    float beta1;
    float beta2;
    float eps;
    float weight_decay;
    float grad_scale;
    long long* step;
    unsigned int* counter;
    int lr_ring;
} bmnas_adam_params;
int bmnas_adam_step(const bmnas_adam_params* p, void* stream);

/* ------------------------------------------------------------------------
 * Data-parallel optimiser step fused with its collective, over NVLink peer memory (one launch per rank and bucket):
 *   reduce-scatter of the gradient bucket (P2P loads, rank-ordered sum) -> Adam on this rank's 1/world shard ->
 *   all-gather of the UPDATED parameters (P2P stores into every replica)
 * replaces nn.DataParallel's gradient reduce-add (ntu_darts_searchable.py:50-51) + torch.optim.Adam.step on every
 * replica (:42 weights, :46-47 architecture; architect.py:24) -- i.e. ncclAllReduce + bmnas_adam_step.
 * grad_ptrs / param_ptrs / signal_ptrs are DEVICE arrays of `world` pointers: rank r's gradient bucket, parameter bucket
 * (same element layout as the gradients) and signal pad, all in peer-mapped ("symmetric") memory.  The launch uses
 * 2 * world uint32 slots of every signal pad from signal_base; they must start at 0.  m / v hold this rank's shard only
 * (ceil(n / 4 / world) * 4 floats).  n % 4 == 0.  Every rank must issue the same launch; the kernels wait for one another.
 * ---------------------------------------------------------------------- */
typedef struct bmnas_dp_adam_params {
    int world;
    int rank;
    long long n;
    const float* const* grad_ptrs;
    float* const* param_ptrs;
    unsigned int* const* signal_ptrs;
    int signal_base;
    float* m;
    float* v;
    const float* lr;
    float beta1;
    float beta2;
    float eps;
    float weight_decay;
    float grad_scale;
    long long* step;
    unsigned int* epoch;
    unsigned int* done_counter;
    int lr_ring;
} bmnas_dp_adam_params;
int bmnas_dp_adam_step(const bmnas_dp_adam_params* p, void* stream);

/* stream-ordered zero fill (cudaMemsetAsync) and ABI self-description for binding tests */
int bmnas_zero(void* ptr, long long nbytes, void* stream);
/* device-to-device copy as a kernel (16-byte aligned pointers, nbytes % 16 == 0): the step's packed input batch into the
 * static buffers the captured graphs read (SearchStep.load_step) */
int bmnas_copy(void* dst, const void* src, long long nbytes, void* stream);
int bmnas_sizeof_params(int which); /* 0 mix, 1 conv, 2 node, 3 ln, 4 loss, 5 adam_tensor, 6 adam */

/* validate-only mode: every entry point checks its parameter block and returns before launching
 * (lets host-side logic and bindings be tested on a machine without a GPU). */
int bmnas_set_validate_only(int on);

/* GEMM engine behind bmnas_conv_*: 0 = fp32 FFMA tiles, 1 = tcgen05 tensor cores with 3xTF32 operand
 * splitting (fp32-class accuracy; default), 2 = tcgen05 single-pass TF32 (reduced precision), 3 = as 1 with bf16
 * operands (kind::f16, fp32 accumulation in tensor memory) in the fused mixed-op forward bmnas_mixed_fwd -- the
 * north_star's reduced-precision mode, parity gate 2e-2 (weight image format 2, forward only).  Shapes the
 * tensor-core path cannot take (L, K or a concat width not a multiple of 4, unaligned tensors) use mode 0. */
int bmnas_set_gemm_mode(int mode);
int bmnas_get_gemm_mode(void);

/* programmatic dependent launch: every kernel is launched with cudaLaunchAttributeProgrammaticStreamSerialization
 * (a programmatic graph edge under stream capture) and follows the protocol [early section] -> griddepcontrol.wait
 * -> griddepcontrol.launch_dependents -> body (csrc/common.cuh).  early_ok in bmnas_conv_params / bmnas_node_params
 * tells a kernel that its early inputs (weight images; the x / y tiles of a node op) were produced at least two
 * kernels before it in the stream, so it may read them before the wait and overlap its predecessor: the conv
 * kernels prefetch the weight tile, the node kernels run the whole attention primitive (forward) / its
 * recomputation (backward) there.  The caller sets it; 0 is always safe. */
int bmnas_set_pdl(int on);

/* Kernel variant behind bmnas_node_fwd / bmnas_node_bwd: 0 = by batch size (default: one CTA per sample while the
 * launch is latency bound -- B < 768 forward, B < 640 backward -- one warp per sample beyond, where it is
 * throughput bound), 1 = always CTA per sample, 2 = warp per sample whenever the shape is eligible (L in
 * {4,8,16}, ceil(C/32)*L <= 32, 16-byte aligned tensors, at most 3 conv-backed primitives of which at most one
 * LinearGLU; x is y -- the searchable cell -- double-buffers its tiles, x != y -- the found cell -- stages them
 * per sample).  All variants draw identical dropout masks and agree to fp32 rounding. */
int bmnas_set_node_variant(int v);

/* Engine behind the tensor-core bmnas_conv_fwd / bmnas_conv_dgrad when bmnas_wprep images exist: enable = 1
 * (default) the warp-specialised persistent kernel (csrc/gemm_ws.cu: producer warps, one MMA-issuing warp, a TMA warp
 * for the weight ring, epilogue warps over two accumulator sets in tensor memory), 0 = the phase-serial panel kernel
 * it replaced (kept for A/B measurements).  max_ctas > 0 caps the CTAs per accumulator row tile (test hook: more
 * column tiles per CTA on small problems), 0 = one CTA per SM. */
int bmnas_set_ws_gemm(int enable, int max_ctas);
int bmnas_get_node_variant(void);

/* rng_state = {seed, step}: advance the step counter on-device (one launch per search step) */
int bmnas_rng_advance(unsigned long long* rng_state, void* stream);

/* Test hook for the in-kernel dropout streams: writes the keep decision (1 = kept) that every fused kernel draws for
 * dropout site `uid` (bmnas_node_params.op_uid / bmnas_ln_params.op_uid) with probability p under rng_state =
 * {seed, step}, for B samples of per_sample (= C*L) elements starting at global sample sample_offset.  A parity test
 * runs a Philox-mode forward/backward, reads the masks of that very step through this entry and hands them to the
 * oracle (nn.Dropout draws, node_operations.py:29,48,89, node_search.py:46 -- bit-matching torch's stream is not a
 * goal; matching the SAME mask in forward, backward and every kernel variant is). */
int bmnas_philox_keep_mask(const unsigned long long* rng_state, unsigned int uid, float p, long long sample_offset, long long per_sample, long long B, unsigned char* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BMNAS_B200_H */
