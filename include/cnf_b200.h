/*
 * cnf_b200.h - C ABI of the B200-native coupling-layer hot path of CategoricalNF.
 *
 * The reference (phlippe/CategoricalNF) is pure Python/PyTorch and has no FFI layer; the
 * boundary it exposes is the Python `FlowLayer` API (layers/flows/flow_layer.py:5-32).  This
 * header is the native boundary that sits directly *under* those modules: one entry point per
 * reference function on the path.  The Python modules in
 * `categoricalnf_b200/layers/` bind these symbols with ctypes and keep the reference's
 * signatures; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - every entry point is `int f(const <args>*, cnf_stream_t)`; 0 = CNF_OK, otherwise an error
 *     code whose text is returned by cnf_last_error_string() (thread local).  No C++ exceptions
 *     cross the boundary.
 *   - pointers named *_host are HOST memory, read synchronously at call time; every other
 *     pointer is DEVICE memory of the current CUDA device and is only touched on `stream`.
 *   - the library never allocates device memory and holds no global mutable state: calls are
 *     re-entrant and capturable in CUDA graphs.  Work is launched asynchronously.
 *   - tensors are dense, row-major, float32 unless stated; B = samples, S = positions per
 *     sample (sequence / nodes / node pairs), C = latent channels, K = mixture components.
 *   - numerical health is reported through an optional device status word (`status`), OR-ed
 *     with CNF_FLAG_* bits; the host checks it lazily instead of the reference's per-layer
 *     `assert torch.isnan(...)` host syncs (mixture_cdf_layer.py:82, flow_model.py:42).
 */
#ifndef CNF_B200_H
#define CNF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* cnf_stream_t;

#if defined(__GNUC__)
#define CNF_API __attribute__((visibility("default")))
#else
#define CNF_API
#endif

enum {
    CNF_OK = 0,
    CNF_ERR_INVALID_ARG = 1,  /* null pointer, bad size, misaligned tensor            */
    CNF_ERR_UNSUPPORTED = 2,  /* shape outside the compiled range (C > 64, K > 256 ..) */
    CNF_ERR_CUDA = 3          /* a CUDA runtime call failed (text has the CUDA error)  */
};

/* bits of the device status word */
#define CNF_FLAG_NAN_Z 1u      /* NaN in an output latent (reference: AssertionError)              */
#define CNF_FLAG_NAN_LDJ 2u    /* NaN in a log-det-Jacobian term                                    */
#define CNF_FLAG_CDF_RANGE 4u  /* inverse CDF input outside (0,1) (reference: RuntimeError, :238)   */

#define CNF_MAX_CHANNELS 64
#define CNF_MAX_MIXTURES 256

CNF_API const char* cnf_last_error_string(void);
/* ABI version (bumped on any struct change) and the SM architecture the kernels were built for. */
CNF_API int cnf_abi_version(void);
CNF_API int cnf_built_for_sm(void);

/* ------------------------------------------------------------------------------------------
 * Coupling masks.  `cond_c_host[c] != 0` marks channel c as conditioner input (mask value 1 in
 * coupling_layer.py:101-112); NULL = no channel is a conditioner.  `cond_s_host` is the chess
 * mask over positions (coupling_layer.py:115-121) of period `s_period` (<= 64), applied as
 * mask[s % s_period] exactly like `_prepare_mask` tiles it (coupling_layer.py:67-74);
 * NULL / 0 = none.  An element is transformed iff neither its channel nor its position is a
 * conditioner; `pad` (channel_padding_mask, [B,S], 1 = real element) multiplies in as in
 * mixture_cdf_layer.py:99-101.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    const float* cond_c_host; /* [C] or NULL */
    const float* cond_s_host; /* [s_period] or NULL */
    int32_t s_period;
} cnf_mask;

/* ------------------------------------------------------------------------------------------
 * K1 / K2  logistic-mixture-CDF coupling transform
 *   replaces MixtureCDFCoupling.get_mixt_params + run_with_params
 *   (layers/flows/mixture_cdf_layer.py:95-142, 145-180, 197-276)
 * nn_out record per channel: [t, log_s, log_pi x K, mu x K, log_scale x K].
 * Forward:  F = mixture CDF(x); y = logit F; z_out = (y + t) e^{log_s};
 *           ldj[b] (+)= sum change * (log_s - log F - log(1-F) + log f (+ reg * reg_factor)).
 * Inverse:  y = z e^{-log_s} - t; F = clamp(sigmoid y, 1e-5, 1-1e-5); x = CDF^-1(F) by
 *           bracketed bisection started at 0; ldj[b] (+)= -sum change * (...).
 * Conditioner / padded elements are copied (times pad) and contribute nothing.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int64_t B, S;
    int32_t C, K;
    const float* z;                      /* [B,S,C]                                   */
    const float* nn_out;                 /* [B,S,C*(2+3K)]                            */
    cnf_mask mask;
    const float* pad;                    /* [B,S] or NULL                             */
    const float* scaling_factor;         /* [C]   log of the tanh bound, or NULL      */
    const float* mixture_scaling_factor; /* [C,K] or NULL                             */
    float reg_max;                       /* <= 0 disables the CDF regulariser (:108)  */
    float reg_factor;
    int32_t training;                    /* regulariser only when training (:108)     */
    int32_t accumulate;                  /* 0: ldj/reg_ldj overwritten, 1: added to   */
    int32_t params_prebounded;           /* 1: log_s / log_scale entries of nn_out are
                                            already tanh-bounded (explicit-parameter callers
                                            of run_with_params); scaling factors ignored */
    float* z_out;                        /* [B,S,C] (may alias z)                     */
    float* ldj;                          /* [B]                                       */
    float* reg_ldj;                      /* [B] or NULL (forward only)                */
    uint32_t* status;                    /* device status word or NULL                */
    /* optional fused epilogue (forward only): the ActNorm and 1x1 convolution of the NEXT flow
     * block, applied to the full output row before the store (saves two passes over z).
     * ldj terms of those layers are per-sample constants and are added by cnf_ldj_axpy. */
    const float* next_actnorm_bias;      /* [C] or NULL                               */
    const float* next_actnorm_scales;    /* [C] or NULL                               */
    const float* next_conv_weight;       /* [C,C] row-major (z @ W) or NULL           */
    /* ABI v4: 1 = `nn_out` is COMPACT, [B,S,Ct*(2+3K)] with only the records of the Ct transformed channels (in
     * channel order; they must form one contiguous run c0..c0+Ct-1).  The conditioner half of the network output is
     * never read, so a caller that owns the final projection need not produce it (training: half the GEMM, half the
     * gradient traffic).  Supported by the TMA pipelines (cnf_mixcdf_path 1 and 2) - query cnf_mixcdf_path with this
     * flag set; the staged generic kernel rejects it. */
    int32_t nn_compact;
} cnf_mixcdf_args;

CNF_API int cnf_mixcdf_fwd(const cnf_mixcdf_args* a, cnf_stream_t stream);
CNF_API int cnf_mixcdf_inv(const cnf_mixcdf_args* a, cnf_stream_t stream);
/* 1 when cnf_mixcdf_fwd can apply the fused next-block epilogue (next_* fields) for this shape, mask
 * and alignment, else 0 (the caller then runs cnf_actnorm / cnf_invconv_apply as separate calls). */
CNF_API int cnf_mixcdf_fusable(const cnf_mixcdf_args* a);
/* Which kernel cnf_mixcdf_fwd / _inv would launch for these arguments (pointers are only checked for alignment):
 * 0 staged generic kernel (mixcdf_kernel), 1 TMA pipeline with compile-time (K, Ct) and one thread per element
 * (mixcdf_pipe_kernel), 2 TMA pipeline with lane groups for any K (mixcdf_gpipe_kernel); -1 invalid arguments. */
CNF_API int cnf_mixcdf_path(const cnf_mixcdf_args* a);

/* ------------------------------------------------------------------------------------------
 * K3  affine coupling  (layers/flows/coupling_layer.py:53-65, 76-98)
 * nn_out record per channel: [s, t]; s = tanh(s / max(e^{sf},1)) e^{sf}.
 * forward z_out = (z + t) e^{s}, ldj += sum s ; inverse z_out = z e^{-s} - t, ldj -= sum s.
 * The ldj is not pad-masked (App. B #4).  `ldj` is always accumulated into.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int64_t B, S;
    int32_t C;
    const float* z;              /* [B,S,C]  */
    const float* nn_out;         /* [B,S,2C] */
    cnf_mask mask;
    const float* scaling_factor; /* [C] or NULL */
    int32_t reverse;
    int32_t params_prebounded;   /* 1: s is already tanh-bounded (explicit run_with_params) */
    float* z_out;                /* [B,S,C] */
    float* ldj;                  /* [B] in/out */
    uint32_t* status;
} cnf_affine_args;

CNF_API int cnf_affine_coupling(const cnf_affine_args* a, cnf_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K4  activation normalisation
 *   ActNormFlow.forward        (layers/flows/activation_normalization.py:24-48)
 *   ExtActNormFlow.forward     (:116-144) with the per-element (bias, raw scale) already
 *                              produced by pred_net: `ext` is [B,S,2C] = [bias | raw scale].
 *   data_init statistics       (:55-67)
 * forward z_out = (z + b) e^{s} * pad ; inverse z_out = (z e^{-s} - b) * pad.
 * ActNorm:    ldj[b] += (+/-) sum_c s_c * len_b,  len_b = length[b] | sum_s pad | S.
 * ExtActNorm: ldj[b] += (+/-) sum_{s,c} tanh(raw)_{s,c} * pad; output is NOT pad-multiplied.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int64_t B, S;
    int32_t C;
    const float* z;       /* [B,S,C] */
    const float* bias;    /* [C] */
    const float* scales;  /* [C] */
    const float* pad;     /* [B,S] or NULL */
    const float* length;  /* [B] float or NULL */
    int32_t reverse;
    float* z_out;         /* [B,S,C] */
    float* ldj;           /* [B] in/out, or NULL to skip */
    uint32_t* status;
} cnf_actnorm_args;

CNF_API int cnf_actnorm(const cnf_actnorm_args* a, cnf_stream_t stream);

typedef struct {
    int64_t B, S;
    int32_t C;
    const float* z;    /* [B,S,C]  */
    const float* ext;  /* [B,S,2C] */
    const float* pad;  /* [B,S] or NULL */
    int32_t reverse;
    float* z_out;
    float* ldj;        /* [B] in/out */
    uint32_t* status;
} cnf_ext_actnorm_args;

CNF_API int cnf_ext_actnorm(const cnf_ext_actnorm_args* a, cnf_stream_t stream);

/* Masked per-channel statistics for the data-dependent init: writes bias = -mean,
 * scales = -0.5 log(var) where var = E[(x + bias)^2] over elements with pad = 1.
 * `workspace` needs 3*C doubles, zeroed by the call. */
typedef struct {
    int64_t B, S;
    int32_t C;
    const float* x;     /* [B,S,C] */
    const float* pad;   /* [B,S] or NULL */
    double* workspace;  /* [3*C] */
    float* bias;        /* [C] out */
    float* scales;      /* [C] out */
} cnf_actnorm_init_args;

CNF_API int cnf_actnorm_data_init(const cnf_actnorm_init_args* a, cnf_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K5  invertible 1x1 convolution  (layers/flows/permutation_layers.py:61-136)
 * build:  W = P (L o strict_lower + I)(U o strict_upper + diag(sign_s e^{log_s})),
 *         sldj = sum log_s, W_inv = inverse(W) in float64 rounded to float32 (:77,:85).
 *         For the non-LU parametrisation pass `weight` and sldj = log|det W| is computed by
 *         an in-kernel LU (:64-65).
 * apply:  z_out = (z @ W) * pad ; ldj[b] += (+/-) sldj * len_b.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int32_t C;
    const float* p;       /* [C,C] or NULL when `weight` given */
    const float* l;       /* [C,C] */
    const float* u;       /* [C,C] */
    const float* log_s;   /* [C]   */
    const float* sign_s;  /* [C]   */
    const float* weight;  /* [C,C] direct parametrisation, or NULL */
    float* w_out;         /* [C,C] */
    float* w_inv_out;     /* [C,C] or NULL */
    float* sldj_out;      /* [1]   */
} cnf_invconv_build_args;

CNF_API int cnf_invconv_build(const cnf_invconv_build_args* a, cnf_stream_t stream);

typedef struct {
    int64_t B, S;
    int32_t C;
    const float* z;       /* [B,S,C] */
    const float* weight;  /* [C,C] (W forward, W^-1 reverse) */
    const float* sldj;    /* [1] device scalar */
    const float* pad;     /* [B,S] or NULL */
    const float* length;  /* [B] float or NULL (-> S) */
    int32_t reverse;
    float* z_out;
    float* ldj;           /* [B] in/out or NULL */
    uint32_t* status;
    /* optional (forward only): the ActNorm of the same flow block applied first, a = (z + b) e^{s} pad
     * (its ldj term sum(s) * len is added by the caller, cnf_ldj_axpy), and a second output
     * z_out * out_mask = the network input of the coupling layer that follows (coupling_layer.py:53). */
    const float* pre_actnorm_bias;    /* [C] or NULL */
    const float* pre_actnorm_scales;  /* [C] or NULL */
    const float* out_mask;            /* [C] or NULL */
    float* z_masked_out;              /* [B,S,C] or NULL */
} cnf_invconv_args;

CNF_API int cnf_invconv_apply(const cnf_invconv_args* a, cnf_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K6  mixture-of-logistics categorical encoding (LinearCategoricalEncoding, num_flows = 0)
 *   (layers/categorical_encoding/linear_encoding.py:59-196, distributions.py:117-163)
 * table[v] = [bias_v (D) | raw scale_v (D)] = pred_net(embed(v)).
 * encode: z0 = logit(u (1-1e-4) + 5e-5)/1.81, z = (z0 + b_x) e^{tanh s_x}; exact posterior over
 *         the V classes; ldj[b] += sum_s pad (beta log q(x|z) - log p(z0) + sum tanh s_x).
 *         Noise is either supplied (`u_noise`, parity mode) or drawn in-kernel from Philox4x32-10
 *         keyed by (seed, offset) when u_noise == NULL.
 * decode: x = argmax_v log p(z|v) + prior_v  (:184-196).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int64_t B, S;
    int32_t V, D;
    const int64_t* tokens;        /* [B,S]   */
    const float* u_noise;         /* [B,S,D] U(0,1) or NULL */
    uint64_t seed, offset;        /* Philox key/counter when u_noise == NULL */
    const float* table;           /* [V,2D]  */
    const float* category_prior;  /* [V] log-softmaxed */
    const float* pad;             /* [B,S] or NULL */
    float beta;
    float* z_out;                 /* [B,S,D] */
    float* ldj;                   /* [B] in/out */
    float* class_prob_log;        /* [B,S] or NULL */
    uint32_t* status;
    /* optional fused epilogue: ActNorm and 1x1 convolution of the FIRST flow block applied to the
     * encoded latent before the store (see cnf_mixcdf_args.next_*); check cnf_categ_encode_fusable. */
    const float* next_actnorm_bias;   /* [D] or NULL */
    const float* next_actnorm_scales; /* [D] or NULL */
    const float* next_conv_weight;    /* [D,D] row-major (z @ W) or NULL */
} cnf_categ_encode_args;

CNF_API int cnf_categ_encode(const cnf_categ_encode_args* a, cnf_stream_t stream);
CNF_API int cnf_categ_encode_fusable(const cnf_categ_encode_args* a);

typedef struct {
    int64_t B, S;
    int32_t V, D;
    const float* z;               /* [B,S,D] */
    const float* table;           /* [V,2D]  */
    const float* category_prior;  /* [V]     */
    int64_t* tokens_out;          /* [B,S]   */
} cnf_categ_decode_args;

CNF_API int cnf_categ_decode(const cnf_categ_decode_args* a, cnf_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K7  logistic prior  (layers/flows/distributions.py:129-163)
 * log_prob: out[b] (+)= sum_{s,c} pad * -(softplus(v) + softplus(-v) + log sigma), v=(x-mu)/sigma
 *           (`elementwise` != NULL additionally stores the unreduced values).
 *           `add` != NULL (accumulate = 0): out[b] = add[b] + log_prob[b] - the per-sample log-likelihood ldj + log p(z)
 *           (general/task.py / experiments/.../task.py `_calc_loss`) finished by this kernel.
 *           `total` != NULL: total[0] = sum_b out[b] in float64, total[1] = B - the (sum log-likelihood, count) pair the
 *           ranks all-reduce once per step (replaces nn.DataParallel's gather, general/mutils.py:243-249), folded into
 *           the epilogue of the kernel that finishes the log-likelihood.
 * sample:   x = logit(u (1-eps) + eps/2) sigma + mu, u from `u_noise` or Philox.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int64_t B, S;
    int32_t C;
    const float* x;      /* [B,S,C] */
    const float* pad;    /* [B,S] or NULL */
    float mu, sigma;
    int32_t accumulate;
    float* out;          /* [B] or NULL */
    float* elementwise;  /* [B,S,C] or NULL */
    const float* add;    /* [B] or NULL; needs out, accumulate = 0, must not alias out (ABI v3) */
    double* total;       /* [2] or NULL; needs out, accumulate = 0 (ABI v3) */
} cnf_logistic_logprob_args;

CNF_API int cnf_logistic_logprob(const cnf_logistic_logprob_args* a, cnf_stream_t stream);

typedef struct {
    int64_t n;
    const float* u_noise; /* [n] or NULL */
    uint64_t seed, offset;
    float mu, sigma, eps;
    float* x_out;         /* [n] */
} cnf_logistic_sample_args;

CNF_API int cnf_logistic_sample(const cnf_logistic_sample_args* a, cnf_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Running ldj accumulator helper (layers/flows/flow_model.py:44): y[b] += alpha * x[b] * len[b]
 * with optional device scalar `alpha_dev` (e.g. sldj); used by the fused block path for the
 * per-sample constants of ActNorm / InvConv.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int64_t B;
    float alpha;
    const float* alpha_dev; /* [1] or NULL */
    const float* x;         /* [B] or NULL (-> 1) */
    const float* length;    /* [B] or NULL (-> 1) */
    float* y;               /* [B] */
} cnf_ldj_axpy_args;

CNF_API int cnf_ldj_axpy(const cnf_ldj_axpy_args* a, cnf_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K8  dense projection of the coupling networks on tcgen05 tensor cores
 *   replaces nn.Linear (y = x W^T + b) inside RGCNNet / EdgeGNN / LinearNet
 *   (layers/networks/graph_layers.py:24-25,64-71,192-202,307-315,402-405,574-577,712-716,766-779;
 *    layers/networks/help_layers.py:57-124), optionally followed by nn.GELU (graph_layers.py:69,176).
 * TMA-fed, accumulators in tensor memory.  precision 0: one TF32 pass; precision 1: 3xTF32 split
 * (hi/lo operand decomposition, fp32-level accuracy, ~1e-6 relative).
 * Requirements: K % 4 == 0, 16-byte aligned x / weight / y (CNF_ERR_UNSUPPORTED / INVALID_ARG otherwise).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int64_t M;            /* rows: positions / nodes / node pairs      */
    int32_t N, K;         /* out_features, in_features                 */
    const float* x;       /* [M,K] row-major                           */
    const float* weight;  /* [N,K] row-major (nn.Linear.weight)        */
    const float* bias;    /* [N] or NULL                               */
    int32_t precision;    /* 0 = TF32, 1 = 3xTF32                      */
    int32_t activation;   /* 0 = none, 1 = GELU (erf form, nn.GELU())  */
    float* y;             /* [M,N]                                     */
    const float* weight_lo; /* [N,K] or NULL.  precision 1 only: the low part of a weight split done once by the caller -
                               `weight` then holds the TF32-representable high part rna_tf32(W) and weight_lo =
                               rna_tf32(W - weight) - so that the kernel splits only the activations                      */
    int32_t block_n;      /* 0 = automatic N tile; else a multiple of 32 <= 256 (tuning)                                 */
} cnf_linear_args;

CNF_API int cnf_linear_fwd(const cnf_linear_args* a, cnf_stream_t stream);

/* Backward of the projection (what autograd derives for F.linear in the reference's training loops,
 * general/train.py:148-152), same tcgen05 kernel with the operands read in place (MN-major shared-memory
 * descriptors instead of transposed copies):
 *   grad_x [M,K]  = grad_y W            (overwritten)
 *   grad_weight [N,K] += grad_y^T x     (ACCUMULATED: the reduction over M is split across the grid and the partial
 *                                        tiles are added with red.global.add - zero it, or keep it to accumulate)
 *   grad_bias [N]     += column sums    (ACCUMULATED)
 * NULL output = not computed.  Requirements: N % 4 == 0, K % 4 == 0, 16-byte aligned tensors. */
typedef struct {
    int64_t M;
    int32_t N, K;
    const float* x;        /* [M,K] (read for grad_weight)     */
    const float* weight;   /* [N,K] (read for grad_x)          */
    const float* grad_y;   /* [M,N]                            */
    int32_t precision;     /* 0 = TF32, 1 = 3xTF32             */
    float* grad_x;         /* [M,K] or NULL                    */
    float* grad_weight;    /* [N,K] or NULL, accumulated into  */
    float* grad_bias;      /* [N] or NULL, accumulated into    */
    /* ABI v5, 3xTF32 only: pre-split weight as in cnf_linear_args.weight_lo - `weight` then holds the high parts
     * rna_tf32(W) and weight_lo = rna_tf32(W - weight); the grad_x product reads both through TMA instead of splitting
     * the weight tile again in every CTA and k-block.  NULL: split in the kernel.                                    */
    const float* weight_lo;
} cnf_linear_bwd_args;

CNF_API int cnf_linear_bwd(const cnf_linear_bwd_args* a, cnf_stream_t stream);

/* Backward of cnf_categ_encode with respect to the class table (what autograd derives through the all-class expansion of
 * linear_encoding.py:71-92,153-174).  `z` is the forward output; the noise is a constant of the graph.
 * grad_table [V,2D] is ACCUMULATED into (zero it first); gradients with respect to embed / pred_net follow from
 * table = pred_net(embed.weight) by ordinary autograd on that [V,2D] product. */
typedef struct {
    int64_t B, S;
    int32_t V, D;
    const int64_t* tokens;        /* [B,S]                          */
    const float* z;               /* [B,S,D] output of the forward  */
    const float* table;           /* [V,2D]                         */
    const float* category_prior;  /* [V] log-softmaxed              */
    const float* pad;             /* [B,S] or NULL                  */
    float beta;
    const float* grad_z;          /* [B,S,D] dL/dz                  */
    const float* grad_ldj;        /* [B] dL/dldj or NULL            */
    float* grad_table;            /* [V,2D], accumulated            */
} cnf_categ_encode_bwd_args;

CNF_API int cnf_categ_encode_bwd(const cnf_categ_encode_bwd_args* a, cnf_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Glue of the graph coupling networks around their projections
 *   RGCNNet / RelationGraphConv / RelationGraphAttention / GNNSkipConnection,
 *   layers/networks/graph_layers.py:15-154,157-235,702-733.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int64_t M;            /* rows                                           */
    int32_t H;            /* normalised (last) dimension                    */
    const float* x;       /* [M,H]                                          */
    const float* gamma;   /* [H] nn.LayerNorm.weight                        */
    const float* beta;    /* [H] nn.LayerNorm.bias                          */
    float eps;            /* 1e-5                                           */
    float* y;             /* [M,H]                                          */
} cnf_layernorm_args;

CNF_API int cnf_layernorm(const cnf_layernorm_args* a, cnf_stream_t stream);

/* Attention logits of RelationGraphAttention (graph_layers.py:92-96):
 *   score_s[m,h]     = sum_d hs[m, h*Dh+d]          * attn_weight[h,0,d]
 *   score_r[m,e,h]   = sum_d hr[m, (e*H+h)*Dh+d]    * attn_weight[h,1,d]      e in 0..E (E = self-connection) */
typedef struct {
    int64_t M;                /* nodes (B*N)                                  */
    int32_t E, H, Dh;         /* edge types, heads, features per head         */
    const float* hs;          /* [M, >= H*Dh], row pitch ld_hs                */
    const float* hr;          /* [M, >= (E+1)*H*Dh], row pitch ld_hr          */
    int64_t ld_hs, ld_hr;
    const float* attn_weight; /* [H,2,Dh]                                     */
    float* score_s;           /* [M,H]                                        */
    float* score_r;           /* [M,(E+1)*H]                                  */
} cnf_graph_attn_scores_args;

CNF_API int cnf_graph_attn_scores(const cnf_graph_attn_scores_args* a, cnf_stream_t stream);

/* Neighbour aggregation from the INTEGER adjacency (0 = no edge, 1..E = edge type; the reference one-hot
 * encodes it first, graph_layers.py:205).
 *   mode 0 (RelationGraphConv.forward :37-50, H = 1, Dh = c_out):
 *       out[i] = hs[i] + sum_{j: adj[j][i] > 0} hr[j, adj[j][i]-1, :] / max(num_neighbours[i], 1e-5)
 *   mode 1 (RelationGraphAttention.forward :98-154): neighbours j of row i plus i itself with edge slot E,
 *       p = softmax_j leaky_relu(score_s[i,h] + score_r[j, e, h]),  out[i,h,:] = sum_j p_j hr[j, e, h, :]
 * activation 1 applies GELU to the result (the nn.GELU opening output_projection, :68-71). */
typedef struct {
    int64_t B;
    int32_t N, E, H, Dh;
    const int64_t* adjacency;     /* [B,N,N]                                           */
    const float* hs;              /* mode 0: [B*N, >= Dh] pitch ld_hs; mode 1: unused  */
    const float* hr;              /* [B*N, >= (E or E+1)*H*Dh] pitch ld_hr             */
    int64_t ld_hs, ld_hr;
    const float* score_s;         /* mode 1: [B*N,H], row pitch ld_score_s             */
    const float* score_r;         /* mode 1: [B*N,(E+1)*H], row pitch ld_score_r       */
    int64_t ld_score_s, ld_score_r; /* 0 = dense; the logits may be extra columns of the projection output (they are
                                       linear in its input: hs.a_0 = x (W_hs^T a_0) + b_hs.a_0)                   */
    const float* num_neighbours;  /* mode 0: [B*N] or NULL (-> number of edges found)  */
    int32_t mode;
    float leaky_slope;            /* 0.2                                               */
    int32_t activation;           /* 0 none, 1 GELU                                    */
    float* out;                   /* [B*N, H*Dh]                                       */
} cnf_graph_aggregate_args;

CNF_API int cnf_graph_aggregate(const cnf_graph_aggregate_args* a, cnf_stream_t stream);

/* GNNSkipConnection.forward (graph_layers.py:722-733); `skip` = skip_layer(feat):
 *   config 0: out = orig + skip                              skip [M,H]
 *   config 1: out = orig + val * sigmoid(gate)               skip [M,2H] = [val | gate]
 *   config 2: out = orig * (1 - sigmoid(gate)) + val * sigmoid(gate) */
typedef struct {
    int64_t M;
    int32_t H, config;
    const float* orig;    /* [M,H]           */
    const float* skip;    /* [M,H] or [M,2H] */
    float* out;           /* [M,H]           */
} cnf_skip_gate_args;

CNF_API int cnf_skip_gate(const cnf_skip_gate_args* a, cnf_stream_t stream);

/* Edge-GNN message passing (layers/networks/graph_layers.py:242-336, 388-700).  Node pairs (a < b) are numbered
 * p = a (N-1) - a (a-1)/2 + (b-a-1) (experiments/molecule_generation/mutils.py:5-10); features of the valid pairs are
 * stored compacted [R,*]; rev[b*P + p] = 1 + compact row, 0 = pair not valid (indices_reverse, graph_layers.py:349-355).
 *   mode 0  Edge2NodeAttnLayer (:595-645): w = sigmoid(edge_logit) / max(sum over valid pairs, 1e-5)
 *   mode 1  Edge2NodeQKVAttnLayer (:432-502): w = softmax(scale * q_i.k_j + edge_logit) over the valid pairs
 *   out[i,h,:] = sum_j w_ij (edge_val[pair(i,j),h,:] + node_val[j,h,:])   (0 for a node without valid pairs) */
typedef struct {
    int64_t B;
    int32_t N, H, Dh;
    int64_t R;                   /* number of compact pair rows                       */
    const int64_t* rev;          /* [B, N(N-1)/2]                                     */
    const float* node_val;       /* [B*N, >= H*Dh] pitch ld_node_val                  */
    const float* node_q;         /* mode 1, pitch ld_node_q                           */
    const float* node_k;         /* mode 1, pitch ld_node_k                           */
    const float* edge_val;       /* [R, >= H*Dh] pitch ld_edge_val                    */
    const float* edge_logit;     /* [R, >= H] pitch ld_edge_logit                     */
    int64_t ld_node_val, ld_node_q, ld_node_k, ld_edge_val, ld_edge_logit;
    int32_t mode;
    float scale;                 /* mode 1: Dh^-0.5                                   */
    float* out;                  /* [B*N, H*Dh]                                       */
} cnf_edge_aggregate_args;

CNF_API int cnf_edge_aggregate(const cnf_edge_aggregate_args* a, cnf_stream_t stream);

/* Node2EdgePlainLayer (graph_layers.py:317-336) on the compact pair rows:
 *   out[r,:] = act(edge_lin[r,:] + node_lin[b, x1[p],:] + node_lin[b, x2[p],:])   with flat_indices[r] = b*P + p */
typedef struct {
    int64_t R;
    int32_t N, He;
    const int64_t* flat_indices;  /* [R]                                   */
    const int64_t* x_indices1;    /* [P] first node of pair p              */
    const int64_t* x_indices2;    /* [P] second node of pair p             */
    const float* edge_lin;        /* [R, >= He] pitch ld_edge              */
    const float* node_lin;        /* [B*N, >= He] pitch ld_node            */
    int64_t ld_edge, ld_node;
    int32_t activation;           /* 0 none, 1 GELU                        */
    float* out;                   /* [R, He]                               */
} cnf_pair_combine_args;

CNF_API int cnf_pair_combine(const cnf_pair_combine_args* a, cnf_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K8 + K1/K2 fused: FINAL projection of the coupling network + mixture-CDF coupling transform.
 *   nn_out = features @ weight^T + bias   (last nn.Linear of the network, e.g.
 *            layers/networks/graph_layers.py:198-201,775-778; help_layers.py:84-94)
 *   followed by cnf_mixcdf_fwd / cnf_mixcdf_inv on that nn_out (mixture_cdf_layer.py:95-180),
 *   without nn_out [B,S,C*(2+3K)] ever being written to memory: the records of the transformed
 *   channels are produced by tcgen05.mma into tensor memory and consumed from there.
 * `mix` is read like in cnf_mixcdf_fwd except that mix.nn_out is ignored (may be NULL); the
 * mix.next_* epilogue (ActNorm + 1x1 conv of the next block) is available for every fusable shape.  Shapes: see cnf_linear_mixcdf_fusable (K in {4,8,16}, 4 or 8
 * contiguous transformed channels, C % 4 == 0, C <= 32, H % 4 == 0, 16-byte aligned tensors).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    cnf_mixcdf_args mix;
    int32_t H;              /* in_features of the final projection           */
    int32_t precision;      /* 0 = TF32, 1 = 3xTF32 (see cnf_linear_args)    */
    const float* features;  /* [B,S,H] input of the final projection         */
    const float* weight;    /* [C*(2+3K), H] nn.Linear.weight                */
    const float* bias;      /* [C*(2+3K)] or NULL                            */
    /* with the mix.next_* epilogue (forward only): the NEXT coupling layer's channel mask (1 = conditioner
     * input) and a second output z_out * next_mask = that layer's network input (coupling_layer.py:53)  */
    const float* next_mask; /* [C] or NULL                                   */
    float* z_masked_out;    /* [B,S,C] or NULL                               */
} cnf_linear_mixcdf_args;

CNF_API int cnf_linear_mixcdf_fwd(const cnf_linear_mixcdf_args* a, cnf_stream_t stream);
CNF_API int cnf_linear_mixcdf_inv(const cnf_linear_mixcdf_args* a, cnf_stream_t stream);
CNF_API int cnf_linear_mixcdf_fusable(const cnf_linear_mixcdf_args* a);


/* ------------------------------------------------------------------------------------------
 * Backward passes: the reference trains by autograd through its
 * eager float64 ops (general/train.py:148-152); these entry points evaluate the same chain rule
 * in one kernel per layer.  `grad_z_out` = dL/d(layer output), `grad_ldj` [B] = dL/d(ldj) (NULL
 * = 0).  Parameter gradients (`grad_scaling_factor`, `grad_bias`, ...) are ACCUMULATED into
 * (+=): the caller zero-initialises them.  All other outputs are overwritten.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int64_t B, S;
    int32_t C, K;
    const float* z;                      /* [B,S,C] forward input                     */
    const float* nn_out;                 /* [B,S,C*(2+3K)] forward input              */
    cnf_mask mask;
    const float* pad;                    /* [B,S] or NULL                             */
    const float* scaling_factor;         /* [C] or NULL                               */
    const float* mixture_scaling_factor; /* [C,K] or NULL                             */
    float reg_max, reg_factor;
    int32_t training;
    int32_t params_prebounded;
    const float* grad_z_out;             /* [B,S,C]                                   */
    const float* grad_ldj;               /* [B] or NULL                               */
    float* grad_z;                       /* [B,S,C] direct path only (not through nn) */
    float* grad_nn_out;                  /* [B,S,C*(2+3K)], zeros for conditioners (NULL allowed with proj_weight) */
    float* grad_scaling_factor;          /* [C] += or NULL                            */
    float* grad_mixture_scaling_factor;  /* [C,K] += or NULL                          */
    /* ABI v4: 1 = nn_out AND grad_nn_out are compact, [B,S,Ct*(2+3K)] (see cnf_mixcdf_args.nn_compact): no zeros are
     * written for conditioner channels.  Needs a contiguous run of transformed channels and 16-byte aligned rows. */
    int32_t nn_compact;
    /* ABI v5 (needs nn_compact): [Ct*(2+3K)] += column sums of grad_nn_out over all positions = dL/dbias of the network's
     * final nn.Linear (the reference gets it from autograd through that Linear, general/train.py:148-152).  The persistent
     * compact-layout kernel sums the columns of every tile while its rows are still in shared memory; otherwise a pass
     * over grad_nn_out follows the kernel.  NULL: not wanted.                                                          */
    float* grad_nn_colsum;
    /* ABI v5 (needs nn_compact, C = 16 with 8 contiguous transformed channels at either end, channel mask only): the
     * coupling network is ONE per-position nn.Linear applied to the (masked) block input z, nn_out = (z * mask) W^T + b,
     * and proj_weight [Ct*(2+3K), C] holds the weight rows of the transformed channels' records.  The kernel then also
     * evaluates that Linear's backward from the gradient tile in shared memory, in fp32:
     *   grad_z[pos, conditioner channels] += grad_nn_out[pos, :] . proj_weight[:, conditioner channels]
     *   grad_proj_weight[n, conditioner channels] += sum_pos grad_nn_out[pos, n] z[pos, conditioner channels]   (may be NULL)
     * (columns of masked-out inputs receive nothing: the reference's `z * mask` gives them a zero gradient).  grad_nn_out
     * may then be NULL: the [B,S,Ct*(2+3K)] gradient never leaves the chip.                                               */
    const float* proj_weight;
    float* grad_proj_weight;
} cnf_mixcdf_bwd_args;

/* forward direction of cnf_mixcdf_fwd (training differentiates the density direction only) */
CNF_API int cnf_mixcdf_bwd(const cnf_mixcdf_bwd_args* a, cnf_stream_t stream);

typedef struct {
    int64_t B, S;
    int32_t C;
    const float* z;              /* [B,S,C]  forward input */
    const float* nn_out;         /* [B,S,2C] */
    cnf_mask mask;
    const float* scaling_factor; /* [C] or NULL */
    int32_t reverse;
    int32_t params_prebounded;
    const float* grad_z_out;     /* [B,S,C] */
    const float* grad_ldj;       /* [B] or NULL */
    float* grad_z;               /* [B,S,C] */
    float* grad_nn_out;          /* [B,S,2C] */
    float* grad_scaling_factor;  /* [C] += or NULL */
} cnf_affine_bwd_args;

CNF_API int cnf_affine_coupling_bwd(const cnf_affine_bwd_args* a, cnf_stream_t stream);

typedef struct {
    int64_t B, S;
    int32_t C;
    const float* z;        /* [B,S,C] forward input */
    const float* bias;     /* [C] */
    const float* scales;   /* [C] */
    const float* pad;      /* [B,S] or NULL */
    const float* length;   /* [B] or NULL */
    int32_t reverse;
    const float* grad_z_out;
    const float* grad_ldj; /* [B] or NULL */
    float* grad_z;
    float* grad_bias;      /* [C] += or NULL */
    float* grad_scales;    /* [C] += or NULL */
} cnf_actnorm_bwd_args;

CNF_API int cnf_actnorm_bwd(const cnf_actnorm_bwd_args* a, cnf_stream_t stream);

typedef struct {
    int64_t B, S;
    int32_t C;
    const float* z;        /* [B,S,C]  forward input */
    const float* ext;      /* [B,S,2C] */
    const float* pad;      /* [B,S] or NULL */
    int32_t reverse;
    const float* grad_z_out;
    const float* grad_ldj; /* [B] or NULL */
    float* grad_z;
    float* grad_ext;       /* [B,S,2C] */
} cnf_ext_actnorm_bwd_args;

CNF_API int cnf_ext_actnorm_bwd(const cnf_ext_actnorm_bwd_args* a, cnf_stream_t stream);

typedef struct {
    int64_t B, S;
    int32_t C;
    const float* z;        /* [B,S,C] forward input */
    const float* weight;   /* [C,C] the matrix that was applied (W forward, W^-1 reverse) */
    const float* pad;      /* [B,S] or NULL */
    const float* length;   /* [B] or NULL (-> S) */
    int32_t reverse;
    const float* grad_z_out;
    const float* grad_ldj; /* [B] or NULL */
    float* grad_z;
    float* grad_weight;    /* [C,C] += or NULL : gradient w.r.t. the applied matrix */
    float* grad_sldj;      /* [1] += or NULL */
} cnf_invconv_bwd_args;

CNF_API int cnf_invconv_bwd(const cnf_invconv_bwd_args* a, cnf_stream_t stream);

typedef struct {
    int64_t B, S;
    int32_t C;
    const float* x;                 /* [B,S,C] */
    const float* pad;               /* [B,S] or NULL (applies to grad_out only) */
    float mu, sigma;
    const float* grad_out;          /* [B] gradient of the per-sample sums, or NULL */
    const float* grad_elementwise;  /* [B,S,C] gradient of the element-wise values, or NULL */
    float* grad_x;                  /* [B,S,C] */
} cnf_logistic_logprob_bwd_args;

CNF_API int cnf_logistic_logprob_bwd(const cnf_logistic_logprob_bwd_args* a, cnf_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * SigmoidFlow (layers/flows/sigmoid_layer.py:24-48) and the two ends of VariationalDequantization
 * (layers/categorical_encoding/variational_dequantization.py:31-58).
 *   reverse == 0:  z_out = sigmoid(z),                 ldj_e = -z - 2 softplus(-z)
 *   reverse == 1:  y = z (1-alpha) + alpha/2,  z_out = log y - log(1-y),
 *                  ldj_e = -log y - log(1-y) + log(1-alpha)          (alpha = 1e-5 upstream)
 * `reverse` is the EFFECTIVE direction (the module resolves `reverse_layer XOR reverse`, :29).
 * ldj[b] (+)= sum of ldj_e over the sample (`accumulate`); `ldj_elementwise` additionally stores the
 * unreduced values (sum_ldj=False, :43-46).  `add_tokens` (int64, same element count) is added to
 * z_out: the dequantised value `z.float() + noise` of variational_dequantization.py:47.
 * cnf_dequant_floor: tokens_out = clamp(floor(z), 0, V-1) (:55-56).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int64_t B, n_per_sample;
    const float* z;              /* [B, n_per_sample] */
    int32_t reverse;
    float alpha;
    int32_t accumulate;
    const int64_t* add_tokens;   /* [B, n_per_sample] or NULL */
    float* z_out;                /* [B, n_per_sample] */
    float* ldj;                  /* [B] or NULL */
    float* ldj_elementwise;      /* [B, n_per_sample] or NULL */
    uint32_t* status;            /* CNF_FLAG_* word or NULL */
} cnf_sigmoid_flow_args;

CNF_API int cnf_sigmoid_flow(const cnf_sigmoid_flow_args* a, cnf_stream_t stream);

typedef struct {
    int64_t B, n_per_sample;
    const float* z;                   /* [B, n_per_sample] forward INPUT */
    int32_t reverse;
    float alpha;
    const float* grad_z_out;          /* [B, n_per_sample] or NULL */
    const float* grad_ldj;            /* [B] or NULL */
    const float* grad_ldj_elementwise;/* [B, n_per_sample] or NULL */
    float* grad_z;                    /* [B, n_per_sample] */
} cnf_sigmoid_flow_bwd_args;

CNF_API int cnf_sigmoid_flow_bwd(const cnf_sigmoid_flow_bwd_args* a, cnf_stream_t stream);

typedef struct {
    int64_t n;
    int32_t V;
    const float* z;        /* [n] */
    int64_t* tokens_out;   /* [n] */
} cnf_dequant_floor_args;

CNF_API int cnf_dequant_floor(const cnf_dequant_floor_args* a, cnf_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Backward of the graph coupling networks' glue (csrc/graph_ops_bwd.cu): training GraphNodeFlow / GraphCNF
 * differentiates RGCNNet / EdgeGNN (layers/networks/graph_layers.py) through these instead of autograd over dense
 * [B,N,N,...] torch expressions.  `fwd` repeats the arguments of the forward call (its `out` is ignored); gradients
 * marked "+=" receive scattered contributions through fp32 reductions and must be zero-initialised by the caller,
 * the others are overwritten.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int64_t n;
    const float* x;        /* [n] input of the GELU                                   */
    const float* grad_y;   /* NULL: y = gelu(x);  else: y = grad_y * gelu'(x)         */
    float* y;              /* [n]                                                     */
} cnf_gelu_args;

CNF_API int cnf_gelu(const cnf_gelu_args* a, cnf_stream_t stream);

typedef struct {
    int64_t M;
    int32_t H;
    const float* x;        /* [M,H] forward input    */
    const float* gamma;    /* [H]                    */
    float eps;
    const float* grad_y;   /* [M,H]                  */
    float* grad_x;         /* [M,H]                  */
    float* grad_gamma;     /* [H] += or NULL         */
    float* grad_beta;      /* [H] += or NULL         */
} cnf_layernorm_bwd_args;

CNF_API int cnf_layernorm_bwd(const cnf_layernorm_bwd_args* a, cnf_stream_t stream);

typedef struct {
    int64_t M;
    int32_t H, config;
    const float* orig;      /* [M,H]            */
    const float* skip;      /* [M,H] or [M,2H]  */
    const float* grad_out;  /* [M,H]            */
    float* grad_orig;       /* [M,H]            */
    float* grad_skip;       /* like skip        */
} cnf_skip_gate_bwd_args;

CNF_API int cnf_skip_gate_bwd(const cnf_skip_gate_bwd_args* a, cnf_stream_t stream);

typedef struct {
    cnf_graph_aggregate_args fwd;
    const float* grad_out;       /* [B*N, H*Dh] gradient of the (activated) output              */
    float* grad_hs;              /* mode 0: [B*N, >= Dh] pitch ld_grad_hs (overwritten); mode 1: unused  */
    float* grad_hr;              /* += , pitch ld_grad_hr                                        */
    float* grad_score_s;         /* mode 1: [B*N,H] pitch ld_grad_score_s (overwritten)          */
    float* grad_score_r;         /* mode 1: += , pitch ld_grad_score_r                           */
    int64_t ld_grad_hs, ld_grad_hr, ld_grad_score_s, ld_grad_score_r;
} cnf_graph_aggregate_bwd_args;

CNF_API int cnf_graph_aggregate_bwd(const cnf_graph_aggregate_bwd_args* a, cnf_stream_t stream);

typedef struct {
    cnf_edge_aggregate_args fwd;
    const float* grad_out;       /* [B*N, H*Dh]                  */
    float* grad_node_val;        /* += , pitch ld_grad_node_val  */
    float* grad_node_q;          /* mode 1: +=                   */
    float* grad_node_k;          /* mode 1: +=                   */
    float* grad_edge_val;        /* += [R, >= H*Dh]              */
    float* grad_edge_logit;      /* += [R, >= H]                 */
    int64_t ld_grad_node_val, ld_grad_node_q, ld_grad_node_k, ld_grad_edge_val, ld_grad_edge_logit;
} cnf_edge_aggregate_bwd_args;

CNF_API int cnf_edge_aggregate_bwd(const cnf_edge_aggregate_bwd_args* a, cnf_stream_t stream);

typedef struct {
    cnf_pair_combine_args fwd;
    const float* grad_out;       /* [R, He]                                  */
    float* grad_edge_lin;        /* [R, >= He] pitch ld_grad_edge (overwritten) */
    float* grad_node_lin;        /* += [B*N, >= He] pitch ld_grad_node       */
    int64_t ld_grad_edge, ld_grad_node;
} cnf_pair_combine_bwd_args;

CNF_API int cnf_pair_combine_bwd(const cnf_pair_combine_bwd_args* a, cnf_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CNF_B200_H */
