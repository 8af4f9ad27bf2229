// K1 / K2: logistic-mixture-CDF coupling transform, forward and inverse, for sm_100a.
//
// Replaces MixtureCDFCoupling.get_mixt_params + run_with_params
// (reference layers/flows/mixture_cdf_layer.py:95-142, 145-180, 197-276): ~100 eager float64
// ATen launches per call there, one kernel here.
//
// Work decomposition: a CTA owns a tile of TP consecutive positions.  The parameter records of
// the *transformed* channels only (the conditioner half of nn_out is never read) and the z rows
// are staged in shared memory with cp.async (16-byte when the layout allows), then one thread
// evaluates one (position, channel) element: K mixture components in registers, fp32 with MUFU
// ex2/lg2/rcp.  Elements whose CDF leaves the range where fp32 reproduces the reference's fp64
// arithmetic to 1e-4 take a float64 path that restates the reference formulas literally.
// Per-sample ldj: shared-memory per-position partials -> warp-segmented sum -> one global
// atomicAdd per (CTA, sample).
#include "cnf_common.cuh"

namespace cnf {
namespace {

constexpr int kThreads = 256;

struct MixParams {
    const float* z;
    const float* nn;
    const float* pad;
    const float* sf;
    const float* msf;
    float* z_out;
    float* ldj;
    float* reg_ldj;
    uint32_t* status;
    const float* nx_bias;
    const float* nx_scales;
    const float* nx_w;
    long long P;  // B * S positions
    int S, C, K, PN, TP;
    MaskView mask;
    int vec_params;  // parameter rows can be copied as 16-byte chunks
    float reg_max, reg_factor;
    int use_reg;
    int reverse;
    int pre;  // parameters already tanh-bounded
};

// Mixture parameters of one element, as staged in shared memory.
struct ElemCtx {
    const float* rec;   // [t, log_s, log_pi[K], mu[K], raw_log_scale[K]]
    const float* mfac;  // [K] e^{msf}
    const float* ma2;   // [K] 2 log2(e) / max(e^{msf}, 1)
    float fac, a2;      // same for the output log-scale
    int K;
    bool pre;           // log-scales are already bounded: skip the tanh
};

struct MixEval {
    float F, G, f;  // CDF, 1-CDF, PDF (fast path, linear domain)
};

// ------------------------------------------------------------------------------------------------
// fp32 evaluation of CDF / survival / density of the mixture at x.
// sigma_k = sigmoid(u_k) is formed from e = exp(-|u_k|), r = 1/(1+e): sigma = r or e r, so neither
// tail cancels; weights are softmax numerators normalised once at the end.
// ------------------------------------------------------------------------------------------------
template <int KT>
__device__ __forceinline__ MixEval mix_eval(float x, const ElemCtx& c, float m_l2) {
    const int K = KT > 0 ? KT : c.K;
    const float* lp = c.rec + 2;
    const float* mu = lp + K;
    const float* ms = mu + K;
    float W = 0.f, Fs = 0.f, Gs = 0.f, fs = 0.f;
#pragma unroll(KT > 0 ? KT : 4)
    for (int k = 0; k < K; ++k) {
        const float ls = c.pre ? ms[k] : tanh_from_2log2e(ms[k] * c.ma2[k]) * c.mfac[k];
        const float einv = ex2(-ls * kLog2e);  // exp(-log_scale)
        const float u = (x - mu[k]) * einv;
        const float e = ex2(-fabsf(u) * kLog2e);
        const float r = rcp(1.0f + e);
        const float q = e * r;  // min(sigma, 1 - sigma)
        const bool pos = u >= 0.0f;
        const float w = ex2(fmaf(lp[k], kLog2e, -m_l2));
        W += w;
        Fs = fmaf(w, pos ? r : q, Fs);
        Gs = fmaf(w, pos ? q : r, Gs);
        fs = fmaf(w * (q * r), einv, fs);
    }
    const float iw = rcp(W);
    MixEval o;
    o.F = Fs * iw;
    o.G = Gs * iw;
    o.f = fs * iw;
    return o;
}

template <int KT>
__device__ __forceinline__ float logit_max_l2(const ElemCtx& c) {
    const int K = KT > 0 ? KT : c.K;
    const float* lp = c.rec + 2;
    float m = lp[0];
#pragma unroll(KT > 0 ? KT : 4)
    for (int k = 1; k < K; ++k) m = fmaxf(m, lp[k]);
    return m * kLog2e;
}

// ------------------------------------------------------------------------------------------------
// float64 restatement of the reference formulas (mixture_cdf_layer.py:201-232, 267-276) for the
// rare elements outside the fp32-safe range.  Parameters are bounded in fp32 first, exactly as
// get_mixt_params does before its .double() (:157-178).
// ------------------------------------------------------------------------------------------------
struct SlowOut {
    double log_cdf, log_pdf;
};

__device__ __noinline__ SlowOut mix_eval_f64(double x, const float* rec, const float* mfac, int K) {
    const float* lp = rec + 2;
    const float* mu = lp + K;
    const float* ms = mu + K;
    float m = lp[0];
    for (int k = 1; k < K; ++k) m = fmaxf(m, lp[k]);
    double se = 0.0;
    for (int k = 0; k < K; ++k) se += exp((double)lp[k] - (double)m);
    const double lse = (double)m + log(se);
    double amax = -INFINITY, asum = 0.0, bmax = -INFINITY, bsum = 0.0;
    for (int k = 0; k < K; ++k) {
        const float fk = mfac ? mfac[k] : 1.0f;
        const double ls = mfac ? (double)(tanhf(ms[k] / fmaxf(fk, 1.0f)) * fk) : (double)ms[k];
        const double u = (x - (double)mu[k]) * exp(-ls);
        const double lpi = (double)lp[k] - lse;
        const double a = lpi + (fmin(u, 0.0) - log1p(exp(-fabs(u))));   // log_pi + logsigmoid(u)
        const double sp = u > 20.0 ? u : log1p(exp(u));                  // F.softplus, threshold 20
        const double b = lpi + u - ls - 2.0 * sp;
        if (a > amax) { asum = asum * exp(amax - a) + 1.0; amax = a; } else { asum += exp(a - amax); }
        if (b > bmax) { bsum = bsum * exp(bmax - b) + 1.0; bmax = b; } else { bsum += exp(b - bmax); }
    }
    SlowOut o;
    o.log_cdf = amax + log(asum);
    o.log_pdf = bmax + log(bsum);
    return o;
}

struct ElemResult {
    float z;    // transformed value
    float ldj;  // log_s + mixt_ldj + log f (+ reg * reg_factor), sign already applied for reverse
    float reg;  // CDF regulariser term (forward, training)
};

__device__ __noinline__ ElemResult mix_forward_f64(float x, const float* rec, const float* mfac, int K, float log_s,
                                                   bool use_reg, float reg_max, float reg_factor) {
    const SlowOut s = mix_eval_f64((double)x, rec, mfac, K);
    const double F = exp(s.log_cdf);
    const double lF = log(fmax(F, 1e-22)), lG = log(fmax(1.0 - F, 1e-22));
    const double y = -log(fmax(1.0 / F - 1.0, 1e-22));
    double reg = 0.0;
    if (use_reg) {
        const double il10 = 1.0 / log(10.0);
        reg = (fmin(lF * il10, -(double)reg_max) + (double)reg_max) + (fmin(lG * il10, -(double)reg_max) + (double)reg_max);
    }
    ElemResult r;
    r.z = (float)((y + (double)rec[0]) * exp((double)log_s));
    r.ldj = (float)((double)log_s - lF - lG + s.log_pdf + reg * (double)reg_factor);
    r.reg = (float)reg;
    return r;
}

template <int KT>
__device__ __forceinline__ ElemResult mix_forward_elem(float x, const ElemCtx& c, bool use_reg, float reg_max,
                                                       float reg_factor) {
    const float log_s = c.pre ? c.rec[1] : tanh_from_2log2e(c.rec[1] * c.a2) * c.fac;
    const float m_l2 = logit_max_l2<KT>(c);
    const MixEval e = mix_eval<KT>(x, c, m_l2);
    // fp32 is trusted while F, 1-F and f are far from underflow and 1-F is not in the region
    // where the reference's own float64 `1/F - 1` loses digits; NaNs fail the test as well.
    if (!(e.F >= 1e-30f && e.G >= 1e-12f && e.f >= 1e-30f)) {
        return mix_forward_f64(x, c.rec, c.pre ? nullptr : c.mfac, KT > 0 ? KT : c.K, log_s, use_reg, reg_max, reg_factor);
    }
    const float lF = fast_log(e.F), lG = fast_log(e.G);
    const float lFc = fmaxf(lF, kLog1em22), lGc = fmaxf(lG, kLog1em22);
    ElemResult r;
    r.reg = 0.f;
    if (use_reg) r.reg = (fminf(lFc * kInvLn10, -reg_max) + reg_max) + (fminf(lGc * kInvLn10, -reg_max) + reg_max);
    r.z = (lF - lG + c.rec[0]) * fast_exp(log_s);
    r.ldj = log_s - lFc - lGc + fast_log(e.f) + r.reg * reg_factor;
    return r;
}

// ------------------------------------------------------------------------------------------------
// Inverse: y -> F = clamp(sigmoid y) -> x = CDF^-1(F) by bisection (mixture_cdf_layer.py:124-136,
// 235-264).  The reference bisects every element in float64 until the batch-wide max step is
// <= 1e-10 (~45 halvings, one host sync each).  Here each element bisects in fp32 registers from
// the same start (x = 0) and bracket until its bracket stops shrinking.  Where the CDF is so flat
// that fp32 evaluation error would move the root by more than the parity tolerance
// (min(F, 1-F) / f large), the element is finished in float64 with a bracketed Newton iteration,
// which converges to the root the reference's float64 bisection approaches.
// ------------------------------------------------------------------------------------------------
__device__ __noinline__ ElemResult mix_inverse_f64(float zin, float x0, float margin, const float* rec,
                                                   const float* mfac, int K, float log_s, float lb0, float ub0) {
    const double y = (double)zin * exp(-(double)log_s) - (double)rec[0];
    double target = 1.0 / (1.0 + exp(-y));
    target = fmin(fmax(target, 1e-5), 1.0 - 1e-5);
    const double mixt_ldj = fabs(y) + 2.0 * log1p(exp(-fabs(y)));
    double x = (double)x0, lo = fmax((double)lb0, x - (double)margin), hi = fmin((double)ub0, x + (double)margin);
    SlowOut s = mix_eval_f64(x, rec, mfac, K);
    for (int it = 0; it < 12; ++it) {
        const double Fx = exp(s.log_cdf);
        if (Fx > target) hi = x; else lo = x;
        double xn = x - (Fx - target) * exp(-s.log_pdf);
        if (!(xn > lo && xn < hi)) xn = 0.5 * (lo + hi);
        const bool done = fabs(xn - x) <= 1e-11 * fmax(1.0, fabs(x));
        x = xn;
        s = mix_eval_f64(x, rec, mfac, K);
        if (done) break;
    }
    ElemResult r;
    r.z = (float)x;
    r.ldj = (float)(-((double)log_s + mixt_ldj + s.log_pdf));
    r.reg = 0.f;
    return r;
}

template <int KT>
__device__ __forceinline__ ElemResult mix_inverse_elem(float zin, const ElemCtx& c, uint32_t* status) {
    const int K = KT > 0 ? KT : c.K;
    const float log_s = c.pre ? c.rec[1] : tanh_from_2log2e(c.rec[1] * c.a2) * c.fac;
    const float y = fmaf(zin, fast_exp(-log_s), -c.rec[0]);
    const float mixt_ldj = softplus_pm(y);
    // sigmoid in a form that keeps both tails, then the reference clamp to [1e-5, 1-1e-5] (:130)
    const float ey = ex2(-fabsf(y) * kLog2e);
    const float ry = rcp(1.0f + ey);
    float Ft = y >= 0.f ? ry : ey * ry;  // target CDF
    float Gt = y >= 0.f ? ey * ry : ry;  // 1 - target
    if (!(Ft == Ft)) flag(status, CNF_FLAG_CDF_RANGE);
    Ft = fminf(fmaxf(Ft, 1e-5f), 1.0f - 1e-5f);
    Gt = fminf(fmaxf(Gt, 1e-5f), 1.0f - 1e-5f);
    const bool upper = Ft > 0.5f;  // compare on the smaller of F and 1-F: relative accuracy in both tails

    // bracket (mixture_cdf_layer.py:252-254)
    const float* lp = c.rec + 2;
    const float* mu = lp + K;
    const float* ms = mu + K;
    float span = 0.f;
#pragma unroll(KT > 0 ? KT : 4)
    for (int k = 0; k < K; ++k) span += fast_exp(c.pre ? ms[k] : tanh_from_2log2e(ms[k] * c.ma2[k]) * c.mfac[k]);
    float lb = INFINITY, ub = -INFINITY;
#pragma unroll(KT > 0 ? KT : 4)
    for (int k = 0; k < K; ++k) {
        lb = fminf(lb, fmaf(-20.f, span, mu[k]));
        ub = fmaxf(ub, fmaf(20.f, span, mu[k]));
    }
    const float lb0 = lb, ub0 = ub;
    const float m_l2 = logit_max_l2<KT>(c);
    float x = 0.f;
    MixEval e = mix_eval<KT>(x, c, m_l2);
    for (int it = 0; it < 48; ++it) {
        const bool gt = upper ? (e.G < Gt) : (e.F > Ft);
        if (gt) ub = x; else lb = x;
        const float xn = 0.5f * (lb + ub);
        const bool done = (xn == x) || !(ub - lb > 5e-7f * fmaxf(1.0f, fabsf(xn)));
        x = xn;
        e = mix_eval<KT>(x, c, m_l2);
        if (done) break;
    }
    // fp32 CDF values carry ~5e-7 relative error -> root error ~5e-7 min(F,G)/f
    const float cond = fminf(e.F, e.G);
    if (!(cond <= 8.0f * fmaxf(1.0f, fabsf(x)) * e.f) || !(e.f >= 1e-30f)) {
        const float margin = fmaxf(4.0f * (ub - lb) + 1e-5f * fmaxf(1.0f, fabsf(x)), 1e-5f * cond / fmaxf(e.f, 1e-37f));
        return mix_inverse_f64(zin, x, margin, c.rec, c.pre ? nullptr : c.mfac, K, log_s, lb0, ub0);
    }
    ElemResult r;
    r.z = x;
    r.ldj = -(log_s + mixt_ldj + fast_log(e.f));
    r.reg = 0.f;
    return r;
}

// ------------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------------
template <int KT, bool REV>
__global__ void __launch_bounds__(kThreads) mixcdf_kernel(const MixParams p) {
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x;
    const int K = KT > 0 ? KT : p.K;
    const int PN = 2 + 3 * K;
    const int C = p.C, Ct = p.mask.n_t, TP = p.TP;
    const int L = Ct * PN;  // parameter floats per position (transformed channels only)

    float* s_par = smem;                                   // [TP * L]
    float* s_z = s_par + ((TP * L + 3) & ~3);              // [TP * C]
    float* s_fac = s_z + ((TP * C + 3) & ~3);              // [Ct] e^{sf}
    float* s_a2 = s_fac + Ct;                              // [Ct] 2 log2e / max(e^{sf},1)
    float* s_mfac = s_a2 + Ct;                             // [Ct * K]
    float* s_ma2 = s_mfac + Ct * K;                        // [Ct * K]
    float* s_ldj = s_ma2 + Ct * K;                         // [TP]
    float* s_reg = s_ldj + TP;                             // [TP]

    const long long pos0 = (long long)blockIdx.x * TP;
    const int rows = (int)min((long long)TP, p.P - pos0);

    // ---- stage parameter rows and z rows -------------------------------------------------------
    if (p.vec_params) {
        const int L4 = L >> 2, total = rows * L4;
        const float inv = 1.0f / (float)L4;
        const float4* src = reinterpret_cast<const float4*>(p.nn);
        const long long row4 = ((long long)C * PN) >> 2;
        const int off4 = (p.mask.c0 * PN) >> 2;
        float4* dst = reinterpret_cast<float4*>(s_par);
        for (int i = tid; i < total; i += kThreads) {
            const int r = fast_div(i, inv), q = i - r * L4;
            cp_async16(dst + i, src + (pos0 + r) * row4 + off4 + q);
        }
    } else {
        const int total = rows * L;
        const float inv_pn = 1.0f / (float)PN, inv_ct = 1.0f / (float)Ct;
        for (int i = tid; i < total; i += kThreads) {
            const int e = fast_div(i, inv_pn), pp = i - e * PN;
            const int r = fast_div(e, inv_ct), j = e - r * Ct;
            cp_async4(s_par + i, p.nn + ((pos0 + r) * C + p.mask.tch[j]) * (long long)PN + pp);
        }
    }
    {
        const float* src = p.z + pos0 * C;
        const int n = rows * C;
        if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
            const int n4 = n >> 2;
            for (int i = tid; i < n4; i += kThreads) cp_async16(s_z + 4 * i, src + 4 * i);
            for (int i = (n4 << 2) + tid; i < n; i += kThreads) cp_async4(s_z + i, src + i);
        } else {
            for (int i = tid; i < n; i += kThreads) cp_async4(s_z + i, src + i);
        }
    }
    // ---- per-channel tanh bounds (mixture_cdf_layer.py:157-162) --------------------------------
    for (int i = tid; i < Ct; i += kThreads) {
        const float fac = p.sf ? expf(p.sf[p.mask.tch[i]]) : 1.0f;
        s_fac[i] = fac;
        s_a2[i] = 2.0f * kLog2e / fmaxf(fac, 1.0f);
    }
    for (int i = tid; i < Ct * K; i += kThreads) {
        const int j = i / K, k = i - j * K;
        const float fac = p.msf ? expf(p.msf[p.mask.tch[j] * K + k]) : 1.0f;
        s_mfac[i] = fac;
        s_ma2[i] = 2.0f * kLog2e / fmaxf(fac, 1.0f);
    }
    for (int i = tid; i < TP; i += kThreads) { s_ldj[i] = 0.f; s_reg[i] = 0.f; }
    cp_async_wait_all();
    __syncthreads();

    // ---- one thread per (position, transformed channel) ----------------------------------------
    const int nelem = rows * Ct;
    const float inv_ct = 1.0f / (float)Ct;
    for (int e = tid; e < nelem; e += kThreads) {
        const int r = fast_div(e, inv_ct), j = e - r * Ct;
        const long long pos = pos0 + r;
        if (p.mask.s_period > 0) {
            const int s = (int)(pos % p.S);
            if ((p.mask.cond_s >> (s % p.mask.s_period)) & 1ull) continue;  // conditioner position
        }
        const float padv = p.pad ? p.pad[pos] : 1.0f;
        if (padv == 0.0f) continue;  // padded: copied through (times 0) below, no ldj
        const int ch = p.mask.tch[j];
        ElemCtx c;
        c.rec = s_par + (size_t)e * PN;
        c.mfac = s_mfac + j * K;
        c.ma2 = s_ma2 + j * K;
        c.fac = s_fac[j];
        c.a2 = s_a2[j];
        c.K = K;
        c.pre = p.pre != 0;
        const float x = s_z[r * C + ch];
        ElemResult res;
        if constexpr (!REV) res = mix_forward_elem<KT>(x, c, p.use_reg != 0, p.reg_max, p.reg_factor);
        else res = mix_inverse_elem<KT>(x, c, p.status);
        // z_out = out * change + x * (1 - change) with change = pad (mixture_cdf_layer.py:137-138)
        s_z[r * C + ch] = (padv == 1.0f) ? res.z : fmaf(res.z, padv, x * (1.0f - padv));
        atomicAdd(&s_ldj[r], res.ldj * padv);
        if (p.use_reg) atomicAdd(&s_reg[r], res.reg * padv);
        uint32_t bad = 0u;
        if (res.z != res.z) bad |= CNF_FLAG_NAN_Z;
        if (res.ldj != res.ldj) bad |= CNF_FLAG_NAN_LDJ;
        flag(p.status, bad);
    }
    __syncthreads();

    // ---- per-sample ldj: warp-segmented sum over the tile's positions --------------------------
    for (int r0 = (tid & ~31); r0 < TP; r0 += kThreads) {
        const int r = r0 + (tid & 31);
        const bool valid = r < rows;
        const long long b = valid ? (pos0 + r) / p.S : 0;
        warp_segmented_atomic_add(p.ldj, b, valid ? s_ldj[r] : 0.f, valid);
        if (p.use_reg && p.reg_ldj) warp_segmented_atomic_add(p.reg_ldj, b, valid ? s_reg[r] : 0.f, valid);
    }

    // ---- store z rows (times pad, mixture_cdf_layer.py:76) -------------------------------------
    {
        float* dst = p.z_out + pos0 * C;
        const int n = rows * C;
        const float inv_c = 1.0f / (float)C;
        if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0 && (C & 3) == 0) {
            const int n4 = n >> 2, c4 = C >> 2;
            const float inv_c4 = 1.0f / (float)c4;
            for (int i = tid; i < n4; i += kThreads) {
                float4 v = *reinterpret_cast<const float4*>(s_z + 4 * i);
                if (p.pad) {
                    const float pv = p.pad[pos0 + fast_div(i, inv_c4)];
                    v.x *= pv; v.y *= pv; v.z *= pv; v.w *= pv;
                }
                stg_stream4(reinterpret_cast<float4*>(dst) + i, v);
            }
        } else {
            for (int i = tid; i < n; i += kThreads) {
                float v = s_z[i];
                if (p.pad) v *= p.pad[pos0 + fast_div(i, inv_c)];
                dst[i] = v;
            }
        }
    }
}

size_t smem_bytes(int TP, int L, int C, int Ct, int K) {
    size_t f = ((size_t)TP * L + 3) & ~(size_t)3;
    f += ((size_t)TP * C + 3) & ~(size_t)3;
    f += 2 * (size_t)Ct + 2 * (size_t)Ct * K + 2 * (size_t)TP;
    return f * sizeof(float);
}

template <int KT, bool REV>
int launch2(const MixParams& p, size_t smem, cudaStream_t stream) {
    if (smem > 48 * 1024)
        CNF_CUDA(cudaFuncSetAttribute(mixcdf_kernel<KT, REV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long grid = (p.P + p.TP - 1) / p.TP;
    mixcdf_kernel<KT, REV><<<(unsigned)grid, kThreads, smem, stream>>>(p);
    return launch_status(REV ? "mixcdf_inv_kernel" : "mixcdf_fwd_kernel");
}

template <int KT>
int launch(const MixParams& p, size_t smem, cudaStream_t stream) {
    return p.reverse ? launch2<KT, true>(p, smem, stream) : launch2<KT, false>(p, smem, stream);
}

int run(const cnf_mixcdf_args* a, cnf_stream_t stream_, int reverse) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    CNF_REQUIRE(a != nullptr, "args is NULL");
    CNF_REQUIRE(a->B >= 0 && a->S >= 0, "negative batch/sequence size");
    CNF_REQUIRE(a->C >= 1 && a->K >= 1, "C and K must be >= 1 (got C=%d K=%d)", a->C, a->K);
    CNF_SUPPORTED(a->C <= CNF_MAX_CHANNELS, "C=%d exceeds CNF_MAX_CHANNELS=%d", a->C, CNF_MAX_CHANNELS);
    CNF_SUPPORTED(a->K <= CNF_MAX_MIXTURES, "K=%d exceeds CNF_MAX_MIXTURES=%d", a->K, CNF_MAX_MIXTURES);
    CNF_SUPPORTED(a->S < (1ll << 31) && a->B * a->S < (1ll << 40), "tensor too large");
    CNF_SUPPORTED(a->next_actnorm_bias == nullptr && a->next_actnorm_scales == nullptr && a->next_conv_weight == nullptr,
                  "fused next-block epilogue is not available in this build");
    MixParams p{};
    int rc = build_mask(a->mask, a->C, &p.mask);
    if (rc != CNF_OK) return rc;
    if (a->B == 0) return CNF_OK;
    CNF_REQUIRE(a->ldj != nullptr, "ldj is NULL");
    if (!a->accumulate) {
        CNF_CUDA(cudaMemsetAsync(a->ldj, 0, sizeof(float) * (size_t)a->B, stream));
        if (a->reg_ldj) CNF_CUDA(cudaMemsetAsync(a->reg_ldj, 0, sizeof(float) * (size_t)a->B, stream));
    }
    const long long P = a->B * a->S;
    if (P == 0) return CNF_OK;
    CNF_REQUIRE(a->z && a->nn_out && a->z_out, "z / nn_out / z_out is NULL");
    const size_t nz = (size_t)P * a->C;
    if (p.mask.n_t == 0) {  // nothing is transformed: copy (times pad) - degenerate but legal
        if (a->z_out != a->z) CNF_CUDA(cudaMemcpyAsync(a->z_out, a->z, nz * sizeof(float), cudaMemcpyDeviceToDevice, stream));
        CNF_SUPPORTED(a->pad == nullptr, "mask with no transformed channel together with a padding mask");
        return CNF_OK;
    }
    p.z = a->z; p.nn = a->nn_out; p.pad = a->pad; p.sf = a->scaling_factor; p.msf = a->mixture_scaling_factor;
    p.z_out = a->z_out; p.ldj = a->ldj; p.reg_ldj = a->reg_ldj; p.status = a->status;
    p.P = P; p.S = (int)a->S; p.C = a->C; p.K = a->K; p.PN = 2 + 3 * a->K;
    p.reg_max = a->reg_max; p.reg_factor = a->reg_factor;
    p.use_reg = (!reverse && a->reg_max > 0.f && a->training) ? 1 : 0;
    p.reverse = reverse;
    p.pre = a->params_prebounded;
    const int Ct = p.mask.n_t, L = Ct * p.PN;
    // tile: about one element per thread, bounded by ~44 KB of parameter staging
    int elems = kThreads;
    const int cap = (44 * 1024) / (4 * p.PN);
    if (elems > cap) elems = cap;
    int TP = elems / Ct;
    TP &= ~3;
    if (TP < 4) TP = 4;
    p.TP = TP;
    const long long rowlen = (long long)a->C * p.PN;
    p.vec_params = p.mask.contiguous && (L % 4 == 0) && (rowlen % 4 == 0) && ((p.mask.c0 * p.PN) % 4 == 0) &&
                   ((reinterpret_cast<uintptr_t>(a->nn_out) & 15) == 0);
    const size_t smem = smem_bytes(TP, L, a->C, Ct, a->K);
    CNF_SUPPORTED(smem <= 200 * 1024, "C=%d K=%d needs %zu bytes of shared memory per tile", a->C, a->K, smem);
    switch (a->K) {
        case 4: return launch<4>(p, smem, stream);
        case 8: return launch<8>(p, smem, stream);
        case 16: return launch<16>(p, smem, stream);
        default: return launch<0>(p, smem, stream);
    }
}

}  // namespace
}  // namespace cnf

extern "C" int cnf_mixcdf_fwd(const cnf_mixcdf_args* a, cnf_stream_t stream) { return cnf::run(a, stream, 0); }
extern "C" int cnf_mixcdf_inv(const cnf_mixcdf_args* a, cnf_stream_t stream) { return cnf::run(a, stream, 1); }
