// Element-level arithmetic of the logistic-mixture-CDF coupling (shared by the generic and the
// TMA-pipelined kernels).  Restates mixture_cdf_layer.py:95-142, 197-276 of the reference:
// fp32 fast path with MUFU ex2/lg2/rcp, float64 restatement for the rare elements whose CDF leaves
// the range where fp32 reproduces the reference's float64 arithmetic to 1e-4.
#pragma once
#include "cnf_common.cuh"

namespace cnf {
namespace mixmath {

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

static __device__ __noinline__ SlowOut mix_eval_f64(double x, const float* rec, const float* mfac, int K) {
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

static __device__ __noinline__ ElemResult mix_forward_f64(float x, const float* rec, const float* mfac, int K, float log_s,
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
static __device__ __noinline__ ElemResult mix_inverse_f64(float zin, float x0, float margin, const float* rec,
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
// "Prepared" form for compile-time K (TMA-pipelined kernels): everything that depends on the
// parameters only (bounded log-scales -> e^{-ls}, softmax numerators, their normaliser) is computed
// once per element and held in registers; an evaluation at x then costs 2 MUFU per component (one
// sigmoid).  The forward transform evaluates once, the inverse up to ~49 times.  log2(e) factors
// are folded into the per-component constants so the evaluation loop has no scaling multiplies.
// ------------------------------------------------------------------------------------------------
// Packed fp32 pairs in the per-component loops (cnf_common.cuh: f2_*); -DCNF_NO_F32X2 builds the scalar form for A/B runs.
#ifdef CNF_NO_F32X2
constexpr bool kPackedPairs = false;
#else
constexpr bool kPackedPairs = true;
#endif

template <int KT>
struct MixPrep {
    float mu[KT];
    float einv2[KT];  // e^{-ls_k} * log2(e)
    float w[KT];      // softmax numerators of log_pi
    float iw;         // 1 / sum w
    float t, log_s;
    float span;       // sum_k e^{ls_k}, inverse only (bracket of the bisection)
};

// `bnd[k * BSTRIDE]` holds, per component, (2 log2e / max(e^{msf},1), -e^{msf} log2e) - see
// mixture_cdf_layer.py:160-162.
template <int KT, int BSTRIDE, bool WANT_SPAN>
__device__ __forceinline__ void mix_prepare(MixPrep<KT>& P, const float* rec, const float2* bnd, float fac, float a2) {
    const float* lp = rec + 2;
    const float* mu = lp + KT;
    const float* ms = mu + KT;
    float m = lp[0];
#pragma unroll
    for (int k = 1; k < KT; ++k) m = fmaxf(m, lp[k]);
    const float m_l2 = m * kLog2e;
    float W = 0.f, span = 0.f;
    if constexpr (kPackedPairs && KT % 2 == 0) {
        // two components per instruction (FMUL2 / FADD2 / FFMA2); the MUFU calls stay scalar
        const f2 one = f2_splat(1.0f), mtwo = f2_splat(-2.0f), l2e = f2_splat(kLog2e), nm = f2_splat(-m_l2);
        f2 W2 = f2_splat(0.f);
#pragma unroll
        for (int k = 0; k < KT; k += 2) {
            const float2 b0 = bnd[k * BSTRIDE], b1 = bnd[(k + 1) * BSTRIDE];
            float t0, t1;
            f2_get(f2_mul(f2_make(ms[k], ms[k + 1]), f2_make(b0.x, b1.x)), t0, t1);
            const f2 r = f2_make(rcp(1.0f + ex2(t0)), rcp(1.0f + ex2(t1)));          // tanh = 1 - 2 / (1 + 2^v)
            float n0, n1;
            f2_get(f2_mul(f2_fma(mtwo, r, one), f2_make(b0.y, b1.y)), n0, n1);        // -ls_k * log2(e)
            float e0, e1;
            f2_get(f2_mul(f2_make(ex2(n0), ex2(n1)), l2e), e0, e1);
            P.einv2[k] = e0;
            P.einv2[k + 1] = e1;
            if (WANT_SPAN) span += ex2(-n0) + ex2(-n1);
            P.mu[k] = mu[k];
            P.mu[k + 1] = mu[k + 1];
            float a0, a1;
            f2_get(f2_fma(f2_make(lp[k], lp[k + 1]), l2e, nm), a0, a1);
            P.w[k] = ex2(a0);
            P.w[k + 1] = ex2(a1);
            W2 = f2_add(W2, f2_make(P.w[k], P.w[k + 1]));
        }
        float w0, w1;
        f2_get(W2, w0, w1);
        W = w0 + w1;
    } else {
#pragma unroll
        for (int k = 0; k < KT; ++k) {
            const float2 bk = bnd[k * BSTRIDE];
            const float nls2 = tanh_from_2log2e(ms[k] * bk.x) * bk.y;   // -ls_k * log2(e)
            P.einv2[k] = ex2(nls2) * kLog2e;
            if (WANT_SPAN) span += ex2(-nls2);
            P.mu[k] = mu[k];
            P.w[k] = ex2(fmaf(lp[k], kLog2e, -m_l2));
            W += P.w[k];
        }
    }
    P.iw = rcp(W);
    P.span = span;
    P.t = rec[0];
    P.log_s = tanh_from_2log2e(rec[1] * a2) * fac;
}

template <int KT>
__device__ __forceinline__ MixEval mix_eval_p(float x, const MixPrep<KT>& P) {
    float Fs = 0.f, Gs = 0.f, fs = 0.f;
    if constexpr (kPackedPairs && KT % 2 == 0) {
        const f2 xx = f2_splat(x), one = f2_splat(1.0f), mone = f2_splat(-1.0f);
        f2 F2 = f2_splat(0.f), G2 = f2_splat(0.f), d2 = f2_splat(0.f);
#pragma unroll
        for (int k = 0; k < KT; k += 2) {
            const f2 ei = f2_make(P.einv2[k], P.einv2[k + 1]);
            const f2 w = f2_make(P.w[k], P.w[k + 1]);
            float u0, u1;
            f2_get(f2_mul(f2_fma(f2_make(P.mu[k], P.mu[k + 1]), mone, xx), ei), u0, u1);   // u * log2(e)
            const float e0 = ex2(-fabsf(u0)), e1 = ex2(-fabsf(u1));
            const f2 e = f2_make(e0, e1);
            float s0, s1;
            f2_get(f2_add(e, one), s0, s1);
            const float r0 = rcp(s0), r1 = rcp(s1);
            const f2 r = f2_make(r0, r1);
            const f2 q = f2_mul(e, r);                                                    // min(sigma, 1 - sigma)
            float q0, q1;
            f2_get(q, q0, q1);
            const bool p0 = u0 >= 0.0f, p1 = u1 >= 0.0f;
            F2 = f2_fma(w, f2_make(p0 ? r0 : q0, p1 ? r1 : q1), F2);
            G2 = f2_fma(w, f2_make(p0 ? q0 : r0, p1 ? q1 : r1), G2);
            d2 = f2_fma(f2_mul(w, f2_mul(q, r)), ei, d2);
        }
        float a, b;
        f2_get(F2, a, b); Fs = a + b;
        f2_get(G2, a, b); Gs = a + b;
        f2_get(d2, a, b); fs = a + b;
    } else {
#pragma unroll
        for (int k = 0; k < KT; ++k) {
            const float u2 = (x - P.mu[k]) * P.einv2[k];   // u * log2(e)
            const float e = ex2(-fabsf(u2));
            const float r = rcp(1.0f + e);
            const float q = e * r;  // min(sigma, 1 - sigma)
            const bool pos = u2 >= 0.0f;
            Fs = fmaf(P.w[k], pos ? r : q, Fs);
            Gs = fmaf(P.w[k], pos ? q : r, Gs);
            fs = fmaf(P.w[k] * (q * r), P.einv2[k], fs);
        }
    }
    MixEval o;
    o.F = Fs * P.iw;
    o.G = Gs * P.iw;
    o.f = fs * (P.iw * kLn2);
    return o;
}

// fp32 is trusted while F, 1-F and f are far from underflow and 1-F is not in the region where the
// reference's own float64 `1/F - 1` loses digits; NaNs fail the test as well.
__device__ __forceinline__ bool mix_fast_ok(const MixEval& e) { return e.F >= 1e-30f && e.G >= 1e-12f && e.f >= 1e-30f; }

template <int KT>
__device__ __forceinline__ ElemResult mix_forward_fast(const MixEval& e, const MixPrep<KT>& P, bool use_reg,
                                                       float reg_max, float reg_factor) {
    const float lF = fast_log(e.F), lG = fast_log(e.G);
    const float lFc = fmaxf(lF, kLog1em22), lGc = fmaxf(lG, kLog1em22);
    ElemResult r;
    r.reg = 0.f;
    if (use_reg) r.reg = (fminf(lFc * kInvLn10, -reg_max) + reg_max) + (fminf(lGc * kInvLn10, -reg_max) + reg_max);
    r.z = (lF - lG + P.t) * fast_exp(P.log_s);
    r.ldj = P.log_s - lFc - lGc + fast_log(e.f) + r.reg * reg_factor;
    return r;
}

// Inverse in fp32 registers (safeguarded Newton, see below).  Returns false when the element must be
// finished by mix_inverse_f64 (flat CDF around the root); x / lb / ub then describe where to resume.
template <int KT>
struct InvState {
    float x, lb, ub, lb0, ub0, cond, f, mixt_ldj;
};

template <int KT>
__device__ __forceinline__ bool mix_inverse_fast(float zin, const MixPrep<KT>& P, uint32_t* status, InvState<KT>& st,
                                                 ElemResult& out) {
    const float log_s = P.log_s;
    const float y = fmaf(zin, fast_exp(-log_s), -P.t);
    st.mixt_ldj = softplus_pm(y);
    const float ey = ex2(-fabsf(y) * kLog2e);
    const float ry = rcp(1.0f + ey);
    float Ft = y >= 0.f ? ry : ey * ry;  // target CDF
    float Gt = y >= 0.f ? ey * ry : ry;  // 1 - target
    if (!(Ft == Ft)) flag(status, CNF_FLAG_CDF_RANGE);
    Ft = fminf(fmaxf(Ft, 1e-5f), 1.0f - 1e-5f);   // reference clamp (:130)
    Gt = fminf(fmaxf(Gt, 1e-5f), 1.0f - 1e-5f);
    const bool upper = Ft > 0.5f;  // compare on the smaller of F and 1-F: relative accuracy in both tails
    float lb = INFINITY, ub = -INFINITY;
#pragma unroll
    for (int k = 0; k < KT; ++k) {   // bracket (:252-254)
        lb = fminf(lb, fmaf(-20.f, P.span, P.mu[k]));
        ub = fmaxf(ub, fmaf(20.f, P.span, P.mu[k]));
    }
    st.lb0 = lb; st.ub0 = ub;
    // Root of logit F(x) = logit(target) by Newton's method in logit space (for one logistic the
    // function is linear there, for mixtures close to it), safeguarded by the reference's bracket
    // ("rtsafe"): a step that leaves [lb, ub] or fails to halve the previous one is replaced by the
    // reference's bisection step (:255-262).  The root is unique (the CDF is strictly increasing), so
    // this lands on the point the reference's float64 bisection approaches.
    const float yc = (lg2(Ft) - lg2(Gt)) * kLn2;   // logit of the clamped target
    float x = 0.f, prev = INFINITY;
    MixEval e = mix_eval_p<KT>(x, P);
    for (int it = 0; it < 48; ++it) {
        const bool gt = upper ? (e.G < Gt) : (e.F > Ft);
        if (gt) ub = x; else lb = x;
        const float mid = 0.5f * (lb + ub);
        const float sloc = e.F * e.G * rcp(e.f);                             // local logistic scale 1 / (logit F)'
        const float step = ((lg2(e.F) - lg2(e.G)) * kLn2 - yc) * sloc;
        const float xt = x - step;
        // quadratic convergence: the error after a Newton step is about step^2 / sloc.  A converged
        // step is taken even if rounding noise put it a hair outside the bracket.
        // Newton needs a density that has not underflowed: with f flushed to zero and 1-F still a normal
        // number, sloc and step are +inf and `inf <= inf` would accept x = -inf as converged.
        const bool fin = e.f >= 1e-30f && fabsf(step) < 1e30f;
        const bool conv = fin && step * step <= 2.5e-8f * fmaxf(1.0f, fabsf(xt)) * sloc;
        const bool newton = conv || (fin && xt >= lb && xt <= ub && fabsf(step) < 0.5f * prev);
        const float xn = newton ? xt : mid;
        prev = newton ? fabsf(step) : (ub - lb);
        // a bisection ends like the reference's, when the bracket has collapsed
        const bool done = conv || (!newton && !(ub - lb > 5e-7f * fmaxf(1.0f, fabsf(xn))));
        const bool stuck = xn == x;
        x = xn;
        if (stuck) break;
        e = mix_eval_p<KT>(x, P);
        if (done) break;
    }
    st.x = x; st.lb = lb; st.ub = ub; st.f = e.f;
    // fp32 CDF values carry ~5e-7 relative error -> root error ~5e-7 min(F,G)/f
    st.cond = fminf(e.F, e.G);
    if (!(st.cond <= 8.0f * fmaxf(1.0f, fabsf(x)) * e.f) || !(e.f >= 1e-30f)) return false;
    out.z = x;
    out.ldj = -(log_s + st.mixt_ldj + fast_log(e.f));
    out.reg = 0.f;
    return true;
}

template <int KT>
__device__ __forceinline__ float inv_slow_margin(const InvState<KT>& st) {
    return fmaxf(4.0f * (st.ub - st.lb) + 1e-5f * fmaxf(1.0f, fabsf(st.x)), 1e-5f * st.cond / fmaxf(st.f, 1e-37f));
}


// ------------------------------------------------------------------------------------------------
// Group-cooperative prepared form for a RUN-TIME number of components (generic kernel): G = 2^g
// adjacent lanes share one element, lane `sub` owns the components sub, sub + G, ... (at most NC of
// them, held in registers as a MixPrep<NC>); partial sums meet in xor-butterflies over the group's
// lanes, which leave bit-identical values in every lane, so the Newton iteration below stays
// uniform inside a group.  G = 1 degenerates to one thread per element.  Same arithmetic as the
// compile-time form above (mix_prepare / mix_eval_p / mix_inverse_fast).
// ------------------------------------------------------------------------------------------------
struct LaneGroup {
    unsigned mask;   // lanes of this group
    int G, sub;
    __device__ __forceinline__ float sum(float v) const {
        for (int d = G >> 1; d > 0; d >>= 1) v += __shfl_xor_sync(mask, v, d);
        return v;
    }
    __device__ __forceinline__ float max(float v) const {
        for (int d = G >> 1; d > 0; d >>= 1) v = fmaxf(v, __shfl_xor_sync(mask, v, d));
        return v;
    }
    __device__ __forceinline__ float min(float v) const {
        for (int d = G >> 1; d > 0; d >>= 1) v = fminf(v, __shfl_xor_sync(mask, v, d));
        return v;
    }
};

template <int NC>
__device__ __forceinline__ void mix_prepare_g(MixPrep<NC>& P, const ElemCtx& c, const LaneGroup& g) {
    const int K = c.K;
    const float* lp = c.rec + 2;
    const float* mu = lp + K;
    const float* ms = mu + K;
    float m = -INFINITY;
#pragma unroll
    for (int i = 0; i < NC; ++i) {
        const int k = g.sub + i * g.G;
        if (k < K) m = fmaxf(m, lp[k]);
    }
    const float m_l2 = g.max(m) * kLog2e;
    float W = 0.f, span = 0.f;
#pragma unroll
    for (int i = 0; i < NC; ++i) {
        const int k = g.sub + i * g.G;
        const bool v = k < K;
        const int kk = v ? k : g.sub;   // G <= K: component `sub` always exists; padding slots copy it with weight 0
        const float nls2 = c.pre ? -ms[kk] * kLog2e : tanh_from_2log2e(ms[kk] * c.ma2[kk]) * (-c.mfac[kk] * kLog2e);
        P.einv2[i] = ex2(nls2) * kLog2e;
        P.mu[i] = mu[kk];
        const float w = v ? ex2(fmaf(lp[kk], kLog2e, -m_l2)) : 0.f;
        P.w[i] = w;
        W += w;
        if (v) span += ex2(-nls2);
    }
    P.iw = rcp(g.sum(W));
    P.span = g.sum(span);
    P.t = c.rec[0];
    P.log_s = c.pre ? c.rec[1] : tanh_from_2log2e(c.rec[1] * c.a2) * c.fac;
}

template <int NC>
__device__ __forceinline__ MixEval mix_eval_g(float x, const MixPrep<NC>& P, const LaneGroup& g) {
    MixEval e = mix_eval_p<NC>(x, P);
    e.F = g.sum(e.F);
    e.G = g.sum(e.G);
    e.f = g.sum(e.f);
    return e;
}

// mix_inverse_fast for a lane group (see there for the method).
template <int NC>
__device__ __forceinline__ bool mix_inverse_g(float zin, const MixPrep<NC>& P, const LaneGroup& g, uint32_t* status,
                                              InvState<NC>& st, ElemResult& out) {
    const float log_s = P.log_s;
    const float y = fmaf(zin, fast_exp(-log_s), -P.t);
    st.mixt_ldj = softplus_pm(y);
    const float ey = ex2(-fabsf(y) * kLog2e);
    const float ry = rcp(1.0f + ey);
    float Ft = y >= 0.f ? ry : ey * ry;
    float Gt = y >= 0.f ? ey * ry : ry;
    if (!(Ft == Ft)) flag(status, CNF_FLAG_CDF_RANGE);
    Ft = fminf(fmaxf(Ft, 1e-5f), 1.0f - 1e-5f);   // reference clamp (mixture_cdf_layer.py:130)
    Gt = fminf(fmaxf(Gt, 1e-5f), 1.0f - 1e-5f);
    const bool upper = Ft > 0.5f;
    float lb = INFINITY, ub = -INFINITY;
#pragma unroll
    for (int k = 0; k < NC; ++k) {   // bracket (:252-254)
        lb = fminf(lb, fmaf(-20.f, P.span, P.mu[k]));
        ub = fmaxf(ub, fmaf(20.f, P.span, P.mu[k]));
    }
    lb = g.min(lb);
    ub = g.max(ub);
    st.lb0 = lb; st.ub0 = ub;
    const float yc = (lg2(Ft) - lg2(Gt)) * kLn2;
    float x = 0.f, prev = INFINITY;
    MixEval e = mix_eval_g<NC>(x, P, g);
    for (int it = 0; it < 48; ++it) {
        const bool gt = upper ? (e.G < Gt) : (e.F > Ft);
        if (gt) ub = x; else lb = x;
        const float mid = 0.5f * (lb + ub);
        const float sloc = e.F * e.G * rcp(e.f);
        const float step = ((lg2(e.F) - lg2(e.G)) * kLn2 - yc) * sloc;
        const float xt = x - step;
        const bool fin = e.f >= 1e-30f && fabsf(step) < 1e30f;
        const bool conv = fin && step * step <= 2.5e-8f * fmaxf(1.0f, fabsf(xt)) * sloc;
        const bool newton = conv || (fin && xt >= lb && xt <= ub && fabsf(step) < 0.5f * prev);
        const float xn = newton ? xt : mid;
        prev = newton ? fabsf(step) : (ub - lb);
        const bool done = conv || (!newton && !(ub - lb > 5e-7f * fmaxf(1.0f, fabsf(xn))));
        const bool stuck = xn == x;
        x = xn;
        if (stuck) break;
        e = mix_eval_g<NC>(x, P, g);
        if (done) break;
    }
    st.x = x; st.lb = lb; st.ub = ub; st.f = e.f;
    st.cond = fminf(e.F, e.G);
    if (!(st.cond <= 8.0f * fmaxf(1.0f, fabsf(x)) * e.f) || !(e.f >= 1e-30f)) return false;
    out.z = x;
    out.ldj = -(log_s + st.mixt_ldj + fast_log(e.f));
    out.reg = 0.f;
    return true;
}

}  // namespace mixmath
}  // namespace cnf
