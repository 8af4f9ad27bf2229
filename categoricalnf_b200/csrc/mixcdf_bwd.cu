// Backward of K1 (logistic-mixture-CDF coupling, forward direction): gradients of a loss with respect to
// z, the network output nn_out and the two scaling-factor parameters, given dL/dz_out and dL/dldj.
//
// The reference differentiates ~100 eager float64 ops per coupling through autograd
// (layers/flows/mixture_cdf_layer.py:95-180 under general/train.py:148-152); here the chain rule of
// SURVEY.md Appendix C is evaluated in one kernel:
//   u_k = (x - mu_k) e^{-ls_k}, sigma_k = sigmoid(u_k), pi = softmax(log_pi)
//   F = sum pi sigma, G = 1 - F = sum pi (1 - sigma), f = sum pi sigma (1 - sigma) e^{-ls}
//   y = log F - log G, z_out = (y + t) e^{log_s}, ldj_e = log_s - log F - log G + log f (+ reg * reg_factor)
// with log_s and ls_k tanh-bounded by the learned scaling factors (:157-162).
//
// One CTA per tile of TP positions, any K / C / mask: the parameter records of the transformed channels,
// the z rows and the incoming dL/dz_out rows are staged with cp.async; one thread per (position, channel)
// makes two passes over the K components (totals, then per-component gradients) and overwrites the record
// IN PLACE with its gradient; the tile's full dL/dnn_out rows (zeros for conditioner channels) are then
// written with coalesced 8-byte stores.  Scaling-factor gradients are reduced in shared memory and leave
// with one atomic per (CTA, parameter).  Elements outside the fp32-safe range (same test as the forward
// kernels) are differentiated in float64.
//
// Two kernels share the element function mix_backward_elem:
//   mixcdf_bwd_kernel       the tile kernel described above (any layout / mask)
//   mixcdf_bwd_pipe_kernel  compact layout (nn_out / dL/dnn_out hold the transformed channels' records only): persistent
//                           CTAs, tiles moved by cp.async.bulk into two shared-memory buffers and out by bulk stores,
//                           per-CTA tables and atomics, optional column sums of dL/dnn_out (the final Linear's bias
//                           gradient) and - PROJ = true, the network is one per-position Linear on z - that Linear's
//                           grad_x / grad_W products from the gradient tile in shared memory, so that dL/dnn_out never
//                           reaches HBM (cnf_mixcdf_bwd_args.grad_nn_colsum / proj_weight, ABI v5).
#include <stdlib.h>

#include "cnf_common.cuh"
#include "tc_ptx.cuh"

namespace cnf {

int tc_colsum(const float* gy, float* gb, long long M, int N, cudaStream_t stream);      // linear_tc.cu: gb[n] += sum_m gy[m,n]

namespace {

constexpr int kThreads = 256;

// shared -> global bulk copy (TMA engine, no tensor map), tracked by the thread's bulk async-group
__device__ __forceinline__ void bulk_store_rows(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(tc::smem_addr(ssrc)), "r"(bytes)
                 : "memory");
}

struct BwdParams {
    const float* z;
    const float* nn;
    const float* pad;
    const float* sf;
    const float* msf;
    const float* gz_out;
    const float* gldj;
    float* gz;
    float* gnn;
    float* gsf;
    float* gmsf;
    float* gcol;      // [Ct * PN] += column sums of dL/dnn_out (= the bias gradient of a final Linear), pipe kernel only
    const float* proj_w;   // [Ct * PN, C] weight block of a final Linear whose input is z itself (pipe kernel, C 16 / Ct 8)
    float* gproj_w;        // [Ct * PN, C] += dL/dnn_out^T z (conditioner columns)
    long long P;
    int S, C, K, PN, TP;
    MaskView mask;
    int vec_params, vec_out;
    int compact;      // nn / gnn hold the transformed channels' records only: [P, Ct * PN] (needs vec_params && vec_out)
    int bulk;         // compact layout + 16-byte aligned rows: mixcdf_bwd_pipe_kernel (persistent CTAs, cp.async.bulk)
    int nbuf;         // its shared-memory buffers per CTA
    float reg_max, reg_factor;
    int use_reg;
    int pre;
};

template <typename T>
struct Mth;
template <>
struct Mth<float> {
    static __device__ __forceinline__ float ex(float v) { return fast_exp(v); }
    static __device__ __forceinline__ float lg(float v) { return fast_log(v); }
    static __device__ __forceinline__ float th(float v) { return tanh_from_2log2e(v * (2.0f * kLog2e)); }
    static __device__ __forceinline__ float rc(float v) { return rcp(v); }     // 1 ulp: no IEEE division sequences
};
template <>
struct Mth<double> {
    static __device__ __forceinline__ double ex(double v) { return exp(v); }
    static __device__ __forceinline__ double lg(double v) { return log(v); }
    static __device__ __forceinline__ double th(double v) { return tanh(v); }
    static __device__ __forceinline__ double rc(double v) { return 1.0 / v; }
};

// Gradient of one transformed element.  `rec` (shared memory) holds [t, raw log_s, log_pi[K], mu[K],
// raw log_scale[K]] on entry and the gradient with respect to those entries on exit; ``gmsf_k[K]`` receives the
// element's contribution to d/d mixture_scaling_factor (written only when the parameters are not pre-bounded).
// Returns false (nothing written) when T = float and the element is outside the fp32-safe range.
template <typename T, int KT>
__device__ __forceinline__ bool mix_backward_elem(float xf, float* rec, const float* mfac, const float* imf, float sfac, int Krt, bool pre,
                                                  float g_out_f, float gl_f, bool use_reg, float reg_max, float reg_factor,
                                                  float* gx_out, float* gsf_out, float* gmsf_k) {
    using M = Mth<T>;
    const int K = KT > 0 ? KT : Krt;
    const T x = (T)xf, g_out = (T)g_out_f, gl = (T)gl_f;
    float* lp = rec + 2;
    float* mu = lp + K;
    float* ms = mu + K;
    float m = lp[0];
#pragma unroll(KT > 0 ? KT : 4)
    for (int k = 1; k < K; ++k) m = fmaxf(m, lp[k]);
    // The record is read (and, in the second pass, overwritten) in 8-byte pieces when K is even: lp / mu / ms start at even
    // offsets of the 8-byte aligned record, and with records 2 + 3K floats apart the 16 lanes of a half warp then hit 32
    // different banks - scalar accesses at that pitch are 2-way conflicted (35 M conflicts per launch at the LM shape).
    constexpr bool kPairs = KT > 0 && (KT % 2) == 0;
    T W = 0, Fs = 0, Gs = 0, fs = 0;
    auto pass1 = [&](int k, float ms_k, float mu_k, float lp_k) {
        const float mf = pre ? 1.0f : mfac[k];
        // the reference bounds the raw log-scales in float32 before its .double() (:157-178)
        // imf[k] = 1 / max(e^{msf}, 1): a per-(channel, component) constant, tabulated by the caller for the fp32 path (it
        // was a MUFU.RCP per component and pass: 16 of ~127 MUFU operations per element)
        const T ls = pre ? (T)ms_k : M::th((T)ms_k * (imf ? (T)imf[k] : M::rc((T)fmaxf(mf, 1.0f)))) * (T)mf;
        const T e = M::ex(-ls);
        const T u = (x - (T)mu_k) * e;
        const T ea = M::ex(-fabs(u));
        const T r = M::rc((T)1 + ea);
        const T q = ea * r;
        const T sg = u >= 0 ? r : q, tg = u >= 0 ? q : r;   // sigma, 1 - sigma
        const T w = M::ex((T)lp_k - (T)m);
        W += w;
        Fs += w * sg;
        Gs += w * tg;
        fs += w * (q * r) * e;
    };
    if constexpr (kPairs) {
#pragma unroll
        for (int k = 0; k < KT; k += 2) {
            const float2 a = *reinterpret_cast<const float2*>(ms + k), b = *reinterpret_cast<const float2*>(mu + k),
                         c = *reinterpret_cast<const float2*>(lp + k);
            pass1(k, a.x, b.x, c.x);
            pass1(k + 1, a.y, b.y, c.y);
        }
    } else {
#pragma unroll(KT > 0 ? KT : 4)
        for (int k = 0; k < K; ++k) pass1(k, ms[k], mu[k], lp[k]);
    }
    const T iw = M::rc(W);
    const T F = Fs * iw, G = Gs * iw, f = fs * iw;
    if (sizeof(T) == sizeof(float)) {
        if (!(F >= (T)1e-30 && G >= (T)1e-12 && f >= (T)1e-30)) return false;
    }
    const float sf_max = fmaxf(sfac, 1.0f);
    const T ths = pre ? (T)0 : M::th((T)rec[1] * M::rc((T)sf_max));
    const T log_s = pre ? (T)rec[1] : ths * (T)sfac;
    const T c22 = (T)-50.65687204586900;   // log(1e-22), safe_log clamp (:197-198)
    const T lF = M::lg(F), lG = M::lg(G);
    const bool cF = lF >= c22, cG = lG >= c22;           // clamp of -log(max(F,1e-22)) - log(max(1-F,1e-22))
    const bool cy = (lG - lF) >= c22;                    // clamp of y = -log(max(1/F - 1, 1e-22))
    const T es = M::ex(log_s);
    const T y = cy ? lF - lG : -c22;
    const T z_out = (y + (T)rec[0]) * es;
    const T A = cy ? g_out * es : (T)0;                  // dL/dy
    const T g_t = g_out * es;
    const T g_logs = g_out * z_out + gl;
    T LlF = A - (cF ? gl : (T)0), LlG = -A - (cG ? gl : (T)0);
    if (use_reg) {   // reg = min(log10 F, -reg_max) + reg_max + same for 1-F, added to the ldj times reg_factor (:108-123)
        const T il10 = (T)0.43429448190325182;
        if (cF && lF * il10 < -(T)reg_max) LlF += gl * (T)reg_factor * il10;
        if (cG && lG * il10 < -(T)reg_max) LlG += gl * (T)reg_factor * il10;
    }
    const T LF = LlF * M::rc(F), LG = LlG * M::rc(G), Lf = gl * M::rc(f);
    const T Ssum = LF * F + LG * G + Lf * f;             // sum_j pi_j P_j
    T gx = 0;
    auto pass2 = [&](int k, float ms_k, float mu_k, float lp_k, float& o_ms, float& o_mu, float& o_lp) {
        const float mf = pre ? 1.0f : mfac[k];
        const float mf_max = fmaxf(mf, 1.0f);
        const T raw = (T)ms_k;
        const T imf_max = imf ? (T)imf[k] : M::rc((T)mf_max);
        const T thk = pre ? (T)0 : M::th(raw * imf_max);
        const T ls = pre ? raw : thk * (T)mf;
        const T e = M::ex(-ls);
        const T u = (x - (T)mu_k) * e;
        const T ea = M::ex(-fabs(u));
        const T r = M::rc((T)1 + ea);
        const T q = ea * r;
        const T sg = u >= 0 ? r : q, tg = u >= 0 ? q : r;
        const T d = q * r;
        const T pi = M::ex((T)lp_k - (T)m) * iw;
        const T Pk = LF * sg + LG * tg + Lf * d * e;
        const T Lsig = pi * ((LF - LG) + Lf * (tg - sg) * e);
        const T Lu = Lsig * d;
        const T Lue = Lu * e;
        gx += Lue;
        const T Lls = -u * Lu - Lf * pi * d * e;
        o_lp = (float)(pi * (Pk - Ssum));
        o_mu = (float)(-Lue);
        if (pre) {
            o_ms = (float)Lls;
        } else {
            const T sech2 = (T)1 - thk * thk;
            o_ms = (float)(Lls * sech2 * (T)mf * imf_max);
            // ls = tanh(raw / max(M,1)) M, M = e^{msf}: d ls / d msf  (for M > 1: max(M,1) = M)
            const T dmsf = (T)mf * thk - (mf >= 1.0f ? sech2 * raw : (T)0);
            if (KT > 0) gmsf_k[k] = (float)(Lls * dmsf);            // registers, reduced across the warp by the caller
            else atomicAdd(gmsf_k + k, (float)(Lls * dmsf));       // generic K / float64 path: shared-memory accumulator
        }
    };
    if constexpr (kPairs) {
#pragma unroll
        for (int k = 0; k < KT; k += 2) {
            const float2 a = *reinterpret_cast<const float2*>(ms + k), b = *reinterpret_cast<const float2*>(mu + k),
                         c = *reinterpret_cast<const float2*>(lp + k);
            float2 oa, ob, oc;
            pass2(k, a.x, b.x, c.x, oa.x, ob.x, oc.x);
            pass2(k + 1, a.y, b.y, c.y, oa.y, ob.y, oc.y);
            *reinterpret_cast<float2*>(ms + k) = oa;
            *reinterpret_cast<float2*>(mu + k) = ob;
            *reinterpret_cast<float2*>(lp + k) = oc;
        }
    } else {
#pragma unroll(KT > 0 ? KT : 4)
        for (int k = 0; k < K; ++k) pass2(k, ms[k], mu[k], lp[k], ms[k], mu[k], lp[k]);
    }
    rec[0] = (float)g_t;
    if (pre) {
        rec[1] = (float)g_logs;
    } else {
        const T sech2 = (T)1 - ths * ths;
        const T raw = (T)rec[1];
        *gsf_out += (float)(g_logs * ((T)sfac * ths - (sfac >= 1.0f ? sech2 * raw : (T)0)));
        rec[1] = (float)(g_logs * sech2 * (T)sfac * M::rc((T)sf_max));
    }
    *gx_out = (float)gx;
    return true;
}

template <int KT>
__global__ void __launch_bounds__(kThreads) mixcdf_bwd_kernel(const BwdParams p) {
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x;
    const int K = KT > 0 ? KT : p.K;
    const int PN = 2 + 3 * K;
    const int C = p.C, Ct = p.mask.n_t, TP = p.TP;
    const int L = Ct * PN;

    float* s_par = smem;                                   // [TP * L] parameters in, gradients out
    float* s_z = s_par + ((TP * L + 3) & ~3);              // [TP * C] z in
    float* s_g = s_z + ((TP * C + 3) & ~3);                // [TP * C] dL/dz_out in, dL/dz out
    float* s_fac = s_g + ((TP * C + 3) & ~3);              // [Ct] e^{sf}
    float* s_mfac = s_fac + Ct;                            // [Ct * K] e^{msf}
    float* s_gsf = s_mfac + Ct * K;                        // [Ct]
    float* s_gmsf = s_gsf + Ct;                            // [Ct * K]
    float* s_imf = s_gmsf + Ct * K;                        // [Ct * K] 1 / max(e^{msf}, 1)
    int* s_jmap = reinterpret_cast<int*>(s_imf + Ct * K);  // [C] channel -> transformed index or -1

    const long long pos0 = (long long)blockIdx.x * TP;
    const int rows = (int)min((long long)TP, p.P - pos0);

    if (p.vec_params) {
        const int L4 = L >> 2, total = rows * L4;
        const float inv = 1.0f / (float)L4;
        const float4* src = reinterpret_cast<const float4*>(p.nn);
        const long long row4 = p.compact ? (long long)L4 : ((long long)C * PN) >> 2;
        const int off4 = p.compact ? 0 : (p.mask.c0 * PN) >> 2;
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
        const int n = rows * C;
        for (int i = tid; i < n; i += kThreads) {
            cp_async4(s_z + i, p.z + pos0 * C + i);
            cp_async4(s_g + i, p.gz_out + pos0 * C + i);
        }
    }
    for (int i = tid; i < Ct; i += kThreads) {
        s_fac[i] = p.sf ? expf(p.sf[p.mask.tch[i]]) : 1.0f;
        s_gsf[i] = 0.f;
    }
    for (int i = tid; i < Ct * K; i += kThreads) {
        const int j = i / K, k = i - j * K;
        s_mfac[i] = p.msf ? expf(p.msf[p.mask.tch[j] * K + k]) : 1.0f;
        s_imf[i] = 1.0f / fmaxf(s_mfac[i], 1.0f);
        s_gmsf[i] = 0.f;
    }
    for (int i = tid; i < C; i += kThreads) s_jmap[i] = -1;
    cp_async_wait_all();
    __syncthreads();
    for (int i = tid; i < Ct; i += kThreads) s_jmap[p.mask.tch[i]] = i;

    // ---- one thread per (position, transformed channel) ----------------------------------------
    const int nelem = rows * Ct;
    const float inv_ct = 1.0f / (float)Ct;
    // d/d mixture_scaling_factor: per-thread registers, summed over the lanes that share a channel (lane % Ct when Ct
    // divides 32) before touching the shared accumulator - 32/Ct times fewer shared atomics
    const bool warp_reduce = KT > 0 && (32 % Ct) == 0 && (kThreads % Ct) == 0;
    for (int e0 = 0; e0 < nelem; e0 += kThreads) {
        const int e = e0 + tid;
        float gm[KT > 0 ? KT : 1];
#pragma unroll
        for (int k = 0; k < (KT > 0 ? KT : 1); ++k) gm[k] = 0.f;
        float gsf = 0.f;
        int j = 0;
        if (e < nelem) {
            const int r = fast_div(e, inv_ct);
            j = e - r * Ct;
            const long long pos = pos0 + r;
            const int ch = p.mask.tch[j];
            float* rec = s_par + (size_t)e * PN;
            const float padv = p.pad ? p.pad[pos] : 1.0f;
            bool active = padv != 0.0f;
            if (p.mask.s_period > 0) {
                const int s = (int)(pos % p.S);
                if ((p.mask.cond_s >> (s % p.mask.s_period)) & 1ull) active = false;
            }
            const float gzo = s_g[r * C + ch];
            if (!active) {   // copied through (times pad): no parameter gradient
                for (int i = 0; i < PN; ++i) rec[i] = 0.f;
                s_g[r * C + ch] = gzo * padv;
            } else {
                // z_final = (out * pad + x (1 - pad)) * pad, ldj_e * pad (mixture_cdf_layer.py:76,99-101,137-138)
                const float g_out = gzo * padv * padv;
                const float gl = (p.gldj ? p.gldj[pos / p.S] : 0.f) * padv;
                const float x = s_z[r * C + ch];
                float gx = 0.f;
                const bool ok = mix_backward_elem<float, KT>(x, rec, s_mfac + j * K, s_imf + j * K, s_fac[j], K, p.pre != 0, g_out, gl, p.use_reg != 0,
                                                              p.reg_max, p.reg_factor, &gx, &gsf, KT > 0 ? gm : s_gmsf + j * K);
                if (!ok) {
#pragma unroll
                    for (int k = 0; k < (KT > 0 ? KT : 1); ++k) gm[k] = 0.f;
                    mix_backward_elem<double, 0>(x, rec, s_mfac + j * K, nullptr, s_fac[j], K, p.pre != 0, g_out, gl, p.use_reg != 0, p.reg_max,
                                                 p.reg_factor, &gx, &gsf, s_gmsf + j * K);
                }
                s_g[r * C + ch] = gx + gzo * (1.0f - padv) * padv;
            }
        }
        if (p.pre == 0 && (p.gsf != nullptr || p.gmsf != nullptr)) {
            if (warp_reduce) {
                for (int d = 16; d >= Ct; d >>= 1) {
                    gsf += __shfl_xor_sync(0xffffffffu, gsf, d);
#pragma unroll
                    for (int k = 0; k < (KT > 0 ? KT : 1); ++k) gm[k] += __shfl_xor_sync(0xffffffffu, gm[k], d);
                }
                const int jj = (tid & 31) % Ct;      // == j for every lane that holds an element (kThreads % Ct == 0)
                if ((tid & 31) < Ct) {
                    if (gsf != 0.f) atomicAdd(s_gsf + jj, gsf);
#pragma unroll
                    for (int k = 0; k < (KT > 0 ? KT : 1); ++k)
                        if (gm[k] != 0.f) atomicAdd(s_gmsf + jj * K + k, gm[k]);
                }
            } else if (e < nelem) {
                if (gsf != 0.f) atomicAdd(s_gsf + j, gsf);
                if (KT > 0) {
#pragma unroll
                    for (int k = 0; k < (KT > 0 ? KT : 1); ++k)
                        if (gm[k] != 0.f) atomicAdd(s_gmsf + j * K + k, gm[k]);
                }
            }
        }
    }
    __syncthreads();
    // conditioner channels: dL/dz = dL/dz_out * pad
    if (Ct < C) {
        const int n = rows * C;
        const float inv_c = 1.0f / (float)C;
        for (int i = tid; i < n; i += kThreads) {
            const int r = fast_div(i, inv_c), c = i - r * C;
            if (s_jmap[c] < 0 && p.pad) s_g[i] *= p.pad[pos0 + r];
        }
        __syncthreads();
    }

    // ---- dL/dnn_out rows: gradient records of the transformed channels, zeros elsewhere ----------
    if (p.vec_params && p.vec_out) {
        // contiguous, 16-byte aligned run of transformed records per position: copy it with 16-byte stores and zero the
        // conditioner records before / after it the same way (no per-granule record lookup)
        const int L4 = L >> 2, row4 = p.compact ? L4 : (C * PN) >> 2, off4 = p.compact ? 0 : (p.mask.c0 * PN) >> 2;
        const int Z4 = row4 - L4;                               // zero float4 per position (none in the compact layout)
        const float inv_l4 = 1.0f / (float)L4;
        float4* dst = reinterpret_cast<float4*>(p.gnn) + pos0 * (long long)row4;
        const float4* src = reinterpret_cast<const float4*>(s_par);
        for (int i = tid; i < rows * L4; i += kThreads) {
            const int r = fast_div(i, inv_l4), q = i - r * L4;
            stg_stream4(dst + (size_t)r * row4 + off4 + q, src[i]);
        }
        if (Z4 > 0) {
            const float inv_z4 = 1.0f / (float)Z4;
            const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int i = tid; i < rows * Z4; i += kThreads) {
                const int r = fast_div(i, inv_z4), q = i - r * Z4;
                stg_stream4(dst + (size_t)r * row4 + (q < off4 ? q : q + L4), zero);
            }
        }
    } else if ((PN & 1) == 0) {   // even K: 8-byte granules never straddle a record
        const int PN2 = PN >> 1;
        const int row2 = C * PN2, total = rows * row2;
        const float inv_row = 1.0f / (float)row2, inv_pn2 = 1.0f / (float)PN2;
        float2* dst = reinterpret_cast<float2*>(p.gnn + pos0 * (long long)C * PN);
        for (int i = tid; i < total; i += kThreads) {
            const int r = fast_div(i, inv_row), q = i - r * row2;
            const int c = fast_div(q, inv_pn2), i2 = q - c * PN2;
            const int j = s_jmap[c];
            float2 v = make_float2(0.f, 0.f);
            if (j >= 0) v = *reinterpret_cast<const float2*>(s_par + ((size_t)r * Ct + j) * PN + 2 * i2);
            dst[i] = v;
        }
    } else {
        const int rowlen = C * PN, total = rows * rowlen;
        const float inv_row = 1.0f / (float)rowlen, inv_pn = 1.0f / (float)PN;
        float* dst = p.gnn + pos0 * (long long)C * PN;
        for (int i = tid; i < total; i += kThreads) {
            const int r = fast_div(i, inv_row), q = i - r * rowlen;
            const int c = fast_div(q, inv_pn), i1 = q - c * PN;
            const int j = s_jmap[c];
            dst[i] = j >= 0 ? s_par[((size_t)r * Ct + j) * PN + i1] : 0.f;
        }
    }
    // ---- dL/dz rows ----------------------------------------------------------------------------
    {
        const int n = rows * C;
        float* dst = p.gz + pos0 * C;
        for (int i = tid; i < n; i += kThreads) dst[i] = s_g[i];
    }
    // ---- scaling-factor gradients: one atomic per (CTA, parameter) -----------------------------
    if (p.gsf)
        for (int i = tid; i < Ct; i += kThreads)
            if (s_gsf[i] != 0.f) atomicAdd(p.gsf + p.mask.tch[i], s_gsf[i]);
    if (p.gmsf)
        for (int i = tid; i < Ct * K; i += kThreads)
            if (s_gmsf[i] != 0.f) atomicAdd(p.gmsf + p.mask.tch[i / K] * K + i % K, s_gmsf[i]);
}

// Compact layout (nn / dL/dnn_out hold the transformed channels' records only, rows 16-byte aligned): PERSISTENT CTAs over
// the tiles, data movement by the TMA engine.  A tile's parameter records, z rows and incoming gradient rows are three
// contiguous runs of global memory: one cp.async.bulk each into one of `nbuf` shared-memory buffers, completion on the
// buffer's mbarrier; the gradient records (written in place) and the dL/dz rows leave as two bulk stores.  One thread
// issues everything: after its own element of tile k it waits until the store of tile k-1 has been read out of shared
// memory (long done) and loads tile k + nbuf - 1 into that buffer, so up to nbuf - 1 loads are in flight under the math.
// The per-CTA tables (e^{sf}, e^{msf}, 1 / max(e^{msf}, 1)) are built once and the scaling-factor gradients leave with one
// atomic per (CTA, parameter) at the very end - the tile kernel above pays both per tile of 32 positions (32768 CTAs x 72
// global atomics at the LM shape).
#ifndef CNF_BWD_PROJ_MIN_CTAS
#define CNF_BWD_PROJ_MIN_CTAS 3      // 80 registers: 3 CTAs per SM (15.5 ms per LM training step; 1 CTA at 150 registers: 21.1 ms)
#endif
template <int KT, bool PROJ>
__global__ void __launch_bounds__(kThreads, PROJ ? CNF_BWD_PROJ_MIN_CTAS : 1) mixcdf_bwd_pipe_kernel(const BwdParams p) {
    extern __shared__ __align__(16) float smem[];
    __shared__ __align__(8) uint64_t s_full[4];
    const int tid = threadIdx.x;
    const int K = KT > 0 ? KT : p.K;
    const int PN = 2 + 3 * K;
    const int C = p.C, Ct = p.mask.n_t, TP = p.TP, nbuf = p.nbuf;
    const int L = Ct * PN;
    const int buf_floats = TP * L + 2 * TP * C;            // [TP * L | TP * C | TP * C], all multiples of 4 floats
    float* s_fac = smem + (size_t)nbuf * buf_floats;       // [Ct] e^{sf}
    float* s_mfac = s_fac + Ct;                            // [Ct * K] e^{msf}
    float* s_gsf = s_mfac + Ct * K;                        // [Ct]
    float* s_gmsf = s_gsf + Ct;                            // [Ct * K]
    float* s_imf = s_gmsf + Ct * K;                        // [Ct * K] 1 / max(e^{msf}, 1)
    int* s_jmap = reinterpret_cast<int*>(s_imf + Ct * K);  // [C] channel -> transformed index or -1
    float* s_col = reinterpret_cast<float*>(s_jmap + C);   // [L] column sums of the gradient records (p.gcol)
    float* s_pw = s_col + ((L + 3) & ~3);                  // [L][8] conditioner columns of the projection weight (p.proj_w)
    const int cb = p.mask.c0 == 0 ? Ct : 0;                // first conditioner channel (C = 16, Ct = 8 in this mode)

    for (int i = tid; i < Ct; i += kThreads) {
        s_fac[i] = p.sf ? expf(p.sf[p.mask.tch[i]]) : 1.0f;
        s_gsf[i] = 0.f;
    }
    for (int i = tid; i < Ct * K; i += kThreads) {
        const int j = i / K, k = i - j * K;
        s_mfac[i] = p.msf ? expf(p.msf[p.mask.tch[j] * K + k]) : 1.0f;
        s_imf[i] = 1.0f / fmaxf(s_mfac[i], 1.0f);
        s_gmsf[i] = 0.f;
    }
    for (int i = tid; i < C; i += kThreads) s_jmap[i] = -1;
    if (p.gcol)
        for (int i = tid; i < L; i += kThreads) s_col[i] = 0.f;
    // two planes [L][4] (conditioner columns 0..3 | 4..7): the eight n-slices of a position read eight consecutive rows of a
    // plane per step = 32 different banks
    if constexpr (PROJ)
        for (int i = tid; i < L * 8; i += kThreads) {
            const int n = i >> 3, j = i & 7;
            s_pw[(j >> 2) * (L * 4) + n * 4 + (j & 3)] = p.proj_w[n * C + cb + j];
        }
    // dL/dW entries of this thread: row n = tid < L, the 8 conditioner columns
    float accw[PROJ ? 2 : 1][4] = {};
    if (tid == 0) {
        for (int b = 0; b < nbuf; ++b) tc::mbar_init(&s_full[b], 1);
        tc::mbar_fence_init();
    }
    __syncthreads();
    for (int i = tid; i < Ct; i += kThreads) s_jmap[p.mask.tch[i]] = i;
    __syncthreads();

    const long long ntiles = (p.P + TP - 1) / TP;
    const int n_my = (int)((ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x);      // tiles blockIdx.x, + gridDim.x, ...
    auto issue_load = [&](int k) {      // thread 0 only
        const long long pos0 = ((long long)blockIdx.x + (long long)k * gridDim.x) * TP;
        const int rows = (int)min((long long)TP, p.P - pos0);
        float* bpar = smem + (size_t)(k % nbuf) * buf_floats;
        const uint32_t b_par = (uint32_t)rows * (uint32_t)L * 4u, b_row = (uint32_t)rows * (uint32_t)C * 4u;
        uint64_t* bar = &s_full[k % nbuf];
        tc::mbar_arrive_expect_tx(bar, b_par + 2u * b_row);
        tc::bulk_load(bpar, p.nn + pos0 * (long long)L, b_par, bar);
        tc::bulk_load(bpar + TP * L, p.z + pos0 * C, b_row, bar);
        tc::bulk_load(bpar + TP * L + TP * C, p.gz_out + pos0 * C, b_row, bar);
    };
    const int ahead = nbuf > 1 ? nbuf - 1 : 1;
    if (tid == 0)
        for (int k = 0; k < ahead && k < n_my; ++k) issue_load(k);

    const bool warp_reduce = KT > 0 && (32 % Ct) == 0 && (kThreads % Ct) == 0;
    const float inv_ct = 1.0f / (float)Ct;
    for (int k = 0; k < n_my; ++k) {
        const long long pos0 = ((long long)blockIdx.x + (long long)k * gridDim.x) * TP;
        const int rows = (int)min((long long)TP, p.P - pos0);
        float* s_par = smem + (size_t)(k % nbuf) * buf_floats;
        float* s_z = s_par + TP * L;
        float* s_g = s_z + TP * C;
        tc::mbar_wait(&s_full[k % nbuf], (uint32_t)((k / nbuf) & 1));

        // ---- one thread per (position, transformed channel): TP * Ct <= kThreads elements per tile ----------------------
        const int nelem = rows * Ct;
        const int e = tid;
        float gm[KT > 0 ? KT : 1];
#pragma unroll
        for (int kk = 0; kk < (KT > 0 ? KT : 1); ++kk) gm[kk] = 0.f;
        float gsf = 0.f;
        int j = 0;
        if (e < nelem) {
            const int r = fast_div(e, inv_ct);
            j = e - r * Ct;
            const long long pos = pos0 + r;
            const int ch = p.mask.tch[j];
            float* rec = s_par + (size_t)e * PN;
            const float padv = p.pad ? p.pad[pos] : 1.0f;
            bool active = padv != 0.0f;
            if (p.mask.s_period > 0) {
                const int sp = (int)(pos % p.S);
                if ((p.mask.cond_s >> (sp % p.mask.s_period)) & 1ull) active = false;
            }
            const float gzo = s_g[r * C + ch];
            if (!active) {   // copied through (times pad): no parameter gradient
                for (int i = 0; i < PN; ++i) rec[i] = 0.f;
                s_g[r * C + ch] = gzo * padv;
            } else {
                const float g_out = gzo * padv * padv;
                const float gl = (p.gldj ? p.gldj[pos / p.S] : 0.f) * padv;
                const float x = s_z[r * C + ch];
                float gx = 0.f;
                const bool ok = mix_backward_elem<float, KT>(x, rec, s_mfac + j * K, s_imf + j * K, s_fac[j], K, p.pre != 0, g_out, gl,
                                                              p.use_reg != 0, p.reg_max, p.reg_factor, &gx, &gsf, KT > 0 ? gm : s_gmsf + j * K);
                if (!ok) {
#pragma unroll
                    for (int kk = 0; kk < (KT > 0 ? KT : 1); ++kk) gm[kk] = 0.f;
                    mix_backward_elem<double, 0>(x, rec, s_mfac + j * K, nullptr, s_fac[j], K, p.pre != 0, g_out, gl, p.use_reg != 0,
                                                 p.reg_max, p.reg_factor, &gx, &gsf, s_gmsf + j * K);
                }
                s_g[r * C + ch] = gx + gzo * (1.0f - padv) * padv;
            }
        }
        if (p.pre == 0 && (p.gsf != nullptr || p.gmsf != nullptr)) {
            if (warp_reduce) {
                for (int d = 16; d >= Ct; d >>= 1) {
                    gsf += __shfl_xor_sync(0xffffffffu, gsf, d);
#pragma unroll
                    for (int kk = 0; kk < (KT > 0 ? KT : 1); ++kk) gm[kk] += __shfl_xor_sync(0xffffffffu, gm[kk], d);
                }
                const int jj = (tid & 31) % Ct;
                if ((tid & 31) < Ct) {
                    if (gsf != 0.f) atomicAdd(s_gsf + jj, gsf);
#pragma unroll
                    for (int kk = 0; kk < (KT > 0 ? KT : 1); ++kk)
                        if (gm[kk] != 0.f) atomicAdd(s_gmsf + jj * K + kk, gm[kk]);
                }
            } else if (e < nelem) {
                if (gsf != 0.f) atomicAdd(s_gsf + j, gsf);
                if (KT > 0) {
#pragma unroll
                    for (int kk = 0; kk < (KT > 0 ? KT : 1); ++kk)
                        if (gm[kk] != 0.f) atomicAdd(s_gmsf + j * K + kk, gm[kk]);
                }
            }
        }
        // conditioner channels: dL/dz = dL/dz_out * pad (entries no element thread touches)
        if (Ct < C && p.pad) {
            const int n = rows * C;
            const float inv_c = 1.0f / (float)C;
            for (int i = tid; i < n; i += kThreads) {
                const int r = fast_div(i, inv_c), c = i - r * C;
                if (s_jmap[c] < 0) s_g[i] *= p.pad[pos0 + r];
            }
        }
        // ---- this tile's stores, next load, column sums ---------------------------------------------------------------
        tc::fence_proxy_async_smem();
        __syncthreads();
        if constexpr (PROJ) {
            // The network IS a per-position Linear on z (training path of MixtureCDFCoupling with the mask folded into the
            // weight): its backward products consume the gradient records right here, from shared memory, in fp32 -
            //   dL/dz[pos, cond] += sum_n G[pos, n] W[n, cond]        thread (pos = tid / 8, slice s = tid % 8 of the n range),
            //                                                          8 conditioner columns per thread, reduced over the 8 slices
            //   dL/dW[n, cond]   += sum_pos G[pos, n] z[pos, cond]     lane-owned entries, accumulated in registers over all tiles
            // - so dL/dnn_out never has to leave the chip (the two GEMMs re-read 0.87 GB each per block at the LM shape).
            {
                const int pp = tid >> 3, sl = tid & 7;
                float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                if (pp < rows) {
                    const float* grow = s_par + (size_t)pp * L;
                    const float* w1p = s_pw + L * 4;
#pragma unroll 2
                    for (int n = sl; n < L; n += 8) {      // slice sl: rows sl, sl + 8, ...
                        const float g = grow[n];
                        const float4 w0 = *reinterpret_cast<const float4*>(s_pw + n * 4), w1 = *reinterpret_cast<const float4*>(w1p + n * 4);
                        a[0] = fmaf(g, w0.x, a[0]); a[1] = fmaf(g, w0.y, a[1]); a[2] = fmaf(g, w0.z, a[2]); a[3] = fmaf(g, w0.w, a[3]);
                        a[4] = fmaf(g, w1.x, a[4]); a[5] = fmaf(g, w1.y, a[5]); a[6] = fmaf(g, w1.z, a[6]); a[7] = fmaf(g, w1.w, a[7]);
                    }
                }
#pragma unroll
                for (int d = 1; d < 8; d <<= 1) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) a[j] += __shfl_xor_sync(0xffffffffu, a[j], d);
                }
                if (sl == 0 && pp < rows) {
                    float4* dst = reinterpret_cast<float4*>(s_g + pp * C + cb);
                    float4 v0 = dst[0], v1 = dst[1];
                    v0.x += a[0]; v0.y += a[1]; v0.z += a[2]; v0.w += a[3];
                    v1.x += a[4]; v1.y += a[5]; v1.z += a[6]; v1.w += a[7];
                    dst[0] = v0; dst[1] = v1;
                }
            }
            const bool col_here = p.gcol != nullptr && p.gproj_w != nullptr;      // column sums ride on the dL/dW loop's reads
            if (p.gproj_w && tid < L) {      // thread n < L owns row n of dL/dW (8 conditioner columns) over all of its tiles
                const float* zc = s_z + cb;
                float csum = 0.f;
#pragma unroll 4
                for (int r = 0; r < rows; ++r) {
                    const float g = s_par[r * L + tid];
                    csum += g;
                    const float4 z0 = *reinterpret_cast<const float4*>(zc + r * C), z1 = *reinterpret_cast<const float4*>(zc + r * C + 4);
                    accw[0][0] = fmaf(g, z0.x, accw[0][0]); accw[0][1] = fmaf(g, z0.y, accw[0][1]);
                    accw[0][2] = fmaf(g, z0.z, accw[0][2]); accw[0][3] = fmaf(g, z0.w, accw[0][3]);
                    accw[1][0] = fmaf(g, z1.x, accw[1][0]); accw[1][1] = fmaf(g, z1.y, accw[1][1]);
                    accw[1][2] = fmaf(g, z1.z, accw[1][2]); accw[1][3] = fmaf(g, z1.w, accw[1][3]);
                }
                if (col_here) s_col[tid] += csum;
            }
            if (p.gcol && !col_here) {
                for (int c = tid; c < L; c += kThreads) {
                    float a = 0.f;
                    for (int r = 0; r < rows; ++r) a += s_par[r * L + c];
                    s_col[c] += a;
                }
            }
            tc::fence_proxy_async_smem();
            __syncthreads();
        }
        if (tid == 0) {
            if (p.gnn) bulk_store_rows(p.gnn + pos0 * (long long)L, s_par, (uint32_t)rows * (uint32_t)L * 4u);
            bulk_store_rows(p.gz + pos0 * C, s_g, (uint32_t)rows * (uint32_t)C * 4u);
            tc::tma_store_commit();
            // the load of tile k + ahead goes into the buffer of tile k - nbuf + ahead: for nbuf > 1 an OLDER tile's, whose
            // store has been read out (all but the group just committed are complete) and whose column sums every thread
            // finished before the barrier above; a single buffer is the one being read right now
            if (k + ahead < n_my) {
                if (nbuf > 1) tc::tma_store_wait_read<1>(); else tc::tma_store_wait_read<0>();
                issue_load(k + ahead);
            }
        }
        if (p.gcol && !PROJ) {
            // d/dbias of the network's final Linear = column sums of dL/dnn_out: column tid of the tile's rows, read while
            // the TMA engine reads the same rows for the store (needs nbuf > 1: a single buffer is re-filled right away)
            for (int c = tid; c < L; c += kThreads) {
                float a = 0.f;
                for (int r = 0; r < rows; ++r) a += s_par[r * L + c];
                s_col[c] += a;
            }
        }
    }
    __syncthreads();
    if (p.gsf)
        for (int i = tid; i < Ct; i += kThreads)
            if (s_gsf[i] != 0.f) atomicAdd(p.gsf + p.mask.tch[i], s_gsf[i]);
    if (p.gmsf)
        for (int i = tid; i < Ct * K; i += kThreads)
            if (s_gmsf[i] != 0.f) atomicAdd(p.gmsf + p.mask.tch[i / K] * K + i % K, s_gmsf[i]);
    if (p.gcol)
        for (int i = tid; i < L; i += kThreads)
            if (s_col[i] != 0.f) atomicAdd(p.gcol + i, s_col[i]);
    if (PROJ && p.gproj_w && tid < L) {
        float* dst = p.gproj_w + (size_t)tid * C + cb;
#pragma unroll
        for (int it = 0; it < (PROJ ? 2 : 1); ++it)
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (accw[it][j] != 0.f) atomicAdd(dst + 4 * it + j, accw[it][j]);
    }
    if (tid == 0) tc::tma_store_wait_read<0>();      // shared memory must outlive the engine's reads of it
}

template <int KT, bool PROJ>
int launch_bwd_pipe_t(BwdParams p, int L, int C, int Ct, int K, cudaStream_t stream) {
    static const int want = getenv("CNF_B200_MIXCDF_BWD_NBUF") ? atoi(getenv("CNF_B200_MIXCDF_BWD_NBUF")) : 2;      // tuning knob
    int nbuf = want < 1 ? 1 : (want > 4 ? 4 : want);
    if (p.gcol != nullptr && nbuf < 2) nbuf = 2;      // the fused column sums read a tile after its store was issued
    const size_t buf = ((size_t)p.TP * L + 2 * (size_t)p.TP * C) * sizeof(float);
    const size_t tables = (2 * (size_t)Ct + 3 * (size_t)Ct * K + (size_t)C + (size_t)((L + 3) & ~3) + (p.proj_w ? 8 * (size_t)L : 0)) * sizeof(float);
    while (nbuf > 2 && nbuf * buf + tables > 200 * 1024) --nbuf;
    if (nbuf * buf + tables > 200 * 1024) return -1;      // caller falls back to the tile kernel
    p.nbuf = nbuf;
    const size_t smem = nbuf * buf + tables;
    CNF_CUDA(cudaFuncSetAttribute(mixcdf_bwd_pipe_kernel<KT, PROJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    CNF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mixcdf_bwd_pipe_kernel<KT, PROJ>, kThreads, smem));
    if (per_sm < 1) per_sm = 1;
    const long long ntiles = (p.P + p.TP - 1) / p.TP;
    long long grid = (long long)per_sm * sm_count();
    if (grid > ntiles) grid = ntiles;
    mixcdf_bwd_pipe_kernel<KT, PROJ><<<(unsigned)grid, kThreads, smem, stream>>>(p);
    return launch_status("mixcdf_bwd_pipe_kernel");
}

template <int KT>
int launch_bwd_pipe(const BwdParams& p, int L, int C, int Ct, int K, cudaStream_t stream) {
    if constexpr (KT > 0) {
        if (p.proj_w != nullptr) return launch_bwd_pipe_t<KT, true>(p, L, C, Ct, K, stream);
    }
    if (p.proj_w != nullptr) return fail(CNF_ERR_UNSUPPORTED, "proj_weight: K must be 4, 8 or 16");
    return launch_bwd_pipe_t<KT, false>(p, L, C, Ct, K, stream);
}

size_t bwd_smem(int TP, int L, int C, int Ct, int K) {
    size_t f = ((size_t)TP * L + 3) & ~(size_t)3;
    f += 2 * (((size_t)TP * C + 3) & ~(size_t)3);
    f += 2 * (size_t)Ct + 3 * (size_t)Ct * K + (size_t)C;
    return f * sizeof(float);
}

template <int KT>
int launch_bwd(const BwdParams& p, size_t smem, cudaStream_t stream) {
    if (smem > 48 * 1024)
        CNF_CUDA(cudaFuncSetAttribute(mixcdf_bwd_kernel<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long grid = (p.P + p.TP - 1) / p.TP;
    mixcdf_bwd_kernel<KT><<<(unsigned)grid, kThreads, smem, stream>>>(p);
    return launch_status("mixcdf_bwd_kernel");
}

}  // namespace
}  // namespace cnf

extern "C" int cnf_mixcdf_bwd(const cnf_mixcdf_bwd_args* a, cnf_stream_t stream_) {
    using namespace cnf;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    CNF_REQUIRE(a != nullptr, "args is NULL");
    CNF_REQUIRE(a->B >= 0 && a->S >= 0, "negative batch/sequence size");
    CNF_REQUIRE(a->C >= 1 && a->K >= 1, "C and K must be >= 1 (got C=%d K=%d)", a->C, a->K);
    CNF_SUPPORTED(a->C <= CNF_MAX_CHANNELS && a->K <= CNF_MAX_MIXTURES, "C=%d / K=%d outside the compiled range", a->C, a->K);
    BwdParams p{};
    int rc = build_mask(a->mask, a->C, &p.mask);
    if (rc != CNF_OK) return rc;
    const long long P = a->B * a->S;
    if (P == 0) return CNF_OK;
    CNF_REQUIRE(a->z && a->nn_out && a->grad_z_out && a->grad_z, "z / nn_out / grad_z_out / grad_z is NULL");
    const size_t nz = (size_t)P * a->C;
    p.PN = 2 + 3 * a->K;
    if (p.mask.n_t == 0) {   // nothing transformed: identity (times pad), no parameter gradient
        CNF_SUPPORTED(a->pad == nullptr, "mask with no transformed channel together with a padding mask");
        CNF_CUDA(cudaMemcpyAsync(a->grad_z, a->grad_z_out, nz * sizeof(float), cudaMemcpyDeviceToDevice, stream));
        if (!a->nn_compact) CNF_CUDA(cudaMemsetAsync(a->grad_nn_out, 0, nz * p.PN * sizeof(float), stream));
        return CNF_OK;
    }
    p.z = a->z; p.nn = a->nn_out; p.pad = a->pad; p.sf = a->scaling_factor; p.msf = a->mixture_scaling_factor;
    p.gz_out = a->grad_z_out; p.gldj = a->grad_ldj; p.gz = a->grad_z; p.gnn = a->grad_nn_out;
    p.gsf = a->params_prebounded ? nullptr : a->grad_scaling_factor;
    p.gmsf = a->params_prebounded ? nullptr : a->grad_mixture_scaling_factor;
    p.P = P; p.S = (int)a->S; p.C = a->C; p.K = a->K;
    p.reg_max = a->reg_max; p.reg_factor = a->reg_factor;
    p.use_reg = (a->reg_max > 0.f && a->training) ? 1 : 0;
    p.pre = a->params_prebounded;
    const int Ct = p.mask.n_t, L = Ct * p.PN;
    int elems = kThreads;
    const int cap = (40 * 1024) / (4 * p.PN);
    if (elems > cap) elems = cap;
    int TP = elems / Ct;
    TP &= ~3;
    if (TP < 4) TP = 4;
    p.TP = TP;
    const long long rowlen = (long long)a->C * p.PN;
    p.vec_params = p.mask.contiguous && (L % 4 == 0) && (rowlen % 4 == 0) && ((p.mask.c0 * p.PN) % 4 == 0) &&
                   ((reinterpret_cast<uintptr_t>(a->nn_out) & 15) == 0);
    p.vec_out = (reinterpret_cast<uintptr_t>(a->grad_nn_out) & 15) == 0 ? 1 : 0;
    p.compact = a->nn_compact ? 1 : 0;
    if (p.compact) {
        p.vec_params = p.mask.contiguous && (L % 4 == 0) && ((reinterpret_cast<uintptr_t>(a->nn_out) & 15) == 0);
        CNF_SUPPORTED(p.vec_params && p.vec_out, "the compact gradient layout needs a contiguous run of transformed channels with "
                                                   "Ct * (2 + 3K) a multiple of 4 and 16-byte aligned nn_out / grad_nn_out");
    }
    static const bool no_bulk = getenv("CNF_B200_MIXCDF_BWD_NO_BULK") != nullptr;      // A/B switch
    p.bulk = (p.compact && !no_bulk && (a->C % 4 == 0) &&
              ((reinterpret_cast<uintptr_t>(a->z) | reinterpret_cast<uintptr_t>(a->grad_z_out) | reinterpret_cast<uintptr_t>(a->grad_z)) & 15) == 0)
                 ? 1 : 0;
    CNF_REQUIRE((reinterpret_cast<uintptr_t>(a->grad_nn_out) & 7) == 0, "grad_nn_out must be 8-byte aligned");
    CNF_SUPPORTED((long long)TP * a->C * p.PN < (1 << 21), "tile too large for the index arithmetic");
    CNF_SUPPORTED(a->grad_nn_colsum == nullptr || p.compact, "grad_nn_colsum needs the compact layout (nn_compact = 1)");
    const bool want_proj = a->proj_weight != nullptr;
    if (want_proj) {
        CNF_SUPPORTED(p.bulk && TP * Ct <= kThreads && (TP * L) % 4 == 0 && a->C == 16 && Ct == 8 && p.mask.contiguous &&
                          (p.mask.c0 == 0 || p.mask.c0 == 8) && p.mask.s_period == 0 && L <= kThreads && TP * 8 <= kThreads,
                      "proj_weight: the fused projection backward is compiled for the compact layout with C = 16 and 8 contiguous "
                      "transformed channels at either end (channel mask only)");
        CNF_REQUIRE((reinterpret_cast<uintptr_t>(a->proj_weight) & 15) == 0, "proj_weight must be 16-byte aligned");
        p.proj_w = a->proj_weight;
        p.gproj_w = a->grad_proj_weight;
    } else {
        CNF_REQUIRE(a->grad_proj_weight == nullptr, "grad_proj_weight needs proj_weight");
        CNF_REQUIRE(a->grad_nn_out != nullptr, "grad_nn_out is NULL");
    }
    if (p.bulk && TP * Ct <= kThreads && (TP * L) % 4 == 0) {
        p.gcol = a->grad_nn_colsum;
        int prc;
        switch (a->K) {
            case 4: prc = launch_bwd_pipe<4>(p, L, a->C, Ct, a->K, stream); break;
            case 8: prc = launch_bwd_pipe<8>(p, L, a->C, Ct, a->K, stream); break;
            case 16: prc = launch_bwd_pipe<16>(p, L, a->C, Ct, a->K, stream); break;
            default: prc = launch_bwd_pipe<0>(p, L, a->C, Ct, a->K, stream); break;
        }
        if (prc != -1) return prc;
        CNF_SUPPORTED(!want_proj, "proj_weight: tile does not fit shared memory");
        p.gcol = nullptr;
    }
    const size_t smem = bwd_smem(TP, L, a->C, Ct, a->K);
    CNF_SUPPORTED(smem <= 200 * 1024, "C=%d K=%d needs %zu bytes of shared memory per tile", a->C, a->K, smem);
    switch (a->K) {
        case 4: rc = launch_bwd<4>(p, smem, stream); break;
        case 8: rc = launch_bwd<8>(p, smem, stream); break;
        case 16: rc = launch_bwd<16>(p, smem, stream); break;
        default: rc = launch_bwd<0>(p, smem, stream); break;
    }
    // the tile kernel does not sum columns: a separate pass over the gradient it just wrote (same result)
    if (rc == CNF_OK && a->grad_nn_colsum != nullptr) rc = tc_colsum(a->grad_nn_out, a->grad_nn_colsum, P, L, stream);
    return rc;
}
