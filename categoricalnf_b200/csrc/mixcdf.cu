// K1 / K2: logistic-mixture-CDF coupling transform, forward and inverse, for sm_100a.
//
// Replaces MixtureCDFCoupling.get_mixt_params + run_with_params
// (reference layers/flows/mixture_cdf_layer.py:95-142, 145-180, 197-276): ~100 eager float64
// ATen launches per call there, one kernel here.
//
// Work decomposition: a CTA owns a tile of TP consecutive positions.  The parameter records of
// the *transformed* channels only (the conditioner half of nn_out is never read) and the z rows
// are staged in shared memory with cp.async (16-byte when the layout allows), then a group of
// G = 2^g adjacent lanes (G = 1 up to K = 8, 8 lanes at K = 64) evaluates one (position, channel)
// element: each lane prepares its <= 8 mixture components once (bounded scales, softmax numerators)
// and keeps them in registers, an evaluation of the mixture is 2 MUFU per component plus a
// butterfly over the group, the inverse is the safeguarded Newton iteration of the pipelined kernel.
// Elements whose CDF leaves the range where fp32 reproduces the reference's fp64 arithmetic to 1e-4
// take a float64 path that restates the reference formulas literally.
// Per-sample ldj: shared-memory per-position partials -> warp-segmented sum -> one global
// atomicAdd per (CTA, sample).
#include <stdlib.h>

#include "cnf_common.cuh"
#include "mixcdf_math.cuh"

namespace cnf {

int mixcdf_pipe_try(const cnf_mixcdf_args* a, const MaskView& mask, int reverse, cudaStream_t stream, int* handled);
int mixcdf_gpipe_try(const cnf_mixcdf_args* a, const MaskView& mask, int reverse, cudaStream_t stream, int* handled);
bool mixcdf_pipe_fusable(const cnf_mixcdf_args* a, const MaskView& mask, int reverse);
bool mixcdf_pipe_eligible(const cnf_mixcdf_args* a, const MaskView& mask);
bool mixcdf_gpipe_eligible(const cnf_mixcdf_args* a, const MaskView& mask);

namespace {
using namespace mixmath;

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
    int G;   // lanes per element (power of two, K <= NC * G)
    MaskView mask;
    int vec_params;  // parameter rows can be copied as 16-byte chunks
    float reg_max, reg_factor;
    int use_reg;
    int reverse;
    int pre;  // parameters already tanh-bounded
};

// ------------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------------
template <int NC, bool REV>
__global__ void __launch_bounds__(kThreads) mixcdf_kernel(const MixParams p) {
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x;
    const int K = p.K;
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

    // ---- one lane group per (position, transformed channel) -------------------------------------
    const int nelem = rows * Ct;
    const float inv_ct = 1.0f / (float)Ct;
    LaneGroup g;
    g.G = p.G;
    g.sub = tid & (p.G - 1);
    g.mask = p.G == 32 ? 0xffffffffu : (((1u << p.G) - 1u) << ((tid & 31) & ~(p.G - 1)));
    const int gshift = 31 - __clz(p.G);
    for (int e = tid >> gshift; e < nelem; e += kThreads >> gshift) {
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
        MixPrep<NC> P;
        mix_prepare_g<NC>(P, c, g);
        ElemResult res;
        if constexpr (!REV) {
            const MixEval ev = mix_eval_g<NC>(x, P, g);
            if (mix_fast_ok(ev)) res = mix_forward_fast<NC>(ev, P, p.use_reg != 0, p.reg_max, p.reg_factor);
            else res = mix_forward_f64(x, c.rec, c.pre ? nullptr : c.mfac, K, P.log_s, p.use_reg != 0, p.reg_max, p.reg_factor);
        } else {
            InvState<NC> st;
            if (!mix_inverse_g<NC>(x, P, g, g.sub == 0 ? p.status : nullptr, st, res))
                res = mix_inverse_f64(x, st.x, inv_slow_margin<NC>(st), c.rec, c.pre ? nullptr : c.mfac, K, P.log_s, st.lb0, st.ub0);
        }
        if (g.sub != 0) continue;   // every lane of the group holds the same result; lane 0 publishes it
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

template <int NC, bool REV>
int launch2(const MixParams& p, size_t smem, cudaStream_t stream) {
    if (smem > 48 * 1024)
        CNF_CUDA(cudaFuncSetAttribute(mixcdf_kernel<NC, REV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long grid = (p.P + p.TP - 1) / p.TP;
    mixcdf_kernel<NC, REV><<<(unsigned)grid, kThreads, smem, stream>>>(p);
    return launch_status(REV ? "mixcdf_inv_kernel" : "mixcdf_fwd_kernel");
}

template <int NC>
int launch(const MixParams& p, size_t smem, cudaStream_t stream) {
    return p.reverse ? launch2<NC, true>(p, smem, stream) : launch2<NC, false>(p, smem, stream);
}

int run(const cnf_mixcdf_args* a, cnf_stream_t stream_, int reverse) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    CNF_REQUIRE(a != nullptr, "args is NULL");
    CNF_REQUIRE(a->B >= 0 && a->S >= 0, "negative batch/sequence size");
    CNF_REQUIRE(a->C >= 1 && a->K >= 1, "C and K must be >= 1 (got C=%d K=%d)", a->C, a->K);
    CNF_SUPPORTED(a->C <= CNF_MAX_CHANNELS, "C=%d exceeds CNF_MAX_CHANNELS=%d", a->C, CNF_MAX_CHANNELS);
    CNF_SUPPORTED(a->K <= CNF_MAX_MIXTURES, "K=%d exceeds CNF_MAX_MIXTURES=%d", a->K, CNF_MAX_MIXTURES);
    CNF_SUPPORTED(a->S < (1ll << 31) && a->B * a->S < (1ll << 40), "tensor too large");
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
    static const bool force_generic = getenv("CNF_B200_MIXCDF_GENERIC") != nullptr;   // A/B switch for profiling
    if (p.mask.n_t > 0 && !force_generic) {
        int handled = 0;
        rc = mixcdf_pipe_try(a, p.mask, reverse, stream, &handled);              // compile-time (K, Ct), thread per element
        if (rc != CNF_OK || handled) return rc;
        rc = mixcdf_gpipe_try(a, p.mask, reverse, stream, &handled);             // any K: lane groups on the same pipeline
        if (rc != CNF_OK || handled) return rc;
    }
    CNF_SUPPORTED(!a->nn_compact, "the compact network-output layout needs one of the TMA pipelines (contiguous transformed "
                                   "channels, 16-byte aligned rows): query cnf_mixcdf_path with nn_compact set");
    CNF_SUPPORTED(a->next_actnorm_bias == nullptr && a->next_actnorm_scales == nullptr && a->next_conv_weight == nullptr,
                  "the fused next-block epilogue needs the pipelined layout (forward, C=16, 8 transformed channels, K=8); "
                  "query cnf_mixcdf_fusable first");
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
    // lanes per element: up to 8 components per lane (4 when K <= 4)
    const int NC = a->K <= 4 ? 4 : 8;
    int G = 1;
    while (NC * G < a->K) G <<= 1;
    CNF_SUPPORTED(G <= 32, "K=%d exceeds %d mixture components", a->K, NC * 32);
    p.G = G;
    // tile: about one element per lane group (two passes when groups are wide), bounded by the parameter staging
    // (~44 KB, 64 KB for wide groups where a record alone is several hundred bytes)
    int elems = G == 1 ? kThreads : 2 * kThreads / G;
    const int cap = ((G == 1 ? 44 : 64) * 1024) / (4 * p.PN);
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
    return NC == 4 ? launch<4>(p, smem, stream) : launch<8>(p, smem, stream);
}

}  // namespace
}  // namespace cnf

extern "C" int cnf_mixcdf_fusable(const cnf_mixcdf_args* a) {
    if (a == nullptr || a->C < 1 || a->C > CNF_MAX_CHANNELS) return 0;
    static const bool force_generic = getenv("CNF_B200_MIXCDF_GENERIC") != nullptr;
    cnf::MaskView mv{};
    if (force_generic || cnf::build_mask(a->mask, a->C, &mv) != CNF_OK) return 0;
    return cnf::mixcdf_pipe_fusable(a, mv, 0) ? 1 : 0;
}
extern "C" int cnf_mixcdf_path(const cnf_mixcdf_args* a) {
    if (a == nullptr || a->C < 1 || a->C > CNF_MAX_CHANNELS || a->K < 1 || a->K > CNF_MAX_MIXTURES) return -1;
    static const bool force_generic = getenv("CNF_B200_MIXCDF_GENERIC") != nullptr;
    cnf::MaskView mv{};
    if (cnf::build_mask(a->mask, a->C, &mv) != CNF_OK) return -1;
    if (force_generic || mv.n_t == 0) return 0;
    if (cnf::mixcdf_pipe_eligible(a, mv)) return 1;
    if (cnf::mixcdf_gpipe_eligible(a, mv)) return 2;
    return 0;
}
extern "C" int cnf_mixcdf_fwd(const cnf_mixcdf_args* a, cnf_stream_t stream) { return cnf::run(a, stream, 0); }
extern "C" int cnf_mixcdf_inv(const cnf_mixcdf_args* a, cnf_stream_t stream) { return cnf::run(a, stream, 1); }
