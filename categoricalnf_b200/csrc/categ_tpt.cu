// K6 fast path: thread-per-token mixture-of-logistics encode / decode for compile-time D.
//
// Same function as categ.cu (reference layers/categorical_encoding/linear_encoding.py:59-196,
// activation_normalization.py:116-144, distributions.py:117-163).  The posterior over the V classes
// costs V*D class-conditional logistic densities per token; this is a MUFU-bound kernel, so the
// arithmetic is arranged to need ONE MUFU op per (class, dimension):
//     sum_d [softplus(v_d) + softplus(-v_d)] = sum_d |v_d| + 2 ln prod_d (1 + e^{-|v_d|})
// (the product of D <= 32 factors in (1, 2] cannot overflow), i.e. one ex2 per (class, dim) and one
// lg2 per class, instead of an ex2 + lg2 pair per (class, dim).  One thread owns one token: its D
// latents stay in registers, every lane walks the class table in lock step, so the table reads are
// shared-memory broadcasts (16-byte loads, two dimensions each) and there are no shuffles and no
// idle lanes.  log2(e)/sigma is folded into the table.  The log-sum-exp over the classes uses the
// token's own forward value as reference (it is the dominant term unless the token is wildly
// unlikely); a running maximum is tracked and the rare overflow case is redone exactly.
#include <stdlib.h>

#include "cnf_common.cuh"
#include "philox.cuh"

namespace cnf {
namespace {

constexpr int kThreadsT = 128;
constexpr float kSigma = (float)(1.0 / 1.81);          // distributions.py:94
constexpr float kLogSigma = -0.59332686459844f;        // log(1/1.81)
constexpr float kEps = 1e-4f;                          // distributions.py:93
constexpr float kScale2 = kLog2e * 1.81f;              // log2(e) / sigma

struct TptParams {
    const long long* tokens; const float* u; const float* table; const float* prior; const float* pad;
    const float* z_in;
    float* z_out; float* ldj; float* cpl; long long* tokens_out; uint32_t* status;
    const float* nx_bias; const float* nx_scales; const float* nx_w;   // fused first-block ActNorm + 1x1 conv
    long long T;
    int S, V;
    float beta;
    unsigned long long seed, offset;
};

// shared tables for D dims (DP = D rounded up to even):
//   eb   [V][DP] float2: (e^{-s_vd} log2e/sigma, b_vd log2e/sigma)   -> v2 = z e' - b' = v log2(e)
//   own  [V][2D+1] floats: (b_vd | tanh s_vd), odd stride (gather by token class, conflict free)
//   cst  [V]: -D log sigma - sum_d s_vd + prior_v
template <int D>
struct Smem {
    static constexpr int DP = (D + 1) & ~1;
    static constexpr int OWN = 2 * D + 1;
    static size_t bytes(int V) {
        return sizeof(float) * ((size_t)V * DP * 2 + (size_t)V * OWN + 2 * (size_t)V + 4 + (size_t)D * D + 2 * D + 4);
    }
};

template <int D>
__device__ __forceinline__ void load_tables(const TptParams& p, float2* eb, float* own, float* cst, float* prior) {
    constexpr int DP = Smem<D>::DP, OWN = Smem<D>::OWN;
    const int V = p.V;
    for (int i = threadIdx.x; i < V * DP; i += kThreadsT) {
        const int v = i / DP, d = i - v * DP;
        float2 e = make_float2(0.f, 0.f);   // padding dimension: v2 = 0 is handled by the caller (never read)
        if (d < D) {
            const float b = p.table[v * 2 * D + d];
            const float s = tanh_from_2log2e(p.table[v * 2 * D + D + d] * (2.0f * kLog2e));
            e = make_float2(fast_exp(-s) * kScale2, b * kScale2);
            own[v * OWN + d] = b;
            own[v * OWN + D + d] = s;
        }
        eb[i] = e;
    }
    __syncthreads();
    for (int v = threadIdx.x; v < V; v += kThreadsT) {
        float a = 0.f;
        for (int d = 0; d < D; ++d) a += own[v * OWN + D + d];
        const float pr = p.prior ? p.prior[v] : 0.f;
        prior[v] = pr;
        cst[v] = -(float)D * kLogSigma - a + pr;
    }
    __syncthreads();
}

// log p(z | class v) + prior_v; z2 are the latents, table entries pre-scaled (see above)
template <int D>
__device__ __forceinline__ float class_score(const float (&z)[D], const float2* ebv, float cstv) {
    float sabs = 0.f, prod = 1.f;
#pragma unroll
    for (int d = 0; d + 1 < D; d += 2) {
        const float4 t = *reinterpret_cast<const float4*>(ebv + d);   // (e'_d, b'_d, e'_{d+1}, b'_{d+1})
        const float a0 = fabsf(fmaf(z[d], t.x, -t.y)), a1 = fabsf(fmaf(z[d + 1], t.z, -t.w));
        sabs += a0 + a1;
        prod *= (1.0f + ex2(-a0)) * (1.0f + ex2(-a1));
    }
    if (D & 1) {
        const float2 t = ebv[D - 1];
        const float a0 = fabsf(fmaf(z[D - 1], t.x, -t.y));
        sabs += a0;
        prod *= 1.0f + ex2(-a0);
    }
    // -(sum softplus(v)+softplus(-v)) = -ln2 (sum |v2| + 2 log2 prod)
    return fmaf(-kLn2, fmaf(2.0f, lg2(prod), sabs), cstv);
}

template <int D>
__global__ void __launch_bounds__(kThreadsT) categ_encode_tpt_kernel(const TptParams p) {
    extern __shared__ __align__(16) float sm[];
    constexpr int DP = Smem<D>::DP, OWN = Smem<D>::OWN;
    const int V = p.V;
    float2* eb = reinterpret_cast<float2*>(sm);
    float* own = sm + (size_t)V * DP * 2;
    float* cst = own + (size_t)V * OWN;
    float* prior = cst + V;
    float* s_w = prior + V + (4 - ((2 * V * DP + V * OWN + 2 * V) & 3)) % 4;   // 16-byte aligned [D][D] next conv weight
    float* s_nb = s_w + D * D;
    float* s_ne = s_nb + D;
    const bool fuse = p.nx_w != nullptr || p.nx_bias != nullptr || p.nx_scales != nullptr;
    if (fuse) {
        for (int i = threadIdx.x; i < D * D; i += kThreadsT) s_w[i] = p.nx_w ? p.nx_w[i] : (i / D == i % D ? 1.0f : 0.f);
        for (int i = threadIdx.x; i < D; i += kThreadsT) {
            s_nb[i] = p.nx_bias ? p.nx_bias[i] : 0.f;
            s_ne[i] = p.nx_scales ? expf(p.nx_scales[i]) : 1.0f;
        }
    }
    load_tables<D>(p, eb, own, cst, prior);

    const long long stride = (long long)gridDim.x * kThreadsT;
    const long long n_round = (p.T + 31) & ~31ll;
    for (long long tk = (long long)blockIdx.x * kThreadsT + threadIdx.x; tk < n_round; tk += stride) {
        const bool in = tk < p.T;
        float ldj_tok = 0.f;
        if (in) {
            const long long tok = p.tokens[tk];
            const float padv = p.pad ? p.pad[tk] : 1.0f;
            const bool tok_ok = tok >= 0 && tok < V;
            const int x = tok_ok ? (int)tok : 0;
            // ---- forward: z0 ~ Logistic(0, sigma), z = (z0 + b_x) e^{s_x} (linear_encoding.py:76-78) ----
            float z[D];
            float init_log_p = 0.f, ldj_fwd = 0.f;
#pragma unroll
            for (int d0 = 0; d0 < D; d0 += 4) {
                float u[4];
                if (p.u) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) u[i] = (d0 + i < D) ? p.u[tk * D + d0 + i] : 0.5f;
                } else {
                    // element e = tk*D + d uses draw (e & 3) of Philox counter offset + (e >> 2), as categ.cu
                    const unsigned long long e0 = (unsigned long long)tk * D + d0;
                    if ((D & 3) == 0) {
                        philox_uniform4(p.seed, p.offset + (e0 >> 2), u);
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            float r[4];
                            philox_uniform4(p.seed, p.offset + ((e0 + i) >> 2), r);
                            u[i] = r[(e0 + i) & 3];
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int d = d0 + i;
                    if (d < D) {
                        const float z0 = __fmul_rn(logistic_from_uniform(u[i], kEps), kSigma);
                        init_log_p -= softplus_pm(__fdiv_rn(z0, kSigma)) + kLogSigma;
                        const float th = own[x * OWN + D + d];
                        ldj_fwd += th;
                        z[d] = (z0 + own[x * OWN + d]) * fast_exp(th);
                    }
                }
            }
            const float log_point = init_log_p - ldj_fwd + prior[x];
            // ---- exact posterior over the V classes (:153-174), reference point = own forward value ----
            float ssum = 0.f, mx = -INFINITY;
            for (int v = 0; v < V; ++v) {
                float sc = class_score<D>(z, eb + v * DP, cst[v]);
                if (v == x) sc = log_point;   // own class: forward value (:167-168)
                mx = fmaxf(mx, sc);
                ssum += ex2((sc - log_point) * kLog2e);
            }
            float lse;
            if (mx - log_point < 80.0f && ssum == ssum) {
                lse = log_point + fast_log(ssum);
            } else {   // another class dominates by e^80 (or NaN upstream): exact running-max form
                float m = -INFINITY, s2 = 0.f;
                for (int v = 0; v < V; ++v) {
                    float sc = class_score<D>(z, eb + v * DP, cst[v]);
                    if (v == x) sc = log_point;
                    const float mn = fmaxf(m, sc);
                    if (mn > -INFINITY) s2 = s2 * fast_exp(m - mn) + fast_exp(sc - mn);
                    m = mn;
                }
                lse = m + fast_log(s2);
            }
            const float cpl = log_point - lse;
            ldj_tok = (p.beta * cpl - (init_log_p - ldj_fwd)) * padv;
            if (fuse) {
                // first block: a = (z pad + bias) e^{scales} pad ; y = (a @ W) pad  (ActNormFlow, InvertibleConv)
                float a[D];
#pragma unroll
                for (int d = 0; d < D; ++d) a[d] = (z[d] * padv + s_nb[d]) * s_ne[d] * padv;
#pragma unroll
                for (int d = 0; d < D; ++d) z[d] = 0.f;
#pragma unroll
                for (int c = 0; c < D; ++c) {
                    if constexpr ((D & 3) == 0) {
#pragma unroll
                        for (int o = 0; o < D; o += 4) {
                            const float4 w = *reinterpret_cast<const float4*>(s_w + c * D + o);   // broadcast load
                            z[o] = fmaf(a[c], w.x, z[o]); z[o + 1] = fmaf(a[c], w.y, z[o + 1]);
                            z[o + 2] = fmaf(a[c], w.z, z[o + 2]); z[o + 3] = fmaf(a[c], w.w, z[o + 3]);
                        }
                    } else {
#pragma unroll
                        for (int o = 0; o < D; ++o) z[o] = fmaf(a[c], s_w[c * D + o], z[o]);
                    }
                }
            }
            float* zo = p.z_out + tk * D;
            if ((D & 3) == 0) {
#pragma unroll
                for (int d = 0; d < D; d += 4)
                    *reinterpret_cast<float4*>(zo + d) = make_float4(z[d] * padv, z[d + 1] * padv, z[d + 2] * padv, z[d + 3] * padv);
            } else {
#pragma unroll
                for (int d = 0; d < D; ++d) zo[d] = z[d] * padv;
            }
            if (p.cpl) p.cpl[tk] = cpl;
            if (ldj_tok != ldj_tok || !tok_ok)
                flag(p.status, (ldj_tok != ldj_tok ? CNF_FLAG_NAN_LDJ : 0u) | (!tok_ok ? CNF_FLAG_CDF_RANGE : 0u));
        }
        warp_segmented_atomic_add(p.ldj, in ? tk / p.S : 0, ldj_tok, in);
    }
}

template <int D>
__global__ void __launch_bounds__(kThreadsT) categ_decode_tpt_kernel(const TptParams p) {
    extern __shared__ __align__(16) float sm[];
    constexpr int DP = Smem<D>::DP, OWN = Smem<D>::OWN;
    const int V = p.V;
    float2* eb = reinterpret_cast<float2*>(sm);
    float* own = sm + (size_t)V * DP * 2;
    float* cst = own + (size_t)V * OWN;
    float* prior = cst + V;
    load_tables<D>(p, eb, own, cst, prior);
    const long long stride = (long long)gridDim.x * kThreadsT;
    for (long long tk = (long long)blockIdx.x * kThreadsT + threadIdx.x; tk < p.T; tk += stride) {
        float z[D];
#pragma unroll
        for (int d = 0; d < D; ++d) z[d] = p.z_in[tk * D + d];
        float best = -INFINITY;
        int arg = 0;
        for (int v = 0; v < V; ++v) {   // first maximum wins, like torch.argmax (:196)
            const float sc = class_score<D>(z, eb + v * DP, cst[v]);
            if (sc > best || v == 0) { best = sc; arg = v; }
        }
        p.tokens_out[tk] = arg;
    }
}

template <int D>
int launch_encode(const TptParams& p, cudaStream_t stream) {
    const size_t smem = Smem<D>::bytes(p.V);
    if (smem > 48 * 1024)
        CNF_CUDA(cudaFuncSetAttribute(categ_encode_tpt_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long blocks = (p.T + kThreadsT - 1) / kThreadsT;
    const long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    categ_encode_tpt_kernel<D><<<(unsigned)blocks, kThreadsT, smem, stream>>>(p);
    return launch_status("categ_encode_tpt_kernel");
}

template <int D>
int launch_decode(const TptParams& p, cudaStream_t stream) {
    const size_t smem = Smem<D>::bytes(p.V);
    if (smem > 48 * 1024)
        CNF_CUDA(cudaFuncSetAttribute(categ_decode_tpt_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long blocks = (p.T + kThreadsT - 1) / kThreadsT;
    const long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    categ_decode_tpt_kernel<D><<<(unsigned)blocks, kThreadsT, smem, stream>>>(p);
    return launch_status("categ_decode_tpt_kernel");
}

size_t smem_for(int V, int D) {
    const int DP = (D + 1) & ~1;
    return sizeof(float) * ((size_t)V * DP * 2 + (size_t)V * (2 * D + 1) + 2 * (size_t)V + 4);
}

}  // namespace

#define CNF_TPT_DISPATCH(FN)              \
    switch (D) {                          \
        case 1: return FN<1>(p, stream);  \
        case 2: return FN<2>(p, stream);  \
        case 3: return FN<3>(p, stream);  \
        case 4: return FN<4>(p, stream);  \
        case 6: return FN<6>(p, stream);  \
        case 8: return FN<8>(p, stream);  \
        case 12: return FN<12>(p, stream); \
        case 16: return FN<16>(p, stream); \
        default: break;                   \
    }

static bool tpt_eligible(int V, int D, long long T) {
    if (!(D == 1 || D == 2 || D == 3 || D == 4 || D == 6 || D == 8 || D == 12 || D == 16)) return false;
    if (smem_for(V, D) > 96 * 1024) return false;
    static const bool force_warp = getenv("CNF_B200_CATEG_WARP") != nullptr;   // A/B switch
    if (force_warp) return false;
    // one thread sweeps all V*D class terms of a token: with few tokens and a big table the
    // warp-per-token kernel (categ.cu), which spreads a token over 32 lanes, has the lower latency
    return T >= 2048 || V * D <= 128;
}

int categ_encode_tpt_try(const cnf_categ_encode_args* a, cudaStream_t stream, int* handled) {
    *handled = 0;
    const int D = a->D;
    const long long T = a->B * a->S;
    if (!tpt_eligible(a->V, D, T)) return CNF_OK;
    if ((D & 3) == 0 && (reinterpret_cast<uintptr_t>(a->z_out) & 15)) return CNF_OK;
    TptParams p{};
    p.tokens = reinterpret_cast<const long long*>(a->tokens); p.u = a->u_noise; p.table = a->table;
    p.prior = a->category_prior; p.pad = a->pad; p.z_out = a->z_out; p.ldj = a->ldj; p.cpl = a->class_prob_log;
    p.status = a->status; p.T = T; p.S = (int)a->S; p.V = a->V; p.beta = a->beta; p.seed = a->seed; p.offset = a->offset;
    p.nx_bias = a->next_actnorm_bias; p.nx_scales = a->next_actnorm_scales; p.nx_w = a->next_conv_weight;
    *handled = 1;
    CNF_TPT_DISPATCH(launch_encode)
    *handled = 0;
    return CNF_OK;
}

bool categ_encode_tpt_fusable(const cnf_categ_encode_args* a) {
    if (!tpt_eligible(a->V, a->D, a->B * a->S)) return false;
    return !((a->D & 3) == 0 && (reinterpret_cast<uintptr_t>(a->z_out) & 15));
}

int categ_decode_tpt_try(const cnf_categ_decode_args* a, cudaStream_t stream, int* handled) {
    *handled = 0;
    const int D = a->D;
    const long long T = a->B * a->S;
    if (!tpt_eligible(a->V, D, T)) return CNF_OK;
    TptParams p{};
    p.z_in = a->z; p.table = a->table; p.prior = a->category_prior;
    p.tokens_out = reinterpret_cast<long long*>(a->tokens_out); p.T = T; p.S = (int)a->S; p.V = a->V;
    *handled = 1;
    CNF_TPT_DISPATCH(launch_decode)
    *handled = 0;
    return CNF_OK;
}

}  // namespace cnf
