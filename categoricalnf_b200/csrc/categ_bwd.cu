// Backward of K6 (mixture-of-logistics categorical encoding, LinearCategoricalEncoding with num_flows = 0):
// gradient of a loss with respect to the class table [V, 2D] = pred_net(embed(v)) = (bias_v | raw log-scale_v),
// given dL/dz and dL/dldj.  The reference gets it from autograd through the all-class expansion
// (layers/categorical_encoding/linear_encoding.py:71-92,153-174: a [B*S*V,1,D] tensor pushed backwards through
// ExtActNorm + embedding + Linear: ~110 ms of a 156 ms training step at the LM shape when done with eager ops on the GPU).
//
// Per token with class t, latent z = (z0 + b_t) e^{s_t} (s = tanh raw), per-sample upstream gradients g_z, g_l:
//   den_c = sum_d logp(v_cd) - sum_d s_cd + prior_c,  v_cd = z_d e^{-s_cd} - b_cd      (den_t = the forward value: v_t = z0)
//   q_c = softmax(den)_c,  log q(t|z) = den_t - logsumexp(den),  ldj_tok = beta log q(t|z) - log p(z0) + sum_d s_td
//   G_c = g_l beta (delta_ct - q_c)
//   c != t:  d/db_cd = -G_c lp'(v_cd),  d/ds_cd = G_c (lp'(v_cd) (-z_d e^{-s_cd}) - 1),  Z_d += G_c lp'(v_cd) e^{-s_cd}
//   c == t:  d/ds_td = -G_t + g_l + (g_z_d + Z_d) z_d,   d/db_td = (g_z_d + Z_d) e^{s_td}      (z0 is a constant: noise)
//   lp'(v) = -tanh(v / 2 sigma) / sigma,   d/draw = d/ds (1 - s^2)
//
// Persistent CTAs, tiles of kTile tokens:  A1 thread per (token, class): den;  A2 thread per token: softmax -> G;
// B thread per (class, dim) pair walking the tile's tokens with its gradient pair in registers (no atomics), Z_d
// reduced over the classes through warp shuffles + shared atomics;  C thread per (token, dim): own-class terms into a
// shared [V,2D] accumulator.  One global atomic per (CTA, table entry) at the end.
#include <stdlib.h>

#include "cnf_common.cuh"

namespace cnf {
namespace {

constexpr int kThreadsB = 512;
constexpr int kTile = 64;
constexpr int kMaxPairsPerThread = 8;
constexpr float kSigmaB = (float)(1.0 / 1.81);          // distributions.py:94
constexpr float kInvSigmaB = 1.81f;

struct EncBwdParams {
    const long long* tokens; const float* z; const float* table; const float* prior; const float* pad;
    const float* gz; const float* gldj; float* gtable;
    long long M;
    int S, V, D;
    float beta;
};

__global__ void __launch_bounds__(kThreadsB, 1) categ_encode_bwd_kernel(const EncBwdParams p) {
    extern __shared__ __align__(16) float sm[];
    const int V = p.V, D = p.D, VD = V * D;
    const int DS = D | 1;                  // odd row pitch of the class tables: lanes walking the classes hit distinct banks
    float* s_b = sm;                       // [V][DS] bias
    float* s_E = s_b + V * DS;             // [V][DS] e^{-s}
    float* s_s = s_E + V * DS;             // [V][DS] s = tanh(raw)
    float* s_cst = s_s + V * DS;           // [V]   -sum_d s_cd + prior_c
    float* s_acc = s_cst + V;              // [V][2D] own-class gradient accumulator (db | ds)
    float* s_z = s_acc + 2 * VD;           // [kTile][D]
    float* s_gz = s_z + kTile * D;         // [kTile][D] g_z * pad
    float* s_Z = s_gz + kTile * D;         // [kTile][D]
    float* s_G = s_Z + kTile * D;          // [kTile][V] den, then G
    float* s_gl = s_G + kTile * V;         // [kTile] g_l * pad
    int* s_tok = reinterpret_cast<int*>(s_gl + kTile);   // [kTile], -1 = inactive token
    // D | 32: per-warp partial sums of Z (reduced over the classes a warp holds) - plain stores, no atomics
    float* s_Zw = reinterpret_cast<float*>(s_tok + kTile);   // [kThreadsB/32][kTile][D]
    const bool zw = (32 % D) == 0;

    const int tid = threadIdx.x;
    for (int i = tid; i < VD; i += kThreadsB) {
        const int c = i / D, d = i - c * D;
        const float s = tanhf(p.table[(size_t)c * 2 * D + D + d]);
        s_b[c * DS + d] = p.table[(size_t)c * 2 * D + d];
        s_s[c * DS + d] = s;
        s_E[c * DS + d] = __expf(-s);
    }
    for (int i = tid; i < 2 * VD; i += kThreadsB) s_acc[i] = 0.f;
    __syncthreads();
    for (int c = tid; c < V; c += kThreadsB) {
        float a = p.prior[c];
        for (int d = 0; d < D; ++d) a -= s_s[c * DS + d];
        s_cst[c] = a;
    }

    // this thread's (class, dim) pairs and their register accumulators
    float gb[kMaxPairsPerThread], gs[kMaxPairsPerThread];
#pragma unroll
    for (int k = 0; k < kMaxPairsPerThread; ++k) { gb[k] = 0.f; gs[k] = 0.f; }

    const long long ntiles = (p.M + kTile - 1) / kTile;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long m0 = tile * kTile;
        const int rows = (int)min((long long)kTile, p.M - m0);
        __syncthreads();   // previous tile fully consumed
        for (int i = tid; i < kTile * D; i += kThreadsB) {
            const int r = i / D;
            float zv = 0.f, gv = 0.f;
            if (r < rows) {
                const float pd = p.pad ? p.pad[m0 + r] : 1.0f;
                zv = p.z[(m0 + r) * D + (i - r * D)];
                gv = p.gz[(m0 + r) * D + (i - r * D)] * pd;
            }
            s_z[i] = zv;
            s_gz[i] = gv;
            s_Z[i] = 0.f;
        }
        for (int r = tid; r < kTile; r += kThreadsB) {
            int t = -1;
            float gl = 0.f;
            if (r < rows) {
                const float pd = p.pad ? p.pad[m0 + r] : 1.0f;
                if (pd != 0.0f) {
                    t = (int)p.tokens[m0 + r];
                    gl = (p.gldj ? p.gldj[(m0 + r) / p.S] : 0.f) * pd;
                }
            }
            s_tok[r] = t;
            s_gl[r] = gl;
        }
        __syncthreads();

        // ---- A1: den[token][class] ---------------------------------------------------------------------------------
        for (int i = tid; i < kTile * V; i += kThreadsB) {
            const int r = i / V, c = i - r * V;
            float den = 0.f;
            if (s_tok[r] >= 0) {
                // sum_d logp(v) = -sum |w| - 2 ln prod (1 + e^{-|w|}) - D log sigma (the constant cancels in the softmax)
                float sabs = 0.f, prod = 1.f;
                const float* zr = s_z + r * D;
                for (int d = 0; d < D; ++d) {
                    const float w = fabsf(fmaf(zr[d], s_E[c * DS + d], -s_b[c * DS + d])) * kInvSigmaB;
                    sabs += w;
                    prod *= 1.0f + fast_exp(-w);
                    if ((d & 15) == 15) { sabs += 2.0f * fast_log(prod); prod = 1.f; }     // keep the product in range
                }
                den = s_cst[c] - sabs - 2.0f * fast_log(prod);
            }
            s_G[i] = den;
        }
        __syncthreads();
        // ---- A2: G_c = g_l beta (delta_ct - softmax(den)_c): one warp per token, lanes over the classes ------------------
        for (int r = tid >> 5; r < kTile; r += kThreadsB / 32) {
            const int t = s_tok[r], lane = tid & 31;
            float* g = s_G + r * V;
            if (t < 0) {
                for (int c = lane; c < V; c += 32) g[c] = 0.f;
                continue;
            }
            float mx = -3.0e38f;
            for (int c = lane; c < V; c += 32) mx = fmaxf(mx, g[c]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            float sum = 0.f;
            for (int c = lane; c < V; c += 32) sum += fast_exp(g[c] - mx);
            sum = warp_sum(sum);
            const float scale = s_gl[r] * p.beta, inv = 1.0f / sum;
            for (int c = lane; c < V; c += 32) g[c] = scale * ((c == t ? 1.0f : 0.0f) - fast_exp(g[c] - mx) * inv);
        }
        __syncthreads();
        // ---- B: per (class, dim) pair over the tile's tokens -------------------------------------------------------
#pragma unroll
        for (int k = 0; k < kMaxPairsPerThread; ++k) {
            const int q = tid + k * kThreadsB;
            if (q - (q & 31) >= VD) continue;                // whole warp beyond the last pair (warp-uniform)
            const bool live = q < VD;
            const int c = live ? q / D : 0, d = live ? q - c * D : 0;
            const float E = s_E[c * DS + d], bb = s_b[c * DS + d];
            float ab = 0.f, as = 0.f;
            const float* gcol = s_G + c;
            const float* zcol = s_z + d;
            float* slot = s_Zw + (size_t)(tid >> 5) * kTile * D + d;
            const bool writer = zw && (tid & 31) < D;
            // branch-free body: G is zero for inactive tokens, the own class (c == t) contributes only -G to d/ds
#pragma unroll 4
            for (int r = 0; r < rows; ++r) {
                const float g = live ? gcol[r * V] : 0.f;
                const float zd = zcol[r * D];
                const float v = fmaf(zd, E, -bb);
                const float lp = -kInvSigmaB * tanh_from_2log2e(v * (kInvSigmaB * kLog2e));   // -tanh(v / 2 sigma) / sigma
                const float glp = (c == s_tok[r]) ? 0.f : g * lp;
                ab -= glp;
                as -= fmaf(glp, zd * E, g);
                float zc = glp * E;
                // Z_d: sum over the classes.  With D | 32 the lanes l, l+D, ... of a warp share d: combine them, then the
                // warp keeps its partial in its own slot (this thread owns the slot for all of its pair groups k)
                if (zw) {
                    for (int o = 16; o >= D; o >>= 1) zc += __shfl_xor_sync(0xffffffffu, zc, o);
                    if (writer) slot[r * D] = (k == 0 ? 0.f : slot[r * D]) + zc;
                } else if (zc != 0.f) {
                    atomicAdd(s_Z + r * D + d, zc);
                }
            }
            gb[k] += ab;
            gs[k] += as;
        }
        __syncthreads();
        // ---- C: own-class terms ------------------------------------------------------------------------------------
        for (int i = tid; i < rows * D; i += kThreadsB) {
            const int r = i / D, d = i - r * D;
            const int t = s_tok[r];
            if (t < 0) continue;
            float zsum = s_Z[i];
            if (zw) {
                const int nw = min(kThreadsB / 32, (VD + 31) / 32);       // warps that own at least one pair in group k = 0
                for (int w = 0; w < nw; ++w) zsum += s_Zw[((size_t)w * kTile + r) * D + d];
            }
            const float zt = s_gz[i] + zsum;
            const float zd = s_z[i];
            atomicAdd(s_acc + t * 2 * D + d, zt / s_E[t * DS + d]);           // (g_z + Z) e^{s_t}
            atomicAdd(s_acc + t * 2 * D + D + d, fmaf(zt, zd, s_gl[r]));       // (g_z + Z) z + g_l
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kMaxPairsPerThread; ++k) {
        const int q = tid + k * kThreadsB;
        if (q >= VD) continue;
        const int c = q / D, d = q - c * D;
        const float s = s_s[c * DS + d];
        const float db = gb[k] + s_acc[c * 2 * D + d];
        const float ds = gs[k] + s_acc[c * 2 * D + D + d];
        if (db != 0.f) atomicAdd(p.gtable + (size_t)c * 2 * D + d, db);
        const float draw = ds * (1.0f - s * s);
        if (draw != 0.f) atomicAdd(p.gtable + (size_t)c * 2 * D + D + d, draw);
    }
}


// ------------------------------------------------------------------------------------------------------------------
// D = 16: thread per TOKEN (the decomposition of the forward kernel categ_encode_tpt_kernel, which sustains the MUFU
// rate on the same V*D class terms).  A thread keeps its token's latents in registers and walks the class table twice -
// every lane in lock step, so the table reads are shared-memory broadcasts:
//   sweep 1: den_c for all classes (one ex2 per (class, dim), one lg2 per class), kept in a thread-private column of
//            shared memory; running maximum -> logsumexp
//   sweep 2: G_c, then per (class, dim) lp'(v) from e = 2^{-|v2|} (ex2 + rcp), Z_d in registers (the sum over the
//            classes is thread-local here - the tile kernel below needs shuffles + a barrier for it), and the 32
//            per-class table-gradient terms (16 d/db, 16 d/ds) reduced over the warp's 32 tokens by a transposing
//            reduction: 31 shuffles for all 32 values (a butterfly per value would need 160), after which lane l holds
//            the warp's sum of term l and adds it to the CTA's [V][32] accumulator with one conflict-free shared atomic.
// No __syncthreads in the token loop.  Measured at the LM shape (1 M tokens, V 51): 3.16 ms (tile kernel) -> see DESIGN.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kThreadsT16 = 128;
constexpr float kScale2B = kLog2e * kInvSigmaB;        // log2(e) / sigma: v2 = v log2(e) / sigma
constexpr float kUnscale2B = kSigmaB * kLn2;           // E = E' * sigma / log2(e)

// sum over the 32 lanes of a warp of 32 per-lane values v[0..31]; lane l returns the total of v[l]
__device__ __forceinline__ float warp_transpose_sum32(float (&v)[32], int lane) {
#pragma unroll
    for (int half = 16; half >= 1; half >>= 1) {
        const bool up = (lane & half) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const float keep = up ? v[i + half] : v[i];
            const float send = up ? v[i] : v[i + half];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
        }
    }
    return v[0];
}

__global__ void __launch_bounds__(kThreadsT16) categ_encode_bwd_tpt16_kernel(const EncBwdParams p) {
    constexpr int D = 16;
    extern __shared__ __align__(16) float sm[];
    const int V = p.V;
    float4* s_eb = reinterpret_cast<float4*>(sm);          // [V][8] (E'_d, b'_d, E'_{d+1}, b'_{d+1})
    float* s_s = sm + (size_t)V * 32;                      // [V][17] s = tanh(raw) (gathered by token class: odd pitch)
    float* s_Eo = s_s + (size_t)V * 17;                    // [V][17] e^{-s}
    float* s_cst = s_Eo + (size_t)V * 17;                  // [V] -sum_d s_cd + prior_c
    float* s_acc = s_cst + V + ((4 - (V & 3)) & 3);        // [V][32] (db_0..15 | ds_0..15)
    float* s_den = s_acc + (size_t)V * 32;                 // [V][kThreadsT16] thread-private columns
    const int tid = threadIdx.x, lane = tid & 31;

    for (int i = tid; i < V * D; i += kThreadsT16) {
        const int c = i >> 4, d = i & 15;
        const float s = tanhf(p.table[(size_t)c * 2 * D + D + d]);
        const float E = __expf(-s);
        float* q = reinterpret_cast<float*>(s_eb + c * 8 + (d >> 1)) + 2 * (d & 1);
        q[0] = E * kScale2B;
        q[1] = p.table[(size_t)c * 2 * D + d] * kScale2B;
        s_s[c * 17 + d] = s;
        s_Eo[c * 17 + d] = E;
    }
    for (int i = tid; i < V * 32; i += kThreadsT16) s_acc[i] = 0.f;
    __syncthreads();
    for (int c = tid; c < V; c += kThreadsT16) {
        float a = p.prior[c];
        for (int d = 0; d < D; ++d) a -= s_s[c * 17 + d];
        s_cst[c] = a;
    }
    __syncthreads();

    float* den = s_den + tid;
    const long long n_round = (p.M + 31) & ~31ll;
    const long long stride = (long long)gridDim.x * kThreadsT16;
    for (long long tk = (long long)blockIdx.x * kThreadsT16 + tid; tk < n_round; tk += stride) {
        const bool in = tk < p.M;
        float z[D], Z[D];
        float pd = 0.f, gl = 0.f;
        int t = -1;
        if (in) {
            pd = p.pad ? p.pad[tk] : 1.0f;
            if (pd != 0.0f) {
                const long long tok = p.tokens[tk];
                if (tok >= 0 && tok < V) {      // (the forward kernel flags out-of-range tokens; here they get no gradient)
                    t = (int)tok;
                    gl = (p.gldj ? p.gldj[tk / p.S] : 0.f) * pd;
                }
            }
        }
#pragma unroll
        for (int d4 = 0; d4 < D; d4 += 4) {
            float4 zv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (in) zv = *reinterpret_cast<const float4*>(p.z + tk * D + d4);
            z[d4] = zv.x; z[d4 + 1] = zv.y; z[d4 + 2] = zv.z; z[d4 + 3] = zv.w;
            Z[d4] = Z[d4 + 1] = Z[d4 + 2] = Z[d4 + 3] = 0.f;
        }
        // ---- sweep 1: den_c, running max ---------------------------------------------------------------------
        float mx = -3.0e38f;
        for (int c = 0; c < V; ++c) {
            const float4* eb = s_eb + c * 8;
            float sabs = 0.f, prod = 1.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 tb = eb[i];
                const float a0 = fabsf(fmaf(z[2 * i], tb.x, -tb.y)), a1 = fabsf(fmaf(z[2 * i + 1], tb.z, -tb.w));
                sabs += a0 + a1;
                prod *= (1.0f + ex2(-a0)) * (1.0f + ex2(-a1));
            }
            // sum_d logp(v) = -ln2 (sum |v2| + 2 log2 prod) - D log sigma (the constant cancels in the softmax)
            const float dc = fmaf(-kLn2, fmaf(2.0f, lg2(prod), sabs), s_cst[c]);
            den[c * kThreadsT16] = dc;
            mx = fmaxf(mx, dc);
        }
        float ssum = 0.f;
        for (int c = 0; c < V; ++c) ssum += ex2((den[c * kThreadsT16] - mx) * kLog2e);
        const float scale = gl * p.beta, nlse2 = -(mx * kLog2e + lg2(ssum));      // -logsumexp * log2(e)
        // ---- sweep 2: table-gradient terms of every class, Z_d ----------------------------------------------------
        for (int c = 0; c < V; ++c) {
            const bool own = c == t;
            const float G = (t >= 0) ? scale * ((own ? 1.0f : 0.0f) - ex2(fmaf(den[c * kThreadsT16], kLog2e, nlse2))) : 0.f;
            const float Gl = own ? 0.f : G * (-kInvSigmaB);      // G lp' = Gl * tanh(v / 2 sigma); own class: only -G in d/ds
            const float4* eb = s_eb + c * 8;
            float val[32];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 tb = eb[i];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int d = 2 * i + h;
                    const float E2 = h ? tb.z : tb.x, b2 = h ? tb.w : tb.y;
                    const float v2 = fmaf(z[d], E2, -b2);
                    const float e = ex2(-fabsf(v2));
                    const float th = copysignf((1.0f - e) * rcp(1.0f + e), v2);      // tanh(v / 2 sigma)
                    const float glp = Gl * th;
                    const float gE = glp * (E2 * kUnscale2B);                         // G lp' e^{-s}
                    val[d] = -glp;
                    val[16 + d] = -fmaf(gE, z[d], G);
                    Z[d] += gE;
                }
            }
            const float tot = warp_transpose_sum32(val, lane);
            if (tot != 0.f) atomicAdd(s_acc + c * 32 + lane, tot);
        }
        // ---- own-class terms ---------------------------------------------------------------------------------------
        if (t >= 0) {
#pragma unroll
            for (int d4 = 0; d4 < D; d4 += 4) {
                const float4 g4 = *reinterpret_cast<const float4*>(p.gz + tk * D + d4);
                const float gq[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int d = d4 + j;
                    const float zt = fmaf(gq[j], pd, Z[d]);
                    atomicAdd(s_acc + t * 32 + d, zt / s_Eo[t * 17 + d]);              // (g_z + Z) e^{s_t}
                    atomicAdd(s_acc + t * 32 + 16 + d, fmaf(zt, z[d], gl));            // (g_z + Z) z + g_l
                }
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < V * 32; i += kThreadsT16) {
        const int c = i >> 5, j = i & 31;
        const float a = s_acc[i];
        if (a == 0.f) continue;
        if (j < 16) {
            atomicAdd(p.gtable + (size_t)c * 2 * D + j, a);
        } else {
            const float s = s_s[c * 17 + (j - 16)];
            atomicAdd(p.gtable + (size_t)c * 2 * D + D + (j - 16), a * (1.0f - s * s));
        }
    }
}

}  // namespace
}  // namespace cnf

extern "C" int cnf_categ_encode_bwd(const cnf_categ_encode_bwd_args* a, cnf_stream_t stream_) {
    using namespace cnf;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    CNF_REQUIRE(a != nullptr, "cnf_categ_encode_bwd: null args");
    CNF_REQUIRE(a->B >= 0 && a->S >= 1 && a->V >= 1 && a->D >= 1, "cnf_categ_encode_bwd: bad shape");
    if (a->B == 0) return CNF_OK;
    CNF_REQUIRE(a->tokens && a->z && a->table && a->category_prior && a->grad_z && a->grad_table, "cnf_categ_encode_bwd: null tensor");
    CNF_SUPPORTED(a->D <= 64 && (long long)a->V * a->D <= (long long)kThreadsB * kMaxPairsPerThread,
                  "cnf_categ_encode_bwd: V*D = %lld exceeds %d", (long long)a->V * a->D, kThreadsB * kMaxPairsPerThread);
    EncBwdParams p{};
    p.tokens = reinterpret_cast<const long long*>(a->tokens);
    p.z = a->z; p.table = a->table; p.prior = a->category_prior; p.pad = a->pad;
    p.gz = a->grad_z; p.gldj = a->grad_ldj; p.gtable = a->grad_table;
    p.M = a->B * a->S; p.S = (int)a->S; p.V = a->V; p.D = a->D; p.beta = a->beta;
    static const bool force_tile = getenv("CNF_B200_CATEG_BWD_TILE") != nullptr;      // A/B switch
    if (a->D == 16 && !force_tile && ((reinterpret_cast<uintptr_t>(a->z) | reinterpret_cast<uintptr_t>(a->grad_z)) & 15) == 0) {
        const size_t smem16 = sizeof(float) * ((size_t)a->V * (32 + 17 + 17 + 1 + 32) + 4 + (size_t)a->V * kThreadsT16);
        if (smem16 <= 100 * 1024) {
            if (smem16 > 48 * 1024)
                CNF_CUDA(cudaFuncSetAttribute(categ_encode_bwd_tpt16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem16));
            long long blocks = (p.M + kThreadsT16 - 1) / kThreadsT16;
            const long long cap = (long long)sm_count() * 4;
            if (blocks > cap) blocks = cap;
            categ_encode_bwd_tpt16_kernel<<<(unsigned)blocks, kThreadsT16, smem16, stream>>>(p);
            return launch_status("categ_encode_bwd_tpt16_kernel");
        }
    }
    const size_t VD = (size_t)a->V * a->D;
    const size_t smem = sizeof(float) * (3 * (size_t)a->V * (a->D | 1) + a->V + 2 * VD + 3 * (size_t)kTile * a->D + (size_t)kTile * a->V + 2 * kTile +
                                         ((32 % a->D) == 0 ? (size_t)(kThreadsB / 32) * kTile * a->D : 0));
    CNF_SUPPORTED(smem <= 200 * 1024, "cnf_categ_encode_bwd: V=%d D=%d needs %zu bytes of shared memory", a->V, a->D, smem);
    if (smem > 48 * 1024)
        CNF_CUDA(cudaFuncSetAttribute(categ_encode_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long ntiles = (p.M + kTile - 1) / kTile;
    long long grid = 2ll * sm_count();
    if (grid > ntiles) grid = ntiles;
    categ_encode_bwd_kernel<<<(unsigned)grid, kThreadsB, smem, stream>>>(p);
    return launch_status("categ_encode_bwd_kernel");
}
