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
