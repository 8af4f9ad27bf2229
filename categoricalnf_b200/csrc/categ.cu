// K6: mixture-of-logistics categorical encoding / decoding.
// Replaces LinearCategoricalEncoding.forward for the mixture model (num_flows = 0):
//   reference layers/categorical_encoding/linear_encoding.py:59-196 (+ ExtActNormFlow
//   activation_normalization.py:116-144, LogisticDistribution distributions.py:117-163).
// The reference materialises a [B*S*V, 1, D] tensor to evaluate every class-conditional density;
// here one warp owns one token: lanes 0..D-1 draw / transform the latent, then the lanes sweep
// the V classes (class-major shared-memory tables, conflict free) and reduce the posterior with
// shuffles.  Nothing but tokens (8 B), optional noise (4D B) and z (4D B) touches HBM.
#include "cnf_common.cuh"
#include "philox.cuh"

namespace cnf {
namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr float kSigma = (float)(1.0 / 1.81);          // distributions.py:94
constexpr float kLogSigma = -0.59332686459844f;        // log(1/1.81)
constexpr float kEps = 1e-4f;                          // distributions.py:93

struct CategParams {
    const long long* tokens; const float* u; const float* table; const float* prior; const float* pad;
    const float* z_in;
    float* z_out; float* ldj; float* cpl; long long* tokens_out; uint32_t* status;
    long long T;  // B*S tokens
    int S, V, D;
    float beta;
    unsigned long long seed, offset;
};

// shared tables: bT/eT are [D][V] (class index fastest) for the posterior sweep, th/bv are [V][D]
struct Tables {
    float *bT, *eT, *th, *bv, *ssum, *prior;
};

__device__ __forceinline__ Tables carve(float* sm, int V, int D) {
    Tables t;
    t.bT = sm; t.eT = t.bT + V * D; t.th = t.eT + V * D; t.bv = t.th + V * D;
    t.ssum = t.bv + V * D; t.prior = t.ssum + V;
    return t;
}

__device__ __forceinline__ void load_tables(const CategParams& p, const Tables& t) {
    const int V = p.V, D = p.D;
    for (int i = threadIdx.x; i < V * D; i += kThreads) {
        const int v = i / D, d = i - v * D;
        const float b = p.table[v * 2 * D + d];
        const float s = tanh_from_2log2e(p.table[v * 2 * D + D + d] * (2.0f * kLog2e));
        t.bv[i] = b; t.th[i] = s;
        t.bT[d * V + v] = b; t.eT[d * V + v] = fast_exp(-s);
    }
    __syncthreads();
    for (int v = threadIdx.x; v < V; v += kThreads) {
        float a = 0.f;
        for (int d = 0; d < D; ++d) a += t.th[v * D + d];
        t.ssum[v] = a;
        t.prior[v] = p.prior ? p.prior[v] : 0.f;
    }
    __syncthreads();
}

// log p(z | class v) + prior_v for the classes owned by this lane; z_d lives in lane d.
__device__ __forceinline__ float class_score(const Tables& t, int V, int D, int v, float z_lane) {
    float acc = 0.f;
    for (int d = 0; d < D; ++d) {
        const float zd = __shfl_sync(0xffffffffu, z_lane, d);
        if (v < V) acc += softplus_pm(fmaf(zd, t.eT[d * V + v], -t.bT[d * V + v]) * (1.0f / kSigma));
    }
    return v < V ? -(acc + (float)D * kLogSigma) - t.ssum[v] + t.prior[v] : -INFINITY;
}

__global__ void __launch_bounds__(kThreads) categ_encode_kernel(const CategParams p) {
    extern __shared__ __align__(16) float sm[];
    const Tables t = carve(sm, p.V, p.D);
    load_tables(p, t);
    const int lane = threadIdx.x & 31, V = p.V, D = p.D;
    const long long nwarps = (long long)gridDim.x * kWarps;
    const long long wid = (long long)blockIdx.x * kWarps + (threadIdx.x >> 5);
    const long long chunk = (p.T + nwarps - 1) / nwarps;
    const long long t0 = wid * chunk, t1 = min(p.T, t0 + chunk);
    long long cur_b = -1;
    float run = 0.f;
    for (long long tk = t0; tk < t1; ++tk) {
        const long long tok = p.tokens[tk];
        const float padv = p.pad ? p.pad[tk] : 1.0f;
        const bool tok_ok = tok >= 0 && tok < V;
        const int x = tok_ok ? (int)tok : 0;
        // ---- forward: z0 ~ Logistic(0, sigma), z = (z0 + b_x) e^{s_x} -------------------------
        float z = 0.f, lp0 = 0.f, th = 0.f;
        if (lane < D) {
            float u;
            if (p.u) u = p.u[tk * D + lane];
            else {
                float r[4];
                const unsigned long long e = (unsigned long long)tk * D + lane;
                philox_uniform4(p.seed, p.offset + (e >> 2), r);
                u = r[e & 3];
            }
            const float z0 = __fmul_rn(logistic_from_uniform(u, kEps), kSigma);
            lp0 = -(softplus_pm(__fdiv_rn(z0, kSigma)) + kLogSigma);
            th = t.th[x * D + lane];
            z = (z0 + t.bv[x * D + lane]) * fast_exp(th);
        }
        const float init_log_p = warp_sum(lp0);
        const float ldj_fwd = warp_sum(th);
        const float log_point = init_log_p - ldj_fwd + t.prior[x];
        // ---- exact posterior over the V classes (linear_encoding.py:153-174) -------------------
        float m = -INFINITY, ssum = 0.f;
        for (int v0 = 0; v0 < V; v0 += 32) {
            const int v = v0 + lane;
            float sc = class_score(t, V, D, v, z);
            if (v == x) sc = log_point;  // own class: forward value (:167-168)
            const float mn = fmaxf(m, sc);
            if (mn > -INFINITY) ssum = ssum * fast_exp(m - mn) + (v < V ? fast_exp(sc - mn) : 0.f);
            m = mn;
        }
        float gm = m;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) gm = fmaxf(gm, __shfl_xor_sync(0xffffffffu, gm, d));
        const float part = (m > -INFINITY) ? ssum * fast_exp(m - gm) : 0.f;
        const float lse = gm + fast_log(warp_sum(part));
        const float cpl = log_point - lse;
        const float ldj_tok = (p.beta * cpl - (init_log_p - ldj_fwd)) * padv;
        if (lane < D) p.z_out[tk * D + lane] = z * padv;
        if (lane == 0) {
            if (p.cpl) p.cpl[tk] = cpl;
            uint32_t bad = 0u;
            if (ldj_tok != ldj_tok) bad |= CNF_FLAG_NAN_LDJ;
            if (!tok_ok) bad |= CNF_FLAG_CDF_RANGE;
            flag(p.status, bad);
            const long long b = tk / p.S;
            if (b != cur_b) {
                if (cur_b >= 0) atomicAdd(p.ldj + cur_b, run);
                cur_b = b; run = 0.f;
            }
            run += ldj_tok;
        }
    }
    if (lane == 0 && cur_b >= 0) atomicAdd(p.ldj + cur_b, run);
}

__global__ void __launch_bounds__(kThreads) categ_decode_kernel(const CategParams p) {
    extern __shared__ __align__(16) float sm[];
    const Tables t = carve(sm, p.V, p.D);
    load_tables(p, t);
    const int lane = threadIdx.x & 31, V = p.V, D = p.D;
    const long long nwarps = (long long)gridDim.x * kWarps;
    for (long long tk = (long long)blockIdx.x * kWarps + (threadIdx.x >> 5); tk < p.T; tk += nwarps) {
        const float z = lane < D ? p.z_in[tk * D + lane] : 0.f;
        float best = -INFINITY;
        int arg = 0x7fffffff;
        for (int v0 = 0; v0 < V; v0 += 32) {
            const int v = v0 + lane;
            const float sc = class_score(t, V, D, v, z);
            if (v < V && (sc > best || arg == 0x7fffffff)) { best = sc; arg = v; }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, d);
            const int oa = __shfl_xor_sync(0xffffffffu, arg, d);
            if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
        }
        if (lane == 0) p.tokens_out[tk] = arg;
    }
}

size_t table_smem(int V, int D) { return sizeof(float) * (4 * (size_t)V * D + 2 * (size_t)V); }

}  // namespace
}  // namespace cnf

using namespace cnf;

namespace cnf {
int categ_encode_tpt_try(const cnf_categ_encode_args* a, cudaStream_t stream, int* handled);
int categ_decode_tpt_try(const cnf_categ_decode_args* a, cudaStream_t stream, int* handled);
bool categ_encode_tpt_fusable(const cnf_categ_encode_args* a);
}

static int check_dims(int V, int D, size_t* smem) {
    CNF_REQUIRE(V >= 1 && D >= 1, "V and D must be >= 1");
    CNF_SUPPORTED(D <= 32, "D=%d > 32 latent dimensions per token", D);
    *smem = table_smem(V, D);
    CNF_SUPPORTED(*smem <= 200 * 1024, "class tables of V=%d x D=%d do not fit in shared memory", V, D);
    return CNF_OK;
}

extern "C" int cnf_categ_encode(const cnf_categ_encode_args* a, cnf_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    CNF_REQUIRE(a != nullptr, "args is NULL");
    CNF_REQUIRE(a->B >= 0 && a->S >= 0, "bad sizes");
    size_t smem;
    int rc = check_dims(a->V, a->D, &smem);
    if (rc != CNF_OK) return rc;
    const long long T = a->B * a->S;
    if (T == 0) return CNF_OK;
    CNF_REQUIRE(a->tokens && a->table && a->z_out && a->ldj, "tokens / table / z_out / ldj is NULL");
    {
        int handled = 0;
        rc = categ_encode_tpt_try(a, stream, &handled);
        if (rc != CNF_OK || handled) return rc;
    }
    CNF_SUPPORTED(a->next_actnorm_bias == nullptr && a->next_actnorm_scales == nullptr && a->next_conv_weight == nullptr,
                  "the fused first-block epilogue needs the thread-per-token encode kernel; query cnf_categ_encode_fusable");
    CategParams p{};
    p.tokens = reinterpret_cast<const long long*>(a->tokens); p.u = a->u_noise; p.table = a->table;
    p.prior = a->category_prior; p.pad = a->pad; p.z_out = a->z_out; p.ldj = a->ldj; p.cpl = a->class_prob_log;
    p.status = a->status; p.T = T; p.S = (int)a->S; p.V = a->V; p.D = a->D; p.beta = a->beta;
    p.seed = a->seed; p.offset = a->offset;
    if (smem > 48 * 1024)
        CNF_CUDA(cudaFuncSetAttribute(categ_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long blocks = (T + kWarps - 1) / kWarps;
    const long long cap = (long long)sm_count() * (smem > 24 * 1024 ? 2 : 6);
    if (blocks > cap) blocks = cap;
    categ_encode_kernel<<<(unsigned)blocks, kThreads, smem, stream>>>(p);
    return launch_status("categ_encode_kernel");
}

extern "C" int cnf_categ_encode_fusable(const cnf_categ_encode_args* a) {
    if (a == nullptr || a->V < 1 || a->D < 1) return 0;
    return cnf::categ_encode_tpt_fusable(a) ? 1 : 0;
}

extern "C" int cnf_categ_decode(const cnf_categ_decode_args* a, cnf_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    CNF_REQUIRE(a != nullptr, "args is NULL");
    CNF_REQUIRE(a->B >= 0 && a->S >= 0, "bad sizes");
    size_t smem;
    int rc = check_dims(a->V, a->D, &smem);
    if (rc != CNF_OK) return rc;
    const long long T = a->B * a->S;
    if (T == 0) return CNF_OK;
    CNF_REQUIRE(a->z && a->table && a->tokens_out, "z / table / tokens_out is NULL");
    {
        int handled = 0;
        rc = categ_decode_tpt_try(a, stream, &handled);
        if (rc != CNF_OK || handled) return rc;
    }
    CategParams p{};
    p.z_in = a->z; p.table = a->table; p.prior = a->category_prior;
    p.tokens_out = reinterpret_cast<long long*>(a->tokens_out); p.T = T; p.S = (int)a->S; p.V = a->V; p.D = a->D;
    if (smem > 48 * 1024)
        CNF_CUDA(cudaFuncSetAttribute(categ_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long blocks = (T + kWarps - 1) / kWarps;
    const long long cap = (long long)sm_count() * (smem > 24 * 1024 ? 2 : 6);
    if (blocks > cap) blocks = cap;
    categ_decode_kernel<<<(unsigned)blocks, kThreads, smem, stream>>>(p);
    return launch_status("categ_decode_kernel");
}
