// SigmoidFlow and the two ends of VariationalDequantization.
//   layers/flows/sigmoid_layer.py:24-48
//   layers/categorical_encoding/variational_dequantization.py:31-58
// Element-wise, HBM-bound: one pass, per-sample ldj reduced in the warp before one atomic per
// (warp, sample); the dequantised value z.float() + noise is formed in the same pass.
// The logit direction reproduces the reference's fp32 rounding sequence (mul, add - no fma
// contraction): close to y = 1 the value 1 - y is exact given y, so the result is decided by how y
// itself was rounded.
#include "cnf_common.cuh"

namespace cnf {
namespace {

constexpr int kThreads = 256;

inline unsigned grid_for(long long work_items, int per_sm = 8) {
    long long blocks = (work_items + kThreads - 1) / kThreads;
    long long cap = (long long)sm_count() * per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (unsigned)blocks;
}

struct SigParams {
    const float* z; const long long* add; float* z_out; float* ldj; float* elem; uint32_t* status;
    long long n, per;
    int reverse;
    float one_m_alpha, half_alpha, log_one_m_alpha;
};

__device__ __forceinline__ float squeeze_unit(float z, float one_m_alpha, float half_alpha) {
    return __fadd_rn(__fmul_rn(z, one_m_alpha), half_alpha);   // sigmoid_layer.py:35
}

__global__ void __launch_bounds__(kThreads) sigmoid_flow_kernel(const SigParams p) {
    const long long stride = (long long)gridDim.x * kThreads;
    const long long n_round = (p.n + 31) & ~31ll;
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n_round; i += stride) {
        const bool in = i < p.n;
        float l = 0.f;
        long long b = 0;
        if (in) {
            b = i / p.per;
            const float z = p.z[i];
            float out;
            if (!p.reverse) {
                // sigmoid: log sigma'(z) = -z - 2 softplus(-z) = -|z| - 2 log(1 + e^{-|z|})   (:32-33)
                const float a = fabsf(z);
                const float e = expf(-a);
                l = -a - 2.0f * log1pf(e);
                const float r = 1.0f / (1.0f + e);
                out = z >= 0.f ? r : e * r;
            } else {
                const float y = squeeze_unit(z, p.one_m_alpha, p.half_alpha);
                const float ly = logf(y), l1y = logf(1.0f - y);
                l = -ly - l1y + p.log_one_m_alpha;                                          // :36
                out = ly - l1y;                                                             // :37
            }
            uint32_t bits = 0u;
            if (out != out) bits |= CNF_FLAG_NAN_Z;
            if (l != l) bits |= CNF_FLAG_NAN_LDJ;
            flag(p.status, bits);
            if (p.elem) p.elem[i] = l;
            if (p.add) out += (float)p.add[i];
            p.z_out[i] = out;
        }
        if (p.ldj) warp_segmented_atomic_add(p.ldj, b, l, in);
    }
}

struct SigBwdParams {
    const float* z; const float* g_out; const float* g_ldj; const float* g_elem; float* g_z;
    long long n, per;
    int reverse;
    float one_m_alpha, half_alpha;
};

__global__ void __launch_bounds__(kThreads) sigmoid_flow_bwd_kernel(const SigBwdParams p) {
    const long long stride = (long long)gridDim.x * kThreads;
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < p.n; i += stride) {
        const float z = p.z[i];
        float gl = p.g_ldj ? p.g_ldj[i / p.per] : 0.f;
        if (p.g_elem) gl += p.g_elem[i];
        const float go = p.g_out ? p.g_out[i] : 0.f;
        float dz, dl;
        if (!p.reverse) {
            const float a = fabsf(z);
            const float e = expf(-a);
            const float r = 1.0f / (1.0f + e);
            const float s = z >= 0.f ? r : e * r;          // sigmoid(z)
            dz = e * r * r;                                // sigma (1 - sigma)
            dl = 1.0f - 2.0f * s;                          // d/dz (-z - 2 softplus(-z))
        } else {
            const float y = squeeze_unit(z, p.one_m_alpha, p.half_alpha);
            const float inv = p.one_m_alpha / (y * (1.0f - y));
            dz = inv;                                      // d logit(y) / dz
            dl = (2.0f * y - 1.0f) * inv;                  // d(-log y - log(1-y)) / dz
        }
        p.g_z[i] = go * dz + gl * dl;
    }
}

struct FloorParams { const float* z; long long* out; long long n; int V; };

__global__ void __launch_bounds__(kThreads) dequant_floor_kernel(const FloorParams p) {
    const long long stride = (long long)gridDim.x * kThreads;
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < p.n; i += stride) {
        float f = floorf(p.z[i]);
        f = fminf(fmaxf(f, 0.0f), (float)(p.V - 1));     // clamp(min=0, max=V-1) (:55); NaN -> 0
        p.out[i] = (long long)f;
    }
}

}  // namespace
}  // namespace cnf

using namespace cnf;

extern "C" int cnf_sigmoid_flow(const cnf_sigmoid_flow_args* a, cnf_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    CNF_REQUIRE(a != nullptr, "args is NULL");
    CNF_REQUIRE(a->B >= 0 && a->n_per_sample >= 0, "bad sizes");
    CNF_REQUIRE(a->alpha >= 0.f && a->alpha < 1.f, "alpha must be in [0,1)");
    if (a->ldj && !a->accumulate && a->B > 0) CNF_CUDA(cudaMemsetAsync(a->ldj, 0, sizeof(float) * (size_t)a->B, stream));
    SigParams p{};
    p.n = a->B * a->n_per_sample;
    if (p.n == 0) return CNF_OK;
    CNF_REQUIRE(a->z && a->z_out, "z / z_out is NULL");
    p.z = a->z; p.add = reinterpret_cast<const long long*>(a->add_tokens); p.z_out = a->z_out;
    p.ldj = a->ldj; p.elem = a->ldj_elementwise; p.status = a->status;
    p.per = a->n_per_sample; p.reverse = a->reverse ? 1 : 0;
    // the reference multiplies an fp32 tensor by the Python scalars (1 - alpha) and alpha * 0.5: both are rounded to fp32
    p.one_m_alpha = (float)(1.0 - (double)a->alpha);
    p.half_alpha = (float)((double)a->alpha * 0.5);
    p.log_one_m_alpha = (float)log(1.0 - (double)a->alpha);
    sigmoid_flow_kernel<<<grid_for(p.n), kThreads, 0, stream>>>(p);
    return launch_status("sigmoid_flow_kernel");
}

extern "C" int cnf_sigmoid_flow_bwd(const cnf_sigmoid_flow_bwd_args* a, cnf_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    CNF_REQUIRE(a != nullptr, "args is NULL");
    CNF_REQUIRE(a->B >= 0 && a->n_per_sample >= 0, "bad sizes");
    SigBwdParams p{};
    p.n = a->B * a->n_per_sample;
    if (p.n == 0) return CNF_OK;
    CNF_REQUIRE(a->z && a->grad_z, "z / grad_z is NULL");
    p.z = a->z; p.g_out = a->grad_z_out; p.g_ldj = a->grad_ldj; p.g_elem = a->grad_ldj_elementwise; p.g_z = a->grad_z;
    p.per = a->n_per_sample; p.reverse = a->reverse ? 1 : 0;
    p.one_m_alpha = (float)(1.0 - (double)a->alpha);
    p.half_alpha = (float)((double)a->alpha * 0.5);
    sigmoid_flow_bwd_kernel<<<grid_for(p.n), kThreads, 0, stream>>>(p);
    return launch_status("sigmoid_flow_bwd_kernel");
}

extern "C" int cnf_dequant_floor(const cnf_dequant_floor_args* a, cnf_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    CNF_REQUIRE(a != nullptr && a->n >= 0 && a->V >= 1, "bad args");
    if (a->n == 0) return CNF_OK;
    CNF_REQUIRE(a->z && a->tokens_out, "z / tokens_out is NULL");
    FloorParams p{a->z, reinterpret_cast<long long*>(a->tokens_out), a->n, a->V};
    dequant_floor_kernel<<<grid_for(p.n), kThreads, 0, stream>>>(p);
    return launch_status("dequant_floor_kernel");
}
