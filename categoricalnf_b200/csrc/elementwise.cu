// K3 / K4 / K7: bandwidth-bound element-wise flow layers with fused per-sample ldj reduction.
//   affine coupling       layers/flows/coupling_layer.py:53-65, 76-98
//   ActNorm / ExtActNorm  layers/flows/activation_normalization.py:24-48, 116-144, 55-67
//   logistic prior        layers/flows/distributions.py:117-163
//   ldj accumulator       layers/flows/flow_model.py:44
// Each replaces 3-12 eager launches with one pass over the tensor: 16-byte loads/stores, grid a
// multiple of the SM count, one global atomic per (warp, sample) for the ldj.
#include "cnf_common.cuh"
#include "philox.cuh"

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

// --------------------------------------------------------------------------------------------
// affine coupling: one thread per (position, channel); nn_out record [s, t] is one 8-byte load
// --------------------------------------------------------------------------------------------
struct AffineParams {
    const float* z; const float2* nn; const float* sf;
    float* z_out; float* ldj; uint32_t* status;
    long long n;   // B*S*C
    long long SC;  // S*C elements per sample
    int S, C, reverse, pre;
    MaskView mask;
};

__global__ void __launch_bounds__(kThreads) affine_kernel(const AffineParams p) {
    const long long stride = (long long)gridDim.x * kThreads;
    const long long n_round = (p.n + 31) & ~31ll;
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n_round; i += stride) {
        const bool in = i < p.n;
        float contrib = 0.f;
        long long b = 0;
        if (in) {
            const long long pos = i / p.C;
            const int c = (int)(i - pos * p.C);
            b = i / p.SC;
            bool cond = (p.mask.cond_c >> c) & 1ull;
            if (p.mask.s_period > 0) cond = cond || ((p.mask.cond_s >> ((pos % p.S) % p.mask.s_period)) & 1ull);
            const float x = p.z[i];
            float out = x;
            if (!cond) {
                const float2 st = p.nn[i];
                float s = st.x;
                if (!p.pre) {
                    const float fac = p.sf ? expf(p.sf[c]) : 1.0f;
                    s = tanh_from_2log2e(st.x * (2.0f * kLog2e / fmaxf(fac, 1.0f))) * fac;
                }
                if (!p.reverse) { out = (x + st.y) * fast_exp(s); contrib = s; }
                else { out = fmaf(x, fast_exp(-s), -st.y); contrib = -s; }
                if (out != out) flag(p.status, CNF_FLAG_NAN_Z);
            }
            p.z_out[i] = out;
        }
        warp_segmented_atomic_add(p.ldj, b, contrib, in);
    }
}

// --------------------------------------------------------------------------------------------
// ActNorm: (z + b) e^{s} * pad, per-sample constant ldj term handled by the first B threads
// --------------------------------------------------------------------------------------------
struct ActNormParams {
    const float* z; const float* bias; const float* scales; const float* pad; const float* length;
    float* z_out; float* ldj; uint32_t* status;
    long long n, B;
    int S, C, reverse;
};

__global__ void __launch_bounds__(kThreads) actnorm_kernel(const ActNormParams p) {
    extern __shared__ float sm[];
    float* s_b = sm;          // [C]
    float* s_e = sm + p.C;    // [C] e^{+-s}
    float ssum = 0.f;
    for (int c = threadIdx.x; c < p.C; c += kThreads) {
        s_b[c] = p.bias[c];
        s_e[c] = expf(p.reverse ? -p.scales[c] : p.scales[c]);
    }
    __syncthreads();
    const long long gtid = (long long)blockIdx.x * kThreads + threadIdx.x;
    const long long stride = (long long)gridDim.x * kThreads;
    // ldj[b] += (+/-) sum_c s_c * len_b   (activation_normalization.py:27-40)
    if (p.ldj != nullptr) {
        for (long long b = gtid; b < p.B; b += stride) {
            if (ssum == 0.f) for (int c = 0; c < p.C; ++c) ssum += p.scales[c];
            float len;
            if (p.length) len = p.length[b];
            else if (p.pad) { len = 0.f; for (int s = 0; s < p.S; ++s) len += p.pad[b * p.S + s]; }
            else len = (float)p.S;
            const float v = p.ldj[b] + (p.reverse ? -ssum : ssum) * len;
            p.ldj[b] = v;
            if (v != v) flag(p.status, CNF_FLAG_NAN_LDJ);
        }
    }
    const int C = p.C;
    if ((C & 3) == 0 && ((reinterpret_cast<uintptr_t>(p.z) | reinterpret_cast<uintptr_t>(p.z_out)) & 15) == 0) {
        const long long n4 = p.n >> 2;
        const int c4n = C >> 2;
        for (long long i = gtid; i < n4; i += stride) {
            const long long pos = i / c4n;
            const int c = (int)(i - pos * c4n) << 2;
            float4 v = ldg_stream4(reinterpret_cast<const float4*>(p.z) + i);
            const float pv = p.pad ? p.pad[pos] : 1.0f;
            if (!p.reverse) {
                v.x = (v.x + s_b[c]) * s_e[c] * pv; v.y = (v.y + s_b[c + 1]) * s_e[c + 1] * pv;
                v.z = (v.z + s_b[c + 2]) * s_e[c + 2] * pv; v.w = (v.w + s_b[c + 3]) * s_e[c + 3] * pv;
            } else {
                v.x = fmaf(v.x, s_e[c], -s_b[c]) * pv; v.y = fmaf(v.y, s_e[c + 1], -s_b[c + 1]) * pv;
                v.z = fmaf(v.z, s_e[c + 2], -s_b[c + 2]) * pv; v.w = fmaf(v.w, s_e[c + 3], -s_b[c + 3]) * pv;
            }
            if (v.x != v.x || v.y != v.y || v.z != v.z || v.w != v.w) flag(p.status, CNF_FLAG_NAN_Z);
            stg_stream4(reinterpret_cast<float4*>(p.z_out) + i, v);
        }
    } else {
        for (long long i = gtid; i < p.n; i += stride) {
            const long long pos = i / C;
            const int c = (int)(i - pos * C);
            const float pv = p.pad ? p.pad[pos] : 1.0f;
            const float x = p.z[i];
            const float v = (!p.reverse ? (x + s_b[c]) * s_e[c] : fmaf(x, s_e[c], -s_b[c])) * pv;
            if (v != v) flag(p.status, CNF_FLAG_NAN_Z);
            p.z_out[i] = v;
        }
    }
}

// --------------------------------------------------------------------------------------------
// ExtActNorm: per-element (bias, raw scale) from the class embedding; ldj masked by pad
// --------------------------------------------------------------------------------------------
struct ExtParams {
    const float* z; const float* ext; const float* pad;
    float* z_out; float* ldj; uint32_t* status;
    long long n, SC;
    int C, reverse;
};

__global__ void __launch_bounds__(kThreads) ext_actnorm_kernel(const ExtParams p) {
    const long long stride = (long long)gridDim.x * kThreads;
    const long long n_round = (p.n + 31) & ~31ll;
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n_round; i += stride) {
        const bool in = i < p.n;
        float contrib = 0.f;
        long long b = 0;
        if (in) {
            const long long pos = i / p.C;
            const int c = (int)(i - pos * p.C);
            b = i / p.SC;
            const float bias = p.ext[pos * 2 * p.C + c];
            const float s = tanh_from_2log2e(p.ext[pos * 2 * p.C + p.C + c] * (2.0f * kLog2e));
            const float pv = p.pad ? p.pad[pos] : 1.0f;
            const float x = p.z[i];
            float out;
            if (!p.reverse) { out = (x + bias) * fast_exp(s); contrib = s * pv; }
            else { out = fmaf(x, fast_exp(-s), -bias); contrib = -s * pv; }
            if (out != out) flag(p.status, CNF_FLAG_NAN_Z);
            p.z_out[i] = out;
        }
        warp_segmented_atomic_add(p.ldj, b, contrib, in);
    }
}

// --------------------------------------------------------------------------------------------
// ActNorm data-dependent init: masked per-channel sum, sum of squares and count in float64
// --------------------------------------------------------------------------------------------
struct InitParams {
    const float* x; const float* pad; double* ws; float* bias; float* scales;
    long long P; int C;
};

__global__ void __launch_bounds__(kThreads) actnorm_stats_kernel(const InitParams p) {
    // thread t handles channel t % C of positions t / C, t / C + step, ... (C <= 256 assumed by launcher)
    const int C = p.C;
    const int lanes_pos = kThreads / C;
    const int c = threadIdx.x % C, slot = threadIdx.x / C;
    if (slot >= lanes_pos) return;
    double s1 = 0.0, s2 = 0.0, cnt = 0.0;
    for (long long pos = (long long)blockIdx.x * lanes_pos + slot; pos < p.P; pos += (long long)gridDim.x * lanes_pos) {
        const double m = p.pad ? (double)p.pad[pos] : 1.0;
        const double v = (double)p.x[pos * C + c];
        s1 += v * m; s2 += v * v * m; cnt += m;
    }
    atomicAdd(p.ws + c, s1);
    atomicAdd(p.ws + C + c, s2);
    atomicAdd(p.ws + 2 * C + c, cnt);
}

__global__ void actnorm_finalize_kernel(const InitParams p) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= p.C) return;
    const double n = p.ws[2 * p.C + c];
    const double mean = p.ws[c] / n;
    const double var = p.ws[p.C + c] / n - mean * mean;  // E[(x - mean)^2]
    p.bias[c] = (float)(-mean);
    p.scales[c] = (float)(-0.5 * log(var));
}

// --------------------------------------------------------------------------------------------
// logistic prior
// --------------------------------------------------------------------------------------------
struct LogProbParams {
    const float* x; const float* pad; float* out; float* elem;
    const float* add; double* total;
    long long n, SC, B; int C;
    float mu, inv_sigma, log_sigma;
};

// Block-level tail shared by both log-prob kernels: the block's partial of sum_b out[b] goes to total[0] with ONE double
// atomic per block (total[1] = number of samples, written once).  This is the (sum log-likelihood, count) pair the ranks
// all-reduce once per step - produced by the kernel that finishes the log-likelihood instead of by a separate reduction.
__device__ __forceinline__ void block_total_add(const LogProbParams& p, float part) {
    if (p.total == nullptr) return;
    __shared__ float s_part[kThreads / 32];
    part = warp_sum(part);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) t += (double)s_part[w];
        if (t != 0.0) atomicAdd(p.total, t);
        if (blockIdx.x == 0) p.total[1] = (double)p.B;
    }
}

__global__ void __launch_bounds__(kThreads) logistic_logprob_kernel(const LogProbParams p) {
    const long long stride = (long long)gridDim.x * kThreads;
    const long long n_round = (p.n + 31) & ~31ll;
    float part = 0.f;
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n_round; i += stride) {
        const bool in = i < p.n;
        float lp = 0.f;
        long long b = 0;
        if (in) {
            b = i / p.SC;
            lp = -(softplus_pm((p.x[i] - p.mu) * p.inv_sigma) + p.log_sigma);
            if (p.elem) p.elem[i] = lp;
            if (p.pad) lp *= p.pad[i / p.C];
            if (p.add && i == b * p.SC) lp += p.add[b];      // first element of the sample carries add[b] into out[b]
        }
        part += lp;
        if (p.out) warp_segmented_atomic_add(p.out, b, lp, in);
    }
    block_total_add(p, part);
}

// Per-sample sums only, S*C a multiple of 1024: a warp owns 1024 consecutive elements of ONE sample per step (eight
// 16-byte loads per lane in flight), so the reduction is one warp sum and one atomic per 4 KB of input - the kernel runs at
// HBM speed instead of at the issue rate of a segmented scan per element.
__global__ void __launch_bounds__(kThreads) logistic_logprob_rows_kernel(const LogProbParams p) {
    const int lane = threadIdx.x & 31;
    const long long nchunks = p.n >> 10;
    const long long nwarps = (long long)gridDim.x * (kThreads / 32);
    const int c4 = p.C >> 2;   // p.C % 4 == 0 on this path: a 16-byte load never straddles two positions
    float part = 0.f;          // lane 0: this warp's share of sum_b out[b]
    for (long long w = (long long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); w < nchunks; w += nwarps) {
        const long long base = w << 10;
        float4 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __ldcs(reinterpret_cast<const float4*>(p.x + base) + j * 32 + lane);
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float s = softplus_pm((v[j].x - p.mu) * p.inv_sigma) + softplus_pm((v[j].y - p.mu) * p.inv_sigma) +
                      softplus_pm((v[j].z - p.mu) * p.inv_sigma) + softplus_pm((v[j].w - p.mu) * p.inv_sigma) + 4.0f * p.log_sigma;
            if (p.pad) s *= p.pad[((base >> 2) + j * 32 + lane) / c4];
            acc -= s;
        }
        acc = warp_sum(acc);
        if (lane == 0) {
            const long long b = base / p.SC;
            if (p.add && base == b * p.SC) acc += p.add[b];      // the sample's first chunk carries add[b] into out[b]
            atomicAdd(p.out + b, acc);
            part += acc;
        }
    }
    if (p.total != nullptr) block_total_add(p, (lane == 0) ? part : 0.f);
}

struct SampleParams {
    const float* u; float* x; long long n;
    unsigned long long seed, offset;
    float mu, sigma, eps;
};

__global__ void __launch_bounds__(kThreads) logistic_sample_kernel(const SampleParams p) {
    const long long stride = (long long)gridDim.x * kThreads;
    const long long n4 = (p.n + 3) >> 2;
    for (long long q = (long long)blockIdx.x * kThreads + threadIdx.x; q < n4; q += stride) {
        float u[4];
        if (p.u == nullptr) philox_uniform4(p.seed, p.offset + (unsigned long long)q, u);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const long long i = 4 * q + j;
            if (i >= p.n) break;
            const float uu = p.u ? p.u[i] : u[j];
            p.x[i] = logistic_from_uniform(uu, p.eps) * p.sigma + p.mu;
        }
    }
}

struct AxpyParams {
    const float* alpha_dev; const float* x; const float* length; float* y; long long B; float alpha;
};

__global__ void ldj_axpy_kernel(const AxpyParams p) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= p.B) return;
    float a = p.alpha;
    if (p.alpha_dev) a *= p.alpha_dev[0];
    if (p.x) a *= p.x[b];
    if (p.length) a *= p.length[b];
    p.y[b] += a;
}

}  // namespace
}  // namespace cnf

using namespace cnf;

extern "C" int cnf_affine_coupling(const cnf_affine_args* a, cnf_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    CNF_REQUIRE(a != nullptr, "args is NULL");
    CNF_REQUIRE(a->B >= 0 && a->S >= 0 && a->C >= 1, "bad sizes");
    AffineParams p{};
    int rc = build_mask(a->mask, a->C, &p.mask);
    if (rc != CNF_OK) return rc;
    p.n = a->B * a->S * a->C;
    if (p.n == 0) return CNF_OK;
    CNF_REQUIRE(a->z && a->nn_out && a->z_out && a->ldj, "z / nn_out / z_out / ldj is NULL");
    CNF_REQUIRE((reinterpret_cast<uintptr_t>(a->nn_out) & 7) == 0, "nn_out must be 8-byte aligned");
    p.z = a->z; p.nn = reinterpret_cast<const float2*>(a->nn_out); p.sf = a->scaling_factor;
    p.z_out = a->z_out; p.ldj = a->ldj; p.status = a->status;
    p.SC = a->S * a->C; p.S = (int)a->S; p.C = a->C; p.reverse = a->reverse; p.pre = a->params_prebounded;
    affine_kernel<<<grid_for(p.n), kThreads, 0, stream>>>(p);
    return launch_status("affine_kernel");
}

extern "C" int cnf_actnorm(const cnf_actnorm_args* a, cnf_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    CNF_REQUIRE(a != nullptr, "args is NULL");
    CNF_REQUIRE(a->B >= 0 && a->S >= 0 && a->C >= 1, "bad sizes");
    CNF_SUPPORTED(a->C <= 4096, "C=%d too large", a->C);
    ActNormParams p{};
    p.n = a->B * a->S * a->C; p.B = a->B;
    if (a->B == 0) return CNF_OK;
    CNF_REQUIRE(a->bias && a->scales, "bias / scales is NULL");
    CNF_REQUIRE(p.n == 0 || (a->z && a->z_out), "z / z_out is NULL");
    p.z = a->z; p.bias = a->bias; p.scales = a->scales; p.pad = a->pad; p.length = a->length;
    p.z_out = a->z_out; p.ldj = a->ldj; p.status = a->status;
    p.S = (int)a->S; p.C = a->C; p.reverse = a->reverse;
    const long long work = ((a->C & 3) == 0) ? (p.n >> 2) : p.n;
    actnorm_kernel<<<grid_for(work > a->B ? work : a->B), kThreads, 2 * a->C * sizeof(float), stream>>>(p);
    return launch_status("actnorm_kernel");
}

extern "C" int cnf_ext_actnorm(const cnf_ext_actnorm_args* a, cnf_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    CNF_REQUIRE(a != nullptr, "args is NULL");
    CNF_REQUIRE(a->B >= 0 && a->S >= 0 && a->C >= 1, "bad sizes");
    ExtParams p{};
    p.n = a->B * a->S * a->C;
    if (p.n == 0) return CNF_OK;
    CNF_REQUIRE(a->z && a->ext && a->z_out && a->ldj, "z / ext / z_out / ldj is NULL");
    p.z = a->z; p.ext = a->ext; p.pad = a->pad; p.z_out = a->z_out; p.ldj = a->ldj; p.status = a->status;
    p.SC = a->S * a->C; p.C = a->C; p.reverse = a->reverse;
    ext_actnorm_kernel<<<grid_for(p.n), kThreads, 0, stream>>>(p);
    return launch_status("ext_actnorm_kernel");
}

extern "C" int cnf_actnorm_data_init(const cnf_actnorm_init_args* a, cnf_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    CNF_REQUIRE(a != nullptr, "args is NULL");
    CNF_REQUIRE(a->B >= 1 && a->S >= 1 && a->C >= 1, "bad sizes");
    CNF_SUPPORTED(a->C <= kThreads, "C=%d > %d", a->C, kThreads);
    CNF_REQUIRE(a->x && a->workspace && a->bias && a->scales, "x / workspace / bias / scales is NULL");
    InitParams p{a->x, a->pad, a->workspace, a->bias, a->scales, a->B * a->S, a->C};
    CNF_CUDA(cudaMemsetAsync(a->workspace, 0, 3 * sizeof(double) * (size_t)a->C, stream));
    const int lanes_pos = kThreads / a->C;
    actnorm_stats_kernel<<<grid_for((p.P + lanes_pos - 1) / lanes_pos * kThreads, 4), kThreads, 0, stream>>>(p);
    actnorm_finalize_kernel<<<(a->C + 127) / 128, 128, 0, stream>>>(p);
    return launch_status("actnorm_data_init");
}

extern "C" int cnf_logistic_logprob(const cnf_logistic_logprob_args* a, cnf_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    CNF_REQUIRE(a != nullptr, "args is NULL");
    CNF_REQUIRE(a->B >= 0 && a->S >= 0 && a->C >= 1 && a->sigma > 0.f, "bad sizes / sigma");
    if (a->out && !a->accumulate) CNF_CUDA(cudaMemsetAsync(a->out, 0, sizeof(float) * (size_t)a->B, stream));
    if (a->total) CNF_CUDA(cudaMemsetAsync(a->total, 0, 2 * sizeof(double), stream));
    CNF_REQUIRE(!(a->add || a->total) || (a->out && !a->accumulate), "add / total need `out` with accumulate = 0");
    CNF_REQUIRE(!a->add || a->add != a->out, "add must not alias out");
    LogProbParams p{};
    p.n = a->B * a->S * a->C;
    if (p.n == 0) return CNF_OK;
    CNF_REQUIRE(a->x && (a->out || a->elementwise), "x is NULL or no output requested");
    p.x = a->x; p.pad = a->pad; p.out = a->out; p.elem = a->elementwise; p.add = a->add; p.total = a->total; p.B = a->B;
    p.SC = a->S * a->C; p.C = a->C; p.mu = a->mu; p.inv_sigma = 1.0f / a->sigma; p.log_sigma = logf(a->sigma);
    if (p.out != nullptr && p.elem == nullptr && (p.SC & 1023) == 0 && (p.C & 3) == 0 &&
        (reinterpret_cast<uintptr_t>(p.x) & 15) == 0) {
        logistic_logprob_rows_kernel<<<grid_for((p.n >> 10) * 32, 8), kThreads, 0, stream>>>(p);
        return launch_status("logistic_logprob_rows_kernel");
    }
    logistic_logprob_kernel<<<grid_for(p.n), kThreads, 0, stream>>>(p);
    return launch_status("logistic_logprob_kernel");
}

extern "C" int cnf_logistic_sample(const cnf_logistic_sample_args* a, cnf_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    CNF_REQUIRE(a != nullptr && a->n >= 0, "bad args");
    if (a->n == 0) return CNF_OK;
    CNF_REQUIRE(a->x_out != nullptr, "x_out is NULL");
    SampleParams p{a->u_noise, a->x_out, a->n, a->seed, a->offset, a->mu, a->sigma, a->eps};
    logistic_sample_kernel<<<grid_for((a->n + 3) / 4), kThreads, 0, stream>>>(p);
    return launch_status("logistic_sample_kernel");
}

extern "C" int cnf_ldj_axpy(const cnf_ldj_axpy_args* a, cnf_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    CNF_REQUIRE(a != nullptr && a->B >= 0, "bad args");
    if (a->B == 0) return CNF_OK;
    CNF_REQUIRE(a->y != nullptr, "y is NULL");
    AxpyParams p{a->alpha_dev, a->x, a->length, a->y, a->B, a->alpha};
    ldj_axpy_kernel<<<(unsigned)((a->B + 255) / 256), 256, 0, stream>>>(p);
    return launch_status("ldj_axpy_kernel");
}
