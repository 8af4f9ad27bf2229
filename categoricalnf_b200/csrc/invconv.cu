// K5: invertible 1x1 convolution (reference layers/flows/permutation_layers.py:61-136).
//   build: W from the LU factors, sum(log_s) (or log|det W| by in-kernel LU for the direct
//          parametrisation), and W^-1 by float64 Gauss-Jordan with partial pivoting - one CTA.
//   apply: z @ W per position.  At C = 16 this is 4 flop/byte, far below the B200 ridge, so it
//          is bandwidth bound: one thread owns one position, keeps its row in registers (16-byte
//          loads) and streams W from shared memory as broadcast 16-byte reads.
#include "cnf_common.cuh"

namespace cnf {

int invconv_tile_try(const cnf_invconv_args* a, cudaStream_t stream, int* handled);      // invconv_tile.cu (C = 16, TMA tiles)

namespace {

constexpr int kThreads = 256;

struct BuildParams {
    const float *p, *l, *u, *log_s, *sign_s, *weight;
    float *w_out, *w_inv_out, *sldj_out;
    int C;
};

// dynamic smem: float lo[C*C], up[C*C], lu[C*C], w[C*C]; double aug[C*2C]
__global__ void __launch_bounds__(kThreads) invconv_build_kernel(const BuildParams q) {
    extern __shared__ __align__(16) unsigned char raw[];
    const int C = q.C, CC = C * C, tid = threadIdx.x;
    float* lo = reinterpret_cast<float*>(raw);
    float* up = lo + CC;
    float* lu = up + CC;
    float* w = lu + CC;
    double* aug = reinterpret_cast<double*>(w + ((CC + 1) & ~1));
    __shared__ int s_piv;
    __shared__ double s_logdet;

    if (q.weight == nullptr) {
        for (int i = tid; i < CC; i += kThreads) {
            const int r = i / C, c = i - r * C;
            lo[i] = (r > c ? q.l[i] : 0.f) + (r == c ? 1.f : 0.f);
            up[i] = (r < c ? q.u[i] : 0.f) + (r == c ? q.sign_s[r] * expf(q.log_s[r]) : 0.f);
        }
        __syncthreads();
        for (int i = tid; i < CC; i += kThreads) {
            const int r = i / C, c = i - r * C;
            float acc = 0.f;
            for (int k = 0; k < C; ++k) acc = fmaf(lo[r * C + k], up[k * C + c], acc);
            lu[i] = acc;
        }
        __syncthreads();
        for (int i = tid; i < CC; i += kThreads) {
            const int r = i / C, c = i - r * C;
            float acc = 0.f;
            for (int k = 0; k < C; ++k) acc = fmaf(q.p[r * C + k], lu[k * C + c], acc);
            w[i] = acc;
        }
    } else {
        for (int i = tid; i < CC; i += kThreads) w[i] = q.weight[i];
    }
    __syncthreads();
    for (int i = tid; i < CC; i += kThreads) q.w_out[i] = w[i];

    // [W | I] -> [I | W^-1] in float64, partial pivoting; log|det| = sum log|pivot|
    const int W2 = 2 * C;
    for (int i = tid; i < C * W2; i += kThreads) {
        const int r = i / W2, c = i - r * W2;
        aug[i] = c < C ? (double)w[r * C + c] : (c - C == r ? 1.0 : 0.0);
    }
    if (tid == 0) s_logdet = 0.0;
    __syncthreads();
    for (int col = 0; col < C; ++col) {
        if (tid == 0) {
            int best = col;
            double bv = fabs(aug[col * W2 + col]);
            for (int r = col + 1; r < C; ++r) {
                const double v = fabs(aug[r * W2 + col]);
                if (v > bv) { bv = v; best = r; }
            }
            s_piv = best;
            s_logdet += log(bv);
        }
        __syncthreads();
        const int pr = s_piv;
        if (pr != col) {
            for (int c = tid; c < W2; c += kThreads) {
                const double t = aug[col * W2 + c];
                aug[col * W2 + c] = aug[pr * W2 + c];
                aug[pr * W2 + c] = t;
            }
        }
        __syncthreads();
        const double inv_p = 1.0 / aug[col * W2 + col];
        __syncthreads();
        for (int c = tid; c < W2; c += kThreads) aug[col * W2 + c] *= inv_p;
        __syncthreads();
        for (int i = tid; i < C * W2; i += kThreads) {
            const int r = i / W2, c = i - r * W2;
            if (r != col && c != col) aug[i] -= aug[r * W2 + col] * aug[col * W2 + c];
        }
        __syncthreads();
        for (int r = tid; r < C; r += kThreads)
            if (r != col) aug[r * W2 + col] = 0.0;
        __syncthreads();
    }
    if (q.w_inv_out)
        for (int i = tid; i < CC; i += kThreads) {
            const int r = i / C, c = i - r * C;
            q.w_inv_out[i] = (float)aug[r * W2 + C + c];
        }
    if (tid == 0) {
        if (q.weight == nullptr) {
            float s = 0.f;
            for (int c = 0; c < C; ++c) s += q.log_s[c];
            q.sldj_out[0] = s;
        } else {
            q.sldj_out[0] = (float)s_logdet;
        }
    }
}

struct ApplyParams {
    const float *z, *w, *sldj, *pad, *length;
    float *z_out, *ldj;
    uint32_t* status;
    long long P, B;
    int S, C, reverse;
    // optional: ActNorm of the same block applied first, a = (z + b) e^{s} pad (activation_normalization.py:35-43),
    // and a second output y * out_mask = the masked network input of the coupling layer that follows
    const float *pre_b, *pre_s, *omask;
    float* z_masked;
};

__device__ __forceinline__ void ldj_term(const ApplyParams& p, long long gtid, long long stride) {
    if (p.ldj == nullptr) return;
    for (long long b = gtid; b < p.B; b += stride) {
        const float len = p.length ? p.length[b] : (float)p.S;
        const float t = p.sldj[0] * len;
        const float v = p.reverse ? p.ldj[b] - t : p.ldj[b] + t;
        p.ldj[b] = v;
        if (v != v) flag(p.status, CNF_FLAG_NAN_LDJ);
    }
}

// register-tiled path: one thread per position, C in {4, 8, 16, 32}
template <int C>
__global__ void __launch_bounds__(kThreads) invconv_rows_kernel(const ApplyParams p) {
    __shared__ __align__(16) float s_w[C * C];
    __shared__ __align__(16) float s_pb[C], s_pe[C], s_om[C];
    for (int i = threadIdx.x; i < C * C; i += kThreads) s_w[i] = p.w[i];
    for (int i = threadIdx.x; i < C; i += kThreads) {
        s_pb[i] = p.pre_b ? p.pre_b[i] : 0.f;
        s_pe[i] = p.pre_s ? expf(p.pre_s[i]) : 1.0f;
        s_om[i] = p.omask ? p.omask[i] : 1.0f;
    }
    __syncthreads();
    const bool pre = p.pre_b != nullptr || p.pre_s != nullptr;
    const long long gtid = (long long)blockIdx.x * kThreads + threadIdx.x;
    const long long stride = (long long)gridDim.x * kThreads;
    ldj_term(p, gtid, stride);
    for (long long pos = gtid; pos < p.P; pos += stride) {
        float x[C], y[C];
        const float pv = p.pad ? p.pad[pos] : 1.0f;
        const float4* src = reinterpret_cast<const float4*>(p.z + pos * C);
#pragma unroll
        for (int j = 0; j < C / 4; ++j) {
            const float4 v = ldg_stream4(src + j);
            x[4 * j] = v.x; x[4 * j + 1] = v.y; x[4 * j + 2] = v.z; x[4 * j + 3] = v.w;
        }
        if (pre) {
#pragma unroll
            for (int c = 0; c < C; ++c) x[c] = (x[c] + s_pb[c]) * s_pe[c] * pv;
        }
#pragma unroll
        for (int co = 0; co < C; ++co) y[co] = 0.f;
#pragma unroll
        for (int ci = 0; ci < C; ++ci) {
#pragma unroll
            for (int j = 0; j < C / 4; ++j) {
                const float4 wv = *reinterpret_cast<const float4*>(s_w + ci * C + 4 * j);
                y[4 * j] = fmaf(x[ci], wv.x, y[4 * j]);
                y[4 * j + 1] = fmaf(x[ci], wv.y, y[4 * j + 1]);
                y[4 * j + 2] = fmaf(x[ci], wv.z, y[4 * j + 2]);
                y[4 * j + 3] = fmaf(x[ci], wv.w, y[4 * j + 3]);
            }
        }
        float4* dst = reinterpret_cast<float4*>(p.z_out + pos * C);
        float4* dstm = p.z_masked ? reinterpret_cast<float4*>(p.z_masked + pos * C) : nullptr;
        bool bad = false;
#pragma unroll
        for (int j = 0; j < C / 4; ++j) {
            float4 v = make_float4(y[4 * j] * pv, y[4 * j + 1] * pv, y[4 * j + 2] * pv, y[4 * j + 3] * pv);
            bad = bad || v.x != v.x || v.y != v.y || v.z != v.z || v.w != v.w;
            stg_stream4(dst + j, v);
            if (dstm != nullptr)
                dstm[j] = make_float4(v.x * s_om[4 * j], v.y * s_om[4 * j + 1], v.z * s_om[4 * j + 2], v.w * s_om[4 * j + 3]);
        }
        if (bad) flag(p.status, CNF_FLAG_NAN_Z);
    }
}

// generic path: one thread per (position, output channel)
__global__ void __launch_bounds__(kThreads) invconv_generic_kernel(const ApplyParams p) {
    extern __shared__ float s_w[];
    const int C = p.C;
    for (int i = threadIdx.x; i < C * C; i += kThreads) s_w[i] = p.w[i];
    __syncthreads();
    const long long gtid = (long long)blockIdx.x * kThreads + threadIdx.x;
    const long long stride = (long long)gridDim.x * kThreads;
    ldj_term(p, gtid, stride);
    const long long n = p.P * C;
    for (long long i = gtid; i < n; i += stride) {
        const long long pos = i / C;
        const int co = (int)(i - pos * C);
        const float* row = p.z + pos * C;
        const float pv = p.pad ? p.pad[pos] : 1.0f;
        float acc = 0.f;
        if (p.pre_b != nullptr || p.pre_s != nullptr) {
            for (int ci = 0; ci < C; ++ci) {
                const float a = (row[ci] + (p.pre_b ? p.pre_b[ci] : 0.f)) * (p.pre_s ? __expf(p.pre_s[ci]) : 1.0f) * pv;
                acc = fmaf(a, s_w[ci * C + co], acc);
            }
        } else {
            for (int ci = 0; ci < C; ++ci) acc = fmaf(row[ci], s_w[ci * C + co], acc);
        }
        acc *= pv;
        if (acc != acc) flag(p.status, CNF_FLAG_NAN_Z);
        p.z_out[i] = acc;
        if (p.z_masked) p.z_masked[i] = acc * (p.omask ? p.omask[co] : 1.0f);
    }
}

}  // namespace
}  // namespace cnf

using namespace cnf;

extern "C" int cnf_invconv_build(const cnf_invconv_build_args* a, cnf_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    CNF_REQUIRE(a != nullptr, "args is NULL");
    CNF_SUPPORTED(a->C >= 1 && a->C <= CNF_MAX_CHANNELS, "C=%d outside [1, %d]", a->C, CNF_MAX_CHANNELS);
    CNF_REQUIRE(a->w_out && a->sldj_out, "w_out / sldj_out is NULL");
    CNF_REQUIRE(a->weight || (a->p && a->l && a->u && a->log_s && a->sign_s), "LU factors incomplete");
    BuildParams q{a->p, a->l, a->u, a->log_s, a->sign_s, a->weight, a->w_out, a->w_inv_out, a->sldj_out, a->C};
    const int CC = a->C * a->C;
    const size_t smem = sizeof(float) * (3 * (size_t)CC + ((CC + 1) & ~1)) + sizeof(double) * 2 * (size_t)CC;
    if (smem > 48 * 1024)
        CNF_CUDA(cudaFuncSetAttribute(invconv_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    invconv_build_kernel<<<1, kThreads, smem, stream>>>(q);
    return launch_status("invconv_build_kernel");
}

extern "C" int cnf_invconv_apply(const cnf_invconv_args* a, cnf_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    CNF_REQUIRE(a != nullptr, "args is NULL");
    CNF_REQUIRE(a->B >= 0 && a->S >= 0, "bad sizes");
    CNF_SUPPORTED(a->C >= 1 && a->C <= CNF_MAX_CHANNELS, "C=%d outside [1, %d]", a->C, CNF_MAX_CHANNELS);
    if (a->B == 0) return CNF_OK;
    CNF_REQUIRE(a->weight && a->sldj, "weight / sldj is NULL");
    ApplyParams p{a->z, a->weight, a->sldj, a->pad, a->length, a->z_out, a->ldj, a->status,
                  a->B * a->S, a->B, (int)a->S, a->C, a->reverse,
                  a->pre_actnorm_bias, a->pre_actnorm_scales, a->out_mask, a->z_masked_out};
    CNF_SUPPORTED(!(a->reverse && (a->pre_actnorm_bias || a->pre_actnorm_scales)), "the fused ActNorm prologue is forward only");
    CNF_REQUIRE(a->z_masked_out == nullptr || (reinterpret_cast<uintptr_t>(a->z_masked_out) & 15) == 0, "z_masked_out must be 16-byte aligned");
    CNF_REQUIRE(p.P == 0 || (a->z && a->z_out), "z / z_out is NULL");
    CNF_REQUIRE(a->z != a->z_out, "invconv cannot run in place");
    const bool aligned = ((reinterpret_cast<uintptr_t>(a->z) | reinterpret_cast<uintptr_t>(a->z_out)) & 15) == 0;
    if (aligned && p.P > 0) {
        int handled = 0;
        const int rc = invconv_tile_try(a, stream, &handled);
        if (rc != CNF_OK || handled) return rc;
    }
    const int sms = sm_count();
    auto grid = [&](long long items) {
        long long blocks = (items + kThreads - 1) / kThreads;
        if (blocks < (a->B + kThreads - 1) / kThreads) blocks = (a->B + kThreads - 1) / kThreads;
        if (blocks > (long long)sms * 8) blocks = (long long)sms * 8;
        return (unsigned)(blocks < 1 ? 1 : blocks);
    };
    if (aligned && a->C == 4) invconv_rows_kernel<4><<<grid(p.P), kThreads, 0, stream>>>(p);
    else if (aligned && a->C == 8) invconv_rows_kernel<8><<<grid(p.P), kThreads, 0, stream>>>(p);
    else if (aligned && a->C == 16) invconv_rows_kernel<16><<<grid(p.P), kThreads, 0, stream>>>(p);
    else if (aligned && a->C == 32) invconv_rows_kernel<32><<<grid(p.P), kThreads, 0, stream>>>(p);
    else invconv_generic_kernel<<<grid(p.P * a->C), kThreads, sizeof(float) * a->C * a->C, stream>>>(p);
    return launch_status("invconv_apply");
}
