// Backward passes of the bandwidth-bound flow layers, so that the reference's
// unchanged training loops (loss.backward(), general/train.py:148-152) differentiate through the fused
// forward kernels:
//   affine coupling   layers/flows/coupling_layer.py:53-65, 76-98
//   ActNorm           layers/flows/activation_normalization.py:24-48
//   ExtActNorm        layers/flows/activation_normalization.py:116-144
//   1x1 convolution   layers/flows/permutation_layers.py:106-136   (dL/dz, dL/dW, dL/dsldj)
//   logistic prior    layers/flows/distributions.py:154-163
// Element index -> channel is fixed per thread (block size and grid stride are multiples of C), so the
// per-channel parameter gradients accumulate in registers and leave through one shared-memory reduction
// and one global atomic per (CTA, channel).
#include "cnf_common.cuh"

namespace cnf {
namespace {

// threads per block: the largest multiple of `period` that is <= 256 (period <= 128)
inline int block_for(int period) { return (256 / period) * period; }

inline unsigned grid_for_threads(long long work_items, int threads, int per_sm = 8) {
    long long blocks = (work_items + threads - 1) / threads;
    long long cap = (long long)sm_count() * per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (unsigned)blocks;
}

__device__ __forceinline__ float tanh_acc(float v) { return 1.0f - 2.0f / (1.0f + __expf(2.0f * v)); }

// reduce `v` (one value per thread, channel = threadIdx.x % C) over the block, then atomicAdd to dst[c]
__device__ __forceinline__ void block_channel_add(float* s_acc, float* dst, float v, int c, int C) {
    if (dst == nullptr) return;
    for (int i = threadIdx.x; i < C; i += blockDim.x) s_acc[i] = 0.f;
    __syncthreads();
    if (v != 0.f) atomicAdd(s_acc + c, v);
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x)
        if (s_acc[i] != 0.f) atomicAdd(dst + i, s_acc[i]);
    __syncthreads();
}

// --------------------------------------------------------------------------------------------
// affine coupling
// --------------------------------------------------------------------------------------------
struct AffineBwdParams {
    const float* z; const float2* nn; const float* sf; const float* gz_out; const float* gldj;
    float* gz; float2* gnn; float* gsf;
    long long n, SC;
    int S, C, reverse, pre;
    MaskView mask;
};

__global__ void affine_bwd_kernel(const AffineBwdParams p) {
    __shared__ float s_acc[CNF_MAX_CHANNELS];
    const long long stride = (long long)gridDim.x * blockDim.x;
    const int c = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) % p.C);
    const bool cond_c = (p.mask.cond_c >> c) & 1ull;
    const float fac = (p.sf && !p.pre) ? expf(p.sf[c]) : 1.0f;
    const float fmax_ = fmaxf(fac, 1.0f);
    float gsf = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += stride) {
        const long long pos = i / p.C;
        bool cond = cond_c;
        if (p.mask.s_period > 0) cond = cond || ((p.mask.cond_s >> ((pos % p.S) % p.mask.s_period)) & 1ull);
        const float go = p.gz_out[i];
        if (cond) {
            p.gz[i] = go;
            p.gnn[i] = make_float2(0.f, 0.f);
            continue;
        }
        const float2 st = p.nn[i];
        const float x = p.z[i];
        const float gl = p.gldj ? p.gldj[i / p.SC] : 0.f;
        const float th = p.pre ? 0.f : tanh_acc(st.x / fmax_);
        const float s = p.pre ? st.x : th * fac;
        float gx, gt, gs;
        if (!p.reverse) {          // out = (x + t) e^s, ldj += s
            const float es = __expf(s);
            gx = go * es;
            gt = gx;
            gs = go * (x + st.y) * es + gl;
        } else {                   // out = x e^{-s} - t, ldj -= s
            const float es = __expf(-s);
            gx = go * es;
            gt = -go;
            gs = -go * x * es - gl;
        }
        p.gz[i] = gx;
        if (p.pre) {
            p.gnn[i] = make_float2(gs, gt);
        } else {
            const float sech2 = 1.0f - th * th;
            p.gnn[i] = make_float2(gs * sech2 * fac / fmax_, gt);
            gsf += gs * fac * (th - (fac >= 1.0f ? sech2 * st.x / fac : 0.f));   // torch.clamp passes the gradient AT the bound (fac == 1: every freshly built layer, scaling_factor = 0)
        }
    }
    block_channel_add(s_acc, p.gsf, gsf, c, p.C);
}

// --------------------------------------------------------------------------------------------
// ActNorm
// --------------------------------------------------------------------------------------------
struct ActNormBwdParams {
    const float* z; const float* bias; const float* scales; const float* pad; const float* length;
    const float* gz_out; const float* gldj;
    float* gz; float* gbias; float* gscales;
    long long n, B;
    int S, C, reverse;
};

__global__ void actnorm_bwd_kernel(const ActNormBwdParams p) {
    __shared__ float s_acc[CNF_MAX_CHANNELS];
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const int c = (int)(gtid % p.C);
    const float b = p.bias[c], s = p.scales[c];
    const float e = __expf(p.reverse ? -s : s);
    float gb = 0.f, gs = 0.f;
    for (long long i = gtid; i < p.n; i += stride) {
        const float pv = p.pad ? p.pad[i / p.C] : 1.0f;
        const float go = p.gz_out[i] * pv;
        const float x = p.z[i];
        p.gz[i] = go * e;
        if (!p.reverse) {          // out = (z + b) e^s pad
            gb += go * e;
            gs += go * (x + b) * e;
        } else {                   // out = (z e^{-s} - b) pad
            gb -= go;
            gs -= go * x * e;
        }
    }
    // ldj[b] += (+/-) sum_c s_c len_b  ->  dL/ds_c += (+/-) sum_b gldj[b] len_b (same for every channel)
    if (p.gldj != nullptr && blockIdx.x == 0) {
        float t = 0.f;
        for (long long bi = threadIdx.x / p.C; bi < p.B; bi += blockDim.x / p.C) {
            float len;
            if (p.length) len = p.length[bi];
            else if (p.pad) { len = 0.f; for (int q = 0; q < p.S; ++q) len += p.pad[bi * p.S + q]; }
            else len = (float)p.S;
            t += p.gldj[bi] * len;
        }
        gs += p.reverse ? -t : t;
    }
    block_channel_add(s_acc, p.gbias, gb, c, p.C);
    block_channel_add(s_acc, p.gscales, gs, c, p.C);
}

// C % 4 == 0: four channels per thread with 16-byte streaming accesses; the thread's channels stay fixed over its
// grid-stride loop (block size and grid stride are multiples of C / 4), the position index - needed for the padding mask
// only - comes from a float-reciprocal division instead of a 64-bit one per element (the scalar kernel above: 116 us for
// the 201 MB of the LM shape, instruction bound).
__global__ void actnorm_bwd4_kernel(const ActNormBwdParams p) {
    __shared__ float s_acc[CNF_MAX_CHANNELS];
    const int C4 = p.C >> 2;
    const long long n4 = p.n >> 2;
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const int c0 = (int)(gtid % C4) * 4;
    float e[4], b[4], gb[4], gs[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float s = p.scales[c0 + j];
        e[j] = __expf(p.reverse ? -s : s);
        b[j] = p.bias[c0 + j];
        gb[j] = 0.f;
        gs[j] = 0.f;
    }
    for (long long i = gtid; i < n4; i += stride) {
        const float pv = p.pad ? p.pad[i / C4] : 1.0f;
        const float4 g = ldg_stream4(reinterpret_cast<const float4*>(p.gz_out) + i);
        const float4 x = ldg_stream4(reinterpret_cast<const float4*>(p.z) + i);
        const float go[4] = {g.x * pv, g.y * pv, g.z * pv, g.w * pv};
        const float xv[4] = {x.x, x.y, x.z, x.w};
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            o[j] = go[j] * e[j];
            if (!p.reverse) {
                gb[j] += o[j];
                gs[j] += o[j] * (xv[j] + b[j]);
            } else {
                gb[j] -= go[j];
                gs[j] -= o[j] * xv[j];
            }
        }
        stg_stream4(reinterpret_cast<float4*>(p.gz) + i, make_float4(o[0], o[1], o[2], o[3]));
    }
    if (p.gldj != nullptr && blockIdx.x == 0) {
        // ldj[b] += (+/-) sum_c s_c len_b -> dL/ds_c += (+/-) sum_b gldj[b] len_b (same for every channel): the threads of
        // the first block that share a channel group split the samples
        float t = 0.f;
        for (long long bi = threadIdx.x / C4; bi < p.B; bi += blockDim.x / C4) {
            float len;
            if (p.length) len = p.length[bi];
            else if (p.pad) { len = 0.f; for (int q = 0; q < p.S; ++q) len += p.pad[bi * p.S + q]; }
            else len = (float)p.S;
            t += p.gldj[bi] * len;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) gs[j] += p.reverse ? -t : t;
    }
    for (int i = threadIdx.x; i < 2 * p.C; i += blockDim.x) s_acc[i] = 0.f;      // [gb | gs], 2 C <= CNF_MAX_CHANNELS checked by the launcher
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (gb[j] != 0.f) atomicAdd(s_acc + c0 + j, gb[j]);
        if (gs[j] != 0.f) atomicAdd(s_acc + p.C + c0 + j, gs[j]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < p.C; i += blockDim.x) {
        if (p.gbias && s_acc[i] != 0.f) atomicAdd(p.gbias + i, s_acc[i]);
        if (p.gscales && s_acc[p.C + i] != 0.f) atomicAdd(p.gscales + i, s_acc[p.C + i]);
    }
}

// --------------------------------------------------------------------------------------------
// ExtActNorm: ext = [bias | raw scale] per element, s = tanh(raw)
// --------------------------------------------------------------------------------------------
struct ExtBwdParams {
    const float* z; const float* ext; const float* pad; const float* gz_out; const float* gldj;
    float* gz; float* gext;
    long long n, SC;
    int C, reverse;
};

__global__ void ext_actnorm_bwd_kernel(const ExtBwdParams p) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += stride) {
        const long long pos = i / p.C;
        const int c = (int)(i - pos * p.C);
        const float b = p.ext[pos * 2 * p.C + c], raw = p.ext[pos * 2 * p.C + p.C + c];
        const float s = tanh_acc(raw);
        const float pv = p.pad ? p.pad[pos] : 1.0f;
        const float gl = (p.gldj ? p.gldj[i / p.SC] : 0.f) * pv;
        const float go = p.gz_out[i];
        const float x = p.z[i];
        float gx, gb, gs;
        if (!p.reverse) {          // out = (z + b) e^s ; ldj += s pad
            const float es = __expf(s);
            gx = go * es;
            gb = gx;
            gs = go * (x + b) * es + gl;
        } else {                   // out = z e^{-s} - b ; ldj -= s pad
            const float es = __expf(-s);
            gx = go * es;
            gb = -go;
            gs = -go * x * es - gl;
        }
        p.gz[i] = gx;
        p.gext[pos * 2 * p.C + c] = gb;
        p.gext[pos * 2 * p.C + p.C + c] = gs * (1.0f - s * s);
    }
}

// --------------------------------------------------------------------------------------------
// 1x1 convolution: out = (z @ W) pad
//   dL/dz = (g pad) @ W^T ; dL/dW[c][o] = sum_pos z[pos][c] g[pos][o] pad ; dL/dsldj = (+/-) sum_b gldj[b] len_b
// --------------------------------------------------------------------------------------------
struct ConvBwdParams {
    const float* z; const float* w; const float* pad; const float* length; const float* gz_out; const float* gldj;
    float* gz; float* gw; float* gsldj;
    long long P, B;
    int S, C, reverse, TP;
};

__global__ void __launch_bounds__(256) invconv_bwd_kernel(const ConvBwdParams p) {
    extern __shared__ float sm[];
    const int C = p.C, TP = p.TP;
    const int CP = C | 1;            // odd pitch: the lanes of a warp hold different rows c of W -> distinct banks
    float* s_w = sm;                 // [C][CP]
    float* s_z = s_w + C * CP;       // [TP*C]
    float* s_g = s_z + TP * C;       // [TP*C]  dL/dout * pad
    for (int i = threadIdx.x; i < C * C; i += 256) s_w[(i / C) * CP + (i % C)] = p.w[i];
    // per-pair accumulators: pair index = threadIdx.x + k*256 < C*C, at most C*C/256 = 16 per thread (C <= 64)
    float acc[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = 0.f;
    const long long ntiles = (p.P + TP - 1) / TP;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const long long pos0 = t * TP;
        const int rows = (int)min((long long)TP, p.P - pos0);
        __syncthreads();
        for (int i = threadIdx.x; i < rows * C; i += 256) {
            const float pv = p.pad ? p.pad[pos0 + i / C] : 1.0f;
            s_z[i] = p.z[pos0 * C + i];
            s_g[i] = p.gz_out[pos0 * C + i] * pv;
        }
        __syncthreads();
        // dL/dz[pos][c] = sum_o g[pos][o] W[c][o]
        for (int i = threadIdx.x; i < rows * C; i += 256) {
            const int r = i / C, c = i - r * C;
            float a = 0.f;
            for (int o = 0; o < C; ++o) a = fmaf(s_g[r * C + o], s_w[c * CP + o], a);
            p.gz[pos0 * C + i] = a;
        }
        if (p.gw != nullptr) {
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const int pair = threadIdx.x + k * 256;
                if (pair < C * C) {
                    const int c = pair / C, o = pair - c * C;
                    float a = acc[k];
                    for (int r = 0; r < rows; ++r) a = fmaf(s_z[r * C + c], s_g[r * C + o], a);
                    acc[k] = a;
                }
            }
        }
    }
    if (p.gw != nullptr) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int pair = threadIdx.x + k * 256;
            if (pair < C * C && acc[k] != 0.f) atomicAdd(p.gw + pair, acc[k]);
        }
    }
    if (p.gsldj != nullptr && p.gldj != nullptr && blockIdx.x == 0) {
        float tsum = 0.f;
        for (long long b = threadIdx.x; b < p.B; b += 256) tsum += p.gldj[b] * (p.length ? p.length[b] : (float)p.S);
        tsum = warp_sum(tsum);
        if ((threadIdx.x & 31) == 0 && tsum != 0.f) atomicAdd(p.gsldj, p.reverse ? -tsum : tsum);
    }
}

// C = 16: thread per position.  The tile kernel above issues two shared-memory loads per FMA (l1tex-bound: 164 us for the
// 201 MB of the LM shape); here a thread keeps its position's z and g rows in registers:
//   dL/dz  : 16 dot products against W rows read as 16-byte shared-memory broadcasts (4 FMA per load)
//   dL/dW  : the warp's 32 rows are staged in a per-warp shared-memory tile (16-byte chunk j of row r at chunk
//            j ^ ((r >> 1) & 3): conflict-free row writes), then lane l accumulates the 8 entries (c = l / 2,
//            o = 8 (l % 2) ..) of z^T g over the 32 rows - 3 loads per 8 FMA - in registers across all of its tiles;
//            one shared + one global atomic per entry and CTA at the end.
// No CTA barrier in the position loop.
__global__ void __launch_bounds__(256) invconv_bwd16_kernel(const ConvBwdParams p) {
    constexpr int C = 16;
    __shared__ __align__(16) float s_w[C * C];            // W[c][o]
    __shared__ __align__(16) float s_zt[8][32 * C];       // per warp: z rows of the current 32 positions (swizzled)
    __shared__ __align__(16) float s_gt[8][32 * C];       // per warp: g rows
    __shared__ float s_gw[C * C];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < C * C; i += 256) { s_w[i] = p.w[i]; s_gw[i] = 0.f; }
    __syncthreads();
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    const int ac = lane >> 1, ao = (lane & 1) * 8;        // this lane's dL/dW entries: row ac, columns ao .. ao + 7
    float* zt = s_zt[warp];
    float* gt = s_gt[warp];
    const int sw = (lane >> 1) & 3;
    const long long n_round = (p.P + 31) & ~31ll;
    const long long stride = (long long)gridDim.x * 256;
    for (long long pos = (long long)blockIdx.x * 256 + tid; pos < n_round; pos += stride) {
        const bool in = pos < p.P;
        float4 z4[4], g4[4];
        const float pv = (in && p.pad) ? p.pad[pos] : 1.0f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            z4[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            g4[j] = z4[j];
            if (in) {
                z4[j] = ldg_stream4(reinterpret_cast<const float4*>(p.z + pos * C) + j);
                const float4 g = ldg_stream4(reinterpret_cast<const float4*>(p.gz_out + pos * C) + j);
                g4[j] = make_float4(g.x * pv, g.y * pv, g.z * pv, g.w * pv);
            }
        }
        if (p.gw != nullptr) {
            __syncwarp();      // the previous tile's reads are done
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                reinterpret_cast<float4*>(zt + lane * C)[j ^ sw] = z4[j];
                reinterpret_cast<float4*>(gt + lane * C)[j ^ sw] = g4[j];
            }
            __syncwarp();
#pragma unroll 8
            for (int r = 0; r < 32; ++r) {
                const int rs = (r >> 1) & 3;
                const float zc = zt[r * C + ((((ac >> 2) ^ rs) << 2) | (ac & 3))];
                const float4 ga = reinterpret_cast<const float4*>(gt + r * C)[(ao >> 2) ^ rs];
                const float4 gb = reinterpret_cast<const float4*>(gt + r * C)[((ao >> 2) + 1) ^ rs];
                acc[0] = fmaf(zc, ga.x, acc[0]); acc[1] = fmaf(zc, ga.y, acc[1]);
                acc[2] = fmaf(zc, ga.z, acc[2]); acc[3] = fmaf(zc, ga.w, acc[3]);
                acc[4] = fmaf(zc, gb.x, acc[4]); acc[5] = fmaf(zc, gb.y, acc[5]);
                acc[6] = fmaf(zc, gb.z, acc[6]); acc[7] = fmaf(zc, gb.w, acc[7]);
            }
        }
        if (in) {
            // dL/dz[c] = sum_o g[o] W[c][o]
            const float g[16] = {g4[0].x, g4[0].y, g4[0].z, g4[0].w, g4[1].x, g4[1].y, g4[1].z, g4[1].w,
                                 g4[2].x, g4[2].y, g4[2].z, g4[2].w, g4[3].x, g4[3].y, g4[3].z, g4[3].w};
            float out[16];
#pragma unroll
            for (int c = 0; c < C; ++c) {
                float a = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 wv = reinterpret_cast<const float4*>(s_w + c * C)[j];
                    a = fmaf(g[4 * j], wv.x, a); a = fmaf(g[4 * j + 1], wv.y, a);
                    a = fmaf(g[4 * j + 2], wv.z, a); a = fmaf(g[4 * j + 3], wv.w, a);
                }
                out[c] = a;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
                stg_stream4(reinterpret_cast<float4*>(p.gz + pos * C) + j, make_float4(out[4 * j], out[4 * j + 1], out[4 * j + 2], out[4 * j + 3]));
        }
    }
    if (p.gw != nullptr) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (acc[k] != 0.f) atomicAdd(s_gw + ac * C + ao + k, acc[k]);
        __syncthreads();
        for (int i = tid; i < C * C; i += 256)
            if (s_gw[i] != 0.f) atomicAdd(p.gw + i, s_gw[i]);
    }
    if (p.gsldj != nullptr && p.gldj != nullptr && blockIdx.x == 0) {
        float tsum = 0.f;
        for (long long b = tid; b < p.B; b += 256) tsum += p.gldj[b] * (p.length ? p.length[b] : (float)p.S);
        tsum = warp_sum(tsum);
        if (lane == 0 && tsum != 0.f) atomicAdd(p.gsldj, p.reverse ? -tsum : tsum);
    }
}

// --------------------------------------------------------------------------------------------
// logistic log-density: d/dx -(softplus(v) + softplus(-v) + log sigma) = -tanh(v / 2) / sigma, v = (x - mu) / sigma
// --------------------------------------------------------------------------------------------
struct LogProbBwdParams {
    const float* x; const float* pad; const float* g_elem; const float* g_sum;
    float* gx;
    long long n, SC;
    int C;
    float mu, inv_sigma;
};

__global__ void logistic_logprob_bwd_kernel(const LogProbBwdParams p) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += stride) {
        const float v = (p.x[i] - p.mu) * p.inv_sigma;
        float g = p.g_elem ? p.g_elem[i] : 0.f;
        if (p.g_sum) g += p.g_sum[i / p.SC] * (p.pad ? p.pad[i / p.C] : 1.0f);
        p.gx[i] = -g * tanh_acc(0.5f * v) * p.inv_sigma;
    }
}

}  // namespace
}  // namespace cnf

using namespace cnf;

extern "C" int cnf_affine_coupling_bwd(const cnf_affine_bwd_args* a, cnf_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    CNF_REQUIRE(a != nullptr && a->B >= 0 && a->S >= 0 && a->C >= 1, "bad arguments");
    CNF_SUPPORTED(a->C <= CNF_MAX_CHANNELS, "C=%d exceeds CNF_MAX_CHANNELS", a->C);
    AffineBwdParams p{};
    int rc = build_mask(a->mask, a->C, &p.mask);
    if (rc != CNF_OK) return rc;
    p.n = a->B * a->S * a->C;
    if (p.n == 0) return CNF_OK;
    CNF_REQUIRE(a->z && a->nn_out && a->grad_z_out && a->grad_z && a->grad_nn_out, "null tensor");
    CNF_REQUIRE((reinterpret_cast<uintptr_t>(a->nn_out) & 7) == 0 && (reinterpret_cast<uintptr_t>(a->grad_nn_out) & 7) == 0,
                "nn_out / grad_nn_out must be 8-byte aligned");
    p.z = a->z; p.nn = reinterpret_cast<const float2*>(a->nn_out); p.sf = a->scaling_factor; p.gz_out = a->grad_z_out;
    p.gldj = a->grad_ldj; p.gz = a->grad_z; p.gnn = reinterpret_cast<float2*>(a->grad_nn_out);
    p.gsf = a->params_prebounded ? nullptr : a->grad_scaling_factor;
    p.SC = a->S * a->C; p.S = (int)a->S; p.C = a->C; p.reverse = a->reverse; p.pre = a->params_prebounded;
    const int threads = block_for(a->C);
    affine_bwd_kernel<<<grid_for_threads(p.n, threads), threads, 0, stream>>>(p);
    return launch_status("affine_bwd_kernel");
}

extern "C" int cnf_actnorm_bwd(const cnf_actnorm_bwd_args* a, cnf_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    CNF_REQUIRE(a != nullptr && a->B >= 0 && a->S >= 0 && a->C >= 1, "bad arguments");
    CNF_SUPPORTED(a->C <= CNF_MAX_CHANNELS, "C=%d exceeds CNF_MAX_CHANNELS", a->C);
    ActNormBwdParams p{};
    p.n = a->B * a->S * a->C;
    CNF_REQUIRE(a->bias && a->scales, "bias / scales is NULL");
    if (a->B == 0) return CNF_OK;
    CNF_REQUIRE(p.n == 0 || (a->z && a->grad_z_out && a->grad_z), "null tensor");
    p.z = a->z; p.bias = a->bias; p.scales = a->scales; p.pad = a->pad; p.length = a->length;
    p.gz_out = a->grad_z_out; p.gldj = a->grad_ldj; p.gz = a->grad_z; p.gbias = a->grad_bias; p.gscales = a->grad_scales;
    p.B = a->B; p.S = (int)a->S; p.C = a->C; p.reverse = a->reverse;
    if (a->C % 4 == 0 && 2 * a->C <= CNF_MAX_CHANNELS && p.n > 0 &&
        ((reinterpret_cast<uintptr_t>(a->z) | reinterpret_cast<uintptr_t>(a->grad_z_out) | reinterpret_cast<uintptr_t>(a->grad_z)) & 15) == 0) {
        const int threads4 = block_for(a->C / 4);
        actnorm_bwd4_kernel<<<grid_for_threads(p.n / 4, threads4, 4), threads4, 0, stream>>>(p);
        return launch_status("actnorm_bwd4_kernel");
    }
    const int threads = block_for(a->C);
    actnorm_bwd_kernel<<<grid_for_threads(p.n > 0 ? p.n : 1, threads), threads, 0, stream>>>(p);
    return launch_status("actnorm_bwd_kernel");
}

extern "C" int cnf_ext_actnorm_bwd(const cnf_ext_actnorm_bwd_args* a, cnf_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    CNF_REQUIRE(a != nullptr && a->B >= 0 && a->S >= 0 && a->C >= 1, "bad arguments");
    ExtBwdParams p{};
    p.n = a->B * a->S * a->C;
    if (p.n == 0) return CNF_OK;
    CNF_REQUIRE(a->z && a->ext && a->grad_z_out && a->grad_z && a->grad_ext, "null tensor");
    p.z = a->z; p.ext = a->ext; p.pad = a->pad; p.gz_out = a->grad_z_out; p.gldj = a->grad_ldj; p.gz = a->grad_z; p.gext = a->grad_ext;
    p.SC = a->S * a->C; p.C = a->C; p.reverse = a->reverse;
    ext_actnorm_bwd_kernel<<<grid_for_threads(p.n, 256), 256, 0, stream>>>(p);
    return launch_status("ext_actnorm_bwd_kernel");
}

extern "C" int cnf_invconv_bwd(const cnf_invconv_bwd_args* a, cnf_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    CNF_REQUIRE(a != nullptr && a->B >= 0 && a->S >= 0 && a->C >= 1, "bad arguments");
    CNF_SUPPORTED(a->C <= CNF_MAX_CHANNELS, "C=%d exceeds CNF_MAX_CHANNELS", a->C);
    ConvBwdParams p{};
    p.P = a->B * a->S;
    CNF_REQUIRE(a->weight != nullptr, "weight is NULL");
    if (a->B == 0) return CNF_OK;
    CNF_REQUIRE(p.P == 0 || (a->z && a->grad_z_out && a->grad_z), "null tensor");
    p.z = a->z; p.w = a->weight; p.pad = a->pad; p.length = a->length; p.gz_out = a->grad_z_out; p.gldj = a->grad_ldj;
    p.gz = a->grad_z; p.gw = a->grad_weight; p.gsldj = a->grad_sldj;
    p.B = a->B; p.S = (int)a->S; p.C = a->C; p.reverse = a->reverse;
    if (a->C == 16 && p.P > 0 &&
        ((reinterpret_cast<uintptr_t>(a->z) | reinterpret_cast<uintptr_t>(a->grad_z_out) | reinterpret_cast<uintptr_t>(a->grad_z)) & 15) == 0) {
        long long grid16 = (p.P + 255) / 256;
        const long long cap16 = (long long)sm_count() * 4;
        if (grid16 > cap16) grid16 = cap16;
        invconv_bwd16_kernel<<<(unsigned)grid16, 256, 0, stream>>>(p);
        return launch_status("invconv_bwd16_kernel");
    }
    p.TP = a->C <= 16 ? 256 : (a->C <= 32 ? 128 : 64);
    const size_t smem = ((size_t)a->C * (a->C | 1) + 2 * (size_t)p.TP * a->C) * sizeof(float);
    long long grid = (p.P + p.TP - 1) / p.TP;
    const long long cap = (long long)sm_count() * 4;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    if (smem > 48 * 1024) CNF_CUDA(cudaFuncSetAttribute(invconv_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    invconv_bwd_kernel<<<(unsigned)grid, 256, smem, stream>>>(p);
    return launch_status("invconv_bwd_kernel");
}

extern "C" int cnf_logistic_logprob_bwd(const cnf_logistic_logprob_bwd_args* a, cnf_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    CNF_REQUIRE(a != nullptr && a->B >= 0 && a->S >= 0 && a->C >= 1, "bad arguments");
    CNF_REQUIRE(a->sigma > 0.f, "sigma must be positive");
    LogProbBwdParams p{};
    p.n = a->B * a->S * a->C;
    if (p.n == 0) return CNF_OK;
    CNF_REQUIRE(a->x && a->grad_x && (a->grad_elementwise || a->grad_out), "null tensor");
    p.x = a->x; p.pad = a->pad; p.g_elem = a->grad_elementwise; p.g_sum = a->grad_out; p.gx = a->grad_x;
    p.SC = a->S * a->C; p.C = a->C; p.mu = a->mu; p.inv_sigma = 1.0f / a->sigma;
    logistic_logprob_bwd_kernel<<<grid_for_threads(p.n, 256), 256, 0, stream>>>(p);
    return launch_status("logistic_logprob_bwd_kernel");
}
