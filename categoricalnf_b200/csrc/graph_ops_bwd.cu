// Backward passes of the graph coupling networks' glue (csrc/graph_ops.cu), so that training the graph flows
// (general/train.py:148-152 differentiates RGCNNet / EdgeGNN through autograd) stays on hand-written kernels:
//   cnf_gelu                 nn.GELU forward / backward where it is not fused into a projection epilogue
//   cnf_layernorm_bwd        nn.LayerNorm                     (layers/networks/graph_layers.py:22,63,190,313-314)
//   cnf_skip_gate_bwd        GNNSkipConnection                (:722-733)
//   cnf_graph_aggregate_bwd  RelationGraphConv / RelationGraphAttention neighbour aggregation (:27-50, :76-154)
//   cnf_edge_aggregate_bwd   Edge2NodeAttnLayer / Edge2NodeQKVAttnLayer (:432-502, :595-645)
//   cnf_pair_combine_bwd     Node2EdgePlainLayer              (:317-336)
// Each aggregation backward is ONE kernel with the forward's decomposition (one CTA per receiving node): it rebuilds the
// neighbour list and the attention weights, recomputes the pre-activation output, and scatters the gradients to the
// neighbours' rows with fp32 reductions in L2 (red.global.add) - gradient buffers that receive scattered contributions
// are accumulated into and must be zero-initialised by the caller, as documented per entry point in cnf_b200.h.
#include "cnf_common.cuh"

namespace cnf {
namespace {

constexpr float kInvSqrt2 = 0.70710678118654752f;
constexpr float kInvSqrt2Pi = 0.3989422804014327f;

__device__ __forceinline__ float gelu_erf(float v) { return 0.5f * v * (1.0f + erff(v * kInvSqrt2)); }
__device__ __forceinline__ float gelu_grad(float v) {
    return fmaf(v * kInvSqrt2Pi, expf(-0.5f * v * v), 0.5f * (1.0f + erff(v * kInvSqrt2)));
}

// Four scalar fp32 reductions (RED.E.ADD.F32).  The 16-byte form (REDG.E.ADD.F32x4 via atomicAdd(float4*)) was measured
// SLOWER on B200 for this scatter pattern (graph_aggregate_bwd 703 us vs 520 us at B*N = 20480, 768 features), so the
// vector paths below only vectorise the loads.
__device__ __forceinline__ void red_add4(float* addr, float a, float b, float c, float d) {
    atomicAdd(addr, a); atomicAdd(addr + 1, b); atomicAdd(addr + 2, c); atomicAdd(addr + 3, d);
}

inline unsigned capped_grid(long long blocks, int per_sm) {
    const long long cap = (long long)sm_count() * per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (unsigned)blocks;
}

// ---- GELU ----------------------------------------------------------------------------------------------------------
struct GeluParams { const float* x; const float* g; float* y; long long n; };

__global__ void __launch_bounds__(256) gelu_kernel(const GeluParams p) {
    const long long stride = (long long)gridDim.x * 256;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < p.n; i += stride) {
        const float v = p.x[i];
        p.y[i] = p.g ? p.g[i] * gelu_grad(v) : gelu_erf(v);
    }
}

// ---- LayerNorm backward: one warp per row, parameter gradients reduced per CTA in shared memory ---------------------
struct LnBwdParams {
    const float* x; const float* gamma; const float* g; float* gx; float* ggamma; float* gbeta;
    long long M; int H; float eps;
};

__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const LnBwdParams p) {
    extern __shared__ float ln_smem[];       // [2][H]: d gamma, d beta of this CTA
    float* sG = ln_smem;
    float* sB = ln_smem + p.H;
    const bool params = p.ggamma != nullptr;
    if (params) for (int c = threadIdx.x; c < 2 * p.H; c += 256) ln_smem[c] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long nwarps = (long long)gridDim.x * 8;
    const float invH = 1.0f / (float)p.H;
    for (long long m = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); m < p.M; m += nwarps) {
        const float* row = p.x + m * p.H;
        const float* grow = p.g + m * p.H;
        float s = 0.f;
        for (int c = lane; c < p.H; c += 32) s += row[c];
        const float mean = warp_sum(s) * invH;
        float q = 0.f;
        for (int c = lane; c < p.H; c += 32) { const float a = row[c] - mean; q += a * a; }
        const float rstd = rsqrtf(warp_sum(q) * invH + p.eps);
        float s1 = 0.f, s2 = 0.f;
        for (int c = lane; c < p.H; c += 32) {
            const float gy = grow[c] * p.gamma[c];
            s1 += gy;
            s2 = fmaf(gy, (row[c] - mean) * rstd, s2);
        }
        s1 = warp_sum(s1) * invH;
        s2 = warp_sum(s2) * invH;
        float* out = p.gx + m * p.H;
        for (int c = lane; c < p.H; c += 32) {
            const float xhat = (row[c] - mean) * rstd;
            const float gv = grow[c];
            out[c] = rstd * (gv * p.gamma[c] - s1 - xhat * s2);
            if (params) { atomicAdd(sG + c, gv * xhat); atomicAdd(sB + c, gv); }
        }
    }
    __syncthreads();
    if (params)
        for (int c = threadIdx.x; c < p.H; c += 256) { atomicAdd(p.ggamma + c, sG[c]); atomicAdd(p.gbeta + c, sB[c]); }
}

// ---- skip connection backward -----------------------------------------------------------------------------------------
struct GateBwdParams {
    const float* orig; const float* s; const float* g; float* g_orig; float* g_s;
    long long M; int H, config;
};

__global__ void __launch_bounds__(256) skip_gate_bwd_kernel(const GateBwdParams p) {
    const long long total = p.M * p.H;
    const long long stride = (long long)gridDim.x * 256;
    const int ld = p.config == 0 ? p.H : 2 * p.H;
    for (long long idx = (long long)blockIdx.x * 256 + threadIdx.x; idx < total; idx += stride) {
        const long long m = idx / p.H;
        const int c = (int)(idx - m * p.H);
        const float g = p.g[idx];
        if (p.config == 0) {
            p.g_orig[idx] = g;
            p.g_s[m * ld + c] = g;
        } else {
            const float val = p.s[m * ld + c];
            const float sg = 1.0f / (1.0f + expf(-p.s[m * ld + p.H + c]));
            const float dsg = sg * (1.0f - sg);
            p.g_s[m * ld + c] = g * sg;
            if (p.config == 1) {
                p.g_orig[idx] = g;
                p.g_s[m * ld + p.H + c] = g * val * dsg;
            } else {
                p.g_orig[idx] = g * (1.0f - sg);
                p.g_s[m * ld + p.H + c] = g * (val - p.orig[idx]) * dsg;
            }
        }
    }
}

// ---- neighbour aggregation backward: one CTA per receiving node --------------------------------------------------------
struct AggBwdParams {
    const long long* adj;
    const float* hs; const float* hr; const float* score_s; const float* score_r; const float* num_neighbours;
    const float* g_out;
    float* g_hs; float* g_hr; float* g_ss; float* g_sr;
    long long ld_hs, ld_hr, ld_ss, ld_sr, ld_ghs, ld_ghr, ld_gss, ld_gsr;
    int N, E, H, Dh, mode, act, vec;
    float slope;
};

constexpr int kAggThreads = 256;
constexpr int kMaxHeads = 16;

__global__ void __launch_bounds__(kAggThreads) graph_aggregate_bwd_kernel(const AggBwdParams p) {
    extern __shared__ __align__(16) unsigned char aggb_smem[];
    const int N1 = p.N + 1;
    const int HD = p.H * p.Dh;
    float* gp = reinterpret_cast<float*>(aggb_smem);              // [HD] gradient at the pre-activation output (16-byte aligned)
    int* nb_row = reinterpret_cast<int*>(gp + ((HD + 3) & ~3));   // [N+1]
    int* nb_e = nb_row + N1;                                      // [N+1]
    float* wgt = reinterpret_cast<float*>(nb_e + N1);             // [H][N+1] softmax weights (mode 1) / [0] = 1/n (mode 0)
    float* dlk = wgt + p.H * N1;                                  // [H][N+1] derivative of the leaky ReLU at the logit
    float* dsum = dlk + p.H * N1;                                 // [H] sum_j p_ij <g, hr_j>  = <g, out_pre>
    float* gss = dsum + p.H;                                      // [H] gradient of score_s
    __shared__ int s_cnt;

    const long long node = blockIdx.x;
    const long long b = node / p.N;
    const int i = (int)(node - b * p.N);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0) {
        const long long* base = p.adj + b * p.N * p.N;
        int cnt = 0;
        for (int j0 = 0; j0 < p.N; j0 += 32) {
            const int j = j0 + lane;
            long long e = 0;
            if (j < p.N) e = p.mode == 1 ? base[(long long)i * p.N + j] : base[(long long)j * p.N + i];
            const bool valid = e > 0 && e <= p.E;
            const unsigned m = __ballot_sync(0xffffffffu, valid);
            if (valid) {
                const int pos = cnt + __popc(m & ((1u << lane) - 1u));
                nb_row[pos] = j;
                nb_e[pos] = (int)e - 1;
            }
            cnt += __popc(m);
        }
        if (p.mode == 1 && lane == 0) { nb_row[cnt] = i; nb_e[cnt] = p.E; cnt += 1; }
        if (lane == 0) s_cnt = cnt;
    }
    if (threadIdx.x < p.H) { dsum[threadIdx.x] = 0.f; gss[threadIdx.x] = 0.f; }
    __syncthreads();
    const int cnt = s_cnt;

    if (p.mode == 1) {
        for (int h = warp; h < p.H; h += kAggThreads / 32) {
            const float s_self = p.score_s[node * p.ld_ss + h];
            float mx = -3.0e38f;
            for (int n = lane; n < cnt; n += 32) {
                float l = s_self + p.score_r[(b * p.N + nb_row[n]) * p.ld_sr + nb_e[n] * p.H + h];
                dlk[h * N1 + n] = l > 0.f ? 1.0f : p.slope;
                l = l > 0.f ? l : l * p.slope;
                wgt[h * N1 + n] = l;
                mx = fmaxf(mx, l);
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
            float sum = 0.f;
            for (int n = lane; n < cnt; n += 32) {
                const float e = expf(wgt[h * N1 + n] - mx);
                wgt[h * N1 + n] = e;
                sum += e;
            }
            sum = warp_sum(sum);
            const float inv = 1.0f / sum;
            for (int n = lane; n < cnt; n += 32) wgt[h * N1 + n] *= inv;
        }
    } else if (threadIdx.x == 0) {
        const float nn = p.num_neighbours != nullptr ? p.num_neighbours[node] : (float)cnt;
        wgt[0] = 1.0f / fmaxf(nn, 1e-5f);
    }
    __syncthreads();

    // pass 1: pre-activation output of this node, gradient at it
    for (int f = threadIdx.x; f < HD; f += kAggThreads) {
        const int h = f / p.Dh;
        float acc = 0.f;
        if (p.act == 1 || p.mode == 1) {
            for (int n = 0; n < cnt; ++n) {
                const float w = p.mode == 1 ? wgt[h * N1 + n] : 1.0f;
                acc = fmaf(w, p.hr[(b * p.N + nb_row[n]) * p.ld_hr + (long long)nb_e[n] * HD + f], acc);
            }
            if (p.mode == 0) acc = fmaf(acc, wgt[0], p.hs[node * p.ld_hs + f]);
        }
        float g = p.g_out[node * HD + f];
        if (p.act == 1) g *= gelu_grad(acc);
        gp[f] = g;
        if (p.mode == 0) p.g_hs[node * p.ld_ghs + f] = g;
        else atomicAdd(dsum + h, g * acc);
    }
    __syncthreads();

    // pass 2: scatter to the neighbours' rows
    if (p.mode == 0) {
        const float inv = wgt[0];
        for (int n = 0; n < cnt; ++n) {
            float* dst = p.g_hr + (b * p.N + nb_row[n]) * p.ld_ghr + (long long)nb_e[n] * HD;
            if (p.vec) {
                for (int f = threadIdx.x * 4; f < HD; f += kAggThreads * 4)
                    red_add4(dst + f, gp[f] * inv, gp[f + 1] * inv, gp[f + 2] * inv, gp[f + 3] * inv);
            } else {
                for (int f = threadIdx.x; f < HD; f += kAggThreads) atomicAdd(dst + f, gp[f] * inv);
            }
        }
    } else {
        for (int w = warp; w < cnt * p.H; w += kAggThreads / 32) {
            const int n = w / p.H, h = w - n * p.H;
            const long long row = b * p.N + nb_row[n];
            const float* src = p.hr + row * p.ld_hr + (long long)nb_e[n] * HD + h * p.Dh;
            float* dst = p.g_hr + row * p.ld_ghr + (long long)nb_e[n] * HD + h * p.Dh;
            const float pw = wgt[h * N1 + n];
            float dot = 0.f;
            if (p.vec) {
                for (int d = lane * 4; d < p.Dh; d += 128) {
                    const float4 gv = *reinterpret_cast<const float4*>(gp + h * p.Dh + d);
                    const float4 sv = *reinterpret_cast<const float4*>(src + d);
                    dot = fmaf(gv.x, sv.x, fmaf(gv.y, sv.y, fmaf(gv.z, sv.z, fmaf(gv.w, sv.w, dot))));
                    red_add4(dst + d, pw * gv.x, pw * gv.y, pw * gv.z, pw * gv.w);
                }
            } else {
                for (int d = lane; d < p.Dh; d += 32) {
                    const float gv = gp[h * p.Dh + d];
                    dot = fmaf(gv, src[d], dot);
                    atomicAdd(dst + d, pw * gv);
                }
            }
            dot = warp_sum(dot);
            if (lane == 0) {
                const float gl = pw * (dot - dsum[h]) * dlk[h * N1 + n];      // softmax, then leaky ReLU
                atomicAdd(p.g_sr + row * p.ld_gsr + nb_e[n] * p.H + h, gl);
                atomicAdd(gss + h, gl);
            }
        }
        __syncthreads();
        if (threadIdx.x < p.H) p.g_ss[node * p.ld_gss + threadIdx.x] = gss[threadIdx.x];
    }
}

// ---- Edge-GNN edge -> node attention backward: one CTA per receiving node ----------------------------------------------
struct EdgeAggBwdParams {
    const long long* rev;
    const float* node_val; const float* node_q; const float* node_k; const float* edge_val; const float* edge_logit;
    const float* g_out;
    float* g_node_val; float* g_node_q; float* g_node_k; float* g_edge_val; float* g_edge_logit;
    long long ld_nv, ld_q, ld_k, ld_ev, ld_el, ld_gnv, ld_gq, ld_gk, ld_gev, ld_gel;
    int N, P, H, Dh, mode, vec;
    float scale;
};

__device__ __forceinline__ int pair_index(int a, int b, int N) { return a * (N - 1) - (a * (a - 1)) / 2 + (b - a - 1); }

__global__ void __launch_bounds__(kAggThreads) edge_aggregate_bwd_kernel(const EdgeAggBwdParams p) {
    extern __shared__ __align__(16) unsigned char eaggb_smem[];
    const int HD = p.H * p.Dh;
    float* gp = reinterpret_cast<float*>(eaggb_smem);             // [HD] gradient row of this node (16-byte aligned)
    int* nb_node = reinterpret_cast<int*>(gp + ((HD + 3) & ~3));  // [N]
    int* nb_row = nb_node + p.N;                                  // [N]
    float* wgt = reinterpret_cast<float*>(nb_row + p.N);          // [H][N] attention weights
    float* sg = wgt + p.H * p.N;                                  // [H][N] mode 0: sigmoid(edge_logit)
    float* tn = sg + p.H * p.N;                                   // [H][N] <g, edge_val + node_val>, later d logit
    float* dsum = tn + p.H * p.N;                                 // [H]
    float* sinv = dsum + p.H;                                     // [H] mode 0: 1 / max(sum, 1e-5)
    float* clamped = sinv + p.H;                                  // [H] mode 0: 1 when the sum was clamped
    __shared__ int s_cnt;
    const long long node = blockIdx.x;
    const long long b = node / p.N;
    const int i = (int)(node - b * p.N);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0) {
        const long long* rev = p.rev + b * p.P;
        int cnt = 0;
        for (int j0 = 0; j0 < p.N; j0 += 32) {
            const int j = j0 + lane;
            long long r = 0;
            if (j < p.N && j != i) r = rev[j < i ? pair_index(j, i, p.N) : pair_index(i, j, p.N)];
            const bool valid = r > 0;
            const unsigned m = __ballot_sync(0xffffffffu, valid);
            if (valid) {
                const int pos = cnt + __popc(m & ((1u << lane) - 1u));
                nb_node[pos] = j;
                nb_row[pos] = (int)(r - 1);
            }
            cnt += __popc(m);
        }
        if (lane == 0) s_cnt = cnt;
    }
    for (int f = threadIdx.x; f < HD; f += kAggThreads) gp[f] = p.g_out[node * HD + f];
    __syncthreads();
    const int cnt = s_cnt;
    if (cnt == 0) return;

    if (p.mode == 1) {
        for (int w = warp; w < cnt * p.H; w += kAggThreads / 32) {
            const int n = w / p.H, h = w - n * p.H;
            const float* q = p.node_q + node * p.ld_q + h * p.Dh;
            const float* k = p.node_k + (b * p.N + nb_node[n]) * p.ld_k + h * p.Dh;
            float acc = 0.f;
            for (int d = lane; d < p.Dh; d += 32) acc = fmaf(q[d], k[d], acc);
            acc = warp_sum(acc);
            if (lane == 0) wgt[h * p.N + n] = fmaf(acc, p.scale, p.edge_logit[(long long)nb_row[n] * p.ld_el + h]);
        }
        __syncthreads();
        for (int h = warp; h < p.H; h += kAggThreads / 32) {
            float mx = -3.0e38f;
            for (int n = lane; n < cnt; n += 32) mx = fmaxf(mx, wgt[h * p.N + n]);
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
            float sum = 0.f;
            for (int n = lane; n < cnt; n += 32) {
                const float e = expf(wgt[h * p.N + n] - mx);
                wgt[h * p.N + n] = e;
                sum += e;
            }
            sum = warp_sum(sum);
            const float inv = 1.0f / sum;
            for (int n = lane; n < cnt; n += 32) wgt[h * p.N + n] *= inv;
        }
    } else {
        for (int h = warp; h < p.H; h += kAggThreads / 32) {
            float sum = 0.f;
            for (int n = lane; n < cnt; n += 32) {
                const float s = 1.0f / (1.0f + expf(-p.edge_logit[(long long)nb_row[n] * p.ld_el + h]));
                sg[h * p.N + n] = s;
                sum += s;
            }
            sum = warp_sum(sum);
            const float inv = 1.0f / fmaxf(sum, 1e-5f);
            if (lane == 0) { sinv[h] = inv; clamped[h] = sum < 1e-5f ? 1.0f : 0.0f; }
            for (int n = lane; n < cnt; n += 32) wgt[h * p.N + n] = sg[h * p.N + n] * inv;
        }
    }
    __syncthreads();

    // value gradients (scattered) and t_n = <g_h, edge_val_n + node_val_n>
    for (int w = warp; w < cnt * p.H; w += kAggThreads / 32) {
        const int n = w / p.H, h = w - n * p.H;
        const long long erow = nb_row[n], nrow = b * p.N + nb_node[n];
        const float* ev = p.edge_val + erow * p.ld_ev + h * p.Dh;
        const float* nv = p.node_val + nrow * p.ld_nv + h * p.Dh;
        float* gev = p.g_edge_val + erow * p.ld_gev + h * p.Dh;
        float* gnv = p.g_node_val + nrow * p.ld_gnv + h * p.Dh;
        const float pw = wgt[h * p.N + n];
        float dot = 0.f;
        if (p.vec) {
            for (int d = lane * 4; d < p.Dh; d += 128) {
                const float4 gv = *reinterpret_cast<const float4*>(gp + h * p.Dh + d);
                const float4 e4 = *reinterpret_cast<const float4*>(ev + d);
                const float4 n4 = *reinterpret_cast<const float4*>(nv + d);
                dot = fmaf(gv.x, e4.x + n4.x, fmaf(gv.y, e4.y + n4.y, fmaf(gv.z, e4.z + n4.z, fmaf(gv.w, e4.w + n4.w, dot))));
                red_add4(gev + d, pw * gv.x, pw * gv.y, pw * gv.z, pw * gv.w);
                red_add4(gnv + d, pw * gv.x, pw * gv.y, pw * gv.z, pw * gv.w);
            }
        } else {
            for (int d = lane; d < p.Dh; d += 32) {
                const float gv = gp[h * p.Dh + d];
                dot = fmaf(gv, ev[d] + nv[d], dot);
                atomicAdd(gev + d, pw * gv);
                atomicAdd(gnv + d, pw * gv);
            }
        }
        dot = warp_sum(dot);
        if (lane == 0) tn[h * p.N + n] = dot;
    }
    __syncthreads();
    for (int h = warp; h < p.H; h += kAggThreads / 32) {
        float s = 0.f;
        for (int n = lane; n < cnt; n += 32) s = fmaf(wgt[h * p.N + n], tn[h * p.N + n], s);
        s = warp_sum(s);
        if (lane == 0) dsum[h] = s;
    }
    __syncthreads();

    // gradient of the logits
    for (int idx = threadIdx.x; idx < cnt * p.H; idx += kAggThreads) {
        const int n = idx / p.H, h = idx - n * p.H;
        float gl;
        if (p.mode == 1) {
            gl = wgt[h * p.N + n] * (tn[h * p.N + n] - dsum[h]);
        } else {
            const float s = sg[h * p.N + n];
            const float gs = clamped[h] != 0.f ? tn[h * p.N + n] * 1e5f : (tn[h * p.N + n] - dsum[h]) * sinv[h];
            gl = gs * s * (1.0f - s);
        }
        atomicAdd(p.g_edge_logit + (long long)nb_row[n] * p.ld_gel + h, gl);
        tn[h * p.N + n] = gl;
    }
    if (p.mode != 1) return;
    __syncthreads();
    // logit = scale q_i . k_j + edge bias:  d q_i = scale sum_n gl_n k_n (this node only), d k_n += scale gl_n q_i
    for (int f = threadIdx.x; f < HD; f += kAggThreads) {
        const int h = f / p.Dh;
        float acc = 0.f;
        for (int n = 0; n < cnt; ++n) acc = fmaf(tn[h * p.N + n], p.node_k[(b * p.N + nb_node[n]) * p.ld_k + f], acc);
        atomicAdd(p.g_node_q + node * p.ld_gq + f, acc * p.scale);
    }
    for (int w = warp; w < cnt * p.H; w += kAggThreads / 32) {
        const int n = w / p.H, h = w - n * p.H;
        const float* q = p.node_q + node * p.ld_q + h * p.Dh;
        float* gk = p.g_node_k + (b * p.N + nb_node[n]) * p.ld_gk + h * p.Dh;
        const float c = tn[h * p.N + n] * p.scale;
        if (p.vec) {
            for (int d = lane * 4; d < p.Dh; d += 128) {
                const float4 q4 = *reinterpret_cast<const float4*>(q + d);
                red_add4(gk + d, c * q4.x, c * q4.y, c * q4.z, c * q4.w);
            }
        } else {
            for (int d = lane; d < p.Dh; d += 32) atomicAdd(gk + d, c * q[d]);
        }
    }
}

// ---- node -> edge combine backward: one warp per compact pair row -----------------------------------------------------
struct PairCombineBwdParams {
    const long long* flat; const long long* idx1; const long long* idx2;
    const float* edge_lin; const float* node_lin; const float* g_out;
    float* g_edge_lin; float* g_node_lin;
    long long R, ld_e, ld_n, ld_ge, ld_gn;
    int N, P, He, act;
};

__global__ void __launch_bounds__(256) pair_combine_bwd_kernel(const PairCombineBwdParams p) {
    const int lane = threadIdx.x & 31;
    const long long nwarps = (long long)gridDim.x * 8;
    for (long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); r < p.R; r += nwarps) {
        const long long fp = p.flat[r];
        const long long b = fp / p.P;
        const int pr = (int)(fp - b * p.P);
        const long long r1 = b * p.N + p.idx1[pr], r2 = b * p.N + p.idx2[pr];
        const float* n1 = p.node_lin + r1 * p.ld_n;
        const float* n2 = p.node_lin + r2 * p.ld_n;
        const float* e = p.edge_lin + r * p.ld_e;
        for (int c = lane; c < p.He; c += 32) {
            float g = p.g_out[r * p.He + c];
            if (p.act == 1) g *= gelu_grad(e[c] + (n1[c] + n2[c]));
            p.g_edge_lin[r * p.ld_ge + c] = g;
            atomicAdd(p.g_node_lin + r1 * p.ld_gn + c, g);
            atomicAdd(p.g_node_lin + r2 * p.ld_gn + c, g);
        }
    }
}

}  // namespace
}  // namespace cnf

using namespace cnf;

extern "C" int cnf_gelu(const cnf_gelu_args* a, cnf_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    CNF_REQUIRE(a != nullptr && a->n >= 0, "cnf_gelu: bad args");
    if (a->n == 0) return CNF_OK;
    CNF_REQUIRE(a->x && a->y, "cnf_gelu: null tensor");
    GeluParams p{a->x, a->grad_y, a->y, a->n};
    gelu_kernel<<<capped_grid((a->n + 255) / 256, 8), 256, 0, stream>>>(p);
    return launch_status("gelu_kernel");
}

extern "C" int cnf_layernorm_bwd(const cnf_layernorm_bwd_args* a, cnf_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    CNF_REQUIRE(a != nullptr, "cnf_layernorm_bwd: null args");
    CNF_REQUIRE(a->M >= 0 && a->H >= 1, "cnf_layernorm_bwd: bad shape");
    if (a->M == 0) return CNF_OK;
    CNF_REQUIRE(a->x && a->gamma && a->grad_y && a->grad_x, "cnf_layernorm_bwd: null tensor");
    CNF_REQUIRE((a->grad_gamma == nullptr) == (a->grad_beta == nullptr), "cnf_layernorm_bwd: grad_gamma and grad_beta go together");
    CNF_SUPPORTED(a->H <= 12 * 1024 / 2, "cnf_layernorm_bwd: H <= 6144");
    LnBwdParams p{a->x, a->gamma, a->grad_y, a->grad_x, a->grad_gamma, a->grad_beta, a->M, a->H, a->eps};
    layernorm_bwd_kernel<<<capped_grid((a->M + 7) / 8, 4), 256, 2 * sizeof(float) * (size_t)a->H, stream>>>(p);
    return launch_status("layernorm_bwd_kernel");
}

extern "C" int cnf_skip_gate_bwd(const cnf_skip_gate_bwd_args* a, cnf_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    CNF_REQUIRE(a != nullptr, "cnf_skip_gate_bwd: null args");
    CNF_REQUIRE(a->M >= 0 && a->H >= 1 && a->config >= 0 && a->config <= 2, "cnf_skip_gate_bwd: bad shape / config");
    if (a->M == 0) return CNF_OK;
    CNF_REQUIRE(a->orig && a->skip && a->grad_out && a->grad_orig && a->grad_skip, "cnf_skip_gate_bwd: null tensor");
    GateBwdParams p{a->orig, a->skip, a->grad_out, a->grad_orig, a->grad_skip, a->M, a->H, a->config};
    skip_gate_bwd_kernel<<<capped_grid((a->M * a->H + 255) / 256, 8), 256, 0, stream>>>(p);
    return launch_status("skip_gate_bwd_kernel");
}

extern "C" int cnf_graph_aggregate_bwd(const cnf_graph_aggregate_bwd_args* a, cnf_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    CNF_REQUIRE(a != nullptr, "cnf_graph_aggregate_bwd: null args");
    const cnf_graph_aggregate_args& f = a->fwd;
    CNF_REQUIRE(f.B >= 0 && f.N >= 1 && f.E >= 1 && f.H >= 1 && f.Dh >= 1, "cnf_graph_aggregate_bwd: bad shape");
    CNF_REQUIRE(f.mode == 0 || f.mode == 1, "cnf_graph_aggregate_bwd: mode must be 0 or 1");
    CNF_REQUIRE(f.activation == 0 || f.activation == 1, "cnf_graph_aggregate_bwd: activation must be 0 or 1");
    if (f.B == 0) return CNF_OK;
    CNF_REQUIRE(f.adjacency && f.hr && a->grad_out && a->grad_hr, "cnf_graph_aggregate_bwd: null tensor");
    if (f.mode == 1) CNF_REQUIRE(f.score_s && f.score_r && a->grad_score_s && a->grad_score_r, "cnf_graph_aggregate_bwd: attention mode needs the scores and their gradients");
    else CNF_REQUIRE(f.hs && a->grad_hs && f.H == 1, "cnf_graph_aggregate_bwd: mean mode needs hs, grad_hs and H = 1");
    CNF_SUPPORTED(f.H <= kMaxHeads && f.N <= 4096, "cnf_graph_aggregate_bwd: H <= %d, N <= 4096", kMaxHeads);
    AggBwdParams p{};
    p.adj = reinterpret_cast<const long long*>(f.adjacency);
    p.hs = f.hs; p.hr = f.hr; p.score_s = f.score_s; p.score_r = f.score_r; p.num_neighbours = f.num_neighbours;
    p.g_out = a->grad_out; p.g_hs = a->grad_hs; p.g_hr = a->grad_hr; p.g_ss = a->grad_score_s; p.g_sr = a->grad_score_r;
    p.ld_hs = f.ld_hs; p.ld_hr = f.ld_hr;
    p.ld_ss = f.ld_score_s > 0 ? f.ld_score_s : f.H;
    p.ld_sr = f.ld_score_r > 0 ? f.ld_score_r : (long long)(f.E + 1) * f.H;
    p.ld_ghs = a->ld_grad_hs; p.ld_ghr = a->ld_grad_hr;
    p.ld_gss = a->ld_grad_score_s > 0 ? a->ld_grad_score_s : f.H;
    p.ld_gsr = a->ld_grad_score_r > 0 ? a->ld_grad_score_r : (long long)(f.E + 1) * f.H;
    p.N = f.N; p.E = f.E; p.H = f.H; p.Dh = f.Dh; p.mode = f.mode; p.act = f.activation; p.slope = f.leaky_slope;
    {
        const uintptr_t bits = reinterpret_cast<uintptr_t>(f.hr) | reinterpret_cast<uintptr_t>(a->grad_hr);
        p.vec = ((f.Dh & 3) == 0 && (p.ld_hr & 3) == 0 && (p.ld_ghr & 3) == 0 && (bits & 15) == 0) ? 1 : 0;
    }
    const size_t smem = (size_t)(f.N + 1) * 8 + ((size_t)2 * f.H * (f.N + 1) + (size_t)f.H * f.Dh + 4 + 2 * f.H) * 4;
    CNF_SUPPORTED(smem <= 48 * 1024, "cnf_graph_aggregate_bwd: working set does not fit shared memory");
    graph_aggregate_bwd_kernel<<<(unsigned)(f.B * f.N), kAggThreads, smem, stream>>>(p);
    return launch_status("graph_aggregate_bwd_kernel");
}

extern "C" int cnf_edge_aggregate_bwd(const cnf_edge_aggregate_bwd_args* a, cnf_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    CNF_REQUIRE(a != nullptr, "cnf_edge_aggregate_bwd: null args");
    const cnf_edge_aggregate_args& f = a->fwd;
    CNF_REQUIRE(f.B >= 0 && f.N >= 2 && f.H >= 1 && f.Dh >= 1, "cnf_edge_aggregate_bwd: bad shape");
    CNF_REQUIRE(f.mode == 0 || f.mode == 1, "cnf_edge_aggregate_bwd: mode must be 0 or 1");
    if (f.B == 0 || f.R == 0) return CNF_OK;
    CNF_REQUIRE(f.rev && f.node_val && f.edge_val && f.edge_logit && a->grad_out, "cnf_edge_aggregate_bwd: null tensor");
    CNF_REQUIRE(a->grad_node_val && a->grad_edge_val && a->grad_edge_logit, "cnf_edge_aggregate_bwd: null gradient tensor");
    if (f.mode == 1) CNF_REQUIRE(f.node_q && f.node_k && a->grad_node_q && a->grad_node_k, "cnf_edge_aggregate_bwd: mode 1 needs q / k and their gradients");
    CNF_SUPPORTED(f.H <= kMaxHeads && f.N <= 2048, "cnf_edge_aggregate_bwd: H <= %d, N <= 2048", kMaxHeads);
    EdgeAggBwdParams p{};
    p.rev = reinterpret_cast<const long long*>(f.rev);
    p.node_val = f.node_val; p.node_q = f.node_q; p.node_k = f.node_k; p.edge_val = f.edge_val; p.edge_logit = f.edge_logit;
    p.g_out = a->grad_out;
    p.g_node_val = a->grad_node_val; p.g_node_q = a->grad_node_q; p.g_node_k = a->grad_node_k;
    p.g_edge_val = a->grad_edge_val; p.g_edge_logit = a->grad_edge_logit;
    p.ld_nv = f.ld_node_val; p.ld_q = f.ld_node_q; p.ld_k = f.ld_node_k; p.ld_ev = f.ld_edge_val; p.ld_el = f.ld_edge_logit;
    p.ld_gnv = a->ld_grad_node_val; p.ld_gq = a->ld_grad_node_q; p.ld_gk = a->ld_grad_node_k;
    p.ld_gev = a->ld_grad_edge_val; p.ld_gel = a->ld_grad_edge_logit;
    p.N = f.N; p.P = f.N * (f.N - 1) / 2; p.H = f.H; p.Dh = f.Dh; p.mode = f.mode; p.scale = f.scale;
    {
        uintptr_t bits = reinterpret_cast<uintptr_t>(f.node_val) | reinterpret_cast<uintptr_t>(f.edge_val) |
                         reinterpret_cast<uintptr_t>(a->grad_node_val) | reinterpret_cast<uintptr_t>(a->grad_edge_val);
        long long lds = p.ld_nv | p.ld_ev | p.ld_gnv | p.ld_gev;
        if (f.mode == 1) {
            bits |= reinterpret_cast<uintptr_t>(f.node_q) | reinterpret_cast<uintptr_t>(a->grad_node_k);
            lds |= p.ld_q | p.ld_gk;
        }
        p.vec = ((f.Dh & 3) == 0 && (lds & 3) == 0 && (bits & 15) == 0) ? 1 : 0;
    }
    const size_t smem = (size_t)f.N * 8 + ((size_t)3 * f.H * f.N + (size_t)f.H * f.Dh + 4 + 3 * f.H) * 4;
    CNF_SUPPORTED(smem <= 48 * 1024, "cnf_edge_aggregate_bwd: working set does not fit shared memory");
    edge_aggregate_bwd_kernel<<<(unsigned)(f.B * f.N), kAggThreads, smem, stream>>>(p);
    return launch_status("edge_aggregate_bwd_kernel");
}

extern "C" int cnf_pair_combine_bwd(const cnf_pair_combine_bwd_args* a, cnf_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    CNF_REQUIRE(a != nullptr, "cnf_pair_combine_bwd: null args");
    const cnf_pair_combine_args& f = a->fwd;
    CNF_REQUIRE(f.R >= 0 && f.N >= 2 && f.He >= 1, "cnf_pair_combine_bwd: bad shape");
    if (f.R == 0) return CNF_OK;
    CNF_REQUIRE(f.flat_indices && f.x_indices1 && f.x_indices2 && f.edge_lin && f.node_lin, "cnf_pair_combine_bwd: null tensor");
    CNF_REQUIRE(a->grad_out && a->grad_edge_lin && a->grad_node_lin, "cnf_pair_combine_bwd: null gradient tensor");
    PairCombineBwdParams p{};
    p.flat = reinterpret_cast<const long long*>(f.flat_indices);
    p.idx1 = reinterpret_cast<const long long*>(f.x_indices1);
    p.idx2 = reinterpret_cast<const long long*>(f.x_indices2);
    p.edge_lin = f.edge_lin; p.node_lin = f.node_lin; p.g_out = a->grad_out;
    p.g_edge_lin = a->grad_edge_lin; p.g_node_lin = a->grad_node_lin;
    p.R = f.R; p.ld_e = f.ld_edge; p.ld_n = f.ld_node; p.ld_ge = a->ld_grad_edge; p.ld_gn = a->ld_grad_node;
    p.N = f.N; p.P = f.N * (f.N - 1) / 2; p.He = f.He; p.act = f.activation;
    pair_combine_bwd_kernel<<<capped_grid((f.R + 7) / 8, 8), 256, 0, stream>>>(p);
    return launch_status("pair_combine_bwd_kernel");
}
