// Glue of the graph coupling networks around their dense projections:
//   cnf_layernorm          nn.LayerNorm in front of every projection   (layers/networks/graph_layers.py:22,63,190,313-314)
//   cnf_graph_attn_scores  per-node attention logits hs.a_0 / hr.a_1     (RelationGraphAttention.forward, :92-96)
//   cnf_graph_aggregate    neighbour aggregation straight from the integer adjacency matrix:
//                            mode 0  RelationGraphConv  (:27-50)  out_i = hs_i + sum_j hr[j, e(j,i)] / n_i
//                            mode 1  RelationGraphAttention (:76-154) softmax over {neighbours, self} of
//                                    leaky_relu(hs_attn_i + hr_attn[j, e(i,j)]) then the weighted sum of hr rows
//                          The reference builds one-hot adjacencies [B,N,N,E], pads every node to the batch-wide maximum
//                          degree with masked_select / index_select and materialises [B,N,max_deg,E+1,H,Dh] products; here
//                          one CTA per node walks its adjacency row once and reads only the rows of real neighbours.
//   cnf_skip_gate          GNNSkipConnection (:722-733): residual / gated / highway combination
// All HBM / L2 bound, no host synchronisation (the reference's max_neighbours.item() at :107 is one per layer).
#include "cnf_common.cuh"

namespace cnf {
namespace {

__device__ __forceinline__ float gelu_erf(float v) { return 0.5f * v * (1.0f + erff(v * 0.70710678118654752f)); }

// ---- LayerNorm: one warp per row ---------------------------------------------------------------------------------
struct LnParams {
    const float* x; const float* gamma; const float* beta; float* y;
    long long M; int H; float eps;
};

__global__ void __launch_bounds__(256) layernorm_kernel(const LnParams p) {
    const int lane = threadIdx.x & 31;
    const long long warp = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * 8;
    const bool vec = (p.H & 3) == 0;
    for (long long m = warp; m < p.M; m += nwarps) {
        const float* row = p.x + m * p.H;
        float s = 0.f;
        if (vec) {
            for (int c = lane * 4; c < p.H; c += 128) {
                const float4 v = *reinterpret_cast<const float4*>(row + c);
                s += (v.x + v.y) + (v.z + v.w);
            }
        } else {
            for (int c = lane; c < p.H; c += 32) s += row[c];
        }
        const float mean = warp_sum(s) / (float)p.H;
        float q = 0.f;
        if (vec) {
            for (int c = lane * 4; c < p.H; c += 128) {
                const float4 v = *reinterpret_cast<const float4*>(row + c);
                const float a = v.x - mean, b = v.y - mean, cc = v.z - mean, d = v.w - mean;
                q += (a * a + b * b) + (cc * cc + d * d);
            }
        } else {
            for (int c = lane; c < p.H; c += 32) { const float a = row[c] - mean; q += a * a; }
        }
        const float rstd = rsqrtf(warp_sum(q) / (float)p.H + p.eps);
        float* out = p.y + m * p.H;
        if (vec) {
            for (int c = lane * 4; c < p.H; c += 128) {
                const float4 v = *reinterpret_cast<const float4*>(row + c);
                const float4 g = *reinterpret_cast<const float4*>(p.gamma + c);
                const float4 b = *reinterpret_cast<const float4*>(p.beta + c);
                float4 o;
                o.x = fmaf((v.x - mean) * rstd, g.x, b.x);
                o.y = fmaf((v.y - mean) * rstd, g.y, b.y);
                o.z = fmaf((v.z - mean) * rstd, g.z, b.z);
                o.w = fmaf((v.w - mean) * rstd, g.w, b.w);
                *reinterpret_cast<float4*>(out + c) = o;
            }
        } else {
            for (int c = lane; c < p.H; c += 32) out[c] = fmaf((row[c] - mean) * rstd, p.gamma[c], p.beta[c]);
        }
    }
}

// ---- attention logits: one warp per (node, slot) -----------------------------------------------------------------
// slot < H: hs head -> dot with attn_weight[h,0,:];  slot >= H: hr (e, h) -> dot with attn_weight[h,1,:]
struct ScoreParams {
    const float* hs; const float* hr; const float* aw;
    float* score_s; float* score_r;
    long long M, ld_hs, ld_hr;
    int H, Dh, EH;   // EH = (E+1)*H slots in hr
};

__global__ void __launch_bounds__(256) attn_scores_kernel(const ScoreParams p) {
    const int lane = threadIdx.x & 31;
    const int slots = p.H + p.EH;
    const long long total = p.M * slots;
    const long long nwarps = (long long)gridDim.x * 8;
    for (long long w = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); w < total; w += nwarps) {
        const long long m = w / slots;
        const int slot = (int)(w - m * slots);
        const float* src;
        const float* a;
        if (slot < p.H) {
            src = p.hs + m * p.ld_hs + (long long)slot * p.Dh;
            a = p.aw + (long long)slot * 2 * p.Dh;
        } else {
            const int s2 = slot - p.H;
            src = p.hr + m * p.ld_hr + (long long)s2 * p.Dh;
            a = p.aw + (long long)(s2 % p.H) * 2 * p.Dh + p.Dh;
        }
        float acc = 0.f;
        for (int d = lane; d < p.Dh; d += 32) acc = fmaf(src[d], __ldg(a + d), acc);
        acc = warp_sum(acc);
        if (lane == 0) {
            if (slot < p.H) p.score_s[m * p.H + slot] = acc;
            else p.score_r[m * p.EH + (slot - p.H)] = acc;
        }
    }
}

// ---- neighbour aggregation: one CTA per node ---------------------------------------------------------------------
struct AggParams {
    const long long* adj;            // [B,N,N], 0 = no edge, e in 1..E = edge type
    const float* hs; const float* hr; const float* score_s; const float* score_r; const float* num_neighbours;
    float* out;
    long long ld_hs, ld_hr, ld_ss, ld_sr;
    int N, E, H, Dh, mode, act, vec;
    float slope;
};

constexpr int kAggThreads = 256;
constexpr int kMaxHeads = 16;

__global__ void __launch_bounds__(kAggThreads) graph_aggregate_kernel(const AggParams p) {
    extern __shared__ unsigned char agg_smem[];
    // neighbour list (row index of the neighbour's features, hr column block) and per-head weights
    int* nb_row = reinterpret_cast<int*>(agg_smem);          // [N+1]
    int* nb_e = nb_row + (p.N + 1);                            // [N+1]
    float* wgt = reinterpret_cast<float*>(nb_e + (p.N + 1));   // [H][N+1] (mode 1) or [1] (mode 0)
    __shared__ int s_cnt;

    const long long node = blockIdx.x;                 // b*N + i
    const long long b = node / p.N;
    const int i = (int)(node - b * p.N);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int HD = p.H * p.Dh;

    if (warp == 0) {
        // ordered compaction of the adjacency row (mode 1: row i, neighbours j with adj[i][j]) or column (mode 0:
        // adj[j][i], the reference sums over the first node index, :45) -> ascending j, deterministic summation
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
        if (p.mode == 1 && lane == 0) {   // self-connection = extra edge type E (:99-102)
            nb_row[cnt] = i;
            nb_e[cnt] = p.E;
            cnt += 1;
        }
        if (lane == 0) s_cnt = cnt;
    }
    __syncthreads();
    const int cnt = s_cnt;

    if (p.mode == 1) {
        // softmax over the neighbour list, one warp per head (:143-149)
        for (int h = warp; h < p.H; h += kAggThreads / 32) {
            const float s_self = p.score_s[node * p.ld_ss + h];
            float mx = -3.0e38f;
            for (int n = lane; n < cnt; n += 32) {
                float l = s_self + p.score_r[(b * p.N + nb_row[n]) * p.ld_sr + nb_e[n] * p.H + h];
                l = l > 0.f ? l : l * p.slope;
                wgt[h * (p.N + 1) + n] = l;
                mx = fmaxf(mx, l);
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
            float sum = 0.f;
            for (int n = lane; n < cnt; n += 32) {
                const float e = __expf(wgt[h * (p.N + 1) + n] - mx);
                wgt[h * (p.N + 1) + n] = e;
                sum += e;
            }
            sum = warp_sum(sum);
            const float inv = 1.0f / sum;
            for (int n = lane; n < cnt; n += 32) wgt[h * (p.N + 1) + n] *= inv;
        }
    } else if (threadIdx.x == 0) {
        const float nn = p.num_neighbours != nullptr ? p.num_neighbours[node] : (float)cnt;
        wgt[0] = 1.0f / fmaxf(nn, 1e-5f);                                       // :47
    }
    __syncthreads();

    float* out = p.out + node * HD;
    if (p.vec) {
        for (int f = threadIdx.x * 4; f < HD; f += kAggThreads * 4) {
            const int h = f / p.Dh;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int n = 0; n < cnt; ++n) {
                const float w = p.mode == 1 ? wgt[h * (p.N + 1) + n] : 1.0f;
                const float4 v = *reinterpret_cast<const float4*>(p.hr + (b * p.N + nb_row[n]) * p.ld_hr + (long long)nb_e[n] * HD + f);
                acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y); acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
            }
            if (p.mode == 0) {
                const float inv = wgt[0];
                const float4 s = *reinterpret_cast<const float4*>(p.hs + node * p.ld_hs + f);
                acc.x = fmaf(acc.x, inv, s.x); acc.y = fmaf(acc.y, inv, s.y); acc.z = fmaf(acc.z, inv, s.z); acc.w = fmaf(acc.w, inv, s.w);
            }
            if (p.act == 1) { acc.x = gelu_erf(acc.x); acc.y = gelu_erf(acc.y); acc.z = gelu_erf(acc.z); acc.w = gelu_erf(acc.w); }
            *reinterpret_cast<float4*>(out + f) = acc;
        }
    } else {
        for (int f = threadIdx.x; f < HD; f += kAggThreads) {
            const int h = f / p.Dh;
            float acc = 0.f;
            for (int n = 0; n < cnt; ++n) {
                const float w = p.mode == 1 ? wgt[h * (p.N + 1) + n] : 1.0f;
                acc = fmaf(w, p.hr[(b * p.N + nb_row[n]) * p.ld_hr + (long long)nb_e[n] * HD + f], acc);
            }
            if (p.mode == 0) acc = fmaf(acc, wgt[0], p.hs[node * p.ld_hs + f]);
            if (p.act == 1) acc = gelu_erf(acc);
            out[f] = acc;
        }
    }
}

// ---- skip connection ---------------------------------------------------------------------------------------------
struct GateParams {
    const float* orig; const float* s; float* out;
    long long M; int H, config;
};

__global__ void __launch_bounds__(256) skip_gate_kernel(const GateParams p) {
    const long long total = p.M * p.H;
    const long long stride = (long long)gridDim.x * 256;
    const int ld = p.config == 0 ? p.H : 2 * p.H;
    for (long long idx = (long long)blockIdx.x * 256 + threadIdx.x; idx < total; idx += stride) {
        const long long m = idx / p.H;
        const int c = (int)(idx - m * p.H);
        const float o = p.orig[idx];
        const float val = p.s[m * ld + c];
        float r;
        if (p.config == 0) {
            r = o + val;
        } else {
            const float g = 1.0f / (1.0f + __expf(-p.s[m * ld + p.H + c]));
            r = p.config == 1 ? fmaf(val, g, o) : fmaf(o, 1.0f - g, val * g);
        }
        p.out[idx] = r;
    }
}

inline unsigned capped_grid(long long blocks, int per_sm) {
    const long long cap = (long long)sm_count() * per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (unsigned)blocks;
}

}  // namespace
}  // namespace cnf

extern "C" int cnf_layernorm(const cnf_layernorm_args* a, cnf_stream_t stream_) {
    using namespace cnf;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    CNF_REQUIRE(a != nullptr, "cnf_layernorm: null args");
    CNF_REQUIRE(a->M >= 0 && a->H >= 1, "cnf_layernorm: bad shape M=%lld H=%d", (long long)a->M, a->H);
    if (a->M == 0) return CNF_OK;
    CNF_REQUIRE(a->x && a->gamma && a->beta && a->y, "cnf_layernorm: null tensor");
    if ((a->H & 3) == 0)
        CNF_REQUIRE(((reinterpret_cast<uintptr_t>(a->x) | reinterpret_cast<uintptr_t>(a->y) | reinterpret_cast<uintptr_t>(a->gamma) |
                      reinterpret_cast<uintptr_t>(a->beta)) & 15) == 0, "cnf_layernorm: tensors must be 16-byte aligned");
    LnParams p{a->x, a->gamma, a->beta, a->y, a->M, a->H, a->eps};
    layernorm_kernel<<<capped_grid((a->M + 7) / 8, 8), 256, 0, stream>>>(p);
    return launch_status("layernorm_kernel");
}

extern "C" int cnf_graph_attn_scores(const cnf_graph_attn_scores_args* a, cnf_stream_t stream_) {
    using namespace cnf;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    CNF_REQUIRE(a != nullptr, "cnf_graph_attn_scores: null args");
    CNF_REQUIRE(a->M >= 0 && a->H >= 1 && a->Dh >= 1 && a->E >= 0, "cnf_graph_attn_scores: bad shape");
    if (a->M == 0) return CNF_OK;
    CNF_REQUIRE(a->hs && a->hr && a->attn_weight && a->score_s && a->score_r, "cnf_graph_attn_scores: null tensor");
    ScoreParams p{};
    p.hs = a->hs; p.hr = a->hr; p.aw = a->attn_weight; p.score_s = a->score_s; p.score_r = a->score_r;
    p.M = a->M; p.ld_hs = a->ld_hs; p.ld_hr = a->ld_hr; p.H = a->H; p.Dh = a->Dh; p.EH = (a->E + 1) * a->H;
    const long long warps = a->M * (long long)(p.H + p.EH);
    attn_scores_kernel<<<capped_grid((warps + 7) / 8, 8), 256, 0, stream>>>(p);
    return launch_status("attn_scores_kernel");
}

extern "C" int cnf_graph_aggregate(const cnf_graph_aggregate_args* a, cnf_stream_t stream_) {
    using namespace cnf;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    CNF_REQUIRE(a != nullptr, "cnf_graph_aggregate: null args");
    CNF_REQUIRE(a->B >= 0 && a->N >= 1 && a->E >= 1 && a->H >= 1 && a->Dh >= 1, "cnf_graph_aggregate: bad shape");
    CNF_REQUIRE(a->mode == 0 || a->mode == 1, "cnf_graph_aggregate: mode must be 0 (mean + self) or 1 (attention)");
    CNF_REQUIRE(a->activation == 0 || a->activation == 1, "cnf_graph_aggregate: activation must be 0 or 1 (GELU)");
    if (a->B == 0) return CNF_OK;
    CNF_REQUIRE(a->adjacency && a->hr && a->out, "cnf_graph_aggregate: null tensor");
    if (a->mode == 1) CNF_REQUIRE(a->score_s && a->score_r, "cnf_graph_aggregate: attention mode needs the scores");
    else CNF_REQUIRE(a->hs != nullptr && a->H == 1, "cnf_graph_aggregate: mean mode needs hs and H = 1");
    CNF_SUPPORTED(a->H <= kMaxHeads && a->N <= 4096, "cnf_graph_aggregate: H <= %d, N <= 4096", kMaxHeads);
    CNF_SUPPORTED(a->B * (long long)a->N < (1ll << 31), "cnf_graph_aggregate: B*N too large");
    AggParams p{};
    p.adj = reinterpret_cast<const long long*>(a->adjacency);
    p.hs = a->hs; p.hr = a->hr; p.score_s = a->score_s; p.score_r = a->score_r; p.num_neighbours = a->num_neighbours;
    p.out = a->out; p.ld_hs = a->ld_hs; p.ld_hr = a->ld_hr;
    p.ld_ss = a->ld_score_s > 0 ? a->ld_score_s : a->H;
    p.ld_sr = a->ld_score_r > 0 ? a->ld_score_r : (long long)(a->E + 1) * a->H;
    p.N = a->N; p.E = a->E; p.H = a->H; p.Dh = a->Dh; p.mode = a->mode; p.act = a->activation; p.slope = a->leaky_slope;
    uintptr_t bits = reinterpret_cast<uintptr_t>(a->hr) | reinterpret_cast<uintptr_t>(a->out);
    if (a->mode == 0) bits |= reinterpret_cast<uintptr_t>(a->hs);
    p.vec = ((a->Dh & 3) == 0 && (a->ld_hr & 3) == 0 && (a->mode == 1 || (a->ld_hs & 3) == 0) && (bits & 15) == 0) ? 1 : 0;
    const size_t smem = (size_t)(a->N + 1) * 8 + (size_t)(a->mode == 1 ? a->H * (a->N + 1) : 1) * 4;
    CNF_SUPPORTED(smem <= 48 * 1024, "cnf_graph_aggregate: neighbour list does not fit shared memory");
    graph_aggregate_kernel<<<(unsigned)(a->B * a->N), kAggThreads, smem, stream>>>(p);
    return launch_status("graph_aggregate_kernel");
}

extern "C" int cnf_skip_gate(const cnf_skip_gate_args* a, cnf_stream_t stream_) {
    using namespace cnf;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    CNF_REQUIRE(a != nullptr, "cnf_skip_gate: null args");
    CNF_REQUIRE(a->M >= 0 && a->H >= 1, "cnf_skip_gate: bad shape");
    CNF_REQUIRE(a->config >= 0 && a->config <= 2, "cnf_skip_gate: config must be 0, 1 or 2");
    if (a->M == 0) return CNF_OK;
    CNF_REQUIRE(a->orig && a->skip && a->out, "cnf_skip_gate: null tensor");
    GateParams p{a->orig, a->skip, a->out, a->M, a->H, a->config};
    skip_gate_kernel<<<capped_grid((a->M * a->H + 255) / 256, 8), 256, 0, stream>>>(p);
    return launch_status("skip_gate_kernel");
}

// =====================================================================================================================
// Edge-GNN glue (layers/networks/graph_layers.py:242-336, 388-700): node <-> node-pair message passing of GraphCNF.
// Node pairs (a < b) are numbered node-major, p(a,b) = a (N-1) - a (a-1)/2 + (b-a-1)  (molecule_generation/mutils.py:5-10);
// the features of the VALID pairs are stored compacted ([R, *]), `rev[b*P + p]` = 1 + row of pair p of graph b, 0 = not
// valid (the reference's indices_reverse, graph_layers.py:349-355).
//   cnf_edge_aggregate  Edge2NodeAttnLayer (:595-645 / 647-699)    mode 0: sigmoid weights normalised over the valid pairs
//                       Edge2NodeQKVAttnLayer (:432-502 / 504-558) mode 1: softmax(q_i.k_j * scale + edge bias)
//                       out[i,h,:] = sum_j w_ij (edge_val[pair(i,j),h,:] + node_val[j,h,:]);  one CTA per node walks its
//                       N-1 pairs once.  The reference sorts / pads every pair list to [B,H,N,N-1,*] tensors (full form) or
//                       gathers top-k neighbour lists (sparse form, one .item() sync per layer); both forms reduce to this.
//   cnf_pair_combine    Node2EdgePlainLayer (:317-336): out[r] = GELU(edge_lin[r] + node_lin[a(r)] + node_lin[b(r)])
// =====================================================================================================================
namespace cnf {
namespace {

struct EdgeAggParams {
    const long long* rev;        // [B*P]
    const float* node_val;       // [B*N, >= H*Dh] pitch ld_nv : context features (mode 0) / values (mode 1)
    const float* node_q;         // mode 1: queries [B*N, >= H*Dh] pitch ld_q
    const float* node_k;         // mode 1: keys, pitch ld_k
    const float* edge_val;       // [R, >= H*Dh] pitch ld_ev
    const float* edge_logit;     // [R, >= H] pitch ld_el
    float* out;                  // [B*N, H*Dh]
    long long ld_nv, ld_q, ld_k, ld_ev, ld_el;
    int N, P, H, Dh, mode, vec;
    float scale;
};

__device__ __forceinline__ int pair_index(int a, int b, int N) { return a * (N - 1) - (a * (a - 1)) / 2 + (b - a - 1); }

__global__ void __launch_bounds__(kAggThreads) edge_aggregate_kernel(const EdgeAggParams p) {
    extern __shared__ unsigned char eagg_smem[];
    int* nb_node = reinterpret_cast<int*>(eagg_smem);          // [N]
    int* nb_row = nb_node + p.N;                                 // [N] compact edge row
    float* wgt = reinterpret_cast<float*>(nb_row + p.N);         // [H][N]
    __shared__ int s_cnt;
    const long long node = blockIdx.x;
    const long long b = node / p.N;
    const int i = (int)(node - b * p.N);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int HD = p.H * p.Dh;

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
    __syncthreads();
    const int cnt = s_cnt;

    if (p.mode == 1) {
        // logits: warp per (neighbour, head) dot product q_i . k_j
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
                const float e = __expf(wgt[h * p.N + n] - mx);
                wgt[h * p.N + n] = e;
                sum += e;
            }
            sum = warp_sum(sum);
            const float inv = cnt > 0 ? 1.0f / sum : 0.f;
            for (int n = lane; n < cnt; n += 32) wgt[h * p.N + n] *= inv;
        }
    } else {
        for (int h = warp; h < p.H; h += kAggThreads / 32) {
            float sum = 0.f;
            for (int n = lane; n < cnt; n += 32) {
                const float s = 1.0f / (1.0f + __expf(-p.edge_logit[(long long)nb_row[n] * p.ld_el + h]));
                wgt[h * p.N + n] = s;
                sum += s;
            }
            sum = warp_sum(sum);
            const float inv = 1.0f / fmaxf(sum, 1e-5f);                                  // :629 / :688
            for (int n = lane; n < cnt; n += 32) wgt[h * p.N + n] *= inv;
        }
    }
    __syncthreads();

    float* out = p.out + node * HD;
    if (p.vec) {
        for (int f = threadIdx.x * 4; f < HD; f += kAggThreads * 4) {
            const int h = f / p.Dh;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int n = 0; n < cnt; ++n) {
                const float w = wgt[h * p.N + n];
                const float4 e = *reinterpret_cast<const float4*>(p.edge_val + (long long)nb_row[n] * p.ld_ev + f);
                const float4 v = *reinterpret_cast<const float4*>(p.node_val + (b * p.N + nb_node[n]) * p.ld_nv + f);
                acc.x = fmaf(w, e.x + v.x, acc.x); acc.y = fmaf(w, e.y + v.y, acc.y);
                acc.z = fmaf(w, e.z + v.z, acc.z); acc.w = fmaf(w, e.w + v.w, acc.w);
            }
            *reinterpret_cast<float4*>(out + f) = acc;
        }
    } else {
        for (int f = threadIdx.x; f < HD; f += kAggThreads) {
            const int h = f / p.Dh;
            float acc = 0.f;
            for (int n = 0; n < cnt; ++n)
                acc = fmaf(wgt[h * p.N + n], p.edge_val[(long long)nb_row[n] * p.ld_ev + f] + p.node_val[(b * p.N + nb_node[n]) * p.ld_nv + f], acc);
            out[f] = acc;
        }
    }
}

struct PairCombineParams {
    const long long* flat;       // [R] = b*P + p
    const long long* idx1; const long long* idx2;   // [P] node indices of pair p
    const float* edge_lin;       // [R, He] pitch ld_e
    const float* node_lin;       // [B*N, He] pitch ld_n
    float* out;                  // [R, He]
    long long R, ld_e, ld_n;
    int N, P, He, act;
};

__global__ void __launch_bounds__(256) pair_combine_kernel(const PairCombineParams p) {
    const int lane = threadIdx.x & 31;
    const long long nwarps = (long long)gridDim.x * 8;
    for (long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); r < p.R; r += nwarps) {
        const long long fp = p.flat[r];
        const long long b = fp / p.P;
        const int pr = (int)(fp - b * p.P);
        const float* n1 = p.node_lin + (b * p.N + p.idx1[pr]) * p.ld_n;
        const float* n2 = p.node_lin + (b * p.N + p.idx2[pr]) * p.ld_n;
        const float* e = p.edge_lin + r * p.ld_e;
        float* o = p.out + r * p.He;
        for (int c = lane; c < p.He; c += 32) {
            float v = e[c] + (n1[c] + n2[c]);
            if (p.act == 1) v = gelu_erf(v);
            o[c] = v;
        }
    }
}

}  // namespace
}  // namespace cnf

extern "C" int cnf_edge_aggregate(const cnf_edge_aggregate_args* a, cnf_stream_t stream_) {
    using namespace cnf;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    CNF_REQUIRE(a != nullptr, "cnf_edge_aggregate: null args");
    CNF_REQUIRE(a->B >= 0 && a->N >= 2 && a->H >= 1 && a->Dh >= 1, "cnf_edge_aggregate: bad shape");
    CNF_REQUIRE(a->mode == 0 || a->mode == 1, "cnf_edge_aggregate: mode must be 0 (sigmoid) or 1 (query-key softmax)");
    if (a->B == 0) return CNF_OK;
    CNF_REQUIRE(a->rev && a->node_val && a->out, "cnf_edge_aggregate: null tensor");
    CNF_REQUIRE(a->R == 0 || (a->edge_val && a->edge_logit), "cnf_edge_aggregate: null edge tensor");
    if (a->mode == 1) CNF_REQUIRE(a->node_q && a->node_k, "cnf_edge_aggregate: mode 1 needs queries and keys");
    CNF_SUPPORTED(a->H <= kMaxHeads && a->N <= 2048, "cnf_edge_aggregate: H <= %d, N <= 2048", kMaxHeads);
    EdgeAggParams p{};
    p.rev = reinterpret_cast<const long long*>(a->rev);
    p.node_val = a->node_val; p.node_q = a->node_q; p.node_k = a->node_k; p.edge_val = a->edge_val; p.edge_logit = a->edge_logit;
    p.out = a->out;
    p.ld_nv = a->ld_node_val; p.ld_q = a->ld_node_q; p.ld_k = a->ld_node_k; p.ld_ev = a->ld_edge_val; p.ld_el = a->ld_edge_logit;
    p.N = a->N; p.P = a->N * (a->N - 1) / 2; p.H = a->H; p.Dh = a->Dh; p.mode = a->mode; p.scale = a->scale;
    uintptr_t bits = reinterpret_cast<uintptr_t>(a->node_val) | reinterpret_cast<uintptr_t>(a->edge_val) | reinterpret_cast<uintptr_t>(a->out);
    p.vec = ((a->Dh & 3) == 0 && (p.ld_nv & 3) == 0 && (p.ld_ev & 3) == 0 && (bits & 15) == 0) ? 1 : 0;
    const size_t smem = (size_t)a->N * 8 + (size_t)a->H * a->N * 4;
    CNF_SUPPORTED(smem <= 48 * 1024, "cnf_edge_aggregate: neighbour list does not fit shared memory");
    edge_aggregate_kernel<<<(unsigned)(a->B * a->N), kAggThreads, smem, stream>>>(p);
    return launch_status("edge_aggregate_kernel");
}

extern "C" int cnf_pair_combine(const cnf_pair_combine_args* a, cnf_stream_t stream_) {
    using namespace cnf;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    CNF_REQUIRE(a != nullptr, "cnf_pair_combine: null args");
    CNF_REQUIRE(a->R >= 0 && a->N >= 2 && a->He >= 1, "cnf_pair_combine: bad shape");
    if (a->R == 0) return CNF_OK;
    CNF_REQUIRE(a->flat_indices && a->x_indices1 && a->x_indices2 && a->edge_lin && a->node_lin && a->out, "cnf_pair_combine: null tensor");
    PairCombineParams p{};
    p.flat = reinterpret_cast<const long long*>(a->flat_indices);
    p.idx1 = reinterpret_cast<const long long*>(a->x_indices1);
    p.idx2 = reinterpret_cast<const long long*>(a->x_indices2);
    p.edge_lin = a->edge_lin; p.node_lin = a->node_lin; p.out = a->out;
    p.R = a->R; p.ld_e = a->ld_edge; p.ld_n = a->ld_node; p.N = a->N; p.P = a->N * (a->N - 1) / 2; p.He = a->He; p.act = a->activation;
    pair_combine_kernel<<<capped_grid((a->R + 7) / 8, 8), 256, 0, stream>>>(p);
    return launch_status("pair_combine_kernel");
}
