// K1 / K2 fast path: persistent, warp-specialised, TMA-fed logistic-mixture-CDF coupling kernel.
//
// Same arithmetic as mixcdf.cu (mixcdf_math.cuh; reference mixture_cdf_layer.py:95-142,145-180),
// different data movement, for the common layout where the transformed channels of a position form
// one 16-byte aligned run of the network output (every channel mask of create_channel_mask with
// even K, chess masks):
//
//   grid      2 persistent CTAs per SM, each owning a contiguous range of position tiles
//   producer  one warp; per tile each lane issues one bulk-async copy (TMA engine,
//             cp.async.bulk ... mbarrier::complete_tx) of a position's parameter run - only the
//             transformed channels are ever read from HBM - plus one copy of the tile's z rows,
//             into a ring of shared-memory stages guarded by full/empty mbarriers
//   consumers 8 warps, one thread per (position, transformed channel): the 2+3K record is pulled
//             from shared memory into registers with 8-byte loads (bank-conflict free for the
//             208-float row pitch), the stage is released immediately, then the element is
//             evaluated in fp32 (MUFU ex2/lg2/rcp) with the float64 escape for extreme tails
//   outputs   z rows are written straight from registers as full 32-byte sectors; the ldj is
//             reduced over a position's channels with shuffles and accumulated per warp across
//             consecutive tiles of the same sample -> about one global atomic per (warp, sample)
// There is no CTA-wide barrier in the steady state.
#include "cnf_common.cuh"
#include "mixcdf_math.cuh"

namespace cnf {
namespace {
using namespace mixmath;

constexpr int kConsumerWarps = 8;
constexpr int kConsumers = kConsumerWarps * 32;
constexpr int kThreadsPipe = kConsumers + 32;

struct PipeParams {
    const float* z;
    const float* nn;
    const float* pad;
    const float* sf;
    const float* msf;
    float* z_out;
    float* ldj;
    float* reg_ldj;
    uint32_t* status;
    const float* nx_bias;    // fused epilogue: ActNorm + 1x1 conv of the next flow block (or NULL)
    const float* nx_scales;
    const float* nx_w;
    long long P;       // positions
    long long ntiles;
    int S, C, c0;      // c0: first transformed channel
    long long nn_row;  // floats per position of nn_out (C * PN, or Ct * PN for the compact layout)
    long long nn_off;  // floats from the start of a position's row to its first transformed record
    int s_period;
    unsigned long long cond_s;
    float reg_max, reg_factor;
    int use_reg;
};

// ---- mbarrier / bulk-copy primitives (sm_90+; SASS: SYNCS.*, UBLKCP) ------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

template <int KT>
constexpr int stages_for() { return KT >= 16 ? 2 : 3; }

// FC > 0: the next block's ActNorm and 1x1 convolution (C = FC channels) are applied to the full
// output row before it is stored (activation_normalization.py:35-43, permutation_layers.py:111-121):
//   a = (z + bias) e^{scales} pad ;  y = (a @ W) pad
// saving two full read+write passes over z per flow block.  Their ldj terms are per-sample constants
// added by the caller (cnf_ldj_axpy).
template <int KT, int CT, bool REV, int FC>
__global__ void __launch_bounds__(kThreadsPipe, 2) mixcdf_pipe_kernel(const PipeParams p) {
    constexpr bool FUSE = FC > 0;
    constexpr int NO = FUSE ? FC / CT : 1;   // output channels per lane in the fused epilogue
    constexpr int kStages = stages_for<KT>();
    constexpr int PN = 2 + 3 * KT;
    constexpr int L = CT * PN;            // parameter floats per position (transformed channels)
    constexpr int TP = kConsumers / CT;   // positions per tile
    constexpr int RW = 32 / CT;           // positions per warp
    static_assert(kConsumers % CT == 0 && 32 % CT == 0, "CT must divide the warp size");
    static_assert(PN % 2 == 0, "records are read with 8-byte loads");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int C = p.C;
    const int zrow = TP * C;                                      // floats of one z tile
    float* s_par = reinterpret_cast<float*>(smem_raw);            // [kStages][TP * L]
    float* s_z = s_par + kStages * TP * L;                        // [kStages][TP * C]
    float2* s_bnd = reinterpret_cast<float2*>(s_z + kStages * zrow);  // [CT * KT] tanh bounds, see mix_prepare
    float* s_mfac = reinterpret_cast<float*>(s_bnd + CT * KT);    // [CT * KT] e^{msf} (float64 escape only)
    uint64_t* full = reinterpret_cast<uint64_t*>(s_mfac + CT * KT + ((CT * KT) & 1));
    uint64_t* empty = full + kStages;
    // fused epilogue tables: bias [FC], e^{scales} [FC], W regrouped so lane j finds its NO columns
    // (j, j+CT, ..) of row c contiguously at (c*CT + j)*NO, and one scratch row block per warp
    float* s_nb = reinterpret_cast<float*>(empty + kStages);
    float* s_ne = s_nb + FC;
    float* s_wt = s_ne + FC;                       // [FC * FC]
    float* s_row = s_wt + FC * FC;                 // [kConsumerWarps][RW * FC]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long t0 = (p.ntiles * (long long)blockIdx.x) / gridDim.x;
    const int tiles = (int)((p.ntiles * (long long)(blockIdx.x + 1)) / gridDim.x - t0);

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kConsumerWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < CT * KT; i += kThreadsPipe) {
        const float mf = p.msf ? expf(p.msf[(p.c0 + i / KT) * KT + i % KT]) : 1.0f;
        s_mfac[i] = mf;
        // [k][j] layout: the CT lanes of a position read consecutive 8-byte words (no bank conflict)
        s_bnd[(i % KT) * CT + i / KT] = make_float2(2.0f * kLog2e / fmaxf(mf, 1.0f), -mf * kLog2e);
    }
    if constexpr (FUSE) {
        for (int i = tid; i < FC; i += kThreadsPipe) {
            s_nb[i] = p.nx_bias ? p.nx_bias[i] : 0.f;
            s_ne[i] = p.nx_scales ? expf(p.nx_scales[i]) : 1.0f;
        }
        for (int i = tid; i < FC * FC; i += kThreadsPipe) {
            const int c = i / FC, o = i % FC;      // W[c][o], z @ W
            const float w = p.nx_w ? p.nx_w[i] : (c == o ? 1.0f : 0.f);
            s_wt[(c * CT + o % CT) * NO + o / CT] = w;
        }
    }
    __syncthreads();

    if (warp == kConsumerWarps) {
        // ---------------- producer warp ------------------------------------------------------------
        int stage = 0;
        uint32_t phase = 0;
        long long pos0 = t0 * TP;
        for (int it = 0; it < tiles; ++it, pos0 += TP) {
            mbar_wait(&empty[stage], phase ^ 1u);
            const int rows = (int)min((long long)TP, p.P - pos0);
            if (lane == 0) mbar_arrive_expect_tx(&full[stage], (uint32_t)(rows * (L + C) * 4));
            __syncwarp();
            float* dpar = s_par + stage * (TP * L);
            for (int r = lane; r < rows; r += 32)
                bulk_g2s(dpar + r * L, p.nn + (pos0 + r) * p.nn_row + p.nn_off, L * 4, &full[stage]);
            if (lane == 0) bulk_g2s(s_z + stage * zrow, p.z + pos0 * C, (uint32_t)(rows * C * 4), &full[stage]);
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        return;
    }

    // ---------------- consumer warps: thread = (position r, transformed channel j) -----------------
    const int r = tid / CT, j = tid % CT;
    const int ch = p.c0 + j;
    const float2* bnd = s_bnd + j;   // component k at bnd[k * CT]
    const float fac = p.sf ? expf(p.sf[ch]) : 1.0f;   // tanh bound of log_s (mixture_cdf_layer.py:157-159)
    const float a2 = 2.0f * kLog2e / fmaxf(fac, 1.0f);
    const bool use_reg = p.use_reg != 0;
    const int rec_off = (r * CT + j) * PN, z_off = r * C;
    constexpr int kMaxCopy = (32 - CT + CT - 1) / CT;   // conditioner channels copied per thread (C <= 32)

    // running per-warp ldj accumulator over consecutive positions of one sample
    long long cur_b = -1;
    float acc = 0.f, acc_reg = 0.f;
    auto flush = [&]() {
        if (lane == 0 && cur_b >= 0) {
            atomicAdd(p.ldj + cur_b, acc);
            if (use_reg && p.reg_ldj) atomicAdd(p.reg_ldj + cur_b, acc_reg);
        }
    };
    // (sample, position in sample) of the current tile's first position
    long long pos0 = t0 * TP;
    long long tb = pos0 / p.S;
    int ts = (int)(pos0 - tb * p.S);
    float* orow = p.z_out + (pos0 + r) * C;
    const float* padp = p.pad ? p.pad + pos0 + r : nullptr;

    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; it < tiles; ++it) {
        const bool full_tile = pos0 + TP <= p.P;
        const bool one_sample = full_tile && ts + TP <= p.S;   // whole tile inside one sample (CTA-uniform)
        const bool valid = full_tile || pos0 + r < p.P;
        mbar_wait(&full[stage], phase);

        // ---- shared memory -> registers, then hand the stage back ---------------------------------
        float rec[PN];
        const float2* src = reinterpret_cast<const float2*>(s_par + stage * (TP * L) + rec_off);
#pragma unroll
        for (int i = 0; i < PN / 2; ++i) {
            const float2 v = src[i];
            rec[2 * i] = v.x;
            rec[2 * i + 1] = v.y;
        }
        const float* zr = s_z + stage * zrow + z_off;
        const float x = zr[ch];
        float cval[kMaxCopy > 0 ? kMaxCopy : 1];
#pragma unroll
        for (int n = 0; n < kMaxCopy; ++n) {
            const int c = j + n * CT;                  // c-th conditioner channel
            const int cc = (c < p.c0) ? c : c + CT;    // skip the transformed run
            cval[n] = (c < C - CT) ? zr[cc] : 0.f;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
        if (++stage == kStages) { stage = 0; phase ^= 1u; }

        // ---- element ------------------------------------------------------------------------------
        float padv = 1.0f;
        if (padp != nullptr && valid) padv = *padp;
        bool active = valid && padv != 0.0f;
        if (p.s_period > 0) {   // chess mask: conditioner positions are copied through
            int s_in = ts + r;
            while (s_in >= p.S) s_in -= p.S;
            if ((p.cond_s >> (s_in % p.s_period)) & 1ull) active = false;
        }
        float out = x, eldj = 0.f, ereg = 0.f;
        if (active) {
            MixPrep<KT> P;
            mix_prepare<KT, CT, REV>(P, rec, bnd, fac, a2);
            ElemResult res;
            if constexpr (!REV) {
                const MixEval e = mix_eval_p<KT>(x, P);
                if (mix_fast_ok(e)) {
                    res = mix_forward_fast<KT>(e, P, use_reg, p.reg_max, p.reg_factor);
                } else {
                    const float* rec_slow = p.nn + (pos0 + r) * p.nn_row + p.nn_off + (long long)j * PN;
                    res = mix_forward_f64(x, rec_slow, s_mfac + j * KT, KT, P.log_s, use_reg, p.reg_max, p.reg_factor);
                }
            } else {
                InvState<KT> st;
                if (!mix_inverse_fast<KT>(x, P, p.status, st, res)) {
                    const float* rec_slow = p.nn + (pos0 + r) * p.nn_row + p.nn_off + (long long)j * PN;
                    res = mix_inverse_f64(x, st.x, inv_slow_margin<KT>(st), rec_slow, s_mfac + j * KT, KT, P.log_s, st.lb0,
                                          st.ub0);
                }
            }
            // z_out = out * change + x * (1 - change), change = pad (mixture_cdf_layer.py:137-138)
            out = (padv == 1.0f) ? res.z : fmaf(res.z, padv, x * (1.0f - padv));
            eldj = res.ldj * padv;
            ereg = res.reg * padv;
            if ((res.z != res.z) | (res.ldj != res.ldj))
                flag(p.status, (res.z != res.z ? CNF_FLAG_NAN_Z : 0u) | (res.ldj != res.ldj ? CNF_FLAG_NAN_LDJ : 0u));
        }
        // ---- z row: transformed channel + conditioner copies, times pad (:76) ----------------------
        if constexpr (!FUSE) {
            if (valid) {
                orow[ch] = out * padv;
#pragma unroll
                for (int n = 0; n < kMaxCopy; ++n) {
                    const int c = j + n * CT;
                    const int cc = (c < p.c0) ? c : c + CT;
                    if (c < C - CT) orow[cc] = cval[n] * padv;
                }
            }
        } else {
            // next block's ActNorm on this lane's channels -> warp scratch row -> 1x1 conv columns
            float* srow = s_row + (warp * RW + (r % RW)) * FC;
            srow[ch] = (out * padv + s_nb[ch]) * s_ne[ch] * padv;
#pragma unroll
            for (int n = 0; n < kMaxCopy; ++n) {
                const int c = j + n * CT;
                const int cc = (c < p.c0) ? c : c + CT;
                if (c < FC - CT) srow[cc] = (cval[n] * padv + s_nb[cc]) * s_ne[cc] * padv;
            }
            __syncwarp();
            float y[NO];
#pragma unroll
            for (int n = 0; n < NO; ++n) y[n] = 0.f;
#pragma unroll
            for (int c4 = 0; c4 < FC; c4 += 4) {
                const float4 a = *reinterpret_cast<const float4*>(srow + c4);
                const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float* wr = s_wt + ((c4 + i) * CT + j) * NO;
#pragma unroll
                    for (int n = 0; n < NO; ++n) y[n] = fmaf(av[i], wr[n], y[n]);
                }
            }
            __syncwarp();
            if (valid) {
#pragma unroll
                for (int n = 0; n < NO; ++n) orow[j + n * CT] = y[n] * padv;
            }
        }
        // ---- ldj -----------------------------------------------------------------------------------
        if (one_sample) {
            eldj = warp_sum(eldj);
            if (use_reg) ereg = warp_sum(ereg);
            if (tb != cur_b) { flush(); cur_b = tb; acc = 0.f; acc_reg = 0.f; }
            acc += eldj;
            acc_reg += ereg;
        } else {
            // tile straddles samples or is ragged: per-position sums, then sequential per-sample accumulation
#pragma unroll
            for (int d = 1; d < CT; d <<= 1) {
                eldj += __shfl_xor_sync(0xffffffffu, eldj, d);
                ereg += __shfl_xor_sync(0xffffffffu, ereg, d);
            }
            for (int q = 0; q < RW; ++q) {
                const float v = __shfl_sync(0xffffffffu, eldj, q * CT);
                const float vr = __shfl_sync(0xffffffffu, ereg, q * CT);
                const long long pq = pos0 + warp * RW + q;
                if (pq >= p.P) break;
                const long long bq = pq / p.S;
                if (bq != cur_b) { flush(); cur_b = bq; acc = 0.f; acc_reg = 0.f; }
                acc += v;
                acc_reg += vr;
            }
        }
        // advance to the next tile
        pos0 += TP;
        ts += TP;
        while (ts >= p.S) { ts -= p.S; ++tb; }
        orow += TP * C;
        if (padp != nullptr) padp += TP;
    }
    flush();
}

template <int KT, int CT, int FC>
size_t pipe_smem(int C) {
    constexpr int PN = 2 + 3 * KT, L = CT * PN, TP = kConsumers / CT, kStages = stages_for<KT>();
    size_t f = (size_t)kStages * TP * L + (size_t)kStages * TP * C + 3 * (size_t)CT * KT + ((CT * KT) & 1);
    f += 2 * (size_t)FC + (size_t)FC * FC + (size_t)kConsumers / CT * FC;   // fused epilogue tables + scratch rows
    return f * sizeof(float) + 2 * kStages * sizeof(uint64_t);
}

template <int KT, int CT, bool REV, int FC = 0>
int launch_pipe(const PipeParams& p, cudaStream_t stream) {
    const size_t smem = pipe_smem<KT, CT, FC>(p.C);
    constexpr int kMaxSmem = 113 * 1024;   // two CTAs per SM
    CNF_SUPPORTED(smem <= (size_t)kMaxSmem, "pipelined mixcdf tile needs %zu bytes of shared memory", smem);
    static thread_local int configured_dev = -1;
    int dev = 0;
    CNF_CUDA(cudaGetDevice(&dev));
    if (configured_dev != dev) {
        CNF_CUDA(cudaFuncSetAttribute(mixcdf_pipe_kernel<KT, CT, REV, FC>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
        configured_dev = dev;
    }
    long long grid = 2ll * sm_count();
    if (grid > p.ntiles) grid = p.ntiles;
    mixcdf_pipe_kernel<KT, CT, REV, FC><<<(unsigned)grid, kThreadsPipe, smem, stream>>>(p);
    return launch_status(REV ? "mixcdf_pipe_kernel<inv>" : "mixcdf_pipe_kernel<fwd>");
}

template <int KT, int CT>
int launch_pipe_dir(const PipeParams& p, int reverse, cudaStream_t stream) {
    return reverse ? launch_pipe<KT, CT, true>(p, stream) : launch_pipe<KT, CT, false>(p, stream);
}

}  // namespace

// Layouts the pipelined kernel takes: compile-time (K, Ct) pair, one contiguous transformed run per
// position, everything 16-byte aligned for the bulk copies.
static bool pipe_eligible(const cnf_mixcdf_args* a, const MaskView& mask) {
    const int K = a->K, C = a->C, Ct = mask.n_t, PN = 2 + 3 * K;
    if (a->params_prebounded) return false;   // explicit-parameter callers: generic kernel
    if (!(K == 8 || K == 4 || K == 16)) return false;
    if (!mask.contiguous || !(Ct == 8 || Ct == 16 || Ct == 4)) return false;
    if (C % 4 != 0 || C > 32) return false;
    if ((Ct * PN) % 4 != 0) return false;
    if (!a->nn_compact && ((mask.c0 * PN) % 4 != 0 || (C * PN) % 4 != 0)) return false;
    if ((reinterpret_cast<uintptr_t>(a->nn_out) & 15) || (reinterpret_cast<uintptr_t>(a->z) & 15)) return false;
    if (K == 16 && Ct == 16) return false;   // two stages of 16 x 800 floats would not fit twice per SM
    return true;
}

bool mixcdf_pipe_eligible(const cnf_mixcdf_args* a, const MaskView& mask) {
    const bool fuse = a->next_actnorm_bias || a->next_actnorm_scales || a->next_conv_weight;
    return pipe_eligible(a, mask) && !(fuse && !(a->K == 8 && mask.n_t == 8 && a->C == 16));
}

// The fused next-block epilogue is compiled for the LM layout (C = 16, Ct = 8, K = 8), forward only.
bool mixcdf_pipe_fusable(const cnf_mixcdf_args* a, const MaskView& mask, int reverse) {
    return !reverse && pipe_eligible(a, mask) && a->K == 8 && mask.n_t == 8 && a->C == 16;
}

// Returns CNF_OK and sets *handled = 1 when the pipelined kernel was launched; *handled = 0 when the
// layout is not eligible and the caller must use the generic kernel.  ldj / reg_ldj have already
// been zeroed (or are accumulated into) by the caller.
int mixcdf_pipe_try(const cnf_mixcdf_args* a, const MaskView& mask, int reverse, cudaStream_t stream, int* handled) {
    *handled = 0;
    if (!pipe_eligible(a, mask)) return CNF_OK;
    const int K = a->K, C = a->C, Ct = mask.n_t;
    const bool fuse = a->next_actnorm_bias || a->next_actnorm_scales || a->next_conv_weight;
    if (fuse && !mixcdf_pipe_fusable(a, mask, reverse)) return CNF_OK;   // caller reports "unsupported"
    const long long P = a->B * a->S;
    PipeParams p{};
    p.z = a->z; p.nn = a->nn_out; p.pad = a->pad; p.sf = a->scaling_factor; p.msf = a->mixture_scaling_factor;
    p.z_out = a->z_out; p.ldj = a->ldj; p.reg_ldj = a->reg_ldj; p.status = a->status;
    p.nx_bias = a->next_actnorm_bias; p.nx_scales = a->next_actnorm_scales; p.nx_w = a->next_conv_weight;
    p.P = P; p.S = (int)a->S; p.C = C; p.c0 = mask.c0; p.s_period = mask.s_period; p.cond_s = mask.cond_s;
    p.nn_row = a->nn_compact ? (long long)Ct * (2 + 3 * K) : (long long)C * (2 + 3 * K);
    p.nn_off = a->nn_compact ? 0 : (long long)mask.c0 * (2 + 3 * K);
    p.reg_max = a->reg_max; p.reg_factor = a->reg_factor;
    p.use_reg = (!reverse && a->reg_max > 0.f && a->training) ? 1 : 0;
    const int TP = kConsumers / Ct;
    p.ntiles = (P + TP - 1) / TP;
    *handled = 1;
    if (fuse) return launch_pipe<8, 8, false, 16>(p, stream);
#define CNF_PIPE_CASE(KK, CC) \
    if (K == KK && Ct == CC) return launch_pipe_dir<KK, CC>(p, reverse, stream);
    CNF_PIPE_CASE(8, 8)
    CNF_PIPE_CASE(8, 16)
    CNF_PIPE_CASE(8, 4)
    CNF_PIPE_CASE(4, 8)
    CNF_PIPE_CASE(4, 16)
    CNF_PIPE_CASE(4, 4)
    CNF_PIPE_CASE(16, 8)
    CNF_PIPE_CASE(16, 4)
#undef CNF_PIPE_CASE
    *handled = 0;
    return CNF_OK;
}

}  // namespace cnf
