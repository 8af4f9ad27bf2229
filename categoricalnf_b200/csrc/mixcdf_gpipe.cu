// K1 / K2 for ANY number of mixture components on the TMA pipeline: the lane-group arithmetic of the generic kernel
// (mixcdf.cu / mixcdf_math.cuh: G = 2^g adjacent lanes share one element, <= 8 components per lane in registers) fed by the
// persistent, warp-specialised bulk-copy ring of mixcdf_pipe.cu instead of stage-then-compute per CTA.
//
// Reference: MixtureCDFCoupling.get_mixt_params + run_with_params (layers/flows/mixture_cdf_layer.py:95-180).  Shapes this
// kernel exists for: K = 64 - the reference's own language-modelling default (experiments/language_modeling/train.py:79),
// records of 776 bytes per element that no thread can hold - and the graph flows' C = 6 / K = 16 and C = 2 / K = 8 layouts,
// whose transformed run per position is not a multiple of 16 bytes (the compile-time kernel of mixcdf_pipe.cu wants both).
//
//   grid      persistent CTAs (as many per SM as the ring's shared memory allows, at most 3), each owning a contiguous
//             range of position tiles
//   producer  one warp: per position one bulk-async copy (cp.async.bulk ... mbarrier::complete_tx) of the 16-byte-aligned
//             HULL of its transformed-channel parameter run - only transformed channels (+ at most 24 bytes of slack) are
//             read from HBM - plus one copy of the tile's z rows, into a ring of stages guarded by full / empty mbarriers
//   consumers 8 warps; lane group g evaluates elements g, g + 256/G, ... of the tile from shared memory: every lane prepares
//             its components once, sums meet in xor-butterflies; the inverse is the safeguarded Newton iteration shared with
//             the other kernels.  Results go back into the stage's z tile; after one consumer barrier the rows leave as
//             coalesced 16-byte stores and the tile's per-position ldj partials as ~1 atomic per (tile, sample)
#include <stdlib.h>

#include "cnf_common.cuh"
#include "mixcdf_math.cuh"
#include "tc_ptx.cuh"

namespace cnf {
namespace {
using namespace mixmath;
using tc::bulk_load;
using tc::mbar_arrive;
using tc::mbar_arrive_expect_tx;
using tc::mbar_fence_init;
using tc::mbar_init;
using tc::mbar_wait;

constexpr int kCWarps = 8;
constexpr int kCons = kCWarps * 32;
constexpr int kThreadsG = kCons + 32;

struct GPipeParams {
    const float* z;
    const float* nn;
    const float* pad;
    const float* sf;
    const float* msf;
    float* z_out;
    float* ldj;
    float* reg_ldj;
    uint32_t* status;
    long long P, ntiles;
    int S, C, K, PN, Ct, c0;
    int TP;        // positions per tile
    int G;         // lanes per element
    int stages;
    int compact;     // nn_out holds the transformed channels' records only ([P, Ct * PN])
    int whole_rows;  // 1: a tile's parameter rows arrive as ONE bulk copy of whole rows (short transformed runs)
    int mis;       // floats between the 16-byte-aligned hull start and the first transformed record
    int hull;      // floats copied per position (multiple of 4)
    int s_period;
    unsigned long long cond_s;
    float reg_max, reg_factor;
    int use_reg, pre;
};

__device__ __forceinline__ void cons_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(kCons) : "memory"); }

// ---- compile-time lane groups: K = 8 * GT, lane `sub` of a group owns components 8 sub .. 8 sub + 7 ------------------------
// All 32 lanes of a warp run the element loop in lock step (ragged tails are clamped, padded elements computed and
// dropped), so the butterflies use the full-warp mask: xor distances below GT never leave the aligned group.
template <int GT>
struct FullWarpGroup {
    int sub;
    __device__ __forceinline__ float sum(float v) const {
#pragma unroll
        for (int d = GT >> 1; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
        return v;
    }
    __device__ __forceinline__ float max(float v) const {
#pragma unroll
        for (int d = GT >> 1; d > 0; d >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, d));
        return v;
    }
};

// MixPrep of this lane's 8 components.  `rec` = the element's record in shared memory (8-byte aligned), `bnd` = the channel's
// (2 log2e / max(e^{msf},1), -e^{msf} log2e) table [K].  Same arithmetic as mix_prepare (mixcdf_math.cuh), with the softmax
// reference point and the normalisation taken over the whole group.
template <int GT, bool WANT_SPAN>
__device__ __forceinline__ void mix_prepare_lane(MixPrep<8>& P, const float* rec, const float2* bnd, float fac, float a2,
                                                 const FullWarpGroup<GT>& g) {
    constexpr int K = 8 * GT;
    // lane `sub` owns the component PAIRS sub, sub + GT, sub + 2 GT, sub + 3 GT (components 2p, 2p + 1): for a fixed i the
    // lanes of a group read consecutive 8-byte words - no bank conflict inside the group (a lane-contiguous split, 32 bytes
    // apart, cost 4.6 extra wavefronts per load: r02b ncu)
    const float2* lp2 = reinterpret_cast<const float2*>(rec + 2) + g.sub;
    const float2* mu2 = lp2 + K / 2;
    const float2* ms2 = mu2 + K / 2;
    const float4* bn4 = reinterpret_cast<const float4*>(bnd) + g.sub;      // two (a2, -mf log2e) entries per pair
    float lp[8], ms[8];
    float2 bn[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 a = lp2[i * GT], b = mu2[i * GT], c = ms2[i * GT];
        const float4 d = bn4[i * GT];
        lp[2 * i] = a.x; lp[2 * i + 1] = a.y;
        P.mu[2 * i] = b.x; P.mu[2 * i + 1] = b.y;
        ms[2 * i] = c.x; ms[2 * i + 1] = c.y;
        bn[2 * i] = make_float2(d.x, d.y);
        bn[2 * i + 1] = make_float2(d.z, d.w);
    }
    float m = lp[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) m = fmaxf(m, lp[i]);
    const float m_l2 = g.max(m) * kLog2e;
    const f2 one = f2_splat(1.0f), mtwo = f2_splat(-2.0f), l2e = f2_splat(kLog2e), nm = f2_splat(-m_l2);
    f2 W2 = f2_splat(0.f);
    float span = 0.f;
#pragma unroll
    for (int k = 0; k < 8; k += 2) {
        const float2 b0 = bn[k], b1 = bn[k + 1];
        float t0, t1;
        f2_get(f2_mul(f2_make(ms[k], ms[k + 1]), f2_make(b0.x, b1.x)), t0, t1);
        const f2 r = f2_make(rcp(1.0f + ex2(t0)), rcp(1.0f + ex2(t1)));          // tanh = 1 - 2 / (1 + 2^v)
        float n0, n1;
        f2_get(f2_mul(f2_fma(mtwo, r, one), f2_make(b0.y, b1.y)), n0, n1);        // -ls_k * log2(e)
        float e0, e1;
        f2_get(f2_mul(f2_make(ex2(n0), ex2(n1)), l2e), e0, e1);
        P.einv2[k] = e0;
        P.einv2[k + 1] = e1;
        if (WANT_SPAN) span += ex2(-n0) + ex2(-n1);
        float a0, a1;
        f2_get(f2_fma(f2_make(lp[k], lp[k + 1]), l2e, nm), a0, a1);
        P.w[k] = ex2(a0);
        P.w[k + 1] = ex2(a1);
        W2 = f2_add(W2, f2_make(P.w[k], P.w[k + 1]));
    }
    float w0, w1;
    f2_get(W2, w0, w1);
    P.iw = rcp(g.sum(w0 + w1));
    P.span = WANT_SPAN ? g.sum(span) : 0.f;
    P.t = rec[0];
    P.log_s = tanh_from_2log2e(rec[1] * a2) * fac;
}

// GT > 0: K = 8 GT at compile time, full-warp butterflies (forward) - the lean path; GT = 0: run-time K through the
// LaneGroup helpers of the generic kernel.  NC = components per lane (8, or 4 for K <= 4 at GT = 0).
template <int NC, bool REV, int GT, int MINB>
__global__ void __launch_bounds__(kThreadsG, MINB) mixcdf_gpipe_kernel(const GPipeParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int C = p.C, K = p.K, PN = p.PN, Ct = p.Ct, TP = p.TP;
    const int par_stage = TP * p.hull;                   // floats
    const int z_stage = (TP * C + 3) & ~3;
    float* s_par = reinterpret_cast<float*>(smem_raw);   // [stages][TP * hull]
    float* s_z = s_par + p.stages * par_stage;           // [stages][TP * C]
    float2* s_bnd = reinterpret_cast<float2*>(s_z + p.stages * z_stage);   // [Ct * K] (2 log2e / max(e^{msf},1), -e^{msf} log2e);
                                                         // 16-byte aligned (read as float4 pairs): the blocks before it are
                                                         // multiples of 4 floats
    float* s_fac = reinterpret_cast<float*>(s_bnd + Ct * K);   // [Ct] e^{sf}
    float* s_a2 = s_fac + Ct;                            // [Ct]
    float* s_mfac = s_a2 + Ct;                           // [Ct * K]
    float* s_ma2 = s_mfac + Ct * K;                      // [Ct * K]
    float* s_ldj = s_ma2 + Ct * K;                       // [2][TP]  (alternating per tile)
    float* s_reg = s_ldj + 2 * TP;                       // [2][TP]
    uint64_t* full = reinterpret_cast<uint64_t*>(s_reg + 2 * TP);
    uint64_t* empty = full + p.stages;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long t0 = (p.ntiles * (long long)blockIdx.x) / gridDim.x;
    const int tiles = (int)((p.ntiles * (long long)(blockIdx.x + 1)) / gridDim.x - t0);

    if (tid == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kCWarps);
        }
        mbar_fence_init();
    }
    for (int i = tid; i < Ct; i += kThreadsG) {
        const float fac = (p.sf && !p.pre) ? expf(p.sf[p.c0 + i]) : 1.0f;
        s_fac[i] = fac;
        s_a2[i] = 2.0f * kLog2e / fmaxf(fac, 1.0f);
    }
    for (int i = tid; i < Ct * K; i += kThreadsG) {
        const int j = i / K, k = i - j * K;
        const float fac = (p.msf && !p.pre) ? expf(p.msf[(p.c0 + j) * K + k]) : 1.0f;
        s_mfac[i] = fac;
        s_ma2[i] = 2.0f * kLog2e / fmaxf(fac, 1.0f);
        s_bnd[i] = make_float2(2.0f * kLog2e / fmaxf(fac, 1.0f), -fac * kLog2e);
    }
    for (int i = tid; i < 2 * TP; i += kThreadsG) { s_ldj[i] = 0.f; s_reg[i] = 0.f; }
    __syncthreads();

    if (warp == kCWarps) {
        // ---------------- producer warp ------------------------------------------------------------
        int stage = 0;
        uint32_t phase = 0;
        long long pos0 = t0 * TP;
        const long long row = p.compact ? (long long)Ct * PN : (long long)C * PN;
        const long long off = p.compact ? 0 : (long long)p.c0 * PN - p.mis;   // hull start inside a position's row (multiple of 4)
        for (int it = 0; it < tiles; ++it, pos0 += TP) {
            mbar_wait(&empty[stage], phase ^ 1u);
            const int rows = (int)min((long long)TP, p.P - pos0);
            // z rows of the tile: one bulk copy of the 16-byte multiple; a ragged LAST tile (rows * C not a multiple of 4,
            // C = 2 or 6) leaves <= 3 floats that three lanes move by hand before the barrier's release
            const int zfl = rows * C, zbulk = zfl & ~3;
            float* dz = s_z + stage * z_stage;
            if (lane < zfl - zbulk) dz[zbulk + lane] = p.z[pos0 * C + zbulk + lane];
            __syncwarp();
            if (lane == 0) mbar_arrive_expect_tx(&full[stage], (uint32_t)((rows * p.hull + zbulk) * 4));
            __syncwarp();
            float* dpar = s_par + stage * par_stage;
            if (p.whole_rows) {
                // hull = row: the tile's rows are contiguous in memory; chunks of <= 16 KB, one per lane
                const long long total = (long long)rows * p.hull;
                for (long long c = (long long)lane * 4096; c < total; c += 32 * 4096) {
                    const int n = (int)min((long long)4096, total - c);
                    bulk_load(dpar + c, p.nn + pos0 * row + c, (uint32_t)(n * 4), &full[stage]);
                }
            } else {
            for (int r = lane; r < rows; r += 32)
                bulk_load(dpar + r * p.hull, p.nn + (pos0 + r) * row + off, (uint32_t)(p.hull * 4), &full[stage]);
            }
            if (lane == 0 && zbulk > 0) bulk_load(dz, p.z + pos0 * C, (uint32_t)(zbulk * 4), &full[stage]);
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
        return;
    }

    // ---------------- consumer warps: one lane group per (position, transformed channel) -------------
    constexpr int kG = GT > 0 ? GT : 1;
    LaneGroup g;                                 // run-time groups (GT = 0, and the inverse's Newton loop at any GT)
    g.G = GT > 0 ? GT : p.G;
    g.sub = tid & (g.G - 1);
    g.mask = g.G == 32 ? 0xffffffffu : (((1u << g.G) - 1u) << (lane & ~(g.G - 1)));
    const int gshift = 31 - __clz(g.G);
    const int ngroups = kCons >> gshift;
    const float inv_ct = 1.0f / (float)Ct;
    const bool use_reg = p.use_reg != 0;
    FullWarpGroup<kG> fg;
    fg.sub = g.sub;

    int stage = 0;
    uint32_t phase = 0;
    long long pos0 = t0 * TP;
    for (int it = 0; it < tiles; ++it, pos0 += TP) {
        const int rows = (int)min((long long)TP, p.P - pos0);
        const int nelem = rows * Ct;
        float* zt = s_z + stage * z_stage;
        const float* par = s_par + stage * par_stage + p.mis;
        float* l_ldj = s_ldj + (it & 1) * TP;
        float* l_reg = s_reg + (it & 1) * TP;
        mbar_wait(&full[stage], phase);

        if constexpr (GT > 0 && !REV) {
            // ---- forward, K = 8 GT: every lane of the warp takes part in every iteration ----------------------------
            // element e -> (position 2a + b, channel j) with e = ((a Ct + j) << 1) | b: ADJACENT lane groups work on the same
            // channel of two consecutive positions, whose records lie `hull` floats apart (= 16 mod 32 banks at K = 64,
            // Ct = 8) instead of PN apart (= 2 mod 32: overlapping bank ranges)
            const int nloop = ((rows + 1) & ~1) * Ct;
            for (int base = 0; base < nloop; base += ngroups) {
                const int e_raw = base + (tid >> gshift);
                const int e = e_raw < nloop ? e_raw : nloop - 1;
                const int aj = e >> 1;
                const int a = fast_div(aj, inv_ct), j = aj - a * Ct;
                const int r_raw = 2 * a + (e & 1);
                const bool in = e_raw < nloop && r_raw < rows;
                const int r = r_raw < rows ? r_raw : rows - 1;
                const long long pos = pos0 + r;
                bool active = in;
                if (p.s_period > 0) {
                    const int sp = (int)(pos % p.S);
                    if ((p.cond_s >> (sp % p.s_period)) & 1ull) active = false;      // conditioner position
                }
                const float padv = p.pad ? p.pad[pos] : 1.0f;
                if (padv == 0.0f) active = false;                                  // padded: copied through (times 0) below
                const int ch = p.c0 + j;
                const float* rec = par + (size_t)r * p.hull + j * PN;
                const float x = zt[r * C + ch];
                MixPrep<8> P;
                mix_prepare_lane<kG, false>(P, rec, s_bnd + j * K, s_fac[j], s_a2[j], fg);
                MixEval ev = mix_eval_p<8>(x, P);
                ev.F = fg.sum(ev.F);
                ev.G = fg.sum(ev.G);
                ev.f = fg.sum(ev.f);
                if (!active || g.sub != 0) continue;      // no shuffle below this line
                ElemResult res;
                if (mix_fast_ok(ev)) res = mix_forward_fast<8>(ev, P, use_reg, p.reg_max, p.reg_factor);
                else res = mix_forward_f64(x, rec, s_mfac + j * K, K, P.log_s, use_reg, p.reg_max, p.reg_factor);
                zt[r * C + ch] = (padv == 1.0f) ? res.z : fmaf(res.z, padv, x * (1.0f - padv));
                atomicAdd(&l_ldj[r], res.ldj * padv);
                if (use_reg) atomicAdd(&l_reg[r], res.reg * padv);
                if ((res.z != res.z) | (res.ldj != res.ldj))
                    flag(p.status, (res.z != res.z ? CNF_FLAG_NAN_Z : 0u) | (res.ldj != res.ldj ? CNF_FLAG_NAN_LDJ : 0u));
            }
        } else {
        for (int e = tid >> gshift; e < nelem; e += ngroups) {
            const int r = fast_div(e, inv_ct), j = e - r * Ct;
            const long long pos = pos0 + r;
            if (p.s_period > 0) {
                const int s = (int)(pos % p.S);
                if ((p.cond_s >> (s % p.s_period)) & 1ull) continue;      // conditioner position (group-uniform)
            }
            const float padv = p.pad ? p.pad[pos] : 1.0f;
            if (padv == 0.0f) continue;                                    // padded: copied through (times 0) below, no ldj
            const int ch = p.c0 + j;
            ElemCtx c;
            c.rec = par + (size_t)r * p.hull + j * PN;
            c.mfac = s_mfac + j * K;
            c.ma2 = s_ma2 + j * K;
            c.fac = s_fac[j];
            c.a2 = s_a2[j];
            c.K = K;
            c.pre = p.pre != 0;
            const float x = zt[r * C + ch];
            MixPrep<NC> P;
            mix_prepare_g<NC>(P, c, g);
            ElemResult res;
            if constexpr (!REV) {
                const MixEval ev = mix_eval_g<NC>(x, P, g);
                if (mix_fast_ok(ev)) res = mix_forward_fast<NC>(ev, P, use_reg, p.reg_max, p.reg_factor);
                else res = mix_forward_f64(x, c.rec, c.pre ? nullptr : c.mfac, K, P.log_s, use_reg, p.reg_max, p.reg_factor);
            } else {
                InvState<NC> st;
                if (!mix_inverse_g<NC>(x, P, g, g.sub == 0 ? p.status : nullptr, st, res))
                    res = mix_inverse_f64(x, st.x, inv_slow_margin<NC>(st), c.rec, c.pre ? nullptr : c.mfac, K, P.log_s, st.lb0,
                                          st.ub0);
            }
            if (g.sub != 0) continue;   // every lane of the group holds the same result; lane 0 publishes it
            zt[r * C + ch] = (padv == 1.0f) ? res.z : fmaf(res.z, padv, x * (1.0f - padv));
            atomicAdd(&l_ldj[r], res.ldj * padv);
            if (use_reg) atomicAdd(&l_reg[r], res.reg * padv);
            if ((res.z != res.z) | (res.ldj != res.ldj))
                flag(p.status, (res.z != res.z ? CNF_FLAG_NAN_Z : 0u) | (res.ldj != res.ldj ? CNF_FLAG_NAN_LDJ : 0u));
        }
        }
        cons_barrier();      // every result of the tile is in shared memory

        // ---- per-sample ldj: warp-segmented sum over the tile's positions; the partials are cleared for tile it + 2 ----
        for (int r0 = (tid & ~31); r0 < TP; r0 += kCons) {
            const int r = r0 + lane;
            const bool valid = r < rows;
            const long long b = valid ? (pos0 + r) / p.S : 0;
            const float v = (r < TP) ? l_ldj[r] : 0.f;
            warp_segmented_atomic_add(p.ldj, b, valid ? v : 0.f, valid);
            if (use_reg && p.reg_ldj) warp_segmented_atomic_add(p.reg_ldj, b, valid ? l_reg[r] : 0.f, valid);
            if (r < TP) { l_ldj[r] = 0.f; l_reg[r] = 0.f; }
        }
        // ---- z rows out (times pad, mixture_cdf_layer.py:76): coalesced 16-byte stores from the stage ----------------
        {
            float* dst = p.z_out + pos0 * C;
            const int n = rows * C;
            if ((C & 3) == 0 && (reinterpret_cast<uintptr_t>(p.z_out) & 15) == 0) {
                const int n4 = n >> 2;
                const float inv_c4 = 4.0f / (float)C;
                for (int i = tid; i < n4; i += kCons) {
                    float4 v = *reinterpret_cast<const float4*>(zt + 4 * i);
                    if (p.pad) {
                        const float pv = p.pad[pos0 + fast_div(i, inv_c4)];
                        v.x *= pv; v.y *= pv; v.z *= pv; v.w *= pv;
                    }
                    stg_stream4(reinterpret_cast<float4*>(dst) + i, v);
                }
            } else {
                const float inv_c = 1.0f / (float)C;
                for (int i = tid; i < n; i += kCons) {
                    float v = zt[i];
                    if (p.pad) v *= p.pad[pos0 + fast_div(i, inv_c)];
                    dst[i] = v;
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
    }
}

size_t gpipe_smem(const GPipeParams& p) {
    size_t f = (size_t)p.stages * p.TP * p.hull + (size_t)p.stages * ((p.TP * p.C + 3) & ~3);
    f += 2 * (size_t)p.Ct + 4 * (size_t)p.Ct * p.K + 4 * (size_t)p.TP + 2;
    return f * sizeof(float) + 2 * (size_t)p.stages * sizeof(uint64_t) + 16;
}

template <int NC, bool REV, int GT, int MINB = 1>
int launch_gpipe(const GPipeParams& p, cudaStream_t stream) {
    const size_t smem = gpipe_smem(p);
    CNF_CUDA(cudaFuncSetAttribute(mixcdf_gpipe_kernel<NC, REV, GT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // persistent grid = exactly the CTAs that are resident at once (registers and shared memory both count)
    int per_sm = 0;
    CNF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mixcdf_gpipe_kernel<NC, REV, GT, MINB>, kThreadsG, smem));
    if (per_sm > 4) per_sm = 4;
    CNF_SUPPORTED(per_sm >= 1, "lane-group pipeline: a CTA with %zu bytes of shared memory does not fit an SM", smem);
    long long grid = (long long)per_sm * sm_count();
    if (grid > p.ntiles) grid = p.ntiles;
    mixcdf_gpipe_kernel<NC, REV, GT, MINB><<<(unsigned)grid, kThreadsG, smem, stream>>>(p);
    return launch_status(REV ? "mixcdf_gpipe_kernel<inv>" : "mixcdf_gpipe_kernel<fwd>");
}

}  // namespace

// Layouts the lane-group pipeline takes: one contiguous transformed run per position, rows a multiple of 16 bytes (so the
// aligned hull of the run has the same shape at every position and never leaves the row), z rows 16-byte aligned.
// *handled = 0: not eligible, the caller uses the staged generic kernel.
static bool gpipe_plan(const cnf_mixcdf_args* a, const MaskView& mask, GPipeParams* out) {
    static const bool disabled = getenv("CNF_B200_MIXCDF_NOGPIPE") != nullptr;      // A/B switch for profiling
    if (disabled) return false;
    const int K = a->K, C = a->C, Ct = mask.n_t, PN = 2 + 3 * K;
    if (!mask.contiguous || Ct < 1) return false;
    if (a->next_actnorm_bias || a->next_actnorm_scales || a->next_conv_weight) return false;
    if (((long long)(a->nn_compact ? Ct : C) * PN) % 4 != 0) return false;
    if ((reinterpret_cast<uintptr_t>(a->nn_out) & 15) || (reinterpret_cast<uintptr_t>(a->z) & 15)) return false;
    GPipeParams p{};
    p.C = C; p.K = K; p.PN = PN; p.Ct = Ct; p.c0 = mask.c0;
    p.compact = a->nn_compact ? 1 : 0;
    const int L = Ct * PN;
    const int NC = K <= 4 ? 4 : 8;
    int G = 1;
    while (NC * G < K) G <<= 1;
    if (G > 32) return false;
    p.G = G;
    // What one bulk copy brings in per position: the aligned hull of the transformed run - or, when that run is short
    // (< 512 bytes: C = 2 / K = 8 edge flows, 104 of 208 bytes), the WHOLE row, so that a tile is ONE contiguous copy
    // instead of hundreds of tiny ones (the TMA engine is request-bound there; the extra sectors were mostly being
    // fetched anyway, a 104-byte run touches 4-5 of the row's 6.5 sectors).
    p.whole_rows = (L * 4 < 512 || p.compact) ? 1 : 0;      // compact: the row IS the run - one copy per tile
    if (p.compact) {
        p.mis = 0;
        p.hull = L;
    } else if (p.whole_rows) {
        p.mis = mask.c0 * PN;
        p.hull = C * PN;
    } else {
        p.mis = (mask.c0 * PN) & 3;
        p.hull = (p.mis + L + 3) & ~3;
    }
    // tile: a whole number of elements per lane group (1 or 2 rounds), <= ~32 KB of parameters per stage, z tile a
    // multiple of 16 bytes
    const int ngroups = kCons / G;
    // K = 64 (G = 8): two rounds of elements per tile halve the per-tile barrier / ring hand-shake per element and two CTAs
    // with ~96 registers beat three with 72 (measured, r02: 0.461 -> 0.429 ms at B 1024 x S 256); K = 32 and below are
    // fastest with one round and three CTAs
    const size_t cap = (G >= 8 ? 52 : (p.whole_rows ? 48 : 32)) * 1024;
    int best = 0;
    float best_eff = 0.f;
    for (int tp = 4; tp <= 1024; tp += 4) {
        if ((size_t)tp * p.hull * 4 > cap && best) break;
        if ((tp * C) % 4 != 0) continue;
        const int elems = tp * Ct, rounds = (elems + ngroups - 1) / ngroups;
        if (rounds > 2 && best) break;
        const float eff = (float)elems / (float)(rounds * ngroups);      // busy lane groups in the tile's last round
        if (eff >= best_eff - 0.02f) { best = tp; best_eff = eff > best_eff ? eff : best_eff; }
    }
    if (!best) return false;
    p.TP = best;
    if (const char* e = getenv("CNF_GPIPE_TP")) {      // experiment knob
        const int tp = atoi(e);
        if (tp >= 4 && tp % 4 == 0 && (tp * C) % 4 == 0) p.TP = tp;
    }
    // three CTAs per SM (24 consumer warps) matter more than a third stage: the ring of a CTA stays under ~72 KB
    p.stages = 3;
    if (gpipe_smem(p) > 72 * 1024) p.stages = 2;
    if (const char* e = getenv("CNF_GPIPE_STAGES")) {  // experiment knob
        const int st = atoi(e);
        if (st >= 2 && st <= 4) p.stages = st;
    }
    if (gpipe_smem(p) > 200 * 1024) return false;
    *out = p;
    return true;
}

bool mixcdf_gpipe_eligible(const cnf_mixcdf_args* a, const MaskView& mask) {
    GPipeParams p{};
    return gpipe_plan(a, mask, &p);
}

int mixcdf_gpipe_try(const cnf_mixcdf_args* a, const MaskView& mask, int reverse, cudaStream_t stream, int* handled) {
    *handled = 0;
    GPipeParams p{};
    if (!gpipe_plan(a, mask, &p)) return CNF_OK;
    const long long P = a->B * a->S;
    p.z = a->z; p.nn = a->nn_out; p.pad = a->pad; p.sf = a->scaling_factor; p.msf = a->mixture_scaling_factor;
    p.z_out = a->z_out; p.ldj = a->ldj; p.reg_ldj = a->reg_ldj; p.status = a->status;
    p.P = P; p.S = (int)a->S;
    p.s_period = mask.s_period; p.cond_s = mask.cond_s;
    p.reg_max = a->reg_max; p.reg_factor = a->reg_factor;
    p.use_reg = (!reverse && a->reg_max > 0.f && a->training) ? 1 : 0;
    p.pre = a->params_prebounded;
    p.ntiles = (P + p.TP - 1) / p.TP;
    *handled = 1;
    // forward with K = 8 G and bounded (not pre-bounded) parameters: compile-time lane groups, full-warp butterflies
    if (!reverse && !p.pre && p.K == 8 * p.G) {
        switch (p.G) {
            case 1: return launch_gpipe<8, false, 1, 3>(p, stream);
            case 2: return launch_gpipe<8, false, 2, 3>(p, stream);
            case 4: return launch_gpipe<8, false, 4, 3>(p, stream);
            case 8: return launch_gpipe<8, false, 8, 2>(p, stream);
            default: break;
        }
    }
    if (p.K <= 4) return reverse ? launch_gpipe<4, true, 0>(p, stream) : launch_gpipe<4, false, 0>(p, stream);
    return reverse ? launch_gpipe<8, true, 0>(p, stream) : launch_gpipe<8, false, 0>(p, stream);
}

}  // namespace cnf
