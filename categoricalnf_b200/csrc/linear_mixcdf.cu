// K8 + K1 / K2 fused: the coupling network's FINAL dense projection and the logistic-mixture-CDF coupling
// transform in one kernel - the [B,S,C*(2+3K)] network output never exists in HBM.
//
// Reference path replaced (two steps there): the last nn.Linear of the coupling network, e.g.
// layers/networks/graph_layers.py:198-201,775-778, help_layers.py:84-94, followed by
// MixtureCDFCoupling.get_mixt_params + run_with_params (layers/flows/mixture_cdf_layer.py:95-180).
//
//   per position the projection produces, for each TRANSFORMED channel, one record
//   [t, log_s, log_pi x K, mu x K, log_scale x K]; only those rows of the weight are multiplied.
//
//   persistent CTA per SM, contiguous range of 128-position tiles:
//     warp 0        TMA producer: feature tile [128 x 32] and, per transformed channel, the weight rows of
//                   its record [PNP x 32] (PNP = record padded to 16/32/64 rows) into a stage ring
//     warp 1        tcgen05.mma kind::tf32, M=128, N=CT*PNP; accumulator = the records, in tensor memory,
//                   two accumulator stages: the mixture math of tile i overlaps the GEMM of tile i+1
//     warps 2-3     3xTF32 operand split (precision 1) ; warp 2 owns the tensor-memory allocation
//     warps 4-19    16 epilogue warps: TMEM lane = position; warp (q, g) handles lane quadrant q and the
//                   channels j = g, g+4, ... (2 per thread at 8 transformed channels).  The CTA starts with 96
//                   registers per thread; the service warpgroup (warps 0-3) releases down to 32 and the four
//                   epilogue warpgroups grow to 112 (setmaxnreg - the pool is per CTA, so the two sides must
//                   balance), which keeps the element math free of spills.  12 warps at 128 registers: 0.272 ms,
//                   16 at 112: 0.250 ms.  tcgen05.ld pulls the record of one (position, channel) into registers, bias is
//                   added, and the element is transformed exactly like mixcdf_pipe.cu (same mixmath code).
//                   The z tile sits in shared memory (bulk-async load), is updated in place and leaves with
//                   one bulk-async store; ldj through warp sums and ~1 atomic per (warp, sample).
#include <cuda.h>
#include <stdlib.h>

#include "cnf_common.cuh"
#include "mixcdf_math.cuh"
#include "tc_ptx.cuh"

namespace cnf {

int tc_encode_2d(CUtensorMap* map, const void* base, long long inner, long long outer, int box_inner, int box_outer,
                 int atom32);

namespace {
using namespace tc;
using namespace mixmath;

constexpr int kBM = 128;
constexpr int kBK = 32;
constexpr int kABytes = kBM * 128;
constexpr int kStageCols = 256;
#ifndef CNF_FUSED_UNROLL
#define CNF_FUSED_UNROLL 1      // elements of one thread evaluated one after the other (2: interleaved - A/B build knob)
#endif
#ifndef CNF_FUSED_EPI_WARPS
#define CNF_FUSED_EPI_WARPS 16
#endif
// 4 channel groups x 4 lane quadrants.  640 threads start with 96 registers each; the four service warps (one warpgroup)
// then release theirs down to kAuxRegs and the four epilogue warpgroups grow to kEpiRegs (setmaxnreg), so the element
// math keeps the ~110 registers it needs without spilling while 16 instead of 12 warps hide its MUFU latency, and the
// 8 transformed channels split evenly (2 per thread instead of 3/3/2).
constexpr int kEpiWarps = CNF_FUSED_EPI_WARPS;
constexpr int kAuxRegs = 32, kEpiRegs = 112;
// setmaxnreg.inc draws from the CTA's own pool only: what the service warpgroup releases must cover what the epilogue
// warpgroups claim
static_assert(128 * (96 - kAuxRegs) >= 32 * kEpiWarps * (kEpiRegs - 96), "register hand-over does not balance");
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kThreadsFused = 128 + kEpiThreads;
constexpr int kZStages = 3;

struct FusedParams {
    const float* z;
    const float* pad;
    const float* sf;
    const float* msf;
    const float* bias;
    float* z_out;
    float* ldj;
    float* reg_ldj;
    uint32_t* status;
    long long P, ntiles;
    int S, C, c0;
    int k_blocks, stages;
    int s_period;
    unsigned long long cond_s;
    float reg_max, reg_factor;
    int use_reg;
    // optional epilogue: ActNorm + 1x1 convolution of the NEXT flow block on the finished row
    // (activation_normalization.py:35-43, permutation_layers.py:111-121), and the next coupling's masked input
    const float* nx_bias;
    const float* nx_scales;
    const float* nx_w;       // [C,C] row-major, z @ W
    const float* nx_mask;    // [C] mask of the next coupling (1 = conditioner input) or NULL
    float* z_masked_out;     // [P,C] = z_out * nx_mask, or NULL
    int next;                // 1: epilogue enabled
    uint32_t sleep_ns;       // back-off of the service warps' barrier polls (0: spin)
};

constexpr int padded_record(int pn) { return pn <= 16 ? 16 : (pn <= 32 ? 32 : 64); }

__device__ __forceinline__ float rna_tf32(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}
__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_addr(ssrc)), "r"(bytes)
                 : "memory");
}
// mbarrier wait for the SERVICE warps (TMA producer, MMA issuer, operand split): between failed polls the warp sleeps, so
// that its SYNCS / BRA pairs stop competing with the epilogue warps for the issue port and the MIO queue (with the plain
// spin the four service warps issued ~4400 polls per tile: 36 % of all warp instructions, mio_throttle the top stall)
__device__ __forceinline__ void mbar_wait_svc(uint64_t* bar, uint32_t parity, uint32_t sleep_ns) {
    if (sleep_ns == 0) { mbar_wait(bar, parity); return; }
    for (;;) {
        uint32_t ok;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(smem_addr(bar)), "r"(parity)
            : "memory");
        if (ok) return;
        __nanosleep(sleep_ns);
    }
}
// The z tile is handled per lane QUADRANT (32 positions): the four epilogue warps that share a quadrant - one per channel
// group - synchronise among themselves only, on named barrier 1 + q.  A CTA-wide barrier per tile kept all 16 epilogue
// warps in lock step (everybody loads tensor memory, then everybody is in the MUFU-heavy part, ...) and cost 16 % of the
// kernel (0.232 -> 0.195 ms without it, r02 timing experiment).
__device__ __forceinline__ void quad_barrier(int q) { asm volatile("bar.sync %0, 128;" ::"r"(q + 1) : "memory"); }

// record of one (position, channel): PN consecutive tensor-memory columns of this thread's lane
template <int PN>
__device__ __forceinline__ void load_record(uint32_t taddr, float (&rec)[PN]) {
    static_assert(PN == 14 || PN == 26 || PN == 50, "record sizes for K = 4, 8, 16");
    uint32_t a[16];
    tmem_ld16(taddr, a);
    if constexpr (PN == 14) {
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 14; ++i) rec[i] = __uint_as_float(a[i]);
    } else if constexpr (PN == 26) {
        uint32_t b[8], c[2];
        tmem_ld8(taddr + 16u, b);
        tmem_ld2(taddr + 24u, c);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) rec[i] = __uint_as_float(a[i]);
#pragma unroll
        for (int i = 0; i < 8; ++i) rec[16 + i] = __uint_as_float(b[i]);
        rec[24] = __uint_as_float(c[0]);
        rec[25] = __uint_as_float(c[1]);
    } else {
        uint32_t b[16], c[16], d[2];
        tmem_ld16(taddr + 16u, b);
        tmem_ld16(taddr + 32u, c);
        tmem_ld2(taddr + 48u, d);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            rec[i] = __uint_as_float(a[i]);
            rec[16 + i] = __uint_as_float(b[i]);
            rec[32 + i] = __uint_as_float(c[i]);
        }
        rec[48] = __uint_as_float(d[0]);
        rec[49] = __uint_as_float(d[1]);
    }
}

// RESW ("resident weights", in_features <= 32 = one k-block): the weight rows of the transformed channels' records are
// loaded - and for 3xTF32 split into high / low parts - ONCE per CTA and stay in shared memory; the stage ring then carries
// only the 16 KB feature tile (+ its low part).  Per tile that removes 32 KB of TMA traffic and two thirds of the operand
// split work, which the two split warps otherwise do while competing with the epilogue warps for the same issue ports.
template <int KT, int CT, bool REV, bool STRICT, bool RESW>
__global__ void __launch_bounds__(kThreadsFused, 1)
linear_mixcdf_kernel(const __grid_constant__ CUtensorMap tm_h, const __grid_constant__ CUtensorMap tm_w, const FusedParams p) {
    constexpr int PN = 2 + 3 * KT;
    constexpr int PNP = padded_record(PN);
    constexpr int BN = CT * PNP;               // UMMA N: one padded record per transformed channel
    constexpr int kGroups = kEpiWarps / 4;     // channel j of a position is handled by group j % kGroups
    constexpr int kBBytes = BN * 128;
    constexpr int kHalf = RESW ? kABytes : kABytes + kBBytes;      // bytes one TMA fill of a stage brings in
    constexpr int kStageBytes = STRICT ? 2 * kHalf : kHalf;
    constexpr int kWBytes = RESW ? (STRICT ? 2 * kBBytes : kBBytes) : 0;   // resident weight block in front of the ring
    static_assert(BN <= 256 && BN % 16 == 0 && CT % 4 == 0, "tile shape");

    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem_w = smem_dyn + ((1024u - (smem_addr(smem_dyn) & 1023u)) & 1023u);    // [W hi | W lo] (RESW)
    unsigned char* smem = smem_w + kWBytes;                                                   // stage ring
    const int C = p.C;
    const int ztile = kBM * C;                                    // floats per z tile
    float* s_z = reinterpret_cast<float*>(smem + p.stages * kStageBytes);   // [kZStages][128 * C]
    float* s_bias = s_z + kZStages * ztile;                       // [CT * PN]
    float2* s_bnd = reinterpret_cast<float2*>(s_bias + CT * PN + ((CT * PN) & 1));   // [KT][CT]
    float* s_mfac = reinterpret_cast<float*>(s_bnd + CT * KT);    // [CT * KT]
    float* s_scr = s_mfac + CT * KT;                              // [kEpiWarps][PNP] float64-escape scratch
    float2* s_fa = reinterpret_cast<float2*>(s_scr + kEpiWarps * PNP);   // [CT] (e^{sf}, 2 log2e / max(e^{sf},1))
    // next-block epilogue: bias, e^{scales}, W^T (output-major), next mask; two output tiles (+ two masked tiles)
    float* s_nb = reinterpret_cast<float*>(s_fa + CT);
    float* s_ne = s_nb + (p.next ? C : 0);
    float* s_wT = s_ne + (p.next ? C : 0);
    float* s_nm = s_wT + (p.next ? C * C : 0);
    float* s_out = s_nm + (p.next ? C : 0);                       // [2][128 * C]
    float* s_msk = s_out + (p.next ? 2 * ztile : 0);              // [2][128 * C]
    float* s_end = s_msk + ((p.next && p.z_masked_out) ? 2 * ztile : 0);
    uint64_t* full = reinterpret_cast<uint64_t*>(s_end + ((s_end - reinterpret_cast<float*>(smem)) & 1));
    uint64_t* empty = full + p.stages;
    uint64_t* ready = empty + p.stages;
    uint64_t* tmem_full = ready + p.stages;
    uint64_t* tmem_empty = tmem_full + 2;
    uint64_t* zfull = tmem_empty + 2;
    uint64_t* wfull = zfull + 4 * kZStages;    // RESW: the weight block has landed / has been split
    uint64_t* wready = wfull + 1;
    uint64_t* zdone = wready + 1;              // lean mode: every epilogue warp has finished z stage b (outputs in place)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(zdone + kZStages);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long t0 = (p.ntiles * (long long)blockIdx.x) / gridDim.x;
    const int tiles = (int)((p.ntiles * (long long)(blockIdx.x + 1)) / gridDim.x - t0);

    // Lean mode (resident weights, no next-block epilogue): the z tile is loaded and stored by a SERVICE thread (warp 3,
    // lane 0; the operand split is then warp 2 alone), which waits on the mbarrier `zdone` that the 16 epilogue warps arrive
    // on when their in-place results are in the tile - so the epilogue warps never wait for each other.  (Measured r02:
    // CTA-wide barrier 0.232 ms, barrier per lane quadrant 0.222 ms, outputs by scattered global stores 0.248 ms.)
    constexpr bool lean = RESW;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_h);
        tma_prefetch_desc(&tm_w);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
            mbar_init(&ready[s], RESW ? 32 : 64);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tmem_full[s], 1);
            mbar_init(&tmem_empty[s], kEpiWarps);
        }
        for (int s = 0; s < 4 * kZStages; ++s) mbar_init(&zfull[s], 1);      // [stage][quadrant]
        mbar_init(wfull, 1);
        mbar_init(wready, RESW ? 32 : 64);
        for (int s = 0; s < kZStages; ++s) mbar_init(&zdone[s], kEpiWarps);
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 512);
    for (int i = tid; i < CT * PN; i += kThreadsFused) s_bias[i] = p.bias ? p.bias[p.c0 * PN + i] : 0.f;
    for (int i = tid; i < CT; i += kThreadsFused) {
        const float fac = p.sf ? expf(p.sf[p.c0 + i]) : 1.0f;   // tanh bound of log_s (mixture_cdf_layer.py:157-159)
        s_fa[i] = make_float2(fac, 2.0f * kLog2e / fmaxf(fac, 1.0f));
    }
    for (int i = tid; i < CT * KT; i += kThreadsFused) {
        const float mf = p.msf ? expf(p.msf[(p.c0 + i / KT) * KT + i % KT]) : 1.0f;
        s_mfac[i] = mf;
        s_bnd[(i % KT) * CT + i / KT] = make_float2(2.0f * kLog2e / fmaxf(mf, 1.0f), -mf * kLog2e);
    }
    if (p.next) {
        for (int i = tid; i < C; i += kThreadsFused) {
            s_nb[i] = p.nx_bias ? p.nx_bias[i] : 0.f;
            s_ne[i] = p.nx_scales ? expf(p.nx_scales[i]) : 1.0f;
            s_nm[i] = p.nx_mask ? p.nx_mask[i] : 1.0f;
        }
        for (int i = tid; i < C * C; i += kThreadsFused) {
            const int c = i / C, o = i - c * C;   // W[c][o]
            s_wT[o * C + c] = p.nx_w ? p.nx_w[i] : (c == o ? 1.0f : 0.f);
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (warp < 4) {
    // register hand-over between warpgroups: 128 * (96 - 32) = 512 * (112 - 96)
    if constexpr (kEpiWarps == 16) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kAuxRegs));
    if (warp == 0) {
        // ---------------- TMA producer --------------------------------------------------------------
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            if constexpr (RESW) {      // the whole weight block, once
                mbar_arrive_expect_tx(wfull, (uint32_t)kBBytes);
#pragma unroll
                for (int j = 0; j < CT; ++j) tma_load_2d(smem_w + j * (PNP * 128), &tm_w, wfull, 0, (p.c0 + j) * PN);
            }
            for (int it = 0; it < tiles; ++it) {
                const int m0 = (int)((t0 + it) * kBM);
                for (int kb = 0; kb < p.k_blocks; ++kb) {
                    mbar_wait_svc(&empty[stage], phase ^ 1u, p.sleep_ns);
                    unsigned char* sa = smem + stage * kStageBytes;
                    mbar_arrive_expect_tx(&full[stage], (uint32_t)kHalf);
                    tma_load_2d(sa, &tm_h, &full[stage], kb * kBK, m0);
                    if constexpr (!RESW) {
#pragma unroll
                        for (int j = 0; j < CT; ++j)   // weight rows of channel c0+j's record (+ padding rows, ignored)
                            tma_load_2d(sa + kABytes + j * (PNP * 128), &tm_w, &full[stage], kb * kBK, (p.c0 + j) * PN);
                    }
                    if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer ----------------------------------------------------------------
        if (lane == 0) {
            constexpr uint32_t idesc = idesc_tf32(kBM, BN);
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            if constexpr (RESW) mbar_wait(STRICT ? wready : wfull, 0u);
            for (int it = 0; it < tiles; ++it) {
                mbar_wait_svc(&tmem_empty[acc], acc_phase ^ 1u, p.sleep_ns);
                tcgen05_fence_after();
                const uint32_t d = tmem_base + (uint32_t)(acc * kStageCols);
                for (int kb = 0; kb < p.k_blocks; ++kb) {
                    mbar_wait_svc(STRICT ? &ready[stage] : &full[stage], phase, p.sleep_ns);
                    tcgen05_fence_after();
                    unsigned char* sa = smem + stage * kStageBytes;
                    const unsigned char* sb = RESW ? smem_w : sa + kABytes;
                    const unsigned char* sbl = RESW ? smem_w + kBBytes : sa + kHalf + kABytes;
                    const uint64_t da = smem_desc_k128(sa), db = smem_desc_k128(sb);
#pragma unroll
                    for (int k = 0; k < kBK / 8; ++k) {
                        const uint32_t off = (uint32_t)k * 32u;
                        mma_tf32(d, smem_desc_advance(da, off), smem_desc_advance(db, off), idesc, kb > 0 || k > 0);
                        if constexpr (STRICT) {
                            const uint64_t dal = smem_desc_k128(sa + kHalf), dbl = smem_desc_k128(sbl);
                            mma_tf32(d, smem_desc_advance(dal, off), smem_desc_advance(db, off), idesc, true);
                            mma_tf32(d, smem_desc_advance(da, off), smem_desc_advance(dbl, off), idesc, true);
                        }
                    }
                    mma_commit(&empty[stage]);
                    if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                }
                mma_commit(&tmem_full[acc]);
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1u;
            }
        }
    } else {
        if (lean && warp == 3) {
            // ---------------- z manager (lean mode): tile rows in, finished rows out ---------------------
            if (lane == 0) {
                auto z_load = [&](int it) {
                    const long long pos0 = (t0 + it) * kBM;
                    const int rows = (int)min((long long)kBM, p.P - pos0);
                    const int b = it % kZStages;
                    mbar_arrive_expect_tx(&zfull[b * 4], (uint32_t)(rows * C * 4));
                    bulk_load(s_z + b * ztile, p.z + pos0 * C, (uint32_t)(rows * C * 4), &zfull[b * 4]);
                };
                if (tiles > 0) z_load(0);
                if (tiles > 1) z_load(1);
                for (int it = 0; it < tiles; ++it) {
                    const int b = it % kZStages;
                    mbar_wait_svc(&zdone[b], (uint32_t)((it / kZStages) & 1), p.sleep_ns);     // all 16 epilogue warps are through
                    const long long pos0 = (t0 + it) * kBM;
                    const int rows = (int)min((long long)kBM, p.P - pos0);
                    bulk_store(p.z_out + pos0 * C, s_z + b * ztile, (uint32_t)(rows * C * 4));
                    tma_store_commit();
                    if (it + 2 < tiles) {
                        tma_store_wait_read<1>();      // the store of tile it-1 has drained buffer (it+2) % 3
                        z_load(it + 2);
                    }
                }
                tma_store_wait<0>();
            }
        } else if constexpr (STRICT) {
        // ---------------- 3xTF32 split (see linear_tc.cu); lean mode: warp 2 alone ---------------------
            constexpr int kSplit = RESW ? 32 : 64;
            const int tt = tid - 64;
            int stage = 0;
            uint32_t phase = 0;
            if constexpr (RESW) {      // split the resident weight block once
                mbar_wait(wfull, 0u);
                float4* hi = reinterpret_cast<float4*>(smem_w);
                float4* lo = reinterpret_cast<float4*>(smem_w + kBBytes);
                for (int i = tt; i < (kBBytes >> 4); i += kSplit) {
                    const float4 x = hi[i];
                    float4 h, l;
                    h.x = rna_tf32(x.x); h.y = rna_tf32(x.y); h.z = rna_tf32(x.z); h.w = rna_tf32(x.w);
                    l.x = rna_tf32(x.x - h.x); l.y = rna_tf32(x.y - h.y); l.z = rna_tf32(x.z - h.z); l.w = rna_tf32(x.w - h.w);
                    hi[i] = h;
                    lo[i] = l;
                }
                fence_proxy_async_smem();
                mbar_arrive(wready);
            }
            for (int it = 0; it < tiles; ++it) {
                for (int kb = 0; kb < p.k_blocks; ++kb) {
                    mbar_wait_svc(&full[stage], phase, p.sleep_ns);
                    float4* hi = reinterpret_cast<float4*>(smem + stage * kStageBytes);
                    float4* lo = reinterpret_cast<float4*>(smem + stage * kStageBytes + kHalf);
                    for (int i = tt; i < (kHalf >> 4); i += kSplit) {
                        const float4 x = hi[i];
                        float4 h, l;
                        h.x = rna_tf32(x.x); h.y = rna_tf32(x.y); h.z = rna_tf32(x.z); h.w = rna_tf32(x.w);
                        l.x = rna_tf32(x.x - h.x); l.y = rna_tf32(x.y - h.y); l.z = rna_tf32(x.z - h.z); l.w = rna_tf32(x.w - h.w);
                        hi[i] = h;
                        lo[i] = l;
                    }
                    fence_proxy_async_smem();
                    mbar_arrive(&ready[stage]);
                    if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    }
    } else {
        // ---------------- epilogue: thread = (position row, channel group g) ------------------------
        if constexpr (kEpiWarps == 16) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kEpiRegs));
        const int ew = warp - 4;
        const int q = ew & 3, g = ew >> 2;
        const int row = q * 32 + lane;
        const bool manager = (g == 0 && lane == 0);      // one per quadrant: moves the quadrant's 32 z rows in and out
        const bool use_reg = p.use_reg != 0;
        float* scr = s_scr + ew * PNP;
        const int qoff = q * 32 * C;                     // the quadrant's rows inside a z / output tile (floats)

        // rows of tile `it` that belong to this quadrant (0 on the ragged end: nothing is loaded, waited for or stored)
        auto quad_rows = [&](int it) {
            const long long left = p.P - ((t0 + it) * kBM + q * 32);
            return (int)(left < 0 ? 0 : (left > 32 ? 32 : left));
        };
        auto z_load = [&](int it) {   // manager only
            const int rows = quad_rows(it);
            if (rows == 0) return;
            const long long posq = (t0 + it) * kBM + q * 32;
            const int b = it % kZStages;
            mbar_arrive_expect_tx(&zfull[b * 4 + q], (uint32_t)(rows * C * 4));
            bulk_load(s_z + b * ztile + qoff, p.z + posq * C, (uint32_t)(rows * C * 4), &zfull[b * 4 + q]);
        };
        if (manager && !lean) {
            if (tiles > 0) z_load(0);
            if (tiles > 1) z_load(1);
        }

        long long cur_b = -1;
        float acc_ldj = 0.f, acc_reg = 0.f;
        auto flush = [&]() {
            if (lane == 0 && cur_b >= 0) {
                atomicAdd(p.ldj + cur_b, acc_ldj);
                if (use_reg && p.reg_ldj) atomicAdd(p.reg_ldj + cur_b, acc_reg);
            }
        };

        int acc = 0;
        uint32_t acc_phase = 0;
        for (int it = 0; it < tiles; ++it) {
            const long long pos0 = (t0 + it) * kBM;
            const long long pos = pos0 + row;
            const bool valid = pos < p.P;
            const int zb = it % kZStages;
            float* zt = s_z + zb * ztile + row * C;
            float padv = 1.0f;
            if (p.pad != nullptr && valid) padv = __ldg(p.pad + pos);
            const long long b_idx = pos / p.S;
            bool active = valid && padv != 0.0f;
            if (p.s_period > 0) {
                const int s_in = (int)(pos - b_idx * p.S);
                if ((p.cond_s >> (s_in % p.s_period)) & 1ull) active = false;
            }
            const int qrows = quad_rows(it);
            if (lean) {
                mbar_wait(&zfull[zb * 4], (uint32_t)((it / kZStages) & 1));      // whole tile, loaded by the z manager
            } else if (qrows > 0) {
                mbar_wait(&zfull[zb * 4 + q], (uint32_t)((it / kZStages) & 1));
            }
            mbar_wait(&tmem_full[acc], acc_phase);
            tcgen05_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * kStageCols);

            float eldj = 0.f, ereg = 0.f;
            constexpr int kUnrollJ = CNF_FUSED_UNROLL;
#pragma unroll kUnrollJ
            for (int j = g; j < CT; j += kGroups) {
                const int ch = p.c0 + j;
                float rec[PN];
                __syncwarp();
#ifdef CNF_EXP_NOTMEMLD      // timing experiment only: what do the tensor-memory loads cost?
#pragma unroll
                for (int i = 0; i < PN; ++i) rec[i] = zt[(ch + i) & 15] * 0.37f;
#else
                load_record<PN>(taddr + (uint32_t)(j * PNP), rec);   // warp-collective: outside the divergent part
#endif
                const float x = zt[ch];
                float out = x;
                if (active) {
                    const float2* bj = reinterpret_cast<const float2*>(s_bias + j * PN);   // PN is even
#ifndef CNF_EXP_NOBIAS       // timing experiment only
#pragma unroll
                    for (int i = 0; i < PN / 2; ++i) {
                        const float2 bv = bj[i];
                        rec[2 * i] += bv.x;
                        rec[2 * i + 1] += bv.y;
                    }
#endif
                    const float2 fa = s_fa[j];
                    MixPrep<KT> P;
                    mix_prepare<KT, CT, REV>(P, rec, s_bnd + j, fa.x, fa.y);
                    ElemResult res;
                    bool slow;
                    InvState<KT> st;
                    if constexpr (!REV) {
                        const MixEval e = mix_eval_p<KT>(x, P);
                        slow = !mix_fast_ok(e);
                        if (!slow) res = mix_forward_fast<KT>(e, P, use_reg, p.reg_max, p.reg_factor);
                    } else {
                        slow = !mix_inverse_fast<KT>(x, P, p.status, st, res);
                    }
                    // rare float64 escape: the record goes through this warp's scratch row, one lane at a time
                    unsigned need = __ballot_sync(__activemask(), slow);
                    const unsigned peers = __activemask();
                    while (need) {
                        const int leader = __ffs(need) - 1;
                        if (lane == leader) {
#pragma unroll
                            for (int i = 0; i < PN; ++i) scr[i] = rec[i];
                            if constexpr (!REV)
                                res = mix_forward_f64(x, scr, s_mfac + j * KT, KT, P.log_s, use_reg, p.reg_max, p.reg_factor);
                            else
                                res = mix_inverse_f64(x, st.x, inv_slow_margin<KT>(st), scr, s_mfac + j * KT, KT, P.log_s, st.lb0, st.ub0);
                        }
                        __syncwarp(peers);
                        need &= need - 1;
                    }
                    out = (padv == 1.0f) ? res.z : fmaf(res.z, padv, x * (1.0f - padv));
                    eldj += res.ldj * padv;
                    ereg += res.reg * padv;
                    if ((res.z != res.z) | (res.ldj != res.ldj))
                        flag(p.status, (res.z != res.z ? CNF_FLAG_NAN_Z : 0u) | (res.ldj != res.ldj ? CNF_FLAG_NAN_LDJ : 0u));
                }
                if (valid) zt[ch] = out * padv;
            }
            // conditioner channels pass through, times pad (mixture_cdf_layer.py:76,137-138)
            if (p.pad != nullptr && valid && padv != 1.0f) {
                for (int c = g; c < C - CT; c += kGroups) {
                    const int cc = (c < p.c0) ? c : c + CT;
                    zt[cc] *= padv;
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1u;

            // ---- ldj: 32 consecutive positions per warp -------------------------------------------
            const long long b_first = __shfl_sync(0xffffffffu, b_idx, 0), b_last = __shfl_sync(0xffffffffu, b_idx, 31);
            const bool all_valid = pos0 + q * 32 + 31 < p.P;
            if (all_valid && b_first == b_last) {
                eldj = warp_sum(eldj);
                if (use_reg) ereg = warp_sum(ereg);
                if (b_first != cur_b) { flush(); cur_b = b_first; acc_ldj = 0.f; acc_reg = 0.f; }
                acc_ldj += eldj;
                acc_reg += ereg;
            } else {
                warp_segmented_atomic_add(p.ldj, b_idx, eldj, valid);
                if (use_reg && p.reg_ldj) warp_segmented_atomic_add(p.reg_ldj, b_idx, ereg, valid);
            }

            // ---- z tile out, next z tile in -------------------------------------------------------
            if (lean) {      // hand the finished part of the tile to the z manager and move on
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&zdone[zb]);
                continue;
            }
            fence_proxy_async_smem();
            quad_barrier(q);
            const long long posq = pos0 + q * 32;
            if (!p.next) {
                if (manager) {
                    if (qrows > 0) bulk_store(p.z_out + posq * C, s_z + zb * ztile + qoff, (uint32_t)(qrows * C * 4));
                    tma_store_commit();
                    if (it + 2 < tiles) {
                        tma_store_wait_read<1>();   // the store of tile it-1 has drained buffer (it+2) % 3
                        z_load(it + 2);
                    }
                }
            } else {
                // next block on the finished row: a = (z + b) e^{s} pad ; y = (a @ W) pad.  Thread (row, g) makes
                // the output channels o = g, g+3, ..; rows are read from the z tile and written to an output tile.
                float* so = s_out + (it & 1) * ztile + row * C;
                float* sk = s_msk + (it & 1) * ztile + row * C;
                if (valid) {
                    for (int o = g; o < C; o += kGroups) {
                        const float* wr = s_wT + o * C;
                        float y = 0.f;
                        for (int c4 = 0; c4 < C; c4 += 4) {
                            const float4 zv = *reinterpret_cast<const float4*>(zt + c4);
                            const float4 nb = *reinterpret_cast<const float4*>(s_nb + c4);
                            const float4 ne = *reinterpret_cast<const float4*>(s_ne + c4);
                            const float4 wv = *reinterpret_cast<const float4*>(wr + c4);
                            y = fmaf((zv.x + nb.x) * ne.x * padv, wv.x, y);
                            y = fmaf((zv.y + nb.y) * ne.y * padv, wv.y, y);
                            y = fmaf((zv.z + nb.z) * ne.z * padv, wv.z, y);
                            y = fmaf((zv.w + nb.w) * ne.w * padv, wv.w, y);
                        }
                        y *= padv;
                        so[o] = y;
                        if (p.z_masked_out) sk[o] = y * s_nm[o];
                    }
                }
                fence_proxy_async_smem();
                quad_barrier(q);
                if (manager) {
                    if (qrows > 0) {
                        bulk_store(p.z_out + posq * C, s_out + (it & 1) * ztile + qoff, (uint32_t)(qrows * C * 4));
                        if (p.z_masked_out)
                            bulk_store(p.z_masked_out + posq * C, s_msk + (it & 1) * ztile + qoff, (uint32_t)(qrows * C * 4));
                    }
                    tma_store_commit();
                    tma_store_wait_read<1>();       // the stores of tile it-1 have drained the other output tile
                    if (it + 2 < tiles) z_load(it + 2);   // buffer (it+2) % 3 was last read before this tile's barriers
                }
            }
        }
        flush();
        if (manager && !lean) tma_store_wait<0>();
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

template <int KT, int CT>
size_t fused_smem(int C, int stages, bool strict, bool resw, int next = 0, int masked = 0) {
    constexpr int PN = 2 + 3 * KT, PNP = padded_record(PN), BN = CT * PNP;
    const size_t stage = (size_t)(kABytes + (resw ? 0 : BN * 128)) * (strict ? 2 : 1);
    size_t f = 1024 + stages * stage + (size_t)kZStages * kBM * C * 4;
    if (resw) f += (size_t)BN * 128 * (strict ? 2 : 1);      // resident weight block (high + low part)
    f += ((size_t)CT * PN + 1 + 3 * (size_t)CT * KT + 1 + (size_t)kEpiWarps * PNP + 2 * CT) * 4;
    f += (3 * (size_t)stages + 6 + 5 * kZStages) * 8 + 256;
    if (next) f += ((size_t)3 * C + (size_t)C * C + (size_t)(masked ? 4 : 2) * kBM * C) * 4;
    return f;
}

template <int KT, int CT, bool REV, bool STRICT, bool RESW>
int launch_fused(const CUtensorMap& tm_h, const CUtensorMap& tm_w, FusedParams p, cudaStream_t stream) {
    constexpr size_t kMaxSmem = 232448;
    const int msk = p.z_masked_out != nullptr;
    int stages = 4;
    while (stages > 1 && fused_smem<KT, CT>(p.C, stages, STRICT, RESW, p.next, msk) > kMaxSmem) --stages;
    const size_t need = fused_smem<KT, CT>(p.C, stages, STRICT, RESW, p.next, msk);
    CNF_SUPPORTED(need <= kMaxSmem, "fused projection + mixture tile does not fit shared memory");
    p.stages = stages;
    const size_t smem = need;
    CNF_CUDA(cudaFuncSetAttribute(linear_mixcdf_kernel<KT, CT, REV, STRICT, RESW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)kMaxSmem));
    long long grid = sm_count();
    if (grid > p.ntiles) grid = p.ntiles;
    linear_mixcdf_kernel<KT, CT, REV, STRICT, RESW><<<(unsigned)grid, kThreadsFused, smem, stream>>>(tm_h, tm_w, p);
    return launch_status("linear_mixcdf_kernel");
}

template <int KT, int CT, bool RESW>
int launch_fused_rs(const CUtensorMap& tm_h, const CUtensorMap& tm_w, const FusedParams& p, int reverse, int strict, cudaStream_t stream) {
    if (reverse) return strict ? launch_fused<KT, CT, true, true, RESW>(tm_h, tm_w, p, stream) : launch_fused<KT, CT, true, false, RESW>(tm_h, tm_w, p, stream);
    return strict ? launch_fused<KT, CT, false, true, RESW>(tm_h, tm_w, p, stream) : launch_fused<KT, CT, false, false, RESW>(tm_h, tm_w, p, stream);
}

template <int KT, int CT>
int launch_fused_kc(const CUtensorMap& tm_h, const CUtensorMap& tm_w, const FusedParams& p, int reverse, int strict, cudaStream_t stream) {
    // one k-block (in_features <= 32) without the next-block epilogue: the weights stay resident in shared memory
    static const bool no_resw = getenv("CNF_B200_FUSED_NORESW") != nullptr;      // A/B switch for profiling
    if (p.k_blocks == 1 && !p.next && !no_resw) return launch_fused_rs<KT, CT, true>(tm_h, tm_w, p, reverse, strict, stream);
    return launch_fused_rs<KT, CT, false>(tm_h, tm_w, p, reverse, strict, stream);
}

bool fused_shape_ok(int K, int Ct) { return (K == 8 && (Ct == 8 || Ct == 4)) || (K == 16 && Ct == 4) || (K == 4 && (Ct == 8 || Ct == 4)); }

int check_fusable(const cnf_linear_mixcdf_args* a, MaskView* mv, const char** why) {
    const cnf_mixcdf_args& m = a->mix;
    *why = nullptr;
    if (m.C < 1 || m.C > CNF_MAX_CHANNELS || m.K < 1) { *why = "bad C / K"; return 0; }
    if (build_mask(m.mask, m.C, mv) != CNF_OK) { *why = "mask"; return 0; }
    if (!mv->contiguous || !fused_shape_ok(m.K, mv->n_t)) { *why = "needs K in {4,8,16} with 4 or 8 contiguous transformed channels"; return 0; }
    if (m.C % 4 != 0 || m.C > 32) { *why = "C must be a multiple of 4, at most 32"; return 0; }
    if (a->H < 1 || a->H % 4 != 0) { *why = "in_features must be a multiple of 4"; return 0; }
    if (m.params_prebounded) { *why = "pre-bounded parameters"; return 0; }
    if ((reinterpret_cast<uintptr_t>(a->features) | reinterpret_cast<uintptr_t>(a->weight) | reinterpret_cast<uintptr_t>(m.z) |
         reinterpret_cast<uintptr_t>(m.z_out)) & 15) { *why = "features / weight / z / z_out must be 16-byte aligned"; return 0; }
    if (m.B * m.S >= (1ll << 31) - 256) { *why = "too many positions for 32-bit TMA coordinates"; return 0; }
    return 1;
}

int run_fused(const cnf_linear_mixcdf_args* a, cnf_stream_t stream_, int reverse) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    CNF_REQUIRE(a != nullptr, "args is NULL");
    const cnf_mixcdf_args& m = a->mix;
    CNF_REQUIRE(m.B >= 0 && m.S >= 0, "negative batch/sequence size");
    CNF_REQUIRE(a->precision == 0 || a->precision == 1, "precision must be 0 (TF32) or 1 (3xTF32)");
    MaskView mv{};
    const char* why = nullptr;
    if (!check_fusable(a, &mv, &why)) return fail(CNF_ERR_UNSUPPORTED, "cnf_linear_mixcdf: %s (query cnf_linear_mixcdf_fusable first)", why);
    if (m.B == 0) return CNF_OK;
    CNF_REQUIRE(m.ldj != nullptr, "ldj is NULL");
    if (!m.accumulate) {
        CNF_CUDA(cudaMemsetAsync(m.ldj, 0, sizeof(float) * (size_t)m.B, stream));
        if (m.reg_ldj) CNF_CUDA(cudaMemsetAsync(m.reg_ldj, 0, sizeof(float) * (size_t)m.B, stream));
    }
    const long long P = m.B * m.S;
    if (P == 0) return CNF_OK;
    CNF_REQUIRE(m.z && m.z_out && a->features && a->weight, "z / z_out / features / weight is NULL");
    const int PN = 2 + 3 * m.K, PNP = padded_record(PN);

    FusedParams p{};
    p.z = m.z; p.pad = m.pad; p.sf = m.scaling_factor; p.msf = m.mixture_scaling_factor; p.bias = a->bias;
    p.z_out = m.z_out; p.ldj = m.ldj; p.reg_ldj = m.reg_ldj; p.status = m.status;
    p.P = P; p.ntiles = (P + kBM - 1) / kBM; p.S = (int)m.S; p.C = m.C; p.c0 = mv.c0;
    p.k_blocks = (a->H + kBK - 1) / kBK;
    p.s_period = mv.s_period; p.cond_s = mv.cond_s;
    p.reg_max = m.reg_max; p.reg_factor = m.reg_factor;
    p.use_reg = (!reverse && m.reg_max > 0.f && m.training) ? 1 : 0;
    p.nx_bias = m.next_actnorm_bias; p.nx_scales = m.next_actnorm_scales; p.nx_w = m.next_conv_weight;
    p.next = (p.nx_bias || p.nx_scales || p.nx_w) ? 1 : 0;
    CNF_SUPPORTED(!(p.next && reverse), "the next-block epilogue exists for the forward direction only");
    CNF_REQUIRE(p.next || (a->next_mask == nullptr && a->z_masked_out == nullptr), "next_mask / z_masked_out need the next-block epilogue");
    p.nx_mask = a->next_mask; p.z_masked_out = a->z_masked_out;
    static const int sleep_ns = getenv("CNF_B200_FUSED_SLEEP_NS") ? atoi(getenv("CNF_B200_FUSED_SLEEP_NS")) : 0;
    p.sleep_ns = (uint32_t)(sleep_ns > 0 ? sleep_ns : 0);
    CNF_REQUIRE(a->z_masked_out == nullptr || (reinterpret_cast<uintptr_t>(a->z_masked_out) & 15) == 0, "z_masked_out must be 16-byte aligned");

    CUtensorMap tm_h, tm_w;
    int rc = tc_encode_2d(&tm_h, a->features, a->H, P, kBK, kBM, 0);
    if (rc != CNF_OK) return rc;
    rc = tc_encode_2d(&tm_w, a->weight, a->H, (long long)m.C * PN, kBK, PNP, 0);
    if (rc != CNF_OK) return rc;
    const int K = m.K, Ct = mv.n_t;
#define CNF_FUSED_CASE(KK, CC) \
    if (K == KK && Ct == CC) return launch_fused_kc<KK, CC>(tm_h, tm_w, p, reverse, a->precision, stream);
    CNF_FUSED_CASE(8, 8)
    CNF_FUSED_CASE(8, 4)
    CNF_FUSED_CASE(16, 4)
    CNF_FUSED_CASE(4, 8)
    CNF_FUSED_CASE(4, 4)
#undef CNF_FUSED_CASE
    return fail(CNF_ERR_UNSUPPORTED, "cnf_linear_mixcdf: K=%d with %d transformed channels is not compiled", K, Ct);
}

}  // namespace
}  // namespace cnf

extern "C" int cnf_linear_mixcdf_fusable(const cnf_linear_mixcdf_args* a) {
    if (a == nullptr) return 0;
    cnf::MaskView mv{};
    const char* why = nullptr;
    return cnf::check_fusable(a, &mv, &why);
}
extern "C" int cnf_linear_mixcdf_fwd(const cnf_linear_mixcdf_args* a, cnf_stream_t stream) { return cnf::run_fused(a, stream, 0); }
extern "C" int cnf_linear_mixcdf_inv(const cnf_linear_mixcdf_args* a, cnf_stream_t stream) { return cnf::run_fused(a, stream, 1); }
