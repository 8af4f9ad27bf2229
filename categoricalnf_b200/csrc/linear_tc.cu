// K8: dense projection y = x W^T + b (nn.Linear) of the coupling networks on the 5th-generation tensor
// cores (reference call sites layers/networks/graph_layers.py:24-25,64-71,
// 192-202,307-315,402-405,574-577,712-716,766-779 and help_layers.py:57-124).
//
//   persistent grid, one CTA per SM, warp-specialised:
//     warp 0      TMA producer: x tile [128 x 32] and W tile [BN x 32] fp32, 128-byte swizzle, into a ring of
//                 shared-memory stages (cp.async.bulk.tensor, mbarrier complete_tx)
//     warp 1      one elected thread issues tcgen05.mma kind::tf32 (M=128, N=BN<=256, K=8 per instruction);
//                 accumulators live in tensor memory, two accumulator stages (2 x 256 columns) so the epilogue
//                 of tile i overlaps the main loop of tile i+1
//     warp 2      tensor-memory allocation
//     warps 4-7   epilogue: tcgen05.ld (thread = output row) -> + bias (+ GELU) -> swizzled staging tile in
//                 shared memory -> TMA store (clips the M / N tails); or guarded direct stores
//     warps 8-11  (3xTF32 only) write lo = x - trunc_tf32(x) next to every landed stage (the tensor core truncates
//                 fp32 operands to TF32 itself, so the landed tile serves as the high part unchanged)
//
//   precision 0  one TF32 pass (operands rounded to 10 mantissa bits: ~5e-4 relative per product)
//   precision 1  3xTF32: x W^T ~= hi hi + lo hi + hi lo, fp32 accumulation in tensor memory - error ~1e-6,
//                inside the 1e-4 parity budget of the flow transforms fed by this projection
#include <cuda.h>
#include <stdlib.h>

#include "cnf_common.cuh"
#include "tc_ptx.cuh"

namespace cnf {
namespace {
using namespace tc;

constexpr int kBM = 128;           // rows per tile (UMMA M)
constexpr int kBK = 32;            // fp32 per 128-byte swizzle row
constexpr int kABytes = kBM * 128;
constexpr int kStageCols = 256;    // tensor-memory columns per accumulator stage
constexpr int kEpiWarps = 4;
constexpr int kStagingBytes = kEpiWarps * 2 * 4096;

struct LinearParams {
    const float* bias;
    float* y;
    long long M;
    int N, K;
    int block_n, n_tiles, k_blocks, stages;
    long long m_tiles, tiles;   // tiles = splits * m_tiles * n_tiles
    int kb_per_split;           // split-K: tile t reduces k-blocks [split * kb_per_split, +kb_per_split) (grad_weight)
    int a_mn, b_mn;             // operand layout in memory: 0 = reduction dim contiguous (K-major), 1 = MN-major
    int act;        // 0 none, 1 GELU (erf)
    int w_lo;       // 3xTF32: the low part of the B operand is loaded from global memory (tm_blo) instead of being split here
    int store;      // 0: guarded direct stores; 1: shared memory + TMA store; 2: red.global.add (split-K partial sums)
};

struct TileCoord {
    long long m0;
    int n0, kb0, kb1;
};
__device__ __forceinline__ TileCoord tile_coord(const LinearParams& p, long long t) {
    TileCoord c;
    c.n0 = (int)(t % p.n_tiles) * p.block_n;
    t /= p.n_tiles;
    c.m0 = (t % p.m_tiles) * 128;
    c.kb0 = (int)(t / p.m_tiles) * p.kb_per_split;
    c.kb1 = c.kb0 + p.kb_per_split < p.k_blocks ? c.kb0 + p.kb_per_split : p.k_blocks;
    return c;
}

__device__ __forceinline__ float trunc_tf32(float v) { return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }
// low part of an fp32 operand whose high part the tensor core forms by truncation; rounded to nearest TF32 so that the
// hardware's own truncation of it is a no-op (a truncated low part would bias every product the same way)
__device__ __forceinline__ float lo_tf32(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v - trunc_tf32(v)));
    return __uint_as_float(r);
}
__device__ __forceinline__ float gelu_erf(float v) { return 0.5f * v * (1.0f + erff(v * 0.70710678118654752f)); }

template <bool STRICT>
__global__ void __launch_bounds__(STRICT ? 384 : 256, 1)
linear_tc_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                 const __grid_constant__ CUtensorMap tm_y, const __grid_constant__ CUtensorMap tm_blo, const LinearParams p) {
    extern __shared__ unsigned char smem_dyn[];
    // 1024-byte alignment: swizzle-128B tiles repeat every 8 rows x 128 bytes
    unsigned char* smem = smem_dyn + ((1024u - (smem_addr(smem_dyn) & 1023u)) & 1023u);
    const int b_bytes = p.block_n * 128;
    const int half_bytes = kABytes + b_bytes;                 // [x tile | W tile]
    const int stage_bytes = STRICT ? 2 * half_bytes : half_bytes;   // (+ [x lo | W lo])
    unsigned char* staging = smem + p.stages * stage_bytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(staging + kStagingBytes);
    uint64_t* empty = full + p.stages;
    uint64_t* ready = empty + p.stages;            // STRICT: split done
    uint64_t* tmem_full = ready + p.stages;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_x);
        tma_prefetch_desc(&tm_w);
        tma_prefetch_desc(&tm_y);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
            mbar_init(&ready[s], 128);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tmem_full[s], 1);
            mbar_init(&tmem_empty[s], kEpiWarps);
        }
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 512);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ---------------- TMA producer --------------------------------------------------------------
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (long long t = blockIdx.x; t < p.tiles; t += gridDim.x) {
                const TileCoord tc = tile_coord(p, t);
                const int m0 = (int)tc.m0, n0 = tc.n0;
                for (int kb = tc.kb0; kb < tc.kb1; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1u);
                    unsigned char* sa = smem + stage * stage_bytes;
                    mbar_arrive_expect_tx(&full[stage], (uint32_t)(half_bytes + (STRICT && p.w_lo ? b_bytes : 0)));
                    if (STRICT && p.w_lo) {      // pre-split B: its low part arrives like the high part, from its own tensor
                        if (p.b_mn) {
                            for (int j = 0; j < p.block_n / 32; ++j)
                                tma_load_2d(sa + half_bytes + kABytes + j * 4096, &tm_blo, &full[stage], n0 + 32 * j, kb * kBK);
                        } else {
                            tma_load_2d(sa + half_bytes + kABytes, &tm_blo, &full[stage], kb * kBK, n0);
                        }
                    }
                    if (p.a_mn) {   // [32 reduction rows x 32 MN] slabs, one per 128-byte swizzle atom along MN
                        for (int j = 0; j < kBM / 32; ++j) tma_load_2d(sa + j * 4096, &tm_x, &full[stage], m0 + 32 * j, kb * kBK);
                    } else {
                        tma_load_2d(sa, &tm_x, &full[stage], kb * kBK, m0);
                    }
                    if (p.b_mn) {
                        for (int j = 0; j < p.block_n / 32; ++j)
                            tma_load_2d(sa + kABytes + j * 4096, &tm_w, &full[stage], n0 + 32 * j, kb * kBK);
                    } else {
                        tma_load_2d(sa + kABytes, &tm_w, &full[stage], kb * kBK, n0);
                    }
                    if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer ----------------------------------------------------------------
        if (lane == 0) {
            const uint32_t idesc = idesc_tf32(kBM, (uint32_t)p.block_n, p.a_mn != 0, p.b_mn != 0);
            // one instruction covers 8 reduction elements: 32 bytes along a K-major row, or 8 rows (1024 bytes = two
            // 4-row swizzle atoms) of an MN-major slab
            const uint32_t step_a = p.a_mn ? 1024u : 32u, step_b = p.b_mn ? 1024u : 32u;
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (long long t = blockIdx.x; t < p.tiles; t += gridDim.x) {
                const TileCoord tc = tile_coord(p, t);
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);
                tcgen05_fence_after();
                const uint32_t d = tmem_base + (uint32_t)(acc * kStageCols);
                for (int kb = tc.kb0; kb < tc.kb1; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tcgen05_fence_after();
                    unsigned char* sa = smem + stage * stage_bytes;
                    const uint64_t da = p.a_mn ? smem_desc_mn128(sa, 4096u) : smem_desc_k128(sa);
                    const uint64_t db = p.b_mn ? smem_desc_mn128(sa + kABytes, 4096u) : smem_desc_k128(sa + kABytes);
#pragma unroll
                    for (int k = 0; k < kBK / 8; ++k)
                        mma_tf32(d, smem_desc_advance(da, (uint32_t)k * step_a), smem_desc_advance(db, (uint32_t)k * step_b), idesc,
                                 kb > tc.kb0 || k > 0);
                    if constexpr (STRICT) {
                        // hi*hi (above) and hi*lo_B with a pre-split B need only what TMA delivered: they are issued
                        // before waiting for the splitter warps, whose latency they hide
                        const uint64_t dal = p.a_mn ? smem_desc_mn128(sa + half_bytes, 4096u) : smem_desc_k128(sa + half_bytes);
                        const uint64_t dbl = p.b_mn ? smem_desc_mn128(sa + half_bytes + kABytes, 4096u)
                                                    : smem_desc_k128(sa + half_bytes + kABytes);
                        if (p.w_lo) {
#pragma unroll
                            for (int k = 0; k < kBK / 8; ++k)
                                mma_tf32(d, smem_desc_advance(da, (uint32_t)k * step_a), smem_desc_advance(dbl, (uint32_t)k * step_b), idesc, true);
                        }
                        mbar_wait(&ready[stage], phase);
                        tcgen05_fence_after();
                        if (!p.w_lo) {
#pragma unroll
                            for (int k = 0; k < kBK / 8; ++k)
                                mma_tf32(d, smem_desc_advance(da, (uint32_t)k * step_a), smem_desc_advance(dbl, (uint32_t)k * step_b), idesc, true);
                        }
#pragma unroll
                        for (int k = 0; k < kBK / 8; ++k)
                            mma_tf32(d, smem_desc_advance(dal, (uint32_t)k * step_a), smem_desc_advance(db, (uint32_t)k * step_b), idesc, true);
                    }
                    mma_commit(&empty[stage]);     // stage reusable once these MMAs have read it
                    if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                }
                mma_commit(&tmem_full[acc]);       // accumulator complete -> epilogue
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1u;
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // ---------------- epilogue: thread = output row ---------------------------------------------
        const int e = warp - 4;
        int acc = 0;
        uint32_t acc_phase = 0;
        unsigned chunk_no = 0;
        for (long long t = blockIdx.x; t < p.tiles; t += gridDim.x) {
            const TileCoord tc = tile_coord(p, t);
            const long long m0 = tc.m0;
            const int n0 = tc.n0;
            mbar_wait(&tmem_full[acc], acc_phase);
            tcgen05_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(e * 32) << 16) + (uint32_t)(acc * kStageCols);
            const long long row = m0 + e * 32 + lane;
            for (int c0 = 0; c0 < p.block_n; c0 += 32) {
                uint32_t r0[16], r1[16];
                tmem_ld16(taddr + (uint32_t)c0, r0);
                tmem_ld16(taddr + (uint32_t)c0 + 16u, r1);
                tmem_ld_wait();
                float v[32];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    v[i] = __uint_as_float(r0[i]);
                    v[16 + i] = __uint_as_float(r1[i]);
                }
                const int col0 = n0 + c0;
                if (p.bias != nullptr) {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (col0 + i < p.N) v[i] += __ldg(p.bias + col0 + i);
                }
                if (p.act == 1) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
                }
                if (p.store == 1) {
                    unsigned char* buf = staging + (e * 2 + (int)(chunk_no & 1u)) * 4096;
                    if (lane == 0) tma_store_wait_read<1>();   // the store issued two chunks ago has read this buffer
                    __syncwarp();
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        float4 q = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
                        *reinterpret_cast<float4*>(buf + lane * 128 + ((c ^ (lane & 7)) << 4)) = q;
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&tm_y, buf, col0, (int)(m0 + e * 32));
                        tma_store_commit();
                    }
                    ++chunk_no;
                } else if (p.store == 2) {
                    if (row < p.M) {
                        float* dst = p.y + row * (long long)p.N + col0;
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (col0 + i < p.N) atomicAdd(dst + i, v[i]);
                    }
                } else if (row < p.M) {
                    float* dst = p.y + row * (long long)p.N + col0;
                    if ((p.N & 3) == 0 && col0 + 32 <= p.N) {
#pragma unroll
                        for (int c = 0; c < 8; ++c)
                            *reinterpret_cast<float4*>(dst + 4 * c) = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (col0 + i < p.N) dst[i] = v[i];
                    }
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1u;
        }
        if (p.store == 1 && lane == 0) tma_store_wait<0>();
    } else if (STRICT && warp >= 8) {
        // ---------------- 3xTF32 split.  tcgen05 kind::tf32 TRUNCATES the 13 low mantissa bits of its fp32 operands
        // (measured: tools/tf32_probe.py), so the landed tile already is the high part as far as the tensor core is
        // concerned; only lo = rna_tf32(x - trunc(x)) has to be written.  The one dropped term, lo*lo, is <= 2^-20 of
        // each product (and of random sign once the B operand's split is rounding based, as the pre-split weights are)
        const int tt = threadIdx.x - 256;
        const int n4 = (p.w_lo ? kABytes : half_bytes) >> 4;     // pre-split B: only the A tile is split here
        int stage = 0;
        uint32_t phase = 0;
        for (long long t = blockIdx.x; t < p.tiles; t += gridDim.x) {
            const TileCoord tc = tile_coord(p, t);
            for (int kb = tc.kb0; kb < tc.kb1; ++kb) {
                mbar_wait(&full[stage], phase);
                const float4* hi = reinterpret_cast<const float4*>(smem + stage * stage_bytes);
                float4* lo = reinterpret_cast<float4*>(smem + stage * stage_bytes + half_bytes);
                for (int i = tt; i < n4; i += 128) {
                    const float4 x = hi[i];
                    float4 l;
                    l.x = lo_tf32(x.x); l.y = lo_tf32(x.y); l.z = lo_tf32(x.z); l.w = lo_tf32(x.w);
                    lo[i] = l;
                }
                fence_proxy_async_smem();
                mbar_arrive(&ready[stage]);
                if (++stage == p.stages) { stage = 0; phase ^= 1u; }
            }
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ---- host -----------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess) return nullptr;
        return q == cudaDriverEntryPointSuccess ? reinterpret_cast<EncodeTiledFn>(ptr) : nullptr;
    }();
    return fn;
}

}  // namespace

// fp32 row-major [outer, inner] matrix, box [box_outer, box_inner], 128-byte swizzle (16-byte chunks, or 32-byte
// chunks with atom32 - the layout of MN-major fp32 tensor-core operands), zero fill out of bounds
int tc_encode_2d(CUtensorMap* map, const void* base, long long inner, long long outer, int box_inner, int box_outer,
                 int atom32) {
    EncodeTiledFn fn = encode_fn();
    if (fn == nullptr) return fail(CNF_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
    const cuuint64_t strides[1] = {(cuuint64_t)inner * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(CNF_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return CNF_OK;
}

// fp32 row-major [outer, inner] matrix whose rows are exactly 64 bytes (inner = 16): box [box_outer, 16], 64-byte swizzle
// (16-byte chunk index ^= (row >> 1) & 3 - a thread per row then reads / writes its row without bank conflicts)
int tc_encode_2d_rows64(CUtensorMap* map, const void* base, long long outer, int box_outer) {
    EncodeTiledFn fn = encode_fn();
    if (fn == nullptr) return fail(CNF_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[2] = {16, (cuuint64_t)outer};
    const cuuint64_t strides[1] = {64};
    const cuuint32_t box[2] = {16, (cuuint32_t)box_outer};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(CNF_ERR_CUDA, "cuTensorMapEncodeTiled (64-byte rows) failed with CUresult %d", (int)r);
    return CNF_OK;
}

// N split into equal tiles of a multiple of 32 columns, at most 256 each
void tc_pick_block_n(int N, int* block_n, int* n_tiles) {
    const int nt = (N + 255) / 256;
    int bn = ((N + nt - 1) / nt + 31) / 32 * 32;
    if (bn > 256) bn = 256;
    *block_n = bn;
    *n_tiles = (N + bn - 1) / bn;
}

// ---- generic launcher: D[Mg,Ng] (+)= sum_r A(m,r) B(n,r) ------------------------------------------------------
//   operand in memory   K-major:  [rows, R] row-major (reduction contiguous)      -> nn.Linear forward operands
//                       MN-major: [R, rows] row-major (the M / N index contiguous) -> backward operands, no transposes
struct GemmDesc {
    const float* a; int a_mn;
    const float* b; int b_mn;
    long long Mg; int Ng; long long R;
    const float* bias; int act; int precision;
    float* y;
    const float* b_lo;   // 3xTF32, B K-major: b - trunc_tf32(b), same layout as b, or NULL (split in the kernel)
    int block_n;      // 0 = automatic; else the N tile (multiple of 32, <= 256) - tuning knob
    int split_k;      // 1: split the reduction over the grid and red.add the partial tiles into y (y pre-initialised)
};

static bool auto_block_n_off() {
    static const bool off = getenv("CNF_B200_LINEAR_WIDE_TILES") != nullptr;      // A/B switch: always the widest N tile
    return off;
}

int launch_gemm(const GemmDesc& g, const char* who, cudaStream_t stream) {
    CNF_SUPPORTED(g.Mg < (1ll << 31) - 256 && g.R < (1ll << 31) - 256, "%s: dimension too large for 32-bit TMA coordinates", who);
    CNF_SUPPORTED((g.a_mn ? g.Mg : g.R) % 4 == 0 && (g.b_mn ? (long long)g.Ng : g.R) % 4 == 0,
                  "%s: operand rows must be a multiple of 4 floats (16-byte pitch for TMA)", who);
    LinearParams p{};
    p.bias = g.bias; p.y = g.y; p.M = g.Mg; p.N = g.Ng; p.K = (int)g.R; p.act = g.act;
    p.a_mn = g.a_mn; p.b_mn = g.b_mn;
    tc_pick_block_n(g.Ng, &p.block_n, &p.n_tiles);
    p.k_blocks = (int)((g.R + kBK - 1) / kBK);
    p.m_tiles = (g.Mg + kBM - 1) / kBM;
    if (g.block_n >= 32 && g.block_n <= 256 && g.block_n % 32 == 0) {
        p.block_n = g.block_n;
        p.n_tiles = (g.Ng + g.block_n - 1) / g.block_n;
    } else if (!g.split_k && !auto_block_n_off() && p.m_tiles * p.n_tiles <= ((g.a_mn || g.b_mn) ? 1 : 2) * sm_count()) {
        // (The model below was fitted on forward products, both operands K-major; the MN-major backward products use it
        // only below ONE wave, where SMs would otherwise idle: between one and two waves it cost them 1 %, 331 vs 327 ms per
        // GraphCNF training step at 512 molecules.)
        // Few tiles - less than two waves of the widest N tile; larger problems keep it (measured 0.91 of the TF32 peak
        // there) - e.g. the projections of a 64-molecule shard, M = 2432 is 19 row tiles: the widest N tile leaves most SMs
        // idle and - at 96 KB per 3xTF32 stage - runs on a 2-deep ring.  Choose the N tile that minimises
        // waves x time per tile, time per tile ~ k_blocks x (a + b block_n) + c (a: the A tile and per-k-block latencies,
        // b: B tile traffic + MMA time per column, c: prologue + epilogue; measured on B200: one 256-wide 3xTF32 k-block
        // ~1.5 us).  Narrower tiles also buy ring depth: 64 KB per stage at 128 columns (3 stages), 48 KB at 64 (4).
        const long long sms_ = sm_count();
        double best = 1e30;
        int best_bn = p.block_n;
        for (int bn = 256; bn >= 32; bn -= 32) {
            const int nt = (g.Ng + bn - 1) / bn;
            if (bn != 256 && (nt - 1) * bn >= g.Ng) continue;
            if (bn > 32 && (g.Ng + nt - 1) / nt + 31 < bn) continue;      // same tile count fits a narrower tile: that one is listed too
            const long long tiles = p.m_tiles * nt;
            const double waves = (double)((tiles + sms_ - 1) / sms_);
            const double t = waves * (p.k_blocks * (0.25 + 1.25 * bn / 256.0) + 3.0);
            if (t < best - 1e-9) { best = t; best_bn = bn; }
        }
        p.block_n = best_bn;
        p.n_tiles = (g.Ng + best_bn - 1) / best_bn;
    }
    const long long base_tiles = p.m_tiles * p.n_tiles;
    const long long sms = sm_count();
    long long splits = 1;
    if (g.split_k) {
        splits = (sms + base_tiles - 1) / base_tiles;
        if (splits > p.k_blocks) splits = p.k_blocks;
        if (splits < 1) splits = 1;
    }
    p.kb_per_split = (int)((p.k_blocks + splits - 1) / splits);
    splits = (p.k_blocks + p.kb_per_split - 1) / p.kb_per_split;
    p.tiles = base_tiles * splits;
    p.store = g.split_k ? 2 : ((g.Ng % 4 == 0) ? 1 : 0);
    const bool strict = g.precision == 1;
    const int stage_bytes = (kABytes + p.block_n * 128) * (strict ? 2 : 1);
    const int fixed = 1024 + kStagingBytes + 512;
    int stages = (232448 - fixed) / stage_bytes;
    if (stages > 6) stages = 6;
    CNF_SUPPORTED(stages >= 2, "%s: tile does not fit shared memory", who);
    p.stages = stages;
    const size_t smem = (size_t)fixed + (size_t)stages * stage_bytes;

    CUtensorMap tm_a, tm_b, tm_y, tm_blo;
    p.w_lo = (strict && g.b_lo != nullptr) ? 1 : 0;
    int rc = g.a_mn ? tc_encode_2d(&tm_a, g.a, g.Mg, g.R, 32, kBK, 1) : tc_encode_2d(&tm_a, g.a, g.R, g.Mg, kBK, kBM, 0);
    if (rc != CNF_OK) return rc;
    rc = g.b_mn ? tc_encode_2d(&tm_b, g.b, g.Ng, g.R, 32, kBK, 1) : tc_encode_2d(&tm_b, g.b, g.R, g.Ng, kBK, p.block_n, 0);
    if (rc != CNF_OK) return rc;
    if (p.w_lo) {
        rc = g.b_mn ? tc_encode_2d(&tm_blo, g.b_lo, g.Ng, g.R, 32, kBK, 1) : tc_encode_2d(&tm_blo, g.b_lo, g.R, g.Ng, kBK, p.block_n, 0);
        if (rc != CNF_OK) return rc;
    } else {
        tm_blo = tm_b;
    }
    if (p.store == 1) {
        rc = tc_encode_2d(&tm_y, g.y, g.Ng, g.Mg, 32, 32, 0);
        if (rc != CNF_OK) return rc;
    } else {
        tm_y = tm_a;
    }
    long long grid = sms;
    if (grid > p.tiles) grid = p.tiles;
    if (strict) {
        CNF_CUDA(cudaFuncSetAttribute(linear_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        linear_tc_kernel<true><<<(unsigned)grid, 384, smem, stream>>>(tm_a, tm_b, tm_y, tm_blo, p);
    } else {
        CNF_CUDA(cudaFuncSetAttribute(linear_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        linear_tc_kernel<false><<<(unsigned)grid, 256, smem, stream>>>(tm_a, tm_b, tm_y, tm_blo, p);
    }
    return launch_status("linear_tc_kernel");
}

// grad_bias[n] += sum_m grad_y[m,n]: CTA = 32 columns x a slice of rows, 8 row lanes per column
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ gy, float* __restrict__ gb, long long M, int N,
                                                     long long rows_per_cta) {
    __shared__ float part[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int col = blockIdx.x * 32 + tx;
    const long long r0 = (long long)blockIdx.y * rows_per_cta;
    const long long r1 = r0 + rows_per_cta < M ? r0 + rows_per_cta : M;
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
    if (col < N) {
        long long r = r0 + ty;
        for (; r + 24 < r1; r += 32) {
            acc0 += __ldg(gy + r * N + col);
            acc1 += __ldg(gy + (r + 8) * N + col);
            acc2 += __ldg(gy + (r + 16) * N + col);
            acc3 += __ldg(gy + (r + 24) * N + col);
        }
        for (; r < r1; r += 8) acc0 += __ldg(gy + r * N + col);
    }
    part[ty][tx] = (acc0 + acc1) + (acc2 + acc3);
    __syncthreads();
    if (ty == 0 && col < N) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += part[i][tx];
        atomicAdd(gb + col, s);
    }
}

int tc_colsum(const float* gy, float* gb, long long M, int N, cudaStream_t stream) {
    const int col_blocks = (N + 31) / 32;
    long long row_blocks = (4ll * sm_count() + col_blocks - 1) / col_blocks;
    if (row_blocks > (M + 255) / 256) row_blocks = (M + 255) / 256;
    if (row_blocks < 1) row_blocks = 1;
    if (row_blocks > 65535) row_blocks = 65535;
    const long long rows_per_cta = (M + row_blocks - 1) / row_blocks;
    colsum_kernel<<<dim3((unsigned)col_blocks, (unsigned)row_blocks), 256, 0, stream>>>(gy, gb, M, N, rows_per_cta);
    return launch_status("colsum_kernel");
}

}  // namespace cnf

extern "C" int cnf_linear_fwd(const cnf_linear_args* a, cnf_stream_t stream_) {
    using namespace cnf;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    CNF_REQUIRE(a != nullptr, "cnf_linear_fwd: null args");
    CNF_REQUIRE(a->M >= 0 && a->N >= 1 && a->K >= 1, "cnf_linear_fwd: bad shape M=%lld N=%d K=%d", (long long)a->M, a->N, a->K);
    if (a->M == 0) return CNF_OK;
    CNF_REQUIRE(a->x && a->weight && a->y, "cnf_linear_fwd: null tensor");
    CNF_REQUIRE(a->precision == 0 || a->precision == 1, "cnf_linear_fwd: precision must be 0 (TF32) or 1 (3xTF32)");
    CNF_REQUIRE(a->activation == 0 || a->activation == 1, "cnf_linear_fwd: activation must be 0 (none) or 1 (GELU)");
    CNF_SUPPORTED(a->K % 4 == 0, "cnf_linear_fwd: in_features K=%d must be a multiple of 4 (16-byte rows for TMA)", a->K);
    CNF_REQUIRE(((reinterpret_cast<uintptr_t>(a->x) | reinterpret_cast<uintptr_t>(a->weight) | reinterpret_cast<uintptr_t>(a->y)) & 15) == 0,
                "cnf_linear_fwd: x, weight and y must be 16-byte aligned");
    GemmDesc g{};
    g.a = a->x; g.a_mn = 0; g.b = a->weight; g.b_mn = 0;
    g.Mg = a->M; g.Ng = a->N; g.R = a->K;
    g.bias = a->bias; g.act = a->activation; g.precision = a->precision; g.y = a->y; g.split_k = 0;
    g.b_lo = a->weight_lo;
    g.block_n = a->block_n;
    return launch_gemm(g, "cnf_linear_fwd", stream);
}

// Backward of y = x W^T + b.  All three products read the tensors where they lie (no transposed copies):
//   grad_x [M,K] = grad_y [M,N] . W [N,K]        A = grad_y K-major,  B = W MN-major
//   grad_W [N,K] += grad_y^T . x                   A = grad_y MN-major, B = x MN-major, reduction over M split across the
//                                                  grid, partial tiles added with red.global.add.f32
//   grad_b [N]   += column sums of grad_y
extern "C" int cnf_linear_bwd(const cnf_linear_bwd_args* a, cnf_stream_t stream_) {
    using namespace cnf;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    CNF_REQUIRE(a != nullptr, "cnf_linear_bwd: null args");
    CNF_REQUIRE(a->M >= 0 && a->N >= 1 && a->K >= 1, "cnf_linear_bwd: bad shape M=%lld N=%d K=%d", (long long)a->M, a->N, a->K);
    if (a->M == 0) return CNF_OK;
    CNF_REQUIRE(a->grad_y != nullptr, "cnf_linear_bwd: null grad_y");
    CNF_REQUIRE(a->precision == 0 || a->precision == 1, "cnf_linear_bwd: precision must be 0 (TF32) or 1 (3xTF32)");
    CNF_SUPPORTED(a->K % 4 == 0 && a->N % 4 == 0, "cnf_linear_bwd: N=%d and K=%d must be multiples of 4 (16-byte rows for TMA)",
                  a->N, a->K);
    uintptr_t bits = reinterpret_cast<uintptr_t>(a->grad_y);
    if (a->grad_x) bits |= reinterpret_cast<uintptr_t>(a->grad_x) | reinterpret_cast<uintptr_t>(a->weight);
    if (a->grad_weight) bits |= reinterpret_cast<uintptr_t>(a->grad_weight) | reinterpret_cast<uintptr_t>(a->x);
    CNF_REQUIRE((bits & 15) == 0, "cnf_linear_bwd: tensors must be 16-byte aligned");
    if (a->grad_x != nullptr) {
        CNF_REQUIRE(a->weight != nullptr, "cnf_linear_bwd: grad_x needs weight");
        GemmDesc g{};
        g.a = a->grad_y; g.a_mn = 0; g.b = a->weight; g.b_mn = 1;
        g.Mg = a->M; g.Ng = a->K; g.R = a->N;
        g.precision = a->precision; g.y = a->grad_x;
        if (a->weight_lo != nullptr) {
            CNF_REQUIRE((reinterpret_cast<uintptr_t>(a->weight_lo) & 15) == 0, "cnf_linear_bwd: weight_lo must be 16-byte aligned");
            g.b_lo = a->weight_lo;
        }
        const int rc = launch_gemm(g, "cnf_linear_bwd (grad_x)", stream);
        if (rc != CNF_OK) return rc;
    }
    if (a->grad_weight != nullptr) {
        CNF_REQUIRE(a->x != nullptr, "cnf_linear_bwd: grad_weight needs x");
        GemmDesc g{};
        g.a = a->grad_y; g.a_mn = 1; g.b = a->x; g.b_mn = 1;
        g.Mg = a->N; g.Ng = a->K; g.R = a->M;
        g.precision = a->precision; g.y = a->grad_weight; g.split_k = 1;
        const int rc = launch_gemm(g, "cnf_linear_bwd (grad_weight)", stream);
        if (rc != CNF_OK) return rc;
    }
    if (a->grad_bias != nullptr) return tc_colsum(a->grad_y, a->grad_bias, a->M, a->N, stream);
    return CNF_OK;
}
