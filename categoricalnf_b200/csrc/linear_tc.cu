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
//     warps 8-11  (3xTF32 only) split every landed stage in place into hi = tf32(x) and lo = x - hi
//
//   precision 0  one TF32 pass (operands rounded to 10 mantissa bits: ~5e-4 relative per product)
//   precision 1  3xTF32: x W^T ~= hi hi + lo hi + hi lo, fp32 accumulation in tensor memory - error ~1e-6,
//                inside the 1e-4 parity budget of the flow transforms fed by this projection
#include <cuda.h>

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
    long long tiles;
    int act;        // 0 none, 1 GELU (erf)
    int store_tma;  // 1: epilogue through shared memory + TMA store; 0: guarded direct stores
};

__device__ __forceinline__ float rna_tf32(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}
__device__ __forceinline__ float gelu_erf(float v) { return 0.5f * v * (1.0f + erff(v * 0.70710678118654752f)); }

template <bool STRICT>
__global__ void __launch_bounds__(STRICT ? 384 : 256, 1)
linear_tc_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                 const __grid_constant__ CUtensorMap tm_y, const LinearParams p) {
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
                const int m0 = (int)(t / p.n_tiles) * kBM, n0 = (int)(t % p.n_tiles) * p.block_n;
                for (int kb = 0; kb < p.k_blocks; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1u);
                    unsigned char* sa = smem + stage * stage_bytes;
                    mbar_arrive_expect_tx(&full[stage], (uint32_t)half_bytes);
                    tma_load_2d(sa, &tm_x, &full[stage], kb * kBK, m0);
                    tma_load_2d(sa + kABytes, &tm_w, &full[stage], kb * kBK, n0);
                    if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer ----------------------------------------------------------------
        if (lane == 0) {
            const uint32_t idesc = idesc_tf32(kBM, (uint32_t)p.block_n);
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (long long t = blockIdx.x; t < p.tiles; t += gridDim.x) {
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);
                tcgen05_fence_after();
                const uint32_t d = tmem_base + (uint32_t)(acc * kStageCols);
                for (int kb = 0; kb < p.k_blocks; ++kb) {
                    mbar_wait(STRICT ? &ready[stage] : &full[stage], phase);
                    tcgen05_fence_after();
                    unsigned char* sa = smem + stage * stage_bytes;
                    const uint64_t da = smem_desc_k128(sa), db = smem_desc_k128(sa + kABytes);
#pragma unroll
                    for (int k = 0; k < kBK / 8; ++k) {
                        const uint32_t off = (uint32_t)k * 32u;   // 8 fp32 along K inside the swizzle atom
                        mma_tf32(d, smem_desc_advance(da, off), smem_desc_advance(db, off), idesc, kb > 0 || k > 0);
                        if constexpr (STRICT) {
                            const uint64_t dal = smem_desc_k128(sa + half_bytes), dbl = smem_desc_k128(sa + half_bytes + kABytes);
                            mma_tf32(d, smem_desc_advance(dal, off), smem_desc_advance(db, off), idesc, true);
                            mma_tf32(d, smem_desc_advance(da, off), smem_desc_advance(dbl, off), idesc, true);
                        }
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
            const long long m0 = (t / p.n_tiles) * kBM;
            const int n0 = (int)(t % p.n_tiles) * p.block_n;
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
                if (p.store_tma) {
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
        if (p.store_tma && lane == 0) tma_store_wait<0>();
    } else if (STRICT && warp >= 8) {
        // ---------------- 3xTF32 split: hi = tf32(x) (in place), lo = tf32(x - hi), both round-to-nearest
        // so that whatever the tensor core does with the low 13 mantissa bits, they are zero -------------
        const int tt = threadIdx.x - 256;
        const int n4 = half_bytes >> 4;
        int stage = 0;
        uint32_t phase = 0;
        for (long long t = blockIdx.x; t < p.tiles; t += gridDim.x) {
            for (int kb = 0; kb < p.k_blocks; ++kb) {
                mbar_wait(&full[stage], phase);
                float4* hi = reinterpret_cast<float4*>(smem + stage * stage_bytes);
                float4* lo = reinterpret_cast<float4*>(smem + stage * stage_bytes + half_bytes);
                for (int i = tt; i < n4; i += 128) {
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

// fp32 row-major [outer, inner] matrix, box [box_outer, box_inner], 128-byte swizzle, zero fill out of bounds
int tc_encode_2d(CUtensorMap* map, const void* base, long long inner, long long outer, int box_inner, int box_outer) {
    EncodeTiledFn fn = encode_fn();
    if (fn == nullptr) return fail(CNF_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
    const cuuint64_t strides[1] = {(cuuint64_t)inner * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(CNF_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
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
    CNF_SUPPORTED(a->M < (1ll << 31) - 256, "cnf_linear_fwd: M=%lld too large for 32-bit TMA coordinates", (long long)a->M);
    CNF_REQUIRE(((reinterpret_cast<uintptr_t>(a->x) | reinterpret_cast<uintptr_t>(a->weight) | reinterpret_cast<uintptr_t>(a->y)) & 15) == 0,
                "cnf_linear_fwd: x, weight and y must be 16-byte aligned");

    LinearParams p{};
    p.bias = a->bias; p.y = a->y; p.M = a->M; p.N = a->N; p.K = a->K; p.act = a->activation;
    tc_pick_block_n(a->N, &p.block_n, &p.n_tiles);
    p.k_blocks = (a->K + kBK - 1) / kBK;
    p.tiles = ((a->M + kBM - 1) / kBM) * p.n_tiles;
    p.store_tma = (a->N % 4 == 0) ? 1 : 0;
    const bool strict = a->precision == 1;
    const int stage_bytes = (kABytes + p.block_n * 128) * (strict ? 2 : 1);
    const int fixed = 1024 + kStagingBytes + 512;
    int stages = (232448 - fixed) / stage_bytes;
    if (stages > 6) stages = 6;
    CNF_SUPPORTED(stages >= 2, "cnf_linear_fwd: tile does not fit shared memory");
    p.stages = stages;
    const size_t smem = (size_t)fixed + (size_t)stages * stage_bytes;

    CUtensorMap tm_x, tm_w, tm_y;
    int rc = tc_encode_2d(&tm_x, a->x, a->K, a->M, kBK, kBM);
    if (rc != CNF_OK) return rc;
    rc = tc_encode_2d(&tm_w, a->weight, a->K, a->N, kBK, p.block_n);
    if (rc != CNF_OK) return rc;
    if (p.store_tma) {
        rc = tc_encode_2d(&tm_y, a->y, a->N, a->M, 32, 32);
        if (rc != CNF_OK) return rc;
    } else {
        tm_y = tm_x;
    }
    long long grid = sm_count();
    if (grid > p.tiles) grid = p.tiles;
    if (strict) {
        CNF_CUDA(cudaFuncSetAttribute(linear_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        linear_tc_kernel<true><<<(unsigned)grid, 384, smem, stream>>>(tm_x, tm_w, tm_y, p);
    } else {
        CNF_CUDA(cudaFuncSetAttribute(linear_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        linear_tc_kernel<false><<<(unsigned)grid, 256, smem, stream>>>(tm_x, tm_w, tm_y, p);
    }
    return launch_status("linear_tc_kernel");
}
