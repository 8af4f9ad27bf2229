// K5 apply, C = 16 (the LM layout): invertible 1x1 convolution z @ W per position (reference
// layers/flows/permutation_layers.py:106-121) as a TMA-staged streaming kernel.
//
// The row kernel of invconv.cu (one thread per position, 16-byte global loads / stores at a 64-byte stride) is bound by
// L1 wavefronts, not by HBM: every strided LDG.128 / STG.128 costs ~16 wavefronts and W is re-read from shared memory 64
// times per position - l1tex 79 %, 3.1 TB/s (profiles/r02_invconv_rows_plain_ncu.txt).  Here the tiles move through the
// TMA engine instead of the LSU:
//
//   tensor maps   z, z_out (and the masked second output) as [P, 16] fp32 matrices, box = 256 rows x 64 bytes,
//                 CU_TENSOR_MAP_SWIZZLE_64B: the 16-byte chunk j of row r lands at chunk j ^ ((r >> 1) & 3), so a thread
//                 per row reads / writes its 4 chunks with conflict-free 16-byte shared-memory accesses
//   pipeline      persistent CTAs (2 per SM), 3 input stages (2 with the masked output) + 2 output tiles; one thread issues
//                 cp.async.bulk.tensor loads two tiles ahead (mbarrier complete_tx) and the tile stores (bulk groups)
//   compute       thread = position: row in registers, optional ActNorm prologue (activation_normalization.py:35-43),
//                 16 x 16 product with W broadcast from shared memory, pad mask, optional second output y * out_mask
//                 (the masked network input of the coupling layer that follows, coupling_layer.py:53)
// The per-sample ldj term (sum log|s| * length, :114-117) is added by the first threads of the grid.
#include <cuda.h>
#include <stdlib.h>

#include "cnf_common.cuh"
#include "tc_ptx.cuh"

namespace cnf {

int tc_encode_2d_rows64(CUtensorMap* map, const void* base, long long outer, int box_outer);

namespace {
using namespace tc;

constexpr int kTP = 256;                 // positions per tile = threads per CTA
constexpr int kC = 16;
constexpr int kTileBytes = kTP * kC * 4; // 16 KB
constexpr int kOut = 2;      // output tiles; input stages: 3, or 2 with the masked second output (two CTAs per SM either way)

struct TileParams {
    const float *w, *sldj, *pad, *length;
    float* ldj;
    uint32_t* status;
    long long P, B, ntiles;
    int S, reverse, masked, nin;
    const float *pre_b, *pre_s, *omask;
};

__global__ void __launch_bounds__(kTP, 2)
invconv_tile_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_out,
                    const __grid_constant__ CUtensorMap tm_msk, const TileParams p) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = smem_dyn + ((1024u - (smem_addr(smem_dyn) & 1023u)) & 1023u);
    unsigned char* s_in = smem;                                  // [nin][16 KB]
    unsigned char* s_out = s_in + p.nin * kTileBytes;            // [kOut][16 KB]
    unsigned char* s_msk = s_out + kOut * kTileBytes;            // [kOut][16 KB] (masked only)
    float* s_w = reinterpret_cast<float*>(s_msk + (p.masked ? kOut * kTileBytes : 0));   // [16 * 16]
    float* s_pb = s_w + kC * kC;
    float* s_pe = s_pb + kC;
    float* s_om = s_pe + kC;
    uint64_t* full = reinterpret_cast<uint64_t*>(s_om + kC);

    const int tid = threadIdx.x;
    for (int i = tid; i < kC * kC; i += kTP) s_w[i] = p.w[i];
    if (tid < kC) {
        s_pb[tid] = p.pre_b ? p.pre_b[tid] : 0.f;
        s_pe[tid] = p.pre_s ? expf(p.pre_s[tid]) : 1.0f;
        s_om[tid] = p.omask ? p.omask[tid] : 1.0f;
    }
    if (tid == 0) {
        tma_prefetch_desc(&tm_in);
        tma_prefetch_desc(&tm_out);
        if (p.masked) tma_prefetch_desc(&tm_msk);
        for (int s = 0; s < p.nin; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    // per-sample ldj term (:114-117): ldj[b] +/-= sum log|s| * length[b]
    if (p.ldj != nullptr) {
        for (long long b = (long long)blockIdx.x * kTP + tid; b < p.B; b += (long long)gridDim.x * kTP) {
            const float len = p.length ? p.length[b] : (float)p.S;
            const float t = p.sldj[0] * len;
            const float v = p.reverse ? p.ldj[b] - t : p.ldj[b] + t;
            p.ldj[b] = v;
            if (v != v) flag(p.status, CNF_FLAG_NAN_LDJ);
        }
    }

    const bool pre = p.pre_b != nullptr || p.pre_s != nullptr;
    const int tiles = (int)((p.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x);     // tiles blockIdx.x, + gridDim.x, ...
    auto tile_row0 = [&](int it) { return (long long)(blockIdx.x + (long long)it * gridDim.x) * kTP; };
    auto load = [&](int it) {      // thread 0 only
        const int st = it % p.nin;
        mbar_arrive_expect_tx(&full[st], (uint32_t)kTileBytes);
        tma_load_2d(s_in + st * kTileBytes, &tm_in, &full[st], 0, (int)tile_row0(it));
    };
    if (tid == 0) {
        for (int it = 0; it < p.nin - 1 && it < tiles; ++it) load(it);
    }
    // swizzled position of this thread's row: chunk j sits at j ^ ((row >> 1) & 3)
    const int sw = (tid >> 1) & 3;
    const int row_off = tid * (kC * 4);

    for (int it = 0; it < tiles; ++it) {
        const int st = it % p.nin, ob = it & 1;
        if (tid == 0) {
            if (it + p.nin - 1 < tiles) load(it + p.nin - 1);   // that stage was last read before the barriers of tile it - 1
            tma_store_wait_read<1>();                // the stores of tile it - 2 have drained output tile `ob`
        }
        mbar_wait(&full[st], (uint32_t)((it / p.nin) & 1));
        const long long pos = tile_row0(it) + tid;
        const bool valid = pos < p.P;
        const float pv = (p.pad != nullptr && valid) ? p.pad[pos] : 1.0f;
        float x[kC], y[kC];
        const unsigned char* src = s_in + st * kTileBytes + row_off;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float4 v = *reinterpret_cast<const float4*>(src + ((j ^ sw) << 4));
            x[4 * j] = v.x; x[4 * j + 1] = v.y; x[4 * j + 2] = v.z; x[4 * j + 3] = v.w;
        }
        if (pre) {
#pragma unroll
            for (int c = 0; c < kC; ++c) x[c] = (x[c] + s_pb[c]) * s_pe[c] * pv;
        }
#pragma unroll
        for (int co = 0; co < kC; ++co) y[co] = 0.f;
#pragma unroll
        for (int ci = 0; ci < kC; ++ci) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 wv = *reinterpret_cast<const float4*>(s_w + ci * kC + 4 * j);
                y[4 * j] = fmaf(x[ci], wv.x, y[4 * j]);
                y[4 * j + 1] = fmaf(x[ci], wv.y, y[4 * j + 1]);
                y[4 * j + 2] = fmaf(x[ci], wv.z, y[4 * j + 2]);
                y[4 * j + 3] = fmaf(x[ci], wv.w, y[4 * j + 3]);
            }
        }
        __syncthreads();     // every row of the input stage is in registers; thread 0's store-drain wait is done
        unsigned char* dst = s_out + ob * kTileBytes + row_off;
        unsigned char* dstm = s_msk + ob * kTileBytes + row_off;
        bool bad = false;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float4 v = make_float4(y[4 * j] * pv, y[4 * j + 1] * pv, y[4 * j + 2] * pv, y[4 * j + 3] * pv);
            bad = bad || v.x != v.x || v.y != v.y || v.z != v.z || v.w != v.w;
            *reinterpret_cast<float4*>(dst + ((j ^ sw) << 4)) = v;
            if (p.masked)
                *reinterpret_cast<float4*>(dstm + ((j ^ sw) << 4)) =
                    make_float4(v.x * s_om[4 * j], v.y * s_om[4 * j + 1], v.z * s_om[4 * j + 2], v.w * s_om[4 * j + 3]);
        }
        if (bad && valid) flag(p.status, CNF_FLAG_NAN_Z);
        fence_proxy_async_smem();
        __syncthreads();     // the output tile is complete and visible to the async proxy
        if (tid == 0) {
            tma_store_2d(&tm_out, s_out + ob * kTileBytes, 0, (int)tile_row0(it));
            if (p.masked) tma_store_2d(&tm_msk, s_msk + ob * kTileBytes, 0, (int)tile_row0(it));
            tma_store_commit();
        }
    }
    if (tid == 0) tma_store_wait<0>();
}

}  // namespace

// *handled = 1 when the tile kernel was launched (C = 16, 16-byte aligned rows, 32-bit row coordinates).
int invconv_tile_try(const cnf_invconv_args* a, cudaStream_t stream, int* handled) {
    *handled = 0;
    static const bool disabled = getenv("CNF_B200_INVCONV_ROWS") != nullptr;      // A/B switch: keep the row kernel
    if (disabled || a->C != kC) return CNF_OK;
    const long long P = a->B * a->S;
    if (P < 4 * kTP || P >= (1ll << 31) - kTP) return CNF_OK;
    if ((reinterpret_cast<uintptr_t>(a->z) | reinterpret_cast<uintptr_t>(a->z_out) | reinterpret_cast<uintptr_t>(a->z_masked_out)) & 15)
        return CNF_OK;
    TileParams p{};
    p.w = a->weight; p.sldj = a->sldj; p.pad = a->pad; p.length = a->length; p.ldj = a->ldj; p.status = a->status;
    p.P = P; p.B = a->B; p.ntiles = (P + kTP - 1) / kTP; p.S = (int)a->S; p.reverse = a->reverse;
    p.masked = a->z_masked_out != nullptr;
    p.pre_b = a->pre_actnorm_bias; p.pre_s = a->pre_actnorm_scales; p.omask = a->out_mask;
    CUtensorMap tm_in, tm_out, tm_msk;
    int rc = tc_encode_2d_rows64(&tm_in, a->z, P, kTP);
    if (rc != CNF_OK) return rc;
    rc = tc_encode_2d_rows64(&tm_out, a->z_out, P, kTP);
    if (rc != CNF_OK) return rc;
    rc = tc_encode_2d_rows64(&tm_msk, p.masked ? a->z_masked_out : a->z_out, P, kTP);
    if (rc != CNF_OK) return rc;
    p.nin = p.masked ? 2 : 3;
    const size_t smem = 1024 + (size_t)(p.nin + kOut + (p.masked ? kOut : 0)) * kTileBytes + (kC * kC + 3 * kC) * 4 + 3 * 8 + 64;
    static thread_local int configured_dev = -1;
    int dev = 0;
    CNF_CUDA(cudaGetDevice(&dev));
    if (configured_dev != dev) {
        CNF_CUDA(cudaFuncSetAttribute(invconv_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 104 * 1024));
        configured_dev = dev;
    }
    long long grid = 2ll * sm_count();
    if (grid > p.ntiles) grid = p.ntiles;
    *handled = 1;
    invconv_tile_kernel<<<(unsigned)grid, kTP, smem, stream>>>(tm_in, tm_out, tm_msk, p);
    return launch_status("invconv_tile_kernel");
}

}  // namespace cnf
