// Fused centred covariance for WCT — SURVEY.md §2b K5, §8 a6; reference: network/wct_rp.py:85-94
// (`cF - mean`, `cF @ cF.t() / (HW - 1)` [+ I on the content side]).
//
// One kernel reads the fp32 features ONCE and feeds the tensor cores directly: no packed-operand round trip
// (the separate pack pass read 268 MB and wrote 222 MB per operand before the GEMM read them again).
//   warps 1..8  converters: 128-bit global loads of a [C x 64 positions] k-tile, subtract a per-channel SHIFT,
//               split into bf16 hi + lo, store into the K-major SWIZZLE_128B operand tile in shared memory
//               (A and B of the SYRK are the same tile), and keep fp32 row sums of the shifted values
//   warp 0      one thread issues tcgen05.mma: X X^T over this CTA's slice of H*W (split-K over all SMs) into TMEM;
//               only the blocks on and above the diagonal are computed (SYRK symmetry: 3 of 4 128x128 blocks)
//   epilogue    TMEM -> per-CTA fp32 partial Gram; a fixed-order fp64 reduction kernel forms
//               cov = (G - S S^T / HW) / (HW - 1) (+ diag) and the exact means mu = shift + S / HW
// The shift is a cheap sub-sampled mean: centring with ANY shift is exact after the rank-1 correction, and a shift
// close to the mean keeps the correction small (no cancellation), so no full statistics pass is needed.
#include <cuda_bf16.h>

#include "common.cuh"
#include "umma.cuh"

namespace rpst {
namespace {

constexpr int kCovConvWarps = 8;
constexpr int kCovThreads = 32 * (1 + kCovConvWarps);
constexpr int kCovMaxC = 256;
constexpr uint32_t kCovPartBytes = 2 * kTileBytes;          // [256 rows x 64 positions] bf16 = 32 KiB

struct CovParams {
    const float* x;          // [c, hw] one sample, row stride hw
    const float* shift;      // [cp] per-channel shift (rows >= c: 0)
    float* partial;          // [grid, cp, cp] per-CTA partial Gram (blocks on/above the diagonal only)
    float* rowsum;           // [grid, cp] per-CTA sums of (x - shift)
    int64_t hw;
    int c, cp;               // channels, padded to 128 / 256
    int k_tiles;             // ceil(hw / 64)
    int passes;              // 1: bf16, 3: bf16x3
};

template <int PARTS>   // 1: hi only, 2: hi + lo
__global__ void __launch_bounds__(kCovThreads, 1) cov_fused_kernel(CovParams p) {
    constexpr int NST = PARTS == 2 ? 3 : 6;
    constexpr uint32_t kStage = kCovPartBytes * PARTS;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t full[NST], empty[NST], acc_full;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tiles = p.cp / 128;
    // this CTA's slice of the k-tiles
    const int per = (p.k_tiles + gridDim.x - 1) / gridDim.x;
    const int kt0 = blockIdx.x * per;
    const int kt1 = min(p.k_tiles, kt0 + per);
    const int nkt = max(0, kt1 - kt0);

    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; ++s) {
            mbar_init(&full[s], kCovConvWarps);
            mbar_init(&empty[s], 1);
        }
        mbar_init(&acc_full, 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = tmem_slot;

    if (warp == 0) {
        if (lane == 0 && nkt > 0) {
            // SYRK: M-tile 0 against all columns, M-tile 1 against columns 128..255 only
            const uint32_t idesc0 = umma_idesc_bf16(128, p.cp);
            const uint32_t idesc1 = umma_idesc_bf16(128, 128);
            for (int it = 0; it < nkt; ++it) {
                const int s = it % NST;
                mbar_wait(&full[s], (uint32_t)(it / NST) & 1u);
                tcgen05_fence_after();
                const uint32_t hi = smem_u32(smem + (size_t)s * kStage), lo = hi + kCovPartBytes;
#pragma unroll
                for (int k = 0; k < kTileK / kUmmaK; ++k) {
                    const uint32_t ko = k * kUmmaK * 2;
                    const bool acc = it > 0 || k > 0;
                    umma_bf16_ss(tmem_base, umma_desc_k_sw128(hi + ko), umma_desc_k_sw128(hi + ko), idesc0, acc);
                    if (PARTS == 2) {
                        umma_bf16_ss(tmem_base, umma_desc_k_sw128(hi + ko), umma_desc_k_sw128(lo + ko), idesc0, true);
                        umma_bf16_ss(tmem_base, umma_desc_k_sw128(lo + ko), umma_desc_k_sw128(hi + ko), idesc0, true);
                    }
                    if (m_tiles == 2) {
                        const uint32_t h1 = hi + kTileBytes + ko, l1 = lo + kTileBytes + ko;
                        umma_bf16_ss(tmem_base + 256, umma_desc_k_sw128(h1), umma_desc_k_sw128(h1), idesc1, acc);
                        if (PARTS == 2) {
                            umma_bf16_ss(tmem_base + 256, umma_desc_k_sw128(h1), umma_desc_k_sw128(l1), idesc1, true);
                            umma_bf16_ss(tmem_base + 256, umma_desc_k_sw128(l1), umma_desc_k_sw128(h1), idesc1, true);
                        }
                    }
                }
                umma_commit(&empty[s]);
            }
            umma_commit(&acc_full);
        }
    } else {
        // ------------------------------------------------------------------ converters: warp cw owns rows [32 cw, 32 cw + 32)
        const int cw = warp - 1;
        const int sub = lane >> 4;                 // which of the two rows of a load
        const int p4 = (lane & 15) * 4;            // first of this lane's 4 positions inside the k-tile
        float sh[16], acc[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int row = cw * 32 + 2 * i + sub;
            sh[i] = row < p.cp ? __ldg(p.shift + row) : 0.f;
            acc[i] = 0.f;
        }
        const bool rows_used = cw * 32 < p.cp;     // C <= 128: the upper converter warps only keep the handshake going
        for (int it = 0; it < nkt; ++it) {
            const int kt = kt0 + it;
            const int64_t pos = (int64_t)kt * kTileK + p4;
            float4 v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int row = cw * 32 + 2 * i + sub;
                if (rows_used && row < p.c && pos + 4 <= p.hw) {
                    v[i] = __ldcs(reinterpret_cast<const float4*>(p.x + (int64_t)row * p.hw + pos));
                } else if (rows_used && row < p.c && pos < p.hw) {       // ragged last k-tile (hw % 64 != 0, hw % 4 == 0 never lands here)
                    const float* src = p.x + (int64_t)row * p.hw + pos;
                    v[i].x = src[0];
                    v[i].y = pos + 1 < p.hw ? src[1] : sh[i];
                    v[i].z = pos + 2 < p.hw ? src[2] : sh[i];
                    v[i].w = sh[i];
                } else {
                    v[i] = make_float4(sh[i], sh[i], sh[i], sh[i]);      // padding: (x - shift) = 0
                }
            }
            const int s = it % NST;
            mbar_wait(&empty[s], ((uint32_t)(it / NST) & 1u) ^ 1u);
            unsigned char* hi = smem + (size_t)s * kStage;
            unsigned char* lo = hi + kCovPartBytes;
            if (rows_used) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int row = cw * 32 + 2 * i + sub;
                    const float a = v[i].x - sh[i], b = v[i].y - sh[i], c2 = v[i].z - sh[i], d = v[i].w - sh[i];
                    acc[i] += (a + b) + (c2 + d);
                    __nv_bfloat16 h0, h1, h2, h3, l0, l1, l2, l3;
                    split_bf16(a, h0, l0); split_bf16(b, h1, l1); split_bf16(c2, h2, l2); split_bf16(d, h3, l3);
                    const uint32_t off = (uint32_t)row * 128u + (uint32_t)((((p4 >> 3) ^ (row & 7)) << 4) + ((lane & 1) << 3));
                    __nv_bfloat162 a01 = __halves2bfloat162(h0, h1), a23 = __halves2bfloat162(h2, h3);
                    uint2 w;
                    w.x = *reinterpret_cast<uint32_t*>(&a01); w.y = *reinterpret_cast<uint32_t*>(&a23);
                    *reinterpret_cast<uint2*>(hi + off) = w;
                    if (PARTS == 2) {
                        __nv_bfloat162 b01 = __halves2bfloat162(l0, l1), b23 = __halves2bfloat162(l2, l3);
                        w.x = *reinterpret_cast<uint32_t*>(&b01); w.y = *reinterpret_cast<uint32_t*>(&b23);
                        *reinterpret_cast<uint2*>(lo + off) = w;
                    }
                }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[s]);
        }
        // row sums of the shifted values: the 16 lanes sharing a row add up, fixed order
        if (rows_used) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                float a = acc[i];
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
                const int row = cw * 32 + 2 * i + sub;
                if ((lane & 15) == 0 && row < p.cp) p.rowsum[(size_t)blockIdx.x * p.cp + row] = a;
            }
        }
        // ------------------------------------------------------------------ epilogue (warps 1..4 = TMEM quarters 1,2,3,0)
        if (warp <= 4) {
            const int q = warp & 3;
            const int r = q * 32 + lane;
            float* out = p.partial + (size_t)blockIdx.x * p.cp * p.cp;
            if (nkt > 0) {
                mbar_wait(&acc_full, 0);
                tcgen05_fence_after();
            }
            float vals[32];
            for (int c0 = 0; c0 < p.cp; c0 += 32) {              // M-tile 0: rows 0..127, all columns
                if (nkt > 0) tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, vals);
                else
#pragma unroll
                    for (int j = 0; j < 32; ++j) vals[j] = 0.f;
                float4* dst = reinterpret_cast<float4*>(out + (size_t)r * p.cp + c0);
#pragma unroll
                for (int j = 0; j < 8; ++j) dst[j] = make_float4(vals[4 * j], vals[4 * j + 1], vals[4 * j + 2], vals[4 * j + 3]);
            }
            if (m_tiles == 2) {
                for (int c0 = 0; c0 < 128; c0 += 32) {           // M-tile 1: rows 128..255, columns 128..255
                    if (nkt > 0) tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + 256u + (uint32_t)c0, vals);
                    else
#pragma unroll
                        for (int j = 0; j < 32; ++j) vals[j] = 0.f;
                    float4* dst = reinterpret_cast<float4*>(out + (size_t)(128 + r) * p.cp + 128 + c0);
#pragma unroll
                    for (int j = 0; j < 8; ++j) dst[j] = make_float4(vals[4 * j], vals[4 * j + 1], vals[4 * j + 2], vals[4 * j + 3]);
                }
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// sub-sampled per-channel mean (up to 32 segments of 64 positions spread over the plane): the centring shift
__global__ void __launch_bounds__(128) cov_shift_kernel(const float* __restrict__ x, int c, int cp, int64_t hw,
                                                        float* __restrict__ shift) {
    const int row = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= cp) return;
    if (row >= c) {
        if (lane == 0) shift[row] = 0.f;
        return;
    }
    const int64_t segs = hw / 64 > 32 ? 32 : (hw / 64 > 0 ? hw / 64 : 1);
    const int64_t seg_len = hw / 64 > 0 ? 64 : hw;
    const int64_t stride = hw / segs;
    float a = 0.f;
    for (int64_t i = lane; i < segs * seg_len; i += 32) a += x[(int64_t)row * hw + (i / seg_len) * stride + (i % seg_len)];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) shift[row] = a / (float)(segs * seg_len);
}

// fixed-order fp64 reduction of the per-CTA partials + rank-1 centring correction; also the exact means
__global__ void __launch_bounds__(256) cov_finalize_kernel(const float* __restrict__ partial, const float* __restrict__ rowsum,
                                                           const float* __restrict__ shift, int parts, int c, int cp, double hw,
                                                           double diag_add, double* __restrict__ cov, float* __restrict__ mean) {
    __shared__ double s_sum[kCovMaxC];
    for (int r = threadIdx.x; r < cp; r += blockDim.x) {
        double a = 0.0;
        for (int z = 0; z < parts; ++z) a += (double)rowsum[(size_t)z * cp + r];
        s_sum[r] = a;
    }
    __syncthreads();
    const int i = blockIdx.x;                       // one row of the covariance per block
    if (mean && threadIdx.x == 0) mean[i] = (float)((double)shift[i] + s_sum[i] / hw);
    for (int j = threadIdx.x; j < c; j += blockDim.x) {
        // block (1,0) was not computed: take the mirrored entry
        const bool lower = i >= 128 && j < 128;
        const size_t idx = lower ? (size_t)j * cp + i : (size_t)i * cp + j;
        double g = 0.0;
        for (int z = 0; z < parts; ++z) g += (double)partial[(size_t)z * cp * cp + idx];
        cov[(size_t)i * c + j] = (g - s_sum[i] * s_sum[j] / hw) / (hw - 1.0) + (i == j ? diag_add : 0.0);
    }
}

}  // namespace

bool cov_fused_supported(const float* x, int64_t c, int64_t hw) {
    return c >= 1 && c <= kCovMaxC && hw >= 64 && hw % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15u) == 0;
}

size_t cov_fused_workspace_bytes(int64_t c) {
    const size_t cp = c <= 128 ? 128 : 256;
    const size_t g = (size_t)sm_count();
    return align_up(g * cp * cp * sizeof(float), 256) + align_up(g * cp * sizeof(float), 256) + align_up(cp * sizeof(float), 256);
}

// cov [c,c] fp64 = centred covariance of x [c,hw] (+ diag_add on the diagonal); mean [c] fp32 exact channel means
int cov_fused(const float* x, int64_t c, int64_t hw, int passes, double diag_add, double* cov, float* mean, void* workspace,
              cudaStream_t st) {
    const int cp = c <= 128 ? 128 : 256;
    const int g = sm_count();
    char* w = static_cast<char*>(workspace);
    float* partial = reinterpret_cast<float*>(w);
    float* rowsum = reinterpret_cast<float*>(w + align_up((size_t)g * cp * cp * sizeof(float), 256));
    float* shift = reinterpret_cast<float*>(reinterpret_cast<char*>(rowsum) + align_up((size_t)g * cp * sizeof(float), 256));
    cov_shift_kernel<<<(cp + 3) / 4, 128, 0, st>>>(x, (int)c, cp, hw, shift);
    RPST_CUDA(cudaGetLastError());
    CovParams p{};
    p.x = x; p.shift = shift; p.partial = partial; p.rowsum = rowsum; p.hw = hw; p.c = (int)c; p.cp = cp;
    p.k_tiles = (int)((hw + kTileK - 1) / kTileK);
    p.passes = passes;
    int grid = g < p.k_tiles ? g : p.k_tiles;
    const int per = (p.k_tiles + grid - 1) / grid;
    grid = (p.k_tiles + per - 1) / per;             // no CTA without work
    static PerDeviceFlag configured_on;
    bool& configured = configured_on.get();
    constexpr size_t smem = 1024 + 3 * 2 * kCovPartBytes;   // 193 KiB for both instantiations
    if (!configured) {
        RPST_CUDA(cudaFuncSetAttribute(cov_fused_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        RPST_CUDA(cudaFuncSetAttribute(cov_fused_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    if (passes == 3) cov_fused_kernel<2><<<grid, kCovThreads, smem, st>>>(p);
    else cov_fused_kernel<1><<<grid, kCovThreads, smem, st>>>(p);
    RPST_CUDA(cudaGetLastError());
    cov_finalize_kernel<<<(unsigned)c, 256, 0, st>>>(partial, rowsum, shift, grid, (int)c, cp, (double)hw, diag_add, cov, mean);
    RPST_CUDA(cudaGetLastError());
    return RPST_OK;
}

}  // namespace rpst
