// Fused colouring apply for WCT — SURVEY.md §2b K7, §8 a6; reference: network/wct_rp.py:110,113
// (`T @ (cF - mean) + s_mean`).
//
//   out[o, n] = sum_c T[o, c] (x[c, n] - mu_c[c]) + mu_s[o]          x, out: [C, HW] fp32, T: [C, C]
//
// computed as D[n, o] (M = 128 positions per tile, N = C output channels, K = C input channels) so that the
// accumulator lanes are POSITIONS and the epilogue's stores are contiguous in `out`.  The A operand is x itself in
// its native orientation — positions contiguous = an MN-MAJOR tcgen05 operand (SWIZZLE_128B atoms of 64 positions x
// 8 channels): converter warps read fp32 x with 128-bit position-coalesced loads (one 512-byte channel row of the
// tile per warp instruction), centre, split into bf16 hi + lo and store 8 bytes — no transposition anywhere, and the
// pack pass that used to read 268 MB and write 268 MB before the GEMM read it again is gone.  T is tiny (<= 256 KiB
// packed) and streams from L2 by TMA, one k-block per stage.
//
//   warp 0        TMA producer of the T k-blocks            warps 2..5   epilogue (TMEM -> + mu_s -> out)
//   warp 1        tcgen05.mma issuer + TMEM owner           warps 6..13  converters (position group x channel half)
// Two 256-column TMEM accumulators: the epilogue of tile i overlaps the MMAs of tile i+1.
#include <cuda_bf16.h>

#include "common.cuh"
#include "umma.cuh"

namespace rpst {
namespace {

constexpr int kApConvWarps = 8;
constexpr int kApThreads = 32 * (6 + kApConvWarps);
constexpr int kApMaxC = 256;

struct ApplyParams {
    const float* x;          // [c, hw]
    const float* mu_c;       // [c]
    const float* mu_s;       // [c]
    const char* t_hi;        // packed T tiles: rows = output channels (cp/128 row blocks), K = input channels (kb tiles)
    const char* t_lo;
    float* out;              // [c, hw]
    int64_t hw;
    int c, cp, kb;           // channels, padded to 128/256, k-blocks of 64 input channels
    int tiles;               // ceil(hw / 128)
    int vec_ok;              // x 16-byte aligned and hw % 4 == 0: 128-bit loads
};

// A tile of one k-block: [8 K-atoms (8 channels each)][2 MN-atoms (64 positions each)][8 channel rows][128 B]
constexpr uint32_t kApLBO = 1024;    // between the two 64-position atoms
constexpr uint32_t kApSBO = 2048;    // between groups of 8 channels

// shared-memory descriptor: MN-major, SWIZZLE_128B (leading byte offset = MN-atom stride, stride byte offset = K-atom stride)
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fffu);
    d |= (uint64_t)(kApLBO >> 4) << 16;
    d |= (uint64_t)(kApSBO >> 4) << 32;
    d |= (uint64_t)1u << 46;
    d |= (uint64_t)2u << 61;
    return d;
}
constexpr uint32_t kIdescAMajorMN = 1u << 15;

__device__ __forceinline__ uint32_t ap_pack_bf16x2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

template <int PARTS>
__global__ void __launch_bounds__(kApThreads, 1) wct_apply_kernel(ApplyParams p) {
    // stage: A hi [128 pos x 64 ch] 16 KiB (+ lo 16 KiB), B hi [256 out x 64 ch] 32 KiB (+ lo 32 KiB)
    constexpr uint32_t kA = kTileBytes, kB = 2 * kTileBytes;
    constexpr uint32_t kStage = PARTS * (kA + kB);
    constexpr int NST = PARTS == 2 ? 2 : 4;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t full_a[NST], full_b[NST], empty[NST], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_slot;
    __shared__ float s_mu[kApMaxC];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_sub = p.cp / 128;                 // 16 KiB B tiles per part

    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; ++s) {
            mbar_init(&full_a[s], kApConvWarps);
            mbar_init(&full_b[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], 4);
        }
        mbar_fence_init();
    }
    for (int r = threadIdx.x; r < kApMaxC; r += blockDim.x) s_mu[r] = r < p.c ? __ldg(p.mu_c + r) : 0.f;
    if (warp == 1) tmem_alloc(&tmem_slot, 512);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            const uint64_t pol = policy_evict_last();
            uint32_t it = 0;
            for (int t = blockIdx.x; t < p.tiles; t += gridDim.x) {
                for (int kb = 0; kb < p.kb; ++kb, ++it) {
                    const int s = it % NST;
                    mbar_wait(&empty[s], ((it / NST) & 1u) ^ 1u);
                    unsigned char* st = smem + (size_t)s * kStage;
                    mbar_arrive_expect_tx(&full_b[s], (uint32_t)(PARTS * n_sub) * kTileBytes);
                    for (int part = 0; part < PARTS; ++part)
                        for (int sub = 0; sub < n_sub; ++sub)
                            tma_load_1d(st + PARTS * kA + part * kB + sub * kTileBytes,
                                        (part ? p.t_lo : p.t_hi) + ((int64_t)sub * p.kb + kb) * kTileBytes, kTileBytes, &full_b[s], pol);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16(128, p.cp) | kIdescAMajorMN;
            uint32_t it = 0;
            int ti = 0;
            for (int t = blockIdx.x; t < p.tiles; t += gridDim.x, ++ti) {
                const int buf = ti & 1;
                mbar_wait(&acc_empty[buf], ((ti >> 1) & 1) ^ 1);
                tcgen05_fence_after();
                const uint32_t d = tmem_base + (uint32_t)(buf * 256);
                for (int kb = 0; kb < p.kb; ++kb, ++it) {
                    const int s = it % NST;
                    mbar_wait(&full_a[s], (it / NST) & 1u);
                    mbar_wait(&full_b[s], (it / NST) & 1u);
                    tcgen05_fence_after();
                    const uint32_t a_hi = smem_u32(smem + (size_t)s * kStage), a_lo = a_hi + kA;
                    const uint32_t b_hi = a_hi + PARTS * kA, b_lo = b_hi + kB;
#pragma unroll
                    for (int k = 0; k < kTileK / kUmmaK; ++k) {
                        const uint32_t ko = k * kUmmaK * 2;            // B: 16 channels = 32 bytes along its K-major rows
                        const uint32_t ka = k * 2 * kApSBO;            // A: 16 channels = two 8-channel atoms
                        umma_bf16_ss(d, umma_desc_mn_sw128(a_hi + ka), umma_desc_k_sw128(b_hi + ko), idesc, kb > 0 || k > 0);
                        if (PARTS == 2) {
                            umma_bf16_ss(d, umma_desc_mn_sw128(a_hi + ka), umma_desc_k_sw128(b_lo + ko), idesc, true);
                            umma_bf16_ss(d, umma_desc_mn_sw128(a_lo + ka), umma_desc_k_sw128(b_hi + ko), idesc, true);
                        }
                    }
                    umma_commit(&empty[s]);
                    if (kb == p.kb - 1) umma_commit(&acc_full[buf]);
                }
            }
        }
    } else if (warp < 6) {
        // ------------------------------------------------------------------ epilogue
        const int q = warp & 3;
        int ti = 0;
        for (int t = blockIdx.x; t < p.tiles; t += gridDim.x, ++ti) {
            const int buf = ti & 1;
            mbar_wait(&acc_full[buf], (ti >> 1) & 1);
            tcgen05_fence_after();
            const int64_t n = (int64_t)t * 128 + q * 32 + lane;
            float v[32];
#pragma unroll 1
            for (int c0 = 0; c0 < p.cp; c0 += 32) {
                tmem_ld_32x32(tmem_base + (uint32_t)(buf * 256) + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
                if (n < p.hw) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (c0 + j < p.c) __stcs(p.out + (int64_t)(c0 + j) * p.hw + n, v[j] + __ldg(p.mu_s + c0 + j));
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);
        }
    } else {
        // ------------------------------------------------------------------ converters: centred bf16 hi / lo, MN-major tile
        const int cw = warp - 6;                           // warp cw converts channels [8 cw, 8 cw + 8) of every k-block
        // lane -> positions 4 lane .. 4 lane + 3 of the tile: MN-atom lane / 16, 16-byte chunk (lane % 16) / 2, half lane & 1
        const uint32_t lane_off = (uint32_t)(lane >> 4) * kApLBO + ((uint32_t)lane & 1u) * 8u;
        const uint32_t chunk = ((uint32_t)lane & 15u) >> 1;
        // Work units u = (tile, k-block) in issue order; three register buffers rotate so that the loads of units u+1 and
        // u+2 are in flight while unit u is converted (the kernel is load-latency bound: 8 loads per thread in flight
        // gave 1.7 TB/s).
        const int my_tiles = blockIdx.x < p.tiles ? (p.tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
        const uint32_t units = (uint32_t)my_tiles * (uint32_t)p.kb;
        auto load_unit = [&](float4 (&v)[8], uint32_t u) {
            if (u >= units) return;
            const int t = (int)blockIdx.x + (int)(u / p.kb) * (int)gridDim.x, kb = (int)(u % p.kb);
            const int64_t n = (int64_t)t * 128 + 4 * lane;
            const bool vec = p.vec_ok && n + 4 <= p.hw;               // aligned 128-bit loads; otherwise per-element
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = kb * 64 + cw * 8 + j;
                if (c < p.c && vec) {
                    v[j] = __ldcs(reinterpret_cast<const float4*>(p.x + (int64_t)c * p.hw + n));
                } else if (c < p.c) {
                    const float* src = p.x + (int64_t)c * p.hw + n;
                    const float mu = s_mu[c];
                    v[j].x = n + 0 < p.hw ? src[0] : mu;
                    v[j].y = n + 1 < p.hw ? src[1] : mu;
                    v[j].z = n + 2 < p.hw ? src[2] : mu;
                    v[j].w = n + 3 < p.hw ? src[3] : mu;
                } else {
                    v[j] = make_float4(0.f, 0.f, 0.f, 0.f);          // s_mu is 0 beyond the last channel
                }
            }
        };
        auto conv_unit = [&](float4 (&v)[8], uint32_t u) {
            if (u >= units) return;
            const int kb = (int)(u % p.kb);
            const int s = u % NST;
            mbar_wait(&empty[s], ((u / NST) & 1u) ^ 1u);
            unsigned char* a_hi = smem + (size_t)s * kStage + (size_t)cw * kApSBO + lane_off;
            unsigned char* a_lo = a_hi + kA;
#pragma unroll
            for (int j = 0; j < 8; ++j) {              // channel row j of this warp's 8-channel atom
                const float mu = s_mu[(kb * 64 + cw * 8 + j) & (kApMaxC - 1)];
                const float a = v[j].x - mu, b = v[j].y - mu, c2 = v[j].z - mu, d = v[j].w - mu;
                const uint32_t off = (uint32_t)j * 128u + ((chunk ^ (uint32_t)j) << 4);
                const uint32_t h01 = ap_pack_bf16x2(a, b), h23 = ap_pack_bf16x2(c2, d);
                *reinterpret_cast<uint2*>(a_hi + off) = make_uint2(h01, h23);
                if (PARTS == 2) {
                    const uint32_t l01 = ap_pack_bf16x2(a - __uint_as_float(h01 << 16), b - __uint_as_float(h01 & 0xffff0000u));
                    const uint32_t l23 = ap_pack_bf16x2(c2 - __uint_as_float(h23 << 16), d - __uint_as_float(h23 & 0xffff0000u));
                    *reinterpret_cast<uint2*>(a_lo + off) = make_uint2(l01, l23);
                }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full_a[s]);
        };
        float4 v0[8], v1[8], v2[8];
        load_unit(v0, 0);
        load_unit(v1, 1);
        for (uint32_t u = 0; u < units; u += 3) {
            load_unit(v2, u + 2);
            conv_unit(v0, u);
            load_unit(v0, u + 3);
            conv_unit(v1, u + 1);
            load_unit(v1, u + 4);
            conv_unit(v2, u + 2);
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace

bool wct_apply_fused_supported(int64_t c, int64_t hw) { return c >= 1 && c <= kApMaxC && hw >= 1; }

// out = T (x - mu_c) + mu_s with T given as packed tiles (pack_operand of the [c,c] fp32 matrix: rows = output channels)
int wct_apply_fused(const float* x, const float* mu_c, const float* mu_s, const void* t_hi, const void* t_lo, float* out,
                    int64_t c, int64_t hw, int passes, cudaStream_t st) {
    ApplyParams p{};
    p.x = x; p.mu_c = mu_c; p.mu_s = mu_s;
    p.t_hi = static_cast<const char*>(t_hi); p.t_lo = static_cast<const char*>(t_lo);
    p.out = out; p.hw = hw; p.c = (int)c;
    p.cp = c <= 128 ? 128 : 256;
    p.kb = (int)((c + kTileK - 1) / kTileK);
    p.tiles = (int)((hw + 127) / 128);
    p.vec_ok = (hw % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15u) == 0) ? 1 : 0;
    constexpr size_t smem = 1024 + 2 * 2 * (kTileBytes + 2 * kTileBytes);   // 193 KiB for both instantiations
    static PerDeviceFlag configured_on;
    bool& configured = configured_on.get();
    if (!configured) {
        RPST_CUDA(cudaFuncSetAttribute(wct_apply_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        RPST_CUDA(cudaFuncSetAttribute(wct_apply_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    int grid = sm_count();
    if (grid > p.tiles) grid = p.tiles;
    if (passes == 3) wct_apply_kernel<2><<<grid, kApThreads, smem, st>>>(p);
    else wct_apply_kernel<1><<<grid, kApThreads, smem, st>>>(p);
    RPST_CUDA(cudaGetLastError());
    return RPST_OK;
}

}  // namespace rpst

RPST_WATCHDOG_SETTER(wct_apply)
