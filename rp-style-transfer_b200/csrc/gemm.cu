// Dense-contraction building blocks on the 5th-gen tensor cores (tcgen05 + TMEM), used by the MRF
// normalised-cross-correlation, the WCT colouring apply and (as a reference point) SANet:
//
//   pack_operand_kernel   fp32 matrix (any row/column stride, optional per-row scale) -> bf16 hi (+lo)
//                         operand tiles in the K-major SW128 tile format of umma.cuh
//   gemm_packed_kernel    D[M,N] (fp32) = alpha * A.B^T (+ row_add[i] + col_add[j]) with A [M,K], B [N,K]
//                         given as packed tiles.  precision passes: 1 = bf16 x bf16; 3 = "bf16x3"
//                         (hi.hi + hi.lo + lo.hi with fp32 accumulation in TMEM, ~2^-16 relative error,
//                         i.e. fp32-grade results at one third of the bf16 tensor rate).
//
// Kernel anatomy (persistent, one CTA per SM walking 128 x 256 output tiles, 6 warps; 18 with the top-k epilogue):
//   warp 0 / 1 thread : TMA producer — one 1-D bulk copy per 16 KiB operand tile into a 4-stage ring
//   warp 1            : allocates TMEM (2 x 256 columns); 1 thread issues tcgen05.mma (M=128, N=256, K=16,
//                       4 per stage), commits the stage back to the producer and the finished
//                       accumulator to the epilogue
//   warps 2..5 [..17] : epilogue — tcgen05.ld (warp w: TMEM lane quarter w % 4, column slice (w - 2) / 4), scale/add,
//                       store (256-bit), or top-k selection, or packed bf16 operand tiles for a following product;
//                       runs concurrently with the MMAs of the next tile (double-buffered accumulator)
#include <cuda_fp16.h>

#include "common.cuh"
#include "umma.cuh"

namespace rpst {
namespace {

// ------------------------------------------------------------------------------------------ pack
struct PackParams {
    const float* x;
    int64_t rows, k;            // logical extent
    int64_t stride_r, stride_k; // element strides of x
    const float* row_scale;     // may be null
    const float* row_shift;     // may be null: value subtracted before scaling (centring)
    __nv_bfloat16* hi;
    __nv_bfloat16* lo;          // may be null
    int64_t row_tiles, k_tiles;
    // batch (blockIdx.y): element stride of x, byte stride of the tile buffers, element stride of scale/shift
    int64_t x_batch, tile_batch_bytes, vec_batch;
    int vec_k;                  // stride_k == 1, x and every row start 16-byte aligned: 128-bit source loads
    int fp16;                   // 1: emit IEEE half (11-bit significand) into `hi` instead of bf16 (flash-attention V operand)
};

// One thread produces one 16-byte chunk (8 consecutive k of one row).  Threads of a warp walk the
// unit-stride direction of the source so the fp32 reads coalesce.
__global__ void __launch_bounds__(256) pack_operand_kernel(PackParams p) {
    const int64_t tile = blockIdx.x;               // (row tile, k tile)
    const int64_t rb = tile / p.k_tiles, kb = tile % p.k_tiles;
    const int64_t bi = blockIdx.y;
    p.x += bi * p.x_batch;
    if (p.row_scale) p.row_scale += bi * p.vec_batch;
    if (p.row_shift) p.row_shift += bi * p.vec_batch;
    char* hi_tile = reinterpret_cast<char*>(p.hi) + bi * p.tile_batch_bytes + tile * kTileBytes;
    char* lo_tile = p.lo ? reinterpret_cast<char*>(p.lo) + bi * p.tile_batch_bytes + tile * kTileBytes : nullptr;
    for (int item = threadIdx.x; item < kTileRows * 8; item += blockDim.x) {
        int r, c;
        if (p.stride_r == 1) { r = item % kTileRows; c = item / kTileRows; }   // MN-major source
        else { c = item % 8; r = item / 8; }                                   // K-major source
        const int64_t row = rb * kTileRows + r;
        const int64_t k0 = kb * kTileK + c * 8;
        const float sc = (p.row_scale && row < p.rows) ? __ldg(p.row_scale + row) : 1.f;
        const float sh = (p.row_shift && row < p.rows) ? __ldg(p.row_shift + row) : 0.f;
        __align__(16) __nv_bfloat16 h[8];
        __align__(16) __nv_bfloat16 l[8];
        float v[8];
        if (p.vec_k && row < p.rows && k0 + 8 <= p.k) {   // K-major source, 16-byte aligned rows: two 128-bit loads
            const float4* src = reinterpret_cast<const float4*>(p.x + row * p.stride_r + k0);
            const float4 a = __ldg(src), b = __ldg(src + 1);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e)
                v[e] = (row < p.rows && k0 + e < p.k) ? __ldg(p.x + row * p.stride_r + (k0 + e) * p.stride_k) : sh;
        }
        const uint32_t off = tile_chunk_offset(r, c);
        if (p.fp16) {
            __align__(16) __half hh[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) hh[e] = __float2half_rn((row < p.rows && k0 + e < p.k) ? (v[e] - sh) * sc : 0.f);
            *reinterpret_cast<uint4*>(hi_tile + off) = *reinterpret_cast<const uint4*>(hh);
            continue;
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) split_bf16((row < p.rows && k0 + e < p.k) ? (v[e] - sh) * sc : 0.f, h[e], l[e]);
        *reinterpret_cast<uint4*>(hi_tile + off) = *reinterpret_cast<const uint4*>(h);
        if (lo_tile) *reinterpret_cast<uint4*>(lo_tile + off) = *reinterpret_cast<const uint4*>(l);
    }
}

// ------------------------------------------------------------------------------------------ gemm
constexpr int kTopkSubs = 4;   // TOPK epilogue: 4 warps per TMEM lane quarter, 64 columns of the tile each (one warp per quarter
                               // took 2.5x the tile's MMA time for its two passes).  The storing epilogues use them only for short
                               // K: with 16 warps the 16384^2 K = 512 fp32 product went 755 -> 715 us, but the K = 16384
                               // products (P.V, f_psi) lost 2-8 %.
constexpr int kGemmThreads = 192;
constexpr int kGemmTopkThreads = 64 + 128 * kTopkSubs;
constexpr int kGemmBN = 256;                         // output tile: 128 x 256
constexpr int kGemmBTiles = kGemmBN / kTileRows;     // 16 KiB B tiles per stage and precision part
constexpr uint32_t kGemmPartBytes = kTileBytes * (1 + kGemmBTiles);   // A tile + B tiles of one part (48 KiB)
// A stage holds one k-tile of both operands.  bf16: hi parts only, 4 stages.  bf16x3: hi AND lo parts
// (96 KiB), 2 stages — every operand tile is fetched once per output tile and the three MMA groups
// (hi.hi, hi.lo, lo.hi) are issued from the same stage instead of re-streaming tiles once per pass.
constexpr int kGemmMaxStages = 4;
constexpr size_t kGemmSmemBytes = (size_t)kGemmMaxStages * kGemmPartBytes + 1024;

struct GemmParams {
    const __nv_bfloat16* a_hi;
    const __nv_bfloat16* a_lo;
    const __nv_bfloat16* b_hi;
    const __nv_bfloat16* b_lo;
    float* out;
    int64_t m, n;        // valid output extent
    int64_t ldo;         // output row stride (elements)
    int k_tiles;         // K / 64 (padded): tile pitch of the packed operands
    int k_per_split;     // k tiles handled by one split-K slice (== k_tiles without split-K)
    int splits;
    int64_t split_stride;  // elements between the partial outputs of consecutive slices
    int m_tiles, n_tiles;  // 128-row A tiles; 256-column output tiles
    int n_tiles128;        // 128-row B tiles actually present in the packed operand
    int passes;          // 1 or 3
    float alpha;
    const float* row_add;  // [m] or null
    const float* col_add;  // [n] or null
    // batch of independent products sharing one launch (tile index space = batch x splits x tiles)
    int batch;
    int64_t a_batch_bytes, b_batch_bytes;   // byte strides of the packed operands (0 = shared by the batch)
    int64_t out_batch;                      // element stride of the output
    // TOPK epilogue (MRF): instead of storing the tile, every row keeps its `topk` largest values of the
    // tile's columns (ties: lower column first) -> cand_val / cand_idx [m, n_tiles * 4, topk]; `out` is not written
    int topk;
    float* cand_val;
    int* cand_idx;
    // PACKED epilogue (TOPK = -1): the product leaves as bf16 hi (+ lo) K-major operand tiles [m x n(K)] for the next GEMM
    // (rows = this product's rows, K = its columns) instead of fp32: no fp32 round trip and no pack pass in between
    char* pk_hi;
    char* pk_lo;
    int pk_ktiles;           // ceil(n / 64)
};

__device__ __forceinline__ void gemm_tile(int rem, int m_tiles, int n_tiles, int& mb, int& nb) {
    // concurrently running CTAs (consecutive tile indices) should share the LARGER operand's tiles in L2:
    // walk the short dimension fastest
    if (m_tiles <= n_tiles) { nb = rem / m_tiles; mb = rem % m_tiles; }
    else { mb = rem / n_tiles; nb = rem % n_tiles; }
}

// Persistent: each CTA walks output tiles t = blockIdx.x, +gridDim.x, ... (A-tile-major order so that
// concurrently running CTAs share operand tiles in L2).  Two 256-column TMEM accumulators: the epilogue
// of tile i overlaps the MMAs of tile i+1.
// TOPK 0: store the tile; 4 / 8: top-k epilogue tracking that many values (k <= TOPK); -1: packed operand tiles.
// WIDE: 16 epilogue warps instead of 4 (always with the top-k epilogue; for the storing epilogues when K <= 1024, where the
// epilogue of a tile takes longer than its MMAs).
template <int TOPK, bool WIDE>
__global__ void __launch_bounds__(WIDE ? kGemmTopkThreads : kGemmThreads, 1) gemm_packed_kernel(GemmParams p) {
    constexpr int kSubs = WIDE ? kTopkSubs : 1;
    constexpr int kEpiWarps = 4 * kSubs;
    extern __shared__ unsigned char smem_raw[];
    // SWIZZLE_128B operand tiles must start on a 1024-byte boundary of the shared address space
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t full[kGemmMaxStages], empty[kGemmMaxStages], acc_full[2], acc_empty[2];
    const int parts = p.passes == 3 ? 2 : 1;                  // hi only, or hi + lo
    const int stages = kGemmMaxStages / parts;
    const uint32_t stage_bytes = kGemmPartBytes * parts;
    __shared__ uint32_t tmem_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_per_split = p.m_tiles * p.n_tiles;
    const int tiles_per_batch = tiles_per_split * p.splits;
    const int total_tiles = tiles_per_batch * p.batch;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kGemmMaxStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], kEpiWarps);   // one arrival per epilogue warp
        }
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(&tmem_slot, 2 * kGemmBN);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            const uint64_t pol = policy_evict_last();   // operands are re-read by other CTAs: keep in L2
            uint32_t it = 0;                             // running k-block counter across tiles
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                const int bi = t / tiles_per_batch, tb = t % tiles_per_batch;
                const int z = tb / tiles_per_split, rem = tb % tiles_per_split;
                int mb, nb;
                gemm_tile(rem, p.m_tiles, p.n_tiles, mb, nb);
                const int k_begin = z * p.k_per_split;
                const int k_count = min(p.k_per_split, p.k_tiles - k_begin);
                const int nsub = min(kGemmBTiles, p.n_tiles128 - nb * kGemmBTiles);
                const int64_t a_off = (int64_t)bi * p.a_batch_bytes, b_off = (int64_t)bi * p.b_batch_bytes;
                for (int kb = 0; kb < k_count; ++kb, ++it) {
                    const int s = it % stages;
                    mbar_wait(&empty[s], ((it / stages) & 1) ^ 1);
                    const int kk = k_begin + kb;
                    unsigned char* st = smem + (size_t)s * stage_bytes;
                    mbar_arrive_expect_tx(&full[s], kTileBytes * (1 + nsub) * parts);
                    for (int part = 0; part < parts; ++part) {
                        const __nv_bfloat16* a_src = part ? p.a_lo : p.a_hi;
                        const __nv_bfloat16* b_src = part ? p.b_lo : p.b_hi;
                        unsigned char* sp = st + (size_t)part * kGemmPartBytes;
                        tma_load_1d(sp, reinterpret_cast<const char*>(a_src) + a_off + ((int64_t)mb * p.k_tiles + kk) * kTileBytes,
                                    kTileBytes, &full[s], pol);
                        for (int sub = 0; sub < nsub; ++sub)
                            tma_load_1d(sp + kTileBytes * (1 + sub),
                                        reinterpret_cast<const char*>(b_src) + b_off +
                                            (((int64_t)nb * kGemmBTiles + sub) * p.k_tiles + kk) * kTileBytes,
                                        kTileBytes, &full[s], pol);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            uint32_t it = 0;
            int ti = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++ti) {
                const int tb = t % tiles_per_batch;
                const int z = tb / tiles_per_split, rem = tb % tiles_per_split;
                int mb, nb;
                gemm_tile(rem, p.m_tiles, p.n_tiles, mb, nb);
                (void)mb;
                const int k_begin = z * p.k_per_split;
                const int k_count = min(p.k_per_split, p.k_tiles - k_begin);
                const int nsub = min(kGemmBTiles, p.n_tiles128 - nb * kGemmBTiles);
                const uint32_t idesc = umma_idesc_bf16(128, 128 * nsub);
                const int buf = ti & 1;
                const uint32_t tmem_acc = tmem_base + (uint32_t)(buf * kGemmBN);
                mbar_wait(&acc_empty[buf], ((ti >> 1) & 1) ^ 1);    // epilogue drained this accumulator
                tcgen05_fence_after();
                for (int kb = 0; kb < k_count; ++kb, ++it) {
                    const int s = it % stages;
                    mbar_wait(&full[s], (it / stages) & 1);
                    tcgen05_fence_after();
                    const uint32_t hi_a = smem_u32(smem + (size_t)s * stage_bytes);
                    const uint32_t hi_b = hi_a + kTileBytes;
                    const uint32_t lo_a = hi_a + kGemmPartBytes, lo_b = hi_b + kGemmPartBytes;
#pragma unroll
                    for (int k = 0; k < kTileK / kUmmaK; ++k) {
                        // B rows 128..255 live in the next 16 KiB tile: the 1024-byte group stride still holds
                        const uint32_t ko = k * kUmmaK * 2;
                        umma_bf16_ss(tmem_acc, umma_desc_k_sw128(hi_a + ko), umma_desc_k_sw128(hi_b + ko), idesc,
                                     kb > 0 || k > 0);
                        if (parts == 2) {
                            umma_bf16_ss(tmem_acc, umma_desc_k_sw128(hi_a + ko), umma_desc_k_sw128(lo_b + ko), idesc, true);
                            umma_bf16_ss(tmem_acc, umma_desc_k_sw128(lo_a + ko), umma_desc_k_sw128(hi_b + ko), idesc, true);
                        }
                    }
                    umma_commit(&empty[s]);                           // stage free once these MMAs have read it
                    if (kb == k_count - 1) umma_commit(&acc_full[buf]);   // accumulator complete
                }
            }
        }
    } else {
        // epilogue warps: TMEM lane quarter = warp % 4, column slice = (warp - 2) / 4
        const int q = warp & 3;
        int ti = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++ti) {
            const int bi = t / tiles_per_batch, tb = t % tiles_per_batch;
            const int z = tb / tiles_per_split, rem = tb % tiles_per_split;
            int mb, nb;
            gemm_tile(rem, p.m_tiles, p.n_tiles, mb, nb);
            const int buf = ti & 1;
            float* const out = p.out + (int64_t)z * p.split_stride + (int64_t)bi * p.out_batch;
            mbar_wait(&acc_full[buf], (ti >> 1) & 1);
            tcgen05_fence_after();
            const int64_t row = (int64_t)mb * 128 + q * 32 + lane;
            const float radd = (p.row_add && row < p.m) ? __ldg(p.row_add + row) : 0.f;
            const int ncols = (int)min((int64_t)kGemmBN, p.n - (int64_t)nb * kGemmBN);   // valid columns of this tile
            const int sub = (warp - 2) >> 2;                                  // this warp's column slice of the tile
            const int cbeg = sub * (kGemmBN / kSubs), cend = min(ncols, cbeg + kGemmBN / kSubs);
            if (TOPK > 0) {
                // Top-k of this row over the tile's columns without a divergent insertion (a lane inserts at ~k ln(256/k)
                // of the 256 columns, but SOME lane of the warp inserts at nearly every column, so a conditional
                // insertion chain ran for the whole warp almost every time: 0.67 ms against 0.44 for the materialised
                // path).  Pass 1: branch-free min/max chain on the VALUES only -> the tile's k-th largest value tau and
                // the number of strictly larger ones.  Pass 2 (TMEM read again): emit every x > tau and the first
                // (k - greater) columns with x == tau (ascending column order = the reference's tie order,
                // network/base.py:338-344), unsorted; the merge kernel orders them.
                constexpr int KM = TOPK > 0 ? TOPK : 1;   // levels of the chain: always all of them (a runtime bound put the list
                                                      // into local memory: 137 cycles per element); k only picks tau
                const int kk = p.topk;
                float bv[KM];
#pragma unroll
                for (int r = 0; r < KM; ++r) bv[r] = -INFINITY;
                const uint32_t tacc = tmem_base + (uint32_t)(buf * kGemmBN) + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
                for (int c0 = cbeg; c0 < cend; c0 += 32) {
                    float v[32];
                    tmem_ld_32x32(tacc + (uint32_t)c0, v);
                    const int col0 = nb * kGemmBN + c0;
                    const bool ragged = col0 + 32 > p.n;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        float x = p.alpha * v[j];
                        if (ragged && col0 + j >= p.n) x = -INFINITY;
#pragma unroll
                        for (int r = 0; r < KM; ++r) {
                            const float hi = fmaxf(bv[r], x);
                            x = fminf(bv[r], x);
                            bv[r] = hi;
                        }
                    }
                }
                float tau = -INFINITY;
                int greater = 0;
#pragma unroll
                for (int r = 0; r < KM; ++r) tau = r == kk - 1 ? bv[r] : tau;
#pragma unroll
                for (int r = 0; r < KM; ++r) greater += (r < kk && bv[r] > tau) ? 1 : 0;
                int need_eq = kk - greater, pos = 0;
                const int64_t base = ((((int64_t)bi * p.m + row) * p.n_tiles + nb) * kTopkSubs + sub) * kk;
                const bool live = row < p.m;
#pragma unroll 1
                for (int c0 = cbeg; c0 < cend; c0 += 32) {
                    float v[32];
                    tmem_ld_32x32(tacc + (uint32_t)c0, v);
                    const int col0 = nb * kGemmBN + c0;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float x = p.alpha * v[j];
                        const bool valid = col0 + j < p.n;
                        const bool gt = valid && x > tau;
                        const bool eq = valid && x == tau && need_eq > 0;
                        if ((gt || eq) && live && pos < kk) {
                            p.cand_val[base + pos] = x;
                            p.cand_idx[base + pos] = col0 + j;
                            ++pos;
                            if (eq) --need_eq;
                        }
                    }
                }
                // fewer than k valid columns in this tile: pad with "nothing" entries
                if (live)
                    for (; pos < kk; ++pos) {
                        p.cand_val[base + pos] = -INFINITY;
                        p.cand_idx[base + pos] = 0x7fffffff;
                    }
            } else if (TOPK < 0) {
                // one 16-byte chunk (8 columns) per store at its swizzled place; columns past n are zero (K padding)
                const int r = q * 32 + lane;
                char* hi_rb = p.pk_hi + (int64_t)bi * p.out_batch + (int64_t)mb * p.pk_ktiles * kTileBytes;
                char* lo_rb = p.pk_lo ? p.pk_lo + (int64_t)bi * p.out_batch + (int64_t)mb * p.pk_ktiles * kTileBytes : nullptr;
#pragma unroll 1
                for (int c0 = cbeg; c0 < cbeg + kGemmBN / kSubs; c0 += 32) {
                    const int64_t col0 = (int64_t)nb * kGemmBN + c0;
                    if (col0 >= (int64_t)p.pk_ktiles * kTileK) break;
                    float v[32];
                    tmem_ld_32x32(tmem_base + (uint32_t)(buf * kGemmBN) + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
                    for (int ch = 0; ch < 4; ++ch) {
                        __align__(16) __nv_bfloat16 h[8];
                        __align__(16) __nv_bfloat16 l[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const int64_t col = col0 + ch * 8 + e;
                            const float x = col < p.n ? fmaf(p.alpha, v[ch * 8 + e], radd + (p.col_add ? __ldg(p.col_add + col) : 0.f)) : 0.f;
                            split_bf16(x, h[e], l[e]);
                        }
                        const int64_t kt = col0 / kTileK;
                        const int cidx = (int)((col0 % kTileK) / 8) + ch;
                        const size_t off = (size_t)kt * kTileBytes + tile_chunk_offset(r, cidx);
                        *reinterpret_cast<uint4*>(hi_rb + off) = *reinterpret_cast<const uint4*>(h);
                        if (lo_rb) *reinterpret_cast<uint4*>(lo_rb + off) = *reinterpret_cast<const uint4*>(l);
                    }
                }
            } else {
#pragma unroll 1
            for (int c0 = cbeg; c0 < cend; c0 += 32) {
                float v[32];
                tmem_ld_32x32(tmem_base + (uint32_t)(buf * kGemmBN) + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
                const int64_t col0 = (int64_t)nb * kGemmBN + c0;
                if (row < p.m) {
                    float* dst = out + row * p.ldo + col0;
                    if (col0 + 32 <= p.n && ((reinterpret_cast<uintptr_t>(dst) & 31u) == 0)) {
                        // whole 32-byte sectors per lane (256-bit stores)
#pragma unroll
                        for (int j = 0; j < 32; j += 8) {
                            float o[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e)
                                o[e] = fmaf(p.alpha, v[j + e], radd + (p.col_add ? __ldg(p.col_add + col0 + j + e) : 0.f));
                            asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                                         :: "l"(dst + j), "f"(o[0]), "f"(o[1]), "f"(o[2]), "f"(o[3]), "f"(o[4]), "f"(o[5]), "f"(o[6]),
                                            "f"(o[7]) : "memory");
                        }
                    } else if (col0 + 32 <= p.n && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0)) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            float4 o;
                            o.x = fmaf(p.alpha, v[j + 0], radd + (p.col_add ? __ldg(p.col_add + col0 + j + 0) : 0.f));
                            o.y = fmaf(p.alpha, v[j + 1], radd + (p.col_add ? __ldg(p.col_add + col0 + j + 1) : 0.f));
                            o.z = fmaf(p.alpha, v[j + 2], radd + (p.col_add ? __ldg(p.col_add + col0 + j + 2) : 0.f));
                            o.w = fmaf(p.alpha, v[j + 3], radd + (p.col_add ? __ldg(p.col_add + col0 + j + 3) : 0.f));
                            *reinterpret_cast<float4*>(dst + j) = o;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (col0 + j < p.n)
                                dst[j] = fmaf(p.alpha, v[j], radd + (p.col_add ? __ldg(p.col_add + col0 + j) : 0.f));
                    }
                }
            }
            }
            // hand the accumulator back to the MMA thread
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, 2 * kGemmBN);
    }
}

}  // namespace

// ---- internal host API (used by mrf.cu / wct.cu) -----------------------------------------------
size_t packed_operand_bytes(int64_t rows, int64_t k) {
    const int64_t rt = (rows + kTileRows - 1) / kTileRows, kt = (k + kTileK - 1) / kTileK;
    return (size_t)rt * kt * kTileBytes;
}

int pack_operand_batched(const float* x, int64_t rows, int64_t k, int64_t stride_r, int64_t stride_k,
                         const float* row_scale, const float* row_shift, void* hi, void* lo, int batch, int64_t x_batch,
                         int64_t tile_batch_bytes, int64_t vec_batch, cudaStream_t stream);

int pack_operand_shift(const float* x, int64_t rows, int64_t k, int64_t stride_r, int64_t stride_k,
                       const float* row_scale, const float* row_shift, void* hi, void* lo, cudaStream_t stream) {
    return pack_operand_batched(x, rows, k, stride_r, stride_k, row_scale, row_shift, hi, lo, 1, 0, 0, 0, stream);
}

int pack_operand_impl(const float* x, int64_t rows, int64_t k, int64_t stride_r, int64_t stride_k,
                      const float* row_scale, const float* row_shift, void* hi, void* lo, int batch, int64_t x_batch,
                      int64_t tile_batch_bytes, int64_t vec_batch, int fp16, cudaStream_t stream);

// `batch` independent operands in one launch: sample i reads x + i*x_batch (elements), scale/shift + i*vec_batch,
// and writes its tiles at hi/lo + i*tile_batch_bytes.
int pack_operand_batched(const float* x, int64_t rows, int64_t k, int64_t stride_r, int64_t stride_k,
                         const float* row_scale, const float* row_shift, void* hi, void* lo, int batch, int64_t x_batch,
                         int64_t tile_batch_bytes, int64_t vec_batch, cudaStream_t stream) {
    return pack_operand_impl(x, rows, k, stride_r, stride_k, row_scale, row_shift, hi, lo, batch, x_batch, tile_batch_bytes,
                             vec_batch, 0, stream);
}

// same tile format with IEEE-half elements (values scaled by row_scale first; no lo part)
int pack_operand_batched_f16(const float* x, int64_t rows, int64_t k, int64_t stride_r, int64_t stride_k,
                             const float* row_scale, void* out, int batch, int64_t x_batch, int64_t tile_batch_bytes,
                             int64_t vec_batch, cudaStream_t stream) {
    return pack_operand_impl(x, rows, k, stride_r, stride_k, row_scale, nullptr, out, nullptr, batch, x_batch, tile_batch_bytes,
                             vec_batch, 1, stream);
}

int pack_operand_impl(const float* x, int64_t rows, int64_t k, int64_t stride_r, int64_t stride_k,
                      const float* row_scale, const float* row_shift, void* hi, void* lo, int batch, int64_t x_batch,
                      int64_t tile_batch_bytes, int64_t vec_batch, int fp16, cudaStream_t stream) {
    PackParams p{};
    p.fp16 = fp16;
    p.x_batch = x_batch; p.tile_batch_bytes = tile_batch_bytes; p.vec_batch = vec_batch;
    p.vec_k = (stride_k == 1 && stride_r % 4 == 0 && x_batch % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15u) == 0) ? 1 : 0;
    p.x = x; p.rows = rows; p.k = k; p.stride_r = stride_r; p.stride_k = stride_k; p.row_scale = row_scale;
    p.row_shift = row_shift;
    p.hi = static_cast<__nv_bfloat16*>(hi);
    p.lo = static_cast<__nv_bfloat16*>(lo);
    p.row_tiles = (rows + kTileRows - 1) / kTileRows;
    p.k_tiles = (k + kTileK - 1) / kTileK;
    const int64_t tiles = p.row_tiles * p.k_tiles;
    if (tiles == 0) return RPST_OK;
    RPST_CHECK_ARG(tiles < (1ll << 31), "pack_operand: too many tiles");
    RPST_CHECK_ARG(batch >= 1 && batch < 65536, "pack_operand: bad batch");
    pack_operand_kernel<<<dim3((unsigned)tiles, (unsigned)batch), 256, 0, stream>>>(p);
    RPST_CUDA(cudaGetLastError());
    return RPST_OK;
}

int pack_operand(const float* x, int64_t rows, int64_t k, int64_t stride_r, int64_t stride_k, const float* row_scale,
                 void* hi, void* lo, cudaStream_t stream) {
    return pack_operand_shift(x, rows, k, stride_r, stride_k, row_scale, nullptr, hi, lo, stream);
}

int gemm_packed_splitk(const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo, float* out, int64_t m,
                       int64_t n, int64_t k, int64_t ldo, int passes, float alpha, const float* row_add,
                       const float* col_add, int splits, int64_t split_stride, cudaStream_t stream);

// which epilogue a launch takes besides the plain fp32 store
struct GemmEpilogue {
    int topk = 0;            // > 0: top-k selection into cand_val / cand_idx
    float* cand_val = nullptr;
    int* cand_idx = nullptr;
    char* pk_hi = nullptr;   // non-null: packed bf16 operand tiles
    char* pk_lo = nullptr;
};
static int gemm_packed_launch(const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo, float* out, int64_t m,
                              int64_t n, int64_t k, int64_t ldo, int passes, float alpha, const float* row_add,
                              const float* col_add, int splits, int64_t split_stride, int batch, int64_t a_batch_bytes,
                              int64_t b_batch_bytes, int64_t out_batch, const GemmEpilogue& epi, cudaStream_t stream);

int gemm_packed_batched(const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo, float* out, int64_t m,
                        int64_t n, int64_t k, int64_t ldo, int passes, float alpha, const float* row_add,
                        const float* col_add, int splits, int64_t split_stride, int batch, int64_t a_batch_bytes,
                        int64_t b_batch_bytes, int64_t out_batch, cudaStream_t stream);

int gemm_packed(const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo, float* out, int64_t m,
                int64_t n, int64_t k, int64_t ldo, int passes, float alpha, const float* row_add,
                const float* col_add, cudaStream_t stream) {
    return gemm_packed_splitk(a_hi, a_lo, b_hi, b_lo, out, m, n, k, ldo, passes, alpha, row_add, col_add, 1, 0, stream);
}

// splits > 1: slice z of blockIdx.z reduces k tiles [z*kps, (z+1)*kps) into out + z*split_stride (the
// caller sums the partial outputs; every slice is non-empty).
int gemm_packed_splitk(const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo, float* out, int64_t m,
                       int64_t n, int64_t k, int64_t ldo, int passes, float alpha, const float* row_add,
                       const float* col_add, int splits, int64_t split_stride, cudaStream_t stream) {
    return gemm_packed_batched(a_hi, a_lo, b_hi, b_lo, out, m, n, k, ldo, passes, alpha, row_add, col_add, splits,
                               split_stride, 1, 0, 0, 0, stream);
}

// `batch` independent products in ONE persistent launch (sample i: operands + i*{a,b}_batch_bytes, output +
// i*out_batch elements); a stride of 0 shares that operand across the batch.
int gemm_packed_batched(const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo, float* out, int64_t m,
                        int64_t n, int64_t k, int64_t ldo, int passes, float alpha, const float* row_add,
                        const float* col_add, int splits, int64_t split_stride, int batch, int64_t a_batch_bytes,
                        int64_t b_batch_bytes, int64_t out_batch, cudaStream_t stream) {
    return gemm_packed_launch(a_hi, a_lo, b_hi, b_lo, out, m, n, k, ldo, passes, alpha, row_add, col_add, splits, split_stride, batch,
                              a_batch_bytes, b_batch_bytes, out_batch, GemmEpilogue{}, stream);
}

static int gemm_packed_launch(const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo, float* out, int64_t m,
                              int64_t n, int64_t k, int64_t ldo, int passes, float alpha, const float* row_add,
                              const float* col_add, int splits, int64_t split_stride, int batch, int64_t a_batch_bytes,
                              int64_t b_batch_bytes, int64_t out_batch, const GemmEpilogue& epi, cudaStream_t stream) {
    RPST_CHECK_ARG(passes == 1 || passes == 3, "gemm_packed: passes must be 1 or 3");
    RPST_CHECK_ARG(batch >= 1, "gemm_packed: bad batch");
    RPST_CHECK_ARG(passes == 1 || (a_lo && b_lo), "gemm_packed: bf16x3 needs the lo operands");
    if (m == 0 || n == 0) return RPST_OK;
    RPST_CHECK_ARG(k > 0, "gemm_packed: K must be positive");
    GemmParams p{};
    p.a_hi = static_cast<const __nv_bfloat16*>(a_hi);
    p.a_lo = static_cast<const __nv_bfloat16*>(a_lo);
    p.b_hi = static_cast<const __nv_bfloat16*>(b_hi);
    p.b_lo = static_cast<const __nv_bfloat16*>(b_lo);
    p.out = out; p.m = m; p.n = n; p.ldo = ldo;
    p.k_tiles = (int)((k + kTileK - 1) / kTileK);
    if (splits < 1) splits = 1;
    if (splits > p.k_tiles) splits = p.k_tiles;
    p.k_per_split = (p.k_tiles + splits - 1) / splits;
    splits = (p.k_tiles + p.k_per_split - 1) / p.k_per_split;   // no empty slice
    p.splits = splits;
    p.split_stride = split_stride;
    p.m_tiles = (int)((m + 127) / 128);
    p.n_tiles = (int)((n + kGemmBN - 1) / kGemmBN);
    p.n_tiles128 = (int)((n + 127) / 128);
    p.passes = passes; p.alpha = alpha; p.row_add = row_add; p.col_add = col_add;
    constexpr size_t smem = kGemmSmemBytes;
    static PerDeviceFlag configured_on;
    bool& configured = configured_on.get();
    if (!configured) {
        RPST_CUDA(cudaFuncSetAttribute(gemm_packed_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        RPST_CUDA(cudaFuncSetAttribute(gemm_packed_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        RPST_CUDA(cudaFuncSetAttribute(gemm_packed_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        RPST_CUDA(cudaFuncSetAttribute(gemm_packed_kernel<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        RPST_CUDA(cudaFuncSetAttribute(gemm_packed_kernel<-1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        RPST_CUDA(cudaFuncSetAttribute(gemm_packed_kernel<-1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    p.batch = batch; p.a_batch_bytes = a_batch_bytes; p.b_batch_bytes = b_batch_bytes; p.out_batch = out_batch;
    const int64_t total_tiles = (int64_t)p.m_tiles * p.n_tiles * splits * batch;
    RPST_CHECK_ARG(total_tiles < (1ll << 30), "gemm_packed: too many output tiles");
    int64_t grid = sm_count();
    if (grid > total_tiles) grid = total_tiles;
    const bool wide = p.k_per_split <= 16;
    if (epi.topk > 0) {
        p.topk = epi.topk; p.cand_val = epi.cand_val; p.cand_idx = epi.cand_idx;
        if (p.topk <= 4) gemm_packed_kernel<4, true><<<(unsigned)grid, kGemmTopkThreads, smem, stream>>>(p);
        else gemm_packed_kernel<8, true><<<(unsigned)grid, kGemmTopkThreads, smem, stream>>>(p);
    } else if (epi.pk_hi) {
        p.pk_hi = epi.pk_hi; p.pk_lo = epi.pk_lo; p.pk_ktiles = (int)((n + kTileK - 1) / kTileK);
        if (wide) gemm_packed_kernel<-1, true><<<(unsigned)grid, kGemmTopkThreads, smem, stream>>>(p);
        else gemm_packed_kernel<-1, false><<<(unsigned)grid, kGemmThreads, smem, stream>>>(p);
    } else {
        if (wide) gemm_packed_kernel<0, true><<<(unsigned)grid, kGemmTopkThreads, smem, stream>>>(p);
        else gemm_packed_kernel<0, false><<<(unsigned)grid, kGemmThreads, smem, stream>>>(p);
    }
    RPST_CUDA(cudaGetLastError());
    return RPST_OK;
}

// alpha A.B^T (+ row_add + col_add) written as packed bf16 hi (+ lo) operand tiles [m x n] (K = n) for a following
// gemm_packed call: the tile buffers must hold packed_operand_bytes(m, n); rows past m / columns past n are zero-filled
// by the epilogue for every tile it visits (m rounded up to 128 rows is covered, n up to a multiple of 64).
int gemm_packed_to_tiles(const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo, int64_t m, int64_t n, int64_t k,
                         int passes, float alpha, void* out_hi, void* out_lo, cudaStream_t stream) {
    RPST_CHECK_ARG(out_hi != nullptr, "gemm_packed_to_tiles: null output");
    GemmEpilogue epi;
    epi.pk_hi = static_cast<char*>(out_hi); epi.pk_lo = static_cast<char*>(out_lo);
    return gemm_packed_launch(a_hi, a_lo, b_hi, b_lo, nullptr, m, n, k, n, passes, alpha, nullptr, nullptr, 1, 0, 1, 0, 0, 0, epi, stream);
}

// lists of candidates per row that gemm_packed_topk emits for n columns
int gemm_topk_lists(int64_t n) { return (int)((n + kGemmBN - 1) / kGemmBN) * kTopkSubs; }

// A.B^T with the per-row top-k epilogue: candidates [m, gemm_topk_lists(n), topk] (value, column), unsorted inside a
// list; the tile itself is never stored.  alpha = -1 ranks the negated product (MRF `reverse`).
int gemm_packed_topk(const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo, int64_t m, int64_t n, int64_t k,
                     int passes, float alpha, int topk, float* cand_val, int* cand_idx, cudaStream_t stream) {
    RPST_CHECK_ARG(topk >= 1 && topk <= 8 && cand_val && cand_idx, "gemm_packed_topk: bad top-k request");
    GemmEpilogue epi;
    epi.topk = topk; epi.cand_val = cand_val; epi.cand_idx = cand_idx;
    return gemm_packed_launch(a_hi, a_lo, b_hi, b_lo, nullptr, m, n, k, n, passes, alpha, nullptr, nullptr, 1, 0, 1, 0, 0, 0, epi, stream);
}

}  // namespace rpst

using namespace rpst;

extern "C" size_t rpst_packed_operand_bytes(int64_t rows, int64_t k) { return packed_operand_bytes(rows, k); }

extern "C" int rpst_pack_operand(const float* x, int64_t rows, int64_t k, int64_t stride_r, int64_t stride_k,
                                 const float* row_scale, void* hi, void* lo, void* stream) {
    RPST_CHECK_ARG(rows >= 0 && k >= 0, "pack_operand: negative size");
    RPST_CHECK_ARG(rows == 0 || k == 0 || (x && hi), "pack_operand: null pointer");
    RPST_CHECK_ARG((reinterpret_cast<uintptr_t>(hi) & 127u) == 0 && (reinterpret_cast<uintptr_t>(lo) & 127u) == 0,
                   "pack_operand: tile buffers must be 128-byte aligned");
    return pack_operand(x, rows, k, stride_r, stride_k, row_scale, hi, lo, static_cast<cudaStream_t>(stream));
}

extern "C" int rpst_gemm_packed(const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo, float* out,
                                int64_t m, int64_t n, int64_t k, int64_t ldo, int passes, float alpha,
                                const float* row_add, const float* col_add, void* stream) {
    RPST_CHECK_ARG(m >= 0 && n >= 0 && k >= 0 && ldo >= n, "gemm_packed: bad shape");
    RPST_CHECK_ARG(m == 0 || n == 0 || (a_hi && b_hi && out), "gemm_packed: null pointer");
    return gemm_packed(a_hi, a_lo, b_hi, b_lo, out, m, n, k, ldo, passes, alpha, row_add, col_add,
                       static_cast<cudaStream_t>(stream));
}

RPST_WATCHDOG_SETTER(gemm)
