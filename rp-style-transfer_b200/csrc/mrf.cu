// MRF patch matching — SURVEY.md §8 a12/a13/a14; reference: network/base.py:317-360 (cal_affinity_map,
// cal_dist) and network/mrf_rp.py:12-23 (MRFLoss.forward).
//
//   1. mrf_norm_kernel     per-position channel L2 norms of content / style (F.normalize, eps 1e-12)
//   2. pack_operand        normalised features -> bf16 hi/lo K-major tiles (positions x channels)
//   3. gemm_packed_topk x2 NCC = c^T s  and  NCC^T = s^T c on the tensor cores, bf16x3 (fp32-grade) so that top-k
//                          INDICES agree with the reference on tie-free inputs; the L x L products are NEVER stored
//                          (SURVEY 2b K10): the GEMM epilogue keeps every row's running top-k (k <= 8) of each
//                          64-column slice of a tile straight out of TMEM (two branch-free passes) -> [L, L/64, k] candidates
//   4. topk_merge_kernel   warp per row: merges the per-tile candidate lists (values descending, ties -> lower index)
//   5. mrf_loss_kernel     sum over the union of both top-k sets of ||a_i - b_j||^2, evaluated sparsely
//                          from the NCC values (no L x L affinity / distance temporaries), fp64 tree
//   (optional) mrf_scatter_kernel  the dense binary [L,L] affinity map the reference returns
#include "common.cuh"

namespace rpst {

size_t packed_operand_bytes(int64_t rows, int64_t k);
int pack_operand(const float* x, int64_t rows, int64_t k, int64_t stride_r, int64_t stride_k, const float* row_scale,
                 void* hi, void* lo, cudaStream_t stream);
int gemm_packed(const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo, float* out, int64_t m,
                int64_t n, int64_t k, int64_t ldo, int passes, float alpha, const float* row_add,
                const float* col_add, cudaStream_t stream);
int gemm_topk_lists(int64_t n);
int gemm_packed_topk(const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo, int64_t m, int64_t n, int64_t k,
                     int passes, float alpha, int topk, float* cand_val, int* cand_idx, cudaStream_t stream);

namespace {

constexpr int kMaxTopK = 8;

// x [c, l] -> sq[l] = sum_c x^2, nrm[l] = max(sqrt(sq), 1e-12), inv[l] = 1/nrm.
// Block = 32 positions (lanes) x 8 channel slices (warps): every load is one 128-byte line, 8 independent lines in flight
// per warp; the slices meet in shared memory and are added in slice order.
__global__ void __launch_bounds__(256) mrf_norm_kernel(const float* __restrict__ x, int64_t c, int64_t l,
                                                       float* __restrict__ sq, float* __restrict__ nrm,
                                                       float* __restrict__ inv) {
    __shared__ float red[8][33];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t pos = blockIdx.x * 32ll + lane;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (pos < l) {
        int64_t ch = w;
        for (; ch + 24 < c; ch += 32) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float v = __ldg(x + (ch + 8 * u) * l + pos);
                acc[u] = fmaf(v, v, acc[u]);
            }
        }
        for (; ch < c; ch += 8) {
            const float v = __ldg(x + ch * l + pos);
            acc[0] = fmaf(v, v, acc[0]);
        }
    }
    red[w][lane] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
    __syncthreads();
    if (w == 0 && pos < l) {
        float s = red[0][lane];
#pragma unroll
        for (int q = 1; q < 8; ++q) s += red[q][lane];
        const float n = fmaxf(sqrtf(s), 1e-12f);
        sq[pos] = s;
        nrm[pos] = n;
        inv[pos] = 1.f / n;
    }
}

// Merge of the per-tile candidate lists of the GEMM's top-k epilogue: cand_* [rows, nt, K], every list sorted (value
// descending, ties -> lower column).  One warp per row; lane t folds tiles t, t + 32, ... into its own sorted list,
// then K rounds of warp arg-max over the list heads.  Values sorted descending, ties towards the lower column index;
// idx_out[row*stride_row + r*stride_k], val_out likewise.
template <int K>
__global__ void __launch_bounds__(256) topk_merge_kernel(const float* __restrict__ cand_val, const int* __restrict__ cand_idx,
                                                         int64_t rows, int nt, int64_t* __restrict__ idx_out,
                                                         float* __restrict__ val_out, int64_t stride_row, int64_t stride_k) {
    const int lane = threadIdx.x & 31;
    const int64_t row = blockIdx.x * (int64_t)(blockDim.x / 32) + (threadIdx.x >> 5);
    if (row >= rows) return;
    float v[K];
    int id[K];
#pragma unroll
    for (int r = 0; r < K; ++r) { v[r] = -INFINITY; id[r] = 0x7fffffff; }
    for (int t = lane; t < nt; t += 32) {
        const float* cv = cand_val + (row * nt + t) * K;
        const int* ci = cand_idx + (row * nt + t) * K;
#pragma unroll
        for (int e = 0; e < K; ++e) {
            float x = __ldg(cv + e);
            int xi = __ldg(ci + e);
            if (x > v[K - 1] || (x == v[K - 1] && xi < id[K - 1])) {
#pragma unroll
                for (int r = 0; r < K; ++r) {
                    const bool before = x > v[r] || (x == v[r] && xi < id[r]);
                    if (before) {
                        const float tv = v[r]; const int ti = id[r];
                        v[r] = x; id[r] = xi; x = tv; xi = ti;
                    }
                }
            }
        }
    }
    for (int r = 0; r < K; ++r) {
        float bv = v[0];
        int bi = id[0];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (id[0] == bi && v[0] == bv) {  // the winning lane pops its head (indices are unique)
#pragma unroll
            for (int q = 0; q < K - 1; ++q) { v[q] = v[q + 1]; id[q] = id[q + 1]; }
            v[K - 1] = -INFINITY; id[K - 1] = 0x7fffffff;
        }
        if (lane == 0) {
            idx_out[row * stride_row + r * stride_k] = bi;
            val_out[row * stride_row + r * stride_k] = bv;
        }
    }
}

struct LossParams {
    const int64_t* idx1;   // [l, k]  (content i -> style j)
    const float* val1;
    const int64_t* idx0;   // [k, l]  (style j -> content i)
    const float* val0;
    const float* a_sq; const float* a_nrm;   // content
    const float* b_sq; const float* b_nrm;   // style
    int64_t l;
    int k;
    float sign;            // -1 if the map was negated (reverse=True)
    double scale;          // 1/(l*k) or 1/(l*l)
    float* loss;
    double* partial;       // [kLossBlocks]
    int* ticket;           // zero on entry, reset by the kernel
};

// dist_ij = |a_i|^2 + |b_j|^2 - 2 a_i.b_j with a_i.b_j = ncc_ij * |a_i| * |b_j|.  kLossBlocks blocks write fp64 partial
// sums, the last one to finish (ticket) adds them in block order: deterministic, one launch.
constexpr int kLossBlocks = 64;
__global__ void __launch_bounds__(256) mrf_loss_kernel(LossParams p) {
    __shared__ double red[8];
    __shared__ bool last;
    double acc = 0.0;
    const int64_t n = p.l * p.k;
    for (int64_t t = blockIdx.x * 256ll + threadIdx.x; t < n; t += kLossBlocks * 256ll) {
        {   // row set: (i, idx1[i][r])
            const int64_t i = t / p.k;
            const int64_t j = p.idx1[t];
            const double dot = (double)(p.sign * p.val1[t]) * (double)p.a_nrm[i] * (double)p.b_nrm[j];
            acc += (double)p.a_sq[i] + (double)p.b_sq[j] - 2.0 * dot;
        }
        {   // column set: (idx0[r][j], j) unless already counted in the row set of that i
            const int64_t j = t % p.l;
            const int64_t i = p.idx0[t];
            bool dup = false;
            for (int r = 0; r < p.k; ++r) dup |= (p.idx1[i * p.k + r] == j);
            if (!dup) {
                const double dot = (double)(p.sign * p.val0[t]) * (double)p.a_nrm[i] * (double)p.b_nrm[j];
                acc += (double)p.a_sq[i] + (double)p.b_sq[j] - 2.0 * dot;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double v = 0.0;
        for (int q = 0; q < 8; ++q) v += red[q];
        p.partial[blockIdx.x] = v;
        __threadfence();
        last = atomicAdd(p.ticket, 1) == kLossBlocks - 1;
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence();
        double v = 0.0;
        for (int q = 0; q < kLossBlocks; ++q) v += __ldcg(p.partial + q);
        *p.loss = (float)(v * p.scale);
        *p.ticket = 0;                       // ready for the next call on this workspace
    }
}

__global__ void mrf_scatter_kernel(const int64_t* __restrict__ idx1, const int64_t* __restrict__ idx0, int64_t l, int k,
                                   float* __restrict__ aff) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= l * k) return;
    aff[(t / k) * l + idx1[t]] = 1.f;        // scatter_(1, index, 1)
    aff[idx0[t] * l + (t % l)] = 1.f;        // scatter_(0, index, 1)
}

struct MrfLayout {
    size_t tiles, vec;             // bytes of one packed operand / one [l] float vector
    size_t off_tiles[4];           // c_hi, c_lo, s_hi, s_lo
    size_t off_vec[6];             // c_sq, c_nrm, c_inv, s_sq, s_nrm, s_inv
    size_t off_cval, off_cidx, off_val1, off_val0, off_loss, total;    // top-k candidates [l, gemm_topk_lists(l), k] (shared by both products)
    int nt;
};

MrfLayout mrf_layout(int64_t c, int64_t l, int k) {
    MrfLayout m;
    m.tiles = packed_operand_bytes(l, c);
    m.vec = align_up((size_t)l * sizeof(float), 256);
    size_t o = 0;
    for (int i = 0; i < 4; ++i) { m.off_tiles[i] = o; o += align_up(m.tiles, 256); }
    for (int i = 0; i < 6; ++i) { m.off_vec[i] = o; o += m.vec; }
    m.nt = gemm_topk_lists(l);
    m.off_cval = o; o += align_up((size_t)l * m.nt * k * sizeof(float), 256);
    m.off_cidx = o; o += align_up((size_t)l * m.nt * k * sizeof(int), 256);
    m.off_val1 = o; o += align_up((size_t)l * k * sizeof(float), 256);
    m.off_val0 = o; o += align_up((size_t)l * k * sizeof(float), 256);
    m.off_loss = o; o += 1024;                 // 64 fp64 partial sums + the ticket
    m.total = o;
    return m;
}

template <int K>
int launch_merge(const float* cv, const int* ci, int64_t rows, int nt, int64_t* idx, float* val, int64_t sr, int64_t sk,
                 cudaStream_t st) {
    topk_merge_kernel<K><<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(cv, ci, rows, nt, idx, val, sr, sk);
    RPST_CUDA(cudaGetLastError());
    return RPST_OK;
}

int merge_dispatch(int k, const float* cv, const int* ci, int64_t rows, int nt, int64_t* idx, float* val, int64_t sr,
                   int64_t sk, cudaStream_t st) {
    switch (k) {
        case 1: return launch_merge<1>(cv, ci, rows, nt, idx, val, sr, sk, st);
        case 2: return launch_merge<2>(cv, ci, rows, nt, idx, val, sr, sk, st);
        case 3: return launch_merge<3>(cv, ci, rows, nt, idx, val, sr, sk, st);
        case 4: return launch_merge<4>(cv, ci, rows, nt, idx, val, sr, sk, st);
        case 5: return launch_merge<5>(cv, ci, rows, nt, idx, val, sr, sk, st);
        case 6: return launch_merge<6>(cv, ci, rows, nt, idx, val, sr, sk, st);
        case 7: return launch_merge<7>(cv, ci, rows, nt, idx, val, sr, sk, st);
        default: return launch_merge<8>(cv, ci, rows, nt, idx, val, sr, sk, st);
    }
}

}  // namespace

// per-position channel norms (shared with the adaptive SANet cosine affinity)
int channel_norms(const float* x, int64_t c, int64_t l, float* sq, float* nrm, float* inv, cudaStream_t st) {
    mrf_norm_kernel<<<(unsigned)((l + 31) / 32), 256, 0, st>>>(x, c, l, sq, nrm, inv);
    RPST_CUDA(cudaGetLastError());
    return RPST_OK;
}

}  // namespace rpst

using namespace rpst;

extern "C" size_t rpst_mrf_workspace_bytes(int64_t c, int64_t l, int k) {
    if (c <= 0 || l <= 0 || k <= 0) return 256;
    return mrf_layout(c, l, k).total;
}

extern "C" int rpst_mrf_match(const float* content, const float* style, int64_t c, int64_t l, int k, int reverse,
                              int passes, int64_t* idx_dim0, int64_t* idx_dim1, float* affinity, float* loss,
                              int loss_mean_over_all, void* workspace, size_t workspace_bytes, void* stream) {
    RPST_CHECK_ARG(c > 0 && l > 0, "mrf: empty input");
    RPST_CHECK_ARG(k >= 1 && k <= kMaxTopK && k <= l, "mrf: k must be in [1, %d] and <= H*W (got %d)", kMaxTopK, k);
    RPST_CHECK_ARG(l < (1ll << 31), "mrf: too many positions");
    RPST_CHECK_ARG(content && style && idx_dim0 && idx_dim1, "mrf: null pointer");
    RPST_CHECK_ARG(passes == 1 || passes == 3, "mrf: passes must be 1 (bf16) or 3 (bf16x3, fp32-grade)");
    const MrfLayout m = mrf_layout(c, l, k);
    if (!workspace || workspace_bytes < m.total) {
        set_error("mrf: workspace too small (%zu < %zu bytes)", workspace_bytes, m.total);
        return RPST_ERR_WORKSPACE;
    }
    RPST_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, "mrf: workspace must be 256-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    char* w = static_cast<char*>(workspace);
    float* vec[6];
    for (int i = 0; i < 6; ++i) vec[i] = reinterpret_cast<float*>(w + m.off_vec[i]);
    const unsigned nb = (unsigned)((l + 31) / 32);
    mrf_norm_kernel<<<nb, 256, 0, st>>>(content, c, l, vec[0], vec[1], vec[2]);
    mrf_norm_kernel<<<nb, 256, 0, st>>>(style, c, l, vec[3], vec[4], vec[5]);
    RPST_CUDA(cudaGetLastError());
    void* c_hi = w + m.off_tiles[0]; void* c_lo = w + m.off_tiles[1];
    void* s_hi = w + m.off_tiles[2]; void* s_lo = w + m.off_tiles[3];
    int rc = pack_operand(content, l, c, 1, l, vec[2], c_hi, passes == 3 ? c_lo : nullptr, st);
    if (rc) return rc;
    rc = pack_operand(style, l, c, 1, l, vec[5], s_hi, passes == 3 ? s_lo : nullptr, st);
    if (rc) return rc;
    float* cval = reinterpret_cast<float*>(w + m.off_cval);
    int* cidx = reinterpret_cast<int*>(w + m.off_cidx);
    const float alpha = reverse ? -1.f : 1.f;
    float* val1 = reinterpret_cast<float*>(w + m.off_val1);
    float* val0 = reinterpret_cast<float*>(w + m.off_val0);
    // NCC rows (content i -> style j: torch.topk(dim=1)) and NCC^T rows (style j -> content i: dim=0); the candidate
    // buffers are reused by the second product (stream order)
    rc = gemm_packed_topk(c_hi, c_lo, s_hi, s_lo, l, l, c, passes, alpha, k, cval, cidx, st);
    if (rc) return rc;
    rc = merge_dispatch(k, cval, cidx, l, m.nt, idx_dim1, val1, k, 1, st);      // [l, k]
    if (rc) return rc;
    rc = gemm_packed_topk(s_hi, s_lo, c_hi, c_lo, l, l, c, passes, alpha, k, cval, cidx, st);
    if (rc) return rc;
    rc = merge_dispatch(k, cval, cidx, l, m.nt, idx_dim0, val0, 1, l, st);      // [k, l]
    if (rc) return rc;
    if (affinity) {
        RPST_CUDA(cudaMemsetAsync(affinity, 0, (size_t)l * l * sizeof(float), st));
        mrf_scatter_kernel<<<(unsigned)((l * k + 255) / 256), 256, 0, st>>>(idx_dim1, idx_dim0, l, k, affinity);
        RPST_CUDA(cudaGetLastError());
    }
    if (loss) {
        LossParams lp{};
        lp.idx1 = idx_dim1; lp.val1 = val1; lp.idx0 = idx_dim0; lp.val0 = val0;
        lp.a_sq = vec[0]; lp.a_nrm = vec[1]; lp.b_sq = vec[3]; lp.b_nrm = vec[4];
        lp.l = l; lp.k = k; lp.sign = alpha;
        lp.scale = loss_mean_over_all ? 1.0 / ((double)l * (double)l) : 1.0 / ((double)l * (double)k);
        lp.loss = loss;
        lp.partial = reinterpret_cast<double*>(w + m.off_loss);
        lp.ticket = reinterpret_cast<int*>(w + m.off_loss + 768);
        RPST_CUDA(cudaMemsetAsync(lp.ticket, 0, sizeof(int), st));
        mrf_loss_kernel<<<kLossBlocks, 256, 0, st>>>(lp);
        RPST_CUDA(cudaGetLastError());
    }
    return RPST_OK;
}

extern "C" size_t rpst_pairwise_sqdist_workspace_bytes(int64_t d, int64_t m, int64_t n) {
    if (d <= 0 || m <= 0 || n <= 0) return 256;
    return 2 * (align_up(packed_operand_bytes(m, d), 256) + align_up(packed_operand_bytes(n, d), 256)) +
           6 * align_up((size_t)(m > n ? m : n) * sizeof(float), 256);
}

extern "C" int rpst_pairwise_sqdist(const float* a, const float* b, int64_t d, int64_t m, int64_t n, float* out,
                                    void* workspace, size_t workspace_bytes, void* stream) {
    RPST_CHECK_ARG(d > 0 && m > 0 && n > 0, "pairwise_sqdist: empty input");
    RPST_CHECK_ARG(a && b && out, "pairwise_sqdist: null pointer");
    const size_t need = rpst_pairwise_sqdist_workspace_bytes(d, m, n);
    if (!workspace || workspace_bytes < need) {
        set_error("pairwise_sqdist: workspace too small (%zu < %zu bytes)", workspace_bytes, need);
        return RPST_ERR_WORKSPACE;
    }
    RPST_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, "pairwise_sqdist: workspace must be 256-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    char* w = static_cast<char*>(workspace);
    const size_t ta = align_up(packed_operand_bytes(m, d), 256), tb = align_up(packed_operand_bytes(n, d), 256);
    const size_t vb = align_up((size_t)(m > n ? m : n) * sizeof(float), 256);
    void* a_hi = w; void* a_lo = w + ta; void* b_hi = w + 2 * ta; void* b_lo = w + 2 * ta + tb;
    float* v = reinterpret_cast<float*>(w + 2 * ta + 2 * tb);
    float* a_sq = v; float* a_t1 = v + vb / 4; float* a_t2 = v + 2 * vb / 4;
    float* b_sq = v + 3 * vb / 4; float* b_t1 = v + 4 * vb / 4; float* b_t2 = v + 5 * vb / 4;
    mrf_norm_kernel<<<(unsigned)((m + 31) / 32), 256, 0, st>>>(a, d, m, a_sq, a_t1, a_t2);
    mrf_norm_kernel<<<(unsigned)((n + 31) / 32), 256, 0, st>>>(b, d, n, b_sq, b_t1, b_t2);
    RPST_CUDA(cudaGetLastError());
    int rc = pack_operand(a, m, d, 1, m, nullptr, a_hi, a_lo, st);
    if (rc) return rc;
    rc = pack_operand(b, n, d, 1, n, nullptr, b_hi, b_lo, st);
    if (rc) return rc;
    // network/base.py:356-359: |a_i|^2 + |b_j|^2 - 2 a_i.b_j
    return gemm_packed(a_hi, a_lo, b_hi, b_lo, out, m, n, d, n, 3, -2.f, a_sq, b_sq, st);
}
