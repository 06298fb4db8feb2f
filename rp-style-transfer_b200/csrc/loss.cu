// Loss statistics: the reference's `calc_style_loss`
// (network/adain_rp.py:84-88, network/sanet.py:232-236: two calc_mean_std + two MSE, i.e. >= 12 launches
// and two full reads per tensor) and `calc_content_loss(norm=True)` (network/sanet.py:226-230:
// two mean_variance_norm materialisations + an MSE = 6 tensor passes) in ONE streaming pass over the
// pair (x, y): 2*E*4 algorithmic bytes, HBM-bound.
//
// pair_moments_kernel accumulates, per 8192-element chunk, the first and second CENTRED moments of the
// pair: (n, mean_x, M2_x, mean_y, M2_y, C_xy) with Chan's pairwise update extended to the co-moment.
// pair_finalize_kernel merges the chunk records of a plane in fp64 (fixed order => deterministic) and
// emits per-plane statistics plus the two batch losses
//     style   = mean_p (mu_x-mu_y)^2 + mean_p (sd_x-sd_y)^2            (sd = sqrt(M2/(HW-1)+eps))
//     content = 1/(P*HW) * sum_p [ M2_x/sd_x^2 + M2_y/sd_y^2 - 2 C_xy/(sd_x sd_y) ]
// the latter being sum((x-mu_x)/sd_x - (y-mu_y)/sd_y)^2 expanded, so the normalised tensors are never
// materialised.  The last block to finish (ticket) adds the block partials in index order.
// plane_affine2_kernel is the matching backward pass: out = ax[p]*x + ay[p]*y + b[p].
#include "common.cuh"
#include "plane_io.cuh"

namespace rpst {
namespace {

constexpr int kLossThreads = 256;
constexpr int kLossVecs = 8;                                  // vectors per thread per tensor
struct PairRec {                                              // 32 bytes, written as two float4
    float n, mean_x, m2_x, mean_y;
    float m2_y, c_xy, pad0, pad1;
};

struct PairM {
    float n, mx, m2x, my, m2y, cxy;
};

__device__ __forceinline__ PairM pair_merge(const PairM& a, const PairM& b) {
    PairM r;
    r.n = a.n + b.n;
    const float inv = r.n > 0.f ? 1.f / r.n : 0.f;
    const float w = b.n * inv;
    const float dx = b.mx - a.mx, dy = b.my - a.my;
    const float k = a.n * w;
    r.mx = fmaf(dx, w, a.mx);
    r.my = fmaf(dy, w, a.my);
    r.m2x = a.m2x + b.m2x + dx * dx * k;
    r.m2y = a.m2y + b.m2y + dy * dy * k;
    r.cxy = a.cxy + b.cxy + dx * dy * k;
    return r;
}

__device__ __forceinline__ PairM pair_shfl_xor(const PairM& m, int o) {
    PairM r;
    r.n = __shfl_xor_sync(0xffffffffu, m.n, o);
    r.mx = __shfl_xor_sync(0xffffffffu, m.mx, o);
    r.m2x = __shfl_xor_sync(0xffffffffu, m.m2x, o);
    r.my = __shfl_xor_sync(0xffffffffu, m.my, o);
    r.m2y = __shfl_xor_sync(0xffffffffu, m.m2y, o);
    r.cxy = __shfl_xor_sync(0xffffffffu, m.cxy, o);
    return r;
}

struct PairParams {
    const float* x;
    const float* y;
    int64_t planes, hw;
    int cpp;              // chunks per plane
    int64_t items;
    PairRec* part;        // [planes*cpp]
    float eps;
    float* stats;         // [planes,8] (mu_x, sd_x, mu_y, sd_y, M2_x, M2_y, C_xy, 0) or null
    float* losses;        // [2] style, content-norm
    double* block_part;   // [finalize blocks][2]
    unsigned* ticket;     // zeroed before the launch
};

struct PairD {
    double n, mx, m2x, my, m2y, cxy;
};
__device__ __forceinline__ PairD pair_merge(const PairD& a, const PairD& b) {
    PairD r;
    r.n = a.n + b.n;
    const double inv = r.n > 0.0 ? 1.0 / r.n : 0.0;
    const double w = b.n * inv;
    const double dx = b.mx - a.mx, dy = b.my - a.my;
    const double k = a.n * w;
    r.mx = a.mx + dx * w;
    r.my = a.my + dy * w;
    r.m2x = a.m2x + b.m2x + dx * dx * k;
    r.m2y = a.m2y + b.m2y + dy * dy * k;
    r.cxy = a.cxy + b.cxy + dx * dy * k;
    return r;
}
__device__ __forceinline__ PairD pair_shfl_xor(const PairD& m, int o) {
    PairD r;
    r.n = __shfl_xor_sync(0xffffffffu, m.n, o);
    r.mx = __shfl_xor_sync(0xffffffffu, m.mx, o);
    r.m2x = __shfl_xor_sync(0xffffffffu, m.m2x, o);
    r.my = __shfl_xor_sync(0xffffffffu, m.my, o);
    r.m2y = __shfl_xor_sync(0xffffffffu, m.m2y, o);
    r.cxy = __shfl_xor_sync(0xffffffffu, m.cxy, o);
    return r;
}

// per-plane loss terms from the merged moments (fp64: the content term cancels when x ~ y)
__device__ __forceinline__ void pair_terms(const PairParams& p, const PairD& m, int64_t plane, bool write_stats, double& style_term,
                                           double& content_term) {
    const double denom = (double)p.hw - 1.0;  // HW==1 -> 0/0 = NaN like torch.var
    const double vx = m.m2x / denom + (double)p.eps, vy = m.m2y / denom + (double)p.eps;
    const double sdx = sqrt(vx), sdy = sqrt(vy);
    const double dm = m.mx - m.my, ds = sdx - sdy;
    style_term = dm * dm + ds * ds;
    content_term = m.m2x / vx + m.m2y / vy - 2.0 * m.cxy / (sdx * sdy);
    if (write_stats && p.stats) {
        float4* st = reinterpret_cast<float4*>(p.stats + plane * 8);
        st[0] = make_float4((float)m.mx, (float)sdx, (float)m.my, (float)sdy);
        st[1] = make_float4((float)m.m2x, (float)m.m2y, (float)m.cxy, 0.f);
    }
}

// last block (ticket): fixed-order sum of the block partials -> the two losses
__device__ __forceinline__ void pair_last_block_sum(const PairParams& p, unsigned blocks, double* s_a, double* s_b) {
    double a = 0.0, b = 0.0;
    for (unsigned i = threadIdx.x; i < blocks; i += 256) {
        a += __ldcg(p.block_part + 2 * i);
        b += __ldcg(p.block_part + 2 * i + 1);
    }
    s_a[threadIdx.x] = a; s_b[threadIdx.x] = b;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) { s_a[threadIdx.x] += s_a[threadIdx.x + s]; s_b[threadIdx.x] += s_b[threadIdx.x + s]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        p.losses[0] = (float)(s_a[0] / (double)p.planes);
        p.losses[1] = (float)(s_b[0] / ((double)p.planes * (double)p.hw));
    }
}

// VECS = vectors per thread and tensor: 8 for the general case; 4 for planes that fit half a chunk (64x64 VGG relu4_1
// planes): half the registers, so five CTAs instead of three are resident per SM and hide the load latency that
// bounds small planes (the chunk geometry, one record per plane, is the same).
// FUSED (one chunk per plane and at most 256 planes per CTA): the CTA keeps its planes' records in shared memory and,
// after its last item, finalizes them itself (one thread per plane, block partial, last-block ticket) — no second
// kernel, which on 64x64 planes was a quarter of the time.
template <int VEC, int VECS, bool FUSED>
__global__ void __launch_bounds__(kLossThreads, VECS == kLossVecs ? 3 : 5) pair_moments_kernel(PairParams p) {
    constexpr int CHUNK = kLossThreads * kLossVecs * VEC;
    static_assert(kLossThreads == 256, "finalize helpers assume 256 threads");
    __shared__ PairM s_warp[kLossThreads / 32];
    __shared__ PairM s_rec[FUSED ? kLossThreads : 1];
    int my_items = 0;
    const uint64_t pol = policy_evict_first();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int64_t item = blockIdx.x; item < p.items; item += gridDim.x) {
        const int64_t plane = item / p.cpp;
        const int chunk = (int)(item % p.cpp);
        const int64_t e0 = (int64_t)chunk * CHUNK;
        const int64_t rem = p.hw - e0;
        const int nvec = (int)((rem < CHUNK ? rem : CHUNK) / VEC);
        const float* xb = p.x + plane * p.hw + e0;
        const float* yb = p.y + plane * p.hw + e0;
        float xv[VECS][VEC], yv[VECS][VEC];
#pragma unroll
        for (int j = 0; j < VECS; ++j) {
            const int idx = j * kLossThreads + threadIdx.x;
            if (idx < nvec) {
                load_vec<VEC>(xv[j], xb + (int64_t)idx * VEC, pol, true);
            } else {
#pragma unroll
                for (int e = 0; e < VEC; ++e) xv[j][e] = 0.f;
            }
        }
#pragma unroll
        for (int j = 0; j < VECS; ++j) {
            const int idx = j * kLossThreads + threadIdx.x;
            if (idx < nvec) {
                load_vec<VEC>(yv[j], yb + (int64_t)idx * VEC, pol, true);
            } else {
#pragma unroll
                for (int e = 0; e < VEC; ++e) yv[j][e] = 0.f;
            }
        }
        // exact two-pass moments of this thread's registers
        float sx = 0.f, sy = 0.f;
        int cnt = 0;
#pragma unroll
        for (int j = 0; j < VECS; ++j) {
            if (j * kLossThreads + (int)threadIdx.x < nvec) {
#pragma unroll
                for (int e = 0; e < VEC; ++e) { sx += xv[j][e]; sy += yv[j][e]; }
                cnt += VEC;
            }
        }
        PairM m = {(float)cnt, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (cnt > 0) {
            m.mx = sx / m.n;
            m.my = sy / m.n;
#pragma unroll
            for (int j = 0; j < VECS; ++j) {
                if (j * kLossThreads + (int)threadIdx.x < nvec) {
#pragma unroll
                    for (int e = 0; e < VEC; ++e) {
                        const float dx = xv[j][e] - m.mx, dy = yv[j][e] - m.my;
                        m.m2x = fmaf(dx, dx, m.m2x);
                        m.m2y = fmaf(dy, dy, m.m2y);
                        m.cxy = fmaf(dx, dy, m.cxy);
                    }
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = pair_merge(m, pair_shfl_xor(m, o));
        if (lane == 0) s_warp[warp] = m;
        __syncthreads();
        if (warp == 0) {
            PairM t = lane < kLossThreads / 32 ? s_warp[lane] : PairM{0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) t = pair_merge(t, pair_shfl_xor(t, o));
            if (lane == 0) {
                if constexpr (FUSED) {
                    s_rec[my_items] = t;
                } else {
                    float4* dst = reinterpret_cast<float4*>(p.part + item);
                    dst[0] = make_float4(t.n, t.mx, t.m2x, t.my);
                    dst[1] = make_float4(t.m2y, t.cxy, 0.f, 0.f);
                }
            }
        }
        ++my_items;
        __syncthreads();
    }
    if constexpr (FUSED) {
        __shared__ double s_a[256], s_b[256];
        __shared__ bool s_last;
        double style_term = 0.0, content_term = 0.0;
        if ((int)threadIdx.x < my_items) {
            const PairM r = s_rec[threadIdx.x];
            const PairD m = {(double)r.n, (double)r.mx, (double)r.m2x, (double)r.my, (double)r.m2y, (double)r.cxy};
            pair_terms(p, m, (int64_t)blockIdx.x + (int64_t)threadIdx.x * gridDim.x, true, style_term, content_term);
        }
        s_a[threadIdx.x] = style_term; s_b[threadIdx.x] = content_term;
        __syncthreads();
        for (int s = 128; s > 0; s >>= 1) {
            if ((int)threadIdx.x < s) { s_a[threadIdx.x] += s_a[threadIdx.x + s]; s_b[threadIdx.x] += s_b[threadIdx.x + s]; }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            p.block_part[2 * blockIdx.x] = s_a[0];
            p.block_part[2 * blockIdx.x + 1] = s_b[0];
            __threadfence();
            s_last = atomicAdd(p.ticket, 1u) == gridDim.x - 1;
        }
        __syncthreads();
        if (!s_last) return;
        __threadfence();
        pair_last_block_sum(p, gridDim.x, s_a, s_b);
    }
}

// one warp per plane; block partial sums -> last block (ticket) adds them in index order
__global__ void __launch_bounds__(256) pair_finalize_kernel(PairParams p) {
    __shared__ double s_style[8], s_content[8];
    __shared__ bool s_last;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t plane = (int64_t)blockIdx.x * 8 + warp;
    double style_term = 0.0, content_term = 0.0;
    if (plane < p.planes) {
        PairD m = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        const float4* recs = reinterpret_cast<const float4*>(p.part + plane * p.cpp);
        for (int k = lane; k < p.cpp; k += 32) {
            const float4 a = __ldcg(recs + 2 * k), b = __ldcg(recs + 2 * k + 1);
            const PairD r = {(double)a.x, (double)a.y, (double)a.z, (double)a.w, (double)b.x, (double)b.y};
            m = pair_merge(m, r);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = pair_merge(m, pair_shfl_xor(m, o));
        pair_terms(p, m, plane, lane == 0, style_term, content_term);
    }
    if (lane == 0) { s_style[warp] = style_term; s_content[warp] = content_term; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < 8; ++w) { a += s_style[w]; b += s_content[w]; }
        p.block_part[2 * blockIdx.x] = a;
        p.block_part[2 * blockIdx.x + 1] = b;
        __threadfence();
        s_last = atomicAdd(p.ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    __shared__ double s_a[256], s_b[256];
    pair_last_block_sum(p, gridDim.x, s_a, s_b);
}

template <int VEC>
__global__ void __launch_bounds__(256) plane_affine2_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                            const float* __restrict__ ax, const float* __restrict__ ay,
                                                            const float* __restrict__ b, float* __restrict__ out,
                                                            int64_t planes, int64_t hw, int cpp) {
    constexpr int CHUNK = 256 * kPerThread * VEC;
    const uint64_t pol = policy_evict_first();
    const int64_t items = planes * cpp;
    for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
        const int64_t plane = item / cpp;
        const int64_t e0 = (item % cpp) * (int64_t)CHUNK;
        const int64_t rem = hw - e0;
        const int nvec = (int)((rem < CHUNK ? rem : CHUNK) / VEC);
        const float kx = __ldg(ax + plane), ky = __ldg(ay + plane), kb = b ? __ldg(b + plane) : 0.f;
        const float* xb = x + plane * hw + e0;
        const float* yb = y + plane * hw + e0;
        float* ob = out + plane * hw + e0;
#pragma unroll
        for (int bt = 0; bt < kBatches; ++bt) {
            float xv[kBatch][VEC], yv[kBatch][VEC];
            load_batch<VEC, 256>(xv, xb, bt, nvec, pol, true);
            load_batch<VEC, 256>(yv, yb, bt, nvec, pol, true);
#pragma unroll
            for (int j = 0; j < kBatch; ++j) {
                const int idx = (bt * kBatch + j) * 256 + threadIdx.x;
                if (idx < nvec) {
                    float o[VEC];
#pragma unroll
                    for (int e = 0; e < VEC; ++e) o[e] = fmaf(kx, xv[j][e], fmaf(ky, yv[j][e], kb));
                    store_vec<VEC>(ob + (int64_t)idx * VEC, o, pol, true);
                }
            }
        }
    }
}

// Backward of the two losses w.r.t. one tensor of the pair, coefficients derived per plane from the
// saved statistics (no host-side tensor algebra):  out = ax*u + ay*v + b, u = the tensor being
// differentiated, v = the other one, g = upstream gradient (device scalar), P planes, n = HW.
//   style  : d/du_i = 2g/P [ (mu_u-mu_v)/n + (sd_u-sd_v)(u_i-mu_u)/((n-1) sd_u) ]            (ay = 0, v not read)
//   content: d/du_i = 2g/(P n sd_u) [ u^_i - v^_i - u^_i K ],  K = (M2_u/sd_u^2 - C/(sd_u sd_v))/(n-1)
template <int VEC>
__global__ void __launch_bounds__(256) pair_loss_bwd_kernel(const float* __restrict__ u, const float* __restrict__ v,
                                                            const float* __restrict__ stats, const float* __restrict__ grad,
                                                            int which, int wrt_second, float* __restrict__ out,
                                                            int64_t planes, int64_t hw, int cpp) {
    constexpr int CHUNK = 256 * kPerThread * VEC;
    const uint64_t pol = policy_evict_first();
    const int64_t items = planes * cpp;
    const float g = __ldg(grad);
    const float n = (float)hw, P = (float)planes;
    for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
        const int64_t plane = item / cpp;
        const int64_t e0 = (item % cpp) * (int64_t)CHUNK;
        const int64_t rem = hw - e0;
        const int nvec = (int)((rem < CHUNK ? rem : CHUNK) / VEC);
        const float4 s0 = __ldg(reinterpret_cast<const float4*>(stats + plane * 8));
        const float4 s1 = __ldg(reinterpret_cast<const float4*>(stats + plane * 8) + 1);
        // (mu, sd, M2) of u and of v
        const float mu_u = wrt_second ? s0.z : s0.x, sd_u = wrt_second ? s0.w : s0.y, m2_u = wrt_second ? s1.y : s1.x;
        const float mu_v = wrt_second ? s0.x : s0.z, sd_v = wrt_second ? s0.y : s0.w;
        float ax, ay, b;
        if (which == 0) {
            ax = 2.f * g * (sd_u - sd_v) / (P * (n - 1.f) * sd_u);
            ay = 0.f;
            b = 2.f * g * (mu_u - mu_v) / (P * n) - ax * mu_u;
        } else {
            const float sc = 2.f * g / (P * n);
            const float k = (m2_u / (sd_u * sd_u) - s1.z / (sd_u * sd_v)) / (n - 1.f);
            ax = sc * (1.f - k) / (sd_u * sd_u);
            ay = -sc / (sd_u * sd_v);
            b = -ax * mu_u - ay * mu_v;
        }
        const float* ub = u + plane * hw + e0;
        const float* vb = v + plane * hw + e0;
        float* ob = out + plane * hw + e0;
#pragma unroll
        for (int bt = 0; bt < kBatches; ++bt) {
            float uv[kBatch][VEC], vv[kBatch][VEC];
            load_batch<VEC, 256>(uv, ub, bt, nvec, pol, true);
            if (which != 0) load_batch<VEC, 256>(vv, vb, bt, nvec, pol, true);
#pragma unroll
            for (int j = 0; j < kBatch; ++j) {
                const int idx = (bt * kBatch + j) * 256 + threadIdx.x;
                if (idx < nvec) {
                    float o[VEC];
#pragma unroll
                    for (int e = 0; e < VEC; ++e) o[e] = which != 0 ? fmaf(ax, uv[j][e], fmaf(ay, vv[j][e], b)) : fmaf(ax, uv[j][e], b);
                    store_vec<VEC>(ob + (int64_t)idx * VEC, o, pol, true);
                }
            }
        }
    }
}

constexpr int64_t kFusedMaxGrid = 4096;   // CTAs of the fused moments+finalize launch (block partials reserved for them)
struct PairLayout {
    size_t part_off, block_off, total;
    int cpp;
    int64_t fin_blocks;
};
PairLayout pair_layout(int64_t planes, int64_t hw, int vec) {
    PairLayout l;
    const int64_t chunk = (int64_t)kLossThreads * kLossVecs * vec;
    l.cpp = (int)((hw + chunk - 1) / chunk);
    l.fin_blocks = (planes + 7) / 8;
    l.part_off = 256;
    l.block_off = align_up(l.part_off + (size_t)planes * l.cpp * sizeof(PairRec), 256);
    l.total = align_up(l.block_off + (size_t)(l.fin_blocks + kFusedMaxGrid) * 2 * sizeof(double), 256);
    return l;
}

}  // namespace
}  // namespace rpst

using namespace rpst;

extern "C" size_t rpst_pair_stats_workspace_bytes(int64_t planes, int64_t hw) {
    if (planes <= 0 || hw <= 0) return 256;
    return pair_layout(planes, hw, 1).total;  // scalar chunks are the smaller ones: upper bound for both paths
}

extern "C" int rpst_pair_stats(const float* x, const float* y, int64_t planes, int64_t hw, float eps, float* stats,
                               float* losses, void* workspace, size_t workspace_bytes, void* stream) {
    RPST_CHECK_ARG(planes > 0 && hw > 0, "pair_stats: empty input (planes=%lld, hw=%lld)", (long long)planes, (long long)hw);
    RPST_CHECK_ARG(x && y && losses, "pair_stats: null pointer");
    RPST_CHECK_ARG(stats == nullptr || aligned16(stats), "pair_stats: stats must be 16-byte aligned");
    const bool vec = hw % 4 == 0 && aligned16(x) && aligned16(y);
    const PairLayout l = pair_layout(planes, hw, vec ? 4 : 1);
    if (workspace == nullptr || workspace_bytes < l.total) {
        set_error("pair_stats: workspace too small (%zu < %zu bytes)", workspace_bytes, l.total);
        return RPST_ERR_WORKSPACE;
    }
    RPST_CHECK_ARG(aligned16(workspace), "pair_stats: workspace must be 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    char* base = static_cast<char*>(workspace);
    PairParams p{};
    p.x = x; p.y = y; p.planes = planes; p.hw = hw; p.cpp = l.cpp; p.items = planes * l.cpp;
    p.part = reinterpret_cast<PairRec*>(base + l.part_off);
    p.eps = eps; p.stats = stats; p.losses = losses;
    p.block_part = reinterpret_cast<double*>(base + l.block_off);
    p.ticket = reinterpret_cast<unsigned*>(base);
    RPST_CHECK_ARG(l.fin_blocks < (1ll << 31), "pair_stats: too many planes");
    RPST_CUDA(cudaMemsetAsync(base, 0, 256, st));
    // persistent grid = exactly the co-resident CTAs (64 KiB of loads in flight each)
    const bool half = hw <= (int64_t)kLossThreads * (kLossVecs / 2) * (vec ? 4 : 1);   // the plane fits half a chunk
    static int per_sm[2][2] = {{0, 0}, {0, 0}};
    int& nbs = per_sm[vec ? 1 : 0][half ? 1 : 0];
    if (nbs == 0) {
        int nb = 0;
        if (vec && half) RPST_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, pair_moments_kernel<4, kLossVecs / 2, true>, kLossThreads, 0));
        else if (vec) RPST_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, pair_moments_kernel<4, kLossVecs, true>, kLossThreads, 0));
        else if (half) RPST_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, pair_moments_kernel<1, kLossVecs / 2, true>, kLossThreads, 0));
        else RPST_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, pair_moments_kernel<1, kLossVecs, true>, kLossThreads, 0));
        nbs = nb > 0 ? nb : 1;
    }
    int64_t grid = (int64_t)sm_count() * nbs;
    if (grid > p.items) grid = p.items;
    // one chunk per plane: the moments kernel finalizes its own planes (no second launch)
    const bool fused = l.cpp == 1 && grid <= kFusedMaxGrid && p.items <= grid * (int64_t)kLossThreads;
    if (fused) {
        if (vec && half) pair_moments_kernel<4, kLossVecs / 2, true><<<(int)grid, kLossThreads, 0, st>>>(p);
        else if (vec) pair_moments_kernel<4, kLossVecs, true><<<(int)grid, kLossThreads, 0, st>>>(p);
        else if (half) pair_moments_kernel<1, kLossVecs / 2, true><<<(int)grid, kLossThreads, 0, st>>>(p);
        else pair_moments_kernel<1, kLossVecs, true><<<(int)grid, kLossThreads, 0, st>>>(p);
        RPST_CUDA(cudaGetLastError());
        return RPST_OK;
    }
    if (vec && half) pair_moments_kernel<4, kLossVecs / 2, false><<<(int)grid, kLossThreads, 0, st>>>(p);
    else if (vec) pair_moments_kernel<4, kLossVecs, false><<<(int)grid, kLossThreads, 0, st>>>(p);
    else if (half) pair_moments_kernel<1, kLossVecs / 2, false><<<(int)grid, kLossThreads, 0, st>>>(p);
    else pair_moments_kernel<1, kLossVecs, false><<<(int)grid, kLossThreads, 0, st>>>(p);
    RPST_CUDA(cudaGetLastError());
    pair_finalize_kernel<<<(unsigned)l.fin_blocks, 256, 0, st>>>(p);
    RPST_CUDA(cudaGetLastError());
    return RPST_OK;
}

extern "C" int rpst_plane_affine2(const float* x, const float* y, const float* ax, const float* ay, const float* b,
                                  float* out, int64_t planes, int64_t hw, void* stream) {
    RPST_CHECK_ARG(planes >= 0 && hw >= 0, "plane_affine2: negative size");
    if (planes == 0 || hw == 0) return RPST_OK;
    RPST_CHECK_ARG(x && y && ax && ay && out, "plane_affine2: null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool vec = hw % 4 == 0 && aligned16(x) && aligned16(y) && aligned16(out);
    const int64_t chunk = 256ll * kPerThread * (vec ? 4 : 1);
    const int cpp = (int)((hw + chunk - 1) / chunk);
    int64_t items = planes * cpp;
    int64_t grid = (int64_t)sm_count() * 8;
    if (grid > items) grid = items;
    if (vec) plane_affine2_kernel<4><<<(int)grid, 256, 0, st>>>(x, y, ax, ay, b, out, planes, hw, cpp);
    else plane_affine2_kernel<1><<<(int)grid, 256, 0, st>>>(x, y, ax, ay, b, out, planes, hw, cpp);
    RPST_CUDA(cudaGetLastError());
    return RPST_OK;
}

extern "C" int rpst_pair_loss_bwd(const float* x, const float* y, const float* stats, const float* grad, int which,
                                  int wrt_second, float* out, int64_t planes, int64_t hw, void* stream) {
    RPST_CHECK_ARG(planes >= 0 && hw >= 0, "pair_loss_bwd: negative size");
    if (planes == 0 || hw == 0) return RPST_OK;
    RPST_CHECK_ARG(x && y && stats && grad && out, "pair_loss_bwd: null pointer");
    RPST_CHECK_ARG(which == 0 || which == 1, "pair_loss_bwd: which must be 0 (style) or 1 (normalised content)");
    RPST_CHECK_ARG(aligned16(stats), "pair_loss_bwd: stats must be 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const float* u = wrt_second ? y : x;
    const float* v = wrt_second ? x : y;
    const bool vec = hw % 4 == 0 && aligned16(u) && aligned16(v) && aligned16(out);
    const int64_t chunk = 256ll * kPerThread * (vec ? 4 : 1);
    const int cpp = (int)((hw + chunk - 1) / chunk);
    int64_t items = planes * cpp;
    int64_t grid = (int64_t)sm_count() * 8;
    if (grid > items) grid = items;
    if (vec) pair_loss_bwd_kernel<4><<<(int)grid, 256, 0, st>>>(u, v, stats, grad, which, wrt_second ? 1 : 0, out, planes, hw, cpp);
    else pair_loss_bwd_kernel<1><<<(int)grid, 256, 0, st>>>(u, v, stats, grad, which, wrt_second ? 1 : 0, out, planes, hw, cpp);
    RPST_CUDA(cudaGetLastError());
    return RPST_OK;
}
