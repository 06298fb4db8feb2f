// AdaIN family: per-(n,c) plane statistics, normalise-affine,
// multiscale blend, concat-write — hand-written for sm_100a, HBM-bound by design.
//
// Two kernels:
//  * adain_direct_kernel  — planes that fit one CTA's registers (<= 32 floats/thread, <= 512 threads): one CTA per
//    plane, content is read ONCE from HBM and stays in registers between the statistics and the
//    apply phase.
//  * adain_pipe_kernel    — larger planes (512x512 = 1 MiB, 1024x2048 = 8 MiB ...): a persistent
//    kernel whose CTAs pull 32 KiB work items from a global ticket counter.  Statistics items of
//    plane p+D are interleaved with apply items of plane p, so that the second read of the content
//    (apply) is served by the 126 MB L2 (the content lines are loaded with an evict_last policy,
//    every other stream with evict_first) and HBM sees each tensor exactly once: the algorithmic
//    3*E*4 (AdaIN) / 4*E*4 (blend) bytes of SURVEY.md §8d.  An apply item waits on a per-plane
//    release/acquire flag set by the CTA that finished the plane's last statistics item; items
//    are issued in ticket order and statistics items never block, so the wait cannot deadlock.
//
// Numerics: unbiased variance, eps inside the sqrt (network/base.py:404-405);
// moments are accumulated as (count, mean, M2) with a two-pass evaluation over each thread's
// registers and Chan merges above that, never as raw sum/sum-of-squares.
#include "common.cuh"

namespace rpst {
namespace {

constexpr int kBatch = 4;          // vectors per load batch
constexpr int kBatches = 2;        // batches per thread per item
constexpr int kPerThread = kBatch * kBatches;
constexpr int kPipeThreads = 256;

struct Tuning {
    int64_t lag_bytes = 16ll << 20;
    int64_t hints = 1;
    int64_t ctas_per_sm = 3;
};
Tuning g_tuning;

struct AdainParams {
    const float* content;
    const float* style;  // may be null
    const float* prev;   // may be null
    float* out;
    float* mean_out;     // stats-only mode
    float* std_out;
    float* saved;        // [planes,4] or null
    int64_t planes, hw, channels, out_batch_stride;
    float eps;
    int stats_only;
    int hints;
    // pipelined kernel only
    int ipp;             // items (chunks) per plane
    int lag;             // planes between statistics and apply
    unsigned total_items;
    unsigned* ticket;
    int* done;           // [planes]
    int* ready;          // [planes]
    float4* coef;        // [planes] (mu_c, a, mu_s, unused)
    float2* part_c;      // [planes*ipp] (mean, m2)
    float2* part_s;
};

template <int VEC>
struct VecT;
template <>
struct VecT<4> {
    using type = float4;
};
template <>
struct VecT<1> {
    using type = float;
};

template <int VEC>
__device__ __forceinline__ void load_vec(float (&dst)[VEC], const float* p, uint64_t pol, bool hint) {
    if constexpr (VEC == 4) {
        float4 v = hint ? ldg_f4_hint(p, pol) : __ldg(reinterpret_cast<const float4*>(p));
        dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
    } else {
        dst[0] = hint ? ldg_f1_hint(p, pol) : __ldg(p);
    }
}
template <int VEC>
__device__ __forceinline__ void store_vec(float* p, const float (&src)[VEC], uint64_t pol, bool hint) {
    if constexpr (VEC == 4) {
        float4 v = make_float4(src[0], src[1], src[2], src[3]);
        if (hint) stg_f4_hint(p, v, pol); else *reinterpret_cast<float4*>(p) = v;
    } else {
        if (hint) stg_f1_hint(p, src[0], pol); else *p = src[0];
    }
}

// Load one batch (kBatch vectors, strided by THREADS vectors) of a chunk that holds `nvec` vectors.
template <int VEC, int THREADS>
__device__ __forceinline__ void load_batch(float (&v)[kBatch][VEC], const float* base, int batch, int nvec,
                                           uint64_t pol, bool hint) {
#pragma unroll
    for (int j = 0; j < kBatch; ++j) {
        int idx = (batch * kBatch + j) * THREADS + threadIdx.x;
        if (idx < nvec) {
            load_vec<VEC>(v[j], base + (int64_t)idx * VEC, pol, hint);
        } else {
#pragma unroll
            for (int e = 0; e < VEC; ++e) v[j][e] = 0.f;
        }
    }
}

// Exact two-pass moments of the valid part of a register batch, merged into `acc`.
template <int VEC, int THREADS>
__device__ __forceinline__ void batch_moments(Moments& acc, const float (&v)[kBatch][VEC], int batch, int nvec) {
    float sum = 0.f;
    int cnt = 0;
#pragma unroll
    for (int j = 0; j < kBatch; ++j) {
        int idx = (batch * kBatch + j) * THREADS + threadIdx.x;
        if (idx < nvec) {
#pragma unroll
            for (int e = 0; e < VEC; ++e) sum += v[j][e];
            cnt += VEC;
        }
    }
    if (cnt == 0) return;
    Moments m;
    m.n = (float)cnt;
    m.mean = sum / m.n;
    float m2 = 0.f;
#pragma unroll
    for (int j = 0; j < kBatch; ++j) {
        int idx = (batch * kBatch + j) * THREADS + threadIdx.x;
        if (idx < nvec) {
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                float d = v[j][e] - m.mean;
                m2 = fmaf(d, d, m2);
            }
        }
    }
    m.m2 = m2;
    acc = merge(acc, m);
}

__device__ __forceinline__ float std_from(const Moments& m, float hw, float eps) {
    // unbiased: divide by HW-1 (0/0 -> NaN for HW==1, like torch.var)
    return sqrtf(m.m2 / (hw - 1.f) + eps);
}

// ------------------------------------------------------------------------------------------
// direct kernel: one CTA per plane, content stays in registers
// ------------------------------------------------------------------------------------------
template <int VEC, int THREADS>
__global__ void __launch_bounds__(THREADS) adain_direct_kernel(AdainParams p) {
    __shared__ Moments scratch[2][32];
    const uint64_t pol_first = policy_evict_first();
    const bool hint = p.hints != 0;
    const int nvec = (int)(p.hw / VEC);
    const float hwf = (float)p.hw;

    for (int64_t plane = blockIdx.x; plane < p.planes; plane += gridDim.x) {
        const float* cbase = p.content + plane * p.hw;
        float c[kBatches][kBatch][VEC];
        Moments mc = {0.f, 0.f, 0.f}, ms = {0.f, 0.f, 0.f};
#pragma unroll
        for (int b = 0; b < kBatches; ++b) load_batch<VEC, THREADS>(c[b], cbase, b, nvec, pol_first, hint);
        if (p.style != nullptr) {
            const float* sbase = p.style + plane * p.hw;
#pragma unroll
            for (int b = 0; b < kBatches; ++b) {
                float s[kBatch][VEC];
                load_batch<VEC, THREADS>(s, sbase, b, nvec, pol_first, hint);
                batch_moments<VEC, THREADS>(ms, s, b, nvec);
            }
        }
#pragma unroll
        for (int b = 0; b < kBatches; ++b) batch_moments<VEC, THREADS>(mc, c[b], b, nvec);

        mc = block_merge<THREADS>(mc, scratch[0]);
        float mu_c = mc.mean, sd_c = std_from(mc, hwf, p.eps);
        float mu_s = 0.f, sd_s = 1.f;
        if (p.style != nullptr) {
            ms = block_merge<THREADS>(ms, scratch[1]);
            mu_s = ms.mean;
            sd_s = std_from(ms, hwf, p.eps);
        }
        if (threadIdx.x == 0) {
            if (p.mean_out) p.mean_out[plane] = mu_c;
            if (p.std_out) p.std_out[plane] = sd_c;
            if (p.saved) reinterpret_cast<float4*>(p.saved)[plane] = make_float4(mu_c, sd_c, mu_s, sd_s);
        }
        if (!p.stats_only) {
            const float a = sd_s / sd_c;
            const int64_t n_idx = plane / p.channels, ch = plane % p.channels;
            float* obase = p.out + n_idx * p.out_batch_stride + ch * p.hw;
            const float* pbase = p.prev ? p.prev + plane * p.hw : nullptr;
#pragma unroll
            for (int b = 0; b < kBatches; ++b) {
                float pv[kBatch][VEC];
                if (pbase) load_batch<VEC, THREADS>(pv, pbase, b, nvec, pol_first, hint);
#pragma unroll
                for (int j = 0; j < kBatch; ++j) {
                    int idx = (b * kBatch + j) * THREADS + threadIdx.x;
                    if (idx < nvec) {
                        float o[VEC];
#pragma unroll
                        for (int e = 0; e < VEC; ++e) {
                            float y = fmaf(c[b][j][e] - mu_c, a, mu_s);
                            o[e] = pbase ? y + pv[j][e] : y;
                        }
                        store_vec<VEC>(obase + (int64_t)idx * VEC, o, pol_first, hint);
                    }
                }
            }
        }
        __syncthreads();  // scratch reuse across the grid-stride loop
    }
}

// ------------------------------------------------------------------------------------------
// pipelined kernel: persistent CTAs + ticket counter, statistics D planes ahead of apply
// ------------------------------------------------------------------------------------------
struct Item {
    int kind;  // 0 statistics, 1 apply
    int64_t plane;
    int chunk;
};

__device__ __forceinline__ Item decode_ticket(unsigned t, const AdainParams& p) {
    Item it;
    const int64_t ipp = p.ipp;
    if (p.stats_only) {
        it.kind = 0; it.plane = t / ipp; it.chunk = (int)(t % ipp);
        return it;
    }
    const int64_t lag = p.lag < p.planes ? p.lag : p.planes;
    int64_t tt = t;
    if (tt < lag * ipp) {  // prologue: statistics only
        it.kind = 0; it.plane = tt / ipp; it.chunk = (int)(tt % ipp);
        return it;
    }
    tt -= lag * ipp;
    const int64_t steady = (p.planes - lag) * 2 * ipp;
    if (tt < steady) {  // steady state: alternate statistics(plane r+lag) / apply(plane r)
        int64_t r = tt / (2 * ipp), u = tt % (2 * ipp);
        it.kind = (int)(u & 1);
        it.chunk = (int)(u >> 1);
        it.plane = it.kind ? r : r + lag;
        return it;
    }
    tt -= steady;  // epilogue: apply only
    it.kind = 1; it.plane = (p.planes - lag) + tt / ipp; it.chunk = (int)(tt % ipp);
    return it;
}

template <int VEC, int MINB>
__global__ void __launch_bounds__(kPipeThreads, MINB) adain_pipe_kernel(AdainParams p) {
    constexpr int T = kPipeThreads;
    constexpr int CHUNK = T * kPerThread * VEC;
    __shared__ Moments scratch[2][32];
    __shared__ unsigned s_ticket;
    __shared__ int s_last;
    __shared__ float4 s_coef;

    const uint64_t pol_first = policy_evict_first();
    const uint64_t pol_last = policy_evict_last();
    const bool hint = p.hints != 0;
    const float hwf = (float)p.hw;

    if (threadIdx.x == 0) s_ticket = atomicAdd(p.ticket, 1u);
    __syncthreads();
    unsigned t = s_ticket;

    while (t < p.total_items) {
        unsigned next = 0;
        if (threadIdx.x == 0) next = atomicAdd(p.ticket, 1u);  // latency hidden behind the item

        const Item it = decode_ticket(t, p);
        const int64_t e0 = (int64_t)it.chunk * CHUNK;
        const int64_t rem = p.hw - e0;
        const int nvec = (int)((rem < CHUNK ? rem : CHUNK) / VEC);
        const float* cbase = p.content + it.plane * p.hw + e0;

        if (it.kind == 0) {
            // ---------------- statistics item
            Moments mc = {0.f, 0.f, 0.f}, ms = {0.f, 0.f, 0.f};
            const uint64_t cpol = p.stats_only ? pol_first : pol_last;
            {
                float c[kBatches][kBatch][VEC];
#pragma unroll
                for (int b = 0; b < kBatches; ++b) load_batch<VEC, T>(c[b], cbase, b, nvec, cpol, hint);
#pragma unroll
                for (int b = 0; b < kBatches; ++b) batch_moments<VEC, T>(mc, c[b], b, nvec);
            }
            if (p.style != nullptr) {
                const float* sbase = p.style + it.plane * p.hw + e0;
                float s[kBatches][kBatch][VEC];
#pragma unroll
                for (int b = 0; b < kBatches; ++b) load_batch<VEC, T>(s[b], sbase, b, nvec, pol_first, hint);
#pragma unroll
                for (int b = 0; b < kBatches; ++b) batch_moments<VEC, T>(ms, s[b], b, nvec);
                ms = block_merge<T>(ms, scratch[1]);
            }
            mc = block_merge<T>(mc, scratch[0]);

            if (threadIdx.x == 0) {
                const int64_t slot = it.plane * p.ipp + it.chunk;
                __stcg(&p.part_c[slot], make_float2(mc.mean, mc.m2));
                if (p.style != nullptr) __stcg(&p.part_s[slot], make_float2(ms.mean, ms.m2));
                __threadfence();
                int old = atomicAdd(&p.done[it.plane], 1);
                s_last = (old == p.ipp - 1);
            }
            __syncthreads();
            if (s_last && threadIdx.x < 32) {
                // last statistics item of this plane: merge the chunk partials (one warp)
                __threadfence();
                Moments tc = {0.f, 0.f, 0.f}, ts = {0.f, 0.f, 0.f};
                for (int k = threadIdx.x; k < p.ipp; k += 32) {
                    int64_t r = p.hw - (int64_t)k * CHUNK;
                    float n = (float)(r < CHUNK ? r : CHUNK);
                    float2 pc = __ldcg(&p.part_c[it.plane * p.ipp + k]);
                    tc = merge(tc, Moments{n, pc.x, pc.y});
                    if (p.style != nullptr) {
                        float2 ps = __ldcg(&p.part_s[it.plane * p.ipp + k]);
                        ts = merge(ts, Moments{n, ps.x, ps.y});
                    }
                }
                tc = warp_merge(tc);
                ts = warp_merge(ts);
                if (threadIdx.x == 0) {
                    float mu_c = tc.mean, sd_c = std_from(tc, hwf, p.eps);
                    float mu_s = 0.f, sd_s = 1.f;
                    if (p.style != nullptr) { mu_s = ts.mean; sd_s = std_from(ts, hwf, p.eps); }
                    if (p.mean_out) p.mean_out[it.plane] = mu_c;
                    if (p.std_out) p.std_out[it.plane] = sd_c;
                    if (p.saved) reinterpret_cast<float4*>(p.saved)[it.plane] = make_float4(mu_c, sd_c, mu_s, sd_s);
                    if (!p.stats_only) {
                        __stcg(&p.coef[it.plane], make_float4(mu_c, sd_s / sd_c, mu_s, 0.f));
                        __threadfence();
                        st_release(&p.ready[it.plane], 1);
                    }
                }
            }
        } else {
            // ---------------- apply item: content comes back from L2, prev streams from HBM
            const float* pbase = p.prev ? p.prev + it.plane * p.hw + e0 : nullptr;
            const int64_t n_idx = it.plane / p.channels, ch = it.plane % p.channels;
            float* obase = p.out + n_idx * p.out_batch_stride + ch * p.hw + e0;
            float c[kBatch][VEC], pv[kBatch][VEC];
            load_batch<VEC, T>(c, cbase, 0, nvec, pol_first, hint);
            if (pbase) load_batch<VEC, T>(pv, pbase, 0, nvec, pol_first, hint);
            if (threadIdx.x == 0) {
                while (ld_acquire(&p.ready[it.plane]) == 0) __nanosleep(100);
                s_coef = __ldcg(&p.coef[it.plane]);
            }
            __syncthreads();
            const float mu_c = s_coef.x, a = s_coef.y, mu_s = s_coef.z;
#pragma unroll
            for (int b = 0; b < kBatches; ++b) {
                if (b > 0) {
                    load_batch<VEC, T>(c, cbase, b, nvec, pol_first, hint);
                    if (pbase) load_batch<VEC, T>(pv, pbase, b, nvec, pol_first, hint);
                }
#pragma unroll
                for (int j = 0; j < kBatch; ++j) {
                    int idx = (b * kBatch + j) * T + threadIdx.x;
                    if (idx < nvec) {
                        float o[VEC];
#pragma unroll
                        for (int e = 0; e < VEC; ++e) {
                            float y = fmaf(c[j][e] - mu_c, a, mu_s);
                            o[e] = pbase ? y + pv[j][e] : y;
                        }
                        store_vec<VEC>(obase + (int64_t)idx * VEC, o, pol_first, hint);
                    }
                }
            }
        }
        __syncthreads();  // everyone is done with scratch / s_last / s_coef / s_ticket
        if (threadIdx.x == 0) s_ticket = next;
        __syncthreads();
        t = s_ticket;
    }
}

// ------------------------------------------------------------------------------------------
// backward.  With x^ = (c-mu_c)/sd_c, G1 = sum(dy), G2 = sum(dy*x^) over the plane (HW = n):
//   dc = (sd_s/sd_c) * (dy - G1/n - x^ * G2/(n-1))        (unbiased variance => n-1)
//   ds = G1/n + (s-mu_s)/((n-1)*sd_s) * G2
// Same two-phase structure as the forward: reduce items (dy, c from HBM, kept in L2) run D planes
// ahead of write items (dy, c from L2; s from HBM; dc, ds to HBM): 5*E*4 algorithmic bytes.
// ------------------------------------------------------------------------------------------
struct BwdParams {
    const float* dy;
    const float* content;
    const float* style;   // may be null
    const float4* saved;  // [planes] (mu_c, sd_c, mu_s, sd_s)
    float* dcontent;
    float* dstyle;        // may be null
    int64_t planes, hw;
    int hints;
    int ipp, lag;
    unsigned total_items;
    unsigned* ticket;
    int* done;
    int* ready;
    float2* sums;         // [planes] (G1, G2)
    float2* part;         // [planes*ipp]
};

template <int THREADS>
__device__ __forceinline__ float2 block_sum2(float2 v, float2* scratch) {
    constexpr int W = THREADS / 32;
    v.x = warp_sum(v.x);
    v.y = warp_sum(v.y);
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    float2 t = scratch[0];
#pragma unroll
    for (int w = 1; w < W; ++w) { t.x += scratch[w].x; t.y += scratch[w].y; }
    return t;
}

template <int VEC, int THREADS>
__device__ __forceinline__ void bwd_write_batch(const float (&dy)[kBatch][VEC], const float (&c)[kBatch][VEC],
                                                const float* sbase, float* dcbase, float* dsbase, int b, int nvec,
                                                float4 st, float2 g, float hwf, uint64_t pol, bool hint) {
    const float inv_sd_c = 1.f / st.y;
    const float a = st.w * inv_sd_c;
    const float g1n = g.x / hwf;
    const float g2n = g.y / (hwf - 1.f);
    float sv[kBatch][VEC];
    if (dsbase) load_batch<VEC, THREADS>(sv, sbase, b, nvec, pol, hint);
#pragma unroll
    for (int j = 0; j < kBatch; ++j) {
        int idx = (b * kBatch + j) * THREADS + threadIdx.x;
        if (idx < nvec) {
            float o[VEC];
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                float xh = (c[j][e] - st.x) * inv_sd_c;
                o[e] = a * (dy[j][e] - g1n - xh * g2n);
            }
            store_vec<VEC>(dcbase + (int64_t)idx * VEC, o, pol, hint);
            if (dsbase) {
                const float k = g2n / st.w;
#pragma unroll
                for (int e = 0; e < VEC; ++e) o[e] = fmaf(sv[j][e] - st.z, k, g1n);
                store_vec<VEC>(dsbase + (int64_t)idx * VEC, o, pol, hint);
            }
        }
    }
}

template <int VEC, int THREADS>
__device__ __forceinline__ float2 bwd_partial(const float (&dy)[kBatch][VEC], const float (&c)[kBatch][VEC], int b,
                                              int nvec, float4 st) {
    const float inv_sd_c = 1.f / st.y;
    float2 acc = make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < kBatch; ++j) {
        int idx = (b * kBatch + j) * THREADS + threadIdx.x;
        if (idx < nvec) {
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                acc.x += dy[j][e];
                acc.y = fmaf(dy[j][e], (c[j][e] - st.x) * inv_sd_c, acc.y);
            }
        }
    }
    return acc;
}

template <int VEC, int THREADS>
__global__ void __launch_bounds__(THREADS) adain_bwd_direct_kernel(BwdParams p) {
    __shared__ float2 scratch[32];
    const uint64_t pol = policy_evict_first();
    const bool hint = p.hints != 0;
    const int nvec = (int)(p.hw / VEC);
    const float hwf = (float)p.hw;
    for (int64_t plane = blockIdx.x; plane < p.planes; plane += gridDim.x) {
        const float4 st = __ldg(&p.saved[plane]);
        const float* dybase = p.dy + plane * p.hw;
        const float* cbase = p.content + plane * p.hw;
        float dy[kBatches][kBatch][VEC], c[kBatches][kBatch][VEC];
        float2 acc = make_float2(0.f, 0.f);
#pragma unroll
        for (int b = 0; b < kBatches; ++b) {
            load_batch<VEC, THREADS>(dy[b], dybase, b, nvec, pol, hint);
            load_batch<VEC, THREADS>(c[b], cbase, b, nvec, pol, hint);
        }
#pragma unroll
        for (int b = 0; b < kBatches; ++b) {
            float2 t = bwd_partial<VEC, THREADS>(dy[b], c[b], b, nvec, st);
            acc.x += t.x; acc.y += t.y;
        }
        const float2 g = block_sum2<THREADS>(acc, scratch);
#pragma unroll
        for (int b = 0; b < kBatches; ++b)
            bwd_write_batch<VEC, THREADS>(dy[b], c[b], p.style ? p.style + plane * p.hw : nullptr,
                                          p.dcontent + plane * p.hw, p.dstyle ? p.dstyle + plane * p.hw : nullptr,
                                          b, nvec, st, g, hwf, pol, hint);
        __syncthreads();
    }
}

template <int VEC>
__global__ void __launch_bounds__(kPipeThreads, 3) adain_bwd_pipe_kernel(BwdParams p) {
    constexpr int T = kPipeThreads;
    constexpr int CHUNK = T * kPerThread * VEC;
    __shared__ float2 scratch[32];
    __shared__ unsigned s_ticket;
    __shared__ int s_last;
    __shared__ float2 s_sums;
    const uint64_t pol_first = policy_evict_first();
    const uint64_t pol_last = policy_evict_last();
    const bool hint = p.hints != 0;
    const float hwf = (float)p.hw;

    AdainParams dec{};  // reuse the ticket decoder
    dec.planes = p.planes; dec.ipp = p.ipp; dec.lag = p.lag; dec.stats_only = 0;

    if (threadIdx.x == 0) s_ticket = atomicAdd(p.ticket, 1u);
    __syncthreads();
    unsigned t = s_ticket;
    while (t < p.total_items) {
        unsigned next = 0;
        if (threadIdx.x == 0) next = atomicAdd(p.ticket, 1u);
        const Item it = decode_ticket(t, dec);
        const int64_t e0 = (int64_t)it.chunk * CHUNK;
        const int64_t rem = p.hw - e0;
        const int nvec = (int)((rem < CHUNK ? rem : CHUNK) / VEC);
        const int64_t off = it.plane * p.hw + e0;
        const float4 st = __ldg(&p.saved[it.plane]);
        if (it.kind == 0) {
            float2 acc = make_float2(0.f, 0.f);
#pragma unroll
            for (int b = 0; b < kBatches; ++b) {
                float dy[kBatch][VEC], c[kBatch][VEC];
                load_batch<VEC, T>(dy, p.dy + off, b, nvec, pol_last, hint);
                load_batch<VEC, T>(c, p.content + off, b, nvec, pol_last, hint);
                float2 tt = bwd_partial<VEC, T>(dy, c, b, nvec, st);
                acc.x += tt.x; acc.y += tt.y;
            }
            const float2 g = block_sum2<T>(acc, scratch);
            if (threadIdx.x == 0) {
                __stcg(&p.part[it.plane * p.ipp + it.chunk], g);
                __threadfence();
                int old = atomicAdd(&p.done[it.plane], 1);
                s_last = (old == p.ipp - 1);
            }
            __syncthreads();
            if (s_last && threadIdx.x < 32) {
                __threadfence();
                float2 tot = make_float2(0.f, 0.f);
                for (int k = threadIdx.x; k < p.ipp; k += 32) {
                    float2 v = __ldcg(&p.part[it.plane * p.ipp + k]);
                    tot.x += v.x; tot.y += v.y;
                }
                tot.x = warp_sum(tot.x);
                tot.y = warp_sum(tot.y);
                if (threadIdx.x == 0) {
                    __stcg(&p.sums[it.plane], tot);
                    __threadfence();
                    st_release(&p.ready[it.plane], 1);
                }
            }
        } else {
            float dy[kBatch][VEC], c[kBatch][VEC];
            load_batch<VEC, T>(dy, p.dy + off, 0, nvec, pol_first, hint);
            load_batch<VEC, T>(c, p.content + off, 0, nvec, pol_first, hint);
            if (threadIdx.x == 0) {
                while (ld_acquire(&p.ready[it.plane]) == 0) __nanosleep(100);
                s_sums = __ldcg(&p.sums[it.plane]);
            }
            __syncthreads();
            const float2 g = s_sums;
#pragma unroll
            for (int b = 0; b < kBatches; ++b) {
                if (b > 0) {
                    load_batch<VEC, T>(dy, p.dy + off, b, nvec, pol_first, hint);
                    load_batch<VEC, T>(c, p.content + off, b, nvec, pol_first, hint);
                }
                bwd_write_batch<VEC, T>(dy, c, p.style ? p.style + off : nullptr, p.dcontent + off,
                                        p.dstyle ? p.dstyle + off : nullptr, b, nvec, st, g, hwf, pol_first, hint);
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) s_ticket = next;
        __syncthreads();
        t = s_ticket;
    }
}

// ------------------------------------------------------------------------------------------
// plane affine: out = x*scale[p] + shift[p]   (SELayer gating, network/attention.py:22)
// ------------------------------------------------------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(256) plane_affine_kernel(const float* __restrict__ x, const float* __restrict__ scale,
                                                           const float* __restrict__ shift, float* __restrict__ out,
                                                           int64_t planes, int64_t hw, int chunks_per_plane) {
    constexpr int CHUNK = 256 * kPerThread * VEC;
    const uint64_t pol = policy_evict_first();
    const int64_t items = planes * chunks_per_plane;
    for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
        int64_t plane = item / chunks_per_plane;
        int64_t e0 = (item % chunks_per_plane) * CHUNK;
        int64_t rem = hw - e0;
        int nvec = (int)((rem < CHUNK ? rem : CHUNK) / VEC);
        const float a = __ldg(scale + plane);
        const float b = shift ? __ldg(shift + plane) : 0.f;
        const float* xb = x + plane * hw + e0;
        float* ob = out + plane * hw + e0;
#pragma unroll
        for (int bt = 0; bt < kBatches; ++bt) {
            float v[kBatch][VEC];
            load_batch<VEC, 256>(v, xb, bt, nvec, pol, true);
#pragma unroll
            for (int j = 0; j < kBatch; ++j) {
                int idx = (bt * kBatch + j) * 256 + threadIdx.x;
                if (idx < nvec) {
                    float o[VEC];
#pragma unroll
                    for (int e = 0; e < VEC; ++e) o[e] = fmaf(v[j][e], a, b);
                    store_vec<VEC>(ob + (int64_t)idx * VEC, o, pol, true);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

struct PipeLayout {
    size_t counters_bytes;  // ticket + done + ready (zeroed before launch)
    size_t coef_off, part_c_off, part_s_off, total;
};

PipeLayout pipe_layout(int64_t planes, int64_t hw) {
    // sized for the scalar (VEC=1) chunking, the smaller chunk => the larger item count
    const int64_t chunk = (int64_t)kPipeThreads * kPerThread;
    const int64_t ipp = (hw + chunk - 1) / chunk;
    PipeLayout l;
    l.counters_bytes = align_up(256 + (size_t)planes * 2 * sizeof(int), 256);
    l.coef_off = l.counters_bytes;
    l.part_c_off = align_up(l.coef_off + (size_t)planes * sizeof(float4), 256);
    l.part_s_off = align_up(l.part_c_off + (size_t)planes * ipp * sizeof(float2), 256);
    l.total = align_up(l.part_s_off + (size_t)planes * ipp * sizeof(float2), 256);
    return l;
}

template <int VEC>
int launch_direct(const AdainParams& p, cudaStream_t stream) {
    const int64_t nvec = p.hw / VEC;
    const int64_t grid64 = p.planes < (int64_t)sm_count() * 32 ? p.planes : (int64_t)sm_count() * 32;
    const int grid = (int)grid64;
    if (nvec <= 128 * kPerThread) adain_direct_kernel<VEC, 128><<<grid, 128, 0, stream>>>(p);
    else if (nvec <= 256 * kPerThread) adain_direct_kernel<VEC, 256><<<grid, 256, 0, stream>>>(p);
    else adain_direct_kernel<VEC, 512><<<grid, 512, 0, stream>>>(p);
    RPST_CUDA(cudaGetLastError());
    return RPST_OK;
}

template <int VEC>
int launch_pipe(AdainParams p, void* ws, size_t ws_bytes, cudaStream_t stream) {
    constexpr int64_t CHUNK = (int64_t)kPipeThreads * kPerThread * VEC;
    const PipeLayout l = pipe_layout(p.planes, p.hw);
    if (ws == nullptr || ws_bytes < l.total) {
        set_error("adain: workspace too small (%zu < %zu bytes)", ws_bytes, l.total);
        return RPST_ERR_WORKSPACE;
    }
    RPST_CHECK_ARG(aligned16(ws), "adain: workspace must be 16-byte aligned");
    char* base = static_cast<char*>(ws);
    p.ipp = (int)((p.hw + CHUNK - 1) / CHUNK);
    const int64_t plane_bytes = p.hw * (int64_t)sizeof(float);
    int64_t lag = (g_tuning.lag_bytes + plane_bytes - 1) / plane_bytes;
    if (lag < 3) lag = 3;
    p.lag = (int)(lag < p.planes ? lag : p.planes);
    const int64_t total = p.planes * p.ipp * (p.stats_only ? 1 : 2);
    RPST_CHECK_ARG(total < (1ll << 31), "adain: too many work items (%lld); split the call", (long long)total);
    p.total_items = (unsigned)total;
    p.ticket = reinterpret_cast<unsigned*>(base);
    p.done = reinterpret_cast<int*>(base + 256);
    p.ready = p.done + p.planes;
    p.coef = reinterpret_cast<float4*>(base + l.coef_off);
    p.part_c = reinterpret_cast<float2*>(base + l.part_c_off);
    p.part_s = reinterpret_cast<float2*>(base + l.part_s_off);
    RPST_CUDA(cudaMemsetAsync(base, 0, l.counters_bytes, stream));
    int64_t grid = (int64_t)sm_count() * g_tuning.ctas_per_sm;
    if (grid > total) grid = total;
    if (g_tuning.ctas_per_sm >= 4) adain_pipe_kernel<VEC, 4><<<(int)grid, kPipeThreads, 0, stream>>>(p);
    else adain_pipe_kernel<VEC, 3><<<(int)grid, kPipeThreads, 0, stream>>>(p);
    RPST_CUDA(cudaGetLastError());
    return RPST_OK;
}

struct BwdLayout {
    size_t counters_bytes, sums_off, part_off, total;
};
BwdLayout bwd_layout(int64_t planes, int64_t hw) {
    const int64_t chunk = (int64_t)kPipeThreads * kPerThread;
    const int64_t ipp = (hw + chunk - 1) / chunk;
    BwdLayout l;
    l.counters_bytes = align_up(256 + (size_t)planes * 2 * sizeof(int), 256);
    l.sums_off = l.counters_bytes;
    l.part_off = align_up(l.sums_off + (size_t)planes * sizeof(float2), 256);
    l.total = align_up(l.part_off + (size_t)planes * ipp * sizeof(float2), 256);
    return l;
}

template <int VEC>
int launch_bwd(BwdParams p, void* ws, size_t ws_bytes, cudaStream_t stream) {
    const int64_t nvec = p.hw / VEC;
    if (nvec <= 512 * kPerThread) {
        const int64_t g64 = p.planes < (int64_t)sm_count() * 32 ? p.planes : (int64_t)sm_count() * 32;
        const int grid = (int)g64;
        if (nvec <= 128 * kPerThread) adain_bwd_direct_kernel<VEC, 128><<<grid, 128, 0, stream>>>(p);
        else if (nvec <= 256 * kPerThread) adain_bwd_direct_kernel<VEC, 256><<<grid, 256, 0, stream>>>(p);
        else adain_bwd_direct_kernel<VEC, 512><<<grid, 512, 0, stream>>>(p);
        RPST_CUDA(cudaGetLastError());
        return RPST_OK;
    }
    constexpr int64_t CHUNK = (int64_t)kPipeThreads * kPerThread * VEC;
    const BwdLayout l = bwd_layout(p.planes, p.hw);
    if (ws == nullptr || ws_bytes < l.total) {
        set_error("adain_bwd: workspace too small (%zu < %zu bytes)", ws_bytes, l.total);
        return RPST_ERR_WORKSPACE;
    }
    char* base = static_cast<char*>(ws);
    p.ipp = (int)((p.hw + CHUNK - 1) / CHUNK);
    const int64_t plane_bytes = 2 * p.hw * (int64_t)sizeof(float);  // dy and content stay resident
    int64_t lag = (g_tuning.lag_bytes + plane_bytes - 1) / plane_bytes;
    if (lag < 3) lag = 3;
    p.lag = (int)(lag < p.planes ? lag : p.planes);
    const int64_t total = p.planes * p.ipp * 2;
    RPST_CHECK_ARG(total < (1ll << 31), "adain_bwd: too many work items (%lld); split the call", (long long)total);
    p.total_items = (unsigned)total;
    p.ticket = reinterpret_cast<unsigned*>(base);
    p.done = reinterpret_cast<int*>(base + 256);
    p.ready = p.done + p.planes;
    p.sums = reinterpret_cast<float2*>(base + l.sums_off);
    p.part = reinterpret_cast<float2*>(base + l.part_off);
    RPST_CUDA(cudaMemsetAsync(base, 0, l.counters_bytes, stream));
    int64_t grid = (int64_t)sm_count() * 3;
    if (grid > total) grid = total;
    adain_bwd_pipe_kernel<VEC><<<(int)grid, kPipeThreads, 0, stream>>>(p);
    RPST_CUDA(cudaGetLastError());
    return RPST_OK;
}

int run_adain(AdainParams p, void* ws, size_t ws_bytes, cudaStream_t stream) {
    if (p.planes == 0 || p.hw == 0) return RPST_OK;
    p.hints = (int)g_tuning.hints;
    bool vec = (p.hw % 4 == 0) && aligned16(p.content) && (!p.style || aligned16(p.style)) &&
               (!p.prev || aligned16(p.prev)) && (!p.out || (aligned16(p.out) && p.out_batch_stride % 4 == 0));
    if (vec) {
        if (p.hw / 4 <= 512 * kPerThread) return launch_direct<4>(p, stream);
        return launch_pipe<4>(p, ws, ws_bytes, stream);
    }
    if (p.hw <= 512 * kPerThread) return launch_direct<1>(p, stream);
    return launch_pipe<1>(p, ws, ws_bytes, stream);
}

}  // namespace

int set_adain_tuning(const char* name, int64_t v, bool set, int64_t* out) {
    int64_t* slot = nullptr;
    if (!strcmp(name, "adain_lag_bytes")) slot = &g_tuning.lag_bytes;
    else if (!strcmp(name, "adain_hints")) slot = &g_tuning.hints;
    else if (!strcmp(name, "adain_ctas_per_sm")) slot = &g_tuning.ctas_per_sm;
    if (!slot) return 0;
    if (set) *slot = v;
    if (out) *out = *slot;
    return 1;
}

}  // namespace rpst

using namespace rpst;

extern "C" size_t rpst_stats_workspace_bytes(int64_t planes, int64_t hw) {
    if (planes <= 0 || hw <= 0) return 256;
    return pipe_layout(planes, hw).total;
}

extern "C" int rpst_stats_nchw(const float* x, int64_t planes, int64_t hw, float eps, float* mean, float* std,
                               void* workspace, size_t workspace_bytes, void* stream) {
    RPST_CHECK_ARG(planes >= 0 && hw >= 0, "stats: negative size");
    RPST_CHECK_ARG(planes == 0 || hw == 0 || x != nullptr, "stats: null input");
    AdainParams p{};
    p.content = x;
    p.mean_out = mean;
    p.std_out = std;
    p.planes = planes;
    p.hw = hw;
    p.channels = 1;
    p.eps = eps;
    p.stats_only = 1;
    return run_adain(p, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

extern "C" size_t rpst_adain_workspace_bytes(int64_t n, int64_t c, int64_t hw) {
    if (n <= 0 || c <= 0 || hw <= 0) return 256;
    return pipe_layout(n * c, hw).total;
}

extern "C" int rpst_adain_fwd(const float* content, const float* style, const float* prev, float* out,
                              int64_t n, int64_t c, int64_t hw, int64_t out_batch_stride, float eps,
                              float* saved_stats, void* workspace, size_t workspace_bytes, void* stream) {
    RPST_CHECK_ARG(n >= 0 && c >= 0 && hw >= 0, "adain: negative size");
    if (n == 0 || c == 0 || hw == 0) return RPST_OK;
    RPST_CHECK_ARG(content != nullptr && out != nullptr, "adain: null content/out");
    RPST_CHECK_ARG(out_batch_stride >= c * hw, "adain: out_batch_stride (%lld) < c*hw (%lld)",
                   (long long)out_batch_stride, (long long)(c * hw));
    RPST_CHECK_ARG(out != content && out != style && out != prev, "adain: out must not alias an input");
    AdainParams p{};
    p.content = content;
    p.style = style;
    p.prev = prev;
    p.out = out;
    p.saved = saved_stats;
    p.planes = n * c;
    p.hw = hw;
    p.channels = c;
    p.out_batch_stride = out_batch_stride;
    p.eps = eps;
    return run_adain(p, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

extern "C" size_t rpst_adain_bwd_workspace_bytes(int64_t n, int64_t c, int64_t hw) {
    if (n <= 0 || c <= 0 || hw <= 0) return 256;
    return bwd_layout(n * c, hw).total;
}

extern "C" int rpst_adain_bwd(const float* grad_out, const float* content, const float* style,
                              const float* saved_stats, float* grad_content, float* grad_style,
                              int64_t n, int64_t c, int64_t hw, void* workspace, size_t workspace_bytes,
                              void* stream) {
    RPST_CHECK_ARG(n >= 0 && c >= 0 && hw >= 0, "adain_bwd: negative size");
    if (n == 0 || c == 0 || hw == 0) return RPST_OK;
    RPST_CHECK_ARG(grad_out && content && saved_stats && grad_content, "adain_bwd: null pointer");
    RPST_CHECK_ARG(grad_style == nullptr || style != nullptr, "adain_bwd: grad_style requested without style");
    RPST_CHECK_ARG(aligned16(saved_stats), "adain_bwd: saved_stats must be 16-byte aligned");
    BwdParams p{};
    p.dy = grad_out;
    p.content = content;
    p.style = grad_style ? style : nullptr;
    p.saved = reinterpret_cast<const float4*>(saved_stats);
    p.dcontent = grad_content;
    p.dstyle = grad_style;
    p.planes = n * c;
    p.hw = hw;
    p.hints = (int)g_tuning.hints;
    const bool vec = hw % 4 == 0 && aligned16(grad_out) && aligned16(content) && aligned16(grad_content) &&
                     (!grad_style || (aligned16(style) && aligned16(grad_style)));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return vec ? launch_bwd<4>(p, workspace, workspace_bytes, st) : launch_bwd<1>(p, workspace, workspace_bytes, st);
}

extern "C" int rpst_plane_affine(const float* x, const float* scale, const float* shift, float* out,
                                 int64_t planes, int64_t hw, void* stream) {
    RPST_CHECK_ARG(planes >= 0 && hw >= 0, "plane_affine: negative size");
    if (planes == 0 || hw == 0) return RPST_OK;
    RPST_CHECK_ARG(x && scale && out, "plane_affine: null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool vec = hw % 4 == 0 && aligned16(x) && aligned16(out);
    const int64_t chunk = 256ll * kPerThread * (vec ? 4 : 1);
    const int cpp = (int)((hw + chunk - 1) / chunk);
    int64_t items = planes * cpp;
    int64_t grid = (int64_t)sm_count() * 8;
    if (grid > items) grid = items;
    if (vec) plane_affine_kernel<4><<<(int)grid, 256, 0, st>>>(x, scale, shift, out, planes, hw, cpp);
    else plane_affine_kernel<1><<<(int)grid, 256, 0, st>>>(x, scale, shift, out, planes, hw, cpp);
    RPST_CUDA(cudaGetLastError());
    return RPST_OK;
}
