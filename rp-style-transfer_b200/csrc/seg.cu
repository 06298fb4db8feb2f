// Segment (mask-label) AdaIN — SURVEY.md §8 a4; reference: network/base.py:421-530 and the batch
// loop network/adain_rp.py:313-319.
//
// The reference walks the label set on the host (np.where per label, index_select gather, ~12
// launches, index_copy_ scatter; per label, per level, per sample).  Here the whole batch is two
// launches:
//   seg_hist_kernel  — per sample: pixel count and first pixel index of every label value (0..255)
//                      in the content and in the style map (labels are shared by all channels).
//   seg_pipe_kernel  — the ticket-pipelined structure of adain.cu: statistics items (content and
//                      style chunks of plane p+D) run ahead of apply items (plane p), the content
//                      is re-read from L2.  Statistics are per (plane,label) shifted sums
//                      S1 = sum(x-K_l), S2 = sum((x-K_l)^2) with K_l = the first pixel of label l in
//                      that plane, accumulated run-length-compressed in registers, warp-aggregated,
//                      reduced in shared memory and finally with one fp32 atomic per (item,label).
// The validity rule (network/base.py:435: label taken from the CONTENT map; cnt_c>10, cnt_s>10,
// cnt_c/cnt_s<100, cnt_s/cnt_c<100) is evaluated on the device from the histogram; pixels of
// unusable labels are copied through bit-exactly.
#include "common.cuh"
#include "plane_io.cuh"

namespace rpst {
namespace {

constexpr int kLabels = 256;
constexpr int kNoIndex = 0x7f7f7f7f;  // memset(0x7f) pattern = "no pixel seen"

struct SegParams {
    const float* content;
    const float* style;
    const uint8_t* c_lab;   // [n, hw_c]
    const uint8_t* s_lab;   // [n, hw_s]
    const float* prev;      // may be null
    float* out;
    int64_t n, channels, hw_c, hw_s;
    float eps;
    int hints;
    int ipp_c, ipp_s, lag;
    unsigned total_items;
    // workspace
    unsigned* ticket;
    int* done;              // [planes]
    int* ready;             // [planes]
    int* cnt;               // [n][2][256]
    int* first;             // [n][2][256]
    float2* gsum;           // [planes][2][256] (S1, S2)
    float4* coef;           // [planes][256] (mu_c, a, mu_s, valid)
};

__global__ void __launch_bounds__(256) seg_hist_kernel(const uint8_t* __restrict__ c_lab, const uint8_t* __restrict__ s_lab,
                                                       int64_t hw_c, int64_t hw_s, int* __restrict__ cnt,
                                                       int* __restrict__ first) {
    __shared__ int h[kLabels];
    __shared__ int f[kLabels];
    const int which = blockIdx.z;  // 0 content, 1 style
    const int64_t n = blockIdx.y;
    const int64_t hw = which ? hw_s : hw_c;
    const uint8_t* lab = (which ? s_lab : c_lab) + n * hw;
    h[threadIdx.x] = 0;
    f[threadIdx.x] = kNoIndex;
    __syncthreads();
    const int64_t per_block = (hw + gridDim.x - 1) / gridDim.x;
    const int64_t beg = blockIdx.x * per_block;
    const int64_t end = beg + per_block < hw ? beg + per_block : hw;
    for (int64_t i = beg + threadIdx.x; i < end; i += blockDim.x) {
        int l = lab[i];
        atomicAdd(&h[l], 1);
        atomicMin(&f[l], (int)i);
    }
    __syncthreads();
    const int64_t slot = (n * 2 + which) * kLabels + threadIdx.x;
    if (h[threadIdx.x]) {
        atomicAdd(&cnt[slot], h[threadIdx.x]);
        atomicMin(&first[slot], f[threadIdx.x]);
    }
}

__device__ __forceinline__ bool label_usable(int nc, int ns) {
    // network/base.py:435 with the float ratios written as exact integer comparisons
    return nc > 10 && ns > 10 && (int64_t)nc < 100ll * ns && (int64_t)ns < 100ll * nc;
}

struct SegItem {
    int kind;  // 0 content statistics, 1 style statistics, 2 apply
    int64_t plane;
    int chunk;
};

__device__ __forceinline__ SegItem seg_decode(unsigned t, const SegParams& p) {
    const int64_t planes = p.n * p.channels;
    const int64_t st = p.ipp_c + p.ipp_s;  // statistics items per plane
    const int64_t lag = p.lag < planes ? p.lag : planes;
    SegItem it;
    int64_t tt = t;
    auto stats = [&](int64_t plane, int64_t u) {
        it.plane = plane;
        if (u < p.ipp_c) { it.kind = 0; it.chunk = (int)u; }
        else { it.kind = 1; it.chunk = (int)(u - p.ipp_c); }
    };
    if (tt < lag * st) { stats(tt / st, tt % st); return it; }
    tt -= lag * st;
    const int64_t round = st + p.ipp_c;
    const int64_t steady = (planes - lag) * round;
    if (tt < steady) {
        int64_t r = tt / round, u = tt % round;
        if (u < st) stats(r + lag, u);
        else { it.kind = 2; it.plane = r; it.chunk = (int)(u - st); }
        return it;
    }
    tt -= steady;
    it.kind = 2; it.plane = (planes - lag) + tt / p.ipp_c; it.chunk = (int)(tt % p.ipp_c);
    return it;
}

template <int VEC>
__device__ __forceinline__ void load_labels(int (&lab)[VEC], const uint8_t* p) {
    if constexpr (VEC == 4) {
        uchar4 v = __ldg(reinterpret_cast<const uchar4*>(p));
        lab[0] = v.x; lab[1] = v.y; lab[2] = v.z; lab[3] = v.w;
    } else {
        lab[0] = __ldg(p);
    }
}

template <int VEC>
__global__ void __launch_bounds__(kPipeThreads, 3) seg_pipe_kernel(SegParams p) {
    constexpr int T = kPipeThreads;
    constexpr int CHUNK = T * kPerThread * VEC;
    static_assert(T == kLabels, "one thread per label value");
    __shared__ float s_shift[kLabels];
    __shared__ float s_acc[kLabels][2];
    __shared__ float4 s_coef[kLabels];
    __shared__ unsigned char s_use[kLabels];
    __shared__ unsigned s_ticket;
    __shared__ int s_last;

    const uint64_t pol_first = policy_evict_first();
    const uint64_t pol_last = policy_evict_last();
    const bool hint = p.hints != 0;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) s_ticket = atomicAdd(p.ticket, 1u);
    __syncthreads();
    unsigned t = s_ticket;

    while (t < p.total_items) {
        unsigned next = 0;
        if (threadIdx.x == 0) next = atomicAdd(p.ticket, 1u);
        const SegItem it = seg_decode(t, p);
        const int64_t sample = it.plane / p.channels;
        const int* cnt_c = p.cnt + (sample * 2 + 0) * kLabels;
        const int* cnt_s = p.cnt + (sample * 2 + 1) * kLabels;

        if (it.kind < 2) {
            // ------------------------------------------------ statistics item (content or style)
            const bool is_style = it.kind == 1;
            const int64_t hw = is_style ? p.hw_s : p.hw_c;
            const float* x = (is_style ? p.style : p.content) + it.plane * hw;
            const uint8_t* labels = (is_style ? p.s_lab : p.c_lab) + sample * hw;
            const int* first = p.first + (sample * 2 + it.kind) * kLabels;
            const int64_t e0 = (int64_t)it.chunk * CHUNK;
            const int64_t rem = hw - e0;
            const int nvec = (int)((rem < CHUNK ? rem : CHUNK) / VEC);
            {
                const int l = threadIdx.x;
                const bool use = label_usable(__ldg(cnt_c + l), __ldg(cnt_s + l));
                s_use[l] = use;
                s_shift[l] = use ? __ldg(x + __ldg(first + l)) : 0.f;
                s_acc[l][0] = 0.f;
                s_acc[l][1] = 0.f;
            }
            __syncthreads();
            int cur = -1;
            float shift = 0.f, a1 = 0.f, a2 = 0.f;
            const uint64_t xpol = is_style ? pol_first : pol_last;
#pragma unroll
            for (int b = 0; b < kBatches; ++b) {
                float v[kBatch][VEC];
                load_batch<VEC, T>(v, x + e0, b, nvec, xpol, hint);
#pragma unroll
                for (int j = 0; j < kBatch; ++j) {
                    const int idx = (b * kBatch + j) * T + threadIdx.x;
                    if (idx < nvec) {
                        int lab[VEC];
                        load_labels<VEC>(lab, labels + e0 + (int64_t)idx * VEC);
#pragma unroll
                        for (int e = 0; e < VEC; ++e) {
                            const int l = lab[e];
                            if (l != cur) {
                                if (cur >= 0) {  // rare mid-item flush (label run ended)
                                    atomicAdd(&s_acc[cur][0], a1);
                                    atomicAdd(&s_acc[cur][1], a2);
                                }
                                if (s_use[l]) { cur = l; shift = s_shift[l]; }
                                else cur = -1;
                                a1 = 0.f; a2 = 0.f;
                            }
                            if (cur >= 0) {
                                const float d = v[j][e] - shift;
                                a1 += d;
                                a2 = fmaf(d, d, a2);
                            }
                        }
                    }
                }
            }
            // final flush, aggregated per distinct label inside the warp
            unsigned remaining = __ballot_sync(0xffffffffu, cur >= 0);
            while (remaining) {
                const int leader = __ffs(remaining) - 1;
                const int l = __shfl_sync(0xffffffffu, cur, leader);
                const bool mine = cur == l;
                const float r1 = warp_sum(mine ? a1 : 0.f);
                const float r2 = warp_sum(mine ? a2 : 0.f);
                if (lane == leader) {
                    atomicAdd(&s_acc[l][0], r1);
                    atomicAdd(&s_acc[l][1], r2);
                }
                remaining &= ~__ballot_sync(0xffffffffu, mine);
            }
            __syncthreads();
            {
                const int l = threadIdx.x;
                const float r1 = s_acc[l][0], r2 = s_acc[l][1];
                if (r1 != 0.f || r2 != 0.f) {
                    float2* g = p.gsum + (it.plane * 2 + it.kind) * kLabels + l;
                    atomicAdd(&g->x, r1);
                    atomicAdd(&g->y, r2);
                }
            }
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0) {
                int old = atomicAdd(&p.done[it.plane], 1);
                s_last = (old == p.ipp_c + p.ipp_s - 1);
            }
            __syncthreads();
            if (s_last) {
                // last statistics item of the plane: turn the sums into per-label AdaIN coefficients
                __threadfence();
                const int l = threadIdx.x;
                const int nc = __ldg(cnt_c + l), ns = __ldg(cnt_s + l);
                float4 cf = make_float4(0.f, 1.f, 0.f, 0.f);  // identity for unusable labels
                if (label_usable(nc, ns)) {
                    const float kc = __ldg(p.content + it.plane * p.hw_c + __ldg(p.first + (sample * 2 + 0) * kLabels + l));
                    const float ks = __ldg(p.style + it.plane * p.hw_s + __ldg(p.first + (sample * 2 + 1) * kLabels + l));
                    const float2 gc = __ldcg(p.gsum + (it.plane * 2 + 0) * kLabels + l);
                    const float2 gs = __ldcg(p.gsum + (it.plane * 2 + 1) * kLabels + l);
                    const float fnc = (float)nc, fns = (float)ns;
                    const float mu_c = kc + gc.x / fnc;
                    const float mu_s = ks + gs.x / fns;
                    const float m2c = fmaxf(gc.y - gc.x * gc.x / fnc, 0.f);
                    const float m2s = fmaxf(gs.y - gs.x * gs.x / fns, 0.f);
                    const float sd_c = sqrtf(m2c / (fnc - 1.f) + p.eps);
                    const float sd_s = sqrtf(m2s / (fns - 1.f) + p.eps);
                    cf = make_float4(mu_c, sd_s / sd_c, mu_s, 1.f);
                }
                __stcg(&p.coef[it.plane * kLabels + l], cf);
                __threadfence();
                __syncthreads();
                if (threadIdx.x == 0) st_release(&p.ready[it.plane], 1);
            }
        } else {
            // ------------------------------------------------ apply item
            const int64_t e0 = (int64_t)it.chunk * CHUNK;
            const int64_t rem = p.hw_c - e0;
            const int nvec = (int)((rem < CHUNK ? rem : CHUNK) / VEC);
            const float* cbase = p.content + it.plane * p.hw_c + e0;
            const float* pbase = p.prev ? p.prev + it.plane * p.hw_c + e0 : nullptr;
            float* obase = p.out + it.plane * p.hw_c + e0;
            const uint8_t* labels = p.c_lab + sample * p.hw_c + e0;
            float c[kBatch][VEC], pv[kBatch][VEC];
            load_batch<VEC, T>(c, cbase, 0, nvec, pol_first, hint);
            if (pbase) load_batch<VEC, T>(pv, pbase, 0, nvec, pol_first, hint);
            if (threadIdx.x == 0) {
                while (ld_acquire(&p.ready[it.plane]) == 0) __nanosleep(100);
            }
            __syncthreads();
            s_coef[threadIdx.x] = __ldcg(&p.coef[it.plane * kLabels + threadIdx.x]);
            __syncthreads();
#pragma unroll
            for (int b = 0; b < kBatches; ++b) {
                if (b > 0) {
                    load_batch<VEC, T>(c, cbase, b, nvec, pol_first, hint);
                    if (pbase) load_batch<VEC, T>(pv, pbase, b, nvec, pol_first, hint);
                }
#pragma unroll
                for (int j = 0; j < kBatch; ++j) {
                    const int idx = (b * kBatch + j) * T + threadIdx.x;
                    if (idx < nvec) {
                        int lab[VEC];
                        load_labels<VEC>(lab, labels + (int64_t)idx * VEC);
                        float o[VEC];
#pragma unroll
                        for (int e = 0; e < VEC; ++e) {
                            const float4 cf = s_coef[lab[e]];
                            // unusable label: (c-0)*1+0 == c bit-exactly (pass-through)
                            const float y = fmaf(c[j][e] - cf.x, cf.y, cf.z);
                            o[e] = pbase ? y + pv[j][e] : y;
                        }
                        store_vec<VEC>(obase + (int64_t)idx * VEC, o, pol_first, hint);
                    }
                }
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) s_ticket = next;
        __syncthreads();
        t = s_ticket;
    }
}

__global__ void seg_export_info_kernel(const int* __restrict__ cnt, int32_t* __restrict__ info, int64_t n) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n * kLabels) return;
    const int64_t s = i / kLabels, l = i % kLabels;
    const int nc = cnt[(s * 2 + 0) * kLabels + l], ns = cnt[(s * 2 + 1) * kLabels + l];
    info[i * 3 + 0] = nc;
    info[i * 3 + 1] = ns;
    info[i * 3 + 2] = label_usable(nc, ns) ? 1 : 0;
}

struct SegLayout {
    size_t zero_bytes;   // ticket, done, ready, cnt, gsum (memset 0)
    size_t first_off, first_bytes, coef_off, total;
    size_t done_off, cnt_off, gsum_off;
};

SegLayout seg_layout(int64_t n, int64_t c) {
    const size_t planes = (size_t)(n * c);
    SegLayout l;
    l.done_off = 256;
    l.cnt_off = align_up(l.done_off + planes * 2 * sizeof(int), 256);
    l.gsum_off = align_up(l.cnt_off + (size_t)n * 2 * kLabels * sizeof(int), 256);
    l.zero_bytes = align_up(l.gsum_off + planes * 2 * kLabels * sizeof(float2), 256);
    l.first_off = l.zero_bytes;
    l.first_bytes = align_up((size_t)n * 2 * kLabels * sizeof(int), 256);
    l.coef_off = l.first_off + l.first_bytes;
    l.total = align_up(l.coef_off + planes * kLabels * sizeof(float4), 256);
    return l;
}

inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }

}  // namespace

int64_t adain_tuning_value(const char* name);

}  // namespace rpst

using namespace rpst;

extern "C" size_t rpst_seg_adain_workspace_bytes(int64_t n, int64_t c, int64_t hw_c, int64_t hw_s) {
    (void)hw_c; (void)hw_s;
    if (n <= 0 || c <= 0) return 256;
    return seg_layout(n, c).total;
}

extern "C" int rpst_seg_adain_fwd(const float* content, const float* style, const uint8_t* c_labels,
                                  const uint8_t* s_labels, const float* prev, float* out, int64_t n, int64_t c,
                                  int64_t hw_c, int64_t hw_s, float eps, int32_t* label_info, void* workspace,
                                  size_t workspace_bytes, void* stream) {
    RPST_CHECK_ARG(n >= 0 && c >= 0 && hw_c >= 0 && hw_s >= 0, "seg_adain: negative size");
    if (n == 0 || c == 0 || hw_c == 0) return RPST_OK;
    RPST_CHECK_ARG(content && style && c_labels && s_labels && out, "seg_adain: null pointer");
    RPST_CHECK_ARG(hw_s > 0, "seg_adain: empty style map");
    RPST_CHECK_ARG(hw_c < (1ll << 31) && hw_s < (1ll << 31), "seg_adain: plane too large");
    RPST_CHECK_ARG(out != content && out != style && out != prev, "seg_adain: out must not alias an input");
    const SegLayout l = seg_layout(n, c);
    if (workspace == nullptr || workspace_bytes < l.total) {
        set_error("seg_adain: workspace too small (%zu < %zu bytes)", workspace_bytes, l.total);
        return RPST_ERR_WORKSPACE;
    }
    RPST_CHECK_ARG(aligned(workspace, 16), "seg_adain: workspace must be 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    char* base = static_cast<char*>(workspace);
    SegParams p{};
    p.content = content; p.style = style; p.c_lab = c_labels; p.s_lab = s_labels; p.prev = prev; p.out = out;
    p.n = n; p.channels = c; p.hw_c = hw_c; p.hw_s = hw_s; p.eps = eps;
    p.hints = (int)adain_tuning_value("adain_hints");
    p.ticket = reinterpret_cast<unsigned*>(base);
    p.done = reinterpret_cast<int*>(base + l.done_off);
    p.ready = p.done + n * c;
    p.cnt = reinterpret_cast<int*>(base + l.cnt_off);
    p.gsum = reinterpret_cast<float2*>(base + l.gsum_off);
    p.first = reinterpret_cast<int*>(base + l.first_off);
    p.coef = reinterpret_cast<float4*>(base + l.coef_off);
    RPST_CUDA(cudaMemsetAsync(base, 0, l.zero_bytes, st));
    RPST_CUDA(cudaMemsetAsync(base + l.first_off, 0x7f, l.first_bytes, st));

    const int64_t hw_max = hw_c > hw_s ? hw_c : hw_s;
    int hist_blocks = (int)((hw_max + 16383) / 16384);
    if (hist_blocks > 256) hist_blocks = 256;
    seg_hist_kernel<<<dim3(hist_blocks, (unsigned)n, 2), 256, 0, st>>>(c_labels, s_labels, hw_c, hw_s, p.cnt, p.first);
    RPST_CUDA(cudaGetLastError());
    if (label_info) {
        const int64_t tot = n * kLabels;
        seg_export_info_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(p.cnt, label_info, n);
        RPST_CUDA(cudaGetLastError());
    }

    const bool vec = hw_c % 4 == 0 && hw_s % 4 == 0 && aligned(content, 16) && aligned(style, 16) && aligned(out, 16) &&
                     (!prev || aligned(prev, 16)) && aligned(c_labels, 4) && aligned(s_labels, 4);
    const int64_t chunk = (int64_t)kPipeThreads * kPerThread * (vec ? 4 : 1);
    p.ipp_c = (int)((hw_c + chunk - 1) / chunk);
    p.ipp_s = (int)((hw_s + chunk - 1) / chunk);
    const int64_t plane_bytes = hw_c * (int64_t)sizeof(float);
    int64_t lag = (adain_tuning_value("adain_lag_bytes") + plane_bytes - 1) / plane_bytes;
    if (lag < 3) lag = 3;
    const int64_t planes = n * c;
    p.lag = (int)(lag < planes ? lag : planes);
    const int64_t total = planes * (2ll * p.ipp_c + p.ipp_s);
    RPST_CHECK_ARG(total < (1ll << 31), "seg_adain: too many work items (%lld); split the call", (long long)total);
    p.total_items = (unsigned)total;
    int64_t grid = (int64_t)sm_count() * 3;
    if (grid > total) grid = total;
    if (vec) seg_pipe_kernel<4><<<(int)grid, kPipeThreads, 0, st>>>(p);
    else seg_pipe_kernel<1><<<(int)grid, kPipeThreads, 0, st>>>(p);
    RPST_CUDA(cudaGetLastError());
    return RPST_OK;
}
