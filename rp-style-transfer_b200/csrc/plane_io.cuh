// Register-batch plane I/O shared by the AdaIN-family kernels (adain.cu, seg.cu).
#pragma once
#include "common.cuh"

namespace rpst {

constexpr int kBatch = 4;          // vectors per load batch
constexpr int kBatches = 2;        // batches per thread per item
constexpr int kPerThread = kBatch * kBatches;
constexpr int kPipeThreads = 256;

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
                                           uint64_t pol, bool hint, int tid = threadIdx.x) {
#pragma unroll
    for (int j = 0; j < kBatch; ++j) {
        int idx = (batch * kBatch + j) * THREADS + tid;
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
__device__ __forceinline__ void batch_moments(Moments& acc, const float (&v)[kBatch][VEC], int batch, int nvec,
                                              int tid = threadIdx.x) {
    float sum = 0.f;
    int cnt = 0;
#pragma unroll
    for (int j = 0; j < kBatch; ++j) {
        int idx = (batch * kBatch + j) * THREADS + tid;
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
        int idx = (batch * kBatch + j) * THREADS + tid;
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

}  // namespace rpst
