// fp64 pipe microbenchmark: mma.sync.m8n8k4.f64 (DMMA) vs DFMA peak.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/fp64_rate tools/micro/fp64_rate.cu
#include <cstdio>
__global__ void k(double* out, int iters) {
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
    double c[16][2];
    for (int i = 0; i < 16; ++i) { c[i][0] = 0; c[i][1] = 0; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0; for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void kf(double* out, int iters) {
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
    double c[32];
    for (int i = 0; i < 32; ++i) c[i] = i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 32; ++i) c[i] = fma(a, b, c[i]);
    }
    double s = 0; for (int i = 0; i < 32; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    double* d; cudaMalloc(&d, 148 * 4 * 256 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) {
        int iters = 20000;
        cudaEventRecord(e0); k<<<148 * 4, 256>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double fl = 148.0 * 4 * 8 * iters * 16 * 512.0;   // warps * mma * 2*8*8*4
        printf("dmma %.3f ms %.2f TFLOP/s\n", ms, fl / ms / 1e9);
        cudaEventRecord(e0); kf<<<148 * 4, 256>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        fl = 148.0 * 4 * 256 * (double)iters * 32 * 2;
        printf("dfma %.3f ms %.2f TFLOP/s\n", ms, fl / ms / 1e9);
    }
    return 0;
}
