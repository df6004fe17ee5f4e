// FP64 roofline probe for B200 (sm_100a): DFMA, DMMA (mma.sync m8n8k4 / m16n8k8 f64),
// cuBLAS DGEMM and a streaming-store bandwidth kernel. Prints one JSON object.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peak tools/fp64_peak.cu -lcublas
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <cublas_v2.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__global__ void dfma_kernel(double* out, int iters, double a, double b) {
    double x[16];
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int NACC>
__global__ void dmma_kernel(double* out, int iters, double a, double b) {
    double c0[NACC], c1[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) { c0[i] = threadIdx.x * 1e-9 + i; c1[i] = i; }
    double aa = a + threadIdx.x * 1e-12, bb = b;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) dmma884(c0[i], c1[i], aa, bb);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma1688(double* c, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__global__ void dmma1688_kernel(double* out, int iters, double a0, double b0) {
    double c[8][4];
#pragma unroll
    for (int i = 0; i < 8; i++) for (int j = 0; j < 4; j++) c[i][j] = threadIdx.x * 1e-9 + i + j;
    double a[4] = {a0, a0 + 1e-12, a0 + 2e-12, a0 + threadIdx.x * 1e-12}, b[2] = {b0, b0 * 0.5};
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) dmma1688(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) for (int j = 0; j < 4; j++) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void store_kernel(double2* out, size_t n2, double v) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n2; i += stride) out[i] = make_double2(v, v + 1.0);
}

template <class F> float time_ms(F f, int reps) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));
    printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, sms);
    // DFMA: blocks = sms*4, 256 thr
    for (int tpb : {128, 256, 512, 1024}) {
        int iters = 20000, blocks = sms * (2048 / tpb);
        float ms = time_ms([&] { dfma_kernel<<<blocks, tpb>>>(out, iters, 1.000001, 1e-9); }, 5);
        double fl = 2.0 * 16 * iters * (double)blocks * tpb;
        printf(", \"dfma_tflops_tpb%d\": %.3f", tpb, fl / ms * 1e-9);
    }
    for (int tpb : {128, 256, 512}) {
        int iters = 5000, blocks = sms * (1024 / tpb);
        float ms = time_ms([&] { dmma_kernel<16><<<blocks, tpb>>>(out, iters, 1.000001, 1e-9); }, 5);
        double fl = 2.0 * 256 * 16 * iters * (double)blocks * (tpb / 32);
        printf(", \"dmma884_acc16_tflops_tpb%d\": %.3f", tpb, fl / ms * 1e-9);
    }
    {
        int tpb = 256, iters = 5000, blocks = sms * 2;
        float ms = time_ms([&] { dmma_kernel<32><<<blocks, tpb>>>(out, iters, 1.000001, 1e-9); }, 5);
        double fl = 2.0 * 256 * 32 * iters * (double)blocks * (tpb / 32);
        printf(", \"dmma884_acc32_tflops_tpb256x2\": %.3f", fl / ms * 1e-9);
        blocks = sms;
        ms = time_ms([&] { dmma_kernel<32><<<blocks, tpb>>>(out, iters, 1.000001, 1e-9); }, 5);
        fl = 2.0 * 256 * 32 * iters * (double)blocks * (tpb / 32);
        printf(", \"dmma884_acc32_tflops_tpb256x1\": %.3f", fl / ms * 1e-9);
        blocks = sms; tpb = 128;
        ms = time_ms([&] { dmma_kernel<32><<<blocks, tpb>>>(out, iters, 1.000001, 1e-9); }, 5);
        fl = 2.0 * 256 * 32 * iters * (double)blocks * (tpb / 32);
        printf(", \"dmma884_acc32_tflops_tpb128x1\": %.3f", fl / ms * 1e-9);
    }
    {
        int tpb = 256, iters = 2000, blocks = sms * 2;
        float ms = time_ms([&] { dmma1688_kernel<<<blocks, tpb>>>(out, iters, 1.000001, 1e-9); }, 5);
        double fl = 2.0 * 16 * 8 * 8 * 8 * iters * (double)blocks * (tpb / 32);
        printf(", \"dmma1688_tflops\": %.3f", fl / ms * 1e-9);
    }
    // store bandwidth: 4 GiB
    {
        size_t bytes = (size_t)4 << 30; double2* buf; CK(cudaMalloc(&buf, bytes));
        float ms = time_ms([&] { store_kernel<<<sms * 8, 512>>>(buf, bytes / 16, 1.0); }, 5);
        printf(", \"store_gbs\": %.1f", bytes / ms * 1e-6);
        ms = time_ms([&] { CK(cudaMemsetAsync(buf, 0, bytes)); }, 5);
        printf(", \"memset_gbs\": %.1f", bytes / ms * 1e-6);
        CK(cudaFree(buf));
    }
    // cuBLAS DGEMM
    {
        cublasHandle_t h; cublasCreate(&h);
        for (int n : {4096, 8192, 16384}) {
            double *A, *B, *C; size_t sz = sizeof(double) * (size_t)n * n;
            CK(cudaMalloc(&A, sz)); CK(cudaMalloc(&B, sz)); CK(cudaMalloc(&C, sz));
            CK(cudaMemset(A, 0, sz)); CK(cudaMemset(B, 0, sz)); CK(cudaMemset(C, 0, sz));
            double al = 1.0, be = 0.0;
            float ms = time_ms([&] { cublasDgemm(h, CUBLAS_OP_T, CUBLAS_OP_N, n, n, n, &al, A, n, B, n, &be, C, n); }, 3);
            printf(", \"cublas_dgemm_tn_%d_tflops\": %.3f", n, 2.0 * n * (double)n * n / ms * 1e-9);
            if (n == 8192) {
                // sustained: 40 back-to-back
                cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
                cudaEventRecord(e0);
                for (int r = 0; r < 40; r++) cublasDgemm(h, CUBLAS_OP_T, CUBLAS_OP_N, n, n, n, &al, A, n, B, n, &be, C, n);
                cudaEventRecord(e1); cudaEventSynchronize(e1); float t; cudaEventElapsedTime(&t, e0, e1);
                printf(", \"cublas_dgemm_8192_sustained_tflops\": %.3f", 40 * 2.0 * n * (double)n * n / t * 1e-9);
                ms = time_ms([&] { cublasDsyrk(h, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, n, 512, &al, A, 512, &be, C, n); }, 3);
                printf(", \"cublas_dsyrk_n8192_k512_tflops\": %.3f", 1.0 * n * (double)n * 512 / ms * 1e-9);
            }
            cudaFree(A); cudaFree(B); cudaFree(C);
        }
        cublasDestroy(h);
    }
    printf("}\n");
    return 0;
}
