// Micro-benchmark: 4x4 fp64 mat-vec per thread with the matrix (a) broadcast from shared memory,
// (b) read from __constant__ memory with a warp-uniform index.  Also with a per-thread shared-memory
// round trip of the vector (as the fused kernel does for its newest partial).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o ldc_bench ldc_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#define NE 62
#define NC 4
__constant__ double cM[NC * NE * 16];
__constant__ int cprog[NE];

template <int MODE, int RT>
__global__ void __launch_bounds__(384, 1) k(const double *gM, const int *gprog, double *out, int iters)
{
    extern __shared__ double sm[];
    double *sM = sm;                       // NC*NE*16
    int *sprog = (int *)(sM + NC * NE * 16);
    double *cur = (double *)(sprog + 64);  // [NC][4][384]
    for (int i = threadIdx.x; i < NC * NE * 16; i += blockDim.x) sM[i] = gM[i];
    for (int i = threadIdx.x; i < NE; i += blockDim.x) sprog[i] = gprog[i];
    __syncthreads();
    double v[NC][4];
    for (int c = 0; c < NC; c++) for (int j = 0; j < 4; j++) v[c][j] = 1.0 + threadIdx.x * 1e-3 + j;
    if (RT) for (int c = 0; c < NC; c++) for (int j = 0; j < 4; j++) cur[(c * 4 + j) * 384 + threadIdx.x] = v[c][j];
    double acc = 0;
    for (int it = 0; it < iters; it++) {
#pragma unroll 1
        for (int i = 0; i < NE; i++) {
            const int e = (MODE == 0) ? sprog[i] : cprog[i];
#pragma unroll 2
            for (int c = 0; c < NC; c++) {
                double x[4];
                if (RT) { for (int j = 0; j < 4; j++) x[j] = cur[(c * 4 + j) * 384 + threadIdx.x]; }
                else { for (int j = 0; j < 4; j++) x[j] = v[c][j]; }
                double y[4];
                if (MODE == 0) {
                    const double *M = sM + (c * NE + e) * 16;
                    for (int r = 0; r < 4; r++) {
                        const double2 a = *(const double2 *)(M + r * 4), b = *(const double2 *)(M + r * 4 + 2);
                        y[r] = a.x * x[0] + a.y * x[1] + b.x * x[2] + b.y * x[3];
                    }
                } else {
                    const double *M = cM + (c * NE + e) * 16;
                    for (int r = 0; r < 4; r++) y[r] = M[r * 4] * x[0] + M[r * 4 + 1] * x[1] + M[r * 4 + 2] * x[2] + M[r * 4 + 3] * x[3];
                }
                if (RT) { for (int j = 0; j < 4; j++) cur[(c * 4 + j) * 384 + threadIdx.x] = y[j]; }
                else { for (int j = 0; j < 4; j++) v[c][j] = y[j]; }
            }
        }
    }
    if (RT) for (int c = 0; c < NC; c++) for (int j = 0; j < 4; j++) acc += cur[(c * 4 + j) * 384 + threadIdx.x];
    else for (int c = 0; c < NC; c++) for (int j = 0; j < 4; j++) acc += v[c][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE, int RT>
static void run(const char *name, const double *gM, const int *gprog, double *out)
{
    const int iters = 200, grid = 148, bd = 384;
    size_t smem = sizeof(double) * NC * NE * 16 + 64 * 4 + sizeof(double) * NC * 4 * 384;
    cudaFuncSetAttribute(k<MODE, RT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE, RT><<<grid, bd, smem>>>(gM, gprog, out, 2);
    cudaEventRecord(a);
    k<MODE, RT><<<grid, bd, smem>>>(gM, gprog, out, iters);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double mv = (double)grid * bd * iters * NE * NC;
    printf("%-34s %8.3f ms  %.3e matvec/s  %.2f TFLOP/s  (%s)\n", name, ms, mv / (ms * 1e-3), mv * 32 / (ms * 1e-3) / 1e12,
           cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    double *hM = new double[NC * NE * 16];
    for (int i = 0; i < NC * NE * 16; i++) hM[i] = 0.25 + 1e-4 * (i % 7);
    int hp[NE]; for (int i = 0; i < NE; i++) hp[i] = (i * 37) % NE;
    double *gM, *out; int *gp;
    cudaMalloc(&gM, sizeof(double) * NC * NE * 16); cudaMalloc(&gp, sizeof(hp)); cudaMalloc(&out, 8 * 148 * 384);
    cudaMemcpy(gM, hM, sizeof(double) * NC * NE * 16, cudaMemcpyHostToDevice);
    cudaMemcpy(gp, hp, sizeof(hp), cudaMemcpyHostToDevice);
    cudaMemcpyToSymbol(cM, hM, sizeof(double) * NC * NE * 16);
    cudaMemcpyToSymbol(cprog, hp, sizeof(hp));
    run<0, 0>("smem matrix, vector in regs", gM, gp, out);
    run<1, 0>("const matrix, vector in regs", gM, gp, out);
    run<0, 1>("smem matrix, vector via smem", gM, gp, out);
    run<1, 1>("const matrix, vector via smem", gM, gp, out);
    return 0;
}
