// FP64 tensor-pipe micro-measurements behind the design of phyly_b200/csrc/dmma.cu:
//  (1) throughput of mma.sync m8n8k4.f64 and m16n8k16.f64 as a function of warps per SM and of independent
//      accumulator chains per warp (how much instruction-level parallelism one warp needs to fill the pipe);
//  (2) the fragment layout of m16n8k16.f64, found by one-hot probing: row of every A register, column of every
//      B register, and which A registers pair with which B registers along k.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -cudart shared -o dmma_probe dmma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define MMA884(c0, c1, a, b) \
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b))
#define MMA16816(c, a, b) \
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};" \
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) \
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]))

template <int CH>
__global__ void k884(double *out, int iters)
{
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
    double c[CH][2];
#pragma unroll
    for (int i = 0; i < CH; i++) { c[i][0] = 0; c[i][1] = 0; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) MMA884(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CH>
__global__ void k16816(double *out, int iters)
{
    double a[8], b[4];
    for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 1e-3 + i;
    for (int i = 0; i < 4; i++) b[i] = 1.0 + threadIdx.x * 1e-6 + i;
    double c[CH][4];
#pragma unroll
    for (int i = 0; i < CH; i++) { c[i][0] = 0; c[i][1] = 0; c[i][2] = 0; c[i][3] = 0; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) MMA16816(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}


// the loop shape of dm_gemm_chunk: 2 NB = 16 k-steps, per k-step two 16-byte LDS of B fragments feeding 4 DMMAs that share
// the A register; 4 accumulator pairs.  Measures what the tensor pipe sustains when every DMMA reads a fresh B operand
// from shared memory (no operand reuse), as a function of the warps per SM.
__global__ void k884_lds(double *out, int iters)
{
    extern __shared__ double2 slot[];          // [16 k-steps][2 pairs][32 lanes] double2 = 16 KB
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 16 * 2 * 32; i += blockDim.x) slot[i] = make_double2(1.0 + i * 1e-9, 1.0 - i * 1e-9);
    __syncthreads();
    double A[16];
    for (int i = 0; i < 16; i++) A[i] = threadIdx.x * 1e-3 + i;
    double c[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
    const double2 *sl = slot + lane;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int kk = 0; kk < 16; kk++) {
#pragma unroll
            for (int p = 0; p < 2; p++) {
                const double2 b = sl[(kk * 2 + p) * 32];
                MMA884(c[2 * p][0], c[2 * p][1], A[kk], b.x);
                MMA884(c[2 * p + 1][0], c[2 * p + 1][1], A[kk], b.y);
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c[0][0] + c[0][1] + c[1][0] + c[1][1] + c[2][0] + c[2][1] + c[3][0] + c[3][1];
}

// the same DMMA stream with the B operands in registers (distinct registers, no shared memory)
__global__ void k884_regs(double *out, int iters)
{
    double A[16], B[8];
    for (int i = 0; i < 16; i++) A[i] = threadIdx.x * 1e-3 + i;
    for (int i = 0; i < 8; i++) B[i] = 1.0 + threadIdx.x * 1e-6 + i;
    double c[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int kk = 0; kk < 16; kk++) {
#pragma unroll
            for (int p = 0; p < 2; p++) {
                MMA884(c[2 * p][0], c[2 * p][1], A[kk], B[(2 * kk + p) & 7]);
                MMA884(c[2 * p + 1][0], c[2 * p + 1][1], A[kk], B[(2 * kk + p + 3) & 7]);
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c[0][0] + c[0][1] + c[1][0] + c[1][1] + c[2][0] + c[2][1] + c[3][0] + c[3][1];
}

template <typename K>
static double tflops_smem(K k, double *out, int sms, int warps_per_sm, int iters, double flop_per_warp_iter, size_t smem)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<<<sms, warps_per_sm * 32, smem>>>(out, iters / 10);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0);
        k<<<sms, warps_per_sm * 32, smem>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return flop_per_warp_iter * iters * (double)sms * warps_per_sm / (best * 1e-3) / 1e12;
}

template <typename K>
static double tflops(K k, double *out, int sms, int warps_per_sm, int iters, double flop_per_warp_iter)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<<<sms, warps_per_sm * 32>>>(out, iters / 10);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0);
        k<<<sms, warps_per_sm * 32>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return flop_per_warp_iter * iters * (double)sms * warps_per_sm / (best * 1e-3) / 1e12;
}

// one-hot probe of m16n8k16: A element (la, ra) = 1, B element (lb, rb) = 1
__global__ void probe16816(int *rowA, int *colB, int *pairs)
{
    const int lane = threadIdx.x;
    // rows of A registers: B all ones
    for (int la = 0; la < 32; la++) for (int ra = 0; ra < 8; ra++) {
        double a[8], b[4], c[4] = {0, 0, 0, 0};
        for (int i = 0; i < 8; i++) a[i] = (lane == la && i == ra) ? 1.0 : 0.0;
        for (int i = 0; i < 4; i++) b[i] = 1.0;
        MMA16816(c, a, b);
        // C layout (known): c0,c1 row g; c2,c3 row g+8
        if (c[0] != 0.0 && (lane & 3) == 0) rowA[la * 8 + ra] = lane >> 2;
        if (c[2] != 0.0 && (lane & 3) == 0) rowA[la * 8 + ra] = (lane >> 2) + 8;
    }
    // columns of B registers: A all ones
    for (int lb = 0; lb < 32; lb++) for (int rb = 0; rb < 4; rb++) {
        double a[8], b[4], c[4] = {0, 0, 0, 0};
        for (int i = 0; i < 8; i++) a[i] = 1.0;
        for (int i = 0; i < 4; i++) b[i] = (lane == lb && i == rb) ? 1.0 : 0.0;
        MMA16816(c, a, b);
        if (lane < 4) {
            if (c[0] != 0.0) colB[lb * 4 + rb] = 2 * lane;
            if (c[1] != 0.0) colB[lb * 4 + rb] = 2 * lane + 1;
        }
    }
    // pairing along k: for the A registers of lanes 0..3 (row 0) against every B register
    for (int la = 0; la < 4; la++) for (int ra = 0; ra < 8; ra += 2) {
        for (int lb = 0; lb < 32; lb++) for (int rb = 0; rb < 4; rb++) {
            double a[8], b[4], c[4] = {0, 0, 0, 0};
            for (int i = 0; i < 8; i++) a[i] = (lane == la && i == ra) ? 1.0 : 0.0;
            for (int i = 0; i < 4; i++) b[i] = (lane == lb && i == rb) ? 1.0 : 0.0;
            MMA16816(c, a, b);
            unsigned any = __ballot_sync(0xffffffffu, c[0] != 0.0 || c[1] != 0.0 || c[2] != 0.0 || c[3] != 0.0);
            if (any && lane == 0) pairs[((la * 4 + ra / 2) * 32 + lb) * 4 + rb] = 1;
        }
    }
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    double *out; cudaMalloc(&out, sizeof(double) * sms * 1024);
    const int it = 40000;
    printf("{\"gpu\": \"%s\", \"sms\": %d,\n \"m8n8k4_tflops\": {", p.name, sms);
    const int ws[4] = {4, 8, 16, 32};
    for (int wi = 0; wi < 4; wi++) {
        const int w = ws[wi];
        const double f = 2.0 * 8 * 8 * 4;
        printf("%s\"warps%d\": {\"chains1\": %.2f, \"chains2\": %.2f, \"chains4\": %.2f, \"chains8\": %.2f}", wi ? ", " : "", w,
               tflops(k884<1>, out, sms, w, it, f * 1), tflops(k884<2>, out, sms, w, it, f * 2),
               tflops(k884<4>, out, sms, w, it, f * 4), tflops(k884<8>, out, sms, w, it, f * 8));
    }
    printf("},\n \"m8n8k4_gemm_loop_tflops\": {");
    for (int wi = 0; wi < 4; wi++) {
        const int w = ws[wi];
        const double f = 2.0 * 8 * 8 * 4 * 64;        /* 64 DMMAs per iteration */
        printf("%s\"warps%d\": {\"B_from_shared_memory\": %.2f, \"B_in_registers\": %.2f}", wi ? ", " : "", w,
               tflops_smem(k884_lds, out, sms, w, it / 16, f, 16384), tflops_smem(k884_regs, out, sms, w, it / 16, f, 0));
    }
    printf("},\n \"m16n8k16_tflops\": {");
    for (int wi = 0; wi < 3; wi++) {
        const int w = ws[wi];
        const double f = 2.0 * 16 * 8 * 16;
        printf("%s\"warps%d\": {\"chains1\": %.2f, \"chains2\": %.2f, \"chains4\": %.2f}", wi ? ", " : "", w,
               tflops(k16816<1>, out, sms, w, it / 8, f * 1), tflops(k16816<2>, out, sms, w, it / 8, f * 2),
               tflops(k16816<4>, out, sms, w, it / 8, f * 4));
    }
    printf("},\n");
    int *rowA, *colB, *pairs;
    cudaMallocManaged(&rowA, sizeof(int) * 256); cudaMallocManaged(&colB, sizeof(int) * 128); cudaMallocManaged(&pairs, sizeof(int) * 16 * 128);
    for (int i = 0; i < 256; i++) rowA[i] = -1;
    for (int i = 0; i < 128; i++) colB[i] = -1;
    for (int i = 0; i < 16 * 128; i++) pairs[i] = 0;
    probe16816<<<1, 32>>>(rowA, colB, pairs);
    cudaDeviceSynchronize();
    printf(" \"m16n8k16_A_row\": [");
    for (int l = 0; l < 32; l++) { printf("%s[", l ? ", " : ""); for (int r = 0; r < 8; r++) printf("%s%d", r ? "," : "", rowA[l * 8 + r]); printf("]"); }
    printf("],\n \"m16n8k16_B_col\": [");
    for (int l = 0; l < 32; l++) { printf("%s[", l ? ", " : ""); for (int r = 0; r < 4; r++) printf("%s%d", r ? "," : "", colB[l * 4 + r]); printf("]"); }
    printf("],\n \"m16n8k16_k_pairs\": {");
    for (int la = 0; la < 4; la++) for (int rh = 0; rh < 4; rh++) {
        printf("%s\"A(lane%d,reg%d)\": [", (la || rh) ? ", " : "", la, 2 * rh);
        bool first = true;
        for (int lb = 0; lb < 32; lb++) for (int rb = 0; rb < 4; rb++)
            if (pairs[((la * 4 + rh) * 32 + lb) * 4 + rb] && (lb >> 2) == 0) { printf("%s\"B(lane%d,reg%d)\"", first ? "" : ",", lb, rb); first = false; }
        printf("]");
    }
    printf("}}\n");
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { fprintf(stderr, "cuda error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
