// Issue-rate microbenchmark for the instructions of the rectangle slot (one block of 1024 threads per SM, 8 independent
// chains per thread): prints warp instructions per clock per SM.   nvcc -arch=sm_100a -o expt/ubench_alu tools/ubench_alu.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CH 8
#define ITERS 4096
template <int OP>
__global__ void k(unsigned int *out, unsigned int seed, long long *cycles)
{
    unsigned int a[CH], b = seed + threadIdx.x, c = seed * 3u + 7u + threadIdx.x;
#pragma unroll
    for (int i = 0; i < CH; i++) a[i] = seed + i * 977u + threadIdx.x;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) {
            if (OP == 0) asm volatile("min.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));                                    // VIMNMX
            if (OP == 1) asm volatile("{.reg .pred p; setp.gt.u32 p, %0, %2; @p min.u32 %0, %0, %1;}" : "+r"(a[i]) : "r"(b), "r"(c));  // ISETP + @p VIMNMX
            if (OP == 2) asm volatile("lop3.b32 %0, %0, %1, %2, 0x36;" : "+r"(a[i]) : "r"(b), "r"(c));                // LOP3
            if (OP == 3) asm volatile("{.reg .pred p; setp.le.u32 p, %0, %1; selp.b32 %0, %1, %0, p;}" : "+r"(a[i]) : "r"(b));  // ISETP + SEL
            if (OP == 4) asm volatile("{.reg .pred p; setp.gt.u32 p, %0, %2; @p mov.b32 %0, %1;}" : "+r"(a[i]) : "r"(b ^ a[(i + 1) % CH]), "r"(c));  // LOP3 + ISETP + @p MOV
            if (OP == 5) asm volatile("min.f32 %0, %0, %1;" : "+f"(*(float *)&a[i]) : "f"(__uint_as_float(b)));       // FMNMX
            if (OP == 6) asm volatile("{.reg .pred p; setp.gt.u32 p, %0, %2; @p fma.rn.f32 %0, %0, %1, %1;}" : "+r"(a[i]) : "r"(b), "r"(c));      // ISETP + @p FFMA
            if (OP == 7) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(*(float *)&a[i]) : "f"(__uint_as_float(b)), "f"(__uint_as_float(c)));   // FFMA
        }
    }
    long long t1 = clock64();
    unsigned int s = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) s ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}
template <int OP> void run(const char *name, int per_iter)
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    unsigned int *out; long long *cyc, h[1024];
    cudaMalloc(&out, sms * 1024 * 4); cudaMalloc(&cyc, sms * 8);
    k<OP><<<sms, 1024>>>(out, 12345u, cyc); cudaDeviceSynchronize();
    k<OP><<<sms, 1024>>>(out, 12345u, cyc); cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, sms * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < sms; i++) avg += (double)h[i]; avg /= sms;
    printf("%-28s %6.2f warp-instr/clk/SM (%d instr per step)  err=%s\n", name, 32.0 * ITERS * CH * per_iter / avg, per_iter, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc);
}
int main()
{
    run<7>("FFMA", 1); run<2>("LOP3", 1); run<0>("VIMNMX.U32", 1); run<5>("FMNMX", 1); run<6>("ISETP + @p FFMA", 2);
    run<1>("ISETP + @p VIMNMX", 2); run<3>("ISETP + SEL", 2); run<4>("LOP3 + ISETP + @p MOV", 3);
    return 0;
}
