// Issue-rate microbenchmark for packed FP32 (fma.rn.f32x2 -> FFMA2) against scalar FFMA in its operand forms, alone and mixed
// with the half-rate ALU pipe, on sm_100a.  One block of 1024 threads per SM, 8 independent chains per thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o expt/ubench_f32x2 tools/ubench_f32x2.cu
// Prints warp instructions per clock per SM and FP32 FMA lanes per clock per SM (the FLOP rate, FMA = 1 lane-op).
#include <cstdio>
#include <cuda_runtime.h>
#define CH 8
#define ITERS 4096

__device__ __forceinline__ unsigned long long pack2(float a, float b)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}

template <int OP>
__global__ void __launch_bounds__(1024, 1) k(float *out, float seed, long long *cycles)
{
    float a[CH], b[CH], c[CH];
    unsigned long long A[CH], B[CH], Cc[CH];
    unsigned int key = 0xFFFFFFFFu;
#pragma unroll
    for (int i = 0; i < CH; i++) {
        a[i] = seed + i * 0.37f + threadIdx.x * 1e-3f; b[i] = 1.0f + (seed + i) * 1e-6f; c[i] = (seed - i) * 1e-5f;
        A[i] = pack2(a[i], a[i] + 1.f); B[i] = pack2(b[i], b[i] * 1.000001f); Cc[i] = pack2(c[i], c[i] * 2.f);
    }
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) {
            if (OP == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b[0]), "f"(c[0]));                       // FFMA, b and c shared by the chains
            if (OP == 1) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b[i]), "f"(c[i]));                       // FFMA, three distinct registers
            if (OP == 2) asm volatile("fma.rn.f32 %0, %0, 0f3F800054, 0f38D1B717;" : "+f"(a[i]));                             // FFMA, two immediates
            if (OP == 3) asm volatile("fma.rn.f32 %0, %0, %1, 0f38D1B717;" : "+f"(a[i]) : "f"(b[i]));                          // FFMA reg, reg, imm
            if (OP == 4) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(A[i]) : "l"(B[i]), "l"(Cc[i]));                    // FFMA2, three distinct pairs
            if (OP == 5) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(A[i]) : "l"(B[0]), "l"(Cc[0]));                    // FFMA2, b and c shared
            if (OP == 6) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(A[i]) : "l"(B[i]));                                    // FADD2
            if (OP == 7) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(A[i]) : "l"(B[i]));                                    // FMUL2
            if (OP == 8) {   // the scalar rectangle slot: FADD, 3 FFMA, 2 ISETP, LOP3, @p VIMNMX (8 instructions)
                float t = fmaf(b[i] - a[i], c[i], -1.401298464e-45f);
                float wu = fmaf(c[(i + 1) % CH], t, a[i] - 3.0f), wv = fmaf(b[(i + 1) % CH], t, a[i] - 5.0f);
                unsigned int mine;
                asm volatile("lop3.b32 %0, %1, %2, %3, 0x36;" : "=r"(mine) : "r"(__float_as_uint(t)), "r"(63u - (unsigned)i), "r"(__float_as_uint(seed)));
                asm volatile("{.reg .pred p; setp.le.u32 p, %1, 0x42000000; setp.le.and.u32 p, %2, 0x42100000, p; @p min.u32 %0, %0, %3;}"
                             : "+r"(key) : "r"(__float_as_uint(wu)), "r"(__float_as_uint(wv)), "r"(mine));
                a[i] = wu;
            }
            if (OP == 9 && (i & 1) == 0) {   // two rectangles of one axis class per packed op: FADD2, 3 FFMA2, then 2 x (2 ISETP, LOP3, @p VIMNMX) = 12 for two
                unsigned long long T, WU, WV;
                asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(T) : "l"(B[i]), "l"(A[i]));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(T) : "l"(Cc[i]), "l"(Cc[i + 1]));
                asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(WU) : "l"(B[i + 1]), "l"(T), "l"(A[i + 1]));
                asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(WV) : "l"(Cc[i + 1]), "l"(T), "l"(A[i + 1]));
                unsigned int t0b, t1b, u0, u1, v0, v1, m0, m1;
                asm volatile("mov.b64 {%0, %1}, %2;" : "=r"(t0b), "=r"(t1b) : "l"(T));
                asm volatile("mov.b64 {%0, %1}, %2;" : "=r"(u0), "=r"(u1) : "l"(WU));
                asm volatile("mov.b64 {%0, %1}, %2;" : "=r"(v0), "=r"(v1) : "l"(WV));
                asm volatile("lop3.b32 %0, %1, %2, %3, 0x36;" : "=r"(m0) : "r"(t0b), "r"(63u - (unsigned)i), "r"(__float_as_uint(seed)));
                asm volatile("lop3.b32 %0, %1, %2, %3, 0x36;" : "=r"(m1) : "r"(t1b), "r"(62u - (unsigned)i), "r"(__float_as_uint(seed)));
                asm volatile("{.reg .pred p; setp.le.u32 p, %1, 0x42000000; setp.le.and.u32 p, %2, 0x42100000, p; @p min.u32 %0, %0, %3;}" : "+r"(key) : "r"(u0), "r"(v0), "r"(m0));
                asm volatile("{.reg .pred p; setp.le.u32 p, %1, 0x42000000; setp.le.and.u32 p, %2, 0x42100000, p; @p min.u32 %0, %0, %3;}" : "+r"(key) : "r"(u1), "r"(v1), "r"(m1));
                A[i] = WU;
            }
            if (OP == 10) {  // sphere scan, scalar: 7 FFMA (one immediate operand each), FSETP, @p LOP (9 instructions)
                float bb, cc, dd;      // (asm volatile: the loop-invariant operands must not be hoisted)
                asm volatile("fma.rn.f32 %0, %1, 0f3FA00000, %2;" : "=f"(bb) : "f"(a[0]), "f"(a[3]));
                asm volatile("fma.rn.f32 %0, %1, 0f40200000, %0;" : "+f"(bb) : "f"(a[1]));
                asm volatile("fma.rn.f32 %0, %1, 0fC0700000, %0;" : "+f"(bb) : "f"(a[2]));
                asm volatile("fma.rn.f32 %0, %1, 0f3FA00000, %2;" : "=f"(cc) : "f"(b[0]), "f"(b[3]));
                asm volatile("fma.rn.f32 %0, %1, 0f40200000, %0;" : "+f"(cc) : "f"(b[1]));
                asm volatile("fma.rn.f32 %0, %1, 0fC0700000, %0;" : "+f"(cc) : "f"(b[2]));
                asm volatile("fma.rn.f32 %0, %1, %1, %2;" : "=f"(dd) : "f"(bb), "f"(cc));
                asm volatile("{.reg .pred p; setp.ge.f32 p, %1, 0f42C80000; @p or.b32 %0, %0, %2;}" : "+r"(key) : "f"(dd), "r"(1u << i));
            }
            if (OP == 11 && (i & 1) == 0) {  // sphere scan, two spheres per packed op: 7 FFMA2 + 2 x (FSETP, @p LOP) = 11 for two (table pairs in registers)
                unsigned long long BB, CC, D;
                asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(BB) : "l"(B[i]), "l"(A[2]), "l"(A[3]));
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(BB) : "l"(B[i + 1]), "l"(A[1]));
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(BB) : "l"(Cc[i]), "l"(A[0]));
                asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(CC) : "l"(B[i]), "l"(A[6]), "l"(A[7]));
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(CC) : "l"(B[i + 1]), "l"(A[5]));
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(CC) : "l"(Cc[i]), "l"(A[4]));
                asm volatile("fma.rn.f32x2 %0, %1, %1, %2;" : "=l"(D) : "l"(BB), "l"(CC));
                float d0, d1;
                asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(D));
                asm volatile("{.reg .pred p; setp.ge.f32 p, %1, 0f42C80000; @p or.b32 %0, %0, %2;}" : "+r"(key) : "f"(d0), "r"(1u << i));
                asm volatile("{.reg .pred p; setp.ge.f32 p, %1, 0f42C80000; @p or.b32 %0, %0, %2;}" : "+r"(key) : "f"(d1), "r"(2u << i));
            }
            if (OP == 12) {  // FFMA2 + 2 ALU instructions per step: do the two pipes overlap?
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(A[i]) : "l"(B[i]), "l"(Cc[i]));
                asm volatile("{.reg .pred p; setp.le.u32 p, %1, 0x42000000; @p min.u32 %0, %0, %2;}" : "+r"(key) : "r"(__float_as_uint(a[i])), "r"(__float_as_uint(b[i])));
            }
            if (OP == 13) {  // 2 FFMA + 2 ALU per step
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b[i]), "f"(c[i]));
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(c[i]) : "f"(b[i]), "f"(a[(i + 1) % CH]));
                asm volatile("{.reg .pred p; setp.le.u32 p, %1, 0x42000000; @p min.u32 %0, %0, %2;}" : "+r"(key) : "r"(__float_as_uint(b[i])), "r"(__float_as_uint(b[(i + 3) % CH])));
            }
        }
    }
    long long t1 = clock64();
    float s = __uint_as_float(key);
#pragma unroll
    for (int i = 0; i < CH; i++) s += a[i] + c[i] + __uint_as_float((unsigned int)A[i]) + __uint_as_float((unsigned int)(A[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP> void run(const char *name, double instr_per_step, double fma_lanes_per_step)
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float *out; long long *cyc, h[1024];
    cudaMalloc(&out, sms * 1024 * 4); cudaMalloc(&cyc, sms * 8);
    k<OP><<<sms, 1024>>>(out, 1.5f, cyc); cudaDeviceSynchronize();
    k<OP><<<sms, 1024>>>(out, 1.5f, cyc); cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, sms * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < sms; i++) avg += (double)h[i]; avg /= sms;
    const double steps = 32.0 * ITERS * CH;       // warp-steps per SM
    printf("%-62s %6.2f warp-instr/clk/SM  %7.1f FP32 FMA-lanes/clk/SM  (%.1f instr, %.1f FMA per step)  %s\n", name, steps * instr_per_step / avg,
           32.0 * steps * fma_lanes_per_step / avg, instr_per_step, fma_lanes_per_step, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    run<0>("FFMA  (b, c shared by the chains)", 1, 1);
    run<1>("FFMA  (three distinct registers)", 1, 1);
    run<2>("FFMA  (reg, imm, imm)", 1, 1);
    run<3>("FFMA  (reg, reg, imm)", 1, 1);
    run<4>("FFMA2 (three distinct pairs)", 1, 2);
    run<5>("FFMA2 (b, c pairs shared)", 1, 2);
    run<6>("FADD2", 1, 2);
    run<7>("FMUL2", 1, 2);
    run<8>("rectangle slot, scalar (8 instr per rectangle)", 8, 4);
    run<9>("rectangle slot, packed pairs (12 instr + 3 unpack per 2 rect.)", 6, 4);      // per chain step: half a pair
    run<10>("sphere scan, scalar (9 instr per sphere)", 9, 7);
    run<11>("sphere scan, packed pairs (11 instr + unpack per 2 spheres)", 5.5, 7);
    run<12>("FFMA2 + ISETP + @p VIMNMX", 3, 2);
    run<13>("2 FFMA + ISETP + @p VIMNMX", 4, 2);
    return 0;
}
