// ubench.cu -- issue/pipe throughput of the instructions the SWE stage kernels are made of (sm_100a).
// Prints warp-instructions per cycle per SM sub-partition (SMSP) for each op mix.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ubench ubench.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITER 2048
#define NACC 8

enum Op { FFMA, FFMA2, FADD2, FMUL2, FMNMX, FMNMX3, FSEL, MUFU_RCP, MUFU_SQRT, MUFU_RSQ, SHFL, MIX_FFMA_FMNMX,
          MIX_FFMA2_FMNMX, MIX_FFMA2_FMNMX_2to1, FADD, MIX_FFMA2_MUFU, FSET, NOPS };
const char* names[] = {"FFMA", "FFMA2", "FADD2", "FMUL2", "FMNMX", "FMNMX3", "FSETP+FSEL", "MUFU.RCP", "MUFU.SQRT",
                       "MUFU.RSQ", "SHFL", "FFMA+FMNMX 1:1", "FFMA2+FMNMX 1:1", "FFMA2+FMNMX 2:1", "FADD",
                       "FFMA2+MUFU 8:1", "FSET.BF"};
const int instr_per_acc[] = {1, 1, 1, 1, 1, 1, 2, 1, 1, 1, 1, 2, 2, 3, 1, 9, 1};

template <int OP>
__global__ void k(float* out, long long* cyc, float s0, float s1) {
    float a[NACC];
    unsigned long long p[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
        a[i] = threadIdx.x * 0.001f + i;
        float lo = a[i], hi = a[i] + 0.5f;
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(lo), "f"(hi));
    }
    unsigned long long ps0, ps1;
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(ps0) : "f"(s0), "f"(s0));
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(ps1) : "f"(s1), "f"(s1));
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            if (OP == FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(s0), "f"(s1));
            if (OP == FADD) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(s1));
            if (OP == FFMA2) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(ps0), "l"(ps1));
            if (OP == FADD2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(ps1));
            if (OP == FMUL2) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(ps0));
            if (OP == FMNMX) asm volatile("min.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(s0));
            if (OP == FMNMX3) asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(s0), "f"(s1));
            if (OP == FSEL)
                asm volatile("{.reg .pred q; setp.gt.f32 q, %0, %1; selp.f32 %0, %2, %0, q;}" : "+f"(a[i]) : "f"(s0), "f"(s1));
            if (OP == FSET) asm volatile("set.gt.f32.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(s0));
            if (OP == MUFU_RCP) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (OP == MUFU_SQRT) asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (OP == MUFU_RSQ) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (OP == SHFL) a[i] = __shfl_down_sync(0xffffffffu, a[i], 1);
            if (OP == MIX_FFMA_FMNMX) {
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(s0), "f"(s1));
                asm volatile("min.f32 %0, %0, %1;" : "+f"(a[(i + 4) % NACC]) : "f"(s0));
            }
            if (OP == MIX_FFMA2_FMNMX) {
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(ps0), "l"(ps1));
                asm volatile("min.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(s0));
            }
            if (OP == MIX_FFMA2_FMNMX_2to1) {
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(ps0), "l"(ps1));
                asm volatile("min.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(s0));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(ps1), "l"(ps0));
            }
            if (OP == MIX_FFMA2_MUFU) {
#pragma unroll
                for (int r = 0; r < 8; ++r)
                    asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(ps0), "l"(ps1));
                asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            }
        }
    }
    long long t1 = clock64();
    float acc = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
        float lo, hi;
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p[i]));
        acc += a[i] + lo + hi;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(int warps_per_smsp) {
    const int nb = 148, nt = warps_per_smsp * 4 * 32;
    float* out;
    long long* cyc;
    cudaMalloc(&out, sizeof(float) * nb * nt);
    cudaMalloc(&cyc, sizeof(long long) * nb);
    k<OP><<<nb, nt>>>(out, cyc, 0.999f, 0.001f);
    k<OP><<<nb, nt>>>(out, cyc, 0.999f, 0.001f);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < nb; ++i) avg += (double)h[i];
    avg /= nb;
    const double instr = (double)ITER * NACC * instr_per_acc[OP] * warps_per_smsp;   // warp-instr per SMSP
    printf("%-18s warps/SMSP=%d  %.3f warp-instr/cyc/SMSP  (%.2f cyc/instr)\n", names[OP], warps_per_smsp, instr / avg,
           avg / instr);
    cudaFree(out);
    cudaFree(cyc);
}

template <int OP>
void both() {
    run<OP>(4);
    run<OP>(8);
}

int main() {
    both<FFMA>(); both<FADD>(); both<FFMA2>(); both<FADD2>(); both<FMUL2>(); both<FMNMX>(); both<FMNMX3>(); both<FSEL>();
    both<FSET>(); both<MUFU_RCP>(); both<MUFU_SQRT>(); both<MUFU_RSQ>(); both<SHFL>(); both<MIX_FFMA_FMNMX>();
    both<MIX_FFMA2_FMNMX>(); both<MIX_FFMA2_FMNMX_2to1>(); both<MIX_FFMA2_MUFU>();
    cudaError_t e = cudaGetLastError();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
