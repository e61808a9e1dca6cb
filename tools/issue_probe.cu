// issue_probe.cu -- does a non-FP64 instruction cost the FP64 pipe a slot on sm_100a?
//
// The scalar rule kernel (cub_gm_scalar) executes 62 % FP64 instructions and 38 % others (reciprocal seeds, range
// checks, selects, addresses) and its FP64 pipe is busy 78 % of the time whatever the occupancy, CTA size or load
// latency (profiles/r02/summary.md, section 6).  Model: a sub-partition issues at most one warp instruction per
// cycle and a warp-wide FP64 instruction occupies it for two (16 FP64 lanes per sub-partition), so a stream of N64
// FP64 and Nx other instructions needs 2 N64 + Nx issue cycles and the pipe can be busy at most 2 N64 / (2 N64 + Nx)
// of the time.  This probe measures exactly that: 8 independent DFMA chains per thread plus K integer instructions
// per 8 DFMAs, 16 warps per SM like the rule kernel, and prints the FP64 rate next to the model's.
//
// Build (no GPU needed): nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/issue_probe tools/issue_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int K, int MIX>
__global__ void __launch_bounds__(128, 4) k_probe(double* out, int iters, double b, double c) {
    double a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = 1.0 + 1e-3 * (threadIdx.x + j);
    unsigned r0 = threadIdx.x, r1 = threadIdx.x * 7u + blockIdx.x, r2 = 0x9e3779b9u + threadIdx.x, r3 = 0x7f4a7c15u ^ threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (MIX == 0 || j % 3 == 0)
                a[j] = fma(a[j], b, c);
            else if (j % 3 == 1)
                a[j] = a[j] * b;  // (b = 1 - 2^-40: stays finite)
            else
                a[j] = a[j] + c;
            // K integer instructions spread over the 8 FP64 ones (independent of them)
            if (K > 0 && (j * K) / 8 != ((j + 1) * K) / 8) {
#pragma unroll
                for (int q = (j * K) / 8; q < ((j + 1) * K) / 8; ++q) {
                    if (q & 1)
                        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r0) : "r"(r2), "r"(r3));
                    else
                        asm volatile("add.u32 %0, %0, %1;" : "+r"(r1) : "r"(r3));
                }
            }
        }
    }
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += a[j];
    if (s == 123.456 || r0 + r1 == 0x12345u) out[0] = s;  // keep everything alive
}

template <int K, int MIX>
static void run(const char* label, int sms, double* d_out) {
    const int iters = 1 << 15, grid = sms * 4;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_probe<K, MIX><<<grid, 128>>>(d_out, 1 << 10, 1.0 - 9.094947017729282e-13, 1e-9);  // warm-up
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k_probe<K, MIX><<<grid, 128>>>(d_out, iters, 1.0 - 9.094947017729282e-13, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double fp64_warp_insts = (double)grid * 4 /*warps*/ * iters * 8;
    const double per_sm_per_us = fp64_warp_insts / sms / (best * 1e3);
    printf("%-34s K=%d  %.3f ms  FP64 warp-insts per SM per us %.1f  model busy %.3f\n", label, K, best, per_sm_per_us,
           16.0 / (16.0 + K));
}

int main() {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, 0) != cudaSuccess) {
        fprintf(stderr, "no CUDA device\n");
        return 1;
    }
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    printf("%s, %d SMs, nominal %d MHz: peak = 2 FP64 warp-insts per SM per cycle = %.1f per us\n", p.name, p.multiProcessorCount,
           khz / 1000, 2.0 * khz / 1000.0);
    double* d_out = nullptr;
    cudaMalloc(&d_out, 8);
    const int sms = p.multiProcessorCount;
    run<0, 0>("DFMA only", sms, d_out);
    run<0, 1>("DFMA/DMUL/DADD mix", sms, d_out);
    run<1, 0>("DFMA + 1 int per 8", sms, d_out);
    run<2, 0>("DFMA + 2 int per 8", sms, d_out);
    run<4, 0>("DFMA + 4 int per 8", sms, d_out);
    run<5, 1>("mix + 5 int per 8 (rule kernel)", sms, d_out);
    run<8, 0>("DFMA + 8 int per 8", sms, d_out);
    run<16, 0>("DFMA + 16 int per 8", sms, d_out);
    cudaFree(d_out);
    return 0;
}
