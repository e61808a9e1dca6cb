// rcp_check.cu -- exhaustive-style check that dm_rcp_flag (detmath.h, CUB_FAST_RCP) returns the bits of the IEEE
// division 1.0/x whenever it does not raise its `bad` flag.  Build: nvcc -arch=sm_100a -fmad=false -DCUB_FAST_RCP=1
// -o tools/rcp_check tools/rcp_check.cu ; run on a GPU box.  Patterns: random mantissas x random exponents over the
// whole double range, plus adversarial mantissas (all ones, near powers of two, single bits, low-word edge cases).
#include <cstdio>
#include <cstdlib>
#define CUB_FAST_RCP 1
#include "../cubature_b200/csrc/detmath.h"

__device__ unsigned long long splitmix(unsigned long long& s) {
    unsigned long long z = (s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

__global__ void k_check(unsigned long long seed, int iters, unsigned long long* stats, unsigned long long* first_bad) {
    unsigned long long s = seed + 0x1234567ULL * (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x);
    unsigned long long mism = 0, flagged = 0, tested = 0;
    for (int it = 0; it < iters; ++it) {
        const unsigned long long r = splitmix(s);
        unsigned long long mant = r & 0xFFFFFFFFFFFFFULL;
        const unsigned kind = (unsigned)(r >> 52) & 15u;
        const unsigned long long r2 = splitmix(s);
        switch (kind) {
            case 0: mant = 0xFFFFFFFFFFFFFULL - (r2 & 0xFF); break;                // all ones minus a little
            case 1: mant = r2 & 0xFFF; break;                                      // just above a power of two
            case 2: mant = 1ULL << (r2 % 52); break;                               // single bit
            case 3: mant = 0xFFFFFFFFFFFFFULL ^ (1ULL << (r2 % 52)); break;        // single zero
            case 4: mant = (mant & 0xFFFFF00000000ULL) | (r2 & 0xFFFFFFFFULL ? 0xFFFFFFFFULL : 0); break;  // low word all ones
            case 5: mant = mant & 0xFFFFF00000000ULL; break;                       // low word zero
            case 6: mant = (r2 & 0xFFFFFULL) << 32 | 0xFFFFFFFFULL; break;
            default: break;                                                        // random
        }
        unsigned long long expo = (r2 >> 20) % 2047ULL;  // includes 0 (subnormal/zero); 2047 excluded (inf/nan tested below)
        if ((r2 & 7) < 5) expo = 1023 - 40 + (r2 >> 40) % 80;  // most mass near 1.0, where the integrands live
        const unsigned long long sign = (r2 >> 3) & 1ULL;
        const double x = dm_from_bits((sign << 63) | (expo << 52) | mant);
        int bad = 0;
        const double fast = dm_rcp_flag(x, &bad);
        const double exact = 1.0 / x;
        ++tested;
        if (bad) {
            ++flagged;
        } else if (dm_bits(fast) != dm_bits(exact)) {
            if (mism == 0) atomicCAS(first_bad, 0ULL, dm_bits(x));
            ++mism;
        }
    }
    atomicAdd(&stats[0], tested);
    atomicAdd(&stats[1], flagged);
    atomicAdd(&stats[2], mism);
}

__global__ void k_special(unsigned long long* out) {
    const double xs[8] = {0.0, -0.0, dm_from_bits(0x7FF0000000000000ULL), dm_from_bits(0x7FF8000000000000ULL), 4.9e-324, 1.7976931348623157e308,
                          2.2250738585072014e-308, dm_from_bits(0x7FD0000000000000ULL)};
    for (int i = 0; i < 8; ++i) {
        int bad = 0;
        const double f = dm_rcp_flag(xs[i], &bad);
        const double e = 1.0 / xs[i];
        out[i] = bad ? 2 : (dm_bits(f) == dm_bits(e) ? 0 : 1);
    }
}

int main(int argc, char** argv) {
    const int launches = argc > 1 ? atoi(argv[1]) : 16;
    unsigned long long *d_stats, *d_first, h_stats[3], h_first, h_sp[8];
    cudaMalloc(&d_stats, 64 + 8 * 8);
    cudaMemset(d_stats, 0, 64 + 8 * 8);
    d_first = d_stats + 4;
    for (int l = 0; l < launches; ++l) k_check<<<148 * 8, 256>>>(0xC0FFEEULL * (l + 1), 1 << 14, d_stats, d_first);
    k_special<<<1, 1>>>(d_stats + 8);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h_stats, d_stats, 24, cudaMemcpyDeviceToHost);
    cudaMemcpy(&h_first, d_first, 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(h_sp, d_stats + 8, 64, cudaMemcpyDeviceToHost);
    printf("rcp_check: %s tested %llu flagged(out of range) %llu MISMATCHES %llu first_bad_x 0x%016llx\n", cudaGetErrorString(e), h_stats[0],
           h_stats[1], h_stats[2], h_first);
    printf("specials (0 ok, 1 MISMATCH, 2 flagged): ");
    int sp_bad = 0;
    for (int i = 0; i < 8; ++i) {
        printf("%llu ", h_sp[i]);
        sp_bad |= h_sp[i] == 1;
    }
    printf("\n");
    return (e != cudaSuccess || h_stats[2] != 0 || sp_bad) ? 1 : 0;
}
