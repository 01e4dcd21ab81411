// NOTE: only the DFMA / IADD3 / 32-bit IMAD rows are meaningful; ptxas moves the IMAD.WIDE row onto the uniform datapath
// (UIMAD.WIDE) because its operands are warp-uniform.  The IMAD.WIDE cost used in DESIGN.md comes from the ncu capture.
// Microbenchmark (developer tool): warp-instruction throughput of IMAD.WIDE.U32, IMAD, DFMA, IADD3 on sm_100a, alone and mixed.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 4096
template <int MODE>
__global__ void k(uint64_t* out, uint32_t a0, double d0) {
    uint32_t a = a0 + threadIdx.x, b = a0 * 3 + 1;
    uint64_t x0 = threadIdx.x, x1 = 1, x2 = 2, x3 = 3, x4 = 4, x5 = 5, x6 = 6, x7 = 7;
    double f0 = d0, f1 = d0 + 1, f2 = d0 + 2, f3 = d0 + 3, f4 = d0 + 4, f5 = d0 + 5, f6 = d0 + 6, f7 = d0 + 7;
    double m = d0 * 0.5 + 1.0;
    uint32_t i0 = a, i1 = b, i2 = a ^ b, i3 = a + b;
    for (int it = 0; it < ITERS; ++it) {
        if (MODE == 0 || MODE == 3 || MODE == 5) {   // 8 independent IMAD.WIDE
            // loop-carried multiplicand (low word of the accumulator) so that the product cannot be hoisted out of the loop
            asm volatile("{ .reg .u32 t0,t1,t2,t3,t4,t5,t6,t7;\n\t"
                         "cvt.u32.u64 t0, %0; cvt.u32.u64 t1, %1; cvt.u32.u64 t2, %2; cvt.u32.u64 t3, %3;\n\t"
                         "cvt.u32.u64 t4, %4; cvt.u32.u64 t5, %5; cvt.u32.u64 t6, %6; cvt.u32.u64 t7, %7;\n\t"
                         "mad.wide.u32 %0, t0, %9, %0; mad.wide.u32 %1, t1, %9, %1; mad.wide.u32 %2, t2, %9, %2; mad.wide.u32 %3, t3, %9, %3;\n\t"
                         "mad.wide.u32 %4, t4, %9, %4; mad.wide.u32 %5, t5, %9, %5; mad.wide.u32 %6, t6, %9, %6; mad.wide.u32 %7, t7, %9, %7; }"
                         : "+l"(x0), "+l"(x1), "+l"(x2), "+l"(x3), "+l"(x4), "+l"(x5), "+l"(x6), "+l"(x7) : "r"(a), "r"(b));
        }
        if (MODE == 1 || MODE == 3 || MODE == 4) {   // 8 independent DFMA
            asm volatile("fma.rn.f64 %0, %0, %8, %0; fma.rn.f64 %1, %1, %8, %1; fma.rn.f64 %2, %2, %8, %2; fma.rn.f64 %3, %3, %8, %3;"
                         "fma.rn.f64 %4, %4, %8, %4; fma.rn.f64 %5, %5, %8, %5; fma.rn.f64 %6, %6, %8, %6; fma.rn.f64 %7, %7, %8, %7;"
                         : "+d"(f0), "+d"(f1), "+d"(f2), "+d"(f3), "+d"(f4), "+d"(f5), "+d"(f6), "+d"(f7) : "d"(m));
        }
        if (MODE == 2 || MODE == 4 || MODE == 5) {   // 8 IADD3-ish (add.cc chains -> IADD3 / IADD3.X)
            asm volatile("add.cc.u32 %0, %0, %4; addc.cc.u32 %1, %1, %4; addc.cc.u32 %2, %2, %4; addc.u32 %3, %3, %4;"
                         "add.cc.u32 %0, %0, %5; addc.cc.u32 %1, %1, %5; addc.cc.u32 %2, %2, %5; addc.u32 %3, %3, %5;"
                         : "+r"(i0), "+r"(i1), "+r"(i2), "+r"(i3) : "r"(a), "r"(b));
        }
        if (MODE == 6) {   // 8 independent 32-bit IMAD (lo)
            asm volatile("mad.lo.u32 %0, %0, %4, %5; mad.lo.u32 %1, %1, %4, %5; mad.lo.u32 %2, %2, %4, %5; mad.lo.u32 %3, %3, %4, %5;"
                         "mad.lo.u32 %0, %0, %5, %4; mad.lo.u32 %1, %1, %5, %4; mad.lo.u32 %2, %2, %5, %4; mad.lo.u32 %3, %3, %5, %4;"
                         : "+r"(i0), "+r"(i1), "+r"(i2), "+r"(i3) : "r"(a), "r"(b));
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7 ^ (uint64_t)(f0 + f1 + f2 + f3 + f4 + f5 + f6 + f7) ^ i0 ^ i1 ^ i2 ^ i3;
}
template <int MODE> void run(const char* name, int per_iter) {
    uint64_t* d; cudaMalloc(&d, 148 * 8 * 256 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 8, 256>>>(d, 12345, 1.000001);
    cudaEventRecord(e0);
    k<MODE><<<148 * 8, 256>>>(d, 12345, 1.000001);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double warp_insts = 148.0 * 8 * 8 * ITERS * per_iter;   // blocks * warps/block * iters * insts
    double per_sm_clk = warp_insts * 32 / (ms * 1e-3) / 148 / 1.965e9;
    printf("%-28s %8.3f ms  %.1f lanes/clk/SM (at 1965 MHz)\n", name, ms, per_sm_clk);
    cudaFree(d);
}
int main() {
    run<0>("IMAD.WIDE.U32 x8", 8);
    run<6>("IMAD (32-bit) x8", 8);
    run<1>("DFMA x8", 8);
    run<2>("IADD3(.X) x8", 8);
    run<3>("IMAD.WIDE x8 + DFMA x8", 16);
    run<4>("DFMA x8 + IADD3 x8", 16);
    run<5>("IMAD.WIDE x8 + IADD3 x8", 16);
    return 0;
}
