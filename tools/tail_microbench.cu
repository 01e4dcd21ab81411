// Developer microbenchmark: latency (SM cycles) of the serial-tail building blocks, one warp.
#include <cstdio>
#include "../snark_verifier_b200/csrc/g1.cuh"
using namespace snarkv;
__global__ void k(unsigned long long* out, uint8_t* sink) {
    __shared__ Fq xch[4];
    const int lane = threadIdx.x & 3;
    Fq gx = fp_one<FQ>(); Fq gy = fp_dbl(gx);
    G1Xyzz p = xyzz_dbl_affine(gx, gy);
    G1Xyzz q = xyzz_dbl(p);
    long long t0 = clock64();
    G1Xyzz a = p;
    for (int i = 0; i < 64; ++i) a = xyzz_dbl(a);                 // serial doubling
    long long t1 = clock64();
    G1Xyzz b = p;
    for (int i = 0; i < 64; ++i) b = xyzz_dbl_x4(b, lane, xch);        // 4-lane doubling
    long long t2 = clock64();
    G1Xyzz c = q;
    for (int i = 0; i < 16; ++i) c = xyzz_add_x4(c, a, lane, xch);     // 4-lane add
    long long t3 = clock64();
    Fq inv1 = fp_inv_serial(a.zz);
    long long t4 = clock64();
    Fq inv2 = fp_inv(a.zz);
    long long t5 = clock64();
    Fq m = a.x;
    for (int i = 0; i < 64; ++i) m = fp_mul(m, a.y);              // dependent mulmod chain
    long long t6 = clock64();
    Fq s = a.x;
    for (int i = 0; i < 64; ++i) s = fq_exchange4(xch, fp_mul(fq_sel4(lane, s, a.y, a.x, s), fq_sel4(lane, a.y, s, s, a.x))).v[0];
    long long t7 = clock64();
    if (threadIdx.x == 0) {
        out[0] = (t1 - t0) / 64; out[1] = (t2 - t1) / 64; out[2] = (t3 - t2) / 16; out[3] = t4 - t3; out[4] = t5 - t4;
        out[5] = (t6 - t5) / 64; out[6] = (t7 - t6) / 64;
        out[7] = fp_eq(a.x, b.x) && fp_eq(a.zzz, b.zzz) && fp_eq(inv1, inv2);
        xyzz_store(sink, 0, c); fp_store<FQ>(sink + 128, m); fp_store<FQ>(sink + 160, s);
    }
}
int main() {
    unsigned long long* d; uint8_t* sink; cudaMalloc(&d, 64); cudaMalloc(&sink, 256);
    for (int rep = 0; rep < 2; ++rep) k<<<1, 32>>>(d, sink);
    unsigned long long h[8]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
    printf("cycles: dbl_serial=%llu dbl_x4=%llu add_x4=%llu inv_bgcd=%llu inv_fermat=%llu mulmod_dep=%llu level_x4=%llu consistent=%llu\n",
           h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7]);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
