// Developer microbenchmark: latency (SM cycles, one thread) of the two serial inversions of fp_inv.cuh.
#include <cstdio>
#include "../snark_verifier_b200/csrc/fp.cuh"
using namespace snarkv;
__global__ void k(unsigned long long* out, uint32_t seed) {
    U256 x, p;
    for (int i = 0; i < 8; ++i) { p.v[i] = fp_mod_limb<FQ>(i); x.v[i] = (seed * 2654435761u) ^ (0x9e3779b9u * (i + 1)); }
    x.v[7] &= 0x0fffffffu;
    long long t0 = clock64();
    U256 a = u256_inv_mod(x, p);
    long long t1 = clock64();
    U256 b = u256_inv_mod_fast(x, p);
    long long t2 = clock64();
    uint32_t same = 1;
    for (int i = 0; i < 8; ++i) same &= (a.v[i] == b.v[i]);
    out[0] = t1 - t0; out[1] = t2 - t1; out[2] = same;
}
int main() {
    unsigned long long* d; cudaMalloc(&d, 64);
    for (int rep = 0; rep < 3; ++rep) {
        k<<<1, 1>>>(d, 12345u + rep);
        unsigned long long h[3]; cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
        printf("cycles: inv_bitwise=%llu inv_pornin31=%llu same=%llu\n", h[0], h[1], h[2]);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
