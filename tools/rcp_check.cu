// Accuracy of the branch-free positive reciprocal used on the Cholesky chain (MUFU seed + 3 FMAs) against 1.0/x.
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
__device__ __forceinline__ double rcp_pos(double d) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    double e = fma(-d, r, 1.0);
    e = fma(e, e, e);
    return fma(r, e, r);
}
__global__ void chk(double* maxerr, double* seederr) {
    unsigned long long s = 0x9E3779B97F4A7C15ULL * (blockIdx.x * blockDim.x + threadIdx.x + 1);
    double worst = 0.0, worst0 = 0.0;
    for (int i = 0; i < 20000; ++i) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        int ex = (int)((s >> 52) % 600) - 300;
        double x = ldexp(1.0 + (double)(s & 0xFFFFFFFFFFFFFULL) / 4503599627370496.0, ex);
        double ref = 1.0 / x;
        double r = rcp_pos(x);
        double r0; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(x));
        worst = fmax(worst, fabs(r - ref) / ref);
        worst0 = fmax(worst0, fabs(r0 - ref) / ref);
    }
    atomicMax((unsigned long long*)maxerr, __double_as_longlong(worst));
    atomicMax((unsigned long long*)seederr, __double_as_longlong(worst0));
}
int main() {
    double* d; cudaMalloc(&d, 16); cudaMemset(d, 0, 16);
    chk<<<64, 256>>>(d, d + 1);
    double h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("rcp_pos max rel err %.3e (%.2f ulp of 2^-53), seed max rel err %.3e (2^%.1f)  %s\n", h[0], h[0] / 1.1102230246251565e-16,
           h[1], log2(h[1]), cudaGetErrorString(cudaDeviceSynchronize()));
}
