// Dependent-issue latencies of the FP64 ops on the Cholesky chain (one warp, clock64 around 512 dependent ops).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void lat(double* out, long long* cyc, double a, double b) {
    __shared__ double sm[64];
    double x = a + threadIdx.x * 1e-9, y = b, z = 0.0;
    long long t0, t1;
    const int N = 512;
    // DFMA
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = fma(x, y, y);
    t1 = clock64(); if (threadIdx.x == 0) cyc[0] = t1 - t0;
    // DMUL -> DFMA pair (2 dependent)
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) { z = x * y; x = fma(z, y, x); }
    t1 = clock64(); if (threadIdx.x == 0) cyc[1] = t1 - t0;
    // DMMA dependent through accumulator
    double c0 = x, c1 = z;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) dmma(c0, c1, y, y);
    t1 = clock64(); if (threadIdx.x == 0) cyc[2] = t1 - t0;
    x += c0 + c1;
    // drcp
    x = 1.0 + threadIdx.x * 1e-3;
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; ++i) x = __drcp_rn(x);
    t1 = clock64(); if (threadIdx.x == 0) cyc[3] = t1 - t0;
    // 1.0/x
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; ++i) x = 1.0 / x;
    t1 = clock64(); if (threadIdx.x == 0) cyc[4] = t1 - t0;
    // rsqrt
    x = 1.0 + threadIdx.x * 1e-3;
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; ++i) x = rsqrt(x);
    t1 = clock64(); if (threadIdx.x == 0) cyc[5] = t1 - t0;
    // shfl double
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = __shfl_sync(0xffffffffu, x, (i + 1) & 31);
    t1 = clock64(); if (threadIdx.x == 0) cyc[6] = t1 - t0;
    // STS -> syncwarp -> LDS round trip
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; ++i) { sm[threadIdx.x] = x; __syncwarp(); x = sm[(threadIdx.x + 1) & 31]; __syncwarp(); }
    t1 = clock64(); if (threadIdx.x == 0) cyc[7] = t1 - t0;
    // float rcp approx + 2 Newton steps in double (candidate fast reciprocal)
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; ++i) {
        float xf = (float)x; float rf; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rf) : "f"(xf));
        double r = (double)rf; double e = fma(-x, r, 1.0); r = fma(r, e, r); e = fma(-x, r, 1.0); r = fma(r, e, r); x = r;
    }
    t1 = clock64(); if (threadIdx.x == 0) cyc[8] = t1 - t0;
    // independent DFMA issue rate: 8 accumulators
    double f[8]; for (int i = 0; i < 8; ++i) f[i] = x + i;
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; ++i) {
#pragma unroll
        for (int q = 0; q < 8; ++q) f[q] = fma(f[q], y, y);
    }
    t1 = clock64(); if (threadIdx.x == 0) cyc[9] = t1 - t0;
    for (int i = 0; i < 8; ++i) x += f[i];
    out[threadIdx.x] = x + z;
}
int main() {
    double* d; long long* c; cudaMalloc(&d, 1024); cudaMalloc(&c, 128);
    lat<<<1, 32>>>(d, c, 1.0, 1.0000001); lat<<<1, 32>>>(d, c, 1.0, 1.0000001);
    long long h[16]; cudaMemcpy(h, c, 128, cudaMemcpyDeviceToHost);
    const char* nm[] = {"DFMA dep", "DMUL+DFMA dep pair", "DMMA dep (accum)", "__drcp_rn dep", "1.0/x dep", "rsqrt dep", "shfl f64 dep",
                        "STS+syncwarp+LDS+syncwarp", "f32 rcp + 2 NR (f64)", "8 indep DFMA (per group)"};
    for (int i = 0; i < 10; ++i) printf("%-28s %7.1f cycles/op\n", nm[i], h[i] / 512.0);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
