// Micro-probe: do the FP64 tensor sub-pipe (DMMA.8x8x4) and the plain FP64 pipe (DFMA) issue concurrently on sm_100a?
// Prints flops/clk/SM for DMMA only, DFMA only and mixes, for 8 and 16 resident warps per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/build/pipe_probe tools/pipe_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NM, int NF>
__global__ void __launch_bounds__(512) probe(double* out, int iters, double a0, double b0) {
    double c[16][2];
    double f[32];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i][0] = c[i][1] = 0.0;
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] = threadIdx.x * 1e-9 + i;
    double a = a0 + threadIdx.x * 1e-12, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < NM / 4; ++i) dmma(c[i % 16][0], c[i % 16][1], a, b);
#pragma unroll
            for (int i = 0; i < NF / 4; ++i) asm volatile("fma.rn.f64 %0, %1, %2, %0;\n" : "+d"(f[i % 32]) : "d"(a), "d"(b));
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < 32; ++i) s += f[i];
    if (s == 123.456) out[0] = s;
}

template <int NM, int NF>
void run(int warps_per_sm, int sms, double clock_ghz, double* d) {
    int iters = 20000;
    int threads = warps_per_sm * 32;
    probe<NM, NF><<<sms, threads>>>(d, 100, 1.0, 1.0);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<NM, NF><<<sms, threads>>>(d, iters, 1.0, 1.0);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double warp_instr = (double)iters * sms * warps_per_sm;
    double fl_m = warp_instr * NM * 512.0, fl_f = warp_instr * NF * 64.0;
    double tf = (fl_m + fl_f) / (ms * 1e-3) / 1e12;
    printf("warps/SM %2d  NM %3d NF %3d  %8.3f ms  total %6.2f TF/s  (dmma %6.2f + dfma %6.2f)  [%s]\n", warps_per_sm, NM, NF,
           ms, tf, fl_m / (ms * 1e-3) / 1e12, fl_f / (ms * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    double* d; cudaMalloc(&d, 64);
    printf("%s  SMs %d  clock %.0f MHz\n", p.name, sms, p.clockRate / 1e3);
    for (int w : {4, 8, 16}) {
        run<32, 0>(w, sms, 0, d);
        run<0, 32>(w, sms, 0, d);
        run<0, 128>(w, sms, 0, d);
        run<32, 32>(w, sms, 0, d);
        run<32, 64>(w, sms, 0, d);
        run<32, 128>(w, sms, 0, d);
        run<32, 256>(w, sms, 0, d);
    }
    return 0;
}
