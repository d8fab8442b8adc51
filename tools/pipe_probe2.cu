// Which ingredient of the SYRK main loop costs DMMA duty?  Same loop skeleton, ingredients switched on one at a time.
//   flags: 1 = fragments from shared memory (LDS.64, padded layout), 2 = DMUL scaling of the B fragments,
//          4 = __syncthreads per 16-k slab, 8 = cp.async of the next slab from global memory
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem) : "memory");
}
constexpr int A_LD = 20;
template <int MI, int NI, int FLAGS>
__global__ void __launch_bounds__(512, 1) probe(double* out, const double* gsrc, int slabs) {
    extern __shared__ double smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const int nw = blockDim.x >> 5;
    const int wm = (nw == 16) ? (warp >> 2) : (warp >> 2), wn = warp & 3;
    const int stage_doubles = 2 * 128 * A_LD + 16;
    for (int i = tid; i < 4 * stage_doubles; i += blockDim.x) smem[i] = 1.0 + 1e-9 * i;
    __syncthreads();
    double acc[MI][NI][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    double a[MI], b[NI];
#pragma unroll
    for (int i = 0; i < MI; ++i) a[i] = 1.0 + i;
#pragma unroll
    for (int j = 0; j < NI; ++j) b[j] = 1.0 + j;
    const double* gp = gsrc + (size_t)blockIdx.x * 65536 + tid * 2;
    for (int s = 0; s < slabs; ++s) {
        if (FLAGS & 4) __syncthreads();
        if (FLAGS & 8) {
            double* dst = smem + ((s + 3) & 3) * stage_doubles;
            const int per = 2048 / blockDim.x;           // 2048 16-byte chunks per stage (A + B slabs)
            for (int q = 0; q < per; ++q) cp_async16(dst + (tid + q * blockDim.x) * 2, gp + ((s & 15) * 4096) + q * 1024);
            asm volatile("cp.async.commit_group;\n" ::: "memory");
            asm volatile("cp.async.wait_group 2;\n" ::: "memory");
        }
        const double* As = smem + (s & 3) * stage_doubles;
        const double* Bs = As + 128 * A_LD;
        const double* Xs = Bs + 128 * A_LD;
        const double* ap = As + (wm * (MI * 8) + g) * A_LD + t;
        const double* bp = Bs + (wn * (NI * 8) + g) * A_LD + t;
        if (FLAGS & 16) {      // register double buffering of the fragments (8-warp configuration of round 1)
            double a2[2][MI], b2[2][NI];
            const double xv0 = (FLAGS & 2) ? Xs[t] : 1.0;
#pragma unroll
            for (int i = 0; i < MI; ++i) a2[0][i] = ap[i * 8 * A_LD];
#pragma unroll
            for (int j = 0; j < NI; ++j) b2[0][j] = bp[j * 8 * A_LD] * xv0;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const int cur = kk & 1, nxt = cur ^ 1;
                if (kk + 1 < 4) {
                    const double xv = (FLAGS & 2) ? Xs[(kk + 1) * 4 + t] : 1.0;
#pragma unroll
                    for (int i = 0; i < MI; ++i) a2[nxt][i] = ap[i * 8 * A_LD + (kk + 1) * 4];
#pragma unroll
                    for (int j = 0; j < NI; ++j) b2[nxt][j] = bp[j * 8 * A_LD + (kk + 1) * 4] * xv;
                }
#pragma unroll
                for (int i = 0; i < MI; ++i)
#pragma unroll
                    for (int j = 0; j < NI; ++j) dmma(acc[i][j][0], acc[i][j][1], a2[cur][i], b2[cur][j]);
            }
            continue;
        }
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            if (FLAGS & 1) {
#pragma unroll
                for (int i = 0; i < MI; ++i) a[i] = ap[i * 8 * A_LD + kk * 4];
#pragma unroll
                for (int j = 0; j < NI; ++j) b[j] = bp[j * 8 * A_LD + kk * 4];
            }
            if (FLAGS & 2) {
                const double xv = Xs[kk * 4 + t];
#pragma unroll
                for (int j = 0; j < NI; ++j) b[j] *= xv;
            }
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NI; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    double sum = 0.0;
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) sum += acc[i][j][0] + acc[i][j][1];
    if (sum == 123.456) out[0] = sum;
}
template <int MI, int NI, int FLAGS>
void run(int threads, int sms, double* d, const double* gsrc) {
    int slabs = 4000;
    size_t smem = 4 * (2 * 128 * A_LD + 16) * 8;
    cudaFuncSetAttribute(probe<MI, NI, FLAGS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe<MI, NI, FLAGS><<<sms, threads, smem>>>(d, gsrc, 50);
    printf("warm: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<MI, NI, FLAGS><<<sms, threads, smem>>>(d, gsrc, slabs);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fl = (double)slabs * 4 * MI * NI * 512.0 * (threads / 32) * sms;
    printf("warps %2d tile %dx%d flags %2d : %7.3f ms  %6.2f TF/s  [%s]\n", threads / 32, MI * 8, NI * 8, FLAGS, ms, fl / (ms * 1e-3) / 1e12,
           cudaGetErrorString(cudaGetLastError()));
}
int main() {
    setvbuf(stdout, NULL, _IONBF, 0);
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    printf("sms %d\n", sms);
    double* d; cudaMalloc(&d, 64);
    double* gsrc; cudaMalloc(&gsrc, (size_t)sms * 65536 * 8 + (1 << 20)); cudaMemset(gsrc, 0, (size_t)sms * 65536 * 8 + (1 << 20));
    run<8, 4, 0>(256, sms, d, gsrc);  run<8, 4, 1>(256, sms, d, gsrc);  run<8, 4, 3>(256, sms, d, gsrc);
    run<8, 4, 7>(256, sms, d, gsrc);  run<8, 4, 15>(256, sms, d, gsrc); run<8, 4, 13>(256, sms, d, gsrc);
    run<8, 4, 17>(256, sms, d, gsrc); run<8, 4, 19>(256, sms, d, gsrc); run<8, 4, 23>(256, sms, d, gsrc); run<8, 4, 31>(256, sms, d, gsrc);
    run<4, 4, 17>(512, sms, d, gsrc); run<4, 4, 19>(512, sms, d, gsrc);
    run<4, 4, 0>(512, sms, d, gsrc);  run<4, 4, 1>(512, sms, d, gsrc);  run<4, 4, 3>(512, sms, d, gsrc);
    run<4, 4, 7>(512, sms, d, gsrc);  run<4, 4, 15>(512, sms, d, gsrc); run<4, 4, 13>(512, sms, d, gsrc);
    return 0;
}
