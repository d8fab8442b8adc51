// What slows the in-warp 32x32 factorisation down when other warps of the CTA are active?  (one CTA, clock64 around factor32)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/build/factor_probe tools/factor_probe.cu \
//        accbpg_and_fw_b200/csrc/build/prof.o accbpg_and_fw_b200/csrc/build/vecops.o
#include "../accbpg_and_fw_b200/csrc/chol.cu"
#include <vector>
using namespace accbpg;

__global__ void __launch_bounds__(256, 1) probe(const double* M, long long* out, int mode, int* gflag, int rolled) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* sD = reinterpret_cast<double*>(smem_raw);
    double* sS = sD + CBUF;
    double* rinv = sS + CBUF;
    double* colbuf = rinv + 64;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int e = tid; e < CB * CB; e += blockDim.x) sD[(e >> 6) * CLD + (e & 63)] = M[e];
    __syncthreads();
    bool bad = false;
    long long t0 = 0, t1 = 0;
    double sink = 0.0;
    if (warp == 0) {
        t0 = clock64();
        sink = factor32(sD, 0, rinv, colbuf, sS, &bad);
        t1 = clock64();
        __syncwarp();
        long long t2 = clock64();
        invert32(sS, 0, rinv, sD + 32 * CLD + 32);
        long long t3 = clock64();
        if (lane == 0) { out[0] = t1 - t0; out[1] = t3 - t2; }
        if (lane == 5) { out[4] = __double_as_longlong(sD[5 * CLD + 3]); out[5] = __double_as_longlong(sD[(32 + 7) * CLD + 32 + 2]); }
    } else {
        if (mode == 1) {                       // spin on a global flag with nanosleep
            if (lane == 0) { for (int q = 0; q < 60; ++q) { if (ld_acquire_gpu(gflag) == 12345) break; __nanosleep(100); } }
        } else if (mode == 2) {                // DMMA work from shared memory
            double acc[4][2][2];
            acc_zero(acc);
            for (int rep = 0; rep < 6; ++rep) mma64<true, true>(acc, sD, sD, 0, CB, (warp >> 2), (warp & 3), lane >> 2, lane & 3);
            sink = acc[0][0][0];
        } else if (mode == 3) {                // tight spin without sleeping
            if (lane == 0) { for (int q = 0; q < 4000; ++q) { if (ld_acquire_gpu(gflag) == 12345) break; } }
        } else if (mode == 4) {                // named-barrier wait of the other warps only
            asm volatile("bar.sync 3, 224;" ::: "memory");
        }
    }
    __syncthreads();
    if (sink == 1.2345) out[3] = 1;
}

int main() {
    std::vector<double> A(64 * 64), M(64 * 64, 0.0);
    srand(1);
    for (auto& a : A) a = rand() / (double)RAND_MAX - 0.5;
    for (int i = 0; i < 64; ++i)
        for (int j = 0; j < 64; ++j) {
            double s = (i == j) ? 6.4 : 0.0;
            for (int k = 0; k < 64; ++k) s += A[i * 64 + k] * A[j * 64 + k];
            M[i * 64 + j] = s;
        }
    double* dM; long long* dout; int* gflag;
    cudaMalloc(&dM, 64 * 64 * 8); cudaMalloc(&dout, 64); cudaMalloc(&gflag, 64); cudaMemset(gflag, 0, 64);
    cudaMemcpy(dM, M.data(), 64 * 64 * 8, cudaMemcpyHostToDevice);
    const char* names[] = {"others at __syncthreads", "one lane per warp polls global + nanosleep(100)", "others run DMMA products",
                           "one lane per warp polls global, no sleep", "others wait at a named barrier"};
    for (int rolled : {0}) {
        int smem = 3 * CBUF * 8 + 2048;
        cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        for (int threads : {32, 256})
            for (int mode = 0; mode < 5; ++mode) {
                if (threads == 32 && mode > 0) continue;
                long long h[8];
                for (int rep = 0; rep < 3; ++rep) probe<<<1, threads, smem>>>(dM, dout, mode, gflag, rolled);
                cudaMemcpy(h, dout, 64, cudaMemcpyDeviceToHost);
                printf("%s  threads %3d  %-50s factor32 = %6lld  invert32 = %6lld cycles  L[5][3] %016llx X[7][2] %016llx (%s)\n",
                       rolled ? "rolled  " : "unrolled", threads, names[mode], h[0], h[1], (unsigned long long)h[4], (unsigned long long)h[5],
                       cudaGetErrorString(cudaGetLastError()));
            }
    }
    return 0;
}
