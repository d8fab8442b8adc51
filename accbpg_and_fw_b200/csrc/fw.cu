// D_opt_FW / D_opt_FW_away on the device (reference: accbpg/D_opt_alg.py:9-88, :91-185).
//
// The reference loop is O(mn) per iteration: argmax / masked argmin over w, a scalar step rule, a rank-one
// (Sherman-Morrison) update of Hinv = (V X V^T)^{-1}, and one GEMV-T pass over V that refreshes every w_j.
// Here the whole iteration stays on the GPU: five stream-ordered kernels per iteration, the scalar decision is
// taken by the last block of the selection kernel, histories (F, SP, SN, T) are written to device arrays and a
// stop flag turns the remaining launches of a batch into no-ops, so the host only looks at the control block once
// per batch.  F_k: D_opt_FW tracks det by the O(1) update (D_opt_alg.py:52,80); D_opt_FW_away recomputes
// log det(Hinv) by LU every iteration (:136) -- the same quantity follows the determinant-lemma increment, tracked
// here in a compensated (two-sum) accumulator; agreement with the per-iteration LU is ~1e-13 relative over
// thousands of iterations (tests/test_fw_*.py).
// Compiled with -fmad=false: the step-size formulas round as in NumPy.
#include <cooperative_groups.h>
#include "common.cuh"

namespace accbpg {

enum FwCtrl {
    C_STOP = 0,      // 1.0 once the optimality test fired
    C_KSTOP = 1,     // iteration index at which it fired
    C_LOGDET_HI = 2, // log det(V X V^T), compensated accumulator
    C_LOGDET_LO = 3,
    C_WMAX = 4, C_IMAX = 5,
    C_WMIN = 6, C_JMIN = 7,
    C_MODE = 8,      // 0 toward vertex i, 1 away from vertex j
    C_T = 9,         // step
    C_CS = 10,       // signed rank-one coefficient: Hinv <- (Hinv - cs*u u^T)/den
    C_DEN = 11,      // 1 - t (toward) or 1 + t (away)
    C_IDX = 12,      // chosen column (global index)
    C_TSIGN = 13,    // +t (toward) or -t (away): x[idx] += tsign
    C_NITER = 14,    // iterations executed (number of history entries written)
};

constexpr int FW_THREADS = 256;
#define FW_INF (__longlong_as_double(0x7ff0000000000000LL))

__device__ __forceinline__ void fw_arg_combine(double& v, long long& i, double v2, long long i2) {
    if (v2 < v || (v2 == v && i2 < i)) { v = v2; i = i2; }
}
__device__ __forceinline__ void fw_block_argmin(double& v, long long& i, double* shv, long long* shi) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double v2 = __shfl_xor_sync(0xffffffffu, v, o);
        long long i2 = __shfl_xor_sync(0xffffffffu, i, o);
        fw_arg_combine(v, i, v2, i2);
    }
    __syncthreads();
    if (lane == 0) { shv[wid] = v; shi[wid] = i; }
    __syncthreads();
    v = (lane < nw) ? shv[lane] : FW_INF;
    i = (lane < nw) ? shi[lane] : 0x7fffffffffffffffLL;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double v2 = __shfl_xor_sync(0xffffffffu, v, o);
        long long i2 = __shfl_xor_sync(0xffffffffu, i, o);
        fw_arg_combine(v, i, v2, i2);
    }
}

// i = argmax w (first occurrence)            D_opt_alg.py:59 / :145
__global__ void __launch_bounds__(FW_THREADS) fw_argmax_kernel(int64_t n, const double* __restrict__ w,
                                                               double* partials, long long* ipartials,
                                                               unsigned int* counter, double* ctrl) {
    __shared__ double shv[32];
    __shared__ long long shi[32];
    __shared__ bool is_last;
    if (ld_cg(&ctrl[C_STOP]) != 0.0) return;
    double v = FW_INF;
    long long idx = 0x7fffffffffffffffLL;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double wi = -w[i];
        if (wi < v) { v = wi; idx = i; }
    }
    fw_block_argmin(v, idx, shv, shi);
    if (threadIdx.x == 0) { partials[blockIdx.x] = v; ipartials[blockIdx.x] = idx; }
    if (last_block_ticket(counter, &is_last)) {
        v = FW_INF; idx = 0x7fffffffffffffffLL;
        for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x)
            fw_arg_combine(v, idx, ld_cg(&partials[b]), __ldcg(&ipartials[b]));
        fw_block_argmin(v, idx, shv, shi);
        if (threadIdx.x == 0) { ctrl[C_WMAX] = -v; ctrl[C_IMAX] = (double)idx; }
    }
}

// masked argmin, then (last block) the scalar step rule, history entry k and the gather of the chosen column
struct FwDecideParams {
    int64_t n, ldv;
    const double* V;
    const double* x;
    const double* w;
    int m, away, k;
    double eps;
    double* ctrl;
    double* v;           // m doubles: the chosen column of V
    double* hist_F; double* hist_SP; double* hist_SN; double* hist_T;
    double* partials; long long* ipartials; unsigned int* counter;
};

__global__ void __launch_bounds__(FW_THREADS) fw_argmin_decide_kernel(FwDecideParams p) {
    __shared__ double shv[32];
    __shared__ long long shi[32];
    __shared__ bool is_last;
    __shared__ long long sh_idx;
    __shared__ int sh_go;
    if (ld_cg(&p.ctrl[C_STOP]) != 0.0) return;
    const double wmax = ld_cg(&p.ctrl[C_WMAX]);
    double v = FW_INF;
    long long idx = 0x7fffffffffffffffLL;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += stride) {
        double xi = p.x[i], val;
        if (p.away) {
            // j = argmin((w - w[i]) * [x > 1e-8])     D_opt_alg.py:146-147
            val = (xi > 1.0e-8) ? (p.w[i] - wmax) : 0.0;
        } else {
            // j = argmin(w[x > 0])                    D_opt_alg.py:60-61
            if (!(xi > 0.0)) continue;
            val = p.w[i];
        }
        if (val < v) { v = val; idx = i; }
    }
    fw_block_argmin(v, idx, shv, shi);
    if (threadIdx.x == 0) { p.partials[blockIdx.x] = v; p.ipartials[blockIdx.x] = idx; }
    if (!last_block_ticket(p.counter, &is_last)) return;

    v = FW_INF; idx = 0x7fffffffffffffffLL;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x)
        fw_arg_combine(v, idx, ld_cg(&p.partials[b]), __ldcg(&p.ipartials[b]));
    fw_block_argmin(v, idx, shv, shi);
    if (threadIdx.x == 0) {
        double* c = p.ctrl;
        const double md = (double)p.m;
        const long long imax = (long long)c[C_IMAX];
        const long long jmin = (idx < p.n) ? idx : 0;   // empty support cannot happen for x on the simplex
        const double wj = p.w[jmin];
        const double xj = p.x[jmin];
        const double logdet = c[C_LOGDET_HI] + c[C_LOGDET_LO];
        const double eps_pos = wmax / md - 1.0;
        const double eps_neg = 1.0 - wj / md;
        p.hist_F[p.k] = -logdet;                   // D_opt_alg.py:52 (-log det M) / :136 (log det Hinv)
        p.hist_SP[p.k] = eps_pos;
        p.hist_SN[p.k] = eps_neg;
        unsigned long long ns;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
        p.hist_T[p.k] = (double)ns;
        c[C_WMIN] = wj; c[C_JMIN] = (double)jmin;
        c[C_NITER] = (double)(p.k + 1);
        int go = 1;
        if (eps_pos <= p.eps && eps_neg <= p.eps) {   // :72-73 / :159-160
            c[C_STOP] = 1.0; c[C_KSTOP] = (double)p.k;
            go = 0;
        } else {
            double t, cs, den, inc, tsign;
            long long chosen;
            int mode;
            if (!p.away) {                                            // :75-82
                t = (wmax / md - 1.0) / (wmax - 1.0);
                double q = 1.0 + t * (wmax - 1.0);
                cs = t / q; den = 1.0 - t; chosen = imax; mode = 0; tsign = t;
                inc = (md - 1.0) * log(1.0 - t) + log(q);
            } else if (eps_pos >= eps_neg) {                          // :162-170
                t = (wmax / md - 1.0) / (wmax - 1.0);
                double q = 1.0 - t + t * wmax;
                cs = t / q; den = 1.0 - t; chosen = imax; mode = 0; tsign = t;
                inc = (md - 1.0) * log1p(-t) + log(q);
            } else {                                                  // :171-179 away step
                t = fmin((1.0 - wj / md) / (wj - 1.0), xj / (1.0 - xj));
                double q = 1.0 + t - t * wj;
                cs = -(t / q); den = 1.0 + t; chosen = jmin; mode = 1; tsign = -t;
                inc = (md - 1.0) * log1p(t) + log(q);
            }
            // compensated accumulation of log det(V X V^T)
            double hi = c[C_LOGDET_HI];
            double s = hi + inc;
            double bb = s - hi;
            double err = (hi - (s - bb)) + (inc - bb);
            c[C_LOGDET_HI] = s;
            c[C_LOGDET_LO] += err;
            c[C_MODE] = (double)mode; c[C_T] = t; c[C_CS] = cs; c[C_DEN] = den;
            c[C_IDX] = (double)chosen; c[C_TSIGN] = tsign;
            sh_idx = chosen;
        }
        sh_go = go;
    }
    __syncthreads();
    if (sh_go) {
        const long long col = sh_idx;
        for (int r = threadIdx.x; r < p.m; r += blockDim.x) p.v[r] = p.V[(int64_t)r * p.ldv + col];
    }
}

// u = Hinv v (warp per row)            D_opt_alg.py:78 / :165 / :174
__global__ void __launch_bounds__(256) fw_hv_kernel(const double* __restrict__ Hinv, int m, const double* __restrict__ v,
                                                    double* __restrict__ u, const double* ctrl) {
    if (ld_cg(&ctrl[C_STOP]) != 0.0) return;
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= m) return;
    const double* hr = Hinv + (size_t)row * m;
    double s = 0.0;
    for (int c = lane; c < m; c += 32) s += hr[c] * __ldcg(v + c);
    s = warp_sum(s);
    if (lane == 0) u[row] = s;
}

// Hinv <- (Hinv - cs * u u^T) / den     D_opt_alg.py:79 / :166 / :175
__global__ void __launch_bounds__(256) fw_rank1_kernel(double* __restrict__ Hinv, int m, const double* __restrict__ u,
                                                       const double* ctrl) {
    if (ld_cg(&ctrl[C_STOP]) != 0.0) return;
    const double cs = ld_cg(&ctrl[C_CS]), den = ld_cg(&ctrl[C_DEN]);
    const int64_t total = (int64_t)m * m;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        int r = (int)(e / m), c = (int)(e - (int64_t)r * m);
        double o = __ldcg(u + r) * __ldcg(u + c);
        Hinv[e] = (Hinv[e] - cs * o) / den;
    }
}

// the one pass over V:  p_j = u^T v_j ;  w_j <- (w_j - cs p_j^2)/den ;  x_j <- x_j*den (+- t at the chosen column)
constexpr int FWP_THREADS = 64;      // 2 columns per thread -> 128 columns per CTA
constexpr int FWP_UNROLL = 16;       // rows in flight per thread (16 x 16 B)
constexpr int FWP_MAXSPLIT = 4;      // CTAs of one cluster split the rows of V; partials meet in distributed smem

template <bool VEC2>
__global__ void __launch_bounds__(FWP_THREADS) fw_pass_kernel(const double* __restrict__ V, int m, int64_t n, int64_t ldv,
                                                              const double* __restrict__ u, double* __restrict__ x,
                                                              double* __restrict__ w, const double* ctrl) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ double us[];                         // this CTA's slice of u
    __shared__ double part[2 * FWP_THREADS];
    if (ld_cg(&ctrl[C_STOP]) != 0.0) return;               // uniform over the whole grid
    const unsigned nsplit = cluster.num_blocks(), rank = cluster.block_rank();
    const int rows_per = (m + nsplit - 1) / nsplit;
    const int r0 = rank * rows_per;
    const int r1 = min(m, r0 + rows_per);
    for (int r = r0 + threadIdx.x; r < r1; r += blockDim.x) us[r - r0] = __ldcg(u + r);
    __syncthreads();
    const int64_t j = ((int64_t)blockIdx.x * FWP_THREADS + threadIdx.x) * 2;
    double s0 = 0.0, s1 = 0.0;
    if (j < n) {
        const double* col = V + j;
        int r = r0;
        if (VEC2 && j + 1 < n) {
            for (; r + FWP_UNROLL <= r1; r += FWP_UNROLL) {
                double2 a[FWP_UNROLL];
#pragma unroll
                for (int q = 0; q < FWP_UNROLL; ++q)
                    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];"
                                 : "=d"(a[q].x), "=d"(a[q].y) : "l"(col + (int64_t)(r + q) * ldv));
#pragma unroll
                for (int q = 0; q < FWP_UNROLL; ++q) {
                    double uq = us[r - r0 + q];
                    s0 += uq * a[q].x;
                    s1 += uq * a[q].y;
                }
            }
            for (; r < r1; ++r) {
                double2 a;
                asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];"
                             : "=d"(a.x), "=d"(a.y) : "l"(col + (int64_t)r * ldv));
                double uq = us[r - r0];
                s0 += uq * a.x;
                s1 += uq * a.y;
            }
        } else {
            const bool two = (j + 1 < n);
            for (; r < r1; ++r) {
                double uq = us[r - r0];
                s0 += uq * __ldcs(col + (int64_t)r * ldv);
                if (two) s1 += uq * __ldcs(col + (int64_t)r * ldv + 1);
            }
        }
    }
    part[2 * threadIdx.x] = s0;
    part[2 * threadIdx.x + 1] = s1;
    cluster.sync();
    if (rank == 0 && j < n) {
        for (unsigned q = 1; q < nsplit; ++q) {             // fixed order: row chunks 0, 1, 2, ...
            const double* rp = cluster.map_shared_rank(part, q);
            s0 += rp[2 * threadIdx.x];
            s1 += rp[2 * threadIdx.x + 1];
        }
        const double cs = ld_cg(&ctrl[C_CS]), den = ld_cg(&ctrl[C_DEN]);
        const int64_t idx = (int64_t)ld_cg(&ctrl[C_IDX]);
        const double tsign = ld_cg(&ctrl[C_TSIGN]);
        w[j] = (w[j] - cs * (s0 * s0)) / den;
        double xn = x[j] * den;
        if (j == idx) xn = xn + tsign;
        x[j] = xn;
        if (j + 1 < n) {
            w[j + 1] = (w[j + 1] - cs * (s1 * s1)) / den;
            xn = x[j + 1] * den;
            if (j + 1 == idx) xn = xn + tsign;
            x[j + 1] = xn;
        }
    }
    cluster.sync();                                         // remote shared memory must outlive rank 0's reads
}

// Hinv = Linv^T Linv  (setup only; Linv lower triangular, zero padded, leading dimension mp)
__global__ void __launch_bounds__(256) fw_hinv_kernel(const double* __restrict__ Linv, int m, int mp,
                                                      double* __restrict__ Hinv) {
    __shared__ double Ar[32][33], Ac[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;       // 32 x 8
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    if (c0 > r0) return;                                          // lower blocks only, mirrored below
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int k0 = r0; k0 < mp; k0 += 32) {                        // Linv[k][r] = 0 for k < r
        __syncthreads();
        for (int e = ty; e < 32; e += 8) {
            Ar[e][tx] = Linv[(size_t)(k0 + e) * mp + r0 + tx];
            Ac[e][tx] = Linv[(size_t)(k0 + e) * mp + c0 + tx];
        }
        __syncthreads();
#pragma unroll 8
        for (int kk = 0; kk < 32; ++kk) {
            double b = Ac[kk][tx];
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[q] = fma(Ar[kk][ty + 8 * q], b, acc[q]);
        }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        int r = r0 + ty + 8 * q, c = c0 + tx;
        if (r < m && c < m) {
            Hinv[(size_t)r * m + c] = acc[q];
            Hinv[(size_t)c * m + r] = acc[q];
        }
    }
}

__global__ void fw_ctrl_init_from_slot_kernel(double* ctrl, const double* neg_logdet_slot) {
    for (int i = threadIdx.x; i < ACCBPG_FW_CTRL_DOUBLES; i += blockDim.x) ctrl[i] = 0.0;
    __syncthreads();
    if (threadIdx.x == 0) ctrl[C_LOGDET_HI] = -neg_logdet_slot[0];
}
__global__ void __launch_bounds__(256) negate_kernel(int64_t n, double* x) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) x[i] = -x[i];
}

}  // namespace accbpg

using namespace accbpg;

extern "C" {

size_t accbpg_fw_workspace_bytes(int m, int64_t n_local) {
    // dopt workspace (gram / factor / Linv / gradient partials) + v and u vectors
    size_t base = accbpg_dopt_workspace_bytes(m, n_local);
    return base + 2 * (((size_t)m * 8 + 255) / 256 * 256);
}

// D_opt_alg.py:39-45 / :123-129:  M = V diag(x0) V^T, Hinv = M^{-1}, w_j = v_j^T Hinv v_j, log det M
int accbpg_fw_setup(void* ctx, void* stream, const double* V, int m, int64_t n, int64_t ldv, const double* x0,
                    void* ws, double* Hinv, double* w, double* ctrl) {
    Ctx* c = (Ctx*)ctx;
    cudaStream_t s = (cudaStream_t)stream;
    if (!c || !V || !x0 || !ws || !Hinv || !w || !ctrl) return arg_err("fw_setup: NULL pointer");
    double* slot = c->d_slots + 247;
    // func_grad(flag=2): slot <- -log det M, w <- gradient = -(v_j^T Hinv v_j); Linv stays in the workspace
    int rc = accbpg_dopt_func_grad(ctx, stream, V, m, n, ldv, x0, 2, ws, slot, w);
    if (rc) return rc;
    int g = grid_for(c, n, 256, 2, 8);
    negate_kernel<<<g, 256, 0, s>>>(n, w);
    ACCBPG_LAUNCHED("negate_kernel");
    fw_ctrl_init_from_slot_kernel<<<1, 32, 0, s>>>(ctrl, slot);
    ACCBPG_LAUNCHED("fw_ctrl_init");
    // Linv sits at a fixed offset of the dopt workspace: recover it through the same plan
    int mp = 0;
    size_t off = dopt_linv_offset(m, n, c->sm_count, &mp);
    const double* Linv = (const double*)((const char*)ws + off);
    dim3 grid((m + 31) / 32, (m + 31) / 32);
    fw_hinv_kernel<<<grid, 256, 0, s>>>(Linv, m, mp, Hinv);
    ACCBPG_LAUNCHED("fw_hinv_kernel");
    return ACCBPG_OK;
}

// run iterations k_start .. k_start+k_count-1 (no-ops after the stop flag is raised)
int accbpg_fw_run(void* ctx, void* stream, const double* V, int m, int64_t n, int64_t ldv, int away, double eps,
                  int k_start, int k_count, void* ws, double* Hinv, double* x, double* w, double* ctrl,
                  double* hist_F, double* hist_SP, double* hist_SN, double* hist_T) {
    Ctx* c = (Ctx*)ctx;
    cudaStream_t s = (cudaStream_t)stream;
    if (!c || !V || !ws || !Hinv || !x || !w || !ctrl || !hist_F || !hist_SP || !hist_SN || !hist_T)
        return arg_err("fw_run: NULL pointer");
    if (m < 1 || n < 1 || ldv < n || k_count < 0) return arg_err("fw_run: shape");
    size_t base = accbpg_dopt_workspace_bytes(m, n);
    double* v = (double*)((char*)ws + base);
    double* u = v + ((size_t)m + 31) / 32 * 32;
    const int sel_grid = grid_for(c, n, FW_THREADS, 4, 4);
    const int hv_grid = (m + 7) / 8;
    const int r1_grid = grid_for(c, (int64_t)m * m, 256, 2, 8);
    const int64_t pass_grid = (n + 2 * FWP_THREADS - 1) / (2 * FWP_THREADS);
    if (pass_grid > 2147483647LL) return arg_err("fw_run: n too large");
    // split the rows over a small cluster when the column blocks alone cannot fill the GPU evenly
    int nsplit = 1;
    while (nsplit < FWP_MAXSPLIT && pass_grid * nsplit < (int64_t)c->sm_count * 8 && m / (nsplit * 2) >= 64) nsplit *= 2;
    const size_t pass_smem = (size_t)((m + nsplit - 1) / nsplit) * sizeof(double);
    if (pass_smem > 200 * 1024) return arg_err("fw_run: m too large for the shared-memory copy of u");
    const bool pass_vec = ((reinterpret_cast<uintptr_t>(V) & 15u) == 0) && (ldv % 2 == 0);
    auto pass_fn = pass_vec ? fw_pass_kernel<true> : fw_pass_kernel<false>;
    if (pass_smem > 48 * 1024)
        ACCBPG_CUDA(cudaFuncSetAttribute(pass_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pass_smem));
    cudaLaunchConfig_t pass_cfg = {};
    pass_cfg.gridDim = dim3((unsigned)pass_grid, (unsigned)nsplit, 1);
    pass_cfg.blockDim = dim3(FWP_THREADS, 1, 1);
    pass_cfg.dynamicSmemBytes = pass_smem;
    pass_cfg.stream = s;
    cudaLaunchAttribute pass_attr[1];
    pass_attr[0].id = cudaLaunchAttributeClusterDimension;
    pass_attr[0].val.clusterDim.x = 1;
    pass_attr[0].val.clusterDim.y = (unsigned)nsplit;
    pass_attr[0].val.clusterDim.z = 1;
    pass_cfg.attrs = pass_attr;
    pass_cfg.numAttrs = 1;
    FwDecideParams p;
    p.n = n; p.ldv = ldv; p.V = V; p.x = x; p.w = w; p.m = m; p.away = away; p.eps = eps;
    p.ctrl = ctrl; p.v = v; p.hist_F = hist_F; p.hist_SP = hist_SP; p.hist_SN = hist_SN; p.hist_T = hist_T;
    p.partials = c->d_partials; p.ipartials = c->d_ipartials; p.counter = c->d_counter;
    for (int k = k_start; k < k_start + k_count; ++k) {
        ProfScope ps_iter(P_FW_ITER, s);
        fw_argmax_kernel<<<sel_grid, FW_THREADS, 0, s>>>(n, w, c->d_partials, c->d_ipartials, c->d_counter, ctrl);
        ACCBPG_LAUNCHED("fw_argmax_kernel");
        p.k = k;
        fw_argmin_decide_kernel<<<sel_grid, FW_THREADS, 0, s>>>(p);
        ACCBPG_LAUNCHED("fw_argmin_decide_kernel");
        fw_hv_kernel<<<hv_grid, 256, 0, s>>>(Hinv, m, v, u, ctrl);
        ACCBPG_LAUNCHED("fw_hv_kernel");
        fw_rank1_kernel<<<r1_grid, 256, 0, s>>>(Hinv, m, u, ctrl);
        ACCBPG_LAUNCHED("fw_rank1_kernel");
        {
            ProfScope ps(P_FW_PASS, s);
            ACCBPG_CUDA(cudaLaunchKernelEx(&pass_cfg, pass_fn, V, m, n, ldv, (const double*)u, x, w, (const double*)ctrl));
        }
        ACCBPG_LAUNCHED("fw_pass_kernel");
    }
    return ACCBPG_OK;
}

}  // extern "C"
