// Poisson / KL regression oracles (PoissonRegression.func_grad, KLdivRegression.func_grad,
// accbpg/functions.py:102-120, :140-158):   Ax,  f(Ax, b),  r(Ax, b),  g = A^T r.
// Both matrix passes are HBM-bound: every element of A is read exactly once per pass with 128-bit loads,
// x / r are re-used from registers (4 rows per CTA) or L1, partial results are combined in a fixed order.
// Compiled with -fmad=false so the scalar objective terms round as NumPy's do.
#include "common.cuh"

namespace accbpg {

constexpr int MV_THREADS = 256;
constexpr int MV_ROWS = 4;             // rows per CTA in A*x: x is loaded once per 4 rows of A
constexpr int64_t MV_SEG = 1 << 16;    // columns per CTA segment (64 Ki doubles = 512 KiB per row)

__device__ __forceinline__ double2 ld_stream2(const double* p) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ double ld_stream1(const double* p) {
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

// y[i] (or partial[seg][i]) = sum_j A[i][j] x[j] over this CTA's column segment
template <bool VEC2>
__global__ void __launch_bounds__(MV_THREADS) matvec_kernel(const double* __restrict__ A, int64_t m, int64_t n,
                                                            int64_t lda, const double* __restrict__ x,
                                                            double* __restrict__ out, int64_t out_stride) {
    __shared__ double sh[32];
    const int64_t row0 = (int64_t)blockIdx.x * MV_ROWS;
    const int64_t c0 = (int64_t)blockIdx.y * MV_SEG;
    int64_t c1 = c0 + MV_SEG;
    if (c1 > n) c1 = n;
    const double* rowp[MV_ROWS];
#pragma unroll
    for (int r = 0; r < MV_ROWS; ++r) {
        int64_t row = row0 + r;
        if (row >= m) row = m - 1;          // clamp: result discarded below
        rowp[r] = A + row * lda;
    }
    double acc[MV_ROWS];
#pragma unroll
    for (int r = 0; r < MV_ROWS; ++r) acc[r] = 0.0;
    if (VEC2) {
        int64_t j = c0 + 2 * (int64_t)threadIdx.x;
        // main loop: two 16-byte column chunks per row in flight per thread
        for (; j + 2 * MV_THREADS + 1 < c1; j += 4 * MV_THREADS) {
            double2 x0 = *reinterpret_cast<const double2*>(x + j);
            double2 x1 = *reinterpret_cast<const double2*>(x + j + 2 * MV_THREADS);
            double2 a0[MV_ROWS], a1[MV_ROWS];
#pragma unroll
            for (int r = 0; r < MV_ROWS; ++r) {
                a0[r] = ld_stream2(rowp[r] + j);
                a1[r] = ld_stream2(rowp[r] + j + 2 * MV_THREADS);
            }
#pragma unroll
            for (int r = 0; r < MV_ROWS; ++r) {
                acc[r] += a0[r].x * x0.x;
                acc[r] += a0[r].y * x0.y;
                acc[r] += a1[r].x * x1.x;
                acc[r] += a1[r].y * x1.y;
            }
        }
        for (; j < c1; j += 2 * MV_THREADS) {
            if (j + 1 < c1) {
                double2 x0 = *reinterpret_cast<const double2*>(x + j);
#pragma unroll
                for (int r = 0; r < MV_ROWS; ++r) {
                    double2 a = ld_stream2(rowp[r] + j);
                    acc[r] += a.x * x0.x;
                    acc[r] += a.y * x0.y;
                }
            } else {
                double xs = x[j];
#pragma unroll
                for (int r = 0; r < MV_ROWS; ++r) acc[r] += ld_stream1(rowp[r] + j) * xs;
            }
        }
    } else {
        for (int64_t j = c0 + threadIdx.x; j < c1; j += MV_THREADS) {
            double xs = x[j];
#pragma unroll
            for (int r = 0; r < MV_ROWS; ++r) acc[r] += ld_stream1(rowp[r] + j) * xs;
        }
    }
#pragma unroll
    for (int r = 0; r < MV_ROWS; ++r) {
        double s = block_sum(acc[r], sh);
        if (threadIdx.x == 0 && row0 + r < m) out[(int64_t)blockIdx.y * out_stride + row0 + r] = s;
    }
}

// out[i] = sum_s partial[s][i], segments added in index order
__global__ void __launch_bounds__(256) seg_reduce_kernel(const double* partial, int nseg, int64_t stride, int64_t len,
                                                         double* out, double scale) {
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += step) {
        double s = 0.0;
        for (int k = 0; k < nseg; ++k) s += __ldcg(partial + (int64_t)k * stride + i);
        out[i] = s * scale;
    }
}

// objective value and residual (functions.py:106,110,119 / :144,148,157)
__global__ void __launch_bounds__(256) value_resid_kernel(int kind, int64_t m, const double* __restrict__ Ax,
                                                          const double* __restrict__ b, double* __restrict__ r,
                                                          double* partials, unsigned int* counter, double* f_out) {
    __shared__ double sh[32];
    __shared__ bool is_last;
    double acc = 0.0;
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += step) {
        double a = Ax[i], bi = b[i];
        if (kind == ACCBPG_LINREG_POISSON) {
            double q = bi / a;
            acc += (bi * log(q) + a) - bi;
            if (r) r[i] = 1.0 - q;
        } else {
            double lg = log(a / bi);
            acc += (a * lg - a) + bi;
            if (r) r[i] = lg;
        }
    }
    acc = block_sum(acc, sh);
    if (threadIdx.x == 0) partials[blockIdx.x] = acc;
    if (last_block_ticket(counter, &is_last)) {
        double s = 0.0;
        for (int k = threadIdx.x; k < (int)gridDim.x; k += blockDim.x) s += ld_cg(&partials[k]);
        s = block_sum(s, sh);
        if (threadIdx.x == 0 && f_out) f_out[0] = s;
    }
}

// partial[chunk][j] = sum over this chunk's rows of r[i] * A[i][j]; thread owns 2 adjacent columns
constexpr int RMV_THREADS = 256;
constexpr int RMV_UNROLL = 8;

template <bool VEC2>
__global__ void __launch_bounds__(RMV_THREADS) rmatvec_kernel(const double* __restrict__ A, int64_t m, int64_t n,
                                                              int64_t lda, const double* __restrict__ r,
                                                              int64_t rows_per_chunk, double* __restrict__ partial,
                                                              int64_t pstride) {
    const int64_t i0 = (int64_t)blockIdx.y * rows_per_chunk;
    int64_t i1 = i0 + rows_per_chunk;
    if (i1 > m) i1 = m;
    if (VEC2) {
        const int64_t j = ((int64_t)blockIdx.x * RMV_THREADS + threadIdx.x) * 2;
        if (j >= n) return;
        const bool pair = (j + 1 < n);
        double s0 = 0.0, s1 = 0.0;
        const double* col = A + j;
        int64_t i = i0;
        if (pair) {
            for (; i + RMV_UNROLL <= i1; i += RMV_UNROLL) {
                double2 a[RMV_UNROLL];
                double rv[RMV_UNROLL];
#pragma unroll
                for (int u = 0; u < RMV_UNROLL; ++u) {
                    a[u] = ld_stream2(col + (i + u) * lda);
                    rv[u] = __ldg(r + i + u);
                }
#pragma unroll
                for (int u = 0; u < RMV_UNROLL; ++u) {
                    s0 += rv[u] * a[u].x;
                    s1 += rv[u] * a[u].y;
                }
            }
            for (; i < i1; ++i) {
                double2 a = ld_stream2(col + i * lda);
                double rv = __ldg(r + i);
                s0 += rv * a.x;
                s1 += rv * a.y;
            }
            *reinterpret_cast<double2*>(partial + (int64_t)blockIdx.y * pstride + j) = make_double2(s0, s1);
        } else {
            for (; i < i1; ++i) s0 += __ldg(r + i) * ld_stream1(col + i * lda);
            partial[(int64_t)blockIdx.y * pstride + j] = s0;
        }
    } else {
        const int64_t j = (int64_t)blockIdx.x * RMV_THREADS * 2 + threadIdx.x;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int64_t jj = j + h * RMV_THREADS;
            if (jj >= n) continue;
            double s0 = 0.0;
            for (int64_t i = i0; i < i1; ++i) s0 += __ldg(r + i) * ld_stream1(A + i * lda + jj);
            partial[(int64_t)blockIdx.y * pstride + jj] = s0;
        }
    }
}

struct LinregPlan {
    int nseg;            // column segments in A*x
    int nchunk;          // row chunks in A^T r
    int64_t rows_per_chunk, pstride_mv, pstride_rmv;
    size_t off_mv, off_rmv, total;
};

static LinregPlan linreg_plan(int64_t m, int64_t n, int sm_count) {
    LinregPlan pl;
    pl.nseg = (int)((n + MV_SEG - 1) / MV_SEG);
    if (pl.nseg < 1) pl.nseg = 1;
    int64_t colblocks = (n + 2 * RMV_THREADS - 1) / (2 * RMV_THREADS);
    int64_t want = ((int64_t)sm_count * 16 + colblocks - 1) / colblocks;     // ~16 CTAs per SM in total
    int64_t maxchunk = (m + 63) / 64;                                       // at least 64 rows per chunk
    if (want > maxchunk) want = maxchunk;
    if (want < 1) want = 1;
    if (want > 256) want = 256;
    pl.nchunk = (int)want;
    pl.rows_per_chunk = (m + pl.nchunk - 1) / pl.nchunk;
    pl.pstride_mv = (m + 1) / 2 * 2;
    pl.pstride_rmv = (n + 1) / 2 * 2;
    size_t a = 0;
    pl.off_mv = a;  a += ((size_t)pl.nseg * pl.pstride_mv * 8 + 255) / 256 * 256;
    pl.off_rmv = a; a += ((size_t)pl.nchunk * pl.pstride_rmv * 8 + 255) / 256 * 256;
    pl.total = a;
    return pl;
}

static int device_sms() {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}
static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace accbpg

using namespace accbpg;

extern "C" {

size_t accbpg_linreg_workspace_bytes(int64_t m, int64_t n_local) {
    if (m < 1 || n_local < 1) return 0;
    return linreg_plan(m, n_local, device_sms()).total;
}

// the workspace is only touched when n_local > 65536 (several column segments per row)
int accbpg_linreg_matvec(void* ctx, void* stream, const double* A, int64_t m, int64_t n, int64_t lda,
                         const double* x, void* ws, double* Ax) {
    Ctx* c = (Ctx*)ctx;
    cudaStream_t s = (cudaStream_t)stream;
    if (!c || !A || !x || !Ax) return arg_err("linreg_matvec: NULL pointer");
    ACCBPG_ON_DEVICE(c);
    if (m < 1 || n < 1 || lda < n) return arg_err("linreg_matvec: shape");
    LinregPlan pl = linreg_plan(m, n, c->sm_count);
    if (pl.nseg > 1 && !ws) return arg_err("linreg_matvec: workspace required when n_local > 65536");
    int64_t rowblocks = (m + MV_ROWS - 1) / MV_ROWS;
    if (rowblocks > 2147483647LL) return arg_err("linreg_matvec: m too large");
    dim3 grid((unsigned)rowblocks, (unsigned)pl.nseg);
    bool vec = al16(A) && al16(x) && (lda % 2 == 0);
    double* out = (pl.nseg > 1) ? (double*)((char*)ws + pl.off_mv) : Ax;
    {
        ProfScope ps(P_MATVEC, s);
        if (vec) matvec_kernel<true><<<grid, MV_THREADS, 0, s>>>(A, m, n, lda, x, out, pl.pstride_mv);
        else     matvec_kernel<false><<<grid, MV_THREADS, 0, s>>>(A, m, n, lda, x, out, pl.pstride_mv);
    }
    ACCBPG_LAUNCHED("matvec_kernel");
    if (pl.nseg > 1) {
        int g = grid_for(c, m, 256, 2, 8);
        seg_reduce_kernel<<<g, 256, 0, s>>>(out, pl.nseg, pl.pstride_mv, m, Ax, 1.0);
        ACCBPG_LAUNCHED("seg_reduce_kernel");
    }
    return ACCBPG_OK;
}

int accbpg_linreg_value_resid(void* ctx, void* stream, int kind, int64_t m, const double* Ax, const double* b,
                              double* d_f_out, double* r) {
    Ctx* c = (Ctx*)ctx;
    cudaStream_t s = (cudaStream_t)stream;
    if (!c || !Ax || !b) return arg_err("linreg_value_resid: NULL pointer");
    ACCBPG_ON_DEVICE(c);
    if (kind != ACCBPG_LINREG_POISSON && kind != ACCBPG_LINREG_KL) return arg_err("linreg kind");
    if (m < 1) return arg_err("linreg_value_resid: m");
    int g = grid_for(c, m, 256, 4, 4);
    value_resid_kernel<<<g, 256, 0, s>>>(kind, m, Ax, b, r, c->d_partials, c->d_counter, d_f_out);
    ACCBPG_LAUNCHED("value_resid_kernel");
    return ACCBPG_OK;
}

int accbpg_linreg_rmatvec(void* ctx, void* stream, const double* A, int64_t m, int64_t n, int64_t lda,
                          const double* r, void* ws, double* g) {
    Ctx* c = (Ctx*)ctx;
    cudaStream_t s = (cudaStream_t)stream;
    if (!c || !A || !r || !ws || !g) return arg_err("linreg_rmatvec: NULL pointer");
    ACCBPG_ON_DEVICE(c);
    if (m < 1 || n < 1 || lda < n) return arg_err("linreg_rmatvec: shape");
    LinregPlan pl = linreg_plan(m, n, c->sm_count);
    double* partial = (double*)((char*)ws + pl.off_rmv);
    int64_t colblocks = (n + 2 * RMV_THREADS - 1) / (2 * RMV_THREADS);
    if (colblocks > 2147483647LL) return arg_err("linreg_rmatvec: n too large");
    dim3 grid((unsigned)colblocks, (unsigned)pl.nchunk);
    bool vec = al16(A) && (lda % 2 == 0);
    {
        ProfScope ps(P_RMATVEC, s);
        if (vec) rmatvec_kernel<true><<<grid, RMV_THREADS, 0, s>>>(A, m, n, lda, r, pl.rows_per_chunk, partial, pl.pstride_rmv);
        else     rmatvec_kernel<false><<<grid, RMV_THREADS, 0, s>>>(A, m, n, lda, r, pl.rows_per_chunk, partial, pl.pstride_rmv);
    }
    ACCBPG_LAUNCHED("rmatvec_kernel");
    int fg = grid_for(c, n, 256, 2, 8);
    seg_reduce_kernel<<<fg, 256, 0, s>>>(partial, pl.nchunk, pl.pstride_rmv, n, g, 1.0);
    ACCBPG_LAUNCHED("seg_reduce_kernel");
    return ACCBPG_OK;
}

}  // extern "C"
