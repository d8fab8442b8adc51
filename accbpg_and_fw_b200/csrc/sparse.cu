// D-optimal design objective on a SPARSE design matrix (LIBSVM data: accbpg/applications.py:17-33 densifies X with
// .toarray('C'); utils.py:22-95 parses it).  H (m x n, m = features, n = samples) is kept in compressed-column form - one
// column of H per sample, exactly the order of the LIBSVM file - and never densified in HBM:
//   K1  M = H diag(x) H^T   column tiles are scattered into a dense m x 64 tile in shared memory, every thread owns a fixed
//                            set of entries of the lower triangle and adds the tile's contribution in column order; CTA
//                            partials are summed in CTA order (no floating-point atomics: bit-reproducible)
//   K4  g_j = -|| L^{-1} h_j ||^2   warp per column over the nonzeros of h_j, L^{-1} (left in the workspace by
//                            accbpg_dopt_factor) staged in shared memory
// K2 / K3 (Cholesky, L^{-1}) are the dense m x m chain of chol.cu.  m <= 128 (the LIBSVM regression sets have m <= 14).
#include "dmma.cuh"

namespace accbpg {

constexpr int SP_TC = 64;            // columns per tile
constexpr int SP_THREADS = 256;
constexpr int SP_MAX_M = 128;
constexpr int SP_GRID = 296;         // CTA partials of the Gram matrix

__global__ void __launch_bounds__(SP_THREADS) sparse_gram_kernel(const int64_t* __restrict__ colptr, const int* __restrict__ rowidx,
                                                                 const double* __restrict__ vals, int m, int64_t n,
                                                                 const double* __restrict__ x, double* __restrict__ partials,
                                                                 uint32_t* status) {
    extern __shared__ __align__(16) double sp_sm[];
    const int T = m * (m + 1) / 2;
    double* acc = sp_sm;                         // [T] lower triangle, row-major packed
    double* D = acc + T;                         // [m][SP_TC + 1]
    double* xs = D + (size_t)m * (SP_TC + 1);    // [SP_TC]
    const int tid = threadIdx.x;
    for (int e = tid; e < T; e += SP_THREADS) acc[e] = 0.0;
    const int64_t ntiles = (n + SP_TC - 1) / SP_TC;
    bool neg = false;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t j0 = tile * SP_TC;
        const int tc = (int)((n - j0) < SP_TC ? (n - j0) : SP_TC);
        __syncthreads();
        for (int e = tid; e < m * (SP_TC + 1); e += SP_THREADS) D[e] = 0.0;
        if (tid < SP_TC) {
            double xv = 0.0;
            if (tid < tc) { xv = x[j0 + tid]; if (xv < 0.0) neg = true; }
            xs[tid] = xv;
        }
        __syncthreads();
        const int64_t p0 = colptr[j0], p1 = colptr[j0 + tc];
        // scatter the tile's nonzeros: the column of an entry by a search over the (<= 65) column pointers of the tile
        for (int64_t q = p0 + tid; q < p1; q += SP_THREADS) {
            int lo = 0, hi = tc;                 // largest c with colptr[j0 + c] <= q
            while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (colptr[j0 + mid] <= q) lo = mid; else hi = mid; }
            const int r = rowidx[q];
            if (r >= 0 && r < m) D[r * (SP_TC + 1) + lo] = vals[q];
        }
        __syncthreads();
        for (int e = tid; e < T; e += SP_THREADS) {
            int a = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
            while ((a + 1) * (a + 2) / 2 <= e) ++a;
            while (a * (a + 1) / 2 > e) --a;
            const int b = e - a * (a + 1) / 2;
            const double* da = D + a * (SP_TC + 1);
            const double* db = D + b * (SP_TC + 1);
            double s = 0.0;
            for (int c = 0; c < tc; ++c) s = fma(da[c] * xs[c], db[c], s);
            acc[e] += s;
        }
    }
    __syncthreads();
    for (int e = tid; e < T; e += SP_THREADS) partials[(size_t)blockIdx.x * T + e] = acc[e];
    if (neg) atomicOr(status, ACCBPG_ST_X_NEGATIVE);
}

__global__ void __launch_bounds__(256) sparse_gram_reduce_kernel(const double* __restrict__ partials, int nparts, int m,
                                                                 double* __restrict__ M) {
    const int T = m * (m + 1) / 2;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < T; e += gridDim.x * blockDim.x) {
        int a = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
        while ((a + 1) * (a + 2) / 2 <= e) ++a;
        while (a * (a + 1) / 2 > e) --a;
        const int b = e - a * (a + 1) / 2;
        double s = 0.0;
        for (int q = 0; q < nparts; ++q) s += __ldcg(partials + (size_t)q * T + e);
        M[(size_t)a * m + b] = s;
        M[(size_t)b * m + a] = s;
    }
}

__global__ void __launch_bounds__(SP_THREADS) sparse_grad_kernel(const int64_t* __restrict__ colptr, const int* __restrict__ rowidx,
                                                                 const double* __restrict__ vals, int m, int mp, int64_t n,
                                                                 const double* __restrict__ Linv, double* __restrict__ g) {
    extern __shared__ __align__(16) double sp_sm[];
    double* Ls = sp_sm;                          // [m][m + 1], lower triangle of L^{-1}
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int e = tid; e < m * m; e += SP_THREADS) {
        const int i = e / m, a = e - i * m;
        Ls[i * (m + 1) + a] = (a <= i) ? __ldcg(Linv + (size_t)i * mp + a) : 0.0;
    }
    __syncthreads();
    const int64_t wstride = (int64_t)gridDim.x * (SP_THREADS / 32);
    for (int64_t j = (int64_t)blockIdx.x * (SP_THREADS / 32) + wid; j < n; j += wstride) {
        const int64_t p0 = colptr[j], p1 = colptr[j + 1];
        double accv = 0.0;
        for (int i = lane; i < m; i += 32) {
            const double* li = Ls + i * (m + 1);
            double s = 0.0;
            for (int64_t q = p0; q < p1; ++q) {
                const int a = rowidx[q];
                if (a >= 0 && a <= i) s = fma(li[a], vals[q], s);
            }
            accv = fma(s, s, accv);
        }
        accv = warp_sum(accv);
        if (lane == 0) g[j] = -accv;
    }
}

static size_t sp_head_bytes(int m) { return (accbpg_dopt_workspace_bytes(m, 2) + 255) / 256 * 256; }

static int sp_ensure_attrs(const Ctx* c) {
    static bool attr_done[kMaxDevices] = {};
    if (c->device >= 0 && c->device < kMaxDevices && attr_done[c->device]) return ACCBPG_OK;
    ACCBPG_CUDA(cudaFuncSetAttribute(sparse_gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    ACCBPG_CUDA(cudaFuncSetAttribute(sparse_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    if (c->device >= 0 && c->device < kMaxDevices) attr_done[c->device] = true;
    return ACCBPG_OK;
}

}  // namespace accbpg

using namespace accbpg;

extern "C" {

size_t accbpg_dopt_sparse_workspace_bytes(int m) {
    if (m < 1 || m > SP_MAX_M) return 0;
    return sp_head_bytes(m) + (size_t)SP_GRID * (m * (m + 1) / 2) * sizeof(double) + 256;
}

int accbpg_dopt_sparse_gram(void* ctx, void* stream, const int64_t* colptr, const int* rowidx, const double* vals, int m,
                            int64_t n, const double* x, void* ws, double* M) {
    Ctx* c = (Ctx*)ctx;
    cudaStream_t s = (cudaStream_t)stream;
    if (!c || !colptr || !rowidx || !vals || !x || !ws || !M) return arg_err("dopt_sparse_gram: NULL pointer");
    ACCBPG_ON_DEVICE(c);
    if (m < 1 || m > SP_MAX_M || n < 1) return arg_err("dopt_sparse_gram: shape (m <= 128)");
    const int T = m * (m + 1) / 2;
    double* partials = (double*)((char*)ws + sp_head_bytes(m));
    const int64_t ntiles = (n + SP_TC - 1) / SP_TC;
    int grid = (int)(ntiles < SP_GRID ? ntiles : SP_GRID);
    if (grid > 2 * c->sm_count) grid = 2 * c->sm_count;
    const size_t smem = ((size_t)T + (size_t)m * (SP_TC + 1) + SP_TC) * sizeof(double);
    int rc = sp_ensure_attrs(c);
    if (rc) return rc;
    sparse_gram_kernel<<<grid, SP_THREADS, smem, s>>>(colptr, rowidx, vals, m, n, x, partials, c->d_status);
    ACCBPG_LAUNCHED("sparse_gram_kernel");
    sparse_gram_reduce_kernel<<<(T + 255) / 256, 256, 0, s>>>(partials, grid, m, M);
    ACCBPG_LAUNCHED("sparse_gram_reduce_kernel");
    return ACCBPG_OK;
}

int accbpg_dopt_sparse_grad(void* ctx, void* stream, const int64_t* colptr, const int* rowidx, const double* vals, int m,
                            int64_t n, void* ws, double* g) {
    Ctx* c = (Ctx*)ctx;
    cudaStream_t s = (cudaStream_t)stream;
    if (!c || !colptr || !rowidx || !vals || !ws || !g) return arg_err("dopt_sparse_grad: NULL pointer");
    ACCBPG_ON_DEVICE(c);
    if (m < 1 || m > SP_MAX_M || n < 1) return arg_err("dopt_sparse_grad: shape (m <= 128)");
    int mp = 0;
    const size_t off = dopt_linv_offset(m, 2, c->sm_count, &mp);
    const double* Linv = (const double*)((const char*)ws + off);
    const size_t smem = (size_t)m * (m + 1) * sizeof(double);
    int rc = sp_ensure_attrs(c);
    if (rc) return rc;
    int64_t want = (n + SP_THREADS / 32 - 1) / (SP_THREADS / 32);
    int grid = (int)(want < 4 * c->sm_count ? want : 4 * c->sm_count);
    sparse_grad_kernel<<<grid, SP_THREADS, smem, s>>>(colptr, rowidx, vals, m, mp, n, Linv, g);
    ACCBPG_LAUNCHED("sparse_grad_kernel");
    return ACCBPG_OK;
}

}  // extern "C"
