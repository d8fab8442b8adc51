// K2 / K3 of the D-optimal design oracle: Cholesky M = L L^T (with -log det M from the pivots) and L^{-1}.
// Both replicate on every GPU (they act on the all-reduced m x m Gram matrix).
//
// Cholesky: right-looking, 32-wide panels, ONE launch per panel and no dependency between the CTAs of a launch.
//   Launch j holds the trailing matrix (already updated by panels < j) in Win.  Every CTA owns one 64x64 tile
//   (R, C) of the next trailing matrix and recomputes what it needs redundantly: it factors the 32x32 diagonal
//   block in warp 0 (rows in registers, pivots broadcast by shuffles -- the serial chain of the algorithm, ~3 us),
//   solves its own panel rows  L[R,J] = Win[R,J] Ld^{-T}  and  L[C,J]  likewise, and writes
//   Wout[R,C] = Win[R,C] - L[R,J] L[C,J]^T.  Win/Wout ping-pong, so nothing read in a launch is written in it.
//   The critical path per panel is one kernel (factor chain + a 64x64x32 update) instead of a K-growing update.
// Triangular inverse: 128x128 diagonal blocks are inverted inside one CTA each (32x32 in-warp substitutions, then
//   two in-smem doubling levels); larger levels use recursive doubling  inv([A 0; B C]) = [A^-1 0; -C^-1 B A^-1 C^-1]
//   with both products on the FP64 DMMA GEMM.
#include "dmma.cuh"

namespace accbpg {

constexpr int CS_NB = 32;          // panel width
constexpr int CS_T = 64;           // trailing tile edge
constexpr int CS_LD = CS_T + 1;

__global__ void __launch_bounds__(256, 2) chol_step_kernel(const double* __restrict__ Win, double* __restrict__ Wout,
                                                           double* __restrict__ L, int m, int j0, int ntile,
                                                           double* logdet_acc, uint32_t* status) {
    __shared__ double Ld[CS_NB][CS_NB + 1];
    __shared__ double rinv[CS_NB];
    __shared__ double Sr[CS_NB][CS_LD];       // Sr[k][row]: panel entries of the tile's rows R
    __shared__ double Sc[CS_NB][CS_LD];       // same for the tile's columns C (rows C of the panel)
    const int tid = threadIdx.x;
    const int nb = min(CS_NB, m - j0);
    const int t0 = j0 + CS_NB;
    int ti = 0, tc = 0;
    if (ntile > 0) {
        int tt = blockIdx.x;
        ti = (int)((sqrt(8.0 * tt + 1.0) - 1.0) * 0.5);
        while ((ti + 1) * (ti + 2) / 2 <= tt) ++ti;
        while (ti * (ti + 1) / 2 > tt) --ti;
        tc = tt - ti * (ti + 1) / 2;
    }
    const bool diag_tile = (ti == tc);
    const int rbase = t0 + ti * CS_T, cbase = t0 + tc * CS_T;

    // ---- phase 0: stage the diagonal block, the two panel strips and this thread's 4x4 outputs
    for (int e = tid; e < CS_NB * CS_NB; e += 256) {
        int r = e >> 5, c = e & 31;
        Ld[r][c] = (r < nb && c < nb) ? Win[(size_t)(j0 + r) * m + j0 + c] : ((r == c) ? 1.0 : 0.0);
    }
    const int ty = tid >> 4, tx = tid & 15;
    double acc[4][4];
    if (ntile > 0) {
        for (int e = tid; e < CS_T * CS_NB; e += 256) {
            int row = e >> 5, k = e & 31;
            int gr = rbase + row, gc = cbase + row;
            Sr[k][row] = (gr < m && k < nb) ? Win[(size_t)gr * m + j0 + k] : 0.0;
            if (!diag_tile) Sc[k][row] = (gc < m && k < nb) ? Win[(size_t)gc * m + j0 + k] : 0.0;
        }
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            int gr = rbase + ty + 16 * a;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                int gc = cbase + tx + 16 * b;
                acc[a][b] = (gr < m && gc < m) ? Win[(size_t)gr * m + gc] : 0.0;
            }
        }
    }
    __syncthreads();

    // ---- phase 1: warp 0 factors the diagonal block (lane r holds row r)
    if (tid < 32) {
        const int lane = tid;
        double row[CS_NB];
#pragma unroll
        for (int c = 0; c < CS_NB; ++c) row[c] = Ld[lane][c];
        double mypiv = 1.0, mydiag = 1.0;        // lane k keeps pivot k and L_kk: log / reciprocal leave the chain
        bool bad = false;
#pragma unroll
        for (int k = 0; k < CS_NB; ++k) {
            double pkk = __shfl_sync(0xffffffffu, row[k], k);
            if (!(pkk > 0.0)) { bad = true; pkk = 1.0; }
            double ri = rsqrt(pkk);
            double lrk = row[k] * ri;
            row[k] = lrk;
            if (lane == k) { mypiv = pkk; mydiag = lrk; }
#pragma unroll
            for (int c = k + 1; c < CS_NB; ++c) {
                double lck = __shfl_sync(0xffffffffu, lrk, c);
                row[c] = fma(-lrk, lck, row[c]);
            }
        }
        rinv[lane] = 1.0 / mydiag;
#pragma unroll
        for (int c = 0; c < CS_NB; ++c) Ld[lane][c] = (c <= lane) ? row[c] : 0.0;
        if (blockIdx.x == 0) {
            double logsum = warp_sum(log(mypiv));
            if (lane == 0) logdet_acc[0] += logsum;          // one writer per launch; launches are stream ordered
            if (bad && lane == 0) atomicOr(status, ACCBPG_ST_NOT_PD);
#pragma unroll
            for (int c = 0; c < CS_NB; ++c)
                if (lane < nb && c <= lane) L[(size_t)(j0 + lane) * m + j0 + c] = row[c];
        }
    }
    __syncthreads();
    if (ntile == 0) return;

    // ---- phase 2: X = S Ld^{-T} for the rows of R (threads 0..63) and of C (threads 64..127)
    if (tid < 128 && !(tid >= 64 && diag_tile)) {
        double (*S)[CS_LD] = (tid < 64) ? Sr : Sc;
        const int row = tid & 63;
        double sv[CS_NB];
#pragma unroll
        for (int c = 0; c < CS_NB; ++c) sv[c] = S[c][row];
#pragma unroll
        for (int k = 0; k < CS_NB; ++k) {
            double xk = sv[k] * rinv[k];
            sv[k] = xk;
#pragma unroll
            for (int c = k + 1; c < CS_NB; ++c) sv[c] = fma(-xk, Ld[c][k], sv[c]);
        }
#pragma unroll
        for (int c = 0; c < CS_NB; ++c) S[c][row] = sv[c];
    }
    __syncthreads();

    // the diagonal tiles own the final panel rows L[R, J]
    if (diag_tile) {
        for (int e = tid; e < CS_T * CS_NB; e += 256) {
            int row = e >> 5, k = e & 31;
            int gr = rbase + row;
            if (gr < m && k < nb) L[(size_t)gr * m + j0 + k] = Sr[k][row];
        }
    }

    // ---- phase 3: Wout[R,C] = Win[R,C] - L[R,J] L[C,J]^T
    double (*Sb)[CS_LD] = diag_tile ? Sr : Sc;
#pragma unroll 8
    for (int k = 0; k < CS_NB; ++k) {
        double ar[4], bc[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) ar[a] = Sr[k][ty + 16 * a];
#pragma unroll
        for (int b = 0; b < 4; ++b) bc[b] = Sb[k][tx + 16 * b];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = fma(-ar[a], bc[b], acc[a][b]);
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        int gr = rbase + ty + 16 * a;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            int gc = cbase + tx + 16 * b;
            if (gr < m && gc < m) Wout[(size_t)gr * m + gc] = acc[a][b];
        }
    }
}

__global__ void store_neg_kernel(const double* src, double* dst) { dst[0] = -src[0]; }

// ------------------------------------------------------------------------------------------ L^{-1}: 128-blocks
constexpr int TI_B = 128;
constexpr int TI_LD = TI_B + 1;
constexpr int TI_SMEM = (TI_B * TI_LD + 2 * 64 * 65) * 8;      // Lw (phase a) aliases S1/S2 (phase b)

// b x b products on shared memory operands, all 256 threads:  dst[i][j] = alpha * sum_k A[i][k] * B[k][j]
__device__ __forceinline__ void cta_gemm(double* dst, int ldd, const double* A, int lda, const double* B, int ldb, int b,
                                         double alpha) {
    for (int e = threadIdx.x; e < b * b; e += 256) {
        int i = e / b, j = e - i * b;
        double s = 0.0;
        for (int k = 0; k < b; ++k) s = fma(A[i * lda + k], B[k * ldb + j], s);
        dst[i * ldd + j] = alpha * s;
    }
}

__global__ void __launch_bounds__(256) trinv_diag128_kernel(const double* __restrict__ L, int m, double* __restrict__ Linv,
                                                            int mp) {
    extern __shared__ double smem_d[];
    double* Xs = smem_d;                       // [128][129] the block inverse being assembled
    double* S1 = Xs + TI_B * TI_LD;            // [64][65] staged off-diagonal block of L
    double* S2 = S1 + 64 * 65;                 // [64][65] B * A^-1
    double* Lw = S1;                           // [4][32][33] diagonal 32-blocks of L (dead before S1/S2 are used)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int base = blockIdx.x * TI_B;
    for (int e = tid; e < TI_B * TI_LD; e += 256) Xs[e] = 0.0;
    __syncthreads();
    // (a) 32x32 diagonal inverses, one warp each: lane c solves column c of  Lqq X = I
    if (warp < 4) {
        double* Ls = Lw + warp * 32 * 33;
        const int q0 = base + warp * 32;
        for (int r = 0; r < 32; ++r) {
            int gr = q0 + r, gc = q0 + lane;
            double v = (r == lane) ? 1.0 : 0.0;
            if (gr < m && gc < m && lane <= r) v = L[(size_t)gr * m + gc];
            Ls[r * 33 + lane] = v;
        }
        __syncwarp();
        double b[32];
#pragma unroll
        for (int r = 0; r < 32; ++r) b[r] = (r == lane) ? 1.0 : 0.0;
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            double xk = b[k] / Ls[k * 33 + k];
            b[k] = xk;
#pragma unroll
            for (int r = k + 1; r < 32; ++r) b[r] = fma(-Ls[r * 33 + k], xk, b[r]);
        }
#pragma unroll
        for (int r = 0; r < 32; ++r)
            if (lane <= r) Xs[(warp * 32 + r) * TI_LD + warp * 32 + lane] = b[r];
    }
    __syncthreads();
    // (b) doubling inside the block: b = 32 (two pairs), then b = 64 (one pair)
    for (int b = 32; b <= 64; b <<= 1) {
        for (int r0 = 0; r0 + b < TI_B; r0 += 2 * b) {
            const int c0 = r0 + b;
            for (int e = tid; e < b * b; e += 256) {          // stage B = L[c0.., r0..]
                int i = e / b, k = e - i * b;
                int gr = base + c0 + i, gc = base + r0 + k;
                S1[i * 65 + k] = (gr < m && gc < m) ? L[(size_t)gr * m + gc] : 0.0;
            }
            __syncthreads();
            cta_gemm(S2, 65, S1, 65, Xs + r0 * TI_LD + r0, TI_LD, b, 1.0);                 // T = B A^-1
            __syncthreads();
            cta_gemm(Xs + c0 * TI_LD + r0, TI_LD, Xs + c0 * TI_LD + c0, TI_LD, S2, 65, b, -1.0);   // X = -C^-1 T
            __syncthreads();
        }
    }
    for (int e = tid; e < TI_B * TI_B; e += 256) {
        int r = e >> 7, c = e & 127;
        if (c <= r) Linv[(size_t)(base + r) * mp + base + c] = Xs[r * TI_LD + c];
    }
}

// ------------------------------------------------------------------------------------------ DMMA GEMM, stored
// C[i][j] = alpha * sum_k A[i][k] B[k][j];  A rows >= a_rows and k >= kdim read as zero; C rows < c_rows stored.
struct GemmNN {
    const double* A; int64_t lda; int a_rows;
    const double* B; int64_t ldb;
    double* C; int64_t ldc; int c_rows;
    int ncols, kdim;
    double alpha;
};

template <bool A_ALIGNED16>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_nn_store_kernel(GemmNN p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* smem = reinterpret_cast<double*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp >> 2, wn = warp & 3;
    const int i0 = blockIdx.x * BM;
    const int64_t j0 = (int64_t)blockIdx.y * BN;
    const int KT = (p.kdim + BK - 1) / BK;
    double acc[MI][NI][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    auto load_stage = [&](int stage, int kt) {
        double* As = smem + stage * NN_STAGE_DOUBLES;
        double* Bs = As + BM * A_LD;
        load_kmajor_slab<A_ALIGNED16>(As, p.A, p.lda, i0, p.a_rows, (int64_t)kt * BK, p.kdim, tid);
        load_nmajor_slab<true>(Bs, p.B, p.ldb, kt * BK, p.kdim, j0, p.ncols, tid);
    };
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < KT) load_stage(s, s);
        cp_async_commit();
    }
    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        int nk = kt + STAGES - 1;
        if (nk < KT) load_stage(nk % STAGES, nk);
        cp_async_commit();
        const double* As = smem + (kt % STAGES) * NN_STAGE_DOUBLES;
        mma_nn_slab(acc, As, As + BM * A_LD, wm, wn, g, t);
    }
    cp_async_wait<0>();
#pragma unroll
    for (int i = 0; i < MI; ++i) {
        int row = i0 + wm * 64 + i * 8 + g;
        if (row >= p.c_rows) continue;
#pragma unroll
        for (int j = 0; j < NI; ++j) {
            int64_t col = j0 + wn * 32 + j * 8 + 2 * t;
            if (col + 1 < p.ncols)
                *reinterpret_cast<double2*>(p.C + (size_t)row * p.ldc + col) =
                    make_double2(p.alpha * acc[i][j][0], p.alpha * acc[i][j][1]);
            else if (col < p.ncols)
                p.C[(size_t)row * p.ldc + col] = p.alpha * acc[i][j][0];
        }
    }
}

static bool g_chol_attr = false;
static int ensure_attrs() {
    if (g_chol_attr) return ACCBPG_OK;
    ACCBPG_CUDA(cudaFuncSetAttribute(trinv_diag128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TI_SMEM));
    ACCBPG_CUDA(cudaFuncSetAttribute(gemm_nn_store_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, NN_SMEM));
    ACCBPG_CUDA(cudaFuncSetAttribute(gemm_nn_store_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, NN_SMEM));
    g_chol_attr = true;
    return ACCBPG_OK;
}

static int launch_gemm(cudaStream_t s, const GemmNN& p) {
    dim3 grid((p.c_rows + BM - 1) / BM, (p.ncols + BN - 1) / BN);
    bool al = ((reinterpret_cast<uintptr_t>(p.A) & 15u) == 0) && (p.lda % 2 == 0);
    if (al) gemm_nn_store_kernel<true><<<grid, GEMM_THREADS, NN_SMEM, s>>>(p);
    else    gemm_nn_store_kernel<false><<<grid, GEMM_THREADS, NN_SMEM, s>>>(p);
    ACCBPG_LAUNCHED("gemm_nn_store_kernel");
    return ACCBPG_OK;
}

// ------------------------------------------------------------------------------------------ host entry points
int chol_factor(Ctx* c, cudaStream_t s, int m, const double* M, double* L, double* Wa, double* Wb, double* acc,
                double* d_out) {                     // acc: device scalar, running sum of log pivots
    ACCBPG_CUDA(cudaMemsetAsync(acc, 0, sizeof(double), s));
    ACCBPG_CUDA(cudaMemsetAsync(L, 0, (size_t)m * m * sizeof(double), s));
    {
        ProfScope ps(P_CHOL, s);
        const double* win = M;
        double* wout = Wa;
        for (int j0 = 0; j0 < m; j0 += CS_NB) {
            int below = m - j0 - CS_NB;
            int ntile = below > 0 ? (below + CS_T - 1) / CS_T : 0;
            int grid = ntile > 0 ? ntile * (ntile + 1) / 2 : 1;
            chol_step_kernel<<<grid, 256, 0, s>>>(win, wout, L, m, j0, ntile, acc, c->d_status);
            ACCBPG_LAUNCHED("chol_step_kernel");
            win = wout;
            wout = (wout == Wa) ? Wb : Wa;
        }
    }
    store_neg_kernel<<<1, 1, 0, s>>>(acc, d_out);
    ACCBPG_LAUNCHED("store_neg_kernel");
    return ACCBPG_OK;
}

int tri_inverse(Ctx* c, cudaStream_t s, int m, int mp, const double* L, double* Linv, double* T) {
    (void)c;
    int rc = ensure_attrs();
    if (rc) return rc;
    ACCBPG_CUDA(cudaMemsetAsync(Linv, 0, (size_t)mp * mp * sizeof(double), s));
    ProfScope ps(P_TRINV, s);
    trinv_diag128_kernel<<<mp / TI_B, 256, TI_SMEM, s>>>(L, m, Linv, mp);
    ACCBPG_LAUNCHED("trinv_diag128_kernel");
    for (int b = TI_B; b < mp; b <<= 1) {
        for (int r0 = 0; r0 + b < mp; r0 += 2 * b) {
            const int c0 = r0 + b;
            if (c0 >= m) continue;                               // the whole lower block is padding
            const int rows_c = (mp - c0 < b) ? (mp - c0) : b;
            GemmNN g1;                                           // T = B A^-1
            g1.A = L + (size_t)c0 * m + r0; g1.lda = m; g1.a_rows = (m - c0 < rows_c) ? (m - c0) : rows_c;
            g1.B = Linv + (size_t)r0 * mp + r0; g1.ldb = mp;
            g1.C = T + (size_t)c0 * mp + r0; g1.ldc = mp; g1.c_rows = rows_c;
            g1.ncols = b; g1.kdim = b; g1.alpha = 1.0;
            rc = launch_gemm(s, g1);
            if (rc) return rc;
            GemmNN g2;                                           // X = -C^-1 T
            g2.A = Linv + (size_t)c0 * mp + c0; g2.lda = mp; g2.a_rows = rows_c;
            g2.B = T + (size_t)c0 * mp + r0; g2.ldb = mp;
            g2.C = Linv + (size_t)c0 * mp + r0; g2.ldc = mp; g2.c_rows = rows_c;
            g2.ncols = b; g2.kdim = rows_c; g2.alpha = -1.0;
            rc = launch_gemm(s, g2);
            if (rc) return rc;
        }
    }
    return ACCBPG_OK;
}

}  // namespace accbpg
