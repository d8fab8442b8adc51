// K2 + K3 of the D-optimal design oracle in one chain: Cholesky M = L L^T (with -log det M from the pivots) and
// L^{-1}, both replicated on every GPU (they act on the all-reduced m x m Gram matrix).
//
// Right-looking over 64-wide block columns, ONE launch per block column J and no dependency between the CTAs of a
// launch.  Every CTA stages the 64x64 diagonal block W[J,J] and redoes the serial part itself -- two in-warp 32x32
// factorisations (rows in registers, pivots broadcast by shuffles) and two in-warp 32x32 triangular inverses, giving
// L_JJ and X = L_JJ^{-1} -- and then does one 64x64 job on the FP64 tensor pipe (DMMA.8x8x4 from padded shared memory):
//
//   trailing tile (R,C), J < C <= R      W[R,C] -= (W[R,J] X^T)(W[C,J] X^T)^T          (L[R,J] = W[R,J] X^T)
//   inverse tile  (C,r), r <= J < C      Y[C,r] -= (W[C,J] X^T) Linv[J,r],  Linv[J,r] = X Y[J,r]  (= X when r = J)
//   row finish    (J,r), r < J           Linv[J,r] = X Y[J,r]
//
// Y[C,r] = -sum_{K=r}^{C-1} L[C,K] Linv[K,r] is the running sum of the block forward substitution  L Linv = I,
// so when the last launch retires both L and L^{-1} are complete: the triangular inverse costs no launch of its own
// and no serial chain beyond the factorisation's.  Tiles are written by exactly one CTA and the block column J that a
// launch reads is never written in it, so the trailing matrix is updated in place.
#include "dmma.cuh"

namespace accbpg {

constexpr int CB = 64;             // block edge
constexpr int CLD = CB + 4;        // 68 doubles: (g*68 + t) and (t*68 + g) mod 16 distinct over a half warp
constexpr int CBUF = CB * CLD;     // doubles per staged block
constexpr int CHOL_SMEM = (6 * CBUF + 64) * 8;

struct CholStep {
    const double* src;   // lower blocks of the matrix this launch reads (M itself at J = 0, W afterwards), ld = m
    double* W;           // trailing matrix being updated in place, ld = m
    double* L;           // optional output (NULL: not wanted), ld = m
    double* Y;           // running sums of the inverse, ld = mp
    double* Linv;        // L^{-1}, ld = mp (zeroed by the host before the first launch)
    double* logacc;      // running sum of log pivots (device scalar)
    double* d_out;       // -log det M, written by the last launch
    uint32_t* status;
    int m, mp, J, nblk, want_inv;
    int nT, nI, nF;      // job counts: trailing tiles, inverse tiles, row-finish tiles
};

// ---- staging ------------------------------------------------------------------------------------------------
// 64x64 block at (r0, c0) of a row-major matrix with `rows` x `cols` valid entries -> smem [64][CLD]; outside: 0
__device__ __forceinline__ void stage_block(double* dst, const double* src, int64_t ld, int r0, int c0, int rows,
                                            int cols, int tid) {
#pragma unroll 4
    for (int e = tid; e < CB * CB; e += 256) {
        int r = e >> 6, c = e & 63;
        int gr = r0 + r, gc = c0 + c;
        dst[r * CLD + c] = (gr < rows && gc < cols) ? __ldcg(src + (int64_t)gr * ld + gc) : 0.0;
    }
}

// ---- the serial part: 32x32 Cholesky and 32x32 lower-triangular inverse inside one warp -------------------------
// Factor the 32x32 block at (off, off) of sD in place (lower factor, zeros above the diagonal).
// Returns sum(log pivot) on every lane; sets *bad when a pivot is not positive.
__device__ __forceinline__ double factor32(double* sD, int off, double* rinv, bool* bad) {
    const int lane = threadIdx.x & 31;
    double row[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) row[c] = sD[(off + lane) * CLD + off + c];
    double mypiv = 1.0, mydiag = 1.0;        // lane k keeps pivot k and L_kk: log / reciprocal leave the chain
#pragma unroll
    for (int k = 0; k < 32; ++k) {
        double pkk = __shfl_sync(0xffffffffu, row[k], k);
        if (!(pkk > 0.0)) { *bad = true; pkk = 1.0; }
        double ri = rsqrt(pkk);
        double lrk = row[k] * ri;
        row[k] = lrk;
        if (lane == k) { mypiv = pkk; mydiag = lrk; }
#pragma unroll
        for (int c = k + 1; c < 32; ++c) {
            double lck = __shfl_sync(0xffffffffu, lrk, c);
            row[c] = fma(-lrk, lck, row[c]);
        }
    }
    rinv[off + lane] = 1.0 / mydiag;
#pragma unroll
    for (int c = 0; c < 32; ++c) sD[(off + lane) * CLD + off + c] = (c <= lane) ? row[c] : 0.0;
    return warp_sum(log(mypiv));
}

// X[off.., off..] = inverse of the lower-triangular 32x32 block of sD at (off, off); lane c solves column c.
__device__ __forceinline__ void invert32(const double* sD, int off, const double* rinv, double* sX) {
    const int lane = threadIdx.x & 31;
    double b[32];
#pragma unroll
    for (int r = 0; r < 32; ++r) b[r] = (r == lane) ? 1.0 : 0.0;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
        double xk = b[k] * rinv[off + k];
        b[k] = xk;
#pragma unroll
        for (int r = k + 1; r < 32; ++r) b[r] = fma(-sD[(off + r) * CLD + off + k], xk, b[r]);
    }
#pragma unroll
    for (int r = 0; r < 32; ++r) sX[(off + r) * CLD + off + lane] = (lane <= r) ? b[r] : 0.0;
}

// ---- 64x64 products on the DMMA pipe ------------------------------------------------------------------------------
// 8 warps as 2 (rows) x 4 (cols); warp tile 32 x 16 = 4 x 2 mma tiles.
// acc[i][j][e] <-> row wm*32 + i*8 + g, col wn*16 + j*8 + 2t + e.
// B_T: the B operand is stored [n][k] (product A * B^T), otherwise [k][n].  NEG: acc -= A*B.
template <bool B_T, bool NEG>
__device__ __forceinline__ void mma64(double (&acc)[4][2][2], const double* A, const double* B, int kbeg, int kend,
                                      int wm, int wn, int g, int t) {
    const double* ap = A + (wm * 32 + g) * CLD + t;
    const double* bp = B_T ? (B + (wn * 16 + g) * CLD + t) : (B + t * CLD + wn * 16 + g);
#pragma unroll 4
    for (int k = kbeg; k < kend; k += 4) {
        double a[4], b[2];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            a[i] = ap[i * 8 * CLD + k];
            if (NEG) a[i] = -a[i];
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) b[j] = B_T ? bp[j * 8 * CLD + k] : bp[k * CLD + j * 8];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
}

__device__ __forceinline__ void acc_zero(double (&acc)[4][2][2]) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
}
__device__ __forceinline__ void acc_to_smem(const double (&acc)[4][2][2], double* dst, int wm, int wn, int g, int t) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
            *reinterpret_cast<double2*>(dst + (wm * 32 + i * 8 + g) * CLD + wn * 16 + j * 8 + 2 * t) =
                make_double2(acc[i][j][0], acc[i][j][1]);
}
// global tile at (r0, c0), valid extent rows x cols
__device__ __forceinline__ void acc_from_global(double (&acc)[4][2][2], const double* src, int64_t ld, int r0, int c0,
                                                int rows, int cols, int wm, int wn, int g, int t) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int gr = r0 + wm * 32 + i * 8 + g;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            int gc = c0 + wn * 16 + j * 8 + 2 * t;
            acc[i][j][0] = (gr < rows && gc < cols) ? __ldcg(src + (int64_t)gr * ld + gc) : 0.0;
            acc[i][j][1] = (gr < rows && gc + 1 < cols) ? __ldcg(src + (int64_t)gr * ld + gc + 1) : 0.0;
        }
    }
}
__device__ __forceinline__ void acc_to_global(const double (&acc)[4][2][2], double* dst, int64_t ld, int r0, int c0,
                                              int rows, int cols, int wm, int wn, int g, int t) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int gr = r0 + wm * 32 + i * 8 + g;
        if (gr >= rows) continue;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            int gc = c0 + wn * 16 + j * 8 + 2 * t;
            if (gc < cols) dst[(int64_t)gr * ld + gc] = acc[i][j][0];
            if (gc + 1 < cols) dst[(int64_t)gr * ld + gc + 1] = acc[i][j][1];
        }
    }
}
__device__ __forceinline__ void smem_to_global(const double* src, double* dst, int64_t ld, int r0, int c0, int rows,
                                               int cols, int tid) {
    for (int e = tid; e < CB * CB; e += 256) {
        int r = e >> 6, c = e & 63;
        int gr = r0 + r, gc = c0 + c;
        if (gr < rows && gc < cols) dst[(int64_t)gr * ld + gc] = src[r * CLD + c];
    }
}

// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 1) chol_inv_step_kernel(CholStep p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* sD = reinterpret_cast<double*>(smem_raw);   // diagonal block -> L_JJ
    double* sX = sD + CBUF;                             // X = L_JJ^{-1}
    double* sA = sX + CBUF;                             // W[R,J]  or  Y[J,r]
    double* sB = sA + CBUF;                             // W[C,J]
    double* sP = sB + CBUF;                             // L[R,J]  or  Linv[J,r]
    double* sQ = sP + CBUF;                             // L[C,J]
    double* rinv = sQ + CBUF;                           // 64 reciprocals of the diagonal of L_JJ
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp >> 2, wn = warp & 3;
    const int m = p.m, J = p.J, j0 = J * CB;

    // ---- which job
    enum { JOB_NONE = 0, JOB_TRAIL, JOB_INV, JOB_FIN };
    int job = JOB_NONE, R = 0, C = 0, r = 0;
    {
        int b = blockIdx.x;
        if (b < p.nT) {
            job = JOB_TRAIL;
            int ti = (int)((sqrt(8.0 * b + 1.0) - 1.0) * 0.5);
            while ((ti + 1) * (ti + 2) / 2 <= b) ++ti;
            while (ti * (ti + 1) / 2 > b) --ti;
            R = J + 1 + ti;
            C = J + 1 + (b - ti * (ti + 1) / 2);
        } else if (b < p.nT + p.nI) {
            job = JOB_INV;
            int e = b - p.nT;
            C = J + 1 + e / (J + 1);
            r = e % (J + 1);
        } else if (b < p.nT + p.nI + p.nF) {
            job = JOB_FIN;
            r = b - p.nT - p.nI;
        }
    }

    // ---- stage everything this CTA reads (the loads overlap the serial part below)
    for (int e = tid; e < CB * CB; e += 256) {
        int rr = e >> 6, cc = e & 63;
        int gr = j0 + rr, gc = j0 + cc;
        double v = (rr == cc) ? 1.0 : 0.0;                       // identity on the padding
        if (gr < m && gc < m) v = __ldcg(p.src + (int64_t)gr * m + gc);
        sD[rr * CLD + cc] = v;
    }
    double acc[4][2][2];
    if (job == JOB_TRAIL) {
        stage_block(sA, p.src, m, R * CB, j0, m, m, tid);
        if (R != C) stage_block(sB, p.src, m, C * CB, j0, m, m, tid);
        acc_from_global(acc, p.src, m, R * CB, C * CB, m, m, wm, wn, g, t);
    } else if (job == JOB_INV) {
        stage_block(sB, p.src, m, C * CB, j0, m, m, tid);
        if (r < J) {
            stage_block(sA, p.Y, p.mp, j0, r * CB, p.mp, p.mp, tid);
            acc_from_global(acc, p.Y, p.mp, C * CB, r * CB, p.mp, p.mp, wm, wn, g, t);
        } else {
            acc_zero(acc);
        }
    } else if (job == JOB_FIN) {
        stage_block(sA, p.Y, p.mp, j0, r * CB, p.mp, p.mp, tid);
    }
    __syncthreads();

    // ---- serial part: L_JJ and X = L_JJ^{-1} from 32x32 pieces
    bool bad = false;
    double logsum = 0.0;
    if (warp == 0) {
        logsum = factor32(sD, 0, rinv, &bad);
        __syncwarp();
        invert32(sD, 0, rinv, sX);
    } else {
        // zero the upper-right quadrants while warp 0 works
        for (int e = tid - 32; e < 32 * 32; e += 224) {
            int rr = e >> 5, cc = 32 + (e & 31);
            sX[rr * CLD + cc] = 0.0;
        }
    }
    __syncthreads();
    {   // L10 = D10 X00^T, then D11 -= L10 L10^T          (32 x 32 outputs, 4 per thread)
        const int i = tid >> 3, c0 = tid & 7;                     // thread owns (i, c0 + 8q), q = 0..3
        double o[4] = {0.0, 0.0, 0.0, 0.0};
        for (int k = 0; k < 32; ++k) {
            double a = sD[(32 + i) * CLD + k];
#pragma unroll
            for (int q = 0; q < 4; ++q) o[q] = fma(a, sX[(c0 + 8 * q) * CLD + k], o[q]);     // X00[c][k] = 0 for k > c
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            sD[(32 + i) * CLD + c0 + 8 * q] = o[q];
            sD[i * CLD + 32 + c0 + 8 * q] = 0.0;
        }
        __syncthreads();
        double d[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) d[q] = sD[(32 + i) * CLD + 32 + c0 + 8 * q];
        for (int k = 0; k < 32; ++k) {
            double a = sD[(32 + i) * CLD + k];
#pragma unroll
            for (int q = 0; q < 4; ++q) d[q] = fma(-a, sD[(32 + c0 + 8 * q) * CLD + k], d[q]);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) sD[(32 + i) * CLD + 32 + c0 + 8 * q] = d[q];
    }
    __syncthreads();
    if (warp == 0) {
        logsum += factor32(sD, 32, rinv, &bad);
        __syncwarp();
        invert32(sD, 32, rinv, sX);
    }
    __syncthreads();
    {   // X10 = -X11 (L10 X00)
        const int i = tid >> 3, c0 = tid & 7;
        double o[4] = {0.0, 0.0, 0.0, 0.0};
        for (int k = 0; k < 32; ++k) {
            double a = sD[(32 + i) * CLD + k];
#pragma unroll
            for (int q = 0; q < 4; ++q) o[q] = fma(a, sX[k * CLD + c0 + 8 * q], o[q]);
        }
        // park T = L10 X00 in the (still unused) lower-left quadrant of sP
#pragma unroll
        for (int q = 0; q < 4; ++q) sP[(32 + i) * CLD + c0 + 8 * q] = o[q];
        __syncthreads();
        double x[4] = {0.0, 0.0, 0.0, 0.0};
        for (int k = 0; k < 32; ++k) {
            double a = sX[(32 + i) * CLD + 32 + k];
#pragma unroll
            for (int q = 0; q < 4; ++q) x[q] = fma(-a, sP[(32 + k) * CLD + c0 + 8 * q], x[q]);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) sX[(32 + i) * CLD + c0 + 8 * q] = x[q];
    }
    __syncthreads();

    // ---- CTA 0 publishes the diagonal results
    if (blockIdx.x == 0) {
        if (p.L) smem_to_global(sD, p.L, m, j0, j0, m, m, tid);
        if (p.want_inv) smem_to_global(sX, p.Linv, p.mp, j0, j0, p.mp, p.mp, tid);
        if (tid == 0) {
            double tot = (J == 0 ? 0.0 : p.logacc[0]) + logsum;
            p.logacc[0] = tot;
            if (J == p.nblk - 1) p.d_out[0] = -tot;
            if (bad) atomicOr(p.status, ACCBPG_ST_NOT_PD);
        }
    }
    if (job == JOB_NONE) return;

    // ---- the 64x64 job
    double tmp[4][2][2];
    if (job == JOB_TRAIL) {
        acc_zero(tmp);
        mma64<true, false>(tmp, sA, sX, 0, wn * 16 + 16, wm, wn, g, t);       // L[R,J] = W[R,J] X^T  (X[n][k] = 0, k > n)
        acc_to_smem(tmp, sP, wm, wn, g, t);
        const double* pc = sP;
        if (R != C) {
            acc_zero(tmp);
            mma64<true, false>(tmp, sB, sX, 0, wn * 16 + 16, wm, wn, g, t);   // L[C,J]
            acc_to_smem(tmp, sQ, wm, wn, g, t);
            pc = sQ;
        }
        __syncthreads();
        mma64<true, true>(acc, sP, pc, 0, CB, wm, wn, g, t);                  // W[R,C] -= L[R,J] L[C,J]^T
        acc_to_global(acc, p.W, m, R * CB, C * CB, m, m, wm, wn, g, t);
        if (R == C && p.L) smem_to_global(sP, p.L, m, R * CB, j0, m, m, tid);
        return;
    }
    if (job == JOB_INV) {
        acc_zero(tmp);
        mma64<true, false>(tmp, sB, sX, 0, wn * 16 + 16, wm, wn, g, t);       // L[C,J]
        acc_to_smem(tmp, sQ, wm, wn, g, t);
        const double* q = sX;                                                 // Linv[J,J] = X
        int kbeg = wn * 16;                                                   // X[k][n] = 0 for k < n
        if (r < J) {
            acc_zero(tmp);
            mma64<false, false>(tmp, sX, sA, 0, wm * 32 + 32, wm, wn, g, t);  // Linv[J,r] = X Y[J,r]  (X[i][k] = 0, k > i)
            acc_to_smem(tmp, sP, wm, wn, g, t);
            q = sP;
            kbeg = 0;
        }
        __syncthreads();
        mma64<false, true>(acc, sQ, q, kbeg, CB, wm, wn, g, t);               // Y[C,r] -= L[C,J] Linv[J,r]
        acc_to_global(acc, p.Y, p.mp, C * CB, r * CB, p.mp, p.mp, wm, wn, g, t);
        return;
    }
    // JOB_FIN
    acc_zero(tmp);
    mma64<false, false>(tmp, sX, sA, 0, wm * 32 + 32, wm, wn, g, t);
    acc_to_global(tmp, p.Linv, p.mp, j0, r * CB, p.mp, p.mp, wm, wn, g, t);
}

static bool g_chol_attr = false;
static int ensure_attrs() {
    if (g_chol_attr) return ACCBPG_OK;
    ACCBPG_CUDA(cudaFuncSetAttribute(chol_inv_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CHOL_SMEM));
    g_chol_attr = true;
    return ACCBPG_OK;
}

// ------------------------------------------------------------------------------------------ host entry point
// M (m x m, symmetric, only read) -> d_out[0] = -log det M;  L (may be NULL) <- lower factor;  when want_inv:
// Linv (mp x mp, ld mp) <- L^{-1}, zero above the diagonal and on rows / columns >= 64*ceil(m/64).
// W: m x m scratch (trailing matrix), Y: mp x mp scratch (running sums), acc: device scalar.
int chol_factor_inv(Ctx* c, cudaStream_t s, int m, int mp, const double* M, double* L, int want_inv, double* Linv,
                    double* W, double* Y, double* acc, double* d_out) {
    int rc = ensure_attrs();
    if (rc) return rc;
    if (L) ACCBPG_CUDA(cudaMemsetAsync(L, 0, (size_t)m * m * sizeof(double), s));
    if (want_inv) ACCBPG_CUDA(cudaMemsetAsync(Linv, 0, (size_t)mp * mp * sizeof(double), s));
    ProfScope ps(P_CHOL, s);
    CholStep p;
    p.W = W; p.L = L; p.Y = Y; p.Linv = Linv; p.logacc = acc; p.d_out = d_out; p.status = c->d_status;
    p.m = m; p.mp = mp; p.want_inv = want_inv;
    p.nblk = (m + CB - 1) / CB;
    for (int J = 0; J < p.nblk; ++J) {
        const int below = p.nblk - 1 - J;
        p.J = J;
        p.src = (J == 0) ? M : W;
        p.nT = below * (below + 1) / 2;
        p.nI = want_inv ? (J + 1) * below : 0;
        p.nF = want_inv ? J : 0;
        int grid = p.nT + p.nI + p.nF;
        if (grid < 1) grid = 1;
        chol_inv_step_kernel<<<grid, 256, CHOL_SMEM, s>>>(p);
        ACCBPG_LAUNCHED("chol_inv_step_kernel");
    }
    return ACCBPG_OK;
}

}  // namespace accbpg
