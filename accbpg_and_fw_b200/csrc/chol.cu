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
#include <functional>
#include <map>
#include "dmma.cuh"

namespace accbpg {

constexpr int CB = 64;             // block edge
constexpr int CLD = CB + 4;        // 68 doubles: (g*68 + t) and (t*68 + g) mod 16 distinct over a half warp
constexpr int CBUF = CB * CLD;     // doubles per staged block
constexpr int CHOL_SMEM = (6 * CBUF + 64 + 96 + 32 * 34) * 8;

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
#ifdef CHOL_TRACE
    long long* trace;    // [nblk][16] clock64 stamps of CTA 0 (tools/chol_trace.cu)
#endif
};

#ifdef CHOL_TRACE
#define CHOL_STAMP(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) p.trace[p.J * 16 + (i)] = clock64(); } while (0)
#define CHOL_STAMP_NS(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) { long long t_; \
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); p.trace[p.J * 16 + (i)] = t_; } } while (0)
#else
#define CHOL_STAMP(i)
#define CHOL_STAMP_NS(i)
#endif

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

// the same through cp.async (zero-filled outside): nothing waits until cp_async_wait.  AL16: base pointer 16-byte
// aligned and ld, rows, cols even, so an aligned pair of columns is inside or outside as a whole.
template <bool AL16>
__device__ __forceinline__ void stage_block_async(double* dst, const double* src, int64_t ld, int r0, int c0, int rows,
                                                  int cols, int tid) {
    if (AL16) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            int e = tid + q * 256;
            int r = e >> 5, c = (e & 31) * 2;
            int gr = r0 + r, gc = c0 + c;
            bool ok = (gr < rows && gc < cols);
            cp_async16(dst + r * CLD + c, ok ? (src + (int64_t)gr * ld + gc) : src, ok ? 16 : 0);
        }
    } else {
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            int e = tid + q * 256;
            int r = e >> 6, c = e & 63;
            int gr = r0 + r, gc = c0 + c;
            bool ok = (gr < rows && gc < cols);
            cp_async8(dst + r * CLD + c, ok ? (src + (int64_t)gr * ld + gc) : src, ok ? 8 : 0);
        }
    }
}

// ---- the serial part: 32x32 Cholesky and 32x32 lower-triangular inverse inside one warp -------------------------
// reciprocal without the library's special-case branch (pivots are positive, normal numbers): MUFU seed + the same
// five-FMA refinement the compiler's own division uses
__device__ __forceinline__ double rcp_pos(double d) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    double e = fma(-d, r, 1.0);
    e = fma(e, e, e);
    return fma(r, e, r);           // relative error ~ (seed error)^3: below one ulp (tools/lat_probe.cu checks it)
}

// Factor the 32x32 block at (off, off) of sD in place (lower factor, zeros above the diagonal); lane r holds row r.
// The loop is the square-root-free recurrence  a_rc -= (a_rk / d_k) a_ck : the reciprocal of the pivot is the only
// long-latency operation on the column-to-column chain (the square roots are taken after the loop, all at once), and
// column k reaches the other lanes through shared memory (one store + broadcast loads) instead of 2(31-k) shuffles.
// The loop is software pipelined by one column: as soon as column k has updated entry k+1 of every row, column k+1's
// pivot broadcast and reciprocal are issued, and the remaining updates of column k fill that latency.
// Lt (ld 34) receives the transposed factor for invert32.  Returns sum(log pivot) on every lane.
constexpr int LT_LD = 34;
__device__ __noinline__ double factor32(double* sD, int off, double* rinv, double* colbuf, double* Lt, bool* bad) {
    const int lane = threadIdx.x & 31;
    double row[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) row[c] = sD[(off + lane) * CLD + off + c];
    double mypiv = 1.0;                       // lane k keeps pivot k
    double v, crit;                           // crit: entry k+1 of column k, the one operand the chain waits for
    {
        const double u = row[0];
        colbuf[lane] = u;
        const double dk = __shfl_sync(0xffffffffu, u, 0);
        __syncwarp();
        crit = colbuf[1];
        if (lane == 0) mypiv = dk;
        v = u * rcp_pos(dk);
    }
    // column buffers rotate over three slots: the store of column k+1 cannot meet a straggling read of column k-1
#pragma unroll
    for (int k = 0; k < 32; ++k) {
        const double* cb = colbuf + (k % 3) * 32;
        double vn = 0.0, critn = 0.0;
        if (k + 1 < 32) {
            double* cbn = colbuf + ((k + 1) % 3) * 32;
            row[k + 1] = fma(-v, crit, row[k + 1]);
            const double u = row[k + 1];
            cbn[lane] = u;
            const double dk = __shfl_sync(0xffffffffu, u, k + 1);
            __syncwarp();
            if (k + 2 < 32) critn = cbn[k + 2];
            if (lane == k + 1) mypiv = dk;
            vn = u * rcp_pos(dk);
        }
#pragma unroll
        for (int c = k + 2; c < 32; ++c) row[c] = fma(-v, cb[c], row[c]);
        v = vn;
        crit = critn;
    }
    // the positivity test is off the chain: a pivot <= 0 (or NaN) poisons what follows, and the caller raises anyway
    if (__any_sync(0xffffffffu, !(mypiv > 0.0))) { *bad = true; mypiv = 1.0; }
    // L_rk = a_rk / sqrt(d_k)
    const double rs = rsqrt(mypiv);
    __syncwarp();
    colbuf[lane] = rs;
    rinv[off + lane] = rs;                    // 1 / L_ll
    __syncwarp();
#pragma unroll
    for (int c = 0; c < 32; ++c) {
        const double l = (c <= lane) ? row[c] * colbuf[c] : 0.0;
        sD[(off + lane) * CLD + off + c] = l;
        Lt[c * LT_LD + lane] = l;
    }
    return warp_sum(log(mypiv));
}

// X[off.., off..] = inverse of the lower-triangular 32x32 block whose transpose sits in Lt; lane c solves column c.
// Same one-step look-ahead: x_{k+1} is formed before the rest of column k's updates are issued.
__device__ __noinline__ void invert32(const double* Lt, int off, const double* rinv, double* sX) {
    const int lane = threadIdx.x & 31;
    double b[32];
#pragma unroll
    for (int r = 0; r < 32; ++r) b[r] = (r == lane) ? 1.0 : 0.0;
    double xk = b[0] * rinv[off];
#pragma unroll
    for (int k = 0; k < 32; ++k) {
        b[k] = xk;
        double xn = 0.0;
        if (k + 1 < 32) {
            b[k + 1] = fma(-Lt[k * LT_LD + k + 1], xk, b[k + 1]);
            xn = b[k + 1] * rinv[off + k + 1];
        }
        // remaining rows of column k, two per 16-byte broadcast load (row k*LT_LD is 16-byte aligned: LT_LD is even)
        if (((k + 2) & 1) && k + 2 < 32) b[k + 2] = fma(-Lt[k * LT_LD + k + 2], xk, b[k + 2]);
#pragma unroll
        for (int r2 = (k + 3) & ~1; r2 < 32; r2 += 2) {
            const double2 ll = *reinterpret_cast<const double2*>(Lt + k * LT_LD + r2);
            b[r2] = fma(-ll.x, xk, b[r2]);
            b[r2 + 1] = fma(-ll.y, xk, b[r2 + 1]);
        }
        xk = xn;
    }
#pragma unroll
    for (int r = 0; r < 32; ++r) sX[(off + r) * CLD + off + lane] = (lane <= r) ? b[r] : 0.0;
}

// 32x32 products on the DMMA pipe, all 8 warps: warp w owns rows (w>>1)*8.., cols (w&1)*16.. (1 x 2 mma tiles).
// acc[j][e] <-> row (w>>1)*8 + g, col (w&1)*16 + j*8 + 2t + e.  A is [i][k]; B is [n][k] (B_T) or [k][n].
template <bool B_T, bool NEG>
__device__ __forceinline__ void mma32(double (&acc)[2][2], const double* A, const double* B, int warp, int g, int t) {
    const double* ap = A + ((warp >> 1) * 8 + g) * CLD + t;
    const double* bp = B_T ? (B + ((warp & 1) * 16 + g) * CLD + t) : (B + t * CLD + (warp & 1) * 16 + g);
#pragma unroll
    for (int k = 0; k < 32; k += 4) {
        double a = ap[k];
        if (NEG) a = -a;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const double b = B_T ? bp[j * 8 * CLD + k] : bp[k * CLD + j * 8];
            dmma884(acc[j][0], acc[j][1], a, b);
        }
    }
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
template <bool AL16>
__device__ __forceinline__ void acc_from_global(double (&acc)[4][2][2], const double* src, int64_t ld, int r0, int c0,
                                                int rows, int cols, int wm, int wn, int g, int t) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int gr = r0 + wm * 32 + i * 8 + g;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            int gc = c0 + wn * 16 + j * 8 + 2 * t;
            if (AL16) {
                double2 v = make_double2(0.0, 0.0);
                if (gr < rows && gc < cols) v = __ldcg(reinterpret_cast<const double2*>(src + (int64_t)gr * ld + gc));
                acc[i][j][0] = v.x; acc[i][j][1] = v.y;
            } else {
                acc[i][j][0] = (gr < rows && gc < cols) ? __ldcg(src + (int64_t)gr * ld + gc) : 0.0;
                acc[i][j][1] = (gr < rows && gc + 1 < cols) ? __ldcg(src + (int64_t)gr * ld + gc + 1) : 0.0;
            }
        }
    }
}
template <bool AL16>
__device__ __forceinline__ void acc_to_global(const double (&acc)[4][2][2], double* dst, int64_t ld, int r0, int c0,
                                              int rows, int cols, int wm, int wn, int g, int t) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int gr = r0 + wm * 32 + i * 8 + g;
        if (gr >= rows) continue;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            int gc = c0 + wn * 16 + j * 8 + 2 * t;
            if (AL16) {
                if (gc < cols)
                    *reinterpret_cast<double2*>(dst + (int64_t)gr * ld + gc) = make_double2(acc[i][j][0], acc[i][j][1]);
            } else {
                if (gc < cols) dst[(int64_t)gr * ld + gc] = acc[i][j][0];
                if (gc + 1 < cols) dst[(int64_t)gr * ld + gc + 1] = acc[i][j][1];
            }
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
// AL16: m is even and M / W / L are 16-byte aligned (Y and Linv always are: their leading dimension is a multiple of 128)
template <bool AL16>
__global__ void __launch_bounds__(256, 1) chol_inv_step_kernel(CholStep p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* sD = reinterpret_cast<double*>(smem_raw);   // diagonal block -> L_JJ
    double* sX = sD + CBUF;                             // X = L_JJ^{-1}
    double* sA = sX + CBUF;                             // W[R,J]  or  Y[J,r]
    double* sB = sA + CBUF;                             // W[C,J]
    double* sP = sB + CBUF;                             // L[R,J]  or  Linv[J,r]
    double* sQ = sP + CBUF;                             // L[C,J]
    double* rinv = sQ + CBUF;                           // 64 reciprocals of the diagonal of L_JJ
    double* colbuf = rinv + 64;                         // 3 x 32: the columns being eliminated (rotating slots)
    double* sLt = colbuf + 96;                          // 32 x 34: transposed 32x32 factor for the in-warp inverse
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp >> 2, wn = warp & 3;
    const int m = p.m, J = p.J, j0 = J * CB;

    CHOL_STAMP_NS(14);
    CHOL_STAMP(0);
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

    // ---- stage.  Order matters for the serial part that follows: the diagonal block's loads go out first, then the
    //      job's operand blocks as cp.async (they land while the serial part runs), then the accumulator tile.
    // Everything before this point overlapped the previous launch's tail (programmatic dependent launch).
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    {
        double v[16];
        if (AL16) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                int e = tid + q * 256;
                int rr = e >> 5, cc = (e & 31) * 2;
                int gr = j0 + rr, gc = j0 + cc;
                double2 d2 = make_double2(rr == cc ? 1.0 : 0.0, rr == cc + 1 ? 1.0 : 0.0);     // identity on the padding
                if (gr < m && gc < m) d2 = __ldcg(reinterpret_cast<const double2*>(p.src + (int64_t)gr * m + gc));
                v[2 * q] = d2.x; v[2 * q + 1] = d2.y;
            }
        } else {
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                int e = tid + q * 256;
                int rr = e >> 6, cc = e & 63;
                int gr = j0 + rr, gc = j0 + cc;
                v[q] = (rr == cc) ? 1.0 : 0.0;
                if (gr < m && gc < m) v[q] = __ldcg(p.src + (int64_t)gr * m + gc);
            }
        }
        if (job == JOB_TRAIL) {
            stage_block_async<AL16>(sA, p.src, m, R * CB, j0, m, m, tid);
            if (R != C) stage_block_async<AL16>(sB, p.src, m, C * CB, j0, m, m, tid);
        } else if (job == JOB_INV) {
            stage_block_async<AL16>(sB, p.src, m, C * CB, j0, m, m, tid);
            if (r < J) stage_block_async<true>(sA, p.Y, p.mp, j0, r * CB, p.mp, p.mp, tid);
        } else if (job == JOB_FIN) {
            stage_block_async<true>(sA, p.Y, p.mp, j0, r * CB, p.mp, p.mp, tid);
        }
        cp_async_commit();
        if (AL16) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                int e = tid + q * 256;
                *reinterpret_cast<double2*>(sD + (e >> 5) * CLD + (e & 31) * 2) = make_double2(v[2 * q], v[2 * q + 1]);
            }
        } else {
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                int e = tid + q * 256;
                sD[(e >> 6) * CLD + (e & 63)] = v[q];
            }
        }
    }
    double acc[4][2][2];
    if (job == JOB_TRAIL) acc_from_global<AL16>(acc, p.src, m, R * CB, C * CB, m, m, wm, wn, g, t);
    else if (job == JOB_INV && r < J) acc_from_global<true>(acc, p.Y, p.mp, C * CB, r * CB, p.mp, p.mp, wm, wn, g, t);
    else acc_zero(acc);
    __syncthreads();
    CHOL_STAMP(1);

    // ---- serial part: L_JJ and X = L_JJ^{-1} from 32x32 pieces
    bool bad = false;
    double logsum = 0.0;
    if (warp == 0) {
        logsum = factor32(sD, 0, rinv, colbuf, sLt, &bad);
        __syncwarp();
        CHOL_STAMP(2);
        invert32(sLt, 0, rinv, sX);
        CHOL_STAMP(3);
    } else {
        // zero the upper-right quadrants while warp 0 works
        for (int e = tid - 32; e < 32 * 32; e += 224) {
            int rr = e >> 5, cc = 32 + (e & 31);
            sX[rr * CLD + cc] = 0.0;
            sD[rr * CLD + cc] = 0.0;
        }
    }
    __syncthreads();
    const int sr = (warp >> 1) * 8 + g, sc = (warp & 1) * 16 + 2 * t;     // this thread's 32x32 fragment origin
    {   // L10 = D10 X00^T, then D11 -= L10 L10^T
        double o[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
        mma32<true, false>(o, sD + 32 * CLD, sX, warp, g, t);
        __syncthreads();                                              // every warp has read D10
#pragma unroll
        for (int j = 0; j < 2; ++j)
            *reinterpret_cast<double2*>(sD + (32 + sr) * CLD + sc + j * 8) = make_double2(o[j][0], o[j][1]);
        __syncthreads();
        double d[2][2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const double2 dd = *reinterpret_cast<const double2*>(sD + (32 + sr) * CLD + 32 + sc + j * 8);
            d[j][0] = dd.x; d[j][1] = dd.y;
        }
        mma32<true, true>(d, sD + 32 * CLD, sD + 32 * CLD, warp, g, t);
#pragma unroll
        for (int j = 0; j < 2; ++j)
            *reinterpret_cast<double2*>(sD + (32 + sr) * CLD + 32 + sc + j * 8) = make_double2(d[j][0], d[j][1]);
    }
    __syncthreads();
    CHOL_STAMP(4);
    if (warp == 0) {
        logsum += factor32(sD, 32, rinv, colbuf, sLt, &bad);
        __syncwarp();
        CHOL_STAMP(5);
        invert32(sLt, 32, rinv, sX);
    }
    __syncthreads();
    CHOL_STAMP(6);
    {   // X10 = -X11 (L10 X00); T = L10 X00 parks in the (still unused) lower-left quadrant of sP
        double o[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
        mma32<false, false>(o, sD + 32 * CLD, sX, warp, g, t);
#pragma unroll
        for (int j = 0; j < 2; ++j)
            *reinterpret_cast<double2*>(sP + (32 + sr) * CLD + sc + j * 8) = make_double2(o[j][0], o[j][1]);
        __syncthreads();
        double x[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
        mma32<false, true>(x, sX + 32 * CLD + 32, sP + 32 * CLD, warp, g, t);
#pragma unroll
        for (int j = 0; j < 2; ++j)
            *reinterpret_cast<double2*>(sX + (32 + sr) * CLD + sc + j * 8) = make_double2(x[j][0], x[j][1]);
    }
    cp_async_wait<0>();
    __syncthreads();
    CHOL_STAMP(7);

    // ---- one CTA publishes the diagonal results: the last one, whose own job is the lightest
    if (blockIdx.x == gridDim.x - 1) {
        if (p.L) smem_to_global(sD, p.L, m, j0, j0, m, m, tid);
        if (p.want_inv) smem_to_global(sX, p.Linv, p.mp, j0, j0, m, m, tid);      // the padding of Linv stays zero
        if (tid == 0) {
            double tot = (J == 0 ? 0.0 : p.logacc[0]) + logsum;
            p.logacc[0] = tot;
            if (J == p.nblk - 1) p.d_out[0] = -tot;
            if (bad) atomicOr(p.status, ACCBPG_ST_NOT_PD);
        }
    }
    CHOL_STAMP(8);
    if (job == JOB_NONE) { CHOL_STAMP_NS(15); return; }

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
        CHOL_STAMP(9);
        mma64<true, true>(acc, sP, pc, 0, CB, wm, wn, g, t);                  // W[R,C] -= L[R,J] L[C,J]^T
        CHOL_STAMP(10);
        acc_to_global<AL16>(acc, p.W, m, R * CB, C * CB, m, m, wm, wn, g, t);
        if (R == C && p.L) smem_to_global(sP, p.L, m, R * CB, j0, m, m, tid);
        CHOL_STAMP(11);
        CHOL_STAMP_NS(15);
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
        acc_to_global<true>(acc, p.Y, p.mp, C * CB, r * CB, p.mp, p.mp, wm, wn, g, t);
        return;
    }
    // JOB_FIN
    acc_zero(tmp);
    mma64<false, false>(tmp, sX, sA, 0, wm * 32 + 32, wm, wn, g, t);
    acc_to_global<true>(tmp, p.Linv, p.mp, j0, r * CB, p.mp, p.mp, wm, wn, g, t);
}


// ==================================================================================================================
// Data-flow form of the same chain: TWO launches for the whole factorisation + inverse instead of one per block column.
//
//   spine kernel (one CTA)   for J = 0 .. nb-1:  [J > 0: L[J,J-1] = W[J,J-1] X_{J-1}^T,  W[J,J] -= L[J,J-1] L[J,J-1]^T]
//                            then the serial part: L_JJ and X_J = L_JJ^{-1} (the same in-warp 32x32 pieces as above).
//                            The critical path of the factorisation never leaves this CTA's shared memory.
//   worker kernel            every other 64x64 tile job, claimed from a counter in topological order:
//        T(J,R)    L[R,J]   = W[R,J] X_J^T                      R > J+1
//        U(J;R,C)  W[R,C]  -= L[R,J] L[C,J]^T                   J < C <= R, (R,C) != (J+1,J+1)
//        F(J,r)    Linv[J,r] = X_J Y[J,r]                       r < J
//        V(K;C,r)  Y[C,r]  -= L[C,K] Linv[K,r]                  r <= K < C
//   Each job waits for its operands on per-tile progress counters in global memory (ld.acquire / st.release at GPU
//   scope): every tile's updates are applied in the fixed order J = 0, 1, 2, ..., so the result does not depend on
//   which CTA runs which job or when (bit-reproducible, and bit-identical to the launch-per-column kernel above, which
//   does the same arithmetic in the same order).  A job only ever waits for jobs with a lower number or for the spine,
//   and jobs are claimed in increasing order by CTAs that are running, so the scheme cannot deadlock whatever share of
//   the GPU the workers get; the worker launch is programmatically dependent on the spine launch, which releases it
//   with its first instruction, so the spine is resident before any worker can wait for it.
//   Counters carry the call's epoch in their upper bits (kept per aux block on the host; the last worker to leave
//   re-arms the job counter): no memset between calls.  The aux block and Linv must be zero when first used (the
//   workspace is zero-initialised once by its owner); Linv's strictly upper part is never written, so it stays zero
//   from call to call.  The triangular GEMM reads the same counters to start on row blocks of L^-1 as they become final.
#ifdef CHOL_TRACE
__device__ long long* g_df_trace_dev = nullptr;
#define DF_STAMP(i) do { if (threadIdx.x == 0 && g_df_trace_dev) g_df_trace_dev[df_trace_row * 16 + (i)] = clock64(); } while (0)
#else
#define DF_STAMP(i)
#endif
constexpr int DF_MAXNB = 64;
constexpr int DF_AUX_HEAD = 16;                 // ints: [1] job counter, [2] workers that have left, [3] spine done
constexpr int DF_WORKER_SMEM = 2 * CBUF * 8;
constexpr int DF_SPINE_SMEM = (6 * CBUF + 64 + 128) * 8;     // > half an SM's shared memory: the spine has its SM alone

struct DfParams {
    const double* M;     // input matrix (only read)
    double* W;           // trailing matrix / L panels, ld = m
    double* L;           // optional output
    double* Y;           // ld = mp
    double* Linv;        // ld = mp
    double* X;           // nb blocks of 64 x 64 (ld 64): X_J as published by the spine
    double* d_out;
    uint32_t* status;
    int* aux;            // DF_AUX_HEAD ints, then cW[nb*nb], cY[nb*nb], fLi[nb*nb], fX[nb]
    int m, mp, nb, want_inv, njobs, nworkers, epoch;
    int dbg;             // ACCBPG_DF_DBG (timing experiments only): 1 = the spine never waits and no workers run (results are garbage)
    int step_off[DF_MAXNB + 1];
};

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// wait until the counter at p carries this call's epoch and a count >= want; bounded (a protocol error traps after 2 s
// instead of hanging the GPU)
__device__ __forceinline__ void df_wait(const int* p, int epoch, int want) {
    if (want <= 0) return;
    unsigned long long t0 = 0;
    unsigned spin = 0;
    for (;;) {
        const int v = ld_acquire_gpu(p);
        if ((v >> 8) == epoch && (v & 255) >= want) return;
        if ((++spin & 0x3ffu) == 0) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > 2000000000ULL) __trap();
        }
    }
}
__device__ __forceinline__ void df_post(int* p, int epoch, int count) { st_release_gpu(p, (epoch << 8) | count); }

// 64 x 64 block with leading dimension 64 (the X buffer) -> smem [64][CLD]
__device__ __forceinline__ void stage_x_async(double* dst, const double* src, int tid) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        int e = tid + q * 256;
        int r = e >> 5, c = (e & 31) * 2;
        cp_async16(dst + r * CLD + c, src + r * CB + c, 16);
    }
}

// Tiles that other CTAs rewrite during the kernel must never be served from this SM's L1: the 16-byte cp.async is .cg,
// the unaligned path falls back to ld.global.cg + st.shared (the 8-byte cp.async exists only as .ca).
template <bool AL16>
__device__ __forceinline__ void df_stage(double* dst, const double* src, int64_t ld, int r0, int c0, int rows, int cols,
                                         int tid) {
    if (AL16) stage_block_async<true>(dst, src, ld, r0, c0, rows, cols, tid);
    else stage_block(dst, src, ld, r0, c0, rows, cols, tid);
}

// sub-tile product on the DMMA pipe: acc (MI_ x NJ_ mma tiles; acc[i][j][e] <-> row row0 + i*8 + g, col col0 + j*8 + 2t + e)
// (+/-)= A[row0 .., k] * B, k in [kbeg, kend); B_T: B stored [n][k] (product A B^T), else [k][n]
template <bool B_T, bool NEG, int MI_, int NJ_>
__device__ __forceinline__ void mma_sub(double (&acc)[MI_][NJ_][2], const double* A, int row0, const double* B, int col0,
                                        int kbeg, int kend, int g, int t) {
    const double* ap = A + (row0 + g) * CLD + t;
    const double* bp = B_T ? (B + (col0 + g) * CLD + t) : (B + t * CLD + col0 + g);
#pragma unroll 4
    for (int k = kbeg; k < kend; k += 4) {
        double a[MI_], b[NJ_];
#pragma unroll
        for (int i = 0; i < MI_; ++i) {
            a[i] = ap[i * 8 * CLD + k];
            if (NEG) a[i] = -a[i];
        }
#pragma unroll
        for (int j = 0; j < NJ_; ++j) b[j] = B_T ? bp[j * 8 * CLD + k] : bp[k * CLD + j * 8];
#pragma unroll
        for (int i = 0; i < MI_; ++i)
#pragma unroll
            for (int j = 0; j < NJ_; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
}
template <int MI_, int NJ_>
__device__ __forceinline__ void sub_load(double (&acc)[MI_][NJ_][2], const double* S, int row0, int col0, int g, int t) {
#pragma unroll
    for (int i = 0; i < MI_; ++i)
#pragma unroll
        for (int j = 0; j < NJ_; ++j) {
            const double2 v = *reinterpret_cast<const double2*>(S + (row0 + i * 8 + g) * CLD + col0 + j * 8 + 2 * t);
            acc[i][j][0] = v.x; acc[i][j][1] = v.y;
        }
}
template <int MI_, int NJ_>
__device__ __forceinline__ void sub_store(const double (&acc)[MI_][NJ_][2], double* S, int row0, int col0, int g, int t) {
#pragma unroll
    for (int i = 0; i < MI_; ++i)
#pragma unroll
        for (int j = 0; j < NJ_; ++j)
            *reinterpret_cast<double2*>(S + (row0 + i * 8 + g) * CLD + col0 + j * 8 + 2 * t) =
                make_double2(acc[i][j][0], acc[i][j][1]);
}
__device__ __forceinline__ void bar_sync_n(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive_n(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// The spine.  Within a step the critical path is  [P0 = rows 0..31 of L[J,J-1]] -> [D00 update] -> 32x32 factor + inverse ->
// 32x32 products -> 32x32 factor + inverse -> 32x32 products; everything else of the step runs on warps 1..7 under the
// first in-warp factorisation (rows 32..63 of the panel product, the other three quadrants of the diagonal update, the
// global stores of L[J,J-1] and of the previous X with their fences and counter posts) and warp 7 fetches the next
// step's two tiles into shared memory as soon as their counters allow (under the second factorisation).
template <bool AL16>
__global__ void __launch_bounds__(256, 1) chol_spine_kernel(DfParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* sD = reinterpret_cast<double*>(smem_raw);   // diagonal block -> L_JJ
    double* sX = sD + CBUF;                             // X_J (kept for the next step's panel product)
    double* sA0 = sX + CBUF;                            // W[J,J-1] -> L[J,J-1], ping
    double* sA1 = sA0 + CBUF;                           //                        pong (the next step's tile lands here)
    double* sDN = sA1 + CBUF;                           // W[J,J] as fetched (the diagonal tile before this step's update)
    double* sS = sDN + CBUF;                            // scratch of the serial part
    double* rinv = sS + CBUF;
    double* colbuf = rinv + 64;
    __shared__ int s_pf;                                // the next step's tiles have been requested (cp.async by warp 7)
    __shared__ volatile int s_w0;                       // serial phases warp 0 has finished (2J+1: first half, 2J+2: second)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int m = p.m, nb = p.nb;
    // the worker grid may start now (it never waits for this grid to finish, only for its counters)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int epoch = p.epoch;
    int* cW = p.aux + DF_AUX_HEAD;
    int* fX = cW + 3 * nb * nb;
    double logtot = 0.0;
    bool bad = false;
    if (tid == 0) { s_pf = 0; s_w0 = 0; }

    // warp 7: fetch the tiles of step Jn (panel tile (Jn,Jn-1) -> dstA, diagonal tile (Jn,Jn) -> sDN) as soon as their
    // counters allow, watching them only for as long as warp 0 is busy with serial phase `phase` (never holds the CTA up)
    auto try_prefetch = [&](int Jn, double* dstA, int phase) {
        if (Jn >= nb || s_pf) return;
        int ok = 0;
        if (lane == 0) {
            while (!ok) {
                const int need = (p.dbg & 1) ? 0 : Jn - 1;
                bool r1 = true, r2 = true;
                if (need > 0) {
                    const int v1 = ld_acquire_gpu(cW + Jn * nb + (Jn - 1)), v2 = ld_acquire_gpu(cW + Jn * nb + Jn);
                    r1 = (v1 >> 8) == epoch && (v1 & 255) >= need;
                    r2 = (v2 >> 8) == epoch && (v2 & 255) >= need;
                }
                ok = (r1 && r2) ? 1 : 0;
                if (ok || s_w0 >= phase) break;
                __nanosleep(40);
            }
        }
        ok = __shfl_sync(0xffffffffu, ok, 0);
        if (!ok) return;
        const double* src = (Jn == 1) ? p.M : p.W;
        if (AL16) {
#pragma unroll 8
            for (int q = 0; q < 64; ++q) {
                const int e = lane + q * 32;
                const int r = e >> 5, c = (e & 31) * 2;
                const int gr = Jn * CB + r;
                {
                    const int gc = (Jn - 1) * CB + c;
                    const bool in = gr < m && gc < m;
                    cp_async16(dstA + r * CLD + c, in ? (src + (int64_t)gr * m + gc) : src, in ? 16 : 0);
                }
                {
                    const int gc = Jn * CB + c;
                    const bool in = gr < m && gc < m;
                    cp_async16(sDN + r * CLD + c, in ? (src + (int64_t)gr * m + gc) : src, in ? 16 : 0);
                }
            }
        } else {
            for (int e = lane; e < CB * CB; e += 32) {
                const int r = e >> 6, c = e & 63;
                const int gr = Jn * CB + r, gc1 = (Jn - 1) * CB + c, gc2 = Jn * CB + c;
                dstA[r * CLD + c] = (gr < m && gc1 < m) ? __ldcg(src + (int64_t)gr * m + gc1) : 0.0;
                sDN[r * CLD + c] = (gr < m && gc2 < m) ? __ldcg(src + (int64_t)gr * m + gc2) : 0.0;
            }
        }
        cp_async_commit();
        if (lane == 0) s_pf = 1;
    };

    for (int J = 0; J < nb; ++J) {
        const int j0 = J * CB;
        const int df_trace_row = J;
        (void)df_trace_row;
        double* sA = (J & 1) ? sA1 : sA0;
        double* sAn = (J & 1) ? sA0 : sA1;
        DF_STAMP(0);
        if (J == 0) {
            df_stage<AL16>(sD, p.M, m, 0, 0, m, m, tid);
            cp_async_commit();
            cp_async_wait<0>();
            __syncthreads();
            for (int e = tid; e < 32 * 32; e += 256) {          // the upper right quadrants stay zero for the whole chain
                const int rr = e >> 5, cc = 32 + (e & 31);
                sX[rr * CLD + cc] = 0.0;
                sD[rr * CLD + cc] = 0.0;
            }
            __syncthreads();
        } else {
            // ---- the step's two tiles: already on their way (warp 7 asked for them during the last step), else now
            if (!s_pf) {
                const double* src = (J == 1) ? p.M : p.W;
                if (!(p.dbg & 1)) {
                    if (tid == 0) df_wait(cW + J * nb + (J - 1), epoch, J - 1);
                    if (tid == 32) df_wait(cW + J * nb + J, epoch, J - 1);
                }
                __syncthreads();
                df_stage<AL16>(sA, src, m, j0, j0 - CB, m, m, tid);
                df_stage<AL16>(sDN, src, m, j0, j0, m, m, tid);
                cp_async_commit();
            }
            cp_async_wait<0>();
            __syncthreads();
            DF_STAMP(1);
            {   // ---- P0: rows 0..31 of L[J,J-1] = W[J,J-1] X^T (X[n][k] = 0 for k > n), written over the tile in place
                const int wr = warp >> 2, wc = warp & 3;
                double a0[2][2][2] = {};
                mma_sub<true, false, 2, 2>(a0, sA, wr * 16, sX, wc * 16, 0, wc * 16 + 16, g, t);
                __syncthreads();
                sub_store<2, 2>(a0, sA, wr * 16, wc * 16, g, t);
                __syncthreads();
            }
            DF_STAMP(2);
            {   // ---- D00 = W[J,J][0:32,0:32] - P0 P0^T
                const int wr = warp >> 1, wc = warp & 1;
                double d0[1][2][2];
                sub_load<1, 2>(d0, sDN, wr * 8, wc * 16, g, t);
                mma_sub<true, true, 1, 2>(d0, sA, wr * 8, sA, wc * 16, 0, CB, g, t);
                sub_store<1, 2>(d0, sD, wr * 8, wc * 16, g, t);
            }
            __syncthreads();
            if (tid < 32 && j0 + tid >= m) sD[tid * CLD + tid] = 1.0;         // identity on the padding of a ragged last block
            __syncwarp();
        }
        DF_STAMP(3);
        // ---- warp 0: first in-warp factorisation; warps 1..7: the rest of the step's tile work
        double* sLt = sS;
        double lsum = 0.0;
        if (warp == 0) {
            lsum = factor32(sD, 0, rinv, colbuf, sLt, &bad);
            __syncwarp();
            DF_STAMP(4);
            if (J > 0) bar_sync_n(2, 256);              // warps 1..7 have finished reading the previous X
            invert32(sLt, 0, rinv, sX);
            if (lane == 0) s_w0 = 2 * J + 1;
            DF_STAMP(5);
        } else {
            const int w7 = warp - 1;
            if (J > 0) {
                // the previous step's X (and Linv / L diagonal blocks) were stored at its end: fence, then tell the workers
                __threadfence();
                bar_sync_n(3, 224);
                if (tid == 32) df_post(fX + (J - 1), epoch, 1);
                // P1: rows 32..63 of L[J,J-1], eight 16x16 sub-tiles over seven warps
                double a1[2][2][2][2] = {};
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int st = w7 + q * 7;
                    if (st < 8) {
                        const int r2 = st >> 2, c4 = st & 3;
                        mma_sub<true, false, 2, 2>(a1[q], sA, 32 + r2 * 16, sX, c4 * 16, 0, c4 * 16 + 16, g, t);
                    }
                }
                bar_arrive_n(2, 256);                   // done with the previous X: warp 0 may overwrite it
                bar_sync_n(3, 224);                     // every read of the tile's rows 32..63 is done
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int st = w7 + q * 7;
                    if (st < 8) sub_store<2, 2>(a1[q], sA, 32 + (st >> 2) * 16, (st & 3) * 16, g, t);
                }
                bar_sync_n(3, 224);
                // L[J,J-1] to global memory, then fence and post
                for (int e = tid - 32; e < CB * CB / 2; e += 224) {
                    const int r = e >> 5, c = (e & 31) * 2;
                    const int gr = j0 + r, gc = j0 - CB + c;
                    if (gr < m) {
                        const double2 v = *reinterpret_cast<const double2*>(sA + r * CLD + c);
                        if (AL16) {
                            *reinterpret_cast<double2*>(p.W + (int64_t)gr * m + gc) = v;
                            if (p.L) *reinterpret_cast<double2*>(p.L + (int64_t)gr * m + gc) = v;
                        } else {
                            p.W[(int64_t)gr * m + gc] = v.x; p.W[(int64_t)gr * m + gc + 1] = v.y;
                            if (p.L) { p.L[(int64_t)gr * m + gc] = v.x; p.L[(int64_t)gr * m + gc + 1] = v.y; }
                        }
                    }
                }
                __threadfence();
                bar_sync_n(3, 224);
                if (tid == 32) df_post(cW + J * nb + (J - 1), epoch, J);      // posted before the diagonal update: the workers' jobs on block column J start early
                // D10 (four 16x16 sub-tiles) and the lower three of D11: one per warp
                {
                    int r0, c0;
                    const double* Bp;
                    int brow;
                    if (w7 < 4) { r0 = 32 + (w7 >> 1) * 16; c0 = (w7 & 1) * 16; brow = c0; }
                    else { const int q = w7 - 4; const int r2 = q == 0 ? 0 : 1, c2 = q == 2 ? 1 : 0; r0 = 32 + r2 * 16; c0 = 32 + c2 * 16; brow = c0; }
                    Bp = sA;
                    double d1[2][2][2];
                    sub_load<2, 2>(d1, sDN, r0, c0, g, t);
                    mma_sub<true, true, 2, 2>(d1, sA, r0, Bp, brow, 0, CB, g, t);
                    sub_store<2, 2>(d1, sD, r0, c0, g, t);
                    if (w7 == 4) {                      // the strictly upper sub-tile of D11 is never used: keep it finite
                        double u[2][2][2];
                        sub_load<2, 2>(u, sDN, 32, 48, g, t);
                        sub_store<2, 2>(u, sD, 32, 48, g, t);
                    }
                }
                bar_sync_n(3, 224);                     // every read of sDN is done: warp 7 may refill it
                for (int e = tid - 32; e < 32; e += 224)
                    if (j0 + 32 + e >= m) sD[(32 + e) * CLD + 32 + e] = 1.0;          // ragged last block
            }
            if (warp == 7) {
                if (lane == 0) s_pf = 0;
                __syncwarp();
                try_prefetch(J + 1, sAn, 2 * J + 1);
            }
        }
        __syncthreads();
        const int sr = (warp >> 1) * 8 + g, sc = (warp & 1) * 16 + 2 * t;
        {   // L10 = D10 X00^T, then D11 -= L10 L10^T
            double o[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
            mma32<true, false>(o, sD + 32 * CLD, sX, warp, g, t);
            __syncthreads();
#pragma unroll
            for (int j = 0; j < 2; ++j)
                *reinterpret_cast<double2*>(sD + (32 + sr) * CLD + sc + j * 8) = make_double2(o[j][0], o[j][1]);
            __syncthreads();
            double d[2][2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const double2 dd = *reinterpret_cast<const double2*>(sD + (32 + sr) * CLD + 32 + sc + j * 8);
                d[j][0] = dd.x; d[j][1] = dd.y;
            }
            mma32<true, true>(d, sD + 32 * CLD, sD + 32 * CLD, warp, g, t);
#pragma unroll
            for (int j = 0; j < 2; ++j)
                *reinterpret_cast<double2*>(sD + (32 + sr) * CLD + 32 + sc + j * 8) = make_double2(d[j][0], d[j][1]);
        }
        __syncthreads();
        DF_STAMP(6);
        if (warp == 0) {
            lsum += factor32(sD, 32, rinv, colbuf, sLt, &bad);
            __syncwarp();
            DF_STAMP(7);
            invert32(sLt, 32, rinv, sX);
            if (lane == 0) s_w0 = 2 * J + 2;
            logtot += lsum;
        } else if (warp == 7) {
            try_prefetch(J + 1, sAn, 2 * J + 2);
        }
        __syncthreads();
        DF_STAMP(8);
        {   // X10 = -X11 (L10 X00); T = L10 X00 parks in the lower half of the scratch block
            double* sT = sS;
            double o[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
            mma32<false, false>(o, sD + 32 * CLD, sX, warp, g, t);
#pragma unroll
            for (int j = 0; j < 2; ++j)
                *reinterpret_cast<double2*>(sT + (32 + sr) * CLD + sc + j * 8) = make_double2(o[j][0], o[j][1]);
            __syncthreads();
            double x[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
            mma32<false, true>(x, sX + 32 * CLD + 32, sT + 32 * CLD, warp, g, t);
#pragma unroll
            for (int j = 0; j < 2; ++j)
                *reinterpret_cast<double2*>(sX + (32 + sr) * CLD + sc + j * 8) = make_double2(x[j][0], x[j][1]);
        }
        __syncthreads();
        DF_STAMP(9);
        // ---- warps 1..7 store X_J (workers read it from the X buffer; Linv[J,J] = X_J) and L_JJ; the fence and the post
        //      follow at the start of the next step, under its first factorisation
        if (warp > 0) {
            double* xdst = p.X + (size_t)J * CB * CB;
            for (int e = tid - 32; e < CB * CB / 2; e += 224) {
                const int r = e >> 5, c = (e & 31) * 2;
                const double2 v = *reinterpret_cast<const double2*>(sX + r * CLD + c);
                *reinterpret_cast<double2*>(xdst + r * CB + c) = v;
                const int gr = j0 + r, gc = j0 + c;
                if (p.want_inv && gr < m) {
                    if (gc + 1 < m) *reinterpret_cast<double2*>(p.Linv + (int64_t)gr * p.mp + gc) = v;
                    else if (gc < m) p.Linv[(int64_t)gr * p.mp + gc] = v.x;
                }
                if (p.L && gr < m) {
                    const double2 l = *reinterpret_cast<const double2*>(sD + r * CLD + c);
                    if (gc < m) p.L[(int64_t)gr * m + gc] = l.x;
                    if (gc + 1 < m) p.L[(int64_t)gr * m + gc + 1] = l.y;
                }
            }
        }
        DF_STAMP(10);
    }
    if (warp > 0) {
        __threadfence();
        bar_sync_n(3, 224);
        if (tid == 32) df_post(fX + (nb - 1), epoch, 1);
        bar_arrive_n(2, 256);
    } else {
        bar_sync_n(2, 256);                             // the last X has been posted
        if (tid == 0) {
            p.d_out[0] = -logtot;
            if (bad) atomicOr(p.status, ACCBPG_ST_NOT_PD);
            __threadfence();
            df_post(p.aux + 3, epoch, 1);               // the last worker leaves only after this: the pair of launches is complete
        }
    }
}

template <bool AL16>
__global__ void __launch_bounds__(256, 2) chol_worker_kernel(DfParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* sA = reinterpret_cast<double*>(smem_raw);
    double* sB = sA + CBUF;
    __shared__ int s_job[4];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp >> 2, wn = warp & 3;
    const int m = p.m, mp = p.mp, nb = p.nb;
    const int epoch = p.epoch;
    int* cW = p.aux + DF_AUX_HEAD;
    int* cY = cW + nb * nb;
    int* fLi = cY + nb * nb;
    int* fX = fLi + nb * nb;
    enum { JOB_T = 0, JOB_U, JOB_F, JOB_V };
    for (;;) {
        if (tid == 0) {
            const int jb = atomicAdd(p.aux + 1, 1);
            int type = -1, a = 0, b = 0, c = 0;
            if (jb < p.njobs) {
                int J = 0;
                while (p.step_off[J + 1] <= jb) ++J;
                int i = jb - p.step_off[J];
                const int below = nb - 1 - J;
                const int nT = below > 1 ? below - 1 : 0;
                const int nF = p.want_inv ? J : 0;
                const int nV1 = (p.want_inv && below >= 1) ? J + 1 : 0;
                const int nU2 = below > 1 ? (below - 1) * below / 2 : 0;
                if (i < nT) { type = JOB_T; a = J; b = J + 2 + i; }
                else if ((i -= nT) < nT) { type = JOB_U; a = J; b = J + 2 + i; c = J + 1; }          // column J+1
                else if ((i -= nT) < nF) { type = JOB_F; a = J; b = i; }
                else if ((i -= nF) < nV1) { type = JOB_V; a = J; b = J + 1; c = i; }                  // row block J+1
                else if ((i -= nV1) < nU2) {                                                          // columns >= J+2
                    // column-major over the lower triangle of the (below-1) x (below-1) trailing blocks
                    int cc = 0, left = i, len = below - 1;
                    while (left >= len) { left -= len; --len; ++cc; }
                    type = JOB_U; a = J; c = J + 2 + cc; b = c + left;
                } else {                                                                              // V, rows >= J+2
                    i -= nU2;
                    type = JOB_V; a = J; b = J + 2 + i / (J + 1); c = i % (J + 1);
                }
            }
            s_job[0] = type; s_job[1] = a; s_job[2] = b; s_job[3] = c;
        }
        __syncthreads();
        const int type = s_job[0], ja = s_job[1], jb2 = s_job[2], jc = s_job[3];
        __syncthreads();
        if (type < 0) break;
        double acc[4][2][2];
        if (type == JOB_T) {
            const int J = ja, R = jb2;
            if (tid == 0) df_wait(fX + J, epoch, 1);
            if (tid == 32) df_wait(cW + R * nb + J, epoch, J);
            __syncthreads();
            df_stage<AL16>(sA, J == 0 ? p.M : p.W, m, R * CB, J * CB, m, m, tid);
            stage_x_async(sB, p.X + (size_t)J * CB * CB, tid);
            cp_async_commit();
            cp_async_wait<0>();
            __syncthreads();
            acc_zero(acc);
            mma64<true, false>(acc, sA, sB, 0, wn * 16 + 16, wm, wn, g, t);
            acc_to_global<AL16>(acc, p.W, m, R * CB, J * CB, m, m, wm, wn, g, t);
            if (p.L) acc_to_global<AL16>(acc, p.L, m, R * CB, J * CB, m, m, wm, wn, g, t);
            __threadfence();
            __syncthreads();
            if (tid == 0) df_post(cW + R * nb + J, epoch, J + 1);
        } else if (type == JOB_U) {
            const int J = ja, R = jb2, C = jc;
            if (tid == 0) df_wait(cW + R * nb + J, epoch, J + 1);
            if (tid == 32) df_wait(cW + C * nb + J, epoch, J + 1);
            if (tid == 64) df_wait(cW + R * nb + C, epoch, J);
            __syncthreads();
            df_stage<AL16>(sA, p.W, m, R * CB, J * CB, m, m, tid);
            if (R != C) df_stage<AL16>(sB, p.W, m, C * CB, J * CB, m, m, tid);
            cp_async_commit();
            acc_from_global<AL16>(acc, J == 0 ? p.M : p.W, m, R * CB, C * CB, m, m, wm, wn, g, t);
            cp_async_wait<0>();
            __syncthreads();
            mma64<true, true>(acc, sA, R != C ? sB : sA, 0, CB, wm, wn, g, t);
            acc_to_global<AL16>(acc, p.W, m, R * CB, C * CB, m, m, wm, wn, g, t);
            __threadfence();
            __syncthreads();
            if (tid == 0) df_post(cW + R * nb + C, epoch, J + 1);
        } else if (type == JOB_F) {
            const int J = ja, r = jb2;
            if (tid == 0) df_wait(fX + J, epoch, 1);
            if (tid == 32) df_wait(cY + J * nb + r, epoch, J - r);
            __syncthreads();
            stage_x_async(sA, p.X + (size_t)J * CB * CB, tid);
            stage_block_async<true>(sB, p.Y, mp, J * CB, r * CB, mp, mp, tid);
            cp_async_commit();
            cp_async_wait<0>();
            __syncthreads();
            acc_zero(acc);
            mma64<false, false>(acc, sA, sB, 0, wm * 32 + 32, wm, wn, g, t);      // Linv[J,r] = X Y[J,r]
            acc_to_global<true>(acc, p.Linv, mp, J * CB, r * CB, mp, mp, wm, wn, g, t);
            __threadfence();
            __syncthreads();
            if (tid == 0) df_post(fLi + J * nb + r, epoch, 1);
        } else {
            const int K = ja, C = jb2, r = jc;
            if (tid == 0) df_wait(cW + C * nb + K, epoch, K + 1);
            if (tid == 32) { if (r < K) df_wait(fLi + K * nb + r, epoch, 1); else df_wait(fX + K, epoch, 1); }
            if (tid == 64) df_wait(cY + C * nb + r, epoch, K - r);
            __syncthreads();
            df_stage<AL16>(sA, p.W, m, C * CB, K * CB, m, m, tid);
            if (r < K) stage_block_async<true>(sB, p.Linv, mp, K * CB, r * CB, mp, mp, tid);
            else stage_x_async(sB, p.X + (size_t)K * CB * CB, tid);
            cp_async_commit();
            if (r < K) acc_from_global<true>(acc, p.Y, mp, C * CB, r * CB, mp, mp, wm, wn, g, t);
            else acc_zero(acc);
            cp_async_wait<0>();
            __syncthreads();
            mma64<false, true>(acc, sA, sB, r < K ? 0 : wn * 16, CB, wm, wn, g, t);   // Y[C,r] -= L[C,K] Linv[K,r]
            acc_to_global<true>(acc, p.Y, mp, C * CB, r * CB, mp, mp, wm, wn, g, t);
            __threadfence();
            __syncthreads();
            if (tid == 0) df_post(cY + C * nb + r, epoch, K - r + 1);
        }
        __syncthreads();         // sA / sB are free again
    }
    // leave: the last worker re-arms the counters for the next call on this aux block
    if (tid == 0) {
        __threadfence();
        const int left = atomicAdd(p.aux + 2, 1);
        if (left == p.nworkers - 1) {
            df_wait(p.aux + 3, epoch, 1);      // the spine's last stores (d_out, status) are done
            p.aux[2] = 0;
            p.aux[1] = 0;
        }
    }
}

size_t chol_df_aux_bytes(int m) {
    const int nb = (m + CB - 1) / CB;
    size_t ints = DF_AUX_HEAD + (size_t)3 * nb * nb + nb;
    size_t a = (ints * 4 + 255) / 256 * 256;
    return a + (size_t)nb * CB * CB * 8;
}

bool chol_df_enabled(int m) {
    static int v = -1;
    if (v < 0) { const char* e = getenv("ACCBPG_CHOL_DF"); v = (e && e[0] == '0') ? 0 : 1; }
    return v == 1 && (m + CB - 1) / CB <= DF_MAXNB && m > CB;
}

static bool g_df_attr[kMaxDevices] = {};
static int ensure_df_attrs() {
    int dev = 0;
    ACCBPG_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < kMaxDevices && g_df_attr[dev]) return ACCBPG_OK;
    ACCBPG_CUDA(cudaFuncSetAttribute(chol_spine_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, DF_SPINE_SMEM));
    ACCBPG_CUDA(cudaFuncSetAttribute(chol_spine_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, DF_SPINE_SMEM));
    ACCBPG_CUDA(cudaFuncSetAttribute(chol_worker_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, DF_WORKER_SMEM));
    ACCBPG_CUDA(cudaFuncSetAttribute(chol_worker_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, DF_WORKER_SMEM));
    if (dev >= 0 && dev < kMaxDevices) g_df_attr[dev] = true;
    return ACCBPG_OK;
}

// aux: chol_df_aux_bytes(m) bytes, zero when first used.  Linv (when wanted) must have been zero when first used and
// only ever written by this function since.
int chol_factor_inv_df(Ctx* c, cudaStream_t s, int m, int mp, const double* M, double* L, int want_inv, double* Linv,
                       double* W, double* Y, void* aux, double* d_out, DfGate* gate_out) {
    int rc = ensure_df_attrs();
    if (rc) return rc;
    if (L) ACCBPG_CUDA(cudaMemsetAsync(L, 0, (size_t)m * m * sizeof(double), s));
    ProfScope ps(P_CHOL, s);
    DfParams p;
    p.M = M; p.W = W; p.L = L; p.Y = Y; p.Linv = Linv; p.d_out = d_out; p.status = c->d_status;
    p.m = m; p.mp = mp; p.want_inv = want_inv ? 1 : 0;
    const int nb = (m + CB - 1) / CB;
    p.nb = nb;
    p.aux = (int*)aux;
    {   // the epoch of this call on this aux block (1, 2, 3, ...: zero-filled memory never matches)
        static std::map<void*, int> epochs;
        int& e = epochs[aux];
        e = (e >= (1 << 22)) ? 1 : e + 1;
        p.epoch = e;
    }
    if (gate_out) {
        gate_out->fX = p.aux + DF_AUX_HEAD + 3 * nb * nb;
        gate_out->fLi = p.aux + DF_AUX_HEAD + 2 * nb * nb;
        gate_out->nb = nb;
        gate_out->epoch = p.epoch;
    }
    size_t ints = DF_AUX_HEAD + (size_t)3 * nb * nb + nb;
    p.X = (double*)((char*)aux + (ints * 4 + 255) / 256 * 256);
    int off = 0;
    for (int J = 0; J < nb; ++J) {
        p.step_off[J] = off;
        const int below = nb - 1 - J;
        const int nT = below > 1 ? below - 1 : 0;
        const int nF = want_inv ? J : 0;
        const int nV1 = (want_inv && below >= 1) ? J + 1 : 0;
        const int nU2 = below > 1 ? (below - 1) * below / 2 : 0;
        const int nV2 = (want_inv && below > 1) ? (below - 1) * (J + 1) : 0;
        off += nT + nT + nF + nV1 + nU2 + nV2;
    }
    for (int J = nb; J <= DF_MAXNB; ++J) p.step_off[J] = off;
    p.njobs = off;
    static int wcap = -1;
    if (wcap < 0) { const char* e = getenv("ACCBPG_CHOL_WORKERS"); wcap = e ? atoi(e) : 0; }
    // enough CTAs to keep one block column's jobs in flight, at most two per SM
    int widest = 0;
    for (int J = 0; J < nb; ++J) widest = max(widest, p.step_off[J + 1] - p.step_off[J]);
    int nworkers = widest + widest / 2 + 4;
    if (nworkers > 2 * (c->sm_count - 1)) nworkers = 2 * (c->sm_count - 1);
    if (wcap > 0 && nworkers > wcap) nworkers = wcap;
    if (nworkers > p.njobs) nworkers = p.njobs;
    if (nworkers < 1) nworkers = 1;
    p.nworkers = nworkers;
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("ACCBPG_DF_DBG"); dbg = e ? atoi(e) : 0; }
    p.dbg = dbg;
    auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
    const bool al16 = (m % 2 == 0) && al(M) && al(W) && (!L || al(L));
    if (al16) chol_spine_kernel<true><<<1, 256, DF_SPINE_SMEM, s>>>(p);
    else      chol_spine_kernel<false><<<1, 256, DF_SPINE_SMEM, s>>>(p);
    ACCBPG_LAUNCHED("chol_spine_kernel");
    if (dbg & 1) return ACCBPG_OK;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(nworkers, 1, 1);
    cfg.blockDim = dim3(256, 1, 1);
    cfg.dynamicSmemBytes = DF_WORKER_SMEM;
    cfg.stream = s;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (al16) ACCBPG_CUDA(cudaLaunchKernelEx(&cfg, chol_worker_kernel<true>, p));
    else      ACCBPG_CUDA(cudaLaunchKernelEx(&cfg, chol_worker_kernel<false>, p));
    ACCBPG_LAUNCHED("chol_worker_kernel");
    return ACCBPG_OK;
}

#ifdef CHOL_TRACE
long long* g_chol_trace = nullptr;
#endif
static bool g_chol_attr[kMaxDevices] = {};        // the attribute is per device
static int ensure_attrs() {
    int dev = 0;
    ACCBPG_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < kMaxDevices && g_chol_attr[dev]) return ACCBPG_OK;
    ACCBPG_CUDA(cudaFuncSetAttribute(chol_inv_step_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, CHOL_SMEM));
    ACCBPG_CUDA(cudaFuncSetAttribute(chol_inv_step_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, CHOL_SMEM));
    if (dev >= 0 && dev < kMaxDevices) g_chol_attr[dev] = true;
    return ACCBPG_OK;
}

// ------------------------------------------------------------------------------------------ host entry point
// M (m x m, symmetric, only read) -> d_out[0] = -log det M;  L (may be NULL) <- lower factor;  when want_inv:
// Linv (mp x mp, ld mp) <- L^{-1}, zero above the diagonal and on rows / columns >= m.
// W: m x m scratch (trailing matrix), Y: mp x mp scratch (running sums), acc: device scalar.
// Launches are chained with programmatic dependent launch: launch J+1 is resident and past its prologue when
// launch J retires, so the ~4 us launch gap drops out of the critical path.
int chol_factor_inv(Ctx* c, cudaStream_t s, int m, int mp, const double* M, double* L, int want_inv, double* Linv,
                    double* W, double* Y, double* acc, double* d_out, const std::function<int(int)>* after_step, void* aux,
                    DfGate* gate) {
    if (aux && !after_step && chol_df_enabled(m))
        return chol_factor_inv_df(c, s, m, mp, M, L, want_inv, Linv, W, Y, aux, d_out, gate);
    int rc = ensure_attrs();
    if (rc) return rc;
    if (L) ACCBPG_CUDA(cudaMemsetAsync(L, 0, (size_t)m * m * sizeof(double), s));
    if (want_inv) ACCBPG_CUDA(cudaMemsetAsync(Linv, 0, (size_t)mp * mp * sizeof(double), s));
    ProfScope ps(P_CHOL, s);
    CholStep p;
    p.W = W; p.L = L; p.Y = Y; p.Linv = Linv; p.logacc = acc; p.d_out = d_out; p.status = c->d_status;
    p.m = m; p.mp = mp; p.want_inv = want_inv;
    p.nblk = (m + CB - 1) / CB;
#ifdef CHOL_TRACE
    p.trace = g_chol_trace;
#endif
    auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
    const bool al16 = (m % 2 == 0) && al(M) && al(W) && (!L || al(L));
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(256, 1, 1);
    cfg.dynamicSmemBytes = CHOL_SMEM;
    cfg.stream = s;
    cfg.attrs = attr;
    bool hooked = false;
    for (int J = 0; J < p.nblk; ++J) {
        const int below = p.nblk - 1 - J;
        p.J = J;
        p.src = (J == 0) ? M : W;
        p.nT = below * (below + 1) / 2;
        p.nI = want_inv ? (J + 1) * below : 0;
        p.nF = want_inv ? J : 0;
        int grid = p.nT + p.nI + p.nF;
        if (grid < 1) grid = 1;
        cfg.gridDim = dim3(grid, 1, 1);
        cfg.numAttrs = (J > 0 && !hooked) ? 1 : 0;      // plain ordering after memsets / event records / other work
        if (al16) ACCBPG_CUDA(cudaLaunchKernelEx(&cfg, chol_inv_step_kernel<true>, p));
        else      ACCBPG_CUDA(cudaLaunchKernelEx(&cfg, chol_inv_step_kernel<false>, p));
        ACCBPG_LAUNCHED("chol_inv_step_kernel");
        hooked = false;
        if (after_step) {                        // rows of block column J of L^{-1} are final from here on
            int hr = (*after_step)(J);
            if (hr < 0) return -hr;
            hooked = hr > 0;                     // the hook put something on the stream: the next launch is not programmatic
        }
    }
    return ACCBPG_OK;
}

}  // namespace accbpg
