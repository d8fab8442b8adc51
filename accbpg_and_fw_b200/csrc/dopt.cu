// D-optimal design objective on sm_100a:  f(x) = -log det(H diag(x) H^T),  g_j = -h_j^T M^{-1} h_j
// (reference: DOptimalObj.func_grad, accbpg/functions.py:43-59).
//
//   K1  gram     M = H diag(x) H^T     FP64 DMMA (mma.sync m8n8k4) SYRK, lower 128x128 tiles, split over the
//                                      n (column) dimension, partials reduced in a fixed order (deterministic)
//   K2  factor   M = L L^T             left-looking blocked Cholesky (32-wide panels), -log det from the pivots
//   K3  trinv    Linv = L^{-1}         diagonal 32x32 blocks inverted in-warp, then one CTA per block column
//   K4  grad     g_j = -||Linv h_j||^2 DMMA triangular GEMM Linv*H streamed over column panels of H; the epilogue
//                                      squares and column-reduces the 128x128 tile, so M^{-1}H never exists in HBM
//
// Blackwell's tcgen05 tensor path has no FP64 kind, so FP64 tensor work is issued as warp-level DMMA.8x8x4
// (every f64 mma.sync shape lowers to that SASS on sm_100a) fed from shared memory through a 4-stage cp.async pipeline.
#include "common.cuh"

namespace accbpg {

// ------------------------------------------------------------------------------------------ tile geometry
constexpr int BM = 128;            // CTA tile rows
constexpr int BN = 128;            // CTA tile cols
constexpr int BK = 16;             // k-slab per pipeline stage
constexpr int STAGES = 4;
constexpr int GEMM_THREADS = 256;  // 8 warps: 2 (rows) x 4 (cols); warp tile 64 x 32 = 8 x 4 DMMA tiles
constexpr int A_LD = BK + 4;       // 20 doubles: (g*20 + t) mod 16 distinct over a half warp -> conflict-free LDS.64
constexpr int BT_LD = BN + 4;      // 132 doubles for the k-major B tile of the triangular GEMM
constexpr int MI = 8, NI = 4;

constexpr int SYRK_STAGE_DOUBLES = BM * A_LD + BN * A_LD + BK;
constexpr int SYRK_SMEM = STAGES * SYRK_STAGE_DOUBLES * 8;
constexpr int TRMM_STAGE_DOUBLES = BM * A_LD + BK * BT_LD;
constexpr int TRMM_SMEM = STAGES * TRMM_STAGE_DOUBLES * 8;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(sa), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, int src_bytes) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(sa), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// Copy a [rows x 16] slab (k contiguous in global memory) into smem rows of A_LD doubles, zero-filling rows
// >= row_limit and columns >= k_limit.  256 threads, 128 rows.
template <bool ALIGNED16>
__device__ __forceinline__ void load_kmajor_slab(double* dst, const double* src, int64_t ld, int row0, int row_limit,
                                                 int64_t k0, int64_t k_limit, int tid) {
    if (ALIGNED16) {
#pragma unroll
        for (int i = 0; i < (BM * BK / 2) / GEMM_THREADS; ++i) {
            int c = tid + i * GEMM_THREADS;
            int row = c >> 3, ch = c & 7;
            int grow = row0 + row;
            int64_t gcol = k0 + ch * 2;
            int64_t left = (k_limit - gcol) * 8;
            int bytes = (grow < row_limit && left > 0) ? (left >= 16 ? 16 : 8) : 0;
            const double* s = bytes ? (src + (int64_t)grow * ld + gcol) : src;
            cp_async16(dst + row * A_LD + ch * 2, s, bytes);
        }
    } else {
#pragma unroll
        for (int i = 0; i < (BM * BK) / GEMM_THREADS; ++i) {
            int c = tid + i * GEMM_THREADS;
            int row = c >> 4, col = c & 15;
            int grow = row0 + row;
            int64_t gcol = k0 + col;
            int bytes = (grow < row_limit && gcol < k_limit) ? 8 : 0;
            const double* s = bytes ? (src + (int64_t)grow * ld + gcol) : src;
            cp_async8(dst + row * A_LD + col, s, bytes);
        }
    }
}

// ------------------------------------------------------------------------------------------ K1: SYRK
struct SyrkParams {
    const double* H;
    const double* x;
    double* P;          // [splits][mp][mp] partial tiles (lower tiles only are written)
    int m, mp, splits;
    int64_t n, ldh, kchunk;
    uint32_t* status;
};

template <bool ALIGNED16>
__global__ void __launch_bounds__(GEMM_THREADS, 1) syrk_dmma_kernel(SyrkParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* smem = reinterpret_cast<double*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp >> 2, wn = warp & 3;

    // lower-triangular tile index -> (bi, bj), bi >= bj
    int tt = blockIdx.x;
    int bi = (int)((sqrt(8.0 * tt + 1.0) - 1.0) * 0.5);
    while ((bi + 1) * (bi + 2) / 2 <= tt) ++bi;
    while (bi * (bi + 1) / 2 > tt) --bi;
    const int bj = tt - bi * (bi + 1) / 2;

    const int64_t k_begin = (int64_t)blockIdx.y * p.kchunk;
    int64_t k_end = k_begin + p.kchunk;
    if (k_end > p.n) k_end = p.n;
    const int KT = (k_end > k_begin) ? (int)((k_end - k_begin + BK - 1) / BK) : 0;

    double acc[MI][NI][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    auto load_stage = [&](int stage, int kt) {
        double* As = smem + stage * SYRK_STAGE_DOUBLES;
        double* Bs = As + BM * A_LD;
        double* Xs = Bs + BN * A_LD;
        const int64_t k0 = k_begin + (int64_t)kt * BK;
        load_kmajor_slab<ALIGNED16>(As, p.H, p.ldh, bi * BM, p.m, k0, k_end, tid);
        load_kmajor_slab<ALIGNED16>(Bs, p.H, p.ldh, bj * BN, p.m, k0, k_end, tid);
        if (ALIGNED16) {
            if (tid < 8) {
                int64_t gk = k0 + tid * 2;
                int64_t left = (k_end - gk) * 8;
                int bytes = left > 0 ? (left >= 16 ? 16 : 8) : 0;
                cp_async16(Xs + tid * 2, bytes ? (p.x + gk) : p.x, bytes);
            }
        } else {
            if (tid < 16) {
                int64_t gk = k0 + tid;
                int bytes = gk < k_end ? 8 : 0;
                cp_async8(Xs + tid, bytes ? (p.x + gk) : p.x, bytes);
            }
        }
    };

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < KT) load_stage(s, s);
        cp_async_commit();
    }

    bool neg = false;
    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            int nk = kt + STAGES - 1;
            if (nk < KT) load_stage(nk % STAGES, nk);
            cp_async_commit();
        }
        const double* As = smem + (kt % STAGES) * SYRK_STAGE_DOUBLES;
        const double* Bs = As + BM * A_LD;
        const double* Xs = Bs + BN * A_LD;
        const double* ap = As + (wm * 64 + g) * A_LD + t;
        const double* bp = Bs + (wn * 32 + g) * A_LD + t;
#pragma unroll
        for (int kk = 0; kk < BK / 4; ++kk) {
            const double xv = Xs[kk * 4 + t];
            neg |= (xv < 0.0);
            double a[MI], b[NI];
#pragma unroll
            for (int i = 0; i < MI; ++i) a[i] = ap[i * 8 * A_LD + kk * 4];
#pragma unroll
            for (int j = 0; j < NI; ++j) b[j] = bp[j * 8 * A_LD + kk * 4] * xv;
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NI; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    cp_async_wait<0>();
    if (neg) atomicOr(p.status, ACCBPG_ST_X_NEGATIVE);

    // partial tile -> workspace (mp is a multiple of 128: no bounds checks, 16-byte stores)
    double* P = p.P + (size_t)blockIdx.y * p.mp * p.mp;
#pragma unroll
    for (int i = 0; i < MI; ++i) {
        int row = bi * BM + wm * 64 + i * 8 + g;
#pragma unroll
        for (int j = 0; j < NI; ++j) {
            int col = bj * BN + wn * 32 + j * 8 + 2 * t;
            double2 v = make_double2(acc[i][j][0], acc[i][j][1]);
            *reinterpret_cast<double2*>(P + (size_t)row * p.mp + col) = v;
        }
    }
}

// M[i][j] = M[j][i] = sum_s P[s][i][j]  (i >= j), splits added in index order
__global__ void __launch_bounds__(256) syrk_reduce_kernel(const double* P, int splits, int m, int mp, double* M) {
    int j = blockIdx.x * 32 + (threadIdx.x & 31);
    int i = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (i >= m || j > i) return;
    const size_t off = (size_t)i * mp + j, sz = (size_t)mp * mp;
    double s = 0.0;
    for (int k = 0; k < splits; ++k) s += __ldcg(P + k * sz + off);
    M[(size_t)i * m + j] = s;
    M[(size_t)j * m + i] = s;
}

// ------------------------------------------------------------------------------------------ K2: Cholesky
// Left-looking, 32-wide panels, one launch per panel, out of place: M (symmetric) is only read, the factor goes to
// L, so the CTAs of one launch never read what another CTA of the same launch writes.  CTA b owns rows R_b = [j0+32+64b, +64) and recomputes the
// 32x32 diagonal block redundantly:   [D; S] = M[J u R_b, J] - L[J u R_b, 0:j0] L[J, 0:j0]^T,  D = Ld Ld^T (warp 0,
// rows in registers, shuffles),  X = S Ld^{-T}.  CTA 0 stores Ld and adds sum log(pivot) to the log-det slot.
constexpr int CH_NB = 32;
constexpr int CH_ROWS = 64;
constexpr int CH_TR = CH_NB + CH_ROWS;   // 96 rows per CTA in the update GEMM
constexpr int CH_LDT = CH_TR + 1;        // transposed k-slab [32][97]

__global__ void __launch_bounds__(256) chol_panel_kernel(const double* __restrict__ M, double* L, int m, int j0,
                                                         double* logdet_acc, uint32_t* status) {
    __shared__ double Lt[CH_NB][CH_LDT];          // Lt[kk][row]: rows 0..31 = J, 32..95 = R_b
    __shared__ double Ld[CH_NB][CH_NB + 1];       // factored diagonal block
    __shared__ double rinv[CH_NB];
    const int tid = threadIdx.x;
    const int ty = tid >> 3, tx = tid & 7;        // 32 x 8
    const int nb = min(CH_NB, m - j0);
    const int r0 = j0 + CH_NB + blockIdx.x * CH_ROWS;

    // global row of tile row r (0..95); -1 when outside the matrix
    auto grow = [&](int r) -> int {
        int gr = (r < CH_NB) ? (j0 + r) : (r0 + r - CH_NB);
        if (r < CH_NB && r >= nb) return -1;
        return (gr < m) ? gr : -1;
    };

    double acc[3][4];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[a][c] = 0.0;

    for (int k0 = 0; k0 < j0; k0 += CH_NB) {      // j0 is a multiple of 32
        __syncthreads();
        // load 96 x 32 slab, coalesced along k, stored transposed
        for (int e = tid; e < CH_TR * CH_NB; e += 256) {
            int r = e >> 5, kk = e & 31;
            int gr = grow(r);
            Lt[kk][r] = (gr >= 0) ? L[(size_t)gr * m + k0 + kk] : 0.0;
        }
        __syncthreads();
#pragma unroll 8
        for (int kk = 0; kk < CH_NB; ++kk) {
            double a0 = Lt[kk][ty], a1 = Lt[kk][ty + 32], a2 = Lt[kk][ty + 64];
            double b[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) b[c] = Lt[kk][tx * 4 + c];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                acc[0][c] = fma(a0, b[c], acc[0][c]);
                acc[1][c] = fma(a1, b[c], acc[1][c]);
                acc[2][c] = fma(a2, b[c], acc[2][c]);
            }
        }
    }
    __syncthreads();
    // S = M - acc for the 96 x 32 tile; stage it in Lt as St[col][row]  (same shape)
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        int r = ty + 32 * a;
        int gr = grow(r);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            int col = tx * 4 + c;
            double v = 0.0;
            if (gr >= 0 && col < nb) v = M[(size_t)gr * m + j0 + col] - acc[a][c];
            if (r < CH_NB && (gr < 0 || col >= nb)) v = (r == col) ? 1.0 : 0.0;   // identity padding of a short panel
            Lt[col][r] = v;
        }
    }
    __syncthreads();

    // ---- factor the diagonal block in warp 0: lane r holds row r
    if (tid < 32) {
        const int lane = tid;
        double row[CH_NB];
#pragma unroll
        for (int c = 0; c < CH_NB; ++c) row[c] = Lt[c][lane];
        double logsum = 0.0;
        bool bad = false;
#pragma unroll
        for (int k = 0; k < CH_NB; ++k) {
            double pkk = __shfl_sync(0xffffffffu, row[k], k);
            if (!(pkk > 0.0)) { bad = true; pkk = 1.0; }
            if (lane == k) logsum = log(pkk);
            double ri = rsqrt(pkk);
            double lrk = row[k] * ri;
            row[k] = lrk;
            if (lane == k) rinv[k] = 1.0 / lrk;
#pragma unroll
            for (int c = k + 1; c < CH_NB; ++c) {
                double lck = __shfl_sync(0xffffffffu, lrk, c);
                row[c] = fma(-lrk, lck, row[c]);
            }
        }
#pragma unroll
        for (int c = 0; c < CH_NB; ++c) Ld[lane][c] = (c <= lane) ? row[c] : 0.0;
        logsum = warp_sum(logsum);
        if (blockIdx.x == 0) {
            if (lane == 0) logdet_acc[0] += logsum;        // single writer, launches are stream ordered
            if (bad && lane == 0) atomicOr(status, ACCBPG_ST_NOT_PD);
#pragma unroll
            for (int c = 0; c < CH_NB; ++c)
                if (lane < nb && c <= lane) L[(size_t)(j0 + lane) * m + j0 + c] = row[c];
        }
    }
    __syncthreads();

    // ---- X = S Ld^{-T}: thread r (< 64) solves one row, right-looking so the 32 steps expose ILP
    if (tid < CH_ROWS) {
        int gr = r0 + tid;
        if (gr < m) {
            double sv[CH_NB];
#pragma unroll
            for (int c = 0; c < CH_NB; ++c) sv[c] = Lt[c][CH_NB + tid];
#pragma unroll
            for (int k = 0; k < CH_NB; ++k) {
                double xk = sv[k] * rinv[k];
                sv[k] = xk;
#pragma unroll
                for (int c = k + 1; c < CH_NB; ++c) sv[c] = fma(-xk, Ld[c][k], sv[c]);
            }
#pragma unroll
            for (int c = 0; c < CH_NB; ++c)
                if (c < nb) L[(size_t)gr * m + j0 + c] = sv[c];
        }
    }
}

__global__ void store_neg_kernel(const double* src, double* dst) { dst[0] = -src[0]; }

// ------------------------------------------------------------------------------------------ K3: L^{-1}
// (a) invert every 32x32 diagonal block: lane c solves column c of  Ljj X = I  (right-looking substitution)
__global__ void __launch_bounds__(32) trinv_diag_kernel(const double* L, int m, double* Linv, int mp) {
    __shared__ double Ls[CH_NB][CH_NB + 1];
    const int j0 = blockIdx.x * CH_NB, lane = threadIdx.x;
    const int nb = min(CH_NB, m - j0);
    for (int r = 0; r < CH_NB; ++r) {
        double v = (r == lane) ? 1.0 : 0.0;
        if (r < nb && lane < nb && lane <= r) v = L[(size_t)(j0 + r) * m + j0 + lane];
        Ls[r][lane] = v;
    }
    __syncwarp();
    double b[CH_NB];
#pragma unroll
    for (int r = 0; r < CH_NB; ++r) b[r] = (r == lane) ? 1.0 : 0.0;
#pragma unroll
    for (int k = 0; k < CH_NB; ++k) {
        double xk = b[k] / Ls[k][k];
        b[k] = xk;
#pragma unroll
        for (int r = k + 1; r < CH_NB; ++r) b[r] = fma(-Ls[r][k], xk, b[r]);
    }
#pragma unroll
    for (int r = 0; r < CH_NB; ++r)
        if (r < nb && lane < nb && lane <= r) Linv[(size_t)(j0 + r) * mp + j0 + lane] = b[r];
}

// (b) block column j of Linv, top to bottom:  X_ij = -Linv_ii * sum_{k=j}^{i-1} L_ik X_kj   (one CTA per j)
__global__ void __launch_bounds__(256) trinv_cols_kernel(const double* L, int m, double* Linv, int mp) {
    __shared__ double Ak[CH_NB][CH_NB + 1];   // L_ik   [r][kk]
    __shared__ double Xk[CH_NB][CH_NB + 1];   // X_kj   [kk][c]
    __shared__ double Tt[CH_NB][CH_NB + 1];   // T      [r][c]
    const int nblk = (m + CH_NB - 1) / CH_NB;
    const int j = blockIdx.x, tid = threadIdx.x;
    const int r = tid >> 3, c4 = (tid & 7) * 4;        // thread: row r, cols c4..c4+3
    for (int i = j + 1; i < nblk; ++i) {
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        for (int k = j; k < i; ++k) {
            __syncthreads();
            for (int e = tid; e < CH_NB * CH_NB; e += 256) {
                int rr = e >> 5, cc = e & 31;
                int gi = i * CH_NB + rr, gk = k * CH_NB + cc;
                Ak[rr][cc] = (gi < m && gk < m) ? L[(size_t)gi * m + gk] : 0.0;
                int gkr = k * CH_NB + rr, gj = j * CH_NB + cc;
                Xk[rr][cc] = Linv[(size_t)gkr * mp + gj];       // zero padded, written earlier by this CTA
            }
            __syncthreads();
#pragma unroll 8
            for (int kk = 0; kk < CH_NB; ++kk) {
                double a = Ak[r][kk];
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[c] = fma(a, Xk[kk][c4 + c], acc[c]);
            }
        }
        __syncthreads();
#pragma unroll
        for (int c = 0; c < 4; ++c) Tt[r][c4 + c] = acc[c];
        for (int e = tid; e < CH_NB * CH_NB; e += 256) {
            int rr = e >> 5, cc = e & 31;
            Ak[rr][cc] = Linv[(size_t)(i * CH_NB + rr) * mp + i * CH_NB + cc];   // Linv_ii (zero above diag)
        }
        __syncthreads();
        double o[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 8
        for (int kk = 0; kk < CH_NB; ++kk) {
            double a = Ak[r][kk];
#pragma unroll
            for (int c = 0; c < 4; ++c) o[c] = fma(a, Tt[kk][c4 + c], o[c]);
        }
        int gi = i * CH_NB + r;
        if (gi < m) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                int gj = j * CH_NB + c4 + c;
                if (gj < m) Linv[(size_t)gi * mp + gj] = -o[c];
            }
        }
        __threadfence_block();
    }
}

// ------------------------------------------------------------------------------------------ K4: gradient
struct TrmmParams {
    const double* Linv;   // [mp][mp], zero above the diagonal and in the padding
    const double* H;
    double* part;         // [nib][npad] partial column sums of squares
    int m, mp, nib;
    int64_t n, ldh, npad;
};

template <bool ALIGNED16>
__global__ void __launch_bounds__(GEMM_THREADS, 1) trmm_colnorm_kernel(TrmmParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* smem = reinterpret_cast<double*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp >> 2, wn = warp & 3;
    const int ib = p.nib - 1 - blockIdx.x;            // heavy row blocks first
    const int64_t j0 = (int64_t)blockIdx.y * BN;
    int kmax = (ib + 1) * BM;
    const int m16 = (p.m + BK - 1) / BK * BK;
    if (kmax > m16) kmax = m16;
    const int KT = kmax / BK;

    double acc[MI][NI][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    auto load_stage = [&](int stage, int kt) {
        double* As = smem + stage * TRMM_STAGE_DOUBLES;
        double* Bs = As + BM * A_LD;
        const int k0 = kt * BK;
        // A: Linv rows ib*128.., cols k0..k0+15 (padded buffer: always in range, 16-byte aligned)
#pragma unroll
        for (int i = 0; i < (BM * BK / 2) / GEMM_THREADS; ++i) {
            int c = tid + i * GEMM_THREADS;
            int row = c >> 3, ch = c & 7;
            cp_async16(As + row * A_LD + ch * 2, p.Linv + (size_t)(ib * BM + row) * p.mp + k0 + ch * 2, 16);
        }
        // B: H rows k0..k0+15, cols j0..j0+127
        if (ALIGNED16) {
#pragma unroll
            for (int i = 0; i < (BK * BN / 2) / GEMM_THREADS; ++i) {
                int c = tid + i * GEMM_THREADS;
                int kr = c >> 6, ch = c & 63;
                int gk = k0 + kr;
                int64_t gcol = j0 + ch * 2;
                int64_t left = (p.n - gcol) * 8;
                int bytes = (gk < p.m && left > 0) ? (left >= 16 ? 16 : 8) : 0;
                const double* s = bytes ? (p.H + (int64_t)gk * p.ldh + gcol) : p.H;
                cp_async16(Bs + kr * BT_LD + ch * 2, s, bytes);
            }
        } else {
#pragma unroll
            for (int i = 0; i < (BK * BN) / GEMM_THREADS; ++i) {
                int c = tid + i * GEMM_THREADS;
                int kr = c >> 7, col = c & 127;
                int gk = k0 + kr;
                int64_t gcol = j0 + col;
                int bytes = (gk < p.m && gcol < p.n) ? 8 : 0;
                const double* s = bytes ? (p.H + (int64_t)gk * p.ldh + gcol) : p.H;
                cp_async8(Bs + kr * BT_LD + col, s, bytes);
            }
        }
    };

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < KT) load_stage(s, s);
        cp_async_commit();
    }
    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            int nk = kt + STAGES - 1;
            if (nk < KT) load_stage(nk % STAGES, nk);
            cp_async_commit();
        }
        const double* As = smem + (kt % STAGES) * TRMM_STAGE_DOUBLES;
        const double* Bs = As + BM * A_LD;
        const double* ap = As + (wm * 64 + g) * A_LD + t;
        const double* bp = Bs + t * BT_LD + wn * 32 + g;
#pragma unroll
        for (int kk = 0; kk < BK / 4; ++kk) {
            double a[MI], b[NI];
#pragma unroll
            for (int i = 0; i < MI; ++i) a[i] = ap[i * 8 * A_LD + kk * 4];
#pragma unroll
            for (int j = 0; j < NI; ++j) b[j] = bp[kk * 4 * BT_LD + j * 8];
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NI; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    cp_async_wait<0>();
    __syncthreads();      // pipeline buffers are dead: reuse the front of smem for the column exchange

    // epilogue: column sums of squares over this tile's 128 rows, fixed order
    double* colx = smem;                              // [2][128]
#pragma unroll
    for (int j = 0; j < NI; ++j) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            double s = 0.0;
#pragma unroll
            for (int i = 0; i < MI; ++i) s = fma(acc[i][j][e], acc[i][j][e], s);
            s += __shfl_xor_sync(0xffffffffu, s, 4);
            s += __shfl_xor_sync(0xffffffffu, s, 8);
            s += __shfl_xor_sync(0xffffffffu, s, 16);
            if (g == 0) colx[wm * BN + wn * 32 + j * 8 + 2 * t + e] = s;
        }
    }
    __syncthreads();
    if (tid < BN) {
        int64_t col = j0 + tid;
        if (col < p.n) p.part[(size_t)ib * p.npad + col] = colx[tid] + colx[BN + tid];
    }
}

__global__ void __launch_bounds__(256) grad_finalize_kernel(const double* part, int nib, int64_t npad, int64_t n,
                                                            double* g) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
        double s = 0.0;
        for (int b = 0; b < nib; ++b) s += __ldcg(part + (size_t)b * npad + j);
        g[j] = -s;
    }
}

// ------------------------------------------------------------------------------------------ host side plan
struct DoptPlan {
    int mp, nt, ntri, nib, splits;
    int64_t kchunk, npad;
    size_t off_P, off_Linv, off_part, off_M, off_L, total;
};

static DoptPlan make_plan(int m, int64_t n, int sm_count) {
    DoptPlan pl;
    pl.nt = (m + BM - 1) / BM;
    pl.mp = pl.nt * BM;
    pl.ntri = pl.nt * (pl.nt + 1) / 2;
    pl.nib = pl.nt;
    pl.npad = (n + 1) / 2 * 2;
    int64_t ktiles = (n + BK - 1) / BK;
    int64_t max_splits = ktiles / 64;                 // at least 64 k-slabs (1024 columns) per CTA
    if (max_splits < 1) max_splits = 1;
    if (max_splits > 64) max_splits = 64;
    int best = 1;
    double best_eff = 0.0;
    for (int w = 1; w <= 2; ++w) {
        int64_t s = (int64_t)sm_count * w / pl.ntri;
        if (s < 1) s = 1;
        if (s > max_splits) s = max_splits;
        int64_t ctas = (int64_t)pl.ntri * s;
        int64_t waves = (ctas + sm_count - 1) / sm_count;
        double eff = (double)ctas / (double)(waves * sm_count);
        if (eff > best_eff + 0.02) { best_eff = eff; best = (int)s; }
    }
    pl.splits = best;
    int64_t per = (ktiles + pl.splits - 1) / pl.splits;
    pl.kchunk = per * BK;
    size_t a = 0;
    pl.off_P = a;    a += (size_t)pl.splits * pl.mp * pl.mp * 8;
    pl.off_Linv = a; a += (size_t)pl.mp * pl.mp * 8;
    pl.off_part = a; a += (size_t)pl.nib * pl.npad * 8;
    pl.off_M = a;    a += ((size_t)m * m * 8 + 255) / 256 * 256;
    pl.off_L = a;    a += ((size_t)m * m * 8 + 255) / 256 * 256;
    pl.total = a;
    return pl;
}

static int device_sm_count() {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}

static bool g_attr_done = false;
static int ensure_smem_attrs() {
    if (g_attr_done) return ACCBPG_OK;
    ACCBPG_CUDA(cudaFuncSetAttribute(syrk_dmma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SYRK_SMEM));
    ACCBPG_CUDA(cudaFuncSetAttribute(syrk_dmma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SYRK_SMEM));
    ACCBPG_CUDA(cudaFuncSetAttribute(trmm_colnorm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TRMM_SMEM));
    ACCBPG_CUDA(cudaFuncSetAttribute(trmm_colnorm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TRMM_SMEM));
    g_attr_done = true;
    return ACCBPG_OK;
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

size_t dopt_linv_offset(int m, int64_t n, int sm_count, int* mp_out) {
    DoptPlan pl = make_plan(m, n, sm_count);
    if (mp_out) *mp_out = pl.mp;
    return pl.off_Linv;
}

}  // namespace accbpg

using namespace accbpg;

extern "C" {

size_t accbpg_dopt_workspace_bytes(int m, int64_t n_local) {
    if (m < 1 || n_local < 1) return 0;
    return make_plan(m, n_local, device_sm_count()).total;
}

int accbpg_dopt_gram(void* ctx, void* stream, const double* H, int m, int64_t n, int64_t ldh, const double* x,
                     void* ws, double* M) {
    Ctx* c = (Ctx*)ctx;
    cudaStream_t s = (cudaStream_t)stream;
    if (!c || !H || !x || !ws || !M) return arg_err("dopt_gram: NULL pointer");
    if (m < 1 || n < 1 || ldh < n) return arg_err("dopt_gram: shape");
    int rc = ensure_smem_attrs();
    if (rc) return rc;
    DoptPlan pl = make_plan(m, n, c->sm_count);
    SyrkParams p;
    p.H = H; p.x = x; p.P = (double*)((char*)ws + pl.off_P);
    p.m = m; p.mp = pl.mp; p.splits = pl.splits; p.n = n; p.ldh = ldh; p.kchunk = pl.kchunk;
    p.status = c->d_status;
    dim3 grid(pl.ntri, pl.splits);
    bool al = aligned16(H) && aligned16(x) && (ldh % 2 == 0);
    {
        ProfScope ps(P_SYRK, s);
        if (al) syrk_dmma_kernel<true><<<grid, GEMM_THREADS, SYRK_SMEM, s>>>(p);
        else    syrk_dmma_kernel<false><<<grid, GEMM_THREADS, SYRK_SMEM, s>>>(p);
    }
    ACCBPG_LAUNCHED("syrk_dmma_kernel");
    dim3 rg((m + 31) / 32, (m + 7) / 8);
    {
        ProfScope ps(P_SYRK_REDUCE, s);
        syrk_reduce_kernel<<<rg, 256, 0, s>>>(p.P, pl.splits, m, pl.mp, M);
    }
    ACCBPG_LAUNCHED("syrk_reduce_kernel");
    return ACCBPG_OK;
}

int accbpg_dopt_factor(void* ctx, void* stream, int m, const double* M, double* L, double* d_out) {
    Ctx* c = (Ctx*)ctx;
    cudaStream_t s = (cudaStream_t)stream;
    if (!c || !M || !L || !d_out) return arg_err("dopt_factor: NULL pointer");
    if (m < 1) return arg_err("dopt_factor: m");
    if (M == L) return arg_err("dopt_factor: in-place factorisation is not supported");
    double* acc = c->d_slots + 248;                  // running sum of log pivots
    ACCBPG_CUDA(cudaMemsetAsync(acc, 0, sizeof(double), s));
    ACCBPG_CUDA(cudaMemsetAsync(L, 0, (size_t)m * m * sizeof(double), s));
    {
        ProfScope ps(P_CHOL, s);
        for (int j0 = 0; j0 < m; j0 += CH_NB) {
            int below = m - j0 - CH_NB;
            int grid = below > 0 ? (below + CH_ROWS - 1) / CH_ROWS : 1;
            chol_panel_kernel<<<grid, 256, 0, s>>>(M, L, m, j0, acc, c->d_status);
            ACCBPG_LAUNCHED("chol_panel_kernel");
        }
    }
    store_neg_kernel<<<1, 1, 0, s>>>(acc, d_out);
    ACCBPG_LAUNCHED("store_neg_kernel");
    return ACCBPG_OK;
}

int accbpg_dopt_grad(void* ctx, void* stream, const double* H, int m, int64_t n, int64_t ldh, const double* L,
                     void* ws, double* g) {
    Ctx* c = (Ctx*)ctx;
    cudaStream_t s = (cudaStream_t)stream;
    if (!c || !H || !L || !ws || !g) return arg_err("dopt_grad: NULL pointer");
    if (m < 1 || n < 1 || ldh < n) return arg_err("dopt_grad: shape");
    int rc = ensure_smem_attrs();
    if (rc) return rc;
    DoptPlan pl = make_plan(m, n, c->sm_count);
    double* Linv = (double*)((char*)ws + pl.off_Linv);
    double* part = (double*)((char*)ws + pl.off_part);
    ACCBPG_CUDA(cudaMemsetAsync(Linv, 0, (size_t)pl.mp * pl.mp * 8, s));
    int nblk = (m + CH_NB - 1) / CH_NB;
    {
        ProfScope ps(P_TRINV, s);
        trinv_diag_kernel<<<nblk, 32, 0, s>>>(L, m, Linv, pl.mp);
        ACCBPG_LAUNCHED("trinv_diag_kernel");
        if (nblk > 1) {
            trinv_cols_kernel<<<nblk - 1, 256, 0, s>>>(L, m, Linv, pl.mp);
            ACCBPG_LAUNCHED("trinv_cols_kernel");
        }
    }
    TrmmParams p;
    p.Linv = Linv; p.H = H; p.part = part; p.m = m; p.mp = pl.mp; p.nib = pl.nib;
    p.n = n; p.ldh = ldh; p.npad = pl.npad;
    int64_t npanels = (n + BN - 1) / BN;
    if (npanels > 65535) return arg_err("dopt_grad: n_local too large for one launch (max 65535*128 columns)");
    dim3 grid(pl.nib, (unsigned)npanels);
    bool al = aligned16(H) && (ldh % 2 == 0);
    {
        ProfScope ps(P_TRMM, s);
        if (al) trmm_colnorm_kernel<true><<<grid, GEMM_THREADS, TRMM_SMEM, s>>>(p);
        else    trmm_colnorm_kernel<false><<<grid, GEMM_THREADS, TRMM_SMEM, s>>>(p);
    }
    ACCBPG_LAUNCHED("trmm_colnorm_kernel");
    int fg = grid_for(c, n, 256, 2, 8);
    {
        ProfScope ps(P_GRAD_FIN, s);
        grad_finalize_kernel<<<fg, 256, 0, s>>>(part, pl.nib, pl.npad, n, g);
    }
    ACCBPG_LAUNCHED("grad_finalize_kernel");
    return ACCBPG_OK;
}

int accbpg_dopt_func_grad(void* ctx, void* stream, const double* H, int m, int64_t n, int64_t ldh, const double* x,
                          int flag, void* ws, double* d_f_out, double* g) {
    Ctx* c = (Ctx*)ctx;
    if (!c || !ws) return arg_err("dopt_func_grad: NULL pointer");
    if (flag < 0 || flag > 2) return arg_err("dopt_func_grad: flag");
    if (flag >= 1 && !g) return arg_err("dopt_func_grad: gradient buffer is NULL");
    DoptPlan pl = make_plan(m, n, c->sm_count);
    double* M = (double*)((char*)ws + pl.off_M);
    double* L = (double*)((char*)ws + pl.off_L);
    int rc = accbpg_dopt_gram(ctx, stream, H, m, n, ldh, x, ws, M);
    if (rc) return rc;
    rc = accbpg_dopt_factor(ctx, stream, m, M, L, d_f_out ? d_f_out : (c->d_slots + 249));
    if (rc) return rc;
    if (flag >= 1) rc = accbpg_dopt_grad(ctx, stream, H, m, n, ldh, L, ws, g);
    return rc;
}

}  // extern "C"
