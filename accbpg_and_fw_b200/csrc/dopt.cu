// D-optimal design objective on sm_100a:  f(x) = -log det(H diag(x) H^T),  g_j = -h_j^T M^{-1} h_j
// (reference: DOptimalObj.func_grad, accbpg/functions.py:43-59).
//
//   K1  gram     M = H diag(x) H^T     FP64 DMMA (mma.sync m8n8k4) SYRK, lower 128x128 tiles, split over the
//                                      n (column) dimension, partials reduced in a fixed order (deterministic)
//   K2+K3 factor M = L L^T, L^{-1}     chol.cu: right-looking 64-block Cholesky whose launches also carry the block
//                                      forward substitution for L^{-1}; -log det from the pivots
//   K4  grad     g_j = -||Linv h_j||^2 DMMA triangular GEMM Linv*H streamed over column panels of H; the epilogue
//                                      squares and column-reduces the 128x128 tile, so M^{-1}H never exists in HBM
//
// Blackwell's tcgen05 tensor path has no FP64 kind, so FP64 tensor work is issued as warp-level DMMA.8x8x4
// (every f64 mma.sync shape lowers to that SASS on sm_100a) fed from shared memory through a 4-stage cp.async pipeline.
#include "dmma.cuh"

namespace accbpg {

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

    // Per-thread copy descriptors of the fast path (aligned, k-slab fully inside the chunk): the same 16-byte column
    // chunk of rows lr, lr+32, lr+64, lr+96 of the A block and of the B block.  Pointers advance by BK per k-slab, so
    // a stage costs 8-9 cp.async and no address arithmetic beyond one add each.
    const int ch = tid & 7, lr = tid >> 3;
    const double* ga[4];
    const double* gb[4];
    int ba[4], bb[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int ra = bi * BM + lr + 32 * i, rb = bj * BN + lr + 32 * i;
        ba[i] = ra < p.m ? 16 : 0;
        bb[i] = rb < p.m ? 16 : 0;
        ga[i] = p.H + (int64_t)(ba[i] ? ra : 0) * p.ldh + k_begin + ch * 2;
        gb[i] = p.H + (int64_t)(bb[i] ? rb : 0) * p.ldh + k_begin + ch * 2;
    }
    const int soff = lr * A_LD + ch * 2;
    const int KT_full = (k_end > k_begin) ? (int)((k_end - k_begin) / BK) : 0;

    auto load_stage = [&](int stage, int kt) {
        double* As = smem + stage * KMAJOR_STAGE_DOUBLES;
        double* Bs = As + BM * A_LD;
        double* Xs = Bs + BN * A_LD;
        if (ALIGNED16 && kt < KT_full) {
            const int64_t ko = (int64_t)kt * BK;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                cp_async16(As + soff + i * 32 * A_LD, ga[i] + ko, ba[i]);
                cp_async16(Bs + soff + i * 32 * A_LD, gb[i] + ko, bb[i]);
            }
            if (tid < 8) cp_async16(Xs + tid * 2, p.x + k_begin + ko + tid * 2, 16);
            return;
        }
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
        // The two warps that share an SM sub-partition (wm = 0 / 1) issue their share of the next stage's copies at
        // different points of the k-slab, so one of them is always feeding the DMMA pipe.
        const int nk = kt + STAGES - 1;
        if (wm == 0 && nk < KT) load_stage(nk % STAGES, nk);
        const double* As = smem + (kt % STAGES) * KMAJOR_STAGE_DOUBLES;
        const double* Bs = As + BM * A_LD;
        const double* Xs = Bs + BN * A_LD;
        const double* ap = As + (wm * 64 + g) * A_LD + t;
        const double* bp = Bs + (wn * 32 + g) * A_LD + t;
        // fragments double buffered in registers: LDS (+ the diag(x) scaling) of step kk+1 run under the DMMAs of kk
        double a[2][MI], b[2][NI];
        {
            const double xv = Xs[t];
            neg |= (xv < 0.0);
#pragma unroll
            for (int i = 0; i < MI; ++i) a[0][i] = ap[i * 8 * A_LD];
#pragma unroll
            for (int j = 0; j < NI; ++j) b[0][j] = bp[j * 8 * A_LD] * xv;
        }
#pragma unroll
        for (int kk = 0; kk < BK / 4; ++kk) {
            const int cur = kk & 1, nxt = cur ^ 1;
            if (kk == 2 && wm == 1 && nk < KT) load_stage(nk % STAGES, nk);
            if (kk + 1 < BK / 4) {
                const double xv = Xs[(kk + 1) * 4 + t];
                neg |= (xv < 0.0);
#pragma unroll
                for (int i = 0; i < MI; ++i) a[nxt][i] = ap[i * 8 * A_LD + (kk + 1) * 4];
#pragma unroll
                for (int j = 0; j < NI; ++j) b[nxt][j] = bp[j * 8 * A_LD + (kk + 1) * 4] * xv;
            }
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NI; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[cur][i], b[cur][j]);
        }
        cp_async_commit();
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

// ------------------------------------------------------------------------------------------ K4: gradient
struct TrmmParams {
    const double* Linv;   // [mp][mp], zero above the diagonal (identity on the padded diagonal)
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
    int kmax = (ib + 1) * BM;                         // Linv is lower triangular: k <= row
    const int m16 = (p.m + BK - 1) / BK * BK;
    if (kmax > m16) kmax = m16;
    const int KT = kmax / BK;

    double acc[MI][NI][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    auto load_stage = [&](int stage, int kt) {
        double* As = smem + stage * NN_STAGE_DOUBLES;
        double* Bs = As + BM * A_LD;
        // A: Linv rows ib*128.., cols k0..k0+15 (padded buffer: always in range, 16-byte aligned)
        load_kmajor_slab<true>(As, p.Linv, p.mp, ib * BM, p.mp, (int64_t)kt * BK, p.mp, tid);
        // B: H rows k0..k0+15 (rows >= m read as zero), cols j0..j0+127
        load_nmajor_slab<ALIGNED16>(Bs, p.H, p.ldh, kt * BK, p.m, j0, p.n, tid);
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
        const double* As = smem + (kt % STAGES) * NN_STAGE_DOUBLES;
        mma_nn_slab(acc, As, As + BM * A_LD, wm, wn, g, t);
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
    size_t off_P, off_Linv, off_Y, off_part, off_M, off_L, off_W, off_M2, off_W2, total;
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
    // layout: the m-only buffers first (their offsets do not depend on n_local), then the n-dependent ones
    const size_t mm = ((size_t)m * m * 8 + 255) / 256 * 256;
    size_t a = 0;
    pl.off_M = a;    a += mm;
    pl.off_L = a;    a += mm;
    pl.off_W = a;    a += mm;      // trailing matrix of the factorisation
    pl.off_M2 = a;   a += mm;      // second set: the value-only evaluation that runs on the side stream
    pl.off_W2 = a;   a += mm;
    pl.off_Linv = a; a += (size_t)pl.mp * pl.mp * 8;
    pl.off_Y = a;    a += (size_t)pl.mp * pl.mp * 8;      // running sums of the block forward substitution
    pl.off_P = a;    a += (size_t)pl.splits * pl.mp * pl.mp * 8;
    pl.off_part = a; a += (size_t)pl.nib * pl.npad * 8;
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
    ACCBPG_CUDA(cudaFuncSetAttribute(syrk_dmma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, KMAJOR_SMEM));
    ACCBPG_CUDA(cudaFuncSetAttribute(syrk_dmma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, KMAJOR_SMEM));
    ACCBPG_CUDA(cudaFuncSetAttribute(trmm_colnorm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, NN_SMEM));
    ACCBPG_CUDA(cudaFuncSetAttribute(trmm_colnorm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, NN_SMEM));
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
        if (al) syrk_dmma_kernel<true><<<grid, GEMM_THREADS, KMAJOR_SMEM, s>>>(p);
        else    syrk_dmma_kernel<false><<<grid, GEMM_THREADS, KMAJOR_SMEM, s>>>(p);
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

int accbpg_dopt_factor(void* ctx, void* stream, int m, const double* M, double* L, int want_inverse, void* ws,
                       double* d_out) {
    Ctx* c = (Ctx*)ctx;
    cudaStream_t s = (cudaStream_t)stream;
    if (!c || !M || !ws || !d_out) return arg_err("dopt_factor: NULL pointer");
    if (m < 1) return arg_err("dopt_factor: m");
    if (M == L) return arg_err("dopt_factor: in-place factorisation is not supported");
    // the factor scratch lives in the m-only head of the workspace (independent of n_local)
    DoptPlan pl = make_plan(m, 2, c->sm_count);
    double* W = (double*)((char*)ws + pl.off_W);
    if (M == W || L == W) return arg_err("dopt_factor: M / L alias the factor scratch");
    return chol_factor_inv(c, s, m, pl.mp, M, L, want_inverse ? 1 : 0, (double*)((char*)ws + pl.off_Linv), W,
                           (double*)((char*)ws + pl.off_Y), c->d_slots + 248, d_out);
}

int accbpg_dopt_grad(void* ctx, void* stream, const double* H, int m, int64_t n, int64_t ldh, void* ws, double* g) {
    Ctx* c = (Ctx*)ctx;
    cudaStream_t s = (cudaStream_t)stream;
    if (!c || !H || !ws || !g) return arg_err("dopt_grad: NULL pointer");
    if (m < 1 || n < 1 || ldh < n) return arg_err("dopt_grad: shape");
    int rc = ensure_smem_attrs();
    if (rc) return rc;
    DoptPlan pl = make_plan(m, n, c->sm_count);
    double* Linv = (double*)((char*)ws + pl.off_Linv);      // left there by accbpg_dopt_factor(want_inverse = 1)
    double* part = (double*)((char*)ws + pl.off_part);
    TrmmParams p;
    p.Linv = Linv; p.H = H; p.part = part; p.m = m; p.mp = pl.mp; p.nib = pl.nib;
    p.n = n; p.ldh = ldh; p.npad = pl.npad;
    int64_t npanels = (n + BN - 1) / BN;
    if (npanels > 65535) return arg_err("dopt_grad: n_local too large for one launch (max 65535*128 columns)");
    dim3 grid(pl.nib, (unsigned)npanels);
    bool al = aligned16(H) && (ldh % 2 == 0);
    {
        ProfScope ps(P_TRMM, s);
        if (al) trmm_colnorm_kernel<true><<<grid, GEMM_THREADS, NN_SMEM, s>>>(p);
        else    trmm_colnorm_kernel<false><<<grid, GEMM_THREADS, NN_SMEM, s>>>(p);
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
    int rc = accbpg_dopt_gram(ctx, stream, H, m, n, ldh, x, ws, M);
    if (rc) return rc;
    rc = accbpg_dopt_factor(ctx, stream, m, M, NULL, flag >= 1, ws, d_f_out ? d_f_out : (c->d_slots + 249));
    if (rc) return rc;
    if (flag >= 1) rc = accbpg_dopt_grad(ctx, stream, H, m, n, ldh, ws, g);
    return rc;
}

// f(xf) and (f(yg), grad f(yg)) in one call: the two Gram matrices are formed back to back on the main stream, then
// the value-only Cholesky runs on the context's side stream while the main stream factors M(yg), inverts L and
// streams the gradient.  Both chains are latency bound on a handful of SMs, so they overlap almost completely.
int accbpg_dopt_pair(void* ctx, void* stream, const double* H, int m, int64_t n, int64_t ldh, const double* xf,
                     const double* yg, int flag_y, void* ws, double* d_fx_out, double* d_fy_out, double* g) {
    Ctx* c = (Ctx*)ctx;
    cudaStream_t s = (cudaStream_t)stream;
    if (!c || !ws || !xf || !yg || !d_fx_out) return arg_err("dopt_pair: NULL pointer");
    if (flag_y < 1 || flag_y > 2 || !g) return arg_err("dopt_pair: flag_y must be 1 or 2 with a gradient buffer");
    DoptPlan pl = make_plan(m, n, c->sm_count);
    double* M1 = (double*)((char*)ws + pl.off_M);
    double* M2 = (double*)((char*)ws + pl.off_M2);
    int rc = accbpg_dopt_gram(ctx, stream, H, m, n, ldh, xf, ws, M2);
    if (rc) return rc;
    ACCBPG_CUDA(cudaEventRecord(c->ev_fork, s));
    ACCBPG_CUDA(cudaStreamWaitEvent(c->side, c->ev_fork, 0));
    rc = chol_factor_inv(c, c->side, m, pl.mp, M2, NULL, 0, NULL, (double*)((char*)ws + pl.off_W2), NULL,
                         c->d_slots + 246, d_fx_out);
    if (rc) return rc;
    ACCBPG_CUDA(cudaEventRecord(c->ev_join, c->side));
    rc = accbpg_dopt_gram(ctx, stream, H, m, n, ldh, yg, ws, M1);
    if (rc) return rc;
    rc = accbpg_dopt_factor(ctx, stream, m, M1, NULL, 1, ws, d_fy_out ? d_fy_out : (c->d_slots + 249));
    if (rc) return rc;
    rc = accbpg_dopt_grad(ctx, stream, H, m, n, ldh, ws, g);
    if (rc) return rc;
    ACCBPG_CUDA(cudaStreamWaitEvent(s, c->ev_join, 0));
    return ACCBPG_OK;
}

// Same, starting from Gram matrices the caller already holds (M is linear in x, so the drivers can form
// M((1-t)x + t z) = (1-t)M(x) + t M(z) instead of running another SYRK).  Mx may be NULL (gradient side only).
int accbpg_dopt_pair_from_gram(void* ctx, void* stream, const double* H, int m, int64_t n, int64_t ldh,
                               const double* Mx, const double* My, int flag_y, void* ws, double* d_fx_out,
                               double* d_fy_out, double* g) {
    Ctx* c = (Ctx*)ctx;
    cudaStream_t s = (cudaStream_t)stream;
    if (!c || !ws || !My) return arg_err("dopt_pair_from_gram: NULL pointer");
    if (flag_y < 0 || flag_y > 2 || (flag_y >= 1 && !g)) return arg_err("dopt_pair_from_gram: flag_y / gradient buffer");
    if (Mx && !d_fx_out) return arg_err("dopt_pair_from_gram: d_fx_out is NULL");
    DoptPlan pl = make_plan(m, n, c->sm_count);
    int rc;
    if (Mx) {
        ACCBPG_CUDA(cudaEventRecord(c->ev_fork, s));
        ACCBPG_CUDA(cudaStreamWaitEvent(c->side, c->ev_fork, 0));
        rc = chol_factor_inv(c, c->side, m, pl.mp, Mx, NULL, 0, NULL, (double*)((char*)ws + pl.off_W2), NULL,
                             c->d_slots + 246, d_fx_out);
        if (rc) return rc;
        ACCBPG_CUDA(cudaEventRecord(c->ev_join, c->side));
    }
    rc = accbpg_dopt_factor(ctx, stream, m, My, NULL, flag_y >= 1, ws, d_fy_out ? d_fy_out : (c->d_slots + 249));
    if (rc) return rc;
    if (flag_y >= 1) {
        rc = accbpg_dopt_grad(ctx, stream, H, m, n, ldh, ws, g);
        if (rc) return rc;
    }
    if (Mx) ACCBPG_CUDA(cudaStreamWaitEvent(s, c->ev_join, 0));
    return ACCBPG_OK;
}

}  // extern "C"
