// LinearSolver_Cholesky on the device (reference: SolverFunctions.cs:332-444 — ALGLIB's
// sparseconverttosks + sparsecholeskyskyline(isupper) + sparsecholeskysolvesks).
//
// The reference factorises the upper skyline in place, one scalar column at a time.  On the GPU the
// same envelope is kept, but in dense 64x64 blocks: block column J stores block rows F[J]..J
// contiguously ("block skyline"), F made non-decreasing so that block row K's stored blocks are the
// contiguous range K..E[K].  Factorisation A = U^T U is right-looking over block rows:
//     panel(K):   U_KK = chol(A_KK);  U_KJ = U_KK^-T A_KJ          for J in (K, E[K]]
//     update(K):  A_IJ -= U_KI^T U_KJ                              for K < I <= J <= E[K]
// which is FP64-FMA bound (n * bandwidth^2 flops), not HBM bound like the CG path.  U^T y = b is
// carried through the factorisation as one more column; U x = y is one sweep over the block columns
// in descending order.  SPC-fixed DOFs stay in the system as
// identity rows (as in the CG path), so eliminating them only adds exact zeros; the tail of the last
// block is padded with identity rows.  Every sum has a fixed order, so results are reproducible.
// Single GPU: a skyline that does not fit one device is beyond what a direct solver is for here.
#include "common.cuh"
#include "bulk.cuh"

namespace stan {

namespace {

constexpr int CB = 64;           // dense block edge (scalar DOFs)
constexpr int CBB = CB * CB;     // doubles per block
constexpr int ERR_NOT_SPD = 5;   // slot in d_err

struct BandDev {
    double *band;
    double *diag;                // factored diagonal blocks U_KK, one 64x64 block per block row (see k_chol_panel)
    const int32_t *F;            // first stored block row of block column J
    const int64_t *pofs;         // blocks before block column J
    __device__ __forceinline__ double *blk(int I, int J) const {
        return band + (pofs[J] + (int64_t)(I - F[J])) * CBB;
    }
};

// smallest column node of every block row == first row of the node's three skyline columns / 3
__global__ void k_min_col(int64_t n_nodes, const int32_t *__restrict__ brow_ptr, const int32_t *__restrict__ bcol,
                          int32_t *__restrict__ minnb) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    int32_t m = (int32_t)i;
    for (int s = brow_ptr[i]; s < brow_ptr[i + 1]; s++) m = min(m, bcol[s]);
    minnb[i] = m;
}

// upper triangle of the assembled block rows -> block skyline; fixed DOFs become identity rows/columns
__global__ void k_band_scatter(int64_t n_nodes, const int32_t *__restrict__ brow_ptr, const int32_t *__restrict__ bcol,
                               const double *__restrict__ vals, const uint8_t *__restrict__ fixed, BandDev B) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= 3 * n_nodes) return;
    const int64_t i = t / 3;
    const int a = (int)(t % 3);
    const int64_t r = t;
    const int s0 = brow_ptr[i], s1 = brow_ptr[i + 1], nb = s1 - s0;
    const double *rowv = vals + 9 * (int64_t)s0 + (int64_t)a * 3 * nb;
    const bool rfix = fixed[r] != 0;
    const int I = (int)(r / CB), rr = (int)(r % CB);
    for (int s = s0; s < s1; s++) {
        const int64_t q = bcol[s];
        for (int b = 0; b < 3; b++) {
            const int64_t c = 3 * q + b;
            if (c < r) continue;
            double v = rowv[3 * (s - s0) + b];
            if (rfix || fixed[c]) v = (c == r) ? 1.0 : 0.0;
            const int J = (int)(c / CB);
            B.blk(I, J)[rr * CB + (int)(c % CB)] = v;
        }
    }
}

__global__ void k_band_pad(int64_t n, int64_t npad, BandDev B) {
    int64_t r = n + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= npad) return;
    const int J = (int)(r / CB), rr = (int)(r % CB);
    B.blk(J, J)[rr * CB + rr] = 1.0;
}

// Programmatic dependent launch: every step kernel lets its successor start early and run its
// prologue (loads that do not depend on the predecessor) while the predecessor drains; pdl_wait()
// returns once the predecessor grid has completed and its stores are visible.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---- panel: diagonal block Cholesky + row of triangular solves -------------------------------
// 256 threads.  Every CTA factorises the 64x64 diagonal block (identical arithmetic, so identical
// bits; cheaper than a second launch per block row) and CTA 0 stores the factor into B.diag[K] — never
// over A_KK itself, which CTAs of a later wave (grid larger than one resident wave) still have to read.  The block lives in
// registers: thread (c, q) owns A[q + 4i][c], i < 16.  Step j costs one barrier: the owners publish
// the raw row j, then every thread scales it by rsqrt(a_jj) itself and applies the rank-1 update to
// its 16 entries.  The stored diagonal is sqrt(a_jj) exactly; off-diagonals are a_jc * rsqrt(a_jj)
// (within 1 ulp of a_jc / sqrt(a_jj), the reference's form).  Each 64-thread quarter then solves
// U_KK^T X = A_KJ for one block J, a column of X per thread in registers: no barriers in the
// 2016-FMA substitution, U_KK reads are shared-memory broadcasts, diagonals enter as reciprocals.
// `tiles` (1, 2 or 4) quarters of a CTA take a block each: one per CTA while the row is short, so the
// substitutions spread over the SMs instead of sharing one FP64 pipe.
// The right-hand side rides along as one more column of the matrix (U^T y = b): warp 7 substitutes
// y_K = U_KK^-T w_K, and the thread that holds column c of the solved block U_KJ subtracts its dot
// product with y_K from w_J[c] — one owner per entry, fixed order — so no forward sweep is needed.
__global__ void __launch_bounds__(256, 1) k_chol_panel(BandDev B, int K, int m, int tiles, double *__restrict__ w,
                                                       double *__restrict__ y, int *__restrict__ err) {
    __shared__ double D[CBB];
    __shared__ double rowbuf[2][CB];
    __shared__ double rinv[CB];
    __shared__ double ys[CB];
    const int tid = threadIdx.x;
    const int c = tid & 63, rq = tid >> 6;
    pdl_trigger();
    pdl_wait();
    double *gkk = B.blk(K, K);
    double a[16];
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = gkk[(rq + 4 * i) * CB + c];
    bool bad = false;
#pragma unroll
    for (int j = 0; j < CB; j++) {
        const int ij = j >> 2, qj = j & 3;
        double *rb = rowbuf[j & 1];
        if (rq == qj) rb[c] = a[ij];
        __syncthreads();
        double ajj = rb[j];
        if (!(ajj > 0.0)) { bad = true; ajj = 1.0; }      // sparsecholeskyskyline returns false here
        const double rs = rsqrt(ajj);
        const double ujc = rb[c] * rs;
        if (rq == qj && c >= j) D[j * CB + c] = (c == j) ? ajj : ujc;     // diagonal: raw pivot, rooted below
#pragma unroll
        for (int i = ij; i < 16; i++) {
            const int r = rq + 4 * i;
            if (r > j && r <= c) a[i] -= (rb[r] * rs) * ujc;
        }
    }
    __syncthreads();
    if (tid < CB) {                                       // exact square roots, off the elimination chain
        const double d = sqrt(D[tid * CB + tid]);
        D[tid * CB + tid] = d;
        rinv[tid] = 1.0 / d;
    }
    __syncthreads();
    if (bad && tid == 0 && blockIdx.x == 0) atomicOr(err + ERR_NOT_SPD, 1);
    if (blockIdx.x == 0) {
        double *dkk = B.diag + (int64_t)K * CBB;
        for (int i = tid; i < CBB; i += 256) dkk[i] = ((i >> 6) <= (i & 63)) ? D[i] : 0.0;
    }
    __syncthreads();

    if (tid >= 224) {                                     // warp 7: forward substitution of the right-hand side
        const int lane = tid & 31;
        double v0 = w[(int64_t)K * CB + lane], v1 = w[(int64_t)K * CB + lane + 32];
        const double ri0 = rinv[lane], ri1 = rinv[lane + 32];
#pragma unroll 8
        for (int r = 0; r < 32; r++) {
            const double yr = __shfl_sync(0xffffffffu, v0 * ri0, r);
            if (lane == r) v0 = yr;
            if (lane > r) v0 -= D[r * CB + lane] * yr;
            v1 -= D[r * CB + lane + 32] * yr;
        }
#pragma unroll 8
        for (int r = 32; r < CB; r++) {
            const double yr = __shfl_sync(0xffffffffu, v1 * ri1, r - 32);
            if (lane + 32 == r) v1 = yr;
            if (lane + 32 > r) v1 -= D[r * CB + lane + 32] * yr;
        }
        ys[lane] = v0; ys[lane + 32] = v1;
        if (blockIdx.x == 0) { y[(int64_t)K * CB + lane] = v0; y[(int64_t)K * CB + lane + 32] = v1; }
    }
    const int jt = blockIdx.x * tiles + rq;
    const bool active = rq < tiles && jt < m;
    double t[CB];
    if (active) {
        double *g = B.blk(K, K + 1 + jt);
#pragma unroll
        for (int r = 0; r < CB; r++) t[r] = g[r * CB + c];
#pragma unroll
        for (int r = 0; r < CB; r++) {
            t[r] = t[r] * rinv[r];
#pragma unroll
            for (int r2 = r + 1; r2 < CB; r2++) t[r2] -= D[r * CB + r2] * t[r];
        }
#pragma unroll
        for (int r = 0; r < CB; r++) g[r * CB + c] = t[r];
    }
    __syncthreads();                                      // y_K is in shared memory
    if (active) {
        double dot = 0.0;
#pragma unroll
        for (int r = 0; r < CB; r++) dot += t[r] * ys[r];
        w[(int64_t)(K + 1 + jt) * CB + c] -= dot;
    }
}

// ---- trailing update: C_IJ -= U_KI^T U_KJ ----------------------------------------------------
// 128 threads per 64x64 block, 8x4 accumulators each: per k a thread reads 8 values of U_KI (two
// addresses per warp: broadcast) and 2+2 of U_KJ (conflict-free 16-byte lanes) for 32 FMAs, so the
// loop is bound by the FP64 pipe.
constexpr uint32_t BLK_BYTES = CBB * sizeof(double);

__device__ __forceinline__ void gemm64(const double *__restrict__ As, const double *__restrict__ Bs, int ty, int tx,
                                       double (&acc)[8][4]) {
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.0;
#pragma unroll 4
    for (int k = 0; k < CB; k++) {
        double av[8], bv[4];
        const double2 *pa = reinterpret_cast<const double2 *>(As + k * CB + ty * 8);
#pragma unroll
        for (int i = 0; i < 4; i++) { double2 v = pa[i]; av[2 * i] = v.x; av[2 * i + 1] = v.y; }
        const double2 b0v = *reinterpret_cast<const double2 *>(Bs + k * CB + 2 * tx);
        const double2 b1v = *reinterpret_cast<const double2 *>(Bs + k * CB + 32 + 2 * tx);
        bv[0] = b0v.x; bv[1] = b0v.y; bv[2] = b1v.x; bv[3] = b1v.y;
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) acc[i][j] += av[i] * bv[j];
    }
}
// a thread's 8x4 piece of a row-major 64x64 block: rows ty*8.., columns {2tx, 2tx+1, 32+2tx, 33+2tx}
__device__ __forceinline__ void load_piece(const double *blk, int ty, int tx, double2 (&c0)[8], double2 (&c1)[8]) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
        c0[i] = *reinterpret_cast<const double2 *>(blk + (ty * 8 + i) * CB + 2 * tx);
        c1[i] = *reinterpret_cast<const double2 *>(blk + (ty * 8 + i) * CB + 32 + 2 * tx);
    }
}
__device__ __forceinline__ void sub_store_piece(double *blk, int ty, int tx, double2 (&c0)[8], double2 (&c1)[8],
                                                const double (&acc)[8][4], bool sub) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
        if (sub) { c0[i].x -= acc[i][0]; c0[i].y -= acc[i][1]; c1[i].x -= acc[i][2]; c1[i].y -= acc[i][3]; }
        *reinterpret_cast<double2 *>(blk + (ty * 8 + i) * CB + 2 * tx) = c0[i];
        *reinterpret_cast<double2 *>(blk + (ty * 8 + i) * CB + 32 + 2 * tx) = c1[i];
    }
}

// One CTA owns up to `ch` consecutive blocks of block row I = K+1+a: U_KI is fetched once, the U_KJ
// blocks stream through a two-deep shared-memory ring with cp.async.bulk (one elected thread, byte-
// counting mbarriers), and the C block is prefetched into registers before the 64-deep product so
// its latency hides behind the FMAs.  Call after pdl_wait() with the three mbarriers initialised.
__device__ __forceinline__ void update_strip(const BandDev &B, int K, int m, int ch, int a, int chunk, double *sm,
                                             uint64_t *bar) {
    const int b0 = a + chunk * ch;
    if (b0 >= m) return;
    const int nb = min(ch, m - b0);
    double *As = sm;
    const int tid = threadIdx.x;
    const int I = K + 1 + a;
    if (tid == 0) {
        mbar_expect_tx(&bar[2], BLK_BYTES); bulk_g2s(As, B.blk(K, I), BLK_BYTES, &bar[2]);
        mbar_expect_tx(&bar[0], BLK_BYTES); bulk_g2s(sm + CBB, B.blk(K, K + 1 + b0), BLK_BYTES, &bar[0]);
        if (nb > 1) { mbar_expect_tx(&bar[1], BLK_BYTES); bulk_g2s(sm + 2 * CBB, B.blk(K, K + 2 + b0), BLK_BYTES, &bar[1]); }
    }
    const int ty = tid >> 4, tx = tid & 15;
    for (int t = 0; t < nb; t++) {
        const int J = K + 1 + b0 + t;
        double *gc = B.blk(I, J);
        double2 c0[8], c1[8];
        load_piece(gc, ty, tx, c0, c1);
        if (t == 0) mbar_wait(&bar[2], 0);
        mbar_wait(&bar[t & 1], (uint32_t)((t >> 1) & 1));
        double acc[8][4];
        gemm64(As, sm + (1 + (t & 1)) * CBB, ty, tx, acc);
        sub_store_piece(gc, ty, tx, c0, c1, acc, true);
        __syncthreads();                                      // ring slot t&1 is free again
        if (tid == 0 && t + 2 < nb) {
            mbar_expect_tx(&bar[t & 1], BLK_BYTES);
            bulk_g2s(sm + (1 + (t & 1)) * CBB, B.blk(K, J + 2), BLK_BYTES, &bar[t & 1]);
        }
    }
}

// Grid (chunks, rows); blockIdx.x runs fastest, which dispatches block row K+1 — the next panel's
// input — first.
__global__ void __launch_bounds__(128) k_chol_update(BandDev B, int K, int m, int ch) {
    extern __shared__ __align__(128) double sm[];
    __shared__ uint64_t bar[3];
    if ((int)blockIdx.y + (int)blockIdx.x * ch >= m) return;
    if (threadIdx.x == 0) {
        mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_init(&bar[2], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    pdl_trigger();
    __syncthreads();
    pdl_wait();
    update_strip(B, K, m, ch, blockIdx.y, blockIdx.x, sm, bar);
}

// ---- U x = y, block column J (descending): x_J = U_JJ^-1 y_J, then y_I -= U_IJ x_J for I in [F[J], J) ----
__global__ void __launch_bounds__(256) k_chol_bwd(BandDev B, int J, int cnt, double *__restrict__ y,
                                                  double *__restrict__ x) {
    __shared__ double D[CB * (CB + 1)];
    __shared__ double xs[CB];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    pdl_trigger();
    const double *gjj = B.diag + (int64_t)J * CBB;
#pragma unroll
    for (int i = 0; i < CBB / 256; i++) {
        const int e = tid + 256 * i;
        D[(e >> 6) * (CB + 1) + (e & 63)] = gjj[e];
    }
    const bool has = (int)blockIdx.x < cnt;
    const int I = B.F[J] + blockIdx.x;
    double g0[8], g1[8];
    if (has) {
        const double *g = B.blk(I, J);
#pragma unroll
        for (int i = 0; i < 8; i++) { g0[i] = g[(warp * 8 + i) * CB + lane]; g1[i] = g[(warp * 8 + i) * CB + lane + 32]; }
    }
    pdl_wait();
    if (tid < CB) xs[tid] = y[(int64_t)J * CB + tid];
    __syncthreads();
    if (tid < 32) {
        double v0 = xs[lane], v1 = xs[lane + 32];
        const double ri0 = 1.0 / D[lane * (CB + 1) + lane], ri1 = 1.0 / D[(lane + 32) * (CB + 1) + lane + 32];
#pragma unroll 8
        for (int cc = CB - 1; cc >= 32; cc--) {
            const double xc = __shfl_sync(0xffffffffu, v1 * ri1, cc - 32);
            if (lane + 32 == cc) v1 = xc;
            if (lane + 32 < cc) v1 -= D[(lane + 32) * (CB + 1) + cc] * xc;
            v0 -= D[lane * (CB + 1) + cc] * xc;
        }
#pragma unroll 8
        for (int cc = 31; cc >= 0; cc--) {
            const double xc = __shfl_sync(0xffffffffu, v0 * ri0, cc);
            if (lane == cc) v0 = xc;
            if (lane < cc) v0 -= D[lane * (CB + 1) + cc] * xc;
        }
        xs[lane] = v0; xs[lane + 32] = v1;
    }
    __syncthreads();
    if (blockIdx.x == 0 && tid < CB) x[(int64_t)J * CB + tid] = xs[tid];
    if (!has) return;
    const double x0 = xs[lane], x1 = xs[lane + 32];
    double p[8];
#pragma unroll
    for (int i = 0; i < 8; i++) p[i] = g0[i] * x0 + g1[i] * x1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int i = 0; i < 8; i++) p[i] += __shfl_xor_sync(0xffffffffu, p[i], o);
    if (lane < 8) {
        double v = p[0];
#pragma unroll
        for (int i = 1; i < 8; i++) if (lane == i) v = p[i];
        y[(int64_t)I * CB + warp * 8 + lane] -= v;
    }
}

template <typename... KArgs, typename... Args>
cudaError_t launch_step(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, bool pdl, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

}  // namespace

int solve_cholesky(stan_handle *h, stan_chol_report *rep) {
    if (!h->assembled) { set_error("stan_solve_cholesky: assemble first"); return STAN_E_STATE; }
    if (h->world != 1) { set_error("stan_solve_cholesky is single-GPU only (the skyline is not partitioned)"); return STAN_E_STATE; }
    cudaStream_t s = h->stream;
    const int64_t nn = h->n_nodes, n = 3 * nn;
    const int64_t nbk = (n + CB - 1) / CB, npad = nbk * CB;
    if (nbk > 0x7fffffff / 2) { set_error("stan_solve_cholesky: too many rows"); return STAN_E_ARG; }
    memset(rep, 0, sizeof *rep);
    STAN_CUDA(cudaEventRecord(h->ev0, s));

    // ---- block envelope (host: nbk integers) ----
    DevBuf<int32_t> d_min;
    STAN_TRY(d_min.alloc(nn, s));
    k_min_col<<<div_up(nn, 256), 256, 0, s>>>(nn, h->d_brow_ptr.p, h->d_bcol.p, d_min.p);
    std::vector<int32_t> minnb(nn);
    STAN_CUDA(cudaMemcpyAsync(minnb.data(), d_min.p, nn * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    STAN_CUDA(cudaStreamSynchronize(s));
    d_min.release(s);
    std::vector<int32_t> F(nbk), E(nbk);
    std::vector<int64_t> pofs(nbk + 1);
    for (int64_t J = 0; J < nbk; J++) {
        int64_t f = J;
        const int64_t q0 = (J * CB) / 3, q1 = std::min(nn - 1, (J * CB + CB - 1) / 3);
        for (int64_t q = q0; q <= q1; q++) f = std::min<int64_t>(f, (3 * (int64_t)minnb[q]) / CB);
        F[J] = (int32_t)f;
    }
    for (int64_t J = nbk - 2; J >= 0; J--) F[J] = std::min(F[J], F[J + 1]);
    pofs[0] = 0;
    for (int64_t J = 0; J < nbk; J++) pofs[J + 1] = pofs[J] + (J - F[J] + 1);
    {
        int64_t J = 0;
        for (int64_t K = 0; K < nbk; K++) {
            if (J < K) J = K;
            while (J + 1 < nbk && F[J + 1] <= K) J++;
            E[K] = (int32_t)J;
        }
    }
    const int64_t total_blocks = pofs[nbk];
    const double band_bytes = (double)total_blocks * CBB * sizeof(double);
    size_t mem_free = 0, mem_total = 0;
    STAN_CUDA(cudaMemGetInfo(&mem_free, &mem_total));
    if (band_bytes + (double)nbk * CBB * sizeof(double) > 0.95 * (double)mem_total) {
        set_error("stan_solve_cholesky: the skyline needs %.1f GB (%lld blocks of 64x64), the device has %.1f GB; use CG",
                  band_bytes / 1e9, (long long)total_blocks, mem_total / 1e9);
        return STAN_E_NOMEM;
    }
    DevBuf<double> band, diag, w, y, x;
    DevBuf<int32_t> dF;
    DevBuf<int64_t> dP;
    auto free_all = [&]() { band.release(s); diag.release(s); w.release(s); y.release(s); x.release(s); dF.release(s); dP.release(s); };
    if (band.alloc((size_t)total_blocks * CBB, s) != STAN_OK) {
        free_all();
        (void)cudaGetLastError();                         // the failed allocation must not poison later calls
        set_error("stan_solve_cholesky: cannot allocate the %.1f GB skyline; use CG", band_bytes / 1e9);
        return STAN_E_NOMEM;
    }
    STAN_TRY(diag.alloc((size_t)nbk * CBB, s));
    STAN_TRY(w.alloc(npad, s)); STAN_TRY(y.alloc(npad, s)); STAN_TRY(x.alloc(npad, s));
    STAN_TRY(dF.alloc(nbk, s)); STAN_TRY(dP.alloc(nbk + 1, s));
    STAN_CUDA(cudaMemcpyAsync(dF.p, F.data(), nbk * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    STAN_CUDA(cudaMemcpyAsync(dP.p, pofs.data(), (nbk + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, s));
    STAN_CUDA(cudaMemsetAsync(band.p, 0, (size_t)total_blocks * CBB * sizeof(double), s));
    STAN_CUDA(cudaMemsetAsync(h->d_err.p + ERR_NOT_SPD, 0, sizeof(int32_t), s));
    BandDev B{band.p, diag.p, dF.p, dP.p};
    k_band_scatter<<<div_up(n, 128), 128, 0, s>>>(nn, h->d_brow_ptr.p, h->d_bcol.p, h->d_vals.p, h->d_fixed.p, B);
    if (npad > n) k_band_pad<<<1, CB, 0, s>>>(n, npad, B);
    STAN_CUDA(cudaMemsetAsync(w.p, 0, npad * sizeof(double), s));
    STAN_CUDA(cudaMemcpyAsync(w.p, h->d_b.p, n * sizeof(double), cudaMemcpyDeviceToDevice, s));
    int64_t launches = 3;

    // ---- factorisation ----
    // per device and cheap: a process may hold handles on several GPUs
    STAN_CUDA(cudaFuncSetAttribute(k_chol_update, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * CBB * (int)sizeof(double)));
    const char *tiles_env = getenv("STAN_CHOL_TILES");     // tests force 1 to run panels of more than one wave
    const int tiles_forced = tiles_env ? atoi(tiles_env) : 0;
    STAN_CUDA(cudaEventRecord(h->ev1, s));
    double flops = 0.0;
    const double b3 = (double)CB * CB * CB;
    static const bool use_pdl = !(getenv("STAN_PDL") && atoi(getenv("STAN_PDL")) == 0);
    for (int64_t K = 0; K < nbk; K++) {
        const int m = E[K] - (int)K;
        const int tiles = (tiles_forced == 1 || tiles_forced == 2 || tiles_forced == 4)
                              ? tiles_forced : (m > 2 * h->sm_count ? 4 : m > h->sm_count ? 2 : 1);
        STAN_CUDA(launch_step(k_chol_panel, dim3(std::max(1, div_up(m, tiles))), dim3(256), 0, s, use_pdl && K > 0,
                              B, (int)K, m, tiles, w.p, y.p, h->d_err.p));
        launches++;
        flops += b3 / 3 + (double)m * b3;
        if (m > 0) {
            // blocks per CTA along a row: long strips reuse U_KI, but small trailing matrices need the CTAs
            const int ch = std::max(1, std::min(4, (m * (m + 1) / 2) / (4 * h->sm_count)));
            STAN_CUDA(launch_step(k_chol_update, dim3(div_up(m, ch), m), dim3(128), 3 * CBB * sizeof(double), s, use_pdl,
                                  B, (int)K, m, ch));
            launches++;
            flops += (double)m * (m + 1) * b3;      // m(m+1)/2 blocks x 2*64^3
        }
    }
    STAN_CUDA(cudaEventRecord(h->ev2, s));

    // ---- back substitution (U^T y = b was carried through the factorisation) ----
    // the first sweep kernel must see the finished factor before its prologue: no early start for it
    for (int64_t J = nbk - 1; J >= 0; J--) {
        const int cnt = (int)J - F[J];
        STAN_CUDA(launch_step(k_chol_bwd, dim3(std::max(1, cnt)), dim3(256), 0, s, use_pdl && J < nbk - 1, B, (int)J, cnt,
                              y.p, x.p));
    }
    launches += nbk;
    STAN_CUDA(cudaEventRecord(h->ev3, s));
    STAN_CUDA(cudaGetLastError());

    STAN_TRY(h->d_x.alloc(n, s));
    int32_t herr[8] = {0};
    STAN_CUDA(cudaMemcpyAsync(herr, h->d_err.p, sizeof herr, cudaMemcpyDeviceToHost, s));
    STAN_CUDA(cudaStreamSynchronize(s));
    const bool spd = herr[ERR_NOT_SPD] == 0;
    if (spd) STAN_CUDA(cudaMemcpyAsync(h->d_x.p, x.p, n * sizeof(double), cudaMemcpyDeviceToDevice, s));
    else     STAN_CUDA(cudaMemsetAsync(h->d_x.p, 0, n * sizeof(double), s));   // "filled by zeros" (SolverFunctions.cs:420)
    STAN_CUDA(cudaStreamSynchronize(s));
    float t01 = 0, t12 = 0, t23 = 0;
    cudaEventElapsedTime(&t01, h->ev0, h->ev1);
    cudaEventElapsedTime(&t12, h->ev1, h->ev2);
    cudaEventElapsedTime(&t23, h->ev2, h->ev3);
    free_all();
    h->x_in_alt = false;
    h->sol = h->d_x.p;
    h->solved = true;
    h->launches += launches;
    rep->terminationtype = spd ? 1 : -3;
    rep->block = CB;
    rep->n = n;
    rep->n_blocks = total_blocks;
    rep->skyline_bytes = (int64_t)band_bytes;
    rep->flops = flops;
    rep->setup_ms = t01;
    rep->factor_ms = t12;
    rep->solve_ms = t23;
    rep->kernel_launches = launches;
    return STAN_OK;
}

}  // namespace stan
