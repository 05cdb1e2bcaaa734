// Jacobi-preconditioned Conjugate Gradient on the device.
//
// Replaces SolverFunctions.LinearSolver_CG (/root/reference/src/STAN_Solver/SolverFunctions.cs:270-330),
// i.e. alglib.lincgsolvesparse with its defaults: x0 = 0, diagonal preconditioner applied as
// z = r * (1/sqrt(A_ii))^2, stop on ||r||2 <= EpsF*||b||2, true-residual refresh plus energy
// functional check every 10th iteration (terminationtype 7), MaxIts (5), breakdown (-4/-5).
// The recurrences follow SURVEY.md Appendix A; ALGLIB's source is not in the reference checkout.
//
// Device design: three kernels per iteration —
//   spmv_dot   mv = A p and p.mv       (HBM-bound: 8 B/value + 4 B per 3x3 block + vectors)
//   update     x += a p, r -= a mv, r.r and r.(M^-1 r)
//   direction  p = M^-1 r + b p
// Every scalar ALGLIB computes on the host (alpha, beta, stopping tests, counters) is computed by
// one thread of the last CTA to finish a reduction and kept in a device-resident CgState, so the
// host enqueues whole batches of iterations and only reads the state back between batches.
// Reductions are two-level with a fixed tree: bitwise reproducible run to run.
#include <cmath>

#include "common.cuh"
#include "bulk.cuh"

namespace stan {

namespace {

constexpr int VEC_THREADS = 256;
constexpr int SPMV_THREADS = 256;
constexpr int SPMV_WARPS = SPMV_THREADS / 32;

enum ScalarStep { SC_NONE = 0, SC_INIT = 1, SC_AFTER_SPMV = 2, SC_AFTER_UPDATE = 3, SC_AFTER_REFRESH = 4 };

// trajectory record of iteration k (stan_get_cg_history); merit = NaN except on refresh iterations
__device__ __forceinline__ void record_history(CgState *st, int k, double r2, double beta, double merit) {
    if (st->hist && k >= 1 && k <= st->hist_cap) {
        double *hrow = st->hist + 4 * (size_t)(k - 1);
        hrow[0] = r2; hrow[1] = st->alpha; hrow[2] = beta; hrow[3] = merit;
    }
}

__device__ __forceinline__ void finish_iteration(CgState *st, double cr2, double rznew, double merit_k) {
    const int k = st->k + 1;
    st->k = k;
    st->r2 = cr2;
    record_history(st, k, cr2, 0.0, merit_k);
    if (sqrt(cr2) <= st->epsf_bnorm) { st->type = 1; st->done = 1; return; }
    if (st->maxits > 0 && k >= st->maxits) { st->type = 5; st->done = 1; return; }
    const int64_t kk = k - st->counter_off;
    double beta = 0.0;                                   // restart: p = z
    if (kk % st->restart != 0) {
        const double uvar = st->rz;
        if (!isfinite(uvar) || uvar == 0.0 || !isfinite(rznew)) { st->type = -4; st->done = 1; return; }
        beta = rznew / uvar;
        record_history(st, k, cr2, beta, merit_k);
    }
    st->beta = beta;
    st->rz = rznew;
}

__device__ void scalar_step(CgState *st, int step) {
    if (step == SC_INIT) {                               // partial: [0] = b.b, [1] = r.z with r = b
        const double bn = sqrt(st->partial[0]);
        st->bnorm = bn;
        st->epsf_bnorm *= bn;                            // host stored EpsF here
        st->r2 = st->partial[0];
        st->rz = st->partial[1];
        st->nmv = 1;                                     // r0 = b - A*x0 counted as ALGLIB does
        if (bn == 0.0 || sqrt(st->r2) <= st->epsf_bnorm) { st->type = 1; st->done = 1; }
        else if (!isfinite(st->r2)) { st->type = -4; st->done = 1; }
    } else if (step == SC_AFTER_SPMV) {                  // partial[0] = p.(A p)
        const double vmv = st->partial[0];
        st->vmv = vmv;
        st->nmv++;
        if (!isfinite(vmv)) { st->type = -4; st->done = 1; }
        else if (vmv <= 0.0) { st->type = -5; st->done = 1; }
        else {
            const double alpha = st->rz / vmv;
            if (!isfinite(alpha)) { st->type = -4; st->done = 1; }
            st->alpha = alpha;
        }
    } else if (step == SC_AFTER_UPDATE) {                // partial: [0] = r.r, [1] = r.z
        finish_iteration(st, st->partial[0], st->partial[1], nan(""));
        if (st->done) st->x_pending = 1;                 // x += alpha p of this iteration rides in k_direction, which now skips
    } else if (step == SC_AFTER_REFRESH) {               // + [2] = 2 b.cx, [3] = (A cx).cx from the SpMV
        st->nmv++;
        const double v1 = st->partial[3] - st->partial[2];
        if (st->merit_check && !(v1 < st->merit)) {      // rounding stagnation: previous x is returned
            st->k += 1;
            st->type = 7;
            st->done = 1;
            record_history(st, st->k, st->r2, 0.0, v1);
            return;
        }
        st->merit = v1;
        st->x_in_alt ^= 1;
        finish_iteration(st, st->partial[0], st->partial[1], v1);
    }
}

// Multi-GPU, peer-memory mode: warp 0 of the CTA that holds this rank's partial sums exchanges them with every
// peer over NVLink and adds the W contributions in rank order — every rank obtains bitwise identical sums, so
// all ranks take the same branches.  Lane r talks to rank r.  A double travels as two 8-byte words, each
// carrying 32 payload bits and the 32-bit sequence tag of this reduction: an aligned 8-byte store arrives
// whole, so the reader needs no separate flag and the writer no fence between payload and flag — one NVLink
// one-way latency per reduction.  Slots are double-buffered by sequence parity; a rank cannot be two reductions
// ahead of a peer because each reduction needs that peer's contribution.
__device__ __forceinline__ void st_volatile_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ void cross_rank_sum(CgState *st, int cnt, int lane) {
    CommDev *cd = st->comm;
    const int W = cd->world, me = cd->rank;
    const unsigned long long seq = st->red_seq + 1;
    const int par = (int)(seq & 1);
    const unsigned long long tag = ((seq % 0xffffffffull) + 1) << 32;     // never 0: a fresh window is all zeros
    double mine[4];
    for (int i = 0; i < 4; i++) mine[i] = i < cnt ? st->partial[i] : 0.0;
    if (lane < W && lane != me)
        for (int i = 0; i < cnt; i++) {
            const unsigned long long bits = (unsigned long long)__double_as_longlong(mine[i]);
            unsigned long long *dst = cd->ctrl[lane]->red[par][me][i];
            st_volatile_u64(dst, (bits & 0xffffffffull) | tag);
            st_volatile_u64(dst + 1, (bits >> 32) | tag);
        }
    double got[4] = {0.0, 0.0, 0.0, 0.0};
    bool timed_out = false;
    if (lane < W) {
        if (lane == me) {
            for (int i = 0; i < 4; i++) got[i] = mine[i];
        } else {
            const long long t0 = clock64();
            for (int i = 0; i < cnt && !timed_out; i++) {
                const unsigned long long *src = cd->ctrl[me]->red[par][lane][i];
                unsigned long long w0, w1;
                for (;;) {
                    w0 = ld_volatile_u64(src); w1 = ld_volatile_u64(src + 1);
                    if ((w0 & 0xffffffff00000000ull) == tag && (w1 & 0xffffffff00000000ull) == tag) break;
                    if (clock64() - t0 > 8000000000LL) { timed_out = true; break; }       // ~4 s: a peer died
                }
                got[i] = __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
            }
        }
    }
    const bool any_timeout = __any_sync(0xffffffffu, timed_out);
    double sum[4] = {0.0, 0.0, 0.0, 0.0};
    for (int r = 0; r < W; r++)
        for (int i = 0; i < 4; i++) sum[i] += __shfl_sync(0xffffffffu, got[i], r);     // rank order, same on every rank
    if (lane == 0) {
        if (any_timeout) { atomicOr(cd->err + 4, 1); st->type = -4; st->done = 1; }
        for (int i = 0; i < cnt; i++) st->partial[i] = sum[i];
        st->red_seq = seq;
    }
    __syncwarp();
}

__global__ void k_scalar(CgState *st, int step) {
    if (st->done) return;
    scalar_step(st, step);
}

// Block-level sum of NV values, published to partials[]; the last CTA to arrive folds all CTA
// partials in a fixed order into st->partial[slot0..] and (single GPU) runs the scalar step.
// Returns true in every thread of the last CTA.
template <int NV>
__device__ __forceinline__ bool grid_reduce(double (&v)[NV], double *partials, unsigned int *counter, CgState *st,
                                            int slot0, int step, bool run_scalar) {
    __shared__ double s_red[NV][32];
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; i++) {
        double x = v[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) s_red[i][warp] = x;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int i = 0; i < NV; i++) {
            double x = lane < nw ? s_red[i][lane] : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            if (lane == 0) partials[(size_t)blockIdx.x * NV + i] = x;
        }
    }
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int ticket = atomicInc(counter, gridDim.x - 1);   // wraps to 0 for the next launch
        s_last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return false;
    __threadfence();
    double acc[NV];
#pragma unroll
    for (int i = 0; i < NV; i++) acc[i] = 0.0;
    for (unsigned int b = threadIdx.x; b < gridDim.x; b += blockDim.x)
#pragma unroll
        for (int i = 0; i < NV; i++) acc[i] += __ldcg(&partials[(size_t)b * NV + i]);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; i++) {
        double x = acc[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) s_red[i][warp] = x;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int i = 0; i < NV; i++) {
            double x = lane < nw ? s_red[i][lane] : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            if (lane == 0) st->partial[slot0 + i] = x;
        }
        if (run_scalar && step != SC_NONE) {
            __syncwarp();
            if (lane == 0) trace_mark(st, step == SC_AFTER_SPMV ? TR_SPMV_LOCAL_DONE : TR_UPDATE_LOCAL_DONE);
            if (st->comm) cross_rank_sum(st, step == SC_AFTER_REFRESH ? 4 : (step == SC_AFTER_SPMV ? 1 : 2), lane);
            if (lane == 0) trace_mark(st, step == SC_AFTER_SPMV ? TR_SPMV_END : (step == SC_AFTER_REFRESH ? TR_REFRESH_END : TR_UPDATE_END));
            if (lane == 0 && !st->done) scalar_step(st, step);
        }
    }
    return true;
}

// ---- SpMV -----------------------------------------------------------------------------------
// One warp per block row.  A block row stores three scalar rows of length 3*nb back to back, so
// lanes stream 256 B segments of each; the column of entry j is 3*bcol[j/3] + j%3, i.e. one int
// per nine values.  x is gathered through L1/L2 (neighbouring rows share their columns).
__device__ __forceinline__ double ld_stream(const double *p) {
    double v;   // not volatile: the scheduler must be free to batch these ahead of their uses
    asm("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

template <bool DOT, int MINB>
__global__ void __launch_bounds__(SPMV_THREADS, MINB)
k_spmv(int64_t nrows, const int32_t *__restrict__ row_list, const int32_t *__restrict__ brow_ptr,
       const int32_t *__restrict__ bcol, const double *__restrict__ vals, const double *__restrict__ x,
       double *__restrict__ y, double *partials, unsigned int *counter, CgState *st, int slot, int step,
       bool run_scalar) {
    if (st && st->done) return;
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * SPMV_WARPS + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * SPMV_WARPS;
    double dsum = 0.0;
    for (int64_t it = warp0; it < nrows; it += nwarps) {
        const int64_t row = row_list ? row_list[it] : it;
        const int s = brow_ptr[row];
        const int len = 3 * (brow_ptr[row + 1] - s);
        const double *v0 = vals + 9 * (int64_t)s;
        const double *v1 = v0 + len, *v2 = v1 + len;
        const int32_t *cols = bcol + s;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0;
        for (int j0 = 0; j0 < len; j0 += 96) {
            int jj[3];
            double xv[3], m0[3], m1[3], m2[3];
#pragma unroll
            for (int u = 0; u < 3; u++) {
                const int j = j0 + lane + 32 * u;
                jj[u] = j < len ? j : len - 1;
            }
#pragma unroll
            for (int u = 0; u < 3; u++) {
                m0[u] = ld_stream(v0 + jj[u]);
                m1[u] = ld_stream(v1 + jj[u]);
                m2[u] = ld_stream(v2 + jj[u]);
            }
#pragma unroll
            for (int u = 0; u < 3; u++) {
                const int blk = jj[u] / 3;
                const int col = __ldg(cols + blk);
                xv[u] = (j0 + lane + 32 * u < len) ? x[3 * (int64_t)col + (jj[u] - 3 * blk)] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 3; u++) {
                a0 += m0[u] * xv[u];
                a1 += m1[u] * xv[u];
                a2 += m2[u] * xv[u];
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a0 += __shfl_xor_sync(0xffffffffu, a0, o);
            a1 += __shfl_xor_sync(0xffffffffu, a1, o);
            a2 += __shfl_xor_sync(0xffffffffu, a2, o);
        }
        if (lane < 3) {
            const double yv = lane == 0 ? a0 : (lane == 1 ? a1 : a2);
            y[3 * row + lane] = yv;
            if (DOT) dsum += yv * x[3 * row + lane];
        }
    }
    if (DOT) {
        double v[1] = {dsum};
        grid_reduce<1>(v, partials, counter, st, slot, step, run_scalar);
    }
}

// ---- SpMV, bulk-copy pipeline ------------------------------------------------------------------
// Persistent CTAs (one per SM).  The matrix is one contiguous stream, so a single elected thread
// moves it with cp.async.bulk (the TMA engine; UBLKCP in SASS) in chunks of BULK_ROWS block rows —
// values, block columns and row pointers — into a ring of shared-memory stages, each guarded by
// an mbarrier that completes on byte count.  Memory-level parallelism is then set by the ring
// depth (>= 100 KB in flight per SM) instead of by how many loads the compiler keeps outstanding
// per warp; the 16 consumer warps only touch global memory for the x gather and the y store.
constexpr int BULK_ROWS = 16;
constexpr int BULK_WARPS = 16;                         // consumer warps; warp BULK_WARPS is the producer
constexpr int BULK_THREADS = 32 * (BULK_WARPS + 1);

struct BulkLayout {          // per-stage byte offsets inside dynamic shared memory
    int vals_off, cols_off, rp_off, stage_bytes, stages;
};

template <bool DOT>
__global__ void __launch_bounds__(BULK_THREADS)
k_spmv_bulk(int64_t nrows, const int32_t *__restrict__ brow_ptr, const int32_t *__restrict__ bcol,
            const double *__restrict__ vals, const double *__restrict__ x, double *__restrict__ y, BulkLayout L,
            double *partials, unsigned int *counter, CgState *st, int slot, int step, bool run_scalar) {
    if (st && st->done) return;
    extern __shared__ __align__(128) unsigned char s_raw[];
    __shared__ __align__(8) uint64_t s_full[8], s_empty[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int S = L.stages;
    const int64_t nchunks = (nrows + BULK_ROWS - 1) / BULK_ROWS;
    const int64_t my_n = blockIdx.x < nchunks ? (nchunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    if (tid == 0) {
        for (int i = 0; i < S; i++) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], BULK_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    double dsum = 0.0;
    if (warp == BULK_WARPS) {
        // ---- producer: one lane streams this CTA's chunks through the ring ----
        if (lane == 0) {
            for (int64_t i = 0; i < my_n; i++) {
                const int stg = (int)(i % S);
                if (i >= S) mbar_wait(&s_empty[stg], (uint32_t)(((i / S) - 1) & 1));
                const int64_t c = blockIdx.x + i * (int64_t)gridDim.x;
                const int64_t r0 = c * BULK_ROWS, r1 = (r0 + BULK_ROWS < nrows) ? r0 + BULK_ROWS : nrows;
                unsigned char *base = s_raw + (size_t)stg * L.stage_bytes;
                const int64_t b0 = brow_ptr[r0], b1 = brow_ptr[r1];
                const int64_t voff = 72 * b0, voff_al = voff & ~(int64_t)15;
                const uint32_t vbytes = (uint32_t)(((voff - voff_al) + 72 * (b1 - b0) + 15) & ~(int64_t)15);
                const int64_t coff = 4 * b0, coff_al = coff & ~(int64_t)15;
                const uint32_t cbytes = (uint32_t)(((coff - coff_al) + 4 * (b1 - b0) + 15) & ~(int64_t)15);
                const uint32_t rbytes = (uint32_t)((4 * (r1 - r0 + 1) + 15) & ~(int64_t)15);   // r0 % 16 == 0
                mbar_expect_tx(&s_full[stg], vbytes + cbytes + rbytes);
                bulk_g2s(base + L.vals_off, (const unsigned char *)vals + voff_al, vbytes, &s_full[stg]);
                bulk_g2s(base + L.cols_off, (const unsigned char *)bcol + coff_al, cbytes, &s_full[stg]);
                bulk_g2s(base + L.rp_off, (const unsigned char *)(brow_ptr + r0), rbytes, &s_full[stg]);
            }
        }
    } else {
        // ---- consumers: one block row per warp and chunk; the x gather of the next chunk's row is
        // issued before the current row is reduced, so its L2 latency hides behind useful work ----
        struct RowCtx { const double *v0; int len, jj[3]; double xv[3]; int64_t row; bool have; };
        auto fetch = [&](int64_t i, RowCtx &r) {
            const int stg = (int)(i % S);
            mbar_wait(&s_full[stg], (uint32_t)((i / S) & 1));
            const unsigned char *base = s_raw + (size_t)stg * L.stage_bytes;
            const int32_t *rp = reinterpret_cast<const int32_t *>(base + L.rp_off);
            const int64_t r0 = (blockIdx.x + i * (int64_t)gridDim.x) * BULK_ROWS;
            const int nr = (int)((nrows - r0) < BULK_ROWS ? (nrows - r0) : BULK_ROWS);
            r.have = warp < nr;
            if (!r.have) return;
            const int b0 = rp[0];
            const int sblk = rp[warp] - b0;
            r.len = 3 * (rp[warp + 1] - rp[warp]);
            r.row = r0 + warp;
            r.v0 = reinterpret_cast<const double *>(base + L.vals_off) + ((9 * (int64_t)b0) & 1) + 9 * sblk;
            const int32_t *cols = reinterpret_cast<const int32_t *>(base + L.cols_off) + (b0 & 3) + sblk;
#pragma unroll
            for (int u = 0; u < 3; u++) {
                const int j = lane + 32 * u;
                r.jj[u] = j < r.len ? j : r.len - 1;
                const int blk = r.jj[u] / 3;
                const double xg = x[3 * (int64_t)cols[blk] + (r.jj[u] - 3 * blk)];
                r.xv[u] = j < r.len ? xg : 0.0;
            }
        };
        RowCtx cur, nxt;
        cur.have = nxt.have = false;
        if (my_n > 0) fetch(0, cur);
        for (int64_t i = 0; i < my_n; i++) {
            const int stg = (int)(i % S);
            if (i + 1 < my_n) fetch(i + 1, nxt);
            if (cur.have) {
                const double *v0 = cur.v0, *v1 = v0 + cur.len, *v2 = v1 + cur.len;
                double a0 = 0.0, a1 = 0.0, a2 = 0.0;
#pragma unroll
                for (int u = 0; u < 3; u++) {
                    a0 += v0[cur.jj[u]] * cur.xv[u];
                    a1 += v1[cur.jj[u]] * cur.xv[u];
                    a2 += v2[cur.jj[u]] * cur.xv[u];
                }
                if (cur.len > 96) {                             // rows wider than 32 blocks: rest without prefetch
                    const unsigned char *base = s_raw + (size_t)stg * L.stage_bytes;
                    const int32_t *rp = reinterpret_cast<const int32_t *>(base + L.rp_off);
                    const int b0 = rp[0];
                    const int32_t *cols = reinterpret_cast<const int32_t *>(base + L.cols_off) + (b0 & 3) + (rp[warp] - b0);
                    for (int j = 96 + lane; j < cur.len; j += 32) {
                        const int blk = j / 3;
                        const double xg = x[3 * (int64_t)cols[blk] + (j - 3 * blk)];
                        a0 += v0[j] * xg; a1 += v1[j] * xg; a2 += v2[j] * xg;
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    a0 += __shfl_xor_sync(0xffffffffu, a0, o);
                    a1 += __shfl_xor_sync(0xffffffffu, a1, o);
                    a2 += __shfl_xor_sync(0xffffffffu, a2, o);
                }
                if (lane < 3) {
                    const double yv = lane == 0 ? a0 : (lane == 1 ? a1 : a2);
                    y[3 * cur.row + lane] = yv;
                    if (DOT) dsum += yv * x[3 * cur.row + lane];
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[stg]);          // this warp is done with the stage
            cur = nxt;
        }
    }
    if (DOT) {
        double v[1] = {dsum};
        grid_reduce<1>(v, partials, counter, st, slot, step, run_scalar);
    }
}

// ---- SpMV, bulk-copy tiles with one thread per scalar row ---------------------------------------
// ncu on the warp-per-row kernels (profiles/r01_spmv.md) shows both pinned at ~5.0 TB/s by the L1
// data pipe: ~100 wavefronts per block row (27 for misaligned value loads, ~24 for the x gather,
// 30 for the shuffle reduction).  Here the TMA engine stages TILE_ROWS block rows in shared memory
// and every thread owns one scalar row: values and columns come from shared memory conflict-free,
// the row sum needs no shuffles, x is gathered through L1 where the lanes of a warp (consecutive
// rows) hit the same few lines, and y is stored coalesced.  ~45 wavefronts and ~35 instructions
// per block row.  G consumer groups each own one stage of the ring; a producer warp refills a
// stage as soon as its group releases it.
template <bool DOT, int TILE_ROWS, int GROUPS>
__global__ void __launch_bounds__(32 * (GROUPS * ((3 * TILE_ROWS + 31) / 32) + 1))
k_spmv_tile(int64_t nrows, const int32_t *__restrict__ brow_ptr, const int32_t *__restrict__ bcol,
            const double *__restrict__ vals, const double *__restrict__ x, double *__restrict__ y, BulkLayout L,
            double *partials, unsigned int *counter, CgState *st, int slot, int step, bool run_scalar) {
    if (st && st->done) return;
    constexpr int GW = (3 * TILE_ROWS + 31) / 32;          // warps per consumer group
    extern __shared__ __align__(128) unsigned char s_raw[];
    __shared__ __align__(8) uint64_t s_full[GROUPS], s_empty[GROUPS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t nchunks = (nrows + TILE_ROWS - 1) / TILE_ROWS;
    const int64_t my_n = blockIdx.x < nchunks ? (nchunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    if (tid == 0) {
        for (int i = 0; i < GROUPS; i++) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], GW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    double dsum = 0.0;
    if (warp == GROUPS * GW) {
        if (lane == 0) {                                   // ---- producer ----
            for (int64_t i = 0; i < my_n; i++) {
                const int g = (int)(i % GROUPS);
                const int64_t k = i / GROUPS;
                if (k > 0) mbar_wait(&s_empty[g], (uint32_t)((k - 1) & 1));
                const int64_t r0 = (blockIdx.x + i * (int64_t)gridDim.x) * TILE_ROWS;
                const int64_t r1 = (r0 + TILE_ROWS < nrows) ? r0 + TILE_ROWS : nrows;
                unsigned char *base = s_raw + (size_t)g * L.stage_bytes;
                const int64_t b0 = brow_ptr[r0], b1 = brow_ptr[r1];
                const int64_t voff = 72 * b0, voff_al = voff & ~(int64_t)15;
                const uint32_t vbytes = (uint32_t)(((voff - voff_al) + 72 * (b1 - b0) + 15) & ~(int64_t)15);
                const int64_t coff = 4 * b0, coff_al = coff & ~(int64_t)15;
                const uint32_t cbytes = (uint32_t)(((coff - coff_al) + 4 * (b1 - b0) + 15) & ~(int64_t)15);
                const uint32_t rbytes = (uint32_t)((4 * (r1 - r0 + 1) + 15) & ~(int64_t)15);   // TILE_ROWS % 4 == 0
                mbar_expect_tx(&s_full[g], vbytes + cbytes + rbytes);
                bulk_g2s(base + L.vals_off, (const unsigned char *)vals + voff_al, vbytes, &s_full[g]);
                bulk_g2s(base + L.cols_off, (const unsigned char *)bcol + coff_al, cbytes, &s_full[g]);
                bulk_g2s(base + L.rp_off, (const unsigned char *)(brow_ptr + r0), rbytes, &s_full[g]);
            }
        }
    } else {
        const int g = warp / GW;                           // ---- consumers ----
        const int t = tid - g * GW * 32;                   // scalar row of the tile owned by this thread
        const int br = t / 3, a = t - 3 * br;
        const unsigned char *base = s_raw + (size_t)g * L.stage_bytes;
        const int32_t *rp = reinterpret_cast<const int32_t *>(base + L.rp_off);
        for (int64_t i = g, k = 0; i < my_n; i += GROUPS, k++) {
            mbar_wait(&s_full[g], (uint32_t)(k & 1));
            const int64_t r0 = (blockIdx.x + i * (int64_t)gridDim.x) * TILE_ROWS;
            const int nr = (int)((nrows - r0) < TILE_ROWS ? (nrows - r0) : TILE_ROWS);
            if (br < nr) {
                const int b0 = rp[0];
                const int sblk = rp[br] - b0, nb = rp[br + 1] - rp[br];
                const double *v = reinterpret_cast<const double *>(base + L.vals_off) + ((9 * (int64_t)b0) & 1) +
                                  9 * sblk + a * 3 * nb;
                const int32_t *cols = reinterpret_cast<const int32_t *>(base + L.cols_off) + (b0 & 3) + sblk;
                // batches of 9 blocks: all column and x loads of a batch are issued before the first
                // FMA, and three accumulators keep the DFMA chain short
                double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0;
                int c = 0;
                for (; c + 9 <= nb; c += 9) {
                    int q[9];
                    double xa[9], xb[9], xc[9];
#pragma unroll
                    for (int u = 0; u < 9; u++) q[u] = cols[c + u];
#pragma unroll
                    for (int u = 0; u < 9; u++) {
                        const double *xp = x + 3 * (int64_t)q[u];
                        xa[u] = xp[0]; xb[u] = xp[1]; xc[u] = xp[2];
                    }
#pragma unroll
                    for (int u = 0; u < 9; u++) {
                        acc0 += v[3 * (c + u)] * xa[u];
                        acc1 += v[3 * (c + u) + 1] * xb[u];
                        acc2 += v[3 * (c + u) + 2] * xc[u];
                    }
                }
                for (; c < nb; c++) {
                    const double *xp = x + 3 * (int64_t)cols[c];
                    acc0 += v[3 * c] * xp[0];
                    acc1 += v[3 * c + 1] * xp[1];
                    acc2 += v[3 * c + 2] * xp[2];
                }
                const double acc = (acc0 + acc1) + acc2;
                const int64_t dof = 3 * (r0 + br) + a;
                y[dof] = acc;
                if (DOT) dsum += acc * x[dof];
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[g]);
        }
    }
    if (DOT) {
        double v[1] = {dsum};
        grid_reduce<1>(v, partials, counter, st, slot, step, run_scalar);
    }
}

// ---- SpMV, bulk-copy tiles, three threads per scalar row ----------------------------------------
// The tile kernel above is bound by the serial chain inside each thread (column -> x -> DFMA, 27
// times; ptxas keeps only ~1 block of lookahead), so a stage is held for ~8000 cycles.  Here a tile
// of 32 block rows (96 scalar rows = 3 full warps) is processed by 9 warps: warp (part, rw) owns
// the blocks [part*nb/3, (part+1)*nb/3) of scalar rows 32*rw .. 32*rw+31.  Lanes of a warp still
// read the same column slot of consecutive rows (few L1 lines per x gather); the three partial
// sums of a row meet in shared memory after a group-wide named barrier, in a fixed order.
constexpr int T3_ROWS = 32, T3_PARTS = 3, T3_GROUPS = 3;
constexpr int T3_GW = 3 * T3_PARTS;                       // warps per group
constexpr int T3_THREADS = 32 * (T3_GROUPS * T3_GW + 1);

// HALO = true (several GPUs, peer-memory mode) adds the halo exchange of the input vector to the product itself:
//  * tiles are taken in the order of ha.tile_order — the ones whose rows have no halo column first;
//  * the producer warp, once the first ring stages are in flight, stores this rank's boundary entries of x into
//    the halo tails of the neighbours' copies of the same vector (NVLink stores), fences, and the last CTA to
//    get there raises the neighbours' sequence flags;
//  * a consumer reaching the first tile with halo columns waits (once) for the flags of the ranks it reads from.
// The exchange therefore overlaps the interior rows, there is no separate push / wait launch and no copy: the
// peers write straight into the tail of x, which starts on a fresh 128-byte line (HALO_ALIGN) so no L1 line
// fetched for owned entries can hold stale halo entries.  x is not __restrict__/read-only in this instantiation:
// its tail changes while the kernel runs.
template <bool HALO> struct XArg { typedef const double *__restrict__ type; };   // read-only for the whole kernel
template <> struct XArg<true> { typedef const double *type; };                      // its halo tail is written by peers

template <bool DOT, bool HALO>
__global__ void __launch_bounds__(T3_THREADS, 1)
k_spmv_tile3(int64_t nrows, const int32_t *__restrict__ brow_ptr, const int32_t *__restrict__ bcol,
             const double *__restrict__ vals, typename XArg<HALO>::type x, double *__restrict__ y, BulkLayout L,
             double *partials, unsigned int *counter, CgState *st, int slot, int step, bool run_scalar, HaloArgs ha) {
    if (st && st->done) return;
    if (blockIdx.x == 0 && threadIdx.x == 0) trace_mark(st, TR_SPMV_BEGIN);
    extern __shared__ __align__(128) unsigned char s_raw[];
    __shared__ __align__(8) uint64_t s_full[T3_GROUPS], s_empty[T3_GROUPS];
    __shared__ double s_part[T3_GROUPS][T3_PARTS][3 * T3_ROWS];
    __shared__ int s_tile[T3_GROUPS];                      // HALO: tile held by each ring stage (written by the producer)
#define X_AT(i) x[i]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t nchunks = (nrows + T3_ROWS - 1) / T3_ROWS;
    const int64_t my_n = blockIdx.x < nchunks ? (nchunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const unsigned long long expect = HALO ? st->halo_seq + 1 : 0;     // every CTA reads it before any CTA can finish
    if (tid == 0) {
        for (int i = 0; i < T3_GROUPS; i++) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], T3_GW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto tile_of = [&](int64_t i) -> int64_t {             // i-th tile of this CTA
        const int64_t g = blockIdx.x + i * (int64_t)gridDim.x;
        return HALO ? (int64_t)ha.tile_order[g] : g;
    };

    double dsum = 0.0;
    if (warp == T3_GROUPS * T3_GW) {
        // ---- producer (lane 0) ----
        // Three groups finish a tile every ~1.3 us between them, so nothing on this loop's path may wait for
        // global memory: the tile id is fetched two tiles ahead and the row pointers one tile ahead.
        if (lane == 0 && my_n > 0) {
            int64_t tile_a = tile_of(0), tile_b = my_n > 1 ? tile_of(1) : 0;      // tiles i and i + 1
            int64_t r0 = tile_a * T3_ROWS, r1 = (r0 + T3_ROWS < nrows) ? r0 + T3_ROWS : nrows;
            int64_t b0 = brow_ptr[r0], b1 = brow_ptr[r1];
            for (int64_t i = 0; i < my_n; i++) {
                const int g = (int)(i % T3_GROUPS);
                const int64_t k = i / T3_GROUPS;
                const int64_t tile_c = i + 2 < my_n ? tile_of(i + 2) : 0;
                const int64_t nr0 = tile_b * T3_ROWS, nr1 = (nr0 + T3_ROWS < nrows) ? nr0 + T3_ROWS : nrows;
                int64_t nb0 = 0, nb1 = 0;
                if (i + 1 < my_n) { nb0 = brow_ptr[nr0]; nb1 = brow_ptr[nr1]; }
                if (k > 0) mbar_wait(&s_empty[g], (uint32_t)((k - 1) & 1));
                if (HALO) s_tile[g] = (int)tile_a;         // published by the arrive.expect_tx (release) that follows
                unsigned char *base = s_raw + (size_t)g * L.stage_bytes;
                const int64_t voff = 72 * b0, voff_al = voff & ~(int64_t)15;
                const uint32_t vbytes = (uint32_t)(((voff - voff_al) + 72 * (b1 - b0) + 15) & ~(int64_t)15);
                const int64_t coff = 4 * b0, coff_al = coff & ~(int64_t)15;
                const uint32_t cbytes = (uint32_t)(((coff - coff_al) + 4 * (b1 - b0) + 15) & ~(int64_t)15);
                const uint32_t rbytes = (uint32_t)((4 * (r1 - r0 + 1) + 15) & ~(int64_t)15);
                mbar_expect_tx(&s_full[g], vbytes + cbytes + rbytes);
                bulk_g2s(base + L.vals_off, (const unsigned char *)vals + voff_al, vbytes, &s_full[g]);
                bulk_g2s(base + L.cols_off, (const unsigned char *)bcol + coff_al, cbytes, &s_full[g]);
                bulk_g2s(base + L.rp_off, (const unsigned char *)(brow_ptr + r0), rbytes, &s_full[g]);
                tile_a = tile_b; tile_b = tile_c;
                r0 = nr0; r1 = nr1; b0 = nb0; b1 = nb1;
            }
        }
    } else {
        const int g = warp / T3_GW, wg = warp % T3_GW;     // ---- consumers ----
        const int part = wg / 3, t = (wg % 3) * 32 + lane; // scalar row of the tile
        const int br = t / 3, a = t - 3 * br;
        const unsigned char *base = s_raw + (size_t)g * L.stage_bytes;
        const int32_t *rp = reinterpret_cast<const int32_t *>(base + L.rp_off);
        bool halo_ready = !HALO;
        if (HALO) {
            // Push: all 27 consumer warps, while the producer's first bulk copies are still in flight.  A thread
            // handles about one entry: one index load, one x load, one NVLink store.
            constexpr int NC = 32 * T3_GROUPS * T3_GW;
            const CommDev *cd = ha.cd;
            const int W = cd->world, me = cd->rank;
            for (int peer = 0; peer < W; peer++) {
                const long long lo3 = 3 * cd->send_off[peer], hi3 = 3 * cd->send_off[peer + 1];
                if (hi3 == lo3) continue;
                double *dst = cd->vec[peer][ha.vec_id] + cd->tail_off[peer] - lo3;
                const int32_t *rows = cd->send_rows;
                for (long long e = lo3 + (long long)blockIdx.x * NC + tid; e < hi3; e += (long long)NC * gridDim.x) {
                    const long long n = e / 3;
                    dst[e] = x[3 * (long long)rows[n] + (e - 3 * n)];
                }
            }
            // one system-scope fence per CTA, by the thread that takes the ticket, after a barrier that orders
            // every consumer's stores before it (fences are cumulative; 864 threads fencing individually cost
            // ~60 us per launch)
            asm volatile("bar.sync 5, %0;" ::"n"(NC) : "memory");
            if (tid == 0) {
                __threadfence_system();
                if (atomicInc(cd->ticket, gridDim.x - 1) == gridDim.x - 1) {
                    __threadfence_system();                // acquire side: every CTA fenced before its ticket
                    for (int peer = 0; peer < W; peer++)
                        if (peer != me && cd->send_off[peer + 1] > cd->send_off[peer])
                            *(volatile unsigned long long *)&cd->ctrl[peer]->hflag[me] = expect;
                }
            }
        }
        for (int64_t i = g, k = 0; i < my_n; i += T3_GROUPS, k++) {
            if (HALO && !halo_ready && blockIdx.x + i * (int64_t)gridDim.x >= ha.n_interior) {
                const CommDev *cd = ha.cd;                 // first tile with halo columns: wait for the neighbours
                if (lane < cd->n_recv_peers) {
                    const unsigned long long *f = &cd->ctrl[cd->rank]->hflag[cd->recv_peer[lane]];
                    const long long t0 = clock64();
                    unsigned long long seen;
                    do {
                        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(f) : "memory");
                        if (clock64() - t0 > 8000000000LL) { atomicOr(cd->err + 4, 1); break; }   // ~4 s: a peer died
                    } while (seen < expect);
                }
                __syncwarp();
                halo_ready = true;
            }
            mbar_wait(&s_full[g], (uint32_t)(k & 1));
            const int64_t r0 = (HALO ? (int64_t)s_tile[g] : blockIdx.x + i * (int64_t)gridDim.x) * T3_ROWS;
            const int nr = (int)((nrows - r0) < T3_ROWS ? (nrows - r0) : T3_ROWS);
            double acc = 0.0;
            if (br < nr) {
                const int b0 = rp[0];
                const int sblk = rp[br] - b0, nb = rp[br + 1] - rp[br];
                const int per = (nb + T3_PARTS - 1) / T3_PARTS;
                const int c0 = part * per, c1 = (c0 + per < nb) ? c0 + per : nb;
                const double *v = reinterpret_cast<const double *>(base + L.vals_off) + ((9 * (int64_t)b0) & 1) +
                                  9 * sblk + a * 3 * nb;
                const int32_t *cols = reinterpret_cast<const int32_t *>(base + L.cols_off) + (b0 & 3) + sblk;
                double acc1 = 0.0, acc2 = 0.0;
#pragma unroll 3
                for (int c = c0; c < c1; c++) {
                    const int64_t xo = 3 * (int64_t)cols[c];
                    acc += v[3 * c] * X_AT(xo);
                    acc1 += v[3 * c + 1] * X_AT(xo + 1);
                    acc2 += v[3 * c + 2] * X_AT(xo + 2);
                }
                acc = (acc + acc1) + acc2;
            }
            s_part[g][part][t] = acc;
            asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "r"(32 * T3_GW) : "memory");
            if (part == 0 && br < nr) {
                const double sum = (s_part[g][0][t] + s_part[g][1][t]) + s_part[g][2][t];
                const int64_t dof = 3 * (r0 + br) + a;
                y[dof] = sum;
                if (DOT) dsum += sum * X_AT(dof);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[g]);
        }
    }
#undef X_AT
    if (DOT) {
        double v[1] = {dsum};
        const bool last = grid_reduce<1>(v, partials, counter, st, slot, step, run_scalar);
        if (HALO && last && tid == 0) st->halo_seq = expect;          // this exchange is complete on this rank
    }
}

// ---- vector kernels --------------------------------------------------------------------------
// init: x = 0, r = b, p = z = b d^2; sums b.b and r.z
__global__ void __launch_bounds__(VEC_THREADS)
k_cg_init(int64_t n, const double *__restrict__ b, const double *__restrict__ d2, double *__restrict__ x,
          double *__restrict__ r, double *__restrict__ p, double *partials, unsigned int *counter, CgState *st,
          bool run_scalar) {
    double v[2] = {0.0, 0.0};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double bi = b[i], z = bi * d2[i];
        x[i] = 0.0;
        r[i] = bi;
        p[i] = z;
        v[0] += bi * bi;
        v[1] += bi * z;
    }
    grid_reduce<2>(v, partials, counter, st, 0, SC_INIT, run_scalar);
}

// r -= alpha mv; sums r.r and r.(r d^2).  The matching x += alpha p is applied by k_direction, which
// reads p anyway (one vector less per iteration), or by k_x_tail when this iteration ends the solve.
__global__ void __launch_bounds__(VEC_THREADS)
k_update(int64_t n, double *__restrict__ r, const double *__restrict__ mv, const double *__restrict__ d2,
         double *partials, unsigned int *counter, CgState *st, bool run_scalar) {
    if (st->done) return;
    if (blockIdx.x == 0 && threadIdx.x == 0) trace_mark(st, TR_UPDATE_BEGIN);
    const double alpha = st->alpha;
    double v[2] = {0.0, 0.0};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double ri = r[i] - alpha * mv[i];
        r[i] = ri;
        v[0] += ri * ri;
        v[1] += (ri * d2[i]) * ri;
    }
    grid_reduce<2>(v, partials, counter, st, 0, SC_AFTER_UPDATE, run_scalar);
}

// refresh iteration, first half: cx = x + alpha p (the accepted x is left untouched)
__global__ void __launch_bounds__(VEC_THREADS)
k_candidate(int64_t n, const double *__restrict__ x, const double *__restrict__ p, double *__restrict__ cx,
            const CgState *st) {
    if (st->done) return;
    const double alpha = st->alpha;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        cx[i] = x[i] + alpha * p[i];
}

// refresh iteration, second half: r = b - A cx; sums r.r, r.z, 2 b.cx
__global__ void __launch_bounds__(VEC_THREADS)
k_refresh(int64_t n, const double *__restrict__ b, const double *__restrict__ mv, const double *__restrict__ cx,
          const double *__restrict__ d2, double *__restrict__ r, double *partials, unsigned int *counter,
          CgState *st, bool run_scalar) {
    if (st->done) return;
    if (blockIdx.x == 0 && threadIdx.x == 0) trace_mark(st, TR_REFRESH_BEGIN);
    double v[3] = {0.0, 0.0, 0.0};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double ri = b[i] - mv[i];
        r[i] = ri;
        v[0] += ri * ri;
        v[1] += (ri * d2[i]) * ri;
        v[2] += 2 * b[i] * cx[i];
    }
    // slots 0..2 are written here; slot 3 ((A cx).cx) was produced by the SpMV before this kernel
    grid_reduce<3>(v, partials, counter, st, 0, SC_AFTER_REFRESH, run_scalar);
}

// x += alpha p (ordinary iterations; a refresh iteration has already installed its candidate x), then
// p = r d^2 + beta p
// PUSH (several GPUs, peer-memory mode): the thread that writes a boundary entry of p also stores it into the halo
// tails of the ranks that read it (per-row send map), the last CTA raises the sequence flags and waits for the
// neighbours' — the halo exchange of the next product costs no launch of its own (device timeline: the stand-alone
// push kernel and the kernel boundary around it were 16 of ~530 us per iteration on 8 GPUs).
template <bool WITH_X, bool PUSH>
__global__ void __launch_bounds__(VEC_THREADS)
k_direction(int64_t n, const double *__restrict__ r, const double *__restrict__ d2, double *__restrict__ p,
            double *__restrict__ x, CgState *st, const CommDev *__restrict__ cd) {
    if (st->done) return;
    if (blockIdx.x == 0 && threadIdx.x == 0) trace_mark(st, TR_DIRECTION_BEGIN);
    const double beta = st->beta, alpha = st->alpha;
    bool pushed = false;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double pi = p[i];
        if (WITH_X) x[i] = x[i] + alpha * pi;
        const double pn = r[i] * d2[i] + beta * pi;
        p[i] = pn;
        if (PUSH) {
            const int64_t node = i / 3;
            const int c = cd->send_map[node];
            if (c >= 0) {
                pushed = true;
                const int a = (int)(i - 3 * node);
                for (int e = cd->send_ptr[c]; e < cd->send_ptr[c + 1]; e++) {
                    const unsigned long long d = cd->send_dst[e];
                    cd->vec[d >> 48][0][(d & 0xffffffffffffull) + a] = pn;
                }
            }
        }
    }
    if (PUSH) {
        const unsigned long long seq = st->halo_seq + 1;   // (read before any CTA can be last)
        __shared__ bool last;
        const bool any_pushed = __syncthreads_or(pushed);
        if (threadIdx.x == 0) {                            // one fence per CTA that stored remotely, cumulative over its stores
            if (any_pushed) __threadfence_system();
            last = (atomicInc(cd->ticket, gridDim.x - 1) == gridDim.x - 1);
        }
        __syncthreads();
        if (last && threadIdx.x < 32) {
            const int W = cd->world, me = cd->rank, lane = threadIdx.x;
            if (lane == 0) trace_mark(st, TR_PUSH_FLAGS);
            if (lane < W && lane != me && cd->send_off[lane + 1] > cd->send_off[lane]) {
                // acquire side of the ticket (the other CTAs fenced before taking theirs) and release side of the flag
                __threadfence_system();
                *(volatile unsigned long long *)&cd->ctrl[lane]->hflag[me] = seq;
            }
            if (lane < cd->n_recv_peers) {
                volatile unsigned long long *f = &cd->ctrl[me]->hflag[cd->recv_peer[lane]];
                const long long t0 = clock64();
                while (*f < seq) {
                    if (clock64() - t0 > 8000000000LL) { atomicOr(cd->err + 4, 1); break; }   // ~4 s: a peer died
                }
                __threadfence_system();
            }
            __syncwarp();
            if (lane == 0) { st->halo_seq = seq; trace_mark(st, TR_WAIT_END); }
        }
    }
}

// the x += alpha p of the iteration that ended the solve (p and alpha are frozen once the state says done)
__global__ void __launch_bounds__(VEC_THREADS)
k_x_tail(int64_t n, double *__restrict__ x0, double *__restrict__ x1, const double *__restrict__ p, const CgState *st) {
    if (!st->x_pending) return;
    double *x = st->x_in_alt ? x1 : x0;
    const double alpha = st->alpha;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        x[i] = x[i] + alpha * p[i];
}

}  // namespace

int64_t spmv_algorithmic_bytes(const stan_handle *h) {
    // values 8 B each (9 per block) + one int32 column per block + per row: 2 brow_ptr reads
    // amortised to 4 B, 3 x reads and 3 y writes of 8 B (SURVEY §8d formula for the stored format)
    const int64_t nloc = h->row1 - h->row0;
    return 72 * h->n_blocks + 4 * h->n_blocks + nloc * (4 + 24 + 24);
}

// grids are sized to exactly one resident wave (SM count x CTAs that fit per SM) and loop
template <typename K>
static int resident_grid(const stan_handle *h, K kernel, int threads, size_t smem, int64_t want) {
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem);
    if (per_sm < 1) per_sm = 1;
    const int64_t cap = (int64_t)h->sm_count * per_sm;
    return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

static int vec_grid(const stan_handle *h, int64_t n) {
    return resident_grid(h, k_update, VEC_THREADS, 0, (n + VEC_THREADS * 2 - 1) / (VEC_THREADS * 2));
}

static int spmv_variant() {
    static int v = -1;
    if (v < 0) {
        // profiles/r01_spmv_variants.md, 10M beam:
        // 4 = bulk-copy tiles, 3 threads per scalar row (default, 6.9 TB/s; falls back to 0 when a
        //     32-row tile does not fit shared memory)
        // 0 = warp-per-row LDG kernel (5.3 TB/s), 1 = bulk-copy ring + warp per row (5.0 TB/s),
        // 2/3 = bulk-copy tiles, 1 thread per scalar row (5.5 / 4.8 TB/s)
        const char *e = getenv("STAN_SPMV");
        v = e ? atoi(e) : 4;
    }
    return v;
}

static BulkLayout bulk_layout(const stan_handle *h, int tile_rows = 16) {
    BulkLayout L;
    int mb = tile_rows == 32 ? h->max_group_blocks : h->max_group16;
    if (mb < 1) mb = 1;
    auto up = [](int v) { return (v + 127) & ~127; };
    L.vals_off = 0;
    L.cols_off = up(72 * mb + 16);
    L.rp_off = L.cols_off + up(4 * mb + 32);
    L.stage_bytes = L.rp_off + 256;                       // row pointers: up to 33 ints rounded to 16 B
    static int ctas = getenv("STAN_BULK_CTAS") ? atoi(getenv("STAN_BULK_CTAS")) : 1;
    int st = (int)((200 * 1024 / ctas) / L.stage_bytes);
    L.stages = st > 8 ? 8 : st;
    return L;
}

struct SpmvPlan { int variant, grid, minb; BulkLayout L; size_t smem; };

static int spmv_plan(const stan_handle *h, int64_t nrows, SpmvPlan *p) {
    p->variant = spmv_variant();
    p->L = bulk_layout(h);
    if (p->variant == 1 && p->L.stages < 2) p->variant = 0;   // rows too wide for the shared-memory ring
    if (p->variant == 4) {                                    // 32-row tiles x 3 groups, 3 threads per scalar row
        p->L = bulk_layout(h, T3_ROWS);
        p->L.stages = T3_GROUPS;
        p->smem = (size_t)T3_GROUPS * p->L.stage_bytes;
        if (p->smem > 215 * 1024) p->variant = 0;
        else {
            STAN_CUDA((cudaFuncSetAttribute(k_spmv_tile3<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem)));
            STAN_CUDA((cudaFuncSetAttribute(k_spmv_tile3<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem)));
            STAN_CUDA((cudaFuncSetAttribute(k_spmv_tile3<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem)));
            const int64_t nchunks = (nrows + T3_ROWS - 1) / T3_ROWS;
            p->grid = (int)(nchunks < h->sm_count ? (nchunks > 0 ? nchunks : 1) : h->sm_count);
            return STAN_OK;
        }
    }
    if (p->variant == 2 || p->variant == 3) {                 // 2: 16-row tiles x 6 groups, 3: 32-row tiles x 3 groups
        const int tr = p->variant == 2 ? 16 : 32, groups = p->variant == 2 ? 6 : 3;
        p->L = bulk_layout(h, tr);
        p->L.stages = groups;
        p->smem = (size_t)groups * p->L.stage_bytes;
        if (p->smem > 220 * 1024) p->variant = 0;
        else {
            if (tr == 16) {
                STAN_CUDA(cudaFuncSetAttribute(k_spmv_tile<true, 16, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem));
                STAN_CUDA(cudaFuncSetAttribute(k_spmv_tile<false, 16, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem));
            } else {
                STAN_CUDA(cudaFuncSetAttribute(k_spmv_tile<true, 32, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem));
                STAN_CUDA(cudaFuncSetAttribute(k_spmv_tile<false, 32, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem));
            }
            const int64_t nchunks = (nrows + tr - 1) / tr;
            p->grid = (int)(nchunks < h->sm_count ? (nchunks > 0 ? nchunks : 1) : h->sm_count);
            return STAN_OK;
        }
    }
    if (p->variant == 1) {
        p->smem = (size_t)p->L.stages * p->L.stage_bytes;
        STAN_CUDA(cudaFuncSetAttribute(k_spmv_bulk<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem));
        STAN_CUDA(cudaFuncSetAttribute(k_spmv_bulk<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem));
        const int64_t nchunks = (nrows + BULK_ROWS - 1) / BULK_ROWS;
        const int64_t cap = (int64_t)h->sm_count * (getenv("STAN_BULK_CTAS") ? atoi(getenv("STAN_BULK_CTAS")) : 1);
        p->grid = (int)(nchunks < cap ? (nchunks > 0 ? nchunks : 1) : cap);
    } else {
        p->smem = 0;
        // 6 CTAs/SM (<= 40 registers, 48 warps): 5.27 TB/s vs 5.00 at 4 CTAs/SM on the 10M beam
        static int minb = getenv("STAN_SPMV_MINB") ? atoi(getenv("STAN_SPMV_MINB")) : 6;
        p->minb = minb;
        const int64_t want = (nrows + SPMV_WARPS - 1) / SPMV_WARPS;
        p->grid = minb == 8 ? resident_grid(h, k_spmv<true, 8>, SPMV_THREADS, 0, want)
                : minb == 6 ? resident_grid(h, k_spmv<true, 6>, SPMV_THREADS, 0, want)
                            : resident_grid(h, k_spmv<true, 4>, SPMV_THREADS, 0, want);
    }
    return STAN_OK;
}

// ha != nullptr: the product exchanges the halo of `in` itself (variant 4, peer-memory mode, inside a solve)
static void launch_spmv(const stan_handle *h, const SpmvPlan &p, bool dot, int64_t nrows, const double *in, double *out,
                        double *partials, unsigned int *counter, CgState *st, int slot, int step, bool run_scalar,
                        cudaStream_t s, const HaloArgs *ha = nullptr) {
#define STAN_LAUNCH_TILE(D, R, G)                                                                                  \
    k_spmv_tile<D, R, G><<<p.grid, 32 * (G * ((3 * R + 31) / 32) + 1), p.smem, s>>>(                               \
        nrows, h->d_brow_ptr.p, h->bcol_x, h->d_vals.p, in, out, p.L, partials, counter, st, slot, step, run_scalar)
    if (p.variant == 4) {
        const HaloArgs none = {nullptr, nullptr, 0, 0};
        if (ha)
            k_spmv_tile3<true, true><<<p.grid, T3_THREADS, p.smem, s>>>(nrows, h->d_brow_ptr.p, h->bcol_x, h->d_vals.p, in, out,
                                                                        p.L, partials, counter, st, slot, step, run_scalar, *ha);
        else if (dot)
            k_spmv_tile3<true, false><<<p.grid, T3_THREADS, p.smem, s>>>(nrows, h->d_brow_ptr.p, h->bcol_x, h->d_vals.p, in, out,
                                                                         p.L, partials, counter, st, slot, step, run_scalar, none);
        else
            k_spmv_tile3<false, false><<<p.grid, T3_THREADS, p.smem, s>>>(nrows, h->d_brow_ptr.p, h->bcol_x, h->d_vals.p, in, out,
                                                                          p.L, partials, counter, st, slot, step, run_scalar, none);
        return;
    }
    if (p.variant == 2) { if (dot) STAN_LAUNCH_TILE(true, 16, 6); else STAN_LAUNCH_TILE(false, 16, 6); return; }
    if (p.variant == 3) { if (dot) STAN_LAUNCH_TILE(true, 32, 3); else STAN_LAUNCH_TILE(false, 32, 3); return; }
#undef STAN_LAUNCH_TILE
    if (p.variant == 1) {
        if (dot)
            k_spmv_bulk<true><<<p.grid, BULK_THREADS, p.smem, s>>>(nrows, h->d_brow_ptr.p, h->bcol_x, h->d_vals.p, in, out,
                                                                   p.L, partials, counter, st, slot, step, run_scalar);
        else
            k_spmv_bulk<false><<<p.grid, BULK_THREADS, p.smem, s>>>(nrows, h->d_brow_ptr.p, h->bcol_x, h->d_vals.p, in, out,
                                                                    p.L, partials, counter, st, slot, step, run_scalar);
    } else {
#define STAN_LAUNCH_LDG(D, M)                                                                                     \
    k_spmv<D, M><<<p.grid, SPMV_THREADS, 0, s>>>(nrows, nullptr, h->d_brow_ptr.p, h->bcol_x, h->d_vals.p, in, out, \
                                                 partials, counter, st, slot, step, run_scalar)
        if (p.minb == 8) { if (dot) STAN_LAUNCH_LDG(true, 8); else STAN_LAUNCH_LDG(false, 8); }
        else if (p.minb == 6) { if (dot) STAN_LAUNCH_LDG(true, 6); else STAN_LAUNCH_LDG(false, 6); }
        else { if (dot) STAN_LAUNCH_LDG(true, 4); else STAN_LAUNCH_LDG(false, 4); }
#undef STAN_LAUNCH_LDG
    }
}

int solve_cg(stan_handle *h, const stan_cg_options *o, stan_cg_report *rep) {
    cudaStream_t s = h->stream;
    const int64_t nloc = h->row1 - h->row0, n = 3 * nloc, nx = 3 * (h->nloc_pad + h->n_halo);
    const bool multi = h->world > 1;
    const bool p2p = multi && comm_p2p_active(h);
    const bool single = !multi || p2p;      // reductions finish inside the producing kernel
    double epsf = o->epsf;
    if (epsf == 0.0 && o->maxits == 0) epsf = 1.0e-6;      // lincgsetcond note (SolverFunctions.cs:292-293)
    const int rupd = o->its_before_rupdate;
    const int64_t n_global_free = 3 * h->n_nodes - h->n_fixed;
    const int64_t restart = o->its_before_restart > 0 ? o->its_before_restart : (n_global_free > 0 ? n_global_free : 1);
    const int off = o->zero_based_counter ? 1 : 0;

    // the vectors a product reads (p, x, xalt): [owned | pad | halo]; in peer-memory mode they live in the window
    double *vp = nullptr, *vx = nullptr, *vxalt = nullptr;
    STAN_TRY(comm_cg_vectors(h, &vp, &vx, &vxalt, s));
    STAN_TRY(h->d_r.alloc(n, s)); STAN_TRY(h->d_mv.alloc(n, s));
    SpmvPlan plan;
    STAN_TRY(spmv_plan(h, nloc, &plan));
    const int gv = vec_grid(h, n), gs = plan.grid;
    const int gmax = gv > gs ? gv : gs;
    STAN_TRY(h->d_partials.alloc((size_t)gmax * 4, s));
    STAN_TRY(h->d_state.alloc(1, s));
    STAN_TRY(h->d_counter.alloc(4, s));
    STAN_CUDA(cudaMemsetAsync(h->d_counter.p, 0, 4 * sizeof(unsigned int), s));
    if (h->n_halo && !p2p) {                 // (peer-memory mode: the tails belong to the peers, nothing to clear)
        STAN_CUDA(cudaMemsetAsync(vp + n, 0, (nx - n) * sizeof(double), s));
        STAN_CUDA(cudaMemsetAsync(vx + n, 0, (nx - n) * sizeof(double), s));
        STAN_CUDA(cudaMemsetAsync(vxalt + n, 0, (nx - n) * sizeof(double), s));
    }
    CgState init;
    memset(&init, 0, sizeof init);
    init.epsf_bnorm = epsf;
    init.maxits = o->maxits; init.rupdate = rupd; init.merit_check = o->merit_check; init.counter_off = off;
    init.restart = restart;
    init.comm = p2p ? comm_dev(h) : nullptr;
    init.red_seq = h->red_seq;              // flags in the peer windows persist across solves
    init.halo_seq = p2p ? comm_next_epoch(h) << 32 : 0;   // a new epoch: flags of earlier solves can never satisfy a wait
    if (h->hist_cap > 0) {
        STAN_TRY(h->d_hist.alloc((size_t)4 * h->hist_cap, s));
        init.hist = h->d_hist.p;
        init.hist_cap = h->hist_cap;
    }
    const char *tr_env = getenv("STAN_CG_TRACE");          // profiling aid: device-side timeline of the loop
    const int tr_cap = tr_env ? atoi(tr_env) : 0;
    if (tr_cap > 0) {
        STAN_TRY(h->d_trace.alloc((size_t)2 * tr_cap, s));
        init.trace = h->d_trace.p;
        init.trace_cap = tr_cap;
        init.trace_from = getenv("STAN_CG_TRACE_FROM") ? atoi(getenv("STAN_CG_TRACE_FROM")) : 100;
    }
    if (!h->h_state) STAN_CUDA(cudaMallocHost((void **)&h->h_state, sizeof(CgState)));
    CgState *hst = h->h_state;
    *hst = init;
    STAN_CUDA(cudaMemcpyAsync(h->d_state.p, hst, sizeof(CgState), cudaMemcpyHostToDevice, s));
    CgState *st = h->d_state.p;
    double *x = vx, *xalt = vxalt;
    // STAN_FUSED_HALO=1: the product kernel exchanges its own halo, overlapped with the tiles that need none.
    // Measured (profiles/r02_multi_gpu_iteration.md) it is within +-1 % of the two small push / wait launches on
    // 2 and 8 GPUs — the iteration is bound by the three cross-rank synchronisations, not by the halo bytes —
    // and 1 % slower on 8, so the stand-alone kernels are the default.
    const char *fh = getenv("STAN_FUSED_HALO");
    const bool fused_halo = p2p && plan.variant == 4 && fh && atoi(fh) == 1;
    // STAN_DIR_PUSH=0: stand-alone push / wait launch before every product (A/B timing)
    const char *dp = getenv("STAN_DIR_PUSH");
    const bool dir_push = p2p && !fused_halo && !(dp && atoi(dp) == 0);
    const CommDev *cdp = comm_dev_ptr(h);
    int64_t launches = 0;
    int spmv_launches = 0;
    float spmv_ms = 0.f;
    const bool timek = o->time_kernels != 0;
    std::vector<cudaEvent_t> evs;

    STAN_CUDA(cudaEventRecord(h->ev0, s));
    auto reduce_tail = [&](int step) -> int {              // multi-GPU: all-reduce the sums, then one scalar thread
        if (single) return STAN_OK;   // one GPU, or peer-memory mode: cross_rank_sum ran in the kernel
        const int cnt = step == SC_AFTER_SPMV ? 1 : (step == SC_AFTER_REFRESH ? 4 : 2);
        STAN_TRY(comm_allreduce_sum(h, (double *)((char *)st + offsetof(CgState, partial)), cnt, s));
        k_scalar<<<1, 1, 0, s>>>(st, step);
        launches++;
        return STAN_OK;
    };
    auto spmv = [&](double *in, int slot, int step) -> int {
        const int vec_id = in == vp ? 0 : (in == vx ? 1 : 2);
        HaloArgs ha;
        const bool fuse = fused_halo && comm_halo_args(h, vec_id, &ha);
        // p's halo was exchanged by the kernel that wrote it (k_direction<.., PUSH>, or the explicit exchange after k_cg_init)
        if (multi && !fuse && !(dir_push && in == vp)) STAN_TRY(comm_halo_exchange(h, in, vec_id, s, st));
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        if (timek) {                                       // events come from a grow-only pool kept on the handle
            if (h->ev_pool.size() < evs.size() + 2) {
                h->ev_pool.resize(evs.size() + 2, nullptr);
                cudaEventCreate(&h->ev_pool[evs.size()]); cudaEventCreate(&h->ev_pool[evs.size() + 1]);
            }
            e0 = h->ev_pool[evs.size()]; e1 = h->ev_pool[evs.size() + 1];
            cudaEventRecord(e0, s);
        }
        launch_spmv(h, plan, true, nloc, in, h->d_mv.p, h->d_partials.p, h->d_counter.p + 2, st, slot, step, single, s,
                    fuse ? &ha : nullptr);
        if (timek) { cudaEventRecord(e1, s); evs.push_back(e0); evs.push_back(e1); }
        launches++; spmv_launches++;
        if (step != SC_NONE) STAN_TRY(reduce_tail(step));
        return STAN_OK;
    };

    k_cg_init<<<gv, VEC_THREADS, 0, s>>>(n, h->d_b.p, h->d_d2.p, x, h->d_r.p, vp, h->d_partials.p,
                                         h->d_counter.p, st, single);
    launches++;
    STAN_TRY(reduce_tail(SC_INIT));
    if (dir_push) STAN_TRY(comm_halo_exchange(h, vp, 0, s, st));   // p of k_cg_init; afterwards k_direction pushes what it writes

    // One batch = two refresh periods, so the x / xalt roles are back where they started and every batch
    // is the same launch sequence.  On one GPU without per-launch timing the batch is captured once into
    // a CUDA graph and replayed: the loop is launch-bound on small systems (100k elements: 17 of 67 us
    // per iteration were gaps between kernels).  Kernels enqueued after the state flags `done` return
    // immediately, so replaying a whole batch past the end is harmless.
    const int batch = 2 * (rupd > 0 ? rupd : 10);
    int k = 0;
    auto enqueue_batch = [&]() -> int {
        for (int bi = 0; bi < batch; bi++) {
            k++;
            const int kk = k - off;
            const bool refresh = rupd > 0 && kk % rupd == 0;
            STAN_TRY(spmv(vp, 0, SC_AFTER_SPMV));
            if (!refresh) {
                k_update<<<gv, VEC_THREADS, 0, s>>>(n, h->d_r.p, h->d_mv.p, h->d_d2.p, h->d_partials.p, h->d_counter.p, st,
                                                    single);
                launches++;
                STAN_TRY(reduce_tail(SC_AFTER_UPDATE));
            } else {
                k_candidate<<<gv, VEC_THREADS, 0, s>>>(n, x, vp, xalt, st);
                launches++;
                STAN_TRY(spmv(xalt, 3, SC_NONE));
                k_refresh<<<gv, VEC_THREADS, 0, s>>>(n, h->d_b.p, h->d_mv.p, xalt, h->d_d2.p, h->d_r.p,
                                                     h->d_partials.p, h->d_counter.p, st, single);
                launches++;
                STAN_TRY(reduce_tail(SC_AFTER_REFRESH));
                double *t = x; x = xalt; xalt = t;         // accepted unless the state says type 7
            }
            if (dir_push) {
                if (refresh) k_direction<false, true><<<gv, VEC_THREADS, 0, s>>>(n, h->d_r.p, h->d_d2.p, vp, x, st, cdp);
                else         k_direction<true, true><<<gv, VEC_THREADS, 0, s>>>(n, h->d_r.p, h->d_d2.p, vp, x, st, cdp);
            } else {
                if (refresh) k_direction<false, false><<<gv, VEC_THREADS, 0, s>>>(n, h->d_r.p, h->d_d2.p, vp, x, st, nullptr);
                else         k_direction<true, false><<<gv, VEC_THREADS, 0, s>>>(n, h->d_r.p, h->d_d2.p, vp, x, st, nullptr);
            }
            launches++;
        }
        return STAN_OK;
    };
    static const bool graphs_on = !(getenv("STAN_GRAPH") && atoi(getenv("STAN_GRAPH")) == 0);
    // replayable when no launch argument changes between batches: one GPU, or peer-memory mode (all sequence
    // numbers live in the device state; the NCCL data plane is enqueued call by call)
    const bool use_graph = graphs_on && !timek && (!multi || p2p);
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t gexec = nullptr;
    int64_t launches_per_batch = 0;
    int spmv_per_batch = 0;
    if (use_graph) {
        const int64_t l0 = launches;
        const int s0 = spmv_launches;
        STAN_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        const int rc = enqueue_batch();
        cudaError_t ce = cudaStreamEndCapture(s, &graph);
        if (rc != STAN_OK) return rc;
        STAN_CUDA(ce);
        STAN_CUDA(cudaGraphInstantiate(&gexec, graph, 0));
        launches_per_batch = launches - l0;
        spmv_per_batch = spmv_launches - s0;
        launches = l0; spmv_launches = s0;                  // counted per replay below
    }
    bool done = false;
    while (!done) {
        if (use_graph) {
            STAN_CUDA(cudaGraphLaunch(gexec, s));
            launches += launches_per_batch;
            spmv_launches += spmv_per_batch;
        } else {
            STAN_TRY(enqueue_batch());
        }
        STAN_CUDA(cudaMemcpyAsync(hst, st, sizeof(CgState), cudaMemcpyDeviceToHost, s));
        STAN_CUDA(cudaStreamSynchronize(s));
        STAN_CUDA(cudaGetLastError());
        done = hst->done != 0;
    }
    if (gexec) cudaGraphExecDestroy(gexec);
    if (graph) cudaGraphDestroy(graph);
    if (hst->x_pending) {
        k_x_tail<<<gv, VEC_THREADS, 0, s>>>(n, vx, vxalt, vp, st);
        launches++;
    }
    STAN_CUDA(cudaEventRecord(h->ev1, s));
    STAN_CUDA(cudaStreamSynchronize(s));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->ev0, h->ev1);
    for (size_t i = 0; i + 1 < evs.size(); i += 2) {
        float t = 0.f;
        // launches enqueued after the state flagged done return immediately; they are still counted
        cudaEventElapsedTime(&t, evs[i], evs[i + 1]);
        spmv_ms += t;
    }
    // accepted iterate: d_x unless an odd number of refreshes were accepted
    h->x_in_alt = hst->x_in_alt != 0;
    h->sol = h->x_in_alt ? vxalt : vx;
    h->red_seq = hst->red_seq;
    h->hist_count = hst->k < h->hist_cap ? hst->k : h->hist_cap;
    if (tr_cap > 0) {                                      // dump: one "ns,iteration,code" line per event
        const int n_ev = hst->trace_n < tr_cap ? hst->trace_n : tr_cap;
        std::vector<unsigned long long> ev((size_t)2 * n_ev);
        if (n_ev) STAN_CUDA(cudaMemcpy(ev.data(), h->d_trace.p, ev.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        const char *dir = getenv("STAN_CG_TRACE_DIR");
        char path[512];
        snprintf(path, sizeof path, "%s/stan_cg_trace_rank%d.csv", dir ? dir : ".", h->rank);
        if (FILE *f = fopen(path, "w")) {
            for (int i = 0; i < n_ev; i++)
                fprintf(f, "%llu,%llu,%llu\n", ev[2 * i], ev[2 * i + 1] >> 8, ev[2 * i + 1] & 0xffull);
            fclose(f);
        }
    }
    rep->terminationtype = hst->type;
    rep->iterationscount = hst->k;
    rep->nmv = hst->nmv;
    rep->spmv_launches = spmv_launches;
    rep->r2 = hst->r2;
    rep->bnorm = hst->bnorm;
    rep->solve_ms = ms;
    rep->spmv_ms = spmv_ms;
    rep->spmv_bytes = spmv_algorithmic_bytes(h);
    rep->iter_bytes = rep->spmv_bytes + 80 * n;
    rep->kernel_launches = launches;
    h->launches += launches;
    if (p2p) {                                             // a peer never raised its flag (spin timed out)
        int32_t herr[8] = {0};
        STAN_CUDA(cudaMemcpy(herr, h->d_err.p, sizeof herr, cudaMemcpyDeviceToHost));
        if (herr[4]) { set_error("peer-memory exchange timed out: a rank stopped participating"); return STAN_E_COMM; }
    }
    h->solved = true;
    return STAN_OK;
}

int spmv_full(stan_handle *h, const double *x_full, double *y_full) {
    if (h->world != 1) { set_error("stan_spmv is single-GPU only"); return STAN_E_STATE; }
    cudaStream_t s = h->stream;
    const int64_t n = 3 * h->n_nodes;
    DevBuf<double> x, y;
    STAN_TRY(x.alloc(n, s)); STAN_TRY(y.alloc(n, s));
    STAN_CUDA(cudaMemcpyAsync(x.p, x_full, n * sizeof(double), cudaMemcpyHostToDevice, s));
    SpmvPlan plan;
    STAN_TRY(spmv_plan(h, h->n_nodes, &plan));
    launch_spmv(h, plan, false, h->n_nodes, x.p, y.p, nullptr, nullptr, nullptr, 0, SC_NONE, false, s);
    STAN_CUDA(cudaGetLastError());
    STAN_CUDA(cudaMemcpyAsync(y_full, y.p, n * sizeof(double), cudaMemcpyDeviceToHost, s));
    STAN_CUDA(cudaStreamSynchronize(s));
    x.release(s); y.release(s);
    h->launches += 1;
    return STAN_OK;
}

int time_spmv(stan_handle *h, int reps, double *ms_out, int64_t *bytes) {
    cudaStream_t s = h->stream;
    const int64_t nloc = h->row1 - h->row0, nx = 3 * (h->nloc_pad + h->n_halo);
    DevBuf<double> x, y;
    STAN_TRY(x.alloc(nx, s)); STAN_TRY(y.alloc(3 * nloc, s));
    STAN_CUDA(cudaMemsetAsync(x.p, 0, nx * sizeof(double), s));
    SpmvPlan plan;
    STAN_TRY(spmv_plan(h, nloc, &plan));
    for (int w = 0; w < 3; w++)
        launch_spmv(h, plan, false, nloc, x.p, y.p, nullptr, nullptr, nullptr, 0, SC_NONE, false, s);
    STAN_CUDA(cudaEventRecord(h->ev2, s));
    for (int r = 0; r < reps; r++)
        launch_spmv(h, plan, false, nloc, x.p, y.p, nullptr, nullptr, nullptr, 0, SC_NONE, false, s);
    STAN_CUDA(cudaEventRecord(h->ev3, s));
    STAN_CUDA(cudaStreamSynchronize(s));
    STAN_CUDA(cudaGetLastError());
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->ev2, h->ev3);
    *ms_out = ms / reps;
    if (bytes) *bytes = spmv_algorithmic_bytes(h);
    x.release(s); y.release(s);
    h->launches += reps + 3;
    return STAN_OK;
}

int scatter_solution(stan_handle *h) {
    cudaStream_t s = h->stream;
    const int64_t nloc = h->row1 - h->row0;
    STAN_TRY(h->d_ufull.alloc(3 * h->n_nodes, s));
    const double *x = h->sol;
    if (h->world == 1) {
        STAN_CUDA(cudaMemcpyAsync(h->d_ufull.p, x, 3 * nloc * sizeof(double), cudaMemcpyDeviceToDevice, s));
    } else {
        STAN_TRY(comm_allgather_rows(h, x, h->d_ufull.p, s));
    }
    return STAN_OK;
}

}  // namespace stan
