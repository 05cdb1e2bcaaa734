// hex8 stiffness integration and deterministic row-owner assembly.
//
// Replaces Element.K_Initial (/root/reference/src/STAN_Database/Element.cs:118-155, with
// Jacobian :274-292, BL0_Matrix :297-328 and the MatrixST products MatrixST.cs:404-427) and the
// locked scatter loop of ParallelAssembly_K (/root/reference/src/STAN_Solver/SolverFunctions.cs:129-174).
//
// Work decomposition: one thread per (row node p, incident element e).  The
// thread integrates only the 3x24 row block of Ke that belongs to p — B_i^T D B_j is evaluated in
// closed form for the isotropic D of Material.cs:39-53,
//     K_ij[a][b] = w|J| ( lambda dNi[a] dNj[b] + G dNi[b] dNj[a] + delta_ab G dNi.dNj ),
// which skips the structural zeros of BL and D that the reference multiplies through.  A CTA owns
// 32 consecutive rows; their CSR storage is contiguous, so it is accumulated in shared memory and
// written once, coalesced.  Contributions to a row are added in ascending element order by
// barrier-separated rounds: no atomics, bitwise reproducible, every matrix value written once.
#include "common.cuh"

namespace stan {

namespace {

// dN_dLocal tables: entries 0..7 = the 2x2x2 points of HEX8_G2 (FE_Library.cs:119-129),
// entry 8 = the centre point of HEX8_G1 (FE_Library.cs:83-87); [point][3][8].
__constant__ double c_dNl[9][24];

constexpr int ROWS_PER_CTA = 32;
constexpr int ASM_THREADS = 256;

__device__ __forceinline__ double det3(const double *m) {
    // term order of MatrixST.Det3 (MatrixST.cs:274-279)
    return m[0] * m[4] * m[8] + m[3] * m[7] * m[2] + m[6] * m[1] * m[5] - m[2] * m[4] * m[6] -
           m[0] * m[5] * m[7] - m[8] * m[1] * m[3];
}

// Row block i (3 x 24) of Ke.  s_tab is a shared-memory copy of c_dNl (the per-thread index i
// would serialise constant-cache reads).  Returns true when a Jacobian determinant is zero.
__device__ __forceinline__ bool hex8_row_block(int type, const double (&X)[24], int i, double lam, double G,
                                               const double *s_tab, double (&K)[72]) {
    const int g0 = (type == STAN_HEX8_G2) ? 0 : 8;
    const int ng = (type == STAN_HEX8_G2) ? 8 : 1;
    const double w = (type == STAN_HEX8_G2) ? 1.0 : 8.0;   // GaussWeight, FE_Library.cs:72,100
#pragma unroll
    for (int k = 0; k < 72; k++) K[k] = 0.0;
    bool bad = false;
    for (int g = g0; g < g0 + ng; g++) {
        double J[9];
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int c = 0; c < 3; c++) {
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < 8; k++) s += c_dNl[g][r * 8 + k] * X[k * 3 + c];
                J[r * 3 + c] = s;
            }
        const double det = det3(J);
        if (det == 0.0) bad = true;
        const double inv = 1.0 / det;
        double Ji[9];                                        // MatrixST.Inverse, MatrixST.cs:303-311
        Ji[0] = inv * (J[4] * J[8] - J[5] * J[7]);
        Ji[1] = inv * (J[2] * J[7] - J[1] * J[8]);
        Ji[2] = inv * (J[1] * J[5] - J[2] * J[4]);
        Ji[3] = inv * (J[5] * J[6] - J[3] * J[8]);
        Ji[4] = inv * (J[0] * J[8] - J[2] * J[6]);
        Ji[5] = inv * (J[2] * J[3] - J[0] * J[5]);
        Ji[6] = inv * (J[3] * J[7] - J[4] * J[6]);
        Ji[7] = inv * (J[1] * J[6] - J[0] * J[7]);
        Ji[8] = inv * (J[0] * J[4] - J[1] * J[3]);
        const double wdet = det * w;
        const double t0 = s_tab[g * 24 + i], t1 = s_tab[g * 24 + 8 + i], t2 = s_tab[g * 24 + 16 + i];
        double l[3], u[3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const double di = Ji[c * 3 + 0] * t0 + Ji[c * 3 + 1] * t1 + Ji[c * 3 + 2] * t2;
            l[c] = lam * wdet * di;
            u[c] = G * wdet * di;
        }
#pragma unroll
        for (int j = 0; j < 8; j++) {
            double dj[3];
#pragma unroll
            for (int c = 0; c < 3; c++)
                dj[c] = Ji[c * 3 + 0] * c_dNl[g][j] + Ji[c * 3 + 1] * c_dNl[g][8 + j] + Ji[c * 3 + 2] * c_dNl[g][16 + j];
            const double s = u[0] * dj[0] + u[1] * dj[1] + u[2] * dj[2];
#pragma unroll
            for (int a = 0; a < 3; a++) {
#pragma unroll
                for (int b = 0; b < 3; b++) K[a * 24 + 3 * j + b] += l[a] * dj[b] + u[b] * dj[a];
                K[a * 24 + 3 * j + a] += s;
            }
        }
    }
    return bad;
}

__device__ __forceinline__ void load_element(const int32_t *__restrict__ conn, const double *__restrict__ xyz,
                                             int64_t e, double (&X)[24]) {
    const int4 c0 = *reinterpret_cast<const int4 *>(conn + 8 * e);
    const int4 c1 = *reinterpret_cast<const int4 *>(conn + 8 * e + 4);
    const int nd[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const double *p = xyz + 3 * (int64_t)nd[k];
        X[3 * k] = p[0]; X[3 * k + 1] = p[1]; X[3 * k + 2] = p[2];
    }
}

__global__ void __launch_bounds__(ASM_THREADS, 1)
k_assemble_rows(int64_t nloc, int64_t row0, const int32_t *__restrict__ inc_ptr, const int32_t *__restrict__ inc,
                const int32_t *__restrict__ brow_ptr, const int32_t *__restrict__ bcol,
                const int32_t *__restrict__ conn, const double *__restrict__ xyz,
                const int32_t *__restrict__ node_index, const uint8_t *__restrict__ etype,
                const int32_t *__restrict__ emat, const double *__restrict__ lam_tab, const double *__restrict__ G_tab,
                const uint8_t *__restrict__ fixed, double *__restrict__ vals, double *__restrict__ d2, int32_t *err) {
    extern __shared__ double s_buf[];
    __shared__ double s_tab[9 * 24];
    __shared__ int s_inc[ROWS_PER_CTA + 1], s_brow[ROWS_PER_CTA + 1];
    const int tid = threadIdx.x;
    const int64_t r0 = (int64_t)blockIdx.x * ROWS_PER_CTA;
    const int nr = (int)((nloc - r0) < ROWS_PER_CTA ? (nloc - r0) : ROWS_PER_CTA);
    if (tid <= nr) { s_inc[tid] = inc_ptr[r0 + tid]; s_brow[tid] = brow_ptr[r0 + tid]; }
    if (tid < 216) s_tab[tid] = (&c_dNl[0][0])[tid];
    __syncthreads();
    const int nvals = 9 * (s_brow[nr] - s_brow[0]);
    for (int t = tid; t < nvals; t += ASM_THREADS) s_buf[t] = 0.0;
    __syncthreads();

    const int ebeg = s_inc[0], eend = s_inc[nr];
    for (int base = ebeg; base < eend; base += ASM_THREADS) {
        const int idx = base + tid;
        const bool active = idx < eend;
        int rl = 0, rank = 0;
        int64_t e = 0;
        double K[72];
        if (active) {
            int lo = 0, hi = nr - 1;                        // last row with s_inc[row] <= idx
            while (lo < hi) {
                int mid = (lo + hi + 1) >> 1;
                if (s_inc[mid] <= idx) lo = mid; else hi = mid - 1;
            }
            rl = lo;
            rank = idx - s_inc[rl];
            const int ent = inc[idx];
            e = ent >> 3;
            double X[24];
            load_element(conn, xyz, e, X);
            const int mat = emat[e];
            if (hex8_row_block(etype[e], X, ent & 7, lam_tab[mat], G_tab[mat], s_tab, K)) atomicOr(err + 2, 1);
        }
        // contributions to one row land in ascending (element, local node) order
        for (int k = 0;; k++) {
            if (active && rank == k) {
                const int64_t p = row0 + r0 + rl;
                const int nb = s_brow[rl + 1] - s_brow[rl];
                double *rowbase = s_buf + 9 * (s_brow[rl] - s_brow[0]);
                const int32_t *cols = bcol + s_brow[rl];
                const bool fa[3] = {fixed[3 * p] != 0, fixed[3 * p + 1] != 0, fixed[3 * p + 2] != 0};
                const int4 c0 = *reinterpret_cast<const int4 *>(conn + 8 * e);
                const int4 c1 = *reinterpret_cast<const int4 *>(conn + 8 * e + 4);
                const int nd[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const int32_t q = node_index[nd[j]];
                    int lo = 0, hi = nb - 1;
                    while (lo < hi) {
                        int mid = (lo + hi) >> 1;
                        if (cols[mid] < q) lo = mid + 1; else hi = mid;
                    }
                    const uint8_t *fq = fixed + 3 * (int64_t)q;
#pragma unroll
                    for (int b = 0; b < 3; b++) {
                        if (fq[b]) continue;                // column of a fixed DOF: dropped (SolverFunctions.cs:160)
#pragma unroll
                        for (int a = 0; a < 3; a++)
                            if (!fa[a]) rowbase[a * 3 * nb + 3 * lo + b] += K[a * 24 + 3 * j + b];
                    }
                }
            }
            if (!__syncthreads_or(active && rank > k)) break;
        }
    }
    __syncthreads();
    // diagonal: identity on fixed rows; Jacobi scaling d^2 = (1/sqrt(A_ii))^2 as ALGLIB forms it
    if (tid < 3 * nr) {
        const int rl = tid / 3, a = tid % 3;
        const int64_t p = row0 + r0 + rl;
        const int nb = s_brow[rl + 1] - s_brow[rl];
        double *rowbase = s_buf + 9 * (s_brow[rl] - s_brow[0]);
        const int32_t *cols = bcol + s_brow[rl];
        int lo = 0, hi = nb - 1;
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (cols[mid] < (int32_t)p) lo = mid + 1; else hi = mid;
        }
        double *dg = rowbase + a * 3 * nb + 3 * lo + a;
        if (fixed[3 * p + a]) *dg = 1.0;
        const double v = *dg;
        const double d = v > 0.0 ? 1.0 / sqrt(v) : 1.0;
        d2[3 * (r0 + rl) + a] = d * d;
    }
    __syncthreads();
    double *out = vals + 9 * (int64_t)s_brow[0];
    for (int t = tid; t < nvals; t += ASM_THREADS) out[t] = s_buf[t];
}

// ---- assembly, two-kernel path: integrate every element once, then gather rows ------------------
// ncu on k_assemble_rows (profiles/r01_path_10m_ncu.md): FP64 pipe 16 %, 14 of 32 lanes active, and every
// element is integrated once per incident node (2.8x the flops).  The two kernels below restore the
// textbook split while staying deterministic:
//
//   k_hex8_ke_batch  one CTA per batch of 8 elements.  Phase 1: thread (element, Gauss point) forms J,
//                    J^-1 and the global shape-function derivatives once and stages dN (8x3) and w|J|
//                    in shared memory.  Phase 2: thread (element, upper block (i <= j)) contracts the
//                    staged derivatives over the Gauss points into one 3x3 block — 9 accumulators, so
//                    the kernel runs at high occupancy — and stores the 36 upper blocks of Ke (2592 B per
//                    element) with fully coalesced writes.  Ke is symmetric by construction.
//   k_assemble_gather  one warp per matrix row, lanes = block slots.  It walks the row's incident
//                    (element, local node i) entries in ascending order, finds which element column j
//                    lands in each lane's slot with eight shuffles, and adds Ke block (i, j) (or the
//                    transpose of (j, i)) — fixed order, no atomics, every stored value written once,
//                    coalesced.  The incidence list is the element-to-slot map.
constexpr int KB_ELEMS = 8;                       // elements per CTA
constexpr int KE_BLK = 10;                        // doubles per stored 3x3 block: 9 values + 1 pad = 80 B, 16-byte aligned
constexpr int KE_ELEM = 36 * KE_BLK;              // doubles per element in the Ke store
constexpr int KB_THREADS = KB_ELEMS * 36;         // one thread per upper block in phase 2
__constant__ unsigned char c_blk_i[36], c_blk_j[36];

__device__ __forceinline__ int upper_index(int i, int j) { return i * 8 - (i * (i - 1)) / 2 + (j - i); }   // i <= j

__global__ void __launch_bounds__(KB_THREADS, 3)
k_hex8_ke_batch(int64_t n_local, const int32_t *__restrict__ lelem, const int32_t *__restrict__ conn,
                const double *__restrict__ xyz, const uint8_t *__restrict__ etype, const int32_t *__restrict__ emat,
                const double *__restrict__ lam_tab, const double *__restrict__ G_tab, double *__restrict__ ke_store,
                int32_t *err) {
    __shared__ double s_dn[KB_ELEMS][8][25];       // per (element, Gauss point): dN[node][xyz] and w|J|
    __shared__ double s_lam[KB_ELEMS], s_G[KB_ELEMS];
    __shared__ int s_ng[KB_ELEMS];
    const int tid = threadIdx.x;
    const int64_t le0 = (int64_t)blockIdx.x * KB_ELEMS;
    if (tid < KB_ELEMS * 8) {                      // ---- phase 1: Jacobians and derivatives, once per (e, g)
        const int el = tid >> 3, g = tid & 7;
        const int64_t le = le0 + el;
        if (le < n_local) {
            const int64_t e = lelem ? lelem[le] : le;
            const int type = etype[e];
            const int ng = (type == STAN_HEX8_G2) ? 8 : 1;
            if (g == 0) { s_ng[el] = ng; const int mat = emat[e]; s_lam[el] = lam_tab[mat]; s_G[el] = G_tab[mat]; }
            if (g < ng) {
                const int gp = (type == STAN_HEX8_G2) ? g : 8;
                const double w = (type == STAN_HEX8_G2) ? 1.0 : 8.0;
                double X[24];
                load_element(conn, xyz, e, X);
                double J[9];
#pragma unroll
                for (int r = 0; r < 3; r++)
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        double sum = 0.0;
#pragma unroll
                        for (int k = 0; k < 8; k++) sum += c_dNl[gp][r * 8 + k] * X[k * 3 + c];
                        J[r * 3 + c] = sum;
                    }
                const double det = det3(J);
                if (det == 0.0) atomicOr(err + 2, 1);
                const double inv = 1.0 / det;
                double Ji[9];
                Ji[0] = inv * (J[4] * J[8] - J[5] * J[7]);
                Ji[1] = inv * (J[2] * J[7] - J[1] * J[8]);
                Ji[2] = inv * (J[1] * J[5] - J[2] * J[4]);
                Ji[3] = inv * (J[5] * J[6] - J[3] * J[8]);
                Ji[4] = inv * (J[0] * J[8] - J[2] * J[6]);
                Ji[5] = inv * (J[2] * J[3] - J[0] * J[5]);
                Ji[6] = inv * (J[3] * J[7] - J[4] * J[6]);
                Ji[7] = inv * (J[1] * J[6] - J[0] * J[7]);
                Ji[8] = inv * (J[0] * J[4] - J[1] * J[3]);
                double *out = s_dn[el][g];
#pragma unroll
                for (int k = 0; k < 8; k++)
#pragma unroll
                    for (int c = 0; c < 3; c++)
                        out[3 * k + c] = Ji[c * 3 + 0] * c_dNl[gp][k] + Ji[c * 3 + 1] * c_dNl[gp][8 + k] + Ji[c * 3 + 2] * c_dNl[gp][16 + k];
                out[24] = det * w;
            }
        }
    }
    __syncthreads();
    // ---- phase 2: one upper block per thread, contracted over the Gauss points ----
    const int el = tid / 36, blk = tid - 36 * el;
    const int64_t le = le0 + el;
    if (le >= n_local) return;
    const int i = c_blk_i[blk], j = c_blk_j[blk];
    const double lam = s_lam[el], G = s_G[el];
    double K[9];
#pragma unroll
    for (int q = 0; q < 9; q++) K[q] = 0.0;
    const int ng = s_ng[el];
    for (int g = 0; g < ng; g++) {
        const double *d = s_dn[el][g];
        const double wdet = d[24];
        const double di0 = d[3 * i], di1 = d[3 * i + 1], di2 = d[3 * i + 2];
        const double dj[3] = {d[3 * j], d[3 * j + 1], d[3 * j + 2]};
        const double l[3] = {lam * wdet * di0, lam * wdet * di1, lam * wdet * di2};
        const double u[3] = {G * wdet * di0, G * wdet * di1, G * wdet * di2};
        const double sdot = u[0] * dj[0] + u[1] * dj[1] + u[2] * dj[2];
#pragma unroll
        for (int a = 0; a < 3; a++) {
#pragma unroll
            for (int b = 0; b < 3; b++) K[3 * a + b] += l[a] * dj[b] + u[b] * dj[a];
            K[3 * a + a] += sdot;
        }
    }
    double *out = ke_store + (le * 36 + blk) * KE_BLK;     // 80-byte slots: 16-byte aligned vector stores / loads
#pragma unroll
    for (int q = 0; q < 8; q += 2) *reinterpret_cast<double2 *>(out + q) = make_double2(K[q], K[q + 1]);
    out[8] = K[8];
}

constexpr int GA_WARPS = 8;
constexpr int GA_GROUP = 4;                       // incidence entries whose blocks are in flight together

__global__ void __launch_bounds__(32 * GA_WARPS, 2)
k_assemble_gather(int64_t nloc, int64_t row0, const int32_t *__restrict__ inc_ptr, const int32_t *__restrict__ inc,
                  const int32_t *__restrict__ brow_ptr, const int32_t *__restrict__ bcol,
                  const int32_t *__restrict__ conn, const int32_t *__restrict__ node_index,
                  const int32_t *__restrict__ g2l, const double *__restrict__ ke_store,
                  const uint8_t *__restrict__ fixed, double *__restrict__ vals, double *__restrict__ d2) {
    const int lane = threadIdx.x & 31;
    const int64_t rl = (int64_t)blockIdx.x * GA_WARPS + (threadIdx.x >> 5);
    if (rl >= nloc) return;
    const int64_t p = row0 + rl;
    const int s0 = brow_ptr[rl], nb = brow_ptr[rl + 1] - s0;
    const int t0 = inc_ptr[rl], t1 = inc_ptr[rl + 1];
    const bool fr[3] = {fixed[3 * p] != 0, fixed[3 * p + 1] != 0, fixed[3 * p + 2] != 0};
    for (int sb = 0; sb < nb; sb += 32) {          // one pass unless the row has more than 32 blocks
        const int s = sb + lane;
        const bool have = s < nb;
        const int32_t myq = have ? bcol[s0 + s] : -1;
        double acc[9];
#pragma unroll
        for (int q = 0; q < 9; q++) acc[q] = 0.0;
        // Entries are taken GA_GROUP at a time.  Pass 1 finds, per entry, which element columns fall into this
        // lane's slot (a bit mask; more than one bit only for degenerate elements that repeat a node) and
        // where the block lives; pass 2 loads the blocks — independent 72-byte reads in flight — and
        // adds them in ascending (element, local node, column) order: the fixed summation order.
        for (int tg = t0; tg < t1; tg += GA_GROUP) {
            const double *blk[GA_GROUP];
            int bits[GA_GROUP], irow[GA_GROUP];
#pragma unroll
            for (int u = 0; u < GA_GROUP; u++) {
                bits[u] = 0; irow[u] = 0; blk[u] = ke_store;
                if (tg + u < t1) {                  // warp-uniform
                    const int ent = inc[tg + u];
                    const int64_t e = ent >> 3;
                    const int32_t qj_mine = node_index[conn[8 * e + (lane & 7)]];
                    int m = 0;
#pragma unroll
                    for (int j = 0; j < 8; j++)
                        if (__shfl_sync(0xffffffffu, qj_mine, j) == myq) m |= 1 << j;
                    bits[u] = have ? m : 0;
                    irow[u] = ent & 7;
                    blk[u] = ke_store + (int64_t)(g2l ? g2l[e] : e) * KE_ELEM;
                }
            }
            double v[GA_GROUP][9];
#pragma unroll
            for (int u = 0; u < GA_GROUP; u++) {
                if (bits[u]) {
                    const int j = __ffs(bits[u]) - 1, i = irow[u];
                    const double *src = blk[u] + (i <= j ? upper_index(i, j) : upper_index(j, i)) * KE_BLK;
#pragma unroll
                    for (int q = 0; q < 8; q += 2) {
                        const double2 t2 = *reinterpret_cast<const double2 *>(src + q);
                        v[u][q] = t2.x; v[u][q + 1] = t2.y;
                    }
                    v[u][8] = src[8];
                }
            }
#pragma unroll
            for (int u = 0; u < GA_GROUP; u++) {
                if (bits[u]) {
                    int rest = bits[u];
                    const int i = irow[u];
                    int j = __ffs(rest) - 1;
                    rest &= rest - 1;
                    if (i <= j) {
#pragma unroll
                        for (int q = 0; q < 9; q++) acc[q] += v[u][q];
                    } else {                        // lower block = transpose of the stored upper one
#pragma unroll
                        for (int a = 0; a < 3; a++)
#pragma unroll
                            for (int b = 0; b < 3; b++) acc[3 * a + b] += v[u][3 * b + a];
                    }
                    while (rest) {                  // degenerate element: the node appears again at column j
                        j = __ffs(rest) - 1;
                        rest &= rest - 1;
                        const double *src = blk[u] + (i <= j ? upper_index(i, j) : upper_index(j, i)) * KE_BLK;
                        for (int a = 0; a < 3; a++)
                            for (int b = 0; b < 3; b++) acc[3 * a + b] += (i <= j) ? src[3 * a + b] : src[3 * b + a];
                    }
                }
            }
        }
        if (have) {
            const uint8_t *fq = fixed + 3 * (int64_t)myq;
            double *out = vals + 9 * (int64_t)s0 + 3 * s;
#pragma unroll
            for (int a = 0; a < 3; a++)
#pragma unroll
                for (int b = 0; b < 3; b++) {
                    double v = acc[3 * a + b];
                    if (fr[a] || fq[b]) v = (fr[a] && myq == (int32_t)p && a == b) ? 1.0 : 0.0;   // SPC rows/columns
                    out[(int64_t)a * 3 * nb + b] = v;
                    if (myq == (int32_t)p && a == b) {                                              // Jacobi scaling
                        const double d = v > 0.0 ? 1.0 / sqrt(v) : 1.0;
                        d2[3 * rl + a] = d * d;
                    }
                }
        }
    }
}

__global__ void k_flag_local_elems(int64_t n_inc, const int32_t *__restrict__ inc, int32_t *__restrict__ flag) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t < n_inc) flag[inc[t] >> 3] = 1;
}

__global__ void k_compact_local_elems(int64_t n_elem, const int32_t *__restrict__ flag, const int32_t *__restrict__ pos,
                                      int32_t *__restrict__ g2l, int32_t *__restrict__ lelem) {
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= n_elem) return;
    if (flag[e]) { g2l[e] = pos[e]; lelem[pos[e]] = (int32_t)e; }
    else g2l[e] = -1;
}

// Element.K_Initial for a range of elements, 24x24 row-major each: one thread per matrix row block.
__global__ void __launch_bounds__(128)
k_element_ke(int64_t first, int64_t count, const int32_t *__restrict__ conn, const double *__restrict__ xyz,
             const uint8_t *__restrict__ etype, const int32_t *__restrict__ emat, const double *__restrict__ lam_tab,
             const double *__restrict__ G_tab, double *__restrict__ ke, int32_t *err) {
    __shared__ double s_tab[9 * 24];
    for (int t = threadIdx.x; t < 216; t += blockDim.x) s_tab[t] = (&c_dNl[0][0])[t];
    __syncthreads();
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= 8 * count) return;
    const int64_t e = first + (t >> 3);
    const int i = (int)(t & 7);
    double X[24], K[72];
    load_element(conn, xyz, e, X);
    const int mat = emat[e];
    if (hex8_row_block(etype[e], X, i, lam_tab[mat], G_tab[mat], s_tab, K)) atomicOr(err + 2, 1);
    double *out = ke + (t >> 3) * 576 + (int64_t)i * 72;
#pragma unroll
    for (int k = 0; k < 72; k++) out[k] = K[k];
}

// Table entries use the expression form of FE_Library.cs:243-273 so they are bit-identical to the
// reference's: 1/8 * (c0 + c1*u + c2*v + c3*u*v).
void host_diff_shape(double xi, double eta, double zeta, double *dN) {
    static const double S[8][3] = {{-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1},
                                   {-1, -1, 1},  {1, -1, 1},  {1, 1, 1},  {-1, 1, 1}};
    for (int k = 0; k < 8; k++) {
        const double sx = S[k][0], sy = S[k][1], sz = S[k][2];
        dN[k] = 1.0 / 8.0 * (sx + (sx * sy) * eta + (sx * sz) * zeta + (sx * sy * sz) * (eta * zeta));
        dN[8 + k] = 1.0 / 8.0 * (sy + (sy * sx) * xi + (sy * sz) * zeta + (sy * sx * sz) * (xi * zeta));
        dN[16 + k] = 1.0 / 8.0 * (sz + (sz * sx) * xi + (sz * sy) * eta + (sz * sx * sy) * (xi * eta));
    }
}

}  // namespace

void host_fe_tables(double *tab /*9*24*/) {
    static const double S[8][3] = {{-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1},
                                   {-1, -1, 1},  {1, -1, 1},  {1, 1, 1},  {-1, 1, 1}};
    const double g = sqrt(1.0 / 3.0);                       // FE_Library.cs:103
    for (int q = 0; q < 8; q++) host_diff_shape(S[q][0] * g, S[q][1] * g, S[q][2] * g, tab + 24 * q);
    host_diff_shape(0.0, 0.0, 0.0, tab + 24 * 8);
}

int upload_fe_tables() {
    double tab[9 * 24];
    host_fe_tables(tab);
    STAN_CUDA(cudaMemcpyToSymbol(c_dNl, tab, sizeof tab));
    // constant memory is per device: stan_create calls this for every handle's device
    unsigned char bi[36], bj[36];
    int n = 0;
    for (int i = 0; i < 8; i++) for (int j = i; j < 8; j++) { bi[n] = (unsigned char)i; bj[n] = (unsigned char)j; n++; }
    STAN_CUDA(cudaMemcpyToSymbol(c_blk_i, bi, sizeof bi));
    STAN_CUDA(cudaMemcpyToSymbol(c_blk_j, bj, sizeof bj));
    return STAN_OK;
}

// Returns 0 when the matrix was assembled, 1 when the caller should use the fused kernel instead
// (Ke store does not fit), < 0 on error.
static int run_assembly_two_kernel(stan_handle *h) {
    cudaStream_t s = h->stream;
    const int64_t nloc = h->row1 - h->row0;
    // elements that touch an owned row: all of them on one GPU, a compacted list otherwise
    int64_t n_local = h->n_elem;
    ScratchBuf<int32_t> g2l(&h->scratch[6]), lelem(&h->scratch[7]);
    if (h->world > 1) {
        int32_t n_inc = 0;
        STAN_CUDA(cudaMemcpyAsync(&n_inc, h->d_inc_ptr.p + nloc, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        ScratchBuf<int32_t> flag(&h->scratch[0]), pos(&h->scratch[1]);
        STAN_TRY(flag.alloc(h->n_elem + 1, s)); STAN_TRY(pos.alloc(h->n_elem + 1, s)); STAN_TRY(g2l.alloc(h->n_elem, s));
        STAN_CUDA(cudaMemsetAsync(flag.p, 0, (h->n_elem + 1) * sizeof(int32_t), s));
        STAN_CUDA(cudaStreamSynchronize(s));
        k_flag_local_elems<<<div_up(n_inc, 256), 256, 0, s>>>(n_inc, h->d_inc.p, flag.p);
        STAN_TRY(device_exclusive_scan_i32(h, flag.p, pos.p, h->n_elem + 1, s));
        int32_t cnt = 0;
        STAN_CUDA(cudaMemcpyAsync(&cnt, pos.p + h->n_elem, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        STAN_CUDA(cudaStreamSynchronize(s));
        n_local = cnt;
        STAN_TRY(lelem.alloc(n_local, s));
        k_compact_local_elems<<<div_up(h->n_elem, 256), 256, 0, s>>>(h->n_elem, flag.p, pos.p, g2l.p, lelem.p);
        flag.release(s); pos.release(s);
        h->launches += 3;
    }
    size_t free_b = 0, total_b = 0;
    const size_t need = (size_t)n_local * KE_ELEM * sizeof(double);
    cudaMemGetInfo(&free_b, &total_b);
    if (h->d_ke.n * sizeof(double) < need && need + ((size_t)2 << 30) > free_b) {   // would not fit next to the matrix
        g2l.release(s); lelem.release(s);
        return 1;
    }
    STAN_TRY(h->d_ke.alloc((size_t)n_local * KE_ELEM, s));
    k_hex8_ke_batch<<<div_up(n_local, KB_ELEMS), KB_THREADS, 0, s>>>(n_local, h->world > 1 ? lelem.p : nullptr, h->d_conn.p,
                                                                    h->d_xyz.p, h->d_etype.p, h->d_emat.p, h->d_lambda.p,
                                                                    h->d_G.p, h->d_ke.p, h->d_err.p);
    k_assemble_gather<<<div_up(nloc, GA_WARPS), 32 * GA_WARPS, 0, s>>>(nloc, h->row0, h->d_inc_ptr.p, h->d_inc.p,
                                                                       h->d_brow_ptr.p, h->d_bcol.p, h->d_conn.p,
                                                                       h->d_node_index.p, h->world > 1 ? g2l.p : nullptr,
                                                                       h->d_ke.p, h->d_fixed.p, h->d_vals.p, h->d_d2.p);
    STAN_CUDA(cudaGetLastError());
    g2l.release(s); lelem.release(s);
    h->launches += 2;
    return 0;
}

int run_assembly(stan_handle *h) {
    cudaStream_t s = h->stream;
    const int64_t nloc = h->row1 - h->row0;
    STAN_TRY(h->d_vals.alloc((size_t)9 * h->n_blocks + 2, s));   // +2: 16-byte granules of the bulk-copy SpMV
    STAN_TRY(h->d_d2.alloc(3 * nloc, s));
    {   // two-kernel path unless the Ke store does not fit in free memory or STAN_ASM=1 asks for the fused kernel
        static const int want = getenv("STAN_ASM") ? atoi(getenv("STAN_ASM")) : 2;
        int rc = want == 2 ? run_assembly_two_kernel(h) : 1;
        if (rc <= 0) return rc;                    // 0 = done, < 0 = error, 1 = fall through to the fused kernel
    }
    const size_t smem = (size_t)9 * h->max_group_blocks * sizeof(double);
    if (smem > 200 * 1024) {
        set_error("32 consecutive rows couple to %d blocks; the assembly tile holds at most %d", h->max_group_blocks,
                  (int)(200 * 1024 / 72));
        return STAN_E_NOMEM;
    }
    STAN_CUDA(cudaFuncSetAttribute(k_assemble_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_assemble_rows<<<div_up(nloc, ROWS_PER_CTA), ASM_THREADS, smem, s>>>(
        nloc, h->row0, h->d_inc_ptr.p, h->d_inc.p, h->d_brow_ptr.p, h->d_bcol.p, h->d_conn.p, h->d_xyz.p,
        h->d_node_index.p, h->d_etype.p, h->d_emat.p, h->d_lambda.p, h->d_G.p, h->d_fixed.p, h->d_vals.p, h->d_d2.p,
        h->d_err.p);
    STAN_CUDA(cudaGetLastError());
    h->launches += 1;
    return STAN_OK;
}

int element_stiffness(stan_handle *h, int64_t first, int64_t count, double *ke_host) {
    cudaStream_t s = h->stream;
    STAN_TRY(h->d_err.alloc(8, s));
    STAN_CUDA(cudaMemsetAsync(h->d_err.p, 0, 8 * sizeof(int32_t), s));
    DevBuf<double> ke;
    STAN_TRY(ke.alloc((size_t)count * 576, s));
    k_element_ke<<<div_up(8 * count, 128), 128, 0, s>>>(first, count, h->d_conn.p, h->d_xyz.p, h->d_etype.p,
                                                        h->d_emat.p, h->d_lambda.p, h->d_G.p, ke.p, h->d_err.p);
    STAN_CUDA(cudaGetLastError());
    int32_t herr[4];
    STAN_CUDA(cudaMemcpyAsync(ke_host, ke.p, (size_t)count * 576 * sizeof(double), cudaMemcpyDeviceToHost, s));
    STAN_CUDA(cudaMemcpyAsync(herr, h->d_err.p, sizeof herr, cudaMemcpyDeviceToHost, s));
    STAN_CUDA(cudaStreamSynchronize(s));
    ke.release(s);
    h->launches += 1;
    if (herr[2]) { set_error("singular Jacobian (det == 0) in element range"); return STAN_E_SINGULAR; }
    return STAN_OK;
}

}  // namespace stan
