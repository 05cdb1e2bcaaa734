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

// ---- assembly, version 2: integrate, stage, gather by slot -------------------------------------
// ncu on k_assemble_rows (profiles/r01_path_10m_ncu.md): 14 of 32 lanes active on average and the
// eight barrier-separated accumulation rounds cost as much as the integration.  Here thread
// (row r, rank k) = t integrates the k-th incident element of row r exactly as before, parks its
// 3x24 row block in shared memory (column-major, conflict-free) and records, per (row, block slot),
// which of its eight element columns land there (one mask byte per rank).  Then every thread owns
// output entries and *gathers*: out(r, a, slot, b) = sum over ranks k ascending, columns j ascending
// — the same fixed element order, no atomics, no serialised rounds, all lanes busy.  16 rows and
// 128 threads per CTA keep two CTAs resident per SM so one integrates while the other gathers.
constexpr int A2_ROWS = 16;
constexpr int A2_THREADS = 128;
constexpr int A2_STRIDE = A2_THREADS + 1;      // padded stride of the staged row blocks (doubles)

__global__ void __launch_bounds__(A2_THREADS, 2)
k_assemble_rows2(int64_t nloc, int64_t row0, const int32_t *__restrict__ inc_ptr, const int32_t *__restrict__ inc,
                 const int32_t *__restrict__ brow_ptr, const int32_t *__restrict__ bcol,
                 const int32_t *__restrict__ conn, const double *__restrict__ xyz,
                 const int32_t *__restrict__ node_index, const uint8_t *__restrict__ etype,
                 const int32_t *__restrict__ emat, const double *__restrict__ lam_tab, const double *__restrict__ G_tab,
                 const uint8_t *__restrict__ fixed, double *__restrict__ vals, double *__restrict__ d2, int32_t *err,
                 int max_blocks) {
    extern __shared__ __align__(16) unsigned char s_dyn[];
    double *s_stage = reinterpret_cast<double *>(s_dyn);                       // [72][A2_STRIDE]
    double *s_out = s_stage + 72 * A2_STRIDE;                                   // [9 * max_blocks]
    unsigned long long *s_mask = reinterpret_cast<unsigned long long *>(s_out + 9 * max_blocks);   // [max_blocks]: 8 rank bytes
    __shared__ double s_tab[9 * 24];
    __shared__ int s_inc[A2_ROWS + 1], s_brow[A2_ROWS + 1];
    const int tid = threadIdx.x;
    const int64_t r0 = (int64_t)blockIdx.x * A2_ROWS;
    const int nr = (int)((nloc - r0) < A2_ROWS ? (nloc - r0) : A2_ROWS);
    if (tid <= nr) { s_inc[tid] = inc_ptr[r0 + tid]; s_brow[tid] = brow_ptr[r0 + tid]; }
    for (int t = tid; t < 216; t += A2_THREADS) s_tab[t] = (&c_dNl[0][0])[t];
    __syncthreads();
    const int b0 = s_brow[0], nblk = s_brow[nr] - b0, nvals = 9 * nblk;
    for (int t = tid; t < nvals; t += A2_THREADS) s_out[t] = 0.0;

    const int r = tid >> 3, k = tid & 7;                    // row of the tile, rank slot
    int maxcnt = 0;
    for (int i = 0; i < nr; i++) maxcnt = max(maxcnt, s_inc[i + 1] - s_inc[i]);
    for (int c0 = 0; c0 < maxcnt; c0 += 8) {                // rows with more than 8 incident elements: several passes
        for (int t = tid; t < nblk; t += A2_THREADS) s_mask[t] = 0ull;
        __syncthreads();
        const int idx = (r < nr) ? s_inc[r] + c0 + k : 0;
        const bool active = r < nr && idx < s_inc[r + 1];
        if (active) {
            const int ent = inc[idx];
            const int64_t e = ent >> 3;
            double X[24], K[72];
            load_element(conn, xyz, e, X);
            const int mat = emat[e];
            if (hex8_row_block(etype[e], X, ent & 7, lam_tab[mat], G_tab[mat], s_tab, K)) atomicOr(err + 2, 1);
#pragma unroll
            for (int q = 0; q < 72; q++) s_stage[q * A2_STRIDE + tid] = K[q];
            // which element column j lands in which block slot of row r
            const int nb = s_brow[r + 1] - s_brow[r];
            const int32_t *cols = bcol + s_brow[r];
            unsigned char *mrow = reinterpret_cast<unsigned char *>(s_mask + (s_brow[r] - b0));
            const int4 c0v = *reinterpret_cast<const int4 *>(conn + 8 * e);
            const int4 c1v = *reinterpret_cast<const int4 *>(conn + 8 * e + 4);
            const int nd[8] = {c0v.x, c0v.y, c0v.z, c0v.w, c1v.x, c1v.y, c1v.z, c1v.w};
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int32_t q = node_index[nd[j]];
                int lo = 0, hi = nb - 1;
                while (lo < hi) {
                    int mid = (lo + hi) >> 1;
                    if (cols[mid] < q) lo = mid + 1; else hi = mid;
                }
                mrow[8 * lo + k] |= (unsigned char)(1u << j);   // byte (slot, rank) belongs to this thread only
            }
        }
        __syncthreads();
        // gather: entry w of the tile's value storage = (row rr, scalar row a, slot s, column b)
        for (int w = tid; w < nvals; w += A2_THREADS) {
            int rr = 0;
            while (rr + 1 < nr && 9 * (s_brow[rr + 1] - b0) <= w) rr++;
            const int nbr = s_brow[rr + 1] - s_brow[rr];
            const int local = w - 9 * (s_brow[rr] - b0);
            const int a = local / (3 * nbr), rem = local - a * 3 * nbr;
            const int s = rem / 3, b = rem - 3 * s;
            unsigned long long m = s_mask[s_brow[rr] - b0 + s];
            double sum = 0.0;
            for (int kk = 0; kk < 8 && m; kk++, m >>= 8) {
                unsigned int bits = (unsigned int)(m & 0xffull);
                while (bits) {
                    const int j = __ffs(bits) - 1;
                    bits &= bits - 1;
                    sum += s_stage[(a * 24 + 3 * j + b) * A2_STRIDE + rr * 8 + kk];
                }
            }
            s_out[w] += sum;
        }
        __syncthreads();
    }
    // SPC masking (fixed rows/columns dropped, SolverFunctions.cs:158-160), identity on fixed rows, Jacobi scaling
    for (int w = tid; w < nvals; w += A2_THREADS) {
        int rr = 0;
        while (rr + 1 < nr && 9 * (s_brow[rr + 1] - b0) <= w) rr++;
        const int nbr = s_brow[rr + 1] - s_brow[rr];
        const int local = w - 9 * (s_brow[rr] - b0);
        const int a = local / (3 * nbr), rem = local - a * 3 * nbr;
        const int s = rem / 3, b = rem - 3 * s;
        const int64_t p = row0 + r0 + rr;
        const int64_t q = bcol[s_brow[rr] + s];
        const bool fr = fixed[3 * p + a] != 0, fc = fixed[3 * q + b] != 0;
        if (fr || fc) s_out[w] = (fr && q == p && a == b) ? 1.0 : 0.0;
    }
    __syncthreads();
    if (tid < 3 * nr) {
        const int rl = tid / 3, a = tid % 3;
        const int64_t p = row0 + r0 + rl;
        const int nb = s_brow[rl + 1] - s_brow[rl];
        const int32_t *cols = bcol + s_brow[rl];
        int lo = 0, hi = nb - 1;
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (cols[mid] < (int32_t)p) lo = mid + 1; else hi = mid;
        }
        const double v = s_out[9 * (s_brow[rl] - b0) + a * 3 * nb + 3 * lo + a];
        const double d = v > 0.0 ? 1.0 / sqrt(v) : 1.0;
        d2[3 * (r0 + rl) + a] = d * d;
    }
    double *out = vals + 9 * (int64_t)b0;
    for (int t = tid; t < nvals; t += A2_THREADS) out[t] = s_out[t];
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
    return STAN_OK;
}

int run_assembly(stan_handle *h) {
    cudaStream_t s = h->stream;
    const int64_t nloc = h->row1 - h->row0;
    STAN_TRY(h->d_vals.alloc((size_t)9 * h->n_blocks + 2, s));   // +2: 16-byte granules of the bulk-copy SpMV
    STAN_TRY(h->d_d2.alloc(3 * nloc, s));
    {   // version 2 (stage + gather) is opt-in (STAN_ASM=2): measured 180 ms vs 114 ms for version 1 on the 10M
        // beam — decoding (row, scalar row, slot, column) per output entry costs more than the rounds it removes
        static const int want = getenv("STAN_ASM") ? atoi(getenv("STAN_ASM")) : 1;
        const int mb = h->max_group16 > 0 ? h->max_group16 : 1;
        const size_t smem2 = (size_t)72 * A2_STRIDE * sizeof(double) + (size_t)mb * (72 + 8);
        if (want == 2 && smem2 <= 220 * 1024) {
            STAN_CUDA(cudaFuncSetAttribute(k_assemble_rows2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
            k_assemble_rows2<<<div_up(nloc, A2_ROWS), A2_THREADS, smem2, s>>>(
                nloc, h->row0, h->d_inc_ptr.p, h->d_inc.p, h->d_brow_ptr.p, h->d_bcol.p, h->d_conn.p, h->d_xyz.p,
                h->d_node_index.p, h->d_etype.p, h->d_emat.p, h->d_lambda.p, h->d_G.p, h->d_fixed.p, h->d_vals.p,
                h->d_d2.p, h->d_err.p, mb);
            STAN_CUDA(cudaGetLastError());
            h->launches += 1;
            return STAN_OK;
        }
    }
    const size_t smem = (size_t)9 * h->max_group_blocks * sizeof(double);
    if (smem > 200 * 1024) {
        set_error("32 consecutive rows couple to %d blocks; the assembly tile holds at most %d", h->max_group_blocks,
                  (int)(200 * 1024 / 72));
        return STAN_E_CAPACITY;
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
