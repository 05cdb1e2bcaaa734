// hex8 stiffness integration and deterministic row-owner assembly.
//
// Replaces Element.K_Initial (/root/reference/src/STAN_Database/Element.cs:118-155, with
// Jacobian :274-292, BL0_Matrix :297-328 and the MatrixST products MatrixST.cs:404-427) and the
// locked scatter loop of ParallelAssembly_K (/root/reference/src/STAN_Solver/SolverFunctions.cs:129-174).
//
// Arithmetic.  The reference forms K += ((BL^T D) BL) * (det J * w) with naive triple
// loops over matrices that are mostly structural zeros.  A product with an exact zero leaves a running
// sum unchanged, so every entry of Ke is determined by its few non-zero terms taken in the reference's
// order (k ascending).  For the isotropic D of Material.cs:39-53 and the BL0 of Element.cs:316-324, with
// d = grad N_i and e = grad N_j at a Gauss point:
//     (BL^T D)[ia][k] is a single product: d_a * (k == a ? lambda + 2G : lambda) for k < 3, d_. * G for the
//     shear rows; and E[ia][jb] = sum_k (BL^T D)[ia][k] BL[k][jb] has three terms when a == b, two otherwise
//     (the expressions are written out in block_terms()).
// Every operation is an explicit __dmul_rn / __dadd_rn (no FMA contraction: the C# JIT and the CPU restatement
// round after each multiply), so Ke is BIT-IDENTICAL to the dense triple loops — tests/ compares with
// array_equal — at about a quarter of their operation count.
//
// Only the upper triangle of K exists in the reference (col >= row, SolverFunctions.cs:155).  For a node
// pair the stored 3x3 block is the one whose ROW node has the smaller DOF index; the full-storage matrix
// used by the SpMV takes the transpose for the mirrored block, exactly as the symmetric product of an
// upper-triangle matrix does.  Diagonal blocks use their upper triangle, mirrored.
//
// Work decomposition: two kernels, deterministic, no atomics.
//   k_hex8_ke      one CTA per batch of 8 elements.  Phase 1: thread (element, Gauss point) forms J, det J,
//                  J^-1 and the global shape-function derivatives once and stages them in shared memory.
//                  Phase 2: thread (element, node pair i <= j) contracts them over the Gauss points into one
//                  3x3 block (9 accumulators) and stores it in an 80-byte slot of the Ke store, fully
//                  coalesced; the DOF indices of the element's 8 nodes go to a 32-byte record next to it.
//   k_assemble_gather  one warp per matrix row, lanes = block slots.  It walks the row's incident (element,
//                  local node) entries in ascending order — the order a serial ParallelAssembly_K would add
//                  them in — matches the element's 8 column indices against the lane's column, and adds the
//                  block (or its transpose).  Every stored value is written once.
// Rows are processed in chunks sized so that the Ke store of a chunk fits in free device memory (one chunk
// at every BASELINE size on one B200); elements on a chunk boundary are integrated by both chunks.
#include <algorithm>

#include "common.cuh"

namespace stan {

namespace {

// dN_dLocal tables: entries 0..7 = the 2x2x2 points of HEX8_G2 (FE_Library.cs:119-129),
// entry 8 = the centre point of HEX8_G1 (FE_Library.cs:83-87); [point][3][8].
__constant__ double c_dNl[9][24];
__constant__ unsigned char c_blk_i[36], c_blk_j[36];

constexpr int KB_ELEMS = 8;                       // elements per CTA
constexpr int KE_BLK = 10;                        // doubles per stored 3x3 block: 9 values + 1 zero pad = 80 B, 16-byte aligned
constexpr int KE_ELEM = 36 * KE_BLK;              // doubles per element in the Ke store
constexpr int KB_THREADS = KB_ELEMS * 36;         // one thread per node pair in phase 2

__device__ __forceinline__ int upper_index(int i, int j) { return i * 8 - (i * (i - 1)) / 2 + (j - i); }   // i <= j

// IEEE operations that the compiler may not contract into FMAs
__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }

__device__ __forceinline__ double det3(const double *m) {
    // MatrixST.Det3, terms and association as written at MatrixST.cs:274-279
    double r = mul(mul(m[0], m[4]), m[8]);
    r = add(r, mul(mul(m[3], m[7]), m[2]));
    r = add(r, mul(mul(m[6], m[1]), m[5]));
    r = sub(r, mul(mul(m[2], m[4]), m[6]));
    r = sub(r, mul(mul(m[0], m[5]), m[7]));
    r = sub(r, mul(mul(m[8], m[1]), m[3]));
    return r;
}

// K[3a + b] += E[ia][jb] * s for the block (row node i with gradient d, column node j with gradient e).
// Strain rows {xx, yy, zz, xy, yz, xz}: column (i, x) of BL has dx in row 0, dy in row 3, dz in row 5;
// (i, y): dy in 1, dx in 3, dz in 4; (i, z): dz in 2, dy in 4, dx in 5 (Element.cs:316-324).
__device__ __forceinline__ void block_terms(const double *d, const double *e, double lam, double G, double D0, double s,
                                            double (&K)[9]) {
    const double dx = d[0], dy = d[1], dz = d[2], ex = e[0], ey = e[1], ez = e[2];
    const double dxD = mul(dx, D0), dyD = mul(dy, D0), dzD = mul(dz, D0);
    const double dxL = mul(dx, lam), dyL = mul(dy, lam), dzL = mul(dz, lam);
    const double dxG = mul(dx, G), dyG = mul(dy, G), dzG = mul(dz, G);
    double E[9];
    E[0] = add(add(mul(dxD, ex), mul(dyG, ey)), mul(dzG, ez));      // k = 0, 3, 5
    E[1] = add(mul(dxL, ey), mul(dyG, ex));                         // k = 1, 3
    E[2] = add(mul(dxL, ez), mul(dzG, ex));                         // k = 2, 5
    E[3] = add(mul(dyL, ex), mul(dxG, ey));                         // k = 0, 3
    E[4] = add(add(mul(dyD, ey), mul(dxG, ex)), mul(dzG, ez));      // k = 1, 3, 4
    E[5] = add(mul(dyL, ez), mul(dzG, ey));                         // k = 2, 4
    E[6] = add(mul(dzL, ex), mul(dxG, ez));                         // k = 0, 5
    E[7] = add(mul(dzL, ey), mul(dyG, ez));                         // k = 1, 4
    E[8] = add(add(mul(dzD, ez), mul(dyG, ey)), mul(dxG, ex));      // k = 2, 4, 5
#pragma unroll
    for (int q = 0; q < 9; q++) K[q] = add(K[q], mul(E[q], s));     // MultiplyScalar, then K += (Element.cs:151)
}

// One Gauss point of one element: out[3k + c] = dN_k/dx_c (k = node), out[24] = det J * w.
// J = dN_dLocal * X with k-ascending sums (MatrixST operator*), J^-1 = adjugate * (1/det) (MatrixST.cs:294-319),
// dN = J^-1 * dN_dLocal.  Returns false when det J == 0 (the reference throws there).
// tab: the 3 x 8 dN_dLocal table of the Gauss point (constant memory, or a shared-memory copy when the lanes of a
// warp work on different Gauss points and the constant cache would serialise them)
__device__ __forceinline__ bool gauss_point_geometry(const double (&X)[24], const double *tab, double w, double *out) {
    double J[9];
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) {
            double s = mul(tab[r * 8], X[c]);
#pragma unroll
            for (int k = 1; k < 8; k++) s = add(s, mul(tab[r * 8 + k], X[k * 3 + c]));
            J[r * 3 + c] = s;
        }
    const double det = det3(J);
    const double inv = __ddiv_rn(1.0, det);
    double Ji[9];
    Ji[0] = mul(inv, sub(mul(J[4], J[8]), mul(J[5], J[7])));
    Ji[1] = mul(inv, sub(mul(J[2], J[7]), mul(J[1], J[8])));
    Ji[2] = mul(inv, sub(mul(J[1], J[5]), mul(J[2], J[4])));
    Ji[3] = mul(inv, sub(mul(J[5], J[6]), mul(J[3], J[8])));
    Ji[4] = mul(inv, sub(mul(J[0], J[8]), mul(J[2], J[6])));
    Ji[5] = mul(inv, sub(mul(J[2], J[3]), mul(J[0], J[5])));
    Ji[6] = mul(inv, sub(mul(J[3], J[7]), mul(J[4], J[6])));
    Ji[7] = mul(inv, sub(mul(J[1], J[6]), mul(J[0], J[7])));
    Ji[8] = mul(inv, sub(mul(J[0], J[4]), mul(J[1], J[3])));
#pragma unroll
    for (int k = 0; k < 8; k++)
#pragma unroll
        for (int c = 0; c < 3; c++)
            out[3 * k + c] = add(add(mul(Ji[c * 3], tab[k]), mul(Ji[c * 3 + 1], tab[8 + k])), mul(Ji[c * 3 + 2], tab[16 + k]));
    out[24] = mul(det, w);
    return det != 0.0;
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

// lelem: elements to integrate (nullptr: first + le); node_index: DOF map (nullptr: local node order decides
// the orientation of a pair block — stan_element_stiffness before a DOF map exists).
__global__ void __launch_bounds__(KB_THREADS, 3)
k_hex8_ke(int64_t n_local, int64_t first, const int32_t *__restrict__ lelem, const int32_t *__restrict__ conn,
          const double *__restrict__ xyz, const int32_t *__restrict__ node_index, const uint8_t *__restrict__ etype,
          const int32_t *__restrict__ emat, const double *__restrict__ lam_tab, const double *__restrict__ G_tab,
          double *__restrict__ ke_store, int32_t *__restrict__ qrec, int32_t *err) {
    __shared__ double s_dn[KB_ELEMS][8][25];       // per (element, Gauss point): dN[node][xyz] and w|J|
    __shared__ double s_lam[KB_ELEMS], s_G[KB_ELEMS];
    __shared__ int s_ng[KB_ELEMS];
    __shared__ int s_q[KB_ELEMS][8];
    const int tid = threadIdx.x;
    const int64_t le0 = (int64_t)blockIdx.x * KB_ELEMS;
    if (tid < KB_ELEMS * 8) {                      // ---- phase 1: Jacobians and derivatives, once per (e, g)
        const int el = tid >> 3, g = tid & 7;
        const int64_t le = le0 + el;
        if (le < n_local) {
            const int64_t e = lelem ? lelem[le] : first + le;
            const int type = etype[e];
            const int ng = (type == STAN_HEX8_G2) ? 8 : 1;
            if (g == 0) { s_ng[el] = ng; const int mat = emat[e]; s_lam[el] = lam_tab[mat]; s_G[el] = G_tab[mat]; }
            const int q = node_index ? node_index[conn[8 * e + g]] : g;      // thread g also fetches node g's DOF index
            s_q[el][g] = q;
            qrec[8 * le + g] = q;
            if (g < ng) {
                double X[24];
                load_element(conn, xyz, e, X);
                // GaussWeight: 1 at the 2x2x2 points, 8 at the centre point (FE_Library.cs:72,100)
                if (!gauss_point_geometry(X, c_dNl[type == STAN_HEX8_G2 ? g : 8], type == STAN_HEX8_G2 ? 1.0 : 8.0, s_dn[el][g]))
                    atomicOr(err + 2, 1);
            }
        }
    }
    __syncthreads();
    // ---- phase 2: one node pair per thread, contracted over the Gauss points ----
    const int el = tid / 36, blk = tid - 36 * el;
    const int64_t le = le0 + el;
    if (le >= n_local) return;
    int i = c_blk_i[blk], j = c_blk_j[blk];
    if (s_q[el][j] < s_q[el][i]) { const int t = i; i = j; j = t; }      // row node = the one with the smaller DOF index
    const double lam = s_lam[el], G = s_G[el], D0 = add(lam, mul(2.0, G));   // Material.cs:42: lambda + (2 * G)
    double K[9];
#pragma unroll
    for (int q = 0; q < 9; q++) K[q] = 0.0;
    const int ng = s_ng[el];
    for (int g = 0; g < ng; g++) {
        const double *d = s_dn[el][g];
        block_terms(d + 3 * i, d + 3 * j, lam, G, D0, d[24], K);
    }
    // 80-byte slots: 16-byte aligned vector stores / loads; the pad is written too, so every sector is written
    // whole and L2 never has to fetch a line to complete it
    double *out = ke_store + (le * 36 + blk) * KE_BLK;
#pragma unroll
    for (int q = 0; q < 8; q += 2) *reinterpret_cast<double2 *>(out + q) = make_double2(K[q], K[q + 1]);
    *reinterpret_cast<double2 *>(out + 8) = make_double2(K[8], 0.0);
}

// ---- warp-specialised variant --------------------------------------------------------------------------------
// ncu on k_hex8_ke: the top stall is the barrier between the phases (9.4 warps per issue) — 224 of 288 threads
// wait while 64 form the geometry, then those 64 mostly wait.  Here the two phases run concurrently on different
// warps of a persistent CTA: two producer groups of 64 threads each prepare the geometry of later batches (group
// it % 2 fills buffer it % 3) while the 288 consumer threads contract the current one; named barriers
// FULL / EMPTY per buffer replace __syncthreads.  Same arithmetic, same bits.
constexpr int WS_CONS = KB_THREADS;               // 288 consumer threads = warps 0..8
constexpr int WS_PROD = 64;                       // per producer group: warps 9,10 and 11,12
constexpr int WS_THREADS = WS_CONS + 2 * WS_PROD;
constexpr int WS_BUFS = 3;

__device__ __forceinline__ void named_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

__global__ void __launch_bounds__(WS_THREADS, 2)
k_hex8_ke_ws(int64_t n_local, int64_t first, const int32_t *__restrict__ lelem, const int32_t *__restrict__ conn,
             const double *__restrict__ xyz, const int32_t *__restrict__ node_index, const uint8_t *__restrict__ etype,
             const int32_t *__restrict__ emat, const double *__restrict__ lam_tab, const double *__restrict__ G_tab,
             double *__restrict__ ke_store, int32_t *__restrict__ qrec, int32_t *err) {
    __shared__ double s_dn[WS_BUFS][KB_ELEMS][8][25];
    __shared__ double s_lam[WS_BUFS][KB_ELEMS], s_G[WS_BUFS][KB_ELEMS];
    __shared__ int s_ng[WS_BUFS][KB_ELEMS];
    __shared__ int s_q[WS_BUFS][KB_ELEMS][8];
    __shared__ double s_tab[9 * 25];               // dN_dLocal in 25-double rows: the 8 Gauss points of a producer warp on distinct banks
    const int tid = threadIdx.x;
    const int64_t n_batches = (n_local + KB_ELEMS - 1) / KB_ELEMS;
    constexpr int NB = WS_CONS + WS_PROD;          // participants of every named barrier
    if (tid < 216) s_tab[25 * (tid / 24) + tid % 24] = (&c_dNl[0][0])[tid];
    __syncthreads();
    if (tid >= WS_CONS) {
        // ---- producers: group g2 prepares iterations g2, g2 + 2, ...
        const int g2 = (tid - WS_CONS) / WS_PROD, t = (tid - WS_CONS) % WS_PROD;
        const int el = t >> 3, g = t & 7;
        for (int64_t it = g2, b = blockIdx.x + (int64_t)g2 * gridDim.x; b < n_batches; it += 2, b += 2 * (int64_t)gridDim.x) {
            const int buf = (int)(it % WS_BUFS);
            if (it >= WS_BUFS) named_sync(4 + buf, NB);                // the consumers have released this buffer
            const int64_t le = b * KB_ELEMS + el;
            if (le < n_local) {
                const int64_t e = lelem ? lelem[le] : first + le;
                const int type = etype[e];
                const int ng = (type == STAN_HEX8_G2) ? 8 : 1;
                if (g == 0) { s_ng[buf][el] = ng; const int mat = emat[e]; s_lam[buf][el] = lam_tab[mat]; s_G[buf][el] = G_tab[mat]; }
                const int q = node_index ? node_index[conn[8 * e + g]] : g;
                s_q[buf][el][g] = q;
                qrec[8 * le + g] = q;
                if (g < ng) {
                    double X[24];
                    load_element(conn, xyz, e, X);
                    if (!gauss_point_geometry(X, s_tab + 25 * (type == STAN_HEX8_G2 ? g : 8), type == STAN_HEX8_G2 ? 1.0 : 8.0, s_dn[buf][el][g]))
                        atomicOr(err + 2, 1);
                }
            } else if (g == 0) s_ng[buf][el] = 0;
            named_arrive(1 + buf, NB);                                 // buffer full
        }
        return;
    }
    // ---- consumers
    const int el = tid / 36, blk = tid - 36 * el;
    const int bi = c_blk_i[blk], bj = c_blk_j[blk];
    for (int64_t it = 0, b = blockIdx.x; b < n_batches; it++, b += gridDim.x) {
        const int buf = (int)(it % WS_BUFS);
        named_sync(1 + buf, NB);
        const int ng = s_ng[buf][el];
        if (ng > 0) {
            int i = bi, j = bj;
            if (s_q[buf][el][j] < s_q[buf][el][i]) { const int t = i; i = j; j = t; }
            const double lam = s_lam[buf][el], G = s_G[buf][el], D0 = add(lam, mul(2.0, G));
            double K[9];
#pragma unroll
            for (int q = 0; q < 9; q++) K[q] = 0.0;
            for (int g = 0; g < ng; g++) {
                const double *d = s_dn[buf][el][g];
                block_terms(d + 3 * i, d + 3 * j, lam, G, D0, d[24], K);
            }
            double *out = ke_store + ((b * KB_ELEMS + el) * 36 + blk) * KE_BLK;
#pragma unroll
            for (int q = 0; q < 8; q += 2) *reinterpret_cast<double2 *>(out + q) = make_double2(K[q], K[q + 1]);
            *reinterpret_cast<double2 *>(out + 8) = make_double2(K[8], 0.0);
        }
        // release the buffer for the producer group that fills it three iterations from now (if there is one)
        if (b + (int64_t)WS_BUFS * gridDim.x < n_batches) named_arrive(4 + buf, NB);
    }
}

constexpr int GA_WARPS = 8;

// g2l: local index of an element in this chunk's Ke store (nullptr: the element id itself)
__global__ void __launch_bounds__(32 * GA_WARPS, 4)
k_assemble_gather(int64_t r_begin, int64_t r_end, int64_t row0, const int32_t *__restrict__ inc_ptr,
                  const int32_t *__restrict__ inc, const int32_t *__restrict__ brow_ptr, const int32_t *__restrict__ bcol,
                  const int32_t *__restrict__ qrec, const int32_t *__restrict__ g2l, const double *__restrict__ ke_store,
                  const uint8_t *__restrict__ fixed, double *__restrict__ vals, double *__restrict__ d2) {
    const int lane = threadIdx.x & 31;
    const int64_t rl = r_begin + (int64_t)blockIdx.x * GA_WARPS + (threadIdx.x >> 5);
    if (rl >= r_end) return;
    const int32_t p = (int32_t)(row0 + rl);
    const int s0 = brow_ptr[rl], nb = brow_ptr[rl + 1] - s0;
    const int t0 = inc_ptr[rl], t1 = inc_ptr[rl + 1];
    const bool fr[3] = {fixed[3 * (int64_t)p] != 0, fixed[3 * (int64_t)p + 1] != 0, fixed[3 * (int64_t)p + 2] != 0};
    for (int sb = 0; sb < nb; sb += 32) {          // one pass unless the row has more than 32 blocks
        const int s = sb + lane;
        const bool have = s < nb;
        const int32_t myq = have ? bcol[s0 + s] : -1;
        double acc[9];
#pragma unroll
        for (int q = 0; q < 9; q++) acc[q] = 0.0;
        // incidence entries ascending = (element, local node) ascending: the serial order of SolverFunctions.cs:143-172
        for (int t = t0; t < t1; t++) {
            const int ent = inc[t];                // warp-uniform
            const int64_t le = g2l ? g2l[ent >> 3] : (ent >> 3);
            const int i = ent & 7;
            const int4 qa = *reinterpret_cast<const int4 *>(qrec + 8 * le), qb = *reinterpret_cast<const int4 *>(qrec + 8 * le + 4);
            const int q8[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
            int bits = 0;                          // element columns that land in this lane's slot (> 1 bit: repeated node)
#pragma unroll
            for (int j = 0; j < 8; j++) bits |= (q8[j] == myq) << j;
            while (bits) {
                const int j = __ffs(bits) - 1;
                bits &= bits - 1;
                const double *src = ke_store + (le * 36 + (i <= j ? upper_index(i, j) : upper_index(j, i))) * KE_BLK;
                double v[9];
#pragma unroll
                for (int q = 0; q < 8; q += 2) {
                    const double2 t2 = *reinterpret_cast<const double2 *>(src + q);
                    v[q] = t2.x; v[q + 1] = t2.y;
                }
                v[8] = src[8];
                // The stored block has the node with the smaller DOF index as its row node.  A pair of local nodes
                // that are the same global node (collapsed hexahedron) is stored as K[min(i,j)][max(i,j)]; the other
                // orientation is taken as its transpose.
                const bool as_stored = (p < myq) || (p == myq && i <= j);
                if (as_stored) {
#pragma unroll
                    for (int q = 0; q < 9; q++) acc[q] = add(acc[q], v[q]);
                } else {
#pragma unroll
                    for (int a = 0; a < 3; a++)
#pragma unroll
                        for (int b = 0; b < 3; b++) acc[3 * a + b] = add(acc[3 * a + b], v[3 * b + a]);
                }
            }
        }
        if (have) {
            if (myq == p) { acc[3] = acc[1]; acc[6] = acc[2]; acc[7] = acc[5]; }   // diagonal block: upper triangle mirrored
            const uint8_t *fq = fixed + 3 * (int64_t)myq;
            double *out = vals + 9 * (int64_t)s0 + 3 * s;
#pragma unroll
            for (int a = 0; a < 3; a++)
#pragma unroll
                for (int b = 0; b < 3; b++) {
                    double v = acc[3 * a + b];
                    if (fr[a] || fq[b]) v = (fr[a] && myq == p && a == b) ? 1.0 : 0.0;   // SPC rows/columns
                    out[(int64_t)a * 3 * nb + b] = v;
                    if (myq == p && a == b) {                                            // Jacobi scaling
                        const double d = v > 0.0 ? 1.0 / sqrt(v) : 1.0;
                        d2[3 * rl + a] = d * d;
                    }
                }
        }
    }
}

__global__ void k_flag_chunk_elems(int64_t t_begin, int64_t t_end, const int32_t *__restrict__ inc, int32_t *__restrict__ flag) {
    int64_t t = t_begin + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t < t_end) flag[inc[t] >> 3] = 1;
}

__global__ void k_compact_chunk_elems(int64_t n_elem, const int32_t *__restrict__ flag, const int32_t *__restrict__ pos,
                                      int32_t *__restrict__ g2l, int32_t *__restrict__ lelem) {
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= n_elem) return;
    if (flag[e]) { g2l[e] = pos[e]; lelem[pos[e]] = (int32_t)e; }
    else g2l[e] = -1;
}

// 24 x 24 per element from the 36 stored blocks: a pair block and its transpose, diagonal blocks with their
// upper triangle mirrored — the values the assembly uses, laid out like Element.K_Initial's result
__global__ void k_expand_ke(int64_t count, const double *__restrict__ ke_store, const int32_t *__restrict__ qrec,
                            double *__restrict__ ke) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= count * 64) return;
    const int64_t le = t >> 6;
    const int i = (int)(t & 63) >> 3, j = (int)(t & 7);
    const double *src = ke_store + (le * 36 + (i <= j ? upper_index(i, j) : upper_index(j, i))) * KE_BLK;
    const int qi = qrec[8 * le + i], qj = qrec[8 * le + j];
    const bool as_stored = (qi < qj) || (qi == qj && i <= j);
    double *out = ke + le * 576 + (int64_t)(3 * i) * 24 + 3 * j;
    for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) {
            double v = as_stored ? src[3 * a + b] : src[3 * b + a];
            if (i == j && b < a) v = src[3 * b + a];
            out[a * 24 + b] = v;
        }
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

static int launch_ke(stan_handle *h, int64_t n_local, int64_t first, const int32_t *lelem, const int32_t *node_index,
                     double *ke_store, int32_t *qrec) {
    const char *v = getenv("STAN_KE");                 // 0 = two-phase kernel, default = warp-specialised persistent kernel
    if (v && atoi(v) == 0) {
        k_hex8_ke<<<div_up(n_local, KB_ELEMS), KB_THREADS, 0, h->stream>>>(n_local, first, lelem, h->d_conn.p, h->d_xyz.p, node_index,
                                                                           h->d_etype.p, h->d_emat.p, h->d_lambda.p, h->d_G.p,
                                                                           ke_store, qrec, h->d_err.p);
    } else {
        const int64_t nb = div_up(n_local, KB_ELEMS);
        const int grid = (int)std::min<int64_t>(nb, 2 * (int64_t)h->sm_count);
        k_hex8_ke_ws<<<grid, WS_THREADS, 0, h->stream>>>(n_local, first, lelem, h->d_conn.p, h->d_xyz.p, node_index,
                                                         h->d_etype.p, h->d_emat.p, h->d_lambda.p, h->d_G.p, ke_store, qrec,
                                                         h->d_err.p);
    }
    STAN_CUDA(cudaGetLastError());
    h->launches += 1;
    return STAN_OK;
}

int run_assembly(stan_handle *h) {
    cudaStream_t s = h->stream;
    const int64_t nloc = h->row1 - h->row0;
    STAN_TRY(h->d_vals.alloc((size_t)9 * h->n_blocks + 2, s));   // +2: 16-byte granules of the bulk-copy SpMV
    STAN_TRY(h->d_d2.alloc(3 * nloc, s));
    // Rows per chunk: all of them when the whole Ke store fits next to what is already allocated
    // (STAN_ASM_CHUNK_ROWS forces a chunk size — tests use it to run the multi-chunk path on small meshes).
    std::vector<int32_t> h_inc_ptr;
    int64_t n_inc_total = 0;
    {
        int32_t last = 0;
        STAN_CUDA(cudaMemcpyAsync(&last, h->d_inc_ptr.p + nloc, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        STAN_CUDA(cudaStreamSynchronize(s));
        n_inc_total = last;
    }
    const size_t per_elem = (size_t)KE_ELEM * sizeof(double) + 8 * sizeof(int32_t);
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    const size_t have = h->d_ke.n * sizeof(double);              // a store kept from an earlier assemble is reused
    const size_t budget = std::max<size_t>(have, free_b > ((size_t)3 << 30) ? free_b - ((size_t)2 << 30) : free_b / 2);
    const char *force = getenv("STAN_ASM_CHUNK_ROWS");
    int64_t chunk_rows = nloc;
    const int64_t n_touch_all = h->world == 1 ? h->n_elem : std::min<int64_t>(h->n_elem, n_inc_total);   // upper bound
    if (force && atoll(force) > 0) chunk_rows = std::min<int64_t>(nloc, atoll(force));
    else if ((size_t)n_touch_all * per_elem > budget) {
        // elements touched by a row range ~ incidence entries / 8 plus its boundary: half the budget for the estimate
        const double rows_fit = (double)budget / (double)per_elem * 8.0 / std::max<double>(1.0, (double)n_inc_total / nloc) * 0.5;
        chunk_rows = std::max<int64_t>(1024, std::min<int64_t>(nloc, (int64_t)rows_fit));
    }
    const bool single_chunk = chunk_rows >= nloc && h->world == 1;
    ScratchBuf<int32_t> g2l(&h->scratch[6]), lelem(&h->scratch[7]), qrec(&h->scratch[10]);
    ScratchBuf<int32_t> flag(&h->scratch[0]), pos(&h->scratch[1]);
    if (!single_chunk) {
        STAN_TRY(flag.alloc(h->n_elem + 1, s)); STAN_TRY(pos.alloc(h->n_elem + 1, s)); STAN_TRY(g2l.alloc(h->n_elem, s));
        if (chunk_rows < nloc) {
            h_inc_ptr.resize((size_t)nloc + 1);
            STAN_CUDA(cudaMemcpyAsync(h_inc_ptr.data(), h->d_inc_ptr.p, (nloc + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
            STAN_CUDA(cudaStreamSynchronize(s));
        }
    }
    for (int64_t r0 = 0; r0 < nloc; r0 += chunk_rows) {
        const int64_t r1 = std::min(nloc, r0 + chunk_rows);
        int64_t n_local = h->n_elem;
        if (!single_chunk) {                        // the elements that touch rows [r0, r1): flag, scan, compact
            const int64_t tb = h_inc_ptr.empty() ? 0 : h_inc_ptr[r0], te = h_inc_ptr.empty() ? n_inc_total : h_inc_ptr[r1];
            STAN_CUDA(cudaMemsetAsync(flag.p, 0, (h->n_elem + 1) * sizeof(int32_t), s));
            if (te > tb) k_flag_chunk_elems<<<div_up(te - tb, 256), 256, 0, s>>>(tb, te, h->d_inc.p, flag.p);
            STAN_TRY(device_exclusive_scan_i32(h, flag.p, pos.p, h->n_elem + 1, s));
            int32_t cnt = 0;
            STAN_CUDA(cudaMemcpyAsync(&cnt, pos.p + h->n_elem, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
            STAN_CUDA(cudaStreamSynchronize(s));
            n_local = cnt;
            STAN_TRY(lelem.alloc(std::max<int64_t>(n_local, 1), s));
            k_compact_chunk_elems<<<div_up(h->n_elem, 256), 256, 0, s>>>(h->n_elem, flag.p, pos.p, g2l.p, lelem.p);
            h->launches += 3;
        }
        if ((size_t)n_local * per_elem > budget && (size_t)n_local * KE_ELEM > h->d_ke.n) {
            set_error("assembly work area: the %lld elements of one row chunk need %.1f GB, %.1f GB are free",
                      (long long)n_local, n_local * (double)per_elem / 1e9, free_b / 1e9);
            return STAN_E_NOMEM;
        }
        STAN_TRY(h->d_ke.alloc((size_t)std::max<int64_t>(n_local, 1) * KE_ELEM, s));
        STAN_TRY(qrec.alloc((size_t)std::max<int64_t>(n_local, 1) * 8, s));
        if (n_local > 0)
            STAN_TRY(launch_ke(h, n_local, 0, single_chunk ? nullptr : lelem.p, h->d_node_index.p, h->d_ke.p, qrec.p));
        k_assemble_gather<<<div_up(r1 - r0, GA_WARPS), 32 * GA_WARPS, 0, s>>>(r0, r1, h->row0, h->d_inc_ptr.p, h->d_inc.p,
                                                                              h->d_brow_ptr.p, h->d_bcol.p, qrec.p,
                                                                              single_chunk ? nullptr : g2l.p, h->d_ke.p,
                                                                              h->d_fixed.p, h->d_vals.p, h->d_d2.p);
        STAN_CUDA(cudaGetLastError());
        h->launches += 1;
    }
    return STAN_OK;
}

// Element.K_Initial through the production integration kernel (see k_expand_ke for the layout)
int element_stiffness(stan_handle *h, int64_t first, int64_t count, double *ke_host) {
    cudaStream_t s = h->stream;
    STAN_TRY(h->d_err.alloc(8, s));
    STAN_CUDA(cudaMemsetAsync(h->d_err.p, 0, 8 * sizeof(int32_t), s));
    DevBuf<double> store, ke;
    DevBuf<int32_t> qrec;
    STAN_TRY(store.alloc((size_t)count * KE_ELEM, s)); STAN_TRY(ke.alloc((size_t)count * 576, s));
    STAN_TRY(qrec.alloc((size_t)count * 8, s));
    STAN_TRY(launch_ke(h, count, first, nullptr, h->have_dof ? h->d_node_index.p : nullptr, store.p, qrec.p));
    k_expand_ke<<<div_up(count * 64, 256), 256, 0, s>>>(count, store.p, qrec.p, ke.p);
    STAN_CUDA(cudaGetLastError());
    int32_t herr[4];
    STAN_CUDA(cudaMemcpyAsync(ke_host, ke.p, (size_t)count * 576 * sizeof(double), cudaMemcpyDeviceToHost, s));
    STAN_CUDA(cudaMemcpyAsync(herr, h->d_err.p, sizeof herr, cudaMemcpyDeviceToHost, s));
    STAN_CUDA(cudaStreamSynchronize(s));
    store.release(s); ke.release(s); qrec.release(s);
    h->launches += 1;
    if (herr[2]) { set_error("singular Jacobian (det == 0) in element range"); return STAN_E_SINGULAR; }
    return STAN_OK;
}

}  // namespace stan
