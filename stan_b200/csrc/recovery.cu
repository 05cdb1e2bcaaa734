// Gauss-point strain/stress recovery and extrapolation to the element nodes.
//
// Replaces Element.Recovery_Stress + Update_StrainStress
// (/root/reference/src/STAN_Database/Element.cs:211-246, 257-267) and the dU_buffer hand-off of
// /root/reference/src/STAN_Solver/Solver.cs:168-178.  The reference keeps BL[g] (6x24) and J[g]
// cached on every element from K_Initial (~9.8 KB per element); here they are recomputed from the
// 8 node coordinates, so nothing but the mesh and U is read and only the 2 x 48 results are written.
//
// One thread per (element, Gauss point) computes eps_g = BL[g] dU and sig_g = D eps_g; the 8
// threads of an element exchange them through shared memory and thread i forms the nodal values
// sum_g N[i][g] * value_g (FE_Library.cs:105-116).  HEX8_G1 has one Gauss point whose value every
// node receives (the reference indexes N[i][g] out of range there and throws).
#include "common.cuh"

namespace stan {

namespace {

__constant__ double c_rdNl[9][24];   // same table as assembly.cu (per-TU constant copy)
__constant__ double c_N[8][8];       // N[i][g], HEX8_ShapeFunctions(node i, 1/sqrt(3)) FE_Library.cs:285-321

constexpr int REC_THREADS = 256;     // 32 elements per CTA

// IEEE operations that the compiler may not contract into FMAs: like the element stiffness (assembly.cu), strain
// and stress follow the reference's operation order exactly — MatrixST.MultiplyVector sums j ascending, and a
// product with a structural zero of BL or D leaves the sum unchanged — so for the same U they are BIT-IDENTICAL
// to the CPU restatement.
__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }

__global__ void __launch_bounds__(REC_THREADS, 3)
k_recover(int64_t e_first, int64_t e_count, const int32_t *__restrict__ conn, const double *__restrict__ xyz,
          const int32_t *__restrict__ node_index, const uint8_t *__restrict__ etype, const int32_t *__restrict__ emat,
          const double *__restrict__ lam_tab, const double *__restrict__ G_tab, const double *__restrict__ ufull,
          double *__restrict__ strain, double *__restrict__ stress, int32_t *err,
          const int32_t *__restrict__ elem_list = nullptr) {
    // per element: 8 nodes x (x, y, z, ux, uy, uz); 49-double rows keep the 4 elements of a warp on distinct banks
    __shared__ double s_xu[REC_THREADS / 8][49];
    // Gauss-point strains (6) and stresses (6) in 16-byte aligned rows; 114 doubles per element put the 4 elements
    // of a warp on distinct banks for the 128-bit extrapolation reads
    __shared__ __align__(16) double s_val[REC_THREADS / 8][114];
    __shared__ double s_tab[9 * 25];               // dN_dLocal, 25-double rows: lanes index it by their own Gauss point
    __shared__ double s_N[64];
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t el = t >> 3;                         // local element index
    const int g = (int)(t & 7), le = threadIdx.x >> 3;
    const bool valid = el < e_count;
    const int64_t e = !valid ? 0 : (elem_list ? (int64_t)elem_list[el] : e_first + el);   // global element index
    if (threadIdx.x < 216) s_tab[25 * (threadIdx.x / 24) + threadIdx.x % 24] = (&c_rdNl[0][0])[threadIdx.x];
    if (threadIdx.x < 64) s_N[threadIdx.x] = (&c_N[0][0])[threadIdx.x];
    // the 8 threads of an element fetch one node each (coordinates and displacement, Element.cs:214-221)
    int type = STAN_HEX8_G2;
    double lam = 0.0, G = 0.0;
    if (valid) {
        const int32_t nd = conn[8 * e + g];
        const double *p = xyz + 3 * (int64_t)nd;
        const double *u = ufull + 3 * (int64_t)node_index[nd];
        double *d = s_xu[le];
        d[3 * g] = p[0]; d[3 * g + 1] = p[1]; d[3 * g + 2] = p[2];
        d[24 + 3 * g] = u[0]; d[24 + 3 * g + 1] = u[1]; d[24 + 3 * g + 2] = u[2];
        type = etype[e];
        const int mat = emat[e];
        lam = lam_tab[mat]; G = G_tab[mat];
    }
    __syncthreads();
    if (valid && (type == STAN_HEX8_G2 || g == 0)) {
        const double *tab = s_tab + 25 * ((type == STAN_HEX8_G2) ? g : 8);
        const double *X = s_xu[le], *U = s_xu[le] + 24;
        double J[9];                                   // J = dN_dLocal * X, each entry summed k ascending (Element.cs:274-292)
#pragma unroll
        for (int k = 0; k < 8; k++) {                  // node by node: 6 shared-memory reads feed 9 products
            const double t0 = tab[k], t1 = tab[8 + k], t2 = tab[16 + k];
            const double x0 = X[3 * k], x1 = X[3 * k + 1], x2 = X[3 * k + 2];
            if (k == 0) {
                J[0] = mul(t0, x0); J[1] = mul(t0, x1); J[2] = mul(t0, x2);
                J[3] = mul(t1, x0); J[4] = mul(t1, x1); J[5] = mul(t1, x2);
                J[6] = mul(t2, x0); J[7] = mul(t2, x1); J[8] = mul(t2, x2);
            } else {
                J[0] = add(J[0], mul(t0, x0)); J[1] = add(J[1], mul(t0, x1)); J[2] = add(J[2], mul(t0, x2));
                J[3] = add(J[3], mul(t1, x0)); J[4] = add(J[4], mul(t1, x1)); J[5] = add(J[5], mul(t1, x2));
                J[6] = add(J[6], mul(t2, x0)); J[7] = add(J[7], mul(t2, x1)); J[8] = add(J[8], mul(t2, x2));
            }
        }
        double det = mul(mul(J[0], J[4]), J[8]);       // MatrixST.Det3, MatrixST.cs:274-279
        det = add(det, mul(mul(J[3], J[7]), J[2]));
        det = add(det, mul(mul(J[6], J[1]), J[5]));
        det = sub(det, mul(mul(J[2], J[4]), J[6]));
        det = sub(det, mul(mul(J[0], J[5]), J[7]));
        det = sub(det, mul(mul(J[8], J[1]), J[3]));
        if (det == 0.0) atomicOr(err + 2, 1);
        const double inv = __ddiv_rn(1.0, det);
        double Ji[9];                                  // MatrixST.Inverse, MatrixST.cs:303-311
        Ji[0] = mul(inv, sub(mul(J[4], J[8]), mul(J[5], J[7])));
        Ji[1] = mul(inv, sub(mul(J[2], J[7]), mul(J[1], J[8])));
        Ji[2] = mul(inv, sub(mul(J[1], J[5]), mul(J[2], J[4])));
        Ji[3] = mul(inv, sub(mul(J[5], J[6]), mul(J[3], J[8])));
        Ji[4] = mul(inv, sub(mul(J[0], J[8]), mul(J[2], J[6])));
        Ji[5] = mul(inv, sub(mul(J[2], J[3]), mul(J[0], J[5])));
        Ji[6] = mul(inv, sub(mul(J[3], J[7]), mul(J[4], J[6])));
        Ji[7] = mul(inv, sub(mul(J[1], J[6]), mul(J[0], J[7])));
        Ji[8] = mul(inv, sub(mul(J[0], J[4]), mul(J[1], J[3])));
        // eps = BL0 dU (BL0_Matrix, Element.cs:316-324; MultiplyVector: columns ascending, i.e. node by node, x y z)
        double ex = 0, ey = 0, ez = 0, exy = 0, eyz = 0, exz = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const double t0 = tab[k], t1 = tab[8 + k], t2 = tab[16 + k];
            const double dx = add(add(mul(Ji[0], t0), mul(Ji[1], t1)), mul(Ji[2], t2));
            const double dy = add(add(mul(Ji[3], t0), mul(Ji[4], t1)), mul(Ji[5], t2));
            const double dz = add(add(mul(Ji[6], t0), mul(Ji[7], t1)), mul(Ji[8], t2));
            const double ux = U[3 * k], uy = U[3 * k + 1], uz = U[3 * k + 2];
            ex = add(ex, mul(dx, ux)); ey = add(ey, mul(dy, uy)); ez = add(ez, mul(dz, uz));
            exy = add(exy, mul(dy, ux)); exy = add(exy, mul(dx, uy));
            eyz = add(eyz, mul(dz, uy)); eyz = add(eyz, mul(dy, uz));
            exz = add(exz, mul(dz, ux)); exz = add(exz, mul(dx, uz));
        }
        const double d0 = add(lam, mul(2.0, G));       // Material.cs:42
        double *v = s_val[le] + 14 * g;
        v[0] = ex; v[1] = ey; v[2] = ez; v[3] = exy; v[4] = eyz; v[5] = exz;
        v[6] = add(add(mul(d0, ex), mul(lam, ey)), mul(lam, ez));      // D.MultiplyVector, Material.cs:42-53
        v[7] = add(add(mul(lam, ex), mul(d0, ey)), mul(lam, ez));
        v[8] = add(add(mul(lam, ex), mul(lam, ey)), mul(d0, ez));
        v[9] = mul(G, exy); v[10] = mul(G, eyz); v[11] = mul(G, exz);
    }
    __syncthreads();
    if (!valid) return;
    const int i = g;                                     // this thread now owns element node i
    double out[12];
    if (type == STAN_HEX8_G2) {
#pragma unroll
        for (int c = 0; c < 12; c++) out[c] = 0.0;
#pragma unroll
        for (int q = 0; q < 8; q++) {                    // Element.cs:238-245, g ascending
            const double w = s_N[8 * i + q];
            const double2 *vq = reinterpret_cast<const double2 *>(s_val[le] + 14 * q);
#pragma unroll
            for (int c = 0; c < 6; c++) {
                const double2 v2 = vq[c];
                out[2 * c] = add(out[2 * c], mul(v2.x, w));
                out[2 * c + 1] = add(out[2 * c + 1], mul(v2.y, w));
            }
        }
    } else {
#pragma unroll
        for (int c = 0; c < 12; c++) out[c] = s_val[le][c];
    }
    double *so = strain + el * 48 + i * 6, *to = stress + el * 48 + i * 6;   // Update_StrainStress :257-267
#pragma unroll
    for (int c = 0; c < 6; c += 2) {
        *reinterpret_cast<double2 *>(so + c) = make_double2(out[c], out[c + 1]);
        *reinterpret_cast<double2 *>(to + c) = make_double2(out[6 + c], out[6 + c + 1]);
    }
}

}  // namespace

static int upload_recovery_tables() {
    double tab[9 * 24];
    host_fe_tables(tab);
    STAN_CUDA(cudaMemcpyToSymbol(c_rdNl, tab, sizeof tab));
    static const double S[8][3] = {{-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1},
                                   {-1, -1, 1},  {1, -1, 1},  {1, 1, 1},  {-1, 1, 1}};
    const double gl = sqrt(1.0 / 3.0);
    double N[64];
    for (int i = 0; i < 8; i++) {
        const double xi = S[i][0] / gl, eta = S[i][1] / gl, zeta = S[i][2] / gl;
        for (int k = 0; k < 8; k++)
            N[i * 8 + k] = 1.0 / 8.0 * (1 + S[k][0] * xi) * (1 + S[k][1] * eta) * (1 + S[k][2] * zeta);
    }
    STAN_CUDA(cudaMemcpyToSymbol(c_N, N, sizeof N));
    return STAN_OK;
}

// strain/stress of an arbitrary element list (post-processing of a rank's nodes needs the elements
// around them, wherever ElemLib puts them); same kernel, same arithmetic as run_recovery
int recover_elements(stan_handle *h, const int32_t *d_list, int64_t count, double *d_strain, double *d_stress) {
    cudaStream_t s = h->stream;
    if (count <= 0) return STAN_OK;
    k_recover<<<div_up(8 * count, REC_THREADS), REC_THREADS, 0, s>>>(
        0, count, h->d_conn.p, h->d_xyz.p, h->d_node_index.p, h->d_etype.p, h->d_emat.p, h->d_lambda.p, h->d_G.p,
        h->d_ufull.p, d_strain, d_stress, h->d_err.p, d_list);
    STAN_CUDA(cudaGetLastError());
    h->launches += 1;
    return STAN_OK;
}

int run_recovery(stan_handle *h, stan_recovery_stats *stats) {
    cudaStream_t s = h->stream;
    STAN_TRY(upload_recovery_tables());
    STAN_TRY(scatter_solution(h));
    // every rank recovers (and later downloads) its own contiguous slice of ElemLib
    h->elem0 = h->n_elem * (int64_t)h->rank / h->world;
    h->elem1 = h->n_elem * (int64_t)(h->rank + 1) / h->world;
    const int64_t ne = h->elem1 - h->elem0;
    STAN_TRY(h->d_strain.alloc((size_t)48 * ne, s));
    STAN_TRY(h->d_stress.alloc((size_t)48 * ne, s));
    STAN_CUDA(cudaMemsetAsync(h->d_err.p, 0, 8 * sizeof(int32_t), s));
    STAN_CUDA(cudaEventRecord(h->ev0, s));
    k_recover<<<div_up(8 * ne, REC_THREADS), REC_THREADS, 0, s>>>(
        h->elem0, ne, h->d_conn.p, h->d_xyz.p, h->d_node_index.p, h->d_etype.p, h->d_emat.p, h->d_lambda.p, h->d_G.p,
        h->d_ufull.p, h->d_strain.p, h->d_stress.p, h->d_err.p);
    STAN_CUDA(cudaGetLastError());
    STAN_CUDA(cudaEventRecord(h->ev1, s));
    int32_t herr[4];
    STAN_CUDA(cudaMemcpyAsync(herr, h->d_err.p, sizeof herr, cudaMemcpyDeviceToHost, s));
    STAN_CUDA(cudaStreamSynchronize(s));
    if (herr[2]) { set_error("singular Jacobian (det == 0) during recovery"); return STAN_E_SINGULAR; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->ev0, h->ev1);
    h->launches += 1;
    if (stats) {
        stats->recover_ms = ms;
        stats->recover_bytes = ne * (32 + 768) + 24 * h->n_nodes * 2 / h->world;
        stats->kernel_launches = 1;
    }
    h->recovered = true;
    return STAN_OK;
}

}  // namespace stan
