/*
 * stan_oracle.c — CPU restatement of STAN's linear-static hot path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library, and only as the
 * checker (or the timed CPU baseline).  The product path (stan_b200/) never calls into it.
 *
 * PARITY UNPINNED: the reference (galuszkm/STAN, C#/.NET 4.6.1 + alglib.net 3.16.0) cannot
 * run in this environment (no dotnet/mono), ships no tests or golden vectors, and the ALGLIB
 * source is not part of the checkout.  Every function below therefore restates the cited
 * reference lines in the same arithmetic order (IEEE double, no FMA contraction: build with
 * -ffp-contract=off -fno-fast-math) and is pinned only by analytic known-answer tests
 * (tests/test_oracle_*.py) and by scipy cross-checks — never by reference outputs.
 *
 * Flat model convention (flattening of the reference's Dictionary<int,...> object graph,
 * enumeration order = insertion order = file order):
 *   xyz[3*n_nodes]      node coordinates in NodeLib order
 *   conn[8*n_elem]      0-based indices into the node array, ElemLib order, CHEXA node order
 *   elem_type[n_elem]   1 = HEX8_G1, 2 = HEX8_G2
 *   elem_mat[n_elem]    0-based index into E[]/nu[]
 *   node_index[n_nodes] BFS index of each node (DOF = 3*index + {0,1,2}, Node.cs:218-223)
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define OX __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------ */
/* MatrixST-order dense helpers (MatrixST.cs:404-427 operator*, :270-319 Det3/Inverse)     */
/* ------------------------------------------------------------------------------------ */

/* C = A(ar x ac) * B(ac x bc); C starts at zero and accumulates k ascending. */
static void mat_mul(const double *A, int ar, int ac, const double *B, int bc, double *C) {
    for (int i = 0; i < ar; i++)
        for (int j = 0; j < bc; j++) {
            double c = 0.0;
            for (int k = 0; k < ac; k++) c += A[i * ac + k] * B[k * bc + j];
            C[i * bc + j] = c;
        }
}

/* MatrixST.Det3, term order as written at MatrixST.cs:274-279. */
static double det3(const double *m) {
    return m[0] * m[4] * m[8] + m[3] * m[7] * m[2] + m[6] * m[1] * m[5] - m[2] * m[4] * m[6] -
           m[0] * m[5] * m[7] - m[8] * m[1] * m[3];
}

/* MatrixST.Inverse (adjugate times 1/det), MatrixST.cs:294-319.  Returns -1 when det == 0
 * (the reference throws ArgumentException there). */
static int inv3(const double *m, double *inv) {
    double det = det3(m);
    if (det == 0.0) return -1;
    double X = 1.0 / det;
    inv[0] = X * (m[4] * m[8] - m[5] * m[7]);
    inv[1] = X * (m[2] * m[7] - m[1] * m[8]);
    inv[2] = X * (m[1] * m[5] - m[2] * m[4]);
    inv[3] = X * (m[5] * m[6] - m[3] * m[8]);
    inv[4] = X * (m[0] * m[8] - m[2] * m[6]);
    inv[5] = X * (m[2] * m[3] - m[0] * m[5]);
    inv[6] = X * (m[3] * m[7] - m[4] * m[6]);
    inv[7] = X * (m[1] * m[6] - m[0] * m[7]);
    inv[8] = X * (m[0] * m[4] - m[1] * m[3]);
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* FE tables (FE_Library.cs:63-131, 206-321)                                              */
/* ------------------------------------------------------------------------------------ */

/* natural-coordinate signs of the 8 hex nodes, FE_Library.cs:108-115 / 216-223 */
static const double SGN[8][3] = {{-1, -1, -1}, {+1, -1, -1}, {+1, +1, -1}, {-1, +1, -1},
                                 {-1, -1, +1}, {+1, -1, +1}, {+1, +1, +1}, {-1, +1, +1}};

/* HEX8_Diff_ShapeFunctions: dN[r*8+k], r = xi/eta/zeta.  Each entry is
 * 1/8 * (c0 + c1*u + c2*v + c3*u*v) with the term order of FE_Library.cs:243-273
 * (for d/dxi: u = eta, v = zeta; d/deta: u = xi, v = zeta; d/dzeta: u = xi, v = eta). */
static void hex8_diff_shape(double xi, double eta, double zeta, double *dN) {
    for (int k = 0; k < 8; k++) {
        double sx = SGN[k][0], sy = SGN[k][1], sz = SGN[k][2];
        dN[0 * 8 + k] = 1.0 / 8.0 * (sx + (sx * sy) * eta + (sx * sz) * zeta + (sx * sy * sz) * (eta * zeta));
        dN[1 * 8 + k] = 1.0 / 8.0 * (sy + (sy * sx) * xi + (sy * sz) * zeta + (sy * sx * sz) * (xi * zeta));
        dN[2 * 8 + k] = 1.0 / 8.0 * (sz + (sz * sx) * xi + (sz * sy) * eta + (sz * sx * sy) * (xi * eta));
    }
}

/* HEX8_ShapeFunctions(Node_Coord, GaussPointLoc), FE_Library.cs:285-321 */
static void hex8_shape_extrap(const double *node_coord, double gloc, double *n) {
    double xi = node_coord[0] / gloc, eta = node_coord[1] / gloc, zeta = node_coord[2] / gloc;
    for (int k = 0; k < 8; k++)
        n[k] = 1.0 / 8.0 * (1 + SGN[k][0] * xi) * (1 + SGN[k][1] * eta) * (1 + SGN[k][2] * zeta);
}

/* type: 1 = HEX8_G1 (FE_Library.cs:63-89), 2 = HEX8_G2 (:91-131).
 * dN_dLocal: ngp x 3 x 8.  N: G2 -> 8 x 8 as N[i][g]; G1 -> single row of ones (N[0][0..7]). */
OX int stan_oracle_hex8_tables(int type, int *ngp, double *weight, double *dN_dLocal, double *N) {
    if (type == 1) {
        *ngp = 1;
        *weight = 2.0 * 2.0 * 2.0;
        hex8_diff_shape(0.0, 0.0, 0.0, dN_dLocal);
        if (N) for (int k = 0; k < 8; k++) N[k] = 1.0;
        return 0;
    }
    if (type == 2) {
        *ngp = 8;
        *weight = 1.0;
        double g = sqrt(1.0 / 3.0);
        for (int q = 0; q < 8; q++) {
            hex8_diff_shape(SGN[q][0] * g, SGN[q][1] * g, SGN[q][2] * g, dN_dLocal + q * 24);
            if (N) hex8_shape_extrap(SGN[q], g, N + q * 8);
        }
        return 0;
    }
    return -1;
}

/* Material.SetElastic, Material.cs:31-56 */
OX void stan_oracle_elastic_D(double E, double Poisson, double *D) {
    memset(D, 0, 36 * sizeof(double));
    double lambda = (E * Poisson) / ((1 - 2 * Poisson) * (1 + Poisson));
    double G = (0.5 * E) / (1 + Poisson);
    D[0] = lambda + (2 * G); D[1] = lambda; D[2] = lambda;
    D[6] = lambda; D[7] = lambda + (2 * G); D[8] = lambda;
    D[12] = lambda; D[13] = lambda; D[14] = lambda + (2 * G);
    D[21] = G; D[28] = G; D[35] = G;
}

/* ------------------------------------------------------------------------------------ */
/* R2: Element.K_Initial (Element.cs:118-155) with Jacobian (:274-292), BL0 (:297-328)     */
/* ------------------------------------------------------------------------------------ */

/* BL0_Matrix, Element.cs:314-325.  BL is 6 x 24 row-major. */
static void bl0_matrix(const double *dN /*3x8*/, double *BL) {
    memset(BL, 0, 144 * sizeof(double));
    for (int i = 0; i < 8; i++) {
        BL[0 * 24 + 3 * i + 0] = dN[0 * 8 + i];
        BL[1 * 24 + 3 * i + 1] = dN[1 * 8 + i];
        BL[2 * 24 + 3 * i + 2] = dN[2 * 8 + i];
        BL[3 * 24 + 3 * i + 0] = dN[1 * 8 + i];
        BL[3 * 24 + 3 * i + 1] = dN[0 * 8 + i];
        BL[4 * 24 + 3 * i + 1] = dN[2 * 8 + i];
        BL[4 * 24 + 3 * i + 2] = dN[1 * 8 + i];
        BL[5 * 24 + 3 * i + 0] = dN[2 * 8 + i];
        BL[5 * 24 + 3 * i + 2] = dN[0 * 8 + i];
    }
}

/* One element.  X: 8x3 nodal coordinates.  K: 24x24 row-major.  Optional outputs Jg
 * (ngp x 9) and BLg (ngp x 144) are the per-Gauss-point caches the reference keeps on the
 * element (Element.cs:127,143).  BL1 is identically zero in linear statics because
 * GetDisp(1, .) == 0 at this point (Element.cs:135-143, Node.cs:109-116), so BL = BL0.
 * Returns -1 if a Jacobian determinant is exactly zero (MatrixST.cs:298,317 throws). */
OX int stan_oracle_k_initial(int type, const double *X, const double *D, double *K, double *Jg,
                             double *BLg) {
    int ngp;
    double w, dNl[8 * 24];
    if (stan_oracle_hex8_tables(type, &ngp, &w, dNl, NULL)) return -2;
    memset(K, 0, 576 * sizeof(double));
    double J[9], Ji[9], dN[24], BL[144], BT[144], T1[144], T2[576];
    for (int g = 0; g < ngp; g++) {
        mat_mul(dNl + g * 24, 3, 8, X, 3, J);              /* Element.cs:127, :290 */
        if (inv3(J, Ji)) return -1;
        mat_mul(Ji, 3, 3, dNl + g * 24, 8, dN);            /* Element.cs:130 */
        bl0_matrix(dN, BL);                                /* Element.cs:143 */
        for (int r = 0; r < 6; r++)
            for (int c = 0; c < 24; c++) BT[c * 6 + r] = BL[r * 24 + c];
        mat_mul(BT, 24, 6, D, 6, T1);                      /* (BL^T * D)            :151 */
        mat_mul(T1, 24, 6, BL, 24, T2);                    /* ... * BL              :151 */
        double s = det3(J) * w;                            /* J.Det3() * GaussWeight     */
        for (int i = 0; i < 576; i++) K[i] = 0.0 + (K[i] + T2[i] * s); /* MultiplyScalar, operator+ */
        if (Jg) memcpy(Jg + g * 9, J, sizeof J);
        if (BLg) memcpy(BLg + g * 144, BL, sizeof BL);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* R0: Database.AssignDOF (Database.cs:140-234), literal duplicate-tolerant FIFO           */
/* ------------------------------------------------------------------------------------ */

/* node -> element lists in element order with per-node de-duplication
 * (AddElem2Nodes + RemoveElemDuplicates, Database.cs:143-158). */
static int build_n2e(int n_nodes, int n_elem, const int32_t *conn, int64_t **ptr_out, int32_t **idx_out) {
    int64_t *ptr = calloc((size_t)n_nodes + 1, sizeof(int64_t));
    int32_t *last = malloc((size_t)n_nodes * sizeof(int32_t));
    if (!ptr || !last) return -1;
    for (int i = 0; i < n_nodes; i++) last[i] = -1;
    for (int e = 0; e < n_elem; e++)
        for (int k = 0; k < 8; k++) {
            int n = conn[8 * (int64_t)e + k];
            if (n < 0 || n >= n_nodes) { free(ptr); free(last); return -2; }
            if (last[n] != e) { last[n] = e; ptr[n + 1]++; }
        }
    for (int i = 0; i < n_nodes; i++) ptr[i + 1] += ptr[i];
    int32_t *idx = malloc((size_t)(ptr[n_nodes] > 0 ? ptr[n_nodes] : 1) * sizeof(int32_t));
    int64_t *fill = malloc((size_t)n_nodes * sizeof(int64_t));
    if (!idx || !fill) return -1;
    memcpy(fill, ptr, (size_t)n_nodes * sizeof(int64_t));
    for (int i = 0; i < n_nodes; i++) last[i] = -1;
    for (int e = 0; e < n_elem; e++)
        for (int k = 0; k < 8; k++) {
            int n = conn[8 * (int64_t)e + k];
            if (last[n] != e) { last[n] = e; idx[fill[n]++] = e; }
        }
    free(fill); free(last);
    *ptr_out = ptr; *idx_out = idx;
    return 0;
}

/* Returns 0 on success; -3 no start node with 1..6 incident elements (reference: FirstNode = 0
 * -> KeyNotFoundException); -4 disconnected mesh (reference: index out of range at :218). */
OX int stan_oracle_assign_dof(int n_nodes, int n_elem, const int32_t *conn, int32_t *node_index) {
    int64_t *ptr; int32_t *idx;
    int rc = build_n2e(n_nodes, n_elem, conn, &ptr, &idx);
    if (rc) return rc;
    /* Database.cs:178-196: first node (NodeLib order) with exactly i incident elements, i=1..6 */
    int first = -1;
    for (int i = 1; i < 7 && first < 0; i++)
        for (int n = 0; n < n_nodes; n++)
            if (ptr[n + 1] - ptr[n] == i) { first = n; break; }
    if (first < 0) { free(ptr); free(idx); return -3; }

    uint8_t *done = calloc((size_t)n_nodes, 1);
    int32_t *stamp = malloc((size_t)n_nodes * sizeof(int32_t)); /* Distinct() helper */
    for (int i = 0; i < n_nodes; i++) stamp[i] = -1;
    size_t qcap = (size_t)n_nodes * 4 + 64, qlen = 0;
    int32_t *queue = malloc(qcap * sizeof(int32_t));
    int32_t nb[4096];

    /* Neighbors[N]: concatenated NLists of incident elements, first-occurrence Distinct,
     * self removed (Database.cs:161-176).  Computed lazily with identical order. */
#define NEIGHBORS(N, OUT, CNT)                                                    \
    do {                                                                          \
        CNT = 0;                                                                  \
        stamp[N] = N; /* self is removed after Distinct: never emit it */          \
        for (int64_t t = ptr[N]; t < ptr[N + 1]; t++)                             \
            for (int k = 0; k < 8; k++) {                                         \
                int m = conn[8 * (int64_t)idx[t] + k];                            \
                if (stamp[m] != N) { stamp[m] = N; if (CNT < 4096) OUT[CNT] = m; CNT++; } \
            }                                                                     \
    } while (0)

    int index = 0, cnt;
    node_index[first] = index++;                       /* :203-205 */
    done[first] = 1;
    NEIGHBORS(first, nb, cnt);
    if (cnt > 4096) { rc = -5; goto out; }
    for (int i = 0; i < cnt; i++) queue[qlen++] = nb[i]; /* NextNode = Neighbors[FirstNode] */
    size_t index2 = 0;
    while (index < n_nodes) {                          /* :210-233 */
        if (index2 >= qlen) { rc = -4; goto out; }
        int nid = queue[index2];
        if (!done[nid]) {
            node_index[nid] = index++;
            done[nid] = 1;
            NEIGHBORS(nid, nb, cnt);
            if (cnt > 4096) { rc = -5; goto out; }
            if (qlen + (size_t)cnt > qcap) {
                qcap = qcap * 2 + (size_t)cnt;
                queue = realloc(queue, qcap * sizeof(int32_t));
            }
            for (int i = 0; i < cnt; i++)
                if (!done[nb[i]]) queue[qlen++] = nb[i];
        }
        index2++;
    }
out:
    free(queue); free(stamp); free(done); free(ptr); free(idx);
    return rc;
}

/* ------------------------------------------------------------------------------------ */
/* R1: SPC elimination + RHS (Solver.cs:104-152), Include/Exclude_BC_DOF (SolverFunctions.cs:520-555) */
/* ------------------------------------------------------------------------------------ */

/* spc_node[i] (0-based node), spc_val[3*i+dir]: a DOF is fixed when the value == 1.
 * Returns the number of fixed DOFs (|Distinct(Fix_DOF)|); ndof_reduction as Solver.cs:121-132. */
OX int stan_oracle_spc_reduction(int n_nodes, const int32_t *node_index, int n_spc,
                                 const int32_t *spc_node, const double *spc_val,
                                 int32_t *ndof_reduction) {
    int ndof = 3 * n_nodes;
    memset(ndof_reduction, 0, (size_t)ndof * sizeof(int32_t));
    for (int i = 0; i < n_spc; i++)
        for (int d = 0; d < 3; d++)
            if (spc_val[3 * i + d] == 1) ndof_reduction[3 * node_index[spc_node[i]] + d] = -1;
    int reduc = 0;
    for (int i = 0; i < ndof; i++) {
        if (ndof_reduction[i] == -1) reduc++;
        else ndof_reduction[i] = reduc;
    }
    return reduc;
}

/* F[reduced] += load for free DOFs, in list order (Solver.cs:136-152). F has n_free entries. */
OX void stan_oracle_build_rhs(int n_nodes, const int32_t *node_index, const int32_t *ndof_reduction,
                              int n_load, const int32_t *load_node, const double *load_val, double *F) {
    int ndof = 3 * n_nodes, nfix = 0;
    for (int i = 0; i < ndof; i++) nfix += (ndof_reduction[i] == -1);
    memset(F, 0, (size_t)(ndof - nfix) * sizeof(double));
    for (int i = 0; i < n_load; i++)
        for (int dir = 0; dir < 3; dir++) {
            int dof = 3 * node_index[load_node[i]] + dir;
            if (ndof_reduction[dof] != -1) F[dof - ndof_reduction[dof]] += load_val[3 * i + dir];
        }
}

OX void stan_oracle_include_bc_dof(int ndof, const int32_t *ndof_reduction, const double *A, double *A_full) {
    for (int i = 0; i < ndof; i++)
        A_full[i] = (ndof_reduction[i] == -1) ? 0.0 : A[i - ndof_reduction[i]];
}

OX void stan_oracle_exclude_bc_dof(int ndof, const int32_t *ndof_reduction, const double *A, double *A_red) {
    for (int i = 0; i < ndof; i++)
        if (ndof_reduction[i] != -1) A_red[i - ndof_reduction[i]] = A[i];
}

/* ------------------------------------------------------------------------------------ */
/* R3: ParallelAssembly_K (SolverFunctions.cs:117-180) + sparseconverttocrs (:275)         */
/* ------------------------------------------------------------------------------------ */

typedef struct {
    int64_t n;        /* reduced dimension */
    int64_t nnz;
    int64_t *rowptr;  /* n + 1 */
    int32_t *col;     /* ascending within row, col >= row */
    double *val;
} oracle_csr;

static int cmp_i32(const void *a, const void *b) {
    int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
    return (x > y) - (x < y);
}

OX void stan_oracle_csr_free(oracle_csr *m) {
    if (!m) return;
    free(m->rowptr); free(m->col); free(m->val); free(m);
}
/* An upper-triangle CRS built from caller arrays: lets the tests feed hand-worked matrices to the
 * ALGLIB restatements (lincg, skyline Cholesky) without going through the mesh assembly. */
OX oracle_csr *stan_oracle_csr_from_arrays(int64_t n, const int64_t *rowptr, const int32_t *col, const double *val) {
    oracle_csr *m = calloc(1, sizeof *m);
    m->n = n;
    m->nnz = rowptr[n];
    m->rowptr = malloc((size_t)(n + 1) * sizeof(int64_t));
    m->col = malloc((size_t)(m->nnz + 1) * sizeof(int32_t));
    m->val = malloc((size_t)(m->nnz + 1) * sizeof(double));
    memcpy(m->rowptr, rowptr, (size_t)(n + 1) * sizeof(int64_t));
    memcpy(m->col, col, (size_t)m->nnz * sizeof(int32_t));
    memcpy(m->val, val, (size_t)m->nnz * sizeof(double));
    return m;
}
OX int64_t stan_oracle_csr_n(const oracle_csr *m) { return m->n; }
OX int64_t stan_oracle_csr_nnz(const oracle_csr *m) { return m->nnz; }
OX void stan_oracle_csr_copy(const oracle_csr *m, int64_t *rowptr, int32_t *col, double *val) {
    if (rowptr) memcpy(rowptr, m->rowptr, (size_t)(m->n + 1) * sizeof(int64_t));
    if (col) memcpy(col, m->col, (size_t)m->nnz * sizeof(int32_t));
    if (val) memcpy(val, m->val, (size_t)m->nnz * sizeof(double));
}

/*
 * Upper-triangle (col >= row) CRS of the reduced system, rows/cols ascending.
 * Element contributions are added in ElemLib order (the reference's order is
 * nondeterministic: Parallel.ForEach + lock, SolverFunctions.cs:129,162).
 *   prune == 0: structural pattern — every free (row, col >= row) pair coupled by an element.
 *   prune == 1: ALGLIB hash-table semantics as recalled in SURVEY.md Appendix A — sparseadd()
 *               ignores v == 0 and drops an entry whose running sum becomes exactly 0.0.
 * *exact_zero receives the count of structural slots whose final value is exactly 0.0.
 * Returns NULL on error (*err: -1 singular Jacobian, -6 out of memory, -7 bad material).
 */
OX oracle_csr *stan_oracle_assemble_upper(int n_nodes, const double *xyz, int n_elem, const int32_t *conn,
                                          const uint8_t *elem_type, const int32_t *elem_mat, int n_mat,
                                          const double *E, const double *nu, const int32_t *node_index,
                                          const int32_t *ndof_reduction, int prune, int64_t *exact_zero,
                                          int *err) {
    *err = 0;
    int ndof = 3 * n_nodes;
    int64_t *nptr; int32_t *nidx;
    if (build_n2e(n_nodes, n_elem, conn, &nptr, &nidx)) { *err = -6; return NULL; }
    int32_t *inv = malloc((size_t)n_nodes * sizeof(int32_t));
    for (int i = 0; i < n_nodes; i++) inv[node_index[i]] = i;

    /* block pattern in BFS space: neighbours q >= p of each BFS node p, sorted */
    int64_t *bptr = calloc((size_t)n_nodes + 1, sizeof(int64_t));
    int32_t *stamp = malloc((size_t)n_nodes * sizeof(int32_t));
    int32_t *bcol_tmp = NULL;
    for (int pass = 0; pass < 2; pass++) {
        for (int i = 0; i < n_nodes; i++) stamp[i] = -1;
        if (pass == 1) {
            for (int p = 0; p < n_nodes; p++) bptr[p + 1] += bptr[p];
            bcol_tmp = malloc((size_t)(bptr[n_nodes] + 1) * sizeof(int32_t));
        }
        for (int p = 0; p < n_nodes; p++) {
            int node = inv[p];
            int64_t w = (pass == 1) ? bptr[p] : 0, w0 = w;
            for (int64_t t = nptr[node]; t < nptr[node + 1]; t++)
                for (int k = 0; k < 8; k++) {
                    int q = node_index[conn[8 * (int64_t)nidx[t] + k]];
                    if (q >= p && stamp[q] != p) {
                        stamp[q] = p;
                        if (pass == 1) bcol_tmp[w] = q;
                        w++;
                    }
                }
            if (pass == 0) bptr[p + 1] = w;
            else qsort(bcol_tmp + w0, (size_t)(w - w0), sizeof(int32_t), cmp_i32);
        }
    }
    {
        {
            /* scalar reduced rows */
            int nfix = 0;
            for (int i = 0; i < ndof; i++) nfix += (ndof_reduction[i] == -1);
            int64_t n = ndof - nfix;
            oracle_csr *m = calloc(1, sizeof *m);
            m->n = n;
            m->rowptr = calloc((size_t)n + 1, sizeof(int64_t));
            for (int p = 0; p < n_nodes; p++)
                for (int a = 0; a < 3; a++) {
                    int d = 3 * p + a;
                    if (ndof_reduction[d] == -1) continue;
                    int64_t cnt = 0;
                    for (int64_t s = bptr[p]; s < bptr[p + 1]; s++)
                        for (int b = 0; b < 3; b++) {
                            int c = 3 * bcol_tmp[s] + b;
                            if (c >= d && ndof_reduction[c] != -1) cnt++;
                        }
                    m->rowptr[d - ndof_reduction[d] + 1] = cnt;
                }
            for (int64_t i = 0; i < n; i++) m->rowptr[i + 1] += m->rowptr[i];
            m->nnz = m->rowptr[n];
            m->col = malloc((size_t)(m->nnz + 1) * sizeof(int32_t));
            m->val = calloc((size_t)m->nnz + 1, sizeof(double));
            uint8_t *present = prune ? calloc((size_t)m->nnz + 1, 1) : NULL;
            if (!m->col || !m->val) { *err = -6; return NULL; }
            for (int p = 0; p < n_nodes; p++)
                for (int a = 0; a < 3; a++) {
                    int d = 3 * p + a;
                    if (ndof_reduction[d] == -1) continue;
                    int64_t w2 = m->rowptr[d - ndof_reduction[d]];
                    for (int64_t s = bptr[p]; s < bptr[p + 1]; s++)
                        for (int b = 0; b < 3; b++) {
                            int c = 3 * bcol_tmp[s] + b;
                            if (c >= d && ndof_reduction[c] != -1) m->col[w2++] = c - ndof_reduction[c];
                        }
                }
            free(bcol_tmp);

            /* element loop: Ke in parallel chunks, scatter strictly in element order */
            double *Dm = malloc((size_t)n_mat * 36 * sizeof(double));
            for (int i = 0; i < n_mat; i++) stan_oracle_elastic_D(E[i], nu[i], Dm + 36 * i);
            const int CH = 4096;
            double *Kbuf = malloc((size_t)CH * 576 * sizeof(double));
            int bad = 0;
            for (int e0 = 0; e0 < n_elem && !bad; e0 += CH) {
                int e1 = e0 + CH < n_elem ? e0 + CH : n_elem;
#pragma omp parallel for schedule(static) reduction(| : bad)
                for (int e = e0; e < e1; e++) {
                    double X[24];
                    for (int k = 0; k < 8; k++)
                        for (int c = 0; c < 3; c++) X[3 * k + c] = xyz[3 * (int64_t)conn[8 * (int64_t)e + k] + c];
                    int mat = elem_mat[e];
                    if (mat < 0 || mat >= n_mat) { bad |= 2; continue; }
                    if (stan_oracle_k_initial(elem_type[e], X, Dm + 36 * mat, Kbuf + (size_t)(e - e0) * 576, NULL, NULL))
                        bad |= 1;
                }
                if (bad) break;
                for (int e = e0; e < e1; e++) {
                    const double *k = Kbuf + (size_t)(e - e0) * 576;
                    const int32_t *cn = conn + 8 * (int64_t)e;
                    /* SolverFunctions.cs:143-172 loop nest: i, m, j, n */
                    for (int i = 0; i < 8; i++)
                        for (int mm = 0; mm < 3; mm++) {
                            int row = 3 * node_index[cn[i]] + mm;
                            if (ndof_reduction[row] == -1) continue;
                            int64_t rr = row - ndof_reduction[row];
                            for (int j = 0; j < 8; j++)
                                for (int nn = 0; nn < 3; nn++) {
                                    int col = 3 * node_index[cn[j]] + nn;
                                    if (col < row || ndof_reduction[col] == -1) continue;
                                    int32_t cc = col - ndof_reduction[col];
                                    int64_t lo = m->rowptr[rr], hi = m->rowptr[rr + 1] - 1;
                                    while (lo < hi) {
                                        int64_t mid = (lo + hi) >> 1;
                                        if (m->col[mid] < cc) lo = mid + 1; else hi = mid;
                                    }
                                    double v = k[(i * 3 + mm) * 24 + j * 3 + nn];
                                    if (!prune) m->val[lo] += v;
                                    else if (v != 0.0) { /* alglib.sparseadd semantics */
                                        if (!present[lo]) { present[lo] = 1; m->val[lo] = v; }
                                        else { m->val[lo] += v; if (m->val[lo] == 0.0) present[lo] = 0; }
                                    }
                                }
                        }
                }
            }
            free(Kbuf); free(Dm);
            if (bad) { *err = (bad & 1) ? -1 : -7; stan_oracle_csr_free(m); m = NULL; }
            if (m) {
                int64_t z = 0;
                for (int64_t i = 0; i < m->nnz; i++) z += (m->val[i] == 0.0);
                if (exact_zero) *exact_zero = z;
                if (prune) { /* compact to the entries still present (sparseconverttocrs) */
                    int64_t w3 = 0, r0 = 0;
                    for (int64_t r = 0; r < n; r++) {
                        int64_t r1 = m->rowptr[r + 1];
                        for (int64_t t = r0; t < r1; t++)
                            if (present[t]) { m->col[w3] = m->col[t]; m->val[w3] = m->val[t]; w3++; }
                        r0 = r1;
                        m->rowptr[r + 1] = w3;
                    }
                    m->nnz = w3;
                }
            }
            free(present); free(bptr); free(stamp); free(inv); free(nptr); free(nidx);
            return m;
        }
    }
}

/* ------------------------------------------------------------------------------------ */
/* R5: alglib.lincg as used by LinearSolver_CG (SolverFunctions.cs:270-330)                 */
/*     restated from SURVEY.md Appendix A (ALGLIB 3.16 source is not in the checkout).      */
/* ------------------------------------------------------------------------------------ */

typedef struct {
    int32_t terminationtype;   /* 1, 5, 7, -4, -5 (SolverFunctions.cs:311-319) */
    int32_t iterationscount;
    int32_t nmv;
    int32_t reserved;
    double r2;                 /* squared 2-norm of the final residual */
    double bnorm;
} oracle_cg_report;

typedef struct {
    double epsf;               /* lincgsetcond EpsF; (0,0) -> 1e-6 */
    int32_t maxits;            /* 0 = unlimited */
    int32_t its_before_rupdate;/* ALGLIB default 10; 0 = never refresh */
    int32_t its_before_restart;/* ALGLIB default n (pass 0) */
    int32_t merit_check;       /* 1 = ALGLIB behaviour (type 7 when the energy functional stalls) */
    int32_t zero_based_counter;/* 0: refresh/restart tests use k = 1,2,... (Appendix A as written);
                                  1: tests use k-1 (counter incremented at loop end) */
    int32_t parallel_spmv;     /* 0: serial symmetric product in ALGLIB's sparsesmv order;
                                  1: expanded full CSR, OpenMP row-parallel (CPU baseline timing) */
    int32_t dot_mode;          /* 0: plain left-to-right sums (ALGLIB); 1: blocked sums (1024-term partials added
                                  in order) — a second, equally valid rounding of the same recurrences, used to
                                  measure how far summation order alone moves the trajectory */
} oracle_cg_opts;

/* y = (U + U^T - diag) x from the upper CRS, in the accumulation order of ALGLIB's sparsesmv
 * for CRS storage: rows ascending; y[i] += d_ii x_i; then for j > i: vy += a_ij x_j and
 * y[j] += a_ij x_i; finally y[i] += vy. */
static void sym_spmv_upper(const oracle_csr *A, const double *x, double *y) {
    int64_t n = A->n;
    for (int64_t i = 0; i < n; i++) y[i] = 0.0;
    for (int64_t i = 0; i < n; i++) {
        int64_t t = A->rowptr[i], t1 = A->rowptr[i + 1];
        if (t < t1 && A->col[t] == i) { y[i] += A->val[t] * x[i]; t++; }
        double vy = 0.0, vx = x[i];
        for (; t < t1; t++) {
            int32_t id = A->col[t];
            double v = A->val[t];
            vy += x[id] * v;
            y[id] += vx * v;
        }
        y[i] += vy;
    }
}

typedef struct { int64_t n; int64_t *rowptr; int32_t *col; double *val; } full_csr;

static full_csr *expand_full(const oracle_csr *A) {
    int64_t n = A->n;
    full_csr *F = calloc(1, sizeof *F);
    F->n = n;
    F->rowptr = calloc((size_t)n + 1, sizeof(int64_t));
    for (int64_t i = 0; i < n; i++)
        for (int64_t t = A->rowptr[i]; t < A->rowptr[i + 1]; t++) {
            F->rowptr[i + 1]++;
            if (A->col[t] != i) F->rowptr[A->col[t] + 1]++;
        }
    for (int64_t i = 0; i < n; i++) F->rowptr[i + 1] += F->rowptr[i];
    F->col = malloc((size_t)(F->rowptr[n] + 1) * sizeof(int32_t));
    F->val = malloc((size_t)(F->rowptr[n] + 1) * sizeof(double));
    int64_t *fill = malloc((size_t)n * sizeof(int64_t));
    memcpy(fill, F->rowptr, (size_t)n * sizeof(int64_t));
    /* lower parts first (rows ascending gives ascending columns), then the upper row */
    for (int64_t i = 0; i < n; i++)
        for (int64_t t = A->rowptr[i]; t < A->rowptr[i + 1]; t++)
            if (A->col[t] != i) { int64_t j = A->col[t]; F->col[fill[j]] = (int32_t)i; F->val[fill[j]++] = A->val[t]; }
    for (int64_t i = 0; i < n; i++)
        for (int64_t t = A->rowptr[i]; t < A->rowptr[i + 1]; t++) { F->col[fill[i]] = A->col[t]; F->val[fill[i]++] = A->val[t]; }
    free(fill);
    return F;
}

static void full_spmv(const full_csr *F, const double *x, double *y) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < F->n; i++) {
        double s = 0.0;
        for (int64_t t = F->rowptr[i]; t < F->rowptr[i + 1]; t++) s += F->val[t] * x[F->col[t]];
        y[i] = s;
    }
}

static int g_dot_mode = 0;   /* set per solve from oracle_cg_opts.dot_mode (solves are not concurrent) */

static double dot_seq(const double *a, const double *b, int64_t n) {
    double s = 0.0;
    if (!g_dot_mode) {
        for (int64_t i = 0; i < n; i++) s += a[i] * b[i];
        return s;
    }
    for (int64_t i0 = 0; i0 < n; i0 += 1024) {
        int64_t i1 = i0 + 1024 < n ? i0 + 1024 : n;
        double t = 0.0;
        for (int64_t i = i0; i < i1; i++) t += a[i] * b[i];
        s += t;
    }
    return s;
}

/* One row of `hist` per completed iteration k = 1, 2, ...: { ||r_k||^2, alpha_k, beta_k (0 on a restart or
 * when the solve ended at k), energy functional x'Ax - 2b'x on refresh iterations else NaN }.
 * The same four numbers are exported by libstan_b200's stan_get_cg_history; the trajectory test compares them. */
OX int stan_oracle_lincg_hist(const oracle_csr *A, const double *b, const oracle_cg_opts *o, double *x,
                              oracle_cg_report *rep, double *hist, int64_t hist_cap) {
    int64_t n = A->n;
    g_dot_mode = o->dot_mode;
#define HIST(K, R2, AL, BE, ME) do { if (hist && (K) >= 1 && (K) <= hist_cap) { double *h_ = hist + 4 * ((K) - 1); \
        h_[0] = (R2); h_[1] = (AL); h_[2] = (BE); h_[3] = (ME); } } while (0)
    double epsf = o->epsf;
    int maxits = o->maxits;
    if (epsf == 0.0 && maxits == 0) epsf = 1.0e-6;          /* lincgsetcond note, SolverFunctions.cs:292-293 */
    int rupd = o->its_before_rupdate;
    int64_t restart = o->its_before_restart > 0 ? o->its_before_restart : n;
    int off = o->zero_based_counter ? 1 : 0;
    full_csr *F = o->parallel_spmv ? expand_full(A) : NULL;
#define SPMV(X, Y) do { if (F) full_spmv(F, X, Y); else sym_spmv_upper(A, X, Y); rep->nmv++; } while (0)

    double *d2 = malloc((size_t)n * sizeof(double));
    double *r = malloc((size_t)n * sizeof(double)), *z = malloc((size_t)n * sizeof(double));
    double *p = malloc((size_t)n * sizeof(double)), *mv = malloc((size_t)n * sizeof(double));
    double *cx = malloc((size_t)n * sizeof(double)), *cr = malloc((size_t)n * sizeof(double));
    memset(rep, 0, sizeof *rep);
    int rc = 0;

    /* diagonal scaling: d = 1/sqrt(A_ii) if A_ii > 0 else 1; preconditioner = r * d^2 */
    for (int64_t i = 0; i < n; i++) {
        double v = 0.0;
        int64_t t = A->rowptr[i];
        if (t < A->rowptr[i + 1] && A->col[t] == i) v = A->val[t];
        double d = v > 0.0 ? 1 / sqrt(v) : 1.0;
        d2[i] = d * d;
    }
    double bnorm = sqrt(dot_seq(b, b, n));
    rep->bnorm = bnorm;
    for (int64_t i = 0; i < n; i++) x[i] = 0.0;
    if (bnorm == 0.0) { rep->terminationtype = 1; goto done; }

    SPMV(x, mv);                                            /* r0 = b - A x0 */
    double r2 = 0.0, merit = 0.0;
    for (int64_t i = 0; i < n; i++) {
        r[i] = b[i] - mv[i];
        r2 += r[i] * r[i];
        merit += mv[i] * x[i] - 2 * b[i] * x[i];
    }
    rep->r2 = r2;
    if (!isfinite(r2)) { rep->terminationtype = -4; goto done; }
    if (sqrt(r2) <= epsf * bnorm) { rep->terminationtype = 1; goto done; }
    for (int64_t i = 0; i < n; i++) { z[i] = r[i] * d2[i]; p[i] = z[i]; }

    for (int k = 1;; k++) {
        SPMV(p, mv);
        double vmv = dot_seq(p, mv, n);
        if (!isfinite(vmv)) { rep->terminationtype = -4; goto done; }
        if (vmv <= 0.0) { rep->terminationtype = -5; goto done; }
        double rz = dot_seq(r, z, n);
        double alpha = rz / vmv;
        if (!isfinite(alpha)) { rep->terminationtype = -4; goto done; }
        for (int64_t i = 0; i < n; i++) cx[i] = x[i] + alpha * p[i];
        int kk = k - off;
        double merit_k = NAN;
        if (rupd == 0 || kk % rupd != 0) {
            for (int64_t i = 0; i < n; i++) cr[i] = r[i] - alpha * mv[i];
        } else {
            SPMV(cx, mv);
            for (int64_t i = 0; i < n; i++) cr[i] = b[i] - mv[i];
            double v1 = 0.0, v2 = 0.0;
            for (int64_t i = 0; i < n; i++) { v1 += mv[i] * cx[i]; v2 += 2 * b[i] * cx[i]; }
            v1 = v1 - v2;
            merit_k = v1;
            if (o->merit_check && !(v1 < merit)) {         /* rounding stagnation: keep previous x */
                rep->terminationtype = 7;
                rep->iterationscount = k;
                HIST(k, rep->r2, alpha, 0.0, v1);
                goto done;
            }
            merit = v1;
        }
        double cr2 = dot_seq(cr, cr, n);
        HIST(k, cr2, alpha, 0.0, merit_k);
        for (int64_t i = 0; i < n; i++) x[i] = cx[i];
        rep->iterationscount = k;
        rep->r2 = cr2;
        if (sqrt(cr2) <= epsf * bnorm) { rep->terminationtype = 1; goto done; }
        if (maxits > 0 && k >= maxits) { rep->terminationtype = 5; goto done; }
        /* z_new = M^-1 r_new; beta = (r_new . z_new) / (r . z) */
        double beta_num = 0.0;
        if (kk % restart != 0) {
            if (!g_dot_mode) {
                for (int64_t i = 0; i < n; i++) { double czi = cr[i] * d2[i]; beta_num += czi * cr[i]; }
            } else {
                for (int64_t i0 = 0; i0 < n; i0 += 1024) {
                    double t = 0.0;
                    for (int64_t i = i0; i < n && i < i0 + 1024; i++) { double czi = cr[i] * d2[i]; t += czi * cr[i]; }
                    beta_num += t;
                }
            }
            double uvar = rz;
            if (!isfinite(uvar) || uvar == 0.0 || !isfinite(beta_num)) { rep->terminationtype = -4; goto done; }
            double beta = beta_num / uvar;
            HIST(k, cr2, alpha, beta, merit_k);
            for (int64_t i = 0; i < n; i++) { double czi = cr[i] * d2[i]; p[i] = czi + beta * p[i]; z[i] = czi; r[i] = cr[i]; }
        } else {
            for (int64_t i = 0; i < n; i++) { double czi = cr[i] * d2[i]; p[i] = czi; z[i] = czi; r[i] = cr[i]; }
        }
    }
done:
    if (F) { free(F->rowptr); free(F->col); free(F->val); free(F); }
    free(d2); free(r); free(z); free(p); free(mv); free(cx); free(cr);
    g_dot_mode = 0;
    return rc;
#undef SPMV
#undef HIST
}

OX int stan_oracle_lincg(const oracle_csr *A, const double *b, const oracle_cg_opts *o, double *x,
                         oracle_cg_report *rep) {
    return stan_oracle_lincg_hist(A, b, o, x, rep, NULL, 0);
}

/* symmetric product exposed for tests */
OX void stan_oracle_sym_spmv(const oracle_csr *A, const double *x, double *y) { sym_spmv_upper(A, x, y); }

/* ------------------------------------------------------------------------------------ */
/* R3b: LinearSolver_Cholesky (SolverFunctions.cs:332-444): sparseconverttosks +           */
/* sparsecholeskyskyline(isupper=true) + sparsecholeskysolvesks of alglib.net 3.16.0       */
/* ------------------------------------------------------------------------------------ */

/* Skyline (SKS) storage of the upper triangle: column j holds rows first[j]..j, where first[j]
 * is the smallest row with a stored entry in that column (SolverFunctions.cs:385 converts the
 * CRS matrix; fill-in stays inside this envelope).  Factorisation A = U^T U by the bordering
 * scheme ALGLIB documents for sparsecholeskyskyline ("in-place ... best on low-profile
 * matrices", quoted at SolverFunctions.cs:336-381): for column j, rows ascending,
 *     u_ij = (a_ij - sum_{k=max(first[i],first[j])}^{i-1} u_ki u_kj) / u_ii,
 *     u_jj = sqrt(a_jj - sum_k u_kj^2), and the routine reports failure when the radicand <= 0.
 * The solve follows sparsecholeskysolvesks (SolverFunctions.cs:398-427): U^T y = b by column
 * dot products, then U x = y by column sweeps from the last column; terminationtype 1 on
 * success, -3 with x = 0 on failure.  Sum orders (k ascending, dot then subtract) are the
 * published algorithm's natural ones; like the rest of the ALGLIB restatement they are unpinned. */
OX int stan_oracle_cholesky_skyline(const oracle_csr *A, const double *b, double *x, int64_t *envelope_out) {
    int64_t n = A->n;
    int64_t *first = malloc((size_t)(n + 1) * sizeof(int64_t));
    int64_t *cp = malloc((size_t)(n + 1) * sizeof(int64_t));
    for (int64_t j = 0; j < n; j++) first[j] = j;
    for (int64_t i = 0; i < n; i++)
        for (int64_t t = A->rowptr[i]; t < A->rowptr[i + 1]; t++) {
            int64_t j = A->col[t];
            if (j >= i && i < first[j]) first[j] = i;
        }
    cp[0] = 0;
    for (int64_t j = 0; j < n; j++) cp[j + 1] = cp[j] + (j - first[j] + 1);
    if (envelope_out) *envelope_out = cp[n];
    double *S = calloc((size_t)cp[n] + 1, sizeof(double));
    for (int64_t i = 0; i < n; i++)
        for (int64_t t = A->rowptr[i]; t < A->rowptr[i + 1]; t++) {
            int64_t j = A->col[t];
            if (j >= i) S[cp[j] + (i - first[j])] = A->val[t];
        }
#define SK(i, j) S[cp[j] + ((i) - first[j])]
    int ok = 1;
    for (int64_t j = 0; j < n && ok; j++) {
        int64_t fj = first[j];
        for (int64_t i = fj; i < j; i++) {
            int64_t k0 = first[i] > fj ? first[i] : fj;
            double v = 0.0;
            for (int64_t k = k0; k < i; k++) v += SK(k, i) * SK(k, j);
            SK(i, j) = (SK(i, j) - v) / SK(i, i);
        }
        double d = 0.0;
        for (int64_t k = fj; k < j; k++) d += SK(k, j) * SK(k, j);
        double r = SK(j, j) - d;
        if (!(r > 0.0)) { ok = 0; break; }
        SK(j, j) = sqrt(r);
    }
    if (!ok) {
        for (int64_t i = 0; i < n; i++) x[i] = 0.0;
        free(S); free(first); free(cp);
        return -3;
    }
    for (int64_t j = 0; j < n; j++) {                     /* U^T y = b */
        double v = 0.0;
        for (int64_t k = first[j]; k < j; k++) v += SK(k, j) * x[k];
        x[j] = (b[j] - v) / SK(j, j);
    }
    for (int64_t j = n - 1; j >= 0; j--) {                /* U x = y */
        x[j] = x[j] / SK(j, j);
        double xj = x[j];
        for (int64_t k = first[j]; k < j; k++) x[k] -= SK(k, j) * xj;
    }
#undef SK
    free(S); free(first); free(cp);
    return 1;
}

/* ------------------------------------------------------------------------------------ */
/* R4: Recovery_Stress + Update_StrainStress (Element.cs:211-246, 257-267)                  */
/* ------------------------------------------------------------------------------------ */

/* U_full[3*n_nodes] in DOF order.  strain/stress: n_elem x 8 x 6 row-major (Strain[1], Stress[1]).
 * G1: the reference indexes N[i][g] with i = 0..7 on a one-row table (FE_Library.cs:77-81 vs
 * Element.cs:242) and throws; the evident intent N[g][i] = 1 is used (every node receives the
 * single Gauss-point value).  SURVEY.md §8a R4. */
OX int stan_oracle_recover(int n_nodes, const double *xyz, int n_elem, const int32_t *conn,
                           const uint8_t *elem_type, const int32_t *elem_mat, int n_mat, const double *E,
                           const double *nu, const int32_t *node_index, const double *U_full,
                           double *strain, double *stress) {
    (void)n_nodes;
    double *Dm = malloc((size_t)n_mat * 36 * sizeof(double));
    for (int i = 0; i < n_mat; i++) stan_oracle_elastic_D(E[i], nu[i], Dm + 36 * i);
    double tab_dN[3][8 * 24], tab_N[3][64], tab_w[3];
    int tab_ngp[3];
    for (int t = 1; t <= 2; t++) stan_oracle_hex8_tables(t, &tab_ngp[t], &tab_w[t], tab_dN[t], tab_N[t]);
    int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
    for (int e = 0; e < n_elem; e++) {
        int type = elem_type[e];
        if (type != 1 && type != 2) { bad |= 2; continue; }
        int ngp = tab_ngp[type];
        const double *D = Dm + 36 * elem_mat[e];
        double X[24], dU[24], K[576], BLg[8 * 144];
        for (int k = 0; k < 8; k++) {
            int nd = conn[8 * (int64_t)e + k];
            for (int c = 0; c < 3; c++) {
                X[3 * k + c] = xyz[3 * (int64_t)nd + c];
                dU[3 * k + c] = U_full[3 * (int64_t)node_index[nd] + c];   /* Element.cs:214-221 */
            }
        }
        /* the reference reuses BL[g] cached by K_Initial; recompute with identical arithmetic */
        if (stan_oracle_k_initial(type, X, D, K, NULL, BLg)) { bad |= 1; continue; }
        double dEg[8][6], dSg[8][6];
        for (int g = 0; g < ngp; g++) {
            for (int i = 0; i < 6; i++) {                 /* BL[g].MultiplyVector(dU) :230 */
                double c = 0.0;
                for (int j = 0; j < 24; j++) c += BLg[g * 144 + i * 24 + j] * dU[j];
                dEg[g][i] = c;
            }
            for (int i = 0; i < 6; i++) {                 /* D.MultiplyVector(dE) :231 */
                double c = 0.0;
                for (int j = 0; j < 6; j++) c += D[i * 6 + j] * dEg[g][j];
                dSg[g][i] = c;
            }
        }
        for (int i = 0; i < 8; i++) {                     /* extrapolation :238-245 */
            double dE[6] = {0}, dS[6] = {0};
            for (int g = 0; g < ngp; g++) {
                double Nig = (type == 2) ? tab_N[2][i * 8 + g] : 1.0;
                for (int c = 0; c < 6; c++) {
                    dE[c] = 0.0 + (dE[c] + dEg[g][c] * Nig);
                    dS[c] = 0.0 + (dS[c] + dSg[g][c] * Nig);
                }
            }
            for (int c = 0; c < 6; c++) {                 /* Update_StrainStress :257-267 */
                strain[(int64_t)e * 48 + i * 6 + c] = dE[c];
                stress[(int64_t)e * 48 + i * 6 + c] = dS[c];
            }
        }
    }
    free(Dm);
    return bad ? ((bad & 1) ? -1 : -2) : 0;
}

/* ------------------------------------------------------------------------------------ */
/* Whole path (Solver.cs:97-210) with wall-clock breakdown — CPU baseline leg of bench.py   */
/* ------------------------------------------------------------------------------------ */

static double now_s(void) {
#ifdef _OPENMP
    return omp_get_wtime();
#else
    struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec;
#endif
}

typedef struct {
    double t_assign_dof, t_assembly, t_solve, t_recovery, t_total;
    int64_t n_free, nnz_upper;
    int32_t threads, pad;
    oracle_cg_report cg;
} oracle_path_stats;

OX int stan_oracle_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* torchrun exports OMP_NUM_THREADS=1 to every worker; the CPU baseline legs set the team size explicitly. */
OX void stan_oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

OX int stan_oracle_linear_statics(int n_nodes, const double *xyz, int n_elem, const int32_t *conn,
                                  const uint8_t *elem_type, const int32_t *elem_mat, int n_mat, const double *E,
                                  const double *nu, int n_spc, const int32_t *spc_node, const double *spc_val,
                                  int n_load, const int32_t *load_node, const double *load_val,
                                  const oracle_cg_opts *opts, int32_t *node_index_out, double *U_full,
                                  double *strain, double *stress, oracle_path_stats *st) {
    memset(st, 0, sizeof *st);
    st->threads = stan_oracle_threads();
    double t0 = now_s();
    int rc = stan_oracle_assign_dof(n_nodes, n_elem, conn, node_index_out);
    if (rc) return rc;
    double t1 = now_s();
    st->t_assign_dof = t1 - t0;
    int ndof = 3 * n_nodes;
    int32_t *red = malloc((size_t)ndof * sizeof(int32_t));
    int nfix = stan_oracle_spc_reduction(n_nodes, node_index_out, n_spc, spc_node, spc_val, red);
    double *F = malloc((size_t)(ndof - nfix + 1) * sizeof(double));
    stan_oracle_build_rhs(n_nodes, node_index_out, red, n_load, load_node, load_val, F);
    int err;
    oracle_csr *K = stan_oracle_assemble_upper(n_nodes, xyz, n_elem, conn, elem_type, elem_mat, n_mat, E, nu,
                                               node_index_out, red, 0, NULL, &err);
    if (!K) { free(red); free(F); return err; }
    double t2 = now_s();
    st->t_assembly = t2 - t1;
    st->n_free = K->n; st->nnz_upper = K->nnz;
    double *U = malloc((size_t)(K->n + 1) * sizeof(double));
    stan_oracle_lincg(K, F, opts, U, &st->cg);
    double t3 = now_s();
    st->t_solve = t3 - t2;
    stan_oracle_include_bc_dof(ndof, red, U, U_full);
    rc = stan_oracle_recover(n_nodes, xyz, n_elem, conn, elem_type, elem_mat, n_mat, E, nu, node_index_out,
                             U_full, strain, stress);
    double t4 = now_s();
    st->t_recovery = t4 - t3;
    st->t_total = t4 - t0;
    stan_oracle_csr_free(K);
    free(U); free(F); free(red);
    return rc;
}
