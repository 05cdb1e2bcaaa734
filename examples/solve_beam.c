/*
 * solve_beam.c — the whole linear-static path through the C ABI from plain C99, the way a
 * P/Invoke / cgo / FFI caller drives it: a cantilever of nx*ny*nz unit cubes,
 * clamped at z = 0, loaded in +x at z = nz.
 *
 *   gcc -std=c99 -I include examples/solve_beam.c -o solve_beam -L stan_b200/lib -lstan_b200 \
 *       -Wl,-rpath,$PWD/stan_b200/lib -lm
 *   ./solve_beam [nx ny nz] [cg|cholesky]
 *
 * Prints one JSON line: tip deflection against Timoshenko beam theory, solver report, device times.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "stan_b200.h"

#define CHECK(call)                                                                   \
    do {                                                                              \
        int rc_ = (call);                                                             \
        if (rc_ != STAN_OK) {                                                         \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, stan_last_error());   \
            return 1;                                                                 \
        }                                                                             \
    } while (0)

int main(int argc, char **argv) {
    const int nx = argc > 3 ? atoi(argv[1]) : 4, ny = argc > 3 ? atoi(argv[2]) : 4, nz = argc > 3 ? atoi(argv[3]) : 40;
    const int cholesky = argc > 1 && !strcmp(argv[argc - 1], "cholesky");
    const int nxn = nx + 1, nyn = ny + 1, nzn = nz + 1;
    const int64_t n_nodes = (int64_t)nxn * nyn * nzn, n_elem = (int64_t)nx * ny * nz;

    /* flat model: node index = i + nxn (j + nyn k); CHEXA node order of FE_Library.cs:108-115 */
    double *xyz = malloc(sizeof(double) * 3 * n_nodes);
    int32_t *conn = malloc(sizeof(int32_t) * 8 * n_elem), *emat = calloc(n_elem, sizeof(int32_t));
    uint8_t *etype = malloc(n_elem);
    for (int k = 0; k < nzn; k++)
        for (int j = 0; j < nyn; j++)
            for (int i = 0; i < nxn; i++) {
                const int64_t n = i + (int64_t)nxn * (j + (int64_t)nyn * k);
                xyz[3 * n] = i; xyz[3 * n + 1] = j; xyz[3 * n + 2] = k;
            }
    const int32_t off[8] = {0, 1, 1 + nxn, nxn, nxn * nyn, nxn * nyn + 1, nxn * nyn + 1 + nxn, nxn * nyn + nxn};
    int64_t e = 0;
    for (int k = 0; k < nz; k++)
        for (int j = 0; j < ny; j++)
            for (int i = 0; i < nx; i++, e++) {
                const int32_t base = i + nxn * (j + nyn * k);
                for (int q = 0; q < 8; q++) conn[8 * e + q] = base + off[q];
                etype[e] = STAN_HEX8_G2;
            }
    const int64_t n_face = (int64_t)nxn * nyn;
    int32_t *spc_node = malloc(sizeof(int32_t) * n_face), *load_node = malloc(sizeof(int32_t) * n_face);
    double *spc_val = malloc(sizeof(double) * 3 * n_face), *load_val = calloc(3 * n_face, sizeof(double));
    const double total_load = 1000.0, E = 210000.0, nu = 0.3;
    for (int64_t q = 0; q < n_face; q++) {
        spc_node[q] = (int32_t)q;                                  /* k = 0 face: all three DOFs fixed (value 1) */
        spc_val[3 * q] = spc_val[3 * q + 1] = spc_val[3 * q + 2] = 1.0;
        load_node[q] = (int32_t)(q + n_face * nz);                 /* k = nz face */
        load_val[3 * q] = total_load / (double)n_face;
    }

    stan_options opt = {-1, 0, 1, 0};
    stan_handle *h = NULL;
    CHECK(stan_create(&opt, &h));
    CHECK(stan_set_mesh(h, n_nodes, xyz, n_elem, conn, etype, emat));
    CHECK(stan_set_materials(h, 1, &E, &nu));
    int32_t *node_index = malloc(sizeof(int32_t) * n_nodes);
    CHECK(stan_assign_dof(h, node_index));                         /* Database.AssignDOF */
    CHECK(stan_set_spc(h, n_face, spc_node, spc_val));
    CHECK(stan_set_loads(h, n_face, load_node, load_val));
    stan_assembly_stats as;
    CHECK(stan_assemble(h, &as));                                  /* ParallelAssembly_K + nDOF_reduction + F */
    stan_cg_report cg;
    stan_chol_report ch;
    memset(&cg, 0, sizeof cg);
    memset(&ch, 0, sizeof ch);
    if (cholesky) {
        CHECK(stan_solve_cholesky(h, &ch));                        /* LinearSolver_Cholesky */
    } else {
        stan_cg_options o = {1e-8, 0, 10, 0, 1, 0, 0, 0};          /* ALGLIB defaults, EpsF = 1e-8 */
        CHECK(stan_solve_cg(h, &o, &cg));                          /* LinearSolver_CG */
    }
    stan_recovery_stats rs;
    CHECK(stan_recover(h, &rs));                                   /* Recovery_Stress + Update_StrainStress */
    double *U = malloc(sizeof(double) * 3 * n_nodes);
    double *strain = malloc(sizeof(double) * 48 * n_elem), *stress = malloc(sizeof(double) * 48 * n_elem);
    CHECK(stan_get_displacements(h, U));
    CHECK(stan_get_strain_stress(h, strain, stress));

    double tip = 0.0, smax = 0.0;
    for (int64_t q = 0; q < n_face; q++) tip += U[3 * (int64_t)node_index[load_node[q]]] / (double)n_face;
    for (int64_t q = 0; q < 48 * n_elem; q++) smax = fmax(smax, fabs(stress[q]));
    const double L = nz, A = (double)nx * ny, I = (double)ny * nx * nx * nx / 12.0, G = E / (2 * (1 + nu));
    const double theory = total_load * L * L * L / (3 * E * I) + total_load * L / (5.0 / 6.0 * G * A);
    printf("{\"elements\": %lld, \"dof\": %lld, \"solver\": \"%s\", \"terminationtype\": %d, \"iterations\": %d, "
           "\"tip_ux\": %.9g, \"beam_theory\": %.9g, \"max_abs_stress\": %.6g, \"assemble_ms\": %.3f, \"solve_ms\": %.3f, "
           "\"recover_ms\": %.3f}\n",
           (long long)n_elem, (long long)as.n_dof, cholesky ? "Cholesky" : "CG", cholesky ? ch.terminationtype : cg.terminationtype,
           cg.iterationscount, tip, theory, smax, as.total_ms, cholesky ? ch.setup_ms + ch.factor_ms + ch.solve_ms : cg.solve_ms,
           rs.recover_ms);
    CHECK(stan_destroy(h));
    free(xyz); free(conn); free(emat); free(etype); free(spc_node); free(load_node); free(spc_val); free(load_val);
    free(node_index); free(U); free(strain); free(stress);
    return 0;
}
