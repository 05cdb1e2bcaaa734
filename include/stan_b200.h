/*
 * stan_b200.h — C ABI of libstan_b200.so, the B200-native replacement for the body of
 * Solver.SolverLinearStatics in galuszkm/STAN (src/STAN_Solver/Solver.cs:97-210).
 *
 * The reference has no FFI: its seams are managed calls.  Each entry point below names the
 * managed code it replaces; INTEGRATION.md shows the P/Invoke stub (interop/StanNative.cs)
 * and the patch to SolverLinearStatics.  Plain pointers and sizes only; every array argument
 * is owned by the caller and is copied before the call returns (P/Invoke pins blittable
 * arrays only for the duration of a call).  All functions return 0 on success or a negative
 * STAN_E_* code; stan_last_error() gives a thread-local message.  No exceptions cross the ABI.
 * cdecl, x86-64, not re-entrant per handle; one handle drives one GPU.
 *
 * Flat model = the reference's object graph in dictionary insertion order
 * (SURVEY.md §8a R6): node i is the i-th entry of Database.NodeLib, element e the e-th entry
 * of Database.ElemLib, connectivity holds 0-based node positions in CHEXA order
 * (src/STAN_Database/FE_Library.cs:108-115).
 */
#ifndef STAN_B200_H
#define STAN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)   /* the library is built with -fvisibility=hidden */
#endif

#define STAN_OK            0
#define STAN_E_ARG        -1  /* bad argument / out-of-range index */
#define STAN_E_CUDA       -2  /* CUDA runtime failure (message has the detail) */
#define STAN_E_SINGULAR   -3  /* Jacobian determinant == 0 (MatrixST.cs:298,317 throws) */
#define STAN_E_STATE      -4  /* call order violated (e.g. solve before assemble) */
/* -5 was STAN_E_CAPACITY (row wider than 96 blocks) in round 1: rows of any width are accepted now */
#define STAN_E_DOFMAP     -6  /* AssignDOF failed: no start node / disconnected mesh (Database.cs:178-196, :218) */
#define STAN_E_COMM       -7  /* NCCL / multi-GPU failure */
#define STAN_E_NOMEM      -8  /* device memory: the direct solver's skyline (or an assembly work area) does not fit */

#define STAN_HEX8_G1 1        /* Element.Type "HEX8_G1", FE_Library.cs:63-89 */
#define STAN_HEX8_G2 2        /* Element.Type "HEX8_G2", FE_Library.cs:91-131 */

typedef struct stan_handle stan_handle;

typedef struct {
    int32_t device;            /* CUDA ordinal; -1 = current device */
    int32_t rank;              /* this process's partition, 0 <= rank < world */
    int32_t world;             /* number of partitions (GPUs); 1 = single GPU */
    int32_t flags;             /* reserved, 0 */
} stan_options;

/* Mirrors Analysis.LinSolverTolerance / LinSolverIterMax (Analysis.cs:10-11) plus the ALGLIB
 * lincg internals that LinearSolver_CG leaves at their defaults (SolverFunctions.cs:270-330). */
typedef struct {
    double  epsf;              /* lincgsetcond EpsF: stop when ||r||2 <= epsf*||b||2; (0,0) -> 1e-6 */
    int32_t maxits;            /* lincgsetcond MaxIts; 0 = unlimited */
    int32_t its_before_rupdate;/* true-residual refresh period, ALGLIB default 10; 0 = never */
    int32_t its_before_restart;/* direction restart period, ALGLIB default n (pass 0) */
    int32_t merit_check;       /* 1 = ALGLIB: terminationtype 7 when the energy functional stalls */
    int32_t zero_based_counter;/* 0 = period tests on k = 1,2,..; 1 = on k-1 */
    int32_t time_kernels;      /* 1 = bracket every SpMV launch with CUDA events (no graph replay) */
    int32_t reserved;
} stan_cg_options;

/* alglib.lincgreport as consumed at SolverFunctions.cs:305-327, plus device timings. */
typedef struct {
    int32_t terminationtype;   /* 1, 5, 7, -4, -5 (SolverFunctions.cs:311-319) */
    int32_t iterationscount;
    int32_t nmv;               /* matrix-vector products counted as ALGLIB does (incl. r0 = b - A*0) */
    int32_t spmv_launches;     /* SpMV kernels actually launched by this rank */
    double  r2;                /* squared 2-norm of the final residual */
    double  bnorm;             /* ||b||2 */
    double  solve_ms;          /* device time of the whole solve (CUDA events) */
    double  spmv_ms;           /* summed device time of the SpMV launches (time_kernels = 1), else 0 */
    int64_t spmv_bytes;        /* algorithmic bytes one SpMV launch moves on this rank */
    int64_t iter_bytes;        /* algorithmic bytes of one full CG iteration on this rank */
    int64_t kernel_launches;   /* all kernels launched by the solve */
} stan_cg_report;

typedef struct {
    int64_t n_dof;             /* 3 * n_nodes (Database.nDOF) */
    int64_t n_fixed;           /* |Distinct(Fix_DOF)| (Solver.cs:117) */
    int64_t n_rows_local;      /* block rows (nodes) owned by this rank */
    int64_t n_blocks_local;    /* stored 3x3 blocks on this rank */
    int64_t nnz_upper;         /* entries of the reference's upper-triangle CRS; counted lazily: 0 until a
                                  stan_get_csr_upper_size call on this model, its result afterwards */
    int64_t assembly_bytes;    /* algorithmic bytes of the assembly kernel */
    double  assembly_flops;    /* algorithmic flops of the element integration: 17.3 k per HEX8_G2
                                  element + 2.2 k per HEX8_G1 element */
    double  pattern_ms;        /* incidence + block pattern build */
    double  assembly_ms;       /* hex8 integration + deterministic row assembly kernel */
    double  total_ms;          /* whole stan_assemble call on the device */
    int64_t kernel_launches;
} stan_assembly_stats;

/* What LinearSolver_Cholesky prints (SolverFunctions.cs:386-437) plus device figures. */
typedef struct {
    int32_t terminationtype;   /* sparsecholeskysolvesks: 1 = solved, -3 = not positive definite (U = 0) */
    int32_t block;             /* edge of the dense blocks the skyline is stored in (64) */
    int64_t n;                 /* rows factorised: nDOF, fixed DOFs kept as identity rows */
    int64_t n_blocks;          /* dense blocks inside the block skyline */
    int64_t skyline_bytes;     /* device bytes of the factor */
    double  flops;             /* flops of the factorisation as executed (dense blocks) */
    double  setup_ms;          /* envelope + CRS -> skyline (sparseconverttosks) */
    double  factor_ms;         /* sparsecholeskyskyline, with U^T y = b carried along as one more column */
    double  solve_ms;          /* sparsecholeskysolvesks: the remaining back substitution U x = y */
    int64_t kernel_launches;
} stan_chol_report;

typedef struct {
    double  recover_ms;
    int64_t recover_bytes;     /* algorithmic bytes */
    int64_t kernel_launches;
} stan_recovery_stats;

const char *stan_last_error(void);
int stan_version(void);

int stan_create(const stan_options *opts, stan_handle **out);
int stan_destroy(stan_handle *h);

/* --- model upload: replaces the object-graph reads of Solver.cs:81-152 --------------------- */
/* Node.X/Y/Z (Node.cs:12-14), Element.NList/Type/MatID (Element.cs:15-18). */
int stan_set_mesh(stan_handle *h, int64_t n_nodes, const double *xyz, int64_t n_elem, const int32_t *conn,
                  const uint8_t *elem_type, const int32_t *elem_mat);
/* Material.SetElastic(E, Poisson) for every "Elastic" material (Solver.cs:33-39, Material.cs:31-56). */
int stan_set_materials(stan_handle *h, int32_t n_mat, const double *E, const double *nu);
/* Node.DOF[0]/3 per node when the managed side already ran Database.AssignDOF (Solver.cs:46). */
int stan_set_dof_map(stan_handle *h, const int32_t *node_index);
/* Native Database.AssignDOF (Database.cs:140-234): computes, stores and returns the BFS index. */
int stan_assign_dof(stan_handle *h, int32_t *node_index_out);
/* BCLib entries of Type "SPC": a DOF is fixed iff its value == 1 (Solver.cs:106-114). */
int stan_set_spc(stan_handle *h, int64_t n, const int32_t *node, const double *val3);
/* BCLib entries of Type "PointLoad", accumulated in list order (Solver.cs:136-152). */
int stan_set_loads(stan_handle *h, int64_t n, const int32_t *node, const double *fxyz);

/* --- the hot path ------------------------------------------------------------------------- */
/* Fun.ParallelAssembly_K(DB, nDOF_reduction, 1, "Initial") + nDOF_reduction + F
 * (Solver.cs:104-156, SolverFunctions.cs:117-180, Element.cs:118-155). */
int stan_assemble(stan_handle *h, stan_assembly_stats *stats);
/* Fun.LinearSolver_CG(K, F, AnalysisLib) (Solver.cs:162, SolverFunctions.cs:270-330). */
int stan_solve_cg(stan_handle *h, const stan_cg_options *opts, stan_cg_report *report);
/* Fun.LinearSolver_Cholesky(K, F) (Solver.cs:163, SolverFunctions.cs:332-444): skyline U^T U
 * factorisation and two triangular solves on one GPU; leaves U where stan_solve_cg leaves it. */
int stan_solve_cholesky(stan_handle *h, stan_chol_report *report);
/* Include_BC_DOF + dU_buffer + Elem.Recovery_Stress + Update_StrainStress
 * (Solver.cs:168-210, Element.cs:211-246, 257-267). */
int stan_recover(stan_handle *h, stan_recovery_stats *stats);

/* --- results: what Solver.cs:171-178,203-210 writes back into Node / Element --------------- */
/* U_Full[nDOF] indexed by DOF (zeros at fixed DOFs). */
int stan_get_displacements(stan_handle *h, double *u_full);
/* The rows this rank owns, [first_row, last_row) of stan_get_partition, 3 doubles per row: what a caller that
 * merges the ranks' results itself downloads instead of the whole vector (same as stan_get_displacements on one GPU). */
int stan_get_displacements_local(stan_handle *h, double *u_rows);
/* Node.dU_buffer for every node in NodeLib order (Solver.cs:171-178: dU_buffer[i] = U_Full[DOF[i]]), n_nodes x 3;
 * the gather through the DOF map runs on the device. */
int stan_get_node_displacements(stan_handle *h, double *disp);
/* Element.Strain[1] / Stress[1]: 8 x 6 row-major per element, for the elements [first, last) of
 * stan_get_element_range — all of ElemLib on one GPU, this rank's contiguous slice otherwise. */
int stan_get_strain_stress(stan_handle *h, double *strain, double *stress);
int stan_get_element_range(stan_handle *h, int64_t *first, int64_t *last);

/* --- post-processing: Part.Load_Scalar, Part.cs:231-528 ------------------ */
/* 24 scalar fields (displacement x/y/z/total, stress xx..xz, P1..P3, von Mises, strain xx..xz,
 * P1..P3, effective strain) as float32: cell = n_elem x 24 x {max, average, min} over the element's
 * nodes, point = n_nodes x 24 averaged over the elements containing the node (NodeLib order).
 * Multi-GPU: cell covers this rank's element slice (stan_get_element_range), point covers its rows
 * (stan_get_partition) in DOF-map order: point[r - first_row] belongs to the node with DOF[0]/3 == r. */
int stan_postprocess(stan_handle *h, double *device_ms);
int stan_get_scalars(stan_handle *h, float *cell, float *point);

/* Page-locked host memory for callers that want the copies at the boundary to be plain DMA transfers (any
 * host pointer is accepted everywhere; pageable memory goes through the library's staging buffers). */
int stan_host_alloc(size_t bytes, void **out);
int stan_host_free(void *p);

/* --- parity / inspection (SURVEY §8b "optional") ------------------------------------------- */
int stan_get_dof_reduction(stan_handle *h, int32_t *ndof_reduction);          /* Solver.cs:121-132 */
int stan_get_rhs(stan_handle *h, double *F_reduced);                          /* Solver.cs:136-152 */
int stan_get_solution_reduced(stan_handle *h, double *U_reduced);             /* Exclude_BC_DOF(U_Full) */
/* Reduced upper-triangle CRS as alglib.sparseconverttocrs leaves it (structural pattern). */
int stan_get_csr_upper_size(stan_handle *h, int64_t *n, int64_t *nnz);
int stan_get_csr_upper(stan_handle *h, int64_t *rowptr, int32_t *col, double *val);
/* Element.K_Initial for elements [first, first+count): count x 24 x 24 row-major. */
int stan_element_stiffness(stan_handle *h, int64_t first, int64_t count, double *ke);
/* y = K x on the stored matrix, x and y in full DOF space (fixed DOFs act as identity rows). */
int stan_spmv(stan_handle *h, const double *x_full, double *y_full);
/* Trajectory of the CG recurrences, for pinning R5 iteration by iteration against the CPU restatement of
 * alglib.lincgiteration (SolverFunctions.cs:304): with capacity > 0 the next stan_solve_cg records, for
 * k = 1 .. min(capacity, iterationscount), the row { ||r_k||^2, alpha_k, beta_k (0 on a restart and at the
 * final iteration), energy functional x'Ax - 2b'x on true-residual refresh iterations else NaN }.
 * stan_get_cg_history returns the number of rows and copies count x 4 doubles (hist4 may be NULL). */
int stan_set_cg_history(stan_handle *h, int32_t capacity);
int stan_get_cg_history(stan_handle *h, int32_t *count, double *hist4);
/* Device time of `reps` back-to-back SpMV launches on the assembled matrix (CUDA events). */
int stan_time_spmv(stan_handle *h, int32_t reps, double *ms_per_launch, int64_t *bytes_per_launch);

/* CUDA events on the library's own stream, for callers that time a region on the device:
 * record slot a, run calls, record slot b, then elapsed(a, b) synchronises and returns ms. */
int stan_event_record(stan_handle *h, int32_t slot);          /* slot in [0, 8) */
int stan_event_elapsed(stan_handle *h, int32_t slot_a, int32_t slot_b, double *ms);
/* Kernels launched through this handle so far (bench.py's gpu_launches). */
int64_t stan_kernel_launches(stan_handle *h);

/* --- multi-GPU (one process per GPU) ------------------------------------------- */
/* Rank 0 creates the 128-byte NCCL id, the host side broadcasts it, every rank joins. */
int stan_comm_unique_id(void *id128);
int stan_comm_init(stan_handle *h, const void *id128);
/* Row range [first, last) of BFS nodes owned by this rank after stan_assemble. */
int stan_get_partition(stan_handle *h, int64_t *first_row, int64_t *last_row);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* STAN_B200_H */
