// Shared declarations of libstan_b200 (sm_100a).  See include/stan_b200.h for the ABI and
// DESIGN.md for the data layout.  Nothing here depends on PyTorch.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/stan_b200.h"

namespace stan {

void set_error(const char *fmt, ...);

#define STAN_CUDA(call)                                                                              \
    do {                                                                                             \
        cudaError_t e__ = (call);                                                                    \
        if (e__ != cudaSuccess) {                                                                    \
            stan::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));    \
            return STAN_E_CUDA;                                                                      \
        }                                                                                            \
    } while (0)

#define STAN_TRY(call)                   \
    do {                                 \
        int rc__ = (call);               \
        if (rc__ != STAN_OK) return rc__; \
    } while (0)

// Device buffer owned by the handle; stream-ordered allocation from the device's default pool so
// repeated assemble/solve cycles reuse memory instead of paying cudaMalloc every step.
template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    int alloc(size_t count, cudaStream_t s) {
        if (count <= n && p) return STAN_OK;
        release(s);
        if (count == 0) count = 1;
        cudaError_t e = cudaMallocAsync((void **)&p, count * sizeof(T), s);
        if (e != cudaSuccess) {
            p = nullptr;
            n = 0;
            set_error("cudaMallocAsync(%zu bytes): %s", count * sizeof(T), cudaGetErrorString(e));
            return STAN_E_CUDA;
        }
        n = count;
        return STAN_OK;
    }
    void release(cudaStream_t s) {
        if (p) cudaFreeAsync(p, s);
        p = nullptr;
        n = 0;
    }
};

// Grow-only scratch slots owned by the handle.  Temporaries of assemble/solve live here so that the
// steady state performs no device allocation at all (one observed 600 ms stall in a stream-ordered
// allocation in the middle of a run was enough reason).  Same interface as DevBuf; release() is a no-op.
struct ScratchSlot { void *p = nullptr; size_t bytes = 0; };
template <typename T>
struct ScratchBuf {
    T *p = nullptr;
    size_t n = 0;
    ScratchSlot *slot;
    explicit ScratchBuf(ScratchSlot *s) : slot(s) {}
    int alloc(size_t count, cudaStream_t s) {
        const size_t need = (count ? count : 1) * sizeof(T);
        if (need > slot->bytes) {
            if (slot->p) cudaFreeAsync(slot->p, s);
            const size_t cap = need + need / 4 + 256;
            cudaError_t e = cudaMallocAsync(&slot->p, cap, s);
            if (e != cudaSuccess) {
                slot->p = nullptr; slot->bytes = 0;
                set_error("cudaMallocAsync(%zu bytes): %s", cap, cudaGetErrorString(e));
                return STAN_E_CUDA;
            }
            slot->bytes = cap;
        }
        p = (T *)slot->p;
        n = count;
        return STAN_OK;
    }
    void release(cudaStream_t) {}
};

// Peer-memory window of one rank (multi-GPU, comm.cu).  Every rank cudaMalloc's one window, the
// ranks exchange CUDA IPC handles once per assemble, and from then on halo values and partial sums
// travel as plain stores into the neighbour's window over NVLink — no collective library in the CG
// loop.  Control block first, then the three CG vectors an SpMV can be applied to (p, x, xalt), each
// laid out as [owned rows padded to 16 nodes | halo rows]: a peer stores its boundary entries straight
// into the halo tail of the vector the product will read — there is no landing zone and no copy.
constexpr int P2P_MAX_RANKS = 32;
constexpr int HALO_ALIGN = 16;                    // nodes: 16 * 3 * 8 B = 3 cache lines, so no 128-byte line
                                                  // holds both owned entries and halo entries of a vector
struct P2PCtrl {
    unsigned long long hflag[P2P_MAX_RANKS];      // halo arrival per source rank: (solve epoch << 32) | exchange count
    unsigned long long red[2][P2P_MAX_RANKS][4][2];   // partial sums per source rank as two flag-carrying words each
};
struct CommDev {                                   // device-resident view used by kernels
    int rank, world;
    int n_recv_peers;
    int recv_peer[P2P_MAX_RANKS];                 // ranks that own halo rows of mine
    P2PCtrl *ctrl[P2P_MAX_RANKS];                 // ctrl[r]: window of rank r (own or IPC-mapped)
    double *vec[P2P_MAX_RANKS][3];                // the window vectors of rank r: 0 = p, 1 = x, 2 = xalt
    long long tail_off[P2P_MAX_RANKS];            // doubles from the start of rank r's vectors to where my rows land
    long long send_off[P2P_MAX_RANKS + 1];        // my send list grouped by destination rank (nodes)
    long long recv_cnt[P2P_MAX_RANKS];            // halo nodes I receive from rank r
    const int32_t *send_rows;                     // local row of every node to send, grouped by destination
    const int32_t *send_map;                      // per local row: -1, or c with send_dst[send_ptr[c] .. send_ptr[c+1]) its destinations
    const int32_t *send_ptr;
    const unsigned long long *send_dst;           // (peer << 48) | offset in doubles inside the peer's vectors
    unsigned int *ticket;                         // last-CTA detection of the push
    int *err;                                      // device error flags of the handle
};

// What the fused SpMV needs to exchange the halo of its input vector itself (cg.cu: k_spmv_tile3<.., true>).
struct HaloArgs {
    const CommDev *cd;
    const int32_t *tile_order;                    // tiles without halo columns first
    long long n_interior;                         // how many of them
    int vec_id;                                   // which window vector the product reads
};

// Device-resident CG scalars: every decision ALGLIB's lincgiteration takes on the host is taken
// here by one thread, so a batch of iterations can be enqueued without a host round trip.
struct CgState {
    double bnorm, epsf_bnorm;
    double rz;        // r.z of the current residual
    double vmv;       // p.(A p)
    double alpha, beta;
    double r2;        // ||r||^2 of the accepted iterate
    double merit;     // best x'Ax - 2b'x so far
    double partial[4];// all-reduced partial sums of the running phase
    int32_t k;        // completed iterations
    int32_t done;     // 0 = running
    int32_t type;     // terminationtype once done
    int32_t nmv;
    int32_t maxits, rupdate, merit_check, counter_off;
    int64_t restart;
    int32_t x_in_alt; // 1 when the accepted iterate lives in the alternate x buffer
    int32_t x_pending;// 1 when the solve ended at an ordinary iteration whose x += alpha p is still owed
    CommDev *comm;    // peer-memory reductions (multi-GPU P2P mode), else nullptr
    unsigned long long red_seq;   // reductions published so far (same on every rank)
    unsigned long long halo_seq;  // (solve epoch << 32) | halo exchanges completed in this solve (same on every rank)
    double *hist;     // stan_set_cg_history: rows k = 1.. of {||r_k||^2, alpha_k, beta_k, merit or NaN}, else nullptr
    int32_t hist_cap;
    int32_t trace_from;       // first iteration to trace
    unsigned long long *trace;    // STAN_CG_TRACE: (globaltimer ns, event code) pairs, else nullptr
    int32_t trace_cap;
    int32_t trace_n;
};

// Device-side timeline of the CG loop (profiling aid, STAN_CG_TRACE=<events>): one thread per kernel stamps
// %globaltimer at the points that matter for the multi-GPU iteration.  tools/cg_timeline.py reads the dump.
enum TraceCode { TR_SPMV_BEGIN = 1, TR_SPMV_LOCAL_DONE = 2, TR_SPMV_END = 3, TR_UPDATE_BEGIN = 4, TR_UPDATE_LOCAL_DONE = 5,
                 TR_UPDATE_END = 6, TR_DIRECTION_BEGIN = 7, TR_PUSH_BEGIN = 8, TR_PUSH_FLAGS = 9, TR_WAIT_BEGIN = 10,
                 TR_WAIT_END = 11, TR_REFRESH_BEGIN = 12, TR_REFRESH_END = 13 };
#ifdef __CUDACC__
__device__ __forceinline__ void trace_mark(CgState *st, int code) {
    if (!st || !st->trace || st->k < st->trace_from) return;
    const int i = atomicAdd(&st->trace_n, 1);
    if (i < st->trace_cap) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        st->trace[2 * i] = t;
        st->trace[2 * i + 1] = ((unsigned long long)st->k << 8) | (unsigned long long)code;
    }
}
#endif

struct Comm;  // comm.cu

}  // namespace stan

struct stan_handle {
    int device = 0;
    int rank = 0, world = 1;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
    cudaEvent_t user_ev[8] = {};
    std::vector<cudaEvent_t> ev_pool;       // per-launch timing events (time_kernels), reused across solves
    void *stage[2] = {nullptr, nullptr};   // pinned staging for large device-to-host results
    cudaEvent_t stage_ev[2] = {};

    // ---- model (global, every rank holds the whole mesh) ----
    int64_t n_nodes = 0, n_elem = 0, n_elem_g2 = 0;
    int32_t n_mat = 0, max_mat_index = 0;
    bool have_mesh = false, have_mat = false, have_dof = false, assembled = false, solved = false,
         recovered = false;
    std::vector<int32_t> h_conn;        // host copy of the connectivity, fetched back only if the host AssignDOF runs
    stan::DevBuf<double> d_xyz;         // 3*n_nodes
    stan::DevBuf<int32_t> d_conn;       // 8*n_elem
    stan::DevBuf<uint8_t> d_etype;      // n_elem
    stan::DevBuf<int32_t> d_emat;       // n_elem
    stan::DevBuf<double> d_lambda, d_G; // n_mat (Material.cs:39-40)
    stan::DevBuf<int32_t> d_node_index; // n_nodes: BFS index of node (DOF/3)
    stan::DevBuf<int32_t> d_inv;        // n_nodes: node at BFS index
    std::vector<int32_t> h_spc_node;
    std::vector<double> h_spc_val;
    std::vector<int32_t> h_load_node;
    std::vector<double> h_load_val;

    // ---- partition ----
    int64_t row0 = 0, row1 = 0;         // owned BFS rows [row0, row1)
    std::vector<int64_t> bounds;        // world + 1 row bounds, the same on every rank (comm.cu: partition_rows)
    int64_t n_halo = 0;                 // halo nodes appended after the owned rows in x vectors
    int64_t nloc_pad = 0;               // node index where the halo tail starts: nloc (1 GPU) or nloc rounded up to HALO_ALIGN
    int64_t elem0 = 0, elem1 = 0;       // elements whose strain/stress this rank recovers

    // ---- assembled system (local rows) ----
    stan::DevBuf<int32_t> d_inc_ptr;    // nloc+1: incidence CSR (row -> (elem<<3 | local node))
    stan::DevBuf<int32_t> d_inc;        //
    stan::DevBuf<int32_t> d_brow_ptr;   // nloc+1: block-row pointers
    stan::DevBuf<int32_t> d_bcol;       // global BFS column node of every block
    stan::DevBuf<int32_t> d_bcol_loc;   // local x index of every block (owned: q-row0, halo: nloc+slot)
    const int32_t *bcol_x = nullptr;    // what SpMV indexes x with: d_bcol (1 GPU) or d_bcol_loc
    stan::DevBuf<double> d_ke;          // 36 upper 3x3 blocks per local element (assembly scratch, kept in the pool)
    stan::DevBuf<double> d_vals;        // 9 per block; per block row: 3 scalar rows of length 3*nb
    stan::DevBuf<uint8_t> d_fixed;      // 3*n_nodes (global): 1 = SPC-fixed DOF
    stan::DevBuf<int32_t> d_red;        // 3*n_nodes: nDOF_reduction (Solver.cs:121-132)
    stan::DevBuf<double> d_b;           // 3*nloc RHS in full DOF space (0 at fixed)
    stan::DevBuf<double> d_d2;          // 3*nloc: (1/sqrt(A_ii))^2
    stan::DevBuf<int32_t> d_err;        // device error flags
    int64_t n_blocks = 0, n_fixed = 0, nnz_upper = 0;
    int32_t max_group_blocks = 0;       // max blocks in a 32-row group (assembly smem sizing)
    int32_t max_group16 = 0;            // max blocks in a 16-row group (SpMV stage sizing)

    // ---- CG work vectors (3*(nloc+n_halo) where halo is needed) ----
    stan::DevBuf<double> d_x, d_xalt, d_r, d_p, d_mv, d_partials;
    stan::DevBuf<stan::CgState> d_state;
    stan::DevBuf<unsigned int> d_counter;
    stan::DevBuf<double> d_hist;        // 4 doubles per iteration (stan_set_cg_history)
    stan::DevBuf<unsigned long long> d_trace;   // STAN_CG_TRACE
    int32_t hist_cap = 0, hist_count = 0;
    bool x_in_alt = false;
    double *sol = nullptr;              // accepted solution of the last solve (owned rows), wherever the solver keeps it
    unsigned long long red_seq = 0;     // cross-rank reductions published so far (peer-memory mode)

    // ---- results ----
    stan::DevBuf<double> d_ufull;       // 3*n_nodes in DOF order (all ranks after the gather)
    stan::DevBuf<double> d_strain, d_stress;  // 48*n_elem
    stan::DevBuf<float> d_cell, d_point;      // post-processing scalars: [elem][24][3], [node][24]
    bool postprocessed = false;

    stan::ScratchSlot scratch[12];      // see ScratchBuf: 0-5 pattern/halo/assembly temporaries, 6-7 element maps, 8 scan, 9 checks
    stan::CgState *h_state = nullptr;   // pinned mirror of the device CG state
    stan::Comm *comm = nullptr;
    int64_t launches = 0;
};

namespace stan {

// dofmap.cpp
int assign_dof_host(int64_t n_nodes, int64_t n_elem, const int32_t *conn, int32_t *node_index);

// dofmap_gpu.cu
int assign_dof_device(stan_handle *h, int32_t *node_index, bool *narrow);

// pattern.cu
int build_system_pattern(stan_handle *h);
int build_rhs(stan_handle *h);
int device_exclusive_scan_i32(stan_handle *h, const int32_t *in, int32_t *out, int64_t n, cudaStream_t s);
int export_csr_upper_size(stan_handle *h, int64_t *n, int64_t *nnz);
int export_csr_upper(stan_handle *h, int64_t *rowptr, int32_t *col, double *val);

// assembly.cu
int upload_fe_tables();
void host_fe_tables(double *tab /*9*24*/);
int run_assembly(stan_handle *h);
int element_stiffness(stan_handle *h, int64_t first, int64_t count, double *ke_host);

// cg.cu
int solve_cg(stan_handle *h, const stan_cg_options *o, stan_cg_report *rep);
int spmv_full(stan_handle *h, const double *x_full, double *y_full);
int time_spmv(stan_handle *h, int reps, double *ms, int64_t *bytes);
int64_t spmv_algorithmic_bytes(const stan_handle *h);
int scatter_solution(stan_handle *h);

// cholesky.cu
int solve_cholesky(stan_handle *h, stan_chol_report *rep);

// recovery.cu
int run_recovery(stan_handle *h, stan_recovery_stats *st);
int recover_elements(stan_handle *h, const int32_t *d_list, int64_t count, double *d_strain, double *d_stress);

// postprocess.cu
int run_postprocess(stan_handle *h, double *ms);

// comm.cu
int comm_unique_id(void *id128);
int comm_init(stan_handle *h, const void *id128);
void comm_destroy(stan_handle *h);
int comm_allreduce_sum(stan_handle *h, double *d_buf, int count, cudaStream_t s);
int comm_halo_exchange(stan_handle *h, double *d_vec, int vec_id, cudaStream_t s, CgState *st);
int comm_cg_vectors(stan_handle *h, double **p, double **x, double **xalt, cudaStream_t s);
bool comm_halo_args(const stan_handle *h, int vec_id, HaloArgs *out);
unsigned long long comm_next_epoch(stan_handle *h);
CommDev *comm_dev_ptr(const stan_handle *h);
int comm_allgather_rows(stan_handle *h, const double *d_local, double *d_full, cudaStream_t s);
int comm_build_halo(stan_handle *h);
bool comm_p2p_active(const stan_handle *h);
CommDev *comm_dev(const stan_handle *h);

static inline int div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

}  // namespace stan
