// C ABI of libstan_b200.so (include/stan_b200.h): argument checking, call-order state machine,
// host<->device copies at the boundary.  Orchestrates pattern.cu / assembly.cu / cg.cu / recovery.cu
// in the order of Solver.SolverLinearStatics (/root/reference/src/STAN_Solver/Solver.cs:97-210).
#include <chrono>
#include <cmath>
#include <thread>

#include "common.cuh"

namespace stan {

static thread_local char g_err[1024] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

int partition_rows(stan_handle *h);   // comm.cu

static int check(stan_handle *h) {
    if (!h) { set_error("null handle"); return STAN_E_ARG; }
    cudaError_t e = cudaSetDevice(h->device);
    if (e != cudaSuccess) { set_error("cudaSetDevice(%d): %s", h->device, cudaGetErrorString(e)); return STAN_E_CUDA; }
    return STAN_OK;
}

// Large results go to the caller's (pageable, usually untouched) buffer through two pinned staging
// buffers: the DMA engine fills one while several host threads copy the other out — first-touch page
// faults on the destination are what bound a plain cudaMemcpy to ~4 GB/s.
static bool is_pinned(const void *p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

static int copy_to_host(stan_handle *h, void *dst, const void *d_src, size_t bytes) {
    cudaStream_t s = h->stream;
    const size_t CH = (size_t)64 << 20;
    if (bytes < 2 * CH || is_pinned(dst)) {          // page-locked destination (stan_host_alloc): one DMA, no staging
        STAN_CUDA(cudaMemcpyAsync(dst, d_src, bytes, cudaMemcpyDeviceToHost, s));
        STAN_CUDA(cudaStreamSynchronize(s));
        return STAN_OK;
    }
    if (!h->stage[0]) {
        for (int k = 0; k < 2; k++) {
            STAN_CUDA(cudaMallocHost(&h->stage[k], CH));
            STAN_CUDA(cudaEventCreateWithFlags(&h->stage_ev[k], cudaEventDisableTiming));
        }
    }
    const unsigned hw = std::thread::hardware_concurrency();
    const int nt = (int)std::min<unsigned>(8, hw ? hw : 1);
    const size_t nch = (bytes + CH - 1) / CH;
    auto issue = [&](size_t c) -> cudaError_t {
        const size_t off = c * CH, n = std::min(CH, bytes - off);
        cudaError_t e = cudaMemcpyAsync(h->stage[c & 1], (const char *)d_src + off, n, cudaMemcpyDeviceToHost, s);
        return e != cudaSuccess ? e : cudaEventRecord(h->stage_ev[c & 1], s);
    };
    STAN_CUDA(issue(0));
    for (size_t c = 0; c < nch; c++) {
        if (c + 1 < nch) STAN_CUDA(issue(c + 1));            // buffer (c+1)&1 was drained in the previous round
        STAN_CUDA(cudaEventSynchronize(h->stage_ev[c & 1]));
        const size_t off = c * CH, n = std::min(CH, bytes - off);
        std::vector<std::thread> th;
        const size_t per = ((n + nt - 1) / nt + 4095) & ~(size_t)4095;
        for (int t = 0; t < nt; t++) {
            const size_t a = (size_t)t * per;
            if (a >= n) break;
            const size_t len = std::min(per, n - a);
            th.emplace_back([=] { memcpy((char *)dst + off + a, (const char *)h->stage[c & 1] + a, len); });
        }
        for (auto &t : th) t.join();
    }
    return STAN_OK;
}

// The model is checked where it lands: a pass over 80 M connectivity entries costs the host ~0.1 s and the
// device ~0.1 ms.  res[0..2] = first element with a bad node index / type / material index, res[3] = largest
// material index, res[4] = number of HEX8_G2 elements.
__global__ void k_check_mesh(long long n_elem, long long n_nodes, const int32_t *__restrict__ conn,
                             const uint8_t *__restrict__ etype, const int32_t *__restrict__ emat, long long *res) {
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const bool in = e < n_elem;
    bool g2 = false;
    if (in) {
        const int4 c0 = *reinterpret_cast<const int4 *>(conn + 8 * e), c1 = *reinterpret_cast<const int4 *>(conn + 8 * e + 4);
        const int lo = min(min(min(c0.x, c0.y), min(c0.z, c0.w)), min(min(c1.x, c1.y), min(c1.z, c1.w)));
        const int hi = max(max(max(c0.x, c0.y), max(c0.z, c0.w)), max(max(c1.x, c1.y), max(c1.z, c1.w)));
        if (lo < 0 || hi >= n_nodes) atomicMin((unsigned long long *)&res[0], (unsigned long long)e);
        const int t = etype[e];
        if (t != STAN_HEX8_G1 && t != STAN_HEX8_G2) atomicMin((unsigned long long *)&res[1], (unsigned long long)e);
        g2 = t == STAN_HEX8_G2;
        const int m = emat[e];
        if (m < 0) atomicMin((unsigned long long *)&res[2], (unsigned long long)e);
        else if (m > 0) atomicMax((unsigned long long *)&res[3], (unsigned long long)m);
    }
    const unsigned b = __ballot_sync(0xffffffffu, g2);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd((unsigned long long *)&res[4], (unsigned long long)__popc(b));
}

// node_index must be a permutation of 0..n-1: scatter the inverse, then look for holes (a duplicate leaves one)
__global__ void k_perm_scatter(long long n, const int32_t *__restrict__ node_index, int32_t *__restrict__ inv, long long *bad) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t p = node_index[i];
    if (p < 0 || p >= n) atomicMin((unsigned long long *)bad, (unsigned long long)i);
    else inv[p] = (int32_t)i;
}
__global__ void k_perm_holes(long long n, const int32_t *__restrict__ inv, long long *bad) {
    const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (p < n && inv[p] < 0) atomicMin((unsigned long long *)(bad + 1), (unsigned long long)p);
}

__global__ void k_gather_rows3(long long n, const int32_t *__restrict__ node_index, const double *__restrict__ ufull,
                               double *__restrict__ out) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= 3 * n) return;
    const long long i = t / 3;
    out[t] = ufull[3 * (long long)node_index[i] + (t - 3 * i)];
}

// d_node_index holds the candidate map; verifies it on the device and leaves the inverse in d_inv
static int check_dof_map(stan_handle *h) {
    cudaStream_t s = h->stream;
    const int64_t n = h->n_nodes;
    STAN_TRY(h->d_inv.alloc(n, s));
    ScratchBuf<long long> bad(&h->scratch[9]);
    STAN_TRY(bad.alloc(2, s));
    STAN_CUDA(cudaMemsetAsync(h->d_inv.p, 0xff, n * sizeof(int32_t), s));
    STAN_CUDA(cudaMemsetAsync(bad.p, 0x7f, 2 * sizeof(long long), s));
    k_perm_scatter<<<div_up(n, 256), 256, 0, s>>>(n, h->d_node_index.p, h->d_inv.p, bad.p);
    k_perm_holes<<<div_up(n, 256), 256, 0, s>>>(n, h->d_inv.p, bad.p);
    long long hb[2];
    STAN_CUDA(cudaMemcpyAsync(hb, bad.p, sizeof hb, cudaMemcpyDeviceToHost, s));
    STAN_CUDA(cudaStreamSynchronize(s));
    h->launches += 2;
    if (hb[0] < n) { set_error("dof map is not a permutation (node %lld maps outside [0,%lld))", hb[0], (long long)n); return STAN_E_ARG; }
    if (hb[1] < n) { set_error("dof map is not a permutation (no node maps to index %lld)", hb[1]); return STAN_E_ARG; }
    return STAN_OK;
}

static int upload_dof_map(stan_handle *h, const int32_t *node_index, bool verify) {
    cudaStream_t s = h->stream;
    h->have_dof = false;
    h->assembled = h->solved = h->recovered = h->postprocessed = false;
    STAN_TRY(h->d_node_index.alloc(h->n_nodes, s));
    STAN_CUDA(cudaMemcpyAsync(h->d_node_index.p, node_index, h->n_nodes * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    if (verify) STAN_TRY(check_dof_map(h));
    else STAN_CUDA(cudaStreamSynchronize(s));
    h->have_dof = true;
    return STAN_OK;
}

}  // namespace stan

using namespace stan;

extern "C" {

const char *stan_last_error(void) { return g_err; }
int stan_version(void) { return 100; }

int stan_create(const stan_options *opts, stan_handle **out) {
    if (!out) { set_error("null out pointer"); return STAN_E_ARG; }
    *out = nullptr;
    int dev = opts ? opts->device : -1;
    if (dev < 0) STAN_CUDA(cudaGetDevice(&dev));
    STAN_CUDA(cudaSetDevice(dev));
    stan_handle *h = new stan_handle();
    h->device = dev;
    h->rank = opts ? opts->rank : 0;
    h->world = opts && opts->world > 0 ? opts->world : 1;
    if (h->rank < 0 || h->rank >= h->world || h->world > 32) {
        set_error("rank %d outside world %d (max 32)", h->rank, h->world);
        delete h;
        return STAN_E_ARG;
    }
    cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, dev);
    auto init = [&]() -> int {
        STAN_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        STAN_CUDA(cudaStreamCreateWithFlags(&h->comm_stream, cudaStreamNonBlocking));
        STAN_CUDA(cudaEventCreate(&h->ev0)); STAN_CUDA(cudaEventCreate(&h->ev1));
        STAN_CUDA(cudaEventCreate(&h->ev2)); STAN_CUDA(cudaEventCreate(&h->ev3));
        cudaMemPool_t pool;
        STAN_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
        uint64_t keep = UINT64_MAX;     // keep freed blocks cached: assemble/solve cycles reuse them
        STAN_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        STAN_TRY(upload_fe_tables());
        return STAN_OK;
    };
    const int rc = init();
    if (rc != STAN_OK) {                 // release whatever was created; the error message stays
        if (h->stream) cudaStreamDestroy(h->stream);
        if (h->comm_stream) cudaStreamDestroy(h->comm_stream);
        for (cudaEvent_t e : {h->ev0, h->ev1, h->ev2, h->ev3}) if (e) cudaEventDestroy(e);
        delete h;
        return rc;
    }
    *out = h;
    return STAN_OK;
}

int stan_destroy(stan_handle *h) {
    if (!h) return STAN_OK;
    cudaSetDevice(h->device);
    cudaStream_t s = h->stream;
    cudaStreamSynchronize(s);
    comm_destroy(h);
    h->d_xyz.release(s); h->d_conn.release(s); h->d_etype.release(s); h->d_emat.release(s);
    h->d_lambda.release(s); h->d_G.release(s); h->d_node_index.release(s); h->d_inv.release(s);
    h->d_inc_ptr.release(s); h->d_inc.release(s); h->d_brow_ptr.release(s); h->d_bcol.release(s);
    h->d_bcol_loc.release(s); h->d_vals.release(s); h->d_fixed.release(s); h->d_red.release(s);
    h->d_b.release(s); h->d_d2.release(s); h->d_err.release(s); h->d_x.release(s); h->d_xalt.release(s);
    h->d_r.release(s); h->d_p.release(s); h->d_mv.release(s); h->d_partials.release(s); h->d_state.release(s);
    h->d_counter.release(s); h->d_ufull.release(s); h->d_strain.release(s); h->d_stress.release(s);
    h->d_cell.release(s); h->d_point.release(s); h->d_ke.release(s); h->d_hist.release(s); h->d_trace.release(s);
    for (auto &sl : h->scratch) if (sl.p) cudaFreeAsync(sl.p, s);
    if (h->h_state) cudaFreeHost(h->h_state);
    for (cudaEvent_t e : h->ev_pool) if (e) cudaEventDestroy(e);
    cudaStreamSynchronize(s);
    cudaEventDestroy(h->ev0); cudaEventDestroy(h->ev1); cudaEventDestroy(h->ev2); cudaEventDestroy(h->ev3);
    for (int i = 0; i < 8; i++) if (h->user_ev[i]) cudaEventDestroy(h->user_ev[i]);
    for (int k = 0; k < 2; k++) if (h->stage[k]) { cudaFreeHost(h->stage[k]); cudaEventDestroy(h->stage_ev[k]); }
    cudaStreamDestroy(h->stream); cudaStreamDestroy(h->comm_stream);
    delete h;
    return STAN_OK;
}

int stan_set_mesh(stan_handle *h, int64_t n_nodes, const double *xyz, int64_t n_elem, const int32_t *conn,
                  const uint8_t *elem_type, const int32_t *elem_mat) {
    STAN_TRY(check(h));
    if (n_nodes <= 0 || n_elem <= 0 || !xyz || !conn || !elem_type || !elem_mat) {
        set_error("stan_set_mesh: empty mesh or null array"); return STAN_E_ARG;
    }
    if (3 * n_nodes > INT32_MAX) { set_error("more than %d DOFs", INT32_MAX); return STAN_E_ARG; }
    // Upload first, check on the device, and only then replace the previous model: a rejected call leaves the
    // handle without a mesh rather than with a half-updated one.
    cudaStream_t s = h->stream;
    h->have_mesh = h->have_dof = h->assembled = h->solved = h->recovered = h->postprocessed = false;
    h->h_conn.clear(); h->h_conn.shrink_to_fit();
    STAN_TRY(h->d_xyz.alloc(3 * n_nodes, s)); STAN_TRY(h->d_conn.alloc(8 * n_elem, s));
    STAN_TRY(h->d_etype.alloc(n_elem, s)); STAN_TRY(h->d_emat.alloc(n_elem, s));
    STAN_CUDA(cudaMemcpyAsync(h->d_xyz.p, xyz, 3 * n_nodes * sizeof(double), cudaMemcpyHostToDevice, s));
    STAN_CUDA(cudaMemcpyAsync(h->d_conn.p, conn, 8 * n_elem * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    STAN_CUDA(cudaMemcpyAsync(h->d_etype.p, elem_type, n_elem, cudaMemcpyHostToDevice, s));
    STAN_CUDA(cudaMemcpyAsync(h->d_emat.p, elem_mat, n_elem * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    ScratchBuf<long long> res(&h->scratch[9]);
    STAN_TRY(res.alloc(5, s));
    const long long init[5] = {INT64_MAX, INT64_MAX, INT64_MAX, 0, 0};
    STAN_CUDA(cudaMemcpyAsync(res.p, init, sizeof init, cudaMemcpyHostToDevice, s));
    k_check_mesh<<<div_up(n_elem, 256), 256, 0, s>>>(n_elem, n_nodes, h->d_conn.p, h->d_etype.p, h->d_emat.p, res.p);
    long long r[5];
    STAN_CUDA(cudaMemcpyAsync(r, res.p, sizeof r, cudaMemcpyDeviceToHost, s));
    STAN_CUDA(cudaStreamSynchronize(s));
    STAN_CUDA(cudaGetLastError());
    h->launches += 1;
    if (r[0] < n_elem) {
        const int32_t *c = conn + 8 * r[0];
        int32_t badv = c[0];
        for (int k = 0; k < 8; k++) if (c[k] < 0 || c[k] >= n_nodes) { badv = c[k]; break; }
        set_error("element %lld references node %d outside [0,%lld)", r[0], badv, (long long)n_nodes);
        return STAN_E_ARG;
    }
    if (r[1] < n_elem) {
        set_error("element %lld has unsupported type %d (only HEX8_G1/HEX8_G2)", r[1], (int)elem_type[r[1]]);
        return STAN_E_ARG;
    }
    if (r[2] < n_elem) { set_error("element %lld has a negative material index", r[2]); return STAN_E_ARG; }
    h->n_nodes = n_nodes; h->n_elem = n_elem; h->n_elem_g2 = r[4];
    h->max_mat_index = (int32_t)r[3];
    h->nnz_upper = 0;
    h->have_mesh = true;
    h->h_spc_node.clear(); h->h_spc_val.clear(); h->h_load_node.clear(); h->h_load_val.clear();
    return STAN_OK;
}

int stan_set_materials(stan_handle *h, int32_t n_mat, const double *E, const double *nu) {
    STAN_TRY(check(h));
    if (n_mat <= 0 || !E || !nu) { set_error("stan_set_materials: no materials"); return STAN_E_ARG; }
    std::vector<double> lam(n_mat), G(n_mat);
    for (int i = 0; i < n_mat; i++) {                       // Material.SetElastic, Material.cs:39-40
        const double Young = E[i], Poisson = nu[i];
        lam[i] = (Young * Poisson) / ((1 - 2 * Poisson) * (1 + Poisson));
        G[i] = (0.5 * Young) / (1 + Poisson);
        if (!std::isfinite(lam[i]) || !std::isfinite(G[i])) { set_error("material %d: E=%g nu=%g gives a non-finite D", i, Young, Poisson); return STAN_E_ARG; }
    }
    cudaStream_t s = h->stream;
    STAN_TRY(h->d_lambda.alloc(n_mat, s)); STAN_TRY(h->d_G.alloc(n_mat, s));
    STAN_CUDA(cudaMemcpyAsync(h->d_lambda.p, lam.data(), n_mat * sizeof(double), cudaMemcpyHostToDevice, s));
    STAN_CUDA(cudaMemcpyAsync(h->d_G.p, G.data(), n_mat * sizeof(double), cudaMemcpyHostToDevice, s));
    STAN_CUDA(cudaStreamSynchronize(s));
    h->n_mat = n_mat;
    h->have_mat = true;
    h->assembled = h->solved = h->recovered = h->postprocessed = false;
    return STAN_OK;
}

int stan_set_dof_map(stan_handle *h, const int32_t *node_index) {
    STAN_TRY(check(h));
    if (!h->have_mesh) { set_error("stan_set_dof_map before stan_set_mesh"); return STAN_E_STATE; }
    if (!node_index) { set_error("null dof map"); return STAN_E_ARG; }
    return upload_dof_map(h, node_index, true);
}

int stan_assign_dof(stan_handle *h, int32_t *node_index_out) {
    STAN_TRY(check(h));
    if (!h->have_mesh) { set_error("stan_assign_dof before stan_set_mesh"); return STAN_E_STATE; }
    std::vector<int32_t> ni((size_t)h->n_nodes);
    // Same numbering either way (tested bit for bit).  The level-synchronous device traversal pays
    // ~10 launches per BFS level, so it is for big meshes with wide levels; STAN_DOF=gpu|host forces one.
    const char *mode = getenv("STAN_DOF");
    bool on_device = mode ? !strcmp(mode, "gpu") : h->n_nodes >= 1000000;
    if (on_device) {
        bool narrow = false;
        STAN_TRY(assign_dof_device(h, ni.data(), &narrow));
        on_device = !narrow;
    }
    if (!on_device) {
        if (h->h_conn.empty()) {                           // host copy of the connectivity, only when this path needs it
            h->h_conn.resize((size_t)8 * h->n_elem);
            STAN_CUDA(cudaMemcpyAsync(h->h_conn.data(), h->d_conn.p, h->h_conn.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
            STAN_CUDA(cudaStreamSynchronize(h->stream));
        }
        STAN_TRY(assign_dof_host(h->n_nodes, h->n_elem, h->h_conn.data(), ni.data()));
    }
    if (node_index_out) memcpy(node_index_out, ni.data(), ni.size() * sizeof(int32_t));
    return upload_dof_map(h, ni.data(), false);
}

int stan_set_spc(stan_handle *h, int64_t n, const int32_t *node, const double *val3) {
    STAN_TRY(check(h));
    if (!h->have_mesh) { set_error("stan_set_spc before stan_set_mesh"); return STAN_E_STATE; }
    if (n < 0 || (n > 0 && (!node || !val3))) { set_error("stan_set_spc: bad arguments"); return STAN_E_ARG; }
    for (int64_t i = 0; i < n; i++)
        if (node[i] < 0 || node[i] >= h->n_nodes) { set_error("SPC entry %lld: node %d not in the mesh", (long long)i, node[i]); return STAN_E_ARG; }
    h->h_spc_node.assign(node, node + n);
    h->h_spc_val.assign(val3, val3 + 3 * n);
    h->assembled = h->solved = h->recovered = h->postprocessed = false;
    return STAN_OK;
}

int stan_set_loads(stan_handle *h, int64_t n, const int32_t *node, const double *fxyz) {
    STAN_TRY(check(h));
    if (!h->have_mesh) { set_error("stan_set_loads before stan_set_mesh"); return STAN_E_STATE; }
    if (n < 0 || (n > 0 && (!node || !fxyz))) { set_error("stan_set_loads: bad arguments"); return STAN_E_ARG; }
    for (int64_t i = 0; i < n; i++)
        if (node[i] < 0 || node[i] >= h->n_nodes) { set_error("load entry %lld: node %d not in the mesh", (long long)i, node[i]); return STAN_E_ARG; }
    h->h_load_node.assign(node, node + n);
    h->h_load_val.assign(fxyz, fxyz + 3 * n);
    h->assembled = h->solved = h->recovered = h->postprocessed = false;
    return STAN_OK;
}

int stan_assemble(stan_handle *h, stan_assembly_stats *stats) {
    STAN_TRY(check(h));
    if (!h->have_mesh || !h->have_mat || !h->have_dof) {
        set_error("stan_assemble needs mesh, materials and a dof map"); return STAN_E_STATE;
    }
    if (h->max_mat_index >= h->n_mat) {                       // DB.MatLib[MatID] would throw KeyNotFound
        set_error("an element uses material index %d but only %d materials were set", h->max_mat_index, h->n_mat);
        return STAN_E_ARG;
    }
    cudaStream_t s = h->stream;
    const int64_t launches0 = h->launches;
    h->assembled = h->solved = h->recovered = h->postprocessed = false;
    STAN_TRY(partition_rows(h));
    static const bool trace = getenv("STAN_TRACE") != nullptr;      // host wall clock of the phases, to stderr
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double w0 = now();
    STAN_CUDA(cudaEventRecord(h->ev2, s));
    STAN_TRY(build_system_pattern(h));
    const double w1 = now();
    STAN_TRY(comm_build_halo(h));
    const double w2 = now();
    STAN_TRY(build_rhs(h));
    const double w3 = now();
    STAN_CUDA(cudaEventRecord(h->ev0, s));
    STAN_TRY(run_assembly(h));
    if (trace) {
        cudaStreamSynchronize(s);
        fprintf(stderr, "[stan rank %d] assemble: pattern %.1f ms, halo/p2p %.1f ms, rhs %.1f ms, kernels %.1f ms\n", h->rank,
                w1 - w0, w2 - w1, w3 - w2, now() - w3);
    }
    STAN_CUDA(cudaEventRecord(h->ev1, s));
    int32_t herr[4];
    STAN_CUDA(cudaMemcpyAsync(herr, h->d_err.p, sizeof herr, cudaMemcpyDeviceToHost, s));
    STAN_CUDA(cudaEventRecord(h->ev3, s));
    STAN_CUDA(cudaStreamSynchronize(s));
    STAN_CUDA(cudaGetLastError());
    if (herr[2]) { set_error("singular Jacobian: an element has det(J) == 0 at a Gauss point"); return STAN_E_SINGULAR; }
    h->assembled = true;
    if (stats) {
        float t_pat = 0.f, t_asm = 0.f, t_all = 0.f;
        cudaEventElapsedTime(&t_pat, h->ev2, h->ev0);
        cudaEventElapsedTime(&t_asm, h->ev0, h->ev1);
        cudaEventElapsedTime(&t_all, h->ev2, h->ev3);
        memset(stats, 0, sizeof *stats);
        stats->n_dof = 3 * h->n_nodes;
        stats->n_fixed = h->n_fixed;
        stats->n_rows_local = h->row1 - h->row0;
        stats->n_blocks_local = h->n_blocks;
        stats->nnz_upper = h->nnz_upper;
        // each stored value written once + connectivity/coordinates read once
        stats->assembly_bytes = 72 * h->n_blocks + h->n_elem * 40 + 24 * h->n_nodes;
        // structured minimum per element: G2 17.3 kflop, G1 2.2 kflop, by the actual type counts
        stats->assembly_flops = (double)h->n_elem_g2 * 17300.0 + (double)(h->n_elem - h->n_elem_g2) * 2200.0;
        stats->pattern_ms = t_pat;
        stats->assembly_ms = t_asm;
        stats->total_ms = t_all;
        stats->kernel_launches = h->launches - launches0;
    }
    return STAN_OK;
}

int stan_solve_cg(stan_handle *h, const stan_cg_options *opts, stan_cg_report *report) {
    STAN_TRY(check(h));
    if (!h->assembled) { set_error("stan_solve_cg before stan_assemble"); return STAN_E_STATE; }
    if (!opts || !report) { set_error("stan_solve_cg: null options/report"); return STAN_E_ARG; }
    if (opts->epsf < 0 || opts->maxits < 0) { set_error("stan_solve_cg: negative EpsF/MaxIts"); return STAN_E_ARG; }
    if (!opts->merit_check && opts->maxits == 0 && opts->epsf < 1e-12) {
        set_error("merit_check = 0 with EpsF < 1e-12 needs MaxIts > 0 (FP64 cannot reach it; the loop would not end)");
        return STAN_E_ARG;
    }
    memset(report, 0, sizeof *report);
    h->recovered = h->postprocessed = false;
    return solve_cg(h, opts, report);
}

int stan_set_cg_history(stan_handle *h, int32_t capacity) {
    STAN_TRY(check(h));
    if (capacity < 0) { set_error("stan_set_cg_history: negative capacity"); return STAN_E_ARG; }
    h->hist_cap = capacity;
    h->hist_count = 0;
    return STAN_OK;
}

int stan_get_cg_history(stan_handle *h, int32_t *count, double *hist4) {
    STAN_TRY(check(h));
    if (!h->solved || !count) { set_error("stan_get_cg_history: no CG solve yet or null count"); return STAN_E_STATE; }
    *count = h->hist_count;
    if (hist4 && h->hist_count > 0) {
        STAN_CUDA(cudaMemcpyAsync(hist4, h->d_hist.p, (size_t)4 * h->hist_count * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        STAN_CUDA(cudaStreamSynchronize(h->stream));
    }
    return STAN_OK;
}

int stan_solve_cholesky(stan_handle *h, stan_chol_report *report) {
    STAN_TRY(check(h));
    if (!h->assembled) { set_error("stan_solve_cholesky before stan_assemble"); return STAN_E_STATE; }
    if (!report) { set_error("stan_solve_cholesky: null report"); return STAN_E_ARG; }
    h->recovered = h->postprocessed = false;
    return solve_cholesky(h, report);
}

int stan_recover(stan_handle *h, stan_recovery_stats *stats) {
    STAN_TRY(check(h));
    if (!h->solved) { set_error("stan_recover before a solve"); return STAN_E_STATE; }
    return run_recovery(h, stats);
}

int stan_get_displacements(stan_handle *h, double *u_full) {
    STAN_TRY(check(h));
    if (!h->solved) { set_error("no solution yet"); return STAN_E_STATE; }
    if (!u_full) { set_error("null output"); return STAN_E_ARG; }
    if (!h->recovered) STAN_TRY(scatter_solution(h));
    return copy_to_host(h, u_full, h->d_ufull.p, 3 * h->n_nodes * sizeof(double));
}

int stan_get_displacements_local(stan_handle *h, double *u_rows) {
    STAN_TRY(check(h));
    if (!h->solved) { set_error("no solution yet"); return STAN_E_STATE; }
    if (!u_rows) { set_error("null output"); return STAN_E_ARG; }
    return copy_to_host(h, u_rows, h->sol, (size_t)3 * (h->row1 - h->row0) * sizeof(double));
}

int stan_get_node_displacements(stan_handle *h, double *disp) {
    STAN_TRY(check(h));
    if (!h->solved) { set_error("no solution yet"); return STAN_E_STATE; }
    if (!disp) { set_error("null output"); return STAN_E_ARG; }
    if (!h->recovered) STAN_TRY(scatter_solution(h));
    cudaStream_t s = h->stream;
    DevBuf<double> tmp;
    STAN_TRY(tmp.alloc((size_t)3 * h->n_nodes, s));
    k_gather_rows3<<<div_up(3 * h->n_nodes, 256), 256, 0, s>>>(h->n_nodes, h->d_node_index.p, h->d_ufull.p, tmp.p);
    h->launches += 1;
    const int rc = copy_to_host(h, disp, tmp.p, (size_t)3 * h->n_nodes * sizeof(double));
    tmp.release(s);
    return rc;
}

int stan_host_alloc(size_t bytes, void **out) {
    if (!out) { set_error("null out pointer"); return STAN_E_ARG; }
    *out = nullptr;
    STAN_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable));
    return STAN_OK;
}

int stan_host_free(void *p) {
    if (p) STAN_CUDA(cudaFreeHost(p));
    return STAN_OK;
}

int stan_get_strain_stress(stan_handle *h, double *strain, double *stress) {
    STAN_TRY(check(h));
    if (!h->recovered) { set_error("stan_get_strain_stress before stan_recover"); return STAN_E_STATE; }
    const size_t bytes = (size_t)48 * (h->elem1 - h->elem0) * sizeof(double);
    if (strain) STAN_TRY(copy_to_host(h, strain, h->d_strain.p, bytes));
    if (stress) STAN_TRY(copy_to_host(h, stress, h->d_stress.p, bytes));
    return STAN_OK;
}

int stan_postprocess(stan_handle *h, double *ms) {
    STAN_TRY(check(h));
    if (!h->recovered) { set_error("stan_postprocess before stan_recover"); return STAN_E_STATE; }
    return run_postprocess(h, ms);
}

int stan_get_scalars(stan_handle *h, float *cell, float *point) {
    STAN_TRY(check(h));
    if (!h->postprocessed || !h->recovered) { set_error("stan_get_scalars before stan_postprocess"); return STAN_E_STATE; }
    // one GPU: every element / every node; partitioned: the rank's element slice and its rows
    if (cell) STAN_CUDA(cudaMemcpyAsync(cell, h->d_cell.p, (size_t)(h->elem1 - h->elem0) * 72 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    if (point) STAN_CUDA(cudaMemcpyAsync(point, h->d_point.p, (size_t)(h->row1 - h->row0) * 24 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    STAN_CUDA(cudaStreamSynchronize(h->stream));
    return STAN_OK;
}

int stan_get_dof_reduction(stan_handle *h, int32_t *ndof_reduction) {
    STAN_TRY(check(h));
    if (!h->assembled) { set_error("not assembled"); return STAN_E_STATE; }
    STAN_CUDA(cudaMemcpyAsync(ndof_reduction, h->d_red.p, 3 * h->n_nodes * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    STAN_CUDA(cudaStreamSynchronize(h->stream));
    return STAN_OK;
}

static int reduced_from_full(stan_handle *h, const double *d_full_local, double *out) {
    if (h->world != 1) { set_error("reduced-vector export is single-GPU only"); return STAN_E_STATE; }
    const int64_t ndof = 3 * h->n_nodes;
    std::vector<double> full((size_t)ndof);
    std::vector<int32_t> red((size_t)ndof);
    STAN_CUDA(cudaMemcpyAsync(full.data(), d_full_local, ndof * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    STAN_CUDA(cudaMemcpyAsync(red.data(), h->d_red.p, ndof * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    STAN_CUDA(cudaStreamSynchronize(h->stream));
    for (int64_t i = 0; i < ndof; i++)                      // Exclude_BC_DOF, SolverFunctions.cs:540-555
        if (red[i] != -1) out[i - red[i]] = full[i];
    return STAN_OK;
}

int stan_get_rhs(stan_handle *h, double *F_reduced) {
    STAN_TRY(check(h));
    if (!h->assembled) { set_error("not assembled"); return STAN_E_STATE; }
    return reduced_from_full(h, h->d_b.p, F_reduced);
}

int stan_get_solution_reduced(stan_handle *h, double *U_reduced) {
    STAN_TRY(check(h));
    if (!h->solved) { set_error("no solution yet"); return STAN_E_STATE; }
    return reduced_from_full(h, h->sol, U_reduced);
}

int stan_get_csr_upper_size(stan_handle *h, int64_t *n, int64_t *nnz) {
    STAN_TRY(check(h));
    if (!h->assembled) { set_error("not assembled"); return STAN_E_STATE; }
    return export_csr_upper_size(h, n, nnz);
}

int stan_get_csr_upper(stan_handle *h, int64_t *rowptr, int32_t *col, double *val) {
    STAN_TRY(check(h));
    if (!h->assembled) { set_error("not assembled"); return STAN_E_STATE; }
    return export_csr_upper(h, rowptr, col, val);
}

int stan_element_stiffness(stan_handle *h, int64_t first, int64_t count, double *ke) {
    STAN_TRY(check(h));
    if (!h->have_mesh || !h->have_mat) { set_error("needs mesh and materials"); return STAN_E_STATE; }
    if (first < 0 || count <= 0 || first + count > h->n_elem || !ke) { set_error("bad element range"); return STAN_E_ARG; }
    return element_stiffness(h, first, count, ke);
}

int stan_spmv(stan_handle *h, const double *x_full, double *y_full) {
    STAN_TRY(check(h));
    if (!h->assembled) { set_error("not assembled"); return STAN_E_STATE; }
    return spmv_full(h, x_full, y_full);
}

int stan_time_spmv(stan_handle *h, int32_t reps, double *ms_per_launch, int64_t *bytes_per_launch) {
    STAN_TRY(check(h));
    if (!h->assembled || reps <= 0 || !ms_per_launch) { set_error("stan_time_spmv: not assembled or bad args"); return STAN_E_STATE; }
    return time_spmv(h, reps, ms_per_launch, bytes_per_launch);
}

int stan_comm_unique_id(void *id128) { return comm_unique_id(id128); }

int stan_comm_init(stan_handle *h, const void *id128) {
    STAN_TRY(check(h));
    return comm_init(h, id128);
}

int stan_get_element_range(stan_handle *h, int64_t *first, int64_t *last) {
    if (!h) return STAN_E_ARG;
    if (!h->recovered) { set_error("stan_get_element_range before stan_recover"); return STAN_E_STATE; }
    if (first) *first = h->elem0;
    if (last) *last = h->elem1;
    return STAN_OK;
}

int stan_get_partition(stan_handle *h, int64_t *first_row, int64_t *last_row) {
    if (!h) return STAN_E_ARG;
    if (first_row) *first_row = h->row0;
    if (last_row) *last_row = h->row1;
    return STAN_OK;
}

int64_t stan_kernel_launches(stan_handle *h) { return h ? h->launches : 0; }

int stan_event_record(stan_handle *h, int32_t slot) {
    STAN_TRY(check(h));
    if (slot < 0 || slot >= 8) { set_error("event slot %d outside [0,8)", slot); return STAN_E_ARG; }
    if (!h->user_ev[slot]) STAN_CUDA(cudaEventCreate(&h->user_ev[slot]));
    STAN_CUDA(cudaEventRecord(h->user_ev[slot], h->stream));
    return STAN_OK;
}

int stan_event_elapsed(stan_handle *h, int32_t a, int32_t b, double *ms) {
    STAN_TRY(check(h));
    if (a < 0 || a >= 8 || b < 0 || b >= 8 || !ms || !h->user_ev[a] || !h->user_ev[b]) {
        set_error("stan_event_elapsed: slots not recorded"); return STAN_E_ARG;
    }
    STAN_CUDA(cudaEventSynchronize(h->user_ev[b]));
    float f = 0.f;
    STAN_CUDA(cudaEventElapsedTime(&f, h->user_ev[a], h->user_ev[b]));
    *ms = f;
    return STAN_OK;
}

}  // extern "C"
