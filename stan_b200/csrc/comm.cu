// Multi-GPU plumbing: one process per GPU, contiguous BFS-node row ranges.
//
// The reference is single-process (its only parallelism is Parallel.ForEach,
// /root/reference/src/STAN_Solver/SolverFunctions.cs:129); this layer is new.  NCCL is resolved at
// run time with dlopen so the single-GPU library has no NCCL dependency and a process that already
// loaded torch's libnccl.so.2 shares that copy.  Per SpMV each rank sends the entries of p that
// its neighbours' boundary rows reference and receives its own halo; CG dot products are one
// small all-reduce each.  Because adjacency is symmetric, a rank derives both its halo list and
// its send lists from its own block columns — no set-up communication is needed.
//
// Default data plane (peer memory over NVLink, no library call in the CG loop): the three vectors an
// SpMV reads live in a cudaMalloc'ed window that every peer maps through CUDA IPC.  The product kernel
// itself stores its boundary entries into the halo tails of the neighbours' copies of that vector while
// its consumer warps work through the tiles that need no halo, raises a sequence flag, and only the
// tiles with halo columns wait for the neighbours' flags (cg.cu: k_spmv_tile3<.., true>).  NCCL is used
// for set-up, and as the whole data plane with STAN_COMM=nccl or when IPC is unavailable.
#include <cub/device/device_scan.cuh>
#include <dlfcn.h>

#include "common.cuh"

namespace stan {

typedef struct { char internal[128]; } nccl_uid;
typedef void *nccl_comm;

struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(nccl_uid *) = nullptr;
    int (*CommInitRank)(nccl_comm *, int, nccl_uid, int) = nullptr;
    int (*CommDestroy)(nccl_comm) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, nccl_comm, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, nccl_comm, cudaStream_t) = nullptr;
    int (*Send)(const void *, size_t, int, int, nccl_comm, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, nccl_comm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};

static NcclApi g_nccl;
constexpr int NCCL_F64 = 8, NCCL_SUM = 0;

static int load_nccl() {
    if (g_nccl.lib) return STAN_OK;
    const char *names[] = {"libnccl.so.2", "libnccl.so", nullptr};
    void *lib = nullptr;
    for (int i = 0; names[i] && !lib; i++) lib = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    if (!lib) {
        const char *env = getenv("STAN_NCCL_LIB");
        if (env) lib = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
    }
    if (!lib) { set_error("cannot dlopen libnccl.so.2 (set STAN_NCCL_LIB): %s", dlerror()); return STAN_E_COMM; }
#define SYM(field, name)                                                              \
    *(void **)(&g_nccl.field) = dlsym(lib, name);                                     \
    if (!g_nccl.field) { set_error("libnccl lacks %s", name); return STAN_E_COMM; }
    SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy")
    SYM(AllReduce, "ncclAllReduce") SYM(AllGather, "ncclAllGather") SYM(Send, "ncclSend") SYM(Recv, "ncclRecv")
    SYM(GroupStart, "ncclGroupStart") SYM(GroupEnd, "ncclGroupEnd") SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    g_nccl.lib = lib;
    return STAN_OK;
}

#define STAN_NCCL(call)                                                                            \
    do {                                                                                           \
        int r__ = (call);                                                                          \
        if (r__ != 0) { set_error("%s -> %s", #call, g_nccl.GetErrorString(r__)); return STAN_E_COMM; } \
    } while (0)

struct Comm {
    nccl_comm comm = nullptr;
    // ---- peer-memory path ----
    bool p2p = false, p2p_failed = false;
    void *window = nullptr;                // this rank's cudaMalloc'ed window (P2PCtrl + the vectors p, x, xalt)
    size_t window_bytes = 0;
    size_t vec_stride = 0;                 // doubles between the window vectors
    unsigned long long epoch = 0;          // solves so far: high half of every halo sequence number
    DevBuf<int32_t> d_tile_order;          // SpMV tiles, the ones without halo columns first
    int64_t n_interior = 0;
    std::vector<int32_t> h_send_rows;      // host copy of d_send_rows (the send map is built from it)
    DevBuf<int32_t> d_send_map, d_send_ptr;
    DevBuf<unsigned long long> d_send_dst;
    void *peer_window[P2P_MAX_RANKS] = {};   // IPC mappings of the other ranks' windows
    DevBuf<CommDev> d_dev;
    DevBuf<unsigned int> d_ticket;
    unsigned long long halo_seq = 0;       // monotonic across solves
    std::vector<int64_t> bound;            // world+1 row bounds
    std::vector<int64_t> recv_off, recv_cnt;   // per peer: segment of the halo region (nodes)
    std::vector<int64_t> send_off, send_cnt;   // per peer: segment of the send list (nodes)
    DevBuf<int32_t> d_send_rows;           // local row index of every node to send, grouped by peer
    DevBuf<double> d_sendbuf;              // 3 doubles per send node
    DevBuf<double> d_gather;               // padded all-gather staging
    int64_t n_send = 0, max_rows = 0;
};

int comm_unique_id(void *id128) {
    STAN_TRY(load_nccl());
    STAN_NCCL(g_nccl.GetUniqueId((nccl_uid *)id128));
    return STAN_OK;
}

int comm_init(stan_handle *h, const void *id128) {
    if (h->world <= 1) return STAN_OK;
    STAN_TRY(load_nccl());
    if (!h->comm) h->comm = new Comm();
    nccl_uid id;
    memcpy(&id, id128, sizeof id);
    STAN_CUDA(cudaSetDevice(h->device));
    STAN_NCCL(g_nccl.CommInitRank(&h->comm->comm, h->world, id, h->rank));
    return STAN_OK;
}

static void p2p_release(stan_handle *h) {
    Comm *c = h->comm;
    if (!c) return;
    for (int r = 0; r < P2P_MAX_RANKS; r++)
        if (c->peer_window[r]) { cudaIpcCloseMemHandle(c->peer_window[r]); c->peer_window[r] = nullptr; }
    if (c->window) { cudaFree(c->window); c->window = nullptr; }
    c->p2p = false;
}

bool comm_p2p_active(const stan_handle *h) { return h->comm && h->comm->p2p; }
CommDev *comm_dev(const stan_handle *h) { return comm_p2p_active(h) ? h->comm->d_dev.p : nullptr; }

void comm_destroy(stan_handle *h) {
    if (!h->comm) return;
    cudaStreamSynchronize(h->stream);
    p2p_release(h);
    h->comm->d_dev.release(h->stream);
    h->comm->d_ticket.release(h->stream);
    if (h->comm->comm) g_nccl.CommDestroy(h->comm->comm);
    h->comm->d_send_rows.release(h->stream);
    h->comm->d_sendbuf.release(h->stream);
    h->comm->d_gather.release(h->stream);
    h->comm->d_tile_order.release(h->stream);
    h->comm->d_send_map.release(h->stream); h->comm->d_send_ptr.release(h->stream); h->comm->d_send_dst.release(h->stream);
    delete h->comm;
    h->comm = nullptr;
}

int comm_allreduce_sum(stan_handle *h, double *d_buf, int count, cudaStream_t s) {
    if (h->world <= 1) return STAN_OK;
    STAN_NCCL(g_nccl.AllReduce(d_buf, d_buf, (size_t)count, NCCL_F64, NCCL_SUM, h->comm->comm, s));
    return STAN_OK;
}

namespace {

__global__ void k_mark_halo(int64_t nblk, const int32_t *__restrict__ bcol, int64_t row0, int64_t row1,
                            int32_t *__restrict__ flag) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= nblk) return;
    int32_t q = bcol[t];
    if (q < row0 || q >= row1) flag[q] = 1;
}

__global__ void k_localize_cols(int64_t nblk, const int32_t *__restrict__ bcol, int64_t row0, int64_t row1,
                                int64_t halo_base, const int32_t *__restrict__ slot, int32_t *__restrict__ out) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= nblk) return;
    int32_t q = bcol[t];
    out[t] = (q >= row0 && q < row1) ? (int32_t)(q - row0) : (int32_t)halo_base + slot[q];
}

// bit s of mask[row] is set when the row has a column owned by rank s
__global__ void k_peer_mask(int64_t nloc, const int32_t *__restrict__ brow_ptr, const int32_t *__restrict__ bcol,
                            int world, const int64_t *__restrict__ bound, int64_t row0, int64_t row1,
                            uint32_t *__restrict__ mask) {
    int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= nloc) return;
    uint32_t m = 0;
    for (int s = brow_ptr[p]; s < brow_ptr[p + 1]; s++) {
        int64_t q = bcol[s];
        if (q >= row0 && q < row1) continue;
        int r = 0;
        while (r + 1 < world && q >= bound[r + 1]) r++;
        m |= 1u << r;
    }
    mask[p] = m;
}

// Stand-alone halo exchange: the first exchange of a solve, the refresh products (x) and every exchange when
// k_direction does not push itself (STAN_DIR_PUSH=0, or SpMV variants other than the tile kernel).  Push: my
// boundary entries of window vector `vec_id` go straight into the halo tails of the ranks that read them; the
// last CTA raises the sequence flags after a system-scope fence and then waits for the neighbours' flags.  Every
// CTA takes a ticket even when the state says done, so the counter stays consistent.
__global__ void __launch_bounds__(256)
k_halo_push(const CommDev *__restrict__ cd, const double *__restrict__ vec, int vec_id, CgState *st) {
    const bool skip = st->done != 0;
    if (!skip && blockIdx.x == 0 && threadIdx.x == 0) trace_mark(st, TR_PUSH_BEGIN);
    const unsigned long long seq = st->halo_seq + 1;
    const int W = cd->world, me = cd->rank;
    if (!skip) {
        const int32_t *rows = cd->send_rows;
        for (int peer = 0; peer < W; peer++) {             // one contiguous destination range per peer
            const long long lo3 = 3 * cd->send_off[peer], hi3 = 3 * cd->send_off[peer + 1];
            if (hi3 == lo3) continue;
            double *dst = cd->vec[peer][vec_id] + cd->tail_off[peer] - lo3;
            for (long long e = lo3 + blockIdx.x * (long long)blockDim.x + threadIdx.x; e < hi3; e += (long long)gridDim.x * blockDim.x) {
                const long long i = e / 3;
                dst[e] = vec[3 * (long long)rows[i] + (e - 3 * i)];
            }
        }
    }
    __shared__ bool last;
    __syncthreads();
    if (threadIdx.x == 0) {                                // one fence per CTA, cumulative over the CTA's stores
        __threadfence_system();
        last = (atomicInc(cd->ticket, gridDim.x - 1) == gridDim.x - 1);
    }
    __syncthreads();
    if (last && !skip && threadIdx.x < W && threadIdx.x != me && cd->send_off[threadIdx.x + 1] > cd->send_off[threadIdx.x]) {
        __threadfence_system();
        *(volatile unsigned long long *)&cd->ctrl[threadIdx.x]->hflag[me] = seq;
    }
    if (last && !skip && threadIdx.x == 0) trace_mark(st, TR_PUSH_FLAGS);
    // The CTA that raised the flags also waits for the neighbours' (they do not depend on this wait: no cycle), so
    // the exchange is one launch.
    if (last && !skip && threadIdx.x < 32) {
        if (threadIdx.x == 0) trace_mark(st, TR_WAIT_BEGIN);
        if ((int)threadIdx.x < cd->n_recv_peers) {
            volatile unsigned long long *f = &cd->ctrl[me]->hflag[cd->recv_peer[threadIdx.x]];
            const long long t0 = clock64();
            while (*f < seq) {
                if (clock64() - t0 > 8000000000LL) { atomicOr(cd->err + 4, 1); break; }   // ~4 s: a peer died
            }
            __threadfence_system();
        }
        __syncwarp();
        if (threadIdx.x == 0) { st->halo_seq = seq; trace_mark(st, TR_WAIT_END); }
    }
}

__global__ void k_pack(int64_t n_send, const int32_t *__restrict__ rows, const double *__restrict__ vec,
                       double *__restrict__ buf) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= 3 * n_send) return;
    buf[t] = vec[3 * (int64_t)rows[t / 3] + t % 3];
}

}  // namespace

namespace {

// number of elements incident to every BFS row (the whole mesh: every rank computes the same bounds)
__global__ void k_valence(int64_t n_ent, const int32_t *__restrict__ conn, const int32_t *__restrict__ node_index,
                          int32_t *__restrict__ val) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t < n_ent) atomicAdd(&val[node_index[conn[t]]], 1);
}

// stored blocks of a row, estimated from its valence: 8 -> 27, 4 -> 15, 2 -> 9, 1 -> 6 on hexahedral meshes
__global__ void k_row_weight(int64_t n, const int32_t *__restrict__ val, int64_t *__restrict__ w) {
    int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p < n) w[p] = 3 * (int64_t)val[p] + 3;
}

// bound[r] = first row whose inclusive prefix weight reaches r/W of the total (bound[0] = 0, bound[W] = n)
__global__ void k_weight_bounds(int64_t n, const int64_t *__restrict__ prefix, int world, int64_t *__restrict__ bound) {
    const int r = threadIdx.x;
    if (r > world) return;
    if (r == 0) { bound[0] = 0; return; }
    if (r == world) { bound[r] = n; return; }
    const int64_t total = prefix[n - 1];
    const int64_t target = (total / world) * r + (total % world) * r / world;
    int64_t lo = 0, hi = n - 1;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (prefix[mid] < target) lo = mid + 1; else hi = mid;
    }
    bound[r] = lo + 1 < n ? lo + 1 : n;               // rows [.., lo] carry the target weight
}

}  // namespace

// Contiguous row ranges with (approximately) equal numbers of stored blocks: the SpMV of a rank
// costs its blocks, not its rows.  One GPU: the whole range.
int partition_rows(stan_handle *h) {
    const int W = h->world;
    h->bounds.assign((size_t)W + 1, 0);
    h->bounds[W] = h->n_nodes;
    if (W > 1) {
        cudaStream_t s = h->stream;
        const int64_t nn = h->n_nodes, ne8 = 8 * h->n_elem;
        ScratchBuf<int32_t> val(&h->scratch[0]);
        ScratchBuf<int64_t> w(&h->scratch[1]), pre(&h->scratch[2]), db(&h->scratch[3]);
        STAN_TRY(val.alloc(nn, s)); STAN_TRY(w.alloc(nn, s)); STAN_TRY(pre.alloc(nn, s)); STAN_TRY(db.alloc(W + 1, s));
        STAN_CUDA(cudaMemsetAsync(val.p, 0, nn * sizeof(int32_t), s));
        k_valence<<<div_up(ne8, 256), 256, 0, s>>>(ne8, h->d_conn.p, h->d_node_index.p, val.p);
        k_row_weight<<<div_up(nn, 256), 256, 0, s>>>(nn, val.p, w.p);
        {
            size_t bytes = 0;
            STAN_CUDA(cub::DeviceScan::InclusiveSum(nullptr, bytes, w.p, pre.p, nn, s));
            ScratchBuf<unsigned char> tmp(&h->scratch[8]);
            STAN_TRY(tmp.alloc(bytes ? bytes : 1, s));
            STAN_CUDA(cub::DeviceScan::InclusiveSum(tmp.p, bytes, w.p, pre.p, nn, s));
        }
        k_weight_bounds<<<1, P2P_MAX_RANKS + 1, 0, s>>>(nn, pre.p, W, db.p);
        STAN_CUDA(cudaMemcpyAsync(h->bounds.data(), db.p, (W + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
        STAN_CUDA(cudaStreamSynchronize(s));
        STAN_CUDA(cudaGetLastError());
        for (int r = 0; r < W; r++)                        // degenerate inputs (fewer rows than ranks): keep the ranges ordered
            if (h->bounds[r + 1] < h->bounds[r]) h->bounds[r + 1] = h->bounds[r];
        h->launches += 3;
    }
    h->row0 = h->bounds[h->rank];
    h->row1 = h->bounds[h->rank + 1];
    return STAN_OK;
}

// Peer-memory set-up: allocate this rank's window, swap IPC handles and landing offsets with an
// NCCL all-gather (set-up only), map every peer window.  Falls back to the NCCL path when IPC is
// unavailable or STAN_COMM=nccl.
static size_t window_ctrl_bytes() { return (sizeof(P2PCtrl) + 255) & ~(size_t)255; }

static int p2p_setup(stan_handle *h) {
    Comm *c = h->comm;
    cudaStream_t s = h->stream;
    const int W = h->world, me = h->rank;
    const char *mode = getenv("STAN_COMM");
    if ((mode && !strcmp(mode, "nccl")) || W > P2P_MAX_RANKS || c->p2p_failed) return STAN_OK;

    // phase 1: sizes and landing offsets (also a barrier: every rank has left its previous solve)
    struct Hello { long long nloc_pad, n_halo, need, cap; long long recv_off[P2P_MAX_RANKS]; };
    struct Hello2 { cudaIpcMemHandle_t handle; long long vec_stride; int ok, pad; };
    const size_t ctrl_bytes = window_ctrl_bytes();
    // one vector: owned rows padded, then the halo rows; rounded so the next vector starts on a 256-byte boundary
    const size_t vec_doubles = ((size_t)3 * (h->nloc_pad + std::max<int64_t>(h->n_halo, 1)) + 31) & ~(size_t)31;
    Hello mine;
    memset(&mine, 0, sizeof mine);
    mine.nloc_pad = h->nloc_pad;
    mine.n_halo = h->n_halo;
    mine.need = (long long)(ctrl_bytes + 3 * vec_doubles * sizeof(double));
    mine.cap = c->window ? (long long)c->window_bytes : 0;
    for (int r = 0; r < W; r++) mine.recv_off[r] = c->recv_off[r];
    std::vector<Hello> all(W);
    {
        ScratchBuf<Hello> dmine(&h->scratch[4]), dall(&h->scratch[5]);
        STAN_TRY(dmine.alloc(1, s)); STAN_TRY(dall.alloc(W, s));
        STAN_CUDA(cudaMemcpyAsync(dmine.p, &mine, sizeof mine, cudaMemcpyHostToDevice, s));
        STAN_NCCL(g_nccl.AllGather(dmine.p, dall.p, sizeof(Hello), 0 /* ncclInt8 */, c->comm, s));
        STAN_CUDA(cudaMemcpyAsync(all.data(), dall.p, W * sizeof(Hello), cudaMemcpyDeviceToHost, s));
        STAN_CUDA(cudaStreamSynchronize(s));
        dmine.release(s); dall.release(s);
    }
    const bool grow = mine.need > mine.cap;

    // phase 2: (new) allocations, IPC handles, vector strides; windows that did not grow keep their mapping
    Hello2 m2;
    memset(&m2, 0, sizeof m2);
    m2.ok = 1;
    if (grow) {
        if (c->window) { cudaFree(c->window); c->window = nullptr; }
        c->window_bytes = (size_t)mine.need + (size_t)mine.need / 8;
        m2.ok = cudaMalloc(&c->window, c->window_bytes) == cudaSuccess;
        if (m2.ok) cudaMemset(c->window, 0, c->window_bytes);
    }
    // the vectors are re-laid inside the (possibly larger) window for this model's size
    c->vec_stride = vec_doubles;
    m2.vec_stride = (long long)vec_doubles;
    if (m2.ok) m2.ok = cudaIpcGetMemHandle(&m2.handle, c->window) == cudaSuccess;
    cudaGetLastError();
    std::vector<Hello2> all2(W);
    {
        ScratchBuf<Hello2> dmine(&h->scratch[4]), dall(&h->scratch[5]);
        STAN_TRY(dmine.alloc(1, s)); STAN_TRY(dall.alloc(W, s));
        STAN_CUDA(cudaMemcpyAsync(dmine.p, &m2, sizeof m2, cudaMemcpyHostToDevice, s));
        STAN_NCCL(g_nccl.AllGather(dmine.p, dall.p, sizeof(Hello2), 0, c->comm, s));
        STAN_CUDA(cudaMemcpyAsync(all2.data(), dall.p, W * sizeof(Hello2), cudaMemcpyDeviceToHost, s));
        STAN_CUDA(cudaStreamSynchronize(s));
        dmine.release(s); dall.release(s);
    }
    bool ok = true;
    for (int r = 0; r < W; r++) ok = ok && all2[r].ok;
    for (int r = 0; r < W && ok; r++) {
        if (r == me || !(all[r].need > all[r].cap)) continue;       // unchanged windows keep their mapping
        if (c->peer_window[r]) { cudaIpcCloseMemHandle(c->peer_window[r]); c->peer_window[r] = nullptr; }
        if (cudaIpcOpenMemHandle(&c->peer_window[r], all2[r].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) ok = false;
    }
    cudaGetLastError();
    {   // every rank must take the same path: agree on success
        double flag = ok ? 0.0 : 1.0;
        DevBuf<double> dflag;
        STAN_TRY(dflag.alloc(1, s));
        STAN_CUDA(cudaMemcpyAsync(dflag.p, &flag, sizeof flag, cudaMemcpyHostToDevice, s));
        STAN_NCCL(g_nccl.AllReduce(dflag.p, dflag.p, 1, NCCL_F64, NCCL_SUM, c->comm, s));
        STAN_CUDA(cudaMemcpyAsync(&flag, dflag.p, sizeof flag, cudaMemcpyDeviceToHost, s));
        STAN_CUDA(cudaStreamSynchronize(s));
        dflag.release(s);
        if (flag != 0.0) { p2p_release(h); c->p2p_failed = true; return STAN_OK; }   // NCCL path on every rank
    }
    CommDev dev;
    memset(&dev, 0, sizeof dev);
    dev.rank = me; dev.world = W; dev.err = h->d_err.p;
    for (int r = 0; r < W; r++) {
        char *base = (char *)(r == me ? c->window : c->peer_window[r]);
        dev.ctrl[r] = (P2PCtrl *)base;
        for (int v = 0; v < 3; v++) dev.vec[r][v] = (double *)(base + ctrl_bytes) + (size_t)v * all2[r].vec_stride;
        dev.tail_off[r] = 3 * (all[r].nloc_pad + all[r].recv_off[me]);
        dev.recv_cnt[r] = c->recv_cnt[r];
        dev.send_off[r] = c->send_off[r];
        if (r != me && c->recv_cnt[r] > 0) dev.recv_peer[dev.n_recv_peers++] = r;
    }
    dev.send_off[W] = c->n_send;
    dev.send_rows = c->d_send_rows.p;
    {   // per-row send map: lets the kernel that writes p store its boundary entries into the peers' tails itself
        const int64_t nloc = h->row1 - h->row0;
        std::vector<int32_t> cnt((size_t)nloc, 0), map((size_t)nloc, -1);
        for (int32_t p : c->h_send_rows) cnt[p]++;
        std::vector<int32_t> ptr;
        ptr.push_back(0);
        for (int64_t p = 0; p < nloc; p++)
            if (cnt[p]) { map[p] = (int32_t)ptr.size() - 1; ptr.push_back(ptr.back() + cnt[p]); }
        std::vector<unsigned long long> dst((size_t)ptr.back());
        std::vector<int32_t> fill(ptr.begin(), ptr.end() - 1);
        for (int r = 0; r < W; r++)
            for (int64_t i = c->send_off[r]; i < c->send_off[r] + c->send_cnt[r]; i++) {
                const int32_t p = c->h_send_rows[i];
                dst[fill[map[p]]++] = ((unsigned long long)r << 48) | (unsigned long long)(dev.tail_off[r] + 3 * (i - c->send_off[r]));
            }
        STAN_TRY(c->d_send_map.alloc(nloc, s)); STAN_TRY(c->d_send_ptr.alloc(ptr.size(), s)); STAN_TRY(c->d_send_dst.alloc(dst.size(), s));
        STAN_CUDA(cudaMemcpyAsync(c->d_send_map.p, map.data(), nloc * sizeof(int32_t), cudaMemcpyHostToDevice, s));
        STAN_CUDA(cudaMemcpyAsync(c->d_send_ptr.p, ptr.data(), ptr.size() * sizeof(int32_t), cudaMemcpyHostToDevice, s));
        if (!dst.empty())
            STAN_CUDA(cudaMemcpyAsync(c->d_send_dst.p, dst.data(), dst.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice, s));
        STAN_CUDA(cudaStreamSynchronize(s));              // the host vectors go out of scope
        dev.send_map = c->d_send_map.p; dev.send_ptr = c->d_send_ptr.p; dev.send_dst = c->d_send_dst.p;
    }
    STAN_TRY(c->d_dev.alloc(1, s));
    if (!c->d_ticket.p) {
        STAN_TRY(c->d_ticket.alloc(1, s));
        STAN_CUDA(cudaMemsetAsync(c->d_ticket.p, 0, sizeof(unsigned int), s));
    }
    dev.ticket = c->d_ticket.p;
    STAN_CUDA(cudaMemcpyAsync(c->d_dev.p, &dev, sizeof dev, cudaMemcpyHostToDevice, s));
    STAN_CUDA(cudaStreamSynchronize(s));
    c->p2p = true;
    return STAN_OK;
}

int comm_build_halo(stan_handle *h) {
    cudaStream_t s = h->stream;
    const int64_t nloc = h->row1 - h->row0, nn = h->n_nodes;
    h->n_halo = 0;
    h->nloc_pad = nloc;
    if (h->world <= 1) { h->bcol_x = h->d_bcol.p; return STAN_OK; }
    if (!h->comm || !h->comm->comm) { set_error("world > 1 but stan_comm_init was not called"); return STAN_E_STATE; }
    Comm *c = h->comm;
    const int W = h->world;
    h->nloc_pad = (nloc + HALO_ALIGN - 1) / HALO_ALIGN * HALO_ALIGN;
    c->bound.assign(h->bounds.begin(), h->bounds.end());   // partition_rows(): balanced by estimated stored blocks
    c->max_rows = 0;
    for (int r = 0; r < W; r++) c->max_rows = std::max(c->max_rows, c->bound[r + 1] - c->bound[r]);

    ScratchBuf<int32_t> flag(&h->scratch[0]), slot(&h->scratch[1]);
    STAN_TRY(flag.alloc(nn + 1, s)); STAN_TRY(slot.alloc(nn + 1, s));
    STAN_CUDA(cudaMemsetAsync(flag.p, 0, (nn + 1) * sizeof(int32_t), s));
    k_mark_halo<<<div_up(h->n_blocks, 256), 256, 0, s>>>(h->n_blocks, h->d_bcol.p, h->row0, h->row1, flag.p);
    {
        size_t bytes = 0;
        STAN_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, flag.p, slot.p, nn + 1, s));
        ScratchBuf<unsigned char> tmp(&h->scratch[8]);
        STAN_TRY(tmp.alloc(bytes ? bytes : 1, s));
        STAN_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, bytes, flag.p, slot.p, nn + 1, s));
    }
    STAN_TRY(h->d_bcol_loc.alloc((size_t)h->n_blocks + 4, s));
    k_localize_cols<<<div_up(h->n_blocks, 256), 256, 0, s>>>(h->n_blocks, h->d_bcol.p, h->row0, h->row1, h->nloc_pad,
                                                             slot.p, h->d_bcol_loc.p);
    h->bcol_x = h->d_bcol_loc.p;
    // halo segment of every owner = difference of the scan at its bounds
    std::vector<int32_t> at(W + 1);
    for (int r = 0; r <= W; r++)
        STAN_CUDA(cudaMemcpyAsync(&at[r], slot.p + c->bound[r], sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    // send lists from the peer mask of the owned rows
    ScratchBuf<uint32_t> mask(&h->scratch[2]); ScratchBuf<int64_t> dbound(&h->scratch[3]);
    STAN_TRY(mask.alloc(nloc, s)); STAN_TRY(dbound.alloc(W + 1, s));
    STAN_CUDA(cudaMemcpyAsync(dbound.p, c->bound.data(), (W + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, s));
    k_peer_mask<<<div_up(nloc, 256), 256, 0, s>>>(nloc, h->d_brow_ptr.p, h->d_bcol.p, W, dbound.p, h->row0, h->row1,
                                                  mask.p);
    std::vector<uint32_t> hmask((size_t)nloc);
    STAN_CUDA(cudaMemcpyAsync(hmask.data(), mask.p, nloc * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    STAN_CUDA(cudaStreamSynchronize(s));
    STAN_CUDA(cudaGetLastError());
    flag.release(s); slot.release(s); mask.release(s); dbound.release(s);

    c->recv_off.assign(W, 0); c->recv_cnt.assign(W, 0); c->send_off.assign(W, 0); c->send_cnt.assign(W, 0);
    for (int r = 0; r < W; r++) { c->recv_off[r] = at[r]; c->recv_cnt[r] = at[r + 1] - at[r]; }
    h->n_halo = at[W];
    std::vector<int32_t> rows;
    for (int r = 0; r < W; r++) {
        c->send_off[r] = (int64_t)rows.size();
        if (r != h->rank)
            for (int64_t p = 0; p < nloc; p++)
                if (hmask[p] & (1u << r)) rows.push_back((int32_t)p);
        c->send_cnt[r] = (int64_t)rows.size() - c->send_off[r];
    }
    c->n_send = (int64_t)rows.size();
    c->h_send_rows = rows;
    STAN_TRY(c->d_send_rows.alloc(rows.size(), s));
    STAN_TRY(c->d_sendbuf.alloc(3 * rows.size(), s));
    if (!rows.empty())
        STAN_CUDA(cudaMemcpyAsync(c->d_send_rows.p, rows.data(), rows.size() * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    // SpMV tile order (32-row tiles, cg.cu): the tiles whose rows have no halo column first, so the product can
    // start on them while the halo is still in flight; a row has a halo column iff its peer mask is non-zero
    {
        const int64_t n_tiles = (nloc + 31) / 32;
        std::vector<int32_t> order((size_t)n_tiles);
        int64_t w = 0;
        std::vector<uint8_t> boundary((size_t)n_tiles, 0);
        for (int64_t p = 0; p < nloc; p++)
            if (hmask[p]) boundary[p >> 5] = 1;
        for (int64_t t = 0; t < n_tiles; t++) if (!boundary[t]) order[w++] = (int32_t)t;
        c->n_interior = w;
        for (int64_t t = 0; t < n_tiles; t++) if (boundary[t]) order[w++] = (int32_t)t;
        STAN_TRY(c->d_tile_order.alloc((size_t)n_tiles, s));
        STAN_CUDA(cudaMemcpyAsync(c->d_tile_order.p, order.data(), n_tiles * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    }
    STAN_CUDA(cudaStreamSynchronize(s));
    h->launches += 4;
    return p2p_setup(h);
}

// The vectors an SpMV is applied to: inside the peer window in peer-memory mode, pool allocations otherwise.
// Layout [3*nloc owned | pad | 3*n_halo halo] with the tail starting at 3*nloc_pad.
int comm_cg_vectors(stan_handle *h, double **p, double **x, double **xalt, cudaStream_t s) {
    const size_t nx = (size_t)3 * (h->nloc_pad + h->n_halo);
    if (comm_p2p_active(h)) {
        Comm *c = h->comm;
        double *base = (double *)((char *)c->window + window_ctrl_bytes());
        *p = base; *x = base + c->vec_stride; *xalt = base + 2 * c->vec_stride;
        return STAN_OK;
    }
    STAN_TRY(h->d_p.alloc(nx, s)); STAN_TRY(h->d_x.alloc(nx, s)); STAN_TRY(h->d_xalt.alloc(nx, s));
    *p = h->d_p.p; *x = h->d_x.p; *xalt = h->d_xalt.p;
    return STAN_OK;
}

bool comm_halo_args(const stan_handle *h, int vec_id, HaloArgs *out) {
    if (!comm_p2p_active(h)) return false;
    out->cd = h->comm->d_dev.p;
    out->tile_order = h->comm->d_tile_order.p;
    out->n_interior = h->comm->n_interior;
    out->vec_id = vec_id;
    return true;
}

unsigned long long comm_next_epoch(stan_handle *h) { return h->comm ? ++h->comm->epoch : 0; }

CommDev *comm_dev_ptr(const stan_handle *h) { return comm_p2p_active(h) ? h->comm->d_dev.p : nullptr; }

// vec holds 3*nloc owned entries, padding, and 3*n_halo halo entries from 3*nloc_pad on
int comm_halo_exchange(stan_handle *h, double *d_vec, int vec_id, cudaStream_t s, CgState *st) {
    if (h->world <= 1) return STAN_OK;
    Comm *c = h->comm;
    if (c->p2p) {
        const int gp = (int)std::min<int64_t>(std::max<int64_t>(div_up(3 * c->n_send, 256), 1), 148);
        k_halo_push<<<gp, 256, 0, s>>>(c->d_dev.p, d_vec, vec_id, st);      // push, raise flags, wait: one launch
        h->launches += 1;
        return STAN_OK;
    }
    if (c->n_send) {
        k_pack<<<div_up(3 * c->n_send, 256), 256, 0, s>>>(c->n_send, c->d_send_rows.p, d_vec, c->d_sendbuf.p);
        h->launches += 1;
    }
    STAN_NCCL(g_nccl.GroupStart());
    for (int r = 0; r < h->world; r++) {
        if (r == h->rank) continue;
        if (c->send_cnt[r])
            STAN_NCCL(g_nccl.Send(c->d_sendbuf.p + 3 * c->send_off[r], (size_t)(3 * c->send_cnt[r]), NCCL_F64, r, c->comm, s));
        if (c->recv_cnt[r])
            STAN_NCCL(g_nccl.Recv(d_vec + 3 * (h->nloc_pad + c->recv_off[r]), (size_t)(3 * c->recv_cnt[r]), NCCL_F64, r, c->comm, s));
    }
    STAN_NCCL(g_nccl.GroupEnd());
    return STAN_OK;
}

// every rank ends with the full DOF-ordered vector (rows are contiguous per rank)
int comm_allgather_rows(stan_handle *h, const double *d_local, double *d_full, cudaStream_t s) {
    Comm *c = h->comm;
    const int W = h->world;
    const int64_t nloc = h->row1 - h->row0;
    STAN_TRY(c->d_gather.alloc((size_t)(3 * c->max_rows) * (W + 1), s));
    double *stage = c->d_gather.p, *all = c->d_gather.p + 3 * c->max_rows;
    STAN_CUDA(cudaMemcpyAsync(stage, d_local, 3 * nloc * sizeof(double), cudaMemcpyDeviceToDevice, s));
    STAN_NCCL(g_nccl.AllGather(stage, all, (size_t)(3 * c->max_rows), NCCL_F64, c->comm, s));
    for (int r = 0; r < W; r++)
        STAN_CUDA(cudaMemcpyAsync(d_full + 3 * c->bound[r], all + (size_t)r * 3 * c->max_rows,
                                  3 * (c->bound[r + 1] - c->bound[r]) * sizeof(double), cudaMemcpyDeviceToDevice, s));
    return STAN_OK;
}

}  // namespace stan
