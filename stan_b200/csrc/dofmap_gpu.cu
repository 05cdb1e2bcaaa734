// Database.AssignDOF on the device (/root/reference/src/STAN_Database/Database.cs:140-234), for
// meshes whose breadth-first levels are wide enough to feed a GPU.
//
// The reference numbers nodes in FIFO discovery order; the order in which a popped node offers its
// neighbours is "incident elements in ElemLib order x NList order".  A level-synchronous traversal
// reproduces that order exactly if every node discovered in a level is ranked by the smallest
// (rank of the parent inside its level, position inside the parent's neighbour list) that reached it:
// each frontier thread proposes that key with a 64-bit atomicMin, the level's new nodes are sorted by
// key (CUB radix sort over just the bits in use) and appended to the queue.  Same numbering as
// dofmap.cpp bit for bit (tests/test_gpu_parity.py); which of the two runs is a question of
// frontier width only — a 4x4xN beam has 25-node levels and is a serial problem.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace stan {

namespace {

constexpr int POS_BITS = 12;                               // position inside a neighbour list: < 512 incident elements
constexpr unsigned long long KEY_NONE = ~0ull;

__device__ __forceinline__ bool first_in_element(const int32_t *nl, int k) {
    for (int q = 0; q < k; q++)
        if (nl[q] == nl[k]) return false;                  // RemoveElemDuplicates, Database.cs:152-158
    return true;
}

__global__ void k_ne_count(int64_t n_ent, const int32_t *__restrict__ conn, int32_t *__restrict__ cnt) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n_ent) return;
    const int k = (int)(t & 7);
    const int32_t *nl = conn + (t - k);
    if (first_in_element(nl, k)) atomicAdd(&cnt[nl[k]], 1);
}

__global__ void k_ne_fill(int64_t n_ent, const int32_t *__restrict__ conn, int32_t *__restrict__ cursor,
                          int32_t *__restrict__ idx) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n_ent) return;
    const int k = (int)(t & 7);
    const int32_t *nl = conn + (t - k);
    if (first_in_element(nl, k)) idx[atomicAdd(&cursor[nl[k]], 1)] = (int32_t)(t >> 3);
}

// ascending element order per node (the atomics above only decided positions), and the statistics
// the host needs: first node with exactly c incident elements for c = 1..6, largest list
__global__ void k_ne_sort(int64_t n_nodes, const int32_t *__restrict__ ptr, int32_t *__restrict__ idx,
                          int32_t *__restrict__ first_with, int32_t *__restrict__ max_cnt) {
    const int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (n >= n_nodes) return;
    const int s = ptr[n], e = ptr[n + 1];
    for (int i = s + 1; i < e; i++) {
        const int32_t v = idx[i];
        int j = i - 1;
        while (j >= s && idx[j] > v) { idx[j + 1] = idx[j]; j--; }
        idx[j + 1] = v;
    }
    const int c = e - s;
    if (c >= 1 && c <= 6) atomicMin(&first_with[c - 1], (int32_t)n);
    atomicMax(max_cnt, c);
}

// 8 threads per frontier node, one incident element each (strided when a node has more)
__global__ void k_bfs_expand(int lo, int hi, const int32_t *__restrict__ queue, const int32_t *__restrict__ ptr,
                             const int32_t *__restrict__ idx, const int32_t *__restrict__ conn,
                             const int32_t *__restrict__ order, unsigned long long *__restrict__ key,
                             int32_t *__restrict__ cand, int32_t *__restrict__ ncand) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int r = (int)(t >> 3), sub = (int)(t & 7);
    if (r >= hi - lo) return;
    const int32_t u = queue[lo + r];
    const int s = ptr[u], cnt = ptr[u + 1] - s;
    for (int q = sub; q < cnt; q += 8) {
        const int32_t *nl = conn + 8 * (int64_t)idx[s + q];
        for (int k = 0; k < 8; k++) {
            const int32_t v = nl[k];
            if (order[v] >= 0) continue;                   // numbered in this or an earlier level
            const unsigned long long nk = ((unsigned long long)r << POS_BITS) | (unsigned)(q * 8 + k);
            if (atomicMin(&key[v], nk) == KEY_NONE) cand[atomicAdd(ncand, 1)] = v;
        }
    }
}

__global__ void k_bfs_keys(int n, const int32_t *__restrict__ cand, const unsigned long long *__restrict__ key,
                           unsigned long long *__restrict__ ck) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) ck[i] = key[cand[i]];
}

__global__ void k_bfs_assign(int n, int base, const int32_t *__restrict__ sorted, int32_t *__restrict__ queue,
                             int32_t *__restrict__ order) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t v = sorted[i];
    queue[base + i] = v;
    order[v] = base + i;
}

__global__ void k_bfs_seed(int32_t first, int32_t *__restrict__ queue, int32_t *__restrict__ order) {
    queue[0] = first;
    order[first] = 0;
}

}  // namespace

// Returns STAN_OK with the numbering in node_index (host), or `*narrow = true` (nothing written) when
// the mesh's levels are too small for this path and the caller should run the serial traversal.
int assign_dof_device(stan_handle *h, int32_t *node_index, bool *narrow) {
    cudaStream_t s = h->stream;
    const int64_t nn = h->n_nodes, ne = h->n_elem, n_ent = 8 * ne;
    *narrow = false;
    if (n_ent >= INT32_MAX) { *narrow = true; return STAN_OK; }
    DevBuf<int32_t> cnt, ptr, idx, stats, order, queue, cand, sorted, ncand;
    DevBuf<unsigned long long> key, ck, ck_sorted;
    DevBuf<unsigned char> tmp;
    auto free_all = [&]() {
        cnt.release(s); ptr.release(s); idx.release(s); stats.release(s); order.release(s); queue.release(s);
        cand.release(s); sorted.release(s); ncand.release(s); key.release(s); ck.release(s); ck_sorted.release(s);
        tmp.release(s);
    };
    STAN_TRY(cnt.alloc(nn + 1, s)); STAN_TRY(ptr.alloc(nn + 1, s)); STAN_TRY(idx.alloc(n_ent, s));
    STAN_TRY(stats.alloc(8, s));
    STAN_CUDA(cudaMemsetAsync(cnt.p, 0, (nn + 1) * sizeof(int32_t), s));
    k_ne_count<<<div_up(n_ent, 256), 256, 0, s>>>(n_ent, h->d_conn.p, cnt.p);
    STAN_TRY(device_exclusive_scan_i32(h, cnt.p, ptr.p, nn + 1, s));
    STAN_CUDA(cudaMemcpyAsync(cnt.p, ptr.p, nn * sizeof(int32_t), cudaMemcpyDeviceToDevice, s));   // cursors
    k_ne_fill<<<div_up(n_ent, 256), 256, 0, s>>>(n_ent, h->d_conn.p, cnt.p, idx.p);
    int32_t hstats[8] = {INT32_MAX, INT32_MAX, INT32_MAX, INT32_MAX, INT32_MAX, INT32_MAX, 0, 0};
    STAN_CUDA(cudaMemcpyAsync(stats.p, hstats, sizeof hstats, cudaMemcpyHostToDevice, s));
    k_ne_sort<<<div_up(nn, 256), 256, 0, s>>>(nn, ptr.p, idx.p, stats.p, stats.p + 6);
    STAN_CUDA(cudaMemcpyAsync(hstats, stats.p, sizeof hstats, cudaMemcpyDeviceToHost, s));
    STAN_CUDA(cudaStreamSynchronize(s));
    h->launches += 5;
    int32_t first = -1;
    for (int c = 0; c < 6 && first < 0; c++)               // Database.cs:178-196
        if (hstats[c] != INT32_MAX) first = hstats[c];
    if (first < 0) {
        free_all();
        set_error("AssignDOF: no node with 1..6 incident elements (reference would look up node 0 and throw)");
        return STAN_E_DOFMAP;
    }
    if (hstats[6] * 8 >= (1 << POS_BITS)) { free_all(); *narrow = true; return STAN_OK; }

    STAN_TRY(order.alloc(nn, s)); STAN_TRY(queue.alloc(nn, s)); STAN_TRY(cand.alloc(nn, s)); STAN_TRY(sorted.alloc(nn, s));
    STAN_TRY(ncand.alloc(1, s)); STAN_TRY(key.alloc(nn, s)); STAN_TRY(ck.alloc(nn, s)); STAN_TRY(ck_sorted.alloc(nn, s));
    size_t tmp_bytes = 0;
    STAN_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, ck.p, ck_sorted.p, cand.p, sorted.p, (int)nn, 0, 64, s));
    STAN_TRY(tmp.alloc(tmp_bytes, s));
    STAN_CUDA(cudaMemsetAsync(order.p, 0xff, nn * sizeof(int32_t), s));
    STAN_CUDA(cudaMemsetAsync(key.p, 0xff, nn * sizeof(unsigned long long), s));
    STAN_CUDA(cudaMemsetAsync(ncand.p, 0, sizeof(int32_t), s));
    k_bfs_seed<<<1, 1, 0, s>>>(first, queue.p, order.p);
    int lo = 0, hi = 1, levels = 0;
    int64_t launches = 1;
    while (true) {
        const int width = hi - lo;
        k_bfs_expand<<<div_up(8 * (int64_t)width, 256), 256, 0, s>>>(lo, hi, queue.p, ptr.p, idx.p, h->d_conn.p, order.p,
                                                                     key.p, cand.p, ncand.p);
        int32_t n_new = 0;
        STAN_CUDA(cudaMemcpyAsync(&n_new, ncand.p, sizeof n_new, cudaMemcpyDeviceToHost, s));
        STAN_CUDA(cudaStreamSynchronize(s));
        launches++;
        if (n_new == 0) break;
        int rank_bits = 1;
        while ((1ll << rank_bits) < width) rank_bits++;
        k_bfs_keys<<<div_up(n_new, 256), 256, 0, s>>>(n_new, cand.p, key.p, ck.p);
        size_t bytes = tmp_bytes;
        STAN_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, ck.p, ck_sorted.p, cand.p, sorted.p, n_new, 0,
                                                  POS_BITS + rank_bits, s));
        k_bfs_assign<<<div_up(n_new, 256), 256, 0, s>>>(n_new, hi, sorted.p, queue.p, order.p);
        STAN_CUDA(cudaMemsetAsync(ncand.p, 0, sizeof(int32_t), s));
        launches += 4;
        lo = hi;
        hi += n_new;
        levels++;
        if (levels == 64 && hi < 64 * 512 && hi < nn) {    // mean level below 512 nodes: a serial problem
            free_all();
            h->launches += launches;
            *narrow = true;
            return STAN_OK;
        }
    }
    STAN_CUDA(cudaGetLastError());
    h->launches += launches;
    if (hi != nn) {
        free_all();
        set_error("AssignDOF: mesh is disconnected (%d of %lld nodes reached; the reference runs off its queue)", hi,
                  (long long)nn);
        return STAN_E_DOFMAP;
    }
    STAN_CUDA(cudaMemcpyAsync(node_index, order.p, nn * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    STAN_CUDA(cudaStreamSynchronize(s));
    free_all();
    return STAN_OK;
}

}  // namespace stan
