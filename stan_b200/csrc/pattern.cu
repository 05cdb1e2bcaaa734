// Sparsity pattern and SPC bookkeeping of the global stiffness matrix.
//
// Replaces, on the device: the implicit pattern that alglib.sparseadd builds one locked hash
// insert at a time (/root/reference/src/STAN_Solver/SolverFunctions.cs:143-172), the
// nDOF_reduction map and the right-hand side (/root/reference/src/STAN_Solver/Solver.cs:104-152).
//
// Layout: rows are nodes in BFS (DOF) order.  Block row p lists the nodes q that
// share an element with p, ascending; each (p,q) is a 3x3 block because the three DOFs of a node
// are consecutive (Node.cs:218-223).  SPC-fixed DOFs stay in the system as identity rows, which
// leaves the CG iterates of the free DOFs unchanged and keeps the 3x3 structure intact; the
// reference's reduced upper-triangle CRS is produced on demand by export_csr_upper().
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <unordered_map>

#include "common.cuh"

namespace stan {

namespace {

__global__ void k_invert_perm(int64_t n, const int32_t *__restrict__ node_index, int32_t *__restrict__ inv,
                              int32_t *err) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    int32_t p = node_index[i];
    if (p < 0 || p >= n) { atomicOr(err, 2); return; }
    inv[p] = (int32_t)i;
}

// incidence count: one thread per (element, local node)
__global__ void k_inc_count(int64_t n_ent, const int32_t *__restrict__ conn, const int32_t *__restrict__ node_index,
                            int64_t row0, int64_t nloc, int32_t *__restrict__ cnt) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n_ent) return;
    int64_t p = (int64_t)node_index[conn[t]] - row0;
    if (p >= 0 && p < nloc) atomicAdd(&cnt[p], 1);
}

__global__ void k_inc_fill(int64_t n_ent, const int32_t *__restrict__ conn, const int32_t *__restrict__ node_index,
                           int64_t row0, int64_t nloc, int32_t *__restrict__ cursor, int32_t *__restrict__ inc) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n_ent) return;
    int64_t p = (int64_t)node_index[conn[t]] - row0;
    if (p >= 0 && p < nloc) inc[atomicAdd(&cursor[p], 1)] = (int32_t)t;   // t = elem*8 + local node
}

// Atomics above leave each row's entries in arbitrary order; ascending (element, local node)
// order restores determinism — contributions are later summed in ElemLib order.
__global__ void k_inc_sort(int64_t nloc, const int32_t *__restrict__ ptr, int32_t *__restrict__ inc) {
    int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= nloc) return;
    int s = ptr[p], e = ptr[p + 1];
    for (int i = s + 1; i < e; i++) {
        int32_t v = inc[i];
        int j = i - 1;
        while (j >= s && inc[j] > v) { inc[j + 1] = inc[j]; j--; }
        inc[j + 1] = v;
    }
}

// Sorted unique neighbour list of a row (the node itself included).  FILL = false counts.
// Rows that couple to at most ROW_FAST nodes (every structured hex mesh: 27) are built in registers /
// local memory by one thread.  A wider row — the reference's hash-table matrix takes any valence
// (SolverFunctions.cs:123, 162-165) — is appended to `wide_rows` by the count pass and built by
// k_wide_rows in a global scratch list of 8 candidates per incident element; the fill pass skips it.
constexpr int ROW_FAST = 96;

template <bool FILL>
__global__ void k_row_neighbors(int64_t nloc, const int32_t *__restrict__ inc_ptr, const int32_t *__restrict__ inc,
                                const int32_t *__restrict__ conn, const int32_t *__restrict__ node_index,
                                int32_t *__restrict__ cnt, const int32_t *__restrict__ brow_ptr,
                                int32_t *__restrict__ bcol, int32_t *__restrict__ wide_rows, int32_t *n_wide,
                                const uint8_t *__restrict__ slow_flag) {
    int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= nloc) return;
    if (!slow_flag[p]) return;                        // this row was handled by the warp-per-row kernels
    int32_t nb[ROW_FAST];
    int n = 0;
    bool overflow = false;
    for (int t = inc_ptr[p]; t < inc_ptr[p + 1]; t++) {
        const int32_t *nl = conn + 8 * (int64_t)(inc[t] >> 3);
        for (int k = 0; k < 8; k++) {
            int32_t q = node_index[nl[k]];
            int lo = 0, hi = n;                       // lower bound in nb[0..n)
            while (lo < hi) {
                int mid = (lo + hi) >> 1;
                if (nb[mid] < q) lo = mid + 1; else hi = mid;
            }
            if (lo < n && nb[lo] == q) continue;
            if (n == ROW_FAST) { overflow = true; continue; }
            for (int j = n; j > lo; j--) nb[j] = nb[j - 1];
            nb[lo] = q;
            n++;
        }
    }
    if (overflow) {                                   // list order is irrelevant: every wide row is built on its own
        if (!FILL) { cnt[p] = 0; wide_rows[atomicAdd(n_wide, 1)] = (int32_t)p; }
        return;
    }
    if (!FILL) cnt[p] = n;
    else {
        int32_t *out = bcol + brow_ptr[p];
        for (int j = 0; j < n; j++) out[j] = nb[j];
    }
}

// Warp-per-row version of the count pass for rows with at most 8 incident elements (every row of a structured
// hex mesh): the 64 candidate columns are fetched by 32 lanes at once (three dependent loads in total instead of
// three per candidate), sorted with a 64-key bitonic network in registers, and the unique ones are written to a
// 32-column staging row so that the fill pass is a copy.  Rows with more incident elements, or more than 32
// distinct columns, are flagged in slow[] and left to k_row_neighbors.
__device__ __forceinline__ void cmp_swap(int32_t &a, int32_t &b, bool up) {
    const int32_t lo = min(a, b), hi = max(a, b);
    a = up ? lo : hi; b = up ? hi : lo;
}

__global__ void __launch_bounds__(256)
k_row_neighbors_warp(int64_t nloc, const int32_t *__restrict__ inc_ptr, const int32_t *__restrict__ inc,
                     const int32_t *__restrict__ conn, const int32_t *__restrict__ node_index,
                     int32_t *__restrict__ cnt, int32_t *__restrict__ staged, uint8_t *__restrict__ slow) {
    const int lane = threadIdx.x & 31;
    const int64_t p = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (p >= nloc) return;
    const int t0 = inc_ptr[p], n_ent = inc_ptr[p + 1] - t0;
    if (n_ent > 8) {
        if (lane == 0) { cnt[p] = 0; slow[p] = 1; }
        return;
    }
    constexpr int32_t NONE = 0x7fffffff;
    int32_t a = NONE, b = NONE;                        // keys lane and 32 + lane: entry (lane / 8 [+ 4]), element node lane % 8
    {
        const int ta = lane >> 3, k = lane & 7;
        if (ta < n_ent) a = node_index[conn[8 * (int64_t)(inc[t0 + ta] >> 3) + k]];
        if (ta + 4 < n_ent) b = node_index[conn[8 * (int64_t)(inc[t0 + ta + 4] >> 3) + k]];
    }
    // bitonic sort of 64 keys, ascending; key index = lane for a, 32 + lane for b
#pragma unroll
    for (int size = 2; size <= 64; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            if (stride == 32) {
                cmp_swap(a, b, true);                  // size 64: the whole sequence ascending
            } else {
                const int32_t oa = __shfl_xor_sync(0xffffffffu, a, stride), ob = __shfl_xor_sync(0xffffffffu, b, stride);
                const bool lower = (lane & stride) == 0;
                const bool up_a = (lane & size) == 0 || size == 64, up_b = ((32 + lane) & size) == 0 || size == 64;
                a = (lower == up_a) ? min(a, oa) : max(a, oa);
                b = (lower == up_b) ? min(b, ob) : max(b, ob);
            }
        }
    }
    // unique: a key counts when it differs from its predecessor in sorted order
    const int32_t pa = __shfl_up_sync(0xffffffffu, a, 1), a31 = __shfl_sync(0xffffffffu, a, 31);
    int32_t pb = __shfl_up_sync(0xffffffffu, b, 1);
    if (lane == 0) pb = a31;
    const bool fa = a != NONE && (lane == 0 || a != pa), fb = b != NONE && b != pb;
    const unsigned ma = __ballot_sync(0xffffffffu, fa), mb = __ballot_sync(0xffffffffu, fb);
    const int na = __popc(ma), n = na + __popc(mb);
    if (n > 32) {
        if (lane == 0) { cnt[p] = 0; slow[p] = 1; }
        return;
    }
    const unsigned lt = (1u << lane) - 1;
    if (fa) staged[32 * p + __popc(ma & lt)] = a;
    if (fb) staged[32 * p + na + __popc(mb & lt)] = b;
    if (lane == 0) { cnt[p] = n; slow[p] = 0; }
}

// fill pass for the rows k_row_neighbors_warp staged
__global__ void k_copy_staged(int64_t nloc, const int32_t *__restrict__ staged, const uint8_t *__restrict__ slow,
                              const int32_t *__restrict__ brow_ptr, int32_t *__restrict__ bcol) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t p = t >> 5;
    if (p >= nloc || slow[p]) return;
    const int j = (int)(t & 31), s0 = brow_ptr[p];
    if (j < brow_ptr[p + 1] - s0) bcol[s0 + j] = staged[t];
}

__global__ void k_wide_caps(int n_wide, const int32_t *__restrict__ wide_rows, const int32_t *__restrict__ inc_ptr,
                            int32_t *__restrict__ cap) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n_wide) return;
    cap[i] = i < n_wide ? 8 * (inc_ptr[wide_rows[i] + 1] - inc_ptr[wide_rows[i]]) : 0;
}

// One warp per wide row.  STAGE 0: gather the 8 nodes of every incident element into the row's scratch
// segment, sort (odd-even transposition over the warp: the segment is a few hundred entries), drop
// duplicates in place and publish the count.  STAGE 1 (after the row-pointer scan): copy to bcol.
template <int STAGE>
__global__ void k_wide_rows(int n_wide, const int32_t *__restrict__ wide_rows, const int32_t *__restrict__ off,
                            const int32_t *__restrict__ inc_ptr, const int32_t *__restrict__ inc,
                            const int32_t *__restrict__ conn, const int32_t *__restrict__ node_index,
                            int32_t *__restrict__ scratch, int32_t *__restrict__ cnt,
                            const int32_t *__restrict__ brow_ptr, int32_t *__restrict__ bcol) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= n_wide) return;
    const int32_t p = wide_rows[w];
    int32_t *seg = scratch + off[w];
    if (STAGE == 1) {
        const int n = cnt[p];
        for (int j = lane; j < n; j += 32) bcol[brow_ptr[p] + j] = seg[j];
        return;
    }
    const int t0 = inc_ptr[p], m = 8 * (inc_ptr[p + 1] - t0);
    for (int j = lane; j < m; j += 32) seg[j] = node_index[conn[8 * (int64_t)(inc[t0 + (j >> 3)] >> 3) + (j & 7)]];
    __syncwarp();
    for (int round = 0; round < m; round++) {         // odd-even transposition sort, m rounds
        for (int j = (round & 1) + 2 * lane; j + 1 < m; j += 64) {
            const int32_t a = seg[j], b = seg[j + 1];
            if (a > b) { seg[j] = b; seg[j + 1] = a; }
        }
        __syncwarp();
    }
    if (lane == 0) {                                  // unique in place (sequential: runs once per wide row)
        int n = 0;
        for (int j = 0; j < m; j++)
            if (n == 0 || seg[n - 1] != seg[j]) seg[n++] = seg[j];
        cnt[p] = n;
    }
}

__global__ void k_gather_i32(int64_t n, const int32_t *__restrict__ idx, const int32_t *__restrict__ src,
                             int32_t *__restrict__ out) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t < n) out[t] = src[idx[t]];
}

__global__ void k_group_max(int64_t nloc, int rows_per_group, const int32_t *__restrict__ brow_ptr, int32_t *out) {
    int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t r0 = g * rows_per_group;
    if (r0 >= nloc) return;
    int64_t r1 = r0 + rows_per_group < nloc ? r0 + rows_per_group : nloc;
    atomicMax(out, brow_ptr[r1] - brow_ptr[r0]);
}

// Solver.cs:106-114: a DOF is fixed when an SPC entry holds exactly 1 in that direction
__global__ void k_mark_fixed(int64_t n_spc, const int32_t *__restrict__ node, const double *__restrict__ val,
                             const int32_t *__restrict__ node_index, uint8_t *__restrict__ fixed) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= 3 * n_spc) return;
    if (val[t] == 1.0) fixed[3 * (int64_t)node_index[node[t / 3]] + t % 3] = 1;
}

__global__ void k_fixed_to_int(int64_t n, const uint8_t *__restrict__ fixed, int32_t *__restrict__ out) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = fixed[i];
}

// Solver.cs:121-132: -1 at fixed DOFs, else the number of fixed DOFs below
__global__ void k_finish_reduction(int64_t n, const uint8_t *__restrict__ fixed, int32_t *__restrict__ red) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n && fixed[i]) red[i] = -1;
}

__global__ void k_scatter_rhs(int64_t n, const int64_t *__restrict__ dof, const double *__restrict__ val,
                              const uint8_t *__restrict__ fixed, int64_t dof0, int64_t ndof_loc,
                              double *__restrict__ b) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n) return;
    int64_t d = dof[t];
    if (fixed[d]) return;                              // Solver.cs:142
    int64_t l = d - dof0;
    if (l >= 0 && l < ndof_loc) b[l] = val[t];
}

// ---- reduced upper-triangle CRS export -------------------------------------------------------
template <bool FILL>
__global__ void k_upper_rows(int64_t nloc, int64_t row0, const int32_t *__restrict__ brow_ptr,
                             const int32_t *__restrict__ bcol, const double *__restrict__ vals,
                             const uint8_t *__restrict__ fixed, const int32_t *__restrict__ red,
                             int64_t *__restrict__ cnt, const int64_t *__restrict__ rowptr,
                             int32_t *__restrict__ col, double *__restrict__ val) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= 3 * nloc) return;
    int64_t pl = t / 3;
    int a = (int)(t % 3);
    int64_t d = 3 * (row0 + pl) + a;
    if (fixed[d]) return;
    int64_t r = d - red[d];
    int s0 = brow_ptr[pl], s1 = brow_ptr[pl + 1], nb = s1 - s0;
    const double *rowv = vals + 9 * (int64_t)s0 + (int64_t)a * 3 * nb;
    int64_t w = FILL ? rowptr[r] : 0;
    for (int s = s0; s < s1; s++) {
        int64_t q = bcol[s];
        for (int b = 0; b < 3; b++) {
            int64_t c = 3 * q + b;
            if (c < d || fixed[c]) continue;
            if (FILL) { col[w] = (int32_t)(c - red[c]); val[w] = rowv[3 * (s - s0) + b]; }
            w++;
        }
    }
    if (!FILL) cnt[r] = w;
}

template <typename T>
int exclusive_scan(stan_handle *h, const T *in, T *out, int64_t n, cudaStream_t s) {
    size_t bytes = 0;
    STAN_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, n, s));
    ScratchBuf<unsigned char> tmp(&h->scratch[8]);
    STAN_TRY(tmp.alloc(bytes ? bytes : 1, s));
    STAN_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, bytes, in, out, n, s));
    return STAN_OK;
}

}  // namespace

int device_exclusive_scan_i32(stan_handle *h, const int32_t *in, int32_t *out, int64_t n, cudaStream_t s) {
    return exclusive_scan(h, in, out, n, s);
}

int build_system_pattern(stan_handle *h) {
    cudaStream_t s = h->stream;
    const int T = 256;
    const int64_t nn = h->n_nodes, ne = h->n_elem, nloc = h->row1 - h->row0;
    STAN_TRY(h->d_err.alloc(8, s));
    STAN_CUDA(cudaMemsetAsync(h->d_err.p, 0, 8 * sizeof(int32_t), s));

    STAN_TRY(h->d_inv.alloc(nn, s));
    STAN_CUDA(cudaMemsetAsync(h->d_inv.p, 0xff, nn * sizeof(int32_t), s));
    k_invert_perm<<<div_up(nn, T), T, 0, s>>>(nn, h->d_node_index.p, h->d_inv.p, h->d_err.p);

    // incidence of the owned rows
    STAN_TRY(h->d_inc_ptr.alloc(nloc + 1, s));
    ScratchBuf<int32_t> cnt(&h->scratch[0]);
    STAN_TRY(cnt.alloc(nloc + 1, s));
    STAN_CUDA(cudaMemsetAsync(cnt.p, 0, (nloc + 1) * sizeof(int32_t), s));
    k_inc_count<<<div_up(8 * ne, T), T, 0, s>>>(8 * ne, h->d_conn.p, h->d_node_index.p, h->row0, nloc, cnt.p);
    STAN_TRY(exclusive_scan(h, cnt.p, h->d_inc_ptr.p, nloc + 1, s));
    int32_t n_inc = 0;
    STAN_CUDA(cudaMemcpyAsync(&n_inc, h->d_inc_ptr.p + nloc, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    STAN_CUDA(cudaStreamSynchronize(s));
    STAN_TRY(h->d_inc.alloc(n_inc, s));
    STAN_CUDA(cudaMemcpyAsync(cnt.p, h->d_inc_ptr.p, nloc * sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
    k_inc_fill<<<div_up(8 * ne, T), T, 0, s>>>(8 * ne, h->d_conn.p, h->d_node_index.p, h->row0, nloc, cnt.p, h->d_inc.p);
    k_inc_sort<<<div_up(nloc, T), T, 0, s>>>(nloc, h->d_inc_ptr.p, h->d_inc.p);

    // block rows
    STAN_TRY(h->d_brow_ptr.alloc(nloc + 1 + 4, s));      // +4: the bulk-copy SpMV reads 16-byte granules
    STAN_CUDA(cudaMemsetAsync(cnt.p, 0, (nloc + 1) * sizeof(int32_t), s));
    ScratchBuf<int32_t> wide(&h->scratch[2]);            // rows wider than ROW_FAST (none on structured meshes)
    STAN_TRY(wide.alloc(nloc + 1, s));
    int32_t *n_wide_d = h->d_err.p + 6;                  // zeroed with the error flags above
    // rows with <= 8 incident elements and <= 32 columns: one warp each, staged; the rest: one thread each
    ScratchBuf<int32_t> staged(&h->scratch[11]);
    ScratchBuf<uint8_t> slow(&h->scratch[9]);
    STAN_TRY(staged.alloc((size_t)32 * nloc, s)); STAN_TRY(slow.alloc(nloc, s));
    k_row_neighbors_warp<<<div_up(32 * nloc, 256), 256, 0, s>>>(nloc, h->d_inc_ptr.p, h->d_inc.p, h->d_conn.p,
                                                               h->d_node_index.p, cnt.p, staged.p, slow.p);
    k_row_neighbors<false><<<div_up(nloc, 128), 128, 0, s>>>(nloc, h->d_inc_ptr.p, h->d_inc.p, h->d_conn.p,
                                                             h->d_node_index.p, cnt.p, nullptr, nullptr, wide.p, n_wide_d,
                                                             slow.p);
    STAN_TRY(exclusive_scan(h, cnt.p, h->d_brow_ptr.p, nloc + 1, s));
    int32_t nblk = 0, herr[8];
    STAN_CUDA(cudaMemcpyAsync(&nblk, h->d_brow_ptr.p + nloc, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    STAN_CUDA(cudaMemcpyAsync(herr, h->d_err.p, sizeof herr, cudaMemcpyDeviceToHost, s));
    STAN_CUDA(cudaStreamSynchronize(s));
    if (herr[0] & 2) { set_error("dof map is not a permutation of 0..n_nodes-1"); cnt.release(s); return STAN_E_ARG; }
    const int n_wide = herr[6];
    ScratchBuf<int32_t> wcap(&h->scratch[3]), woff(&h->scratch[4]), wscr(&h->scratch[5]);
    if (n_wide > 0) {
        STAN_TRY(wcap.alloc(n_wide + 1, s)); STAN_TRY(woff.alloc(n_wide + 1, s));
        k_wide_caps<<<div_up(n_wide + 1, T), T, 0, s>>>(n_wide, wide.p, h->d_inc_ptr.p, wcap.p);
        STAN_TRY(exclusive_scan(h, wcap.p, woff.p, n_wide + 1, s));
        int32_t total = 0;
        STAN_CUDA(cudaMemcpyAsync(&total, woff.p + n_wide, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        STAN_CUDA(cudaStreamSynchronize(s));
        STAN_TRY(wscr.alloc(total, s));
        k_wide_rows<0><<<div_up(32 * (int64_t)n_wide, 128), 128, 0, s>>>(n_wide, wide.p, woff.p, h->d_inc_ptr.p, h->d_inc.p,
                                                                        h->d_conn.p, h->d_node_index.p, wscr.p, cnt.p,
                                                                        nullptr, nullptr);
        STAN_TRY(exclusive_scan(h, cnt.p, h->d_brow_ptr.p, nloc + 1, s));
        STAN_CUDA(cudaMemcpyAsync(&nblk, h->d_brow_ptr.p + nloc, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        STAN_CUDA(cudaStreamSynchronize(s));
        h->launches += 4;
    }
    h->n_blocks = nblk;
    STAN_TRY(h->d_bcol.alloc((size_t)nblk + 4, s));
    k_copy_staged<<<div_up(32 * nloc, 256), 256, 0, s>>>(nloc, staged.p, slow.p, h->d_brow_ptr.p, h->d_bcol.p);
    k_row_neighbors<true><<<div_up(nloc, 128), 128, 0, s>>>(nloc, h->d_inc_ptr.p, h->d_inc.p, h->d_conn.p,
                                                            h->d_node_index.p, nullptr, h->d_brow_ptr.p, h->d_bcol.p,
                                                            nullptr, nullptr, slow.p);
    h->launches += 2;
    if (n_wide > 0) {
        k_wide_rows<1><<<div_up(32 * (int64_t)n_wide, 128), 128, 0, s>>>(n_wide, wide.p, woff.p, h->d_inc_ptr.p, h->d_inc.p,
                                                                        h->d_conn.p, h->d_node_index.p, wscr.p, cnt.p,
                                                                        h->d_brow_ptr.p, h->d_bcol.p);
        h->launches += 1;
    }
    wide.release(s); wcap.release(s); woff.release(s); wscr.release(s); staged.release(s); slow.release(s);
    k_group_max<<<div_up(div_up(nloc, 32), T), T, 0, s>>>(nloc, 32, h->d_brow_ptr.p, h->d_err.p + 1);
    k_group_max<<<div_up(div_up(nloc, 16), T), T, 0, s>>>(nloc, 16, h->d_brow_ptr.p, h->d_err.p + 3);
    cnt.release(s);

    // SPC flags and nDOF_reduction over the global DOF range
    const int64_t ndof = 3 * nn;
    STAN_TRY(h->d_fixed.alloc(ndof, s));
    STAN_TRY(h->d_red.alloc(ndof + 1, s));
    STAN_CUDA(cudaMemsetAsync(h->d_fixed.p, 0, ndof, s));
    const int64_t nspc = (int64_t)h->h_spc_node.size();
    if (nspc) {
        ScratchBuf<int32_t> dn(&h->scratch[2]); ScratchBuf<double> dv(&h->scratch[3]);
        STAN_TRY(dn.alloc(nspc, s)); STAN_TRY(dv.alloc(3 * nspc, s));
        STAN_CUDA(cudaMemcpyAsync(dn.p, h->h_spc_node.data(), nspc * sizeof(int32_t), cudaMemcpyHostToDevice, s));
        STAN_CUDA(cudaMemcpyAsync(dv.p, h->h_spc_val.data(), 3 * nspc * sizeof(double), cudaMemcpyHostToDevice, s));
        k_mark_fixed<<<div_up(3 * nspc, T), T, 0, s>>>(nspc, dn.p, dv.p, h->d_node_index.p, h->d_fixed.p);
        dn.release(s); dv.release(s);
    }
    {
        ScratchBuf<int32_t> tmp(&h->scratch[1]);
        STAN_TRY(tmp.alloc(ndof + 1, s));
        STAN_CUDA(cudaMemsetAsync(tmp.p + ndof, 0, sizeof(int32_t), s));
        k_fixed_to_int<<<div_up(ndof, T), T, 0, s>>>(ndof, h->d_fixed.p, tmp.p);
        STAN_TRY(exclusive_scan(h, tmp.p, h->d_red.p, ndof + 1, s));
        tmp.release(s);
        int32_t nfix = 0;
        STAN_CUDA(cudaMemcpyAsync(&nfix, h->d_red.p + ndof, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        k_finish_reduction<<<div_up(ndof, T), T, 0, s>>>(ndof, h->d_fixed.p, h->d_red.p);
        STAN_CUDA(cudaMemcpyAsync(herr, h->d_err.p, sizeof herr, cudaMemcpyDeviceToHost, s));
        STAN_CUDA(cudaStreamSynchronize(s));
        h->n_fixed = nfix;
        h->max_group_blocks = herr[1];
        h->max_group16 = herr[3];
    }
    h->launches += 12;
    return STAN_OK;
}

// F (Solver.cs:136-152).  Loads are accumulated on the host in list order (the += of the
// reference), one value per DOF is uploaded and scattered into the full-space vector.
int build_rhs(stan_handle *h) {
    cudaStream_t s = h->stream;
    const int64_t nloc = h->row1 - h->row0;
    STAN_TRY(h->d_b.alloc(3 * nloc, s));
    STAN_CUDA(cudaMemsetAsync(h->d_b.p, 0, 3 * nloc * sizeof(double), s));
    const int64_t nl = (int64_t)h->h_load_node.size();
    if (!nl) return STAN_OK;
    std::vector<int32_t> h_node_index((size_t)nl);        // DOF index of the loaded nodes only, position i <-> load entry i
    {
        ScratchBuf<int32_t> dn(&h->scratch[2]), di(&h->scratch[3]);
        STAN_TRY(dn.alloc(nl, s)); STAN_TRY(di.alloc(nl, s));
        STAN_CUDA(cudaMemcpyAsync(dn.p, h->h_load_node.data(), nl * sizeof(int32_t), cudaMemcpyHostToDevice, s));
        k_gather_i32<<<div_up(nl, 256), 256, 0, s>>>(nl, dn.p, h->d_node_index.p, di.p);
        STAN_CUDA(cudaMemcpyAsync(h_node_index.data(), di.p, nl * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        STAN_CUDA(cudaStreamSynchronize(s));
        h->launches += 1;
    }
    std::unordered_map<int64_t, size_t> slot;
    std::vector<int64_t> dofs;
    std::vector<double> vals;
    slot.reserve((size_t)nl * 3);
    for (int64_t i = 0; i < nl; i++)
        for (int d = 0; d < 3; d++) {
            int64_t dof = 3 * (int64_t)h_node_index[i] + d;
            auto it = slot.find(dof);
            if (it == slot.end()) { slot.emplace(dof, dofs.size()); dofs.push_back(dof); vals.push_back(0.0 + h->h_load_val[3 * i + d]); }
            else vals[it->second] += h->h_load_val[3 * i + d];
        }
    ScratchBuf<int64_t> dd(&h->scratch[2]); ScratchBuf<double> dv(&h->scratch[3]);
    STAN_TRY(dd.alloc(dofs.size(), s)); STAN_TRY(dv.alloc(vals.size(), s));
    STAN_CUDA(cudaMemcpyAsync(dd.p, dofs.data(), dofs.size() * sizeof(int64_t), cudaMemcpyHostToDevice, s));
    STAN_CUDA(cudaMemcpyAsync(dv.p, vals.data(), vals.size() * sizeof(double), cudaMemcpyHostToDevice, s));
    k_scatter_rhs<<<div_up((int64_t)dofs.size(), 256), 256, 0, s>>>((int64_t)dofs.size(), dd.p, dv.p, h->d_fixed.p,
                                                                    3 * h->row0, 3 * nloc, h->d_b.p);
    STAN_CUDA(cudaStreamSynchronize(s));   // host staging vectors go out of scope
    dd.release(s); dv.release(s);
    h->launches += 1;
    return STAN_OK;
}

static int upper_counts(stan_handle *h, DevBuf<int64_t> &rowptr, int64_t *n_out, int64_t *nnz_out) {
    cudaStream_t s = h->stream;
    const int64_t nloc = h->row1 - h->row0;
    const int64_t n = 3 * h->n_nodes - h->n_fixed;
    DevBuf<int64_t> cnt;
    STAN_TRY(cnt.alloc(n + 1, s));
    STAN_TRY(rowptr.alloc(n + 1, s));
    STAN_CUDA(cudaMemsetAsync(cnt.p, 0, (n + 1) * sizeof(int64_t), s));
    k_upper_rows<false><<<div_up(3 * nloc, 128), 128, 0, s>>>(nloc, h->row0, h->d_brow_ptr.p, h->d_bcol.p, h->d_vals.p,
                                                              h->d_fixed.p, h->d_red.p, cnt.p, nullptr, nullptr, nullptr);
    STAN_TRY(exclusive_scan(h, cnt.p, rowptr.p, n + 1, s));
    int64_t nnz = 0;
    STAN_CUDA(cudaMemcpyAsync(&nnz, rowptr.p + n, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    STAN_CUDA(cudaStreamSynchronize(s));
    cnt.release(s);
    *n_out = n;
    *nnz_out = nnz;
    return STAN_OK;
}

int export_csr_upper_size(stan_handle *h, int64_t *n, int64_t *nnz) {
    if (h->world != 1) { set_error("csr export is single-GPU only"); return STAN_E_STATE; }
    DevBuf<int64_t> rowptr;
    int rc = upper_counts(h, rowptr, n, nnz);
    rowptr.release(h->stream);
    if (rc == STAN_OK) h->nnz_upper = *nnz;
    return rc;
}

int export_csr_upper(stan_handle *h, int64_t *rowptr_out, int32_t *col_out, double *val_out) {
    if (h->world != 1) { set_error("csr export is single-GPU only"); return STAN_E_STATE; }
    cudaStream_t s = h->stream;
    const int64_t nloc = h->row1 - h->row0;
    DevBuf<int64_t> rowptr;
    int64_t n, nnz;
    STAN_TRY(upper_counts(h, rowptr, &n, &nnz));
    DevBuf<int32_t> col; DevBuf<double> val;
    STAN_TRY(col.alloc(nnz, s)); STAN_TRY(val.alloc(nnz, s));
    k_upper_rows<true><<<div_up(3 * nloc, 128), 128, 0, s>>>(nloc, h->row0, h->d_brow_ptr.p, h->d_bcol.p, h->d_vals.p,
                                                             h->d_fixed.p, h->d_red.p, nullptr, rowptr.p, col.p, val.p);
    if (rowptr_out) STAN_CUDA(cudaMemcpyAsync(rowptr_out, rowptr.p, (n + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    if (col_out) STAN_CUDA(cudaMemcpyAsync(col_out, col.p, nnz * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    if (val_out) STAN_CUDA(cudaMemcpyAsync(val_out, val.p, nnz * sizeof(double), cudaMemcpyDeviceToHost, s));
    STAN_CUDA(cudaStreamSynchronize(s));
    rowptr.release(s); col.release(s); val.release(s);
    return STAN_OK;
}

}  // namespace stan
