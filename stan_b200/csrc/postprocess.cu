// Post-processing scalars: the 24 result fields PrePost derives from a solved database
// (Part.Load_Scalar, /root/reference/src/STAN_Database/Part.cs:231-528) — SURVEY.md §8f row 3.
//
// Per (element, element node): displacement x/y/z/total, the six stress components, the principal
// stresses P1 >= P2 >= P3 (eigenvalues of the symmetric tensor, MathNet Evd in the reference), von
// Mises stress sqrt(((P1-P2)^2+(P2-P3)^2+(P3-P1)^2)/2), the six strain components (the engineering
// shear values are used as off-diagonals exactly as the reference does), principal strains and
// effective strain (2/3 of the same expression).  Cell data = max / average / min over the 8 nodes
// of an element (Part.cs:383-391); point data = average over the elements that contain the node
// of the per-element values (Part.cs:431-519; principal values are averaged, not recomputed).
// Stored as float32 like the reference's vtkFloatArray.
#include <cfloat>

#include "common.cuh"

namespace stan {

namespace {

constexpr int NS = 24;

// eigenvalues of a symmetric 3x3 matrix, descending (closed form, FP64)
__device__ __forceinline__ void eig3(double a00, double a11, double a22, double a01, double a12, double a02, double &e1,
                                     double &e2, double &e3) {
    const double p1 = a01 * a01 + a02 * a02 + a12 * a12;
    if (p1 == 0.0) {
        double x = a00, y = a11, z = a22, t;
        if (x < y) { t = x; x = y; y = t; }
        if (y < z) { t = y; y = z; z = t; }
        if (x < y) { t = x; x = y; y = t; }
        e1 = x; e2 = y; e3 = z;
        return;
    }
    const double q = (a00 + a11 + a22) / 3.0;
    const double b00 = a00 - q, b11 = a11 - q, b22 = a22 - q;
    const double p2 = b00 * b00 + b11 * b11 + b22 * b22 + 2.0 * p1;
    const double p = sqrt(p2 / 6.0);
    const double ip = 1.0 / p;
    const double c00 = b00 * ip, c11 = b11 * ip, c22 = b22 * ip, c01 = a01 * ip, c12 = a12 * ip, c02 = a02 * ip;
    double r = 0.5 * (c00 * (c11 * c22 - c12 * c12) - c01 * (c01 * c22 - c12 * c02) + c02 * (c01 * c12 - c11 * c02));
    r = r < -1.0 ? -1.0 : (r > 1.0 ? 1.0 : r);
    const double phi = acos(r) / 3.0;
    e1 = q + 2.0 * p * cos(phi);
    e3 = q + 2.0 * p * cos(phi + 2.0943951023931954923);   // + 2 pi / 3
    e2 = 3.0 * q - e1 - e3;
}

__device__ __forceinline__ void node_scalars(const double *__restrict__ u, const double *__restrict__ sig,
                                             const double *__restrict__ eps, double (&v)[NS]) {
    v[0] = u[0]; v[1] = u[1]; v[2] = u[2];
    v[3] = sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
    double p1, p2, p3;
#pragma unroll
    for (int c = 0; c < 6; c++) v[4 + c] = sig[c];
    eig3(sig[0], sig[1], sig[2], sig[3], sig[4], sig[5], p1, p2, p3);
    v[10] = p1; v[11] = p2; v[12] = p3;
    v[13] = sqrt(((p1 - p2) * (p1 - p2) + (p2 - p3) * (p2 - p3) + (p3 - p1) * (p3 - p1)) / 2);
#pragma unroll
    for (int c = 0; c < 6; c++) v[14 + c] = eps[c];
    eig3(eps[0], eps[1], eps[2], eps[3], eps[4], eps[5], p1, p2, p3);
    v[20] = p1; v[21] = p2; v[22] = p3;
    v[23] = (2.0 / 3.0) * sqrt(((p1 - p2) * (p1 - p2) + (p2 - p3) * (p2 - p3) + (p3 - p1) * (p3 - p1)) / 2);
}

// cell data: [element][scalar][max, average, min]
// strain/stress/out are indexed from the first element of the slice, conn by global element
__global__ void __launch_bounds__(128)
k_cell_scalars(int64_t e_first, int64_t n_elem, const int32_t *__restrict__ conn_all, const int32_t *__restrict__ node_index,
               const double *__restrict__ ufull, const double *__restrict__ strain, const double *__restrict__ stress,
               float *__restrict__ out) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= n_elem) return;
    const int32_t *conn = conn_all + 8 * e_first;
    double mx[NS], mn[NS], sm[NS];
#pragma unroll
    for (int s = 0; s < NS; s++) { mx[s] = -DBL_MAX; mn[s] = DBL_MAX; sm[s] = 0.0; }
    for (int i = 0; i < 8; i++) {
        double v[NS];
        node_scalars(ufull + 3 * (int64_t)node_index[conn[8 * e + i]], stress + e * 48 + i * 6, strain + e * 48 + i * 6, v);
#pragma unroll
        for (int s = 0; s < NS; s++) { mx[s] = fmax(mx[s], v[s]); mn[s] = fmin(mn[s], v[s]); sm[s] += v[s]; }
    }
    float *o = out + e * (NS * 3);
#pragma unroll
    for (int s = 0; s < NS; s++) { o[3 * s] = (float)mx[s]; o[3 * s + 1] = (float)(sm[s] / 8); o[3 * s + 2] = (float)mn[s]; }
}

// point data: average over incident elements in ElemLib order.  One GPU: out[node (NodeLib order)][scalar].
// Partitioned: p counts this rank's rows from row0, strain/stress hold the touched elements compacted
// through `slot`, and out[p][scalar] stays in row order (the host maps rows to nodes with the DOF map).
__global__ void __launch_bounds__(128)
k_point_scalars(int64_t n_rows, int64_t row0, const int32_t *__restrict__ inc_ptr, const int32_t *__restrict__ inc,
                const int32_t *__restrict__ inv, const int32_t *__restrict__ slot, const double *__restrict__ ufull,
                const double *__restrict__ strain, const double *__restrict__ stress, float *__restrict__ out) {
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;   // local BFS row
    if (p >= n_rows) return;
    double sm[NS];
#pragma unroll
    for (int s = 0; s < NS; s++) sm[s] = 0.0;
    const int t0 = inc_ptr[p], t1 = inc_ptr[p + 1];
    int cnt = 0;
    int64_t last_e = -1;
    for (int t = t0; t < t1; t++) {
        const int64_t e = inc[t] >> 3;
        if (e == last_e) continue;                              // EList holds an element once; IndexOf = first position
        last_e = e;
        const int i = inc[t] & 7;
        const int64_t es = slot ? (int64_t)slot[e] : e;
        double v[NS];
        node_scalars(ufull + 3 * (row0 + p), stress + es * 48 + i * 6, strain + es * 48 + i * 6, v);
#pragma unroll
        for (int s = 0; s < NS; s++) sm[s] += v[s];
        cnt++;
    }
    float *o = out + (slot ? p : (int64_t)inv[p]) * NS;
#pragma unroll
    for (int s = 0; s < NS; s++) o[s] = (float)(sm[s] / cnt);
}

__global__ void k_touch(int64_t n_ent, const int32_t *__restrict__ inc, int32_t *__restrict__ touch) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t < n_ent) touch[inc[t] >> 3] = 1;
}
__global__ void k_touch_list(int64_t n_elem, const int32_t *__restrict__ touch, const int32_t *__restrict__ slot,
                             int32_t *__restrict__ list) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e < n_elem && touch[e]) list[slot[e]] = (int32_t)e;
}

}  // namespace

// One GPU: cell data for every element, point data for every node (NodeLib order).  Partitioned: cell
// data for the rank's element slice [elem0, elem1) and point data for its rows [row0, row1); the
// elements around those rows are recovered again locally (recovery is 0.7 ns per element) instead of
// shipping 768 B per element between ranks.
int run_postprocess(stan_handle *h, double *ms_out) {
    cudaStream_t s = h->stream;
    const bool multi = h->world > 1;
    const int64_t ne = h->elem1 - h->elem0, nloc = h->row1 - h->row0;
    STAN_TRY(h->d_cell.alloc((size_t)ne * NS * 3, s));
    STAN_TRY(h->d_point.alloc((size_t)nloc * NS, s));
    STAN_CUDA(cudaEventRecord(h->ev0, s));
    k_cell_scalars<<<div_up(ne, 128), 128, 0, s>>>(h->elem0, ne, h->d_conn.p, h->d_node_index.p, h->d_ufull.p,
                                                   h->d_strain.p, h->d_stress.p, h->d_cell.p);
    int64_t launches = 2;
    if (!multi) {
        k_point_scalars<<<div_up(nloc, 128), 128, 0, s>>>(nloc, 0, h->d_inc_ptr.p, h->d_inc.p, h->d_inv.p, nullptr,
                                                          h->d_ufull.p, h->d_strain.p, h->d_stress.p, h->d_point.p);
    } else {
        DevBuf<int32_t> touch, slot, list;
        DevBuf<double> t_strain, t_stress;
        STAN_TRY(touch.alloc(h->n_elem + 1, s)); STAN_TRY(slot.alloc(h->n_elem + 1, s));
        STAN_CUDA(cudaMemsetAsync(touch.p, 0, (h->n_elem + 1) * sizeof(int32_t), s));
        int32_t n_ent = 0, n_touch = 0;
        STAN_CUDA(cudaMemcpyAsync(&n_ent, h->d_inc_ptr.p + nloc, sizeof n_ent, cudaMemcpyDeviceToHost, s));
        STAN_CUDA(cudaStreamSynchronize(s));
        if (n_ent > 0) k_touch<<<div_up(n_ent, 256), 256, 0, s>>>(n_ent, h->d_inc.p, touch.p);
        STAN_TRY(device_exclusive_scan_i32(h, touch.p, slot.p, h->n_elem + 1, s));
        STAN_CUDA(cudaMemcpyAsync(&n_touch, slot.p + h->n_elem, sizeof n_touch, cudaMemcpyDeviceToHost, s));
        STAN_CUDA(cudaStreamSynchronize(s));
        STAN_TRY(list.alloc(n_touch, s));
        STAN_TRY(t_strain.alloc((size_t)48 * n_touch, s)); STAN_TRY(t_stress.alloc((size_t)48 * n_touch, s));
        k_touch_list<<<div_up(h->n_elem, 256), 256, 0, s>>>(h->n_elem, touch.p, slot.p, list.p);
        STAN_TRY(recover_elements(h, list.p, n_touch, t_strain.p, t_stress.p));
        k_point_scalars<<<div_up(nloc, 128), 128, 0, s>>>(nloc, h->row0, h->d_inc_ptr.p, h->d_inc.p, h->d_inv.p, slot.p,
                                                          h->d_ufull.p, t_strain.p, t_stress.p, h->d_point.p);
        touch.release(s); slot.release(s); list.release(s); t_strain.release(s); t_stress.release(s);
        launches += 4;
    }
    STAN_CUDA(cudaGetLastError());
    STAN_CUDA(cudaEventRecord(h->ev1, s));
    STAN_CUDA(cudaStreamSynchronize(s));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->ev0, h->ev1);
    if (ms_out) *ms_out = ms;
    h->launches += launches;
    h->postprocessed = true;
    return STAN_OK;
}

}  // namespace stan
