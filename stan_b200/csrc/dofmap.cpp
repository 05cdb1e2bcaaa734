// Native Database.AssignDOF (/root/reference/src/STAN_Database/Database.cs:140-234).
//
// The reference keeps a FIFO that tolerates duplicates and numbers a node when it is first
// popped.  Numbering on first pop of a FIFO equals numbering on first push, so this is a plain
// discovery-order BFS whose neighbour order is "incident elements in ElemLib order x NList
// order" — O(nodes + 8*elements) time, no per-node neighbour lists are materialised.
#include <cstdint>
#include <cstdlib>
#include <vector>

#include "../../include/stan_b200.h"

namespace stan {

void set_error(const char *fmt, ...);

int assign_dof_host(int64_t n_nodes, int64_t n_elem, const int32_t *conn, int32_t *node_index) {
    // node -> incident elements, element order, duplicates of one element collapsed
    // (AddElem2Nodes + RemoveElemDuplicates, Database.cs:143-158)
    std::vector<int64_t> ptr((size_t)n_nodes + 1, 0);
    std::vector<int32_t> last((size_t)n_nodes, -1);
    for (int64_t e = 0; e < n_elem; e++)
        for (int k = 0; k < 8; k++) {
            int32_t n = conn[8 * e + k];
            if (n < 0 || n >= n_nodes) {
                set_error("element %lld references node %d outside [0,%lld)", (long long)e, n, (long long)n_nodes);
                return STAN_E_ARG;
            }
            if (last[n] != (int32_t)e) { last[n] = (int32_t)e; ptr[n + 1]++; }
        }
    for (int64_t i = 0; i < n_nodes; i++) ptr[i + 1] += ptr[i];
    std::vector<int32_t> idx((size_t)ptr[n_nodes]);
    {
        std::vector<int64_t> fill(ptr.begin(), ptr.end() - 1);
        std::fill(last.begin(), last.end(), -1);
        for (int64_t e = 0; e < n_elem; e++)
            for (int k = 0; k < 8; k++) {
                int32_t n = conn[8 * e + k];
                if (last[n] != (int32_t)e) { last[n] = (int32_t)e; idx[fill[n]++] = (int32_t)e; }
            }
    }
    // first node (NodeLib order) with exactly 1, else 2, ... 6 incident elements (Database.cs:178-196)
    int64_t first = -1;
    for (int want = 1; want < 7 && first < 0; want++)
        for (int64_t n = 0; n < n_nodes; n++)
            if (ptr[n + 1] - ptr[n] == want) { first = n; break; }
    if (first < 0) {
        set_error("AssignDOF: no node with 1..6 incident elements (reference would look up node 0 and throw)");
        return STAN_E_DOFMAP;
    }
    for (int64_t i = 0; i < n_nodes; i++) node_index[i] = -1;
    std::vector<int32_t> queue((size_t)n_nodes);
    int64_t head = 0, tail = 0;
    int32_t next = 0;
    node_index[first] = next++;
    queue[tail++] = (int32_t)first;
    while (head < tail) {
        int32_t v = queue[head++];
        for (int64_t t = ptr[v]; t < ptr[v + 1]; t++) {
            const int32_t *nl = conn + 8 * (int64_t)idx[t];
            for (int k = 0; k < 8; k++) {
                int32_t w = nl[k];
                if (node_index[w] < 0) { node_index[w] = next++; queue[tail++] = w; }
            }
        }
    }
    if (next != n_nodes) {
        set_error("AssignDOF: mesh is disconnected (%d of %lld nodes reached; the reference runs off its queue)",
                  next, (long long)n_nodes);
        return STAN_E_DOFMAP;
    }
    return STAN_OK;
}

}  // namespace stan
