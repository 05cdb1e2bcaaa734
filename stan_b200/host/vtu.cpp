// See vtu.hpp.  Layout of the file follows what the reference assembles before handing it to VTK:
// every part contributes its own point list — the distinct node IDs of its elements in ascending
// order (Part.DetectPartNodes, Part.cs:721-747), moved by the displacement of the increment
// (Part.UpdateNode, :581-594) — and its hexahedra (VTK cell type 12, :869-886); parts are appended one
// after the other in ascending PID order (PartLib) without merging points (vtkAppendFilter), so interface nodes appear once per part.
// Point data = the selected nodal-averaged arrays, named without the " INC n" suffix (:920-934), in
// the order Displacement, Strain, Stress (ExportWindow.xaml.cs:66-68).
#include "vtu.hpp"

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <type_traits>
#include <unordered_map>

namespace vtu {

namespace {

const char *NAMES[24] = {"Displacement X", "Displacement Y", "Displacement Z", "Total Displacement",
                         "Stress XX", "Stress YY", "Stress ZZ", "Stress XY", "Stress YZ", "Stress XZ",
                         "Stress P1", "Stress P2", "Stress P3", "von Mises Stress",
                         "Strain XX", "Strain YY", "Strain ZZ", "Strain XY", "Strain YZ", "Strain XZ",
                         "Strain P1", "Strain P2", "Strain P3", "Effective Strain"};

std::string base64(const unsigned char *p, size_t n) {
    static const char T[] = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
    std::string o;
    o.reserve((n + 2) / 3 * 4);
    for (size_t i = 0; i < n; i += 3) {
        const unsigned a = p[i], b = i + 1 < n ? p[i + 1] : 0, c = i + 2 < n ? p[i + 2] : 0;
        o += T[a >> 2];
        o += T[((a & 3) << 4) | (b >> 4)];
        o += i + 1 < n ? T[((b & 15) << 2) | (c >> 6)] : '=';
        o += i + 2 < n ? T[c & 63] : '=';
    }
    return o;
}

template <typename T>
void data_array(FILE *f, const char *type, const char *name, int ncomp, const std::vector<T> &v, bool ascii) {
    fprintf(f, "        <DataArray type=\"%s\"", type);
    if (name) fprintf(f, " Name=\"%s\"", name);
    if (ncomp > 1) fprintf(f, " NumberOfComponents=\"%d\"", ncomp);
    fprintf(f, " format=\"%s\">\n", ascii ? "ascii" : "binary");
    if (ascii) {
        for (size_t i = 0; i < v.size(); i++) {
            if (sizeof(T) == 1) fprintf(f, "%d", (int)v[i]);
            else if (std::is_floating_point<T>::value) fprintf(f, "%.9g", (double)v[i]);
            else fprintf(f, "%lld", (long long)v[i]);
            fputc((i + 1) % 12 == 0 || i + 1 == v.size() ? '\n' : ' ', f);
        }
    } else {                                              // VTK inline binary: base64(UInt32 byte count) + base64(data)
        const uint32_t nbytes = (uint32_t)(v.size() * sizeof(T));
        fputs(base64(reinterpret_cast<const unsigned char *>(&nbytes), 4).c_str(), f);
        fputs(base64(reinterpret_cast<const unsigned char *>(v.data()), nbytes).c_str(), f);
        fputc('\n', f);
    }
    fprintf(f, "        </DataArray>\n");
}

}  // namespace

bool write_increment(const stdb::Database &db, const std::vector<double> &disp, const std::vector<float> &point,
                     const std::string &prefix, bool ascii, std::string &path_out, std::string &err) {
    const size_t nn = db.nodes.size();
    if (disp.size() != 3 * nn || point.size() != 24 * nn) { err = "vtu: result arrays do not match the node count"; return false; }
    std::unordered_map<int32_t, size_t> pos;
    pos.reserve(nn * 2);
    for (size_t i = 0; i < nn; i++) pos[db.nodes[i].id] = i;

    // PartLib = distinct PIDs of ElemLib, sorted (Database.cs:98-108): std::map iterates in that order
    std::map<int32_t, std::vector<int32_t>> part_nodes;
    for (const auto &e : db.elems) {
        if (e.type.find("HEX") == std::string::npos) continue;
        auto &ids = part_nodes[e.pid];
        ids.insert(ids.end(), e.nlist.begin(), e.nlist.end());
    }
    std::vector<float> xyz;
    std::vector<size_t> src;                              // node position behind every output point
    std::vector<int64_t> conn, offsets;
    std::vector<uint8_t> types;
    for (auto &part : part_nodes) {
        const int32_t pid = part.first;
        auto &ids = part.second;
        std::sort(ids.begin(), ids.end());
        ids.erase(std::unique(ids.begin(), ids.end()), ids.end());
        const int64_t base = (int64_t)src.size();
        std::unordered_map<int32_t, int64_t> local;
        local.reserve(ids.size() * 2);
        for (size_t k = 0; k < ids.size(); k++) {
            auto p = pos.find(ids[k]);
            if (p == pos.end()) { err = "vtu: element references a missing node"; return false; }
            local[ids[k]] = base + (int64_t)k;
            src.push_back(p->second);
            const stdb::Node &n = db.nodes[p->second];
            xyz.push_back((float)(disp[3 * p->second] + n.x));
            xyz.push_back((float)(disp[3 * p->second + 1] + n.y));
            xyz.push_back((float)(disp[3 * p->second + 2] + n.z));
        }
        for (const auto &e : db.elems) {
            if (e.pid != pid || e.type.find("HEX") == std::string::npos || e.nlist.size() != 8) continue;
            for (int k = 0; k < 8; k++) conn.push_back(local[e.nlist[k]]);
            offsets.push_back((int64_t)conn.size());
            types.push_back(12);                          // VTK_HEXAHEDRON
        }
    }

    path_out = prefix + "_001.vtu";                       // Prefix + "_" + inc.ToString("000") + ".vtu"
    FILE *f = fopen(path_out.c_str(), "wb");
    if (!f) { err = "vtu: cannot open " + path_out; return false; }
    fprintf(f, "<?xml version=\"1.0\"?>\n<VTKFile type=\"UnstructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\">\n");
    fprintf(f, "  <UnstructuredGrid>\n    <Piece NumberOfPoints=\"%zu\" NumberOfCells=\"%zu\">\n", src.size(), types.size());
    fprintf(f, "      <PointData>\n");
    static const int ORDER[24] = {0, 1, 2, 3, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13};
    std::vector<float> col(src.size());
    for (int a = 0; a < 24; a++) {
        const int sidx = ORDER[a];
        for (size_t i = 0; i < src.size(); i++) col[i] = point[24 * src[i] + sidx];
        data_array(f, "Float32", NAMES[sidx], 1, col, ascii);
    }
    fprintf(f, "      </PointData>\n      <CellData>\n      </CellData>\n      <Points>\n");
    data_array(f, "Float32", "Points", 3, xyz, ascii);
    fprintf(f, "      </Points>\n      <Cells>\n");
    data_array(f, "Int64", "connectivity", 1, conn, ascii);
    data_array(f, "Int64", "offsets", 1, offsets, ascii);
    data_array(f, "UInt8", "types", 1, types, ascii);
    fprintf(f, "      </Cells>\n    </Piece>\n  </UnstructuredGrid>\n</VTKFile>\n");
    const bool ok = fclose(f) == 0;
    if (!ok) err = "vtu: write failed for " + path_out;
    return ok;
}

}  // namespace vtu
