// See model_build.hpp.
#include "model_build.hpp"

#include <cerrno>
#include <cstdlib>
#include <cstring>
#include <unordered_set>

namespace model_build {

namespace {

std::vector<std::string> split_keep_empty(const std::string &s, char sep) {       // string.Split(char)
    std::vector<std::string> out;
    size_t a = 0;
    while (true) {
        const size_t b = s.find(sep, a);
        out.push_back(s.substr(a, b == std::string::npos ? std::string::npos : b - a));
        if (b == std::string::npos) break;
        a = b + 1;
    }
    return out;
}

std::string trim(const std::string &s) {                                           // .NET parsers allow outer white space
    const char *ws = " \t\r\n\v\f";
    const size_t a = s.find_first_not_of(ws);
    if (a == std::string::npos) return "";
    return s.substr(a, s.find_last_not_of(ws) - a + 1);
}

bool parse_int(const std::string &raw, int32_t &v) {                               // int.Parse
    const std::string s = trim(raw);
    if (s.empty()) return false;
    size_t i = (s[0] == '+' || s[0] == '-') ? 1 : 0;
    if (i == s.size()) return false;
    for (size_t k = i; k < s.size(); k++)
        if (s[k] < '0' || s[k] > '9') return false;
    errno = 0;
    const long long x = strtoll(s.c_str(), nullptr, 10);
    if (errno || x < INT32_MIN || x > INT32_MAX) return false;
    v = (int32_t)x;
    return true;
}

bool parse_double(const std::string &raw, double &v) {                             // double.Parse(s, InvariantCulture)
    const std::string s = trim(raw);
    if (s.empty()) return false;
    for (char c : s)
        if (!(c >= '0' && c <= '9') && c != '+' && c != '-' && c != '.' && c != 'e' && c != 'E') return false;
    char *end = nullptr;
    v = strtod(s.c_str(), &end);
    return end && *end == '\0';
}

}  // namespace

std::vector<BcRow> parse_bc_text(const std::string &text) {
    std::vector<BcRow> rows;
    const std::vector<std::string> lines = split_keep_empty(text, '\n');
    if (lines.size() <= 1) return rows;                                            // `if (text.Length > 1)`
    for (const std::string &s : lines) {
        std::vector<std::string> f = split_keep_empty(s, ',');
        if (f.size() != 4) {
            f = split_keep_empty(s, ' ');
            if (f.size() != 4) f = split_keep_empty(s, '\t');
        }
        if (f.size() != 4) continue;
        BcRow r;
        if (parse_int(f[0], r.nid) && parse_double(f[1], r.v[0]) && parse_double(f[2], r.v[1]) && parse_double(f[3], r.v[2]))
            rows.push_back(r);                                                     // `catch { }` otherwise
    }
    return rows;
}

void add_material(stdb::Database &db, double E, double poisson) {
    stdb::Material m;
    m.id = db.mats.empty() ? 1 : db.mats.back().id + 1;                            // MainWindow.xaml.cs:393-401
    m.type = "Elastic";
    m.name = "New Material";
    m.colorid = m.id % 9;
    m.E = E;
    m.poisson = poisson;
    db.mats.push_back(m);
}

void set_part_material(stdb::Database &db, int32_t pid, int32_t matid) {
    for (auto &e : db.elems)
        if (pid < 0 || e.pid == pid) e.matid = matid;
}

void set_hex_type(stdb::Database &db, int32_t pid, const std::string &type) {
    for (auto &e : db.elems)
        if ((pid < 0 || e.pid == pid) && e.type.find("HEX") != std::string::npos) e.type = type;
}

bool add_bc(stdb::Database &db, const std::string &type, const std::string &name, const std::vector<BcRow> &rows,
            std::string &err) {
    std::unordered_set<int32_t> known, seen;
    known.reserve(db.nodes.size() * 2);
    for (const auto &n : db.nodes) known.insert(n.id);
    stdb::BoundaryCondition bc;
    bc.type = type;
    bc.name = name;
    bc.id = db.bcs.empty() ? 1 : db.bcs.back().id + 1;                             // MainWindow.xaml.cs:420-429
    bc.colorid = bc.id % 9;
    for (const BcRow &r : rows) {
        if (!known.count(r.nid)) continue;                                         // BoundaryCondition.cs:89
        if (!seen.insert(r.nid).second) {
            err = "node " + std::to_string(r.nid) + " listed twice in boundary condition '" + name +
                  "' (Dictionary.Add throws in the reference)";
            return false;
        }
        stdb::MatrixST m;
        m.rows = 3; m.cols = 1;
        m.M.assign(r.v, r.v + 3);
        bc.nodal.emplace_back(r.nid, m);
    }
    db.bcs.push_back(bc);
    db.bc_keys.push_back(bc.id);
    return true;
}

void set_analysis(stdb::Database &db, const std::string &linsolver, double tolerance, int32_t itermax) {
    db.analysis.present = true;
    db.analysis.type = "Linear_Statics";
    db.analysis.linsolver = linsolver;
    db.analysis.tolerance = tolerance;
    db.analysis.itermax = itermax;
    db.analysis.incnumb = 0;
    db.analysis.result_stepno = 0;
}

}  // namespace model_build
