#include "bdf.hpp"

#include <cstdlib>
#include <fstream>
#include <iterator>
#include <unordered_set>

namespace bdf {

namespace {

bool starts_with(const std::string &s, const char *p) { return s.rfind(p, 0) == 0; }

// int.Parse / int.TryParse: optional surrounding blanks, optional sign, digits only
bool parse_int(const std::string &s, int32_t &out) {
    size_t a = 0, b = s.size();
    auto ws = [](char c) { return c == ' ' || (c >= '\t' && c <= '\r'); };   // NumberStyles.AllowLeading/TrailingWhite
    while (a < b && ws(s[a])) a++;
    while (b > a && ws(s[b - 1])) b--;
    if (a == b) return false;
    size_t i = a;
    if (s[i] == '+' || s[i] == '-') i++;
    if (i == b) return false;
    long long v = 0;
    for (size_t k = i; k < b; k++) {
        if (s[k] < '0' || s[k] > '9') return false;
        v = v * 10 + (s[k] - '0');
        if (v > 2147483648LL) return false;
    }
    if (s[a] == '-') v = -v;
    if (v > 2147483647LL) return false;
    out = (int32_t)v;
    return true;
}

// double.Parse(s, InvariantCulture) for what a deck can contain: [ws][sign]digits[.digits][e[sign]digits][ws]
// (strtod alone would also take hex floats, "inf" and "nan", which .NET rejects in this spelling)
bool parse_double(const std::string &s, double &out) {
    size_t a = 0, b = s.size();
    auto ws = [](char c) { return c == ' ' || (c >= '\t' && c <= '\r'); };
    while (a < b && ws(s[a])) a++;
    while (b > a && ws(s[b - 1])) b--;
    size_t i = a;
    if (i < b && (s[i] == '+' || s[i] == '-')) i++;
    size_t nd = 0;
    while (i < b && s[i] >= '0' && s[i] <= '9') { i++; nd++; }
    if (i < b && s[i] == '.') { i++; while (i < b && s[i] >= '0' && s[i] <= '9') { i++; nd++; } }
    if (nd == 0) return false;
    if (i < b && (s[i] == 'e' || s[i] == 'E')) {
        i++;
        if (i < b && (s[i] == '+' || s[i] == '-')) i++;
        size_t ne = 0;
        while (i < b && s[i] >= '0' && s[i] <= '9') { i++; ne++; }
        if (ne == 0) return false;
    }
    if (i != b) return false;
    out = strtod(s.substr(a, b - a).c_str(), nullptr);
    return true;
}

void replace_all(std::string &s, const std::string &from, const std::string &to) {
    size_t pos = 0;
    while ((pos = s.find(from, pos)) != std::string::npos) { s.replace(pos, from.size(), to); pos += to.size(); }
}

// Node(string input), Node.cs:25-80
bool parse_grid(const std::string &input, stdb::Node &n) {
    std::vector<std::string> data;
    for (size_t i = 0; i < input.size() / 8; i++) {
        std::string text = input.substr(i * 8, 8);
        replace_all(text, " ", "");
        if (text.empty()) continue;
        bool blank = true;
        for (char c : text) if (c != '\t' && c != '\r') blank = false;
        if (blank) continue;                                     // IsNullOrWhiteSpace
        if (text.find('e') == std::string::npos && text.find('E') == std::string::npos) {
            if (text.substr(1).find('-') != std::string::npos) {
                if (text[0] == '-') { std::string t = text.substr(1); replace_all(t, "-", "e-"); text = "-" + t; }
                else replace_all(text, "-", "e-");
            }
            // text.Replace("+", "e+") is computed and thrown away in the reference (Node.cs:55)
        }
        if (text[0] == '.') text = "0" + text;
        data.push_back(text);
    }
    if (data.size() < 5) return false;
    if (!parse_int(data[1], n.id)) return false;
    if (!parse_double(data[2], n.x) || !parse_double(data[3], n.y) || !parse_double(data[4], n.z)) return false;
    n.dof.assign(3, 0);                                          // DOF = new int[3]
    n.dispx = {0.0}; n.dispy = {0.0}; n.dispz = {0.0};          // Disp at time 0
    return true;
}

// Element(string input), Element.cs:35-73
bool parse_element(const std::string &input, stdb::Element &e) {
    std::vector<std::string> data;                               // Regex.Split(input, @"\s+")
    std::string cur;
    bool in_ws = false;
    for (char c : input) {
        bool ws = (c == ' ' || c == '\t' || c == '\r' || c == '\n' || c == '\f' || c == '\v');
        if (ws) { if (!in_ws) { data.push_back(cur); cur.clear(); in_ws = true; } }
        else { cur.push_back(c); in_ws = false; }
    }
    data.push_back(cur);
    if (data.size() < 3) return false;
    if (!parse_int(data[1], e.id) || !parse_int(data[2], e.pid)) return false;
    for (size_t i = 3; i < data.size(); i++) {
        std::string t = data[i];
        replace_all(t, "+", "");
        int32_t v;
        if (parse_int(t, v)) e.nlist.push_back(v);
    }
    if (data[0] == "CHEXA") e.type = "HEX8_G2";
    if (data[0] == "CPENTA") e.type = "PENTA6_G2";
    if (data[0] == "CTETRA") e.type = "TET4_G2";
    e.matid = 0;
    return true;
}

}  // namespace

bool read_nastran_mesh(const std::string &path, stdb::Database &db, ImportReport &rep, std::string &err) {
    std::ifstream in(path, std::ios::binary);
    if (!in) { err = "cannot open " + path; return false; }
    const std::string all((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    std::vector<std::string> data;                               // File.ReadAllLines: \r\n, \n and lone \r end a line
    for (size_t a = 0; a < all.size();) {
        size_t b = a;
        while (b < all.size() && all[b] != '\n' && all[b] != '\r') b++;
        data.push_back(all.substr(a, b - a));
        a = b + ((b + 1 < all.size() && all[b] == '\r' && all[b + 1] == '\n') ? 2 : 1);
    }
    std::unordered_set<int32_t> node_ids, elem_ids;
    for (size_t i = 0; i < data.size(); i++) {
        if (starts_with(data[i], "$")) continue;
        if (data[i].find("CHEXA") != std::string::npos) {        // Elem_types_allowed = { "CHEXA" }
            std::string temp = data[i];
            for (size_t j = i + 1; j < data.size(); j++) {
                if (starts_with(data[j], "+") || starts_with(data[j], " ")) { temp += data[j]; i = j; }
                else break;
            }
            stdb::Element e;
            if (parse_element(temp, e) && elem_ids.insert(e.id).second) db.elems.push_back(std::move(e));
            else rep.errors.push_back(temp);
        }
        if (starts_with(data[i], "GRID")) {
            stdb::Node n;
            if (parse_grid(data[i], n) && node_ids.insert(n.id).second) db.nodes.push_back(std::move(n));
            else rep.errors.push_back(data[i]);
        }
    }
    return true;
}

}  // namespace bdf
