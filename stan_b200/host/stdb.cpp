// Protobuf wire codec for the STdb contracts (see stdb.hpp).  Hand-written: the image has no
// protoc/protobuf C++ runtime and the format is seven small messages.
#include "stdb.hpp"

#include <cstdio>
#include <cstring>

namespace stdb {

namespace {

enum { WT_VARINT = 0, WT_I64 = 1, WT_LEN = 2, WT_I32 = 5 };

struct Reader {
    const uint8_t *p, *end;
    bool ok = true;
    Reader(const uint8_t *b, const uint8_t *e) : p(b), end(e) {}
    bool more() const { return ok && p < end; }
    uint64_t varint() {
        uint64_t v = 0;
        for (int shift = 0; shift < 64; shift += 7) {
            if (p >= end) { ok = false; return 0; }
            uint8_t b = *p++;
            v |= (uint64_t)(b & 0x7f) << shift;
            if (!(b & 0x80)) return v;
        }
        ok = false;
        return 0;
    }
    double f64() {
        if (end - p < 8) { ok = false; return 0; }
        double d;
        memcpy(&d, p, 8);
        p += 8;
        return d;
    }
    Reader sub() {
        uint64_t n = varint();
        if (!ok || (uint64_t)(end - p) < n) { ok = false; return Reader(p, p); }
        Reader r(p, p + n);
        p += n;
        return r;
    }
    std::string str() {
        Reader r = sub();
        return std::string((const char *)r.p, (size_t)(r.end - r.p));
    }
    bool tag(uint32_t &field, int &wt) {
        uint64_t t = varint();
        field = (uint32_t)(t >> 3);
        wt = (int)(t & 7);
        return ok && field != 0;
    }
    void skip(int wt) {
        if (wt == WT_VARINT) varint();
        else if (wt == WT_I64) { if (end - p < 8) ok = false; else p += 8; }
        else if (wt == WT_LEN) sub();
        else if (wt == WT_I32) { if (end - p < 4) ok = false; else p += 4; }
        else ok = false;
    }
    // repeated scalar, packed (LEN) or unpacked
    void rep_i32(int wt, std::vector<int32_t> &out) {
        if (wt == WT_LEN) { Reader r = sub(); while (r.more()) out.push_back((int32_t)r.varint()); ok = ok && r.ok; }
        else if (wt == WT_VARINT) out.push_back((int32_t)varint());
        else ok = false;
    }
    void rep_f64(int wt, std::vector<double> &out) {
        if (wt == WT_LEN) { Reader r = sub(); while (r.more()) out.push_back(r.f64()); ok = ok && r.ok; }
        else if (wt == WT_I64) out.push_back(f64());
        else ok = false;
    }
};

struct Writer {
    std::string buf;
    void varint(uint64_t v) {
        while (v >= 0x80) { buf.push_back((char)(v | 0x80)); v >>= 7; }
        buf.push_back((char)v);
    }
    void tag(uint32_t field, int wt) { varint(((uint64_t)field << 3) | (uint64_t)wt); }
    void i32(uint32_t field, int32_t v, bool force = false) {
        if (v == 0 && !force) return;
        tag(field, WT_VARINT);
        varint((uint64_t)(int64_t)v);              // negative values are sign-extended to 64 bits
    }
    void f64(uint32_t field, double v, bool force = false) {
        if (v == 0.0 && !force) return;            // protobuf-net omits default values (also -0.0 != 0 bitwise is rare)
        tag(field, WT_I64);
        char b[8];
        memcpy(b, &v, 8);
        buf.append(b, 8);
    }
    void str(uint32_t field, const std::string &s) {
        if (s.empty()) return;
        tag(field, WT_LEN); varint(s.size()); buf += s;
    }
    void msg(uint32_t field, const std::string &payload) {
        tag(field, WT_LEN); varint(payload.size()); buf += payload;
    }
};

bool read_matrix(Reader r, MatrixST &m) {
    uint32_t f; int wt;
    while (r.more() && r.tag(f, wt)) {
        if (f == 1) r.rep_f64(wt, m.M);
        else if (f == 2 && wt == WT_VARINT) m.rows = (int32_t)r.varint();
        else if (f == 3 && wt == WT_VARINT) m.cols = (int32_t)r.varint();
        else r.skip(wt);
    }
    return r.ok;
}

std::string write_matrix(const MatrixST &m) {
    Writer w;
    for (double v : m.M) w.f64(1, v, true);
    w.i32(2, m.rows); w.i32(3, m.cols);
    return w.buf;
}

bool read_node(Reader r, Node &n) {
    uint32_t f; int wt;
    while (r.more() && r.tag(f, wt)) {
        switch (f) {
            case 1: n.id = (int32_t)r.varint(); break;
            case 2: n.x = r.f64(); break;
            case 3: n.y = r.f64(); break;
            case 4: n.z = r.f64(); break;
            case 5: r.rep_i32(wt, n.elist); break;
            case 6: r.rep_i32(wt, n.dof); break;
            case 7: r.rep_f64(wt, n.dispx); break;
            case 8: r.rep_f64(wt, n.dispy); break;
            case 9: r.rep_f64(wt, n.dispz); break;
            default: r.skip(wt);
        }
    }
    return r.ok;
}

std::string write_node(const Node &n) {
    Writer w;
    w.i32(1, n.id); w.f64(2, n.x); w.f64(3, n.y); w.f64(4, n.z);
    for (int32_t v : n.elist) w.i32(5, v, true);
    for (int32_t v : n.dof) w.i32(6, v, true);
    for (double v : n.dispx) w.f64(7, v, true);
    for (double v : n.dispy) w.f64(8, v, true);
    for (double v : n.dispz) w.f64(9, v, true);
    return w.buf;
}

bool read_element(Reader r, Element &e) {
    uint32_t f; int wt;
    while (r.more() && r.tag(f, wt)) {
        switch (f) {
            case 1: e.id = (int32_t)r.varint(); break;
            case 2: e.type = r.str(); break;
            case 3: e.pid = (int32_t)r.varint(); break;
            case 4: e.matid = (int32_t)r.varint(); break;
            case 5: r.rep_i32(wt, e.nlist); break;
            case 6: { MatrixST m; if (!read_matrix(r.sub(), m)) return false; e.strain.push_back(std::move(m)); break; }
            case 7: { MatrixST m; if (!read_matrix(r.sub(), m)) return false; e.stress.push_back(std::move(m)); break; }
            default: r.skip(wt);
        }
    }
    return r.ok;
}

std::string write_element(const Element &e) {
    Writer w;
    w.i32(1, e.id); w.str(2, e.type); w.i32(3, e.pid); w.i32(4, e.matid);
    for (int32_t v : e.nlist) w.i32(5, v, true);
    for (const MatrixST &m : e.strain) w.msg(6, write_matrix(m));
    for (const MatrixST &m : e.stress) w.msg(7, write_matrix(m));
    return w.buf;
}

bool read_material(Reader r, Material &m) {
    uint32_t f; int wt;
    while (r.more() && r.tag(f, wt)) {
        switch (f) {
            case 1: m.id = (int32_t)r.varint(); break;
            case 2: m.type = r.str(); break;
            case 3: m.name = r.str(); break;
            case 4: m.E = r.f64(); break;
            case 5: m.poisson = r.f64(); break;
            case 6: m.colorid = (int32_t)r.varint(); break;
            default: r.skip(wt);
        }
    }
    return r.ok;
}

std::string write_material(const Material &m) {
    Writer w;
    w.i32(1, m.id); w.str(2, m.type); w.str(3, m.name); w.f64(4, m.E); w.f64(5, m.poisson); w.i32(6, m.colorid);
    return w.buf;
}

bool read_bc(Reader r, BoundaryCondition &b) {
    uint32_t f; int wt;
    while (r.more() && r.tag(f, wt)) {
        switch (f) {
            case 1: b.type = r.str(); break;
            case 2: b.name = r.str(); break;
            case 3: b.id = (int32_t)r.varint(); break;
            case 4: {
                Reader e = r.sub();
                int32_t key = 0; MatrixST m;
                uint32_t ef; int ewt;
                while (e.more() && e.tag(ef, ewt)) {
                    if (ef == 1) key = (int32_t)e.varint();
                    else if (ef == 2) { if (!read_matrix(e.sub(), m)) return false; }
                    else e.skip(ewt);
                }
                if (!e.ok) return false;
                b.nodal.emplace_back(key, std::move(m));
                break;
            }
            case 5: b.colorid = (int32_t)r.varint(); break;
            default: r.skip(wt);
        }
    }
    return r.ok;
}

std::string write_bc(const BoundaryCondition &b) {
    Writer w;
    w.str(1, b.type); w.str(2, b.name); w.i32(3, b.id);
    for (const auto &kv : b.nodal) {
        Writer e;
        e.i32(1, kv.first);
        e.msg(2, write_matrix(kv.second));
        w.msg(4, e.buf);
    }
    w.i32(5, b.colorid);
    return w.buf;
}

bool read_analysis(Reader r, Analysis &a) {
    uint32_t f; int wt;
    a.present = true;
    while (r.more() && r.tag(f, wt)) {
        switch (f) {
            case 1: a.type = r.str(); break;
            case 2: a.linsolver = r.str(); break;
            case 3: a.tolerance = r.f64(); break;
            case 4: a.itermax = (int32_t)r.varint(); break;
            case 5: a.incnumb = (int32_t)r.varint(); break;
            case 6: a.result_stepno = (int32_t)r.varint(); break;
            default: r.skip(wt);
        }
    }
    return r.ok;
}

std::string write_analysis(const Analysis &a) {
    Writer w;
    w.str(1, a.type); w.str(2, a.linsolver); w.f64(3, a.tolerance); w.i32(4, a.itermax); w.i32(5, a.incnumb);
    w.i32(6, a.result_stepno);
    return w.buf;
}

// one Dictionary<int,T> entry: {1: key, 2: value}
template <typename T, typename F>
bool read_entry(Reader e, int32_t &key, T &value, F read_value) {
    uint32_t f; int wt;
    while (e.more() && e.tag(f, wt)) {
        if (f == 1 && wt == WT_VARINT) key = (int32_t)e.varint();
        else if (f == 2 && wt == WT_LEN) { if (!read_value(e.sub(), value)) return false; }
        else e.skip(wt);
    }
    return e.ok;
}

std::string entry(int32_t key, const std::string &payload) {
    Writer e;
    e.i32(1, key);
    e.msg(2, payload);
    return e.buf;
}

}  // namespace

bool decode(const std::string &bytes, Database &db, std::string &err) {
    db = Database();
    Reader r((const uint8_t *)bytes.data(), (const uint8_t *)bytes.data() + bytes.size());
    uint32_t f; int wt;
    while (r.more() && r.tag(f, wt)) {
        int32_t key = 0;
        if (f == 1 && wt == WT_LEN) {
            Node n;
            if (!read_entry(r.sub(), key, n, read_node)) { err = "malformed NodeLib entry"; return false; }
            if (n.id == 0) n.id = key;
            db.nodes.push_back(std::move(n));
        } else if (f == 2 && wt == WT_LEN) {
            Element e;
            if (!read_entry(r.sub(), key, e, read_element)) { err = "malformed ElemLib entry"; return false; }
            if (e.id == 0) e.id = key;
            db.elems.push_back(std::move(e));
        } else if (f == 3 && wt == WT_LEN) {
            Material m;
            if (!read_entry(r.sub(), key, m, read_material)) { err = "malformed MatLib entry"; return false; }
            if (m.id == 0) m.id = key;
            db.mats.push_back(std::move(m));
        } else if (f == 4 && wt == WT_LEN) {
            BoundaryCondition b;
            if (!read_entry(r.sub(), key, b, read_bc)) { err = "malformed BCLib entry"; return false; }
            db.bcs.push_back(std::move(b));
            db.bc_keys.push_back(key);
        } else if (f == 5 && wt == WT_VARINT) {
            db.ndof = (int32_t)r.varint();
        } else if (f == 6 && wt == WT_LEN) {
            if (!read_analysis(r.sub(), db.analysis)) { err = "malformed Analysis"; return false; }
        } else if (f == 7 && wt == WT_LEN) {
            db.info_raw = r.str();
            db.has_info = true;
        } else {
            r.skip(wt);
        }
    }
    if (!r.ok) { err = "truncated or malformed STdb stream"; return false; }
    return true;
}

std::string encode(const Database &db) {
    Writer w;
    for (const Node &n : db.nodes) w.msg(1, entry(n.id, write_node(n)));
    for (const Element &e : db.elems) w.msg(2, entry(e.id, write_element(e)));
    for (const Material &m : db.mats) w.msg(3, entry(m.id, write_material(m)));
    for (size_t i = 0; i < db.bcs.size(); i++)
        w.msg(4, entry(i < db.bc_keys.size() ? db.bc_keys[i] : db.bcs[i].id, write_bc(db.bcs[i])));
    w.i32(5, db.ndof);
    if (db.analysis.present) w.msg(6, write_analysis(db.analysis));
    if (db.has_info) w.msg(7, db.info_raw);
    return w.buf;
}

bool read_file(const std::string &path, std::string &bytes, std::string &err) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) { err = "cannot open " + path; return false; }
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    bytes.resize((size_t)n);
    size_t got = n ? fread(&bytes[0], 1, (size_t)n, f) : 0;
    fclose(f);
    if (got != (size_t)n) { err = "short read on " + path; return false; }
    return true;
}

bool write_file(const std::string &path, const std::string &bytes, std::string &err) {
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) { err = "cannot write " + path; return false; }
    size_t put = bytes.empty() ? 0 : fwrite(bytes.data(), 1, bytes.size(), f);
    fclose(f);
    if (put != bytes.size()) { err = "short write on " + path; return false; }
    return true;
}

}  // namespace stdb
