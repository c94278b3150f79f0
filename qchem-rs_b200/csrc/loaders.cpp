// loaders.cpp -- basis-set and molecule loaders behind the C ABI: the native stand-in for the two
// `molint` loader calls the reference makes before an SCF run,
//     BasisSet::load(path)                 qchem-cli/src/main.rs:76, :120
//     MolecularSystem::load(path, &basis)  qchem-cli/src/main.rs:77, :121
// (the crate itself is an absent path dependency, Cargo.toml:12).  File formats, SURVEY.md 2 rows 8-9:
//   basis    : MolSSI-BSE "complete" JSON -- elements[Z].electron_shells[*] with string-encoded `exponents` and
//              `coefficients[k]` (one row per entry of `angular_momentum`; fused SP shells carry [0, 1]),
//              `function_type` in gto / gto_cartesian / gto_spherical;
//   molecule : JSON array of {"element": "<Z>", "position": [x, y, z]}, positions in bohr (rhf.rs:116-117).
// Conventions (identical to qchem-rs_b200/basis.py, which the tests compare against): SP shells split into s then p,
// zero contraction coefficients dropped, shells atom-major in file order, coefs[k] = c_k * N(a_k; l,0,0).
// Pure host code: no CUDA call in this file.
#include "../../include/qcfock.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

namespace {

// ---- a small JSON reader (objects, arrays, strings, numbers, literals) ------------------------------------
struct JValue {
    enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
    double num = 0;
    bool b = false;
    std::string str;
    std::vector<JValue> arr;
    std::vector<std::pair<std::string, JValue>> obj;
    const JValue* get(const std::string& k) const {
        for (auto& kv : obj) if (kv.first == k) return &kv.second;
        return nullptr;
    }
};

struct JParser {
    const std::string& s;
    size_t i = 0;
    std::string err;
    explicit JParser(const std::string& t) : s(t) {}
    void ws() { while (i < s.size() && (s[i] == ' ' || s[i] == '\n' || s[i] == '\t' || s[i] == '\r')) ++i; }
    bool fail(const std::string& m) { if (err.empty()) err = m + " at byte " + std::to_string(i); return false; }
    bool parse(JValue& v, int depth = 0) {
        if (depth > 64) return fail("nesting too deep");
        ws();
        if (i >= s.size()) return fail("unexpected end");
        const char c = s[i];
        if (c == '{') {
            v.kind = JValue::Obj; ++i; ws();
            if (i < s.size() && s[i] == '}') { ++i; return true; }
            while (true) {
                ws();
                JValue k;
                if (i >= s.size() || s[i] != '"' || !string(k.str)) return fail("object key expected");
                ws();
                if (i >= s.size() || s[i] != ':') return fail("':' expected");
                ++i;
                JValue val;
                if (!parse(val, depth + 1)) return false;
                v.obj.emplace_back(std::move(k.str), std::move(val));
                ws();
                if (i < s.size() && s[i] == ',') { ++i; continue; }
                if (i < s.size() && s[i] == '}') { ++i; return true; }
                return fail("',' or '}' expected");
            }
        }
        if (c == '[') {
            v.kind = JValue::Arr; ++i; ws();
            if (i < s.size() && s[i] == ']') { ++i; return true; }
            while (true) {
                JValue e;
                if (!parse(e, depth + 1)) return false;
                v.arr.push_back(std::move(e));
                ws();
                if (i < s.size() && s[i] == ',') { ++i; continue; }
                if (i < s.size() && s[i] == ']') { ++i; return true; }
                return fail("',' or ']' expected");
            }
        }
        if (c == '"') { v.kind = JValue::Str; return string(v.str); }
        if (!s.compare(i, 4, "true")) { v.kind = JValue::Bool; v.b = true; i += 4; return true; }
        if (!s.compare(i, 5, "false")) { v.kind = JValue::Bool; v.b = false; i += 5; return true; }
        if (!s.compare(i, 4, "null")) { v.kind = JValue::Null; i += 4; return true; }
        char* end = nullptr;
        v.num = std::strtod(s.c_str() + i, &end);
        if (end == s.c_str() + i) return fail("value expected");
        v.kind = JValue::Num;
        i = end - s.c_str();
        return true;
    }
    bool string(std::string& out) {
        ++i;   // opening quote
        while (i < s.size() && s[i] != '"') {
            if (s[i] == '\\') {
                if (++i >= s.size()) return fail("bad escape");
                switch (s[i]) {
                    case 'n': out += '\n'; break; case 't': out += '\t'; break; case 'r': out += '\r'; break;
                    case 'b': out += '\b'; break; case 'f': out += '\f'; break;
                    case 'u': if (i + 4 >= s.size()) return fail("bad \\u escape"); out += '?'; i += 4; break;
                    default: out += s[i];
                }
                ++i;
            } else out += s[i++];
        }
        if (i >= s.size()) return fail("unterminated string");
        ++i;
        return true;
    }
};

bool read_file(const char* path, std::string& out, std::string& err) {
    std::ifstream f(path, std::ios::binary);
    if (!f) { err = std::string("cannot open ") + path; return false; }
    std::stringstream ss;
    ss << f.rdbuf();
    out = ss.str();
    return true;
}

// numbers appear both as JSON numbers and as strings ("0.1543289673E+00") in BSE files
bool as_number(const JValue& v, double& x) {
    if (v.kind == JValue::Num) { x = v.num; return true; }
    if (v.kind == JValue::Str) {
        char* end = nullptr;
        x = std::strtod(v.str.c_str(), &end);
        return end != v.str.c_str();
    }
    return false;
}

double dfact(int n) { double r = 1; while (n > 1) { r *= n; n -= 2; } return r; }
// normalisation of x^l exp(-a r^2)
double prim_norm(double a, int l) {
    return std::pow(2.0 * a / M_PI, 0.75) * std::pow(4.0 * a, l / 2.0) / std::sqrt(dfact(2 * l - 1));
}

struct ShellDef { int l; std::vector<double> exps, coefs; std::string ftype; };

}  // namespace

struct qcf_system {
    std::string err;
    std::vector<int> Z, shell_atom, shell_l, shell_nprim, shell_prim_off;
    std::vector<double> xyz, exps, coefs;
    qcf_basis basis{};
    int n_electrons = 0, n_basis = 0;
};

extern "C" {

int qcf_system_load(const char* basis_json, const char* molecule_json, qcf_system** out) {
    if (!basis_json || !molecule_json || !out) return QCF_ERR_ARG;
    qcf_system* sys = new qcf_system();
    *out = sys;
    auto fail = [&](const std::string& m) { sys->err = m; return QCF_ERR_ARG; };
    std::string btxt, mtxt;
    if (!read_file(basis_json, btxt, sys->err) || !read_file(molecule_json, mtxt, sys->err)) return QCF_ERR_ARG;
    JValue bdoc, mdoc;
    { JParser p(btxt); if (!p.parse(bdoc)) return fail(std::string(basis_json) + ": " + p.err); }
    { JParser p(mtxt); if (!p.parse(mdoc)) return fail(std::string(molecule_json) + ": " + p.err); }
    const JValue* elements = bdoc.get("elements");
    if (bdoc.kind != JValue::Obj || !elements || elements->kind != JValue::Obj) return fail("basis file has no \"elements\" object");
    // BasisSet::load
    std::map<int, std::vector<ShellDef>> table;
    for (auto& kv : elements->obj) {
        const int z = std::atoi(kv.first.c_str());
        std::vector<ShellDef> shells;
        const JValue* es = kv.second.get("electron_shells");
        if (es && es->kind == JValue::Arr)
            for (const JValue& sh : es->arr) {
                const JValue* am = sh.get("angular_momentum");
                const JValue* ex = sh.get("exponents");
                const JValue* co = sh.get("coefficients");
                if (!am || !ex || !co || am->kind != JValue::Arr || ex->kind != JValue::Arr || co->kind != JValue::Arr)
                    return fail("malformed electron shell of element " + kv.first);
                const JValue* ft = sh.get("function_type");
                const std::string ftype = ft && ft->kind == JValue::Str ? ft->str : "gto";
                std::vector<double> exps(ex->arr.size());
                for (size_t k = 0; k < exps.size(); ++k)
                    if (!as_number(ex->arr[k], exps[k])) return fail("bad exponent of element " + kv.first);
                const size_t nrow = std::min(am->arr.size(), co->arr.size());   // zip(angular_momentum, coefficients)
                for (size_t r = 0; r < nrow; ++r) {
                    double lval = 0;
                    if (!as_number(am->arr[r], lval)) return fail("bad angular momentum of element " + kv.first);
                    const JValue& row = co->arr[r];
                    if (row.kind != JValue::Arr || row.arr.size() != exps.size()) return fail("coefficient row length mismatch, element " + kv.first);
                    ShellDef sd; sd.l = (int)lval; sd.ftype = ftype;
                    for (size_t k = 0; k < exps.size(); ++k) {
                        double c = 0;
                        if (!as_number(row.arr[k], c)) return fail("bad coefficient of element " + kv.first);
                        if (c != 0.0) { sd.exps.push_back(exps[k]); sd.coefs.push_back(c); }
                    }
                    shells.push_back(std::move(sd));
                }
            }
        table[z] = std::move(shells);
    }
    // MolecularSystem::load
    if (mdoc.kind != JValue::Arr) return fail("molecule file must be a JSON array of atoms");
    for (const JValue& a : mdoc.arr) {
        const JValue* el = a.get("element");
        const JValue* pos = a.get("position");
        double z = 0;
        if (!el || !pos || !as_number(*el, z) || pos->kind != JValue::Arr || pos->arr.size() != 3) return fail("malformed atom entry");
        const int Z = (int)z;
        const int ia = (int)sys->Z.size();
        sys->Z.push_back(Z);
        for (int k = 0; k < 3; ++k) {
            double x = 0;
            if (!as_number(pos->arr[k], x)) return fail("bad atom position");
            sys->xyz.push_back(x);
        }
        auto it = table.find(Z);
        if (it == table.end()) return fail("basis set has no element Z=" + std::to_string(Z));
        for (const ShellDef& sd : it->second) {
            if (sd.l > 2) return fail("angular momentum l=" + std::to_string(sd.l) + " > 2 is not supported");
            if (sd.l >= 2 && sd.ftype == "gto_spherical") return fail("spherical d shells are not supported (Cartesian 6d only)");
            if (sd.exps.empty()) return fail("shell without primitives");
            sys->shell_atom.push_back(ia);
            sys->shell_l.push_back(sd.l);
            sys->shell_nprim.push_back((int)sd.exps.size());
            sys->shell_prim_off.push_back((int)sys->exps.size());
            for (size_t k = 0; k < sd.exps.size(); ++k) {
                sys->exps.push_back(sd.exps[k]);
                sys->coefs.push_back(sd.coefs[k] * prim_norm(sd.exps[k], sd.l));
            }
            sys->n_basis += (sd.l + 1) * (sd.l + 2) / 2;
        }
        sys->n_electrons += Z;     // rhf.rs:36 (neutral molecules only)
    }
    if (sys->Z.empty() || sys->shell_l.empty()) return fail("empty molecule or basis");
    qcf_basis& b = sys->basis;
    b.n_atoms = (int)sys->Z.size(); b.Z = sys->Z.data(); b.xyz = sys->xyz.data();
    b.n_shells = (int)sys->shell_l.size(); b.shell_atom = sys->shell_atom.data(); b.shell_l = sys->shell_l.data();
    b.shell_nprim = sys->shell_nprim.data(); b.shell_prim_off = sys->shell_prim_off.data();
    b.exps = sys->exps.data(); b.coefs = sys->coefs.data(); b.cartesian = 1;
    return QCF_OK;
}

const qcf_basis* qcf_system_basis(const qcf_system* sys) { return sys && sys->err.empty() && sys->basis.n_shells > 0 ? &sys->basis : nullptr; }
int qcf_system_n_electrons(const qcf_system* sys) { return sys ? sys->n_electrons : QCF_ERR_ARG; }
int qcf_system_n_basis(const qcf_system* sys) { return sys ? sys->n_basis : QCF_ERR_ARG; }

// rhf.rs:110-122
double qcf_system_nuclear_repulsion(const qcf_system* sys) {
    if (!sys) return 0.0;
    double e = 0;
    const int n = (int)sys->Z.size();
    for (int i = 0; i < n; ++i)
        for (int j = i + 1; j < n; ++j) {
            double r2 = 0;
            for (int k = 0; k < 3; ++k) { const double d = sys->xyz[3 * j + k] - sys->xyz[3 * i + k]; r2 += d * d; }
            e += (double)(sys->Z[i] * sys->Z[j]) / std::sqrt(r2);
        }
    return e;
}

const char* qcf_system_error(const qcf_system* sys) { return sys ? sys->err.c_str() : "null system"; }
void qcf_system_free(qcf_system* sys) { delete sys; }

}  // extern "C"
