// cq_json.cpp — *.static.json loader (host only, dependency-free).
// Replaces StaticMeshLoader.loadStaticMeshAsset(named:) / buildAsset (StaticMeshLoader.swift:30-125);
// schema per StaticMeshLoader.swift:168-197:
//   {"version":1,"meshes":[{"name":..,"transform":[16 row-major],"mesh":{"positions":[3V],"normals":[..],
//     "uvs":[..],"indices":[3T],"submeshes":[{start,count,material}]?},"collisionHulls":[{positions,indices}]?}]}
// JSON numbers are parsed as doubles and narrowed to float, as Foundation's JSONDecoder does for Float.
// Parts with invalid positions / no indices are skipped (the reference prints and continues, :53-61);
// a missing file gives CQ_ERR_IO and malformed JSON CQ_ERR_PARSE where the reference returns nil.
#include <cctype>
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../../include/cq.h"

namespace cq {
void set_error(const char *fmt, ...);
}

namespace {

struct JValue {
    enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
    bool b = false;
    double num = 0;
    std::string str;
    std::vector<double> nums;          // fast path: array of plain numbers
    bool numericArray = false;
    std::vector<JValue> items;         // generic array
    std::vector<std::pair<std::string, JValue>> members;
    const JValue *get(const char *key) const {
        for (auto &m : members)
            if (m.first == key) return &m.second;
        return nullptr;
    }
};

struct Parser {
    const char *p, *end;
    std::string err;
    void ws() {
        while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) p++;
    }
    bool fail(const char *m) {
        if (err.empty()) err = m;
        return false;
    }
    bool parseString(std::string &out) {
        if (p >= end || *p != '"') return fail("expected string");
        p++;
        out.clear();
        while (p < end && *p != '"') {
            if (*p == '\\') {
                p++;
                if (p >= end) return fail("bad escape");
                switch (*p) {
                case 'n': out += '\n'; break;
                case 't': out += '\t'; break;
                case 'r': out += '\r'; break;
                case 'b': out += '\b'; break;
                case 'f': out += '\f'; break;
                case 'u': { // keep BMP code points as UTF-8
                    if (end - p < 5) return fail("bad \\u escape");
                    for (int k = 1; k <= 4; k++)
                        if (!isxdigit((unsigned char)p[k])) return fail("bad \\u escape");
                    unsigned cp = (unsigned)strtoul(std::string(p + 1, p + 5).c_str(), nullptr, 16);
                    p += 4;
                    if (cp < 0x80) out += (char)cp;
                    else if (cp < 0x800) {
                        out += (char)(0xC0 | (cp >> 6));
                        out += (char)(0x80 | (cp & 0x3F));
                    } else {
                        out += (char)(0xE0 | (cp >> 12));
                        out += (char)(0x80 | ((cp >> 6) & 0x3F));
                        out += (char)(0x80 | (cp & 0x3F));
                    }
                    break;
                }
                default: out += *p; break;
                }
                p++;
            } else {
                out += *p++;
            }
        }
        if (p >= end) return fail("unterminated string");
        p++;
        return true;
    }
    // RFC 8259 number grammar only: -? (0 | [1-9][0-9]*) (. [0-9]+)? ([eE] [+-]? [0-9]+)?  — strtod alone would also take
    // "inf", "nan", hex floats, a leading '+' or leading zeros, none of which JSONDecoder accepts
    bool parseNumber(double &out) {
        const char *q = p;
        if (q < end && *q == '-') q++;
        if (q >= end || *q < '0' || *q > '9') return fail("expected number");
        if (*q == '0') q++;
        else
            while (q < end && *q >= '0' && *q <= '9') q++;
        if (q < end && *q == '.') {
            q++;
            if (q >= end || *q < '0' || *q > '9') return fail("bad number");
            while (q < end && *q >= '0' && *q <= '9') q++;
        }
        if (q < end && (*q == 'e' || *q == 'E')) {
            q++;
            if (q < end && (*q == '+' || *q == '-')) q++;
            if (q >= end || *q < '0' || *q > '9') return fail("bad number");
            while (q < end && *q >= '0' && *q <= '9') q++;
        }
        char *e = nullptr;
        out = strtod(p, &e); // the buffer is NUL-terminated (std::string); the token is a prefix strtod accepts in full
        if (e != q) return fail("bad number");
        p = q;
        return true;
    }
    bool parseValue(JValue &v, int depth) {
        if (depth > 64) return fail("nesting too deep");
        ws();
        if (p >= end) return fail("unexpected end");
        char c = *p;
        if (c == '{') {
            v.kind = JValue::Obj;
            p++;
            ws();
            if (p < end && *p == '}') {
                p++;
                return true;
            }
            while (true) {
                ws();
                std::string key;
                if (!parseString(key)) return false;
                ws();
                if (p >= end || *p != ':') return fail("expected ':'");
                p++;
                v.members.emplace_back(key, JValue());
                if (!parseValue(v.members.back().second, depth + 1)) return false;
                ws();
                if (p < end && *p == ',') {
                    p++;
                    continue;
                }
                if (p < end && *p == '}') {
                    p++;
                    return true;
                }
                return fail("expected ',' or '}'");
            }
        }
        if (c == '[') {
            v.kind = JValue::Arr;
            p++;
            ws();
            if (p < end && *p == ']') {
                p++;
                v.numericArray = true;
                return true;
            }
            // numeric fast path
            if (p < end && (*p == '-' || (*p >= '0' && *p <= '9'))) {
                v.numericArray = true;
                while (true) {
                    ws();
                    if (p < end && *p != '-' && (*p < '0' || *p > '9')) {
                        // not a plain number array after all (only unknown keys can hold such a thing): carry on generically
                        v.numericArray = false;
                        v.items.resize(v.nums.size());
                        for (size_t k = 0; k < v.nums.size(); k++) v.items[k].kind = JValue::Num, v.items[k].num = v.nums[k];
                        v.nums.clear();
                        break;
                    }
                    double d;
                    if (!parseNumber(d)) return false;
                    v.nums.push_back(d);
                    ws();
                    if (p < end && *p == ',') {
                        p++;
                        continue;
                    }
                    if (p < end && *p == ']') {
                        p++;
                        return true;
                    }
                    return fail("expected ',' or ']' in number array");
                }
            }
            while (true) {
                v.items.emplace_back();
                if (!parseValue(v.items.back(), depth + 1)) return false;
                ws();
                if (p < end && *p == ',') {
                    p++;
                    continue;
                }
                if (p < end && *p == ']') {
                    p++;
                    return true;
                }
                return fail("expected ',' or ']'");
            }
        }
        if (c == '"') {
            v.kind = JValue::Str;
            return parseString(v.str);
        }
        if (!strncmp(p, "true", 4) && end - p >= 4) {
            v.kind = JValue::Bool, v.b = true, p += 4;
            return true;
        }
        if (!strncmp(p, "false", 5) && end - p >= 5) {
            v.kind = JValue::Bool, v.b = false, p += 5;
            return true;
        }
        if (!strncmp(p, "null", 4) && end - p >= 4) {
            v.kind = JValue::Null, p += 4;
            return true;
        }
        v.kind = JValue::Num;
        return parseNumber(v.num);
    }
};

struct Geometry {
    std::vector<float> positions;
    std::vector<uint32_t> indices;
};
struct Part {
    std::string name;
    float transform[16]; // column-major
    Geometry mesh;
    std::vector<Geometry> hulls;
};

bool numberArray(const JValue *v, std::vector<double> &out) {
    if (!v || v->kind != JValue::Arr) return false;
    if (v->numericArray) {
        out = v->nums;
        return true;
    }
    out.clear();
    for (auto &it : v->items) {
        if (it.kind != JValue::Num) return false;
        out.push_back(it.num);
    }
    return true;
}

bool indexArray(const std::vector<double> &in, std::vector<uint32_t> &out) {
    out.resize(in.size());
    for (size_t i = 0; i < in.size(); i++) {
        double d = in[i];
        if (!(d >= 0.0) || d > 4294967295.0 || d != std::floor(d)) return false; // [UInt32] decode would throw
        out[i] = (uint32_t)d;
    }
    return true;
}

} // namespace

struct cq_static_mesh_asset {
    std::vector<Part> parts;
};

extern "C" {

int cq_static_mesh_load(const char *path, cq_static_mesh_asset **out) {
    if (!path || !out) return CQ_ERR_INVALID;
    *out = nullptr;
    FILE *f = fopen(path, "rb");
    if (!f) {
        cq::set_error("StaticMeshLoader: missing json: %s", path);
        return CQ_ERR_IO;
    }
    std::string text;
    char buf[1 << 16];
    size_t got;
    while ((got = fread(buf, 1, sizeof(buf), f)) > 0) text.append(buf, got);
    fclose(f);
    Parser ps{text.data(), text.data() + text.size(), {}};
    JValue root;
    if (!ps.parseValue(root, 0) || root.kind != JValue::Obj) {
        cq::set_error("StaticMeshLoader: failed to load json: %s (%s)", path, ps.err.empty() ? "not an object" : ps.err.c_str());
        return CQ_ERR_PARSE;
    }
    ps.ws();
    if (ps.p != ps.end) { // JSONDecoder: "garbage at end"
        cq::set_error("StaticMeshLoader: failed to load json: %s (trailing characters after the document)", path);
        return CQ_ERR_PARSE;
    }
    const JValue *version = root.get("version"), *meshes = root.get("meshes");
    if (!version || version->kind != JValue::Num || version->num != std::floor(version->num) /* `let version: Int` */ || !meshes || meshes->kind != JValue::Arr || (meshes->numericArray && !meshes->nums.empty())) {
        cq::set_error("StaticMeshLoader: failed to load json: %s (missing version/meshes)", path);
        return CQ_ERR_PARSE;
    }
    std::unique_ptr<cq_static_mesh_asset> asset(new cq_static_mesh_asset());
    for (const JValue &entry : meshes->items) {
        if (entry.kind != JValue::Obj) {
            cq::set_error("StaticMeshLoader: mesh entry is not an object");
            return CQ_ERR_PARSE;
        }
        const JValue *name = entry.get("name"), *transform = entry.get("transform"), *mesh = entry.get("mesh");
        std::vector<double> tr, pos, idx, tmp;
        // required keys of the Codable structs (StaticMeshLoader.swift:173-185): a missing one fails the whole decode
        if (!name || name->kind != JValue::Str || !numberArray(transform, tr) || !mesh || mesh->kind != JValue::Obj ||
            !numberArray(mesh->get("positions"), pos) || !numberArray(mesh->get("normals"), tmp) ||
            !numberArray(mesh->get("uvs"), tmp) || !numberArray(mesh->get("indices"), idx)) {
            cq::set_error("StaticMeshLoader: failed to load json: %s (mesh entry is missing a required key)", path);
            return CQ_ERR_PARSE;
        }
        Part part;
        part.name = name->str;
        size_t vCount = pos.size() / 3;
        if (vCount == 0 || pos.size() != vCount * 3) continue; // "invalid positions" -> part skipped (:53-57)
        if (idx.empty()) continue;                             // "missing indices"   -> part skipped (:58-61)
        part.mesh.positions.resize(pos.size());
        for (size_t i = 0; i < pos.size(); i++) part.mesh.positions[i] = (float)pos[i];
        if (!indexArray(idx, part.mesh.indices)) {
            cq::set_error("StaticMeshLoader: failed to load json: %s (index is not a UInt32)", path);
            return CQ_ERR_PARSE;
        }
        if (tr.size() == 16) { // matrixFromArrayRowMajor (:127-134): row-major file -> column-major simd
            for (int r = 0; r < 4; r++)
                for (int c = 0; c < 4; c++) part.transform[c * 4 + r] = (float)tr[r * 4 + c];
        } else { // matrix_identity_float4x4 (:115)
            for (int k = 0; k < 16; k++) part.transform[k] = (k % 5 == 0) ? 1.0f : 0.0f;
        }
        const JValue *hulls = entry.get("collisionHulls");
        if (hulls && hulls->kind == JValue::Arr) { // buildCollisionHulls (:136-165)
            for (const JValue &h : hulls->items) {
                std::vector<double> hp, hi;
                if (h.kind != JValue::Obj || !numberArray(h.get("positions"), hp) || !numberArray(h.get("indices"), hi)) {
                    cq::set_error("StaticMeshLoader: failed to load json: %s (bad collision hull)", path);
                    return CQ_ERR_PARSE;
                }
                size_t hv = hp.size() / 3;
                if (hv == 0 || hp.size() != hv * 3) continue;
                if (hi.empty()) continue;
                Geometry g;
                g.positions.resize(hp.size());
                for (size_t i = 0; i < hp.size(); i++) g.positions[i] = (float)hp[i];
                if (!indexArray(hi, g.indices)) {
                    cq::set_error("StaticMeshLoader: failed to load json: %s (hull index is not a UInt32)", path);
                    return CQ_ERR_PARSE;
                }
                part.hulls.push_back(std::move(g));
            }
        }
        asset->parts.push_back(std::move(part));
    }
    *out = asset.release();
    return CQ_OK;
}

void cq_static_mesh_free(cq_static_mesh_asset *a) { delete a; }

int32_t cq_static_mesh_part_count(const cq_static_mesh_asset *a) { return a ? (int32_t)a->parts.size() : 0; }

const char *cq_static_mesh_part_name(const cq_static_mesh_asset *a, int32_t part) {
    if (!a || part < 0 || part >= (int32_t)a->parts.size()) return nullptr;
    return a->parts[part].name.c_str();
}

int cq_static_mesh_part_transform(const cq_static_mesh_asset *a, int32_t part, float out_colmajor[16]) {
    if (!a || !out_colmajor || part < 0 || part >= (int32_t)a->parts.size()) return CQ_ERR_INVALID;
    memcpy(out_colmajor, a->parts[part].transform, sizeof(float) * 16);
    return CQ_OK;
}

int32_t cq_static_mesh_hull_count(const cq_static_mesh_asset *a, int32_t part) {
    if (!a || part < 0 || part >= (int32_t)a->parts.size()) return 0;
    return (int32_t)a->parts[part].hulls.size();
}

int cq_static_mesh_geometry(const cq_static_mesh_asset *a, int32_t part, int32_t hull, const float **positions_xyz,
                            int32_t *n_verts, const uint32_t **indices, int32_t *n_indices) {
    if (!a || part < 0 || part >= (int32_t)a->parts.size()) return CQ_ERR_INVALID;
    const Part &p = a->parts[part];
    if (hull < -1 || hull >= (int32_t)p.hulls.size()) return CQ_ERR_INVALID;
    const Geometry &g = hull < 0 ? p.mesh : p.hulls[hull];
    if (positions_xyz) *positions_xyz = g.positions.data();
    if (n_verts) *n_verts = (int32_t)(g.positions.size() / 3);
    if (indices) *indices = g.indices.data();
    if (n_indices) *n_indices = (int32_t)g.indices.size();
    return CQ_OK;
}

} // extern "C"
