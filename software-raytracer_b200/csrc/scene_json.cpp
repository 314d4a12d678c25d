// scene_json.cpp - reader/writer for the reference's scene files (see scene_json.h).
#include "scene_json.h"

#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <utility>

namespace rtb {
namespace {

// ---- a small JSON value tree ------------------------------------------------------------
struct JVal {
    enum Type { Null, Bool, Num, Str, Arr, Obj } type = Null;
    bool b = false;
    double num = 0.0;
    std::string str;
    std::vector<JVal> arr;
    std::vector<std::pair<std::string, JVal>> obj;

    const JVal* find(const char* key) const {           // last duplicate wins, like a std::map assignment
        if (type != Obj) return nullptr;
        const JVal* r = nullptr;
        for (const auto& kv : obj) if (kv.first == key) r = &kv.second;
        return r;
    }
};

struct ParseError { std::string msg; };

class Parser {
public:
    explicit Parser(const std::string& s) : p_(s.data()), end_(s.data() + s.size()), begin_(s.data()) {}
    JVal parse_document() {
        // UTF-8 BOM is skipped, as the reference's parser does
        if (end_ - p_ >= 3 && (unsigned char)p_[0] == 0xEF && (unsigned char)p_[1] == 0xBB && (unsigned char)p_[2] == 0xBF) p_ += 3;
        JVal v = parse_value(0);
        skip_ws();
        if (p_ != end_) fail("unexpected trailing characters");
        return v;
    }

private:
    const char* p_; const char* end_; const char* begin_;

    [[noreturn]] void fail(const std::string& what) {
        throw ParseError{"parse error at byte " + std::to_string((long long)(p_ - begin_)) + ": " + what};
    }
    void skip_ws() { while (p_ < end_ && (*p_ == ' ' || *p_ == '\t' || *p_ == '\n' || *p_ == '\r')) ++p_; }
    bool match(const char* lit) {
        size_t n = strlen(lit);
        if ((size_t)(end_ - p_) >= n && memcmp(p_, lit, n) == 0) { p_ += n; return true; }
        return false;
    }
    static void append_utf8(std::string& out, unsigned cp) {
        if (cp < 0x80) out += (char)cp;
        else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
        else if (cp < 0x10000) { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
        else { out += (char)(0xF0 | (cp >> 18)); out += (char)(0x80 | ((cp >> 12) & 0x3F)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
    }
    unsigned parse_hex4() {
        if (end_ - p_ < 4) fail("truncated \\u escape");
        unsigned v = 0;
        for (int i = 0; i < 4; ++i) {
            char c = *p_++;
            v <<= 4;
            if (c >= '0' && c <= '9') v |= (unsigned)(c - '0');
            else if (c >= 'a' && c <= 'f') v |= (unsigned)(c - 'a' + 10);
            else if (c >= 'A' && c <= 'F') v |= (unsigned)(c - 'A' + 10);
            else fail("bad \\u escape");
        }
        return v;
    }
    std::string parse_string() {
        ++p_;  // opening quote
        std::string out;
        for (;;) {
            if (p_ >= end_) fail("unterminated string");
            unsigned char c = (unsigned char)*p_++;
            if (c == '"') return out;
            if (c < 0x20) fail("control character in string");
            if (c != '\\') { out += (char)c; continue; }
            if (p_ >= end_) fail("unterminated escape");
            char e = *p_++;
            switch (e) {
                case '"': out += '"'; break;
                case '\\': out += '\\'; break;
                case '/': out += '/'; break;
                case 'b': out += '\b'; break;
                case 'f': out += '\f'; break;
                case 'n': out += '\n'; break;
                case 'r': out += '\r'; break;
                case 't': out += '\t'; break;
                case 'u': {
                    unsigned cp = parse_hex4();
                    if (cp >= 0xD800 && cp <= 0xDBFF) {
                        if (!(end_ - p_ >= 2 && p_[0] == '\\' && p_[1] == 'u')) fail("lone surrogate");
                        p_ += 2;
                        unsigned lo = parse_hex4();
                        if (lo < 0xDC00 || lo > 0xDFFF) fail("bad low surrogate");
                        cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                    } else if (cp >= 0xDC00 && cp <= 0xDFFF) fail("lone surrogate");
                    append_utf8(out, cp);
                    break;
                }
                default: fail("bad escape");
            }
        }
    }
    JVal parse_number() {
        const char* s = p_;
        if (p_ < end_ && *p_ == '-') ++p_;
        if (p_ >= end_) fail("bad number");
        if (*p_ == '0') ++p_;
        else if (*p_ >= '1' && *p_ <= '9') { while (p_ < end_ && *p_ >= '0' && *p_ <= '9') ++p_; }
        else fail("bad number");
        if (p_ < end_ && *p_ == '.') {
            ++p_;
            if (p_ >= end_ || *p_ < '0' || *p_ > '9') fail("bad fraction");
            while (p_ < end_ && *p_ >= '0' && *p_ <= '9') ++p_;
        }
        if (p_ < end_ && (*p_ == 'e' || *p_ == 'E')) {
            ++p_;
            if (p_ < end_ && (*p_ == '+' || *p_ == '-')) ++p_;
            if (p_ >= end_ || *p_ < '0' || *p_ > '9') fail("bad exponent");
            while (p_ < end_ && *p_ >= '0' && *p_ <= '9') ++p_;
        }
        std::string tok(s, p_);
        JVal v; v.type = JVal::Num;
        v.num = strtod(tok.c_str(), nullptr);   // correctly rounded decimal -> double, then (float) at use
        return v;
    }
    JVal parse_value(int depth) {
        if (depth > 256) fail("nesting too deep");
        skip_ws();
        if (p_ >= end_) fail("unexpected end of input");
        JVal v;
        char c = *p_;
        if (c == '{') {
            ++p_; v.type = JVal::Obj;
            skip_ws();
            if (p_ < end_ && *p_ == '}') { ++p_; return v; }
            for (;;) {
                skip_ws();
                if (p_ >= end_ || *p_ != '"') fail("expected object key");
                std::string key = parse_string();
                skip_ws();
                if (p_ >= end_ || *p_ != ':') fail("expected ':'");
                ++p_;
                JVal child = parse_value(depth + 1);
                v.obj.emplace_back(std::move(key), std::move(child));
                skip_ws();
                if (p_ < end_ && *p_ == ',') { ++p_; continue; }
                if (p_ < end_ && *p_ == '}') { ++p_; return v; }
                fail("expected ',' or '}'");
            }
        }
        if (c == '[') {
            ++p_; v.type = JVal::Arr;
            skip_ws();
            if (p_ < end_ && *p_ == ']') { ++p_; return v; }
            for (;;) {
                v.arr.push_back(parse_value(depth + 1));
                skip_ws();
                if (p_ < end_ && *p_ == ',') { ++p_; continue; }
                if (p_ < end_ && *p_ == ']') { ++p_; return v; }
                fail("expected ',' or ']'");
            }
        }
        if (c == '"') { v.type = JVal::Str; v.str = parse_string(); return v; }
        if (c == '-' || (c >= '0' && c <= '9')) return parse_number();
        if (match("true")) { v.type = JVal::Bool; v.b = true; return v; }
        if (match("false")) { v.type = JVal::Bool; v.b = false; return v; }
        if (match("null")) return v;
        fail("unexpected character");
    }
};

// ---- typed access with the reference's failure modes ---------------------------------------
struct EntryError { std::string msg; };

float as_float(const JVal* v, const char* what) {
    if (!v) throw EntryError{std::string("missing ") + what};
    if (v->type == JVal::Num) return (float)v->num;
    if (v->type == JVal::Bool) return v->b ? 1.f : 0.f;   // json bool -> arithmetic is allowed by the reference's library
    throw EntryError{std::string(what) + " is not a number"};
}
void as_float3(const JVal* v, const char* what, float out[3]) {
    if (!v) throw EntryError{std::string("missing ") + what};
    if (v->type != JVal::Arr || v->arr.size() < 3) throw EntryError{std::string(what) + " is not an array of 3 numbers"};
    for (int i = 0; i < 3; ++i) out[i] = as_float(&v->arr[i], what);
}
inline float clamp0(float v) { return v < 0 ? 0 : v; }     // Color ctor (Common.hpp:253-262)
void as_color(const JVal* v, const char* what, float out[3]) {
    as_float3(v, what, out);
    for (int i = 0; i < 3; ++i) out[i] = clamp0(out[i]);
}

// ---- writer: nlohmann dump(4) look-alike ------------------------------------------------------
void dump_string(std::string& out, const std::string& s) {
    out += '"';
    for (unsigned char c : s) {
        switch (c) {
            case '"': out += "\\\""; break;
            case '\\': out += "\\\\"; break;
            case '\b': out += "\\b"; break;
            case '\f': out += "\\f"; break;
            case '\n': out += "\\n"; break;
            case '\r': out += "\\r"; break;
            case '\t': out += "\\t"; break;
            default:
                if (c < 0x20) { char buf[8]; snprintf(buf, sizeof buf, "\\u%04x", c); out += buf; }
                else out += (char)c;
        }
    }
    out += '"';
}

// Shortest round-trip digits of a double, laid out the way the reference's library prints
// number_float: integral values get ".0", fixed notation for decimal exponents in (-4, 15],
// otherwise d.ddde[+-]XX with at least two exponent digits.
void dump_number(std::string& out, float f) {
    double v = (double)f;                        // json stores number_float as double (Object.hpp:32-40)
    if (!std::isfinite(v)) { out += "null"; return; }
    if (v == 0) { out += std::signbit(v) ? "-0.0" : "0.0"; return; }
    char buf[64];
    auto res = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::scientific);
    std::string sci(buf, res.ptr);               // [-]d[.ddd]e[+-]XX
    bool neg = sci[0] == '-';
    size_t epos = sci.find('e');
    std::string mant = sci.substr(neg ? 1 : 0, epos - (neg ? 1 : 0));
    int exp10 = atoi(sci.c_str() + epos + 1);
    std::string digits;
    for (char c : mant) if (c != '.') digits += c;
    int k = (int)digits.size();                  // number of significant digits
    // The reference's library (Grisu2) and std::to_chars (shortest, round-half-even) agree except
    // when the exact binary value lies exactly half-way between two k-digit candidates: Grisu2
    // keeps the candidate of larger magnitude (e.g. -1001.20001220703125 -> ...0313). Detect the
    // tie on the exact decimal expansion and round half away from zero.
    {
        char exact[400];
        snprintf(exact, sizeof exact, "%.330e", std::fabs(v));      // glibc prints the exact expansion
        std::string ed;
        for (const char* q = exact; *q && *q != 'e'; ++q) if (*q != '.') ed += *q;
        bool tie = (int)ed.size() > k && ed[(size_t)k] == '5';
        for (size_t i = (size_t)k + 1; tie && i < ed.size(); ++i) if (ed[i] != '0') tie = false;
        if (tie && ed.compare(0, (size_t)k, digits) == 0) {          // to_chars rounded down: bump the last digit
            int i = k - 1;
            while (i >= 0 && digits[(size_t)i] == '9') { digits[(size_t)i] = '0'; --i; }
            if (i >= 0) digits[(size_t)i]++;
            else { digits.insert(digits.begin(), '1'); digits.pop_back(); ++exp10; }
        }
    }
    int n = exp10 + 1;                           // position of the decimal point relative to the digits
    if (neg) out += '-';
    const int min_exp = -4, max_exp = 15;
    if (k <= n && n <= max_exp) {                // digits[000].0
        out += digits; out.append((size_t)(n - k), '0'); out += ".0";
    } else if (0 < n && n <= max_exp) {          // dig.its
        out += digits.substr(0, (size_t)n); out += '.'; out += digits.substr((size_t)n);
    } else if (min_exp < n && n <= 0) {          // 0.[000]digits
        out += "0."; out.append((size_t)(-n), '0'); out += digits;
    } else {                                     // d[.igits]e+XX
        out += digits[0];
        if (k > 1) { out += '.'; out += digits.substr(1); }
        out += 'e';
        int e = n - 1;
        out += e < 0 ? '-' : '+';
        if (e < 0) e = -e;
        char eb[8]; snprintf(eb, sizeof eb, e < 10 ? "0%d" : "%d", e);
        out += eb;
    }
}

void indent(std::string& out, int n) { out.append((size_t)n, ' '); }
void dump_vec3(std::string& out, const float v[3], int ind) {
    out += "[\n";
    for (int i = 0; i < 3; ++i) {
        indent(out, ind + 4); dump_number(out, v[i]); out += i < 2 ? ",\n" : "\n";
    }
    indent(out, ind); out += "]";
}

}  // namespace

rt_object make_object(int type) {
    rt_object o;
    memset(&o, 0, sizeof o);
    o.type = type;
    o.smoothness = 0.5f; o.spec_amount = 0.0f;              // Material() Common.hpp:313-318
    for (int i = 0; i < 3; ++i) { o.base[i] = 1.f; o.spec_color[i] = 1.f; o.emissive[i] = 0.f; }
    return o;
}

int HostScene::LoadFromString(const std::string& text, std::string& err) {
    Unload();                                                // sceneObjects.clear() Scene.hpp:28
    JVal root;
    try {
        root = Parser(text).parse_document();
    } catch (const ParseError& e) { err = e.msg; return RT_ERR_PARSE; }
    try {
        const JVal* sn = root.find("SceneName");             // Scene.hpp:35
        if (!sn || sn->type != JVal::Str) throw EntryError{"SceneName missing or not a string"};
        scene_name = sn->str;
        const JVal* list = root.find("SceneObjects");        // Scene.hpp:36
        if (!list || (list->type != JVal::Arr && list->type != JVal::Obj)) return RT_OK;   // null -> empty iteration
        std::vector<const JVal*> entries;
        if (list->type == JVal::Arr) for (const JVal& e : list->arr) entries.push_back(&e);
        else for (const auto& kv : list->obj) entries.push_back(&kv.second);
        for (size_t i = 0; i < entries.size(); ++i) {
            const JVal& value = *entries[i];
            try {
                float pos[3];
                HostMesh mesh;
                as_float3(value.find("Position"), "Position", pos);          // Scene.hpp:40,57
                const JVal* rend = value.find("Renderer");
                const JVal* type = rend ? rend->find("Type") : nullptr;
                rt_object o;
                if (type && type->type == JVal::Str && type->str == "Sphere") {             // Scene.hpp:43-45
                    o = make_object(RT_OBJ_SPHERE);
                    o.radius = as_float(rend->find("Radius"), "Renderer.Radius");
                } else if (type && type->type == JVal::Str && type->str == "Cube") {        // Scene.hpp:46-52
                    o = make_object(RT_OBJ_CUBE);
                    as_float3(rend->find("Size"), "Renderer.Size", o.half);
                } else if (type && type->type == JVal::Str && type->str == "Mesh") {        // extension (mesh.h)
                    o = make_object(RT_OBJ_MESH);
                    const JVal* file = rend->find("File");
                    if (file && file->type == JVal::Str) {
                        std::string path = file->str;
                        if (!path.empty() && path[0] != '/') {                              // relative to the scene file
                            const size_t slash = file_name.find_last_of('/');
                            if (slash != std::string::npos) path = file_name.substr(0, slash + 1) + path;
                        }
                        std::string merr;
                        if (!load_obj(path, mesh, merr)) throw EntryError{merr};
                        mesh.file = file->str;
                    } else {
                        const JVal* vs = rend->find("Vertices"); const JVal* ts = rend->find("Triangles");
                        if (!vs || vs->type != JVal::Arr || !ts || ts->type != JVal::Arr) throw EntryError{"Mesh needs File or Vertices + Triangles"};
                        for (const JVal& x : vs->arr) mesh.vertices.push_back(as_float(&x, "Renderer.Vertices"));
                        for (const JVal& x : ts->arr) mesh.indices.push_back((int32_t)as_float(&x, "Renderer.Triangles"));
                        if (mesh.vertices.size() % 3 || mesh.indices.size() % 3) throw EntryError{"Mesh arrays must hold triples"};
                    }
                } else {
                    o = make_object(RT_OBJ_NONE);                                           // Scene.hpp:53-55
                }
                memcpy(o.pos, pos, sizeof pos);
                if (const JVal* m = value.find("Material")) {                               // Scene.hpp:59-69
                    const JVal* v;
                    o.smoothness = (v = m->find("Smoothness")) ? as_float(v, "Material.Smoothness") : 0.5f;
                    o.spec_amount = (v = m->find("SpecularAmount")) ? as_float(v, "Material.SpecularAmount") : 0.1f;
                    if ((v = m->find("SpecularColor"))) as_color(v, "Material.SpecularColor", o.spec_color);
                    else o.spec_color[0] = o.spec_color[1] = o.spec_color[2] = 1.f;
                    if ((v = m->find("Color"))) as_color(v, "Material.Color", o.base);
                    else o.base[0] = o.base[1] = o.base[2] = 1.f;
                    if ((v = m->find("Emissive"))) as_color(v, "Material.Emissive", o.emissive);
                    else o.emissive[0] = o.emissive[1] = o.emissive[2] = 0.f;
                }
                const JVal* name = value.find("Name");                                      // Scene.hpp:71
                if (!name || name->type != JVal::Str) throw EntryError{"Name missing or not a string"};
                objects.push_back(o);
                names.push_back(name->str);
                meshes.resize(objects.size());
                meshes.back() = std::move(mesh);
            } catch (const EntryError& e) {
                // Scene.hpp:75-77: the exception ends the load; earlier objects stay.
                err = "SceneObjects[" + std::to_string(i) + "]: " + e.msg;
                return RT_ERR_PARSE;
            }
        }
    } catch (const EntryError& e) { err = e.msg; return RT_ERR_PARSE; }
    return RT_OK;
}

int HostScene::Load(const std::string& path, std::string& err) {
    file_name = path;
    Unload();
    std::ifstream f(path, std::ios::binary);
    if (!f.good()) { err = "cannot open scene file: " + path; return RT_ERR_IO; }   // Scene.hpp:30-32
    std::stringstream ss; ss << f.rdbuf();
    return LoadFromString(ss.str(), err);
}

std::string HostScene::Dump() const {
    std::string out = "{\n";
    indent(out, 4); out += "\"SceneName\": "; dump_string(out, scene_name); out += ",\n";
    indent(out, 4); out += "\"SceneObjects\": ";
    if (objects.empty()) { out += "[]\n}"; return out; }
    out += "[\n";
    for (size_t i = 0; i < objects.size(); ++i) {
        const rt_object& o = objects[i];
        indent(out, 8); out += "{\n";
        indent(out, 12); out += "\"Material\": {\n";
        indent(out, 16); out += "\"Color\": "; dump_vec3(out, o.base, 16); out += ",\n";
        indent(out, 16); out += "\"Emissive\": "; dump_vec3(out, o.emissive, 16); out += ",\n";
        indent(out, 16); out += "\"Metalness\": "; dump_number(out, o.spec_amount); out += ",\n";   // Object.hpp:33
        indent(out, 16); out += "\"Smoothness\": "; dump_number(out, o.smoothness); out += ",\n";
        indent(out, 16); out += "\"SpecularAmount\": "; dump_number(out, o.spec_amount); out += ",\n";
        indent(out, 16); out += "\"SpecularColor\": "; dump_vec3(out, o.spec_color, 16); out += "\n";
        indent(out, 12); out += "},\n";
        indent(out, 12); out += "\"Name\": "; dump_string(out, i < names.size() ? names[i] : std::string()); out += ",\n";
        indent(out, 12); out += "\"Position\": "; dump_vec3(out, o.pos, 12); out += ",\n";
        indent(out, 12); out += "\"Renderer\": {\n";
        if (o.type == RT_OBJ_SPHERE) {
            indent(out, 16); out += "\"Radius\": "; dump_number(out, o.radius); out += ",\n";
            indent(out, 16); out += "\"Type\": \"Sphere\"\n";
        } else if (o.type == RT_OBJ_CUBE) {
            indent(out, 16); out += "\"Size\": "; dump_vec3(out, o.half, 16); out += ",\n";
            indent(out, 16); out += "\"Type\": \"Cube\"\n";
        } else if (o.type == RT_OBJ_MESH && i < meshes.size()) {
            const HostMesh& m = meshes[i];
            if (!m.file.empty()) {
                indent(out, 16); out += "\"File\": "; dump_string(out, m.file); out += ",\n";
                indent(out, 16); out += "\"Type\": \"Mesh\"\n";
            } else {
                indent(out, 16); out += "\"Triangles\": [";
                for (size_t k = 0; k < m.indices.size(); ++k) { if (k) out += ", "; out += std::to_string(m.indices[k]); }
                out += "],\n";
                indent(out, 16); out += "\"Type\": \"Mesh\",\n";
                indent(out, 16); out += "\"Vertices\": [";
                for (size_t k = 0; k < m.vertices.size(); ++k) { if (k) out += ", "; dump_number(out, m.vertices[k]); }
                out += "]\n";
            }
        } else {
            indent(out, 16); out += "\"Type\": \"None\"\n";
        }
        indent(out, 12); out += "}\n";
        indent(out, 8); out += i + 1 < objects.size() ? "},\n" : "}\n";
    }
    indent(out, 4); out += "]\n}";
    return out;
}

int HostScene::SaveAs(const std::string& path, std::string& err) {
    file_name = path;                                        // Scene.hpp:101-104
    std::ofstream f(path, std::ios::binary);
    if (!f.good()) { err = "cannot write scene file: " + path; return RT_ERR_IO; }
    f << Dump();
    return RT_OK;
}

}  // namespace rtb
