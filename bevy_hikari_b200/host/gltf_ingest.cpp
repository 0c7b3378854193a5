// gltf_ingest.cpp — SURVEY 8(f) rank 4: run-time scene ingest for the host mirror.  What a Bevy app gets from
// `asset_server.load("models/x.glb#Scene0")` + the extraction of src/mesh_material/{mesh,material,instance}.rs — one hikari::Mesh per
// glTF primitive (POSITION / NORMAL / TEXCOORD_0 / indices; triangle lists and strips, mod.rs:433-450), one StandardMaterial per glTF
// material (material.rs:139-203: base colour / emissive / roughness / metallic factors and the five texture slots), one instance per
// (node, primitive) with the node's world matrix (glam f32 arithmetic, depth-first node order), one texture per (image, colour space)
// with the sampler's wrap modes and filter (material.rs:55-87: base colour and emissive are sRGB, the other slots linear) — straight
// into a MeshMaterialWorld, from .glb or .gltf (side-car or data: URI buffers).  Same rules as the offline converter
// tools/make_assets.py, which produced scenes/*.npz; tests/test_gltf_ingest.py holds the two against each other.
// PNG images are decoded here (zlib inflate + the five scanline filters; 8 / 16-bit grey, grey+alpha, RGB, RGBA, palette;
// non-interlaced); any other encoding (JPEG) goes through the caller's decoder, as image decoding is the host application's job in
// the reference too (bevy_render's image loader).  Pure CPU; part of libhikari_host.so.
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <zlib.h>

#include <map>
#include <string>
#include <utility>
#include <vector>

#include "hikari.hpp"
#include "hikari_host.h"

namespace hikari {
namespace {

// ------------------------------------------------------------------------------------------------ JSON
struct JVal {
    enum Type { Null, Bool, Num, Str, Arr, Obj } type = Null;
    bool b = false;
    double num = 0.0;
    std::string str;
    std::vector<JVal> arr;
    std::vector<std::pair<std::string, JVal>> obj;
    const JVal* get(const char* key) const {
        if (type != Obj) return nullptr;
        for (const auto& kv : obj)
            if (kv.first == key) return &kv.second;
        return nullptr;
    }
    double number(const char* key, double dflt) const { const JVal* v = get(key); return v && v->type == Num ? v->num : dflt; }
    long index(const char* key) const { const JVal* v = get(key); return v && v->type == Num ? (long)v->num : -1; }
    size_t size() const { return type == Arr ? arr.size() : 0; }
};

struct JParser {
    const char* p; const char* end; bool ok = true; int depth = 0;
    void ws() { while (p < end && (*p == ' ' || *p == '\n' || *p == '\r' || *p == '\t')) ++p; }
    bool lit(const char* s) { size_t n = strlen(s); if ((size_t)(end - p) >= n && memcmp(p, s, n) == 0) { p += n; return true; } return false; }
    static void utf8(std::string& out, unsigned cp) {
        if (cp < 0x80) out += (char)cp;
        else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
        else if (cp < 0x10000) { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
        else { out += (char)(0xF0 | (cp >> 18)); out += (char)(0x80 | ((cp >> 12) & 0x3F)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
    }
    std::string string() {
        std::string s;
        ++p;   // opening quote
        while (p < end && *p != '"') {
            if (*p == '\\' && p + 1 < end) {
                ++p;
                switch (*p) {
                    case 'n': s += '\n'; break; case 't': s += '\t'; break; case 'r': s += '\r'; break;
                    case 'b': s += '\b'; break; case 'f': s += '\f'; break;
                    case 'u': {
                        unsigned cp = 0;
                        for (int i = 1; i <= 4 && p + i < end; ++i) {
                            char c = p[i];
                            cp = cp * 16 + (unsigned)(c >= '0' && c <= '9' ? c - '0' : (c | 32) - 'a' + 10);
                        }
                        p += 4;
                        utf8(s, cp);
                        break;
                    }
                    default: s += *p;
                }
                ++p;
            } else {
                s += *p++;
            }
        }
        if (p >= end) ok = false; else ++p;
        return s;
    }
    JVal value() {
        JVal v;
        ws();
        if (p >= end || ++depth > 256) { ok = false; return v; }
        if (*p == '{') {
            v.type = JVal::Obj; ++p; ws();
            if (p < end && *p == '}') { ++p; --depth; return v; }
            while (ok) {
                ws();
                if (p >= end || *p != '"') { ok = false; break; }
                std::string k = string();
                ws();
                if (p >= end || *p != ':') { ok = false; break; }
                ++p;
                v.obj.emplace_back(std::move(k), value());
                ws();
                if (p < end && *p == ',') { ++p; continue; }
                if (p < end && *p == '}') { ++p; break; }
                ok = false;
            }
        } else if (*p == '[') {
            v.type = JVal::Arr; ++p; ws();
            if (p < end && *p == ']') { ++p; --depth; return v; }
            while (ok) {
                v.arr.push_back(value());
                ws();
                if (p < end && *p == ',') { ++p; continue; }
                if (p < end && *p == ']') { ++p; break; }
                ok = false;
            }
        } else if (*p == '"') {
            v.type = JVal::Str; v.str = string();
        } else if (lit("true")) { v.type = JVal::Bool; v.b = true; }
        else if (lit("false")) { v.type = JVal::Bool; }
        else if (lit("null")) { }
        else {
            char* e = nullptr;
            std::string tmp(p, (size_t)std::min<ptrdiff_t>(end - p, 64));
            v.num = strtod(tmp.c_str(), &e);
            if (e == tmp.c_str()) ok = false;
            else { v.type = JVal::Num; p += e - tmp.c_str(); }
        }
        --depth;
        return v;
    }
};

bool read_file(const std::string& path, std::vector<uint8_t>* out) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    out->resize(n > 0 ? (size_t)n : 0);
    bool ok = n >= 0 && fread(out->data(), 1, out->size(), f) == out->size();
    fclose(f);
    return ok;
}

std::vector<uint8_t> base64(const char* s, size_t n) {
    std::vector<uint8_t> out;
    unsigned acc = 0; int bits = 0;
    for (size_t i = 0; i < n; ++i) {
        const char c = s[i];
        int v;
        if (c >= 'A' && c <= 'Z') v = c - 'A';
        else if (c >= 'a' && c <= 'z') v = c - 'a' + 26;
        else if (c >= '0' && c <= '9') v = c - '0' + 52;
        else if (c == '+' || c == '-') v = 62;
        else if (c == '/' || c == '_') v = 63;
        else continue;      // padding, line breaks
        acc = (acc << 6) | (unsigned)v; bits += 6;
        if (bits >= 8) { bits -= 8; out.push_back((uint8_t)(acc >> bits)); }
    }
    return out;
}

std::string dir_of(const std::string& path) {
    size_t k = path.find_last_of("/\\");
    return k == std::string::npos ? std::string(".") : path.substr(0, k);
}

std::string uri_decode(const std::string& s) {   // %20 etc. in side-car file names
    std::string o;
    for (size_t i = 0; i < s.size(); ++i) {
        if (s[i] == '%' && i + 2 < s.size()) {
            auto hex = [](char c) { return c >= '0' && c <= '9' ? c - '0' : (c | 32) - 'a' + 10; };
            o += (char)(hex(s[i + 1]) * 16 + hex(s[i + 2]));
            i += 2;
        } else {
            o += s[i];
        }
    }
    return o;
}

// ------------------------------------------------------------------------------------------------ PNG
bool inflate_all(const uint8_t* src, size_t n, std::vector<uint8_t>* out, size_t expect) {
    out->resize(expect);
    z_stream z;
    memset(&z, 0, sizeof(z));
    if (inflateInit(&z) != Z_OK) return false;
    z.next_in = const_cast<Bytef*>(src); z.avail_in = (uInt)n;
    z.next_out = out->data(); z.avail_out = (uInt)expect;
    int rc = inflate(&z, Z_FINISH);
    const bool ok = (rc == Z_STREAM_END || rc == Z_OK || rc == Z_BUF_ERROR) && z.total_out == expect;
    inflateEnd(&z);
    return ok;
}

uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

bool decode_png(const uint8_t* d, size_t n, GltfImage* out) {
    static const uint8_t SIG[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    if (n < 8 || memcmp(d, SIG, 8) != 0) return false;
    size_t off = 8;
    uint32_t w = 0, h = 0; int depth = 0, ctype = 0, interlace = 0;
    std::vector<uint8_t> idat, plte, trns;
    while (off + 12 <= n) {
        const uint32_t len = be32(d + off);
        const uint8_t* type = d + off + 4;
        const uint8_t* body = d + off + 8;
        if (off + 12 + (size_t)len > n) return false;
        if (!memcmp(type, "IHDR", 4) && len >= 13) { w = be32(body); h = be32(body + 4); depth = body[8]; ctype = body[9]; interlace = body[12]; }
        else if (!memcmp(type, "PLTE", 4)) plte.assign(body, body + len);
        else if (!memcmp(type, "tRNS", 4)) trns.assign(body, body + len);
        else if (!memcmp(type, "IDAT", 4)) idat.insert(idat.end(), body, body + len);
        else if (!memcmp(type, "IEND", 4)) break;
        off += 12 + (size_t)len;
    }
    if (!w || !h || interlace != 0 || (uint64_t)w * h > (1ull << 28)) return false;
    int channels;
    switch (ctype) { case 0: channels = 1; break; case 2: channels = 3; break; case 3: channels = 1; break; case 4: channels = 2; break; case 6: channels = 4; break; default: return false; }
    if (!(depth == 8 || depth == 16 || ((ctype == 0 || ctype == 3) && (depth == 1 || depth == 2 || depth == 4)))) return false;
    if (ctype == 3 && depth == 16) return false;
    const size_t bits_pp = (size_t)channels * depth, bpp = bits_pp >= 8 ? bits_pp / 8 : 1;
    const size_t row = ((size_t)w * bits_pp + 7) / 8;
    std::vector<uint8_t> raw;
    if (!inflate_all(idat.data(), idat.size(), &raw, (row + 1) * h)) return false;
    std::vector<uint8_t> img(row * h);
    for (uint32_t y = 0; y < h; ++y) {      // the five scanline filters (PNG 1.2, section 6)
        const uint8_t f = raw[(row + 1) * y];
        const uint8_t* s = raw.data() + (row + 1) * y + 1;
        uint8_t* o = img.data() + row * y;
        const uint8_t* up = y ? img.data() + row * (y - 1) : nullptr;
        for (size_t x = 0; x < row; ++x) {
            const int a = x >= bpp ? o[x - bpp] : 0, b = up ? up[x] : 0, c = (up && x >= bpp) ? up[x - bpp] : 0;
            int v;
            switch (f) {
                case 0: v = s[x]; break;
                case 1: v = s[x] + a; break;
                case 2: v = s[x] + b; break;
                case 3: v = s[x] + ((a + b) >> 1); break;
                case 4: { const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
                          v = s[x] + ((pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c)); break; }
                default: return false;
            }
            o[x] = (uint8_t)v;
        }
    }
    out->width = w; out->height = h;
    out->rgba.resize((size_t)w * h * 4);
    auto sample = [&](const uint8_t* line, size_t x, int ch) -> unsigned {      // channel `ch` of pixel x, at the file's bit depth
        if (depth == 8) return line[x * channels + ch];
        if (depth == 16) return ((unsigned)line[(x * channels + ch) * 2] << 8) | line[(x * channels + ch) * 2 + 1];
        const size_t bit = x * depth;
        return (line[bit >> 3] >> (8 - depth - (bit & 7))) & ((1u << depth) - 1u);
    };
    const unsigned maxv = (1u << depth) - 1u;
    auto to8 = [&](unsigned v) -> uint8_t {      // PIL's conversions: 16-bit -> high byte; sub-byte grey -> scaled to 0..255
        if (depth == 8) return (uint8_t)v;
        if (depth == 16) return (uint8_t)(v >> 8);
        return (uint8_t)(v * 255u / maxv);
    };
    for (uint32_t y = 0; y < h; ++y) {
        const uint8_t* line = img.data() + row * y;
        uint8_t* o = out->rgba.data() + (size_t)w * 4 * y;
        for (uint32_t x = 0; x < w; ++x, o += 4) {
            switch (ctype) {
                case 0: {
                    const unsigned g = sample(line, x, 0);
                    o[0] = o[1] = o[2] = to8(g);
                    o[3] = (trns.size() >= 2 && g == (((unsigned)trns[0] << 8) | trns[1])) ? 0 : 255;
                    break;
                }
                case 2:
                    o[0] = to8(sample(line, x, 0)); o[1] = to8(sample(line, x, 1)); o[2] = to8(sample(line, x, 2));
                    o[3] = 255;
                    if (trns.size() >= 6 && sample(line, x, 0) == (((unsigned)trns[0] << 8) | trns[1]) &&
                        sample(line, x, 1) == (((unsigned)trns[2] << 8) | trns[3]) && sample(line, x, 2) == (((unsigned)trns[4] << 8) | trns[5])) o[3] = 0;
                    break;
                case 3: {
                    const unsigned i = sample(line, x, 0);
                    if (3 * (size_t)i + 2 >= plte.size()) return false;
                    o[0] = plte[3 * i]; o[1] = plte[3 * i + 1]; o[2] = plte[3 * i + 2];
                    o[3] = i < trns.size() ? trns[i] : 255;
                    break;
                }
                case 4: o[0] = o[1] = o[2] = to8(sample(line, x, 0)); o[3] = to8(sample(line, x, 1)); break;
                default: o[0] = to8(sample(line, x, 0)); o[1] = to8(sample(line, x, 1)); o[2] = to8(sample(line, x, 2)); o[3] = to8(sample(line, x, 3));
            }
        }
    }
    return true;
}

// ------------------------------------------------------------------------------------------------ glTF
struct Doc {
    JVal js;
    std::vector<std::vector<uint8_t>> buffers;
    std::string base_dir;
};

bool load_document(const std::string& path, Doc* doc, std::string* err) {
    std::vector<uint8_t> file;
    if (!read_file(path, &file)) { *err = "cannot read " + path; return false; }
    doc->base_dir = dir_of(path);
    const char* json = nullptr; size_t json_len = 0;
    std::vector<uint8_t> bin;
    bool have_bin = false;
    if (file.size() >= 12 && !memcmp(file.data(), "glTF", 4)) {       // binary container: 12-byte header, then chunks
        size_t off = 12;
        while (off + 8 <= file.size()) {
            uint32_t clen, ctype;
            memcpy(&clen, file.data() + off, 4); memcpy(&ctype, file.data() + off + 4, 4);
            if (off + 8 + (size_t)clen > file.size()) { *err = "truncated GLB chunk"; return false; }
            if (ctype == 0x4E4F534Au) { json = (const char*)file.data() + off + 8; json_len = clen; }
            else if (ctype == 0x004E4942u && !have_bin) { bin.assign(file.begin() + (long)off + 8, file.begin() + (long)(off + 8 + clen)); have_bin = true; }
            off += 8 + (size_t)clen;
        }
    } else {
        json = (const char*)file.data(); json_len = file.size();
    }
    if (!json) { *err = "no JSON chunk"; return false; }
    JParser p{json, json + json_len};
    doc->js = p.value();
    if (!p.ok || doc->js.type != JVal::Obj) { *err = "malformed glTF JSON"; return false; }
    const JVal* bufs = doc->js.get("buffers");
    for (size_t i = 0; bufs && i < bufs->size(); ++i) {
        const JVal* uri = bufs->arr[i].get("uri");
        if (!uri || uri->type != JVal::Str) {                          // the GLB's own BIN chunk
            if (!have_bin) { *err = "buffer without uri and no BIN chunk"; return false; }
            doc->buffers.push_back(bin);
        } else if (uri->str.compare(0, 5, "data:") == 0) {
            size_t comma = uri->str.find(',');
            if (comma == std::string::npos) { *err = "bad data: URI"; return false; }
            doc->buffers.push_back(base64(uri->str.c_str() + comma + 1, uri->str.size() - comma - 1));
        } else {
            std::vector<uint8_t> b;
            if (!read_file(doc->base_dir + "/" + uri_decode(uri->str), &b)) { *err = "cannot read buffer " + uri->str; return false; }
            doc->buffers.push_back(std::move(b));
        }
    }
    return true;
}

int component_size(long t) { return t == 5120 || t == 5121 ? 1 : (t == 5122 || t == 5123 ? 2 : (t == 5125 || t == 5126 ? 4 : 0)); }
int type_components(const std::string& t) { return t == "SCALAR" ? 1 : t == "VEC2" ? 2 : t == "VEC3" ? 3 : t == "VEC4" ? 4 : t == "MAT4" ? 16 : 0; }

// accessor `idx` as floats (integers converted, normalised ones scaled) — `ncomp` components per element
bool read_accessor(const Doc& d, long idx, int ncomp, std::vector<float>* f_out, std::vector<uint32_t>* u_out, std::string* err) {
    const JVal* accs = d.js.get("accessors");
    if (!accs || idx < 0 || (size_t)idx >= accs->size()) { *err = "accessor index out of range"; return false; }
    const JVal& a = accs->arr[(size_t)idx];
    const long ct = a.index("componentType"), count = a.index("count"), view = a.index("bufferView");
    const JVal* type = a.get("type");
    const int cs = component_size(ct), nc = type ? type_components(type->str) : 0;
    if (!cs || nc != ncomp || count < 0) { *err = "unsupported accessor layout"; return false; }
    if (a.get("sparse")) { *err = "sparse accessors are not supported"; return false; }
    const JVal* views = d.js.get("bufferViews");
    if (view < 0) {                                                     // no view: zeros (glTF 2.0, 5.1.1)
        if (f_out) f_out->assign((size_t)count * ncomp, 0.0f);
        if (u_out) u_out->assign((size_t)count * ncomp, 0u);
        return true;
    }
    if (!views || (size_t)view >= views->size()) { *err = "bufferView index out of range"; return false; }
    const JVal& bv = views->arr[(size_t)view];
    const long buffer = bv.index("buffer");
    if (buffer < 0 || (size_t)buffer >= d.buffers.size()) { *err = "buffer index out of range"; return false; }
    const std::vector<uint8_t>& buf = d.buffers[(size_t)buffer];
    const size_t start = (size_t)bv.number("byteOffset", 0) + (size_t)a.number("byteOffset", 0);
    size_t stride = (size_t)bv.number("byteStride", 0);
    if (!stride) stride = (size_t)cs * ncomp;
    if (count && start + stride * ((size_t)count - 1) + (size_t)cs * ncomp > buf.size()) { *err = "accessor reaches beyond its buffer"; return false; }
    const JVal* nrm = a.get("normalized");
    const bool normalized = nrm && nrm->type == JVal::Bool && nrm->b;
    if (f_out) f_out->resize((size_t)count * ncomp);
    if (u_out) u_out->resize((size_t)count * ncomp);
    for (size_t i = 0; i < (size_t)count; ++i) {
        const uint8_t* p = buf.data() + start + stride * i;
        for (int c = 0; c < ncomp; ++c, p += cs) {
            double v; uint32_t u = 0;
            switch (ct) {
                case 5120: { int8_t t; memcpy(&t, p, 1); v = normalized ? std::max(t / 127.0, -1.0) : t; u = (uint32_t)t; break; }
                case 5121: { uint8_t t; memcpy(&t, p, 1); v = normalized ? t / 255.0 : t; u = t; break; }
                case 5122: { int16_t t; memcpy(&t, p, 2); v = normalized ? std::max(t / 32767.0, -1.0) : t; u = (uint32_t)t; break; }
                case 5123: { uint16_t t; memcpy(&t, p, 2); v = normalized ? t / 65535.0 : t; u = t; break; }
                case 5125: { uint32_t t; memcpy(&t, p, 4); v = t; u = t; break; }
                default: { float t; memcpy(&t, p, 4); v = t; u = (uint32_t)t; if (f_out) { (*f_out)[i * ncomp + c] = t; } }
            }
            if (f_out && ct != 5126) (*f_out)[i * ncomp + c] = (float)v;
            if (u_out) (*u_out)[i * ncomp + c] = u;
        }
    }
    return true;
}

// glam f32: Mat4::from_scale_rotation_translation / Mat4 * Mat4, as tools/make_assets.py restates them
void node_local_matrix(const JVal& node, float m[16]) {
    const JVal* mat = node.get("matrix");
    if (mat && mat->size() == 16) { for (int i = 0; i < 16; ++i) m[i] = (float)mat->arr[i].num; return; }
    float t[3] = {0, 0, 0}, s[3] = {1, 1, 1}, q[4] = {0, 0, 0, 1};
    const JVal* v;
    if ((v = node.get("translation")) && v->size() == 3) for (int i = 0; i < 3; ++i) t[i] = (float)v->arr[i].num;
    if ((v = node.get("scale")) && v->size() == 3) for (int i = 0; i < 3; ++i) s[i] = (float)v->arr[i].num;
    if ((v = node.get("rotation")) && v->size() == 4) for (int i = 0; i < 4; ++i) q[i] = (float)v->arr[i].num;
    const float x = q[0], y = q[1], z = q[2], w = q[3];
    const float x2 = x + x, y2 = y + y, z2 = z + z;
    const float xx = x * x2, xy = x * y2, xz = x * z2, yy = y * y2, yz = y * z2, zz = z * z2, wx = w * x2, wy = w * y2, wz = w * z2;
    const float cols[3][3] = {{1.0f - (yy + zz), xy + wz, xz - wy}, {xy - wz, 1.0f - (xx + zz), yz + wx}, {xz + wy, yz - wx, 1.0f - (xx + yy)}};
    for (int c = 0; c < 3; ++c) {
        for (int r = 0; r < 3; ++r) m[4 * c + r] = cols[c][r] * s[c];
        m[4 * c + 3] = 0.0f;
    }
    m[12] = t[0]; m[13] = t[1]; m[14] = t[2]; m[15] = 1.0f;
}
void mat_mul(const float a[16], const float b[16], float out[16]) {   // a * b, column-major, accumulated left to right
    for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 4; ++r) {
            float acc = 0.0f;
            for (int k = 0; k < 4; ++k) acc = acc + a[4 * k + r] * b[4 * c + k];
            out[4 * c + r] = acc;
        }
}

struct Loader {
    MeshMaterialWorld& world;
    const Doc& doc;
    GltfImageDecoder decoder; void* user;
    GltfLoadResult& res;
    std::map<std::pair<long, long>, uint32_t> mesh_of;        // (glTF mesh, primitive) -> world mesh id
    std::vector<long> material_of_mesh;                       // world mesh id - first id -> glTF material index or -1
    std::vector<uint32_t> material_ids;                       // glTF material -> world material id
    long default_material = -1;
    std::map<std::pair<long, int>, uint32_t> texture_of;      // (image, srgb) -> world texture id
    std::string err;

    const JVal* array(const char* key) const { const JVal* v = doc.js.get(key); return v && v->type == JVal::Arr ? v : nullptr; }

    bool image_bytes(long image, std::vector<uint8_t>* bytes, std::string* mime) {
        const JVal* images = array("images");
        if (!images || image < 0 || (size_t)image >= images->size()) { err = "image index out of range"; return false; }
        const JVal& im = images->arr[(size_t)image];
        const JVal* mt = im.get("mimeType");
        if (mt && mt->type == JVal::Str) *mime = mt->str;
        const long view = im.index("bufferView");
        if (view >= 0) {
            const JVal* views = array("bufferViews");
            if (!views || (size_t)view >= views->size()) { err = "image bufferView out of range"; return false; }
            const JVal& bv = views->arr[(size_t)view];
            const long buffer = bv.index("buffer");
            const size_t off = (size_t)bv.number("byteOffset", 0), len = (size_t)bv.number("byteLength", 0);
            if (buffer < 0 || (size_t)buffer >= doc.buffers.size() || off + len > doc.buffers[(size_t)buffer].size()) { err = "image bytes out of range"; return false; }
            bytes->assign(doc.buffers[(size_t)buffer].begin() + (long)off, doc.buffers[(size_t)buffer].begin() + (long)(off + len));
            return true;
        }
        const JVal* uri = im.get("uri");
        if (!uri || uri->type != JVal::Str) { err = "image without bufferView or uri"; return false; }
        if (uri->str.compare(0, 5, "data:") == 0) {
            size_t comma = uri->str.find(',');
            if (comma == std::string::npos) { err = "bad data: URI"; return false; }
            if (mime->empty()) *mime = uri->str.substr(5, uri->str.find(';') == std::string::npos ? comma - 5 : uri->str.find(';') - 5);
            *bytes = base64(uri->str.c_str() + comma + 1, uri->str.size() - comma - 1);
            return true;
        }
        if (!read_file(doc.base_dir + "/" + uri_decode(uri->str), bytes)) { err = "cannot read image " + uri->str; return false; }
        return true;
    }

    // texture slot `info` ({index, texCoord}) -> world texture id; one texture per (image, colour space), in first-use order
    bool texture_id(const JVal* info, bool srgb, uint32_t* out) {
        *out = 0xFFFFFFFFu;
        if (!info || info->type != JVal::Obj) return true;
        const JVal* textures = array("textures");
        const long ti = info->index("index");
        if (!textures || ti < 0 || (size_t)ti >= textures->size()) { err = "texture index out of range"; return false; }
        const JVal& tex = textures->arr[(size_t)ti];
        const long source = tex.index("source"), sampler = tex.index("sampler");
        auto key = std::make_pair(source, srgb ? 1 : 0);
        auto it = texture_of.find(key);
        if (it != texture_of.end()) { *out = it->second; return true; }
        std::vector<uint8_t> bytes; std::string mime;
        if (!image_bytes(source, &bytes, &mime)) return false;
        GltfImage img;
        if (!decode_png(bytes.data(), bytes.size(), &img)) {
            if (!decoder || !decoder(bytes.data(), bytes.size(), mime.c_str(), user, &img) || img.rgba.size() != (size_t)img.width * img.height * 4 || !img.width) {
                err = "image " + std::to_string(source) + " (" + (mime.empty() ? "unknown type" : mime) + ") is not a PNG this loader decodes and no decoder took it";
                return false;
            }
        }
        hk_texture_desc td;
        memset(&td, 0, sizeof(td));
        td.width = img.width; td.height = img.height;
        const JVal* samplers = array("samplers");
        long wrap_s = 10497, wrap_t = 10497, mag = 9729;
        if (samplers && sampler >= 0 && (size_t)sampler < samplers->size()) {
            const JVal& s = samplers->arr[(size_t)sampler];
            wrap_s = (long)s.number("wrapS", 10497); wrap_t = (long)s.number("wrapT", 10497); mag = (long)s.number("magFilter", 9729);
        }
        auto mode = [](long w) { return w == 33071 ? 1u : (w == 33648 ? 2u : 0u); };     // CLAMP_TO_EDGE, MIRRORED_REPEAT, else REPEAT
        td.address_mode_u = mode(wrap_s); td.address_mode_v = mode(wrap_t);
        td.filter_linear = mag == 9728 ? 0u : 1u;
        td.srgb = srgb ? 1u : 0u;
        td.rgba8 = img.rgba.data();
        *out = world.add_texture(td, img.rgba.data());
        texture_of[key] = *out;
        res.textures.push_back(*out);
        return true;
    }

    bool materials() {
        const JVal* mats = array("materials");
        for (size_t i = 0; mats && i < mats->size(); ++i) {
            const JVal& m = mats->arr[i];
            StandardMaterial s;
            s.perceptual_roughness = 1.0f; s.metallic = 1.0f;      // glTF defaults of the factors (bevy_gltf load_material)
            const JVal* pbr = m.get("pbrMetallicRoughness");
            const JVal* v;
            if (pbr && (v = pbr->get("baseColorFactor")) && v->size() == 4) for (int k = 0; k < 4; ++k) s.base_color[(size_t)k] = (float)v->arr[(size_t)k].num;
            if ((v = m.get("emissiveFactor")) && v->size() == 3) for (int k = 0; k < 3; ++k) s.emissive[(size_t)k] = (float)v->arr[(size_t)k].num;
            s.emissive[3] = 1.0f;
            if (pbr) { s.perceptual_roughness = (float)pbr->number("roughnessFactor", 1.0); s.metallic = (float)pbr->number("metallicFactor", 1.0); }
            if (!texture_id(pbr ? pbr->get("baseColorTexture") : nullptr, true, &s.base_color_texture)) return false;
            if (!texture_id(m.get("emissiveTexture"), true, &s.emissive_texture)) return false;
            if (!texture_id(pbr ? pbr->get("metallicRoughnessTexture") : nullptr, false, &s.metallic_roughness_texture)) return false;
            if (!texture_id(m.get("normalTexture"), false, &s.normal_map_texture)) return false;
            if (!texture_id(m.get("occlusionTexture"), false, &s.occlusion_texture)) return false;
            material_ids.push_back(world.add_material(s));
            res.materials.push_back(material_ids.back());
        }
        return true;
    }

    bool meshes() {
        const JVal* ms = array("meshes");
        for (size_t mi = 0; ms && mi < ms->size(); ++mi) {
            const JVal* prims = ms->arr[mi].get("primitives");
            for (size_t pi = 0; prims && pi < prims->size(); ++pi) {
                const JVal& p = prims->arr[pi];
                Mesh mesh;
                const long mode = (long)p.number("mode", 4);
                mesh.topology = mode == 4 ? PrimitiveTopology::TriangleList : (mode == 5 ? PrimitiveTopology::TriangleStrip : PrimitiveTopology::Other);
                const JVal* at = p.get("attributes");
                std::vector<float> f;
                const long pos = at ? at->index("POSITION") : -1, nrm = at ? at->index("NORMAL") : -1, uv = at ? at->index("TEXCOORD_0") : -1;
                // a missing attribute leaves the array empty: GpuMesh::try_from reports it and the instance is dropped (mod.rs:301-308)
                if (pos >= 0) {
                    if (!read_accessor(doc, pos, 3, &f, nullptr, &err)) return false;
                    mesh.positions.resize(f.size() / 3);
                    memcpy(mesh.positions.data(), f.data(), f.size() * 4);
                }
                if (nrm >= 0) {
                    if (!read_accessor(doc, nrm, 3, &f, nullptr, &err)) return false;
                    mesh.normals.resize(f.size() / 3);
                    memcpy(mesh.normals.data(), f.data(), f.size() * 4);
                }
                if (uv >= 0) {
                    if (!read_accessor(doc, uv, 2, &f, nullptr, &err)) return false;
                    mesh.uvs.resize(f.size() / 2);
                    memcpy(mesh.uvs.data(), f.data(), f.size() * 4);
                }
                const long ind = p.index("indices");
                mesh.has_indices = ind >= 0;
                if (ind >= 0 && !read_accessor(doc, ind, 1, nullptr, &mesh.indices, &err)) return false;
                for (uint32_t ix : mesh.indices)
                    if (ix >= mesh.positions.size()) { err = "primitive index beyond its vertices"; return false; }
                if ((!mesh.normals.empty() && mesh.normals.size() != mesh.positions.size()) || (!mesh.uvs.empty() && mesh.uvs.size() != mesh.positions.size())) {
                    err = "vertex attributes of different lengths"; return false;
                }
                const uint32_t id = world.add_mesh(mesh);
                mesh_of[{(long)mi, (long)pi}] = id;
                res.meshes.push_back(id);
                material_of_mesh.push_back(p.index("material"));
            }
        }
        return true;
    }

    uint32_t material_for(size_t mesh_slot) {
        const long m = material_of_mesh[mesh_slot];
        if (m >= 0 && (size_t)m < material_ids.size()) return material_ids[(size_t)m];
        if (default_material < 0) {                              // a primitive without material: StandardMaterial::default()
            default_material = (long)world.add_material(StandardMaterial());
            res.materials.push_back((uint32_t)default_material);
        }
        return (uint32_t)default_material;
    }

    bool walk(long ni, const float parent[16], int depth) {
        const JVal* nodes = array("nodes");
        if (!nodes || ni < 0 || (size_t)ni >= nodes->size() || depth > 512) { err = "node index out of range (or a cycle)"; return false; }
        const JVal& node = nodes->arr[(size_t)ni];
        float local[16], worldm[16];
        node_local_matrix(node, local);
        mat_mul(parent, local, worldm);
        const long mesh = node.index("mesh");
        if (mesh >= 0) {
            const JVal* ms = array("meshes");
            const JVal* prims = ms && (size_t)mesh < ms->size() ? ms->arr[(size_t)mesh].get("primitives") : nullptr;
            for (size_t pi = 0; prims && pi < prims->size(); ++pi) {
                auto it = mesh_of.find({mesh, (long)pi});
                if (it == mesh_of.end()) continue;
                InstanceDesc d;
                d.mesh = it->second;
                d.material = material_for((size_t)(it->second - res.meshes.front()));
                memcpy(d.transform, worldm, 64);
                res.instances.push_back(world.add_instance(d));
            }
        }
        const JVal* kids = node.get("children");
        for (size_t c = 0; kids && c < kids->size(); ++c)
            if (!walk((long)kids->arr[c].num, worldm, depth + 1)) return false;
        return true;
    }
};

}  // namespace

bool decode_png_rgba8(const uint8_t* bytes, size_t n, GltfImage* out) { return decode_png(bytes, n, out); }

bool load_gltf(MeshMaterialWorld& world, const char* path, const float* parent_transform, GltfImageDecoder decoder, void* user, GltfLoadResult* out) {
    GltfLoadResult local;
    GltfLoadResult& res = out ? *out : local;
    res = GltfLoadResult();
    Doc doc;
    if (!path || !load_document(path, &doc, &res.error)) { if (res.error.empty()) res.error = "no path"; return false; }
    Loader L{world, doc, decoder, user, res};
    if (!L.materials() || !L.meshes()) { res.error = L.err; return false; }
    static const float IDENTITY[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    const float* parent = parent_transform ? parent_transform : IDENTITY;
    const JVal* scenes = doc.js.get("scenes");
    const long scene = doc.js.index("scene") >= 0 ? doc.js.index("scene") : 0;
    if (scenes && (size_t)scene < scenes->size()) {
        const JVal* roots = scenes->arr[(size_t)scene].get("nodes");
        for (size_t i = 0; roots && i < roots->size(); ++i)
            if (!L.walk((long)roots->arr[i].num, parent, 0)) { res.error = L.err; return false; }
    }
    return true;
}

}  // namespace hikari

// ------------------------------------------------------------------------------------------------ C shims (include/hikari_host.h)
extern "C" {

int hikari_world_load_gltf(hikari_world* w, const char* path, const float* parent_transform16, hikari_image_decoder decoder, void* user,
                           hikari_gltf_counts* counts, char* error, size_t error_capacity) {
    struct Thunk { hikari_image_decoder fn; void* user; } thunk{decoder, user};
    auto bridge = [](const uint8_t* bytes, size_t n, const char* mime, void* u, hikari::GltfImage* out) -> bool {
        Thunk* t = static_cast<Thunk*>(u);
        uint32_t width = 0, height = 0;
        if (!t->fn(bytes, n, mime, t->user, nullptr, &width, &height) || !width || !height) return false;     // first call: the size
        out->width = width; out->height = height;
        out->rgba.resize((size_t)width * height * 4);
        return t->fn(bytes, n, mime, t->user, out->rgba.data(), &width, &height) != 0;                            // second call: the pixels
    };
    hikari::GltfLoadResult res;
    const bool ok = hikari::load_gltf(*reinterpret_cast<hikari::MeshMaterialWorld*>(w), path, parent_transform16,
                                      decoder ? +bridge : nullptr, &thunk, &res);
    if (counts) {
        counts->first_mesh = res.meshes.empty() ? 0u : res.meshes.front(); counts->mesh_count = (uint32_t)res.meshes.size();
        counts->first_material = res.materials.empty() ? 0u : res.materials.front(); counts->material_count = (uint32_t)res.materials.size();
        counts->first_instance = res.instances.empty() ? 0u : res.instances.front(); counts->instance_count = (uint32_t)res.instances.size();
        counts->first_texture = res.textures.empty() ? 0u : res.textures.front(); counts->texture_count = (uint32_t)res.textures.size();
    }
    if (error && error_capacity) { strncpy(error, res.error.c_str(), error_capacity - 1); error[error_capacity - 1] = 0; }
    return ok ? 1 : 0;
}

int hikari_decode_png(const uint8_t* bytes, size_t n, uint8_t* rgba_out, uint32_t* width, uint32_t* height) {
    hikari::GltfImage img;
    if (!hikari::decode_png_rgba8(bytes, n, &img)) return 0;
    if (width) *width = img.width;
    if (height) *height = img.height;
    if (rgba_out) memcpy(rgba_out, img.rgba.data(), img.rgba.size());
    return 1;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------ Mesh::from(shape::*)
// The bevy 0.9 shape generators the reference's examples spawn next to their glTF scenes (examples/scene.rs:86-113: Plane + UVSphere,
// examples/city.rs:64-90: Plane + UVSphere, examples/minimal.rs: Plane + Cube), restated like bevy_hikari_b200/scenes.py restates
// them: vertex order, winding, normals and UVs are bevy_render::mesh::shape's; angles are evaluated in double and rounded once.
namespace hikari {
namespace shape {

Mesh plane(float size) {
    const float e = size / 2.0f;
    Mesh m;
    m.positions = {{e, 0, -e}, {e, 0, e}, {-e, 0, e}, {-e, 0, -e}};
    m.normals.assign(4, {0, 1, 0});
    m.uvs = {{1, 0}, {1, 1}, {0, 1}, {0, 0}};
    m.indices = {0, 2, 1, 0, 3, 2};
    return m;
}

Mesh uv_sphere(float radius_f, uint32_t sectors, uint32_t stacks) {
    const double PI = 3.141592653589793, radius = radius_f;
    Mesh m;
    for (uint32_t i = 0; i <= stacks; ++i) {
        const double stack_angle = PI / 2 - i * PI / stacks;
        const double xy = radius * cos(stack_angle), z = radius * sin(stack_angle);
        for (uint32_t j = 0; j <= sectors; ++j) {
            const double a = j * 2 * PI / sectors;
            const double x = xy * cos(a), y = xy * sin(a);
            m.positions.push_back({(float)x, (float)y, (float)z});
            m.normals.push_back({(float)(x / radius), (float)(y / radius), (float)(z / radius)});
            m.uvs.push_back({(float)((double)j / sectors), (float)((double)i / stacks)});
        }
    }
    for (uint32_t i = 0; i < stacks; ++i) {
        uint32_t k1 = i * (sectors + 1), k2 = (i + 1) * (sectors + 1);
        for (uint32_t j = 0; j < sectors; ++j, ++k1, ++k2) {
            if (i != 0) { m.indices.push_back(k1); m.indices.push_back(k2); m.indices.push_back(k1 + 1); }
            if (i != stacks - 1) { m.indices.push_back(k1 + 1); m.indices.push_back(k2); m.indices.push_back(k2 + 1); }
        }
    }
    return m;
}

Mesh box(float sx, float sy, float sz) {     // shape::Box::new; shape::Cube { size } = box(size, size, size)
    const float x0 = -sx / 2, x1 = sx / 2, y0 = -sy / 2, y1 = sy / 2, z0 = -sz / 2, z1 = sz / 2;
    struct V { float p[3], n[3], t[2]; };
    const V v[24] = {
        {{x0, y0, z1}, {0, 0, 1}, {0, 0}}, {{x1, y0, z1}, {0, 0, 1}, {1, 0}}, {{x1, y1, z1}, {0, 0, 1}, {1, 1}}, {{x0, y1, z1}, {0, 0, 1}, {0, 1}},
        {{x0, y1, z0}, {0, 0, -1}, {1, 0}}, {{x1, y1, z0}, {0, 0, -1}, {0, 0}}, {{x1, y0, z0}, {0, 0, -1}, {0, 1}}, {{x0, y0, z0}, {0, 0, -1}, {1, 1}},
        {{x1, y0, z0}, {1, 0, 0}, {0, 0}}, {{x1, y1, z0}, {1, 0, 0}, {1, 0}}, {{x1, y1, z1}, {1, 0, 0}, {1, 1}}, {{x1, y0, z1}, {1, 0, 0}, {0, 1}},
        {{x0, y0, z1}, {-1, 0, 0}, {1, 0}}, {{x0, y1, z1}, {-1, 0, 0}, {0, 0}}, {{x0, y1, z0}, {-1, 0, 0}, {0, 1}}, {{x0, y0, z0}, {-1, 0, 0}, {1, 1}},
        {{x1, y1, z0}, {0, 1, 0}, {1, 0}}, {{x0, y1, z0}, {0, 1, 0}, {0, 0}}, {{x0, y1, z1}, {0, 1, 0}, {0, 1}}, {{x1, y1, z1}, {0, 1, 0}, {1, 1}},
        {{x1, y0, z1}, {0, -1, 0}, {0, 0}}, {{x0, y0, z1}, {0, -1, 0}, {1, 0}}, {{x0, y0, z0}, {0, -1, 0}, {1, 1}}, {{x1, y0, z0}, {0, -1, 0}, {0, 1}},
    };
    Mesh m;
    for (const V& k : v) {
        m.positions.push_back({k.p[0], k.p[1], k.p[2]}); m.normals.push_back({k.n[0], k.n[1], k.n[2]}); m.uvs.push_back({k.t[0], k.t[1]});
    }
    for (uint32_t k = 0; k < 24; k += 4)
        for (uint32_t o : {0u, 1u, 2u, 2u, 3u, 0u}) m.indices.push_back(k + o);
    return m;
}

}  // namespace shape
}  // namespace hikari

extern "C" uint32_t hikari_world_add_shape(hikari_world* w, uint32_t kind, const float* params) {
    hikari::MeshMaterialWorld* world = reinterpret_cast<hikari::MeshMaterialWorld*>(w);
    switch (kind) {
        case HIKARI_SHAPE_PLANE: return world->add_mesh(hikari::shape::plane(params[0]));
        case HIKARI_SHAPE_UV_SPHERE: return world->add_mesh(hikari::shape::uv_sphere(params[0], (uint32_t)params[1], (uint32_t)params[2]));
        case HIKARI_SHAPE_BOX: return world->add_mesh(hikari::shape::box(params[0], params[1], params[2]));
        default: return 0xFFFFFFFFu;
    }
}
