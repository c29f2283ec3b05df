// File boundary (include/prb_io.h): oriented-point readers (.ply / .bnpts / ASCII) and the
// triangle-mesh PLY writer.  Host-only; replaces PointStream.inl + the Greg Turk PLY library
// (plyfile.cu, 2907 lines) for the two things the reference's main() does with them.
//
// Readers load the whole file with one read() and parse in place: the binary paths are a
// strided gather (multi-threaded), the ASCII paths a hand-rolled strtof loop.  The ASCII writer
// formats chunks of elements in parallel threads into per-chunk buffers and writes them in
// order, producing the same bytes as the reference's per-item fprintf("%g ").
#include "prb_io.h"
#include <algorithm>
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace prb {
void set_error(const std::string& msg);
}

namespace {

enum { kErrArg = -1, kErrIo = -5, kErrFormat = -6 };

int fail(int code, const std::string& msg) {
    prb::set_error(msg);
    return code;
}

bool read_file(const char* path, std::vector<char>& buf) {
    FILE* fp = std::fopen(path, "rb");
    if (!fp) return false;
    if (std::fseek(fp, 0, SEEK_END) != 0) { std::fclose(fp); return false; }
    long long sz = std::ftell(fp);
    if (sz < 0 || std::fseek(fp, 0, SEEK_SET) != 0) { std::fclose(fp); return false; }     // not seekable (pipe, directory)
    buf.resize((size_t)sz + 1);
    size_t got = sz ? std::fread(buf.data(), 1, (size_t)sz, fp) : 0;
    std::fclose(fp);
    buf[got] = 0;
    buf.resize(got + 1);   // NUL-terminated for the ASCII parsers
    return true;
}

int n_threads(size_t items, size_t perThread) {
    unsigned hw = std::thread::hardware_concurrency();
    if (hw == 0) hw = 1;
    size_t want = (items + perThread - 1) / perThread;
    return (int)std::max<size_t>(1, std::min<size_t>(want, std::min<unsigned>(hw, 32u)));
}
template <class F>
void parallel_chunks(size_t n, size_t perThread, F f) {
    int T = n_threads(n, perThread);
    if (T <= 1) { f((size_t)0, n); return; }
    std::vector<std::thread> th;
    size_t step = (n + T - 1) / T;
    for (int t = 0; t < T; t++) {
        size_t a = std::min(n, (size_t)t * step), b = std::min(n, a + step);
        if (a < b) th.emplace_back([=] { f(a, b); });
    }
    for (auto& x : th) x.join();
}

std::string ext_of(const char* path) {   // GetFileExtension (CmdLineParser.inl): text after the last '.'
    const char* dot = std::strrchr(path, '.');
    const char* slash = std::strrchr(path, '/');
    if (!dot || (slash && dot < slash)) return "";
    std::string e(dot + 1);
    for (auto& c : e) c = (char)std::tolower((unsigned char)c);
    return e;
}

// ---- PLY scalar types (plyfile.cu:41-60 type_names + the int8/uint8/... aliases)
struct PlyType { const char* name; int size; int kind; };   // kind: 0 signed int, 1 unsigned int, 2 float
const PlyType kTypes[] = {
    {"char", 1, 0},   {"short", 2, 0},   {"int", 4, 0},   {"uchar", 1, 1},  {"ushort", 2, 1},  {"uint", 4, 1},   {"float", 4, 2},   {"double", 8, 2},
    {"int8", 1, 0},   {"int16", 2, 0},   {"int32", 4, 0}, {"uint8", 1, 1},  {"uint16", 2, 1},  {"uint32", 4, 1}, {"float32", 4, 2}, {"float64", 8, 2},
    {"longlong", 8, 0}, {"ulonglong", 8, 1}, {"int64", 8, 0}, {"uint64", 8, 1},
};
const PlyType* find_type(const std::string& s) {
    for (auto& t : kTypes) if (s == t.name) return &t;
    return nullptr;
}
inline double load_scalar(const unsigned char* p, const PlyType& t, bool swap) {
    unsigned char b[8];
    if (swap) { for (int i = 0; i < t.size; i++) b[i] = p[t.size - 1 - i]; p = b; }
    switch (t.kind * 16 + t.size) {
        case 0 * 16 + 1: { signed char v; std::memcpy(&v, p, 1); return v; }
        case 0 * 16 + 2: { short v; std::memcpy(&v, p, 2); return v; }
        case 0 * 16 + 4: { int v; std::memcpy(&v, p, 4); return v; }
        case 0 * 16 + 8: { long long v; std::memcpy(&v, p, 8); return (double)v; }
        case 1 * 16 + 1: return *p;
        case 1 * 16 + 2: { unsigned short v; std::memcpy(&v, p, 2); return v; }
        case 1 * 16 + 4: { unsigned v; std::memcpy(&v, p, 4); return v; }
        case 1 * 16 + 8: { unsigned long long v; std::memcpy(&v, p, 8); return (double)v; }
        case 2 * 16 + 4: { float v; std::memcpy(&v, p, 4); return v; }
        default: { double v; std::memcpy(&v, p, 8); return v; }
    }
}

struct PlyProp { std::string name; const PlyType* type = nullptr; bool isList = false; const PlyType* countType = nullptr; };

std::vector<std::string> split_ws(const std::string& s) {
    std::vector<std::string> out;
    size_t i = 0;
    while (i < s.size()) {
        while (i < s.size() && std::isspace((unsigned char)s[i])) i++;
        size_t j = i;
        while (j < s.size() && !std::isspace((unsigned char)s[j])) j++;
        if (j > i) out.emplace_back(s.substr(i, j - i));
        i = j;
    }
    return out;
}

int read_ply(const char* path, std::vector<char>& buf, float*& xyz, float*& nrm, int64_t& n) {
    const size_t size = buf.size() - 1;
    size_t pos = 0;
    auto next_line = [&](std::string& line) -> bool {
        if (pos >= size) return false;
        size_t e = pos;
        while (e < size && buf[e] != '\n') e++;
        line.assign(buf.data() + pos, e - pos);
        if (!line.empty() && line.back() == '\r') line.pop_back();
        pos = e < size ? e + 1 : e;
        return true;
    };
    std::string line;
    if (!next_line(line) || split_ws(line).empty() || split_ws(line)[0] != "ply")
        return fail(kErrFormat, std::string("[ERROR] Failed to open ply file for reading: ") + path);
    int fmt = -1;   // 0 ascii, 1 LE, 2 BE
    std::vector<PlyProp> props;   // of the first element
    std::string firstElem;
    int64_t firstCount = 0;
    int nElems = 0;
    bool ended = false;
    while (next_line(line)) {
        auto w = split_ws(line);
        if (w.empty()) continue;
        if (w[0] == "format" && w.size() >= 2) fmt = w[1] == "ascii" ? 0 : w[1] == "binary_little_endian" ? 1 : w[1] == "binary_big_endian" ? 2 : -1;
        else if (w[0] == "element" && w.size() >= 3) {
            if (nElems++ == 0) { firstElem = w[1]; firstCount = std::atoll(w[2].c_str()); }
        } else if (w[0] == "property" && nElems == 1) {
            PlyProp p;
            if (w.size() >= 5 && w[1] == "list") { p.isList = true; p.countType = find_type(w[2]); p.type = find_type(w[3]); p.name = w[4]; }
            else if (w.size() >= 3) { p.type = find_type(w[1]); p.name = w[2]; }
            if (!p.type || (p.isList && !p.countType)) return fail(kErrFormat, "[ERROR] unknown PLY property type in: " + line);
            props.push_back(p);
        } else if (w[0] == "end_header") { ended = true; break; }
    }
    if (!ended || fmt < 0) return fail(kErrFormat, std::string("[ERROR] Failed to open ply file for reading: ") + path);
    // the reference pulls vertices right after the header: `vertex` has to be the first element
    if (firstElem != "vertex") return fail(kErrFormat, "[ERROR] Could not find vertices in ply file");
    static const char* want[6] = {"x", "y", "z", "nx", "ny", "nz"};
    int col[6];
    for (int k = 0; k < 6; k++) {
        col[k] = -1;
        for (size_t j = 0; j < props.size(); j++) if (!props[j].isList && props[j].name == want[k]) col[k] = (int)j;
        if (col[k] < 0) return fail(kErrFormat, std::string("[ERROR] Failed to find property in ply file: ") + want[k]);
    }
    n = firstCount;
    if (n < 0) return fail(kErrFormat, "[ERROR] negative vertex count");
    // the header's count is untrusted: every record takes at least one byte per property (binary) or two
    // characters per property (ascii), so a count the body cannot hold is rejected BEFORE any allocation
    // (division, not multiplication: no wrap-around for absurd counts)
    {
        size_t minRec = 0;
        for (auto& p : props) minRec += (fmt == 0) ? 2 : (size_t)(p.isList ? p.countType->size : p.type->size);
        if (minRec == 0) minRec = 1;
        const size_t bodyBytes = size - pos;
        if ((size_t)n > bodyBytes / minRec + 1) return fail(kErrFormat, fmt == 0 ? "[ERROR] truncated ascii ply body" : "[ERROR] truncated binary ply body");
    }
    xyz = (float*)std::malloc(sizeof(float) * 3 * (size_t)std::max<int64_t>(n, 1));
    nrm = (float*)std::malloc(sizeof(float) * 3 * (size_t)std::max<int64_t>(n, 1));
    if (!xyz || !nrm) return fail(-4, "out of host memory");
    bool anyList = false;
    for (auto& p : props) anyList |= p.isList;
    if (fmt == 0) {
        const char* s = buf.data() + pos;
        std::vector<double> row(props.size());
        for (int64_t i = 0; i < n; i++) {
            for (size_t j = 0; j < props.size(); j++) {
                char* e;
                if (props[j].isList) {
                    long cnt = std::strtol(s, &e, 10);
                    if (e == s) return fail(kErrFormat, "[ERROR] truncated ascii ply body");
                    s = e;
                    for (long q = 0; q < cnt; q++) { std::strtod(s, &e); if (e == s) return fail(kErrFormat, "[ERROR] truncated ascii ply body"); s = e; }
                    row[j] = 0;
                } else {
                    row[j] = std::strtod(s, &e);
                    if (e == s) return fail(kErrFormat, "[ERROR] truncated ascii ply body");
                    s = e;
                }
            }
            for (int k = 0; k < 3; k++) { xyz[3 * i + k] = (float)row[col[k]]; nrm[3 * i + k] = (float)row[col[3 + k]]; }
        }
        return 0;
    }
    const bool swap = (fmt == 2);   // host is little endian (x86-64 / aarch64-le)
    const unsigned char* body = (const unsigned char*)buf.data() + pos;
    const size_t avail = size - pos;
    if (!anyList) {
        std::vector<int> off(props.size());
        int stride = 0;
        for (size_t j = 0; j < props.size(); j++) { off[j] = stride; stride += props[j].type->size; }
        if (stride <= 0 || (size_t)n > avail / (size_t)stride) return fail(kErrFormat, "[ERROR] truncated binary ply body");
        bool plainF32 = !swap;
        for (int k = 0; k < 6; k++) plainF32 &= (props[col[k]].type->kind == 2 && props[col[k]].type->size == 4);
        parallel_chunks((size_t)n, 1 << 18, [&](size_t a, size_t b) {
            for (size_t i = a; i < b; i++) {
                const unsigned char* r = body + i * (size_t)stride;
                if (plainF32) {
                    for (int k = 0; k < 3; k++) { std::memcpy(&xyz[3 * i + k], r + off[col[k]], 4); std::memcpy(&nrm[3 * i + k], r + off[col[3 + k]], 4); }
                } else {
                    for (int k = 0; k < 3; k++) {
                        xyz[3 * i + k] = (float)load_scalar(r + off[col[k]], *props[col[k]].type, swap);
                        nrm[3 * i + k] = (float)load_scalar(r + off[col[3 + k]], *props[col[3 + k]].type, swap);
                    }
                }
            }
        });
        return 0;
    }
    // variable-length records: sequential walk
    const unsigned char* r = body;
    const unsigned char* end = body + avail;
    std::vector<double> row(props.size());
    for (int64_t i = 0; i < n; i++) {
        for (size_t j = 0; j < props.size(); j++) {
            if (props[j].isList) {
                if (r + props[j].countType->size > end) return fail(kErrFormat, "[ERROR] truncated binary ply body");
                long long cnt = (long long)load_scalar(r, *props[j].countType, swap);
                r += props[j].countType->size;
                if (cnt < 0 || (unsigned long long)cnt > (unsigned long long)(end - r) / (unsigned long long)props[j].type->size)
                    return fail(kErrFormat, "[ERROR] truncated binary ply body");
                r += cnt * props[j].type->size;
                row[j] = 0;
            } else {
                if (r + props[j].type->size > end) return fail(kErrFormat, "[ERROR] truncated binary ply body");
                row[j] = load_scalar(r, *props[j].type, swap);
                r += props[j].type->size;
            }
        }
        for (int k = 0; k < 3; k++) { xyz[3 * i + k] = (float)row[col[k]]; nrm[3 * i + k] = (float)row[col[3 + k]]; }
    }
    return 0;
}

int read_bnpts(std::vector<char>& buf, float*& xyz, float*& nrm, int64_t& n) {
    const size_t size = buf.size() - 1;
    n = (int64_t)(size / 24);   // fread(..., 24, count): a trailing partial record is dropped (PointStream.inl:86)
    xyz = (float*)std::malloc(sizeof(float) * 3 * (size_t)std::max<int64_t>(n, 1));
    nrm = (float*)std::malloc(sizeof(float) * 3 * (size_t)std::max<int64_t>(n, 1));
    if (!xyz || !nrm) return fail(-4, "out of host memory");
    const char* b = buf.data();
    parallel_chunks((size_t)n, 1 << 18, [&](size_t a, size_t e) {
        for (size_t i = a; i < e; i++) { std::memcpy(xyz + 3 * i, b + 24 * i, 12); std::memcpy(nrm + 3 * i, b + 24 * i + 12, 12); }
    });
    return 0;
}

int read_ascii(std::vector<char>& buf, float*& xyz, float*& nrm, int64_t& n) {
    // fscanf(" %f %f %f %f %f %f ") until a record fails (PointStream.inl:44-52)
    std::vector<float> v;
    v.reserve(buf.size() / 8);
    const char* s = buf.data();
    for (;;) {
        float c[6];
        int k = 0;
        for (; k < 6; k++) {
            char* e;
            c[k] = std::strtof(s, &e);
            if (e == s) break;
            s = e;
        }
        if (k < 6) break;
        v.insert(v.end(), c, c + 6);
    }
    n = (int64_t)(v.size() / 6);
    xyz = (float*)std::malloc(sizeof(float) * 3 * (size_t)std::max<int64_t>(n, 1));
    nrm = (float*)std::malloc(sizeof(float) * 3 * (size_t)std::max<int64_t>(n, 1));
    if (!xyz || !nrm) return fail(-4, "out of host memory");
    for (int64_t i = 0; i < n; i++) { std::memcpy(xyz + 3 * i, &v[6 * i], 12); std::memcpy(nrm + 3 * i, &v[6 * i + 3], 12); }
    return 0;
}

}  // namespace

extern "C" {

void prbio_free(void* p) { std::free(p); }

int prbio_read_points(const char* path, float** xyz, float** normals, int64_t* n) {
    if (!path || !xyz || !normals || !n) return fail(kErrArg, "prbio_read_points: null argument");
    *xyz = *normals = nullptr;
    *n = 0;
    std::vector<char> buf;
    if (!read_file(path, buf)) {
        std::string e = ext_of(path);
        return fail(kErrIo, std::string(e == "ply" ? "[ERROR] Failed to open ply file for reading: " : "Failed to open file for reading: ") + path);
    }
    std::string e = ext_of(path);
    float *p = nullptr, *q = nullptr;
    int64_t cnt = 0;
    int r = e == "ply" ? read_ply(path, buf, p, q, cnt) : e == "bnpts" ? read_bnpts(buf, p, q, cnt) : read_ascii(buf, p, q, cnt);
    if (r != 0) { std::free(p); std::free(q); return r; }
    *xyz = p; *normals = q; *n = cnt;
    return 0;
}

// Optional cross-pass weld: vertices with bit-identical positions become one (first occurrence kept, order preserved), triangles are
// re-indexed in place.  The reference never welds: every pass appends its own copy of the vertices on a seam (main.cu:3220-3245).
int prbio_weld_mesh(float* v, int64_t nv, int32_t* t, int64_t nt, int64_t* nv_out) {
    if (nv < 0 || nt < 0 || (nv > 0 && !v) || (nt > 0 && !t) || !nv_out) return fail(kErrArg, "prbio_weld_mesh: bad argument");
    struct Key { uint32_t a, b, c; };
    auto key_of = [&](int64_t i) { Key k; std::memcpy(&k, v + 3 * i, 12); if (k.a == 0x80000000u) k.a = 0; if (k.b == 0x80000000u) k.b = 0; if (k.c == 0x80000000u) k.c = 0; return k; };
    // open addressing over a power-of-two table of vertex ids
    size_t cap = 16;
    while (cap < 2 * (size_t)nv + 1) cap <<= 1;
    std::vector<int32_t> slot(cap, -1), remap((size_t)nv);
    int64_t out = 0;
    for (int64_t i = 0; i < nv; i++) {
        const Key k = key_of(i);
        size_t h = ((size_t)k.a * 0x9E3779B97F4A7C15ull) ^ ((size_t)k.b * 0xC2B2AE3D27D4EB4Full) ^ ((size_t)k.c * 0x165667B19E3779F9ull);
        h &= cap - 1;
        for (;;) {
            const int32_t j = slot[h];
            if (j < 0) {
                slot[h] = (int32_t)out;
                if (out != i) std::memcpy(v + 3 * out, v + 3 * i, 12);
                remap[(size_t)i] = (int32_t)out++;
                break;
            }
            const Key kj = key_of(j);
            if (kj.a == k.a && kj.b == k.b && kj.c == k.c) { remap[(size_t)i] = j; break; }
            h = (h + 1) & (cap - 1);
        }
    }
    for (int64_t q = 0; q < 3 * nt; q++) {
        if (t[q] < 0 || t[q] >= nv) return fail(kErrArg, "prbio_weld_mesh: triangle index out of range");
        t[q] = remap[(size_t)t[q]];
    }
    *nv_out = out;
    return 0;
}

int prbio_write_mesh(const char* path, const float* v, int64_t nv, const int32_t* t, int64_t nt, const float center[3], float scale, int binary) {
    if (!path || nv < 0 || nt < 0 || (nv > 0 && !v) || (nt > 0 && !t) || !center) return fail(kErrArg, "prbio_write_mesh: bad argument");
    std::string name(path);
    if (name.size() < 4 || name.compare(name.size() - 4, 4, ".ply") != 0) name += ".ply";
    FILE* fp = std::fopen(name.c_str(), "wb");
    if (!fp) return fail(kErrIo, "Failed to open file for writing: " + name);
    std::fprintf(fp, "ply\nformat %s 1.0\nelement vertex %lld\nproperty float x\nproperty float y\nproperty float z\n"
                     "element face %lld\nproperty list uchar int vertex_indices\nend_header\n",
                 binary ? "binary_little_endian" : "ascii", (long long)nv, (long long)nt);
    bool ok = true;
    auto xf = [&](int64_t i, int a) { return v[3 * i + a] * scale + center[a]; };   // float arithmetic, plyfile.cu:2801-2803
    if (binary) {
        std::vector<float> vb(3 * (size_t)nv);
        parallel_chunks((size_t)nv, 1 << 18, [&](size_t a, size_t b) { for (size_t i = a; i < b; i++) for (int k = 0; k < 3; k++) vb[3 * i + k] = xf((int64_t)i, k); });
        ok &= nv == 0 || std::fwrite(vb.data(), 12, (size_t)nv, fp) == (size_t)nv;
        std::vector<unsigned char> fb(13 * (size_t)nt);
        parallel_chunks((size_t)nt, 1 << 18, [&](size_t a, size_t b) { for (size_t i = a; i < b; i++) { fb[13 * i] = 3; std::memcpy(&fb[13 * i + 1], t + 3 * i, 12); } });
        ok &= nt == 0 || std::fwrite(fb.data(), 13, (size_t)nt, fp) == (size_t)nt;
    } else {
        // chunks formatted in parallel, written in order
        auto emit = [&](size_t n, size_t maxPerItem, auto fmtOne) {
            const size_t chunk = 1 << 16;
            const size_t nChunks = (n + chunk - 1) / chunk;
            const size_t wave = (size_t)n_threads(nChunks, 1) * 4;
            for (size_t c0 = 0; c0 < nChunks && ok; c0 += wave) {
                size_t c1 = std::min(nChunks, c0 + wave);
                std::vector<std::vector<char>> out(c1 - c0);
                parallel_chunks(c1 - c0, 1, [&](size_t a, size_t b) {
                    for (size_t c = a; c < b; c++) {
                        size_t i0 = (c0 + c) * chunk, i1 = std::min(n, i0 + chunk);
                        std::vector<char>& o = out[c];
                        o.resize((i1 - i0) * maxPerItem);
                        size_t w = 0;
                        for (size_t i = i0; i < i1; i++) w += fmtOne(o.data() + w, i);
                        o.resize(w);
                    }
                });
                for (auto& o : out) ok &= o.empty() || std::fwrite(o.data(), 1, o.size(), fp) == o.size();
            }
        };
        emit((size_t)nv, 3 * 32 + 2, [&](char* o, size_t i) {
            return (size_t)std::snprintf(o, 3 * 32 + 2, "%g %g %g \n", (double)xf((int64_t)i, 0), (double)xf((int64_t)i, 1), (double)xf((int64_t)i, 2));
        });
        emit((size_t)nt, 4 * 12 + 2, [&](char* o, size_t i) { return (size_t)std::snprintf(o, 4 * 12 + 2, "3 %d %d %d \n", t[3 * i], t[3 * i + 1], t[3 * i + 2]); });
    }
    ok &= std::fclose(fp) == 0;
    if (!ok) return fail(kErrIo, "write failed: " + name);
    return 0;
}

}  // extern "C"
