// scene_loader.cpp — scene ingest of the host layer: the reference's scenes/*.json schema (SceneData.cpp:61-225)
// -> flattened structure-of-arrays buffers (include/ipt_abi.h: ipt_scene).
//
// The reference parses with nlohmann::json into an AoS std::vector<ObjectData> and every CUDA thread then rebuilds
// polymorphic objects from it (Renderer.cu:69-86, Plane.cu:32-45).  Here a small pull parser reads the file once
// without building a DOM (a 1M-object scene is a ~200 MB file) and the per-rectangle constants are computed once,
// in fp64, on the host.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <charconv>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "host_scene.hpp"

namespace {

// ------------------------------------------------------------------------------------------------ JSON pull parser
struct Reader {
    const char* p;
    const char* end;
    bool ok = true;

    void ws() { while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) p++; }
    bool eat(char c) { ws(); if (p < end && *p == c) { p++; return true; } return false; }
    char peek() { ws(); return p < end ? *p : '\0'; }
    bool fail() { ok = false; return false; }

    bool string(std::string& out)
    {
        out.clear();
        if (!eat('"')) return fail();
        while (p < end && *p != '"') {
            if (*p == '\\') {
                if (++p >= end) return fail();
                switch (*p) {
                    case 'n': out += '\n'; break; case 't': out += '\t'; break; case 'r': out += '\r'; break;
                    case 'b': out += '\b'; break; case 'f': out += '\f'; break;
                    case 'u': out += '?'; p += (end - p > 4) ? 4 : 0; break;   // keys/types of the schema are ASCII
                    default: out += *p;
                }
                p++;
            } else out += *p++;
        }
        if (p >= end) return fail();
        p++;
        return true;
    }
    bool number(double& v)
    {
        ws();
        const char* q = p;
        if (q < end && *q == '+') q++;
        auto r = std::from_chars(q, end, v);
        if (r.ec != std::errc()) return fail();
        p = r.ptr;
        return true;
    }
    // width and height are read into uint32_t upstream (SceneData.cpp:113-114): nlohmann's conversion to an arithmetic type other
    // than its own number types also takes true / false as 1 / 0.  Doubles (radius, vector components) and the enum do not.
    bool arithmetic(double& v)
    {
        const char c = peek();
        if (c == 't' && end - p >= 4 && !std::memcmp(p, "true", 4)) { p += 4; v = 1.0; return true; }
        if (c == 'f' && end - p >= 5 && !std::memcmp(p, "false", 5)) { p += 5; v = 0.0; return true; }
        return number(v);
    }
    bool skip()   // any value
    {
        const char c = peek();
        if (c == '"') { std::string s; return string(s); }
        if (c == '{') {
            p++;
            if (eat('}')) return true;
            do { std::string k; if (!string(k) || !eat(':') || !skip()) return fail(); } while (eat(','));
            return eat('}') || fail();
        }
        if (c == '[') {
            p++;
            if (eat(']')) return true;
            do { if (!skip()) return fail(); } while (eat(','));
            return eat(']') || fail();
        }
        if (c == 't' && end - p >= 4 && !std::memcmp(p, "true", 4)) { p += 4; return true; }
        if (c == 'f' && end - p >= 5 && !std::memcmp(p, "false", 5)) { p += 5; return true; }
        if (c == 'n' && end - p >= 4 && !std::memcmp(p, "null", 4)) { p += 4; return true; }
        double d;
        return number(d);
    }
};

struct Vec3Opt {
    bool present = false;        // key exists
    bool key[3] = {false, false, false};   // "xx" / "yy" / "zz" exist in it, whatever their values are
    bool has[3] = {false, false, false};   // ... and are numbers
    double v[3] = {0, 0, 0};
    // validateVec3tor checks that xx and yy (and yy again) EXIST: SceneData.cpp:30-33.  That is all the reference checks before it
    // prints one of its messages; a missing zz or a component that is not a number makes nlohmann throw or assert later, when the
    // Vec3 is built (:143-145, :196-199, :218-224), and the program dies there without a message.
    bool valid2() const { return present && key[0] && key[1]; }
    bool valid3() const { return valid2() && has[0] && has[1] && has[2]; }
};

bool read_vec(Reader& r, Vec3Opt& out)
{
    out = Vec3Opt();
    out.present = true;
    if (r.peek() != '{') return r.skip();   // present but not an object: fails validation later
    r.p++;
    if (r.eat('}')) return true;
    do {
        std::string k;
        if (!r.string(k) || !r.eat(':')) return r.fail();
        const int i = k == "xx" ? 0 : k == "yy" ? 1 : k == "zz" ? 2 : -1;
        if (i >= 0) { out.key[i] = true; out.has[i] = false; }      // (a repeated key: the last one counts, as in nlohmann)
        if (i >= 0 && (r.peek() == '-' || (r.peek() >= '0' && r.peek() <= '9'))) {
            if (!r.number(out.v[i])) return false;
            out.has[i] = true;
        } else if (!r.skip()) return false;
    } while (r.eat(','));
    return r.eat('}') || r.fail();
}

struct ObjTmp {
    bool hasType = false, typeIsString = false, hasRefl = false, hasRadius = false;
    bool reflIsNumber = false, radiusIsNumber = false;   // anything else makes nlohmann's get<>() throw upstream (uncaught: the reference dies)
    std::string type;
    double refl = 0, radius = 0;
    Vec3Opt color, emission, position, north, east;
};

bool read_object(Reader& r, ObjTmp& o)
{
    o = ObjTmp();
    if (r.peek() != '{') return r.skip();
    r.p++;
    if (r.eat('}')) return true;
    do {
        std::string k;
        if (!r.string(k) || !r.eat(':')) return r.fail();
        if (k == "type") { o.hasType = true; if (r.peek() == '"') { o.typeIsString = true; if (!r.string(o.type)) return false; } else if (!r.skip()) return false; }
        else if (k == "reflection") { o.hasRefl = true; const char c = r.peek(); o.reflIsNumber = c == '-' || (c >= '0' && c <= '9'); if (o.reflIsNumber) { if (!r.number(o.refl)) return false; } else if (!r.skip()) return false; }
        else if (k == "radius") { o.hasRadius = true; const char c = r.peek(); o.radiusIsNumber = c == '-' || (c >= '0' && c <= '9'); if (o.radiusIsNumber) { if (!r.number(o.radius)) return false; } else if (!r.skip()) return false; }
        else if (k == "color") { if (!read_vec(r, o.color)) return false; }
        else if (k == "emission") { if (!read_vec(r, o.emission)) return false; }
        else if (k == "position") { if (!read_vec(r, o.position)) return false; }
        else if (k == "north") { if (!read_vec(r, o.north)) return false; }
        else if (k == "east") { if (!read_vec(r, o.east)) return false; }
        else if (!r.skip()) return false;
    } while (r.eat(','));
    return r.eat('}') || r.fail();
}

void msg(char* out, size_t n, const char* s)
{
    if (out && n) std::snprintf(out, n, "%s", s);
}

inline void norm3(double* v)   // Vec3.hpp:48-51: v * (1/sqrt(v.v))
{
    const double s = 1 / std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    v[0] *= s; v[1] *= s; v[2] *= s;
}
inline void cross3(const double* a, const double* b, double* o)   // Vec3.hpp:69-72
{
    o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}
inline double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

}  // namespace

// Appends one primitive in object (JSON) order.  The rectangle constants restate Plane.cu:32-45 (normal) and solve
// Plane.cu:87-100 for the hit position: with s = signed distance of the hit from the centre line along the in-plane
// axis perpendicular to an edge pair, h = half the distance between the two edge lines and L = the side length the
// reference compares against, the reference's test "d(bottom)+d(top) == L +- 1e-4" reads |L - 2 max(h,|s|)| <= 1e-4,
// i.e. lo <= |s| <= hi with hi = (L+1e-4)/2 and lo = 0 when north is perpendicular to east (L = 2h), else (L-1e-4)/2.
void ipt_host_scene::add(int type, double radius, const double* north, const double* east, const double* position,
                         const double* emission, const double* color, int reflection)
{
    const uint32_t obj = (uint32_t)mat_reflection.size();
    for (int i = 0; i < 3; i++) { mat_color.push_back(color[i]); mat_emission.push_back(emission[i]); }
    mat_reflection.push_back(reflection);
    if (type == 0) {
        sphere_cxyzr.insert(sphere_cxyzr.end(), {position[0], position[1], position[2], radius});
        sphere_object.push_back(obj);
        return;
    }
    const double M = 1e-4;   // scene/cuda/objects/Constants.hpp:8
    double n[3], u[3], v[3];
    cross3(north, east, n);
    norm3(n);                                   // Plane.cu:36
    cross3(east, n, u); norm3(u);               // in-plane, perpendicular to east  ("vertical" distances, Plane.cu:89-92)
    cross3(north, n, v); norm3(v);              // in-plane, perpendicular to north ("horizontal" distances, :94-97)
    const double L_v = 2 * std::sqrt(dot3(north, north));   // distanceVertical_   = |bottomLeft - topLeft|     (:44)
    const double L_h = 2 * std::sqrt(dot3(east, east));     // distanceHorizontal_ = |bottomLeft - bottomRight| (:43)
    const double h_u = std::fabs(dot3(north, u)), h_v = std::fabs(dot3(east, v));
    const double u_lo = (L_v - 2 * h_u <= M) ? 0.0 : (L_v - M) / 2, u_hi = (L_v + M) / 2;
    const double v_lo = (L_h - 2 * h_v <= M) ? 0.0 : (L_h - M) / 2, v_hi = (L_h + M) / 2;
    rect_plane.insert(rect_plane.end(), {n[0], n[1], n[2], dot3(n, position)});
    rect_u.insert(rect_u.end(), {u[0], u[1], u[2], dot3(u, position)});
    rect_v.insert(rect_v.end(), {v[0], v[1], v[2], dot3(v, position)});
    rect_bounds.insert(rect_bounds.end(), {u_lo, u_hi, v_lo, v_hi});
    rect_object.push_back(obj);
    rect_center.insert(rect_center.end(), {position[0], position[1], position[2]});
    rect_north.insert(rect_north.end(), {north[0], north[1], north[2]});
    rect_east.insert(rect_east.end(), {east[0], east[1], east[2]});
}

void ipt_host_scene::refresh_view()
{
    view.n_spheres = (uint32_t)sphere_object.size();
    view.n_rects = (uint32_t)rect_object.size();
    view.n_objects = (uint32_t)mat_reflection.size();
    view.sphere_cxyzr = sphere_cxyzr.data(); view.sphere_object = sphere_object.data();
    view.rect_plane = rect_plane.data(); view.rect_u = rect_u.data(); view.rect_v = rect_v.data();
    view.rect_bounds = rect_bounds.data(); view.rect_object = rect_object.data();
    view.mat_color = mat_color.data(); view.mat_emission = mat_emission.data(); view.mat_reflection = mat_reflection.data();
    view.n_bvh_nodes = (uint32_t)bvh_nodes.size(); view.n_bvh_slots = (uint32_t)bvh_slot_prim.size();
    view.bvh_nodes = bvh_nodes.empty() ? nullptr : bvh_nodes.data();
    view.bvh_slot_prim = bvh_slot_prim.empty() ? nullptr : bvh_slot_prim.data();
    const bool grid = !grid_cell_start.empty() && !bvh_nodes.empty();
    for (int k = 0; k < 3; k++) { view.grid_res[k] = grid ? grid_res[k] : 0; view.grid_lo[k] = grid_lo[k]; view.grid_cell[k] = grid_cell[k]; }
    view.n_grid_big = grid ? (uint32_t)grid_big.size() : 0; view.n_grid_refs = grid ? (uint32_t)grid_refs.size() : 0; view.reserved1 = 0;
    view.grid_cell_start = grid ? grid_cell_start.data() : nullptr;
    view.grid_refs = grid && !grid_refs.empty() ? grid_refs.data() : nullptr;
    view.grid_big = grid && !grid_big.empty() ? grid_big.data() : nullptr;
}

namespace {
// The bytes of a scene file: regular files are mapped (config 5's file is 200 MB: reading it into a string cost a zero fill and
// a copy, 0.17 s), anything else that can be read (a pipe, /proc) is read in pieces up to 1 GB; a directory is not a scene.
struct FileBytes {
    const char* data = nullptr;
    size_t size = 0;
    void* mapped = nullptr;
    std::string owned;
    bool open(const char* path)
    {
        const int fd = path ? ::open(path, O_RDONLY) : -1;
        if (fd < 0) return false;
        struct stat st;
        if (fstat(fd, &st) != 0 || S_ISDIR(st.st_mode)) { ::close(fd); return false; }
        if (S_ISREG(st.st_mode) && st.st_size > 0) {
            void* m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
            if (m != MAP_FAILED) { mapped = m; data = (const char*)m; size = (size_t)st.st_size; ::close(fd); return true; }
        }
        char piece[1 << 16];
        for (;;) {
            const ssize_t n = ::read(fd, piece, sizeof(piece));
            if (n <= 0) break;
            owned.append(piece, (size_t)n);
            if (owned.size() > ((size_t)1 << 30)) { ::close(fd); return false; }
        }
        ::close(fd);
        data = owned.data(); size = owned.size();
        return true;
    }
    ~FileBytes() { if (mapped) munmap(mapped, size); }
};
ipt_host_scene* load_scene(const char* path, char* message, size_t message_len);
}  // namespace

extern "C" ipt_host_scene* ipt_host_load_scene(const char* path, char* message, size_t message_len)
{
    try { return load_scene(path, message, message_len); }
    catch (...) { msg(message, message_len, "Could not load provided json file!"); return nullptr; }   // nothing is thrown across the C ABI
}

namespace {
ipt_host_scene* load_scene(const char* path, char* message, size_t message_len)
{
    msg(message, message_len, "");
    const auto T0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) { if (std::getenv("IPT_VERBOSE")) std::fprintf(stderr, "[load] %s %.3f s\n", what, std::chrono::duration<double>(std::chrono::steady_clock::now() - T0).count()); };
    FileBytes buf;
    if (!buf.open(path)) { msg(message, message_len, "Could not load provided json file!"); return nullptr; }   // SceneData.cpp:66-70
    lap("read");

    Reader r{buf.data, buf.data + buf.size};
    bool hasW = false, hasH = false, hasCamera = false, hasObjects = false, objectsIsArray = false;
    double W = 0, H = 0;
    Vec3Opt camDir, camPos, camOri;
    std::vector<ObjTmp> objs;
    // Malformed JSON makes the reference abort with an uncaught nlohmann exception; here it is a load failure.
    bool parsed = r.eat('{');
    if (parsed && !r.eat('}')) {
        do {
            std::string k;
            if (!r.string(k) || !r.eat(':')) { parsed = false; break; }
            if (k == "width") { hasW = true; if (!r.arithmetic(W)) { parsed = false; break; } }
            else if (k == "height") { hasH = true; if (!r.arithmetic(H)) { parsed = false; break; } }
            else if (k == "camera") {
                hasCamera = true;
                if (r.peek() != '{') { if (!r.skip()) { parsed = false; break; } continue; }
                r.p++;
                if (r.eat('}')) continue;
                bool okc = true;
                do {
                    std::string ck;
                    if (!r.string(ck) || !r.eat(':')) { okc = false; break; }
                    if (ck == "direction") okc = read_vec(r, camDir);
                    else if (ck == "position") okc = read_vec(r, camPos);
                    else if (ck == "orientation") okc = read_vec(r, camOri);
                    else okc = r.skip();
                } while (okc && r.eat(','));
                if (!okc || !r.eat('}')) { parsed = false; break; }
            } else if (k == "objects") {
                hasObjects = true;
                if (r.peek() == '{') {
                    // not an array but an object: the reference's range-for (SceneData.cpp:158) walks its VALUES, in nlohmann's
                    // std::map order (sorted by key, a repeated key keeps its last value)
                    r.p++;
                    std::vector<std::pair<std::string, ObjTmp>> kv;
                    bool okm = true;
                    if (!r.eat('}')) {
                        do {
                            std::string key;
                            ObjTmp t;
                            if (!r.string(key) || !r.eat(':') || !read_object(r, t)) { okm = false; break; }
                            auto it = std::find_if(kv.begin(), kv.end(), [&](const std::pair<std::string, ObjTmp>& e) { return e.first == key; });
                            if (it != kv.end()) it->second = t; else kv.emplace_back(key, t);
                        } while (r.eat(','));
                        if (okm && !r.eat('}')) okm = false;
                    }
                    if (!okm) { parsed = false; break; }
                    std::stable_sort(kv.begin(), kv.end(), [](const std::pair<std::string, ObjTmp>& a, const std::pair<std::string, ObjTmp>& b) { return a.first < b.first; });
                    for (auto& e : kv) objs.push_back(e.second);
                    continue;
                }
                if (r.peek() != '[') {
                    // a scalar: nlohmann iterates it as ONE element (which then fails validateObject), null as none
                    const bool is_null = r.peek() == 'n';
                    if (!r.skip()) { parsed = false; break; }
                    if (!is_null) objs.emplace_back();
                    continue;
                }
                objectsIsArray = true;
                r.p++;
                if (r.eat(']')) continue;
                bool oko = true;
                const size_t parallel_min = std::getenv("IPT_PARSE_PARALLEL_MIN") ? (size_t)std::atoll(std::getenv("IPT_PARSE_PARALLEL_MIN")) : (size_t)(8u << 20);
                if (buf.size < parallel_min) {
                    do { objs.emplace_back(); oko = read_object(r, objs.back()); } while (oko && r.eat(','));
                } else {
                    // large scene (a 1M-object file is ~200 MB): a structural scan finds where every array element starts, then
                    // the elements are parsed on all host threads into their slots (order preserved).  The scan itself runs on
                    // all threads too (it was 40 % of the load): the rest of the file is cut into equal parts, every part is
                    // scanned as if it began outside a string and keeps the commas at the lowest nesting depth it reaches (the
                    // depth at its first byte is not known yet, but the commas between array elements are the ones at the lowest
                    // depth that occurs before the array ends); one serial pass then chains the parts - a part whose assumption
                    // was wrong (it began inside a string, or the array ends inside it) is scanned again with its true state.
                    const char* const a0 = r.p;
                    unsigned T = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
                    if (const char* e = std::getenv("IPT_HOST_THREADS")) T = (unsigned)std::max(1, std::min(256, std::atoi(e)));
                    // Whether a part begins inside a string is not known either (a third of the bytes of a scene file are inside
                    // strings), so every part is scanned for both cases at once: a quote flips both, a structural byte counts
                    // for the case in which it lies outside a string.
                    struct Acc { std::vector<const char*> commas; int depth = 0, lowest = 0; };
                    struct Part { Acc acc[2]; bool in_str0 = false; const char* next = nullptr; };   // acc[1]: the part began inside a string
                    const unsigned NP = T * 4;             // a part scanned again (the one the array ends in) costs 1 / NP of the scan
                    std::vector<Part> parts(NP);
                    auto cut = [&](unsigned t) { return a0 + (size_t)(r.end - a0) * t / NP; };
                    auto scan_guess = [&](unsigned t) {
                        Part& P = parts[t];
                        const char* q = cut(t);
                        const char* const e = cut(t + 1);
                        bool in0 = false;                  // inside a string, if the part began outside one
                        for (; q < e; q++) {
                            const char ch = *q;
                            if (ch == '\\') { q++; continue; }      // escapes the next byte (valid only inside a string, whichever case is true)
                            if (ch == '"') { in0 = !in0; continue; }
                            Acc& A = P.acc[in0 ? 1 : 0];
                            if (ch == '{' || ch == '[') A.depth++;
                            else if (ch == '}' || ch == ']') { if (--A.depth < A.lowest) { A.lowest = A.depth; A.commas.clear(); } }
                            else if (ch == ',' && A.depth == A.lowest) A.commas.push_back(q);
                        }
                        P.in_str0 = in0; P.next = q;       // q > e: the first byte of the next part is escaped
                    };
                    {
                        std::vector<std::thread> th;
                        for (unsigned w = 0; w < T; w++) th.emplace_back([&, w] { for (unsigned t = w; t < NP; t += T) scan_guess(t); });
                        for (auto& x : th) x.join();
                    }
                    lap("scan (parts)");
                    std::vector<const char*> commas;
                    const char* close = nullptr;          // the array's closing bracket
                    const char* q = a0;                    // everything before q is accounted for
                    int depth = 0;
                    bool in_str = false;
                    unsigned rescans = 0;
                    for (unsigned t = 0; t < NP && !close; t++) {
                        const Part& P = parts[t];
                        const Acc& A = P.acc[in_str ? 1 : 0];
                        const char* const e = cut(t + 1);
                        if (q == cut(t) && depth + A.lowest >= 0) {                     // the array goes on beyond this part
                            if (depth + A.lowest == 0) commas.insert(commas.end(), A.commas.begin(), A.commas.end());
                            depth += A.depth; in_str = in_str ? !P.in_str0 : P.in_str0; q = P.next;
                            continue;
                        }
                        rescans++;
                        for (; q < e; q++) {                                            // byte by byte, with the true state
                            const char ch = *q;
                            if (in_str) { if (ch == '\\') q++; else if (ch == '"') in_str = false; continue; }
                            if (ch == '"') in_str = true;
                            else if (ch == '{' || ch == '[') depth++;
                            else if (ch == '}' || ch == ']') { if (depth == 0) { close = q; break; } depth--; }
                            else if (ch == ',' && depth == 0) commas.push_back(q);
                        }
                    }
                    std::vector<const char*> starts;
                    auto after_ws = [&](const char* c) { while (c < r.end && (*c == ' ' || *c == '\n' || *c == '\t' || *c == '\r')) c++; return c; };
                    if (!close) oko = false;
                    else {
                        starts.reserve(commas.size() + 2);
                        starts.push_back(after_ws(a0));
                        for (const char* c : commas) starts.push_back(after_ws(c + 1));
                    }
                    q = close;
                    if (std::getenv("IPT_VERBOSE")) std::fprintf(stderr, "[load] %u parts, %u scanned again, %zu elements\n", NP, rescans, starts.size());
                    lap("scan");
                    if (oko) {
                        starts.push_back(q);
                        const size_t n = starts.size() - 1;
                        objs.resize(n);
                        std::vector<char> okv(T, 1);
                        std::vector<std::thread> th;
                        for (unsigned t = 0; t < T; t++)
                            th.emplace_back([&, t] {
                                for (size_t i = n * t / T; i < n * (t + 1) / T; i++) {
                                    Reader e{starts[i], starts[i + 1]};
                                    if (!read_object(e, objs[i]) || !e.ok) { okv[t] = 0; return; }
                                    e.ws();                                   // nothing but the separating comma may follow
                                    if (e.p < e.end && *e.p == ',') { e.p++; e.ws(); }
                                    if (e.p != e.end) { okv[t] = 0; return; }
                                }
                            });
                        for (auto& x : th) x.join();
                        for (char v : okv) oko = oko && v;
                        r.p = q;
                    }
                }
                if (!oko || !r.eat(']')) { parsed = false; break; }
            } else if (!r.skip()) { parsed = false; break; }
        } while (r.eat(','));
        if (parsed && !r.eat('}')) parsed = false;
    }
    if (!parsed || !r.ok) { msg(message, message_len, "Could not load provided json file!"); return nullptr; }
    lap("parse");

    // validation in the reference's order: basic data, camera, objects (SceneData.cpp:72-94)
    if (!hasH || !hasW) { msg(message, message_len, "Missing height or witdh data!"); return nullptr; }            // :98-111
    if (!hasCamera) { msg(message, message_len, "No camera data!"); return nullptr; }                              // :119-123
    if (!camDir.present || !camPos.present || !camOri.present) { msg(message, message_len, "Camera data could not be read!"); return nullptr; }   // :126-131
    if (!camDir.valid2() || !camPos.valid2() || !camOri.valid2()) { msg(message, message_len, "Camera data could not be parsed!"); return nullptr; }   // :137-141
    if (!camDir.valid3() || !camPos.valid3() || !camOri.valid3()) { msg(message, message_len, "Camera data could not be parsed!"); return nullptr; }   // upstream dies in :143-145 (zz unchecked, values unchecked)
    if (!hasObjects) { msg(message, message_len, "No objects data!"); return nullptr; }                            // :152-156
    (void)objectsIsArray;

    auto* s = new ipt_host_scene();
    s->view.width = (uint32_t)W; s->view.height = (uint32_t)H;
    norm3(camDir.v); norm3(camOri.v);                                                                             // :143-145
    std::memcpy(s->view.cam_origin, camPos.v, 24); std::memcpy(s->view.cam_dir, camDir.v, 24); std::memcpy(s->view.cam_orient, camOri.v, 24);
    const double zero[3] = {0, 0, 0};
    {   // room for everything at once: a million objects would otherwise grow sixteen arrays by doubling
        size_t n_sph = 0;
        for (const ObjTmp& o : objs) n_sph += o.type == "sphere";
        const size_t n_rec = objs.size() - n_sph;
        s->mat_color.reserve(3 * objs.size()); s->mat_emission.reserve(3 * objs.size()); s->mat_reflection.reserve(objs.size());
        s->sphere_cxyzr.reserve(4 * n_sph); s->sphere_object.reserve(n_sph);
        s->rect_plane.reserve(4 * n_rec); s->rect_u.reserve(4 * n_rec); s->rect_v.reserve(4 * n_rec); s->rect_bounds.reserve(4 * n_rec);
        s->rect_object.reserve(n_rec); s->rect_center.reserve(3 * n_rec); s->rect_north.reserve(3 * n_rec); s->rect_east.reserve(3 * n_rec);
    }
    for (const ObjTmp& o : objs) {
        const char* err = nullptr;
        if (!o.color.present || !o.emission.present || !o.position.present || !o.hasRefl || !o.hasType ||
            !o.color.valid2() || !o.emission.valid2() || !o.position.valid2())
            err = "Could not validate object data!";                                                               // :35-51,:160-164
        else if (!o.typeIsString || (o.type != "sphere" && o.type != "plane")) err = "Unknown object type";        // :166-177
        else if (o.type == "sphere" && !o.hasRadius) err = "Broken sphere object! ";                               // :185-189
        // a "radius" / "reflection" that is not a number: upstream nlohmann throws type_error out of main (the program dies
        // without a message, :196-199, :218-224); here the file is refused with the message of the check next to it
        else if (o.type == "sphere" && !o.radiusIsNumber) err = "Broken sphere object! ";
        else if (o.type == "plane" && (!o.north.present || !o.east.present)) err = "Broken plane object! ";        // :205-209
        // from here on the reference has no message: it dies building the object (missing zz, values that are not numbers)
        else if (o.type == "plane" && (!o.north.valid3() || !o.east.valid3())) err = "Broken plane object! ";
        else if (!o.color.valid3() || !o.emission.valid3() || !o.position.valid3() || !o.reflIsNumber) err = "Could not validate object data!";
        if (err) { msg(message, message_len, err); delete s; return nullptr; }
        if (o.type == "sphere") s->add(0, o.radius, zero, zero, o.position.v, o.emission.v, o.color.v, (int)o.refl);
        else s->add(1, 0.0, o.north.v, o.east.v, o.position.v, o.emission.v, o.color.v, (int)o.refl);
    }
    if (s->mat_reflection.empty()) { msg(message, message_len, "Object list empty! Cannot build scene"); delete s; return nullptr; }   // :87-91
    s->refresh_view();
    lap("flatten");
    return s;
}
}  // namespace

static ipt_host_scene* from_objects_impl(const void* objects, uint32_t n, uint32_t width, uint32_t height, const double* cam)
{
    if (!objects || !cam || n == 0) return nullptr;
    auto* s = new ipt_host_scene();
    s->view.width = width; s->view.height = height;
    std::memcpy(s->view.cam_origin, cam, 24); std::memcpy(s->view.cam_dir, cam + 3, 24); std::memcpy(s->view.cam_orient, cam + 6, 24);
    const char* base = (const char*)objects;
    for (uint32_t i = 0; i < n; i++) {
        const char* o = base + (size_t)i * 144;   // ObjectData.hpp:15-31
        int32_t type, refl;
        double radius, north[3], east[3], pos[3], emi[3], col[3];
        std::memcpy(&type, o, 4); std::memcpy(&radius, o + 8, 8); std::memcpy(north, o + 16, 24); std::memcpy(east, o + 40, 24);
        std::memcpy(pos, o + 64, 24); std::memcpy(emi, o + 88, 24); std::memcpy(col, o + 112, 24); std::memcpy(&refl, o + 136, 4);
        s->add(type == 0 ? 0 : 1, radius, north, east, pos, emi, col, refl);
    }
    s->refresh_view();
    return s;
}

extern "C" void ipt_host_free_scene(ipt_host_scene* s) { delete s; }
extern "C" const ipt_scene* ipt_host_scene_view(const ipt_host_scene* s) { return s ? &s->view : nullptr; }
extern "C" void ipt_host_set_size(ipt_host_scene* s, uint32_t w, uint32_t h) { if (s) { s->view.width = w; s->view.height = h; } }

extern "C" ipt_host_scene* ipt_host_from_objects(const void* objects, uint32_t n, uint32_t width, uint32_t height, const double* cam)
{
    try { return from_objects_impl(objects, n, width, height, cam); }
    catch (...) { return nullptr; }                 // nothing is thrown across the C ABI
}
