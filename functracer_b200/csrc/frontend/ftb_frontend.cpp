// ftb_frontend.cpp — see ftb_frontend.h.  Host-side only; no CUDA.
#include "ftb_frontend.h"

#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include <zlib.h>

namespace {

thread_local std::string g_err;

struct V3 {
    double x, y, z;
};
inline V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 normalise(V3 v)  // CommonTypes.fs:63-67
{
    double l = std::sqrt(dot(v, v));
    if (l < 0.0000001) return v;
    double s = 1.0 / l;
    return {s * v.x, s * v.y, s * v.z};
}
const double kPi = 3.14159265358979323846;
inline double degToRad(double d) { return d * 1.0 * (kPi / 180.0); }  // CommonTypes.fs:98-99

// ---- Transform.matrix / inverse (Transform.fs:47-71) ------------------------------------------
void matTranslate(double* m, double x, double y, double z)
{
    const double t[12] = {1, 0, 0, x, 0, 1, 0, y, 0, 0, 1, z};
    std::memcpy(m, t, sizeof(t));
}
void matScale(double* m, double x, double y, double z)
{
    const double t[12] = {x, 0, 0, 0, 0, y, 0, 0, 0, 0, z, 0};
    std::memcpy(m, t, sizeof(t));
}
void matRotate(double* m, V3 u, double angle)  // u already normalised (Transform.fs:37-38)
{
    double c = std::cos(angle), invc = 1.0 - c, s = std::sin(angle);
    const double t[12] = {c + invc * u.x * u.x,       invc * u.x * u.y - s * u.z, invc * u.x * u.z + s * u.y, 0.0,
                          invc * u.x * u.y + s * u.z, c + invc * u.y * u.y,       invc * u.y * u.z - s * u.x, 0.0,
                          invc * u.x * u.z - s * u.y, invc * u.y * u.z + s * u.x, c + invc * u.z * u.z,       0.0};
    std::memcpy(m, t, sizeof(t));
}

struct Tri {
    V3 a, b, c;
};

// ---- Plane.intersect for a general plane (Plane.fs:9-20), Plane.isAbove (:22-23) ----------------
inline bool isAbove(V3 p0, V3 n, V3 point) { return dot(sub(point, p0), n) >= 0.0; }
V3 edgeIntersection(V3 p0, V3 n, V3 a, V3 b)  // Triangle.fs:8-10 (Seq.first read as tryHead)
{
    const double eps = 0.0000001;
    V3 o = a, d = normalise(sub(b, a));
    double num = dot(sub(p0, o), n);
    double denom = dot(d, n);
    if (std::fabs(denom) < eps) return o;  // t = 0 hit, or no hit (the reference would throw on .Value)
    double t = num / denom;
    return {o.x + t * d.x, o.y + t * d.y, o.z + t * d.z};
}
typedef std::vector<Tri> Tris;
void sliceP(V3 p0, V3 n, const Tri& t, Tris& single, Tris& two)  // slice' (Triangle.fs:13-22)
{
    auto inter = [&](V3 x, V3 y) { return edgeIntersection(p0, n, x, y); };
    single.push_back({t.a, inter(t.a, t.b), inter(t.a, t.c)});
    two.push_back({inter(t.b, t.a), t.b, t.c});
    two.push_back({t.c, inter(t.c, t.a), inter(t.b, t.a)});
}
void slice(V3 p0, V3 n, const Tri& t, Tris& fst, Tris& snd)  // Triangle.fs:24-41
{
    bool aAbove = isAbove(p0, n, t.a), bAbove = isAbove(p0, n, t.b), cAbove = isAbove(p0, n, t.c);
    Tris x, y;  // the pair before the final flip
    if (aAbove == bAbove && bAbove == cAbove) {
        x.push_back(t);
    } else if (aAbove == bAbove) {
        sliceP(p0, n, {t.c, t.a, t.b}, y, x);  // |> flip true
    } else if (aAbove == cAbove) {
        sliceP(p0, n, {t.b, t.c, t.a}, y, x);  // |> flip true
    } else {
        sliceP(p0, n, t, x, y);
    }
    if (!aAbove) std::swap(x, y);  // |> flip (not aAbove)
    fst.insert(fst.end(), x.begin(), x.end());
    snd.insert(snd.end(), y.begin(), y.end());
}

// ---- scene under construction -------------------------------------------------------------------
struct Builder {
    std::vector<ftb_node> nodes;
    std::vector<int32_t> children;
    std::vector<ftb_transform> transforms;
    std::vector<ftb_material> materials;
    std::vector<ftb_texture> textures;
    std::vector<ftb_image> images;
    std::vector<std::vector<uint8_t>> imageData;
    std::vector<ftb_mesh> meshes;
    std::vector<ftb_bsp_node> bspNodes;
    std::vector<ftb_bsp_leaf> bspLeaves;
    std::vector<double> triangles;
    std::vector<ftb_light> lights;
    std::string assetDir;

    int addNode(int kind, int a, int b)
    {
        nodes.push_back({kind, a, b, 0});
        return (int)nodes.size() - 1;
    }
    int addTriangle(const Tri& t)
    {
        const double v[9] = {t.a.x, t.a.y, t.a.z, t.b.x, t.b.y, t.b.z, t.c.x, t.c.y, t.c.z};
        triangles.insert(triangles.end(), v, v + 9);
        return (int)(triangles.size() / 9) - 1;
    }
    // BspMesh.compile (BspMesh.fs:51-65) -> link (>=0 branch, <0 ~leaf)
    int compileBsp(int maxDepth, const Tris& tris)
    {
        auto leaf = [&]() {
            ftb_bsp_leaf lf = {(int32_t)(triangles.size() / 9), (int32_t)tris.size()};
            for (const Tri& t : tris) addTriangle(t);
            bspLeaves.push_back(lf);
            return ~((int)bspLeaves.size() - 1);
        };
        if (maxDepth == 0 || tris.empty()) return leaf();
        // trianglesBoundry / pointsBoundry (BspMesh.fs:49, BoundingBox.fs:9-22)
        V3 mn = tris[0].a, mx = tris[0].a;
        for (const Tri& t : tris)
            for (V3 p : {t.a, t.b, t.c}) {
                mn.x = std::min(mn.x, p.x); mn.y = std::min(mn.y, p.y); mn.z = std::min(mn.z, p.z);
                mx.x = std::max(mx.x, p.x); mx.y = std::max(mx.y, p.y); mx.z = std::max(mx.z, p.z);
            }
        // optimalSplit (BspMesh.fs:30-46)
        double wx = std::fabs(mx.x - mn.x) / 2.0, wy = std::fabs(mx.y - mn.y) / 2.0, wz = std::fabs(mx.z - mn.z) / 2.0;
        V3 p0, n;
        if (wx > wy && wx > wz) { p0 = {(mn.x + mx.x) / 2.0, 0.0, 0.0}; n = {1, 0, 0}; }
        else if (wy > wz) { p0 = {0.0, (mn.y + mx.y) / 2.0, 0.0}; n = {0, 1, 0}; }
        else { p0 = {0.0, 0.0, (mn.z + mx.z) / 2.0}; n = {0, 0, 1}; }
        Tris left, right;
        for (const Tri& t : tris) slice(p0, n, t, left, right);
        size_t triCount = tris.size();
        if (left.size() >= triCount || right.size() >= triCount) return leaf();
        int idx = (int)bspNodes.size();
        bspNodes.push_back({{mn.x, mn.y, mn.z}, {mx.x, mx.y, mx.z}, 0, 0});
        int l = compileBsp(maxDepth - 1, left);
        int r = compileBsp(maxDepth - 1, right);
        bspNodes[idx].left = l;
        bspNodes[idx].right = r;
        return idx;
    }
};

std::string baseName(const std::string& p)
{
    size_t k = p.find_last_of("/\\");
    return k == std::string::npos ? p : p.substr(k + 1);
}
bool readFile(const std::string& path, std::string& out)
{
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    std::ostringstream ss;
    ss << f.rdbuf();
    out = ss.str();
    return true;
}
bool resolveAsset(const Builder& b, const std::string& name, const char* altExt, std::string& data, std::string& used)
{
    std::vector<std::string> cands;
    cands.push_back(name);
    std::string fixed = name;
    for (char& c : fixed) if (c == '\\') c = '/';
    cands.push_back(fixed);
    std::string base = baseName(name);
    if (!b.assetDir.empty()) {
        cands.push_back(b.assetDir + "/" + fixed);
        cands.push_back(b.assetDir + "/" + base);
        if (altExt) {
            size_t dotp = base.find_last_of('.');
            cands.push_back(b.assetDir + "/" + (dotp == std::string::npos ? base : base.substr(0, dotp)) + altExt);
        }
    }
    for (const std::string& c : cands)
        if (readFile(c, data)) { used = c; return true; }
    return false;
}

// ---- PlyParser.fs ----------------------------------------------------------------------------------
Tris parsePly(const std::string& text)
{
    std::istringstream in(text);
    std::string line;
    auto fail = [](const std::string& m) -> void { throw std::runtime_error("PLY: " + m); };
    if (!std::getline(in, line) || line.substr(0, 3) != "ply") fail("expected 'ply'");  // :20
    long nv = -1, nf = -1;
    bool ended = false;
    while (std::getline(in, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (line.rfind("format", 0) == 0 || line.rfind("comment", 0) == 0 || line.rfind("property", 0) == 0) continue;  // :22-26
        if (line.rfind("element vertex ", 0) == 0) { nv = std::atol(line.c_str() + 15); continue; }                    // :28
        if (line.rfind("element face ", 0) == 0) { nf = std::atol(line.c_str() + 13); continue; }                      // :29
        if (line.rfind("end_header", 0) == 0) { ended = true; break; }                                                   // :39
        fail("unexpected header line '" + line + "'");
    }
    if (!ended || nv < 0 || nf < 0) fail("incomplete header");
    std::vector<V3> verts((size_t)nv);
    for (long i = 0; i < nv; ++i) {  // pvertex :42-49: 5 floats, x y z confidence intensity
        if (!std::getline(in, line)) fail("missing vertex line");
        double x, y, z, c, s;
        if (std::sscanf(line.c_str(), "%lf %lf %lf %lf %lf", &x, &y, &z, &c, &s) != 5) fail("vertex needs 5 floats");
        verts[(size_t)i] = {x, y, z};
    }
    Tris tris;
    tris.reserve((size_t)nf);
    for (long i = 0; i < nf; ++i) {  // pface :51-57: "3 a b c"
        if (!std::getline(in, line)) fail("missing face line");
        long a, b, c;
        if (std::sscanf(line.c_str(), "3 %ld %ld %ld", &a, &b, &c) != 3) fail("face must be '3 a b c'");
        if (a < 0 || b < 0 || c < 0 || a >= nv || b >= nv || c >= nv) fail("face index out of range");
        tris.push_back({verts[(size_t)a], verts[(size_t)b], verts[(size_t)c]});
    }
    return tris;
}

// Binary PPM (P6, maxval 255) -> Rgb24, the stand-in for Image.Load<Rgb24> (Textures/Image.fs:23-26)
bool decodePpm(const std::string& data, std::vector<uint8_t>& rgb, int& w, int& h)
{
    size_t pos = 0;
    auto token = [&]() {
        for (;;) {
            while (pos < data.size() && std::isspace((unsigned char)data[pos])) ++pos;
            if (pos < data.size() && data[pos] == '#') { while (pos < data.size() && data[pos] != '\n') ++pos; continue; }
            break;
        }
        size_t s = pos;
        while (pos < data.size() && !std::isspace((unsigned char)data[pos])) ++pos;
        return data.substr(s, pos - s);
    };
    if (token() != "P6") return false;
    w = std::atoi(token().c_str());
    h = std::atoi(token().c_str());
    int maxv = std::atoi(token().c_str());
    if (w <= 0 || h <= 0 || maxv != 255) return false;
    ++pos;  // single whitespace after maxval
    if (data.size() - pos < (size_t)w * h * 3) return false;
    rgb.assign(data.begin() + (long)pos, data.begin() + (long)pos + (long)w * h * 3);
    return true;
}

// ---- SceneParser.fs ----------------------------------------------------------------------------------
struct ParseError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

typedef std::function<int(int)> GFunc;  // SceneGraph -> SceneGraph on node indices

struct Options {
    ftb_camera camera;
    int width, height, spp, sampling;
};

struct Parser {
    const std::string& s;
    size_t pos = 0;
    Builder& b;
    Options& opt;
    Parser(const std::string& text, Builder& bb, Options& oo) : s(text), b(bb), opt(oo) {}

    [[noreturn]] void fail(const std::string& expected)
    {
        size_t line = 1, col = 1;
        for (size_t i = 0; i < pos && i < s.size(); ++i) {
            if (s[i] == '\n') { ++line; col = 1; } else ++col;
        }
        std::ostringstream m;
        m << "Error in Ln: " << line << " Col: " << col << ": expecting " << expected;
        throw ParseError(m.str());
    }
    bool eof() const { return pos >= s.size(); }
    char peek() const { return pos < s.size() ? s[pos] : '\0'; }
    bool isNewlineAt() const { return peek() == '\n' || peek() == '\r'; }
    bool skipNewline()
    {
        if (peek() == '\r') { ++pos; if (peek() == '\n') ++pos; return true; }
        if (peek() == '\n') { ++pos; return true; }
        return false;
    }
    void ws() { while (peek() == ' ' || peek() == '\t') ++pos; }                  // :18
    void ws1() { if (!(peek() == ' ' || peek() == '\t')) fail("space or tab"); ws(); }  // :19
    void anyWhitespace() { for (;;) { if (peek() == ' ' || peek() == '\t') ++pos; else if (!skipNewline()) break; } }  // :21-22
    bool skipTrivia()  // :24-25
    {
        if (skipNewline()) return true;
        if (peek() == ';') { while (!eof() && !isNewlineAt()) ++pos; skipNewline(); return true; }
        return false;
    }
    bool skipTrailingTrivia1() { if (!skipTrivia()) return false; while (skipTrivia()) {} return true; }  // :26
    bool tryStringCI(const char* kw)  // skipStringCI, atomic
    {
        size_t n = std::strlen(kw);
        if (pos + n > s.size()) return false;
        for (size_t i = 0; i < n; ++i)
            if (std::tolower((unsigned char)s[pos + i]) != std::tolower((unsigned char)kw[i])) return false;
        pos += n;
        return true;
    }
    bool tryKeyword(const char* kw) { if (!tryStringCI(kw)) return false; anyWhitespace(); return true; }  // pkeyword :52-53
    void expectKeyword(const char* kw) { if (!tryKeyword(kw)) fail(std::string("'") + kw + "'"); }
    void expectChar(char c) { if (peek() != c) fail(std::string("'") + c + "'"); ++pos; }
    void openBracket() { expectChar('('); anyWhitespace(); }   // inBrackets :28
    void closeBracket() { anyWhitespace(); expectChar(')'); }

    // numberLiteral with the given options (:32-50); returns false without consuming if no number starts here
    bool tryNumber(bool allowMinus, bool allowFraction, bool allowExponent, double& out, bool allowPlus = false)
    {
        size_t p = pos;
        if ((allowMinus && peek() == '-') || (allowPlus && peek() == '+')) ++p;
        size_t digits = p;
        while (p < s.size() && std::isdigit((unsigned char)s[p])) ++p;
        if (p == digits) return false;
        if (allowFraction && p < s.size() && s[p] == '.') {
            ++p;
            while (p < s.size() && std::isdigit((unsigned char)s[p])) ++p;
        }
        if (allowExponent && p < s.size() && (s[p] == 'e' || s[p] == 'E')) {
            size_t q = p + 1;
            if (q < s.size() && (s[q] == '+' || s[q] == '-')) ++q;
            size_t ed = q;
            while (q < s.size() && std::isdigit((unsigned char)s[q])) ++q;
            if (q > ed) p = q;
        }
        out = std::strtod(s.substr(pos, p - pos).c_str(), nullptr);
        pos = p;
        return true;
    }
    double pnumber() { double v; if (!tryNumber(true, true, true, v)) fail("number"); return v; }                 // :41-42
    double pnonNegativeNumber() { double v; if (!tryNumber(false, true, true, v)) fail("non-negative number"); return v; }  // :49-50
    double pfloat() { double v; if (!tryNumber(true, true, true, v, true)) fail("floating-point number"); return v; }      // FParsec pfloat
    int pint() { double v; if (!tryNumber(false, false, false, v)) fail("integer"); return (int)v; }               // :44-45
    int pint32() { double v; if (!tryNumber(true, false, false, v, true)) fail("integer number (32-bit, signed)"); return (int)v; }

    bool tryTriple(double t[3])  // ptriple :55-60 (fails without consuming if no '(')
    {
        if (peek() != '(') return false;
        openBracket();
        t[0] = pnumber(); ws();
        expectChar(','); ws(); t[1] = pnumber(); ws();
        expectChar(','); ws(); t[2] = pnumber(); ws();
        closeBracket();
        return true;
    }
    void ptriple(double t[3]) { if (!tryTriple(t)) fail("comma-separated list of 3 numbers in parens"); }
    void ppair(double t[2])  // :62-67
    {
        if (peek() != '(') fail("comma-separated list of 2 numbers in parens");
        openBracket();
        t[0] = pnumber(); ws();
        expectChar(','); ws(); t[1] = pnumber(); ws();
        closeBracket();
    }
    void pcolour(double c[3])  // :69-87
    {
        double v;
        if (tryTriple(c)) return;
        if (tryNumber(true, true, true, v)) { c[0] = c[1] = c[2] = v; return; }
        if (peek() == '#') {
            ++pos;
            if (pos + 6 > s.size()) fail("6 hex digits");
            for (int i = 0; i < 3; ++i) {
                std::string h = s.substr(pos + 2 * (size_t)i, 2);
                char* end = nullptr;
                long byte = std::strtol(h.c_str(), &end, 16);
                if (end != h.c_str() + 2) fail("hex digits");
                c[i] = (double)byte / 255.0;
            }
            pos += 6;
            return;
        }
        fail("colour");
    }
    std::string pfile()  // :89-91
    {
        expectChar('"');
        size_t st = pos;
        while (!eof() && peek() != '"') ++pos;
        std::string f = s.substr(st, pos - st);
        expectChar('"');
        return f;
    }

    // ---- textures :159-185
    int ptexture()
    {
        if (tryKeyword("grid")) {
            ftb_texture t = {};
            t.kind = FTB_TEX_GRID; t.inner = -1; t.image = -1;
            pcolour(t.p); ws1(); pcolour(t.p + 3);
            b.textures.push_back(t);
            return (int)b.textures.size() - 1;
        }
        if (tryKeyword("image")) {
            std::string file = pfile();
            std::string data, used;
            if (!resolveAsset(b, file, ".ppm", data, used)) throw ParseError("cannot open texture image '" + file + "'");
            std::vector<uint8_t> rgb;
            int w = 0, h = 0;
            if (!decodePpm(data, rgb, w, h)) throw ParseError("texture image '" + used + "' is not a binary PPM (P6, maxval 255)");
            b.imageData.push_back(std::move(rgb));
            b.images.push_back({nullptr, w, h});
            ftb_texture t = {};
            t.kind = FTB_TEX_IMAGE; t.inner = -1; t.image = (int)b.images.size() - 1;
            b.textures.push_back(t);
            return (int)b.textures.size() - 1;
        }
        if (peek() == '(') {  // appliedTextureFunction
            openBracket();
            ftb_texture t = {};
            t.image = -1;
            if (tryKeyword("scale")) {
                t.kind = FTB_TEX_SCALE;
                ppair(t.p);
            } else if (tryKeyword("rotate")) {
                t.kind = FTB_TEX_ROTATE;
                double a = degToRad(pnumber() * 1.0);  // pangle :46-47
                // matrix (rotate (Vector (0,1,0)) angle): m00 = c + invc*0*0, m02 = invc*0*0 + s*1 (Transform.fs:60-69)
                double m[12];
                matRotate(m, normalise(V3{0.0, 1.0, 0.0}), a);
                t.p[0] = a; t.p[1] = m[0]; t.p[2] = m[2];
            } else
                fail("'scale' or 'rotate'");
            anyWhitespace();
            t.inner = ptexture();
            closeBracket();
            b.textures.push_back(t);
            return (int)b.textures.size() - 1;
        }
        fail("texture");
    }

    // ---- geometry functions :155-263
    bool tryGeometryFunction(GFunc& out)
    {
        Builder* bb = &b;
        if (tryStringCI("IgnoreLight")) {  // :253
            out = [bb](int g) { return bb->addNode(FTB_NODE_IGNORELIGHT, 0, g); };
            return true;
        }
        if (tryKeyword("texture")) {  // :187-189
            int t = ptexture();
            out = [bb, t](int g) { return bb->addNode(FTB_NODE_TEXTURE, t, g); };
            return true;
        }
        if (tryKeyword("hueShift")) {  // :155-157
            pfloat();
            out = [bb](int g) { return bb->addNode(FTB_NODE_HUESHIFT, 0, g); };
            return true;
        }
        if (tryKeyword("material")) {  // :99-111, 200-202
            ftb_material m = {{1.0, 1.0, 1.0}, 0.0, 0.0, 0.0, 1, 0};
            if (tryKeyword("diffuse")) { pcolour(m.colour); ws1(); }
            if (tryKeyword("roughness")) { m.roughness = pfloat(); ws1(); }
            if (tryKeyword("reflectance")) { m.reflectance = pfloat(); ws1(); }
            if (tryKeyword("shineyness")) { m.shineyness = pfloat(); }
            b.materials.push_back(m);
            int mi = (int)b.materials.size() - 1;
            out = [bb, mi](int g) { return bb->addNode(FTB_NODE_MATERIAL, mi, g); };
            return true;
        }
        if (tryKeyword("repeat")) {  // :241-251
            int count = pint32();
            anyWhitespace();
            GFunc f;
            if (!tryGeometryFunction(f)) fail("geometry function");
            anyWhitespace();
            out = [bb, count, f](int g) {
                std::vector<int> items;
                int cur = g;
                for (int c = count;; --c) {  // factory: [f g; f (f g); ...], count+1 items
                    cur = f(cur);
                    items.push_back(cur);
                    if (c == 0) break;
                }
                int first = (int)bb->children.size();
                for (int it : items) bb->children.push_back(it);
                return bb->addNode(FTB_NODE_GROUP, first, (int)items.size());
            };
            return true;
        }
        if (tryKeyword("scale")) {  // :191-198
            double t[3];
            if (!tryTriple(t)) { double v = pnumber(); t[0] = t[1] = t[2] = v; }
            ws1();
            ftb_transform x;
            matScale(x.m2w, t[0], t[1], t[2]);
            matScale(x.w2m, 1.0 / t[0], 1.0 / t[1], 1.0 / t[2]);  // Transform.fs:49
            b.transforms.push_back(x);
            int ti = (int)b.transforms.size() - 1;
            out = [bb, ti](int g) { return bb->addNode(FTB_NODE_TRANSFORM, ti, g); };
            return true;
        }
        if (tryKeyword("translate")) {  // :215-219
            double t[3];
            ptriple(t);
            ftb_transform x;
            matTranslate(x.m2w, t[0], t[1], t[2]);
            matTranslate(x.w2m, -t[0], -t[1], -t[2]);  // Transform.fs:48
            b.transforms.push_back(x);
            int ti = (int)b.transforms.size() - 1;
            out = [bb, ti](int g) { return bb->addNode(FTB_NODE_TRANSFORM, ti, g); };
            return true;
        }
        if (tryKeyword("rotate")) {  // :204-213
            double t[3];
            ptriple(t);
            ws1();
            double angle = degToRad(pfloat() * 1.0);
            V3 axis = normalise(V3{t[0], t[1], t[2]});
            ftb_transform x;
            matRotate(x.m2w, axis, angle);
            matRotate(x.w2m, axis, -angle);  // Transform.fs:50
            b.transforms.push_back(x);
            int ti = (int)b.transforms.size() - 1;
            out = [bb, ti](int g) { return bb->addNode(FTB_NODE_TRANSFORM, ti, g); };
            return true;
        }
        if (peek() == '(') {  // composed :235-239:  (f1) . (f2)  ==  f1 >> f2
            size_t save = pos;
            openBracket();
            GFunc f1, f2;
            if (!tryGeometryFunction(f1)) { pos = save; return false; }
            closeBracket();
            anyWhitespace();
            expectChar('.');
            anyWhitespace();
            if (peek() != '(') fail("'('");
            openBracket();
            if (!tryGeometryFunction(f2)) fail("geometry function");
            closeBracket();
            out = [f1, f2](int g) { return f2(f1(g)); };
            return true;
        }
        return false;
    }

    bool tryPrimitive(int& node)  // :116-154
    {
        if (tryKeyword("mesh")) {
            std::string file = pfile(), data, used;
            if (!resolveAsset(b, file, nullptr, data, used)) throw ParseError("cannot open mesh '" + file + "'");
            Tris tris = parsePly(data);
            int first = (int)b.children.size();
            std::vector<int> items;
            for (const Tri& t : tris) items.push_back(b.addNode(FTB_NODE_PRIMITIVE, FTB_PRIM_TRIANGLE, b.addTriangle(t)));
            first = (int)b.children.size();
            for (int it : items) b.children.push_back(it);
            node = b.addNode(FTB_NODE_GROUP, first, (int)items.size());
            return true;
        }
        if (tryKeyword("bspMesh")) {
            int depth = pint();
            ws1();
            std::string file = pfile(), data, used;
            if (!resolveAsset(b, file, nullptr, data, used)) throw ParseError("cannot open mesh '" + file + "'");
            Tris tris = parsePly(data);
            int root = b.compileBsp(depth, tris);  // BspMesh.bspMesh false depth triangles (:88-97)
            b.meshes.push_back({root, 0});
            node = b.addNode(FTB_NODE_PRIMITIVE, FTB_PRIM_BSPMESH, (int)b.meshes.size() - 1);
            return true;
        }
        static const struct { const char* name; int kind; } named[] = {
            {"circle", FTB_PRIM_CIRCLE}, {"square", FTB_PRIM_SQUARE}, {"cube", FTB_PRIM_CUBE}, {"sphere", FTB_PRIM_SPHERE},
            {"plane", FTB_PRIM_PLANE}, {"cone", FTB_PRIM_CONE}, {"solidCylinder", FTB_PRIM_SOLIDCYLINDER}, {"cylinder", FTB_PRIM_CYLINDER}};
        for (const auto& np : named)
            if (tryStringCI(np.name)) { node = b.addNode(FTB_NODE_PRIMITIVE, np.kind, 0); return true; }
        return false;
    }

    bool tryGeometry(int& node)  // :264  primitive <|> inBrackets appliedFunction
    {
        if (tryPrimitive(node)) return true;
        if (peek() != '(') return false;
        openBracket();
        node = appliedFunction();
        closeBracket();
        return true;
    }
    int geometry()
    {
        int node;
        if (!tryGeometry(node)) fail("primitive or '('");
        return node;
    }
    int appliedFunction()  // :255-261
    {
        static const struct { const char* name; int kind; } ops[] = {
            {"union", FTB_NODE_UNION}, {"subtract", FTB_NODE_SUBTRACT}, {"intersect", FTB_NODE_INTERSECT}, {"exclude", FTB_NODE_EXCLUDE}};
        for (const auto& op : ops)
            if (tryKeyword(op.name)) {  // binaryGeometryFunction :221-226
                int a = geometry();
                ws1();
                int bnode = geometry();
                return b.addNode(op.kind, a, bnode);
            }
        if (tryKeyword("group")) {  // :228-231
            std::vector<int> items;
            int node;
            while (tryGeometry(node)) {  // many (geometry .>> anyWhitespace)
                items.push_back(node);
                anyWhitespace();
            }
            int first = (int)b.children.size();
            for (int it : items) b.children.push_back(it);
            return b.addNode(FTB_NODE_GROUP, first, (int)items.size());
        }
        GFunc f;
        if (!tryGeometryFunction(f)) fail("geometry function, 'group' or CSG operator");
        anyWhitespace();
        int g = geometry();
        return f(g);
    }

    // ---- options :272-317
    bool tryOption()
    {
        if (tryKeyword("camera")) {
            ftb_camera c = {};
            double t[3], pr[2];
            expectKeyword("pos"); ptriple(c.o); ws1();
            expectKeyword("lookat"); ptriple(c.look_at); ws1();
            expectKeyword("up"); ptriple(t); ws1();
            V3 up = normalise(V3{t[0], t[1], t[2]});
            c.up[0] = up.x; c.up[1] = up.y; c.up[2] = up.z;
            expectKeyword("fov"); c.fov_y_rad = degToRad(1.0 * pnonNegativeNumber()); ws1();
            expectKeyword("ratio"); c.aspect_ratio = pnonNegativeNumber(); ws();
            if (tryKeyword("focus")) {
                ppair(pr);
                c.has_focus = 1; c.focal_length = pr[0]; c.aperture_rad = degToRad(pr[1] * 1.0);
            }
            opt.camera = c;
            return true;
        }
        if (tryKeyword("samples")) {
            double v;
            if (tryNumber(true, false, false, v, true)) { opt.sampling = FTB_SAMPLING_JITTER; opt.spp = (int)v; return true; }
            if (s.compare(pos, 6, "corner") == 0) { pos += 6; opt.sampling = FTB_SAMPLING_CORNER; return true; }
            fail("positive number or 'corner'");
        }
        if (tryKeyword("res")) {
            opt.width = pint32(); ws1(); opt.height = pint32();
            return true;
        }
        return false;
    }

    // ---- lights :319-351
    bool tryLight()
    {
        ftb_light L = {};
        double t[3];
        if (tryKeyword("directional")) {
            expectKeyword("dir"); ptriple(t); ws1();
            expectKeyword("colour"); pcolour(L.colour);
            V3 d = normalise(V3{t[0], t[1], t[2]});  // Light.fs:19-20
            L.kind = FTB_LIGHT_DIRECTIONAL; L.v[0] = d.x; L.v[1] = d.y; L.v[2] = d.z;
            b.lights.push_back(L);
            return true;
        }
        if (tryKeyword("softdirectional")) {
            expectKeyword("dir"); ptriple(t); ws1();
            expectKeyword("samples"); L.samples = pint32(); ws1();
            expectKeyword("scatter"); double sc = pnonNegativeNumber(); ws1();
            expectKeyword("colour"); ptriple(L.colour);
            V3 d = normalise(V3{t[0], t[1], t[2]});  // Light.fs:22-23
            L.kind = FTB_LIGHT_SOFT_DIRECTIONAL; L.v[0] = d.x; L.v[1] = d.y; L.v[2] = d.z;
            L.scatter_rad = degToRad(1.0) * sc;  // SceneParser.fs:329
            b.lights.push_back(L);
            return true;
        }
        if (tryKeyword("positional")) {
            expectKeyword("pos"); ptriple(L.v); ws1();
            expectKeyword("falloff"); ptriple(L.falloff); ws1();
            expectKeyword("colour"); ptriple(L.colour);
            L.kind = FTB_LIGHT_POINT;
            b.lights.push_back(L);
            return true;
        }
        return false;
    }

    // pscenegraph :353-358
    int parseAll()
    {
        while (skipTrivia()) {}
        // poptions = sepEndBy (option .>> ws) skipTrailingTrivia1
        while (tryOption()) { ws(); if (!skipTrailingTrivia1()) break; }
        std::vector<int> objects;
        int node;
        while (tryGeometry(node)) {  // pobjects = sepEndBy (geometry .>> ws) skipTrailingTrivia1
            objects.push_back(node);
            ws();
            if (!skipTrailingTrivia1()) break;
        }
        while (tryLight()) { ws(); if (!skipTrailingTrivia1()) break; }
        if (!eof()) fail("end of input");
        int first = (int)b.children.size();
        for (int it : objects) b.children.push_back(it);
        return b.addNode(FTB_NODE_GROUP, first, (int)objects.size());
    }
};

}  // namespace

struct ftbf_scene {
    Builder b;
    Options opt;
    ftb_scene_desc desc;
};

extern "C" {

const char* ftbf_last_error(void) { return g_err.c_str(); }

int ftbf_parse(const char* text, const char* asset_dir, ftbf_scene** out)
{
    if (!text || !out) { g_err = "null argument"; return FTB_ERR_BAD_ARG; }
    ftbf_scene* sc = new ftbf_scene();
    sc->b.assetDir = asset_dir ? asset_dir : "";
    // SceneOptions.Default (Scene.fs:61-65)
    ftb_camera c = {};
    c.look_at[2] = 1.0; c.up[1] = 1.0; c.fov_y_rad = degToRad(50.0); c.aspect_ratio = 1.0;
    sc->opt = {c, 400, 400, 8, FTB_SAMPLING_JITTER};
    std::string src(text);
    try {
        Parser p(src, sc->b, sc->opt);
        int root = p.parseAll();
        Builder& b = sc->b;
        for (size_t i = 0; i < b.images.size(); ++i) b.images[i].rgb24 = b.imageData[i].data();
        ftb_scene_desc& d = sc->desc;
        std::memset(&d, 0, sizeof(d));
        d.root = root;
        d.n_nodes = (int)b.nodes.size(); d.nodes = b.nodes.data();
        d.n_children = (int)b.children.size(); d.children = b.children.data();
        d.n_transforms = (int)b.transforms.size(); d.transforms = b.transforms.data();
        d.n_materials = (int)b.materials.size(); d.materials = b.materials.data();
        d.n_textures = (int)b.textures.size(); d.textures = b.textures.data();
        d.n_images = (int)b.images.size(); d.images = b.images.data();
        d.n_meshes = (int)b.meshes.size(); d.meshes = b.meshes.data();
        d.n_bsp_nodes = (int)b.bspNodes.size(); d.bsp_nodes = b.bspNodes.data();
        d.n_bsp_leaves = (int)b.bspLeaves.size(); d.bsp_leaves = b.bspLeaves.data();
        d.n_triangles = (int)(b.triangles.size() / 9); d.triangles = b.triangles.data();
        d.n_lights = (int)b.lights.size(); d.lights = b.lights.data();
    } catch (const std::exception& e) {
        g_err = e.what();
        delete sc;
        *out = nullptr;
        return FTB_ERR_BAD_SCENE;
    }
    *out = sc;
    return FTB_OK;
}

void ftbf_destroy(ftbf_scene* s) { delete s; }
const ftb_scene_desc* ftbf_desc(const ftbf_scene* s) { return s ? &s->desc : nullptr; }
const ftb_camera* ftbf_camera(const ftbf_scene* s) { return s ? &s->opt.camera : nullptr; }
void ftbf_options(const ftbf_scene* s, int* width, int* height, int* spp, int* sampling)
{
    if (!s) return;
    if (width) *width = s->opt.width;
    if (height) *height = s->opt.height;
    if (spp) *spp = s->opt.spp;
    if (sampling) *sampling = s->opt.sampling;
}

void ftbf_jitter_pattern(uint64_t seed, int spp, double* xy)
{
    uint64_t state = seed * 0x9E3779B97F4A7C15ULL + 0x2545F4914F6CDD1DULL;
    auto next = [&]() {  // splitmix64 -> [0,1) with 53 bits, stands in for Random.NextDouble()
        uint64_t z = (state += 0x9E3779B97F4A7C15ULL);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        z ^= z >> 31;
        return (double)(z >> 11) * (1.0 / 9007199254740992.0);
    };
    for (int i = 0; i < spp; ++i)
        for (;;) {  // Jitter.circle (Jitter.fs:15-21)
            double x = 2.0 * next() - 1.0, y = 2.0 * next() - 1.0;
            if ((x * x + y * y) > 1.0) continue;
            xy[2 * i] = x;
            xy[2 * i + 1] = y;
            break;
        }
}

void ftbf_slice_triangle(const double* p0, const double* n, const double* tri, double* above, int* n_above, double* below, int* n_below)
{
    Tris a, bl;
    slice(V3{p0[0], p0[1], p0[2]}, V3{n[0], n[1], n[2]}, Tri{{tri[0], tri[1], tri[2]}, {tri[3], tri[4], tri[5]}, {tri[6], tri[7], tri[8]}}, a, bl);
    auto put = [](const Tris& ts, double* out, int* cnt) {
        *cnt = (int)ts.size();
        for (size_t i = 0; i < ts.size(); ++i) {
            const double v[9] = {ts[i].a.x, ts[i].a.y, ts[i].a.z, ts[i].b.x, ts[i].b.y, ts[i].b.z, ts[i].c.x, ts[i].c.y, ts[i].c.z};
            std::memcpy(out + 9 * i, v, sizeof(v));
        }
    };
    put(a, above, n_above);
    put(bl, below, n_below);
}

int ftbf_parse_colour(const char* text, double* rgb)
{
    Builder b;
    Options o = {};
    std::string src(text ? text : "");
    try {
        Parser p(src, b, o);
        p.pcolour(rgb);
    } catch (const std::exception& e) {
        g_err = e.what();
        return FTB_ERR_BAD_SCENE;
    }
    return FTB_OK;
}

int ftbf_write_png(const char* path, int width, int height, const uint8_t* rgba)
{
    if (!path || !rgba || width <= 0 || height <= 0) { g_err = "bad png arguments"; return FTB_ERR_BAD_ARG; }
    std::vector<uint8_t> raw((size_t)height * ((size_t)width * 4 + 1));
    for (int y = 0; y < height; ++y) {
        raw[(size_t)y * ((size_t)width * 4 + 1)] = 0;  // filter: none
        std::memcpy(&raw[(size_t)y * ((size_t)width * 4 + 1) + 1], rgba + (size_t)y * width * 4, (size_t)width * 4);
    }
    uLongf clen = compressBound((uLong)raw.size());
    std::vector<uint8_t> comp(clen);
    if (compress2(comp.data(), &clen, raw.data(), (uLong)raw.size(), 6) != Z_OK) { g_err = "zlib failure"; return FTB_ERR_OOM; }
    FILE* f = std::strcmp(path, "-") == 0 ? stdout : std::fopen(path, "wb");
    if (!f) { g_err = std::string("cannot open ") + path; return FTB_ERR_BAD_ARG; }
    auto be32 = [](uint8_t* p, uint32_t v) { p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v; };
    auto chunk = [&](const char* type, const uint8_t* data, uint32_t len) {
        uint8_t hdr[8];
        be32(hdr, len);
        std::memcpy(hdr + 4, type, 4);
        std::fwrite(hdr, 1, 8, f);
        if (len) std::fwrite(data, 1, len, f);
        uLong crc = crc32(0L, (const Bytef*)type, 4);
        if (len) crc = crc32(crc, data, len);
        uint8_t c[4];
        be32(c, (uint32_t)crc);
        std::fwrite(c, 1, 4, f);
    };
    const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    std::fwrite(sig, 1, 8, f);
    uint8_t ihdr[13];
    be32(ihdr, (uint32_t)width); be32(ihdr + 4, (uint32_t)height);
    ihdr[8] = 8; ihdr[9] = 6; ihdr[10] = 0; ihdr[11] = 0; ihdr[12] = 0;  // 8-bit RGBA
    chunk("IHDR", ihdr, 13);
    chunk("IDAT", comp.data(), (uint32_t)clen);
    chunk("IEND", nullptr, 0);
    if (f != stdout) std::fclose(f); else std::fflush(f);
    return FTB_OK;
}
}
