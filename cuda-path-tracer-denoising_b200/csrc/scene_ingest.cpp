// scene_ingest.cpp -- SURVEY.md 8(f) N2: the reference's scene description -> the arrays svgf_create() takes.
//
// Reads the reference's text scene format (MATERIAL / OBJECT / CAMERA blocks, src/scene.cpp:9-238) and Wavefront OBJ
// meshes, transforms the triangles to world space (scene.cpp:240-311) and builds the SAH BVH (src/bvhtree.cpp:21-182), so
// that a caller no longer needs the reference's Scene loader in front of the hot path. The on-disk contract stays the
// reference's; the output is bit-for-bit what its loader produces (tests/test_scene_ingest.py compares with the arrays
// exported from the reference's own loader), which matters because closest-hit ties and the 64-deep traversal stack make
// the rendered image depend on triangle order and tree shape. That fixes the arithmetic:
//   * matrices in glm 0.9.6.3's expression order (gtc/matrix_transform.inl translate/rotate/scale, detail/type_mat4x4.inl
//     operator*, compute_inverse; gtc/matrix_inverse.inl inverseTranspose), fp32, no FMA contraction (-ffp-contract=off);
//   * OBJ numbers through the decimal parser of the tinyobjloader version the reference vendors (mantissa accumulated digit
//     by digit in double, tiny_obj_loader.cc:119-233) -- not strtod, which rounds differently in the last bit;
//   * the BVH builder's quirks: an all-zero box is "empty" for unions, 9 SAH buckets, <= 10 triangles per leaf,
//     std::partition / std::nth_element for the splits, right subtree built before the left one (the order g++ evaluates
//     the two recursive calls in MakeNode's argument list, bvhtree.cpp:87,133), pre-order flattening.
// Textures are reported by file name; their pixels are either attached by the caller (svgf_scene_set_texture) or decoded here
// from the JPEG files (svgf_scene_load_textures -> csrc/jpeg_decode.cpp).
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/svgf_b200.h"

// csrc/jpeg_decode.cpp
bool svgf_jpeg_decode(const unsigned char *bytes, size_t n, int *width, int *height, int *components, std::vector<unsigned char> &pixels, std::string &err);

void svgf_mat4_inverse(const float *m, float *out16);      // camera.cpp (glm compute_inverse)

namespace {
const float kPi = 3.1415926535897932384626422832795028841971f;     // utilities.h:12

struct V3 { float x, y, z; };
struct V4 { float x, y, z, w; };
struct M4 { V4 c[4]; };                                             // column-major, like glm::mat4

inline V4 mulvs(V4 a, float s) { return V4{a.x * s, a.y * s, a.z * s, a.w * s}; }
inline V4 addv(V4 a, V4 b) { return V4{a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
inline V4 mulvv(V4 a, V4 b) { return V4{a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w}; }
inline M4 identity() { return M4{{{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}}}; }

// m1 * m2: every column is ((A0*b0 + A1*b1) + A2*b2) + A3*b3 (type_mat4x4.inl:686-706)
M4 mul(const M4 &a, const M4 &b) {
    M4 r;
    for (int i = 0; i < 4; i++)
        r.c[i] = addv(addv(addv(mulvs(a.c[0], b.c[i].x), mulvs(a.c[1], b.c[i].y)), mulvs(a.c[2], b.c[i].z)), mulvs(a.c[3], b.c[i].w));
    return r;
}
// m * v = (m0*v0 + m1*v1) + (m2*v2 + m3*v3) (type_mat4x4.inl:617-628)
V4 mulmv(const M4 &m, V4 v) {
    return addv(addv(mulvs(m.c[0], v.x), mulvs(m.c[1], v.y)), addv(mulvs(m.c[2], v.z), mulvs(m.c[3], v.w)));
}
M4 translate(const M4 &m, V3 v) {       // gtc/matrix_transform.inl: Result[3] = m[0]*v0 + m[1]*v1 + m[2]*v2 + m[3]
    M4 r = m;
    r.c[3] = addv(addv(addv(mulvs(m.c[0], v.x), mulvs(m.c[1], v.y)), mulvs(m.c[2], v.z)), m.c[3]);
    return r;
}
M4 rotate(const M4 &m, float angle, V3 v) {
    const float c = cosf(angle), s = sinf(angle);
    const float inv = 1.0f / sqrtf(v.x * v.x + v.y * v.y + v.z * v.z);       // normalize = v * inversesqrt(dot(v, v))
    const V3 ax = {v.x * inv, v.y * inv, v.z * inv};
    const float k = 1.0f - c;
    const V3 t = {k * ax.x, k * ax.y, k * ax.z};
    float R[3][3];
    R[0][0] = c + t.x * ax.x;           R[0][1] = 0 + t.x * ax.y + s * ax.z;  R[0][2] = 0 + t.x * ax.z - s * ax.y;
    R[1][0] = 0 + t.y * ax.x - s * ax.z; R[1][1] = c + t.y * ax.y;           R[1][2] = 0 + t.y * ax.z + s * ax.x;
    R[2][0] = 0 + t.z * ax.x + s * ax.y; R[2][1] = 0 + t.z * ax.y - s * ax.x; R[2][2] = c + t.z * ax.z;
    M4 r;
    for (int i = 0; i < 3; i++) r.c[i] = addv(addv(mulvs(m.c[0], R[i][0]), mulvs(m.c[1], R[i][1])), mulvs(m.c[2], R[i][2]));
    r.c[3] = m.c[3];
    return r;
}
M4 scale(const M4 &m, V3 v) { return M4{{mulvs(m.c[0], v.x), mulvs(m.c[1], v.y), mulvs(m.c[2], v.z), m.c[3]}}; }

// utilityCore::buildTransformationMatrix (src/utilities.cpp:63-71)
M4 build_transform(V3 t, V3 r, V3 s) {
    const M4 tm = translate(identity(), t);
    M4 rm = rotate(identity(), r.x * kPi / 180, V3{1, 0, 0});
    rm = mul(rm, rotate(identity(), r.y * kPi / 180, V3{0, 1, 0}));
    rm = mul(rm, rotate(identity(), r.z * kPi / 180, V3{0, 0, 1}));
    return mul(mul(tm, rm), scale(identity(), s));
}
// glm::inverseTranspose(mat4) (gtc/matrix_inverse.inl:95-148): cofactors, then every element divided by the determinant
M4 inverse_transpose(const M4 &mm) {
    float m[4][4];
    for (int i = 0; i < 4; i++) { m[i][0] = mm.c[i].x; m[i][1] = mm.c[i].y; m[i][2] = mm.c[i].z; m[i][3] = mm.c[i].w; }
    const float s00 = m[2][2] * m[3][3] - m[3][2] * m[2][3], s01 = m[2][1] * m[3][3] - m[3][1] * m[2][3];
    const float s02 = m[2][1] * m[3][2] - m[3][1] * m[2][2], s03 = m[2][0] * m[3][3] - m[3][0] * m[2][3];
    const float s04 = m[2][0] * m[3][2] - m[3][0] * m[2][2], s05 = m[2][0] * m[3][1] - m[3][0] * m[2][1];
    const float s06 = m[1][2] * m[3][3] - m[3][2] * m[1][3], s07 = m[1][1] * m[3][3] - m[3][1] * m[1][3];
    const float s08 = m[1][1] * m[3][2] - m[3][1] * m[1][2], s09 = m[1][0] * m[3][3] - m[3][0] * m[1][3];
    const float s10 = m[1][0] * m[3][2] - m[3][0] * m[1][2], s11 = m[1][1] * m[3][3] - m[3][1] * m[1][3];
    const float s12 = m[1][0] * m[3][1] - m[3][0] * m[1][1], s13 = m[1][2] * m[2][3] - m[2][2] * m[1][3];
    const float s14 = m[1][1] * m[2][3] - m[2][1] * m[1][3], s15 = m[1][1] * m[2][2] - m[2][1] * m[1][2];
    const float s16 = m[1][0] * m[2][3] - m[2][0] * m[1][3], s17 = m[1][0] * m[2][2] - m[2][0] * m[1][2];
    const float s18 = m[1][0] * m[2][1] - m[2][0] * m[1][1];
    float o[4][4];
    o[0][0] = +((m[1][1] * s00 - m[1][2] * s01) + m[1][3] * s02);
    o[0][1] = -((m[1][0] * s00 - m[1][2] * s03) + m[1][3] * s04);
    o[0][2] = +((m[1][0] * s01 - m[1][1] * s03) + m[1][3] * s05);
    o[0][3] = -((m[1][0] * s02 - m[1][1] * s04) + m[1][2] * s05);
    o[1][0] = -((m[0][1] * s00 - m[0][2] * s01) + m[0][3] * s02);
    o[1][1] = +((m[0][0] * s00 - m[0][2] * s03) + m[0][3] * s04);
    o[1][2] = -((m[0][0] * s01 - m[0][1] * s03) + m[0][3] * s05);
    o[1][3] = +((m[0][0] * s02 - m[0][1] * s04) + m[0][2] * s05);
    o[2][0] = +((m[0][1] * s06 - m[0][2] * s07) + m[0][3] * s08);
    o[2][1] = -((m[0][0] * s06 - m[0][2] * s09) + m[0][3] * s10);
    o[2][2] = +((m[0][0] * s11 - m[0][1] * s09) + m[0][3] * s12);
    o[2][3] = -((m[0][0] * s08 - m[0][1] * s10) + m[0][2] * s12);
    o[3][0] = -((m[0][1] * s13 - m[0][2] * s14) + m[0][3] * s15);
    o[3][1] = +((m[0][0] * s13 - m[0][2] * s16) + m[0][3] * s17);
    o[3][2] = -((m[0][0] * s14 - m[0][1] * s16) + m[0][3] * s18);
    o[3][3] = +((m[0][0] * s15 - m[0][1] * s17) + m[0][2] * s18);
    const float det = ((+m[0][0] * o[0][0] + m[0][1] * o[0][1]) + m[0][2] * o[0][2]) + m[0][3] * o[0][3];
    M4 r;
    for (int i = 0; i < 4; i++) r.c[i] = V4{o[i][0] / det, o[i][1] / det, o[i][2] / det, o[i][3] / det};
    return r;
}
inline void store(float *dst, const M4 &m) { memcpy(dst, &m, 64); }

// ---- text helpers (utilityCore::safeGetline / tokenizeString, src/utilities.cpp:73-110) ----
void get_line(std::istream &is, std::string &t) {      // handles \n, \r\n, \r and a last line without terminator
    t.clear();
    if (!is.good()) return;
    for (;;) {
        const int ch = is.get();        // at end of file this also sets eofbit, which ends the callers' `while (in.good())`
        if (ch == '\n' || ch == EOF) return;
        if (ch == '\r') { if (is.peek() == '\n') is.get(); return; }
        t += (char)ch;
    }
}
std::vector<std::string> tokens_of(const std::string &s) {
    std::istringstream ss(s);
    std::vector<std::string> out;
    std::string w;
    while (ss >> w) out.push_back(w);
    return out;
}
inline float num(const std::vector<std::string> &t, size_t i) { return i < t.size() ? (float)atof(t[i].c_str()) : 0.0f; }

// ---- OBJ numbers: the vendored tinyobjloader's tryParseDouble (tiny_obj_loader.cc:119-233) ----
bool parse_decimal(const char *s, const char *s_end, double *result) {
    if (s >= s_end) return false;
    double mantissa = 0.0;
    int exponent = 0;
    char sign = '+', exp_sign = '+';
    const char *curr = s;
    int read = 0;
    bool more = false;
    if (*curr == '+' || *curr == '-') { sign = *curr; curr++; }
    else if (!isdigit((unsigned char)*curr)) return false;
    while ((more = (curr != s_end)) && isdigit((unsigned char)*curr)) { mantissa *= 10; mantissa += (int)(*curr - '0'); curr++; read++; }
    if (read == 0) return false;
    bool assemble = !more;
    if (!assemble) {
        if (*curr == '.') {
            curr++; read = 1;
            while ((more = (curr != s_end)) && isdigit((unsigned char)*curr)) { mantissa += (int)(*curr - '0') * pow(10, -read); read++; curr++; }
        } else if (*curr != 'e' && *curr != 'E') {
            assemble = true;
        }
        if (!assemble && more && (*curr == 'e' || *curr == 'E')) {
            curr++;
            if ((more = (curr != s_end)) && (*curr == '+' || *curr == '-')) { exp_sign = *curr; curr++; }
            else if (!isdigit((unsigned char)*curr)) return false;
            read = 0;
            while ((more = (curr != s_end)) && isdigit((unsigned char)*curr)) { exponent *= 10; exponent += (int)(*curr - '0'); curr++; read++; }
            exponent *= (exp_sign == '+' ? 1 : -1);
            if (read == 0) return false;
        }
    }
    *result = (sign == '+' ? 1 : -1) * ldexp(mantissa * pow(5, exponent), exponent);
    return true;
}
float obj_float(const char *&tok) {
    tok += strspn(tok, " \t");
    const char *end = tok + strcspn(tok, " \t\r");
    double v = 0.0;
    parse_decimal(tok, end, &v);
    tok = end;
    return (float)v;
}
inline int fix_index(int idx, int n) { return idx > 0 ? idx - 1 : (idx == 0 ? 0 : n + idx); }
struct Corner { int v, vt, vn; };
Corner parse_corner(const char *&tok, int nv, int nvn, int nvt) {       // i, i/j, i//k, i/j/k
    Corner c{-1, -1, -1};
    c.v = fix_index(atoi(tok), nv);
    tok += strcspn(tok, "/ \t\r");
    if (tok[0] != '/') return c;
    tok++;
    if (tok[0] == '/') { tok++; c.vn = fix_index(atoi(tok), nvn); tok += strcspn(tok, "/ \t\r"); return c; }
    c.vt = fix_index(atoi(tok), nvt);
    tok += strcspn(tok, "/ \t\r");
    if (tok[0] != '/') return c;
    tok++;
    c.vn = fix_index(atoi(tok), nvn);
    tok += strcspn(tok, "/ \t\r");
    return c;
}

// ---- BVH (src/bvhtree.cpp, src/boundingbox.{h,cpp}) ----
struct Box { V3 mn, mx; };
inline float fmin_glm(float a, float b) { return b < a ? b : a; }      // glm::min / glm::max
inline float fmax_glm(float a, float b) { return a < b ? b : a; }
inline bool is_zero(const Box &b) { return b.mn.x == 0.f && b.mx.x == 0.f && b.mn.y == 0.f && b.mx.y == 0.f && b.mn.z == 0.f && b.mx.z == 0.f; }
inline Box unite(const Box &a, const Box &b) {         // BoundingBox::operator||(BoundingBox&): an all-zero `a` means "empty"
    if (is_zero(a)) return b;
    return Box{{fmin_glm(a.mn.x, b.mn.x), fmin_glm(a.mn.y, b.mn.y), fmin_glm(a.mn.z, b.mn.z)},
               {fmax_glm(a.mx.x, b.mx.x), fmax_glm(a.mx.y, b.mx.y), fmax_glm(a.mx.z, b.mx.z)}};
}
inline Box unite_pt(const Box &a, V3 p) {              // operator||(vec3): plain min/max
    return Box{{fmin_glm(a.mn.x, p.x), fmin_glm(a.mn.y, p.y), fmin_glm(a.mn.z, p.z)}, {fmax_glm(a.mx.x, p.x), fmax_glm(a.mx.y, p.y), fmax_glm(a.mx.z, p.z)}};
}
inline float area(const Box &b) {
    const float dx = b.mx.x - b.mn.x, dy = b.mx.y - b.mn.y, dz = b.mx.z - b.mn.z;
    return 2.0f * ((dx * dy + dx * dz) + dy * dz);
}
inline float comp(V3 v, int a) { return a == 0 ? v.x : (a == 1 ? v.y : v.z); }
inline int longest_axis(const Box &b) {
    const float dx = b.mx.x - b.mn.x, dy = b.mx.y - b.mn.y, dz = b.mx.z - b.mn.z;
    if (dx > dy && dx > dz) return 0;
    return dy > dz ? 1 : 2;
}
inline float offset_on(const Box &b, V3 p, int a) {   // BoundingBox::getOffset(p)[a]
    float o = comp(p, a) - comp(b.mn, a);
    if (comp(b.mx, a) > comp(b.mn, a)) o /= (comp(b.mx, a) - comp(b.mn, a));
    return o;
}
struct Prim { int index; Box bounds; V3 centroid; };
struct Node { Box box; Node *l = nullptr, *r = nullptr; int axis = -1, first = 0, count = 0; };

struct Builder {
    std::vector<Prim> prims;
    const std::vector<svgf_triangle> *tris = nullptr;
    std::vector<svgf_triangle> ordered;
    int total = 0;
    static int cmp_axis;

    Node *leaf(Node *n, int start, int end, const Box &bounds) {
        n->first = (int)ordered.size();
        for (int i = start; i < end; i++) ordered.push_back((*tris)[prims[i].index]);
        n->count = end - start; n->box = bounds; n->axis = -1;
        return n;
    }
    Node *interior(Node *n, int axis, int start, int mid, int end) {
        Node *r = build(mid, end);          // the right subtree first, see the header of this file
        Node *l = build(start, mid);
        n->l = l; n->r = r; n->box = unite(l->box, r->box); n->axis = axis; n->count = 0;
        return n;
    }
    Node *build(int start, int end) {
        Node *n = new Node();
        total++;
        Box bounds = prims[start].bounds;
        for (int i = start; i < end; i++) bounds = unite(bounds, prims[i].bounds);
        const int ntris = end - start;
        if (ntris == 1) return leaf(n, start, end, bounds);
        Box cb{prims[start].centroid, prims[start].centroid};
        for (int i = start; i < end; i++) cb = unite_pt(cb, prims[i].centroid);
        const int ax = longest_axis(cb);
        if (comp(cb.mx, ax) == comp(cb.mn, ax)) return leaf(n, start, end, bounds);
        if (ntris == 2) {
            const float midf = (1.f * (start + end)) / 2.f;
            const int mid = (int)midf;
            cmp_axis = ax;
            std::nth_element(&prims[start], &prims[mid], &prims[end - 1] + 1,
                             [](const Prim &a, const Prim &b) { return comp(a.centroid, cmp_axis) < comp(b.centroid, cmp_axis); });
            return interior(n, ax, start, mid, end);
        }
        const int NB = 9;
        struct Bucket { int count = 0; Box bounds{{0, 0, 0}, {0, 0, 0}}; } bucket[NB];
        for (int i = start; i < end; i++) {
            int b = (int)(NB * offset_on(cb, prims[i].centroid, ax));
            if (b >= NB) b = NB - 1;
            if (b < 0) b = 0;
            bucket[b].bounds = unite(bucket[b].bounds, prims[i].bounds);
            bucket[b].count++;
        }
        float cost[NB - 1];
        for (int i = 0; i < NB - 1; i++) {
            int ca = 0, cbn = 0;
            Box A{{0, 0, 0}, {0, 0, 0}}, B{{0, 0, 0}, {0, 0, 0}};
            for (int j = 0; j <= i; j++) { A = unite(A, bucket[j].bounds); ca += bucket[j].count; }
            for (int j = i + 1; j < NB; j++) { B = unite(B, bucket[j].bounds); cbn += bucket[j].count; }
            cost[i] = 1.f + (ca * area(A) + cbn * area(B)) / area(bounds);
        }
        float min_cost = FLT_MAX;
        int split = 0;
        for (int i = 0; i < NB - 1; i++) if (cost[i] < min_cost) { min_cost = cost[i]; split = i; }
        if (min_cost < ntris || ntris > 10) {
            Prim *midp = std::partition(&prims[start], &prims[end - 1] + 1, [=](const Prim &p) {
                int b = (int)(NB * offset_on(cb, p.centroid, ax));
                if (b == NB) b = NB - 1;
                return b <= split;
            });
            const int mid = (int)(midp - &prims[0]);
            if (mid == start || mid == end) return leaf(n, start, end, bounds);   // cannot happen with finite extents; never recurse on it
            return interior(n, ax, start, mid, end);
        }
        return leaf(n, start, end, bounds);
    }
    int flatten(Node *n, std::vector<svgf_bvh_node> &out, int &offset) {
        svgf_bvh_node &o = out[offset];
        memcpy(o.bounds_min, &n->box.mn, 12); memcpy(o.bounds_max, &n->box.mx, 12);
        const int me = offset++;
        if (n->count > 0) { o.primitivesOffset = n->first; o.primitive_count = n->count; return me; }
        o.primitive_count = 0; o.axis = n->axis;
        flatten(n->l, out, offset);
        const int right = flatten(n->r, out, offset);
        out[me].rightchildoffset = right;
        return me;
    }
    static void destroy(Node *n) { if (!n) return; destroy(n->l); destroy(n->r); delete n; }
};
int Builder::cmp_axis = 0;
}  // namespace

struct svgf_scene {
    std::vector<svgf_geom> geoms;
    std::vector<svgf_material> materials;
    std::vector<svgf_triangle> triangles;          // BVH order after load
    std::vector<svgf_bvh_node> bvh;
    std::vector<Box> mesh_boxes;                    // Scene::BoudningBoxs (one per MESH geom; not used by the hot path)
    std::vector<std::string> texture_files;
    struct Tex { int w = 0, h = 0, comp = 0; std::vector<unsigned char> px; };
    std::vector<Tex> textures;
    std::vector<svgf_texture_desc> tex_desc;
    float eye[3] = {0, 0, 0}, lookat[3] = {0, 0, 0}, up[3] = {0, 0, 0}, fovy = 0.f;
    int res[2] = {0, 0};
    int tri_counter = 0;
    std::string err;
};

namespace {
bool load_obj(svgf_scene &sc, const std::string &path, svgf_geom &g, const M4 &xf, const M4 &it) {
    std::ifstream in(path.c_str());
    if (!in) { sc.err = "cannot open OBJ file " + path; return false; }
    std::vector<float> v, vn, vt;
    g.BoundIdx = (int)sc.mesh_boxes.size();
    g.T_startidx = (int)sc.triangles.size();
    float mx[3] = {FLT_MIN, FLT_MIN, FLT_MIN}, mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX};     // sic: FLT_MIN, scene.cpp:255
    std::string line;
    std::vector<Corner> face;
    while (std::getline(in, line)) {
        if (!line.empty() && line[line.size() - 1] == '\r') line.erase(line.size() - 1);
        const char *tok = line.c_str();
        tok += strspn(tok, " \t");
        if (tok[0] == '\0' || tok[0] == '#') continue;
        const bool sp1 = tok[1] == ' ' || tok[1] == '\t', sp2 = tok[1] != '\0' && (tok[2] == ' ' || tok[2] == '\t');
        if (tok[0] == 'v' && sp1) { tok += 2; for (int k = 0; k < 3; k++) v.push_back(obj_float(tok)); continue; }
        if (tok[0] == 'v' && tok[1] == 'n' && sp2) { tok += 3; for (int k = 0; k < 3; k++) vn.push_back(obj_float(tok)); continue; }
        if (tok[0] == 'v' && tok[1] == 't' && sp2) { tok += 3; for (int k = 0; k < 2; k++) vt.push_back(obj_float(tok)); continue; }
        if (!(tok[0] == 'f' && sp1)) continue;     // groups, materials, smoothing: no effect on the triangle stream
        tok += 2;
        tok += strspn(tok, " \t");
        face.clear();
        while (!(tok[0] == '\r' || tok[0] == '\n' || tok[0] == '\0')) {
            face.push_back(parse_corner(tok, (int)(v.size() / 3), (int)(vn.size() / 3), (int)(vt.size() / 2)));
            tok += strspn(tok, " \t\r");
        }
        for (size_t k = 2; k < face.size(); k++) {         // polygon -> triangle fan (tiny_obj_loader.cc:375-397)
            const Corner cs[3] = {face[0], face[k - 1], face[k]};
            svgf_triangle t;
            memset(&t, 0, sizeof(t));
            V3 wp[3];
            for (int q = 0; q < 3; q++) {
                if (cs[q].v < 0 || (size_t)(3 * cs[q].v + 2) >= v.size()) { sc.err = "OBJ face references a missing vertex in " + path; return false; }
                const V4 p = mulmv(xf, V4{v[3 * cs[q].v], v[3 * cs[q].v + 1], v[3 * cs[q].v + 2], 1.0f});
                wp[q] = V3{p.x, p.y, p.z};
                // the BVH builder (and the tracer) need ordered coordinates with finite extents
                if (!(fabsf(p.x) <= 1e18f && fabsf(p.y) <= 1e18f && fabsf(p.z) <= 1e18f)) {
                    sc.err = "non-finite or out-of-range vertex after the object's transform in " + path;
                    return false;
                }
                memcpy(t.verts[q].pos, &wp[q], 12);
            }
            for (int a = 0; a < 3; a++) {                   // utilityCore::compareThreeVertex + running totals (scene.cpp:273-281)
                const float p0 = comp(wp[0], a), p1 = comp(wp[1], a), p2 = comp(wp[2], a);
                mn[a] = fmin_glm(mn[a], fmin_glm(p0, fmin_glm(p1, p2)));
                mx[a] = fmax_glm(mx[a], fmax_glm(p0, fmax_glm(p1, p2)));
            }
            if (!vn.empty())
                for (int q = 0; q < 3; q++) {
                    if (cs[q].vn < 0 || (size_t)(3 * cs[q].vn + 2) >= vn.size()) continue;
                    const V4 n = mulmv(it, V4{vn[3 * cs[q].vn], vn[3 * cs[q].vn + 1], vn[3 * cs[q].vn + 2], 0.0f});
                    t.verts[q].normal[0] = n.x; t.verts[q].normal[1] = n.y; t.verts[q].normal[2] = n.z;
                }
            if (!vt.empty())
                for (int q = 0; q < 3; q++) {
                    if (cs[q].vt < 0 || (size_t)(2 * cs[q].vt + 1) >= vt.size()) continue;
                    t.verts[q].uv[0] = vt[2 * cs[q].vt]; t.verts[q].uv[1] = vt[2 * cs[q].vt + 1];
                }
            t.id = sc.tri_counter++;
            sc.triangles.push_back(t);
        }
    }
    sc.mesh_boxes.push_back(Box{{mn[0], mn[1], mn[2]}, {mx[0], mx[1], mx[2]}});
    g.T_endidx = (int)sc.triangles.size();
    return true;
}

void build_bvh(svgf_scene &sc) {
    if (sc.triangles.empty()) return;
    Builder b;
    b.tris = &sc.triangles;
    b.prims.resize(sc.triangles.size());
    for (size_t i = 0; i < sc.triangles.size(); i++) {
        const svgf_triangle &t = sc.triangles[i];
        Box bb;
        float *mnp = &bb.mn.x, *mxp = &bb.mx.x;
        for (int a = 0; a < 3; a++) {
            mnp[a] = fmin_glm(t.verts[0].pos[a], fmin_glm(t.verts[1].pos[a], t.verts[2].pos[a]));
            mxp[a] = fmax_glm(t.verts[0].pos[a], fmax_glm(t.verts[1].pos[a], t.verts[2].pos[a]));
        }
        b.prims[i].index = (int)i; b.prims[i].bounds = bb;
        b.prims[i].centroid = V3{0.5f * (bb.mn.x + bb.mx.x), 0.5f * (bb.mn.y + bb.mx.y), 0.5f * (bb.mn.z + bb.mx.z)};
    }
    b.ordered.reserve(sc.triangles.size());
    Node *root = b.build(0, (int)sc.triangles.size());
    sc.bvh.assign((size_t)b.total, svgf_bvh_node());
    memset(sc.bvh.data(), 0, sc.bvh.size() * sizeof(svgf_bvh_node));
    int off = 0;
    b.flatten(root, sc.bvh, off);
    Builder::destroy(root);
    sc.triangles.swap(b.ordered);
}

bool parse_scene(svgf_scene &sc, const std::string &path, const std::string &models_dir) {
    std::ifstream in(path.c_str(), std::ios::binary);
    if (!in.is_open()) { sc.err = "cannot open scene file " + path; return false; }
    std::string line;
    while (in.good()) {
        get_line(in, line);
        if (line.empty()) continue;
        std::vector<std::string> tk = tokens_of(line);
        if (tk.empty()) continue;
        if (tk[0] == "MATERIAL") {
            const int id = tk.size() > 1 ? atoi(tk[1].c_str()) : -1;
            if (id != (int)sc.materials.size()) { sc.err = "MATERIAL id does not match its position in the file"; return false; }
            svgf_material m;
            memset(&m, 0, sizeof(m));
            for (int i = 0; i < 7; i++) {           // exactly seven property lines (scene.cpp:187)
                get_line(in, line);
                const std::vector<std::string> t = tokens_of(line);
                if (t.empty()) continue;
                if (t[0] == "RGB") { m.color[0] = num(t, 1); m.color[1] = num(t, 2); m.color[2] = num(t, 3); }
                else if (t[0] == "SPECEX") m.specular_exponent = num(t, 1);
                else if (t[0] == "SPECRGB") { m.specular_color[0] = num(t, 1); m.specular_color[1] = num(t, 2); m.specular_color[2] = num(t, 3); }
                else if (t[0] == "REFL") m.hasReflective = num(t, 1);
                else if (t[0] == "REFR") m.hasRefractive = num(t, 1);
                else if (t[0] == "REFRIOR") m.indexOfRefraction = num(t, 1);
                else if (t[0] == "EMITTANCE") m.emittance = num(t, 1);
            }
            m.texid = -1; m.matid = id;
            get_line(in, line);
            while (!line.empty() && in.good()) {    // optional extras until the blank line
                const std::vector<std::string> t = tokens_of(line);
                if (t.size() > 1 && t[0] == "TEXTURE") { m.texid = (int)sc.texture_files.size(); sc.texture_files.push_back(t[1]); }
                get_line(in, line);
            }
            sc.materials.push_back(m);
        } else if (tk[0] == "OBJECT") {
            const int id = tk.size() > 1 ? atoi(tk[1].c_str()) : -1;
            if (id != (int)sc.geoms.size()) { sc.err = "OBJECT id does not match its position in the file"; return false; }
            svgf_geom g;
            memset(&g, 0, sizeof(g));
            get_line(in, line);
            bool mesh = false;
            if (line == "sphere") g.type = 0; else if (line == "cube") g.type = 1; else if (line == "mesh") { g.type = 2; mesh = true; }
            else { sc.err = "unknown object type '" + line + "'"; return false; }
            get_line(in, line);
            { const std::vector<std::string> t = tokens_of(line); g.materialid = t.size() > 1 ? atoi(t[1].c_str()) : 0; }
            for (int i = 0; i < 3; i++) {
                get_line(in, line);
                const std::vector<std::string> t = tokens_of(line);
                if (t.empty()) continue;
                float *dst = t[0] == "TRANS" ? g.translation : (t[0] == "ROTAT" ? g.rotation : (t[0] == "SCALE" ? g.scale : nullptr));
                if (dst) { dst[0] = num(t, 1); dst[1] = num(t, 2); dst[2] = num(t, 3); }
            }
            const M4 xf = build_transform(V3{g.translation[0], g.translation[1], g.translation[2]}, V3{g.rotation[0], g.rotation[1], g.rotation[2]},
                                          V3{g.scale[0], g.scale[1], g.scale[2]});
            store(g.transform, xf);
            svgf_mat4_inverse(g.transform, g.inverseTransform);
            const M4 it = inverse_transpose(xf);
            store(g.invTranspose, it);
            if (mesh) {
                get_line(in, line);
                if (!load_obj(sc, models_dir + "/" + line, g, xf, it)) return false;
            }
            get_line(in, line);
            while (!line.empty() && in.good()) get_line(in, line);
            sc.geoms.push_back(g);
        } else if (tk[0] == "CAMERA") {
            for (int i = 0; i < 3; i++) {           // only the first three lines are scanned for RES / FOVY / FILE (scene.cpp:126-139)
                get_line(in, line);
                const std::vector<std::string> t = tokens_of(line);
                if (t.empty()) continue;
                if (t[0] == "RES") { sc.res[0] = t.size() > 1 ? atoi(t[1].c_str()) : 0; sc.res[1] = t.size() > 2 ? atoi(t[2].c_str()) : 0; }
                else if (t[0] == "FOVY") sc.fovy = num(t, 1);
            }
            get_line(in, line);
            while (!line.empty() && in.good()) {
                const std::vector<std::string> t = tokens_of(line);
                float *dst = t.empty() ? nullptr : (t[0] == "EYE" ? sc.eye : (t[0] == "LOOKAT" ? sc.lookat : (t[0] == "UP" ? sc.up : nullptr)));
                if (dst) { dst[0] = num(t, 1); dst[1] = num(t, 2); dst[2] = num(t, 3); }
                get_line(in, line);
            }
        }
    }
    for (const svgf_geom &g : sc.geoms)
        if (g.materialid < 0 || g.materialid >= (int)sc.materials.size()) { sc.err = "OBJECT references a MATERIAL that does not exist"; return false; }
    build_bvh(sc);
    sc.textures.resize(sc.texture_files.size());
    return true;
}
}  // namespace

extern "C" {

int svgf_scene_load(svgf_scene **out, const char *scene_file, const char *models_dir) {
    if (!out || !scene_file) return SVGF_ERR_INVALID;
    svgf_scene *sc = new svgf_scene();
    *out = sc;                  // returned even on failure so that svgf_scene_error() can explain; free it either way
    std::string md;
    if (models_dir) md = models_dir;
    else {                      // the reference resolves meshes as ../scenes/Models/<file> (scene.cpp:236): <scene dir>/Models here
        const std::string p(scene_file);
        const size_t k = p.find_last_of('/');
        md = (k == std::string::npos ? std::string(".") : p.substr(0, k)) + "/Models";
    }
    return parse_scene(*sc, scene_file, md) ? SVGF_OK : SVGF_ERR_INVALID;
}

void svgf_scene_free(svgf_scene *sc) { delete sc; }

const char *svgf_scene_error(const svgf_scene *sc) { return sc ? sc->err.c_str() : "no scene"; }

int svgf_scene_describe(svgf_scene *sc, int width, int height, svgf_scene_desc *d) {
    if (!sc || !d || width <= 0 || height <= 0) return SVGF_ERR_INVALID;
    memset(d, 0, sizeof(*d));
    d->geoms = sc->geoms.data(); d->n_geoms = (int)sc->geoms.size();
    d->materials = sc->materials.data(); d->n_materials = (int)sc->materials.size();
    d->triangles = sc->triangles.data(); d->n_triangles = (int)sc->triangles.size();
    d->bvh_nodes = sc->bvh.data(); d->n_bvh_nodes = (int)sc->bvh.size();
    sc->tex_desc.resize(sc->textures.size());
    for (size_t i = 0; i < sc->textures.size(); i++) {
        if (sc->textures[i].px.empty()) { sc->err = "texture '" + sc->texture_files[i] + "' has no pixels yet (svgf_scene_set_texture)"; return SVGF_ERR_INVALID; }
        sc->tex_desc[i] = svgf_texture_desc{sc->textures[i].w, sc->textures[i].h, sc->textures[i].comp, sc->textures[i].px.data()};
    }
    d->textures = sc->tex_desc.data(); d->n_textures = (int)sc->tex_desc.size();
    d->width = width; d->height = height;
    return SVGF_OK;
}

int svgf_scene_camera(const svgf_scene *sc, float eye[3], float lookat[3], float up[3], float *fovy, int res[2]) {
    if (!sc) return SVGF_ERR_INVALID;
    if (eye) memcpy(eye, sc->eye, 12);
    if (lookat) memcpy(lookat, sc->lookat, 12);
    if (up) memcpy(up, sc->up, 12);
    if (fovy) *fovy = sc->fovy;
    if (res) { res[0] = sc->res[0]; res[1] = sc->res[1]; }
    return SVGF_OK;
}

int svgf_scene_num_textures(const svgf_scene *sc) { return sc ? (int)sc->texture_files.size() : 0; }
const char *svgf_scene_texture_file(const svgf_scene *sc, int i) {
    return (sc && i >= 0 && i < (int)sc->texture_files.size()) ? sc->texture_files[i].c_str() : nullptr;
}
int svgf_scene_set_texture(svgf_scene *sc, int i, int width, int height, int components, const unsigned char *pixels) {
    if (!sc || i < 0 || i >= (int)sc->textures.size() || width <= 0 || height <= 0 || components <= 0 || !pixels) return SVGF_ERR_INVALID;
    svgf_scene::Tex &t = sc->textures[i];
    t.w = width; t.h = height; t.comp = components;
    t.px.assign(pixels, pixels + (size_t)width * height * components);
    return SVGF_OK;
}
// Decodes every texture that has no pixels yet from `<textures_dir>/<file name>` (JPEG: csrc/jpeg_decode.cpp, the bytes the
// reference's stb_image produces). Returns the number of textures that now have pixels, or a negative status.
int svgf_scene_load_textures(svgf_scene *sc, const char *textures_dir) {
    if (!sc || !textures_dir) return SVGF_ERR_INVALID;
    int have = 0;
    for (size_t i = 0; i < sc->textures.size(); i++) {
        svgf_scene::Tex &t = sc->textures[i];
        if (!t.px.empty()) { have++; continue; }
        const std::string path = std::string(textures_dir) + "/" + sc->texture_files[i];
        FILE *f = fopen(path.c_str(), "rb");
        if (!f) { sc->err = "cannot open texture '" + path + "'"; continue; }
        std::vector<unsigned char> bytes;
        unsigned char buf[65536];
        size_t n;
        while ((n = fread(buf, 1, sizeof(buf), f)) > 0) bytes.insert(bytes.end(), buf, buf + n);
        fclose(f);
        std::string why;
        if (!svgf_jpeg_decode(bytes.data(), bytes.size(), &t.w, &t.h, &t.comp, t.px, why)) { t.px.clear(); sc->err = "texture '" + path + "': " + why; continue; }
        have++;
    }
    return have;
}
int svgf_jpeg_decode_memory(const unsigned char *bytes, size_t n, int *width, int *height, int *components, unsigned char *out, size_t out_bytes) {
    if (!bytes || !width || !height || !components) return SVGF_ERR_INVALID;
    std::vector<unsigned char> px; std::string why;
    if (!svgf_jpeg_decode(bytes, n, width, height, components, px, why)) return SVGF_ERR_INVALID;
    if (out) {
        if (out_bytes < px.size()) return SVGF_ERR_INVALID;
        memcpy(out, px.data(), px.size());
    }
    return SVGF_OK;
}
int svgf_scene_mesh_boxes(const svgf_scene *sc, float *out6, int max_boxes) {
    if (!sc) return 0;
    const int n = (int)sc->mesh_boxes.size();
    for (int i = 0; i < n && i < max_boxes && out6; i++) memcpy(out6 + 6 * i, &sc->mesh_boxes[i], 24);
    return n;
}

}  // extern "C"
