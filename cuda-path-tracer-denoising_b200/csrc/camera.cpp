// camera.cpp -- host-side camera logic either side of the hot path.
//   svgf_view_matrix   <- GetViewMatrix (src/denoise.cu:342-347): inverse(mat4(right, up, view, position)), computed in
//                         fp32 with glm 0.9.6.3's cofactor expansion (external/include/glm/detail/type_mat4x4.inl:37-92)
//                         so the matrix handed to the temporal kernel has the reference's bits.
//   svgf_camera_init   <- loadCamera's resolution-dependent part (src/scene.cpp:159-168) + resetCamera (src/main.cpp:77-101)
//   svgf_camera_step   <- camera automation + the `camchanged` block of runCuda (src/main.cpp:156-190)
#include <cmath>
#include <cstring>

#include "../../include/svgf_b200.h"

namespace {
const float kPi = 3.1415926535897932384626422832795028841971f;     // utilities.h:12

struct V { float x, y, z; };
inline V sub(V a, V b) { return V{a.x - b.x, a.y - b.y, a.z - b.z}; }
inline float dot3(V a, V b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline V scale(V a, float s) { return V{a.x * s, a.y * s, a.z * s}; }
inline V unit(V a) { return scale(a, 1.0f / sqrtf(dot3(a, a))); }
inline V crs(V a, V b) { return V{a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
inline V get(const float *p) { return V{p[0], p[1], p[2]}; }
inline void put(float *p, V v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }

// 2x2 minors of rows (r0, r1) taken over column pairs, as glm names them (CoefNN)
inline float minor2(const float *m, int ca, int cb, int ra, int rb) {
    return m[ca * 4 + ra] * m[cb * 4 + rb] - m[cb * 4 + ra] * m[ca * 4 + rb];
}
}  // namespace

// glm::inverse(mat4) = detail::compute_inverse (external/include/glm/detail/type_mat4x4.inl:37-92), column-major, fp32.
// Also used by the scene ingest for Geom::inverseTransform (src/scene.cpp:103).
void svgf_mat4_inverse(const float *m, float *out16) {
    // coef[k][*] = {Coef(k*4), Coef(k*4), Coef(k*4+2), Coef(k*4+3)}: minors over rows (2,3) (1,3) (1,2) [k=0], ...
    float fac[6][4];
    const int rowpair[6][2] = {{2, 3}, {1, 3}, {1, 2}, {0, 3}, {0, 2}, {0, 1}};
    for (int k = 0; k < 6; k++) {
        const int ra = rowpair[k][0], rb = rowpair[k][1];
        const float c0 = minor2(m, 2, 3, ra, rb), c2 = minor2(m, 1, 3, ra, rb), c3 = minor2(m, 1, 2, ra, rb);
        fac[k][0] = c0; fac[k][1] = c0; fac[k][2] = c2; fac[k][3] = c3;
    }
    float vec[4][4];
    for (int r = 0; r < 4; r++) { vec[r][0] = m[1 * 4 + r]; vec[r][1] = vec[r][2] = vec[r][3] = m[0 * 4 + r]; }
    float inv[4][4];
    for (int i = 0; i < 4; i++) {
        const float sa = (i & 1) ? -1.f : 1.f, sb = -sa;
        inv[0][i] = ((vec[1][i] * fac[0][i] - vec[2][i] * fac[1][i]) + vec[3][i] * fac[2][i]) * sa;
        inv[1][i] = ((vec[0][i] * fac[0][i] - vec[2][i] * fac[3][i]) + vec[3][i] * fac[4][i]) * sb;
        inv[2][i] = ((vec[0][i] * fac[1][i] - vec[1][i] * fac[3][i]) + vec[3][i] * fac[5][i]) * sa;
        inv[3][i] = ((vec[0][i] * fac[2][i] - vec[1][i] * fac[4][i]) + vec[2][i] * fac[5][i]) * sb;
    }
    const float d0 = m[0] * inv[0][0], d1 = m[1] * inv[1][0], d2 = m[2] * inv[2][0], d3 = m[3] * inv[3][0];
    const float rdet = 1.0f / ((d0 + d1) + (d2 + d3));
    for (int cidx = 0; cidx < 4; cidx++) for (int r = 0; r < 4; r++) out16[cidx * 4 + r] = inv[cidx][r] * rdet;
}

void svgf_view_matrix(const svgf_camera *cam, float *out16) {
    const float m[16] = {cam->right[0], cam->right[1], cam->right[2], 0.f, cam->up[0], cam->up[1], cam->up[2], 0.f,
                         cam->view[0], cam->view[1], cam->view[2], 0.f, cam->position[0], cam->position[1], cam->position[2], 1.f};
    svgf_mat4_inverse(m, out16);
}

extern "C" void svgf_camera_init(svgf_camera *cam, svgf_camera_rig *rig, const float eye[3], const float lookat[3],
                                 const float up[3], float fovy, int width, int height) {
    memset(cam, 0, sizeof(*cam));
    memset(rig, 0, sizeof(*rig));
    cam->resolution[0] = width; cam->resolution[1] = height;
    const float yscaled = tanf(fovy * (kPi / 180));
    const float xscaled = (yscaled * width) / height;
    cam->fov[0] = (atanf(xscaled) * 180) / kPi; cam->fov[1] = fovy;
    cam->pixelLength[0] = 2 * xscaled / (float)width; cam->pixelLength[1] = 2 * yscaled / (float)height;
    put(cam->position, get(eye)); put(cam->lookAt, get(lookat)); put(cam->up, get(up));
    const V view = unit(sub(get(lookat), get(eye)));
    put(cam->view, view);
    rig->fovy = fovy;
    rig->phi = acosf(dot3(unit(V{view.x, 0.0f, view.z}), V{0, 0, -1}));
    rig->theta = acosf(dot3(unit(V{0.0f, view.y, view.z}), V{0, 1, 0}));
    const V d = sub(get(eye), get(lookat));
    rig->zoom = sqrtf(dot3(d, d));
}

extern "C" void svgf_camera_step(svgf_camera *cam, svgf_camera_rig *rig, int automate, const float sp[5]) {
    if (automate) {
        rig->tx += sp[0]; rig->ty += sp[1]; rig->tz += sp[2]; rig->ttheta += sp[3]; rig->tphi += sp[4];
        cam->lookAt[0] = 0.0f + 2.0f * sinf(rig->tx);
        cam->lookAt[1] = 5.0f + 1.0f * sinf(rig->ty);
        cam->lookAt[2] = 0.0f + 1.5f * sinf(rig->tz);
        rig->theta = kPi * 0.5f + kPi / 18 * sinf(rig->ttheta);
        rig->phi = kPi * 0.0f + kPi / 12 * sinf(rig->tphi);
    }
    V cp;
    cp.x = rig->zoom * sinf(rig->phi) * sinf(rig->theta);
    cp.y = rig->zoom * cosf(rig->theta);
    cp.z = rig->zoom * cosf(rig->phi) * sinf(rig->theta);
    const V v = scale(unit(cp), -1.0f);
    const V r = crs(v, V{0, 1, 0});
    put(cam->view, v); put(cam->up, crs(r, v)); put(cam->right, r);
    const V la = get(cam->lookAt);
    put(cam->position, V{cp.x + la.x, cp.y + la.y, cp.z + la.z});
}
