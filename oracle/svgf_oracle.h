/* oracle/svgf_oracle.h -- TEST INFRASTRUCTURE. CPU restatement of the reference's pathtrace() -> denoise()
 * hot path (see svgf_oracle.cpp for the file:line map). Only tests/, bench.py's cpu_baseline/reference legs
 * and __graft_entry__.smoke() may load this library; the product (libsvgf_b200.so) never does.
 *
 * Pinning: tests/test_oracle_vs_reference.py checks this code BIT FOR BIT against the reference's own
 * sources executed on the CPU emulator (oracle/_ref/libref_cpu*.so, built from /root/reference), and the
 * committed fixtures in tests/golden/ hold outputs of the reference's own CUDA build run on a B200.
 */
#ifndef SVGF_ORACLE_H
#define SVGF_ORACLE_H
#include "../include/svgf_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_scene orc_scene;
typedef struct orc_state orc_state;

/* variance handling of the a-trous filter */
enum { ORC_VAR_JACOBI = 0,      /* double-buffered: what libref_*_jacobi.so computes */
       ORC_VAR_INPLACE_SEQ = 1  /* in place, pixels visited in the CPU emulator's launch order
                                   (8x8 blocks row-major, threads row-major inside): what libref_cpu.so computes */ };

orc_scene *orc_scene_load(const char *blob_path);
void orc_scene_free(orc_scene *);
int orc_scene_counts(const orc_scene *, int *out6);
int orc_scene_camera(const orc_scene *, svgf_camera *cam_out, float *fovy_out);   /* loader camera (view/up/eye/lookAt) */
/* raw arrays for handing the scene to svgf_create (the product has no loader of its own yet) */
int orc_scene_desc(const orc_scene *, svgf_scene_desc *out, svgf_texture_desc *tex_out, int max_tex);

orc_state *orc_create(const orc_scene *, int W, int H);
void orc_destroy(orc_state *);
void orc_reset(orc_state *);    /* pathtraceInit + denoiseInit semantics */

/* pathtrace(pbo, frame) == orc_pathtrace + (denoise_enable ? orc_denoise : copy) + orc_pack_pbo */
int orc_frame(orc_state *, const svgf_camera *, const svgf_params *, int frame, int variance_mode, int threads);
/* denoise(output, input, gbuffer) on caller buffers (vec3 colour, 52-byte texels), src/denoise.h:8 */
int orc_denoise(orc_state *, float *output, const float *input, const svgf_gbuffer_texel *gbuffer, const svgf_camera *,
                const svgf_params *, int variance_mode, int threads);
int orc_fetch(orc_state *, const char *name, void *host, size_t bytes);
int orc_host_intersect(const orc_scene *, const float *origin, const float *dir, float *t, float *normal,
                       float *uv, int *geomId, int *materialId);

/* Stand-alone kernels on caller buffers (reference AoS layouts). */
int orc_atrous_level(float *color_out, float *variance_out, const float *color_in, const float *variance_in,
                     const svgf_gbuffer_texel *gbuffer, int W, int H, int level, int is_last,
                     float sigma_c, float sigma_n, float sigma_x, int blur_variance, int addcolor,
                     int variance_mode, int threads);
int orc_max_threads(void);

/* Harness duties restated (src/main.cpp:77-101, 154-190; src/scene.cpp:159-166) */
void orc_camera_init(svgf_camera *cam, svgf_camera_rig *rig, const float eye[3], const float lookat[3],
                     const float up[3], float fovy, int W, int H);
void orc_camera_step(svgf_camera *cam, svgf_camera_rig *rig, int automate, const float speeds[5]);
void orc_view_matrix(const svgf_camera *cam, float *out16);   /* GetViewMatrix, denoise.cu:342-347 */

#ifdef __cplusplus
}
#endif
#endif
