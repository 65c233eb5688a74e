// oracle/ref/tu_denoise.cu -- TEST INFRASTRUCTURE (oracle/), not product code.
// Translation unit that compiles the reference's src/denoise.cu *textually, from where it lies*
// (the staged, portability-patched copy produced by stage.sh; -I<stage>/src) and adds accessors for
// its file-static device buffers (src/denoise.cu:14-27) so tests can read every intermediate.
#include "denoise.cu"
#include <cstring>

extern "C" int refh_fetch_denoise(const char *name, void *host, size_t bytes) {
    if (!hst_scene) return -1;
    const Camera &cam = hst_scene->state.camera;
    const size_t px = (size_t)cam.resolution.x * cam.resolution.y;
    const void *src = NULL; size_t need = 0;
    if      (!strcmp(name, "variance"))              { src = dev_variance;              need = px * 4; }
    else if (!strcmp(name, "color_acc"))             { src = dev_color_acc;             need = px * 12; }
    else if (!strcmp(name, "color_history"))         { src = dev_color_history;         need = px * 12; }
    else if (!strcmp(name, "moment_acc"))            { src = dev_moment_acc;            need = px * 8; }
    else if (!strcmp(name, "moment_history"))        { src = dev_moment_history;        need = px * 8; }
    else if (!strcmp(name, "history_length"))        { src = dev_history_length;        need = px * 4; }
    else if (!strcmp(name, "history_length_update")) { src = dev_history_length_update; need = px * 4; }
    else if (!strcmp(name, "gbuffer_prev"))          { src = dev_gbuffer_prev;          need = px * sizeof(GBufferTexel); }
    else if (!strcmp(name, "temp0"))                 { src = dev_temp[0];               need = px * 12; }
    else if (!strcmp(name, "temp1"))                 { src = dev_temp[1];               need = px * 12; }
    else return 1;  // not ours
    if (bytes != need || !src) return -2;
    cudaMemcpy(host, src, need, cudaMemcpyDeviceToHost);
    return 0;
}

extern "C" void refh_get_view_matrix_prev(float *out16) { memcpy(out16, &view_matrix_prev, 64); }
