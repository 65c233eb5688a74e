// oracle/ref/shim_hooks.cpp -- TEST INFRASTRUCTURE. The headless harness (harness.cpp) dumps the reference TUs' file statics through
// refh_fetch_pathtrace / refh_fetch_denoise; when the harness is linked against the PRODUCT's drop-in shim instead
// (libref_shim.so), these two functions answer from the shim's context through the public C ABI. They live here, not in the
// shim a maintainer compiles (cuda-path-tracer-denoising_b200/shim/svgf_shim.cpp).
#include <cstring>
#include "svgf_b200.h"

extern "C" svgf_ctx *svgf_shim_context(void);

extern "C" int refh_fetch_pathtrace(const char *name, void *host, size_t bytes) {
    svgf_ctx *ctx = svgf_shim_context();
    if (!ctx) return -1;
    if (!strcmp(name, "image") || !strcmp(name, "denoised") || !strcmp(name, "gbuffer")) return svgf_fetch(ctx, name, host, bytes) ? -2 : 0;
    return 1;
}
extern "C" int refh_fetch_denoise(const char *name, void *host, size_t bytes) {
    svgf_ctx *ctx = svgf_shim_context();
    if (!ctx) return -1;
    int rc = svgf_fetch(ctx, name, host, bytes);
    return rc == SVGF_ERR_UNKNOWN_NAME ? 1 : (rc ? -2 : 0);
}
