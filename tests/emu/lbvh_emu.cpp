// lbvh_emu.cpp -- host emulation of svgf_rebuild_bvh (csrc/lbvh.cu): the steps of csrc/lbvh_core.h, the very functions the
// CUDA kernels call, run index by index; std::sort stands in for cub's radix sort (keys are unique, so the order is the same).
// TEST INFRASTRUCTURE (CPU suite): checks the tree and renders through it with the oracle; the GPU suite checks that the
// device build produces the same arrays.
#include <algorithm>
#include <vector>

#include "../../cuda-path-tracer-denoising_b200/csrc/lbvh_core.h"

// tri_hot: 3 x {x,y,z,w} per triangle ({v0, id}, {e1, .}, {e2, .}). out_nodes: 2 x {x,y,z,w} per node, 2n-1 nodes. out_order: the
// slot (in the input order) of the triangle at every sorted position. Returns the number of nodes.
extern "C" int lbvh_emu_build(const float *tri_hot_f, int n, float *out_nodes_f, int *out_order) {
    if (n <= 0) return 0;
    const LbvhF4 *hot = reinterpret_cast<const LbvhF4 *>(tri_hot_f);
    LbvhF4 *nodes = reinterpret_cast<LbvhF4 *>(out_nodes_f);
    const int nn = 2 * n - 1;
    std::vector<float> tri_b6(6 * (size_t)n), node_b6(6 * (size_t)nn);
    float scene6[6] = {3e38f, 3e38f, 3e38f, -3e38f, -3e38f, -3e38f};
    for (int k = 0; k < n; k++) lbvh_tri_bounds(hot, k, &tri_b6[6 * (size_t)k]);
    for (int k = 0; k < n; k++)
        for (int a = 0; a < 3; a++) { scene6[a] = lbvh_min(scene6[a], tri_b6[6 * (size_t)k + a]); scene6[3 + a] = lbvh_max(scene6[3 + a], tri_b6[6 * (size_t)k + 3 + a]); }
    std::vector<uint64_t> keys(n);
    for (int k = 0; k < n; k++) keys[k] = lbvh_key(&tri_b6[6 * (size_t)k], scene6, k);
    std::sort(keys.begin(), keys.end());
    std::vector<int> left(n), right(n), axis(n), flags(n, 0), parent(2 * (size_t)n, 0), size(2 * (size_t)n, 0);
    for (int i = 0; i < n - 1; i++) lbvh_internal(keys.data(), n, i, left.data(), right.data(), parent.data(), axis.data());
    for (int k = 0; k < n; k++)
        lbvh_climb(k, n, keys.data(), tri_b6.data(), left.data(), right.data(), parent.data(), node_b6.data(), size.data(),
                   [&flags](int p) { return flags[p]++; });
    for (int v = 0; v < nn; v++) lbvh_emit(v, n, left.data(), right.data(), parent.data(), size.data(), axis.data(), node_b6.data(), nodes);
    for (int k = 0; k < n; k++) out_order[k] = (int)(uint32_t)keys[k];
    return nn;
}
