// halo_sync.cuh -- device side of the cross-GPU ordering of a sharded frame, shared by the producer kernels.
//
// A frame is sharded by row strips, one GPU per strip. A stage's producer kernel stores the rows its neighbours will tap
// straight into THEIR copy of the plane as well ("dual stores": the same registers go to local HBM and over NVLink, no
// separate copy kernel), and the last block of the kernel to finish raises the stage's sequence flag in those neighbours'
// memory. The consumer side polls its own (local) copy of the flags in the few blocks that touch halo rows. Only the ranks
// whose strips lie within reach of each other take part: a level is never gated by a GPU it does not exchange rows with.
//
// Memory ordering: every thread fences its peer stores at system scope before its block is counted; the block that
// observes the full count fences again and then writes the flags, so a rank that sees flag >= seq sees all rows.
#pragma once
#include "svgf_internal.h"

#ifdef __CUDACC__
// Call once per block after all of the block's stores, by ALL threads that have not exited (no early returns before it).
// `stored_to_peers` (block-uniform): did any thread of this block store into a neighbour's planes? Only such blocks have
// something to publish and pay for the system-scope fence; the others are merely counted (every block must be: the flag also
// tells the neighbours that this rank is done READING the planes their next stage overwrites). Measured on 2 x B200 at 4K:
// with the fence in every block a non-final level took 231 us against 140 us for the strip's own work.
__device__ __forceinline__ void halo_block_done(const HaloOut &h, bool stored_to_peers) {
    if (h.peers.n == 0 || !h.signal) return;
    if (stored_to_peers) __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0 && threadIdx.z == 0) {
        const unsigned total = gridDim.x * gridDim.y * gridDim.z;
        if (atomicAdd(h.counter, 1u) == total - 1u) {
            *h.counter = 0u;                // the next launch of this stage starts from zero (stream order)
            __threadfence_system();
            for (int i = 0; i < h.peers.n; i++) *reinterpret_cast<volatile unsigned *>(h.flag[i]) = h.seq;
        }
    }
}

// Does any row of [y0, y1] (inclusive) belong to the rows some peer taps? (block-uniform argument for halo_block_done)
__device__ __forceinline__ bool halo_rows_touch(const HaloPeers &p, int y0, int y1) {
    bool t = false;
    for (int i = 0; i < p.n; i++) t |= (y0 < p.hi[i] && y1 >= p.lo[i]);
    return t;
}

// Poll by ONE thread of the block (then __syncthreads by the caller). Bounded: a peer that died must not hang the GPU.
__device__ __forceinline__ void halo_wait(const HaloIn &w) {
    const long long t0 = clock64();
    for (int i = 0; i < w.n; i++) {
        const volatile unsigned *f = w.flag[i];
        while ((int)(*f - w.seq) < 0) {
            if (clock64() - t0 > 4000000000LL) { *reinterpret_cast<volatile unsigned *>(w.err) = 1u; break; }
            __nanosleep(100);
        }
    }
    __threadfence_system();
}

// Which peers tap row y of my strip (bit i = peers.rank[i]).
__device__ __forceinline__ unsigned halo_targets(const HaloPeers &p, int y) {
    unsigned m = 0;
    for (int i = 0; i < p.n; i++) m |= (y >= p.lo[i] && y < p.hi[i]) ? (1u << i) : 0u;
    return m;
}
#endif
