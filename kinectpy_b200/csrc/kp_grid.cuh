// kp_grid.cuh -- device view of the uniform-grid spatial hash (K3) shared by the neighbour
// kernels (kp_grid.cu) and the ICP correspondence kernel (kp_icp.cu).
#pragma once
#include "kp_common.cuh"

struct KpGridDev {
    const float4 *pts;        // sorted by cell key; .w = original index (int bits)
    const uint4 *slots;       // open addressing, 16-byte slots {key lo, key hi, start, end}; key ~0 = empty.
                              // One 128-bit load returns the key AND the cell's [start,end) run in pts.
    uint32_t hmask;
    int sh_x, sh_y;           // key = cx << sh_x | cy << sh_y | cz   (cz is the low field: the cells of a
                              // z-row are adjacent in the sorted array)
    int dim[3];
    int npts;                 // rows of pts (NaN rows, if any, sit at the end)
    const uint32_t *bitmap;   // one bit per cell, index (cx*dim1 + cy)*dim2 + cz; NULL when the grid is too large.
                              // Surfaces are thin: most of a query's 27 cells are empty, and a bitmap word comes
                              // from L1 while a hash probe goes to L2
    double org[3];
    double cell, inv_cell;
};

#ifdef __CUDACC__
__device__ __forceinline__ int kp_cell_coord(const KpGridDev &g, double v, int c)
{
    return (int)floor(__dmul_rn(__dsub_rn(v, g.org[c]), g.inv_cell));
}
__device__ __forceinline__ uint64_t kp_cell_key(const KpGridDev &g, int cx, int cy, int cz)
{
    return ((uint64_t)(uint32_t)cx << g.sh_x) | ((uint64_t)(uint32_t)cy << g.sh_y) | (uint64_t)(uint32_t)cz;
}
__device__ __forceinline__ uint64_t kp_slot_key(const uint4 &s) { return ((uint64_t)s.y << 32) | s.x; }
__device__ __forceinline__ bool kp_cell_bit(const KpGridDev &g, int cx, int cy, int cz)
{
    const long long b = ((long long)cx * g.dim[1] + cy) * g.dim[2] + cz;
    return (__ldg(g.bitmap + (b >> 5)) >> (b & 31)) & 1u;
}
// resolve a probe sequence that starts at slot h with the already loaded slot s
__device__ __forceinline__ int2 kp_slot_resolve(const KpGridDev &g, uint64_t key, uint32_t h, uint4 s)
{
    for (;;) {
        const uint64_t k = kp_slot_key(s);
        if (k == key) return make_int2((int)s.z, (int)s.w);
        if (k == ~0ull) return make_int2(0, 0);
        h = (h + 1) & g.hmask;
        s = __ldg(g.slots + h);
    }
}
// [start,end) of a cell, (0,0) when empty or outside the grid
__device__ __forceinline__ int2 kp_cell_range(const KpGridDev &g, int cx, int cy, int cz)
{
    if ((unsigned)cx >= (unsigned)g.dim[0] || (unsigned)cy >= (unsigned)g.dim[1] || (unsigned)cz >= (unsigned)g.dim[2])
        return make_int2(0, 0);
    if (g.bitmap && !kp_cell_bit(g, cx, cy, cz)) return make_int2(0, 0);
    const uint64_t key = kp_cell_key(g, cx, cy, cz);
    const uint32_t h = (uint32_t)kp_mix64(key) & g.hmask;
    return kp_slot_resolve(g, key, h, __ldg(g.slots + h));
}
// Merged [start,end) of the z-row (cx, cy, cz-1 .. cz+1): the three cells are adjacent in pts, so the
// union of the non-empty ones is one contiguous run.  Written so the (up to) three bitmap tests are issued
// together and then the (up to) three first hash probes are issued together: two dependent memory round
// trips per row instead of up to nine.
__device__ __forceinline__ int2 kp_row_range(const KpGridDev &g, int cx, int cy, int cz)
{
    if ((unsigned)cx >= (unsigned)g.dim[0] || (unsigned)cy >= (unsigned)g.dim[1]) return make_int2(0, 0);
    bool act[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) act[j] = (unsigned)(cz - 1 + j) < (unsigned)g.dim[2];
    if (g.bitmap) {
        bool bit[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) bit[j] = act[j] ? kp_cell_bit(g, cx, cy, cz - 1 + j) : false;
#pragma unroll
        for (int j = 0; j < 3; ++j) act[j] = bit[j];
    }
    uint64_t key[3];
    uint32_t h[3];
    uint4 s[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        key[j] = kp_cell_key(g, cx, cy, cz - 1 + j);
        h[j] = (uint32_t)kp_mix64(key[j]) & g.hmask;
        s[j] = make_uint4(0xffffffffu, 0xffffffffu, 0u, 0u);
        if (act[j]) s[j] = __ldg(g.slots + h[j]);
    }
    int a = 0x7fffffff, b = 0;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        if (!act[j]) continue;
        const int2 r = kp_slot_resolve(g, key[j], h[j], s[j]);
        if (r.y > r.x) { a = min(a, r.x); b = max(b, r.y); }
    }
    return b > 0 ? make_int2(a, b) : make_int2(0, 0);
}
#endif

KpGridDev kp_grid_dev(const KpGrid &g);
