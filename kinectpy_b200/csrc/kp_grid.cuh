// kp_grid.cuh -- device view of the uniform-grid spatial hash (K3) shared by the neighbour
// kernels (kp_grid.cu) and the ICP correspondence kernel (kp_icp.cu).
#pragma once
#include "kp_common.cuh"

struct KpGridDev {
    const float4 *pts;        // sorted by cell key; .w = original index (int bits)
    const uint64_t *hkeys;    // open addressing, ~0ull = empty
    const int2 *hvals;        // [start, end) into pts
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
// [start,end) of a cell, (0,0) when empty or outside the grid
__device__ __forceinline__ int2 kp_cell_range(const KpGridDev &g, int cx, int cy, int cz)
{
    if ((unsigned)cx >= (unsigned)g.dim[0] || (unsigned)cy >= (unsigned)g.dim[1] || (unsigned)cz >= (unsigned)g.dim[2])
        return make_int2(0, 0);
    if (g.bitmap) {
        const long long b = ((long long)cx * g.dim[1] + cy) * g.dim[2] + cz;
        if (!((__ldg(g.bitmap + (b >> 5)) >> (b & 31)) & 1u)) return make_int2(0, 0);
    }
    const uint64_t key = kp_cell_key(g, cx, cy, cz);
    uint32_t h = (uint32_t)kp_mix64(key) & g.hmask;
    for (;;) {
        const uint64_t k = __ldg(g.hkeys + h);
        if (k == key) return __ldg(g.hvals + h);
        if (k == ~0ull) return make_int2(0, 0);
        h = (h + 1) & g.hmask;
    }
}
#endif

KpGridDev kp_grid_dev(const KpGrid &g);
