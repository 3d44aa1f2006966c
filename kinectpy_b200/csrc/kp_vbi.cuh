// kp_vbi.cuh -- voxel-brick index: the spatial index of a voxel-downsampled cloud.
//
// Every neighbour search of the frame path (SOR after voxel_down_sample, normals and ICP correspondences on the
// down-sampled master: preprocessing/filtering.py:23-24, registration.py:8-13, 78-84) runs on a cloud with AT MOST ONE
// point per voxel, and that point (the voxel mean) lies inside its voxel.  The index is therefore an occupancy BITMAP
// at voxel resolution: bricks of 2 x 4 x 4 voxels, 32 bits each (everything a query does with a brick is 32-bit
// integer arithmetic), in a dense brick table {occupancy, position of the brick's first point} (one 8-byte load), the
// points sorted by (brick, bit).  The position of a voxel's point is first + popcount(occupancy below its bit): no
// run table, no hashing, and a query prunes 32 voxels with one AND of the occupancy word against a box mask built
// from three per-axis ranges.
#pragma once
#include "kp_common.cuh"

constexpr int KP_VBI_PAD = 2;          // empty bricks around the grid: a 5 x 5 x 5 brick neighbourhood never leaves the table

struct KpVbiDev {
    const uint2 *bricks;      // [nb0 * nb1 * nb2] {occupancy, first position}
    const float4 *pts;        // sorted by (brick, bit); .w = original index (int bits)
    const uint32_t *vijk;     // packed voxel coordinates (ix << 20 | iy << 10 | iz) of pts[]
    double minb[3];           // voxel grid origin (voxel_down_sample's min_bound - voxel / 2)
    double voxel, inv_voxel;
    double eps;               // a stored point lies inside its voxel's box widened by eps (float32 rounding of the mean)
    int nb[3];                // bricks per axis, padding included
    int nvox[3];              // voxels per axis the grid spans (brick-aligned)
    int npts;
    int ok;                   // 0: not built (empty cloud, or the extent exceeds the brick table / 10-bit coordinates)
};

#ifdef __CUDACC__
// bit of voxel (vx, vy, vz) inside its brick: vx * 16 + vy * 4 + vz   (vx 0..1, vy, vz 0..3)
__device__ __forceinline__ unsigned kp_vbi_zmask(unsigned zm) { return zm * 0x11111111u; }
__device__ __forceinline__ unsigned kp_vbi_ymask(unsigned ym)
{
    const unsigned t = ((ym & 1u) | ((ym & 2u) << 3) | ((ym & 4u) << 6) | ((ym & 8u) << 9)) * 0xFu;   // nibble vy = 0xF iff bit vy of ym
    return t * 0x00010001u;
}
__device__ __forceinline__ unsigned kp_vbi_xmask(unsigned xm) { return ((xm & 1u) ? 0x0000FFFFu : 0u) | ((xm & 2u) ? 0xFFFF0000u : 0u); }
// mask of the voxels lo..hi (grid coordinates) that fall into brick b (unpadded brick coordinate) along an axis with
// 4 (y, z) or 2 (x) voxels per brick
__device__ __forceinline__ unsigned kp_vbi_axis_mask4(int lo, int hi, int b)
{
    const int a = max(lo - 4 * b, 0), c = min(hi - 4 * b, 3);
    return a <= c ? ((2u << c) - 1u) & ~((1u << a) - 1u) : 0u;
}
__device__ __forceinline__ unsigned kp_vbi_axis_mask2(int lo, int hi, int b)
{
    const int a = max(lo - 2 * b, 0), c = min(hi - 2 * b, 1);
    return a <= c ? ((2u << c) - 1u) & ~((1u << a) - 1u) : 0u;
}
__device__ __forceinline__ long long kp_vbi_brick_index(const KpVbiDev &v, int bx, int by, int bz)   // unpadded brick coordinates
{
    return ((long long)(bx + KP_VBI_PAD) * v.nb[1] + (by + KP_VBI_PAD)) * v.nb[2] + (bz + KP_VBI_PAD);
}
// visits every point of the index inside the voxel box lo..hi (grid coordinates, inside the grid):
// fn(position, point); two candidates per trip so that both point loads are in flight together
struct KpVbiBox { int x0, x1, y0, y1, z0, z1; };
template <class F>
__device__ __forceinline__ void kp_vbi_visit(const KpVbiDev &v, const KpVbiBox bx6, F &&fn)
{
    const int bz0 = bx6.z0 >> 2, bz1 = bx6.z1 >> 2;
    for (int bx = bx6.x0 >> 1; bx <= bx6.x1 >> 1; ++bx) {
        const unsigned mx = kp_vbi_xmask(kp_vbi_axis_mask2(bx6.x0, bx6.x1, bx));
        for (int by = bx6.y0 >> 2; by <= bx6.y1 >> 2; ++by) {
            const unsigned mxy = mx & kp_vbi_ymask(kp_vbi_axis_mask4(bx6.y0, bx6.y1, by));
            const uint2 *row = v.bricks + kp_vbi_brick_index(v, bx, by, 0);
            uint2 e = __ldg(row + bz0);
            for (int bz = bz0; bz <= bz1; ++bz) {
                const uint2 cur = e;
                if (bz < bz1) e = __ldg(row + bz + 1);            // next brick in flight while this one is walked
                unsigned m = cur.x & mxy & kp_vbi_zmask(kp_vbi_axis_mask4(bx6.z0, bx6.z1, bz));
                while (m) {
                    const int b0 = __ffs((int)m) - 1;
                    m &= m - 1;
                    const bool two = m != 0;
                    const int b1 = two ? __ffs((int)m) - 1 : b0;
                    m &= m - 1;                                   // (0 & anything = 0)
                    const int p0 = (int)cur.y + __popc(cur.x & ((1u << b0) - 1u));
                    const int p1 = (int)cur.y + __popc(cur.x & ((1u << b1) - 1u));
                    const float4 c0 = __ldg(v.pts + p0), c1 = __ldg(v.pts + p1);
                    fn(p0, c0);
                    if (two) fn(p1, c1);
                }
            }
        }
    }
}
// voxel box of the ball of radius `rad` around q, clipped to the grid; false when it misses the grid altogether
__device__ __forceinline__ bool kp_vbi_box(const KpVbiDev &v, double qx, double qy, double qz, double rad, KpVbiBox &o)
{
    const double ax = floor((qx - rad - v.minb[0]) * v.inv_voxel), bx = floor((qx + rad - v.minb[0]) * v.inv_voxel);
    const double ay = floor((qy - rad - v.minb[1]) * v.inv_voxel), by = floor((qy + rad - v.minb[1]) * v.inv_voxel);
    const double az = floor((qz - rad - v.minb[2]) * v.inv_voxel), bz = floor((qz + rad - v.minb[2]) * v.inv_voxel);
    const double mx = (double)(v.nvox[0] - 1), my = (double)(v.nvox[1] - 1), mz = (double)(v.nvox[2] - 1);
    // (the negated comparisons also catch NaN)
    const bool any = (bx >= 0.0) && (ax <= mx) && (by >= 0.0) && (ay <= my) && (bz >= 0.0) && (az <= mz);
    o.x0 = (int)fmax(ax, 0.0); o.x1 = (int)fmin(bx, mx);
    o.y0 = (int)fmax(ay, 0.0); o.y1 = (int)fmin(by, my);
    o.z0 = (int)fmax(az, 0.0); o.z1 = (int)fmin(bz, mz);
    return any;
}
#endif
