// kp_voxel.cu -- K2: voxel-grid downsample = key generation -> stable radix sort -> segmented mean.
// Replaces PointCloud.voxel_down_sample (preprocessing/filtering.py:23, registration.py:8,100-101,
// utils/processing.py:308).  Voxel index = floor((p - (min - v/2)) / v) in double, IEEE division.
#include <math.h>
#include "kp_common.cuh"

namespace {
struct VoxParams {
    double minb[3];
    double voxel;
    int sh_x, sh_y;          // key = ix << sh_x | iy << sh_y | iz
    unsigned long long sentinel;
};

template <class K>
__global__ void __launch_bounds__(256) k_voxel_keys(const float *xyz, int64_t n, const __grid_constant__ VoxParams vp, K *keys)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
    K key;
    if (isnan(x)) {
        key = (K)vp.sentinel;
    } else {
        long long ix = (long long)floor(__ddiv_rn(__dsub_rn((double)x, vp.minb[0]), vp.voxel));
        long long iy = (long long)floor(__ddiv_rn(__dsub_rn((double)y, vp.minb[1]), vp.voxel));
        long long iz = (long long)floor(__ddiv_rn(__dsub_rn((double)z, vp.minb[2]), vp.voxel));
        key = (K)(((unsigned long long)ix << vp.sh_x) | ((unsigned long long)iy << vp.sh_y) | (unsigned long long)iz);
    }
    keys[i] = key;
}

// one thread per voxel run: sums its points in input order (the sort is stable) in double, like
// the insertion loop of Open3D's hash-map accumulator, then one division and one rounding.
template <class K>
__global__ void __launch_bounds__(128) k_voxel_mean(const K *keys, const int32_t *vals, const int32_t *run_start, int R,
                                                    int nvalid, const float *xyz, const float *colors,
                                                    const float *normals, const __grid_constant__ VoxParams vp,
                                                    float *out_xyz, float *out_colors, float *out_normals,
                                                    int32_t *out_ijk, int32_t *point_voxel)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    int a = run_start[r], b = (r + 1 < R) ? run_start[r + 1] : nvalid;
    double sx = 0, sy = 0, sz = 0, cr = 0, cg = 0, cb = 0, nx = 0, ny = 0, nz = 0;
    for (int t = a; t < b; ++t) {
        int64_t i = vals[t];
        sx = __dadd_rn(sx, (double)xyz[3 * i]);
        sy = __dadd_rn(sy, (double)xyz[3 * i + 1]);
        sz = __dadd_rn(sz, (double)xyz[3 * i + 2]);
        if (colors) {
            cr = __dadd_rn(cr, (double)colors[3 * i]);
            cg = __dadd_rn(cg, (double)colors[3 * i + 1]);
            cb = __dadd_rn(cb, (double)colors[3 * i + 2]);
        }
        if (normals) {
            nx = __dadd_rn(nx, (double)normals[3 * i]);
            ny = __dadd_rn(ny, (double)normals[3 * i + 1]);
            nz = __dadd_rn(nz, (double)normals[3 * i + 2]);
        }
        if (point_voxel) point_voxel[i] = r;
    }
    double cnt = (double)(b - a);
    out_xyz[3 * (int64_t)r] = (float)__ddiv_rn(sx, cnt);
    out_xyz[3 * (int64_t)r + 1] = (float)__ddiv_rn(sy, cnt);
    out_xyz[3 * (int64_t)r + 2] = (float)__ddiv_rn(sz, cnt);
    if (colors && out_colors) {
        out_colors[3 * (int64_t)r] = (float)__ddiv_rn(cr, cnt);
        out_colors[3 * (int64_t)r + 1] = (float)__ddiv_rn(cg, cnt);
        out_colors[3 * (int64_t)r + 2] = (float)__ddiv_rn(cb, cnt);
    }
    if (normals && out_normals) {
        double ax = __ddiv_rn(nx, cnt), ay = __ddiv_rn(ny, cnt), az = __ddiv_rn(nz, cnt);
        double nn = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(ax, ax), __dmul_rn(ay, ay)), __dmul_rn(az, az)));
        if (nn > 0.0) { ax = __ddiv_rn(ax, nn); ay = __ddiv_rn(ay, nn); az = __ddiv_rn(az, nn); }
        out_normals[3 * (int64_t)r] = (float)ax;
        out_normals[3 * (int64_t)r + 1] = (float)ay;
        out_normals[3 * (int64_t)r + 2] = (float)az;
    }
    if (out_ijk) {
        unsigned long long key = (unsigned long long)keys[a];
        out_ijk[3 * (int64_t)r] = (int32_t)(key >> vp.sh_x);
        out_ijk[3 * (int64_t)r + 1] = (int32_t)((key >> vp.sh_y) & ((1ull << (vp.sh_x - vp.sh_y)) - 1ull));
        out_ijk[3 * (int64_t)r + 2] = (int32_t)(key & ((1ull << vp.sh_y) - 1ull));
    }
}

__global__ void k_fill_i32(int32_t *p, int64_t n, int32_t v)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

static int bit_length(long long v)
{
    int b = 0;
    while (v > 0) { ++b; v >>= 1; }
    return b < 1 ? 1 : b;
}

template <class K>
int voxel_run(kp_ctx *ctx, const float *d_xyz, const float *d_colors, const float *d_normals, int64_t n, int64_t nvalid,
              const VoxParams &vp, int total_bits, float *d_xyz_out, float *d_colors_out, float *d_normals_out,
              int32_t *d_ijk, int32_t *d_point_voxel, int64_t *h_m)
{
    K *keys, *keys_tmp, *keys_sorted;
    int32_t *vals, *vals_tmp, *vals_sorted, *run_start;
    KP_TRY(kp_ws(ctx, (size_t)n, &keys));
    KP_TRY(kp_ws(ctx, (size_t)n, &keys_tmp));
    KP_TRY(kp_ws(ctx, (size_t)n, &vals));
    KP_TRY(kp_ws(ctx, (size_t)n, &vals_tmp));
    KP_TRY(kp_ws(ctx, (size_t)nvalid + 1, &run_start));
    {
        KP_PROFB(ctx, "voxel_keys", (double)n * (12.0 + sizeof(K)));
        k_voxel_keys<K><<<kp_blocks(n, 256), 256, 0, ctx->stream>>>(d_xyz, n, vp, keys);
        KP_LAUNCH_CHECK(ctx);
    }
    if (sizeof(K) == 8)
        KP_TRY(kp_prim_sort_pairs_u64(ctx, n, total_bits, (uint64_t *)keys, (uint64_t *)keys_tmp, vals, vals_tmp,
                                      (uint64_t **)&keys_sorted, &vals_sorted));
    else
        KP_TRY(kp_prim_sort_pairs_u32(ctx, n, total_bits, (uint32_t *)keys, (uint32_t *)keys_tmp, vals, vals_tmp,
                                      (uint32_t **)&keys_sorted, &vals_sorted));
    int32_t *d_total = (int32_t *)ctx->d_scratch;
    if (sizeof(K) == 8) KP_TRY(kp_prim_run_starts_u64(ctx, nvalid, (const uint64_t *)keys_sorted, run_start, d_total));
    else KP_TRY(kp_prim_run_starts_u32(ctx, nvalid, (const uint32_t *)keys_sorted, run_start, d_total));
    KP_TRY(kp_fetch_scratch(ctx, sizeof(int32_t)));
    int R = *(int32_t *)ctx->h_scratch;
    *h_m = R;
    if (d_point_voxel && n > nvalid) {
        k_fill_i32<<<kp_blocks(n, 256), 256, 0, ctx->stream>>>(d_point_voxel, n, -1);
        KP_LAUNCH_CHECK(ctx);
    }
    if (R > 0) {
        KP_PROFB(ctx, "voxel_mean", (double)nvalid * (4.0 + 12.0) + (double)R * (12.0 + sizeof(K) + 4.0));
        k_voxel_mean<K><<<kp_blocks(R, 128), 128, 0, ctx->stream>>>(keys_sorted, vals_sorted, run_start, R, (int)nvalid, d_xyz,
                                                                    d_colors, d_normals, vp, d_xyz_out, d_colors_out,
                                                                    d_normals_out, d_ijk, d_point_voxel);
        KP_LAUNCH_CHECK(ctx);
    }
    return KP_OK;
}
}  // namespace

int kp_voxel_device(kp_ctx *ctx, const float *d_xyz, const float *d_colors, const float *d_normals, int64_t n,
                    double voxel, const float *h_bounds6, int64_t nvalid, float *d_xyz_out, float *d_colors_out,
                    float *d_normals_out, int32_t *d_ijk, int32_t *d_point_voxel, double *h_min_bound, int64_t *h_m)
{
    *h_m = 0;
    if (!(voxel > 0.0)) return kp_set_err(ctx, KP_E_ARG, "voxel_size <= 0");
    if (n > 2147483000LL) return kp_set_err(ctx, KP_E_ARG, "more than 2^31 points in one call");
    if (nvalid <= 0) {
        if (d_point_voxel && n > 0) {
            k_fill_i32<<<kp_blocks(n, 256), 256, 0, ctx->stream>>>(d_point_voxel, n, -1);
            KP_LAUNCH_CHECK(ctx);
        }
        return KP_OK;
    }
    VoxParams vp;
    int bits[3];
    for (int c = 0; c < 3; ++c) {
        vp.minb[c] = (double)h_bounds6[c] - voxel * 0.5;
        double maxb = (double)h_bounds6[3 + c] + voxel * 0.5;
        if (voxel * 2147483647.0 < maxb - vp.minb[c]) return kp_set_err(ctx, KP_E_RANGE, "voxel_size is too small for the cloud extent");
        long long imax = (long long)floor(((double)h_bounds6[3 + c] - vp.minb[c]) / voxel);
        bits[c] = bit_length(imax);
        if (h_min_bound) h_min_bound[c] = vp.minb[c];
    }
    vp.voxel = voxel;
    vp.sh_y = bits[2];
    vp.sh_x = bits[2] + bits[1];
    int total = bits[0] + bits[1] + bits[2];
    bool has_nan = n > nvalid;
    int sort_bits = total + (has_nan ? 1 : 0);
    if (sort_bits > 64) return kp_set_err(ctx, KP_E_RANGE, "voxel grid needs %d key bits (max 64): voxel_size too small for the extent", sort_bits);
    vp.sentinel = has_nan ? (total >= 64 ? 0ull : (1ull << total)) : 0ull;
    if (sort_bits <= 32)
        return voxel_run<uint32_t>(ctx, d_xyz, d_colors, d_normals, n, nvalid, vp, sort_bits, d_xyz_out, d_colors_out,
                                   d_normals_out, d_ijk, d_point_voxel, h_m);
    return voxel_run<uint64_t>(ctx, d_xyz, d_colors, d_normals, n, nvalid, vp, sort_bits, d_xyz_out, d_colors_out,
                               d_normals_out, d_ijk, d_point_voxel, h_m);
}

extern "C" int kp_voxel_downsample(kp_ctx *ctx, const float *d_xyz, const float *d_colors, const float *d_normals,
                                   int64_t n, double voxel_size, float *d_xyz_out, float *d_colors_out,
                                   float *d_normals_out, int32_t *d_ijk, int32_t *d_point_voxel, double *h_min_bound,
                                   int64_t *h_m)
{
    if (!ctx || !h_m || (n > 0 && (!d_xyz || !d_xyz_out))) return kp_set_err(ctx, KP_E_ARG, "kp_voxel_downsample: NULL argument");
    if (!(voxel_size > 0.0)) return kp_set_err(ctx, KP_E_ARG, "voxel_size <= 0");
    kp_enter(ctx);
    float b6[6];
    int64_t nvalid = 0;
    KP_TRY(kp_prim_bounds_fetch(ctx, d_xyz, n, b6, &nvalid));
    return kp_voxel_device(ctx, d_xyz, d_colors, d_normals, n, voxel_size, b6, nvalid, d_xyz_out, d_colors_out,
                           d_normals_out, d_ijk, d_point_voxel, h_min_bound, h_m);
}
