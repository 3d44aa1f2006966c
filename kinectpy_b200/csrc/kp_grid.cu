// kp_grid.cu -- K3: uniform-grid spatial hash and the exact neighbour kernels built on it
// (kNN / hybrid, statistical outlier removal, radius outlier removal, normal estimation).
// Replaces the nanoflann KD-tree queries Open3D runs under remove_statistical_outlier
// (preprocessing/filtering.py:24, floor_removal.py:73), estimate_normals
// (preprocessing/registration.py:11-13) and remove_radius_outlier.
//
// Search scheme: one warp per query.  The query's 27-cell block is probed by 27 lanes at once
// (one hash probe each), the three cells of a z-row are merged into one contiguous range of the
// cell-sorted point array, and the warp streams those ranges 32 candidates at a time
// (one coalesced 16-byte load per lane).  Candidates that beat the current k-th entry are
// appended to a per-warp shared-memory buffer which is bitonic-sorted and truncated to k when it
// fills and at the end of every ring.  After ring r every point closer than r*cell has been seen,
// so the search stops as soon as the k-th distance is inside that radius; otherwise the next
// shell of cells is probed (isolated outliers -- the points SOR exists to find -- take this path).
#include <math.h>
#include <stdlib.h>
#include "kp_grid.cuh"

KpGridDev kp_grid_dev(const KpGrid &g)
{
    KpGridDev d;
    d.pts = g.d_sorted; d.slots = g.d_slots; d.hmask = g.hmask;
    d.sh_x = g.sh_x; d.sh_y = g.sh_y;
    for (int c = 0; c < 3; ++c) { d.dim[c] = g.dim[c]; d.org[c] = g.org[c]; }
    d.cell = g.cell; d.inv_cell = g.inv_cell;
    d.npts = g.n;
    d.bitmap = g.d_bitmap;
    return d;
}

namespace {
// ------------------------------------------------------------- build ------
template <class K>
__global__ void __launch_bounds__(256) k_grid_keys(const float *xyz, int64_t n, const __grid_constant__ KpGridDev g,
                                                   unsigned long long sentinel, K *keys)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
    K key;
    if (isnan(x)) key = (K)sentinel;
    else {
        int cx = min(max(kp_cell_coord(g, (double)x, 0), 0), g.dim[0] - 1);
        int cy = min(max(kp_cell_coord(g, (double)y, 1), 0), g.dim[1] - 1);
        int cz = min(max(kp_cell_coord(g, (double)z, 2), 0), g.dim[2] - 1);
        key = (K)kp_cell_key(g, cx, cy, cz);
    }
    keys[i] = key;
}

template <class K>
__global__ void __launch_bounds__(256) k_grid_insert(const K *keys_sorted, const int32_t *run_start, const int32_t *d_R,
                                                     int n, unsigned long long sentinel, int has_sentinel,
                                                     uint4 *slots, uint32_t hmask, uint32_t *bitmap, int sh_x, int sh_y,
                                                     int dim1, int dim2)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    int R = *d_R;
    if (r >= R) return;
    int a = run_start[r], b = (r + 1 < R) ? run_start[r + 1] : n;
    unsigned long long key = (unsigned long long)keys_sorted[a];
    if (has_sentinel && key == sentinel) return;
    if (bitmap) {
        const long long cx = (long long)(key >> sh_x), cy = (long long)((key >> sh_y) & ((1ull << (sh_x - sh_y)) - 1ull)),
                        cz = (long long)(key & ((1ull << sh_y) - 1ull));
        const long long bit = (cx * dim1 + cy) * dim2 + cz;
        atomicOr(bitmap + (bit >> 5), 1u << (bit & 31));
    }
    uint32_t h = (uint32_t)kp_mix64(key) & hmask;
    for (;;) {
        // claim the key half of the slot, then fill the value half (readers only run in later kernels)
        unsigned long long prev = atomicCAS((unsigned long long *)(slots + h), ~0ull, key);
        if (prev == ~0ull) { reinterpret_cast<int2 *>(slots + h)[1] = make_int2(a, b); return; }
        h = (h + 1) & hmask;
    }
}

__global__ void __launch_bounds__(256) k_grid_gather(const float *xyz, const int32_t *vals_sorted, int64_t n, float4 *sorted)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    int i = vals_sorted[t];
    sorted[t] = make_float4(xyz[3 * (int64_t)i], xyz[3 * (int64_t)i + 1], xyz[3 * (int64_t)i + 2], __int_as_float(i));
}

int bit_length_u(long long v)
{
    int b = 0;
    while (v > 0) { ++b; v >>= 1; }
    return b < 1 ? 1 : b;
}

template <class K>
int grid_sort_build(kp_ctx *ctx, const float *d_xyz, int64_t n, KpGrid *g, int total_bits, unsigned long long sentinel,
                    int32_t **run_start_out, int32_t **d_R_out, bool build_hash)
{
    K *keys, *keys_tmp, *keys_sorted;
    int32_t *vals, *vals_tmp, *vals_sorted, *run_start, *d_R;
    KP_TRY(kp_ws(ctx, (size_t)n, &keys));
    KP_TRY(kp_ws(ctx, (size_t)n, &keys_tmp));
    KP_TRY(kp_ws(ctx, (size_t)n, &vals));
    KP_TRY(kp_ws(ctx, (size_t)n, &vals_tmp));
    KP_TRY(kp_ws(ctx, (size_t)n + 1, &run_start));
    KP_TRY(kp_ws(ctx, 4, &d_R));
    KpGridDev gd = kp_grid_dev(*g);
    {
        KP_PROFB(ctx, "grid_keys", (double)n * (12.0 + sizeof(K)));
        k_grid_keys<K><<<kp_blocks(n, 256), 256, 0, ctx->stream>>>(d_xyz, n, gd, sentinel, keys);
        KP_LAUNCH_CHECK(ctx);
    }
    if (sizeof(K) == 8) {
        KP_TRY(kp_prim_sort_pairs_u64(ctx, n, total_bits + 1, (uint64_t *)keys, (uint64_t *)keys_tmp, vals, vals_tmp,
                                      (uint64_t **)&keys_sorted, &vals_sorted));
        KP_TRY(kp_prim_run_starts_u64(ctx, n, (const uint64_t *)keys_sorted, run_start, d_R));
    } else {
        KP_TRY(kp_prim_sort_pairs_u32(ctx, n, total_bits + 1, (uint32_t *)keys, (uint32_t *)keys_tmp, vals, vals_tmp,
                                      (uint32_t **)&keys_sorted, &vals_sorted));
        KP_TRY(kp_prim_run_starts_u32(ctx, n, (const uint32_t *)keys_sorted, run_start, d_R));
    }
    if (run_start_out) *run_start_out = run_start;
    if (d_R_out) *d_R_out = d_R;
    if (!build_hash) return KP_OK;
    KP_PROFB(ctx, "grid_hash", ((double)g->hmask + 1.0) * 16.0 + (double)n * (4.0 + 12.0 + 16.0));
    KP_CUDA(ctx, cudaMemsetAsync(g->d_slots, 0xff, sizeof(uint4) * ((size_t)g->hmask + 1), ctx->stream));
    if (g->d_bitmap) {
        size_t words = (size_t)(((long long)g->dim[0] * g->dim[1] * g->dim[2] + 31) / 32);
        KP_CUDA(ctx, cudaMemsetAsync(g->d_bitmap, 0, words * sizeof(uint32_t), ctx->stream));
    }
    k_grid_insert<K><<<kp_blocks(n, 256), 256, 0, ctx->stream>>>(keys_sorted, run_start, d_R, (int)n, sentinel, 1, g->d_slots,
                                                                 g->hmask, g->d_bitmap, g->sh_x, g->sh_y, g->dim[1],
                                                                 g->dim[2]);
    KP_LAUNCH_CHECK(ctx);
    k_grid_gather<<<kp_blocks(n, 256), 256, 0, ctx->stream>>>(d_xyz, vals_sorted, n, g->d_sorted);
    KP_LAUNCH_CHECK(ctx);
    return KP_OK;
}

int grid_layout(kp_ctx *ctx, const float *h_bounds6, double cell, KpGrid *g, int *total_bits)
{
    // grow the cell until every axis fits 21 bits
    for (;;) {
        bool ok = true;
        for (int c = 0; c < 3; ++c) {
            double ext = (double)h_bounds6[3 + c] - (double)h_bounds6[c];
            if (!(ext >= 0)) ext = 0;
            if (ext / cell > 2000000.0) ok = false;
        }
        if (ok) break;
        cell *= 2.0;
    }
    g->cell = cell;
    g->inv_cell = 1.0 / cell;
    int bits[3];
    for (int c = 0; c < 3; ++c) {
        g->org[c] = (double)h_bounds6[c];
        double ext = (double)h_bounds6[3 + c] - g->org[c];
        if (!(ext >= 0)) ext = 0;
        g->dim[c] = (int)floor(ext * g->inv_cell) + 2;   // +1 for the top cell, +1 rounding margin
        bits[c] = bit_length_u(g->dim[c] - 1);
    }
    g->sh_y = bits[2];
    g->sh_x = bits[2] + bits[1];
    *total_bits = bits[0] + bits[1] + bits[2];
    (void)ctx;
    return KP_OK;
}
}  // namespace

int kp_grid_build(kp_ctx *ctx, const float *d_xyz, int64_t n, double cell, const float *h_bounds6, KpGrid *g)
{
    *g = KpGrid();
    if (!(cell > 0.0)) return kp_set_err(ctx, KP_E_ARG, "grid cell must be > 0");
    if (n > 2147483000LL) return kp_set_err(ctx, KP_E_ARG, "more than 2^31 points in one call");
    g->n = (int32_t)n;
    if (n <= 0) return KP_OK;
    float b6[6];
    if (!h_bounds6) {
        int64_t nv = 0;
        KP_TRY(kp_prim_bounds_fetch(ctx, d_xyz, n, b6, &nv));
        if (nv == 0) { b6[0] = b6[1] = b6[2] = 0; b6[3] = b6[4] = b6[5] = 0; }
        h_bounds6 = b6;
    }
    int total_bits = 0;
    KP_TRY(grid_layout(ctx, h_bounds6, cell, g, &total_bits));
    unsigned long long sentinel = 1ull << total_bits;   // total_bits <= 63
    // hash capacity: power of two >= 2 * min(n, #cells)
    double ncell = (double)g->dim[0] * g->dim[1] * g->dim[2];
    double want = 2.0 * ((double)n < ncell ? (double)n : ncell);
    uint32_t cap = 1024;
    while ((double)cap < want) cap <<= 1;
    g->hmask = cap - 1;
    KP_TRY(kp_ws(ctx, (size_t)cap, &g->d_slots));
    KP_TRY(kp_ws(ctx, (size_t)n, &g->d_sorted));
    if (ncell <= 134217728.0)   // <= 16 MiB of occupancy bits
        KP_TRY(kp_ws(ctx, (size_t)(ncell / 32.0) + 2, &g->d_bitmap));
    if (total_bits + 1 <= 32) return grid_sort_build<uint32_t>(ctx, d_xyz, n, g, total_bits, sentinel, nullptr, nullptr, true);
    return grid_sort_build<uint64_t>(ctx, d_xyz, n, g, total_bits, sentinel, nullptr, nullptr, true);
}

int kp_grid_auto_cell(kp_ctx *ctx, const float *d_xyz, int64_t n, const float *h_bounds6, double target_per_cell,
                      double *cell_out)
{
    // Trial grid at extent/256, count occupied cells, rescale assuming the cloud is a 2-D surface
    // (occupancy ~ cell^2).  Only the cost of the search depends on this choice, never its result.
    float b6[6];
    if (!h_bounds6) {
        int64_t nv = 0;
        KP_TRY(kp_prim_bounds_fetch(ctx, d_xyz, n, b6, &nv));
        if (nv == 0) { *cell_out = 1.0; return KP_OK; }
        h_bounds6 = b6;
    }
    double e = 0;
    for (int c = 0; c < 3; ++c) e = fmax(e, (double)h_bounds6[3 + c] - (double)h_bounds6[c]);
    if (!(e > 0)) { *cell_out = 1.0; return KP_OK; }
    KpGrid g;
    g.n = (int32_t)n;
    int total_bits = 0;
    double cell0 = e / 256.0;
    KP_TRY(grid_layout(ctx, h_bounds6, cell0, &g, &total_bits));
    int32_t *d_R = nullptr;
    unsigned long long sentinel = 1ull << total_bits;
    if (total_bits + 1 <= 32) KP_TRY(grid_sort_build<uint32_t>(ctx, d_xyz, n, &g, total_bits, sentinel, nullptr, &d_R, false));
    else KP_TRY(grid_sort_build<uint64_t>(ctx, d_xyz, n, &g, total_bits, sentinel, nullptr, &d_R, false));
    KP_CUDA(ctx, cudaMemcpyAsync(ctx->d_scratch, d_R, sizeof(int32_t), cudaMemcpyDeviceToDevice, ctx->stream));
    KP_TRY(kp_fetch_scratch(ctx, sizeof(int32_t)));
    int R = *(int32_t *)ctx->h_scratch;
    double occ = R > 0 ? (double)n / (double)R : 1.0;
    double cell = g.cell * sqrt(target_per_cell / occ);
    if (cell < g.cell / 64.0) cell = g.cell / 64.0;
    if (cell > e) cell = e;
    *cell_out = cell;
    return KP_OK;
}

// ============================================================ queries =====
namespace {
constexpr int KQ_WARPS = 4;
enum { KQ_MODE_KNN = 0, KQ_MODE_RADIUS = 1, KQ_MODE_NORMALS = 2 };

struct KnnParams {
    KpGridDev g;
    const float *queries;   // NULL -> the cloud queries itself, in cell-sorted order
    int64_t nq;
    int k, cap, mode;
    double r2cap;           // > 0: only neighbours with d2 < r2cap
    int32_t *idx; double *d2; int32_t *count; double *mean;
    const float *cloud; float *normals;
    int32_t *rcount;
    const int32_t *qlist; const int32_t *qcount;   // visit only these query positions (rows of qpts)
    uint8_t *strag_flags;                          // thread kernel: flags[q] = 1 for queries it could not certify
    const float4 *qpts;                            // self-query source rows (level-0 cell-sorted array)
};

__device__ __forceinline__ bool kq_less(double d, int i, double td, int ti) { return d < td || (d == td && i < ti); }

// bitonic sort of cap (power of two) entries by (d2, idx), one warp, shared memory
__device__ void kq_sort(double *bd, int *bi, int cap, int lane)
{
    for (int size = 2; size <= cap; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = lane; t < (cap >> 1); t += 32) {
                int i = ((t / stride) * 2 * stride) + (t % stride);
                int j = i + stride;
                bool up = (i & size) == 0;
                double di = bd[i], dj = bd[j];
                int ii = bi[i], ij = bi[j];
                bool sw = kq_less(dj, ij, di, ii) == up;
                if (sw) { bd[i] = dj; bd[j] = di; bi[i] = ij; bi[j] = ii; }
            }
            __syncwarp();
        }
    }
}

struct KqState {
    double *bd; int *bi;
    int n_buf, k, cap;
    double tau_d; int tau_i;
    double qx, qy, qz;
    int lane;
    int rcount;
};

__device__ __forceinline__ void kq_truncate(KqState &s)
{
    for (int t = s.n_buf + s.lane; t < s.cap; t += 32) { s.bd[t] = INFINITY; s.bi[t] = 0x7fffffff; }
    __syncwarp();
    kq_sort(s.bd, s.bi, s.cap, s.lane);
    if (s.n_buf > s.k) s.n_buf = s.k;
    if (s.n_buf == s.k) { s.tau_d = s.bd[s.k - 1]; s.tau_i = s.bi[s.k - 1]; }
    __syncwarp();
}

// ---- buffer maintenance without sorting: the buffer always holds EVERY candidate seen so far that is
// below the bound tau.  When it fills, tau is lowered to a sampled pivot that still has >= k entries at or
// below it (one counting pass + one in-place compaction); the exact order is only established once, by the
// final kq_truncate.  (A bitonic sort per overflow, the first version, cost more than the search itself.)
__device__ __forceinline__ int kq_warp_sum(int c)
{
#pragma unroll
    for (int sft = 16; sft >= 1; sft >>= 1) c += __shfl_xor_sync(KP_FULL, c, sft);
    return c;
}
__device__ __forceinline__ int kq_count_le(const KqState &s, double pd, int pi)
{
    int c = 0;
    for (int t = s.lane; t < s.n_buf; t += 32) c += !kq_less(pd, pi, s.bd[t], s.bi[t]);
    return kq_warp_sum(c);
}
__device__ __forceinline__ int kq_count_d_le(const KqState &s, double lim)
{
    int c = 0;
    for (int t = s.lane; t < s.n_buf; t += 32) c += s.bd[t] <= lim;
    return kq_warp_sum(c);
}
__device__ __forceinline__ void kq_keep_le(KqState &s, double pd, int pi)
{
    int out = 0;
    const unsigned lt = (1u << s.lane) - 1u;
    for (int base = 0; base < s.n_buf; base += 32) {
        const int t = base + s.lane;
        const bool v = t < s.n_buf;
        const double d = v ? s.bd[t] : 0.0;
        const int i = v ? s.bi[t] : 0;
        const bool keep = v && !kq_less(pd, pi, d, i);
        const unsigned m = __ballot_sync(KP_FULL, keep);
        __syncwarp();   // the whole chunk is in registers before anything (at an index <= t) is overwritten
        if (keep) { int pos = out + __popc(m & lt); s.bd[pos] = d; s.bi[pos] = i; }
        out += __popc(m);
        __syncwarp();
    }
    s.n_buf = out;
}
__device__ void kq_tighten(KqState &s)
{
    // 32 evenly spaced samples, sorted across the lanes (bitonic network on shuffles)
    const int si = (int)(((long long)s.lane * s.n_buf) >> 5);
    double sd = s.bd[si];
    int sid = s.bi[si];
#pragma unroll
    for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            const double od = __shfl_xor_sync(KP_FULL, sd, stride);
            const int oi = __shfl_xor_sync(KP_FULL, sid, stride);
            const bool keep_min = ((s.lane & stride) == 0) == ((s.lane & size) == 0);
            const bool other_less = kq_less(od, oi, sd, sid);
            if (keep_min ? other_less : kq_less(sd, sid, od, oi)) { sd = od; sid = oi; }
        }
    }
    // lowest sample quantile expected to keep ~1.5 k entries; verify by counting, move up if it keeps too few
    int j = (int)((48LL * s.k) / s.n_buf);
    if (j > 31) j = 31;
    for (int tries = 0; tries < 6; ++tries) {
        const double pd = __shfl_sync(KP_FULL, sd, j);
        const int pi = __shfl_sync(KP_FULL, sid, j);
        const int c = kq_count_le(s, pd, pi);
        if (c >= s.k) {
            if (c < s.n_buf) { kq_keep_le(s, pd, pi); s.tau_d = pd; s.tau_i = pi; }
            break;
        }
        if (j == 31) break;
        j = min(31, 2 * j + 1);
    }
    if (s.n_buf > s.cap - 32) kq_truncate(s);   // unlucky samples: fall back to the exact sort
}

__device__ __forceinline__ void kq_candidate(KqState &s, bool valid, float4 p, int mode)
{
    double d = kp_d2(s.qx, s.qy, s.qz, (double)p.x, (double)p.y, (double)p.z);
    int id = __float_as_int(p.w);
    if (mode == KQ_MODE_RADIUS) {
        s.rcount += __popc(__ballot_sync(KP_FULL, valid && d < s.tau_d));
        return;
    }
    bool pass = valid && kq_less(d, id, s.tau_d, s.tau_i);
    unsigned m = __ballot_sync(KP_FULL, pass);
    if (m == 0) return;
    if (pass) {
        int pos = s.n_buf + __popc(m & ((1u << s.lane) - 1u));
        s.bd[pos] = d; s.bi[pos] = id;
    }
    s.n_buf += __popc(m);
    __syncwarp();
    if (s.n_buf > s.cap - 32) kq_tighten(s);
}

__device__ __forceinline__ void kq_scan_ranges(const KpGridDev &g, int rs, int re, KqState &s, int mode)
{
    unsigned m = __ballot_sync(KP_FULL, re > rs);
    while (m) {
        int j = __ffs(m) - 1;
        m &= m - 1;
        int a = __shfl_sync(KP_FULL, rs, j), b = __shfl_sync(KP_FULL, re, j);
        for (int t = a; t < b; t += 32) {
            bool v = t + s.lane < b;
            float4 p = v ? __ldg(g.pts + t + s.lane) : make_float4(0.f, 0.f, 0.f, 0.f);
            kq_candidate(s, v, p, mode);
        }
    }
}

// analytic smallest-eigenvector of a symmetric 3x3; A = {xx,xy,xz,yy,yz,zz}
__device__ void kq_cross(const double *a, const double *b, double *o)
{
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ void kq_evec0(const double *A, double ev, double *out)
{
    double r0[3] = {A[0] - ev, A[1], A[2]}, r1[3] = {A[1], A[3] - ev, A[4]}, r2[3] = {A[2], A[4], A[5] - ev};
    double c01[3], c02[3], c12[3];
    kq_cross(r0, r1, c01); kq_cross(r0, r2, c02); kq_cross(r1, r2, c12);
    double d0 = c01[0] * c01[0] + c01[1] * c01[1] + c01[2] * c01[2];
    double d1 = c02[0] * c02[0] + c02[1] * c02[1] + c02[2] * c02[2];
    double d2 = c12[0] * c12[0] + c12[1] * c12[1] + c12[2] * c12[2];
    double dm = d0; double b0 = c01[0], b1 = c01[1], b2 = c01[2];
    if (d1 > dm) { dm = d1; b0 = c02[0]; b1 = c02[1]; b2 = c02[2]; }
    if (d2 > dm) { dm = d2; b0 = c12[0]; b1 = c12[1]; b2 = c12[2]; }
    if (dm > 0) { double sq = sqrt(dm); out[0] = b0 / sq; out[1] = b1 / sq; out[2] = b2 / sq; }
    else { out[0] = out[1] = out[2] = 0.0; }
}
__device__ void kq_evec1(const double *A, const double *e0, double ev1, double *out)
{
    double U[3], V[3];
    if (fabs(e0[0]) > fabs(e0[1])) { double il = 1.0 / sqrt(e0[0] * e0[0] + e0[2] * e0[2]); U[0] = -e0[2] * il; U[1] = 0; U[2] = e0[0] * il; }
    else { double il = 1.0 / sqrt(e0[1] * e0[1] + e0[2] * e0[2]); U[0] = 0; U[1] = e0[2] * il; U[2] = -e0[1] * il; }
    kq_cross(e0, U, V);
    double AU[3] = {A[0] * U[0] + A[1] * U[1] + A[2] * U[2], A[1] * U[0] + A[3] * U[1] + A[4] * U[2], A[2] * U[0] + A[4] * U[1] + A[5] * U[2]};
    double AV[3] = {A[0] * V[0] + A[1] * V[1] + A[2] * V[2], A[1] * V[0] + A[3] * V[1] + A[4] * V[2], A[2] * V[0] + A[4] * V[1] + A[5] * V[2]};
    double m00 = U[0] * AU[0] + U[1] * AU[1] + U[2] * AU[2] - ev1;
    double m01 = U[0] * AV[0] + U[1] * AV[1] + U[2] * AV[2];
    double m11 = V[0] * AV[0] + V[1] * AV[1] + V[2] * AV[2] - ev1;
    double a00 = fabs(m00), a01 = fabs(m01), a11 = fabs(m11);
    double cu, cv;
    if (a00 >= a11) {
        if (fmax(a00, a01) > 0) {
            if (a00 >= a01) { m01 /= m00; m00 = 1 / sqrt(1 + m01 * m01); m01 *= m00; }
            else { m00 /= m01; m01 = 1 / sqrt(1 + m00 * m00); m00 *= m01; }
            cu = m01; cv = -m00;
        } else { cu = 1; cv = 0; }
    } else {
        if (fmax(a11, a01) > 0) {
            if (a11 >= a01) { m01 /= m11; m11 = 1 / sqrt(1 + m01 * m01); m01 *= m11; }
            else { m11 /= m01; m01 = 1 / sqrt(1 + m11 * m11); m11 *= m01; }
            cu = m11; cv = -m01;
        } else { cu = 1; cv = 0; }
    }
    for (int c = 0; c < 3; ++c) out[c] = cu * U[c] + cv * V[c];
}
__device__ void kq_smallest_eigvec(const double *cov, double *nrm)
{
    double A[6];
    double mc = cov[0];
    for (int i = 1; i < 6; ++i) if (cov[i] > mc) mc = cov[i];
    if (mc == 0.0) { nrm[0] = nrm[1] = nrm[2] = 0.0; return; }
    for (int i = 0; i < 6; ++i) A[i] = cov[i] / mc;
    double norm = A[1] * A[1] + A[2] * A[2] + A[4] * A[4];
    if (norm > 0) {
        double q = (A[0] + A[3] + A[5]) / 3.0;
        double b00 = A[0] - q, b11 = A[3] - q, b22 = A[5] - q;
        double p = sqrt((b00 * b00 + b11 * b11 + b22 * b22 + norm * 2.0) / 6.0);
        double c00 = b11 * b22 - A[4] * A[4];
        double c01 = A[1] * b22 - A[4] * A[2];
        double c02 = A[1] * A[4] - b11 * A[2];
        double det = (b00 * c00 - A[1] * c01 + A[2] * c02) / (p * p * p);
        double half = det * 0.5;
        half = fmin(fmax(half, -1.0), 1.0);
        double angle = acos(half) / 3.0;
        const double two_thirds_pi = 2.09439510239319549;
        double beta2 = cos(angle) * 2.0;
        double beta0 = cos(angle + two_thirds_pi) * 2.0;
        double beta1 = -(beta0 + beta2);
        double e0 = q + p * beta0, e1 = q + p * beta1, e2 = q + p * beta2;
        double v0[3], v1[3], v2[3];
        if (half >= 0) {
            kq_evec0(A, e2, v2);
            kq_evec1(A, v2, e1, v1);
            kq_cross(v1, v2, v0);
        } else {
            kq_evec0(A, e0, v0);
        }
        nrm[0] = v0[0]; nrm[1] = v0[1]; nrm[2] = v0[2];
    } else {
        if (A[0] < A[3] && A[0] < A[5]) { nrm[0] = 1; nrm[1] = 0; nrm[2] = 0; }
        else if (A[3] < A[0] && A[3] < A[5]) { nrm[0] = 0; nrm[1] = 1; nrm[2] = 0; }
        else { nrm[0] = 0; nrm[1] = 0; nrm[2] = 1; }
    }
}

// Results of one query, written by ONE thread from its list sorted ascending by (d2, index)
// (entry t at bd[t*stride], bi[t*stride]).  Sums run sequentially in that order -- the order Open3D's
// std::accumulate sees (the KD-tree returns neighbours by ascending distance) and the oracle uses.
__device__ void kq_finalize(const KnnParams &p, int64_t row, int cnt, const double *bd, const int *bi, int stride)
{
    if (p.mode == KQ_MODE_KNN) {
        if (p.idx) for (int t = 0; t < p.k; ++t) p.idx[row * p.k + t] = t < cnt ? bi[t * stride] : -1;
        if (p.d2) for (int t = 0; t < p.k; ++t) p.d2[row * p.k + t] = t < cnt ? bd[t * stride] : INFINITY;
        if (p.count) p.count[row] = cnt;
        if (p.mean) {
            double acc = 0.0;
            for (int t = 0; t < cnt; ++t) acc = __dadd_rn(acc, sqrt(bd[t * stride]));
            p.mean[row] = cnt > 0 ? __ddiv_rn(acc, (double)cnt) : -1.0;
        }
        return;
    }
    // normals: covariance from cumulants over the neighbourhood, smallest eigenvector
    double nr[3] = {0.0, 0.0, 1.0};
    if (cnt >= 3) {
        double sm[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int t = 0; t < cnt; ++t) {
            int64_t j = bi[t * stride];
            double x = (double)p.cloud[3 * j], y = (double)p.cloud[3 * j + 1], z = (double)p.cloud[3 * j + 2];
            sm[0] += x; sm[1] += y; sm[2] += z;
            sm[3] += x * x; sm[4] += x * y; sm[5] += x * z; sm[6] += y * y; sm[7] += y * z; sm[8] += z * z;
        }
        double inv = (double)cnt;
        for (int c = 0; c < 9; ++c) sm[c] /= inv;
        double cov[6] = {sm[3] - sm[0] * sm[0], sm[4] - sm[0] * sm[1], sm[5] - sm[0] * sm[2],
                         sm[6] - sm[1] * sm[1], sm[7] - sm[1] * sm[2], sm[8] - sm[2] * sm[2]};
        kq_smallest_eigvec(cov, nr);
        if (nr[0] == 0.0 && nr[1] == 0.0 && nr[2] == 0.0) nr[2] = 1.0;
    }
    p.normals[3 * row] = (float)nr[0]; p.normals[3 * row + 1] = (float)nr[1]; p.normals[3 * row + 2] = (float)nr[2];
}

// ---- warp-per-query kernel: any k, ring expansion, linear-scan fallback.  Used for large k, for external
// queries and for the stragglers the thread-per-query kernel hands over (qlist / qcount).
__global__ void __launch_bounds__(KQ_WARPS * 32) k_knn(const __grid_constant__ KnnParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const KpGridDev &g = p.g;
    const int64_t nq = p.qcount ? (int64_t)*p.qcount : p.nq;
    for (int64_t w = (int64_t)blockIdx.x * KQ_WARPS + warp; w < nq; w += (int64_t)gridDim.x * KQ_WARPS) {
        const int64_t q = p.qlist ? (int64_t)p.qlist[w] : w;
        KqState s;
        s.bd = reinterpret_cast<double *>(smem_raw) + (size_t)warp * p.cap;
        s.bi = reinterpret_cast<int *>(reinterpret_cast<double *>(smem_raw) + (size_t)KQ_WARPS * p.cap) + (size_t)warp * p.cap;
        s.n_buf = 0; s.k = p.k; s.cap = p.cap; s.lane = lane; s.rcount = 0;
        s.tau_d = p.r2cap > 0 ? p.r2cap : INFINITY;
        s.tau_i = p.r2cap > 0 ? (int)0x80000000 : 0x7fffffff;   // strict d2 < r2cap: equality never passes

        int64_t row;
        if (p.queries) {
            s.qx = (double)p.queries[3 * q]; s.qy = (double)p.queries[3 * q + 1]; s.qz = (double)p.queries[3 * q + 2];
            row = q;
        } else {
            float4 me = __ldg(p.qpts + q);
            s.qx = (double)me.x; s.qy = (double)me.y; s.qz = (double)me.z;
            row = __float_as_int(me.w);
        }
        const bool qnan = isnan(s.qx);
        if (!qnan && g.dim[0] > 0) {
            // Rings are centred on the in-grid cell nearest to the query.  For a query inside the grid that is
            // its own cell; for an outside query, a grid point within distance rho of the query still lies
            // within ceil(rho/cell) cells of the clamped cell on every axis, so the ring-r guarantee
            // ("everything closer than r*cell has been seen") holds unchanged.
            const int cx = min(max(kp_cell_coord(g, s.qx, 0), 0), g.dim[0] - 1);
            const int cy = min(max(kp_cell_coord(g, s.qy, 1), 0), g.dim[1] - 1);
            const int cz = min(max(kp_cell_coord(g, s.qz, 2), 0), g.dim[2] - 1);
            // rings needed to cover the whole grid from this cell
            int maxring = max(max(max(cx, g.dim[0] - 1 - cx), max(cy, g.dim[1] - 1 - cy)), max(cz, g.dim[2] - 1 - cz));
            if (maxring < 1) maxring = 1;
            // ---- ring 0+1: the 27-cell block, z-rows merged
            {
                int2 r = make_int2(0, 0);
                if (lane < 27) r = kp_cell_range(g, cx + lane / 9 - 1, cy + (lane / 3) % 3 - 1, cz + lane % 3 - 1);
                bool ne = r.y > r.x;
                int a = ne ? r.x : 0x7fffffff, b = ne ? r.y : 0;
                int a1 = __shfl_down_sync(KP_FULL, a, 1), b1 = __shfl_down_sync(KP_FULL, b, 1);
                int a2 = __shfl_down_sync(KP_FULL, a, 2), b2 = __shfl_down_sync(KP_FULL, b, 2);
                if (lane < 27 && lane % 3 == 0) { a = min(a, min(a1, a2)); b = max(b, max(b1, b2)); if (b == 0) a = 0; }
                else { a = 0; b = 0; }
                kq_scan_ranges(g, a, b, s, p.mode);
            }
            int ring = 1;
            if (p.mode != KQ_MODE_RADIUS) {
                for (;;) {
                    double safe = (double)ring * g.cell * (1.0 - 1.0 / 1048576.0);
                    double s2 = safe * safe;
                    // certified as soon as k of the candidates seen so far lie inside the scanned radius
                    if (s.n_buf >= s.k && kq_count_d_le(s, s2) >= s.k) break;
                    if (p.r2cap > 0 && p.r2cap <= s2) break;
                    if (ring >= maxring) break;
                    ++ring;
                    {
                        // A shell this large costs more probes than the cloud has points: finish with one exact
                        // linear pass over the sorted array instead (bounds the work of far-away stragglers).
                        const long long w3 = 2LL * ring + 1;
                        if (w3 * w3 * w3 > (long long)g.npts) {
                            s.n_buf = 0;
                            s.tau_d = p.r2cap > 0 ? p.r2cap : INFINITY;
                            s.tau_i = p.r2cap > 0 ? (int)0x80000000 : 0x7fffffff;
                            kq_scan_ranges(g, 0, lane == 0 ? g.npts : 0, s, p.mode);
                            break;
                        }
                    }
                    // ---- shell `ring`: two full slabs dz = +-ring, then the perimeter of every layer in between
                    const long long sw = 2LL * ring + 1;
                    const long long slab = sw * sw, per = 4 * (sw - 1);
                    const long long ncell = 2 * slab + (sw - 2) * per;
                    for (long long e0 = 0; e0 < ncell; e0 += 32) {
                        long long e = e0 + lane;
                        int2 r = make_int2(0, 0);
                        if (e < ncell) {
                            int dx, dy, dz;
                            if (e < 2 * slab) {
                                int sl = (int)(e / slab);
                                long long rem = e % slab;
                                dx = (int)(rem / sw) - ring; dy = (int)(rem % sw) - ring; dz = sl ? ring : -ring;
                            } else {
                                long long e2 = e - 2 * slab;
                                int layer = (int)(e2 / per), pi = (int)(e2 % per);
                                dz = -ring + 1 + layer;
                                int side = pi / (int)(sw - 1), t = pi % (int)(sw - 1);
                                if (side == 0) { dx = -ring + t; dy = -ring; }
                                else if (side == 1) { dx = ring; dy = -ring + t; }
                                else if (side == 2) { dx = ring - t; dy = ring; }
                                else { dx = -ring; dy = ring - t; }
                            }
                            r = kp_cell_range(g, cx + dx, cy + dy, cz + dz);
                        }
                        kq_scan_ranges(g, r.x, r.y, s, p.mode);
                    }
                }
            }
        }
        if (p.mode == KQ_MODE_RADIUS) {
            if (lane == 0) p.rcount[row] = s.rcount;
        } else {
            kq_truncate(s);   // the one exact sort: ascending (d2, index), cut to k
            if (lane == 0) kq_finalize(p, row, qnan ? 0 : s.n_buf, s.bd, s.bi, 1);
        }
        __syncwarp();
    }
}

// ---- thread-per-query kernel (the fast path, k <= TQ_KMAX, the cloud queries itself).
// Queries are taken in cell-sorted order, so the lanes of a warp sit in the same or adjacent cells and
// walk (almost) the same candidate ranges: their 16-byte candidate loads hit the same L1 lines, which is
// the shared staging of the cell neighbourhood.  Each thread keeps a max-heap of its k best (d2, index)
// pairs in shared memory, laid out [slot][thread] so that any slot pattern is bank-conflict free.
// Rows of cells whose nearest face is already farther than the current k-th distance are skipped.
// A query whose k-th distance is not certified by the 27-cell block (isolated points: what SOR is
// looking for) is appended to a straggler list and finished by the warp kernel above.
constexpr int TQ_THREADS = 128;
constexpr int TQ_KMAX = 64;

__device__ __forceinline__ bool tq_greater(double d, int i, double e, int j) { return d > e || (d == e && i > j); }

__global__ void __launch_bounds__(TQ_THREADS) k_knn_tq(const __grid_constant__ KnnParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const KpGridDev &g = p.g;
    const int tid = threadIdx.x;
    const int64_t w = (int64_t)blockIdx.x * TQ_THREADS + tid;
    if (w >= p.nq) return;
    const int64_t q = p.qlist ? (int64_t)p.qlist[w] : w;
    double *hd = reinterpret_cast<double *>(smem_raw) + tid;                                   // hd[slot * TQ_THREADS]
    int *hi = reinterpret_cast<int *>(reinterpret_cast<double *>(smem_raw) + (size_t)p.k * TQ_THREADS) + tid;
    const int k = p.k;
    const float4 me = __ldg(p.qpts + q);
    const int64_t row = __float_as_int(me.w);
    const double qx = (double)me.x, qy = (double)me.y, qz = (double)me.z;
    if (isnan(qx)) { kq_finalize(p, row, 0, hd, hi, TQ_THREADS); return; }
    const int cx = kp_cell_coord(g, qx, 0), cy = kp_cell_coord(g, qy, 1), cz = kp_cell_coord(g, qz, 2);
    // distance from the query to the faces of its own cell, per axis (low, high), slightly shrunk so
    // that rounding in the cell assignment can never make a pruned row hold a closer point
    const double shrink = 1.0 - 1.0 / 1048576.0;
    double glo[3], ghi[3];
    {
        const double qq[3] = {qx, qy, qz};
        const int cc[3] = {cx, cy, cz};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            double f = (qq[c] - g.org[c]) * g.inv_cell - (double)cc[c];
            f = fmin(fmax(f, 0.0), 1.0);
            glo[c] = f * g.cell * shrink;
            ghi[c] = (1.0 - f) * g.cell * shrink;
        }
    }
    int n = 0;
    const bool capped = p.r2cap > 0;
    // the heap's root (current k-th best) is mirrored in registers: the common "reject" test touches no memory
    double topd = INFINITY;
    int topi = 0x7fffffff;
    auto offer = [&](double d, int id) {
        if (capped && !(d < p.r2cap)) return;
        if (n < k) {
            int pos = n++;   // push: sift up
            while (pos > 0) {
                int par = (pos - 1) >> 1;
                double pd = hd[par * TQ_THREADS]; int pi = hi[par * TQ_THREADS];
                if (!tq_greater(d, id, pd, pi)) break;
                hd[pos * TQ_THREADS] = pd; hi[pos * TQ_THREADS] = pi;
                pos = par;
            }
            hd[pos * TQ_THREADS] = d; hi[pos * TQ_THREADS] = id;
            if (n == k) { topd = hd[0]; topi = hi[0]; }
        } else if (tq_greater(topd, topi, d, id)) {
            int pos = 0;     // replace the current worst: sift down
            for (;;) {
                int ch = 2 * pos + 1;
                if (ch >= k) break;
                double cd = hd[ch * TQ_THREADS]; int ci = hi[ch * TQ_THREADS];
                if (ch + 1 < k) {
                    double ed = hd[(ch + 1) * TQ_THREADS]; int ei = hi[(ch + 1) * TQ_THREADS];
                    if (tq_greater(ed, ei, cd, ci)) { cd = ed; ci = ei; ++ch; }
                }
                if (!tq_greater(cd, ci, d, id)) break;
                hd[pos * TQ_THREADS] = cd; hi[pos * TQ_THREADS] = ci;
                pos = ch;
            }
            hd[pos * TQ_THREADS] = d; hi[pos * TQ_THREADS] = id;
            topd = hd[0]; topi = hi[0];
        }
    };
    // rows ordered so the query's own row comes first (tightens the k-th distance early)
    const int order[9] = {4, 1, 3, 5, 7, 0, 2, 6, 8};
    for (int oi = 0; oi < 9; ++oi) {
        const int dx = order[oi] / 3 - 1, dy = order[oi] % 3 - 1;
        if (n == k || capped) {
            const double gx = dx < 0 ? glo[0] : (dx > 0 ? ghi[0] : 0.0), gy = dy < 0 ? glo[1] : (dy > 0 ? ghi[1] : 0.0);
            const double m2 = gx * gx + gy * gy;
            if (n == k && m2 > topd) continue;
            if (capped && m2 >= p.r2cap) continue;
        }
        const int2 rr = kp_row_range(g, cx + dx, cy + dy, cz);
        const int a = rr.x, b = rr.y;
        // four candidates per trip: the loads and the four distance chains are independent, only the
        // offers are sequential (keeps the FP64 pipe busy at the low occupancy a per-thread heap allows)
        for (int t = a; t < b; t += 4) {
            const int m = b - t;
            const float4 c0 = __ldg(g.pts + t);
            const float4 c1 = __ldg(g.pts + (m > 1 ? t + 1 : t));
            const float4 c2 = __ldg(g.pts + (m > 2 ? t + 2 : t));
            const float4 c3 = __ldg(g.pts + (m > 3 ? t + 3 : t));
            const double d0 = kp_d2(qx, qy, qz, (double)c0.x, (double)c0.y, (double)c0.z);
            const double d1 = kp_d2(qx, qy, qz, (double)c1.x, (double)c1.y, (double)c1.z);
            const double d2 = kp_d2(qx, qy, qz, (double)c2.x, (double)c2.y, (double)c2.z);
            const double d3 = kp_d2(qx, qy, qz, (double)c3.x, (double)c3.y, (double)c3.z);
            offer(d0, __float_as_int(c0.w));
            if (m > 1) offer(d1, __float_as_int(c1.w));
            if (m > 2) offer(d2, __float_as_int(c2.w));
            if (m > 3) offer(d3, __float_as_int(c3.w));
        }
    }
    // certificate: every point closer than (cell + distance to the nearest face of the own cell) was seen
    if (!capped || p.r2cap > g.cell * g.cell * shrink * shrink) {
        double mg = fmin(fmin(fmin(glo[0], ghi[0]), fmin(glo[1], ghi[1])), fmin(glo[2], ghi[2]));
        double safe = g.cell * shrink + mg;
        double s2 = safe * safe;
        bool exact = (n == k && topd <= s2) || (capped && p.r2cap <= s2);
        if (!exact) {
            p.strag_flags[q] = 1;   // flag, not append: the list is compacted in query order so that the
            return;                 // next level's warps again hold spatial neighbours
        }
    }
    // heapsort in place -> ascending (d2, index)
    for (int end = n - 1; end > 0; --end) {
        const double d = hd[end * TQ_THREADS]; const int id = hi[end * TQ_THREADS];
        hd[end * TQ_THREADS] = hd[0]; hi[end * TQ_THREADS] = hi[0];
        int pos = 0;
        for (;;) {
            int ch = 2 * pos + 1;
            if (ch >= end) break;
            double cd = hd[ch * TQ_THREADS]; int ci = hi[ch * TQ_THREADS];
            if (ch + 1 < end) {
                double ed = hd[(ch + 1) * TQ_THREADS]; int ei = hi[(ch + 1) * TQ_THREADS];
                if (tq_greater(ed, ei, cd, ci)) { cd = ed; ci = ei; ++ch; }
            }
            if (!tq_greater(cd, ci, d, id)) break;
            hd[pos * TQ_THREADS] = cd; hi[pos * TQ_THREADS] = ci;
            pos = ch;
        }
        hd[pos * TQ_THREADS] = d; hi[pos * TQ_THREADS] = id;
    }
    kq_finalize(p, row, n, hd, hi, TQ_THREADS);
}

int next_pow2(int v)
{
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

int knn_launch_warp(kp_ctx *ctx, KnnParams &p, int64_t grid_queries)
{
    p.cap = next_pow2(4 * p.k > 128 ? 4 * p.k : 128);
    if ((size_t)KQ_WARPS * p.cap * 12 > 96 * 1024) p.cap = next_pow2(p.k + 64 > 128 ? p.k + 64 : 128);
    size_t smem = (size_t)KQ_WARPS * p.cap * (sizeof(double) + sizeof(int));
    if (smem > 200 * 1024) return kp_set_err(ctx, KP_E_ARG, "k = %d neighbours is too many for the per-warp buffer", p.k);
    if (smem > 48 * 1024) KP_CUDA(ctx, cudaFuncSetAttribute(k_knn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t blocks = (grid_queries + KQ_WARPS - 1) / KQ_WARPS;
    int64_t cap_blocks = (int64_t)ctx->sm_count * 32;
    if (p.qlist && blocks > cap_blocks) blocks = cap_blocks;     // persistent loop over the straggler list
    if (blocks < 1) blocks = 1;
    k_knn<<<(unsigned)blocks, KQ_WARPS * 32, smem, ctx->stream>>>(p);
    KP_LAUNCH_CHECK(ctx);
    return KP_OK;
}

// d_xyz: the cloud the grid was built from (needed to build the coarser cascade levels; may be NULL)
int knn_launch(kp_ctx *ctx, KnnParams &p, const char *name, const float *d_xyz = nullptr)
{
    if (p.nq <= 0) return KP_OK;
    // compulsory HBM traffic: the cell-sorted float4 array once + the per-query outputs
    double out_b = p.mode == KQ_MODE_RADIUS ? 4.0 : p.mode == KQ_MODE_NORMALS ? 12.0 + 12.0
                   : (p.idx ? 4.0 * p.k : 0.0) + (p.d2 ? 8.0 * p.k : 0.0) + (p.count ? 4.0 : 0.0) + (p.mean ? 8.0 : 0.0);
    KP_PROFB(ctx, name, (double)p.g.npts * 16.0 + (double)p.nq * (out_b + (p.queries ? 12.0 : 0.0)));
    p.qlist = nullptr; p.qcount = nullptr; p.strag_flags = nullptr;
    p.qpts = p.g.pts;
    const bool fast = !p.queries && p.mode != KQ_MODE_RADIUS && p.k <= TQ_KMAX;
    if (!fast) return knn_launch_warp(ctx, p, p.nq);
    const KpGridDev g0 = p.g;
    const int64_t nq0 = p.nq;
    int32_t *listA, *counts;
    uint8_t *flags;
    KP_TRY(kp_ws(ctx, (size_t)nq0, &listA));
    KP_TRY(kp_ws(ctx, (size_t)nq0, &flags));
    KP_TRY(kp_ws(ctx, 8, &counts));
    KP_CUDA(ctx, cudaMemsetAsync(counts, 0, 8 * sizeof(int32_t), ctx->stream));
    KP_CUDA(ctx, cudaMemsetAsync(flags, 0, (size_t)nq0, ctx->stream));
    size_t smem = (size_t)p.k * TQ_THREADS * (sizeof(double) + sizeof(int));
    if (smem > 48 * 1024) KP_CUDA(ctx, cudaFuncSetAttribute(k_knn_tq, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // level 0: every query against the caller's grid
    p.strag_flags = flags;
    {
        KP_PROF(ctx, "knn_level0");
        k_knn_tq<<<kp_blocks(nq0, TQ_THREADS), TQ_THREADS, smem, ctx->stream>>>(p);
        KP_LAUNCH_CHECK(ctx);
    }
    // A radius-capped search on a grid whose cell covers the radius is always certified: nothing to hand over.
    const double shrink = 1.0 - 1.0 / 1048576.0;   // same test as the kernel's certificate
    const bool always_exact = p.r2cap > 0 && !(p.r2cap > g0.cell * g0.cell * shrink * shrink);
    if (always_exact) return KP_OK;
    // Queries the 27-cell block could not certify: flying pixels and other isolated points (exactly what SOR is
    // looking for) plus the sparsest fringes of the cloud.  Their k-th neighbour can be tens of cells away, so
    // they are finished by the ring-expanding warp kernel on a much COARSER grid (few rings, long contiguous
    // candidate runs that a warp streams 32 at a time).  The list is compacted in query order.
    // (sweep in profiles/r01_c_knn_base_coarse_sweep.log: a 4x coarser grid halves the straggler pass for
    // k = 20; for k = 50 the per-warp buffer sorts dominate and the level-0 grid is as good)
    const double coarse_mult = getenv("KP_KNN_COARSE_MULT") ? atof(getenv("KP_KNN_COARSE_MULT")) : (p.k <= 32 ? 4.0 : 1.0);
    KP_TRY(kp_prim_compact_mask(ctx, nq0, flags, 0, nullptr, nullptr, listA, counts));
    KP_CUDA(ctx, cudaMemcpyAsync(ctx->d_scratch, counts, sizeof(int32_t), cudaMemcpyDeviceToDevice, ctx->stream));
    KP_TRY(kp_fetch_scratch(ctx, sizeof(int32_t)));
    const int32_t n_strag = *(int32_t *)ctx->h_scratch;
    if (getenv("KP_DEBUG_KNN"))
        fprintf(stderr, "[kp knn] %s k=%d n=%lld cell=%g uncertified=%d\n", name, p.k, (long long)nq0, g0.cell, n_strag);
    if (n_strag <= 0) return KP_OK;
    p.g = g0;
    if (d_xyz && coarse_mult > 1.0) {
        float b6[6];
        for (int c = 0; c < 3; ++c) {
            b6[c] = (float)g0.org[c];
            b6[3 + c] = (float)(g0.org[c] + g0.cell * (double)g0.dim[c]);
        }
        KpGrid gl;
        KP_TRY(kp_grid_build(ctx, d_xyz, g0.npts, g0.cell * coarse_mult, b6, &gl));
        p.g = kp_grid_dev(gl);
    }
    p.qpts = g0.pts;
    p.qlist = listA;
    p.qcount = counts;
    p.nq = nq0;
    {
        KP_PROF(ctx, "knn_stragglers");
        return knn_launch_warp(ctx, p, n_strag);
    }
}

// SOR statistics: the two canonical sums run on transformed copies of the mean array
__global__ void k_sor_pos(const double *mean, int64_t n, double *out)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { double m = mean[i]; out[i] = m > 0 ? m : 0.0; }
}
__global__ void k_sor_sq(const double *mean, int64_t n, const double *sum, double valid, double *out, double *mu_out)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double mu = __ddiv_rn(*sum, valid);
    if (i == 0) *mu_out = mu;
    if (i < n) { double m = mean[i]; double d = __dsub_rn(m, mu); out[i] = m > 0 ? __dmul_rn(d, d) : 0.0; }
}
__global__ void k_sor_mask(const double *mean, int64_t n, const double *mu_p, const double *sq_p, double valid, double ratio,
                           uint8_t *keep, double *stats)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double mu = *mu_p;
    double sd = sqrt(__ddiv_rn(*sq_p, __dsub_rn(valid, 1.0)));
    double thr = __dadd_rn(mu, __dmul_rn(ratio, sd));
    if (i == 0) { stats[0] = mu; stats[1] = sd; stats[2] = thr; }
    if (i < n) { double m = mean[i]; keep[i] = (m > 0 && m < thr) ? 1 : 0; }
}
__global__ void k_radius_mask(const int32_t *cnt, int64_t n, int nb, uint8_t *keep)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keep[i] = cnt[i] > nb;
}
}  // namespace

int kp_knn_device(kp_ctx *ctx, const KpGrid &g, const float *d_queries, int64_t nq, int k, double radius, int32_t *d_idx,
                  double *d_d2, int32_t *d_count, double *d_mean, const float *d_xyz)
{
    KnnParams p;
    p.g = kp_grid_dev(g);
    p.queries = d_queries; p.nq = d_queries ? nq : g.n; p.k = k; p.mode = KQ_MODE_KNN;
    p.r2cap = radius > 0 ? radius * radius : 0.0;
    p.idx = d_idx; p.d2 = d_d2; p.count = d_count; p.mean = d_mean;
    p.cloud = nullptr; p.normals = nullptr; p.rcount = nullptr;
    return knn_launch(ctx, p, "knn", d_xyz);
}

int kp_sor_device(kp_ctx *ctx, const float *d_xyz, int64_t n, int k, double ratio, double cell_hint,
                  const float *h_bounds6, uint8_t *d_keep, double *d_mean, double *h_stats, int64_t *h_kept)
{
    if (k < 1 || !(ratio > 0.0)) return kp_set_err(ctx, KP_E_ARG, "remove_statistical_outlier: nb_neighbors < 1 or std_ratio <= 0");
    if (h_kept) *h_kept = 0;
    if (n <= 0) return KP_OK;
    // valid = number of points that get a neighbour list (NaN rows do not; compacted clouds have none)
    float b6[6];
    int64_t nvalid = n;
    if (!h_bounds6) {
        KP_TRY(kp_prim_bounds_fetch(ctx, d_xyz, n, b6, &nvalid));
        h_bounds6 = b6;
        if (nvalid == 0) {
            KP_CUDA(ctx, cudaMemsetAsync(d_keep, 0, (size_t)n, ctx->stream));
            return KP_OK;
        }
    }
    double cell = cell_hint;
    if (!(cell > 0.0)) KP_TRY(kp_grid_auto_cell(ctx, d_xyz, n, h_bounds6, 0.5 * k > 4 ? 0.5 * k : 4, &cell));
    KpGrid g;
    KP_TRY(kp_grid_build(ctx, d_xyz, n, cell, h_bounds6, &g));
    double *mean = d_mean, *tmp, *red;
    if (!mean) KP_TRY(kp_ws(ctx, (size_t)n, &mean));
    KP_TRY(kp_ws(ctx, (size_t)n, &tmp));
    KP_TRY(kp_ws(ctx, (size_t)n / 1024 + (size_t)n / 1048576 + 16, &red));
    {
        KnnParams p;
        p.g = kp_grid_dev(g);
        p.queries = nullptr; p.nq = n; p.k = k; p.mode = KQ_MODE_KNN; p.r2cap = 0.0;
        p.idx = nullptr; p.d2 = nullptr; p.count = nullptr; p.mean = mean;
        p.cloud = nullptr; p.normals = nullptr; p.rcount = nullptr;
        KP_TRY(knn_launch(ctx, p, "sor_knn", d_xyz));
    }
    KP_PROFB(ctx, "sor_stats", (double)n * (7.0 * 8.0 + 2.0));
    // NaN points never get a neighbour list: they count as "not computed" (mean = -1), like upstream's
    // empty-result branch; valid = number of points with a computed mean
    double *d_sum = (double *)ctx->d_scratch, *d_mu = d_sum + 1, *d_sq = d_sum + 2, *d_stats = d_sum + 4;
    int32_t *d_cnt = (int32_t *)(d_sum + 8);
    unsigned nb = kp_blocks(n, 256);
    k_sor_pos<<<nb, 256, 0, ctx->stream>>>(mean, n, tmp);
    KP_LAUNCH_CHECK(ctx);
    KP_TRY(kp_prim_csum(ctx, tmp, n, red, d_sum));
    k_sor_sq<<<nb, 256, 0, ctx->stream>>>(mean, n, d_sum, (double)nvalid, tmp, d_mu);
    KP_LAUNCH_CHECK(ctx);
    KP_TRY(kp_prim_csum(ctx, tmp, n, red, d_sq));
    k_sor_mask<<<nb, 256, 0, ctx->stream>>>(mean, n, d_mu, d_sq, (double)nvalid, ratio, d_keep, d_stats);
    KP_LAUNCH_CHECK(ctx);
    KP_TRY(kp_prim_count_u8(ctx, d_keep, n, d_cnt));
    KP_TRY(kp_fetch_scratch(ctx, 10 * sizeof(double)));
    const double *hs = (const double *)ctx->h_scratch;
    if (h_stats) { h_stats[0] = hs[4]; h_stats[1] = hs[5]; h_stats[2] = hs[6]; }
    if (h_kept) *h_kept = *(const int32_t *)(hs + 8);
    return KP_OK;
}

int kp_normals_device(kp_ctx *ctx, const float *d_xyz, int64_t n, double radius, int max_nn, const float *h_bounds6,
                      float *d_normals)
{
    if (max_nn < 1) return kp_set_err(ctx, KP_E_ARG, "estimate_normals: max_nn < 1");
    if (n <= 0) return KP_OK;
    double cell;
    if (radius > 0) cell = radius * (1.0 + 1e-6);
    else KP_TRY(kp_grid_auto_cell(ctx, d_xyz, n, h_bounds6, 0.5 * max_nn > 4 ? 0.5 * max_nn : 4, &cell));
    KpGrid g;
    KP_TRY(kp_grid_build(ctx, d_xyz, n, cell, h_bounds6, &g));
    KnnParams p;
    p.g = kp_grid_dev(g);
    p.queries = nullptr; p.nq = n; p.k = max_nn; p.mode = KQ_MODE_NORMALS;
    p.r2cap = radius > 0 ? radius * radius : 0.0;
    p.idx = nullptr; p.d2 = nullptr; p.count = nullptr; p.mean = nullptr;
    p.cloud = d_xyz; p.normals = d_normals; p.rcount = nullptr;
    return knn_launch(ctx, p, "normals");
}

extern "C" {

int kp_knn(kp_ctx *ctx, const float *d_xyz, int64_t n, const float *d_queries, int64_t nq, int k, double radius,
           double cell_hint, int32_t *d_idx, double *d_d2, int32_t *d_count)
{
    if (!ctx || (n > 0 && !d_xyz)) return kp_set_err(ctx, KP_E_ARG, "kp_knn: NULL argument");
    if (k < 1) return kp_set_err(ctx, KP_E_ARG, "kp_knn: k < 1");
    kp_enter(ctx);
    if (!d_queries) nq = n;
    if (nq <= 0) return KP_OK;
    double cell = cell_hint;
    if (!(cell > 0.0)) {
        if (radius > 0) cell = radius * (1.0 + 1e-6);
        else KP_TRY(kp_grid_auto_cell(ctx, d_xyz, n, nullptr, 0.5 * k > 4 ? 0.5 * k : 4, &cell));
    }
    KpGrid g;
    KP_TRY(kp_grid_build(ctx, d_xyz, n, cell, nullptr, &g));
    return kp_knn_device(ctx, g, d_queries, nq, k, radius, d_idx, d_d2, d_count, nullptr, d_xyz);
}

int kp_sor_mask(kp_ctx *ctx, const float *d_xyz, int64_t n, int nb_neighbors, double std_ratio, double cell_hint,
                uint8_t *d_keep, double *d_mean, double *h_stats, int64_t *h_kept)
{
    if (!ctx || (n > 0 && (!d_xyz || !d_keep))) return kp_set_err(ctx, KP_E_ARG, "kp_sor_mask: NULL argument");
    kp_enter(ctx);
    return kp_sor_device(ctx, d_xyz, n, nb_neighbors, std_ratio, cell_hint, nullptr, d_keep, d_mean, h_stats, h_kept);
}

int kp_radius_mask(kp_ctx *ctx, const float *d_xyz, int64_t n, int nb_points, double radius, uint8_t *d_keep,
                   int32_t *d_counts, int64_t *h_kept)
{
    if (!ctx || (n > 0 && (!d_xyz || !d_keep))) return kp_set_err(ctx, KP_E_ARG, "kp_radius_mask: NULL argument");
    if (nb_points < 1 || !(radius > 0.0)) return kp_set_err(ctx, KP_E_ARG, "remove_radius_outlier: nb_points < 1 or radius <= 0");
    kp_enter(ctx);
    if (h_kept) *h_kept = 0;
    if (n <= 0) return KP_OK;
    KpGrid g;
    KP_TRY(kp_grid_build(ctx, d_xyz, n, radius * (1.0 + 1e-6), nullptr, &g));
    int32_t *cnt = d_counts;
    if (!cnt) KP_TRY(kp_ws(ctx, (size_t)n, &cnt));
    KnnParams p;
    p.g = kp_grid_dev(g);
    p.queries = nullptr; p.nq = n; p.k = 1; p.mode = KQ_MODE_RADIUS; p.r2cap = radius * radius;
    p.idx = nullptr; p.d2 = nullptr; p.count = nullptr; p.mean = nullptr; p.cloud = nullptr; p.normals = nullptr;
    p.rcount = cnt;
    KP_TRY(knn_launch(ctx, p, "radius_count"));
    k_radius_mask<<<kp_blocks(n, 256), 256, 0, ctx->stream>>>(cnt, n, nb_points, d_keep);
    KP_LAUNCH_CHECK(ctx);
    int32_t *d_tot = (int32_t *)ctx->d_scratch;
    KP_TRY(kp_prim_count_u8(ctx, d_keep, n, d_tot));
    KP_TRY(kp_fetch_scratch(ctx, sizeof(int32_t)));
    if (h_kept) *h_kept = *(int32_t *)ctx->h_scratch;
    return KP_OK;
}

int kp_estimate_normals(kp_ctx *ctx, const float *d_xyz, int64_t n, double radius, int max_nn, float *d_normals)
{
    if (!ctx || (n > 0 && (!d_xyz || !d_normals))) return kp_set_err(ctx, KP_E_ARG, "kp_estimate_normals: NULL argument");
    kp_enter(ctx);
    return kp_normals_device(ctx, d_xyz, n, radius, max_nn, nullptr, d_normals);
}

}  // extern "C"
