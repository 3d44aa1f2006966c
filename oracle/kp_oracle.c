/*
 * kp_oracle.c -- CPU restatement of the KinectPy per-frame point-cloud path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under kinectpy_b200/ may import, link or
 * execute this file; it is the checker for tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs.
 *
 * PARITY UNPINNED: the reference (tiborcamargo/KinectPy) has no arithmetic of
 * its own on this path -- every step is a call into Open3D (version not pinned,
 * not present in this image) or, for depth->XYZ, into
 * the Azure Kinect SDK inside an external binary.  The reference ships no
 * tests, golden vectors or fixtures.  This file therefore restates the
 * *published* Open3D / k4a semantics at the reference's
 * own call sites, cited per function below, and tests/ cross-check it against
 * independent implementations (numpy brute force, scipy cKDTree, numpy.linalg).
 *
 * Arithmetic contract (shared with the CUDA kernels):
 *   - points are float32 in storage; every decision is taken in IEEE double on
 *     the exactly-converted values, in the operation order written here, with
 *     no fused multiply-add (build with -ffp-contract=off);
 *   - global / per-query sums that feed a decision use the canonical 32-lane
 *     tree `csum` below so that a parallel device reduction can be bit-equal;
 *   - k-nearest-neighbour results are ordered by (d^2, index) lexicographically.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define KPO_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ RNG -- */
/* Counter-based generator shared by oracle, CUDA kernels and the synthetic
 * scene generator (splitmix64 finaliser).  Replaces Open3D's global mt19937
 * in SegmentPlane so hypotheses are reproducible. */
static inline uint64_t kpo_mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
KPO_API uint64_t kpo_rng(uint64_t seed, uint64_t a, uint64_t b)
{
    return kpo_mix64(kpo_mix64(kpo_mix64(seed) + a) + b);
}

/* -------------------------------------------------------- canonical sum -- */
/* csum: groups of 1024 consecutive values; "lane" t adds x[t], x[t+32], ...
 * sequentially, then a 5-step xor butterfly (16,8,4,2,1); recurse on the
 * group sums until one value is left.  n == 0 -> 0.0. */
static double kpo_group1024(const double *x, long n)
{
    double v[32], w[32];
    for (int t = 0; t < 32; ++t) {
        double acc = 0.0;
        for (long i = t; i < n; i += 32) acc = acc + x[i];
        v[t] = acc;
    }
    for (int s = 16; s >= 1; s >>= 1) {
        for (int t = 0; t < 32; ++t) w[t] = v[t] + v[t ^ s];
        memcpy(v, w, sizeof v);
    }
    return v[0];
}
KPO_API double kpo_csum(const double *x, long n)
{
    if (n <= 0) return 0.0;
    long m = (n + 1023) / 1024;
    double *p = (double *)malloc(sizeof(double) * (size_t)m);
    const double *cur = x;
    long cn = n;
    double *owned = NULL;
    for (;;) {
        m = (cn + 1023) / 1024;
        for (long g = 0; g < m; ++g) {
            long lo = g * 1024, len = cn - lo < 1024 ? cn - lo : 1024;
            p[g] = kpo_group1024(cur + lo, len);
        }
        if (m == 1) break;
        double *nxt = (double *)malloc(sizeof(double) * (size_t)m);
        memcpy(nxt, p, sizeof(double) * (size_t)m);
        free(owned);
        owned = nxt;
        cur = nxt;
        cn = m;
    }
    double r = p[0];
    free(p);
    free(owned);
    return r;
}

/* --------------------------------------------------- K1 unproject (a1-a5) -- */
/* Depth -> XYZ through the xy-table, then the per-sensor extrinsic.
 * Follows: k4a transformation_depth_image_to_point_cloud semantics
 * (pinned in the reference only by the int16 (H*W,3) layout
 * read at utils/io.py:15-20), the validity rule of utils/io.py:36 (keep iff
 * x,y,z all != 0) and PointCloud.transform at preprocessing/data.py:46-48.
 * flags: bit0 = k4a int16 rounding mode, bit1 = drop_any_zero rule.
 * Invalid pixels produce NaN,NaN,NaN (and valid[i] = 0). */
#define KPO_F_INT16 1
#define KPO_F_DROP_ANY_ZERO 2
KPO_API void kpo_unproject(const uint16_t *depth, const float *xytab, const double *T,
                           int B, int S, long P, int flags, double scale, float *xyz,
                           uint8_t *valid, int16_t *xyz16)
{
    const float qnan = nanf("");
#pragma omp parallel for schedule(static)
    for (long bs = 0; bs < (long)B * S; ++bs) {
        int s = (int)(bs % S);
        const double *M = T ? T + 16 * s : NULL;
        for (long p = 0; p < P; ++p) {
            long i = bs * P + p;
            float xt = xytab[((long)s * P + p) * 2 + 0];
            float yt = xytab[((long)s * P + p) * 2 + 1];
            uint16_t z = depth[i];
            double X, Y, Z;
            int ok = !(isnan(xt) || isnan(yt)) && z != 0;
            int16_t xi = 0, yi = 0, zi = 0;
            if (ok) {
                if (flags & KPO_F_INT16) {
                    float zf = (float)z;
                    float fx = xt * zf;            /* fp32, separate mul and add */
                    float fy = yt * zf;
                    xi = (int16_t)(int32_t)floorf(fx + 0.5f);
                    yi = (int16_t)(int32_t)floorf(fy + 0.5f);
                    zi = (int16_t)z;
                    X = (double)xi * scale;
                    Y = (double)yi * scale;
                    Z = (double)zi * scale;
                } else {
                    X = ((double)xt * (double)z) * scale;
                    Y = ((double)yt * (double)z) * scale;
                    Z = (double)z * scale;
                }
                if ((flags & KPO_F_DROP_ANY_ZERO) && (X == 0.0 || Y == 0.0 || Z == 0.0)) ok = 0;
            }
            if (xyz16) {
                xyz16[3 * i + 0] = (flags & KPO_F_INT16) && !(isnan(xt) || isnan(yt)) && z ? xi : 0;
                xyz16[3 * i + 1] = (flags & KPO_F_INT16) && !(isnan(xt) || isnan(yt)) && z ? yi : 0;
                xyz16[3 * i + 2] = (flags & KPO_F_INT16) && !(isnan(xt) || isnan(yt)) && z ? zi : 0;
            }
            if (!ok) {
                xyz[3 * i] = xyz[3 * i + 1] = xyz[3 * i + 2] = qnan;
                if (valid) valid[i] = 0;
                continue;
            }
            if (M) {
                double x2 = ((M[0] * X + M[1] * Y) + M[2] * Z) + M[3];
                double y2 = ((M[4] * X + M[5] * Y) + M[6] * Z) + M[7];
                double z2 = ((M[8] * X + M[9] * Y) + M[10] * Z) + M[11];
                X = x2; Y = y2; Z = z2;
            }
            xyz[3 * i + 0] = (float)X;
            xyz[3 * i + 1] = (float)Y;
            xyz[3 * i + 2] = (float)Z;
            if (valid) valid[i] = 1;
        }
    }
}

/* PointCloud.transform (preprocessing/data.py:46-48), fp32
 * storage, double math, one rounding.  rotate_only = 1 for normals. */
KPO_API void kpo_transform(float *xyz, long n, const double *M, int rotate_only)
{
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) {
        double X = xyz[3 * i], Y = xyz[3 * i + 1], Z = xyz[3 * i + 2];
        double x2 = (M[0] * X + M[1] * Y) + M[2] * Z;
        double y2 = (M[4] * X + M[5] * Y) + M[6] * Z;
        double z2 = (M[8] * X + M[9] * Y) + M[10] * Z;
        if (!rotate_only) { x2 = x2 + M[3]; y2 = y2 + M[7]; z2 = z2 + M[11]; }
        xyz[3 * i] = (float)x2; xyz[3 * i + 1] = (float)y2; xyz[3 * i + 2] = (float)z2;
    }
}

/* ---------------------------------------------- K2 voxel downsample (a7) -- */
/* PointCloud.voxel_down_sample at preprocessing/filtering.py:23,
 * preprocessing/registration.py:8,100-101, utils/processing.py:308
 * (SURVEY.md A.2).  Canonical output order: voxels sorted by (ix,iy,iz);
 * inside a voxel points are accumulated in input order (what Open3D's
 * insertion loop does).  NaN points are skipped.  Returns M, or -1 on a bad
 * argument, -2 if the grid would overflow int32 indices. */
typedef struct { int32_t i, j, k; int64_t idx; } kpo_vox_t;
static int kpo_vox_cmp(const void *a, const void *b)
{
    const kpo_vox_t *p = (const kpo_vox_t *)a, *q = (const kpo_vox_t *)b;
    if (p->i != q->i) return p->i < q->i ? -1 : 1;
    if (p->j != q->j) return p->j < q->j ? -1 : 1;
    if (p->k != q->k) return p->k < q->k ? -1 : 1;
    return p->idx < q->idx ? -1 : (p->idx > q->idx);
}
KPO_API long kpo_voxel_downsample(const float *xyz, const float *colors, const float *normals,
                                  long n, double voxel, float *out_xyz, float *out_colors,
                                  float *out_normals, int32_t *out_ijk, int32_t *point_voxel,
                                  double *min_bound_out)
{
    if (!(voxel > 0.0)) return -1;
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    long nv = 0;
    for (long i = 0; i < n; ++i) {
        if (isnan(xyz[3 * i])) continue;
        for (int c = 0; c < 3; ++c) {
            float v = xyz[3 * i + c];
            if (v < mn[c]) mn[c] = v;
            if (v > mx[c]) mx[c] = v;
        }
        ++nv;
    }
    if (point_voxel) for (long i = 0; i < n; ++i) point_voxel[i] = -1;
    if (nv == 0) return 0;
    double minb[3], maxb[3];
    for (int c = 0; c < 3; ++c) {
        minb[c] = (double)mn[c] - voxel * 0.5;
        maxb[c] = (double)mx[c] + voxel * 0.5;
        if (voxel * 2147483647.0 < maxb[c] - minb[c]) return -2;
        if (min_bound_out) min_bound_out[c] = minb[c];
    }
    kpo_vox_t *v = (kpo_vox_t *)malloc(sizeof(kpo_vox_t) * (size_t)nv);
    long m = 0;
    for (long i = 0; i < n; ++i) {
        if (isnan(xyz[3 * i])) continue;
        v[m].i = (int32_t)floor(((double)xyz[3 * i + 0] - minb[0]) / voxel);
        v[m].j = (int32_t)floor(((double)xyz[3 * i + 1] - minb[1]) / voxel);
        v[m].k = (int32_t)floor(((double)xyz[3 * i + 2] - minb[2]) / voxel);
        v[m].idx = i;
        ++m;
    }
    qsort(v, (size_t)nv, sizeof(kpo_vox_t), kpo_vox_cmp);
    long M = 0;
    for (long a = 0; a < nv;) {
        long b = a;
        double sx = 0, sy = 0, sz = 0, cr = 0, cg = 0, cb = 0, nx = 0, ny = 0, nz = 0;
        while (b < nv && v[b].i == v[a].i && v[b].j == v[a].j && v[b].k == v[a].k) {
            long i = v[b].idx;
            sx = sx + (double)xyz[3 * i]; sy = sy + (double)xyz[3 * i + 1]; sz = sz + (double)xyz[3 * i + 2];
            if (colors) { cr = cr + (double)colors[3 * i]; cg = cg + (double)colors[3 * i + 1]; cb = cb + (double)colors[3 * i + 2]; }
            if (normals) { nx = nx + (double)normals[3 * i]; ny = ny + (double)normals[3 * i + 1]; nz = nz + (double)normals[3 * i + 2]; }
            if (point_voxel) point_voxel[i] = (int32_t)M;
            ++b;
        }
        double cnt = (double)(b - a);
        out_xyz[3 * M] = (float)(sx / cnt); out_xyz[3 * M + 1] = (float)(sy / cnt); out_xyz[3 * M + 2] = (float)(sz / cnt);
        if (colors && out_colors) {
            out_colors[3 * M] = (float)(cr / cnt); out_colors[3 * M + 1] = (float)(cg / cnt); out_colors[3 * M + 2] = (float)(cb / cnt);
        }
        if (normals && out_normals) {
            double ax = nx / cnt, ay = ny / cnt, az = nz / cnt;
            double nn = sqrt((ax * ax + ay * ay) + az * az);
            if (nn > 0.0) { ax = ax / nn; ay = ay / nn; az = az / nn; }
            out_normals[3 * M] = (float)ax; out_normals[3 * M + 1] = (float)ay; out_normals[3 * M + 2] = (float)az;
        }
        if (out_ijk) { out_ijk[3 * M] = v[a].i; out_ijk[3 * M + 1] = v[a].j; out_ijk[3 * M + 2] = v[a].k; }
        ++M;
        a = b;
    }
    free(v);
    return M;
}

/* ----------------------------------------------- neighbour search (K3) -- */
/* Exact k-nearest / hybrid search with canonical (d^2, index) order.
 * Stands in for the nanoflann KD-tree under remove_statistical_outlier
 * (preprocessing/filtering.py:24, floor_removal.py:73), estimate_normals
 * (preprocessing/registration.py:11-13) and registration_icp
 * (preprocessing/registration.py:78-84).  The result of an exact search does
 * not depend on the index structure, so a uniform grid with ring expansion is
 * used here; tests pin it against brute force and scipy.cKDTree.
 * radius > 0 adds the hybrid cap d^2 < radius^2 (strict). */
typedef struct {
    long n;
    const float *pts;
    double org[3], cell;
    int dim[3];
    int64_t *start;  /* dim0*dim1*dim2 + 1 */
    int32_t *order;  /* point ids sorted by cell, ascending id inside a cell */
} kpo_grid_t;

static inline double kpo_d2(const float *a, const float *b)
{
    double dx = (double)a[0] - (double)b[0];
    double dy = (double)a[1] - (double)b[1];
    double dz = (double)a[2] - (double)b[2];
    return (dx * dx + dy * dy) + dz * dz;
}
static inline int kpo_cell_of(const kpo_grid_t *g, double v, int c)
{
    double f = floor((v - g->org[c]) / g->cell);
    if (f < 0) return -1;
    if (f >= g->dim[c]) return g->dim[c];
    return (int)f;
}
static int kpo_grid_build(kpo_grid_t *g, const float *pts, long n, double cell_hint)
{
    g->n = n; g->pts = pts; g->start = NULL; g->order = NULL;
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    long nv = 0;
    for (long i = 0; i < n; ++i) {
        if (isnan(pts[3 * i])) continue;
        ++nv;
        for (int c = 0; c < 3; ++c) {
            if (pts[3 * i + c] < mn[c]) mn[c] = pts[3 * i + c];
            if (pts[3 * i + c] > mx[c]) mx[c] = pts[3 * i + c];
        }
    }
    if (nv == 0) { g->dim[0] = g->dim[1] = g->dim[2] = 0; return 0; }
    double ext[3] = {(double)mx[0] - mn[0], (double)mx[1] - mn[1], (double)mx[2] - mn[2]};
    double cell = cell_hint;
    if (!(cell > 0)) {
        double e = fmax(ext[0], fmax(ext[1], ext[2]));
        cell = e > 0 ? e / 64.0 : 1.0;
    }
    /* bound the table to ~16M cells */
    for (;;) {
        double tot = 1;
        for (int c = 0; c < 3; ++c) tot *= floor(ext[c] / cell) + 1.0;
        if (tot <= 16.0e6) break;
        cell *= 1.26;
    }
    g->cell = cell;
    for (int c = 0; c < 3; ++c) { g->org[c] = mn[c]; g->dim[c] = (int)floor(ext[c] / cell) + 1; }
    long nc = (long)g->dim[0] * g->dim[1] * g->dim[2];
    g->start = (int64_t *)calloc((size_t)nc + 1, sizeof(int64_t));
    g->order = (int32_t *)malloc(sizeof(int32_t) * (size_t)(nv ? nv : 1));
    int64_t *cid = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
    for (long i = 0; i < n; ++i) {
        if (isnan(pts[3 * i])) { cid[i] = -1; continue; }
        int cx = kpo_cell_of(g, pts[3 * i], 0), cy = kpo_cell_of(g, pts[3 * i + 1], 1), cz = kpo_cell_of(g, pts[3 * i + 2], 2);
        if (cx >= g->dim[0]) cx = g->dim[0] - 1;
        if (cy >= g->dim[1]) cy = g->dim[1] - 1;
        if (cz >= g->dim[2]) cz = g->dim[2] - 1;
        cid[i] = ((int64_t)cx * g->dim[1] + cy) * g->dim[2] + cz;
        g->start[cid[i] + 1]++;
    }
    for (long c = 0; c < nc; ++c) g->start[c + 1] += g->start[c];
    int64_t *fill = (int64_t *)malloc(sizeof(int64_t) * (size_t)nc);
    memcpy(fill, g->start, sizeof(int64_t) * (size_t)nc);
    for (long i = 0; i < n; ++i) if (cid[i] >= 0) g->order[fill[cid[i]]++] = (int32_t)i;
    free(fill); free(cid);
    return 0;
}
static void kpo_grid_free(kpo_grid_t *g) { free(g->start); free(g->order); }

/* insert (d2, id) into the sorted list kept in (bd, bi) of length *cnt <= k */
static inline void kpo_topk_insert(double *bd, int32_t *bi, int *cnt, int k, double d2, int32_t id)
{
    int c = *cnt;
    if (c == k) {
        if (d2 > bd[k - 1] || (d2 == bd[k - 1] && id > bi[k - 1])) return;
        c = k - 1;
    }
    int p = c;
    while (p > 0 && (bd[p - 1] > d2 || (bd[p - 1] == d2 && bi[p - 1] > id))) {
        bd[p] = bd[p - 1]; bi[p] = bi[p - 1]; --p;
    }
    bd[p] = d2; bi[p] = id;
    *cnt = c + 1;
}
static void kpo_query(const kpo_grid_t *g, const float *q, int k, double r2cap, double *bd, int32_t *bi, int *out_cnt)
{
    int cnt = 0;
    *out_cnt = 0;
    if (g->dim[0] == 0) return;
    /* rings are centred on the in-grid cell nearest to the query: a grid point within rho of an outside
     * query is still within ceil(rho/cell) cells of that clamped cell on every axis */
    int c0[3];
    int maxring = 0;
    for (int c = 0; c < 3; ++c) {
        c0[c] = kpo_cell_of(g, q[c], c);
        if (c0[c] < 0) c0[c] = 0;
        if (c0[c] >= g->dim[c]) c0[c] = g->dim[c] - 1;
        if (c0[c] > maxring) maxring = c0[c];
        if (g->dim[c] - 1 - c0[c] > maxring) maxring = g->dim[c] - 1 - c0[c];
    }
    for (int ring = 0; ring <= maxring; ++ring) {
        for (int dx = -ring; dx <= ring; ++dx) {
            int cx = c0[0] + dx; if (cx < 0 || cx >= g->dim[0]) continue;
            for (int dy = -ring; dy <= ring; ++dy) {
                int cy = c0[1] + dy; if (cy < 0 || cy >= g->dim[1]) continue;
                int onshell_xy = (abs(dx) == ring || abs(dy) == ring);
                for (int dz = -ring; dz <= ring; dz += (onshell_xy || ring == 0) ? 1 : 2 * ring) {
                    int cz = c0[2] + dz; if (cz < 0 || cz >= g->dim[2]) continue;
                    long cid = ((long)cx * g->dim[1] + cy) * g->dim[2] + cz;
                    for (int64_t t = g->start[cid]; t < g->start[cid + 1]; ++t) {
                        int32_t id = g->order[t];
                        double d2 = kpo_d2(q, g->pts + 3 * (long)id);
                        if (r2cap > 0 && !(d2 < r2cap)) continue;
                        kpo_topk_insert(bd, bi, &cnt, k, d2, id);
                    }
                }
            }
        }
        /* every point closer than ring*cell (conservatively shrunk) has been seen */
        double safe = (double)ring * g->cell * (1.0 - 1.0 / 1048576.0);
        double s2 = safe * safe;
        if (cnt == k && bd[k - 1] <= s2) break;
        if (r2cap > 0 && r2cap <= s2) break;
    }
    *out_cnt = cnt;
}
KPO_API void kpo_knn(const float *pts, long n, const float *queries, long nq, int k, double radius,
                     int32_t *idx, double *d2, int32_t *cnt)
{
    kpo_grid_t g;
    if (!queries) { queries = pts; nq = n; }
    kpo_grid_build(&g, pts, n, radius > 0 ? radius : 0.0);
    double r2 = radius > 0 ? radius * radius : 0.0;
#pragma omp parallel
    {
        double *bd = (double *)malloc(sizeof(double) * (size_t)k);
        int32_t *bi = (int32_t *)malloc(sizeof(int32_t) * (size_t)k);
#pragma omp for schedule(dynamic, 256)
        for (long i = 0; i < nq; ++i) {
            int c = 0;
            if (!isnan(queries[3 * i])) kpo_query(&g, queries + 3 * i, k, r2, bd, bi, &c);
            for (int j = 0; j < k; ++j) {
                if (idx) idx[i * k + j] = j < c ? bi[j] : -1;
                if (d2) d2[i * k + j] = j < c ? bd[j] : INFINITY;
            }
            if (cnt) cnt[i] = c;
        }
        free(bd); free(bi);
    }
    kpo_grid_free(&g);
}

/* ------------------------------------- statistical outlier removal (a8) -- */
/* PointCloud.remove_statistical_outlier at preprocessing/filtering.py:24,
 * floor_removal.py:73, utils/processing.py:309.
 * stats[0]=mu, stats[1]=std, stats[2]=threshold.  Returns number kept. */
KPO_API long kpo_sor(const float *pts, long n, int k, double std_ratio, double cell_hint,
                     uint8_t *keep, double *mean_out, double *stats)
{
    if (k < 1 || !(std_ratio > 0.0)) return -1;
    if (n == 0) return 0;
    double *mean = (double *)malloc(sizeof(double) * (size_t)n);
    kpo_grid_t g;
    kpo_grid_build(&g, pts, n, cell_hint);
#pragma omp parallel
    {
        double *bd = (double *)malloc(sizeof(double) * (size_t)k);
        int32_t *bi = (int32_t *)malloc(sizeof(int32_t) * (size_t)k);
#pragma omp for schedule(dynamic, 256)
        for (long i = 0; i < n; ++i) {
            int c = 0;
            if (!isnan(pts[3 * i])) kpo_query(&g, pts + 3 * i, k, 0.0, bd, bi, &c);
            if (c == 0) { mean[i] = -1.0; continue; }
            /* std::accumulate over the neighbour distances in ascending (d2, index) order */
            double acc = 0.0;
            for (int j = 0; j < c; ++j) acc = acc + sqrt(bd[j]);
            mean[i] = acc / (double)c;
        }
        free(bd); free(bi);
    }
    kpo_grid_free(&g);
    long valid = 0;
    double *tmp = (double *)malloc(sizeof(double) * (size_t)n);
    for (long i = 0; i < n; ++i) { if (mean[i] != -1.0) ++valid; tmp[i] = mean[i] > 0 ? mean[i] : 0.0; }
    double mu = kpo_csum(tmp, n) / (double)valid;
    for (long i = 0; i < n; ++i) tmp[i] = mean[i] > 0 ? (mean[i] - mu) * (mean[i] - mu) : 0.0;
    double sd = sqrt(kpo_csum(tmp, n) / (double)(valid - 1));
    double thr = mu + std_ratio * sd;
    long kept = 0;
    for (long i = 0; i < n; ++i) {
        int kp = mean[i] > 0 && mean[i] < thr;
        keep[i] = (uint8_t)kp; kept += kp;
    }
    if (mean_out) memcpy(mean_out, mean, sizeof(double) * (size_t)n);
    if (stats) { stats[0] = mu; stats[1] = sd; stats[2] = thr; }
    free(tmp); free(mean);
    return kept;
}

/* ------------------------------------------- radius outlier removal (a9) -- */
/* PointCloud.remove_radius_outlier (no call site in the reference; north-star
 * scope): count strict d^2 < r^2 including self; keep iff
 * count > nb_points. */
KPO_API long kpo_radius_outlier(const float *pts, long n, int nb_points, double radius,
                                uint8_t *keep, int32_t *counts)
{
    if (nb_points < 1 || !(radius > 0.0)) return -1;
    kpo_grid_t g;
    kpo_grid_build(&g, pts, n, radius);
    double r2 = radius * radius;
    long kept = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : kept)
    for (long i = 0; i < n; ++i) {
        const float *q = pts + 3 * i;
        int cnt = 0;
        if (!isnan(q[0])) {
            int c0[3];
            for (int c = 0; c < 3; ++c) c0[c] = kpo_cell_of(&g, q[c], c);
            int R = (int)ceil(radius / g.cell) + 1;
            for (int dx = -R; dx <= R; ++dx) for (int dy = -R; dy <= R; ++dy) for (int dz = -R; dz <= R; ++dz) {
                int cx = c0[0] + dx, cy = c0[1] + dy, cz = c0[2] + dz;
                if (cx < 0 || cy < 0 || cz < 0 || cx >= g.dim[0] || cy >= g.dim[1] || cz >= g.dim[2]) continue;
                long cid = ((long)cx * g.dim[1] + cy) * g.dim[2] + cz;
                for (int64_t t = g.start[cid]; t < g.start[cid + 1]; ++t)
                    if (kpo_d2(q, pts + 3 * (long)g.order[t]) < r2) ++cnt;
            }
        }
        if (counts) counts[i] = cnt;
        keep[i] = (uint8_t)(cnt > nb_points);
        kept += cnt > nb_points;
    }
    kpo_grid_free(&g);
    return kept;
}

/* ----------------------------------------------------- normals (a14) -- */
/* estimate_normals(KDTreeSearchParamHybrid(radius, max_nn)) at
 * preprocessing/registration.py:11-13: covariance from
 * cumulants over the hybrid neighbourhood, eigenvector of the smallest
 * eigenvalue by the analytic symmetric 3x3 solver, (0,0,1) fallback. */
static void kpo_cross(const double *a, const double *b, double *o)
{
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
static void kpo_evec0(const double A[6], double ev, double *out)
{   /* A = {xx,xy,xz,yy,yz,zz}; eigenvector for eigenvalue ev via largest row cross product */
    double r0[3] = {A[0] - ev, A[1], A[2]}, r1[3] = {A[1], A[3] - ev, A[4]}, r2[3] = {A[2], A[4], A[5] - ev};
    double c01[3], c02[3], c12[3];
    kpo_cross(r0, r1, c01); kpo_cross(r0, r2, c02); kpo_cross(r1, r2, c12);
    double d0 = c01[0] * c01[0] + c01[1] * c01[1] + c01[2] * c01[2];
    double d1 = c02[0] * c02[0] + c02[1] * c02[1] + c02[2] * c02[2];
    double d2 = c12[0] * c12[0] + c12[1] * c12[1] + c12[2] * c12[2];
    const double *best = c01; double dm = d0;
    if (d1 > dm) { dm = d1; best = c02; }
    if (d2 > dm) { dm = d2; best = c12; }
    if (dm > 0) { double s = sqrt(dm); out[0] = best[0] / s; out[1] = best[1] / s; out[2] = best[2] / s; }
    else { out[0] = out[1] = out[2] = 0.0; }
}
static void kpo_evec1(const double A[6], const double *e0, double ev1, double *out)
{   /* second eigenvector in the plane orthogonal to e0 (Eberly's robust 3x3 scheme) */
    double U[3], V[3];
    if (fabs(e0[0]) > fabs(e0[1])) { double il = 1.0 / sqrt(e0[0] * e0[0] + e0[2] * e0[2]); U[0] = -e0[2] * il; U[1] = 0; U[2] = e0[0] * il; }
    else { double il = 1.0 / sqrt(e0[1] * e0[1] + e0[2] * e0[2]); U[0] = 0; U[1] = e0[2] * il; U[2] = -e0[1] * il; }
    kpo_cross(e0, U, V);
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
KPO_API void kpo_smallest_eigvec(const double cov_in[6], double *nrm)
{   /* cov_in = {xx,xy,xz,yy,yz,zz} */
    double A[6];
    double mc = cov_in[0];
    for (int i = 1; i < 6; ++i) if (cov_in[i] > mc) mc = cov_in[i];
    if (mc == 0.0) { nrm[0] = nrm[1] = nrm[2] = 0.0; return; }
    for (int i = 0; i < 6; ++i) A[i] = cov_in[i] / mc;
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
        if (half < -1.0) half = -1.0;
        if (half > 1.0) half = 1.0;
        double angle = acos(half) / 3.0;
        const double two_thirds_pi = 2.09439510239319549;
        double beta2 = cos(angle) * 2.0;
        double beta0 = cos(angle + two_thirds_pi) * 2.0;
        double beta1 = -(beta0 + beta2);
        double e0 = q + p * beta0, e1 = q + p * beta1, e2 = q + p * beta2;
        double v0[3], v1[3], v2[3];
        if (half >= 0) {
            kpo_evec0(A, e2, v2);
            kpo_evec1(A, v2, e1, v1);
            kpo_cross(v1, v2, v0);
        } else {
            kpo_evec0(A, e0, v0);
        }
        (void)e1;
        nrm[0] = v0[0]; nrm[1] = v0[1]; nrm[2] = v0[2];
    } else {
        if (A[0] < A[3] && A[0] < A[5]) { nrm[0] = 1; nrm[1] = 0; nrm[2] = 0; }
        else if (A[3] < A[0] && A[3] < A[5]) { nrm[0] = 0; nrm[1] = 1; nrm[2] = 0; }
        else { nrm[0] = 0; nrm[1] = 0; nrm[2] = 1; }
    }
}
KPO_API void kpo_estimate_normals(const float *pts, long n, double radius, int max_nn, float *normals)
{
    kpo_grid_t g;
    kpo_grid_build(&g, pts, n, radius > 0 ? radius : 0.0);
    double r2 = radius > 0 ? radius * radius : 0.0;
#pragma omp parallel
    {
        double *bd = (double *)malloc(sizeof(double) * (size_t)max_nn);
        int32_t *bi = (int32_t *)malloc(sizeof(int32_t) * (size_t)max_nn);
#pragma omp for schedule(dynamic, 256)
        for (long i = 0; i < n; ++i) {
            int c = 0;
            double nr[3] = {0, 0, 1};
            if (!isnan(pts[3 * i])) kpo_query(&g, pts + 3 * i, max_nn, r2, bd, bi, &c);
            if (c >= 3) {
                double s[9] = {0};
                for (int j = 0; j < c; ++j) {
                    double x = pts[3 * (long)bi[j]], y = pts[3 * (long)bi[j] + 1], z = pts[3 * (long)bi[j] + 2];
                    s[0] += x; s[1] += y; s[2] += z;
                    s[3] += x * x; s[4] += x * y; s[5] += x * z; s[6] += y * y; s[7] += y * z; s[8] += z * z;
                }
                for (int j = 0; j < 9; ++j) s[j] /= (double)c;
                double cov[6] = {s[3] - s[0] * s[0], s[4] - s[0] * s[1], s[5] - s[0] * s[2],
                                 s[6] - s[1] * s[1], s[7] - s[1] * s[2], s[8] - s[2] * s[2]};
                kpo_smallest_eigvec(cov, nr);
                if (nr[0] == 0.0 && nr[1] == 0.0 && nr[2] == 0.0) { nr[2] = 1.0; }
            }
            normals[3 * i] = (float)nr[0]; normals[3 * i + 1] = (float)nr[1]; normals[3 * i + 2] = (float)nr[2];
        }
        free(bd); free(bi);
    }
    kpo_grid_free(&g);
}

/* -------------------------------------------- RANSAC plane (a11, a13) -- */
/* PointCloud.segment_plane at floor_removal.py:70.
 * Hypothesis h samples ransac_n distinct indices kpo_rng(seed,h,j) % n,
 * j = 0,1,... with rejection of repeats.  Sequential best / early-exit rule.
 * plane[4] is the refit plane, inlier_mask belongs to the pre-refit best
 * plane (as upstream).  Returns #inliers (0 if no valid hypothesis), <0 on
 * bad arguments. */
static int kpo_fit_plane(const float *pts, const int64_t *ids, long m, double *pl, int force_cov)
{
    if (m == 3 && !force_cov) {
        const float *p0 = pts + 3 * ids[0], *p1 = pts + 3 * ids[1], *p2 = pts + 3 * ids[2];
        double e1[3] = {(double)p1[0] - p0[0], (double)p1[1] - p0[1], (double)p1[2] - p0[2]};
        double e2[3] = {(double)p2[0] - p0[0], (double)p2[1] - p0[1], (double)p2[2] - p0[2]};
        double a = e1[1] * e2[2] - e1[2] * e2[1];
        double b = e1[2] * e2[0] - e1[0] * e2[2];
        double c = e1[0] * e2[1] - e1[1] * e2[0];
        double nn = sqrt((a * a + b * b) + c * c);
        if (nn == 0.0) return 0;
        a = a / nn; b = b / nn; c = c / nn;
        pl[0] = a; pl[1] = b; pl[2] = c;
        pl[3] = -((a * (double)p0[0] + b * (double)p0[1]) + c * (double)p0[2]);
        return 1;
    }
    double cx = 0, cy = 0, cz = 0;
    for (long j = 0; j < m; ++j) { cx = cx + pts[3 * ids[j]]; cy = cy + pts[3 * ids[j] + 1]; cz = cz + pts[3 * ids[j] + 2]; }
    cx = cx / (double)m; cy = cy / (double)m; cz = cz / (double)m;
    double xx = 0, xy = 0, xz = 0, yy = 0, yz = 0, zz = 0;
    for (long j = 0; j < m; ++j) {
        double x = (double)pts[3 * ids[j]] - cx, y = (double)pts[3 * ids[j] + 1] - cy, z = (double)pts[3 * ids[j] + 2] - cz;
        xx = xx + x * x; xy = xy + x * y; xz = xz + x * z; yy = yy + y * y; yz = yz + y * z; zz = zz + z * z;
    }
    double dx = yy * zz - yz * yz, dy = xx * zz - xz * xz, dz = xx * yy - xy * xy;
    double a, b, c;
    if (dx >= dy && dx >= dz) { a = dx; b = xz * yz - xy * zz; c = xy * yz - xz * yy; }
    else if (dy >= dx && dy >= dz) { a = xz * yz - xy * zz; b = dy; c = xy * xz - yz * xx; }
    else { a = xy * yz - xz * yy; b = xy * xz - yz * xx; c = dz; }
    double nn = sqrt((a * a + b * b) + c * c);
    if (nn == 0.0) return 0;
    a = a / nn; b = b / nn; c = c / nn;
    pl[0] = a; pl[1] = b; pl[2] = c;
    pl[3] = -((a * cx + b * cy) + c * cz);
    return 1;
}
KPO_API void kpo_ransac_sample(uint64_t seed, long h, long n, int ransac_n, int64_t *ids)
{
    int got = 0;
    for (uint64_t j = 0; got < ransac_n; ++j) {
        int64_t v = (int64_t)(kpo_rng(seed, (uint64_t)h, j) % (uint64_t)n);
        int dup = 0;
        for (int t = 0; t < got; ++t) if (ids[t] == v) { dup = 1; break; }
        if (!dup) ids[got++] = v;
    }
}
KPO_API long kpo_ransac_plane(const float *pts, long n, double thr, int ransac_n, int iters,
                              double probability, uint64_t seed, double *plane, uint8_t *inlier_mask,
                              int32_t *best_iter_out, int64_t *counts_out)
{
    if (ransac_n < 3 || n < ransac_n || !(thr > 0) || iters < 1) return -1;
    double *planes = (double *)malloc(sizeof(double) * 4 * (size_t)iters);
    uint8_t *pvalid = (uint8_t *)malloc((size_t)iters);
    int64_t *cnt = (int64_t *)calloc((size_t)iters, sizeof(int64_t));
    double *sq = (double *)calloc((size_t)iters, sizeof(double));
#pragma omp parallel
    {
        int64_t *ids = (int64_t *)malloc(sizeof(int64_t) * (size_t)ransac_n);
#pragma omp for schedule(dynamic, 4)
        for (int h = 0; h < iters; ++h) {
            kpo_ransac_sample(seed, h, n, ransac_n, ids);
            pvalid[h] = (uint8_t)kpo_fit_plane(pts, ids, ransac_n, planes + 4 * h, 0);
            if (!pvalid[h]) continue;
            const double *pl = planes + 4 * h;
            int64_t c = 0; double s = 0;
            for (long i = 0; i < n; ++i) {
                double d = fabs(((pl[0] * (double)pts[3 * i] + pl[1] * (double)pts[3 * i + 1]) + pl[2] * (double)pts[3 * i + 2]) + pl[3]);
                if (d < thr) { ++c; s = s + d * d; }
            }
            cnt[h] = c; sq[h] = s;
        }
        free(ids);
    }
    /* sequential replay of Open3D's best / early-exit rule */
    int best = -1; double best_fit = 0, best_rmse = 0;
    double break_it = (double)iters; long done = 0;
    for (int h = 0; h < iters; ++h) {
        if ((double)done > break_it) continue;
        if (!pvalid[h]) continue;                 /* degenerate sample: skipped, not counted (upstream `continue`) */
        {
            double fit = (double)cnt[h] / (double)n;
            double rm = cnt[h] ? sq[h] / sqrt((double)cnt[h]) : 0.0;
            if (fit > best_fit || (fit == best_fit && rm < best_rmse && best >= 0)) {
                best = h; best_fit = fit; best_rmse = rm;
                if (fit < 1.0) {
                    double b = log(1.0 - probability) / log(1.0 - pow(fit, (double)ransac_n));
                    break_it = b < (double)iters ? b : (double)iters;
                } else break_it = 0;
            }
        }
        ++done;
    }
    if (counts_out) memcpy(counts_out, cnt, sizeof(int64_t) * (size_t)iters);
    if (best_iter_out) *best_iter_out = best;
    long ninl = 0;
    memset(inlier_mask, 0, (size_t)n);
    if (best >= 0) {
        const double *pl = planes + 4 * best;
        for (long i = 0; i < n; ++i) {
            double d = fabs(((pl[0] * (double)pts[3 * i] + pl[1] * (double)pts[3 * i + 1]) + pl[2] * (double)pts[3 * i + 2]) + pl[3]);
            if (d < thr) { inlier_mask[i] = 1; ++ninl; }
        }
        int64_t *ids = (int64_t *)malloc(sizeof(int64_t) * (size_t)(ninl ? ninl : 1));
        long t = 0;
        for (long i = 0; i < n; ++i) if (inlier_mask[i]) ids[t++] = i;
        memcpy(plane, pl, sizeof(double) * 4);
        /* refit on the final inliers with the covariance fit (always, as upstream);
         * a degenerate refit keeps the best hypothesis plane */
        double rp[4];
        if (kpo_fit_plane(pts, ids, ninl, rp, 1)) memcpy(plane, rp, sizeof rp);
        free(ids);
    } else {
        plane[0] = plane[1] = plane[2] = plane[3] = 0.0;
    }
    free(planes); free(pvalid); free(cnt); free(sq);
    return ninl;
}
/* pcd_above_plane (floor_removal.py:39-51): keep iff a*x+b*y+c*z+d < 0,
 * evaluated left to right in double as the Python loop at :43 does. */
KPO_API long kpo_plane_side(const float *pts, long n, double a, double b, double c, double d, uint8_t *keep)
{
    long kept = 0;
    for (long i = 0; i < n; ++i) {
        double v = ((a * (double)pts[3 * i] + b * (double)pts[3 * i + 1]) + c * (double)pts[3 * i + 2]) + d;
        keep[i] = (uint8_t)(v < 0.0); kept += v < 0.0;   /* NaN -> dropped */
    }
    return kept;
}

/* ------------------------------------------ point-to-plane ICP (a15) -- */
/* registration_icp(..., TransformationEstimationPointToPlane()) at
 * preprocessing/registration.py:78-84.  src is moved in
 * double, in place, update by update, like upstream's pcd.Transform(update). */
static int kpo_solve6(double A[6][6], double b[6], double x[6])
{   /* Gaussian elimination with partial pivoting on the SPD-ish normal matrix */
    double M[6][7];
    for (int i = 0; i < 6; ++i) { for (int j = 0; j < 6; ++j) M[i][j] = A[i][j]; M[i][6] = b[i]; }
    for (int c = 0; c < 6; ++c) {
        int p = c; double best = fabs(M[c][c]);
        for (int r = c + 1; r < 6; ++r) if (fabs(M[r][c]) > best) { best = fabs(M[r][c]); p = r; }
        if (!(best > 1e-300)) return 0;
        if (p != c) for (int j = 0; j < 7; ++j) { double t = M[c][j]; M[c][j] = M[p][j]; M[p][j] = t; }
        for (int r = c + 1; r < 6; ++r) {
            double f = M[r][c] / M[c][c];
            for (int j = c; j < 7; ++j) M[r][j] -= f * M[c][j];
        }
    }
    for (int r = 5; r >= 0; --r) {
        double s = M[r][6];
        for (int j = r + 1; j < 6; ++j) s -= M[r][j] * x[j];
        x[r] = s / M[r][r];
        if (!isfinite(x[r])) return 0;
    }
    return 1;
}
static void kpo_mat4_mul(const double *A, const double *B, double *C)
{
    double R[16];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) {
        double s = 0;
        for (int k = 0; k < 4; ++k) s = s + A[4 * i + k] * B[4 * k + j];
        R[4 * i + j] = s;
    }
    memcpy(C, R, sizeof R);
}
static void kpo_x6_to_mat4(const double *x, double *M)
{   /* R = Rz(g) * Ry(b) * Rx(a), t = x[3..5] */
    double ca = cos(x[0]), sa = sin(x[0]), cb = cos(x[1]), sb = sin(x[1]), cg = cos(x[2]), sg = sin(x[2]);
    M[0] = cg * cb; M[1] = cg * sb * sa - sg * ca; M[2] = cg * sb * ca + sg * sa; M[3] = x[3];
    M[4] = sg * cb; M[5] = sg * sb * sa + cg * ca; M[6] = sg * sb * ca - cg * sa; M[7] = x[4];
    M[8] = -sb;     M[9] = cb * sa;                M[10] = cb * ca;               M[11] = x[5];
    M[12] = 0; M[13] = 0; M[14] = 0; M[15] = 1;
}
/* ---- Eigen::umeyama without scaling (TransformationEstimationPointToPoint::ComputeTransformation,
 * manual_pointcloud_registration.py:90-98): R = U diag(1,1,det(U)det(V)) V^T of the cross-covariance
 * (1/n) sum (t - mt)(s - ms)^T, t = mt - R ms.  SVD by one-sided (Hestenes) Jacobi rotations. */
static void kpo_svd3(const double C[3][3], double U[3][3], double S[3], double V[3][3])
{
    double A[3][3];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { A[i][j] = C[i][j]; V[i][j] = i == j; }
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0;
        for (int p = 0; p < 2; ++p) for (int q = p + 1; q < 3; ++q) {
            double alpha = 0, beta = 0, gamma = 0;
            for (int k = 0; k < 3; ++k) { alpha += A[k][p] * A[k][p]; beta += A[k][q] * A[k][q]; gamma += A[k][p] * A[k][q]; }
            if (gamma == 0.0) continue;
            off = fmax(off, fabs(gamma) / sqrt(alpha * beta + 1e-300));
            double zeta = (beta - alpha) / (2.0 * gamma);
            double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
            double c = 1.0 / sqrt(1.0 + t * t), sn = c * t;
            for (int k = 0; k < 3; ++k) {
                double ap = A[k][p], aq = A[k][q];
                A[k][p] = c * ap - sn * aq; A[k][q] = sn * ap + c * aq;
                double vp = V[k][p], vq = V[k][q];
                V[k][p] = c * vp - sn * vq; V[k][q] = sn * vp + c * vq;
            }
        }
        if (off < 1e-15) break;
    }
    int o[3] = {0, 1, 2};
    double nrm[3];
    for (int j = 0; j < 3; ++j) nrm[j] = sqrt(A[0][j] * A[0][j] + A[1][j] * A[1][j] + A[2][j] * A[2][j]);
    for (int a2 = 0; a2 < 2; ++a2) for (int b2 = a2 + 1; b2 < 3; ++b2) if (nrm[o[b2]] > nrm[o[a2]]) { int t = o[a2]; o[a2] = o[b2]; o[b2] = t; }
    double Vs[3][3];
    for (int j = 0; j < 3; ++j) {
        S[j] = nrm[o[j]];
        for (int k = 0; k < 3; ++k) { Vs[k][j] = V[k][o[j]]; U[k][j] = S[j] > 1e-300 ? A[k][o[j]] / S[j] : 0.0; }
    }
    memcpy(V, Vs, sizeof Vs);
    /* complete U to an orthonormal basis when singular values vanish */
    if (!(S[1] > 1e-12 * S[0])) {
        int ax = fabs(U[0][0]) <= fabs(U[1][0]) && fabs(U[0][0]) <= fabs(U[2][0]) ? 0 : (fabs(U[1][0]) <= fabs(U[2][0]) ? 1 : 2);
        double e[3] = {0, 0, 0}; e[ax] = 1;
        double d = U[ax][0], l = 0;
        for (int k = 0; k < 3; ++k) { U[k][1] = e[k] - d * U[k][0]; l += U[k][1] * U[k][1]; }
        l = sqrt(l);
        for (int k = 0; k < 3; ++k) U[k][1] /= l;
    }
    if (!(S[2] > 1e-12 * S[0])) {
        /* third left vector: any unit vector completing the basis; its sign is absorbed by diag(1,1,det U det V) */
        U[0][2] = U[1][0] * U[2][1] - U[2][0] * U[1][1];
        U[1][2] = U[2][0] * U[0][1] - U[0][0] * U[2][1];
        U[2][2] = U[0][0] * U[1][1] - U[1][0] * U[0][1];
    }
}
static double kpo_det3(const double M[3][3])
{
    return M[0][0] * (M[1][1] * M[2][2] - M[1][2] * M[2][1]) - M[0][1] * (M[1][0] * M[2][2] - M[1][2] * M[2][0]) +
           M[0][2] * (M[1][0] * M[2][1] - M[1][1] * M[2][0]);
}
/* src, tgt: double [n][3] matched rows; out: 4x4 row-major */
KPO_API void kpo_umeyama(const double *src, const double *tgt, long n, double *T)
{
    memset(T, 0, sizeof(double) * 16);
    T[0] = T[5] = T[10] = T[15] = 1;
    if (n <= 0) return;
    double ms[3] = {0, 0, 0}, mt[3] = {0, 0, 0}, C[3][3] = {{0}};
    for (long i = 0; i < n; ++i) for (int c = 0; c < 3; ++c) { ms[c] += src[3 * i + c]; mt[c] += tgt[3 * i + c]; }
    for (int c = 0; c < 3; ++c) { ms[c] /= (double)n; mt[c] /= (double)n; }
    for (long i = 0; i < n; ++i)
        for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) C[a][b] += (tgt[3 * i + a] - mt[a]) * (src[3 * i + b] - ms[b]);
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) C[a][b] /= (double)n;
    double U[3][3], S[3], V[3][3];
    kpo_svd3(C, U, S, V);
    if (!(S[0] > 1e-300)) return;
    double d = kpo_det3(U) * kpo_det3(V) < 0 ? -1.0 : 1.0;
    for (int a = 0; a < 3; ++a) {
        for (int b = 0; b < 3; ++b) T[4 * a + b] = U[a][0] * V[b][0] + U[a][1] * V[b][1] + d * U[a][2] * V[b][2];
    }
    for (int a = 0; a < 3; ++a) T[4 * a + 3] = mt[a] - (T[4 * a] * ms[0] + T[4 * a + 1] * ms[1] + T[4 * a + 2] * ms[2]);
}

/* ---- InitializePointCloudForColoredICP: tangent-plane colour gradient per point (preprocessing/
 * registration.py:108-113 via registration_colored_icp); hybrid search (radius, max_nn), >= 4 neighbours. */
KPO_API int kpo_color_gradient(const float *pts, const float *colors, const float *nrm, long n, double radius, int max_nn,
                               float *inten_out, float *grad_out)
{
    if (!(radius > 0) || max_nn < 1) return -1;
    for (long i = 0; i < n; ++i)
        inten_out[i] = (float)((((double)colors[3 * i] + (double)colors[3 * i + 1]) + (double)colors[3 * i + 2]) / 3.0);
    kpo_grid_t g;
    kpo_grid_build(&g, pts, n, radius);
    double r2 = radius * radius;
#pragma omp parallel
    {
        double *bd = (double *)malloc(sizeof(double) * (size_t)max_nn);
        int32_t *bi = (int32_t *)malloc(sizeof(int32_t) * (size_t)max_nn);
#pragma omp for schedule(dynamic, 256)
        for (long i = 0; i < n; ++i) {
            double x[3] = {0, 0, 0};
            int cnt = 0;
            if (!isnan(pts[3 * i])) kpo_query(&g, pts + 3 * i, max_nn, r2, bd, bi, &cnt);
            if (cnt >= 4) {
                double vx = pts[3 * i], vy = pts[3 * i + 1], vz = pts[3 * i + 2];
                double nx = nrm[3 * i], ny = nrm[3 * i + 1], nz = nrm[3 * i + 2], it = inten_out[i];
                double A[6] = {0, 0, 0, 0, 0, 0}, b[3] = {0, 0, 0};
                for (int t = 1; t < cnt; ++t) {
                    long j = bi[t];
                    double ax = (double)pts[3 * j] - vx, ay = (double)pts[3 * j + 1] - vy, az = (double)pts[3 * j + 2] - vz;
                    double dn = (ax * nx + ay * ny) + az * nz;
                    double px = ax - dn * nx, py = ay - dn * ny, pz = az - dn * nz;
                    double di = (double)inten_out[j] - it;
                    A[0] += px * px; A[1] += px * py; A[2] += px * pz; A[3] += py * py; A[4] += py * pz; A[5] += pz * pz;
                    b[0] += px * di; b[1] += py * di; b[2] += pz * di;
                }
                double w = (double)(cnt - 1);
                A[0] += w * nx * w * nx; A[1] += w * nx * w * ny; A[2] += w * nx * w * nz;
                A[3] += w * ny * w * ny; A[4] += w * ny * w * nz; A[5] += w * nz * w * nz;
                double M[3][4] = {{A[0], A[1], A[2], b[0]}, {A[1], A[3], A[4], b[1]}, {A[2], A[4], A[5], b[2]}};
                int ok = 1;
                for (int c = 0; c < 3 && ok; ++c) {
                    int pv = c;
                    for (int r = c + 1; r < 3; ++r) if (fabs(M[r][c]) > fabs(M[pv][c])) pv = r;
                    if (!(fabs(M[pv][c]) > 1e-300)) { ok = 0; break; }
                    if (pv != c) for (int j = 0; j < 4; ++j) { double tv = M[c][j]; M[c][j] = M[pv][j]; M[pv][j] = tv; }
                    for (int r = c + 1; r < 3; ++r) { double f = M[r][c] / M[c][c]; for (int j = c; j < 4; ++j) M[r][j] -= f * M[c][j]; }
                }
                if (ok) {
                    x[2] = M[2][3] / M[2][2];
                    x[1] = (M[1][3] - M[1][2] * x[2]) / M[1][1];
                    x[0] = (M[0][3] - M[0][1] * x[1] - M[0][2] * x[2]) / M[0][0];
                    if (!(isfinite(x[0]) && isfinite(x[1]) && isfinite(x[2]))) x[0] = x[1] = x[2] = 0;
                }
            }
            grad_out[3 * i] = (float)x[0]; grad_out[3 * i + 1] = (float)x[1]; grad_out[3 * i + 2] = (float)x[2];
        }
        free(bd); free(bi);
    }
    kpo_grid_free(&g);
    return 0;
}

/* ---- the ICP loop of registration_icp for the three estimators the reference uses:
 * mode 0 point-to-plane (preprocessing/registration.py:78-84), 1 point-to-point
 * (manual_pointcloud_registration.py:94-98), 2 coloured (preprocessing/registration.py:108-113;
 * src_int / tgt_int / tgt_grad from kpo_color_gradient, lambda = lambda_geometric). */
KPO_API int kpo_icp(int mode, const float *src, long ns, const float *tgt, const float *tgt_n, long nt,
                    const float *src_int, const float *tgt_int, const float *tgt_grad, double lambda,
                    double max_corr, const double *init, int max_iter, double rel_fit,
                    double rel_rmse, double *T_out, double *fitness_out, double *rmse_out,
                    int *iters_out, int64_t *ncorr_out)
{
    if (!(max_corr > 0) || (mode != 1 && !tgt_n)) return -1;
    kpo_grid_t g;
    kpo_grid_build(&g, tgt, nt, max_corr);
    double r2 = max_corr * max_corr;
    double slg = sqrt(lambda), slp = sqrt(1.0 - lambda);
    double *cur = (double *)malloc(sizeof(double) * 3 * (size_t)(ns ? ns : 1));
    double *ms = (double *)malloc(sizeof(double) * 3 * (size_t)(ns ? ns : 1));
    double *mt = (double *)malloc(sizeof(double) * 3 * (size_t)(ns ? ns : 1));
    int32_t *corr = (int32_t *)malloc(sizeof(int32_t) * (size_t)(ns ? ns : 1));
    double T[16];
    memcpy(T, init, sizeof T);
    for (long i = 0; i < ns; ++i) {
        double X = src[3 * i], Y = src[3 * i + 1], Z = src[3 * i + 2];
        cur[3 * i] = ((T[0] * X + T[1] * Y) + T[2] * Z) + T[3];
        cur[3 * i + 1] = ((T[4] * X + T[5] * Y) + T[6] * Z) + T[7];
        cur[3 * i + 2] = ((T[8] * X + T[9] * Y) + T[10] * Z) + T[11];
    }
    double fit = 0, rmse = 0, pfit = 0, prmse = 0;
    long nc = 0;
    int it_done = 0;
    for (int pass = 0; pass <= max_iter; ++pass) {
        /* correspondences: 1-NN in target with d^2 < max_corr^2 (query is double) */
        double err2 = 0; nc = 0;
#pragma omp parallel for schedule(dynamic, 256)
        for (long i = 0; i < ns; ++i) {
            corr[i] = -1;
            if (isnan(cur[3 * i]) || g.dim[0] == 0) continue;
            int c0[3];
            for (int c = 0; c < 3; ++c) c0[c] = kpo_cell_of(&g, cur[3 * i + c], c);
            int R = (int)ceil(max_corr / g.cell) + 1;
            double bd = INFINITY; int32_t bi = -1;
            for (int dx = -R; dx <= R; ++dx) for (int dy = -R; dy <= R; ++dy) for (int dz = -R; dz <= R; ++dz) {
                int cx = c0[0] + dx, cy = c0[1] + dy, cz = c0[2] + dz;
                if (cx < 0 || cy < 0 || cz < 0 || cx >= g.dim[0] || cy >= g.dim[1] || cz >= g.dim[2]) continue;
                long cid = ((long)cx * g.dim[1] + cy) * g.dim[2] + cz;
                for (int64_t t = g.start[cid]; t < g.start[cid + 1]; ++t) {
                    int32_t id = g.order[t];
                    double ex = cur[3 * i] - (double)tgt[3 * (long)id], ey = cur[3 * i + 1] - (double)tgt[3 * (long)id + 1], ez = cur[3 * i + 2] - (double)tgt[3 * (long)id + 2];
                    double d2 = (ex * ex + ey * ey) + ez * ez;
                    if (d2 < r2 && (d2 < bd || (d2 == bd && id < bi))) { bd = d2; bi = id; }
                }
            }
            corr[i] = bi;
        }
        double A[6][6] = {{0}}, b[6] = {0};
        for (long i = 0; i < ns; ++i) {
            if (corr[i] < 0) continue;
            long j = corr[i];
            double sx = cur[3 * i], sy = cur[3 * i + 1], sz = cur[3 * i + 2];
            double ex = sx - (double)tgt[3 * j], ey = sy - (double)tgt[3 * j + 1], ez = sz - (double)tgt[3 * j + 2];
            err2 += (ex * ex + ey * ey) + ez * ez;
            if (mode == 1) {
                ms[3 * nc] = sx; ms[3 * nc + 1] = sy; ms[3 * nc + 2] = sz;
                mt[3 * nc] = tgt[3 * j]; mt[3 * nc + 1] = tgt[3 * j + 1]; mt[3 * nc + 2] = tgt[3 * j + 2];
                ++nc;
                continue;
            }
            ++nc;
            double nx = tgt_n[3 * j], ny = tgt_n[3 * j + 1], nz = tgt_n[3 * j + 2];
            double r = (ex * nx + ey * ny) + ez * nz;
            double J[6] = {sy * nz - sz * ny, sz * nx - sx * nz, sx * ny - sy * nx, nx, ny, nz};
            if (mode == 2) {
                double is = src_int[i], it = tgt_int[j];
                double gx = tgt_grad[3 * j], gy = tgt_grad[3 * j + 1], gz = tgt_grad[3 * j + 2];
                double px = ex - r * nx, py = ey - r * ny, pz = ez - r * nz;       /* vs_proj - vt */
                double is_proj = ((gx * px + gy * py) + gz * pz) + it;
                double gn = (gx * nx + gy * ny) + gz * nz;
                double mx = -(gx - gn * nx), my = -(gy - gn * ny), mz = -(gz - gn * nz);   /* -dit^T (I - n n^T) */
                double J2[6] = {slp * (sy * mz - sz * my), slp * (sz * mx - sx * mz), slp * (sx * my - sy * mx), slp * mx, slp * my, slp * mz};
                double r2c = slp * (is - is_proj);
                for (int p = 0; p < 6; ++p) J[p] *= slg;
                r *= slg;
                for (int p = 0; p < 6; ++p) { b[p] += J2[p] * r2c; for (int q = 0; q < 6; ++q) A[p][q] += J2[p] * J2[q]; }
            }
            for (int p = 0; p < 6; ++p) { b[p] += J[p] * r; for (int q = 0; q < 6; ++q) A[p][q] += J[p] * J[q]; }
        }
        pfit = fit; prmse = rmse;
        fit = ns ? (double)nc / (double)ns : 0.0;
        rmse = nc ? sqrt(err2 / (double)nc) : 0.0;
        if (pass > 0) {
            it_done = pass;
            if (fabs(pfit - fit) < rel_fit && fabs(prmse - rmse) < rel_rmse) break;
        }
        if (pass == max_iter) break;
        double x[6], U[16];
        memset(U, 0, sizeof U); U[0] = U[5] = U[10] = U[15] = 1;
        if (mode == 1) {
            if (nc > 0) kpo_umeyama(ms, mt, nc, U);
        } else {
            for (int p = 0; p < 6; ++p) b[p] = -b[p];
            if (nc > 0 && kpo_solve6(A, b, x)) kpo_x6_to_mat4(x, U);
        }
        kpo_mat4_mul(U, T, T);
        for (long i = 0; i < ns; ++i) {
            double X = cur[3 * i], Y = cur[3 * i + 1], Z = cur[3 * i + 2];
            cur[3 * i] = ((U[0] * X + U[1] * Y) + U[2] * Z) + U[3];
            cur[3 * i + 1] = ((U[4] * X + U[5] * Y) + U[6] * Z) + U[7];
            cur[3 * i + 2] = ((U[8] * X + U[9] * Y) + U[10] * Z) + U[11];
        }
    }
    memcpy(T_out, T, sizeof T);
    if (fitness_out) *fitness_out = fit;
    if (rmse_out) *rmse_out = rmse;
    if (iters_out) *iters_out = it_done;
    if (ncorr_out) *ncorr_out = nc;
    free(cur); free(corr); free(ms); free(mt);
    kpo_grid_free(&g);
    return 0;
}
KPO_API int kpo_icp_point_to_plane(const float *src, long ns, const float *tgt, const float *tgt_n, long nt,
                                   double max_corr, const double *init, int max_iter, double rel_fit,
                                   double rel_rmse, double *T_out, double *fitness_out, double *rmse_out,
                                   int *iters_out, int64_t *ncorr_out)
{
    return kpo_icp(0, src, ns, tgt, tgt_n, nt, NULL, NULL, NULL, 1.0, max_corr, init, max_iter, rel_fit, rel_rmse, T_out,
                   fitness_out, rmse_out, iters_out, ncorr_out);
}

/* ------------------------------------------- global registration (f3) -- */
/* compute_fpfh_feature (preprocessing/registration.py:17-20): Open3D's ComputePairFeatures /
 * ComputeSPFHFeature / ComputeFPFHFeature restated; hybrid neighbours in canonical (d2, index) order. */
static void kpo_pair_features(const double *p1, const double *n1, const double *p2, const double *n2, double *f)
{
    double d[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
    double len = sqrt((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]);
    f[0] = f[1] = f[2] = 0.0;
    if (len == 0.0) return;
    double a[3] = {n1[0], n1[1], n1[2]}, b[3] = {n2[0], n2[1], n2[2]};
    double angle1 = ((a[0] * d[0] + a[1] * d[1]) + a[2] * d[2]) / len;
    double angle2 = ((b[0] * d[0] + b[1] * d[1]) + b[2] * d[2]) / len;
    if (acos(fabs(angle1)) > acos(fabs(angle2))) {
        for (int c = 0; c < 3; ++c) { double t = a[c]; a[c] = b[c]; b[c] = t; d[c] = -d[c]; }
        f[2] = -angle2;
    } else f[2] = angle1;
    double v[3] = {d[1] * a[2] - d[2] * a[1], d[2] * a[0] - d[0] * a[2], d[0] * a[1] - d[1] * a[0]};
    double vn = sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]);
    if (vn == 0.0) { f[2] = 0.0; return; }
    for (int c = 0; c < 3; ++c) v[c] /= vn;
    double w[3] = {a[1] * v[2] - a[2] * v[1], a[2] * v[0] - a[0] * v[2], a[0] * v[1] - a[1] * v[0]};
    f[1] = (v[0] * b[0] + v[1] * b[1]) + v[2] * b[2];
    f[0] = atan2((w[0] * b[0] + w[1] * b[1]) + w[2] * b[2], (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]);
}
static int kpo_bin11(double x) { int h = (int)floor(x); return h < 0 ? 0 : (h > 10 ? 10 : h); }
KPO_API int kpo_fpfh(const float *pts, const float *nrm, long n, double radius, int max_nn, double *feat /* [n][33] */)
{
    if (max_nn < 1 || !(radius > 0)) return -1;
    kpo_grid_t g;
    kpo_grid_build(&g, pts, n, radius);
    double r2 = radius * radius;
    int32_t *idx = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n ? n : 1) * (size_t)max_nn);
    double *d2 = (double *)malloc(sizeof(double) * (size_t)(n ? n : 1) * (size_t)max_nn);
    int *cnt = (int *)calloc((size_t)(n ? n : 1), sizeof(int));
    double *spfh = (double *)calloc((size_t)(n ? n : 1) * 33, sizeof(double));
#pragma omp parallel for schedule(dynamic, 256)
    for (long i = 0; i < n; ++i) {
        if (isnan(pts[3 * i])) continue;
        kpo_query(&g, pts + 3 * i, max_nn, r2, d2 + (size_t)i * max_nn, idx + (size_t)i * max_nn, &cnt[i]);
        int c = cnt[i];
        if (c <= 1) continue;
        double p1[3] = {pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]}, n1[3] = {nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]};
        double incr = 100.0 / (double)(c - 1);
        double *h = spfh + (size_t)i * 33;
        for (int t = 1; t < c; ++t) {
            long j = idx[(size_t)i * max_nn + t];
            double p2[3] = {pts[3 * j], pts[3 * j + 1], pts[3 * j + 2]}, n2[3] = {nrm[3 * j], nrm[3 * j + 1], nrm[3 * j + 2]};
            double f[3];
            kpo_pair_features(p1, n1, p2, n2, f);
            h[kpo_bin11(11.0 * (f[0] + M_PI) / (2.0 * M_PI))] += incr;
            h[11 + kpo_bin11(11.0 * (f[1] + 1.0) * 0.5)] += incr;
            h[22 + kpo_bin11(11.0 * (f[2] + 1.0) * 0.5)] += incr;
        }
    }
#pragma omp parallel for schedule(dynamic, 256)
    for (long i = 0; i < n; ++i) {
        double *f = feat + (size_t)i * 33;
        for (int j = 0; j < 33; ++j) f[j] = 0.0;
        int c = cnt[i];
        if (c <= 1) continue;
        double sum[3] = {0, 0, 0};
        for (int t = 1; t < c; ++t) {
            double dist = d2[(size_t)i * max_nn + t];
            if (dist == 0.0) continue;
            const double *sp = spfh + (size_t)idx[(size_t)i * max_nn + t] * 33;
            for (int j = 0; j < 33; ++j) { double val = sp[j] / dist; sum[j / 11] += val; f[j] += val; }
        }
        for (int q = 0; q < 3; ++q) if (sum[q] != 0.0) sum[q] = 100.0 / sum[q];
        for (int j = 0; j < 33; ++j) f[j] = f[j] * sum[j / 11] + spfh[(size_t)i * 33 + j];
    }
    free(idx); free(d2); free(cnt); free(spfh);
    kpo_grid_free(&g);
    return 0;
}
/* the two KD-tree searches inside registration_ransac_based_on_feature_matching: exact 1-NN in feature
 * space, squared L2 in double summed in dimension order, ties to the lower index */
KPO_API void kpo_feature_match(const double *fa, long na, const double *fb, long nb, int dim, int32_t *nn, double *nn_d2)
{
#pragma omp parallel for schedule(static)
    for (long i = 0; i < na; ++i) {
        double best = INFINITY; int32_t bi = -1;
        for (long t = 0; t < nb; ++t) {
            double s = 0.0;
            for (int j = 0; j < dim; ++j) { double d = fa[(size_t)i * dim + j] - fb[(size_t)t * dim + j]; s += d * d; }
            if (s < best) { best = s; bi = (int32_t)t; }
        }
        nn[i] = bi;
        if (nn_d2) nn_d2[i] = best;
    }
}
/* RegistrationRANSACBasedOnCorrespondence (second half of registration_ransac_based_on_feature_matching,
 * preprocessing/registration.py:50-57): sequential over the hypothesis index with the counter-based sampler */
KPO_API int kpo_ransac_correspondence(const float *src, const float *tgt, const int32_t *corres, long m, double max_corr,
                                      int ransac_n, double edge_sim, double dist_thr, int max_iter, double confidence,
                                      uint64_t seed, double *T_out, double *fitness_out, double *rmse_out, int32_t *best_iter_out,
                                      int64_t *validated_out)
{
    memset(T_out, 0, sizeof(double) * 16);
    T_out[0] = T_out[5] = T_out[10] = T_out[15] = 1;
    *fitness_out = 0; *rmse_out = 0; *best_iter_out = -1; *validated_out = 0;
    if (ransac_n < 3 || ransac_n > 8 || !(max_corr > 0)) return -1;
    if (m < ransac_n) return 0;
    double max_d2 = max_corr * max_corr, best_fit = 0, best_rmse = 0, exit_itr = (double)max_iter;
    /* validity and scores of every hypothesis (the early exit only skips the comparison, as on the device, so that
     * `validated` counts the same thing) */
    char *valid = (char *)calloc((size_t)(max_iter ? max_iter : 1), 1);
    double *Ts = (double *)malloc(sizeof(double) * 16 * (size_t)(max_iter ? max_iter : 1));
#pragma omp parallel for schedule(dynamic, 1024)
    for (int h = 0; h < max_iter; ++h) {
        double s[8][3], t[8][3];
        for (int j = 0; j < ransac_n; ++j) {
            long c = (long)(kpo_rng(seed, (uint64_t)h, (uint64_t)j) % (uint64_t)m);
            for (int d = 0; d < 3; ++d) { s[j][d] = src[3 * (long)corres[2 * c] + d]; t[j][d] = tgt[3 * (long)corres[2 * c + 1] + d]; }
        }
        int ok = 1;
        if (edge_sim > 0)
            for (int i = 0; i < ransac_n && ok; ++i) for (int j = i + 1; j < ransac_n; ++j) {
                double a0 = s[i][0] - s[j][0], a1 = s[i][1] - s[j][1], a2 = s[i][2] - s[j][2];
                double b0 = t[i][0] - t[j][0], b1 = t[i][1] - t[j][1], b2 = t[i][2] - t[j][2];
                double ds = sqrt((a0 * a0 + a1 * a1) + a2 * a2), dt = sqrt((b0 * b0 + b1 * b1) + b2 * b2);
                if (ds < dt * edge_sim || dt < ds * edge_sim) { ok = 0; break; }
            }
        if (!ok) continue;
        double *T = Ts + 16 * (size_t)h;
        kpo_umeyama(&s[0][0], &t[0][0], ransac_n, T);
        if (dist_thr > 0)
            for (int j = 0; j < ransac_n; ++j) {
                double x = ((T[0] * s[j][0] + T[1] * s[j][1]) + T[2] * s[j][2]) + T[3];
                double y = ((T[4] * s[j][0] + T[5] * s[j][1]) + T[6] * s[j][2]) + T[7];
                double z = ((T[8] * s[j][0] + T[9] * s[j][1]) + T[10] * s[j][2]) + T[11];
                double e0 = x - t[j][0], e1 = y - t[j][1], e2 = z - t[j][2];
                if (sqrt((e0 * e0 + e1 * e1) + e2 * e2) > dist_thr) { ok = 0; break; }
            }
        valid[h] = (char)ok;
    }
    for (int h = 0; h < max_iter; ++h) {
        if (!valid[h]) continue;
        ++*validated_out;
        if ((double)h >= exit_itr) continue;
        const double *T = Ts + 16 * (size_t)h;
        long good = 0; double err2 = 0;
        for (long c = 0; c < m; ++c) {
            const float *sp = src + 3 * (long)corres[2 * c], *tp = tgt + 3 * (long)corres[2 * c + 1];
            double x = ((T[0] * sp[0] + T[1] * sp[1]) + T[2] * sp[2]) + T[3];
            double y = ((T[4] * sp[0] + T[5] * sp[1]) + T[6] * sp[2]) + T[7];
            double z = ((T[8] * sp[0] + T[9] * sp[1]) + T[10] * sp[2]) + T[11];
            double e0 = x - tp[0], e1 = y - tp[1], e2 = z - tp[2];
            double d2 = (e0 * e0 + e1 * e1) + e2 * e2;
            if (d2 < max_d2) { ++good; err2 += d2; }
        }
        double fit = good ? (double)good / (double)m : 0.0, rmse = good ? sqrt(err2 / (double)good) : 0.0;
        if (fit > best_fit || (fit == best_fit && rmse < best_rmse)) {
            best_fit = fit; best_rmse = rmse; *best_iter_out = h;
            memcpy(T_out, T, sizeof(double) * 16);
            if (confidence < 1.0) {
                double k_est = log(1.0 - confidence) / log(1.0 - pow(fit, (double)ransac_n));
                if (k_est < exit_itr) exit_itr = ceil(k_est);
            }
        }
    }
    *fitness_out = best_fit; *rmse_out = best_rmse;
    free(valid); free(Ts);
    return 0;
}

/* threads the OpenMP loops of this library use from now on (a launcher may have exported OMP_NUM_THREADS=1) */
KPO_API void kpo_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

KPO_API int kpo_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
