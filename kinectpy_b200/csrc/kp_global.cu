// kp_global.cu -- global registration: FPFH features, feature matching and the
// RANSAC over correspondences that initialises the ICP.
// Replaces, at preprocessing/registration.py:15-21 and :50-57,
//   o3d.pipelines.registration.compute_fpfh_feature(pcd_down, KDTreeSearchParamHybrid(5 * voxel, fpfh_nn))
//   o3d.pipelines.registration.registration_ransac_based_on_feature_matching(source_down, target_down,
//       source_fpfh, target_fpfh, True, 1.5 * voxel, TransformationEstimationPointToPoint(False), 3,
//       [CorrespondenceCheckerBasedOnEdgeLength(0.95), CorrespondenceCheckerBasedOnDistance(1.5 * voxel)],
//       RANSACConvergenceCriteria(250000, 0.999)).
//
// FPFH: the hybrid neighbour lists come from the grid search of kp_grid.cu (ascending (d2, index), self
// first); one thread per point builds its SPFH (three 11-bin histograms of the Darboux-frame angles to each
// neighbour), a second pass blends the neighbours' SPFHs weighted by 1 / d2 and adds the point's own.  Sums
// run sequentially in neighbour order, in double.
// Matching: exact 1-NN in the 33-dimensional feature space by a tiled brute-force scan (target tile staged in
// shared memory, one query per thread, double, fixed summation order, ties to the lower index).
// RANSAC: upstream draws from a global generator inside an OpenMP loop; here hypothesis h draws its ransac_n
// correspondences with the counter-based kp_rng(seed, h, j) (with replacement, as upstream), ALL max_iteration
// hypotheses are checked (edge length, then the alignment and the distance check) in one launch, the survivors
// are scored against every correspondence in a second, and the host replays upstream's sequential
// "better than the best so far" rule with its confidence-based early exit over the scores.
#include <math.h>
#include <string.h>
#include <vector>
#include "kp_grid.cuh"
#include "kp_umeyama.cuh"

namespace {
constexpr int FPFH_DIM = 33;

// ComputePairFeatures: (atan2 term, v.n2, cos angle of n1 with the connecting line); zero when degenerate
__device__ __forceinline__ bool pair_features(const double *p1, const double *n1, const double *p2, const double *n2, double *f)
{
    double d[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
    const double len = sqrt((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]);
    f[0] = f[1] = f[2] = 0.0;
    if (len == 0.0) return false;
    double a[3] = {n1[0], n1[1], n1[2]}, b[3] = {n2[0], n2[1], n2[2]};
    const double angle1 = ((a[0] * d[0] + a[1] * d[1]) + a[2] * d[2]) / len;
    const double angle2 = ((b[0] * d[0] + b[1] * d[1]) + b[2] * d[2]) / len;
    if (acos(fabs(angle1)) > acos(fabs(angle2))) {
        for (int c = 0; c < 3; ++c) { const double t = a[c]; a[c] = b[c]; b[c] = t; d[c] = -d[c]; }
        f[2] = -angle2;
    } else {
        f[2] = angle1;
    }
    double v[3] = {d[1] * a[2] - d[2] * a[1], d[2] * a[0] - d[0] * a[2], d[0] * a[1] - d[1] * a[0]};
    const double vn = sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]);
    if (vn == 0.0) { f[2] = 0.0; return false; }
    for (int c = 0; c < 3; ++c) v[c] /= vn;
    const double w[3] = {a[1] * v[2] - a[2] * v[1], a[2] * v[0] - a[0] * v[2], a[0] * v[1] - a[1] * v[0]};
    f[1] = (v[0] * b[0] + v[1] * b[1]) + v[2] * b[2];
    f[0] = atan2((w[0] * b[0] + w[1] * b[1]) + w[2] * b[2], (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]);
    return true;
}
__device__ __forceinline__ int bin11(double x)
{
    int h = (int)floor(x);
    return h < 0 ? 0 : (h > 10 ? 10 : h);
}

__global__ void __launch_bounds__(128) k_spfh(const float *xyz, const float *nrm, int64_t n, const int32_t *idx, const int32_t *cnt,
                                              int k, double *spfh)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double h[FPFH_DIM];
#pragma unroll
    for (int j = 0; j < FPFH_DIM; ++j) h[j] = 0.0;
    const int c = cnt[i];
    if (c > 1) {
        const double p1[3] = {(double)xyz[3 * i], (double)xyz[3 * i + 1], (double)xyz[3 * i + 2]};
        const double n1[3] = {(double)nrm[3 * i], (double)nrm[3 * i + 1], (double)nrm[3 * i + 2]};
        const double incr = 100.0 / (double)(c - 1);
        const double two_pi = 6.283185307179586476925286766559, pi = 3.14159265358979323846;
        for (int t = 1; t < c; ++t) {
            const int64_t j = idx[i * k + t];
            const double p2[3] = {(double)xyz[3 * j], (double)xyz[3 * j + 1], (double)xyz[3 * j + 2]};
            const double n2[3] = {(double)nrm[3 * j], (double)nrm[3 * j + 1], (double)nrm[3 * j + 2]};
            double f[3];
            pair_features(p1, n1, p2, n2, f);
            // dynamic bin index: the histogram lives in local memory (L1), 3 updates per neighbour
            h[bin11(11.0 * (f[0] + pi) / two_pi)] += incr;
            h[11 + bin11(11.0 * (f[1] + 1.0) * 0.5)] += incr;
            h[22 + bin11(11.0 * (f[2] + 1.0) * 0.5)] += incr;
        }
    }
    for (int j = 0; j < FPFH_DIM; ++j) spfh[i * FPFH_DIM + j] = h[j];
}

__global__ void __launch_bounds__(128) k_fpfh(int64_t n, const int32_t *idx, const double *d2, const int32_t *cnt, int k,
                                              const double *spfh, double *out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double f[FPFH_DIM];
#pragma unroll
    for (int j = 0; j < FPFH_DIM; ++j) f[j] = 0.0;
    const int c = cnt[i];
    if (c > 1) {
        double sum[3] = {0.0, 0.0, 0.0};
        for (int t = 1; t < c; ++t) {
            const double dist = d2[i * k + t];
            if (dist == 0.0) continue;
            const double *sp = spfh + (int64_t)idx[i * k + t] * FPFH_DIM;
#pragma unroll
            for (int j = 0; j < FPFH_DIM; ++j) {
                const double val = sp[j] / dist;
                sum[j / 11] += val;
                f[j] += val;
            }
        }
#pragma unroll
        for (int g = 0; g < 3; ++g) if (sum[g] != 0.0) sum[g] = 100.0 / sum[g];
#pragma unroll
        for (int j = 0; j < FPFH_DIM; ++j) f[j] = f[j] * sum[j / 11] + spfh[i * FPFH_DIM + j];
    }
#pragma unroll
    for (int j = 0; j < FPFH_DIM; ++j) out[i * FPFH_DIM + j] = f[j];
}

// ---- exact 1-NN in feature space: one query per thread, targets staged tile by tile in shared memory
constexpr int FM_THREADS = 128;
constexpr int FM_TILE = 64;
template <int DIM>
__global__ void __launch_bounds__(FM_THREADS) k_feat_nn(const double *qa, int64_t na, const double *tb, int64_t nb, int32_t *nn,
                                                       double *nn_d2)
{
    __shared__ double tile[FM_TILE * DIM];
    const int64_t i = (int64_t)blockIdx.x * FM_THREADS + threadIdx.x;
    double q[DIM];
    const bool live = i < na;
#pragma unroll
    for (int j = 0; j < DIM; ++j) q[j] = live ? qa[i * DIM + j] : 0.0;
    double best = INFINITY;
    int bi = -1;
    for (int64_t t0 = 0; t0 < nb; t0 += FM_TILE) {
        const int m = (int)(nb - t0 < FM_TILE ? nb - t0 : FM_TILE);
        __syncthreads();
        for (int e = threadIdx.x; e < m * DIM; e += FM_THREADS) tile[e] = tb[t0 * DIM + e];
        __syncthreads();
        for (int t = 0; t < m; ++t) {
            double s = 0.0;
#pragma unroll
            for (int j = 0; j < DIM; ++j) { const double d = q[j] - tile[t * DIM + j]; s += d * d; }
            if (s < best) { best = s; bi = (int)(t0 + t); }     // ascending scan: ties keep the lower index
        }
    }
    if (live) { nn[i] = bi; if (nn_d2) nn_d2[i] = best; }
}

// ---- RANSAC over correspondences
constexpr int GR_MAXN = 8;
struct GrParams {
    const float *src, *tgt;
    const int32_t *corres;    // [m][2]
    int m, ransac_n, max_iter;
    double edge_sim;          // CorrespondenceCheckerBasedOnEdgeLength threshold, <= 0: off
    double dist_thr;          // CorrespondenceCheckerBasedOnDistance threshold, <= 0: off
    double max_d2;
    uint64_t seed;
    uint8_t *valid;           // [max_iter]
    double *T;                // [max_iter][12]
};

__global__ void __launch_bounds__(128) k_gr_hypotheses(const __grid_constant__ GrParams p)
{
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= p.max_iter) return;
    double s[GR_MAXN][3], t[GR_MAXN][3];
    const int rn = p.ransac_n;
    for (int j = 0; j < rn; ++j) {
        const int c = (int)(kp_rng(p.seed, (uint64_t)h, (uint64_t)j) % (uint64_t)p.m);
        const int a = p.corres[2 * c], b = p.corres[2 * c + 1];
        for (int d = 0; d < 3; ++d) { s[j][d] = (double)p.src[3 * (int64_t)a + d]; t[j][d] = (double)p.tgt[3 * (int64_t)b + d]; }
    }
    bool ok = true;
    if (p.edge_sim > 0.0) {
        for (int i = 0; i < rn && ok; ++i)
            for (int j = i + 1; j < rn; ++j) {
                const double ds = sqrt(kp_d2(s[i][0], s[i][1], s[i][2], s[j][0], s[j][1], s[j][2]));
                const double dt = sqrt(kp_d2(t[i][0], t[i][1], t[i][2], t[j][0], t[j][1], t[j][2]));
                if (ds < dt * p.edge_sim || dt < ds * p.edge_sim) { ok = false; break; }
            }
    }
    double Un[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    if (ok) {
        double tot[15];
        for (int e = 0; e < 15; ++e) tot[e] = 0.0;
        for (int j = 0; j < rn; ++j) {
            for (int d = 0; d < 3; ++d) { tot[d] += s[j][d]; tot[3 + d] += t[j][d]; }
            for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) tot[6 + 3 * a + b] += t[j][a] * s[j][b];
        }
        kp_umeyama(tot, (double)rn, Un);
        if (p.dist_thr > 0.0) {
            for (int j = 0; j < rn; ++j) {
                const double x = kp_affine(Un[0], Un[1], Un[2], Un[3], s[j][0], s[j][1], s[j][2]);
                const double y = kp_affine(Un[4], Un[5], Un[6], Un[7], s[j][0], s[j][1], s[j][2]);
                const double z = kp_affine(Un[8], Un[9], Un[10], Un[11], s[j][0], s[j][1], s[j][2]);
                if (sqrt(kp_d2(x, y, z, t[j][0], t[j][1], t[j][2])) > p.dist_thr) { ok = false; break; }
            }
        }
    }
    p.valid[h] = ok ? 1 : 0;
    if (ok) for (int e = 0; e < 12; ++e) p.T[(int64_t)h * 12 + e] = Un[e];
}

// one surviving hypothesis per thread; correspondences staged in shared memory; sums in correspondence order
constexpr int GS_THREADS = 128;
constexpr int GS_TILE = 256;
__global__ void __launch_bounds__(GS_THREADS) k_gr_score(const __grid_constant__ GrParams p, const int32_t *list, const int32_t *n_list,
                                                         int32_t *good_out, double *err2_out)
{
    __shared__ float sm_s[GS_TILE][3], sm_t[GS_TILE][3];
    const int nl = *n_list;
    const int w = blockIdx.x * GS_THREADS + threadIdx.x;
    const bool live = w < nl;
    double T[12];
    const int h = live ? list[w] : 0;
    for (int e = 0; e < 12; ++e) T[e] = live ? p.T[(int64_t)h * 12 + e] : 0.0;
    int good = 0;
    double err2 = 0.0;
    for (int c0 = 0; c0 < p.m; c0 += GS_TILE) {
        const int mm = p.m - c0 < GS_TILE ? p.m - c0 : GS_TILE;
        __syncthreads();
        for (int e = threadIdx.x; e < mm; e += GS_THREADS) {
            const int a = p.corres[2 * (c0 + e)], b = p.corres[2 * (c0 + e) + 1];
            for (int d = 0; d < 3; ++d) { sm_s[e][d] = p.src[3 * (int64_t)a + d]; sm_t[e][d] = p.tgt[3 * (int64_t)b + d]; }
        }
        __syncthreads();
        if (!live) continue;
        for (int e = 0; e < mm; ++e) {
            const double sx = (double)sm_s[e][0], sy = (double)sm_s[e][1], sz = (double)sm_s[e][2];
            const double x = kp_affine(T[0], T[1], T[2], T[3], sx, sy, sz);
            const double y = kp_affine(T[4], T[5], T[6], T[7], sx, sy, sz);
            const double z = kp_affine(T[8], T[9], T[10], T[11], sx, sy, sz);
            const double d2 = kp_d2(x, y, z, (double)sm_t[e][0], (double)sm_t[e][1], (double)sm_t[e][2]);
            if (d2 < p.max_d2) { ++good; err2 = __dadd_rn(err2, d2); }
        }
    }
    if (live) { good_out[w] = good; err2_out[w] = err2; }
}
}  // namespace

extern "C" {

int kp_fpfh(kp_ctx *ctx, const float *d_xyz, const float *d_normals, int64_t n, double radius, int max_nn, double *d_feat)
{
    if (!ctx || (n > 0 && (!d_xyz || !d_normals || !d_feat))) return kp_set_err(ctx, KP_E_ARG, "kp_fpfh: NULL argument");
    if (max_nn < 1) return kp_set_err(ctx, KP_E_ARG, "compute_fpfh_feature: max_nn < 1");
    kp_enter(ctx);
    if (n <= 0) return KP_OK;
    double cell;
    if (radius > 0) cell = radius * (1.0 + 4e-6);
    else KP_TRY(kp_grid_auto_cell(ctx, d_xyz, n, nullptr, 0.5 * max_nn > 4 ? 0.5 * max_nn : 4, &cell));
    KpGrid g;
    KP_TRY(kp_grid_build_knn(ctx, d_xyz, n, cell, max_nn, nullptr, &g));
    int32_t *idx, *cnt;
    double *d2, *spfh;
    KP_TRY(kp_ws(ctx, (size_t)n * (size_t)max_nn, &idx));
    KP_TRY(kp_ws(ctx, (size_t)n * (size_t)max_nn, &d2));
    KP_TRY(kp_ws(ctx, (size_t)n, &cnt));
    KP_TRY(kp_ws(ctx, (size_t)n * FPFH_DIM, &spfh));
    KP_TRY(kp_knn_device(ctx, g, nullptr, n, max_nn, radius, idx, d2, cnt, nullptr, d_xyz));
    KP_PROFB(ctx, "fpfh", (double)n * (12.0 * max_nn * (4.0 + 24.0) + 2.0 * 8.0 * FPFH_DIM * (1.0 + max_nn)));
    k_spfh<<<kp_blocks(n, 128), 128, 0, ctx->stream>>>(d_xyz, d_normals, n, idx, cnt, max_nn, spfh);
    KP_LAUNCH_CHECK(ctx);
    k_fpfh<<<kp_blocks(n, 128), 128, 0, ctx->stream>>>(n, idx, d2, cnt, max_nn, spfh, d_feat);
    KP_LAUNCH_CHECK(ctx);
    return KP_OK;
}

int kp_feature_match(kp_ctx *ctx, const double *d_feat_a, int64_t na, const double *d_feat_b, int64_t nb, int dim,
                     int32_t *d_nn, double *d_nn_d2)
{
    if (!ctx || (na > 0 && (!d_feat_a || !d_nn)) || (nb > 0 && !d_feat_b)) return kp_set_err(ctx, KP_E_ARG, "kp_feature_match: NULL argument");
    if (dim != FPFH_DIM) return kp_set_err(ctx, KP_E_ARG, "kp_feature_match: only 33-dimensional (FPFH) features are supported");
    if (nb > 2147483000LL || na > 2147483000LL) return kp_set_err(ctx, KP_E_ARG, "kp_feature_match: more than 2^31 features");
    kp_enter(ctx);
    if (na <= 0) return KP_OK;
    KP_PROFB(ctx, "feature_match", (double)na * 8.0 * FPFH_DIM + (double)kp_blocks(na, FM_THREADS) * (double)nb * 8.0 * FPFH_DIM);
    k_feat_nn<FPFH_DIM><<<kp_blocks(na, FM_THREADS), FM_THREADS, 0, ctx->stream>>>(d_feat_a, na, d_feat_b, nb, d_nn, d_nn_d2);
    KP_LAUNCH_CHECK(ctx);
    return KP_OK;
}

int kp_ransac_correspondence(kp_ctx *ctx, const float *d_src, int64_t n_src, const float *d_tgt, int64_t n_tgt,
                             const int32_t *d_corres, int64_t m, double max_corr, int ransac_n, double edge_similarity,
                             double distance_threshold, int max_iteration, double confidence, uint64_t seed, double *h_T16,
                             double *h_fitness, double *h_rmse, int32_t *h_best_iter, int64_t *h_validated)
{
    if (!ctx || !h_T16) return kp_set_err(ctx, KP_E_ARG, "kp_ransac_correspondence: NULL argument");
    if (ransac_n < 3 || ransac_n > GR_MAXN || !(max_corr > 0.0) || max_iteration < 0 || !(confidence >= 0.0 && confidence <= 1.0))
        return kp_set_err(ctx, KP_E_ARG, "registration_ransac_based_on_correspondence: need 3 <= ransac_n <= 8, max_correspondence_distance > 0, "
                                         "0 <= confidence <= 1");
    kp_enter(ctx);
    for (int i = 0; i < 16; ++i) h_T16[i] = (i % 5 == 0) ? 1.0 : 0.0;
    if (h_fitness) *h_fitness = 0.0;
    if (h_rmse) *h_rmse = 0.0;
    if (h_best_iter) *h_best_iter = -1;
    if (h_validated) *h_validated = 0;
    (void)n_src; (void)n_tgt;
    if (m < ransac_n || max_iteration == 0) return KP_OK;
    if (m > 2147483000LL) return kp_set_err(ctx, KP_E_ARG, "too many correspondences");
    GrParams p;
    p.src = d_src; p.tgt = d_tgt; p.corres = d_corres; p.m = (int)m; p.ransac_n = ransac_n; p.max_iter = max_iteration;
    p.edge_sim = edge_similarity; p.dist_thr = distance_threshold; p.max_d2 = max_corr * max_corr; p.seed = seed;
    int32_t *list, *d_n, *good;
    double *err2;
    KP_TRY(kp_ws(ctx, (size_t)max_iteration, &p.valid));
    KP_TRY(kp_ws(ctx, (size_t)max_iteration * 12, &p.T));
    KP_TRY(kp_ws(ctx, (size_t)max_iteration, &list));
    KP_TRY(kp_ws(ctx, 4, &d_n));
    {
        KP_PROFB(ctx, "gr_hypotheses", (double)max_iteration * (ransac_n * 32.0 + 97.0));
        k_gr_hypotheses<<<kp_blocks(max_iteration, 128), 128, 0, ctx->stream>>>(p);
        KP_LAUNCH_CHECK(ctx);
    }
    KP_TRY(kp_prim_compact_mask(ctx, max_iteration, p.valid, 0, nullptr, nullptr, list, d_n));
    KP_CUDA(ctx, cudaMemcpyAsync(ctx->d_scratch, d_n, sizeof(int32_t), cudaMemcpyDeviceToDevice, ctx->stream));
    KP_TRY(kp_fetch_scratch(ctx, sizeof(int32_t)));
    const int nl = *(int32_t *)ctx->h_scratch;
    if (h_validated) *h_validated = nl;
    if (nl <= 0) return KP_OK;
    KP_TRY(kp_ws(ctx, (size_t)nl, &good));
    KP_TRY(kp_ws(ctx, (size_t)nl, &err2));
    {
        KP_PROFB(ctx, "gr_score", (double)kp_blocks(nl, GS_THREADS) * (double)m * 32.0 + (double)nl * 108.0);
        k_gr_score<<<kp_blocks(nl, GS_THREADS), GS_THREADS, 0, ctx->stream>>>(p, list, d_n, good, err2);
        KP_LAUNCH_CHECK(ctx);
    }
    std::vector<int32_t> h_list((size_t)nl), h_good((size_t)nl);
    std::vector<double> h_err((size_t)nl);
    KP_CUDA(ctx, cudaMemcpyAsync(h_list.data(), list, sizeof(int32_t) * (size_t)nl, cudaMemcpyDeviceToHost, ctx->stream));
    KP_CUDA(ctx, cudaMemcpyAsync(h_good.data(), good, sizeof(int32_t) * (size_t)nl, cudaMemcpyDeviceToHost, ctx->stream));
    KP_CUDA(ctx, cudaMemcpyAsync(h_err.data(), err2, sizeof(double) * (size_t)nl, cudaMemcpyDeviceToHost, ctx->stream));
    KP_TRY(kp_stream_wait(ctx));
    // upstream's sequential rule over the hypotheses in order: better = higher fitness, then lower rmse;
    // every improvement tightens the exit iteration ceil(log(1 - confidence) / log(1 - fitness^ransac_n))
    double best_fit = 0.0, best_rmse = 0.0;
    int best_w = -1;
    double exit_itr = (double)max_iteration;
    for (int w = 0; w < nl; ++w) {
        if ((double)h_list[(size_t)w] >= exit_itr) break;
        const int gd = h_good[(size_t)w];
        const double fit = gd > 0 ? (double)gd / (double)m : 0.0;
        const double rmse = gd > 0 ? sqrt(h_err[(size_t)w] / (double)gd) : 0.0;
        if (fit > best_fit || (fit == best_fit && rmse < best_rmse)) {   // RegistrationResult::IsBetterRANSACThan
            best_fit = fit; best_rmse = rmse; best_w = w;
            if (confidence < 1.0) {
                const double k_est = log(1.0 - confidence) / log(1.0 - pow(fit, (double)ransac_n));
                if (k_est < exit_itr) exit_itr = ceil(k_est);
            }
        }
    }
    if (best_w < 0) return KP_OK;
    double T12[12];
    KP_CUDA(ctx, cudaMemcpyAsync(ctx->d_scratch, p.T + (int64_t)h_list[(size_t)best_w] * 12, sizeof(double) * 12, cudaMemcpyDeviceToDevice, ctx->stream));
    KP_TRY(kp_fetch_scratch(ctx, sizeof(double) * 12));
    memcpy(T12, ctx->h_scratch, sizeof T12);
    for (int i = 0; i < 12; ++i) h_T16[i] = T12[i];
    h_T16[12] = h_T16[13] = h_T16[14] = 0.0; h_T16[15] = 1.0;
    if (h_fitness) *h_fitness = best_fit;
    if (h_rmse) *h_rmse = best_rmse;
    if (h_best_iter) *h_best_iter = h_list[(size_t)best_w];
    return KP_OK;
}

}  // extern "C"
