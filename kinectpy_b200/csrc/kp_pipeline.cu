// kp_pipeline.cu -- batched whole-frame driver (BASELINE config C4).
// Per frame: unproject + transform + fuse (preprocessing/data.py:44-58) -> filter_outliers
// (preprocessing/filtering.py:23-24: voxel + SOR) -> floor removal (floor_removal.py:64-73: band,
// segment_plane, invert-select, merge, SOR) -> point-to-plane ICP refinement of every sub
// extrinsic (preprocessing/registration.py:65-86: voxel, normals, registration_icp).
// Frames are independent (the reference loop at preprocessing/data.py:35 is a pure map), so a
// pipeline owns `n_streams` workers -- one host thread, one kp_ctx (= one CUDA stream and
// workspace) and one set of frame buffers each -- that pull frame indices from a shared counter.
// The per-stage counts that must reach the host (voxel count, kept counts, RANSAC tallies)
// cost a stream sync each; with several frames in flight another worker's kernels fill the gap.
#include <atomic>
#include <string.h>
#include <thread>
#include <vector>
#include "kp_common.cuh"

struct KpWorker {
    kp_ctx *ctx = nullptr;
    uint16_t *d_depth = nullptr;
    float *fused = nullptr, *raw = nullptr;
    float *A = nullptr, *B = nullptr, *C = nullptr, *D = nullptr, *E = nullptr, *Nrm = nullptr;
    uint8_t *mask = nullptr, *mask2 = nullptr;
    int32_t *pos = nullptr, *enc = nullptr;
    int rc = KP_OK;
};

struct kp_pipeline {
    kp_pipeline_cfg cfg;
    int device = 0;
    std::vector<KpWorker> workers;
    float *d_tab = nullptr;
    std::vector<double> T_fuse, T_icp;
    std::string err;
};

namespace {
#define PL_CUDA(p, call)                                                                         \
    do {                                                                                         \
        cudaError_t e__ = (call);                                                                \
        if (e__ != cudaSuccess) {                                                                \
            (p)->err = std::string(#call) + ": " + cudaGetErrorString(e__);                      \
            return KP_E_CUDA;                                                                    \
        }                                                                                        \
    } while (0)

int compact_rows(kp_ctx *ctx, KpWorker &w, int64_t n, const uint8_t *mask, int invert, const float *in, float *out)
{
    int32_t *d_total = (int32_t *)ctx->d_scratch + 128;     // clear of the K1 bounds rows fetched through the same scratch
    KP_TRY(kp_prim_compact_mask(ctx, n, mask, invert, in, w.pos, nullptr, d_total));
    return kp_prim_gather3(ctx, n, w.pos, in, out);
}

int run_frame(kp_pipeline *pl, KpWorker &w, const uint16_t *depth_f, bool on_dev, kp_frame_result *res, float *d_out,
              float *h_out = nullptr)
{
    const kp_pipeline_cfg &c = pl->cfg;
    kp_ctx *ctx = w.ctx;
    const int S = c.S;
    const int64_t P = c.P, NP = (int64_t)S * P;
    memset(res, 0, sizeof *res);
    kp_ws_reset(ctx);
    const uint16_t *d_depth = depth_f;
    if (!on_dev) {
        KP_CUDA(ctx, cudaMemcpyAsync(w.d_depth, depth_f, sizeof(uint16_t) * (size_t)NP, cudaMemcpyHostToDevice, ctx->stream));
        d_depth = w.d_depth;
    }
    // ---- K1: fused cloud in the master frame (+ raw sub clouds for ICP); bounds and counts per sensor ride along
    const bool icp = c.do_icp && S > 1;
    const int nrows = S + (icp ? S - 1 : 0);
    KP_TRY(kp_unproject_device(ctx, d_depth, pl->d_tab, pl->T_fuse.data(), 1, S, P, c.unproject_flags | KP_UP_BOUNDS_PER_SENSOR,
                               c.scale, w.fused, nullptr, nullptr, w.enc));
    if (icp)
        KP_TRY(kp_unproject_device(ctx, d_depth + P, pl->d_tab + 2 * P, nullptr, 1, S - 1, P,
                                   c.unproject_flags | KP_UP_BOUNDS_PER_SENSOR, c.scale, w.raw, nullptr, nullptr, w.enc + 8 * S));
    KP_CUDA(ctx, cudaMemcpyAsync(ctx->d_scratch, w.enc, (size_t)nrows * 8 * sizeof(int32_t), cudaMemcpyDeviceToDevice, ctx->stream));
    KP_TRY(kp_fetch_scratch(ctx, (size_t)nrows * 8 * sizeof(int32_t)));
    float sb6[2 * 6][6];          // per row: min xyz, max xyz
    int64_t scount[2 * 6];
    {
        const int32_t *h = (const int32_t *)ctx->h_scratch;
        for (int r = 0; r < nrows; ++r) {
            for (int k = 0; k < 6; ++k) {
                int32_t b = h[8 * r + k] >= 0 ? h[8 * r + k] : h[8 * r + k] ^ 0x7fffffff;
                memcpy(&sb6[r][k], &b, 4);
            }
            scount[r] = h[8 * r + 6];
        }
    }
    float b6[6];
    int64_t nvalid = 0;
    for (int k = 0; k < 3; ++k) { b6[k] = INFINITY; b6[3 + k] = -INFINITY; }
    for (int s = 0; s < S; ++s) {
        nvalid += scount[s];
        if (scount[s] <= 0) continue;
        for (int k = 0; k < 3; ++k) { b6[k] = fminf(b6[k], sb6[s][k]); b6[3 + k] = fmaxf(b6[3 + k], sb6[s][3 + k]); }
    }
    res->n_fused = nvalid;
    // ---- filter_outliers: voxel + SOR
    int64_t M = 0, kept = 0;
    KP_TRY(kp_voxel_device(ctx, w.fused, nullptr, nullptr, NP, c.voxel_size, b6, nvalid, w.A, nullptr, nullptr, nullptr,
                           nullptr, nullptr, &M));
    res->n_voxel = M;
    kp_ws_reset(ctx);
    float *cur = w.A;
    int64_t ncur = M;
    if (c.sor_k > 0 && M > 0) {
        KP_TRY(kp_sor_device(ctx, w.A, M, c.sor_k, c.sor_ratio, kp_knn_cell_from_voxel(c.voxel_size, c.sor_k), b6, w.mask,
                             nullptr, nullptr, &kept));
        KP_TRY(compact_rows(ctx, w, M, w.mask, 0, w.A, w.B));
        cur = w.B; ncur = kept;
        kp_ws_reset(ctx);
    }
    res->n_sor = ncur;
    // ---- floor removal
    if (c.do_floor && ncur >= c.ransac_n) {
        int64_t nlo = 0, ninl = 0;
        KP_TRY(kp_band_mask_device(ctx, cur, ncur, 1, c.floor_band, w.mask, nullptr, &nlo));
        int64_t nup = ncur - nlo;
        KP_TRY(compact_rows(ctx, w, ncur, w.mask, 0, cur, w.C));          // lower band
        KP_TRY(compact_rows(ctx, w, ncur, w.mask, 1, cur, w.D));          // upper part
        double plane[4];
        int32_t best = -1;
        if (nlo >= c.ransac_n) {
            KP_TRY(kp_ransac_device(ctx, w.C, nlo, c.ransac_thr, c.ransac_n, c.ransac_iters, 0.99999999, c.seed, plane,
                                    w.mask2, &ninl, &best, nullptr));
        } else {
            KP_CUDA(ctx, cudaMemsetAsync(w.mask2, 0, (size_t)(nlo > 0 ? nlo : 1), ctx->stream));
        }
        res->n_floor_inliers = ninl;
        // outlier_cloud + upper (floor_removal.py:71-72): band outliers first, then the upper part
        KP_TRY(compact_rows(ctx, w, nlo, w.mask2, 1, w.C, w.E));
        int64_t nrest = nlo - ninl;
        if (nup > 0)
            KP_CUDA(ctx, cudaMemcpyAsync(w.E + 3 * nrest, w.D, sizeof(float) * 3 * (size_t)nup, cudaMemcpyDeviceToDevice, ctx->stream));
        int64_t nmerged = nrest + nup;
        kp_ws_reset(ctx);
        cur = w.E; ncur = nmerged;
        if (c.floor_sor_k > 0 && nmerged > 0) {
            KP_TRY(kp_sor_device(ctx, w.E, nmerged, c.floor_sor_k, c.floor_sor_ratio,
                                 kp_knn_cell_from_voxel(c.voxel_size, c.floor_sor_k), b6, w.mask, nullptr, nullptr, &kept));
            KP_TRY(compact_rows(ctx, w, nmerged, w.mask, 0, w.E, w.C));
            cur = w.C; ncur = kept;
            kp_ws_reset(ctx);
        }
    }
    res->n_out = ncur;
    if (d_out && ncur > 0)
        KP_CUDA(ctx, cudaMemcpyAsync(d_out, cur, sizeof(float) * 3 * (size_t)ncur, cudaMemcpyDeviceToDevice, ctx->stream));
    if (h_out && ncur > 0)
        KP_CUDA(ctx, cudaMemcpyAsync(h_out, cur, sizeof(float) * 3 * (size_t)ncur, cudaMemcpyDeviceToHost, ctx->stream));
    // ---- ICP refinement of every sub extrinsic: target = master cloud, source = sub cloud in its own frame
    if (c.do_icp && S > 1) {
        const float *tb = sb6[0];         // the master cloud is sensor 0 of the fused one
        int64_t tn = scount[0], Mt = 0;
        KP_TRY(kp_voxel_device(ctx, w.fused, nullptr, nullptr, P, c.icp_voxel, tb, tn, w.A, nullptr, nullptr, nullptr, nullptr,
                               nullptr, &Mt));
        kp_ws_reset(ctx);
        KP_TRY(kp_normals_device(ctx, w.A, Mt, c.normals_radius, c.normals_max_nn, tb, w.Nrm));
        kp_ws_reset(ctx);
        KpGrid g;
        KP_TRY(kp_grid_build(ctx, w.A, Mt, c.icp_max_corr * (1.0 + 1e-6), tb, &g));
        for (int s = 1; s < S && s <= 5; ++s) {
            const float *raw_s = w.raw + 3 * (size_t)(s - 1) * P;
            const float *sb = sb6[S + s - 1];
            int64_t sn = scount[S + s - 1], Ms = 0;
            KP_TRY(kp_voxel_device(ctx, raw_s, nullptr, nullptr, P, c.icp_voxel, sb, sn, w.B, nullptr, nullptr, nullptr,
                                   nullptr, nullptr, &Ms));
            int64_t ncorr = 0;
            KP_TRY(kp_icp_device(ctx, w.B, Ms, g, w.Nrm, c.icp_max_corr, pl->T_icp.data() + 16 * s, c.icp_max_iter, 1e-6, 1e-6,
                                 res->icp_T[s - 1], &res->icp_fitness[s - 1], &res->icp_rmse[s - 1], &res->icp_iters[s - 1],
                                 &ncorr));
        }
        kp_ws_reset(ctx);
    }
    return kp_stream_wait(ctx);
}
}  // namespace

extern "C" {

int kp_pipeline_create(int device, const kp_pipeline_cfg *cfg, const float *h_xytab, const double *h_T,
                       kp_pipeline **out)
{
    if (!cfg || !h_xytab || !out) return kp_set_err(nullptr, KP_E_ARG, "kp_pipeline_create: NULL argument");
    if (cfg->S < 1 || cfg->S > 6 || cfg->P < 1) return kp_set_err(nullptr, KP_E_ARG, "kp_pipeline_create: need 1..6 sensors and P >= 1");
    *out = nullptr;
    kp_pipeline *p = new kp_pipeline();
    p->cfg = *cfg;
    p->device = device;
    int nw = cfg->n_streams < 1 ? 1 : (cfg->n_streams > 16 ? 16 : cfg->n_streams);
    {
        char hint[16];
        snprintf(hint, sizeof hint, "%d", nw);
        setenv("KP_WORKERS_HINT", hint, 0);     // lets kp_stream_wait pick spin or sleep for this box (kp_ctx.cu)
    }
    p->workers.resize(nw);
    const int S = cfg->S;
    const size_t NP = (size_t)S * cfg->P;
    p->T_fuse.assign(16 * S, 0.0);
    p->T_icp.assign(16 * S, 0.0);
    for (int s = 0; s < S; ++s)
        for (int i = 0; i < 16; ++i) {
            // h_T: [2][S][16] = fusion extrinsics, then ICP initial guesses; NULL -> identity for both
            p->T_fuse[16 * s + i] = h_T ? h_T[16 * s + i] : (i % 5 == 0 ? 1.0 : 0.0);
            p->T_icp[16 * s + i] = h_T ? h_T[16 * (S + s) + i] : (i % 5 == 0 ? 1.0 : 0.0);
        }
    for (auto &w : p->workers) {
        int rc = kp_ctx_create(device, &w.ctx);
        if (rc != KP_OK) {
            std::string m = kp_last_error(nullptr);
            kp_pipeline_destroy(p);
            return kp_set_err(nullptr, rc, "%s", m.c_str());
        }
    }
    cudaSetDevice(device);
    auto fail = [&](const char *what) {
        std::string m = std::string("kp_pipeline_create: ") + what + ": " + cudaGetErrorString(cudaGetLastError());
        kp_pipeline_destroy(p);
        return kp_set_err(nullptr, KP_E_NOMEM, "%s", m.c_str());
    };
    if (cudaMalloc((void **)&p->d_tab, sizeof(float) * 2 * NP) != cudaSuccess) return fail("table");
    if (cudaMemcpy(p->d_tab, h_xytab, sizeof(float) * 2 * NP, cudaMemcpyHostToDevice) != cudaSuccess) return fail("table copy");
    for (auto &w : p->workers) {
        float **bufs[] = {&w.fused, &w.raw, &w.A, &w.B, &w.C, &w.D, &w.E, &w.Nrm};
        for (float **b : bufs)
            if (cudaMalloc((void **)b, sizeof(float) * 3 * NP) != cudaSuccess) return fail("frame buffers");
        if (cudaMalloc((void **)&w.d_depth, sizeof(uint16_t) * NP) != cudaSuccess) return fail("depth staging");
        if (cudaMalloc((void **)&w.mask, NP) != cudaSuccess) return fail("mask");
        if (cudaMalloc((void **)&w.mask2, NP) != cudaSuccess) return fail("mask");
        if (cudaMalloc((void **)&w.pos, sizeof(int32_t) * NP) != cudaSuccess) return fail("pos");
        if (cudaMalloc((void **)&w.enc, sizeof(int32_t) * 8 * 16) != cudaSuccess) return fail("enc");
    }
    *out = p;
    return KP_OK;
}

int kp_pipeline_destroy(kp_pipeline *p)
{
    if (!p) return KP_OK;
    cudaSetDevice(p->device);
    for (auto &w : p->workers) {
        if (w.ctx) cudaStreamSynchronize(w.ctx->stream);
        float *bufs[] = {w.fused, w.raw, w.A, w.B, w.C, w.D, w.E, w.Nrm};
        for (float *b : bufs) if (b) cudaFree(b);
        if (w.d_depth) cudaFree(w.d_depth);
        if (w.mask) cudaFree(w.mask);
        if (w.mask2) cudaFree(w.mask2);
        if (w.pos) cudaFree(w.pos);
        if (w.enc) cudaFree(w.enc);
        if (w.ctx) kp_ctx_destroy(w.ctx);
    }
    if (p->d_tab) cudaFree(p->d_tab);
    delete p;
    return KP_OK;
}

const char *kp_pipeline_last_error(kp_pipeline *p) { return p ? p->err.c_str() : kp_last_error(nullptr); }

static int pipeline_run_impl(kp_pipeline *p, const uint16_t *depth, int depth_on_device, int64_t F,
                             kp_frame_result *h_results, float *d_out_xyz, float *h_out_xyz, int64_t out_stride);

int kp_pipeline_run(kp_pipeline *p, const uint16_t *depth, int depth_on_device, int64_t F, kp_frame_result *h_results,
                    float *d_out_xyz, int64_t out_stride)
{
    return pipeline_run_impl(p, depth, depth_on_device, F, h_results, d_out_xyz, nullptr, out_stride);
}

int kp_pipeline_run_host(kp_pipeline *p, const uint16_t *h_depth, int64_t F, kp_frame_result *h_results, float *h_out_xyz,
                         int64_t out_stride)
{
    return pipeline_run_impl(p, h_depth, 0, F, h_results, nullptr, h_out_xyz, out_stride);
}

static int pipeline_run_impl(kp_pipeline *p, const uint16_t *depth, int depth_on_device, int64_t F,
                             kp_frame_result *h_results, float *d_out_xyz, float *h_out_xyz, int64_t out_stride)
{
    if (!p || !depth || !h_results || F < 0) return kp_set_err(nullptr, KP_E_ARG, "kp_pipeline_run: bad argument");
    const int64_t NP = (int64_t)p->cfg.S * p->cfg.P;
    std::atomic<int64_t> next(0);
    auto body = [&](KpWorker *w) {
        cudaSetDevice(p->device);
        w->rc = KP_OK;
        for (;;) {
            int64_t f = next.fetch_add(1);
            if (f >= F) break;
            float *out = d_out_xyz ? d_out_xyz + 3 * f * out_stride : nullptr;
            float *hout = h_out_xyz ? h_out_xyz + 3 * f * out_stride : nullptr;
            int rc = run_frame(p, *w, depth + f * NP, depth_on_device != 0, &h_results[f], out, hout);
            h_results[f].status = rc;
            if (rc != KP_OK) { w->rc = rc; break; }
        }
    };
    if (p->workers.size() == 1 || F <= 1) {
        body(&p->workers[0]);
    } else {
        std::vector<std::thread> th;
        for (auto &w : p->workers) th.emplace_back(body, &w);
        for (auto &t : th) t.join();
    }
    for (auto &w : p->workers)
        if (w.rc != KP_OK) { p->err = w.ctx->err; return w.rc; }
    return KP_OK;
}

int64_t kp_pipeline_launch_count(kp_pipeline *p)
{
    int64_t n = 0;
    if (p) for (auto &w : p->workers) n += w.ctx->launches;
    return n;
}

int kp_pipeline_profile(kp_pipeline *p, int enable_or_read, int max_entries, const char **h_names, double *h_ms,
                        int64_t *h_calls, double *h_bytes, int *h_n)
{
    // enable_or_read: 1 = enable + reset, 0 = disable, 2 = read (summed over the workers)
    if (!p) return KP_E_ARG;
    cudaSetDevice(p->device);
    if (enable_or_read == 1 || enable_or_read == 0) {
        for (auto &w : p->workers) { kp_profile_reset(w.ctx); kp_profile_enable(w.ctx, enable_or_read); }
        return KP_OK;
    }
    std::vector<const char *> names;
    std::vector<double> ms, bytes;
    std::vector<int64_t> calls;
    for (auto &w : p->workers) {
        const char *nm[64]; double m[64], by[64]; int64_t cl[64]; int n = 0;
        int rc = kp_profile_read(w.ctx, 64, nm, m, cl, by, &n);
        if (rc != KP_OK) return rc;
        for (int i = 0; i < n; ++i) {
            size_t j = 0;
            for (; j < names.size(); ++j) if (strcmp(names[j], nm[i]) == 0) break;
            if (j == names.size()) { names.push_back(nm[i]); ms.push_back(0); calls.push_back(0); bytes.push_back(0); }
            ms[j] += m[i]; calls[j] += cl[i]; bytes[j] += by[i];
        }
    }
    int n = 0;
    for (size_t j = 0; j < names.size() && n < max_entries; ++j, ++n) {
        h_names[n] = names[j]; h_ms[n] = ms[j]; h_calls[n] = calls[j];
        if (h_bytes) h_bytes[n] = bytes[j];
    }
    *h_n = n;
    return KP_OK;
}

}  // extern "C"
