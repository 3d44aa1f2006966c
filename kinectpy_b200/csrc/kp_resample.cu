// kp_resample.cu -- K6: fixed-N point resampling feeding PointNet (BASELINE config C5).
// Replaces select_points_randomly (utils/processing.py:259-275: np.random.choice(n, N, replace=False),
// called at datasets/kinect_dataset.py:103-104) and the prefix variant `points[:N]` of
// datasets/kinect_dataset_npz.py:96-97.
//
// np.random.choice draws from numpy's global Mersenne state, which no parallel machine reproduces;
// the subset here is defined by a counter-based key instead: point i of sample stream s gets
// key = min(kp_rng(seed, s, i) >> 32, 2^32 - 2), NaN rows get 2^32 - 1, and the sample is the N
// points with the smallest (key, index), in that order -- a uniformly random N-subset in uniformly
// random order, like the reference's, and one a CPU restatement can replay bit for bit.  The selection is a
// stable LSD radix sort of the 32-bit keys (HBM-bound streaming) followed by a gather of the first N.
#include "kp_common.cuh"

namespace {
__global__ void __launch_bounds__(256) k_resample_keys(const float *xyz, int64_t n, uint64_t seed, uint64_t stream,
                                                       uint32_t *keys, int32_t *vals)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t key = 0xffffffffu;
    if (!isnan(xyz[3 * i])) {
        key = (uint32_t)(kp_rng(seed, stream, (uint64_t)i) >> 32);
        if (key == 0xffffffffu) key = 0xfffffffeu;
    }
    keys[i] = key;
    vals[i] = (int32_t)i;
}
__global__ void __launch_bounds__(256) k_resample_gather(const float *xyz, const int32_t *order, int64_t N, float *out,
                                                         int32_t *index_out)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N) return;
    const int64_t i = order[t];
    out[3 * t] = xyz[3 * i]; out[3 * t + 1] = xyz[3 * i + 1]; out[3 * t + 2] = xyz[3 * i + 2];
    if (index_out) index_out[t] = (int32_t)i;
}

// one cloud; asynchronous except for the one 4-byte read that validates N <= #valid
int resample_one(kp_ctx *ctx, const float *d_xyz, int64_t n, int64_t N, int mode, uint64_t seed, uint64_t stream,
                 float *d_out, int32_t *d_index_out, int64_t *h_count)
{
    if (mode == KP_RESAMPLE_PREFIX) {
        const int64_t m = n < N ? n : N;
        if (m > 0) KP_CUDA(ctx, cudaMemcpyAsync(d_out, d_xyz, sizeof(float) * 3 * (size_t)m, cudaMemcpyDeviceToDevice, ctx->stream));
        if (h_count) *h_count = m;
        return KP_OK;
    }
    if (N > n) return kp_set_err(ctx, KP_E_ARG, "select_points_randomly: cannot take a larger sample (%lld) than population (%lld) when replace=False",
                                 (long long)N, (long long)n);
    if (h_count) *h_count = N;
    if (N == 0) return KP_OK;
    uint32_t *keys, *keys_tmp, *keys_sorted;
    int32_t *vals, *vals_tmp, *vals_sorted;
    KP_TRY(kp_ws(ctx, (size_t)n, &keys));
    KP_TRY(kp_ws(ctx, (size_t)n, &keys_tmp));
    KP_TRY(kp_ws(ctx, (size_t)n, &vals));
    KP_TRY(kp_ws(ctx, (size_t)n, &vals_tmp));
    {
        KP_PROFB(ctx, "resample_keys", (double)n * (4.0 + 8.0));
        k_resample_keys<<<kp_blocks(n, 256), 256, 0, ctx->stream>>>(d_xyz, n, seed, stream, keys, vals);
        KP_LAUNCH_CHECK(ctx);
    }
    KP_TRY(kp_prim_sort_pairs_u32(ctx, n, 32, keys, keys_tmp, vals, vals_tmp, &keys_sorted, &vals_sorted));
    {
        KP_PROFB(ctx, "resample_gather", (double)N * (4.0 + 12.0 + 12.0));
        k_resample_gather<<<kp_blocks(N, 256), 256, 0, ctx->stream>>>(d_xyz, vals_sorted, N, d_out, d_index_out);
        KP_LAUNCH_CHECK(ctx);
    }
    // the N-th key is the NaN sentinel <=> fewer than N valid points
    KP_CUDA(ctx, cudaMemcpyAsync(ctx->d_scratch, keys_sorted + (N - 1), sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
    KP_TRY(kp_fetch_scratch(ctx, sizeof(uint32_t)));
    if (*(uint32_t *)ctx->h_scratch == 0xffffffffu)
        return kp_set_err(ctx, KP_E_ARG, "select_points_randomly: fewer than %lld valid (non-NaN) points", (long long)N);
    return KP_OK;
}
}  // namespace

extern "C" {

int kp_resample_fixed_n(kp_ctx *ctx, const float *d_xyz, int64_t n, int64_t N, int mode, uint64_t seed, uint64_t stream,
                        float *d_out, int32_t *d_index_out, int64_t *h_count)
{
    if (!ctx || (n > 0 && !d_xyz) || (N > 0 && !d_out)) return kp_set_err(ctx, KP_E_ARG, "kp_resample_fixed_n: NULL argument");
    if (N < 0 || n < 0 || n > 2147483000LL) return kp_set_err(ctx, KP_E_ARG, "kp_resample_fixed_n: bad size");
    if (mode != KP_RESAMPLE_RANDOM && mode != KP_RESAMPLE_PREFIX) return kp_set_err(ctx, KP_E_ARG, "kp_resample_fixed_n: bad mode");
    kp_enter(ctx);
    return resample_one(ctx, d_xyz, n, N, mode, seed, stream, d_out, d_index_out, h_count);
}

int kp_resample_batch(kp_ctx *ctx, const float *d_xyz, const int64_t *h_offsets, int B, int64_t N, int mode, uint64_t seed,
                      uint64_t first_stream, float *d_out, int64_t *h_counts)
{
    if (!ctx || !h_offsets || B < 0 || (B > 0 && N > 0 && !d_out)) return kp_set_err(ctx, KP_E_ARG, "kp_resample_batch: NULL argument");
    if (mode != KP_RESAMPLE_RANDOM && mode != KP_RESAMPLE_PREFIX) return kp_set_err(ctx, KP_E_ARG, "kp_resample_batch: bad mode");
    kp_enter(ctx);
    for (int b = 0; b < B; ++b) {
        const int64_t n = h_offsets[b + 1] - h_offsets[b];
        if (n < 0 || n > 2147483000LL) return kp_set_err(ctx, KP_E_ARG, "kp_resample_batch: offsets must be non-decreasing");
        kp_ws_reset(ctx);
        int64_t cnt = 0;
        KP_TRY(resample_one(ctx, d_xyz + 3 * h_offsets[b], n, N, mode, seed, first_stream + (uint64_t)b,
                            d_out + 3 * (size_t)b * (size_t)N, nullptr, &cnt));
        // prefix mode on a short cloud: the tail of the [N][3] slot is zero-filled (a dense batch tensor)
        if (cnt < N)
            KP_CUDA(ctx, cudaMemsetAsync(d_out + 3 * ((size_t)b * (size_t)N + (size_t)cnt), 0, sizeof(float) * 3 * (size_t)(N - cnt), ctx->stream));
        if (h_counts) h_counts[b] = cnt;
    }
    return KP_OK;
}

}  // extern "C"
