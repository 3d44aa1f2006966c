// kp_primitives.cu -- device-wide building blocks written for this library:
// single-pass ordered stream compaction (ballot scan + decoupled look-back), stable LSD radix sort of (key, index) pairs,
// run-head extraction, bounds, canonical double sums.  No CUB/Thrust on the product path.
#include <math.h>
#include "kp_common.cuh"
#include "kp_batch.cuh"

// ===================================================== ordered compaction ==
// Tile = 256 threads x 8 items, striped: item j of thread t is element base + j*256 + t, so the
// element order inside a tile is (j, warp, lane) and every load is coalesced.  Flags are scanned
// with ballots: rank inside the warp from popc, then a 64-entry (j, warp) table per tile.
namespace {
constexpr int SC_THREADS = 256;
constexpr int SC_ITEMS = 8;
constexpr int SC_TILE = SC_THREADS * SC_ITEMS;

struct FlagMask {
    const uint8_t *mask; int invert; const float *nan_src;
    __device__ bool operator()(int64_t i) const
    {
        if (mask) return (mask[i] != 0) != (invert != 0);
        return !isnan(nan_src[3 * i]);
    }
};
template <class K>
struct FlagRunHead {
    const K *keys;
    __device__ bool operator()(int64_t i) const { return i == 0 || keys[i] != keys[i - 1]; }
};
struct EmitPos {
    int32_t *pos_out; int32_t *index_out;
    __device__ void operator()(int64_t i, bool f, int32_t pos) const
    {
        if (pos_out) pos_out[i] = f ? pos : -1;
        if (f && index_out) index_out[pos] = (int32_t)i;
    }
};
struct EmitRunStart {
    int32_t *run_start;
    __device__ void operator()(int64_t i, bool f, int32_t pos) const
    {
        if (f) run_start[pos] = (int32_t)i;
    }
};

template <class Flag>
__global__ void __launch_bounds__(SC_THREADS) k_flag_count(int64_t n, Flag flag, int32_t *block_sums)
{
    __shared__ int warp_cnt[SC_THREADS / 32];
    int64_t base = (int64_t)blockIdx.x * SC_TILE;
    int cnt = 0;
#pragma unroll
    for (int j = 0; j < SC_ITEMS; ++j) {
        int64_t i = base + j * SC_THREADS + threadIdx.x;
        bool f = i < n && flag(i);
        cnt += __popc(__ballot_sync(KP_FULL, f));   // identical in all lanes of the warp
    }
    if ((threadIdx.x & 31) == 0) warp_cnt[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int w = 0; w < SC_THREADS / 32; ++w) s += warp_cnt[w];
        block_sums[blockIdx.x] = s;
    }
}

// single block: in-place exclusive scan of `nb` int32 values, total to *total
__global__ void __launch_bounds__(1024) k_scan_block_sums(int32_t *v, int nb, int32_t *total)
{
    __shared__ int warp_tot[32];
    __shared__ int carry_s;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += 1024) {
        int i = base + threadIdx.x;
        int x = i < nb ? v[i] : 0;
        int incl = x;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            int y = __shfl_up_sync(KP_FULL, incl, s);
            if (lane >= s) incl += y;
        }
        if (lane == 31) warp_tot[w] = incl;
        __syncthreads();
        if (w == 0) {
            int t = warp_tot[lane];
            int ti = t;
#pragma unroll
            for (int s = 1; s < 32; s <<= 1) {
                int y = __shfl_up_sync(KP_FULL, ti, s);
                if (lane >= s) ti += y;
            }
            warp_tot[lane] = ti - t;   // exclusive warp offsets
        }
        __syncthreads();
        int excl = carry_s + warp_tot[w] + incl - x;
        if (i < nb) v[i] = excl;
        __syncthreads();               // everyone has read carry_s and warp_tot
        if (threadIdx.x == 1023) carry_s = excl + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry_s;
}

template <class Flag, class Emit>
__global__ void __launch_bounds__(SC_THREADS) k_flag_scatter(int64_t n, Flag flag, Emit emit, const int32_t *block_off)
{
    __shared__ int cnt[SC_ITEMS * (SC_THREADS / 32) + 1];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int64_t base = (int64_t)blockIdx.x * SC_TILE;
    bool f[SC_ITEMS];
    int rank[SC_ITEMS];
#pragma unroll
    for (int j = 0; j < SC_ITEMS; ++j) {
        int64_t i = base + j * SC_THREADS + threadIdx.x;
        f[j] = i < n && flag(i);
        unsigned b = __ballot_sync(KP_FULL, f[j]);
        rank[j] = __popc(b & ((1u << lane) - 1u));
        if (lane == 0) cnt[j * (SC_THREADS / 32) + w] = __popc(b);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = block_off[blockIdx.x];
        for (int e = 0; e < SC_ITEMS * (SC_THREADS / 32); ++e) { int c = cnt[e]; cnt[e] = s; s += c; }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < SC_ITEMS; ++j) {
        int64_t i = base + j * SC_THREADS + threadIdx.x;
        if (i < n) emit(i, f[j], cnt[j * (SC_THREADS / 32) + w] + rank[j]);
    }
}

// ---- single-pass scan: tile offsets in two levels.  Tile t publishes its aggregate; the last tile of every group
// of 32 adds its group's aggregates and publishes the group total; tile t's exclusive prefix is then (totals of
// the groups before it) + (aggregates of the tiles before it in its own group): two dependent reads, whatever t
// is.  (A frame's arrays are a few hundred tiles that are all resident and in step.  A classic decoupled
// look-back, where a tile walks back until it meets a finished predecessor, was measured first: inclusive prefixes
// appear late in that regime, tiles near the end made ~17 dependent round trips, and 62 % of the kernel was the
// other warps waiting at the barrier behind that walk.)  A word carries {epoch:30, flag:2, value:32}; words of
// earlier calls have another epoch and read as "not ready", so the state arrays are never cleared.  Tiles are
// numbered by blockIdx.x (see k_flag_compact), so every tile a running tile waits for has started.
constexpr unsigned LB_READY = 1u;
__device__ __forceinline__ unsigned long long lb_pack(unsigned epoch, unsigned flag, unsigned v)
{
    return ((unsigned long long)((epoch << 2) | flag) << 32) | v;
}
__device__ __forceinline__ unsigned lb_wait(volatile unsigned long long *word, unsigned epoch)
{
    unsigned long long x;
    do { x = *word; } while ((unsigned)(x >> 34) != epoch);
    return (unsigned)x;
}
// executed by one full warp; returns the exclusive prefix of `tile` to every lane.
// state: [LB_TILES] tile aggregates, then [LB_TILES / 32] group totals.
__device__ __forceinline__ unsigned lb_exclusive(volatile unsigned long long *state, unsigned tile, unsigned agg,
                                                 unsigned epoch, int lane)
{
    volatile unsigned long long *gstate = state + kp_ctx::LB_TILES;
    if (lane == 0) state[tile] = lb_pack(epoch, LB_READY, agg);
    const unsigned g = tile >> 5, r = tile & 31u;
    // the tiles before me in my group first: a group total must not wait for the totals of the groups before it
    unsigned in_group = (unsigned)lane < r ? lb_wait(state + (g << 5) + lane, epoch) : 0u;
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) in_group += __shfl_xor_sync(KP_FULL, in_group, s);
    if (r == 31u && lane == 0) gstate[g] = lb_pack(epoch, LB_READY, in_group + agg);
    // then the totals of the groups before mine
    unsigned v = 0;
    for (unsigned base = 0; base < g; base += 32)
        if (base + lane < g) v += lb_wait(gstate + base + lane, epoch);
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) v += __shfl_xor_sync(KP_FULL, v, s);
    return v + in_group;
}

template <class Flag, class Emit>
__global__ void __launch_bounds__(SC_THREADS) k_flag_compact(int64_t n, Flag flag, Emit emit, unsigned long long *state,
                                                             unsigned epoch, int32_t *total)
{
    __shared__ int cnt[SC_ITEMS * (SC_THREADS / 32)];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    // Tiles are numbered by blockIdx.x: CTAs of a 1-D grid are dispatched in index order, so every predecessor
    // of a running tile has started.  (Numbering them with an atomic ticket was measured: a few hundred
    // same-address atomics serialise at ~5 ns each, a third of this kernel's run time.)
    const unsigned tile = blockIdx.x;
    const int64_t base = (int64_t)tile * SC_TILE;
    bool f[SC_ITEMS];
    int rank[SC_ITEMS];
#pragma unroll
    for (int j = 0; j < SC_ITEMS; ++j) {
        int64_t i = base + j * SC_THREADS + threadIdx.x;
        f[j] = i < n && flag(i);
        unsigned b = __ballot_sync(KP_FULL, f[j]);
        rank[j] = __popc(b & ((1u << lane) - 1u));
        if (lane == 0) cnt[j * (SC_THREADS / 32) + w] = __popc(b);
    }
    __syncthreads();
    if (w == 0) {
        // element order inside the tile is (j, warp, lane) = entry order of cnt[]; a lane owns two entries
        const int c0 = cnt[2 * lane], c1 = cnt[2 * lane + 1];
        const int s = c0 + c1;
        int incl = s;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int y = __shfl_up_sync(KP_FULL, incl, d);
            if (lane >= d) incl += y;
        }
        const unsigned agg = (unsigned)__shfl_sync(KP_FULL, incl, 31);
        const unsigned excl = lb_exclusive(state, tile, agg, epoch, lane);
        cnt[2 * lane] = (int)excl + incl - s;
        cnt[2 * lane + 1] = (int)excl + incl - s + c0;
        if (tile == gridDim.x - 1 && lane == 0) *total = (int32_t)(excl + agg);
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < SC_ITEMS; ++j) {
        int64_t i = base + j * SC_THREADS + threadIdx.x;
        if (i < n) emit(i, f[j], cnt[j * (SC_THREADS / 32) + w] + rank[j]);
    }
}

// next look-back epoch of the context (0 is never valid: the state starts zeroed)
int lb_next_epoch(kp_ctx *ctx, unsigned *epoch)
{
    if (++ctx->lb_epoch >= (1u << 30)) {
        KP_CUDA(ctx, cudaMemsetAsync(ctx->d_lb_state, 0, sizeof(unsigned long long) * (kp_ctx::LB_TILES + kp_ctx::LB_TILES / 32), ctx->stream));
        ctx->lb_epoch = 1;
    }
    *epoch = ctx->lb_epoch;
    return KP_OK;
}

template <class Flag, class Emit>
int compact_generic(kp_ctx *ctx, int64_t n, Flag flag, Emit emit, int32_t *d_total)
{
    if (n <= 0) {
        KP_CUDA(ctx, cudaMemsetAsync(d_total, 0, sizeof(int32_t), ctx->stream));
        return KP_OK;
    }
    unsigned nb = kp_blocks(n, SC_TILE);
    // The single-pass kernel numbers its tiles by blockIdx.x and a tile waits for the words of the tiles before it: that
    // only makes progress if every earlier tile has started, which is guaranteed while the whole grid is resident (four
    // CTAs of 256 threads per SM fit in every configuration of this kernel) and merely customary beyond that.  Larger
    // arrays take the count / scan / scatter path below, where no CTA waits for another one.
    if (nb <= (unsigned)kp_ctx::LB_TILES && nb <= (unsigned)ctx->sm_count * 4u) {
        unsigned epoch;
        KP_TRY(lb_next_epoch(ctx, &epoch));
        k_flag_compact<<<nb, SC_THREADS, 0, ctx->stream>>>(n, flag, emit, ctx->d_lb_state, epoch, d_total);
        KP_LAUNCH_CHECK(ctx);
        return KP_OK;
    }
    // more tiles than are resident at once: count / scan / scatter
    int32_t *d_sums;
    KP_TRY(kp_ws(ctx, nb, &d_sums));
    k_flag_count<<<nb, SC_THREADS, 0, ctx->stream>>>(n, flag, d_sums);
    KP_LAUNCH_CHECK(ctx);
    k_scan_block_sums<<<1, 1024, 0, ctx->stream>>>(d_sums, (int)nb, d_total);
    KP_LAUNCH_CHECK(ctx);
    k_flag_scatter<<<nb, SC_THREADS, 0, ctx->stream>>>(n, flag, emit, d_sums);
    KP_LAUNCH_CHECK(ctx);
    return KP_OK;
}

__global__ void k_gather3(int64_t n, const int32_t *pos, const float *in, float *out)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int p = pos[i];
    if (p < 0) return;
    float a = in[3 * i], b = in[3 * i + 1], c = in[3 * i + 2];
    out[3 * (int64_t)p] = a; out[3 * (int64_t)p + 1] = b; out[3 * (int64_t)p + 2] = c;
}
}  // namespace

int kp_prim_compact_mask(kp_ctx *ctx, int64_t n, const uint8_t *d_mask, int invert, const float *d_nan_src,
                         int32_t *d_pos_out, int32_t *d_index_out, int32_t *d_total)
{
    KP_PROFB(ctx, "compact_scan", (double)n * (d_mask ? 2.0 + 4.0 : 2.0 * 12.0 + 4.0));
    FlagMask fl{d_mask, invert, d_nan_src};
    EmitPos em{d_pos_out, d_index_out};
    return compact_generic(ctx, n, fl, em, d_total);
}

int kp_prim_gather3(kp_ctx *ctx, int64_t n, const int32_t *d_pos, const float *d_in, float *d_out)
{
    if (n <= 0) return KP_OK;
    KP_PROFB(ctx, "compact_gather", (double)n * (4.0 + 12.0 + 12.0));
    k_gather3<<<kp_blocks(n, 256), 256, 0, ctx->stream>>>(n, d_pos, d_in, d_out);
    KP_LAUNCH_CHECK(ctx);
    return KP_OK;
}

int kp_prim_run_starts_u64(kp_ctx *ctx, int64_t n, const uint64_t *d_keys, int32_t *d_run_start, int32_t *d_total)
{
    KP_PROFB(ctx, "run_heads", (double)n * 2.0 * 8.0);
    return compact_generic(ctx, n, FlagRunHead<uint64_t>{d_keys}, EmitRunStart{d_run_start}, d_total);
}
int kp_prim_run_starts_u32(kp_ctx *ctx, int64_t n, const uint32_t *d_keys, int32_t *d_run_start, int32_t *d_total)
{
    KP_PROFB(ctx, "run_heads", (double)n * 2.0 * 4.0);
    return compact_generic(ctx, n, FlagRunHead<uint32_t>{d_keys}, EmitRunStart{d_run_start}, d_total);
}

// ============================================================ radix sort ==
// Stable LSD sort, 8 bits per pass, three kernels per pass:
//   hist   : per-tile digit histogram           -> g_hist[digit][tile]
//   scan   : per-digit exclusive scan over tiles (one block per digit) -> g_hist in place + g_tot[digit]
//   scatter: per-warp stable ranking with __match_any_sync, then direct scatter
// Tile = 256 threads x RS_ITEMS keys; warp w owns a contiguous slab of RS_ITEMS*32 keys so the
// order inside a tile is (warp, item, lane) = input order.
namespace {
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 8;    // 2048-key tiles: enough CTAs to fill 148 SMs already at ~1 M keys
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;

template <class K>
__global__ void __launch_bounds__(RS_THREADS) k_rs_hist(const K *keys, int64_t n, int shift, int32_t *g_hist, int nb,
                                                        int64_t seg_stride)
{
    // blockIdx.y = segment of a batched sort (kp_b_sort_pairs_u32): every array advances by its per-segment stride
    keys += (int64_t)blockIdx.y * seg_stride;
    g_hist += (int64_t)blockIdx.y * 256 * nb;
    __shared__ int hist[256];
    hist[threadIdx.x] = 0;
    __syncthreads();
    int64_t base = (int64_t)blockIdx.x * RS_TILE;
    if (base + RS_TILE <= n) {
        // full tile: a thread takes RS_ITEMS CONSECUTIVE keys (16-byte loads) and sends one shared-memory atomic per
        // run of equal digits.  Keys arrive spatially ordered, so above the lowest digit a thread's keys mostly
        // share their digit: the kernel is bound by these atomics, and this halves them.
        constexpr int PER16 = 16 / (int)sizeof(K), NV = RS_ITEMS / PER16;
        const uint4 *src = reinterpret_cast<const uint4 *>(keys + base) + (int64_t)threadIdx.x * NV;
        uint4 raw[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) raw[v] = src[v];
        const K *kk = reinterpret_cast<const K *>(raw);
        unsigned run_d = (unsigned)(kk[0] >> shift) & 0xFFu;
        int run_n = 1;
#pragma unroll
        for (int j = 1; j < RS_ITEMS; ++j) {
            const unsigned d = (unsigned)(kk[j] >> shift) & 0xFFu;
            if (d == run_d) ++run_n;
            else { atomicAdd(&hist[run_d], run_n); run_d = d; run_n = 1; }
        }
        atomicAdd(&hist[run_d], run_n);
    } else {
#pragma unroll 4
        for (int j = 0; j < RS_ITEMS; ++j) {
            int64_t i = base + j * RS_THREADS + threadIdx.x;
            if (i < n) atomicAdd(&hist[(unsigned)(keys[i] >> shift) & 0xFFu], 1);
        }
    }
    __syncthreads();
    g_hist[(int64_t)threadIdx.x * nb + blockIdx.x] = hist[threadIdx.x];
}

// one block per digit: exclusive scan of that digit's per-tile counts (row of g_hist), total to g_tot[digit]
__global__ void __launch_bounds__(256) k_rs_scan_rows(int32_t *g_hist, int nb, int32_t *g_tot)
{
    g_hist += (int64_t)blockIdx.y * 256 * nb;
    g_tot += (int64_t)blockIdx.y * 256;
    __shared__ int warp_tot[8];
    const int d = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    int32_t *row = g_hist + (int64_t)d * nb;
    const int seg = (nb + 255) / 256;
    const int lo = min(tid * seg, nb), hi = min(lo + seg, nb);
    int sum = 0;
    for (int i = lo; i < hi; ++i) sum += row[i];
    int incl = sum;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
        int y = __shfl_up_sync(KP_FULL, incl, s);
        if (lane >= s) incl += y;
    }
    if (lane == 31) warp_tot[w] = incl;
    __syncthreads();
    int woff = 0;
#pragma unroll
    for (int ww = 0; ww < 8; ++ww) woff += ww < w ? warp_tot[ww] : 0;
    int run = woff + incl - sum;
    for (int i = lo; i < hi; ++i) { int c = row[i]; row[i] = run; run += c; }
    if (tid == 255) g_tot[d] = run;
}

template <class K>
__global__ void __launch_bounds__(RS_THREADS) k_rs_scatter(const K *keys_in, const int32_t *vals_in, K *keys_out,
                                                           int32_t *vals_out, int64_t n, int shift,
                                                           const int32_t *g_hist, const int32_t *g_tot, int nb,
                                                           int64_t seg_stride)
{
    keys_in += (int64_t)blockIdx.y * seg_stride; keys_out += (int64_t)blockIdx.y * seg_stride;
    if (vals_in) vals_in += (int64_t)blockIdx.y * seg_stride;
    vals_out += (int64_t)blockIdx.y * seg_stride;
    g_hist += (int64_t)blockIdx.y * 256 * nb; g_tot += (int64_t)blockIdx.y * 256;
    __shared__ int cnt[RS_WARPS][256];
    __shared__ int dig_wtot[RS_WARPS];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    // digit threadIdx.x: its total and this tile's offset inside it are requested now, used after the ranking
    const int dig_tot = g_tot[threadIdx.x];
    const int dig_off = g_hist[(int64_t)threadIdx.x * nb + blockIdx.x];
    for (int e = threadIdx.x; e < RS_WARPS * 256; e += RS_THREADS) (&cnt[0][0])[e] = 0;
    __syncthreads();
    const int64_t wbase = (int64_t)blockIdx.x * RS_TILE + (int64_t)w * (RS_ITEMS * 32);
    K kreg[RS_ITEMS];
    int32_t vreg[RS_ITEMS];
    int rank[RS_ITEMS];
    const unsigned lt = (1u << lane) - 1u;
    // all of the thread's keys and values are requested before the first one is ranked: one memory round trip
    // instead of one per item (the ranking below is a chain of warp votes that cannot start without its key)
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) {
        const int64_t i = wbase + j * 32 + lane;
        kreg[j] = 0; vreg[j] = 0;
        if (i < n) { kreg[j] = keys_in[i]; vreg[j] = vals_in ? vals_in[i] : (int32_t)i; }
    }
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) {
        int64_t i = wbase + j * 32 + lane;
        bool in = i < n;
        unsigned act = __ballot_sync(KP_FULL, in);
        rank[j] = 0;
        if (in) {
            K key = kreg[j];
            unsigned d = (unsigned)(key >> shift) & 0xFFu;
            unsigned peers = __match_any_sync(act, d);
            int leader = __ffs(peers) - 1;
            int old = 0;
            if (lane == leader) { old = cnt[w][d]; cnt[w][d] = old + __popc(peers); }
            old = __shfl_sync(peers, old, leader);
            rank[j] = old + __popc(peers & lt);
        }
        __syncwarp();
    }
    __syncthreads();
    {
        int d = threadIdx.x;   // 256 threads <-> 256 digits
        // global base of digit d = exclusive scan over the 256 digit totals (block scan, recomputed per tile)
        int tot = dig_tot, incl = tot;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            int y = __shfl_up_sync(KP_FULL, incl, s);
            if (lane >= s) incl += y;
        }
        if (lane == 31) dig_wtot[w] = incl;
        __syncthreads();
        int base = incl - tot;
#pragma unroll
        for (int ww = 0; ww < RS_WARPS; ++ww) base += ww < w ? dig_wtot[ww] : 0;
        int off = base + dig_off;
#pragma unroll
        for (int ww = 0; ww < RS_WARPS; ++ww) { int c = cnt[ww][d]; cnt[ww][d] = off; off += c; }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) {
        int64_t i = wbase + j * 32 + lane;
        if (i < n) {
            unsigned d = (unsigned)(kreg[j] >> shift) & 0xFFu;
            int pos = cnt[w][d] + rank[j];
            keys_out[pos] = kreg[j];
            vals_out[pos] = vreg[j];
        }
    }
}

template <class K>
int sort_pairs(kp_ctx *ctx, int64_t n, int bits, K *d_keys, K *d_keys_tmp, int32_t *d_vals, int32_t *d_vals_tmp,
               K **keys_sorted, int32_t **vals_sorted, bool vals_are_iota)
{
    *keys_sorted = d_keys;
    *vals_sorted = d_vals;
    if (n <= 0) return KP_OK;
    // algorithmic bytes: ceil(b/8) passes x 2 x (key + 4-byte index) x n, plus one histogram read
    KP_PROFB(ctx, "radix_sort", (double)n * (((bits + 7) / 8 < 1 ? 1 : (bits + 7) / 8) * 2.0 * (sizeof(K) + 4.0) + sizeof(K)));
    int nb = (int)kp_blocks(n, RS_TILE);
    int32_t *g_hist, *g_base;
    KP_TRY(kp_ws(ctx, (size_t)256 * nb, &g_hist));
    KP_TRY(kp_ws(ctx, 256, &g_base));
    int passes = (bits + 7) / 8;
    if (passes < 1) passes = 1;
    K *kin = d_keys, *kout = d_keys_tmp;
    int32_t *vin = d_vals, *vout = d_vals_tmp;
    for (int p = 0; p < passes; ++p) {
        int shift = 8 * p;
        k_rs_hist<K><<<nb, RS_THREADS, 0, ctx->stream>>>(kin, n, shift, g_hist, nb, 0);
        KP_LAUNCH_CHECK(ctx);
        k_rs_scan_rows<<<256, 256, 0, ctx->stream>>>(g_hist, nb, g_base);
        KP_LAUNCH_CHECK(ctx);
        k_rs_scatter<K><<<nb, RS_THREADS, 0, ctx->stream>>>(kin, (p == 0 && vals_are_iota) ? nullptr : vin, kout, vout, n,
                                                            shift, g_hist, g_base, nb, 0);
        KP_LAUNCH_CHECK(ctx);
        K *tk = kin; kin = kout; kout = tk;
        int32_t *tv = vin; vin = vout; vout = tv;
    }
    *keys_sorted = kin;
    *vals_sorted = vin;
    return KP_OK;
}
}  // namespace

// batched form (kp_batch.cuh): `nseg` independent sorts of n rows each, one launch per kernel and pass
size_t kp_b_sort_hist_elems(int64_t n) { return (size_t)256 * kp_blocks(n > 0 ? n : 1, RS_TILE); }
int kp_b_sort_pairs_u32(const BLaunch &L, const BSort &W, int64_t n, int passes, uint32_t *keys, uint32_t *keys_tmp,
                        int32_t *vals, int32_t *vals_tmp, int64_t stride)
{
    if (n <= 0 || L.nseg <= 0) return KP_OK;
    kp_ctx *ctx = L.ctx;
    KP_PROFB(ctx, "radix_sort", (double)L.nseg * (double)n * (passes * 2.0 * 8.0 + 4.0));
    const int nb = (int)kp_blocks(n, RS_TILE);
    uint32_t *kin = keys, *kout = keys_tmp;
    int32_t *vin = vals, *vout = vals_tmp;
    for (int p = 0; p < passes; ++p) {
        const int shift = 8 * p;
        k_rs_hist<uint32_t><<<dim3(nb, L.nseg), RS_THREADS, 0, ctx->stream>>>(kin, n, shift, W.g_hist, nb, stride);
        KP_LAUNCH_CHECK(ctx);
        k_rs_scan_rows<<<dim3(256, L.nseg), 256, 0, ctx->stream>>>(W.g_hist, nb, W.g_tot);
        KP_LAUNCH_CHECK(ctx);
        k_rs_scatter<uint32_t><<<dim3(nb, L.nseg), RS_THREADS, 0, ctx->stream>>>(kin, p == 0 ? nullptr : vin, kout, vout, n, shift,
                                                                                 W.g_hist, W.g_tot, nb, stride);
        KP_LAUNCH_CHECK(ctx);
        uint32_t *tk = kin; kin = kout; kout = tk;
        int32_t *tv = vin; vin = vout; vout = tv;
    }
    return KP_OK;
}

// The value of element i entering the first pass is its input position i (iota), synthesised in
// the scatter kernel instead of being read, so d_vals need not be initialised.
int kp_prim_sort_pairs_u64(kp_ctx *ctx, int64_t n, int bits, uint64_t *d_keys, uint64_t *d_keys_tmp, int32_t *d_vals,
                           int32_t *d_vals_tmp, uint64_t **d_keys_sorted, int32_t **d_vals_sorted)
{
    return sort_pairs<uint64_t>(ctx, n, bits, d_keys, d_keys_tmp, d_vals, d_vals_tmp, d_keys_sorted, d_vals_sorted, true);
}
int kp_prim_sort_pairs_u32(kp_ctx *ctx, int64_t n, int bits, uint32_t *d_keys, uint32_t *d_keys_tmp, int32_t *d_vals,
                           int32_t *d_vals_tmp, uint32_t **d_keys_sorted, int32_t **d_vals_sorted)
{
    return sort_pairs<uint32_t>(ctx, n, bits, d_keys, d_keys_tmp, d_vals, d_vals_tmp, d_keys_sorted, d_vals_sorted, true);
}

// ================================================================ bounds ==
namespace {
__global__ void k_bounds_init(int32_t *enc)
{
    if (threadIdx.x < 3) enc[threadIdx.x] = kp_f2ord(INFINITY);
    else if (threadIdx.x < 6) enc[threadIdx.x] = kp_f2ord(-INFINITY);
    else if (threadIdx.x < 8) enc[threadIdx.x] = 0;
}
__global__ void __launch_bounds__(256) k_bounds(const float *xyz, int64_t n, int32_t *enc)
{
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    int cnt = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
        if (isnan(x)) continue;
        mn[0] = fminf(mn[0], x); mn[1] = fminf(mn[1], y); mn[2] = fminf(mn[2], z);
        mx[0] = fmaxf(mx[0], x); mx[1] = fmaxf(mx[1], y); mx[2] = fmaxf(mx[2], z);
        ++cnt;
    }
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            mn[c] = fminf(mn[c], __shfl_xor_sync(KP_FULL, mn[c], s));
            mx[c] = fmaxf(mx[c], __shfl_xor_sync(KP_FULL, mx[c], s));
        }
        cnt += __shfl_xor_sync(KP_FULL, cnt, s);
    }
    if ((threadIdx.x & 31) == 0 && cnt > 0) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            atomicMin(&enc[c], kp_f2ord(mn[c]));
            atomicMax(&enc[3 + c], kp_f2ord(mx[c]));
        }
        atomicAdd(&enc[6], cnt);
    }
}
}  // namespace

int kp_prim_bounds(kp_ctx *ctx, const float *d_xyz, int64_t n, int32_t *d_bounds_enc)
{
    KP_PROFB(ctx, "bounds", (double)n * 12.0);
    k_bounds_init<<<1, 32, 0, ctx->stream>>>(d_bounds_enc);
    KP_LAUNCH_CHECK(ctx);
    if (n > 0) {
        unsigned nb = kp_blocks(n, 256 * 8);
        unsigned cap = (unsigned)ctx->sm_count * 8;
        if (nb > cap) nb = cap;
        k_bounds<<<nb, 256, 0, ctx->stream>>>(d_xyz, n, d_bounds_enc);
        KP_LAUNCH_CHECK(ctx);
    }
    return KP_OK;
}

static inline float kp_host_ord2f(int32_t i)
{
    int32_t b = i >= 0 ? i : i ^ 0x7fffffff;
    float f;
    memcpy(&f, &b, 4);
    return f;
}

int kp_prim_bounds_fetch(kp_ctx *ctx, const float *d_xyz, int64_t n, float *h_bounds6, int64_t *h_nvalid)
{
    int32_t *enc = (int32_t *)ctx->d_scratch;
    KP_TRY(kp_prim_bounds(ctx, d_xyz, n, enc));
    KP_TRY(kp_fetch_scratch(ctx, 8 * sizeof(int32_t)));
    const int32_t *h = (const int32_t *)ctx->h_scratch;
    for (int c = 0; c < 6; ++c) h_bounds6[c] = kp_host_ord2f(h[c]);
    if (h_nvalid) *h_nvalid = h[6];
    return KP_OK;
}

// ========================================================= canonical sum ==
namespace {
// one warp per group of 1024 values: lane t adds x[t], x[t+32], ... in order, then the butterfly.
// mode selects what is summed (SOR statistics: the two sums run on transformed views of the mean array, so the
// transform rides in the first level instead of a pass that writes a transformed copy):
//   0: x[i]      1: x[i] > 0 ? x[i] : 0      2: x[i] > 0 ? (x[i] - mu)^2 : 0  with mu = *aux / aux_div
__global__ void __launch_bounds__(256) k_csum_level(const double *x, int64_t n, double *out, int64_t groups, int mode,
                                                    const double *aux, double aux_div)
{
    int64_t g = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= groups) return;
    const int lane = threadIdx.x & 31;
    const double mu = mode == 2 ? __ddiv_rn(*aux, aux_div) : 0.0;
    int64_t lo = g * 1024;
    double acc = 0.0;
#pragma unroll 4
    for (int r = 0; r < 32; ++r) {
        int64_t i = lo + r * 32 + lane;
        if (i < n) {
            double v = x[i];
            if (mode == 1) v = v > 0 ? v : 0.0;
            else if (mode == 2) { const double d = __dsub_rn(v, mu); v = v > 0 ? __dmul_rn(d, d) : 0.0; }
            acc = __dadd_rn(acc, v);
        }
    }
    acc = kp_butterfly_sum(acc);
    if (lane == 0) out[g] = acc;
}
// every further level in one CTA: cur[0..cn) -> groups of 1024 -> ... -> out[0]; scratch follows cur
__global__ void __launch_bounds__(1024) k_csum_rest(double *cur, int64_t cn, double *out)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *nxt = cur + cn;
    for (;;) {
        const int64_t groups = (cn + 1023) / 1024;
        double *dst = groups == 1 ? out : nxt;
        for (int64_t g = warp; g < groups; g += 32) {
            const int64_t lo = g * 1024;
            double acc = 0.0;
            for (int r = 0; r < 32; ++r) {
                const int64_t i = lo + r * 32 + lane;
                if (i < cn) acc = __dadd_rn(acc, cur[i]);
            }
            acc = kp_butterfly_sum(acc);
            if (lane == 0) dst[g] = acc;
        }
        if (groups == 1) break;
        __syncthreads();
        cur = nxt; nxt = nxt + groups; cn = groups;
    }
}
__global__ void k_count_u8(const uint8_t *m, int64_t n, int32_t *total)
{
    int c = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) c += m[i] != 0;
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) c += __shfl_xor_sync(KP_FULL, c, s);
    // one atomic per CTA (same-address atomics serialise at ~5 ns each: per warp they were the whole kernel)
    __shared__ int wc[32];
    if ((threadIdx.x & 31) == 0) wc[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) c += wc[w];
        if (c) atomicAdd(total, c);
    }
}
}  // namespace

int kp_prim_csum_mode(kp_ctx *ctx, const double *d_x, int64_t n, double *d_tmp, double *d_out, int mode, const double *d_aux,
                      double aux_div)
{
    if (n <= 0) {
        KP_CUDA(ctx, cudaMemsetAsync(d_out, 0, sizeof(double), ctx->stream));
        return KP_OK;
    }
    const int64_t groups = (n + 1023) / 1024;
    k_csum_level<<<kp_blocks(groups, 8), 256, 0, ctx->stream>>>(d_x, n, groups == 1 ? d_out : d_tmp, groups, mode, d_aux, aux_div);
    KP_LAUNCH_CHECK(ctx);
    if (groups > 1) {
        k_csum_rest<<<1, 1024, 0, ctx->stream>>>(d_tmp, groups, d_out);
        KP_LAUNCH_CHECK(ctx);
    }
    return KP_OK;
}

int kp_prim_csum(kp_ctx *ctx, const double *d_x, int64_t n, double *d_tmp, double *d_out)
{
    return kp_prim_csum_mode(ctx, d_x, n, d_tmp, d_out, 0, nullptr, 1.0);
}

int kp_prim_count_u8(kp_ctx *ctx, const uint8_t *d_mask, int64_t n, int32_t *d_total)
{
    KP_CUDA(ctx, cudaMemsetAsync(d_total, 0, sizeof(int32_t), ctx->stream));
    if (n <= 0) return KP_OK;
    unsigned nb = kp_blocks(n, 256 * 16);
    unsigned cap = (unsigned)ctx->sm_count * 8;
    if (nb > cap) nb = cap;
    k_count_u8<<<nb, 256, 0, ctx->stream>>>(d_mask, n, d_total);
    KP_LAUNCH_CHECK(ctx);
    return KP_OK;
}
