// kp_ctx.cu -- context, workspace, memory and timing entry points of the C ABI.
#include <sched.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <map>
#include <thread>
#include "kp_common.cuh"

static thread_local std::string g_last_err;

int kp_set_err(kp_ctx *ctx, int code, const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_err = buf;
    if (ctx) ctx->err = buf;
    return code;
}

extern "C" {

const char *kp_version(void) { return "kinectpy_b200 0.1 (sm_100a)"; }

int kp_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char *kp_last_error(kp_ctx *ctx) { return ctx ? ctx->err.c_str() : g_last_err.c_str(); }

int kp_ctx_create(int device, kp_ctx **out)
{
    if (!out) return kp_set_err(nullptr, KP_E_ARG, "kp_ctx_create: out is NULL");
    *out = nullptr;
    int n = kp_device_count();
    if (n <= 0) return kp_set_err(nullptr, KP_E_NODEVICE, "no CUDA device visible: kinectpy_b200 has no CPU path");
    if (device < 0 || device >= n) return kp_set_err(nullptr, KP_E_ARG, "device %d out of range (have %d)", device, n);
    kp_ctx *ctx = new kp_ctx();
    ctx->device = device;
    KP_CUDA(ctx, cudaSetDevice(device));
    KP_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    KP_CUDA(ctx, cudaMallocHost((void **)&ctx->h_scratch, ctx->scratch_bytes));
    KP_CUDA(ctx, cudaMalloc((void **)&ctx->d_scratch, ctx->scratch_bytes));
    KP_CUDA(ctx, cudaMemsetAsync(ctx->d_scratch, 0, ctx->scratch_bytes, ctx->stream));
    KP_CUDA(ctx, cudaMalloc((void **)&ctx->d_lb_state, sizeof(unsigned long long) * (kp_ctx::LB_TILES + kp_ctx::LB_TILES / 32 + 32)));
    KP_CUDA(ctx, cudaMemsetAsync(ctx->d_lb_state, 0, sizeof(unsigned long long) * (kp_ctx::LB_TILES + kp_ctx::LB_TILES / 32 + 32), ctx->stream));
    KP_CUDA(ctx, cudaEventCreate(&ctx->t0));
    KP_CUDA(ctx, cudaEventCreate(&ctx->t1));
    cudaDeviceProp prop;
    KP_CUDA(ctx, cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;
    *out = ctx;
    return KP_OK;
}

int kp_ctx_destroy(kp_ctx *ctx)
{
    if (!ctx) return KP_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto &b : ctx->blocks) cudaFree(b.ptr);
    for (auto &e : ctx->ev_pool) cudaEventDestroy(e);
    for (auto &p : ctx->prof) { cudaEventDestroy(p.e0); cudaEventDestroy(p.e1); }
    if (ctx->l2_flush) cudaFree(ctx->l2_flush);
    cudaFreeHost(ctx->h_scratch);
    cudaFree(ctx->d_scratch);
    if (ctx->d_lb_state) cudaFree(ctx->d_lb_state);
    if (ctx->ev_block) cudaEventDestroy(ctx->ev_block);
    cudaEventDestroy(ctx->t0);
    cudaEventDestroy(ctx->t1);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return KP_OK;
}

void *kp_ctx_stream(kp_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int kp_sync(kp_ctx *ctx)
{
    KP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KP_OK;
}

int kp_malloc(kp_ctx *ctx, size_t bytes, void **d_ptr)
{
    if (!ctx || !d_ptr) return kp_set_err(ctx, KP_E_ARG, "kp_malloc: NULL argument");
    KP_CUDA(ctx, cudaSetDevice(ctx->device));
    KP_CUDA(ctx, cudaMallocAsync(d_ptr, bytes ? bytes : 16, ctx->stream));
    return KP_OK;
}

int kp_free(kp_ctx *ctx, void *d_ptr)
{
    if (!d_ptr) return KP_OK;
    KP_CUDA(ctx, cudaFreeAsync(d_ptr, ctx->stream));
    return KP_OK;
}

int kp_host_alloc(size_t bytes, void **h_ptr)
{
    if (cudaMallocHost(h_ptr, bytes ? bytes : 16) != cudaSuccess)
        return kp_set_err(nullptr, KP_E_NOMEM, "cudaMallocHost(%zu) failed: %s", bytes, cudaGetErrorString(cudaGetLastError()));
    return KP_OK;
}

int kp_host_free(void *h_ptr)
{
    if (h_ptr) cudaFreeHost(h_ptr);
    return KP_OK;
}

int kp_memcpy_h2d(kp_ctx *ctx, void *d_dst, const void *h_src, size_t bytes)
{
    if (bytes) KP_CUDA(ctx, cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return KP_OK;
}

int kp_memcpy_d2h(kp_ctx *ctx, void *h_dst, const void *d_src, size_t bytes)
{
    if (bytes) KP_CUDA(ctx, cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    KP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KP_OK;
}

int kp_memcpy_d2d(kp_ctx *ctx, void *d_dst, const void *d_src, size_t bytes)
{
    if (bytes) KP_CUDA(ctx, cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return KP_OK;
}

int kp_memset(kp_ctx *ctx, void *d_dst, int value, size_t bytes)
{
    if (bytes) KP_CUDA(ctx, cudaMemsetAsync(d_dst, value, bytes, ctx->stream));
    return KP_OK;
}

int kp_timer_start(kp_ctx *ctx)
{
    KP_CUDA(ctx, cudaEventRecord(ctx->t0, ctx->stream));
    return KP_OK;
}

int kp_timer_stop(kp_ctx *ctx, float *h_ms)
{
    KP_CUDA(ctx, cudaEventRecord(ctx->t1, ctx->stream));
    KP_CUDA(ctx, cudaEventSynchronize(ctx->t1));
    KP_CUDA(ctx, cudaEventElapsedTime(h_ms, ctx->t0, ctx->t1));
    return KP_OK;
}

int64_t kp_launch_count(kp_ctx *ctx) { return ctx ? ctx->launches : 0; }

int kp_flush_l2(kp_ctx *ctx)
{
    const size_t bytes = 256u << 20;
    if (!ctx->l2_flush) KP_CUDA(ctx, cudaMalloc(&ctx->l2_flush, bytes));
    KP_CUDA(ctx, cudaMemsetAsync(ctx->l2_flush, 0x5a, bytes, ctx->stream));
    return KP_OK;
}

int kp_profile_enable(kp_ctx *ctx, int on)
{
    ctx->prof_on = on != 0;
    return KP_OK;
}

int kp_profile_reset(kp_ctx *ctx)
{
    KP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (auto &p : ctx->prof) { ctx->ev_pool.push_back(p.e0); ctx->ev_pool.push_back(p.e1); }
    ctx->prof.clear();
    return KP_OK;
}

int kp_profile_read(kp_ctx *ctx, int max_entries, const char **h_names, double *h_ms, int64_t *h_calls,
                    double *h_bytes, int *h_n)
{
    KP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    struct Acc { double ms = 0, bytes = 0; int64_t calls = 0; };
    std::vector<const char *> order;
    std::map<std::string, Acc> acc;
    for (auto &p : ctx->prof) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, p.e0, p.e1) != cudaSuccess) { cudaGetLastError(); continue; }
        auto it = acc.find(p.name);
        if (it == acc.end()) { order.push_back(p.name); it = acc.emplace(p.name, Acc()).first; }
        it->second.ms += ms; it->second.bytes += p.bytes; it->second.calls += 1;
    }
    int n = 0;
    for (auto nm : order) {
        if (n >= max_entries) break;
        const Acc &a = acc[nm];
        h_names[n] = nm;
        h_ms[n] = a.ms;
        h_calls[n] = a.calls;
        if (h_bytes) h_bytes[n] = a.bytes;
        ++n;
    }
    *h_n = n;
    return KP_OK;
}

}  // extern "C"

// ------------------------------------------------------------- workspace --
void kp_ws_reset(kp_ctx *ctx)
{
    if (ctx->blocks.size() > 1) {
        // the previous call outgrew the arena: merge into one block of the combined size
        cudaStreamSynchronize(ctx->stream);
        size_t total = 0;
        for (auto &b : ctx->blocks) { total += b.cap; cudaFree(b.ptr); }
        ctx->blocks.clear();
        char *p = nullptr;
        if (cudaMalloc((void **)&p, total) == cudaSuccess) ctx->blocks.push_back({p, total});
        else cudaGetLastError();
    }
    ctx->ws_off = 0;
}

int kp_ws_alloc(kp_ctx *ctx, size_t bytes, void **out)
{
    bytes = (bytes + 255) & ~(size_t)255;
    if (!ctx->blocks.empty()) {
        auto &b = ctx->blocks.back();
        if (ctx->ws_off + bytes <= b.cap) {
            *out = b.ptr + ctx->ws_off;
            ctx->ws_off += bytes;
            return KP_OK;
        }
    }
    size_t cap = ctx->blocks.empty() ? (size_t)64 << 20 : ctx->blocks.back().cap * 2;
    if (cap < bytes) cap = bytes;
    char *p = nullptr;
    KP_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaError_t e = cudaMalloc((void **)&p, cap);
    if (e != cudaSuccess) { cudaGetLastError(); return kp_set_err(ctx, KP_E_NOMEM, "workspace cudaMalloc(%zu) failed: %s", cap, cudaGetErrorString(e)); }
    ctx->blocks.push_back({p, cap});
    *out = p;
    ctx->ws_off = bytes;
    return KP_OK;
}

int kp_fetch_scratch(kp_ctx *ctx, size_t bytes)
{
    KP_CUDA(ctx, cudaMemcpyAsync(ctx->h_scratch, ctx->d_scratch, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return kp_stream_wait(ctx);
}

// How a host thread waits for its stream: cudaStreamSynchronize.  The frame engine has one host thread per GPU and waits
// once per call, so there is no wait policy to choose; `KP_SYNC=block` (an explicit request, e.g. many single-operation
// callers sharing few cores) sleeps on a blocking event instead.  Nothing is inferred from the environment.
static int kp_sync_blocking()
{
    static int mode = -1;
    if (mode >= 0) return mode;
    const char *e = getenv("KP_SYNC");
    return mode = (e && !strcmp(e, "block")) ? 1 : 0;
}

int kp_stream_wait(kp_ctx *ctx)
{
    if (!kp_sync_blocking()) {
        KP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return KP_OK;
    }
    if (!ctx->ev_block) KP_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_block, cudaEventBlockingSync | cudaEventDisableTiming));
    KP_CUDA(ctx, cudaEventRecord(ctx->ev_block, ctx->stream));
    KP_CUDA(ctx, cudaEventSynchronize(ctx->ev_block));
    return KP_OK;
}

KpProfScope::KpProfScope(kp_ctx *c, const char *name, double bytes) : ctx(c), idx(0), on(c->prof_on)
{
    if (!on) return;
    KpProfEntry e;
    e.name = name;
    e.bytes = bytes;
    for (cudaEvent_t *ev : {&e.e0, &e.e1}) {
        if (!ctx->ev_pool.empty()) { *ev = ctx->ev_pool.back(); ctx->ev_pool.pop_back(); }
        else cudaEventCreate(ev);
    }
    cudaEventRecord(e.e0, ctx->stream);
    idx = ctx->prof.size();
    ctx->prof.push_back(e);
}
KpProfScope::~KpProfScope()
{
    if (on) cudaEventRecord(ctx->prof[idx].e1, ctx->stream);
}
