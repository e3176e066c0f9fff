// api.cu — the C ABI of libila_b200 (include/ila_b200.h): host-pointer entry points that own device memory,
// streams and the multi-GPU sharding, plus the device-pointer (`_dev`) forms.  No torch types, no CPU fallback.
#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "ila_common.cuh"
#include "ila_kernels.cuh"

namespace ila {

thread_local char g_last_error[512] = "";
std::atomic<int64_t> g_launches{0};

// ---- per-device grow-only buffer pool (host entry points only) ---------------------------------------
constexpr int kMaxDev = 16;
constexpr int kSlots = 24;
constexpr int kPipe = 4;   // copy/compute pipeline depth of the host entry points (= staging workers per device)
constexpr int kMaxChunkEv = 48;   // chunks of the pinned-buffer pipeline (the last one takes the remainder)
constexpr int kStage = 2;  // pinned staging slots per pipeline stream (double buffering)
struct DevState {
    bool init = false;
    cudaStream_t stream = nullptr;
    cudaStream_t pipe[kPipe] = {};
    void *buf[kSlots] = {};
    size_t cap[kSlots] = {};
    // pinned staging buffers of the pageable-caller path (grow-only, cudaHostAlloc) and their completion events
    void *hin[kPipe][kStage] = {}, *hout[kPipe][kStage] = {};
    size_t hin_cap[kPipe][kStage] = {}, hout_cap[kPipe][kStage] = {};
    cudaEvent_t hev[kPipe][kStage] = {};
    cudaEvent_t cev[kMaxChunkEv][2] = {};   // per chunk of the pinned pipeline: input arrived / kernel done
};
static DevState g_dev[kMaxDev];
static std::mutex g_mutex;

static int dev_init(int d)
{
    ILA_CUDA_CHECK(cudaSetDevice(d));
    if (!g_dev[d].init) {
        ILA_CUDA_CHECK(cudaStreamCreateWithFlags(&g_dev[d].stream, cudaStreamNonBlocking));
        for (int i = 0; i < kPipe; i++) ILA_CUDA_CHECK(cudaStreamCreateWithFlags(&g_dev[d].pipe[i], cudaStreamNonBlocking));
        for (int i = 0; i < kPipe; i++)
            for (int k = 0; k < kStage; k++) ILA_CUDA_CHECK(cudaEventCreateWithFlags(&g_dev[d].hev[i][k], cudaEventDisableTiming));
        for (int i = 0; i < kMaxChunkEv; i++)
            for (int k = 0; k < 2; k++) ILA_CUDA_CHECK(cudaEventCreateWithFlags(&g_dev[d].cev[i][k], cudaEventDisableTiming));
        g_dev[d].init = true;
    }
    return 0;
}
static int dev_buf(int d, int slot, size_t bytes, void **out)
{
    DevState &S = g_dev[d];
    if (bytes == 0) bytes = 8;
    if (S.cap[slot] < bytes) {
        if (S.buf[slot]) ILA_CUDA_CHECK(cudaFree(S.buf[slot]));
        S.buf[slot] = nullptr;
        S.cap[slot] = 0;
        ILA_CUDA_CHECK(cudaMalloc(&S.buf[slot], bytes));
        S.cap[slot] = bytes;
    }
    *out = S.buf[slot];
    return 0;
}
static int host_stage_buf(void **buf, size_t *cap, size_t bytes)
{
    if (*cap >= bytes) return 0;
    if (*buf) ILA_CUDA_CHECK(cudaFreeHost(*buf));
    *buf = nullptr;
    *cap = 0;
    ILA_CUDA_CHECK(cudaHostAlloc(buf, bytes, cudaHostAllocPortable));
    *cap = bytes;
    return 0;
}
// true when the driver can DMA straight from/to this host pointer (cudaHostAlloc / cudaHostRegister / managed memory)
static bool host_ptr_pinned(const void *p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
}
static std::atomic<int64_t> g_chunk_min{0};   // tuning hooks of the pinned pipeline (ila_set_host_chunking): smallest chunk,
static std::atomic<int> g_chunk_div{0};       // and the divisor of the geometric schedule (chunk = remaining / div)
static std::atomic<int> g_host_copy_mode{0};   // 0 auto, 1 always direct cudaMemcpyAsync, 2 always staged (ila_set_host_copy_mode)

// Persistent host worker threads of the staged copy path (created on first use, never destroyed: they sleep on a
// condition variable between calls; host entry points are serialised by g_mutex, so one call uses the pool at a time).
class WorkerPool {
    std::mutex m_;
    std::condition_variable cv_, done_;
    std::deque<std::function<void()>> q_;
    std::vector<std::thread> th_;
    int pending_ = 0;

  public:
    void ensure(int n)
    {
        while ((int)th_.size() < n)
            th_.emplace_back([this] {
                for (;;) {
                    std::function<void()> job;
                    {
                        std::unique_lock<std::mutex> lk(m_);
                        cv_.wait(lk, [this] { return !q_.empty(); });
                        job = std::move(q_.front());
                        q_.pop_front();
                    }
                    job();
                    {
                        std::lock_guard<std::mutex> lk(m_);
                        if (--pending_ == 0) done_.notify_all();
                    }
                }
            });
    }
    void submit(std::function<void()> f)
    {
        {
            std::lock_guard<std::mutex> lk(m_);
            q_.push_back(std::move(f));
            pending_++;
        }
        cv_.notify_one();
    }
    void wait()
    {
        std::unique_lock<std::mutex> lk(m_);
        done_.wait(lk, [this] { return pending_ == 0; });
    }
};
static WorkerPool *g_pool = nullptr;   // leaked on purpose (its threads must not be joined from a static destructor)

template <class T>
static int dev_buf_t(int d, int slot, size_t count, T **out)
{
    void *p = nullptr;
    int rc = dev_buf(d, slot, count * sizeof(T), &p);
    *out = static_cast<T *>(p);
    return rc;
}

// logical shard d -> physical device: a single-GPU call runs on the CURRENT device (one process per GPU under
// torchrun selects it with ila_set_device); a multi-GPU call uses devices 0..G-1 of this process.
static int g_base_dev = 0;
static inline int phys(int d, int G) { return G == 1 ? g_base_dev : d; }

static int resolve_ngpus(int ngpus, int *out)
{
    int cnt = 0;
    cudaError_t e = cudaGetDeviceCount(&cnt);
    if (e != cudaSuccess || cnt <= 0) {
        cudaGetLastError();
        return set_error(ILA_ERR_NOGPU, "no CUDA device visible: libila_b200 has no CPU fallback");
    }
    if (cnt > kMaxDev) cnt = kMaxDev;
    if (ngpus <= 0 || ngpus > cnt) ngpus = cnt;
    *out = ngpus;
    return 0;
}

// Every host entry point holds g_mutex for its whole duration and declares one HostGuard right after taking it: the guard
// records the caller's current device (the device single-GPU calls run on) UNDER the lock, and on every exit that did not
// reach the normal end (`done`) it drains the streams of all devices the call may have used and restores the caller's
// device — so an error return never leaves async copies into the caller's buffers in flight.
struct HostGuard {
    int G;
    bool done = false;
    explicit HostGuard(int G_) : G(G_)
    {
        int cur = 0;
        if (cudaGetDevice(&cur) != cudaSuccess) { cudaGetLastError(); cur = 0; }
        if (cur < 0 || cur >= kMaxDev) cur = 0;
        g_base_dev = cur;
    }
    ~HostGuard()
    {
        if (done) return;
        for (int d = 0; d < G; d++) {
            const int pd = phys(d, G);
            if (pd < 0 || pd >= kMaxDev || !g_dev[pd].init) continue;
            if (cudaSetDevice(pd) != cudaSuccess) { cudaGetLastError(); continue; }
            cudaStreamSynchronize(g_dev[pd].stream);
            for (int i = 0; i < kPipe; i++) cudaStreamSynchronize(g_dev[pd].pipe[i]);
        }
        cudaSetDevice(g_base_dev);
        cudaGetLastError();
    }
};

// ---- FP64 peak probe ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dfma_probe_kernel(double *out, int iters)
{
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// dependent-chain latency probe (one warp): cycles per dependent DFMA / DMUL+DADD pair / sqrt / div
__global__ void latency_probe_kernel(double *out, double seed)
{
    const int iters = 4096;
    double a = seed + threadIdx.x * 1e-9;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) a = fma(a, 0.999999, 1e-7);
    long long t1 = clock64();
    double b = a;
    for (int i = 0; i < iters; i++) b = __dadd_rn(__dmul_rn(b, 0.999999), 1e-7);
    long long t2 = clock64();
    double c = fabs(b) + 2.0;
    for (int i = 0; i < iters; i++) c = __dsqrt_rn(c) + 1.5;
    long long t3 = clock64();
    double d = c;
    for (int i = 0; i < iters; i++) d = __ddiv_rn(3.0, d) + 0.5;
    long long t4 = clock64();
    if (threadIdx.x == 0) {
        out[0] = (double)(t1 - t0) / iters;
        out[1] = (double)(t2 - t1) / iters;
        out[2] = (double)(t3 - t2) / iters;
        out[3] = (double)(t4 - t3) / iters;
        out[4] = a + b + c + d;
    }
}

} // namespace ila

using namespace ila;

extern "C" {

int ila_version(void) { return 100; }

int ila_device_count(void)
{
    int cnt = 0;
    if (cudaGetDeviceCount(&cnt) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return cnt;
}

const char *ila_last_error(void) { return g_last_error; }

int ila_set_device(int dev)
{
    ILA_CUDA_CHECK(cudaSetDevice(dev));
    return 0;
}

int ila_get_device(void)
{
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return -1; }
    return dev;
}

int ila_release(void)
{
    // frees the per-device buffer pools and streams of the host entry points (they are re-created on demand)
    std::lock_guard<std::mutex> lock(g_mutex);
    int cur = 0;
    if (cudaGetDevice(&cur) != cudaSuccess) { cudaGetLastError(); return 0; }
    for (int d = 0; d < kMaxDev; d++) {
        DevState &S = g_dev[d];
        if (!S.init) continue;
        if (cudaSetDevice(d) != cudaSuccess) { cudaGetLastError(); continue; }
        for (int k = 0; k < kSlots; k++) {
            if (S.buf[k]) cudaFree(S.buf[k]);
            S.buf[k] = nullptr;
            S.cap[k] = 0;
        }
        if (S.stream) cudaStreamDestroy(S.stream);
        for (int i = 0; i < kPipe; i++) {
            if (S.pipe[i]) cudaStreamDestroy(S.pipe[i]);
            for (int k = 0; k < kStage; k++) {
                if (S.hin[i][k]) cudaFreeHost(S.hin[i][k]);
                if (S.hout[i][k]) cudaFreeHost(S.hout[i][k]);
                if (S.hev[i][k]) cudaEventDestroy(S.hev[i][k]);
            }
        }
        for (int i = 0; i < kMaxChunkEv; i++)
            for (int k = 0; k < 2; k++)
                if (S.cev[i][k]) cudaEventDestroy(S.cev[i][k]);
        S = DevState{};
    }
    cudaSetDevice(cur);
    return 0;
}

int64_t ila_kernel_launches(void) { return g_launches.load(); }

int ila_fp64_peak_probe(double *tflops_out, double *ms_out)
{
    int cnt = ila_device_count();
    if (cnt <= 0) return set_error(ILA_ERR_NOGPU, "no CUDA device visible");
    int dev = 0, sms = 0;
    ILA_CUDA_CHECK(cudaGetDevice(&dev));
    ILA_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int blocks = sms * 8, threads = 256, iters = 1 << 15;
    double *d = nullptr;
    ILA_CUDA_CHECK(cudaMalloc(&d, sizeof(double) * blocks * threads));
    cudaEvent_t e0, e1;
    ILA_CUDA_CHECK(cudaEventCreate(&e0));
    ILA_CUDA_CHECK(cudaEventCreate(&e1));
    dfma_probe_kernel<<<blocks, threads>>>(d, 1 << 10);   // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        ILA_CUDA_CHECK(cudaEventRecord(e0));
        dfma_probe_kernel<<<blocks, threads>>>(d, iters);
        ILA_CUDA_CHECK(cudaEventRecord(e1));
        ILA_CUDA_CHECK(cudaEventSynchronize(e1));
        float ms = 0;
        ILA_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    count_launch(4);
    ILA_CUDA_CHECK(cudaGetLastError());
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    const double flops = 2.0 * 8.0 * (double)iters * (double)blocks * threads;
    if (tflops_out) *tflops_out = flops / (best * 1e-3) / 1e12;
    if (ms_out) *ms_out = best;
    return 0;
}

/* cycles per dependent op on one warp: out[0] DFMA, out[1] DMUL+DADD, out[2] DSQRT+DADD, out[3] DDIV+DADD */
int ila_fp64_latency_probe(double *out4)
{
    if (ila_device_count() <= 0) return set_error(ILA_ERR_NOGPU, "no CUDA device visible");
    double *d = nullptr;
    ILA_CUDA_CHECK(cudaMalloc(&d, sizeof(double) * 8));
    latency_probe_kernel<<<1, 32>>>(d, 1.0);
    latency_probe_kernel<<<1, 32>>>(d, 1.0);
    count_launch(2);
    ILA_CUDA_CHECK(cudaDeviceSynchronize());
    double h[8];
    ILA_CUDA_CHECK(cudaMemcpy(h, d, sizeof(double) * 5, cudaMemcpyDeviceToHost));
    cudaFree(d);
    for (int i = 0; i < 4; i++) out4[i] = h[i];
    return 0;
}

// ================================ K3: Toeplitz-tail QL =================================================

// One device's share of a K3 host call through pinned staging buffers (the caller's arrays are ordinary pageable memory —
// what a Julia `ccall` passes).  kPipe worker threads, each with its own stream and kStage pinned slots: memcpy λ-chunk
// into a slot -> H2D -> kernel -> D2H into the slot -> (when the slot comes round again) memcpy the results out.  The
// host copies of the workers run in parallel with each other and with the GPU work of the other slots.
struct TzStageJob {
    bool real_path;
    int out_mode, l, u, branch_rank, pd;
    const double *a, *lambda;
    int64_t lo, cnt, chunk;
    double *d_lam, *d_L;
    void *d_st;
    double *L11;
    void *status;
    bool stage_in, stage_out;   // which side goes through the pinned slots (a pinned caller buffer is used directly)
};

static int tz_stage_worker(const TzStageJob &J, int t)
{
    DevState &S = g_dev[J.pd];
    ILA_CUDA_CHECK(cudaSetDevice(J.pd));
    const int w = J.real_path ? 1 : 2, ow = (J.out_mode == 1) ? 1 : w, sb = (J.out_mode == 1) ? 1 : 4;
    const size_t in_b = sizeof(double) * w, out_b = sizeof(double) * ow;
    // slot layout: L11 chunk, then the status chunk at an 8-aligned offset
    const size_t st_off = (size_t)J.chunk * out_b;
    cudaStream_t s = S.pipe[t];
    const int64_t nchunks = (J.cnt + J.chunk - 1) / J.chunk;
    struct Pending { int64_t c0 = -1, cc = 0; } pend[kStage];
    auto flush = [&](int k) -> int {
        if (pend[k].c0 < 0) return 0;
        ILA_CUDA_CHECK(cudaEventSynchronize(S.hev[t][k]));
        const char *src = static_cast<const char *>(S.hout[t][k]);
        memcpy(reinterpret_cast<char *>(J.L11) + (size_t)(J.lo + pend[k].c0) * out_b, src, (size_t)pend[k].cc * out_b);
        if (J.status)
            memcpy(static_cast<char *>(J.status) + (size_t)(J.lo + pend[k].c0) * sb, src + st_off, (size_t)pend[k].cc * sb);
        pend[k].c0 = -1;
        return 0;
    };
    int rc = 0, it = 0;
    for (int64_t ci = t; ci < nchunks; ci += kPipe, it++) {
        const int k = it % kStage;
        const int64_t c0 = ci * J.chunk, cc = (c0 + J.chunk <= J.cnt) ? J.chunk : J.cnt - c0;
        if ((rc = flush(k))) return rc;
        const double *src_lam = J.lambda + (size_t)(J.lo + c0) * w;
        if (J.stage_in) {
            if (it >= kStage && !J.stage_out) ILA_CUDA_CHECK(cudaEventSynchronize(S.hev[t][k]));   // slot's last H2D done
            memcpy(S.hin[t][k], src_lam, (size_t)cc * in_b);
            src_lam = static_cast<const double *>(S.hin[t][k]);
        }
        ILA_CUDA_CHECK(cudaMemcpyAsync(J.d_lam + c0 * w, src_lam, (size_t)cc * in_b, cudaMemcpyHostToDevice, s));
        if (J.out_mode == 1)
            rc = launch_ql_toeplitz_l11re(J.l, J.u, J.a, J.branch_rank, J.d_lam + c0 * w, cc, J.d_L + c0,
                                          J.status ? static_cast<uint8_t *>(J.d_st) + c0 : nullptr, s);
        else
            rc = launch_ql_toeplitz(J.real_path, J.l, J.u, J.a, J.branch_rank, J.d_lam + c0 * w, cc, J.d_L + c0 * w, nullptr,
                                    static_cast<int32_t *>(J.d_st) + c0, s);
        if (rc) return rc;
        if (J.stage_out) {
            char *dst = static_cast<char *>(S.hout[t][k]);
            ILA_CUDA_CHECK(cudaMemcpyAsync(dst, J.d_L + c0 * ow, (size_t)cc * out_b, cudaMemcpyDeviceToHost, s));
            if (J.status)
                ILA_CUDA_CHECK(cudaMemcpyAsync(dst + st_off, static_cast<char *>(J.d_st) + (size_t)c0 * sb, (size_t)cc * sb, cudaMemcpyDeviceToHost, s));
            pend[k].c0 = c0;
            pend[k].cc = cc;
        } else {
            ILA_CUDA_CHECK(cudaMemcpyAsync(reinterpret_cast<char *>(J.L11) + (size_t)(J.lo + c0) * out_b, J.d_L + c0 * ow, (size_t)cc * out_b, cudaMemcpyDeviceToHost, s));
            if (J.status)
                ILA_CUDA_CHECK(cudaMemcpyAsync(static_cast<char *>(J.status) + (size_t)(J.lo + c0) * sb, static_cast<char *>(J.d_st) + (size_t)c0 * sb, (size_t)cc * sb, cudaMemcpyDeviceToHost, s));
        }
        ILA_CUDA_CHECK(cudaEventRecord(S.hev[t][k], s));
    }
    for (int j = 0; j < kStage; j++)
        if ((rc = flush((it + j) % kStage))) return rc;
    return 0;
}

// out_mode 0: L11 has the symbol's scalar type, status int32.  out_mode 1 (complex symbols only): L11 is Float64 (L[1,1] of a
// complex-eltype factorization is real), status uint8 or NULL (the NaN payload carries it) — 24-25 B/shift instead of 36.
static int ql_toeplitz_host(bool real_path, int out_mode, int l, int u, const double *a, const double *lambda, int64_t n,
                            int branch_rank, double *L11, double *tail_opt, void *status, int ngpus)
{
    if (n < 0 || !a || (n > 0 && (!lambda || !L11 || (!status && out_mode != 1)))) return set_error(ILA_ERR_ARG, "ql_toeplitz: null pointer / negative n");
    int G = 0;
    int rc = resolve_ngpus(ngpus, &G);
    if (rc) return rc;
    if (n == 0) return 0;
    std::lock_guard<std::mutex> lock(g_mutex);
    HostGuard guard(G);
    const int w = real_path ? 1 : 2;             // doubles per scalar
    const int ow = (out_mode == 1) ? 1 : w;      // doubles per L11 entry
    const int sb = (out_mode == 1) ? 1 : 4;      // bytes per status entry
    const int m = l + 2, tl = 2 * m + 2;
    if (G > n) G = (int)n;
    const int64_t per = (n + G - 1) / G;
    // pageable caller buffers (a Julia Vector): the driver's own pageable path stages every copy through one bounce buffer and
    // blocks the calling thread; kPipe workers with pinned slots keep PCIe busy in both directions instead
    const int mode = g_host_copy_mode.load();
    const bool pin_in = host_ptr_pinned(lambda), pin_out = host_ptr_pinned(L11) && (!status || host_ptr_pinned(status));
    const bool staged = !tail_opt && (mode == 2 || (mode == 0 && !(pin_in && pin_out) && n >= 4096));
    const bool stage_in = staged && (mode == 2 || !pin_in), stage_out = staged && (mode == 2 || !pin_out);
    std::vector<TzStageJob> jobs;
    std::vector<std::function<int()>> direct;      // per-device enqueue loops of the pinned path
    for (int d = 0; d < G; d++) {
        const int64_t lo = d * per, cnt = (lo + per <= n ? per : n - lo);
        if (cnt <= 0) continue;
        const int pd = phys(d, G);
        if ((rc = dev_init(pd))) return rc;
        double *d_lam, *d_L, *d_tail = nullptr;
        char *d_st = nullptr;
        if ((rc = dev_buf_t(pd, 0, (size_t)cnt * w, &d_lam))) return rc;
        if ((rc = dev_buf_t(pd, 1, (size_t)cnt * ow, &d_L))) return rc;
        if ((rc = dev_buf_t(pd, 2, (size_t)cnt * 4, &d_st))) return rc;
        if (tail_opt && (rc = dev_buf_t(pd, 3, (size_t)cnt * tl * w, &d_tail))) return rc;
        if (staged) {
            // chunk: at least 8 per device so that every worker has work, at most 64 Ki shifts (1 MiB of λ) per slot
            int64_t chunk = (cnt + 2 * kPipe - 1) / (2 * kPipe);
            if (chunk > (1 << 16)) chunk = 1 << 16;
            if (chunk < 1024) chunk = cnt < 1024 ? cnt : 1024;
            for (int t = 0; t < kPipe; t++)
                for (int k = 0; k < kStage; k++) {
                    if ((rc = host_stage_buf(&g_dev[pd].hin[t][k], &g_dev[pd].hin_cap[t][k], (size_t)chunk * sizeof(double) * w))) return rc;
                    if ((rc = host_stage_buf(&g_dev[pd].hout[t][k], &g_dev[pd].hout_cap[t][k], (size_t)chunk * (sizeof(double) * ow + 8)))) return rc;
                }
            jobs.push_back(TzStageJob{real_path, out_mode, l, u, branch_rank, pd, a, lambda, lo, cnt, chunk, d_lam, d_L, d_st, L11, status,
                                      stage_in, stage_out});
            continue;
        }
        // pinned caller buffers.  The call is bound by the H2D copy of λ (16 B/shift over PCIe), so the pipeline is built around
        // keeping that copy engine busy from the first microsecond to the last: ALL input chunks go back to back on one
        // stream; the kernels (two alternating streams) and the D2H copies (their own stream) hang off events.  Chunk sizes
        // shrink geometrically so that what is left after the last input byte has arrived — one small kernel and its
        // small D2H — is short.  With several devices each device's enqueue loop runs on its own pool thread (a single
        // thread needs ~250 µs to queue 8 devices' worth of copies and launches — longer than the work itself).
        direct.push_back([=]() -> int {
            ILA_CUDA_CHECK(cudaSetDevice(pd));
            DevState &S = g_dev[pd];
            cudaStream_t s_in = S.pipe[0], s_out = S.pipe[1];
            const int64_t min_chunk = g_chunk_min.load() > 0 ? g_chunk_min.load() : (1 << 16);
            const int shrink = g_chunk_div.load() > 1 ? g_chunk_div.load() : 3;
            int ci = 0, rc2 = 0;
            for (int64_t c0 = 0; c0 < cnt; ci++) {
                int64_t cc = (cnt - c0) / shrink;
                cc = (cc + 1023) & ~(int64_t)1023;           // whole grid rows of the usual power-of-two widths
                if (cc < min_chunk) cc = min_chunk;
                if (c0 + cc > cnt || cnt - (c0 + cc) < min_chunk / 2) cc = cnt - c0;
                if (ci >= kMaxChunkEv) cc = cnt - c0;
                cudaStream_t s_k = S.pipe[2 + (ci & 1)];
                ILA_CUDA_CHECK(cudaMemcpyAsync(d_lam + c0 * w, lambda + (lo + c0) * w, sizeof(double) * cc * w, cudaMemcpyHostToDevice, s_in));
                ILA_CUDA_CHECK(cudaEventRecord(S.cev[ci][0], s_in));
                ILA_CUDA_CHECK(cudaStreamWaitEvent(s_k, S.cev[ci][0], 0));
                if (out_mode == 1)
                    rc2 = launch_ql_toeplitz_l11re(l, u, a, branch_rank, d_lam + c0 * w, cc, d_L + c0,
                                                   status ? reinterpret_cast<uint8_t *>(d_st) + c0 : nullptr, s_k);
                else
                    rc2 = launch_ql_toeplitz(real_path, l, u, a, branch_rank, d_lam + c0 * w, cc, d_L + c0 * w,
                                             d_tail ? d_tail + c0 * tl * w : nullptr, reinterpret_cast<int32_t *>(d_st) + c0, s_k);
                if (rc2) return rc2;
                ILA_CUDA_CHECK(cudaEventRecord(S.cev[ci][1], s_k));
                ILA_CUDA_CHECK(cudaStreamWaitEvent(s_out, S.cev[ci][1], 0));
                ILA_CUDA_CHECK(cudaMemcpyAsync(L11 + (lo + c0) * ow, d_L + c0 * ow, sizeof(double) * cc * ow, cudaMemcpyDeviceToHost, s_out));
                if (status)
                    ILA_CUDA_CHECK(cudaMemcpyAsync(static_cast<char *>(status) + (lo + c0) * sb, d_st + c0 * sb, (size_t)cc * sb, cudaMemcpyDeviceToHost, s_out));
                if (tail_opt)
                    ILA_CUDA_CHECK(cudaMemcpyAsync(tail_opt + (lo + c0) * tl * w, d_tail + c0 * tl * w, sizeof(double) * cc * tl * w, cudaMemcpyDeviceToHost, s_out));
                c0 += cc;
            }
            return 0;
        });
    }
    if (!direct.empty()) {
        std::vector<int> rcs(direct.size(), 0);
        std::vector<std::string> errs(direct.size());
        if (direct.size() > 1) {
            if (!g_pool) g_pool = new WorkerPool();
            g_pool->ensure((int)direct.size() - 1);
            for (size_t j = 1; j < direct.size(); j++)
                g_pool->submit([&, j] {
                    rcs[j] = direct[j]();
                    if (rcs[j]) errs[j] = g_last_error;
                });
        }
        rcs[0] = direct[0]();
        if (direct.size() > 1) g_pool->wait();
        for (size_t j = 0; j < rcs.size(); j++)
            if (rcs[j]) {
                if (j) snprintf(g_last_error, sizeof(g_last_error), "%s", errs[j].c_str());
                return rcs[j];
            }
    }
    if (!jobs.empty()) {
        // workers: kPipe per device from the persistent pool; worker (job 0, stream 0) runs on the calling thread
        if (!g_pool) g_pool = new WorkerPool();
        g_pool->ensure((int)jobs.size() * kPipe - 1);
        std::vector<int> rcs(jobs.size() * kPipe, 0);
        std::vector<std::string> errs(jobs.size() * kPipe);
        for (size_t j = 0; j < jobs.size(); j++)
            for (int t = 0; t < kPipe; t++) {
                if (j == 0 && t == 0) continue;
                g_pool->submit([&, j, t] {
                    rcs[j * kPipe + t] = tz_stage_worker(jobs[j], t);
                    if (rcs[j * kPipe + t]) errs[j * kPipe + t] = g_last_error;   // thread-local text -> caller
                });
            }
        rcs[0] = tz_stage_worker(jobs[0], 0);
        g_pool->wait();
        for (size_t k = 0; k < rcs.size(); k++)
            if (rcs[k]) {
                if (k) snprintf(g_last_error, sizeof(g_last_error), "%s", errs[k].c_str());
                return rcs[k];
            }
    }
    for (int d = 0; d < G; d++) {
        const int pd = phys(d, G);
        if (!g_dev[pd].init) continue;
        ILA_CUDA_CHECK(cudaSetDevice(pd));
        for (int i = 0; i < kPipe; i++) ILA_CUDA_CHECK(cudaStreamSynchronize(g_dev[pd].pipe[i]));
    }
    guard.done = true;
    ILA_CUDA_CHECK(cudaSetDevice(g_base_dev));
    return 0;
}

int ila_ql_toeplitz_l11_c64(int l, int u, const ila_c64 *a, const ila_c64 *lambda, int64_t n, int branch_rank,
                            ila_c64 *L11, ila_c64 *tail_opt, int32_t *status, int ngpus)
{
    return ql_toeplitz_host(false, 0, l, u, (const double *)a, (const double *)lambda, n, branch_rank, (double *)L11,
                            (double *)tail_opt, status, ngpus);
}

int ila_ql_toeplitz_l11_f64(int l, int u, const double *a, const double *lambda, int64_t n, int branch_rank,
                            double *L11, double *tail_opt, int32_t *status, int ngpus)
{
    return ql_toeplitz_host(true, 0, l, u, a, lambda, n, branch_rank, L11, tail_opt, status, ngpus);
}

int ila_ql_toeplitz_l11re_c64(int l, int u, const ila_c64 *a, const ila_c64 *lambda, int64_t n, int branch_rank,
                              double *L11re, uint8_t *status8_opt, int ngpus)
{
    return ql_toeplitz_host(false, 1, l, u, (const double *)a, (const double *)lambda, n, branch_rank, L11re, nullptr,
                            status8_opt, ngpus);
}

int ila_ql_toeplitz_l11re_c64_dev(int l, int u, const ila_c64 *a, const ila_c64 *d_lambda, int64_t n, int branch_rank,
                                  double *d_L11re, uint8_t *d_status8_opt, void *stream)
{
    if (ila_device_count() <= 0) return set_error(ILA_ERR_NOGPU, "no CUDA device visible: libila_b200 has no CPU fallback");
    if (n < 0 || !a || (n > 0 && (!d_lambda || !d_L11re))) return set_error(ILA_ERR_ARG, "ql_toeplitz_l11re: null pointer / negative n");
    return launch_ql_toeplitz_l11re(l, u, (const double *)a, branch_rank, (const double *)d_lambda, n, d_L11re, d_status8_opt,
                                    (cudaStream_t)stream);
}

int ila_set_host_chunking(int64_t min_chunk_shifts, int divisor)
{
    if (min_chunk_shifts < 0 || divisor < 0) return set_error(ILA_ERR_ARG, "host chunking: arguments must be >= 0 (0 = default)");
    g_chunk_min.store(min_chunk_shifts);
    g_chunk_div.store(divisor);
    return 0;
}

int ila_set_host_copy_mode(int mode)
{
    if (mode < 0 || mode > 2) return set_error(ILA_ERR_ARG, "host copy mode must be 0 (auto), 1 (direct) or 2 (staged)");
    g_host_copy_mode.store(mode);
    return 0;
}

/* pinned host memory for callers that want the direct (zero-staging) path */
int ila_host_alloc(void **ptr, int64_t bytes)
{
    if (!ptr || bytes < 0) return set_error(ILA_ERR_ARG, "host_alloc: bad arguments");
    if (ila_device_count() <= 0) return set_error(ILA_ERR_NOGPU, "no CUDA device visible: libila_b200 has no CPU fallback");
    ILA_CUDA_CHECK(cudaHostAlloc(ptr, bytes > 0 ? (size_t)bytes : 8, cudaHostAllocPortable));
    return 0;
}
int ila_host_free(void *ptr)
{
    if (ptr) ILA_CUDA_CHECK(cudaFreeHost(ptr));
    return 0;
}
int ila_host_register(void *ptr, int64_t bytes)
{
    if (!ptr || bytes <= 0) return set_error(ILA_ERR_ARG, "host_register: bad arguments");
    if (ila_device_count() <= 0) return set_error(ILA_ERR_NOGPU, "no CUDA device visible: libila_b200 has no CPU fallback");
    ILA_CUDA_CHECK(cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable));
    return 0;
}
int ila_host_unregister(void *ptr)
{
    if (ptr) ILA_CUDA_CHECK(cudaHostUnregister(ptr));
    return 0;
}

int ila_ql_toeplitz_l11_c64_dev(int l, int u, const ila_c64 *a, const ila_c64 *d_lambda, int64_t n, int branch_rank,
                                ila_c64 *d_L11, ila_c64 *d_tail_opt, int32_t *d_status, void *stream)
{
    if (ila_device_count() <= 0) return set_error(ILA_ERR_NOGPU, "no CUDA device visible: libila_b200 has no CPU fallback");
    return launch_ql_toeplitz(false, l, u, (const double *)a, branch_rank, (const double *)d_lambda, n, (double *)d_L11,
                              (double *)d_tail_opt, d_status, (cudaStream_t)stream);
}

int ila_ql_toeplitz_l11_f64_dev(int l, int u, const double *a, const double *d_lambda, int64_t n, int branch_rank,
                                double *d_L11, double *d_tail_opt, int32_t *d_status, void *stream)
{
    if (ila_device_count() <= 0) return set_error(ILA_ERR_NOGPU, "no CUDA device visible: libila_b200 has no CPU fallback");
    return launch_ql_toeplitz(true, l, u, a, branch_rank, d_lambda, n, d_L11, d_tail_opt, d_status, (cudaStream_t)stream);
}

int ila_ql_toeplitz_absl11_grid_c64_dev(int l, int u, const ila_c64 *a, const double *d_x, int nx, const double *d_y, int64_t ny,
                                        int branch_rank, double *d_out, void *stream)
{
    if (ila_device_count() <= 0) return set_error(ILA_ERR_NOGPU, "no CUDA device visible: libila_b200 has no CPU fallback");
    return launch_ql_toeplitz_grid(l, u, (const double *)a, branch_rank, d_x, nx, d_y, ny, d_out, (cudaStream_t)stream);
}

int ila_ql_toeplitz_absl11_grid_c64(int l, int u, const ila_c64 *a, const double *x, int nx, const double *y, int64_t ny,
                                    int branch_rank, double *out, int ngpus)
{
    if (nx < 0 || ny < 0 || !a || ((int64_t)nx * ny > 0 && (!x || !y || !out))) return set_error(ILA_ERR_ARG, "ql_toeplitz grid: bad arguments");
    int G = 0;
    int rc = resolve_ngpus(ngpus, &G);
    if (rc) return rc;
    if ((int64_t)nx * ny == 0) return 0;
    std::lock_guard<std::mutex> lock(g_mutex);
    HostGuard guard(G);
    if (G > ny) G = (int)ny;
    const int64_t per = (ny + G - 1) / G;           // contiguous row slabs per device
    for (int d = 0; d < G; d++) {
        const int64_t lo = d * per, cnt = (lo + per <= ny ? per : ny - lo);
        if (cnt <= 0) continue;
        const int pd = phys(d, G);
        if ((rc = dev_init(pd))) return rc;
        double *d_x, *d_y, *d_o;
        if ((rc = dev_buf_t(pd, 0, (size_t)nx, &d_x))) return rc;
        if ((rc = dev_buf_t(pd, 1, (size_t)cnt, &d_y))) return rc;
        if ((rc = dev_buf_t(pd, 2, (size_t)cnt * nx, &d_o))) return rc;
        cudaStream_t s0 = g_dev[pd].pipe[0];
        ILA_CUDA_CHECK(cudaMemcpyAsync(d_x, x, sizeof(double) * nx, cudaMemcpyHostToDevice, s0));
        ILA_CUDA_CHECK(cudaMemcpyAsync(d_y, y + lo, sizeof(double) * cnt, cudaMemcpyHostToDevice, s0));
        ILA_CUDA_CHECK(cudaStreamSynchronize(s0));
        // row-chunked pipeline: the D2H of chunk c overlaps the kernel of chunk c+1
        const int64_t rchunk = cnt > 256 ? 128 : cnt;
        int ci = 0;
        for (int64_t r0 = 0; r0 < cnt; r0 += rchunk, ci++) {
            const int64_t rc_ = (r0 + rchunk <= cnt) ? rchunk : cnt - r0;
            cudaStream_t s = g_dev[pd].pipe[ci % kPipe];
            if ((rc = launch_ql_toeplitz_grid(l, u, (const double *)a, branch_rank, d_x, nx, d_y + r0, rc_, d_o + r0 * nx, s))) return rc;
            ILA_CUDA_CHECK(cudaMemcpyAsync(out + (lo + r0) * nx, d_o + r0 * nx, sizeof(double) * rc_ * nx, cudaMemcpyDeviceToHost, s));
        }
    }
    for (int d = 0; d < G; d++) {
        const int pd = phys(d, G);
        if (!g_dev[pd].init) continue;
        ILA_CUDA_CHECK(cudaSetDevice(pd));
        for (int i = 0; i < kPipe; i++) ILA_CUDA_CHECK(cudaStreamSynchronize(g_dev[pd].pipe[i]));
    }
    guard.done = true;
    ILA_CUDA_CHECK(cudaSetDevice(g_base_dev));
    return 0;
}

// ================================ K7: solve with the Toeplitz QL ========================================

static int ql_solve_args(int l, int u, int nb, int64_t n, int64_t nout)
{
    if (n < 0 || nout < 1 || nb < 1) return set_error(ILA_ERR_ARG, "ql_toeplitz_solve: bad sizes");
    if (u != 1 || l < 1 || l > 7) return set_error(ILA_ERR_UNSUPPORTED, "ql_toeplitz_solve: only u == 1, l in 1..7");
    return 0;
}

static int ql_solve_dev(bool real_path, int l, int u, const double *a, const double *d_lambda, int64_t n, int branch_rank,
                        int nb, const double *b, int64_t nout, double *d_work, double *d_x, int32_t *d_status, void *stream)
{
    if (ila_device_count() <= 0) return set_error(ILA_ERR_NOGPU, "no CUDA device visible: libila_b200 has no CPU fallback");
    int rc = ql_solve_args(l, u, nb, n, nout);
    if (rc) return rc;
    if (!a || !b || (n > 0 && (!d_lambda || !d_work || !d_x || !d_status))) return set_error(ILA_ERR_ARG, "ql_toeplitz_solve: null pointer");
    // one fused launch: the K3 kernel solves from its registers; d_work receives L[1,1] (n scalars)
    return launch_ql_toeplitz_solve_fused(real_path, l, u, a, branch_rank, d_lambda, n, nb, b, nout, d_work, d_x, d_status,
                                          (cudaStream_t)stream);
}

static int ql_solve_host(bool real_path, int l, int u, const double *a, const double *lambda, int64_t n, int branch_rank,
                         int nb, const double *b, int64_t nout, double *x, int32_t *status, int ngpus, int mode = 0,
                         int64_t *nstop = nullptr)
{
    int rc = ql_solve_args(l, u, nb, n, nout);
    if (rc) return rc;
    if (!a || !b || (n > 0 && (!lambda || !x || !status))) return set_error(ILA_ERR_ARG, "ql_toeplitz_solve: null pointer");
    int G = 0;
    if ((rc = resolve_ngpus(ngpus, &G))) return rc;
    if (n == 0) return 0;
    std::lock_guard<std::mutex> lock(g_mutex);
    HostGuard guard(G);
    const int w = real_path ? 1 : 2;
    if (G > n) G = (int)n;
    const int64_t per = (n + G - 1) / G;
    for (int d = 0; d < G; d++) {
        const int64_t lo = d * per, cnt = (lo + per <= n ? per : n - lo);
        if (cnt <= 0) continue;
        const int pd = phys(d, G);
        if ((rc = dev_init(pd))) return rc;
        double *d_lam, *d_L, *d_x;
        int32_t *d_st;
        if ((rc = dev_buf_t(pd, 0, (size_t)cnt * w, &d_lam))) return rc;
        if ((rc = dev_buf_t(pd, 1, (size_t)cnt * w, &d_L))) return rc;
        if ((rc = dev_buf_t(pd, 2, (size_t)cnt, &d_st))) return rc;
        if ((rc = dev_buf_t(pd, 4, (size_t)cnt * nout * w, &d_x))) return rc;
        int64_t *d_ns = nullptr;
        if (nstop && (rc = dev_buf_t(pd, 5, (size_t)cnt, &d_ns))) return rc;
        cudaStream_t s = g_dev[pd].stream;
        ILA_CUDA_CHECK(cudaMemcpyAsync(d_lam, lambda + lo * w, sizeof(double) * cnt * w, cudaMemcpyHostToDevice, s));
        if ((rc = launch_ql_toeplitz_solve_fused(real_path, l, u, a, branch_rank, d_lam, cnt, nb, b, nout, d_L, d_x, d_st, s, mode, d_ns))) return rc;
        if (nstop) ILA_CUDA_CHECK(cudaMemcpyAsync(nstop + lo, d_ns, sizeof(int64_t) * cnt, cudaMemcpyDeviceToHost, s));
        ILA_CUDA_CHECK(cudaMemcpyAsync(status + lo, d_st, sizeof(int32_t) * cnt, cudaMemcpyDeviceToHost, s));
        // x is column-major n x nout on the host, cnt x nout on the device: one strided copy
        ILA_CUDA_CHECK(cudaMemcpy2DAsync(x + lo * w, sizeof(double) * n * w, d_x, sizeof(double) * cnt * w,
                                         sizeof(double) * cnt * w, (size_t)nout, cudaMemcpyDeviceToHost, s));
    }
    for (int d = 0; d < G; d++) {
        const int pd = phys(d, G);
        if (!g_dev[pd].init) continue;
        ILA_CUDA_CHECK(cudaSetDevice(pd));
        ILA_CUDA_CHECK(cudaStreamSynchronize(g_dev[pd].stream));
    }
    guard.done = true;
    ILA_CUDA_CHECK(cudaSetDevice(g_base_dev));
    return 0;
}

int ila_ql_toeplitz_solve_c64(int l, int u, const ila_c64 *a, const ila_c64 *lambda, int64_t n, int branch_rank, int nb,
                              const ila_c64 *b, int64_t nout, ila_c64 *x, int32_t *status, int ngpus)
{
    return ql_solve_host(false, l, u, (const double *)a, (const double *)lambda, n, branch_rank, nb, (const double *)b, nout,
                         (double *)x, status, ngpus);
}
int ila_ql_toeplitz_solve_f64(int l, int u, const double *a, const double *lambda, int64_t n, int branch_rank, int nb,
                              const double *b, int64_t nout, double *x, int32_t *status, int ngpus)
{
    return ql_solve_host(true, l, u, a, lambda, n, branch_rank, nb, b, nout, x, status, ngpus);
}
int ila_ql_toeplitz_applyq_c64(int l, int u, const ila_c64 *a, const ila_c64 *lambda, int64_t n, int branch_rank, int adjoint,
                               int nb, const ila_c64 *b, int64_t nout, ila_c64 *y, int64_t *nstop_opt, int32_t *status, int ngpus)
{
    return ql_solve_host(false, l, u, (const double *)a, (const double *)lambda, n, branch_rank, nb, (const double *)b, nout,
                         (double *)y, status, ngpus, adjoint ? 1 : 2, adjoint ? nullptr : nstop_opt);
}
int ila_ql_toeplitz_applyq_f64(int l, int u, const double *a, const double *lambda, int64_t n, int branch_rank, int adjoint,
                               int nb, const double *b, int64_t nout, double *y, int64_t *nstop_opt, int32_t *status, int ngpus)
{
    return ql_solve_host(true, l, u, a, lambda, n, branch_rank, nb, b, nout, y, status, ngpus, adjoint ? 1 : 2,
                         adjoint ? nullptr : nstop_opt);
}
int ila_ql_toeplitz_solve_c64_dev(int l, int u, const ila_c64 *a, const ila_c64 *d_lambda, int64_t n, int branch_rank, int nb,
                                  const ila_c64 *b, int64_t nout, ila_c64 *d_work, ila_c64 *d_x, int32_t *d_status, void *stream)
{
    return ql_solve_dev(false, l, u, (const double *)a, (const double *)d_lambda, n, branch_rank, nb, (const double *)b, nout,
                        (double *)d_work, (double *)d_x, d_status, stream);
}
int ila_ql_toeplitz_solve_f64_dev(int l, int u, const double *a, const double *d_lambda, int64_t n, int branch_rank, int nb,
                                  const double *b, int64_t nout, double *d_work, double *d_x, int32_t *d_status, void *stream)
{
    return ql_solve_dev(true, l, u, a, d_lambda, n, branch_rank, nb, b, nout, d_work, d_x, d_status, stream);
}

// ================================ K4: perturbed-Toeplitz head ==========================================

static int ql_pert_host(bool real_path, int l, int u, int p, const double *head, const double *a, const double *lambda,
                        int64_t n, int branch_rank, double *L11, double *head_factors_opt, int32_t *status, int ngpus,
                        int nb = 0, const double *b = nullptr, int64_t nout = 0, double *x = nullptr)
{
    if (n < 0 || !a || !head || (n > 0 && (!lambda || !status || (!L11 && !x)))) return set_error(ILA_ERR_ARG, "ql_perttoeplitz: bad arguments");
    if (x && (!b || nb < 1 || nout < 1)) return set_error(ILA_ERR_ARG, "ql_perttoeplitz_solve: bad right-hand side / nout");
    int G = 0;
    int rc = resolve_ngpus(ngpus, &G);
    if (rc) return rc;
    if (n == 0) return 0;
    std::lock_guard<std::mutex> lock(g_mutex);
    HostGuard guard(G);
    const int w = real_path ? 1 : 2;
    const int m = l + 2, tl = 2 * m + 2;
    if (G > n) G = (int)n;
    const int64_t per = (n + G - 1) / G;
    for (int d = 0; d < G; d++) {
        const int64_t lo = d * per, cnt = (lo + per <= n ? per : n - lo);
        if (cnt <= 0) continue;
        const int pd = phys(d, G);
        if ((rc = dev_init(pd))) return rc;
        double *d_lam, *d_L, *d_tail, *d_hf = nullptr, *d_x = nullptr;
        int32_t *d_st;
        if ((rc = dev_buf_t(pd, 0, (size_t)cnt * w, &d_lam))) return rc;
        if ((rc = dev_buf_t(pd, 1, (size_t)cnt * w, &d_L))) return rc;
        if ((rc = dev_buf_t(pd, 2, (size_t)cnt, &d_st))) return rc;
        if ((rc = dev_buf_t(pd, 3, (size_t)cnt * tl * w, &d_tail))) return rc;
        if (head_factors_opt && (rc = dev_buf_t(pd, 23, (size_t)cnt * p * p * w, &d_hf))) return rc;
        if (x && (rc = dev_buf_t(pd, 4, (size_t)cnt * nout * w, &d_x))) return rc;
        cudaStream_t s = g_dev[pd].stream;
        ILA_CUDA_CHECK(cudaMemcpyAsync(d_lam, lambda + lo * w, sizeof(double) * cnt * w, cudaMemcpyHostToDevice, s));
        if ((rc = launch_ql_toeplitz(real_path, l, u, a, branch_rank, d_lam, cnt, d_L, d_tail, d_st, s))) return rc;
        if ((rc = launch_ql_pert_head_ex(real_path, l, p, a, head, d_lam, cnt, d_tail, d_st, d_L, d_hf, nb, b, nout, d_x, s))) return rc;
        if (L11) ILA_CUDA_CHECK(cudaMemcpyAsync(L11 + lo * w, d_L, sizeof(double) * cnt * w, cudaMemcpyDeviceToHost, s));
        if (x)
            ILA_CUDA_CHECK(cudaMemcpy2DAsync(x + lo * w, sizeof(double) * n * w, d_x, sizeof(double) * cnt * w,
                                             sizeof(double) * cnt * w, (size_t)nout, cudaMemcpyDeviceToHost, s));
        ILA_CUDA_CHECK(cudaMemcpyAsync(status + lo, d_st, sizeof(int32_t) * cnt, cudaMemcpyDeviceToHost, s));
        if (head_factors_opt)
            ILA_CUDA_CHECK(cudaMemcpyAsync(head_factors_opt + lo * p * p * w, d_hf, sizeof(double) * cnt * p * p * w, cudaMemcpyDeviceToHost, s));
    }
    for (int d = 0; d < G; d++) {
        const int pd = phys(d, G);
        if (!g_dev[pd].init) continue;
        ILA_CUDA_CHECK(cudaSetDevice(pd));
        ILA_CUDA_CHECK(cudaStreamSynchronize(g_dev[pd].stream));
    }
    guard.done = true;
    ILA_CUDA_CHECK(cudaSetDevice(g_base_dev));
    return 0;
}

int ila_ql_perttoeplitz_l11_c64(int l, int u, int p, const ila_c64 *head, const ila_c64 *a, const ila_c64 *lambda, int64_t n,
                                int branch_rank, ila_c64 *L11, ila_c64 *head_factors_opt, int32_t *status, int ngpus)
{
    return ql_pert_host(false, l, u, p, (const double *)head, (const double *)a, (const double *)lambda, n, branch_rank,
                        (double *)L11, (double *)head_factors_opt, status, ngpus);
}

int ila_ql_perttoeplitz_l11_f64(int l, int u, int p, const double *head, const double *a, const double *lambda, int64_t n,
                                int branch_rank, double *L11, double *head_factors_opt, int32_t *status, int ngpus)
{
    return ql_pert_host(true, l, u, p, head, a, lambda, n, branch_rank, L11, head_factors_opt, status, ngpus);
}

int ila_ql_perttoeplitz_solve_c64(int l, int u, int p, const ila_c64 *head, const ila_c64 *a, const ila_c64 *lambda, int64_t n,
                                  int branch_rank, int nb, const ila_c64 *b, int64_t nout, ila_c64 *x, int32_t *status, int ngpus)
{
    return ql_pert_host(false, l, u, p, (const double *)head, (const double *)a, (const double *)lambda, n, branch_rank,
                        nullptr, nullptr, status, ngpus, nb, (const double *)b, nout, (double *)x);
}

int ila_ql_perttoeplitz_solve_f64(int l, int u, int p, const double *head, const double *a, const double *lambda, int64_t n,
                                  int branch_rank, int nb, const double *b, int64_t nout, double *x, int32_t *status, int ngpus)
{
    return ql_pert_host(true, l, u, p, head, a, lambda, n, branch_rank, nullptr, nullptr, status, ngpus, nb, b, nout, x);
}

// ================================ K8: adaptive reverse Cholesky (symmetric tridiagonal) =================

int ila_revchol_tridiag_f64(int64_t batch, const double *ga, const double *gb, const double *shift, int hp,
                            const double *a_head, const double *b_head, int64_t n, double *la, double *lb, int64_t *N_used,
                            int32_t *status, int ngpus)
{
    if (batch < 0 || n < 1 || hp < 0 || !ga || !gb || (hp > 0 && (!a_head || !b_head)) ||
        (batch > 0 && (!la || !lb || !N_used || !status)))
        return set_error(ILA_ERR_ARG, "revchol_tridiag: bad arguments");
    if (n > (1 << 19)) return set_error(ILA_ERR_UNSUPPORTED, "revchol_tridiag: n must not exceed 2^19 (N starts at 4n, MAX_TRIDIAG_CHOL_N = 2^21 - 1)");
    int G = 0;
    int rc = resolve_ngpus(ngpus, &G);
    if (rc) return rc;
    if (batch == 0) return 0;
    std::lock_guard<std::mutex> lock(g_mutex);
    HostGuard guard(G);
    if (G > batch) G = (int)batch;
    const int64_t per = (batch + G - 1) / G;
    for (int d = 0; d < G; d++) {
        const int64_t lo = d * per, cnt = (lo + per <= batch ? per : batch - lo);
        if (cnt <= 0) continue;
        const int pd = phys(d, G);
        if ((rc = dev_init(pd))) return rc;
        double *d_ga, *d_gb, *d_sh = nullptr, *d_ah = nullptr, *d_bh = nullptr, *d_la, *d_lb;
        int64_t *d_N;
        int32_t *d_st;
        if ((rc = dev_buf_t(pd, 0, (size_t)cnt * 5, &d_ga))) return rc;
        if ((rc = dev_buf_t(pd, 1, (size_t)cnt * 5, &d_gb))) return rc;
        if (shift && (rc = dev_buf_t(pd, 2, (size_t)cnt, &d_sh))) return rc;
        if (hp > 0 && (rc = dev_buf_t(pd, 3, (size_t)hp, &d_ah))) return rc;
        if (hp > 0 && (rc = dev_buf_t(pd, 4, (size_t)hp, &d_bh))) return rc;
        if ((rc = dev_buf_t(pd, 5, (size_t)cnt * n, &d_la))) return rc;
        if ((rc = dev_buf_t(pd, 6, (size_t)cnt * n, &d_lb))) return rc;
        if ((rc = dev_buf_t(pd, 7, (size_t)cnt, &d_N))) return rc;
        if ((rc = dev_buf_t(pd, 8, (size_t)cnt, &d_st))) return rc;
        cudaStream_t s = g_dev[pd].stream;
        ILA_CUDA_CHECK(cudaMemcpyAsync(d_ga, ga + lo * 5, sizeof(double) * cnt * 5, cudaMemcpyHostToDevice, s));
        ILA_CUDA_CHECK(cudaMemcpyAsync(d_gb, gb + lo * 5, sizeof(double) * cnt * 5, cudaMemcpyHostToDevice, s));
        if (shift) ILA_CUDA_CHECK(cudaMemcpyAsync(d_sh, shift + lo, sizeof(double) * cnt, cudaMemcpyHostToDevice, s));
        if (hp > 0) {
            ILA_CUDA_CHECK(cudaMemcpyAsync(d_ah, a_head, sizeof(double) * hp, cudaMemcpyHostToDevice, s));
            ILA_CUDA_CHECK(cudaMemcpyAsync(d_bh, b_head, sizeof(double) * hp, cudaMemcpyHostToDevice, s));
        }
        RevCholArgs a{cnt, n, d_ga, d_gb, d_sh, hp, d_ah, d_bh, d_la, d_lb, d_N, d_st};
        if ((rc = launch_revchol_tridiag(a, s))) return rc;
        ILA_CUDA_CHECK(cudaMemcpyAsync(la + lo * n, d_la, sizeof(double) * cnt * n, cudaMemcpyDeviceToHost, s));
        ILA_CUDA_CHECK(cudaMemcpyAsync(lb + lo * n, d_lb, sizeof(double) * cnt * n, cudaMemcpyDeviceToHost, s));
        ILA_CUDA_CHECK(cudaMemcpyAsync(N_used + lo, d_N, sizeof(int64_t) * cnt, cudaMemcpyDeviceToHost, s));
        ILA_CUDA_CHECK(cudaMemcpyAsync(status + lo, d_st, sizeof(int32_t) * cnt, cudaMemcpyDeviceToHost, s));
    }
    for (int d = 0; d < G; d++) {
        const int pd = phys(d, G);
        if (!g_dev[pd].init) continue;
        ILA_CUDA_CHECK(cudaSetDevice(pd));
        ILA_CUDA_CHECK(cudaStreamSynchronize(g_dev[pd].stream));
    }
    guard.done = true;
    ILA_CUDA_CHECK(cudaSetDevice(g_base_dev));
    return 0;
}

// ================================ K10: tridiagonal / bidiagonal conjugation ==============================

int ila_triconj_bands_f64_dev(int nbands, int64_t batch, int64_t n, const double *d_U, int64_t uld, const double *d_X,
                              int64_t xld, int x_shared, double *d_Y, double *d_UX_opt, void *stream)
{
    if (ila_device_count() <= 0) return set_error(ILA_ERR_NOGPU, "no CUDA device visible: libila_b200 has no CPU fallback");
    if (batch < 0 || n < 1 || uld < n + 1 || xld < n + 1 || !d_U || !d_X || !d_Y) return set_error(ILA_ERR_ARG, "triconj_bands: bad arguments (bands need n+1 entries)");
    TriConjArgs a{};
    a.batch = batch; a.n = n; a.nbands = nbands; a.U = d_U; a.uld = uld; a.X = d_X; a.xld = xld; a.xstride = x_shared ? 0 : 3 * xld;
    a.Y = d_Y; a.UX = d_UX_opt;
    return launch_triconj(a, (cudaStream_t)stream);
}

int ila_triconj_records_f64_dev(int nbands, int nch, int64_t batch, int64_t n, const int64_t *d_gbase, const void *d_work,
                                const int32_t *d_status, const double *d_X, int64_t xld, int x_shared, double *d_Y, void *stream)
{
    if (ila_device_count() <= 0) return set_error(ILA_ERR_NOGPU, "no CUDA device visible: libila_b200 has no CPU fallback");
    if (batch < 0 || n < 1 || nch < 1 || 2 * nch < nbands || xld < n + 1 || !d_gbase || !d_work || !d_X || !d_Y)
        return set_error(ILA_ERR_ARG, "triconj_records: bad arguments");
    TriConjArgs a{};
    a.batch = batch; a.n = n; a.nbands = nbands; a.gbase = d_gbase; a.work = (const double2 *)d_work; a.nch = nch; a.status = d_status;
    a.X = d_X; a.xld = xld; a.xstride = x_shared ? 0 : 3 * xld; a.Y = d_Y;
    return launch_triconj(a, (cudaStream_t)stream);
}

int ila_triconj_bands_f64(int nbands, int64_t batch, int64_t n, const double *U, int64_t uld, const double *X, int64_t xld,
                          int x_shared, double *Y, double *UX_opt, int ngpus)
{
    if (batch < 0 || n < 1 || uld < n + 1 || xld < n + 1 || !U || !X || (batch > 0 && !Y)) return set_error(ILA_ERR_ARG, "triconj_bands: bad arguments (bands need n+1 entries)");
    if (nbands != 2 && nbands != 3) return set_error(ILA_ERR_UNSUPPORTED, "triconj_bands: nbands must be 2 or 3");
    int G = 0;
    int rc = resolve_ngpus(ngpus, &G);
    if (rc) return rc;
    if (batch == 0) return 0;
    std::lock_guard<std::mutex> lock(g_mutex);
    HostGuard guard(G);
    if (G > batch) G = (int)batch;
    const int64_t per = (batch + G - 1) / G;
    for (int d = 0; d < G; d++) {
        const int64_t lo = d * per, cnt = (lo + per <= batch ? per : batch - lo);
        if (cnt <= 0) continue;
        const int pd = phys(d, G);
        if ((rc = dev_init(pd))) return rc;
        double *d_U, *d_X, *d_Y, *d_UX = nullptr;
        const size_t xcount = (size_t)(x_shared ? 1 : cnt) * 3 * xld;
        if ((rc = dev_buf_t(pd, 0, (size_t)cnt * nbands * uld, &d_U))) return rc;
        if ((rc = dev_buf_t(pd, 1, xcount, &d_X))) return rc;
        if ((rc = dev_buf_t(pd, 2, (size_t)cnt * 3 * n, &d_Y))) return rc;
        if (UX_opt && (rc = dev_buf_t(pd, 3, (size_t)cnt * 3 * n, &d_UX))) return rc;
        cudaStream_t s = g_dev[pd].stream;
        ILA_CUDA_CHECK(cudaMemcpyAsync(d_U, U + lo * nbands * uld, sizeof(double) * cnt * nbands * uld, cudaMemcpyHostToDevice, s));
        ILA_CUDA_CHECK(cudaMemcpyAsync(d_X, X + (x_shared ? 0 : lo * 3 * xld), sizeof(double) * xcount, cudaMemcpyHostToDevice, s));
        if ((rc = ila_triconj_bands_f64_dev(nbands, cnt, n, d_U, uld, d_X, xld, x_shared, d_Y, d_UX, s))) return rc;
        ILA_CUDA_CHECK(cudaMemcpyAsync(Y + lo * 3 * n, d_Y, sizeof(double) * cnt * 3 * n, cudaMemcpyDeviceToHost, s));
        if (UX_opt) ILA_CUDA_CHECK(cudaMemcpyAsync(UX_opt + lo * 3 * n, d_UX, sizeof(double) * cnt * 3 * n, cudaMemcpyDeviceToHost, s));
    }
    for (int d = 0; d < G; d++) {
        const int pd = phys(d, G);
        if (!g_dev[pd].init) continue;
        ILA_CUDA_CHECK(cudaSetDevice(pd));
        ILA_CUDA_CHECK(cudaStreamSynchronize(g_dev[pd].stream));
    }
    guard.done = true;
    ILA_CUDA_CHECK(cudaSetDevice(g_base_dev));
    return 0;
}

int ila_biconj_bands_f64(int upper, int64_t batch, int64_t n, const double *U, const double *Cm, int64_t ld, double *dv,
                         double *ev, int ngpus)
{
    if (batch < 0 || n < 1 || ld < n + 1 || !U || !Cm || (batch > 0 && (!dv || !ev))) return set_error(ILA_ERR_ARG, "biconj_bands: bad arguments (bands need n+1 entries)");
    int G = 0;
    int rc = resolve_ngpus(ngpus, &G);
    if (rc) return rc;
    if (batch == 0) return 0;
    std::lock_guard<std::mutex> lock(g_mutex);
    HostGuard guard(G);
    if (G > batch) G = (int)batch;
    const int64_t per = (batch + G - 1) / G;
    for (int d = 0; d < G; d++) {
        const int64_t lo = d * per, cnt = (lo + per <= batch ? per : batch - lo);
        if (cnt <= 0) continue;
        const int pd = phys(d, G);
        if ((rc = dev_init(pd))) return rc;
        double *d_U, *d_C, *d_dv, *d_ev;
        if ((rc = dev_buf_t(pd, 0, (size_t)cnt * 3 * ld, &d_U))) return rc;
        if ((rc = dev_buf_t(pd, 1, (size_t)cnt * 3 * ld, &d_C))) return rc;
        if ((rc = dev_buf_t(pd, 2, (size_t)cnt * n, &d_dv))) return rc;
        if ((rc = dev_buf_t(pd, 3, (size_t)cnt * n, &d_ev))) return rc;
        cudaStream_t s = g_dev[pd].stream;
        ILA_CUDA_CHECK(cudaMemcpyAsync(d_U, U + lo * 3 * ld, sizeof(double) * cnt * 3 * ld, cudaMemcpyHostToDevice, s));
        ILA_CUDA_CHECK(cudaMemcpyAsync(d_C, Cm + lo * 3 * ld, sizeof(double) * cnt * 3 * ld, cudaMemcpyHostToDevice, s));
        BiConjArgs a{cnt, n, ld, upper ? 1 : 0, d_U, d_C, d_dv, d_ev};
        if ((rc = launch_biconj(a, s))) return rc;
        ILA_CUDA_CHECK(cudaMemcpyAsync(dv + lo * n, d_dv, sizeof(double) * cnt * n, cudaMemcpyDeviceToHost, s));
        ILA_CUDA_CHECK(cudaMemcpyAsync(ev + lo * n, d_ev, sizeof(double) * cnt * n, cudaMemcpyDeviceToHost, s));
    }
    for (int d = 0; d < G; d++) {
        const int pd = phys(d, G);
        if (!g_dev[pd].init) continue;
        ILA_CUDA_CHECK(cudaSetDevice(pd));
        ILA_CUDA_CHECK(cudaStreamSynchronize(g_dev[pd].stream));
    }
    guard.done = true;
    ILA_CUDA_CHECK(cudaSetDevice(g_base_dev));
    return 0;
}

// ================================ K11: adaptive block-banded QR solve ====================================

int ila_qr_blockbanded_set_mode(int mode)
{
    if (mode != 0 && mode != 1) return set_error(ILA_ERR_ARG, "qr_blockbanded mode must be 0 (blocked reflectors) or 1 (one reflector per pass)");
    qr_blockbanded_set_mode(mode);
    return 0;
}

int ila_qr_blockbanded_work_doubles(int maxblocks, int64_t *per_operator)
{
    if (maxblocks < 3 || maxblocks > 64 || !per_operator) return set_error(ILA_ERR_ARG, "qr_blockbanded_work_doubles: bad arguments");
    *per_operator = (int64_t)(maxblocks - 1) * maxblocks * (maxblocks + 1);
    return 0;
}

int ila_qr_blockbanded_solve_f64_dev(int64_t batch, const double *d_alpha, const double *d_beta, const double *d_lam,
                                     const double *d_T1, const double *d_T2, int64_t tld, int nb, const double *d_b, double tol,
                                     int colgrowth, int maxblocks, double *d_work, double *d_qtb, double *d_x, int64_t xld,
                                     int64_t *d_ncols, int32_t *d_nblocks, int32_t *d_status, void *stream)
{
    if (ila_device_count() <= 0) return set_error(ILA_ERR_NOGPU, "no CUDA device visible: libila_b200 has no CPU fallback");
    if (batch < 0 || nb < 1 || colgrowth < 1 || !d_work || !d_qtb || !d_x) return set_error(ILA_ERR_ARG, "qr_blockbanded_solve_dev: bad arguments");
    if (maxblocks < 3 || maxblocks > 64) return set_error(ILA_ERR_UNSUPPORTED, "qr_blockbanded_solve: maxblocks must be in 3..64 (the panel lives in shared memory)");
    if (tld < maxblocks + 1 || xld < (int64_t)maxblocks * (maxblocks + 1) / 2) return set_error(ILA_ERR_ARG, "qr_blockbanded_solve_dev: tld / xld too small");
    BlockQrArgs a{};
    a.batch = batch; a.alpha = d_alpha; a.beta = d_beta; a.lam = d_lam; a.T1 = d_T1; a.T2 = d_T2; a.tld = tld; a.b = d_b; a.nb = nb;
    a.tol = tol; a.colgrowth = colgrowth; a.maxblocks = maxblocks; a.work = d_work;
    a.work_stride = (int64_t)(maxblocks - 1) * maxblocks * (maxblocks + 1); a.qtb = d_qtb;
    a.x = d_x; a.xld = xld; a.ncols = d_ncols; a.nblocks = d_nblocks; a.status = d_status;
    return launch_qr_blockbanded(a, (cudaStream_t)stream);
}

int ila_qr_blockbanded_solve_f64(int64_t batch, const double *alpha, const double *beta, const double *lam, const double *T1,
                                 const double *T2, int64_t tld, int nb, const double *b, double tol, int colgrowth,
                                 int maxblocks, double *x, int64_t xld, int64_t *ncols, int32_t *nblocks, double *qtb_opt,
                                 int32_t *status, int ngpus)
{
    if (batch < 0 || !alpha || !beta || !lam || !T1 || !T2 || nb < 1 || !b || colgrowth < 1 ||
        (batch > 0 && (!x || !ncols || !nblocks || !status)))
        return set_error(ILA_ERR_ARG, "qr_blockbanded_solve: bad arguments");
    if (maxblocks < 3 || maxblocks > 64) return set_error(ILA_ERR_UNSUPPORTED, "qr_blockbanded_solve: maxblocks must be in 3..64 (the panel lives in shared memory)");
    const int64_t nmax = (int64_t)maxblocks * (maxblocks + 1) / 2;
    if (tld < maxblocks + 1 || xld < nmax || nb > nmax) return set_error(ILA_ERR_ARG, "qr_blockbanded_solve: need tld >= maxblocks+1, xld >= maxblocks(maxblocks+1)/2 >= nb");
    int G = 0;
    int rc = resolve_ngpus(ngpus, &G);
    if (rc) return rc;
    if (batch == 0) return 0;
    std::lock_guard<std::mutex> lock(g_mutex);
    HostGuard guard(G);
    if (G > batch) G = (int)batch;
    const int64_t per = (batch + G - 1) / G;
    const int64_t wstride = (int64_t)(maxblocks - 1) * maxblocks * (maxblocks + 1);
    for (int d = 0; d < G; d++) {
        const int64_t lo = d * per, cnt = (lo + per <= batch ? per : batch - lo);
        if (cnt <= 0) continue;
        const int pd = phys(d, G);
        if ((rc = dev_init(pd))) return rc;
        cudaStream_t s = g_dev[pd].stream;
        double *d_al, *d_be, *d_la, *d_T1, *d_T2, *d_b, *d_work, *d_qtb, *d_x;
        int64_t *d_ncols;
        int32_t *d_nblk, *d_st;
        if ((rc = dev_buf_t(pd, 0, (size_t)cnt, &d_al))) return rc;
        if ((rc = dev_buf_t(pd, 1, (size_t)cnt, &d_be))) return rc;
        if ((rc = dev_buf_t(pd, 2, (size_t)cnt, &d_la))) return rc;
        if ((rc = dev_buf_t(pd, 3, (size_t)3 * tld, &d_T1))) return rc;
        if ((rc = dev_buf_t(pd, 4, (size_t)3 * tld, &d_T2))) return rc;
        if ((rc = dev_buf_t(pd, 5, (size_t)cnt * nb, &d_b))) return rc;
        if ((rc = dev_buf_t(pd, 16, (size_t)cnt * wstride, &d_work))) return rc;
        if ((rc = dev_buf_t(pd, 6, (size_t)cnt * xld, &d_qtb))) return rc;
        if ((rc = dev_buf_t(pd, 7, (size_t)cnt * xld, &d_x))) return rc;
        if ((rc = dev_buf_t(pd, 8, (size_t)cnt, &d_ncols))) return rc;
        if ((rc = dev_buf_t(pd, 9, (size_t)cnt, &d_nblk))) return rc;
        if ((rc = dev_buf_t(pd, 10, (size_t)cnt, &d_st))) return rc;
        ILA_CUDA_CHECK(cudaMemcpyAsync(d_al, alpha + lo, sizeof(double) * cnt, cudaMemcpyHostToDevice, s));
        ILA_CUDA_CHECK(cudaMemcpyAsync(d_be, beta + lo, sizeof(double) * cnt, cudaMemcpyHostToDevice, s));
        ILA_CUDA_CHECK(cudaMemcpyAsync(d_la, lam + lo, sizeof(double) * cnt, cudaMemcpyHostToDevice, s));
        ILA_CUDA_CHECK(cudaMemcpyAsync(d_T1, T1, sizeof(double) * 3 * tld, cudaMemcpyHostToDevice, s));
        ILA_CUDA_CHECK(cudaMemcpyAsync(d_T2, T2, sizeof(double) * 3 * tld, cudaMemcpyHostToDevice, s));
        ILA_CUDA_CHECK(cudaMemcpyAsync(d_b, b + lo * nb, sizeof(double) * cnt * nb, cudaMemcpyHostToDevice, s));
        ILA_CUDA_CHECK(cudaMemsetAsync(d_qtb, 0, sizeof(double) * cnt * xld, s));
        BlockQrArgs a{};
        a.batch = cnt; a.alpha = d_al; a.beta = d_be; a.lam = d_la; a.T1 = d_T1; a.T2 = d_T2; a.tld = tld; a.b = d_b; a.nb = nb;
        a.tol = tol; a.colgrowth = colgrowth; a.maxblocks = maxblocks; a.work = d_work; a.work_stride = wstride; a.qtb = d_qtb;
        a.x = d_x; a.xld = xld; a.ncols = d_ncols; a.nblocks = d_nblk; a.status = d_st;
        if ((rc = launch_qr_blockbanded(a, s))) return rc;
        ILA_CUDA_CHECK(cudaMemcpyAsync(x + lo * xld, d_x, sizeof(double) * cnt * xld, cudaMemcpyDeviceToHost, s));
        if (qtb_opt) ILA_CUDA_CHECK(cudaMemcpyAsync(qtb_opt + lo * xld, d_qtb, sizeof(double) * cnt * xld, cudaMemcpyDeviceToHost, s));
        ILA_CUDA_CHECK(cudaMemcpyAsync(ncols + lo, d_ncols, sizeof(int64_t) * cnt, cudaMemcpyDeviceToHost, s));
        ILA_CUDA_CHECK(cudaMemcpyAsync(nblocks + lo, d_nblk, sizeof(int32_t) * cnt, cudaMemcpyDeviceToHost, s));
        ILA_CUDA_CHECK(cudaMemcpyAsync(status + lo, d_st, sizeof(int32_t) * cnt, cudaMemcpyDeviceToHost, s));
    }
    for (int d = 0; d < G; d++) {
        const int pd = phys(d, G);
        if (!g_dev[pd].init) continue;
        ILA_CUDA_CHECK(cudaSetDevice(pd));
        ILA_CUDA_CHECK(cudaStreamSynchronize(g_dev[pd].stream));
    }
    guard.done = true;
    ILA_CUDA_CHECK(cudaSetDevice(g_base_dev));
    return 0;
}

// ================================ K12: adaptive almost-banded QR solve ===================================

int ila_qr_almostbanded_solve_f64(int lD, int uD, int r, int64_t batch, const double *dg, const double *bg, const double *shift,
                                  int nb, const double *b, double tol, int colgrowth, int64_t ncap, double *x, int64_t *ncols,
                                  int64_t *nconv, int32_t *status, int ngpus)
{
    if (batch < 0 || !dg || !bg || nb < 1 || !b || colgrowth < 1 || ncap < 1 || (batch > 0 && (!x || !ncols || !nconv || !status)))
        return set_error(ILA_ERR_ARG, "qr_almostbanded_solve: bad arguments");
    if (r < 1 || r > 2 || lD < 0 || uD < 0 || lD + r > 4 || lD + uD + 1 > 12 || nb > ncap)
        return set_error(ILA_ERR_UNSUPPORTED, "qr_almostbanded_solve: need 1 <= r <= 2, lD + r <= 4, lD + uD + 1 <= 12, nb <= ncap");
    int G = 0;
    int rc = resolve_ngpus(ngpus, &G);
    if (rc) return rc;
    if (batch == 0) return 0;
    std::lock_guard<std::mutex> lock(g_mutex);
    HostGuard guard(G);
    if (G > batch) G = (int)batch;
    const int nd = lD + uD + 1;
    const int64_t per = (batch + G - 1) / G, wstride = ncap * (int64_t)(nd + r + 1);
    for (int d = 0; d < G; d++) {
        const int64_t lo = d * per, cnt = (lo + per <= batch ? per : batch - lo);
        if (cnt <= 0) continue;
        const int pd = phys(d, G);
        if ((rc = dev_init(pd))) return rc;
        cudaStream_t s = g_dev[pd].stream;
        double *d_dg, *d_bg, *d_sh = nullptr, *d_b, *d_work, *d_x;
        int64_t *d_ncols, *d_nconv;
        int32_t *d_st;
        if ((rc = dev_buf_t(pd, 0, (size_t)cnt * nd * 3, &d_dg))) return rc;
        if ((rc = dev_buf_t(pd, 1, (size_t)cnt * r * 4, &d_bg))) return rc;
        if (shift && (rc = dev_buf_t(pd, 2, (size_t)cnt, &d_sh))) return rc;
        if ((rc = dev_buf_t(pd, 3, (size_t)cnt * nb, &d_b))) return rc;
        if ((rc = dev_buf_t(pd, 16, (size_t)((cnt + 31) / 32) * 32 * wstride, &d_work))) return rc;     // whole warps (interleaved records)
        if ((rc = dev_buf_t(pd, 5, (size_t)cnt * ncap, &d_x))) return rc;
        if ((rc = dev_buf_t(pd, 6, (size_t)cnt, &d_ncols))) return rc;
        if ((rc = dev_buf_t(pd, 7, (size_t)cnt, &d_nconv))) return rc;
        if ((rc = dev_buf_t(pd, 8, (size_t)cnt, &d_st))) return rc;
        ILA_CUDA_CHECK(cudaMemcpyAsync(d_dg, dg + lo * nd * 3, sizeof(double) * cnt * nd * 3, cudaMemcpyHostToDevice, s));
        ILA_CUDA_CHECK(cudaMemcpyAsync(d_bg, bg + lo * r * 4, sizeof(double) * cnt * r * 4, cudaMemcpyHostToDevice, s));
        if (shift) ILA_CUDA_CHECK(cudaMemcpyAsync(d_sh, shift + lo, sizeof(double) * cnt, cudaMemcpyHostToDevice, s));
        ILA_CUDA_CHECK(cudaMemcpyAsync(d_b, b + lo * nb, sizeof(double) * cnt * nb, cudaMemcpyHostToDevice, s));
        AlmostBandedArgs a{cnt, lD, uD, r, d_dg, d_bg, d_sh, d_b, nb, colgrowth, tol, ncap, d_work, wstride, d_x, d_ncols, d_nconv, d_st};
        if ((rc = launch_qr_almostbanded(a, s))) return rc;
        ILA_CUDA_CHECK(cudaMemcpyAsync(x + lo * ncap, d_x, sizeof(double) * cnt * ncap, cudaMemcpyDeviceToHost, s));
        ILA_CUDA_CHECK(cudaMemcpyAsync(ncols + lo, d_ncols, sizeof(int64_t) * cnt, cudaMemcpyDeviceToHost, s));
        ILA_CUDA_CHECK(cudaMemcpyAsync(nconv + lo, d_nconv, sizeof(int64_t) * cnt, cudaMemcpyDeviceToHost, s));
        ILA_CUDA_CHECK(cudaMemcpyAsync(status + lo, d_st, sizeof(int32_t) * cnt, cudaMemcpyDeviceToHost, s));
    }
    for (int d = 0; d < G; d++) {
        const int pd = phys(d, G);
        if (!g_dev[pd].init) continue;
        ILA_CUDA_CHECK(cudaSetDevice(pd));
        ILA_CUDA_CHECK(cudaStreamSynchronize(g_dev[pd].stream));
    }
    guard.done = true;
    ILA_CUDA_CHECK(cudaSetDevice(g_base_dev));
    return 0;
}

// ================================ K9: adaptive finite-section QL =========================================

} // extern "C" (templates below have C++ linkage)

static inline int launch_fsql_any(int l, int u, const FsqlArgs &a, cudaStream_t s) { return launch_ql_finite_section(l, u, a, s); }
static inline int launch_fsql_any(int l, int u, const FsqlArgsC &a, cudaStream_t s) { return launch_ql_finite_section_c(l, u, a, s); }

template <typename T>
static int fsql_host(int l, int u, int64_t batch, const T *g, const T *shift, int hp, const T *head, int64_t j, double tol,
                     int64_t maxN, T *Lband, T *V, T *tau, int64_t *N_used, int32_t *status, int ngpus)
{
    if (batch < 0 || !g || hp < 0 || (hp > 0 && !head) || (batch > 0 && (!Lband || !V || !tau || !N_used || !status)))
        return set_error(ILA_ERR_ARG, "ql_finite_section: bad arguments");
    if (l < 1 || u < 1 || j < l + u + 2 || maxN < j) return set_error(ILA_ERR_ARG, "ql_finite_section: need l,u >= 1, j >= l+u+2, maxN >= j");
    int G = 0;
    int rc = resolve_ngpus(ngpus, &G);
    if (rc) return rc;
    if (batch == 0) return 0;
    std::lock_guard<std::mutex> lock(g_mutex);
    HostGuard guard(G);
    if (G > batch) G = (int)batch;
    const int nd = l + u + 1;
    const int64_t per = (batch + G - 1) / G;
    for (int d = 0; d < G; d++) {
        const int64_t lo = d * per, cnt = (lo + per <= batch ? per : batch - lo);
        if (cnt <= 0) continue;
        const int pd = phys(d, G);
        if ((rc = dev_init(pd))) return rc;
        T *d_g, *d_sh = nullptr, *d_head = nullptr, *d_L, *d_V, *d_tau;
        int64_t *d_N;
        int32_t *d_st;
        if ((rc = dev_buf_t(pd, 0, (size_t)cnt * nd * 5, &d_g))) return rc;
        if (shift && (rc = dev_buf_t(pd, 1, (size_t)cnt, &d_sh))) return rc;
        if (hp > 0 && (rc = dev_buf_t(pd, 2, (size_t)nd * hp, &d_head))) return rc;
        if ((rc = dev_buf_t(pd, 3, (size_t)cnt * j * nd, &d_L))) return rc;
        if ((rc = dev_buf_t(pd, 4, (size_t)cnt * j * u, &d_V))) return rc;
        if ((rc = dev_buf_t(pd, 5, (size_t)cnt * j, &d_tau))) return rc;
        if ((rc = dev_buf_t(pd, 6, (size_t)cnt, &d_N))) return rc;
        if ((rc = dev_buf_t(pd, 7, (size_t)cnt, &d_st))) return rc;
        cudaStream_t s = g_dev[pd].stream;
        ILA_CUDA_CHECK(cudaMemcpyAsync(d_g, g + lo * nd * 5, sizeof(T) * cnt * nd * 5, cudaMemcpyHostToDevice, s));
        if (shift) ILA_CUDA_CHECK(cudaMemcpyAsync(d_sh, shift + lo, sizeof(T) * cnt, cudaMemcpyHostToDevice, s));
        if (hp > 0) ILA_CUDA_CHECK(cudaMemcpyAsync(d_head, head, sizeof(T) * nd * hp, cudaMemcpyHostToDevice, s));
        FsqlArgsT<T> a{cnt, j, maxN, d_g, d_sh, hp, d_head, tol, d_L, d_V, d_tau, d_N, d_st};
        if ((rc = launch_fsql_any(l, u, a, s))) return rc;
        ILA_CUDA_CHECK(cudaMemcpyAsync(Lband + lo * j * nd, d_L, sizeof(T) * cnt * j * nd, cudaMemcpyDeviceToHost, s));
        ILA_CUDA_CHECK(cudaMemcpyAsync(V + lo * j * u, d_V, sizeof(T) * cnt * j * u, cudaMemcpyDeviceToHost, s));
        ILA_CUDA_CHECK(cudaMemcpyAsync(tau + lo * j, d_tau, sizeof(T) * cnt * j, cudaMemcpyDeviceToHost, s));
        ILA_CUDA_CHECK(cudaMemcpyAsync(N_used + lo, d_N, sizeof(int64_t) * cnt, cudaMemcpyDeviceToHost, s));
        ILA_CUDA_CHECK(cudaMemcpyAsync(status + lo, d_st, sizeof(int32_t) * cnt, cudaMemcpyDeviceToHost, s));
    }
    for (int d = 0; d < G; d++) {
        const int pd = phys(d, G);
        if (!g_dev[pd].init) continue;
        ILA_CUDA_CHECK(cudaSetDevice(pd));
        ILA_CUDA_CHECK(cudaStreamSynchronize(g_dev[pd].stream));
    }
    guard.done = true;
    ILA_CUDA_CHECK(cudaSetDevice(g_base_dev));
    return 0;
}

extern "C" {

int ila_ql_finite_section_f64(int l, int u, int64_t batch, const double *g, const double *shift, int hp, const double *head,
                              int64_t j, double tol, int64_t maxN, double *Lband, double *V, double *tau, int64_t *N_used,
                              int32_t *status, int ngpus)
{
    return fsql_host<double>(l, u, batch, g, shift, hp, head, j, tol, maxN, Lband, V, tau, N_used, status, ngpus);
}

int ila_ql_finite_section_c64(int l, int u, int64_t batch, const ila_c64 *g, const ila_c64 *shift, int hp, const ila_c64 *head,
                              int64_t j, double tol, int64_t maxN, ila_c64 *Lband, ila_c64 *V, ila_c64 *tau, int64_t *N_used,
                              int32_t *status, int ngpus)
{
    return fsql_host<cplx>(l, u, batch, (const cplx *)g, (const cplx *)shift, hp, (const cplx *)head, j, tol, maxN,
                           (cplx *)Lband, (cplx *)V, (cplx *)tau, N_used, status, ngpus);
}

// ================================ K6: block-tridiagonal Toeplitz-tail QL ================================

} // extern "C" (templates below have C++ linkage)

template <typename T>
static int blocktri_params(int n, const T *c, const T *a, const T *b, int n_dl, const T *dl_head, int n_d, const T *d_head,
                           int n_du, const T *du_head, double atol, int maxit, BlockTriParamsT<T> *P)
{
    if (n < 1 || n > kBtMaxN) return set_error(ILA_ERR_UNSUPPORTED, "ql_blocktri: block size n must be in 1..4");
    if (n_dl < 0 || n_d < 0 || n_du < 0 || n_dl > kBtMaxHead || n_d > kBtMaxHead || n_du + 1 > kBtMaxHead)
        return set_error(ILA_ERR_UNSUPPORTED, "ql_blocktri: at most 4 head block columns");
    if (!c || !a || !b || (n_dl && !dl_head) || (n_d && !d_head) || (n_du && !du_head) || maxit < 1)
        return set_error(ILA_ERR_ARG, "ql_blocktri: bad arguments");
    memset(P, 0, sizeof(*P));
    P->n = n; P->n_dl = n_dl; P->n_d = n_d; P->n_du = n_du; P->maxit = maxit; P->atol = atol;
    const int nn = n * n;
    memcpy(P->c, c, sizeof(T) * nn);
    memcpy(P->a, a, sizeof(T) * nn);
    memcpy(P->b, b, sizeof(T) * nn);
    for (int k = 0; k < n_dl; k++) memcpy(P->dl_head[k], dl_head + k * nn, sizeof(T) * nn);
    for (int k = 0; k < n_d; k++) memcpy(P->d_head[k], d_head + k * nn, sizeof(T) * nn);
    for (int k = 0; k < n_du; k++) memcpy(P->du_head[k], du_head + k * nn, sizeof(T) * nn);
    return 0;
}

template <typename T>
static int blocktri_dev(int n, const T *c, const T *a, const T *b, int n_dl, const T *dl_head, int n_d, const T *d_head,
                        int n_du, const T *du_head, const T *d_energy, int64_t batch, double atol, int maxit, T *d_L11,
                        int32_t *d_iters, int32_t *d_status, void *stream)
{
    if (ila_device_count() <= 0) return set_error(ILA_ERR_NOGPU, "no CUDA device visible: libila_b200 has no CPU fallback");
    BlockTriParamsT<T> P;
    int rc = blocktri_params<T>(n, c, a, b, n_dl, dl_head, n_d, d_head, n_du, du_head, atol, maxit, &P);
    if (rc) return rc;
    return launch_ql_blocktri_any(P, d_energy, batch, d_L11, d_iters, d_status, (cudaStream_t)stream);
}

template <typename T>
static int blocktri_host(int n, const T *c, const T *a, const T *b, int n_dl, const T *dl_head, int n_d, const T *d_head,
                         int n_du, const T *du_head, const T *energy, int64_t batch, double atol, int maxit, T *L11,
                         int32_t *iters, int32_t *status, int ngpus)
{
    if (batch < 0 || (batch > 0 && (!energy || !L11 || !iters || !status))) return set_error(ILA_ERR_ARG, "ql_blocktri: bad arguments");
    BlockTriParamsT<T> P;
    int rc = blocktri_params<T>(n, c, a, b, n_dl, dl_head, n_d, d_head, n_du, du_head, atol, maxit, &P);
    if (rc) return rc;
    int G = 0;
    if ((rc = resolve_ngpus(ngpus, &G))) return rc;
    if (batch == 0) return 0;
    std::lock_guard<std::mutex> lock(g_mutex);
    HostGuard guard(G);
    if (G > batch) G = (int)batch;
    const int64_t per = (batch + G - 1) / G;
    for (int d = 0; d < G; d++) {
        const int64_t lo = d * per, cnt = (lo + per <= batch ? per : batch - lo);
        if (cnt <= 0) continue;
        const int pd = phys(d, G);
        if ((rc = dev_init(pd))) return rc;
        T *d_e, *d_L;
        int32_t *d_it, *d_st;
        if ((rc = dev_buf_t(pd, 0, (size_t)cnt, &d_e))) return rc;
        if ((rc = dev_buf_t(pd, 1, (size_t)cnt, &d_L))) return rc;
        if ((rc = dev_buf_t(pd, 2, (size_t)cnt, &d_st))) return rc;
        if ((rc = dev_buf_t(pd, 3, (size_t)cnt, &d_it))) return rc;
        cudaStream_t s = g_dev[pd].stream;
        ILA_CUDA_CHECK(cudaMemcpyAsync(d_e, energy + lo, sizeof(T) * cnt, cudaMemcpyHostToDevice, s));
        if ((rc = launch_ql_blocktri_any(P, d_e, cnt, d_L, d_it, d_st, s))) return rc;
        ILA_CUDA_CHECK(cudaMemcpyAsync(L11 + lo, d_L, sizeof(T) * cnt, cudaMemcpyDeviceToHost, s));
        ILA_CUDA_CHECK(cudaMemcpyAsync(iters + lo, d_it, sizeof(int32_t) * cnt, cudaMemcpyDeviceToHost, s));
        ILA_CUDA_CHECK(cudaMemcpyAsync(status + lo, d_st, sizeof(int32_t) * cnt, cudaMemcpyDeviceToHost, s));
    }
    for (int d = 0; d < G; d++) {
        const int pd = phys(d, G);
        if (!g_dev[pd].init) continue;
        ILA_CUDA_CHECK(cudaSetDevice(pd));
        ILA_CUDA_CHECK(cudaStreamSynchronize(g_dev[pd].stream));
    }
    guard.done = true;
    ILA_CUDA_CHECK(cudaSetDevice(g_base_dev));
    return 0;
}

// the whole block factorization + Q'b (host entry): per energy the fixed point P (2n x 3n), τ tail, the factored head
// (Nn x Nn), its τ and, for a finitely supported b, (Q'b)[1:nout]
template <typename T>
static int blocktri_factors_host(int n, const T *c, const T *a, const T *b, int n_dl, const T *dl_head, int n_d, const T *d_head,
                                 int n_du, const T *du_head, const T *energy, int64_t batch, double atol, int maxit, T *L11,
                                 T *tail, T *tau_tail, T *head, T *head_tau, int nb, const T *rhs, int64_t nout, T *qtb,
                                 int32_t *iters, int32_t *status, int ngpus)
{
    if (batch < 0 || (batch > 0 && (!energy || !L11 || !iters || !status))) return set_error(ILA_ERR_ARG, "ql_blocktri_factors: bad arguments");
    if (qtb && (!rhs || nb < 1 || nb > kBtMaxB || nout < 1 || nout > kBtMaxB + 2 * n + 1))
        return set_error(ILA_ERR_ARG, "ql_blocktri_factors: Q'b needs 1 <= nb <= 16 and 1 <= nout <= 16 + 2n + 1");
    BlockTriParamsT<T> P;
    int rc = blocktri_params<T>(n, c, a, b, n_dl, dl_head, n_d, d_head, n_du, du_head, atol, maxit, &P);
    if (rc) return rc;
    int N = n_du + 1;
    if (n_d > N) N = n_d;
    if (n_dl > N) N = n_dl;
    const int64_t hn = (int64_t)N * n, tl = 6 * (int64_t)n * n;
    int G = 0;
    if ((rc = resolve_ngpus(ngpus, &G))) return rc;
    if (batch == 0) return 0;
    std::lock_guard<std::mutex> lock(g_mutex);
    HostGuard guard(G);
    if (G > batch) G = (int)batch;
    const int64_t per = (batch + G - 1) / G;
    for (int d = 0; d < G; d++) {
        const int64_t lo = d * per, cnt = (lo + per <= batch ? per : batch - lo);
        if (cnt <= 0) continue;
        const int pd = phys(d, G);
        if ((rc = dev_init(pd))) return rc;
        T *d_e, *d_L;
        int32_t *d_it, *d_st;
        BlockTriOutT<T> O;
        if ((rc = dev_buf_t(pd, 0, (size_t)cnt, &d_e))) return rc;
        if ((rc = dev_buf_t(pd, 1, (size_t)cnt, &d_L))) return rc;
        if ((rc = dev_buf_t(pd, 2, (size_t)cnt, &d_st))) return rc;
        if ((rc = dev_buf_t(pd, 3, (size_t)cnt, &d_it))) return rc;
        if (tail && (rc = dev_buf_t(pd, 4, (size_t)cnt * tl, &O.tail))) return rc;
        if (tau_tail && (rc = dev_buf_t(pd, 5, (size_t)cnt * n, &O.tau_tail))) return rc;
        if (head && (rc = dev_buf_t(pd, 6, (size_t)cnt * hn * hn, &O.head))) return rc;
        if (head_tau && (rc = dev_buf_t(pd, 7, (size_t)cnt * hn, &O.head_tau))) return rc;
        if (qtb) {
            if ((rc = dev_buf_t(pd, 8, (size_t)cnt * nout, &O.qtb))) return rc;
            O.nb = nb;
            O.nout = (int)nout;
            for (int k = 0; k < kBtMaxB; k++) O.b[k] = k < nb ? rhs[k] : T{};
        }
        cudaStream_t s = g_dev[pd].stream;
        ILA_CUDA_CHECK(cudaMemcpyAsync(d_e, energy + lo, sizeof(T) * cnt, cudaMemcpyHostToDevice, s));
        if ((rc = launch_ql_blocktri_any(P, d_e, cnt, d_L, d_it, d_st, s, O))) return rc;
        ILA_CUDA_CHECK(cudaMemcpyAsync(L11 + lo, d_L, sizeof(T) * cnt, cudaMemcpyDeviceToHost, s));
        ILA_CUDA_CHECK(cudaMemcpyAsync(iters + lo, d_it, sizeof(int32_t) * cnt, cudaMemcpyDeviceToHost, s));
        ILA_CUDA_CHECK(cudaMemcpyAsync(status + lo, d_st, sizeof(int32_t) * cnt, cudaMemcpyDeviceToHost, s));
        if (tail) ILA_CUDA_CHECK(cudaMemcpyAsync(tail + lo * tl, O.tail, sizeof(T) * cnt * tl, cudaMemcpyDeviceToHost, s));
        if (tau_tail) ILA_CUDA_CHECK(cudaMemcpyAsync(tau_tail + lo * n, O.tau_tail, sizeof(T) * cnt * n, cudaMemcpyDeviceToHost, s));
        if (head) ILA_CUDA_CHECK(cudaMemcpyAsync(head + lo * hn * hn, O.head, sizeof(T) * cnt * hn * hn, cudaMemcpyDeviceToHost, s));
        if (head_tau) ILA_CUDA_CHECK(cudaMemcpyAsync(head_tau + lo * hn, O.head_tau, sizeof(T) * cnt * hn, cudaMemcpyDeviceToHost, s));
        if (qtb) ILA_CUDA_CHECK(cudaMemcpyAsync(qtb + lo * nout, O.qtb, sizeof(T) * cnt * nout, cudaMemcpyDeviceToHost, s));
    }
    for (int d = 0; d < G; d++) {
        const int pd = phys(d, G);
        if (!g_dev[pd].init) continue;
        ILA_CUDA_CHECK(cudaSetDevice(pd));
        ILA_CUDA_CHECK(cudaStreamSynchronize(g_dev[pd].stream));
    }
    guard.done = true;
    ILA_CUDA_CHECK(cudaSetDevice(g_base_dev));
    return 0;
}

extern "C" {

int ila_ql_blocktri_factors_f64(int n, const double *c, const double *a, const double *b, int n_dl, const double *dl_head,
                                int n_d, const double *d_head, int n_du, const double *du_head, const double *energy,
                                int64_t batch, double atol, int maxit, double *L11, double *tail_opt, double *tau_tail_opt,
                                double *head_opt, double *head_tau_opt, int nb, const double *rhs, int64_t nout, double *qtb_opt,
                                int32_t *iters, int32_t *status, int ngpus)
{
    return blocktri_factors_host<double>(n, c, a, b, n_dl, dl_head, n_d, d_head, n_du, du_head, energy, batch, atol, maxit, L11,
                                         tail_opt, tau_tail_opt, head_opt, head_tau_opt, nb, rhs, nout, qtb_opt, iters, status, ngpus);
}

int ila_ql_blocktri_factors_c64(int n, const ila_c64 *c, const ila_c64 *a, const ila_c64 *b, int n_dl, const ila_c64 *dl_head,
                                int n_d, const ila_c64 *d_head, int n_du, const ila_c64 *du_head, const ila_c64 *energy,
                                int64_t batch, double atol, int maxit, ila_c64 *L11, ila_c64 *tail_opt, ila_c64 *tau_tail_opt,
                                ila_c64 *head_opt, ila_c64 *head_tau_opt, int nb, const ila_c64 *rhs, int64_t nout,
                                ila_c64 *qtb_opt, int32_t *iters, int32_t *status, int ngpus)
{
    return blocktri_factors_host<cplx>(n, (const cplx *)c, (const cplx *)a, (const cplx *)b, n_dl, (const cplx *)dl_head, n_d,
                                       (const cplx *)d_head, n_du, (const cplx *)du_head, (const cplx *)energy, batch, atol, maxit,
                                       (cplx *)L11, (cplx *)tail_opt, (cplx *)tau_tail_opt, (cplx *)head_opt, (cplx *)head_tau_opt,
                                       nb, (const cplx *)rhs, nout, (cplx *)qtb_opt, iters, status, ngpus);
}

int ila_ql_blocktri_l11_f64_dev(int n, const double *c, const double *a, const double *b, int n_dl, const double *dl_head,
                                int n_d, const double *d_head, int n_du, const double *du_head, const double *d_energy,
                                int64_t batch, double atol, int maxit, double *d_L11, int32_t *d_iters, int32_t *d_status,
                                void *stream)
{
    return blocktri_dev<double>(n, c, a, b, n_dl, dl_head, n_d, d_head, n_du, du_head, d_energy, batch, atol, maxit, d_L11,
                                d_iters, d_status, stream);
}

int ila_ql_blocktri_l11_f64(int n, const double *c, const double *a, const double *b, int n_dl, const double *dl_head,
                            int n_d, const double *d_head, int n_du, const double *du_head, const double *energy,
                            int64_t batch, double atol, int maxit, double *L11, int32_t *iters, int32_t *status, int ngpus)
{
    return blocktri_host<double>(n, c, a, b, n_dl, dl_head, n_d, d_head, n_du, du_head, energy, batch, atol, maxit, L11,
                                 iters, status, ngpus);
}

int ila_ql_blocktri_l11_c64_dev(int n, const ila_c64 *c, const ila_c64 *a, const ila_c64 *b, int n_dl, const ila_c64 *dl_head,
                                int n_d, const ila_c64 *d_head, int n_du, const ila_c64 *du_head, const ila_c64 *d_energy,
                                int64_t batch, double atol, int maxit, ila_c64 *d_L11, int32_t *d_iters, int32_t *d_status,
                                void *stream)
{
    return blocktri_dev<cplx>(n, (const cplx *)c, (const cplx *)a, (const cplx *)b, n_dl, (const cplx *)dl_head, n_d,
                              (const cplx *)d_head, n_du, (const cplx *)du_head, (const cplx *)d_energy, batch, atol, maxit,
                              (cplx *)d_L11, d_iters, d_status, stream);
}

int ila_ql_blocktri_l11_c64(int n, const ila_c64 *c, const ila_c64 *a, const ila_c64 *b, int n_dl, const ila_c64 *dl_head,
                            int n_d, const ila_c64 *d_head, int n_du, const ila_c64 *du_head, const ila_c64 *energy,
                            int64_t batch, double atol, int maxit, ila_c64 *L11, int32_t *iters, int32_t *status, int ngpus)
{
    return blocktri_host<cplx>(n, (const cplx *)c, (const cplx *)a, (const cplx *)b, n_dl, (const cplx *)dl_head, n_d,
                               (const cplx *)d_head, n_du, (const cplx *)du_head, (const cplx *)energy, batch, atol, maxit,
                               (cplx *)L11, iters, status, ngpus);
}

int ila_ql_set_max_sweeps(int fp64_sweeps, int scout_sweeps)
{
    if (fp64_sweeps < 0 || scout_sweeps < 0) return set_error(ILA_ERR_ARG, "sweep limits must be >= 0 (0 = default)");
    tz_set_max_sweeps(fp64_sweeps, scout_sweeps);
    return 0;
}

/* Debug build (-DILA_BOUNDS, libila_b200_dbg.so): out-of-range accesses recorded by the bounds-checked pointers of K1/K2,
 * K5 and K11 since the last call.  count = -1 in the release build (not instrumented). */
int ila_debug_bounds_report(int64_t *count, int64_t *site, int64_t *index, int64_t *extent)
{
    unsigned long long v[3][4] = {};
    int inst = 0;
    int rc;
    if ((rc = qr_banded_bounds_report(v[0])) < 0) return rc;
    inst += rc;
    if ((rc = chol_banded_bounds_report(v[1])) < 0) return rc;
    inst += rc;
    if ((rc = qr_blockbanded_bounds_report(v[2])) < 0) return rc;
    inst += rc;
    int64_t c = 0, s = 0, i = 0, e = 0;
    for (int k = 0; k < 3; k++) {
        if (v[k][0] && !c) { s = (int64_t)v[k][1]; i = (int64_t)v[k][2]; e = (int64_t)v[k][3]; }
        c += (int64_t)v[k][0];
    }
    if (count) *count = inst ? c : -1;
    if (site) *site = s;
    if (index) *index = i;
    if (extent) *extent = e;
    return 0;
}

int ila_ql_set_cold_start(int closed_form_seeds)
{
    tz_set_seeds(closed_form_seeds);
    return 0;
}

int ila_ql_set_shifts_per_thread(int spt)
{
    if (spt < 0 || spt > 16) return set_error(ILA_ERR_ARG, "shifts per thread must be in 0..16 (0 = automatic)");
    tz_set_spt(spt);
    return 0;
}

// ================================ K1+K2: adaptive banded QR solve =======================================

static int g_qr_direct_x = -1;   // -1 auto, 0 always interleaved + compaction, 1 always direct

int ila_qr_set_arith_mode(int mode)
{
    // 0 / 1: arithmetic of the sweeps; 2 / 3 / 4 (validation knobs): solution output auto / always compacted / always direct
    if (mode >= 2 && mode <= 4) { g_qr_direct_x = mode == 2 ? -1 : (mode == 3 ? 0 : 1); return 0; }
    if (mode == 5 || mode == 6) { qr_set_bulk_loader(mode == 6); return 0; }   // back-substitution loader: LDGSTS / cp.async.bulk
    if (mode >= 1000) { qr_set_fwd_fault(mode - 1000); return 0; }   // test hook: the scout of the three-warp sweep misrounds a candidate every (mode - 1000) columns
    if (mode == 7 || mode == 8 || mode == 11) { qr_set_fwd_duo(mode == 7 ? 0 : (mode == 8 ? 1 : 2)); return 0; }   // forward sweep (l = 1): one warp / three warps for batches of at most one group per SM / three warps always
    if (mode != 0 && mode != 1) return set_error(ILA_ERR_ARG, "qr arith mode must be 0 (branch-free exact), 1 (IEEE intrinsics), 2-4 (output path), 5-6 (loader) or 7, 8, 11 (forward sweep warps)");
    qr_set_arith_mode(mode);
    return 0;
}

int ila_qr_banded_work_layout(int l, int u, int64_t batch, int64_t ncap, const int64_t *ncap_per, int64_t *gbase_host,
                              int64_t *work_bytes, int64_t *xi_bytes)
{
    if (batch < 0 || !gbase_host) return set_error(ILA_ERR_ARG, "qr_banded_work_layout: bad arguments");
    const int64_t groups = (batch + 31) / 32;
    int64_t acc = 0;
    for (int64_t g = 0; g < groups; g++) {
        gbase_host[g] = acc;
        int64_t mx = 0;
        for (int64_t i = g * 32; i < batch && i < (g + 1) * 32; i++) {
            const int64_t c = ncap_per ? ncap_per[i] : ncap;
            if (c > mx) mx = c;
        }
        acc += (mx + 7) & ~(int64_t)7;
    }
    gbase_host[groups] = acc;
    // records sized for the widest form (with reflectors) so one buffer serves both modes
    if (work_bytes) *work_bytes = acc * 32 * (int64_t)sizeof(double) * qr_record_doubles(l, u, true);
    if (xi_bytes) *xi_bytes = acc * 32 * (int64_t)sizeof(double);
    return 0;
}

int ila_qr_banded_state_doubles(int l, int u) { return qr_state_doubles(l, u); }
int ila_chol_banded_state_doubles(int u) { return chol_state_doubles(u); }

int ila_qr_banded_factor_resume_f64_dev(int l, int u, int64_t batch, const double *d_p, const double *d_q, const double *d_r,
                                        const double *d_shift, int hp, const double *d_head, int nb, const double *d_b,
                                        double tol, int colgrowth, int64_t ncap, const int64_t *d_ncap_per,
                                        const int64_t *d_gbase, void *d_work, int64_t *d_ncols, int64_t *d_nconv,
                                        int64_t *d_nswept, int64_t *d_xoff, int32_t *d_status, int want_reflectors,
                                        double *d_state, int resume, void *stream)
{
    if (ila_device_count() <= 0) return set_error(ILA_ERR_NOGPU, "no CUDA device visible: libila_b200 has no CPU fallback");
    if (batch < 0 || nb < 1 || colgrowth < 1 || hp < 0) return set_error(ILA_ERR_ARG, "qr_banded: bad sizes");
    if (resume && !d_state) return set_error(ILA_ERR_ARG, "qr_banded: resume needs the state buffer of the earlier call");
    QrFwdArgs a{batch, d_p, d_q, d_r, d_shift, d_head, d_b, hp, nb, colgrowth, tol, d_gbase, d_ncap_per, ncap,
                (double2 *)d_work, d_ncols, d_nconv, d_nswept, d_status};
    a.state = d_state;
    a.resume = resume ? 1 : 0;
    int rc = launch_qr_forward(l, u, a, want_reflectors != 0, (cudaStream_t)stream);
    if (rc) return rc;
    return launch_qr_scan(d_ncols, batch, d_xoff, (cudaStream_t)stream);
}

int ila_qr_banded_factor_f64_dev(int l, int u, int64_t batch, const double *d_p, const double *d_q, const double *d_r,
                                 const double *d_shift, int hp, const double *d_head, int nb, const double *d_b,
                                 double tol, int colgrowth, int64_t ncap, const int64_t *d_ncap_per,
                                 const int64_t *d_gbase, void *d_work, int64_t *d_ncols, int64_t *d_nconv,
                                 int64_t *d_nswept, int64_t *d_xoff, int32_t *d_status, int want_reflectors, void *stream)
{
    return ila_qr_banded_factor_resume_f64_dev(l, u, batch, d_p, d_q, d_r, d_shift, hp, d_head, nb, d_b, tol, colgrowth, ncap,
                                               d_ncap_per, d_gbase, d_work, d_ncols, d_nconv, d_nswept, d_xoff, d_status,
                                               want_reflectors, nullptr, 0, stream);
}

int ila_qr_banded_partialqr_f64_dev(int l, int u, int64_t batch, const double *d_p, const double *d_q, const double *d_r,
                                    const double *d_shift, int hp, const double *d_head, int64_t n, int colgrowth,
                                    const int64_t *d_gbase, void *d_work, double *d_state, int resume, int64_t *d_ncols,
                                    int64_t *d_nswept, int64_t *d_xoff, int32_t *d_status, void *stream)
{
    if (ila_device_count() <= 0) return set_error(ILA_ERR_NOGPU, "no CUDA device visible: libila_b200 has no CPU fallback");
    if (batch < 0 || n < 0 || hp < 0 || colgrowth < 1 || !d_state || !d_ncols || !d_nswept || !d_xoff || !d_status)
        return set_error(ILA_ERR_ARG, "qr_banded_partialqr: bad arguments");
    QrFwdArgs a{batch, d_p, d_q, d_r, d_shift, d_head, nullptr, hp, 0, colgrowth, 0.0, d_gbase, nullptr, n,
                (double2 *)d_work, d_ncols, d_nswept, d_nswept, d_status};     // nconv aliases nswept (written first)
    a.state = d_state;
    a.resume = resume ? 1 : 0;
    a.factor_only = 1;
    int rc = launch_qr_forward(l, u, a, true, (cudaStream_t)stream);
    if (rc) return rc;
    return launch_qr_scan(d_ncols, batch, d_xoff, (cudaStream_t)stream);
}

int ila_qr_banded_export_f64_dev(int l, int u, int want_reflectors, int64_t batch, const int64_t *d_gbase, const void *d_work,
                                 const int64_t *d_ncols, const int64_t *d_nswept, const int64_t *d_xoff, double *d_factors,
                                 double *d_tau, double *d_qtb, void *stream)
{
    if (ila_device_count() <= 0) return set_error(ILA_ERR_NOGPU, "no CUDA device visible: libila_b200 has no CPU fallback");
    if (batch < 0 || !d_gbase || !d_work || !d_ncols || !d_nswept || !d_xoff) return set_error(ILA_ERR_ARG, "qr_banded_export: bad arguments");
    if (batch == 0) return 0;
    return launch_qr_export(l, u, want_reflectors ? 1 : 0, batch, d_gbase, (const double2 *)d_work, d_ncols, d_nswept, d_xoff,
                            d_factors, d_tau, d_qtb, (cudaStream_t)stream);
}

int ila_qr_banded_backsolve_f64_dev(int l, int u, int64_t batch, const int64_t *d_gbase, const void *d_work, double *d_xi,
                                    const int64_t *d_ncols, const int64_t *d_nswept, const int64_t *d_xoff,
                                    const int32_t *d_status, double *d_x, int64_t xcap, int want_reflectors, void *stream)
{
    if (ila_device_count() <= 0) return set_error(ILA_ERR_NOGPU, "no CUDA device visible: libila_b200 has no CPU fallback");
    // few groups (latency-bound regime: at most two solver warps per SM): the solver writes x straight into the ragged
    // output and the compaction pass disappears; many groups (HBM-bound regime): coalesced interleaved stream + compaction
    int sms = 148;
    {
        int dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const bool direct = g_qr_direct_x >= 0 ? g_qr_direct_x != 0 : (batch + 31) / 32 <= 2 * (int64_t)sms;
    QrBwdArgs a{batch, d_gbase, (const double2 *)d_work, d_xi, d_ncols, d_nswept, d_status, direct ? d_x : nullptr, d_xoff};
    int rc = launch_qr_backsolve(l, u, a, want_reflectors != 0, (cudaStream_t)stream);
    if (rc || direct) return rc;
    return launch_qr_compact(batch, d_gbase, d_xi, d_ncols, d_nswept, d_xoff, d_x, xcap, (cudaStream_t)stream);
}

int ila_chol_banded_factor_resume_f64_dev(int u, int64_t batch, const double *d_p, const double *d_q, const double *d_r,
                                          const double *d_shift, int hp, const double *d_head, int nb, const double *d_b,
                                          double tol, int colgrowth, int64_t ncap, const int64_t *d_ncap_per,
                                          const int64_t *d_gbase, void *d_work, int64_t *d_ncols, int64_t *d_nconv,
                                          int64_t *d_nswept, int64_t *d_xoff, int32_t *d_status, int full_sweep,
                                          double *d_state, int resume, void *stream)
{
    if (ila_device_count() <= 0) return set_error(ILA_ERR_NOGPU, "no CUDA device visible: libila_b200 has no CPU fallback");
    if (batch < 0 || nb < 1 || colgrowth < 1 || hp < 0) return set_error(ILA_ERR_ARG, "chol_banded: bad sizes");
    if (resume && !d_state) return set_error(ILA_ERR_ARG, "chol_banded: resume needs the state buffer of the earlier call");
    CholFwdArgs a{batch, d_p, d_q, d_r, d_shift, d_head, d_b, hp, nb, colgrowth, full_sweep, tol, d_gbase, d_ncap_per, ncap,
                  (double2 *)d_work, d_ncols, d_nconv, d_nswept, d_status};
    a.state = d_state;
    a.resume = resume ? 1 : 0;
    int rc = launch_chol_forward(u, a, (cudaStream_t)stream);
    if (rc) return rc;
    return launch_qr_scan(d_ncols, batch, d_xoff, (cudaStream_t)stream);
}

int ila_chol_banded_factor_f64_dev(int u, int64_t batch, const double *d_p, const double *d_q, const double *d_r,
                                   const double *d_shift, int hp, const double *d_head, int nb, const double *d_b,
                                   double tol, int colgrowth, int64_t ncap, const int64_t *d_ncap_per,
                                   const int64_t *d_gbase, void *d_work, int64_t *d_ncols, int64_t *d_nconv,
                                   int64_t *d_nswept, int64_t *d_xoff, int32_t *d_status, int full_sweep, void *stream)
{
    return ila_chol_banded_factor_resume_f64_dev(u, batch, d_p, d_q, d_r, d_shift, hp, d_head, nb, d_b, tol, colgrowth, ncap,
                                                 d_ncap_per, d_gbase, d_work, d_ncols, d_nconv, d_nswept, d_xoff, d_status,
                                                 full_sweep, nullptr, 0, stream);
}

// shared host driver of the two adaptive banded solves: kind 0 = QR (K1+K2), kind 1 = Cholesky (K5+K2, l == 0)
static int banded_solve_host(int kind, int l, int u, int64_t batch, const double *p, const double *q, const double *r,
                            const double *shift, int hp, const double *head, int nb, const double *b, double tol,
                            int colgrowth, int64_t ncap, const int64_t *ncap_per, double *x, int64_t xcap,
                            int64_t *xoff, int64_t *ncols, int64_t *nconv, double *factors_opt, double *tau_opt,
                            double *qtb_opt, int32_t *status, int ngpus)
{
    if (batch < 0 || nb < 1 || colgrowth < 1 || hp < 0 || !p || !q || !b || !xoff || !ncols || !nconv || !status || (hp > 0 && !head))
        return set_error(ILA_ERR_ARG, "qr_banded_solve: bad arguments");
    if (kind == 0 && l < 1) return set_error(ILA_ERR_UNSUPPORTED, "qr_banded_solve: l == 0 (triangular operator, Q = I) is not on the accelerated path");
    if (kind == 1 && (tau_opt || l != 0)) return set_error(ILA_ERR_ARG, "chol_banded_solve: bad arguments");
    int G = 0;
    int rc = resolve_ngpus(ngpus, &G);
    if (rc) return rc;
    if (batch == 0) { xoff[0] = 0; return 0; }
    std::lock_guard<std::mutex> lock(g_mutex);
    HostGuard guard(G);
    if (G > batch) G = (int)batch;
    const int nd = l + u + 1, LD = 2 * l + u + 1;
    const bool refl = kind == 0 && (factors_opt || tau_opt);
    const bool want_export = factors_opt || tau_opt || qtb_opt;
    const int REC = qr_record_doubles(l, u, refl);

    // round-robin sharding: instance i -> device i % G, local index i / G
    struct Shard {
        int64_t cnt = 0;
        std::vector<double> p, q, r, shift, b;
        std::vector<int64_t> gbase, caps, ncols, nconv, nswept, xoff;
        int64_t work_bytes = 0, xi_bytes = 0;
        std::vector<int32_t> status;
        double *d_p, *d_q, *d_r, *d_shift, *d_head, *d_b, *d_x, *d_xi, *d_fac, *d_tau, *d_qtb;
        char *d_work;
        int64_t *d_gbase, *d_caps, *d_ncols, *d_nconv, *d_nswept, *d_xoff;
        int32_t *d_status;
        int64_t total = 0;
    };
    std::vector<Shard> sh(G);
    for (int d = 0; d < G; d++) {
        Shard &S = sh[d];
        S.cnt = (batch - d + G - 1) / G;
        S.p.resize(S.cnt * nd); S.q.resize(S.cnt * nd); S.b.resize(S.cnt * nb);
        if (r) S.r.resize(S.cnt * nd);
        if (shift) S.shift.resize(S.cnt);
        S.caps.resize(S.cnt);
        std::vector<int64_t> &caps = S.caps;
        for (int64_t j = 0; j < S.cnt; j++) {
            const int64_t i = j * G + d;
            memcpy(&S.p[j * nd], p + i * nd, sizeof(double) * nd);
            memcpy(&S.q[j * nd], q + i * nd, sizeof(double) * nd);
            if (r) memcpy(&S.r[j * nd], r + i * nd, sizeof(double) * nd);
            if (shift) S.shift[j] = shift[i];
            memcpy(&S.b[j * nb], b + i * nb, sizeof(double) * nb);
            caps[j] = ncap_per ? ncap_per[i] : ncap;
        }
        S.gbase.resize((S.cnt + 31) / 32 + 1);
        ila_qr_banded_work_layout(l, u, S.cnt, 0, caps.data(), S.gbase.data(), &S.work_bytes, &S.xi_bytes);
        S.ncols.resize(S.cnt); S.nconv.resize(S.cnt); S.nswept.resize(S.cnt); S.status.resize(S.cnt); S.xoff.resize(S.cnt + 1);
    }
    // phase 1 on every device
    for (int d = 0; d < G; d++) {
        Shard &S = sh[d];
        const int pd = phys(d, G);
        if ((rc = dev_init(pd))) return rc;
        cudaStream_t s = g_dev[pd].stream;
        if ((rc = dev_buf_t(pd, 4, S.p.size(), &S.d_p))) return rc;
        if ((rc = dev_buf_t(pd, 5, S.q.size(), &S.d_q))) return rc;
        S.d_r = nullptr; S.d_shift = nullptr; S.d_head = nullptr;
        if (r && (rc = dev_buf_t(pd, 6, S.r.size(), &S.d_r))) return rc;
        if (shift && (rc = dev_buf_t(pd, 7, S.shift.size(), &S.d_shift))) return rc;
        if (hp > 0 && (rc = dev_buf_t(pd, 8, (size_t)nd * hp, &S.d_head))) return rc;
        if ((rc = dev_buf_t(pd, 9, S.b.size(), &S.d_b))) return rc;
        if ((rc = dev_buf_t(pd, 10, S.gbase.size(), &S.d_gbase))) return rc;
        if ((rc = dev_buf_t(pd, 21, (size_t)S.cnt, &S.d_caps))) return rc;
        if ((rc = dev_buf_t(pd, 11, (size_t)S.cnt, &S.d_ncols))) return rc;
        if ((rc = dev_buf_t(pd, 12, (size_t)S.cnt, &S.d_nconv))) return rc;
        if ((rc = dev_buf_t(pd, 13, (size_t)S.cnt, &S.d_nswept))) return rc;
        if ((rc = dev_buf_t(pd, 14, (size_t)S.cnt + 1, &S.d_xoff))) return rc;
        if ((rc = dev_buf_t(pd, 15, (size_t)S.cnt, &S.d_status))) return rc;
        if ((rc = dev_buf_t(pd, 16, (size_t)S.work_bytes, &S.d_work))) return rc;
        if ((rc = dev_buf_t(pd, 22, (size_t)S.xi_bytes / sizeof(double), &S.d_xi))) return rc;
        ILA_CUDA_CHECK(cudaMemcpyAsync(S.d_p, S.p.data(), sizeof(double) * S.p.size(), cudaMemcpyHostToDevice, s));
        ILA_CUDA_CHECK(cudaMemcpyAsync(S.d_q, S.q.data(), sizeof(double) * S.q.size(), cudaMemcpyHostToDevice, s));
        if (r) ILA_CUDA_CHECK(cudaMemcpyAsync(S.d_r, S.r.data(), sizeof(double) * S.r.size(), cudaMemcpyHostToDevice, s));
        if (shift) ILA_CUDA_CHECK(cudaMemcpyAsync(S.d_shift, S.shift.data(), sizeof(double) * S.shift.size(), cudaMemcpyHostToDevice, s));
        if (hp > 0) ILA_CUDA_CHECK(cudaMemcpyAsync(S.d_head, head, sizeof(double) * nd * hp, cudaMemcpyHostToDevice, s));
        ILA_CUDA_CHECK(cudaMemcpyAsync(S.d_b, S.b.data(), sizeof(double) * S.b.size(), cudaMemcpyHostToDevice, s));
        ILA_CUDA_CHECK(cudaMemcpyAsync(S.d_gbase, S.gbase.data(), sizeof(int64_t) * S.gbase.size(), cudaMemcpyHostToDevice, s));
        ILA_CUDA_CHECK(cudaMemcpyAsync(S.d_caps, S.caps.data(), sizeof(int64_t) * S.cnt, cudaMemcpyHostToDevice, s));
        if (kind == 0)
            rc = ila_qr_banded_factor_f64_dev(l, u, S.cnt, S.d_p, S.d_q, S.d_r, S.d_shift, hp, S.d_head, nb, S.d_b, tol, colgrowth,
                                              0, S.d_caps, S.d_gbase, S.d_work, S.d_ncols, S.d_nconv, S.d_nswept, S.d_xoff, S.d_status, refl, s);
        else
            rc = ila_chol_banded_factor_f64_dev(u, S.cnt, S.d_p, S.d_q, S.d_r, S.d_shift, hp, S.d_head, nb, S.d_b, tol, colgrowth,
                                                0, S.d_caps, S.d_gbase, S.d_work, S.d_ncols, S.d_nconv, S.d_nswept, S.d_xoff, S.d_status, want_export ? 1 : 0, s);
        if (rc) return rc;
        ILA_CUDA_CHECK(cudaMemcpyAsync(S.ncols.data(), S.d_ncols, sizeof(int64_t) * S.cnt, cudaMemcpyDeviceToHost, s));
        ILA_CUDA_CHECK(cudaMemcpyAsync(S.nconv.data(), S.d_nconv, sizeof(int64_t) * S.cnt, cudaMemcpyDeviceToHost, s));
        ILA_CUDA_CHECK(cudaMemcpyAsync(S.status.data(), S.d_status, sizeof(int32_t) * S.cnt, cudaMemcpyDeviceToHost, s));
        ILA_CUDA_CHECK(cudaMemcpyAsync(S.xoff.data(), S.d_xoff, sizeof(int64_t) * (S.cnt + 1), cudaMemcpyDeviceToHost, s));
    }
    for (int d = 0; d < G; d++) {
        ILA_CUDA_CHECK(cudaSetDevice(phys(d, G)));
        ILA_CUDA_CHECK(cudaStreamSynchronize(g_dev[phys(d, G)].stream));
    }
    // the one host round trip per call (not per column): sizes of the ragged result
    int64_t acc = 0;
    for (int64_t i = 0; i < batch; i++) {
        const Shard &S = sh[i % G];
        const int64_t j = i / G;
        xoff[i] = acc;
        ncols[i] = S.ncols[j];
        nconv[i] = S.nconv[j];
        status[i] = S.status[j];
        acc += S.ncols[j];
    }
    xoff[batch] = acc;
    if (acc > xcap || !x) {
        guard.done = true;
    ILA_CUDA_CHECK(cudaSetDevice(g_base_dev));
        return set_error(ILA_ERR_CAPACITY, "qr_banded_solve: x capacity too small; xoff[batch] holds the required number of doubles");
    }
    // phase 2
    for (int d = 0; d < G; d++) {
        Shard &S = sh[d];
        const int pd = phys(d, G);
        ILA_CUDA_CHECK(cudaSetDevice(pd));
        cudaStream_t s = g_dev[pd].stream;
        S.total = S.xoff[S.cnt];
        if ((rc = dev_buf_t(pd, 17, (size_t)S.total, &S.d_x))) return rc;
        rc = ila_qr_banded_backsolve_f64_dev(l, u, S.cnt, S.d_gbase, S.d_work, S.d_xi, S.d_ncols, S.d_nswept, S.d_xoff, S.d_status, S.d_x,
                                             S.total, refl, s);
        if (rc) return rc;
        S.d_fac = S.d_tau = S.d_qtb = nullptr;
        if (want_export) {
            if (kind == 0 && !refl) return set_error(ILA_ERR_UNSUPPORTED, "qr_banded_solve: qtb_opt requires factors_opt or tau_opt");
            if (factors_opt) {
                if ((rc = dev_buf_t(pd, 18, (size_t)S.total * LD, &S.d_fac))) return rc;
                ILA_CUDA_CHECK(cudaMemsetAsync(S.d_fac, 0, sizeof(double) * S.total * LD, s));
            }
            if (tau_opt && (rc = dev_buf_t(pd, 19, (size_t)S.total, &S.d_tau))) return rc;
            if (qtb_opt && (rc = dev_buf_t(pd, 20, (size_t)S.total, &S.d_qtb))) return rc;
            if ((rc = launch_qr_export(l, u, refl ? 1 : 0, S.cnt, S.d_gbase, (const double2 *)S.d_work, S.d_ncols, S.d_nswept, S.d_xoff, S.d_fac, S.d_tau, S.d_qtb, s))) return rc;
        }
        if (G == 1) {
            ILA_CUDA_CHECK(cudaMemcpyAsync(x, S.d_x, sizeof(double) * S.total, cudaMemcpyDeviceToHost, s));
            if (factors_opt) ILA_CUDA_CHECK(cudaMemcpyAsync(factors_opt, S.d_fac, sizeof(double) * S.total * LD, cudaMemcpyDeviceToHost, s));
            if (tau_opt) ILA_CUDA_CHECK(cudaMemcpyAsync(tau_opt, S.d_tau, sizeof(double) * S.total, cudaMemcpyDeviceToHost, s));
            if (qtb_opt) ILA_CUDA_CHECK(cudaMemcpyAsync(qtb_opt, S.d_qtb, sizeof(double) * S.total, cudaMemcpyDeviceToHost, s));
        } else {
            for (int64_t j = 0; j < S.cnt; j++) {
                const int64_t i = j * G + d, nn = S.ncols[j];
                if (nn <= 0) continue;
                ILA_CUDA_CHECK(cudaMemcpyAsync(x + xoff[i], S.d_x + S.xoff[j], sizeof(double) * nn, cudaMemcpyDeviceToHost, s));
                if (factors_opt) ILA_CUDA_CHECK(cudaMemcpyAsync(factors_opt + xoff[i] * LD, S.d_fac + S.xoff[j] * LD, sizeof(double) * nn * LD, cudaMemcpyDeviceToHost, s));
                if (tau_opt) ILA_CUDA_CHECK(cudaMemcpyAsync(tau_opt + xoff[i], S.d_tau + S.xoff[j], sizeof(double) * nn, cudaMemcpyDeviceToHost, s));
                if (qtb_opt) ILA_CUDA_CHECK(cudaMemcpyAsync(qtb_opt + xoff[i], S.d_qtb + S.xoff[j], sizeof(double) * nn, cudaMemcpyDeviceToHost, s));
            }
        }
    }
    for (int d = 0; d < G; d++) {
        ILA_CUDA_CHECK(cudaSetDevice(phys(d, G)));
        ILA_CUDA_CHECK(cudaStreamSynchronize(g_dev[phys(d, G)].stream));
    }
    guard.done = true;
    ILA_CUDA_CHECK(cudaSetDevice(g_base_dev));
    return 0;
}

int ila_qr_banded_solve_f64(int l, int u, int64_t batch, const double *p, const double *q, const double *r,
                            const double *shift, int hp, const double *head, int nb, const double *b, double tol,
                            int colgrowth, int64_t ncap, const int64_t *ncap_per, double *x, int64_t xcap,
                            int64_t *xoff, int64_t *ncols, int64_t *nconv, double *factors_opt, double *tau_opt,
                            double *qtb_opt, int32_t *status, int ngpus)
{
    return banded_solve_host(0, l, u, batch, p, q, r, shift, hp, head, nb, b, tol, colgrowth, ncap, ncap_per, x, xcap, xoff,
                             ncols, nconv, factors_opt, tau_opt, qtb_opt, status, ngpus);
}

// Complex eltype: p, q, shift, head, b complex; r real.  Contiguous shards; x ragged instance-major (xoff), complex.
int ila_qr_banded_solve_c64(int l, int u, int64_t batch, const ila_c64 *p, const ila_c64 *q, const double *r,
                            const ila_c64 *shift, int hp, const ila_c64 *head, int nb, const ila_c64 *b, double tol,
                            int colgrowth, int64_t ncap, const int64_t *ncap_per, ila_c64 *x, int64_t xcap, int64_t *xoff,
                            int64_t *ncols, int64_t *nconv, int32_t *status, int ngpus)
{
    if (batch < 0 || nb < 1 || colgrowth < 1 || hp < 0 || !p || !q || !b || !xoff || !ncols || !nconv || !status || (hp > 0 && !head))
        return set_error(ILA_ERR_ARG, "qr_banded_solve_c64: bad arguments");
    if (l < 1) return set_error(ILA_ERR_UNSUPPORTED, "qr_banded_solve_c64: l == 0 (triangular operator, Q = I) is not on the accelerated path");
    int G = 0;
    int rc = resolve_ngpus(ngpus, &G);
    if (rc) return rc;
    if (batch == 0) { xoff[0] = 0; return 0; }
    std::lock_guard<std::mutex> lock(g_mutex);
    HostGuard guard(G);
    if (G > batch) G = (int)batch;
    const int nd = l + u + 1, REC = qr_c64_record_chunks(l, u);
    const int64_t per = (batch + G - 1) / G;
    struct Shard {
        int64_t lo = 0, cnt = 0, total = 0;
        std::vector<int64_t> gbase, xoff;
        cplx *d_p, *d_q, *d_shift, *d_head, *d_b, *d_x;
        double *d_r;
        double2 *d_work;
        int64_t *d_gbase, *d_caps, *d_ncols, *d_nconv, *d_nswept, *d_xoff;
        int32_t *d_status;
    };
    std::vector<Shard> sh(G);
    for (int d = 0; d < G; d++) {
        Shard &S = sh[d];
        S.lo = d * per;
        S.cnt = (S.lo + per <= batch) ? per : batch - S.lo;
        if (S.cnt <= 0) { S.cnt = 0; continue; }
        const int64_t ng = (S.cnt + 31) / 32;
        S.gbase.assign(ng + 1, 0);
        for (int64_t g = 0; g < ng; g++) {
            int64_t mx = 0;
            for (int64_t j = g * 32; j < S.cnt && j < (g + 1) * 32; j++) {
                const int64_t c = ncap_per ? ncap_per[S.lo + j] : ncap;
                if (c < 1) return set_error(ILA_ERR_ARG, "qr_banded_solve_c64: non-positive capacity");
                mx = c > mx ? c : mx;
            }
            S.gbase[g + 1] = S.gbase[g] + mx;
        }
        const int pd = phys(d, G);
        if ((rc = dev_init(pd))) return rc;
        cudaStream_t s = g_dev[pd].stream;
        if ((rc = dev_buf_t(pd, 4, (size_t)S.cnt * nd, &S.d_p))) return rc;
        if ((rc = dev_buf_t(pd, 5, (size_t)S.cnt * nd, &S.d_q))) return rc;
        S.d_r = nullptr; S.d_shift = nullptr; S.d_head = nullptr; S.d_caps = nullptr;
        if (r && (rc = dev_buf_t(pd, 6, (size_t)S.cnt * nd, &S.d_r))) return rc;
        if (shift && (rc = dev_buf_t(pd, 7, (size_t)S.cnt, &S.d_shift))) return rc;
        if (hp > 0 && (rc = dev_buf_t(pd, 8, (size_t)nd * hp, &S.d_head))) return rc;
        if ((rc = dev_buf_t(pd, 9, (size_t)S.cnt * nb, &S.d_b))) return rc;
        if ((rc = dev_buf_t(pd, 10, S.gbase.size(), &S.d_gbase))) return rc;
        if (ncap_per && (rc = dev_buf_t(pd, 21, (size_t)S.cnt, &S.d_caps))) return rc;
        if ((rc = dev_buf_t(pd, 11, (size_t)S.cnt, &S.d_ncols))) return rc;
        if ((rc = dev_buf_t(pd, 12, (size_t)S.cnt, &S.d_nconv))) return rc;
        if ((rc = dev_buf_t(pd, 13, (size_t)S.cnt, &S.d_nswept))) return rc;
        if ((rc = dev_buf_t(pd, 14, (size_t)S.cnt + 1, &S.d_xoff))) return rc;
        if ((rc = dev_buf_t(pd, 15, (size_t)S.cnt, &S.d_status))) return rc;
        if ((rc = dev_buf_t(pd, 16, (size_t)S.gbase[ng] * REC * 32, &S.d_work))) return rc;
        ILA_CUDA_CHECK(cudaMemcpyAsync(S.d_p, p + S.lo * nd, sizeof(cplx) * S.cnt * nd, cudaMemcpyHostToDevice, s));
        ILA_CUDA_CHECK(cudaMemcpyAsync(S.d_q, q + S.lo * nd, sizeof(cplx) * S.cnt * nd, cudaMemcpyHostToDevice, s));
        if (r) ILA_CUDA_CHECK(cudaMemcpyAsync(S.d_r, r + S.lo * nd, sizeof(double) * S.cnt * nd, cudaMemcpyHostToDevice, s));
        if (shift) ILA_CUDA_CHECK(cudaMemcpyAsync(S.d_shift, shift + S.lo, sizeof(cplx) * S.cnt, cudaMemcpyHostToDevice, s));
        if (hp > 0) ILA_CUDA_CHECK(cudaMemcpyAsync(S.d_head, head, sizeof(cplx) * nd * hp, cudaMemcpyHostToDevice, s));
        ILA_CUDA_CHECK(cudaMemcpyAsync(S.d_b, b + S.lo * nb, sizeof(cplx) * S.cnt * nb, cudaMemcpyHostToDevice, s));
        ILA_CUDA_CHECK(cudaMemcpyAsync(S.d_gbase, S.gbase.data(), sizeof(int64_t) * S.gbase.size(), cudaMemcpyHostToDevice, s));
        if (ncap_per) ILA_CUDA_CHECK(cudaMemcpyAsync(S.d_caps, ncap_per + S.lo, sizeof(int64_t) * S.cnt, cudaMemcpyHostToDevice, s));
        QrC64FwdArgs a{S.cnt, S.d_p, S.d_q, S.d_r, S.d_shift, S.d_head, S.d_b, hp, nb, colgrowth, tol, S.d_gbase, S.d_caps, ncap,
                       S.d_work, S.d_ncols, S.d_nconv, S.d_nswept, S.d_status};
        if ((rc = launch_qr_forward_c64(l, u, a, s))) return rc;
        if ((rc = launch_qr_scan(S.d_ncols, S.cnt, S.d_xoff, s))) return rc;
        S.xoff.resize(S.cnt + 1);
        ILA_CUDA_CHECK(cudaMemcpyAsync(ncols + S.lo, S.d_ncols, sizeof(int64_t) * S.cnt, cudaMemcpyDeviceToHost, s));
        ILA_CUDA_CHECK(cudaMemcpyAsync(nconv + S.lo, S.d_nconv, sizeof(int64_t) * S.cnt, cudaMemcpyDeviceToHost, s));
        ILA_CUDA_CHECK(cudaMemcpyAsync(status + S.lo, S.d_status, sizeof(int32_t) * S.cnt, cudaMemcpyDeviceToHost, s));
        ILA_CUDA_CHECK(cudaMemcpyAsync(S.xoff.data(), S.d_xoff, sizeof(int64_t) * (S.cnt + 1), cudaMemcpyDeviceToHost, s));
    }
    for (int d = 0; d < G; d++) {
        if (sh[d].cnt == 0) continue;
        ILA_CUDA_CHECK(cudaSetDevice(phys(d, G)));
        ILA_CUDA_CHECK(cudaStreamSynchronize(g_dev[phys(d, G)].stream));
    }
    int64_t acc = 0;
    for (int d = 0; d < G; d++) {
        Shard &S = sh[d];
        for (int64_t j = 0; j < S.cnt; j++) xoff[S.lo + j] = acc + S.xoff[j];
        S.total = S.cnt ? S.xoff[S.cnt] : 0;
        acc += S.total;
    }
    xoff[batch] = acc;
    if (acc > xcap || !x) {
        guard.done = true;
    ILA_CUDA_CHECK(cudaSetDevice(g_base_dev));
        return set_error(ILA_ERR_CAPACITY, "qr_banded_solve_c64: x capacity too small; xoff[batch] holds the required number of scalars");
    }
    for (int d = 0; d < G; d++) {
        Shard &S = sh[d];
        if (S.cnt == 0) continue;
        const int pd = phys(d, G);
        ILA_CUDA_CHECK(cudaSetDevice(pd));
        cudaStream_t s = g_dev[pd].stream;
        if ((rc = dev_buf_t(pd, 17, (size_t)S.total, &S.d_x))) return rc;
        if ((rc = launch_qr_backsolve_c64(l, u, S.cnt, S.d_gbase, S.d_work, S.d_ncols, S.d_nswept, S.d_xoff, (double2 *)S.d_x, s))) return rc;
        if (S.total > 0)
            ILA_CUDA_CHECK(cudaMemcpyAsync(x + xoff[S.lo], S.d_x, sizeof(cplx) * S.total, cudaMemcpyDeviceToHost, s));
    }
    for (int d = 0; d < G; d++) {
        if (sh[d].cnt == 0) continue;
        ILA_CUDA_CHECK(cudaSetDevice(phys(d, G)));
        ILA_CUDA_CHECK(cudaStreamSynchronize(g_dev[phys(d, G)].stream));
    }
    guard.done = true;
    ILA_CUDA_CHECK(cudaSetDevice(g_base_dev));
    return 0;
}

int ila_chol_banded_solve_f64(int u, int64_t batch, const double *p, const double *q, const double *r,
                              const double *shift, int hp, const double *head, int nb, const double *b, double tol,
                              int colgrowth, int64_t ncap, const int64_t *ncap_per, double *x, int64_t xcap,
                              int64_t *xoff, int64_t *ncols, int64_t *nconv, double *factors_opt, double *y_opt,
                              int32_t *status, int ngpus)
{
    if (u < 1) return set_error(ILA_ERR_UNSUPPORTED, "chol_banded_solve: u == 0 (diagonal operator) is not on the accelerated path");
    return banded_solve_host(1, 0, u, batch, p, q, r, shift, hp, head, nb, b, tol, colgrowth, ncap, ncap_per, x, xcap, xoff,
                             ncols, nconv, factors_opt, nullptr, y_opt, status, ngpus);
}

// K5 -> K10 fused behind the device-resident factor records: cholesky(S_i).U never leaves the device.
// The Cholesky sweep is driven through its adaptive entry with b = e1 and an infinite tolerance, so that the convergence test
// passes at the first opportunity (column colgrowth+1) and full_sweep factorizes 2*colgrowth + u >= n + 3 columns; the
// blocking of partialcholesky! is then two blocks of colgrowth (+u) columns (upstream the partition depends on the
// caller's access history: 100 columns at construction, then up to n+3 — tridiagonalconjugation.jl:244-245, 276).
int ila_chol_triconj_f64(int u, int64_t batch, const double *p, const double *q, const double *r, const double *shift,
                         int hp, const double *head, int64_t n, const double *X, int64_t xld, int x_shared, double *Y,
                         int32_t *status, int ngpus)
{
    if (batch < 0 || n < 1 || hp < 0 || !p || !q || !X || xld < n + 1 || (hp > 0 && !head) || (batch > 0 && (!Y || !status)))
        return set_error(ILA_ERR_ARG, "chol_triconj: bad arguments (X bands need n+1 entries)");
    if (u < 1 || u > 4) return set_error(ILA_ERR_UNSUPPORTED, "chol_triconj: upper bandwidth u must be in 1..4");
    int G = 0;
    int rc = resolve_ngpus(ngpus, &G);
    if (rc) return rc;
    if (batch == 0) return 0;
    std::lock_guard<std::mutex> lock(g_mutex);
    HostGuard guard(G);
    if (G > batch) G = (int)batch;
    const int nd = u + 1;
    const int64_t cg = (n + 4) / 2, ncap = 2 * cg + 3 * u + 2;
    const int nch = qr_record_doubles(0, u, false) / 2;
    const double one = 1.0;
    const int64_t per = (batch + G - 1) / G;
    for (int d = 0; d < G; d++) {
        const int64_t lo = d * per, cnt = (lo + per <= batch ? per : batch - lo);
        if (cnt <= 0) continue;
        const int pd = phys(d, G);
        if ((rc = dev_init(pd))) return rc;
        cudaStream_t s = g_dev[pd].stream;
        std::vector<int64_t> gbase((cnt + 31) / 32 + 1);
        int64_t work_bytes = 0, xi_bytes = 0;
        ila_qr_banded_work_layout(0, u, cnt, ncap, nullptr, gbase.data(), &work_bytes, &xi_bytes);
        std::vector<double> bvec((size_t)cnt, one);
        double *d_p, *d_q, *d_r = nullptr, *d_shift = nullptr, *d_head = nullptr, *d_b, *d_X, *d_Y;
        int64_t *d_gbase, *d_ncols, *d_nconv, *d_nswept, *d_xoff;
        int32_t *d_status;
        char *d_work;
        const size_t xcount = (size_t)(x_shared ? 1 : cnt) * 3 * xld;
        if ((rc = dev_buf_t(pd, 4, (size_t)cnt * nd, &d_p))) return rc;
        if ((rc = dev_buf_t(pd, 5, (size_t)cnt * nd, &d_q))) return rc;
        if (r && (rc = dev_buf_t(pd, 6, (size_t)cnt * nd, &d_r))) return rc;
        if (shift && (rc = dev_buf_t(pd, 7, (size_t)cnt, &d_shift))) return rc;
        if (hp > 0 && (rc = dev_buf_t(pd, 8, (size_t)nd * hp, &d_head))) return rc;
        if ((rc = dev_buf_t(pd, 9, (size_t)cnt, &d_b))) return rc;
        if ((rc = dev_buf_t(pd, 10, gbase.size(), &d_gbase))) return rc;
        if ((rc = dev_buf_t(pd, 11, (size_t)cnt, &d_ncols))) return rc;
        if ((rc = dev_buf_t(pd, 12, (size_t)cnt, &d_nconv))) return rc;
        if ((rc = dev_buf_t(pd, 13, (size_t)cnt, &d_nswept))) return rc;
        if ((rc = dev_buf_t(pd, 14, (size_t)cnt + 1, &d_xoff))) return rc;
        if ((rc = dev_buf_t(pd, 15, (size_t)cnt, &d_status))) return rc;
        if ((rc = dev_buf_t(pd, 16, (size_t)work_bytes, &d_work))) return rc;
        if ((rc = dev_buf_t(pd, 1, xcount, &d_X))) return rc;
        if ((rc = dev_buf_t(pd, 2, (size_t)cnt * 3 * n, &d_Y))) return rc;
        ILA_CUDA_CHECK(cudaMemcpyAsync(d_p, p + lo * nd, sizeof(double) * cnt * nd, cudaMemcpyHostToDevice, s));
        ILA_CUDA_CHECK(cudaMemcpyAsync(d_q, q + lo * nd, sizeof(double) * cnt * nd, cudaMemcpyHostToDevice, s));
        if (r) ILA_CUDA_CHECK(cudaMemcpyAsync(d_r, r + lo * nd, sizeof(double) * cnt * nd, cudaMemcpyHostToDevice, s));
        if (shift) ILA_CUDA_CHECK(cudaMemcpyAsync(d_shift, shift + lo, sizeof(double) * cnt, cudaMemcpyHostToDevice, s));
        if (hp > 0) ILA_CUDA_CHECK(cudaMemcpyAsync(d_head, head, sizeof(double) * nd * hp, cudaMemcpyHostToDevice, s));
        ILA_CUDA_CHECK(cudaMemcpyAsync(d_b, bvec.data(), sizeof(double) * cnt, cudaMemcpyHostToDevice, s));
        ILA_CUDA_CHECK(cudaMemcpyAsync(d_gbase, gbase.data(), sizeof(int64_t) * gbase.size(), cudaMemcpyHostToDevice, s));
        ILA_CUDA_CHECK(cudaMemcpyAsync(d_X, X + (x_shared ? 0 : lo * 3 * xld), sizeof(double) * xcount, cudaMemcpyHostToDevice, s));
        ILA_CUDA_CHECK(cudaStreamSynchronize(s));     // bvec / gbase are stack-scoped host buffers
        if ((rc = ila_chol_banded_factor_f64_dev(u, cnt, d_p, d_q, d_r, d_shift, hp, d_head, 1, d_b, HUGE_VAL, (int)cg, ncap, nullptr,
                                                 d_gbase, d_work, d_ncols, d_nconv, d_nswept, d_xoff, d_status, 1, s))) return rc;
        if ((rc = ila_triconj_records_f64_dev(u + 1 > 3 ? 3 : u + 1, nch, cnt, n, d_gbase, d_work, d_status, d_X, xld, x_shared, d_Y, s))) return rc;
        ILA_CUDA_CHECK(cudaMemcpyAsync(Y + lo * 3 * n, d_Y, sizeof(double) * cnt * 3 * n, cudaMemcpyDeviceToHost, s));
        ILA_CUDA_CHECK(cudaMemcpyAsync(status + lo, d_status, sizeof(int32_t) * cnt, cudaMemcpyDeviceToHost, s));
    }
    for (int d = 0; d < G; d++) {
        const int pd = phys(d, G);
        if (!g_dev[pd].init) continue;
        ILA_CUDA_CHECK(cudaSetDevice(pd));
        ILA_CUDA_CHECK(cudaStreamSynchronize(g_dev[pd].stream));
    }
    guard.done = true;
    ILA_CUDA_CHECK(cudaSetDevice(g_base_dev));
    return 0;
}

} // extern "C"
