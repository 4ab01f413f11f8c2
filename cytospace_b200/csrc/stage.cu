// Host -> device upload of PAGEABLE memory (the numpy arrays a CytoSPACE user holds: main_cytospace hands
// `scRNA_norm_np` / `st_norm_np` to the solver as plain float64 arrays, cytospace/cytospace.py:398-443) through a ring of
// pinned pieces: worker threads copy pageable -> pinned with non-temporal stores (no read-for-ownership of the ring, which
// is written once and read once by the DMA engine) and each worker enqueues the DMA of its own piece, so the copy engine
// sees a steady queue of 8 MB transfers while the other workers are still filling.  cudaMemcpy on pageable memory stages
// through a single-threaded driver path (11 GB/s measured); from pinned memory the box delivers 51.8 GB/s, this ring 50.8
// with 8 workers and 8 MB pieces (profiles/r02_stage_sweep.txt; the Python thread-pool ring it replaces: 40 GB/s).
//
// One stager per process (lazily built, sized by CYB_STAGE_THREADS / CYB_STAGE_PIECE_MB / CYB_STAGE_PIECES); calls are
// serialised by a mutex.  The call returns when every piece is ENQUEUED on the internal copy stream and `stream` has been
// made to wait for it: asynchronous towards the device like the rest of the ABI, and the source may be reused on return.
#include "common.h"

#include <immintrin.h>

#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

namespace {

void copy_nontemporal(char *dst, const char *src, size_t n) {
    // dst is 64-byte aligned by construction (piece starts); src is whatever numpy gave us
    size_t lines = n / 64;
    for (size_t i = 0; i < lines; ++i) {
        const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src));
        const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + 16));
        const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + 32));
        const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + 48));
        _mm_stream_si128(reinterpret_cast<__m128i *>(dst), a);
        _mm_stream_si128(reinterpret_cast<__m128i *>(dst + 16), b);
        _mm_stream_si128(reinterpret_cast<__m128i *>(dst + 32), c);
        _mm_stream_si128(reinterpret_cast<__m128i *>(dst + 48), d);
        src += 64; dst += 64;
    }
    _mm_sfence();
    if (n & 63) memcpy(dst, src, n & 63);
}

// dst[i] = (float) src[i], non-temporal stores; returns false when some value does not survive the round trip exactly
// (dst is 64-byte aligned: piece starts).  AVX2 where the host has it (8 values per step), else SSE2.
__attribute__((target("avx2"))) bool narrow_nontemporal_avx2(float *dst, const double *src, size_t n) {
    __m256d ok = _mm256_castsi256_pd(_mm256_set1_epi32(-1));
    const size_t oct = n / 8;
    for (size_t i = 0; i < oct; ++i) {
        const __m256d a = _mm256_loadu_pd(src), b = _mm256_loadu_pd(src + 4);
        const __m128 fa = _mm256_cvtpd_ps(a), fb = _mm256_cvtpd_ps(b);
        _mm256_stream_ps(dst, _mm256_set_m128(fb, fa));
        ok = _mm256_and_pd(ok, _mm256_and_pd(_mm256_cmp_pd(_mm256_cvtps_pd(fa), a, _CMP_EQ_OQ),
                                             _mm256_cmp_pd(_mm256_cvtps_pd(fb), b, _CMP_EQ_OQ)));
        src += 8; dst += 8;
    }
    _mm_sfence();
    bool exact = _mm256_movemask_pd(ok) == 15;
    for (size_t i = 0; i < (n & 7); ++i) {
        dst[i] = (float)src[i];
        exact = exact && (double)dst[i] == src[i];
    }
    return exact;
}

bool narrow_nontemporal(float *dst, const double *src, size_t n) {
    static const bool have_avx2 = __builtin_cpu_supports("avx2");
    if (have_avx2) return narrow_nontemporal_avx2(dst, src, n);
    __m128d ok = _mm_castsi128_pd(_mm_set1_epi32(-1));
    const size_t quads = n / 4;
    for (size_t i = 0; i < quads; ++i) {
        const __m128d a = _mm_loadu_pd(src), b = _mm_loadu_pd(src + 2);
        const __m128 fa = _mm_cvtpd_ps(a), fb = _mm_cvtpd_ps(b);                 // MXCSR rounding: nearest even, like numpy's astype
        _mm_stream_ps(dst, _mm_movelh_ps(fa, fb));
        ok = _mm_and_pd(ok, _mm_and_pd(_mm_cmpeq_pd(_mm_cvtps_pd(fa), a), _mm_cmpeq_pd(_mm_cvtps_pd(fb), b)));
        src += 4; dst += 4;
    }
    _mm_sfence();
    bool exact = _mm_movemask_pd(ok) == 3;
    for (size_t i = 0; i < (n & 3); ++i) {
        dst[i] = (float)src[i];
        exact = exact && (double)dst[i] == src[i];
    }
    return exact;
}

struct Stager {
    int device = -1, n_threads = 0, n_pieces = 0;
    size_t piece = 0;
    char *ring = nullptr;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t fence = nullptr;
    std::vector<cudaEvent_t> drained;             // per ring slot: the DMA that last read it
    std::vector<char> used;                       // slot has a recorded event
    std::vector<std::thread> workers;
    std::mutex m;
    std::condition_variable cv_work, cv_done, cv_slot;
    // the job in flight
    const char *src = nullptr;
    char *dst = nullptr;
    size_t bytes = 0, n_jobs = 0, next = 0, finished = 0;      // bytes: of the DEVICE buffer (= pinned bytes)
    int narrow = 0;                                            // 1: src holds float64, the pieces and dst hold float32
    int error = 0, inexact = 0;
    bool quit = false;

    void run() {
        cudaSetDevice(device);
        std::unique_lock<std::mutex> lk(m);
        for (;;) {
            cv_work.wait(lk, [&] { return quit || next < n_jobs; });
            if (quit) return;
            const size_t p = next++;
            // a ring slot is refilled only after the worker of piece p - n_pieces has enqueued its DMA (host side, here) and
            // that DMA has drained (device side, the slot's event below); pieces are handed out in order, so no cycle
            if (p >= (size_t)n_pieces) cv_slot.wait(lk, [&] { return done_flag[p - n_pieces] != 0; });
            lk.unlock();
            const int slot = (int)(p % (size_t)n_pieces);
            const size_t lo = p * piece, len = std::min(piece, bytes - lo);
            char *buf = ring + (size_t)slot * piece;
            bool ok = !used[slot] || cudaEventSynchronize(drained[slot]) == cudaSuccess;
            bool exact = true;
            if (ok) {
                if (narrow) exact = narrow_nontemporal(reinterpret_cast<float *>(buf), reinterpret_cast<const double *>(src + 2 * lo), len / 4);
                else copy_nontemporal(buf, src + lo, len);
                ok = cudaMemcpyAsync(dst + lo, buf, len, cudaMemcpyHostToDevice, copy_stream) == cudaSuccess &&
                     cudaEventRecord(drained[slot], copy_stream) == cudaSuccess;
            }
            lk.lock();
            used[slot] = 1;
            done_flag[p] = 1;
            if (!ok) error = 1;
            if (!exact) inexact = 1;
            ++finished;
            cv_slot.notify_all();
            if (finished == n_jobs) cv_done.notify_all();
        }
    }
    std::vector<char> done_flag;                  // per piece of the current job: DMA enqueued
};

Stager *g_stager = nullptr;
std::mutex g_stager_mutex;

int env_int(const char *name, int dflt, int lo, int hi) {
    const char *e = getenv(name);
    int v = e ? atoi(e) : dflt;
    return std::max(lo, std::min(hi, v));
}

}  // namespace

namespace {
int stage_upload(const void *src_host, void *dst_dev, size_t bytes, int narrow, int32_t *inexact_host, cudaStream_t stream) {
    std::lock_guard<std::mutex> guard(g_stager_mutex);
    int dev = 0;
    CYB_CUDA_CHECK(cudaGetDevice(&dev));
    if (g_stager && g_stager->device != dev) {
        // one ring per process: wait for the old device's transfers and move the ring's stream / events over
        Stager *s = g_stager;
        cudaSetDevice(s->device);
        cudaStreamSynchronize(s->copy_stream);
        {
            std::lock_guard<std::mutex> lk(s->m);
            s->quit = true;
        }
        s->cv_work.notify_all();
        for (auto &t : s->workers) t.join();
        for (auto &e : s->drained) cudaEventDestroy(e);
        cudaEventDestroy(s->fence);
        cudaStreamDestroy(s->copy_stream);
        cudaFreeHost(s->ring);
        delete s;
        g_stager = nullptr;
        CYB_CUDA_CHECK(cudaSetDevice(dev));
    }
    if (!g_stager) {
        Stager *s = new Stager();
        s->device = dev;
        const int hw = (int)std::thread::hardware_concurrency();
        s->n_threads = env_int("CYB_STAGE_THREADS", std::max(1, std::min(16, hw > 0 ? hw : 8)), 1, 64);
        s->piece = (size_t)env_int("CYB_STAGE_PIECE_MB", 8, 1, 256) << 20;
        s->n_pieces = env_int("CYB_STAGE_PIECES", 3 * s->n_threads, 2, 512);
        const size_t ring_bytes = s->piece * (size_t)s->n_pieces;
        if (cudaHostAlloc(reinterpret_cast<void **>(&s->ring), ring_bytes, cudaHostAllocDefault) != cudaSuccess) {
            cudaGetLastError();
            delete s;
            return cyb::set_error(CYB_ERR_CUDA, "cyb_stage_upload: cannot pin %zu bytes of host memory", ring_bytes);
        }
        CYB_CUDA_CHECK(cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
        CYB_CUDA_CHECK(cudaEventCreateWithFlags(&s->fence, cudaEventDisableTiming));
        s->drained.resize(s->n_pieces);
        s->used.assign(s->n_pieces, 0);
        for (auto &e : s->drained) CYB_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        for (int t = 0; t < s->n_threads; ++t) s->workers.emplace_back([s] { s->run(); });
        g_stager = s;
    }
    Stager *s = g_stager;
    // the copies may not overtake work already queued on the caller's stream that still uses dst
    CYB_CUDA_CHECK(cudaEventRecord(s->fence, stream));
    CYB_CUDA_CHECK(cudaStreamWaitEvent(s->copy_stream, s->fence, 0));
    int err;
    {
        std::unique_lock<std::mutex> lk(s->m);
        s->src = static_cast<const char *>(src_host);
        s->dst = static_cast<char *>(dst_dev);
        s->bytes = bytes;
        s->narrow = narrow;
        s->n_jobs = (bytes + s->piece - 1) / s->piece;
        s->next = 0; s->finished = 0; s->error = 0; s->inexact = 0;
        s->done_flag.assign(s->n_jobs, 0);
        s->cv_work.notify_all();
        s->cv_done.wait(lk, [&] { return s->finished == s->n_jobs; });
        err = s->error;
        if (inexact_host) *inexact_host = s->inexact;
    }
    if (err) {
        cudaGetLastError();
        return cyb::set_error(CYB_ERR_CUDA, "cyb_stage_upload: a piece transfer failed");
    }
    CYB_CUDA_CHECK(cudaEventRecord(s->fence, s->copy_stream));
    CYB_CUDA_CHECK(cudaStreamWaitEvent(stream, s->fence, 0));
    return CYB_OK;
}
}  // namespace

extern "C" int cyb_stage_upload(const void *src_host, void *dst_dev, size_t bytes, void *stream_v) {
    if (bytes == 0) return CYB_OK;
    if (!src_host || !dst_dev) return cyb::set_error(CYB_ERR_INVALID, "cyb_stage_upload: null pointer argument");
    return stage_upload(src_host, dst_dev, bytes, 0, nullptr, static_cast<cudaStream_t>(stream_v));
}

extern "C" int cyb_stage_upload_f64_as_f32(const double *src_host, float *dst_dev, size_t n, int32_t *inexact_host,
                                           void *stream_v) {
    if (inexact_host) *inexact_host = 0;
    if (n == 0) return CYB_OK;
    if (!src_host || !dst_dev) return cyb::set_error(CYB_ERR_INVALID, "cyb_stage_upload_f64_as_f32: null pointer argument");
    return stage_upload(src_host, dst_dev, n * sizeof(float), 1, inexact_host, static_cast<cudaStream_t>(stream_v));
}
