// Multi-GPU data plane of the chunked assignment: NCCL point-to-point / broadcast /
// all-gather of expression blocks and assignment indices behind the C ABI, plus the device column gather that
// cuts a chunk's columns out of the expression matrix.
//
// Replaces the process-pool fan-out of apply_linear_assignment, cytospace/cytospace.py:430-467: the reference
// pickles `scRNA_norm_np[:, index_sc_list[i]]` / `st_norm_np[:, index_st_list[i]]` (:434-443) to one worker
// process per chunk and collects the per-chunk index lists as they complete (:453-467).  Here rank 0 gathers the
// columns on its GPU (cyb_gather_columns), the blocks travel over NVLink (cyb_dist_send / cyb_dist_recv; the ST
// block every --sampling-sub-spots chunk shares, :438, with ONE cyb_dist_broadcast) and the indices come back with
// one cyb_dist_all_gather.  Nothing is exchanged during a solve.
//
// NCCL is bound at run time (dlopen of libnccl.so.2 -- inside a PyTorch process that is the copy torch already
// loaded), so the library has no link-time dependency on it; without NCCL every cyb_dist_* call returns
// CYB_ERR_UNSUPPORTED.

#include <dlfcn.h>
#include <nccl.h>

#include <cstdint>
#include <cstring>
#include <mutex>

#include "common.h"

namespace {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId *);
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
    const char *(*GetErrorString)(ncclResult_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    bool ok = false;
};

NcclApi &api() {
    static NcclApi a;
    static std::once_flag once;
    std::call_once(once, [] {
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return;
#define CYB_SYM(field, name) a.field = reinterpret_cast<decltype(a.field)>(dlsym(h, name))
        CYB_SYM(GetUniqueId, "ncclGetUniqueId");
        CYB_SYM(CommInitRank, "ncclCommInitRank");
        CYB_SYM(CommDestroy, "ncclCommDestroy");
        CYB_SYM(Broadcast, "ncclBroadcast");
        CYB_SYM(Send, "ncclSend");
        CYB_SYM(Recv, "ncclRecv");
        CYB_SYM(AllGather, "ncclAllGather");
        CYB_SYM(GetErrorString, "ncclGetErrorString");
        CYB_SYM(GroupStart, "ncclGroupStart");
        CYB_SYM(GroupEnd, "ncclGroupEnd");
#undef CYB_SYM
        a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.Broadcast && a.Send && a.Recv && a.AllGather &&
               a.GetErrorString && a.GroupStart && a.GroupEnd;
    });
    return a;
}

#define CYB_NCCL_READY()                                                                          \
    do {                                                                                          \
        if (!api().ok) return cyb::set_error(CYB_ERR_UNSUPPORTED, "libnccl.so.2 could not be loaded"); \
    } while (0)
#define CYB_NCCL_CHECK(expr)                                                                      \
    do {                                                                                          \
        ncclResult_t _r = (expr);                                                                 \
        if (_r != ncclSuccess)                                                                    \
            return cyb::set_error(CYB_ERR_CUDA, "%s failed: %s", #expr, api().GetErrorString(_r)); \
    } while (0)

template <typename T>
__global__ void gather_columns_kernel(const T *__restrict__ x, long long ld_x, long long n_rows,
                                      const int32_t *__restrict__ cols, long long n_out, T *__restrict__ out,
                                      long long ld_out) {
    // one row of the output per blockIdx.y; coalesced writes, gathered reads (a row of x is contiguous: the
    // reads of one warp fall into as many sectors as the selected columns span)
    const long long r = blockIdx.y;
    for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < n_out; j += (long long)gridDim.x * blockDim.x)
        out[r * ld_out + j] = x[r * ld_x + cols[j]];
}

// float64 -> float32 with an exactness flag: one pass, 12 bytes of traffic per element.
__global__ void narrow_kernel(const double *__restrict__ x, float *__restrict__ out, long long n, int *__restrict__ inexact) {
    int bad = 0;
    const long long n2 = n >> 1;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x) {
        const double2 v = __ldcs(reinterpret_cast<const double2 *>(x) + i);
        const float2 f = make_float2((float)v.x, (float)v.y);
        bad |= ((double)f.x != v.x) | ((double)f.y != v.y);
        __stcs(reinterpret_cast<float2 *>(out) + i, f);
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const double v = x[n - 1];
        out[n - 1] = (float)v;
        bad |= ((double)(float)v != v);
    }
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(inexact, 1);
}

}  // namespace

static_assert(sizeof(ncclUniqueId) == CYB_DIST_ID_BYTES, "ncclUniqueId size");

extern "C" int cyb_dist_unique_id(void *id_out) {
    CYB_NCCL_READY();
    if (!id_out) return cyb::set_error(CYB_ERR_INVALID, "cyb_dist_unique_id: null pointer");
    ncclUniqueId id;
    CYB_NCCL_CHECK(api().GetUniqueId(&id));
    memcpy(id_out, &id, sizeof(id));
    return CYB_OK;
}

extern "C" int cyb_dist_init(const void *id_bytes, int n_ranks, int rank, void **comm_out) {
    CYB_NCCL_READY();
    if (!id_bytes || !comm_out || n_ranks < 1 || rank < 0 || rank >= n_ranks)
        return cyb::set_error(CYB_ERR_INVALID, "cyb_dist_init: bad arguments (ranks=%d rank=%d)", n_ranks, rank);
    ncclUniqueId id;
    memcpy(&id, id_bytes, sizeof(id));
    ncclComm_t comm = nullptr;
    CYB_NCCL_CHECK(api().CommInitRank(&comm, n_ranks, id, rank));
    *comm_out = comm;
    return CYB_OK;
}

extern "C" int cyb_dist_destroy(void *comm) {
    CYB_NCCL_READY();
    if (comm) CYB_NCCL_CHECK(api().CommDestroy(static_cast<ncclComm_t>(comm)));
    return CYB_OK;
}

extern "C" int cyb_dist_broadcast(void *comm, void *buf_dev, size_t bytes, int root, void *stream) {
    CYB_NCCL_READY();
    if (!comm || (!buf_dev && bytes)) return cyb::set_error(CYB_ERR_INVALID, "cyb_dist_broadcast: null pointer");
    CYB_NCCL_CHECK(api().Broadcast(buf_dev, buf_dev, bytes, ncclUint8, root, static_cast<ncclComm_t>(comm),
                                   static_cast<cudaStream_t>(stream)));
    return CYB_OK;
}

extern "C" int cyb_dist_send(void *comm, const void *buf_dev, size_t bytes, int peer, void *stream) {
    CYB_NCCL_READY();
    if (!comm || (!buf_dev && bytes)) return cyb::set_error(CYB_ERR_INVALID, "cyb_dist_send: null pointer");
    CYB_NCCL_CHECK(api().Send(buf_dev, bytes, ncclUint8, peer, static_cast<ncclComm_t>(comm), static_cast<cudaStream_t>(stream)));
    return CYB_OK;
}

extern "C" int cyb_dist_recv(void *comm, void *buf_dev, size_t bytes, int peer, void *stream) {
    CYB_NCCL_READY();
    if (!comm || (!buf_dev && bytes)) return cyb::set_error(CYB_ERR_INVALID, "cyb_dist_recv: null pointer");
    CYB_NCCL_CHECK(api().Recv(buf_dev, bytes, ncclUint8, peer, static_cast<ncclComm_t>(comm), static_cast<cudaStream_t>(stream)));
    return CYB_OK;
}

extern "C" int cyb_dist_group_start(void) {
    CYB_NCCL_READY();
    CYB_NCCL_CHECK(api().GroupStart());
    return CYB_OK;
}

extern "C" int cyb_dist_group_end(void) {
    CYB_NCCL_READY();
    CYB_NCCL_CHECK(api().GroupEnd());
    return CYB_OK;
}

extern "C" int cyb_dist_all_gather(void *comm, const void *send_dev, void *recv_dev, size_t bytes_per_rank, void *stream) {
    CYB_NCCL_READY();
    if (!comm || !send_dev || !recv_dev) return cyb::set_error(CYB_ERR_INVALID, "cyb_dist_all_gather: null pointer");
    CYB_NCCL_CHECK(api().AllGather(send_dev, recv_dev, bytes_per_rank, ncclUint8, static_cast<ncclComm_t>(comm),
                                   static_cast<cudaStream_t>(stream)));
    return CYB_OK;
}

extern "C" int cyb_gather_columns(const void *x_dev, int x_dtype, int64_t n_rows, int64_t ld_x, const int32_t *cols_dev,
                                  int64_t n_cols_out, void *out_dev, int64_t ld_out, void *stream_v) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    if (!x_dev || !cols_dev || !out_dev) return cyb::set_error(CYB_ERR_INVALID, "cyb_gather_columns: null pointer");
    if (n_rows <= 0 || n_cols_out <= 0 || ld_out < n_cols_out || n_rows > 65535 * 64ll)
        return cyb::set_error(CYB_ERR_INVALID, "cyb_gather_columns: bad shape (%lld x %lld)", (long long)n_rows, (long long)n_cols_out);
    if (x_dtype != CYB_F64 && x_dtype != CYB_F32) return cyb::set_error(CYB_ERR_INVALID, "cyb_gather_columns: bad dtype");
    const int threads = 256;
    const int bx = (int)std::min<long long>((n_cols_out + threads - 1) / threads, 64);
    // rows go to blockIdx.y in slabs of 65535 (the grid limit)
    for (int64_t r0 = 0; r0 < n_rows; r0 += 65535) {
        const int rows = (int)std::min<int64_t>(65535, n_rows - r0);
        if (x_dtype == CYB_F64)
            gather_columns_kernel<double><<<dim3(bx, rows), threads, 0, stream>>>(
                static_cast<const double *>(x_dev) + r0 * ld_x, ld_x, rows, cols_dev, n_cols_out,
                static_cast<double *>(out_dev) + r0 * ld_out, ld_out);
        else
            gather_columns_kernel<float><<<dim3(bx, rows), threads, 0, stream>>>(
                static_cast<const float *>(x_dev) + r0 * ld_x, ld_x, rows, cols_dev, n_cols_out,
                static_cast<float *>(out_dev) + r0 * ld_out, ld_out);
        CYB_CUDA_CHECK(cudaGetLastError());
    }
    return CYB_OK;
}

extern "C" int cyb_narrow_f64_to_f32(const double *x_dev, int64_t n, float *out_dev, int32_t *inexact_dev, void *stream_v) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    if (!x_dev || !out_dev || !inexact_dev || n <= 0) return cyb::set_error(CYB_ERR_INVALID, "cyb_narrow_f64_to_f32: bad arguments");
    if ((reinterpret_cast<uintptr_t>(x_dev) & 15) || (reinterpret_cast<uintptr_t>(out_dev) & 7))
        return cyb::set_error(CYB_ERR_INVALID, "cyb_narrow_f64_to_f32: x must be 16-byte, out 8-byte aligned");
    int dev = 0, sms = 0;
    CYB_CUDA_CHECK(cudaGetDevice(&dev));
    CYB_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    narrow_kernel<<<sms * 8, 512, 0, stream>>>(x_dev, out_dev, (long long)n, inexact_dev);
    CYB_CUDA_CHECK(cudaGetLastError());
    return CYB_OK;
}
