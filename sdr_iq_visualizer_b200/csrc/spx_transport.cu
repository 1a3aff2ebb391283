// spx_transport.cu -- data movement around the kernels: device/peer copies, the measured host-copy ceiling of the
// box, and the C-level collective step of the sharded capture:
//   * spx_peer_reduce / spx_peer_push_rows: the fused path's last hop over CUDA-IPC mapped peer memory (NVLink), callable
//     from plain C so that a C consumer can finish config 5 without Python;
//   * spx_nccl_*: the plain NCCL all-reduce (Welch SUM f64, max-hold MAX f32, frame count SUM i64) and the gather of
//     uint8 rows to one rank.  libnccl is loaded with dlopen at first use (no link-time dependency: single-GPU users
//     never need it); torch's bundled libnccl.so.2 or the system one both work.
// The reference is single-process and has no collective (app/sdr/streamer.py:58 starts one thread); this is new
// capability defined by SURVEY.md 8(e).
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <mutex>

#include "spx_plan.h"

using namespace spx;

// ------------------------------------------------------------------ minimal NCCL surface (ABI-stable subset of nccl.h)
namespace {
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclInt8 = 0, ncclUint8 = 1, ncclInt64 = 4, ncclFloat32 = 7, ncclFloat64 = 8 };
enum { ncclSum = 0, ncclMax = 2 };

struct NcclApi {
    void* h = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
std::mutex g_nccl_mu;

int nccl_load() {
    std::lock_guard<std::mutex> g(g_nccl_mu);
    if (g_nccl.h) return SPX_OK;
    const char* names[] = {getenv("SPX_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names) {
        if (!n || !*n) continue;
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) return spx_set_error(SPX_E_UNSUPPORTED, "libnccl not found (set SPX_NCCL_LIB to its path): %s", dlerror());
    NcclApi a;
    a.h = h;
#define SPX_SYM(field, name)                                                                 \
    *(void**)(&a.field) = dlsym(h, name);                                                    \
    if (!a.field) { dlclose(h); return spx_set_error(SPX_E_UNSUPPORTED, "libnccl lacks %s", name); }
    SPX_SYM(GetUniqueId, "ncclGetUniqueId")
    SPX_SYM(CommInitRank, "ncclCommInitRank")
    SPX_SYM(CommDestroy, "ncclCommDestroy")
    SPX_SYM(AllReduce, "ncclAllReduce")
    SPX_SYM(Send, "ncclSend")
    SPX_SYM(Recv, "ncclRecv")
    SPX_SYM(GroupStart, "ncclGroupStart")
    SPX_SYM(GroupEnd, "ncclGroupEnd")
    SPX_SYM(GetErrorString, "ncclGetErrorString")
#undef SPX_SYM
    g_nccl = a;
    return SPX_OK;
}

#define SPX_NCCL(expr)                                                                                        \
    do {                                                                                                      \
        int _r = (expr);                                                                                      \
        if (_r != ncclSuccess)                                                                                \
            return spx_set_error(SPX_E_CUDA, "%s failed: %s", #expr, g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "?"); \
    } while (0)
}  // namespace

struct spx_comm {
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1, device = 0;
    long long* d_count = nullptr;
};

extern "C" {

// ------------------------------------------------------------------ copies
int spx_memcpy_d2d_async(int device, void* dst, const void* src, size_t bytes, void* stream) {
    SPX_CUDA(cudaSetDevice(device));
    // cudaMemcpyDefault: either side may be another GPU's memory mapped with spx_ipc_open (the copy engine then
    // moves it over NVLink)
    SPX_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, (cudaStream_t)stream));
    return SPX_OK;
}

int spx_stream_create(int device, void** stream_out) {
    if (!stream_out) return spx_set_error(SPX_E_INVALID, "stream_out is NULL");
    SPX_CUDA(cudaSetDevice(device));
    cudaStream_t s = nullptr;
    SPX_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    *stream_out = (void*)s;
    return SPX_OK;
}

int spx_stream_destroy(int device, void* stream) {
    if (!stream) return SPX_OK;
    SPX_CUDA(cudaSetDevice(device));
    SPX_CUDA(cudaStreamDestroy((cudaStream_t)stream));
    return SPX_OK;
}

int spx_stream_wait_stream(int device, void* waiter, void* signaller) {
    SPX_CUDA(cudaSetDevice(device));
    cudaEvent_t e = nullptr;
    SPX_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    cudaError_t r = cudaEventRecord(e, (cudaStream_t)signaller);
    if (r == cudaSuccess) r = cudaStreamWaitEvent((cudaStream_t)waiter, e, 0);
    cudaEventDestroy(e);      // released once the recorded work has completed
    if (r != cudaSuccess) return spx_set_error(SPX_E_CUDA, "spx_stream_wait_stream: %s", cudaGetErrorString(r));
    return SPX_OK;
}

int spx_stream_sync(int device, void* stream) {
    SPX_CUDA(cudaSetDevice(device));
    SPX_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return SPX_OK;
}

int spx_copy_ceiling(int device, const void* host_in, size_t h2d_bytes, void* host_out, size_t d2h_bytes, size_t piece_bytes,
                     int iters, double* seconds_out) {
    if (!seconds_out || iters < 1) return spx_set_error(SPX_E_INVALID, "bad argument");
    if ((h2d_bytes && !host_in) || (d2h_bytes && !host_out)) return spx_set_error(SPX_E_INVALID, "NULL host buffer");
    if (piece_bytes < 4096) piece_bytes = 16u << 20;
    SPX_CUDA(cudaSetDevice(device));
    void *d_in = nullptr, *d_out = nullptr;
    cudaStream_t s_in = nullptr, s_out = nullptr;
    int rc = SPX_OK;
    cudaError_t e = cudaSuccess;
    if (h2d_bytes) e = cudaMalloc(&d_in, h2d_bytes);
    if (e == cudaSuccess && d2h_bytes) e = cudaMalloc(&d_out, d2h_bytes);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking);
    if (e == cudaSuccess && d2h_bytes) e = cudaMemset(d_out, 0, d2h_bytes);
    if (e == cudaSuccess) {
        std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
        for (int it = -1; it < iters && e == cudaSuccess; ++it) {   // iteration -1 is the warm-up
            if (it == 0) {
                cudaStreamSynchronize(s_in);
                cudaStreamSynchronize(s_out);
            }
            if (it == 0) t0 = std::chrono::steady_clock::now();
            size_t a = 0, b = 0;
            while ((a < h2d_bytes || b < d2h_bytes) && e == cudaSuccess) {   // interleave the enqueues like the pipeline does
                if (a < h2d_bytes) {
                    const size_t n = h2d_bytes - a < piece_bytes ? h2d_bytes - a : piece_bytes;
                    e = cudaMemcpyAsync((char*)d_in + a, (const char*)host_in + a, n, cudaMemcpyHostToDevice, s_in);
                    a += n;
                }
                if (b < d2h_bytes && e == cudaSuccess) {
                    const size_t n = d2h_bytes - b < piece_bytes ? d2h_bytes - b : piece_bytes;
                    e = cudaMemcpyAsync((char*)host_out + b, (const char*)d_out + b, n, cudaMemcpyDeviceToHost, s_out);
                    b += n;
                }
            }
            if (it == iters - 1 && e == cudaSuccess) {
                e = cudaStreamSynchronize(s_in);
                if (e == cudaSuccess) e = cudaStreamSynchronize(s_out);
                *seconds_out = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / iters;
            }
        }
    }
    if (e != cudaSuccess) rc = spx_set_error(SPX_E_CUDA, "copy ceiling probe: %s", cudaGetErrorString(e));
    if (s_in) cudaStreamDestroy(s_in);
    if (s_out) cudaStreamDestroy(s_out);
    if (d_in) cudaFree(d_in);
    if (d_out) cudaFree(d_out);
    return rc;
}

// ------------------------------------------------------------------ collective step over peer memory (CUDA IPC + NVLink)
int spx_peer_reduce(int device, const double* welch_local, const float* maxhold_local, double* welch_owner,
                    float* maxhold_owner, int64_t n, void* stream) {
    if (n < 0) return spx_set_error(SPX_E_INVALID, "n < 0");
    if ((welch_local && !welch_owner) || (maxhold_local && !maxhold_owner)) return spx_set_error(SPX_E_INVALID, "owner buffer is NULL");
    SPX_CUDA(cudaSetDevice(device));
    NvtxRange r("spx_peer_reduce");
    return peer_reduce_launch(welch_local, maxhold_local, welch_owner, maxhold_owner, n, (cudaStream_t)stream);
}

int spx_peer_push_rows(int device, void* rows_owner, const void* rows_local, size_t bytes, void* stream) {
    if (bytes && (!rows_owner || !rows_local)) return spx_set_error(SPX_E_INVALID, "NULL rows");
    NvtxRange r("spx_peer_push_rows");
    return spx_memcpy_d2d_async(device, rows_owner, rows_local, bytes, stream);
}

// ------------------------------------------------------------------ collective step over NCCL
int spx_nccl_unique_id(void* id_out_128) {
    if (!id_out_128) return spx_set_error(SPX_E_INVALID, "id_out is NULL");
    SPX_TRY(nccl_load());
    ncclUniqueId id;
    SPX_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(id_out_128, &id, sizeof(id));
    return SPX_OK;
}

int spx_nccl_init(spx_comm** out, int device, int rank, int nranks, const void* unique_id_128) {
    if (!out || !unique_id_128 || rank < 0 || nranks < 1 || rank >= nranks) return spx_set_error(SPX_E_INVALID, "bad argument");
    *out = nullptr;
    SPX_TRY(nccl_load());
    SPX_CUDA(cudaSetDevice(device));
    spx_comm* c = new (std::nothrow) spx_comm();
    if (!c) return spx_set_error(SPX_E_NOMEM, "out of host memory");
    c->rank = rank;
    c->nranks = nranks;
    c->device = device;
    ncclUniqueId id;
    memcpy(&id, unique_id_128, sizeof(id));
    int r = g_nccl.CommInitRank(&c->comm, nranks, id, rank);
    if (r != ncclSuccess) {
        delete c;
        return spx_set_error(SPX_E_CUDA, "ncclCommInitRank failed: %s", g_nccl.GetErrorString(r));
    }
    if (cudaMalloc((void**)&c->d_count, sizeof(long long)) != cudaSuccess) {
        g_nccl.CommDestroy(c->comm);
        delete c;
        return spx_set_error(SPX_E_NOMEM, "cudaMalloc failed");
    }
    *out = c;
    return SPX_OK;
}

int spx_nccl_destroy(spx_comm* c) {
    if (!c) return SPX_OK;
    cudaSetDevice(c->device);
    if (c->d_count) cudaFree(c->d_count);
    if (c->comm) g_nccl.CommDestroy(c->comm);
    delete c;
    return SPX_OK;
}

int spx_allreduce_welch(spx_comm* c, double* welch_acc, float* maxhold, int64_t n, int64_t* n_frames_inout, void* stream) {
    if (!c) return spx_set_error(SPX_E_INVALID, "comm is NULL");
    SPX_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    NvtxRange r("spx_allreduce_welch");
    if (n_frames_inout) {
        const long long v = *n_frames_inout;
        SPX_CUDA(cudaMemcpyAsync(c->d_count, &v, sizeof(v), cudaMemcpyHostToDevice, st));
    }
    SPX_NCCL(g_nccl.GroupStart());
    if (welch_acc && n) SPX_NCCL(g_nccl.AllReduce(welch_acc, welch_acc, (size_t)n, ncclFloat64, ncclSum, c->comm, st));
    if (maxhold && n) SPX_NCCL(g_nccl.AllReduce(maxhold, maxhold, (size_t)n, ncclFloat32, ncclMax, c->comm, st));
    if (n_frames_inout) SPX_NCCL(g_nccl.AllReduce(c->d_count, c->d_count, 1, ncclInt64, ncclSum, c->comm, st));
    SPX_NCCL(g_nccl.GroupEnd());
    if (n_frames_inout) {
        long long v = 0;
        SPX_CUDA(cudaMemcpyAsync(&v, c->d_count, sizeof(v), cudaMemcpyDeviceToHost, st));
        SPX_CUDA(cudaStreamSynchronize(st));
        *n_frames_inout = v;
    }
    return SPX_OK;
}

int spx_gather_rows(spx_comm* c, const void* rows_local, int64_t local_bytes, void* rows_all, const int64_t* bytes_per_rank,
                    int dst, void* stream) {
    if (!c || !bytes_per_rank || dst < 0 || dst >= c->nranks) return spx_set_error(SPX_E_INVALID, "bad argument");
    if (bytes_per_rank[c->rank] != local_bytes) return spx_set_error(SPX_E_INVALID, "bytes_per_rank[rank] != local_bytes");
    SPX_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    NvtxRange r("spx_gather_rows");
    if (c->rank == dst) {
        if (!rows_all) return spx_set_error(SPX_E_INVALID, "rows_all is NULL on the destination rank");
        size_t off = 0;
        SPX_NCCL(g_nccl.GroupStart());
        for (int q = 0; q < c->nranks; ++q) {
            const size_t nb = (size_t)bytes_per_rank[q];
            if (q != dst && nb) SPX_NCCL(g_nccl.Recv((char*)rows_all + off, nb, ncclUint8, q, c->comm, st));
            off += nb;
        }
        SPX_NCCL(g_nccl.GroupEnd());
        off = 0;
        for (int q = 0; q < dst; ++q) off += (size_t)bytes_per_rank[q];
        if (local_bytes && (char*)rows_all + off != (const char*)rows_local)
            SPX_CUDA(cudaMemcpyAsync((char*)rows_all + off, rows_local, (size_t)local_bytes, cudaMemcpyDeviceToDevice, st));
    } else if (local_bytes) {
        SPX_NCCL(g_nccl.Send(rows_local, (size_t)local_bytes, ncclUint8, dst, c->comm, st));
    }
    return SPX_OK;
}

}  // extern "C"
