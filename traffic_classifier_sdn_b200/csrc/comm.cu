// comm.cu -- the one collective of the path: an all-gather of the per-shard label vectors.
//
// Rows are independent and the models are replicated, so `predict` itself never communicates: rank r classifies its
// contiguous block of ceil(n / world) rows.  Only when every rank wants the FULL label vector is there an exchange, and
// it is a single ncclAllGather of int32 class indices.  NCCL is bound at run time (dlopen of libnccl.so.2): the library
// has no link-time dependency on it, a process that never gathers never loads it, and inside a PyTorch process the
// already loaded NCCL is the one that gets used.
#include <dlfcn.h>
#include <nccl.h>   // types only; every entry point is looked up with dlsym

#include <algorithm>
#include <mutex>
#include <new>

#include "common.h"

namespace tcsdn {

struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi g_nccl;
static std::mutex g_nccl_mu;

static int nccl_load() {
    std::lock_guard<std::mutex> lock(g_nccl_mu);
    if (g_nccl.lib) return TCSDN_OK;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { set_error("cannot load libnccl.so.2: %s", dlerror()); return TCSDN_ECUDA; }
    NcclApi a;
    a.lib = h;
    a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
    a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
    a.AllGather = reinterpret_cast<decltype(a.AllGather)>(dlsym(h, "ncclAllGather"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    if (!a.GetUniqueId || !a.CommInitRank || !a.AllGather || !a.CommDestroy || !a.GetErrorString) {
        set_error("libnccl.so.2 lacks an expected entry point");
        return TCSDN_ECUDA;
    }
    g_nccl = a;
    return TCSDN_OK;
}

#define TCSDN_NCCL(expr)                                                                           \
    do {                                                                                           \
        ncclResult_t _r = (expr);                                                                  \
        if (_r != ncclSuccess) {                                                                   \
            tcsdn::set_error("%s failed: %s", #expr, tcsdn::g_nccl.GetErrorString(_r));            \
            return TCSDN_ECUDA;                                                                    \
        }                                                                                          \
    } while (0)

}  // namespace tcsdn

struct tcsdn_comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    int32_t *d_pad = nullptr;   // staging for a short last shard (n_local < n_block)
    int64_t pad_cap = 0;
    uint8_t *d_bytes = nullptr; // byte-wide wire format: [n_block] packed local labels, then [world * n_block] gathered
    int64_t bytes_cap = 0;      // in units of n_block
};

namespace tcsdn {

// int32 labels -> one byte each (0xFF = the -1 padding of a short shard), 16 labels per thread per step
__global__ void labels_pack_u8(const int32_t *__restrict__ src, int64_t n_src, uint8_t *__restrict__ dst, int64_t n_dst) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i * 4 < n_dst; i += stride) {
        uint32_t w = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int64_t e = i * 4 + k;
            const uint32_t v = e < n_src ? (uint32_t)src[e] & 0xFFu : 0xFFu;
            w |= v << (8 * k);
        }
        if (i * 4 + 4 <= n_dst) reinterpret_cast<uint32_t *>(dst)[i] = w;
        else for (int k = 0; i * 4 + k < n_dst; ++k) dst[i * 4 + k] = (uint8_t)(w >> (8 * k));
    }
}

__global__ void labels_unpack_u8(const uint8_t *__restrict__ src, int32_t *__restrict__ dst, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i * 4 < n; i += stride) {
        if (i * 4 + 4 <= n) {
            const uint32_t w = reinterpret_cast<const uint32_t *>(src)[i];
            int4 o;
            o.x = (w & 0xFFu) == 0xFFu ? -1 : (int)(w & 0xFFu);
            o.y = ((w >> 8) & 0xFFu) == 0xFFu ? -1 : (int)((w >> 8) & 0xFFu);
            o.z = ((w >> 16) & 0xFFu) == 0xFFu ? -1 : (int)((w >> 16) & 0xFFu);
            o.w = (w >> 24) == 0xFFu ? -1 : (int)(w >> 24);
            reinterpret_cast<int4 *>(dst)[i] = o;
        } else {
            for (int64_t e = i * 4; e < n; ++e) dst[e] = src[e] == 0xFF ? -1 : (int)src[e];
        }
    }
}

}  // namespace tcsdn

using namespace tcsdn;

extern "C" {

int tcsdn_comm_unique_id(void *id_out) {
    if (!id_out) { set_error("comm_unique_id: NULL output"); return TCSDN_EINVAL; }
    TCSDN_TRY(nccl_load());
    ncclUniqueId id;
    TCSDN_NCCL(g_nccl.GetUniqueId(&id));
    static_assert(sizeof(ncclUniqueId) == TCSDN_COMM_ID_BYTES, "ncclUniqueId size");
    memcpy(id_out, &id, sizeof(id));
    return TCSDN_OK;
}

int tcsdn_comm_init(int32_t rank, int32_t world, const void *unique_id, tcsdn_comm_t **out) {
    if (!out || !unique_id || world < 1 || rank < 0 || rank >= world) {
        set_error("comm_init: bad arguments (rank %d of %d)", rank, world);
        return TCSDN_EINVAL;
    }
    *out = nullptr;
    TCSDN_TRY(nccl_load());
    ncclUniqueId id;
    memcpy(&id, unique_id, sizeof(id));
    tcsdn_comm *c = new (std::nothrow) tcsdn_comm();
    if (!c) { set_error("comm_init: out of memory"); return TCSDN_ENOMEM; }
    c->rank = rank; c->world = world;
    ncclResult_t r = g_nccl.CommInitRank(&c->comm, world, id, rank);
    if (r != ncclSuccess) {
        set_error("ncclCommInitRank failed: %s", g_nccl.GetErrorString(r));
        delete c;
        return TCSDN_ECUDA;
    }
    *out = c;
    return TCSDN_OK;
}

int tcsdn_allgather_labels(tcsdn_comm_t *c, const int32_t *local, int64_t n_local, int64_t n_block, int32_t *all,
                           void *cuda_stream) {
    if (!c || !all || n_block < 0 || n_local < 0 || n_local > n_block || (n_local > 0 && !local)) {
        set_error("allgather_labels: bad arguments");
        return TCSDN_EINVAL;
    }
    if (n_block == 0) return TCSDN_OK;
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    const int32_t *send = local;
    if (n_local < n_block) {   // short (or empty) last shard: pad with -1 so that every rank contributes n_block entries
        if (c->pad_cap < n_block) {
            if (c->d_pad) cudaFree(c->d_pad);
            c->d_pad = nullptr; c->pad_cap = 0;
            TCSDN_CUDA(cudaMalloc(&c->d_pad, (size_t)n_block * sizeof(int32_t)));
            c->pad_cap = n_block;
        }
        TCSDN_CUDA(cudaMemsetAsync(c->d_pad, 0xFF, (size_t)n_block * sizeof(int32_t), st));
        if (n_local > 0)
            TCSDN_CUDA(cudaMemcpyAsync(c->d_pad, local, (size_t)n_local * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
        send = c->d_pad;
    }
    TCSDN_NCCL(g_nccl.AllGather(send, all, (size_t)n_block, ncclInt32, c->comm, st));
    return TCSDN_OK;
}

// Same result, a quarter of the bytes on the wire: class indices below 255 travel as one byte each; every rank
// packs its block, one ncclAllGather of bytes, every rank widens the gathered vector back to int32.  Everything is
// enqueued on `cuda_stream` with no host synchronisation, so predict + gather can be captured into one CUDA graph
// (staging buffers are sized on the first call with a given n_block: make that call outside the capture).
int tcsdn_allgather_labels_u8(tcsdn_comm_t *c, const int32_t *local, int64_t n_local, int64_t n_block, int32_t *all,
                              int32_t n_classes, void *cuda_stream) {
    if (!c || !all || n_block < 0 || n_local < 0 || n_local > n_block || (n_local > 0 && !local)) {
        set_error("allgather_labels_u8: bad arguments");
        return TCSDN_EINVAL;
    }
    if (n_classes < 1 || n_classes > 255) return tcsdn_allgather_labels(c, local, n_local, n_block, all, cuda_stream);
    if (n_block == 0) return TCSDN_OK;
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    const int64_t nb4 = (n_block + 3) & ~(int64_t)3;   // blocks start on 4-byte boundaries
    if (c->bytes_cap < nb4) {
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(st, &cap);
        if (cap != cudaStreamCaptureStatusNone) {
            set_error("allgather_labels_u8: the first call with n_block=%lld allocates staging memory; make it outside the graph capture", (long long)n_block);
            return TCSDN_EINVAL;
        }
        if (c->d_bytes) cudaFree(c->d_bytes);
        c->d_bytes = nullptr; c->bytes_cap = 0;
        TCSDN_CUDA(cudaMalloc(&c->d_bytes, (size_t)nb4 * (size_t)(c->world + 1)));
        c->bytes_cap = nb4;
    }
    uint8_t *mine = c->d_bytes, *gathered = c->d_bytes + c->bytes_cap;
    const int threads = 256;
    int64_t blocks = (nb4 / 4 + threads - 1) / threads;
    if (blocks > 148 * 8) blocks = 148 * 8;
    labels_pack_u8<<<(unsigned)blocks, threads, 0, st>>>(local, n_local, mine, nb4);
    TCSDN_CUDA(cudaGetLastError());
    TCSDN_NCCL(g_nccl.AllGather(mine, gathered, (size_t)nb4, ncclUint8, c->comm, st));
    if (nb4 == n_block) {
        const int64_t total = n_block * c->world;
        blocks = (total / 4 + threads - 1) / threads;
        if (blocks > 148 * 8) blocks = 148 * 8;
        labels_unpack_u8<<<(unsigned)blocks, threads, 0, st>>>(gathered, all, total);
    } else {
        for (int r = 0; r < c->world; ++r)   // odd block length: per-rank segments (destination blocks are n_block apart)
            labels_unpack_u8<<<(unsigned)std::max<int64_t>(1, std::min<int64_t>(blocks, 148 * 8)), threads, 0, st>>>(
                gathered + (size_t)r * nb4, all + (size_t)r * n_block, n_block);
    }
    TCSDN_CUDA(cudaGetLastError());
    return TCSDN_OK;
}

void tcsdn_comm_destroy(tcsdn_comm_t *c) {
    if (!c) return;
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    if (c->d_pad) cudaFree(c->d_pad);
    if (c->d_bytes) cudaFree(c->d_bytes);
    delete c;
}

}  // extern "C"
