// Row-band exchange transports (see comm.cuh): NCCL over NVLink, and an in-process loopback for single-GPU tests.
#include "comm.cuh"
#include "tapgemm.cuh"  // set_error
#include <chrono>
#include <condition_variable>
#include <dlfcn.h>
#include <mutex>
#include <string.h>
#include <vector>

namespace wfd {

#define CUDA_CK(expr)                                                           \
    do {                                                                        \
        cudaError_t e__ = (expr);                                               \
        if (e__ != cudaSuccess) {                                               \
            set_error("%s: %s", #expr, cudaGetErrorString(e__));                \
            return -8;                                                          \
        }                                                                       \
    } while (0)

// ------------------------------------------------------------------------------------------------ NCCL (dlopen)
// Only the handful of entry points used here are declared; the values follow nccl.h (2.27 / 2.28: ncclUint8 = 1,
// ncclInt64 = 4, ncclFloat32 = 7, ncclSum = 0, a 128-byte unique id passed by value).
namespace {
struct NcclId {
    char internal[128];
};
typedef void* NcclComm;
enum { kNcclUint8 = 1, kNcclInt64 = 4, kNcclFloat32 = 7, kNcclSum = 0 };
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(NcclId*) = nullptr;
    int (*CommInitRank)(NcclComm*, int, NcclId, int) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Send)(const void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
    int (*ReduceScatter)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
};
NcclApi g_nccl;
std::mutex g_nccl_mu;

int load_nccl(const char* path) {
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (g_nccl.lib) return 0;
    const char* cand[3] = {path && path[0] ? path : nullptr, "libnccl.so.2", "libnccl.so"};
    void* lib = nullptr;
    for (const char* c : cand) {
        if (!c) continue;
        lib = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
        if (lib) break;
    }
    if (!lib) {
        set_error("row bands: cannot load libnccl (%s)", dlerror());
        return -20;
    }
#define SYM(field, name)                                                 \
    do {                                                                 \
        *reinterpret_cast<void**>(&g_nccl.field) = dlsym(lib, name);     \
        if (!g_nccl.field) {                                             \
            set_error("row bands: libnccl has no symbol %s", name);      \
            dlclose(lib);                                                \
            return -20;                                                  \
        }                                                                \
    } while (0)
    SYM(GetUniqueId, "ncclGetUniqueId");
    SYM(CommInitRank, "ncclCommInitRank");
    SYM(CommDestroy, "ncclCommDestroy");
    SYM(GetErrorString, "ncclGetErrorString");
    SYM(GroupStart, "ncclGroupStart");
    SYM(GroupEnd, "ncclGroupEnd");
    SYM(Send, "ncclSend");
    SYM(Recv, "ncclRecv");
    SYM(AllReduce, "ncclAllReduce");
    SYM(AllGather, "ncclAllGather");
    SYM(ReduceScatter, "ncclReduceScatter");
#undef SYM
    g_nccl.lib = lib;
    return 0;
}

#define NCCL_CK(expr)                                                             \
    do {                                                                          \
        int r__ = (expr);                                                         \
        if (r__ != 0) {                                                           \
            set_error("%s: %s", #expr, g_nccl.GetErrorString(r__));               \
            return -21;                                                           \
        }                                                                         \
    } while (0)

struct NcclComm_ : Comm {
    NcclComm comm = nullptr;
    ~NcclComm_() override {
        if (comm) g_nccl.CommDestroy(comm);
    }
    int halo(void* top, const void* first, const void* last, void* bottom, size_t bytes, cudaStream_t s) override {
        if (rank == 0) CUDA_CK(cudaMemsetAsync(top, 0, bytes, s));
        if (rank == nranks - 1) CUDA_CK(cudaMemsetAsync(bottom, 0, bytes, s));
        if (nranks == 1) return 0;
        NCCL_CK(g_nccl.GroupStart());
        if (rank > 0) {
            NCCL_CK(g_nccl.Send(first, bytes, kNcclUint8, rank - 1, comm, s));
            NCCL_CK(g_nccl.Recv(top, bytes, kNcclUint8, rank - 1, comm, s));
        }
        if (rank < nranks - 1) {
            NCCL_CK(g_nccl.Send(last, bytes, kNcclUint8, rank + 1, comm, s));
            NCCL_CK(g_nccl.Recv(bottom, bytes, kNcclUint8, rank + 1, comm, s));
        }
        NCCL_CK(g_nccl.GroupEnd());
        return 0;
    }
    int allreduce_i64(long long* buf, size_t n, cudaStream_t s) override {
        if (nranks == 1) return 0;
        NCCL_CK(g_nccl.AllReduce(buf, buf, n, kNcclInt64, kNcclSum, comm, s));
        return 0;
    }
    int allreduce_f32(float* buf, size_t n, cudaStream_t s) override {
        if (nranks == 1) return 0;
        NCCL_CK(g_nccl.AllReduce(buf, buf, n, kNcclFloat32, kNcclSum, comm, s));
        return 0;
    }
    int allgather(void* buf, size_t bytes_per_rank, cudaStream_t s) override {
        if (nranks == 1) return 0;
        const char* mine = static_cast<const char*>(buf) + static_cast<size_t>(rank) * bytes_per_rank;
        NCCL_CK(g_nccl.AllGather(mine, buf, bytes_per_rank, kNcclUint8, comm, s));
        return 0;
    }
    int reduce_scatter_f32(float* buf, size_t n_per_rank, cudaStream_t s) override {
        if (nranks == 1) return 0;
        NCCL_CK(g_nccl.ReduceScatter(buf, buf + static_cast<size_t>(rank) * n_per_rank, n_per_rank, kNcclFloat32, kNcclSum,
                                     comm, s));
        return 0;
    }
};
}  // namespace

int nccl_unique_id(const char* libnccl_path, void* id128) {
    if (int rc = load_nccl(libnccl_path)) return rc;
    NcclId id;
    NCCL_CK(g_nccl.GetUniqueId(&id));
    memcpy(id128, id.internal, 128);
    return 0;
}

int nccl_comm_create(const char* libnccl_path, const void* id128, int rank, int nranks, Comm** out) {
    if (int rc = load_nccl(libnccl_path)) return rc;
    if (rank < 0 || rank >= nranks) {
        set_error("row bands: rank %d outside [0, %d)", rank, nranks);
        return -3;
    }
    NcclId id;
    memcpy(id.internal, id128, 128);
    NcclComm_* c = new NcclComm_();
    c->rank = rank;
    c->nranks = nranks;
    int r = g_nccl.CommInitRank(&c->comm, nranks, id, rank);
    if (r != 0) {
        set_error("ncclCommInitRank: %s", g_nccl.GetErrorString(r));
        c->comm = nullptr;
        delete c;
        return -21;
    }
    *out = c;
    return 0;
}

// ------------------------------------------------------------------------------------------------ loopback
// Every band lives in this process on the same GPU and is driven by its own host thread. Each step is
//   sync my stream -> publish my pointers -> barrier -> read the peers' data -> barrier (-> write my result)
// so it is slow and fully serialising, which is fine for what it is for: running the band plan on one GPU in the tests.
struct LoopGroup {
    int nranks = 0;
    std::mutex mu;
    std::condition_variable cv;
    int waiting = 0;
    long long generation = 0;
    std::vector<const void*> a, b;
    bool broken = false;
    // false when a peer never arrives (it failed before its matching call): every later barrier fails at once
    bool barrier() {
        std::unique_lock<std::mutex> lk(mu);
        if (broken) return false;
        const long long gen = generation;
        if (++waiting == nranks) {
            waiting = 0;
            ++generation;
            cv.notify_all();
            return true;
        }
        if (!cv.wait_for(lk, std::chrono::seconds(120), [&] { return generation != gen || broken; }) || broken) {
            broken = true;
            cv.notify_all();
            return false;
        }
        return true;
    }
};

LoopGroup* loop_group_create(int nranks) {
    if (nranks < 1) return nullptr;
    LoopGroup* g = new LoopGroup();
    g->nranks = nranks;
    g->a.assign(nranks, nullptr);
    g->b.assign(nranks, nullptr);
    return g;
}
void loop_group_destroy(LoopGroup* g) { delete g; }

#define LOOP_BARRIER()                                                                   \
    do {                                                                                 \
        if (!g->barrier()) {                                                             \
            set_error("row bands (loopback): a peer band did not reach the exchange");   \
            return -22;                                                                  \
        }                                                                                \
    } while (0)

namespace {
struct LoopComm : Comm {
    LoopGroup* g = nullptr;
    int halo(void* top, const void* first, const void* last, void* bottom, size_t bytes, cudaStream_t s) override {
        CUDA_CK(cudaStreamSynchronize(s));
        g->a[rank] = first;
        g->b[rank] = last;
        LOOP_BARRIER();
        if (rank > 0) CUDA_CK(cudaMemcpyAsync(top, g->b[rank - 1], bytes, cudaMemcpyDeviceToDevice, s));
        else CUDA_CK(cudaMemsetAsync(top, 0, bytes, s));
        if (rank < nranks - 1) CUDA_CK(cudaMemcpyAsync(bottom, g->a[rank + 1], bytes, cudaMemcpyDeviceToDevice, s));
        else CUDA_CK(cudaMemsetAsync(bottom, 0, bytes, s));
        CUDA_CK(cudaStreamSynchronize(s));
        LOOP_BARRIER();
        return 0;
    }
    template <typename T>
    int allreduce(T* buf, size_t n, cudaStream_t s) {
        CUDA_CK(cudaStreamSynchronize(s));
        g->a[rank] = buf;
        LOOP_BARRIER();
        std::vector<T> acc(n, T(0)), tmp(n);
        for (int r = 0; r < nranks; ++r) {
            CUDA_CK(cudaMemcpy(tmp.data(), g->a[r], n * sizeof(T), cudaMemcpyDeviceToHost));
            for (size_t i = 0; i < n; ++i) acc[i] += tmp[i];
        }
        LOOP_BARRIER();
        CUDA_CK(cudaMemcpy(buf, acc.data(), n * sizeof(T), cudaMemcpyHostToDevice));
        return 0;
    }
    int allreduce_i64(long long* buf, size_t n, cudaStream_t s) override { return allreduce<long long>(buf, n, s); }
    int allreduce_f32(float* buf, size_t n, cudaStream_t s) override { return allreduce<float>(buf, n, s); }
    int allgather(void* buf, size_t bytes_per_rank, cudaStream_t s) override {
        CUDA_CK(cudaStreamSynchronize(s));
        g->a[rank] = buf;
        LOOP_BARRIER();
        for (int r = 0; r < nranks; ++r) {
            if (r == rank) continue;
            const size_t off = static_cast<size_t>(r) * bytes_per_rank;
            CUDA_CK(cudaMemcpyAsync(static_cast<char*>(buf) + off, static_cast<const char*>(g->a[r]) + off, bytes_per_rank,
                                    cudaMemcpyDeviceToDevice, s));
        }
        CUDA_CK(cudaStreamSynchronize(s));
        LOOP_BARRIER();
        return 0;
    }
    int reduce_scatter_f32(float* buf, size_t n_per_rank, cudaStream_t s) override {
        CUDA_CK(cudaStreamSynchronize(s));
        g->a[rank] = buf;
        LOOP_BARRIER();
        std::vector<float> acc(n_per_rank, 0.f), tmp(n_per_rank);
        const size_t off = static_cast<size_t>(rank) * n_per_rank;
        for (int r = 0; r < nranks; ++r) {
            CUDA_CK(cudaMemcpy(tmp.data(), static_cast<const float*>(g->a[r]) + off, n_per_rank * sizeof(float),
                               cudaMemcpyDeviceToHost));
            for (size_t i = 0; i < n_per_rank; ++i) acc[i] += tmp[i];
        }
        LOOP_BARRIER();
        CUDA_CK(cudaMemcpy(buf + off, acc.data(), n_per_rank * sizeof(float), cudaMemcpyHostToDevice));
        return 0;
    }
};
}  // namespace

int loop_comm_create(LoopGroup* g, int rank, Comm** out) {
    if (!g || rank < 0 || rank >= g->nranks) {
        set_error("row bands: bad loopback group / rank");
        return -3;
    }
    LoopComm* c = new LoopComm();
    c->g = g;
    c->rank = rank;
    c->nranks = g->nranks;
    *out = c;
    return 0;
}

}  // namespace wfd
