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

// ---- peer mailboxes: small exchanges as stores into the neighbour's memory over NVLink
// Every rank owns one mailbox (cudaMalloc + CUDA IPC, mapped by all ranks). Collectives are numbered by a counter that lives
// IN the mailbox (seq), so a captured CUDA graph replays correctly: a kernel reads seq, uses slot seq & 1, publishes flag
// = seq + 1 in the receivers' mailboxes and bumps its own seq when it is done. All ranks issue the same collectives in
// the same order; a rank can run at most one collective ahead of a rank it exchanges with (it needs that rank's flag to
// finish), so two data slots per kind are enough.
constexpr int PEER_MAX_RANKS = 16;
constexpr int PEER_AR_MAX = 32;          // values per all-reduce (GroupNorm: 16 int64, loss: 1 float)
struct PeerHdr {
    unsigned int seq;                     // collectives this rank has completed
    unsigned int arrive;                  // blocks of the running halo kernel that have pushed their share
    unsigned int halo_flag[2];            // [0] set by rank - 1 (my top halo has landed), [1] by rank + 1
    unsigned int ar_flag[PEER_MAX_RANKS]; // set by each source rank
    unsigned int pad[44];                 // 256 bytes
};
struct PeerView {                         // device-side view of one mailbox
    PeerHdr* hdr;
    unsigned long long* ar;               // [2][PEER_MAX_RANKS][PEER_AR_MAX]
    unsigned char* halo;                  // [2][2][row_cap]
};
struct PeerTable {
    PeerView box[PEER_MAX_RANKS];
    size_t row_cap;
    int rank, nranks;
};

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ bool seq_reached(unsigned int flag, unsigned int want) {
    return static_cast<int>(flag - want) >= 0;  // wrap-safe
}

constexpr int PEER_HALO_BLOCKS = 16;
// my first row -> bottom halo of rank - 1, my last row -> top halo of rank + 1, through their mailboxes; then my own halos
// are copied out of my mailbox (or zeroed at the map border). bytes is a multiple of 16.
__global__ void __launch_bounds__(256)
peer_halo_k(PeerTable t, unsigned char* top, const unsigned char* first, const unsigned char* last, unsigned char* bottom,
            size_t bytes) {
    __shared__ unsigned int s_seq, s_last;
    PeerHdr* me = t.box[t.rank].hdr;
    if (threadIdx.x == 0) s_seq = *reinterpret_cast<volatile unsigned int*>(&me->seq);
    __syncthreads();
    const unsigned int seq = s_seq;
    const size_t slot = (seq & 1u) * 2u;
    const size_t nv = bytes / 16;
    const size_t tid = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x, nth = static_cast<size_t>(gridDim.x) * blockDim.x;
    if (t.rank > 0) {
        uint4* dst = reinterpret_cast<uint4*>(t.box[t.rank - 1].halo + (slot + 1) * t.row_cap);
        const uint4* src = reinterpret_cast<const uint4*>(first);
        for (size_t i = tid; i < nv; i += nth) dst[i] = src[i];
    }
    if (t.rank < t.nranks - 1) {
        uint4* dst = reinterpret_cast<uint4*>(t.box[t.rank + 1].halo + (slot + 0) * t.row_cap);
        const uint4* src = reinterpret_cast<const uint4*>(last);
        for (size_t i = tid; i < nv; i += nth) dst[i] = src[i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(&me->arrive, 1u);
        s_last = (prev == gridDim.x - 1) ? 1u : 0u;
        if (s_last) {
            // every block's rows are on their way: publish, reset the arrival counter, count this collective
            __threadfence_system();
            if (t.rank > 0) st_release_sys(&t.box[t.rank - 1].hdr->halo_flag[1], seq + 1);
            if (t.rank < t.nranks - 1) st_release_sys(&t.box[t.rank + 1].hdr->halo_flag[0], seq + 1);
            me->arrive = 0;
            *reinterpret_cast<volatile unsigned int*>(&me->seq) = seq + 1;
        }
        if (t.rank > 0)
            while (!seq_reached(ld_acquire_sys(&me->halo_flag[0]), seq + 1)) {}
        if (t.rank < t.nranks - 1)
            while (!seq_reached(ld_acquire_sys(&me->halo_flag[1]), seq + 1)) {}
    }
    __syncthreads();
    {
        uint4* dst = reinterpret_cast<uint4*>(top);
        const uint4* src = reinterpret_cast<const uint4*>(t.box[t.rank].halo + (slot + 0) * t.row_cap);
        if (t.rank > 0) for (size_t i = tid; i < nv; i += nth) dst[i] = __ldcv(src + i);
        else for (size_t i = tid; i < nv; i += nth) dst[i] = make_uint4(0, 0, 0, 0);
    }
    {
        uint4* dst = reinterpret_cast<uint4*>(bottom);
        const uint4* src = reinterpret_cast<const uint4*>(t.box[t.rank].halo + (slot + 1) * t.row_cap);
        if (t.rank < t.nranks - 1) for (size_t i = tid; i < nv; i += nth) dst[i] = __ldcv(src + i);
        else for (size_t i = tid; i < nv; i += nth) dst[i] = make_uint4(0, 0, 0, 0);
    }
}

// all-reduce of n <= PEER_AR_MAX 8-byte integers or 4-byte floats: every rank stores its values into every mailbox,
// flags, waits for all flags and sums the nranks contributions in rank order (the same order on every rank)
template <typename T>
__global__ void __launch_bounds__(512)
peer_allreduce_k(PeerTable t, T* buf, int n) {
    __shared__ unsigned int s_seq;
    PeerHdr* me = t.box[t.rank].hdr;
    if (threadIdx.x == 0) s_seq = *reinterpret_cast<volatile unsigned int*>(&me->seq);
    __syncthreads();
    const unsigned int seq = s_seq;
    const size_t slot = (seq & 1u) * PEER_MAX_RANKS * PEER_AR_MAX;
    for (int j = threadIdx.x; j < t.nranks * n; j += blockDim.x) {
        const int r = j / n, i = j - r * n;
        unsigned long long v = 0;
        if (sizeof(T) == 8) v = static_cast<unsigned long long>(reinterpret_cast<const long long*>(buf)[i]);
        else v = __float_as_uint(reinterpret_cast<const float*>(buf)[i]);
        t.box[r].ar[slot + t.rank * PEER_AR_MAX + i] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < t.nranks && threadIdx.x != t.rank) {
        st_release_sys(&t.box[threadIdx.x].hdr->ar_flag[t.rank], seq + 1);
        while (!seq_reached(ld_acquire_sys(&me->ar_flag[threadIdx.x]), seq + 1)) {}
    }
    __syncthreads();
    if (threadIdx.x < n) {
        const volatile unsigned long long* mine = t.box[t.rank].ar + slot;
        if (sizeof(T) == 8) {
            long long acc = 0;
            for (int r = 0; r < t.nranks; ++r) acc += static_cast<long long>(mine[r * PEER_AR_MAX + threadIdx.x]);
            reinterpret_cast<long long*>(buf)[threadIdx.x] = acc;
        } else {
            float acc = 0.f;
            for (int r = 0; r < t.nranks; ++r) acc += __uint_as_float(static_cast<unsigned int>(mine[r * PEER_AR_MAX + threadIdx.x]));
            reinterpret_cast<float*>(buf)[threadIdx.x] = acc;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) *reinterpret_cast<volatile unsigned int*>(&me->seq) = seq + 1;
}

struct NcclComm_ : Comm {
    NcclComm comm = nullptr;
    // peer mailboxes (optional)
    void* box_local = nullptr;
    size_t box_bytes = 0, row_cap = 0;
    void* box_mapped[PEER_MAX_RANKS] = {nullptr};
    PeerTable table;
    bool peer_on = false;
    ~NcclComm_() override {
        for (int r = 0; r < PEER_MAX_RANKS; ++r)
            if (box_mapped[r] && r != rank) cudaIpcCloseMemHandle(box_mapped[r]);
        if (box_local) cudaFree(box_local);
        if (comm) g_nccl.CommDestroy(comm);
    }
    static size_t ar_bytes() { return static_cast<size_t>(2) * PEER_MAX_RANKS * PEER_AR_MAX * 8; }
    int peer_local_handle(size_t max_row_bytes, void* handle64) override {
        if (nranks > PEER_MAX_RANKS || !handle64) {
            set_error("row bands: peer mailboxes support up to %d ranks", PEER_MAX_RANKS);
            return -3;
        }
        if (!box_local) {
            row_cap = (max_row_bytes + 255) & ~size_t(255);
            box_bytes = sizeof(PeerHdr) + ar_bytes() + 4 * row_cap;
            CUDA_CK(cudaMalloc(&box_local, box_bytes));
            CUDA_CK(cudaMemset(box_local, 0, box_bytes));
            CUDA_CK(cudaDeviceSynchronize());
        }
        cudaIpcMemHandle_t h;
        CUDA_CK(cudaIpcGetMemHandle(&h, box_local));
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
        memcpy(handle64, &h, 64);
        return 0;
    }
    int peer_attach(const void* handles) override {
        if (!box_local || !handles) {
            set_error("row bands: peer_attach before peer_local_handle");
            return -3;
        }
        memset(&table, 0, sizeof(table));
        for (int r = 0; r < nranks; ++r) {
            void* base = box_local;
            if (r != rank) {
                cudaIpcMemHandle_t h;
                memcpy(&h, static_cast<const char*>(handles) + 64 * r, 64);
                CUDA_CK(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
                box_mapped[r] = base;
            }
            unsigned char* b = static_cast<unsigned char*>(base);
            table.box[r].hdr = reinterpret_cast<PeerHdr*>(b);
            table.box[r].ar = reinterpret_cast<unsigned long long*>(b + sizeof(PeerHdr));
            table.box[r].halo = b + sizeof(PeerHdr) + ar_bytes();
        }
        table.row_cap = row_cap;
        table.rank = rank;
        table.nranks = nranks;
        peer_on = true;
        return 0;
    }
    int halo(void* top, const void* first, const void* last, void* bottom, size_t bytes, cudaStream_t s) override {
        if (peer_on && nranks > 1 && bytes <= row_cap && bytes % 16 == 0) {
            peer_halo_k<<<PEER_HALO_BLOCKS, 256, 0, s>>>(table, static_cast<unsigned char*>(top), static_cast<const unsigned char*>(first),
                                                        static_cast<const unsigned char*>(last), static_cast<unsigned char*>(bottom), bytes);
            CUDA_CK(cudaGetLastError());
            return 0;
        }
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
        if (peer_on && n <= PEER_AR_MAX) {
            peer_allreduce_k<long long><<<1, 512, 0, s>>>(table, buf, static_cast<int>(n));
            CUDA_CK(cudaGetLastError());
            return 0;
        }
        NCCL_CK(g_nccl.AllReduce(buf, buf, n, kNcclInt64, kNcclSum, comm, s));
        return 0;
    }
    int allreduce_f32(float* buf, size_t n, cudaStream_t s) override {
        if (nranks == 1) return 0;
        if (peer_on && n <= PEER_AR_MAX) {
            peer_allreduce_k<float><<<1, 512, 0, s>>>(table, buf, static_cast<int>(n));
            CUDA_CK(cudaGetLastError());
            return 0;
        }
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
