// Exchange steps of a row-band plan: one 256x256 map split into bands of rows, one band per GPU.
// The reference has no multi-GPU path for one map (it only widens `width/height`, pipeline_wfd.py:283-290); the
// exchanges below are exactly what its single-device ops need once the plane is cut along y:
//   * halo       - a 3x3 conv (resnet.py:416,432, wf_vae.py:218,232) and the vertical WFC pairs (losses.py:60-66) read
//                  one row above and below the band
//   * allreduce  - GroupNorm statistics are per (sample, group) over the whole plane (resnet.py:407,423); the loss is a
//                  sum over all pairs (losses.py:69-75)
//   * allgather / reduce-scatter - the mid-block attention (attention.py:329-380) attends over every pixel: keys and
//                  values are gathered, the key/value gradients are summed over the query bands
// Two transports: NCCL (one process per GPU, NVLink) loaded from the libnccl the process already uses, and an
// in-process loopback (several bands on one GPU, one host thread per band) used by the single-GPU parity tests.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace wfd {

struct Comm {
    int rank = 0, nranks = 1;
    virtual ~Comm() {}
    // my first row -> bottom halo of rank-1, my last row -> top halo of rank+1; halos at the map border are zeroed
    virtual int halo(void* top, const void* first, const void* last, void* bottom, size_t bytes, cudaStream_t s) = 0;
    virtual int allreduce_i64(long long* buf, size_t n, cudaStream_t s) = 0;
    virtual int allreduce_f32(float* buf, size_t n, cudaStream_t s) = 0;
    // in place: rank r's contribution sits at buf + r * bytes_per_rank
    virtual int allgather(void* buf, size_t bytes_per_rank, cudaStream_t s) = 0;
    // in place: every rank holds nranks * n_per_rank partial values; the sum of slice r lands in rank r's slice r
    virtual int reduce_scatter_f32(float* buf, size_t n_per_rank, cudaStream_t s) = 0;
    // Peer mailboxes (NCCL transport only; see comm.cu): the small exchanges - halo rows, the 16-value GroupNorm sums, the
    // loss - then go over NVLink as direct stores into the neighbour's memory plus a flag, one short kernel each, instead
    // of ncclSend/Recv/AllReduce. peer_local_handle allocates this rank's mailbox and returns its CUDA IPC handle (64
    // bytes); peer_attach maps every rank's mailbox (handles of all ranks, nranks x 64 bytes, in rank order).
    virtual int peer_local_handle(size_t max_row_bytes, void* handle64) { (void)max_row_bytes; (void)handle64; return -30; }
    virtual int peer_attach(const void* handles) { (void)handles; return -30; }
};

int nccl_unique_id(const char* libnccl_path, void* id128);
int nccl_comm_create(const char* libnccl_path, const void* id128, int rank, int nranks, Comm** out);

struct LoopGroup;
LoopGroup* loop_group_create(int nranks);
void loop_group_destroy(LoopGroup* g);
int loop_comm_create(LoopGroup* g, int rank, Comm** out);

}  // namespace wfd
