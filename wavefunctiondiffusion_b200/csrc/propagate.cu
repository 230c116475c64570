// Wave-seeded WFC constraint propagation on the device: the information-propagation step of the reference's
// prior-distribution solver (wfd/wfc/wfc_prior_dist.py:136-174 propagate_information, :176-199 double_check) as bit-set
// arc consistency over the whole map, plus the kernel that turns a wave into the solver's allowed-tile sets.
//
// Data: allowed  uint32 [B, H, W, Wd]   bit (t & 31) of word (t >> 5) = tile t still possible in the cell (Wd = ceil(T/32))
//       conn     uint32 [4, T, Wd]      conn[d][s] = tiles that may sit at direction d of tile s; d as the reference orders
//                                       them (wfc_prior_dist.py:25): 0 left (-1,0), 1 up (0,-1), 2 right (1,0), 3 down (0,1).
//                                       Handed over once (host memory) and re-encoded for the device as SUPPORTER lists:
//                                       for direction d and target tile t the source tiles s with conn[d][s][t] set. The
//                                       StarCraft rules are 0.2 % dense, so a list is ~7 16-bit ids (CSR); lists longer
//                                       than 255 (the null tile's, :207-212, which everything supports) stay bit rows
//       boundary uint8  [B, H, W]       non-zero = cell takes no part (is_grid_allowed false), may be NULL
// One sweep, for every cell n and direction d with source m = n - dir_d inside the map:
//       allowed[n] &= OR_{s in allowed[m]} conn[d][s]            (an empty source does not propagate: the reference
//                                                                  returns False there without touching its neighbours)
// Sweeps repeat until nothing changes. Every sweep reads the previous state and writes the next one (two buffers), so the
// result does not depend on the order in which the CTAs run: it is the state that numpy reaches sweep by sweep
// (oracle/wfc_propagate_oracle.py), bit for bit, contradictions included. While no cell runs empty the update is
// monotone, and the fixed point is the one the reference's raster-order depth-1 propagation reaches when it is repeated
// until nothing changes. A cell is re-evaluated only when it or one of its neighbours changed in the previous sweep.
// Integer / bit work: per (cell, direction) the source's words are staged in shared memory; every tile the cell still
// allows (one lane per bit of a state word) walks its supporter list until it finds a source tile that is still possible,
// and a warp ballot is the narrowed word. No lists of set bits, no atomics. The supporter lists (~200 KB for ice) live
// in L1 / L2; HBM sees the state words of the cells that are re-evaluated. No tensor cores.
#include "kernels.cuh"
#include <cuda_fp16.h>
#include <new>
#include <vector>
#include "../../include/wfd_b200.h"

namespace wfd {

#define WFD_CHECK_LAUNCH_P(name)                                        \
    do {                                                                \
        cudaError_t e__ = cudaGetLastError();                           \
        if (e__ != cudaSuccess) {                                       \
            set_error("%s launch: %s", name, cudaGetErrorString(e__));  \
            return -7;                                                  \
        }                                                               \
        ++g_launch_count;                                               \
        if (debug_sync_enabled()) {                                     \
            int rc__ = debug_sync_check(name, s);                       \
            if (rc__ < 0) return rc__;                                  \
        }                                                               \
    } while (0)

constexpr int PROP_THREADS = 128;
constexpr int PROP_MAX_WD = 256;          // T <= 8192
constexpr int PROP_MAX_SWEEPS = 4096;     // one change counter per sweep

// ------------------------------------------------------------------------------------------------ wave -> allowed sets
// One warp per cell. allowed[n, t] = (first_tile <= t < T) and (wave[n, t] >= thr or (keep_argmax and t == argmax)),
// the argmax over [first_tile, T) with numpy's rule (first maximum wins). dtype: 0 fp32, 1 fp16, 2 bf16.
template <int DT>
__device__ __forceinline__ float ld_wave(const void* p, long long i) {
    if (DT == 0) return __ldg(reinterpret_cast<const float*>(p) + i);
    const unsigned short h = __ldg(reinterpret_cast<const unsigned short*>(p) + i);
    if (DT == 1) return __half2float(__ushort_as_half(h));
    return __uint_as_float(static_cast<unsigned int>(h) << 16);
}
template <int DT>
__global__ void __launch_bounds__(256)
wave_to_allowed_k(const void* __restrict__ wave, uint32_t* __restrict__ allowed, long long rows, int ld, int T, int Wd,
                  int first_tile, float thr, int keep_argmax) {
    const int lane = threadIdx.x & 31;
    const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const long long base = row * ld;
    int best_i = -1;
    if (keep_argmax) {
        float best = -INFINITY;
        for (int t = first_tile + lane; t < T; t += 32) {
            const float v = ld_wave<DT>(wave, base + t);
            if (best_i < 0 || v > best) {  // strictly greater: the first maximum of this lane's stride wins
                best = v;
                best_i = t;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
            if (oi >= 0 && (best_i < 0 || ov > best || (ov == best && oi < best_i))) {
                best = ov;
                best_i = oi;
            }
        }
    }
    uint32_t mine = 0u;  // lane l keeps word (w0 + l) of the current group of 32 words: coalesced 128-byte stores
    for (int w = 0; w < Wd; ++w) {
        const int t = w * 32 + lane;
        bool on = false;
        if (t >= first_tile && t < T) on = ld_wave<DT>(wave, base + t) >= thr || t == best_i;
        const uint32_t word = __ballot_sync(0xffffffffu, on);
        if ((w & 31) == lane) mine = word;
        if ((w & 31) == 31 || w == Wd - 1) {
            const int wi = (w & ~31) + lane;
            if (wi <= w) allowed[row * Wd + wi] = mine;
        }
    }
}

// ------------------------------------------------------------------------------------------------ one sweep
struct PropRules {          // device pointers, built by wfd_propagator_create
    const int* ptr;         // [4 * T + 1] CSR row starts of the supporter lists, row index d * T + t
    const uint16_t* col;    // supporter tile ids, ascending; a list kept as bits is the pair {0xFFFF, index into dense}
    const uint32_t* dense;  // [n_dense][Wd]
};

template <int WPT>
__global__ void __launch_bounds__(PROP_THREADS)
wfc_propagate_sweep_k(const PropRules rules, const uint32_t* __restrict__ state_in, uint32_t* __restrict__ state_out,
                      const uint8_t* __restrict__ boundary, uint8_t* dirty_in, uint8_t* dirty_out, unsigned int* changed,
                      int sweep, int H, int W, int T, int Wd) {
    __shared__ uint32_t nb[PROP_MAX_WD];    // the source cell's words
    __shared__ uint32_t curw[PROP_MAX_WD];  // this cell's words, narrowed direction by direction
    if (sweep > 0 && changed[sweep - 1] == 0u) return;  // the previous sweep changed nothing: the fixed point is reached
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cell = blockIdx.x;
    const long long n = static_cast<long long>(blockIdx.y) * H * W + cell;
    if (boundary && boundary[n]) return;  // never rewritten: both buffers hold its initial state
    if (sweep > 0) {
        // Only cells that changed, or have a neighbour that changed, during the previous sweep. Everybody else holds the
        // same words in both buffers (its last evaluation wrote back what it had read).
        const uint8_t dflag = dirty_in[n];
        __syncthreads();
        if (!dflag) return;
        if (tid == 0) dirty_in[n] = 0;
    }
    const int y = cell / W, x = cell - y * W;
    uint32_t orig[WPT];
#pragma unroll
    for (int k = 0; k < WPT; ++k) {
        const int w = tid + k * PROP_THREADS;
        orig[k] = w < Wd ? __ldg(state_in + n * Wd + w) : 0u;
        if (w < Wd) curw[w] = orig[k];
    }
#pragma unroll 1
    for (int d = 0; d < 4; ++d) {
        const int dx = d == 0 ? -1 : (d == 2 ? 1 : 0), dy = d == 1 ? -1 : (d == 3 ? 1 : 0);
        const int my = y - dy, mx = x - dx;  // the source whose direction-d neighbour this cell is
        if (my < 0 || my >= H || mx < 0 || mx >= W) continue;
        const long long m = static_cast<long long>(blockIdx.y) * H * W + my * W + mx;
        if (boundary && boundary[m]) continue;
        bool some = false;
#pragma unroll
        for (int k = 0; k < WPT; ++k) {
            const int w = tid + k * PROP_THREADS;
            if (w < Wd) {
                const uint32_t v = __ldg(state_in + m * Wd + w);
                nb[w] = v;
                some = some || v != 0u;
            }
        }
        // the barrier also publishes nb and the previous direction's curw
        if (!__syncthreads_or(some)) continue;  // an empty source does not propagate (wfc_prior_dist.py:152-153)
        const int* rp = rules.ptr + static_cast<size_t>(d) * T;
        for (int w = warp; w < Wd; w += PROP_THREADS / 32) {
            const uint32_t cw = curw[w];
            if (cw == 0u) continue;  // warp-uniform
            const int t = w * 32 + lane;
            bool keep = false;
            if ((cw >> lane) & 1u) {
                const int beg = __ldg(rp + t), end = __ldg(rp + t + 1);
                for (int e = beg; e < end; ++e) {
                    const int sid = __ldg(rules.col + e);
                    if (sid == 0xFFFF) {  // a supporter list kept as bits
                        const uint32_t* row = rules.dense + static_cast<size_t>(__ldg(rules.col + e + 1)) * Wd;
                        for (int ww = 0; ww < Wd && !keep; ++ww) keep = (nb[ww] & __ldg(row + ww)) != 0u;
                        break;
                    }
                    if ((nb[sid >> 5] >> (sid & 31)) & 1u) {
                        keep = true;
                        break;
                    }
                }
            }
            const uint32_t word = __ballot_sync(0xffffffffu, keep);
            if (lane == 0) curw[w] = word;
        }
        __syncthreads();  // nb is rewritten for the next direction; curw is read by other warps
    }
    __syncthreads();
    bool diff = false;
#pragma unroll
    for (int k = 0; k < WPT; ++k) {
        const int w = tid + k * PROP_THREADS;
        if (w < Wd) {
            const uint32_t v = curw[w];
            state_out[n * Wd + w] = v;
            diff = diff || v != orig[k];
        }
    }
    if (__syncthreads_or(diff) && tid == 0) {
        atomicAdd(changed + sweep, 1u);
        dirty_out[n] = 1;  // evaluated again next sweep, which also brings the other buffer up to date
        if (x > 0) dirty_out[n - 1] = 1;
        if (x + 1 < W) dirty_out[n + 1] = 1;
        if (y > 0) dirty_out[n - W] = 1;
        if (y + 1 < H) dirty_out[n + W] = 1;
    }
}

// number of tiles each cell still allows (is_fixed / is_zero of the reference, wfc_prior_dist.py:86-90): one warp per cell
__global__ void __launch_bounds__(256)
allowed_count_k(const uint32_t* __restrict__ allowed, int* __restrict__ count, long long cells, int Wd) {
    const int lane = threadIdx.x & 31;
    const long long n = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (n >= cells) return;
    int c = 0;
    for (int w = lane; w < Wd; w += 32) c += __popc(__ldg(allowed + n * Wd + w));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) count[n] = c;
}

static int check_shape(const char* who, int B, int H, int W, int T) {
    if (B < 1 || H < 1 || W < 1 || T < 1 || (T + 31) / 32 > PROP_MAX_WD) {
        set_error("%s: bad shape B=%d H=%d W=%d T=%d (T <= %d)", who, B, H, W, T, PROP_MAX_WD * 32);
        return -3;
    }
    return 0;
}

}  // namespace wfd

using namespace wfd;

struct wfd_propagator {
    int T = 0, Wd = 0, n_dense = 0;
    long long entries = 0;
    int* ptr = nullptr;
    uint16_t* col = nullptr;
    uint32_t* dense = nullptr;
    void* scratch = nullptr;
    size_t scratch_bytes = 0;
};

extern "C" {

int wfd_wave_to_allowed(const void* wave, int dtype, long long rows, int ld, int T, int first_tile, float threshold,
                        int keep_argmax, uint32_t* allowed, void* stream) {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    ProfScope prof_scope("wave_to_allowed", s);
    if (!wave || !allowed) {
        set_error("wfd_wave_to_allowed: null argument");
        return -1;
    }
    if (rows < 0 || T < 1 || ld < T || first_tile < 0 || dtype < 0 || dtype > 2 || (T + 31) / 32 > PROP_MAX_WD) {
        set_error("wfd_wave_to_allowed: bad arguments rows=%lld ld=%d T=%d first_tile=%d dtype=%d", rows, ld, T, first_tile, dtype);
        return -3;
    }
    if (rows == 0) return 0;
    const int Wd = (T + 31) / 32;
    const unsigned int blocks = static_cast<unsigned int>((rows + 7) / 8);
    if (dtype == 0) wave_to_allowed_k<0><<<blocks, 256, 0, s>>>(wave, allowed, rows, ld, T, Wd, first_tile, threshold, keep_argmax);
    else if (dtype == 1) wave_to_allowed_k<1><<<blocks, 256, 0, s>>>(wave, allowed, rows, ld, T, Wd, first_tile, threshold, keep_argmax);
    else wave_to_allowed_k<2><<<blocks, 256, 0, s>>>(wave, allowed, rows, ld, T, Wd, first_tile, threshold, keep_argmax);
    WFD_CHECK_LAUNCH_P("wave_to_allowed");
    return 0;
}

int wfd_propagator_create(const uint32_t* conn_bits, int T, wfd_propagator** out) {
    if (!conn_bits || !out) {
        set_error("wfd_propagator_create: null argument");
        return -1;
    }
    if (T < 1 || (T + 31) / 32 > PROP_MAX_WD) {
        set_error("wfd_propagator_create: T = %d outside [1, %d]", T, PROP_MAX_WD * 32);
        return -3;
    }
    const int Wd = (T + 31) / 32;
    // supporter lists: for direction d and target tile t, the source tiles s with conn[d][s][t] set
    std::vector<uint32_t> dense;
    std::vector<int> ptr(static_cast<size_t>(4) * T + 1, 0);
    std::vector<uint16_t> col;
    int n_dense = 0;
    for (int d = 0; d < 4; ++d) {
        std::vector<std::vector<uint16_t>> sup(T);
        for (int s_ = 0; s_ < T; ++s_) {
            const uint32_t* row = conn_bits + (static_cast<size_t>(d) * T + s_) * Wd;
            for (int w = 0; w < Wd; ++w) {
                uint32_t v = row[w];
                while (v) {
                    const int t = w * 32 + __builtin_ctz(v);
                    v &= v - 1;
                    if (t < T) sup[t].push_back(static_cast<uint16_t>(s_));
                }
            }
        }
        for (int t = 0; t < T; ++t) {
            ptr[static_cast<size_t>(d) * T + t] = static_cast<int>(col.size());
            if (sup[t].size() > 255 && n_dense < 65535) {  // kept as a bit row (everything supports the null tile)
                col.push_back(0xFFFF);
                col.push_back(static_cast<uint16_t>(n_dense++));
                std::vector<uint32_t> bits(Wd, 0u);
                for (uint16_t s_ : sup[t]) bits[s_ >> 5] |= 1u << (s_ & 31);
                dense.insert(dense.end(), bits.begin(), bits.end());
            } else {
                col.insert(col.end(), sup[t].begin(), sup[t].end());
            }
        }
    }
    ptr[static_cast<size_t>(4) * T] = static_cast<int>(col.size());
    col.push_back(0);
    if (dense.empty()) dense.assign(Wd, 0u);
    wfd_propagator* p = new (std::nothrow) wfd_propagator();
    if (!p) {
        set_error("wfd_propagator_create: out of memory");
        return -2;
    }
    p->T = T;
    p->Wd = Wd;
    p->entries = static_cast<long long>(col.size()) - 1;
    p->n_dense = n_dense;
    cudaError_t e = cudaMalloc(&p->ptr, ptr.size() * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&p->col, col.size() * sizeof(uint16_t));
    if (e == cudaSuccess) e = cudaMalloc(&p->dense, dense.size() * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMemcpy(p->ptr, ptr.data(), ptr.size() * sizeof(int), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(p->col, col.data(), col.size() * sizeof(uint16_t), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(p->dense, dense.data(), dense.size() * sizeof(uint32_t), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        set_error("wfd_propagator_create: %s", cudaGetErrorString(e));
        wfd_propagator_destroy(p);
        return -8;
    }
    *out = p;
    return 0;
}

void wfd_propagator_destroy(wfd_propagator* p) {
    if (!p) return;
    if (p->ptr) cudaFree(p->ptr);
    if (p->col) cudaFree(p->col);
    if (p->dense) cudaFree(p->dense);
    if (p->scratch) cudaFree(p->scratch);
    delete p;
}

long long wfd_propagator_entries(const wfd_propagator* p) { return p ? p->entries : -1; }

int wfd_propagator_run(wfd_propagator* p, uint32_t* allowed, const uint8_t* boundary, int B, int H, int W, int max_sweeps,
                       int* sweeps_done, int* converged, void* stream) {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (!p || !allowed) {
        set_error("wfd_propagator_run: null argument");
        return -1;
    }
    const int T = p->T, Wd = p->Wd;
    if (int rc = check_shape("wfd_propagator_run", B, H, W, T)) return rc;
    if (max_sweeps < 1 || max_sweeps > PROP_MAX_SWEEPS) {
        set_error("wfd_propagator_run: max_sweeps %d outside [1, %d]", max_sweeps, PROP_MAX_SWEEPS);
        return -3;
    }
    const size_t cells = static_cast<size_t>(B) * H * W;
    const size_t flags_bytes = ((2 * cells + 255) / 256) * 256 + PROP_MAX_SWEEPS * sizeof(unsigned int);
    const size_t need = flags_bytes + cells * Wd * sizeof(uint32_t);
    cudaError_t e = cudaSuccess;
    if (p->scratch_bytes < need) {  // grow-only: flags + the second state buffer
        if (p->scratch) {
            e = cudaStreamSynchronize(s);
            cudaFree(p->scratch);
            p->scratch = nullptr;
            p->scratch_bytes = 0;
        }
        if (e == cudaSuccess) e = cudaMalloc(&p->scratch, need);
        if (e != cudaSuccess) {
            set_error("wfd_propagator_run: %s", cudaGetErrorString(e));
            return -8;
        }
        p->scratch_bytes = need;
    }
    uint8_t* dirty = static_cast<uint8_t*>(p->scratch);
    unsigned int* changed = reinterpret_cast<unsigned int*>(dirty + ((2 * cells + 255) / 256) * 256);
    uint32_t* other = reinterpret_cast<uint32_t*>(dirty + flags_bytes);
    e = cudaMemsetAsync(p->scratch, 0, flags_bytes, s);
    // cells that are never evaluated (boundary) must read the same in both buffers
    if (e == cudaSuccess) e = cudaMemcpyAsync(other, allowed, cells * Wd * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s);
    if (e != cudaSuccess) {
        set_error("wfd_propagator_run: %s", cudaGetErrorString(e));
        return -8;
    }
    PropRules rules;
    rules.ptr = p->ptr;
    rules.col = p->col;
    rules.dense = p->dense;
    const dim3 grid(static_cast<unsigned int>(H) * W, B);
    const int check_every = 4;  // sweeps between two reads of the change counters (a finished run makes the rest no-ops)
    int done = 0, conv = 0;
    unsigned int host_changed[check_every];
    while (done < max_sweeps && !conv) {
        const int first = done;
        for (int i = 0; i < check_every && done < max_sweeps; ++i, ++done) {
            ProfScope prof_scope("wfc_propagate_sweep", s);
            uint8_t* din = dirty + (done & 1) * cells;
            uint8_t* dout = dirty + ((done + 1) & 1) * cells;
            const uint32_t* sin = (done & 1) ? other : allowed;  // sweep k reads buffer k & 1 (0 = the caller's)
            uint32_t* sout = (done & 1) ? allowed : other;
            if (Wd <= PROP_THREADS)
                wfc_propagate_sweep_k<1><<<grid, PROP_THREADS, 0, s>>>(rules, sin, sout, boundary, din, dout, changed, done, H, W, T, Wd);
            else
                wfc_propagate_sweep_k<2><<<grid, PROP_THREADS, 0, s>>>(rules, sin, sout, boundary, din, dout, changed, done, H, W, T, Wd);
            WFD_CHECK_LAUNCH_P("wfc_propagate_sweep");
        }
        e = cudaMemcpyAsync(host_changed, changed + first, (done - first) * sizeof(unsigned int), cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) {
            set_error("wfd_propagator_run: %s", cudaGetErrorString(e));
            return -8;
        }
        for (int i = 0; i < done - first; ++i)
            if (host_changed[i] == 0u) {
                conv = 1;
                done = first + i + 1;  // the sweep that found nothing to change is the last one that did any work
                break;
            }
    }
    // After a sweep without changes both buffers agree on every cell; otherwise (max_sweeps ran out) the newest state is
    // the output of the last sweep, which sits in the second buffer when the sweep count is odd.
    if (!conv && (done & 1)) {
        e = cudaMemcpyAsync(allowed, other, cells * Wd * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s);
        if (e != cudaSuccess) {
            set_error("wfd_propagator_run: %s", cudaGetErrorString(e));
            return -8;
        }
    }
    if (sweeps_done) *sweeps_done = done;
    if (converged) *converged = conv;
    return 0;
}

int wfd_wfc_allowed_count(const uint32_t* allowed, int B, int H, int W, int T, int* count, void* stream) {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    ProfScope prof_scope("allowed_count", s);
    if (!allowed || !count) {
        set_error("wfd_wfc_allowed_count: null argument");
        return -1;
    }
    if (int rc = check_shape("wfd_wfc_allowed_count", B, H, W, T)) return rc;
    const long long cells = static_cast<long long>(B) * H * W;
    allowed_count_k<<<static_cast<unsigned int>((cells + 7) / 8), 256, 0, s>>>(allowed, count, cells, (T + 31) / 32);
    WFD_CHECK_LAUNCH_P("allowed_count");
    return 0;
}

}  // extern "C"
