// Sparse form of the WFC loss contraction: the same a[n, i] that wfc.cu obtains from the dense tensor-core
// product,
//     a[n,i] = sum_j Ax[i,j] p[right(n),j] + Ax[j,i] p[left(n),j] + Ay[i,j] p[below(n),j] + Ay[j,i] p[above(n),j],
// as a gather over the allowed pairs only (0.15 - 0.55 % of the T x T entries for the StarCraft tilesets): per cell
// 89 316 additions for ice instead of 4 x 3074^2 multiply-adds. CUDA cores + shared memory, no tensor cores; what bounds it
// is the shared-memory gather rate (two 32-bit gathers per pair and four cells, random banks), not HBM and not arithmetic.
//
// One CTA = four consecutive cells of a map row. The 14 wave rows they need (six of their own row, four above, four
// below; zero outside the map) are staged in shared memory as seven arrays of [T_pad] 32-bit words, each word the 16-bit
// values of two horizontally adjacent cells, so that one gather serves two cells:
//     array 0 = (x0-1, x0)   1 = (x0+1, x0+2)   2 = (x0+3, x0+4)   3 / 4 = row above (x0, x0+1) / (x0+2, x0+3)   5 / 6 = row below
// For the cell pair (x0, x0+1) the left neighbours are array 0, the right ones array 1, above 3, below 5; for the pair
// (x0+2, x0+3) it is always the NEXT array. A list entry is therefore one 16-bit word index `array * T_pad + j`.
// The lists are sliced ELL: output rows are cut into virtual rows of at most 128 entries (the longest ice row has 666),
// sorted by length, 32 virtual rows per slice stored in blocks of eight entries per lane, so that a warp reads 512
// contiguous bytes per block and its lanes run the same trip count. Virtual rows add into an fp32 staging row in shared memory, which is then rounded to
// the 16-bit `a` (coalesced) and dotted with the cells' own wave rows for rowdot (fixed point 2^40, like the dense path).
#include "decoder.cuh"
#include "kernels.cuh"
#include <cuda_fp16.h>
#include <algorithm>
#include <vector>

namespace wfd {

constexpr int SP_THREADS = 512;
constexpr int SP_CELLS = 4;
constexpr int SP_CAP = 128;  // entries per virtual row
constexpr int SP_BLK = 8;    // entries of one lane stored together (a 16-byte read)

#define SP_CUDA_CK(expr)                                                        \
    do {                                                                        \
        cudaError_t e__ = (expr);                                               \
        if (e__ != cudaSuccess) {                                               \
            set_error("%s: %s", #expr, cudaGetErrorString(e__));                \
            return -8;                                                          \
        }                                                                       \
    } while (0)

WfcSparse::~WfcSparse() {
    if (ent) cudaFree(ent);
    if (slice_off) cudaFree(slice_off);
    if (row_of) cudaFree(row_of);
    if (len_of) cudaFree(len_of);
}

int WfcSparse::init(const uint8_t* mx, const uint8_t* my, int T_, int T_pad_) {
    T = T_;
    T_pad = T_pad_;
    if (7LL * T_pad > 65535) {
        set_error("wfc sparse: 7 * T_pad = %lld does not fit the 16-bit list entries", 7LL * T_pad);
        return -3;
    }
    struct VRow {
        int row;
        std::vector<uint16_t> e;
    };
    std::vector<VRow> vrows;
    nnz = 0;
    for (int i = 0; i < T; ++i) {
        std::vector<uint16_t> e;
        for (int j = 0; j < T; ++j) {
            if (mx[static_cast<size_t>(j) * T + i]) e.push_back(static_cast<uint16_t>(0 * T_pad + j));  // left:   Ax[j,i]
            if (mx[static_cast<size_t>(i) * T + j]) e.push_back(static_cast<uint16_t>(1 * T_pad + j));  // right:  Ax[i,j]
            if (my[static_cast<size_t>(j) * T + i]) e.push_back(static_cast<uint16_t>(3 * T_pad + j));  // above:  Ay[j,i]
            if (my[static_cast<size_t>(i) * T + j]) e.push_back(static_cast<uint16_t>(5 * T_pad + j));  // below:  Ay[i,j]
        }
        nnz += static_cast<long long>(e.size());
        for (size_t o = 0; o < e.size(); o += SP_CAP) {
            VRow v;
            v.row = i;
            v.e.assign(e.begin() + o, e.begin() + std::min(e.size(), o + SP_CAP));
            vrows.push_back(std::move(v));
        }
    }
    std::stable_sort(vrows.begin(), vrows.end(), [](const VRow& a, const VRow& b) { return a.e.size() > b.e.size(); });
    n_slices = static_cast<int>((vrows.size() + 31) / 32);
    if (n_slices == 0) n_slices = 1;
    std::vector<int> off(n_slices + 1, 0);
    std::vector<uint16_t> rows(static_cast<size_t>(n_slices) * 32, 0), lens(static_cast<size_t>(n_slices) * 32, 0), ents;
    for (int s = 0; s < n_slices; ++s) {
        int L = 0;
        for (int l = 0; l < 32; ++l) {
            const size_t v = static_cast<size_t>(s) * 32 + l;
            if (v < vrows.size()) {
                rows[v] = static_cast<uint16_t>(vrows[v].row);
                lens[v] = static_cast<uint16_t>(vrows[v].e.size());
                L = std::max(L, static_cast<int>(vrows[v].e.size()));
            }
        }
        L = (L + SP_BLK - 1) / SP_BLK * SP_BLK;  // whole blocks of eight entries: one 16-byte read per lane and block
        off[s] = static_cast<int>(ents.size());
        ents.resize(ents.size() + static_cast<size_t>(L) * 32, 0);
        for (int l = 0; l < 32; ++l) {
            const size_t v = static_cast<size_t>(s) * 32 + l;
            if (v >= vrows.size()) continue;
            for (size_t k = 0; k < vrows[v].e.size(); ++k)
                ents[off[s] + ((k / SP_BLK) * 32 + l) * SP_BLK + k % SP_BLK] = vrows[v].e[k];
        }
    }
    off[n_slices] = static_cast<int>(ents.size());
    padded = static_cast<long long>(ents.size());
    if (ents.empty()) ents.push_back(0);
    SP_CUDA_CK(cudaMalloc(&ent, ents.size() * sizeof(uint16_t)));
    SP_CUDA_CK(cudaMalloc(&slice_off, off.size() * sizeof(int)));
    SP_CUDA_CK(cudaMalloc(&row_of, rows.size() * sizeof(uint16_t)));
    SP_CUDA_CK(cudaMalloc(&len_of, lens.size() * sizeof(uint16_t)));
    SP_CUDA_CK(cudaMemcpy(ent, ents.data(), ents.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    SP_CUDA_CK(cudaMemcpy(slice_off, off.data(), off.size() * sizeof(int), cudaMemcpyHostToDevice));
    SP_CUDA_CK(cudaMemcpy(row_of, rows.data(), rows.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    SP_CUDA_CK(cudaMemcpy(len_of, lens.data(), lens.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    return 0;
}

template <bool BF16>
__device__ __forceinline__ float2 unpack2(uint32_t w) {
    if (BF16) return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
    return __half22float2(*reinterpret_cast<const __half2*>(&w));
}
template <bool BF16>
__device__ __forceinline__ uint32_t pack2f(float a, float b) {
    if (BF16) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
        return *reinterpret_cast<const uint32_t*>(&h);
    }
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}

template <bool BF16>
__global__ void __launch_bounds__(SP_THREADS, 1)
wfc_sparse_k(const uint16_t* __restrict__ wave, uint16_t* __restrict__ a_out, acc_t* __restrict__ rowdot,
             const uint16_t* __restrict__ ent, const int* __restrict__ slice_off, const uint16_t* __restrict__ row_of,
             const uint16_t* __restrict__ len_of, int n_slices, int H, int W, int T_pad, int strips) {
    extern __shared__ __align__(16) uint32_t sp_smem[];
    uint32_t* P = sp_smem;                                         // [7][T_pad] pairs of cells
    float* A = reinterpret_cast<float*>(sp_smem + 7 * T_pad);      // [4][T_pad] fp32 staging of a
    __shared__ float red[SP_CELLS][SP_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int sx = blockIdx.x % strips;
    const int y = (blockIdx.x / strips) % H;
    const int b = blockIdx.x / (strips * H);
    const int x0 = sx * SP_CELLS;
    const long long plane = static_cast<long long>(b) * H * W;
    // ---- stage the seven pair arrays (32-bit global reads = two channels of one cell; 64-bit shared stores)
    const int half_t = T_pad >> 1;
#pragma unroll 1
    for (int arr = 0; arr < 7; ++arr) {
        const int yy = arr < 3 ? y : (arr < 5 ? y - 1 : y + 1);
        const int xa = arr < 3 ? x0 - 1 + 2 * arr : x0 + 2 * ((arr - 3) & 1), xb = xa + 1;
        const bool row_ok = yy >= 0 && yy < H;
        const bool oka = row_ok && xa >= 0 && xa < W, okb = row_ok && xb >= 0 && xb < W;
        const uint32_t* ra = reinterpret_cast<const uint32_t*>(wave + (plane + static_cast<long long>(yy) * W + xa) * T_pad);
        const uint32_t* rb = reinterpret_cast<const uint32_t*>(wave + (plane + static_cast<long long>(yy) * W + xb) * T_pad);
        uint2* dst = reinterpret_cast<uint2*>(P + arr * T_pad);
#pragma unroll 4  // eight row reads in flight per thread: the staging is L2-latency bound otherwise
        for (int j2 = tid; j2 < half_t; j2 += SP_THREADS) {
            const uint32_t va = oka ? __ldg(ra + j2) : 0u, vb = okb ? __ldg(rb + j2) : 0u;
            dst[j2] = make_uint2((va & 0xffffu) | (vb << 16), (va >> 16) | (vb & 0xffff0000u));
        }
    }
    for (int i = tid; i < SP_CELLS * T_pad; i += SP_THREADS) A[i] = 0.f;
    __syncthreads();
    // ---- gather: one virtual row per lane, 32 per slice, slices dealt round-robin to the warps (they are sorted by length)
    for (int s = warp; s < n_slices; s += SP_THREADS / 32) {
        const int base = __ldg(slice_off + s), nblk = (__ldg(slice_off + s + 1) - base) / (32 * SP_BLK);
        const int my_len = __ldg(len_of + s * 32 + lane);
        const int row = __ldg(row_of + s * 32 + lane);
        float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
        // eight entries of this lane per 16-byte read (512 contiguous bytes per warp), the next block in flight while the
        // current one is gathered: the list comes from L2 and its latency is what an unpipelined loop waits for
        const uint4* ep = reinterpret_cast<const uint4*>(ent + base) + lane;
        uint4 cur = nblk > 0 ? __ldg(ep) : make_uint4(0u, 0u, 0u, 0u);
        for (int kb = 0; kb < nblk; ++kb) {
            const uint4 nxt = kb + 1 < nblk ? __ldg(ep + (kb + 1) * 32) : make_uint4(0u, 0u, 0u, 0u);
            const uint32_t wds[4] = {cur.x, cur.y, cur.z, cur.w};
            const int k0 = kb * SP_BLK;
#pragma unroll
            for (int u = 0; u < SP_BLK; ++u) {
                if (k0 + u < my_len) {
                    const uint32_t e = (u & 1) ? (wds[u >> 1] >> 16) : (wds[u >> 1] & 0xffffu);
                    const float2 pa = unpack2<BF16>(P[e]), pb = unpack2<BF16>(P[e + T_pad]);
                    acc0 += pa.x;
                    acc1 += pa.y;
                    acc2 += pb.x;
                    acc3 += pb.y;
                }
            }
            cur = nxt;
        }
        if (my_len > 0) {
            atomicAdd(A + 0 * T_pad + row, acc0);
            atomicAdd(A + 1 * T_pad + row, acc1);
            atomicAdd(A + 2 * T_pad + row, acc2);
            atomicAdd(A + 3 * T_pad + row, acc3);
        }
    }
    __syncthreads();
    // ---- store a (16-bit, coalesced) and dot it with the cells' own wave rows: cell c is (array, half) =
    // (0, hi), (1, lo), (1, hi), (2, lo)
    float dot[SP_CELLS] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < SP_CELLS; ++c) {
        if (x0 + c >= W) continue;
        const uint32_t* own = P + ((c + 1) >> 1) * T_pad;
        const bool hi = ((c + 1) & 1) != 0;
        uint32_t* dst = reinterpret_cast<uint32_t*>(a_out + (plane + static_cast<long long>(y) * W + x0 + c) * T_pad);
        const float* src = A + c * T_pad;
        for (int j2 = tid; j2 < half_t; j2 += SP_THREADS) {
            const float v0 = src[2 * j2], v1 = src[2 * j2 + 1];
            dst[j2] = pack2f<BF16>(v0, v1);
            const float2 p0 = unpack2<BF16>(own[2 * j2]), p1 = unpack2<BF16>(own[2 * j2 + 1]);
            dot[c] = fmaf(hi ? p0.y : p0.x, v0, dot[c]);
            dot[c] = fmaf(hi ? p1.y : p1.x, v1, dot[c]);
        }
    }
#pragma unroll
    for (int c = 0; c < SP_CELLS; ++c) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dot[c] += __shfl_xor_sync(0xffffffffu, dot[c], o);
        if (lane == 0) red[c][warp] = dot[c];
    }
    __syncthreads();
    if (tid < SP_CELLS && x0 + tid < W) {
        float sum = 0.f;
        for (int w = 0; w < SP_THREADS / 32; ++w) sum += red[tid][w];
        rowdot[plane + static_cast<long long>(y) * W + x0 + tid] = __float2ll_rn(sum * WFD_FX_GRAD);
    }
}

size_t WfcSparse::smem_bytes() const { return static_cast<size_t>(7 + SP_CELLS) * T_pad * 4; }

int WfcSparse::contract(const void* wave, void* a, acc_t* rowdot, int B, int H, int W, int bf16, cudaStream_t s) const {
    ProfScope prof_scope("wfc_sparse", s);
    if (!ent) {
        set_error("wfc sparse: lists not built");
        return -3;
    }
    const size_t smem = smem_bytes();
    if (smem + 1024 > 227 * 1024 || (T_pad & 1)) {
        set_error("wfc sparse: T_pad = %d needs %zu bytes of shared memory per CTA (max 227 KB)", T_pad, smem);
        return -3;
    }
    const int strips = (W + SP_CELLS - 1) / SP_CELLS;
    const long long blocks = static_cast<long long>(B) * H * strips;
    auto kern = bf16 ? wfc_sparse_k<true> : wfc_sparse_k<false>;
    static size_t attr_set[2] = {0, 0};  // largest dynamic size granted so far (the kernel also has 256 static bytes)
    if (smem > attr_set[bf16 ? 1 : 0]) {
        const cudaError_t ea = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (ea != cudaSuccess) {
            (void)cudaGetLastError();  // not sticky for the launches that follow
            set_error("wfc sparse: %zu bytes of shared memory per CTA refused (T_pad = %d): %s", smem, T_pad, cudaGetErrorString(ea));
            return -8;
        }
        attr_set[bf16 ? 1 : 0] = smem;
    }
    kern<<<static_cast<unsigned int>(blocks), SP_THREADS, smem, s>>>(static_cast<const uint16_t*>(wave), static_cast<uint16_t*>(a),
                                                                     rowdot, ent, slice_off, row_of, len_of, n_slices, H, W, T_pad, strips);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("wfc_sparse launch: %s", cudaGetErrorString(e));
        return -7;
    }
    ++g_launch_count;
    if (debug_sync_enabled()) {
        int rc = debug_sync_check("wfc_sparse", s);
        if (rc < 0) return rc;
    }
    return 0;
}

}  // namespace wfd
