/*
 * wfd_b200_test.h — TEST HOOKS of libwfd_b200.so, kept out of the product ABI (include/wfd_b200.h). Both entry points
 * return -13 unless the process was started with WFD_ENABLE_TEST_HOOKS=1 (tests/conftest.py sets it): a product process
 * can neither reroute its contractions nor run hand-described ones by accident.
 */
#ifndef WFD_B200_TEST_H
#define WFD_B200_TEST_H

#include "wfd_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Runs one tapgemm contraction described field by field, either on the tcgen05 kernel (use_ref = 0) or on the CUDA-core
 * checking kernel (use_ref = 1). */
typedef struct {
    const void* a; int a_C, a_W, a_H, a_B; long long a_ld;     /* channels-last input [a_B, a_H, a_W, a_ld] */
    const void* b; int b_rows; long long ldb, b_zstride; int b_Z; /* K-major matrix [b_Z][b_rows][ldb] */
    int n_taps, cpt; int tap_dx[9], tap_dy[9];
    const int* a_c0;
    int BN, N, W, H, B, w_box, h_box, in_stride;
    int z_count, a_zmul, b_zmul, b_bbmul;
    void* out; int out_fp32; long long ldc, c_zstride;
    const float* bias; const void* residual; float alpha;
    void* rowdot; const void* dotsrc; long long ld_dot;   /* int64 fixed point (2^40) per output row */
    void* gn_sums; int gn_cpg;                              /* int64 fixed point (2^20) {sum, sumsq} per (sample, group) */
    int dtype16;
    int cg;           /* 0: library chooses, 1: single CTA, 2: CTA pair (cta_group::2) */
    int n_group;      /* >1: tile rasterisation in bands of n_group N tiles */
    int reuse;        /* 1: row-reuse schedule (3x3 stride-1 taps, b packed in (dx, cc, dy) chunk order) */
    long long* dbg;   /* optional device int64[8]: cycle counters of CTA 0 per warp role (profiling aid) */
    int a_mn_major;     /* A given transposed in memory: [a_B][a_C = K rows][a_W = M columns], M contiguous, pitch a_ld */
    const int* kc_list; /* optional K-chunk skip list (device), entries dx+1 | (dy+1)<<2 | chunk-in-tap<<4 | chunk<<16; */
    const int* kc_cnt;  /* [n_tiles] entries used per N tile (>= 1); rows n*n_taps*cpt.. of kc_list belong to N tile n */
} wfd_tapgemm_desc;
WFD_API int wfd_test_tapgemm(const wfd_tapgemm_desc* desc, int use_ref, void* stream);
/* when set to 1 every tensor-core contraction is routed to the CUDA-core checking kernel (bisecting aid of the parity tests) */
WFD_API int wfd_debug_set_ref_gemm(int on);


#ifdef __cplusplus
}
#endif
#endif /* WFD_B200_TEST_H */
