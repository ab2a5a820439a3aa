// SASS probe: one product chain and one squaring chain, to count instructions with cuobjdump before going to the GPU.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -cubin -o /tmp/sqr.cubin tools/sqr/sass_probe.cu -I zkt_plonk_b200/csrc
#include "ff.cuh"
using namespace zkb;
extern "C" __global__ void k_mul(fe_t *io, int n) {
    fe_t a = io[threadIdx.x], b = io[threadIdx.x + 32];
    for (int i = 0; i < n; ++i) a = fmul<FqP>(a, b);
    io[threadIdx.x] = a;
}
extern "C" __global__ void k_sqr(fe_t *io, int n) {
    fe_t a = io[threadIdx.x];
    for (int i = 0; i < n; ++i) a = fsqr_tri<FqP>(a);
    io[threadIdx.x] = a;
}
