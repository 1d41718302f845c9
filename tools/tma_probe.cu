// Standalone probe: 2-D TMA tile load (box 40 x 39 of 4-byte elements) with the descriptor (a) as a
// __grid_constant__ kernel parameter and (b) in global memory. Build: nvcc -gencode arch=compute_100a,code=sm_100a
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <vector>

#define BW 40
#define BH 39
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ int run_tma(const void* tmap, int c0, int c1, uint32_t* out, bool fence_tm) {
    extern __shared__ __align__(128) unsigned char sm[];
    uint32_t* buf = (uint32_t*)sm;
    unsigned long long* bar = (unsigned long long*)(sm + 6272);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (fence_tm) asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(tmap) : "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(BW * BH * 4) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(s32(buf)), "l"(tmap), "r"(s32(bar)), "r"(c0), "r"(c1) : "memory");
    }
    int ok = 0;
    for (int spin = 0; spin < (1 << 22); spin++) {
        uint32_t done;
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(s32(bar)), "r"(0) : "memory");
        if (done) { ok = 1; break; }
    }
    if (ok) for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) out[i] = buf[i];
    return ok;
}

__global__ void k_param(const __grid_constant__ CUtensorMap tm, int c0, int c1, uint32_t* out, int* status) {
    int ok = run_tma(&tm, c0, c1, out, false);
    if (threadIdx.x == 0) *status = ok;
}
__global__ void k_global(const CUtensorMap* tm, int c0, int c1, uint32_t* out, int* status, int fence_tm) {
    int ok = run_tma(tm, c0, c1, out, fence_tm != 0);
    if (threadIdx.x == 0) *status = ok;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
    void* f = nullptr; cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
    printf("entry point: %d %d %p\n", (int)e, (int)q, f);
    EncodeTiledFn enc = (EncodeTiledFn)f;
    const int W = 100, H = 77, P = 100;
    std::vector<uint32_t> h(P * H);
    for (int i = 0; i < P * H; i++) h[i] = i;
    uint32_t *d, *out; int* st; CUtensorMap* dtm;
    cudaMalloc(&d, P * H * 4); cudaMalloc(&out, BW * BH * 4); cudaMalloc(&st, 4); cudaMalloc(&dtm, sizeof(CUtensorMap));
    cudaMemcpy(d, h.data(), P * H * 4, cudaMemcpyHostToDevice);
    CUtensorMap tm;
    cuuint64_t gdim[2] = {W, H}; cuuint64_t gstr[1] = {P * 4}; cuuint32_t box[2] = {BW, BH}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode: %d\n", (int)r);
    cudaMemcpy(dtm, &tm, sizeof(tm), cudaMemcpyHostToDevice);
    std::vector<uint32_t> o(BW * BH);
    int hs;
    for (int mode = 0; mode < 3; mode++) {
        for (int c0 : {0, 8, 72}) {
            cudaMemset(out, 0xff, BW * BH * 4); cudaMemset(st, 0, 4);
            if (mode == 0) k_param<<<1, 128, 6272 + 16>>>(tm, c0, 3, out, st);
            else k_global<<<1, 128, 6272 + 16>>>(dtm, c0, 3, out, st, mode == 2);
            cudaError_t ce = cudaDeviceSynchronize();
            cudaMemcpy(&hs, st, 4, cudaMemcpyDeviceToHost);
            cudaMemcpy(o.data(), out, BW * BH * 4, cudaMemcpyDeviceToHost);
            int bad = 0;
            for (int y = 0; y < BH; y++) for (int x = 0; x < BW; x++) {
                uint32_t exp = (c0 + x < W && 3 + y < H) ? (uint32_t)((3 + y) * P + c0 + x) : 0u;
                if (o[y * BW + x] != exp) bad++;
            }
            printf("mode %d c0 %d: cuda=%s done=%d mismatches=%d\n", mode, c0, cudaGetErrorString(ce), hs, bad);
            if (ce != cudaSuccess) return 1;
        }
    }
    return 0;
}
