// Probe 2: which part of the TMA sequence is at fault.  usage: tma_probe2 <impl: p|c> <x0> <fence: m|a> <boxw>
//   impl p = inline PTX (as in ame_kernels.cu), c = libcu++ cuda::device::experimental API
#include <cuda.h>
#include <cuda/barrier>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
namespace cde = cuda::device::experimental;
using barrier = cuda::barrier<cuda::thread_scope_block>;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe_ptx(const __grid_constant__ CUtensorMap tm, uint16_t *out, int x0, int y0, int fenceAsync, int boxw) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint16_t *win = reinterpret_cast<uint16_t *>(smem);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 256 * 16 * 2);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1) : "memory");
        if (fenceAsync) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        else asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(boxw * 16 * 2) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(win)), "l"(&tm),
                     "r"(smem_u32(bar)), "r"(x0), "r"(y0)
                     : "memory");
    }
    asm volatile(
        "{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(smem_u32(bar)), "r"(0)
        : "memory");
    for (int i = threadIdx.x; i < boxw * 16; i += blockDim.x) out[i] = win[i];
}

__global__ void probe_cxx(const __grid_constant__ CUtensorMap tm, uint16_t *out, int x0, int y0, int boxw) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint16_t *win = reinterpret_cast<uint16_t *>(smem);
    __shared__ barrier bar;
    if (threadIdx.x == 0) {
        init(&bar, blockDim.x);
        cde::fence_proxy_async_shared_cta();
    }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        cde::cp_async_bulk_tensor_2d_global_to_shared(win, &tm, x0, y0, bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, boxw * 16 * 2);
    } else {
        token = bar.arrive();
    }
    bar.wait(std::move(token));
    for (int i = threadIdx.x; i < boxw * 16; i += blockDim.x) out[i] = win[i];
}

int main(int argc, char **argv) {
    const char impl = argc > 1 ? argv[1][0] : 'p';
    const int x0 = argc > 2 ? atoi(argv[2]) : 37;
    const int fenceAsync = argc > 3 ? argv[3][0] == 'a' : 0;
    const int boxw = argc > 4 ? atoi(argv[4]) : 96;
    const int W = 736, H = 560, y0 = 11;
    std::vector<uint16_t> h((size_t)W * H);
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) h[(size_t)y * W + x] = (uint16_t)((x * 7 + y * 13) & 1023);
    uint16_t *d, *out;
    cudaMalloc(&d, h.size() * 2);
    cudaMalloc(&out, 256 * 16 * 2);
    cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                 const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    CUtensorMap m;
    const cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H};
    const cuuint64_t strides[1] = {(cuuint64_t)W * 2};
    const cuuint32_t box[2] = {(cuuint32_t)boxw, 16}, es[2] = {1, 1};
    CUresult r = ((EncodeFn)fn)(&m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    const size_t smem = 256 * 16 * 2 + 16;
    if (impl == 'p') probe_ptx<<<1, 128, smem>>>(m, out, x0, y0, fenceAsync, boxw);
    else probe_cxx<<<1, 128, smem>>>(m, out, x0, y0, boxw);
    cudaError_t e = cudaDeviceSynchronize();
    printf("impl %c x0 %d fence %s boxw %d encode %d: %s", impl, x0, fenceAsync ? "proxy.async" : "mbarrier_init", boxw, (int)r, cudaGetErrorString(e));
    if (e != cudaSuccess) { printf("\n"); return 1; }
    std::vector<uint16_t> o(boxw * 16);
    cudaMemcpy(o.data(), out, o.size() * 2, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int y = 0; y < 16; y++)
        for (int x = 0; x < boxw; x++) bad += o[y * boxw + x] != h[(size_t)(y0 + y) * W + x0 + x];
    printf(", %d mismatches\n", bad);
    return bad != 0;
}
