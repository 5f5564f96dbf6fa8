// Probe: cp.async.bulk.tensor.2d with the tensor map (a) in global memory, (b) as a __grid_constant__ parameter,
// (c) in global memory behind fence.proxy.tensormap.  usage: tma_probe a|b|c
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>
__global__ void probe(const void *tmapGlobal, const __grid_constant__ CUtensorMap tmapParam, uint16_t *out, int x0, int y0) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint16_t *win = reinterpret_cast<uint16_t *>(smem);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 96 * 32 * 2);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const void *tm = MODE == 1 ? (const void *)&tmapParam : tmapGlobal;
        if (MODE == 2) asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(tm) : "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(96 * 32 * 2) : "memory");
        for (int r = 0; r < 32; r += 16)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(win + r * 96)),
                         "l"(tm), "r"(smem_u32(bar)), "r"(x0), "r"(y0 + r)
                         : "memory");
    }
    asm volatile(
        "{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(smem_u32(bar)), "r"(0)
        : "memory");
    for (int i = threadIdx.x; i < 96 * 32; i += blockDim.x) out[i] = win[i];
}

int main(int argc, char **argv) {
    const char mode = argc > 1 ? argv[1][0] : 'a';
    const int W = 736, H = 560;
    std::vector<uint16_t> h((size_t)W * H);
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) h[(size_t)y * W + x] = (uint16_t)((x * 7 + y * 13) & 1023);
    uint16_t *d, *out;
    cudaMalloc(&d, h.size() * 2);
    cudaMalloc(&out, 96 * 32 * 2);
    cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                 const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    if (!fn) { printf("no cuTensorMapEncodeTiled\n"); return 2; }
    CUtensorMap m;
    const cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H};
    const cuuint64_t strides[1] = {(cuuint64_t)W * 2};
    const cuuint32_t box[2] = {96, 16}, es[2] = {1, 1};
    CUresult r = ((EncodeFn)fn)(&m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d\n", (int)r);
    void *dm;
    cudaMalloc(&dm, sizeof m);
    cudaMemcpy(dm, &m, sizeof m, cudaMemcpyHostToDevice);
    const int x0 = 37, y0 = 11;   // odd x on purpose
    const size_t smem = 96 * 32 * 2 + 16;
    if (mode == 'a') probe<0><<<1, 128, smem>>>(dm, m, out, x0, y0);
    if (mode == 'b') probe<1><<<1, 128, smem>>>(dm, m, out, x0, y0);
    if (mode == 'c') probe<2><<<1, 128, smem>>>(dm, m, out, x0, y0);
    cudaError_t e = cudaDeviceSynchronize();
    printf("mode %c: %s\n", mode, cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<uint16_t> o(96 * 32);
    cudaMemcpy(o.data(), out, o.size() * 2, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int y = 0; y < 32; y++)
        for (int x = 0; x < 96; x++) bad += o[y * 96 + x] != h[(size_t)(y0 + y) * W + x0 + x];
    printf("mode %c: %d mismatches\n", mode, bad);
    return bad != 0;
}
