// GPU probe (not product code): semantics of cp.async.bulk.tensor.2d...tile::gather4 on sm_100a, the one TMA form that
// fetches four ARBITRARY rows of a 2-D tensor with one request -- what a page-size-1 KV gather wants.
// Questions answered (one process per config, a faulting config must not poison the others):
//   argv[1] = box rows encoded in the tensor map (1, 4 or 16), argv[2] = 0 plain / 1 SWIZZLE_128B (box 64 halves wide),
//   argv[3] = byte offset of the destination inside a 1024-byte aligned buffer (0 or 512: are swizzle patterns a function
//   of the shared-memory ADDRESS, so that rows 4-7 of an 8-row swizzle atom can come from a second request?)
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tools/gather4_probe tools/gather4_probe.cu -lcuda
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

__global__ void probe(const __grid_constant__ CUtensorMap tm, int col0, int r0, int r1, int r2, int r3, int dst_off, int bytes,
                      unsigned char* out) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ __align__(8) unsigned long long bar;
    const unsigned bar_u = (unsigned)__cvta_generic_to_shared(&bar);
    const unsigned dst = (unsigned)__cvta_generic_to_shared(sm) + dst_off;
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = 0xEE;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_u));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_u), "r"(bytes) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
            " [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
            ::"r"(dst), "l"(&tm), "r"(col0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar_u) : "memory");
    }
    unsigned ok = 0;
    long long spins = 0;
    while (!ok && spins < (1 << 22)) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar_u) : "memory");
        ++spins;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) out[i] = sm[i];
    if (threadIdx.x == 0) out[4096] = ok ? 1 : 0;
}

int main(int argc, char** argv) {
    const int box_rows = argc > 1 ? atoi(argv[1]) : 1;
    const int swz = argc > 2 ? atoi(argv[2]) : 0;
    const int dst_off = argc > 3 ? atoi(argv[3]) : 0;
    const int ROWS = 512, COLS = 4096;                        // a KV pool: 512 slots x 32 heads x 128
    const int box_cols = swz ? 64 : 128;
    std::vector<__half> h((size_t)ROWS * COLS);
    for (int r = 0; r < ROWS; ++r)
        for (int c = 0; c < COLS; ++c) h[(size_t)r * COLS + c] = __float2half((float)((r * 7 + c) % 2048));
    __half* d; unsigned char* out;
    cudaMalloc(&d, h.size() * 2); cudaMalloc(&out, 8192);
    cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    auto enc = reinterpret_cast<CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill)>(fn);
    CUtensorMap tm;
    const cuuint64_t dims[2] = {(cuuint64_t)COLS, (cuuint64_t)1 << 24};       // extent far beyond the allocation, as the product does
    const cuuint64_t strides[1] = {(cuuint64_t)COLS * 2};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    const cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     swz ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("config box_rows=%d swizzle=%d dst_off=%d: encode -> %d\n", box_rows, swz, dst_off, (int)r);
    if (r != CUDA_SUCCESS) return 0;
    const int rows[4] = {5, 300, 3, 77}, col0 = 3 * 128;
    const int bytes = 4 * box_cols * 2;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192);
    probe<<<1, 128, 8192>>>(tm, col0, rows[0], rows[1], rows[2], rows[3], dst_off, bytes, out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("  kernel -> %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 0;
    std::vector<unsigned char> o(4097);
    cudaMemcpy(o.data(), out, 4097, cudaMemcpyDeviceToHost);
    printf("  barrier completed: %d\n", (int)o[4096]);
    const __half* oh = reinterpret_cast<const __half*>(o.data() + dst_off);
    // expected layout: row i of the gather at i * box_cols halves; with swizzle the 16-byte chunk c of row i (counted from the
    // 1024-byte aligned base, i.e. row index dst_off/128 + i) sits at chunk c ^ (row & 7)
    int bad_linear = 0, bad_swz = 0;
    for (int i = 0; i < 4; ++i)
        for (int c = 0; c < box_cols; ++c) {
            const float want = (float)((rows[i] * 7 + col0 + c) % 2048);
            if (__half2float(oh[i * box_cols + c]) != want) ++bad_linear;
            const int arow = dst_off / 128 + i, chunk = c / 8;
            const int pos = i * box_cols + ((chunk ^ (arow & 7)) * 8) + (c % 8);
            if (__half2float(oh[pos]) != want) ++bad_swz;
        }
    printf("  mismatches vs linear layout: %d, vs address-swizzled layout: %d (of %d)\n", bad_linear, bad_swz, 4 * box_cols);
    printf("  first halves of each gathered row:");
    for (int i = 0; i < 4; ++i) printf(" [%g %g]", __half2float(oh[i * box_cols]), __half2float(oh[i * box_cols + 1]));
    printf("\n");
    return 0;
}
