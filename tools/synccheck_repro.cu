// Minimal repro for the compute-sanitizer synccheck report "Barrier error detected. Divergent thread(s) in warp" that round 1
// saw at the first named barrier of the batched kernel, only for CTAs of the launch's SECOND wave (profiles/r02_sanitizer.txt).
// This kernel has no divergence at all: every thread executes straight-line code -- a cluster arrive, one named barrier
// (bar.sync 1, 384), a cluster wait.  Launch shape = the batched kernel's: clusters of 4, 384 threads, ~225 KB of dynamic shared
// memory (one CTA per SM), grid 256 > the 132 CTAs that are co-resident, so half of the clusters start when earlier ones retire.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tools/synccheck_repro tools/synccheck_repro.cu
//   compute-sanitizer --tool synccheck ./tools/synccheck_repro [clusters_in_x] [grid_y]
// Round 2 finding: the report follows blockIdx.y >= 1 (the batched kernels launch grid (heads*4, chunks)), not the second wave:
// run with grid_y = 2 to reproduce it on this kernel, which has no divergent code at all.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

__global__ void __launch_bounds__(384, 1) repro(float* out, int spin) {
    extern __shared__ __align__(16) unsigned char sm[];
    float* f = reinterpret_cast<float*>(sm);
    f[threadIdx.x] = (float)threadIdx.x;
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("bar.sync 1, 384;" ::: "memory");
    __syncthreads();
    float a = f[(threadIdx.x + 1) % 384];
    for (int i = 0; i < spin; ++i) a = a * 1.0001f + 0.5f;          // keep the first wave busy for a while (uniform trip count)
    asm volatile("bar.sync 1, 384;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (threadIdx.x == 0) out[blockIdx.y * gridDim.x + blockIdx.x] = a;
}

int main(int argc, char** argv) {
    const int clusters = argc > 1 ? atoi(argv[1]) : 64;
    const int grid_y = argc > 2 ? atoi(argv[2]) : 1;
    const int smem = 225 * 1024;
    float* out;
    cudaMalloc(&out, (size_t)clusters * 4 * grid_y * sizeof(float));
    cudaFuncSetAttribute(repro, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(clusters * 4, grid_y, 1);
    cfg.blockDim = dim3(384, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 4; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, repro, out, 20000);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    printf("synccheck repro: grid (%d, %d) in clusters of 4 x 384 threads, %d KB smem -> %s\n", clusters * 4, grid_y, smem / 1024, cudaGetErrorString(e));
    return 0;
}
