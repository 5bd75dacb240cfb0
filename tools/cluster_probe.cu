// GPU probe (not product code): how many thread-block clusters of a given size and shared-memory footprint can be
// co-resident on this device (cudaOccupancyMaxActiveClusters), and which SMs / GPC-like groups they land on.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tools/cluster_probe tools/cluster_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#include <algorithm>

__global__ void probe_kernel(unsigned* smid_out, unsigned long long* t_out, int spin_us) {
    extern __shared__ char s[];
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    if (threadIdx.x == 0) { smid_out[blockIdx.x] = smid; t_out[blockIdx.x] = t0; s[0] = 1; }
    do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); } while (t - t0 < (unsigned long long)spin_us * 1000ull);
}

int main() {
    cudaDeviceProp pr;
    cudaGetDeviceProperties(&pr, 0);
    printf("device %s, %d SMs, smem/block optin %zu\n", pr.name, pr.multiProcessorCount, pr.sharedMemPerBlockOptin);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    const int smems[] = {16 * 1024, 100 * 1024, 114 * 1024, 230 * 1024};
    for (int smem : smems) {
        cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        for (int cs : {1, 2, 4, 8, 16}) {
            for (int threads : {384}) {
                cudaLaunchConfig_t cfg{};
                cfg.gridDim = dim3(cs * 64, 1, 1);
                cfg.blockDim = dim3(threads, 1, 1);
                cfg.dynamicSmemBytes = smem;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeClusterDimension;
                at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                cfg.attrs = at; cfg.numAttrs = 1;
                int n = -1;
                cudaError_t e = cudaOccupancyMaxActiveClusters(&n, probe_kernel, &cfg);
                printf("smem %6d KB cluster %2d threads %d -> max active clusters %d (%d CTAs) %s\n", smem / 1024, cs, threads, n,
                       n * cs, e == cudaSuccess ? "" : cudaGetErrorString(e));
            }
        }
    }
    // empirical: launch 148/cs clusters of 230 KB, 200 us spin each; CTAs that start > 100 us late were not co-resident
    unsigned* d_smid; unsigned long long* d_t;
    cudaMalloc(&d_smid, 4096 * 4); cudaMalloc(&d_t, 4096 * 8);
    const int smem = 230 * 1024;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int cs : {1, 2, 4, 8, 16}) {
        const int ncl = 148 / cs;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(cs * ncl, 1, 1);
        cfg.blockDim = dim3(384, 1, 1);
        cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        for (int rep = 0; rep < 2; ++rep) {
            cudaError_t e = cudaLaunchKernelEx(&cfg, probe_kernel, d_smid, d_t, 200);
            if (e != cudaSuccess) { printf("launch cs=%d failed: %s\n", cs, cudaGetErrorString(e)); break; }
            cudaDeviceSynchronize();
        }
        std::vector<unsigned> smid(cs * ncl); std::vector<unsigned long long> t(cs * ncl);
        cudaMemcpy(smid.data(), d_smid, smid.size() * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(t.data(), d_t, t.size() * 8, cudaMemcpyDeviceToHost);
        unsigned long long t0 = *std::min_element(t.begin(), t.end());
        int first_wave = 0;
        for (auto v : t) if (v - t0 < 100000ull) ++first_wave;
        printf("cluster %2d: launched %3d CTAs (%d clusters), %3d in the first wave (%d clusters)\n", cs, cs * ncl, ncl, first_wave, first_wave / cs);
        if (cs >= 8) {
            for (int c = 0; c < ncl; ++c) {
                printf("   cluster %2d start +%6.1f us smids:", c, (t[c * cs] - t0) / 1000.0);
                for (int k = 0; k < cs; ++k) printf(" %3u", smid[c * cs + k]);
                printf("\n");
            }
        }
    }
    return 0;
}
