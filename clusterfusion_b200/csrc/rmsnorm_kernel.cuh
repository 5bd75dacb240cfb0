/*
 * Standalone cluster RMSNorm for sm_100a:  out[b] = fp16( x[b] * rsqrt(mean(x[b]^2) + eps) * w ), fp32 math, one rounding.
 *
 * Replaces the reference op `rmsnorm(input, weight)` (/root/reference/include/H100/norm/kernel.cuh:8-76,
 * norm_kernel_dispatch.cu:4-26, pybind.cpp:61-64, :114): one cluster of 2 CTAs per row, each CTA holds half of the row in
 * registers, the two partial sums of squares are exchanged through distributed shared memory.  Differences:
 *  - any batch and any hidden that is a multiple of 16 up to 16384 (the reference binary is fixed at 64 x 8192),
 *  - the scalar exchange is the new cluster_reduce<2, Stage::LINEAR> of include/dsm.cuh (one st.async push per peer,
 *    rank-ordered fold -> bit-identical on both CTAs) instead of a remote atomicAdd between two cluster.sync(),
 *  - launches on the caller's stream, no device syncs, no memset of the output (reference: torch::full + 2 syncs).
 * HBM-bound elementwise op: 4 bytes moved per element; at 64 x 8192 that is 2 MB, i.e. launch-latency territory.
 */
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/dsm.cuh"
#include "llama_decoder_kernel.cuh"

namespace cfb {

constexpr int NORM_CLUSTER = 2;
constexpr int NORM_THREADS = 256;
constexpr int NORM_ITERS = 4;                 // 8-element chunks per thread: slice <= 256 * 8 * 4 = 8192 elements per CTA

__global__ void __launch_bounds__(NORM_THREADS)
rmsnorm_cluster_kernel(const __half* __restrict__ x, const __half* __restrict__ w, __half* __restrict__ out, int hidden, float eps)
{
    __shared__ __align__(16) float src[4];
    __shared__ __align__(16) float recv[2 * NORM_CLUSTER * 4];
    __shared__ __align__(8) uint64_t bar;
    __shared__ float warp_sums[NORM_THREADS / 32];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t rank = dsm::cluster_ctarank();
    const uint32_t row = blockIdx.x / NORM_CLUSTER;
    const int slice = hidden / NORM_CLUSTER;
    const __half* xr = x + (size_t)row * hidden + rank * slice;
    const __half* wr = w + rank * slice;
    __half* orow = out + (size_t)row * hidden + rank * slice;

    const uint32_t bar_u32 = dsm::smem_u32(&bar);
    if (tid == 0) {
        cluster_reduce_arm<NORM_CLUSTER>(bar_u32, 16);
        dsm::mbar_fence_init();
    }
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    // the weight does not depend on the previous kernel in the stream
    uint4 wv[NORM_ITERS];
#pragma unroll
    for (int it = 0; it < NORM_ITERS; ++it) {
        const int e = (it * NORM_THREADS + tid) * 8;
        wv[it] = e < slice ? *reinterpret_cast<const uint4*>(wr + e) : make_uint4(0, 0, 0, 0);
    }
    dsm::cluster_arrive();                    // the peer may push into this CTA only after the barrier is armed
    asm volatile("griddepcontrol.wait;" ::: "memory");

    float f[NORM_ITERS][8];
    float ss = 0.f;
#pragma unroll
    for (int it = 0; it < NORM_ITERS; ++it) {
        const int e = (it * NORM_THREADS + tid) * 8;
        const uint4 v = e < slice ? *reinterpret_cast<const uint4*>(xr + e) : make_uint4(0, 0, 0, 0);
        unpack8(v, f[it]);
#pragma unroll
        for (int k = 0; k < 8; ++k) ss = fmaf(f[it][k], f[it][k], ss);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (lane == 0) warp_sums[warp] = ss;
    __syncthreads();
    if (tid == 0) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < NORM_THREADS / 32; ++i) t += warp_sums[i];
        src[0] = t; src[1] = 0.f; src[2] = 0.f; src[3] = 0.f;
    }
    dsm::cluster_wait();
    uint32_t phase = 0;
    cluster_reduce<NORM_CLUSTER, Stage::LINEAR>(16, tid, 4, rank, dsm::smem_u32(src), dsm::smem_u32(recv), bar_u32, phase,
                                                src, recv);
    const float rstd = rsqrtf(src[0] / (float)hidden + eps);
#pragma unroll
    for (int it = 0; it < NORM_ITERS; ++it) {
        const int e = (it * NORM_THREADS + tid) * 8;
        if (e < slice) {
            float w8[8];
            unpack8(wv[it], w8);
            __align__(16) __half h[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) h[k] = __float2half_rn(f[it][k] * rstd * w8[k]);
            *reinterpret_cast<uint4*>(orow + e) = *reinterpret_cast<const uint4*>(h);
        }
    }
    // keep this CTA's shared memory alive until the peer's push has landed and been consumed on both sides
    dsm::cluster_arrive();
    dsm::cluster_wait();
}

}  // namespace cfb
