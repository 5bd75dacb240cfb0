/*
 * Fused DeepSeek-MLA decoder attention half-layer for sm_100a.
 *
 * Replaces the reference op `deepseek_decoder_layer(input, W_q_nope, W_q_pe, W_uk, W_kv_nope, W_k_pe, W_uv, W_o, ckv_cache,
 * rms_input_weight, rms_ckv_weight, cos, sin) -> [1, hidden]` (/root/reference/include/H100/deepseek/kernel.cuh:9-697,
 * deepseek_kernel_dispatch.cu:4-242, include/pybind.cpp:45-59, :113; shapes config.h:1-8).  Same operator, same weight
 * layouts ([in, out], as the reference's tensor maps read them), different decomposition.
 *
 * The reference runs one 4-CTA cluster per head: every cluster recomputes the shared ckv / k_pe projection (16x the W_kv
 * reads), walks the WHOLE latent cache for its own head (16x the cache reads, CUDA cores, fp16 accumulators), and is
 * fixed at SEQ_LEN 4096 at compile time.  Here one call is three back-to-back kernels on the caller's stream, chained
 * with programmatic dependent launch so that each kernel's weight tiles are already in shared memory when the data it
 * depends on arrives; every byte of weights and cache leaves HBM once:
 *
 *  1. ds_proj_kernel   16 clusters of 8 CTAs, cluster = head, 108 KB of shared memory so that all 16 clusters are resident at
 *     once (at one CTA per SM a B200 holds only 15 clusters of 8: ncu launch__cluster_max_active).  CTA (h, r) takes hidden
 *     rows [256r, 256r+256) of the head's W_q_nope / W_q_pe columns.  The 8 partial q vectors are summed with
 *     cluster_reduce<8, LINEAR> over distributed shared memory (include/dsm.cuh); then CTA r computes latent columns
 *     [64r, 64r+64) of q_lat = q_nope . W_uk[head] (tile fetched into the consumed W_q_nope bytes during the exchange) and
 *     writes them to the workspace (the "all-gather" of the reference, :391-398, becomes a plain store because the next
 *     kernel is not organised by head).
 *  2. ds_attn_kernel   128 CTAs, CTA = a contiguous slice of cache rows, ALL 16 heads at once: the heads are the M = 16
 *     rows of `mma.sync.m16n8k16`, so the cache is read once (not once per head) and QK^T / PV run on the tensor cores
 *     from 128-byte-swizzled TMA boxes (ldmatrix conflict-free).  Each CTA leaves an un-normalised flash-decode state
 *     (m, l, o[16][512]) in the workspace.  The same CTA c also multiplies hidden rows [16c, 16c+16) into the shared W_kv /
 *     W_k_pe columns, so the ckv / k_pe projection is computed ONCE per call, split 128 ways (fp32 `red.global.add.v4`).
 *  3. ds_out_kernel    16 clusters of 8, cluster = head.  Every CTA finishes the current token from the complete sums (fp16
 *     ckv -> RMSNorm, RoPE on k_pe, its head's score): one more flash-decode state.  CTA (h, r) merges the 129 states for latent columns
 *     [64r, 64r+64) of head h, multiplies by its 64 rows of W_uv[head], the 8 partial head outputs are summed with
 *     cluster_reduce<8, LINEAR>, then every CTA multiplies the head output by W_o[128h..128h+128, 256r..256r+256) and adds
 *     its fp32 partial to the workspace; the last of the 16 heads to arrive on a column slice rounds it to fp16.
 *
 * fp16 rounding points are the reference's (xn, q, ckv, k_pe, ckv_n, q_lat, o_lat, head output); every sum is fp32.
 * Algorithmic bytes per call: 27.5 MB of weights + seq_len * 1152 B of cache (4.7 MB at 4096).
 * Workspace (cf_deepseek_workspace_bytes): zero once at allocation; one call at a time per workspace.
 */
#pragma once

#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/dsm.cuh"
#include "llama_decoder_kernel.cuh"

namespace cfb {

constexpr int DS_HIDDEN = 2048;
constexpr int DS_HEADS = 16;
constexpr int DS_NOPE = 128;
constexpr int DS_ROPE = 64;
constexpr int DS_LORA = 512;
constexpr int DS_MLA = DS_LORA + DS_ROPE;          // 576 columns per cache row
constexpr int DS_CLUSTER = 8;
constexpr int DS_THREADS = 256;
constexpr int DS_SPLITS = 128;                     // cache slices = CTAs of the attention kernel
constexpr int DS_STATES = DS_SPLITS + 1;           // + the current token
constexpr int DS_TILE_ROWS = 32;                   // cache rows per TMA tile (two 16-row mma groups)
constexpr int DS_STAGES = 4;
constexpr int DS_TILE_BYTES = DS_TILE_ROWS * DS_MLA * 2;      // 36864: nine {64 cols, 32 rows} swizzled boxes
constexpr int DS_Q_STRIDE = DS_MLA + 8;            // halves; 1168-byte rows -> ldmatrix rows land in 8 different 16-byte bank groups
constexpr int DS_KSLICE = DS_HIDDEN / DS_CLUSTER;  // 256 hidden rows per CTA of a head cluster
constexpr int DS_KV_ROWS = DS_HIDDEN / (DS_HEADS * DS_CLUSTER);   // 16 hidden rows of W_kv / W_k_pe per CTA

constexpr uint32_t DS_FLAG_ROPE_SCORES = 0x100;    // == CF_DS_FLAG_ROPE_SCORES

struct DsParams {
    CUtensorMap tm_wq_nope, tm_wq_pe, tm_wuk, tm_wkv, tm_wk_pe, tm_cache, tm_wuv, tm_wo;
    const __half* x;
    const __half* rms_in_w;
    const __half* rms_ckv_w;
    const float* cos;
    const float* sin;
    __half* out;
    __half* ckv_new;          // optional [512]: the current token's normalised latent
    __half* k_pe_new;         // optional [64]:  the current token's rotated k_pe
    float* ckv_acc;           // ws [576]  fp32 sums of the shared projection, zero between calls
    __half* q;                // ws [16][576] fp16: q_lat ++ rotated q_pe (zeros unless DS_FLAG_ROPE_SCORES)
    float* part_ml;           // ws [129][16][2]  (m in the log2 domain, l)
    float* part_o;            // ws [128][16][512] ++ [512] (the current token's latent, shared by the heads)
    float* out_acc;           // ws [2048], zero between calls
    unsigned* counters;       // ws [8], zero between calls
    int n_rows;               // cache rows that take part: seq_len - 1
    int rows_per_split;       // multiple of DS_TILE_ROWS
    int n_stages;             // tile-ring depth of the attention kernel: 1 or DS_STAGES
    float eps;
    float scale_log2;         // log2(e) / sqrt(192)
    uint32_t flags;
};

// rotate-half RoPE over 64 dims, the reference's index pattern (kernel.cuh:300-320)
__device__ __forceinline__ float ds_rope(const float* v, const float* cos, const float* sin, int i) {
    return i < DS_ROPE / 2 ? v[i] * cos[i] - v[i + DS_ROPE / 2] * sin[i + DS_ROPE / 2]
                           : v[i] * cos[i] + v[i - DS_ROPE / 2] * sin[i - DS_ROPE / 2];
}

// RMSNorm of the whole 2048-wide input row by one 256-thread CTA: xn = fp16(x * rsqrt(mean(x^2) + eps) * w), w8 = this
// thread's 8 weights.  Ends without a barrier: the caller synchronises before reading xn.
__device__ __forceinline__ void ds_rmsnorm_row(const __half* x, const float (&w8)[8], float eps, __half* xn, float* misc,
                                               uint32_t tid, uint32_t lane, uint32_t warp) {
    float f[8];
    unpack8(*reinterpret_cast<const uint4*>(x + tid * 8), f);
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) ss = fmaf(f[k], f[k], ss);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (lane == 0) misc[warp] = ss;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < DS_THREADS / 32; ++i) t += misc[i];
    const float rstd = rsqrtf(t / (float)DS_HIDDEN + eps);
    __align__(16) __half h[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) h[k] = __float2half_rn(f[k] * rstd * w8[k]);
    *reinterpret_cast<uint4*>(xn + tid * 8) = *reinterpret_cast<const uint4*>(h);
}

// ------------------------------------------------------------------------------------------------------------------
// 1. projections
// ------------------------------------------------------------------------------------------------------------------
struct SmemDsProj {
    // 108 KB: two CTAs of this kernel -- or one of this and one of the previous call's output kernel -- fit on an SM.  That
    // matters twice: at one CTA per SM only 15 clusters of 8 are co-resident on a B200 (GPC sizes), so the 16 head clusters
    // would run in TWO waves; and with CF_FLAG_PDL the weight tiles stream in while the previous kernel is still running.
    // The W_q_nope tile is consumed first; its bytes then hold the W_uk tile (fetched while the q_pe product and the cluster
    // exchange run), the exchange buffer and the exchange source.  Peers are held back by the cluster barrier, which this CTA
    // arrives at only after its last read of the tile.
    static constexpr int WQ = 0;                                   // [256][128] halves
    static constexpr int WQPE = WQ + DS_KSLICE * DS_NOPE * 2;      // [256][64]
    static constexpr int XN = WQPE + DS_KSLICE * DS_ROPE * 2;      // [2048] halves
    static constexpr int RED = XN + DS_HIDDEN * 2;                 // 2048 floats of cross-thread scratch
    static constexpr int MISC = RED + 2048 * 4;                    // warp sums
    static constexpr int BARS = MISC + 64;                         // 3 mbarriers
    static constexpr int TOTAL = BARS + 64;
    // aliases inside the consumed W_q_nope tile
    static constexpr int WUK = WQ;                                 // [128][64] halves
    static constexpr int RECV = WUK + DS_NOPE * 64 * 2;            // 8 x 192 floats: one exchange, phase 0 only
    static constexpr int SRC = RECV + DS_CLUSTER * 192 * 4;        // 192 floats
    static_assert(SRC + 192 * 4 <= WQPE, "aliases must fit in the W_q_nope tile");
    static_assert(2 * (TOTAL + 1024) <= 228 * 1024, "two CTAs per SM");
};

__global__ void __launch_bounds__(DS_THREADS, 2)
ds_proj_kernel(const __grid_constant__ DsParams p)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t rank = dsm::cluster_ctarank();
    const uint32_t head = blockIdx.x / DS_CLUSTER;
    const uint32_t sb = dsm::smem_u32(smem);
    const uint32_t bar_q = sb + SmemDsProj::BARS, bar_uk = bar_q + 8, bar_x = bar_q + 16;
    const __half* wq = reinterpret_cast<const __half*>(smem + SmemDsProj::WQ);
    const __half* wqpe = reinterpret_cast<const __half*>(smem + SmemDsProj::WQPE);
    const __half* wuk = reinterpret_cast<const __half*>(smem + SmemDsProj::WUK);
    __half* xn = reinterpret_cast<__half*>(smem + SmemDsProj::XN);
    float* red = reinterpret_cast<float*>(smem + SmemDsProj::RED);
    float* src = reinterpret_cast<float*>(smem + SmemDsProj::SRC);
    float* recv = reinterpret_cast<float*>(smem + SmemDsProj::RECV);
    float* misc = reinterpret_cast<float*>(smem + SmemDsProj::MISC);

    if (tid == 0) {
        dsm::mbar_init(bar_q, 1);
        dsm::mbar_init(bar_uk, 1);
        cluster_reduce_arm<DS_CLUSTER>(bar_x, 192 * 4);
        dsm::mbar_fence_init();
        // weights do not depend on the previous kernel: fetch this CTA's q tiles now
        const uint64_t pol = policy_evict_first();
        dsm::mbar_arrive_expect_tx(bar_q, DS_KSLICE * (DS_NOPE + DS_ROPE) * 2);
        tma_load_2d(sb + SmemDsProj::WQ, &p.tm_wq_nope, head * DS_NOPE, rank * DS_KSLICE, bar_q, pol);
        tma_load_2d(sb + SmemDsProj::WQPE, &p.tm_wq_pe, head * DS_ROPE, rank * DS_KSLICE, bar_q, pol);
    }
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    float w8[8];
    unpack8(*reinterpret_cast<const uint4*>(p.rms_in_w + tid * 8), w8);
    asm volatile("griddepcontrol.wait;" ::: "memory");

    // ---- RMSNorm of the whole input row (4 KB; every CTA needs all of it) ----
    ds_rmsnorm_row(p.x, w8, p.eps, xn, misc, tid, lane, warp);
    __syncthreads();

    // ---- q_nope partial over this CTA's 256 hidden rows: 16 column groups of 8 x 16 row groups of 16 ----
    dsm::mbar_wait(bar_q, 0);
    {
        const int cg = tid & 15, kg = tid >> 4;
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
        for (int i = 0; i < 16; ++i) {
            const int k = kg * 16 + i;
            const float xk = __half2float(xn[rank * DS_KSLICE + k]);
            float w[8];
            unpack8(*reinterpret_cast<const uint4*>(wq + k * DS_NOPE + cg * 8), w);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = fmaf(xk, w[j], acc[j]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) red[kg * DS_NOPE + cg * 8 + j] = acc[j];
    }
    __syncthreads();                             // the W_q_nope tile is dead: its bytes become W_uk / recv / src
    if (tid == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic reads of the tile before the async-proxy write
        dsm::mbar_arrive_expect_tx(bar_uk, DS_NOPE * 64 * 2);
        tma_load_2d(sb + SmemDsProj::WUK, &p.tm_wuk, head * DS_LORA + rank * 64, 0, bar_uk, policy_evict_first());
    }
    dsm::cluster_arrive();                       // peers may push into this CTA from here on (exchange barrier armed at the top)
    if (tid < DS_NOPE) {
        float t = 0.f;
#pragma unroll
        for (int g = 0; g < 16; ++g) t += red[g * DS_NOPE + tid];
        src[tid] = t;
    }
    __syncthreads();
    {
        // q_pe partial: 8 column groups of 8 x 32 row groups of 8
        const int cg = tid & 7, kg = tid >> 3;
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
        for (int i = 0; i < 8; ++i) {
            const int k = kg * 8 + i;
            const float xk = __half2float(xn[rank * DS_KSLICE + k]);
            float w[8];
            unpack8(*reinterpret_cast<const uint4*>(wqpe + k * DS_ROPE + cg * 8), w);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = fmaf(xk, w[j], acc[j]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) red[kg * DS_ROPE + cg * 8 + j] = acc[j];
    }
    __syncthreads();
    if (tid < DS_ROPE) {
        float t = 0.f;
#pragma unroll
        for (int g = 0; g < 32; ++g) t += red[g * DS_ROPE + tid];
        src[DS_NOPE + tid] = t;
    }
    // ---- sum the 8 hidden slices over distributed shared memory: src[0..192) = q_nope ++ q_pe of this head, on every CTA ----
    dsm::cluster_wait();
    uint32_t phase = 0;
    cluster_reduce<DS_CLUSTER, Stage::LINEAR_DEEPSEEK>(192 * 4, tid, 192, rank, dsm::smem_u32(src), dsm::smem_u32(recv), bar_x,
                                                       phase, src, recv);
    if (tid < 192) src[tid] = round_h(src[tid]);
    __syncthreads();

    // ---- q_lat columns [64 rank, 64 rank + 64) = q_nope . W_uk[:, head*512 + ...] ----
    dsm::mbar_wait(bar_uk, 0);
    {
        const int cg = tid & 7, kg = tid >> 3;                   // 8 column groups of 8 x 32 row groups of 4
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int k = kg * 4 + i;
            const float qk = src[k];
            float w[8];
            unpack8(*reinterpret_cast<const uint4*>(wuk + k * 64 + cg * 8), w);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = fmaf(qk, w[j], acc[j]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) red[kg * 64 + cg * 8 + j] = acc[j];
    }
    __syncthreads();
    if (tid < 64) {
        float t = 0.f;
#pragma unroll
        for (int g = 0; g < 32; ++g) t += red[g * 64 + tid];
        p.q[head * DS_MLA + rank * 64 + tid] = __float2half_rn(t);
    } else if (tid < 128 && rank == 0) {
        const int i = tid - 64;
        const float v = (p.flags & DS_FLAG_ROPE_SCORES) ? ds_rope(src + DS_NOPE, p.cos, p.sin, i) : 0.f;
        p.q[head * DS_MLA + DS_LORA + i] = __float2half_rn(v);
    }
    // keep this CTA's shared memory alive until every peer's pushes have landed and been consumed
    dsm::cluster_arrive();
    dsm::cluster_wait();
}

// ------------------------------------------------------------------------------------------------------------------
// 2. attention over the latent cache: all heads per CTA on the tensor cores
// ------------------------------------------------------------------------------------------------------------------
struct SmemDsAttn {
    // [n_stages x 36864 tile ring, 1024-aligned][q][W_kv / W_k_pe rows][xn][misc][barriers].  The ring has ONE stage when every
    // CTA has at most one tile (seq_len <= 4097): 78 KB per CTA, so the CTAs of this kernel become resident NEXT TO the
    // projection kernel's (108 KB) and their tiles are in flight while that kernel still runs; otherwise DS_STAGES.
    static constexpr int Q_BYTES = DS_HEADS * DS_Q_STRIDE * 2;               // [16][DS_Q_STRIDE] halves
    static constexpr int WKV_BYTES = DS_KV_ROWS * DS_MLA * 2;                // [2 boxes][16][256] ++ [16][64] halves
    static constexpr int XN_BYTES = DS_HIDDEN * 2;
    static constexpr int TAIL = Q_BYTES + WKV_BYTES + XN_BYTES + 64 + 64;
    static constexpr int MERGE_STRIDE = DS_HEADS * 128 + 2 * DS_HEADS;       // floats per (column block) slot of the warp-pair merge
    static_assert(4 * MERGE_STRIDE * 4 <= DS_TILE_BYTES, "merge scratch reuses the first stage of the tile ring");
    static_assert(DS_STAGES + 1 <= 8, "barriers: one per stage + one for the projection rows");
    __host__ __device__ static constexpr int total(int n_stages) { return n_stages * DS_TILE_BYTES + TAIL; }
};

__device__ __forceinline__ void ds_issue_tile(const DsParams& p, uint32_t dst, uint32_t bar, int row, uint64_t pol) {
    dsm::mbar_arrive_expect_tx(bar, DS_TILE_BYTES);
#pragma unroll
    for (int j = 0; j < DS_MLA / 64; ++j) tma_load_2d(dst + j * (DS_TILE_ROWS * 128), &p.tm_cache, j * 64, row, bar, pol);
}

__global__ void __launch_bounds__(DS_THREADS, 1)
ds_attn_kernel(const __grid_constant__ DsParams p)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t sb = dsm::smem_u32(smem);
    const int n_stages = p.n_stages;
    const int q_off = n_stages * DS_TILE_BYTES;
    const int wkv_off = q_off + SmemDsAttn::Q_BYTES;
    const int xn_off = wkv_off + SmemDsAttn::WKV_BYTES;
    __half* qs = reinterpret_cast<__half*>(smem + q_off);
    const __half* wkv = reinterpret_cast<const __half*>(smem + wkv_off);
    const __half* wkpe = wkv + DS_KV_ROWS * DS_LORA;
    __half* xn = reinterpret_cast<__half*>(smem + xn_off);
    float* misc = reinterpret_cast<float*>(smem + xn_off + SmemDsAttn::XN_BYTES);
    const uint32_t bars = sb + xn_off + SmemDsAttn::XN_BYTES + 64;
    const uint32_t bar_kv = bars + 8 * DS_STAGES;
    const int split = blockIdx.x;
    const int row0 = split * p.rows_per_split;
    const int row1 = min(row0 + p.rows_per_split, p.n_rows);
    const int ntiles = row1 > row0 ? (row1 - row0 + DS_TILE_ROWS - 1) / DS_TILE_ROWS : 0;

    if (tid == 0) {
        for (int s = 0; s < n_stages; ++s) dsm::mbar_init(bars + 8 * s, 1);
        dsm::mbar_init(bar_kv, 1);
        dsm::mbar_fence_init();
        // The cache and the weights are inputs of the CALL, not products of the projection kernel: their tiles may be fetched
        // before the dependency resolves.  (Without CF_FLAG_PDL the projection kernel -- and so this one -- starts only after
        // everything earlier on the stream has finished; with it the caller promises that no kernel in flight writes the cache.)
        const uint64_t pol = policy_evict_first();
        for (int t = 0; t < min(ntiles, n_stages); ++t)
            ds_issue_tile(p, sb + t * DS_TILE_BYTES, bars + 8 * t, row0 + t * DS_TILE_ROWS, pol);
        dsm::mbar_arrive_expect_tx(bar_kv, SmemDsAttn::WKV_BYTES);
        tma_load_2d(sb + wkv_off, &p.tm_wkv, 0, split * DS_KV_ROWS, bar_kv, pol);
        tma_load_2d(sb + wkv_off + DS_KV_ROWS * 256 * 2, &p.tm_wkv, 256, split * DS_KV_ROWS, bar_kv, pol);
        tma_load_2d(sb + wkv_off + DS_KV_ROWS * DS_LORA * 2, &p.tm_wk_pe, 0, split * DS_KV_ROWS, bar_kv, pol);
    }
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    float w8[8];
    unpack8(*reinterpret_cast<const uint4*>(p.rms_in_w + tid * 8), w8);
    __syncthreads();
    asm volatile("griddepcontrol.wait;" ::: "memory");           // q comes from the projection kernel, x from whatever preceded the call
    // q of all heads -> padded rows in shared memory (72 16-byte chunks per head)
    for (int i = tid; i < DS_HEADS * (DS_MLA / 8); i += DS_THREADS) {
        const int h = i / (DS_MLA / 8), c = i - h * (DS_MLA / 8);
        *reinterpret_cast<uint4*>(qs + h * DS_Q_STRIDE + c * 8) = *reinterpret_cast<const uint4*>(p.q + h * DS_MLA + c * 8);
    }

    // ---- shared ckv / k_pe projection, computed ONCE per call and split 128 ways: this CTA multiplies hidden rows
    //      [16 split, 16 split + 16) into all 576 columns and adds its fp32 partial to the workspace.  It lives here rather than
    //      in the projection kernel to keep that kernel at two CTAs per SM; the sums are complete when this kernel is. ----
    ds_rmsnorm_row(p.x, w8, p.eps, xn, misc, tid, lane, warp);
    __syncthreads();                                             // xn and qs visible
    dsm::mbar_wait(bar_kv, 0);
    if (tid < DS_MLA / 4) {
        const __half* base = tid < 128 ? wkv + (tid >> 6) * (DS_KV_ROWS * 256) + (tid & 63) * 4 : wkpe + (tid - 128) * 4;
        const int stride = tid < 128 ? 256 : DS_ROPE;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
        for (int i = 0; i < DS_KV_ROWS; ++i) {
            const float xk = __half2float(xn[split * DS_KV_ROWS + i]);
            const uint2 u = *reinterpret_cast<const uint2*>(base + i * stride);
            const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
            const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
            a0 = fmaf(xk, lo.x, a0); a1 = fmaf(xk, lo.y, a1); a2 = fmaf(xk, hi.x, a2); a3 = fmaf(xk, hi.y, a3);
        }
        red_add_v4(p.ckv_acc + tid * 4, make_float4(a0, a1, a2, a3));
    }

    // ---- flash-decode over this CTA's rows.  Warp (rg, cb): rows rg*16..rg*16+15 of each tile, output columns cb*128.. ----
    const int rg = warp & 1, cb = warp >> 1;
    const int g4 = lane >> 2, t4 = lane & 3, lrow = lane & 7, lmat = lane >> 3;
    float oacc[16][4];
#pragma unroll
    for (int t = 0; t < 16; ++t) { oacc[t][0] = 0.f; oacc[t][1] = 0.f; oacc[t][2] = 0.f; oacc[t][3] = 0.f; }
    float mrun[2] = {-INFINITY, -INFINITY}, lrun[2] = {0.f, 0.f};     // heads g4 and g4 + 8; l is a quad-partial
    const uint32_t q_lane = sb + q_off + ((lmat & 1) * 8 + lrow) * (DS_Q_STRIDE * 2) + (lmat >> 1) * 16;
    for (int t = 0; t < ntiles; ++t) {
        const int s = t % n_stages;
        dsm::mbar_wait(bars + 8 * s, (t / n_stages) & 1);
        const uint32_t st = sb + s * DS_TILE_BYTES;
        const int rows_left = row1 - (row0 + t * DS_TILE_ROWS + rg * 16);
        if (rows_left > 0) {
            float sc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll 4
            for (int ks = 0; ks < DS_MLA / 16; ++ks) {
                uint32_t a[4];
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                             : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(q_lane + ks * 32));
                const int key = rg * 16 + (lmat >> 1) * 8 + lrow, chunk = (ks & 3) * 2 + (lmat & 1);
                const uint32_t addr = st + (ks >> 2) * (DS_TILE_ROWS * 128) + key * 128 + ((chunk ^ (key & 7)) << 4);
                uint32_t b0, b1, b2, b3;
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                             : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3) : "r"(addr));
                mma16816(sc[0], a, b0, b1);
                mma16816(sc[1], a, b2, b3);
            }
            // online softmax: this lane holds keys 2t4, 2t4+1 (n-tile 0) and 8+2t4, 9+2t4 (n-tile 1) of heads g4 ([0],[1]) and g4+8 ([2],[3])
            const bool ok0 = 2 * t4 < rows_left, ok1 = 2 * t4 + 1 < rows_left, ok2 = 8 + 2 * t4 < rows_left, ok3 = 9 + 2 * t4 < rows_left;
            uint32_t pa[4];
            float corr[2];
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                float s4[4];
                s4[0] = ok0 ? sc[0][2 * hh] * p.scale_log2 : -INFINITY;
                s4[1] = ok1 ? sc[0][2 * hh + 1] * p.scale_log2 : -INFINITY;
                s4[2] = ok2 ? sc[1][2 * hh] * p.scale_log2 : -INFINITY;
                s4[3] = ok3 ? sc[1][2 * hh + 1] * p.scale_log2 : -INFINITY;
                float mx = fmaxf(fmaxf(s4[0], s4[1]), fmaxf(s4[2], s4[3]));
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
                const float m_new = fmaxf(mrun[hh], mx);
                const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
                corr[hh] = dsm::exp2_diff(mrun[hh], m_use);
                mrun[hh] = m_new;
                const __half2 p01 = __floats2half2_rn(dsm::fast_exp2(s4[0] - m_use), dsm::fast_exp2(s4[1] - m_use));
                const __half2 p23 = __floats2half2_rn(dsm::fast_exp2(s4[2] - m_use), dsm::fast_exp2(s4[3] - m_use));
                // the row sum uses the same fp16-rounded probabilities that multiply the rows
                lrun[hh] = lrun[hh] * corr[hh] + (__low2float(p01) + __high2float(p01)) + (__low2float(p23) + __high2float(p23));
                pa[hh] = *reinterpret_cast<const uint32_t*>(&p01);
                pa[2 + hh] = *reinterpret_cast<const uint32_t*>(&p23);
            }
            // O = O * corr + P . rows  (this warp's 128 latent columns = boxes 2cb, 2cb+1 of the tile)
#pragma unroll
            for (int jd = 0; jd < 8; ++jd) {
                const int key = rg * 16 + (lmat & 1) * 8 + lrow, chunk = (jd & 3) * 2 + (lmat >> 1);
                const uint32_t addr = st + (cb * 2 + (jd >> 2)) * (DS_TILE_ROWS * 128) + key * 128 + ((chunk ^ (key & 7)) << 4);
                uint32_t b0, b1, b2, b3;
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                             : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3) : "r"(addr));
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    oacc[2 * jd + u][0] *= corr[0]; oacc[2 * jd + u][1] *= corr[0];
                    oacc[2 * jd + u][2] *= corr[1]; oacc[2 * jd + u][3] *= corr[1];
                }
                mma16816(oacc[2 * jd], pa, b0, b1);
                mma16816(oacc[2 * jd + 1], pa, b2, b3);
            }
        }
        __syncthreads();                                   // every warp is done with stage s
        if (tid == 0 && t + n_stages < ntiles)
            ds_issue_tile(p, st, bars + 8 * s, row0 + (t + n_stages) * DS_TILE_ROWS, policy_evict_first());
    }
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
        lrun[hh] += __shfl_xor_sync(0xffffffffu, lrun[hh], 1);
        lrun[hh] += __shfl_xor_sync(0xffffffffu, lrun[hh], 2);
    }

    // ---- merge the two row groups of a column block through shared memory (the tile ring is idle now), publish the state ----
    float* slot = reinterpret_cast<float*>(smem) + cb * SmemDsAttn::MERGE_STRIDE;
    if (rg == 1) {
        if (t4 == 0) {
            slot[2 * g4] = mrun[0]; slot[2 * g4 + 1] = lrun[0];
            slot[2 * (g4 + 8)] = mrun[1]; slot[2 * (g4 + 8) + 1] = lrun[1];
        }
        float* o = slot + 2 * DS_HEADS;
#pragma unroll
        for (int nt = 0; nt < 16; ++nt) {
            *reinterpret_cast<float2*>(o + g4 * 128 + nt * 8 + t4 * 2) = make_float2(oacc[nt][0], oacc[nt][1]);
            *reinterpret_cast<float2*>(o + (g4 + 8) * 128 + nt * 8 + t4 * 2) = make_float2(oacc[nt][2], oacc[nt][3]);
        }
    }
    __syncthreads();
    if (rg == 0) {
        const float* o = slot + 2 * DS_HEADS;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const int h = g4 + hh * 8;
            const float m2 = slot[2 * h], l2 = slot[2 * h + 1];
            const float M = fmaxf(mrun[hh], m2);
            const float Mu = (M == -INFINITY) ? 0.f : M;
            const float w1 = dsm::exp2_diff(mrun[hh], Mu), w2 = dsm::exp2_diff(m2, Mu);
            float* go = p.part_o + ((size_t)split * DS_HEADS + h) * DS_LORA + cb * 128;
#pragma unroll
            for (int nt = 0; nt < 16; ++nt) {
                const float2 o2 = *reinterpret_cast<const float2*>(o + h * 128 + nt * 8 + t4 * 2);
                *reinterpret_cast<float2*>(go + nt * 8 + t4 * 2) =
                    make_float2(oacc[nt][2 * hh] * w1 + o2.x * w2, oacc[nt][2 * hh + 1] * w1 + o2.y * w2);
            }
            if (cb == 0 && t4 == 0) {
                p.part_ml[(split * DS_HEADS + h) * 2] = M;
                p.part_ml[(split * DS_HEADS + h) * 2 + 1] = lrun[hh] * w1 + l2 * w2;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// 3. state merge, W_uv, W_o
// ------------------------------------------------------------------------------------------------------------------
struct SmemDsOut {
    // 89 KB: see SmemDsProj.  The exchange buffer of the cluster reduction reuses the cross-thread scratch.
    static constexpr int WUV = 0;                                  // [64][128] halves
    static constexpr int WO = WUV + 64 * DS_NOPE * 2;              // [128][256] halves
    static constexpr int RED = WO + DS_NOPE * 256 * 2;             // 2048 floats
    static constexpr int RECV = RED;                               // 8 x 128 floats: one exchange, phase 0 only
    static constexpr int SRC = RED + 2048 * 4;                     // 128 floats
    static constexpr int OLAT = SRC + DS_NOPE * 4;                 // 64 floats
    static constexpr int WGT = OLAT + 64 * 4;                      // 132 floats: merge weights of the states (+ pad)
    static constexpr int TOK = WGT + 132 * 4;                      // 576 floats: the current token's cache row
    static constexpr int MISC = TOK + DS_MLA * 4;
    static constexpr int BARS = MISC + 64;
    static constexpr int TOTAL = BARS + 64;
    static_assert(SmemDsProj::TOTAL + TOTAL + 2048 <= 228 * 1024, "a projection CTA and an output CTA must fit on one SM");
};

__global__ void __launch_bounds__(DS_THREADS, 1)
ds_out_kernel(const __grid_constant__ DsParams p)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t rank = dsm::cluster_ctarank();
    const uint32_t head = blockIdx.x / DS_CLUSTER;
    const uint32_t sb = dsm::smem_u32(smem);
    const uint32_t bar_uv = sb + SmemDsOut::BARS, bar_o = bar_uv + 8, bar_x = bar_uv + 16;
    const __half* wuv = reinterpret_cast<const __half*>(smem + SmemDsOut::WUV);
    const __half* wo = reinterpret_cast<const __half*>(smem + SmemDsOut::WO);
    float* red = reinterpret_cast<float*>(smem + SmemDsOut::RED);
    float* src = reinterpret_cast<float*>(smem + SmemDsOut::SRC);
    float* recv = reinterpret_cast<float*>(smem + SmemDsOut::RECV);
    float* olat = reinterpret_cast<float*>(smem + SmemDsOut::OLAT);
    float* wgt = reinterpret_cast<float*>(smem + SmemDsOut::WGT);
    float* tok = reinterpret_cast<float*>(smem + SmemDsOut::TOK);
    float* misc = reinterpret_cast<float*>(smem + SmemDsOut::MISC);
    __shared__ unsigned s_last, s_zero;

    if (tid == 0) {
        dsm::mbar_init(bar_uv, 1);
        dsm::mbar_init(bar_o, 1);
        cluster_reduce_arm<DS_CLUSTER>(bar_x, DS_NOPE * 4);
        dsm::mbar_fence_init();
        const uint64_t pol = policy_evict_first();
        dsm::mbar_arrive_expect_tx(bar_uv, 64 * DS_NOPE * 2);
        tma_load_2d(sb + SmemDsOut::WUV, &p.tm_wuv, head * DS_NOPE, rank * 64, bar_uv, pol);
        dsm::mbar_arrive_expect_tx(bar_o, DS_NOPE * 256 * 2);
        tma_load_2d(sb + SmemDsOut::WO, &p.tm_wo, rank * 256, head * DS_NOPE, bar_o, pol);
    }
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");

    // ---- merge the flash-decode states of this head for latent columns [64 rank, 64 rank + 64): 128 cache slices from the
    //      attention kernel + the current token, which every CTA finishes for itself from the (now complete) fp32 sums of the
    //      shared projection: fp16 ckv -> RMSNorm, RoPE on k_pe, the head's score ----
    {
        // thread (j, g) folds states g, g+4, ... of column j: their loads are issued first, so that they are in flight
        // while the block works out the current token and the merge weights
        const int j = tid & 63, g = tid >> 6;
        const float* po = p.part_o + (size_t)head * DS_LORA + rank * 64 + j;
        constexpr int NPER = (DS_STATES + 3) / 4;               // 33; the last one of group 0 is the current token
        float v[NPER];
#pragma unroll
        for (int i = 0; i < NPER; ++i) {
            const int s = g + 4 * i;
            v[i] = s < DS_SPLITS ? po[(size_t)s * DS_HEADS * DS_LORA] : 0.f;
        }
        float m = -INFINITY, l = 0.f;
        if (tid < DS_SPLITS) {
            const float2 ml = *reinterpret_cast<const float2*>(p.part_ml + (tid * DS_HEADS + head) * 2);
            m = ml.x; l = ml.y;
        }
        const float c0 = round_h(p.ckv_acc[tid]), c1 = round_h(p.ckv_acc[tid + DS_THREADS]);
        const float kraw = tid < DS_ROPE ? round_h(p.ckv_acc[DS_LORA + tid]) : 0.f;
        const __half* qh = p.q + head * DS_MLA;
        const float q0 = __half2float(qh[tid]), q1 = __half2float(qh[tid + DS_THREADS]);
        const float q2 = tid < DS_ROPE ? __half2float(qh[DS_LORA + tid]) : 0.f;
        const float g0 = __half2float(p.rms_ckv_w[tid]), g1 = __half2float(p.rms_ckv_w[tid + DS_THREADS]);
        if (tid < DS_ROPE) tok[DS_LORA + tid] = kraw;                       // un-rotated k_pe, staged
        float ss = fmaf(c0, c0, c1 * c1), mx = m;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            ss += __shfl_xor_sync(0xffffffffu, ss, o);
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        if (lane == 0) { misc[warp] = ss; misc[8 + warp] = mx; }
        __syncthreads();
        // every thread of this CTA has consumed its words of the shared projection: count the CTA as done with it.  The
        // last CTA of the grid to get here clears the sums for the next call (nobody reads them after its own count).
        if (tid == 0) {
            __threadfence();
            s_zero = atomicAdd(p.counters + DS_CLUSTER, 1u) == gridDim.x - 1 ? 1u : 0u;
        }
        float t = 0.f, Mst = -INFINITY;
#pragma unroll
        for (int i = 0; i < DS_THREADS / 32; ++i) { t += misc[i]; Mst = fmaxf(Mst, misc[8 + i]); }
        const float rstd = rsqrtf(t / (float)DS_LORA + p.eps);
        const float n0 = round_h(c0 * rstd * g0), n1 = round_h(c1 * rstd * g1);
        const float kr = tid < DS_ROPE ? round_h(ds_rope(tok + DS_LORA, p.cos, p.sin, tid)) : 0.f;
        tok[tid] = n0;
        tok[tid + DS_THREADS] = n1;
        if (blockIdx.x == 0 && p.ckv_new) {
            p.ckv_new[tid] = __float2half_rn(n0);
            p.ckv_new[tid + DS_THREADS] = __float2half_rn(n1);
        }
        __syncthreads();                                   // every thread has read the un-rotated k_pe and the warp results
        if (tid < DS_ROPE) {
            tok[DS_LORA + tid] = kr;
            if (blockIdx.x == 0 && p.k_pe_new) p.k_pe_new[tid] = __float2half_rn(kr);
        }
        if (s_zero) {
            for (int i = tid; i < DS_MLA; i += DS_THREADS) p.ckv_acc[i] = 0.f;
            if (tid == 0) p.counters[DS_CLUSTER] = 0u;
        }
        // the head's score against the current token (q_pe is zero without DS_FLAG_ROPE_SCORES)
        float sc = fmaf(q0, n0, q1 * n1) + q2 * kr;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sc += __shfl_xor_sync(0xffffffffu, sc, o);
        if (lane == 0) misc[warp] = sc;
        __syncthreads();                                   // tok complete, score partials written
        float score = 0.f;
#pragma unroll
        for (int i = 0; i < DS_THREADS / 32; ++i) score += misc[i];
        const float m_new = score * p.scale_log2;
        const float M = fmaxf(Mst, m_new);                 // finite: the current token always takes part
        if (g == 0) v[NPER - 1] = tok[rank * 64 + j];      // state DS_SPLITS: o = the token's latent, l = 1
        const float w = tid < DS_SPLITS ? dsm::exp2_diff(m, M) : tid == DS_SPLITS ? dsm::exp2_diff(m_new, M) : 0.f;
        if (tid < 132) wgt[tid] = w;
        float wl = w * (tid == DS_SPLITS ? 1.f : l);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) wl += __shfl_xor_sync(0xffffffffu, wl, o);
        __syncthreads();                                   // misc read by everyone, wgt written
        if (lane == 0) misc[warp] = wl;
        __syncthreads();
        float L = 0.f;
#pragma unroll
        for (int i = 0; i < DS_THREADS / 32; ++i) L += misc[i];
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < NPER; ++i) acc = fmaf(wgt[g + 4 * i], v[i], acc);
        red[g * 64 + j] = acc;
        __syncthreads();
        if (tid < 64) olat[tid] = round_h((red[tid] + red[64 + tid] + red[128 + tid] + red[192 + tid]) / L);
    }
    __syncthreads();

    // ---- partial head output over this CTA's 64 latent rows of W_uv ----
    dsm::mbar_wait(bar_uv, 0);
    {
        const int cg = tid & 15, kg = tid >> 4;                 // 16 column groups of 8 x 16 row groups of 4
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int k = kg * 4 + i;
            const float ok = olat[k];
            float w[8];
            unpack8(*reinterpret_cast<const uint4*>(wuv + k * DS_NOPE + cg * 8), w);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = fmaf(ok, w[j], acc[j]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) red[kg * DS_NOPE + cg * 8 + j] = acc[j];
    }
    __syncthreads();
    if (tid < DS_NOPE) {
        float t = 0.f;
#pragma unroll
        for (int g = 0; g < 16; ++g) t += red[g * DS_NOPE + tid];
        src[tid] = t;
    }
    __syncthreads();                             // last read of `red`, which now becomes the exchange buffer
    dsm::cluster_arrive();
    dsm::cluster_wait();
    uint32_t phase = 0;
    cluster_reduce<DS_CLUSTER, Stage::ATTN_DEEPSEEK>(DS_NOPE * 4, tid, DS_NOPE, rank, dsm::smem_u32(src), dsm::smem_u32(recv),
                                                     bar_x, phase, src, recv);
    if (tid < DS_NOPE) src[tid] = round_h(src[tid]);
    __syncthreads();

    // ---- out[256 rank .. +256) += head output . W_o[128 head .. +128, 256 rank .. +256) ----
    dsm::mbar_wait(bar_o, 0);
    {
        const int cg = tid & 31, kg = tid >> 5;                 // 32 column groups of 8 x 8 row groups of 16
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
        for (int i = 0; i < 16; ++i) {
            const int k = kg * 16 + i;
            const float ak = src[k];
            float w[8];
            unpack8(*reinterpret_cast<const uint4*>(wo + k * 256 + cg * 8), w);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = fmaf(ak, w[j], acc[j]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) red[kg * 256 + cg * 8 + j] = acc[j];
    }
    __syncthreads();
    {
        float t = 0.f;
#pragma unroll
        for (int g = 0; g < 8; ++g) t += red[g * 256 + tid];
        asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p.out_acc + rank * 256 + tid), "f"(t) : "memory");
    }
    // the last of the 16 head clusters to arrive on column slice `rank` rounds it to fp16 and re-zeroes the accumulators
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(p.counters + rank, 1u) == DS_HEADS - 1 ? 1u : 0u;
    __syncthreads();
    if (s_last) {
        __threadfence();
        float* a = p.out_acc + rank * 256 + tid;
        const float v = __ldcg(a);
        p.out[rank * 256 + tid] = __float2half_rn(v);
        *a = 0.f;
        if (tid == 0) p.counters[rank] = 0u;
    }
    dsm::cluster_arrive();
    dsm::cluster_wait();
}

}  // namespace cfb
