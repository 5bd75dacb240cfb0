/*
 * Fused Llama FFN half-layer for sm_100a (B200), batch 1:
 *
 *   h = x + residual -> RMSNorm -> gate/up GEMV -> SiLU(g) * u -> down GEMV -> fp32 cross-CTA reduction
 *
 * SURVEY.md section 8 row f1.  The reference reserves a Stage::FFN in its cluster primitive and an FFN_DIM in its
 * config (/root/reference/include/dsm.cuh:140-153, include/H100/llama/config.h:4, :18-19) but ships no FFN kernel:
 * its FFN stays eager PyTorch (chat/llama/model.py:407-448, :519), which is 2/3 of the bytes of a decoded token.
 *
 * Decomposition (no cluster needed): the FFN intermediate dimension is cut into blocks of 16; CTA c owns blocks
 * c, c + grid, ...  For its blocks a CTA (1) streams the 16 gate rows and 16 up rows of W13 [2*ffn, hidden] and
 * reduces them against the normalised input, (2) applies SiLU * up, (3) streams the matching 16 rows of
 * W2^T [ffn, hidden] and accumulates its partial of ALL `hidden` outputs, (4) adds that partial into an fp32
 * global scratch; the last CTA to arrive converts to fp16.  Both GEMVs are therefore split along the intermediate
 * dimension and there is no dependency between CTAs until the final reduction (a two-kernel or grid-barrier
 * formulation would drain the HBM pipe in the middle of the layer).  W2 is taken TRANSPOSED ([ffn, hidden]) so that
 * a block's slice is 16 contiguous 8 KB rows instead of 4096 strided 32-byte fragments; the caller transposes once
 * at weight-load time, like the reference's own one-time fused-weight build (model.py:292-328).
 *
 * Machinery = the attention kernel's: 24 x 8 KB TMA stages, 12 self-issuing warp streams (tile g -> stage g % 24,
 * warp g % 12), every tile a [16 rows x 256 columns] box.  Per-warp results are combined with shared-memory fp32
 * atomics (row sums of the gate/up tiles, column sums of the down tiles).
 */
#pragma once

#include "llama_decoder_kernel.cuh"

namespace cfb {

constexpr int FFN_BLOCK = 16;          // intermediate dims per work block (= rows of one tile)
constexpr int FFN_NB_MAX = 13;         // max blocks per CTA (13 * 148 >= 28672 / 16; keeps shared memory under 227 KB)
constexpr int FFN_HIDDEN_MAX = 8192;

struct alignas(64) FfnParams {
    CUtensorMap tm_w13;     // [2*ffn][hidden] box {256, 16}
    CUtensorMap tm_w2t;     // [ffn][hidden]   box {256, 16}
    const __half* x;
    const __half* residual_in;
    const __half* rms_w;
    void* out;              // fp16 [hidden] (fp32 with flag bit 0)
    __half* residual_out;
    float* scratch;         // fp32 [hidden], zero between launches
    unsigned* counters;     // [2], zero between launches
    float eps;
    int hidden;
    int ffn;
    unsigned flags;
    int launch_id;          // trace only
};

struct SmemFfn {
    static constexpr int RING = 0;
    static constexpr int XS = RING + NSTAGES * STAGE_BYTES;                 // fp32[hidden] normalised input (phase 1)
    static constexpr int OUT_ACC = XS;                                      // fp32[hidden] this CTA's output partial (phase 2;
                                                                            //   xs is dead by then, a __syncthreads separates them)
    static constexpr int GU = XS + FFN_HIDDEN_MAX * 4;                      // fp32[NB_MAX][2][16] gate / up row sums
    static constexpr int ACT = GU + FFN_NB_MAX * 2 * FFN_BLOCK * 4;          // fp32[NB_MAX][16] silu(g) * u
    static constexpr int RED = ACT + FFN_NB_MAX * FFN_BLOCK * 4;             // fp32[32]
    static constexpr int BARS = RED + 32 * 4;                                // u64 full[NSTAGES]
    static constexpr int FLAGS = BARS + NSTAGES * 8;
    static constexpr int TOTAL = FLAGS + 16;
    static_assert(TOTAL <= 227 * 1024, "shared-memory layout exceeds the 227 KB opt-in limit");
};

#ifdef CF_TRACE
#define FFN_MARK(slot) trace_mark(slot, (int)(p.flags >> 16))      /* tools/trace_ffn.py passes the launch index in the high flag bits */
#else
#define FFN_MARK(slot) ((void)0)
#endif

__global__ void __launch_bounds__(BLOCK_THREADS, 1)
llama_ffn_layer_kernel(const __grid_constant__ FfnParams p)
{
    using S = SmemFfn;
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t smem_base = dsm::smem_u32(smem);
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int hidden = p.hidden, ffn = p.ffn;
    const int wins = hidden / 256;                          // 256-column windows per row
    const int n_blocks = ffn / FFN_BLOCK;
    const int cta = blockIdx.x, grid = gridDim.x;
    const int nb = (n_blocks - cta + grid - 1) / grid;      // blocks owned by this CTA: cta, cta + grid, ...
    const uint32_t full_u32 = smem_base + S::BARS;

    const uint32_t n_gu_tiles = nb * 2 * wins;              // per block: gate wins, then up wins
    const uint32_t n_dn_tiles = nb * wins;
    const uint32_t total_tiles = n_gu_tiles + n_dn_tiles;

    FFN_MARK(0);
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const uint64_t pol = policy_evict_first();
    auto issue_tile = [&](uint32_t g) {
        if (g >= total_tiles || lane != 0) return;
        const uint32_t s = ring_stage(g);
        const uint32_t fb = full_u32 + 8 * s;
        const uint32_t dst = smem_base + S::RING + s * STAGE_BYTES;
        dsm::mbar_arrive_expect_tx(fb, STAGE_BYTES);
        if (g < n_gu_tiles) {
            const int k = g / (2 * wins), r = g % (2 * wins), mat = r / wins, win = r % wins;
            const int blk = cta + k * grid;
            tma_load_2d(dst, &p.tm_w13, win * 256, mat * ffn + blk * FFN_BLOCK, fb, pol);
        } else {
            const uint32_t i = g - n_gu_tiles;
            const int k = i / wins, cb = i % wins;
            const int blk = cta + k * grid;
            tma_load_2d(dst, &p.tm_w2t, cb * 256, blk * FFN_BLOCK, fb, pol);
        }
    };
    if (lane == 0) {
        dsm::mbar_init(full_u32 + 8 * warp, 1);
        dsm::mbar_init(full_u32 + 8 * (warp + CONSUMER_WARPS), 1);
        if (tid == 0) { prefetch_tmap(&p.tm_w13); prefetch_tmap(&p.tm_w2t); }
        dsm::mbar_fence_init();
    }
    __syncwarp();
    issue_tile(warp);
    issue_tile(warp + CONSUMER_WARPS);

    float* xs = reinterpret_cast<float*>(smem + S::XS);
    float* out_acc = reinterpret_cast<float*>(smem + S::OUT_ACC);
    float* gu = reinterpret_cast<float*>(smem + S::GU);
    float* act = reinterpret_cast<float*>(smem + S::ACT);
    float* red = reinterpret_cast<float*>(smem + S::RED);
    uint32_t* sflags = reinterpret_cast<uint32_t*>(smem + S::FLAGS);

    // the RMSNorm weight does not depend on the previous kernel: fetch it before the dependency wait
    constexpr int P0_ITERS = (FFN_HIDDEN_MAX / 8 + CONSUMER_THREADS - 1) / CONSUMER_THREADS;      // 3
    uint4 wraw[P0_ITERS];
#pragma unroll
    for (int it = 0; it < P0_ITERS; ++it) {
        const int e = (it * CONSUMER_THREADS + tid) * 8;
        wraw[it] = e < hidden ? *reinterpret_cast<const uint4*>(p.rms_w + e) : make_uint4(0, 0, 0, 0);
    }
    for (int e = tid; e < FFN_NB_MAX * 2 * FFN_BLOCK; e += CONSUMER_THREADS) gu[e] = 0.f;

    asm volatile("griddepcontrol.wait;" ::: "memory");

    // ---- phase 0: h = x + residual, RMSNorm (every CTA reduces the whole vector itself).  One pass: x and residual are
    //      loaded once and stay in registers across the block reduction. --------------------------------------------
    {
        uint4 xr[P0_ITERS], rr[P0_ITERS];
#pragma unroll
        for (int it = 0; it < P0_ITERS; ++it) {
            const int e = (it * CONSUMER_THREADS + tid) * 8;
            xr[it] = e < hidden ? *reinterpret_cast<const uint4*>(p.x + e) : make_uint4(0, 0, 0, 0);
            rr[it] = e < hidden ? *reinterpret_cast<const uint4*>(p.residual_in + e) : make_uint4(0, 0, 0, 0);
        }
        float f[P0_ITERS][8];
        float ss = 0.f;
#pragma unroll
        for (int it = 0; it < P0_ITERS; ++it) {
            const int e = (it * CONSUMER_THREADS + tid) * 8;
            float r8[8];
            unpack8(xr[it], f[it]);
            unpack8(rr[it], r8);
            __align__(16) __half hs[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                hs[k] = __float2half_rn(f[it][k] + r8[k]);
                f[it][k] = __half2float(hs[k]);
                ss += f[it][k] * f[it][k];
            }
            if (cta == 0 && e < hidden && p.residual_out != p.residual_in)
                *reinterpret_cast<uint4*>(p.residual_out + e) = *reinterpret_cast<const uint4*>(hs);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        if (lane == 0) red[warp] = ss;
        __syncthreads();
        float tot = 0.f;
#pragma unroll
        for (int w = 0; w < CONSUMER_WARPS; ++w) tot += red[w];
        const float rstd = rsqrtf(tot / (float)hidden + p.eps);
#pragma unroll
        for (int it = 0; it < P0_ITERS; ++it) {
            const int e = (it * CONSUMER_THREADS + tid) * 8;
            if (e < hidden) {
                float w8[8];
                unpack8(wraw[it], w8);
                float4 lo, hi;
                lo.x = round_h(round_h(f[it][0] * rstd) * w8[0]); lo.y = round_h(round_h(f[it][1] * rstd) * w8[1]);
                lo.z = round_h(round_h(f[it][2] * rstd) * w8[2]); lo.w = round_h(round_h(f[it][3] * rstd) * w8[3]);
                hi.x = round_h(round_h(f[it][4] * rstd) * w8[4]); hi.y = round_h(round_h(f[it][5] * rstd) * w8[5]);
                hi.z = round_h(round_h(f[it][6] * rstd) * w8[6]); hi.w = round_h(round_h(f[it][7] * rstd) * w8[7]);
                *reinterpret_cast<float4*>(xs + e) = lo;
                *reinterpret_cast<float4*>(xs + e + 4) = hi;
            }
        }
        __syncthreads();
    }

    FFN_MARK(1);
    // ---- phase 1: gate / up rows of this CTA's blocks ------------------------------------------------------
    for (uint32_t g = warp; g < n_gu_tiles; g += CONSUMER_WARPS) {
        const uint32_t s = ring_stage(g);
        const int k = g / (2 * wins), r = g % (2 * wins), mat = r / wins, win = r % wins;
        float x8[8];
        {
            const float4 a = *reinterpret_cast<const float4*>(xs + win * 256 + lane * 8);
            const float4 b = *reinterpret_cast<const float4*>(xs + win * 256 + lane * 8 + 4);
            x8[0] = a.x; x8[1] = a.y; x8[2] = a.z; x8[3] = a.w; x8[4] = b.x; x8[5] = b.y; x8[6] = b.z; x8[7] = b.w;
        }
        ring_wait_full(full_u32, g);
        const uint4* tile = reinterpret_cast<const uint4*>(smem + S::RING + s * STAGE_BYTES);
        float* dst = gu + (k * 2 + mat) * FFN_BLOCK;
#pragma unroll
        for (int grp = 0; grp < FFN_BLOCK / 8; ++grp) {
            float v[8];
#pragma unroll
            for (int rr = 0; rr < 8; ++rr) {
                float w8[8];
                unpack8(tile[(grp * 8 + rr) * 32 + lane], w8);
                float a = 0.f;
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) a = fmaf(x8[kk], w8[kk], a);
                v[rr] = a;
            }
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
                const bool hi = lane & 16;
                const float send = hi ? v[rr] : v[rr + 4];
                const float keep = hi ? v[rr + 4] : v[rr];
                v[rr] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
            }
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                const bool hi = lane & 8;
                const float send = hi ? v[rr] : v[rr + 2];
                const float keep = hi ? v[rr + 2] : v[rr];
                v[rr] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
            }
            {
                const bool hi = lane & 4;
                const float send = hi ? v[0] : v[1];
                const float keep = hi ? v[1] : v[0];
                v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
            }
            v[0] += __shfl_xor_sync(0xffffffffu, v[0], 2);
            v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
            if ((lane & 3) == 0) {
                const int row = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
                atomicAdd(dst + grp * 8 + row, v[0]);          // 16 windows of the same row land here
            }
        }
        __syncwarp();
        issue_tile(g + NSTAGES);
    }
    __syncthreads();
    FFN_MARK(2);

    // ---- SwiGLU on this CTA's nb * 16 intermediate values (fp16 rounding points of the eager model) ---------
    for (int e = tid; e < nb * FFN_BLOCK; e += CONSUMER_THREADS) {
        const int k = e / FFN_BLOCK, j = e % FFN_BLOCK;
        const float gv = round_h(gu[(k * 2 + 0) * FFN_BLOCK + j]);
        const float uv = round_h(gu[(k * 2 + 1) * FFN_BLOCK + j]);
        const float sg = round_h(gv / (1.f + __expf(-gv)));
        act[e] = round_h(sg * uv);
    }
    for (int e = tid; e < hidden; e += CONSUMER_THREADS) out_acc[e] = 0.f;
    __syncthreads();

    FFN_MARK(4);
    // ---- phase 2: down projection, rows of W2^T for this CTA's blocks ---------------------------------------
    for (uint32_t i = first_tile(n_gu_tiles, warp); i < n_dn_tiles; i += CONSUMER_WARPS) {
        const uint32_t g = n_gu_tiles + i, s = ring_stage(g);
        const int k = i / wins, cb = i % wins;
        ring_wait_full(full_u32, g);
        const uint4* tile = reinterpret_cast<const uint4*>(smem + S::RING + s * STAGE_BYTES);
        const float* a16 = act + k * FFN_BLOCK;
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int r = 0; r < FFN_BLOCK; ++r) {
            float w8[8];
            unpack8(tile[r * 32 + lane], w8);
            const float av = a16[r];
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) acc[kk] = fmaf(av, w8[kk], acc[kk]);
        }
        __syncwarp();
        issue_tile(g + NSTAGES);
        float* dst = out_acc + cb * 256 + lane * 8;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) atomicAdd(dst + kk, acc[kk]);
    }
    __syncthreads();
    FFN_MARK(5);

    // ---- cross-CTA reduction: fp32 red into scratch, the last CTA finalises ------------------------------------
    for (int e = tid * 4; e < hidden; e += CONSUMER_THREADS * 4)
        red_add_v4(p.scratch + e, *reinterpret_cast<const float4*>(out_acc + e));
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned prev = atomicAdd(&p.counters[0], 1u);
        sflags[0] = (prev == (unsigned)grid - 1u);
    }
    __syncthreads();
    FFN_MARK(8);
    if (sflags[0]) {
        __threadfence();
        const bool fp32_out = p.flags & 1u;
        // the last CTA is the critical path of the next kernel: all scratch loads go out first (one L2 round trip, not three)
        constexpr int FIN_ITERS = (FFN_HIDDEN_MAX / 4 + CONSUMER_THREADS - 1) / CONSUMER_THREADS;      // 6
        float4 v[FIN_ITERS];
#pragma unroll
        for (int it = 0; it < FIN_ITERS; ++it) {
            const int e = (it * CONSUMER_THREADS + tid) * 4;
            if (e < hidden) v[it] = ld_cg_v4(p.scratch + e);
        }
#pragma unroll
        for (int it = 0; it < FIN_ITERS; ++it) {
            const int e = (it * CONSUMER_THREADS + tid) * 4;
            if (e < hidden) {
                *reinterpret_cast<float4*>(p.scratch + e) = make_float4(0.f, 0.f, 0.f, 0.f);
                if (fp32_out) {
                    *reinterpret_cast<float4*>(static_cast<float*>(p.out) + e) = v[it];
                } else {
                    __align__(8) __half h4[4] = {__float2half_rn(v[it].x), __float2half_rn(v[it].y),
                                                 __float2half_rn(v[it].z), __float2half_rn(v[it].w)};
                    *reinterpret_cast<uint2*>(static_cast<__half*>(p.out) + e) = *reinterpret_cast<const uint2*>(h4);
                }
            }
        }
        if (tid == 0) p.counters[0] = 0u;
        if (p.residual_out == p.residual_in) {
            // in-place residual stream: every other CTA has finished reading it by now
            for (int e = tid * 8; e < hidden; e += CONSUMER_THREADS * 8) {
                float f[8], r8[8];
                unpack8(*reinterpret_cast<const uint4*>(p.x + e), f);
                unpack8(*reinterpret_cast<const uint4*>(p.residual_in + e), r8);
                __align__(16) __half hs[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) hs[k] = __float2half_rn(f[k] + r8[k]);
                *reinterpret_cast<uint4*>(p.residual_out + e) = *reinterpret_cast<const uint4*>(hs);
            }
        }
    }
    FFN_MARK(9);
}

}  // namespace cfb
