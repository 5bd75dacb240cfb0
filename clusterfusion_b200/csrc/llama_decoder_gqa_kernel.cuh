/*
 * FIRST-GENERATION grouped-query kernel (selected only with CF_FLAG_GQA_CLUSTER; the default is the group kernel in
 * llama_decoder_gqa2_kernel.cuh, which is 1.3-2x faster on B200 -- kept for A/B measurement and because it shows the
 * pure-DSMEM formulation).
 * Grouped-query variant of the fused decoder attention half-layer (Llama-3-8B: 32 Q / 8 KV heads; the
 * Llama-2-70B head-parallel shards: 8 Q heads per KV head).  New capability relative to the reference, whose
 * kernels hard-code MHA (KV row stride = HIDDEN_DIM, /root/reference/include/H100/llama/llama_kernel_dispatch.cu:63-64;
 * SURVEY.md section 8 row a7).  nn.Linear ("sglang") weight layout only -- that is how GQA checkpoints are stored.
 *
 * Mapping: one 8- or 16-CTA cluster per (request, KV head, group of NQ = 4 query heads) -- 16 when at most four
 * clusters exist (7 16-CTA / 15 8-CTA clusters are co-resident at this footprint, profiles/r02_cluster_probe.txt), else 8.
 * The 4 query heads share
 * every K/V tile, so K/V are read from HBM once per KV head (SURVEY.md section 8d bytes model); a KV head with 8
 * query heads (70B) gets two clusters, which recompute the small K/V projection and read the cache twice (the
 * second read hits L2).  Inside the cluster: 16-way K-split of the QKV GEMV, 16-way sequence split of the cache,
 * 16-way N-split of the O GEMV -- the same decomposition as the MHA kernel, same single TMA ring / producer warp
 * / warp-private tiles (llama_decoder_kernel.cuh).
 *
 * What changes with 16 CTAs is the collective: an all-to-all push of 3 KB vectors would need 48 KB of receive
 * slots per CTA, so both exchanges are reduce-scatter (cluster_scatter: CTA r receives everyone's slice r and
 * folds it in rank order) followed by all-gather (cluster_reduce<.., QUK_DEEPSEEK>), both in include/dsm.cuh.
 */
#pragma once

#include "llama_decoder_kernel.cuh"

namespace cfb {

constexpr int GQA_KS_MAX = 1024;       // hidden / CLUSTER <= 1024  (hidden 8192 with 8-CTA clusters)

template <int CLUSTER, int NQ>
struct SmemGqa {
    static constexpr int R = (NQ + 2) * HEAD_DIM;                 // q(NQ heads) | k | v rows of one cluster
    static constexpr int SLICE1 = R / CLUSTER;                    // floats per CTA in exchange 1 (48)
    static constexpr int SLICE2 = NQ * HEAD_DIM / CLUSTER;        // attention-output floats per CTA (32)
    static constexpr int PAY2 = SLICE2 + 4;                       // [m, l, -, -, o[32]]
    static constexpr int RING = 0;
    static constexpr int UNION = RING + NSTAGES * STAGE_BYTES;
    //   phase QKV : xs fp32[GQA_KS_MAX] | qkv_part fp32[GQA_KS_MAX/256][R]
    //   phase ATTN: attn_part fp32[12][132] | cta_state fp32[NQ][132]
    //   phase O   : out_part fp32[NQ*128/256][GQA_KS_MAX]
    static constexpr int UNION_BYTES = 12 * (HEAD_DIM + 4) * 4 + NQ * (HEAD_DIM + 4) * 4;      // 8448
    static constexpr int XS = UNION;
    static constexpr int QKV_PART = UNION + GQA_KS_MAX * 4;
    static constexpr int QKV_BYTES = GQA_KS_MAX * 4 + (GQA_KS_MAX / 256) * R * 4;
    static constexpr int UNION_SIZE = QKV_BYTES > UNION_BYTES ? QKV_BYTES : UNION_BYTES;
    static constexpr int ATTN_PART = UNION;
    static constexpr int CTA_STATE = UNION + 12 * (HEAD_DIM + 4) * 4;
    static constexpr int OUT_PART = UNION;
    static constexpr int QKV_SRC = UNION + UNION_SIZE;                     // fp32[R] this CTA's partial sums
    static constexpr int RS1 = QKV_SRC + R * 4;                            // fp32[CLUSTER][SLICE1]
    static constexpr int RED1 = RS1 + CLUSTER * SLICE1 * 4;                // fp32[SLICE1]
    static constexpr int AG1 = RED1 + SLICE1 * 4;                          // fp32[CLUSTER][SLICE1] = full q|k|v
    static constexpr int QKV_FIN = AG1 + CLUSTER * SLICE1 * 4;             // fp32[R] roped q*scale | k | v
    // exchange-2 buffers reuse exchange-1 buffers that are dead by then.  Safe against early peers: a peer can only
    // scatter into RS2 after it finished the exchange-1 all-gather, which needed this CTA's gather contribution,
    // which this CTA sends after it has folded RS1.
    static constexpr int SEND2 = QKV_SRC;                                  // fp32[CLUSTER][PAY2]
    static constexpr int RS2 = RS1;                                        // fp32[CLUSTER][PAY2]
    static constexpr int RED2 = RED1;                                      // fp32[SLICE2]
    static_assert(CLUSTER * PAY2 <= R && CLUSTER * PAY2 <= CLUSTER * SLICE1 && SLICE2 <= SLICE1, "exchange-2 buffers must fit");
    static constexpr int AG2 = QKV_FIN + R * 4;                            // fp32[CLUSTER][SLICE2] = attention output
    static constexpr int RED = AG2 + CLUSTER * SLICE2 * 4;                 // fp32[32]
    static constexpr int BARS = RED + 32 * 4;                              // full[24], xbar[4]
    static constexpr int FLAGS = BARS + (NSTAGES + 4) * 8;
    static constexpr int TOTAL = FLAGS + 16;
    static_assert(TOTAL <= 227 * 1024, "shared-memory layout exceeds the 227 KB opt-in limit");
};

// 16 output rows x 256 input columns of an [out,in] weight tile against 8 activations per lane;
// writes the 16 row sums to out[0..16).  Same inner loop as the MHA kernel's sglang QKV phase.
__device__ __forceinline__ void gemv_tile_16x256(const uint4* tile, const float (&x8)[8], float* out, uint32_t lane) {
#pragma unroll
    for (int grp = 0; grp < ROWS512 / 8; ++grp) {
        float v[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            float w8[8];
            unpack8(tile[(grp * 8 + r) * 32 + lane], w8);
            float a = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) a = fmaf(x8[k], w8[k], a);
            v[r] = a;
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const bool hi = lane & 16;
            const float send = hi ? v[r] : v[r + 4];
            const float keep = hi ? v[r + 4] : v[r];
            v[r] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const bool hi = lane & 8;
            const float send = hi ? v[r] : v[r + 2];
            const float keep = hi ? v[r + 2] : v[r];
            v[r] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
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
            const int r = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
            out[grp * 8 + r] = v[0];
        }
    }
}

template <int VARIANT, int CLUSTER, int NQ>
__global__ void __launch_bounds__(BLOCK_THREADS, 1)
llama_decoder_layer_gqa_kernel(const __grid_constant__ KParams p)
{
    using S = SmemGqa<CLUSTER, NQ>;
    static_assert(VARIANT != CHAT, "GQA uses the nn.Linear weight layout");
    static_assert(S::R % CLUSTER == 0 && (S::SLICE1 * 4) % 16 == 0 && (S::PAY2 * 4) % 16 == 0, "slice alignment");
    constexpr bool kPaged = (VARIANT == PAGED);

    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t smem_base = dsm::smem_u32(smem);
    const uint32_t tid = threadIdx.x;
    const uint32_t warp = tid >> 5;
    const uint32_t lane = tid & 31;
    const uint32_t rank = dsm::cluster_ctarank();
    const uint32_t cid = blockIdx.x / CLUSTER;
    const uint32_t batch = blockIdx.y;

    const int hidden = p.hidden;
    const int Hq = p.n_heads, Hkv = p.n_kv_heads;
    const int qsplit = (Hq / Hkv) / NQ;                 // clusters per KV head
    const int kvh = cid / qsplit;
    const int qh0 = kvh * (Hq / Hkv) + (cid % qsplit) * NQ;     // first query head of this cluster
    const bool writes_kv = (cid % qsplit) == 0;
    const int KS = hidden / CLUSTER;
    const int kv_cols = Hkv * HEAD_DIM;

    const uint32_t full_u32 = smem_base + S::BARS;          // u64 full[NSTAGES]
    const uint32_t xbar_u32 = full_u32 + NSTAGES * 8;        // u64 xbar[4]

    int kv_len, kv_base = 0, new_slot = 0;
    if constexpr (kPaged) {
        kv_base = p.indptr[batch];
        const int end = p.indptr[batch + 1] - 1;
        kv_len = end - kv_base;
        new_slot = p.indices[end];
    } else {
        kv_len = p.kv_len;
    }
    const int chunk = (((kv_len + CLUSTER - 1) / CLUSTER) + ROWS512 - 1) & ~(ROWS512 - 1);
    const int row_begin = min((int)rank * chunk, kv_len);
    const int row_end = min(row_begin + chunk, kv_len);
    const int wins = KS / 256;
    const uint32_t n_qkv_tiles = (S::R / ROWS512) * wins;
    const uint32_t n_kv_tiles = (row_end - row_begin + ROWS512 - 1) / ROWS512;
    const uint32_t n_o_tiles = (KS / ROWS512) * (NQ * HEAD_DIM / 256);

    CF_MARK(0);
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    // ---- tile stream: every warp requests, consumes and re-requests its own tiles (see llama_decoder_kernel.cuh) ----
    const uint64_t pol = policy_evict_first();
    const uint32_t total_tiles = n_qkv_tiles + n_kv_tiles + n_o_tiles;
    const __half* kpool = nullptr;
    const __half* vpool = nullptr;
    if constexpr (kPaged) {
        kpool = reinterpret_cast<const __half*>(p.k_pool_ptrs[p.layer_id]);
        vpool = reinterpret_cast<const __half*>(p.v_pool_ptrs[p.layer_id]);
    }
    auto issue_tile = [&](uint32_t g) {
        if (g >= total_tiles) return;
        const uint32_t s = ring_stage(g);
        const uint32_t fb = full_u32 + 8 * s;
        const uint32_t dst = smem_base + S::RING + s * STAGE_BYTES;
        if (g < n_qkv_tiles) {
            if (lane == 0) {
                constexpr int BPH = HEAD_DIM / ROWS512;          // 16-row blocks per head (8)
                const int rb = g / wins, win = g % wins;         // rb: 16-row block inside q(NQ*128) | k(128) | v(128)
                int row0;
                if (rb < NQ * BPH) row0 = qh0 * HEAD_DIM + rb * ROWS512;
                else if (rb < NQ * BPH + BPH) row0 = Hq * HEAD_DIM + kvh * HEAD_DIM + (rb - NQ * BPH) * ROWS512;
                else row0 = (Hq + Hkv) * HEAD_DIM + kvh * HEAD_DIM + (rb - NQ * BPH - BPH) * ROWS512;
                dsm::mbar_arrive_expect_tx(fb, STAGE_BYTES);
                tma_load_2d(dst, &p.tm_wqkv, rank * KS + win * 256, row0, fb, pol);
            }
        } else if (g < n_qkv_tiles + n_kv_tiles) {
            const uint32_t i = g - n_qkv_tiles;
            if constexpr (!kPaged) {
                if (lane == 0) {
                    const int r0 = row_begin + i * ROWS512;
                    dsm::mbar_arrive_expect_tx(fb, STAGE_BYTES);
                    tma_load_2d(dst, &p.tm_k, kvh * HEAD_DIM, r0, fb, pol);
                    tma_load_2d(dst + STAGE_BYTES / 2, &p.tm_v, kvh * HEAD_DIM, r0, fb, pol);
                }
            } else {
                const int r = row_begin + i * ROWS512 + (lane & 15);
                const bool valid = r < row_end;
                const long long slot = valid ? (long long)p.indices[kv_base + r] : 0;
                const int nvalid = min(ROWS512, row_end - (row_begin + (int)i * ROWS512));
                if (lane == 0) dsm::mbar_arrive_expect_tx(fb, nvalid * 2 * HEAD_DIM * 2);
                __syncwarp();
                if (valid) {
                    const uint32_t d = dst + (lane & 15) * (HEAD_DIM * 2);
                    if (lane < 16) bulk_load_1d(d, kpool + slot * kv_cols + kvh * HEAD_DIM, HEAD_DIM * 2, fb, pol);
                    else bulk_load_1d(d + STAGE_BYTES / 2, vpool + slot * kv_cols + kvh * HEAD_DIM, HEAD_DIM * 2, fb, pol);
                }
            }
        } else {
            if (lane == 0) {
                constexpr int owins = NQ * HEAD_DIM / 256;
                const uint32_t i = g - n_qkv_tiles - n_kv_tiles;
                const int rb = i / owins, win = i % owins;       // Wo [out][in]: 16 output rows x 256 of this cluster's input cols
                dsm::mbar_arrive_expect_tx(fb, STAGE_BYTES);
                tma_load_2d(dst, &p.tm_wo, qh0 * HEAD_DIM + win * 256, rank * KS + rb * ROWS512, fb, pol);
            }
        }
    };

    if (lane == 0) {
        dsm::mbar_init(full_u32 + 8 * warp, 1);
        dsm::mbar_init(full_u32 + 8 * (warp + CONSUMER_WARPS), 1);
        if (tid == 0) {
            prefetch_tmap(&p.tm_wqkv);
            prefetch_tmap(&p.tm_wo);
            if constexpr (!kPaged) { prefetch_tmap(&p.tm_k); prefetch_tmap(&p.tm_v); }
            cluster_reduce_arm<CLUSTER>(xbar_u32, S::SLICE1 * 4);
            cluster_reduce_arm<CLUSTER>(xbar_u32 + 8, S::SLICE1 * 4);
            cluster_reduce_arm<CLUSTER>(xbar_u32 + 16, S::PAY2 * 4);
            cluster_reduce_arm<CLUSTER>(xbar_u32 + 24, S::SLICE2 * 4);
        }
        dsm::mbar_fence_init();
    }
    __syncwarp();
    CF_MARK(12);
    issue_tile(warp);
    issue_tile(warp + CONSUMER_WARPS);
    dsm::cluster_arrive();

    // =============================================================================================
    // ALL WARPS
    // =============================================================================================
    float* xs = reinterpret_cast<float*>(smem + S::XS);
    float* qkv_part = reinterpret_cast<float*>(smem + S::QKV_PART);
    float* attn_part = reinterpret_cast<float*>(smem + S::ATTN_PART);
    float* cta_state = reinterpret_cast<float*>(smem + S::CTA_STATE);
    float* out_part = reinterpret_cast<float*>(smem + S::OUT_PART);
    float* qkv_src = reinterpret_cast<float*>(smem + S::QKV_SRC);
    float* rs1 = reinterpret_cast<float*>(smem + S::RS1);
    float* red1 = reinterpret_cast<float*>(smem + S::RED1);
    float* ag1 = reinterpret_cast<float*>(smem + S::AG1);
    float* qkv_fin = reinterpret_cast<float*>(smem + S::QKV_FIN);
    float* send2 = reinterpret_cast<float*>(smem + S::SEND2);
    float* rs2 = reinterpret_cast<float*>(smem + S::RS2);
    float* red2 = reinterpret_cast<float*>(smem + S::RED2);
    float* ag2 = reinterpret_cast<float*>(smem + S::AG2);
    float* red = reinterpret_cast<float*>(smem + S::RED);
    uint32_t* sflags = reinterpret_cast<uint32_t*>(smem + S::FLAGS);

    const __half* xg = p.x + (size_t)batch * hidden;
    const __half* rg = p.residual_in + (size_t)batch * hidden;
    __half* rout = p.residual_out + (size_t)batch * hidden;
    const bool residual_inplace = (static_cast<const void*>(rout) == static_cast<const void*>(rg));

    asm volatile("griddepcontrol.wait;" ::: "memory");

    // ---- phase 0: fused residual add + RMSNorm ------------------------------------------------------
    {
        float ss = 0.f;
        for (int e = tid * 8; e < hidden; e += CONSUMER_THREADS * 8) {
            float f[8], r8[8];
            unpack8(*reinterpret_cast<const uint4*>(xg + e), f);
            unpack8(*reinterpret_cast<const uint4*>(rg + e), r8);
#pragma unroll
            for (int k = 0; k < 8; ++k) { f[k] = round_h(f[k] + r8[k]); ss += f[k] * f[k]; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        if (lane == 0) red[warp] = ss;
        dsm::named_bar_sync(CONSUMER_BAR, CONSUMER_THREADS);
        float tot = 0.f;
#pragma unroll
        for (int w = 0; w < CONSUMER_WARPS; ++w) tot += red[w];
        const float rstd = rsqrtf(tot / (float)hidden + p.eps);
        for (int e = tid * 8; e < KS; e += CONSUMER_THREADS * 8) {
            const int ge = rank * KS + e;
            float f[8], w8[8], r8[8];
            unpack8(*reinterpret_cast<const uint4*>(xg + ge), f);
            unpack8(*reinterpret_cast<const uint4*>(p.rms_w + ge), w8);
            unpack8(*reinterpret_cast<const uint4*>(rg + ge), r8);
            __align__(16) __half hs[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) { hs[k] = __float2half_rn(f[k] + r8[k]); f[k] = __half2float(hs[k]); }
            if (cid == 0 && !residual_inplace)
                *reinterpret_cast<uint4*>(rout + ge) = *reinterpret_cast<const uint4*>(hs);
#pragma unroll
            for (int k = 0; k < 8; ++k) xs[e + k] = round_h(round_h(f[k] * rstd) * w8[k]);
        }
        dsm::named_bar_sync(CONSUMER_BAR, CONSUMER_THREADS);
    }
    CF_MARK(1);

    uint32_t gbase = 0;
    // ---- phase 1: QKV GEMV over this CTA's K-slice --------------------------------------------------
    for (uint32_t i = first_tile(gbase, warp); i < n_qkv_tiles; i += CONSUMER_WARPS) {
        const uint32_t g = gbase + i, s = ring_stage(g);
        const int rb = i / wins, win = i % wins;
        float x8[8];
        {
            const float4 a = *reinterpret_cast<const float4*>(xs + win * 256 + lane * 8);
            const float4 b = *reinterpret_cast<const float4*>(xs + win * 256 + lane * 8 + 4);
            x8[0] = a.x; x8[1] = a.y; x8[2] = a.z; x8[3] = a.w; x8[4] = b.x; x8[5] = b.y; x8[6] = b.z; x8[7] = b.w;
        }
        ring_wait_full(full_u32, g);
        gemv_tile_16x256(reinterpret_cast<const uint4*>(smem + S::RING + s * STAGE_BYTES), x8,
                         qkv_part + win * S::R + rb * ROWS512, lane);
        __syncwarp();
        issue_tile(g + NSTAGES);
    }
    gbase += n_qkv_tiles;
    dsm::named_bar_sync(CONSUMER_BAR, CONSUMER_THREADS);
    CF_MARK(2);
    for (int o = tid; o < S::R; o += CONSUMER_THREADS) {
        float a = 0.f;
        for (int w = 0; w < wins; ++w) a += qkv_part[w * S::R + o];
        qkv_src[o] = a;
    }

    // ---- exchange 1: reduce-scatter (sum) + all-gather of q|k|v ---------------------------------------
    dsm::cluster_wait();
    uint32_t ph0 = 0, ph1 = 0, ph2 = 0, ph3 = 0;
    cluster_scatter<CLUSTER, CONSUMER_THREADS, CONSUMER_BAR>(S::SLICE1 * 4, tid, rank, smem_base + S::RS1, xbar_u32, ph0,
                                                             qkv_src, rs1);
    if (tid < S::SLICE1) {
        float a = 0.f;
#pragma unroll
        for (int r = 0; r < CLUSTER; ++r) a += rs1[r * S::SLICE1 + tid];
        red1[tid] = round_h(a);                                  // q / k / v leave the projection as fp16 (eager model)
    }
    cluster_reduce<CLUSTER, Stage::QUK_DEEPSEEK, CONSUMER_THREADS, CONSUMER_BAR>(
        S::SLICE1 * 4, tid, S::SLICE1, rank, smem_base + S::RED1, smem_base + S::AG1, xbar_u32 + 8, ph1, red1, ag1);
    CF_MARK(3);

    // ---- RoPE (NeoX), new K/V out ---------------------------------------------------------------------
    {
        constexpr float kScaleLog2 = 0.08838834764831845f * 1.4426950408889634f;
        const float* cosp = p.cos;
        const float* sinp = p.sin;
        if constexpr (kPaged) {
            cosp = p.cos + p.positions[batch] * HEAD_DIM;
            sinp = cosp + HEAD_DIM / 2;
        }
        for (int e = tid; e < S::R; e += CONSUMER_THREADS) {
            const int hd = e >> 7, d = e & 127;                  // hd < NQ: query head; NQ: k; NQ+1: v
            const float a = ag1[e];
            if (hd <= NQ) {
                const float b = ag1[e ^ 64];
                const int i = d & 63;
                const float rot = (d & 64) ? fmaf(a, cosp[i], b * sinp[i]) : fmaf(a, cosp[i], -b * sinp[i]);
                const __half rh = __float2half_rn(rot);
                qkv_fin[e] = hd < NQ ? __half2float(rh) * kScaleLog2 : __half2float(rh);
                if (hd == NQ && rank == 0 && writes_kv) {
                    if constexpr (kPaged) {
                        __half* kpool = reinterpret_cast<__half*>(p.k_pool_ptrs[p.layer_id]);
                        kpool[(size_t)new_slot * kv_cols + kvh * HEAD_DIM + d] = rh;
                    } else {
                        p.k_new[kvh * HEAD_DIM + d] = rh;
                    }
                }
            } else {
                qkv_fin[e] = a;
                if (rank == 0 && writes_kv) {
                    const __half vh = __float2half_rn(a);
                    if constexpr (kPaged) {
                        __half* vpool = reinterpret_cast<__half*>(p.v_pool_ptrs[p.layer_id]);
                        vpool[(size_t)new_slot * kv_cols + kvh * HEAD_DIM + d] = vh;
                    } else {
                        p.v_new[kvh * HEAD_DIM + d] = vh;
                    }
                }
            }
        }
        dsm::named_bar_sync(CONSUMER_BAR, CONSUMER_THREADS);
    }
    CF_MARK(4);

    // ---- phase 2: flash-decode, NQ query heads share each K/V tile -------------------------------------
    {
        const int sub = lane >> 4, c = lane & 15;
        float q8[NQ][8], o8[NQ][8], m[NQ], l[NQ];
#pragma unroll
        for (int h = 0; h < NQ; ++h) {
            m[h] = -INFINITY; l[h] = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) { q8[h][k] = qkv_fin[h * HEAD_DIM + c * 8 + k]; o8[h][k] = 0.f; }
        }
        for (uint32_t i = first_tile(gbase, warp); i < n_kv_tiles; i += CONSUMER_WARPS) {
            const uint32_t g = gbase + i, s = ring_stage(g);
            ring_wait_full(full_u32, g);
            const uint4* kt = reinterpret_cast<const uint4*>(smem + S::RING + s * STAGE_BYTES);
            const uint4* vt = kt + STAGE_BYTES / 32;
            const int rows_left = row_end - (row_begin + (int)i * ROWS512);
            {
                constexpr int half = 0;
                float sc[NQ][8];
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                    const int row = 2 * (half * 8 + jj) + sub;
                    float k8[8];
                    unpack8(kt[row * 16 + c], k8);
#pragma unroll
                    for (int h = 0; h < NQ; ++h) {
                        float a = 0.f;
#pragma unroll
                        for (int k = 0; k < 8; ++k) a = fmaf(q8[h][k], k8[k], a);
                        a += __shfl_xor_sync(0xffffffffu, a, 1);
                        a += __shfl_xor_sync(0xffffffffu, a, 2);
                        a += __shfl_xor_sync(0xffffffffu, a, 4);
                        a += __shfl_xor_sync(0xffffffffu, a, 8);
                        sc[h][jj] = (row < rows_left) ? a : -INFINITY;
                    }
                }
                float mu[NQ];
#pragma unroll
                for (int h = 0; h < NQ; ++h) {
                    float mx = sc[h][0];
#pragma unroll
                    for (int jj = 1; jj < 8; ++jj) mx = fmaxf(mx, sc[h][jj]);
                    const float m_new = fmaxf(m[h], mx);
                    mu[h] = (m_new == -INFINITY) ? 0.f : m_new;
                    const float corr = dsm::exp2_diff(m[h], mu[h]);
                    l[h] *= corr;
#pragma unroll
                    for (int k = 0; k < 8; ++k) o8[h][k] *= corr;
                    m[h] = m_new;
                }
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                    const int row = 2 * (half * 8 + jj) + sub;
                    uint4 raw = vt[row * 16 + c];
                    if constexpr (kPaged) {
                        if (row >= rows_left) raw = make_uint4(0, 0, 0, 0);
                    }
                    float v8[8];
                    unpack8(raw, v8);
#pragma unroll
                    for (int h = 0; h < NQ; ++h) {
                        const float pr = dsm::fast_exp2(sc[h][jj] - mu[h]);
                        l[h] += pr;
#pragma unroll
                        for (int k = 0; k < 8; ++k) o8[h][k] = fmaf(pr, v8[k], o8[h][k]);
                    }
                }
            }
            __syncwarp();
            issue_tile(g + NSTAGES);
        }
        gbase += n_kv_tiles;
        CF_MARK(5);
        // merge the two half-warps (they saw different rows) in registers
#pragma unroll
        for (int h = 0; h < NQ; ++h) {
            const float m2 = __shfl_xor_sync(0xffffffffu, m[h], 16);
            const float l2 = __shfl_xor_sync(0xffffffffu, l[h], 16);
            const float M = fmaxf(m[h], m2);
            const float w1 = dsm::exp2_diff(m[h], M), w2 = dsm::exp2_diff(m2, M);
            l[h] = l[h] * w1 + l2 * w2;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float o2 = __shfl_xor_sync(0xffffffffu, o8[h][k], 16);
                o8[h][k] = o8[h][k] * w1 + o2 * w2;
            }
            m[h] = M;
        }
        // block merge, one head at a time through a 12 x 132 float buffer; rank 0 folds in the current token
#pragma unroll
        for (int h = 0; h < NQ; ++h) {
            if (sub == 0) {
                float* slot = attn_part + warp * (HEAD_DIM + 4);
                if (c == 0) { slot[0] = m[h]; slot[1] = l[h]; }
                *reinterpret_cast<float4*>(slot + 4 + c * 8) = make_float4(o8[h][0], o8[h][1], o8[h][2], o8[h][3]);
                *reinterpret_cast<float4*>(slot + 4 + c * 8 + 4) = make_float4(o8[h][4], o8[h][5], o8[h][6], o8[h][7]);
            }
            if (warp == 0) {
                float a = 0.f;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    a = fmaf(qkv_fin[h * HEAD_DIM + lane * 4 + k], qkv_fin[NQ * HEAD_DIM + lane * 4 + k], a);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
                if (lane == 0) red[CONSUMER_WARPS] = a;
            }
            dsm::named_bar_sync(CONSUMER_BAR, CONSUMER_THREADS);
            if (tid < HEAD_DIM) {
                const bool with_new = (rank == 0);
                float M = with_new ? red[CONSUMER_WARPS] : -INFINITY;
#pragma unroll
                for (int gI = 0; gI < CONSUMER_WARPS; ++gI) M = fmaxf(M, attn_part[gI * (HEAD_DIM + 4)]);
                float L = 0.f, O = 0.f;
#pragma unroll
                for (int gI = 0; gI < CONSUMER_WARPS; ++gI) {
                    const float w = dsm::exp2_diff(attn_part[gI * (HEAD_DIM + 4)], M);
                    L = fmaf(attn_part[gI * (HEAD_DIM + 4) + 1], w, L);
                    O = fmaf(attn_part[gI * (HEAD_DIM + 4) + 4 + tid], w, O);
                }
                if (with_new) {
                    const float w = dsm::exp2_diff(red[CONSUMER_WARPS], M);
                    L += w;
                    O = fmaf(qkv_fin[(NQ + 1) * HEAD_DIM + tid], w, O);
                }
                float* st = cta_state + h * (HEAD_DIM + 4);
                st[4 + tid] = O;
                if (tid == 0) { st[0] = M; st[1] = L; }
            }
            dsm::named_bar_sync(CONSUMER_BAR, CONSUMER_THREADS);
        }
        // ---- exchange 2: reduce-scatter of the softmax states (owner r: head r / (CLUSTER/NQ), 32 dims) ------
        for (int e = tid; e < CLUSTER * S::PAY2; e += CONSUMER_THREADS) {
            const int owner = e / S::PAY2, f = e % S::PAY2;
            const int h = owner / (CLUSTER / NQ), d0 = (owner % (CLUSTER / NQ)) * S::SLICE2;
            const float* st = cta_state + h * (HEAD_DIM + 4);
            send2[e] = f < 2 ? st[f] : (f < 4 ? 0.f : st[4 + d0 + (f - 4)]);
        }
        cluster_scatter<CLUSTER, CONSUMER_THREADS, CONSUMER_BAR>(S::PAY2 * 4, tid, rank, smem_base + S::RS2, xbar_u32 + 16,
                                                                 ph2, send2, rs2);
        if (tid < S::SLICE2) {
            float M = -INFINITY;
#pragma unroll
            for (int r = 0; r < CLUSTER; ++r) M = fmaxf(M, rs2[r * S::PAY2]);
            float L = 0.f, O = 0.f;
#pragma unroll
            for (int r = 0; r < CLUSTER; ++r) {
                const float w = dsm::exp2_diff(rs2[r * S::PAY2], M);
                L = fmaf(rs2[r * S::PAY2 + 1], w, L);
                O = fmaf(rs2[r * S::PAY2 + 4 + tid], w, O);
            }
            red2[tid] = round_h(O / L);                          // attention output leaves as fp16 (eager model)
        }
        cluster_reduce<CLUSTER, Stage::QUK_DEEPSEEK, CONSUMER_THREADS, CONSUMER_BAR>(
            S::SLICE2 * 4, tid, S::SLICE2, rank, smem_base + S::RED2, smem_base + S::AG2, xbar_u32 + 24, ph3, red2, ag2);
    }
    CF_MARK(6);

    // ---- phase 3: O GEMV for output rows [rank*KS, +KS) over this cluster's NQ*128 input columns --------
    {
        constexpr int owins = NQ * HEAD_DIM / 256;
        for (uint32_t i = first_tile(gbase, warp); i < n_o_tiles; i += CONSUMER_WARPS) {
            const uint32_t g = gbase + i, s = ring_stage(g);
            const int rb = i / owins, win = i % owins;
            float a8[8];
            {
                const float4 a = *reinterpret_cast<const float4*>(ag2 + win * 256 + lane * 8);
                const float4 b = *reinterpret_cast<const float4*>(ag2 + win * 256 + lane * 8 + 4);
                a8[0] = a.x; a8[1] = a.y; a8[2] = a.z; a8[3] = a.w; a8[4] = b.x; a8[5] = b.y; a8[6] = b.z; a8[7] = b.w;
            }
            ring_wait_full(full_u32, g);
            gemv_tile_16x256(reinterpret_cast<const uint4*>(smem + S::RING + s * STAGE_BYTES), a8,
                             out_part + win * GQA_KS_MAX + rb * ROWS512, lane);
            __syncwarp();
            issue_tile(g + NSTAGES);
        }
        dsm::named_bar_sync(CONSUMER_BAR, CONSUMER_THREADS);
    }
    CF_MARK(7);

    // ---- cross-cluster reduction: fp32 red into scratch, last arriver of the slice finalises ----------------
    constexpr int owins = NQ * HEAD_DIM / 256;
    float* scratch = p.scratch + (size_t)batch * hidden + rank * KS;
    for (int e = tid * 4; e < KS; e += CONSUMER_THREADS * 4) {
        float4 v = *reinterpret_cast<const float4*>(out_part + e);
#pragma unroll
        for (int w = 1; w < owins; ++w) {
            const float4 u = *reinterpret_cast<const float4*>(out_part + w * GQA_KS_MAX + e);
            v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
        }
        red_add_v4(scratch + e, v);
    }
    __threadfence();
    dsm::named_bar_sync(CONSUMER_BAR, CONSUMER_THREADS);
    unsigned* counters = p.counters + (size_t)batch * 32;
    const unsigned n_clusters = gridDim.x / CLUSTER;
    if (tid == 0) {
        const unsigned prev = atomicAdd(&counters[rank], 1u);
        sflags[0] = (prev == n_clusters - 1u);
    }
    dsm::named_bar_sync(CONSUMER_BAR, CONSUMER_THREADS);
    CF_MARK(8);
    if (sflags[0]) {
        __threadfence();
        const bool fp32_out = p.flags & 1u;
        for (int e = tid * 4; e < KS; e += CONSUMER_THREADS * 4) {
            const float4 v = ld_cg_v4(scratch + e);
            *reinterpret_cast<float4*>(scratch + e) = make_float4(0.f, 0.f, 0.f, 0.f);
            const size_t off = (size_t)batch * hidden + rank * KS + e;
            if (fp32_out) {
                *reinterpret_cast<float4*>(static_cast<float*>(p.out) + off) = v;
            } else {
                __align__(8) __half h4[4] = {__float2half_rn(v.x), __float2half_rn(v.y),
                                             __float2half_rn(v.z), __float2half_rn(v.w)};
                *reinterpret_cast<uint2*>(static_cast<__half*>(p.out) + off) = *reinterpret_cast<const uint2*>(h4);
            }
        }
        if (tid == 0) counters[rank] = 0u;
        if (residual_inplace) {
            __threadfence();
            dsm::named_bar_sync(CONSUMER_BAR, CONSUMER_THREADS);
            if (tid == 0) {
                const unsigned prev = atomicAdd(&counters[16], 1u);
                sflags[1] = (prev == (unsigned)CLUSTER - 1u);
                if (sflags[1]) counters[16] = 0u;
            }
            dsm::named_bar_sync(CONSUMER_BAR, CONSUMER_THREADS);
            if (sflags[1]) {
                for (int e = tid * 8; e < hidden; e += CONSUMER_THREADS * 8) {
                    float f[8], r8[8];
                    unpack8(*reinterpret_cast<const uint4*>(xg + e), f);
                    unpack8(*reinterpret_cast<const uint4*>(rg + e), r8);
                    __align__(16) __half hs[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) hs[k] = __float2half_rn(f[k] + r8[k]);
                    *reinterpret_cast<uint4*>(rout + e) = *reinterpret_cast<const uint4*>(hs);
                }
            }
        }
    }
    CF_MARK(9);
}

}  // namespace cfb
