/*
 * Grouped-query fused decoder attention half-layer, second generation ("group kernel").
 *
 * Why a second kernel: a B200 SM pulls at most ~64 GB/s through one 192 KB TMA ring (measured), so a
 * layer only reaches the HBM roofline when >= ~110 SMs stream at once.  With grouped-query attention the natural unit
 * of fusion is one KV head + its query heads: Llama-3-8B has 8 of them.  A thread-block cluster per unit (the first
 * generation of this kernel, removed in round 2) can give a unit 8 CTAs (64 SMs: 48 % of the roofline) -- at a 1-CTA-per-SM footprint
 * cudaOccupancyMaxActiveClusters answers 15 clusters of 8 and 7 clusters of 16 (tools/cluster_probe.cu,
 * profiles/r02_cluster_probe.txt), one short of the 16 / 8 it would take, and DSMEM does not reach outside a cluster.  Here a unit ("group") is G CTAs, G = 2^k chosen so that groups x G ~ fills the 148 SMs
 * (8B: 8 x 16 = 128; 70B shards: 8 x 16, 4 x 32, 2 x 64), and the two exchanges the fusion needs go through L2:
 *
 *   exchange 1  q|k|v:  the group's [(NQ+2)*128 x hidden] weight block is cut into 16-row x 256-col tiles, dealt to the
 *               G CTAs as contiguous runs in row-block-major order (a row block is shared by at most two CTAs); every
 *               CTA publishes its row sums, every CTA reads all 768 back, adds the (at most two) parts in a fixed order
 *               and applies RoPE itself.
 *   exchange 2  softmax state: CTA r streams KV rows [r*chunk, ...) for all NQ query heads and publishes its
 *               NQ x [m, l, o[128]]; CTA j merges dims [j*512/G, ...) over all G states in rank order and publishes the
 *               normalised fp16-rounded result; every CTA reads the 512 merged values (reduce-scatter + all-gather: an
 *               all-to-all of full states would be 67-270 KB per CTA through a ~64 GB/s SM port).
 *
 * Both exchanges use a flag-in-data protocol (what NCCL calls LL): every 64-bit word carries a float and the launch's
 * epoch, written with one st.relaxed.gpu.b64 and polled with ld.relaxed.gpu.b64 until the epoch matches.  No fence, no
 * counter, no atomics, nothing to re-zero: one L2 round trip per hop (B300_MICROARCH: L2 hit 234-262 cycles).  The first
 * version used red.global.add + __threadfence + a counter + an acquire spin per exchange and cost 3.4 us + 6.4 us per
 * layer (phase timeline, profiles/); the epoch lives in a 256-byte header at the start of the workspace and is bumped by
 * CTA 0 once it has seen every group's output partial.  The TMA ring keeps landing the next phase's tiles while a CTA
 * waits.  The output partials of the groups are summed the same way (batch 1: ll_finalize_columns, deterministic; for
 * head-parallel shards the finalising CTAs also push / poll the peers' words over NVLink = the layer's all-reduce).
 * Attention runs QK^T and PV on the tensor cores (mma.sync m16n8k16 on 128-byte-swizzled K/V tiles) whenever K/V arrive through
 * tensor maps -- a contiguous cache, or a paged pool whose address the caller also passed on the host: full tiles of 16
 * consecutive rows as tiled boxes, everything else (arbitrary page tables, a ragged last tile) through sm_100 tile::gather4
 * requests of 4 rows x 128 B, whose swizzle pattern is a function of the shared-memory address and therefore composes into the
 * same 8-row atoms (profiles/round2_gather4_probe.txt).  With 4 query heads per K/V row the CUDA-core loop is issue-bound;
 * it remains for paged launches without the host-side pool address (linear 256-byte row copies cannot swizzle).
 * Tried in round 2 and rejected on a same-box A/B (profiles/round2_gqa_onehop_vs_round1_same_box.txt, commit b31f9de):
 * a ONE-hop softmax-state exchange (every CTA gathers the G states of one head and owns that head's 128 input columns in the
 * O projection, tensor-core O GEMV over swizzled [32 x 64] boxes): exchange 2 + O phase 1.6 us shorter, but the 4x larger
 * cross-CTA output reduction and the 128-byte row segments of the O tiles cost 3.8 us elsewhere (Llama-3-8B 24.2 / 27.8 us
 * against 21.9 / 25.8).  Everything else -- the single 24 x 8 KB self-issuing
 * tile stream, fp32 reductions, fp16 rounding points of the eager model -- is as in the MHA kernel.
 * nn.Linear weight layout only (SGLANG / PAGED), NQ = 4 query heads per group; a KV head with 8 query heads (70B) gets
 * two groups.
 *
 * Cross-CTA spinning needs the waited-for CTAs to be resident or dispatched eventually: the launcher picks G with
 * groups x G x batch <= #SMs whenever it can (1 CTA / SM), and CTAs of a group are contiguous in blockIdx.x, so a
 * partially resident group only ever waits for blocks that are dispatched as earlier, complete groups drain.
 *
 * Reference: new capability (the reference kernels hard-code MHA, /root/reference/include/H100/llama/config.h:2-38,
 * llama_kernel_dispatch.cu:63-64); same operator signatures (pybind.cpp:14-43).
 */
#pragma once

#include "llama_decoder_kernel.cuh"

namespace cfb {

constexpr int G2_HIDDEN_MAX = 8192;
constexpr int G2_RB_LOCAL_MAX = 8;        // 16-row blocks a CTA can touch in the QKV phase (48 / G whole + 1 partial, G >= 8)
constexpr int G2_OROWS_MAX = 1024;        // hidden / G
constexpr int G2_GROUPS_MAX = 16;         // groups per request
constexpr int G2_G_MAX = 64;              // CTAs per group
constexpr int G2_SLOTS = 160;             // softmax-state slots per request (groups x G <= 148 whenever batch == 1)
constexpr int G2_COUNTERS = 128;          // u32 per request: [0,64) O slices, [64] finalised slices
constexpr int WS_HEADER_BYTES = 256;      // workspace header: u32 [0] epoch, [1] finalised-slice count of the running launch

template <int NQ>
struct SmemGqa2 {
    static constexpr int R = (NQ + 2) * HEAD_DIM;                 // 768 rows of q(NQ heads) | k | v
    static constexpr int PAY = HEAD_DIM + 4;                      // [m, l, -, -, o[128]]
    static constexpr int RING = 0;
    static constexpr int UNION = RING + NSTAGES * STAGE_BYTES;
    //   phase QKV : xs fp16[hidden <= 8192] | part fp32[12 warps][G2_RB_LOCAL_MAX][16]
    //   phase ATTN: attn_part fp32[12][NQ][132] | mg fp32[640]
    //   phase O   : out_part fp32[NQ*128/256][G2_OROWS_MAX]
    static constexpr int XS = UNION;
    static constexpr int PART = UNION + G2_HIDDEN_MAX * 2;
    static constexpr int QKV_BYTES = G2_HIDDEN_MAX * 2 + CONSUMER_WARPS * G2_RB_LOCAL_MAX * ROWS512 * 4;
    static constexpr int ATTN_PART = UNION;                                // fp32 [12 warps][NQ][132]
    static constexpr int MG = UNION + CONSUMER_WARPS * NQ * PAY * 4;       // fp32 [G][S2 + 2] <= 640 floats
    static constexpr int ATTN_BYTES = CONSUMER_WARPS * NQ * PAY * 4 + 640 * 4;
    static constexpr int OUT_PART = UNION;
    static constexpr int OUT_BYTES = (NQ * HEAD_DIM / 256) * G2_OROWS_MAX * 4;
    static constexpr int UNION_SIZE = QKV_BYTES > ATTN_BYTES ? (QKV_BYTES > OUT_BYTES ? QKV_BYTES : OUT_BYTES)
                                                             : (ATTN_BYTES > OUT_BYTES ? ATTN_BYTES : OUT_BYTES);
    static constexpr int QKV_FIN = UNION + UNION_SIZE;                     // fp32[R] roped q*scale | k | v
    static constexpr int AG2 = QKV_FIN + R * 4;                            // fp32[NQ*128] attention output
    static constexpr int RED = AG2 + NQ * HEAD_DIM * 4;                    // fp32[32]
    static constexpr int BARS = RED + 32 * 4;                              // u64 full[NSTAGES]
    static constexpr int FLAGS = BARS + NSTAGES * 8;
    static constexpr int TOTAL = FLAGS + 16;
    static_assert(TOTAL <= 227 * 1024, "shared-memory layout exceeds the 227 KB opt-in limit");
};

// extra kernel parameters of the group kernel (appended to KParams by composition)
struct G2Params {
    KParams k;
    unsigned long long* qkv_ll;    // (float, epoch) words [batch][G2_GROUPS_MAX][2 parts][768]
    unsigned long long* attn_ll;   // (float, epoch) words [batch][G2_SLOTS][NQ*132]
    unsigned long long* ag_ll;     // (float, epoch) words [batch][G2_GROUPS_MAX][NQ*128]
    unsigned* gcounters;           // u32 [batch][G2_COUNTERS], zero between launches
                                   // (k.header: u32 [0] epoch, [1] finalised slices of a batch > 1 launch; k.out_ll: batch == 1)
    int G;                 // CTAs per group (power of two)
    int n_groups;          // groups per request
};

// 16 output rows x 256 input columns of an [out,in] weight tile against 8 activations per lane;
// writes the 16 row sums to out[0..16).  (CUDA-core loop: 8 FMA per lane per row, then a 9-shuffle transpose-reduce.)
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

// 16 output rows x 256 input columns; ACCUMULATES the 16 row sums into acc[0..16) (owned by the calling warp)
__device__ __forceinline__ void gemv_tile_16x256_acc(const uint4* tile, const float (&x8)[8], float* acc, uint32_t lane) {
    float tmp[2];
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
        tmp[grp] = v[0];
    }
    if ((lane & 3) == 0) {
        const int r = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
        acc[r] += tmp[0];
        acc[8 + r] += tmp[1];
    }
}

template <int VARIANT, int NQ>
__global__ void __launch_bounds__(BLOCK_THREADS, 1)
llama_decoder_layer_gqa2_kernel(const __grid_constant__ G2Params gp)
{
    using S = SmemGqa2<NQ>;
    static_assert(VARIANT != CHAT, "GQA uses the nn.Linear weight layout");
    constexpr bool kPaged = (VARIANT == PAGED);
    // contiguous KV arrives through TMA and can be laid out 128-byte-swizzled, which makes ldmatrix conflict-free: that
    // variant runs QK^T and PV on the tensor cores (mma.sync m16n8k16).  Page-size-1 KV lands as linear 256-byte rows
    // (bulk copies cannot swizzle) and keeps the CUDA-core loop.
    constexpr float kScaleLog2 = 0.08838834764831845f * 1.4426950408889634f;      // 1/sqrt(128) * log2(e)
    const KParams& p = gp.k;

    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t smem_base = dsm::smem_u32(smem);
    const uint32_t tid = threadIdx.x;
    const uint32_t warp = tid >> 5;
    const uint32_t lane = tid & 31;
    const int G = gp.G;
    const uint32_t rank = blockIdx.x % G;            // CTA index inside the group
    const uint32_t gid = blockIdx.x / G;             // group index inside the request
    const uint32_t batch = blockIdx.y;

    const int hidden = p.hidden;
    const int Hq = p.n_heads, Hkv = p.n_kv_heads;
    const int qsplit = (Hq / Hkv) / NQ;                 // groups per KV head
    const int kvh = gid / qsplit;
    const int qh0 = kvh * (Hq / Hkv) + (gid % qsplit) * NQ;     // first query head of this group
    const bool writes_kv = (gid % qsplit) == 0;
    const int OROWS = hidden / G;                        // this CTA's slice of the O output dim
    const int kv_cols = Hkv * HEAD_DIM;
    const int wins = hidden / 256;

    const uint32_t full_u32 = smem_base + S::BARS;          // u64 full[NSTAGES]

    int kv_len, kv_base = 0, new_slot = 0;
    if constexpr (kPaged) {
        kv_base = p.indptr[batch];
        const int end = p.indptr[batch + 1] - 1;
        kv_len = end - kv_base;
        new_slot = p.indices[end];
    } else {
        kv_len = p.kv_len;
    }
    const int chunk = (((kv_len + G - 1) / G) + ROWS512 - 1) & ~(ROWS512 - 1);
    const int row_begin = min((int)rank * chunk, kv_len);
    const int row_end = min(row_begin + chunk, kv_len);
    const uint32_t n_qkv_tiles = (uint32_t)((S::R / ROWS512) * wins / G);       // contiguous run of the group's tiles
    const uint32_t t_first = rank * n_qkv_tiles;                                 // group-level index of this CTA's first tile
    const int rb_first = t_first / wins;
    const uint32_t n_kv_tiles = (row_end - row_begin + ROWS512 - 1) / ROWS512;
    constexpr int owins = NQ * HEAD_DIM / 256;
    const uint32_t n_o_tiles = (OROWS / ROWS512) * owins;

    CF_MARK(0);
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    // ---- tile stream: every warp requests, consumes and re-requests its own tiles (see llama_decoder_kernel.cuh) ----
    const uint64_t pol = policy_evict_first();
    const uint32_t total_tiles = n_qkv_tiles + n_kv_tiles + n_o_tiles;
    const __half* kpool = p.k_base;
    const __half* vpool = p.v_base;
    if constexpr (kPaged) {
        kpool = reinterpret_cast<const __half*>(p.k_pool_ptrs[p.layer_id]);
        vpool = reinterpret_cast<const __half*>(p.v_pool_ptrs[p.layer_id]);
    }
    // the host's copy of the pool addresses is used only if it agrees with the device table (see llama_decoder_kernel.cuh)
    const bool pool_maps = !kPaged || (p.k_base != nullptr && kpool == p.k_base && vpool == p.v_base);
    const bool use_mma = pool_maps;                      // uniform over the launch
    // paged KV: page index of this lane's row of KV tile g, fetched one ring cycle ahead (see llama_decoder_kernel.cuh).
    // Rows past the end of the chunk repeat its last row: tile::gather4 needs four valid rows, the scores mask them.
    int pre_slot0 = 0, pre_slot1 = 0;
    uint32_t pre_g0 = 0xffffffffu, pre_g1 = 0xffffffffu;
    auto page_of = [&](uint32_t g) -> int {
        const int r = min(row_begin + (int)(g - n_qkv_tiles) * ROWS512 + (int)(lane & 15), row_end - 1);
        if constexpr (kPaged) return p.indices[kv_base + r];
        else return r;
    };

    auto issue_tile = [&](uint32_t g) {
        if (g >= total_tiles) return;
        const uint32_t s = ring_stage(g);
        const uint32_t fb = full_u32 + 8 * s;
        const uint32_t dst = smem_base + S::RING + s * STAGE_BYTES;
        if (g < n_qkv_tiles) {
            if (lane == 0) {
                constexpr int BPH = HEAD_DIM / ROWS512;          // 16-row blocks per head (8)
                const uint32_t t = t_first + g;
                const int rb = t / wins, win = t % wins;         // rb: 16-row block inside q(NQ*128) | k(128) | v(128)
                int row0;
                if (rb < NQ * BPH) row0 = qh0 * HEAD_DIM + rb * ROWS512;
                else if (rb < NQ * BPH + BPH) row0 = Hq * HEAD_DIM + kvh * HEAD_DIM + (rb - NQ * BPH) * ROWS512;
                else row0 = (Hq + Hkv) * HEAD_DIM + kvh * HEAD_DIM + (rb - NQ * BPH - BPH) * ROWS512;
                dsm::mbar_arrive_expect_tx(fb, STAGE_BYTES);
                tma_load_2d(dst, &p.tm_wqkv, win * 256, row0, fb, pol);
            }
        } else if (g < n_qkv_tiles + n_kv_tiles) {
            const uint32_t i = g - n_qkv_tiles;
            const int r0 = row_begin + (int)i * ROWS512;
            const int nvalid = min(ROWS512, row_end - r0);
            int slot;
            if constexpr (kPaged) {
                const bool odd = (g / CONSUMER_WARPS) & 1u;
                slot = (odd ? pre_g1 : pre_g0) == g ? (odd ? pre_slot1 : pre_slot0) : page_of(g);
            } else {
                slot = min(r0 + (int)(lane & 15), row_end - 1);
            }
            if (use_mma) {
                // stage = K dims 0-63 | K dims 64-127 | V dims 0-63 | V dims 64-127, each [16 rows][128 B] 128-byte swizzled
                int slot0 = r0;
                bool run = nvalid == ROWS512;                // contiguous cache: every full tile is a run, no vote needed
                if constexpr (kPaged) {
                    slot0 = __shfl_sync(0xffffffffu, slot, 0);
                    run = run && __all_sync(0xffffffffu, slot == slot0 + (int)(lane & 15));
                }
                // (the pool maps cover POOL_MAP_ROWS = 2^24 slots -- 34 GB of K per layer at 8 KV heads; the limit is part of the
                //  contract of passing the host-side pool address, include/clusterfusion_b200.h)
                if (run) {                                  // 16 consecutive rows: one tiled box per quarter
                    if (lane == 0) {
                        dsm::mbar_arrive_expect_tx(fb, STAGE_BYTES);
                        tma_load_2d(dst, &p.tm_k, kvh * HEAD_DIM, slot0, fb, pol);
                        tma_load_2d(dst + 2048, &p.tm_k, kvh * HEAD_DIM + 64, slot0, fb, pol);
                        tma_load_2d(dst + 4096, &p.tm_v, kvh * HEAD_DIM, slot0, fb, pol);
                        tma_load_2d(dst + 6144, &p.tm_v, kvh * HEAD_DIM + 64, slot0, fb, pol);
                    }
                } else {                                    // arbitrary rows (or a ragged last tile): tile::gather4, 4 rows x 128 B
                    const int s1 = __shfl_down_sync(0xffffffffu, slot, 1);
                    const int s2 = __shfl_down_sync(0xffffffffu, slot, 2);
                    const int s3 = __shfl_down_sync(0xffffffffu, slot, 3);
                    if (lane == 0) dsm::mbar_arrive_expect_tx(fb, STAGE_BYTES);
                    __syncwarp();
                    if ((lane & 3) == 0) {                  // lanes 0,4,8,12: K rows 4q..4q+3; lanes 16,..,28: V rows
                        const uint32_t d = dst + (lane < 16 ? 0 : 4096) + ((lane & 15) >> 2) * 512;
                        const CUtensorMap* tm = lane < 16 ? &p.tm_kg : &p.tm_vg;
                        tma_gather4_2d(d, tm, kvh * HEAD_DIM, slot, s1, s2, s3, fb, pol);
                        tma_gather4_2d(d + 2048, tm, kvh * HEAD_DIM + 64, slot, s1, s2, s3, fb, pol);
                    }
                }
            } else {
                if constexpr (kPaged) {                     // no tensor map over the pool: linear 256-byte rows, CUDA-core loop
                    if (lane == 0) dsm::mbar_arrive_expect_tx(fb, nvalid * 2 * HEAD_DIM * 2);
                    __syncwarp();
                    if ((int)(lane & 15) < nvalid) {
                        const uint32_t d = dst + (lane & 15) * (HEAD_DIM * 2);
                        if (lane < 16) bulk_load_1d(d, kpool + (long long)slot * kv_cols + kvh * HEAD_DIM, HEAD_DIM * 2, fb, pol);
                        else bulk_load_1d(d + STAGE_BYTES / 2, vpool + (long long)slot * kv_cols + kvh * HEAD_DIM, HEAD_DIM * 2, fb, pol);
                    }
                }
            }
        } else {
            if (lane == 0) {
                const uint32_t i = g - n_qkv_tiles - n_kv_tiles;
                const int rb = i / owins, win = i % owins;       // Wo [out][in]: 16 output rows x 256 of this group's input cols
                dsm::mbar_arrive_expect_tx(fb, STAGE_BYTES);
                tma_load_2d(dst, &p.tm_wo, qh0 * HEAD_DIM + win * 256, rank * OROWS + rb * ROWS512, fb, pol);
            }
        }
        if constexpr (kPaged) {
            const uint32_t g2 = g + NSTAGES;          // the tile that will live in this stage next
            if (g2 >= n_qkv_tiles && g2 < n_qkv_tiles + n_kv_tiles) {
                const int pg = page_of(g2);
                if ((g / CONSUMER_WARPS) & 1u) { pre_slot1 = pg; pre_g1 = g2; } else { pre_slot0 = pg; pre_g0 = g2; }
            }
        }
    };

    if (lane == 0) {
        dsm::mbar_init(full_u32 + 8 * warp, 1);
        dsm::mbar_init(full_u32 + 8 * (warp + CONSUMER_WARPS), 1);
        if (tid == 0) {
            prefetch_tmap(&p.tm_wqkv);
            prefetch_tmap(&p.tm_wo);
            if (pool_maps) { prefetch_tmap(&p.tm_k); prefetch_tmap(&p.tm_v); prefetch_tmap(&p.tm_kg); prefetch_tmap(&p.tm_vg); }
        }
        dsm::mbar_fence_init();
    }
    __syncwarp();
    CF_MARK(12);
    issue_tile(warp);
    issue_tile(warp + CONSUMER_WARPS);

    __half* xs = reinterpret_cast<__half*>(smem + S::XS);
    float* part = reinterpret_cast<float*>(smem + S::PART);
    float* attn_part = reinterpret_cast<float*>(smem + S::ATTN_PART);
    float* mg = reinterpret_cast<float*>(smem + S::MG);
    float* out_part = reinterpret_cast<float*>(smem + S::OUT_PART);
    float* qkv_fin = reinterpret_cast<float*>(smem + S::QKV_FIN);
    float* ag2 = reinterpret_cast<float*>(smem + S::AG2);
    float* red = reinterpret_cast<float*>(smem + S::RED);
    uint32_t* sflags = reinterpret_cast<uint32_t*>(smem + S::FLAGS);

    const __half* xg = p.x + (size_t)batch * hidden;
    const __half* rg = p.residual_in + (size_t)batch * hidden;
    __half* rout = p.residual_out + (size_t)batch * hidden;
    const bool residual_inplace = (static_cast<const void*>(rout) == static_cast<const void*>(rg));
    unsigned long long* qkv_ll = gp.qkv_ll + ((size_t)batch * G2_GROUPS_MAX + gid) * (2 * S::R);
    unsigned long long* attn_ll = gp.attn_ll + ((size_t)batch * G2_SLOTS + (size_t)gid * G) * (NQ * S::PAY);
    unsigned long long* ag_ll = gp.ag_ll + ((size_t)batch * G2_GROUPS_MAX + gid) * (NQ * HEAD_DIM);
    unsigned* gcnt = gp.gcounters + (size_t)batch * G2_COUNTERS;

    // zero this warp's accumulation slots (smem only: legal before griddepcontrol.wait)
    for (int e = lane; e < G2_RB_LOCAL_MAX * ROWS512; e += 32) part[warp * G2_RB_LOCAL_MAX * ROWS512 + e] = 0.f;

    // the RMSNorm weight does not depend on the previous kernel: fetch it before the dependency wait
    constexpr int P0_ITERS = (G2_HIDDEN_MAX + CONSUMER_THREADS * 8 - 1) / (CONSUMER_THREADS * 8);     // 3
    uint4 wraw[P0_ITERS];
#pragma unroll
    for (int it = 0; it < P0_ITERS; ++it) {
        const int e = (it * CONSUMER_THREADS + tid) * 8;
        wraw[it] = e < hidden ? *reinterpret_cast<const uint4*>(p.rms_w + e) : make_uint4(0, 0, 0, 0);
    }

    // the RoPE factors of this thread's two projected rows (rows tid and tid + 384; not needed for v) do not depend on the previous
    // kernel either (nor does `positions`, by the PDL contract): fetched here, not right behind exchange 1
    float rope_c[2] = {0.f, 0.f}, rope_s[2] = {0.f, 0.f};
    {
        const float* cosp = p.cos;
        const float* sinp = p.sin;
        if constexpr (kPaged) {
            cosp = p.cos + p.positions[batch] * HEAD_DIM;
            sinp = cosp + HEAD_DIM / 2;
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int e = (int)tid + u * CONSUMER_THREADS;
            if ((e >> 7) <= NQ) { rope_c[u] = cosp[e & 63]; rope_s[u] = sinp[e & 63]; }
        }
    }

    asm volatile("griddepcontrol.wait;" ::: "memory");

    // this launch's epoch: every flag-in-data word written below carries it (header zeroed once by the caller,
    // bumped by the previous launch's last finaliser; 0 is never used as a flag)
    const unsigned epoch = __ldcg(p.header);
    const unsigned flag = ll_flag_of_epoch(epoch);
    const unsigned tp_flag = p.tp_world > 1 ? ll_flag_of_epoch(__ldcg(p.header + 3)) : 0u;

    // ---- phase 0: fused residual add + RMSNorm over the FULL vector (every CTA needs all of it: rows are split).
    //      One pass: x and residual are loaded once and stay in registers across the block reduction. ----
    {
        float f[P0_ITERS][8];
        float ss = 0.f;
        uint4 xr[P0_ITERS], rr[P0_ITERS];
#pragma unroll
        for (int it = 0; it < P0_ITERS; ++it) {
            const int e = (it * CONSUMER_THREADS + tid) * 8;
            if (e < hidden) {
                xr[it] = *reinterpret_cast<const uint4*>(xg + e);
                rr[it] = *reinterpret_cast<const uint4*>(rg + e);
            } else {
                xr[it] = make_uint4(0, 0, 0, 0);
                rr[it] = make_uint4(0, 0, 0, 0);
            }
        }
        const int own_lo = rank * OROWS, own_hi = own_lo + OROWS;       // residual_out slice written by group 0
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
            if (gid == 0 && !residual_inplace && e >= own_lo && e < own_hi)
                *reinterpret_cast<uint4*>(rout + e) = *reinterpret_cast<const uint4*>(hs);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        if (lane == 0) red[warp] = ss;
        dsm::named_bar_sync(CONSUMER_BAR, CONSUMER_THREADS);
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
                __align__(16) __half xn[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) xn[k] = __float2half_rn(round_h(f[it][k] * rstd) * w8[k]);
                *reinterpret_cast<uint4*>(xs + e) = *reinterpret_cast<const uint4*>(xn);
            }
        }
        dsm::named_bar_sync(CONSUMER_BAR, CONSUMER_THREADS);
    }
    CF_MARK(1);

    uint32_t gbase = 0;
    // ---- phase 1: QKV GEMV over this CTA's run of tiles ---------------------------------------------------
    for (uint32_t i = first_tile(gbase, warp); i < n_qkv_tiles; i += CONSUMER_WARPS) {
        const uint32_t g = gbase + i, s = ring_stage(g);
        const uint32_t t = t_first + i;
        const int rb = t / wins, win = t % wins;
        float x8[8];
        unpack8(*reinterpret_cast<const uint4*>(xs + win * 256 + lane * 8), x8);
        ring_wait_full(full_u32, g);
        gemv_tile_16x256_acc(reinterpret_cast<const uint4*>(smem + S::RING + s * STAGE_BYTES), x8,
                             part + (warp * G2_RB_LOCAL_MAX + (rb - rb_first)) * ROWS512, lane);
        __syncwarp();
        issue_tile(g + NSTAGES);
    }
    gbase += n_qkv_tiles;
    dsm::named_bar_sync(CONSUMER_BAR, CONSUMER_THREADS);
    CF_MARK(2);
    // ---- exchange 1: publish this CTA's row sums as (value, epoch) words; part = position of this CTA among the
    //      (at most two) CTAs that share the row block ------------------------------------------------------------
    {
        const int rb_last = (t_first + n_qkv_tiles - 1) / wins;
        const int n_loc = (rb_last - rb_first + 1) * ROWS512;
        for (int o = tid; o < n_loc; o += CONSUMER_THREADS) {
            float a = 0.f;
#pragma unroll
            for (int w = 0; w < CONSUMER_WARPS; ++w) a += part[w * G2_RB_LOCAL_MAX * ROWS512 + o];
            const int rb = rb_first + o / ROWS512;
            const int part_id = (int)rank - (int)((uint32_t)(rb * wins) / n_qkv_tiles);
            ll_store(qkv_ll + part_id * S::R + rb_first * ROWS512 + o, a, flag);
        }
    }
    CF_MARK(3);

    // ---- RoPE (NeoX), new K/V out ---------------------------------------------------------------------
    {
        // read a projected row: sum of its 1 or 2 published parts, each validated by its own epoch
        auto n_parts = [&](int e) {
            const int rb = e / ROWS512;
            return (int)((uint32_t)(rb * wins + wins - 1) / n_qkv_tiles) - (int)((uint32_t)(rb * wins) / n_qkv_tiles) + 1;
        };
        // one L2 round trip: every thread probes all words of its own two rows back to back, resolves them, and parks
        // the fp16-rounded sums in shared memory (the activation buffer `xs` is dead by now) for the RoPE pairing
        float* qkv_raw = reinterpret_cast<float*>(smem + S::XS);
        {
            static_assert(S::R == 2 * CONSUMER_THREADS, "two projected rows per thread");
            unsigned long long w0[2], w1[2];
            bool two[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int e = tid + u * CONSUMER_THREADS;
                two[u] = n_parts(e) > 1;
                w0[u] = ll_load(qkv_ll + e);
                w1[u] = two[u] ? ll_load(qkv_ll + S::R + e) : 0ull;
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int e = tid + u * CONSUMER_THREADS;
                float v = ll_resolve(qkv_ll + e, w0[u], flag, p.header + 2);
                if (two[u]) v += ll_resolve(qkv_ll + S::R + e, w1[u], flag, p.header + 2);
                qkv_raw[e] = round_h(v);                         // q / k / v leave the projection as fp16 (eager model)
            }
        }
        dsm::named_bar_sync(CONSUMER_BAR, CONSUMER_THREADS);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int e = (int)tid + u * CONSUMER_THREADS;
            const int hd = e >> 7, d = e & 127;                  // hd < NQ: query head; NQ: k; NQ+1: v
            const float a = qkv_raw[e];
            const float bv = qkv_raw[e ^ 64];
            if (hd <= NQ) {
                const float b = bv;
                const float rot = (d & 64) ? fmaf(a, rope_c[u], b * rope_s[u]) : fmaf(a, rope_c[u], -b * rope_s[u]);
                const __half rh = __float2half_rn(rot);
                // CUDA-core loop: q pre-scaled in fp32; mma loop: q must stay an exact fp16 value, the scale goes onto S
                qkv_fin[e] = (hd < NQ && !use_mma) ? __half2float(rh) * kScaleLog2 : __half2float(rh);
                if (hd == NQ && rank == 0 && writes_kv) {
                    if constexpr (kPaged) {
                        __half* kp = reinterpret_cast<__half*>(p.k_pool_ptrs[p.layer_id]);
                        kp[(size_t)new_slot * kv_cols + kvh * HEAD_DIM + d] = rh;
                    } else {
                        p.k_new[kvh * HEAD_DIM + d] = rh;
                    }
                }
            } else {
                qkv_fin[e] = a;
                if (rank == 0 && writes_kv) {
                    const __half vh = __float2half_rn(a);
                    if constexpr (kPaged) {
                        __half* vp = reinterpret_cast<__half*>(p.v_pool_ptrs[p.layer_id]);
                        vp[(size_t)new_slot * kv_cols + kvh * HEAD_DIM + d] = vh;
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
        if (use_mma) {
            // ---- tensor-core loop.  Per 16-key tile and warp: S[head][key] = Q K^T (M = 16 rows of which NQ = 4 are real
            // heads, N = 2 x 8 keys, K = 128 dims: 16 mma), online softmax on the C fragments (a head's 16 scores sit in one
            // quad), then O[head][dim] += P V (P's C fragments are the A fragments of the second product: 16 mma).  ~150
            // instructions per tile instead of ~1150 on the CUDA cores, where 4 heads x 2 FMA per K/V byte made this phase
            // issue-bound (7.3 us for 33.5 MB at kv 8K).  P is rounded to fp16 for the product, like the probabilities of the
            // eager fp16 model.
            const int g4 = lane >> 2, t4 = lane & 3;              // fragment coordinates: row / column pair
            uint32_t qa[8][2];                                    // A fragments of Q: [k-step][k lo / k hi], rows >= NQ are zero
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
#pragma unroll
                for (int hi = 0; hi < 2; ++hi) {
                    const int d = ks * 16 + hi * 8 + t4 * 2;
                    const float q0 = g4 < NQ ? qkv_fin[g4 * HEAD_DIM + d] : 0.f;
                    const float q1 = g4 < NQ ? qkv_fin[g4 * HEAD_DIM + d + 1] : 0.f;
                    const __half2 h2 = __floats2half2_rn(q0, q1);
                    qa[ks][hi] = *reinterpret_cast<const uint32_t*>(&h2);
                }
            }
            float oacc[16][4];
#pragma unroll
            for (int t = 0; t < 16; ++t) { oacc[t][0] = 0.f; oacc[t][1] = 0.f; oacc[t][2] = 0.f; oacc[t][3] = 0.f; }
            float mrun = -INFINITY, lrun = 0.f;                   // this lane's head g4 (replicated over the quad; l is a quad-partial)
            const int lrow = lane & 7, lmat = lane >> 3;
            for (uint32_t i = first_tile(gbase, warp); i < n_kv_tiles; i += CONSUMER_WARPS) {
                const uint32_t g = gbase + i, s = ring_stage(g);
                ring_wait_full(full_u32, g);
                const uint32_t st = smem_base + S::RING + s * STAGE_BYTES;
                const int rows_left = row_end - (row_begin + (int)i * ROWS512);
                // S = Q K^T
                float sc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    const int key = (lmat >> 1) * 8 + lrow, chunk = (ks & 3) * 2 + (lmat & 1);
                    const uint32_t addr = st + (ks >> 2) * 2048 + key * 128 + ((chunk ^ (key & 7)) << 4);
                    uint32_t b0, b1, b2, b3;
                    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                                 : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3) : "r"(addr));
                    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                                 : "+f"(sc[0][0]), "+f"(sc[0][1]), "+f"(sc[0][2]), "+f"(sc[0][3])
                                 : "r"(qa[ks][0]), "r"(0u), "r"(qa[ks][1]), "r"(0u), "r"(b0), "r"(b1));
                    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                                 : "+f"(sc[1][0]), "+f"(sc[1][1]), "+f"(sc[1][2]), "+f"(sc[1][3])
                                 : "r"(qa[ks][0]), "r"(0u), "r"(qa[ks][1]), "r"(0u), "r"(b2), "r"(b3));
                }
                // online softmax for head g4: scores of keys 2*t4, 2*t4+1 (n-tile 0) and 8+2*t4, 9+2*t4 (n-tile 1)
                float s4[4];
                s4[0] = (2 * t4 < rows_left) ? sc[0][0] * kScaleLog2 : -INFINITY;
                s4[1] = (2 * t4 + 1 < rows_left) ? sc[0][1] * kScaleLog2 : -INFINITY;
                s4[2] = (8 + 2 * t4 < rows_left) ? sc[1][0] * kScaleLog2 : -INFINITY;
                s4[3] = (9 + 2 * t4 < rows_left) ? sc[1][1] * kScaleLog2 : -INFINITY;
                float mx = fmaxf(fmaxf(s4[0], s4[1]), fmaxf(s4[2], s4[3]));
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
                const float m_new = fmaxf(mrun, mx);
                const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
                const float corr = dsm::exp2_diff(mrun, m_use);
                mrun = m_new;
                float pr[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) pr[k] = dsm::fast_exp2(s4[k] - m_use);           // -inf -> 0
                const __half2 p01 = __floats2half2_rn(pr[0], pr[1]), p23 = __floats2half2_rn(pr[2], pr[3]);
                // the row sum uses the same fp16-rounded probabilities that multiply V
                lrun = lrun * corr + (__low2float(p01) + __high2float(p01)) + (__low2float(p23) + __high2float(p23));
                const uint32_t pa0 = *reinterpret_cast<const uint32_t*>(&p01), pa2 = *reinterpret_cast<const uint32_t*>(&p23);
                // O = O * corr + P V
#pragma unroll
                for (int jd = 0; jd < 8; ++jd) {
                    const int key = (lmat & 1) * 8 + lrow, chunk = (jd & 3) * 2 + (lmat >> 1);
                    const uint32_t addr = st + 4096 + (jd >> 2) * 2048 + key * 128 + ((chunk ^ (key & 7)) << 4);
                    uint32_t b0, b1, b2, b3;
                    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                                 : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3) : "r"(addr));
                    oacc[2 * jd][0] *= corr; oacc[2 * jd][1] *= corr;
                    oacc[2 * jd + 1][0] *= corr; oacc[2 * jd + 1][1] *= corr;
                    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                                 : "+f"(oacc[2 * jd][0]), "+f"(oacc[2 * jd][1]), "+f"(oacc[2 * jd][2]), "+f"(oacc[2 * jd][3])
                                 : "r"(pa0), "r"(0u), "r"(pa2), "r"(0u), "r"(b0), "r"(b1));
                    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                                 : "+f"(oacc[2 * jd + 1][0]), "+f"(oacc[2 * jd + 1][1]), "+f"(oacc[2 * jd + 1][2]), "+f"(oacc[2 * jd + 1][3])
                                 : "r"(pa0), "r"(0u), "r"(pa2), "r"(0u), "r"(b2), "r"(b3));
                }
                __syncwarp();
                issue_tile(g + NSTAGES);
            }
            // this warp's state of head g4 -> slot [warp][head]; l is summed over the quad first
            lrun += __shfl_xor_sync(0xffffffffu, lrun, 1);
            lrun += __shfl_xor_sync(0xffffffffu, lrun, 2);
            if (g4 < NQ) {
                float* slot = attn_part + (warp * NQ + g4) * S::PAY;
                if (t4 == 0) { slot[0] = mrun; slot[1] = lrun; }
#pragma unroll
                for (int t = 0; t < 16; ++t)
                    *reinterpret_cast<float2*>(slot + 4 + t * 8 + t4 * 2) = make_float2(oacc[t][0], oacc[t][1]);
            }
        } else if constexpr (kPaged) {
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
                float sc[NQ][8];
    #pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                    const int row = 2 * jj + sub;
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
                    const int row = 2 * jj + sub;
                    uint4 raw = vt[row * 16 + c];
                    if (row >= rows_left) raw = make_uint4(0, 0, 0, 0);
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
                __syncwarp();
                issue_tile(g + NSTAGES);
            }
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
            // block merge of all NQ heads in one round: every warp writes its NQ states, then thread (h, d) folds the 12
            // warps (and, on rank 0, the current token) in warp order and publishes the result straight from registers as
            // (value, epoch) words -- exchange 2, hop A.
    #pragma unroll
            for (int h = 0; h < NQ; ++h) {
                if (sub == 0) {
                    float* slot = attn_part + (warp * NQ + h) * S::PAY;
                    if (c == 0) { slot[0] = m[h]; slot[1] = l[h]; }
                    *reinterpret_cast<float4*>(slot + 4 + c * 8) = make_float4(o8[h][0], o8[h][1], o8[h][2], o8[h][3]);
                    *reinterpret_cast<float4*>(slot + 4 + c * 8 + 4) = make_float4(o8[h][4], o8[h][5], o8[h][6], o8[h][7]);
                }
            }
        }
        gbase += n_kv_tiles;
        CF_MARK(5);
        if (warp < NQ) {                                         // score of the current token against query head `warp`
            float a = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                a = fmaf(qkv_fin[warp * HEAD_DIM + lane * 4 + k], qkv_fin[NQ * HEAD_DIM + lane * 4 + k], a);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            if (lane == 0) red[CONSUMER_WARPS + warp] = use_mma ? a * kScaleLog2 : a;
        }
        dsm::named_bar_sync(CONSUMER_BAR, CONSUMER_THREADS);
        for (int e = tid; e < NQ * HEAD_DIM; e += CONSUMER_THREADS) {
            const int h = e >> 7, d = e & 127;
            const bool with_new = (rank == 0);
            const float s_new = red[CONSUMER_WARPS + h];
            float M = with_new ? s_new : -INFINITY;
#pragma unroll
            for (int gI = 0; gI < CONSUMER_WARPS; ++gI) M = fmaxf(M, attn_part[(gI * NQ + h) * S::PAY]);
            float L = 0.f, O = 0.f;
#pragma unroll
            for (int gI = 0; gI < CONSUMER_WARPS; ++gI) {
                const float* sl = attn_part + (gI * NQ + h) * S::PAY;
                const float w = dsm::exp2_diff(sl[0], M);
                L = fmaf(sl[1], w, L);
                O = fmaf(sl[4 + d], w, O);
            }
            if (with_new) {
                const float w = dsm::exp2_diff(s_new, M);
                L += w;
                O = fmaf(qkv_fin[(NQ + 1) * HEAD_DIM + d], w, O);
            }
            unsigned long long* st = attn_ll + (size_t)rank * (NQ * S::PAY) + h * S::PAY;
            ll_store(st + 4 + d, O, flag);
            if (d == 0) { ll_store(st, M, flag); ll_store(st + 1, L, flag); }
        }
        // this CTA owns merged dims [rank*S2, +S2) of the group's NQ*128: gather [m, l, o[S2]] of every rank ...
        const int S2 = NQ * HEAD_DIM / G;                        // 64 / 32 / 16 / 8 for G = 8 / 16 / 32 / 64
        const int hh = (rank * S2) >> 7, d0 = (rank * S2) & 127;
        {
            const int nw = G * (S2 + 2);                         // <= 640 words: at most two per thread
            const unsigned long long* src[2];
            unsigned long long w[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int i = tid + u * CONSUMER_THREADS;
                const int r = i / (S2 + 2), j = i % (S2 + 2);
                src[u] = attn_ll + (size_t)r * (NQ * S::PAY) + hh * S::PAY + (j < 2 ? j : 4 + d0 + (j - 2));
                w[u] = i < nw ? ll_load(src[u]) : 0ull;
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int i = tid + u * CONSUMER_THREADS;
                if (i < nw) mg[i] = ll_resolve(src[u], w[u], flag, p.header + 2);
            }
        }
        dsm::named_bar_sync(CONSUMER_BAR, CONSUMER_THREADS);
        // ... merge them in rank order (deterministic) and publish the normalised, fp16-rounded slice (hop B)
        if ((int)tid < S2) {
            float M = -INFINITY;
            for (int r = 0; r < G; ++r) M = fmaxf(M, mg[r * (S2 + 2)]);
            float L = 0.f, O = 0.f;
            for (int r = 0; r < G; ++r) {
                const float w = dsm::exp2_diff(mg[r * (S2 + 2)], M);
                L = fmaf(mg[r * (S2 + 2) + 1], w, L);
                O = fmaf(mg[r * (S2 + 2) + 2 + tid], w, O);
            }
            ll_store(ag_ll + rank * S2 + tid, round_h(O / L), flag);   // attention output leaves as fp16 (eager model)
        }
        // every CTA reads the whole merged attention output of the group
        {
            constexpr int NA = NQ * HEAD_DIM;                    // 512 words: at most two per thread
            unsigned long long w[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int e = tid + u * CONSUMER_THREADS;
                w[u] = e < NA ? ll_load(ag_ll + e) : 0ull;
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int e = tid + u * CONSUMER_THREADS;
                if (e < NA) ag2[e] = ll_resolve(ag_ll + e, w[u], flag, p.header + 2);
            }
        }
        dsm::named_bar_sync(CONSUMER_BAR, CONSUMER_THREADS);
    }
    CF_MARK(6);

    // ---- phase 3: O GEMV for output rows [rank*OROWS, +OROWS) over this group's NQ*128 input columns --------
    {
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
                             out_part + win * G2_OROWS_MAX + rb * ROWS512, lane);
            __syncwarp();
            issue_tile(g + NSTAGES);
        }
        dsm::named_bar_sync(CONSUMER_BAR, CONSUMER_THREADS);
    }
    CF_MARK(7);

    if (gridDim.y == 1 && p.out_ll != nullptr) {
        // ---- cross-group reduction, batch == 1: publish the fp32 partial of this rank's output slice as (value, epoch)
        //      words; the n_groups CTAs that share the slice each sum 1/n_groups of its columns over all groups in group
        //      order (deterministic; see ll_finalize_columns in llama_decoder_kernel.cuh) ----
        unsigned long long* mine = p.out_ll + (size_t)gid * hidden + rank * OROWS;
        for (int e = tid * 2; e < OROWS; e += CONSUMER_THREADS * 2) {
            float2 v = *reinterpret_cast<const float2*>(out_part + e);
#pragma unroll
            for (int w = 1; w < owins; ++w) {
                const float2 u = *reinterpret_cast<const float2*>(out_part + w * G2_OROWS_MAX + e);
                v.x += u.x; v.y += u.y;
            }
            ll_store2(mine + e, v.x, v.y, flag);
        }
        CF_MARK(8);
        const int ng = gp.n_groups;
        const int lo = (int)((long long)gid * OROWS / ng), hi = (int)((long long)(gid + 1) * OROWS / ng);
        ll_finalize_columns(p, p.out_ll, hidden, ng, rank * OROWS, lo, hi, flag, tp_flag, p.out, (p.flags & 1u) != 0, tid, CONSUMER_THREADS);
        // a CTA that got here has seen every group's partial of its slice, and a group only gets past its exchanges once
        // all of its CTAs are past phase 0: CTA 0 may bump the epoch and (in-place form) overwrite `residual`
        if (blockIdx.x == 0) {
            if (tid == 0) {
                asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(p.header) : "memory");
                if (p.tp_world > 1) asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(p.header + 3) : "memory");
            }
            if (residual_inplace) {
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
        CF_MARK(9);
        return;
    }
    // ---- cross-group reduction, batch > 1: fp32 red into scratch, last arriver of the slice finalises ----------------
    float* scratch = p.scratch + (size_t)batch * hidden + rank * OROWS;
    for (int e = tid * 4; e < OROWS; e += CONSUMER_THREADS * 4) {
        float4 v = *reinterpret_cast<const float4*>(out_part + e);
#pragma unroll
        for (int w = 1; w < owins; ++w) {
            const float4 u = *reinterpret_cast<const float4*>(out_part + w * G2_OROWS_MAX + e);
            v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
        }
        red_add_v4(scratch + e, v);
    }
    __threadfence();
    dsm::named_bar_sync(CONSUMER_BAR, CONSUMER_THREADS);
    if (tid == 0) {
        const unsigned prev = atomicAdd(&gcnt[rank], 1u);
        sflags[0] = (prev == (unsigned)gp.n_groups - 1u);
    }
    dsm::named_bar_sync(CONSUMER_BAR, CONSUMER_THREADS);
    CF_MARK(8);
    if (sflags[0]) {
        __threadfence();
        if (tid == 0) {
            // a slice is finalised only after every group reached the end, so when the launch's last slice is, no CTA
            // will read another flag-in-data word: bump the workspace epoch for the next launch
            const unsigned prevf = atomicAdd(p.header + 1, 1u);
            if (prevf == gridDim.y * (unsigned)G - 1u) { p.header[1] = 0u; p.header[0] = epoch + 1u; }
        }
        const bool fp32_out = p.flags & 1u;
        for (int e = tid * 4; e < OROWS; e += CONSUMER_THREADS * 4) {
            const float4 v = ld_cg_v4(scratch + e);
            *reinterpret_cast<float4*>(scratch + e) = make_float4(0.f, 0.f, 0.f, 0.f);
            const size_t off = (size_t)batch * hidden + rank * OROWS + e;
            if (fp32_out) {
                *reinterpret_cast<float4*>(static_cast<float*>(p.out) + off) = v;
            } else {
                __align__(8) __half h4[4] = {__float2half_rn(v.x), __float2half_rn(v.y),
                                             __float2half_rn(v.z), __float2half_rn(v.w)};
                *reinterpret_cast<uint2*>(static_cast<__half*>(p.out) + off) = *reinterpret_cast<const uint2*>(h4);
            }
        }
        if (tid == 0) gcnt[rank] = 0u;
        if (residual_inplace) {
            __threadfence();
            dsm::named_bar_sync(CONSUMER_BAR, CONSUMER_THREADS);
            if (tid == 0) {
                const unsigned prev = atomicAdd(&gcnt[64], 1u);
                sflags[1] = (prev == (unsigned)G - 1u);
                if (sflags[1]) gcnt[64] = 0u;
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
