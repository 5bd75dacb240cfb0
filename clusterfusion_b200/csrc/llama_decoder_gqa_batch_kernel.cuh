/*
 * Batched paged decode for grouped-query shapes: the weights are streamed ONCE per chunk of up to eight requests
 * (SURVEY.md section 8 row f3 for Llama-3-8B / Llama-2-70B shapes; the MHA counterpart is llama_decoder_batch8_kernel.cuh).
 *
 * Same execution model as the group kernel (llama_decoder_gqa2_kernel.cuh): a (KV head, 4 query heads) group is G plain
 * CTAs (groups x G ~ 128 of the 148 SMs), every warp streams its own 8 KB tiles through the 24-stage ring, and the CTAs of a
 * group talk through (value, epoch) words in L2.  What changes with a chunk of requests on the N = 8 dimension of
 * mma.sync.m16n8k16:
 *
 *   phase 0    the residual add + RMSNorm is K-split: a CTA normalises only its hidden / G input columns of the eight
 *              requests (4 KB of fp16 activations instead of 64 KB); the per-request sums of squares are exchanged as
 *              G x 8 words.
 *   phase 1    QKV projection on the tensor cores over this CTA's column slice: weight tiles are [32 rows x 128 cols] =
 *              two 128-byte-swizzled TMA boxes read with ldmatrix as the A operand, the requests are the B operand.  A warp
 *              owns two of the group's 24 row blocks and keeps their accumulators in registers across the slice.
 *   exchange 1 reduce-scatter: CTA j sums, in rank order, the G partials of the rows that make up RoPE pairs
 *              {i, i + 64 : i in [j*64/G, (j+1)*64/G)} of the six 128-row slots q0..q3 | k | v, rounds to fp16, applies RoPE,
 *              writes the new K / V rows into the pool and publishes the final values packed two fp16 per word.  There is no
 *              all-gather: a CTA reads q (and k, v of the new token) only for the requests whose KV rows it streams.
 *   phase 2    attention.  The KV rows of the chunk's requests are concatenated and cut into G equal ranges, so a CTA streams
 *              one or two SEGMENTS (request, row range) however ragged the batch is; all 12 warps work on the same segment
 *              (tensor-core QK^T / PV as in the group kernel), their states are folded in shared memory and published per
 *              (rank, request).  The CTA that holds a request's last row also folds in the new token.
 *   exchange 2 CTA j merges dims [j*512/G, ...) of every request over the ranks that hold a segment of it, publishes the
 *              normalised fp16 output; every CTA reads the 8 x 512 outputs (2048 words).
 *   phase 3    O projection on the tensor cores: output rows [rank*hidden/G, ...) x this group's 512 input columns, one row
 *              block per warp at a time, C fragments leave as (value, epoch) words; the n_groups CTAs sharing a row slice each
 *              sum 1/n_groups of its rows over all groups in group order (deterministic, no atomics, nothing to re-zero).
 *
 * Tile -> warp mapping: global tile g belongs to warp g % 12 (ring discipline, llama_decoder_kernel.cuh).  For a warp to
 * accumulate a row block across its column tiles, the tiles of a GEMM phase are enumerated so that index i = slot + 12 * k
 * means row block slot + 12 * (k / W), column tile k % W; indices whose row block does not exist are EMPTY tiles (the barrier
 * completes with a plain arrive, nothing is loaded) -- the ring parity stays a pure function of g.
 *
 * PAGED variant only, page size 1, batch >= 2 (batch 1 uses the group kernel).  K/V arrive through tensor maps over the pools when
 * the caller passed the pool addresses on the host (tiled boxes / tile::gather4, 128-byte swizzled), else as linear 128-byte row
 * pieces read without the swizzle.  hidden / G must be a multiple of 128 and <= 1024.
 * Hops that read many words per thread (the reduce-scatter, the gather of the merged outputs, the cross-group sum) use 16-byte
 * flag-in-data loads -- two adjacent words per request, each still validated by its own epoch (8- and 32-byte loads, a one-hop state
 * exchange, L2 prefetches at the stall points: measured and rejected, profiles/round2_gqa_rejected_variants.txt).
 *
 * Reference: /root/reference/include/H100/llama/llama_kernel_batch_sglang_dispatch.cu:89 (the reference launches one cluster
 * per (head, request) and re-reads the weights for every request; grouped-query shapes are a new capability).
 */
#pragma once

#include "llama_decoder_gqa2_kernel.cuh"
#include "llama_decoder_batch_kernel.cuh"

namespace cfb {

constexpr int GB_BC = 8;                  // requests per chunk (the N dimension of the MMA)
constexpr int GB_CTAS_MAX = 160;          // groups x G per chunk (<= resident CTAs)
constexpr int GB_QKVF_WORDS = 384;        // final q|k|v words per request: 6 slots x 64 RoPE pairs
constexpr int GB_STATE_WORDS = 4 * (HEAD_DIM + 4);   // softmax-state words per (rank, request)
constexpr int GB_AG_WORDS = 256;          // merged attention output per request: 512 fp16, two per word

struct SmemGqaB {
    static constexpr int NQ = 4, BC = GB_BC;
    static constexpr int R = (NQ + 2) * HEAD_DIM;                          // 768 rows of q(4 heads) | k | v
    static constexpr int PAY = HEAD_DIM + 4;
    static constexpr int RING = 0;
    static constexpr int UNION = RING + NSTAGES * STAGE_BYTES;
    //   phase 0/1 : xs fp16 [8][BK_XS_STRIDE] | ssall fp32 [64][8] | sseg fp32 [64]
    static constexpr int XS = UNION;
    static constexpr int SSALL = XS + BC * BK_XS_STRIDE * 2;
    static constexpr int SSEG = SSALL + G2_G_MAX * BC * 4;
    static constexpr int QKV_BYTES = SSEG + 64 * 4 - UNION;               // 18816
    //   exchange 1: stage1 fp32 [G parts][768 / G rows][8] = 6144 floats
    static constexpr int STAGE1 = UNION;
    static constexpr int X1_BYTES = R * BC * 4;                            // 24576
    //   phase 2   : attn_part fp32 [12 warps][4 heads][132]
    static constexpr int ATTN_PART = UNION;
    static constexpr int ATTN_BYTES = CONSUMER_WARPS * NQ * PAY * 4;       // 25344
    //   phase 3   : ag16 fp16 [8][512 + 8]
    static constexpr int AG16 = UNION;
    static constexpr int AG_STRIDE = NQ * HEAD_DIM + 8;
    static constexpr int O_BYTES = BC * AG_STRIDE * 2;                     // 8320
    static constexpr int UNION_BYTES = ATTN_BYTES;
    static_assert(QKV_BYTES <= UNION_BYTES && X1_BYTES <= UNION_BYTES && O_BYTES <= UNION_BYTES, "union too small");
    static constexpr int QSEG = UNION + UNION_BYTES;                       // fp16 [768]: q | k_new | v_new of the current segment
    static constexpr int META = QSEG + R * 2;                              // int [8][4] requests, [8][4] segments, u32 [9] tile0, [3] misc
    static constexpr int RSTD = META + (32 + 32 + 12) * 4;                 // fp32 [8]
    static constexpr int RED = RSTD + BC * 4;                              // fp32 [16]
    static constexpr int BARS = RED + 16 * 4;                              // u64 full[NSTAGES]
    static constexpr int TOTAL = BARS + NSTAGES * 8;
    static_assert(BARS % 8 == 0, "mbarrier alignment");
    static_assert(TOTAL <= 227 * 1024, "shared-memory layout exceeds the 227 KB opt-in limit");
};

struct GBParams {
    KParams k;
    unsigned long long* ss_ll;      // (float, epoch) words [chunks][GB_CTAS_MAX][8]            partial sums of squares
    unsigned long long* qkvp_ll;    //                      [chunks][GB_CTAS_MAX][768][8]       QKV partials of a CTA's column slice
    unsigned long long* qkvf_ll;    // (2 x fp16, epoch)    [chunks][G2_GROUPS_MAX][8][384]     final q | k | v
    unsigned long long* attn_ll;    // (float, epoch)       [chunks][GB_CTAS_MAX][8][528]       softmax states per (rank, request)
    unsigned long long* ag_ll;      // (2 x fp16, epoch)    [chunks][G2_GROUPS_MAX][8][256]     merged attention output
    unsigned long long* out_ll;     // (float, epoch)       [chunks][G2_GROUPS_MAX][hidden][8]  O-projection partials per group
    int G;                          // CTAs per group (power of two, hidden / G in {128, ..., 1024})
    int n_groups;                   // groups per request
};

__global__ void __launch_bounds__(BLOCK_THREADS, 1)
llama_decoder_layer_gqa_batch_kernel(const __grid_constant__ GBParams gp)
{
    using S = SmemGqaB;
    constexpr int NQ = S::NQ, BC = S::BC;
    constexpr float kScaleLog2 = 0.08838834764831845f * 1.4426950408889634f;      // 1/sqrt(128) * log2(e)
    const KParams& p = gp.k;

    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t smem_base = dsm::smem_u32(smem);
    const uint32_t tid = threadIdx.x;
    const uint32_t warp = tid >> 5;
    const uint32_t lane = tid & 31;
    const int G = gp.G;
    const uint32_t rank = blockIdx.x % G;            // CTA index inside the group
    const uint32_t gid = blockIdx.x / G;             // group index inside the chunk
    const uint32_t chunk = blockIdx.y;
    const uint32_t cta = blockIdx.x;                 // gid * G + rank < GB_CTAS_MAX
    const int b0 = (int)chunk * BC;
    const int nb = min(BC, p.batch - b0);

    const int hidden = p.hidden;
    const int Hq = p.n_heads, Hkv = p.n_kv_heads;
    const int qsplit = (Hq / Hkv) / NQ;                 // groups per KV head
    const int kvh = gid / qsplit;
    const int qh0 = kvh * (Hq / Hkv) + (gid % qsplit) * NQ;     // first query head of this group
    const bool writes_kv = (gid % qsplit) == 0;
    const int KS = hidden / G;                           // input columns of this CTA in the QKV phase = its output rows in the O phase
    const int WPC = KS / 128;                            // 128-column tiles per row block
    const int RB_O = KS / ROWS256;                       // 32-row blocks of this CTA's O slice
    const int kv_cols = Hkv * HEAD_DIM;

    const uint32_t full_u32 = smem_base + S::BARS;

    __half* xs = reinterpret_cast<__half*>(smem + S::XS);
    float* ssall = reinterpret_cast<float*>(smem + S::SSALL);
    float* sseg = reinterpret_cast<float*>(smem + S::SSEG);
    float* stage1 = reinterpret_cast<float*>(smem + S::STAGE1);
    float* attn_part = reinterpret_cast<float*>(smem + S::ATTN_PART);
    __half* ag16 = reinterpret_cast<__half*>(smem + S::AG16);
    __half* qseg = reinterpret_cast<__half*>(smem + S::QSEG);
    int* mreq = reinterpret_cast<int*>(smem + S::META);                       // [b][4]: kv_base, position, new_slot, first rank | owner << 16
    int* mseg = mreq + 32;                                                    // [s][4]: request | owner << 8, row begin, row end, -
    uint32_t* stile0 = reinterpret_cast<uint32_t*>(mseg + 32);                // [9]: first KV tile of segment s (entries >= n_seg: total)
    int* mmisc = reinterpret_cast<int*>(stile0 + 9);                          // [0] n_seg
    float* rstd_s = reinterpret_cast<float*>(smem + S::RSTD);
    float* red = reinterpret_cast<float*>(smem + S::RED);

    unsigned long long* ss_ll = gp.ss_ll + (size_t)chunk * GB_CTAS_MAX * BC;
    unsigned long long* qkvp_ll = gp.qkvp_ll + (size_t)chunk * GB_CTAS_MAX * S::R * BC;
    unsigned long long* qkvf_ll = gp.qkvf_ll + ((size_t)chunk * G2_GROUPS_MAX + gid) * BC * GB_QKVF_WORDS;
    unsigned long long* attn_ll = gp.attn_ll + (size_t)chunk * GB_CTAS_MAX * BC * GB_STATE_WORDS;
    unsigned long long* ag_ll = gp.ag_ll + ((size_t)chunk * G2_GROUPS_MAX + gid) * BC * GB_AG_WORDS;
    unsigned long long* out_ll = gp.out_ll + (size_t)chunk * G2_GROUPS_MAX * hidden * BC;

    // ---- segments: the chunk's KV rows, concatenated request after request, are cut into G equal ranges (multiples of 16 rows).
    //      Lane b of warp 0 handles request b; offsets are warp prefix sums.  A request's new token is folded in by the rank
    //      that holds its last row ("owner"; for an empty request: the rank its offset falls into). ----
    if (warp == 0) {
        const int b = (int)lane;
        int len = 0, kb = 0, ns = 0, pos = 0;
        if (b < nb) {
            kb = p.indptr[b0 + b];
            const int end = p.indptr[b0 + b + 1] - 1;
            len = end - kb;
            ns = p.indices[end];
            pos = (int)p.positions[b0 + b];              // row of the RoPE table: read here, not behind exchange 1
        }
        int incl = len;
#pragma unroll
        for (int o = 1; o < BC; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if ((int)lane >= o) incl += v;
        }
        const int T = __shfl_sync(0xffffffffu, incl, BC - 1);
        const int off = incl - len;
        const int per = (((T + G - 1) / G) + ROWS512 - 1) & ~(ROWS512 - 1);
        const int c0 = min((int)rank * per, T), c1 = min(c0 + per, T);
        int s0 = max(c0, off) - off, s1 = min(c1, off + len) - off;
        const bool nonempty = b < nb && s1 > s0;
        const int owner = len > 0 ? (off + len - 1) / per : (per > 0 ? min(off / per, G - 1) : b % G);
        const int rf = len > 0 ? off / per : owner;
        const bool has = b < nb && (nonempty || owner == (int)rank);
        if (!nonempty) { s0 = 0; s1 = 0; }
        const uint32_t nt = has ? (uint32_t)((s1 - s0 + ROWS512 - 1) / ROWS512) : 0u;
        const unsigned bal = __ballot_sync(0xffffffffu, has);
        const int idx = __popc(bal & ((1u << lane) - 1u));
        uint32_t tincl = nt;
#pragma unroll
        for (int o = 1; o < BC; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, tincl, o);
            if ((int)lane >= o) tincl += v;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, tincl, BC - 1);
        const int nseg = __popc(bal);
        if (b < BC) { mreq[b * 4 + 0] = kb; mreq[b * 4 + 1] = pos; mreq[b * 4 + 2] = ns; mreq[b * 4 + 3] = rf | (owner << 16); }
        if ((int)lane >= nseg && lane <= BC) stile0[lane] = total;
        __syncwarp();
        if (has) {
            mseg[idx * 4 + 0] = b | ((owner == (int)rank) ? 256 : 0);
            mseg[idx * 4 + 1] = s0;
            mseg[idx * 4 + 2] = s1;
            stile0[idx] = tincl - nt;
        }
        if (lane == 0) mmisc[0] = nseg;
        __syncwarp();
    }
    if (lane == 0) {
        dsm::mbar_init(full_u32 + 8 * warp, 1);
        dsm::mbar_init(full_u32 + 8 * (warp + CONSUMER_WARPS), 1);
        if (tid == 0) {
            prefetch_tmap(&p.tm_wqkv);
            prefetch_tmap(&p.tm_wo);
        }
        dsm::mbar_fence_init();
    }
    __syncthreads();                                                          // meta visible to every warp

    const int n_seg = mmisc[0];
    const uint32_t n_qkv_tiles = (uint32_t)(2 * CONSUMER_WARPS * WPC);        // 24 row blocks x WPC column tiles
    const uint32_t n_kv_tiles = stile0[BC];
    const uint32_t n_o_tiles = (uint32_t)(CONSUMER_WARPS * 4 * ((RB_O + CONSUMER_WARPS - 1) / CONSUMER_WARPS));   // padded with empty tiles
    const uint32_t total_tiles = n_qkv_tiles + n_kv_tiles + n_o_tiles;

    CF_MARK(0);
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    const uint64_t pol = policy_evict_first();
    // The tensor maps over the pools were built from the host's copy of the pool addresses: they are used only if that copy
    // agrees with the device table.  Otherwise (no copy, or a stale one) the same stage layout is filled with 128-byte row
    // pieces that cannot be swizzled, and ldmatrix reads them linearly (bank conflicts, same results).
    const __half* kpool = reinterpret_cast<const __half*>(p.k_pool_ptrs[p.layer_id]);
    const __half* vpool = reinterpret_cast<const __half*>(p.v_pool_ptrs[p.layer_id]);
    const bool pool_maps = p.k_base != nullptr && kpool == p.k_base && vpool == p.v_base;
    const int swz = pool_maps ? 7 : 0;
    if (tid == 0 && pool_maps) { prefetch_tmap(&p.tm_k); prefetch_tmap(&p.tm_v); prefetch_tmap(&p.tm_kg); prefetch_tmap(&p.tm_vg); }

    // KV tile t (phase-local) -> segment, first row, end row, request
    auto seg_of = [&](uint32_t t) -> int {
        int s = 0;
#pragma unroll
        for (int q = 1; q < BC; ++q) s += (t >= stile0[q]) ? 1 : 0;
        return s;
    };
    int pre_slot0 = 0, pre_slot1 = 0;
    uint32_t pre_g0 = 0xffffffffu, pre_g1 = 0xffffffffu;
    // page index of this lane's row of KV tile g; rows past the end of the segment repeat its last row (tile::gather4 needs four
    // valid rows, the scores mask them)
    auto page_of = [&](uint32_t g) -> int {
        const uint32_t t = g - n_qkv_tiles;
        const int s = seg_of(t);
        const int b = mseg[s * 4] & 255;
        const int r = min(mseg[s * 4 + 1] + (int)(t - stile0[s]) * ROWS512 + (int)(lane & 15), mseg[s * 4 + 2] - 1);
        return p.indices[mreq[b * 4] + r];
    };

    auto issue_tile = [&](uint32_t g) {
        if (g >= total_tiles) return;
        const uint32_t s = ring_stage(g);
        const uint32_t fb = full_u32 + 8 * s;
        const uint32_t dst = smem_base + S::RING + s * STAGE_BYTES;
        if (g < n_qkv_tiles) {
            if (lane == 0) {
                const int k = (int)(g / CONSUMER_WARPS);
                const int rb = (int)(g % CONSUMER_WARPS) + CONSUMER_WARPS * (k / WPC), win = k % WPC;   // 32-row block of q | k | v
                int row0;
                if (rb < 16) row0 = qh0 * HEAD_DIM + rb * ROWS256;
                else if (rb < 20) row0 = Hq * HEAD_DIM + kvh * HEAD_DIM + (rb - 16) * ROWS256;
                else row0 = (Hq + Hkv) * HEAD_DIM + kvh * HEAD_DIM + (rb - 20) * ROWS256;
                dsm::mbar_arrive_expect_tx(fb, STAGE_BYTES);
                tma_load_2d(dst, &p.tm_wqkv, rank * KS + win * 128, row0, fb, pol);
                tma_load_2d(dst + 4096, &p.tm_wqkv, rank * KS + win * 128 + 64, row0, fb, pol);
            }
        } else if (g < n_qkv_tiles + n_kv_tiles) {
            const uint32_t t = g - n_qkv_tiles;
            const int sg = seg_of(t);
            const int r0 = mseg[sg * 4 + 1] + (int)(t - stile0[sg]) * ROWS512;
            const int nvalid = min(ROWS512, mseg[sg * 4 + 2] - r0);
            const bool odd = (g / CONSUMER_WARPS) & 1u;
            const int slot = (odd ? pre_g1 : pre_g0) == g ? (odd ? pre_slot1 : pre_slot0) : page_of(g);
            // stage = K dims 0-63 | K dims 64-127 | V dims 0-63 | V dims 64-127, each [16 rows][128 B] 128-byte swizzled
            const int slot0 = __shfl_sync(0xffffffffu, slot, 0);
            const bool run = nvalid == ROWS512 && __all_sync(0xffffffffu, slot == slot0 + (int)(lane & 15));
            if (!pool_maps) {                           // 16 (clamped) rows as 128-byte pieces straight from the pool
                if (lane == 0) dsm::mbar_arrive_expect_tx(fb, STAGE_BYTES);
                __syncwarp();
                const __half* src = (lane < 16 ? kpool : vpool) + (size_t)slot * kv_cols + kvh * HEAD_DIM;
                const uint32_t d = dst + (lane < 16 ? 0 : 4096) + (lane & 15) * 128;
                bulk_load_1d(d, src, 128, fb, pol);
                bulk_load_1d(d + 2048, src + 64, 128, fb, pol);
            } else if (run) {                                  // 16 consecutive slots: one tiled box per quarter
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
            if (lane == 0) {
                const uint32_t i = g - n_qkv_tiles - n_kv_tiles;
                const int k = (int)(i / CONSUMER_WARPS);
                const int rb = (int)(i % CONSUMER_WARPS) + CONSUMER_WARPS * (k / 4), win = k % 4;
                if (rb < RB_O) {
                    dsm::mbar_arrive_expect_tx(fb, STAGE_BYTES);
                    tma_load_2d(dst, &p.tm_wo, qh0 * HEAD_DIM + win * 128, rank * KS + rb * ROWS256, fb, pol);
                    tma_load_2d(dst + 4096, &p.tm_wo, qh0 * HEAD_DIM + win * 128 + 64, rank * KS + rb * ROWS256, fb, pol);
                } else {
                    dsm::mbar_arrive(fb);               // empty tile: the phase completes, nothing lands
                }
            }
        }
        const uint32_t g2 = g + NSTAGES;                // the tile that will live in this stage next
        if (g2 >= n_qkv_tiles && g2 < n_qkv_tiles + n_kv_tiles) {
            const int pg = page_of(g2);
            if ((g / CONSUMER_WARPS) & 1u) { pre_slot1 = pg; pre_g1 = g2; } else { pre_slot0 = pg; pre_g0 = g2; }
        }
    };

    CF_MARK(12);
    issue_tile(warp);
    issue_tile(warp + CONSUMER_WARPS);

    asm volatile("griddepcontrol.wait;" ::: "memory");

    const unsigned epoch = __ldcg(p.header);
    const unsigned flag = ll_flag_of_epoch(epoch);
    const bool residual_inplace = (static_cast<const void*>(p.residual_out) == static_cast<const void*>(p.residual_in));

    // ---- phase 0: fused residual add + RMSNorm, K-split.  This CTA holds columns [rank*KS, +KS) of the 8 requests; the sums
    //      of squares of a request are exchanged as one word per (rank, request). ----
    {
        BatchSlice<BC> slice;
        batch_slice_load<BC>(slice, p, b0, nb, hidden, KS, rank, tid);
        const int cpk = KS / 8;                                   // 8-element items per request (a multiple of 16)
#pragma unroll
        for (int it = 0; it < BatchSlice<BC>::ITEMS; ++it) {
            const int item = (int)tid + it * CONSUMER_THREADS;
            float f[8], r8[8];
            unpack8(slice.x[it], f);
            unpack8(slice.r[it], r8);
            float v = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) { const float h = round_h(f[k] + r8[k]); v = fmaf(h, h, v); }
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if ((lane & 15) == 0 && item < BC * cpk) sseg[item >> 4] = v;
        }
        dsm::named_bar_sync(CONSUMER_BAR, CONSUMER_THREADS);
        if (tid < (uint32_t)BC) {
            const int ns16 = cpk / 16;
            float a = 0.f;
            for (int k = 0; k < ns16; ++k) a += sseg[tid * ns16 + k];
            ll_store(ss_ll + (size_t)cta * BC + tid, a, flag);
        }
        for (int i = tid; i < G * BC; i += CONSUMER_THREADS) {
            const unsigned long long* src = ss_ll + ((size_t)gid * G + i / BC) * BC + (i % BC);
            ssall[i] = ll_resolve(src, ll_load(src), flag, p.header + 2);
        }
        dsm::named_bar_sync(CONSUMER_BAR, CONSUMER_THREADS);
        if (tid < (uint32_t)BC) {
            float tot = 0.f;
            for (int r = 0; r < G; ++r) tot += ssall[r * BC + tid];
            rstd_s[tid] = rsqrtf(tot / (float)hidden + p.eps);
        }
        dsm::named_bar_sync(CONSUMER_BAR, CONSUMER_THREADS);
#pragma unroll
        for (int it = 0; it < BatchSlice<BC>::ITEMS; ++it) {
            const int item = (int)tid + it * CONSUMER_THREADS;
            if (item >= BC * cpk) continue;
            const int b = item / cpk, e = (item % cpk) * 8;
            __align__(16) __half xn[8];
            if (b < nb) {
                const float rstd = rstd_s[b];
                float f[8], w8[8], r8[8];
                unpack8(slice.x[it], f);
                unpack8(slice.w[it], w8);
                unpack8(slice.r[it], r8);
                __align__(16) __half hs[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) { hs[k] = __float2half_rn(f[k] + r8[k]); f[k] = __half2float(hs[k]); }
                if (gid == 0 && !residual_inplace)
                    *reinterpret_cast<uint4*>(p.residual_out + (size_t)(b0 + b) * hidden + rank * KS + e) = *reinterpret_cast<const uint4*>(hs);
#pragma unroll
                for (int k = 0; k < 8; ++k) xn[k] = __float2half_rn(round_h(f[k] * rstd) * w8[k]);
            } else {
#pragma unroll
                for (int k = 0; k < 8; ++k) xn[k] = __float2half_rn(0.f);
            }
            *reinterpret_cast<uint4*>(xs + b * BK_XS_STRIDE + e) = *reinterpret_cast<const uint4*>(xn);
        }
        dsm::named_bar_sync(CONSUMER_BAR, CONSUMER_THREADS);
    }
    CF_MARK(1);

    uint32_t gbase = 0;
    const int g4 = lane >> 2, t4 = lane & 3;
    // ---- phase 1: QKV projection of the column slice on the tensor cores; a row block's partial sums leave as (value, epoch)
    //      words [row][request] as soon as its last column tile is done ----
    {
        float acc[2][4];
        for (uint32_t i = warp; i < n_qkv_tiles; i += CONSUMER_WARPS) {
            const uint32_t g = i, s = ring_stage(g);
            const int k = (int)(i / CONSUMER_WARPS);
            const int rb = (int)warp + CONSUMER_WARPS * (k / WPC), win = k % WPC;
            if (win == 0) {
#pragma unroll
                for (int mb = 0; mb < 2; ++mb) { acc[mb][0] = 0.f; acc[mb][1] = 0.f; acc[mb][2] = 0.f; acc[mb][3] = 0.f; }
            }
            uint32_t xb[8][2];
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
                const __half* xp = xs + g4 * BK_XS_STRIDE + win * 128 + ks * 16 + t4 * 2;
                xb[ks][0] = *reinterpret_cast<const uint32_t*>(xp);
                xb[ks][1] = *reinterpret_cast<const uint32_t*>(xp + 8);
            }
            ring_wait_full(full_u32, g);
            const uint32_t st = smem_base + S::RING + s * STAGE_BYTES;
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
#pragma unroll
                for (int mb = 0; mb < 2; ++mb) {
                    uint32_t af[4];
                    ldsm_a_mrows(af, st + (ks >> 2) * 4096, mb * 16, (ks & 3) * 2, lane);
                    mma16816(acc[mb], af, xb[ks][0], xb[ks][1]);
                }
            }
            __syncwarp();
            issue_tile(g + NSTAGES);
            if (win == WPC - 1) {
#pragma unroll
                for (int mb = 0; mb < 2; ++mb) {
                    unsigned long long* dstw = qkvp_ll + ((size_t)cta * S::R + rb * ROWS256 + mb * 16 + g4) * BC + 2 * t4;
                    ll_store2(dstw, acc[mb][0], acc[mb][1], flag);
                    ll_store2(dstw + 8 * BC, acc[mb][2], acc[mb][3], flag);
                }
            }
        }
        gbase += n_qkv_tiles;
    }
    dsm::named_bar_sync(CONSUMER_BAR, CONSUMER_THREADS);       // every warp is done with xs (the union changes role)
    CF_MARK(2);

    // ---- exchange 1 (reduce-scatter + RoPE): this CTA owns RoPE pairs i in [rank*PR, +PR) of the six 128-row slots ----
    {
        const int PR = 64 / G;                                   // pairs per slot (8 .. 1)
        const int NR = 12 * PR;                                  // rows of this CTA = 768 / G, local row lr = (slot*2 + half)*PR + u
        const int n_out = NR * BC;                               // sums this CTA owns (6144 / G)
        // all G * n_out = 6144 partial words as 3072 pairs of adjacent requests (16-byte loads: half the L2 requests), 8 per
        // thread, probed back to back, then resolved
        {
            unsigned long long w[8][2];
            const unsigned long long* src[8];
            const int n_pairs = n_out / 2;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int i = (int)tid + k * CONSUMER_THREADS;
                const int part = i / n_pairs, o2 = i % n_pairs, lr = o2 / (BC / 2), rqp = o2 % (BC / 2);
                const int sl2 = lr / PR, u = lr % PR;            // sl2 = slot * 2 + half
                const int row = (sl2 >> 1) * HEAD_DIM + (sl2 & 1) * 64 + (int)rank * PR + u;
                src[k] = qkvp_ll + (((size_t)gid * G + part) * S::R + row) * BC + 2 * rqp;
                ll_load2(src[k], w[k][0], w[k][1]);
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int i = (int)tid + k * CONSUMER_THREADS;
                stage1[2 * i] = ll_resolve(src[k], w[k][0], flag, p.header + 2);
                stage1[2 * i + 1] = ll_resolve(src[k] + 1, w[k][1], flag, p.header + 2);
            }
        }
        dsm::named_bar_sync(CONSUMER_BAR, CONSUMER_THREADS);
        for (int o = tid; o < n_out; o += CONSUMER_THREADS) {
            float a = 0.f;
            for (int part = 0; part < G; ++part) a += stage1[part * n_out + o];
            stage1[o] = round_h(a);                              // q / k / v leave the projection as fp16 (eager model)
        }
        dsm::named_bar_sync(CONSUMER_BAR, CONSUMER_THREADS);
        CF_MARK(3);
        // RoPE (NeoX) per pair; the pair (dim i, dim i + 64) travels in one word
        for (int it = tid; it < 6 * PR * BC; it += CONSUMER_THREADS) {
            const int rq = it % BC, u = (it / BC) % PR, sl = it / (BC * PR);
            const int i = (int)rank * PR + u;
            const float a = stage1[((sl * 2) * PR + u) * BC + rq];
            const float bv = stage1[((sl * 2 + 1) * PR + u) * BC + rq];
            __half lo = __float2half_rn(a), hi = __float2half_rn(bv);
            if (rq < nb) {
                if (sl < 5) {
                    const float* cosp = p.cos + (size_t)mreq[rq * 4 + 1] * HEAD_DIM;
                    const float c = cosp[i], sn = cosp[HEAD_DIM / 2 + i];
                    lo = __float2half_rn(fmaf(a, c, -bv * sn));
                    hi = __float2half_rn(fmaf(bv, c, a * sn));
                }
                if (sl >= 4 && writes_kv) {
                    __half* rowp = const_cast<__half*>(sl == 4 ? kpool : vpool) + (size_t)mreq[rq * 4 + 2] * kv_cols + kvh * HEAD_DIM;
                    rowp[i] = lo;
                    rowp[64 + i] = hi;
                }
            }
            const __half2 h2 = __halves2half2(lo, hi);
            ll_store(qkvf_ll + (size_t)rq * GB_QKVF_WORDS + sl * 64 + i, __uint_as_float(*reinterpret_cast<const uint32_t*>(&h2)), flag);
        }
        dsm::named_bar_sync(CONSUMER_BAR, CONSUMER_THREADS);   // stage1 is dead: the union becomes attn_part
    }
    CF_MARK(4);

    // ---- phase 2: flash-decode over this CTA's segments; all 12 warps work on one segment at a time ----
    for (int sg = 0; sg < n_seg; ++sg) {
        const int b = mseg[sg * 4] & 255;
        const bool owner = (mseg[sg * 4] & 256) != 0;
        const int row_begin = mseg[sg * 4 + 1], row_end = mseg[sg * 4 + 2];
        const uint32_t nt = stile0[sg + 1] - stile0[sg];
        const uint32_t gb = gbase + stile0[sg];
        // final q of the request (and k, v of the new token on its owner): one word per thread
        {
            const unsigned long long* src = qkvf_ll + (size_t)b * GB_QKVF_WORDS + tid;
            const bool need = tid < 256u || owner;
            if (need) {
                const uint32_t bits = __float_as_uint(ll_resolve(src, ll_load(src), flag, p.header + 2));
                const __half2 h2 = *reinterpret_cast<const __half2*>(&bits);
                const int sl = tid >> 6, i = tid & 63;
                qseg[sl * HEAD_DIM + i] = __low2half(h2);
                qseg[sl * HEAD_DIM + 64 + i] = __high2half(h2);
            }
        }
        dsm::named_bar_sync(CONSUMER_BAR, CONSUMER_THREADS);
        // tensor-core loop (see llama_decoder_gqa2_kernel.cuh): S = Q K^T with the 4 heads on rows 0-3 of M = 16, online softmax
        // on the C fragments, O += P V
        uint32_t qa[8][2];
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
#pragma unroll
            for (int hi = 0; hi < 2; ++hi)
                qa[ks][hi] = g4 < NQ ? *reinterpret_cast<const uint32_t*>(qseg + g4 * HEAD_DIM + ks * 16 + hi * 8 + t4 * 2) : 0u;
        }
        float oacc[16][4];
#pragma unroll
        for (int t = 0; t < 16; ++t) { oacc[t][0] = 0.f; oacc[t][1] = 0.f; oacc[t][2] = 0.f; oacc[t][3] = 0.f; }
        float mrun = -INFINITY, lrun = 0.f;
        const int lrow = lane & 7, lmat = lane >> 3;
        for (uint32_t i = first_tile(gb, warp); i < nt; i += CONSUMER_WARPS) {
            const uint32_t g = gb + i, s = ring_stage(g);
            ring_wait_full(full_u32, g);
            const uint32_t st = smem_base + S::RING + s * STAGE_BYTES;
            const int rows_left = row_end - (row_begin + (int)i * ROWS512);
            float sc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
                const int key = (lmat >> 1) * 8 + lrow, ch = (ks & 3) * 2 + (lmat & 1);
                const uint32_t addr = st + (ks >> 2) * 2048 + key * 128 + ((ch ^ (key & swz)) << 4);
                uint32_t k0, k1, k2, k3;
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                             : "=r"(k0), "=r"(k1), "=r"(k2), "=r"(k3) : "r"(addr));
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                             : "+f"(sc[0][0]), "+f"(sc[0][1]), "+f"(sc[0][2]), "+f"(sc[0][3])
                             : "r"(qa[ks][0]), "r"(0u), "r"(qa[ks][1]), "r"(0u), "r"(k0), "r"(k1));
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                             : "+f"(sc[1][0]), "+f"(sc[1][1]), "+f"(sc[1][2]), "+f"(sc[1][3])
                             : "r"(qa[ks][0]), "r"(0u), "r"(qa[ks][1]), "r"(0u), "r"(k2), "r"(k3));
            }
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
            for (int k = 0; k < 4; ++k) pr[k] = dsm::fast_exp2(s4[k] - m_use);
            const __half2 p01 = __floats2half2_rn(pr[0], pr[1]), p23 = __floats2half2_rn(pr[2], pr[3]);
            lrun = lrun * corr + (__low2float(p01) + __high2float(p01)) + (__low2float(p23) + __high2float(p23));
            const uint32_t pa0 = *reinterpret_cast<const uint32_t*>(&p01), pa2 = *reinterpret_cast<const uint32_t*>(&p23);
#pragma unroll
            for (int jd = 0; jd < 8; ++jd) {
                const int key = (lmat & 1) * 8 + lrow, ch = (jd & 3) * 2 + (lmat >> 1);
                const uint32_t addr = st + 4096 + (jd >> 2) * 2048 + key * 128 + ((ch ^ (key & swz)) << 4);
                uint32_t v0, v1, v2, v3;
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                             : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(addr));
                oacc[2 * jd][0] *= corr; oacc[2 * jd][1] *= corr;
                oacc[2 * jd + 1][0] *= corr; oacc[2 * jd + 1][1] *= corr;
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                             : "+f"(oacc[2 * jd][0]), "+f"(oacc[2 * jd][1]), "+f"(oacc[2 * jd][2]), "+f"(oacc[2 * jd][3])
                             : "r"(pa0), "r"(0u), "r"(pa2), "r"(0u), "r"(v0), "r"(v1));
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                             : "+f"(oacc[2 * jd + 1][0]), "+f"(oacc[2 * jd + 1][1]), "+f"(oacc[2 * jd + 1][2]), "+f"(oacc[2 * jd + 1][3])
                             : "r"(pa0), "r"(0u), "r"(pa2), "r"(0u), "r"(v2), "r"(v3));
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
        if (owner && warp < NQ) {                                // score of the new token against query head `warp`
            float a = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                a = fmaf(__half2float(qseg[warp * HEAD_DIM + lane * 4 + k]), __half2float(qseg[NQ * HEAD_DIM + lane * 4 + k]), a);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            if (lane == 0) red[warp] = a * kScaleLog2;
        }
        dsm::named_bar_sync(CONSUMER_BAR, CONSUMER_THREADS);
        // block merge in warp order (+ the new token on the owner), published straight from registers
        unsigned long long* stw = attn_ll + ((size_t)cta * BC + b) * GB_STATE_WORDS;
        for (int e = tid; e < NQ * HEAD_DIM; e += CONSUMER_THREADS) {
            const int h = e >> 7, d = e & 127;
            const float s_new = red[h];
            float M = owner ? s_new : -INFINITY;
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
            if (owner) {
                const float w = dsm::exp2_diff(s_new, M);
                L += w;
                O = fmaf(__half2float(qseg[(NQ + 1) * HEAD_DIM + d]), w, O);
            }
            ll_store(stw + h * S::PAY + 4 + d, O, flag);
            if (d == 0) { ll_store(stw + h * S::PAY, M, flag); ll_store(stw + h * S::PAY + 1, L, flag); }
        }
        dsm::named_bar_sync(CONSUMER_BAR, CONSUMER_THREADS);   // attn_part / qseg / red are free for the next segment
    }
    gbase += n_kv_tiles;
    CF_MARK(5);

    // ---- exchange 2: this CTA merges dims [rank*S2, +S2) of every request over the ranks that hold a segment of it (rank
    //      order), publishes the normalised fp16 pair; then every CTA reads the chunk's merged outputs ----
    {
        const int S2 = NQ * HEAD_DIM / G;                        // 64 .. 8, even
        const int hp = S2 / 2;
        for (int it = tid; it < nb * hp; it += CONSUMER_THREADS) {
            const int b = it / hp, d = (int)rank * S2 + 2 * (it % hp);
            const int h = d >> 7, dd = d & 127;
            const int rf = mreq[b * 4 + 3] & 0xffff, rl = mreq[b * 4 + 3] >> 16;
            float M = -INFINITY, L = 0.f, O0 = 0.f, O1 = 0.f;
            for (int r0 = rf; r0 <= rl; r0 += 2) {
                unsigned long long w[2][4];
                const unsigned long long* src[2];
                const bool two = r0 + 1 <= rl;
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    src[u] = attn_ll + (((size_t)gid * G + r0 + u) * BC + b) * GB_STATE_WORDS + h * S::PAY;
                    if (u == 0 || two) {
                        w[u][0] = ll_load(src[u]);
                        w[u][1] = ll_load(src[u] + 1);
                        ll_load2(src[u] + 4 + dd, w[u][2], w[u][3]);
                    }
                }
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    if (u == 1 && !two) break;
                    const float m = ll_resolve(src[u], w[u][0], flag, p.header + 2);
                    const float l = ll_resolve(src[u] + 1, w[u][1], flag, p.header + 2);
                    const float o0 = ll_resolve(src[u] + 4 + dd, w[u][2], flag, p.header + 2);
                    const float o1 = ll_resolve(src[u] + 4 + dd + 1, w[u][3], flag, p.header + 2);
                    const float Mn = fmaxf(M, m);
                    const float wa = dsm::exp2_diff(M, Mn), wb = dsm::exp2_diff(m, Mn);
                    L = L * wa + l * wb;
                    O0 = O0 * wa + o0 * wb;
                    O1 = O1 * wa + o1 * wb;
                    M = Mn;
                }
            }
            const __half2 h2 = __floats2half2_rn(O0 / L, O1 / L);     // attention output leaves as fp16 (eager model)
            ll_store(ag_ll + (size_t)b * GB_AG_WORDS + d / 2, __uint_as_float(*reinterpret_cast<const uint32_t*>(&h2)), flag);
        }
        CF_MARK(10);
        for (int i0 = tid; i0 < BC * GB_AG_WORDS / 2; i0 += CONSUMER_THREADS * 3) {     // pairs of words, 16-byte loads
            unsigned long long w[3][2];
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                const int i = 2 * (i0 + u * CONSUMER_THREADS);
                w[u][0] = 0ull; w[u][1] = 0ull;
                if (i < nb * GB_AG_WORDS) ll_load2(ag_ll + i, w[u][0], w[u][1]);
            }
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                const int i = 2 * (i0 + u * CONSUMER_THREADS);
                if (i >= BC * GB_AG_WORDS) continue;
                uint2 bits = make_uint2(0u, 0u);
                if (i < nb * GB_AG_WORDS) {
                    bits.x = __float_as_uint(ll_resolve(ag_ll + i, w[u][0], flag, p.header + 2));
                    bits.y = __float_as_uint(ll_resolve(ag_ll + i + 1, w[u][1], flag, p.header + 2));
                }
                *reinterpret_cast<uint2*>(ag16 + (i / GB_AG_WORDS) * S::AG_STRIDE + (i % GB_AG_WORDS) * 2) = bits;
            }
        }
        dsm::named_bar_sync(CONSUMER_BAR, CONSUMER_THREADS);
    }
    CF_MARK(6);

    // ---- phase 3: O projection on the tensor cores, one 32-row block per warp at a time over the group's 4 column tiles ----
    {
        float oc[2][4];
        for (uint32_t i = first_tile(gbase, warp); i < n_o_tiles; i += CONSUMER_WARPS) {
            const uint32_t g = gbase + i, s = ring_stage(g);
            const int k = (int)(i / CONSUMER_WARPS);
            const int rb = (int)(i % CONSUMER_WARPS) + CONSUMER_WARPS * (k / 4), win = k % 4;
            if (win == 0) {
#pragma unroll
                for (int mb = 0; mb < 2; ++mb) { oc[mb][0] = 0.f; oc[mb][1] = 0.f; oc[mb][2] = 0.f; oc[mb][3] = 0.f; }
            }
            ring_wait_full(full_u32, g);
            if (rb < RB_O) {
                uint32_t ob[8][2];
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    const __half* ap = ag16 + g4 * S::AG_STRIDE + win * 128 + ks * 16 + t4 * 2;
                    ob[ks][0] = *reinterpret_cast<const uint32_t*>(ap);
                    ob[ks][1] = *reinterpret_cast<const uint32_t*>(ap + 8);
                }
                const uint32_t st = smem_base + S::RING + s * STAGE_BYTES;
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
#pragma unroll
                    for (int mb = 0; mb < 2; ++mb) {
                        uint32_t af[4];
                        ldsm_a_mrows(af, st + (ks >> 2) * 4096, mb * 16, (ks & 3) * 2, lane);
                        mma16816(oc[mb], af, ob[ks][0], ob[ks][1]);
                    }
                }
            }
            __syncwarp();
            issue_tile(g + NSTAGES);
            if (rb < RB_O && win == 3) {
#pragma unroll
                for (int mb = 0; mb < 2; ++mb) {
                    unsigned long long* dstw = out_ll + ((size_t)gid * hidden + rank * KS + rb * ROWS256 + mb * 16 + g4) * BC + 2 * t4;
                    ll_store2(dstw, oc[mb][0], oc[mb][1], flag);
                    ll_store2(dstw + 8 * BC, oc[mb][2], oc[mb][3], flag);
                }
            }
        }
    }
    CF_MARK(7);

    // ---- cross-group reduction: the n_groups CTAs that share output rows [rank*KS, +KS) each finalise 1/n_groups of them, summing
    //      the groups' partials in group order ----
    {
        const int ng = gp.n_groups;
        const int lo = (int)((long long)gid * KS / ng), hi = (int)((long long)(gid + 1) * KS / ng);
        const int nrows = hi - lo;
        const bool fp32_out = p.flags & 1u;
        for (int it = tid; it < nrows * 4; it += CONSUMER_THREADS) {
            const int row = rank * KS + lo + it % nrows, pair = it / nrows;
            const unsigned long long* base = out_ll + ((size_t)row) * BC + 2 * pair;
            float a0 = 0.f, a1 = 0.f;
            for (int c = 0; c < ng; c += 8) {
                unsigned long long w[8][2];
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (c + j < ng) ll_load2(base + (size_t)(c + j) * hidden * BC, w[j][0], w[j][1]);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (c + j < ng) {
                        const unsigned long long* q = base + (size_t)(c + j) * hidden * BC;
                        a0 += ll_resolve(q, w[j][0], flag, p.header + 2);
                        a1 += ll_resolve(q + 1, w[j][1], flag, p.header + 2);
                    }
                }
            }
            const int bA = 2 * pair, bB = 2 * pair + 1;
            if (fp32_out) {
                float* o = static_cast<float*>(p.out);
                if (bA < nb) o[(size_t)(b0 + bA) * hidden + row] = a0;
                if (bB < nb) o[(size_t)(b0 + bB) * hidden + row] = a1;
            } else {
                __half* o = static_cast<__half*>(p.out);
                if (bA < nb) o[(size_t)(b0 + bA) * hidden + row] = __float2half_rn(a0);
                if (bB < nb) o[(size_t)(b0 + bB) * hidden + row] = __float2half_rn(a1);
            }
        }
    }
    CF_MARK(8);
    // A CTA that got here has seen every group's partial of its rows, and a group only gets past its exchanges once all of its
    // CTAs are past phase 0: CTA 0 of the chunk may overwrite `residual` (in-place form), and the launch's last chunk bumps the epoch.
    if (blockIdx.x == 0) {
        if (residual_inplace) {
            for (int e = tid * 8; e < nb * hidden; e += CONSUMER_THREADS * 8) {
                float f[8], r8[8];
                unpack8(*reinterpret_cast<const uint4*>(p.x + (size_t)b0 * hidden + e), f);
                unpack8(*reinterpret_cast<const uint4*>(p.residual_in + (size_t)b0 * hidden + e), r8);
                __align__(16) __half hs[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) hs[k] = __float2half_rn(f[k] + r8[k]);
                *reinterpret_cast<uint4*>(p.residual_out + (size_t)b0 * hidden + e) = *reinterpret_cast<const uint4*>(hs);
            }
        }
        if (tid == 0) {
            const unsigned prev = atomicAdd(p.header + 1, 1u);
            if (prev == gridDim.y - 1u) { p.header[1] = 0u; p.header[0] = epoch + 1u; }
        }
    }
    CF_MARK(9);
}

}  // namespace cfb
