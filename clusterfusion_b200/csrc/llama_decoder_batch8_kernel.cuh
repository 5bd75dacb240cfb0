/*
 * Batched paged decode, chunks of EIGHT requests per head cluster (second step).
 *
 * Same idea as llama_decoder_batch_kernel.cuh (one 4-CTA cluster per head, every Wqkv / Wo tile streamed once per chunk
 * and multiplied against the chunk's activation vectors with ldmatrix + mma.sync.m16n8k16), but the chunk fills the whole
 * N = 8 dimension of the MMA, so a batch of 8 streams the weights once instead of twice.  What makes 8 fit into 227 KB
 * next to the 192 KB tile ring is time-sharing one 19.5 KB region X:
 *
 *   QKV phase      X = the 8 activation vectors, fp16 [8][1024 + 8]
 *   exchange 1     X = reduce-scatter slots (fp32 partials written by the peers straight from their MMA accumulators with 8-byte
 *                  st.async, no staging copy) + all-gather buffers in fp16 (the folded values are fp16-exact): all 8 requests in
 *                  one pass (two passes of four until late in round 2), result kept as fp16 q|k|v [8][384] outside X
 *   attention      over this CTA's KV segments (batch_build_segments), then ONE all-gather of the 8 block-merged states through X
 *   O phase        X = a 1 KB staging tile per warp: the [32 rows x 8 requests] C fragments are transposed through it
 *                  and leave as red.global.add.v4 straight away (no 32 KB out_part; the reds overlap the stream)
 *
 * X is written by peer CTAs (st.async) in the exchange phases, so every change of its role is fenced by a cluster barrier
 * (arrive as soon as this CTA is done with the old role, wait right before the first remote write of the new one): B0
 * at the start (peers have armed their exchange barriers; waited on before the sums-of-squares exchange), B1 after the QKV
 * loop, B2 after RoPE.  PAGED variant, MHA, hidden <= 4096.
 */
#pragma once

#include "llama_decoder_batch_kernel.cuh"

namespace cfb {

struct SmemB8 {
    static constexpr int BC = 8, HB = 4, CL = 4;                          // chunk, half chunk, cluster
    static constexpr int QKV_OUT = 3 * HEAD_DIM;                          // 384
    static constexpr int PAY = HEAD_DIM + 4;
    static constexpr int RING = 0;
    static constexpr int ATTN_PART = RING + NSTAGES * STAGE_BYTES;        // fp32 [12 warps][132]
    static constexpr int X = ATTN_PART + CONSUMER_WARPS * PAY * 4;
    // role 1: activations
    static constexpr int XS = X;                                          // fp16 [8][BK_XS_STRIDE]
    // role 2: exchange 1, all 8 requests in one pass: rank r reduces requests 2r and 2r + 1
    static constexpr int SLICE1 = 2 * QKV_OUT;                            // 768 values per rank
    static constexpr int RS_RECV = X;                                     // fp32 [4 source ranks][384 rows][2 requests]
    static constexpr int RED1 = RS_RECV + CL * SLICE1 * 4;                // fp16 [2][384]   my two folded requests
    static constexpr int AG_RECV = RED1 + SLICE1 * 2;                     // fp16 [4 ranks][2][384] = [8][384]
    static constexpr int X_BYTES = AG_RECV + CL * SLICE1 * 2 - X;         // 19968
    static_assert(BC * BK_XS_STRIDE * 2 <= X_BYTES, "activations must fit into X");
    // role 3: exchange 2, all 8 requests in one pass: slot [rank] of ATTN_RECV is also the source of this CTA's contribution
    static constexpr int ATTN_RECV = X;                                   // fp32 [4 ranks][8][132]
    static_assert(CL * BC * PAY * 4 <= X_BYTES, "exchange-2 buffers must fit into X");
    // role 4: O-phase staging
    static constexpr int OSTAGE = X;                                      // fp32 [12 warps][8][32]
    static_assert(CONSUMER_WARPS * BC * 32 * 4 <= X_BYTES, "staging must fit into X");
    static constexpr int QKV_FIN = X + X_BYTES;                           // fp16 [8][384]  roped q (unscaled) | k | v
    static constexpr int ATTN_OUT = QKV_FIN + BC * QKV_OUT * 2;           // fp16 [8][128]
    static constexpr int RED = ATTN_OUT + BC * HEAD_DIM * 2;              // fp32 [12][8] + [8]
    static constexpr int META = RED + (CONSUMER_WARPS + 1) * BC * 4;      // int [8][4] requests: kv_base, position, new_slot, owner; int [8][4] segments:
                                                                          // request | owner << 8, row begin, row end, -; u32 [9] tile0; [1] n_seg
    static constexpr int BARS = META + 76 * 4;                            // u64 full[NSTAGES], xbar[7]
    static constexpr int FLAGS = BARS + (NSTAGES + 7) * 8;                // u32 [8]
    static constexpr int TOTAL = FLAGS + BC * 4;
    static_assert(BARS % 8 == 0, "mbarrier alignment");
    static_assert(TOTAL <= 227 * 1024, "shared-memory layout exceeds the 227 KB opt-in limit");
};

__global__ void __launch_bounds__(BLOCK_THREADS, 1)
llama_decoder_layer_batch8_kernel(const __grid_constant__ KParams p)
{
    using S = SmemB8;
    constexpr int BC = S::BC, HB = S::HB, CLUSTER = S::CL;
    constexpr float kScaleLog2 = 0.08838834764831845f * 1.4426950408889634f;

    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t smem_base = dsm::smem_u32(smem);
    const uint32_t tid = threadIdx.x;
    const uint32_t warp = tid >> 5;
    const uint32_t lane = tid & 31;
    const uint32_t rank = dsm::cluster_ctarank();
    const uint32_t head = blockIdx.x / CLUSTER;
    const int b0 = blockIdx.y * BC;
    const int nb = min(BC, p.batch - b0);

    const int hidden = p.hidden;
    const int KS = hidden / CLUSTER;
    const int kv_cols = p.n_kv_heads * HEAD_DIM;

    const uint32_t full_u32 = smem_base + S::BARS;
    const uint32_t xbar_u32 = full_u32 + NSTAGES * 8;

    __half* xs = reinterpret_cast<__half*>(smem + S::XS);
    float* attn_part = reinterpret_cast<float*>(smem + S::ATTN_PART);
    float* rs_recv = reinterpret_cast<float*>(smem + S::RS_RECV);
    __half* red1 = reinterpret_cast<__half*>(smem + S::RED1);
    __half* ag_recv = reinterpret_cast<__half*>(smem + S::AG_RECV);
    __half* qkv_fin = reinterpret_cast<__half*>(smem + S::QKV_FIN);
    float* attn_recv = reinterpret_cast<float*>(smem + S::ATTN_RECV);
    __half* attn_out = reinterpret_cast<__half*>(smem + S::ATTN_OUT);
    float* ostage = reinterpret_cast<float*>(smem + S::OSTAGE);
    float* red = reinterpret_cast<float*>(smem + S::RED);
    int* meta = reinterpret_cast<int*>(smem + S::META);                       // [b][4]
    int* mseg = meta + BC * 4;                                                // [s][4]
    uint32_t* tile0 = reinterpret_cast<uint32_t*>(mseg + BC * 4);             // [9]: first KV tile of segment s (entries >= n_seg: total)
    int* mmisc = reinterpret_cast<int*>(tile0 + BC + 1);                      // [0] n_seg
    uint32_t* sflags = reinterpret_cast<uint32_t*>(smem + S::FLAGS);

    // ---- KV segments of this CTA (batch_build_segments, llama_decoder_batch_kernel.cuh): all 12 warps work on the same request at a
    //      time -- one copy of the KV loop and one block merge per segment instead of per-request register states in eight
    //      unrolled copies (207 KB of code, 64 % instruction-cache hit rate, 16 us per half chunk against 9.5 us for the same rows in
    //      the chunks-of-4 kernel of that time) ----
    if (warp == 0) batch_build_segments<BC>(p, b0, nb, (int)rank, CLUSTER, lane, meta, mseg, tile0, mmisc);
    if (lane == 0) {
        dsm::mbar_init(full_u32 + 8 * warp, 1);
        dsm::mbar_init(full_u32 + 8 * (warp + CONSUMER_WARPS), 1);
        if (tid == 0) {
            prefetch_tmap(&p.tm_wqkv);
            prefetch_tmap(&p.tm_wo);
            cluster_reduce_arm<CLUSTER>(xbar_u32, S::SLICE1 * 4);            // scatter: 768 fp32 partials from every peer
            cluster_reduce_arm<CLUSTER>(xbar_u32 + 16, S::SLICE1 * 2);       // gather: 768 fp16 results from every peer
            cluster_reduce_arm<CLUSTER>(xbar_u32 + 32, BC * S::PAY * 4);     // softmax states of the 8 requests
            cluster_reduce_arm<CLUSTER>(xbar_u32 + 48, BC * 4);              // sums of squares
        }
        dsm::mbar_fence_init();
    }
    __syncthreads();                                                          // meta visible to every warp

    const uint32_t n_qkv_tiles = (uint32_t)CONSUMER_WARPS * (KS / 128);
    const uint32_t n_kv_tiles = tile0[BC];
    const uint32_t n_o_tiles = (uint32_t)(KS / ROWS256);
    const uint32_t total_tiles = n_qkv_tiles + n_kv_tiles + n_o_tiles;

    CF_MARK(0);
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    const uint64_t pol = policy_evict_first();
    const __half* kpool = reinterpret_cast<const __half*>(p.k_pool_ptrs[p.layer_id]);
    const __half* vpool = reinterpret_cast<const __half*>(p.v_pool_ptrs[p.layer_id]);
    // host's copy of the pool addresses (tensor maps over the pools): used only if it agrees with the device table
    const bool pool_maps = p.k_base != nullptr && kpool == p.k_base && vpool == p.v_base;
    if (tid == 0 && pool_maps) { prefetch_tmap(&p.tm_k); prefetch_tmap(&p.tm_v); prefetch_tmap(&p.tm_kg); prefetch_tmap(&p.tm_vg); }

    const int n_seg = mmisc[0];
    auto seg_of = [&](uint32_t t) -> int {                // which segment KV tile t (phase-local index) belongs to
        int sgi = 0;
#pragma unroll
        for (int q = 1; q < BC; ++q) sgi += (t >= tile0[q]) ? 1 : 0;
        return sgi;
    };
    int pre_slot0 = 0, pre_slot1 = 0;
    uint32_t pre_g0 = 0xffffffffu, pre_g1 = 0xffffffffu;
    auto page_of = [&](uint32_t g) -> int {               // page index of this lane's row of KV tile g
        const uint32_t t = g - n_qkv_tiles;
        const int sgi = seg_of(t);
        const int r = mseg[sgi * 4 + 1] + (int)(t - tile0[sgi]) * ROWS512 + (int)(lane & 15);
        return (r < mseg[sgi * 4 + 2]) ? p.indices[meta[(mseg[sgi * 4] & 255) * 4] + r] : 0;
    };
    auto issue_tile = [&](uint32_t g) {
        if (g >= total_tiles) return;
        const uint32_t s = ring_stage(g);
        const uint32_t fb = full_u32 + 8 * s;
        const uint32_t dst = smem_base + S::RING + s * STAGE_BYTES;
        if (g < n_qkv_tiles) {
            if (lane == 0) {
                const int rb = (int)(g % CONSUMER_WARPS), ct = (int)(g / CONSUMER_WARPS);
                const int j = rb / 4, sub = rb % 4;
                const int row0 = (j == 0) ? head * HEAD_DIM
                               : (j == 1) ? p.n_heads * HEAD_DIM + head * HEAD_DIM
                                          : (p.n_heads + p.n_kv_heads) * HEAD_DIM + head * HEAD_DIM;
                dsm::mbar_arrive_expect_tx(fb, STAGE_BYTES);
                tma_load_2d(dst, &p.tm_wqkv, rank * KS + ct * 128, row0 + sub * 32, fb, pol);
                tma_load_2d(dst + 4096, &p.tm_wqkv, rank * KS + ct * 128 + 64, row0 + sub * 32, fb, pol);
            }
        } else if (g < n_qkv_tiles + n_kv_tiles) {
            const uint32_t t = g - n_qkv_tiles;
            const int sgi = seg_of(t);
            const int rbeg = mseg[sgi * 4 + 1], rend = mseg[sgi * 4 + 2];
            const int i = (int)(t - tile0[sgi]);
            const bool odd = (g / CONSUMER_WARPS) & 1u;
            const long long slot = (odd ? pre_g1 : pre_g0) == g ? (long long)(odd ? pre_slot1 : pre_slot0) : (long long)page_of(g);
            const int nvalid = min(ROWS512, rend - (rbeg + i * ROWS512));
            issue_kv_stage(p, pool_maps, false, dst, fb, head * HEAD_DIM, slot, nvalid, kpool, vpool, kv_cols, lane, pol);
        } else {
            if (lane == 0) {
                const uint32_t i = g - n_qkv_tiles - n_kv_tiles;
                dsm::mbar_arrive_expect_tx(fb, STAGE_BYTES);
                tma_load_2d(dst, &p.tm_wo, head * HEAD_DIM, rank * KS + i * ROWS256, fb, pol);
                tma_load_2d(dst + 4096, &p.tm_wo, head * HEAD_DIM + 64, rank * KS + i * ROWS256, fb, pol);
            }
        }
        const uint32_t g2 = g + NSTAGES;
        if (g2 >= n_qkv_tiles && g2 < n_qkv_tiles + n_kv_tiles) {
            const int pg = page_of(g2);
            if ((g / CONSUMER_WARPS) & 1u) { pre_slot1 = pg; pre_g1 = g2; } else { pre_slot0 = pg; pre_g0 = g2; }
        }
    };

    CF_MARK(12);
    issue_tile(warp);
    issue_tile(warp + CONSUMER_WARPS);
    dsm::cluster_arrive();                                                    // B0: barriers armed

    asm volatile("griddepcontrol.wait;" ::: "memory");

    // ---- phase 0: fused residual add + RMSNorm for the 8 requests, K-split over the cluster -----------------------------
    dsm::cluster_wait();                                                      // B0: every peer has armed its exchange barriers
    batch_rmsnorm_slice<BC, CLUSTER>(p, b0, nb, hidden, KS, rank, head, tid, red, smem_base + S::RED, xbar_u32 + 48, xs);
    CF_MARK(1);

    uint32_t gbase = 0;
    const int g4 = lane >> 2, t4 = lane & 3;
    // ---- phase 1: QKV GEMV on the tensor cores, all 8 requests on the N dimension ------------------------------------
    float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    for (uint32_t i = warp; i < n_qkv_tiles; i += CONSUMER_WARPS) {
        const uint32_t g = i, s = ring_stage(g);
        const int ct = (int)(i / CONSUMER_WARPS);
        uint32_t xb[8][2];
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
            const __half* xp = xs + g4 * BK_XS_STRIDE + ct * 128 + ks * 16 + t4 * 2;
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
    }
    gbase += n_qkv_tiles;
    dsm::named_bar_sync(CONSUMER_BAR, CONSUMER_THREADS);       // every warp is done reading xs
    CF_MARK(2);
    dsm::cluster_arrive();                                      // B1: this CTA no longer reads xs (X changes role)
    dsm::cluster_wait();

    // ---- exchange 1 + RoPE, all 8 requests in one pass.  A lane's C fragments hold requests 2*t4 and 2*t4 + 1 of rows g4 / g4 + 8:
    //      both belong to rank t4, so the fragments go straight from registers into that rank's receive slots as 8-byte
    //      st.async (no staging copy: the fp32 staging of 8 requests would not fit next to the receive buffers) ----------------
    {
        const uint32_t dst_rank = (uint32_t)t4;
        const uint32_t slot_u32 = smem_base + S::RS_RECV + rank * (S::SLICE1 * 4);      // my slot in the destination's buffer
#pragma unroll
        for (int mb = 0; mb < 2; ++mb) {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int row = (int)warp * 32 + mb * 16 + g4 + hh * 8;
                const float2 v = make_float2(acc[mb][2 * hh], acc[mb][2 * hh + 1]);
                if (dst_rank == rank) *reinterpret_cast<float2*>(rs_recv + rank * S::SLICE1 + row * 2) = v;
                else dsm::st_async_v2(dsm::mapa(slot_u32 + row * 8, dst_rank), v, dsm::mapa(xbar_u32, dst_rank));
            }
        }
        dsm::mbar_wait_cluster(xbar_u32, 0);
        dsm::named_bar_sync(CONSUMER_BAR, CONSUMER_THREADS);
        // fold my two requests in rank order; q / k / v leave the projection as fp16 (eager model)
        for (int e = tid; e < S::SLICE1; e += CONSUMER_THREADS) {
            const int rl = e / S::QKV_OUT, row = e % S::QKV_OUT;
            float a = 0.f;
#pragma unroll
            for (int r = 0; r < CLUSTER; ++r) a += rs_recv[r * S::SLICE1 + row * 2 + rl];
            red1[e] = __float2half_rn(a);
        }
        uint32_t ph_g = 0;
        cluster_reduce<CLUSTER, Stage::QUK_DEEPSEEK, CONSUMER_THREADS, CONSUMER_BAR>(
            S::SLICE1 * 2, tid, S::SLICE1, rank, smem_base + S::RED1, smem_base + S::AG_RECV, xbar_u32 + 16, ph_g,
            reinterpret_cast<float*>(red1), reinterpret_cast<float*>(ag_recv));
        // RoPE (NeoX) for the 8 requests (ag_recv = [8][384]); new K / V rows into the pool; q stays unscaled fp16
        for (int f = tid; f < BC * S::QKV_OUT; f += CONSUMER_THREADS) {
            const int b = f / S::QKV_OUT, e = f % S::QKV_OUT;
            const int which = e >> 7, d = e & 127;
            __half outv = ag_recv[f];
            if (b < nb) {
                if (which < 2) {
                    const float a = __half2float(outv);
                    const float* cosp = p.cos + (size_t)meta[b * 4 + 1] * HEAD_DIM;
                    const float* sinp = cosp + HEAD_DIM / 2;
                    const float bbv = __half2float(ag_recv[f ^ 64]);
                    const int i = d & 63;
                    const float rot = (d & 64) ? fmaf(a, cosp[i], bbv * sinp[i]) : fmaf(a, cosp[i], -bbv * sinp[i]);
                    outv = __float2half_rn(rot);
                    if (which == 1 && rank == 0) {
                        __half* kp = reinterpret_cast<__half*>(p.k_pool_ptrs[p.layer_id]);
                        kp[(size_t)meta[b * 4 + 2] * kv_cols + head * HEAD_DIM + d] = outv;
                    }
                } else if (rank == 0) {
                    __half* vp = reinterpret_cast<__half*>(p.v_pool_ptrs[p.layer_id]);
                    vp[(size_t)meta[b * 4 + 2] * kv_cols + head * HEAD_DIM + d] = outv;
                }
            }
            qkv_fin[b * S::QKV_OUT + e] = outv;
        }
        dsm::named_bar_sync(CONSUMER_BAR, CONSUMER_THREADS);
    }
    CF_MARK(4);
    dsm::cluster_arrive();                                      // B2: done reading ag_recv (X changes role again)

    // ---- phase 2: flash-decode over this CTA's segments (all 12 warps on one request at a time), then the softmax-state
    //      exchange, one half chunk at a time ------------------------------------------------------------------------------
    {
        const int sub = lane >> 4, c = lane & 15;
        // block-merged state of request b on this CTA: thread tid < 128 keeps o[dim tid] in registers, (m, l) sit in red[2b], red[2b+1]
        // (the RMSNorm partials there are dead); requests without a segment here stay (-inf, 0, 0)
        float Ov[BC];
#pragma unroll
        for (int b = 0; b < BC; ++b) Ov[b] = 0.f;
        if (tid == 0) {                                         // (written and read by thread 0 only: no barrier needed when n_seg == 0)
#pragma unroll
            for (int b = 0; b < BC; ++b) { red[2 * b] = -INFINITY; red[2 * b + 1] = 0.f; }
        }
        for (int sg = 0; sg < n_seg; ++sg) {
            const int b = mseg[sg * 4] & 255;
            const bool owner = (mseg[sg * 4] & 256) != 0;
            const int rbeg = mseg[sg * 4 + 1], rend = mseg[sg * 4 + 2];
            const uint32_t nt = tile0[sg + 1] - tile0[sg];
            const uint32_t gb = gbase + tile0[sg];
            float m = -INFINITY, l = 0.f, o8[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) o8[k] = 0.f;
            float q8[8];
            unpack8(*reinterpret_cast<const uint4*>(qkv_fin + b * S::QKV_OUT + c * 8), q8);
#pragma unroll
            for (int k = 0; k < 8; ++k) q8[k] *= kScaleLog2;
            for (uint32_t i = first_tile(gb, warp); i < nt; i += CONSUMER_WARPS) {
                const uint32_t g = gb + i, s = ring_stage(g);
                ring_wait_full(full_u32, g);
                const uint4* kt = reinterpret_cast<const uint4*>(smem + S::RING + s * STAGE_BYTES);
                const uint4* vt = kt + STAGE_BYTES / 32;
                const int rows_left = rend - (rbeg + (int)i * ROWS512);
                float sc[ROWS512 / 2];
#pragma unroll
                for (int jj = 0; jj < ROWS512 / 2; ++jj) {
                    const int row = 2 * jj + sub;
                    float k8[8];
                    unpack8(kt[row * 16 + c], k8);
                    float a = 0.f;
#pragma unroll
                    for (int k = 0; k < 8; ++k) a = fmaf(q8[k], k8[k], a);
                    a += __shfl_xor_sync(0xffffffffu, a, 1);
                    a += __shfl_xor_sync(0xffffffffu, a, 2);
                    a += __shfl_xor_sync(0xffffffffu, a, 4);
                    a += __shfl_xor_sync(0xffffffffu, a, 8);
                    sc[jj] = (row < rows_left) ? a : -INFINITY;
                }
                float mx = sc[0];
#pragma unroll
                for (int jj = 1; jj < ROWS512 / 2; ++jj) mx = fmaxf(mx, sc[jj]);
                const float m_new = fmaxf(m, mx);
                const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
                const float corr = dsm::exp2_diff(m, m_use);
                l *= corr;
#pragma unroll
                for (int k = 0; k < 8; ++k) o8[k] *= corr;
#pragma unroll
                for (int jj = 0; jj < ROWS512 / 2; ++jj) {
                    const int row = 2 * jj + sub;
                    const float pr = dsm::fast_exp2(sc[jj] - m_use);
                    l += pr;
                    uint4 raw = vt[row * 16 + c];
                    if (row >= rows_left) raw = make_uint4(0, 0, 0, 0);
                    float v8[8];
                    unpack8(raw, v8);
#pragma unroll
                    for (int k = 0; k < 8; ++k) o8[k] = fmaf(pr, v8[k], o8[k]);
                }
                m = m_new;
                __syncwarp();
                issue_tile(g + NSTAGES);
            }
            // merge the two half-warps in registers, then the block merge through [12][132]
            {
                const float m2 = __shfl_xor_sync(0xffffffffu, m, 16);
                const float l2 = __shfl_xor_sync(0xffffffffu, l, 16);
                const float M = fmaxf(m, m2);
                const float w1 = dsm::exp2_diff(m, M), w2 = dsm::exp2_diff(m2, M);
                l = l * w1 + l2 * w2;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float o2 = __shfl_xor_sync(0xffffffffu, o8[k], 16);
                    o8[k] = o8[k] * w1 + o2 * w2;
                }
                m = M;
            }
            if (sub == 0) {
                float* slot = attn_part + warp * S::PAY;
                if (c == 0) { slot[0] = m; slot[1] = l; }
                *reinterpret_cast<float4*>(slot + 4 + c * 8) = make_float4(o8[0], o8[1], o8[2], o8[3]);
                *reinterpret_cast<float4*>(slot + 4 + c * 8 + 4) = make_float4(o8[4], o8[5], o8[6], o8[7]);
            }
            if (owner && warp == 0) {                           // score of the request's new token
                float a = 0.f;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    a = fmaf(__half2float(qkv_fin[b * S::QKV_OUT + lane * 4 + k]),
                             __half2float(qkv_fin[b * S::QKV_OUT + HEAD_DIM + lane * 4 + k]), a);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
                if (lane == 0) red[CONSUMER_WARPS * BC + b] = a * kScaleLog2;
            }
            dsm::named_bar_sync(CONSUMER_BAR, CONSUMER_THREADS);
            if (tid < HEAD_DIM) {
                const float s_new = red[CONSUMER_WARPS * BC + b];
                float M = owner ? s_new : -INFINITY;
#pragma unroll
                for (int gI = 0; gI < CONSUMER_WARPS; ++gI) M = fmaxf(M, attn_part[gI * S::PAY]);
                float L = 0.f, Ovv = 0.f;
#pragma unroll
                for (int gI = 0; gI < CONSUMER_WARPS; ++gI) {
                    const float w = dsm::exp2_diff(attn_part[gI * S::PAY], M);
                    L = fmaf(attn_part[gI * S::PAY + 1], w, L);
                    Ovv = fmaf(attn_part[gI * S::PAY + 4 + tid], w, Ovv);
                }
                if (owner) {
                    const float w = dsm::exp2_diff(s_new, M);
                    L += w;
                    Ovv = fmaf(__half2float(qkv_fin[b * S::QKV_OUT + 2 * HEAD_DIM + tid]), w, Ovv);
                }
#pragma unroll
                for (int q = 0; q < BC; ++q) Ov[q] = (q == b) ? Ovv : Ov[q];
                if (tid == 0) { red[2 * b] = M; red[2 * b + 1] = L; }
            }
            dsm::named_bar_sync(CONSUMER_BAR, CONSUMER_THREADS);
        }
        gbase += n_kv_tiles;
        CF_MARK(5);
        // ---- exchange 2: all-gather of the 8 block-merged states (this CTA's go into its own slot of the receive buffer, which
        //      is also the source of the pushes), merged in rank order ----
        dsm::cluster_wait();                                    // B2: every peer is past its RoPE, X may take exchange-2 data
        float* mine = attn_recv + rank * (BC * S::PAY);
        if (tid < HEAD_DIM) {
#pragma unroll
            for (int b = 0; b < BC; ++b) {
                float* stp = mine + b * S::PAY;
                stp[4 + tid] = Ov[b];
                if (tid == 0) { stp[0] = red[2 * b]; stp[1] = red[2 * b + 1]; stp[2] = 0.f; stp[3] = 0.f; }
            }
        }
        CF_MARK(3);
        uint32_t ph_a = 0;
        cluster_reduce<CLUSTER, Stage::QUK_DEEPSEEK, CONSUMER_THREADS, CONSUMER_BAR>(
            BC * S::PAY * 4, tid, BC * S::PAY, rank, smem_base + S::ATTN_RECV + rank * (BC * S::PAY * 4), smem_base + S::ATTN_RECV,
            xbar_u32 + 32, ph_a, mine, attn_recv);
        for (int f = tid; f < BC * HEAD_DIM; f += CONSUMER_THREADS) {
            const int b = f >> 7, d = f & 127;
            float M = -INFINITY;
#pragma unroll
            for (int r = 0; r < CLUSTER; ++r) M = fmaxf(M, attn_recv[(r * BC + b) * S::PAY]);
            float L = 0.f, Ovv = 0.f;
#pragma unroll
            for (int r = 0; r < CLUSTER; ++r) {
                const float* stp = attn_recv + (r * BC + b) * S::PAY;
                const float w = dsm::exp2_diff(stp[0], M);
                L = fmaf(stp[1], w, L);
                Ovv = fmaf(stp[4 + d], w, Ovv);
            }
            attn_out[b * HEAD_DIM + d] = __float2half_rn((b < nb) ? Ovv / L : 0.f);   // fp16, as the eager model
        }
        dsm::named_bar_sync(CONSUMER_BAR, CONSUMER_THREADS);
    }
    CF_MARK(6);

    // ---- phase 3: O GEMV on the tensor cores, C fragments leave through a per-warp staging tile as red.v4 -------------
    {
        uint32_t ob[8][2];
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
            const __half* ap = attn_out + g4 * HEAD_DIM + ks * 16 + t4 * 2;
            ob[ks][0] = *reinterpret_cast<const uint32_t*>(ap);
            ob[ks][1] = *reinterpret_cast<const uint32_t*>(ap + 8);
        }
        float* stg = ostage + warp * (BC * 32);
        for (uint32_t i = first_tile(gbase, warp); i < n_o_tiles; i += CONSUMER_WARPS) {
            const uint32_t g = gbase + i, s = ring_stage(g);
            ring_wait_full(full_u32, g);
            const uint32_t st = smem_base + S::RING + s * STAGE_BYTES;
            float oc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
#pragma unroll
                for (int mb = 0; mb < 2; ++mb) {
                    uint32_t af[4];
                    ldsm_a_mrows(af, st + (ks >> 2) * 4096, mb * 16, (ks & 3) * 2, lane);
                    mma16816(oc[mb], af, ob[ks][0], ob[ks][1]);
                }
            }
            __syncwarp();
            issue_tile(g + NSTAGES);
            // transpose through the staging tile: stg[request][row]
#pragma unroll
            for (int mb = 0; mb < 2; ++mb) {
                stg[(2 * t4) * 32 + mb * 16 + g4] = oc[mb][0];
                stg[(2 * t4 + 1) * 32 + mb * 16 + g4] = oc[mb][1];
                stg[(2 * t4) * 32 + mb * 16 + g4 + 8] = oc[mb][2];
                stg[(2 * t4 + 1) * 32 + mb * 16 + g4 + 8] = oc[mb][3];
            }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int f = lane + 32 * j, n = f >> 3, r4 = f & 7;           // request n, rows r4*4 .. +4
                if (n < nb)
                    red_add_v4(p.scratch + (size_t)(b0 + n) * hidden + rank * KS + i * ROWS256 + r4 * 4,
                               *reinterpret_cast<const float4*>(stg + n * 32 + r4 * 4));
            }
            __syncwarp();
        }
    }
    CF_MARK(7);

    // ---- last arriver of each (request, slice) finalises ----------------------------------------------------------------
    __threadfence();
    dsm::named_bar_sync(CONSUMER_BAR, CONSUMER_THREADS);
    if (tid < (uint32_t)nb) {
        unsigned* counters = p.counters + (size_t)(b0 + tid) * (CLUSTER + 1);
        const unsigned prev = atomicAdd(&counters[rank], 1u);
        sflags[tid] = (prev == (unsigned)p.n_heads - 1u);
    }
    dsm::named_bar_sync(CONSUMER_BAR, CONSUMER_THREADS);
    CF_MARK(8);
    {
        __threadfence();
        const bool fp32_out = p.flags & 1u;
        const int e = tid * 4;                                // KS <= 1024: one float4 per thread
        float4 v[BC];                                         // all eight loads go out before anything is stored: one L2 round trip
#pragma unroll
        for (int b = 0; b < BC; ++b)
            if (b < nb && sflags[b] && e < KS) v[b] = ld_cg_v4(p.scratch + (size_t)(b0 + b) * hidden + rank * KS + e);
#pragma unroll
        for (int b = 0; b < BC; ++b) {
            if (b < nb && sflags[b] && e < KS) {
                const size_t off = (size_t)(b0 + b) * hidden + rank * KS + e;
                *reinterpret_cast<float4*>(p.scratch + off) = make_float4(0.f, 0.f, 0.f, 0.f);
                if (fp32_out) {
                    *reinterpret_cast<float4*>(static_cast<float*>(p.out) + off) = v[b];
                } else {
                    __align__(8) __half h4[4] = {__float2half_rn(v[b].x), __float2half_rn(v[b].y),
                                                 __float2half_rn(v[b].z), __float2half_rn(v[b].w)};
                    *reinterpret_cast<uint2*>(static_cast<__half*>(p.out) + off) = *reinterpret_cast<const uint2*>(h4);
                }
            }
            if (b < nb && sflags[b] && tid == 0) p.counters[(size_t)(b0 + b) * (CLUSTER + 1) + rank] = 0u;
        }
    }
    CF_MARK(9);
}

}  // namespace cfb
