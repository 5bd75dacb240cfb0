/*
 * Batched paged decode with the weights streamed ONCE per chunk of BC requests.
 *
 * The reference's batch kernel launches one cluster per (request, head) -- grid 32*4*bs,
 * /root/reference/include/H100/llama/llama_kernel_batch_sglang_dispatch.cu:89 -- so every request re-streams all of
 * Wqkv and Wo (kernel_batch_sglang.cuh:43-664); so did this repo's first paged path (grid (Hq*4, bs) of the MHA
 * kernel).  On B200 a CTA ingests at most ~64 GB/s, so bs requests cost bs layers.  Here one 4-CTA cluster per head
 * serves a chunk of BC = 4 requests: every Wqkv / Wo tile that lands in shared memory is multiplied against the BC
 * activation vectors on the tensor cores (see below), the
 * chunk's K/V pages follow in the same tile stream, and the two cluster exchanges carry all BC requests at once:
 *
 *   exchange 1  BC x (q|k|v) = 6 KB of fp32 partial sums per CTA: reduce-scatter (cluster_scatter, slice r folded in
 *               rank order by CTA r) + all-gather (cluster_reduce<.., QUK_DEEPSEEK>) -- an all-to-all of 6 KB vectors
 *               would need 24 KB of receive slots;
 *   exchange 2  BC softmax states [m, l, o[128]] per CTA, all-gathered and merged in rank order.
 *
 * Both GEMVs run on the tensor cores: with BC activation vectors the product is a skinny GEMM (N = BC <= 8), and on the
 * CUDA cores it was issue-bound (BC FMAs per weight element: QKV 22.6 us, O 7.2 us per layer at BC = 4).  Weight tiles are
 * [32 rows x 128 cols] = two 128-byte-swizzled TMA boxes, read with ldmatrix.x4 as the A operand of mma.sync.m16n8k16
 * (fp16 in, fp32 accumulate); the requests sit on the N dimension (B fragments straight from the fp16 activations in
 * shared memory, padded row stride -> conflict-free); 16 ldmatrix + 16 mma per 8 KB tile.  QKV tile g = row block g % 12
 * (owned by warp g % 12, accumulated over its 8 column tiles in registers; deterministic), column tile g / 12.
 * Same decomposition otherwise as llama_decoder_kernel.cuh (K-split QKV, sequence-split KV, N-split O, 24 x 8 KB
 * self-issuing ring, fp32 red + last-arriver finalize per request).  PAGED variant, MHA, hidden <= 4096, page size 1.
 */
#pragma once

#include "llama_decoder_kernel.cuh"

namespace cfb {

constexpr int BK_KS_MAX = 1024;             // hidden / CLUSTER
constexpr int BK_XS_STRIDE = BK_KS_MAX + 8; // fp16 activations per request, padded: conflict-free B-fragment loads

template <int BC>
struct SmemB {
    static constexpr int CL = 4;
    static constexpr int QKV_OUT = 3 * HEAD_DIM;                          // 384
    static constexpr int SLICE1 = BC * QKV_OUT / CL;                      // floats per CTA slice in exchange 1 (BC * 96)
    static constexpr int PAY = HEAD_DIM + 4;                              // [m, l, -, -, o[128]]
    static constexpr int RING = 0;
    static constexpr int UNION = RING + NSTAGES * STAGE_BYTES;
    //   phase QKV : xs fp16 [BC][BK_KS_MAX]            phase ATTN: attn_part fp32 [24][132]
    static constexpr int XS = UNION;
    static constexpr int ATTN_PART = UNION;
    static constexpr int UNION_BYTES = (BC * BK_XS_STRIDE * 2 > 24 * PAY * 4) ? BC * BK_XS_STRIDE * 2 : 24 * PAY * 4;
    // X1: exchange-1 buffers; all dead after RoPE
    static constexpr int X1 = UNION + UNION_BYTES;
    static constexpr int QKV_SRC = X1;                                    // fp32 [BC][384]   this CTA's partial sums
    static constexpr int RED1 = QKV_SRC + BC * QKV_OUT * 4;               // fp32 [SLICE1]    my folded slice
    static constexpr int AG_RECV = RED1 + SLICE1 * 4;                     // fp32 [CL][SLICE1] = full [BC][384]
    static constexpr int X1_BYTES = BC * QKV_OUT * 4 + SLICE1 * 4 + CL * SLICE1 * 4;
    // exchange-2 buffers alias X1 (a peer can only push them after it finished exchange 1, which needed this CTA's
    // all-gather contribution, which this CTA sends after it is done with QKV_SRC / RED1; AG_RECV is read before then)
    static constexpr int ATTN_SRC = X1;                                   // fp32 [BC][132]
    static constexpr int ATTN_RECV = ATTN_SRC + BC * PAY * 4;             // fp32 [CL][BC][132]
    static_assert((1 + CL) * BC * PAY * 4 <= X1_BYTES, "exchange-2 buffers must fit into the exchange-1 region");
    static constexpr int RS_RECV = X1 + X1_BYTES;                         // fp32 [CL][SLICE1]  reduce-scatter slots
    static constexpr int QKV_FIN = RS_RECV;                               // fp32 [BC][384]  (RS_RECV is dead after the fold)
    static_assert(CL * SLICE1 == BC * QKV_OUT, "QKV_FIN aliases RS_RECV exactly");
    // phase O: out_part fp32 [BC][BK_KS_MAX] aliases X1 + RS_RECV (everything there is dead once attn_out is written)
    static constexpr int OUT_PART = X1;
    static_assert(BC * BK_KS_MAX * 4 <= X1_BYTES + CL * SLICE1 * 4, "out_part must fit");
    static constexpr int ATTN_OUT = RS_RECV + CL * SLICE1 * 4;            // fp32 [BC][128]
    static constexpr int RED = ATTN_OUT + BC * HEAD_DIM * 4;              // fp32 [12 warps][BC] + [BC] new-token scores
    static constexpr int META = RED + (CONSUMER_WARPS + 1) * BC * 4;      // int [BC][4] requests, [BC][4] segments, u32 [BC+1] tile0, [1] n_seg
    static constexpr int BARS = META + (9 * BC + 2 + 3) / 4 * 16;         // u64 full[NSTAGES], xbar[4]
    static constexpr int FLAGS = BARS + (NSTAGES + 4) * 8;                // u32 [BC]
    static constexpr int TOTAL = FLAGS + ((BC * 4 + 15) & ~15);
    static_assert(BARS % 8 == 0, "mbarrier alignment");
    static_assert(TOTAL <= 227 * 1024, "shared-memory layout exceeds the 227 KB opt-in limit");
};

// Phase 0 of the batched kernels: the normalised activations of this CTA's K-slice for all BC requests.  The (request, 8-element
// chunk) items are dealt over ALL threads; a CTA only ever loads its own slice -- the sums of squares over the whole row are
// completed by a cluster all-gather of the BC slice sums (batch_rmsnorm_slice below; until late in round 2 every CTA re-read the
// whole rows of all BC requests for them: 128 KB of L2 reads per CTA at batch 8).
template <int BC>
struct BatchSlice {
    static constexpr int ITEMS = (BC * (BK_KS_MAX / 8) + CONSUMER_THREADS - 1) / CONSUMER_THREADS;
    uint4 x[ITEMS], r[ITEMS], w[ITEMS];
};
template <int BC>
__device__ __forceinline__ void batch_slice_load(BatchSlice<BC>& sl, const KParams& p, int b0, int nb, int hidden, int KS, uint32_t rank,
                                                 uint32_t tid) {
    const int cpk = KS / 8;
#pragma unroll
    for (int it = 0; it < BatchSlice<BC>::ITEMS; ++it) {
        const int item = (int)tid + it * CONSUMER_THREADS;
        const int b = item / cpk, ge = (int)rank * KS + (item % cpk) * 8;
        sl.x[it] = sl.r[it] = sl.w[it] = make_uint4(0, 0, 0, 0);
        if (item < BC * cpk && b < nb) {
            sl.x[it] = *reinterpret_cast<const uint4*>(p.x + (size_t)(b0 + b) * hidden + ge);
            sl.r[it] = *reinterpret_cast<const uint4*>(p.residual_in + (size_t)(b0 + b) * hidden + ge);
            sl.w[it] = *reinterpret_cast<const uint4*>(p.rms_w + ge);
        }
    }
}
// red[w * BC + b], w < NPARTS, holds the partial sums of squares of request b (written before the barrier the caller just passed)
template <int BC, int NPARTS = CONSUMER_WARPS>
__device__ __forceinline__ void batch_slice_store(const BatchSlice<BC>& sl, const KParams& p, int b0, int nb, int hidden, int KS,
                                                  uint32_t rank, uint32_t head, uint32_t tid, const float* red, __half* xs) {
    const int cpk = KS / 8;
#pragma unroll
    for (int it = 0; it < BatchSlice<BC>::ITEMS; ++it) {
        const int item = (int)tid + it * CONSUMER_THREADS;
        if (item >= BC * cpk) continue;
        const int b = item / cpk, e = (item % cpk) * 8;
        __align__(16) __half xn[8];
        if (b < nb) {
            float tot = 0.f;
#pragma unroll
            for (int w = 0; w < NPARTS; ++w) tot += red[w * BC + b];
            const float rstd = rsqrtf(tot / (float)hidden + p.eps);
            float f[8], w8[8], r8[8];
            unpack8(sl.x[it], f);
            unpack8(sl.w[it], w8);
            unpack8(sl.r[it], r8);
            __align__(16) __half hs[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) { hs[k] = __float2half_rn(f[k] + r8[k]); f[k] = __half2float(hs[k]); }
            if (head == 0)
                *reinterpret_cast<uint4*>(p.residual_out + (size_t)(b0 + b) * hidden + rank * KS + e) = *reinterpret_cast<const uint4*>(hs);
#pragma unroll
            for (int k = 0; k < 8; ++k) xn[k] = __float2half_rn(round_h(f[k] * rstd) * w8[k]);
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) xn[k] = __float2half_rn(0.f);
        }
        *reinterpret_cast<uint4*>(xs + b * BK_XS_STRIDE + e) = *reinterpret_cast<const uint4*>(xn);
    }
}

// KV segments of a chunk (warp 0 of every CTA; result in shared memory, made visible by the caller's block barrier).  The chunk's
// KV rows of this head, concatenated request after request, are cut into `nranks` equal ranges (multiples of 16 rows), so a CTA
// streams one or two SEGMENTS (request, row range) however ragged the batch is.  A request's new token is folded in by the rank
// that holds its last row ("owner"; an empty request: the rank its offset falls into).  Lane b handles request b, offsets are
// warp prefix sums (no single-thread loop in front of a block barrier).
//   meta [BC][4]  kv_base, position, new_slot, owner          mseg [BC][4]  request | owner-is-me << 8, row begin, row end, -
//   tile0 [BC+1]  first KV tile (phase-local) of segment s; entries >= n_seg hold the total          mmisc [0] n_seg
template <int BC>
__device__ __forceinline__ void batch_build_segments(const KParams& p, int b0, int nb, int rank, int nranks, uint32_t lane,
                                                     int* meta, int* mseg, uint32_t* tile0, int* mmisc) {
    const int b = (int)lane;
    int len = 0, kb = 0, ns = 0, pos = 0;
    if (b < nb) {
        kb = p.indptr[b0 + b];
        const int end = p.indptr[b0 + b + 1] - 1;
        len = end - kb;
        ns = p.indices[end];
        pos = (int)p.positions[b0 + b];                  // row of the RoPE table: read here, not behind exchange 1
    }
    int incl = len;
#pragma unroll
    for (int o = 1; o < BC; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if ((int)lane >= o) incl += v;
    }
    const int T = __shfl_sync(0xffffffffu, incl, BC - 1);
    const int off = incl - len;
    const int per = (((T + nranks - 1) / nranks) + ROWS512 - 1) & ~(ROWS512 - 1);
    const int c0 = min(rank * per, T), c1 = min(c0 + per, T);
    int s0 = max(c0, off) - off, s1 = min(c1, off + len) - off;
    const bool nonempty = b < nb && s1 > s0;
    const int owner = len > 0 ? (off + len - 1) / per : (per > 0 ? min(off / per, nranks - 1) : b % nranks);
    const bool has = b < nb && (nonempty || owner == rank);
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
    if (b < BC) { meta[b * 4 + 0] = kb; meta[b * 4 + 1] = pos; meta[b * 4 + 2] = ns; meta[b * 4 + 3] = owner; }
    if ((int)lane >= nseg && (int)lane <= BC) tile0[lane] = total;
    __syncwarp();
    if (has) {
        mseg[idx * 4 + 0] = b | ((owner == rank) ? 256 : 0);
        mseg[idx * 4 + 1] = s0;
        mseg[idx * 4 + 2] = s1;
        tile0[idx] = tincl - nt;
    }
    if (lane == 0) mmisc[0] = nseg;
    __syncwarp();
}

// Fused residual add + RMSNorm of the batched MHA kernels, K-split over the CLUSTER CTAs of a head.  `red` (16-byte aligned,
// >= 9 * BC floats; `red_u32` its shared-memory address) is scratch: [0, 4*BC) per-warp sums, [4*BC, 5*BC) this CTA's BC slice
// sums, [5*BC, 9*BC) the gathered sums of the CLUSTER ranks.  `bar_u32`: an exchange barrier armed with
// cluster_reduce_arm<CLUSTER>(bar, BC * 4); every peer must be known to have armed it (cluster barrier) before the call.
// The CLUSTER slice sums are added in rank order by everybody: all CTAs (and all heads) compute the same rstd.
template <int BC, int CLUSTER>
__device__ __forceinline__ void batch_rmsnorm_slice(const KParams& p, int b0, int nb, int hidden, int KS, uint32_t rank, uint32_t head,
                                                    uint32_t tid, float* red, uint32_t red_u32, uint32_t bar_u32, __half* xs) {
    static_assert(BC * 4 % 16 == 0, "the slice sums travel as 16-byte vectors");
    BatchSlice<BC> slice;
    batch_slice_load<BC>(slice, p, b0, nb, hidden, KS, rank, tid);
    const int cpk = KS / 8;                                   // items per request: a multiple of 32 (KS is a multiple of 256)
    float* sseg = red;
    float* ss_src = red + 4 * BC;
    float* ss_recv = red + 5 * BC;
#pragma unroll
    for (int it = 0; it < BatchSlice<BC>::ITEMS; ++it) {
        const int item = (int)tid + it * CONSUMER_THREADS;
        float f[8], r8[8];
        unpack8(slice.x[it], f);
        unpack8(slice.r[it], r8);
        float v = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) { const float h = round_h(f[k] + r8[k]); v = fmaf(h, h, v); }     // zeros past the end
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((tid & 31) == 0 && item < BC * cpk) sseg[item >> 5] = v;       // a warp's 32 items belong to one request
    }
    dsm::named_bar_sync(CONSUMER_BAR, CONSUMER_THREADS);
    if (tid < (uint32_t)BC) {
        const int nw = cpk / 32;                              // <= 4
        float a = 0.f;
        for (int k = 0; k < nw; ++k) a += sseg[tid * nw + k];
        ss_src[tid] = a;
    }
    uint32_t ph = 0;
    cluster_reduce<CLUSTER, Stage::QUK_DEEPSEEK, CONSUMER_THREADS, CONSUMER_BAR>(
        BC * 4, tid, BC, rank, red_u32 + 4 * BC * 4, red_u32 + 5 * BC * 4, bar_u32, ph, ss_src, ss_recv);
    batch_slice_store<BC, CLUSTER>(slice, p, b0, nb, hidden, KS, rank, head, tid, ss_recv, xs);
    dsm::named_bar_sync(CONSUMER_BAR, CONSUMER_THREADS);
}

template <int BC>
__global__ void __launch_bounds__(BLOCK_THREADS, 1)
llama_decoder_layer_batch_kernel(const __grid_constant__ KParams p)
{
    using S = SmemB<BC>;
    constexpr int CLUSTER = S::CL;

    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t smem_base = dsm::smem_u32(smem);
    const uint32_t tid = threadIdx.x;
    const uint32_t warp = tid >> 5;
    const uint32_t lane = tid & 31;
    const uint32_t rank = dsm::cluster_ctarank();
    const uint32_t head = blockIdx.x / CLUSTER;
    const int b0 = blockIdx.y * BC;                       // first request of this chunk
    const int nb = min(BC, p.batch - b0);      // live requests in the chunk (>= 1)

    const int hidden = p.hidden;
    const int KS = hidden / CLUSTER;
    const int kv_cols = p.n_kv_heads * HEAD_DIM;

    const uint32_t full_u32 = smem_base + S::BARS;
    const uint32_t xbar_u32 = full_u32 + NSTAGES * 8;

    // ---- KV segments of this CTA (batch_build_segments): all 12 warps work on the same request at a time ----------------
    int* meta = reinterpret_cast<int*>(smem + S::META);                       // [b][4]
    int* mseg = meta + BC * 4;                                                // [s][4]
    uint32_t* kv_tile0 = reinterpret_cast<uint32_t*>(mseg + BC * 4);          // [BC + 1]
    int* mmisc = reinterpret_cast<int*>(kv_tile0 + BC + 1);
    if (warp == 0) batch_build_segments<BC>(p, b0, nb, (int)rank, CLUSTER, lane, meta, mseg, kv_tile0, mmisc);
    __syncthreads();                                                          // segments visible to every warp
    const int n_seg = mmisc[0];
    const uint32_t n_qkv_tiles = (uint32_t)CONSUMER_WARPS * (KS / 128);          // 12 row blocks of 32 x KS/128 column tiles
    const uint32_t n_kv_tiles = kv_tile0[BC];
    const uint32_t n_o_tiles = (uint32_t)(KS / ROWS256);
    const uint32_t total_tiles = n_qkv_tiles + n_kv_tiles + n_o_tiles;

    CF_MARK(0);
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    const uint64_t pol = policy_evict_first();
    const __half* kpool = reinterpret_cast<const __half*>(p.k_pool_ptrs[p.layer_id]);
    const __half* vpool = reinterpret_cast<const __half*>(p.v_pool_ptrs[p.layer_id]);
    // host's copy of the pool addresses (tensor maps over the pools): used only if it agrees with the device table
    const bool pool_maps = p.k_base != nullptr && kpool == p.k_base && vpool == p.v_base;

    // page index of this lane's row of KV tile g (phase-global index), fetched one ring cycle ahead of the tile
    int pre_slot0 = 0, pre_slot1 = 0;
    uint32_t pre_g0 = 0xffffffffu, pre_g1 = 0xffffffffu;
    auto seg_of = [&](uint32_t t) -> int {                // which segment KV tile t (phase-local index) belongs to
        int sgi = 0;
#pragma unroll
        for (int q = 1; q < BC; ++q) sgi += (t >= kv_tile0[q]) ? 1 : 0;
        return sgi;
    };
    auto page_of = [&](uint32_t g) -> int {
        const uint32_t t = g - n_qkv_tiles;
        const int sgi = seg_of(t);
        const int r = mseg[sgi * 4 + 1] + (int)(t - kv_tile0[sgi]) * ROWS512 + (int)(lane & 15);
        return (r < mseg[sgi * 4 + 2]) ? p.indices[meta[(mseg[sgi * 4] & 255) * 4] + r] : 0;
    };
    auto issue_tile = [&](uint32_t g) {
        if (g >= total_tiles) return;
        const uint32_t s = ring_stage(g);
        const uint32_t fb = full_u32 + 8 * s;
        const uint32_t dst = smem_base + S::RING + s * STAGE_BYTES;
        if (g < n_qkv_tiles) {
            if (lane == 0) {
                // tile g: row block g % 12 (32 of the head's 384 q|k|v rows, owned by warp g % 12), column tile g / 12
                // (128 input columns) = two [32 rows x 64 cols] boxes, 128-byte swizzled for ldmatrix
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
            const int i = (int)(t - kv_tile0[sgi]);
            const bool odd = (g / CONSUMER_WARPS) & 1u;
            const long long slot = (odd ? pre_g1 : pre_g0) == g ? (long long)(odd ? pre_slot1 : pre_slot0) : (long long)page_of(g);
            const int nvalid = min(ROWS512, rend - (rbeg + i * ROWS512));
            issue_kv_stage(p, pool_maps, false, dst, fb, head * HEAD_DIM, slot, nvalid, kpool, vpool, kv_cols, lane, pol);
        } else {
            if (lane == 0) {
                const uint32_t i = g - n_qkv_tiles - n_kv_tiles;     // Wo [out][in]: 32 output rows x this head's 128 input cols
                dsm::mbar_arrive_expect_tx(fb, STAGE_BYTES);          // as two swizzled [32 x 64] boxes
                tma_load_2d(dst, &p.tm_wo, head * HEAD_DIM, rank * KS + i * ROWS256, fb, pol);
                tma_load_2d(dst + 4096, &p.tm_wo, head * HEAD_DIM + 64, rank * KS + i * ROWS256, fb, pol);
            }
        }
        const uint32_t g2 = g + NSTAGES;              // the tile that will live in this stage next
        if (g2 >= n_qkv_tiles && g2 < n_qkv_tiles + n_kv_tiles) {
            const int pg = page_of(g2);
            if ((g / CONSUMER_WARPS) & 1u) { pre_slot1 = pg; pre_g1 = g2; } else { pre_slot0 = pg; pre_g0 = g2; }
        }
    };

    if (lane == 0) {
        dsm::mbar_init(full_u32 + 8 * warp, 1);
        dsm::mbar_init(full_u32 + 8 * (warp + CONSUMER_WARPS), 1);
        if (tid == 0) {
            prefetch_tmap(&p.tm_wqkv);
            prefetch_tmap(&p.tm_wo);
            if (pool_maps) { prefetch_tmap(&p.tm_k); prefetch_tmap(&p.tm_v); prefetch_tmap(&p.tm_kg); prefetch_tmap(&p.tm_vg); }
            cluster_reduce_arm<CLUSTER>(xbar_u32, S::SLICE1 * 4);
            cluster_reduce_arm<CLUSTER>(xbar_u32 + 8, S::SLICE1 * 4);
            cluster_reduce_arm<CLUSTER>(xbar_u32 + 16, BC * S::PAY * 4);
            cluster_reduce_arm<CLUSTER>(xbar_u32 + 24, BC * 4);              // sums of squares
        }
        dsm::mbar_fence_init();
    }
    __syncwarp();
    CF_MARK(12);
    issue_tile(warp);
    issue_tile(warp + CONSUMER_WARPS);
    dsm::cluster_arrive();

    __half* xs = reinterpret_cast<__half*>(smem + S::XS);
    float* attn_part = reinterpret_cast<float*>(smem + S::ATTN_PART);
    float* qkv_src = reinterpret_cast<float*>(smem + S::QKV_SRC);
    float* rs_recv = reinterpret_cast<float*>(smem + S::RS_RECV);
    float* red1 = reinterpret_cast<float*>(smem + S::RED1);
    float* ag_recv = reinterpret_cast<float*>(smem + S::AG_RECV);
    float* qkv_fin = reinterpret_cast<float*>(smem + S::QKV_FIN);
    float* attn_src = reinterpret_cast<float*>(smem + S::ATTN_SRC);
    float* attn_recv = reinterpret_cast<float*>(smem + S::ATTN_RECV);
    float* attn_out = reinterpret_cast<float*>(smem + S::ATTN_OUT);
    float* out_part = reinterpret_cast<float*>(smem + S::OUT_PART);
    float* red = reinterpret_cast<float*>(smem + S::RED);
    uint32_t* sflags = reinterpret_cast<uint32_t*>(smem + S::FLAGS);

    asm volatile("griddepcontrol.wait;" ::: "memory");

    // ---- phase 0: fused residual add + RMSNorm for the BC requests of the chunk, K-split over the cluster --------
    dsm::cluster_wait();                 // every peer has armed its exchange barriers (it arrived before its dependency wait)
    batch_rmsnorm_slice<BC, CLUSTER>(p, b0, nb, hidden, KS, rank, head, tid, red, smem_base + S::RED, xbar_u32 + 24, xs);

    CF_MARK(1);
    uint32_t gbase = 0;
    // ---- phase 1: QKV GEMV on the tensor cores: D[32 rows][8] += W[32 x 128] * X^T[128 x 8], X = the chunk's BC
    //      activation vectors (columns BC..7 are zero).  Warp w owns row block w (32 of the head's 384 q|k|v rows) and
    //      accumulates over its 8 column tiles in registers: 16 ldmatrix.x4 + 16 mma.sync per 8 KB tile. -----------------
    {
        static_assert(BC <= 8 && BC % 2 == 0, "requests sit on the N = 8 dimension of m16n8k16, two per lane");
        const int g4 = lane >> 2, t4 = lane & 3;
        const int lrow = lane & 7, lmat = lane >> 3;
        float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
        for (uint32_t i = warp; i < n_qkv_tiles; i += CONSUMER_WARPS) {        // phase starts at ring index 0
            const uint32_t g = i, s = ring_stage(g);
            const int ct = (int)(i / CONSUMER_WARPS);                           // column tile: input cols ct*128 .. +128
            // B fragments of this column tile: x[request g4][k], k = ct*128 + ks*16 + 2*t4 (+8)
            uint32_t xb[8][2];
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
                const __half* xp = xs + g4 * BK_XS_STRIDE + ct * 128 + ks * 16 + t4 * 2;
                xb[ks][0] = g4 < BC ? *reinterpret_cast<const uint32_t*>(xp) : 0u;
                xb[ks][1] = g4 < BC ? *reinterpret_cast<const uint32_t*>(xp + 8) : 0u;
            }
            ring_wait_full(full_u32, g);
            const uint32_t st = smem_base + S::RING + s * STAGE_BYTES;
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
#pragma unroll
                for (int mb = 0; mb < 2; ++mb) {
                    const int row = mb * 16 + (lmat & 1) * 8 + lrow, chunk = (ks & 3) * 2 + (lmat >> 1);
                    const uint32_t addr = st + (ks >> 2) * 4096 + row * 128 + ((chunk ^ (row & 7)) << 4);
                    uint32_t a0, a1, a2, a3;
                    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                                 : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "r"(addr));
                    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                                 : "+f"(acc[mb][0]), "+f"(acc[mb][1]), "+f"(acc[mb][2]), "+f"(acc[mb][3])
                                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(xb[ks][0]), "r"(xb[ks][1]));
                }
            }
            __syncwarp();
            issue_tile(g + NSTAGES);
        }
        // C fragment: rows g4 / g4 + 8 of each 16-row block, requests 2*t4 and 2*t4 + 1
        if (2 * t4 < BC) {
#pragma unroll
            for (int mb = 0; mb < 2; ++mb) {
                const int row = (int)warp * 32 + mb * 16 + g4;
                qkv_src[(2 * t4) * S::QKV_OUT + row] = acc[mb][0];
                qkv_src[(2 * t4 + 1) * S::QKV_OUT + row] = acc[mb][1];
                qkv_src[(2 * t4) * S::QKV_OUT + row + 8] = acc[mb][2];
                qkv_src[(2 * t4 + 1) * S::QKV_OUT + row + 8] = acc[mb][3];
            }
        }
    }
    gbase += n_qkv_tiles;
    CF_MARK(2);

    // ---- exchange 1: reduce-scatter (sum, rank order) + all-gather of BC x (q|k|v) ---------------------------
    uint32_t ph0 = 0, ph1 = 0, ph2 = 0;
    cluster_scatter<CLUSTER, CONSUMER_THREADS, CONSUMER_BAR>(S::SLICE1 * 4, tid, rank, smem_base + S::RS_RECV, xbar_u32, ph0,
                                                             qkv_src, rs_recv);
    for (int e = tid; e < S::SLICE1; e += CONSUMER_THREADS) {
        float a = 0.f;
#pragma unroll
        for (int r = 0; r < CLUSTER; ++r) a += rs_recv[r * S::SLICE1 + e];
        red1[e] = round_h(a);                                    // q / k / v leave the projection as fp16 (eager model)
    }
    cluster_reduce<CLUSTER, Stage::QUK_DEEPSEEK, CONSUMER_THREADS, CONSUMER_BAR>(
        S::SLICE1 * 4, tid, S::SLICE1, rank, smem_base + S::RED1, smem_base + S::AG_RECV, xbar_u32 + 8, ph1, red1, ag_recv);

    CF_MARK(3);
    // ---- RoPE (NeoX) per request, new K/V rows into the pool ---------------------------------------------------
    {
        constexpr float kScaleLog2 = 0.08838834764831845f * 1.4426950408889634f;
        for (int f = tid; f < BC * S::QKV_OUT; f += CONSUMER_THREADS) {
            const int b = f / S::QKV_OUT, e = f % S::QKV_OUT;
            const int which = e >> 7, d = e & 127;               // 0: q, 1: k, 2: v
            const float a = ag_recv[f];
            float outv = a;
            if (b < nb) {
                if (which < 2) {
                    const float* cosp = p.cos + (size_t)meta[b * 4 + 1] * HEAD_DIM;
                    const float* sinp = cosp + HEAD_DIM / 2;
                    const float bb = ag_recv[f ^ 64];
                    const int i = d & 63;
                    const float rot = (d & 64) ? fmaf(a, cosp[i], bb * sinp[i]) : fmaf(a, cosp[i], -bb * sinp[i]);
                    const __half rh = __float2half_rn(rot);
                    outv = which == 0 ? __half2float(rh) * kScaleLog2 : __half2float(rh);
                    if (which == 1 && rank == 0) {
                        __half* kp = reinterpret_cast<__half*>(p.k_pool_ptrs[p.layer_id]);
                        kp[(size_t)meta[b * 4 + 2] * kv_cols + head * HEAD_DIM + d] = rh;
                    }
                } else if (rank == 0) {
                    __half* vp = reinterpret_cast<__half*>(p.v_pool_ptrs[p.layer_id]);
                    vp[(size_t)meta[b * 4 + 2] * kv_cols + head * HEAD_DIM + d] = __float2half_rn(a);
                }
            }
            qkv_fin[f] = outv;        // aliases rs_recv: every thread finished folding it before the all-gather barrier
        }
        dsm::named_bar_sync(CONSUMER_BAR, CONSUMER_THREADS);
    }
    // exchange-2 receive slots alias ag_recv: tell the peers this CTA is done reading it (waited on before exchange 2)
    dsm::cluster_arrive();

    CF_MARK(4);
    // ---- phase 2: flash-decode over this CTA's segments, all 12 warps on one request at a time -------------------------
    {
        const int sub = lane >> 4, c = lane & 15;
        // block-merged state of request b on this CTA: thread tid < 128 keeps o[dim tid] in registers, (m, l) in red[2b], red[2b+1]
        // (written and read by thread 0 only); requests without a segment here stay (-inf, 0, 0)
        float Ov[BC];
#pragma unroll
        for (int b = 0; b < BC; ++b) Ov[b] = 0.f;
        if (tid == 0) {
#pragma unroll
            for (int b = 0; b < BC; ++b) { red[2 * b] = -INFINITY; red[2 * b + 1] = 0.f; }
        }
        for (int sg = 0; sg < n_seg; ++sg) {
            const int b = mseg[sg * 4] & 255;
            const bool owner = (mseg[sg * 4] & 256) != 0;
            const int rbeg = mseg[sg * 4 + 1], rend = mseg[sg * 4 + 2];
            const uint32_t nt = kv_tile0[sg + 1] - kv_tile0[sg];
            const uint32_t gb = gbase + kv_tile0[sg];
            float m = -INFINITY, l = 0.f, o8[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) o8[k] = 0.f;
            float q8[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) q8[k] = qkv_fin[b * S::QKV_OUT + c * 8 + k];
            for (uint32_t i = first_tile(gb, warp); i < nt; i += CONSUMER_WARPS) {
                const uint32_t g = gb + i, s = ring_stage(g);
                ring_wait_full(full_u32, g);
                const uint4* kt = reinterpret_cast<const uint4*>(smem + S::RING + s * STAGE_BYTES);
                const uint4* vt = kt + STAGE_BYTES / 32;
                const int rows_left = rend - (rbeg + (int)i * ROWS512);     // >= 1
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
                    const float pr = dsm::fast_exp2(sc[jj] - m_use);       // -inf -> 0
                    l += pr;
                    uint4 raw = vt[row * 16 + c];
                    if (row >= rows_left) raw = make_uint4(0, 0, 0, 0);    // rows past the end were never copied
                    float v8[8];
                    unpack8(raw, v8);
#pragma unroll
                    for (int k = 0; k < 8; ++k) o8[k] = fmaf(pr, v8[k], o8[k]);
                }
                m = m_new;
                __syncwarp();
                issue_tile(g + NSTAGES);
            }
            // block merge through the 24 x 132 buffer (two half-warp states per warp); the owner folds in the request's new token
            {
                const int grp = warp * 2 + sub;
                float* slot = attn_part + grp * S::PAY;
                if (c == 0) { slot[0] = m; slot[1] = l; }
                *reinterpret_cast<float4*>(slot + 4 + c * 8) = make_float4(o8[0], o8[1], o8[2], o8[3]);
                *reinterpret_cast<float4*>(slot + 4 + c * 8 + 4) = make_float4(o8[4], o8[5], o8[6], o8[7]);
            }
            if (owner && warp == 0) {
                float a = 0.f;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    a = fmaf(qkv_fin[b * S::QKV_OUT + lane * 4 + k], qkv_fin[b * S::QKV_OUT + HEAD_DIM + lane * 4 + k], a);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
                if (lane == 0) red[CONSUMER_WARPS * BC + b] = a;
            }
            dsm::named_bar_sync(CONSUMER_BAR, CONSUMER_THREADS);
            if (tid < HEAD_DIM) {
                const float s_new = red[CONSUMER_WARPS * BC + b];
                float M = owner ? s_new : -INFINITY;
#pragma unroll
                for (int gI = 0; gI < 2 * CONSUMER_WARPS; ++gI) M = fmaxf(M, attn_part[gI * S::PAY]);
                float L = 0.f, Ovv = 0.f;
#pragma unroll
                for (int gI = 0; gI < 2 * CONSUMER_WARPS; ++gI) {
                    const float w = dsm::exp2_diff(attn_part[gI * S::PAY], M);
                    L = fmaf(attn_part[gI * S::PAY + 1], w, L);
                    Ovv = fmaf(attn_part[gI * S::PAY + 4 + tid], w, Ovv);
                }
                if (owner) {
                    const float w = dsm::exp2_diff(s_new, M);
                    L += w;
                    Ovv = fmaf(qkv_fin[b * S::QKV_OUT + 2 * HEAD_DIM + tid], w, Ovv);
                }
#pragma unroll
                for (int q = 0; q < BC; ++q) Ov[q] = (q == b) ? Ovv : Ov[q];
                if (tid == 0) { red[2 * b] = M; red[2 * b + 1] = L; }
            }
            dsm::named_bar_sync(CONSUMER_BAR, CONSUMER_THREADS);
        }
        gbase += n_kv_tiles;
        CF_MARK(5);
        if (tid < HEAD_DIM) {
#pragma unroll
            for (int b = 0; b < BC; ++b) {
                float* st = attn_src + b * S::PAY;       // aliases qkv_src / red1: dead since the all-gather
                st[4 + tid] = Ov[b];
                if (tid == 0) { st[0] = red[2 * b]; st[1] = red[2 * b + 1]; st[2] = 0.f; st[3] = 0.f; }
            }
        }
        // ---- exchange 2: all-gather of the BC softmax states, merged in rank order -------------------------------
        dsm::cluster_wait();          // every peer is past RoPE: its exchange-1 buffers may now be overwritten
        cluster_reduce<CLUSTER, Stage::QUK_DEEPSEEK, CONSUMER_THREADS, CONSUMER_BAR>(
            BC * S::PAY * 4, tid, BC * S::PAY, rank, smem_base + S::ATTN_SRC, smem_base + S::ATTN_RECV, xbar_u32 + 16, ph2,
            attn_src, attn_recv);
        for (int f = tid; f < BC * HEAD_DIM; f += CONSUMER_THREADS) {
            const int b = f >> 7, d = f & 127;
            float M = -INFINITY;
#pragma unroll
            for (int r = 0; r < CLUSTER; ++r) M = fmaxf(M, attn_recv[(r * BC + b) * S::PAY]);
            float L = 0.f, Ov = 0.f;
#pragma unroll
            for (int r = 0; r < CLUSTER; ++r) {
                const float* st = attn_recv + (r * BC + b) * S::PAY;
                const float w = dsm::exp2_diff(st[0], M);
                L = fmaf(st[1], w, L);
                Ov = fmaf(st[4 + d], w, Ov);
            }
            attn_out[f] = (b < nb) ? round_h(Ov / L) : 0.f;          // attention output leaves as fp16 (eager model)
        }
        dsm::named_bar_sync(CONSUMER_BAR, CONSUMER_THREADS);
    }

    CF_MARK(6);
    // ---- phase 3: O GEMV on the tensor cores: D[32 out rows][8] = Wo tile[32 x 128] * A^T[128 x 8], A = the BC attention
    //      outputs of this head (fp16-rounded values, exact as fp16) ----------------------------------------------------
    {
        const int g4 = lane >> 2, t4 = lane & 3;
        const int lrow = lane & 7, lmat = lane >> 3;
        uint32_t ob[8][2];
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
#pragma unroll
            for (int hi = 0; hi < 2; ++hi) {
                const int d = ks * 16 + hi * 8 + t4 * 2;
                const __half2 h2 = g4 < BC ? __floats2half2_rn(attn_out[g4 * HEAD_DIM + d], attn_out[g4 * HEAD_DIM + d + 1])
                                           : __floats2half2_rn(0.f, 0.f);
                ob[ks][hi] = *reinterpret_cast<const uint32_t*>(&h2);
            }
        }
        for (uint32_t i = first_tile(gbase, warp); i < n_o_tiles; i += CONSUMER_WARPS) {
            const uint32_t g = gbase + i, s = ring_stage(g);
            ring_wait_full(full_u32, g);
            const uint32_t st = smem_base + S::RING + s * STAGE_BYTES;
            float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
#pragma unroll
                for (int mb = 0; mb < 2; ++mb) {
                    const int row = mb * 16 + (lmat & 1) * 8 + lrow, chunk = (ks & 3) * 2 + (lmat >> 1);
                    const uint32_t addr = st + (ks >> 2) * 4096 + row * 128 + ((chunk ^ (row & 7)) << 4);
                    uint32_t a0, a1, a2, a3;
                    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                                 : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "r"(addr));
                    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                                 : "+f"(acc[mb][0]), "+f"(acc[mb][1]), "+f"(acc[mb][2]), "+f"(acc[mb][3])
                                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(ob[ks][0]), "r"(ob[ks][1]));
                }
            }
            __syncwarp();
            issue_tile(g + NSTAGES);
            if (2 * t4 < BC) {
#pragma unroll
                for (int mb = 0; mb < 2; ++mb) {
                    const int row = (int)i * ROWS256 + mb * 16 + g4;
                    out_part[(2 * t4) * BK_KS_MAX + row] = acc[mb][0];
                    out_part[(2 * t4 + 1) * BK_KS_MAX + row] = acc[mb][1];
                    out_part[(2 * t4) * BK_KS_MAX + row + 8] = acc[mb][2];
                    out_part[(2 * t4 + 1) * BK_KS_MAX + row + 8] = acc[mb][3];
                }
            }
        }
    }
    dsm::named_bar_sync(CONSUMER_BAR, CONSUMER_THREADS);

    CF_MARK(7);
    // ---- cross-head reduction per request: fp32 red into scratch, last arriver of the slice finalises --------------
#pragma unroll
    for (int b = 0; b < BC; ++b) {
        if (b < nb) {
            float* scratch = p.scratch + (size_t)(b0 + b) * hidden + rank * KS;
            for (int e = tid * 4; e < KS; e += CONSUMER_THREADS * 4)
                red_add_v4(scratch + e, *reinterpret_cast<const float4*>(out_part + b * BK_KS_MAX + e));
        }
    }
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
        // finalise every request this CTA arrived last for: all scratch loads first (one L2 round trip), then the stores
        __threadfence();
        const bool fp32_out = p.flags & 1u;
        const int e = tid * 4;                                // KS <= 1024: one float4 per thread
        float4 v[BC];
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
