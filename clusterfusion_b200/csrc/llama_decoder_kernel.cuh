/*
 * Fused Llama decoder attention half-layer for sm_100a (B200):
 *
 *   RMSNorm -> QKV GEMV -> RoPE -> (paged-)KV flash-decode -> O GEMV -> fp32 cross-head reduction
 *
 * One launch per layer, one thread-block cluster per (request, KV-head group).  Replaces the three
 * reference kernels (/root/reference/include/H100/llama/kernel.cuh:20-619, kernel_sglang.cuh:20-633,
 * kernel_batch_sglang.cuh:43-664); the decomposition inside a cluster follows the paper (K-split of
 * the QKV GEMV, sequence-split of the KV cache, N-split of the O GEMV) but the machinery is new:
 *
 *  - one tile stream for EVERY byte the CTA will ever need -- Wqkv tiles, then K/V tiles, then Wo tiles --
 *    through 24 x 8 KB TMA stages.  None of those loads depends on activations, so the stream runs ahead
 *    across phase boundaries: K/V tiles land while the warps are still in the QKV cluster exchange / RoPE,
 *    Wo tiles land during the softmax merge, and under programmatic dependent launch the first 24 tiles are
 *    requested while the previous layer is still finishing.  (The reference drains a 2-stage,
 *    single-thread-producer pipeline five times per layer, kernel.cuh:141-267, :343, :579.)
 *  - 12 warps, each a self-contained double-buffered stream: tile g lives in stage g % 24 and belongs to warp
 *    g % 12; a warp consumes its tile, then re-issues the TMA load for tile g + 24 into the stage it just freed
 *    while its other stage is already landing.  No producer warp, no empty barriers, no head-of-line blocking;
 *    one full mbarrier per stage whose phase parity cannot alias because a stage only ever has one owner.
 *  - every reduction is fp32: per-warp partials land in write-once shared-memory slots and are
 *    folded in a fixed order; the cluster exchanges (q|k|v vector, softmax state) go through the new
 *    cluster_reduce<> in include/dsm.cuh (one st.async push per peer, no cluster.sync());
 *    2 exchanges per layer instead of the reference's 22 cluster.sync().
 *  - the cross-head O reduction is fp32 `red.global.add.v4.f32` into a 4*hidden-byte scratch plus an
 *    arrival counter per output slice; the last-arriving CTA converts to fp16, writes `out`, and
 *    re-zeroes scratch + counter, so there are no memset launches and no fp16 atomics
 *    (reference: 131 072 fp16 atomicAdd per layer, kernel.cuh:600/:618).
 *  - the two GEMVs read their 128-byte-swizzled weight tiles with ldmatrix and multiply them with mma.sync.m16n8k16
 *    (activation vector on column 0 of N = 8): not for tensor throughput -- the layer is HBM-bound -- but because it
 *    halves the issue slots; the MHA flash-decode loop (one FMA per K/V byte) stays on the CUDA cores.
 *
 * Shapes: head_dim 128; hidden % (CLUSTER*256) == 0, hidden/CLUSTER <= 2048.
 */
#pragma once

#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/dsm.cuh"

namespace cfb {

constexpr int HEAD_DIM = 128;
constexpr int CONSUMER_WARPS = 12;
constexpr int CONSUMER_THREADS = CONSUMER_WARPS * 32;   // 384
constexpr int BLOCK_THREADS = CONSUMER_THREADS;         // no producer warp: every warp streams its own tiles
constexpr int STAGE_BYTES = 8192;
constexpr int STAGES_PER_WARP = 2;
constexpr int NSTAGES = STAGES_PER_WARP * CONSUMER_WARPS;   // stage s is consumed by warp s % 12, always (see ring_wait_full)
constexpr int ROWS256 = STAGE_BYTES / 256;              // rows of a {128 halves wide} tile            (32)
constexpr int ROWS512 = STAGE_BYTES / 512;              // rows of a {256 halves wide} tile / KV rows per stage (16)
constexpr int KS_MAX = 2048;                            // max hidden / CLUSTER
constexpr int CONSUMER_BAR = 1;                         // named barrier id for the consumer threads
constexpr long long POOL_MAP_ROWS = 1ll << 24;          // row extent of the tensor maps the host builds over a paged KV pool

enum Variant : int { CHAT = 0, SGLANG = 1, PAGED = 2 };
static_assert(CONSUMER_WARPS % 3 == 0 && CONSUMER_THREADS >= 3 * HEAD_DIM, "chat QKV mapping needs warps % 3 == 0");

struct alignas(64) KParams {
    CUtensorMap tm_wqkv;   // CHAT: [3*hidden][hidden] box {128,64}; else [(Hq+2Hkv)*128][hidden] box {256,32}
    CUtensorMap tm_wo;     // CHAT: [Hq*128][hidden] box {128,64};  else [hidden][Hq*128] box {128,64}
    CUtensorMap tm_k;      // CHAT/SGLANG: [>= 2^24][Hkv*128] from k_cache, box {128,16}; PAGED (k_base given): the same over the pool
    CUtensorMap tm_v;
    CUtensorMap tm_kg;     // PAGED (k_base given): the pool again with box {128,1} for tile::gather4 requests
    CUtensorMap tm_vg;
    const __half* x;
    const __half* residual_in;
    const __half* rms_w;
    void* out;
    __half* residual_out;
    __half* k_new;
    __half* v_new;
    const float* cos;
    const float* sin;
    const int* indptr;
    const int* indices;
    const unsigned long long* k_pool_ptrs;
    const unsigned long long* v_pool_ptrs;
    // CHAT / SGLANG: k_cache / v_cache (tm_k / tm_v then map [2^24+ rows][kv_cols] from the same base: the maps depend on
    // the base pointer only, so a growing cache never re-encodes them; a ragged last tile is fetched row by row).
    // PAGED, optional: the host's copy of k_pool_ptrs[layer_id] / v_pool_ptrs[layer_id]; tm_k / tm_v then map the pools.
    const __half* k_base;
    const __half* v_base;
    const long long* positions;
    float* scratch;          // fp32 [batch][hidden], zero between launches
    unsigned* counters;      // [batch][CLUSTER + 1], zero between launches
    unsigned long long* out_ll;   // batch == 1: (float, epoch) words [clusters][hidden] for the cross-head O reduction
    unsigned* header;        // workspace header: u32 [0] launch epoch (tags every flag-in-data word)
    float eps;
    int hidden;
    int n_heads;             // query heads == clusters per request (MHA kernels)
    int n_kv_heads;
    int kv_len;
    int layer_id;
    unsigned flags;
    int batch;               // requests in the launch (batched paged kernel)
    // head-parallel shards (Llama-2-70B): fused all-reduce of the O-projection partial over NVLink peer memory
    unsigned long long* tp_peer[8];   // rank r's exchange buffer: (float, epoch) words [2 parities][tp_world][hidden]
    int tp_rank, tp_world;            // tp_world <= 1: no peer stage
};

// ------------------------------------------------------------------------------------------------
// shared memory carve-up (dynamic)
// ------------------------------------------------------------------------------------------------
template <int CLUSTER>
struct Smem {
    static constexpr int QKV_OUT = 3 * HEAD_DIM;                 // q | k | v of one head
    static constexpr int ATTN_PAYLOAD = HEAD_DIM + 4;            // [m, l, -, -, o[128]]
    static constexpr int RING = 0;
    static constexpr int UNION = RING + NSTAGES * STAGE_BYTES;
    //   phase QKV : xs fp32[KS_MAX] | qkv_part  (chat only: fp32[12 warps][128]; nn.Linear layout keeps its sums in registers)
    //   phase ATTN: attn_part fp32[24][132]
    //   phase O   : out_part  (chat: fp32[4][KS <= 1024]; sglang: fp32[KS])
    static constexpr int UNION_BYTES = KS_MAX * 4 + 8 * QKV_OUT * 4;   // 20480 >= 2*KS_MAX*4
    static constexpr int XS = UNION;
    static constexpr int QKV_PART = UNION + KS_MAX * 4;
    static constexpr int ATTN_PART = UNION;
    static constexpr int OUT_PART = UNION;
    static constexpr int QKV_SRC = UNION + UNION_BYTES;                    // fp32[384]  contribution / result
    // each exchange barrier is used exactly once per launch (phase 0), so only the phase-0 half of
    // cluster_reduce's double buffer is ever touched and only that half is allocated
    static constexpr int QKV_RECV = QKV_SRC + QKV_OUT * 4;                 // fp32[CLUSTER][384]
    static constexpr int ATTN_SRC = QKV_RECV + CLUSTER * QKV_OUT * 4;      // fp32[132]
    static constexpr int ATTN_RECV = ATTN_SRC + ATTN_PAYLOAD * 4;          // fp32[CLUSTER][132]
    static constexpr int QKV_FINAL = ATTN_RECV + CLUSTER * ATTN_PAYLOAD * 4;  // fp32[384] roped q*scale | k | v
    static constexpr int ATTN_OUT = QKV_FINAL + QKV_OUT * 4;               // fp32[128]
    static constexpr int RED = ATTN_OUT + HEAD_DIM * 4;                    // fp32[32] block-reduce scratch
    static constexpr int BARS = RED + 32 * 4;                              // u64 full[NSTAGES], xbar[2]
    static constexpr int FLAGS = BARS + (NSTAGES + 2) * 8;                 // u32[4]
    static constexpr int TOTAL = FLAGS + 16;
    static_assert(TOTAL <= 227 * 1024, "shared-memory layout exceeds the 227 KB opt-in limit");
};

// ------------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1,
                                            uint32_t bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%2, %3}], [%4], %5;"
        ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(bar), "l"(policy) : "memory");
}
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar,
                                             uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1], %2, [%3], %4;"
        ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy) : "memory");
}
// Four ARBITRARY rows of a 2-D tensor with one TMA request (sm_100 tile::gather4): the map's box is {cols, 1}, the four
// rows land back to back in shared memory.  This is what a page-size-1 KV gather wants: 8 requests per 8 KB stage instead
// of 32 row-sized bulk copies.
__device__ __forceinline__ void tma_gather4_2d(uint32_t dst, const CUtensorMap* tm, int c0, int r0, int r1, int r2, int r3,
                                               uint32_t bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%2, %3, %4, %5, %6}], [%7], %8;"
        ::"r"(dst), "l"(tm), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar), "l"(policy) : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
#ifdef CF_EXPERIMENT_EVICT_NORMAL      /* tools/sweep_build.sh experiment, never defined in the product build */
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
#else
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
#endif
    return p;
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* tm) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}
__device__ __forceinline__ void red_add_v4(float* addr, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                 ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// ---- flag-in-data ("LL") words: low 32 bits = float payload, high 32 bits = epoch of the launch that wrote it.
//      A 64-bit scalar access is single-copy atomic, so a reader that sees the right epoch also sees the payload:
//      no fence, no counter, nothing to re-zero -- one L2 round trip per hop. ----
__device__ __forceinline__ unsigned long long ll_pack(float v, unsigned flag) {
    return (unsigned long long)__float_as_uint(v) | ((unsigned long long)flag << 32);
}
__device__ __forceinline__ void ll_store(unsigned long long* p, float v, unsigned flag) {
    asm volatile("st.relaxed.gpu.global.b64 [%0], %1;" ::"l"(p), "l"(ll_pack(v, flag)) : "memory");
}
__device__ __forceinline__ void ll_store2(unsigned long long* p, float v0, float v1, unsigned flag) {   // p 16-byte aligned
    asm volatile("st.relaxed.gpu.global.v2.b64 [%0], {%1, %2};" ::"l"(p), "l"(ll_pack(v0, flag)), "l"(ll_pack(v1, flag)) : "memory");
}
__device__ __forceinline__ unsigned long long ll_load(const unsigned long long* p) {
    unsigned long long w;
    asm volatile("ld.relaxed.gpu.global.b64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
    return w;
}
// two adjacent words with one 16-byte request (p 16-byte aligned); each word still carries and is validated by its own epoch
__device__ __forceinline__ void ll_load2(const unsigned long long* p, unsigned long long& w0, unsigned long long& w1) {
    asm volatile("ld.relaxed.gpu.global.v2.b64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(p) : "memory");
}
// spin until the word carries this launch's epoch (the first probe `w` was issued by the caller, batched with others).
// The spin is bounded (~1 s of dependent L2 round trips): a publisher that never arrives -- CTAs of a group not
// co-resident, a workspace shared by two streams -- turns into an error word in the workspace header (`err`, header[2])
// and a wrong result instead of a hung GPU.
__device__ __forceinline__ float ll_resolve(const unsigned long long* p, unsigned long long w, unsigned flag, unsigned* err) {
    unsigned spins = 0;
    while ((unsigned)(w >> 32) != flag) {
        w = ll_load(p);
        if (++spins > (1u << 22)) { *err = 1u; break; }
    }
    return __uint_as_float((unsigned)w);
}

// Flag of a launch epoch: never 0 (a zero-filled workspace must not match), different for any two CONSECUTIVE epochs including
// across the 2^32 wrap of the counter (0xFFFFFFFF -> 0), and its low bit alternates strictly (the peer stage double-buffers on it).
__device__ __forceinline__ unsigned ll_flag_of_epoch(unsigned epoch) { return 0x80000000u | (epoch & 0x7fffffffu); }

// ---- the same words at SYSTEM scope: peer GPUs' memory over NVLink (an aligned 64-bit store is one transaction) ----
__device__ __forceinline__ void ll_store_sys(unsigned long long* p, float v, unsigned flag) {
    asm volatile("st.relaxed.sys.global.b64 [%0], %1;" ::"l"(p), "l"(ll_pack(v, flag)) : "memory");
}
__device__ __forceinline__ unsigned long long ll_load_sys(const unsigned long long* p) {
    unsigned long long w;
    asm volatile("ld.relaxed.sys.global.b64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
    return w;
}

// Fused all-reduce stage for head-parallel shards.  `v` is this rank's fp32 partial of output column `col` (already summed
// over the rank's own clusters).  The rank pushes it as a (value, epoch) word into slot [parity][tp_rank][col] of EVERY
// rank's exchange buffer (its own included) with one 8-byte store per peer over NVLink, then polls its own buffer until
// all tp_world slots of the column carry this launch's epoch and adds them in rank order: every rank computes the
// bit-identical sum, with no NCCL launch, no extra kernel and no separate fp32 -> fp16 pass (one collective per layer is
// the whole communication of the 70B config; NCCL's latency for 32 KB is several times this kernel's tail).
// Double-buffered on the epoch parity: a rank can be at most one launch ahead of its slowest peer (it needs that peer's
// words of launch L+1 to finish L+1, and the peer publishes them only after it has read everything of launch L).
// The poll is bounded (about a second) so that a desynchronised peer shows up as a NaN + an error word, not as a hang.
// `flag` comes from the peer stage's own launch counter (header[3], read at kernel start and bumped only by launches that
// use the stage), so launches without a peer stage on the same workspace cannot break the strict parity alternation.
__device__ __forceinline__ float tp_allreduce_column(const KParams& p, float v, int col, unsigned flag) {
    const size_t par = (size_t)(flag & 1u) * p.tp_world;
    for (int r = 0; r < p.tp_world; ++r)
        ll_store_sys(p.tp_peer[r] + (par + p.tp_rank) * p.hidden + col, v, flag);
    const unsigned long long* mine = p.tp_peer[p.tp_rank] + par * p.hidden + col;
    float acc = 0.f;
    for (int r = 0; r < p.tp_world; ++r) {
        unsigned long long w = ll_load_sys(mine + (size_t)r * p.hidden);
        unsigned spins = 0;
        while ((unsigned)(w >> 32) != flag) {
            if (++spins > (1u << 23)) { p.header[2] = 1u; w = ll_pack(__int_as_float(0x7fc00000), flag); break; }
            __nanosleep(64);
            w = ll_load_sys(mine + (size_t)r * p.hidden);
        }
        acc += __uint_as_float((unsigned)w);
    }
    return acc;
}

// Cross-cluster sum of per-cluster partial output slices without atomics (batch == 1 launches).
// Cluster `cid` of `ncl` has published slice words out_ll[cid][slice0 .. slice0+n).  This CTA finalises columns
// [lo, hi) of the slice (the ncl CTAs that share a slice split its columns), summing the ncl partials of a column in
// cluster order -> the result is bit-identical from launch to launch.  4 threads per column poll ncl/4 words each.
// With head-parallel shards (p.tp_world > 1) the column then goes through tp_allreduce_column before it is written.
__device__ __forceinline__ void ll_finalize_columns(const KParams& p, const unsigned long long* out_ll, int hidden, int ncl,
                                                    int slice0, int lo, int hi, unsigned flag, unsigned tp_flag, void* out,
                                                    bool fp32_out, uint32_t tid, int nthreads) {
    const int ncols = hi - lo;
    for (int i = tid; i < ((ncols * 4 + 31) & ~31); i += nthreads) {
        const int col = i >> 2, sub = i & 3;
        const bool active = col < ncols;
        const int c0 = sub * ncl / 4, c1 = (sub + 1) * ncl / 4;
        const unsigned long long* base = out_ll + slice0 + lo + (active ? col : 0);
        float acc = 0.f;
        if (active) {
            for (int c = c0; c < c1; c += 8) {
                unsigned long long w[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) w[j] = (c + j < c1) ? ll_load(base + (size_t)(c + j) * hidden) : 0ull;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (c + j < c1) acc += ll_resolve(base + (size_t)(c + j) * hidden, w[j], flag, p.header + 2);
            }
        }
        // fixed-order combine of the 4 sub-sums: (s0 + s1) + (s2 + s3)
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        if (active && sub == 0) {
            const int o = slice0 + lo + col;
            if (p.tp_world > 1) acc = tp_allreduce_column(p, acc, o, tp_flag);
            if (fp32_out) static_cast<float*>(out)[o] = acc;
            else static_cast<__half*>(out)[o] = __float2half_rn(acc);
        }
    }
}

__device__ __forceinline__ float4 ld_cg_v4(const float* addr) {
    float4 v;
    asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
    const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 t = __half22float2(h[i]);
        f[2 * i] = t.x;
        f[2 * i + 1] = t.y;
    }
}
__device__ __forceinline__ float round_h(float v) { return __half2float(__float2half_rn(v)); }

// ------------------------------------------------------------------------------------------------
// Tensor-core helpers (mma.sync.m16n8k16, fp16 in / fp32 accumulate).  Weight tiles land as 128-byte-swizzled TMA boxes
// of [rows][64 halves]: the 16-byte chunk c of row r sits at r*128 + ((c ^ (r & 7)) << 4), which makes the eight row
// addresses of every 8x8 ldmatrix hit eight different bank groups.  The GEMVs put the weights on the A operand (M = 16
// output rows or columns per step) and the activation vector(s) on the N = 8 dimension; at batch 1 only column 0 is
// real -- the tensor cores are idle anyway and this replaces ~17 LDS/CVT/FFMA per 16 x 8 weight block by one ldmatrix +
// one mma, which matters where a phase is consumption-bound (O projection: its tiles were prefetched during the softmax
// exchange) rather than HBM-bound.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// A fragment (m0..m0+15, 16 k) from a swizzled block whose shared-memory rows are the M rows ([m][k], nn.Linear tiles);
// kchunk0 = first 16-byte chunk of the k-step inside the 64-half row
__device__ __forceinline__ void ldsm_a_mrows(uint32_t (&a)[4], uint32_t block, int m0, int kchunk0, uint32_t lane) {
    const int lmat = lane >> 3, row = m0 + (lmat & 1) * 8 + (lane & 7), chunk = kchunk0 + (lmat >> 1);
    const uint32_t addr = block + row * 128 + ((chunk ^ (row & 7)) << 4);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(addr));
}
// A fragment from a swizzled block whose shared-memory rows are the K rows ([k][m], transposed "chat" tiles): k0 = first
// k row of the step, mchunk0 = first 16-byte chunk of the 16 output columns inside the 64-half row
__device__ __forceinline__ void ldsm_a_krows(uint32_t (&a)[4], uint32_t block, int k0, int mchunk0, uint32_t lane) {
    const int lmat = lane >> 3, row = k0 + (lmat >> 1) * 8 + (lane & 7), chunk = mchunk0 + (lmat & 1);
    const uint32_t addr = block + row * 128 + ((chunk ^ (row & 7)) << 4);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(addr));
}
// B fragment of a single activation vector on column n = 0: {v[k0 + 2*t4], v[k0 + 2*t4 + 1]} for lanes 0-3, zero elsewhere
__device__ __forceinline__ uint32_t bfrag_col0(const float* v, int k0, uint32_t lane) {
    if (lane >= 4) return 0u;
    const __half2 h2 = __floats2half2_rn(v[k0 + 2 * lane], v[k0 + 2 * lane + 1]);
    return *reinterpret_cast<const uint32_t*>(&h2);
}

// Optional phase timeline (build with -DCF_TRACE; tools/trace_timeline.py).  Never compiled into the product.
#ifdef CF_TRACE
__device__ unsigned long long* g_cf_trace = nullptr;     // [grid][16] globaltimer ns, set by cf_debug_set_trace
__device__ __forceinline__ void trace_mark(int slot, int launch_id) {
    if (g_cf_trace && (threadIdx.x & 31) == 0 && (threadIdx.x >> 5) == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_cf_trace[((size_t)launch_id * gridDim.y * gridDim.x + blockIdx.y * gridDim.x + blockIdx.x) * 16 + slot] = t;
    }
}
#define CF_MARK(slot) trace_mark(slot, p.layer_id)   /* chat/sglang: layer_id is free to carry a launch index */
#else
#define CF_MARK(slot) ((void)0)
#endif

// Fill one KV stage of the MHA kernels -- 16 K rows (256 B each, head `col0 / 128`) in the first 4 KB, the 16 V rows in the
// second -- from cache / pool rows `slot` (this lane's row is lane & 15; `nvalid` of the 16 rows exist).  Called warp-converged.
// Three ways, all with the same shared-memory layout:
//   tiled   one {128 x 16} TMA box per tensor -- 16 rows in consecutive slots (contiguous cache: always; paged: a sequence
//           that grew without competition), 2 requests per stage;
//   gather4 four arbitrary pool rows per request (sm_100 tile::gather4), 8 requests per stage;
//   rows    one 256-byte bulk copy per row per tensor, 32 requests per stage -- the ragged last tile of a request, and paged
//           launches whose caller did not pass the pool addresses on the host (`maps` false: no tensor map over the pool).
__device__ __forceinline__ void issue_kv_stage(const KParams& p, bool maps, bool contiguous, uint32_t dst, uint32_t fb, int col0,
                                               long long slot, int nvalid, const __half* kbase, const __half* vbase, int kv_cols,
                                               uint32_t lane, uint64_t pol) {
    bool tiled = maps && nvalid == ROWS512, gather = false;
    if (tiled && !contiguous) {
        // the pool maps cover POOL_MAP_ROWS slots (the pool's real size is not part of the interface): a tile that names a
        // slot beyond them -- a > 100 GB pool -- takes the row-copy path, which addresses the pool directly
        const long long slot0 = __shfl_sync(0xffffffffu, slot, 0);
        const bool run = __all_sync(0xffffffffu, slot == slot0 + (long long)(lane & 15));
        const bool in_map = __all_sync(0xffffffffu, slot < (long long)POOL_MAP_ROWS);
        gather = in_map && !run;
        tiled = in_map && run;
    }
    if (tiled) {
        if (lane == 0) {
            dsm::mbar_arrive_expect_tx(fb, STAGE_BYTES);
            tma_load_2d(dst, &p.tm_k, col0, (int)slot, fb, pol);
            tma_load_2d(dst + STAGE_BYTES / 2, &p.tm_v, col0, (int)slot, fb, pol);
        }
    } else if (gather) {
        const int s1 = (int)__shfl_down_sync(0xffffffffu, slot, 1);
        const int s2 = (int)__shfl_down_sync(0xffffffffu, slot, 2);
        const int s3 = (int)__shfl_down_sync(0xffffffffu, slot, 3);
        if (lane == 0) dsm::mbar_arrive_expect_tx(fb, STAGE_BYTES);
        __syncwarp();
        if ((lane & 3) == 0) {                          // lanes 0,4,8,12: K rows 4q..4q+3; lanes 16,..,28: V rows
            const uint32_t d = dst + (lane & 15) * (HEAD_DIM * 2) + (lane < 16 ? 0 : STAGE_BYTES / 2);
            tma_gather4_2d(d, lane < 16 ? &p.tm_kg : &p.tm_vg, col0, (int)slot, s1, s2, s3, fb, pol);
        }
    } else {
        if (lane == 0) dsm::mbar_arrive_expect_tx(fb, nvalid * 2 * HEAD_DIM * 2);
        __syncwarp();
        if ((int)(lane & 15) < nvalid) {
            const uint32_t d = dst + (lane & 15) * (HEAD_DIM * 2);
            if (lane < 16) bulk_load_1d(d, kbase + slot * kv_cols + col0, HEAD_DIM * 2, fb, pol);
            else bulk_load_1d(d + STAGE_BYTES / 2, vbase + slot * kv_cols + col0, HEAD_DIM * 2, fb, pol);
        }
    }
}

// ring bookkeeping: global tile index g -> (stage, parity)
__device__ __forceinline__ uint32_t ring_stage(uint32_t g) { return g % NSTAGES; }
__device__ __forceinline__ uint32_t ring_parity(uint32_t g) { return (g / NSTAGES) & 1u; }

// Ring discipline.  Global tile index g -> stage g % NSTAGES, owned by warp g % CONSUMER_WARPS, and
// NSTAGES == 2 * CONSUMER_WARPS: a stage always belongs to the same warp, which both consumes it and re-fills it
// (tile g + NSTAGES) -- while it consumes one of its two stages the other is landing, so a warp's cycle is
// max(consume, consume/2 + memory latency) instead of consume + latency.  The affinity is also what makes the
// one-bit phase parity sufficient: a warp cannot test use u of its stage before it has itself consumed use
// u-1, so the parity can never alias to an older phase (it would if consecutive uses of a stage were
// consumed by different warps -- a warp could then look at the barrier before the previous use's TMA landed).
__device__ __forceinline__ void ring_wait_full(uint32_t full_u32, uint32_t g) {
    dsm::mbar_wait(full_u32 + 8 * ring_stage(g), ring_parity(g));
}
// first tile index i >= 0 of a phase starting at global index gbase that belongs to `warp`
__device__ __forceinline__ uint32_t first_tile(uint32_t gbase, uint32_t warp) {
    return (warp + CONSUMER_WARPS - gbase % CONSUMER_WARPS) % CONSUMER_WARPS;
}

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
template <int VARIANT, int CLUSTER>
__global__ void __launch_bounds__(BLOCK_THREADS, 1)
llama_decoder_layer_kernel(const __grid_constant__ KParams p)
{
    using S = Smem<CLUSTER>;
    constexpr bool kChat = (VARIANT == CHAT);
    constexpr bool kPaged = (VARIANT == PAGED);

    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t smem_base = dsm::smem_u32(smem);
    const uint32_t tid = threadIdx.x;
    const uint32_t warp = tid >> 5;
    const uint32_t lane = tid & 31;
    const uint32_t rank = dsm::cluster_ctarank();
    const uint32_t head = blockIdx.x / CLUSTER;
    const uint32_t batch = blockIdx.y;

    const int hidden = p.hidden;
    const int KS = hidden / CLUSTER;     // this CTA's slice of the GEMV reduction dim / of the O output dim
    const int kv_cols = p.n_kv_heads * HEAD_DIM;

    const uint32_t full_u32 = smem_base + S::BARS;          // u64 full[NSTAGES]
    const uint32_t xbar_u32 = full_u32 + NSTAGES * 8;        // u64 xbar[2]

    const uint32_t n_qkv_tiles = kChat ? 3u * (KS / ROWS256) : (uint32_t)(S::QKV_OUT / ROWS256) * (KS / 128);     // >= 24
    const uint32_t n_o_tiles = kChat ? (uint32_t)(HEAD_DIM / ROWS256) * (KS / 128) : (uint32_t)(KS / ROWS256);

    CF_MARK(0);   // kernel entry
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // no-op unless the NEXT launch opted into PDL

    // ---- tile stream ----------------------------------------------------------------------------------
    const uint64_t pol = policy_evict_first();
    // QKV weight tile g (< n_qkv_tiles) into its stage; elected lane only.  Needs nothing but the CTA's coordinates.
    auto issue_qkv_tile = [&](uint32_t g) {
        if (lane != 0) return;
        const uint32_t s = ring_stage(g);
        const uint32_t fb = full_u32 + 8 * s;
        const uint32_t dst = smem_base + S::RING + s * STAGE_BYTES;
        const uint32_t i = g;
        int c0, c1;
        if constexpr (kChat) {           // tile i: matrix j = i % 3, rows t = i / 3 (32 input rows x 128 output cols)
            const int j = i % 3, t = i / 3;
            c0 = head * HEAD_DIM;
            c1 = j * hidden + rank * KS + t * ROWS256;
        } else {                         // tile i: 32 output rows (block rb = i % 12 of the head's 384 q|k|v rows) x 128
            const int rb = i % CONSUMER_WARPS, win = i / CONSUMER_WARPS;     // input cols (window win): window-major, so
            const int j = rb / (HEAD_DIM / ROWS256), sub = rb % (HEAD_DIM / ROWS256);   // warp w only ever sees block w
            const int row0 = (j == 0) ? head * HEAD_DIM
                           : (j == 1) ? p.n_heads * HEAD_DIM + head * HEAD_DIM
                                      : (p.n_heads + p.n_kv_heads) * HEAD_DIM + head * HEAD_DIM;
            c0 = rank * KS + win * 128;
            c1 = row0 + sub * ROWS256;
        }
        dsm::mbar_arrive_expect_tx(fb, STAGE_BYTES);           // two [32 rows x 64 cols] swizzled boxes, either layout
        tma_load_2d(dst, &p.tm_wqkv, c0, c1, fb, pol);
        tma_load_2d(dst + 4096, &p.tm_wqkv, c0 + 64, c1, fb, pol);
    };

    // ---- barrier init + first two tiles of every warp ---------------------------------------------------
    // Each warp initialises the full barriers of its own two stages and requests its first two tiles at once (always QKV
    // weight tiles: n_qkv_tiles >= 24): nothing here depends on the other warps, on the peer CTAs, on the previous kernel in
    // the stream (programmatic dependent launch) or -- paged form -- on the request's page table, whose two dependent index
    // loads would otherwise sit in front of the first TMA request of every layer (round 2: they now overlap it).
    // The exchange barriers are armed by thread 0 before the cluster-wide arrive.
    if (lane == 0) {
        dsm::mbar_init(full_u32 + 8 * warp, 1);
        dsm::mbar_init(full_u32 + 8 * (warp + CONSUMER_WARPS), 1);
        if (tid == 0) {
            prefetch_tmap(&p.tm_wqkv);
            prefetch_tmap(&p.tm_wo);
            cluster_reduce_arm<CLUSTER>(xbar_u32, S::QKV_OUT * 4);
            cluster_reduce_arm<CLUSTER>(xbar_u32 + 8, S::ATTN_PAYLOAD * 4);
        }
        dsm::mbar_fence_init();
    }
    __syncwarp();
    CF_MARK(12);  // first TMA issue
    issue_qkv_tile(warp);
    issue_qkv_tile(warp + CONSUMER_WARPS);
    dsm::cluster_arrive();     // peers may push into this CTA's smem only after every CTA armed its barriers

    // ---- per-request KV range -----------------------------------------------------------------
    int kv_len, kv_base = 0, new_slot = 0;
    if constexpr (kPaged) {
        kv_base = p.indptr[batch];
        const int end = p.indptr[batch + 1] - 1;      // last index = slot of the new token
        kv_len = end - kv_base;
        new_slot = p.indices[end];
    } else {
        kv_len = p.kv_len;
    }
    const int chunk = (((kv_len + CLUSTER - 1) / CLUSTER) + ROWS512 - 1) & ~(ROWS512 - 1);   // KV rows per CTA, tile aligned
    const int row_begin = min((int)rank * chunk, kv_len);
    const int row_end = min(row_begin + chunk, kv_len);
    const uint32_t n_kv_tiles = (row_end - row_begin + ROWS512 - 1) / ROWS512;
    const uint32_t total_tiles = n_qkv_tiles + n_kv_tiles + n_o_tiles;
    const __half* kpool = p.k_base;      // contiguous forms: the cache itself
    const __half* vpool = p.v_base;
    if constexpr (kPaged) {
        kpool = reinterpret_cast<const __half*>(p.k_pool_ptrs[p.layer_id]);
        vpool = reinterpret_cast<const __half*>(p.v_pool_ptrs[p.layer_id]);
    }
    // Paged: the host may pass its own copy of the two pool addresses together with tensor maps over the pools.  They are
    // used only if they agree with what the device-side pointer table says NOW (a stale host copy just disables the fast
    // paths: correctness never depends on it).
    const bool pool_maps = !kPaged || (p.k_base != nullptr && kpool == p.k_base && vpool == p.v_base);
    // Paged KV: the page index of this lane's row of KV tile g.  A warp fetches it one ring cycle ahead (when it issues
    // tile g - 24 into the same stage), so the index load is never on the path between consuming a tile and refilling
    // its stage (the reference gathers with an index load per row on the critical path, kernel_batch_sglang.cuh:356-371;
    // measured here: 3 us per layer at kv 1K before the prefetch).
    int pre_slot0 = 0, pre_slot1 = 0;
    uint32_t pre_g0 = 0xffffffffu, pre_g1 = 0xffffffffu;
    auto page_of = [&](uint32_t g) -> int {
        const int r = row_begin + (int)(g - n_qkv_tiles) * ROWS512 + (int)(lane & 15);
        return (r < row_end) ? p.indices[kv_base + r] : 0;
    };
    // Request global tile g into its stage.  Called warp-converged by the owner warp (g % 12 == warp), after it has
    // finished reading the stage (tile g - 24).  Weights and tiled KV: one elected lane; paged KV: one row per lane.
    auto issue_tile = [&](uint32_t g) {
        if (g >= total_tiles) return;
        const uint32_t s = ring_stage(g);
        const uint32_t fb = full_u32 + 8 * s;
        const uint32_t dst = smem_base + S::RING + s * STAGE_BYTES;
        if (g < n_qkv_tiles) {
            issue_qkv_tile(g);
        } else if (g < n_qkv_tiles + n_kv_tiles) {
            const uint32_t i = g - n_qkv_tiles;                 // 16 KV rows: K in the first 4 KB of the stage, V in the second
            const int r0 = row_begin + (int)i * ROWS512;
            const int nvalid = min(ROWS512, row_end - r0);
            long long slot = r0 + (int)(lane & 15);             // contiguous cache: the row index itself
            if constexpr (kPaged) {
                const bool odd = (g / CONSUMER_WARPS) & 1u;
                slot = (odd ? pre_g1 : pre_g0) == g ? (long long)(odd ? pre_slot1 : pre_slot0) : (long long)page_of(g);
            }
            issue_kv_stage(p, pool_maps, !kPaged, dst, fb, head * HEAD_DIM, slot, nvalid, kpool, vpool, kv_cols, lane, pol);
        } else {
            if (lane == 0) {
                const uint32_t i = g - n_qkv_tiles - n_kv_tiles;
                int c0, c1;
                if constexpr (kChat) {           // Wo^T [in][out]: 32 input rows x 128 output cols
                    c0 = rank * KS + (i >> 2) * 128;
                    c1 = head * HEAD_DIM + (i & 3) * ROWS256;
                } else {                         // Wo [out][in]: 32 output rows x this head's 128 input cols
                    c0 = head * HEAD_DIM;
                    c1 = rank * KS + i * ROWS256;
                }
                dsm::mbar_arrive_expect_tx(fb, STAGE_BYTES);           // two [32 rows x 64 cols] swizzled boxes
                tma_load_2d(dst, &p.tm_wo, c0, c1, fb, pol);
                tma_load_2d(dst + 4096, &p.tm_wo, c0 + 64, c1, fb, pol);
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

    if (tid == 0 && pool_maps) {
        prefetch_tmap(&p.tm_k); prefetch_tmap(&p.tm_v);
        if constexpr (kPaged) { prefetch_tmap(&p.tm_kg); prefetch_tmap(&p.tm_vg); }
    }

    // =============================================================================================
    // ALL WARPS
    // =============================================================================================
    float* xs = reinterpret_cast<float*>(smem + S::XS);
    float* qkv_part = reinterpret_cast<float*>(smem + S::QKV_PART);
    float* attn_part = reinterpret_cast<float*>(smem + S::ATTN_PART);
    float* out_part = reinterpret_cast<float*>(smem + S::OUT_PART);
    float* qkv_src = reinterpret_cast<float*>(smem + S::QKV_SRC);
    float* qkv_recv = reinterpret_cast<float*>(smem + S::QKV_RECV);
    float* attn_src = reinterpret_cast<float*>(smem + S::ATTN_SRC);
    float* attn_recv = reinterpret_cast<float*>(smem + S::ATTN_RECV);
    float* qkv_fin = reinterpret_cast<float*>(smem + S::QKV_FINAL);
    float* attn_out = reinterpret_cast<float*>(smem + S::ATTN_OUT);
    float* red = reinterpret_cast<float*>(smem + S::RED);
    uint32_t* sflags = reinterpret_cast<uint32_t*>(smem + S::FLAGS);

    const __half* xg = p.x + (size_t)batch * hidden;
    const __half* rg = kChat ? nullptr : p.residual_in + (size_t)batch * hidden;
    __half* rout = kChat ? nullptr : p.residual_out + (size_t)batch * hidden;
    const bool residual_inplace = !kChat && (static_cast<const void*>(rout) == static_cast<const void*>(rg));

    // The RMSNorm weight slice does not depend on the previous kernel: fetch it before the dependency wait.
    const int cpk = KS / 8;                    // 8-element chunks in this CTA's slice (<= 256 < CONSUMER_THREADS)
    const int nchunks = hidden / 8;
    uint4 wraw = make_uint4(0, 0, 0, 0);
    if ((int)tid < cpk) wraw = *reinterpret_cast<const uint4*>(p.rms_w + rank * KS + tid * 8);
    // Neither do the RoPE factors of this thread's q / k element (nor `positions`, by the PDL contract): fetched here they cost
    // nothing later, fetched in the RoPE step they put one (paged form: two dependent) L2 round trips right behind exchange 1.
    float rope_c = 0.f, rope_s = 0.f;
    if (tid < 2u * HEAD_DIM) {
        const int d = tid & 127;
        if constexpr (kChat) {
            rope_c = p.cos[d]; rope_s = p.sin[d];
        } else if constexpr (kPaged) {
            const float* cs = p.cos + p.positions[batch] * HEAD_DIM;
            rope_c = cs[d & 63]; rope_s = cs[HEAD_DIM / 2 + (d & 63)];
        } else {
            rope_c = p.cos[d & 63]; rope_s = p.sin[d & 63];
        }
    }

    // Programmatic dependent launch: everything above (and the whole producer warp) may run while the previous
    // kernel in the stream is still finishing; activations, outputs and the workspace may only be touched after it.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const bool ll_out = (gridDim.y == 1) && (p.out_ll != nullptr);
    const unsigned flag = ll_out ? ll_flag_of_epoch(__ldcg(p.header)) : 0u;
    const unsigned tp_flag = (ll_out && p.tp_world > 1) ? ll_flag_of_epoch(__ldcg(p.header + 3)) : 0u;

    // ---- phase 0: RMSNorm ---------------------------------------------------------------------------
    // every CTA reduces the full vector itself (8-16 KB from L2) -> no cluster round trip for a scalar.  One pass:
    // chunk c = (j + rank*cpk) mod nchunks goes to thread j mod 384, so the chunks of this CTA's own slice are the
    // FIRST chunk of threads 0..cpk-1 and stay in registers across the block reduction (one L2 round trip, not two).
    {
        constexpr int P0_ITERS = (KS_MAX * 4 / 8 + CONSUMER_THREADS - 1) / CONSUMER_THREADS;     // hidden <= 4*KS_MAX: 3
        uint4 xr[P0_ITERS], rr[P0_ITERS];
#pragma unroll
        for (int it = 0; it < P0_ITERS; ++it) {
            const int j = tid + it * CONSUMER_THREADS;
            const int c = (j + rank * cpk) % nchunks;
            xr[it] = make_uint4(0, 0, 0, 0);
            rr[it] = make_uint4(0, 0, 0, 0);
            if (j < nchunks) {
                xr[it] = *reinterpret_cast<const uint4*>(xg + c * 8);
                if constexpr (!kChat) rr[it] = *reinterpret_cast<const uint4*>(rg + c * 8);
            }
        }
        float f0[8];
        float ss = 0.f;
#pragma unroll
        for (int it = 0; it < P0_ITERS; ++it) {
            float f[8];
            unpack8(xr[it], f);
            if constexpr (!kChat) {
                float r8[8];
                unpack8(rr[it], r8);
#pragma unroll
                for (int k = 0; k < 8; ++k) f[k] = round_h(f[k] + r8[k]);   // residual_out is fp16; norm the rounded sum
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) ss += f[k] * f[k];                  // chunks past the end are zeros
            if (it == 0) {
#pragma unroll
                for (int k = 0; k < 8; ++k) f0[k] = f[k];
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        if (lane == 0) red[warp] = ss;
        dsm::named_bar_sync(CONSUMER_BAR, CONSUMER_THREADS);
        float tot = 0.f;
#pragma unroll
        for (int w = 0; w < CONSUMER_WARPS; ++w) tot += red[w];
        const float rstd = rsqrtf(tot / (float)hidden + p.eps);
        // normalised slice [rank*KS, +KS) -> fp32 smem, rounded where the eager fp16 model rounds
        if ((int)tid < cpk) {
            const int e = tid * 8;
            float w8[8];
            unpack8(wraw, w8);
            if constexpr (!kChat) {
                if (head == 0 && !residual_inplace) {
                    __align__(16) __half hs[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) hs[k] = __float2half_rn(f0[k]);     // exact: f0 is already fp16-rounded
                    *reinterpret_cast<uint4*>(rout + rank * KS + e) = *reinterpret_cast<const uint4*>(hs);
                }
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) xs[e + k] = round_h(round_h(f0[k] * rstd) * w8[k]);
        }
        dsm::named_bar_sync(CONSUMER_BAR, CONSUMER_THREADS);
    }

    CF_MARK(1);   // RMSNorm done
    uint32_t gbase = 0;   // global ring index of tile 0 of the current phase

    // ---- phase 1: QKV GEMV over this CTA's K-slice ----------------------------------------------------
    if constexpr (kChat) {
        // tile i = 32 input rows (t = i / 3) x 128 output cols of matrix j = i % 3.  The QKV phase starts at
        // ring index 0 and 12 % 3 == 0, so warp w only ever sees matrix j = w % 3: its 8 column sums stay in
        // registers for the whole phase.  lane (sub, c): rows sub + 2s, cols c*8 .. c*8+7.
        // Tensor cores: D[128 cols][8] += tile^T[128 x 32] * x[32 x 8] (x on column 0): 8 m-blocks x 2 k-steps per tile.
        const int g4 = lane >> 2, t4 = lane & 3;
        float acc[8][4];
#pragma unroll
        for (int mb = 0; mb < 8; ++mb) { acc[mb][0] = 0.f; acc[mb][1] = 0.f; acc[mb][2] = 0.f; acc[mb][3] = 0.f; }
        for (uint32_t i = first_tile(gbase, warp); i < n_qkv_tiles; i += CONSUMER_WARPS) {
            const uint32_t g = gbase + i, s = ring_stage(g);
            const float* xrow = xs + (i / 3) * ROWS256;
            uint32_t xb[2][2];
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) { xb[ks][0] = bfrag_col0(xrow, ks * 16, lane); xb[ks][1] = bfrag_col0(xrow, ks * 16 + 8, lane); }
            ring_wait_full(full_u32, g);
            const uint32_t st = smem_base + S::RING + s * STAGE_BYTES;
#pragma unroll
            for (int mb = 0; mb < 8; ++mb) {
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    uint32_t af[4];
                    ldsm_a_krows(af, st + (mb >> 2) * 4096, ks * 16, (mb & 3) * 2, lane);
                    mma16816(acc[mb], af, xb[ks][0], xb[ks][1]);
                }
            }
            __syncwarp();
            issue_tile(g + NSTAGES);          // stage is free again: request the tile that will live in it next
        }
        if (t4 == 0) {       // C fragment column 0: rows g4 and g4 + 8 of each 16-column block; write-once slot [warp][128]
#pragma unroll
            for (int mb = 0; mb < 8; ++mb) {
                qkv_part[warp * HEAD_DIM + mb * 16 + g4] = acc[mb][0];
                qkv_part[warp * HEAD_DIM + mb * 16 + g4 + 8] = acc[mb][2];
            }
        }
    } else {
        // tile = 32 output rows x 128 input cols as two swizzled [32 x 64] boxes (half the TMA requests of round 1's four
        // [16 x 64] boxes per stage).  The tiles are dealt window-major and 12 row blocks = 12 warps, so warp w only ever sees
        // row block w of the head's 384 q|k|v rows: its 32 row sums stay in registers for the whole phase and go straight
        // to the exchange buffer -- no per-window partial slots, no fold pass.  Tensor cores: D[32 rows][8] += tile[32 x 128]
        // * x[128 x 8] (x on column 0): two m-blocks x 8 k-steps per tile.
        static_assert(S::QKV_OUT / ROWS256 == CONSUMER_WARPS, "one 32-row block of q|k|v per warp");
        const int g4 = lane >> 2, t4 = lane & 3;
        float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
        for (uint32_t i = first_tile(gbase, warp); i < n_qkv_tiles; i += CONSUMER_WARPS) {
            const uint32_t g = gbase + i, s = ring_stage(g);
            const float* xw = xs + (i / CONSUMER_WARPS) * 128;
            uint32_t xb[8][2];
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) { xb[ks][0] = bfrag_col0(xw, ks * 16, lane); xb[ks][1] = bfrag_col0(xw, ks * 16 + 8, lane); }
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
            issue_tile(g + NSTAGES);          // stage is free again: request the tile that will live in it next
        }
        if (t4 == 0) {                        // C fragment column 0: rows g4 and g4 + 8 of each 16-row block
#pragma unroll
            for (int mb = 0; mb < 2; ++mb) {
                qkv_src[warp * ROWS256 + mb * 16 + g4] = acc[mb][0];
                qkv_src[warp * ROWS256 + mb * 16 + g4 + 8] = acc[mb][2];
            }
        }
    }
    gbase += n_qkv_tiles;
    dsm::named_bar_sync(CONSUMER_BAR, CONSUMER_THREADS);
    CF_MARK(2);   // QKV tiles consumed (this warp)

    // chat layout: fold the four write-once slots of a matrix in a fixed order -> this CTA's partial q|k|v
    // (nn.Linear layout: every warp already wrote its 32 complete row sums to qkv_src)
    if constexpr (kChat) {
        for (int o = tid; o < S::QKV_OUT; o += CONSUMER_THREADS) {
            float a = 0.f;
            for (int w = o >> 7; w < CONSUMER_WARPS; w += 3) a += qkv_part[w * HEAD_DIM + (o & 127)];   // warps j, j+3, j+6, j+9
            qkv_src[o] = a;
        }
    }
    // first remote access of the kernel: all CTAs of the cluster have armed their barriers by now
    dsm::cluster_wait();
    uint32_t xphase0 = 0, xphase1 = 0;
    cluster_reduce<CLUSTER, Stage::LINEAR, CONSUMER_THREADS, CONSUMER_BAR>(
        S::QKV_OUT * 4, tid, HEAD_DIM, rank,
        smem_base + S::QKV_SRC, smem_base + S::QKV_RECV, xbar_u32, xphase0, qkv_src, qkv_recv);

    CF_MARK(3);   // QKV cluster exchange done
    // ---- RoPE (fp16 rounding points of the eager model), new K/V out -------------------------------
    {
        constexpr float kScaleLog2 = 0.08838834764831845f * 1.4426950408889634f;   // 1/sqrt(128) * log2(e)
        if (tid < 2 * HEAD_DIM) {
            const int which = tid >> 7, d = tid & 127;          // 0: q, 1: k
            const float* v = qkv_src + which * HEAD_DIM;
            const float a = round_h(v[d]);
            float rot;
            if constexpr (kChat) {                              // GPT-J pairs (2i, 2i+1), cos/sin pair-repeated
                const float b = round_h(v[d ^ 1]);
                rot = (d & 1) ? fmaf(a, rope_c, b * rope_s) : fmaf(a, rope_c, -b * rope_s);
            } else {                                            // NeoX pairs (i, i+64), cos/sin [64]
                const float b = round_h(v[d ^ 64]);
                rot = (d & 64) ? fmaf(a, rope_c, b * rope_s) : fmaf(a, rope_c, -b * rope_s);
            }
            const __half rh = __float2half_rn(rot);
            qkv_fin[tid] = which == 0 ? __half2float(rh) * kScaleLog2 : __half2float(rh);
            if (which == 1 && rank == 0) {
                if constexpr (kPaged) {
                    __half* kpool = reinterpret_cast<__half*>(p.k_pool_ptrs[p.layer_id]);
                    kpool[(size_t)new_slot * kv_cols + head * HEAD_DIM + d] = rh;
                } else {
                    p.k_new[head * HEAD_DIM + d] = rh;
                }
            }
        } else if (tid < 3 * HEAD_DIM) {
            const int d = tid - 2 * HEAD_DIM;
            const __half vh = __float2half_rn(qkv_src[2 * HEAD_DIM + d]);
            qkv_fin[2 * HEAD_DIM + d] = __half2float(vh);
            if (rank == 0) {
                if constexpr (kPaged) {
                    __half* vpool = reinterpret_cast<__half*>(p.v_pool_ptrs[p.layer_id]);
                    vpool[(size_t)new_slot * kv_cols + head * HEAD_DIM + d] = vh;
                } else {
                    p.v_new[head * HEAD_DIM + d] = vh;
                }
            }
        }
        dsm::named_bar_sync(CONSUMER_BAR, CONSUMER_THREADS);
    }

    CF_MARK(4);   // RoPE done
    // ---- phase 2: flash-decode over this CTA's KV rows -----------------------------------------------
    {
        const int sub = lane >> 4, c = lane & 15;
        float q8[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) q8[k] = qkv_fin[c * 8 + k];
        float m = -INFINITY, l = 0.f;
        float o8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (uint32_t i = first_tile(gbase, warp); i < n_kv_tiles; i += CONSUMER_WARPS) {
            const uint32_t g = gbase + i, s = ring_stage(g);
            ring_wait_full(full_u32, g);
            const uint4* kt = reinterpret_cast<const uint4*>(smem + S::RING + s * STAGE_BYTES);
            const uint4* vt = kt + STAGE_BYTES / 32;
            const int rows_left = row_end - (row_begin + (int)i * ROWS512);     // >= 1
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
                if (row >= rows_left) raw = make_uint4(0, 0, 0, 0);       // rows past the end were never copied (stale smem)
                float v8[8];
                unpack8(raw, v8);
#pragma unroll
                for (int k = 0; k < 8; ++k) o8[k] = fmaf(pr, v8[k], o8[k]);
            }
            m = m_new;
            __syncwarp();
            issue_tile(g + NSTAGES);          // stage is free again: request the tile that will live in it next
        }
        gbase += n_kv_tiles;
        CF_MARK(5);   // KV tiles consumed (this warp)
        // per half-warp state -> smem group slot (16 groups)
        {
            const int grp = warp * 2 + sub;                    // 2 * CONSUMER_WARPS groups
            float* slot = attn_part + grp * S::ATTN_PAYLOAD;
            if (c == 0) { slot[0] = m; slot[1] = l; }
            *reinterpret_cast<float4*>(slot + 4 + c * 8) = make_float4(o8[0], o8[1], o8[2], o8[3]);
            *reinterpret_cast<float4*>(slot + 4 + c * 8 + 4) = make_float4(o8[4], o8[5], o8[6], o8[7]);
        }
        // the current token's score (rank 0 folds it in as one more group)
        if (warp == 0) {
            float a = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) a = fmaf(qkv_fin[lane * 4 + k], qkv_fin[HEAD_DIM + lane * 4 + k], a);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            if (lane == 0) red[CONSUMER_WARPS] = a;
        }
        dsm::named_bar_sync(CONSUMER_BAR, CONSUMER_THREADS);
        if (tid < HEAD_DIM) {
            const bool with_new = (rank == 0);
            float M = with_new ? red[CONSUMER_WARPS] : -INFINITY;
#pragma unroll
            for (int gI = 0; gI < 2 * CONSUMER_WARPS; ++gI) M = fmaxf(M, attn_part[gI * S::ATTN_PAYLOAD]);
            float L = 0.f, O = 0.f;
#pragma unroll
            for (int gI = 0; gI < 2 * CONSUMER_WARPS; ++gI) {
                const float w = dsm::exp2_diff(attn_part[gI * S::ATTN_PAYLOAD], M);
                L = fmaf(attn_part[gI * S::ATTN_PAYLOAD + 1], w, L);
                O = fmaf(attn_part[gI * S::ATTN_PAYLOAD + 4 + tid], w, O);
            }
            if (with_new) {
                const float w = dsm::exp2_diff(red[CONSUMER_WARPS], M);
                L += w;
                O = fmaf(qkv_fin[2 * HEAD_DIM + tid], w, O);
            }
            attn_src[4 + tid] = O;
            if (tid == 0) { attn_src[0] = M; attn_src[1] = L; attn_src[2] = 0.f; attn_src[3] = 0.f; }
        }
        cluster_reduce<CLUSTER, Stage::ATTN, CONSUMER_THREADS, CONSUMER_BAR>(
            S::ATTN_PAYLOAD * 4, tid, HEAD_DIM, rank,
            smem_base + S::ATTN_SRC, smem_base + S::ATTN_RECV, xbar_u32 + 8, xphase1, attn_src, attn_recv);
        if (tid < HEAD_DIM) attn_out[tid] = round_h(attn_src[4 + tid] / attn_src[1]);
        dsm::named_bar_sync(CONSUMER_BAR, CONSUMER_THREADS);
    }

    CF_MARK(6);   // softmax merge + exchange done
    // ---- phase 3: O GEMV for output columns [rank*KS, +KS), on the tensor cores (attention output on column 0) ------
    if constexpr (kChat) {
        // tile = 32 input rows (a quarter of the head) x 128 output cols: D[128][8] = tile^T[128 x 32] * a[32 x 8]
        const int g4 = lane >> 2, t4 = lane & 3;
        for (uint32_t i = first_tile(gbase, warp); i < n_o_tiles; i += CONSUMER_WARPS) {
            const uint32_t g = gbase + i, s = ring_stage(g);
            const int cb = i >> 2, rh = i & 3;
            const float* arow = attn_out + rh * ROWS256;
            uint32_t ab[2][2];
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) { ab[ks][0] = bfrag_col0(arow, ks * 16, lane); ab[ks][1] = bfrag_col0(arow, ks * 16 + 8, lane); }
            ring_wait_full(full_u32, g);
            const uint32_t st = smem_base + S::RING + s * STAGE_BYTES;
            float acc[8][4];
#pragma unroll
            for (int mb = 0; mb < 8; ++mb) {
                acc[mb][0] = 0.f; acc[mb][1] = 0.f; acc[mb][2] = 0.f; acc[mb][3] = 0.f;
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    uint32_t af[4];
                    ldsm_a_krows(af, st + (mb >> 2) * 4096, ks * 16, (mb & 3) * 2, lane);
                    mma16816(acc[mb], af, ab[ks][0], ab[ks][1]);
                }
            }
            __syncwarp();
            issue_tile(g + NSTAGES);          // stage is free again: request the tile that will live in it next
            if (t4 == 0) {
#pragma unroll
                for (int mb = 0; mb < 8; ++mb) {
                    out_part[rh * KS + cb * 128 + mb * 16 + g4] = acc[mb][0];
                    out_part[rh * KS + cb * 128 + mb * 16 + g4 + 8] = acc[mb][2];
                }
            }
        }
    } else {
        // tile = 32 output rows x 128 input cols as two swizzled [32 x 64] boxes: D[32][8] = tile[32 x 128] * a[128 x 8]
        const int g4 = lane >> 2, t4 = lane & 3;
        uint32_t ab[8][2];
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) { ab[ks][0] = bfrag_col0(attn_out, ks * 16, lane); ab[ks][1] = bfrag_col0(attn_out, ks * 16 + 8, lane); }
        for (uint32_t i = first_tile(gbase, warp); i < n_o_tiles; i += CONSUMER_WARPS) {
            const uint32_t g = gbase + i, s = ring_stage(g);
            ring_wait_full(full_u32, g);
            const uint32_t st = smem_base + S::RING + s * STAGE_BYTES;
            float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
#pragma unroll
                for (int mb = 0; mb < 2; ++mb) {
                    uint32_t af[4];
                    ldsm_a_mrows(af, st + (ks >> 2) * 4096, mb * 16, (ks & 3) * 2, lane);
                    mma16816(acc[mb], af, ab[ks][0], ab[ks][1]);
                }
            }
            __syncwarp();
            issue_tile(g + NSTAGES);          // stage is free again: request the tile that will live in it next
            if (t4 == 0) {
#pragma unroll
                for (int mb = 0; mb < 2; ++mb) {
                    out_part[i * ROWS256 + mb * 16 + g4] = acc[mb][0];
                    out_part[i * ROWS256 + mb * 16 + g4 + 8] = acc[mb][2];
                }
            }
        }
    }
    dsm::named_bar_sync(CONSUMER_BAR, CONSUMER_THREADS);

    CF_MARK(7);   // O tiles consumed, block-reduced
    if (ll_out) {
        // ---- cross-head reduction, opt-in (CF_FLAG_LL_OUT, batch == 1): every cluster publishes its fp32 partial of this
        //      rank's output slice as (value, epoch) words; the n_heads CTAs that share the slice each sum 1/n_heads of its
        //      columns over all heads in head order and write the result.  No atomics, bitwise reproducible -- but every
        //      CTA then lives until the slowest cluster has published, which delays the next layer's CTAs in a PDL chain
        //      (+2.5 us per layer, measured), so the red path below stays the default. ----
        unsigned long long* mine = p.out_ll + (size_t)head * hidden + rank * KS;
        for (int e = tid * 2; e < KS; e += CONSUMER_THREADS * 2) {
            float2 v = *reinterpret_cast<const float2*>(out_part + e);
            if constexpr (kChat) {
#pragma unroll
                for (int q = 1; q < HEAD_DIM / ROWS256; ++q) {
                    const float2 w = *reinterpret_cast<const float2*>(out_part + q * KS + e);
                    v.x += w.x; v.y += w.y;
                }
            }
            ll_store2(mine + e, v.x, v.y, flag);
        }
        CF_MARK(8);   // partial published
        const int nh = p.n_heads;
        const int lo = (int)((long long)head * KS / nh), hi = (int)((long long)(head + 1) * KS / nh);
        ll_finalize_columns(p, p.out_ll, hidden, nh, rank * KS, lo, hi, flag, tp_flag, p.out, (p.flags & 1u) != 0, tid, CONSUMER_THREADS);
        // Any CTA that got here has seen every cluster's partial, so every CTA of the launch is past phase 0 (it read x,
        // residual and the epoch before its cluster's first exchange): CTA 0 may now bump the epoch for the next launch
        // and, for the in-place form, overwrite `residual` (the reference races here).
        if (blockIdx.x == 0) {
            if (tid == 0) {
                asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(p.header) : "memory");
                if (p.tp_world > 1) asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(p.header + 3) : "memory");
            }
            if constexpr (!kChat) {
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
        }
        CF_MARK(9);   // CTA done
        return;
    }
    // ---- cross-head reduction (default): fp32 red into scratch, last arriver of the slice finalises ---------------
    float* scratch = p.scratch + (size_t)batch * hidden + rank * KS;
    for (int e = tid * 4; e < KS; e += CONSUMER_THREADS * 4) {
        float4 v = *reinterpret_cast<const float4*>(out_part + e);
        if constexpr (kChat) {
#pragma unroll
            for (int q = 1; q < HEAD_DIM / ROWS256; ++q) {
                const float4 w = *reinterpret_cast<const float4*>(out_part + q * KS + e);
                v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
            }
        }
        red_add_v4(scratch + e, v);
    }
    __threadfence();
    dsm::named_bar_sync(CONSUMER_BAR, CONSUMER_THREADS);
    unsigned* counters = p.counters + (size_t)batch * (CLUSTER + 1);
    if (tid == 0) {
        const unsigned prev = atomicAdd(&counters[rank], 1u);
        sflags[0] = (prev == (unsigned)p.n_heads - 1u);
    }
    dsm::named_bar_sync(CONSUMER_BAR, CONSUMER_THREADS);
    CF_MARK(8);   // reds issued, counter bumped
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
        if constexpr (!kChat) {
            if (residual_inplace) {
                // in-place residual update is deferred to the very last CTA of the request: every other
                // CTA has finished reading `residual` by then (the reference races here)
                __threadfence();
                dsm::named_bar_sync(CONSUMER_BAR, CONSUMER_THREADS);
                if (tid == 0) {
                    const unsigned prev = atomicAdd(&counters[CLUSTER], 1u);
                    sflags[1] = (prev == (unsigned)CLUSTER - 1u);
                    if (sflags[1]) counters[CLUSTER] = 0u;
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
    }
    CF_MARK(9);   // CTA done
}

}  // namespace cfb
