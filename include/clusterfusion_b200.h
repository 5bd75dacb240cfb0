/*
 * clusterfusion_b200 -- C ABI of the B200-native fused Llama decoder attention half-layer.
 *
 * This is the drop-in boundary below the reference's pybind layer.  One entry point launches
 * the single fused sm_100a kernel for any of the reference's three operator signatures:
 *
 *   variant CF_VARIANT_CHAT    replaces  llama_decoder_layer_sm90
 *       /root/reference/include/H100/llama/llama_kernel_dispatch.cu:4-146   (pybind.cpp:3-12, :110)
 *   variant CF_VARIANT_SGLANG  replaces  llama_decoder_layer_sglang_sm90
 *       /root/reference/include/H100/llama/llama_kernel_sglang_dispatch.cu:4-151 (pybind.cpp:14-25, :111)
 *   variant CF_VARIANT_PAGED   replaces  llama_decoder_layer_batch_sglang_sm90
 *       /root/reference/include/H100/llama/llama_kernel_batch_sglang_dispatch.cu:6-111 (pybind.cpp:27-43, :112)
 *
 * Plain pointers and sizes only: no torch types.  All pointers are device pointers on the
 * current CUDA device; fp16 tensors are IEEE binary16, contiguous.  Calls are asynchronous on
 * `stream` (a cudaStream_t), issue exactly one kernel launch, never synchronise and never
 * allocate, so they may be captured into a CUDA graph.
 *
 * Error convention: 0 = success; > 0 = a cudaError_t / CUresult from the runtime;
 * < 0 = CF_ERR_* (bad argument).  Never throws.  cf_last_error_string() describes the last
 * failure on the calling thread.
 */
#ifndef CLUSTERFUSION_B200_H
#define CLUSTERFUSION_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CF_ABI_VERSION 3

enum {
    CF_VARIANT_CHAT = 0,   /* W^T weights, GPT-J interleaved RoPE, eps fixed by caller (1e-6), no residual   */
    CF_VARIANT_SGLANG = 1, /* [out,in] weights, NeoX RoPE, fused residual add, contiguous KV                  */
    CF_VARIANT_PAGED = 2   /* SGLANG + batch + paged KV (page size 1) + in-pool KV append                     */
};

enum {
    CF_ERR_NULL_ARG = -1,
    CF_ERR_BAD_VARIANT = -2,
    CF_ERR_BAD_SHAPE = -3,     /* unsupported hidden / heads / head_dim combination */
    CF_ERR_BAD_ALIGNMENT = -4, /* a tensor is not 16-byte aligned                   */
    CF_ERR_NO_DEVICE = -5,     /* current device is not sm_100 (B200)               */
    CF_ERR_DRIVER = -6         /* cuTensorMapEncodeTiled entry point unavailable    */
};

/* flags */
#define CF_FLAG_OUT_FP32_PARTIAL 0x1u /* head-parallel shard: write the fp32 O-projection partial to `out`
                                         (float[batch, hidden]) instead of an fp16 result; the caller
                                         all-reduces it across ranks (Llama-2-70B config).             */

#define CF_FLAG_PDL 0x2u /* programmatic dependent launch: the kernel may START (stream its weight / KV tiles)
                            while the previous kernel in `stream` is still running; it touches x, residual,
                            outputs and workspace only after that kernel completed.  Contract: w_qkv, w_o, the KV
                            cache / pools and indptr / indices of THIS call are not being written by kernels still
                            in flight in the stream (true for a decoder stack: layer l+1's weights and cache are
                            not produced by layer l).                                                           */

/* 0x4u was CF_FLAG_GQA_CLUSTER (first-generation grouped-query cluster kernel, removed in ABI 3) */

#define CF_FLAG_LL_OUT 0x8u /* batch-1 MHA launches: reduce the O projection across clusters through flag-in-data words summed
                               in head order (bitwise reproducible output, no atomics) instead of fp32 red.global.add + a
                               last-arriver finalize.  Costs ~2.5 us per layer in a PDL chain: every CTA then lives until the
                               slowest cluster has published, so the next layer's CTAs start later (measured). */

#define CF_FLAG_PER_REQUEST 0x10u /* PAGED, batch >= 2: launch one cluster (MHA) / one set of groups (grouped-query shapes) per
                                     request as the reference does (weights re-streamed per request) instead of the batched
                                     kernels that stream every weight tile once per chunk of 4 / 8 requests (measurement / A-B) */

/* 0x20u was CF_FLAG_BATCH4 (chunks of 4 requests at batch >= 5, removed in ABI 3) */

typedef struct CfLlamaArgs {
    int32_t variant;    /* CF_VARIANT_*                                                             */
    uint32_t flags;     /* CF_FLAG_*                                                                */
    int32_t hidden;     /* model width (input of Wqkv, output of Wo); multiple of 1024, <= 8192     */
    int32_t n_q_heads;  /* query heads held by THIS call (a shard passes its local count)           */
    int32_t n_kv_heads; /* KV heads held by this call; n_q_heads % n_kv_heads == 0                  */
    int32_t head_dim;   /* must be 128                                                              */
    int32_t batch;      /* 1 for CHAT / SGLANG; >= 1 for PAGED                                      */
    uint32_t kv_len;    /* CHAT / SGLANG: rows of k_cache / v_cache (may be 0)                      */
    int32_t layer_id;   /* PAGED: index into k_pool_ptrs / v_pool_ptrs                              */
    float eps;          /* RMSNorm epsilon (the reference's 8-arg form hard-codes 1e-6)             */

    const void* x;           /* fp16 [batch, hidden]                                                */
    const void* residual_in; /* fp16 [batch, hidden]; NULL for CHAT                                 */
    const void* w_qkv;       /* CHAT: fp16 [3*hidden, hidden] = [Wq^T; Wk^T; Wv^T]
                                else: fp16 [(n_q+2*n_kv)*128, hidden]  (nn.Linear layout)          */
    const void* w_o;         /* CHAT: fp16 [n_q*128, hidden] = Wo^T;  else fp16 [hidden, n_q*128]   */
    const void* rms_w;       /* fp16 [hidden]                                                       */

    void* out;          /* fp16 [batch, hidden]  (float with CF_FLAG_OUT_FP32_PARTIAL)              */
    void* residual_out; /* fp16 [batch, hidden]; may alias residual_in (in-place, race-free); NULL for CHAT */
    void* k_new;        /* fp16 [n_kv*128]  post-RoPE K of the new token (CHAT / SGLANG)            */
    void* v_new;        /* fp16 [n_kv*128]                                                          */

    const void* k_cache; /* CHAT / SGLANG: fp16 [kv_len, n_kv*128].  The TMA descriptors are keyed on this base pointer, not on
                            kv_len: a cache that grows in place (chat/llama/model.py:355-372) never re-encodes them.
                            PAGED (optional; may be NULL): the HOST's copy of k_pool_ptrs[layer_id] / v_pool_ptrs[layer_id].
                            With it, full 16-row KV tiles are fetched by tiled TMA (16 consecutive slots) or tile::gather4
                            requests (any slots) instead of 32 row-sized bulk copies.  The kernel compares the copy with the
                            device table and ignores it if they differ (stale copy = slower, never wrong).  Pools of up to
                            2^24 slots.                                                                                    */
    const void* v_cache;

    const int32_t* indptr;       /* PAGED: int32 [batch+1]                                          */
    const int32_t* indices;      /* PAGED: int32 [nnz]; last entry of a request = slot of the new token */
    const uint64_t* k_pool_ptrs; /* PAGED: device array of per-layer pool base pointers             */
    const uint64_t* v_pool_ptrs; /*        each pool fp16 [num_slots, n_kv*128]                     */
    const int64_t* positions;    /* PAGED: int64 [batch]                                            */

    const float* cos; /* CHAT: fp32 [128] pair-repeated; SGLANG: fp32 [>=64]; PAGED: cos_sin fp32 [max_pos,128] */
    const float* sin; /* CHAT / SGLANG as cos; PAGED: unused                                        */

    void* workspace;  /* cf_llama_workspace_bytes(hidden, workspace_batch) bytes, zero-filled ONCE by the caller and
                         opaque afterwards (a launch epoch, zeroed scratch / counters and stale exchange words
                         live in it).  One workspace per stream that may run concurrently; its internal layout
                         depends on (hidden, workspace_batch), so keep both fixed for the life of a workspace.
                         Its first 16 bytes are u32 {launch epoch, -, error, peer-stage launches}: `error` becomes 1 if an
                         exchange poll inside a kernel ever timed out (~1 s); results of that launch are then invalid. */
    int32_t workspace_batch; /* the batch the workspace was sized for; 0 means `batch`.  Lets one workspace sized for
                                the largest batch serve smaller launches (batch <= workspace_batch).            */

    /* Head-parallel shards with the all-reduce FUSED into the kernel (Llama-2-70B config; grouped-query shapes, batch 1).
     * tp_world in [2, 8] enables it: every rank launches the same call on its shard; the kernel pushes its fp32
     * O-projection partial into every rank's exchange buffer over NVLink peer memory and sums all ranks' partials in rank
     * order, so `out` (fp16, or fp32 with CF_FLAG_OUT_FP32_PARTIAL) already holds the all-reduced, rank-identical result:
     * no NCCL call, no extra kernel.  tp_peer[r] = device pointer, valid on THIS device, to rank r's exchange buffer
     * (cf_tp_exchange_bytes(hidden, tp_world) bytes, zero-filled once; tp_peer[tp_rank] is the local one; peers' come
     * from cf_ipc_open).  All ranks must issue the same sequence of tp launches on their workspace.  0 / 1: disabled.  */
    int32_t tp_rank;
    int32_t tp_world;
    void* tp_peer[8];
} CfLlamaArgs;

/* Bytes of zero-initialised device workspace needed for a call with this hidden / batch (about 1 MB per request + 1-4 MB, and
 * for batch >= 2 another 18-22 MB per chunk of 8 requests: the exchange words of the grouped-query batched kernel). */
size_t cf_llama_workspace_bytes(int32_t hidden, int32_t batch);

/* Validate, encode (cached) TMA descriptors, launch the fused kernel on `stream`. */
int cf_llama_decoder_layer_launch(const CfLlamaArgs* args, void* stream);

/* ---- peer memory for the fused all-reduce: cudaMalloc'ed, zero-filled buffers that other processes on the node can map.
 * cf_ipc_alloc: allocate + zero `bytes` on the current device, return the device pointer and a 64-byte handle to send to the
 * peers (any transport: torch.distributed all_gather_object, MPI, a file).  cf_ipc_open: map a peer's handle on the current
 * device (peer access over NVLink).  cf_ipc_close / cf_ipc_free undo them.  All return 0 or a cudaError_t.                */
size_t cf_tp_exchange_bytes(int32_t hidden, int32_t tp_world);
int cf_ipc_alloc(size_t bytes, void** dev_ptr, unsigned char handle[64]);
int cf_ipc_open(const unsigned char handle[64], void** dev_ptr);
int cf_ipc_close(void* dev_ptr);
int cf_ipc_free(void* dev_ptr);

/* ---- workspace status.  The cross-CTA / cross-GPU exchanges inside the kernels poll with a bound (about one second); a
 * publisher that never arrives (peer rank stalled or crashed, CTAs of a group not co-resident because another kernel holds
 * SMs, one workspace shared by two streams) sets the workspace's sticky error word instead of hanging the GPU, and the
 * results of that launch are invalid (the peer stage also writes NaN).  cf_workspace_status: copy the word to the host
 * (synchronises `stream`; not capturable): *status = 0 ok, non-zero = some launch on this workspace since the last clear
 * timed out.  cf_workspace_clear_status: reset it (asynchronous on `stream`).  Both return 0 or a cudaError_t.          */
int cf_workspace_status(const void* workspace, void* stream, uint32_t* status);
int cf_workspace_clear_status(void* workspace, void* stream);

/* Number of cuTensorMapEncodeTiled calls this process has made through the library (tests: a decode loop in steady state
 * must not encode anything). */
uint64_t cf_debug_tensor_map_encodes(void);

/* Algorithmic HBM bytes of one call (SURVEY.md section 8d formula) -- used by bench / tests. */
uint64_t cf_llama_algorithmic_bytes(const CfLlamaArgs* args, uint64_t total_kv_rows);

/* ---- fused FFN half-layer (no counterpart kernel in the reference, whose FFN stays
 * eager PyTorch, chat/llama/model.py:407-448, :519) --------------------------------------------------------------
 * h = x + residual; residual_out = fp16(h); out = W2 (silu(W1 n) * (W3 n)), n = rmsnorm(h) * rms_w.              */
typedef struct CfFfnArgs {
    uint32_t flags;          /* CF_FLAG_OUT_FP32_PARTIAL, CF_FLAG_PDL                                       */
    int32_t hidden;          /* multiple of 256, <= 8192                                                    */
    int32_t ffn;             /* intermediate size, multiple of 16 (11008, 14336, 28672 / N ...)             */
    float eps;
    const void* x;           /* fp16 [hidden]                                                               */
    const void* residual_in; /* fp16 [hidden]                                                               */
    const void* w_gate_up;   /* fp16 [2*ffn, hidden] = [W1; W3], nn.Linear layout                           */
    const void* w_down_t;    /* fp16 [ffn, hidden] = W2^T (transpose W2 once at load time)                  */
    const void* rms_w;       /* fp16 [hidden]                                                               */
    void* out;               /* fp16 [hidden] (float with CF_FLAG_OUT_FP32_PARTIAL)                         */
    void* residual_out;      /* fp16 [hidden]; may alias residual_in                                        */
    void* workspace;         /* cf_llama_workspace_bytes(hidden, workspace_batch) bytes, zeroed once; may be shared
                                with the attention op on the same stream                                   */
    int32_t workspace_batch; /* the batch the workspace was sized for (its layout depends on it); 0 means 1  */
} CfFfnArgs;

int cf_llama_ffn_launch(const CfFfnArgs* args, void* stream);

/* ---- standalone cluster RMSNorm -- replaces the reference op `rmsnorm(input, weight)`
 * (/root/reference/include/H100/norm/norm_kernel_dispatch.cu:4-26, kernel.cuh:8-76, pybind.cpp:61-64, :114):
 * out[b] = fp16(x[b] * rsqrt(mean(x[b]^2) + eps) * weight), fp32 math.  x, out fp16 [batch, hidden]; weight fp16 [hidden];
 * hidden a multiple of 16, <= 16384 (the reference binary is fixed at 64 x 8192, eps 1e-6).  flags: CF_FLAG_PDL.      */
int cf_rmsnorm_launch(const void* x, const void* weight, void* out, int32_t batch, int32_t hidden, float eps,
                      uint32_t flags, void* stream);

/* ---- fused DeepSeek-MLA decoder attention half-layer -- replaces the reference op
 * `deepseek_decoder_layer(input, weight_q_nope, weight_q_pe, weight_uk, weight_kv_nope, weight_k_pe, weight_uv, weight_o,
 * ckv_cache, rms_input_weight, rms_ckv_weight, cos, sin) -> [1, hidden]`
 * (/root/reference/include/H100/deepseek/deepseek_kernel_dispatch.cu:4-242, kernel.cuh:9-697, pybind.cpp:45-59, :113):
 * RMSNorm -> q_nope / q_pe / ckv / k_pe projections -> RoPE on q_pe, k_pe -> RMSNorm(ckv) -> q_nope . W_uk -> decode
 * attention of all heads over the latent cache rows [0, seq_len-1) ++ the current token -> W_uv -> W_o.  No residual.
 * Weight layouts are the reference's ([in, out] row-major).  Shapes are the reference's compile-time ones (config.h:1-8:
 * hidden 2048, 16 heads, nope 128, rope 64, kv_lora_rank 512) except seq_len, which is a run-time argument here
 * (the reference binary is fixed at 4096).  As in the reference kernel the scores use the 512 latent columns only;
 * CF_DS_FLAG_ROPE_SCORES adds the decoupled-RoPE term q_pe . k_pe (cache columns 512..575).  The cache is not written:
 * the current token's row is returned through ckv_new / k_pe_new when those are non-NULL. */
#define CF_DS_FLAG_ROPE_SCORES 0x100u
typedef struct CfDeepseekArgs {
    uint32_t flags;          /* CF_FLAG_PDL, CF_DS_FLAG_ROPE_SCORES                                          */
    int32_t hidden;          /* 2048                                                                         */
    int32_t n_heads;         /* 16                                                                           */
    int32_t seq_len;         /* rows of ckv_cache, >= 1; row seq_len-1 is replaced by the current token      */
    float eps;               /* 1e-6 in the reference (kernel.cuh:46)                                        */
    const void* x;           /* fp16 [hidden]                                                                */
    const void* w_q_nope;    /* fp16 [hidden, n_heads*128]                                                   */
    const void* w_q_pe;      /* fp16 [hidden, n_heads*64]                                                    */
    const void* w_uk;        /* fp16 [128, n_heads*512]                                                      */
    const void* w_kv_nope;   /* fp16 [hidden, 512]                                                           */
    const void* w_k_pe;      /* fp16 [hidden, 64]                                                            */
    const void* w_uv;        /* fp16 [512, n_heads*128]                                                      */
    const void* w_o;         /* fp16 [n_heads*128, hidden]                                                   */
    const void* ckv_cache;   /* fp16 [seq_len, 576]                                                          */
    const void* rms_input_w; /* fp16 [hidden]                                                                */
    const void* rms_ckv_w;   /* fp16 [512]                                                                   */
    const float* cos;        /* fp32 [64]                                                                    */
    const float* sin;        /* fp32 [64]                                                                    */
    void* out;               /* fp16 [hidden]                                                                */
    void* ckv_new;           /* optional fp16 [512]: normalised latent of the current token                  */
    void* k_pe_new;          /* optional fp16 [64]: rotated k_pe of the current token                        */
    void* workspace;         /* cf_deepseek_workspace_bytes() bytes, zeroed once; one call at a time         */
} CfDeepseekArgs;
size_t cf_deepseek_workspace_bytes(void);
int cf_deepseek_decoder_layer_launch(const CfDeepseekArgs* args, void* stream);
size_t cf_sizeof_deepseek_args(void);

/* Unit-test hook for the device primitive in include/dsm.cuh:
 * launches n_clusters clusters of `cluster_size` CTAs; CTA r of cluster c contributes
 * in[(c*cluster_size + r)*n .. +n) (float); stage 0 = LINEAR (sum), 1 = ATTN (softmax-state merge of
 * [m, l, pad, pad, o[n-4]]), 4 = QUK_DEEPSEEK (all-gather).  Every CTA writes its result to
 * out[(c*cluster_size + r) * n_out ..], n_out = n (sum / merge) or n*cluster_size (gather).      */
int cf_test_cluster_reduce(const float* in, float* out, int32_t n, int32_t cluster_size,
                           int32_t n_clusters, int32_t stage, int32_t repeats, void* stream);

const char* cf_last_error_string(void);
int cf_abi_version(void);
/* sizeof(CfLlamaArgs) / sizeof(CfFfnArgs) as compiled into the library: lets a foreign-language binding (ctypes, cgo, JNI)
 * verify its mirror of the structs before the first launch. */
size_t cf_sizeof_llama_args(void);
size_t cf_sizeof_ffn_args(void);

#ifdef __cplusplus
}
#endif
#endif /* CLUSTERFUSION_B200_H */
