/*
 * C-ABI launcher for the fused sm_100a Llama decoder attention half-layer (include/clusterfusion_b200.h).
 * Torch-free translation unit: nothing here includes torch headers, so it compiles in seconds and the
 * shared library can be bound from anything that speaks C (ctypes, cgo, JNI, the pybind shim).
 *
 * Host-side work per call (the reference does 2 cudaFuncSetAttribute + 3 memset kernels + 4 tensor-map
 * encodes + 2 cudaDeviceSynchronize per call, llama_kernel_dispatch.cu:15-21, :48-121, :126-144):
 *   - argument validation,
 *   - tensor-map lookup in a small cache keyed on (pointer, dims, box) -- weights hit every time,
 *     K/V maps are re-encoded only when kv_len or the cache base pointer changes,
 *   - one cudaLaunchKernelEx on the caller's stream.  No sync, no allocation, graph-capturable.
 */
#include "../../include/clusterfusion_b200.h"
#include "llama_decoder_kernel.cuh"
#include "llama_decoder_gqa2_kernel.cuh"
#include "llama_decoder_batch_kernel.cuh"
#include "llama_decoder_batch8_kernel.cuh"
#include "llama_decoder_gqa_batch_kernel.cuh"
#include "llama_ffn_kernel.cuh"
#include "rmsnorm_kernel.cuh"
#include "deepseek_mla_kernel.cuh"

#include <cuda.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>

namespace {

thread_local std::string g_last_error;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                   CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                   CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return (EncodeTiledFn) nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

struct MapKey {
    const void* ptr;
    uint64_t rows, cols;
    uint32_t box_c, box_r;      // bit 31 of box_r: 128-byte swizzle
    bool operator==(const MapKey& o) const {
        return ptr == o.ptr && rows == o.rows && cols == o.cols && box_c == o.box_c && box_r == o.box_r;
    }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const {
        uint64_t h = reinterpret_cast<uint64_t>(k.ptr) * 0x9E3779B97F4A7C15ull;
        h ^= (k.rows + 0x632BE59BD9B4E019ull) + (h << 6) + (h >> 2);
        h ^= (k.cols * 31 + k.box_c * 7 + k.box_r) + (h << 6) + (h >> 2);
        return static_cast<size_t>(h);
    }
};

std::mutex g_map_mutex;
// Every map is keyed on (base pointer, extent, box): weights, and KV caches / pools mapped from their base pointer with a
// fixed (huge) row extent -- nothing in the key moves while a model decodes, so the steady state of a decode loop is zero
// encodes per token.  Never flushed in normal operation: a serving process holds a few hundred entries.
std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_map_cache;
std::atomic<uint64_t> g_encode_calls{0};

// 2-D fp16 row-major tensor [rows][cols], box {box_c, box_r}; OOB rows/cols read as zero.
int get_tensor_map(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_c,
                   uint32_t box_r, bool swizzle128 = false) {
    const MapKey key{ptr, rows, cols, box_c, box_r | (swizzle128 ? 0x80000000u : 0u)};
    {
        std::lock_guard<std::mutex> lk(g_map_mutex);
        auto it = g_map_cache.find(key);
        if (it != g_map_cache.end()) { *out = it->second; return 0; }
    }
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return fail(CF_ERR_DRIVER, "cuTensorMapEncodeTiled entry point not available");
    const cuuint64_t dims[2] = {cols, rows ? rows : 1};
    const cuuint64_t strides[1] = {cols * sizeof(__half)};
    const cuuint32_t box[2] = {box_c, box_r};
    const cuuint32_t estr[2] = {1, 1};
    CUtensorMap m;
    g_encode_calls.fetch_add(1, std::memory_order_relaxed);
    const CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
#ifdef CF_EXPERIMENT_L2_PROMO_128      /* tools/sweep_build.sh experiment, never defined in the product build */
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
#else
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
#endif
    if (r != CUDA_SUCCESS) return fail((int)r, "cuTensorMapEncodeTiled failed (CUresult %d) rows=%llu cols=%llu",
                                       (int)r, (unsigned long long)rows, (unsigned long long)cols);
    {
        std::lock_guard<std::mutex> lk(g_map_mutex);
        // bounded as a safety net only (a process that keeps allocating new tensors): maps are re-encoded on demand
        if (g_map_cache.size() > (1u << 16)) g_map_cache.clear();
        g_map_cache.emplace(key, m);
    }
    *out = m;
    return 0;
}

// Row extent of the base-pointer-keyed K/V maps: a power of two >= 2^24 that covers kv_len.  The extent says nothing about
// the allocation -- the kernel only ever requests rows < kv_len through these maps (full tiles; a ragged last tile is
// fetched row by row), so nothing beyond the caller's tensor is touched.
uint64_t kv_map_rows(uint64_t kv_len) {
    uint64_t r = 1ull << 24;
    while (r < kv_len) r <<= 1;
    return r;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <int CLUSTER, typename Kern, typename Params>
int launch_kernel(Kern kern, int smem_bytes, int slot, const Params& kp, int n_clusters, int batch, bool pdl,
                  cudaStream_t stream) {
    static std::once_flag once[12][16];
    static cudaError_t attr_err[12][16];
    int dev = 0;
    cudaGetDevice(&dev);
    std::call_once(once[slot][dev & 15], [&] {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
        if (e == cudaSuccess && CLUSTER > 8)
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        attr_err[slot][dev & 15] = e;
    });
    const cudaError_t ae = attr_err[slot][dev & 15];
    if (ae != cudaSuccess) return fail((int)ae, "cudaFuncSetAttribute: %s", cudaGetErrorString(ae));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(n_clusters * CLUSTER, batch, 1);
    cfg.blockDim = dim3(cfb::BLOCK_THREADS, 1, 1);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CLUSTER;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 2 : 1;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, kp);
    if (e != cudaSuccess) return fail((int)e, "kernel launch failed: %s", cudaGetErrorString(e));
    return 0;
}

template <int VARIANT, int CLUSTER>
int launch(const cfb::KParams& kp, int n_clusters, int batch, bool pdl, cudaStream_t stream) {
    return launch_kernel<CLUSTER>(cfb::llama_decoder_layer_kernel<VARIANT, CLUSTER>, cfb::Smem<CLUSTER>::TOTAL, VARIANT,
                                  kp, n_clusters, batch, pdl, stream);
}

// second-generation grouped-query kernel: G plain CTAs per (KV head, 4 query heads) group, exchanges through L2
template <int VARIANT>
int launch_gqa2(const cfb::G2Params& gp, int batch, bool pdl, cudaStream_t stream) {
    return launch_kernel<1>(cfb::llama_decoder_layer_gqa2_kernel<VARIANT, 4>, cfb::SmemGqa2<4>::TOTAL, 8 + VARIANT,
                            gp, gp.n_groups * gp.G, batch, pdl, stream);
}

int sm_count_of_current_device() {
    static int sm_count[16] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (sm_count[dev & 15] == 0) cudaDeviceGetAttribute(&sm_count[dev & 15], cudaDevAttrMultiProcessorCount, dev);
    return sm_count[dev & 15];
}

// workspace carve-up (bytes): [header 256][scratch fp32 batch*hidden][counters u32 batch*32]
//                              [qkv_ll u64][attn_ll u64][ag_ll u64][gcounters u32 batch*128][out_ll u64 (hidden/128)*hidden]
// The header sits at a shape-independent offset: its epoch word tags every flag-in-data word of the group kernel, so
// stale words left by launches of any other shape on the same workspace can never match.
size_t ws_off_scratch() { return cfb::WS_HEADER_BYTES; }
size_t ws_off_counters(int hidden, int batch) { return ws_off_scratch() + (size_t)batch * hidden * sizeof(float); }
size_t ws_off_qkv_ll(int hidden, int batch) { return ws_off_counters(hidden, batch) + (size_t)batch * 32 * sizeof(uint32_t); }
size_t ws_off_attn_ll(int hidden, int batch) {
    return ws_off_qkv_ll(hidden, batch) + (size_t)batch * cfb::G2_GROUPS_MAX * 2 * cfb::SmemGqa2<4>::R * sizeof(uint64_t);
}
size_t ws_off_ag_ll(int hidden, int batch) {
    return ws_off_attn_ll(hidden, batch) + (size_t)batch * cfb::G2_SLOTS * 4 * cfb::SmemGqa2<4>::PAY * sizeof(uint64_t);
}
size_t ws_off_gcounters(int hidden, int batch) {
    return ws_off_ag_ll(hidden, batch) + (size_t)batch * cfb::G2_GROUPS_MAX * 4 * cfb::HEAD_DIM * sizeof(uint64_t);
}
size_t ws_off_out_ll_b1(int hidden, int batch) {
    return ws_off_gcounters(hidden, batch) + (size_t)batch * cfb::G2_COUNTERS * sizeof(uint32_t);
}
// grouped-query batched kernel (llama_decoder_gqa_batch_kernel.cuh): per chunk of 8 requests, all (value, epoch) words
size_t ws_off_gb(int hidden, int batch) {
    // batch == 1 launches reduce the O projection across clusters through (value, epoch) words [hidden/128][hidden]
    return ws_off_out_ll_b1(hidden, batch) + (size_t)(hidden / 128) * hidden * sizeof(uint64_t);
}
int gb_chunks(int batch) { return batch >= 2 ? (batch + cfb::GB_BC - 1) / cfb::GB_BC : 0; }
size_t ws_off_gb_qkvp(int hidden, int batch) { return ws_off_gb(hidden, batch) + (size_t)gb_chunks(batch) * cfb::GB_CTAS_MAX * cfb::GB_BC * sizeof(uint64_t); }
size_t ws_off_gb_qkvf(int hidden, int batch) {
    return ws_off_gb_qkvp(hidden, batch) + (size_t)gb_chunks(batch) * cfb::GB_CTAS_MAX * cfb::SmemGqaB::R * cfb::GB_BC * sizeof(uint64_t);
}
size_t ws_off_gb_attn(int hidden, int batch) {
    return ws_off_gb_qkvf(hidden, batch) + (size_t)gb_chunks(batch) * cfb::G2_GROUPS_MAX * cfb::GB_BC * cfb::GB_QKVF_WORDS * sizeof(uint64_t);
}
size_t ws_off_gb_ag(int hidden, int batch) {
    return ws_off_gb_attn(hidden, batch) + (size_t)gb_chunks(batch) * cfb::GB_CTAS_MAX * cfb::GB_BC * cfb::GB_STATE_WORDS * sizeof(uint64_t);
}
size_t ws_off_gb_out(int hidden, int batch) {
    return ws_off_gb_ag(hidden, batch) + (size_t)gb_chunks(batch) * cfb::G2_GROUPS_MAX * cfb::GB_BC * cfb::GB_AG_WORDS * sizeof(uint64_t);
}
size_t ws_total(int hidden, int batch) {
    return ws_off_gb_out(hidden, batch) + (size_t)gb_chunks(batch) * cfb::G2_GROUPS_MAX * hidden * cfb::GB_BC * sizeof(uint64_t);
}

bool device_is_sm100() {
    static int cached[16] = {0};   // 0 unknown, 1 yes, 2 no
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return false;
    int& c = cached[dev & 15];
    if (c == 0) {
        int major = 0;
        cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
        c = (major == 10) ? 1 : 2;
    }
    return c == 1;
}

}  // namespace

extern "C" {

int cf_abi_version(void) { return CF_ABI_VERSION; }

const char* cf_last_error_string(void) { return g_last_error.c_str(); }

uint64_t cf_debug_tensor_map_encodes(void) { return g_encode_calls.load(std::memory_order_relaxed); }

int cf_workspace_status(const void* workspace, void* stream_, uint32_t* status) {
    if (!workspace || !status) return fail(CF_ERR_NULL_ARG, "cf_workspace_status: NULL argument");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    uint32_t word = 0;
    cudaError_t e = cudaMemcpyAsync(&word, static_cast<const uint32_t*>(workspace) + 2, sizeof word, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) return fail((int)e, "cf_workspace_status: %s", cudaGetErrorString(e));
    *status = word;
    return 0;
}

int cf_workspace_clear_status(void* workspace, void* stream_) {
    if (!workspace) return fail(CF_ERR_NULL_ARG, "cf_workspace_clear_status: NULL argument");
    const cudaError_t e = cudaMemsetAsync(static_cast<uint32_t*>(workspace) + 2, 0, sizeof(uint32_t), static_cast<cudaStream_t>(stream_));
    return e == cudaSuccess ? 0 : fail((int)e, "cf_workspace_clear_status: %s", cudaGetErrorString(e));
}

size_t cf_sizeof_llama_args(void) { return sizeof(CfLlamaArgs); }
size_t cf_sizeof_ffn_args(void) { return sizeof(CfFfnArgs); }

size_t cf_llama_workspace_bytes(int32_t hidden, int32_t batch) {
    if (hidden <= 0 || batch <= 0) return 0;
    // header + fp32 scratch [batch][hidden] + counters [batch][32] (any cluster size) + the grouped-query kernel's L2
    // exchange buffers (flag-in-data words for q|k|v, softmax states, merged attention output): ~0.95 MB per request,
    // + the batch-1 cross-cluster output words (1 MB at hidden 4096, 4 MB at 8192)
    return ws_total(hidden, batch);
}

uint64_t cf_llama_algorithmic_bytes(const CfLlamaArgs* a, uint64_t total_kv_rows) {
    if (!a) return 0;
    const uint64_t D = 128, H = a->hidden, hq = a->n_q_heads, hkv = a->n_kv_heads, bs = a->batch;
    uint64_t b = 2 * (hq + 2 * hkv) * D * H + 2 * hq * D * H;          // Wqkv + Wo
    b += 2 * 2 * total_kv_rows * hkv * D;                               // K and V, once per KV head
    uint64_t small = 2 * H /*x*/ + 2 * H /*out*/ + 2 * 2 * hkv * D /*k_new, v_new*/;
    if (a->variant != CF_VARIANT_CHAT) small += 4 * H;                  // residual in / out
    b += bs * small + 2 * H /*rms_w*/ + (a->variant == CF_VARIANT_CHAT ? 2 * 4 * D : bs * 4 * D) /*cos,sin*/;
    return b;
}

int cf_llama_decoder_layer_launch(const CfLlamaArgs* a, void* stream_) {
    if (!a) return fail(CF_ERR_NULL_ARG, "args is NULL");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (a->variant < CF_VARIANT_CHAT || a->variant > CF_VARIANT_PAGED)
        return fail(CF_ERR_BAD_VARIANT, "unknown variant %d", a->variant);
    const bool chat = a->variant == CF_VARIANT_CHAT, paged = a->variant == CF_VARIANT_PAGED;
    constexpr int CL = 4;
    if (a->head_dim != 128) return fail(CF_ERR_BAD_SHAPE, "head_dim must be 128 (got %d)", a->head_dim);
    if (a->n_q_heads <= 0 || a->n_kv_heads <= 0 || a->n_q_heads % a->n_kv_heads != 0)
        return fail(CF_ERR_BAD_SHAPE, "bad head counts q=%d kv=%d", a->n_q_heads, a->n_kv_heads);
    const bool gqa = a->n_q_heads != a->n_kv_heads;
    if (!gqa) {
        const int ks_max = chat ? 1024 : cfb::KS_MAX;      // chat keeps 4 write-once O-projection slots per column
        if (a->hidden <= 0 || a->hidden % (CL * 256) != 0 || a->hidden / CL > ks_max)
            return fail(CF_ERR_BAD_SHAPE, "hidden must be a multiple of %d and <= %d (got %d)", CL * 256,
                        CL * ks_max, a->hidden);
    } else {
        // grouped-query path: 4 query heads per group, nn.Linear weight layout
        if (chat) return fail(CF_ERR_BAD_SHAPE, "grouped-query attention needs the nn.Linear layout (SGLANG / PAGED)");
        if ((a->n_q_heads / a->n_kv_heads) % 4 != 0)
            return fail(CF_ERR_BAD_SHAPE, "grouped-query attention needs a multiple of 4 query heads per KV head (q=%d, kv=%d)",
                        a->n_q_heads, a->n_kv_heads);
        if (a->hidden <= 0 || a->hidden % (16 * 256) != 0 || a->hidden > cfb::G2_HIDDEN_MAX)
            return fail(CF_ERR_BAD_SHAPE, "GQA: hidden must be a multiple of 4096 and <= %d (got %d)", cfb::G2_HIDDEN_MAX, a->hidden);
    }
    if (a->batch < 1 || (!paged && a->batch != 1))
        return fail(CF_ERR_BAD_SHAPE, "batch must be 1 for CHAT/SGLANG and >= 1 for PAGED (got %d)", a->batch);
    if (a->batch > 65535) return fail(CF_ERR_BAD_SHAPE, "batch too large (%d)", a->batch);
    if (!a->x || !a->w_qkv || !a->w_o || !a->rms_w || !a->out || !a->cos || !a->workspace)
        return fail(CF_ERR_NULL_ARG, "x / w_qkv / w_o / rms_w / out / cos / workspace must be non-NULL");
    if (!paged && (!a->sin || !a->k_new || !a->v_new)) return fail(CF_ERR_NULL_ARG, "sin / k_new / v_new must be non-NULL");
    if (!paged && a->kv_len > 0 && (!a->k_cache || !a->v_cache))
        return fail(CF_ERR_NULL_ARG, "k_cache / v_cache must be non-NULL when kv_len > 0");
    if (!chat && (!a->residual_in || !a->residual_out))
        return fail(CF_ERR_NULL_ARG, "residual_in / residual_out must be non-NULL for SGLANG / PAGED");
    if (paged && (!a->indptr || !a->indices || !a->k_pool_ptrs || !a->v_pool_ptrs || !a->positions))
        return fail(CF_ERR_NULL_ARG, "paged arguments must be non-NULL");
    if (a->kv_len > 0x7fffffffu / 2) return fail(CF_ERR_BAD_SHAPE, "kv_len too large");
    const void* al[] = {a->x, a->residual_in, a->w_qkv, a->w_o, a->rms_w, a->out, a->residual_out,
                        a->k_new, a->v_new, a->k_cache, a->v_cache, a->workspace, a->cos, a->sin};
    for (const void* q : al)
        if (q && !aligned16(q)) return fail(CF_ERR_BAD_ALIGNMENT, "all tensors must be 16-byte aligned (%p)", q);
    if (!device_is_sm100()) return fail(CF_ERR_NO_DEVICE, "current CUDA device is not compute capability 10.x");

    cfb::KParams kp;
    memset(&kp, 0, sizeof kp);
    const uint64_t H = a->hidden, qd = (uint64_t)a->n_q_heads * 128, kvd = (uint64_t)a->n_kv_heads * 128;
    int rc;
    // batched paged decode: one cluster per head serves chunks of 4 requests, weights streamed once per chunk
    const bool batched = paged && !gqa && a->batch >= 2 && a->hidden / CL <= cfb::BK_KS_MAX &&
                         a->residual_out != a->residual_in && !(a->flags & CF_FLAG_PER_REQUEST);
    // grouped-query shapes, batch >= 2: weights streamed once per chunk of 8 requests
    // (two requests whose groups are co-resident with 8 CTAs each are faster through the group kernel: 27.9 vs 31.2 us for
    //  Llama-3-8B shapes, profiles/round2_gqa_batch_probe_8b.txt)
    const bool two_coresident = a->batch == 2 && gqa &&
                                a->n_kv_heads * ((a->n_q_heads / a->n_kv_heads) / 4) * 8 * 2 <= sm_count_of_current_device();
    const bool gqa_batched = paged && gqa && a->batch >= 2 && a->tp_world <= 1 && !(a->flags & CF_FLAG_PER_REQUEST) && !two_coresident;
    if (batched || gqa_batched) {
        // tensor-core GEMVs: [32 rows x 64 cols] boxes, 128-byte swizzled (two per 8 KB tile)
        if ((rc = get_tensor_map(&kp.tm_wqkv, a->w_qkv, qd + 2 * kvd, H, 64, 32, true))) return rc;
        if ((rc = get_tensor_map(&kp.tm_wo, a->w_o, H, qd, 64, 32, true))) return rc;
    } else if (chat) {
        // MHA kernels read weight tiles with ldmatrix: [rows x 64 cols] boxes, 128-byte swizzled (2 or 4 per 8 KB tile)
        if ((rc = get_tensor_map(&kp.tm_wqkv, a->w_qkv, 3 * H, H, 64, cfb::ROWS256, true))) return rc;
        if ((rc = get_tensor_map(&kp.tm_wo, a->w_o, qd, H, 64, cfb::ROWS256, true))) return rc;
    } else if (!gqa) {
        if ((rc = get_tensor_map(&kp.tm_wqkv, a->w_qkv, qd + 2 * kvd, H, 64, cfb::ROWS256, true))) return rc;
        if ((rc = get_tensor_map(&kp.tm_wo, a->w_o, H, qd, 64, cfb::ROWS256, true))) return rc;
    } else {
        // group kernel: CUDA-core GEMVs over unswizzled [16 x 256] tiles (512-byte row segments)
        if ((rc = get_tensor_map(&kp.tm_wqkv, a->w_qkv, qd + 2 * kvd, H, 256, cfb::ROWS512))) return rc;
        if ((rc = get_tensor_map(&kp.tm_wo, a->w_o, H, qd, 256, cfb::ROWS512))) return rc;
    }
    const bool group_kernel = gqa;
    if (!paged) {
        // kv_len == 0: no tile is ever requested; point the maps at any valid address
        const void* kc = a->kv_len ? a->k_cache : a->x;
        const void* vc = a->kv_len ? a->v_cache : a->x;
        if (group_kernel) {
            // group kernel (tensor-core attention): 64-dim half rows, 128-byte swizzled, so ldmatrix is conflict-free.  Full
            // tiles come through the {64 x 16} boxes, a ragged last tile through tile::gather4 over the {64 x 1} maps (rows
            // clamped to kv_len - 1), so nothing depends on kv_len here either
            if ((rc = get_tensor_map(&kp.tm_k, kc, kv_map_rows(a->kv_len), kvd, 64, cfb::ROWS512, true))) return rc;
            if ((rc = get_tensor_map(&kp.tm_v, vc, kv_map_rows(a->kv_len), kvd, 64, cfb::ROWS512, true))) return rc;
            if ((rc = get_tensor_map(&kp.tm_kg, kc, kv_map_rows(a->kv_len), kvd, 64, 1, true))) return rc;
            if ((rc = get_tensor_map(&kp.tm_vg, vc, kv_map_rows(a->kv_len), kvd, 64, 1, true))) return rc;
        } else {
            if ((rc = get_tensor_map(&kp.tm_k, kc, kv_map_rows(a->kv_len), kvd, 128, cfb::ROWS512))) return rc;
            if ((rc = get_tensor_map(&kp.tm_v, vc, kv_map_rows(a->kv_len), kvd, 128, cfb::ROWS512))) return rc;
        }
        kp.k_base = static_cast<const __half*>(kc);
        kp.v_base = static_cast<const __half*>(vc);
    } else if (gqa && a->k_cache && a->v_cache) {
        // paged group kernel with the pool addresses known on the host: swizzled maps over the pools, tensor-core attention
        if ((rc = get_tensor_map(&kp.tm_k, a->k_cache, (uint64_t)cfb::POOL_MAP_ROWS, kvd, 64, cfb::ROWS512, true))) return rc;
        if ((rc = get_tensor_map(&kp.tm_v, a->v_cache, (uint64_t)cfb::POOL_MAP_ROWS, kvd, 64, cfb::ROWS512, true))) return rc;
        if ((rc = get_tensor_map(&kp.tm_kg, a->k_cache, (uint64_t)cfb::POOL_MAP_ROWS, kvd, 64, 1, true))) return rc;
        if ((rc = get_tensor_map(&kp.tm_vg, a->v_cache, (uint64_t)cfb::POOL_MAP_ROWS, kvd, 64, 1, true))) return rc;
        kp.k_base = static_cast<const __half*>(a->k_cache);
        kp.v_base = static_cast<const __half*>(a->v_cache);
    } else if (!gqa && a->k_cache && a->v_cache) {
        // optional fast paths of the paged form: the caller knows the pool addresses of this layer on the host (k_cache /
        // v_cache = its copy of k_pool_ptrs[layer_id] / v_pool_ptrs[layer_id]).  Tiles whose 16 rows sit in consecutive
        // slots are fetched with one tiled TMA load per tensor, all other full tiles with tile::gather4 requests (four
        // arbitrary rows each).  The pool's slot count is not part of the interface: the maps cover 2^24 slots.
        if ((rc = get_tensor_map(&kp.tm_k, a->k_cache, (uint64_t)cfb::POOL_MAP_ROWS, kvd, 128, cfb::ROWS512))) return rc;
        if ((rc = get_tensor_map(&kp.tm_v, a->v_cache, (uint64_t)cfb::POOL_MAP_ROWS, kvd, 128, cfb::ROWS512))) return rc;
        if ((rc = get_tensor_map(&kp.tm_kg, a->k_cache, (uint64_t)cfb::POOL_MAP_ROWS, kvd, 128, 1))) return rc;
        if ((rc = get_tensor_map(&kp.tm_vg, a->v_cache, (uint64_t)cfb::POOL_MAP_ROWS, kvd, 128, 1))) return rc;
        kp.k_base = static_cast<const __half*>(a->k_cache);
        kp.v_base = static_cast<const __half*>(a->v_cache);
    }
    kp.x = static_cast<const __half*>(a->x);
    kp.residual_in = static_cast<const __half*>(a->residual_in);
    kp.rms_w = static_cast<const __half*>(a->rms_w);
    kp.out = a->out;
    kp.residual_out = static_cast<__half*>(a->residual_out);
    kp.k_new = static_cast<__half*>(a->k_new);
    kp.v_new = static_cast<__half*>(a->v_new);
    kp.cos = a->cos;
    kp.sin = a->sin;
    kp.indptr = a->indptr;
    kp.indices = a->indices;
    kp.k_pool_ptrs = reinterpret_cast<const unsigned long long*>(a->k_pool_ptrs);
    kp.v_pool_ptrs = reinterpret_cast<const unsigned long long*>(a->v_pool_ptrs);
    kp.positions = reinterpret_cast<const long long*>(a->positions);
    const int wsb = a->workspace_batch > 0 ? a->workspace_batch : a->batch;
    if (wsb < a->batch) return fail(CF_ERR_BAD_SHAPE, "workspace_batch (%d) < batch (%d)", wsb, a->batch);
    kp.scratch = reinterpret_cast<float*>(static_cast<char*>(a->workspace) + ws_off_scratch());
    kp.counters = reinterpret_cast<unsigned*>(static_cast<char*>(a->workspace) + ws_off_counters(a->hidden, wsb));
    kp.header = reinterpret_cast<unsigned*>(a->workspace);
    kp.out_ll = (a->batch == 1 && (gqa || (a->flags & CF_FLAG_LL_OUT))) ? reinterpret_cast<unsigned long long*>(static_cast<char*>(a->workspace) + ws_off_out_ll_b1(a->hidden, wsb))
                              : nullptr;
    kp.eps = a->eps;
    kp.hidden = a->hidden;
    kp.n_heads = a->n_q_heads;
    kp.n_kv_heads = a->n_kv_heads;
    kp.kv_len = (int)a->kv_len;
    kp.layer_id = a->layer_id;
    kp.flags = a->flags;
    kp.batch = a->batch;
    kp.tp_rank = 0;
    kp.tp_world = 1;
    if (a->tp_world > 1) {
        if (a->tp_world > 8 || a->tp_rank < 0 || a->tp_rank >= a->tp_world)
            return fail(CF_ERR_BAD_SHAPE, "tp_world must be in [2, 8] and 0 <= tp_rank < tp_world (got rank %d of %d)", a->tp_rank, a->tp_world);
        if (!gqa || a->batch != 1)
            return fail(CF_ERR_BAD_SHAPE, "the fused all-reduce needs a grouped-query shape, batch 1 and the group kernel");
        for (int r = 0; r < a->tp_world; ++r) {
            if (!a->tp_peer[r] || !aligned16(a->tp_peer[r])) return fail(CF_ERR_NULL_ARG, "tp_peer[%d] must be a 16-byte aligned device pointer", r);
            kp.tp_peer[r] = static_cast<unsigned long long*>(a->tp_peer[r]);
        }
        kp.tp_rank = a->tp_rank;
        kp.tp_world = a->tp_world;
    }

    const bool pdl = (a->flags & CF_FLAG_PDL) != 0;
    if (gqa_batched) {
        const int n_groups = a->n_kv_heads * ((a->n_q_heads / a->n_kv_heads) / 4);
        if (n_groups > cfb::G2_GROUPS_MAX)
            return fail(CF_ERR_BAD_SHAPE, "GQA: at most %d (KV head, 4 query heads) groups per call (got %d)", cfb::G2_GROUPS_MAX, n_groups);
        // the CTAs of a chunk spin on each other through L2: one chunk (groups x G CTAs) must fit the device; later chunks are
        // dispatched as earlier ones drain (chunks are the slow grid dimension, see llama_decoder_gqa2_kernel.cuh)
        static int capacity[16] = {0};
        int dev = 0;
        cudaGetDevice(&dev);
        if (capacity[dev & 15] == 0) {
            int occ = 0;
            cudaFuncSetAttribute(cfb::llama_decoder_layer_gqa_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, cfb::SmemGqaB::TOTAL);
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, cfb::llama_decoder_layer_gqa_batch_kernel, cfb::BLOCK_THREADS,
                                                              cfb::SmemGqaB::TOTAL) != cudaSuccess || occ < 1)
                occ = 1;
            capacity[dev & 15] = occ * sm_count_of_current_device();
        }
        const int n_sm = capacity[dev & 15] < cfb::GB_CTAS_MAX ? capacity[dev & 15] : cfb::GB_CTAS_MAX;
        int G = cfb::G2_G_MAX;
        while (G > 8 && (n_groups * G > n_sm || a->hidden / G < 128)) G >>= 1;
        if (n_groups * G > n_sm || a->hidden / G > cfb::BK_KS_MAX || a->hidden % (G * 128) != 0)
            return fail(CF_ERR_BAD_SHAPE, "GQA batched: %d groups x %d CTAs (hidden %d) do not fit this device (%d CTAs)", n_groups, G, a->hidden, n_sm);
        cfb::GBParams gp;
        memset(&gp, 0, sizeof gp);
        gp.k = kp;
        char* ws = static_cast<char*>(a->workspace);
        gp.ss_ll = reinterpret_cast<unsigned long long*>(ws + ws_off_gb(a->hidden, wsb));
        gp.qkvp_ll = reinterpret_cast<unsigned long long*>(ws + ws_off_gb_qkvp(a->hidden, wsb));
        gp.qkvf_ll = reinterpret_cast<unsigned long long*>(ws + ws_off_gb_qkvf(a->hidden, wsb));
        gp.attn_ll = reinterpret_cast<unsigned long long*>(ws + ws_off_gb_attn(a->hidden, wsb));
        gp.ag_ll = reinterpret_cast<unsigned long long*>(ws + ws_off_gb_ag(a->hidden, wsb));
        gp.out_ll = reinterpret_cast<unsigned long long*>(ws + ws_off_gb_out(a->hidden, wsb));
        gp.G = G;
        gp.n_groups = n_groups;
        return launch_kernel<1>(cfb::llama_decoder_layer_gqa_batch_kernel, cfb::SmemGqaB::TOTAL, 7, gp, n_groups * G,
                                (a->batch + cfb::GB_BC - 1) / cfb::GB_BC, pdl, stream);
    }
    if (gqa) {
        // group kernel: G CTAs per group, G = largest power of two in [8, 64] with groups * G * batch <= #SMs
        const int n_groups = a->n_kv_heads * ((a->n_q_heads / a->n_kv_heads) / 4);
        if (n_groups > cfb::G2_GROUPS_MAX)
            return fail(CF_ERR_BAD_SHAPE, "GQA: at most %d (KV head, 4 query heads) groups per call (got %d)", cfb::G2_GROUPS_MAX, n_groups);
        // The CTAs of a group spin on each other through L2, so a group must be co-resident (for batch 1: the whole grid,
        // because the output columns are summed across groups the same way).  Capacity = what the occupancy calculator says
        // this kernel can keep resident on this device (1 CTA / SM at its shared-memory footprint) -- G is the largest power
        // of two that fits.  For batch > 1 the grid exceeds it by design: requests are the slow grid dimension and a group's
        // CTAs are contiguous in blockIdx.x, so a group only ever waits for CTAs that are dispatched as earlier groups drain.
        // Anything that breaks those assumptions (another kernel holding SMs, MPS partitions) ends in the bounded poll's
        // error word, not a hang: see cf_workspace_status.
        static int capacity[16] = {0};
        int dev = 0;
        cudaGetDevice(&dev);
        if (capacity[dev & 15] == 0) {
            int occ = 0;
            cudaFuncSetAttribute(cfb::llama_decoder_layer_gqa2_kernel<cfb::SGLANG, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 cfb::SmemGqa2<4>::TOTAL);
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, cfb::llama_decoder_layer_gqa2_kernel<cfb::SGLANG, 4>,
                                                              cfb::BLOCK_THREADS, cfb::SmemGqa2<4>::TOTAL) != cudaSuccess || occ < 1)
                occ = 1;
            capacity[dev & 15] = occ * sm_count_of_current_device();
        }
        const int n_sm = capacity[dev & 15];
        int G = cfb::G2_G_MAX;
        while (G > 8 && (long long)n_groups * G * a->batch > n_sm) G >>= 1;
        if (a->batch == 1 && n_groups * G > n_sm)
            return fail(CF_ERR_BAD_SHAPE, "GQA: %d groups x %d CTAs cannot be co-resident on this device (%d CTAs fit)", n_groups, G, n_sm);
        cfb::G2Params gp;
        memset(&gp, 0, sizeof gp);
        gp.k = kp;
        char* ws = static_cast<char*>(a->workspace);
        gp.qkv_ll = reinterpret_cast<unsigned long long*>(ws + ws_off_qkv_ll(a->hidden, wsb));
        gp.attn_ll = reinterpret_cast<unsigned long long*>(ws + ws_off_attn_ll(a->hidden, wsb));
        gp.ag_ll = reinterpret_cast<unsigned long long*>(ws + ws_off_ag_ll(a->hidden, wsb));
        gp.gcounters = reinterpret_cast<unsigned*>(ws + ws_off_gcounters(a->hidden, wsb));
        gp.G = G;
        gp.n_groups = n_groups;
        return paged ? launch_gqa2<cfb::PAGED>(gp, a->batch, pdl, stream) : launch_gqa2<cfb::SGLANG>(gp, 1, pdl, stream);
    }
    if (batched && a->batch >= 5)      // chunks of 8: the whole N dimension of the MMA
        return launch_kernel<CL>(cfb::llama_decoder_layer_batch8_kernel, cfb::SmemB8::TOTAL, 3, kp, a->n_q_heads,
                                 (a->batch + 7) / 8, pdl, stream);
    if (batched)
        return launch_kernel<CL>(cfb::llama_decoder_layer_batch_kernel<4>, cfb::SmemB<4>::TOTAL, 11, kp, a->n_q_heads,
                                 (a->batch + 3) / 4, pdl, stream);
    switch (a->variant) {
        case CF_VARIANT_CHAT: return launch<cfb::CHAT, CL>(kp, a->n_q_heads, 1, pdl, stream);
        case CF_VARIANT_SGLANG: return launch<cfb::SGLANG, CL>(kp, a->n_q_heads, 1, pdl, stream);
        default: return launch<cfb::PAGED, CL>(kp, a->n_q_heads, a->batch, pdl, stream);
    }
}

}  // extern "C"

extern "C" int cf_llama_ffn_launch(const CfFfnArgs* a, void* stream_) {
    if (!a) return fail(CF_ERR_NULL_ARG, "args is NULL");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (a->hidden <= 0 || a->hidden % 256 != 0 || a->hidden > cfb::FFN_HIDDEN_MAX)
        return fail(CF_ERR_BAD_SHAPE, "ffn: hidden must be a multiple of 256 and <= %d (got %d)", cfb::FFN_HIDDEN_MAX, a->hidden);
    if (a->ffn <= 0 || a->ffn % cfb::FFN_BLOCK != 0)
        return fail(CF_ERR_BAD_SHAPE, "ffn: intermediate size must be a multiple of %d (got %d)", cfb::FFN_BLOCK, a->ffn);
    if (!a->x || !a->residual_in || !a->w_gate_up || !a->w_down_t || !a->rms_w || !a->out || !a->residual_out || !a->workspace)
        return fail(CF_ERR_NULL_ARG, "ffn: all pointers must be non-NULL");
    const void* al[] = {a->x, a->residual_in, a->w_gate_up, a->w_down_t, a->rms_w, a->out, a->residual_out, a->workspace};
    for (const void* q : al)
        if (!aligned16(q)) return fail(CF_ERR_BAD_ALIGNMENT, "all tensors must be 16-byte aligned (%p)", q);
    if (!device_is_sm100()) return fail(CF_ERR_NO_DEVICE, "current CUDA device is not compute capability 10.x");
    static int sm_count[16] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (sm_count[dev & 15] == 0) cudaDeviceGetAttribute(&sm_count[dev & 15], cudaDevAttrMultiProcessorCount, dev);
    const int n_blocks = a->ffn / cfb::FFN_BLOCK;
    int grid = sm_count[dev & 15] < n_blocks ? sm_count[dev & 15] : n_blocks;
    if ((n_blocks + grid - 1) / grid > cfb::FFN_NB_MAX)
        return fail(CF_ERR_BAD_SHAPE, "ffn: intermediate size %d too large for %d SMs", a->ffn, grid);

    cfb::FfnParams fp;
    memset(&fp, 0, sizeof fp);
    int rc;
    if ((rc = get_tensor_map(&fp.tm_w13, a->w_gate_up, 2ull * a->ffn, a->hidden, 256, cfb::FFN_BLOCK))) return rc;
    if ((rc = get_tensor_map(&fp.tm_w2t, a->w_down_t, a->ffn, a->hidden, 256, cfb::FFN_BLOCK))) return rc;
    fp.x = static_cast<const __half*>(a->x);
    fp.residual_in = static_cast<const __half*>(a->residual_in);
    fp.rms_w = static_cast<const __half*>(a->rms_w);
    fp.out = a->out;
    fp.residual_out = static_cast<__half*>(a->residual_out);
    fp.scratch = reinterpret_cast<float*>(static_cast<char*>(a->workspace) + ws_off_scratch());
    fp.counters = reinterpret_cast<unsigned*>(static_cast<char*>(a->workspace) + ws_off_counters(a->hidden, a->workspace_batch > 0 ? a->workspace_batch : 1));
    fp.eps = a->eps;
    fp.hidden = a->hidden;
    fp.ffn = a->ffn;
    fp.flags = a->flags;

    static std::once_flag once[16];
    static cudaError_t attr_err[16];
    std::call_once(once[dev & 15], [&] {
        attr_err[dev & 15] = cudaFuncSetAttribute(cfb::llama_ffn_layer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                  cfb::SmemFfn::TOTAL);
    });
    if (attr_err[dev & 15] != cudaSuccess)
        return fail((int)attr_err[dev & 15], "cudaFuncSetAttribute(ffn): %s", cudaGetErrorString(attr_err[dev & 15]));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid, 1, 1);
    cfg.blockDim = dim3(cfb::BLOCK_THREADS, 1, 1);
    cfg.dynamicSmemBytes = cfb::SmemFfn::TOTAL;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = (a->flags & CF_FLAG_PDL) ? 1 : 0;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, cfb::llama_ffn_layer_kernel, fp);
    if (e != cudaSuccess) return fail((int)e, "ffn kernel launch failed: %s", cudaGetErrorString(e));
    return 0;
}

extern "C" size_t cf_tp_exchange_bytes(int32_t hidden, int32_t tp_world) {
    if (hidden <= 0 || tp_world < 1) return 0;
    return (size_t)2 * tp_world * hidden * sizeof(uint64_t);          // [2 parities][tp_world][hidden] (float, epoch) words
}

extern "C" int cf_ipc_alloc(size_t bytes, void** dev_ptr, unsigned char handle[64]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    if (!dev_ptr || !handle || bytes == 0) return fail(CF_ERR_NULL_ARG, "cf_ipc_alloc: bad arguments");
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaSuccess) e = cudaMemset(p, 0, bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        if (p) cudaFree(p);
        return fail((int)e, "cf_ipc_alloc: %s", cudaGetErrorString(e));
    }
    memcpy(handle, &h, 64);
    *dev_ptr = p;
    return 0;
}

extern "C" int cf_ipc_open(const unsigned char handle[64], void** dev_ptr) {
    if (!dev_ptr || !handle) return fail(CF_ERR_NULL_ARG, "cf_ipc_open: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    const cudaError_t e = cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return fail((int)e, "cf_ipc_open: %s", cudaGetErrorString(e));
    return 0;
}

extern "C" int cf_ipc_close(void* dev_ptr) {
    const cudaError_t e = cudaIpcCloseMemHandle(dev_ptr);
    return e == cudaSuccess ? 0 : fail((int)e, "cf_ipc_close: %s", cudaGetErrorString(e));
}

extern "C" int cf_ipc_free(void* dev_ptr) {
    const cudaError_t e = cudaFree(dev_ptr);
    return e == cudaSuccess ? 0 : fail((int)e, "cf_ipc_free: %s", cudaGetErrorString(e));
}

extern "C" int cf_rmsnorm_launch(const void* x, const void* weight, void* out, int32_t batch, int32_t hidden, float eps,
                                 uint32_t flags, void* stream_) {
    if (!x || !weight || !out) return fail(CF_ERR_NULL_ARG, "rmsnorm: x / weight / out must be non-NULL");
    if (batch < 1 || batch > (1 << 20)) return fail(CF_ERR_BAD_SHAPE, "rmsnorm: bad batch %d", batch);
    const int slice_max = cfb::NORM_THREADS * 8 * cfb::NORM_ITERS;
    if (hidden <= 0 || hidden % (8 * cfb::NORM_CLUSTER) != 0 || hidden / cfb::NORM_CLUSTER > slice_max)
        return fail(CF_ERR_BAD_SHAPE, "rmsnorm: hidden must be a multiple of %d and <= %d (got %d)", 8 * cfb::NORM_CLUSTER,
                    slice_max * cfb::NORM_CLUSTER, hidden);
    if (!aligned16(x) || !aligned16(weight) || !aligned16(out)) return fail(CF_ERR_BAD_ALIGNMENT, "rmsnorm: tensors must be 16-byte aligned");
    if (!device_is_sm100()) return fail(CF_ERR_NO_DEVICE, "current CUDA device is not compute capability 10.x");
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(batch * cfb::NORM_CLUSTER, 1, 1);
    cfg.blockDim = dim3(cfb::NORM_THREADS, 1, 1);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = static_cast<cudaStream_t>(stream_);
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cfb::NORM_CLUSTER;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = (flags & CF_FLAG_PDL) ? 2 : 1;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, cfb::rmsnorm_cluster_kernel, static_cast<const __half*>(x),
                                             static_cast<const __half*>(weight), static_cast<__half*>(out), (int)hidden, eps);
    if (e != cudaSuccess) return fail((int)e, "rmsnorm kernel launch failed: %s", cudaGetErrorString(e));
    return 0;
}

// ---- DeepSeek-MLA half-layer: three kernels back to back on the caller's stream (deepseek_mla_kernel.cuh) ----
namespace {
constexpr size_t DS_WS_CKV = 0;                                                  // 576 floats
constexpr size_t DS_WS_OUT = 2560;                                               // 2048 floats
constexpr size_t DS_WS_CNT = DS_WS_OUT + cfb::DS_HIDDEN * 4;                     // 8 counters
constexpr size_t DS_WS_ZERO_END = DS_WS_CNT + 256;                               // everything below must start zeroed
constexpr size_t DS_WS_Q = DS_WS_ZERO_END;                                       // 16 x 576 halves
constexpr size_t DS_WS_ML = DS_WS_Q + cfb::DS_HEADS * cfb::DS_MLA * 2;           // 129 x 16 x 2 floats
constexpr size_t DS_WS_O = (DS_WS_ML + cfb::DS_STATES * cfb::DS_HEADS * 2 * 4 + 255) / 256 * 256;
constexpr size_t DS_WS_TOTAL = DS_WS_O + (size_t)cfb::DS_STATES * cfb::DS_HEADS * cfb::DS_LORA * 4;

template <typename Kern>
int ds_launch(Kern kern, int slot, int grid, int smem_max, int smem_bytes, int cluster, const cfb::DsParams& dp, bool pdl, cudaStream_t stream) {
    static std::once_flag once[3][16];
    static cudaError_t attr_err[3][16];
    int dev = 0;
    cudaGetDevice(&dev);
    std::call_once(once[slot][dev & 15], [&] {
        attr_err[slot][dev & 15] = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max);
    });
    if (attr_err[slot][dev & 15] != cudaSuccess)
        return fail((int)attr_err[slot][dev & 15], "cudaFuncSetAttribute(deepseek %d): %s", slot, cudaGetErrorString(attr_err[slot][dev & 15]));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid, 1, 1);
    cfg.blockDim = dim3(cfb::DS_THREADS, 1, 1);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute at[2];
    int n = 0;
    if (cluster > 1) {
        at[n].id = cudaLaunchAttributeClusterDimension;
        at[n].val.clusterDim.x = cluster;
        at[n].val.clusterDim.y = 1;
        at[n].val.clusterDim.z = 1;
        ++n;
    }
    if (pdl) {
        at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    cfg.attrs = at;
    cfg.numAttrs = n;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, dp);
    if (e != cudaSuccess) return fail((int)e, "deepseek kernel %d launch failed: %s", slot, cudaGetErrorString(e));
    return 0;
}
}  // namespace

extern "C" size_t cf_deepseek_workspace_bytes(void) { return DS_WS_TOTAL; }
extern "C" size_t cf_sizeof_deepseek_args(void) { return sizeof(CfDeepseekArgs); }

extern "C" int cf_deepseek_decoder_layer_launch(const CfDeepseekArgs* a, void* stream_) {
    static_assert(CF_DS_FLAG_ROPE_SCORES == cfb::DS_FLAG_ROPE_SCORES, "flag mirror");
    if (!a) return fail(CF_ERR_NULL_ARG, "args is NULL");
    if (a->hidden != cfb::DS_HIDDEN || a->n_heads != cfb::DS_HEADS)
        return fail(CF_ERR_BAD_SHAPE, "deepseek: hidden must be %d and n_heads %d (got %d, %d)", cfb::DS_HIDDEN, cfb::DS_HEADS,
                    a->hidden, a->n_heads);
    if (a->seq_len < 1 || a->seq_len > (1 << 24)) return fail(CF_ERR_BAD_SHAPE, "deepseek: seq_len must be in [1, 2^24] (got %d)", a->seq_len);
    const void* need[] = {a->x, a->w_q_nope, a->w_q_pe, a->w_uk, a->w_kv_nope, a->w_k_pe, a->w_uv, a->w_o, a->rms_input_w,
                          a->rms_ckv_w, a->cos, a->sin, a->out, a->workspace};
    for (const void* q : need)
        if (!q) return fail(CF_ERR_NULL_ARG, "deepseek: x / weights / rms weights / cos / sin / out / workspace must be non-NULL");
    if (a->seq_len > 1 && !a->ckv_cache) return fail(CF_ERR_NULL_ARG, "deepseek: ckv_cache must be non-NULL when seq_len > 1");
    const void* al[] = {a->x, a->w_q_nope, a->w_q_pe, a->w_uk, a->w_kv_nope, a->w_k_pe, a->w_uv, a->w_o, a->ckv_cache,
                        a->rms_input_w, a->rms_ckv_w, a->out, a->ckv_new, a->k_pe_new, a->workspace};
    for (const void* q : al)
        if (q && !aligned16(q)) return fail(CF_ERR_BAD_ALIGNMENT, "all tensors must be 16-byte aligned (%p)", q);
    if (!device_is_sm100()) return fail(CF_ERR_NO_DEVICE, "current CUDA device is not compute capability 10.x");

    cfb::DsParams dp;
    memset(&dp, 0, sizeof dp);
    const int H = cfb::DS_HIDDEN, NH = cfb::DS_HEADS;
    const int n_rows = a->seq_len - 1;
    int rc;
    if ((rc = get_tensor_map(&dp.tm_wq_nope, a->w_q_nope, H, (uint64_t)NH * cfb::DS_NOPE, cfb::DS_NOPE, cfb::DS_KSLICE))) return rc;
    if ((rc = get_tensor_map(&dp.tm_wq_pe, a->w_q_pe, H, (uint64_t)NH * cfb::DS_ROPE, cfb::DS_ROPE, cfb::DS_KSLICE))) return rc;
    if ((rc = get_tensor_map(&dp.tm_wuk, a->w_uk, cfb::DS_NOPE, (uint64_t)NH * cfb::DS_LORA, 64, cfb::DS_NOPE))) return rc;
    if ((rc = get_tensor_map(&dp.tm_wkv, a->w_kv_nope, H, cfb::DS_LORA, 256, cfb::DS_KV_ROWS))) return rc;
    if ((rc = get_tensor_map(&dp.tm_wk_pe, a->w_k_pe, H, cfb::DS_ROPE, cfb::DS_ROPE, cfb::DS_KV_ROWS))) return rc;
    if ((rc = get_tensor_map(&dp.tm_wuv, a->w_uv, cfb::DS_LORA, (uint64_t)NH * cfb::DS_NOPE, cfb::DS_NOPE, 64))) return rc;
    if ((rc = get_tensor_map(&dp.tm_wo, a->w_o, (uint64_t)NH * cfb::DS_NOPE, H, 256, cfb::DS_NOPE))) return rc;
    if (n_rows > 0 && (rc = get_tensor_map(&dp.tm_cache, a->ckv_cache, n_rows, cfb::DS_MLA, 64, cfb::DS_TILE_ROWS, true))) return rc;
    char* ws = static_cast<char*>(a->workspace);
    dp.x = static_cast<const __half*>(a->x);
    dp.rms_in_w = static_cast<const __half*>(a->rms_input_w);
    dp.rms_ckv_w = static_cast<const __half*>(a->rms_ckv_w);
    dp.cos = a->cos;
    dp.sin = a->sin;
    dp.out = static_cast<__half*>(a->out);
    dp.ckv_new = static_cast<__half*>(a->ckv_new);
    dp.k_pe_new = static_cast<__half*>(a->k_pe_new);
    dp.ckv_acc = reinterpret_cast<float*>(ws + DS_WS_CKV);
    dp.out_acc = reinterpret_cast<float*>(ws + DS_WS_OUT);
    dp.counters = reinterpret_cast<unsigned*>(ws + DS_WS_CNT);
    dp.q = reinterpret_cast<__half*>(ws + DS_WS_Q);
    dp.part_ml = reinterpret_cast<float*>(ws + DS_WS_ML);
    dp.part_o = reinterpret_cast<float*>(ws + DS_WS_O);
    dp.n_rows = n_rows;
    const int per = (n_rows + cfb::DS_SPLITS - 1) / cfb::DS_SPLITS;
    dp.rows_per_split = per <= cfb::DS_TILE_ROWS ? cfb::DS_TILE_ROWS : (per + cfb::DS_TILE_ROWS - 1) / cfb::DS_TILE_ROWS * cfb::DS_TILE_ROWS;
    dp.n_stages = dp.rows_per_split <= cfb::DS_TILE_ROWS ? 1 : cfb::DS_STAGES;
    dp.eps = a->eps;
    dp.scale_log2 = 1.4426950408889634f / sqrtf((float)(cfb::DS_NOPE + cfb::DS_ROPE));
    dp.flags = a->flags;

    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const bool pdl = (a->flags & CF_FLAG_PDL) != 0;
    if ((rc = ds_launch(cfb::ds_proj_kernel, 0, cfb::DS_HEADS * cfb::DS_CLUSTER, cfb::SmemDsProj::TOTAL, cfb::SmemDsProj::TOTAL, cfb::DS_CLUSTER, dp, pdl, stream))) return rc;
    // the two inner launches always overlap their prologues with the kernel before them; CF_FLAG_PDL decides only
    // whether the FIRST kernel may start before the caller's previous kernel on the stream has finished
    if ((rc = ds_launch(cfb::ds_attn_kernel, 1, cfb::DS_SPLITS, cfb::SmemDsAttn::total(cfb::DS_STAGES), cfb::SmemDsAttn::total(dp.n_stages), 1, dp, true, stream))) return rc;
    return ds_launch(cfb::ds_out_kernel, 2, cfb::DS_HEADS * cfb::DS_CLUSTER, cfb::SmemDsOut::TOTAL, cfb::SmemDsOut::TOTAL, cfb::DS_CLUSTER, dp, true, stream);
}

#ifdef CF_TRACE
extern "C" int cf_debug_set_trace(void* dev_ptr) {
    return (int)cudaMemcpyToSymbol(cfb::g_cf_trace, &dev_ptr, sizeof(void*));
}
#endif

// ------------------------------------------------------------------------------------------------
// unit-test hook for include/dsm.cuh
// ------------------------------------------------------------------------------------------------
namespace {

template <int CS, Stage ST>
__global__ void test_cluster_reduce_kernel(const float* in, float* out, int n, int repeats) {
    extern __shared__ __align__(16) uint8_t sm[];
    float* src = reinterpret_cast<float*>(sm);
    float* dst = src + n;
    uint64_t* bar = reinterpret_cast<uint64_t*>(dst + 2 * CS * n);
    const uint32_t rank = dsm::cluster_ctarank();
    const uint32_t cta = blockIdx.x;
    const uint32_t bar_u32 = dsm::smem_u32(bar);
    if (threadIdx.x == 0) {
        cluster_reduce_arm<CS>(bar_u32, n * 4);
        dsm::mbar_fence_init();
    }
    __syncthreads();
    dsm::cluster_arrive();
    dsm::cluster_wait();
    uint32_t phase = 0;
    const int n_out = (ST == Stage::QUK_DEEPSEEK) ? n * CS : n;
    for (int e = threadIdx.x; e < n_out; e += blockDim.x) out[(size_t)cta * n_out + e] = 0.f;
    for (int rep = 0; rep < repeats; ++rep) {
        const float scale = (ST == Stage::ATTN) ? 1.f : (float)(rep + 1);
        for (int e = threadIdx.x; e < n; e += blockDim.x) src[e] = in[(size_t)cta * n + e] * scale;
        cluster_reduce<CS, ST>(n * 4, threadIdx.x, ST == Stage::ATTN ? n - 4 : n, rank,
                               dsm::smem_u32(src), dsm::smem_u32(dst), bar_u32, phase, src, dst);
        if (ST == Stage::QUK_DEEPSEEK) {
            const float* g = dst + ((phase ^ 1u) & 1u) * CS * n;
            for (int e = threadIdx.x; e < n_out; e += blockDim.x) out[(size_t)cta * n_out + e] += g[e];
        } else if (ST == Stage::ATTN) {
            for (int e = threadIdx.x; e < n; e += blockDim.x) out[(size_t)cta * n + e] = src[e];
        } else {
            for (int e = threadIdx.x; e < n; e += blockDim.x) out[(size_t)cta * n + e] += src[e];
        }
        __syncthreads();
    }
    // keep this CTA's shared memory alive until every peer has finished pushing into it
    dsm::cluster_arrive();
    dsm::cluster_wait();
}

template <int CS, Stage ST>
int launch_test(const float* in, float* out, int n, int n_clusters, int repeats, cudaStream_t stream) {
    auto kern = test_cluster_reduce_kernel<CS, ST>;
    const size_t smem = (size_t)(n + 2 * CS * n) * 4 + 16;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess && CS > 8) e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return fail((int)e, "test attr: %s", cudaGetErrorString(e));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(n_clusters * CS, 1, 1);
    cfg.blockDim = dim3(128, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CS;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, kern, in, out, n, repeats);
    if (e != cudaSuccess) return fail((int)e, "test launch: %s", cudaGetErrorString(e));
    return 0;
}

template <int CS>
int launch_test_stage(const float* in, float* out, int n, int n_clusters, int stage, int repeats, cudaStream_t s) {
    switch (stage) {
        case 0: return launch_test<CS, Stage::LINEAR>(in, out, n, n_clusters, repeats, s);
        case 1: return launch_test<CS, Stage::ATTN>(in, out, n, n_clusters, repeats, s);
        case 4: return launch_test<CS, Stage::QUK_DEEPSEEK>(in, out, n, n_clusters, repeats, s);
        default: return fail(CF_ERR_BAD_VARIANT, "stage %d not testable", stage);
    }
}

}  // namespace

extern "C" int cf_test_cluster_reduce(const float* in, float* out, int32_t n, int32_t cluster_size,
                                      int32_t n_clusters, int32_t stage, int32_t repeats, void* stream_) {
    if (!in || !out) return fail(CF_ERR_NULL_ARG, "in/out NULL");
    if (n <= 0 || n % 4 != 0 || n > 2048) return fail(CF_ERR_BAD_SHAPE, "n must be a multiple of 4 in (0, 2048]");
    if (!device_is_sm100()) return fail(CF_ERR_NO_DEVICE, "current CUDA device is not compute capability 10.x");
    cudaStream_t s = static_cast<cudaStream_t>(stream_);
    switch (cluster_size) {
        case 1: return launch_test_stage<1>(in, out, n, n_clusters, stage, repeats, s);
        case 2: return launch_test_stage<2>(in, out, n, n_clusters, stage, repeats, s);
        case 4: return launch_test_stage<4>(in, out, n, n_clusters, stage, repeats, s);
        case 8: return launch_test_stage<8>(in, out, n, n_clusters, stage, repeats, s);
        case 16: return launch_test_stage<16>(in, out, n, n_clusters, stage, repeats, s);
        default: return fail(CF_ERR_BAD_SHAPE, "cluster_size must be 1, 2, 4, 8 or 16");
    }
}
