/*
 * PyTorch C++ extension `_clusterfusion`: the reference's Python-facing operator surface
 * (/root/reference/include/pybind.cpp:108-123) on top of the torch-free C ABI
 * (include/clusterfusion_b200.h).  Same names, same positional argument order, same return
 * conventions as the reference:
 *
 *   llama_decoder_layer(input, weight_qkv, weight_o, k_cache, v_cache, rms_w, cos, sin)
 *        -> (o [1,hidden], k [1,H,128], v [1,H,128])                      pybind.cpp:3-12
 *   llama_decoder_layer_sglang(input, residual, weight_qkv, weight_o, k_cache, v_cache, rms_w,
 *                              eps, cos, sin) -> (o, residual, k, v)      pybind.cpp:14-25
 *        (`residual` is updated in place and returned, llama_kernel_sglang_dispatch.cu:150)
 *   llama_decoder_layer_batch_decode_sglang(output, residual_output, input, residual, weight_qkv,
 *        weight_o, paged_kv_indptr, paged_kv_indices, k_cache_ptrs, v_cache_ptrs, layer_id,
 *        rms_w, eps, positions, cos_sin) -> None                           pybind.cpp:27-43
 *   llama_decoder_layer(<the same 15 arguments>)  -- the call the reference README shows (README.md:55-75)
 *
 * What this shim does that the reference's dispatch files do not: TORCH_CHECKs on
 * device / dtype / contiguity / shape, launches on the CURRENT stream of the input's device,
 * never synchronises, allocates outputs with empty() (no memset kernels), and raises RuntimeError
 * with the C ABI's message on failure.
 */
#include <torch/extension.h>
#include <ATen/cuda/EmptyTensor.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>

#include <atomic>
#include <map>
#include <mutex>
#include <tuple>
#include <utility>
#include <vector>

#include "../../include/clusterfusion_b200.h"

namespace {

using torch::Tensor;

void check_cuda_contig(const Tensor& t, const char* name, c10::ScalarType dtype) {
    TORCH_CHECK(t.is_cuda(), name, " must be a CUDA tensor");
    TORCH_CHECK(t.is_contiguous(), name, " must be contiguous");
    TORCH_CHECK(t.scalar_type() == dtype, name, " must have dtype ", dtype, " (got ", t.scalar_type(), ")");
}

// One zero-initialised workspace per (device, stream, hidden): memset exactly once, opaque afterwards.  Its layout depends
// on the batch it was sized for, so that batch is remembered and passed as CfLlamaArgs::workspace_batch; a larger batch gets
// a fresh (zeroed) workspace sized for it, which then also serves the smaller ones.
// A first call that happens INSIDE a stream capture must not record the allocation's memset into the graph (every replay
// would re-zero the workspace, and a buffer from the graph's private pool would die with the graph): in that case the
// workspace comes from a plain cudaMalloc + synchronous cudaMemset issued in relaxed capture mode, outside the graph.
struct Workspace { Tensor buf; int batch; };
std::mutex g_ws_mu;
std::map<std::tuple<int, void*, int>, Workspace> g_ws_cache;
std::map<std::tuple<int, void*>, Tensor> g_ds_ws_cache;

Tensor zeroed_bytes(const Tensor& like, size_t need, cudaStream_t stream) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    const bool capturing = cudaStreamIsCapturing(stream, &st) == cudaSuccess && st != cudaStreamCaptureStatusNone;
    if (!capturing)
        return torch::zeros({(int64_t)need}, torch::TensorOptions().dtype(torch::kUInt8).device(like.device()));
    // relaxed mode: this thread may call cudaMalloc etc. while its stream is capturing.  The memset runs on a private
    // NON-BLOCKING stream (the legacy stream would implicitly join the capturing stream and invalidate the capture).
    cudaStreamCaptureMode mode = cudaStreamCaptureModeRelaxed;
    cudaThreadExchangeStreamCaptureMode(&mode);
    void* p = nullptr;
    cudaStream_t side = nullptr;
    cudaError_t e = cudaMalloc(&p, need);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMemsetAsync(p, 0, need, side);
    if (e == cudaSuccess) e = cudaStreamSynchronize(side);
    if (side) cudaStreamDestroy(side);
    cudaThreadExchangeStreamCaptureMode(&mode);
    TORCH_CHECK(e == cudaSuccess, "clusterfusion_b200: workspace allocation during stream capture failed: ", cudaGetErrorString(e));
    return torch::from_blob(p, {(int64_t)need}, [](void* q) { cudaFree(q); },
                            torch::TensorOptions().dtype(torch::kUInt8).device(like.device()));
}

Workspace workspace_for(const Tensor& like, int hidden, int batch, cudaStream_t stream) {
    std::lock_guard<std::mutex> lk(g_ws_mu);
    auto key = std::make_tuple((int)like.get_device(), (void*)stream, hidden);
    auto it = g_ws_cache.find(key);
    if (it == g_ws_cache.end() || it->second.batch < batch) {
        Workspace w{zeroed_bytes(like, cf_llama_workspace_bytes(hidden, batch), stream), batch};
        g_ws_cache[key] = w;
        return w;
    }
    return it->second;
}

// Programmatic dependent launch for every op issued through this module (off by default; process-wide switch -- set it once
// at start-up, not per call from several threads).  Contract: see CF_FLAG_PDL in include/clusterfusion_b200.h -- the weights,
// KV cache / pools, page tables (indptr / indices), positions and pool-pointer tables of a call are not written by kernels
// still in flight on the stream: the kernel reads them BEFORE it waits for the previous kernel.
std::atomic<bool> g_pdl{false};
// debug mode: after every launch, read the workspace's sticky error word back (synchronises) and raise if an in-kernel
// exchange poll timed out -- see cf_workspace_status in the C ABI header
std::atomic<bool> g_check_status{false};

// Output tensors of the forms that allocate (8- and 10-argument): straight from the caching allocator, without a trip through
// the dispatcher -- three torch::empty calls were ~3.6 us of the 8-argument form's 10.9 us host cost per call
// (tools/host_overhead_probe.py), in a chat loop that is host-bound.
inline Tensor empty_half_like(const Tensor& like, at::IntArrayRef sizes) {
    return at::detail::empty_cuda(sizes, at::kHalf, like.device(), std::nullopt);
}

// [a, a+na) and [b, b+nb) overlap?
bool overlaps(const Tensor& a, const Tensor& b) {
    const char* pa = static_cast<const char*>(a.data_ptr());
    const char* pb = static_cast<const char*>(b.data_ptr());
    return pa < pb + b.nbytes() && pb < pa + a.nbytes();
}

// Host copies of the device pointer tables of the paged form, fetched once per table (one blocking copy of 8 bytes per layer,
// never during stream capture).  They only enable the tiled / gather4 KV fast paths: the kernel re-checks the address against
// the device table, so a stale copy costs speed, not correctness.
const std::vector<uint64_t>* host_ptr_table(const Tensor& t, cudaStream_t stream) {
    static std::mutex mu;
    static std::map<std::pair<const void*, int64_t>, std::vector<uint64_t>> cache;
    std::lock_guard<std::mutex> lk(mu);
    const auto key = std::make_pair((const void*)t.data_ptr(), (int64_t)t.numel());
    auto it = cache.find(key);
    if (it != cache.end()) return &it->second;
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &st) != cudaSuccess || st != cudaStreamCaptureStatusNone) return nullptr;
    std::vector<uint64_t> h((size_t)t.numel());
    if (cudaMemcpyAsync(h.data(), t.data_ptr(), 8 * (size_t)t.numel(), cudaMemcpyDeviceToHost, stream) != cudaSuccess ||
        cudaStreamSynchronize(stream) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return &(cache[key] = std::move(h));
}

void run(const CfLlamaArgs& a, cudaStream_t stream) {
    const int rc = cf_llama_decoder_layer_launch(&a, stream);
    TORCH_CHECK(rc == 0, "clusterfusion_b200: launch failed (", rc, "): ", cf_last_error_string());
    if (g_check_status.load(std::memory_order_relaxed)) {
        uint32_t st = 0;
        const int rs = cf_workspace_status(a.workspace, stream, &st);
        TORCH_CHECK(rs == 0, "clusterfusion_b200: cf_workspace_status failed (", rs, "): ", cf_last_error_string());
        TORCH_CHECK(st == 0, "clusterfusion_b200: an exchange poll inside the kernel timed out (CTAs of a group not co-resident, "
                             "a stalled peer rank, or one workspace shared by two streams); the results of this launch are invalid");
    }
}

std::tuple<Tensor, Tensor, Tensor> llama_decoder_layer(
    Tensor input, Tensor weight_qkv, Tensor weight_o, Tensor k_cache, Tensor v_cache,
    Tensor rms_input_weight, Tensor cos, Tensor sin)
{
    check_cuda_contig(input, "input", torch::kHalf);
    check_cuda_contig(weight_qkv, "weight_qkv", torch::kHalf);
    check_cuda_contig(weight_o, "weight_o", torch::kHalf);
    check_cuda_contig(k_cache, "k_cache", torch::kHalf);
    check_cuda_contig(v_cache, "v_cache", torch::kHalf);
    check_cuda_contig(rms_input_weight, "rms_input_weight", torch::kHalf);
    check_cuda_contig(cos, "cos", torch::kFloat);
    check_cuda_contig(sin, "sin", torch::kFloat);
    const int64_t hidden = input.size(-1);
    TORCH_CHECK(input.numel() == hidden, "input must hold one token: [1, hidden] or [1, 1, hidden]");
    TORCH_CHECK(weight_qkv.dim() == 2 && weight_qkv.size(0) == 3 * hidden && weight_qkv.size(1) == hidden,
                "weight_qkv must be [3*hidden, hidden] = [Wq^T; Wk^T; Wv^T]");
    TORCH_CHECK(weight_o.dim() == 2 && weight_o.size(0) == hidden && weight_o.size(1) == hidden,
                "weight_o must be [hidden, hidden] = Wo^T");
    TORCH_CHECK(k_cache.dim() == 2 && k_cache.size(1) == hidden, "k_cache must be [kv_len, hidden]");
    TORCH_CHECK(v_cache.sizes() == k_cache.sizes(), "v_cache must match k_cache");
    TORCH_CHECK(rms_input_weight.numel() == hidden, "rms_input_weight must be [hidden]");
    TORCH_CHECK(cos.numel() >= 128 && sin.numel() >= 128, "cos / sin must be [1, 128] (pair-repeated)");
    const int n_heads = (int)(hidden / 128);

    const c10::cuda::CUDAGuard guard(input.device());
    cudaStream_t stream = c10::cuda::getCurrentCUDAStream(input.get_device()).stream();
    Tensor o = empty_half_like(input, {1, hidden});
    Tensor k = empty_half_like(input, {1, n_heads, 128});
    Tensor v = empty_half_like(input, {1, n_heads, 128});
    Workspace ws = workspace_for(input, (int)hidden, 1, stream);

    CfLlamaArgs a{};
    a.variant = CF_VARIANT_CHAT;
    a.hidden = (int)hidden; a.n_q_heads = n_heads; a.n_kv_heads = n_heads; a.head_dim = 128; a.batch = 1;
    a.kv_len = (uint32_t)k_cache.size(0);
    a.eps = 1e-6f;      // fixed by the reference's 8-argument kernel (kernel.cuh:58)
    a.x = input.data_ptr(); a.w_qkv = weight_qkv.data_ptr(); a.w_o = weight_o.data_ptr();
    a.rms_w = rms_input_weight.data_ptr();
    a.out = o.data_ptr(); a.k_new = k.data_ptr(); a.v_new = v.data_ptr();
    a.k_cache = k_cache.data_ptr(); a.v_cache = v_cache.data_ptr();
    a.cos = cos.data_ptr<float>(); a.sin = sin.data_ptr<float>();
    a.workspace = ws.buf.data_ptr(); a.workspace_batch = ws.batch;
    if (g_pdl.load(std::memory_order_relaxed)) a.flags |= CF_FLAG_PDL;
    run(a, stream);
    return std::make_tuple(o, k, v);
}

std::tuple<Tensor, Tensor, Tensor, Tensor> llama_decoder_layer_sglang(
    Tensor input, Tensor residual, Tensor weight_qkv, Tensor weight_o, Tensor k_cache, Tensor v_cache,
    Tensor rms_input_weight, double eps, Tensor cos, Tensor sin)
{
    check_cuda_contig(input, "input", torch::kHalf);
    check_cuda_contig(residual, "residual", torch::kHalf);
    check_cuda_contig(weight_qkv, "weight_qkv", torch::kHalf);
    check_cuda_contig(weight_o, "weight_o", torch::kHalf);
    check_cuda_contig(k_cache, "k_cache", torch::kHalf);
    check_cuda_contig(v_cache, "v_cache", torch::kHalf);
    check_cuda_contig(rms_input_weight, "rms_input_weight", torch::kHalf);
    check_cuda_contig(cos, "cos", torch::kFloat);
    check_cuda_contig(sin, "sin", torch::kFloat);
    const int64_t hidden = input.size(-1);
    TORCH_CHECK(input.numel() == hidden && residual.numel() == hidden, "input / residual must hold one token");
    TORCH_CHECK(weight_o.dim() == 2 && weight_o.size(0) == hidden && weight_o.size(1) % 128 == 0,
                "weight_o must be [hidden, n_heads*128] (nn.Linear layout)");
    const int64_t qd = weight_o.size(1);
    TORCH_CHECK(weight_qkv.dim() == 2 && weight_qkv.size(1) == hidden && weight_qkv.size(0) > qd &&
                (weight_qkv.size(0) - qd) % 256 == 0, "weight_qkv must be [(n_q + 2*n_kv)*128, hidden]");
    const int64_t kvd = (weight_qkv.size(0) - qd) / 2;
    TORCH_CHECK(k_cache.dim() == 2 && k_cache.size(1) == kvd, "k_cache must be [kv_len, n_kv*128]");
    TORCH_CHECK(v_cache.sizes() == k_cache.sizes(), "v_cache must match k_cache");
    TORCH_CHECK(rms_input_weight.numel() == hidden, "rms_input_weight must be [hidden]");
    TORCH_CHECK(cos.numel() >= 64 && sin.numel() >= 64, "cos / sin must hold at least head_dim/2 = 64 floats");
    TORCH_CHECK(!overlaps(input, residual), "input and residual must not overlap (residual is updated in place)");

    const c10::cuda::CUDAGuard guard(input.device());
    cudaStream_t stream = c10::cuda::getCurrentCUDAStream(input.get_device()).stream();
    Tensor o = empty_half_like(input, {1, hidden});
    Tensor k = empty_half_like(input, {1, kvd / 128, 128});
    Tensor v = empty_half_like(input, {1, kvd / 128, 128});
    Workspace ws = workspace_for(input, (int)hidden, 1, stream);

    CfLlamaArgs a{};
    a.variant = CF_VARIANT_SGLANG;
    a.hidden = (int)hidden; a.n_q_heads = (int)(qd / 128); a.n_kv_heads = (int)(kvd / 128); a.head_dim = 128;
    a.batch = 1;
    a.kv_len = (uint32_t)k_cache.size(0);
    a.eps = (float)eps;
    a.x = input.data_ptr(); a.residual_in = residual.data_ptr(); a.residual_out = residual.data_ptr();
    a.w_qkv = weight_qkv.data_ptr(); a.w_o = weight_o.data_ptr(); a.rms_w = rms_input_weight.data_ptr();
    a.out = o.data_ptr(); a.k_new = k.data_ptr(); a.v_new = v.data_ptr();
    a.k_cache = k_cache.data_ptr(); a.v_cache = v_cache.data_ptr();
    a.cos = cos.data_ptr<float>(); a.sin = sin.data_ptr<float>();
    a.workspace = ws.buf.data_ptr(); a.workspace_batch = ws.batch;
    if (g_pdl.load(std::memory_order_relaxed)) a.flags |= CF_FLAG_PDL;
    run(a, stream);
    return std::make_tuple(o, residual, k, v);
}

void llama_decoder_layer_batch_decode_sglang(
    Tensor output, Tensor residual_output, Tensor input, Tensor residual, Tensor weight_qkv, Tensor weight_o,
    Tensor paged_kv_indptr, Tensor paged_kv_indices, Tensor k_cache_ptrs, Tensor v_cache_ptrs, int64_t layer_id,
    Tensor rms_input_weight, double eps, Tensor positions, Tensor cos_sin)
{
    check_cuda_contig(output, "output", torch::kHalf);
    check_cuda_contig(residual_output, "residual_output", torch::kHalf);
    check_cuda_contig(input, "input", torch::kHalf);
    check_cuda_contig(residual, "residual", torch::kHalf);
    check_cuda_contig(weight_qkv, "weight_qkv", torch::kHalf);
    check_cuda_contig(weight_o, "weight_o", torch::kHalf);
    check_cuda_contig(paged_kv_indptr, "paged_kv_indptr", torch::kInt);
    check_cuda_contig(paged_kv_indices, "paged_kv_indices", torch::kInt);
    check_cuda_contig(rms_input_weight, "rms_input_weight", torch::kHalf);
    check_cuda_contig(positions, "positions", torch::kLong);
    check_cuda_contig(cos_sin, "cos_sin", torch::kFloat);
    TORCH_CHECK(k_cache_ptrs.is_cuda() && v_cache_ptrs.is_cuda() && k_cache_ptrs.is_contiguous() &&
                v_cache_ptrs.is_contiguous() && k_cache_ptrs.element_size() == 8 && v_cache_ptrs.element_size() == 8,
                "k_cache_ptrs / v_cache_ptrs must be contiguous CUDA tensors of 64-bit pointers");
    TORCH_CHECK(input.dim() == 2, "input must be [batch, hidden]");
    const int64_t bs = input.size(0), hidden = input.size(1);
    TORCH_CHECK(residual.sizes() == input.sizes() && output.sizes() == input.sizes() &&
                residual_output.sizes() == input.sizes(), "output / residual_output / residual must be [batch, hidden]");
    TORCH_CHECK(weight_o.dim() == 2 && weight_o.size(0) == hidden && weight_o.size(1) % 128 == 0,
                "weight_o must be [hidden, n_heads*128]");
    const int64_t qd = weight_o.size(1);
    TORCH_CHECK(weight_qkv.dim() == 2 && weight_qkv.size(1) == hidden && weight_qkv.size(0) > qd &&
                (weight_qkv.size(0) - qd) % 256 == 0, "weight_qkv must be [(n_q + 2*n_kv)*128, hidden]");
    const int64_t kvd = (weight_qkv.size(0) - qd) / 2;
    TORCH_CHECK(paged_kv_indptr.numel() == bs + 1, "paged_kv_indptr must be [batch + 1]");
    TORCH_CHECK(positions.numel() == bs, "positions must be [batch]");
    TORCH_CHECK(layer_id >= 0 && layer_id < k_cache_ptrs.numel() && layer_id < v_cache_ptrs.numel(), "layer_id out of range");
    TORCH_CHECK(cos_sin.dim() == 2 && cos_sin.size(1) == 128, "cos_sin must be [max_pos, 128] = [cos(64) | sin(64)]");
    // the epilogues write `output` / `residual_output` while other CTAs may still be reading `input` / `residual`: the only
    // supported aliasing is residual_output == residual exactly (in place; written by the request's very last CTA)
    TORCH_CHECK(!overlaps(output, input) && !overlaps(output, residual) && !overlaps(output, residual_output) &&
                !overlaps(residual_output, input), "output / residual_output must not overlap input / residual / each other");
    TORCH_CHECK(residual_output.data_ptr() == residual.data_ptr() || !overlaps(residual_output, residual),
                "residual_output may alias residual only exactly (same tensor, in place)");

    const c10::cuda::CUDAGuard guard(input.device());
    cudaStream_t stream = c10::cuda::getCurrentCUDAStream(input.get_device()).stream();
    Workspace ws = workspace_for(input, (int)hidden, (int)bs, stream);

    CfLlamaArgs a{};
    a.variant = CF_VARIANT_PAGED;
    a.hidden = (int)hidden; a.n_q_heads = (int)(qd / 128); a.n_kv_heads = (int)(kvd / 128); a.head_dim = 128;
    a.batch = (int)bs;
    a.layer_id = (int)layer_id;
    a.eps = (float)eps;
    a.x = input.data_ptr(); a.residual_in = residual.data_ptr(); a.residual_out = residual_output.data_ptr();
    a.w_qkv = weight_qkv.data_ptr(); a.w_o = weight_o.data_ptr(); a.rms_w = rms_input_weight.data_ptr();
    a.out = output.data_ptr();
    a.indptr = paged_kv_indptr.data_ptr<int>(); a.indices = paged_kv_indices.data_ptr<int>();
    a.k_pool_ptrs = static_cast<const uint64_t*>(k_cache_ptrs.data_ptr());
    a.v_pool_ptrs = static_cast<const uint64_t*>(v_cache_ptrs.data_ptr());
    a.positions = positions.data_ptr<int64_t>();
    a.cos = cos_sin.data_ptr<float>();
    a.workspace = ws.buf.data_ptr(); a.workspace_batch = ws.batch;
    if (const auto* hk = host_ptr_table(k_cache_ptrs, stream))
        if (const auto* hv = host_ptr_table(v_cache_ptrs, stream)) {
            const uint64_t kb = (*hk)[(size_t)layer_id], vb = (*hv)[(size_t)layer_id];
            if (kb && vb && kb % 16 == 0 && vb % 16 == 0) {
                a.k_cache = reinterpret_cast<const void*>(kb);
                a.v_cache = reinterpret_cast<const void*>(vb);
            }
        }
    if (g_pdl.load(std::memory_order_relaxed)) a.flags |= CF_FLAG_PDL;
    run(a, stream);
}

void llama_ffn_layer_out(Tensor output, Tensor residual_output, Tensor input, Tensor residual, Tensor weight_gate_up,
                         Tensor weight_down_t, Tensor rms_weight, double eps)
{
    check_cuda_contig(output, "output", torch::kHalf);
    check_cuda_contig(residual_output, "residual_output", torch::kHalf);
    check_cuda_contig(input, "input", torch::kHalf);
    check_cuda_contig(residual, "residual", torch::kHalf);
    check_cuda_contig(weight_gate_up, "weight_gate_up", torch::kHalf);
    check_cuda_contig(weight_down_t, "weight_down_t", torch::kHalf);
    check_cuda_contig(rms_weight, "rms_weight", torch::kHalf);
    const int64_t hidden = input.size(-1);
    TORCH_CHECK(input.numel() == hidden && residual.numel() == hidden && output.numel() == hidden &&
                residual_output.numel() == hidden, "input / residual / output / residual_output must hold one token");
    TORCH_CHECK(weight_down_t.dim() == 2 && weight_down_t.size(1) == hidden, "weight_down_t must be [ffn, hidden] (W2 transposed)");
    const int64_t ffn = weight_down_t.size(0);
    TORCH_CHECK(weight_gate_up.dim() == 2 && weight_gate_up.size(0) == 2 * ffn && weight_gate_up.size(1) == hidden,
                "weight_gate_up must be [2*ffn, hidden] = [W1; W3]");
    TORCH_CHECK(rms_weight.numel() == hidden, "rms_weight must be [hidden]");
    TORCH_CHECK(!overlaps(output, input) && !overlaps(output, residual) && !overlaps(output, residual_output) &&
                !overlaps(residual_output, input), "output / residual_output must not overlap input / residual / each other");
    TORCH_CHECK(residual_output.data_ptr() == residual.data_ptr() || !overlaps(residual_output, residual),
                "residual_output may alias residual only exactly (same tensor, in place)");
    const c10::cuda::CUDAGuard guard(input.device());
    cudaStream_t stream = c10::cuda::getCurrentCUDAStream(input.get_device()).stream();
    Workspace ws = workspace_for(input, (int)hidden, 1, stream);
    CfFfnArgs a{};
    a.flags = g_pdl.load(std::memory_order_relaxed) ? CF_FLAG_PDL : 0u;
    a.hidden = (int)hidden; a.ffn = (int)ffn; a.eps = (float)eps;
    a.x = input.data_ptr(); a.residual_in = residual.data_ptr();
    a.w_gate_up = weight_gate_up.data_ptr(); a.w_down_t = weight_down_t.data_ptr(); a.rms_w = rms_weight.data_ptr();
    a.out = output.data_ptr(); a.residual_out = residual_output.data_ptr(); a.workspace = ws.buf.data_ptr();
    a.workspace_batch = ws.batch;
    const int rc = cf_llama_ffn_launch(&a, stream);
    TORCH_CHECK(rc == 0, "clusterfusion_b200: ffn launch failed (", rc, "): ", cf_last_error_string());
}

std::tuple<Tensor, Tensor> llama_ffn_layer(Tensor input, Tensor residual, Tensor weight_gate_up, Tensor weight_down_t,
                                           Tensor rms_weight, double eps)
{
    TORCH_CHECK(input.is_cuda(), "input must be a CUDA tensor");
    Tensor out = torch::empty({1, input.size(-1)}, input.options());
    Tensor res_out = torch::empty({1, input.size(-1)}, input.options());
    llama_ffn_layer_out(out, res_out, input, residual, weight_gate_up, weight_down_t, rms_weight, eps);
    return std::make_tuple(out, res_out);
}

// rmsnorm(input [batch, hidden], weight [hidden]) -> fp16 [batch, hidden]; eps = 1e-6 as in the reference kernel
// (/root/reference/include/H100/norm/kernel.cuh:28; signature pybind.cpp:61-64)
Tensor rmsnorm(Tensor input, Tensor weight) {
    check_cuda_contig(input, "input", torch::kHalf);
    check_cuda_contig(weight, "weight", torch::kHalf);
    TORCH_CHECK(input.dim() == 2 && weight.numel() == input.size(1), "rmsnorm: input [batch, hidden], weight [hidden]");
    const c10::cuda::CUDAGuard guard(input.device());
    cudaStream_t stream = c10::cuda::getCurrentCUDAStream(input.get_device()).stream();
    Tensor out = torch::empty_like(input);
    const int rc = cf_rmsnorm_launch(input.data_ptr(), weight.data_ptr(), out.data_ptr(), (int)input.size(0), (int)input.size(1),
                                     1e-6f, g_pdl.load(std::memory_order_relaxed) ? CF_FLAG_PDL : 0u, stream);
    TORCH_CHECK(rc == 0, "clusterfusion_b200: rmsnorm launch failed (", rc, "): ", cf_last_error_string());
    return out;
}

// deepseek_decoder_layer(input, weight_q_nope, weight_q_pe, weight_uk, weight_kv_nope, weight_k_pe, weight_uv, weight_o,
//                        ckv_cache, rms_input_weight, rms_ckv_weight, cos, sin) -> fp16 [1, hidden]
// (/root/reference/include/pybind.cpp:45-59, :113; deepseek_kernel_dispatch.cu:4-18).  seq_len = ckv_cache.size(0).
Tensor deepseek_workspace_for(const Tensor& like, cudaStream_t stream) {
    std::lock_guard<std::mutex> lk(g_ws_mu);
    auto key = std::make_tuple((int)like.get_device(), (void*)stream);
    auto it = g_ds_ws_cache.find(key);
    if (it != g_ds_ws_cache.end()) return it->second;
    Tensor w = zeroed_bytes(like, cf_deepseek_workspace_bytes(), stream);
    g_ds_ws_cache[key] = w;
    return w;
}

Tensor deepseek_run(Tensor input, Tensor weight_q_nope, Tensor weight_q_pe, Tensor weight_uk, Tensor weight_kv_nope,
                    Tensor weight_k_pe, Tensor weight_uv, Tensor weight_o, Tensor ckv_cache, Tensor rms_input_weight,
                    Tensor rms_ckv_weight, Tensor cos, Tensor sin, bool rope_scores, Tensor* ckv_new, Tensor* k_pe_new)
{
    const std::pair<const Tensor*, const char*> halves[] = {
        {&input, "input"}, {&weight_q_nope, "weight_q_nope"}, {&weight_q_pe, "weight_q_pe"}, {&weight_uk, "weight_uk"},
        {&weight_kv_nope, "weight_kv_nope"}, {&weight_k_pe, "weight_k_pe"}, {&weight_uv, "weight_uv"}, {&weight_o, "weight_o"},
        {&ckv_cache, "ckv_cache"}, {&rms_input_weight, "rms_input_weight"}, {&rms_ckv_weight, "rms_ckv_weight"}};
    for (const auto& t : halves) check_cuda_contig(*t.first, t.second, torch::kHalf);
    check_cuda_contig(cos, "cos", torch::kFloat);
    check_cuda_contig(sin, "sin", torch::kFloat);
    const int64_t hidden = input.size(-1);
    TORCH_CHECK(input.numel() == hidden, "deepseek_decoder_layer: input must be [1, hidden]");
    TORCH_CHECK(weight_q_nope.dim() == 2 && weight_q_nope.size(0) == hidden && weight_q_nope.size(1) % 128 == 0,
                "weight_q_nope must be [hidden, n_heads*128]");
    const int64_t nh = weight_q_nope.size(1) / 128;
    TORCH_CHECK(weight_q_pe.dim() == 2 && weight_q_pe.size(0) == hidden && weight_q_pe.size(1) == nh * 64, "weight_q_pe must be [hidden, n_heads*64]");
    TORCH_CHECK(weight_uk.dim() == 2 && weight_uk.size(0) == 128 && weight_uk.size(1) == nh * 512, "weight_uk must be [128, n_heads*512]");
    TORCH_CHECK(weight_kv_nope.dim() == 2 && weight_kv_nope.size(0) == hidden && weight_kv_nope.size(1) == 512, "weight_kv_nope must be [hidden, 512]");
    TORCH_CHECK(weight_k_pe.dim() == 2 && weight_k_pe.size(0) == hidden && weight_k_pe.size(1) == 64, "weight_k_pe must be [hidden, 64]");
    TORCH_CHECK(weight_uv.dim() == 2 && weight_uv.size(0) == 512 && weight_uv.size(1) == nh * 128, "weight_uv must be [512, n_heads*128]");
    TORCH_CHECK(weight_o.dim() == 2 && weight_o.size(0) == nh * 128 && weight_o.size(1) == hidden, "weight_o must be [n_heads*128, hidden]");
    TORCH_CHECK(ckv_cache.dim() == 2 && ckv_cache.size(0) >= 1 && ckv_cache.size(1) == 576, "ckv_cache must be [seq_len >= 1, 576]");
    TORCH_CHECK(rms_input_weight.numel() == hidden && rms_ckv_weight.numel() == 512, "rms weights must be [hidden] and [512]");
    TORCH_CHECK(cos.numel() >= 64 && sin.numel() >= 64, "cos / sin must hold 64 floats");
    const c10::cuda::CUDAGuard guard(input.device());
    cudaStream_t stream = c10::cuda::getCurrentCUDAStream(input.get_device()).stream();
    Tensor ws = deepseek_workspace_for(input, stream);
    Tensor out = torch::empty({1, hidden}, input.options());
    CfDeepseekArgs a;
    memset(&a, 0, sizeof a);
    a.flags = (g_pdl.load(std::memory_order_relaxed) ? CF_FLAG_PDL : 0u) | (rope_scores ? CF_DS_FLAG_ROPE_SCORES : 0u);
    a.hidden = (int32_t)hidden;
    a.n_heads = (int32_t)nh;
    a.seq_len = (int32_t)ckv_cache.size(0);
    a.eps = 1e-6f;                              // compiled into the reference kernel (kernel.cuh:46)
    a.x = input.data_ptr();
    a.w_q_nope = weight_q_nope.data_ptr();
    a.w_q_pe = weight_q_pe.data_ptr();
    a.w_uk = weight_uk.data_ptr();
    a.w_kv_nope = weight_kv_nope.data_ptr();
    a.w_k_pe = weight_k_pe.data_ptr();
    a.w_uv = weight_uv.data_ptr();
    a.w_o = weight_o.data_ptr();
    a.ckv_cache = ckv_cache.data_ptr();
    a.rms_input_w = rms_input_weight.data_ptr();
    a.rms_ckv_w = rms_ckv_weight.data_ptr();
    a.cos = cos.data_ptr<float>();
    a.sin = sin.data_ptr<float>();
    a.out = out.data_ptr();
    if (ckv_new) {
        *ckv_new = torch::empty({512}, input.options());
        *k_pe_new = torch::empty({64}, input.options());
        a.ckv_new = ckv_new->data_ptr();
        a.k_pe_new = k_pe_new->data_ptr();
    }
    a.workspace = ws.data_ptr();
    const int rc = cf_deepseek_decoder_layer_launch(&a, stream);
    TORCH_CHECK(rc == 0, "clusterfusion_b200: deepseek launch failed (", rc, "): ", cf_last_error_string());
    return out;
}

Tensor deepseek_decoder_layer(Tensor input, Tensor weight_q_nope, Tensor weight_q_pe, Tensor weight_uk, Tensor weight_kv_nope,
                              Tensor weight_k_pe, Tensor weight_uv, Tensor weight_o, Tensor ckv_cache, Tensor rms_input_weight,
                              Tensor rms_ckv_weight, Tensor cos, Tensor sin)
{
    return deepseek_run(input, weight_q_nope, weight_q_pe, weight_uk, weight_kv_nope, weight_k_pe, weight_uv, weight_o, ckv_cache,
                        rms_input_weight, rms_ckv_weight, cos, sin, false, nullptr, nullptr);
}

// extended form: (out, ckv_new [512], k_pe_new [64]); rope_scores adds the q_pe . k_pe term the reference kernel leaves out
std::tuple<Tensor, Tensor, Tensor> deepseek_decoder_layer_ex(Tensor input, Tensor weight_q_nope, Tensor weight_q_pe, Tensor weight_uk,
                                                             Tensor weight_kv_nope, Tensor weight_k_pe, Tensor weight_uv, Tensor weight_o,
                                                             Tensor ckv_cache, Tensor rms_input_weight, Tensor rms_ckv_weight, Tensor cos,
                                                             Tensor sin, bool rope_scores)
{
    Tensor ckv_new, k_pe_new;
    Tensor out = deepseek_run(input, weight_q_nope, weight_q_pe, weight_uk, weight_kv_nope, weight_k_pe, weight_uv, weight_o, ckv_cache,
                              rms_input_weight, rms_ckv_weight, cos, sin, rope_scores, &ckv_new, &k_pe_new);
    return std::make_tuple(out, ckv_new, k_pe_new);
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.doc() = "clusterfusion_b200: B200-native fused Llama decoder attention half-layer (sm_100a)";
    m.def("llama_decoder_layer", &llama_decoder_layer, "");
    // the README example of the reference calls the 15-argument paged form under this name
    m.def("llama_decoder_layer", &llama_decoder_layer_batch_decode_sglang, "");
    m.def("llama_decoder_layer_sglang", &llama_decoder_layer_sglang, "");
    m.def("llama_decoder_layer_batch_decode_sglang", &llama_decoder_layer_batch_decode_sglang, "");
    // fused FFN half-layer (new op; the reference's FFN is eager PyTorch): (out, residual_out) = f(input, residual, [W1;W3], W2^T, w, eps)
    m.def("llama_ffn_layer", &llama_ffn_layer, "");
    m.def("llama_ffn_layer_out", &llama_ffn_layer_out, "");
    m.def("rmsnorm", &rmsnorm, "");
    m.def("deepseek_decoder_layer", &deepseek_decoder_layer, "");
    m.def("deepseek_decoder_layer_ex", &deepseek_decoder_layer_ex, "");
    m.def("set_pdl", [](bool on) { g_pdl.store(on); },
          "enable / disable programmatic dependent launch for all ops of this module (process-wide; contract: weights, KV "
          "caches / pools, page tables and positions of a call are not written by kernels still in flight on the stream)");
    m.def("get_pdl", []() { return g_pdl.load(); });
    m.def("set_check_status", [](bool on) { g_check_status.store(on); },
          "debug mode: after every launch read the workspace's sticky error word back (synchronises) and raise on an "
          "in-kernel exchange time-out");
    m.def("workspace_status", []() {
              // sticky error words of every workspace this module owns (synchronises their streams): 0 = all launches so far
              // completed their in-kernel exchanges; non-zero = some launch timed out waiting for a peer CTA / rank
              std::vector<std::pair<void*, void*>> ws;
              {
                  std::lock_guard<std::mutex> lk(g_ws_mu);
                  for (auto& kv : g_ws_cache) ws.emplace_back(kv.second.buf.data_ptr(), std::get<1>(kv.first));
              }
              uint32_t worst = 0;
              for (auto& w : ws) {
                  uint32_t st = 0;
                  const int rc = cf_workspace_status(w.first, w.second, &st);
                  TORCH_CHECK(rc == 0, "clusterfusion_b200: cf_workspace_status failed (", rc, "): ", cf_last_error_string());
                  worst |= st;
              }
              return worst;
          },
          "OR of the sticky error words of all workspaces owned by this module (0 = ok); synchronises their streams");
    m.def("tensor_map_encodes", []() { return cf_debug_tensor_map_encodes(); },
          "cuTensorMapEncodeTiled calls made by the library so far (0 per token in a steady-state decode loop)");
    // called from an atexit hook of the Python package: CUDA tensors must not be destroyed by static destructors after the
    // CUDA context is gone
    m.def("_release_workspaces", []() {
        std::lock_guard<std::mutex> lk(g_ws_mu);
        g_ws_cache.clear();
        g_ds_ws_cache.clear();
    });
    m.def("abi_version", []() { return cf_abi_version(); });
    m.def("workspace_bytes", [](int hidden, int batch) { return cf_llama_workspace_bytes(hidden, batch); });
}
