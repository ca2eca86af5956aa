// Context, error reporting and device-memory helpers of liblfgpu.so (product code).
#include <mutex>

#include "lfgpu_internal.cuh"

namespace lfgpu {
namespace {
thread_local std::string g_thread_error;
}
void set_last_error(const lfgpu_ctx* ctx, const std::string& msg) {
  g_thread_error = msg;
  if (ctx != nullptr) const_cast<lfgpu_ctx*>(ctx)->last_error = msg;
}
}  // namespace lfgpu

extern "C" {

const char* lfgpu_version(void) { return "lfgpu 0.1 (sm_100a)"; }

int lfgpu_ctx_create(int device, lfgpu_ctx** out) {
  if (out == nullptr) return LFGPU_ERR_INVALID;
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    // no CPU fallback: fail loudly
    lfgpu::set_last_error(nullptr, std::string("lfgpu needs a CUDA device (sm_100a); none visible: ") +
                                       (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
    return LFGPU_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= n) {
    lfgpu::set_last_error(nullptr, "device index out of range");
    return LFGPU_ERR_INVALID;
  }
  auto* ctx = new lfgpu_ctx;
  ctx->device = device;
  if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) {
    lfgpu::set_last_error(nullptr, std::string("cuda init: ") + cudaGetErrorString(e));
    delete ctx;
    return LFGPU_ERR_CUDA;
  }
  cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
  if ((e = cudaMalloc(&ctx->d_scratch, 4096)) != cudaSuccess) {
    lfgpu::set_last_error(nullptr, std::string("cudaMalloc scratch: ") + cudaGetErrorString(e));
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return LFGPU_ERR_CUDA;
  }
  *out = ctx;
  return LFGPU_OK;
}

void lfgpu_ctx_destroy(lfgpu_ctx* ctx) {
  if (ctx == nullptr) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (auto& e : ctx->table_cache) cudaFree(e.dev);
  for (cudaEvent_t ev : ctx->pipe_events) cudaEventDestroy(ev);
  if (ctx->s_h2d != nullptr) cudaStreamDestroy(ctx->s_h2d);
  if (ctx->s_d2h != nullptr) cudaStreamDestroy(ctx->s_d2h);
  cudaFree(ctx->nodal_tab[0]);
  cudaFree(ctx->nodal_tab[1]);
  cudaFree(ctx->d_scratch);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* lfgpu_last_error(const lfgpu_ctx* ctx) {
  if (ctx != nullptr) return ctx->last_error.c_str();
  return lfgpu::g_thread_error.c_str();
}

int lfgpu_ctx_synchronize(lfgpu_ctx* ctx) {
  if (ctx == nullptr) return LFGPU_ERR_INVALID;
  LFGPU_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  if (ctx->geom_check_pending) {  // degeneracy check queued by lfgpu_mesh_update_node_coords
    ctx->geom_check_pending = false;
    int h[2] = {0, 0};
    LFGPU_CUDA_CHECK(ctx, cudaMemcpy(h, static_cast<char*>(ctx->d_scratch) + 1024, sizeof(h), cudaMemcpyDeviceToHost));
    if (h[1]) LFGPU_FAIL(ctx, LFGPU_ERR_DEGENERATE, "degenerate cell geometry after a coordinate update (collapsed edge or zero area)");
  }
  return LFGPU_OK;
}

void* lfgpu_ctx_stream(lfgpu_ctx* ctx) { return ctx ? static_cast<void*>(ctx->stream) : nullptr; }
int64_t lfgpu_ctx_kernel_launches(const lfgpu_ctx* ctx) { return ctx ? ctx->launches : 0; }

// ---- timing on the ctx stream (CUDA events) ------------------------------------------------------------------------------
int lfgpu_event_create(lfgpu_ctx* ctx, void** ev) {
  if (ctx == nullptr || ev == nullptr) return LFGPU_ERR_INVALID;
  cudaEvent_t e;
  LFGPU_CUDA_CHECK(ctx, cudaEventCreate(&e));
  *ev = e;
  return LFGPU_OK;
}
int lfgpu_event_record(lfgpu_ctx* ctx, void* ev) {
  if (ctx == nullptr || ev == nullptr) return LFGPU_ERR_INVALID;
  LFGPU_CUDA_CHECK(ctx, cudaEventRecord(static_cast<cudaEvent_t>(ev), ctx->stream));
  return LFGPU_OK;
}
int lfgpu_event_elapsed_ms(lfgpu_ctx* ctx, void* start, void* stop, double* ms) {
  if (ctx == nullptr || ms == nullptr) return LFGPU_ERR_INVALID;
  LFGPU_CUDA_CHECK(ctx, cudaEventSynchronize(static_cast<cudaEvent_t>(stop)));
  float f = 0.f;
  LFGPU_CUDA_CHECK(ctx, cudaEventElapsedTime(&f, static_cast<cudaEvent_t>(start), static_cast<cudaEvent_t>(stop)));
  *ms = f;
  return LFGPU_OK;
}
int lfgpu_event_destroy(lfgpu_ctx* ctx, void* ev) {
  if (ctx == nullptr) return LFGPU_ERR_INVALID;
  LFGPU_CUDA_CHECK(ctx, cudaEventDestroy(static_cast<cudaEvent_t>(ev)));
  return LFGPU_OK;
}

int lfgpu_malloc(lfgpu_ctx* ctx, int64_t bytes, void** d_ptr) {
  if (ctx == nullptr || d_ptr == nullptr || bytes < 0) return LFGPU_ERR_INVALID;
  LFGPU_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  LFGPU_CUDA_CHECK(ctx, cudaMalloc(d_ptr, bytes > 0 ? bytes : 1));
  return LFGPU_OK;
}
int lfgpu_free(lfgpu_ctx* ctx, void* d_ptr) {
  if (ctx == nullptr) return LFGPU_ERR_INVALID;
  LFGPU_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  LFGPU_CUDA_CHECK(ctx, cudaFree(d_ptr));
  return LFGPU_OK;
}
int lfgpu_memset(lfgpu_ctx* ctx, void* d_ptr, int value, int64_t bytes) {
  if (ctx == nullptr) return LFGPU_ERR_INVALID;
  LFGPU_CUDA_CHECK(ctx, cudaMemsetAsync(d_ptr, value, bytes, ctx->stream));
  return LFGPU_OK;
}
int lfgpu_memcpy_h2d(lfgpu_ctx* ctx, void* d_dst, const void* h_src, int64_t bytes) {
  if (ctx == nullptr) return LFGPU_ERR_INVALID;
  LFGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return LFGPU_OK;
}
int lfgpu_memcpy_d2h(lfgpu_ctx* ctx, void* h_dst, const void* d_src, int64_t bytes) {
  if (ctx == nullptr) return LFGPU_ERR_INVALID;
  LFGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  return LFGPU_OK;
}
int lfgpu_host_alloc_pinned(lfgpu_ctx* ctx, int64_t bytes, void** h_ptr) {
  if (ctx == nullptr || h_ptr == nullptr) return LFGPU_ERR_INVALID;
  LFGPU_CUDA_CHECK(ctx, cudaHostAlloc(h_ptr, bytes > 0 ? bytes : 1, cudaHostAllocDefault));
  return LFGPU_OK;
}
int lfgpu_host_free_pinned(lfgpu_ctx* ctx, void* h_ptr) {
  if (ctx == nullptr) return LFGPU_ERR_INVALID;
  LFGPU_CUDA_CHECK(ctx, cudaFreeHost(h_ptr));
  return LFGPU_OK;
}

}  // extern "C"
