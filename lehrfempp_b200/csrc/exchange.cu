// Interface-row exchange helpers of the multi-GPU path (product code).
//
// The reference is serial; the partitioned assembly is described in DESIGN.md ("Multi-GPU").  Each GPU assembles the
// contributions of ITS cells; matrix rows touched by cells of several GPUs ("interface rows") hold partial sums that
// are sent to the row's owner.  These kernels move whole row segments values[outer[r] .. outer[r+1]) between the value
// array and a contiguous message buffer; the transport itself (NCCL over NVLink) is driven from the host binding.
#include "lfgpu_internal.cuh"

namespace lfgpu {
namespace {

// one warp per listed row
template <bool UNPACK_ADD>
__global__ void k_rows_copy(int64_t n_rows, const int32_t* __restrict__ rows, const int64_t* __restrict__ offsets,
                            const int32_t* __restrict__ outer, double* __restrict__ values, double* __restrict__ buf) {
  const int64_t w = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= n_rows) return;
  const int32_t r = rows[w];
  const int32_t v0 = outer[r], len = outer[r + 1] - v0;
  const int64_t o = offsets[w];
  for (int k = lane; k < len; k += 32) {
    if (UNPACK_ADD) {
      // a row at a corner of the partition is touched by three or more GPUs and then appears once per sender in the list: the
      // additions of two warps to one entry must not be a plain read-modify-write (found by the 8-GPU parity run of round 2:
      // lost updates, errors of order one; two GPUs never list a row twice)
      atomicAdd(values + v0 + k, buf[o + k]);
    } else {
      buf[o + k] = values[v0 + k];
    }
  }
}

}  // namespace
}  // namespace lfgpu

using namespace lfgpu;

extern "C" {

int lfgpu_rows_pack(lfgpu_ctx* ctx, const lfgpu_pattern* p, const int32_t* d_rows, int64_t n_rows, const int64_t* d_offsets,
                    const double* d_values, double* d_buf) {
  if (ctx == nullptr || p == nullptr) return LFGPU_ERR_INVALID;
  if (n_rows <= 0) return LFGPU_OK;
  k_rows_copy<false><<<static_cast<unsigned>(cdiv(n_rows * 32, 256)), 256, 0, ctx->stream>>>(n_rows, d_rows, d_offsets, p->outer,
                                                                                            const_cast<double*>(d_values), d_buf);
  LFGPU_LAUNCH_CHECK(ctx);
  return LFGPU_OK;
}

int lfgpu_rows_unpack_add(lfgpu_ctx* ctx, const lfgpu_pattern* p, const int32_t* d_rows, int64_t n_rows, const int64_t* d_offsets,
                          const double* d_buf, double* d_values) {
  if (ctx == nullptr || p == nullptr) return LFGPU_ERR_INVALID;
  if (n_rows <= 0) return LFGPU_OK;
  k_rows_copy<true><<<static_cast<unsigned>(cdiv(n_rows * 32, 256)), 256, 0, ctx->stream>>>(n_rows, d_rows, d_offsets, p->outer, d_values,
                                                                                           const_cast<double*>(d_buf));
  LFGPU_LAUNCH_CHECK(ctx);
  return LFGPU_OK;
}

const int32_t* lfgpu_pattern_adj_ptr_device(const lfgpu_pattern* p) { return p ? p->adj_ptr : nullptr; }
const uint32_t* lfgpu_pattern_adj_device(const lfgpu_pattern* p) { return p ? p->adj : nullptr; }
int64_t lfgpu_pattern_num_items(const lfgpu_pattern* p) { return p ? p->n_items : -1; }
const double* lfgpu_mesh_node_coords_device(const lfgpu_mesh* m) { return m ? m->node_coords : nullptr; }
const uint32_t* lfgpu_mesh_cell_nodes_device(const lfgpu_mesh* m) { return m ? m->cell_nodes : nullptr; }

// make the ctx stream wait for / signal an external CUDA event (overlap of the exchange with the interior rows)
int lfgpu_ctx_wait_event(lfgpu_ctx* ctx, void* cuda_event) {
  if (ctx == nullptr || cuda_event == nullptr) return LFGPU_ERR_INVALID;
  LFGPU_CUDA_CHECK(ctx, cudaStreamWaitEvent(ctx->stream, static_cast<cudaEvent_t>(cuda_event), 0));
  return LFGPU_OK;
}

}  // extern "C"
