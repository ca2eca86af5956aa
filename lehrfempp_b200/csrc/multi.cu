// Multi-GPU assembly through the C ABI (product code): one process, one lfgpu_ctx per listed device, the problem cut into
// per-device sub-problems by the distributed-ownership layer of partition.cu.  What a C++ caller of
// lf::assemble::AssembleMatrixLocally (assemble/assembler.h:114-186) gets when it hands over a device LIST instead of one
// device (SURVEY.md section 8b sketched lfgpu_ctx_create(device_ids, n_dev, ...)): Morton cell ranges, every device owns the
// matrix rows of its cells, owner-computes with a one-cell halo, so the numeric pass needs no exchange between the devices and
// no NCCL; the rows of the result stay where they were computed.  Setup runs one host thread per device (each with its own
// context); the numeric pass only queues kernels, so one thread drives all devices.
#include <algorithm>
#include <cstring>
#include <thread>

#include "lfgpu_internal.cuh"

struct lfgpu_multi {
  struct Part {
    lfgpu_ctx* ctx = nullptr;
    lfgpu_submesh* sub = nullptr;
    lfgpu_pattern* pattern = nullptr;
    double* d_values = nullptr;
    bool empty = true;
    int64_t n_cells = 0, n_nodes = 0, n_dofs = 0, n_owned_rows = 0, owned_nnz = 0;
    std::vector<int32_t> l2g_cells, l2g_dofs, l2g_nodes, outer, inner;
    std::vector<uint8_t> owned;
    double* d_alpha = nullptr;  // per-cell / per-point coefficient tables of this device's cells
    double* d_gamma = nullptr;
    int64_t alpha_len = 0, gamma_len = 0;
    int rc = LFGPU_OK;
    std::string err;
  };
  std::vector<Part> parts;
  int major = LFGPU_COL_MAJOR;
  int64_t n_dofs = 0, n_cells = 0;
  bool ready = false;
  std::string last_error;
};

namespace lfgpu {
namespace {

void free_part(lfgpu_multi::Part& p) {
  if (p.ctx == nullptr) return;
  cudaSetDevice(p.ctx->device);
  cudaStreamSynchronize(p.ctx->stream);
  cudaFree(p.d_values);
  cudaFree(p.d_alpha);
  cudaFree(p.d_gamma);
  lfgpu_pattern_destroy(p.pattern);
  lfgpu_submesh_destroy(p.sub);
  p.d_values = p.d_alpha = p.d_gamma = nullptr;
  p.pattern = nullptr;
  p.sub = nullptr;
}

#define PART_CHECK(expr)                                         \
  do {                                                           \
    const int _rc = (expr);                                      \
    if (_rc != LFGPU_OK) {                                       \
      part.rc = _rc;                                             \
      part.err = std::string(#expr) + ": " + lfgpu_last_error(part.ctx); \
      return;                                                    \
    }                                                            \
  } while (0)
#define PART_CUDA(expr)                                          \
  do {                                                           \
    const cudaError_t _e = (expr);                               \
    if (_e != cudaSuccess) {                                     \
      part.rc = LFGPU_ERR_CUDA;                                  \
      part.err = std::string(#expr) + ": " + cudaGetErrorString(_e); \
      return;                                                    \
    }                                                            \
  } while (0)

// everything device k does at setup: upload the flattened problem, cut out its share, drop the rest, symbolic pass
void setup_part(lfgpu_multi::Part& part, int k, int n_parts, int64_t n_nodes, const double* node_coords, int64_t n_cells,
                const uint32_t* cell_nodes, const double* cell_coords, int64_t n_dofs, int stride, const int64_t* cell_dofs,
                const uint8_t* n_ldof, int major) {
  lfgpu_ctx* ctx = part.ctx;
  lfgpu_mesh* mesh = nullptr;
  lfgpu_dofmap* dm = nullptr;
  uint8_t *d_part = nullptr, *d_owner = nullptr, *d_sel = nullptr, *d_owned = nullptr;
  auto cleanup = [&]() {
    cudaFree(d_part); cudaFree(d_owner); cudaFree(d_sel); cudaFree(d_owned);
    lfgpu_dofmap_destroy(dm);
    lfgpu_mesh_destroy(mesh);
  };
  struct Guard {
    decltype(cleanup)& f;
    ~Guard() { f(); }
  } guard{cleanup};
  PART_CHECK(lfgpu_mesh_upload(ctx, n_nodes, node_coords, n_cells, cell_nodes, cell_coords, &mesh));
  PART_CHECK(lfgpu_dofmap_upload(ctx, mesh, n_dofs, stride, cell_dofs, n_ldof, &dm));
  PART_CUDA(cudaMalloc(&d_part, n_cells));
  PART_CUDA(cudaMalloc(&d_owner, n_dofs));
  PART_CUDA(cudaMalloc(&d_sel, n_cells));
  PART_CHECK(lfgpu_partition_morton(ctx, mesh, n_parts, d_part));
  PART_CHECK(lfgpu_partition_dof_owner(ctx, dm, d_part, d_owner));
  PART_CHECK(lfgpu_partition_select_cells(ctx, dm, d_part, d_owner, k, 1, d_sel));
  PART_CHECK(lfgpu_submesh_extract(ctx, mesh, dm, d_sel, &part.sub));
  PART_CHECK(lfgpu_submesh_counts(part.sub, &part.n_cells, &part.n_nodes, &part.n_dofs));
  PART_CUDA(cudaMalloc(&d_owned, part.n_dofs));
  PART_CHECK(lfgpu_submesh_owned_dofs(ctx, part.sub, d_owner, k, d_owned));
  part.l2g_cells.resize(part.n_cells);
  part.l2g_dofs.resize(part.n_dofs);
  part.l2g_nodes.resize(part.n_nodes);
  part.owned.resize(part.n_dofs);
  PART_CUDA(cudaMemcpyAsync(part.l2g_nodes.data(), lfgpu_submesh_l2g_nodes_device(part.sub), 4 * part.n_nodes, cudaMemcpyDeviceToHost, ctx->stream));
  PART_CUDA(cudaMemcpyAsync(part.l2g_cells.data(), lfgpu_submesh_l2g_cells_device(part.sub), 4 * part.n_cells, cudaMemcpyDeviceToHost, ctx->stream));
  PART_CUDA(cudaMemcpyAsync(part.l2g_dofs.data(), lfgpu_submesh_l2g_dofs_device(part.sub), 4 * part.n_dofs, cudaMemcpyDeviceToHost, ctx->stream));
  PART_CUDA(cudaMemcpyAsync(part.owned.data(), d_owned, part.n_dofs, cudaMemcpyDeviceToHost, ctx->stream));
  PART_CUDA(cudaStreamSynchronize(ctx->stream));
  // the global copies go before the pattern is built: the sub-problem is all this device keeps
  cleanup();
  mesh = nullptr;
  dm = nullptr;
  d_part = d_owner = d_sel = d_owned = nullptr;
  lfgpu_dofmap* sd = lfgpu_submesh_dofmap(part.sub);
  PART_CHECK(lfgpu_symbolic(ctx, lfgpu_submesh_mesh(part.sub), sd, sd, major, &part.pattern));
  {
    uint8_t* d_keep = nullptr;
    PART_CUDA(cudaMalloc(&d_keep, part.n_dofs));
    cudaError_t e = cudaMemcpy(d_keep, part.owned.data(), part.n_dofs, cudaMemcpyHostToDevice);
    const int rk = e == cudaSuccess ? lfgpu_pattern_restrict_rows(ctx, part.pattern, d_keep) : LFGPU_ERR_CUDA;
    cudaStreamSynchronize(ctx->stream);
    cudaFree(d_keep);
    PART_CHECK(rk);
  }
  const int64_t nnz = lfgpu_pattern_nnz(part.pattern);
  PART_CUDA(cudaMalloc(&part.d_values, sizeof(double) * (nnz > 0 ? nnz : 1)));
  PART_CUDA(cudaMemsetAsync(part.d_values, 0, sizeof(double) * nnz, ctx->stream));
  part.outer.resize(part.n_dofs + 1);
  part.inner.resize(nnz);
  PART_CHECK(lfgpu_pattern_download(ctx, part.pattern, part.outer.data(), part.inner.data()));
  part.n_owned_rows = 0;
  part.owned_nnz = 0;
  for (int64_t r = 0; r < part.n_dofs; ++r) {
    if (part.owned[r]) {
      ++part.n_owned_rows;
      part.owned_nnz += part.outer[r + 1] - part.outer[r];
    }
  }
  part.empty = true;
}

// coefficient of the global problem -> the same coefficient on one device's cells (tables gathered through l2g_cells)
int local_coeff(lfgpu_multi::Part& part, const lfgpu_coeff* c, double** d_buf, int64_t* buf_len, lfgpu_coeff* out) {
  *out = *c;
  if (c->kind == LFGPU_COEFF_CONST || c->kind == LFGPU_COEFF_CONST_2X2) return LFGPU_OK;
  if (c->data == nullptr) return LFGPU_ERR_INVALID;
  const bool nodal = c->kind == LFGPU_COEFF_NODAL;  // one value per mesh node instead of per cell
  const int64_t per_cell = (nodal || c->kind == LFGPU_COEFF_PER_CELL) ? 1 : (c->kind == LFGPU_COEFF_PER_QP ? c->stride : 4 * c->stride);
  const int64_t n_ent = nodal ? part.n_nodes : part.n_cells;
  const std::vector<int32_t>& l2g = nodal ? part.l2g_nodes : part.l2g_cells;
  const int64_t n = n_ent * per_cell;
  std::vector<double> h(static_cast<size_t>(n));
  for (int64_t i = 0; i < n_ent; ++i)
    std::memcpy(h.data() + i * per_cell, c->data + static_cast<int64_t>(l2g[i]) * per_cell, sizeof(double) * per_cell);
  cudaSetDevice(part.ctx->device);
  if (*buf_len < n) {
    cudaStreamSynchronize(part.ctx->stream);
    cudaFree(*d_buf);
    *d_buf = nullptr;
    if (cudaMalloc(d_buf, sizeof(double) * n) != cudaSuccess) return LFGPU_ERR_CUDA;
    *buf_len = n;
  }
  if (cudaMemcpy(*d_buf, h.data(), sizeof(double) * n, cudaMemcpyHostToDevice) != cudaSuccess) return LFGPU_ERR_CUDA;
  out->data = *d_buf;
  return LFGPU_OK;
}

}  // namespace
}  // namespace lfgpu

using namespace lfgpu;

extern "C" {

int lfgpu_multi_create(const int* device_ids, int n_dev, lfgpu_multi** out) {
  if (out == nullptr || device_ids == nullptr || n_dev < 1 || n_dev > 255) return LFGPU_ERR_INVALID;
  *out = nullptr;
  auto* m = new lfgpu_multi;
  m->parts.resize(n_dev);
  for (int k = 0; k < n_dev; ++k) {
    const int rc = lfgpu_ctx_create(device_ids[k], &m->parts[k].ctx);
    if (rc != LFGPU_OK) {
      lfgpu_multi_destroy(m);
      return rc;
    }
  }
  *out = m;
  return LFGPU_OK;
}

void lfgpu_multi_destroy(lfgpu_multi* m) {
  if (m == nullptr) return;
  for (auto& p : m->parts) {
    free_part(p);
    lfgpu_ctx_destroy(p.ctx);
  }
  delete m;
}

int lfgpu_multi_num_devices(const lfgpu_multi* m) { return m ? static_cast<int>(m->parts.size()) : 0; }
lfgpu_ctx* lfgpu_multi_ctx(lfgpu_multi* m, int k) { return (m && k >= 0 && k < static_cast<int>(m->parts.size())) ? m->parts[k].ctx : nullptr; }
const char* lfgpu_multi_last_error(const lfgpu_multi* m) { return m ? m->last_error.c_str() : ""; }

int lfgpu_multi_setup(lfgpu_multi* m, int64_t n_nodes, const double* node_coords, int64_t n_cells, const uint32_t* cell_nodes,
                      const double* cell_coords, int64_t n_dofs, int stride, const int64_t* cell_dofs, const uint8_t* n_ldof, int major) {
  if (m == nullptr || node_coords == nullptr || cell_nodes == nullptr || cell_dofs == nullptr) return LFGPU_ERR_INVALID;
  if (major != LFGPU_COL_MAJOR && major != LFGPU_ROW_MAJOR) return LFGPU_ERR_INVALID;
  for (auto& p : m->parts) free_part(p);
  m->ready = false;
  m->major = major;
  m->n_dofs = n_dofs;
  m->n_cells = n_cells;
  const int n_parts = static_cast<int>(m->parts.size());
  // one thread per distinct DEVICE; parts that share a device (a list may name one several times) are set up one after the other
  std::vector<int> devices;
  for (int k = 0; k < n_parts; ++k) {
    m->parts[k].rc = LFGPU_OK;
    const int dev = m->parts[k].ctx->device;
    if (std::find(devices.begin(), devices.end(), dev) == devices.end()) devices.push_back(dev);
  }
  std::vector<std::thread> th;
  for (const int dev : devices) {
    th.emplace_back([=]() {
      for (int k = 0; k < n_parts; ++k) {
        if (m->parts[k].ctx->device != dev) continue;
        setup_part(m->parts[k], k, n_parts, n_nodes, node_coords, n_cells, cell_nodes, cell_coords, n_dofs, stride, cell_dofs, n_ldof, major);
      }
    });
  }
  for (auto& t : th) t.join();
  for (int k = 0; k < n_parts; ++k) {
    if (m->parts[k].rc != LFGPU_OK) {
      m->last_error = "device " + std::to_string(m->parts[k].ctx->device) + ": " + m->parts[k].err;
      return m->parts[k].rc;
    }
  }
  m->ready = true;
  return LFGPU_OK;
}

int lfgpu_multi_set_zero(lfgpu_multi* m) {
  if (m == nullptr || !m->ready) return LFGPU_ERR_INVALID;
  for (auto& p : m->parts) {
    cudaSetDevice(p.ctx->device);
    if (cudaMemsetAsync(p.d_values, 0, sizeof(double) * lfgpu_pattern_nnz(p.pattern), p.ctx->stream) != cudaSuccess) return LFGPU_ERR_CUDA;
    p.empty = true;
  }
  return LFGPU_OK;
}

int lfgpu_multi_assemble_reaction_diffusion(lfgpu_multi* m, int degree, const lfgpu_quad* qr_tria, const lfgpu_quad* qr_quad,
                                            const lfgpu_coeff* alpha, const lfgpu_coeff* gamma, int accumulate) {
  if (m == nullptr || !m->ready || alpha == nullptr || gamma == nullptr) return LFGPU_ERR_INVALID;
  // queue the pass on every device, then wait for all of them
  for (auto& p : m->parts) {
    lfgpu_coeff la, lg;
    int rc = local_coeff(p, alpha, &p.d_alpha, &p.alpha_len, &la);
    if (rc == LFGPU_OK) rc = local_coeff(p, gamma, &p.d_gamma, &p.gamma_len, &lg);
    if (rc == LFGPU_OK)
      rc = lfgpu_assemble_reaction_diffusion(p.ctx, lfgpu_submesh_mesh(p.sub), p.pattern, degree, qr_tria, qr_quad, &la, &lg, nullptr,
                                             (accumulate && !p.empty) ? 1.0 : 0.0, p.d_values, LFGPU_ALGO_AUTO);
    if (rc != LFGPU_OK) {
      m->last_error = "device " + std::to_string(p.ctx->device) + ": " + lfgpu_last_error(p.ctx);
      return rc;
    }
    p.empty = false;
  }
  for (auto& p : m->parts) {
    const int rc = lfgpu_ctx_synchronize(p.ctx);
    if (rc != LFGPU_OK) {
      m->last_error = "device " + std::to_string(p.ctx->device) + ": " + lfgpu_last_error(p.ctx);
      return rc;
    }
  }
  return LFGPU_OK;
}

int lfgpu_multi_part_sizes(const lfgpu_multi* m, int k, int64_t* n_rows, int64_t* nnz, int64_t* n_local_cells, int64_t* n_local_rows,
                           int64_t* n_local_nnz) {
  if (m == nullptr || !m->ready || k < 0 || k >= static_cast<int>(m->parts.size())) return LFGPU_ERR_INVALID;
  const auto& p = m->parts[k];
  if (n_rows) *n_rows = p.n_owned_rows;
  if (nnz) *nnz = p.owned_nnz;
  if (n_local_cells) *n_local_cells = p.n_cells;
  if (n_local_rows) *n_local_rows = p.n_dofs;
  if (n_local_nnz) *n_local_nnz = lfgpu_pattern_nnz(p.pattern);
  return LFGPU_OK;
}

int lfgpu_multi_part_download(lfgpu_multi* m, int k, int64_t* rows, int64_t* row_ptr, int32_t* cols, double* values) {
  if (m == nullptr || !m->ready || k < 0 || k >= static_cast<int>(m->parts.size())) return LFGPU_ERR_INVALID;
  auto& p = m->parts[k];
  std::vector<double> h;
  if (values != nullptr) {
    h.resize(p.inner.size());
    cudaSetDevice(p.ctx->device);
    if (cudaMemcpyAsync(h.data(), p.d_values, sizeof(double) * h.size(), cudaMemcpyDeviceToHost, p.ctx->stream) != cudaSuccess ||
        cudaStreamSynchronize(p.ctx->stream) != cudaSuccess) {
      m->last_error = "download of the values failed";
      return LFGPU_ERR_CUDA;
    }
  }
  int64_t nr = 0, pos = 0;
  for (int64_t r = 0; r < p.n_dofs; ++r) {
    if (!p.owned[r]) continue;
    if (rows) rows[nr] = p.l2g_dofs[r];
    if (row_ptr) row_ptr[nr] = pos;
    for (int32_t t = p.outer[r]; t < p.outer[r + 1]; ++t, ++pos) {
      if (cols) cols[pos] = p.l2g_dofs[p.inner[t]];
      if (values) values[pos] = h[t];
    }
    ++nr;
  }
  if (row_ptr) row_ptr[nr] = pos;
  return LFGPU_OK;
}

}  // extern "C"
