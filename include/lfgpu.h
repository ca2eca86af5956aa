/* lfgpu.h -- C ABI of liblfgpu.so: B200-native (sm_100a) finite-element assembly behind LehrFEM++'s assembler API.
 *
 * The reference (craffael/lehrfempp) has no FFI; its seam for this path is the C++ call
 *     lf::assemble::AssembleMatrixLocally(codim, dofh_trial, dofh_test, provider, matrix)   lib/lf/assemble/assembler.h:114-186
 *     lf::assemble::AssembleVectorLocally(codim, dofh, provider, vector)                    lib/lf/assemble/assembler.h:298-327
 * with the lf::uscalfe providers (lib/lf/uscalfe/loc_comp_ellbvp.h:85-339, 562-746).  Every entry point below names the
 * reference interface it stands in for.  The header-only C++ shim include/lf_gpu_shim.hpp keeps the reference call
 * signatures and forwards to these functions; INTEGRATION.md shows the binding a LehrFEM++ maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; the caller owns every host array; handles own device memory
 *   - every function returns 0 on success and a negative lfgpu_status otherwise; nothing throws or aborts
 *     (the shim turns LFGPU_ERR_MISSING_RULE into lf::base::LfException, mirroring loc_comp_ellbvp.h:278-287)
 *   - one lfgpu_ctx per host thread and GPU; calls on one ctx are issued on its single CUDA stream, in order
 *   - there is NO CPU fallback: without a CUDA device lfgpu_ctx_create fails with LFGPU_ERR_NO_DEVICE
 *   - index types follow the reference: entity indices uint32 (lib/lf/base/types.h:20-36, 0xFFFFFFFF = "no 4th vertex"),
 *     global dof indices int64 on the host side (lib/lf/assemble/assembly_types.h:22), int32 in the compressed
 *     pattern like Eigen::SparseMatrix<double> (StorageIndex = int, lib/lf/assemble/coomatrix.h:172-180)
 */
#ifndef LFGPU_H
#define LFGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lfgpu_ctx lfgpu_ctx;
typedef struct lfgpu_mesh lfgpu_mesh;       /* flattened lf::mesh::Mesh (cells, nodes, optional edges) on the device      */
typedef struct lfgpu_dofmap lfgpu_dofmap;   /* flattened lf::assemble::DofHandler: cell -> global dof table on the device  */
typedef struct lfgpu_pattern lfgpu_pattern; /* compressed sparsity pattern + scatter/gather plan (symbolic pass result)    */

typedef enum {
  LFGPU_OK = 0,
  LFGPU_ERR_INVALID = -1,      /* bad argument                                                        */
  LFGPU_ERR_CUDA = -2,         /* CUDA runtime error, text in lfgpu_last_error                        */
  LFGPU_ERR_NO_DEVICE = -3,    /* no CUDA device: there is no CPU fallback                            */
  LFGPU_ERR_MISSING_RULE = -4, /* no quadrature rule / shape functions for a cell type that occurs    */
  LFGPU_ERR_DEGENERATE = -5,   /* degenerate cell geometry (tria_o1.cc:10-48, quad_o1.cc:14-59)       */
  LFGPU_ERR_OVERFLOW = -6,     /* nnz or an index does not fit the reference's int32 storage index    */
  LFGPU_ERR_UNSUPPORTED = -7,
  LFGPU_ERR_NCCL = -8
} lfgpu_status;

#define LFGPU_IDX_NIL 0xFFFFFFFFu

/* storage order of the compressed matrix */
#define LFGPU_COL_MAJOR 0 /* Eigen::SparseMatrix default = what COOMatrix::makeSparse returns (coomatrix.h:172-180) */
#define LFGPU_ROW_MAJOR 1 /* CSR */

/* numeric-pass algorithm */
#define LFGPU_ALGO_AUTO 0
#define LFGPU_ALGO_ATOMIC 1 /* one thread per (cell, local row), FP64 atomics into the values (supports beta = 1)       */
#define LFGPU_ALGO_GATHER 2 /* owner-computes: one thread per matrix row gathers its cells, deterministic, no atomics    */
#define LFGPU_ALGO_FAN 3    /* triangles with constant coefficients: the kernels that own matrix rows in registers -- P1 vertex-fan
                               kernel, P2 / P3 row kernels (AUTO picks them when they apply; this value insists on them)       */

/* ---- context ------------------------------------------------------------------------------------------------------ */
int lfgpu_ctx_create(int device, lfgpu_ctx** out);
void lfgpu_ctx_destroy(lfgpu_ctx* ctx);
const char* lfgpu_last_error(const lfgpu_ctx* ctx); /* ctx may be NULL: last error of the calling thread */
int lfgpu_ctx_synchronize(lfgpu_ctx* ctx);
void* lfgpu_ctx_stream(lfgpu_ctx* ctx);       /* the cudaStream_t all work of this ctx is issued on */
int64_t lfgpu_ctx_kernel_launches(const lfgpu_ctx* ctx); /* number of lfgpu kernels launched so far on this ctx */
const char* lfgpu_version(void);
/* CUDA events recorded on the ctx stream: device-side timing of a sequence of calls */
int lfgpu_event_create(lfgpu_ctx* ctx, void** ev);
int lfgpu_event_record(lfgpu_ctx* ctx, void* ev);
int lfgpu_event_elapsed_ms(lfgpu_ctx* ctx, void* start, void* stop, double* ms); /* waits for `stop` */
int lfgpu_event_destroy(lfgpu_ctx* ctx, void* ev);

/* ---- device memory (plain helpers so that hosts without a CUDA binding can hold results) -------------------------- */
int lfgpu_malloc(lfgpu_ctx* ctx, int64_t bytes, void** d_ptr);
int lfgpu_free(lfgpu_ctx* ctx, void* d_ptr);
int lfgpu_memset(lfgpu_ctx* ctx, void* d_ptr, int value, int64_t bytes);
int lfgpu_memcpy_h2d(lfgpu_ctx* ctx, void* d_dst, const void* h_src, int64_t bytes); /* async on the ctx stream */
int lfgpu_memcpy_d2h(lfgpu_ctx* ctx, void* h_dst, const void* d_src, int64_t bytes); /* async on the ctx stream */
int lfgpu_host_alloc_pinned(lfgpu_ctx* ctx, int64_t bytes, void** h_ptr);
int lfgpu_host_free_pinned(lfgpu_ctx* ctx, void* h_ptr);

/* ---- mesh: stands in for lf::mesh::Mesh as the assembler sees it -------------------------------------------------- */
/* Upload a flattened mesh.  cell_nodes is [n_cells][4] with LFGPU_IDX_NIL in slot 3 of a triangle (hybrid2d/
 * mesh_factory.cc:97-100); cells are in lf::mesh::Mesh::Entities(0) order.  node_coords is [n_nodes][2].
 * cell_coords ([n_cells][4][2], nullable) are the corners taken from cell.Geometry()->Global(RefEl.NodeCoords());
 * pass them when they are not bitwise equal to the node positions (e.g. refined meshes, tria_o1.cc:99-151).      */
int lfgpu_mesh_upload(lfgpu_ctx* ctx, int64_t n_nodes, const double* node_coords, int64_t n_cells,
                      const uint32_t* cell_nodes, const double* cell_coords, lfgpu_mesh** out);
/* Device-side mesh generators with the reference's numbering.
 * tp_tria: lf::mesh::utils::TPTriagMeshBuilder (mesh/utils/tp_triag_mesh_builder.cc:18-178)
 * tp_quad: lf::mesh::utils::TPQuadMeshBuilder  (mesh/utils/tp_quad_mesh_builder.cc:19-95)
 * hybrid : synthetic jittered tri/quad mesh of benchmark config C2 (spec in DESIGN.md)                               */
int lfgpu_mesh_tp_tria(lfgpu_ctx* ctx, uint32_t nx, uint32_t ny, double x0, double y0, double x1, double y1, lfgpu_mesh** out);
int lfgpu_mesh_tp_quad(lfgpu_ctx* ctx, uint32_t nx, uint32_t ny, double x0, double y0, double x1, double y1, lfgpu_mesh** out);
int lfgpu_mesh_hybrid(lfgpu_ctx* ctx, uint32_t n, double jitter, uint64_t seed, lfgpu_mesh** out);
/* refine : lf::refinement::MeshHierarchy::RefineRegular (refinement/mesh_hierarchy.cc:72-114, 368-1262) -- one step of
 *          regular refinement of `parent` with the reference's node / edge / cell numbering (SURVEY.md section 8f row
 *          3): the mesh of benchmark config C4.  Needs a parent whose cell corners are its node positions.            */
int lfgpu_mesh_refine_regular(lfgpu_ctx* ctx, lfgpu_mesh* parent, lfgpu_mesh** out);
/* Edge numbering and orientations exactly as the lf::mesh::hybrid2d::Mesh constructor assigns them
 * (mesh/hybrid2d/mesh.cc:178-810): the n_explicit edges (edge_nodes [n][2], nullable) keep their position as index and
 * their direction; the remaining edges are numbered in ascending (min,max) endpoint order and point along the local
 * direction of the lowest-index adjacent cell.  Needed by lfgpu_dofmap_uniform when edges carry dofs.
 * cell_has_geometry (host uint8 [n_cells], nullable = same policy for all cells) mirrors MeshFactory::AddEntity's
 * optional geometry argument: an edge first met in a cell WITHOUT geometry is reversed when the first later cell WITH
 * geometry runs along it the other way (mesh.cc:413-428, 532-534).
 * The tp_tria generator supplies its explicit edge list itself.                                                       */
int lfgpu_mesh_build_topology(lfgpu_ctx* ctx, lfgpu_mesh* mesh, int64_t n_explicit, const uint32_t* edge_nodes,
                              const uint8_t* cell_has_geometry);
int lfgpu_mesh_counts(const lfgpu_mesh* mesh, int64_t* n_nodes, int64_t* n_edges, int64_t* n_cells, int64_t* n_tria, int64_t* n_quad);
/* every output nullable; layouts as in lfgpu_mesh_upload, cell_edges [n_cells][4], cell_edge_ori int8 [n_cells][4] (+1/-1) */
int lfgpu_mesh_download(lfgpu_ctx* ctx, const lfgpu_mesh* mesh, uint8_t* cell_type, uint32_t* cell_nodes, double* cell_coords,
                        uint32_t* cell_edges, int8_t* cell_edge_ori, uint32_t* edge_nodes, double* node_coords);
/* replace the node positions (per-step input of a moving-mesh / re-assembly loop); host array [n_nodes][2].  The copy is
 * asynchronous on the ctx stream: a page-locked host buffer must stay valid until lfgpu_ctx_synchronize.  The new geometry is
 * checked like that of lfgpu_mesh_upload; a degenerate cell makes the NEXT lfgpu_ctx_synchronize return LFGPU_ERR_DEGENERATE.
 * (The host-buffer assembly calls, which take the positions of the step as an argument, check them before they return.)          */
int lfgpu_mesh_update_node_coords(lfgpu_ctx* ctx, lfgpu_mesh* mesh, const double* node_coords);
void lfgpu_mesh_destroy(lfgpu_mesh* mesh);

/* ---- Gmsh input: stands in for lf::io::GmshReader (io/gmsh_reader.h:55-200, io/gmsh_reader.cc) ----------------------- */
/* Host-side reader of MSH 2.2 and 4.1 files, text or binary (ReadGmshFile, gmsh_reader.cc:629-697), flattened the way
 * GmshReader::InitGmshFile (gmsh_reader.cc:121-340, 343-627) feeds its hybrid2d::MeshFactory -- which fixes the entity
 * numbering every dof number depends on: mesh nodes = the vertices of elements in FILE order (auxiliary nodes of
 * second-order elements are not mesh nodes), explicitly listed edges in file order ahead of all other edges, cells in
 * file order, consecutive repetitions of one element merged.  dim_world must be 2.  Errors (unreadable file,
 * unsupported version or element type, z != 0) return LFGPU_ERR_INVALID with the text in lfgpu_last_error(NULL).       */
typedef struct lfgpu_gmsh lfgpu_gmsh;
int lfgpu_gmsh_read_file(const char* filename, int dim_world, lfgpu_gmsh** out);
int lfgpu_gmsh_read_memory(const void* data, int64_t n_bytes, int dim_world, lfgpu_gmsh** out);
void lfgpu_gmsh_destroy(lfgpu_gmsh* g);
/* geometry_order: 1, or 2 when the file holds second-order elements (every output nullable) */
int lfgpu_gmsh_counts(const lfgpu_gmsh* g, int64_t* n_nodes, int64_t* n_explicit_edges, int64_t* n_cells, int* geometry_order,
                      int* n_physical_names);
/* node_coords [n_nodes][2], edge_nodes [n_explicit_edges][2], cell_nodes [n_cells][4] (LFGPU_IDX_NIL in slot 3 of a
 * triangle): the arguments of AddPoint / AddEntity in call order (every output nullable)                            */
int lfgpu_gmsh_arrays(const lfgpu_gmsh* g, double* node_coords, uint32_t* edge_nodes, uint32_t* cell_nodes);
/* GmshReader::PhysicalEntityNr(entity) (gmsh_reader.cc:115-118): returns the number of physical entity numbers of the
 * entity (codim 0 = cell, 1 = edge, 2 = node; edges beyond the explicit ones have none) and writes min(count, capacity) */
int lfgpu_gmsh_physical_entity_nr(const lfgpu_gmsh* g, int codim, int64_t index, int capacity, uint32_t* out);
/* flags[i] = GmshReader::IsPhysicalEntity(entity i, nr) for the first n entities of the codimension: the selectors
 * the examples build for boundary conditions, ready for active_edges / d_fixed style arguments after upload         */
int lfgpu_gmsh_physical_flags(const lfgpu_gmsh* g, int codim, uint32_t nr, int64_t n, uint8_t* flags);
/* $PhysicalNames entry i: number, codimension (= 2 - dimension), name; returns the name's length */
int lfgpu_gmsh_physical_name(const lfgpu_gmsh* g, int i, uint32_t* nr, int* codim, char* buf, int capacity);
/* GmshReader::PhysicalEntityName2Nr / PhysicalEntityNr2Name (gmsh_reader.cc:30-99); codim < 0 = not specified, which is
 * an error when the name / number exists for several codimensions (as in the reference)                             */
int lfgpu_gmsh_physical_name2nr(const lfgpu_gmsh* g, const char* name, int codim, uint32_t* nr);
int lfgpu_gmsh_physical_nr2name(const lfgpu_gmsh* g, uint32_t nr, int codim, char* buf, int capacity);
/* reader.mesh() on the device: upload + edge numbering with the explicit edges first (lfgpu_mesh_build_topology).
 * LFGPU_ERR_UNSUPPORTED for second-order files (TriaO2 / QuadO2 geometries are outside the device path).            */
int lfgpu_gmsh_mesh(lfgpu_ctx* ctx, const lfgpu_gmsh* g, lfgpu_mesh** out);

/* ---- dof maps: stand in for lf::assemble::DofHandler (assemble/dofhandler.h:112-228) ------------------------------- */
/* From any DofHandler: cell_dofs [n_cells][stride] = GlobalDofIndices(cell), n_ldof [n_cells] = NumLocalDofs(cell)
 * (nullable: then 3/4 * ... is derived as the count of non-negative entries).                                         */
int lfgpu_dofmap_upload(lfgpu_ctx* ctx, const lfgpu_mesh* mesh, int64_t n_dofs, int stride, const int64_t* cell_dofs,
                        const uint8_t* n_ldof, lfgpu_dofmap** out);
/* lf::assemble::UniformFEDofHandler(mesh, {{kPoint,n_pt},{kSegment,n_seg},{kTria,n_tria},{kQuad,n_quad}})
 * (assemble/dofhandler.cc:86-284), numbered on the device.                                                            */
int lfgpu_dofmap_uniform(lfgpu_ctx* ctx, lfgpu_mesh* mesh, int n_pt, int n_seg, int n_tria, int n_quad, lfgpu_dofmap** out);
/* dof layout of lf::uscalfe::FeSpaceLagrangeO{1,2,3} (uscalfe/uniform_scalar_fe_space.h:241-342) */
int lfgpu_dofmap_lagrange(lfgpu_ctx* ctx, lfgpu_mesh* mesh, int degree, lfgpu_dofmap** out);
/* lf::assemble::DynamicFEDofHandler(mesh, locdof) (assemble/dofhandler.h:514-789): variable numbers of interior dofs,
 * numbered on the device.  The LOCALDOFINFO functor is passed tabulated: n_int_node [n_nodes], n_int_edge [n_edges],
 * n_int_cell [n_cells] host arrays = locdof(entity) in entity-index order; NULL = 0 for that codimension.  A cell may
 * carry at most 16 local dofs (LFGPU_ERR_UNSUPPORTED otherwise).  The table stride is the longest cell list.           */
int lfgpu_dofmap_dynamic(lfgpu_ctx* ctx, lfgpu_mesh* mesh, const uint32_t* n_int_node, const uint32_t* n_int_edge,
                         const uint32_t* n_int_cell, lfgpu_dofmap** out);
int64_t lfgpu_dofmap_num_dofs(const lfgpu_dofmap* d);
int lfgpu_dofmap_stride(const lfgpu_dofmap* d);
int lfgpu_dofmap_download(lfgpu_ctx* ctx, const lfgpu_dofmap* d, int64_t* cell_dofs, uint8_t* n_ldof);
void lfgpu_dofmap_destroy(lfgpu_dofmap* d);

/* ---- symbolic pass: COOMatrix + makeSparse structure (assemble/coomatrix.h:87-91,172-180) -------------------------- */
/* Builds the compressed pattern Eigen's setFromTriplets would produce for the triplets AssembleMatrixLocally emits
 * (one stored entry per (row dof, col dof) pair that shares a cell, explicit zeros kept, inner indices ascending,
 * int32 indices) plus the per-cell scatter map and the per-row gather lists of the numeric pass.                     */
int lfgpu_symbolic(lfgpu_ctx* ctx, const lfgpu_mesh* mesh, const lfgpu_dofmap* test, const lfgpu_dofmap* trial, int major,
                   lfgpu_pattern** out);
int64_t lfgpu_pattern_nnz(const lfgpu_pattern* p);
int64_t lfgpu_pattern_rows(const lfgpu_pattern* p);
int64_t lfgpu_pattern_cols(const lfgpu_pattern* p);
int lfgpu_pattern_download(lfgpu_ctx* ctx, const lfgpu_pattern* p, int32_t* outer, int32_t* inner);
const int32_t* lfgpu_pattern_outer_device(const lfgpu_pattern* p);
const int32_t* lfgpu_pattern_inner_device(const lfgpu_pattern* p);
/* Only the outer indices flagged in d_keep (device uint8 [n_outer]) have to be produced by later numeric passes; the values of
 * the others are unspecified afterwards.  For the sub-problems of the distributed-ownership scheme: halo rows are incomplete by
 * construction, so the kernels need not spend a generic pass on those of them that do not fit the fast plans.  Must be called
 * before the first numeric pass on the pattern.                                                                              */
int lfgpu_pattern_restrict_rows(lfgpu_ctx* ctx, lfgpu_pattern* p, const uint8_t* d_keep);
void lfgpu_pattern_destroy(lfgpu_pattern* p);

/* ---- numeric pass ---------------------------------------------------------------------------------------------------- */
/* quadrature rule = lf::quad::QuadRule (quad/quad_rule.h): host arrays, points [2][n] (row 0 = x0, row 1 = x1) */
typedef struct {
  int n;
  const double* points;
  const double* weights;
} lfgpu_quad;

/* coefficient = a MeshFunction evaluated at the quadrature points (mesh/utils/mesh_function_traits.h:148-154) */
#define LFGPU_COEFF_CONST 0        /* MeshFunctionConstant<double>        : c[0]                                     */
#define LFGPU_COEFF_CONST_2X2 1    /* MeshFunctionConstant<Matrix2d>      : c[0..3] row-major                        */
#define LFGPU_COEFF_PER_CELL 2     /* data[n_cells]                                                                  */
#define LFGPU_COEFF_PER_QP 3       /* data[n_cells][stride], value at quadrature point k of the cell's rule          */
#define LFGPU_COEFF_PER_QP_2X2 4   /* data[n_cells][stride][4] row-major 2x2                                         */
#define LFGPU_COEFF_NODAL 5        /* data[n_nodes]: a continuous piecewise (bi)linear function given by its values at the mesh
                                      nodes (lf::fe::MeshFunctionFE of a FeSpaceLagrangeO1 function, fe/mesh_function_fe.h),
                                      evaluated at the quadrature points with the cell's vertex shape functions; scalar.
                                      The caller hands over 8 B per node; the library tabulates the values at the quadrature
                                      points on the device (one small kernel per call) and runs the PER_QP kernels, the fast
                                      P1 row kernel included.  Cell terms only.                                            */
typedef struct {
  int kind;
  double c[4];
  const double* data; /* DEVICE pointer for PER_* kinds (lfgpu_malloc / any CUDA allocation on the ctx device)      */
  int64_t stride;     /* values per cell for PER_QP kinds (>= max number of quadrature points)                      */
} lfgpu_coeff;

/* lf::uscalfe::ReactionDiffusionElementMatrixProvider<double,ALPHA,GAMMA>::Eval (uscalfe/loc_comp_ellbvp.h:266-339)
 * for every active cell + AssembleMatrixLocally's scatter (assembler.h:166-179), fused.
 *   degree 1..3 selects FeLagrangeO{1,2,3}{Tria,Quad} (uscalfe/lagr_fe.h); qr_tria / qr_quad NULL = the provider's
 *   default rule make_QuadRule(ref_el, 2*degree) (loc_comp_ellbvp.h:227-228); active (device uint8 [n_cells],
 *   nullable) = provider.isActive(cell); beta = 0 overwrites d_values, beta = 1 accumulates like the reference's
 *   void overload (assembler.h:84-88).  d_values: device array [nnz] in the pattern's order.                        */
int lfgpu_assemble_reaction_diffusion(lfgpu_ctx* ctx, const lfgpu_mesh* mesh, const lfgpu_pattern* pattern, int degree,
                                      const lfgpu_quad* qr_tria, const lfgpu_quad* qr_quad, const lfgpu_coeff* alpha,
                                      const lfgpu_coeff* gamma, const uint8_t* active, double beta, double* d_values,
                                      int algo);
/* Same, restricted to the outer indices (matrix rows for LFGPU_ROW_MAJOR) listed in d_row_list (device int32 [n_rows]);
 * only those rows of d_values are written.  Building block of the multi-GPU row partition.  LFGPU_ALGO_GATHER only.    */
int lfgpu_assemble_reaction_diffusion_rows(lfgpu_ctx* ctx, const lfgpu_mesh* mesh, const lfgpu_pattern* pattern, int degree,
                                           const lfgpu_quad* qr_tria, const lfgpu_quad* qr_quad, const lfgpu_coeff* alpha,
                                           const lfgpu_coeff* gamma, const uint8_t* active, double beta, double* d_values,
                                           int algo, const int32_t* d_row_list, int64_t n_rows);
/* lf::uscalfe::ScalarLoadElementVectorProvider<double,F>::Eval (loc_comp_ellbvp.h:691-746) + AssembleVectorLocally's
 * scatter (assembler.h:322-324).  d_vec: device array [n_dofs].  algo: LFGPU_ALGO_ATOMIC (= AUTO) one thread per cell
 * and FP64 atomics; LFGPU_ALGO_GATHER one thread per dof adding its cells' entries in ascending cell order, i.e. the
 * reference's order of additions (deterministic; the per-dof lists are built on first use and cached in the dofmap). */
int lfgpu_assemble_load(lfgpu_ctx* ctx, const lfgpu_mesh* mesh, const lfgpu_dofmap* dofmap, int degree,
                        const lfgpu_quad* qr_tria, const lfgpu_quad* qr_quad, const lfgpu_coeff* f, const uint8_t* active,
                        double beta, double* d_vec, int algo);
/* Same, restricted to the contiguous outer range [row0, row0 + n_rows): the row partition of a multi-GPU run by row
 * blocks (every cell is active).  Runs in the kernels that own rows in registers -- the P1 vertex-fan kernel, the P2 / P3
 * row kernels (triangles, constant coefficients, default rule, beta = 0); LFGPU_ERR_UNSUPPORTED otherwise -- pass the
 * rows as a list to lfgpu_assemble_reaction_diffusion_rows then.                                                    */
int lfgpu_assemble_reaction_diffusion_range(lfgpu_ctx* ctx, const lfgpu_mesh* mesh, const lfgpu_pattern* pattern, int degree,
                                            const lfgpu_quad* qr_tria, const lfgpu_quad* qr_quad, const lfgpu_coeff* alpha,
                                            const lfgpu_coeff* gamma, double beta, double* d_values, int algo, int64_t row0,
                                            int64_t n_rows);
/* Host-buffer form of the matrix assembly -- the call a CPU-side user of AssembleMatrixLocally (assembler.h:114-186)
 * makes: this step's node coordinates come from host memory (h_node_coords [n_nodes][2], NULL = keep the device copy),
 * every cell is active, the values are overwritten in d_values (device, [nnz]) and copied to h_values (host [nnz],
 * NULL = no download).  Returns when h_values is complete.  Page-locked host buffers (lfgpu_host_alloc_pinned) let
 * upload, kernel and download overlap: the outer indices are processed in n_blocks (<= 0: default 16) contiguous
 * blocks, each computed as soon as the leading part of the coordinate array it refers to is on the device and
 * downloaded while later blocks are still uploading.  Results are identical to
 * lfgpu_mesh_update_node_coords + lfgpu_assemble_reaction_diffusion + lfgpu_memcpy_d2h.                             */
int lfgpu_assemble_reaction_diffusion_host(lfgpu_ctx* ctx, lfgpu_mesh* mesh, const lfgpu_pattern* pattern, int degree,
                                           const lfgpu_quad* qr_tria, const lfgpu_quad* qr_quad, const lfgpu_coeff* alpha,
                                           const lfgpu_coeff* gamma, const double* h_node_coords, double* d_values,
                                           double* h_values, int algo, int n_blocks);
/* The same for the contiguous outer range [row0, row0 + n_rows) -- the share of one GPU when the rows are split into
 * blocks (lehrfempp_b200/distributed.py, mode "owner_rows"): h_node_coords is still the FULL coordinate array, of
 * which only the window the range refers to is uploaded; h_values_range receives the values of the range only (its
 * element 0 is the first value of row row0).  Fan kernel only (LFGPU_ERR_UNSUPPORTED otherwise).                     */
int lfgpu_assemble_reaction_diffusion_host_range(lfgpu_ctx* ctx, lfgpu_mesh* mesh, const lfgpu_pattern* pattern, int degree,
                                                 const lfgpu_quad* qr_tria, const lfgpu_quad* qr_quad, const lfgpu_coeff* alpha,
                                                 const lfgpu_coeff* gamma, const double* h_node_coords, double* d_values,
                                                 double* h_values_range, int algo, int n_blocks, int64_t row0, int64_t n_rows);
/* ---- edge (codim-1) contributions (second "next" row) ------------------------------------------- */
/* AssembleMatrixLocally(1, dofh, dofh, MassEdgeMatrixProvider(fe_space, gamma[, rule], edge_selector), A)
 * (uscalfe/loc_comp_ellbvp.h:367-529, assembler.h:114-186 with codim 1): for every active edge the mass matrix
 * sum_k w_k |e| gamma(x_k) phi_a phi_b of FeLagrangeO<degree>Segment is ADDED to d_values (the reference accumulates
 * into the COO matrix that already holds the cell terms; the entries are part of the pattern of the symbolic pass).
 * dofmap: from lfgpu_dofmap_lagrange(degree) (edge dofs must be known).  qr_segment: points[n] on [0,1], NULL = the
 * provider's default make_QuadRule(kSegment, 2*degree).  gamma: CONST, PER_CELL (one value per EDGE) or PER_QP (per
 * edge and point).  active_edges: device uint8 [n_edges] = the EDGESELECTOR, NULL = all edges.                      */
int lfgpu_assemble_edge_mass(lfgpu_ctx* ctx, lfgpu_mesh* mesh, const lfgpu_dofmap* dofmap, const lfgpu_pattern* pattern, int degree,
                             const lfgpu_quad* qr_segment, const lfgpu_coeff* gamma, const uint8_t* active_edges, double* d_values);
/* AssembleVectorLocally(1, dofh, ScalarLoadEdgeVectorProvider(fe_space, g[, rule], edge_selector), vec)
 * (uscalfe/loc_comp_ellbvp.h:784-921): accumulates into d_vec [n_dofs].                                             */
int lfgpu_assemble_edge_load(lfgpu_ctx* ctx, lfgpu_mesh* mesh, const lfgpu_dofmap* dofmap, int degree, const lfgpu_quad* qr_segment,
                             const lfgpu_coeff* g, const uint8_t* active_edges, double* d_vec);
/* The same two operations for an explicit list of straight segments -- the form a host caller that owns the
 * DofHandler uses: it passes only the ACTIVE edges (d_seg_xy device [n][4] = x0 y0 x1 y1 of the edge geometry,
 * d_seg_dofs device int32 [n][degree + 1] = DofHandler::GlobalDofIndices(edge)); coefficient tables are per segment.  */
int lfgpu_assemble_segment_mass(lfgpu_ctx* ctx, const lfgpu_pattern* pattern, int degree, const lfgpu_quad* qr_segment,
                                int64_t n_segments, const double* d_seg_xy, const int32_t* d_seg_dofs, const lfgpu_coeff* gamma,
                                double* d_values);
int lfgpu_assemble_segment_load(lfgpu_ctx* ctx, int degree, const lfgpu_quad* qr_segment, int64_t n_segments, const double* d_seg_xy,
                                const int32_t* d_seg_dofs, const lfgpu_coeff* g, int64_t n_dofs, double* d_vec);
/* global coordinates of every edge's quadrature points, SegmentO1::Global (geometry/segment_o1.cc:9-11):
 * d_out device [n_edges][nq_stride][2] -- lets a host tabulate MeshFunctionGlobal lambdas into PER_QP edge tables    */
int lfgpu_edge_qp_coords(lfgpu_ctx* ctx, lfgpu_mesh* mesh, int degree, const lfgpu_quad* qr_segment, int nq_stride, double* d_out);
/* edges with exactly one adjacent cell (mesh/utils flagEntitiesOnBoundary(mesh, 1)): d_flags device uint8 [n_edges]  */
int lfgpu_mesh_boundary_edges(lfgpu_ctx* ctx, lfgpu_mesh* mesh, uint8_t* d_flags);
/* nodes on the boundary (flagEntitiesOnBoundary(mesh, 2)): d_node_flags device uint8 [n_nodes]                         */
int lfgpu_mesh_boundary_nodes(lfgpu_ctx* ctx, lfgpu_mesh* mesh, uint8_t* d_node_flags);
/* dofs on the boundary -- the selector of the examples, `boundary(dofh.Entity(dof))` with flagEntitiesOnBoundary(mesh)
 * (examples/ellbvp_linfe/homDir_linfe_demo.cc:158-165, fe/test/loc_comp_tests.cc:86-91): d_dof_flags device uint8
 * [n_dofs], ready as d_fixed of lfgpu_fix_flagged_solution_components.  Needs a dof map numbered on the device
 * (lfgpu_dofmap_uniform / _lagrange), where dof -> entity is arithmetic; LFGPU_ERR_UNSUPPORTED for uploaded tables. */
int lfgpu_dofmap_boundary_dofs(lfgpu_ctx* ctx, lfgpu_mesh* mesh, const lfgpu_dofmap* dofmap, uint8_t* d_dof_flags);
/* lf::fe::InitEssentialConditionFromFunction (fe/fe_tools.h:301-356) in two device steps.  (1) the dofs of the selected edges
 * -- their interior dofs and those of their end points, GlobalDofIndices(edge) -- as flags: d_edge_sel device uint8 [n_edges]
 * (e.g. lfgpu_mesh_boundary_edges or a physical group), d_flags device uint8 [n_dofs].  (2) the position of every dof
 * (the interpolation node of its Lagrange shape function: vertices; points at 1/2 resp. 1/3, 2/3 of an edge in the edge's
 * own direction; cell interior nodes), d_xy device [n_dofs][2]: the prescribed value of a flagged dof is g at that point
 * (NodalValuesToDofs of the Lagrange elements is the identity).  n_tria / n_quad: interior dofs per cell of the layout
 * (0/0, 0/1, 1/4 for degree 1, 2, 3).  Device-numbered uniform layouts only.                                          */
int lfgpu_dofmap_edge_dof_flags(lfgpu_ctx* ctx, lfgpu_mesh* mesh, const lfgpu_dofmap* dofmap, const uint8_t* d_edge_sel, uint8_t* d_flags);
int lfgpu_dofmap_dof_coords(lfgpu_ctx* ctx, lfgpu_mesh* mesh, const lfgpu_dofmap* dofmap, int n_tria, int n_quad, double* d_xy);
/* ---- essential boundary conditions (first "next" row) ------------------------------------------- */
/* lf::assemble::FixFlaggedSolutionComponents (assemble/fix_dof.h:86-138) on the compressed matrix: with xhat = the
 * prescribed values on the fixed dofs and 0 elsewhere,  rhs -= A * xhat;  rhs[fixed] = xhat;  every entry in a fixed row
 * or column is erased (COOMatrix::setZero, assemble/coomatrix.h:108-115) and the diagonal of a fixed dof becomes 1.
 * d_values / d_rhs are edited in place (erased entries stay as explicit zeros in the pattern of the symbolic pass).
 * If d_outer_out [n+1], d_inner_out [>= nnz], d_values_out [>= nnz] are all non-null the erased entries are also
 * removed, which yields exactly makeSparse() of the reference's edited triplet list; *nnz_out = entries kept.
 * d_fixed: device uint8 [n_dofs] (non-zero = fixed), d_fixed_values: device double [n_dofs] (read where fixed).
 * Errors: non-square matrix (fix_dof.h:90 "Matrix must be square!"), fixed dof without a diagonal entry.            */
int lfgpu_fix_flagged_solution_components(lfgpu_ctx* ctx, const lfgpu_pattern* pattern, double* d_values, double* d_rhs,
                                          const uint8_t* d_fixed, const double* d_fixed_values, int32_t* d_outer_out,
                                          int32_t* d_inner_out, double* d_values_out, int64_t* nnz_out);
/* lf::assemble::FixFlaggedSolutionCompAlt (assemble/fix_dof.h:181-218), the non-symmetric variant: only the ROWS of the
 * fixed dofs become unit rows, rhs[fixed] = xhat, everything else untouched.  Same arguments as above.              */
int lfgpu_fix_flagged_solution_comp_alt(lfgpu_ctx* ctx, const lfgpu_pattern* pattern, double* d_values, double* d_rhs,
                                        const uint8_t* d_fixed, const double* d_fixed_values, int32_t* d_outer_out,
                                        int32_t* d_inner_out, double* d_values_out, int64_t* nnz_out);
/* ---- consumer side: the assembled matrix used in place ------------------------------------ */
/* y = A x on the compressed arrays (either storage order); d_x [cols], d_y [rows] device                             */
int lfgpu_spmv(lfgpu_ctx* ctx, const lfgpu_pattern* pattern, const double* d_values, const double* d_x, double* d_y);
/* Conjugate gradients for the symmetric positive definite systems the path produces (what the reference hands to an
 * Eigen solver after makeSparse(), examples/ellbvp_linfe/homDir_linfe_demo.cc:166-175).  d_x: initial guess in, solution
 * out.  Stops when ||r||_2 <= rel_tol * ||b||_2 or after max_iter iterations; jacobi != 0 = diagonal preconditioner.
 * iters_out / rel_res_out (nullable): iterations done, final relative residual.                                      */
int lfgpu_cg_solve(lfgpu_ctx* ctx, const lfgpu_pattern* pattern, const double* d_values, const double* d_rhs, double* d_x,
                   double rel_tol, int max_iter, int jacobi, int* iters_out, double* rel_res_out);
/* ---- multi-GPU building blocks (DESIGN.md "Multi-GPU"; the reference is serial) --------------------------------------- */
/* Row segments values[outer[r] .. outer[r+1]) of the listed rows <-> a contiguous message buffer.  d_rows device int32
 * [n_rows], d_offsets device int64 [n_rows] = start of each row's segment inside the buffer.  unpack_add ADDS the
 * received partial sums of interface rows into the owner's values.                                                    */
int lfgpu_rows_pack(lfgpu_ctx* ctx, const lfgpu_pattern* pattern, const int32_t* d_rows, int64_t n_rows, const int64_t* d_offsets,
                    const double* d_values, double* d_buf);
int lfgpu_rows_unpack_add(lfgpu_ctx* ctx, const lfgpu_pattern* pattern, const int32_t* d_rows, int64_t n_rows,
                          const int64_t* d_offsets, const double* d_buf, double* d_values);
/* ---- distributed ownership: "the mesh is partitioned across the GPUs ... by Morton-ordered cell ranges, with each GPU owning
 * the matrix rows for its cells" (BASELINE.json north_star).  The reference's cell loop (assemble/assembler.h:125-180) is
 * serial; these calls cut it into per-GPU sub-problems that are ordinary lfgpu_mesh / lfgpu_dofmap pairs with LOCAL indices,
 * so that symbolic pass, plans and numeric kernels run unchanged on 1/N of the matrix (int32 indices stay valid when the global
 * number of stored values exceeds 2^31).  The local numbering is the order-preserving restriction of the global one: the local
 * pattern of an OWNED row, mapped through local -> global, is bit-exactly the global pattern of that row; values are added in
 * the reference's cell order.                                                                                                */
/* cell -> part: cells sorted by the Morton code of their centroid, n_parts contiguous ranges of equal cell counts.
 * d_cell_part: device uint8 [n_cells].                                                                                       */
int lfgpu_partition_morton(lfgpu_ctx* ctx, const lfgpu_mesh* mesh, int n_parts, uint8_t* d_cell_part);
/* dof -> owner = the lowest part with a cell touching the dof; d_dof_owner: device uint8 [n_dofs] (255 = untouched)           */
int lfgpu_partition_dof_owner(lfgpu_ctx* ctx, const lfgpu_dofmap* dofmap, const uint8_t* d_cell_part, uint8_t* d_dof_owner);
/* d_cell_sel [n_cells] = 1 for the cells of part `rank` (halo == 0: every rank assembles its own cells, interface rows are
 * summed at their owner) or for every cell that touches a dof owned by `rank` (halo != 0: owner-computes, the rows a rank owns
 * are complete without any exchange)                                                                                           */
int lfgpu_partition_select_cells(lfgpu_ctx* ctx, const lfgpu_dofmap* dofmap, const uint8_t* d_cell_part, const uint8_t* d_dof_owner,
                                 int rank, int halo, uint8_t* d_cell_sel);
/* the selected cells as a mesh + dof map of their own, and the local -> global index lists (ascending)                        */
typedef struct lfgpu_submesh lfgpu_submesh;
int lfgpu_submesh_extract(lfgpu_ctx* ctx, const lfgpu_mesh* mesh, const lfgpu_dofmap* dofmap, const uint8_t* d_cell_sel,
                          lfgpu_submesh** out);
void lfgpu_submesh_destroy(lfgpu_submesh* s);     /* also destroys its mesh and dof map */
lfgpu_mesh* lfgpu_submesh_mesh(lfgpu_submesh* s);     /* owned by the submesh */
lfgpu_dofmap* lfgpu_submesh_dofmap(lfgpu_submesh* s); /* owned by the submesh */
int lfgpu_submesh_counts(const lfgpu_submesh* s, int64_t* n_cells, int64_t* n_nodes, int64_t* n_dofs);
const int32_t* lfgpu_submesh_l2g_cells_device(const lfgpu_submesh* s);
const int32_t* lfgpu_submesh_l2g_nodes_device(const lfgpu_submesh* s);
const int32_t* lfgpu_submesh_l2g_dofs_device(const lfgpu_submesh* s);
/* d_owned [n_local_dofs] = 1 where the local dof is owned by `rank` (d_dof_owner: the GLOBAL owner array)                     */
int lfgpu_submesh_owned_dofs(lfgpu_ctx* ctx, const lfgpu_submesh* s, const uint8_t* d_dof_owner, int rank, uint8_t* d_owned);
/* ---- several GPUs from one process ("lfgpu_ctx_create(device_ids, n_dev, ...)") ---------------------------
 * The drop-in form of the distributed-ownership scheme for a C / C++ caller: one lfgpu_ctx per listed device inside one handle.
 * lfgpu_multi_setup takes the flattened mesh and dof table exactly as lfgpu_mesh_upload / lfgpu_dofmap_upload do (host arrays of
 * the WHOLE problem), cuts them into one sub-problem per device (Morton cell ranges, every device owns the rows of its cells,
 * one-cell halo) and runs the symbolic pass per device; afterwards no device holds more than its share.  The numeric pass queues
 * the kernels on every device and waits for all of them; it needs no exchange between the devices.                              */
typedef struct lfgpu_multi lfgpu_multi;
int lfgpu_multi_create(const int* device_ids, int n_dev, lfgpu_multi** out);
void lfgpu_multi_destroy(lfgpu_multi* m);
int lfgpu_multi_num_devices(const lfgpu_multi* m);
lfgpu_ctx* lfgpu_multi_ctx(lfgpu_multi* m, int k); /* the context of device k (owned by the handle) */
const char* lfgpu_multi_last_error(const lfgpu_multi* m);
int lfgpu_multi_setup(lfgpu_multi* m, int64_t n_nodes, const double* node_coords, int64_t n_cells, const uint32_t* cell_nodes,
                      const double* cell_coords, int64_t n_dofs, int stride, const int64_t* cell_dofs, const uint8_t* n_ldof, int major);
int lfgpu_multi_set_zero(lfgpu_multi* m);
/* AssembleMatrixLocally with ReactionDiffusionElementMatrixProvider on all devices.  alpha / gamma as in
 * lfgpu_assemble_reaction_diffusion, except that the tables of the PER_* kinds are HOST arrays over the cells of the WHOLE mesh
 * (every device receives the entries of its cells).  accumulate != 0: add to what the matrix holds (assembler.h:84-88).        */
int lfgpu_multi_assemble_reaction_diffusion(lfgpu_multi* m, int degree, const lfgpu_quad* qr_tria, const lfgpu_quad* qr_quad,
                                            const lfgpu_coeff* alpha, const lfgpu_coeff* gamma, int accumulate);
/* share of device k: n_rows / nnz of the rows it OWNS; cells, rows and stored values of its sub-problem incl. the halo           */
int lfgpu_multi_part_sizes(const lfgpu_multi* m, int k, int64_t* n_rows, int64_t* nnz, int64_t* n_local_cells, int64_t* n_local_rows,
                           int64_t* n_local_nnz);
/* the rows device k owns as a compressed block with GLOBAL indices: rows [n_rows] (global outer index, ascending), row_ptr
 * [n_rows + 1], cols [nnz] (global inner index, ascending inside a row = the reference's pattern), values [nnz]; every output
 * nullable.  The blocks of all devices together are the matrix makeSparse() returns (assemble/coomatrix.h:172-180).             */
int lfgpu_multi_part_download(lfgpu_multi* m, int k, int64_t* rows, int64_t* row_ptr, int32_t* cols, double* values);
/* read-only device views used by the host-side partitioner: gather lists (items = cell << 4 | local index), mesh arrays.
 * READ-ONLY: node positions change only through lfgpu_mesh_update_node_coords or the host-buffer assembly calls -- the row-kernel
 * plans keep their own reordered copies of the positions and refresh them when those calls have moved the mesh's version counter. */
const int32_t* lfgpu_pattern_adj_ptr_device(const lfgpu_pattern* p);
const uint32_t* lfgpu_pattern_adj_device(const lfgpu_pattern* p);
int64_t lfgpu_pattern_num_items(const lfgpu_pattern* p);
const double* lfgpu_mesh_node_coords_device(const lfgpu_mesh* m);
const uint32_t* lfgpu_mesh_cell_nodes_device(const lfgpu_mesh* m);
int lfgpu_ctx_wait_event(lfgpu_ctx* ctx, void* cuda_event); /* ctx stream waits for a cudaEvent_t recorded elsewhere */

/* global coordinates of every cell's quadrature points, Geometry::Global (tria_o1.cc:70-74, quad_o1.cc:68-83):
 * d_out device [n_cells][nq_stride][2]; lets a host evaluate MeshFunctionGlobal lambdas into PER_QP tables.         */
int lfgpu_qp_coords(lfgpu_ctx* ctx, const lfgpu_mesh* mesh, int degree, const lfgpu_quad* qr_tria, const lfgpu_quad* qr_quad,
                    int nq_stride, double* d_out);
/* tabulated reference element data the numeric pass uses (PrecomputedScalarReferenceFiniteElement,
 * uscalfe/precomputed_scalar_reference_finite_element.h:72-78): host outputs, phi [nsf][nq], grad [nsf][2*nq]
 * with columns (2k, 2k+1) = (d/dx0, d/dx1) at point k.  Returns nsf (or <0). cell_type 3 = tria, 4 = quad.          */
int lfgpu_fe_tabulate(int degree, int cell_type, const lfgpu_quad* qr, double* phi, double* grad);
/* default rule make_QuadRule(ref_el, degree) (quad/make_quad_rule.cc:21-157): returns n; points [2][n], weights [n] */
int lfgpu_default_quad_rule(int cell_type, int degree, int capacity, double* points, double* weights);

#ifdef __cplusplus
}
#endif
#endif /* LFGPU_H */
