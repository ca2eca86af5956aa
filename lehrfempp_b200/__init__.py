"""lehrfempp_b200 -- B200-native finite-element assembly behind LehrFEM++'s assembler API.

Thin ctypes binding over liblfgpu.so (the C ABI declared in include/lfgpu.h).  The compute path is hand-written
CUDA for sm_100a inside that library; this module is host plumbing only (handles, numpy <-> device copies).
There is no CPU fallback: importing works anywhere, but creating a Context without a CUDA device raises.

Reference interfaces mirrored (paths relative to the LehrFEM++ checkout):
  lf::assemble::AssembleMatrixLocally / AssembleVectorLocally   lib/lf/assemble/assembler.h:114-186, 298-327
  lf::assemble::UniformFEDofHandler                              lib/lf/assemble/dofhandler.h:260-503
  lf::uscalfe::ReactionDiffusionElementMatrixProvider            lib/lf/uscalfe/loc_comp_ellbvp.h:85-339
  lf::uscalfe::ScalarLoadElementVectorProvider                   lib/lf/uscalfe/loc_comp_ellbvp.h:562-746
  lf::assemble::DynamicFEDofHandler                              lib/lf/assemble/dofhandler.h:514-789
  lf::io::GmshReader                                             lib/lf/io/gmsh_reader.h:55-200, gmsh_reader.cc
"""
from .api import (  # noqa: F401
    ALGO_ATOMIC,
    ALGO_AUTO,
    ALGO_FAN,
    ALGO_GATHER,
    COL_MAJOR,
    ROW_MAJOR,
    Coeff,
    Context,
    DeviceArray,
    DofMap,
    GmshReader,
    LfgpuError,
    Mesh,
    MultiAssembler,
    Pattern,
    QuadRule,
    SubMesh,
    build_library,
    default_quad_rule,
    fe_tabulate,
    library_path,
)

__all__ = [
    "ALGO_ATOMIC", "ALGO_AUTO", "ALGO_FAN", "ALGO_GATHER", "COL_MAJOR", "ROW_MAJOR", "Coeff", "Context", "DeviceArray", "DofMap",
    "GmshReader",    "LfgpuError", "Mesh", "MultiAssembler", "Pattern", "QuadRule", "SubMesh", "build_library", "default_quad_rule", "fe_tabulate", "library_path",
]
