"""cellscopes.jl_b200 -- B200-native (sm_100a) implementation of the CellScopes.jl
counts -> normalise -> HVG -> scale -> PCA -> kNN/SNN path.

The directory name carries a dot, so it cannot be imported by name; use the loader at the repo
root (``import csb200 as cs``) or ``importlib`` with ``submodule_search_locations``.

Layout: ``csrc/`` CUDA kernels + the C ABI (include/cellscopes_cuda.h), ``_capi`` ctypes binding,
``processing``/``objects`` the host-side mirror of the reference's operator interface,
``pipeline`` the device-resident, cell-sharded orchestration, ``synth`` the synthetic workloads.
"""
from . import _capi, ingest, integration, objects, pipeline, processing, synth  # noqa: F401
from .integration import harmony_matrix, run_harmony  # noqa: F401
from .ingest import merge_matrices, read_10x, read_10x_h5, run_tf_idf  # noqa: F401
from ._capi import CellScopesCudaError, Context, load_library  # noqa: F401
from .objects import (ClusteringObject, NormCountObject, PCAObject, RawCountObject, ReductionObject,  # noqa: F401
                      ScaleCountObject, UMAPObject, VariableGeneObject, scRNAObject, subset_count, subset_matrix)
from .pipeline import LocalComm, NcclComm, PathParams, TorchComm, partition_cells, run_path  # noqa: F401
from .processing import (find_all_markers, find_markers, find_variable_genes, normalize_object, run_clustering, run_pca,
                         scale_object)  # noqa: F401

__version__ = "0.1.0"
from ._capi import csc_from_arrays  # noqa: E402,F401
