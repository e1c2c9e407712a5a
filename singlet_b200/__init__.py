"""singlet_b200 -- B200-native (sm_100a) ALS-NMF engine behind the reference's run_nmf /
cross_validate_nmf / ard_nmf / project_model interface (zdebruine/singlet).

Importing the package never touches the GPU; the shared library ``libsinglet_cuda.so`` is loaded on
first use and there is no CPU fallback (``SingletCudaError`` is raised instead).
"""
from ._lib import SingletCudaError  # noqa: F401
from .api import (GetBestRank, Handle, Rcpp_predict, ard_nmf, c_ard_nmf, c_ard_nmf_dense, c_ard_nmf_sparse_list, c_linked_nmf, c_nmf, c_nmf_dense,  # noqa: F401
                  c_nmf_sparse_list, c_project_model, weight_by_split, cross_validate_nmf, default_handle, project_model, run_nmf, set_seed)
from .datasets import get_pbmc3k_data, log_normalize  # noqa: F401

__version__ = "0.1.0"
