"""cellula_b200 -- B200-native (sm_100a) hot path of gdagstn/cellula: counts -> SNN graph.

    from cellula_b200 import doNormAndReduce, makeGraphsAndClusters

mirrors the reference's R entry points (see api.py); `Context` is the ctypes binding of the
plain C ABI (include/cellula_b200.h).  Importing the package never touches the GPU; the first
call does, and raises if libcellula_b200.so has not been built or no B200 is present.
"""
from ._lib import Cb200Error, Context, MultiContext, build, lib  # noqa: F401
from .api import (SerialParam, approxSilhouette, doNormAndReduce, get_context, makeGraphsAndClusters,  # noqa: F401
                  makeSNNGraph, options, pairwiseModularity, rebuildModularity, release_context)
from .sce import CSC, ReducedDim, SingleCellExperiment, SNNGraph  # noqa: F401

__version__ = "0.1"
