"""Minimal stand-in for the `SingleCellExperiment` container the reference passes between its
modules (SURVEY.md layer L0): assays as dgCMatrix-style CSC triples, colData / rowData as
dicts of columns, reducedDims, metadata.  Only what doNormAndReduce() /
makeGraphsAndClusters() read or write is modelled."""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


@dataclass
class CSC:
    """dgCMatrix slots: i (0-based row = gene index, sorted per column), p, x, Dim."""
    p: np.ndarray
    i: np.ndarray
    x: np.ndarray
    shape: tuple  # (n_genes, n_cells)


@dataclass
class SNNGraph:
    """What bluster::makeSNNGraph returns, as plain arrays: undirected weighted graph on
    n vertices; from/to are 1-based like igraph::as_edgelist."""
    n: int
    frm: np.ndarray
    to: np.ndarray
    weight: np.ndarray
    k: int = 0
    type: str = "rank"


class ReducedDim(np.ndarray):
    """N x d matrix that can carry runPCA's attributes (varExplained, percentVar, rotation)."""

    def __new__(cls, arr, **attrs):
        obj = np.asarray(arr).view(cls)
        obj.attrs = dict(attrs)
        return obj

    def __array_finalize__(self, obj):
        self.attrs = getattr(obj, "attrs", {})


@dataclass
class SingleCellExperiment:
    assays: dict = field(default_factory=dict)
    colData: dict = field(default_factory=dict)
    rowData: dict = field(default_factory=dict)
    reducedDims: dict = field(default_factory=dict)
    metadata: dict = field(default_factory=dict)
    rownames: list | None = None
    colnames: list | None = None

    @classmethod
    def from_counts(cls, colptr, rowidx, x, n_genes, rownames=None, colnames=None):
        colptr = np.asarray(colptr)
        return cls(assays={"counts": CSC(colptr, np.asarray(rowidx, dtype=np.int32),
                                         np.asarray(x, dtype=np.float64), (int(n_genes), colptr.size - 1))},
                   rownames=rownames, colnames=colnames)

    @property
    def counts(self) -> CSC:
        return self.assays["counts"]

    def nrow(self):
        return self.counts.shape[0]

    def ncol(self):
        return self.counts.shape[1]
