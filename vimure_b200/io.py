"""Edgelist parser: DataFrame (ego, alter, reporter[, layer, weight]) -> X (COO) + reporter mask.

Restates `vimure._io.read_from_edgelist / read_from_csv` (reference `_io.py:132-323`, checks `_io.py:366-513`) with the
same arguments, warnings and errors, but vectorised and without ever materialising the mask: when R is not given the
reference builds, per reporter and layer, a sparse "row m + column m minus the (m,m) tie" matrix (`_io.py:229-242`,
nnz(R) = 2*L*M*(N-1)); here that is `masks.EgoMask(diag=False)`.
"""
import warnings

import numpy as np
import pandas as pd

from . import masks
from .sptensor import sptensor


class RealNetwork:
    """Observed network parsed from real data (counterpart of `vimure._io.RealNetwork`, `_io.py:93-128`)."""

    def __init__(self, X, R, L, N, M, K, nodeNames=None, layerNames=None):
        self.X, self.R, self.L, self.N, self.M, self.K = X, R, L, N, M, K
        if nodeNames is not None:
            self.nodeNames = pd.DataFrame(nodeNames.items(), columns=["id", "name"])
        if layerNames is not None:
            self.layerNames = layerNames

    def getX(self):
        return self.X

    def __repr__(self):
        return f"{self.__class__.__name__} (N={self.N}, M={self.M}, L={self.L}, K={self.K})"


def _check_params_consistency(df, nodes, reporters, reporter, layer, ego, alter, weight):
    """reference `_io.py:366-480`"""
    if not isinstance(df, pd.DataFrame):
        raise ValueError(f"'df' should be a DataFrame, instead it is of type: {type(df)}.")
    missing = [c for c in (ego, alter, reporter) if c not in df.columns]
    if missing:
        raise ValueError(
            f"Required columns not found in data frame: {', '.join(missing)}. Mapping used: "
            f"ego='{ego}', alter='{alter}', reporter='{reporter}'. "
            "Hint: Use params ego,alter,... for mapping column names.")
    if nodes is not None and not isinstance(nodes, list):
        raise ValueError(f"'nodes' should be a list, instead it is of type: {type(nodes)}.")
    if reporters is not None and not isinstance(reporters, list):
        raise ValueError(f"'reporters' should be a list, instead it is of type: {type(reporters)}.")
    if nodes == [] or nodes is None:
        warnings.warn(f"The set of nodes was not informed, using {ego} and {alter} columns to infer nodes.", UserWarning)
        nodes = pd.concat([df[ego], df[alter]]).unique().tolist()
    if np.logical_or(~df[ego].isin(nodes), ~df[alter].isin(nodes)).any():
        raise ValueError("A list of nodes was informed, but it does not contain all nodes in the data frame.")
    df = df.copy()
    if layer not in df.columns:
        df.loc[:, layer] = "1"
    if weight not in df.columns:
        df.loc[:, weight] = 1
    unsupported = ("This survey setup is not currently supported by the package: "
                   " some reporters are not nodes in the network. "
                   "Hint: If this is unexpected behaviour, "
                   f"compare the unique values of the `{str(reporter)}` column "
                   f"with those of the `{str(ego)}` and `{str(alter)}` columns.")
    reporters_in_df = df[reporter].unique().tolist()
    if reporters is None or reporters == []:
        warnings.warn("The set of reporters was not informed, assuming set(reporters) = set(nodes) and N = M.",
                      UserWarning)
        reporters = nodes[:]
        if not set(reporters_in_df).issubset(reporters):
            raise ValueError(unsupported)
    elif not set(reporters_in_df).issubset(reporters):
        raise ValueError("Some reporters in the data frame do not appear in the list of reporters provided. "
                         f"Hint: Compare the unique values of the `{str(reporter)}` column "
                         "with the list of reporters passed as parameter.")
    if not set(reporters).issubset(nodes):
        raise ValueError(unsupported)
    if not set(nodes).issubset(reporters):
        warnings.warn("Not necessarily a problem, but some of the nodes are not reporters.", UserWarning)
    return df, nodes, reporters


def read_from_edgelist(df, nodes=[], reporters=[], is_weighted=False, is_undirected=False, reporter="reporter",
                       layer="layer", ego="ego", alter="alter", weight="weight", K=None, R=None, **kwargs):
    """Parse an edgelist into a `RealNetwork` (X: sptensor of shape (L, N, N, N); R: reporter mask)."""
    df, nodes, reporters = _check_params_consistency(df, nodes, reporters, reporter, layer, ego, alter, weight)
    layers = sorted(df[layer].unique())
    L, N, M = len(layers), len(nodes), len(reporters)
    df = df[[ego, alter, reporter, layer, weight]].drop_duplicates()

    node_id = pd.Series(np.arange(N), index=pd.Index(nodes))
    layer_id = pd.Series(np.arange(L), index=pd.Index(layers))
    i = node_id.reindex(df[ego].values).to_numpy()
    j = node_id.reindex(df[alter].values).to_numpy()
    m = node_id.reindex(df[reporter].values).to_numpy()
    l = layer_id.reindex(df[layer].values).to_numpy()
    w = df[weight].to_numpy()
    data = w if is_weighted else (w > 0).astype("int")
    keep = data > 0
    l, i, j, m, data = (a[keep].astype(np.int64) for a in (l, i, j, m, data))

    if is_undirected and len(data):  # element-wise max(X, X^T) per (reporter, layer)   (_io.py:275-277)
        key = ((l * N + np.minimum(i, j)) * N + np.maximum(i, j)) * N + m
        order = np.argsort(key, kind="stable")
        ks = key[order]
        first = np.concatenate([[True], ks[1:] != ks[:-1]])
        grp = np.cumsum(first) - 1
        mx = np.zeros(grp[-1] + 1, dtype=data.dtype)
        np.maximum.at(mx, grp, data[order])
        lo, io, jo, mo = l[order][first], np.minimum(i, j)[order][first], np.maximum(i, j)[order][first], m[order][first]
        offd = io != jo
        l = np.concatenate([lo, lo[offd]])
        i = np.concatenate([io, jo[offd]])
        j = np.concatenate([jo, io[offd]])
        m = np.concatenate([mo, mo[offd]])
        data = np.concatenate([mx, mx[offd]])

    # same entry order as sptensor_from_list (utils.py:154-175): layer-major, then reporter, then input order
    order = np.lexsort((np.arange(len(l)), m, l))
    X = sptensor((l[order], i[order], j[order], m[order]), data[order], shape=(L, N, N, N))

    if R is None:
        warnings.warn("Reporters Mask was not informed (parameter R). Parser will build it from reporter column, "
                      "assuming a reporter can only report their own ties.", UserWarning)
        rep = np.zeros((L, N), dtype=np.uint8)
        rep[:, node_id.reindex(reporters).to_numpy()] = 1
        R = masks.EgoMask(L, N, N, rep=rep, diag=False)
    elif tuple(R.shape) != (L, N, N, M):
        raise ValueError("Dimensions of reporter mask (R) do not match L x N x N x M")

    if K is None:
        K = int(np.max(X.vals)) + 1
        warnings.warn(f"Parameter K was None. Defaulting to: {K}", UserWarning)

    return RealNetwork(X=X, R=R, L=L, N=N, M=M, K=K, nodeNames={k: v for k, v in enumerate(nodes)}, layerNames=layers)


def read_from_csv(filename, **kwargs):
    """reference `_io.py:297-321`"""
    return read_from_edgelist(pd.read_csv(filename), **kwargs)
