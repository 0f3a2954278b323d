"""Drop-in for ``utils_sparse/graph_utils.py:5-30`` (sparse normalized Laplacian).

The reference forms ``D^-1/2 (D - A) D^-1/2`` with two scipy SpGEMMs against
diagonal matrices.  A product with a diagonal matrix is one multiply per stored
entry, so the same values -- bit for bit, ``(dis[i] * L_ij) * dis[j]`` in that
order -- come out of one O(nnz) pass; zero-degree rows become empty exactly as
``1/sqrt(0) -> inf -> 0`` makes them in the reference (graph_utils.py:20-22).
"""

import numpy as np
import scipy.sparse as sp


def get_normalized_laplacian(adj_matrix):
    """Normalized Laplacian ``D^-1/2 (D - A) D^-1/2`` as scipy CSR (sorted indices)."""
    A = adj_matrix.tocsr()
    degrees = np.array(A.sum(axis=1)).flatten()
    with np.errstate(divide="ignore"):
        d_inv_sqrt = 1.0 / np.sqrt(degrees)
    d_inv_sqrt[np.isinf(d_inv_sqrt)] = 0

    L = (sp.diags(degrees, format="csr") - A).tocsr()
    L.sum_duplicates()
    L.sort_indices()
    n = L.shape[0]
    rows = np.repeat(np.arange(n), np.diff(L.indptr))
    vals = (d_inv_sqrt[rows] * L.data) * d_inv_sqrt[L.indices]
    # scipy's SpGEMM drops entries whose product is exactly zero (isolated rows / columns)
    keep = ((d_inv_sqrt[rows] * L.data) != 0) & (vals != 0)
    counts = np.bincount(rows[keep], minlength=n)
    indptr = np.zeros(n + 1, dtype=L.indptr.dtype)
    np.cumsum(counts, out=indptr[1:])
    out = sp.csr_matrix((vals[keep], L.indices[keep], indptr), shape=L.shape)
    out.has_sorted_indices = True
    return out
