"""Drop-in for ``graph_kernels/utils.py:6-28`` (dense normalized Laplacian).

The reference multiplies by two dense diagonal matrices (2 N^3 flops); a
product with a diagonal matrix is a row / column scaling, so the same entries
-- bit for bit, ``(dis[i] * W_ij) * dis[j]`` -- cost O(N^2) here.  An isolated
node keeps its diagonal 1 (a self-loop of weight 1), as in the reference.
"""

import numpy as np
import scipy.sparse as sp


def get_normalized_laplacian(W, sparse=False):
    """``I - D^-1/2 W D^-1/2`` (ndarray, or scipy sparse when ``sparse=True``)."""
    if sparse:
        degrees = np.array(W.sum(axis=1)).flatten()
        with np.errstate(divide="ignore"):
            dis = np.where(degrees > 0, 1.0 / np.sqrt(degrees), 0.0)
        D_inv_sqrt = sp.diags(dis)
        return sp.eye(W.shape[0]) - D_inv_sqrt @ W @ D_inv_sqrt
    W = np.asarray(W)
    degrees = np.sum(W, axis=1)
    dis = np.zeros(W.shape[0], dtype=float)
    valid = degrees > 0
    dis[valid] = 1.0 / np.sqrt(degrees[valid])
    return np.eye(W.shape[0]) - (dis[:, None] * W) * dis[None, :]
