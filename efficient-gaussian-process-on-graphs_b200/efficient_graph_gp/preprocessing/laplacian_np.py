"""Dense Laplacians for the GPflow kernels -- API of ``preprocessing/laplacian_np.py:3-35``.

Here a zero-degree node is given degree 1 before the scaling (so its row is
just the identity row), unlike ``graph_kernels/utils.py``; both behaviours are
kept, each under its own name, as in the reference.
"""

import numpy as np


def get_normalized_laplacian(W):
    W = np.asarray(W)
    degrees = np.sum(W, axis=1)
    dis = 1.0 / np.sqrt(np.where(degrees > 0, degrees, 1.0))
    return np.eye(W.shape[0]) - (dis[:, None] * W) * dis[None, :]


def get_laplacian(W):
    W = np.asarray(W)
    return np.diag(np.sum(W, axis=1)) - W
