"""Host helper mirroring ``decagon/utility/preprocessing.py:20-26`` of the reference.

Only ``sparse_to_tuple`` is on the hot path (the reference's ``get_sparse_mat`` is
python-2 only and unused).
"""
import numpy as np
import scipy.sparse as sp


def sparse_to_tuple(sparse_mx):
    """scipy sparse -> ``(coords int32 [nnz, 2], values, shape)``.

    Same contract as the reference (``preprocessing.py:20-26``): the matrix is viewed
    as COO (converted when needed) and the coordinates come back as one ``[nnz, 2]``
    array whose column 0 is the row index and column 1 the column index.
    """
    coo = sparse_mx if sp.isspmatrix_coo(sparse_mx) else sparse_mx.tocoo()
    coords = np.stack([coo.row, coo.col], axis=1)
    return coords, coo.data, coo.shape
