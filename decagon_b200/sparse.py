"""Boundary types the minibatch iterator requires of every adjacency matrix.

Mirrors the attribute contract of ``main/Utils/Sparse.py:5-73`` in the reference:
scipy CSR / COO matrices that additionally carry

* ``id``                 -- process-unique string used as a dictionary key
                            (``minibatch.py:51-52``),
* ``isTranspose``        -- set on BOTH members of a transposed pair
                            (``Sparse.py:57-62``),
* ``transposedMtxLink``  -- the other member of the pair (``minibatch.py:130-135``).

``transpose(copy=..., setId=True)`` is what ``DecagonDataSet._augmentAdjMtxDictWithTranspose``
(``DecagonDataSet.py:212-231``) calls to create the transposed twin of every relation.
"""
import itertools
import os

import scipy.sparse as sp

_coo_ids = itertools.count()
_csr_ids = itertools.count()


def _link(a, b):
    a.isTranspose = True
    a.transposedMtxLink = b
    b.isTranspose = True
    b.transposedMtxLink = a


class RelationCooMatrix(sp.coo_matrix):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.isTranspose = False
        self.transposedMtxLink = None
        self.id = "RelationCooMatrix|%d|%d" % (os.getpid(), next(_coo_ids))

    def transpose(self, axes=None, copy=False, setId=False):
        twin = RelationCooMatrix(super().transpose(axes, copy))
        if setId:
            _link(self, twin)
        return twin

    def tocsr(self, copy=False):
        out = RelationCsrMatrix(super().tocsr(copy))
        out.id, out.isTranspose, out.transposedMtxLink = self.id, self.isTranspose, self.transposedMtxLink
        return out


class RelationCsrMatrix(sp.csr_matrix):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.isTranspose = False
        self.transposedMtxLink = None
        self.id = "RelationCsrMatrix|%d|%d" % (os.getpid(), next(_csr_ids))

    def transpose(self, axes=None, copy=False, setId=False):
        twin = RelationCsrMatrix(super().transpose(axes, copy))
        if setId:
            _link(self, twin)
        return twin

    def tocoo(self, copy=False):
        out = RelationCooMatrix(super().tocoo(copy))
        out.id, out.isTranspose, out.transposedMtxLink = self.id, self.isTranspose, self.transposedMtxLink
        return out
