"""Variable initialisers, drop-in for ``decagon/deep/inits.py``.

TF's RNG is unseeded in the reference; here the draws come from a module-level numpy
``RandomState`` that ``set_seed`` re-seeds (it is separate from the global ``np.random`` stream the
minibatch iterator consumes, like TF's generator is)."""
import numpy as np

from .. import tf_compat as tf

_rng = np.random.RandomState()


def set_seed(seed):
    global _rng
    _rng = np.random.RandomState(seed)


def weight_variable_glorot(input_dim, output_dim, name=""):
    """Glorot & Bengio uniform init, float32 (``inits.py:5-12``)."""
    init_range = np.sqrt(6.0 / (input_dim + output_dim))
    initial = _rng.uniform(-init_range, init_range, size=(input_dim, output_dim)).astype(np.float32)
    return tf.Variable(initial, name=name)


def zeros(input_dim, output_dim, name=None):
    return tf.Variable(np.zeros((input_dim, output_dim), dtype=np.float32), name=name or '')


def ones(input_dim, output_dim, name=None):
    return tf.Variable(np.ones((input_dim, output_dim), dtype=np.float32), name=name or '')
