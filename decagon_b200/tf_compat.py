"""The handful of TensorFlow-1 names the reference's hot-path callers touch, backed by the CUDA
engine instead of a TF graph.

The reference drives its model through ``tf.placeholder`` dictionaries and
``tf.Session.run(fetches, feed_dict)`` (``DecagonDataSet.py:84-120``, ``DecagonTrainer.py:94-100``,
``DecagonAccuracyEvaluator.py:122``).  A maintainer switches a caller over by replacing
``import tensorflow as tf`` with ``from decagon_b200 import tf_compat as tf``.
Nothing here computes: placeholders, variables and tensors are plain handles, and
``Session.run`` (``decagon_b200.session``) translates a fetch list into C-ABI calls.
"""
import contextlib
import types
import weakref

import numpy as np

float32, int32, int64 = np.float32, np.int32, np.int64

_scope = []
MODELS = weakref.WeakSet()  # every DecagonModel alive: what global_variables_initializer() covers


@contextlib.contextmanager
def name_scope(name):
    _scope.append(name)
    try:
        yield
    finally:
        _scope.pop()


variable_scope = name_scope


def current_scope():
    return '/'.join(_scope)


class Placeholder(object):
    """Hashable feed-dict key (``tf.placeholder`` / ``tf.sparse_placeholder``)."""

    def __init__(self, dtype, shape=None, name=None, sparse=False, default=None):
        self.dtype, self.shape, self.name, self.sparse, self.default = dtype, shape, name, sparse, default

    def __repr__(self):
        return 'Placeholder(%s)' % (self.name,)


def placeholder(dtype, shape=None, name=None):
    return Placeholder(dtype, shape, name)


def sparse_placeholder(dtype, shape=None, name=None):
    return Placeholder(dtype, shape, name, sparse=True)


def placeholder_with_default(input, shape, name=None):
    return Placeholder(np.asarray(input).dtype, shape, name, default=input)


class Variable(object):
    """A trainable tensor.  ``initial`` is the float32 numpy value uploaded by
    ``global_variables_initializer``; once a model owns it, ``slot`` = (kind, group, k) names its
    place in the engine's parameter arena."""

    def __init__(self, initial, name=''):
        self.initial = np.asarray(initial, dtype=np.float32)
        scope = current_scope()
        self.name = (scope + '/' if scope else '') + name + ':0'
        self.shape = self.initial.shape
        self.model, self.slot = None, None

    def __repr__(self):
        return 'Variable(%s, shape=%s)' % (self.name, self.shape)


class Tensor(object):
    """Symbolic fetch handle: ``kind`` says what ``Session.run`` has to produce, ``owner`` is the
    model / optimizer / layer it belongs to."""

    def __init__(self, kind, owner, index=None):
        self.kind, self.owner, self.index = kind, owner, index

    def __repr__(self):
        return 'Tensor(%s, %s)' % (self.kind, self.index)


class InitOp(object):
    pass


def global_variables_initializer():
    return InitOp()


class ConfigProto(object):
    intra_op_parallelism_threads = 0
    inter_op_parallelism_threads = 0


class _Flags(object):
    """``tf.app.flags``: DEFINE_* register a default, FLAGS.<name> reads it back."""

    def __init__(self):
        object.__setattr__(self, '_values', {})

    def _define(self, name, default, doc=''):
        self._values[name] = default

    DEFINE_integer = DEFINE_float = DEFINE_boolean = DEFINE_string = _define

    @property
    def FLAGS(self):
        return self

    def __getattr__(self, name):
        try:
            return object.__getattribute__(self, '_values')[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        self._values[name] = value

    def __contains__(self, name):
        return name in self._values


flags = _Flags()
FLAGS = flags
app = types.SimpleNamespace(flags=flags)
# defaults of the reference's toy script (main.py:227-238); DecagonDataSet._getFlags overrides them
flags.DEFINE_integer('hidden1', 64)
flags.DEFINE_integer('hidden2', 32)
flags.DEFINE_float('learning_rate', 0.001)


def Session(config=None, **kwargs):
    from .session import Session as _Session
    return _Session(config=config, **kwargs)
