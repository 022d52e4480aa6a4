"""Layer classes with the reference's names, constructor signatures and variable names
(``decagon/deep/layers.py``).  They own variables and describe the computation; the arithmetic is
the CUDA engine's (``decagon_b200/csrc``):

* ``GraphConvolutionSparseMulti`` (``layers.py:70-94``)  -> fused multi-relation SpMM, layer 1
* ``GraphConvolutionMulti``       (``layers.py:97-118``) -> projection + fused SpMM, layer 2
* decoders (``layers.py:121-213``) hold the per-relation parameters; calling one yields the
  all-pairs score tensors of its relations (``optimizer.predictions`` is the path the reference
  actually evaluates).

Only ``act = identity`` inside the graph-convolution layers is supported: ``DecagonModel`` always
passes ``lambda x: x`` (``model.py:71,82``) and a per-relation non-linearity would forbid fusing the
sum over relations.  Anything else raises ``NotImplementedError``.
"""
import numpy as np

from .. import tf_compat as tf
from . import inits

flags = tf.app.flags
FLAGS = flags.FLAGS

_LAYER_UIDS = {}


def get_layer_uid(layer_name=''):
    _LAYER_UIDS[layer_name] = _LAYER_UIDS.get(layer_name, 0) + 1
    return _LAYER_UIDS[layer_name]


def identity(x):
    return x


def relu(x):
    return np.maximum(x, 0)


def sigmoid(x):
    return 1. / (1 + np.exp(-x))


def _is_identity(act):
    probe = np.array([-1.5, 0.0, 2.0], dtype=np.float32)
    try:
        return np.array_equal(np.asarray(act(probe)), probe)
    except Exception:
        return False


class MultiLayer(object):
    """Base layer: ``edge_type`` = (i, j) group, ``num_types`` = number of relations K."""

    def __init__(self, edge_type=(), num_types=-1, **kwargs):
        self.edge_type = edge_type
        self.num_types = num_types
        allowed_kwargs = {'name', 'logging'}
        for kwarg in kwargs.keys():
            assert kwarg in allowed_kwargs, 'Invalid keyword argument: ' + kwarg
        name = kwargs.get('name')
        if not name:
            layer = self.__class__.__name__.lower()
            name = layer + '_' + str(get_layer_uid(layer))
        self.name = name
        self.vars = {}
        self.logging = kwargs.get('logging', False)
        self.issparse = False

    def _call(self, inputs):
        return inputs

    def __call__(self, inputs):
        with tf.name_scope(self.name):
            return self._call(inputs)


class _GraphConvolution(MultiLayer):
    layer_no = 0

    def _check_act(self):
        if not _is_identity(self.act):
            raise NotImplementedError(
                '%s: only act = identity is supported inside the layer (the reference model passes '
                'lambda x: x, model.py:71,82); ReLU is applied after the sum over groups' % self.name)

    def _call(self, inputs):
        self.inputs = inputs
        return tf.Tensor('layer%d_group' % self.layer_no, self, self.edge_type)


class GraphConvolutionSparseMulti(_GraphConvolution):
    """Graph convolution layer for sparse inputs: l2norm(sum_k A_k (dropout_k(X) W_k))."""
    layer_no = 1

    def __init__(self, input_dim, output_dim, adj_mats, nonzero_feat, dropout=0., act=relu, **kwargs):
        super(GraphConvolutionSparseMulti, self).__init__(**kwargs)
        self.dropout, self.adj_mats, self.act = dropout, adj_mats, act
        self.issparse = True
        self.nonzero_feat = nonzero_feat
        self.input_dim, self.output_dim = input_dim[self.edge_type[1]], output_dim
        self._check_act()
        with tf.variable_scope('%s_vars' % self.name):
            for k in range(self.num_types):
                self.vars['weights_%d' % k] = inits.weight_variable_glorot(
                    self.input_dim, output_dim, name='weights_%d' % k)


class GraphConvolutionMulti(_GraphConvolution):
    """Dense-input graph convolution: l2norm(sum_k A_k (dropout_k(H) W_k))."""
    layer_no = 2

    def __init__(self, input_dim, output_dim, adj_mats, dropout=0., act=relu, **kwargs):
        super(GraphConvolutionMulti, self).__init__(**kwargs)
        self.adj_mats, self.dropout, self.act = adj_mats, dropout, act
        self.input_dim, self.output_dim = input_dim, output_dim
        self._check_act()
        with tf.variable_scope('%s_vars' % self.name):
            for k in range(self.num_types):
                self.vars['weights_%d' % k] = inits.weight_variable_glorot(
                    input_dim, output_dim, name='weights_%d' % k)


class _Decoder(MultiLayer):
    kind = None

    def __init__(self, input_dim, dropout=0., act=sigmoid, **kwargs):
        super(_Decoder, self).__init__(**kwargs)
        self.dropout, self.act, self.input_dim = dropout, act, input_dim
        with tf.variable_scope('%s_vars' % self.name):
            self._make_vars(input_dim)

    def _make_vars(self, input_dim):
        pass

    def _call(self, inputs):
        """``inputs``: {node type: embedding tensor}.  Returns K tensors act(Z_i M_k Z_j^T)
        (``layers.py:135-147`` etc.; the decoder-side dropout of the reference is not applied --
        this call is never on the reference's train / eval path)."""
        self.inputs = inputs
        return [tf.Tensor('decoder_scores', self, k) for k in range(self.num_types)]


class DEDICOMDecoder(_Decoder):
    """DEDICOM tensor factorisation decoder: Z_i diag(d_k) R diag(d_k) Z_j^T."""
    kind = 'dedicom'

    def _make_vars(self, input_dim):
        self.vars['global_interaction'] = inits.weight_variable_glorot(input_dim, input_dim, name='global_interaction')
        for k in range(self.num_types):
            tmp = inits.weight_variable_glorot(input_dim, 1, name='local_variation_%d' % k)
            tmp.initial = tmp.initial.reshape(-1)  # tf.reshape(tmp, [-1]) (layers.py:133)
            tmp.shape = tmp.initial.shape
            self.vars['local_variation_%d' % k] = tmp


class DistMultDecoder(_Decoder):
    """DistMult decoder: Z_i diag(r_k) Z_j^T."""
    kind = 'distmult'

    def _make_vars(self, input_dim):
        for k in range(self.num_types):
            tmp = inits.weight_variable_glorot(input_dim, 1, name='relation_%d' % k)
            tmp.initial = tmp.initial.reshape(-1)
            tmp.shape = tmp.initial.shape
            self.vars['relation_%d' % k] = tmp


class BilinearDecoder(_Decoder):
    """Bilinear decoder: Z_i M_k Z_j^T."""
    kind = 'bilinear'

    def _make_vars(self, input_dim):
        for k in range(self.num_types):
            self.vars['relation_%d' % k] = inits.weight_variable_glorot(input_dim, input_dim, name='relation_%d' % k)


class InnerProductDecoder(_Decoder):
    """Inner-product decoder: Z_i Z_j^T."""
    kind = 'innerproduct'
