"""``DecagonModel`` with the reference's constructor and attribute surface
(``decagon/deep/model.py:48-137``): two multi-relation graph-convolution layers per (i, j) group,
ReLU after the sum over the groups of a node type in layer 1, plain sum in layer 2, and the
per-relation decoder matrices ``latent_inters`` / ``latent_varies``.  Building the model creates
handles and variables only; ``decagon_b200.session.Session.run`` executes it on the GPU.
"""
from collections import defaultdict

from .. import tf_compat as tf
from .layers import GraphConvolutionMulti, GraphConvolutionSparseMulti, \
    DistMultDecoder, InnerProductDecoder, DEDICOMDecoder, BilinearDecoder, identity

flags = tf.app.flags
FLAGS = flags.FLAGS

_DECODER_CLASSES = {'innerproduct': InnerProductDecoder, 'distmult': DistMultDecoder,
                    'bilinear': BilinearDecoder, 'dedicom': DEDICOMDecoder}


class Model(object):
    def __init__(self, **kwargs):
        allowed_kwargs = {'name', 'logging'}
        for kwarg in kwargs.keys():
            assert kwarg in allowed_kwargs, 'Invalid keyword argument: ' + kwarg
        name = kwargs.get('name')
        if not name:
            name = self.__class__.__name__.lower()
        self.name = name
        self.logging = kwargs.get('logging', False)
        self.vars = {}

    def _build(self):
        raise NotImplementedError

    def build(self):
        with tf.variable_scope(self.name):
            self._build()
        self.vars = {var.name: var for var in self._variables()}

    def fit(self):
        pass

    def predict(self):
        pass


class DecagonModel(Model):
    def __init__(self, placeholders, num_feat, nonzero_feat, edge_types, decoders, **kwargs):
        super(DecagonModel, self).__init__(**kwargs)
        self.edge_types = edge_types
        self.num_edge_types = sum(self.edge_types.values())
        self.num_obj_types = max([i for i, _ in self.edge_types]) + 1
        self.decoders = decoders
        self.inputs = {i: placeholders['feat_%d' % i] for i, _ in self.edge_types}
        self.input_dim = num_feat
        self.nonzero_feat = nonzero_feat
        self.placeholders = placeholders
        self.dropout = placeholders['dropout']
        self.adj_mats = {et: [
            placeholders['adj_mats_%d,%d,%d' % (et[0], et[1], k)] for k in range(n)]
            for et, n in self.edge_types.items()}
        self.hidden1_dim, self.hidden2_dim = FLAGS.hidden1, FLAGS.hidden2
        self.engine = None  # created by Session.run once the adjacency tuples are fed
        self.build()
        tf.MODELS.add(self)

    def _build(self):
        self.layer1, self.layer2 = {}, {}
        self.hidden1 = {}
        for i, j in self.edge_types:
            self.layer1[i, j] = GraphConvolutionSparseMulti(
                input_dim=self.input_dim, output_dim=self.hidden1_dim,
                edge_type=(i, j), num_types=self.edge_types[i, j],
                adj_mats=self.adj_mats, nonzero_feat=self.nonzero_feat,
                act=identity, dropout=self.dropout, logging=self.logging)
            self.layer1[i, j].model = self
            self.layer1[i, j](self.inputs[j])
            # relu(add_n(...)) over the groups of node type i (model.py:74-75)
            self.hidden1[i] = tf.Tensor('hidden1', self, i)

        self.embeddings_reltyp = defaultdict(list)
        for i, j in self.edge_types:
            self.layer2[i, j] = GraphConvolutionMulti(
                input_dim=self.hidden1_dim, output_dim=self.hidden2_dim,
                edge_type=(i, j), num_types=self.edge_types[i, j],
                adj_mats=self.adj_mats, act=identity,
                dropout=self.dropout, logging=self.logging)
            self.layer2[i, j].model = self
            self.embeddings_reltyp[i].append(self.layer2[i, j](self.hidden1[j]))

        # plain add_n, no activation (model.py:85-88)
        self.embeddings = [None] * self.num_obj_types
        for i in self.embeddings_reltyp:
            self.embeddings[i] = tf.Tensor('embeddings', self, i)

        self.edge_type2decoder = {}
        for i, j in self.edge_types:
            decoder = self.decoders[i, j]
            if decoder not in _DECODER_CLASSES:
                raise ValueError('Unknown decoder type')
            self.edge_type2decoder[i, j] = _DECODER_CLASSES[decoder](
                input_dim=self.hidden2_dim, logging=self.logging,
                edge_type=(i, j), num_types=self.edge_types[i, j],
                act=identity, dropout=self.dropout)
            self.edge_type2decoder[i, j].model = self

        # glb / loc of every flat relation (model.py:116-137)
        self.latent_inters, self.latent_varies = [], []
        for edge_type in self.edge_types:
            for k in range(self.edge_types[edge_type]):
                r = len(self.latent_inters)
                self.latent_inters.append(tf.Tensor('latent_inter', self, r))
                self.latent_varies.append(tf.Tensor('latent_vary', self, r))

        self._bind_variables()

    def groups(self):
        return list(self.edge_types)

    def _bind_variables(self):
        """Give every variable its (kind, group, k) slot in the engine's parameter arena."""
        from .._lib import PARAM_W1, PARAM_W2, PARAM_DEC_GLOBAL, PARAM_DEC_LOCAL
        self._var_list = []
        for g in self.edge_types:
            for kind, layers in ((PARAM_W1, self.layer1), (PARAM_W2, self.layer2)):
                for k in range(self.edge_types[g]):
                    v = layers[g].vars['weights_%d' % k]
                    v.model, v.slot = self, (kind, g, k)
                    self._var_list.append(v)
            dec = self.edge_type2decoder[g]
            for name, v in dec.vars.items():
                v.model = self
                if name == 'global_interaction':
                    v.slot = (PARAM_DEC_GLOBAL, g, None)
                else:
                    v.slot = (PARAM_DEC_LOCAL, g, int(name.rsplit('_', 1)[1]))
                self._var_list.append(v)

    def _variables(self):
        return list(self._var_list)
