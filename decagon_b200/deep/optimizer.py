"""``DecagonOptimizer`` with the reference's constructor and attribute surface
(``decagon/deep/optimizer.py:9-127``).

What the reference builds as TF ops is one fused device step here:
* negative sampling per relation, P(v) ~ degrees[i][k][v]^0.75 (``optimizer.py:36-49``) -- only the
  sampler of the batch's relation runs (the reference executes all R and keeps one);
* ``batch_predict`` + ``diag_part`` (``:51-57,63-85``) -- B bilinear forms instead of a B x B matrix;
* hinge loss (``:109,116-120``; ``loss='xent'`` selects ``:122-127``), backward, TF1 Adam
  (``:111-113``).
The attributes are fetch handles for ``Session.run``.
"""
from .. import tf_compat as tf

flags = tf.app.flags
FLAGS = flags.FLAGS


class DecagonOptimizer(object):
    def __init__(self, embeddings, latent_inters, latent_varies,
                 degrees, edge_types, edge_type2dim, placeholders,
                 margin=0.1, neg_sample_weights=1., batch_size=100, loss='hinge'):
        self.embeddings = embeddings
        self.latent_inters = latent_inters
        self.latent_varies = latent_varies
        self.edge_types = edge_types
        self.degrees = degrees
        self.edge_type2dim = edge_type2dim
        self.obj_type2n = {i: self.edge_type2dim[i, j][0][0] for i, j in self.edge_types}
        self.margin = margin
        self.neg_sample_weights = neg_sample_weights
        self.batch_size = batch_size
        if loss not in ('hinge', 'xent'):
            raise ValueError('Unknown loss kind')
        self.loss_kind = loss
        self.learning_rate = FLAGS.learning_rate

        self.model = next(e for e in embeddings if e is not None).owner
        self.model.optimizer = self
        self.placeholders = placeholders
        self.inputs = placeholders['batch']
        self.batch_edge_type_idx = placeholders['batch_edge_type_idx']
        self.batch_row_edge_type = placeholders['batch_row_edge_type']
        self.batch_col_edge_type = placeholders['batch_col_edge_type']

        # range_max = len(degrees[i][k]) must be the node count of the row type (optimizer.py:45)
        for i, j in self.edge_types:
            for k in range(self.edge_types[i, j]):
                if len(self.degrees[i][k]) != self.obj_type2n[i]:
                    raise ValueError('degrees[%d][%d] has %d entries, node type %d has %d nodes'
                                     % (i, k, len(self.degrees[i][k]), i, self.obj_type2n[i]))

        self.row_inputs = tf.Tensor('row_inputs', self)
        self.col_inputs = tf.Tensor('col_inputs', self)
        self.neg_samples = tf.Tensor('neg_samples', self)
        self.preds = tf.Tensor('preds', self)
        self.outputs = tf.Tensor('outputs', self)
        self.neg_preds = tf.Tensor('neg_preds', self)
        self.neg_outputs = tf.Tensor('neg_outputs', self)
        self.predictions = tf.Tensor('predictions', self)
        self.cost = tf.Tensor('cost', self)
        self.opt_op = tf.Tensor('opt_op', self)
        self.grads_vars = [(tf.Tensor('grad', self, v), v) for v in self.model._variables()]
