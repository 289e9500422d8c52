"""Two-layer dense GCN with the reference's API (src/models/gcn.py:9-34)."""
import torch
import torch.nn.functional as F

from ..utils.graph import normalize_adjacency_matrix
from .sampling import FactoredGraph
from .layers import MetaDenseGraphConvolution, MetaModule, get_subdict, sparse_companion

# When true, a sampled graph is propagated through plain differentiable torch ops instead of the custom
# kernels' once-differentiable Functions: needed where a later backward differentiates THROUGH this
# forward's gradient (the unrolled inner steps, src/trainers/inner.py:71).
_DOUBLE_BACKWARD = [False]
# Factored (unrolled) route: run the feature products from a CSR copy of mostly-zero features. Off = dense F.linear with the
# reference's full-matrix dropout mask (tests that compare dropout masks with the dense composable route).
SPARSE_FEATURES = [True]


class double_backward_path:
    def __enter__(self):
        self.previous = _DOUBLE_BACKWARD[0]
        _DOUBLE_BACKWARD[0] = True

    def __exit__(self, *exc):
        _DOUBLE_BACKWARD[0] = self.previous


class MetaDenseGCN(MetaModule):

    def __init__(self, in_features, hidden_features, out_features, dropout, normalize_adj: bool = True):
        super().__init__()
        self.layer_in = MetaDenseGraphConvolution(in_features, hidden_features)
        self.layer_out = MetaDenseGraphConvolution(hidden_features, out_features)
        self.dropout = dropout
        self.normalize_adj = normalize_adj

    def reset_weights(self):
        self.layer_in.reset_weights()
        self.layer_out.reset_weights()

    def forward_to_last_layer(self, node_features, dense_adj, params=None):
        sparse = None
        if isinstance(dense_adj, FactoredGraph):
            if not self.normalize_adj:
                raise NotImplementedError("a FactoredGraph carries the self-looped sample; it is propagated in normalised form only")
            dense_adj = dense_adj.normalized()
            sparse = sparse_companion(node_features) if SPARSE_FEATURES[0] else None
        elif self.normalize_adj:
            factored_ok = not (_DOUBLE_BACKWARD[0] and torch.is_grad_enabled())
            dense_adj = normalize_adjacency_matrix(dense_adj, materialize=not factored_ok)
        if sparse is not None:
            # bag-of-words features: dropout only has to touch the non-zeros (zeros stay zero under any mask), and the feature
            # products run from CSR (layers._SparseProduct) instead of dense N x F SGEMMs
            hidden = sparse.with_values(F.dropout(sparse.val, self.dropout, training=self.training))
        else:
            hidden = F.dropout(node_features, self.dropout, training=self.training)
        hidden = F.relu(self.layer_in(hidden, dense_adj, params=get_subdict(params, "layer_in")))
        hidden = F.dropout(hidden, self.dropout, training=self.training)
        return self.layer_out(hidden, dense_adj, params=get_subdict(params, "layer_out"))

    def forward(self, node_features, dense_adj, params=None):
        return F.log_softmax(self.forward_to_last_layer(node_features, dense_adj, params=params), dim=1)
