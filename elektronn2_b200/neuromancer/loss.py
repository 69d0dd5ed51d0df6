"""Loss-side nodes used by every BASELINE config: Softmax, MultinoulliNLL (sparse
targets), AggregateLoss, Classification / Errors.

Host-side mirror of neuromancer/loss.py:33-93, 141-351, 700-826, 1279-1370.  On the
device the chain Softmax -> MultinoulliNLL -> AggregateLoss (+ Errors) is ONE fused
kernel pair (e2_softmax_nll_fwd / _bwd); the individual nodes exist so that model
files read exactly like the reference's.
"""
import numpy as np

from .graphutils import TaggedShape
from .node_basic import Node
from .variables import VariableParam

EPS = 1e-5


class Softmax(Node):
    """loss.py:33-93."""

    def __init__(self, parent, n_class='auto', n_indep=1, name="softmax", print_repr=True):
        super(Softmax, self).__init__(parent, name, print_repr)
        n_f = parent.shape['f']
        if getattr(parent, 'activation_func', 'lin') != 'lin':
            raise ValueError("The parent of a Softmax-node must have a linear activation function.")
        if n_class == 'auto':
            if n_f % n_indep != 0:
                raise ValueError("Cannot create %i-fold %i-class softmax from %i features."
                                 % (n_indep, n_f // n_indep, n_f))
            n_class = n_f // n_indep
        elif n_class * n_indep != n_f:
            raise ValueError("Cannot create %i-fold %i-class softmax from %i features." % (n_indep, n_class, n_f))
        if n_indep != 1:
            raise NotImplementedError("n_indep > 1 is not on the B200 path")
        self.n_class, self.n_indep = n_class, n_indep


class MultinoulliNLL(Node):
    """loss.py:141-351 (sparse integer targets, no class/example weights or masks)."""

    def __init__(self, pred, target, target_is_sparse=False, class_weights=None, example_weights=None,
                 mask_class_labeled=None, mask_class_not_present=None, name="nll", print_repr=True):
        super(MultinoulliNLL, self).__init__([pred, target], name, print_repr)
        if not target_is_sparse:
            raise NotImplementedError("dense (one-hot) targets are not on the B200 path")
        if any(v is not None for v in (class_weights, example_weights, mask_class_labeled, mask_class_not_present)):
            raise NotImplementedError("class/example weights and masks are not on the B200 path")
        if not isinstance(pred, Softmax):
            raise NotImplementedError("MultinoulliNLL expects a Softmax prediction node")
        self.pred, self.target = pred, target
        self.target_is_sparse = True
        self.axis = pred.shape.tag2index('f')
        self.n_class, self.n_indep = pred.n_class, pred.n_indep
        for i, (a, b) in enumerate(zip(pred.shape.shape, target.shape.shape)):
            if i != self.axis and a != b:
                raise ValueError("Prediction %s and target %s shapes do not match" % (pred.shape, target.shape))

    def _calc_shape(self):
        self.shape = self.parents[0].shape.updateshape(self.axis, 1)


class AggregateLoss(Node):
    """Mean of the (weighted) means of its parents (loss.py:1279-1370)."""

    def __init__(self, parent_nodes, mixing_weights=None, name="total_loss", print_repr=True):
        if not isinstance(parent_nodes, (tuple, list)):
            parent_nodes = [parent_nodes]
        super(AggregateLoss, self).__init__(list(parent_nodes), name, print_repr)
        if mixing_weights is None:
            mixing_weights = np.ones(len(parent_nodes))
        if len(parent_nodes) != len(mixing_weights):
            raise ValueError("Mismatch: len(parent_nodes)=%i, len(weights)=%i" % (len(parent_nodes), len(mixing_weights)))
        if len(parent_nodes) != 1:
            raise NotImplementedError("multi-target losses are not on the B200 path")
        self.mixing_weights = VariableParam(np.array(mixing_weights, 'float32'), name="loss_mixing_weights",
                                            apply_train=False, apply_reg=False)
        self.params['mixing_weights'] = self.mixing_weights

    def _calc_shape(self):
        self.shape = TaggedShape([1], ['f'])

    def _calc_comp_cost(self):
        self.computational_cost = int(sum(p.shape.stripnone_prod for p in self.parents))


class Classification(Node):
    """argmax over the class axis (loss.py:700-757)."""

    def __init__(self, pred, n_class='auto', n_indep='auto', name="cls", print_repr=True):
        super(Classification, self).__init__(pred, name, print_repr)
        self.pred = pred
        self.n_class = getattr(pred, 'n_class', pred.shape['f'])
        self.n_indep = getattr(pred, 'n_indep', 1)

    def _calc_shape(self):
        self.shape = self.parents[0].shape.updateshape(self.parents[0].shape.tag2index('f'), 1)


class _Errors(Node):
    """mean(int16(target) != argmax(pred)) (loss.py:760-814)."""

    def __init__(self, cls, target, target_is_sparse=False, name="errors", print_repr=True):
        super(_Errors, self).__init__([cls, target], name, print_repr)
        if not target_is_sparse:
            raise NotImplementedError("dense (one-hot) targets are not on the B200 path")
        self.cls, self.target = cls, target

    def _calc_shape(self):
        self.shape = TaggedShape([1], ['f'])


def Errors(pred, target, target_is_sparse=False, n_class='auto', n_indep='auto', name="errors", print_repr=True):
    if not isinstance(pred, Classification):
        pred = Classification(pred, n_class=n_class, n_indep=n_indep, name='cls for errors', print_repr=False)
    return _Errors(pred, target, target_is_sparse=target_is_sparse, name=name, print_repr=print_repr)
