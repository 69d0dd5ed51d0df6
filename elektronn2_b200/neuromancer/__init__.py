"""Drop-in face of ``elektronn2.neuromancer`` for the volumetric-CNN hot path
(reference: elektronn2/neuromancer/__init__.py)."""
from .graphutils import TaggedShape, floatX, as_floatX  # noqa: F401
from .variables import VariableParam, VariableWeight, ConstantParam, initweights  # noqa: F401
from .node_basic import Node, Input, Input_like, Concat, model_manager, choose_name  # noqa: F401
from .neural import Conv, UpConv, Pool, Crop, FragmentsToDense, AutoMerge, UpConvMerge  # noqa: F401
from .loss import Softmax, MultinoulliNLL, AggregateLoss, Classification, Errors  # noqa: F401
from .model import Model, rebuild_model, kernel_lists_from_node_descr, modelload, params_from_model_file  # noqa: F401
from . import graphmanager, model, computations  # noqa: F401
from . import optimiser  # noqa: F401
