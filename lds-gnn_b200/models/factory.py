"""Factory for the LDS graph model and its optimiser (src/models/factory.py:9-69). Only the `lds` ingredient
is on the path; the `embedding` / `gae` model families are out of scope."""
from torch.optim import SGD, Adam, Optimizer

from ..config import Ingredient
from ..utils.graph import DenseData
from .graph import BernoulliGraphModel, GraphGenerativeModel


class GraphGenerativeModelFactory:
    _lds_ingredient = Ingredient("lds")
    INGREDIENTS = {"lds": _lds_ingredient}

    def __init__(self, data: DenseData):
        self.data = data

    def create(self, model_name: str) -> GraphGenerativeModel:
        if model_name.lower() == "lds":
            return self.lds(data=self.data)
        raise NotImplementedError(f"Model {model_name} not supported (lds_gnn_b200 covers the LDS model).")

    def optimizer(self, model: GraphGenerativeModel) -> Optimizer:
        if type(model) == BernoulliGraphModel:
            return self.lds_optimizer(model=model)
        raise NotImplementedError(f"Optimizer for model type {type(model)} not implemented.")

    @staticmethod
    @_lds_ingredient.config
    def _lds_config():
        directed: bool = False        # noqa: F841
        lr: float = 1.0               # noqa: F841
        optimizer_type: str = "SGD"   # noqa: F841  (the reference hard-codes SGD; Adam is the north-star's option)

    @staticmethod
    @_lds_ingredient.capture
    def lds(data: DenseData, directed: bool) -> BernoulliGraphModel:
        return BernoulliGraphModel(data.dense_adj, directed=directed)

    @staticmethod
    @_lds_ingredient.capture
    def lds_optimizer(model: BernoulliGraphModel, lr: float, optimizer_type: str = "SGD") -> Optimizer:
        if optimizer_type == "SGD":
            return SGD(model.parameters(), lr=lr)
        if optimizer_type == "Adam":
            return Adam(model.parameters(), lr=lr)
        raise NotImplementedError(f"Optimizer {optimizer_type} not supported")
