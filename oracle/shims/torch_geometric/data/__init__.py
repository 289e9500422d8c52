class Data:
    """Attribute bag (src/utils/graph.py:15-24 subclasses it and calls super(DenseData).__init__)."""

    def __init__(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)

    def to(self, device):
        import torch
        for k, v in list(self.__dict__.items()):
            if torch.is_tensor(v):
                setattr(self, k, v.to(device))
        return self

    @property
    def num_features(self):
        return self.x.size(1)

    @property
    def num_nodes(self):
        return self.x.size(0)
