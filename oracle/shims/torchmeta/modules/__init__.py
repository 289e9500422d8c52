"""MetaModule / MetaLinear restated from torchmeta 1.2.1's documented behaviour:
MetaLinear.forward(x, params) = F.linear(x, params['weight'], params.get('bias')), with
`params` defaulting to the module's own parameters (call sites: src/models/layers.py:35,43)."""
from collections import OrderedDict

import torch.nn as nn
import torch.nn.functional as F


class MetaModule(nn.Module):
    def meta_named_parameters(self, prefix="", recurse=True):
        return self.named_parameters(prefix=prefix, recurse=recurse)

    def meta_parameters(self, recurse=True):
        return self.parameters(recurse=recurse)


class MetaLinear(nn.Linear, MetaModule):
    def forward(self, input, params=None):
        if params is None:
            params = OrderedDict(self.named_parameters())
        bias = params.get("bias", None)
        return F.linear(input, params["weight"], bias)
