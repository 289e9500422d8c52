"""Config plumbing of the drop-in: sacred-style ingredients and loaders for the reference's LDS config files.

The reference wires its hyper-parameters through `sacred.Ingredient`s (`@config` defaults, `@capture`
injection; e.g. src/models/sampling.py:89-106, src/models/factory.py:52-69, src/trainers/outer.py:113-133)
and is launched from sacred JSON (configs/sacred/lds/config.json) or seml YAML
(configs/seml/final/lds.yaml, configs/seml/grid/lds_grid.yaml). sacred/seml are not dependencies here:
`Ingredient` below reproduces the two decorators the path uses, and `loader` reads both file formats.
"""
from .ingredient import Ingredient, REGISTRY, apply_config, current_config      # noqa: F401
from .loader import load_sacred_json, load_seml_yaml, expand_seml               # noqa: F401
