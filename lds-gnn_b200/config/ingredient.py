"""A small capture-style injector with the two sacred decorators the LDS path relies on."""
import functools
import inspect
import sys

REGISTRY = {}


class Ingredient:
    """`Ingredient(name)`: `.config(fn)` runs `fn` once and keeps its local variables as defaults;
    `.capture(fn)` fills arguments the caller omitted from those values (caller arguments win)."""

    def __init__(self, path):
        self.path = path
        self.values = {}
        REGISTRY[path] = self

    def config(self, fn):
        found = {}

        def tracer(frame, event, arg):
            if event == "return" and frame.f_code is fn.__code__:
                found.update(frame.f_locals)

        previous = sys.getprofile()
        sys.setprofile(tracer)
        try:
            fn()
        finally:
            sys.setprofile(previous)
        self.values.update(found)
        return fn

    def capture(self, fn):
        signature = inspect.signature(fn)

        @functools.wraps(fn)
        def injected(*args, **kwargs):
            given = signature.bind_partial(*args, **kwargs).arguments
            for name in signature.parameters:
                if name not in given and name in self.values:
                    kwargs[name] = self.values[name]
            return fn(*args, **kwargs)

        return injected

    def update(self, mapping):
        unknown = set(mapping) - set(self.values)
        if unknown:
            raise KeyError(f"ingredient '{self.path}' has no config entries {sorted(unknown)}")
        self.values.update(mapping)


def apply_config(cfg):
    """Push a parsed config dict ({ingredient: {key: value}, top_level_key: value}) into the registered
    ingredients; returns the remaining top-level entries (the `run(...)` kwargs of src/scripts/bilevel.py:39-52)."""
    rest = {}
    for key, value in cfg.items():
        if isinstance(value, dict) and key in REGISTRY:
            REGISTRY[key].update(value)
        else:
            rest[key] = value
    return rest


def current_config():
    return {name: dict(ing.values) for name, ing in REGISTRY.items()}
