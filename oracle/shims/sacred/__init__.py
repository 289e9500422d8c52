"""Stand-in for `sacred` (absent, not installable offline). TEST INFRASTRUCTURE ONLY.

Provides just enough of `Ingredient.config` / `Ingredient.capture` / `Experiment` for the
reference's modules to import and for captured functions to receive their config defaults
(reference call sites: src/models/sampling.py:89-106, src/models/factory.py:11-19,
src/trainers/outer.py:113-133). Written from sacred's documented behaviour.
"""
import functools
import inspect
import sys


class Ingredient:
    def __init__(self, path="", ingredients=()):
        self.path = path
        self.ingredients = list(ingredients)
        self.cfg = {}

    def config(self, fn):
        # sacred runs the config function and harvests its local variables as defaults
        harvested = {}

        def profiler(frame, event, arg):
            if event == "return" and frame.f_code is fn.__code__:
                harvested.update(frame.f_locals)

        old = sys.getprofile()
        sys.setprofile(profiler)
        try:
            fn()
        finally:
            sys.setprofile(old)
        self.cfg.update(harvested)
        return fn

    def capture(self, fn):
        sig = inspect.signature(fn)
        cfg = self.cfg

        @functools.wraps(fn)
        def wrapper(*args, **kwargs):
            bound = sig.bind_partial(*args, **kwargs)
            for name in sig.parameters:
                if name not in bound.arguments and name in cfg:
                    kwargs[name] = cfg[name]
            return fn(*args, **kwargs)

        return wrapper

    def add_config(self, *a, **kw):
        for d in a:
            self.cfg.update(d)
        self.cfg.update(kw)


class Experiment(Ingredient):
    def __init__(self, name="", ingredients=(), **kw):
        super().__init__(name, ingredients)
        self.observers = []
        self.logger = None

    def automain(self, fn):
        return fn

    def main(self, fn):
        return fn

    def post_run_hook(self, fn):
        return fn
