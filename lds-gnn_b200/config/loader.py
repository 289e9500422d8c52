"""Readers for the reference's two launch formats (only what the LDS files use)."""
import itertools
import json
import random

import yaml

SEML_RESERVED = ("seml", "slurm", "fixed", "grid", "random")


def load_sacred_json(path):
    """configs/sacred/lds/config.json: nested dict, one sub-dict per ingredient."""
    with open(path) as fh:
        return json.load(fh)


def _set(cfg, dotted, value):
    parts = dotted.split(".")
    for p in parts[:-1]:
        cfg = cfg.setdefault(p, {})
    cfg[parts[-1]] = value


def _flatten_fixed(block, prefix=""):
    out = {}
    for k, v in (block or {}).items():
        name = f"{prefix}{k}"
        if isinstance(v, dict):
            if v.get("type") == "parameter_collection":          # lds_grid.yaml nests one under `fixed:`
                v = v.get("params", {})
            out.update(_flatten_fixed(v, name + "."))
        else:
            out[name] = v
    return out


def _grid_axes(block, prefix=""):
    """seml `grid:` -> list of (dotted_name, [options]); `parameter_collection` nests, `choice` lists options."""
    axes = []
    for k, v in (block or {}).items():
        name = f"{prefix}{k}"
        kind = v.get("type")
        if kind == "parameter_collection":
            axes += _grid_axes(v.get("params", {}), name + ".")
        elif kind == "choice":
            axes.append((name, list(v["options"])))
        elif kind == "range":
            lo, hi, step = v["min"], v["max"], v.get("step", 1)
            opts, x = [], lo
            while x < hi:
                opts.append(x)
                x += step
            axes.append((name, opts))
        else:
            raise ValueError(f"unsupported seml grid type '{kind}' for '{name}'")
    return axes


def _random_axes(block, prefix=""):
    axes = []
    for k, v in (block or {}).items():
        if k in ("samples", "seed"):
            continue
        name = f"{prefix}{k}"
        kind = v.get("type")
        if kind == "parameter_collection":
            axes += _random_axes(v.get("params", {}), name + ".")
        elif kind == "randint":
            axes.append((name, ("randint", v["min"], v["max"])))
        elif kind == "uniform":
            axes.append((name, ("uniform", v["min"], v["max"])))
        elif kind == "choice":
            axes.append((name, ("choice", list(v["options"]))))
        else:
            raise ValueError(f"unsupported seml random type '{kind}' for '{name}'")
    return axes


def load_seml_yaml(path):
    """Parse a seml YAML into {"seml":…, "slurm":…, "fixed":…, "grid":…, "random":…, "sub": {name: {...}}}."""
    with open(path) as fh:
        doc = yaml.safe_load(fh)
    out = {k: doc.get(k) for k in SEML_RESERVED}
    out["sub"] = {k: v for k, v in doc.items() if k not in SEML_RESERVED}
    return out


def expand_seml(doc, sub_config=None):
    """All concrete configurations of a parsed seml file (root blocks + optionally one named sub-config),
    as nested dicts in the sacred-JSON shape. Grid axes are crossed; `random:` draws `samples` points."""
    blocks = [doc] + ([doc["sub"][sub_config]] if sub_config else [])
    fixed, axes, rand_axes, samples, seed = {}, [], [], 1, None
    for b in blocks:
        fixed.update(_flatten_fixed(b.get("fixed")))
        axes += _grid_axes(b.get("grid"))
        rnd = b.get("random") or {}
        if rnd:
            samples, seed = rnd.get("samples", samples), rnd.get("seed", seed)
            rand_axes += _random_axes(rnd)
    rng = random.Random(seed)
    configs = []
    names = [a[0] for a in axes]
    for combo in itertools.product(*[a[1] for a in axes]) if axes else [()]:
        for _ in range(samples if rand_axes else 1):
            cfg = {}
            for k, v in fixed.items():
                _set(cfg, k, v)
            for k, v in zip(names, combo):
                _set(cfg, k, v)
            for k, spec in rand_axes:
                if spec[0] == "randint":
                    _set(cfg, k, rng.randrange(spec[1], spec[2]))
                elif spec[0] == "uniform":
                    _set(cfg, k, rng.uniform(spec[1], spec[2]))
                else:
                    _set(cfg, k, rng.choice(spec[1]))
            configs.append(cfg)
    return configs
