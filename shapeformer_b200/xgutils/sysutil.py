"""Plugin loader: `{class: dotted.path, kwargs: {...}}` -> object (reference xgutils/sysutil.py:136-156)."""
import importlib


def load_object(object_path):
    module_path, _, name = object_path.rpartition(".")
    mod = importlib.import_module(module_path)
    if not hasattr(mod, name):
        raise NameError(f"Object {name} not found in {module_path}")
    return getattr(mod, name)


def instantiate_from_opt(opt):
    if opt is None or opt.get("class") is None:
        return None
    return load_object(opt["class"])(**opt.get("kwargs", dict()))


def dictUpdate(base, new):
    """Recursive dict merge (reference xgutils/sysutil.py:46-64): values of `new` override `base`."""
    out = dict(base)
    for k, v in new.items():
        if isinstance(v, dict) and isinstance(out.get(k), dict):
            out[k] = dictUpdate(out[k], v)
        else:
            out[k] = v
    return out


def progbar(it, **kw):
    return it
