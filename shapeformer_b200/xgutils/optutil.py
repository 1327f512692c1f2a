"""YAML option loader with recursive `inherit_from` (reference xgutils/optutil.py:44-70)."""
import os

import yaml

from .sysutil import dictUpdate


def load_option(path):
    with open(path) as f:
        opt = yaml.safe_load(f) or {}
    parent = opt.pop("inherit_from", None)
    if parent:
        base = load_option(os.path.normpath(os.path.join(os.path.dirname(path), parent)))
        opt = dictUpdate(base, opt)
    return opt
