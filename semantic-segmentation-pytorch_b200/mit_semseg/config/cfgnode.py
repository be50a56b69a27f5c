"""A small work-alike of `yacs.config.CfgNode` for the subset the reference's scripts use (config/defaults.py,
train.py:227-268, eval.py, test.py): attribute access on a nested dict, `merge_from_file` (YAML), `merge_from_list`
(KEY VALUE pairs with dotted keys), literal evaluation of strings such as "(300, 375, 450)" or "1e-4", type checking of
overrides against the defaults, `freeze` / `defrost` / `clone`, YAML text from `str()`. New keys may be ASSIGNED at any time
(the scripts add TRAIN.batch_size, TRAIN.max_iters, ... on the fly) but a MERGE may only override keys that exist.
Used when yacs itself is not installed (mit_semseg/config/defaults.py picks whichever is available)."""
import ast
import copy

import yaml


def _literal(v):
    """YAML / command-line strings -> Python literals where they parse ("1e-4" -> 0.0001, "(1, 2)" -> (1, 2))"""
    if isinstance(v, str):
        try:
            return ast.literal_eval(v)
        except (ValueError, SyntaxError):
            return v
    return v


def _coerce(new, old, key):
    """the override must have the default's type; tuple <-> list and int -> float are converted like yacs does"""
    if old is None or type(new) is type(old):
        return new
    if isinstance(old, tuple) and isinstance(new, list):
        return tuple(new)
    if isinstance(old, list) and isinstance(new, tuple):
        return list(new)
    if isinstance(old, float) and isinstance(new, int) and not isinstance(new, bool):
        return float(new)
    raise ValueError("Type mismatch ({} vs. {}) with values ({} vs. {}) for config key: {}".format(
        type(old), type(new), old, new, key))


class CfgNode(dict):
    _FROZEN = "__frozen__"

    def __init__(self, init=None):
        super().__init__()
        self.__dict__[CfgNode._FROZEN] = False
        for k, v in (init or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    # ---- attribute access
    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        if self.__dict__[CfgNode._FROZEN]:
            raise AttributeError("Attempted to set {} to {}, but CfgNode is immutable".format(name, value))
        self[name] = value

    # ---- merging
    def _merge_dict(self, other, path):
        for k, v in other.items():
            full = ".".join(path + [k])
            if k not in self:
                raise KeyError("Non-existent config key: {}".format(full))
            if isinstance(self[k], CfgNode):
                if not isinstance(v, dict):
                    raise ValueError("config key {} is a section".format(full))
                self[k]._merge_dict(v, path + [k])
            else:
                self[k] = _coerce(_literal(copy.deepcopy(v)), self[k], full)

    def merge_from_other_cfg(self, other):
        self._merge_dict(other, [])

    def merge_from_file(self, cfg_filename):
        with open(cfg_filename, "r") as f:
            self._merge_dict(yaml.safe_load(f) or {}, [])

    def merge_from_list(self, cfg_list):
        cfg_list = list(cfg_list or [])
        if len(cfg_list) % 2:
            raise AssertionError("Override list has odd length: {}; it must be a list of pairs".format(cfg_list))
        for full, v in zip(cfg_list[0::2], cfg_list[1::2]):
            node, parts = self, full.split(".")
            for p in parts[:-1]:
                if p not in node:
                    raise KeyError("Non-existent config key: {}".format(full))
                node = node[p]
            if parts[-1] not in node:
                raise KeyError("Non-existent config key: {}".format(full))
            node[parts[-1]] = _coerce(_literal(v), node[parts[-1]], full)

    # ---- state
    def _set_frozen(self, flag):
        self.__dict__[CfgNode._FROZEN] = flag
        for v in self.values():
            if isinstance(v, CfgNode):
                v._set_frozen(flag)

    def freeze(self):
        self._set_frozen(True)

    def defrost(self):
        self._set_frozen(False)

    def is_frozen(self):
        return self.__dict__[CfgNode._FROZEN]

    def clone(self):
        return copy.deepcopy(self)

    def __deepcopy__(self, memo):
        out = CfgNode()
        for k, v in self.items():
            dict.__setitem__(out, k, copy.deepcopy(v, memo))
        out.__dict__[CfgNode._FROZEN] = self.__dict__[CfgNode._FROZEN]
        return out

    # ---- text
    def _plain(self):
        def conv(v):
            if isinstance(v, CfgNode):
                return v._plain()
            if isinstance(v, tuple):
                return list(v)
            return v
        return {k: conv(v) for k, v in self.items()}

    def dump(self, **kwargs):
        return yaml.safe_dump(self._plain(), **kwargs)

    def __str__(self):
        return self.dump(default_flow_style=None, sort_keys=True).rstrip("\n")

    def __repr__(self):
        return "{}({})".format(self.__class__.__name__, dict.__repr__(self))
