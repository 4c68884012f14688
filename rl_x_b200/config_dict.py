"""Minimal stand-in for `ml_collections.config_dict.ConfigDict` (absent from this image): an attribute dict with the
subset of behaviour the reference relies on (runner.py:179-181,266-270; every default_config.py): attribute get/set,
`in`, `.items()`, nested trees, and `--tree.key=value` overrides applied with the type of the default."""


class ConfigDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def __delattr__(self, k):
        del self[k]

    def to_dict(self):
        return {k: (v.to_dict() if isinstance(v, ConfigDict) else v) for k, v in self.items()}

    def copy_and_resolve_references(self):
        return ConfigDict({k: (v.copy_and_resolve_references() if isinstance(v, ConfigDict) else v) for k, v in self.items()})


def config_tree_class():
    """The class default configs are built from: `ml_collections.config_dict.ConfigDict` where ml_collections is installed - the reference's
    runner hands the plugins' defaults to `config_flags.DEFINE_config_dict` (runner.py:179-181), which only accepts that class - and the
    stand-in above where it is not (this image)."""
    try:
        from ml_collections import config_dict as _mlc
        return _mlc.ConfigDict
    except ImportError:
        return ConfigDict


def config_from_defaults(name, defaults):
    """A plugin's default config tree from its (key, value) table; `name` is the registered plugin name."""
    config = config_tree_class()()
    config["name"] = name
    for key, value in defaults:
        config[key] = value
    return config


def _coerce(text, default):
    if isinstance(default, bool):
        if text.lower() in ("true", "1", "yes"):
            return True
        if text.lower() in ("false", "0", "no"):
            return False
        raise ValueError(f"cannot parse {text!r} as bool")
    if isinstance(default, int) and not isinstance(default, bool):
        try:
            return int(text)
        except ValueError:
            return type(default)(float(text))
    if isinstance(default, float):
        return float(text)
    if default is None:
        for cast in (int, float):
            try:
                return cast(text)
            except ValueError:
                pass
        return None if text.lower() == "none" else text
    return type(default)(text)


def apply_overrides(trees, argv):
    """Apply `--runner.x=1 --algorithm.y=2 --environment.z=3` style flags (absl config_flags semantics: the key must exist,
    the value takes the default's type).  Returns the list of explicitly set dotted keys (runner.py:266-270)."""
    explicitly_set = []
    for arg in argv:
        if not arg.startswith("--"):
            raise ValueError(f"unrecognised argument {arg!r}")
        body = arg[2:]
        if "=" not in body:
            raise ValueError(f"flag {arg!r} needs a value (--tree.key=value)")
        dotted, text = body.split("=", 1)
        parts = dotted.split(".")
        if parts[0] not in trees:
            raise ValueError(f"unknown config tree in {arg!r}")
        node = trees[parts[0]]
        for p in parts[1:-1]:
            if p not in node:
                raise KeyError(f"unknown config key {dotted!r}")
            node = node[p]
        leaf = parts[-1]
        if leaf not in node:
            raise KeyError(f"unknown config key {dotted!r}")
        node[leaf] = _coerce(text, node[leaf])
        explicitly_set.append(dotted)
    return explicitly_set
