"""One registry implementation behind both plugin managers.  The contract is the reference's (rl_x/algorithms/algorithm_manager.py:8-25,
rl_x/environments/environment_manager.py:8-25): a plugin package's `__init__.py` derives its dotted name from its own path and registers
three objects under it; the runner looks them up by that name."""
import os
from collections import namedtuple


class PluginRegistry:
    def __init__(self, kind, fields):
        self.kind = kind                      # directory name that anchors the dotted plugin name ("algorithms" / "environments")
        self.Entry = namedtuple(f"{kind.capitalize()}Entry", ("name",) + tuple(fields))
        self.entries = {}

    def name_from_file(self, init_file):
        """'.../<kind>/ppo/b200/__init__.py' -> 'ppo.b200'"""
        parts = os.path.normpath(init_file).split(os.sep)
        start = len(parts) - 1 - parts[::-1].index(self.kind)
        return ".".join(parts[start + 1:-1])

    def register(self, name, *objects):
        self.entries[name] = self.Entry(name, *objects)
        self._mirror_into_reference(name, objects)

    def _mirror_into_reference(self, name, objects):
        """Where the reference package is importable, the plugin also goes into ITS registry.  The reference's runner imports
        `<package>.<kind>.<name>` for every implementation package it is given and then looks the plugin up in rl_x's own manager
        (rl_x/runner/runner.py:232-247, 86-96), so `Runner(implementation_package_names=["rl_x", "rl_x_b200"])` finds `ppo.b200` without
        any bridge package.  Same argument order on both sides (rl_x/algorithms/algorithm_manager.py:12, environment_manager.py:12)."""
        import importlib
        singular = self.kind[:-1]
        try:
            manager = importlib.import_module(f"rl_x.{self.kind}.{singular}_manager")
        except ImportError:
            return                       # no reference here: this package's own runner and registry are all there is
        except Exception as exc:         # a reference install that does not import must not take this package's plugins down with it
            import logging
            logging.getLogger("rl_x").warning(f"rl_x_b200: {name!r} not mirrored into rl_x's {singular} registry ({type(exc).__name__}: {exc})")
            return
        getattr(manager, f"register_{singular}")(name, *objects)

    def lookup(self, name):
        try:
            return self.entries[name]
        except KeyError:
            raise KeyError(f"no {self.kind[:-1]} plugin registered under {name!r} (registered: {sorted(self.entries)})") from None
