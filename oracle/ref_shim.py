"""TEST INFRASTRUCTURE -- never imported by the product (phantom_b200/).

Import shim that lets the *unmodified* reference package under /root/reference be
imported in the build container, where four of its third-party imports are absent
(gymnasium, termcolor, tensorboardX, ray; matplotlib for the supply-chain example).
None of the env-step hot path lives in those packages:
gymnasium contributes `gym.Env` as a base class and `gym.spaces.*` as declarations.

The shim is only usable where /root/reference exists (this container).  It is used by
  * oracle/make_golden.py      -- generates tests/golden/*.npz from the real reference
  * oracle/pin_against_reference.py -- runs the reference's own test-suite against the
                                  oracle restatement, and the reference against itself
Nothing under tests/ -m gpu, bench.py or __graft_entry__.smoke() uses it.
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("PHX_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "phantom"))


from .phantom_oracle.spaces import Box, Dict, Discrete, Env, Space, Tuple  # noqa: E402


def _make_gymnasium() -> types.ModuleType:
    gym = types.ModuleType("gymnasium")
    spaces = types.ModuleType("gymnasium.spaces")
    for cls in (Space, Box, Discrete, Dict, Tuple):
        setattr(spaces, cls.__name__, cls)
    gym.spaces = spaces
    gym.Space = Space
    gym.Env = Env
    gym.__path__ = []  # mark as package
    return gym, spaces


# ---------------------------------------------------------- fabricated ray.* modules
class _FabricatedModule(types.ModuleType):
    """Module whose every attribute exists: Capitalised names are empty classes,
    ALL_CAPS names are strings, anything else is a sub-module."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        if name.isupper():
            value = "/tmp/_phx_fabricated"
        elif name[0].isupper():
            value = type(name, (), {"__init__": lambda self, *a, **k: None})
        else:
            value = _FabricatedModule(f"{self.__name__}.{name}")
            value.__path__ = []
            sys.modules[value.__name__] = value
        setattr(self, name, value)
        return value


class _FabricatingFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    ROOTS = ("ray",)

    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in self.ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        mod = _FabricatedModule(spec.name)
        mod.__path__ = []
        return mod

    def exec_module(self, module):
        pass


_installed = False


def install(with_reference: bool = True) -> None:
    """Install the stub third-party modules; optionally put the reference on sys.path."""
    global _installed
    if _installed:
        return
    _installed = True

    if "gymnasium" not in sys.modules:
        try:
            import gymnasium  # noqa: F401
        except ImportError:
            gym, spaces = _make_gymnasium()
            sys.modules["gymnasium"] = gym
            sys.modules["gymnasium.spaces"] = spaces

    if "termcolor" not in sys.modules:
        try:
            import termcolor  # noqa: F401
        except ImportError:
            tc = types.ModuleType("termcolor")
            tc.colored = lambda text, *a, **k: text
            sys.modules["termcolor"] = tc

    if "tensorboardX" not in sys.modules:
        try:
            import tensorboardX  # noqa: F401
        except ImportError:
            tbx = types.ModuleType("tensorboardX")
            tbx.SummaryWriter = type("SummaryWriter", (), {})
            sys.modules["tensorboardX"] = tbx

    try:
        import matplotlib.pyplot  # noqa: F401
    except ImportError:
        mpl = types.ModuleType("matplotlib")
        mpl.__path__ = []
        mpl.pyplot = types.ModuleType("matplotlib.pyplot")
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = mpl.pyplot

    try:
        import ray  # noqa: F401
    except ImportError:
        sys.meta_path.append(_FabricatingFinder())

    if with_reference:
        if not reference_available():
            raise RuntimeError(
                f"reference tree not found at {REFERENCE_ROOT}; the shim only works "
                "in the build container"
            )
        if REFERENCE_ROOT not in sys.path:
            sys.path.insert(0, REFERENCE_ROOT)


def import_reference():
    """Return the real `phantom` package imported from /root/reference."""
    install(with_reference=True)
    import phantom

    assert phantom.__file__.startswith(REFERENCE_ROOT), phantom.__file__
    return phantom


def import_reference_supply_chain():
    """Import examples/environments/supply_chain/supply_chain.py unmodified.

    The file runs `sys.argv[1]` dispatch at module scope (supply_chain.py:185), so a
    neutral argv is supplied while importing."""
    import importlib.util

    import_reference()
    path = os.path.join(
        REFERENCE_ROOT, "examples/environments/supply_chain/supply_chain.py"
    )
    spec = importlib.util.spec_from_file_location("_ref_supply_chain", path)
    mod = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ["supply_chain.py", "none"]
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.argv = argv
    return mod
