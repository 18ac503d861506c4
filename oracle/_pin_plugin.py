"""TEST INFRASTRUCTURE.  pytest plugin used by oracle/pin_against_reference.py.

PHX_PIN_TARGET=oracle    -> `import phantom` resolves to oracle.phantom_oracle
PHX_PIN_TARGET=reference -> `import phantom` resolves to the unmodified reference
                            (through the third-party stubs of oracle/ref_shim.py)
"""
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_repo = os.path.dirname(_here)
if _repo not in sys.path:
    sys.path.insert(0, _repo)

from oracle import ref_shim  # noqa: E402

target = os.environ.get("PHX_PIN_TARGET", "oracle")
if target == "oracle":
    ref_shim.install(with_reference=False)
    import oracle.phantom_oracle as po

    sys.modules["phantom"] = po
    prefix = po.__name__ + "."
    for name, mod in list(sys.modules.items()):
        if name.startswith(prefix):
            sys.modules["phantom." + name[len(prefix):]] = mod
else:
    ref_shim.import_reference()
