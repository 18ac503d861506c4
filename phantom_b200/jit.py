"""Run-time specialisation of the step kernel to one env class (include/phx.h: phx_jit_source /
phx_load_specialised).

The generic queue engines interpret the lowered env class at run time -- agent loops, kind and
mask tests, per-slot tables are all data driven -- which costs an order of magnitude in
instructions for small env classes.  `specialise(env)` asks libphx for a translation unit in which
this handle's lowered env class is a compile-time constant, compiles it for sm_100a with nvcc
(`-cubin`), caches the cubin under `phantom_b200/_jit/` keyed by the hash of everything that
went into it, and hands it back to the library; the handle's phx_step / phx_rollout then launch
the specialised kernel.  Results are bit-identical (same kernel body, constant-folded):
tests/test_gpu_jit.py.  Opt-in (`env.specialise()`), because a compile takes a few seconds and
the constant includes the handle's num_envs / seed / env_offset.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import shutil
import subprocess
import tempfile
from typing import Optional

from . import _lib as L

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
CACHE = os.environ.get("PHX_JIT_CACHE", os.path.join(HERE, "_jit"))
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc() -> str:
    exe = os.environ.get("PHX_NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("run-time specialisation needs nvcc (set PHX_NVCC)")
    return exe


def _tree_digest() -> bytes:
    """Hash of every source the generated unit can include."""
    h = hashlib.sha256()
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))
             if f.endswith((".cu", ".cuh", ".h"))]
    files.append(os.path.join(os.path.dirname(HERE), "include", "phx.h"))
    for path in files:
        with open(path, "rb") as f:
            h.update(path.encode() + b"\0" + f.read())
    return h.digest()


def source(env) -> str:
    """The specialised translation unit of a live env (text)."""
    env._ensure_handle()
    need = C.c_uint64(0)
    L.check(L.lib.phx_jit_source(env._handle, None, 0, C.byref(need)))
    buf = C.create_string_buffer(need.value)
    L.check(L.lib.phx_jit_source(env._handle, buf, need.value, None))
    return buf.value.decode()


def compile_program(path: str, cache: Optional[str] = None) -> str:
    """A user's device program (a .cu file ending with PHX_USER_PROGRAM(...), csrc/phx_user.cuh)
    -> path of its cubin.  Files it includes with "..." are looked up next to it first."""
    path = os.path.abspath(path)
    with open(path) as f:
        text = f.read()
    return compile_unit(text, cache, include_dirs=[os.path.dirname(path)])


def compile_unit(text: str, cache: Optional[str] = None, include_dirs=()) -> str:
    """text -> path of the cubin (compiled once per distinct text / source tree / nvcc)."""
    cache = cache or CACHE
    try:
        os.makedirs(cache, exist_ok=True)
        if not os.access(cache, os.W_OK):
            raise OSError("not writable")
    except OSError:  # read-only install: fall back to a per-user scratch directory
        cache = os.path.join(tempfile.gettempdir(), f"phx_jit_{os.getuid()}")
        os.makedirs(cache, exist_ok=True)
    nvcc = _nvcc()
    key = hashlib.sha256(text.encode() + _tree_digest() + nvcc.encode()).hexdigest()[:24]
    cubin = os.path.join(cache, key + ".cubin")
    if os.path.exists(cubin):
        return cubin
    with tempfile.TemporaryDirectory(dir=cache) as tmp:
        unit = os.path.join(tmp, "unit.cu")
        with open(unit, "w") as f:
            f.write(text)
        out = os.path.join(tmp, "unit.cubin")
        cmd = [nvcc, *ARCH, "-O3", "-std=c++17", "-lineinfo", "-cubin", "-I", CSRC]
        for d in include_dirs:
            cmd += ["-I", d]
        cmd += ["-o", out, unit]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode != 0:
            raise RuntimeError("nvcc failed on the specialised unit:\n" + proc.stderr[-4000:])
        os.replace(out, cubin)  # atomic: concurrent ranks may compile the same unit
        with open(os.path.join(cache, key + ".cu"), "w") as f:
            f.write(text)
    return cubin


def specialise(env, cache: Optional[str] = None) -> str:
    """Compile (or fetch) and load the build of the step kernel specialised to `env`'s env class.
    Returns the cubin path.  Raises NotLowerableError-like RuntimeError (from libphx) when the env
    class / kernel variant has no specialisation."""
    cubin = compile_unit(source(env), cache)
    L.check(L.lib.phx_load_specialised(env._handle, cubin.encode()))
    return cubin
