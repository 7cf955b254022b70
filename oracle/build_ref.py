"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Recipe that COMPILES the reference's own inference modules, from the sources
where they lie under /root/reference, into oracle/_ref/ (git-ignored; travels to the GPU box with the snapshot like
the repo's own built .so).

The reference is pure Python, so "compiling" is ``py_compile``: every module RAFTStereo.forward / IGEVStereo.forward
imports is translated to sourceless bytecode that keeps the package layout (``meta_arch/raft_stereo/raft_stereo.rpyc`` ...; the
extension is not ``.pyc`` because repository snapshots drop ``*.pyc`` files -- a small meta-path finder below imports
them).  No reference source text enters the repository or oracle/_ref/; the bytecode is the UNMODIFIED reference and is
what ``bench.py --impl reference`` times on the GPU box's host cores (``cpu_baseline.kind = "reference"``).  The GPU
box has the same image, hence the same CPython: the .pyc magic matches (``load()`` falls back to the oracle port and
says so if it does not).

    python -m oracle.build_ref            # run in the build container; __graft_entry__.build() calls it

``load()`` imports the staged modules (namespace stubs for ``meta_arch`` / ``opt_einsum`` / ``timm`` exactly as
oracle/make_golden.py does for the real sources).
"""
from __future__ import annotations

import importlib
import importlib.abc
import importlib.machinery
import importlib.util
import os
import py_compile
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("DKT_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
EXT = ".rpyc"

# modules on the import path of meta_arch.{raft_stereo.raft_stereo, igev_stereo.igev_stereo} (package __init__ files of
# meta_arch itself are NOT staged: they pull the training-side families and timm)
MODULES = [
    "core/__init__.py", "core/utils/__init__.py", "core/utils/utils.py",
    "meta_arch/raft_stereo/raft_stereo.py", "meta_arch/raft_stereo/update.py", "meta_arch/raft_stereo/extractor.py",
    "meta_arch/raft_stereo/corr.py", "meta_arch/raft_stereo/utils/__init__.py", "meta_arch/raft_stereo/utils/utils.py",
    "meta_arch/igev_stereo/igev_stereo.py", "meta_arch/igev_stereo/update.py", "meta_arch/igev_stereo/extractor.py",
    "meta_arch/igev_stereo/geometry.py", "meta_arch/igev_stereo/submodule.py",
    "meta_arch/igev_stereo/utils/__init__.py", "meta_arch/igev_stereo/utils/utils.py",
]
CONFIGS = ["configs/raft_stereo/base.json", "configs/igev_stereo/base.json"]


def build() -> int:
    """Compile MODULES into OUT; returns the number of files written (0 when the reference checkout is absent)."""
    if not os.path.isdir(REF):
        return 0
    n = 0
    for rel in MODULES:
        src = os.path.join(REF, rel)
        dst = os.path.join(OUT, rel[:-3] + EXT)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if os.path.exists(dst) and os.path.getmtime(dst) >= os.path.getmtime(src):
            continue
        # dfile: the path recorded in tracebacks -- the reference-relative name, so a failure cites the reference's lines
        py_compile.compile(src, cfile=dst, dfile=rel, doraise=True, optimize=0)
        n += 1
    return n


def available() -> bool:
    return os.path.exists(os.path.join(OUT, "meta_arch", "raft_stereo", "raft_stereo" + EXT))


class _RefFinder(importlib.abc.MetaPathFinder):
    """Maps a dotted module name onto oracle/_ref/<path>.rpyc (or <path>/__init__.rpyc) and loads it as bytecode."""

    def find_spec(self, fullname, path=None, target=None):
        rel = os.path.join(OUT, *fullname.split("."))
        init = os.path.join(rel, "__init__" + EXT)
        if os.path.exists(init):
            return importlib.util.spec_from_file_location(
                fullname, init, loader=importlib.machinery.SourcelessFileLoader(fullname, init), submodule_search_locations=[rel])
        if os.path.exists(rel + EXT):
            return importlib.util.spec_from_file_location(
                fullname, rel + EXT, loader=importlib.machinery.SourcelessFileLoader(fullname, rel + EXT))
        return None


def _stub(name: str, path=None, **attrs):
    if name not in sys.modules:
        m = types.ModuleType(name)
        if path is not None:
            m.__path__ = [path]
        sys.modules[name] = m
    for k, v in attrs.items():
        setattr(sys.modules[name], k, v)
    return sys.modules[name]


def load(model: str = "raft"):
    """-> the reference's model CLASS (RAFTStereo or IGEVStereo) imported from the staged bytecode."""
    import torch
    if not available():
        raise ImportError("oracle/_ref is empty: run `python -m oracle.build_ref` where /root/reference exists")
    sys.dont_write_bytecode = True
    if not any(isinstance(f, _RefFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _RefFinder())
    _stub("meta_arch", os.path.join(OUT, "meta_arch"))
    _stub("meta_arch.raft_stereo", os.path.join(OUT, "meta_arch", "raft_stereo"))
    _stub("meta_arch.igev_stereo", os.path.join(OUT, "meta_arch", "igev_stereo"))
    _stub("opt_einsum", contract=torch.einsum)
    if model == "raft":
        return importlib.import_module("meta_arch.raft_stereo.raft_stereo").RAFTStereo
    _stub("timm")
    from oracle.make_golden import _timm_stub      # torchvision's MobileNetV2 in place of timm 0.5.4
    _timm_stub()
    return importlib.import_module("meta_arch.igev_stereo.igev_stereo").IGEVStereo


if __name__ == "__main__":
    print(f"[build_ref] {build()} module(s) compiled into {OUT}")
