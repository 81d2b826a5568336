"""Import the UNMODIFIED reference Python (models + modules) from /root/reference on CPU.

TEST INFRASTRUCTURE, build-container only: /root/reference does not exist on the GPU box, so
nothing reachable from `-m gpu` tests, smoke() or bench.py may call this.  It exists to
generate tests/golden/* (oracle/make_golden.py) and to validate oracle/modules_ref.py.

How: the reference's native module `pointops_cuda` is replaced by the C oracle
(oracle.fake_pointops_cuda), `SharedArray` is stubbed (only imported by the seg data loader),
and the legacy constructors torch.cuda.{Int,Float,Long}Tensor that the reference hard-codes
(e.g. seg/modules/repsurface_utils.py:22,268; cls/po/functions/pointops.py:44-45) are aliased
to their CPU counterparts.  The two trees reuse the top-level package names `modules`,
`models`, `util`, so loading one tree evicts the other from sys.modules.
"""
import contextlib
import os
import sys
import types

import torch

from . import oracle as _orc

REF_ROOT = os.environ.get("REPSURF_REFERENCE", "/root/reference")
_TREES = {"cls": "classification", "seg": "segmentation"}
_SHARED = ("modules", "models", "util", "dataset", "pointops_cuda", "SharedArray")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "classification", "modules"))


def _evict():
    for name in list(sys.modules):
        if name.split(".")[0] in _SHARED:
            del sys.modules[name]


@contextlib.contextmanager
def _cpu_legacy_ctors():
    saved = {k: getattr(torch.cuda, k, None) for k in ("IntTensor", "FloatTensor", "LongTensor")}
    torch.cuda.IntTensor, torch.cuda.FloatTensor, torch.cuda.LongTensor = torch.IntTensor, torch.FloatTensor, torch.LongTensor
    try:
        yield
    finally:
        for k, v in saved.items():
            setattr(torch.cuda, k, v)


class RefTree:
    """Handle on one reference tree ('cls' or 'seg'); use as a context manager around every call
    into reference code so that sys.path / legacy constructors are in place."""

    def __init__(self, tree):
        assert available(), f"reference not found at {REF_ROOT}"
        self.tree = tree
        self.root = os.path.join(REF_ROOT, _TREES[tree])
        self._stack = None

    def __enter__(self):
        _evict()
        sys.path.insert(0, self.root)
        sys.modules["pointops_cuda"] = _orc.fake_pointops_cuda(self.tree)
        sys.modules["SharedArray"] = types.ModuleType("SharedArray")
        self._stack = contextlib.ExitStack()
        self._stack.enter_context(_cpu_legacy_ctors())
        return self

    def __exit__(self, *exc):
        self._stack.close()
        sys.path.remove(self.root)
        _evict()
        return False

    def imp(self, name):
        import importlib
        return importlib.import_module(name)


def cls_args(**kw):
    """Flags of classification/scripts/scanobjectnn/repsurf_ssg_umb.sh."""
    a = dict(return_center=True, return_polar=True, return_dist=True, group_size=8, umb_pool="sum",
             cuda_ops=True, num_point=1024, num_class=15)
    a.update(kw)
    return types.SimpleNamespace(**a)


def seg_args(**kw):
    """Flags of segmentation/scripts/s3dis/train_repsurf_umb.sh (+ constants set at seg/tool/train.py:452-470)."""
    a = dict(return_polar=False, in_channel=6, group_size=8, num_class=13)
    a.update(kw)
    return types.SimpleNamespace(**a)
