"""Run an UNMODIFIED reference script (main_cd.py, eval_cd.py, demo.py) on the native module.

    python -m dahitra_b200.launch --ref /path/to/DAHiTra eval_cd.py --net_G newUNetTrans --gpu_ids 0 ...

The reference harness resolves the class at call time inside ``define_G``
(``net = BASE_Transformer_UNet(...)``, reference models/networks.py:163-165), so rebinding
``models.networks.BASE_Transformer_UNet`` before the script runs swaps the implementation without editing a
single reference file.  ``define_G``/``init_net``/``init_weights`` stay the reference's own.

``--stub-missing`` registers empty stand-ins for optional third-party imports the reference pulls in at module
import time but never touches on this path (timm via ChangeFormer, segmentation_models_pytorch via losses),
and keeps ``torchvision``-style ``pretrained=True`` from downloading when no network is available.
"""
from __future__ import annotations

import argparse
import importlib
import os
import runpy
import sys
import types


def _stub(name, **attrs):
    parts = name.split(".")
    for i in range(1, len(parts) + 1):
        n = ".".join(parts[:i])
        if n not in sys.modules:
            sys.modules[n] = types.ModuleType(n)
        if i > 1:
            setattr(sys.modules[".".join(parts[:i - 1])], parts[i - 1], sys.modules[n])
    for k, v in attrs.items():
        setattr(sys.modules[name], k, v)


def install(ref_root: str, stub_missing: bool = False, offline_trunk: bool = False, rebind: bool = True):
    """Import the reference's ``models.networks`` from ``ref_root`` and rebind the network class (``rebind=False``
    leaves the reference's own class in place: only the import shims are applied — used by the reference arm of
    bench.py and the harness tests).  Returns the reference ``models.networks`` module."""
    import torch
    ref_root = os.path.abspath(ref_root)
    if ref_root not in sys.path:
        sys.path.insert(0, ref_root)
    if stub_missing:
        try:
            importlib.import_module("timm.models.layers")
        except Exception:
            class DropPath(torch.nn.Identity):
                def __init__(self, *a, **k):
                    super().__init__()
            _stub("timm.models.layers", DropPath=DropPath, to_2tuple=lambda x: (x, x),
                  trunc_normal_=torch.nn.init.trunc_normal_)
        for mod, attrs in (("segmentation_models_pytorch.losses", dict(DiceLoss=object)), ("tifffile", {}),
                           ("matplotlib.pyplot", dict(imsave=lambda *a, **k: None)),
                           ("skimage.filters", dict(threshold_otsu=lambda x: 0.5))):
            try:
                importlib.import_module(mod)
            except Exception:
                _stub(mod, **attrs)
        # the reference has a top-level `datasets/` directory without __init__.py; make sure it wins over an
        # installed `datasets` distribution
        if os.path.isdir(os.path.join(ref_root, "datasets")):
            ns = types.ModuleType("datasets")
            ns.__path__ = [os.path.join(ref_root, "datasets")]
            sys.modules["datasets"] = ns
    nets = importlib.import_module("models.networks")
    if offline_trunk:
        res = importlib.import_module("models.resnet")
        res._resnet = lambda arch, block, layers, pretrained, progress, **kw: res.ResNet(block, layers, **kw)
    if rebind:
        from dahitra_b200.networks import BASE_Transformer_UNet
        nets.BASE_Transformer_UNet = BASE_Transformer_UNet
    return nets


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--ref", required=True, help="root of the nka77/DAHiTra checkout")
    ap.add_argument("--stub-missing", action="store_true")
    ap.add_argument("script", help="reference script relative to --ref, e.g. eval_cd.py")
    ap.add_argument("script_args", nargs=argparse.REMAINDER)
    a = ap.parse_args(argv)
    install(a.ref, a.stub_missing)
    os.chdir(a.ref)
    sys.argv = [a.script] + a.script_args
    runpy.run_path(os.path.join(a.ref, a.script), run_name="__main__")


if __name__ == "__main__":
    main()
