#!/usr/bin/env python
"""Place an UNMODIFIED copy of the reference's Python sources (and its shipped LEVIR sample images) under
``baseline/_ref/ref`` so that the reference arm of bench.py and the harness tests can import it on the GPU box,
where /root/reference does not exist.

The reference has no setup.py / pyproject.toml, so ``pip install --target baseline/_ref /root/reference`` has nothing
to build ("neither 'setup.py' nor 'pyproject.toml' found"); the files are copied byte for byte instead.
``baseline/_ref/`` is git-ignored (never part of this repository's history) but travels with gpurun snapshots.

    python baseline/install_reference.py [--src /root/reference]
"""
from __future__ import annotations

import argparse
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref", "ref")
# what the newUNetTrans path and its harness (eval_cd.py / main_cd.py / evaluator / trainer / loaders / metrics) import
TREES = ["models", "datasets", "misc", os.path.join("xBD_code", "zoo"), os.path.join("data", "LEVIR_CD")]
FILES = ["utils.py", "data_config.py", "eval_cd.py", "main_cd.py", "demo.py", os.path.join("xBD_code", "utils.py")]


def install(src: str = "/root/reference", dst: str = DST) -> bool:
    if not os.path.isdir(src):
        return False
    for t in TREES:
        s, d = os.path.join(src, t), os.path.join(dst, t)
        if os.path.isdir(s):
            shutil.copytree(s, d, dirs_exist_ok=True, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    for f in FILES:
        s, d = os.path.join(src, f), os.path.join(dst, f)
        if os.path.isfile(s):
            os.makedirs(os.path.dirname(d), exist_ok=True)
            shutil.copy2(s, d)
    return True


def installed(dst: str = DST) -> bool:
    return all(os.path.exists(os.path.join(dst, p)) for p in
               ("models/networks.py", "models/evaluator.py", "models/trainer.py", "utils.py", "datasets/CD_dataset.py",
                "data/LEVIR_CD/train/A"))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    a = ap.parse_args()
    ok = install(a.src)
    print(f"[install_reference] {'copied ' + a.src + ' -> ' + DST if ok else 'source tree ' + a.src + ' not present'}; "
          f"installed={installed()}")
