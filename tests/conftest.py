import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


# the golden fixtures were generated from the reference with its trunk download patched away (fresh BatchNorm statistics): keep a
# resnet18 file that happens to sit in this machine's torch hub cache out of the seeded-init comparisons
os.environ.setdefault("DAHITRA_RESNET18_CKPT", "none")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def levir_template():
    """state_dict template (keys/shapes/dtypes) of the LEVIR variant"""
    import torch
    from dahitra_b200.networks import BASE_Transformer_UNet
    torch.manual_seed(123)
    net = BASE_Transformer_UNet(3, 2, 'learned', resnet_stages_num=4, with_decoder_pos='learned', enc_depth=1, dec_depth=8)
    return net.state_dict()
