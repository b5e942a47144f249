"""Checkpoint helpers: reference checkpoint wrappers, DataParallel prefixes, LEVIR <-> xBD key layouts,
prepared-weight export."""
import numpy as np
import torch

from dahitra_b200 import checkpoints as C
from dahitra_b200.networks import BASE_Transformer_UNet
from dahitra_b200.xbd import BASE_Transformer_UNet as XBD


def _levir():
    torch.manual_seed(1)
    return BASE_Transformer_UNet(3, 2, 'learned', resnet_stages_num=4, with_decoder_pos='learned', enc_depth=1, dec_depth=8)


def _xbd():
    torch.manual_seed(2)
    return XBD(input_nc=3, output_nc=5, token_len=4, resnet_stages_num=4, with_pos="learned", with_decoder_pos="learned",
               enc_depth=1, dec_depth=8)


def test_reference_checkpoint_wrappers(tmp_path):
    net = _levir()
    sd = net.state_dict()
    path = tmp_path / "best_ckpt.pt"                     # what models/trainer.py:150-160 writes
    torch.save({"epoch_id": 3, "best_val_acc": 0.9, "model_G_state_dict": sd, "optimizer_G_state_dict": {}}, path)
    got = C.load_checkpoint(str(path))
    assert list(got) == list(sd) and all(torch.equal(got[k], sd[k]) for k in sd)
    _levir().load_state_dict(got, strict=True)
    xsd = _xbd().state_dict()                            # what xBD_code/train.py:447-457 writes (DataParallel prefixes)
    wrapped = {"epoch": 1, "state_dict": {"module." + k: v for k, v in xsd.items()}, "best_score": 0.5}
    got = C.extract_state_dict(wrapped)
    assert list(got) == list(xsd)
    _xbd().load_state_dict(got, strict=True)


def test_key_layout_conversion_round_trip():
    lev, xbd = _levir(), _xbd()
    lsd, xsd = lev.state_dict(), xbd.state_dict()
    as_xbd = C.levir_to_xbd(lsd)
    assert set(as_xbd) == set(xsd) - {"classifier.weight", "classifier.bias"} | {"classifier.weight", "classifier.bias"}
    assert len(as_xbd) == 700
    for k, v in as_xbd.items():                          # alias keys carry the per-level tensors
        if k not in ("classifier.weight", "classifier.bias"):
            assert v.shape == xsd[k].shape, k
    assert torch.equal(as_xbd["transformer_decoder_layers.1.layers.0.0.fn.norm.weight"], lsd["transformer_decoder_3.layers.0.0.fn.norm.weight"])
    back = C.xbd_to_levir(as_xbd, template=lsd)
    assert list(sorted(back)) == list(sorted(lsd)) and all(torch.equal(back[k], lsd[k]) for k in lsd)
    as_lev = C.xbd_to_levir(xsd)                         # xBD weights under LEVIR keys (5-class head)
    lev5 = BASE_Transformer_UNet(3, 5, 'learned', resnet_stages_num=4, with_decoder_pos='learned', enc_depth=1, dec_depth=8)
    lev5.load_state_dict(as_lev, strict=True)


def test_export_prepared(tmp_path):
    from dahitra_b200.engine import slot_names
    net = _levir().eval()
    shapes = C.export_prepared(net, str(tmp_path / "prepared.npz"))
    z = np.load(tmp_path / "prepared.npz")
    names = slot_names()
    assert set(z.files) <= set(names) and "DH_W_STEM_W" in z.files and "DH_W_CL20A_WT" in z.files
    assert z["DH_W_STEM_W"].shape == (147, 64) and z["DH_W_STEM_W"].dtype == np.float32
    assert z["DH_W_CL20A_WT"].shape == (5, 128, 1152)    # TF32 hi | TF32 lo | bf16 pair | f16 / bf16 pair | scaled f16 remainder
    assert shapes["DH_W_CLS_W"] == tuple(z["DH_W_CLS_W"].shape)
