"""CPU: the oracle restatement against the golden fixtures generated from the real reference
(oracle/pin_against_reference.py), and the module/state_dict contract."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import dahitra_oracle as O
from oracle import synth


class Args:
    net_G = "newUNetTrans"


def _defineG_sd():
    from dahitra_b200.networks import define_G
    torch.manual_seed(0)
    return define_G(Args(), gpu_ids=[]).state_dict()


CASES = {
    "levir_defineG_seed0_normal": dict(weights="defineG", pair=(1, 1, "normal")),
    "levir_synth3_uniform": dict(weights=(3, "default"), pair=(2, 2, "uniform")),
    "levir_synth4_u8": dict(weights=(4, "small"), pair=(1, 5, "u8")),
}


def case_inputs(name, template):
    c = CASES[name]
    sd = _defineG_sd() if c["weights"] == "defineG" else synth.synth_state_dict(template, seed=c["weights"][0], style=c["weights"][1])
    B, seed, kind = c["pair"]
    x1, x2 = synth.synth_pair(B, 256, 256, seed=seed, kind=kind)
    return sd, x1, x2


def test_contract_keys(golden_dir, levir_template):
    con = json.load(open(os.path.join(golden_dir, "state_dict_contract.json")))
    keys = [[k, list(v.shape), str(v.dtype)] for k, v in levir_template.items()]
    assert keys == con["keys_levir"]                       # 425 keys, same order, shapes and dtypes as the reference
    assert len(keys) == 425
    from dahitra_b200.xbd import BASE_Transformer_UNet as X
    net = X(input_nc=3, output_nc=5, token_len=4, resnet_stages_num=4, with_pos="learned",
            with_decoder_pos="learned", enc_depth=1, dec_depth=8)
    keysx = [[k, list(v.shape), str(v.dtype)] for k, v in net.state_dict().items()]
    assert keysx == con["keys_xbd"] and len(keysx) == 700


def test_seeded_init_matches_reference(golden_dir):
    """torch.manual_seed(0); define_G(...) gives the reference's weights (per-key float64 sums recorded at pin time)."""
    con = json.load(open(os.path.join(golden_dir, "state_dict_contract.json")))
    fp = synth.fingerprint(_defineG_sd())
    ref = con["fingerprints"]["levir_defineG_seed0"]
    assert fp.keys() == ref.keys()
    for k in fp:
        assert fp[k] == pytest.approx(ref[k], rel=0, abs=1e-9), k


def test_synth_weights_reproducible(golden_dir, levir_template):
    con = json.load(open(os.path.join(golden_dir, "state_dict_contract.json")))
    fp = synth.fingerprint(synth.synth_state_dict(levir_template, seed=3, style="default"))
    ref = con["fingerprints"]["levir_synth3"]
    for k in fp:
        assert fp[k] == pytest.approx(ref[k], rel=1e-12, abs=1e-9), k


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_reproduces_reference_logits(name, golden_dir, levir_template):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    sd, x1, x2 = case_inputs(name, levir_template)
    taps = {}
    y64 = O.forward_levir(sd, x1, x2, dtype=torch.float64, taps=taps)
    ref64 = torch.from_numpy(g["logits_f64ref"]).double()
    # fixtures hold the fp64 reference rounded to fp32: 2^-24 relative + accumulated representation error
    assert float((y64 - ref64).abs().max()) <= 1e-6 * max(1.0, float(ref64.abs().max()))
    for k in (5, 4, 3):
        assert float((taps[f"tokens_{k}"] - torch.from_numpy(g[f"tokens_{k}"]).double()).abs().max()) < 1e-4
    if name != "levir_synth3_uniform":       # fp32 oracle: within the reference's own fp32 noise
        y = O.forward_levir(sd, x1, x2)
        assert float((y - torch.from_numpy(g["logits"])).abs().max()) < 1e-5


def test_oracle_edge_cases(levir_template):
    """constant images (token softmax is uniform), identical pre/post (difference tokens are exactly zero)."""
    sd = synth.synth_state_dict(levir_template, seed=9, style="default")
    x = torch.zeros(1, 3, 256, 256)
    y = O.forward_levir(sd, x, x)
    assert torch.isfinite(y).all()
    x1, _ = synth.synth_pair(1, 256, 256, seed=11)
    taps = {}
    O.forward_levir(sd, x1, x1.clone(), taps=taps)
    for k in (5, 4, 3):
        t = taps[f"tokens_{k}"]
        assert float((t[:, :4] - t[:, 4:]).abs().max()) > 0      # pos-emb differs between the two halves
