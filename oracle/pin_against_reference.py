"""ORACLE pinning — runs ONLY in the build container (needs /root/reference; CPU).

1. imports the unmodified reference (nka77/DAHiTra) through import shims,
2. checks that dahitra_b200's module tree reproduces the reference state_dict bit for bit under the
   same seed (LEVIR variant through define_G; xBD variant through default init),
3. checks that oracle/dahitra_oracle.py reproduces the reference logits and intermediate taps on
   identical weights and inputs,
4. writes the golden fixtures under tests/golden/ that travel to the GPU box.

    python oracle/pin_against_reference.py            # regenerate fixtures, prints the pin report
"""
from __future__ import annotations

import json
import os
import sys
import types
import warnings

import numpy as np
import torch

warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("DAHITRA_REFERENCE", "/root/reference")
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle import dahitra_oracle as O          # noqa: E402
from oracle import synth                        # noqa: E402


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


def import_reference():
    """-> (define_G, xbd BASE_Transformer_UNet class) from the unmodified reference tree."""
    class DropPath(torch.nn.Identity):
        def __init__(self, *a, **k):
            super().__init__()
    _stub("timm.models.layers", DropPath=DropPath, to_2tuple=lambda x: (x, x),
          trunc_normal_=torch.nn.init.trunc_normal_)
    _stub("matplotlib.pyplot")
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "models" or k.startswith("models.")}
    sys.path.insert(0, REF)
    import models.resnet as R                                         # reference models/resnet.py
    R._resnet = lambda arch, block, layers, pretrained, progress, **kw: R.ResNet(block, layers, **kw)
    from models.networks import define_G                              # reference models/networks.py:130
    cwd = os.getcwd()
    os.chdir(os.path.join(REF, "xBD_code"))
    sys.path.insert(0, os.path.join(REF, "xBD_code"))
    import zoo.model_transformer_encoding as X                        # reference xBD variant
    X.bitmodule._resnet = lambda arch, block, layers, pretrained, progress, **kw: X.bitmodule.ResNet(block, layers, **kw)
    os.chdir(cwd)
    return define_G, X.BASE_Transformer_UNet


class Args:
    net_G = "newUNetTrans"


def maxdiff(a, b):
    return float((a.double() - b.double()).abs().max())


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    ref_define_G, RefXbd = import_reference()
    from dahitra_b200.networks import define_G as my_define_G
    from dahitra_b200.xbd import BASE_Transformer_UNet as MyXbd
    report = {"torch": torch.__version__}

    # ---- 2. state_dict identity under the same seed --------------------------------------------------
    torch.manual_seed(0)
    ref = ref_define_G(Args(), gpu_ids=[]).eval()
    torch.manual_seed(0)
    mine = my_define_G(Args(), gpu_ids=[]).eval()
    sr, sm = ref.state_dict(), mine.state_dict()
    assert list(sr.keys()) == list(sm.keys()), "LEVIR key order differs"
    for k in sr:
        assert sr[k].shape == sm[k].shape and sr[k].dtype == sm[k].dtype, k
        assert torch.equal(sr[k], sm[k]), f"seeded init differs at {k}"
    mine.load_state_dict(sr, strict=True); ref.load_state_dict(sm, strict=True)
    report["levir_keys"] = len(sr)
    report["levir_params"] = int(sum(p.numel() for p in ref.parameters()))
    print(f"[pin] LEVIR state_dict: {len(sr)} keys, bit-identical seeded init, strict load both ways")

    kw = dict(input_nc=3, output_nc=5, token_len=4, resnet_stages_num=4, with_pos="learned",
              with_decoder_pos="learned", enc_depth=1, dec_depth=8)
    torch.manual_seed(0)
    refx = RefXbd(**kw).eval()
    torch.manual_seed(0)
    minex = MyXbd(**kw).eval()
    sxr, sxm = refx.state_dict(), minex.state_dict()
    assert list(sxr.keys()) == list(sxm.keys()), "xBD key order differs"
    for k in sxr:
        assert torch.equal(sxr[k], sxm[k]), f"xBD seeded init differs at {k}"
    minex.load_state_dict(sxr, strict=True)
    report["xbd_keys"] = len(sxr)
    print(f"[pin] xBD state_dict: {len(sxr)} keys, bit-identical seeded init")

    # ---- 3/4. oracle vs reference, fixtures --------------------------------------------------------
    def tap_hooks(net, store):
        hs = []
        for k in (5, 4, 3):
            hs.append(getattr(net, f"transformer_{k}").register_forward_hook(
                lambda m, i, o, k=k: store.__setitem__(f"tokens_{k}", o.detach())))
        return hs

    cases = []
    # (a) define_G init (tiny logits), N(0,1) inputs, B=1
    x1, x2 = synth.synth_pair(1, 256, 256, seed=1, kind="normal")
    cases.append(("levir_defineG_seed0_normal", sr, x1, x2))
    # (b) synthetic default-scale weights, uniform [-1,1] inputs, B=2
    ssd = synth.synth_state_dict(sr, seed=3, style="default")
    x1, x2 = synth.synth_pair(2, 256, 256, seed=2, kind="uniform")
    cases.append(("levir_synth3_uniform", ssd, x1, x2))
    # (c) synthetic small-scale weights, uint8-statistics inputs, B=1
    ssd2 = synth.synth_state_dict(sr, seed=4, style="small")
    x1, x2 = synth.synth_pair(1, 256, 256, seed=5, kind="u8")
    cases.append(("levir_synth4_u8", ssd2, x1, x2))

    fp = {"levir_defineG_seed0": synth.fingerprint(sr)}
    import copy
    ref64 = copy.deepcopy(ref).double()
    for name, sd, x1, x2 in cases:
        ref.load_state_dict(sd, strict=True)
        ref64.load_state_dict(sd, strict=True)
        taps_ref, taps = {}, {}
        hs = tap_hooks(ref, taps_ref)
        with torch.no_grad():
            y_ref = ref(x1, x2)
            y_ref64 = ref64(x1.double(), x2.double())
        for h in hs:
            h.remove()
        y = O.forward_levir(sd, x1, x2, taps=taps)
        y64 = O.forward_levir(sd, x1, x2, dtype=torch.float64)
        sem = maxdiff(y64, y_ref64)                 # semantic pin: fp64 oracle == fp64 reference
        d, noise = maxdiff(y, y_ref), maxdiff(y_ref, y_ref64)
        agree = float((y.argmax(1) == y_ref.argmax(1)).float().mean())
        dt = {k: maxdiff(taps[k], taps_ref[k]) for k in taps_ref}
        print(f"[pin] {name}: fp64 oracle-vs-fp64 ref {sem:.2e}; fp32 oracle-vs-ref {d:.2e}; reference fp32 noise {noise:.2e}; "
              f"argmax agree {agree:.6f}; logit std {float(y_ref.std()):.4f}; token taps {max(dt.values()):.1e}")
        assert sem < 1e-9, name
        assert d <= 4 * noise + 1e-6, name
        report[name] = dict(semantic_fp64=sem, fp32_oracle_vs_ref=d, ref_fp32_noise=noise, argmax_agree=agree,
                            logit_std=float(y_ref.std()))
        np.savez_compressed(os.path.join(GOLD, name + ".npz"),
                            logits=y_ref.numpy().astype(np.float32),
                            logits_f64ref=y_ref64.numpy().astype(np.float32),
                            tokens_5=taps_ref["tokens_5"].numpy(), tokens_4=taps_ref["tokens_4"].numpy(),
                            tokens_3=taps_ref["tokens_3"].numpy(),
                            level_5=taps["level_5"].numpy().astype(np.float32),
                            level_4=taps["level_4"].numpy().astype(np.float32))
    del ref64
    # (d) xBD variant, 1024x1024, B=1, 5 classes (subsampled fixture: every 8th pixel + float64 checksums)
    sxs = synth.synth_state_dict(sxr, seed=6, style="default")
    refx.load_state_dict(sxs, strict=True)
    g = torch.Generator().manual_seed(7)
    xx = torch.randint(0, 256, (1, 6, 1024, 1024), generator=g).float() / 127 - 1     # xBD_code/utils.py:112-116
    with torch.no_grad():
        yx_ref = refx(xx)
        yx_ref64 = copy.deepcopy(refx).double()(xx.double())
    yx = O.forward_xbd(sxs, xx)
    yx64 = O.forward_xbd(sxs, xx, dtype=torch.float64)
    sem, d, noise = maxdiff(yx64, yx_ref64), maxdiff(yx, yx_ref), maxdiff(yx_ref, yx_ref64)
    agree = float((yx.argmax(1) == yx_ref.argmax(1)).float().mean())
    print(f"[pin] xbd_synth6_1024: fp64 oracle-vs-fp64 ref {sem:.2e}; fp32 oracle-vs-ref {d:.2e}; reference fp32 noise {noise:.2e}; "
          f"argmax agree {agree:.6f}; logit std {float(yx_ref.std()):.4f}")
    assert sem < 1e-9 and d <= 4 * noise + 1e-6
    report["xbd_synth6_1024"] = dict(semantic_fp64=sem, fp32_oracle_vs_ref=d, ref_fp32_noise=noise, argmax_agree=agree,
                                     logit_std=float(yx_ref.std()))
    np.savez_compressed(os.path.join(GOLD, "xbd_synth6_1024.npz"),
                        logits_sub8=yx_ref[:, :, ::8, ::8].numpy().astype(np.float32),
                        logits_f64ref_sub8=yx_ref64[:, :, ::8, ::8].numpy().astype(np.float32),
                        logits_f64ref_sum=np.float64(yx_ref64.sum()), logits_f64ref_abs_sum=np.float64(yx_ref64.abs().sum()),
                        argmax_hist=np.bincount(yx_ref64.argmax(1).flatten().numpy(), minlength=5))
    fp["xbd_default_seed0"] = synth.fingerprint(sxr)
    fp["levir_synth3"] = synth.fingerprint(ssd)
    fp["xbd_synth6"] = synth.fingerprint(sxs)
    json.dump({"keys_levir": [[k, list(v.shape), str(v.dtype)] for k, v in sr.items()],
               "keys_xbd": [[k, list(v.shape), str(v.dtype)] for k, v in sxr.items()],
               "fingerprints": fp, "report": report},
              open(os.path.join(GOLD, "state_dict_contract.json"), "w"))
    print("[pin] fixtures written to", GOLD)


if __name__ == "__main__":
    main()
