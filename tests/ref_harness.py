"""Drive the UNMODIFIED reference harness (models/evaluator.py CDEvaluator.eval_models, models/trainer.py CDTrainer)
from a copy of the reference tree, on either the reference's own network class or the native drop-in class
(dahitra_b200.launch.install rebinding), and print one JSON line with what it measured.

Runs as a script in its own process (it changes the working directory and registers import shims):

    python tests/ref_harness.py --ref baseline/_ref/ref --impl native --device cuda --what eval,train

Test infrastructure only.  Reference call sites exercised: models/evaluator.py:28 (define_G), :73 (strict
load_state_dict), :75 (.to), :164 (net_G(img_in1, img_in2)), :166-180 (eval_models), :95-104 (argmax -> numpy ->
confusion matrix); models/trainer.py:29, :39 (AdamW over net_G.parameters()), :247-262 (_forward_pass, _backward_G),
:299-310 (train step), :155 (state_dict in _save_checkpoint).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import tempfile
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", required=True)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--device", default="cuda", choices=["cuda", "cpu"])
    ap.add_argument("--what", default="eval,train")
    ap.add_argument("--mode", default=None)
    a = ap.parse_args()
    import numpy as np
    import torch
    from dahitra_b200 import launch
    ref = os.path.abspath(a.ref)
    nets = launch.install(ref, stub_missing=True, offline_trunk=True, rebind=(a.impl == "native"))
    os.chdir(ref)                                           # data_config.py holds relative data paths
    import utils as ref_utils                               # reference utils.py (loaders)
    tmp = tempfile.mkdtemp(prefix="dahitra_harness_")
    gpu_ids = [0] if a.device == "cuda" else []
    args = types.SimpleNamespace(net_G="newUNetTrans", n_class=2, gpu_ids=gpu_ids, checkpoint_dir=os.path.join(tmp, "ckpt"),
                                 vis_dir=os.path.join(tmp, "vis"), lr=1e-3, max_epochs=1, lr_policy="linear", loss="ce",
                                 batch_size=1, num_workers=0, data_name="LEVIR", dataset="CDDataset", split="train",
                                 split_val="train", img_size=256, project_name="harness")
    os.makedirs(args.checkpoint_dir); os.makedirs(args.vis_dir)
    # a fabricated best_ckpt.pt in the reference's own format (models/trainer.py:150-158): seeded define_G weights
    torch.manual_seed(0)
    net0 = nets.define_G(args=args, gpu_ids=[])
    sd0 = {k: v.detach().clone() for k, v in net0.state_dict().items()}
    torch.save({"epoch_id": 0, "best_val_acc": 0.5, "best_epoch_id": 0, "model_G_state_dict": sd0},
               os.path.join(args.checkpoint_dir, "best_ckpt.pt"))
    out = dict(impl=a.impl, device=a.device, net_class=type(net0).__module__ + "." + type(net0).__name__)
    del net0
    what = a.what.split(",")
    if "eval" in what:
        from models.evaluator import CDEvaluator               # reference, unmodified
        loader = ref_utils.get_loader("LEVIR", img_size=256, batch_size=4, is_train=False, split="train")
        ev = CDEvaluator(args=args, dataloader=loader)
        if a.mode and hasattr(ev.net_G, "set_mode"):
            ev.net_G.set_mode(a.mode)
        ev.eval_models()
        out["eval_scores"] = {k: float(v) for k, v in ev.running_metric.get_scores().items()}
        out["eval_cm"] = np.asarray(ev.running_metric.sum).astype(np.int64).tolist()
        out["eval_net_class"] = type(ev.net_G).__module__
        out["eval_log"] = open(os.path.join(args.checkpoint_dir, "log_test.txt")).read()[-600:]
        # the same batches through the oracle (fp64, CPU) on the checkpoint's weights: what the scores must be
        from oracle import dahitra_oracle as O
        from oracle import metrics_oracle as MO
        cm = np.zeros((2, 2), dtype=np.int64)
        agree, total, maxd = 0, 0, 0.0
        for batch in loader:
            ref_logits = O.forward_levir(sd0, batch["A"], batch["B"], dtype=torch.float64)
            pred = ref_logits.argmax(1).numpy()
            cm += MO.confuse_matrix(2, batch["L"].numpy(), pred).astype(np.int64)
            with torch.no_grad():
                ev._forward_pass(batch)
            got = ev.G_pred.detach().double().cpu()
            maxd = max(maxd, float((got - ref_logits).abs().max()))
            agree += int((got.argmax(1).numpy() == pred).sum()); total += pred.size
        out["oracle_cm"] = cm.tolist()
        out["oracle_scores"] = {k: float(v) for k, v in MO.cm2score(cm).items()}
        out["logits_max_abs_diff_vs_oracle"] = maxd
        out["argmax_agree_vs_oracle"] = agree / total
    if "train" in what:
        from models.trainer import CDTrainer                    # reference, unmodified
        import models.losses as ref_losses
        if a.device == "cpu":                                   # models/losses.py:24 calls .cuda() unconditionally
            torch.Tensor.cuda = lambda self, *x, **k: self       # (plumbing check of this script on a CPU-only machine)
        loaders = ref_utils.get_loaders(args)                   # batch_size 1 -> the cross-entropy branch of _backward_G
        tr = CDTrainer(args=args, dataloaders=loaders)
        tr.net_G.train()
        tr.is_training = True
        losses = []
        it = iter(loaders["train"])
        batch = next(it)
        for _ in range(3):                                      # the body of train_models' inner loop (trainer.py:299-310)
            tr._forward_pass(batch)
            tr.optimizer_G.zero_grad()
            tr._backward_G()
            tr.optimizer_G.step()
            torch.nn.utils.clip_grad_norm_(tr.net_G.parameters(), 0.999)
            losses.append(float(tr.G_loss.detach()))
        tr.net_G.eval()
        with torch.no_grad():
            tr._forward_pass(batch)                             # eval-mode forward on the UPDATED weights (native path)
        y_eval = tr.G_pred.detach().double().cpu()
        from oracle import dahitra_oracle as O
        sd1 = {k: v.detach().cpu() for k, v in tr.net_G.state_dict().items()}
        y_ref = O.forward_levir(sd1, batch["A"], batch["B"], dtype=torch.float64)
        d = (y_eval - y_ref).abs()
        out["train_losses"] = losses
        out["train_net_class"] = type(tr.net_G).__module__
        out["train_eval_after_max_abs_diff_vs_oracle"] = float(d.max())
        out["train_eval_after_outside_tol"] = int((d > 1e-4 + 1e-3 * y_ref.abs()).sum())
        out["train_eval_after_ref_absmax"] = float(y_ref.abs().max())
        tr._save_checkpoint("last_ckpt.pt")
        ck = torch.load(os.path.join(args.checkpoint_dir, "last_ckpt.pt"), map_location="cpu")
        out["ckpt_keys"] = len(ck["model_G_state_dict"])
    print("HARNESS_JSON " + json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
