"""configs[3] smoke: LEVIR-CD training step of the drop-in module (batch 8 per GPU, 256x256, synthetic labels, CE loss, AdamW).
Training runs on the stock-autograd route (DESIGN.md "Training step"); under torchrun the module is wrapped in
DistributedDataParallel (NCCL gradient all-reduce, as BASELINE.json's config 4 describes).  After the last step the
module is switched to eval() and the NATIVE forward is checked against the autograd route on the updated weights.
   python tools/train_step.py [--steps 10]                                   (1 GPU)
   python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/train_step.py    (DDP)
Prints one JSON line from rank 0."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import torch.nn.functional as F
from dahitra_b200.networks import define_G


class Args:
    net_G = "newUNetTrans"


ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--batch", type=int, default=8)
a = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl")
torch.manual_seed(0)
net = define_G(Args(), gpu_ids=[local]).train()
model = net
if world > 1:
    # like the reference, the module owns parameters its forward never uses (scale-2 transformer, conv_pred, layer4)
    model = torch.nn.parallel.DistributedDataParallel(net, device_ids=[local], find_unused_parameters=True)
opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=0.01)           # models/trainer.py:39-40
g = torch.Generator(device="cuda").manual_seed(100 + rank)
x1 = torch.rand(a.batch, 3, 256, 256, device="cuda", generator=g) * 2 - 1
x2 = torch.rand(a.batch, 3, 256, 256, device="cuda", generator=g) * 2 - 1
y = (torch.rand(a.batch, 256, 256, device="cuda", generator=g) < 0.1).long()
losses = []
torch.cuda.synchronize()
t0 = time.time()
for step in range(a.steps):
    opt.zero_grad(set_to_none=True)
    loss = F.cross_entropy(model(x1, x2), y, weight=torch.ones(2, device="cuda"), ignore_index=255)   # models/losses.py:9-26
    loss.backward()
    opt.step()
    losses.append(float(loss.detach()))
torch.cuda.synchronize()
dt = time.time() - t0
no_grad = [n for n, p in net.named_parameters() if p.requires_grad and p.grad is None]
gnorm = float(torch.sqrt(sum((p.grad.float() ** 2).sum() for p in net.parameters() if p.grad is not None)))
if world > 1:                                   # replicas must hold identical weights after the all-reduced steps
    w = torch.cat([p.detach().flatten()[:64] for p in net.parameters()])
    ref = w.clone()
    dist.broadcast(ref, 0)
    same = bool(torch.equal(w, ref))
    flag = torch.tensor([int(same)], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    replicas_equal = bool(flag.item())
else:
    replicas_equal = None
net.eval()
with torch.no_grad():
    y_native = net(x1, x2)                      # native sm_100a forward on the UPDATED weights
y_auto = net._forward_autograd(x1, x2).detach()
native_vs_autograd = float((y_native - y_auto).abs().max())
if rank == 0:
    print(json.dumps(dict(workload=f"LEVIR-CD training step, batch {a.batch} x {world} GPU(s), CE loss, AdamW (autograd route"
                                   + (", DDP/NCCL all-reduce)" if world > 1 else ")"),
                          steps=a.steps, steps_per_s=a.steps / dt, pairs_per_s=a.steps * a.batch * world / dt,
                          loss_first=losses[0], loss_last=losses[-1], grad_norm_last=gnorm, params_without_grad=len(no_grad),
                          replicas_equal=replicas_equal, native_vs_autograd_after_training_max_abs=native_vs_autograd)))
if world > 1:
    dist.destroy_process_group()
