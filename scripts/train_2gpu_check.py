"""Data-parallel training check on 2 GPUs (run with torchrun --nproc-per-node 2 under gpurun --gpus 2).
Each rank trains on its own batch slice (split_render_data order); the library's own bucketed ncclAllReduce runs inside the backward
(JOIN=0: torch.distributed all-reduce after it instead), and the flat gradient on every rank must equal the SUM of the oracle's
per-shard gradients (BatchNorm statistics are per device, like the reference); after the Adam step the replicas must be identical."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import yolo_b200  # noqa: E402
from oracle import nets, train, weights  # noqa: E402
from yolo_b200 import parallel  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
spec = nets.spec_tiny(size=(64, 96), C=10)
B = 4
params = weights.make_params("carnet", spec, seed=11, calib_batch=2)
x, _ = weights.synthetic_frames(B, spec["size"], seed=12)
labels = train.synthetic_labels(B, 4, nobj=1, seed=13, p_box=1.0)
hp = train.V1_HPARAMS
lo, hi = parallel.shard_bounds(B, rank, world)
net = yolo_b200.Net("carnet", spec, precision="fp16x3", max_batch=hi - lo, device=local)
net.load_params(params)
tr = yolo_b200.Trainer(net, learning_rate=0.001, join_nccl=os.environ.get("JOIN", "1") != "0", bucket_bytes=1 << 16)     # small buckets: several exchanges
tr.forward_backward(torch.from_numpy(x[lo:hi]).cuda(), labels[lo:hi], hp["scale"], hp["positive_weight"], hp["negative_weight"])
tr.allreduce_grads()
shapes = dict(net.param_shapes())
ref = [train.train_step("carnet", spec, params, x[a:b], labels[a:b], hp, batch_size=B) for a, b in (parallel.shard_bounds(B, r, world) for r in range(world))]
worst = 0.0
for name in ref[0]["grads"]:
    want = sum(r["grads"][name] for r in ref)
    got = tr.get_param(name, shapes[name], grad=True)
    worst = max(worst, float(np.abs(got - want).max() / max(np.abs(want).max(), 1e-6)))
tr.step(B)
w = torch.from_numpy(tr.get_param("stages.0.weight", shapes["stages.0.weight"])).cuda()
ws = [torch.empty_like(w) for _ in range(world)]
dist.all_gather(ws, w)
same = all(torch.equal(ws[0], t) for t in ws)
print(f"rank {rank} (library NCCL joined: {tr.nccl_joined}): all-reduced gradient vs sum of oracle shard gradients: worst rel err {worst:.2e}; replicas identical after the step: {same}", flush=True)
assert worst < 5e-2 and same      # per-shard batches of 2 images: BatchNorm backward on 12-sample statistics is ill-conditioned in fp32
dist.destroy_process_group()
