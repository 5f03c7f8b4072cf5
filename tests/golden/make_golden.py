"""Generate the committed golden fixtures from the oracle (run from the repo root:
``python tests/golden/make_golden.py``).  The reference itself cannot run here, so these
vectors pin the ORACLE RESTATEMENT: any later drift of oracle/ or of numpy/torch numerics shows up in
``tests/test_oracle.py``, and the CUDA path is compared against the same files in ``tests/test_gpu_*.py``.
Fixtures are self-contained (inputs, parameters, outputs)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle import decode, nets, weights  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def net_fixture(name, net, spec, batch, seed):
    p = weights.make_params(net, spec, seed=seed, calib_batch=4)
    x, u8 = weights.synthetic_frames(batch, spec["size"], seed=seed + 7)
    with torch.no_grad():
        out = nets.forward(net, spec, weights.to_torch(p), torch.from_numpy(x))
    d = {"param:" + k: v for k, v in p.items()}
    d["frames_u8"] = u8
    if net == "carnet":
        heads = [h.numpy() for h in out]
        rows, idx = decode.predict(spec, heads, return_index=True)
        for i, h in enumerate(heads):
            d[f"head{i}"] = h
        d["rows"], d["idx"] = rows, idx
    elif net == "carlpnet":
        heads = [h.numpy() for h in out[0]]
        lp = out[1][0].numpy()
        rows, idx = decode.predict(spec, heads, return_index=True)
        lrows, lidx = decode.predict_LP_batch(spec, lp, return_index=True)
        for i, h in enumerate(heads):
            d[f"head{i}"] = h
        d["lp"], d["rows"], d["idx"], d["lp_rows"], d["lp_idx"] = lp, rows, idx, lrows, lidx
    else:
        o = out.numpy()
        row, idx = decode.predict_LP_single(spec, o, return_index=True)
        d["out"], d["lp_row"], d["lp_idx"] = o, row, np.int32(idx)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(name, sum(v.nbytes for v in d.values()) / 1e6, "MB raw")


def decode_fixture():
    spec = nets.spec_micro(size=(64, 96), C=10)
    heads = weights.synthetic_heads(3, spec, seed=7)
    heads[0][1, 5, 1, 0] = 30.0          # saturated ties: first index must win
    heads[1][1, 2, 0, 0] = 25.0
    heads[2][2, 0, 0, 0] = 18.5
    rows, idx = decode.predict(spec, heads, return_index=True)
    nm = decode.nms(spec, heads, score_thr=0.05, iou_thr=0.3, max_out=16, max_cand=256)
    d = {f"head{i}": h for i, h in enumerate(heads)}
    d["rows"], d["idx"] = rows, idx
    for b, (r, i) in enumerate(nm):
        d[f"nms_rows{b}"], d[f"nms_idx{b}"] = r, i
    np.savez_compressed(os.path.join(OUT, "decode_micro.npz"), **d)
    print("decode_micro", idx, [len(i) for _, i in nm])


def lp_micro_spec():
    s = nets.spec_lp_micro()
    s["size"] = [128, 128]
    return s


if __name__ == "__main__":
    net_fixture("carnet_micro", "carnet", nets.spec_micro(size=(128, 128)), 2, 11)
    net_fixture("carlpnet_micro", "carlpnet", nets.spec_micro(size=(128, 128), lp=True), 2, 12)
    net_fixture("lpdensenet_micro", "lpdensenet", lp_micro_spec(), 2, 13)
    decode_fixture()
