"""Turn an ncu launch list (gpu__time_duration.sum CSV of one bench step) into a per-layer table for Darknet-53 416, B=32.
Usage: python scripts/launch_table.py gpurun_out/launches.csv [passes]   (passes = MMA passes per product: 3 fp16x3, 6 bf16x6, 1 bf16)"""
import csv
import sys

B = 32


def dk53_convs(size=416):
    layers, channels = [1, 2, 8, 8, 4], [32, 64, 128, 256, 512, 1024]
    out = []
    h = size
    out.append(("stages.0", 3, channels[0], 3, h))
    cin = channels[0]
    for s, (n, ch) in enumerate(zip(layers, channels[1:]), 1):
        h //= 2
        out.append((f"stages.{s}.0", cin, ch, 3, h))
        for j in range(1, n + 1):
            out.append((f"stages.{s}.{j}.body.0", ch, ch // 2, 1, h))
            out.append((f"stages.{s}.{j}.body.1", ch // 2, ch, 3, h))
        cin = ch
    pyr = channels[-3:][::-1]
    hs = [size // 32, size // 16, size // 8]
    x = pyr[0]
    for i, pc in enumerate(pyr):
        h = hs[i]
        c = x
        for b in range(2):
            out.append((f"yolo_blocks.{i}.body.{2*b}", c, pc, 1, h)); out.append((f"yolo_blocks.{i}.body.{2*b+1}", pc, 2 * pc, 3, h)); c = 2 * pc
        out.append((f"yolo_blocks.{i}.body.4", 2 * pc, pc, 1, h))
        out.append((f"yolo_blocks.{i}.tip", pc, 2 * pc, 3, h))
        out.append((f"yolo_outputs.{i}", 2 * pc, 90, 1, h))
        if i < 2:
            out.append((f"transitions.{i}", pc, pyr[i + 1], 1, h))
            x = 2 * pyr[i + 1]
    return out


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
    passes = float(sys.argv[2]) if len(sys.argv) > 2 else 3
    names = [r[4] for r in rows]
    # the LAST complete step of the capture: 75 conv layers + decode_kernel<0>.  A layer with a split-K tail is three launches
    # (full waves, split tail with fp32 partial output, fix-up) - or two when it has no full wave; they are merged into one row.
    last = [i for i, n in enumerate(names) if "decode_kernel" in n][-1]
    first = max(i for i, n in enumerate(names[:last]) if "stem" in n)
    raw = rows[first:last + 1]
    order, i = [], 0
    while i < len(raw):
        r = list(raw[i])
        j = i + 1
        while j < len(raw) and "splitk_fixup" not in raw[i][4] and "decode" not in raw[i][4]:
            nxt = raw[j][4]
            is_tail = "conv_umma_kernel<2, 1," in nxt.replace("yb::", "") or "conv_umma_kernel<(int)2, (bool)1" in nxt
            if "splitk_fixup" in nxt or (is_tail and j + 1 < len(raw) and "splitk_fixup" in raw[j + 1][4]):
                r[-1] = str(float(r[-1]) + float(raw[j][-1]))
                j += 1
                if "splitk_fixup" in nxt:
                    break
            else:
                break
        order.append(r)
        i = j
    assert "stem" in order[0][4] and len(order) == 76, (order[0][4], len(order))
    convs = dk53_convs()
    ci, tot, tot_fl = 0, 0.0, 0.0
    groups = {}
    print(f"{'layer':26s} {'Cin':>5s} {'Cout':>5s} k {'HxW':>4s} {'GFLOP':>8s} {'us':>8s} {'TFLOP/s':>8s} {'MMA TF/s':>9s}")
    for r in order:
        t = float(r[-1]) / 1e3
        if "decode" in r[4]:
            print(f"{'decode_kernel<0>':26s} {'':>5s} {'':>5s}   {'':>4s} {'':>8s} {t:8.1f}")
            tot += t
            continue
        name, cin, cout, k, h = convs[ci]; ci += 1
        fl = 2.0 * B * h * h * cout * cin * k * k
        tf = fl / t / 1e6
        print(f"{name:26s} {cin:5d} {cout:5d} {k} {h:4d} {fl/1e9:8.1f} {t:8.1f} {tf:8.1f} {tf*passes:9.1f}")
        tot += t; tot_fl += fl
        key = "stem" if cin == 3 else (f"{k}x{k} @{h}")
        g = groups.setdefault(key, [0.0, 0.0]); g[0] += t; g[1] += fl
    print(f"\ntotal {tot/1e3:.2f} ms, {tot_fl/1e12:.3f} TFLOP algorithmic -> {tot_fl/tot/1e6:.1f} TFLOP/s ({tot_fl/tot/1e6*passes:.0f} TFLOP/s of 16-bit MMA issued)")
    print("\nshare by class (cold-cache, serialised: compare SHARES):")
    for k, (t, fl) in sorted(groups.items(), key=lambda kv: -kv[1][0]):
        print(f"  {k:12s} {t/1e3:7.2f} ms {100*t/tot:5.1f} %  {fl/t/1e6:7.1f} TFLOP/s")


if __name__ == "__main__":
    main()
