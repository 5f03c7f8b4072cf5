"""DRAM traffic and tensor-pipe activity of ONE inference step from an ncu CSV holding, per launch,
dram__bytes_read.sum, dram__bytes_write.sum, gpu__time_duration.sum and sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active.

  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
      --clock-control none --csv --log-file gpurun_out/traffic.csv python bench.py --steps 1 --warmup 3 --no-train --no-secondary --no-parity --no-cpu-baseline
  python scripts/step_traffic.py gpurun_out/traffic.csv profiles/r2_step_traffic_fp16x3.json

The LAST complete step of the capture is used (the launches from the last stem kernel to the last decode_kernel<0>).  bench.py reads the json
for `roofline.traffic` (a committed capture, not a number measured inside the timed run)."""
import csv
import json
import sys


def to_bytes(value, unit):
    v = float(value.replace(",", ""))
    return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
    launches = {}
    order = []
    for r in rows:
        lid = int(r[0])
        if lid not in launches:
            launches[lid] = {"name": r[4]}
            order.append(lid)
        metric, unit, value = r[-3], r[-2], r[-1]
        if metric.startswith("dram__bytes"):
            launches[lid][metric] = to_bytes(value, unit)
        elif metric == "gpu__time_duration.sum":
            launches[lid]["ns"] = float(value.replace(",", "")) * {"ns": 1.0, "us": 1e3, "ms": 1e6}.get(unit, 1.0)
        else:
            launches[lid]["tensor_pct"] = float(value.replace(",", ""))
    names = [launches[i]["name"] for i in order]
    last = max(i for i, n in enumerate(names) if "decode_kernel" in n)
    stems = [i for i, n in enumerate(names[:last]) if "stem" in n]
    first = max(stems) if stems else max(0, last - 75)          # a capture filtered to the tcgen05 kernel has no stem launch
    step = [launches[i] for i in order[first:last + 1]]
    conv = [l for l in step if "decode_kernel" not in l["name"]]
    rd = sum(l.get("dram__bytes_read.sum", 0.0) for l in conv)
    wr = sum(l.get("dram__bytes_write.sum", 0.0) for l in conv)
    umma = [l for l in conv if "conv_umma" in l["name"]]
    t_umma = sum(l["ns"] for l in umma)
    out = {
        "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,"
                  "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active over the last fp16x3 step of the capture "
                  f"({len(step)} launches), B=32, Darknet-53 416; cold-cache serialised",
        "dram_bytes_read_per_step": rd,
        "dram_bytes_write_per_step": wr,
        "conv_launches_dram_bytes_per_step": rd + wr,
        "tensor_pipe_active_pct_time_weighted": sum(l["ns"] * l.get("tensor_pct", 0.0) for l in umma) / max(t_umma, 1.0),
        "sum_kernel_time_ms": sum(l["ns"] for l in step) / 1e6,
        "launches": len(step),
    }
    json.dump(out, open(sys.argv[2], "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
