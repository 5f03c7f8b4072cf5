"""Turn an .ncu-rep (ncu --set full) into the text summary committed under profiles/: one block of key metrics per captured launch.
Usage: python scripts/ncu_export.py gpurun_out/X.ncu-rep > profiles/r2_ncu_X.txt"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__cluster_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__average_warp_latency_per_inst_issued.ratio"]
STALLS = "smsp__average_warps_issue_stalled_"
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
print(f"# {sys.argv[1]}: ncu --set full --clock-control none, one block per captured launch (replayed ~40x: times are cold-cache)")
for r in rows[2:]:
    d = {h: (v, u) for h, v, u in zip(hdr, r, units)}
    print(f"\n== {d['Kernel Name'][0]}   grid {d.get('Grid Size', ('?',))[0]} block {d.get('Block Size', ('?',))[0]}")
    for k in KEYS:
        if k in d:
            print(f"  {k:75s} {d[k][0]:>16s} {d[k][1]}")
    st = sorted(((float(v[0]), h[len(STALLS):-len('_per_issue_active.ratio')]) for h, v in d.items()
                 if h.startswith(STALLS) and h.endswith("_per_issue_active.ratio") and "not_issued" not in h), reverse=True)
    print("  warp stall reasons (warps stalled per issue-active cycle): " + ", ".join(f"{n} {v:.2f}" for v, n in st[:8]))
