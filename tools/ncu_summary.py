"""Condense one `ncu --set full` report (first kernel) into the JSON bench.py reads for `roofline.traffic`
and profiles/README.md quotes.  usage: python tools/ncu_summary.py report.ncu-rep "shape label" out.json"""
import csv
import json
import subprocess
import sys

rep, label, out = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = {
    "gpu__time_duration.sum": "duration", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "smem_tc_wavefronts_pct_of_peak",
    "l1tex__m_xbar2l1tex_read_bytes.sum": "l2_to_sm_read", "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct", "launch__registers_per_thread": "regs_per_thread",
    "launch__grid_size": "grid", "launch__block_size": "block", "launch__shared_mem_per_block_dynamic": "dyn_smem",
    "gpc__cycles_elapsed.max.per_second": "sm_clock", "Kernel Name": "kernel",
}
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
d = {"shape": label, "source": rep.split("/")[-1]}
for h, u, v in zip(hdr, units, vals):
    if h in want:
        k = want[h]
        try:
            x = float(v.replace(",", ""))
        except ValueError:
            d[k] = v
            continue
        if u in scale:
            x *= scale[u]
            k += "_bytes"
        elif u in ("ms", "us", "ns"):
            x *= {"ms": 1e-3, "us": 1e-6, "ns": 1e-9}[u]
            k += "_s"
        elif u == "Ghz":
            k += "_ghz"
        d[k] = x
d["dram_traffic_bytes"] = d.get("dram_read_bytes", 0) + d.get("dram_write_bytes", 0)
json.dump(d, open(out, "w"), indent=1)
print(json.dumps(d))
