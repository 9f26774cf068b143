"""Turn an .ncu-rep (read here on the CPU box with `ncu -i`) into a small committed summary.

    python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r02_xxx.json [algorithmic_flop] [algorithmic_bytes]
"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size",
    "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "sm__inst_executed_pipe_xu_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu_realtime.avg.pct_of_peak_sustained_elapsed",
    "lts__lts2xbar_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
]


def key_of(header: str):
    """The raw page prefixes some metrics with their section ("TPC.TriageCompute.sm__pipe_tensor_..."):
    match the metric name by SUFFIX, so those land in the summary under their plain name."""
    for k in KEYS:
        if header == k or header.endswith("." + k):
            return k
    return None


def main():
    rep, out = sys.argv[1], sys.argv[2]
    flop = float(sys.argv[3]) if len(sys.argv) > 3 else None
    nbytes = float(sys.argv[4]) if len(sys.argv) > 4 else None
    if rep.endswith(".csv"):     # a raw page exported on the GPU box (`ncu -i x.ncu-rep --page raw --csv`)
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    launches = []
    for r in rows[2:]:
        d = {}
        for i, h in enumerate(hdr):
            k = h if h in ("Kernel Name", "ID") else key_of(h)
            if k is None or (k in d and not r[i]):
                continue
            d[k] = r[i] + ((" " + units[i]) if units[i] and k not in ("Kernel Name", "ID") else "")
        launches.append(d)
    summary = {"report": rep, "launches": launches}
    # per-launch DRAM traffic (bytes) of the first profiled launch, for bench.py's roofline.traffic
    def to_bytes(s):
        v, u = s.split()
        return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]
    tr = [to_bytes(l["dram__bytes_read.sum"]) + to_bytes(l["dram__bytes_write.sum"]) for l in launches]
    summary["dram_bytes_per_launch"] = sum(tr) / len(tr)
    if flop:
        summary["algorithmic_flop_per_launch"] = flop
    if nbytes:
        summary["algorithmic_bytes_per_launch"] = nbytes
        summary["traffic_over_algorithmic"] = summary["dram_bytes_per_launch"] / nbytes
    with open(out, "w") as f:
        json.dump(summary, f, indent=1)
    print(json.dumps(summary, indent=1))


if __name__ == "__main__":
    main()
