#!/usr/bin/env python
"""Key counters of `ncu --set full` captures as JSON:  python profiles/ncu_summary.py a.ncu-rep [b.ncu-rep ...] > summary.json
(reads the reports with `ncu -i ... --page raw --csv`; no GPU needed)."""
import csv
import io
import json
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "time",
    "sm__cycles_elapsed.avg.per_second": "sm_clock",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_elapsed_pct",
    "sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active": "shared_pipe_active_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum": "tma_load_bytes",
    "l1tex__m_l1tex2xbar_write_bytes_mem_global_op_tma_st.sum": "tma_store_bytes",
    "smsp__inst_executed.sum": "warp_instructions",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__shared_mem_per_block_dynamic": "dynamic_smem",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
}


def summarise(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for vals in rows[2:]:
        d = {"report": path.split("/")[-1]}
        for h, u, v in zip(hdr, units, vals):
            if h == "Kernel Name":
                d["kernel"] = v
            elif h in KEYS:
                try:
                    d[KEYS[h]] = {"value": float(v.replace(",", "")), "unit": u}
                except ValueError:
                    pass
        out.append(d)
    return out


if __name__ == "__main__":
    res = []
    for p in sys.argv[1:]:
        res += summarise(p)
    print(json.dumps(res, indent=1))
