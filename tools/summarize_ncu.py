"""Summarise an .ncu-rep (from `ncu --set full`) into a small JSON: per captured kernel the duration, DRAM bytes,
throughput percentages, occupancy limits and the top warp-stall reasons.  Run here (no GPU needed):
    python tools/summarize_ncu.py gpurun_out/x.ncu-rep profiles/r1_x.json"""
import csv
import json
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_ncu_peak",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed": "l1tex_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "smsp__inst_executed.sum": "warp_instructions",
    "lts__t_sectors_srcunit_tex_op_read.sum": "l2_read_sectors_from_sm",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts",
    "l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed": "l1_to_l2_request_port_pct",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "l1_data_pipe_pct",
}


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    kernels = []
    for r in rows[2:]:
        k = {"kernel": r[idx["Kernel Name"]].split("(")[0]}
        for key, name in KEYS.items():
            if key in idx and r[idx[key]] != "":
                k[name] = f"{r[idx[key]]} {units[idx[key]]}".strip()
        stalls = {}
        for i, h in enumerate(hdr):
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
                try:
                    stalls[h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = round(float(r[i]), 3)
                except ValueError:
                    pass
        k["top_stalls_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:6])
        kernels.append(k)
    with open(out, "w") as f:
        json.dump({"source": rep, "kernels": kernels}, f, indent=1)
    print(f"wrote {out}: {len(kernels)} kernel launches")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
