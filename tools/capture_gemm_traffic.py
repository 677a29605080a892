#!/usr/bin/env python
"""Re-capture profiles/gemm_traffic.json: one `ncu --set full` pass over the dominant launch of the bench step (the FFN-1 GEGLU
GEMM, tools/geglu_one.py) and a stamp with the sha256 of the csrc/gemm.cu it was taken from (bench.py reports a capture of
another source as stale).  Run on the GPU box: python tools/capture_gemm_traffic.py"""
import csv
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = os.path.join(ROOT, "gpurun_out", "gemm_traffic_capture")
os.makedirs(os.path.dirname(rep), exist_ok=True)
subprocess.run(["ncu", "--set", "full", "--clock-control", "none", "-k", "regex:gemm2_tcgen05_kernel", "-s", "3", "-c", "1", "-o", rep, "-f",
                sys.executable, os.path.join(ROOT, "tools", "geglu_one.py")], check=True, capture_output=True)
out = subprocess.run(["ncu", "-i", rep + ".ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, row = rows[0], rows[1], rows[2]
col = {h: i for i, h in enumerate(hdr)}


def val(name):
    v = float(row[col[name]].replace(",", ""))
    u = units[col[name]]
    return v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(u, 1.0)


rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
path = os.path.join(ROOT, "profiles", "gemm_traffic.json")
d = json.load(open(path)) if os.path.exists(path) else {}
d.update({
    "source": "ncu --set full --clock-control none capture of tools/geglu_one.py by tools/capture_gemm_traffic.py",
    "kernel": row[col["Kernel Name"]][:120] + " (FFN-1: M=125440, N=2x2048, K=768)",
    "dram_bytes_read": int(rd), "dram_bytes_write": int(wr), "dram_bytes_per_launch": int(rd + wr),
    "algorithmic_bytes_per_launch": 125440 * 768 * 2 + 4096 * 768 * 2 + 125440 * 2048 * 2 + 125440 * 4096 * 2,
    "tensor_pipe_active_pct": float(row[col["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]]),
    "duration_us": float(row[col["gpu__time_duration.sum"]].replace(",", "")),
    "gemm_cu_sha256": hashlib.sha256(open(os.path.join(ROOT, "incomplete_multimodal_fusion_b200", "csrc", "gemm.cu"), "rb").read()).hexdigest(),
    "captured_at_commit": os.environ.get("MMF_COMMIT", d.get("captured_at_commit")),
})
json.dump(d, open(os.path.join(ROOT, "gpurun_out", "gemm_traffic.json"), "w"), indent=1)
print(json.dumps({k: d[k] for k in ("dram_bytes_per_launch", "algorithmic_bytes_per_launch", "tensor_pipe_active_pct", "duration_us")}))
