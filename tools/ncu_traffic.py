"""Take the `ncu --set full` capture that bench.py's roofline.traffic quotes (run under gpurun on ONE GPU):
    python tools/ncu_traffic.py [log2n=24] [out_dir=gpurun_out]
Captures the second launch of the bucket-accumulation kernel the library chooses at that size, keeps the .ncu-rep, prints the key
metrics (tools/ncu_summary.py raw) and writes <out_dir>/r02_traffic.json with the kernel-source hash bench.py checks, so that a
capture can never be quoted for code it was not taken on.  Copy the JSON and the text summary into profiles/ afterwards."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (kernel_source_hash)

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 24
out_dir = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out")
os.makedirs(out_dir, exist_ok=True)
rep = os.path.join(out_dir, "r02_accumulate_2p%d" % lg)
cmd = ["ncu", "--set", "full", "--clock-control", "none", "--import-source", "on", "-k", "regex:k_bucket_accumulate", "-s", "1", "-c", "1",
       "-f", "-o", rep, sys.executable, os.path.join(ROOT, "tools", "ba_ncu_target.py"), str(lg), "0", "2"]
print(" ".join(cmd), flush=True)
subprocess.run(cmd, check=True)
raw = subprocess.run(["ncu", "-i", rep + ".ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, vals = rows[0], rows[2]
d = dict(zip(hdr, vals))
unit = dict(zip(hdr, rows[1]))


def to_bytes(key):
    v = float(d[key].replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit[key]]


kernel = "k_bucket_accumulate_affine" if "affine" in d["Kernel Name"] else "k_bucket_accumulate"
import snark_verifier_b200 as sv  # noqa: E402
L = sv.CudaLoader(0)
plan = L.msm_plan(1 << lg)
L.close()
rec = {kernel: {"log_n": lg, "window_bits": plan["window_bits"],
                "dram_bytes_per_launch": to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum"),
                "dram_read_bytes": to_bytes("dram__bytes_read.sum"), "dram_write_bytes": to_bytes("dram__bytes_write.sum"),
                "fmaheavy_pct": float(d["sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"]),
                "dram_pct_of_peak": float(d["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]),
                "kernel_ms_under_ncu": float(d["gpu__time_duration.sum"].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}.get(unit["gpu__time_duration.sum"], 1),
                "source_hash": bench.kernel_source_hash(), "source": "profiles/r02_ncu_%s_2p%d.txt" % (kernel, lg)}}
with open(os.path.join(out_dir, "r02_traffic.json"), "w") as f:
    json.dump(rec, f, indent=1)
print(json.dumps(rec, indent=1))
txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), "raw", rep + ".ncu-rep"], capture_output=True, text=True).stdout
with open(os.path.join(out_dir, "r02_ncu_%s_2p%d.txt" % (kernel, lg)), "w") as f:
    f.write("# ncu --set full --clock-control none --import-source on -k regex:k_bucket_accumulate -s 1 -c 1  (python tools/ba_ncu_target.py %d 0 2)\n" % lg)
    f.write("# kernel sources sha256[:16] = %s\n" % bench.kernel_source_hash())
    f.write(txt)
print(txt)
