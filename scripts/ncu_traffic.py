"""profiles/step_kernel_traffic.json from an `ncu --set full` capture of ONE step-kernel launch of the bench command:
    python scripts/ncu_traffic.py gpurun_out/r02g_step_tc.ncu-rep "capture name / command"
bench.py reports the sum as roofline.traffic as long as the kernel sources named here are unchanged (sha256)."""
import csv, hashlib, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "mobody-model-based-off-dynamics-offline-reinforcement-learning_b200", "csrc")
SOURCES = ["step_tc.cu", "tc_epi.cuh", "tc_prims.cuh", "tc_layout.h", "term.cuh", "philox.cuh", "common.cuh"]


def to_bytes(v, unit):
    x = float(v.replace(",", ""))
    return int(round(x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]))


raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, r = rows[0], rows[1], rows[2]
g = lambda k: (r[hdr.index(k)], units[hdr.index(k)])
h = hashlib.sha256()
for fn in SOURCES:
    with open(os.path.join(CSRC, fn), "rb") as f:
        h.update(f.read())
rec = {"kernel": g("Kernel Name")[0], "grid": g("Grid Size")[0],
       "dram_bytes_read": to_bytes(*g("dram__bytes_read.sum")), "dram_bytes_write": to_bytes(*g("dram__bytes_write.sum")),
       "duration_ns_under_ncu": g("gpu__time_duration.sum")[0],
       "capture": sys.argv[2] if len(sys.argv) > 2 else os.path.basename(sys.argv[1]),
       "sources": SOURCES, "sources_sha256": h.hexdigest()}
with open(os.path.join(ROOT, "profiles", "step_kernel_traffic.json"), "w") as f:
    json.dump(rec, f, indent=1)
print(rec)
