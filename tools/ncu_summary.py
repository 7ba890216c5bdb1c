"""ncu report -> committed JSON summary under profiles/ (one entry per captured launch), tagged with the build id of the
kernel sources (monocular-depth-estimation-cil_b200/build_id.txt) so that bench.py pairs a live timing with a `traffic`
figure only when both come from the same sources.
usage: python tools/ncu_summary.py <report.ncu-rep> <profiles/name.json> [note]   (run right after the capture)"""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, out = sys.argv[1], sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {n: i for i, n in enumerate(hdr)}
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def val(r, name, scale=False):
    if name not in col:
        return None
    try:
        v = float(r[col[name]].replace(",", ""))
    except ValueError:
        return None
    return v * UNIT.get(units[col[name]], 1) if scale else v


launches = []
for r in data:
    launches.append({
        "kernel": r[col["Kernel Name"]], "grid": r[col["Grid Size"]], "block": r[col["Block Size"]],
        "duration_us": val(r, "gpu__time_duration.sum", True),
        "dram_bytes_read": val(r, "dram__bytes_read.sum", True), "dram_bytes_write": val(r, "dram__bytes_write.sum", True),
        "dram_throughput_pct": val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "tensor_pipe_active_pct": val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        "issue_active_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "warp_instructions": val(r, "smsp__inst_executed.sum"),
        "registers_per_thread": val(r, "launch__registers_per_thread"),
        "shared_wavefronts": val(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
        "shared_bank_conflicts": val(r, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
        "warps_active_pct": val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        "sm_clock_mhz": val(r, "smsp__cycles_elapsed.avg.per_second", False),
    })
with open(os.path.join(ROOT, "monocular-depth-estimation-cil_b200", "build_id.txt")) as f:
    build = f.read().strip()
files = json.load(open(os.path.join(ROOT, "monocular-depth-estimation-cil_b200", "build_files.json")))
json.dump({"build_id": build, "source_files": files, "source": os.path.basename(rep), "command": "ncu --set full --clock-control none --import-source on",
           "note": note, "launches": launches}, open(out, "w"), indent=1)
print(out, build, len(launches), "launch(es)")
