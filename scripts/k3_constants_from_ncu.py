"""profiles/k3_ncu_constants.json from an ncu --set full capture of the K3 kernel (raw page): the per-launch DRAM traffic and the
per-shift instruction / FP64-flop counts bench.py's roofline objects use, stamped with the commit the capture was taken at."""
import csv, io, json, subprocess, sys
rep, commit, out = sys.argv[1], sys.argv[2], sys.argv[3]
n = int(sys.argv[4]) if len(sys.argv) > 4 else 1 << 20
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units, r = rows[0], rows[1], rows[-1]
def val(name):
    i = hdr.index(name)
    v = float(r[i].replace(",", ""))
    u = units[i].lower()
    scale = {"kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "byte": 1.0}.get(u, 1.0)
    return v * scale
dram = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
inst = val("smsp__inst_executed.sum") * 32 / n
cyc = val("smsp__cycles_elapsed.avg")
op = lambda o: val(f"smsp__sass_thread_inst_executed_op_{o}_pred_on.sum.per_cycle_elapsed") * cyc      # thread instructions per launch
fl = (2 * op("dfma") + op("dmul") + op("dadd")) / n
f32 = (op("ffma") + op("fmul") + op("fadd")) / n
json.dump({"commit": commit, "kernel": r[hdr.index("Kernel Name")] + ", one 1024x1024 launch (compact output form)",
           "dram_bytes_per_launch": dram, "instr_per_point": inst, "fp64_flops_per_point": fl,
           "fp64_instr_per_point": (op("dfma") + op("dmul") + op("dadd")) / n, "fp32_arith_instr_per_point": f32,
           "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
           "duration_us_under_ncu": val("gpu__time_duration.sum") / 1e3 if "gpu__time_duration.sum" in hdr else None,
           "source": "ncu --set full --clock-control none (scripts/prof_k3.py): dram__bytes_read.sum + dram__bytes_write.sum; "
                     "smsp__inst_executed.sum * 32 / shifts; 2 DFMA + DMUL + DADD thread instructions / shifts (per_cycle_elapsed sums x smsp__cycles_elapsed.avg)"},
          open(out, "w"), indent=1)
print(open(out).read())
