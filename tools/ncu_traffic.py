"""profiles/r02_update_traffic.json from an `ncu --set full` capture of the update kernel:
   python tools/ncu_traffic.py gpurun_out/<report>.ncu-rep STEPS
dram__bytes_read.sum + dram__bytes_write.sum of the (single) captured launch divided by the minibatch steps it ran;
bench.py scales it to its own launch size for the roofline's `traffic` field."""
import csv
import io
import json
import subprocess
import sys

rep, steps = sys.argv[1], int(sys.argv[2])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units, v = rows[0], rows[1], rows[2]
col = {name: i for i, name in enumerate(h)}


def val(name):
    x, u = float(v[col[name]]), units[col[name]]
    return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
out = {"dram_bytes_per_minibatch_step": (rd + wr) / steps, "dram_bytes_read": rd, "dram_bytes_write": wr, "steps_in_capture": steps,
       "kernel": v[col["Kernel Name"]] if "Kernel Name" in col else "spo_update_kernel",
       "source": f"ncu --set full --clock-control none, {rep.split('/')[-1]} (one launch of {steps} minibatch steps; algorithmic 64 x 272 B = 17.4 KB per step)"}
json.dump(out, open("profiles/r02_update_traffic.json", "w"), indent=1)
print(out)
