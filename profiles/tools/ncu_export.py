"""Exports an `ncu --set full` capture as the raw metric table the profile notes quote (one row per metric: name, unit,
value per captured launch), and optionally the DRAM traffic record bench.py reads.

    python profiles/tools/ncu_export.py gpurun_out/r02b_fast.ncu-rep profiles/r02_fast_raw.csv [--traffic profiles/r02_traffic.json]

The CSV is `ncu -i <rep> --page raw --csv` transposed; nothing is typed by hand."""
import csv
import io
import json
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, launches = rows[0], rows[1], rows[2:]
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit"] + ["launch_%d" % i for i in range(len(launches))])
    for c, name in enumerate(hdr):
        w.writerow([name, units[c]] + [r[c] for r in launches])
print("wrote %s: %d metrics x %d launch(es)" % (out, len(hdr), len(launches)))
if "--traffic" in sys.argv:
    path = sys.argv[sys.argv.index("--traffic") + 1]

    def val(name):
        c = hdr.index(name)
        v = float(launches[0][c].replace(",", ""))
        u = units[c].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}[u]

    rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
    rec = {"kernel": launches[0][hdr.index("Kernel Name")].split("(")[0], "dram_bytes_per_launch": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr,
           "gpu_time_ms": float(launches[0][hdr.index("gpu__time_duration.sum")].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[units[hdr.index("gpu__time_duration.sum")]],
           "source": "%s via profiles/tools/ncu_export.py (ncu --set full --clock-control none, one launch of bench.py's config 2)" % rep.split("/")[-1], "raw": out}
    with open(path, "w") as f:
        json.dump(rec, f, indent=1)
    print("wrote", path, rec)
