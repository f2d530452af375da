"""Summarise an ncu report per CUDA source line: python -m graphaligner_b200.tools.ncu_lines report.ncu-rep [topN]
(needs -lineinfo at compile time and --import-source on at capture time)."""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = None
    cur_file = ""
    lines = []
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        elif len(r) > 8 and r[0] == "Line No":
            hdr = r
        elif hdr and len(r) > 8 and r[0] not in ("", "Line No"):
            try:
                samples = float(r[6])
                inst = float(r[7])
                tinst = float(r[8])
            except ValueError:
                continue
            lines.append((samples, inst, tinst, cur_file, r[0], r[1].strip()[:110]))
    ts = sum(l[0] for l in lines) or 1
    ti = sum(l[1] for l in lines) or 1
    print("total samples %.0f, warp instructions %.0f, thread instructions %.0f, lanes/instr %.2f" % (ts, ti, sum(l[2] for l in lines), sum(l[2] for l in lines) / ti))
    print("%6s %6s %6s  %s" % ("smpl%", "inst%", "lanes", "line"))
    for l in sorted(lines, key=lambda x: -x[0])[:top]:
        print("%6.2f %6.2f %6.1f  %s:%s  %s" % (100 * l[0] / ts, 100 * l[1] / ti, l[2] / l[1] if l[1] else 0, l[3], l[4], l[5]))


if __name__ == "__main__":
    main()
