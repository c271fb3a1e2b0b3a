"""Source-line hot spots of one kernel in an .ncu-rep (needs -lineinfo + --import-source on):
    python profiles/ncu_hotspots.py report.ncu-rep kernel_regex [n_lines]
Aggregates the warp-stall samples and executed warp instructions of `ncu --page source --print-source cuda,sass` per CUDA line."""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur, hdr, data, seen = None, None, [], None
for r in rows:
    if not r:
        continue
    if r[0] == "Kernel Name":
        if seen is not None and r[1] != seen:
            break                     # first matching kernel only
        seen = r[1]
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        if seen is None:
            seen = r[1]
        elif r[1] != seen:
            break
        continue
    if r[0] == "Line No":
        hdr = r; si = hdr.index("# Samples"); ie = hdr.index("Instructions Executed"); continue
    if r[0] != "" and hdr:
        try:
            data.append((int(r[si]), int(r[ie]), cur, int(r[0]), r[1].strip()[:110]))
        except Exception:
            pass
tot = sum(d[0] for d in data) or 1; toti = sum(d[1] for d in data) or 1
print(f"kernel {seen}: {tot} stall samples, {toti} warp instructions")
for d in sorted(data, reverse=True)[:top]:
    print(f"{100 * d[0] / tot:5.1f}% samples {100 * d[1] / toti:5.1f}% instr  {d[2]}:{d[3]}  {d[4]}")
