"""Rank source lines of a profiled kernel by executed warp-instructions (ncu --page source --csv)."""
import csv, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
kern, cur, out = None, None, {}
for r in rows:
    if len(r) >= 2 and r[0] == "Function Name":
        kern = r[1]; out.setdefault(kern, {})
    elif len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif len(r) > 8 and r[0].isdigit() and kern:
        try:
            key = (cur, int(r[0]), r[1].strip()[:95])
            out[kern][key] = out[kern].get(key, 0) + int(r[7])
        except ValueError:
            pass
for kern, o in out.items():
    if pat not in kern:
        continue
    tot = sum(o.values()) or 1
    print(f"=== {kern[:100]}: {tot} warp-instructions")
    for (f, ln, src), inst in sorted(o.items(), key=lambda kv: -kv[1])[:top]:
        print(f"  {f}:{ln:4d} {100*inst/tot:5.1f}%  {src}")
