"""Rank source lines of each profiled kernel by stall samples (ncu --page source --csv)."""
import csv, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 18
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
kern, cur, out, hdr = None, None, {}, None
for r in rows:
    if len(r) >= 2 and r[0] == "Function Name":
        kern = r[1][:90]; out.setdefault(kern, [])
    elif len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif len(r) > 8 and r[0] == "Line No":
        hdr = r
    elif len(r) > 8 and r[0].isdigit() and kern:
        try:
            out[kern].append((int(r[7]), int(r[6]) if r[6].isdigit() else 0, cur, int(r[0]), r[1], r))
        except ValueError:
            pass
st_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
for kern, o in out.items():
    tot = sum(x[0] for x in o) or 1; totS = sum(x[1] for x in o) or 1
    print(f"=== {kern}: {tot} warp-instructions, {totS} samples")
    for inst, smp, f, ln, src, r in sorted(o, key=lambda x: -x[1])[:top]:
        stalls = sorted(((int(r[i]) if r[i].isdigit() else 0, hdr[i][6:]) for i in st_cols), reverse=True)[:3]
        print(f"  {f}:{ln:4d} inst {100*inst/tot:5.1f}% samp {100*smp/totS:5.1f}%  {src.strip()[:80]:80s} {stalls}")
