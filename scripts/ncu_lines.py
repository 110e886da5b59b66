"""Per-source-line instruction / stall-sample totals from an .ncu-rep captured with --import-source on.
usage: python scripts/ncu_lines.py rep.ncu-rep [top_n]"""
import csv, io, subprocess, sys, collections

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
fpath, hdr, cur = None, None, None
inst = collections.Counter(); samp = collections.Counter(); text = {}; stalls = collections.Counter()
for r in rows:
    if not r: continue
    if r[0] == "File Path": fpath = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; ii = hdr.index("Instructions Executed"); si = hdr.index("# Samples"); continue
    if hdr is None: continue
    if r[0] != "":
        cur = (fpath, int(r[0])); text[cur] = r[1].strip()
    else:
        try:
            inst[cur] += int(r[ii]); samp[cur] += int(r[si])
            for j, h in enumerate(hdr):
                if h.startswith("stall_") and "Not Issued" not in h and r[j] not in ("", "0"): stalls[h] += int(r[j])
        except (ValueError, IndexError):
            pass
ti, ts = sum(inst.values()), sum(samp.values())
print(f"total warp instructions {ti}, samples {ts}")
print("stall samples: " + ", ".join(f"{k[6:]} {100*v/max(ts,1):.1f}%" for k, v in stalls.most_common(8)))
for k, v in inst.most_common(top):
    print(f"{k[0]:22s}:{k[1]:4d} inst {100*v/ti:5.1f}%  samples {100*samp[k]/max(ts,1):5.1f}%  {text[k][:90]}")
