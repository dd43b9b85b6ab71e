"""Per-source-line summary of an ncu report (needs -lineinfo and --import-source on):
   ncu_lines.py <report.ncu-rep> [top N]   -> instructions executed and stall samples by CUDA source line."""
import collections, csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur = None; per = collections.Counter(); pers = collections.Counter(); src = {}; stall = collections.defaultdict(collections.Counter)
ia = isamp = None; stall_cols = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No":
        hdr = r; ia = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples")
        stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if r[0] != "":
        key = (cur, int(r[0])); src[key] = r[1]
        try:
            per[key] += int(r[ia]); pers[key] += int(r[isamp])
            for i, h in stall_cols:
                if r[i] not in ("", "-", "0"): stall[key][h] += int(r[i])
        except Exception: pass
tot = sum(per.values()) or 1; ts = sum(pers.values()) or 1
print("total warp instructions %d, stall samples %d" % (tot, ts))
allst = collections.Counter()
for k in stall: allst.update(stall[k])
print("stall mix:", ", ".join("%s %.1f%%" % (h[6:], 100.0 * v / ts) for h, v in allst.most_common(8)))
for k, v in pers.most_common(top):
    st = ", ".join("%s %d%%" % (h[6:], 100 * c // max(1, pers[k])) for h, c in stall[k].most_common(2))
    print("%-12s %4d %5.2f%% samp %5.2f%% instr [%s] | %s" % (k[0], k[1], 100.0 * v / ts, 100.0 * per[k] / tot, st, src[k].strip()[:100]))
