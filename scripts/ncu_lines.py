"""Per-source-line instruction and stall-sample shares of one kernel from an .ncu-rep (needs -lineinfo):
python scripts/ncu_lines.py report.ncu-rep [min_pct]"""
import csv, subprocess, sys

def lines(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, data, fname = None, [], ""
    for r in rows:
        if len(r) == 2 and r[0] == "File Path": fname = r[1].split("/")[-1]
        if len(r) > 8 and r[0] == "Line No": hdr = r; continue
        if hdr and len(r) == len(hdr) and r[0].isdigit(): data.append((fname, r))
    ia, iss = hdr.index("Instructions Executed"), hdr.index("# Samples")
    agg = {}
    for f, r in data:
        a = agg.setdefault((f, int(r[0]), r[1].strip()[:90]), [0, 0])
        a[0] += int(r[ia]); a[1] += int(r[iss])
    return agg

if __name__ == "__main__":
    agg = lines(sys.argv[1])
    mn = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
    tot = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
    print("total warp-instructions %d, samples %d" % (tot, ts))
    for k, (n, s) in sorted(agg.items()):
        if n > tot * mn / 100 or s > ts * mn / 100:
            print("%-16s %4d %8.1fM inst %5.1f%% samp %5.1f%%  %s" % (k[0], k[1], n / 1e6, n * 100 / tot, s * 100 / ts, k[2]))
