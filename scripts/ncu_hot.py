"""Hot spots of one kernel in an ncu report: per-instruction sampling grouped around barrier waits.
usage: python scripts/ncu_hot.py report.ncu-rep [kernel-index] [top-n]"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    kidx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    segs, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            segs.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None:
            cur["rows"].append(r)
    seg = segs[kidx]
    h = {k: i for i, k in enumerate(seg["hdr"])}
    S = [int(r[h["# Samples"]]) for r in seg["rows"]]
    tot = sum(S)
    print(seg["name"], "samples", tot, "instructions", len(S))
    stalls = [k for k in seg["hdr"] if k.startswith("stall_") and "Not Issued" not in k]
    agg = {k: sum(int(r[h[k]]) for r in seg["rows"]) for k in stalls}
    print("stall totals:", sorted(agg.items(), key=lambda kv: -kv[1])[:8])
    top = sorted(range(len(S)), key=lambda i: -S[i])[:topn]
    for i in sorted(top):
        r = seg["rows"][i]
        st = sorted(((int(r[h[k]]), k[6:]) for k in stalls), reverse=True)[:2]
        print(f"{i:5d} {S[i]:6d} {100.0 * S[i] / tot:5.1f}%  x{r[h['Instructions Executed']]:>9s}  {r[h['Source']].strip()[:72]:72s} {st}")


if __name__ == "__main__":
    main()
