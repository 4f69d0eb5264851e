"""Shared-memory wavefronts per SASS instruction of one kernel in an ncu report (source page), largest first.
usage: python scripts/ncu_smem.py report.ncu-rep [kernel-index] [top-n]"""
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
    W = [(int(r[h["L1 Wavefronts Shared"]] or 0), int(r[h["L1 Wavefronts Shared Ideal"]] or 0), int(r[h["Instructions Executed"]] or 0), i, r[h["Source"]].strip())
         for i, r in enumerate(seg["rows"])]
    tot = sum(w[0] for w in W)
    ideal = sum(w[1] for w in W)
    print(seg["name"], "shared wavefronts", tot, "ideal", ideal)
    by_op = {}
    for w, idl, n, i, src in W:
        if w:
            op = src.split()[0] if not src.startswith("@") else src.split()[1]
            a = by_op.setdefault(op, [0, 0, 0])
            a[0] += w; a[1] += idl; a[2] += n
    for op, (w, idl, n) in sorted(by_op.items(), key=lambda kv: -kv[1][0]):
        print(f"  {op:28s} wavefronts {w:12d} ({100 * w / tot:5.1f}%)  ideal {idl:12d}  executed {n:10d}  wavefronts/instr {w / max(n, 1):.2f}")
    print("top instructions:")
    for w, idl, n, i, src in sorted(W, reverse=True)[:topn]:
        print(f"  {i:5d} {w:10d} ideal {idl:10d} x{n:9d}  {src[:90]}")


if __name__ == "__main__":
    main()
