"""Region x stall-reason matrix of one kernel from an .ncu-rep source page (first captured launch).
usage: python tools/ncu_regions.py <file.ncu-rep> <source-file-substring> name:first-last [name:first-last ...]"""
import collections, csv, io, subprocess, sys


def main(path, fsub, regions):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    cur, hdr, seen = None, None, set()
    inst, stall = collections.Counter(), collections.defaultdict(collections.Counter)
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1]
            if cur in seen:   # second launch starts
                break
            continue
        if r[0] == "Line No":
            hdr = r; seen.add(cur); continue
        if hdr is None or not r[0].isdigit() or len(r) < len(hdr):
            continue
        if r[2] != "-":   # SASS rows carry an address; keep the per-source-line aggregate rows only
            continue
        ln = int(r[0]); name = cur.split("/")[-1]
        if fsub in cur:
            for n, a, b in regions:
                if a <= ln <= b:
                    name = n
        inst[name] += int(r[hdr.index("Instructions Executed")] or 0)
        for i, h in enumerate(hdr):
            if h.startswith("stall_") and "Not Issued" not in h:
                stall[name][h[6:]] += int(r[i] or 0)
    ti = sum(inst.values()); ts = sum(sum(v.values()) for v in stall.values())
    reasons = [k for k, _ in sum((collections.Counter(v) for v in stall.values()), collections.Counter()).most_common(9)]
    print(f"{'region':26s} inst%  samp%  " + " ".join(f"{r[:9]:>9s}" for r in reasons))
    for n, _ in sorted(inst.items(), key=lambda kv: -sum(stall[kv[0]].values())):
        s = sum(stall[n].values())
        print(f"{n[:26]:26s} {100 * inst[n] / ti:5.1f}  {100 * s / ts:5.1f}  " + " ".join(f"{100 * stall[n][r] / ts:9.1f}" for r in reasons))
    print("total warp instructions", ti, "samples", ts)


if __name__ == "__main__":
    regs = []
    for a in sys.argv[3:]:
        n, rng = a.split(":"); lo, hi = rng.split("-"); regs.append((n, int(lo), int(hi)))
    main(sys.argv[1], sys.argv[2], regs)
