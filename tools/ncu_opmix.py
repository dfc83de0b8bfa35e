"""Opcode mix per source region of one kernel from an .ncu-rep source page (first captured launch; SASS rows are
attributed to the source line they are listed under).
usage: python tools/ncu_opmix.py <file.ncu-rep> <source-file-substring> name:first-last [name:first-last ...]"""
import collections, csv, io, subprocess, sys


def main(path, fsub, regions):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    cur, hdr, seen, name = None, None, set(), None
    mix = collections.defaultdict(collections.Counter)
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1]
            if cur in seen:
                break
            continue
        if r[0] == "Line No":
            hdr = r; seen.add(cur); ii = hdr.index("Instructions Executed"); continue
        if hdr is None or len(r) < len(hdr):
            continue
        if r[0].isdigit():
            ln = int(r[0]); name = cur.split("/")[-1]
            if fsub in cur:
                for n, a, b in regions:
                    if a <= ln <= b:
                        name = n
            continue
        if r[0] == "" and r[2].startswith("0x") and r[ii].isdigit():
            toks = r[3].split()
            op = toks[1] if toks[0].startswith("@") else toks[0]
            mix[name][op.split(".")[0]] += int(r[ii])
    tot = sum(sum(v.values()) for v in mix.values())
    ops = [k for k, _ in sum((collections.Counter(v) for v in mix.values()), collections.Counter()).most_common(16)]
    print(f"{'region':22s} inst%  " + " ".join(f"{o[:6]:>6s}" for o in ops))
    for n, v in sorted(mix.items(), key=lambda kv: -sum(kv[1].values())):
        print(f"{n[:22]:22s} {100 * sum(v.values()) / tot:5.1f}  " + " ".join(f"{100 * v[o] / tot:6.2f}" for o in ops))
    print("total", tot)


if __name__ == "__main__":
    regs = []
    for a in sys.argv[3:]:
        n, r = a.split(":"); lo, hi = r.split("-"); regs.append((n, int(lo), int(hi)))
    main(sys.argv[1], sys.argv[2], regs)
