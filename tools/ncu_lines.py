"""Per-source-line instruction / stall-sample breakdown of one kernel from an .ncu-rep (no GPU needed).
usage: python tools/ncu_lines.py <file.ncu-rep> [top_n]"""
import collections, csv, io, subprocess, sys


def load(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    lines, cur, hdr = [], None, None
    for r in rows:
        if r and r[0] == "File Path":
            cur = r[1].split("/")[-1]; continue
        if r and r[0] == "Line No":
            hdr = r; ii = hdr.index("Instructions Executed"); si = hdr.index("# Samples"); continue
        if r and hdr and r[0] not in ("", "Function Name"):
            try:
                lines.append((cur, int(r[0]), r[1], int(r[ii]), int(r[si])))
            except ValueError:
                pass
    return lines


def sass_stats(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
    by_op, st = collections.Counter(), collections.Counter()
    tot = 0
    for r in rows[2:]:
        if len(r) < len(hdr) or r[ix["Instructions Executed"]] == "Instructions Executed":
            continue
        src = r[ix["Source"]].strip().split()
        op = (src[1] if src[0].startswith("@") else src[0]).split(".")[0]
        n = int(r[ix["Instructions Executed"]]); tot += n; by_op[op] += n
        for h in hdr:
            if h.startswith("stall_") and "Not Issued" not in h:
                st[h] += int(r[ix[h]])
    return tot, by_op, st


if __name__ == "__main__":
    path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    tot, by_op, st = sass_stats(path)
    print("warp instructions", tot)
    print("  ".join(f"{o}:{100 * n / tot:.1f}%" for o, n in by_op.most_common(14)))
    ts = sum(st.values())
    print("  ".join(f"{h[6:]}:{100 * n / ts:.1f}%" for h, n in st.most_common(8)))
    lines = load(path)
    ti, tsm = sum(l[3] for l in lines), sum(l[4] for l in lines)
    print(f"{'file':16s} line  inst%  samp%  source")
    for l in sorted(lines, key=lambda l: -l[4])[:top]:
        print(f"{l[0][:16]:16s} {l[1]:4d} {100 * l[3] / ti:5.1f}  {100 * l[4] / tsm:5.1f}  {l[2][:110]}")
