"""Top stall-sample lines of one kernel of an .ncu-rep (source page; needs -lineinfo + --import-source on).
usage: python scripts/ncu_hot.py rep kernel_regex [launch_index] [cuda|sass] [top_n]"""
import csv
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    idx = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    mode = sys.argv[4] if len(sys.argv) > 4 else "sass"
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", mode, "--csv", "--kernel-name", f"regex:{kern}",
                          "--launch-skip", str(idx), "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
    hdr = rows[hi]
    c = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    data = []
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        try:
            n = float(r[c["# Samples"]])
        except ValueError:
            continue
        if n <= 0:
            continue
        st = sorted(((float(r[c[s]] or 0), s[6:]) for s in stalls), reverse=True)[:2]
        data.append((n, r[c["Source"]].strip()[:120], " ".join(f"{s}:{int(v)}" for v, s in st if v > 0)))
    tot = sum(d[0] for d in data)
    print(f"# {rows[0][1][:100] if rows and len(rows[0]) > 1 else kern}  total samples {int(tot)}")
    for n, src, st in sorted(data, reverse=True)[:top]:
        print(f"{100 * n / tot:5.1f}%  {src:120s} {st}")


if __name__ == "__main__":
    main()
