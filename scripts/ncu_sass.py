"""Per-instruction view of one kernel of an .ncu-rep (SASS source page): totals, stall mix, instruction share per
1 KB code block and the hottest instructions.  usage: python scripts/ncu_sass.py rep [top_n]"""
import collections
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h = rows[1]
    c = {k: i for i, k in enumerate(h)}
    data = rows[2:]

    def f(r, k):
        try:
            return float(r[c[k]])
        except (ValueError, IndexError):
            return 0.0
    tot = sum(f(r, '# Samples') for r in data)
    ti = sum(f(r, 'Instructions Executed') for r in data)
    print('kernel', rows[0][1][:80])
    print('total samples', tot, 'warp-instr executed', ti)
    stalls = [k for k in h if k.startswith('stall_') and 'Not Issued' not in k]
    agg = {s: sum(f(r, s) for r in data) for s in stalls}
    print(sorted(((round(v / tot * 100, 1), k) for k, v in agg.items()), reverse=True)[:9])
    b = collections.defaultdict(lambda: [0, 0])
    for r in data:
        a = int(r[c['Address']][-5:], 16)
        b[a >> 10][0] += f(r, 'Instructions Executed')
        b[a >> 10][1] += f(r, '# Samples')
    for k in sorted(b):
        if b[k][0] > ti * 0.01 or b[k][1] > tot * 0.01:
            print(hex(k << 10), int(b[k][0]), f"{b[k][0] / ti * 100:.1f}% instr", f"{b[k][1] / tot * 100:.1f}% samples")
    for r in sorted(data, key=lambda r: -f(r, '# Samples'))[:top_n]:
        st = sorted(((f(r, s), s[6:]) for s in stalls), reverse=True)[:2]
        print(f"{f(r, '# Samples') / tot * 100:5.1f}% {r[c['Address']][-5:]} exec={int(f(r, 'Instructions Executed')):9d} {r[c['Source']][:70]:70s} {st[0][1]}:{int(st[0][0])} {st[1][1]}:{int(st[1][0])}")


if __name__ == "__main__":
    main()
