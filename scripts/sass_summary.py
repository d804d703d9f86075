"""Per-kernel counts of the SASS instructions that prove the Blackwell paths (tcgen05 MMA / TMEM, TMA, bulk copies, packed math) in the
built library.  usage: python scripts/sass_summary.py [liblocov_b200.so] > profiles/sass_summary.txt"""
import collections
import re
import subprocess
import sys

LIB = sys.argv[1] if len(sys.argv) > 1 else "locov_b200/lib/liblocov_b200.so"
KEYS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "UBLKRED", "SYNCS", "LDGSTS", "ACQBULK", "FFMA2",
        "FMUL2", "FADD2", "F2FP", "REDG", "RED.E", "MUFU.EX2", "ELECT", "CCTL", "MEMBAR.SC.SYS", "MEMBAR.ALL.SYS"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    cur, counts, sizes = None, collections.OrderedDict(), {}
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(.*", "", cur)[:110]
            counts[cur] = collections.Counter()
            sizes[cur] = 0
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        sizes[cur] += 1
        for k in KEYS:
            if op.startswith(k):
                counts[cur][k] += 1
    print(f"# {LIB}: SASS instruction counts per kernel (cuobjdump -sass, sm_100a); only kernels that use one of the listed instructions")
    print("# UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st (TMEM), UTMALDG/UTMASTG = TMA tensor load/store, UBLKCP = cp.async.bulk,")
    print("# SYNCS = mbarrier, LDGSTS = cp.async, FFMA2/FMUL2/FADD2 = packed fp32x2, F2FP = cvt.rn.bf16x2, REDG/RED = vector atomics")
    for k, c in counts.items():
        if not c:
            continue
        print(f"{k}  [{sizes[k]} instr]  " + "  ".join(f"{n}={v}" for n, v in sorted(c.items())))


if __name__ == "__main__":
    main()
