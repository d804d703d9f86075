"""Developer probe: fixed per-tile cost of the GEMM core (K = 64 -> one k block) and its epilogue variants."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SWEEP = os.path.join(ROOT, "scripts", "gemm_sweep.py")
for shp in [(8192, 8192, 64), (8192, 8192, 512), (3200, 768, 64), (3200, 768, 2048)]:
    for bn in ("256", "128"):
        env = dict(os.environ, LOCOV_B200_CM="1", LOCOV_B200_CN="1", LOCOV_B200_BN=bn)
        r = subprocess.run([sys.executable, SWEEP, "child"] + [str(x) for x in shp], env=env, capture_output=True, text=True)
        print(shp, "BN", bn, (r.stdout.strip().splitlines() or [r.stderr[-200:]])[-1][:130], flush=True)
