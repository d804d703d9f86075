"""Developer probe: where does the GEMM core's time go?  Runs the 8192x8192x2048 and 3200x768x2048 GEMMs with the
MMAs skipped (pure TMA ingest) and with the TMA loads skipped (pure tensor pipe + operand reads), for two tile widths,
and profiles cuBLAS' kernel configuration for the same shape.  usage (GPU box): python scripts/gemm_probe.py"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SWEEP = os.path.join(ROOT, "scripts", "gemm_sweep.py")
for shp in [(8192, 8192, 2048), (3200, 768, 2048)]:
    for bn in ("256", "128"):
        for dbg in ("0", "1", "2"):
            env = dict(os.environ, LOCOV_B200_CM="1", LOCOV_B200_CN="1", LOCOV_B200_BN=bn, LOCOV_B200_DEBUG=dbg)
            r = subprocess.run([sys.executable, SWEEP, "child"] + [str(x) for x in shp], env=env, capture_output=True, text=True)
            print(shp, "BN", bn, "debug", dbg, (r.stdout.strip().splitlines() or [r.stderr[-200:]])[-1][:120], flush=True)
