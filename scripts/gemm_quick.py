"""Developer probe: default-configuration GEMM timings vs cuBLAS for the shapes of the path."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SWEEP = os.path.join(ROOT, "scripts", "gemm_sweep.py")
for shp in [(3200, 768, 2048), (8192, 768, 2048), (8192, 2048, 768), (32768, 768, 2048), (8192, 8192, 2048), (8192, 8192, 64)]:
    for cfg in [dict(), dict(LOCOV_B200_BN="256"), dict(LOCOV_B200_BN="128")]:
        r = subprocess.run([sys.executable, SWEEP, "child"] + [str(x) for x in shp], env=dict(os.environ, **cfg), capture_output=True, text=True)
        print(shp, cfg, (r.stdout.strip().splitlines() or [r.stderr[-300:]])[-1][:130], flush=True)
