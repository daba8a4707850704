"""A/B timing of row-wise kernel variants (tools/build_variants.sh rowwise.cu ...) on one GPU."""
import os, subprocess, sys
HERE = os.path.dirname(os.path.abspath(__file__))
for v in sys.argv[1:]:
    env = dict(os.environ, PFN_B200_LIB=os.path.join(HERE, "ubench", "_bin", f"libpfn_{v}.so"))
    r = subprocess.run([sys.executable, os.path.join(HERE, "time_kernels.py"), "row"], env=env, capture_output=True, text=True, timeout=300)
    print(f"{v}: " + " | ".join(l for l in r.stdout.strip().splitlines() if l.startswith("ln")) + r.stderr.strip()[-200:], flush=True)
