import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu_parity.py", "-x", "-q", "-m", "gpu", "-k", "fallback"], cwd=ROOT, capture_output=True, text=True)
print("\n".join(l for l in r.stdout.splitlines() if "out of range" not in l)[-3000:])
