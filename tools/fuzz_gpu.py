"""CUDA path against the fp64 oracle on random buildable version strings (tests/test_gpu_parity.py: fuzz_gpu).
    python tools/fuzz_gpu.py [n=300] [seed=21] [--sizes | --feature]      (--sizes: random frame size, batch, pass size, pair selection too)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import test_gpu_parity as T
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 21
t0 = time.time()
sizes = "--sizes" in sys.argv
if sizes:
    sys.argv.remove("--sizes")
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 21
if "--feature" in sys.argv:
    sys.argv.remove("--feature")
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 21
    failures = T.fuzz_gpu_features(n, seed, log=lambda s: print(s, flush=True))
    print("seed %d: mode='feature' of %d version strings, %d failures, %.0f s" % (seed, n, failures, time.time() - t0))
    sys.exit(min(failures, 100))
worst, failures = (T.fuzz_gpu_sizes if sizes else T.fuzz_gpu)(n, seed, log=lambda s: print(s, flush=True))
print("seed %d: %d version strings, %d failures, worst |gpu - oracle64| = %.2f of the bar (2e-5 + 4.9e-4 max|ref|, inside the "
      "north-star 1e-4 + 1e-3 |ref|), %.0f s" % (seed, n, failures, worst, time.time() - t0))
sys.exit(min(failures, 100))
