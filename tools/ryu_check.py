"""Check spk_format.cuh's py_repr (host build) against Python's repr on a few million doubles: random bit patterns,
count/length ratios like the matrix holds, integers, powers of two and ten, subnormals, boundaries."""
import struct, subprocess, sys
import numpy as np
rng = np.random.default_rng(1)
vals = []
vals += [0.0, -0.0, 1.0, -1.0, 0.1, 0.5, 1e-4, 9.999e-5, 1e-5, 1e16, 9999999999999998.0, 1e15, 123456789012345678.0,
         5e-324, 2.2250738585072014e-308, 1.7976931348623157e308, float("inf"), float("-inf"), float("nan"),
         1e22, 1e23, 9007199254740993.0, 2.0 ** -44, 0.3, 2.5, 1e-7, 123.456, 1e100, 1e-100, 4.35, 0.000123]
vals += [2.0 ** e for e in range(-1074, 1024, 7)] + [10.0 ** e for e in range(-320, 309)]
vals += list(rng.integers(0, 2**63, 1_500_000, dtype=np.uint64).view(np.float64))
vals += list(rng.integers(0, 2**63, 200_000, dtype=np.uint64).view(np.float64) * -1)
c = rng.integers(0, 50000, 1_500_000); l = rng.integers(1, 900_000_000, 1_500_000)
vals += list(c / l)
vals += list(rng.integers(0, 10**9, 200_000).astype(np.float64))
vals += list(rng.random(300_000)) + list(rng.random(100_000) * 1e-300)
arr = np.array(vals, dtype=np.float64)
out = subprocess.run([sys.argv[1] if len(sys.argv) > 1 else "/tmp/ryu_check"], input=arr.tobytes(), capture_output=True).stdout.decode().split("\n")
bad = 0
for v, s in zip(arr.tolist(), out):
    if repr(v) != s:
        bad += 1
        if bad < 10:
            print("MISMATCH", repr(v), s)
print("checked", len(arr), "mismatches", bad)
sys.exit(1 if bad else 0)
