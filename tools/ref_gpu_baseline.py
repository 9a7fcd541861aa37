#!/usr/bin/env python
"""The reference's own GPU sort (oracle/_ref/libreflab.so = SRM/lab.cu compiled unmodified) run on this
GPU as a timing baseline and to confirm on hardware what SURVEY.md could only emulate:
it is launchable for n <= 2^17 only, it mis-sorts low-duplicate inputs from n = 2^10 on (tail
off-by-one, SRM/lab.cu:254,260), and it never terminates on mixed-sign keys (SRM/lab.cu:61,78).
Every case runs in its own process under a timeout (the reference exits on CUDA errors)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASE = r'''
import sys, time
sys.path.insert(0, %r)
import numpy as np, oracle
from b200sort import datagen
import b200sort
dist, log2n = sys.argv[1], int(sys.argv[2])
n = 1 << log2n
keys = datagen.make(dist, n, 1)
want = np.sort(keys)
oracle.ref.order_array(datagen.make(dist, 256, 2))           # warm-up (context, first launches)
t = time.perf_counter(); got = oracle.ref.order_array(keys); ref_ms = (time.perf_counter() - t) * 1e3
ours = keys.copy(); b200sort.order_array(ours.copy())
t = time.perf_counter(); b200sort.order_array(ours); our_ms = (time.perf_counter() - t) * 1e3
wrong = int((got != want).sum())
print(f"{dist:15s} n=2^{log2n:<2d} reference order_array {ref_ms:8.3f} ms  sorted={'yes' if wrong == 0 else 'NO (%%d keys out of place)' %% wrong:28s} | ours order_array {our_ms:7.3f} ms  sorted={'yes' if (ours == want).all() else 'NO'}")
''' % ROOT
def run(dist, log2n, timeout=60):
    try:
        r = subprocess.run([sys.executable, "-c", CASE, dist, str(log2n)], capture_output=True, text=True, timeout=timeout, cwd=ROOT)
        out = (r.stdout.strip().splitlines() or [""])[-1]
        err = [l for l in r.stderr.strip().splitlines() if "GPUassert" in l]
        if r.returncode != 0:
            print(f"{dist:15s} n=2^{log2n:<2d} reference order_array: exit code {r.returncode}  {err[-1] if err else r.stderr.strip()[-200:]}")
        else:
            print(out)
    except subprocess.TimeoutExpired:
        print(f"{dist:15s} n=2^{log2n:<2d} reference order_array: no result after {timeout} s (killed) -- does not terminate")
    sys.stdout.flush()
for log2n in range(8, 18):
    run("lab_rand100", log2n)
for log2n in (8, 9, 10, 11, 12, 14, 16, 17):
    run("uniform_nonneg", log2n)
run("lab_rand100", 18)                  # launch failure: separators_kernel needs 1026 threads per block
if "--dangerous" in sys.argv:
    # recorded once in profiles/r01_reference_gpu_baseline.txt; not repeated by default: the first
    # faults inside a reference kernel, the second leaves a kernel spinning until the timeout kills it
    run("uniform_nonneg", 20)
    run("uniform", 10, timeout=20)      # mixed signs
