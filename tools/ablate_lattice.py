"""Developer tool: time the lattice kernel with parts of the inner loop compiled out (-DB200CTC_ABLATE=n;
results are WRONG in those builds, only the timing is of interest).  Build the variants on the CPU box
(`--build`), run on the GPU box (no arguments).
  1 no phase-1 scratch stores   2 constant emissions (no gathers)   3 no renormalisation (no max tree)
  4 no neighbour shuffles       5 no posterior at all
  8 phase 1 runs the recursion twice per frame (independent duplicate): latency- or throughput-bound?
  9 the helper warps skip the per-symbol reduction (how much does their work slow the lattice warps?)
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pytorch_end2end_speech_recognition_b200 import build as b  # noqa: E402

VARIANTS = [int(a) for a in os.environ.get('ABLATE_VARIANTS', '0,2,5,8').split(',')]


def lib_of(n):
    return os.path.join(b.LIB_DIR, "libb200ctc_ablate%d.so" % n)


if "--build" in sys.argv:
    for n in VARIANTS:
        b.build_library(extra_flags=["-DB200CTC_ABLATE=%d" % n], lib_path=lib_of(n))
    sys.exit(0)
for n in VARIANTS:
    env = dict(os.environ, B200CTC_LIB=lib_of(n))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "20", "--warmup", "5",
                          "--no-cpu-baseline"] + sys.argv[1:], env=env, capture_output=True, text=True)
    try:
        d = json.loads(out.stdout.strip().splitlines()[-1])
        print("ablate %d: lattice %.4f ms  step %.4f ms" % (n, d["roofline"]["kernel_ms"]["lattice_and_cost_sum"], d["ms_per_step"]))
    except Exception as e:  # noqa: BLE001
        print("ablate %d: failed (%s) %s" % (n, e, out.stderr[-300:]))
