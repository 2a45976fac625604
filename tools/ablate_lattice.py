"""Developer tool: time the lattice kernel with parts of the inner loop compiled out (-DB200CTC_ABLATE=<bit mask>,
bit n = experiment n; results are WRONG in those builds, only the timing is of interest).  Build the variants on
the CPU box (`--build`), run on the GPU box (no arguments).  ABLATE_VARIANTS: comma-separated experiments, several
joined with '+' in one build (e.g. "0,5,9,5+9,5+9+10").
  1 no phase-1 scratch stores   2 constant emissions (no gathers)   3 no renormalisation (no max tree)
  4 no neighbour shuffles       5 no posterior at all
  8 phase 1 runs the recursion twice per frame (independent duplicate): latency- or throughput-bound?
  7 phase 2: chunk barrier only at every halo exchange (every 16 frames)
  9 the helper warps skip the per-symbol reduction (how much does their work slow the lattice warps?)
  10 the helper warps fetch no records (use with 5)     11 ... fetch but never wait (ends in a launch failure: do not run)
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pytorch_end2end_speech_recognition_b200 import build as b  # noqa: E402

VARIANTS = os.environ.get('ABLATE_VARIANTS', '0,5,9,5+9').split(',')


def mask_of(v):
    return sum(1 << int(x) for x in v.split('+') if int(x) > 0)


def lib_of(v):
    return os.path.join(b.LIB_DIR, "libb200ctc_ablate%s.so" % v.replace('+', '_'))


if "--build" in sys.argv:
    for v in VARIANTS:
        b.build_library(extra_flags=["-DB200CTC_ABLATE=%d" % mask_of(v)], lib_path=lib_of(v))
    sys.exit(0)
for n in VARIANTS:
    env = dict(os.environ, B200CTC_LIB=lib_of(n))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "20", "--warmup", "5",
                          "--no-cpu-baseline"] + sys.argv[1:], env=env, capture_output=True, text=True)
    try:
        d = json.loads(out.stdout.strip().splitlines()[-1])
        print("ablate %s: lattice %.4f ms  step %.4f ms" % (n, d["roofline"]["kernel_ms"]["lattice_and_cost_sum"], d["ms_per_step"]))
    except Exception as e:  # noqa: BLE001
        print("ablate %s: failed (%s) %s" % (n, e, out.stderr[-300:]))
