#!/bin/bash
# Developer aid (GPU box): run the GPU test-suite and the bench for the named workloads.
#   tools/gpu_round.sh TAG [--no-tests] [workloads...]
TAG=$1; shift
TESTS=1
if [ "$1" == "--no-tests" ]; then TESTS=0; shift; fi
WL=${@:-C3}
mkdir -p gpurun_out
if [ $TESTS == 1 ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
  echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
  tail -3 gpurun_out/${TAG}_pytest.log
fi
for w in $WL; do
  timeout 300 python bench.py --workload $w --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_$w.log 2>> gpurun_out/${TAG}_bench.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench_$w.log").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("$w ms/step %.4f  value %.1f M  lattice %.4f ms softmax %.4f ms  frac %.3f whole %.3f  e2e %.4f ms host %.3f" % (
        d["ms_per_step"], d["value"]/1e6, r["kernel_ms"]["lattice_and_cost_sum"], r["kernel_ms"]["softmax_rows"], r["frac"],
        r["whole_step"]["frac"], d["e2e"]["ms_per_step"], d["host"]["wall_ms_per_step"]))
except Exception as e:
    print("$w failed", e)
PY
done
