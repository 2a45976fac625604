for a in 9 5 1; do
  for w in C4 C3 C1; do
    B200CTC_LIB=pytorch_end2end_speech_recognition_b200/lib/libb200ctc_ab$a.so python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ablate $a $w lattice %.4f softmax %.4f step %.4f'%(d['roofline']['kernel_ms']['lattice_and_cost_sum'], d['roofline']['kernel_ms']['softmax_rows'], d['ms_per_step']))"
  done
done
