for cfg in "32 3" "32 4" "32 2" "64 2" "64 3" "16 4" "16 6" "48 3"; do
  set -- $cfg
  python bench.py --quick --no-cpu --steps 20 --warmup 5 --batch $1 --inflight $2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('batch $1 inflight $2 value %.0f e2e %.0f ms/step %.3f serial %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['ms_per_step_profiled_serial']))"
done
