# usage: bash tools/ab_env.sh VAR v1 v2 ...   -- one short bench.py pass per value of an environment variable
var=$1; shift
for v in "$@"; do env $var=$v python bench.py --quick --no-cpu --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
r={x['kernel']:x for x in d['roofline_all']}
def ms(k): return r[k]['avg_launch_ms']*r[k]['launches']/d['steps'] if k in r else 0
print('$var=$v value %.0f ms/step %.3f serial %.3f | blur %.3f extrema %.3f desc %.3f orient %.3f base %.3f' % (d['value'], d['ms_per_step'], d['ms_per_step_profiled_serial'], ms('sift_blur_tma_kernel'), ms('sift_extrema'), ms('sift_descriptor'), ms('sift_refine_orient'), ms('sift_base_upsample_blur')))"; done
