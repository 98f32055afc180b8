python -m pytest tests -m gpu -x -q -k "sift or vo_frames_with_unique or equals_loop" 2>&1 | tail -2
for v in 0 1 0 1; do VO_BLUR_PACK=$v python bench.py --quick --no-cpu --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
r=[x for x in d['roofline_all'] if x['kernel']=='sift_blur_tma_kernel'][0]
print('PACK=$v value %.0f ms/step %.3f serial %.3f blur ms/step %.3f frac %.3f' % (d['value'], d['ms_per_step'], d['ms_per_step_profiled_serial'], r['avg_launch_ms']*r['launches']/d['steps'], r['frac']))"; done
