for s in 0 12 1 8 4 2 16; do
PF_SKIP=$s python bench.py --no-cpu-baseline --steps 100 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('skip=$s', round(d['ms_per_step']*1e3,1), 'single', round(d['single_stream']['ms_per_step']*1e3,1))"
done
