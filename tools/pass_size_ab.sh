for mb in 512 1024 512 1024 342; do
  python bench.py --no-configs --no-cpu-baseline --no-e2e --max-batch $mb 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print($mb, round(d['value']), d['ms_per_step'], d['clocks']['sm_mhz'])"
done
