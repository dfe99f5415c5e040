for i in 1 2; do
for cfg in "A 0" "B 1" "A 1"; do set -- $cfg
NETCUDA_PDL=$2 NETCUDA_LIB_DIR=$PWD/vit-fpga_b200/lib$( [ $1 = A ] && echo _A ) timeout 150 python bench.py --steps 15 --warmup 4 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1 pdl=$2', round(d['value']), round(d['ms_per_step'],3), d['clocks']['sm_mhz'])"
done; done
