# sweep of the loss-kernel ring geometry (debug env hooks of yb_loss_fwd_bwd)
for cfg in "2 4 20 2 --unfused" "2 4 20 3 --unfused" "2 4 20 2" "3 3 16 2" "2 4 20 2"; do
  set -- $cfg
  echo -n "stages=$1 ctas=$2 tile=$3 warps=$4 $5: "
  YB_LOSS_STAGES=$1 YB_LOSS_CTAS_PER_SM=$2 YB_LOSS_TILE_CELLS=$3 YB_LOSS_WARPS=$4 python bench.py --steps 20 --warmup 5 --no-cpu-baseline $5 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['frac'])"
done
