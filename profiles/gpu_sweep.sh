# C5 sweep points (BASELINE configs[4]): bash profiles/gpu_sweep.sh N TAG "occ occ ..."  -> gpurun_out/sweep_c5_<occ>_<N>gpu_TAG.json
N=${1:-1}
TAG=${2:-r2}
OCCS=${3:-"0.0001 0.001"}
mkdir -p gpurun_out
for OCC in $OCCS; do
  EXTRA="--no-cpu"
  TMO=600
  case $OCC in 0.0001) STEPS="--steps 5 --warmup 3";; 0.001) STEPS="--steps 3 --warmup 1"; EXTRA="$EXTRA --no-e2e --parity-rows 500";; *) STEPS="--steps 1 --warmup 1"; EXTRA="$EXTRA --no-e2e --parity-rows 64"; TMO=540;; esac
  OUT=gpurun_out/sweep_c5_${OCC}_${N}gpu_$TAG
  if [ $N -eq 1 ]; then
    timeout $TMO python bench.py --workload c5 --occupancy $OCC $STEPS $EXTRA > $OUT.json 2> $OUT.err
  else
    timeout $TMO python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --workload c5 --occupancy $OCC $STEPS $EXTRA > $OUT.json 2> $OUT.err
  fi
  echo "c5 occ $OCC N=$N exit $?"; tail -3 $OUT.err | cut -c1-300
  python - $OUT.json <<'PY'
import json,sys
try:
    j=json.load(open(sys.argv[1])); e=j.get("e2e")
    print("  value %.4g frames/s  ms/step %.2f  e2e %s  events %.3g  parity %s" % (j["value"], j["ms_per_step"], ("%.4g (%.1f ms)" % (e["value"], e["ms_per_step"])) if e else "-", j["config"]["events_total"], j.get("parity")))
    for k,v in sorted(j["kernels"].items(), key=lambda kv:-kv[1]["ms_per_step"])[:8]: print("     %-20s %9.3f ms" % (k, v["ms_per_step"]))
except Exception as ex: print("  failed", ex)
PY
done
