# A/B of library variants on the rollout (build them first: python scripts/build_variant.py NAME -D...):  bash scripts/ab_variants.sh
for v in "$@"; do
  export MOBODY_B200_LIB=$PWD/variants/$v/libmobody_b200.so
  echo "=== $v" >> gpurun_out/ab1.log
  python scripts/tc_debug.py 2>&1 | grep "^bf16x2" >> gpurun_out/ab1.log
  python bench.py --no-cpu-baseline --no-train --steps 20 --warmup 3 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value',d['value'],'kernel_ms',r['kernel_ms'],'frac',r['frac'],'e2e',d['e2e']['value'])" >> gpurun_out/ab1.log 2>&1
done
