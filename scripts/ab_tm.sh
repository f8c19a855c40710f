for v in 64 32; do
  export MOBODY_TRAIN_TM=$v
  echo "=== MOBODY_TRAIN_TM=$v" >> gpurun_out/ab_tm.log
  python - >> gpurun_out/ab_tm.log 2>&1 <<'PY'
import bench, torch
import mobody_b200 as mb
dev = torch.device("cuda:0")
r = bench.gpu_train_rate(mb, dev, 4096, steps=100, s_dim=27, a_dim=8)
print(4096, r)
PY
done
MOBODY_TRAIN_TM=32 python -m pytest tests -x -q -m gpu -k "train_on_rows" >> gpurun_out/ab_tm.log 2>&1
