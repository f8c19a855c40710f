# A/B of library variants on the train step: MOBODY_B200_LIB=variants/<name>/libmobody_b200.so
for v in "$@"; do
  if [ "$v" = "main" ]; then unset MOBODY_B200_LIB; else export MOBODY_B200_LIB=$PWD/variants/$v/libmobody_b200.so; fi
  echo "=== $v" >> gpurun_out/ab_train.log
  python - >> gpurun_out/ab_train.log 2>&1 <<'PY'
import bench, torch
import mobody_b200 as mb
dev = torch.device("cuda:0")
for batch, s, a in ((128, None, None), (4096, 27, 8)):
    r = bench.gpu_train_rate(mb, dev, batch, steps=300 if batch == 128 else 100, s_dim=s, a_dim=a)
    print(batch, r)
PY
done
