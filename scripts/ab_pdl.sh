for v in 0 1; do
  export MOBODY_PDL=$v
  echo "=== MOBODY_PDL=$v" >> gpurun_out/ab_pdl.log
  python - >> gpurun_out/ab_pdl.log 2>&1 <<'PY'
import bench, torch
import mobody_b200 as mb
dev = torch.device("cuda:0")
for batch, s, a in ((128, None, None), (4096, 27, 8)):
    r = bench.gpu_train_rate(mb, dev, batch, steps=300 if batch == 128 else 100, s_dim=s, a_dim=a)
    print(batch, r)
PY
done
