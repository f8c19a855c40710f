# Evidence captures for profiles/ (run under gpurun, one GPU): usage  bash scripts/capture_profiles.sh r01p
# Every ncu pass follows a plain run of the same command that exited 0.
tag=${1:-rXX}
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train"
$B > gpurun_out/${tag}_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_bench_launches.csv $B > gpurun_out/${tag}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_tc_kernel -s 3 -c 1 -o gpurun_out/${tag}_step_tc -f $B > gpurun_out/${tag}_ncu2.log 2>&1
python scripts/train_profile.py 4096 > gpurun_out/${tag}_tp.log 2>&1 || exit 1
ncu --set full --clock-control none -k regex:"wgrad|critic_fwd|policy_bwd|actor_q" -s 7 -c 5 -o gpurun_out/${tag}_train4096 -f python scripts/train_profile.py 4096 > gpurun_out/${tag}_ncu3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 22 --csv --log-file gpurun_out/${tag}_train128_launches.csv python scripts/train_profile.py 128 > gpurun_out/${tag}_ncu4.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 22 --csv --log-file gpurun_out/${tag}_train4096_launches.csv python scripts/train_profile.py 4096 > gpurun_out/${tag}_ncu5.log 2>&1
