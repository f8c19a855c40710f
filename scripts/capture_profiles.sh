# Evidence captures for profiles/ (run under gpurun, one GPU): usage  bash scripts/capture_profiles.sh r02g
# Every ncu pass follows a plain run of the same command that exited 0.
tag=${1:-rXX}
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train"
$B > gpurun_out/${tag}_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_bench_launches.csv $B > gpurun_out/${tag}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_tc_kernel -s 3 -c 1 -o gpurun_out/${tag}_step_tc -f $B > gpurun_out/${tag}_ncu2.log 2>&1
python scripts/ncu_select.py gpurun_out/${tag}_step_tc.ncu-rep gpurun_out/${tag}_step_tc_ncu_full_selected.csv
python scripts/train_profile.py 4096 > gpurun_out/${tag}_tp.log 2>&1 || exit 1
# one update = 24 launches at batch 4096; skip the first three updates, capture the GEMM tiles + helpers of one
ncu --set full --clock-control none -k regex:"gemm_kernel|head_bwd|rowdot|adam_kernel|pack_b" -s 66 -c 24 -o gpurun_out/${tag}_train4096 -f python scripts/train_profile.py 4096 > gpurun_out/${tag}_ncu3.log 2>&1
python scripts/ncu_select.py gpurun_out/${tag}_train4096.ncu-rep gpurun_out/${tag}_train4096_ncu_full_selected.csv
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^(critic|actor|policy|wgrad|adam|sample_rows|head_bwd|rowdot|td)_" -s 30 -c 22 --csv --log-file gpurun_out/${tag}_train128_launches.csv python scripts/train_profile.py 128 > gpurun_out/${tag}_ncu4.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^(critic|actor|policy|wgrad|adam|sample_rows|head_bwd|rowdot|td|gemm|pack_b)_" -s 81 -c 54 --csv --log-file gpurun_out/${tag}_train4096_launches.csv python scripts/train_profile.py 4096 > gpurun_out/${tag}_ncu5.log 2>&1
# dynamics fitting step: launch list of two mini-batches (43 launches each) + full capture of one mini-batch's GEMM tiles
python scripts/fit_profile.py > gpurun_out/${tag}_fit.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^(prep|reparam|combine|fake|lossgrad|loss|finish|dfake|dza|do3|adam|gemm)_kernel" -s 86 -c 86 --csv --log-file gpurun_out/${tag}_dynfit_launches.csv python scripts/fit_profile.py > gpurun_out/${tag}_ncu6.log 2>&1
ncu --set full --clock-control none -k regex:"^gemm_kernel" -s 64 -c 32 -o gpurun_out/${tag}_dynfit -f python scripts/fit_profile.py > gpurun_out/${tag}_ncu7.log 2>&1
python scripts/ncu_select.py gpurun_out/${tag}_dynfit.ncu-rep gpurun_out/${tag}_dynfit_ncu_full_selected.csv
