# round-2 ncu captures (a gpurun call returns at most 64 MiB): bash scripts/ncu_round2.sh
# every profiled command first runs once without ncu in the same call (exit code checked by &&)
set -x
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-sub"
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $B > gpurun_out/r02_plain_cfg2.log 2>&1 &&
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_cfg2_launches.csv $B > gpurun_out/r02_ncu_l.log 2>&1
timeout 400 $NCU -k regex:aa_stream -s 3 -c 1 -o gpurun_out/prof_r02_cfg2_stream $B > gpurun_out/r02_ncu_f2.log 2>&1
timeout 300 $B --config cfg3 > gpurun_out/r02_plain_cfg3.log 2>&1 &&
timeout 400 $NCU -k regex:aa_vmma -s 3 -c 1 -o gpurun_out/prof_r02_cfg3_vmma_full $B --config cfg3 > gpurun_out/r02_ncu_f3.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_cfg3_launches.csv $B --config cfg3 > gpurun_out/r02_ncu_l3.log 2>&1
timeout 300 $B --config cfg4 > gpurun_out/r02_plain_cfg4.log 2>&1 &&
timeout 400 $NCU -k regex:aa_tile -s 3 -c 1 -o gpurun_out/prof_r02_cfg4_tile $B --config cfg4 > gpurun_out/r02_ncu_f4.log 2>&1
ls -la gpurun_out/ | grep r02
