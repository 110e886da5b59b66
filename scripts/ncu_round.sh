# ncu captures of one round, in two parts (a gpurun call returns at most 64 MiB): bash scripts/ncu_round.sh a|b
set -x
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
NCU="ncu --set full --clock-control none --import-source on -f"
if [ "$1" = "a" ]; then
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_cfg2_launches.csv $B > gpurun_out/ncu_l.log 2>&1
timeout 400 $NCU -k regex:aa_stream -s 3 -c 1 -o gpurun_out/prof_r01_cfg2_final $B > gpurun_out/ncu_f2.log 2>&1
timeout 400 $NCU -k regex:aa_stream -s 3 -c 1 -o gpurun_out/prof_r01_cfg3_final $B --config cfg3 --images 16 > gpurun_out/ncu_f3.log 2>&1
timeout 400 $NCU -k regex:aa_tile -s 3 -c 1 -o gpurun_out/prof_r01_cfg4_final $B --config cfg4 > gpurun_out/ncu_f4.log 2>&1
else
timeout 300 $NCU -k regex:aa_band -s 2 -c 1 -o gpurun_out/prof_r01_band_l075 python scripts/run_case.py fwd linear 3 0 1024 1024 768 768 > gpurun_out/ncu_f5.log 2>&1
timeout 300 $NCU -k regex:aa_tile -s 2 -c 1 -o gpurun_out/prof_r01_tile_c2 python scripts/run_case.py fwd cubic 3 0 1024 1024 2048 2048 > gpurun_out/ncu_f6.log 2>&1
timeout 300 $NCU -k regex:aa_stream -s 2 -c 1 -o gpurun_out/prof_r01_stream_c025 python scripts/run_case.py fwd cubic 1 0 1024 1024 256 256 > gpurun_out/ncu_f7.log 2>&1
fi
ls -la gpurun_out/ | tail -8
