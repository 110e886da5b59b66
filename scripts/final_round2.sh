# final measurements of round 2 -> gpurun_out/ (copied into profiles/ afterwards): bash scripts/final_round2.sh
set -x
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r02_final_gputest.log 2>&1; tail -3 gpurun_out/r02_final_gputest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_final_smoke.log 2>&1; tail -2 gpurun_out/r02_final_smoke.log
timeout 900 python bench.py > gpurun_out/bench_r02_final.json 2> gpurun_out/bench_r02_final.err; tail -2 gpurun_out/bench_r02_final.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_r02_final_reference.json 2>> gpurun_out/bench_r02_final.err
timeout 300 python bench.py --config cfg3 > gpurun_out/bench_r02_final_cfg3.json 2>> gpurun_out/bench_r02_final.err
timeout 300 python bench.py --config cfg4 > gpurun_out/bench_r02_final_cfg4.json 2>> gpurun_out/bench_r02_final.err
timeout 600 python scripts/bench_sweep.py > gpurun_out/r02_sweep_1gpu.txt 2>&1
timeout 300 python scripts/strong_probe.py > gpurun_out/r02_strong_probe.txt 2>&1
timeout 300 python scripts/redo_cost_probe.py > gpurun_out/r02_drain_cost.txt 2>&1
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-sub"
timeout 300 $B --config cfg3 > gpurun_out/r02_plain_cfg3.log 2>&1 &&
timeout 400 ncu --set full --clock-control none --import-source on -f -k regex:aa_vmma -s 3 -c 1 -o gpurun_out/prof_r02_cfg3_vmma_full $B --config cfg3 > gpurun_out/r02_ncu_f3.log 2>&1
ls -la gpurun_out | grep -E "r02_final|bench_r02|r02_sweep|r02_strong"
