# one ncu --set full capture per class of sweep points below 0.70 that had none yet: bash scripts/ncu_round2_classes.sh
# (each case first runs without ncu in the same call)
set -x
NCU="ncu --set full --clock-control none --import-source on -f"
run() {  # name kernel-regex args...
  local name=$1 rx=$2; shift 2
  timeout 200 python scripts/run_case.py "$@" > gpurun_out/r02_plain_$name.log 2>&1 &&
  timeout 400 $NCU -k regex:$rx -s 2 -c 1 -o gpurun_out/prof_r02_$name python scripts/run_case.py "$@" > gpurun_out/r02_ncu_$name.log 2>&1
}
run u8_bicubic_05  aa_vmma   fwd8 cubic  3 0 1024 1024 512 512 64
run bwd_bicubic_15 aa_stream bwd  cubic  3 0 1024 1024 1536 1536 8
run fwd_bilinear_05_cl aa_stream fwd linear 3 1 1024 1024 512 512 16
run fwd_bicubic_033 aa_stream fwd cubic 3 0 1024 1024 341 341 16
ls -la gpurun_out | grep prof_r02_
