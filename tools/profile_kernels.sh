#!/bin/bash
# ncu --set full captures of the path's kernels at BASELINE configs[1] shapes (run under gpurun, one GPU).
set -x
mkdir -p gpurun_out
python tools/profile_target.py fit 20000 > gpurun_out/pk_plain_fit.log 2>&1 || exit 1
for k in gram_cross_kernel colreduce_kernel gram_sym_kernel potf2_inv_kernel trsm_leaf_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/pk_$k \
      python tools/profile_target.py fit 20000 > gpurun_out/pk_ncu_$k.log 2>&1
done
python tools/gemm_bench.py quick > gpurun_out/pk_plain_gemm.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_f64 -s 1 -c 1 -f -o gpurun_out/pk_gemm_f64 \
    python tools/gemm_bench.py quick > gpurun_out/pk_ncu_gemm.log 2>&1
ls -la gpurun_out/pk_*.ncu-rep
