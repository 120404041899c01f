#!/bin/bash
# Runs the GPU test groups in separate processes (a trapped kernel poisons its CUDA context) and logs to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name" ; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary.txt; tail -5 gpurun_out/$name.log; }
PT="python -m pytest -q --maxfail=200 -p no:cacheprovider --timeout 300 --timeout-method=thread -m gpu"
run ops $PT tests/test_gpu_ops.py
run gemm_tc $PT tests/test_gpu_gemm_tc.py
GCT_B200_SIMT_GEMM=1 run model_simt $PT tests/test_gpu_model.py
run sampling_fp32 $PT tests/test_gpu_sampling.py -k "not bf16"
run model_tc $PT tests/test_gpu_model.py -k "bf16"
run sampling_bf16 $PT tests/test_gpu_sampling.py -k "bf16"
run smoke python -c "import __graft_entry__ as g; g.smoke()"
cat gpurun_out/summary.txt
