python -m pytest tests/test_gpu_mlp.py tests/test_gpu_train.py -q 2>&1 | tail -15
MMX_MLP_TC_FP32=1 python tools/tc_model_check.py 2>&1 | grep -v "Warn\|detach\|print(" | grep -A3 "fp32:"
env B=4096 PDROP=0.1 MMX_PRECISION=tf32 python tools/quick_bench.py
env B=4096 PDROP=0.1 python tools/quick_bench.py
env B=4096 PDROP=0.1 MMX_MLP_TC_FP32=0 python tools/quick_bench.py
