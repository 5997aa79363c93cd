nvidia-smi -L
python -m pytest tests/test_gpu_arith.py -x -q -m gpu 2>&1 | tail -15
python -m pytest tests/test_gpu_msm.py -x -q -m gpu 2>&1 | tail -15
python tools/quick_bench.py 2>&1 | tail -20
BPGPU_PROFILE=1 python tools/quick_bench.py 20 2>&1 | grep "bpgpu msm" | tail -3
BPGPU_PROFILE=1 python tools/quick_bench.py 10 2>&1 | grep "bpgpu msm" | tail -2
