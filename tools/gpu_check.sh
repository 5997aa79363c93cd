# one GPU box call: parity tests, smoke, bench, then the ncu launch list of the same bench command
nvidia-smi -L
python -m pytest tests -x -q -m gpu 2>&1 | tail -5
python __graft_entry__.py smoke 2>&1 | tail -2
python tools/quick_bench.py 10 12 14 16 18 20 22 2>&1 | tail -9
BPGPU_PROFILE=1 python tools/quick_bench.py 14 2>&1 | grep "bpgpu msm" | tail -1
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
