export SNB_TC_CG2=1
for d in 0 2 3 4 5 6; do echo "--- stagger $d"; SNB_TC_STAGGER=$d timeout 120 python tools/time_decoder.py 2>&1 | tail -2; done
echo "--- cg1 stagger 4"; SNB_TC_CG2=0 SNB_TC_STAGGER=4 timeout 120 python tools/time_decoder.py 2>&1 | tail -2
SNB_TC_STAGGER=4 timeout 240 python -m pytest tests/test_gpu_bf16.py -x -q 2>&1 | tail -3
SNB_TC_STAGGER=4 timeout 120 python tools/trace_pipeline.py fwd 2>&1 | tail -22
