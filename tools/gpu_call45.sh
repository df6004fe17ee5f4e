#!/bin/bash
# round 2, call 45: posterior prediction timings at C2-C4
mkdir -p gpurun_out
(timeout 600 python tools/predict_rate.py 2> gpurun_out/r02_c45_predict.err | grep "^{") > gpurun_out/r02_c45_predict_rate.jsonl
