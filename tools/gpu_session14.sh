#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/s14_parity_scale.jsonl
: > $L
timeout 900 python tools/parity_scale.py --code nr5g:2:384 --impl HLMinstarapproxf32 --frames 16384 --ebn0 0.25 --max-iter 50 2>&1 | tail -1 | tee -a $L
timeout 900 python tools/parity_scale.py --code nr5g:1:384 --impl Aminstarf32 --frames 8192 --ebn0 1.0 --max-iter 50 2>&1 | tail -1 | tee -a $L
timeout 900 python tools/parity_scale.py --code nr5g:1:384 --impl HLAminstarf32 --frames 8192 --ebn0 0.75 --max-iter 50 2>&1 | tail -1 | tee -a $L
timeout 900 python tools/parity_scale.py --code ar4ja:1/2:1024 --impl Phif64 --frames 16384 --ebn0 1.5 --max-iter 100 2>&1 | tail -1 | tee -a $L
