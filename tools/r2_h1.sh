#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SDIMB_HOST_PROFILE=1 timeout 300 python tools/probe_e2e.py > gpurun_out/h1_e2e.txt 2>&1
grep -v Warning gpurun_out/h1_e2e.txt | tail -14
