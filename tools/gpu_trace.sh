#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/trace_chain2.py > gpurun_out/trace2.log 2>&1; echo "trace exit $?"; grep -n "===\|epilogue span\|MMA thread\|producer\|D-encode" gpurun_out/trace2.log
