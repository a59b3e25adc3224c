#!/bin/bash
# Runs GPU test groups in separate processes with timeouts (a trapped kernel poisons its CUDA
# context, so groups are isolated); logs go to gpurun_out/.  Each argument is "file" or "file@kexpr".
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
i=0
for t in "$@"; do
  i=$((i+1))
  f="${t%%@*}"; k=""; [[ "$t" == *@* ]] && k="${t#*@}"
  name="g${i}"
  echo "=== $name: $f -k '$k'"
  if [ -n "$k" ]; then timeout 400 python -m pytest -x -q -m gpu "$f" -k "$k" -s > gpurun_out/$name.log 2>&1
  else timeout 400 python -m pytest -x -q -m gpu "$f" -s > gpurun_out/$name.log 2>&1; fi
  echo "exit $? ($name)"; tail -n ${TAILN:-15} gpurun_out/$name.log
done
