#!/bin/bash
set -x
mkdir -p gpurun_out
{
echo "== f32 sorted wide classic"; NINT_PIPE=0 timeout 300 python tools/prof_case.py --case sorted --dtype f32 --points 1e9 --launches 4
echo "== f32 sorted wide pipe"; timeout 300 python tools/prof_case.py --case sorted --dtype f32 --points 1e9 --launches 4
echo "== f32 morton wide classic"; NINT_PIPE=0 timeout 300 python tools/prof_case.py --case morton --dtype f32 --points 1e9 --launches 4
echo "== f32 morton wide pipe"; timeout 300 python tools/prof_case.py --case morton --dtype f32 --points 1e9 --launches 4
echo "== nearest sorted classic"; NINT_PIPE=0 timeout 300 python tools/prof_case.py --case sorted --strategy nearest --points 1e9 --launches 4
echo "== nearest sorted pipe"; timeout 300 python tools/prof_case.py --case sorted --strategy nearest --points 1e9 --launches 4
} > gpurun_out/r2y_times.log 2>&1
grep -v "^+" gpurun_out/r2y_times.log | grep -v "launch 0" | grep -v "launch 1"
