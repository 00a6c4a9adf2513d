#!/bin/bash
# final validation + profile captures in one call
bash tools/gpu_final.sh
sed -i 's/cap sorted "interp_pipe_kernel" 1/cap sorted "interp_kernel" 1/; s/cap tilesorted "interp_pipe_kernel" 1/cap tilesorted "interp_kernel" 1/' tools/gpu_profiles.sh
bash tools/gpu_profiles.sh
