#!/bin/bash
# final validation + profile captures in one call
bash tools/gpu_final.sh
bash tools/gpu_profiles.sh
