#!/usr/bin/env bash
set -u
bash tools/gpu_variants.sh
bash tools/gpu_step_prof.sh 1200
