#!/bin/bash
bash tools/r02_variants.sh r02_var2 lib_head lib_A2 lib_B2
bash tools/r02_phase.sh r02_phase2 lib_A2_timing
