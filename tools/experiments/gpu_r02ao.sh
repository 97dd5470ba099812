#!/bin/bash
for s in "hdri-test 64" "cornell-glossy 64" "quads 64" "cornell 64"; do bash tools/ab_run.sh $s 2>&1 | cut -c1-130; done
