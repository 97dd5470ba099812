#!/bin/bash
for s in "cornell-lucy 64" "random 64" "cornell-smoke 64"; do bash tools/ab_run.sh $s 2>&1 | cut -c1-150; done
