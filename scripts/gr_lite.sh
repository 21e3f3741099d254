#!/bin/bash
# gpurun without the staged reference checkpoints (95 MB of the 120 MB push): for kernel experiments that use synthetic
# weights only.  usage: scripts/gr_lite.sh <timeout_s> '<command>'
cd "$(dirname "$0")/.."
cp .gpurunignore .gpurunignore.bak
printf 'baseline/_ref/saved_check_point\nbaseline/_ref/saved_check_point/\nsaved_check_point\n*.ckpt\n' >> .gpurunignore
/usr/local/graft/bin/gpurun --timeout "$1" -- "$2"
rc=$?
mv .gpurunignore.bak .gpurunignore
exit $rc
