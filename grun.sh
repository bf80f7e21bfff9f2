#!/bin/bash
# helper: rebuild the library, then run a command on a B200 (usage: ./grun.sh <timeout_s> '<cmd>')
set -e
cd /root/repo
python gemmgan_b200/build.py > /dev/null
T=$1; shift
/usr/local/graft/bin/gpurun --timeout $T -- "$@"
