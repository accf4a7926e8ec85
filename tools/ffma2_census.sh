#!/bin/bash
# FFMA2 census of libpsx.so: total packed FMAs and how many take their tap pair from a uniform register
# (2-cycle issue); a drop in the second column after a source change means ptxas fell off the uniform datapath.
cuobjdump -sass "${1:-samplers_b200/_lib/libpsx.so}" | awk '/Function :/{fn=$3} /FFMA2/{t[fn]++; if ($0 ~ /UR[0-9]/) u[fn]++} END{for (f in t) printf "%6d %6d %s\n", t[f], u[f], substr(f,1,70)}' | sort -k3
