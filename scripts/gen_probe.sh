#!/bin/bash
# developer aid: all hazard probes of scripts/gen_probe.py in one GPU call
for d in 0 1 2 4 8 16 32 64 128 256 512; do
  Z3D_GEN_DBG=$d timeout 120 python scripts/gen_probe.py ${1:-64,50,16} 4
done
