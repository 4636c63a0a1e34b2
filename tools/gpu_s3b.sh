#!/bin/bash
for n in 64 512; do for lim in 1 2; do for fr in 4 0; do echo "studies $n limit $lim DPHY_SPR_FRONTIER=$fr"; DPHY_SPR_FRONTIER=$fr timeout 300 python tools/spr_bounded_timing.py $lim $n 2>&1 | grep "limit="; done; done; done
