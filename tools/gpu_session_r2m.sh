#!/bin/bash
# Round-2 session M: shape 2 at 4 and 3 blocks/SM (126 / 138 registers) against 5 (96).
timeout 300 python tools/ab_atrous.py --workload c2 --frames 20 --shapes "11,12" 2>&1 | cut -c1-330
timeout 300 python tools/ab_atrous.py --workload c4 --frames 10 --shapes "11" 2>&1 | cut -c1-330
