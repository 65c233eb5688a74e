#!/bin/bash
# 2 x 3 patches (shape 13) and 2 x 1 patches (shape 14) against the default 2 x 2
timeout -s INT 600 python -m pytest tests -m gpu -q -x -k "every_tile_shape" 2>&1 | tail -3
timeout 300 python tools/ab_atrous.py --workload c2 --frames 20 --shapes "14" 2>&1 | cut -c1-330
