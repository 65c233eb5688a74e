#!/bin/bash
timeout -s INT 600 python -m pytest tests -m gpu -q -x -k "atrous or multirank" 2>&1 | tail -3
for w in c2 c4; do timeout 200 python tools/ab_atrous.py --workload $w --frames 20 --shapes "" 2>&1 | cut -c1-330; done
timeout 200 python tools/ab_atrous.py --workload c4 --frames 20 --shapes "" --strip 945,1215 2>&1 | cut -c1-330
