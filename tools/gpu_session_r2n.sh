#!/bin/bash
# Round-2 session N: programmatic dependent launch in the a-trous chain: parity (a-trous, graph, multirank) and A/B.
timeout -s INT 900 python -m pytest tests -m gpu -q -x -k "atrous or async or multirank or graph" 2>&1 | tail -4
for w in c2 c3; do timeout 200 python tools/ab_atrous.py --workload $w --frames 20 --shapes "" --extra "SVGF_PDL=0" 2>&1 | cut -c1-330; done
timeout 200 python tools/ab_atrous.py --workload c4 --frames 20 --shapes "" --strip 945,1215 --extra "SVGF_PDL=0" 2>&1 | cut -c1-330
