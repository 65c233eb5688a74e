#!/bin/bash
# Bisect of the rt_kernel time (0.80 ms at the end of round 1 -> 1.62 ms): the libraries of three older commits, each run through its
# own package + tool (bisect/<commit>/, built by hand, git-ignored), and the A/B builds of the current pathtrace.cu (tools/build_rt_ab.sh).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for c in 9ff4667 beb0b74 85b0eab; do
  (cd bisect/$c && timeout 200 python tools/ab_atrous.py --workload c2 --frames 20 --shapes "" > ../../gpurun_out/bisect_$c.jsonl 2> ../../gpurun_out/bisect_$c.err)
  echo "$c: $(cut -c1-260 gpurun_out/bisect_$c.jsonl)"; tail -2 gpurun_out/bisect_$c.err
done
timeout 200 python tools/ab_atrous.py --workload c2 --frames 20 --shapes "" --extra "SVGF_RT_MINBLOCKS=4" > gpurun_out/bisect_head.jsonl 2> gpurun_out/bisect_head.err; echo "head: $(cut -c1-260 gpurun_out/bisect_head.jsonl)"
for v in nolq nopad nolq_nopad; do
  SVGF_LIB_PATH=$PWD/cuda-path-tracer-denoising_b200/ab/libsvgf_$v.so timeout 200 python tools/ab_atrous.py --workload c2 --frames 20 --shapes "" > gpurun_out/bisect_$v.jsonl 2> gpurun_out/bisect_$v.err
  echo "$v: $(cut -c1-260 gpurun_out/bisect_$v.jsonl)"; tail -2 gpurun_out/bisect_$v.err
done
