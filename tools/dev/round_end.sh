#!/bin/bash
# Round-end evidence on ONE B200 (run through gpurun): bench lines, ncu captures of the shipped kernel, launch list, sanitizers.
# Everything lands in gpurun_out/; summaries are made afterwards on the build host (tools/ncu_summary.py) and copied to profiles/.
tag=${1:-r02_final}
out=gpurun_out
mkdir -p $out
timeout 900 python bench.py --impl reference > $out/bench_${tag}_reference.json 2> $out/bench_${tag}_reference.err
timeout 900 python bench.py > $out/bench_${tag}.json 2> $out/bench_${tag}.err
tail -c 600 $out/bench_${tag}.json
for spec in "batch:prof_batch.py 64 3" "single4k:prof_single.py 3840 2160 3" "single1080nomap:prof_single.py 1920 1080 3 nomap"; do
  name=${spec%%:*}; cmd=${spec#*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:ssim_fused -s 2 -c 1 -o $out/prof_${tag}_${name} -f python tools/dev/$cmd > /dev/null 2>&1
  ncu -i $out/prof_${tag}_${name}.ncu-rep --page raw --csv > $out/prof_${tag}_${name}_raw.csv 2>/dev/null
  ncu -i $out/prof_${tag}_${name}.ncu-rep --page source --csv > $out/prof_${tag}_${name}_src.csv 2>/dev/null
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"ssim_fused_kernel|pack_|deinterleave|scatter_map" -c 300 --csv --log-file $out/launches_${tag}.csv python bench.py --steps 2 --warmup 1 > /dev/null 2>&1
{
  for tool in memcheck racecheck synccheck initcheck; do
    echo "== compute-sanitizer --tool $tool python tools/dev/sanitize.py"
    timeout 900 compute-sanitizer --tool $tool python tools/dev/sanitize.py 2>&1 | tail -12
  done
} > $out/sanitizer_${tag}.txt 2>&1
tail -5 $out/sanitizer_${tag}.txt
ls -la $out | tail -20
