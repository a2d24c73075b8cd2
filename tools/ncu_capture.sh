#!/bin/bash
# ncu evidence of one round (run on the GPU box through gpurun): launch list of the bench command + --set full
# captures of the hot kernels, exported to CSV on the box; the .ncu-rep files stay there (gpurun_out/ is capped at 64 MiB).
#   bash tools/ncu_capture.sh <tag>
tag=${1:-r2}
out=gpurun_out
set -x
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu --no-config5-full > $out/${tag}_plain.log 2>&1 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-config5-full > $out/${tag}_bench_under_ncu.log 2>&1
for part in interp2 interp1 edm; do
  timeout 200 python tools/prof_driver.py $part > $out/${tag}_prof_plain_$part.log 2>&1 || exit 1
  case $part in
    interp2) rx='interp2_scattered_smem|interp2_grid_kernel'; skip=2; cnt=3;;
    interp1) rx='interp1_vec'; skip=0; cnt=8;;
    edm) rx='edm_evolve'; skip=1; cnt=1;;
  esac
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c $cnt -o /tmp/prof_${tag}_$part \
      python tools/prof_driver.py $part > $out/${tag}_prof_ncu_$part.log 2>&1
  ncu -i /tmp/prof_${tag}_$part.ncu-rep --page raw --csv > $out/prof_${tag}_${part}_raw.csv 2>/dev/null
  ncu -i /tmp/prof_${tag}_$part.ncu-rep --page source --csv > $out/prof_${tag}_${part}_src.csv 2>/dev/null
  ls -la /tmp/prof_${tag}_$part.ncu-rep
done
du -sh $out
