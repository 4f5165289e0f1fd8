#!/bin/bash
# One-GPU measurement pass of a round: parity tests, bench lines of every BASELINE config, the reference arm,
# the ncu launch list and the ncu full captures the profiles/ summaries are made from.
#   gpurun --timeout 1500 -- 'bash tools/measure_round.sh r2m'
# Everything lands in gpurun_out/<tag>_*; tools/summarize_profiles.py turns it into profiles/.
set -u
TAG=${1:-r2}
O=gpurun_out
mkdir -p $O
(time python -m pytest tests -m gpu -x -q) > $O/${TAG}_pytest.log 2>&1
tail -2 $O/${TAG}_pytest.log
python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference_arm.json 2> $O/${TAG}_bench.err
python bench.py > $O/${TAG}_bench_c3_1gpu.json 2>> $O/${TAG}_bench.err
for w in c1 c2 c4a c4b; do
  python bench.py --workload $w --cpu-seconds 6 > $O/${TAG}_bench_${w}_1gpu.json 2>> $O/${TAG}_bench.err
done
for w in c3o x_ycbcr x_cmyk x_ycck x_rgb411 x_rgb444; do
  python bench.py --workload $w --no-cpu > $O/${TAG}_bench_${w}_1gpu.json 2>> $O/${TAG}_bench.err
done
python bench.py --workload c4a --batch 16 --no-cpu > $O/${TAG}_bench_c4a_x16_1gpu.json 2>> $O/${TAG}_bench.err
python bench.py --workload c5 --no-cpu > $O/${TAG}_bench_c5_1gpu.json 2>> $O/${TAG}_bench.err
python bench.py --workload c5o --no-cpu > $O/${TAG}_bench_c5o_1gpu.json 2>> $O/${TAG}_bench.err
# launch list of the bench command (cold-cache, serialised: shares of the step, not absolute times)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${TAG}_launches_c3_b64.csv \
    python bench.py --workload c3 --batch 64 --steps 2 --warmup 1 --no-cpu --no-c5 > $O/${TAG}_ncu_launches.log 2>&1
# full captures of the dominant kernels (one launch each, after the parity gate and the warm-up)
ncu --set full --clock-control none --import-source on -k "regex:stage_a_warp|encode_chunks|place_chunks|stuff_scatter|count_ff" -s 10 -c 5 -o $O/${TAG}_full \
    python bench.py --workload c3 --batch 64 --steps 2 --warmup 1 --no-cpu --no-c5 > $O/${TAG}_ncu_full.log 2>&1
for w in c4a c4b; do
  ncu --set full --clock-control none -k regex:stage_a_warp -s 2 -c 1 -o $O/${TAG}_stage_a_$w \
      python bench.py --workload $w --steps 2 --warmup 1 --no-cpu > $O/${TAG}_ncu_$w.log 2>&1
done
ncu --set full --clock-control none -k regex:stage_a_warp -s 2 -c 1 -o $O/${TAG}_stage_a_x_rgb444 \
    python bench.py --workload x_rgb444 --batch 64 --steps 2 --warmup 1 --no-cpu --no-c5 > $O/${TAG}_ncu_x_rgb444.log 2>&1
# the progressive coder (C5: blocks staged once for all scans -- dram__bytes_read per block) and the histogram of the optimized batch
ncu --set full --clock-control none -k regex:encode_chunks -s 2 -c 1 -o $O/${TAG}_coder_c5 \
    python bench.py --workload c5 --steps 2 --warmup 1 --no-cpu --no-c5 > $O/${TAG}_ncu_c5.log 2>&1
ncu --set full --clock-control none -k regex:histogram_kernel -s 2 -c 1 -o $O/${TAG}_histogram_c3o \
    python bench.py --workload c3o --batch 64 --steps 2 --warmup 1 --no-cpu --no-c5 > $O/${TAG}_ncu_c3o.log 2>&1
# planar input: the plane kernel against the generic kernel
python tools/planar_stage_time.py > $O/${TAG}_planar_stage_a.txt 2>&1
# gpurun brings back at most 64 MiB: keep the raw metric pages (what tools/summarize_profiles.py reads), not the reports
for r in $O/${TAG}_*.ncu-rep; do
  ncu -i $r --page raw --csv > ${r%.ncu-rep}.raw.csv 2>/dev/null && rm -f $r
done
du -sh $O
cut -c1-300 $O/${TAG}_bench_c3_1gpu.json
