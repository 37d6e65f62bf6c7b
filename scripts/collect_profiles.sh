#!/bin/bash
# Run ON THE GPU BOX (under gpurun): the ncu evidence that profiles/ summarises.  Usage: scripts/collect_profiles.sh <tag>
# Every ncu command is preceded by the same command without ncu (B200_PROFILING.md); numbers printed under ncu are never
# bench values.  Outputs land in gpurun_out/ and are summarised here (CPU box) by scripts/summarize_ncu_launches.py and
# scripts/summarize_ncu_aux.py.
set -u
TAG=${1:-v1}
OUT=gpurun_out
python bench.py --print-digest > $OUT/digest_$TAG.json 2>/dev/null   # identifies the tree these captures were taken on
M_STEP="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"
M_AUX="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed"
for W in cfg2 cfg3 cfg4; do
  python bench.py --workload $W --steps 2 --warmup 3 --no-cpu-baseline > $OUT/plain_$W.log 2>&1 &&
  timeout 900 ncu --nvtx --nvtx-include "sib_timed/" --metrics $M_STEP --clock-control none --csv \
      --log-file $OUT/launches_${W}_$TAG.csv python bench.py --workload $W --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_$W.log 2>&1
done
python scripts/aux_microbench.py > $OUT/aux_plain_$TAG.txt 2>&1 &&
timeout 900 ncu --metrics $M_AUX --clock-control none --csv --log-file $OUT/aux_ncu_$TAG.csv \
    -k regex:'mel_kernel|znorm|layernorm|conv0_gn|conv1d_cout1|pack_int16|si_sdr|abs_diff|resample' \
    python scripts/aux_microbench.py > $OUT/ncu_aux.log 2>&1
if [ "${2:-full}" != "lists" ]; then
for CASE in s4k3d1 s3k7d3 s2k3d5; do
  python scripts/resunit_microbench.py --only $CASE --iters 1 > $OUT/plain_ru_$CASE.log 2>&1 &&
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:resunit -c 2 -o $OUT/prof_resunit_${CASE}_$TAG -f \
      python scripts/resunit_microbench.py --only $CASE --iters 1 > $OUT/ncu_ru_$CASE.log 2>&1
done
for CASE in s2k3c2n ffn1g; do
  python scripts/tc_microbench.py --only $CASE --iters 1 > $OUT/plain_tc_$CASE.log 2>&1 &&
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv1d_bf16 -c 2 -o $OUT/prof_conv_${CASE}_$TAG -f \
      python scripts/tc_microbench.py --only $CASE --iters 1 > $OUT/ncu_tc_$CASE.log 2>&1
done
python scripts/attn_microbench.py --only base_4s --iters 1 > $OUT/plain_attn.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc -c 1 -o $OUT/prof_attn_$TAG -f \
    python scripts/attn_microbench.py --only base_4s --iters 1 > $OUT/ncu_attn.log 2>&1
python scripts/aux_microbench.py --only mel > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mel_kernel -c 1 -o $OUT/prof_mel_$TAG -f \
    python scripts/aux_microbench.py > $OUT/ncu_mel.log 2>&1
fi
ls -la $OUT/*$TAG* | head -30
