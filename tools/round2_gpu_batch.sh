#!/bin/bash
# One gpurun call: GPU tests, compute-sanitizer over every kernel, the synccheck repro, the reference DeepSeek kernel under
# racecheck, ncu captures.  Outputs under gpurun_out/.
set -u
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -8 > $O/round2_pytest_gpu.txt; cat $O/round2_pytest_gpu.txt
for tool in memcheck racecheck synccheck; do
  echo "== $tool" ; timeout 600 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize_targets.py 2>&1 | grep -E "SUMMARY|sanitize targets done|Barrier error|Device Frame|hazard|Invalid|at " | sort | uniq -c | sort -rn | head -14
done > $O/round2_sanitizer.txt 2>&1
{ echo "== synccheck on the minimal repro (64 clusters = 2 waves)"; timeout 300 compute-sanitizer --tool synccheck --print-limit 3 ./tools/synccheck_repro 64 2>&1 | grep -E "SUMMARY|repro|Barrier error|Device Frame" | sort | uniq -c | head;
  echo "== synccheck on the minimal repro (32 clusters = 1 wave)"; timeout 300 compute-sanitizer --tool synccheck --print-limit 3 ./tools/synccheck_repro 32 2>&1 | grep -E "SUMMARY|repro|Barrier error|Device Frame" | sort | uniq -c | head; } >> $O/round2_sanitizer.txt 2>&1
cat $O/round2_sanitizer.txt
{ echo "== reference DeepSeek kernel (oracle/_ref), plain"; timeout 300 python tools/ref_deepseek_debug.py 4 2>&1 | grep -v Warn | tail -6;
  for tool in racecheck initcheck memcheck; do echo "== reference DeepSeek kernel under $tool"; timeout 900 compute-sanitizer --tool $tool --print-limit 6 python tools/ref_deepseek_debug.py 1 2>&1 | grep -E "SUMMARY|run 0|hazard|Uninitialized|Invalid|at .*kernel|Race reported|Device Frame" | sort | uniq -c | sort -rn | head -12; done; } > $O/round2_ref_deepseek.txt 2>&1
cat $O/round2_ref_deepseek.txt
CF_PROFILE=1 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:llama_decoder_layer_kernel -c 3 -o $O/round2_mha_kv16384 python bench.py --steps 1 --warmup 1 --kv-len 16384 --no-sweep --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:llama_decoder_layer_kernel -s 8 -c 3 -o $O/round2_paged_kv16384_random python tools/ncu_targets.py paged 16384 random 2>&1 | tail -1
ncu --set full --clock-control none --import-source on -k regex:llama_decoder_layer_kernel -s 8 -c 3 -o $O/round2_paged_kv16384_sequential python tools/ncu_targets.py paged 16384 sequential 2>&1 | tail -1
ncu --set full --clock-control none --import-source on -k regex:gqa2 -s 8 -c 3 -o $O/round2_gqa8k python tools/ncu_targets.py gqa 8192 2>&1 | tail -1
ncu --set full --clock-control none --import-source on -k regex:ds_ -s 24 -c 9 -o $O/round2_deepseek python tools/ncu_targets.py deepseek 4096 2>&1 | tail -1
CF_PROFILE=1 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/round2_launches.csv python bench.py --steps 2 --warmup 1 --no-sweep --no-cpu-baseline > /dev/null 2>&1
ls -la $O/*.ncu-rep $O/round2_launches.csv
