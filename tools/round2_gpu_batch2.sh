#!/bin/bash
set -u
O=gpurun_out
timeout 900 compute-sanitizer --tool memcheck python oracle/gen_golden_deepseek_ref.py $O/deepseek_ref_kernel_seq4096.npz 2>&1 | grep -E "reference kernel|wrote|ERROR SUMMARY"
timeout 900 compute-sanitizer --tool initcheck python oracle/gen_golden_deepseek_ref.py $O/deepseek_ref_kernel_seq4096_initcheck.npz 2>&1 | grep -E "reference kernel|wrote|ERROR SUMMARY"
CF_PROFILE=1 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:llama_decoder_layer_kernel -c 3 -o $O/round2_mha_kv16384 python bench.py --steps 1 --warmup 1 --kv-len 16384 --no-sweep --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:llama_decoder_layer_kernel -s 8 -c 3 -o $O/round2_paged_kv16384_random python tools/ncu_targets.py paged 16384 random 2>&1 | tail -1
ncu --set full --clock-control none --import-source on -k regex:gqa2 -s 8 -c 3 -o $O/round2_gqa8k python tools/ncu_targets.py gqa 8192 2>&1 | tail -1
ncu --set full --clock-control none --cache-control none --import-source on -k regex:llama_decoder_layer_kernel -s 8 -c 3 -o $O/round2_paged_kv16384_random_nocachectl python tools/ncu_targets.py paged 16384 random 2>&1 | tail -1
for r in round2_mha_kv16384 round2_paged_kv16384_random round2_gqa8k round2_paged_kv16384_random_nocachectl; do ncu -i $O/$r.ncu-rep --page raw --csv > $O/$r.raw.csv 2>/dev/null; done
ncu --set full --clock-control none -k regex:ds_ -s 24 -c 9 -o $O/round2_deepseek python tools/ncu_targets.py deepseek 4096 2>&1 | tail -1
ncu -i $O/round2_deepseek.ncu-rep --page raw --csv > $O/round2_deepseek.raw.csv 2>/dev/null; rm -f $O/round2_deepseek.ncu-rep
CF_PROFILE=1 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/round2_launches.csv python bench.py --steps 2 --warmup 1 --no-sweep --no-cpu-baseline > /dev/null 2>&1
ls -la $O | head -30; du -sh $O
