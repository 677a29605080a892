#!/bin/bash
# Round-2 profiling batch (run under gpurun, one GPU): in-step timeline, ncu launch list of the bench command, `ncu --set full`
# captures of the attention kernels and of the dominant GEMM launch, compute-sanitizer racecheck / synccheck of the
# mbarrier / TMEM kernels at toy shapes.  Everything lands in gpurun_out/; summaries are copied to profiles/ by hand.
set -u
O=gpurun_out
mkdir -p $O
python tools/timeline_step.py --batch 256 --out $O/r02_timeline_step_cfg2.txt > $O/r02_timeline.log 2>&1
# launch list of the default bench command (cold-cache, serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file $O/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-e2e > $O/r02_ncu_bench.log 2>&1
python tools/summarize_launches.py $O/r02_launches_bench.csv > $O/r02_launches_bench.summary.txt 2>&1
# full captures: attention forward / dQ / dK,dV at the cfg-2 shape, and the FFN-1 GEGLU GEMM (traffic)
ncu --set full --clock-control none --import-source on -k regex:attn_fwd_tc2 -c 1 -o $O/r02_attn_fwd -f python tools/kernel_bench.py attn > $O/r02_ncu_attn_fwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_bwd_dq_tc2 -c 1 -o $O/r02_attn_dq -f python tools/kernel_bench.py attn > $O/r02_ncu_attn_dq.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_bwd_dkv -c 1 -o $O/r02_attn_dkv -f python tools/kernel_bench.py attn > $O/r02_ncu_attn_dkv.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm2_tcgen05_kernel -s 3 -c 1 -o $O/r02_gemm_geglu -f python tools/geglu_one.py > $O/r02_ncu_gemm.log 2>&1
for f in r02_attn_fwd r02_attn_dq r02_attn_dkv r02_gemm_geglu; do python tools/ncu_summary.py $O/$f.ncu-rep > $O/${f}_summary.txt 2>&1; done
# sanitizer: shared-memory races / barrier misuse in the tcgen05 kernels (toy shapes through the kernel tests)
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_kernels_gpu.py -q -x -k "attention_fwd_bwd or gemm_layouts or gemm_geglu" > $O/r02_sanitizer_racecheck.log 2>&1
timeout 900 compute-sanitizer --tool synccheck --print-limit 20 python -m pytest tests/test_kernels_gpu.py -q -x -k "attention_fwd_bwd or gemm_layouts or gemm_geglu" > $O/r02_sanitizer_synccheck.log 2>&1
tail -5 $O/r02_sanitizer_racecheck.log $O/r02_sanitizer_synccheck.log
cat $O/r02_timeline_step_cfg2.txt | head -40
