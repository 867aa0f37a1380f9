#!/bin/bash
# Round-2 evidence run (one GPU): smoke x3, GPU suite (default and CTA pairs forced), sanitizer passes over the hand-rolled
# mbarrier / TMEM / cluster pipelines, kernel micro-benchmarks, bench lines, launch list.  Outputs under gpurun_out/r2_*.
set -u
O=gpurun_out
for i in 1 2 3; do timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1 | cut -c1-200; done > $O/r2_smoke.log
timeout 500 python -m pytest tests -m gpu -q 2>&1 | tail -8 > $O/r2_pytest_gpu.log
QTB200_CTA_GROUP=2 timeout 500 python -m pytest tests -m gpu -q 2>&1 | tail -8 > $O/r2_pytest_gpu_cta_group2.log
for tool in racecheck synccheck memcheck; do
  timeout 420 compute-sanitizer --tool $tool --print-limit 5 python -m pytest -q -x -m gpu \
    "tests/test_gpu_fp4.py::test_gemm_f4_accumulators_exact" "tests/test_gpu_requant.py::test_batchnorm_folded_into_epilogue" \
    "tests/test_gpu_convchain.py::test_residual_epilogue_matches_composition" "tests/test_gpu_parity.py::test_config1_linearbin_1024_b512" \
    2>&1 | tail -12 > $O/r2_sanitizer_$tool.log
done
timeout 200 python scratch/kern_bench.py > $O/r2_kern_bench.log 2>&1
timeout 400 python bench.py --steps 20 --warmup 5 2> $O/r2_bench_1gpu.err | tail -1 > $O/r2_bench_1gpu.json
timeout 200 python bench.py --impl reference --steps 5 --warmup 2 2>/dev/null | tail -1 > $O/r2_bench_reference.json
for c in alexnet_w4a4 resnet18_t2a8 vgg_w8a8; do
  timeout 400 python bench.py --config $c --steps 20 --warmup 3 2> $O/r2_${c}_1gpu.err | tail -1 > $O/r2_${c}_1gpu.json
done
QTB200_BENCH_GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 > /dev/null 2>&1
for c in alexnet_w4a4 resnet18_t2a8 vgg_w8a8; do
  QTB200_BENCH_GRAPH=0 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'tc_gemm|image_|pool_|act_quant|rowsum|gemm|elementwise|reduce' -c 400 --csv --log-file $O/r2_launches_$c.csv python bench.py --config $c --steps 1 --warmup 3 > /dev/null 2>&1
done
# full captures: the hot kernels of the headline workloads (profiles/prof_kernels.py) and of one ResNet-18 forward (stem + layer 1)
timeout 500 ncu --set full --clock-control none -k regex:'tc_gemm|act_quant|weight_expand|simt_gemm' -s 19 -c 19 -o /tmp/r2_full_headline python profiles/prof_kernels.py > $O/r2_full_headline.log 2>&1
python profiles/ncu_summary.py /tmp/r2_full_headline.ncu-rep > $O/r2_ncu_full_headline.json 2>/dev/null
timeout 500 ncu --set full --clock-control none -k regex:'tc_gemm|image_windows|pool_quant' -c 7 -o /tmp/r2_full_resnet python bench.py --config resnet18_t2a8 --steps 1 --warmup 3 > $O/r2_full_resnet.log 2>&1
python profiles/ncu_summary.py /tmp/r2_full_resnet.ncu-rep > $O/r2_ncu_full_resnet.json 2>/dev/null
timeout 200 python scratch/r2_ns.py > $O/r2_ns_probe.json 2>/dev/null
timeout 300 python profiles/measure_peaks.py > $O/r2_peaks.log 2>&1
for f in r2_smoke.log r2_pytest_gpu.log r2_pytest_gpu_cta_group2.log r2_sanitizer_racecheck.log r2_sanitizer_synccheck.log r2_sanitizer_memcheck.log; do echo "== $f"; tail -4 $O/$f; done
for f in r2_bench_1gpu r2_alexnet_w4a4_1gpu r2_resnet18_t2a8_1gpu r2_vgg_w8a8_1gpu; do cut -c1-140 $O/$f.json; done
du -sh $O
