#!/bin/bash
# Runs on the GPU box (via gpurun): parity tests, benches and an ncu capture.
# Usage: bash scripts/gpu_check.sh <tag> [steps...]   steps: frontend simt tc quant bench ncu_melif ncu_assign launches
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
tag=$1; shift
steps="$*"
[ -z "$steps" ] && steps="frontend quant tc bench"
for s in $steps; do
  case $s in
    quick)
      # a hang here (e.g. a kernel deadlock) must not burn the GPU budget: stop the whole script
      timeout 150 python -c "
import torch
from interactive_spectrogram_inpainting_b200.utils import synthetic
from interactive_spectrogram_inpainting_b200.utils.spectrograms_helper import MelSpectrogramsHelper
h = MelSpectrogramsHelper().to('cuda:0')
for n in (1, 4, 300):
    s = h.to_spectrogram(synthetic.synthetic_notes(min(n, 8)).repeat((n + 7) // 8, 1)[:n].to('cuda:0'))
    torch.cuda.synchronize(); print('quick ok', n, tuple(s.shape), float(s.abs().mean()))
" 2>&1 | tail -4
      if [ "${PIPESTATUS[0]}" != "0" ]; then echo "quick check FAILED or hung: stopping"; exit 1; fi ;;
    frontend)
      timeout 300 python -m pytest tests/test_gpu_frontend.py -q 2>&1 | tail -40 > gpurun_out/pytest_frontend_$tag.log
      tail -3 gpurun_out/pytest_frontend_$tag.log ;;
    extraction)
      timeout 900 python -m pytest tests/test_gpu_extraction.py -q 2>&1 | tail -40 > gpurun_out/pytest_extraction_$tag.log
      tail -12 gpurun_out/pytest_extraction_$tag.log ;;
    simt)
      timeout 600 python bench.py --steps 10 --warmup 3 --assign-algo simt --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_simt_$tag.json ;;
    tc)
      timeout 300 python -m pytest tests/test_gpu_quantizer_tc.py -x -q -s 2>&1 | tail -60 > gpurun_out/pytest_tc_$tag.log
      tail -15 gpurun_out/pytest_tc_$tag.log ;;
    quant)
      timeout 900 python -m pytest tests/test_gpu_quantizer.py -q 2>&1 | tail -40 > gpurun_out/pytest_quant_$tag.log
      tail -3 gpurun_out/pytest_quant_$tag.log ;;
    bench)
      timeout 900 python bench.py --steps 20 --warmup 3 $BENCH_ARGS 2>&1 | tail -1 > gpurun_out/bench_$tag.json
      cut -c1-400 gpurun_out/bench_$tag.json ;;
    ncu_melif)
      timeout 600 ncu --set full --clock-control none --import-source on -k regex:'^melif_(ws_)?kernel|isi::melif' -s 3 -c 1 \
        -o gpurun_out/melif_$tag python bench.py --steps 2 --warmup 3 --no-cpu-baseline --assign-algo simt > gpurun_out/ncu_melif_$tag.log 2>&1
      tail -1 gpurun_out/ncu_melif_$tag.log | cut -c1-200 ;;
    ncu_assign)
      timeout 600 ncu --set full --clock-control none --import-source on -k regex:vq_assign_tc -s 2 -c 1 \
        -o gpurun_out/assign_$tag python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_assign_$tag.log 2>&1
      tail -1 gpurun_out/ncu_assign_$tag.log | cut -c1-200 ;;
    launches)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv \
        --log-file gpurun_out/launches_$tag.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_$tag.log 2>&1
      tail -1 gpurun_out/ncu_launches_$tag.log | cut -c1-200 ;;
    ncu_stats)
      timeout 600 ncu --set full --clock-control none --import-source on -k regex:vq_gather_stats_smem -s 3 -c 1 \
        -o gpurun_out/stats_$tag python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_stats_$tag.log 2>&1
      tail -1 gpurun_out/ncu_stats_$tag.log | cut -c1-200 ;;
    multi)
      n=${NGPUS:-2}
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 3 2>&1 | tail -2 > gpurun_out/bench_multi${n}_$tag.json
      cut -c1-300 gpurun_out/bench_multi${n}_$tag.json
      timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $n --steps 3 --warmup 1 2>&1 | tail -1 | cut -c1-200 ;;
    sanitize)
      for tool in memcheck racecheck synccheck; do
        for part in frontend quantizer; do
          timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_smoke.py $part > gpurun_out/sanitize_${tool}_${part}_$tag.log 2>&1
          echo "$tool $part rc=$? $(grep -c "=========     at\|========= Error\|========= Warning\|hazard" gpurun_out/sanitize_${tool}_${part}_$tag.log) findings; $(grep "ERROR SUMMARY\|RACECHECK SUMMARY" gpurun_out/sanitize_${tool}_${part}_$tag.log | tail -1)"
        done
      done ;;
    multi_test)
      timeout 600 python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -5 ;;
    inverse)
      timeout 900 python -m pytest tests/test_gpu_inverse.py -q -s 2>&1 | tail -40 > gpurun_out/pytest_inverse_$tag.log
      tail -8 gpurun_out/pytest_inverse_$tag.log ;;
    inverse_bench)
      timeout 300 python tools/bench_inverse.py > gpurun_out/inverse_$tag.json 2> gpurun_out/inverse_$tag.err; tail -3 gpurun_out/inverse_$tag.err; head -c 1500 gpurun_out/inverse_$tag.json ;;
    inverse_sanitize)
      for tool in memcheck racecheck; do
        timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_smoke.py inverse > gpurun_out/sanitize_${tool}_inverse_$tag.log 2>&1
        echo "$tool inverse rc=$? $(grep -c "=========     at\|========= Error\|========= Warning\|hazard" gpurun_out/sanitize_${tool}_inverse_$tag.log) findings; $(grep "ERROR SUMMARY\|RACECHECK SUMMARY" gpurun_out/sanitize_${tool}_inverse_$tag.log | tail -1)"
      done ;;
    ncu_inverse)
      timeout 600 ncu --set full --clock-control none --import-source on -k regex:imelif_kernel -s 4 -c 1 \
        -o gpurun_out/imelif_$tag python tools/bench_inverse.py --only 444 > gpurun_out/ncu_imelif_$tag.log 2>&1
      tail -1 gpurun_out/ncu_imelif_$tag.log | cut -c1-200 ;;
    launches_inverse)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
        --log-file gpurun_out/launches_inverse_$tag.csv python tools/bench_inverse.py --kernel-only > gpurun_out/ncu_launches_inverse_$tag.log 2>&1
      grep -c imelif gpurun_out/launches_inverse_$tag.csv ;;
    projection)
      timeout 240 python -m pytest tests/test_gpu_projection.py -x -q -s 2>&1 | tail -30 > gpurun_out/pytest_projection_$tag.log
      tail -12 gpurun_out/pytest_projection_$tag.log ;;
    ncu_project)
      timeout 600 ncu --set full --clock-control none --import-source on -k regex:vq_project_tc -s 5 -c 1 \
        -o gpurun_out/project_$tag python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_project_$tag.log 2>&1
      tail -1 gpurun_out/ncu_project_$tag.log | cut -c1-200 ;;
    projection_sanitize)
      for tool in memcheck racecheck; do
        timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_smoke.py projection > gpurun_out/sanitize_${tool}_projection_$tag.log 2>&1
        echo "$tool projection rc=$? $(grep -c "=========     at\|========= Error\|========= Warning\|hazard" gpurun_out/sanitize_${tool}_projection_$tag.log) findings; $(grep "ERROR SUMMARY\|RACECHECK SUMMARY" gpurun_out/sanitize_${tool}_projection_$tag.log | tail -1)"
      done ;;
    smoke)
      timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 ;;
    probe)
      timeout 120 ./tools/umma_probe.bin 2>&1 | tee gpurun_out/umma_probe_$tag.log ;;
    dft_probe)
      timeout 90 ./tools/dft_tc_probe.bin 2>&1 | tee gpurun_out/dft_tc_probe_$tag.json ;;
    stress)
      timeout 600 python -m pytest tests/test_gpu_stats_stress.py -q 2>&1 | tail -5 ;;
    ncu_pair)
      timeout 600 ncu --set full --clock-control none --import-source on -k regex:vq_assign_pair_kernel -s 3 -c 1 \
        -o gpurun_out/assign_pair_$tag python tools/ncu_targets.py pair > gpurun_out/ncu_pair_$tag.log 2>&1
      tail -1 gpurun_out/ncu_pair_$tag.log | cut -c1-200 ;;
    ncu_pstream)
      timeout 600 ncu --set full --clock-control none --import-source on -k regex:vq_assign_pstream -s 3 -c 1 \
        -o gpurun_out/assign_pstream_$tag python tools/ncu_targets.py pstream > gpurun_out/ncu_pstream_$tag.log 2>&1
      tail -1 gpurun_out/ncu_pstream_$tag.log | cut -c1-200 ;;
    train_test)
      timeout 300 python -m pytest tests/test_gpu_train_step.py -q 2>&1 | tail -15 ;;
    sanitize_frontend)
      for tool in memcheck synccheck racecheck; do
        timeout 400 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_smoke.py frontend > gpurun_out/sanitize_${tool}_frontend_$tag.log 2>&1
        echo "$tool frontend rc=$? $(grep -c "=========     at\|========= Error\|========= Warning\|hazard" gpurun_out/sanitize_${tool}_frontend_$tag.log) findings; $(grep "ERROR SUMMARY\|RACECHECK SUMMARY" gpurun_out/sanitize_${tool}_frontend_$tag.log | tail -1)"
      done ;;
    fb4)
      ISI_MELIF_WS_FB=4 timeout 300 python -m pytest tests/test_gpu_frontend.py -q 2>&1 | tail -3
      ISI_MELIF_WS_FB=4 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_fb4_$tag.json
      cut -c1-200 gpurun_out/bench_fb4_$tag.json ;;
    time_melif)
      timeout 500 python tools/time_melif.py $TIME_MELIF_CONFIGS 2>&1 | tee gpurun_out/time_melif_$tag.log ;;
    diag)
      timeout 200 python tools/diag_melif.py 2>&1 | tail -40 ;;
    e2e_parity)
      timeout 600 python -m pytest tests/test_gpu_e2e_parity.py -x -q -s 2>&1 | tail -30 > gpurun_out/pytest_e2e_parity_$tag.log
      tail -12 gpurun_out/pytest_e2e_parity_$tag.log ;;
    all_tests)
      timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_all_$tag.log
      tail -5 gpurun_out/pytest_all_$tag.log ;;
  esac
done
