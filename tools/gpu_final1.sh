#!/bin/bash
# round-end evidence run: full -m gpu suite, N=1 bench (+ reference arm), ncu --set full of the dominant kernel,
# ncu launch lists of a 2-layer decode and a 2-layer encoder (same kernels, shapes and launch order as the bench)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_final.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu_final.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_final2_n1.json 2> gpurun_out/bench_final2_n1.err
echo "bench exit $?"; tail -2 gpurun_out/bench_final2_n1.err; cut -c1-1500 gpurun_out/bench_final2_n1.json
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_final2_ref.json 2> gpurun_out/bench_final2_ref.err
echo "ref exit $?"; cut -c1-600 gpurun_out/bench_final2_ref.json
timeout 120 python tools/ncu_decode_attention.py > gpurun_out/da_plain.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:decode_attention_stream -s 1 -c 2 -f -o gpurun_out/prof_decode_attention_final python tools/ncu_decode_attention.py > gpurun_out/da_ncu.log 2>&1
echo "ncu da exit $?"
timeout 120 python tools/ncu_decode_small.py 12 > gpurun_out/ds_plain.log 2>&1 && \
TWB200_GRAPH=0 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/decode_small_final.csv python tools/ncu_decode_small.py 12 > gpurun_out/ds_ncu.log 2>&1
echo "ncu ds exit $?"; wc -l gpurun_out/decode_small_final.csv
timeout 120 python tools/ncu_encoder_small.py 16 > gpurun_out/es_plain.log 2>&1 && \
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/encoder_small_final.csv python tools/ncu_encoder_small.py 16 > gpurun_out/es_ncu.log 2>&1
echo "ncu es exit $?"; wc -l gpurun_out/encoder_small_final.csv
